// logging.hpp -- EVA_VERBOSITY (reference eva/util/logging.cpp:12-70): silent / info / debug / trace
#pragma once
#include <algorithm>
#include <cstdlib>
#include <string>

namespace evab {

// ---- logging (reference eva/util/logging.cpp:12-70: EVA_VERBOSITY) ----
inline int verbosity() {
  static int v = -1;
  if (v < 0) {
    v = 0;
    if (const char *e = std::getenv("EVA_VERBOSITY")) {
      std::string s(e);
      for (auto &c : s) c = (char)std::tolower(c);
      if (s == "silent") v = 0; else if (s == "info") v = 1; else if (s == "debug") v = 2; else if (s == "trace") v = 3;
      else v = std::max(0, std::atoi(e));
    }
  }
  return v;
}

}  // namespace evab
