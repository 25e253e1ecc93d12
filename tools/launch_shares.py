"""launch_shares.py <ncu launch list csv> -- per-kernel share of the device time of ONE program execute, from an
`ncu --metrics gpu__time_duration.sum --csv` pass over `bench.py --no-graph` (cold-cache, serialised launches: compare
SHARES, not absolutes).  The roofline micro-benchmark launches of the same run (8192-residue grids) are left out."""
import csv, sys, collections, re
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
ix = {h: i for i, h in enumerate(rows[0])}
per = collections.OrderedDict(); execs = 0
for r in rows[1:]:
    name = r[ix["Kernel Name"]]
    if "at::" in name or "elementwise" in name:
        continue
    grid = [int(x) for x in re.findall(r"\d+", r[ix["Grid Size"]])]
    if grid[1] >= 1024:          # the NTT roofline launch (q = 2048 polynomials)
        continue
    short = re.sub(r"\(.*", "", name).replace("void ", "")
    t = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    t *= {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(unit, 1)
    e = per.setdefault(short, [0, 0.0]); e[0] += 1; e[1] += t
    if short.startswith("k_enc_uniform"):
        execs += 1
execs = max(execs, 1)
tot = sum(v[1] for v in per.values())
print("executes in the capture: %d;  device time per execute (serialised, cold): %.1f us;  launches per execute: %.1f" % (execs, tot / execs, sum(v[0] for v in per.values()) / execs))
print("| kernel | launches / execute | us / execute | share |"); print("|---|---|---|---|")
fam = collections.Counter()
for k, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %.1f | %.1f | %.1f %% |" % (k, n / execs, t / execs, 100 * t / tot))
    fam["NTT / iNTT (all variants)" if k.startswith("k_ntt") else k] += t
print(); print("| family | share |"); print("|---|---|")
for k, t in fam.most_common(): print("| %s | %.1f %% |" % (k, 100 * t / tot))
