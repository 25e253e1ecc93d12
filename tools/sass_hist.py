#!/usr/bin/env python3
"""sass_hist.py -- opcode histogram per kernel from `cuobjdump -sass` (optionally only between two
labels / within the biggest loop).  Usage: sass_hist.py <binary-or-so> [name-substring]"""
import re, subprocess, sys, collections
def main():
    out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
    pat = sys.argv[2] if len(sys.argv) > 2 else ""
    cur = None; funcs = collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m: cur = m.group(1); funcs[cur] = []; continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m and cur: funcs[cur].append((int(m.group(1), 16), m.group(2)))
    for name, ins in funcs.items():
        if pat not in name: continue
        # biggest backward branch = main loop
        lo, hi = 0, len(ins)
        best = None
        for i, (addr, txt) in enumerate(ins):
            m = re.search(r"BRA\S*\s+(?:!?U?P\d,\s*)?`\(\.L_x_\d+\)|BRA\s.*0x([0-9a-f]+)", txt)
            m2 = re.search(r"0x([0-9a-f]+)", txt) if "BRA" in txt else None
            if m2:
                tgt = int(m2.group(1), 16)
                if tgt < addr and (best is None or addr - tgt > best[1] - best[0]): best = (tgt, addr)
        body = ins
        if best and "--loop" in sys.argv:
            body = [x for x in ins if best[0] <= x[0] <= best[1]]
        h = collections.Counter()
        for _, txt in body:
            t = txt.split()
            op = t[1] if t[0].startswith("@") else t[0]
            h[op] += 1
        print(f"== {name}: {len(body)} instructions" + (f" (loop {best[0]:#x}..{best[1]:#x})" if best and '--loop' in sys.argv else ""))
        for op, c in h.most_common(40): print(f"   {c:6d} {op}")
main()
