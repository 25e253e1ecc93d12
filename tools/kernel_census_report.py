"""kernel_census_report.py <ncu csv> <census stdout> -> markdown table: per kernel launch of tools/kernel_census.py the duration,
DRAM read / written, L2 sectors, fma-heavy / ALU pipe utilisation, issue utilisation; per op the measured DRAM traffic against the
algorithmic bytes and the fraction of the measured HBM rate."""
import csv, json, sys, collections, os
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_sectors_lookup_hit.sum",
           "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread"]
if __name__ == "__main__":
    if len(sys.argv) == 2 and sys.argv[1] == "--metrics":
        print(",".join(METRICS)); sys.exit(0)
    rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
    hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
    launches = collections.OrderedDict()
    for r in rows[1:]:
        lid = r[ix["ID"]]
        e = launches.setdefault(lid, {"name": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]]})
        v = r[ix["Metric Value"]].replace(",", "")
        unit = r[ix["Metric Unit"]]
        try:
            x = float(v)
        except ValueError:
            continue
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "second": 1}.get(unit, 1)
        e[r[ix["Metric Name"]]] = x * scale
    ops = [json.loads(l) for l in open(sys.argv[2]) if l.startswith("{")]
    peak = 6486.5
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    L = [k for k in launches.values() if "at::" not in k["name"]]     # drop torch's L2-flush fills between the ops
    pos = 0
    print("| op (64 instances per launch unless noted) | kernel | grid x block | regs | µs | DRAM rd+wr MB | L2 sectors M (hit %) | fmaheavy % | alu % | issue % |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    summary = []
    for op in ops:
        ks = L[pos:pos + op["launches"]]; pos += op["launches"]
        t = sum(k.get("gpu__time_duration.sum", 0) for k in ks)
        dram = sum(k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0) for k in ks)
        algo = op["algorithmic_bytes_per_instance"] * op["instances"]
        summary.append((op["op"], len(ks), t, dram, algo))
        for k in ks:
            name = k["name"].split("(")[0].replace("void ", "")
            sec = k.get("lts__t_sectors.sum", 0)
            hit = 100 * k.get("lts__t_sectors_lookup_hit.sum", 0) / sec if sec else 0
            print("| %s | `%s` | %s x %s | %d | %.1f | %.1f | %.2f (%.0f) | %.1f | %.1f | %.1f |" % (
                op["op"], name, k["grid"].replace(" ", ""), k["block"].replace(" ", ""), k.get("launch__registers_per_thread", 0), k.get("gpu__time_duration.sum", 0) * 1e6,
                (k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0)) / 1e6, sec / 1e6, hit,
                k.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", 0), k.get("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", 0),
                k.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0)))
    print()
    print("| op | launches | µs per launch set | µs per instance | algorithmic MB | measured DRAM MB | DRAM / algorithmic | algorithmic GB/s | of measured HBM (%.1f GB/s) |" % peak)
    print("|---|---|---|---|---|---|---|---|---|")
    for name, n, t, dram, algo in summary:
        inst = next(o["instances"] for o in ops if o["op"] == name)
        print("| %s | %d | %.1f | %.2f | %.1f | %.1f | %.2f | %.0f | %.1f %% |" % (name, n, t * 1e6, t * 1e6 / inst, algo / 1e6, dram / 1e6, dram / algo if algo else 0, algo / t / 1e9 if t else 0, 100 * algo / t / 1e9 / peak if t else 0))
