"""Summarise an `ncu --set full` capture: per launch duration, DRAM bytes, tensor-pipe / L2 / issue utilisation; writes the
text summary and merges `dram_bytes_per_launch` of the kernel class into profiles/r2_ncu_traffic.json (read by bench.py).

    ncu -i capture.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_summary.py raw.csv <config> <op name in bench.py> <out.txt>
"""
import csv, json, os, sys

raw, config, op, out = sys.argv[1:5]
rows = list(csv.reader(open(raw)))
hdr, data = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB rd"), ("dram__bytes_write.sum", "MB wr"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %")]
lines = [f"{raw}: {len(data)} launches ({config}, op {op}); units as reported by ncu (duration us, DRAM Mbyte)"]
lines.append("%-44s %-14s " % ("kernel", "grid") + " ".join("%9s" % c[1] for c in cols))
tot_b = tot_t = 0.0
for r in data:
    vals = [float(r[ix[c]].replace(",", "")) if c in ix and r[ix[c]] not in ("", "n/a") else float("nan") for c, _ in cols]
    tot_t += vals[0]
    tot_b += (vals[1] + vals[2]) * 1e6
    lines.append("%-44s %-14s " % (r[ix["Kernel Name"]].replace("void <unnamed>::", "")[:44], r[ix["Grid Size"]]) + " ".join("%9.2f" % v for v in vals))
n = len(data)
lines.append(f"mean per launch: {tot_t / n:.2f} us, DRAM read+write {tot_b / n / 1e6:.2f} MB")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_ncu_traffic.json")
d = json.load(open(p)) if os.path.exists(p) else {}
d.setdefault(config, {})[op] = dict(dram_bytes_per_launch=tot_b / n, launches_captured=n, mean_us=tot_t / n, source=os.path.basename(out))
json.dump(d, open(p, "w"), indent=1)
