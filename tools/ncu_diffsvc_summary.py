"""Condense the ncu launch list of tools/ncu_diffsvc_target.py (gpu__time_duration.sum + DRAM bytes per launch) into the
launches of the LAST sampler step (from the last diffembed_kernel on) and per-kernel totals.
    python tools/ncu_diffsvc_summary.py gpurun_out/launches_diffsvc_fp32.csv profiles/r02_ncu_launches_diffsvc_fp32.csv"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
ix = {h: i for i, h in enumerate(rows[hi])}
launches = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= ix["Metric Value"]:
        continue
    d = launches.setdefault(int(r[ix["ID"]]), {"kernel": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]]})
    val = float(r[ix["Metric Value"]].replace(",", ""))
    unit, name = r[ix["Metric Unit"]], r[ix["Metric Name"]]
    if name == "gpu__time_duration.sum":
        val *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
    else:
        val *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    d[name] = val
ids = sorted(launches)
start = max(i for i in ids if "diffembed" in launches[i]["kernel"])
step = [launches[i] for i in ids if i >= start and "bvg::" in launches[i]["kernel"]]
short = lambda k: re.sub(r"<.*$", "", re.sub(r"\(.*$", "", re.sub(r"^void ", "", k))).replace("bvg::", "")
out = csv.writer(open(sys.argv[2], "w", newline=""))
out.writerow(["#", "kernel", "grid", "block", "time_us", "dram_read_MB", "dram_write_MB"])
tot = collections.OrderedDict()
for n, d in enumerate(step):
    t, rd, wr = d.get("gpu__time_duration.sum", 0.0), d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
    out.writerow([n, short(d["kernel"]), d["grid"], d["block"], f"{t:.2f}", f"{rd / 1e6:.3f}", f"{wr / 1e6:.3f}"])
    a = tot.setdefault(short(d["kernel"]), [0, 0.0])
    a[0] += 1
    a[1] += t
print(f"{len(step)} launches, {sum(v[1] for v in tot.values()):.1f} us summed")
for k, (n, t) in tot.items():
    print(f"  {k:22s} n={n:3d} {t:8.1f} us  ({t / n:.2f} us each)")
