"""Condense an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log of
tools/ncu_target.py into the launch list of the LAST forward (one row per launch: kernel, grid, time,
DRAM bytes) and per-kernel-class totals.
    python tools/ncu_launch_summary.py gpurun_out/launches_fp32_v7.csv profiles/r01_ncu_launches_fp32_v7.csv"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
ix = {h: i for i, h in enumerate(rows[hi])}
launches = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= ix["Metric Value"]:
        continue
    d = launches.setdefault(int(r[ix["ID"]]), {"kernel": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]]})
    val = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    name = r[ix["Metric Name"]]
    if name == "gpu__time_duration.sum":
        val *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)   # -> us
    else:
        val *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    d[name] = val
ids = sorted(launches)
# last forward = from the last pack_mel launch on
start = max(i for i in ids if "pack_mel" in launches[i]["kernel"])
fwd = [launches[i] for i in ids if i >= start]
def short(k):
    k = re.sub(r"^void ", "", k)
    k = re.sub(r"\(.*$", "", k)
    return k.replace("bvg::", "")
out = csv.writer(open(sys.argv[2], "w", newline=""))
out.writerow(["launch", "kernel", "grid", "block", "time_us", "dram_read_bytes", "dram_write_bytes"])
cls = collections.OrderedDict()
for n, d in enumerate(fwd):
    t, rd, wr = d.get("gpu__time_duration.sum", 0.0), d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
    out.writerow([n, short(d["kernel"]), d["grid"], d["block"], f"{t:.2f}", int(rd), int(wr)])
    key = re.sub(r"<.*", "", short(d["kernel"]))
    c = cls.setdefault(key, [0, 0.0, 0.0, 0.0])
    c[0] += 1; c[1] += t; c[2] += rd; c[3] += wr
tot = sum(c[1] for c in cls.values())
print(f"# last forward: {len(fwd)} launches, {tot/1e3:.2f} ms summed launch durations (cold-cache, serialised)")
print("| kernel | launches | sum of durations | share | DRAM read | DRAM write | mean traffic / launch |")
print("|---|---|---|---|---|---|---|")
for k, c in cls.items():
    print(f"| {k} | {c[0]} | {c[1]/1e3:.2f} ms | {100*c[1]/tot:.1f} % | {c[2]/1e9:.2f} GB | {c[3]/1e9:.2f} GB | {(c[2]+c[3])/c[0]/1e6:.1f} MB |")
