"""Per-kernel SASS mnemonic counts of the built library (evidence that the hot kernels are tcgen05 / TMEM / TMA code):
    python tools/sass_counts.py [path/to/libbvg_b200.so] > profiles/rNN_sass_counts.md
Runs `cuobjdump -sass`, demangles the kernel names with c++filt and counts the instruction families that matter:
UTCHMMA (tcgen05.mma), UTMALDG (TMA tensor load), LDTM (tcgen05.ld from TMEM), UTCBAR (tcgen05.commit), SYNCS (mbarrier),
HMMA (mma.sync), FFMA2 / FMUL2 / FADD2 (two-lane fp32), FFMA, MUFU, LDGSTS (cp.async), LDSM / STSM (ldmatrix / stmatrix)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "svc_inference_pipeline_b200", "libbvg_b200.so")
FAMILIES = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS", "HMMA", "FFMA2", "FMUL2", "FADD2", "FFMA", "MUFU", "LDGSTS", "LDSM", "STSM", "LDG", "STG"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
counts, total, name = collections.OrderedDict(), {}, None
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        name = m.group(1)
        counts[name] = collections.Counter()
        total[name] = 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and name:
        total[name] += 1
        full = m.group(1)
        op = full.split(".")[0]
        if full.startswith("UTCHMMA") and ".2CTA" in full:
            counts[name]["UTCHMMA.2CTA"] += 1
            continue
        for fam in FAMILIES:
            if op == fam or (op.startswith(fam) and fam not in ("FFMA", "LDG", "STG")) or (fam in ("FFMA", "LDG", "STG") and op == fam):
                counts[name][fam] += 1
                break
names = list(counts)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
short = {}
for n, d in zip(names, dem):
    d = re.sub(r"\(.*\)$", "", d).replace("void bvg::", "").replace("bvg::", "")
    d = d.replace("(bvg_dtype)", "").replace("(bool)", "")
    short[n] = d
print(f"# SASS mnemonic counts per kernel: `cuobjdump -sass {os.path.relpath(lib, ROOT)}` (sm_100a), `tools/sass_counts.py`\n")
print("| kernel | instr | " + " | ".join(FAMILIES) + " |")
print("|---|---|" + "---|" * len(FAMILIES))
agg = collections.OrderedDict()
for n in names:
    base = short[n].split("<")[0]
    a = agg.setdefault(base, [0, 0, collections.Counter()])
    a[0] += 1
    a[1] += total[n]
    a[2].update(counts[n])
for base, (k, tot, c) in agg.items():
    print(f"| {base} ({k} instantiation{'s' if k > 1 else ''}) | {tot} | " + " | ".join(str(c[f]) if c[f] else "" for f in FAMILIES) + " |")
print("\nLargest instantiations of the two hot kernels:\n")
print("| kernel | instr | " + " | ".join(FAMILIES) + " |")
print("|---|---|" + "---|" * len(FAMILIES))
for n in sorted(names, key=lambda q: -total[q]):
    if any(s in short[n] for s in ("conv_pair_kernel", "conv_umma_kernel", "amp_kernel_p2", "amp_mma_kernel")):
        print(f"| {short[n]} | {total[n]} | " + " | ".join(str(counts[n][f]) if counts[n][f] else "" for f in FAMILIES) + " |")
