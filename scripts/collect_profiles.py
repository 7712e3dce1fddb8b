"""Turns the captures of scripts/round2_ncu.sh into the tracked evidence under profiles/:
  profiles/<kernel>_<tag>.txt      ncu_summary of every gpurun_out/prof_*_<tag>.ncu-rep
  profiles/launches_<tag>_step.txt one steady-state training step from the launch list
  profiles/traffic.json            dram bytes per launch of the raster / projection kernels (bench.py's roofline.traffic)
  profiles/sass_summary.txt        cuobjdump -sass mnemonic counts per kernel of libdnr_b200.so (UBLKCP / SYNCS / RED / FFMA2 ...)
usage: python scripts/collect_profiles.py <tag> [--n-gauss 1000000 --width 1920 --height 1080]"""
import argparse
import collections
import csv
import glob
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("tag")
ap.add_argument("--n-gauss", type=int, default=1_000_000)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
args = ap.parse_args()
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)

traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"prof_*_{args.tag}.ncu-rep"))):
    name = os.path.basename(rep)[len("prof_"):-len(".ncu-rep")]
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    open(os.path.join(ROOT, "profiles", f"{name}.txt"), "w").write(out)
    rd = re.search(r"dram__bytes_read.sum\s+([\d.]+) (\w+)", out)
    wr = re.search(r"dram__bytes_write.sum\s+([\d.]+) (\w+)", out)
    unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    if rd and wr:
        b = float(rd.group(1)) * unit[rd.group(2)] + float(wr.group(1)) * unit[wr.group(2)]
        k = name[: -len("_" + args.tag)]
        traffic[f"{k}:{args.n_gauss}:{args.width}x{args.height}:n"] = {
            "dram_bytes_per_launch": int(b), "source": f"profiles/{name}.txt ({rd.group(1)} {rd.group(2)} read + {wr.group(1)} {wr.group(2)} written)"}
    print("summarised", name)
traffic["_comment"] = ("DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from ncu --set full captures of bench.py's "
                       "workload; key = kernel:N:WxH:(n)ormals|(c)olour-only")
json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)

lst = os.path.join(ROOT, "gpurun_out", f"launches_{args.tag}.csv")
if os.path.exists(lst):
    rows = list(csv.reader(open(lst)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    cols = rows[hdr]
    data = [r for r in rows[hdr + 1:] if len(r) == len(cols)]
    ki, vi = cols.index("Kernel Name"), cols.index("Metric Value")
    names = [r[ki] for r in data]
    vals = [float(r[vi].replace(",", "")) for r in data]
    idx = [i for i, n in enumerate(names) if "adam" in n]
    if len(idx) >= 2:
        a, b = idx[-2] + 1, idx[-1] + 1
        tot = sum(vals[a:b])
        lines = [f"{v / 1000:9.1f} us  {100 * v / tot:5.1f}%  {n[:120]}" for n, v in zip(names[a:b], vals[a:b])]
        open(os.path.join(ROOT, "profiles", f"launches_{args.tag}_step.txt"), "w").write(
            "# one steady-state training step (eager launches, --no-graph, 24-view ring), ncu --metrics gpu__time_duration.sum "
            "--clock-control none\n# per-launch times are cold-cache and serialised: compare SHARES; source: gpurun_out/"
            f"launches_{args.tag}.csv\n" + "\n".join(lines) + f"\nstep total {tot / 1000:.1f} us, {b - a} launches\n")
        print("step total us", tot / 1000, "launches", b - a)

# SASS evidence: which Blackwell-specific instructions each kernel really contains
lib = os.path.join(ROOT, "dn_splatter_b200", "libdnr_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
want = ("UBLKCP", "SYNCS", "REDG", "RED", "ATOMG", "FFMA2", "FMUL2", "FADD2", "MUFU", "LDS", "STS", "SHFL", "LDG", "STG", "BAR", "FFMA")
per = collections.OrderedDict()
fn = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(anonymous namespace\)::", "", fn)[:90]
        per[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        per[fn][m.group(1)] += 1
with open(os.path.join(ROOT, "profiles", "sass_summary.txt"), "w") as f:
    f.write("# cuobjdump -sass dn_splatter_b200/libdnr_b200.so: instruction counts per kernel (sm_100a cubins only).\n"
            "# UBLKCP = cp.async.bulk (TMA engine), SYNCS = mbarrier ops, REDG/RED = reduction atomics, FFMA2/FMUL2/FADD2 = packed f32x2\n")
    for fn, c in per.items():
        if not any(k in fn for k in ("dnr", "raster", "project", "emit", "count", "offsets", "pad", "finalize", "loss", "ssim", "adam",
                                      "knn", "density", "l1_", "u8_to", "peer", "normal_from", "photometric", "iota")) or "cub::" in fn:
            continue
        f.write(f"{fn}\n    total {sum(c.values())}  " + "  ".join(f"{k} {c[k]}" for k in want if c[k]) + "\n")
print("wrote profiles/sass_summary.txt")
