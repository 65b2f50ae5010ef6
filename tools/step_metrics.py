"""Summarise an ncu per-launch metrics csv of one step (tools/gpu_profile.sh) per kernel family:
time, DRAM bytes, tensor-pipe activity.  Writes a markdown table and (optionally) the traffic json bench.py reads."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))  # repo root
import collections
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ki, ni, vi, ui, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
launch = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("jb::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    d = launch.setdefault(r[idi], {"name": name})
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    scale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    d[r[ni]] = v * scale


def family(n):
    if n.startswith("mrf_pair_kernel") or n.startswith("conv_bf16_tma_kernel"):
        return "hifigan_conv"
    if n.startswith("gemm_split"):
        return "fs2_split_gemm"
    if "attention" in n:
        return "fs2_attention"
    return "other"


fam = collections.OrderedDict()
per_kernel = collections.OrderedDict()
for d in launch.values():
    for key, table in ((family(d["name"]), fam), (d["name"], per_kernel)):
        a = table.setdefault(key, {"n": 0, "t": 0.0, "rd": 0.0, "wr": 0.0, "tensor_w": 0.0})
        t = d.get("gpu__time_duration.sum", 0.0)
        a["n"] += 1
        a["t"] += t
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
        a["tensor_w"] += t * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
tot = sum(a["t"] for a in fam.values())
print("| kernel | launches | ms | share | DRAM read MB | DRAM write MB | tensor pipe active (time-weighted) |")
print("|---|---|---|---|---|---|---|")
for table in (fam, per_kernel):
    for k, a in sorted(table.items(), key=lambda x: -x[1]["t"]):
        if a["t"] < 0.05e-3:
            continue
        print(f"| {k} | {a['n']} | {a['t'] * 1e3:.3f} | {100 * a['t'] / tot:.1f} % | {a['rd'] / 1e6:.0f} | {a['wr'] / 1e6:.0f} | {a['tensor_w'] / max(a['t'], 1e-12):.1f} % |")
    print("| | | | | | | |")
print(f"\ntotal {tot * 1e3:.3f} ms, {len(launch)} launches")
if len(sys.argv) > 2:
    a = fam["hifigan_conv"]
    json.dump({"hifigan_conv_dram_bytes_per_step": a["rd"] + a["wr"], "launches": a["n"], "ms_under_ncu": a["t"] * 1e3,
               "tensor_pipe_active_pct_time_weighted": a["tensor_w"] / a["t"],
               "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over every HiFi-GAN convolution launch of one "
                         "bench step (tools/gpu_profile.sh -> " + sys.argv[1].split("/")[-1] + ")"}, open(sys.argv[2], "w"), indent=1)
