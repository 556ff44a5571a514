"""profiles/r02_ncu_facts.json from an ncu --set full report of tools/profile_cmd.py: per dominant kernel instance the
per-launch DRAM bytes, tensor-pipe / issue / xu / fma utilisation and duration (last captured launch of each instance).
    python tools/ncu_facts.py gpurun_out/prof_r02.ncu-rep profiles/r02_ncu_facts.json"""
import csv
import json
import subprocess
import sys

import os

out, reps = sys.argv[1], sys.argv[2:]
facts = json.load(open(out)) if os.path.exists(out) else {}
col, units = {}, []


def num(r, name):
    if name not in col or r[col[name]] in ("", "n/a"):
        return None
    v = float(r[col[name]].replace(",", ""))
    u = units[col[name]]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
    return v * scale


import re


def classify(name):
    """kernel instance -> (key used by bench.py, cta_pairs); template arguments DT, MODE, NPG, C8[, PAIR]."""
    flat = re.sub(r"\((int|bool)\)", "", name.replace(" ", ""))
    m = re.search(r"rff_fused_kernel<(\d+),(\d+),(\d+),(\w+)(?:,(\w+))?>", flat)
    if m:
        dt, mode, npg = int(m.group(1)), int(m.group(2)), int(m.group(3))
        c8 = m.group(4) in ("true", "1")
        pair = m.group(5) in ("true", "1")
        shape = {(6, 0): "c3", (2, 0): "c2", (4, 2): "c4d", (4, 0): "c4"}.get((dt, mode))
        if shape is None:
            return None, False
        return f"k1_{shape}_{'fast' if c8 else 'fp32'}", pair
    if re.search(r"mpes_all_perms_kernel<5,3,", flat):
        return "k2_c5", False
    return None, False


# what one captured launch of tools/profile_cmd.py processes (bench.py scales the measured DRAM bytes by its own units / these)
UNITS = {"k1_c3_fp32": {"units_per_launch": 1 << 21, "unit": "candidates"},
         "k1_c3_fast": {"units_per_launch": 1 << 21, "unit": "candidates"},
         "k1_c3_fp32_pass1": {"units_per_launch": 1 << 21, "unit": "candidates"},
         "k1_c3_fast_pass1": {"units_per_launch": 1 << 21, "unit": "candidates"},
         "k1_c2_fp32": {"units_per_launch": 1 << 20, "unit": "candidates"},
         "k1_c4d_fast": {"units_per_launch": 2048 * 256, "unit": "candidates"},
         "k2_c5": {"units_per_launch": 100000, "unit": "query sets"}}
rows_all = []
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    if len(rr) < 3:
        print("no rows in", rep, file=sys.stderr)
        continue
    rows_all.append((rep, rr))
for rep, rows in rows_all:
  hdr, units = rows[0], rows[1]
  col = {h: i for i, h in enumerate(hdr)}
  for r in rows[2:]:
    name = r[col["Kernel Name"]]
    key, pair = classify(name)
    if key is None:
        continue
    if "pass1" in rep:                 # the hi*hi-only pass of the two-pass argmax: same kernel instance, other facts
        key += "_pass1"
    rd, wr = num(r, "dram__bytes_read.sum"), num(r, "dram__bytes_write.sum")
    facts[key] = {
        "kernel": name, "cta_pairs": bool(pair) if key.startswith("k1") else None,
        "source": f"ncu --set full --clock-control none on tools/profile_cmd.py ({rep.split('/')[-1]}), profiles/r02_*_summary.txt",
        "duration_ms": num(r, "gpu__time_duration.sum"),
        "dram_bytes_read": rd, "dram_bytes_write": wr,
        "dram_bytes_per_launch": (rd or 0.0) + (wr or 0.0),
        "tensor_pipe_active_pct": num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": num(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
        "xu_pct": num(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        "fma_pct": num(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "fp64_pct": num(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "lts_throughput_pct": num(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": num(r, "launch__registers_per_thread"),
        "sm_cycles": num(r, "sm__cycles_elapsed.avg")}
    facts[key].update(UNITS.get(key, {}))
with open(out, "w") as fh:
    json.dump(facts, fh, indent=1)
print(json.dumps(facts, indent=1))
