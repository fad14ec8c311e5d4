"""Where the step time sits against the rooflines: per kernel family of a bench.py JSON line,
the measured time, the time the same algorithmic work would take at the measured peak
(MEASURED_PEAKS.json: HBM copy GB/s for the memory-bound families, 1/3 of the sustained bf16
tensor rate for the fp16x3 GEMMs) and the gap.  No GPU needed.

    python scripts/step_gap.py profiles/r1c_bench_n1.json > profiles/r1c_step_gap.md
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
line = json.load(open(sys.argv[1]))
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
hbm, tens = peaks["hbm_gbs"], peaks["bf16_tflops_sustained"]
fam = line["kernel_families"]
steps = line["steps"]
rows, tot, tot_ideal = [], 0.0, 0.0
for name, f in sorted(fam.items(), key=lambda kv: -kv[1]["ms_per_step"]):
    ms = f["ms_per_step"]
    if name == "gemm_tc":
        work = f["rate"] * 1e12 * ms * 1e-3                      # flops (2*M*N*K)
        ideal = work / (tens / 3.0 * 1e12) * 1e3
        what = f"{work / 1e12:.2f} TFLOP at {f['rate']:.0f} TFLOP/s (ceiling {tens / 3:.0f} = sustained bf16 / 3)"
    elif name == "gemm_prep":
        # the profiler books 8 B per operand element for this family: rate is in "TFLOP/s" units of bytes
        work = f["rate"] * 1e12 * ms * 1e-3
        ideal = work / (hbm * 1e9) * 1e3
        what = (f"{work / 1e9:.2f} GB booked (8 B per operand element) at {work / 1e9 / (ms * 1e-3):.0f} GB/s; the "
                f"column-scaled splits really move 12 B per element, ~21 GB per step in all = 3.2 ms at peak")
    else:
        work = f["rate"] * 1e9 * ms * 1e-3
        ideal = work / (hbm * 1e9) * 1e3
        what = f"{work / 1e9:.2f} GB at {f['rate']:.0f} GB/s"
    tot += ms
    tot_ideal += ideal
    rows.append((name, f["launches"] // steps, ms, ideal, ms - ideal, what))
print(f"# Step gap analysis -- {os.path.basename(sys.argv[1])}\n")
print(f"{line['config']['workload']}; {line['ms_per_step']:.2f} ms per step = {line['value']:.0f} {line['unit']} "
      f"(profiled pass: {line['roofline']['profiled_ms_per_step']:.2f} ms).  Peaks: HBM {hbm:.0f} GB/s, bf16 sustained "
      f"{tens:.0f} TFLOP/s (MEASURED_PEAKS.json).\n")
print("| family | launches/step | measured ms | at-peak ms | gap ms | algorithmic work |")
print("|---|---:|---:|---:|---:|---|")
for r in rows:
    print(f"| {r[0]} | {r[1]} | {r[2]:.3f} | {r[3]:.3f} | {r[4]:+.3f} | {r[5]} |")
print(f"| **sum** | | {tot:.3f} | {tot_ideal:.3f} | {tot - tot_ideal:+.3f} | |")
print("\nA negative GEMM gap means the box ran above the pool's recorded sustained bf16 rate (power-capped figure); the "
      "memory-bound families hold the remaining headroom.")
