#!/usr/bin/env python3
"""Turn the raw gpurun_out/ artefacts of a profiling call into the committed summaries under profiles/.
   python tools/summarize_profiles.py   (needs ncu on PATH to read the .ncu-rep)"""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
# round tag and the gpurun_out artefacts of that round: bench line, reference-arm line, launch list, full capture
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
SRC = {"r01": ("bench_r01.json", "bench_r01_ref.json", "launches_r01.csv", "prof_r01_final.ncu-rep"),
       "r02": ("r2z_bench.json", "r2z_bench_ref.json", "r02_launches.csv", "r02_full.ncu-rep")}[R]
shutil.copy(os.path.join(G, SRC[0]), os.path.join(P, f"{R}_bench_final.json"))
shutil.copy(os.path.join(G, SRC[1]), os.path.join(P, f"{R}_bench_reference_arm.json"))
# ---- launch list
lines = [l for l in open(os.path.join(G, SRC[2])) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    a = agg.setdefault(row["Kernel Name"], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
out = [f"# ncu launch list, {R} (bench.py --steps 1 --warmup 3 ... ; first launches of the run, see the gpurun command in DESIGN.md;",
       "# --metrics gpu__time_duration.sum --clock-control none).  Per-launch times are cold-cache and serialised: compare SHARES.",
       "kernel,launches,total_us,share"]
conv = 0.0
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"\"{k[:100]}\",{n},{t:.1f},{t / tot:.4f}")
    if "conv_" in k or "stem_fused" in k or "block_rows" in k:
        conv += t
out.append(f"# share of conv_umma + conv_umma2 + block_rows + stem_fused kernels: {conv / tot:.4f}")
open(os.path.join(P, f"{R}_launches_final.csv"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:10])); print(out[-1])
# ---- full capture
raw = subprocess.run(["ncu", "-i", os.path.join(G, SRC[3]), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, units, data = rows[0], rows[1], rows[2:]
g = lambda d, n: d[hdr.index(n)]
tb = lambda v, u: float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
out = [f"# ncu --set full --clock-control none, {R} kernels; one chunk (148 segments x 6 heads in r02, 128 in r01) of",
       "# tools/one_chunk.py (r02) / bench.py --batch 256 (r01).  Replayed, cold-cache, unthrottled clocks.",
       "kernel,grid,duration_us,dram_read_MB,dram_write_MB,dram_pct_of_peak,tensor_pipe_pct,sm_throughput_pct,regs,smem_wavefronts,smem_bank_conflicts"]
cb = []
for d in data:
    name = g(d, "Kernel Name").split("(")[0].replace("sad::<unnamed>::", "").replace("void ", "")
    rd, wr = tb(g(d, "dram__bytes_read.sum"), ur), tb(g(d, "dram__bytes_write.sum"), uw)
    out.append(",".join([name, g(d, "Grid Size").replace(",", " "), g(d, "gpu__time_duration.sum"), f"{rd / 1e6:.1f}", f"{wr / 1e6:.1f}",
                         g(d, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")[:5],
                         g(d, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")[:5],
                         g(d, "sm__throughput.avg.pct_of_peak_sustained_elapsed")[:5], g(d, "launch__registers_per_thread"),
                         g(d, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), g(d, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")]))
    if "conv_" in name or "stem_fused" in name or "block_rows" in name:
        cb.append(rd + wr)
open(os.path.join(P, f"{R}_ncu_full_summary.csv"), "w").write("\n".join(out) + "\n")
print("\n".join(out[3:]))
json.dump({"kernel": "conv_umma_kernel / conv_umma2_kernel / block_rows_kernel / stem_fused_kernel (launches of one chunk of 128 segments x 6 heads)",
           "dram_bytes_per_launch_mean": sum(cb) / len(cb), "launches": len(cb),
           "source": f"profiles/{R}_ncu_full_summary.csv (dram__bytes_read.sum + dram__bytes_write.sum)", "chunk": 128 if R == "r01" else 148, "heads": 6},
          open(os.path.join(P, f"{R}_roofline_traffic.json"), "w"), indent=1)
d = json.load(open(os.path.join(G, SRC[0])))
print({k: d[k] for k in ["value", "ms_per_step", "gpu_launches", "clocks", "e2e", "cpu_baseline"]})
print(d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["share_of_step"], d["roofline_frontend"])
