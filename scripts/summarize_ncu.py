"""Summarise gpurun_out/<tag>_prof.ncu-rep + <tag>_launches.csv into profiles/ (tracked):
   python scripts/summarize_ncu.py r01a [round-label]"""
import csv
import re
import io
import json
import os
import subprocess
import sys

tag = sys.argv[1]
label = sys.argv[2] if len(sys.argv) > 2 else tag
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
reps = [os.path.join(ROOT, "gpurun_out", f"{tag}_prof{x}.ncu-rep") for x in ("", "_primary", "_shadow", "_incoh")]
reps = [r for r in reps if os.path.exists(r)]
rows, hdr, units = [None, None], None, None
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    if hdr is None:
        hdr, units = rr[0], rr[1]
        rows = rr
    else:
        h2_ = rr[0]
        for r in rr[2:]:
            d_ = dict(zip(h2_, r))
            rows.append([d_.get(k, "") for k in hdr])
KEYS = ["gpu__time_duration.sum", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "lts__t_bytes.sum.per_second"]
out = [f"# ncu --set full summary, {label} (bench.py --steps 2 --warmup 3 --profile: C2 ray casting, incoherent rays, then C3 MISPT passes; 1 x B200)\n",
       "Per-launch values; cold-cache, serialised replay: compare shares, not absolutes.\n"]
traffic = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"]
    out.append(f"\n## {name.split('(')[0]}\n\n| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in d:
            out.append(f"| {k} | {d[k]} | {units[hdr.index(k)]} |")
    m_ = re.search(r"k_trace<(?:\(bool\))?(\d)(?:, (?:\(int\))?(\d))?(?:, (?:\(int\))?(\d))?>", name)      # <ANYHIT, TREE1, RAYGEN>
    any_hit, tree1, raygen = (int(m_.group(1)), int(m_.group(2) or 0), int(m_.group(3) or 0)) if m_ else (-1, -1, -1)
    if any_hit == 0 and tree1 == 0 and raygen == 0:
        traffic["incoherent_closest"] = {"ms_under_ncu": float(d["gpu__time_duration.sum"])/1e3 if float(d["gpu__time_duration.sum"]) > 100 else float(d["gpu__time_duration.sum"]),
                                         "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
                                         "active_threads_per_warp_instruction": float(d["smsp__thread_inst_executed_per_inst_executed.ratio"]),
                                         "lsu_wavefronts_pct_of_peak": float(d["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]),
                                         "dram_bytes_per_launch": (float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"]))*1e6,
                                         "warp_instructions": float(d["smsp__inst_executed.sum"])}
    if any_hit == 0 and tree1 == 0 and raygen == 1:
        traffic["k_trace_closest_dram_bytes_per_launch"] = (float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"]))*1e6
        traffic["k_trace_closest_ms_under_ncu"] = float(d["gpu__time_duration.sum"])/1e3
        traffic["issue_active_pct"] = float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"])
        traffic["lsu_wavefronts_pct_of_peak"] = float(d["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"])
        traffic["active_threads_per_warp_instruction"] = float(d["smsp__thread_inst_executed_per_inst_executed.ratio"])
        traffic["l1_hit_pct"] = float(d["l1tex__t_sector_hit_rate.pct"])
        traffic["l2_hit_pct"] = float(d["lts__t_sector_hit_rate.pct"])
        traffic["l2_to_l1_bytes_per_launch"] = float(d["lts__t_sectors_srcunit_tex_op_read.sum"])*32.0
        traffic["l1_load_bytes_per_launch"] = float(d["l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"])*32.0
    if any_hit == 1:
        traffic["k_trace_shadow_dram_bytes_per_launch"] = (float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"]))*1e6
shade = os.path.join(ROOT, "gpurun_out", f"{tag}_shade.ncu-rep")
if os.path.exists(shade):
    raw2 = subprocess.run(["ncu", "-i", shade, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows2 = list(csv.reader(io.StringIO(raw2)))
    h2, u2 = rows2[0], rows2[1]
    for r in rows2[2:]:
        d = dict(zip(h2, r))
        out.append(f"\n## {d['Kernel Name'].split('(')[0]} (C3 pass, one bounce)\n\n| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in d:
                out.append(f"| {k} | {d[k]} | {u2[h2.index(k)]} |")
l2rep = os.path.join(ROOT, "gpurun_out", f"{tag}_l2.ncu-rep")
if os.path.exists(l2rep):
    raw3 = subprocess.run(["ncu", "-i", l2rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows3 = list(csv.reader(io.StringIO(raw3)))
    h3, u3 = rows3[0], rows3[1]
    l2 = ["# ncu record of the L2 / HBM streaming-read microbenchmark (hc_measure_read_bandwidth, k_stream_read), " + label + "\n",
          "First launch: 48 MiB working set, 20 sweeps (L2-resident); second: 2 GiB, one sweep (HBM).  lts__t_bytes / duration is the bandwidth the kernel's own",
          "CUDA-event timing reports as `measured_memory_peaks` in the bench line; `lts__throughput` is ncu's utilisation of the busiest L2 sub-unit.\n"]
    for r in rows3[2:]:
        d = dict(zip(h3, r))
        l2.append(f"\n## {d['Kernel Name'].split('(')[0]} (launch id {d['ID']})\n\n| metric | value | unit |\n|---|---|---|")
        for k in ["gpu__time_duration.sum", "lts__t_bytes.sum", "lts__t_bytes.sum.per_second", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
                  "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                  "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]:
            if k in d:
                l2.append(f"| {k} | {d[k]} | {u3[h3.index(k)]} |")
    open(os.path.join(ROOT, "profiles", f"{label}_l2_microbench_ncu.md"), "w").write("\n".join(l2) + "\n")
for rep_name, kre in ((f"{tag}_prof.ncu-rep", "k_trace"), (f"{tag}_prof_primary.ncu-rep", "k_trace"), (f"{tag}_prof_incoh.ncu-rep", "k_trace"), (f"{tag}_shade.ncu-rep", "k_pt_shade")):
    rp = os.path.join(ROOT, "gpurun_out", rep_name)
    if os.path.exists(rp):
        blk = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_blocks.py"), rp, kre, "1.5"], stdout=subprocess.PIPE, text=True).stdout
        out.append(f"\n## per-basic-block view of {kre} (scripts/ncu_blocks.py: share of issued warp-instructions, average active threads, stall samples)\n\n```\n{blk}```")
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
open(os.path.join(ROOT, "profiles", f"{label}_ncu_full_summary.md"), "w").write("\n".join(out) + "\n")
traffic["source"] = f"profiles/{label}_ncu_full_summary.md (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"
json.dump(traffic, open(os.path.join(ROOT, "profiles", "c2_ncu_traffic.json"), "w"), indent=1)

# launch list: share of each kernel in the step
lc = os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")
tot = {}
n = {}
for r in csv.reader(l for l in open(lc) if not l.startswith("==")):
    if len(r) < 5 or r[0] == "ID":
        continue
    try:
        t = float(r[-1])
    except ValueError:
        continue
    k = r[4].split("(")[0]
    tot[k] = tot.get(k, 0.0) + t
    n[k] = n.get(k, 0) + 1
s = sum(tot.values())
lines = [f"# ncu launch list, {label}: gpu__time_duration.sum per kernel (ns), same command as above\n", "| kernel | launches | total ns | share |", "|---|---|---|---|"]
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    lines.append(f"| {k} | {n[k]} | {v:.0f} | {100*v/s:.1f} % |")
open(os.path.join(ROOT, "profiles", f"{label}_ncu_launches.md"), "w").write("\n".join(lines) + "\n")
import shutil
shutil.copy(lc, os.path.join(ROOT, "profiles", f"{label}_ncu_launches.csv"))
print("\n".join(lines))
print(traffic)
