#!/usr/bin/env python
"""Turn the artefacts of tools/gpu_round.sh (gpurun_out/<tag>/) into the tracked summaries under profiles/<tag>/:
launch-list shares, ncu full-capture counters of the fused decode kernel, per-line stall summary, DRAM traffic.
    python tools/summarize_round.py r01c
"""
import collections, csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
src = os.path.join(ROOT, "gpurun_out", tag)
dst = os.path.join(ROOT, "profiles", tag)
os.makedirs(dst, exist_ok=True)
for f in ("bench.json", "timeline.txt", "timeline.json", "gpu.txt"):
    if os.path.exists(os.path.join(src, f)):
        shutil.copy(os.path.join(src, f), dst)

# ---- launch list
rows = list(csv.reader(open(os.path.join(src, "launches.csv"))))
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, start = r, i + 1
        break
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[start:]:
    try:
        v = float(r[mv].replace(",", ""))
    except (ValueError, IndexError):
        continue
    k = r[kn].split("(")[0]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
with open(os.path.join(dst, "launch_list_summary.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 700 python bench.py --steps 2 --warmup 3 --no-cpu\n")
    f.write(f"(cold-cache, serialised per-launch times: compare SHARES) total {tot/1e3:.1f} us over {sum(v[0] for v in agg.values())} launches\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{100*v[1]/tot:6.2f}%  n={v[0]:4d}  {v[1]/1e3:10.1f} us  avg {v[1]/v[0]/1e3:8.1f} us  {k[:100]}\n")

# ---- full capture
rep = os.path.join(src, "decode_mega.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr = rr[0]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]
out = ["ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 4 -c 2   (python bench.py --steps 2 --warmup 3 --no-cpu)",
       "kernel: gv::decode_mega_kernel<8,false>: 148 CTAs x 288 threads (8 consumer warps + producer warp), 8 tokens per launch", ""]
mul = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
rd = wr = None
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        vals = [r[i] for r in rr[2:]]
        out.append(f"{k:78s} {rr[1][i]:16s} " + "  ".join(vals))
        if k == "dram__bytes_read.sum":
            rd = [float(v) * mul[rr[1][i]] for v in vals]
        if k == "dram__bytes_write.sum":
            wr = [float(v) * mul[rr[1][i]] for v in vals]
traffic = sum(a + b for a, b in zip(rd, wr)) / len(rd)
out += ["", f"algorithmic bytes per 8-forward launch: 8 x 1.516 GB + KV (S~60) = 12.13 GB; measured DRAM traffic {traffic/1e9:.3f} GB -> no re-reads"]
open(os.path.join(dst, "decode_mega_ncu_summary.txt"), "w").write("\n".join(out) + "\n")
json.dump({"dram_bytes_per_launch": traffic,
           "source": f"profiles/{tag}/decode_mega_ncu_summary.txt (dram__bytes_read.sum + dram__bytes_write.sum, mean of the captured launches, 8 tokens per launch)"},
          open(os.path.join(ROOT, "profiles", "decode_mega_traffic.json"), "w"))
srcp = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
open("/tmp/_src.csv", "w").write(srcp)
lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), "/tmp/_src.csv", "40"], capture_output=True, text=True).stdout
open(os.path.join(dst, "decode_mega_stall_lines.txt"), "w").write(lines)
print("\n".join(out))
