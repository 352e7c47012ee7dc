#!/usr/bin/env python
"""Turn the artefacts of tools/gpu_evidence.sh (gpurun_out/<tag>/) into tracked summaries under profiles/<tag>/:

* launch_list_summary.txt   shares of the default bench step per kernel (ncu gpu__time_duration.sum)
* <kernel>_ncu_summary.txt  counters of the ncu --set full capture of each kernel family (DRAM bytes / GB/s, % of peak,
                            tensor-pipe %, issue-slot %, registers, shared memory, top stall reasons per issue)
* sass_histogram.txt        per kernel of libgenvc_b200.so: counts of the opcodes that prove which hardware path it
                            uses (UTCHMMA / LDTM / UTCBAR = tcgen05, UBLKCP = bulk TMA copies, HMMA / LDSM = legacy
                            warp-level tensor path, FFMA, SYNCS = mbarrier)
* *_traffic.json            DRAM bytes per forward of the fused decode kernels (bench.py's roofline.traffic)

    python tools/summarize_evidence.py r02c
"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
src = os.path.join(ROOT, "gpurun_out", tag)
dst = os.path.join(ROOT, "profiles", tag)
os.makedirs(dst, exist_ok=True)
for f in ("gpu.txt", "bench.json", "bench_cfg3.json", "bench_cfg4.json", "timeline.txt", "batch_bench.json", "stage_bench.json"):
    if os.path.exists(os.path.join(src, f)):
        shutil.copy(os.path.join(src, f), dst)

# ---------------------------------------------------------------- launch list
ll = os.path.join(src, "launches.csv")
if os.path.exists(ll):
    rows = list(csv.reader(open(ll)))
    hdr, start = None, 0
    for i, r in enumerate(rows):
        if r and r[0] == "ID":
            hdr, start = r, i + 1
            break
    if hdr:
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
            f.write("ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 python bench.py --steps 3 --warmup 3 --no-cpu\n")
            f.write(f"(cold-cache, serialised per-launch times: compare SHARES) total {tot/1e3:.1f} us over {sum(v[0] for v in agg.values())} launches\n")
            for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(f"{100*v[1]/tot:6.2f}%  n={v[0]:4d}  {v[1]/1e3:10.1f} us  avg {v[1]/v[0]/1e3:8.1f} us  {k[:110]}\n")
        print(open(os.path.join(dst, "launch_list_summary.txt")).read())

# ---------------------------------------------------------------- full captures
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
MUL = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}
NOTES = {
    "decode_mega": "fused single-row decode kernel, one launch = stream_chunk_size forwards (bench.py --steps 2 --warmup 3 --no-cpu)",
    "decode_batch": "fused batched decode kernel, 8 rows x 11 forwards (tools/batch_bench.py --rows 8 --tokens 12)",
    "gemm_tc": "tcgen05 3xTF32 GEMM of the prefill (48 rows)",
    "attention": "causal attention of the prefill (48 rows x 4 heads x 256): one CTA per (row, head)",
    "vocoder_conv1d": "HiFi-GAN generator / content-DVAE convolution (tools/vocoder_bench.py)",
    "splitk_ln": "split-K reduction + bias + residual + LayerNorm epilogue of the prefill GEMMs",
    "kv_attention": "single-query KV-cache attention (BASELINE configs[4] microbenchmark)",
    "pc_attention_tc": "perceiver cross-attention on tcgen05 (one CTA per element and head)",
}
for rep in sorted(f for f in os.listdir(src) if f.endswith(".ncu-rep")):
    name = rep[:-8]
    raw = subprocess.run(["ncu", "-i", os.path.join(src, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    if len(rr) < 3:
        continue
    hdr = rr[0]
    kcol = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
    out = [f"ncu --set full --clock-control none --import-source on -k regex:{name} -c 1   ({NOTES.get(name, '')})",
           "kernel: " + (rr[2][kcol] if kcol is not None else name), ""]
    rd = wr = dur = None
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            vals = [r[i] for r in rr[2:]]
            out.append(f"{k:84s} {rr[1][i]:16s} " + "  ".join(vals))
            try:
                if k == "dram__bytes_read.sum":
                    rd = float(vals[0]) * MUL.get(rr[1][i], 1)
                if k == "dram__bytes_write.sum":
                    wr = float(vals[0]) * MUL.get(rr[1][i], 1)
                if k == "gpu__time_duration.sum":
                    dur = float(vals[0]) * {"msecond": 1e-3, "usecond": 1e-6, "second": 1, "nsecond": 1e-9, "ms": 1e-3, "us": 1e-6, "s": 1, "ns": 1e-9}.get(rr[1][i], 1e-9)
            except ValueError:
                pass
    if rd is not None and wr is not None and dur:
        out += ["", f"DRAM traffic {(rd + wr)/1e9:.4f} GB in {dur*1e3:.4f} ms = {(rd + wr)/dur/1e9:.1f} GB/s under the profiler "
                    "(serialised, cold cache: for the absolute rate see bench.py's roofline)"]
        if name == "decode_mega":
            # a launch of the default bench decodes 8 forwards (7 in the first launch after a prefill; -s 4 lands on an 8-forward launch)
            json.dump({"dram_bytes_per_launch": rd + wr, "forwards_per_launch": 8, "dram_bytes_per_forward": (rd + wr) / 8,
                       "source": f"profiles/{tag}/decode_mega_ncu_summary.txt (dram__bytes_read.sum + dram__bytes_write.sum, one 8-forward launch)"},
                      open(os.path.join(ROOT, "profiles", "decode_mega_traffic.json"), "w"))
        if name == "decode_batch":
            json.dump({"dram_bytes_per_launch": rd + wr, "forwards_per_launch": 11, "dram_bytes_per_forward": (rd + wr) / 11,
                       "source": f"profiles/{tag}/decode_batch_ncu_summary.txt (one launch of 8 rows x 11 forwards)"},
                      open(os.path.join(ROOT, "profiles", "decode_batch_traffic.json"), "w"))
    open(os.path.join(dst, f"{name}_ncu_summary.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out), "\n")
    srcp = subprocess.run(["ncu", "-i", os.path.join(src, rep), "--page", "source", "--csv", "--print-source", "cuda,sass"],
                          capture_output=True, text=True).stdout
    open("/tmp/_src.csv", "w").write(srcp)
    lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), "/tmp/_src.csv", "25"], capture_output=True, text=True).stdout
    open(os.path.join(dst, f"{name}_stall_lines.txt"), "w").write(lines)

# ---------------------------------------------------------------- SASS opcode histogram of the shipped library
so = os.path.join(ROOT, "genvc_b200", "libgenvc_b200.so")
if os.path.exists(so):
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTCALLOC", "UBLKCP", "UTMALDG", "UTMASTG", "HMMA", "LDSM", "FFMA", "SYNCS", "USETMAXREG",
             "LDGSTS", "BAR.SYNC", "MEMBAR", "ATOMG", "REDG", "LDS", "STS", "LDG", "STG", "SHFL", "MUFU"]
    cur, hist, total = None, collections.OrderedDict(), {}
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            hist[cur] = collections.Counter()
            total[cur] = 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            op = m.group(1)
            total[cur] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    hist[cur][w] += 1
                    break
    demangle = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
    with open(os.path.join(dst, "sass_histogram.txt"), "w") as f:
        f.write("cuobjdump -sass genvc_b200/libgenvc_b200.so : opcode counts per kernel (tcgen05 = UTCHMMA/LDTM/UTCBAR; bulk TMA copy = UBLKCP;\n"
                "legacy warp-level tensor path = HMMA/LDSM; there is no UTMALDG: every TMA here is the 1-D bulk copy over pre-tiled streams)\n\n")
        for (k, h), name in zip(hist.items(), demangle):
            if total[k] < 50:
                continue
            short = re.sub(r"\(.*", "", name)
            f.write(f"{short[:100]:100s} {total[k]:6d} instr  " + "  ".join(f"{w}={h[w]}" for w in WATCH if h[w]) + "\n")
    print(open(os.path.join(dst, "sass_histogram.txt")).read()[:3000])
