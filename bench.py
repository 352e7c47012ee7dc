#!/usr/bin/env python
"""Benchmark of the GenVC codec-token inference path (BASELINE.json metric: codec tokens/s and
first-chunk latency, GenVC_small streaming, 1/2/4/8 x B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg3|cfg4]

Default workload = BASELINE.json configs[1]: GenVC_small dims (L=30, D=1024, H=4, fp32,
seeded synthetic checkpoint in the reference layout), streaming decode of 1 s source segments: T=13 phonetic
tokens (prefix P=47), batch 1, greedy (top_k=1), 24 new codec tokens per segment with EOS suppressed (fixed
work), latents flushed every ``stream_chunk_size=8`` tokens.  One "step" = one segment: compute_embeddings +
prefill (48 rows) + 24 generated tokens, driven through the drop-in ``GPT.compute_embeddings`` /
``GPT.get_generator`` API.

* ``value``      tokens/s with the step's inputs already resident in HBM (device-timed, CUDA events; no per-launch
                 event recording inside this loop)
* ``e2e``        the same through the same public API from pinned HOST buffers, host->device copies of the
                 inputs and device->host read-back of ids + latents inside the timed region
* ``roofline``   fused decode kernel (decode_mega at batch 1, decode_batch for 2..8 rows), measured in a separate
                 short loop: algorithmic bytes per launch / mean launch duration (CUDA events on the launching
                 stream) against the measured HBM copy bandwidth
* ``cpu_baseline`` the CPU oracle (port of the reference path) timed on this box's host cores
* N > 1, cfg2: replicas; rank 0 packs the weights and broadcasts the blob once over NCCL (``init_ms``); every rank
  decodes its own segments (weak scaling, no data-path collective); the ids are gathered over NCCL at the end of
  the timed region; time = max over ranks.

``--config cfg3`` (BASELINE configs[2]: top_k=20 non-streaming, 32 utterances of 10 s = a 6 s + a 4 s segment each)
and ``--config cfg4`` (configs[3]: "GenVC_large" = second-seed checkpoint, streaming, 8 utterances of 30 s = 5 x 6 s
segments, 5 s reference through the perceiver) are STRONG-scaling jobs: the fixed utterance set is dealt to the ranks
(``replicas.shard_units``), each rank batches its utterances (equal-T segments, up to 8 rows through the batched fused
kernel), ids are all-gathered over NCCL inside the timed region; one step = the whole job.

``--impl reference`` times the CPU oracle on a bounded sample of the same workload (rank 0 only).
"""
from __future__ import annotations

import argparse
import datetime
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

L, D, V = 30, 1024, 1026
CHUNK = 8
CODE_STRIDE = 1024

CONFIGS = {
    "cfg2": dict(
        heads=4, seed=1234, T=13, new_tokens=24, s_mel=282, top_k=1,
        metric="codec tokens/sec, GenVC_small streaming decode (1 s segments, batch 1)",
        workload=("cfg2: GenVC_small L=30 D=1024 H=4 fp32, streaming 1 s segments (T=13, P=47), batch 1 per GPU, greedy top_k=1, "
                  "24 new tokens/segment (EOS suppressed), stream_chunk_size=8; step = compute_embeddings + prefill + 24 tokens"),
        parallelism="replicas (weights broadcast once over NCCL, each rank decodes its own segments, ids gathered at the end)",
        scaling="weak"),
    "cfg3": dict(
        heads=4, seed=1234, n_utt=32, segments=[(75, 141), (50, 94)], s_mel=282, top_k=20, streaming=False,
        metric="codec tokens/sec, GenVC_small non-streaming top_k=20, 32 utterances of 10 s (batched)",
        workload=("cfg3: GenVC_small L=30 D=1024 H=4 fp32, non-streaming generate + teacher-forced latent pass, top_k=20 top_p=0.85, "
                  "32 utterances x (6 s segment T=75 -> 141 tokens + 4 s segment T=50 -> 94 tokens), 3 s reference (282 mel frames) "
                  "through the perceiver per utterance, EOS suppressed; utterances dealt to the ranks, equal-T segments batched up to "
                  "8 rows; step = the whole 64-segment job"),
        parallelism="replicas, fixed utterance set sharded across ranks (strong scaling), ids all-gathered over NCCL",
        scaling="strong"),
    "cfg4": dict(
        heads=4, seed=4321, n_utt=8, segments=[(75, 141)] * 5, s_mel=469, top_k=15, streaming=True,
        metric="codec tokens/sec, GenVC_large streaming, batch 8, 30 s source / 5 s reference",
        workload=("cfg4: GenVC_large (= GenVC_small dims L=30 D=1024 H=4 fp32, second-seed checkpoint: the README distinguishes the "
                  "two by training data only), streaming stream_chunk_size=8, top_k=15 top_p=0.85, 8 utterances x 5 segments of 6 s "
                  "(T=75 -> 141 tokens), 5 s reference (469 mel frames) through the perceiver, EOS suppressed; utterances dealt to "
                  "the ranks and batched; step = the whole 40-segment job"),
        parallelism="replicas, fixed utterance set sharded across ranks (strong scaling), ids all-gathered over NCCL",
        scaling="strong"),
}
L2_NOTE = "inputs larger than L2: 1.516 GB of fp32 weights are re-read for every generated token (L2 = 126 MB)"


def config_dict(name: str) -> dict:
    """The ``config`` object of the JSON line: identical for the two arms of a workload."""
    c = CONFIGS[name]
    return {"workload": c["workload"], "parallelism": c["parallelism"], "l2": L2_NOTE}


def sampling_kw(c: dict, new_tokens: int) -> dict:
    return dict(do_sample=True, top_p=0.85, top_k=c["top_k"], temperature=0.85, num_beams=1, length_penalty=1.0,
                repetition_penalty=2.0, output_attentions=False, num_return_sequences=1, output_hidden_states=True,
                ignore_eos=True, max_new_tokens=new_tokens, stream_chunk_size=CHUNK)


def weight_bytes() -> int:
    """fp32 bytes every decode step must read: blocks + ln_f/final_norm + mel_head + 2 embedding rows."""
    return 4 * (L * (12 * D * D + 13 * D) + 4 * D + V * (D + 1) + 2 * D)


def kv_bytes(S: int) -> int:
    """KV cache bytes of one decode step of ONE row attending S keys: read S-1 cached rows, write 1 (K and V, all layers)."""
    return L * 2 * D * 4 * (S - 1) + L * 2 * D * 4


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_per_forward(kernel: str):
    """DRAM bytes per forward of the fused kernel from the committed ncu capture (profiles/*_traffic.json), or None."""
    tp = os.path.join(ROOT, "profiles", f"{kernel}_traffic.json")
    if os.path.exists(tp):
        try:
            j = json.load(open(tp))
            if "dram_bytes_per_forward" in j:
                return float(j["dram_bytes_per_forward"])
            return float(j["dram_bytes_per_launch"]) / float(j.get("forwards_per_launch", 8))
        except Exception:
            return None
    return None


def synth_inputs(n_segments: int, T: int, s_mel: int, seed: int = 7):
    g = torch.Generator().manual_seed(seed)
    codes = torch.randint(0, 256, (n_segments, 1, T), generator=g)
    g = torch.Generator().manual_seed(11)
    mel = torch.randn((1, 80, s_mel), generator=g)
    return codes, mel


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 100 ms while the timed region runs."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        if os.environ.get("GENVC_BENCH_NOSMI") == "1":  # debug: rule the sampler in or out as a perturbation
            self.first = ""
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        # nvidia-smi spends its first few hundred ms initialising NVML (driver locks that delay kernel launches): wait
        # for its first sample so that only the cheap periodic queries overlap the timed region
        self.first = ""
        if self.proc is not None:
            try:
                import select
                if select.select([self.proc.stdout], [], [], 5.0)[0]:
                    self.first = self.proc.stdout.readline()
            except Exception:
                pass

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        rows = []
        for line in (self.first + out).splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), f[4:8]))
            except Exception:
                continue
        inside = [r for r in rows if t0 - 0.05 <= r[0] <= t1 + 0.05] or rows[-3:]
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in inside for i, v in enumerate(r[3]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(r[1] for r in inside), "sm_max_mhz": inside[0][2], "reasons": reasons,
                "samples": len(inside)}


# ----------------------------------------------------------------------------------------- CPU arm
def cpu_sample(cfg_name: str, n_units: int, warmup: int, want_ids: bool = False):
    """The oracle (CPU port of the reference path) on a bounded sample of the workload.
    Returns (tokens/s, seconds, cores, ids of the first timed unit or None, description of the sample)."""
    from genvc_b200.synth import synth_checkpoint
    from oracle.genvc_oracle import SamplingParams, draw_exponential_noise, load_oracle

    c = CONFIGS[cfg_name]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ck = synth_checkpoint(n_layer=L, d_model=D, n_head=c["heads"], seed=c["seed"])
    o = load_oracle(ck)
    sp = SamplingParams(top_k=c["top_k"], top_p=0.85, temperature=0.85, repetition_penalty=2.0)
    ids0 = None
    with torch.inference_mode():
        if cfg_name == "cfg2":
            codes, mel = synth_inputs(max(n_units + warmup, 1), c["T"], c["s_mel"])
            cond = o.get_gpt_cond_latents([mel])
            for i in range(warmup):
                o.generate(cond, codes[i], sp, max_new_tokens=c["new_tokens"], ignore_eos=True)
            times = []
            for i in range(n_units):
                t = time.perf_counter()
                ids, _ = o.generate(cond, codes[warmup + i], sp, max_new_tokens=c["new_tokens"], ignore_eos=True)
                times.append(time.perf_counter() - t)
                if i == 0:
                    ids0 = ids
            total = sum(times)
            tokens = c["new_tokens"] * n_units
            desc = (f"{n_units} segments x {c['new_tokens']} tokens after {warmup} warm-up segment(s) ({total:.1f} s), "
                    "oracle/genvc_oracle.py on torch CPU fp32")
        else:
            # one batch of 8 rows of the shortest segment shape, a bounded number of tokens (the CPU path costs ~0.1 s/step)
            T, M = min(c["segments"])
            M = min(M, 48)
            B = 8
            g = torch.Generator().manual_seed(7)
            codes = torch.randint(0, 256, (B, T), generator=g)
            mel = torch.randn((B, 80, c["s_mel"]), generator=torch.Generator().manual_seed(11))
            t = time.perf_counter()
            cond = o.get_style_emb(mel).transpose(1, 2).contiguous()
            noise = draw_exponential_noise((M, B, V), torch.Generator().manual_seed(3))
            ids, lats = o.generate(cond, codes, sp, noise=noise, max_new_tokens=M, ignore_eos=True)
            if not c["streaming"]:
                o.forward_latents(codes, ids, cond)
            total = time.perf_counter() - t
            tokens = B * M
            desc = (f"one batch of {B} rows: perceiver ({c['s_mel']} mel frames) + T={T} prefill + {M} tokens per row"
                    + ("" if c["streaming"] else " + teacher-forced latent pass") + f" ({total:.1f} s), oracle/genvc_oracle.py on torch CPU fp32")
    return tokens / total, total, cores, (ids0 if want_ids else None), desc


def run_reference(args, rank: int):
    if rank != 0:
        return
    c = CONFIGS[args.config]
    t0 = time.perf_counter()
    tps, total, cores, _, desc = cpu_sample(args.config, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": c["metric"], "value": round(tps, 3), "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * total / max(args.steps, 1), 3),
        "higher_is_better": True, "scaling": c["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.config),
        "cpu_baseline": {"value": round(tps, 3), "unit": "tokens/s", "cores": cores, "kind": "port",
                         "sample": desc + " (the reference's own modules need /root/reference, absent on the GPU box; the port is "
                                          "pinned to them by tests/test_oracle_vs_reference.py and the golden fixtures)"},
        "e2e": {"value": round(tps, 3), "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": round(time.perf_counter() - t0, 1),
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- GPU arm: shared pieces
def consume(gen):
    toks, lats = [], []
    for tok, lat in gen:
        toks.append(tok)
        lats.append(lat)
    return toks, lats


def init_model(c: dict, dev, rank: int, world: int, max_batch: int, max_mel_frames: int = 576):
    """Rank 0 builds + packs the checkpoint, one NCCL broadcast of the blob, every rank binds its replica.
    Returns (model, init_ms of this rank incl. broadcast and the stream / tensor-core repacking)."""
    from genvc_b200.config import make_config_dict
    from genvc_b200.replicas import init_replica
    from genvc_b200.synth import synth_checkpoint

    cfg = make_config_dict(L, D, c["heads"])
    ck = synth_checkpoint(n_layer=L, d_model=D, n_head=c["heads"], seed=c["seed"]) if rank == 0 else {"config": cfg, "model": None}
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    model = init_replica(ck, dev, rank, world, max_batch=max_batch, max_mel_frames=max_mel_frames)
    torch.cuda.synchronize(dev)
    return model, 1e3 * (time.perf_counter() - t0)


def max_over_ranks(x: float, dev, world: int) -> float:
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def roofline_from_timing(timing, rows: int, kernel: str) -> dict:
    wb = weight_bytes()
    alg, dur, nf = [], [], []
    for s_ev, e_ev, n_fwd, first_S in timing:
        alg.append(sum(wb + rows * kv_bytes(first_S + j) for j in range(n_fwd)))
        dur.append(s_ev.elapsed_time(e_ev))
        nf.append(n_fwd)
    peak, peak_src = measured_peaks()
    achieved = (sum(alg) / len(alg)) / (sum(dur) / len(dur) * 1e-3) / 1e9
    tpf = traffic_per_forward(kernel)
    return {"bound": "hbm", "kernel": kernel + "_kernel", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4),
            # DRAM bytes for the SAME launch mix as bytes_per_launch: ncu's per-forward traffic x mean forwards per launch
            "traffic": None if tpf is None else int(tpf * sum(nf) / len(nf)),
            "peak_source": peak_src, "launches_timed": len(dur), "forwards_per_launch": round(sum(nf) / len(nf), 3),
            "bytes_per_launch": int(sum(alg) / len(alg)), "ms_per_launch": round(sum(dur) / len(dur), 4),
            "decode_ms_per_forward": round(sum(dur) / max(1, sum(nf)), 4)}


# ----------------------------------------------------------------------------------------- GPU arm: cfg2 (headline)
def run_cfg2(args, rank: int, world: int, local_rank: int):
    import torch.distributed as dist

    from genvc_b200.replicas import gather_ids

    c = CONFIGS["cfg2"]
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    model, init_ms = init_model(c, dev, rank, world, max_batch=1)
    g = model.gpt
    eng = g.engine
    kw = sampling_kw(c, c["new_tokens"])
    n_seg = args.warmup + args.steps
    codes_host, mel_host = synth_inputs(n_seg, c["T"], c["s_mel"], seed=7 + rank)
    codes_dev = codes_host.to(dev)
    mel_dev = mel_host.to(dev)
    cond_dev = model.get_gpt_cond_latents_from_mels([mel_dev]).contiguous()
    torch.cuda.synchronize(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def segment(cond, codes):
        fake = g.compute_embeddings(cond, codes)
        return consume(g.get_generator(fake_inputs=fake, **kw))

    # ---- warm-up
    toks = None
    for i in range(args.warmup):
        toks, _ = segment(cond_dev, codes_dev[i])
    if world > 1:
        # the first all-gather builds NCCL's channels (tens of ms): part of the warm-up, not of a timed step
        warm = torch.stack(toks, 1) if toks else torch.zeros((1, 1), dtype=torch.int64, device=dev)
        gather_ids(warm, world, pad=g.stop_audio_token)
    # ---- timed: device-resident inputs, no per-launch instrumentation
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    # the GPU must not idle into the timed region (starting nvidia-smi takes a few hundred ms on rank 0 and an idle GPU
    # drops its clocks): every rank runs two more untimed segments right before the barrier
    eng.timing = None  # (before the last warm-up segments: they must take the same host path as the timed ones)
    import gc
    gc.collect()
    gc.disable()  # a generation-2 collection inside a 100 ms timed region is a 5-10 ms host stall (stays off: the process
    #               only runs the remaining timed loops and exits); collected HERE so that the GPU does not idle afterwards
    for i in range(min(2, args.warmup)):
        segment(cond_dev, codes_dev[i])
    barrier()
    launches0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record(torch.cuda.current_stream(dev))
    all_ids = []
    dbg = os.environ.get("GENVC_BENCH_DEBUG") == "1"
    marks = [time.perf_counter()]
    for i in range(args.steps):
        toks, _ = segment(cond_dev, codes_dev[args.warmup + i])
        all_ids.append(torch.stack(toks, 1))
        if dbg:
            marks.append(time.perf_counter())
    if dbg and rank == 0:
        print("per-step ms:", [round(1e3 * (b - a), 2) for a, b in zip(marks, marks[1:])], file=sys.stderr)
    local_ids = torch.cat(all_ids, 0)  # [steps, new_tokens]
    gathered = gather_ids(local_ids, world, pad=g.stop_audio_token)  # NCCL all-gather of the ids (no-op at N = 1)
    e1.record(torch.cuda.current_stream(dev))
    barrier()
    w1 = time.time()
    launches = eng.launch_count - launches0
    ms_max = max_over_ranks(e0.elapsed_time(e1), dev, world)
    clocks = sampler.stop(w0, w1) if sampler else None
    assert sum(int(t.shape[0]) for t in gathered) == args.steps * world

    # ---- roofline of the fused decode kernel: separate short loop with per-launch CUDA events
    eng.timing = []
    for i in range(min(args.steps, 10)):
        segment(cond_dev, codes_dev[args.warmup + i])
    torch.cuda.synchronize(dev)
    timing, eng.timing = eng.timing, None
    roof = roofline_from_timing(timing, 1, "decode_mega")

    # ---- e2e: pinned host inputs -> public API -> host outputs, copies inside the timed region
    cond_host = cond_dev.cpu().pin_memory()
    codes_pin = codes_host.pin_memory()

    def segment_e2e(i):
        cond = cond_host.to(dev, non_blocking=True)
        codes = codes_pin[i].to(dev, non_blocking=True)
        toks, lats = segment(cond, codes)
        return torch.stack(toks, 1).cpu(), torch.stack(lats, 1).cpu()

    for i in range(min(args.warmup, 2)):
        segment_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        segment_e2e(args.warmup + i)
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0, dev, world)
    h2d = cond_host.numel() * 4 + c["T"] * 8
    d2h = c["new_tokens"] * (8 + D * 4) + (c["new_tokens"] // CHUNK) * 16

    # ---- first-chunk latency (path-only): mel + codes on the device -> perceiver -> embeddings ->
    # prefill -> 8 (id, latent) pairs visible on the host
    lat_ms = []
    for i in range(8):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        cond = model.get_gpt_cond_latents_from_mels([mel_dev])
        fake = g.compute_embeddings(cond, codes_dev[i % n_seg])
        gen = g.get_generator(fake_inputs=fake, **kw)
        toks, lats = [], []
        for _ in range(CHUNK):
            tk, lt = next(gen)
            toks.append(tk)
            lats.append(lt)
        torch.stack(toks).cpu(), torch.stack(lats).cpu()
        lat_ms.append(1e3 * (time.perf_counter() - t0))
        for _ in gen:  # drain
            pass
    first_chunk_ms = statistics.median(lat_ms[2:])

    # ---- first AUDIO (informational; the stage after the path): the same, but the 8 latents stay on the
    # device, go through the x4 interpolation and the CUDA HiFi-GAN generator (synthetic weights of the reference's vocoder
    # config) and the 8192-sample waveform chunk is what reaches the host (inference_utils.py:196-207)
    first_audio_ms = None
    try:  # informational: a failure here must not take the benchmark line down
        from genvc_b200.inference.inference_utils import _vocode
        from genvc_b200.synth import synth_hifigan_state
        from genvc_b200.vocoder import HiFiGAN
        model.hifigan = HiFiGAN.from_config({}, device=dev).load_state_dict(synth_hifigan_state(77))
        aud_ms = []
        for i in range(8):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            cond = model.get_gpt_cond_latents_from_mels([mel_dev])
            fake = g.compute_embeddings(cond, codes_dev[i % n_seg])
            gen = g.get_generator(fake_inputs=fake, **kw)
            lats = [next(gen)[1] for _ in range(CHUNK)]
            wav = _vocode(model, torch.cat(lats, dim=0)[None, :])
            wav.cpu()
            aud_ms.append(1e3 * (time.perf_counter() - t0))
            for _ in gen:  # drain
                pass
        first_audio_ms = round(statistics.median(aud_ms[2:]), 3)
    except Exception as e:  # noqa: BLE001
        print(f"first_audio_ms not measured: {e!r}", file=sys.stderr)

    # ---- prefill alone (compute_embeddings + prefill of 48 rows), CUDA events
    pf = []
    for i in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.compute_embeddings(cond_dev, codes_dev[i % n_seg])
        eng.prefill(g._prefix)
        b.record()
        torch.cuda.synchronize(dev)
        pf.append(a.elapsed_time(b))
    prefill_ms = statistics.median(pf[1:])

    total_tokens = c["new_tokens"] * args.steps * world
    line = {
        "metric": c["metric"], "value": round(total_tokens / (ms_max * 1e-3), 2), "unit": "tokens/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_max / args.steps, 4),
        "higher_is_better": True, "scaling": c["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict("cfg2"),
        "e2e": {"value": round(total_tokens / e2e_s, 2), "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "first_chunk_ms": round(first_chunk_ms, 3),
        "first_audio_ms": first_audio_ms,
        "prefill_ms": round(prefill_ms, 4),
        "decode_ms_per_token": roof["decode_ms_per_forward"],
        "init_ms": round(max_over_ranks(init_ms, dev, world), 1),
        "roofline": roof,
        "clocks": clocks,
    }
    if rank == 0:
        # ---- CPU baseline on a bounded sample of the same workload (N = 1 only)
        if world == 1 and not args.no_cpu:
            n_cpu = max(2, min(8, args.steps))
            tps, total, cores, ids_cpu, desc = cpu_sample("cfg2", n_cpu, 1, want_ids=True)
            line["cpu_baseline"] = {"value": round(tps, 3), "unit": "tokens/s", "cores": cores, "kind": "port", "sample": desc}
            # the oracle's first timed segment uses codes[1] of seed 7 (rank 0) = the GPU's warm-up segment 1
            toks, _ = segment(cond_dev, codes_dev[1])
            line["parity"] = {"ids_equal_oracle": bool(torch.equal(torch.stack(toks, 1).cpu(), ids_cpu))}
        print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- GPU arm: cfg3 / cfg4 (batched, strong scaling)
def run_utterances(args, name: str, rank: int, world: int, local_rank: int):
    import torch.distributed as dist

    from genvc_b200.replicas import gather_ids, plan_batches, shard_units

    c = CONFIGS[name]
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    mine = shard_units(c["n_utt"], rank, world)
    groups = plan_batches(len(mine), 8)  # utterance groups of up to 8 rows (all utterances share the segment shapes)
    rows_max = max((len(gr) for gr in groups), default=1)
    model, init_ms = init_model(c, dev, rank, world, max_batch=max(rows_max, 1))
    g = model.gpt
    eng = g.engine
    segs = c["segments"]
    tokens_per_utt = sum(m for _, m in segs)
    # synthetic inputs of the whole job (same on every rank), indexed by utterance
    gen = torch.Generator().manual_seed(7)
    codes_host = [torch.randint(0, 256, (c["n_utt"], T), generator=gen) for T, _ in segs]
    mel_host = torch.randn((c["n_utt"], 80, c["s_mel"]), generator=torch.Generator().manual_seed(11))
    codes_dev = [x.to(dev) for x in codes_host]
    mel_dev = mel_host.to(dev)
    codes_pin = [x.pin_memory() for x in codes_host]
    mel_pin = mel_host.pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def job(from_host: bool):
        """The rank's share of the job; returns the ids [n_local_utt, tokens_per_utt] (device) and host copies if from_host."""
        out_ids = []
        for gi, gr in enumerate(groups):
            utt = torch.tensor([mine[i] for i in gr])
            mel = (mel_pin[utt].to(dev, non_blocking=True) if from_host else mel_dev[utt.to(dev)])
            cond = model.get_gpt_cond_latents_from_mels([mel]).contiguous()  # perceiver, one 5 s / 3 s chunk per utterance
            row_ids = []
            for si, (T, M) in enumerate(segs):
                codes = (codes_pin[si][utt].to(dev, non_blocking=True) if from_host else codes_dev[si][utt.to(dev)])
                kw = sampling_kw(c, M)
                kw["seed"] = 1000 * gi + si + 1
                if c["streaming"]:
                    fake = g.compute_embeddings(cond, codes)
                    toks, lats = consume(g.get_generator(fake_inputs=fake, **kw))
                    ids = torch.stack(toks, 1)
                    lat = torch.stack(lats, 1)
                else:
                    kw.pop("stream_chunk_size")
                    ids = g.generate(cond, codes, **kw)
                    lens = torch.full((len(gr),), M * CODE_STRIDE)
                    lat = g(codes, torch.full((len(gr),), T), ids, lens, cond_latents=cond, return_latent=True)
                if from_host:
                    ids.cpu(), lat.cpu()
                row_ids.append(ids)
            out_ids.append(torch.cat(row_ids, 1))
        local = torch.cat(out_ids, 0) if out_ids else torch.empty((0, tokens_per_utt), dtype=torch.int64, device=dev)
        return gather_ids(local, world, pad=g.stop_audio_token)

    for _ in range(args.warmup):
        job(False)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    import gc
    gc.collect()
    gc.disable()  # see run_cfg2
    if args.warmup > 0:
        job(False)  # keeps the GPU busy while nvidia-smi starts on rank 0 (an idle GPU drops its clocks)
    barrier()
    launches0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record(torch.cuda.current_stream(dev))
    for _ in range(args.steps):
        gathered = job(False)
    e1.record(torch.cuda.current_stream(dev))
    barrier()
    w1 = time.time()
    launches = eng.launch_count - launches0
    ms_max = max_over_ranks(e0.elapsed_time(e1), dev, world)
    clocks = sampler.stop(w0, w1) if sampler else None
    assert sum(int(t.shape[0]) for t in gathered) == c["n_utt"], "gathered ids do not cover the utterance set"

    # roofline of the fused kernel the rank actually ran (decode_batch for 2..8 rows, decode_mega for 1): one instrumented job
    eng.timing = []
    job(False)
    torch.cuda.synchronize(dev)
    timing, eng.timing = eng.timing, None
    roof = roofline_from_timing(timing, rows_max, "decode_batch" if rows_max > 1 else "decode_mega") if timing else None

    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        job(True)
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0, dev, world)
    n_local = len(mine)
    h2d = n_local * (80 * c["s_mel"] * 4 + sum(T for T, _ in segs) * 8)
    d2h = n_local * tokens_per_utt * (8 + D * 4)

    total_tokens = c["n_utt"] * tokens_per_utt * args.steps
    line = {
        "metric": c["metric"], "value": round(total_tokens / (ms_max * 1e-3), 2), "unit": "tokens/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_max / args.steps, 3),
        "higher_is_better": True, "scaling": c["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(name),
        "e2e": {"value": round(total_tokens / e2e_s, 2), "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches), "rows_per_launch": rows_max, "utterances_per_rank": n_local,
        "init_ms": round(max_over_ranks(init_ms, dev, world), 1),
        "roofline": roof, "clocks": clocks,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu:
            tps, total, cores, _, desc = cpu_sample(name, 1, 0)
            line["cpu_baseline"] = {"value": round(tps, 3), "unit": "tokens/s", "cores": cores, "kind": "port", "sample": desc}
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 40 if args.config == "cfg2" else 3
    if args.warmup is None:
        args.warmup = 3 if args.config == "cfg2" else 1
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        if args.config == "cfg2":
            run_cfg2(args, rank, world, local_rank)
        else:
            run_utterances(args, args.config, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
