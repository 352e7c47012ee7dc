#!/usr/bin/env python
"""Batched decode (BASELINE configs[2]/[3] shapes): ms per step and tokens/s for B rows through the fused kernels
(B = 1: decode_mega; 2..8: decode_batch, rows share one pass of the weight stream) and the per-op path.

    python tools/batch_bench.py [--rows 1 2 4 8] [--heads 4] [--T 75] [--tokens 64] [--windows 2 3 4] [--per-op]

Timed with CUDA events around the decode launches (prefill excluded), inputs resident; the HBM fraction is
algorithmic bytes (weights once + B x KV) / time against MEASURED_PEAKS.json.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from genvc_b200.config import GenVCDims  # noqa: E402
from genvc_b200.engine import Sampling  # noqa: E402
from genvc_b200.gpt import GPT  # noqa: E402
from genvc_b200.synth import synth_checkpoint  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, nargs="*", default=[1, 2, 4, 8])
ap.add_argument("--heads", type=int, default=4)
ap.add_argument("--T", type=int, default=75)
ap.add_argument("--tokens", type=int, default=64)
ap.add_argument("--windows", type=int, nargs="*", default=[2])
ap.add_argument("--top-k", type=int, default=20)
ap.add_argument("--per-op", action="store_true")
ap.add_argument("--out", default=None)
a = ap.parse_args()

L, D, V = 30, 1024, 1026
dev = torch.device("cuda:0")
ck = synth_checkpoint(n_layer=L, d_model=D, n_head=a.heads, seed=1234)
peak = 6532.9
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
W = 4 * (L * (12 * D * D + 13 * D) + 4 * D + V * (D + 1) + 2 * D)
results = []
Bmax = max(a.rows)
g = GPT(GenVCDims.from_config(ck["config"]), device=dev, max_batch=Bmax)
g.load_state_dict(ck["model"])
g.eval().to(dev).init_gpt_for_inference()
eng = g.engine
for B in a.rows:
    gen = torch.Generator().manual_seed(7)
    codes = torch.randint(0, 256, (B, a.T), generator=gen).to(dev)
    cond = torch.randn((B, 32, D), generator=gen).to(dev)
    P = 32 + a.T + 2
    modes = [(2, w) for w in a.windows] + ([(1, 0)] if a.per_op else [])
    for mode, window in modes:
        if mode == 2:
            eng.tune(window=window)
        sp = Sampling(top_k=a.top_k, top_p=0.85, temperature=0.85, repetition_penalty=2.0, ignore_eos=True,
                      max_new_tokens=a.tokens, seed=5)
        best = None
        for rep in range(4):
            g.compute_embeddings(cond, codes)
            eng.prefill(g._prefix)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ch = eng.decode(a.tokens, sp, mode=mode)
            e1.record()
            torch.cuda.synchronize()
            assert ch.status.tolist()[0] == a.tokens
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        n_fwd = a.tokens - 1
        kv = sum(L * 2 * D * 4 * (P + 1 + j) for j in range(n_fwd))  # read S-1 rows, write 1 (per row of the batch)
        alg = n_fwd * W + B * kv
        r = {"B": B, "heads": a.heads, "T": a.T, "tokens": a.tokens, "mode": "fused" if mode == 2 else "per-op", "window": window,
             "ms_per_step": round(best / n_fwd, 4), "tokens_per_s": round(B * n_fwd / (best * 1e-3), 1),
             "hbm_frac": round(alg / (best * 1e-3) / 1e9 / peak, 4)}
        results.append(r)
        print(json.dumps(r), flush=True)
eng.tune(window=2)
if a.out:
    json.dump(results, open(a.out, "w"), indent=1)
