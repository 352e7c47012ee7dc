#!/usr/bin/env python
"""Batched decode (BASELINE configs[2]/[3] shapes) through the per-op path: ms per step and tokens/s for B rows.
    python tools/batch_bench.py [B ...]
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from genvc_b200.config import GenVCDims
from genvc_b200.gpt import GPT
from genvc_b200.synth import synth_checkpoint

dev = torch.device("cuda:0")
ck = synth_checkpoint(n_layer=30, d_model=1024, n_head=4, seed=1234)
for B in [int(x) for x in (sys.argv[1:] or ["1", "4", "8"])]:
    g = GPT(GenVCDims.from_config(ck["config"]), device=dev, max_batch=B)
    g.load_state_dict(ck["model"]); g.eval().to(dev).init_gpt_for_inference()
    gen = torch.Generator().manual_seed(7)
    codes = torch.randint(0, 256, (B, 75), generator=gen).to(dev)   # 6 s segments
    cond = torch.randn((B, 32, 1024), generator=gen).to(dev)
    for label, extra in (("batched per-op kernels", dict(top_k=20, decode_mode=1)), ("default path, greedy", dict(top_k=1))):
        kw = dict(do_sample=True, top_p=0.85, temperature=0.85, repetition_penalty=2.0, ignore_eos=True, max_new_tokens=64, **extra)
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ids = g.generate(cond, codes, **kw)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        n = ids.shape[1]
        print(f"B={B} {label}: {n} steps in {dt*1e3:.1f} ms incl. prefill -> {dt*1e3/n:.3f} ms/step, {B*n/dt:.0f} tokens/s", flush=True)
    del g
    torch.cuda.empty_cache()
