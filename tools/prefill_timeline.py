#!/usr/bin/env python
"""Phase profile of the persistent fused prefill kernel (debug hook: thread 0 of every CTA accumulates cycles per phase kind).
    python tools/prefill_timeline.py [--T 13] [--heads 4]
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from genvc_b200.config import GenVCDims
from genvc_b200.gpt import GPT
from genvc_b200.synth import synth_checkpoint

ap = argparse.ArgumentParser()
ap.add_argument("--T", type=int, default=13)
ap.add_argument("--heads", type=int, default=4)
a = ap.parse_args()
L, D = 30, 1024
dev = torch.device("cuda:0")
ck = synth_checkpoint(n_layer=L, d_model=D, n_head=a.heads, seed=1234)
g = GPT(GenVCDims.from_config(ck["config"]), device=dev)
g.load_state_dict(ck["model"]); g.eval().to(dev).init_gpt_for_inference()
eng = g.engine
gen = torch.Generator().manual_seed(7)
codes = torch.randint(0, 256, (1, a.T), generator=gen).to(dev)
cond = torch.randn((1, 32, D), generator=gen).to(dev)
G = eng.decode_grid
tr = torch.zeros(G * 16, dtype=torch.int64, device=dev)
NAMES = ["c_attn GEMM", "attn c_proj GEMM", "c_fc GEMM", "mlp c_proj GEMM", "barrier after GEMM (x4)", "flat reduce (x2)", "row reduce+LN (x2)",
         "barrier after reduce (x4)", "attention", "barrier after attention"]
for rep in range(3):
    g.compute_embeddings(cond, codes)
    eng._check(eng.lib.genvc_debug_trace(eng._ctx, tr.data_ptr(), 16, 0))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.prefill(g._prefix); e1.record(); torch.cuda.synchronize()
    eng._check(eng.lib.genvc_debug_trace(eng._ctx, None, 0, 0))
t = tr.view(G, 16).cpu().double() / L / 1.965e3  # us per layer
print(f"prefill of {32 + a.T + 3} rows: {e0.elapsed_time(e1):.3f} ms")
print("%-28s %10s %10s %10s" % ("phase (us per layer)", "median", "min", "max"))
NAMES += ["", "(GEMM: loads + stage loop, 4 phases)", "(GEMM: wait for MMAs)", "(GEMM: TMEM epilogue)"]
for i, n in enumerate(NAMES):
    print("%-28s %10.2f %10.2f %10.2f" % (n, t[:, i].median().item(), t[:, i].min().item(), t[:, i].max().item()))
print("%-28s %10.2f" % ("sum of medians", sum(t[:, i].median().item() for i in range(10))))
