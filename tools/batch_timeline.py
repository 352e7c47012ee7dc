#!/usr/bin/env python
"""Phase profile of the batched fused decode kernel (debug hook: thread 0 of every CTA accumulates the cycles between
phase marks over a whole launch): mean us per layer per phase, for the attention CTAs, the reducer CTAs and the rest.
    python tools/batch_timeline.py [--rows 8] [--heads 4] [--T 75] [--tokens 32]
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from genvc_b200.config import GenVCDims
from genvc_b200.engine import Sampling
from genvc_b200.gpt import GPT
from genvc_b200.synth import synth_checkpoint

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, nargs="*", default=[2, 8])
ap.add_argument("--heads", type=int, default=4)
ap.add_argument("--T", type=int, default=75)
ap.add_argument("--tokens", type=int, default=32)
ap.add_argument("--top-k", type=int, default=20)
a = ap.parse_args()
L, D = 30, 1024
dev = torch.device("cuda:0")
ck = synth_checkpoint(n_layer=L, d_model=D, n_head=a.heads, seed=1234)
g = GPT(GenVCDims.from_config(ck["config"]), device=dev, max_batch=max(a.rows))
g.load_state_dict(ck["model"]); g.eval().to(dev).init_gpt_for_inference()
eng = g.engine
NAMES = ["x2 hop+load", "stats+QKV", "att item", "AO hop", "merge", "PROJ", "x1 hop+load", "stats+FC", "P2", "reduce", "head", "sample+tok"]
GHZ = 1.965
for B in a.rows:
    gen = torch.Generator().manual_seed(7)
    codes = torch.randint(0, 256, (B, a.T), generator=gen).to(dev)
    cond = torch.randn((B, 32, D), generator=gen).to(dev)
    sp = Sampling(top_k=a.top_k, top_p=0.85, temperature=0.85, repetition_penalty=2.0, ignore_eos=True, max_new_tokens=a.tokens, seed=5)
    G = eng.decode_grid
    tr = torch.zeros(G * 16, dtype=torch.int64, device=dev)
    for rep in range(2):
        g.compute_embeddings(cond, codes); eng.prefill(g._prefix)
        eng._check(eng.lib.genvc_debug_trace(eng._ctx, tr.data_ptr(), 16, 0))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ch = eng.decode(a.tokens, sp, mode=2); e1.record(); torch.cuda.synchronize()
        eng._check(eng.lib.genvc_debug_trace(eng._ctx, None, 0, 0))
    t = tr.view(G, 16).cpu().double()
    nf = a.tokens - 1
    per_layer = t / (nf * L) / (GHZ * 1e3)  # us per layer (head / sample: per step / L)
    P = 32 + a.T + 2
    print(f"B={B} H={a.heads}: {e0.elapsed_time(e1)/nf:.4f} ms/step  ({e0.elapsed_time(e1)/nf/L*1e3:.1f} us per layer incl. head/sampling)")
    n_items = min(G, B * a.heads * 8)
    groups = {"CTA 0": slice(0, 1), "first 32 (attention+reducer)": slice(0, 32), "CTAs 32..127": slice(32, 128), "CTAs 128..147 (no reduce)": slice(128, G)}
    print("%-14s" % "phase" + "".join("%30s" % k for k in groups))
    for i, nm in enumerate(NAMES):
        print("%-14s" % nm + "".join("%30.2f" % per_layer[sl, i].mean().item() for sl in groups.values()))
    for i, nm in ((12, "(tile wait dot)"), (13, "(tile wait P2)"), (14, "(P2 stores)"), (15, "(bar after FC)")):
        print("%-14s" % nm + "".join("%30.2f" % per_layer[sl, i].mean().item() for sl in groups.values()))
    print("%-14s" % "sum" + "".join("%30.2f" % per_layer[sl, :12].sum(1).mean().item() for sl in groups.values()))
