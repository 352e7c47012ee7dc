#!/usr/bin/env python
"""Sweep the fused decode kernel's tuning knobs (genvc_debug_tune) and print ms/token for each setting."""
import sys, os, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from genvc_b200.config import GenVCDims
from genvc_b200.gpt import GPT
from genvc_b200.synth import synth_checkpoint

dev = torch.device("cuda:0")
ck = synth_checkpoint(n_layer=30, d_model=1024, n_head=4, seed=1234)
g = GPT(GenVCDims.from_config(ck["config"]), device=dev)
g.load_state_dict(ck["model"]); g.eval().to(dev).init_gpt_for_inference()
eng = g.engine
gen = torch.Generator().manual_seed(7)
codes = torch.randint(0, 256, (1, 13), generator=gen).to(dev)
cond = torch.randn((1, 32, 1024), generator=gen).to(dev)
kw = dict(do_sample=True, top_p=0.85, top_k=1, temperature=0.85, repetition_penalty=2.0, ignore_eos=True, max_new_tokens=24, stream_chunk_size=8)
settles = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,50,100,200,400".split(","))]
windows = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "4".split(","))]
holds = [int(x) for x in (sys.argv[3].split(",") if len(sys.argv) > 3 else "0".split(","))]
aheads = [int(x) for x in (sys.argv[4].split(",") if len(sys.argv) > 4 else "-1".split(","))]
for w, sset, hh, ah in itertools.product(windows, settles, holds, aheads):
    eng.tune(window=w, hop_settle_ns=sset, hop_hold=hh, l2_ahead=ah)
    best = 1e9
    for rep in range(4):
        eng.timing = []
        fake = g.compute_embeddings(cond, codes)
        for _ in g.get_generator(fake_inputs=fake, **kw):
            pass
        torch.cuda.synchronize()
        t = sum(a.elapsed_time(b) for a, b, _, _ in eng.timing) / sum(n for _, _, n, _ in eng.timing)
        best = min(best, t)
    print(f"window {w} settle {sset:4d} ns hold {hh} ahead {ah}: {best:.4f} ms/token", flush=True)
