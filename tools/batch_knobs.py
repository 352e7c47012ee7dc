import os, sys, json
sys.path.insert(0, "/root/repo")
import torch
from genvc_b200.config import GenVCDims
from genvc_b200.engine import Sampling
from genvc_b200.gpt import GPT
from genvc_b200.synth import synth_checkpoint
L, D = 30, 1024
dev = torch.device("cuda:0")
ck = synth_checkpoint(n_layer=L, d_model=D, n_head=4, seed=1234)
g = GPT(GenVCDims.from_config(ck["config"]), device=dev, max_batch=8)
g.load_state_dict(ck["model"]); g.eval().to(dev).init_gpt_for_inference()
eng = g.engine
def run(B, tokens=48):
    gen = torch.Generator().manual_seed(7)
    codes = torch.randint(0, 256, (B, 75), generator=gen).to(dev)
    cond = torch.randn((B, 32, D), generator=gen).to(dev)
    sp = Sampling(top_k=20, top_p=0.85, temperature=0.85, repetition_penalty=2.0, ignore_eos=True, max_new_tokens=tokens, seed=5)
    best = 1e9
    for rep in range(3):
        g.compute_embeddings(cond, codes); eng.prefill(g._prefix)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ch = eng.decode(tokens, sp, mode=2); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / (tokens - 1))
    return best
for B in (2, 8):
    for window in (1, 2, 3, 4, 8):
        eng.tune(window=window)
        print(f"B={B} window={window}: {run(B):.4f} ms/step", flush=True)
    eng.tune(window=2)
    for near, near_ao in ((0, 0), (1, 1), (4, 4), (8, 4), (16, 8), (32, 8)):
        eng.tune(hop_hold=1000 + near + 100 * near_ao)
        print(f"B={B} near={near} near_ao={near_ao}: {run(B):.4f} ms/step", flush=True)
    eng.tune(hop_hold=1000 + 4 + 100 * 4)
    for settle in (0, 100, 300, 600):
        eng.tune(hop_settle_ns=settle)
        print(f"B={B} settle={settle}: {run(B):.4f} ms/step", flush=True)
    eng.tune(hop_settle_ns=0)
