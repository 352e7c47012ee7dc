#!/usr/bin/env python
"""Single-row fused decode on a long generation (6 s segment: T = 75 content codes, 140 new tokens): ms per token with and
without the projected-value cache variant.   GENVC_VW=0|1 python tools/long_gen_bench.py"""
import os, sys, json
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
codes = torch.randint(0, 256, (1, 75), generator=gen).to(dev)
cond = torch.randn((1, 32, 1024), generator=gen).to(dev)
N = int(os.environ.get("TOKENS", "140"))
kw = dict(do_sample=True, top_p=0.85, top_k=1, temperature=0.85, repetition_penalty=2.0, ignore_eos=True, max_new_tokens=N, stream_chunk_size=8)
best = 1e9
for rep in range(5):
    eng.timing = []
    fake = g.compute_embeddings(cond, codes)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in g.get_generator(fake_inputs=fake, **kw):
        pass
    b.record()
    torch.cuda.synchronize()
    t = sum(x.elapsed_time(y) for x, y, _, _ in eng.timing) / sum(n for _, _, n, _ in eng.timing)
    best = min(best, t)
    total = a.elapsed_time(b)
print(json.dumps({"vw": os.environ.get("GENVC_VW", "0"), "tokens": N, "decode_ms_per_token_best": round(best, 4), "segment_ms_last": round(total, 2)}))
