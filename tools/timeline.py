#!/usr/bin/env python
"""Phase timeline of the fused decode kernel (debug hook genvc_debug_trace): where a decode step's
time goes — GEMV/attention compute vs. grid-exchange wait — per phase, median over layers and CTAs.

    python tools/timeline.py [--heads 4] [--T 13] [--out gpurun_out/timeline.json]
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

PHASES = ["QKV", "ATT", "PROJ", "FC", "PROJ2"]


def analyse(tr: torch.Tensor, L: int) -> dict:
    tr = tr.cpu().double()  # [G, slots] ns
    G = tr.shape[0]
    start = tr[:, L * 10 + 3]
    t0 = float(start.min())
    out = {"grid": G, "phases": {}, "layers": []}
    acc = {p: {"compute_med": [], "compute_max": [], "wait_med": [], "skew": [], "span": []} for p in PHASES + ["HEAD"]}
    prev_end = start.clone()
    for l in range(L + 1):
        names = PHASES if l < L else ["HEAD"]
        for k, name in enumerate(names):
            ce = tr[:, l * 10 + 2 * k]
            be = tr[:, l * 10 + 2 * k + 1]
            comp = ce - prev_end
            wait = be - ce
            a = acc[name]
            a["compute_med"].append(float(comp.median()))
            a["compute_max"].append(float(comp.max()))
            a["wait_med"].append(float(wait.median()))
            a["skew"].append(float(ce.max() - ce.min()))
            a["span"].append(float(be.max() - prev_end.max()))
            prev_end = be
    for name, a in acc.items():
        out["phases"][name] = {k: round(statistics.mean(v), 1) for k, v in a.items()}
        out["phases"][name]["n"] = len(a["span"])
    sample_end = tr[:, L * 10 + 2]
    out["step_ns"] = float(sample_end.max() - t0)
    out["sample_ns"] = float((sample_end - prev_end).median())
    out["sum_span_ns"] = sum(sum(a["span"]) for a in acc.values())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--heads", type=int, default=4)
    ap.add_argument("--T", type=int, default=13)
    ap.add_argument("--layers", type=int, default=30)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--window", type=int, nargs="*", default=[0])
    ap.add_argument("--nosync", type=int, default=0)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline.json"))
    args = ap.parse_args()
    from genvc_b200.config import GenVCDims
    from genvc_b200.gpt import GPT
    from genvc_b200.synth import synth_checkpoint

    dev = torch.device("cuda:0")
    ck = synth_checkpoint(n_layer=args.layers, d_model=args.dim, n_head=args.heads, seed=1234)
    g = GPT(GenVCDims.from_config(ck["config"]), device=dev)
    g.load_state_dict(ck["model"])
    g.eval().to(dev).init_gpt_for_inference()
    eng = g.engine
    gen = torch.Generator().manual_seed(7)
    codes = torch.randint(0, 256, (1, args.T), generator=gen).to(dev)
    cond = torch.randn((1, 32, args.dim), generator=gen).to(dev)
    kw = dict(do_sample=True, top_p=0.85, top_k=1, temperature=0.85, repetition_penalty=2.0, ignore_eos=True,
              max_new_tokens=24, stream_chunk_size=8, decode_mode=args.mode)
    res = {}
    for win in args.window:
        eng.tune(window=win, nosync=bool(args.nosync))
        for rep in range(3):  # warm-up, then two traced runs (steps 2 and 6 of the 2nd launch)
            tr = eng.trace(step=(2 if rep < 2 else 6))
            fake = g.compute_embeddings(cond, codes)
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in g.get_generator(fake_inputs=fake, **kw):
                pass
            t1.record()
            torch.cuda.synchronize()
            if rep > 0:
                res[f"win{win}_run{rep}"] = analyse(tr, args.layers)
                tt = eng.tile_trace.cpu()
                for c in (0, 5, 77):
                    row = tt[c]
                    base = int(row[row > 0].min()) if (row > 0).any() else 0
                    res[f"win{win}_run{rep}"][f"tiles_cta{c}"] = [[int(v - base) if v > 0 else -1 for v in t] for t in row.tolist()]
                res[f"win{win}_run{rep}"]["segment_ms"] = t0.elapsed_time(t1)
    eng.trace(None)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    for k, v in res.items():
        print(k, "step_us", round(v["step_ns"] / 1e3, 1), "sample_us", round(v["sample_ns"] / 1e3, 2), "segment_ms", round(v["segment_ms"], 3))
        for c in (0, 5, 77):
            if f"tiles_cta{c}" in v:
                print(f"  tiles cta{c} (issue, wait_begin, wait_end ns):", " ".join(f"({a},{b},{d})" for a, b, d in v[f"tiles_cta{c}"] if a >= 0 or b >= 0))
        for name, a in v["phases"].items():
            print(f"  {name:6s} span {a['span']/1e3:7.2f} us  compute med/max {a['compute_med']/1e3:6.2f}/{a['compute_max']/1e3:6.2f}"
                  f"  wait med {a['wait_med']/1e3:6.2f}  skew {a['skew']/1e3:6.2f}")


if __name__ == "__main__":
    main()
