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

# trace slots per layer (include/genvc_b200.h)
TPL = 28
SEGMENTS = [
    ("x2 hop + load", None, 0), ("LN1", 0, 1), ("QKV gemv", 1, 2), ("attention item", 2, 3), ("AO hop + merge", 3, 4),
    ("PROJ gemv", 4, 5), ("x1 hop + load", 5, 6), ("LN2", 6, 7), ("FC gemv", 7, 8), ("P2 gemv", 8, 9), ("PP hop", 9, 10),
    ("gather + reduce", 10, 11),
]


def analyse(tr: torch.Tensor, L: int) -> dict:
    """Per segment: median over CTAs of the segment duration, averaged over layers; plus the critical path
    (latest CTA) view: time between the latest stamp of the segment's end and the latest of its start."""
    tr = tr.cpu().double()  # [G, slots] ns
    G = tr.shape[0]
    start = tr[:, L * TPL + 4]
    t0 = float(start.min())
    out = {"grid": G, "segments": {}}
    acc = {name: {"med": [], "max": [], "crit": []} for name, _, _ in SEGMENTS}
    for l in range(L):
        for name, a, b in SEGMENTS:
            if a is None:
                if l == 0:
                    continue
                ta = tr[:, (l - 1) * TPL + 11]
                ta = torch.where(ta > 0, ta, tr[:, (l - 1) * TPL + 9])  # non-reducer CTAs skip slots 10, 11
            else:
                ta = tr[:, l * TPL + a]
            tb = tr[:, l * TPL + b]
            ok = (ta > 0) & (tb > 0)
            if not ok.any():
                continue
            d = (tb - ta)[ok]
            acc[name]["med"].append(float(d.median()))
            acc[name]["max"].append(float(d.max()))
            acc[name]["crit"].append(float(tb[ok].max() - ta[ok].max()))
    for name, a in acc.items():
        out["segments"][name] = {k: round(statistics.mean(v), 1) if v else 0.0 for k, v in a.items()}
    head = L * TPL
    ww = tr[:, : L * TPL].view(G, L, TPL)
    out["weight_wait_ns"] = {"QKV+PROJ med": float(ww[:, :, 12].mean(1).median()), "QKV+PROJ max": float(ww[:, :, 12].mean(1).max()),
                             "FC+P2 med": float(ww[:, :, 13].mean(1).median()), "FC+P2 max": float(ww[:, :, 13].mean(1).max())}
    out["head"] = {
        "x2 hop + load": float((tr[:, head + 0] - torch.where(tr[:, head - 3] > 0, tr[:, head - 3], tr[:, head - 5])).median()),
        "2xLN + head gemv": float((tr[:, head + 1] - tr[:, head + 0]).median()),
        "logits hop + load": float((tr[:, head + 2] - tr[:, head + 1]).median()),
        "sample": float((tr[:, head + 3] - tr[:, head + 2]).median()),
    }
    # projected-value variant: stamps 14..19 are absolute times inside the PROJ section
    pv = tr[:, : L * TPL].view(G, L, TPL)[:, 1:, :]
    if float(pv[:, :, 14].max()) > 1e15:  # (globaltimer values, not cycle counts)
        names = ["score item->xq hop seen", "v load", "score hop wait", "per-head proj gemv", "scores validated+max", "exp+sum+scale", "barrier",
                 "8 fma", "new position", "barrier2", "reduce+store"]
        idx = [(3, 14), (14, 15), (15, 16), (16, 4), (4, 20), (20, 21), (21, 18), (18, 22), (22, 23), (23, 19), (19, 5)]
        out["pvw_ns"] = {n: float((pv[:, :, b] - pv[:, :, a]).median()) for n, (a, b) in zip(names, idx)}
    att = ww[:, 1:, 14:20]
    sel = att[:, :, 0] > 0
    out["att_cycles"] = [float(att[:, :, k][sel].median()) if sel.any() else 0.0 for k in range(6)]
    out["layer_ns"] = float((tr[:, (L - 1) * TPL].max() - tr[:, 0].max()) / max(L - 1, 1)) if L > 1 else 0.0
    out["step_ns"] = float(tr[:, head + 3].max() - t0)
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
    ap.add_argument("--ahead", type=int, nargs="*", default=[-1])
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
    for win, ahead in [(w, a) for w in args.window for a in args.ahead]:
        eng.tune(window=win, nosync=bool(args.nosync), l2_ahead=ahead)
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
                res[f"win{win}_ahead{ahead}_run{rep}"] = analyse(tr, args.layers)
                res[f"win{win}_ahead{ahead}_run{rep}"]["segment_ms"] = t0.elapsed_time(t1)
    eng.trace(None)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    for k, v in res.items():
        print(k, "step_us", round(v["step_ns"] / 1e3, 1), "layer_us", round(v["layer_ns"] / 1e3, 2), "segment_ms",
              round(v["segment_ms"], 3))
        for name, a in v["segments"].items():
            print(f"  {name:20s} med {a['med']/1e3:6.2f} us  max {a['max']/1e3:6.2f}  critical-path {a['crit']/1e3:6.2f}")
        print("  head:", {n: round(x / 1e3, 2) for n, x in v["head"].items()})
        if "pvw_ns" in v:
            print("  projected-value section (median over CTAs and layers, us):", {n: round(x / 1e3, 2) for n, x in v["pvw_ns"].items()})
        print("  attention item cycles (prefetch issue, poll until q seen, dots+shuffles, softmax weights+PV, later batches, barrier+merge+store):", [int(x) for x in v["att_cycles"]])
        print("  weight wait per layer (thread 0):", {n: int(x) for n, x in v["weight_wait_ns"].items()}, "cycles")


if __name__ == "__main__":
    main()
