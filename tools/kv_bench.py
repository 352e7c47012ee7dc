#!/usr/bin/env python
"""KV-cache attention microbenchmark (BASELINE.json configs[4]): single-token attention
over S cached keys, d_model = 1024 (H x hd = 4 x 256 and 16 x 64), fp32 K/V, seq_len sweep 64 -> 4096.
>= 64 independent caches are cycled so the working set exceeds the 126 MB L2; reports achieved HBM GB/s
(algorithmic bytes 2*S*D*4 + 2*D*4 per cache) against the measured copy bandwidth.

    python tools/kv_bench.py [--out gpurun_out/kv_bench.json]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "kv_bench.json"))
    args = ap.parse_args()
    from genvc_b200.config import GenVCDims, make_config_dict
    from genvc_b200.engine import Engine

    dev = torch.device("cuda:0")
    eng = Engine(GenVCDims.from_config(make_config_dict(2, 128, 4)), dev)  # only the library handle is needed
    peak = 6532.9
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    D = 1024
    rows = []
    for H, hd in ((4, 256), (16, 64)):
        for S in (64, 128, 256, 512, 1024, 2048, 4096):
            per_cache = 2 * S * D * 4
            N = max(64, (768 << 20) // per_cache)  # >= 768 MB of K/V: far beyond L2
            N = min(N, 4096)
            k = torch.randn((N, H, S, hd), device=dev)
            v = torch.randn((N, H, S, hd), device=dev)
            q = torch.randn((N, H, hd), device=dev)
            out = eng.kv_attention(q, k, v, S)
            if S <= 256:  # spot check against torch on a few caches
                ref = torch.softmax(torch.einsum("nhd,nhsd->nhs", q[:4], k[:4]) / hd ** 0.5, -1)
                ref = torch.einsum("nhs,nhsd->nhd", ref, v[:4])
                assert torch.allclose(out[:4], ref, atol=2e-5, rtol=1e-4), float((out[:4] - ref).abs().max())
            for _ in range(3):
                eng.kv_attention(q, k, v, S)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps):
                eng.kv_attention(q, k, v, S)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            bytes_ = N * (per_cache + 2 * D * 4)
            gbs = bytes_ / (ms * 1e-3) / 1e9
            rows.append({"H": H, "hd": hd, "S": S, "caches": N, "ms": round(ms, 4), "GBps": round(gbs, 1), "frac": round(gbs / peak, 3)})
            print(rows[-1], flush=True)
            del k, v, q, out
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"peak_GBps": peak, "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
