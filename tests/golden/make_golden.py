"""Generate golden fixtures by running the REAL reference modules on CPU.

Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_golden.py [--only NAME] [--skip-full]

Each fixture ``tests/golden/<name>.pt`` holds the recipe for a seeded synthetic
checkpoint (regenerated at test time by ``genvc_b200.synth``; a digest guards
against RNG drift), the synthetic inputs, and the outputs of the reference's own
``GPT`` / ``GPT2InferenceModel`` / ``PerceiverResampler`` code driven by
``oracle/ref_shim.py::ref_generate`` (the restated HF-4.33 ``sample`` loop using
HF's own logits processors).  The reference repo has no tests or golden vectors
of its own, so these are the pin.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from genvc_b200.synth import state_dict_digest, synth_checkpoint  # noqa: E402
from oracle import ref_shim  # noqa: E402

# name -> recipe
CASES = {
    # toy models: every step's logits/latents kept
    "toy_d128_greedy": dict(model=dict(n_layer=2, d_model=128, n_head=4, seed=3), T=13, S_mel=120, B=1,
                            top_k=1, top_p=0.85, new_tokens=48, keep_all=True),
    "toy_d128_topk20": dict(model=dict(n_layer=2, d_model=128, n_head=4, seed=3), T=13, S_mel=120, B=1,
                            top_k=20, top_p=0.85, new_tokens=48, keep_all=True, noise_seed=77),
    "toy_d128_topk0_topp1": dict(model=dict(n_layer=2, d_model=128, n_head=4, seed=3), T=9, S_mel=64, B=1,
                                 top_k=0, top_p=1.0, new_tokens=24, keep_all=True, noise_seed=78),
    "toy_d128_eos": dict(model=dict(n_layer=2, d_model=128, n_head=4, seed=3, eos_bias=1.0), T=13, S_mel=120, B=1,
                         top_k=1, top_p=0.85, new_tokens=None, keep_all=True),
    "toy_d128_batch3": dict(model=dict(n_layer=2, d_model=128, n_head=4, seed=3, eos_bias=1.5), T=11, S_mel=90, B=3,
                            top_k=1, top_p=0.85, new_tokens=64, keep_all=True),
    "toy_d256_h4_greedy": dict(model=dict(n_layer=3, d_model=256, n_head=4, seed=4), T=38, S_mel=282, B=1,
                               top_k=1, top_p=0.85, new_tokens=40, keep_all=True),
    "toy_d512_h2_greedy": dict(model=dict(n_layer=2, d_model=512, n_head=2, seed=5), T=13, S_mel=100, B=1,
                               top_k=1, top_p=0.85, new_tokens=24, keep_all=True),
    # BASELINE.json configs[0]: GenVC_small greedy, 3 s src + 3 s ref, full run to the 602 cap / EOS
    "full_h4_cfg1": dict(model=dict(n_layer=30, d_model=1024, n_head=4, seed=1234), T=38, S_mel=282, B=1,
                         top_k=1, top_p=0.85, new_tokens=None, keep_all=False, full=True),
    "full_h16_greedy": dict(model=dict(n_layer=30, d_model=1024, n_head=16, seed=1234), T=13, S_mel=282, B=1,
                            top_k=1, top_p=0.85, new_tokens=64, keep_all=False, full=True),
    "full_h4_topk20": dict(model=dict(n_layer=30, d_model=1024, n_head=4, seed=1234), T=75, S_mel=469, B=1,
                           top_k=20, top_p=0.85, new_tokens=48, keep_all=False, full=True, noise_seed=79),
    # ---- round 2: batches (BASELINE configs[2]/[3]: equal-T rows, shared position index, finished rows padded) ----
    "toy_d128_batch8_eos": dict(model=dict(n_layer=2, d_model=128, n_head=4, seed=3, eos_bias=1.5), T=11, S_mel=90, B=8,
                                top_k=1, top_p=0.85, new_tokens=96, keep_all=True),
    "toy_d256_batch5_topk20": dict(model=dict(n_layer=3, d_model=256, n_head=4, seed=4), T=20, S_mel=100, B=5,
                                   top_k=20, top_p=0.85, new_tokens=40, keep_all=True, noise_seed=80),
    "full_h4_batch4_greedy": dict(model=dict(n_layer=30, d_model=1024, n_head=4, seed=1234), T=25, S_mel=282, B=4,
                                  top_k=1, top_p=0.85, new_tokens=20, keep_all=False, full=True),
    "full_h4_batch8_topk20": dict(model=dict(n_layer=30, d_model=1024, n_head=4, seed=1234), T=50, S_mel=469, B=8,
                                  top_k=20, top_p=0.85, new_tokens=40, keep_all=False, full=True, noise_seed=81),
    "full_h16_batch8_greedy": dict(model=dict(n_layer=30, d_model=1024, n_head=16, seed=1234), T=13, S_mel=282, B=8,
                                   top_k=1, top_p=0.85, new_tokens=24, keep_all=False, full=True),
    # shipped sampling defaults (top_k 15) on the 6 s reference chunk the pipeline produces (563 mel frames)
    "full_h4_topk15_mel563": dict(model=dict(n_layer=30, d_model=1024, n_head=4, seed=1234), T=75, S_mel=563, B=1,
                                  top_k=15, top_p=0.85, new_tokens=40, keep_all=False, full=True, noise_seed=82),
    # "GenVC_large": same dimensions (the README distinguishes the two checkpoints by training data only), second seed
    "full_large_seed4321": dict(model=dict(n_layer=30, d_model=1024, n_head=4, seed=4321), T=38, S_mel=282, B=2,
                                top_k=1, top_p=0.85, new_tokens=32, keep_all=False, full=True),
}


def make_inputs(case):
    g = torch.Generator().manual_seed(7)
    codes = torch.randint(0, 256, (case["B"], case["T"]), generator=g)
    g = torch.Generator().manual_seed(11)
    mel = torch.randn((case["B"], 80, case["S_mel"]), generator=g)
    return codes, mel


@torch.no_grad()
def run_case(name, case):
    t0 = time.time()
    ck = synth_checkpoint(**case["model"])
    ref = ref_shim.build_reference_gpt(ck)
    codes, mel = make_inputs(case)
    style = ref.get_style_emb(mel)  # [B, D, 32]
    cond = style.transpose(1, 2).contiguous()
    V = ck["config"]["model_args"]["gpt_num_audio_tokens"]
    cap = ref.max_gen_mel_tokens if case["new_tokens"] is None else case["new_tokens"]
    noise = None
    if case.get("noise_seed") is not None:
        ng = torch.Generator().manual_seed(case["noise_seed"])
        noise = torch.empty((cap, case["B"], V)).exponential_(1, generator=ng)
    trace = {}
    ids, lats = ref_shim.ref_generate(
        ref, cond, codes, case["top_k"], case["top_p"], 0.85, 2.0,
        noise=noise, max_new_tokens=case["new_tokens"], trace=trace,
    )
    n = ids.shape[1]
    logits = torch.stack(trace["logits"], 1)  # [B, n, V]
    scores = torch.stack(trace["scores"], 1)
    # top-2 gap of the repetition-penalised, temperature-scaled logits (before top-k/top-p):
    # how far greedy decoding is from flipping under fp32 reorder noise
    from oracle.genvc_oracle import SamplingParams, process_logits
    P = 32 + codes.shape[1] + 2
    fake = torch.full((case["B"], P + 1), 1, dtype=torch.long)
    fake[:, -1] = ref.start_audio_token
    gaps = []
    for s in range(n):
        pen = process_logits(torch.cat([fake, ids[:, :s]], 1), logits[:, s], SamplingParams(top_k=0, top_p=1.0))
        t2 = torch.topk(pen, 2, dim=-1)[0]
        gaps.append(t2[..., 0] - t2[..., 1])
    gap = torch.stack(gaps, 1)
    if case["keep_all"]:
        steps = list(range(n))
    else:
        steps = sorted(set(list(range(min(n, 32))) + list(range(0, n, 16)) + [n - 1]))
    # second (latent) pass on row 0's non-stop tokens, like synthesize_utt does
    g0 = ids[0][ids[0] != ref.stop_audio_token]
    lat_pass = None
    if g0.numel() > 1:
        lat_pass = ref(codes[:1], torch.tensor([codes.shape[1]]), g0[None],
                       torch.tensor([g0.numel() * ref.code_stride_len]), cond_latents=cond[:1], return_latent=True)
    fx = dict(
        name=name,
        model=case["model"],
        digest=state_dict_digest(ck["model"]),
        sampling=dict(top_k=case["top_k"], top_p=case["top_p"], temperature=0.85, repetition_penalty=2.0),
        new_tokens=case["new_tokens"],
        noise_seed=case.get("noise_seed"),
        codes=codes,
        mel=mel,
        style_emb=style.clone(),
        ids=ids,
        steps=torch.tensor(steps),
        logits=logits[:, steps].clone(),
        latents=lats[:, steps].clone(),
        gap=gap.clone(),
        latent_pass=None if lat_pass is None else (lat_pass.clone() if case["keep_all"] else lat_pass[:, steps[: min(len(steps), g0.numel())]].clone()),
        latent_pass_steps=None if (lat_pass is None or case["keep_all"]) else torch.tensor(steps[: min(len(steps), g0.numel())]),
        versions=dict(torch=torch.__version__, transformers=__import__("transformers").__version__),
    )
    path = os.path.join(HERE, name + ".pt")
    torch.save(fx, path)
    print(f"{name}: n={n} ids[:8]={ids[0,:8].tolist()} last={ids[0,-1].item()} min_gap={gap.min().item():.3e} "
          f"size={os.path.getsize(path)/1e3:.0f} kB  ({time.time()-t0:.1f}s)", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--skip-full", action="store_true")
    a = ap.parse_args()
    assert ref_shim.reference_available(), "needs /root/reference"
    torch.set_num_threads(os.cpu_count())
    for name, case in CASES.items():
        if a.only and a.only != name:
            continue
        if a.skip_full and case.get("full"):
            continue
        run_case(name, case)


if __name__ == "__main__":
    main()
