"""``model_init(checkpoint_path, device) -> (model, config)`` with the reference's contract
(``inference/model_init.py:9-34``) for the codec-token path.

The checkpoint layout is the reference's: ``torch.load(path)`` -> ``{"config": nested dict,
"model": flat state_dict}``; the GPT's tensors live under the ``gpt.`` prefix and are loaded
non-strictly (everything else in the file — DVAEs, HiFi-GAN, discriminators — is ignored here).
``config`` is returned as a mutable attribute-dict so ``model.config.top_k = ...`` (``infer.py:22``)
keeps working without coqpit.

Stages outside this path (ContentVec, content-DVAE, mel front-end, HiFi-GAN; SURVEY.md §2 #9-#11,
#16) are *attachment points* on the returned model: assign the reference's own modules to
``model.content_extractor``, ``model.content_dvae`` and ``model.hifigan`` (the last two are built on the CUDA library
automatically when the checkpoint holds ``content_dvae.*`` / ``hifigan.*`` weights: ``genvc_b200/content_dvae.py``,
``genvc_b200/vocoder.py``) and
``model.torch_mel_spectrogram_style_encoder`` to run the full pipeline.
"""
from __future__ import annotations

from typing import Optional

import torch

from ..config import GenVCDims, wrap_config
from ..gpt import GPT


class _Unattached:
    def __init__(self, name: str):
        self._name = name

    def __getattr__(self, item):
        raise RuntimeError(
            f"model.{self._name} is outside the genvc_b200 path and has not been attached; assign the "
            f"reference module to model.{self._name} (see INTEGRATION.md)"
        )

    def __call__(self, *a, **k):
        self.__getattr__("__call__")


def _cfg_get(cfg, path, default):
    """Nested lookup in an attribute-dict / dict config; ``default`` when any key is absent."""
    cur = cfg
    for key in path:
        if cur is None:
            return default
        if isinstance(cur, dict):
            cur = cur.get(key)
        else:
            cur = getattr(cur, key, None)
    return default if cur is None else cur


class GenVCModel:
    """The slice of ``trainers/hifigan_trainer.py::HiFiGANTrainer`` the inference drivers touch."""

    def __init__(self, config, device, max_batch: int = 1, max_mel_frames: int = 576):
        self.config = config
        self.dims = GenVCDims.from_config(config)
        self.device = torch.device(device)
        self.gpt = GPT(self.dims, device=device, max_batch=max_batch, max_mel_frames=max_mel_frames)
        # trainers/hifigan_trainer.py:56: latent upsampling = gpt_code_stride_len / vocoder hop length;
        # :146: content sample rate = content_dvae_config.audio.dvae_sample_rate.  Read from the checkpoint's config;
        # the shipped values (1024 / 256 = 4; 16 kHz) only when the keys are absent.
        self.hifigan_scale_factor = _cfg_get(config, ("vocoder_config", "hop_length"), None)
        if self.hifigan_scale_factor is not None:
            self.hifigan_scale_factor = self.dims.code_stride_len / self.hifigan_scale_factor
            if float(self.hifigan_scale_factor).is_integer():
                self.hifigan_scale_factor = int(self.hifigan_scale_factor)
        else:
            self.hifigan_scale_factor = 4
        self.content_sample_rate = int(_cfg_get(config, ("content_dvae_config", "audio", "dvae_sample_rate"), 16000))
        self.content_extractor = _Unattached("content_extractor")
        self.content_dvae = _Unattached("content_dvae")
        self.hifigan = _Unattached("hifigan")
        self.torch_mel_spectrogram_style_encoder = _Unattached("torch_mel_spectrogram_style_encoder")

    def load_state_dict(self, state_dict, strict: bool = False):
        self.gpt.load_state_dict(state_dict, strict=strict)
        return self

    @torch.inference_mode()
    def get_gpt_cond_latents(self, audio: torch.Tensor, sr: int, length: int = 30, chunk_length: int = 6) -> torch.Tensor:
        """Reference audio [1, samples] -> speaker latents [1, 32, D]: clip to ``length`` s, one perceiver
        pass per ``chunk_length``-s chunk (chunks under 0.33 s dropped), mean over chunks
        (``trainers/hifigan_trainer.py:438-455``)."""
        audio = audio[:, : sr * length]
        mels = []
        for i in range(0, audio.shape[1], sr * chunk_length):
            chunk = audio[:, i: i + sr * chunk_length]
            if chunk.size(-1) < sr * 0.33:
                continue
            mels.append(self.torch_mel_spectrogram_style_encoder(chunk.unsqueeze(0)))
        return self.get_gpt_cond_latents_from_mels(mels)

    @torch.inference_mode()
    def inference(self, src_audio: torch.Tensor, cond_latent: torch.Tensor, do_sample: bool = True, top_p: float = 0.85,
                  top_k: int = 15, temperature: float = 0.75, num_beams: int = 1, length_penalty: float = 1.0,
                  repetition_penalty: float = 10.0, output_attentions: bool = False,
                  reuse_decode_latents: bool = False) -> torch.Tensor:
        """One source segment -> waveform (``trainers/hifigan_trainer.py:458-504``): content codes, ``gpt.generate``,
        EOS stripped, latents of the generated codes, x``hifigan_scale_factor`` linear interpolation, vocoder.
        ``reuse_decode_latents``: take the latents the decode kernel emitted instead of the teacher-forced second
        pass (equal up to fp32 rounding)."""
        feat = self.content_extractor.extract_content_features(src_audio)
        codes = self.content_dvae.get_codebook_indices(feat.transpose(1, 2))
        gen = self.gpt.generate(cond_latent, codes, do_sample=do_sample, top_p=top_p, top_k=top_k, temperature=temperature,
                                num_beams=num_beams, length_penalty=length_penalty, repetition_penalty=repetition_penalty,
                                output_attentions=output_attentions)[0]
        keep = (gen != self.gpt.stop_audio_token).nonzero().squeeze()
        if reuse_decode_latents:
            lat = self.gpt.last_latents[0][keep].reshape(1, -1, self.gpt.last_latents.shape[-1])
        else:
            gen = gen[keep]
            out_len = torch.tensor([gen.shape[-1] * self.config.model_args.gpt_code_stride_len], device=self.device)
            content_len = torch.tensor([codes.shape[-1]], device=self.device)
            lat = self.gpt(codes, content_len, gen.unsqueeze(0), out_len, cond_latents=cond_latent, return_latent=True)
        mel_input = torch.nn.functional.interpolate(lat.transpose(1, 2), scale_factor=[self.hifigan_scale_factor],
                                                    mode="linear").squeeze(1)
        return self.hifigan(mel_input)

    @torch.inference_mode()
    def get_gpt_cond_latents_from_mels(self, mel_chunks) -> torch.Tensor:
        """Same, entered after the mel front-end: list of [B, 80, S_i] (or [B, 1, 80, S_i])."""
        if not mel_chunks:
            raise ValueError("no reference chunk of at least 0.33 s")
        embs = [self.gpt.get_style_emb(m.to(self.device), None) for m in mel_chunks]
        return torch.stack(embs).mean(dim=0).transpose(1, 2)


@torch.inference_mode()
def model_init(checkpoint_path, device, max_batch: int = 1, max_mel_frames: int = 576):
    ckpt_states = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
    return model_from_checkpoint(ckpt_states, device, max_batch=max_batch, max_mel_frames=max_mel_frames)


def model_from_checkpoint(ckpt_states: dict, device, max_batch: int = 1, max_mel_frames: int = 576,
                          blob: Optional[torch.Tensor] = None):
    """``blob``: an already packed weight blob (ranks > 0 receive it by NCCL broadcast instead of
    re-reading the checkpoint; see ``genvc_b200.replicas``)."""
    config = wrap_config(ckpt_states["config"])
    config.is_inference = True
    model = GenVCModel(config, device, max_batch=max_batch, max_mel_frames=max_mel_frames)
    if blob is not None:
        model.gpt.load_blob(blob)
    else:
        model.load_state_dict(ckpt_states["model"], strict=False)
    model.gpt.eval()
    model.gpt.to(device)
    model.gpt.init_gpt_for_inference()
    attach_cuda_hifigan(model, ckpt_states.get("model") or {}, ckpt_states["config"])
    attach_cuda_content_dvae(model, ckpt_states.get("model") or {}, ckpt_states["config"])
    attach_cuda_mel_frontend(model, config)
    return model, config


def attach_cuda_mel_frontend(model, config) -> bool:
    """The style-encoder mel front-end (``trainers/hifigan_trainer.py:105-115``) on the CUDA library when its normalisation
    file is available: ``config.model_args.mel_norm_file`` is a readable path (or explicitly None = no normalisation).  A
    checkpoint whose path does not exist here keeps the attachment point."""
    import os

    ma = getattr(config, "model_args", None)
    if ma is None or "mel_norm_file" not in ma or torch.device(model.device).type != "cuda":
        return False
    path = ma["mel_norm_file"]
    if path is not None and not (isinstance(path, str) and os.path.exists(path)):
        return False
    from ..mel import TorchMelSpectrogram

    model.torch_mel_spectrogram_style_encoder = TorchMelSpectrogram(
        filter_length=2048, hop_length=256, win_length=1024, normalize=False, sampling_rate=int(config.audio.sample_rate),
        mel_fmin=0, mel_fmax=8000, n_mel_channels=80, mel_norm_file=path, device=model.device)
    return True


def attach_cuda_content_dvae(model, state_dict: dict, raw_config) -> bool:
    """``content_dvae.*`` weights in the checkpoint -> the tokeniser runs on the CUDA library (``genvc_b200/content_dvae.py``,
    built from ``config.content_dvae_config`` like ``trainers/hifigan_trainer.py:149-160``)."""
    sd = {k[len("content_dvae."):]: v for k, v in state_dict.items() if k.startswith("content_dvae.")}
    if "codebook.embed" not in sd or torch.device(model.device).type != "cuda":
        return False
    from ..content_dvae import DiscreteVAE

    dc = raw_config.get("content_dvae_config", {}) if isinstance(raw_config, dict) else getattr(raw_config, "content_dvae_config", {})
    model.content_dvae = DiscreteVAE.from_config(dc or {}, device=model.device).load_state_dict(sd)
    return True


def attach_cuda_hifigan(model, state_dict: dict, raw_config) -> bool:
    """If the checkpoint carries the generator's weights (``hifigan.*``, as saved by ``trainers/hifigan_trainer.py``), the
    vocoder runs on the CUDA library too (``genvc_b200/vocoder.py``, built from ``config.vocoder_config`` like
    ``trainers/hifigan_trainer.py:47-55``); otherwise ``model.hifigan`` stays an attachment point."""
    sd = {k[len("hifigan."):]: v for k, v in state_dict.items() if k.startswith("hifigan.")}
    if not any(k.startswith("conv_pre.") for k in sd) or torch.device(model.device).type != "cuda":
        return False
    from ..vocoder import HiFiGAN

    vc = raw_config.get("vocoder_config", {}) if isinstance(raw_config, dict) else getattr(raw_config, "vocoder_config", {})
    model.hifigan = HiFiGAN.from_config(vc or {}, device=model.device).load_state_dict(sd)
    return True
