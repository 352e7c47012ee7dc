"""Model dimensions and inference knobs read from a GenVC checkpoint's ``config`` dict.

The reference turns ``ckpt["config"]`` into Coqpit dataclasses
(``inference/model_init.py:11-12``); coqpit is not a dependency here, so the same
fields are read by plain dict access and exposed with attribute syntax, keeping
``model.config.top_k = ...`` (``infer.py:22``) working.

Field names follow ``configs/genVC_configs.py:127-139`` (model_args.gpt_*) and
``configs/genVC_train_configs.py:76-80`` (top_k / top_p / temperature /
length_penalty / repetition_penalty).
"""
from __future__ import annotations

from dataclasses import dataclass


class AttrDict(dict):
    """dict with attribute access, recursively (mutable, like a Coqpit object)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover - error path
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    @staticmethod
    def wrap(obj):
        if isinstance(obj, dict):
            return AttrDict({k: AttrDict.wrap(v) for k, v in obj.items()})
        if isinstance(obj, list):
            return [AttrDict.wrap(v) for v in obj]
        return obj


# Defaults of the reference dataclasses, used when a key is absent from the checkpoint.
MODEL_ARG_DEFAULTS = dict(
    gpt_max_audio_tokens=605,
    gpt_max_text_tokens=402,
    gpt_max_prompt_tokens=70,
    gpt_layers=30,
    gpt_n_model_channels=1024,
    gpt_n_heads=16,
    gpt_number_text_tokens=258,
    gpt_start_text_token=256,
    gpt_stop_text_token=257,
    gpt_num_audio_tokens=1026,
    gpt_start_audio_token=1024,
    gpt_stop_audio_token=1025,
    gpt_code_stride_len=1024,
)
TOP_LEVEL_DEFAULTS = dict(
    temperature=0.85,
    length_penalty=1.0,
    repetition_penalty=2.0,
    top_k=15,
    top_p=0.85,
)

# Perceiver hyper-parameters are hard-coded in the reference (layers/gpt.py:179-188).
PERCEIVER = dict(depth=4, dim_context=80, num_latents=32, dim_head=64, heads=8, ff_mult=4)


@dataclass(frozen=True)
class GenVCDims:
    """Everything the kernels need to know about the model's shape."""

    n_layer: int
    d_model: int
    n_head: int
    n_text_vocab: int
    n_audio_vocab: int
    start_text: int
    stop_text: int
    start_audio: int
    stop_audio: int
    max_audio_tokens: int  # config value (605); position table has +3 rows
    max_text_tokens: int  # config value (402); position table has +2 rows
    max_prompt_tokens: int
    code_stride_len: int
    # perceiver
    pc_depth: int = 4
    pc_dim_context: int = 80
    pc_latents: int = 32
    pc_dim_head: int = 64
    pc_heads: int = 8
    pc_ff_mult: int = 4

    @property
    def head_dim(self) -> int:
        return self.d_model // self.n_head

    @property
    def n_mel_pos(self) -> int:  # layers/gpt.py:132  max_mel_tokens + 2 + max_conditioning_inputs
        return self.max_audio_tokens + 3

    @property
    def n_text_pos(self) -> int:  # layers/gpt.py:133
        return self.max_text_tokens + 2

    @property
    def max_gen_mel_tokens(self) -> int:  # layers/gpt.py:131
        return self.max_audio_tokens - 1 - 2

    @property
    def pc_inner(self) -> int:  # attention inner dim
        return self.pc_dim_head * self.pc_heads

    @property
    def pc_ff_inner(self) -> int:  # layers/perceiver_encoder.py:211
        return int(self.d_model * self.pc_ff_mult * 2 / 3)

    @property
    def max_seq(self) -> int:
        """Largest sequence the decoder can ever see (layers/gpt.py:56-59, 198)."""
        return self.n_mel_pos + self.n_text_pos + self.max_prompt_tokens + 1

    @staticmethod
    def from_config(cfg: dict) -> "GenVCDims":
        ma = dict(MODEL_ARG_DEFAULTS)
        for k, v in (cfg.get("model_args") or {}).items():
            if k in ma and v is not None:
                ma[k] = v
        return GenVCDims(
            n_layer=int(ma["gpt_layers"]),
            d_model=int(ma["gpt_n_model_channels"]),
            n_head=int(ma["gpt_n_heads"]),
            n_text_vocab=int(ma["gpt_number_text_tokens"]),
            n_audio_vocab=int(ma["gpt_num_audio_tokens"]),
            start_text=int(ma["gpt_start_text_token"]),
            stop_text=int(ma["gpt_stop_text_token"]),
            start_audio=int(ma["gpt_start_audio_token"]),
            stop_audio=int(ma["gpt_stop_audio_token"]),
            max_audio_tokens=int(ma["gpt_max_audio_tokens"]),
            max_text_tokens=int(ma["gpt_max_text_tokens"]),
            max_prompt_tokens=int(ma["gpt_max_prompt_tokens"]),
            code_stride_len=int(ma["gpt_code_stride_len"]),
            pc_depth=PERCEIVER["depth"],
            pc_dim_context=PERCEIVER["dim_context"],
            pc_latents=PERCEIVER["num_latents"],
            pc_dim_head=PERCEIVER["dim_head"],
            pc_heads=PERCEIVER["heads"],
            pc_ff_mult=PERCEIVER["ff_mult"],
        )


def make_config_dict(n_layer=30, d_model=1024, n_head=4, **overrides) -> dict:
    """A checkpoint ``config`` dict in the reference's nested layout."""
    ma = dict(MODEL_ARG_DEFAULTS)
    ma.update(gpt_layers=n_layer, gpt_n_model_channels=d_model, gpt_n_heads=n_head)
    top = dict(TOP_LEVEL_DEFAULTS)
    for k, v in overrides.items():
        if k in ma:
            ma[k] = v
        else:
            top[k] = v
    cfg = dict(top)
    cfg["model_args"] = ma
    cfg["audio"] = {"sample_rate": 24000}
    return cfg


def wrap_config(cfg: dict) -> AttrDict:
    """The mutable ``config`` object handed back by ``model_init``."""
    out = AttrDict.wrap(dict(cfg))
    for k, v in TOP_LEVEL_DEFAULTS.items():
        out.setdefault(k, v)
    ma = out.setdefault("model_args", AttrDict())
    for k, v in MODEL_ARG_DEFAULTS.items():
        if ma.get(k) is None:
            ma[k] = v
    out.setdefault("audio", AttrDict({"sample_rate": 24000}))
    return out
