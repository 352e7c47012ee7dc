"""Device engine: PyTorch owns the memory, ``libgenvc_b200.so`` does the work.

One ``Engine`` = one ``genvc_ctx`` on one GPU: the packed fp32 weight blob (reference
state-dict names and orientations), the re-tiled decode weight stream, a static KV cache
``[L][2][max_batch][H][max_seq][hd]`` and a workspace — all ``torch`` tensors whose
``data_ptr()`` is handed to the C ABI.  Every call enqueues on the current CUDA stream and
returns immediately; nothing here synchronises except where a result is read on the host.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .config import GenVCDims
from .lib import GenvcError, GenvcSampling, load_library
from .weights import c_config, pack_state_dict, tensor_table


@dataclass
class Sampling:
    """HF ``generate`` knobs the reference passes (inference/inference_utils.py:55-66, 170-182)."""

    top_k: int = 15
    top_p: float = 0.85
    temperature: float = 0.85
    repetition_penalty: float = 2.0
    ignore_eos: bool = False
    max_new_tokens: int = 0  # <= 0: the reference cap (max_gen_mel_tokens)
    seed: int = 0

    def to_c(self) -> GenvcSampling:
        top_p = 1.0 if self.top_p is None else float(self.top_p)
        return GenvcSampling(
            top_k=int(self.top_k or 0),
            top_p=top_p,
            # `cumulative_probs <= (1 - top_p)` is evaluated in fp32 by torch (HF TopPLogitsWarper)
            top_p_threshold=float(np.float32(1.0 - top_p)),
            temperature=float(1.0 if self.temperature is None else self.temperature),
            repetition_penalty=float(1.0 if self.repetition_penalty is None else self.repetition_penalty),
            ignore_eos=int(bool(self.ignore_eos)),
            max_new_tokens=int(self.max_new_tokens or 0),
            seed=int(self.seed) & 0xFFFFFFFFFFFFFFFF,
        )


@dataclass
class DecodeChunk:
    """Device-side outputs of one ``Engine.decode`` call (valid once the stream reaches it)."""

    ids: torch.Tensor  # [n_steps, B] int64
    latents: torch.Tensor  # [n_steps, B, D] fp32
    logits: Optional[torch.Tensor]  # [n_steps, B, V] raw logits, if requested
    status: torch.Tensor  # int32 [4] = {steps emitted by this call, done, out-of-range ids were clamped, reserved}


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Engine:
    def __init__(self, dims: GenVCDims, device, max_batch: int = 1, max_seq: Optional[int] = None,
                 max_mel_frames: int = 576):
        self.lib = load_library()
        self.dims = dims
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("genvc_b200 runs on CUDA devices only (no CPU fallback); got %s" % (self.device,))
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA is not available: genvc_b200 has no CPU fallback")
        self.dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", self.dev_index)
        if max_seq is None:
            max_seq = (dims.max_seq + 7) // 8 * 8
        self.max_batch, self.max_seq, self.max_mel_frames = int(max_batch), int(max_seq), int(max_mel_frames)
        self.cfg = c_config(dims, self.max_batch, self.max_seq, self.max_mel_frames)
        self._ctx = C.c_void_p()
        rc = self.lib.genvc_create(C.byref(self.cfg), self.dev_index, C.byref(self._ctx))
        if rc != 0:
            msg = self.lib.genvc_last_error(self._ctx).decode() if self._ctx else "genvc_create failed"
            if self._ctx:
                self.lib.genvc_destroy(self._ctx)
                self._ctx = C.c_void_p()
            raise GenvcError(rc, msg)
        self.blob: Optional[torch.Tensor] = None
        self.wstream: Optional[torch.Tensor] = None
        self.wtc: Optional[torch.Tensor] = None
        self.use_tensor_cores = os.environ.get("GENVC_TC", "1") != "0"
        with torch.cuda.device(self.device):
            self.kv = torch.zeros(int(self.lib.genvc_kv_floats(self._ctx)), dtype=torch.float32, device=self.device)
            self.ws = torch.zeros(int(self.lib.genvc_workspace_bytes(self._ctx)), dtype=torch.uint8, device=self.device)
        self._check(self.lib.genvc_bind_buffers(self._ctx, self.kv.data_ptr(), self.kv.numel(), self.ws.data_ptr(),
                                                self.ws.numel()))
        # projected-value cache variant of the single-row fused kernel: opt-in (GENVC_VW=1).  Since the GEMV weights are in
        # registers before their hop (gemv_preload) the K / V attention items win at every length measured: 6 s segment, 140
        # tokens, S up to 250: 0.404 ms/token against 0.509 with the cache, whose weighted sum reads (S - 1) x H rows of 32
        # bytes per CTA and only the first 256 of them are prefetched (tools/long_gen_bench.py).  1.07 GB at L=30, H=4.
        self.vw: Optional[torch.Tensor] = None
        n_vw = int(self.lib.genvc_vw_floats(self._ctx))
        if n_vw > 0 and os.environ.get("GENVC_VW", "0") == "1":
            with torch.cuda.device(self.device):
                self.vw = torch.zeros(n_vw, dtype=torch.float32, device=self.device)
            self._check(self.lib.genvc_bind_vw(self._ctx, self.vw.data_ptr(), n_vw))
        self._B = 0
        self._P = 0
        self._pending = False  # logits of the next step already computed (by prefill)
        self._n = 0  # host mirror of the number of tokens emitted since prefill (upper bound)
        self.validate_device_ids = True
        # bench hook: when a list, every decode() appends (start_event, end_event, n_forward, first_S)
        self.timing: Optional[list] = None

    # ------------------------------------------------------------------ plumbing
    def __del__(self):
        try:
            if getattr(self, "_ctx", None):
                if torch.cuda.is_available():
                    torch.cuda.synchronize(self.device)
                self.lib.genvc_destroy(self._ctx)
                self._ctx = C.c_void_p()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise GenvcError(rc, self.lib.genvc_last_error(self._ctx).decode())

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _dev(self, t: torch.Tensor, dtype, name: str) -> torch.Tensor:
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a tensor")
        if t.device != self.device:
            t = t.to(self.device, non_blocking=True)
        if t.dtype != dtype:
            t = t.to(dtype)
        return t.contiguous()

    def _check_ids(self, ids: torch.Tensor, vocab: int, what: str):
        """Range check of ids the kernels index embedding tables with.  Host tensors are checked here (free).
        Device tensors are NOT synchronised on: the kernels clamp an out-of-range id (no out-of-bounds read) and
        raise a flag that comes back with the next decode status (``DecodeChunk.status[2]``; ``GPT`` turns it
        into an ``IndexError``).  ``validate_device_ids = "sync"`` restores the blocking check."""
        if ids.numel() == 0 or (ids.is_cuda and self.validate_device_ids != "sync"):
            return
        lo, hi = int(ids.min()), int(ids.max())
        if lo < 0 or hi >= vocab:
            raise IndexError(f"{what} outside [0, {vocab}): min {lo}, max {hi}")

    @property
    def decode_grid(self) -> int:
        return int(self.lib.genvc_decode_grid(self._ctx))

    def fused_rows(self, B: int) -> bool:
        """True when a batch of ``B`` rows runs through a fused persistent decode kernel (decode_mega / decode_batch)."""
        return self.wstream is not None and 1 <= B <= int(self.lib.genvc_fused_max_rows(self._ctx))

    @property
    def launch_count(self) -> int:
        return int(self.lib.genvc_launch_count(self._ctx))

    # ------------------------------------------------------------------ weights
    def tensor_table(self) -> List[Tuple[str, int, int, int, int]]:
        return tensor_table(self.lib, self._ctx)

    @property
    def blob_floats(self) -> int:
        return int(self.lib.genvc_blob_floats(self._ctx))

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = True):
        self.load_blob(pack_state_dict(self.dims, state_dict, strict))

    def load_blob(self, blob: torch.Tensor):
        """Bind a packed blob (host or device tensor) and build the decode weight stream."""
        if blob.numel() != self.blob_floats or blob.dtype != torch.float32:
            raise ValueError("blob has the wrong size or dtype")
        with torch.cuda.device(self.device):
            self.blob = blob.to(self.device).contiguous()
            self._check(self.lib.genvc_bind_weights(self._ctx, self.blob.data_ptr(), self.blob.numel()))
            n = int(self.lib.genvc_stream_floats(self._ctx))
            if n > 0:
                self.wstream = torch.empty(n, dtype=torch.float32, device=self.device)
                self._check(self.lib.genvc_pack_stream(self._ctx, self.wstream.data_ptr(), n, self._stream()))
            # TF32 hi|lo pre-tiled copy of the dense matrices for the tcgen05 GEMMs (prefill / latent pass / perceiver)
            n = int(self.lib.genvc_tc_floats(self._ctx))
            if n > 0 and self.use_tensor_cores:
                self.wtc = torch.empty(n, dtype=torch.float32, device=self.device)
                self._check(self.lib.genvc_pack_tc(self._ctx, self.wtc.data_ptr(), n, self._stream()))

    # ------------------------------------------------------------------ the path
    def perceiver(self, mel: torch.Tensor) -> torch.Tensor:
        """mel [B, 80, S] -> speaker latents [B, 32, D]."""
        mel = self._dev(mel, torch.float32, "mel")
        if mel.ndim != 3 or mel.shape[1] != self.dims.pc_dim_context:
            raise ValueError(f"mel must be [B, {self.dims.pc_dim_context}, S]; got {tuple(mel.shape)}")
        B, _, S = mel.shape
        out = torch.empty((B, self.dims.pc_latents, self.dims.d_model), dtype=torch.float32, device=self.device)
        self._check(self.lib.genvc_perceiver(self._ctx, mel.data_ptr(), B, S, out.data_ptr(), self._stream()))
        return out

    def embed_prefix(self, cond: torch.Tensor, text_ids: torch.Tensor) -> torch.Tensor:
        d = self.dims
        self._check_ids(text_ids, d.n_text_vocab, "text token id")
        cond = self._dev(cond, torch.float32, "cond_latents")
        text_ids = self._dev(text_ids, torch.int64, "text_inputs")
        if cond.ndim != 3 or cond.shape[1:] != (d.pc_latents, d.d_model):
            raise ValueError(f"cond_latents must be [B, {d.pc_latents}, {d.d_model}]; got {tuple(cond.shape)}")
        if text_ids.ndim != 2 or text_ids.shape[0] != cond.shape[0]:
            raise ValueError("text_inputs must be [B, T] with the batch of cond_latents")
        B, T = text_ids.shape
        out = torch.empty((B, d.pc_latents + T + 2, d.d_model), dtype=torch.float32, device=self.device)
        self._check(self.lib.genvc_embed_prefix(self._ctx, cond.data_ptr(), text_ids.data_ptr(), B, T, out.data_ptr(),
                                                self._stream()))
        return out

    def prefill(self, prefix: torch.Tensor):
        prefix = self._dev(prefix, torch.float32, "prefix")
        if prefix.ndim != 3 or prefix.shape[2] != self.dims.d_model:
            raise ValueError("prefix must be [B, P, D]")
        B, P, _ = prefix.shape
        self._check(self.lib.genvc_prefill(self._ctx, prefix.data_ptr(), B, P, self._stream()))
        self._B, self._P = B, P
        self._pending, self._n = True, 0
        self._keep = prefix  # stays alive until the stream has consumed it

    def decode(self, n_steps: int, sampling: Sampling, noise: Optional[torch.Tensor] = None,
               forced: Optional[torch.Tensor] = None, want_logits: bool = False, mode: int = 0) -> DecodeChunk:
        """Up to ``n_steps`` iterations of the generation loop on the device (asynchronous)."""
        B, D, V = self._B, self.dims.d_model, self.dims.n_audio_vocab
        if B == 0:
            raise GenvcError(-2, "decode before prefill")
        if noise is not None:
            noise = self._dev(noise, torch.float32, "noise")
            if tuple(noise.shape) != (n_steps, B, V):
                raise ValueError(f"noise must be [{n_steps}, {B}, {V}]")
        if forced is not None:
            self._check_ids(forced, V, "forced token id")
            forced = self._dev(forced, torch.int64, "forced")
            if tuple(forced.shape) != (n_steps, B):
                raise ValueError(f"forced ids must be [{n_steps}, {B}]")
        dev = self.device
        # No fill kernels: the library zeroes `status` on the stream and only the first status[0] steps of the
        # outputs are defined (every consumer slices by it).
        ids = torch.empty((n_steps, B), dtype=torch.int64, device=dev)
        lat = torch.empty((n_steps, B, D), dtype=torch.float32, device=dev)
        lg = torch.empty((n_steps, B, V), dtype=torch.float32, device=dev) if want_logits else None
        status = torch.empty(4, dtype=torch.int32, device=dev)
        sp = sampling.to_c()
        ev = None
        if self.timing is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record(torch.cuda.current_stream(self.device))
        self._check(self.lib.genvc_decode(self._ctx, n_steps, C.byref(sp), _ptr(noise), _ptr(forced), ids.data_ptr(),
                                          lat.data_ptr(), _ptr(lg), status.data_ptr(), mode, self._stream()))
        if ev is not None:
            ev[1].record(torch.cuda.current_stream(self.device))
            # forwards in this call (the first step after prefill samples from the prefill's logits) and
            # the number of keys the first of them attends
            n_fwd = n_steps - (1 if self._pending else 0)
            first_S = self._P + self._n + (1 if self._pending else 0) + 1
            self.timing.append((ev[0], ev[1], n_fwd, first_S))
        self._pending = False
        self._n += n_steps
        self._keep2 = (noise, forced)
        return DecodeChunk(ids, lat, lg, status)

    def forward_latents(self, cond: torch.Tensor, text_ids: torch.Tensor, codes: torch.Tensor) -> torch.Tensor:
        """Teacher-forced pass: cond [B,32,D], text [B,T], codes [B,M] -> latents [B,M,D]."""
        self._check_ids(text_ids, self.dims.n_text_vocab, "text token id")
        self._check_ids(codes, self.dims.n_audio_vocab, "audio code")
        cond = self._dev(cond, torch.float32, "cond_latents")
        text_ids = self._dev(text_ids, torch.int64, "text_inputs")
        codes = self._dev(codes, torch.int64, "audio_codes")
        B, T = text_ids.shape
        M = codes.shape[1]
        if codes.shape[0] != B or cond.shape[0] != B:
            raise ValueError("batch mismatch")
        out = torch.empty((B, M, self.dims.d_model), dtype=torch.float32, device=self.device)
        self._check(self.lib.genvc_forward_latents(self._ctx, cond.data_ptr(), text_ids.data_ptr(), T, codes.data_ptr(), M, B,
                                                   out.data_ptr(), self._stream()))
        return out

    # ------------------------------------------------------------------ debug timeline
    def trace(self, step: Optional[int] = 0) -> Optional[torch.Tensor]:
        """Arm (``step`` = index inside each decode launch) or disarm (``None``) the fused kernel's
        phase timeline; returns the int64 buffer [grid, slots] the kernel writes %globaltimer into."""
        if step is None:
            self._check(self.lib.genvc_debug_trace(self._ctx, None, 0, 0))
            self._trace = None
            return None
        slots = self.dims.n_layer * 28 + 8
        g = self.decode_grid
        self._trace = torch.zeros(g * slots, dtype=torch.int64, device=self.device)
        self._check(self.lib.genvc_debug_trace(self._ctx, self._trace.data_ptr(), slots, int(step)))
        return self._trace.view(g, slots)

    def tune(self, window: int = 0, nosync: bool = False, l2_ahead: int = -1, hop_settle_ns: int = -1, hop_hold: int = -1):
        """Fused-kernel knobs: TMA tiles in flight per SM; ``nosync`` = streaming-rate probe (garbage results);
        ``l2_ahead`` = HBM->L2 prefetch distance in tiles (-1 keeps the current value)."""
        self._check(self.lib.genvc_debug_tune(self._ctx, int(window), int(bool(nosync)), int(l2_ahead), int(hop_settle_ns), int(hop_hold)))

    # ------------------------------------------------------------------ microbenchmark
    def kv_attention(self, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, S: int) -> torch.Tensor:
        """q [N,H,hd]; k, v [N,H,S_max,hd]; attends the first S keys -> [N,H,hd]."""
        N, H, S_max, hd = k.shape
        out = torch.empty_like(q)
        rc = self.lib.genvc_kv_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), N, H, hd, S, S_max, out.data_ptr(),
                                         self._stream())
        if rc != 0:
            raise GenvcError(rc, "genvc_kv_attention failed")
        return out
