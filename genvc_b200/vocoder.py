"""HiFi-GAN generator on the CUDA library — the stage right after the codec-token path.

Host-side mirror of ``layers/hifigan.py::HiFiGAN`` (``:156-243``): same constructor arguments, same state-dict keys
(``conv_pre``, ``ups.N``, ``resblocks.N.convs.M`` / ``convs1|2.M``, ``conv_post``; weight norm as ``weight_g`` / ``weight_v``
or already removed), ``forward(x[B, input_feat_dim, T]) -> [B, 1, T * prod(upsample_rates)]``, ``remove_weight_norm()``.
The arithmetic runs in two kernels of ``libgenvc_b200.so`` (``genvc_conv1d``, ``genvc_conv_transpose1d``: activation,
bias, residual, resblock sum and tanh fused into the convolutions — ``csrc/vocoder.cu``); a forward of a given
``(B, T)`` is captured once as a CUDA graph and replayed.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .lib import GenvcError, load_library

LRELU_SLOPE = 0.1  # layers/hifigan.py:22


def _get_padding(kernel_size: int, dilation: int = 1) -> int:  # utils.py:174-175
    return int((kernel_size * dilation - dilation) / 2)


def _fold(sd: Dict[str, torch.Tensor], prefix: str) -> torch.Tensor:
    """weight norm (dim 0) folded: w = g * v / ||v|| (norm over every dim but the first)."""
    if prefix + ".weight" in sd:
        return sd[prefix + ".weight"].float()
    if prefix + ".parametrizations.weight.original0" in sd:  # torch.nn.utils.parametrizations.weight_norm
        g, v = sd[prefix + ".parametrizations.weight.original0"].float(), sd[prefix + ".parametrizations.weight.original1"].float()
    else:
        g, v = sd[prefix + ".weight_g"].float(), sd[prefix + ".weight_v"].float()
    norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
    return g * v / norm


class _Conv:
    __slots__ = ("w", "b", "cin", "cout", "k", "dil", "pad", "stride", "transposed")


class HiFiGAN:
    """Drop-in for ``layers.hifigan.HiFiGAN`` at inference (``model.hifigan`` of ``inference/model_init.py``)."""

    def __init__(self, input_feat_dim, upsample_initial_channel, resblock_kernel_sizes, resblock_dilation_sizes, upsample_rates,
                 upsample_kernel_sizes, resblock_type="1", device="cuda"):
        self.input_feat_dim = int(input_feat_dim)
        self.c0 = int(upsample_initial_channel)
        self.rks = [int(k) for k in resblock_kernel_sizes]
        self.rds = [[int(d) for d in ds] for ds in resblock_dilation_sizes]
        self.rates = [int(u) for u in upsample_rates]
        self.uks = [int(k) for k in upsample_kernel_sizes]
        self.resblock_type = str(resblock_type)
        self.num_kernels = len(self.rks)
        self.num_upsamples = len(self.rates)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("genvc_b200.vocoder.HiFiGAN runs on a CUDA device only (no CPU fallback)")
        self.lib = load_library()
        self._convs: Dict[str, _Conv] = {}
        self._graphs: Dict[Tuple[int, int], tuple] = {}
        self.use_graphs = os.environ.get("GENVC_STAGE_GRAPHS", "1") != "0"  # 0: every call launches its kernels eagerly
        self.launches = 0
        # split-over-input-channels partial sums of the small layers (csrc/vocoder.cu: pick_ksplit); 8 MB covers 16 slices of
        # the largest layer that is ever split at streaming sizes
        self._scratch = torch.empty(2 * 1024 * 1024, dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------------------------------------ nn.Module surface
    def eval(self):
        return self

    def to(self, device):
        if torch.device(device) != self.device:
            raise RuntimeError("move the vocoder by constructing it on the target device")
        return self

    def remove_weight_norm(self):  # layers/hifigan.py:227-234 — folded at load time here
        return None

    def conv_names(self) -> List[Tuple[str, bool, int, int]]:
        """(state-dict prefix, transposed, kernel/stride info...) in the reference's construction order."""
        out = [("conv_pre", False, 1, 3)]
        for i in range(self.num_upsamples):
            out.append((f"ups.{i}", True, self.rates[i], (self.uks[i] - self.rates[i]) // 2))
        for i in range(self.num_upsamples):
            for j, (k, ds) in enumerate(zip(self.rks, self.rds)):
                for m, d in enumerate(ds):
                    base = f"resblocks.{i * self.num_kernels + j}"
                    if self.resblock_type == "1":
                        out.append((f"{base}.convs1.{m}", False, d, _get_padding(k, d)))
                        out.append((f"{base}.convs2.{m}", False, 1, _get_padding(k, 1)))
                    else:
                        out.append((f"{base}.convs.{m}", False, d, _get_padding(k, d)))
        out.append(("conv_post", False, 1, 3))
        return out

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        self._convs.clear()
        self._graphs.clear()
        for prefix, transposed, dil_or_stride, pad in self.conv_names():
            try:
                w = _fold(sd, prefix)
            except KeyError as e:
                raise KeyError(f"HiFiGAN state dict has no weights for {prefix}") from e
            c = _Conv()
            c.transposed = transposed
            if transposed:  # ConvTranspose1d weight [Cin, Cout, K] -> [Cin][K][Cout]
                c.cin, c.cout, c.k = (int(v) for v in w.shape)
                c.w = w.permute(0, 2, 1).contiguous().to(self.device)
                c.stride, c.dil = dil_or_stride, 1
            else:           # Conv1d weight [Cout, Cin, K] -> [Cin][K][Cout]
                c.cout, c.cin, c.k = (int(v) for v in w.shape)
                c.w = w.permute(1, 2, 0).contiguous().to(self.device)
                c.stride, c.dil = 1, dil_or_stride
            c.pad = pad
            b = sd.get(prefix + ".bias")
            c.b = b.float().contiguous().to(self.device) if b is not None else None
            self._convs[prefix] = c
        return self

    # ------------------------------------------------------------------------------------------------ kernels
    def _check(self, rc: int, what: str):
        if rc != 0:
            raise GenvcError(rc, what)

    def _conv(self, name: str, x, y, T: int, B: int, slope: float, residual=None, accumulate=False, scale=1.0, tanh=False, st=None):
        c = self._convs[name]
        self.launches += 1
        self._check(self.lib.genvc_conv1d(x.data_ptr(), c.w.data_ptr(), c.b.data_ptr() if c.b is not None else None,
                                          residual.data_ptr() if residual is not None else None, y.data_ptr(), B, c.cin, c.cout, T,
                                          c.k, c.dil, c.pad, 1, float(slope), int(accumulate), float(scale), int(tanh),
                                          self._scratch.data_ptr(), self._scratch.numel(), st), name)

    def _up(self, name: str, x, y, Tin: int, B: int, slope: float, st=None):
        c = self._convs[name]
        self.launches += 1
        self._check(self.lib.genvc_conv_transpose1d(x.data_ptr(), c.w.data_ptr(), c.b.data_ptr() if c.b is not None else None,
                                                    y.data_ptr(), B, c.cin, c.cout, Tin, c.k, c.stride, c.pad, float(slope),
                                                    self._scratch.data_ptr(), self._scratch.numel(), st), name)

    def _run(self, x: torch.Tensor, bufs, st) -> torch.Tensor:
        """layers/hifigan.py:210-225 with every elementwise op folded into a convolution."""
        B, _, T = x.shape
        cur, tmp, xs, wav = bufs["cur"], bufs["tmp"], bufs["xs"], bufs["wav"]
        self._conv("conv_pre", x, cur[0], T, B, 1.0, st=st)  # :211
        src, slope = cur[0], LRELU_SLOPE
        for i in range(self.num_upsamples):
            Tn = T * self.rates[i]
            xu = cur[i + 1]
            self._up(f"ups.{i}", src, xu, T, B, slope, st=st)  # leaky_relu(x, 0.1) + ups[i]  (:213-214)
            T = Tn
            for j, ds in enumerate(self.rds):
                base = f"resblocks.{i * self.num_kernels + j}"
                last_block_scale = 1.0 / self.num_kernels if j == self.num_kernels - 1 else 1.0
                a = xu
                if self.resblock_type == "1":  # :98-105
                    for m in range(len(ds)):
                        last = m == len(ds) - 1
                        self._conv(f"{base}.convs1.{m}", a, tmp[i][0], T, B, LRELU_SLOPE, st=st)
                        out = xs[i] if last else tmp[i][1 + (m & 1)]
                        self._conv(f"{base}.convs2.{m}", tmp[i][0], out, T, B, LRELU_SLOPE, residual=a,
                                   accumulate=last and j > 0, scale=last_block_scale if last else 1.0, st=st)
                        a = out
                else:  # ResBlock2 (:147-152)
                    for m in range(len(ds)):
                        last = m == len(ds) - 1
                        out = xs[i] if last else tmp[i][m & 1]
                        self._conv(f"{base}.convs.{m}", a, out, T, B, LRELU_SLOPE, residual=a, accumulate=last and j > 0,
                                   scale=last_block_scale if last else 1.0, st=st)
                        a = out
            src = xs[i]  # = xs / num_kernels (:221)
        self._conv("conv_post", src, wav, T, B, 0.01, tanh=True, st=st)  # F.leaky_relu default slope, conv_post, tanh (:222-224)
        return wav

    def _buffers(self, B: int, T: int):
        dev, f32 = self.device, torch.float32
        cur = [torch.empty((B, self.c0, T), device=dev, dtype=f32)]
        tmp, xs = [], []
        t = T
        for i in range(self.num_upsamples):
            t *= self.rates[i]
            ch = self.c0 // 2 ** (i + 1)
            cur.append(torch.empty((B, ch, t), device=dev, dtype=f32))
            tmp.append([torch.empty((B, ch, t), device=dev, dtype=f32) for _ in range(3)])
            xs.append(torch.empty((B, ch, t), device=dev, dtype=f32))
        return {"cur": cur, "tmp": tmp, "xs": xs, "wav": torch.empty((B, 1, t), device=dev, dtype=f32),
                "x": torch.empty((B, self.input_feat_dim, T), device=dev, dtype=f32)}

    @torch.inference_mode()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self._convs:
            raise RuntimeError("HiFiGAN: load_state_dict() first")
        if x.dim() != 3 or x.shape[1] != self.input_feat_dim:
            raise ValueError(f"expected [B, {self.input_feat_dim}, T], got {tuple(x.shape)}")
        x = x.to(self.device, torch.float32).contiguous()
        B, _, T = (int(v) for v in x.shape)
        with torch.cuda.device(self.device):
            key = (B, T)
            entry = self._graphs.get(key)
            if entry is None:
                bufs = self._buffers(B, T)
                bufs["x"].copy_(x)
                st = torch.cuda.current_stream(self.device)
                self._run(bufs["x"], bufs, C.c_void_p(st.cuda_stream))  # eager first pass (also loads the kernels)
                graph = None
                if self.use_graphs:
                    side = torch.cuda.Stream(self.device)
                    side.wait_stream(st)
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.stream(side):
                        with torch.cuda.graph(graph, stream=side):
                            self._run(bufs["x"], bufs, C.c_void_p(side.cuda_stream))
                    st.wait_stream(side)
                self._graphs[key] = (graph, bufs)
                if len(self._graphs) > 16:  # streaming produces a handful of chunk lengths; keep the cache bounded
                    self._graphs.pop(next(iter(self._graphs)))
                return bufs["wav"].clone()
            graph, bufs = entry
            bufs["x"].copy_(x)
            if graph is not None:
                graph.replay()
            else:
                self._run(bufs["x"], bufs, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
            return bufs["wav"].clone()

    __call__ = forward

    @classmethod
    def from_config(cls, vocoder_config, device="cuda") -> "HiFiGAN":
        """``trainers/hifigan_trainer.py:47-55``: the generator is built from ``config.vocoder_config``."""
        g = (lambda k, d: vocoder_config.get(k, d)) if isinstance(vocoder_config, dict) else (lambda k, d: getattr(vocoder_config, k, d))
        return cls(g("input_feat_dim", 1024), g("upsample_initial_channel", 256), g("resblock_kernel_sizes", [3, 5, 7]),
                   g("resblock_dilation_sizes", [[1, 2], [2, 6], [3, 12]]), g("upsample_rates", [8, 8, 4]),
                   g("upsample_kernal_sizes", g("upsample_kernel_sizes", [16, 16, 8])), g("resblock_type", "2"), device=device)
