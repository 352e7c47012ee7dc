"""Content-DVAE tokeniser on the CUDA library — the stage right before the codec-token path.

Host-side mirror of the inference slice of ``layers/dvae.py::DiscreteVAE`` for ``positional_dims=1`` (the content DVAE of
``trainers/hifigan_trainer.py:149-160``): same constructor arguments, same state-dict keys (``encoder.N...``,
``codebook.embed``; decoder keys are ignored), ``get_codebook_indices(features[B, channels, T]) -> int64 [B, T']``.
Encoder convolutions run in ``genvc_conv1d`` (ReLU and the ResBlock residual folded into the epilogue), the nearest
codebook entry in ``genvc_codebook_argmin``.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List

import torch

from .lib import GenvcError, load_library


class DiscreteVAE:
    def __init__(self, positional_dims=1, num_tokens=512, codebook_dim=512, num_layers=3, num_resnet_blocks=0, hidden_dim=64,
                 channels=3, stride=2, kernel_size=4, use_transposed_convs=True, encoder_norm=False, activation="relu",
                 normalization=None, device="cuda", **_ignored):
        if positional_dims != 1 or encoder_norm or activation != "relu" or normalization is not None or num_layers < 1:
            raise NotImplementedError("genvc_b200 DiscreteVAE covers the content DVAE: 1-d, ReLU, no encoder norm, no input normalisation")
        if stride not in (1, 2):
            raise NotImplementedError("stride 1 or 2")
        self.num_tokens, self.codebook_dim, self.num_layers = int(num_tokens), int(codebook_dim), int(num_layers)
        self.num_resnet_blocks, self.hidden_dim, self.channels = int(num_resnet_blocks), int(hidden_dim), int(channels)
        self.stride, self.kernel_size = int(stride), int(kernel_size)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("genvc_b200.content_dvae.DiscreteVAE runs on a CUDA device only (no CPU fallback)")
        self.lib = load_library()
        self._w: Dict[str, tuple] = {}
        self._embed = None
        self._scratch = torch.empty(2 * 1024 * 1024, dtype=torch.float32, device=self.device)
        self._graphs: Dict[tuple, tuple] = {}
        self.use_graphs = os.environ.get("GENVC_STAGE_GRAPHS", "1") != "0"  # 0: every call launches its kernels eagerly
        self.launches = 0

    def eval(self):
        return self

    def to(self, device):
        if torch.device(device) != self.device:
            raise RuntimeError("construct the tokeniser on the target device")
        return self

    def conv_names(self) -> List[str]:
        names, i = [], 0
        for _ in range(self.num_layers):
            names.append(f"encoder.{i}.0")
            i += 1
        for _ in range(self.num_resnet_blocks):
            names += [f"encoder.{i}.net.0", f"encoder.{i}.net.2", f"encoder.{i}.net.4"]
            i += 1
        names.append(f"encoder.{i}")
        return names

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = False):
        self._w.clear()
        self._graphs.clear()
        for name in self.conv_names():
            w = sd[name + ".weight"].float()  # [Cout, Cin, K] -> [Cin][K][Cout]
            self._w[name] = (w.permute(1, 2, 0).contiguous().to(self.device), sd[name + ".bias"].float().contiguous().to(self.device),
                             int(w.shape[1]), int(w.shape[0]), int(w.shape[2]))
        e = sd["codebook.embed"].float()
        if tuple(e.shape) != (self.codebook_dim, self.num_tokens):
            raise ValueError(f"codebook.embed has shape {tuple(e.shape)}, expected {(self.codebook_dim, self.num_tokens)}")
        self._embed = e.contiguous().to(self.device)
        return self

    def _conv(self, name, x, T, B, stride, pad, act, residual=None, st=None):
        w, b, cin, cout, k = self._w[name]
        To = (T + 2 * pad - (k - 1) - 1) // stride + 1
        y = torch.empty((B, cout, To), dtype=torch.float32, device=self.device)  # (inside a graph capture: the graph's pool)
        self.launches += 1
        rc = self.lib.genvc_conv1d(x.data_ptr(), w.data_ptr(), b.data_ptr(), residual.data_ptr() if residual is not None else None,
                                   y.data_ptr(), B, cin, cout, T, k, 1, pad, stride, 1.0, 0, 1.0, act, self._scratch.data_ptr(),
                                   self._scratch.numel(), st)
        if rc != 0:
            raise GenvcError(rc, name)
        return y, To

    @torch.inference_mode()
    def get_codebook_indices(self, images: torch.Tensor) -> torch.Tensor:
        """layers/dvae.py:324-331 (``log_codes`` is training bookkeeping and is skipped)."""
        if self._embed is None:
            raise RuntimeError("DiscreteVAE: load_state_dict() first")
        if images.dim() != 3 or images.shape[1] != self.channels:
            raise ValueError(f"expected [B, {self.channels}, T], got {tuple(images.shape)}")
        x = images.to(self.device, torch.float32).contiguous()
        B, _, T = (int(v) for v in x.shape)
        with torch.cuda.device(self.device):
            # a call of a given (B, T) is 25 small launches: captured once as a CUDA graph, replayed afterwards
            entry = self._graphs.get((B, T))
            if entry is not None:
                graph, x_in, codes = entry
                x_in.copy_(x)
                graph.replay()
                return codes.clone()
            cur = torch.cuda.current_stream(self.device)
            codes = self._run(x, B, T, C.c_void_p(cur.cuda_stream))  # eager first pass (loads the kernels)
            if self.use_graphs:
                x_in = x.clone()
                side = torch.cuda.Stream(self.device)
                side.wait_stream(cur)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    out = self._run(x_in, B, T, C.c_void_p(side.cuda_stream))
                cur.wait_stream(side)
                self._graphs[(B, T)] = (graph, x_in, out)
                if len(self._graphs) > 16:
                    self._graphs.pop(next(iter(self._graphs)))
            return codes

    def _run(self, x, B, T, st):
        names = self.conv_names()
        pad = (self.kernel_size - 1) // 2
        i = 0
        for _ in range(self.num_layers):  # conv + ReLU (:265)
            x, T = self._conv(names[i], x, T, B, self.stride, pad, 2, st=st)
            i += 1
        for _ in range(self.num_resnet_blocks):  # conv3 - act - conv3 - act - conv1, + input (:171-184)
            h, _ = self._conv(names[i], x, T, B, 1, 1, 2, st=st)
            h, _ = self._conv(names[i + 1], h, T, B, 1, 1, 2, st=st)
            x, _ = self._conv(names[i + 2], h, T, B, 1, 0, 0, residual=x, st=st)
            i += 3
        x, _ = self._conv(names[i], x, T, B, 1, 0, 0, st=st)  # to codebook_dim (:284)
        codes = torch.empty((B, T), dtype=torch.int64, device=self.device)
        self.launches += 1
        rc = self.lib.genvc_codebook_argmin(x.data_ptr(), self._embed.data_ptr(), codes.data_ptr(), B, self.codebook_dim,
                                            self.num_tokens, T, st)
        if rc != 0:
            raise GenvcError(rc, "codebook argmin")
        return codes

    @classmethod
    def from_config(cls, dvae_config, device="cuda") -> "DiscreteVAE":
        """``trainers/hifigan_trainer.py:149-160``."""
        g = (lambda k, d: dvae_config.get(k, d)) if isinstance(dvae_config, dict) else (lambda k, d: getattr(dvae_config, k, d))
        return cls(positional_dims=1, channels=g("num_channels", 256), num_tokens=g("num_tokens", 256), codebook_dim=g("codebook_dim", 512),
                   hidden_dim=g("hidden_dim", 512), num_resnet_blocks=g("num_resnet_blocks", 3), kernel_size=g("kernel_size", 3),
                   num_layers=g("num_layers", 2), use_transposed_convs=False, device=device)
