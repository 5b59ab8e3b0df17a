"""Drop-in for the reference's `wavenet_autoencoder/model1.py`: NSynth-style WaveNet autoencoder.

Reference: class wavenet_autoencoder, wavenet_autoencoder/model1.py:12-268.  Same constructor, attributes,
submodule names and state_dict keys; `forward` runs in libwavenet_b200.so (csrc/ae.cu, fp32 check mode).

The reference draws NEW random conditioning convs (`nn.Conv1d(bottleneck, 2*Dd, 1).cuda()` per layer and one more
after connection_1, model1.py:178-179 and :216-217) on every forward call and never registers or trains them, so its
output is different on every call.  Here those N+1 convs are an explicit, persistent `cond_layers` ModuleList
(initialised exactly like the reference's fresh convs, default Conv1d init with bias); pass `cond_weights=` to
forward() to supply the ones captured from a reference run.  `fresh_cond=True` re-draws them on every call, which is
the reference's literal behaviour.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from .. import _lib as L
from .._engine import SoftmaxRowsFunction, _require_cuda


class wavenet_autoencoder(nn.Module):

    def __init__(self, filter_width, quantization_channel, dilations, en_residual_channel, en_dilation_channel,
                 en_bottleneck_width, en_pool_kernel_size, de_residual_channel, de_dilation_channel, de_skip_channel,
                 use_bias, *, fresh_cond: bool = False):
        super(wavenet_autoencoder, self).__init__()
        self.filter_width = filter_width
        self.quantization_channel = quantization_channel
        self.dilations = dilations
        self.en_residual_channel = en_residual_channel
        self.en_dilation_channel = en_dilation_channel
        self.en_bottleneck_width = en_bottleneck_width
        self.en_pool_kernel_size = en_pool_kernel_size
        self.de_residual_channel = de_residual_channel
        self.de_dilation_channel = de_dilation_channel
        self.de_skip_channel = de_skip_channel
        self.use_bias = use_bias
        self.receptive_field = self._calc_receptive_field()
        self.softmax = nn.Softmax(dim=1)
        self._init_encoding()
        self._init_decoding()
        self._init_causal_layer()
        self._init_connection()
        self.fresh_cond = fresh_cond
        # not registered in the reference (they are throw-away there): kept out of state_dict() on purpose
        object.__setattr__(self, "_cond_layers", self._new_cond_layers())
        self._handle = None
        self._ws = {}

    # ---- identical registration order / names (model1.py:55-134) ------------------------------------------------
    def _init_causal_layer(self):
        self.en_causal_layer = nn.Conv1d(self.quantization_channel, self.en_residual_channel, self.filter_width, bias=self.use_bias)
        self.bottleneck_layer = nn.Conv1d(self.en_residual_channel, self.en_bottleneck_width, 1, bias=self.use_bias)
        self.de_causal_layer = nn.Conv1d(self.quantization_channel, self.de_residual_channel, self.filter_width, bias=self.use_bias)

    def _calc_receptive_field(self):
        return (self.filter_width - 1) * (sum(self.dilations) + 1) + 1

    def _init_encoding(self):
        self.en_dilation_layer_stack = nn.ModuleList()
        self.en_dense_layer_stack = nn.ModuleList()
        for dilation in self.dilations:
            self.en_dilation_layer_stack.append(nn.Conv1d(self.en_residual_channel, self.en_dilation_channel, self.filter_width,
                                                          dilation=dilation, bias=self.use_bias))
            self.en_dense_layer_stack.append(nn.Conv1d(self.en_dilation_channel, self.en_residual_channel, 1, bias=self.use_bias))

    def _init_decoding(self):
        self.de_dilation_layer_stack = nn.ModuleList()
        for dilation in self.dilations:
            self.de_dilation_layer_stack.extend([
                nn.Conv1d(self.de_residual_channel, 2 * self.de_dilation_channel, self.filter_width, dilation=dilation,
                          bias=self.use_bias),                                                          # filter_gate
                nn.Conv1d(self.de_dilation_channel, self.de_residual_channel, kernel_size=1, dilation=dilation, bias=self.use_bias),
                nn.Conv1d(self.de_dilation_channel, self.de_skip_channel, dilation=dilation, kernel_size=1, bias=self.use_bias),
            ])

    def _init_connection(self):
        self.connection_1 = nn.Conv1d(self.de_skip_channel, self.de_skip_channel, 1, bias=self.use_bias)
        self.connection_2 = nn.Conv1d(self.de_skip_channel, self.quantization_channel, 1, bias=self.use_bias)

    def _new_cond_layers(self):
        layers = [nn.Conv1d(self.en_bottleneck_width, 2 * self.de_dilation_channel, 1) for _ in self.dilations]
        layers.append(nn.Conv1d(self.en_bottleneck_width, self.de_skip_channel, 1))
        return layers

    @property
    def cond_layers(self):
        return self._cond_layers

    # ---- engine --------------------------------------------------------------------------------------------------
    def _plan(self):
        if self._handle is None:
            lib = L.load()
            arr = (C.c_int32 * len(self.dilations))(*[int(d) for d in self.dilations])
            cfg = L.wn_ae_config(len(self.dilations), arr, self.quantization_channel, self.en_residual_channel,
                                 self.en_dilation_channel, self.en_bottleneck_width, self.en_pool_kernel_size,
                                 self.de_residual_channel, self.de_dilation_channel, self.de_skip_channel,
                                 int(bool(self.use_bias)), self.filter_width)
            h = C.c_void_p()
            L.check(lib.wn_ae_create(C.byref(cfg), C.byref(h)))
            self._handle = h
            assert lib.wn_ae_param_count(h) == sum(p.numel() for p in self.parameters())
        return self._handle

    def _cond_flat(self, cond_weights, device):
        if cond_weights is None:
            if self.fresh_cond:
                object.__setattr__(self, "_cond_layers", self._new_cond_layers())
            tensors = []
            for c in self._cond_layers:
                tensors += [c.weight.detach(), c.bias.detach()]
        else:
            tensors = list(cond_weights.values()) if isinstance(cond_weights, dict) else list(cond_weights)
        flat = torch.cat([t.reshape(-1).float() for t in tensors]).to(device).contiguous()
        n = int(L.load().wn_ae_cond_param_count(self._plan()))
        if flat.numel() != n:
            raise ValueError(f"cond_weights has {flat.numel()} values, expected {n}")
        return flat

    def forward_logits(self, wave_sample=None, indices=None, cond_weights=None, return_encoding=False):
        """Pre-softmax (B,Q,W) tensor = output of connection_2 (model1.py:221)."""
        src = wave_sample if wave_sample is not None else indices
        _require_cuda(src, "input")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and src.requires_grad:
            raise L.WavenetB200Error("wavenet_autoencoder: only the forward pass is implemented on the GPU so far")
        dev = src.device
        lib = L.init(dev.index if dev.index is not None else torch.cuda.current_device())
        h = self._plan()
        B, Lx = src.shape[0], src.shape[-1]
        W = Lx - self.receptive_field + 1
        if W <= 0:
            raise ValueError("wave sample not long enough")
        x = idx = None
        if wave_sample is not None:
            x = wave_sample.detach().float().contiguous()
        else:
            idx = indices.detach().to(torch.int64).contiguous()
        params = torch.cat([p.detach().reshape(-1).float() for p in self.parameters()]).to(dev).contiguous()
        cond = self._cond_flat(cond_weights, dev)
        key = (B, Lx, str(dev))
        ws = self._ws.get(key)
        if ws is None:
            nbytes = C.c_size_t()
            L.check(lib.wn_ae_workspace_bytes(h, B, Lx, C.byref(nbytes)))
            self._ws.clear()
            ws = torch.zeros(nbytes.value, dtype=torch.uint8, device=dev)
            self._ws[key] = ws
        logits = torch.empty(B, self.quantization_channel, W, dtype=torch.float32, device=dev)
        frames = W // self.en_pool_kernel_size
        enc = torch.empty(B, max(frames, 1), self.en_bottleneck_width, dtype=torch.float32, device=dev) if return_encoding else None
        L.check(lib.wn_ae_forward(h, B, Lx, L.ptr(x), L.ptr(idx), L.ptr(params), L.ptr(cond), L.ptr(ws), L.ptr(logits),
                                  L.ptr(enc), L.stream_ptr()))
        if return_encoding:
            return logits, enc.permute(0, 2, 1).contiguous()        # (B, BW, frames) as `_encode` returns it
        return logits

    def forward(self, wave_sample, cond_weights=None):
        """(B,Q,L) float -> (B*W, Q) probabilities in the reference's (scrambled) row order (model1.py:256-268,
        :222-224: `result.view(-1, Q)` then softmax over dim 1)."""
        logits = self.forward_logits(wave_sample=wave_sample, cond_weights=cond_weights)
        return SoftmaxRowsFunction.apply(logits, L.ROWS_REFERENCE)
