"""Drop-in for the reference's `wavenet_autoencoder/model1.py`: NSynth-style WaveNet autoencoder.

Reference: class wavenet_autoencoder, wavenet_autoencoder/model1.py:12-268.  Same constructor, attributes,
submodule names and state_dict keys; `forward` and its backward run in libwavenet_b200.so (csrc/ae.cu, fp32 check
mode), so `loss.backward()` / `optimizer.step()` of wavenet_autoencoder/train.py:150-160 work unchanged.

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


class _AeFunction(torch.autograd.Function):
    """logits (B,Q,W) = connection_2 output; backward = wn_ae_backward (per-parameter gradients, and gradients of
    the conditioning convs when they take part in autograd)."""

    @staticmethod
    def forward(ctx, net, x, idx, cond, *params):
        src = x if x is not None else idx
        dev = src.device
        lib, h = L.load(), net._plan()
        B, Lx = src.shape[0], src.shape[-1]
        W = Lx - net.receptive_field + 1
        flat = torch.cat([p.detach().reshape(-1).float() for p in params]).contiguous()
        cond = cond.detach().contiguous()
        ws = net._workspace(B, Lx, dev, train=True)
        logits = torch.empty(B, net.quantization_channel, W, dtype=torch.float32, device=dev)
        L.check(lib.wn_ae_forward_train(h, B, Lx, L.ptr(x), L.ptr(idx), L.ptr(flat), L.ptr(cond), L.ptr(ws), L.ptr(logits),
                                        None, L.stream_ptr()))
        net._ws_gen += 1
        ctx.net, ctx.x, ctx.idx, ctx.flat, ctx.cond, ctx.ws, ctx.gen = net, x, idx, flat, cond, ws, net._ws_gen
        ctx.shapes = [p.shape for p in params]
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        net = ctx.net
        if ctx.gen != net._ws_gen:
            raise L.WavenetB200Error("the activation workspace of this forward was overwritten by a later forward; "
                                     "call backward() before running the module again")
        src = ctx.x if ctx.x is not None else ctx.idx
        B, Lx = src.shape[0], src.shape[-1]
        g = torch.empty_like(ctx.flat)
        gc = torch.empty_like(ctx.cond) if ctx.needs_input_grad[3] else None
        L.check(L.load().wn_ae_backward(net._plan(), B, Lx, L.ptr(ctx.x), L.ptr(ctx.idx), L.ptr(ctx.flat), L.ptr(ctx.cond),
                                        L.ptr(ctx.ws), L.ptr(dlogits.contiguous().float()), L.ptr(g), L.ptr(gc), L.stream_ptr()))
        grads, off = [], 0
        for shp in ctx.shapes:
            n = int(np.prod(shp))
            grads.append(g[off:off + n].view(shp))
            off += n
        return (None, None, None, gc, *grads)


class wavenet_autoencoder(nn.Module):

    def __init__(self, filter_width, quantization_channel, dilations, en_residual_channel, en_dilation_channel,
                 en_bottleneck_width, en_pool_kernel_size, de_residual_channel, de_dilation_channel, de_skip_channel,
                 use_bias, *, fresh_cond: bool = False, mode: str = "auto"):
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
        if mode not in L.MODE_NAMES:
            raise ValueError(f"mode must be one of {list(L.MODE_NAMES)}")
        # "bf16": the conditioned decoder (89 % of the FLOPs at the shipped parameters) runs on the tcgen05 WaveNet kernels, the
        # encoder in fp32; "auto" picks that whenever the decoder's shape has kernels; "fp32": the SIMT check mode (1e-4 parity)
        fast_ok = (de_residual_channel <= 64 and de_dilation_channel <= 64 and de_skip_channel in (256, 512)
                   and quantization_channel == 256 and not use_bias)
        if mode == "auto" and not fast_ok:
            import warnings
            warnings.warn("music_b200: no tcgen05 kernels for this decoder shape (residual %d, dilation %d, skip %d, quantization %d, "
                          "bias %s); mode='auto' runs the autoencoder in the fp32 SIMT mode" %
                          (de_residual_channel, de_dilation_channel, de_skip_channel, quantization_channel, bool(use_bias)))
        self.mode = "fp32" if (mode == "fp32" or (mode == "auto" and not fast_ok)) else "bf16"
        # not registered in the reference (they are throw-away there): kept out of state_dict() on purpose
        object.__setattr__(self, "_cond_layers", self._new_cond_layers())
        self._handle = None
        self._ws = {}
        self._ws_gen = 0

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
        for c in layers:                 # throw-away in the reference: not trained unless the caller turns this on
            c.requires_grad_(False)
        return layers

    @property
    def cond_layers(self):
        return self._cond_layers

    # ---- engine --------------------------------------------------------------------------------------------------
    def __del__(self):
        h = self.__dict__.get("_handle")
        if h is not None:
            try:
                L.load().wn_ae_destroy(h)
            except Exception:
                pass
            self.__dict__["_handle"] = None

    def __deepcopy__(self, memo):
        """The C plan handle has one owner: a copy builds its own lazily."""
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k == "_handle" else ({} if k == "_ws" else copy.deepcopy(v, memo))
        return new

    def _plan(self):
        if self._handle is None:
            lib = L.load()
            arr = (C.c_int32 * len(self.dilations))(*[int(d) for d in self.dilations])
            cfg = L.wn_ae_config(len(self.dilations), arr, self.quantization_channel, self.en_residual_channel,
                                 self.en_dilation_channel, self.en_bottleneck_width, self.en_pool_kernel_size,
                                 self.de_residual_channel, self.de_dilation_channel, self.de_skip_channel,
                                 int(bool(self.use_bias)), self.filter_width, 1 if self.mode == "bf16" else 0)
            h = C.c_void_p()
            L.check(lib.wn_ae_create(C.byref(cfg), C.byref(h)))
            self._handle = h
            assert lib.wn_ae_param_count(h) == sum(p.numel() for p in self.parameters())
        return self._handle

    def _workspace(self, B, Lx, dev, train):
        key = (B, Lx, str(dev), train)
        ws = self._ws.get(key)
        if ws is None:
            lib = L.load()
            nbytes = C.c_size_t()
            fn = lib.wn_ae_train_workspace_bytes if train else lib.wn_ae_workspace_bytes
            L.check(fn(self._plan(), B, Lx, C.byref(nbytes)))
            self._ws.clear()
            ws = torch.zeros(nbytes.value, dtype=torch.uint8, device=dev)
            self._ws[key] = ws
        return ws

    def _cond_flat(self, cond_weights, device):
        if cond_weights is None:
            if self.fresh_cond:
                object.__setattr__(self, "_cond_layers", self._new_cond_layers())
            tensors = []
            for c in self._cond_layers:      # kept in the autograd graph: they receive gradients if they require them
                tensors += [c.weight, c.bias]
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
        cond = self._cond_flat(cond_weights, dev)
        plist = list(self.parameters())
        if torch.is_grad_enabled() and (any(p.requires_grad for p in plist) or cond.requires_grad):
            if return_encoding:
                raise L.WavenetB200Error("return_encoding is an inference-only option (use torch.no_grad())")
            for p in plist:
                _require_cuda(p, "parameter")
            return _AeFunction.apply(self, x, idx, cond, *plist)
        params = torch.cat([p.detach().reshape(-1).float() for p in plist]).to(dev).contiguous()
        cond = cond.detach()
        ws = self._workspace(B, Lx, dev, train=False)
        self._ws_gen += 1
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
