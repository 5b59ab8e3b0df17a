"""Host-side engine shared by the drop-in modules: flat parameter storage, packed kernel weights,
workspaces and the autograd bridges onto the C ABI (include/wavenet_b200.h).

PyTorch is plumbing here (device memory, streams, autograd graph, torch.distributed); every
arithmetic step of the hot path runs in libwavenet_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib as L


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise L.WavenetB200Error(
            f"{what} is on {t.device}: music_b200 runs on a B200 (sm_100a) only and has no CPU fallback; "
            "move the module and its inputs to CUDA")


class Engine:
    """Owns the C model plan plus the flat fp32 parameter / gradient vectors of one nn.Module."""

    def __init__(self, dilations, R, D, S, Q, use_bias, filter_width=2):
        self.handle = L.make_model(dilations, R, D, S, Q, use_bias, filter_width)
        self.lib = L.load()
        self.n_params = int(self.lib.wn_model_param_count(self.handle))
        self.rf = int(self.lib.wn_model_receptive_field(self.handle))
        self.Q = Q
        self.flat: Optional[torch.Tensor] = None
        self.gflat: Optional[torch.Tensor] = None
        self._packed = {}          # mode -> (tensor, signature)
        self._ws = {}              # (mode, B, L) -> tensor
        self._ws_gen = 0           # bumped by every forward that writes the workspace
        self._scratch = {}

    def __del__(self):
        try:
            self.lib.wn_model_destroy(self.handle)
        except Exception:
            pass

    # ---- flat parameters ------------------------------------------------------------------
    def ensure_flat(self, params: List[torch.nn.Parameter]) -> torch.Tensor:
        """Make every parameter a view into one contiguous fp32 vector in state_dict order."""
        dev = params[0].device
        _require_cuda(params[0], "module parameters")
        ok = self.flat is not None and self.flat.device == dev
        if ok:
            off = 0
            base = self.flat.data_ptr()
            for p in params:
                if p.data_ptr() != base + 4 * off or p.dtype != torch.float32 or not p.is_contiguous():
                    ok = False
                    break
                off += p.numel()
        if not ok:
            total = sum(p.numel() for p in params)
            if total != self.n_params:
                raise L.WavenetB200Error(f"parameter count {total} != plan {self.n_params}")
            flat = torch.empty(total, dtype=torch.float32, device=dev)
            off = 0
            with torch.no_grad():
                for p in params:
                    n = p.numel()
                    flat[off:off + n].copy_(p.detach().reshape(-1).float())
                    p.data = flat[off:off + n].view(p.shape)
                    off += n
            self.flat = flat
            self.gflat = torch.zeros(total, dtype=torch.float32, device=dev)
            self._packed.clear()
            self._ws.clear()
            L.init(dev.index if dev.index is not None else torch.cuda.current_device())
        return self.flat

    def grad_views(self, params):
        out, off = [], 0
        for p in params:
            n = p.numel()
            out.append(self.gflat[off:off + n].view(p.shape))
            off += n
        return out

    # ---- packed weights -------------------------------------------------------------------
    def packed(self, mode: int, params) -> torch.Tensor:
        sig = (self.flat.data_ptr(), tuple(p._version for p in params), self.flat._version)
        ent = self._packed.get(mode)
        if ent is not None and ent[1] == sig:
            return ent[0]
        if ent is None:
            nbytes = C.c_size_t()
            L.check(self.lib.wn_packed_bytes(self.handle, mode, C.byref(nbytes)))
            buf = torch.zeros(nbytes.value, dtype=torch.uint8, device=self.flat.device)
        else:
            buf = ent[0]
        L.check(self.lib.wn_pack_weights(self.handle, mode, L.ptr(self.flat), L.ptr(buf), L.stream_ptr()))
        self._packed[mode] = (buf, sig)
        return buf

    def invalidate_packed(self):
        for k in list(self._packed):
            self._packed[k] = (self._packed[k][0], None)

    # ---- workspaces -----------------------------------------------------------------------
    def workspace(self, mode: int, B: int, Lx: int) -> torch.Tensor:
        key = (mode, B, Lx)
        ws = self._ws.get(key)
        if ws is None:
            nbytes = C.c_size_t()
            L.check(self.lib.wn_workspace_bytes(self.handle, mode, B, Lx, C.byref(nbytes)))
            if len(self._ws) >= 2:          # keep at most two shapes alive
                self._ws.pop(next(iter(self._ws)))
            ws = torch.zeros(nbytes.value, dtype=torch.uint8, device=self.flat.device)
            self._ws[key] = ws
        return ws

    def scratch(self, name: str, nbytes: int) -> torch.Tensor:
        t = self._scratch.get(name)
        if t is None or t.numel() < nbytes or t.device != self.flat.device:
            t = torch.zeros(nbytes, dtype=torch.uint8, device=self.flat.device)
            self._scratch[name] = t
        return t

    # ---- raw calls ------------------------------------------------------------------------
    def forward_logits(self, mode, x, idx, packed, ws) -> torch.Tensor:
        src = x if x is not None else idx
        B, Lx = src.shape[0], src.shape[-1]
        W = Lx - self.rf + 1
        if W <= 0:
            raise ValueError("wave sample not long enough")          # wavenet/model.py:100-101
        logits = torch.empty(B, self.Q, W, dtype=torch.float32, device=src.device)
        L.check(self.lib.wn_forward(self.handle, mode, B, Lx, L.ptr(x), L.ptr(idx), L.ptr(packed), L.ptr(ws),
                                    L.ptr(logits), L.stream_ptr()))
        self._ws_gen += 1
        return logits

    def backward(self, mode, x, idx, packed, ws, dlogits, gout) -> None:
        src = x if x is not None else idx
        B, Lx = src.shape[0], src.shape[-1]
        L.check(self.lib.wn_backward(self.handle, mode, B, Lx, L.ptr(x), L.ptr(idx), L.ptr(packed), L.ptr(ws),
                                     L.ptr(dlogits), L.ptr(gout), L.stream_ptr()))


class ConvStackFunction(torch.autograd.Function):
    """logits = conv stack(params, input); backward produces per-parameter gradients."""

    @staticmethod
    def forward(ctx, engine: Engine, mode: int, x, idx, *params):
        engine.ensure_flat(list(params))
        packed = engine.packed(mode, params)
        src = x if x is not None else idx
        ws = engine.workspace(mode, src.shape[0], src.shape[-1])
        logits = engine.forward_logits(mode, x, idx, packed, ws)
        ctx.engine, ctx.mode, ctx.x, ctx.idx, ctx.packed, ctx.ws = engine, mode, x, idx, packed, ws
        ctx.gen = engine._ws_gen
        ctx.shapes = [p.shape for p in params]
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        e = ctx.engine
        if ctx.gen != e._ws_gen:
            raise L.WavenetB200Error(
                "the activation workspace of this forward was overwritten by a later forward; "
                "call backward() before running the module again")
        g = torch.empty(e.n_params, dtype=torch.float32, device=dlogits.device)
        e.backward(ctx.mode, ctx.x, ctx.idx, ctx.packed, ctx.ws, dlogits.contiguous().clone(), g)
        grads, off = [], 0
        for shp in ctx.shapes:
            n = 1
            for s in shp:
                n *= s
            grads.append(g[off:off + n].view(shp))
            off += n
        return (None, None, None, None, *grads)


class SoftmaxRowsFunction(torch.autograd.Function):
    """The reference's `total.view(-1, Q)` + `nn.Softmax()` (wavenet/model.py:142-144)."""

    @staticmethod
    def forward(ctx, logits, rows: int):
        lib = L.load()
        B, Q, W = logits.shape
        logits = logits.contiguous()
        probs = torch.empty(B * W, Q, dtype=torch.float32, device=logits.device)
        L.check(lib.wn_softmax_fwd(L.ptr(logits), B, Q, W, rows, L.ptr(probs), L.stream_ptr()))
        ctx.save_for_backward(probs)
        ctx.rows, ctx.shape = rows, (B, Q, W)
        return probs

    @staticmethod
    def backward(ctx, dprobs):
        (probs,) = ctx.saved_tensors
        B, Q, W = ctx.shape
        dprobs = dprobs.contiguous().float()
        dlogits = torch.empty(B, Q, W, dtype=torch.float32, device=probs.device)
        L.check(L.load().wn_softmax_bwd(L.ptr(probs), L.ptr(dprobs), B, Q, W, ctx.rows, L.ptr(dlogits), L.stream_ptr()))
        return dlogits, None


def loss_scratch_bytes(B: int, W: int) -> int:
    nbytes = C.c_size_t()
    L.check(L.load().wn_loss_scratch_bytes(B, W, C.byref(nbytes)))
    return int(nbytes.value)


def fused_loss(logits: torch.Tensor, target: torch.Tensor, rows: int, want_grad: bool, grad_scale: float = 1.0,
               scratch: Optional[torch.Tensor] = None):
    """(loss[1] device tensor, dlogits or None) via wn_loss_fwd_bwd."""
    lib = L.load()
    B, Q, W = logits.shape
    nbytes = C.c_size_t()
    L.check(lib.wn_loss_scratch_bytes(B, W, C.byref(nbytes)))
    if scratch is None or scratch.numel() < nbytes.value:
        scratch = torch.empty(nbytes.value, dtype=torch.uint8, device=logits.device)
    loss = torch.empty(1, dtype=torch.float32, device=logits.device)
    dlogits = torch.empty_like(logits) if want_grad else None
    tgt = target.reshape(-1).to(torch.int64).contiguous()
    L.check(lib.wn_loss_fwd_bwd(L.ptr(logits), L.ptr(tgt), B, Q, W, rows, float(grad_scale), L.ptr(loss),
                                L.ptr(dlogits), L.ptr(scratch), L.stream_ptr()))
    return loss, dlogits
