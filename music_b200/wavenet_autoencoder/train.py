"""Drop-in for the reference's `wavenet_autoencoder/train.py`: optimizer factory, checkpoint helpers, one train step.

Reference: get_optimizer :26-34, save_model :36-41, load_model :44-64, hot loop :117-134
(zero_grad -> net(music_piece) -> CrossEntropyLoss(outputs, target.view(-1)) -> backward -> step).  The loss is
applied to the module's softmax OUTPUT (a second softmax inside CrossEntropyLoss), exactly as in wavenet/train.py.
Forward and backward of the module run in libwavenet_b200.so (csrc/ae.cu, ae_fast.cu); `train_step` keeps the reference's
shape (torch.optim objects on the module's parameters, autograd in between), `AeTrainer` is the fused step without an autograd
graph: flat parameters, loss + dlogits in one kernel, fused optimizer, one all-reduce.
"""
from __future__ import annotations

import os
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.optim as optim

from .. import _lib as L
from .._engine import fused_loss, loss_scratch_bytes, _require_cuda
from ..wavenet.train import all_reduce_mean_
from .model1 import wavenet_autoencoder  # noqa: F401  (re-exported like the reference's `from model1 import ...`)


def get_optimizer(model, optimizer_type, learning_rate, momentum1=False):
    """Names as the reference spells them (:26-34).  Its 'sgd' branch calls the non-existent `optim.sgd`; SGD is meant."""
    if optimizer_type == 'sgd':
        return optim.SGD(model.parameters(), lr=learning_rate, momentum=momentum1)
    if optimizer_type == 'RMSprop':
        return optim.RMSprop(model.parameters(), lr=learning_rate, momentum=momentum1)
    if optimizer_type == 'Adam':
        return optim.Adam(model.parameters(), lr=learning_rate)
    if optimizer_type == 'lbfgs':
        return optim.LBFGS(model.parameters(), lr=learning_rate)


def save_model(model, num_epoch, path):
    model_name = 'wavenet_autoencoder' + str(num_epoch) + '.model'
    checkpoint_path = path + model_name
    print('Storing checkpoint to {}...'.format(path))
    torch.save(OrderedDict((k, v.detach().cpu().clone()) for k, v in model.state_dict().items()), checkpoint_path)
    print('done')


def load_model(model, path, model_name):
    """Returns the model, or None when the file is missing; strips DataParallel's `module.` prefix (:44-64)."""
    checkpoint_path = path + model_name
    print("Trying to restore saved checkpoint from ", "{}".format(checkpoint_path))
    if not os.path.exists(checkpoint_path):
        print("No checkpoint found!")
        return None
    print("Checkpoint found, restoring!")
    state_dict = torch.load(checkpoint_path, map_location="cpu")
    keys = list(state_dict.keys())
    if keys[0][:6] == 'module':
        state_dict = OrderedDict((k[7:], v) for k, v in state_dict.items())
    model.load_state_dict(state_dict)
    return model


def all_reduce_grads_(params, dist=None, group=None):
    """Data-parallel gradient exchange for one process per GPU (the reference wraps the net in nn.DataParallel,
    train.py:93-97): ONE all-reduce of the flattened gradients, averaged over ranks.  Every rank holds the same number of
    loss rows, so this equals the gradient of the mean loss over the gathered batch.  Works on any backend (NCCL on the
    GPUs, gloo in the CPU tests)."""
    if dist is None:
        import torch.distributed as dist
    params = [p for p in params if p.grad is not None]
    if not params or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n


def train_step(net, optimizer, music_piece, target_piece, loss_func=None, cond_weights=None, distributed=False):
    """One iteration of the reference loop (:117-134).  `music_piece` (B,Q,L) float on the GPU, `target_piece` any
    integer tensor with B*W entries.  Returns the loss as a 0-d device tensor (no host sync; the reference's
    `loss.data[0]` is the caller's choice).  distributed=True averages the gradients over the process group first."""
    loss_func = loss_func or nn.CrossEntropyLoss()
    optimizer.zero_grad()
    outputs = net(music_piece, cond_weights=cond_weights)
    loss = loss_func(outputs, target_piece.reshape(-1).to(outputs.device, torch.int64))
    loss.backward()
    if distributed:
        all_reduce_grads_(net.parameters())
    optimizer.step()
    return loss.detach()


class AeTrainer:
    """Fused train step of the autoencoder: zero_grad -> net(piece) -> CrossEntropyLoss(probabilities, target) -> backward ->
    [all-reduce] -> optimizer step (wavenet_autoencoder/train.py:117-134) as four C-ABI calls on flat vectors: wn_ae_forward_train,
    wn_loss_fwd_bwd (flat-chunk softmax rows + double-softmax cross entropy + dlogits, as the reference's objective), wn_ae_backward
    and wn_adam_step / wn_sgd_step / wn_rmsprop_step.  The module's parameters become views into one fp32 vector (state_dict and
    checkpoints are unchanged); after step() every `p.grad` is a view into the flat gradient.

    step(piece, target): `piece` is the dense (B,Q,L) float tensor the reference feeds or a (B,L) integer tensor of mu-law codes;
    `target` any integer tensor with B*W entries.  Returns the loss as a 1-element device tensor (no host sync).
    """

    def __init__(self, net, optimizer_type: str = "Adam", learning_rate: float = 1e-4, momentum: float = 0.9, process_group=None,
                 distributed=None):
        kind = optimizer_type.lower()
        if kind not in ("adam", "sgd", "rmsprop"):
            raise ValueError(optimizer_type)
        self.net, self.kind, self.lr, self.momentum = net, kind, learning_rate, momentum
        self.step_count = 0
        self.flat = self.gflat = None
        self.state = {}
        self._cond_key = self._cond = None
        import torch.distributed as dist
        self.dist = dist if (distributed if distributed is not None else (dist.is_available() and dist.is_initialized())) else None
        self.group = process_group

    def _ensure_flat(self):
        params = list(self.net.parameters())
        _require_cuda(params[0], "module parameters")
        dev = params[0].device
        ok = self.flat is not None and self.flat.device == dev
        if ok:
            off, base = 0, self.flat.data_ptr()
            for p in params:
                if p.data_ptr() != base + 4 * off or p.dtype != torch.float32 or not p.is_contiguous():
                    ok = False
                    break
                off += p.numel()
        if not ok:
            total = sum(p.numel() for p in params)
            flat = torch.empty(total, dtype=torch.float32, device=dev)
            off = 0
            with torch.no_grad():
                for p in params:
                    n = p.numel()
                    flat[off:off + n].copy_(p.detach().reshape(-1).float())
                    p.data = flat[off:off + n].view(p.shape)
                    off += n
            self.flat, self.gflat = flat, torch.zeros(total, dtype=torch.float32, device=dev)
            self.state.clear()
        return params

    def _cond_vector(self, cond_weights, dev):
        net = self.net
        if cond_weights is not None or net.fresh_cond:
            return net._cond_flat(cond_weights, dev).detach()
        key = tuple((c.weight.data_ptr(), c.weight._version, c.bias.data_ptr(), c.bias._version) for c in net.cond_layers)
        if key != self._cond_key or self._cond.device != dev:      # the N + 1 fixed conditioning convs: concatenated once
            self._cond, self._cond_key = net._cond_flat(None, dev).detach(), key
        return self._cond

    def _buf(self, name):
        t = self.state.get(name)
        if t is None or t.device != self.flat.device or t.numel() != self.flat.numel():
            t = torch.zeros_like(self.flat)
            self.state[name] = t
        return t

    def forward_backward(self, piece, target, cond_weights=None):
        import ctypes as C
        net, lib = self.net, L.load()
        self._ensure_flat()
        _require_cuda(piece, "input")
        dev = piece.device
        L.init(dev.index if dev.index is not None else torch.cuda.current_device())
        x = idx = None
        if piece.dtype.is_floating_point:
            x = piece.detach().float().contiguous()
        else:
            idx = piece.detach().to(torch.int64).contiguous()
        src = x if x is not None else idx
        B, Lx = src.shape[0], src.shape[-1]
        W = Lx - net.receptive_field + 1
        if W <= 0:
            raise ValueError("wave sample not long enough")
        h = net._plan()
        cond = self._cond_vector(cond_weights, dev)
        ws = net._workspace(B, Lx, dev, train=True)
        logits = torch.empty(B, net.quantization_channel, W, dtype=torch.float32, device=dev)
        s = L.stream_ptr()
        L.check(lib.wn_ae_forward_train(h, B, Lx, L.ptr(x), L.ptr(idx), L.ptr(self.flat), L.ptr(cond), L.ptr(ws), L.ptr(logits), None, s))
        net._ws_gen += 1
        nb = loss_scratch_bytes(B, W)
        sc = self.state.get("loss_scratch")
        if sc is None or sc.numel() < nb or sc.device != dev:
            sc = self.state["loss_scratch"] = torch.empty(nb, dtype=torch.uint8, device=dev)
        loss, dlogits = fused_loss(logits, target, L.ROWS_REFERENCE, True, 1.0, sc)
        L.check(lib.wn_ae_backward(h, B, Lx, L.ptr(x), L.ptr(idx), L.ptr(self.flat), L.ptr(cond), L.ptr(ws), L.ptr(dlogits),
                                   L.ptr(self.gflat), None, s))
        return loss

    def all_reduce(self):
        if self.dist is not None:
            all_reduce_mean_(self.gflat, self.dist, self.group)

    def apply(self, device_step=None):
        """Optimizer step on the flat vectors.  device_step: int32 device tensor holding the step count (CUDA-graph replay; Adam only)."""
        lib = L.load()
        self.step_count += 1
        n, s = self.flat.numel(), L.stream_ptr()
        if self.kind == "adam" and device_step is not None:
            L.check(lib.wn_adam_step_dev(L.ptr(self.flat), L.ptr(self.gflat), L.ptr(self._buf("m")), L.ptr(self._buf("v")), n,
                                         self.lr, 0.9, 0.999, 1e-8, L.ptr(device_step), s))
        elif self.kind == "adam":
            L.check(lib.wn_adam_step(L.ptr(self.flat), L.ptr(self.gflat), L.ptr(self._buf("m")), L.ptr(self._buf("v")), n,
                                     self.lr, 0.9, 0.999, 1e-8, self.step_count, s))
        elif self.kind == "sgd":
            L.check(lib.wn_sgd_step(L.ptr(self.flat), L.ptr(self.gflat), L.ptr(self._buf("buf")), n, self.lr, self.momentum,
                                    int(self.step_count == 1), s))
        else:
            L.check(lib.wn_rmsprop_step(L.ptr(self.flat), L.ptr(self.gflat), L.ptr(self._buf("sq")), L.ptr(self._buf("buf")), n,
                                        self.lr, 0.99, 1e-8, self.momentum, s))

    def step(self, piece, target, cond_weights=None):
        g = self.__dict__.get("_graph")
        if g is not None and cond_weights is None and piece.shape == g["piece"].shape and piece.dtype == g["piece"].dtype \
                and target.numel() == g["target"].numel():
            g["piece"].copy_(piece, non_blocking=True)
            g["target"].copy_(target.reshape(g["target"].shape), non_blocking=True)
            g["graph"].replay()
            self.step_count += 1
            return g["loss"]
        loss = self.forward_backward(piece, target, cond_weights)
        self.all_reduce()
        self.apply()
        off = 0
        for p in self.net.parameters():
            n = p.numel()
            p.grad = self.gflat[off:off + n].view(p.shape)
            off += n
        return loss

    def capture(self, piece, target, warmup=2):
        """Capture the whole step (forward, loss, backward, [all-reduce], optimizer) for batches of this shape into ONE CUDA graph; later
        step() calls with the same shapes copy the batch into the graph's static buffers and replay it.  At one 64000-target clip per
        GPU the eager step is ~300 launches = 4.1 ms of host time under a 5.2 ms device step; the replay takes the host out of the loop.
        The optimizer state is left exactly as it was (the warm-up steps run on a snapshot that is restored); returns True, or False - with
        the eager path untouched - when the capture is refused (e.g. by the collective library)."""
        self._ensure_flat()
        piece, target = piece.detach(), target.detach()
        snap = {"flat": self.flat.clone(), "count": self.step_count,
                "state": {k: v.clone() for k, v in self.state.items() if torch.is_tensor(v)}}
        d_step = torch.tensor([self.step_count], dtype=torch.int32, device=self.flat.device)
        st = {"piece": piece.clone(), "target": target.clone()}

        def one():
            loss = self.forward_backward(st["piece"], st["target"])
            self.all_reduce()
            self.apply(device_step=d_step if self.kind == "adam" else None)
            return loss

        def restore():
            self.flat.copy_(snap["flat"])
            for k, v in snap["state"].items():
                self.state[k].copy_(v)
            for k in list(self.state):
                if torch.is_tensor(self.state[k]) and k not in snap["state"] and k != "loss_scratch":
                    self.state[k].zero_()
            self.step_count = snap["count"]
            d_step.fill_(snap["count"])
        try:
            side = torch.cuda.Stream(self.flat.device)
            side.wait_stream(torch.cuda.current_stream(self.flat.device))
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    one()
            torch.cuda.current_stream(self.flat.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                st["loss"] = one()
            st["graph"] = graph
        except Exception as e:          # capture refused: stay eager
            import warnings
            warnings.warn(f"music_b200: CUDA-graph capture of the training step failed ({e}); the step stays eager")
            torch.cuda.synchronize()
            restore()
            return False
        restore()
        st["d_step"] = d_step               # the captured optimizer launch reads this tensor on every replay: it must outlive this call
        self.__dict__["_graph"] = st
        off = 0
        for p in self.net.parameters():
            n = p.numel()
            p.grad = self.gflat[off:off + n].view(p.shape)
            off += n
        return True
