"""Drop-in for the reference's `wavenet/train.py`: optimizer factory, checkpoint helpers, the
train loop - plus `Trainer`, the fused data-parallel step the loop uses.

Reference: get_optimizer :28-42, save_model :45-50, load_model :53-73, train :76-222 (hot loop
:169-182).  The reference wraps the net in nn.DataParallel (:121); here it is one process per GPU
with ONE NCCL all-reduce (avg) of the flat gradient vector per step, which is arithmetically the
same thing because every rank holds the same number of loss rows.
"""
from __future__ import annotations

import glob
import json
import os
from collections import OrderedDict
from functools import cmp_to_key

import ctypes as C

import torch
import torch.nn as nn
import torch.optim as optim

from .. import _lib as L
from .._engine import fused_loss, loss_scratch_bytes
from .model import wavenet


def get_params(json_dir):
    with open(json_dir, 'r') as f:
        params = json.load(f)
    return params


def get_arguments(base='./params/'):
    train_params = get_params(os.path.join(base, 'train_params.json'))
    wavenet_params = get_params(os.path.join(base, 'wavenet_params.json'))
    dataset_params = get_params(os.path.join(base, 'dataset_params.json'))
    return train_params, wavenet_params, dataset_params


def get_optimizer(model, optimizer_type, learning_rate, momentum):
    """torch.optim objects exactly as the reference builds them (train.py:28-42); they work on the
    module's parameters (views into the flat vector).  `Trainer` below is the fused equivalent."""
    if optimizer_type == 'sgd':
        return optim.SGD(model.parameters(), lr=learning_rate, momentum=momentum)
    if optimizer_type == 'rmsprop':
        return optim.RMSprop(model.parameters(), lr=learning_rate, momentum=momentum)
    if optimizer_type == 'adam':
        return optim.Adam(model.parameters(), lr=learning_rate)


def save_model(model, num_iter, path):
    model_name = "wavenet" + str(num_iter) + ".model"
    checkpoint_path = path + model_name
    print("Storing checkpoint to {} ...".format(path))
    sd = OrderedDict((k, v.detach().cpu().clone()) for k, v in model.state_dict().items())
    torch.save(sd, checkpoint_path)
    print("Done!")


def load_model(model, path, model_name):
    checkpoint_path = path + model_name
    print("Trying to restore saved checkpoint from ", "{}".format(checkpoint_path))
    if os.path.exists(checkpoint_path):
        print("Checkpoint found, restoring!")
        state_dict = torch.load(checkpoint_path, map_location="cpu")
        keys = list(state_dict.keys())
        if keys[0][:6] == 'module':                       # saved from nn.DataParallel (train.py:61-69)
            new_state_dict = OrderedDict()
            for k, v in state_dict.items():
                new_state_dict[k[7:]] = v
            state_dict = new_state_dict
        model.load_state_dict(state_dict)
        return model
    else:
        print("No checkpoint found!")
        return None


def all_reduce_mean_(flat_grad: torch.Tensor, dist=None, group=None) -> torch.Tensor:
    """Average one flat gradient vector over the data-parallel ranks, in place: ONE collective per step
    (NCCL `avg` on GPUs; gloo has no avg, so sum then divide).  Every rank holds the same number of loss
    rows, so the mean of per-rank mean-loss gradients equals the reference's DataParallel gradient of the
    mean loss over the gathered batch (wavenet/train.py:121,178-181)."""
    if dist is None:
        import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return flat_grad
    ws = dist.get_world_size(group)
    if ws == 1:
        return flat_grad
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat_grad, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
        flat_grad.div_(ws)
    return flat_grad


class Trainer:
    """Fused train step: zero_grad -> forward -> CrossEntropyLoss(probabilities) -> backward ->
    [all-reduce] -> optimizer.step, i.e. wavenet/train.py:171-182, with no autograd graph.

    step(piece, target): `piece` is either the dense (B,Q,L) float tensor the reference feeds or a
    (B,L) integer tensor of mu-law codes (true one-hot input); `target` is (B,W) int64.  Returns
    the loss as a 1-element device tensor (no host sync).
    """

    def __init__(self, net: wavenet, optimizer_type: str = "adam", learning_rate: float = 1e-4,
                 momentum: float = 0.9, process_group=None, distributed=None):
        if optimizer_type not in ("adam", "sgd", "rmsprop"):
            raise ValueError(optimizer_type)
        self.net, self.kind, self.lr, self.momentum = net, optimizer_type, learning_rate, momentum
        self.step_count = 0
        self.state = {}
        import torch.distributed as dist
        self.dist = dist if (distributed if distributed is not None else (dist.is_available() and dist.is_initialized())) else None
        self.group = process_group
        self._comm_stream = None

    def _buf(self, name):
        e = self.net.engine
        t = self.state.get(name)
        if t is None or t.device != e.flat.device or t.numel() != e.n_params:
            t = torch.zeros(e.n_params, dtype=torch.float32, device=e.flat.device)
            self.state[name] = t
        return t

    def forward_backward(self, piece, target):
        net, e = self.net, self.net.engine
        params = net._params()
        e.ensure_flat(params)
        mode, rows = L.MODES[net.mode], L.ROWS[net.parity]
        packed = e.packed(mode, params)
        x = idx = None
        if piece.dtype.is_floating_point:
            x = piece.detach().float().contiguous()
        else:
            idx = piece.detach().to(torch.int64).contiguous()
        src = x if x is not None else idx
        ws = e.workspace(mode, src.shape[0], src.shape[-1])
        logits = e.forward_logits(mode, x, idx, packed, ws)
        loss, dlogits = fused_loss(logits, target, rows, True, 1.0, e.scratch("loss", loss_scratch_bytes(logits.shape[0], logits.shape[2])))
        e.backward(mode, x, idx, packed, ws, dlogits, e.gflat)
        return loss

    # ---- overlapped gradient exchange -------------------------------------------------------------------------------------
    # The backward produces the gradients top-down: head and skip weights first, then blocks N-1 .. 0.  With two buckets split at
    # block K = N / 2 the upper bucket [offset(K), n_params) is final when the backward is half way through; the library records an
    # event at that point (wn_backward_set_split) and its all-reduce runs on a second stream under the lower blocks' kernels.  Only
    # the lower bucket's exchange (and the slowest rank's skew) is left exposed after the last backward kernel.
    def _overlap_setup(self):
        # opt-in (WN_AR_OVERLAP=1, set before the first backward): measured on 2 and 8 GPUs the two-bucket exchange is no faster than
        # one all-reduce (the persistent block kernels leave NCCL no SM before they drain), and the per-layer reductions it needs
        # on the side stream cost the single-GPU step ~1 %
        if self.dist is None or self.dist.get_world_size(self.group) == 1 or os.environ.get("WN_AR_OVERLAP", "0") != "1":
            return False
        os.environ.setdefault("WN_WGRAD_SIDE", "1")
        if self._comm_stream is None:
            e, lib = self.net.engine, L.load()
            n_layers = len(self.net.dilations)
            self._split_layer = n_layers // 2
            self._split_off = int(lib.wn_model_layer_offset(e.handle, self._split_layer))
            self._comm_stream = torch.cuda.Stream(e.flat.device)
            self._split_event = torch.cuda.Event()
            self._split_event.record()          # (creates the underlying cudaEvent_t)
            L.check(lib.wn_backward_set_split(e.handle, self._split_layer, C.c_void_p(self._split_event.cuda_event)))
        return self._split_off > 0

    def all_reduce(self):
        if self.dist is None:
            return
        g = self.net.engine.gflat
        if not self._overlap_setup():
            all_reduce_mean_(g, self.dist, self.group)
            return
        main = torch.cuda.current_stream(g.device)
        self._comm_stream.wait_event(self._split_event)            # recorded inside the backward that was enqueued just now
        with torch.cuda.stream(self._comm_stream):
            all_reduce_mean_(g[self._split_off:], self.dist, self.group)
        all_reduce_mean_(g[:self._split_off], self.dist, self.group)
        main.wait_stream(self._comm_stream)

    def apply(self, device_step=None):
        """Optimizer step on the flat vectors.  device_step: int32 device tensor holding the step count (CUDA-graph replay; Adam only)."""
        e, lib = self.net.engine, L.load()
        self.step_count += 1
        n, s = e.n_params, L.stream_ptr()
        if self.kind == "adam" and device_step is not None:
            L.check(lib.wn_adam_step_dev(L.ptr(e.flat), L.ptr(e.gflat), L.ptr(self._buf("m")), L.ptr(self._buf("v")), n,
                                         self.lr, 0.9, 0.999, 1e-8, L.ptr(device_step), s))
        elif self.kind == "adam":
            L.check(lib.wn_adam_step(L.ptr(e.flat), L.ptr(e.gflat), L.ptr(self._buf("m")), L.ptr(self._buf("v")), n,
                                     self.lr, 0.9, 0.999, 1e-8, self.step_count, s))
        elif self.kind == "sgd":
            L.check(lib.wn_sgd_step(L.ptr(e.flat), L.ptr(e.gflat), L.ptr(self._buf("buf")), n, self.lr, self.momentum,
                                    int(self.step_count == 1), s))
        else:
            L.check(lib.wn_rmsprop_step(L.ptr(e.flat), L.ptr(e.gflat), L.ptr(self._buf("sq")), L.ptr(self._buf("buf")), n,
                                        self.lr, 0.99, 1e-8, self.momentum, s))
        e.invalidate_packed()

    def step(self, piece, target):
        g = self.__dict__.get("_graph")
        if g is not None and piece.shape == g["piece"].shape and piece.dtype == g["piece"].dtype and target.shape == g["target"].shape:
            g["piece"].copy_(piece, non_blocking=True)
            g["target"].copy_(target, non_blocking=True)
            g["graph"].replay()
            self.step_count += 1
            return g["loss"]
        loss = self.forward_backward(piece, target)
        self.all_reduce()
        self.apply()
        params = self.net._params()
        for p, g in zip(params, self.net.engine.grad_views(params)):
            p.grad = g
        return loss

    def capture(self, piece, target, warmup=2):
        """Capture the whole step (weight pack, forward, loss, backward, [all-reduce], optimizer) for batches of this shape into ONE CUDA
        graph; later step() calls with the same shapes copy the batch into the graph's static buffers and replay it.  At the
        benchmarked shape the eager step's 80 launches take 1.9 ms of host time under a 5.2 ms device step, so the replay buys nothing
        there; it matters for small batches (one clip: 126 tiles per layer).  The optimizer state is left exactly as it was (the warm-up
        steps run on a snapshot that is restored); returns True, or False - eager path untouched - when the capture is refused."""
        e = self.net.engine
        e.ensure_flat(self.net._params())
        snap = {"flat": e.flat.clone(), "count": self.step_count, "state": {k: v.clone() for k, v in self.state.items()}}
        d_step = torch.tensor([self.step_count], dtype=torch.int32, device=e.flat.device)
        st = {"piece": piece.detach().clone(), "target": target.detach().clone()}

        def one():
            loss = self.forward_backward(st["piece"], st["target"])
            self.all_reduce()
            self.apply(device_step=d_step if self.kind == "adam" else None)
            return loss

        def restore():
            e.flat.copy_(snap["flat"])
            for k, v in self.state.items():
                if k in snap["state"]:
                    v.copy_(snap["state"][k])
                else:
                    v.zero_()
            self.step_count = snap["count"]
            d_step.fill_(snap["count"])
            e.invalidate_packed()
        try:
            side = torch.cuda.Stream(e.flat.device)
            side.wait_stream(torch.cuda.current_stream(e.flat.device))
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    one()
            torch.cuda.current_stream(e.flat.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                st["loss"] = one()
            st["graph"] = graph
        except Exception as ex:         # capture refused: stay eager
            import warnings
            warnings.warn(f"music_b200: CUDA-graph capture of the training step failed ({ex}); the step stays eager")
            torch.cuda.synchronize()
            restore()
            return False
        restore()
        st["d_step"] = d_step               # the captured optimizer launch reads this tensor on every replay: it must outlive this call
        self.__dict__["_graph"] = st
        params = self.net._params()
        for p, g in zip(params, e.grad_views(params)):
            p.grad = g
        return True

    # ---- optimizer state (the reference checkpoints the model only, train.py:44-50; resuming Adam from zero moments costs a
    #      few hundred steps of re-warm-up, so the fused optimizer's state can be saved next to the `.model` file) ---------------
    def state_dict(self):
        """{'kind', 'step', 'lr', 'momentum', buffers...}: the flat moment vectors on the CPU (parameter order = state_dict order)."""
        out = {"kind": self.kind, "step": self.step_count, "lr": self.lr, "momentum": self.momentum}
        for k, v in self.state.items():
            out["buf." + k] = v.detach().cpu().clone()
        return out

    def load_state_dict(self, sd):
        if sd["kind"] != self.kind:
            raise ValueError(f"optimizer state is for '{sd['kind']}', this trainer runs '{self.kind}'")
        e = self.net.engine
        e.ensure_flat(self.net._params())
        self.step_count = int(sd["step"])
        for k, v in sd.items():
            if k.startswith("buf."):
                if v.numel() != e.n_params:
                    raise ValueError(f"optimizer buffer '{k[4:]}' has {v.numel()} values, the model {e.n_params} parameters")
                self._buf(k[4:]).copy_(v.to(e.flat.device))


def save_optimizer(trainer, num_iter, path):
    """`wavenet<N>.optim` next to `wavenet<N>.model` (same numbering as save_model)."""
    torch.save(trainer.state_dict(), path + "wavenet" + str(num_iter) + ".optim")


def load_optimizer(trainer, path, model_name):
    """Restore the optimizer state saved with the checkpoint `model_name` (`wavenet<N>.model`); False when there is none."""
    f = path + model_name.rsplit(".", 1)[0] + ".optim"
    if not os.path.exists(f):
        return False
    trainer.load_state_dict(torch.load(f, map_location="cpu"))
    return True


# ---- time-axis sharding ----------------------------------------------------------------------------------
def exchange_time_halo(codes, rf, dist=None, group=None):
    """One long clip split along TIME over the ranks: rank r holds the contiguous slice `codes` (B, T_r) of mu-law codes.  The
    network is causal with receptive field rf: the prediction of a slice's first sample needs the rf samples before it (rf - 1 of
    context plus the sample the first window ends on), so each rank sends the last rf codes of its slice to rank r + 1 and receives
    its left neighbour's - one send / recv pair per boundary; rank 0 has no left context, its first rf samples are context only,
    exactly as at the start of an unsharded clip.  Returns the (B, rf + T_r) piece this rank trains on (rank 0: its slice unchanged);
    together the ranks then cover every target of the clip exactly once.  No activation crosses ranks: the halo is INPUT, recomputed
    through the stack.  Works on any backend (NCCL on GPUs, gloo in the CPU test)."""
    if dist is None:
        import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if world == 1:
        return codes
    h = rf
    if codes.shape[-1] < h:
        raise ValueError(f"a time slice of {codes.shape[-1]} samples is shorter than the halo ({h})")
    ops, halo = [], None
    if rank + 1 < world:
        ops.append(dist.P2POp(dist.isend, codes[..., -h:].contiguous(), rank + 1, group))
    if rank > 0:
        halo = torch.empty(codes.shape[:-1] + (h,), dtype=codes.dtype, device=codes.device)
        ops.append(dist.P2POp(dist.irecv, halo, rank - 1, group))
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    return codes if halo is None else torch.cat([halo, codes], dim=-1)


def time_sharded_step(trainer, codes, rf=None):
    """One training step on a clip that spans the ranks along time (exchange_time_halo + the data-parallel step): every rank's
    targets are its own slice (rank 0: its slice minus the first rf samples); each rank's gradient of its mean loss is weighted by
    its share of the clip's targets before the flat gradient is averaged, so the result is the gradient of the mean loss over ALL
    targets of the clip, whatever the slice lengths.  Needs the per-time-step objective (`parity="corrected"`): the reference's
    flat-chunk softmax rows mix time steps of the whole (B, Q, W) tensor and have no time-local definition.  Returns this rank's loss."""
    net = trainer.net
    if net.parity != "corrected":
        raise L.WavenetB200Error("time-axis sharding needs parity='corrected' (the reference objective's softmax rows are not local in time)")
    rf = net.receptive_field if rf is None else rf
    piece = exchange_time_halo(codes, rf, trainer.dist, trainer.group) if trainer.dist is not None else codes
    W = piece.shape[-1] - rf                      # targets: sample t + 1 for every window ending at t (as audio_data_loader pairs them)
    loss = trainer.forward_backward(piece[..., :-1].contiguous(), piece[..., rf:rf + W].contiguous())
    if trainer.dist is not None and trainer.dist.get_world_size(trainer.group) > 1:
        key = (int(W), tuple(codes.shape))
        cache = trainer.__dict__.setdefault("_time_shard_weight", {})
        if key not in cache:                      # this rank's share of the targets x world (one scalar all-reduce per new shape)
            n = torch.tensor([float(W)], device=codes.device)
            trainer.dist.all_reduce(n, group=trainer.group)
            cache[key] = float(W) * trainer.dist.get_world_size(trainer.group) / float(n[0])
        trainer.net.engine.gflat.mul_(cache[key])
    trainer.all_reduce()
    trainer.apply()
    return loss


class BatchStager:
    """Host -> device staging of (audio_piece, audio_target) batches under the running train step: two device slots,
    copies on their own stream, so the copy of batch i + 1 overlaps step i (what `pin_memory=True` + `.cuda(non_blocking)`
    in the reference loop, train.py:186-194, is after - there the copy sits in the compute stream and serialises).

        stager.put(piece, target)          # start the copy of the next batch (pinned host tensors copy asynchronously)
        piece_d, target_d = stager.get()   # the compute stream waits for that copy
        ... trainer.step(piece_d, target_d) ...
        stager.done()                      # the step that used the oldest slot has been enqueued: the slot may be refilled
    """

    def __init__(self, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.slots = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [None, None]
        self.n_put = self.n_get = self.n_done = 0

    def put(self, piece, target):
        if self.n_put - self.n_done >= 2:
            raise L.WavenetB200Error("BatchStager: both slots are in use (call done() after the step that consumed a batch)")
        i = self.n_put % 2
        slot = self.slots[i]
        if (slot is None or slot[0].shape != piece.shape or slot[1].shape != target.shape or slot[0].dtype != piece.dtype
                or slot[1].dtype != target.dtype):
            # A new slot (first use, or a batch of another shape: the short last batch of an epoch).  The caching allocator
            # may hand out a block the compute stream has freed on the host while its kernels are still running, so the
            # copy stream first catches up with everything enqueued on the compute stream; the old slot, if any, may still
            # be read by the step in flight and is kept alive for it by record_stream.
            compute = torch.cuda.current_stream(self.device)
            if slot is not None:
                for t in slot:
                    t.record_stream(compute)
            slot = self.slots[i] = (torch.empty(piece.shape, dtype=piece.dtype, device=self.device),
                                    torch.empty(target.shape, dtype=target.dtype, device=self.device))
            self.copy_stream.wait_stream(compute)
            for t in slot:
                t.record_stream(self.copy_stream)
            self.free[i] = None
        with torch.cuda.stream(self.copy_stream):
            if self.free[i] is not None:
                self.copy_stream.wait_event(self.free[i])     # the step that read this slot two batches ago
            slot[0].copy_(piece, non_blocking=True)
            slot[1].copy_(target, non_blocking=True)
            self.ready[i].record(self.copy_stream)
        self.n_put += 1

    def get(self):
        if self.n_get >= self.n_put:
            raise L.WavenetB200Error("BatchStager: get() without a staged batch")
        i = self.n_get % 2
        torch.cuda.current_stream(self.device).wait_event(self.ready[i])
        self.n_get += 1
        return self.slots[i]

    def done(self):
        i = self.n_done % 2
        if self.free[i] is None:
            self.free[i] = torch.cuda.Event()
        self.free[i].record(torch.cuda.current_stream(self.device))
        self.n_done += 1


def _setup_data_parallel(train_params):
    """One process per GPU (what replaces nn.DataParallel, train.py:117-122).  Under torchrun (WORLD_SIZE > 1) each process
    takes cuda:LOCAL_RANK and joins the NCCL group; returns (rank, world).  A multi-entry `device_ids` without torchrun cannot
    be honoured by a single process and is refused instead of being ignored silently."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ids = train_params.get("device_ids") or []
    if world == 1:
        if len(ids) > 1:
            raise L.WavenetB200Error(
                "train_params['device_ids'] names %d GPUs: music_b200 runs data parallelism as one process per GPU - launch "
                "`torchrun --nproc-per-node %d` (the batch is then split over the ranks as nn.DataParallel would split it)"
                % (len(ids), len(ids)))
        if len(ids) == 1:
            torch.cuda.set_device(int(ids[0]))
        return 0, 1
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return dist.get_rank(), dist.get_world_size()


class _RankShard:
    """The slice of every batch this rank trains on: rows [rank * B / world, (rank + 1) * B / world) - the split
    nn.DataParallel applies to a batch (train.py:117-122 asserts batch_size % num_gpu == 0); batches that do not divide
    (the short last batch of an epoch) are dropped on every rank alike so that the ranks stay in step."""

    def __init__(self, loader, rank, world):
        self.loader, self.rank, self.world = loader, rank, world

    def __iter__(self):
        for b in self.loader:
            n = b["audio_piece"].shape[0]
            if n % self.world:
                continue
            k = n // self.world
            sl = slice(self.rank * k, (self.rank + 1) * k)
            yield {"audio_piece": b["audio_piece"][sl], "audio_target": b["audio_target"][sl]}


def train(base='./params/', dataloader=None, rank=None):
    """The reference loop (train.py:76-222): JSON configs, resume, loss / store logs, checkpoint rotation.
    One process per GPU: run it directly for one GPU, or under `torchrun --nproc-per-node N` for data parallelism -
    every rank then trains its slice of each batch, the flat gradient is all-reduced (NCCL avg) inside `Trainer.step`,
    rank 0 alone writes logs and checkpoints, and all ranks meet at a barrier around the checkpoint rotation."""
    from .faster_audio_data import audio_data_loader, one_hot_encode_device
    if not torch.cuda.is_available():
        raise L.WavenetB200Error("music_b200 trains on a B200 only (no CPU fallback)")
    train_params, wavenet_params, dataset_params = get_arguments(base)
    ddp_rank, world = _setup_data_parallel(train_params)
    rank = ddp_rank if rank is None else rank
    net = wavenet(**wavenet_params)
    epoch_trained = 0
    if train_params["restore_model"]:
        restored = load_model(net, train_params["restore_dir"], train_params["restore_model"])
        if restored is None:
            print("Initialize network and train from scratch.")
        else:
            epoch_trained = int(train_params["restore_model"].split('.')[0][7:])
    if dataloader is None:
        # the dataset hands out integer codes; the reference's (Q, T) "one-hot" (reshape quirk and all) is built on the
        # device from them (one_hot_encode_device), so 8 bytes per sample cross PCIe instead of 1 KB
        dataset_params = dict(dataset_params)
        dataset_params.setdefault("encoding", "codes")
        dataloader = audio_data_loader(**dataset_params)
    device_encoding = getattr(getattr(dataloader, "dataset", None), "encoding", None) == "codes"
    if world > 1:
        dataloader = _RankShard(dataloader, rank, world)
    net = net.cuda()
    if world > 1:                       # identical replicas (restored or freshly initialised on rank 0)
        import torch.distributed as dist
        for p in net.parameters():
            dist.broadcast(p.data, 0)
    print("Start training.")
    print("Writing logging information to ", "{}".format(train_params["log_dir"]))
    print("Models are saved in {}".format(train_params["restore_dir"]))
    trainer = Trainer(net, train_params["optimizer"], train_params["learning_rate"], train_params["momentum"])
    num_trained = 0
    loss_log_file = store_log_file = None
    if rank == 0:
        os.makedirs(train_params["log_dir"], exist_ok=True)
        os.makedirs(train_params["restore_dir"], exist_ok=True)
        loss_log_file = open(train_params["log_dir"] + 'loss_log.log', 'a')
        store_log_file = open(train_params["log_dir"] + 'store_log.log', 'a')
        with open(train_params["log_dir"] + 'loss_log.log', 'r') as f:
            lines = f.readlines()
            num_trained = int(lines[-1].split(' ')[2]) if len(lines) > 0 else 0
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([num_trained], device="cuda")
        dist.broadcast(t, 0)
        num_trained = int(t[0])
    total_loss = torch.zeros(1, device="cuda")
    stager = BatchStager()
    for epoch in range(train_params["num_epochs"]):
        batches = iter(dataloader)
        nxt = next(batches, None)
        if nxt is not None:
            stager.put(nxt["audio_piece"], nxt["audio_target"])
        while nxt is not None:
            piece, target = stager.get()
            nxt = next(batches, None)
            if nxt is not None:                                # staged under the step below
                stager.put(nxt["audio_piece"], nxt["audio_target"])
            if device_encoding:
                piece = one_hot_encode_device(piece, net.quantization_channels)
            total_loss += trainer.step(piece, target)         # accumulated on device: no per-step sync
            stager.done()
            num_trained += 1
            if num_trained % train_params["print_every"] == 0:
                avg_loss = float(total_loss) / train_params["print_every"]
                line = "Trained over " + str(num_trained) + " pieces," + "Average loss is " + str(avg_loss) + "\n"
                if rank == 0:
                    loss_log_file.writelines(line)
                    loss_log_file.flush()
                total_loss.zero_()
        if (epoch + 1) % train_params["check_point_every"] == 0 and world > 1:
            import torch.distributed as dist
            dist.barrier()              # every rank has finished the epoch before rank 0 rotates / writes checkpoints
        if (epoch + 1) % train_params["check_point_every"] == 0 and rank == 0:
            stored_models = glob.glob(train_params["restore_dir"] + "*.model")
            if len(stored_models) == train_params["max_check_points"]:
                def cmp(x, y):
                    x = int(x.split('/')[-1].split('.')[0][7:])
                    y = int(y.split('/')[-1].split('.')[0][7:])
                    return x - y
                stored_models = sorted(stored_models, key=cmp_to_key(cmp))
                os.remove(stored_models[0])
            save_model(net, epoch_trained + epoch + 1, train_params["restore_dir"])
            store_log_file.writelines("Epoch " + str(epoch_trained + epoch + 1) + ", model saved!\n")
            store_log_file.flush()
    if rank == 0:
        loss_log_file.close()
        store_log_file.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    return net


if __name__ == '__main__':
    train()
