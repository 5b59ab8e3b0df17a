"""CPU, world_size 2 over gloo: the data-parallel reduction used by Trainer (one all-reduce of the flat
gradient) reproduces the full-batch gradient of the reference objective."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import wavenet_oracle as O

DIL = [1, 2, 4, 8]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _case():
    st = O.init_wavenet_state(DIL, 8, 8, 16, 256, True, seed=3, scale=2.0)
    rf = O.receptive_field(2, DIL)
    W, B = 21, 4
    g = torch.Generator().manual_seed(3)
    idx = torch.randint(0, 256, (B, rf + W), generator=g)
    return st, idx[:, :-1], idx[:, rf:rf + W].contiguous()


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from music_b200.wavenet.train import all_reduce_mean_
    st, idx, tgt = _case()
    per = idx.shape[0] // world
    sl = slice(rank * per, (rank + 1) * per)
    loss, grads = O.grads(st, DIL, O.one_hot(idx[sl], 256), tgt[sl])
    flat = torch.cat([g.reshape(-1) for g in grads.values()])
    all_reduce_mean_(flat)
    losses = torch.tensor([loss])
    dist.all_reduce(losses)
    if rank == 0:
        np.save(out, np.concatenate([[float(losses[0]) / world], flat.numpy()]))
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_equals_full_batch(tmp_path):
    out = str(tmp_path / "g.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    st, idx, tgt = _case()
    loss, grads = O.grads(st, DIL, O.one_hot(idx, 256), tgt)
    ref = torch.cat([g.reshape(-1) for g in grads.values()]).numpy()
    assert abs(got[0] - loss) < 1e-6
    np.testing.assert_allclose(got[1:], ref, rtol=1e-4, atol=1e-9)


def test_allreduce_is_identity_without_process_group():
    from music_b200.wavenet.train import all_reduce_mean_
    t = torch.arange(5.0)
    assert torch.equal(all_reduce_mean_(t.clone()), t)


# ---- autoencoder: all_reduce_grads_ of wavenet_autoencoder/train.py on two gloo ranks ------------------------------
AE_DIL = [1, 2, 4, 1, 2, 4]
AE_CFG = dict(Re=8, De=8, BW=8, pool=4, Rd=8, Dd=8, Sd=16)


def _ae_case():
    torch.manual_seed(7)
    shapes = O.ae_param_shapes(AE_DIL, AE_CFG["Re"], AE_CFG["De"], AE_CFG["BW"], AE_CFG["Rd"], AE_CFG["Dd"], AE_CFG["Sd"],
                               quantization_channel=256, use_bias=True)
    st = {k: (0.3 * torch.randn(*s)).requires_grad_(True) for k, s in shapes}
    cond = {k: 0.3 * torch.randn(*s) for k, s in O.ae_cond_shapes(len(AE_DIL), AE_CFG["BW"], AE_CFG["Dd"], AE_CFG["Sd"])}
    rf = O.receptive_field(2, AE_DIL)
    W, B = 24, 4
    idx = torch.randint(0, 256, (B, rf + W - 1))
    tgt = torch.randint(0, 256, (B, W))
    return st, cond, idx, tgt


def _ae_grads(st, cond, idx, tgt):
    for v in st.values():
        v.grad = None
    probs = O.ae_forward_probs(st, cond, AE_DIL, O.one_hot(idx, 256), AE_CFG["pool"])
    loss = torch.nn.functional.cross_entropy(probs, tgt.reshape(-1))
    loss.backward()
    return float(loss)


def _ae_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from music_b200.wavenet_autoencoder.train import all_reduce_grads_
    st, cond, idx, tgt = _ae_case()
    per = idx.shape[0] // world
    sl = slice(rank * per, (rank + 1) * per)
    _ae_grads(st, cond, idx[sl], tgt[sl])
    params = [v for v in st.values()]            # the last block's dense conv has no gradient: must be skipped, not crash
    all_reduce_grads_(params)
    if rank == 0:
        np.save(out, torch.cat([p.grad.reshape(-1) for p in params if p.grad is not None]).numpy())
    dist.destroy_process_group()


def test_autoencoder_two_rank_gloo_gradient_average_equals_full_batch(tmp_path):
    out = str(tmp_path / "ae.npy")
    mp.spawn(_ae_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    st, cond, idx, tgt = _ae_case()
    _ae_grads(st, cond, idx, tgt)
    ref = torch.cat([p.grad.reshape(-1) for p in st.values() if p.grad is not None]).numpy()
    np.testing.assert_allclose(got, ref, rtol=2e-4, atol=1e-8)


# ------------------------------------------------------------------------------------------ time-axis sharding
def _time_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from music_b200.wavenet.train import all_reduce_mean_, exchange_time_halo
    st, clip, bounds = _time_case(world)
    rf = O.receptive_field(2, DIL)
    mine = clip[:, bounds[rank]:bounds[rank + 1]].contiguous()
    piece = exchange_time_halo(mine, rf)
    # this rank's targets under the per-time-step objective: one cross entropy per time step over the Q logits (parity="corrected")
    W = piece.shape[1] - rf
    x = O.one_hot(piece[:, :-1], 256)
    tgt = piece[:, rf:rf + W]
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in st.items()}
    logits = O.forward_logits(leaves, DIL, x)                                   # (B, Q, W)
    loss_sum = torch.nn.functional.cross_entropy(logits, tgt, reduction="sum")
    loss_sum.backward()
    flat = torch.cat([(torch.zeros_like(leaves[k]) if leaves[k].grad is None else leaves[k].grad).reshape(-1) for k in st])
    n = torch.tensor([float(tgt.numel())])
    dist.all_reduce(flat)                                                       # sums of per-target gradients ...
    dist.all_reduce(n)                                                          # ... over all targets of the clip
    ok = torch.tensor([1.0 if torch.equal(piece, clip[:, max(bounds[rank] - rf, 0):bounds[rank + 1]]) else 0.0])
    dist.all_reduce(ok)
    if rank == 0:
        np.save(out, np.concatenate([[float(ok[0]), float(n[0])], (flat / n).numpy()]))
    dist.destroy_process_group()


def _time_case(world):
    st = O.init_wavenet_state(DIL, 8, 8, 16, 256, True, seed=5, scale=2.0)
    rf = O.receptive_field(2, DIL)
    T = rf + 90
    g = torch.Generator().manual_seed(9)
    clip = torch.randint(0, 256, (2, T), generator=g)
    first = rf + 29                       # rank 0's slice contains the clip's first rf context samples
    rest = (T - first) // (world - 1)
    bounds = [0, first] + [first + rest * (i + 1) for i in range(world - 1)]
    bounds[-1] = T
    return st, clip, bounds


def test_time_axis_sharding_halo_exchange_equals_whole_clip(tmp_path):
    """A clip split along time over 3 ranks: after exchange_time_halo every rank holds its slice plus the rf samples before it,
    and the target-weighted sum of the per-rank gradients equals the gradient of the whole clip (per-time-step objective)."""
    world = 3
    out = str(tmp_path / "t.npy")
    mp.spawn(_time_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = np.load(out)
    st, clip, bounds = _time_case(world)
    rf = O.receptive_field(2, DIL)
    assert got[0] == world                                   # every rank's piece is the right window of the clip
    W = clip.shape[1] - rf
    assert got[1] == 2 * W                                   # together the ranks cover every target exactly once
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in st.items()}
    logits = O.forward_logits(leaves, DIL, O.one_hot(clip[:, :-1], 256))
    torch.nn.functional.cross_entropy(logits, clip[:, rf:rf + W], reduction="mean").backward()
    ref = torch.cat([(torch.zeros_like(leaves[k]) if leaves[k].grad is None else leaves[k].grad).reshape(-1) for k in st]).numpy()
    np.testing.assert_allclose(got[2:], ref, rtol=2e-4, atol=1e-8)      # (the last block's dense conv is unused: zero gradient)
