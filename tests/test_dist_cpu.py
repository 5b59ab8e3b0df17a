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
