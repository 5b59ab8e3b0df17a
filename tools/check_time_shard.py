"""Time-axis sharding on GPUs (run under torchrun, >= 2 ranks): one clip split along time over the ranks, halo exchange of rf input
samples per boundary, weighted gradient average - against the gradient of the whole clip computed on rank 0 (fp32 check mode,
per-time-step objective).  torchrun --nproc-per-node 2 tools/check_time_shard.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from music_b200.wavenet.model import wavenet
from music_b200.wavenet.train import Trainer, time_sharded_step

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
dil = [2 ** i for i in range(8)] * 2
for mode, tol in (("fp32", 2e-4), ("bf16", 6e-2)):
    torch.manual_seed(0)
    net = wavenet(2, dil, 64, 64, 256, 256, False, mode=mode, parity="corrected").cuda()
    rf = net.receptive_field
    T = rf + 6000
    g = torch.Generator().manual_seed(5)
    clip = torch.randint(0, 256, (2, T), generator=g).cuda()
    cut = [0] + [rf + 1500 + (T - rf - 1500) * i // (world - 1) for i in range(world)]
    cut[-1] = T
    # whole clip on every rank (reference), lr = 0 so that the parameters stay put
    ref = Trainer(net, "sgd", 0.0, momentum=0.0, distributed=False)
    ref.forward_backward(clip[:, :-1].contiguous(), clip[:, rf:].contiguous())
    g_ref = net.engine.gflat.clone()
    tr = Trainer(net, "sgd", 0.0, momentum=0.0)
    time_sharded_step(tr, clip[:, cut[rank]:cut[rank + 1]].contiguous())
    g_sh = net.engine.gflat
    err = float((g_sh - g_ref).norm() / g_ref.norm())
    if rank == 0:
        print(f"mode {mode}: time-sharded gradient vs whole clip, rel-l2 {err:.3e} (bound {tol})", flush=True)
    assert err < tol, err
dist.destroy_process_group()
