"""Two-bucket overlapped gradient exchange vs the single all-reduce (run under torchrun, >= 2 GPUs):
same parameters after a few Adam steps, and the step times of both.  torchrun --nproc-per-node 2 tools/check_overlap.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from music_b200.wavenet.model import wavenet
from music_b200.wavenet.train import Trainer

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
dil = [2 ** i for i in range(10)] * 3
res = {}
os.environ["WN_WGRAD_SIDE"] = "1"      # (read once by the library: the overlapped variant needs the per-layer reductions)
for overlap in ("0", "1"):
    os.environ["WN_AR_OVERLAP"] = overlap
    torch.manual_seed(0)
    net = wavenet(2, dil, 64, 64, 256, 256, False, mode="bf16").cuda()
    tr = Trainer(net, "adam", 1e-3)
    g = torch.Generator().manual_seed(100 + rank)
    rf = net.receptive_field
    idx = torch.randint(0, 256, (16, rf + 16000), generator=g).cuda()
    piece, target = idx[:, :-1].contiguous(), idx[:, rf:rf + 16000].contiguous()
    for _ in range(3):
        tr.step(piece, target)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        tr.step(piece, target)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 20], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[overlap] = (float(t[0]), net.engine.flat.clone())
    chk = net.engine.flat.clone()
    dist.broadcast(chk, 0)
    same = bool(torch.equal(chk, net.engine.flat))
    if rank == 0:
        print(f"overlap={overlap}: {float(t[0]):.3f} ms/step (max over {world} ranks), replicas identical: {same}", flush=True)
if rank == 0:
    d = (res["0"][1] - res["1"][1]).abs().max().item()
    print(f"max |param difference| between the two exchange schemes after 23 Adam steps: {d:.3e}")
dist.destroy_process_group()
