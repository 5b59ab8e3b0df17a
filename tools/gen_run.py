"""A short bf16 fast_generate run at the cfg-4 shape (64 streams) for ncu captures.  STEPS env = steps per launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_b200.wavenet.model import wavenet
from music_b200.wavenet import fast_generate as fg
dil = [2 ** i for i in range(10)] * 3
net = wavenet(2, dil, 64, 64, 256, 256, False, mode="bf16").cuda()
n = int(os.environ.get("STREAMS", 64)); steps = int(os.environ.get("STEPS", 200))
prime = torch.full((n, net.receptive_field), 128, dtype=torch.int64, device="cuda")
first, st, _ = fg._prime(net, prime)
for _ in range(2):
    out, _ = fg._steps(net, st, first, steps)
    first = out[-1].contiguous()
torch.cuda.synchronize()
print("done", tuple(out.shape))
