"""Time the half-precision generation kernel (cfg 4 shape) for a few stream counts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_b200.wavenet.model import wavenet
from music_b200.wavenet import fast_generate as fg
dil = [2 ** i for i in range(10)] * 3
net = wavenet(2, dil, 64, 64, 256, 256, False, mode="bf16").cuda()
steps = int(os.environ.get("STEPS", 500))
for n in [int(x) for x in os.environ.get("STREAMS", "8,64,512,1024,1184").split(",")]:
    prime = torch.full((n, net.receptive_field), 128, dtype=torch.int64, device="cuda")
    first, st, _ = fg._prime(net, prime)
    out, _ = fg._steps(net, st, first, 50)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out, _ = fg._steps(net, st, first, steps)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / steps
    print(f"streams={n:5d}  {us:8.2f} us/step  {1e6 / us:9.0f} samples/s/stream  {n * 1e6 / us / 1e6:8.2f} M samples/s total", flush=True)
