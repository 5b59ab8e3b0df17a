"""Step time of the generation kernel for models of N blocks (timing experiments: WN_GEN_BPC, WN_GEN_GPC, WN_GEN_PIPE)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_b200.wavenet.model import wavenet
from music_b200.wavenet import fast_generate as fg
steps = 500
for N in [int(x) for x in os.environ.get("LAYERS", "2,6,14,30").split(",")]:
    dil = ([2 ** i for i in range(10)] * 3)[:N]
    net = wavenet(2, dil, 64, 64, 256, 256, False, mode="bf16").cuda()
    for n in [int(x) for x in os.environ.get("STREAMS", "8,64").split(",")]:
        prime = torch.full((n, net.receptive_field), 128, dtype=torch.int64, device="cuda")
        first, st, _ = fg._prime(net, prime)
        out, _ = fg._steps(net, st, first, 50)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out, _ = fg._steps(net, st, first, steps)
        e1.record(); torch.cuda.synchronize()
        print(f"N={N:3d} streams={n:5d}  {e0.elapsed_time(e1) * 1e3 / steps:8.2f} us/step", flush=True)
