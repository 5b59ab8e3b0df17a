"""Host-side enqueue time vs device time of one fused autoencoder step (is the step launch-bound?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_b200.wavenet_autoencoder.model1 import wavenet_autoencoder
from music_b200.wavenet_autoencoder.train import AeTrainer
from music_b200.wavenet.model import wavenet
from music_b200.wavenet.train import Trainer

def measure(name, step, n=20):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name}: host enqueue {(t1 - t0) / n * 1e3:.3f} ms/step, device-complete {(t2 - t0) / n * 1e3:.3f} ms/step", flush=True)

dil = [2 ** i for i in range(10)] * 4
net = wavenet_autoencoder(2, 256, dil, 32, 32, 512, 512, 32, 32, 512, False, mode="auto").cuda()
idx = torch.randint(0, 256, (1, net.receptive_field + 64000 - 1), device="cuda")
tgt = idx[:, net.receptive_field - 1:].contiguous()
tr = AeTrainer(net, "Adam", 1e-4, distributed=False)
measure("autoencoder 1 clip", lambda: tr.step(idx, tgt))
dil3 = [2 ** i for i in range(10)] * 3
wn = wavenet(2, dil3, 64, 64, 256, 256, False, mode="bf16").cuda()
rf = wn.receptive_field
x = torch.randint(0, 256, (16, rf + 16000), device="cuda")
tw = Trainer(wn, "adam", distributed=False)
p, t = x[:, :-1].contiguous(), x[:, rf:].contiguous()
measure("wavenet cfg 2", lambda: tw.step(p, t))
