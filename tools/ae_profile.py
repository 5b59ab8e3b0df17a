"""Per-kernel breakdown (CUDA events, wn_profile_*) of one fused autoencoder training step at the configs[4] shape."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_b200 import _lib as L
from music_b200.wavenet_autoencoder.model1 import wavenet_autoencoder
from music_b200.wavenet_autoencoder.train import AeTrainer

mode = sys.argv[1] if len(sys.argv) > 1 else "auto"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dil = [2 ** i for i in range(10)] * 4
torch.manual_seed(0)
net = wavenet_autoencoder(2, 256, dil, 32, 32, 512, 512, 32, 32, 512, False, mode=mode).cuda()
W = 64000
idx = torch.randint(0, 256, (B, net.receptive_field + W - 1), device="cuda")
tgt = idx[:, net.receptive_field - 1:].contiguous()
tr = AeTrainer(net, "Adam", 1e-4, distributed=False)
for _ in range(3):
    tr.step(idx, tgt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    loss = tr.step(idx, tgt)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"mode {net.mode} B {B}: {ms:.3f} ms/step = {B * W / ms * 1e3:.3e} samples/s = {B * W * 8.656e6 / ms / 1e9:.1f} TFLOP/s, loss {float(loss):.5f}")
if os.environ.get("GRAPH") == "1":
    ok = tr.capture(idx, tgt)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        loss = tr.step(idx, tgt)
    e1.record()
    torch.cuda.synchronize()
    msg = e0.elapsed_time(e1) / 20
    print(f"captured={ok}: {msg:.3f} ms/step = {B * W * 8.656e6 / msg / 1e9:.1f} TFLOP/s, loss {float(loss):.5f}")
    sys.exit(0)
lib = L.load()
lib.wn_profile_enable(1)
tr.step(idx, tgt)
torch.cuda.synchronize()
rep = L.profile_report()
lib.wn_profile_enable(0)
tot = 0.0
for name, cnt, ms in rep:
    print(f"  {name:45s} {cnt:5d} {ms:9.3f} ms")
    tot += ms
print("  sum of labelled kernels", tot)
