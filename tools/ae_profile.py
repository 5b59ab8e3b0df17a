"""Per-kernel breakdown (CUDA events, wn_profile_*) of one autoencoder training step at the configs[4] shape."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_b200 import _lib as L
from music_b200._engine import SoftmaxRowsFunction
from music_b200.wavenet_autoencoder.model1 import wavenet_autoencoder

mode = sys.argv[1] if len(sys.argv) > 1 else "auto"
dil = [2 ** i for i in range(10)] * 4
torch.manual_seed(0)
net = wavenet_autoencoder(2, 256, dil, 32, 32, 512, 512, 32, 32, 512, False, mode=mode).cuda()
W = 64000
idx = torch.randint(0, 256, (1, net.receptive_field + W - 1), device="cuda")
tgt = idx[:, net.receptive_field - 1:].reshape(-1)


def step():
    net.zero_grad()
    probs = SoftmaxRowsFunction.apply(net.forward_logits(indices=idx), L.ROWS_REFERENCE)
    loss = torch.nn.functional.cross_entropy(probs, tgt)
    loss.backward()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
lib = L.load()
lib.wn_profile_enable(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
step()
e1.record()
torch.cuda.synchronize()
rep = L.profile_report()
lib.wn_profile_enable(0)
print("mode", net.mode, "step ms", e0.elapsed_time(e1))
tot = 0.0
for name, cnt, ms in rep:
    print(f"  {name:45s} {cnt:5d} {ms:9.3f} ms")
    tot += ms
print("  sum of labelled kernels", tot)
