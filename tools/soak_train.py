"""Soak run: a few hundred bf16 training steps of the cfg-2 model on synthetic (learnable) audio; prints the loss curve.
PARITY=corrected uses the single-softmax, per-time-step objective, whose loss moves visibly; reference = double softmax."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import wavenet_oracle as O
from music_b200.wavenet.model import wavenet
from music_b200.wavenet.train import Trainer
from music_b200.wavenet.audio_func import mu_law_encode
dil = [2 ** i for i in range(10)] * 3
parity = os.environ.get("PARITY", "corrected")
net = wavenet(2, dil, 64, 64, 256, 256, False, mode="bf16", parity=parity).cuda()
B, W = int(os.environ.get("B", 4)), 16000
rf = net.receptive_field; L = rf + W - 1
tr = Trainer(net, "adam", learning_rate=float(os.environ.get("LR", 1e-3)), distributed=False)
g = torch.Generator().manual_seed(0)
steps = int(os.environ.get("STEPS", 300))
losses = []
for it in range(steps):
    t = torch.arange(L + 1, dtype=torch.float32)[None, :] / 16000.0
    f = 110.0 * (1 + torch.randint(0, 6, (B, 1), generator=g).float())
    ph = torch.rand(B, 1, generator=g) * 6.28
    wave = 0.4 * torch.sin(2 * math.pi * f * t + ph) + 0.2 * torch.sin(2 * math.pi * 2 * f * t)
    idx = mu_law_encode(wave.cuda()).to(torch.int64)
    loss = tr.step(idx[:, :L].contiguous(), idx[:, rf:rf + W].contiguous())
    if it % 20 == 0 or it == steps - 1:
        losses.append((it, float(loss)))
        print(f"step {it:4d} loss {float(loss):.4f}", flush=True)
assert all(math.isfinite(l) for _, l in losses)
print("first", losses[0][1], "last", losses[-1][1])
