"""Per-kernel time of one autoencoder training step at the configs[4] shape (fp32 check mode)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_b200 import _lib as L
from music_b200.wavenet_autoencoder.model1 import wavenet_autoencoder
from music_b200.wavenet_autoencoder import train as T
dil = [2 ** i for i in range(10)] * 4
net = wavenet_autoencoder(2, 256, dil, 32, 32, 512, 512, 32, 32, 512, False).cuda()
W = int(os.environ.get("W", 64000)); Lx = net.receptive_field + W - 1
idx = torch.randint(0, 256, (1, Lx)).cuda()
tgt = idx[:, net.receptive_field - 1:].contiguous()
opt = T.get_optimizer(net, 'Adam', 1e-4)
from music_b200._engine import SoftmaxRowsFunction
def step():
    opt.zero_grad()
    probs = SoftmaxRowsFunction.apply(net.forward_logits(indices=idx), L.ROWS_REFERENCE)
    loss = torch.nn.functional.cross_entropy(probs, tgt.reshape(-1)); loss.backward(); opt.step(); return loss
step(); torch.cuda.synchronize()
lib = L.load(); lib.wn_profile_enable(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
rep = L.profile_report(); lib.wn_profile_enable(0)
print("step ms (with profiling events):", e0.elapsed_time(e1))
for n, c, m in rep[:12]: print(f"{n:28s} {c:5d} launches {m:9.3f} ms")
