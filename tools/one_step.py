"""A few bf16 train steps at the cfg-2 shape (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_b200.wavenet.model import wavenet
from music_b200.wavenet.train import Trainer
dil = [2 ** i for i in range(10)] * 3
B = int(os.environ.get("B", 16)); W = 16000
net = wavenet(2, dil, 64, 64, 256, 256, False, mode="bf16").cuda()
L = net.receptive_field + W - 1
idx = torch.randint(0, 256, (B, L + 1)).cuda()
tr = Trainer(net, "adam", distributed=False)
for i in range(int(os.environ.get("STEPS", 2))):
    tr.step(idx[:, :L].contiguous(), idx[:, net.receptive_field:net.receptive_field + W].contiguous())
torch.cuda.synchronize()
print("done")
