"""Smallest bf16 forward, for compute-sanitizer / WN_DEBUG_SYNC runs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import wavenet_oracle as O
from music_b200.wavenet.model import wavenet
dil = [1, 2, 4]
net = wavenet(2, dil, 64, 64, 256, 256, False, mode="bf16").cuda()
rf = O.receptive_field(2, dil)
idx = torch.randint(0, 256, (1, rf + 200 - 1)).cuda()
with torch.no_grad():
    lg = net.forward_logits(indices=idx)
torch.cuda.synchronize()
print("ok", lg.shape, float(lg.abs().max()))
