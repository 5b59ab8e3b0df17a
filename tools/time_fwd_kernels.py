"""Per-kernel forward timing with the library profiler (timing experiments: WN_DBG / WN_FWD_SIMPLE env toggles)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_b200 import _lib as L
from music_b200.wavenet.model import wavenet
dil = [2 ** i for i in range(10)] * 3
B, W = 16, 16000
net = wavenet(2, dil, 64, 64, 256, 256, False, mode="bf16").cuda()
Lx = net.receptive_field + W - 1
idx = torch.randint(0, 256, (B, Lx)).cuda()
with torch.no_grad():
    for _ in range(3): net.forward_logits(indices=idx)
    torch.cuda.synchronize()
    L.load().wn_profile_enable(1)
    for _ in range(5): net.forward_logits(indices=idx)
    rep = L.profile_report()
print(os.environ.get("WN_DBG", "0"), os.environ.get("WN_FWD_SIMPLE", "0"), [(n, round(ms / 5, 3)) for n, c, ms in rep[:3]])
