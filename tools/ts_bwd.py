"""Per-tile clock64 timeline of CTA 0 of block_bwd6 (layer N/2 of the cfg-2 model).  WN_TS=1 WN_BWD6=1 python tools/ts_bwd.py"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_b200 import _lib as L
from music_b200.wavenet.model import wavenet
from music_b200.wavenet.train import Trainer

dil = [2 ** i for i in range(10)] * 3
net = wavenet(2, dil, 64, 64, 256, 256, False, mode="bf16").cuda()
tr = Trainer(net, "adam", distributed=False)
rf = net.receptive_field
idx = torch.randint(0, 256, (16, rf + 16000), device="cuda")
for _ in range(3):
    tr.forward_backward(idx[:, :-1].contiguous(), idx[:, rf:rf + 16000].contiguous())
torch.cuda.synchronize()
n = 16 * 18
buf = (C.c_longlong * n)()
L.check(L.load().wn_debug_ts(buf, n))
names = ["x loads issued", "dx loads issued", "dz committed", "dWd/P/dWfg committed", "f|g committed", "store read done", "E1 start", "dz drained",
         "fg_full passed", "E1 math done", "out_full", "p_full passed", "E2 done"]
t0 = buf[0]
print("tile " + " ".join(f"{i:>7d}" for i in range(13)))
for it in range(17):
    print(f"{it:4d} " + " ".join(f"{buf[it * 16 + k] - t0:7d}" for k in range(13)))
print("columns:", ", ".join(f"{i}={s}" for i, s in enumerate(names)))
