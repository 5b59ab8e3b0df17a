"""clock64 timeline of CTA 1 / group 0 of gen_pipe_kernel.  WN_TS=1 python tools/ts_gen.py [streams]"""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_b200 import _lib as L
from music_b200.wavenet.model import wavenet
from music_b200.wavenet import fast_generate as fg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dil = [2 ** i for i in range(10)] * 3
net = wavenet(2, dil, 64, 64, 256, 256, False, mode="bf16").cuda()
prime = torch.full((n, net.receptive_field), 128, dtype=torch.int64, device="cuda")
first, st, _ = fg._prime(net, prime)
out, _ = fg._steps(net, st, first, 40)
torch.cuda.synchronize()
N = 3 * 16 * 64
buf = (C.c_longlong * N)()
L.check(L.load().wn_debug_ts(buf, -N))
names = ["wait start", "token", "b0 fg mma", "b0 bar1", "b0 dense", "b0 bar2", "b1 fg mma", "b1 bar1", "b1 dense", "x sent", "pushes+skip mma",
         "skip token", "skip barrier", "prefetch", "next pre", "-"]
print("step " + " ".join(f"{i:>6d}" for i in range(15)))
for it in range(20, 30):
    t0 = buf[it * 16 + 1]
    print(f"{it:4d} " + " ".join(f"{buf[it * 16 + k] - t0:6d}" for k in range(15)), " period", buf[it * 16 + 1] - buf[(it - 1) * 16 + 1])
print("columns:", ", ".join(f"{i}={s}" for i, s in enumerate(names)))
print("globaltimer at token arrival, ns after rank 1 (ranks 1..15; rank 15 is the head's skip token)")
for it in range(20, 26):
    g0 = buf[1024 + it * 16 + 1]
    print(f"{it:4d} " + " ".join(f"{buf[1024 + it * 16 + r] - g0:6d}" for r in range(1, 16)), " period ns", g0 - buf[1024 + (it - 1) * 16 + 1])
print("head (clock64 after its skip token): 0=wait start 1=token 2=relu/convert barrier 3=P1 mmas issued 4=P1 barrier 5=P2 mmas issued 6=P2 barrier 7=pick sent")
for it in range(20, 26):
    t0 = buf[2048 + it * 16 + 1]
    print(f"{it:4d} " + " ".join(f"{buf[2048 + it * 16 + k] - t0:6d}" for k in range(8)))
