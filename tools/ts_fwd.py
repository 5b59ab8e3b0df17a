"""Phase timestamps of block_fwd2 (CTA 0, middle layer): WN_TS=1."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["WN_TS"] = "1"
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
e = net.engine
ws = e._ws[(1, B, Lx)]
# X0f offset: query via layout arithmetic is internal; find it as the region after H1: use the C layout through sizes
def al(x): return (x + 1023) // 1024 * 1024
N = 30; Wp = Lx - ((net.receptive_field - 1) // 128) * 128
xs = al(B * Lx * 64 * 2)
off = xs * N + xs * 2 + al(B * Wp * 64 * N * 2) + 2 * al(B * Wp * 256 * 2)
ts = ws[off:off + 16 * 8 * 8].view(torch.int64).view(16, 8).cpu()
mm = ws[off + 1024 * 8: off + 1024 * 8 + 16 * 4 * 8].view(torch.int64).view(16, 4).cpu()
t0 = int(ts[0][0])
for t in range(15):
    print(f"T{t:2d}: load_issued {int(mm[t][2])-t0:7d}  M1_issued {int(mm[t][0])-t0:7d}  epi_start {int(ts[t][0])-t0:7d}  fg_seen {int(ts[t][1])-t0:7d}  z_ready~ {int(ts[t][3])-t0:7d}  M2_issued {int(mm[t][1])-t0:7d}  dense_seen {int(ts[t][4])-t0:7d}  end {int(ts[t][7])-t0:7d}")
names = ["start", "fg_full", "epi1 math", "bar1", "dense_full", "epi2 math", "bar2", "end"]
for t in range(15):
    row = ts[t]
    d = [int(row[k + 1] - row[k]) for k in range(7)]
    print(f"tile {t:2d} total {int(row[7]-row[0]):6d}  " + "  ".join(f"{names[k+1]}:{d[k]}" for k in range(7)), " gap_to_next", int(ts[t+1][0]-row[7]) if t < 14 else "")
