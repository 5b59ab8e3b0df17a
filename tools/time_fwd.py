"""Quick device timing of the forward paths at the cfg-2 shape (not the bench: exploratory)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import wavenet_oracle as O
from music_b200.wavenet.model import wavenet

def timed(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
    return min(ts), sum(ts) / n

dil = [2 ** i for i in range(10)] * 3
B, W = int(os.environ.get("B", 16)), 16000
rf = O.receptive_field(2, dil); L = rf + W - 1
g = torch.Generator().manual_seed(1)
idx = torch.randint(0, 256, (B, L), generator=g).cuda()
for mode in sys.argv[1:] or ["bf16"]:
    net = wavenet(2, dil, 64, 64, 256, 256, False, mode=mode).cuda()
    with torch.no_grad():
        f = lambda: net.forward_logits(indices=idx)
        best, avg = timed(f)
    print(f"mode={mode} B={B} L={L} forward_logits: best {best:.3f} ms avg {avg:.3f} ms -> {B*W/best*1e3:.3e} samples/s fwd-only", flush=True)
    fl = 2 * 21.55e9 * B
    print(f"   fwd TFLOP/s = {fl/best/1e9:.1f}", flush=True)
