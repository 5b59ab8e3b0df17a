"""bf16 forward+backward on a small model against the oracle, per-parameter errors (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import wavenet_oracle as O
from music_b200.wavenet.model import wavenet
from music_b200.wavenet.train import Trainer

def run(dil, B, W, scale, seed=5):
    st = O.init_wavenet_state(dil, 64, 64, 256, 256, False, seed=seed, scale=scale)
    rf = O.receptive_field(2, dil); L = rf + W - 1
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, 256, (B, L + 1), generator=g)
    tgt = idx[:, rf:rf + W].contiguous()
    loss_ref, g_ref = O.grads(st, dil, O.one_hot(idx[:, :L], 256), tgt)
    out = {}
    for mode in ("fp32", "bf16"):
        net = wavenet(2, dil, 64, 64, 256, 256, False, mode=mode)
        net.load_state_dict(st); net = net.cuda()
        tr = Trainer(net, "adam", distributed=False)
        loss = float(tr.forward_backward(idx[:, :L].cuda(), tgt.cuda()))
        torch.cuda.synchronize()
        gv = net.engine.grad_views(net._params())
        worst = 0
        for (k, p), gg in zip(net.named_parameters(), gv):
            r = g_ref[k].numpy(); a = gg.cpu().numpy()
            if np.abs(r).max() == 0:
                e = float(np.abs(a).max())
            else:
                e = float(np.linalg.norm(a - r) / np.linalg.norm(r))
            worst = max(worst, e)
            if mode == "bf16" and (e > 2e-2 or "--all" in sys.argv):
                print(f"   {k:40s} rel-l2 {e:.4f}  |ref| {np.linalg.norm(r):.3e}")
        print(f"dil={len(dil)} B={B} W={W} scale={scale} mode={mode}: loss {loss:.6f} (ref {loss_ref:.6f}) worst grad rel-l2 {worst:.4f}", flush=True)

run([1, 2, 4], 1, 200, 1.5)
run([1, 2, 4, 8, 16, 32], 2, 300, 1.5)
run([2 ** i for i in range(10)] * 3, 2, 300, 1.0)
