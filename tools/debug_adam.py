import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_b200 import _lib as L
lib = L.init(0)
n = 100000
torch.manual_seed(0)
p0 = torch.randn(n, device="cuda"); g = torch.randn(n, device="cuda") * 1e-3
m0 = torch.randn(n, device="cuda") * 1e-3; v0 = torch.rand(n, device="cuda") * 1e-6
for step in (1, 2, 7):
    pa, ma, va = p0.clone(), m0.clone(), v0.clone()
    pb, mb, vb = p0.clone(), m0.clone(), v0.clone()
    L.check(lib.wn_adam_step(L.ptr(pa), L.ptr(g), L.ptr(ma), L.ptr(va), n, 1e-3, 0.9, 0.999, 1e-8, step, L.stream_ptr()))
    d = torch.tensor([step - 1], dtype=torch.int32, device="cuda")
    L.check(lib.wn_adam_step_dev(L.ptr(pb), L.ptr(g), L.ptr(mb), L.ptr(vb), n, 1e-3, 0.9, 0.999, 1e-8, L.ptr(d), L.stream_ptr()))
    torch.cuda.synchronize()
    print(step, int(d[0]), float((pa - pb).abs().max()), float((ma - mb).abs().max()), float((va - vb).abs().max()), float((pa - p0).abs().max()))
