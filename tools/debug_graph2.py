import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_b200.wavenet.model import wavenet
from music_b200.wavenet import train as T
dil = [1, 2, 4, 8, 16, 32]
torch.manual_seed(4)
net_a = wavenet(2, dil, 64, 64, 256, 256, False, mode="bf16", parity="corrected").cuda()
net_b = copy.deepcopy(net_a)
rf = net_a.receptive_field
x1 = torch.randint(0, 256, (2, rf + 300)).cuda()
tr_a = T.Trainer(net_a, "adam", 1e-3, distributed=False)
tr_b = T.Trainer(net_b, "adam", 1e-3, distributed=False)
p, t = x1[:, :-1].contiguous(), x1[:, rf:].contiguous()
for tr in (tr_a, tr_b):
    tr.step(p, t)
ok = tr_b.capture(p, t)
g = tr_b.__dict__["_graph"]
print("state keys", list(tr_b.state), "ptr m", tr_b.state["m"].data_ptr(), tr_a.state["m"].data_ptr())
def d(name, a, b): print(name, float((a - b).abs().max()), float(a.abs().max()), float(b.abs().max()))
d("flat before", net_a.engine.flat, net_b.engine.flat); d("m before", tr_a.state["m"], tr_b.state["m"]); d("v before", tr_a.state["v"], tr_b.state["v"])
tr_a.step(p, t); tr_b.step(p, t); torch.cuda.synchronize()
d("flat after", net_a.engine.flat, net_b.engine.flat); d("m after", tr_a.state["m"], tr_b.state["m"]); d("v after", tr_a.state["v"], tr_b.state["v"])
d("g after", net_a.engine.gflat, net_b.engine.gflat)
# what update would adam give from b's own m, v, g?
