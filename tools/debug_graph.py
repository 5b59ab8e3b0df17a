import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_b200.wavenet.model import wavenet
from music_b200.wavenet import train as T
dil = [1, 2, 4, 8, 16, 32]
for kind in ("sgd", "adam"):
    torch.manual_seed(4)
    net_a = wavenet(2, dil, 64, 64, 256, 256, False, mode="bf16", parity="corrected").cuda()
    net_b = copy.deepcopy(net_a)
    rf = net_a.receptive_field
    x1 = torch.randint(0, 256, (2, rf + 300)).cuda()
    x2 = torch.randint(0, 256, (2, rf + 300)).cuda()
    tr_a = T.Trainer(net_a, kind, 1e-3, distributed=False)
    tr_b = T.Trainer(net_b, kind, 1e-3, distributed=False)
    for tr in (tr_a, tr_b):
        tr.step(x1[:, :-1].contiguous(), x1[:, rf:].contiguous())
    print(kind, "after 1 eager step, param diff", max(float((a - b).abs().max()) for a, b in zip(net_a.parameters(), net_b.parameters())))
    ok = tr_b.capture(x1[:, :-1].contiguous(), x1[:, rf:].contiguous())
    print(kind, "capture", ok, "param diff after capture", max(float((a - b).abs().max()) for a, b in zip(net_a.parameters(), net_b.parameters())),
          "m diff", {k: float((tr_a.state[k] - tr_b.state[k]).abs().max()) for k in tr_a.state if k in tr_b.state})
    for i, x in enumerate((x2, x1, x2)):
        la = tr_a.step(x[:, :-1].contiguous(), x[:, rf:].contiguous())
        lb = tr_b.step(x[:, :-1].contiguous(), x[:, rf:].contiguous())
        print(kind, i, float(la), float(lb), "param diff", max(float((a - b).abs().max()) for a, b in zip(net_a.parameters(), net_b.parameters())),
              "grad diff", float((net_a.engine.gflat - net_b.engine.gflat).abs().max()), float(net_a.engine.gflat.abs().max()))
