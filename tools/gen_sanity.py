"""Small half-precision generation runs through every pipeline geometry (for compute-sanitizer):
compute-sanitizer --tool memcheck python tools/gen_sanity.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_b200.wavenet.model import wavenet
from music_b200.wavenet import fast_generate as fg
for n_layers, bias in ((3, False), (7, True), (12, False), (30, False)):
    dil = [2 ** (i % 10) for i in range(n_layers)]
    net = wavenet(2, dil, 64, 64, 256, 256, bias, mode="bf16").cuda()
    for streams in (5, 19):
        prime = torch.randint(0, 256, (streams, net.receptive_field), device="cuda")
        for bpc, gpc in (("4", "1"), ("2", "1"), ("2", "3")):
            os.environ["WN_GEN_PIPE"], os.environ["WN_GEN_BPC"], os.environ["WN_GEN_GPC"] = "1", bpc, gpc
            codes = fg.generate_codes(net, 6, prime)
            u = torch.rand(6, streams, device="cuda")
            codes2 = fg.generate_codes(net, 6, prime, uniforms=u)
            torch.cuda.synchronize()
            print(n_layers, bias, streams, bpc, gpc, codes[-1, :3].tolist(), codes2[-1, :3].tolist(), flush=True)
