"""GPU parity: wavenet_autoencoder forward (model1.py) against golden vectors captured from the unmodified
reference (including its throw-away conditioning convs) and against the oracle. fp32 check mode: 1e-4."""
import numpy as np
import pytest
import torch

from oracle import wavenet_oracle as O
from tests.util import max_rel, state_of

pytestmark = pytest.mark.gpu


def _build(z):
    from music_b200.wavenet_autoencoder.model1 import wavenet_autoencoder
    dil = [int(d) for d in z["dilations"]]
    net = wavenet_autoencoder(2, int(z["Q"]), dil, int(z["Re"]), int(z["De"]), int(z["BW"]), int(z["pool"]), int(z["Rd"]),
                              int(z["Dd"]), int(z["Sd"]), bool(z["use_bias"]))
    st = state_of(z)
    assert list(net.state_dict().keys()) == list(st.keys())
    net.load_state_dict(st)
    cond = {k[len("cond."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("cond.")}
    return dil, st, cond, net.cuda()


@pytest.mark.parametrize("name", ["ae_tile", "ae_bias"])
def test_forward_vs_reference_golden(golden, name):
    z = golden(name)
    dil, st, cond, net = _build(z)
    idx = torch.from_numpy(z["idx"].astype(np.int64))
    with torch.no_grad():
        lg = net.forward_logits(indices=idx.cuda(), cond_weights=cond)
        probs = net(O.one_hot(idx, int(z["Q"])).cuda(), cond_weights=cond)
    assert max_rel(lg.cpu().numpy(), z["logits"]) < 1e-4
    assert max_rel(probs.cpu().numpy(), z["probs"]) < 1e-4
    assert probs.shape == z["probs"].shape


def test_broadcast_branch_and_encoding_vs_oracle():
    """W divisible by the frame count -> the view/broadcast branch of `_conditon`; also checks `_encode`."""
    from music_b200.wavenet_autoencoder.model1 import wavenet_autoencoder
    dil = [1, 2, 4, 8, 1, 2, 4, 8]
    cfg = dict(Q=256, Re=16, De=24, BW=32, pool=8, Rd=16, Dd=16, Sd=48)
    net = wavenet_autoencoder(2, 256, dil, cfg["Re"], cfg["De"], cfg["BW"], cfg["pool"], cfg["Rd"], cfg["Dd"], cfg["Sd"], True)
    st = {k: v.detach().clone() for k, v in net.state_dict().items()}
    cond = {}
    for i, c in enumerate(net.cond_layers):
        cond[f"cond.{i}.weight"] = c.weight.detach().clone()
        cond[f"cond.{i}.bias"] = c.bias.detach().clone()
    rf = O.receptive_field(2, dil)
    B, W = 2, 96                      # 12 frames, 96 % 12 == 0
    x = torch.randn(B, 256, rf + W - 1)
    ref = O.ae_forward_logits(st, cond, dil, x, cfg["pool"])
    enc_ref = O.ae_encode(st, dil, x, cfg["pool"])
    net = net.cuda()
    with torch.no_grad():
        lg, enc = net.forward_logits(wave_sample=x.cuda(), return_encoding=True)
    assert max_rel(enc.cpu().numpy(), enc_ref.numpy()) < 1e-4
    assert max_rel(lg.cpu().numpy(), ref.numpy()) < 1e-4
    with pytest.raises(ValueError, match="wave sample not long enough"):
        net(torch.zeros(1, 256, rf - 1).cuda())
