"""GPU parity, bf16 tensor-core mode (tcgen05 / TMEM / TMA kernels).
Bar (north star): pre-softmax logits within 1e-2 relative error of the fp32 reference."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import wavenet_oracle as O
from tests.util import build_net, cfg_state, make_net, max_rel, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-2


def test_umma_tma_building_blocks_exact():
    """D = A B^T tiles through TMA -> UMMA -> TMEM for every operand layout the kernels use; inputs are
    small multiples of 1/4 so the expected error is exactly 0."""
    from music_b200 import _lib as L
    lib = L.init(torch.cuda.current_device())
    n = 7
    err = (C.c_float * n)()
    L.check(lib.wn_selftest_umma(err, n, L.stream_ptr()))
    errs = list(err)
    print("selftest max errors:", errs)
    assert errs == [0.0] * n, errs


def _c64_inputs(z):
    idx = torch.from_numpy(z["idx"].astype(np.int64))[:, :int(z["L"])]
    tgt = torch.from_numpy(z["target"].astype(np.int64))
    return idx, tgt


def test_forward_logits_vs_golden_c64(golden):
    z = golden("wn_c64_onehot")
    dil, st = cfg_state(z)
    idx, tgt = _c64_inputs(z)
    net = make_net(z, st, mode="bf16")
    sub = int(z["rows"][1] - z["rows"][0])
    lg = net.forward_logits(indices=idx.cuda()).detach().cpu().numpy()
    assert np.isfinite(lg).all()
    e = max_rel(lg[:, :, ::sub], z["logits_cols"])
    print("bf16 logits max-rel err vs golden:", e)
    assert e < TOL
    probs = net.forward_indices(idx.cuda()).detach().cpu().numpy()
    assert max_rel(probs[z["rows"]], z["probs_rows"]) < TOL


@pytest.mark.parametrize("dense", [False, True])
def test_forward_30_layers_vs_oracle(dense):
    """The cfg-2 model (10 x 3 dilations up to 512, 64/64/256) on a short window, both input paths."""
    dil = [2 ** i for i in range(10)] * 3
    st = O.init_wavenet_state(dil, 64, 64, 256, 256, False, seed=3, scale=1.5)
    rf = O.receptive_field(2, dil)
    B, W = 2, 333
    L = rf + W - 1
    g = torch.Generator().manual_seed(8)
    idx = torch.randint(0, 256, (B, L), generator=g)
    x = O.one_hot(idx, 256)
    ref = O.forward_logits(st, dil, x).numpy()
    net = build_net(dil, 64, 64, 256, 256, False, st, mode="bf16")
    lg = (net.forward_logits(wave_sample=x.cuda()) if dense else net.forward_logits(indices=idx.cuda())).detach().cpu().numpy()
    e = max_rel(lg, ref)
    print("30-layer bf16 logits max-rel err:", e, "rel-l2:", rel_err(lg, ref))
    assert e < TOL
    # fp32 check mode on the same input: 1e-4
    net32 = build_net(dil, 64, 64, 256, 256, False, st, mode="fp32")
    lg32 = net32.forward_logits(indices=idx.cuda()).detach().cpu().numpy()
    assert max_rel(lg32, ref) < 1e-4


def test_forward_with_bias_vs_oracle():
    dil = [1, 2, 4, 8, 16, 32, 64, 128]
    st = O.init_wavenet_state(dil, 64, 64, 256, 256, True, seed=4, scale=1.5)
    rf = O.receptive_field(2, dil)
    B, W = 3, 200
    g = torch.Generator().manual_seed(9)
    idx = torch.randint(0, 256, (B, rf + W - 1), generator=g)
    ref = O.forward_logits(st, dil, O.one_hot(idx, 256)).numpy()
    net = build_net(dil, 64, 64, 256, 256, True, st, mode="bf16")
    lg = net.forward_logits(indices=idx.cuda()).detach().cpu().numpy()
    e = max_rel(lg, ref)
    print("bias bf16 logits max-rel err:", e)
    assert e < TOL


def test_unsupported_shape_fails_loudly():
    from music_b200 import _lib as L
    st = O.init_wavenet_state([1, 2], 16, 16, 32, 256, False)
    net = build_net([1, 2], 16, 16, 32, 256, False, st, mode="bf16")
    with pytest.raises(L.WavenetB200Error, match="specialised"):
        net.forward_logits(indices=torch.zeros(1, 8, dtype=torch.int64).cuda())
