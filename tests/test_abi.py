"""CPU: the C-ABI library loads, exports every symbol the header declares, and refuses to run
without an sm_100 device (no fallback)."""
import os
import re

import pytest

from music_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "wavenet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    lib = L.load()
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/wavenet_b200.h but not exported"
        assert s in L.SIGNATURES, f"{s} has no ctypes signature in music_b200/_lib.py"
    for s in L.SIGNATURES:
        assert s in syms, f"{s} bound in _lib.py but not declared in the header"


def test_version_and_plan_without_gpu():
    lib = L.load()
    assert lib.wn_version() >= 100
    h = L.make_model([1, 2, 4, 8], 16, 16, 32, 256, False)
    assert lib.wn_model_receptive_field(h) == 17
    n = 16 * 256 * 2 + 4 * (2 * 16 * 16 * 2 + 16 * 16 + 32 * 16) + 32 * 32 + 256 * 32
    assert lib.wn_model_param_count(h) == n
    lib.wn_model_destroy(h)
    hb = L.make_model([1, 2], 8, 8, 16, 256, True)
    nb = (8 * 256 * 2 + 8) + 2 * (2 * (8 * 8 * 2 + 8) + 8 * 8 + 8 + 16 * 8 + 16) + 16 * 16 + 16 + 256 * 16 + 256
    assert lib.wn_model_param_count(hb) == nb
    lib.wn_model_destroy(hb)


def test_bad_config_is_rejected():
    with pytest.raises(L.WavenetB200Error):
        L.make_model([1, 2], 8, 8, 16, 256, False, filter_width=3)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = L.load()
    assert lib.wn_init(0) == L.WN_ERR_UNSUPPORTED
    assert "no CPU fallback" in L.last_error()
    # every compute entry point refuses to run before a successful wn_init
    assert lib.wn_softmax_fwd(None, 1, 256, 1, 0, None, None) == L.WN_ERR_UNSUPPORTED
    from music_b200.wavenet.model import wavenet
    net = wavenet(2, [1, 2], 8, 8, 16, 256, False)
    with pytest.raises(L.WavenetB200Error):
        net(torch.zeros(1, 256, 8))
    from music_b200.wavenet.audio_func import mu_law_encode
    with pytest.raises(L.WavenetB200Error):
        mu_law_encode(torch.zeros(4))


def test_state_dict_keys_match_reference_layout():
    from oracle import wavenet_oracle as O
    from music_b200.wavenet.model import wavenet
    for bias in (False, True):
        net = wavenet(2, [1, 2, 4], 8, 8, 16, 256, bias)
        keys = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
        assert keys == [(k, tuple(s)) for k, s in O.wavenet_param_shapes([1, 2, 4], 8, 8, 16, 256, bias)]
        assert net.receptive_field == O.receptive_field(2, [1, 2, 4])
