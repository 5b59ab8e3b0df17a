"""GPU parity: mu-law codec, bit-exact against the golden vectors of the reference."""
import numpy as np
import pytest
import torch

from oracle import wavenet_oracle as O

pytestmark = pytest.mark.gpu


def test_encode_decode_bit_exact_vs_golden(golden):
    from music_b200.wavenet.audio_func import mu_law_decode, mu_law_encode
    z = golden("mulaw")
    enc = mu_law_encode(torch.from_numpy(z["x"]).cuda())
    assert enc.dtype == torch.int64
    assert np.array_equal(enc.cpu().numpy(), z["enc"].astype(np.int64))
    dec = mu_law_decode(torch.arange(256).cuda())
    assert np.array_equal(dec.cpu().numpy(), z["dec"])


def test_encode_bit_exact_vs_oracle_at_bin_edges_and_random():
    from music_b200.wavenet.audio_func import host_tables, mu_law_decode, mu_law_encode
    thr, _ = host_tables(256)
    edges = thr[1:].view(np.int32).astype(np.int64)
    near = (edges[:, None] + np.arange(-48, 48)[None, :]).reshape(-1).astype(np.int32).view(np.float32)
    g = torch.Generator().manual_seed(0)
    x = torch.cat([torch.from_numpy(near.copy()), torch.rand(1 << 20, generator=g) * 2.2 - 1.1,
                   torch.randn(1 << 18, generator=g) * 0.05])
    x = x[: (x.numel() // 16) * 16]
    got = mu_law_encode(x.cuda()).cpu()
    ref = O.mu_law_encode(x)
    assert torch.equal(got, ref)
    # round trip property at full size: decode(encode(x)) re-encodes to the same code
    codes = mu_law_encode(mu_law_decode(got.cuda()))
    assert torch.equal(codes.cpu(), got)
    assert mu_law_encode(torch.zeros(0).cuda()).numel() == 0


def test_other_q():
    from music_b200.wavenet.audio_func import mu_law_decode, mu_law_encode
    x = torch.linspace(-1, 1, 4096)
    for q in (16, 64):
        assert torch.equal(mu_law_encode(x.cuda(), q).cpu(), O.mu_law_encode(x, q))
        assert torch.equal(mu_law_decode(torch.arange(q).cuda(), q).cpu(), O.mu_law_decode(torch.arange(q), q))


def test_device_one_hot_encode_matches_reference_loader(golden):
    """one_hot_encode (wavenet/faster_audio_data.py:62-83) on the device, against the UNMODIFIED reference's own output
    (tests/golden/loader.npz) and against the host mirror on a full-size batch: the reshape quirk bit for bit, and the true
    one-hot; only the integer codes cross PCIe."""
    from music_b200.wavenet.faster_audio_data import one_hot_encode, one_hot_encode_device
    z = golden("loader")
    codes = torch.from_numpy(z["onehot_in"].astype(np.int64))
    got = one_hot_encode_device(codes.cuda(), 256).cpu().numpy()
    assert got.shape == z["onehot_out"].shape and np.array_equal(got, z["onehot_out"])
    g = torch.Generator().manual_seed(3)
    batch = torch.randint(0, 256, (3, 19070), generator=g)
    dev = one_hot_encode_device(batch.cuda(), 256)
    dev_t = one_hot_encode_device(batch.cuda(), 256, transpose=True)
    for b in range(3):
        ref = one_hot_encode({"audio_piece": batch[b], "audio_target": torch.zeros(1)})["audio_piece"]
        assert torch.equal(dev[b].cpu(), ref)
    assert torch.equal(dev_t.argmax(dim=1).cpu(), batch) and float(dev_t.sum()) == batch.numel()
    with pytest.raises(Exception):
        one_hot_encode_device(batch, 256)          # host tensor: refused, no CPU fallback
