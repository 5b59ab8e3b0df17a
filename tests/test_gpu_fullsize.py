"""GPU property tests at the FULL sizes of BASELINE.json (the oracle cannot run these in seconds, so they check
size-independent properties instead of element-wise parity): cfg 2 training shape (30 layers, 64/64/256, 16 k-sample
windows), cfg 4 generation shape (64 streams), 1.6 M-sample mu-law round trips."""
import math

import numpy as np
import pytest
import torch

from oracle import wavenet_oracle as O
from tests.util import build_net, rel_err

pytestmark = pytest.mark.gpu

DIL = [2 ** i for i in range(10)] * 3
W_FULL = 16000


def _net(mode="bf16", seed=3):
    st = O.init_wavenet_state(DIL, 64, 64, 256, 256, False, seed=seed, scale=1.0)
    return build_net(DIL, 64, 64, 256, 256, False, st, mode=mode)


def _batch(B, seed):
    rf = O.receptive_field(2, DIL)
    L = rf + W_FULL - 1
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, 256, (B, L + 1), generator=g)
    return idx[:, :L].contiguous().cuda(), idx[:, rf:rf + W_FULL].contiguous().cuda()


def test_full_size_step_loss_bounds_and_batch_linearity():
    """One bf16 forward/backward at the cfg-2 shape (B = 4 clips of L = 19070):
    * the loss is a cross entropy over the module's SOFTMAX output (train.py:146,178): ln(255 + e) - 1 <= loss <= ln 256;
    * the gradient of the mean loss is linear in the batch: grad(4 clips) = mean(grad(first 2), grad(last 2)) - the property
      data-parallel training rests on (train.py:117-122) - up to bf16 accumulation-order noise."""
    from music_b200.wavenet.train import Trainer
    net = _net()
    tr = Trainer(net, "adam", distributed=False)
    x, y = _batch(4, 11)
    loss = float(tr.forward_backward(x, y))
    assert math.log(255 + math.e) - 1 - 1e-4 <= loss <= math.log(256) + 1e-4, loss
    g_full = net.engine.gflat.clone()
    assert torch.isfinite(g_full).all()
    halves = []
    for sl in (slice(0, 2), slice(2, 4)):
        tr.forward_backward(x[sl].contiguous(), y[sl].contiguous())
        halves.append(net.engine.gflat.clone())
    g_mean = 0.5 * (halves[0] + halves[1])
    assert rel_err(g_full.cpu().numpy(), g_mean.cpu().numpy()) < 2e-3
    # and the same step twice gives the same loss and (up to the order of fp32 atomics in the weight-gradient flushes) gradient
    loss2 = float(tr.forward_backward(x, y))
    assert loss2 == loss
    assert rel_err(net.engine.gflat.cpu().numpy(), g_full.cpu().numpy()) < 1e-5


def test_full_size_forward_rows_are_distributions():
    """Reference forward() at the full window: (B*W, 256) rows of the scrambled softmax are probability vectors."""
    net = _net()
    x, _ = _batch(2, 12)
    with torch.no_grad():
        probs = net.forward_indices(x)
    assert probs.shape == (2 * W_FULL, 256)
    s = probs.sum(dim=1)
    assert float((s - 1).abs().max()) < 1e-4 and float(probs.min()) >= 0.0


def test_generation_64_streams_are_deterministic_and_independent():
    """cfg 4 shape: 64 streams primed identically produce identical greedy sequences (no cross-stream leakage in the
    8-streams-per-CTA kernel), the same launch repeated gives the same codes, and splitting the steps over two
    launches (state carried in the ring buffers) gives the same sequence as one launch."""
    from music_b200.wavenet import fast_generate as FG
    net = _net()
    rf = net.receptive_field
    prime = torch.full((64, rf), 128, dtype=torch.int64, device="cuda")
    prime[32:, -5:] = torch.tensor([3, 200, 77, 12, 250], device="cuda")          # two different groups of 32 streams
    with torch.no_grad():
        first, st, _ = FG._prime(net, prime)
        st2 = st.clone()
        a, _ = FG._steps(net, st, first, 600)
        b1, _ = FG._steps(net, st2, first, 250)
        b2, _ = FG._steps(net, st2, b1[-1].contiguous(), 350)
    assert int(a.min()) >= 0 and int(a.max()) <= 255
    assert torch.equal(a[:, :32], a[:, :1].expand(-1, 32)) and torch.equal(a[:, 32:], a[:, 32:33].expand(-1, 32))
    assert not torch.equal(a[:, 0], a[:, 32])
    assert torch.equal(torch.cat([b1, b2]), a)


def test_mulaw_round_trips_at_full_size():
    """1.6 M samples: encode(decode(q)) == q for every code (idempotence), decode(encode(x)) stays within one
    quantisation bin of x, and encode is monotone (audio_func.py:5-39)."""
    from music_b200.wavenet.audio_func import mu_law_decode, mu_law_encode
    q = torch.arange(256, device="cuda").repeat(6250)
    assert torch.equal(mu_law_encode(mu_law_decode(q)), q)
    g = torch.Generator().manual_seed(4)
    x = (torch.rand(1_600_000, generator=g) * 2 - 1).cuda()
    e = mu_law_encode(x)
    xs, order = torch.sort(x)
    assert bool((e[order][1:] >= e[order][:-1]).all())
    back = mu_law_decode(e)
    width = mu_law_decode(torch.clamp(e + 1, max=255)) - mu_law_decode(torch.clamp(e - 1, min=0))
    assert bool(((back - x).abs() <= width + 1e-6).all())


@pytest.mark.parametrize("n", [1, 3, 9])
def test_generation_ragged_stream_counts_match_the_full_group(n):
    """The half-precision kernel advances 8 streams per CTA; stream counts that do not fill a group (1, 3) or spill into a second
    one (9) must give exactly the sequences the same streams produce inside a full launch of 16."""
    from music_b200.wavenet import fast_generate as FG
    net = _net()
    rf = net.receptive_field
    g = torch.Generator().manual_seed(21)
    prime = torch.randint(0, 256, (16, rf), generator=g).cuda()
    with torch.no_grad():
        first, st, _ = FG._prime(net, prime)
        full, _ = FG._steps(net, st, first, 200)
        first_n, st_n, _ = FG._prime(net, prime[:n].contiguous())
        part, _ = FG._steps(net, st_n, first_n, 200)
    assert torch.equal(first_n, first[:n])
    assert torch.equal(part, full[:, :n])
