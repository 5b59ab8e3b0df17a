"""GPU parity: incremental generation (fast_generate.predict_next) against the golden sequences of
the reference's own predict_next and against the oracle."""
import numpy as np
import pytest
import torch

from oracle import wavenet_oracle as O
from tests.util import build_net, max_rel, state_of

pytestmark = pytest.mark.gpu


def _net(z, mode="fp32"):
    dil = [int(d) for d in z["dilations"]]
    st = state_of(z)
    return dil, st, build_net(dil, int(z["R"]), int(z["D"]), int(z["S"]), int(z["Q"]), bool(z["use_bias"]), st, mode)


@pytest.mark.parametrize("name", ["gen_tiny", "gen_bias"])
def test_sequence_and_logits_match_reference_predict_next(golden, name):
    from music_b200.wavenet.fast_generate import generate_codes
    z = golden(name)
    dil, st, net = _net(z)
    prime = torch.from_numpy(z["prime_idx"].astype(np.int64)).cuda()
    n = len(z["picks"])
    codes, logits = generate_codes(net, n, prime, return_logits=True)
    assert max_rel(logits[:, 0].cpu().numpy(), z["logits"]) < 1e-4
    assert codes[:, 0].cpu().tolist() == [int(p) for p in z["picks"]]       # queue indexing: bit-exact picks


def test_drop_in_predict_next_loop_and_state_layout(golden):
    """The reference's own calling pattern (generate(), fast_generate.py:166-172), one call per
    sample, and the exported state_queue against the reference's final queues."""
    from music_b200.wavenet.fast_generate import predict_next
    z = golden("gen_tiny")
    dil, st, net = _net(z)
    Q = int(z["Q"])
    note = O.one_hot(torch.from_numpy(z["prime_idx"].astype(np.int64)), Q).cuda()
    queue, picks = None, []
    for i in range(len(z["picks"])):
        p, queue = predict_next(net, note, queue)
        k = int(p[0])
        picks.append(k)
        note = torch.zeros(1, Q, 1, device="cuda")
        note[:, k, :] = 1.0
    assert picks == [int(p) for p in z["picks"]]
    keys = list(queue.keys())
    assert keys == ["causal_layer"] + [f"block_{i + 1}" for i in range(len(dil))]
    for k in keys:
        ref = z["queue." + k]
        got = queue[k].cpu().numpy()
        assert got.shape == ref.shape, k
        assert np.abs(got - ref).max() < 1e-4 * max(1.0, np.abs(ref).max()), k


def test_input_push_equals_full_forward(golden):
    from music_b200.wavenet.fast_generate import generate_codes
    z = golden("gen_tiny")
    dil, st, net = _net(z)
    prime = torch.from_numpy(z["prime_idx"].astype(np.int64))
    n = 24
    codes, logits = generate_codes(net, n, prime.cuda(), queue_push="input", return_logits=True)
    ref_codes, ref_logits = O.generate(st, dil, n, start_piece=O.one_hot(prime, int(z["Q"])), queue_push="input",
                                       return_logits=True)
    assert max_rel(logits[:, 0].cpu().numpy(), ref_logits.numpy()) < 1e-4
    assert codes[:, 0].cpu().tolist() == ref_codes


def test_streams_are_independent_and_sampling_follows_uniforms(golden):
    from music_b200.wavenet.fast_generate import generate_codes
    z = golden("gen_bias")
    dil, st, net = _net(z)
    Q, rf = int(z["Q"]), O.receptive_field(2, dil)
    g = torch.Generator().manual_seed(4)
    primes = torch.randint(0, Q, (5, rf), generator=g)
    n = 20
    u = torch.rand(n, 5, generator=g)
    codes = generate_codes(net, n, primes.cuda(), uniforms=u).cpu()
    for s in range(5):
        ref = O.generate(st, dil, n, start_piece=O.one_hot(primes[s:s + 1], Q), uniforms=u[:, s].tolist())
        assert codes[:, s].tolist() == ref, s
    greedy = generate_codes(net, n, primes.cuda()).cpu()
    one = generate_codes(net, n, primes[2:3].cuda()).cpu()
    assert greedy[:, 2].tolist() == one[:, 0].tolist()


def test_import_reference_layout_state(golden):
    """A queue dict in the reference's layout can be imported and continued."""
    from music_b200.wavenet.fast_generate import generate_codes, import_state, predict_next
    z = golden("gen_tiny")
    dil, st, net = _net(z)
    Q = int(z["Q"])
    prime = torch.from_numpy(z["prime_idx"].astype(np.int64))
    # oracle: run 10 steps, hand its queues over, continue 10 more on the GPU
    note, queues, picks = O.one_hot(prime, Q), None, []
    for i in range(10):
        lg, queues = (O.gen_prime(st, dil, note) if queues is None else O.gen_step(st, dil, note, queues))
        k = O.pick_greedy(lg)
        picks.append(k)
        note = torch.zeros(1, Q, 1)
        note[:, k, :] = 1.0
    gstate = import_state(net, queues)
    cont = []
    for i in range(10):
        p, gstate = predict_next(net, note.cuda(), gstate)
        cont.append(int(p[0]))
        note = torch.zeros(1, Q, 1)
        note[:, cont[-1], :] = 1.0
    assert picks + cont == [int(p) for p in z["picks"][:20]]


def test_bf16_generation_kernel_teacher_forced_vs_oracle():
    """bf16-weight generation kernel (two streams per CTA): per-step logits against the fp32 oracle driven
    by the GPU's own picks (teacher forcing, so one flipped argmax cannot hide later agreement).
    Tolerance 1e-2 relative (bf16 weights, fp32 state)."""
    from music_b200.wavenet.fast_generate import generate_codes
    dil = [1, 2, 4, 8, 16, 1, 2, 4, 8, 16]
    Q = 256
    for bias in (False, True):
        st = O.init_wavenet_state(dil, 64, 64, 256, Q, bias, seed=9, scale=1.5)
        rf = O.receptive_field(2, dil)
        net = build_net(dil, 64, 64, 256, Q, bias, st, mode="bf16")
        g = torch.Generator().manual_seed(10)
        primes = torch.randint(0, Q, (3, rf), generator=g)           # odd stream count: one CTA has a single stream
        n = 24
        codes, logits = generate_codes(net, n, primes.cuda(), return_logits=True)
        codes, logits = codes.cpu(), logits.cpu()
        agree = 0
        for s in range(3):
            note, queues = O.one_hot(primes[s:s + 1], Q), None
            for i in range(n):
                lg, queues = (O.gen_prime(st, dil, note) if queues is None else O.gen_step(st, dil, note, queues))
                assert max_rel(logits[i, s].numpy(), lg.numpy()) < 1e-2, (bias, s, i)
                agree += int(O.pick_greedy(lg) == int(codes[i, s]))
                note = torch.zeros(1, Q, 1)
                note[:, int(codes[i, s]), :] = 1.0
        assert agree >= 0.9 * 3 * n, agree
