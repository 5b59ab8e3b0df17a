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


@pytest.mark.parametrize("n_layers", [1, 2, 3, 4, 5, 7, 8, 11, 12, 15])
def test_pipeline_geometries_agree_with_the_one_cta_kernel(n_layers, monkeypatch):
    """The cluster pipeline in both geometries (2 and 4 blocks per CTA; every role split the layer count produces: CTA 0 alone,
    CTA 0 + last, a short CTA in between with the skip-sum bypass, full CTAs with shared-memory fragments; one group per cluster
    and several) against the one-CTA-per-8-streams kernel on the same weights and primes: same picks, logits within 5e-3 relative
    up to the first step where a pick differs (the kernels accumulate in different orders; both round x and z to fp16 once)."""
    from music_b200.wavenet.fast_generate import generate_codes
    dil = [2 ** (i % 5) for i in range(n_layers)]
    Q, bias = 256, bool(n_layers % 2)
    st = O.init_wavenet_state(dil, 64, 64, 256, Q, bias, seed=20 + n_layers, scale=1.5)
    rf = O.receptive_field(2, dil)
    net = build_net(dil, 64, 64, 256, Q, bias, st, mode="bf16")
    g = torch.Generator().manual_seed(11)
    n = 10
    for streams in (5, 19):      # 1 and 3 groups
        primes = torch.randint(0, Q, (streams, rf), generator=g).cuda()
        monkeypatch.setenv("WN_GEN_PIPE", "0")
        ref_codes, ref_logits = generate_codes(net, n, primes, return_logits=True)
        ref_codes, ref_logits = ref_codes.cpu(), ref_logits.cpu()
        u = torch.rand(n, streams, generator=g)
        ref_sampled = generate_codes(net, n, primes, uniforms=u).cpu()
        for bpc, gpc in (("2", "1"), ("2", "3"), ("4", "1")):
            monkeypatch.setenv("WN_GEN_PIPE", "1")
            monkeypatch.setenv("WN_GEN_BPC", bpc)
            monkeypatch.setenv("WN_GEN_GPC", gpc)
            codes, logits = generate_codes(net, n, primes, return_logits=True)
            codes, logits = codes.cpu(), logits.cpu()
            agree = 0
            for s in range(streams):
                for i in range(n):
                    assert max_rel(logits[i, s].numpy(), ref_logits[i, s].numpy()) < 5e-3, (bpc, gpc, streams, s, i)
                    if int(codes[i, s]) != int(ref_codes[i, s]):
                        break
                    agree += 1
                else:
                    continue
            assert agree >= 0.9 * streams * n, (bpc, gpc, streams, agree)
            sampled = generate_codes(net, n, primes, uniforms=u).cpu()      # inverse-CDF picks with the same uniforms
            same = 0
            for s in range(streams):
                for i in range(n):
                    if int(sampled[i, s]) != int(ref_sampled[i, s]):
                        break
                    same += 1
            assert same >= 0.7 * streams * n, (bpc, gpc, streams, same)      # (a sample next to a CDF step may flip; the stream then differs)
        monkeypatch.delenv("WN_GEN_BPC")
        monkeypatch.delenv("WN_GEN_GPC")
        monkeypatch.delenv("WN_GEN_PIPE")


def _save_reference_checkpoint(tmp_path, dil, R, D, S, Q, bias, st, name="wavenet7.model", dataparallel=False):
    """params JSON + a checkpoint in the reference's format (train.py:45-50: torch.save of the state_dict; files saved from
    nn.DataParallel carry a 'module.' prefix, train.py:61-69)."""
    import json
    from collections import OrderedDict
    params = {"filter_width": 2, "dilations": dil, "dilation_channels": D, "residual_channels": R, "skip_channels": S,
              "quantization_channels": Q, "use_bias": bias}
    pj = tmp_path / "wavenet_params.json"
    pj.write_text(json.dumps(params))
    keys = [k for k, _ in O.wavenet_param_shapes(dil, D, R, S, Q, bias)]
    sd = OrderedDict((("module." + k) if dataparallel else k, st[k].clone()) for k in keys)
    (tmp_path / "restore").mkdir(exist_ok=True)
    torch.save(sd, str(tmp_path / "restore" / name))
    return str(pj), str(tmp_path / "restore") + "/", name


@pytest.mark.parametrize("mode,dataparallel", [("fp32", False), ("auto", True)])
def test_generate_entry_point_end_to_end(tmp_path, mode, dataparallel):
    """generate() (fast_generate.py:144-179) as a user calls it: params JSON + checkpoint -> prime with one-hot(128) x rf ->
    duration * sr steps -> mu-law decode -> wav file.  fp32: the written waveform is the oracle's greedy sequence, decoded,
    bit for bit.  auto (-> the half-precision kernel for this 64/64/256 model): valid codes, a readable wav of the right
    length, and the first samples equal the fp32 sequence as long as no near-tie in the logits is broken differently."""
    from scipy.io import wavfile
    from music_b200.wavenet.fast_generate import generate
    dil = [1, 2, 4, 8, 16, 1, 2, 4, 8, 16]
    R = D = 64
    S = Q = 256
    st = O.init_wavenet_state(dil, D, R, S, Q, False, seed=12, scale=1.5)
    pj, mpath, mname = _save_reference_checkpoint(tmp_path, dil, R, D, S, Q, False, st, dataparallel=dataparallel)
    sr, duration = 40, 2
    out_dir = str(tmp_path / "gen") + "/"
    audio = generate(mpath, mname, out_dir, "a.wav", sr=sr, duration=duration, params_path=pj, mode=mode)
    assert audio.shape == (sr * duration,) and audio.dtype == torch.float32
    ref_codes = O.generate(st, dil, sr * duration)
    ref_audio = O.mu_law_decode(torch.tensor(ref_codes), Q)
    rate, wav = wavfile.read(out_dir + "a.wav")
    assert rate == sr and wav.shape == (sr * duration,)
    assert np.array_equal(wav, audio.numpy())
    if mode == "fp32":
        assert (audio.view(torch.int32) - ref_audio.view(torch.int32)).abs().max() <= 1    # decode: 1 ulp (vector/scalar pow loops)
        from music_b200.wavenet.audio_func import mu_law_encode
        assert mu_law_encode(audio.cuda(), Q).cpu().tolist() == ref_codes
    else:
        assert float(audio.abs().max()) <= 1.0
        assert torch.equal(audio[:4].view(torch.int32), ref_audio[:4].view(torch.int32))
    # a user-supplied start piece (one-hot (1,Q,rf), as the reference's signature takes it)
    rf = O.receptive_field(2, dil)
    g = torch.Generator().manual_seed(2)
    piece_idx = torch.randint(0, Q, (1, rf), generator=g)
    audio2 = generate(mpath, mname, out_dir, "b.wav", start_piece=O.one_hot(piece_idx, Q), sr=sr, duration=1, params_path=pj, mode="fp32")
    ref2 = O.mu_law_decode(torch.tensor(O.generate(st, dil, sr, start_piece=O.one_hot(piece_idx, Q))), Q)
    assert (audio2.view(torch.int32) - ref2.view(torch.int32)).abs().max() <= 1


def test_generate_missing_checkpoint_raises(tmp_path):
    from music_b200.wavenet.fast_generate import generate
    dil = [1, 2]
    st = O.init_wavenet_state(dil, 16, 16, 32, 256, False, seed=1)
    pj, mpath, _ = _save_reference_checkpoint(tmp_path, dil, 16, 16, 32, 256, False, st)
    with pytest.raises(FileNotFoundError):
        generate(mpath, "wavenet999.model", str(tmp_path / "g") + "/", "x.wav", sr=10, duration=1, params_path=pj)


@pytest.mark.parametrize("name", ["wn_tiny_onehot", "wn_bias_dense"])
def test_slow_predict_next_matches_reference_rule(golden, name):
    """model.predict_next (wavenet/model.py:148-165): full forward, greedy pick on the LAST ROW of the (scrambled) softmax
    output.  Checked against the same rule applied to the unmodified reference's golden probabilities, and to the oracle."""
    from music_b200.wavenet.model import predict_next
    from tests.util import cfg_state, make_net
    z = golden(name)
    dil, st = cfg_state(z)
    net = make_net(z, st, mode="fp32")
    Q, L = int(z["Q"]), int(z["L"])
    if "x" in z.files:
        x = torch.from_numpy(z["x"]).float()
    else:
        x = O.one_hot(torch.from_numpy(z["idx"].astype(np.int64))[:, :L], Q)
    with torch.no_grad():
        got = predict_next(net, x.cuda(), Q)
    probs = O.forward_probs(st, dil, x)
    want = int(torch.topk(probs.view(-1, Q)[-1], 1)[1])
    assert got.shape == (1,) and got.dtype == torch.int64
    assert int(got[0]) == want
    if "probs" in z.files:                                   # the reference's own output, frozen by oracle/make_golden.py
        assert int(np.argmax(z["probs"].reshape(-1, Q)[-1])) == want
