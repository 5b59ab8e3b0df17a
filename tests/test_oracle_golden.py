"""The oracle (oracle/wavenet_oracle.py) against the golden vectors frozen from the
unmodified reference by oracle/make_golden.py.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import wavenet_oracle as O


def _state(z, prefix="state."):
    return {k[len(prefix):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix)}


def _cfg_state(z):
    dil = [int(d) for d in z["dilations"]]
    if any(k.startswith("state.") for k in z.files):
        st = _state(z)
    else:
        st = O.init_wavenet_state(dil, int(z["D"]), int(z["R"]), int(z["S"]), int(z["Q"]), bool(z["use_bias"]),
                                  seed=int(z["seed"]), scale=float(z["scale"]))
    return dil, st


def _input(z):
    if "x" in z.files:
        return torch.from_numpy(z["x"])
    idx = torch.from_numpy(z["idx"].astype(np.int64))
    return O.one_hot(idx[:, :int(z["L"])], int(z["Q"]))


@pytest.mark.parametrize("name", ["wn_tiny_onehot", "wn_tiny_dense", "wn_bias_dense"])
def test_forward_loss_grads_adam(golden, name):
    z = golden(name)
    dil, st = _cfg_state(z)
    x = _input(z)
    tgt = torch.from_numpy(z["target"].astype(np.int64))
    logits = O.forward_logits(st, dil, x)
    np.testing.assert_allclose(logits.numpy(), z["logits"], rtol=1e-5, atol=1e-6)
    probs = O.scrambled_softmax(logits)
    np.testing.assert_allclose(probs.numpy(), z["probs"], rtol=1e-5, atol=1e-8)
    loss, gr = O.grads(st, dil, x, tgt)
    assert abs(loss - float(z["loss"])) < 1e-6
    for k, g in gr.items():
        np.testing.assert_allclose(g.numpy(), z["grad." + k], rtol=1e-4, atol=1e-9, err_msg=k)
    ts = O.TrainState(st, "adam", lr=float(z["lr"]))
    l2 = O.train_step(ts, dil, x, tgt)
    assert abs(l2 - float(z["loss"])) < 1e-6
    for k, p in ts.params.items():
        np.testing.assert_allclose(p.detach().numpy(), z["after." + k], rtol=1e-5, atol=1e-7, err_msg=k)


@pytest.mark.parametrize("name", ["wn_c64_onehot", "wn_cfg1"])
def test_forward_subsampled(golden, name):
    z = golden(name)
    dil, st = _cfg_state(z)
    x = _input(z)
    tgt = torch.from_numpy(z["target"].astype(np.int64))
    sub = int(z["rows"][1] - z["rows"][0])
    logits = O.forward_logits(st, dil, x)
    np.testing.assert_allclose(logits.numpy()[:, :, ::sub], z["logits_cols"], rtol=1e-4, atol=1e-6)
    probs = O.scrambled_softmax(logits)
    np.testing.assert_allclose(probs.numpy()[z["rows"]], z["probs_rows"], rtol=1e-5, atol=1e-8)
    loss, gr = O.grads(st, dil, x, tgt)
    assert abs(loss - float(z["loss"])) < 1e-6
    for k, g in gr.items():
        assert abs(float(g.double().norm()) - float(z["gradnorm." + k])) <= 1e-4 * float(z["gradnorm." + k]) + 1e-12, k


def test_scramble_is_not_per_timestep(golden):
    """Fact 4: rows are flat 256-chunks of the (B,Q,W) buffer."""
    z = golden("wn_tiny_onehot")
    lg = torch.from_numpy(z["logits"])
    per_t = torch.softmax(lg, dim=1).permute(0, 2, 1).reshape(-1, lg.shape[1])
    assert not np.allclose(per_t.numpy(), z["probs"], atol=1e-6)
    flat = torch.softmax(lg.reshape(-1, 256), dim=1)
    np.testing.assert_allclose(flat.numpy(), z["probs"], rtol=1e-6, atol=1e-9)


def test_loss_bounds(golden):
    """Fact 3: CE over probabilities is confined to [ln(e+255)-1, ln(e+255)]; ln 256 when uniform."""
    z = golden("wn_tiny_dense")
    assert np.log(np.e + 255) - 1 <= float(z["loss"]) <= np.log(np.e + 255)


@pytest.mark.parametrize("name", ["gen_tiny", "gen_bias"])
def test_generate_matches_reference_predict_next(golden, name):
    z = golden(name)
    dil = [int(d) for d in z["dilations"]]
    st = _state(z)
    prime = O.one_hot(torch.from_numpy(z["prime_idx"].astype(np.int64)), int(z["Q"]))
    picks, logits = O.generate(st, dil, len(z["picks"]), start_piece=prime, queue_push="output", return_logits=True)
    assert picks == [int(p) for p in z["picks"]]
    np.testing.assert_allclose(logits.numpy(), z["logits"], rtol=1e-5, atol=1e-6)


def test_generate_queue_bug_is_observable(golden):
    """Fact 5: the reference pushes block outputs; the consistent variant diverges from it
    and equals the full forward."""
    z = golden("gen_tiny")
    dil = [int(d) for d in z["dilations"]]
    st = _state(z)
    Q = int(z["Q"])
    prime_idx = torch.from_numpy(z["prime_idx"].astype(np.int64))
    n = 12
    seq_in, lg_in = O.generate(st, dil, n, start_piece=O.one_hot(prime_idx, Q), queue_push="input", return_logits=True)
    # full-forward greedy continuation
    hist = prime_idx.clone()
    for i in range(n):
        lg = O.forward_logits(st, dil, O.one_hot(hist[:, -O.receptive_field(2, dil):], Q))[0, :, -1]
        np.testing.assert_allclose(lg.numpy(), lg_in[i].numpy(), rtol=1e-4, atol=1e-5)
        hist = torch.cat([hist, torch.tensor([[seq_in[i]]])], 1)
    assert not np.allclose(lg_in.numpy()[:n], z["logits"][:n], atol=1e-4)


def test_mulaw(golden):
    z = golden("mulaw")
    x = torch.from_numpy(z["x"])
    enc = O.mu_law_encode(x)
    assert np.array_equal(enc.numpy(), z["enc"].astype(np.int64))
    assert int(O.mu_law_encode(torch.tensor([0.0]))[0]) == 128
    dec = O.mu_law_decode(torch.arange(256))
    assert np.array_equal(dec.numpy(), z["dec"])
    # threshold table reproduces the reference on the golden inputs
    thr = z["thresholds"]
    xs = np.clip(z["x"], -1, 1)
    via_thr = np.searchsorted(thr[1:], xs, side="right")
    assert np.array_equal(via_thr, z["enc"].astype(np.int64))


def test_loader(golden):
    z = golden("loader")
    items = [z[f"item{i}"] for i in range(int(z["n_items"]))]
    pieces = O.make_data_pieces(items, int(z["rf"]), int(z["window"]))
    assert len(pieces) == int(z["n_pieces"])
    for i, p in enumerate(pieces):
        assert np.array_equal(p["audio_piece"].numpy(), z[f"piece{i}"])
        assert np.array_equal(p["audio_target"].numpy(), z[f"target{i}"])
    oh = O.one_hot_encode_reference(torch.from_numpy(z["onehot_in"]))
    assert np.array_equal(oh.numpy(), z["onehot_out"])
    true_oh = O.one_hot(torch.from_numpy(z["onehot_in"].astype(np.int64))[None], 256)[0]
    assert not np.array_equal(true_oh.numpy(), z["onehot_out"])      # fact 8


@pytest.mark.parametrize("name", ["ae_tile", "ae_bias"])
def test_autoencoder(golden, name):
    z = golden(name)
    dil = [int(d) for d in z["dilations"]]
    st = _state(z)
    cond = {k[len("cond."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("cond.")}
    x = O.one_hot(torch.from_numpy(z["idx"].astype(np.int64)), int(z["Q"]))
    lg = O.ae_forward_logits(st, cond, dil, x, int(z["pool"]))
    np.testing.assert_allclose(lg.numpy(), z["logits"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(O.scrambled_softmax(lg).numpy(), z["probs"], rtol=1e-5, atol=1e-8)
    shapes = O.ae_param_shapes(dil, int(z["Re"]), int(z["De"]), int(z["BW"]), int(z["Rd"]), int(z["Dd"]), int(z["Sd"]),
                               int(z["Q"]), bool(z["use_bias"]))
    assert [k for k, _ in shapes] == [k[len("state."):] for k in z.files if k.startswith("state.")]
    for k, s in shapes:
        assert tuple(st[k].shape) == tuple(s), k
