"""Freeze golden vectors from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py        # writes tests/golden/*.npz

Every array under tests/golden/ is an output of /root/reference code (wavenet/model.py,
wavenet/audio_func.py, wavenet/fast_generate.py:predict_next via the two-line shim of
oracle/ref_loader.py, wavenet/faster_audio_data.py, wavenet_autoencoder/model1.py) executed
by this container's torch CPU build, from inputs and weights that are either stored beside
them or regenerated from seeds by oracle/wavenet_oracle.py.  The oracle is then checked
against these files by tests/test_oracle_golden.py, here and on the GPU box.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_loader as RL          # noqa: E402
from oracle import wavenet_oracle as O       # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
torch.set_num_threads(4)


def _np(d):
    return {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def ref_net(cfg, state):
    M = RL.wavenet_module()
    net = M.wavenet(2, cfg["dilations"], cfg["D"], cfg["R"], cfg["S"], cfg["Q"], cfg["use_bias"])
    net.load_state_dict(state)
    return net


def run_ref_train(net, x, target, lr=1e-3):
    """Reference forward + the restated step of wavenet/train.py:171-182 with Adam."""
    cap = {}
    h = net.post_process_2.register_forward_hook(lambda m, i, o: cap.__setitem__("logits", o.detach().clone()))
    opt = torch.optim.Adam(net.parameters(), lr=lr)
    opt.zero_grad()
    probs = net(x)
    loss = nn.CrossEntropyLoss()(probs, target.view(-1))
    loss.backward()
    grads = {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p))
             for k, p in net.named_parameters()}
    opt.step()
    h.remove()
    after = {k: v.detach().clone() for k, v in net.state_dict().items()}
    return cap["logits"], probs.detach(), float(loss.detach()), grads, after


def case_wavenet(name, cfg, B, W, seed, scale, dense, lr=1e-3, subsample=None, store_state=True):
    state = O.init_wavenet_state(cfg["dilations"], cfg["D"], cfg["R"], cfg["S"], cfg["Q"],
                                 cfg["use_bias"], seed=seed, scale=scale)
    rf = O.receptive_field(2, cfg["dilations"])
    L = rf + W - 1
    g = torch.Generator().manual_seed(seed + 1)
    if dense:
        x = torch.randn(B, cfg["Q"], L, generator=g)
        idx = None
    else:
        idx = O.mu_law_encode(O.synthetic_audio(B, L + 1, seed=seed + 2), cfg["Q"])
        x = O.one_hot(idx[:, :L], cfg["Q"])
    if idx is not None:
        target = idx[:, rf:rf + W].contiguous()
    else:
        target = torch.randint(0, cfg["Q"], (B, W), generator=g)
    net = ref_net(cfg, state)
    logits, probs, loss, grads, after = run_ref_train(net, x, target, lr)
    out = {"dilations": np.asarray(cfg["dilations"]), "R": cfg["R"], "D": cfg["D"], "S": cfg["S"], "Q": cfg["Q"],
           "use_bias": int(cfg["use_bias"]), "B": B, "W": W, "L": L, "seed": seed, "scale": scale, "lr": lr,
           "target": target.numpy().astype(np.int16), "loss": np.float64(loss)}
    if dense:
        out["x"] = x.numpy()
    else:
        out["idx"] = idx.numpy().astype(np.int16)
    if subsample is None:
        out["logits"] = logits.numpy()
        out["probs"] = probs.numpy()
        for k, v in grads.items():
            out["grad." + k] = v.numpy()
        for k, v in after.items():
            out["after." + k] = v.numpy()
    else:
        rows = np.arange(0, probs.shape[0], subsample)
        out["rows"] = rows
        out["probs_rows"] = probs.numpy()[rows]
        out["logits_cols"] = logits.numpy()[:, :, ::subsample]
        for k, v in grads.items():
            out["gradnorm." + k] = np.float64(v.double().norm())
            out["gradhead." + k] = v.reshape(-1)[:64].numpy()
        for k, v in after.items():
            out["afterhead." + k] = v.reshape(-1)[:64].numpy()
    if store_state:
        for k, v in state.items():
            out["state." + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", loss, "L", L)


def case_generate(name, cfg, n_steps, seed, scale):
    state = O.init_wavenet_state(cfg["dilations"], cfg["D"], cfg["R"], cfg["S"], cfg["Q"],
                                 cfg["use_bias"], seed=seed, scale=scale)
    net = ref_net(cfg, state)
    predict_next = RL.predict_next_fn()
    rf = net.receptive_field
    g = torch.Generator().manual_seed(seed + 7)
    prime_idx = torch.randint(0, cfg["Q"], (1, rf), generator=g)
    note = O.one_hot(prime_idx, cfg["Q"])
    cap = []
    h = net.post_process_2.register_forward_hook(lambda m, i, o: cap.append(o.detach().reshape(-1).clone()))
    picks, queue = [], None
    for i in range(n_steps):
        p, queue = predict_next(net, note, queue)
        k = int(p[0])
        picks.append(k)
        note = torch.zeros(1, cfg["Q"], 1)
        note[:, k, :] = 1.0
    h.remove()
    out = {"dilations": np.asarray(cfg["dilations"]), "R": cfg["R"], "D": cfg["D"], "S": cfg["S"], "Q": cfg["Q"],
           "use_bias": int(cfg["use_bias"]), "seed": seed, "scale": scale,
           "prime_idx": prime_idx.numpy().astype(np.int16), "picks": np.asarray(picks, dtype=np.int16),
           "logits": torch.stack(cap).numpy()}
    for k, v in queue.items():
        out["queue." + k] = v.detach().numpy()
    for k, v in state.items():
        out["state." + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "picks", picks[:12])


def case_mulaw():
    A = RL.audio_func_module()
    g = torch.Generator().manual_seed(99)
    x = torch.cat([
        torch.linspace(-1.25, 1.25, 20001),
        torch.randn(20000, generator=g) * 0.3,
        torch.tensor([0.0, -0.0, 1.0, -1.0, 1e-8, -1e-8, 1e-4, -1e-4, 3.0, -3.0, 0.5, -0.5]),
    ]).float()
    enc = A.mu_law_encode(x, 256)
    codes = torch.arange(256)
    dec = A.mu_law_decode(codes, 256)
    # encode thresholds: smallest float with enc(x) >= k, by bisection over the float order, on the reference fn
    def enc1(v):
        t = torch.full((16,), float(v), dtype=torch.float32)
        return int(A.mu_law_encode(t, 256)[0])
    def f2o(f):     # float32 -> monotone int
        b = np.float32(f).view(np.int32).item()
        return b if b >= 0 else -(b & 0x7FFFFFFF)
    def o2f(o):
        b = o if o >= 0 else (-o) | 0x80000000
        return np.uint32(b & 0xFFFFFFFF).view(np.float32).item()
    thr = np.zeros(256, dtype=np.float32)
    thr[0] = -np.inf
    for k in range(1, 256):
        lo, hi = f2o(-1.0), f2o(1.0)          # enc(lo) = 0 < k <= 255 = enc(hi)
        while hi - lo > 1:
            mid = (lo + hi) // 2
            if enc1(o2f(mid)) >= k:
                hi = mid
            else:
                lo = mid
        thr[k] = o2f(hi)
    np.savez_compressed(os.path.join(OUT, "mulaw.npz"), x=x.numpy(), enc=enc.numpy().astype(np.int16),
                        dec=dec.numpy(), thresholds=thr)
    print("mulaw", enc[:5].tolist(), dec[:3].tolist(), thr[126:131])


def case_loader():
    D = RL.data_module()
    rng = np.random.RandomState(5)
    items = [rng.randint(0, 256, size=n).astype(np.int32) for n in (45, 23, 9, 61)]
    ds = D.audio_dataset.__new__(D.audio_dataset)
    ds.receptive_field, ds.window_length = 10, 12
    pieces = ds._make_data_pieces(items)
    out = {"n_items": len(items), "rf": 10, "window": 12, "n_pieces": len(pieces)}
    for i, it in enumerate(items):
        out[f"item{i}"] = it
    for i, p in enumerate(pieces):
        out[f"piece{i}"] = p["audio_piece"].numpy()
        out[f"target{i}"] = p["audio_target"].numpy()
    oh = D.one_hot_encode({"audio_piece": torch.from_numpy(items[2]), "audio_target": torch.zeros(1)}, False, 256)
    out["onehot_in"] = items[2]
    out["onehot_out"] = oh["audio_piece"].numpy()
    np.savez_compressed(os.path.join(OUT, "loader.npz"), **out)
    print("loader pieces", len(pieces))


def case_ae(name, cfg, B, W, seed):
    M = RL.ae_module()
    torch.manual_seed(seed)
    net = M.wavenet_autoencoder(2, cfg["Q"], cfg["dilations"], cfg["Re"], cfg["De"], cfg["BW"], cfg["pool"],
                                cfg["Rd"], cfg["Dd"], cfg["Sd"], cfg["use_bias"])
    rf = net.receptive_field
    L = rf + W - 1
    g = torch.Generator().manual_seed(seed + 1)
    idx = torch.randint(0, cfg["Q"], (B, L), generator=g)
    x = O.one_hot(idx, cfg["Q"])
    cap = {}
    h = net.connection_2.register_forward_hook(lambda m, i, o: cap.__setitem__("logits", o.detach().clone()))
    probs, cond = RL.capture_ae_forward(net, x)
    h.remove()
    out = {"dilations": np.asarray(cfg["dilations"]), "B": B, "W": W, "L": L,
           "idx": idx.numpy().astype(np.int16), "probs": probs.detach().numpy(), "logits": cap["logits"].numpy()}
    for k in ("Q", "Re", "De", "BW", "pool", "Rd", "Dd", "Sd"):
        out[k] = cfg[k]
    out["use_bias"] = int(cfg["use_bias"])
    for k, v in net.state_dict().items():
        out["state." + k] = v.numpy()
    for k, v in cond.items():
        out["cond." + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "L", L, "frames", W // cfg["pool"], "probs", tuple(probs.shape))


def main():
    os.makedirs(OUT, exist_ok=True)
    tiny = dict(dilations=[1, 2, 4, 8, 1, 2, 4, 8], R=16, D=16, S=32, Q=256, use_bias=False)
    case_wavenet("wn_tiny_onehot", tiny, B=2, W=37, seed=11, scale=2.0, dense=False)
    case_wavenet("wn_tiny_dense", tiny, B=2, W=300, seed=12, scale=1.0, dense=True)
    bias = dict(dilations=[1, 2, 4, 1, 2, 4], R=8, D=16, S=24, Q=256, use_bias=True)
    case_wavenet("wn_bias_dense", bias, B=3, W=65, seed=13, scale=1.5, dense=True)
    c64 = dict(dilations=[1, 2, 4, 8, 16, 32], R=64, D=64, S=256, Q=256, use_bias=False)
    case_wavenet("wn_c64_onehot", c64, B=2, W=512, seed=14, scale=1.5, dense=False, subsample=37, store_state=False)
    cfg1 = dict(dilations=[2 ** i for i in range(10)] * 3, R=32, D=32, S=256, Q=256, use_bias=False)
    case_wavenet("wn_cfg1", cfg1, B=1, W=16000 - 3071 + 1, seed=15, scale=1.0, dense=False,
                 subsample=431, store_state=False)
    case_generate("gen_tiny", tiny, n_steps=48, seed=21, scale=3.0)
    gen2 = dict(dilations=[1, 2, 4, 8, 16, 1, 2, 4, 8, 16], R=32, D=32, S=64, Q=256, use_bias=True)
    case_generate("gen_bias", gen2, n_steps=40, seed=22, scale=3.0)
    case_mulaw()
    case_loader()
    ae = dict(dilations=[1, 2, 4, 8, 1, 2, 4, 8], Q=256, Re=16, De=16, BW=32, pool=8, Rd=16, Dd=16, Sd=32, use_bias=False)
    case_ae("ae_tile", ae, B=2, W=88, seed=31)        # per-layer lengths not divisible by 11 frames -> tile branch
    ae2 = dict(dilations=[1, 2, 4], Q=256, Re=8, De=8, BW=16, pool=4, Rd=8, Dd=16, Sd=24, use_bias=True)
    case_ae("ae_bias", ae2, B=1, W=40, seed=32)


if __name__ == "__main__":
    main()
