"""Shared helpers for the GPU parity tests."""
import numpy as np
import torch

from oracle import wavenet_oracle as O


def state_of(z, prefix="state."):
    return {k[len(prefix):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix)}


def cfg_state(z):
    dil = [int(d) for d in z["dilations"]]
    if any(k.startswith("state.") for k in z.files):
        st = state_of(z)
    else:
        st = O.init_wavenet_state(dil, int(z["D"]), int(z["R"]), int(z["S"]), int(z["Q"]), bool(z["use_bias"]),
                                  seed=int(z["seed"]), scale=float(z["scale"]))
    return dil, st


def make_net(z, st, mode="fp32", parity="reference"):
    from music_b200.wavenet.model import wavenet
    dil = [int(d) for d in z["dilations"]]
    net = wavenet(2, dil, int(z["D"]), int(z["R"]), int(z["S"]), int(z["Q"]), bool(z["use_bias"]), mode=mode, parity=parity)
    net.load_state_dict(st)
    return net.cuda()


def build_net(dil, R, D, S, Q, bias, st, mode="fp32", parity="reference"):
    from music_b200.wavenet.model import wavenet
    net = wavenet(2, dil, D, R, S, Q, bias, mode=mode, parity=parity)
    net.load_state_dict(st)
    return net.cuda()


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def max_rel(a, b):
    """max |a-b| / max |b| - the 'relative error' the parity bars are stated in."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
