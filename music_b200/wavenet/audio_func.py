"""Drop-in for the reference's `wavenet/audio_func.py`: mu-law companding on the GPU.

The per-sample work runs in libwavenet_b200.so (csrc/codec.cu).  To be bit-exact with the
reference's fp32 torch arithmetic the kernels use two small tables built once per Q on the host:
the encode thresholds (smallest float the reference formula maps to a code >= k, found by bisection
over the float order) and the Q decode values.  The tables are constants of the codec, like a
CRC table; no sample ever takes a host path.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib as L

_TABLES = {}


def _ref_encode_cpu(audio: torch.Tensor, q: int) -> torch.Tensor:
    # the reference formula, wavenet/audio_func.py:16-22 (used ONLY to tabulate bin edges)
    mu = torch.Tensor([q - 1]).float()
    safe = torch.abs(torch.clamp(audio, -1.0, 1.0))
    mag = torch.log1p(mu * safe) / torch.log1p(mu)
    sig = torch.sign(audio) * mag
    return ((sig + 1) / 2 * mu + 0.5).long()


def _ref_decode_cpu(codes: torch.Tensor, q: int) -> torch.Tensor:
    # wavenet/audio_func.py:35-39
    mu = torch.Tensor([q - 1]).float()
    sig = 2.0 * (codes.float() / mu) - 1.0
    mag = (1.0 / mu) * ((1.0 + mu) ** torch.abs(sig) - 1.0)
    return torch.sign(sig) * mag


def _ord_to_float(o: np.ndarray) -> np.ndarray:
    b = np.where(o >= 0, o, (-o) | 0x80000000).astype(np.uint64) & 0xFFFFFFFF
    return b.astype(np.uint32).view(np.float32)


def host_tables(q: int = 256):
    """(thresholds[q] float32 with [0] = -inf, values[q] float32)."""
    if q in _TABLES:
        return _TABLES[q]
    n = q - 1
    pad = (-n) % 16
    one = int(np.float32(1.0).view(np.int32))
    lo = np.full(n + pad, -one, dtype=np.int64)      # enc(-1) = 0 < k
    hi = np.full(n + pad, one, dtype=np.int64)       # enc(+1) = q-1 >= k
    ks = np.concatenate([np.arange(1, q), np.full(pad, 1)]).astype(np.int64)
    while np.any(hi - lo > 1):
        mid = (lo + hi) // 2
        enc = _ref_encode_cpu(torch.from_numpy(_ord_to_float(mid).copy()), q).numpy()
        ge = enc >= ks
        hi = np.where(ge, mid, hi)
        lo = np.where(ge, lo, mid)
    thr = np.empty(q, dtype=np.float32)
    thr[0] = -np.inf
    thr[1:] = _ord_to_float(hi[:n])
    codes = torch.arange(q + ((-q) % 16))
    vals = _ref_decode_cpu(codes, q).numpy()[:q].astype(np.float32)
    _TABLES[q] = (thr, vals)
    return _TABLES[q]


_DEV_TABLES = {}


def _device_tables(q: int, device):
    key = (q, str(device))
    if key not in _DEV_TABLES:
        thr, vals = host_tables(q)
        _DEV_TABLES[key] = (torch.from_numpy(thr).to(device), torch.from_numpy(vals).to(device))
    return _DEV_TABLES[key]


def _cuda_device(t: torch.Tensor):
    if t.is_cuda:
        return t.device if t.device.index is not None else torch.device("cuda", torch.cuda.current_device())
    if not torch.cuda.is_available():
        raise L.WavenetB200Error("mu-law codec needs a CUDA (sm_100a) device; music_b200 has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def mu_law_encode(audio, quantization_channels=256):
    """float tensor (any shape) -> int64 codes in [0, Q-1]; same device as the input
    (CPU inputs are staged through the GPU).  wavenet/audio_func.py:5-22."""
    dev = _cuda_device(audio)
    x = audio.detach().to(dev, torch.float32).contiguous()
    lib = L.init(dev.index)
    thr, _ = _device_tables(quantization_channels, dev)
    out = torch.empty(x.shape, dtype=torch.int64, device=dev)
    L.check(lib.wn_mulaw_encode(L.ptr(x), x.numel(), quantization_channels, L.ptr(thr), L.ptr(out), L.stream_ptr()))
    return out if audio.is_cuda else out.cpu()


def mu_law_decode(output, quantization_channels=256):
    """int codes -> float32 in [-1, 1].  wavenet/audio_func.py:24-39."""
    dev = _cuda_device(output)
    c = output.detach().to(dev, torch.int64).contiguous()
    lib = L.init(dev.index)
    _, vals = _device_tables(quantization_channels, dev)
    out = torch.empty(c.shape, dtype=torch.float32, device=dev)
    L.check(lib.wn_mulaw_decode(L.ptr(c), c.numel(), quantization_channels, L.ptr(vals), L.ptr(out), L.stream_ptr()))
    return out if output.is_cuda else out.cpu()


def trim_silence(audio, threshold, frame_length=2048):
    """Host-side helper of the offline data prep (wavenet/audio_func.py:41-55); needs librosa."""
    import librosa                                              # not part of the GPU path
    if audio.size < frame_length:
        frame_length = audio.size
    energy = librosa.feature.rms(y=audio, frame_length=frame_length)
    frames = np.nonzero(energy > threshold)
    indices = librosa.core.frames_to_samples(frames)[1]
    return audio[indices[0]:indices[-1]] if indices.size else audio[0:0]
