"""Drop-in for the reference's `wavenet/fast_generate.py`: incremental (fast-wavenet) generation.

Reference: predict_next :13-141, generate :144-179.  The per-sample recurrence runs inside
libwavenet_b200.so (csrc/gen.cu, csrc/fast_gen.cu) with per-block ring buffers instead of the
reference's shift-copied queues; `generate` runs all steps of all streams in one launch sequence
instead of 160 000 Python iterations.

Behavioural switches (defaults = the reference):
  queue_push = "output"  the reference pushes each block's OUTPUT into its queue (:128-129) although
                         the prime branch fills queues with block INPUTS (:42); "input" is the
                         consistent variant that equals the full forward.
  uniforms   = None      greedy topk(1) as in the reference (:139-140); a tensor of uniform draws
                         switches to inverse-CDF sampling (extension).
"""
from __future__ import annotations

import ctypes as C
import json
import os
from collections import OrderedDict

import torch

from .. import _lib as L
from .audio_func import mu_law_decode
from .model import wavenet


class GenState(OrderedDict):
    """The `state_queue` returned by predict_next.  It carries the device-resident ring buffers;
    the reference-layout entries ('causal_layer': (1,Q,1), 'block_k': (1,R,d_k), oldest first) are
    exported from the device lazily, the first time the dict is read."""

    def __init__(self, net, n_streams, state, last_pick_dev, mode):
        super().__init__()
        self._net, self._n, self._state, self._mode = net, n_streams, state, mode
        self._fresh = False

    def _materialize(self):
        if self._fresh:
            return
        net, e, lib = self._net, self._net.engine, L.load()
        sum_d = sum(net.dilations)
        dev = self._state.device
        q = torch.empty(self._n, sum_d, net.residual_channels, dtype=torch.float32, device=dev)
        last = torch.empty(self._n, dtype=torch.int64, device=dev)
        L.check(lib.wn_gen_export(e.handle, self._mode, self._n, L.ptr(self._state), L.ptr(q), L.ptr(last), L.stream_ptr()))
        OrderedDict.clear(self)
        oh = torch.zeros(self._n, net.quantization_channels, 1, device=dev)
        oh.scatter_(1, last.view(-1, 1, 1), 1.0)
        OrderedDict.__setitem__(self, 'causal_layer', oh)
        off = 0
        for i, d in enumerate(net.dilations):
            OrderedDict.__setitem__(self, 'block_' + str(i + 1), q[:, off:off + d, :].permute(0, 2, 1).contiguous())
            off += d
        self._fresh = True

    def _stale(self):
        self._fresh = False

    def __getitem__(self, k):
        self._materialize()
        return OrderedDict.__getitem__(self, k)

    def __iter__(self):
        self._materialize()
        return OrderedDict.__iter__(self)

    def __len__(self):
        return 1 + len(self._net.dilations)

    def keys(self):
        self._materialize()
        return OrderedDict.keys(self)

    def items(self):
        self._materialize()
        return OrderedDict.items(self)

    def values(self):
        self._materialize()
        return OrderedDict.values(self)


def _codes_of(note: torch.Tensor) -> torch.Tensor:
    """(N,Q,T) one-hot float -> (N,T) int64; raises if the note is not one-hot."""
    mx, idx = note.max(dim=1)
    if not bool(((mx == 1.0) & (note.sum(dim=1) == 1.0)).all()):
        raise ValueError("predict_next: `note` must be one-hot encoded (as fast_generate.generate feeds it)")
    return idx.to(torch.int64).contiguous()


def _gen_mode(net) -> int:
    return L.MODES[net.gen_mode]


def _prime(net, codes, uniforms=None, want_logits=False):
    e, lib = net.engine, L.load()
    params = net._params()
    e.ensure_flat(params)
    n, rf = codes.shape
    if uniforms is not None:
        uniforms = uniforms.to(codes.device, torch.float32).contiguous()
    mode32 = L.MODE_FP32                  # the prime is one full forward; it runs in fp32
    packed = e.packed(mode32, params)
    nbytes = C.c_size_t()
    L.check(lib.wn_gen_state_bytes(e.handle, _gen_mode(net), n, C.byref(nbytes)))
    state = torch.zeros(nbytes.value, dtype=torch.uint8, device=codes.device)
    ws = e.workspace(mode32, n, rf)
    out = torch.empty(n, dtype=torch.int64, device=codes.device)
    logits = torch.empty(n, net.quantization_channels, dtype=torch.float32, device=codes.device)
    L.check(lib.wn_gen_prime(e.handle, mode32, n, L.ptr(codes), L.ptr(packed), L.ptr(state), L.ptr(ws), ws.numel(),
                             L.ptr(uniforms), L.ptr(out), L.ptr(logits), L.stream_ptr()))
    e._ws_gen += 1
    return out, state, (logits if want_logits else None)


def _steps(net, state, first_note, n_steps, queue_push="output", uniforms=None, want_logits=False):
    e, lib = net.engine, L.load()
    params = net._params()
    e.ensure_flat(params)
    mode = _gen_mode(net)
    packed = e.packed(mode, params)
    n = first_note.shape[0]
    out = torch.empty(n_steps, n, dtype=torch.int64, device=first_note.device)
    logits = torch.empty(n_steps, n, net.quantization_channels, dtype=torch.float32,
                         device=first_note.device) if want_logits else None
    if uniforms is not None:
        uniforms = uniforms.to(first_note.device, torch.float32).contiguous()
        assert uniforms.shape == (n_steps, n)
    L.check(lib.wn_gen_steps(e.handle, mode, n, n_steps, L.PUSH[queue_push], L.ptr(first_note), L.ptr(packed),
                             L.ptr(state), L.ptr(uniforms), L.ptr(out), L.ptr(logits), L.stream_ptr()))
    return out, logits


def predict_next(net, note, state_queue=None, queue_push="output"):
    """Same contract as the reference (fast_generate.py:13-141): with state_queue=None `note` is the
    (1,Q,rf) one-hot priming piece, otherwise the (1,Q,1) one-hot of the newest sample.  Returns
    (predict: LongTensor[1], state_queue).  Extension: a leading dimension N > 1 runs N independent
    streams and `predict` has N entries."""
    if not note.is_cuda:
        raise L.WavenetB200Error("predict_next: `note` must be a CUDA tensor (no CPU fallback)")
    if state_queue is None:
        assert note.size()[2] == net.receptive_field
        codes = _codes_of(note)
        predict, state, _ = _prime(net, codes)
        return predict, GenState(net, codes.shape[0], state, predict, _gen_mode(net))
    assert note.size()[2] == 1
    if not isinstance(state_queue, GenState):
        state_queue = import_state(net, state_queue)
    codes = _codes_of(note).view(-1)
    out, _ = _steps(net, state_queue._state, codes, 1, queue_push)
    state_queue._stale()
    return out[0], state_queue


def import_state(net, queues) -> GenState:
    """Build a device state from a reference-layout OrderedDict (e.g. one produced by the reference)."""
    e, lib = net.engine, L.load()
    params = net._params()
    e.ensure_flat(params)
    dev = e.flat.device
    n = queues['causal_layer'].shape[0]
    last = _codes_of(queues['causal_layer'].to(dev).float()).view(-1)
    blocks = [queues['block_' + str(i + 1)].to(dev).float().permute(0, 2, 1) for i in range(len(net.dilations))]
    q = torch.cat(blocks, dim=1).contiguous()
    nbytes = C.c_size_t()
    mode = _gen_mode(net)
    L.check(lib.wn_gen_state_bytes(e.handle, mode, n, C.byref(nbytes)))
    state = torch.zeros(nbytes.value, dtype=torch.uint8, device=dev)
    L.check(lib.wn_gen_import(e.handle, mode, n, L.ptr(state), L.ptr(q), L.ptr(last), L.stream_ptr()))
    return GenState(net, n, state, None, mode)


def generate_codes(net, n_samples, start_codes=None, n_streams=1, queue_push="output", uniforms=None,
                   return_logits=False):
    """The loop of generate() (:158-172) on the device: prime, then n_samples-1 steps.
    Returns (n_samples, n_streams) int64 codes [and the (n_samples, n_streams, Q) logits]."""
    dev = next(net.parameters()).device
    Q, rf = net.quantization_channels, net.receptive_field
    if start_codes is None:
        start_codes = torch.full((n_streams, rf), Q // 2, dtype=torch.int64, device=dev)     # one-hot(128), :159-160
    start_codes = start_codes.to(dev, torch.int64).contiguous()
    n_streams = start_codes.shape[0]
    if uniforms is not None:
        uniforms = uniforms.to(dev, torch.float32).contiguous()
    first, state, lg0 = _prime(net, start_codes, None if uniforms is None else uniforms[0].contiguous(), return_logits)
    if n_samples == 1:
        return (first.view(1, -1), lg0.unsqueeze(0)) if return_logits else first.view(1, -1)
    rest, lg = _steps(net, state, first, n_samples - 1, queue_push,
                      None if uniforms is None else uniforms[1:].contiguous(), return_logits)
    codes = torch.cat([first.view(1, -1), rest], dim=0)
    if return_logits:
        return codes, torch.cat([lg0.unsqueeze(0), lg], dim=0)
    return codes


def _write_wav(path, audio, sr):
    import numpy as np
    try:
        from scipy.io import wavfile
        wavfile.write(path, sr, np.asarray(audio, dtype=np.float32))
    except ImportError:                                    # pragma: no cover
        import wave
        pcm = (np.clip(audio, -1, 1) * 32767).astype('<i2')
        with wave.open(path, 'wb') as w:
            w.setnchannels(1)
            w.setsampwidth(2)
            w.setframerate(sr)
            w.writeframes(pcm.tobytes())


def generate(model_path, model_name, generate_path, generate_name, start_piece=None, sr=16000, duration=10,
             params_path='./params/wavenet_params.json', queue_push="output", mode="auto"):
    """Same entry point as the reference (:144-179): load params + checkpoint, prime with
    one-hot(128) x rf unless `start_piece` (1,Q,rf) is given, generate duration*sr samples,
    mu-law decode and write the wav.  Returns the decoded float32 waveform (CPU tensor)."""
    from .train import load_model
    if os.path.exists(generate_path) is False:
        os.makedirs(generate_path)
    with open(params_path, 'r') as f:
        params = json.load(f)
    net = wavenet(**params, mode=mode)
    net = load_model(net, model_path, model_name)
    if net is None:
        raise FileNotFoundError(model_path + model_name)
    net = net.cuda()
    start_codes = None
    if start_piece is not None:
        start_codes = _codes_of(start_piece.cuda().float())
    codes = generate_codes(net, duration * sr, start_codes, queue_push=queue_push)[:, 0]
    print(codes.tolist()[:64], "...")
    audio = mu_law_decode(codes, net.quantization_channels).cpu()
    _write_wav(generate_path + generate_name, audio.numpy(), sr)
    return audio
