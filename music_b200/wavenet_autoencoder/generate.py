"""Drop-in for the reference's `wavenet_autoencoder/generate.py` (the slow, O(receptive field)-per-sample generator).

Reference: predict_next :13-19 (full forward on the window, greedy `topk(1)` over the LAST row of the (scrambled)
softmax output), generate :22-65 (window = the last rf+512 samples, start = one-hot(128) x (rf+512)).  The reference
loop slices `input_wav[:, -rf-511:]`, i.e. the CHANNEL axis, so its window grows by one sample per step instead of
sliding; `slide="time"` (default) slides along time as intended, `slide="reference"` reproduces the literal slicing.
"""
from __future__ import annotations

import os

import torch

from ..wavenet.audio_func import mu_law_decode


def predict_next(net, input_wav, quantization_channel=256, cond_weights=None):
    with torch.no_grad():
        out = net(input_wav, cond_weights=cond_weights).view(-1, quantization_channel)
    _, predict = torch.topk(out[-1, :].view(-1), 1)
    return int(predict)


def generate_codes(net, n_samples, start_piece=None, slide="time", cond_weights=None):
    """The loop of generate() (:44-58): returns the list of picked codes."""
    dev = next(net.parameters()).device
    Q, rf = net.quantization_channel, net.receptive_field
    if start_piece is None:
        start_piece = torch.zeros(1, Q, rf + 512)
        start_piece[:, 128, :] = 1.0
    input_wav = start_piece.to(dev)
    picks = []
    for _ in range(n_samples):
        p = predict_next(net, input_wav, Q, cond_weights)
        picks.append(p)
        note = torch.zeros(1, Q, 1, device=dev)
        note[0, p, 0] = 1.0
        if slide == "reference":
            input_wav = torch.cat((input_wav[:, -rf - 511:], note), 2)           # :58, slices channels: a no-op for Q <= rf+511
        else:
            input_wav = torch.cat((input_wav[:, :, -rf - 511:], note), 2)
    return picks


def generate(model_path, model_name, generate_path, generate_name, start_piece=None, sr=16000, duration=10, model_params=None,
             net=None):
    """Same entry point as the reference (:22-65).  `model_params` replaces its `./params/model_params.json` (which is
    not valid JSON in the reference tree); a ready `net` may be passed instead of a checkpoint."""
    from .model1 import wavenet_autoencoder
    from .train import load_model
    if not os.path.exists(generate_path):
        os.makedirs(generate_path)
    if net is None:
        net = load_model(wavenet_autoencoder(**model_params), model_path, model_name)
        if net is None:
            raise FileNotFoundError(model_path + model_name)
    net = net.cuda()
    codes = generate_codes(net, duration * sr, start_piece)
    wave = mu_law_decode(torch.tensor(codes, dtype=torch.int64, device="cuda"), net.quantization_channel).cpu().numpy()
    from scipy.io import wavfile
    wavfile.write(generate_path + generate_name + ("" if generate_name.endswith(".wav") else ".wav"), sr, wave)
    return wave


# ---- incremental (fast) generation with the decoder: an extension (the reference only has the O(rf)-per-sample loop above) ----
def decoder_as_wavenet(net, mode="fp32"):
    """The autoencoder's decoder (model1.py:158-225) as a `wavenet` module (fp32 mode by default): same stack, the combined
    filter_gate conv split into the gate (first half of its outputs) and the filter (second half, :188-192),
    connection_1 / connection_2 as post_process_1 / post_process_2.  Weights are copied at every call.
    mode="bf16": the generation steps run on the half-precision cluster pipeline (64 residual / 64 dilation / 256 skip / 256
    quantization channels only; the prime is an fp32 forward in every mode)."""
    from ..wavenet.model import wavenet
    D = net.de_dilation_channel
    attr = "_decoder_wavenet" if mode == "fp32" else "_decoder_wavenet_" + mode
    dec = getattr(net, attr, None)
    if dec is None:
        dec = wavenet(net.filter_width, list(net.dilations), D, net.de_residual_channel, net.de_skip_channel, net.quantization_channel,
                      net.use_bias, mode=mode)
        object.__setattr__(net, attr, dec)
    src = net.state_dict()
    sd = {}

    def put(dst, key, rows=None):
        for suffix in ("weight", "bias"):
            k = f"{key}.{suffix}"
            if k in src:
                v = src[k]
                sd[f"{dst}.{suffix}"] = (v if rows is None else v[rows]).detach().clone()
    put("causal_layer", "de_causal_layer")
    for i in range(len(net.dilations)):
        put(f"dilation_layer_stack.{4 * i}", f"de_dilation_layer_stack.{3 * i}", slice(D, 2 * D))        # filter = second half
        put(f"dilation_layer_stack.{4 * i + 1}", f"de_dilation_layer_stack.{3 * i}", slice(0, D))        # gate = first half
        put(f"dilation_layer_stack.{4 * i + 2}", f"de_dilation_layer_stack.{3 * i + 1}")
        put(f"dilation_layer_stack.{4 * i + 3}", f"de_dilation_layer_stack.{3 * i + 2}")
    put("post_process_1", "connection_1")
    put("post_process_2", "connection_2")
    dec.load_state_dict(sd)
    return dec.to(next(net.parameters()).device)


def fast_generate_codes(net, encoding, total_len, n_samples, start_codes, cond_weights=None, forced=None, uniforms=None,
                        return_logits=False, mode="fp32"):
    """Incremental generation of `n_samples` codes per stream with the conditioned decoder, on the GPU.

    encoding   : (n_streams, bottleneck, frames) as `_encode` returns it (model1.py:137-156)
    total_len  : length L of the whole sequence the conditioning refers to (`_conditon`'s frame rule depends on it)
    start_codes: (n_streams, receptive_field) int codes of the first rf samples (the prime)
    forced     : optional (n_samples - 1, n_streams) codes fed instead of the picks (teacher forcing, for parity tests)
    mode       : "fp32" (default; logits equal the full forward's to 1e-4) or "bf16": the steps run on the weights-stationary
                 cluster pipeline with fp16 weight fragments (decoders of 64 / 64 / 256 / 256 channels, up to 576 streams), ~10 us
                 per step for the 30-block stack; logits within 1e-2 of the fp32 kernel's.  Without `forced` all steps are one launch.
    Returns (n_samples, n_streams) codes [and (n_samples, n_streams, Q) logits]: entry j predicts sample rf + j, exactly row j
    of the full forward over the finished sequence.  Per sample the cost is O(layers), not O(receptive field)."""
    import ctypes as C
    from .. import _lib as L
    from ..wavenet import fast_generate as FG
    dev = next(net.parameters()).device
    lib = L.init(dev.index if dev.index is not None else torch.cuda.current_device())
    dec = decoder_as_wavenet(net, mode)
    start_codes = start_codes.to(dev, torch.int64).contiguous()
    n, rf = start_codes.shape
    assert rf == net.receptive_field and total_len >= rf
    enc = encoding.to(dev, torch.float32).permute(0, 2, 1).contiguous()                 # (n, frames, BW)
    frames = enc.shape[1]
    N, D, S = len(net.dilations), net.de_dilation_channel, net.de_skip_channel
    cond = net._cond_flat(cond_weights, dev).detach()
    tab_fg = torch.empty(n, frames, N, 2 * D, dtype=torch.float32, device=dev)
    tab_head = torch.empty(n, frames, S, dtype=torch.float32, device=dev)
    ws = net._workspace(n, rf, dev, train=False)
    L.check(lib.wn_ae_cond_tables(net._plan(), n, frames, L.ptr(enc), L.ptr(cond), L.ptr(ws), L.ptr(tab_fg), L.ptr(tab_head),
                                  L.stream_ptr()))
    desc = L.wn_gen_cond(tab_fg.data_ptr(), tab_head.data_ptr(), frames, int(total_len), 1)
    params = dec._params()
    dec.engine.ensure_flat(params)
    handle = dec.engine.handle
    L.check(lib.wn_set_conditioning(handle, C.byref(desc)))
    try:
        with torch.no_grad():
            first, state, lg0 = FG._prime(dec, start_codes, None if uniforms is None else uniforms[0].contiguous(), True)
            codes, logits = [first], [lg0]
            note = first
            if forced is None and n_samples > 1:      # free-running: every remaining step in one launch
                u = None if uniforms is None else uniforms[1:n_samples].contiguous()
                out, lg = FG._steps(dec, state, first, n_samples - 1, "input", u, True)
                codes += list(out)
                logits += list(lg)
                n_samples = 1
            for j in range(1, n_samples):
                feed = note if forced is None else forced[j - 1].to(dev, torch.int64).contiguous()
                u = None if uniforms is None else uniforms[j:j + 1].contiguous()
                out, lg = FG._steps(dec, state, feed, 1, "input", u, True)
                note = out[0].contiguous()
                codes.append(note)
                logits.append(lg[0])
        torch.cuda.current_stream().synchronize()          # the tables must outlive the launches that read them
    finally:
        L.check(lib.wn_set_conditioning(handle, None))
    codes = torch.stack(codes)
    return (codes, torch.stack(logits)) if return_logits else codes
