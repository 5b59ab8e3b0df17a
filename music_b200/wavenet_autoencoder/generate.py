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
