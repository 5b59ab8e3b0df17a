"""Drop-in for the reference's `wavenet/model.py`: same class name, constructor, attributes,
submodule names, state_dict keys and `forward()` contract, with the arithmetic running in
libwavenet_b200.so (sm_100a CUDA) instead of nn.Conv1d / F.* calls.

Reference: wavenet/model.py:6-145 (`class wavenet`), :148-165 (`predict_next`).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib as L
from .._engine import ConvStackFunction, Engine, SoftmaxRowsFunction, _require_cuda


class wavenet(nn.Module):
    """Same signature as the reference (wavenet/model.py:8-15).  Two keyword-only extensions:

    mode   : "fp32" (SIMT check mode, 1e-4 parity) or "bf16" (tcgen05 tensor-core path).
    parity : "reference" reproduces the reference's output exactly as it is - softmax over flat
             256-chunks of the (B,Q,W) buffer (model.py:142-144); "corrected" gives one softmax per
             time step.
    """

    def __init__(self, filter_width, dilations, dilation_channels, residual_channels, skip_channels,
                 quantization_channels, use_bias, *, mode: str = "fp32", parity: str = "reference"):
        super(wavenet, self).__init__()
        self.filter_width = filter_width
        self.dilations = dilations
        self.dilation_channels = dilation_channels
        self.residual_channels = residual_channels
        self.skip_channels = skip_channels
        self.quantization_channels = quantization_channels
        self.use_bias = use_bias
        self.receptive_field = self.calc_receptive_field()
        self._init_causal_layer()
        self._init_dliation_layer()
        self._init_post_processing_layer()
        self.softmax = nn.Softmax(dim=1)
        if mode not in L.MODES:
            raise ValueError(f"mode must be one of {list(L.MODES)}")
        if parity not in L.ROWS:
            raise ValueError(f"parity must be one of {list(L.ROWS)}")
        self.mode, self.parity = mode, parity
        self._engine = None

    # -- construction: identical submodules / registration order (model.py:43-84) ---------------
    def calc_receptive_field(self):
        return (self.filter_width - 1) * (sum(self.dilations) + 1) + 1

    def _init_causal_layer(self):
        self.causal_layer = nn.Conv1d(self.quantization_channels, self.residual_channels, self.filter_width,
                                      bias=self.use_bias)

    def _init_dliation_layer(self):
        self.dilation_layer_stack = nn.ModuleList()
        for dilation in self.dilations:
            self.dilation_layer_stack.extend([
                nn.Conv1d(self.residual_channels, self.dilation_channels, self.filter_width, dilation=dilation,
                          bias=self.use_bias),                                       # filter
                nn.Conv1d(self.residual_channels, self.dilation_channels, self.filter_width, dilation=dilation,
                          bias=self.use_bias),                                       # gate
                nn.Conv1d(self.dilation_channels, self.residual_channels, 1, bias=self.use_bias),   # dense
                nn.Conv1d(self.dilation_channels, self.skip_channels, 1, bias=self.use_bias),       # skip
            ])

    def _init_post_processing_layer(self):
        self.post_process_1 = nn.Conv1d(self.skip_channels, self.skip_channels, 1, bias=self.use_bias)
        self.post_process_2 = nn.Conv1d(self.skip_channels, self.quantization_channels, 1, bias=self.use_bias)

    # -- engine ----------------------------------------------------------------------------------
    @property
    def engine(self) -> Engine:
        if self._engine is None:
            self._engine = Engine(list(self.dilations), self.residual_channels, self.dilation_channels,
                                  self.skip_channels, self.quantization_channels, self.use_bias, self.filter_width)
        return self._engine

    def _params(self):
        return list(self.parameters())

    # -- forward ---------------------------------------------------------------------------------
    def forward_logits(self, wave_sample=None, indices=None):
        """Pre-softmax (B,Q,W) tensor = output of post_process_2 (model.py:138)."""
        src = wave_sample if wave_sample is not None else indices
        _require_cuda(src, "input")
        L_in = src.shape[-1]
        if L_in - self.receptive_field + 1 <= 0:
            raise ValueError("wave sample not long enough")           # model.py:100-101
        x = idx = None
        if wave_sample is not None:
            if wave_sample.dim() != 3 or wave_sample.shape[1] != self.quantization_channels:
                raise ValueError("wave_sample must be (batch, quantization_channels, length)")
            x = wave_sample.detach().float().contiguous()
        else:
            idx = indices.detach().to(torch.int64).contiguous()
        return ConvStackFunction.apply(self.engine, L.MODES[self.mode], x, idx, *self._params())

    def forward(self, wave_sample):
        """(B,Q,L) float (dense; one-hot in practice) -> (B*W, Q) probabilities, W = L - rf + 1,
        rows in the reference's order (model.py:86-145)."""
        logits = self.forward_logits(wave_sample=wave_sample)
        return SoftmaxRowsFunction.apply(logits, L.ROWS[self.parity])

    def forward_indices(self, indices):
        """Extension: (B,L) integer mu-law codes, equivalent to forward(one_hot(indices)) without
        materialising the (B,Q,L) tensor (the causal layer becomes a gather)."""
        logits = self.forward_logits(indices=indices)
        return SoftmaxRowsFunction.apply(logits, L.ROWS[self.parity])


def predict_next(model, wave_var, quantization_channels=256):
    """Slow-path greedy pick: full forward, argmax of the LAST ROW of the (scrambled) output
    (wavenet/model.py:148-165)."""
    raw_out = model(wave_var)
    out = raw_out.view(-1, quantization_channels)
    last = out[-1, :].view(-1)
    _, predict = torch.topk(last, 1)
    return predict
