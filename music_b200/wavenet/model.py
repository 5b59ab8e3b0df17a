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

    mode   : "auto" (default) = the bf16 tcgen05 tensor-core path whenever this build has kernels for the shape
             (residual, dilation <= 64 channels, skip = quantization = 256), otherwise the fp32 path with a warning;
             "bf16" insists on the tensor-core path (raises for other shapes); "fp32" = SIMT check mode (1e-4 parity).
    parity : "reference" reproduces the reference's output exactly as it is - softmax over flat
             256-chunks of the (B,Q,W) buffer (model.py:142-144); "corrected" gives one softmax per
             time step.
    """

    def __init__(self, filter_width, dilations, dilation_channels, residual_channels, skip_channels,
                 quantization_channels, use_bias, *, mode: str = "auto", parity: str = "reference"):
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
        if mode not in L.MODE_NAMES:
            raise ValueError(f"mode must be one of {list(L.MODE_NAMES)}")
        if parity not in L.ROWS:
            raise ValueError(f"parity must be one of {list(L.ROWS)}")
        self._mode_arg, self.parity = mode, parity
        self._engine = None

    @property
    def mode(self) -> str:
        """The arithmetic mode the training / forward kernels run in ("fp32" or "bf16"); "auto" is resolved against the
        library once (wn_model_supports) and says so loudly when it has to settle for the fp32 SIMT path."""
        if self._mode_arg != "auto":
            return self._mode_arg
        r = getattr(self, "_mode_resolved", None)
        if r is None:
            ok = bool(L.load().wn_model_supports(self.engine.handle, 0))
            r = self._mode_resolved = "bf16" if ok else "fp32"
            if not ok:
                import warnings
                warnings.warn("music_b200: no tcgen05 kernels for residual=%d dilation=%d skip=%d quantization=%d channels; "
                              "mode='auto' runs this model in the fp32 SIMT mode (~20x slower)" %
                              (self.residual_channels, self.dilation_channels, self.skip_channels, self.quantization_channels))
        return r

    @mode.setter
    def mode(self, value: str):
        if value not in L.MODE_NAMES:
            raise ValueError(f"mode must be one of {list(L.MODE_NAMES)}")
        self._mode_arg = value
        self._mode_resolved = None

    @property
    def gen_mode(self) -> str:
        """Mode of the incremental-generation kernels: the half-precision kernel exists for 64/64/256/256 only."""
        if self._mode_arg == "fp32":
            return "fp32"
        ok = bool(L.load().wn_model_supports(self.engine.handle, 1))
        if not ok and self._mode_arg == "bf16":
            return "bf16"          # explicit request: let the library refuse loudly
        return "bf16" if ok else "fp32"

    def invalidate(self):
        """Call after writing parameters through `.data` (p.data.copy_(), legacy optimizers, EMA): such writes do not
        bump the version counters the packed-weight cache is keyed on."""
        if self._engine is not None:
            self._engine.invalidate_packed()

    def __deepcopy__(self, memo):
        """The C plan handle is owned by one Engine: a copy gets its own, built lazily."""
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k == "_engine" else copy.deepcopy(v, memo)
        return new

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_engine"] = None
        return st

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
