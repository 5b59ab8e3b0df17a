"""Loads the UNMODIFIED reference from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box: nothing that runs
there may import this module.  It is used by ``oracle/make_golden.py`` (to freeze golden
vectors under tests/golden/) and by tests that are skipped when the reference is absent.

* ``wavenet/model.py`` and ``wavenet_autoencoder/model1.py`` import and run as they are.
* ``wavenet/audio_func.py`` needs a stub ``librosa`` module to import (librosa is only used
  by ``trim_silence``).
* ``wavenet/fast_generate.py`` cannot be imported (imports librosa/train, calls
  ``generate()`` at import time, :182-186): ``predict_next`` (:13-141) is extracted from the
  source text by AST and exec'd with exactly two textual shims, at :75 and :103, where
  ``layer_input[:, :, -1] = note.data`` assigns a (1,C,1) tensor into a (1,C) slot - legal
  in the torch of the reference's era, a RuntimeError today.
* the autoencoder hard-codes ``.cuda()`` on its per-call conditioning convs
  (model1.py:178,216); ``capture_ae_forward`` makes ``.cuda()`` an identity on CPU and
  records the weights of every Conv1d constructed during the call.
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys
import types
from collections import OrderedDict

import torch
import torch.nn as nn

REF = os.environ.get("MUSIC_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "wavenet", "model.py"))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def wavenet_module():
    return _load(os.path.join(REF, "wavenet", "model.py"), "_ref_wavenet_model")


def ae_module():
    return _load(os.path.join(REF, "wavenet_autoencoder", "model1.py"), "_ref_ae_model1")


def audio_func_module():
    stub = types.ModuleType("librosa")
    had = sys.modules.get("librosa")
    sys.modules["librosa"] = stub
    try:
        return _load(os.path.join(REF, "wavenet", "audio_func.py"), "_ref_audio_func")
    finally:
        if had is None:
            sys.modules.pop("librosa", None)
        else:
            sys.modules["librosa"] = had


def data_module():
    return _load(os.path.join(REF, "wavenet", "faster_audio_data.py"), "_ref_faster_audio_data")


def predict_next_fn():
    """AST-extract fast_generate.predict_next with the two shape shims."""
    path = os.path.join(REF, "wavenet", "fast_generate.py")
    src = open(path).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "predict_next")
    lines = src.splitlines()[fn.lineno - 1:fn.end_lineno]
    text = "\n".join(lines)
    n_before = text.count("= note.data")
    text = text.replace("layer_input[:, :, -1] = note.data",
                        "layer_input[:, :, -1] = note.data.reshape(batch_size, channels)")
    text = text.replace("new_state[:, :, -1] = note.data",
                        "new_state[:, :, -1] = note.data.reshape(state.size(0), state.size(1))")
    assert n_before == 2 and text.count("= note.data.reshape") == 2, "shim sites moved"
    import torch.nn.functional as F
    from torch.autograd import Variable
    ns = {"torch": torch, "F": F, "Variable": Variable, "OrderedDict": OrderedDict}
    exec(compile(text, path, "exec"), ns)
    return ns["predict_next"]


def capture_ae_forward(net, x):
    """Run the reference AE forward on CPU; returns (probs, cond_state) where cond_state holds
    the per-call random conditioning convs as cond.{i}.weight/bias in creation order."""
    created = []
    orig_init = nn.Conv1d.__init__
    orig_cuda = nn.Module.cuda

    def spy_init(self, *a, **k):
        orig_init(self, *a, **k)
        created.append(self)

    nn.Conv1d.__init__ = spy_init
    nn.Module.cuda = lambda self, *a, **k: self
    try:
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):     # forward prints rf on every call (:261)
            out = net(x)
    finally:
        nn.Conv1d.__init__ = orig_init
        nn.Module.cuda = orig_cuda
    cond = OrderedDict()
    for i, c in enumerate(created):
        cond[f"cond.{i}.weight"] = c.weight.detach().clone()
        cond[f"cond.{i}.bias"] = c.bias.detach().clone()
    return out, cond
