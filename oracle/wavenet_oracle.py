"""CPU oracle for the WaveNet hot path of deep-art-project/Music.

TEST INFRASTRUCTURE ONLY.  Nothing under ``music_b200/`` may import this file; it is
used by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs, and only as the checker or as the timed CPU baseline.

What it is: a functional restatement, on torch CPU fp32 (or fp64 with ``dtype=``), of the
algorithm in the reference's Python sources.  The arithmetic of the reference lives in a
third-party dependency, PyTorch (version unpinned by the reference: no requirements file;
API usage dates it to 0.2-0.3).  The oracle therefore executes the same torch CPU kernels
(`conv1d`, `sigmoid`, `tanh`, `relu`, `softmax`, `cross_entropy`, `avg_pool1d`,
`optim.Adam`) through the functional API, from an explicit ``state`` dict whose keys are
the reference ``state_dict`` keys.

Parity pinning: the reference has no tests or golden vectors for this path
(SURVEY.md section 8c).  The oracle is pinned instead against outputs of the reference
itself, imported unmodified from /root/reference in the build container by
``oracle/make_golden.py``; the outputs are committed under ``tests/golden/`` and
``tests/test_oracle_golden.py`` checks the oracle against them on every run.

Citations are relative to /root/reference.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------
# a2: receptive field                                   wavenet/model.py:43-44
# --------------------------------------------------------------------------------------
def receptive_field(filter_width: int, dilations: Sequence[int]) -> int:
    return (filter_width - 1) * (sum(dilations) + 1) + 1


# --------------------------------------------------------------------------------------
# a1: parameter construction                            wavenet/model.py:46-84
# --------------------------------------------------------------------------------------
def wavenet_param_shapes(dilations, dilation_channels, residual_channels, skip_channels,
                         quantization_channels, use_bias, filter_width=2):
    """Ordered (key, shape) list; the order is the reference state_dict order."""
    R, D, S, Q, fw = residual_channels, dilation_channels, skip_channels, quantization_channels, filter_width
    out = [("causal_layer.weight", (R, Q, fw))]
    if use_bias:
        out.append(("causal_layer.bias", (R,)))
    for i in range(len(dilations)):
        j = 4 * i
        for k, shp in enumerate([(D, R, fw), (D, R, fw), (R, D, 1), (S, D, 1)]):
            out.append((f"dilation_layer_stack.{j + k}.weight", shp))
            if use_bias:
                out.append((f"dilation_layer_stack.{j + k}.bias", (shp[0],)))
    out.append(("post_process_1.weight", (S, S, 1)))
    if use_bias:
        out.append(("post_process_1.bias", (S,)))
    out.append(("post_process_2.weight", (Q, S, 1)))
    if use_bias:
        out.append(("post_process_2.bias", (Q,)))
    return out


def init_wavenet_state(dilations, dilation_channels, residual_channels, skip_channels,
                       quantization_channels=256, use_bias=False, seed=0, filter_width=2,
                       scale=1.0) -> State:
    """Conv1d default init U(+-1/sqrt(in*k)) (what `nn.Conv1d` does in the reference's
    constructor, wavenet/model.py:47-84), from a private generator so it travels."""
    g = torch.Generator().manual_seed(seed)
    st: State = OrderedDict()
    for key, shp in wavenet_param_shapes(dilations, dilation_channels, residual_channels,
                                         skip_channels, quantization_channels, use_bias, filter_width):
        if key.endswith(".weight"):
            fan_in = shp[1] * shp[2]
            last_fan_in = fan_in
        else:
            fan_in = last_fan_in
        bound = scale / (fan_in ** 0.5)
        st[key] = (torch.rand(shp, generator=g) * 2 - 1) * bound
    return st


def _b(state: State, key: str):
    return state.get(key[:-len("weight")] + "bias") if key.endswith("weight") else None


# --------------------------------------------------------------------------------------
# a3-a6: forward                                        wavenet/model.py:86-145
# --------------------------------------------------------------------------------------
def forward_logits(state: State, dilations: Sequence[int], x: torch.Tensor,
                   return_intermediates: bool = False):
    """(B,Q,L) dense float -> pre-softmax (B,Q,W) tensor (output of post_process_2,
    wavenet/model.py:138), W = L - rf + 1 (:99)."""
    fw = state["causal_layer.weight"].shape[2]
    rf = receptive_field(fw, dilations)
    B, Q, L = x.shape
    W = L - rf + 1
    if W <= 0:
        raise ValueError("wave sample not long enough")          # :100-101
    cur = F.conv1d(x, state["causal_layer.weight"], state.get("causal_layer.bias"))   # :104
    skips = []
    inter = {"x": [cur], "z": []}
    for i, d in enumerate(dilations):                              # :108
        j = 4 * i
        kf, kg, kd, ks = (f"dilation_layer_stack.{j + k}.weight" for k in range(4))
        f = F.conv1d(cur, state[kf], _b(state, kf), dilation=d)    # :118
        g = F.conv1d(cur, state[kg], _b(state, kg), dilation=d)    # :119
        z = torch.sigmoid(g) * torch.tanh(f)                       # :120
        dense = F.conv1d(z, state[kd], _b(state, kd))              # :121
        cur = dense + cur[:, :, -dense.shape[2]:]                  # :122-124
        skips.append(F.conv1d(z[:, :, -W:], state[ks], _b(state, ks)))   # :127-129
        if return_intermediates:
            inter["x"].append(cur)
            inter["z"].append(z)
    total = sum(skips)                                             # :134
    total = F.relu(total)
    total = F.conv1d(total, state["post_process_1.weight"], state.get("post_process_1.bias"))
    total = F.relu(total)
    total = F.conv1d(total, state["post_process_2.weight"], state.get("post_process_2.bias"))  # :138
    if return_intermediates:
        return total, inter
    return total


def scrambled_softmax(logits: torch.Tensor) -> torch.Tensor:
    """`total.view(-1, Q)` on the contiguous (B,Q,W) buffer then softmax over dim 1
    (wavenet/model.py:142-144; `nn.Softmax()` with implicit dim = 1 for 2-D input).
    Rows are flat 256-chunks, not time steps."""
    Q = logits.shape[1]
    return torch.softmax(logits.contiguous().view(-1, Q), dim=1)


def forward_probs(state: State, dilations: Sequence[int], x: torch.Tensor) -> torch.Tensor:
    """What `wavenet.forward` returns: (B*W, Q) probabilities in reference row order."""
    return scrambled_softmax(forward_logits(state, dilations, x))


# --------------------------------------------------------------------------------------
# a7: loss                                              wavenet/train.py:146,177-179
# --------------------------------------------------------------------------------------
def loss_from_probs(probs: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """`nn.CrossEntropyLoss()(probs, target.view(-1))`: a second (log-)softmax applied to
    the probabilities."""
    return F.cross_entropy(probs, target.reshape(-1))


def loss_and_dlogits(logits: torch.Tensor, target: torch.Tensor):
    """Loss and d loss / d logits through the scrambled softmax + double-softmax CE."""
    lg = logits.detach().clone().requires_grad_(True)
    loss = loss_from_probs(scrambled_softmax(lg), target)
    (g,) = torch.autograd.grad(loss, lg)
    return loss.detach(), g


# --------------------------------------------------------------------------------------
# a8: train step                                        wavenet/train.py:169-182
# --------------------------------------------------------------------------------------
class TrainState:
    """Parameters (leaf tensors requiring grad) + a torch optimizer, as train.py builds
    them (`get_optimizer`, wavenet/train.py:28-42)."""

    def __init__(self, state: State, optimizer: str = "adam", lr: float = 1e-4, momentum: float = 0.9):
        self.params: State = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in state.items())
        ps = list(self.params.values())
        if optimizer == "sgd":
            self.opt = torch.optim.SGD(ps, lr=lr, momentum=momentum)
        elif optimizer == "rmsprop":
            self.opt = torch.optim.RMSprop(ps, lr=lr, momentum=momentum)
        elif optimizer == "adam":
            self.opt = torch.optim.Adam(ps, lr=lr)
        else:
            raise ValueError(optimizer)


def train_step(ts: TrainState, dilations: Sequence[int], x: torch.Tensor, target: torch.Tensor) -> float:
    """zero_grad -> forward -> CE(probabilities) -> backward -> step (train.py:171-182)."""
    ts.opt.zero_grad()
    probs = forward_probs(ts.params, dilations, x)
    loss = loss_from_probs(probs, target)
    loss.backward()
    ts.opt.step()
    return float(loss.detach())


def grads(state: State, dilations: Sequence[int], x: torch.Tensor, target: torch.Tensor):
    """(loss, {key: grad}) of the reference training objective, by autograd."""
    params = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in state.items())
    loss = loss_from_probs(forward_probs(params, dilations, x), target)
    gs = torch.autograd.grad(loss, list(params.values()), allow_unused=True)
    out = OrderedDict()
    for (k, p), g in zip(params.items(), gs):
        out[k] = torch.zeros_like(p) if g is None else g     # last layer's dense conv: unused (:121-124)
    return float(loss.detach()), out


# --------------------------------------------------------------------------------------
# one-hot helpers
# --------------------------------------------------------------------------------------
def one_hot(idx: torch.Tensor, Q: int = 256, dtype=torch.float32) -> torch.Tensor:
    """True one-hot, (B,L) int -> (B,Q,L)."""
    return F.one_hot(idx.long(), Q).permute(0, 2, 1).to(dtype).contiguous()


def one_hot_encode_reference(piece: torch.Tensor, Q: int = 256) -> torch.Tensor:
    """The loader's 'one-hot': (T,Q) one-hot then RESHAPE to (Q,T), not a transpose
    (wavenet/faster_audio_data.py:77-82)."""
    T = piece.shape[0]
    oh = np.zeros((T, Q))
    oh[np.arange(T), piece.numpy()] = 1.0
    return torch.FloatTensor(oh.reshape(Q, T))


def make_data_pieces(items: Sequence[np.ndarray], receptive_field_: int, window_length: int):
    """Windowing of wavenet/faster_audio_data.py:24-40, including its tail behaviour: when a
    remainder shorter than rf+window is left, the previous (piece, target) is appended
    again (`target`/`piece` are stale variables at :34-39)."""
    pieces = []
    piece = target = None
    for item in items:
        item = torch.from_numpy(np.asarray(item))
        while len(item) > receptive_field_:
            if len(item) >= receptive_field_ + window_length:
                piece = item[:receptive_field_ + window_length - 1]
                target = item[receptive_field_:receptive_field_ + window_length]
                item = item[window_length:]
            else:
                item = item[receptive_field_:]
            target = target.long()
            pieces.append({"audio_piece": piece, "audio_target": target})
    return pieces


# --------------------------------------------------------------------------------------
# a14/a15: mu-law codec                                 wavenet/audio_func.py:5-39
# --------------------------------------------------------------------------------------
def mu_law_encode(audio: torch.Tensor, quantization_channels: int = 256) -> torch.Tensor:
    mu = torch.Tensor([quantization_channels - 1]).float()              # :16-17
    safe_audio_abs = torch.abs(torch.clamp(audio, -1.0, 1.0))           # :18
    magnitude = torch.log1p(mu * safe_audio_abs) / torch.log1p(mu)      # :19
    signal = torch.sign(audio) * magnitude                              # :20
    encoded = (signal + 1) / 2 * mu + 0.5                               # :21
    return encoded.long()                                               # :22


def mu_law_decode(output: torch.Tensor, quantization_channels: int = 256) -> torch.Tensor:
    mu = torch.Tensor([quantization_channels - 1]).float()              # :35-36
    signal = 2.0 * (output.float() / mu) - 1.0                          # :37
    magnitude = (1.0 / mu) * ((1.0 + mu) ** torch.abs(signal) - 1.0)    # :38
    return torch.sign(signal) * magnitude                               # :39


# --------------------------------------------------------------------------------------
# a10-a12: incremental generation                       wavenet/fast_generate.py:13-141
# --------------------------------------------------------------------------------------
def _head(state: State, skip_sum: torch.Tensor) -> torch.Tensor:
    """(1,S,1) -> (Q,) pre-softmax (fast_generate.py:130-134)."""
    t = F.relu(skip_sum)
    t = F.conv1d(t, state["post_process_1.weight"], state.get("post_process_1.bias"))
    t = F.relu(t)
    t = F.conv1d(t, state["post_process_2.weight"], state.get("post_process_2.bias"))
    return t.reshape(-1)


def gen_prime(state: State, dilations: Sequence[int], note: torch.Tensor):
    """Prime branch (fast_generate.py:29-65): note is (1,Q,rf). Returns (logits(Q,), queues).
    Queues hold block INPUT history (:42)."""
    fw = state["causal_layer.weight"].shape[2]
    rf = receptive_field(fw, dilations)
    assert note.shape[2] == rf                                               # :30
    Q = note.shape[1]
    R = state["causal_layer.weight"].shape[0]
    queues = OrderedDict()
    cur = F.conv1d(note, state["causal_layer.weight"], state.get("causal_layer.bias"))  # :32
    queues["causal_layer"] = note[:, :, -1].contiguous().view(1, Q, 1).clone()          # :33-38
    skips = []
    for i, d in enumerate(dilations):
        queues[f"block_{i + 1}"] = cur[:, :, -d:].contiguous().view(1, R, d).clone()    # :41-47
        j = 4 * i
        kf, kg, kd, ks = (f"dilation_layer_stack.{j + k}.weight" for k in range(4))
        f = F.conv1d(cur, state[kf], _b(state, kf), dilation=d)
        g = F.conv1d(cur, state[kg], _b(state, kg), dilation=d)
        z = torch.sigmoid(g) * torch.tanh(f)
        dense = F.conv1d(z, state[kd], _b(state, kd))
        cur = dense + cur[:, :, -dense.shape[2]:]
        skips.append(F.conv1d(z[:, :, -1:], state[ks], _b(state, ks)))                  # :62-64
    return _head(state, sum(skips)), queues


def gen_step(state: State, dilations: Sequence[int], note: torch.Tensor, queues,
             queue_push: str = "output"):
    """Step branch (fast_generate.py:66-129): note is (1,Q,1). Each layer sees
    cat(queue, new) of d+1 columns so the dilated conv reads column 0 (oldest) and the new
    one (:71-95). After the layer the queue is shifted left and a vector appended (:99-104):
    the block OUTPUT in the reference (`note_out`, :128-129), the block
    INPUT when queue_push="input" (the mathematically consistent variant)."""
    assert note.shape[2] == 1                                                # :67
    newq = OrderedDict()
    cs = queues["causal_layer"]
    li = torch.cat([cs, note], dim=2)                                        # :73-75
    out = F.conv1d(li, state["causal_layer.weight"], state.get("causal_layer.bias"))   # :79-80
    newq["causal_layer"] = note.clone()                                      # :116 (queue length 1)
    skips = []
    for i, d in enumerate(dilations):
        note_in = out
        q = queues[f"block_{i + 1}"]
        li = torch.cat([q, note_in], dim=2)
        j = 4 * i
        kf, kg, kd, ks = (f"dilation_layer_stack.{j + k}.weight" for k in range(4))
        f = F.conv1d(li, state[kf], _b(state, kf), dilation=d)
        g = F.conv1d(li, state[kg], _b(state, kg), dilation=d)
        z = torch.sigmoid(g) * torch.tanh(f)
        dense = F.conv1d(z, state[kd], _b(state, kd))
        out = dense + li[:, :, -dense.shape[2]:]
        skips.append(F.conv1d(z[:, :, -1:], state[ks], _b(state, ks)))
        pushed = out if queue_push == "output" else note_in
        newq[f"block_{i + 1}"] = torch.cat([q[:, :, 1:], pushed], dim=2)     # :99-104
    return _head(state, sum(skips)), newq


def pick_greedy(logits: torch.Tensor) -> int:
    """softmax then topk(1) (fast_generate.py:138-140)."""
    p = torch.softmax(logits.view(1, -1), dim=1).view(-1)
    return int(torch.topk(p, 1)[1][0])


def pick_sampled(logits: torch.Tensor, u: float) -> int:
    """Extension (not in the reference, which is greedy): inverse-CDF over the fp32 softmax
    probabilities accumulated in index order with fp32 adds; returns the first index whose
    running sum exceeds u*total (total = the final running sum), clamped to Q-1."""
    p = torch.softmax(logits.view(1, -1).float(), dim=1).view(-1).numpy().astype(np.float32)
    c = np.float32(0.0)
    cs = np.empty_like(p)
    for k in range(p.shape[0]):
        c = np.float32(c + p[k])
        cs[k] = c
    thr = np.float32(np.float32(u) * cs[-1])
    k = int(np.searchsorted(cs, thr, side="right"))
    return min(k, p.shape[0] - 1)


def generate(state: State, dilations: Sequence[int], n_steps: int, start_piece: Optional[torch.Tensor] = None,
             queue_push: str = "output", uniforms: Optional[Sequence[float]] = None,
             return_logits: bool = False):
    """fast_generate.generate's loop (:158-172) without file I/O: prime with one-hot(128)*rf
    unless given, then n_steps picks. Returns list of ints (length n_steps)."""
    fw = state["causal_layer.weight"].shape[2]
    Q = state["causal_layer.weight"].shape[1]
    rf = receptive_field(fw, dilations)
    dt = state["causal_layer.weight"].dtype
    if start_piece is None:
        start_piece = torch.zeros(1, Q, rf, dtype=dt)
        start_piece[:, Q // 2, :] = 1.0                                       # :159-160 (128 for Q=256)
    note, queues = start_piece, None
    out, all_logits = [], []
    for i in range(n_steps):
        if queues is None:
            logits, queues = gen_prime(state, dilations, note)
        else:
            logits, queues = gen_step(state, dilations, note, queues, queue_push)
        k = pick_greedy(logits) if uniforms is None else pick_sampled(logits, uniforms[i])
        out.append(k)
        if return_logits:
            all_logits.append(logits.clone())
        note = torch.zeros(1, Q, 1, dtype=dt)
        note[:, k, :] = 1.0                                                   # :170-172
    if return_logits:
        return out, torch.stack(all_logits)
    return out


# --------------------------------------------------------------------------------------
# a17-a20: autoencoder                                  wavenet_autoencoder/model1.py
# --------------------------------------------------------------------------------------
def ae_param_shapes(dilations, en_residual_channel, en_dilation_channel, en_bottleneck_width,
                    de_residual_channel, de_dilation_channel, de_skip_channel,
                    quantization_channel=256, use_bias=False, filter_width=2):
    """Registration order in the reference ctor: _init_encoding, _init_decoding,
    _init_causal_layer, _init_connection (model1.py:55-58)."""
    Q, fw = quantization_channel, filter_width
    Re, De, BW = en_residual_channel, en_dilation_channel, en_bottleneck_width
    Rd, Dd, Sd = de_residual_channel, de_dilation_channel, de_skip_channel
    out = []

    def add(name, shp):
        out.append((name + ".weight", shp))
        if use_bias:
            out.append((name + ".bias", (shp[0],)))
    n = len(dilations)
    for i in range(n):
        add(f"en_dilation_layer_stack.{i}", (De, Re, fw))
    for i in range(n):
        add(f"en_dense_layer_stack.{i}", (Re, De, 1))
    for i in range(n):
        add(f"de_dilation_layer_stack.{3 * i}", (2 * Dd, Rd, fw))
        add(f"de_dilation_layer_stack.{3 * i + 1}", (Rd, Dd, 1))
        add(f"de_dilation_layer_stack.{3 * i + 2}", (Sd, Dd, 1))
    add("en_causal_layer", (Re, Q, fw))
    add("bottleneck_layer", (BW, Re, 1))
    add("de_causal_layer", (Rd, Q, fw))
    add("connection_1", (Sd, Sd, 1))
    add("connection_2", (Q, Sd, 1))
    return out


def ae_cond_shapes(n_layers, en_bottleneck_width, de_dilation_channel, de_skip_channel):
    """The per-call throw-away conditioning convs, always bias=True (model1.py:178,216):
    n_layers of (2*Dd, BW, 1) then one (Sd, BW, 1)."""
    out = []
    for i in range(n_layers):
        out.append((f"cond.{i}.weight", (2 * de_dilation_channel, en_bottleneck_width, 1)))
        out.append((f"cond.{i}.bias", (2 * de_dilation_channel,)))
    out.append((f"cond.{n_layers}.weight", (de_skip_channel, en_bottleneck_width, 1)))
    out.append((f"cond.{n_layers}.bias", (de_skip_channel,)))
    return out


def init_state_from_shapes(shapes, seed=0) -> State:
    g = torch.Generator().manual_seed(seed)
    st: State = OrderedDict()
    last = 1
    for key, shp in shapes:
        if key.endswith(".weight"):
            last = shp[1] * shp[2]
        bound = 1.0 / (last ** 0.5)
        st[key] = (torch.rand(shp, generator=g) * 2 - 1) * bound
    return st


def ae_condition(x: torch.Tensor, enc: torch.Tensor) -> torch.Tensor:
    """`_conditon` (model1.py:227-247): broadcast when len % frames == 0, else TILE."""
    mb, ch, frames = enc.shape
    xl = x.shape[2]
    if xl % frames == 0:
        return (x.reshape(mb, ch, frames, -1) + enc.reshape(mb, ch, frames, 1)).reshape(mb, ch, xl)   # :233-240
    rep = xl // frames
    tiled = torch.cat((enc.repeat(1, 1, rep), enc[:, :, :xl % frames]), 2)                            # :241-245
    return x + tiled


def ae_encode(state: State, dilations, x: torch.Tensor, pool: int) -> torch.Tensor:
    """`_encode` (model1.py:137-156)."""
    s = F.conv1d(x, state["en_causal_layer.weight"], state.get("en_causal_layer.bias"))
    for i, d in enumerate(dilations):
        cur = s
        s = F.relu(s)
        s = F.conv1d(s, state[f"en_dilation_layer_stack.{i}.weight"],
                     state.get(f"en_dilation_layer_stack.{i}.bias"), dilation=d)
        s = F.relu(s)
        s = F.conv1d(s, state[f"en_dense_layer_stack.{i}.weight"], state.get(f"en_dense_layer_stack.{i}.bias"))
        s = s + cur[:, :, -s.shape[2]:]
    s = F.conv1d(s, state["bottleneck_layer.weight"], state.get("bottleneck_layer.bias"))
    return F.avg_pool1d(s, pool)


def ae_decode_logits(state: State, cond: State, dilations, x: torch.Tensor, enc: torch.Tensor, W: int):
    """`_decode` (model1.py:158-221) up to connection_2 (pre-softmax, (B,Q,W))."""
    cur = F.conv1d(x, state["de_causal_layer.weight"], state.get("de_causal_layer.bias"))
    skips = []
    for i, d in enumerate(dilations):
        j = 3 * i
        y = F.conv1d(cur, state[f"de_dilation_layer_stack.{j}.weight"],
                     state.get(f"de_dilation_layer_stack.{j}.bias"), dilation=d)          # :175
        en = F.conv1d(enc, cond[f"cond.{i}.weight"], cond[f"cond.{i}.bias"])              # :178-179
        y = ae_condition(y, en)                                                          # :183
        ch = y.shape[1]
        xg = y[:, :-(ch // 2), :]                                                        # :188 gate = FIRST half
        xf = y[:, -(ch // 2):, :]                                                        # :190 filter = second half
        z = torch.tanh(xf) * torch.sigmoid(xg)                                           # :192
        res = F.conv1d(z, state[f"de_dilation_layer_stack.{j + 1}.weight"],
                       state.get(f"de_dilation_layer_stack.{j + 1}.bias"))
        cur = cur[:, :, -res.shape[2]:] + res                                            # :198-202
        skips.append(F.conv1d(z[:, :, -W:], state[f"de_dilation_layer_stack.{j + 2}.weight"],
                              state.get(f"de_dilation_layer_stack.{j + 2}.bias")))       # :204-208
    r = F.relu(sum(skips))
    r = F.conv1d(r, state["connection_1.weight"], state.get("connection_1.bias"))        # :213
    n = len(dilations)
    en = F.conv1d(enc, cond[f"cond.{n}.weight"], cond[f"cond.{n}.bias"])                 # :216-217
    r = ae_condition(r, en)
    r = F.relu(r)
    return F.conv1d(r, state["connection_2.weight"], state.get("connection_2.bias"))     # :221


def ae_forward_logits(state: State, cond: State, dilations, x: torch.Tensor, pool: int):
    """`forward` (model1.py:256-268) up to the pre-softmax tensor."""
    fw = state["en_causal_layer.weight"].shape[2]
    W = x.shape[2] - receptive_field(fw, dilations) + 1
    enc = ae_encode(state, dilations, x, pool)
    return ae_decode_logits(state, cond, dilations, x, enc, W)


def ae_forward_probs(state: State, cond: State, dilations, x: torch.Tensor, pool: int):
    return scrambled_softmax(ae_forward_logits(state, cond, dilations, x, pool))


# --------------------------------------------------------------------------------------
# synthetic audio used by tests and bench
# --------------------------------------------------------------------------------------
def synthetic_audio(B: int, L: int, seed: int = 1234, sr: int = 16000) -> torch.Tensor:
    """0.2*sine mixture + N(0, 0.05^2) noise, clipped to [-1,1]; float32 (B,L)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(L, dtype=torch.float32) / sr
    freqs = torch.rand(B, 3, generator=g) * 800 + 100
    phase = torch.rand(B, 3, generator=g) * 6.2831853
    w = (0.2 * torch.sin(2 * 3.14159265 * freqs[:, :, None] * t[None, None, :] + phase[:, :, None])).sum(1) / 1.5
    w = w + 0.05 * torch.randn(B, L, generator=g)
    return w.clamp_(-1.0, 1.0)
