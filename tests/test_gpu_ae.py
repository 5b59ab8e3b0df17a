"""GPU parity: wavenet_autoencoder forward (model1.py) against golden vectors captured from the unmodified
reference (including its throw-away conditioning convs) and against the oracle. fp32 check mode: 1e-4."""
import numpy as np
import pytest
import torch

from oracle import wavenet_oracle as O
from tests.util import max_rel, state_of

pytestmark = pytest.mark.gpu


def _build(z):
    from music_b200.wavenet_autoencoder.model1 import wavenet_autoencoder
    dil = [int(d) for d in z["dilations"]]
    net = wavenet_autoencoder(2, int(z["Q"]), dil, int(z["Re"]), int(z["De"]), int(z["BW"]), int(z["pool"]), int(z["Rd"]),
                              int(z["Dd"]), int(z["Sd"]), bool(z["use_bias"]))
    st = state_of(z)
    assert list(net.state_dict().keys()) == list(st.keys())
    net.load_state_dict(st)
    cond = {k[len("cond."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("cond.")}
    return dil, st, cond, net.cuda()


@pytest.mark.parametrize("name", ["ae_tile", "ae_bias"])
def test_forward_vs_reference_golden(golden, name):
    z = golden(name)
    dil, st, cond, net = _build(z)
    idx = torch.from_numpy(z["idx"].astype(np.int64))
    with torch.no_grad():
        lg = net.forward_logits(indices=idx.cuda(), cond_weights=cond)
        probs = net(O.one_hot(idx, int(z["Q"])).cuda(), cond_weights=cond)
    assert max_rel(lg.cpu().numpy(), z["logits"]) < 1e-4
    assert max_rel(probs.cpu().numpy(), z["probs"]) < 1e-4
    assert probs.shape == z["probs"].shape


def test_broadcast_branch_and_encoding_vs_oracle():
    """W divisible by the frame count -> the view/broadcast branch of `_conditon`; also checks `_encode`."""
    from music_b200.wavenet_autoencoder.model1 import wavenet_autoencoder
    dil = [1, 2, 4, 8, 1, 2, 4, 8]
    cfg = dict(Q=256, Re=16, De=24, BW=32, pool=8, Rd=16, Dd=16, Sd=48)
    net = wavenet_autoencoder(2, 256, dil, cfg["Re"], cfg["De"], cfg["BW"], cfg["pool"], cfg["Rd"], cfg["Dd"], cfg["Sd"], True)
    st = {k: v.detach().clone() for k, v in net.state_dict().items()}
    cond = {}
    for i, c in enumerate(net.cond_layers):
        cond[f"cond.{i}.weight"] = c.weight.detach().clone()
        cond[f"cond.{i}.bias"] = c.bias.detach().clone()
    rf = O.receptive_field(2, dil)
    B, W = 2, 96                      # 12 frames, 96 % 12 == 0
    x = torch.randn(B, 256, rf + W - 1)
    ref = O.ae_forward_logits(st, cond, dil, x, cfg["pool"])
    enc_ref = O.ae_encode(st, dil, x, cfg["pool"])
    net = net.cuda()
    with torch.no_grad():
        lg, enc = net.forward_logits(wave_sample=x.cuda(), return_encoding=True)
    assert max_rel(enc.cpu().numpy(), enc_ref.numpy()) < 1e-4
    assert max_rel(lg.cpu().numpy(), ref.numpy()) < 1e-4
    with pytest.raises(ValueError, match="wave sample not long enough"):
        net(torch.zeros(1, 256, rf - 1).cuda())


def _ae_case(dil, cfg, bias, B, W, seed):
    from music_b200.wavenet_autoencoder.model1 import wavenet_autoencoder
    torch.manual_seed(seed)
    net = wavenet_autoencoder(2, 256, dil, cfg["Re"], cfg["De"], cfg["BW"], cfg["pool"], cfg["Rd"], cfg["Dd"], cfg["Sd"], bias)
    with torch.no_grad():
        for p in net.parameters():       # larger weights than the default init so that every ReLU / gate is exercised
            p.mul_(2.0)
    st = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    cond = {}
    for i, c in enumerate(net.cond_layers):
        cond[f"cond.{i}.weight"] = c.weight.detach().clone().requires_grad_(True)
        cond[f"cond.{i}.bias"] = c.bias.detach().clone().requires_grad_(True)
    rf = O.receptive_field(2, dil)
    idx = torch.randint(0, 256, (B, rf + W - 1))
    tgt = torch.randint(0, 256, (B * W,))
    return net, st, cond, idx, tgt


@pytest.mark.parametrize("W,bias,dense", [(96, True, True), (100, False, False), (61, True, False)])
def test_backward_vs_oracle_autograd(W, bias, dense):
    """loss.backward() through the module (wavenet_autoencoder/train.py:150-160: CrossEntropyLoss on the softmax output)
    against torch autograd on the oracle, fp32: every registered parameter and the conditioning convs.
    W=96: 12 frames, broadcast branch of `_conditon`; W=100 / 61: tiled branch, ragged tail after the last pool window."""
    dil = [1, 2, 4, 8, 1, 2, 4, 8]
    cfg = dict(Re=16, De=24, BW=32, pool=8, Rd=16, Dd=16, Sd=48)
    net, st, cond, idx, tgt = _ae_case(dil, cfg, bias, 2, W, seed=W)
    x = O.one_hot(idx, 256)
    if dense:
        x = x + 0.1 * torch.randn_like(x)
    probs = O.ae_forward_probs(st, cond, dil, x, cfg["pool"])
    loss_ref = torch.nn.functional.cross_entropy(probs, tgt)
    loss_ref.backward()

    net = net.cuda()
    for c in net.cond_layers:
        c.requires_grad_(True)
    if dense:
        out = net(x.cuda())
    else:
        out = net.forward_logits(indices=idx.cuda())
        from music_b200._engine import SoftmaxRowsFunction
        from music_b200 import _lib as L
        out = SoftmaxRowsFunction.apply(out, L.ROWS_REFERENCE)
    assert max_rel(out.detach().cpu().numpy(), probs.detach().numpy()) < 1e-4
    loss = torch.nn.functional.cross_entropy(out, tgt.cuda())
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) < 1e-5
    worst = 0.0
    for name, p in net.named_parameters():
        g_ref = st[name].grad
        if g_ref is None:                                   # the last block's dense conv feeds nothing (model1.py:194-202)
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        err = max_rel(p.grad.cpu().numpy(), g_ref.numpy())
        worst = max(worst, err)
        assert err < 2e-4, (name, err)
    for i, c in enumerate(net.cond_layers):
        for nm, t in (("weight", c.weight), ("bias", c.bias)):
            err = max_rel(t.grad.cpu().numpy(), cond[f"cond.{i}.{nm}"].grad.numpy())
            assert err < 2e-4, (i, nm, err)


def test_train_steps_reduce_loss(tmp_path):
    """A few iterations of the reference's AE training loop (wavenet_autoencoder/train.py:117-134) on one batch, through the
    package's own train module, then a checkpoint round trip with the reference's file naming."""
    from music_b200.wavenet_autoencoder import train as T
    dil = [1, 2, 4, 8, 1, 2, 4, 8]
    cfg = dict(Re=16, De=24, BW=32, pool=8, Rd=16, Dd=16, Sd=48)
    net, st, cond, idx, tgt = _ae_case(dil, cfg, False, 2, 96, seed=5)
    net = net.cuda()
    opt = T.get_optimizer(net, 'Adam', 1e-3)
    x = O.one_hot(idx, 256).cuda()
    rf = O.receptive_field(2, dil)
    target = idx[:, rf - 1:]                                 # predict the input itself: learnable
    losses = [float(T.train_step(net, opt, x, target)) for _ in range(12)]
    assert losses[-1] < losses[0] - 1e-3, losses
    T.save_model(net, 3, str(tmp_path) + "/")
    from music_b200.wavenet_autoencoder.model1 import wavenet_autoencoder
    net2 = wavenet_autoencoder(2, 256, dil, cfg["Re"], cfg["De"], cfg["BW"], cfg["pool"], cfg["Rd"], cfg["Dd"], cfg["Sd"], False)
    assert T.load_model(net2, str(tmp_path) + "/", "missing.model") is None
    assert T.load_model(net2, str(tmp_path) + "/", "wavenet_autoencoder3.model") is net2
    for (k, a), (_, b) in zip(net.state_dict().items(), net2.state_dict().items()):
        assert torch.equal(a.cpu(), b), k


def test_slow_generate_matches_oracle_greedy():
    """wavenet_autoencoder/generate.py:13-19: greedy pick over the last row of the scrambled softmax, sliding window."""
    from music_b200.wavenet_autoencoder.generate import generate_codes
    dil = [1, 2, 4, 1, 2, 4]
    cfg = dict(Re=16, De=16, BW=16, pool=8, Rd=16, Dd=16, Sd=32)
    net, st, cond, idx, tgt = _ae_case(dil, cfg, True, 1, 16, seed=11)
    rf = O.receptive_field(2, dil)
    start = O.one_hot(torch.randint(0, 256, (1, rf + 40)), 256)
    st = {k: v.detach() for k, v in st.items()}
    cond = {k: v.detach() for k, v in cond.items()}
    wav, ref = start.clone(), []
    for _ in range(5):
        out = O.ae_forward_probs(st, cond, dil, wav, cfg["pool"]).view(-1, 256)
        p = int(torch.topk(out[-1], 1)[1])
        ref.append(p)
        wav = torch.cat((wav[:, :, -rf - 511:], O.one_hot(torch.tensor([[p]]), 256)), 2)
    got = generate_codes(net.cuda(), 5, start_piece=start)
    assert got == ref


@pytest.mark.parametrize("W,pool,bias", [(24, 4, True), (26, 4, False)])
def test_incremental_decoder_generation_equals_full_forward(W, pool, bias):
    """Extension: incremental generation with the conditioned decoder.  Teacher-forced with the samples of a
    given sequence, step j must reproduce row j of the decoder's full forward over that sequence (`_decode`,
    model1.py:158-225), including `_conditon`'s per-layer frame rule (:227-247), whose branch differs from layer to layer
    (W=24: 6 frames; W=26: 6 frames and a ragged tail)."""
    from music_b200.wavenet_autoencoder.generate import fast_generate_codes
    dil = [1, 2, 4, 1, 2, 4]
    cfg = dict(Re=16, De=16, BW=16, pool=pool, Rd=16, Dd=16, Sd=32)
    net, st, cond, idx, tgt = _ae_case(dil, cfg, bias, 2, W, seed=W)
    st = {k: v.detach() for k, v in st.items()}
    cond = {k: v.detach() for k, v in cond.items()}
    rf = O.receptive_field(2, dil)
    L = rf + W - 1
    frames = W // pool
    enc = torch.randn(2, cfg["BW"], frames)
    x = O.one_hot(idx, 256)
    ref = O.ae_decode_logits(st, cond, dil, x, enc, W)                       # (B, Q, W)
    net = net.cuda()
    forced = idx[:, rf:].t().contiguous()                                    # sample rf + j feeds step j + 1
    codes, logits = fast_generate_codes(net, enc, L, W, idx[:, :rf], cond_weights=cond, forced=forced, return_logits=True)
    got = logits.permute(1, 2, 0).cpu().numpy()                              # (B, Q, W)
    assert max_rel(got, ref.numpy()) < 1e-4
    assert codes.t().cpu().tolist() == ref.argmax(dim=1).tolist()
    # free-running generation (no forcing) is self-consistent: replaying its own picks as forced inputs gives the same codes
    free = fast_generate_codes(net, enc, L, 8, idx[:, :rf], cond_weights=cond)
    again = fast_generate_codes(net, enc, L, 8, idx[:, :rf], cond_weights=cond, forced=free[:-1])
    assert torch.equal(free, again)


@pytest.mark.parametrize("n_layers,W,bias", [(9, 24, False), (12, 26, True), (4, 24, False)])
def test_half_precision_pipeline_decoder_generation_vs_fp32_kernel(n_layers, W, bias):
    """mode="bf16" of the incremental decoder generation: the conditioned 64/64/256/256 decoder on the weights-stationary cluster
    pipeline (fp16 weight fragments; the per-frame conditioning of every block is folded into the W0.old term that is computed
    ahead of the token, post_process_1's is fetched by the head while the token travels).  Teacher-forced with the same samples,
    its logits stay within 1e-2 relative of the fp32 kernel's (which equals the full forward, test above) at every step, in both
    frame-rule branches (W = 24: 6 frames; 26: ragged) and for layer counts that produce every CTA-role split; free-running
    generation in ONE launch equals the step-by-step replay of its own picks."""
    from music_b200.wavenet_autoencoder.generate import fast_generate_codes
    dil = [2 ** (i % 4) for i in range(n_layers)]
    cfg = dict(Re=16, De=16, BW=16, pool=4, Rd=64, Dd=64, Sd=256)
    B = 3
    net, st, cond, idx, tgt = _ae_case(dil, cfg, bias, B, W, seed=W + n_layers)
    cond = {k: v.detach() for k, v in cond.items()}
    rf = O.receptive_field(2, dil)
    L = rf + W - 1
    enc = torch.randn(B, cfg["BW"], W // cfg["pool"])
    net = net.cuda()
    forced = idx[:, rf:].t().contiguous()
    c32, l32 = fast_generate_codes(net, enc, L, W, idx[:, :rf], cond_weights=cond, forced=forced, return_logits=True)
    c16, l16 = fast_generate_codes(net, enc, L, W, idx[:, :rf], cond_weights=cond, forced=forced, return_logits=True, mode="bf16")
    for j in range(W):
        assert max_rel(l16[j].cpu().numpy(), l32[j].cpu().numpy()) < 1e-2, j
    assert (c16 == c32).float().mean().item() >= 0.9
    free = fast_generate_codes(net, enc, L, 12, idx[:, :rf], cond_weights=cond, mode="bf16")
    again = fast_generate_codes(net, enc, L, 12, idx[:, :rf], cond_weights=cond, forced=free[:-1], mode="bf16")
    assert torch.equal(free, again)
    u = torch.rand(12, B)
    sampled = fast_generate_codes(net, enc, L, 12, idx[:, :rf], cond_weights=cond, uniforms=u, mode="bf16")
    replay = fast_generate_codes(net, enc, L, 12, idx[:, :rf], cond_weights=cond, uniforms=u, forced=sampled[:-1], mode="bf16")
    assert torch.equal(sampled, replay)


# ------------------------------------------------------------------------------------------ bf16 tensor-core decoder
@pytest.mark.parametrize("W,Sd,dense", [(96, 512, False), (100, 256, False), (61, 512, True)])
def test_bf16_decoder_forward_backward_vs_oracle(W, Sd, dense):
    """mode="bf16": the conditioned decoder (model1.py:158-247) on the tcgen05 WaveNet kernels - 32-channel stacks zero-padded to
    64, per-frame conditioning added in the block kernels' epilogues (both `_conditon` branches: W = 96 divides into 12 frames,
    100 / 61 tile the encoding), the S = 512 head as streaming GEMMs - and the encoder in fp32.  Logits 1e-2 relative (north
    star's bf16 bar); gradients against torch autograd on the oracle: the decoder's tensors within the bf16 bound of
    tests/test_gpu_fast.py (ordinary weights: ReLU mask flips included), the encoder's - which receive their gradient through
    the bf16 decoder's conditioning sums - likewise; cosine of the whole gradient > 0.99."""
    dil = [1, 2, 4, 8, 16, 1, 2, 4, 8, 16]
    cfg = dict(Re=32, De=32, BW=64, pool=8, Rd=32, Dd=32, Sd=Sd)
    from music_b200.wavenet_autoencoder.model1 import wavenet_autoencoder
    torch.manual_seed(W)
    net = wavenet_autoencoder(2, 256, dil, cfg["Re"], cfg["De"], cfg["BW"], cfg["pool"], cfg["Rd"], cfg["Dd"], cfg["Sd"], False, mode="auto")
    assert net.mode == "bf16"
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(1.5)
    st = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    cond = {}
    for i, c in enumerate(net.cond_layers):
        cond[f"cond.{i}.weight"] = c.weight.detach().clone().requires_grad_(True)
        cond[f"cond.{i}.bias"] = c.bias.detach().clone().requires_grad_(True)
    rf = O.receptive_field(2, dil)
    B = 2
    idx = torch.randint(0, 256, (B, rf + W - 1))
    tgt = torch.randint(0, 256, (B * W,))
    x = O.one_hot(idx, 256)
    if dense:
        x = x + 0.1 * torch.randn_like(x)
    logits_ref = O.ae_forward_logits(st, cond, dil, x, cfg["pool"])
    probs_ref = O.scrambled_softmax(logits_ref)
    loss_ref = torch.nn.functional.cross_entropy(probs_ref, tgt)
    loss_ref.backward()

    net = net.cuda()
    for c in net.cond_layers:
        c.requires_grad_(True)
    lg = net.forward_logits(wave_sample=x.cuda()) if dense else net.forward_logits(indices=idx.cuda())
    e = max_rel(lg.detach().cpu().numpy(), logits_ref.detach().numpy())
    print(f"AE bf16 decoder W={W} Sd={Sd}: logits max-rel {e:.3e}")
    assert e < 1e-2
    from music_b200._engine import SoftmaxRowsFunction
    from music_b200 import _lib as L
    out = SoftmaxRowsFunction.apply(lg, L.ROWS_REFERENCE)
    loss = torch.nn.functional.cross_entropy(out, tgt.cuda())
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) < 1e-4
    ga, gb, worst = [], [], (0.0, "")
    for name, p in net.named_parameters():
        g_ref = st[name].grad
        if g_ref is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        a, b = p.grad.detach().cpu().double().reshape(-1), g_ref.double().reshape(-1)
        err = float((a - b).norm() / (b.norm() + 1e-30))
        worst = max(worst, (err, name))
        ga.append(a)
        gb.append(b)
    for i, c in enumerate(net.cond_layers):
        for nm, t in (("weight", c.weight), ("bias", c.bias)):
            a, b = t.grad.detach().cpu().double().reshape(-1), cond[f"cond.{i}.{nm}"].grad.double().reshape(-1)
            err = float((a - b).norm() / (b.norm() + 1e-30))
            worst = max(worst, (err, f"cond.{i}.{nm}"))
    ga, gb = torch.cat(ga), torch.cat(gb)
    cos = float((ga @ gb) / (ga.norm() * gb.norm()))
    print(f"AE bf16 decoder W={W} Sd={Sd}: worst grad rel-l2 {worst}, cosine {cos:.5f}")
    assert worst[0] < 0.15, worst
    assert cos > 0.99


def test_bf16_decoder_at_shipped_shape_tracks_fp32_mode():
    """The shipped model_params.json shape (40 layers, 32 channels, bottleneck / skip 512, pool 512) on one clip of 8192 targets:
    the bf16-decoder mode against the library's own fp32 check mode (oracle-pinned above) - logits 1e-2, loss 1e-4."""
    from music_b200.wavenet_autoencoder.model1 import wavenet_autoencoder
    dil = [2 ** i for i in range(10)] * 4
    torch.manual_seed(0)
    net16 = wavenet_autoencoder(2, 256, dil, 32, 32, 512, 512, 32, 32, 512, False, mode="bf16").cuda()
    net32 = wavenet_autoencoder(2, 256, dil, 32, 32, 512, 512, 32, 32, 512, False, mode="fp32").cuda()
    net32.load_state_dict(net16.state_dict())
    cond = {}
    for i, c in enumerate(net16.cond_layers):
        cond[f"cond.{i}.weight"] = c.weight.detach().clone()
        cond[f"cond.{i}.bias"] = c.bias.detach().clone()
    W = 8192
    idx = torch.randint(0, 256, (1, net16.receptive_field + W - 1)).cuda()
    with torch.no_grad():
        a = net16.forward_logits(indices=idx, cond_weights=cond)
        b = net32.forward_logits(indices=idx, cond_weights=cond)
    e = max_rel(a.cpu().numpy(), b.cpu().numpy())
    print("AE shipped shape bf16 decoder vs fp32 mode: logits max-rel", e)
    assert e < 1e-2


def test_fused_trainer_matches_autograd_train_step():
    """AeTrainer.step (flat parameters, fused loss / optimizer, no autograd graph) against train_step (autograd + torch.optim.Adam) on the
    same model, batch and conditioning convs: same loss trajectory and parameters after 3 SGD steps.  Shape with the mma.sync
    encoder and the tcgen05 decoder (mode bf16), index input."""
    import copy
    from music_b200.wavenet_autoencoder import train as T
    from music_b200.wavenet_autoencoder.model1 import wavenet_autoencoder
    dil = [1, 2, 4, 8, 16, 1, 2, 4, 8, 16]
    torch.manual_seed(11)
    net_a = wavenet_autoencoder(2, 256, dil, 32, 32, 64, 8, 32, 32, 256, False, mode="auto").cuda()
    assert net_a.mode == "bf16"
    net_b = copy.deepcopy(net_a)
    rf = net_a.receptive_field
    idx = torch.randint(0, 256, (2, rf + 96 - 1)).cuda()
    target = idx[:, rf - 1:].contiguous()
    # plain SGD: Adam would turn the run-to-run noise of a near-zero gradient (fp32 atomics in the weight-gradient reductions) into
    # full-size updates, which says nothing about the two step implementations
    opt = T.get_optimizer(net_a, 'sgd', 0.5, 0.0)
    tr = T.AeTrainer(net_b, 'sgd', 0.5, momentum=0.0, distributed=False)
    from music_b200._engine import SoftmaxRowsFunction
    from music_b200 import _lib as L
    for it in range(3):
        opt.zero_grad()
        probs = SoftmaxRowsFunction.apply(net_a.forward_logits(indices=idx), L.ROWS_REFERENCE)
        la = torch.nn.functional.cross_entropy(probs, target.reshape(-1))
        la.backward()
        opt.step()
        lb = tr.step(idx, target)
        assert abs(float(la) - float(lb)) < 1e-5, (it, float(la), float(lb))
    for (k, a), (_, b) in zip(net_a.state_dict().items(), net_b.state_dict().items()):
        assert float((a - b).norm()) <= 1e-4 * float(a.norm()) + 1e-7, (k, float((a - b).norm()), float(a.norm()))
    assert all(p.grad is not None for p in net_b.parameters())


def test_captured_step_matches_eager_step():
    """AeTrainer.capture: the whole step replayed from ONE CUDA graph (device-side Adam step count) against the eager step - same losses
    over 4 steps on two copies of one model, optimizer state untouched by the capture's warm-up (step_count, moments)."""
    import copy
    from music_b200.wavenet_autoencoder import train as T
    from music_b200.wavenet_autoencoder.model1 import wavenet_autoencoder
    dil = [1, 2, 4, 8, 16, 1, 2, 4, 8, 16]
    torch.manual_seed(21)
    net_a = wavenet_autoencoder(2, 256, dil, 32, 32, 64, 8, 32, 32, 256, False, mode="auto").cuda()
    net_b = copy.deepcopy(net_a)
    rf = net_a.receptive_field
    idx = torch.randint(0, 256, (2, rf + 96 - 1)).cuda()
    idx2 = torch.randint(0, 256, (2, rf + 96 - 1)).cuda()
    tr_a = T.AeTrainer(net_a, 'Adam', 1e-4, distributed=False)
    tr_b = T.AeTrainer(net_b, 'Adam', 1e-4, distributed=False)
    tr_a.step(idx, idx[:, rf - 1:].contiguous())
    tr_b.step(idx, idx[:, rf - 1:].contiguous())
    before = {k: v.detach().clone() for k, v in net_b.state_dict().items()}
    assert tr_b.capture(idx, idx[:, rf - 1:].contiguous()) is True
    assert tr_b.step_count == 1
    for k, v in net_b.state_dict().items():
        assert torch.equal(v, before[k]), k                      # the capture's warm-up steps ran on a snapshot
    for batch in (idx2, idx, idx2):
        la = tr_a.step(batch, batch[:, rf - 1:].contiguous())
        lb = tr_b.step(batch, batch[:, rf - 1:].contiguous())
        assert abs(float(la) - float(lb)) < 5e-5, (float(la), float(lb))      # (two models drifting apart by atomic-order noise under Adam)
    assert tr_b.step_count == tr_a.step_count == 4
    worst = max(float((a - b).abs().max()) for (_, a), (_, b) in zip(net_a.state_dict().items(), net_b.state_dict().items()))
    # Adam moves a weight by up to ~3 lr per step when its gradient is at noise level (the reference objective's gradients are ~1e-3 of
    # a proper cross entropy's), and the fp32-atomic summation order differs between any two runs: the bound only says
    # "no step was lost or doubled" (a wrong step count or a stale moment buffer shows up in the losses above)
    assert worst < 3e-3, worst
