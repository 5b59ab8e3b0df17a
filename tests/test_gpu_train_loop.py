"""The reference training loop (wavenet/train.py:76-222) through the package's `train()`: JSON configs, loss / store logs in
the format `wavenet/vis/visualize.py:7-14` parses, checkpoint naming and rotation, resume from the newest checkpoint."""
import glob
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

DIL = [1, 2, 4, 8, 1, 2, 4, 8]


def _write_params(base, restore_model, epochs):
    os.makedirs(os.path.join(base, "params"), exist_ok=True)
    tp = {"restore_model": restore_model, "restore_dir": base + "/restore/", "log_dir": base + "/log/", "optimizer": "adam",
          "learning_rate": 1e-3, "momentum": 0.9, "num_epochs": epochs, "print_every": 2, "check_point_every": 1, "max_check_points": 2}
    wp = {"filter_width": 2, "dilations": DIL, "dilation_channels": 16, "residual_channels": 16, "skip_channels": 32,
          "quantization_channels": 256, "use_bias": False}
    for name, obj in (("train_params", tp), ("wavenet_params", wp), ("dataset_params", {})):
        with open(os.path.join(base, "params", name + ".json"), "w") as f:
            json.dump(obj, f)
    return base + "/params/"


def _loader(n_batches):
    rf = sum(DIL) + 2                                       # (filter_width - 1) * sum(dilations) + 2
    g = torch.Generator().manual_seed(0)
    W = 40
    out = []
    for _ in range(n_batches):
        idx = torch.randint(0, 256, (2, rf + W), generator=g)
        piece = torch.nn.functional.one_hot(idx[:, :-1], 256).permute(0, 2, 1).float()      # (B, Q, L) as the reference feeds it
        out.append({"audio_piece": piece, "audio_target": idx[:, rf:rf + W].contiguous()})
    return out


def test_train_loop_logs_checkpoints_and_resume(tmp_path):
    from music_b200.wavenet.train import train
    base = str(tmp_path)
    params = _write_params(base, "", epochs=3)
    net = train(base=params, dataloader=_loader(4))
    assert net is not None
    lines = open(base + "/log/loss_log.log").read().splitlines()
    assert len(lines) == 6                                    # 3 epochs x 4 batches / print_every 2
    assert [int(l.split(' ')[2]) for l in lines] == [2, 4, 6, 8, 10, 12]            # the resume parser of train.py:159-165
    losses = [float(l.split(' ')[-1]) for l in lines]                              # the parser of vis/visualize.py:7-14
    assert all(5.0 < x < 5.6 for x in losses)
    assert lines[0].startswith("Trained over 2 pieces,Average loss is ")
    store = open(base + "/log/store_log.log").read().splitlines()
    assert store == ["Epoch 1, model saved!", "Epoch 2, model saved!", "Epoch 3, model saved!"]
    models = sorted(os.path.basename(p) for p in glob.glob(base + "/restore/*.model"))
    assert models == ["wavenet2.model", "wavenet3.model"]     # max_check_points = 2: the oldest was rotated out
    # resume: epoch numbering and the piece counter continue
    params = _write_params(base, "wavenet3.model", epochs=1)
    train(base=params, dataloader=_loader(4))
    lines = open(base + "/log/loss_log.log").read().splitlines()
    assert [int(l.split(' ')[2]) for l in lines][-2:] == [14, 16]
    assert open(base + "/log/store_log.log").read().splitlines()[-1] == "Epoch 4, model saved!"
    models = sorted(os.path.basename(p) for p in glob.glob(base + "/restore/*.model"))
    assert models == ["wavenet3.model", "wavenet4.model"]


def test_batch_stager_orders_copies_and_slot_reuse():
    """Host batches staged on the copy stream arrive intact and in order while the compute stream is busy, a slot is
    refilled only after the step that read it (done()), and misuse is reported."""
    from music_b200 import _lib as L
    from music_b200.wavenet.train import BatchStager
    g = torch.Generator().manual_seed(3)
    host = [(torch.randint(0, 256, (4, 5000), generator=g).pin_memory(), torch.randint(0, 256, (4, 4000), generator=g).pin_memory())
            for _ in range(7)]
    st = BatchStager()
    with pytest.raises(L.WavenetB200Error):
        st.get()
    sums = []
    busy = torch.randn(2048, 2048, device="cuda")
    st.put(*host[0])
    for i in range(len(host)):
        p, t = st.get()
        if i + 1 < len(host):
            st.put(*host[i + 1])
        for _ in range(3):
            busy = busy @ busy * 1e-3                # keeps the compute stream behind the copy stream
        sums.append((p.sum() + (busy[0, 0] * 0).long(), t.sum(), p.clone(), t.clone()))
        st.done()
    torch.cuda.synchronize()
    for (ps, ts, pc, tc_), (hp, ht) in zip(sums, host):
        assert int(ps) == int(hp.sum()) and int(ts) == int(ht.sum())
        assert torch.equal(pc.cpu(), hp) and torch.equal(tc_.cpu(), ht)
    st.put(*host[0])
    st.put(*host[1])
    with pytest.raises(L.WavenetB200Error):
        st.put(*host[2])


@pytest.mark.gpu
def test_optimizer_state_checkpoint_resumes_exactly(tmp_path):
    """save_optimizer / load_optimizer next to the model checkpoint: a trainer restored from both continues like the
    one that kept running (Adam moments and step count included) - the reference checkpoints the model only (train.py:44-50)."""
    import torch
    from music_b200.wavenet.model import wavenet
    from music_b200.wavenet import train as T
    dil = [1, 2, 4, 8, 16, 32]
    torch.manual_seed(3)
    net = wavenet(2, dil, 64, 64, 256, 256, False, mode="bf16", parity="corrected").cuda()
    rf = net.receptive_field
    idx = torch.randint(0, 256, (2, rf + 300)).cuda()
    piece, target = idx[:, :-1].contiguous(), idx[:, rf:].contiguous()
    tr = T.Trainer(net, "adam", 1e-3, distributed=False)
    for _ in range(3):
        tr.step(piece, target)
    base = str(tmp_path) + "/"
    T.save_model(net, 7, base)
    T.save_optimizer(tr, 7, base)
    for _ in range(2):
        tr.step(piece, target)
    want = {k: v.detach().clone() for k, v in net.state_dict().items()}

    net2 = wavenet(2, dil, 64, 64, 256, 256, False, mode="bf16", parity="corrected")
    assert T.load_model(net2, base, "wavenet7.model") is net2
    net2 = net2.cuda()
    tr2 = T.Trainer(net2, "adam", 1e-3, distributed=False)
    assert T.load_optimizer(tr2, base, "wavenet7.model") is True
    assert T.load_optimizer(tr2, base, "wavenet9.model") is False
    assert tr2.step_count == 3
    for _ in range(2):
        tr2.step(piece, target)
    worst = max(float((v - want[k]).abs().max()) for k, v in net2.state_dict().items())
    # (a trainer restarted WITHOUT its moments moves every weight by ~lr = 1e-3 on its first step; fp32 atomics in the causal layer's
    #  gradient leave run-to-run noise far below that)
    assert worst < 2e-5, worst


def test_captured_step_equals_eager_step():
    """Trainer.capture: the step replayed from one CUDA graph (weight pack, forward, loss, backward, Adam with a device-side step count)
    reproduces the eager step on the deterministic bf16 path: same losses and parameters over 3 steps on two copies of a model, and the
    capture's warm-up leaves parameters, moments and the step count as they were."""
    import copy
    from music_b200.wavenet.model import wavenet
    from music_b200.wavenet import train as T
    dil = [1, 2, 4, 8, 16, 32]
    torch.manual_seed(4)
    net_a = wavenet(2, dil, 64, 64, 256, 256, False, mode="bf16", parity="corrected").cuda()
    net_b = copy.deepcopy(net_a)
    rf = net_a.receptive_field
    x1 = torch.randint(0, 256, (2, rf + 300)).cuda()
    x2 = torch.randint(0, 256, (2, rf + 300)).cuda()
    tr_a = T.Trainer(net_a, "adam", 1e-3, distributed=False)
    tr_b = T.Trainer(net_b, "adam", 1e-3, distributed=False)
    for tr in (tr_a, tr_b):
        tr.step(x1[:, :-1].contiguous(), x1[:, rf:].contiguous())
    before = {k: v.detach().clone() for k, v in net_b.state_dict().items()}
    assert tr_b.capture(x1[:, :-1].contiguous(), x1[:, rf:].contiguous()) is True
    assert tr_b.step_count == 1
    for k, v in net_b.state_dict().items():
        assert torch.equal(v, before[k]), k
    for x in (x2, x1, x2):
        la = tr_a.step(x[:, :-1].contiguous(), x[:, rf:].contiguous())
        lb = tr_b.step(x[:, :-1].contiguous(), x[:, rf:].contiguous())
        assert abs(float(la) - float(lb)) < 1e-5, (float(la), float(lb))
    worst = max(float((a - b).abs().max()) for (_, a), (_, b) in zip(net_a.state_dict().items(), net_b.state_dict().items()))
    assert worst < 2e-5, worst
