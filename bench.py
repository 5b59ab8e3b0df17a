#!/usr/bin/env python
"""bench.py - headline benchmark of the WaveNet hot path (BASELINE.json metric: training audio samples/sec).

    python bench.py --gpus N --steps K --warmup W            # this repo (bf16 tensor-core path)
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port), host cores

Workload (N = 1): BASELINE.json configs[1] - WaveNet 30 layers (dilations 1..512 x 3), 64 residual /
64 dilation / 256 skip channels, batch 16 x 16k-sample windows, bf16 compute / fp32 master weights, Adam.
A "step" = zero_grad -> forward -> CrossEntropyLoss(probabilities) -> backward -> [all-reduce] -> Adam
(wavenet/train.py:171-182) on one synthetic batch.  For N > 1 (torchrun) every rank keeps batch 16
(weak scaling; 8 GPUs = the global batch 128 of configs[2]) and the flat gradient is all-reduced over NCCL.

One JSON line on rank 0; see the driver contract in the task statement for the keys.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIL = [2 ** i for i in range(10)] * 3
R = D = 64
S = Q = 256
WINDOW = 16000
BATCH = 16


def flops_per_sample():
    """Algorithmic forward MACs per target sample, x2 FLOPs, x3 for training."""
    rf = sum(DIL) + 2
    L = rf + WINDOW - 1
    macs = (L - 1) * 2 * Q * R
    ln = L - 1
    for d in DIL:
        ln -= d
        macs += ln * (2 * R * 2 * D + D * R) + WINDOW * D * S
    macs += WINDOW * S * S + WINDOW * S * Q
    return 2.0 * macs / WINDOW


def kernel_flops(B):
    """Algorithmic FLOPs per step attributed to each kernel name of the profiler (recompute not counted)."""
    rf = sum(DIL) + 2
    L = rf + WINDOW - 1
    W, N = WINDOW, len(DIL)
    lens, ln = [], L - 1
    for d in DIL:
        ln -= d
        lens.append(ln)
    f = {}
    f["block_fwd"] = sum(2.0 * B * l * (2 * R * 2 * D + (D * R if i + 1 < N else 0)) for i, l in enumerate(lens))
    f["skip_head"] = 2.0 * B * W * (N * D * S + S * S + S * Q)
    f["block_bwd"] = sum(2.0 * B * l * D * R for l in lens[:-1])
    f["gemm_nt_dx"] = sum(2.0 * B * l * 2 * R * 2 * D for l in lens)
    f["gemm_tn_dWfg"] = f["gemm_nt_dx"]
    f["gemm_tn_dWd"] = f["block_bwd"]
    f["block_bwd2"] = f["block_bwd"] + f["gemm_tn_dWfg"] + f["gemm_tn_dWd"]      # dz dgrad + fused weight gradients
    f["block_bwd3"] = f["block_bwd2"]
    f["block_bwd6"] = f["block_bwd2"] + f["gemm_nt_dx"]                          # + the fused data-gradient GEMM of the dilated conv
    f["gemm_nt_dZcat"] = 2.0 * B * W * N * D * S
    f["gemm_tn_dWs"] = f["gemm_nt_dZcat"]
    f["gemm_nt_dH1"] = 2.0 * B * W * S * Q
    f["gemm_nt_dSK"] = 2.0 * B * W * S * S
    f["gemm_tn_head"] = 2.0 * B * W * (S * Q + S * S)
    return f


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed ncu captures
    (profiles/r2_traffic.json, made by tools/traffic_json.py; bytes cannot be measured live outside a profiler).
    None when not captured."""
    for name in ("r2_traffic.json", "r1d_traffic.json", "r1c_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))
            return t["kernels"][kernel]["dram_bytes_per_launch"]
        except Exception:
            continue
    return None


class ClockSampler:
    """SM clock / power / throttle reasons polled through NVML every 20 ms while the timed region runs
    (nvidia-smi's shortest loop interval is too coarse for a ~100 ms region)."""

    def __init__(self, gpu_index):
        self.gpu, self.samples, self.stop_flag, self.ok = gpu_index, [], False, False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.ok = True
        except Exception:
            return
        self.t = threading.Thread(target=self._poll, daemon=True)
        self.t.start()

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, pw, rs))
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.stop_flag = True
        self.t.join(timeout=1)
        nv = self.nv
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        reasons = sorted(n for n, bit in names.items() if any(r & bit for _, _, r in self.samples))
        # under load = the upper half of the power samples (the region is bracketed by idle syncs)
        sm = sorted(s for s, _, _ in self.samples)
        pw = [p for _, p, _ in self.samples]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": reasons}


def cpu_reference_rate(steps, warmup, n_threads=None):
    """The reference's CPU path: oracle port of wavenet/model.py forward + the train step of
    wavenet/train.py:171-182 (torch CPU fp32, all host threads) on ONE clip of the cfg-2 shape
    (batch 1 x 16000 targets; the CPU rate is batch-insensitive).  Returns (samples/s, cores, s/step)."""
    import torch
    from oracle import wavenet_oracle as O
    n_threads = n_threads or os.cpu_count()
    torch.set_num_threads(n_threads)
    st = O.init_wavenet_state(DIL, D, R, S, Q, False, seed=0)
    rf = O.receptive_field(2, DIL)
    L = rf + WINDOW - 1
    idx = O.mu_law_encode(O.synthetic_audio(1, L + 1, seed=1234), Q)
    x = O.one_hot(idx[:, :L], Q)
    tgt = idx[:, rf:rf + WINDOW].contiguous()
    ts = O.TrainState(st, "adam", lr=1e-4)
    for _ in range(warmup):
        O.train_step(ts, DIL, x, tgt)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.train_step(ts, DIL, x, tgt)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return WINDOW / dt, n_threads, dt


def measure_l2_read_peak(dev, n_ctas):
    """Measured L2 -> SM read bandwidth (GB/s) with `n_ctas` CTAs each streaming one L2-resident 2.5 MB buffer - the access
    pattern of the generation kernel's CTAs on their shared weight image (csrc/bench_kernels.cu)."""
    import torch
    from music_b200 import _lib as L
    lib = L.load()
    nbytes = 2_621_440
    buf = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    sink = torch.zeros(4, dtype=torch.int32, device=dev)
    iters = 200
    L.check(lib.wn_bench_l2_read(L.ptr(buf), nbytes, n_ctas, 5, L.ptr(sink), L.stream_ptr()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(lib.wn_bench_l2_read(L.ptr(buf), nbytes, n_ctas, iters, L.ptr(sink), L.stream_ptr()))
    e1.record()
    torch.cuda.synchronize()
    return nbytes * n_ctas * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9


def bench_generation(net, n_streams, n_steps, dev):
    """BASELINE.json configs[3]: fast_generate incremental sampling, 30-layer model, 64 parallel streams x 160 000 samples.
    Prime every stream with one-hot(128) x rf (fast_generate.py:158-161; timed separately), then time n_steps greedy steps
    per stream on the device (CUDA events).
    `roofline` follows SURVEY.md 8(d): ALGORITHMIC bytes per step = streams x 15 368 B (ring read + write of 30 x 64 fp32 per
    stream-step + 8 B of I/O) + the 2.54 MB weight set once, over the step time, against the measured HBM copy peak.
    Stream counts whose groups of 8 all fit into resident clusters run on the weights-stationary cluster pipeline (gen_pipe_kernel;
    `kernel` says which kernel and geometry served the run: gen_pipe4 = 4 blocks per CTA, one group per 10-CTA cluster, up to 128
    streams; gen_pipe = 2 blocks per CTA, up to 8 groups per 16-CTA cluster); its only per-step L2 traffic is the ring state.
    Larger stream counts run on the one-CTA-per-8-streams kernel, whose L2 -> SM traffic (every CTA streams the weight-fragment image each step) is reported in `l2` against an L2 read
    peak measured in this run with the same number of CTAs.
    Also timed: a throughput configuration with 8 streams on every SM-sized slice of the GPU (1024 streams)."""
    import torch
    from music_b200 import _lib as L_
    from music_b200.wavenet import fast_generate as FG
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        prime = torch.full((n_streams, net.receptive_field), Q // 2, dtype=torch.int64, device=dev)
        FG._prime(net, prime)                                  # warm-up (workspace allocation, packing)
        torch.cuda.synchronize()
        e0.record()
        first, state, _ = FG._prime(net, prime)
        e1.record()
        torch.cuda.synchronize()
        ms_prime = e0.elapsed_time(e1)
        FG._steps(net, state, first, 50)                       # warm-up
        torch.cuda.synchronize()
        e0.record()
        codes, _ = FG._steps(net, state, first, n_steps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        del codes
        # throughput configuration: 128 CTAs x 8 streams
        big = 1024
        first_b, state_b, _ = FG._prime(net, torch.full((big, net.receptive_field), Q // 2, dtype=torch.int64, device=dev))
        FG._steps(net, state_b, first_b, 20)
        torch.cuda.synchronize()
        nb = max(50, min(2000, n_steps // 10))
        e0.record()
        FG._steps(net, state_b, first_b, nb)
        e1.record()
        torch.cuda.synchronize()
        ms_big = e0.elapsed_time(e1)
    def kernel_used(state_, first_):      # which generation kernel serves this stream count (profiler label of a short call)
        lib_ = L_.load()
        lib_.wn_profile_enable(1)
        FG._steps(net, state_, first_, 4)
        torch.cuda.synchronize()
        rep = L_.profile_report()
        lib_.wn_profile_enable(0)
        names = [n for n, c, m in rep if n.startswith("gen_")]
        return names[0] if names else "?"
    with torch.no_grad():
        k_small, k_big = kernel_used(state, first), kernel_used(state_b, first_b)
    n_layers = len(DIL)
    w_bytes = 2 * (n_layers * (2 * 64 * 128 + 64 * 64 + 64 * 256) + 2 * 256 * 256) + 4 * 2 * Q * R      # what a CTA streams per step
    w_algo = 2 * 1_269_760                                    # the parameter set in half precision, once per step
    ring_algo = n_layers * 64 * 4 * 2 + 8                     # per stream-step: one fp32 vector read + written per block, + I/O

    def algo_rate(streams, steps, t_ms):
        return steps * (streams * ring_algo + w_algo) / (t_ms * 1e-3) / 1e9

    def l2_rate(streams, steps, t_ms):
        return steps * (((streams + 7) // 8) * w_bytes + streams * (ring_algo - 8)) / (t_ms * 1e-3) / 1e9
    peak = peaks.get("hbm_gbs", 6650.0)
    n_ctas = (n_streams + 7) // 8
    l2_peak, l2_peak_big = measure_l2_read_peak(dev, n_ctas), measure_l2_read_peak(dev, big // 8)
    gbs = algo_rate(n_streams, n_steps, ms)
    many = {"streams": big, "steps": nb, "kernel": k_big, "us_per_step": ms_big * 1e3 / nb, "samples_per_s_per_stream": nb / (ms_big * 1e-3),
            "samples_per_s_total": big * nb / (ms_big * 1e-3),
            "roofline": {"bound": "hbm", "achieved": algo_rate(big, nb, ms_big), "peak": peak, "unit": "GB/s",
                         "frac": algo_rate(big, nb, ms_big) / peak},
            "l2": {"achieved_gbs": l2_rate(big, nb, ms_big), "measured_peak_gbs": l2_peak_big,
                   "frac": l2_rate(big, nb, ms_big) / l2_peak_big, "ctas": big // 8}}
    return {"workload": f"fast_generate incremental sampling, 30-layer 64/64/256 model, {n_streams} streams x {n_steps} steps "
                        "(greedy, queue_push=output as the reference)",
            "samples_per_s_per_stream": n_steps / (ms * 1e-3), "samples_per_s_total": n_streams * n_steps / (ms * 1e-3),
            "us_per_step": ms * 1e3 / n_steps, "prime_ms": ms_prime,
            "dtype": "f16 weights and MMA operands, f32 accumulate / ring state",
            "many_streams": many,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None,
                         "bytes_per_step": n_streams * ring_algo + w_algo,
                         "note": "SURVEY.md 8(d) accounting: algorithmic bytes per step over the step time.  The recurrence is "
                                 "a 31-stage dependency chain per stream (latency-bound at 64 streams): this fraction measures "
                                 "how far the kernel is from streaming its state and weights once, not its L2 efficiency"},
            "kernel": k_small,
            "l2": ({"achieved_gbs": l2_rate(n_streams, n_steps, ms), "measured_peak_gbs": l2_peak, "frac": l2_rate(n_streams, n_steps, ms) / l2_peak,
                    "ctas": n_ctas, "note": "bytes the kernel requests from L2 (weight-fragment image per CTA of 8 streams per step + "
                                            "ring vectors) against the L2 read rate the same number of CTAs reach in wn_bench_l2_read"}
                   if not k_small.startswith("gen_pipe") else
                   {"achieved_gbs": n_steps * n_streams * (ring_algo - 8) / (ms * 1e-3) / 1e9, "measured_peak_gbs": None, "frac": None,
                    "ctas": ((n_streams + 7) // 8) * (10 if k_small == "gen_pipe4" else n_layers // 2 + 1),
                    "note": "gen_pipe: the weights are resident in the registers and shared memory of a cluster (10 CTAs per group of 8 "
                            "streams in the 4-blocks-per-CTA geometry), so the only per-step L2 traffic is the ring vectors (one fp32 "
                            "vector read and written per block and stream) = the algorithmic state bytes; tokens move CTA to CTA "
                            "through distributed shared memory (3 KB st.async per hop and group, 8 KB of skip sums behind it).  The "
                            "step is the ring latency (30 blocks of ~480 cycles + 9 hops of ~280 + head ~2150 cycles), not a bandwidth"})}


def bench_gpu_incumbent(dev, B, steps=3):
    """What a user of the reference gets on this GPU without this library: the same model built from torch.nn.Conv1d modules
    (cuDNN / cuBLAS kernels, torch eager autograd, torch.optim.Adam), dense one-hot input generated on the device, the
    reference's forward + scrambled softmax + CrossEntropyLoss(probabilities) step (wavenet/model.py:86-145,
    train.py:171-182) - in fp32 and under bf16 autocast.  Same shape as the headline (B x 16000 targets), CUDA-event timed."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    class EagerWaveNet(nn.Module):
        def __init__(self):
            super().__init__()
            self.causal = nn.Conv1d(Q, R, 2, bias=False)
            self.filt = nn.ModuleList(nn.Conv1d(R, D, 2, dilation=d, bias=False) for d in DIL)
            self.gate = nn.ModuleList(nn.Conv1d(R, D, 2, dilation=d, bias=False) for d in DIL)
            self.dense = nn.ModuleList(nn.Conv1d(D, R, 1, bias=False) for _ in DIL)
            self.skip = nn.ModuleList(nn.Conv1d(D, S, 1, bias=False) for _ in DIL)
            self.p1 = nn.Conv1d(S, S, 1, bias=False)
            self.p2 = nn.Conv1d(S, Q, 1, bias=False)

        def forward(self, x):
            W = x.shape[2] - (sum(DIL) + 2) + 1
            cur = self.causal(x)
            total = None
            for i in range(len(DIL)):
                z = torch.sigmoid(self.gate[i](cur)) * torch.tanh(self.filt[i](cur))
                dense = self.dense[i](z)
                cur = dense + cur[:, :, -dense.shape[2]:]
                sk = self.skip[i](z[:, :, -W:])
                total = sk if total is None else total + sk
            out = self.p2(F.relu(self.p1(F.relu(total))))
            return torch.softmax(out.contiguous().view(-1, Q).float(), dim=1)

    rf = sum(DIL) + 2
    Lx = rf + WINDOW - 1
    g = torch.Generator(device=dev).manual_seed(7)
    idx = torch.randint(0, Q, (B, Lx + 1), generator=g, device=dev)
    x = F.one_hot(idx[:, :Lx], Q).permute(0, 2, 1).float().contiguous()
    tgt = idx[:, rf:rf + WINDOW].reshape(-1)
    out = {}
    for name, autocast in (("fp32", False), ("bf16_autocast", True)):
        torch.manual_seed(0)
        net = EagerWaveNet().to(dev)
        opt = torch.optim.Adam(net.parameters(), lr=1e-4)

        def step():
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                probs = net(x)
            loss = F.cross_entropy(probs, tgt)
            loss.backward()
            opt.step()
            return loss
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[name] = {"ms_per_step": ms, "samples_per_s": B * WINDOW / (ms * 1e-3), "loss": float(loss)}
        del net, opt
        torch.cuda.empty_cache()
    out["workload"] = (f"torch eager (nn.Conv1d -> cuDNN/cuBLAS, autograd, torch.optim.Adam) on the same B200, same model and "
                       f"shape (batch {B} x 16000 targets, dense one-hot input resident on the device), {steps} steps")
    out["tf32"] = bool(torch.backends.cudnn.allow_tf32)
    return out


def bench_cfg1(dev, cpu_seconds=6.0):
    """BASELINE.json configs[0]: the reference's default shape (10 x 3 layers, 32 residual / 32 dilation / 256 skip channels)
    on one 16000-sample clip (W = 12930 targets), batch 1, Adam - on the GPU through the package (mode='auto': the tcgen05
    kernels with channels zero-padded to 64) and on the host cores with the oracle port, same shape."""
    import torch
    from music_b200.wavenet.model import wavenet
    from music_b200.wavenet.train import Trainer
    from oracle import wavenet_oracle as O
    torch.manual_seed(0)
    net = wavenet(2, DIL, 32, 32, 256, Q, False).to(dev)
    rf = net.receptive_field
    Lc = 16000
    W = Lc - rf + 1
    codes = O.mu_law_encode(O.synthetic_audio(1, Lc + 1, seed=99), Q)
    piece, tgt = codes[:, :Lc].contiguous().to(dev), codes[:, rf:rf + W].contiguous().to(dev)
    tr = Trainer(net, "adam", learning_rate=1e-4, distributed=False)
    for _ in range(3):
        tr.step(piece, tgt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        loss = tr.step(piece, tgt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    # the same step replayed from one CUDA graph (Trainer.capture): one clip is 126 tiles per layer, the eager step is bound by the
    # host's ~80 launches
    ms_graph = None
    if tr.capture(piece, tgt):
        for _ in range(3):
            tr.step(piece, tgt)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            loss_g = tr.step(piece, tgt)
        e1.record()
        torch.cuda.synchronize()
        ms_graph = e0.elapsed_time(e1) / n
    # CPU, same shape
    torch.set_num_threads(os.cpu_count())
    st = O.init_wavenet_state(DIL, 32, 32, 256, Q, False, seed=0)
    x = O.one_hot(codes[:, :Lc], Q)
    ts = O.TrainState(st, "adam", lr=1e-4)
    O.train_step(ts, DIL, x, codes[:, rf:rf + W].contiguous())
    t0 = time.perf_counter()
    k = 0
    while k < 2 or time.perf_counter() - t0 < cpu_seconds:
        O.train_step(ts, DIL, x, codes[:, rf:rf + W].contiguous())
        k += 1
    dt = (time.perf_counter() - t0) / k
    return {"workload": "wavenet/train.py default shape: 30 layers (1..512 x3), 32 residual / 32 dilation / 256 skip ch, 1 clip of "
                        f"16000 samples (W = {W} targets), Adam",
            "gpu": {"samples_per_s": W / (ms * 1e-3), "ms_per_step": ms, "mode": net.mode, "loss": float(loss),
                    "note": "tcgen05 kernels, channels zero-padded 32 -> 64; 1 clip = 126 tiles per layer < 148 SMs",
                    "cuda_graph": None if ms_graph is None else {"ms_per_step": ms_graph, "samples_per_s": W / (ms_graph * 1e-3)}},
            "cpu": {"samples_per_s": W / dt, "s_per_step": dt, "cores": os.cpu_count(), "kind": "port", "steps": k}}


def bench_autoencoder(dev, steps=10, world=1, rank=0):
    """BASELINE.json configs[4] shape: wavenet_autoencoder with the shipped parameters (40 layers, 32 channels, 512
    bottleneck / skip, pool 512), one clip of W = 64000 targets (L = 68093) per GPU, Adam; with N > 1 ranks the gradients are
    averaged with one all-reduce per step (weak scaling) and the rate is the aggregate over all GPUs, timed as the max over
    ranks.  mode "auto" = bf16: conditioned decoder on the tcgen05 WaveNet kernels, encoder / conditioning convs on the
    mma.sync kernels of csrc/ae_fast.cu; the step is the fused AeTrainer.step (music_b200/wavenet_autoencoder/train.py)."""
    import torch
    from music_b200.wavenet_autoencoder.model1 import wavenet_autoencoder
    from music_b200.wavenet_autoencoder import train as T
    dil = [2 ** i for i in range(10)] * 4
    torch.manual_seed(0)
    net = wavenet_autoencoder(2, 256, dil, 32, 32, 512, 512, 32, 32, 512, False, mode="auto").to(dev)
    W = 64000
    L = net.receptive_field + W - 1
    g = torch.Generator().manual_seed(1234 + rank)
    idx = torch.randint(0, 256, (1, L), generator=g).to(dev)
    target = idx[:, net.receptive_field - 1:].contiguous()
    tr = T.AeTrainer(net, 'Adam', 1e-4, distributed=world > 1)

    def timed(idx_, target_, n):
        for _ in range(3):
            tr.step(idx_, target_)
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            loss_ = tr.step(idx_, target_)
        e1.record()
        torch.cuda.synchronize()
        ms_ = e0.elapsed_time(e1) / n
        if world > 1:
            t = torch.tensor([ms_], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_ = float(t[0])
        return ms_, float(loss_)
    ms, loss = timed(idx, target, steps)
    # the same step replayed from one CUDA graph (AeTrainer.capture): ~300 launches = 4.1 ms of host time per eager step
    graphed = None
    if world == 1 and tr.capture(idx, target):
        ms_g, loss_g = timed(idx, target, steps)
        graphed = {"ms_per_step": ms_g, "samples_per_s": W / (ms_g * 1e-3), "tflops": W * 8.656e6 / (ms_g * 1e-3) / 1e12, "loss": loss_g}
        tr.__dict__.pop("_graph", None)
    # the same step with 4 clips per GPU (the per-GPU batch is not fixed by configs[4]; one clip is 532 tiles per layer = 3.6 per SM)
    idx4 = torch.randint(0, 256, (4, L), generator=g).to(dev)
    ms4, _ = timed(idx4, idx4[:, net.receptive_field - 1:].contiguous(), max(3, steps // 2))
    return {"workload": "wavenet_autoencoder 40 layers (1..512 x4), 32 ch, bottleneck/skip 512, pool 512; 1 clip x 64000 targets "
                        f"(L=68093) per GPU on {world} GPU(s), index input, fused Adam step (AeTrainer), mode {net.mode}",
            "samples_per_s": world * W / (ms * 1e-3), "ms_per_step": ms, "dtype": "bf16" if net.mode == "bf16" else "f32",
            "loss": loss, "n_gpus": world, "steps": steps,
            "train_flops_per_sample": 8.656e6, "tflops": world * W * 8.656e6 / (ms * 1e-3) / 1e12,
            "cuda_graph": graphed,
            "clips_per_gpu_4": {"ms_per_step": ms4, "samples_per_s": world * 4 * W / (ms4 * 1e-3),
                                "tflops": world * 4 * W * 8.656e6 / (ms4 * 1e-3) / 1e12}}


def workload_config(world, B):
    """The `config` object both arms print (BASELINE.json configs[1] / configs[2] per GPU)."""
    return {"workload": "wavenet 30 layers (dilations 1..512 x3), 64 residual/64 dilation/256 skip ch, "
                        f"batch {B} x 16k-sample windows per GPU (L=19070), Adam, index (true one-hot) input",
            "global_batch": world * B, "window": WINDOW, "parallelism": f"dp{world}",
            "l2": "per-step working set ~4.6 GB of activations rewritten every step >> 126 MB L2; "
                  "4 distinct input batches rotate"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    rate, cores, dt = cpu_reference_rate(steps, max(1, min(args.warmup, 2)))
    sample = f"1 clip x {WINDOW} targets per step (cfg-2 model), {steps} steps, torch CPU fp32 oracle port"
    out = {"impl": "reference", "metric": "training audio samples/sec", "value": rate, "unit": "samples/s",
           "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": dict(workload_config(max(1, args.gpus), args.batch),
                          reference_note="CPU arm: rank 0 steps a bounded sample (one clip per step) on the host cores; "
                                         "the rate per target sample does not depend on the batch size"),
           "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--gen-steps", type=int, default=160000,
                    help="incremental-generation steps timed per stream (configs[3]: 160000 = 10 s of audio; 0 = skip)")
    ap.add_argument("--no-incumbent", action="store_true", help="skip the torch-eager GPU incumbent leg")
    ap.add_argument("--no-cfg1", action="store_true", help="skip the configs[0] (32/32/256, one clip) leg")
    ap.add_argument("--no-dense-e2e", action="store_true", help="skip the end-to-end variant fed with the reference loader's dense (B,256,L) batch")
    ap.add_argument("--gen-streams", type=int, default=64)
    ap.add_argument("--no-ae", action="store_true", help="skip the autoencoder (configs[4] shape) leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from music_b200 import _lib as L
    from music_b200.wavenet.audio_func import mu_law_encode
    from music_b200.wavenet.model import wavenet
    from music_b200.wavenet.train import Trainer
    from oracle import wavenet_oracle as O      # synthetic audio generator + cpu_baseline leg only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: music_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    K = args.steps
    B = args.batch

    torch.manual_seed(0)
    net = wavenet(2, DIL, D, R, S, Q, False, mode=args.mode).to(dev)
    if world > 1:                       # identical replicas
        for p in net.parameters():
            dist.broadcast(p.data, 0)
    trainer = Trainer(net, "adam", learning_rate=1e-4)
    rf = net.receptive_field
    Lx = rf + WINDOW - 1
    n_batches = 4
    audio = O.synthetic_audio(n_batches * B, Lx + 1, seed=1234 + rank)
    codes = mu_law_encode(audio.to(dev), Q).view(n_batches, B, Lx + 1)
    pieces = [codes[i, :, :Lx].contiguous() for i in range(n_batches)]
    targets = [codes[i, :, rf:rf + WINDOW].contiguous() for i in range(n_batches)]
    host_pieces = [p.cpu().pin_memory() for p in pieces]
    host_targets = [t.cpu().pin_memory() for t in targets]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`)
    for i in range(W):
        trainer.step(pieces[i % n_batches], targets[i % n_batches])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.load().wn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        loss = trainer.step(pieces[i % n_batches], targets[i % n_batches])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.load().wn_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    last_loss = float(loss)

    # ---------------- end to end through the public API with host buffers (`e2e`)
    # Every step's batch is copied from pinned host memory inside the timed region (K copies for K steps) and its loss
    # is read back; the package's BatchStager puts the copy of batch i + 1 on a copy stream under step i.
    from music_b200.wavenet.train import BatchStager
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

    def run_e2e(h_pieces, h_targets, n_steps):
        """n_steps steps fed from pinned host memory through BatchStager; the stager's device slots are allocated and
        the pipeline is primed by two untimed warm-up steps, so the timed region holds exactly one H2D copy per step."""
        stager = BatchStager(dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        total = n_steps + 2
        stager.put(h_pieces[0], h_targets[0])
        for i in range(total):
            if i == 2:
                barrier()
                ev0.record()
            d_piece, d_target = stager.get()
            if i + 1 < total:
                stager.put(h_pieces[(i + 1) % len(h_pieces)], h_targets[(i + 1) % len(h_pieces)])
            l = trainer.step(d_piece, d_target)
            stager.done()
            loss_host.copy_(l, non_blocking=True)
        ev1.record()
        barrier()
        return ev0.elapsed_time(ev1), d_piece.numel() * d_piece.element_size() + d_target.numel() * d_target.element_size()

    ms_e2e, h2d = run_e2e(host_pieces, host_targets, K)
    d2h = 4
    # the same through the reference loader's own batch format: dense (B,256,L) float "one-hot" pieces (faster_audio_data.py:62-83)
    dense_e2e = None
    if not args.no_dense_e2e:
        try:
            nd = 2
            h_dense = [torch.nn.functional.one_hot(host_pieces[i], Q).permute(0, 2, 1).float().contiguous().pin_memory() for i in range(nd)]
            kd = max(3, K // 2)
            ms_d, h2d_d = run_e2e(h_dense, host_targets[:nd], kd)
            td = torch.tensor([ms_d], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(td, op=dist.ReduceOp.MAX)
            dense_e2e = {"value": world * B * WINDOW * kd / (float(td[0]) * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d_d,
                         "d2h_bytes_per_step": 4, "ms_per_step": float(td[0]) / kd,
                         "input": "dense float32 (B,256,L) one-hot pieces as the reference's audio_data_loader yields them"}
            del h_dense
        except Exception as exc:
            dense_e2e = {"error": repr(exc)}

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    samples = world * B * WINDOW * K
    value = samples / (ms * 1e-3)
    e2e_value = samples / (ms_e2e * 1e-3)

    # ---------------- per-kernel breakdown: dominant kernel -> roofline
    roofline = None
    breakdown = None
    lib = L.load()
    n_prof = 3
    if rank == 0:
        lib.wn_profile_enable(1)
    for i in range(n_prof):          # every rank steps (the step contains a collective); only rank 0 records events
        trainer.step(pieces[i % n_batches], targets[i % n_batches])
    barrier()
    if rank == 0:
        rep = L.profile_report()
        lib.wn_profile_enable(0)
        kf = kernel_flops(B)
        tot_ms = sum(r[2] for r in rep) or 1.0
        breakdown = [{"kernel": n, "launches_per_step": c / n_prof, "ms_per_step": m / n_prof,
                      "share": m / tot_ms, "tflops": (kf[n] * n_prof / (m * 1e-3) / 1e12) if n in kf and m > 0 else None}
                     for n, c, m in rep]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
        top = next((r for r in rep if r[0] in kf), None)
        if top is not None:
            n, c, m = top
            ach = kf[n] * n_prof / (m * 1e-3) / 1e12
            roofline = {"bound": "tensor", "kernel": n, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                        "traffic": ncu_traffic(n), "peak_source": peak_src, "launches_per_step": c / n_prof,
                        "avg_launch_ms": m / c,
                        "step_tensor_frac": (value * 3 * flops_per_sample() / world) / (peak * 1e12)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        _, _, dt1 = cpu_reference_rate(1, 1)
        n = max(1, int(args.cpu_seconds / max(dt1, 1e-3)))
        rate, cores, dt = cpu_reference_rate(n, 0)
        cpu = {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"oracle port of wavenet/model.py + train.py:171-182 step, torch CPU fp32, cfg-2 model, "
                         f"1 clip x {WINDOW} targets per step, {n} steps ({dt:.2f} s/step)"}

    # strong-scaling reading of configs[2]: global batch 128 at 2 / 4 GPUs (64 / 32 clips per GPU); 8 GPUs x 16 is the headline itself
    strong = None
    if world in (2, 4):
        try:
            Bs = 128 // world
            audio_s = O.synthetic_audio(Bs, Lx + 1, seed=4321 + rank)
            codes_s = mu_law_encode(audio_s.to(dev), Q)
            p_s, t_s = codes_s[:, :Lx].contiguous(), codes_s[:, rf:rf + WINDOW].contiguous()
            for _ in range(2):
                trainer.step(p_s, t_s)
            barrier()
            es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ks = max(3, K // 2)
            es0.record()
            for _ in range(ks):
                trainer.step(p_s, t_s)
            es1.record()
            barrier()
            ts_ = torch.tensor([es0.elapsed_time(es1)], dtype=torch.float64, device=dev)
            dist.all_reduce(ts_, op=dist.ReduceOp.MAX)
            strong = {"global_batch": 128, "per_gpu_batch": Bs, "ms_per_step": float(ts_[0]) / ks,
                      "value": 128 * WINDOW * ks / (float(ts_[0]) * 1e-3), "unit": "samples/s", "scaling": "strong"}
            del p_s, t_s, codes_s
        except Exception as exc:
            strong = {"error": repr(exc)}

    gen = None
    if rank == 0 and world == 1 and args.gen_steps > 0:
        gen = bench_generation(net, args.gen_streams, args.gen_steps, dev)
    incumbent = cfg1 = None
    if rank == 0 and world == 1 and not args.no_incumbent:
        try:
            incumbent = bench_gpu_incumbent(dev, B)
        except Exception as exc:
            incumbent = {"error": repr(exc)}
    if rank == 0 and world == 1 and not args.no_cfg1:
        try:
            cfg1 = bench_cfg1(dev)
        except Exception as exc:
            cfg1 = {"error": repr(exc)}
    ae = None
    if not args.no_ae:                  # every rank takes part (one gradient all-reduce per step when world > 1)
        try:
            ae = bench_autoencoder(dev, world=world, rank=rank)
        except Exception as exc:        # reported, never fatal for the headline line
            ae = {"error": repr(exc)}

    if rank == 0:
        out = {"metric": "training audio samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": K,
               "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "bf16" if args.mode == "bf16" else "f32", "data": "synthetic",
               "config": workload_config(world, B),
               "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": ms_e2e / K, "input": "int64 mu-law codes (B,L) + targets (B,W), pinned host memory",
                       "dense_input": dense_e2e},
               "gpu_launches": int(launches), "loss": last_loss, "clocks": clocks, "roofline": roofline,
               "cpu_baseline": cpu, "train_flops_per_sample": 3 * flops_per_sample(), "generation": gen,
               "autoencoder": ae, "gpu_incumbent": incumbent, "cfg1": cfg1, "strong_scaling": strong, "kernels": breakdown}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
