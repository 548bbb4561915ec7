#!/usr/bin/env python
"""bench.py — CTC loss+grad frames/sec on B200 (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3] [--impl ours|reference]

ours       one rank per GPU (torchrun for N>1).  A step = one pass of the hot path (nasr_ctc_loss_grad:
           the throughput kernel and the robust kernel's retry pass, + the batch-sum kernel and the
           4-scalar NCCL all-reduce when N>1) over one synthetic batch of
           the workload shape that is already resident in HBM.  `value` = frames all ranks
           processed / max-over-ranks device time.  `e2e` = the same metric through the HOST-buffer
           C-ABI call (nasr_host_ctc_step): pinned H2D of logits/labels, kernel, D2H of grad/loss
           inside the timed region.  `roofline` = algorithmic bytes / CUDA-event kernel time vs the
           measured HBM peak.  `cpu_baseline` = the oracle's C port (TF's algorithm, float, all host
           threads) timed on a bounded sample on rank 0.
reference  the CPU arm: the reference's path is TensorFlow's CPU CTC kernel, not installable here, so
           this times the oracle's C restatement of that algorithm on the host cores (kind "port").
Weak scaling: every rank holds its own B-utterance batch (utterances are independent; the only
exchange is the scalar all-reduce).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {  # BASELINE.json configs; cfg3 is the one the metric is quoted on
    "cfg1": dict(T=500, B=16, C=38, Lmax=100),
    "cfg2": dict(T=800, B=64, C=38, Lmax=150),
    "cfg3": dict(T=1000, B=256, C=38, Lmax=200),
    "cfg4": dict(T=3000, B=32, C=38, Lmax=600),
    "cfg5": dict(T=800, B=128, C=1024, Lmax=150),
}
METRIC = "ctc_loss_grad_frames_per_sec"
UNIT = "frames/s"
L2_BYTES = 126e6


def synth(w, seed):
    """SURVEY.md §8(d) throughput set: logits N(0,1)*3, labels uniform with 10% repeats,
    L ~ U[Lmax/2, Lmax], seq_len = T for every utterance."""
    rng = np.random.default_rng(seed)
    T, B, C, Lmax = w["T"], w["B"], w["C"], w["Lmax"]
    x = rng.standard_normal((T, B, C), dtype=np.float32) * np.float32(3.0)
    L = rng.integers(Lmax // 2, Lmax + 1, size=B)
    L[0] = Lmax
    vals = rng.integers(0, C - 1, size=int(L.sum())).astype(np.int32)
    rep = rng.random(vals.size) < 0.1
    offs = np.zeros(B + 1, dtype=np.int32)
    offs[1:] = np.cumsum(L)
    starts = set(offs[:-1].tolist())
    for i in np.nonzero(rep)[0]:
        if i not in starts and i > 0:
            vals[i] = vals[i - 1]
    seq = np.full(B, T, dtype=np.int32)
    return x, vals, offs, seq


def workload_config(name, w, world):
    """The `config` object of the JSON line: the same for our arm and for the reference arm."""
    return {"workload": "%s: B=%d per GPU,T=%d,C=%d,L in [%d,%d],seq_len=T" % (name, w["B"], w["T"], w["C"], w["Lmax"] // 2, w["Lmax"]),
            "global_batch": w["B"] * world, "parallelism": "utterance-sharded dp%d" % world}


def algorithmic_bytes(w, seq, nlabels):
    """4*C*(2*sum(seq_len) + T*B) + labels/seq_len/loss: logits read twice,
    gradient written once."""
    return 4 * w["C"] * (2 * int(seq.sum()) + w["T"] * w["B"]) + 4 * nlabels + 8 * w["B"]


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None

    def sample(self):
        if self._h is None:
            return
        nv = self._nv
        try:
            self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
            r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            for bit, name in names.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def run(self):
        while not self._stop_evt.is_set():
            self.sample()
            self._stop_evt.wait(0.02)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            self.sample()
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_port_rate(w, x, vals, offs, seq, n_utt, reps, budget_s=20.0):
    """frames/s of the oracle's C port (float, TF's algorithm, all host threads) on the first n_utt
    utterances of the batch.  Returns (frames_per_s, threads, seconds_per_rep)."""
    from oracle import c_oracle
    n_utt = min(n_utt, w["B"])
    xs = np.ascontiguousarray(x[:, :n_utt, :])
    so = offs[: n_utt + 1].copy()
    sv = vals[: so[-1]]
    ss = seq[:n_utt]
    c_oracle.ctc_loss_grad(xs, sv, so, ss, precision="f32", want_grad=True)   # warm-up pass (as the reference arm)
    dts = []
    t_start = time.perf_counter()
    for _ in range(max(1, reps)):
        t0 = time.perf_counter()
        c_oracle.ctc_loss_grad(xs, sv, so, ss, precision="f32", want_grad=True)
        dts.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    mean = sum(dts) / len(dts)   # the mean, like the reference arm's steps, so that the two agree
    return float(ss.sum()) / mean, c_oracle.num_threads(), mean


def run_reference(args, w, rank, world):
    """The reference's CPU path on this box's host cores: the C port of TensorFlow's CTCLossCalculator (TF itself is not
    installable offline) over the WHOLE batch of the workload, utterances sharded over all host threads as TF's op
    shards them.  One step = one pass over the batch (0.1-0.3 s at cfg3)."""
    if rank != 0:
        return
    x, vals, offs, seq = synth(w, 1234)
    from oracle import c_oracle
    for _ in range(args.warmup):
        c_oracle.ctc_loss_grad(x, vals, offs, seq, precision="f32")
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c_oracle.ctc_loss_grad(x, vals, offs, seq, precision="f32")
    dt = time.perf_counter() - t0
    fps = float(seq.sum()) * args.steps / dt
    cores = c_oracle.num_threads()
    sample = "whole %s batch per step (B=%d,T=%d,C=%d), %d threads" % (args.workload, w["B"], w["T"], w["C"], cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args.workload, w, world),
        "reference_kind": "C port of TensorFlow's CPU CTCLossCalculator (TF itself is not installable offline); rank 0's batch",
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_ours(args, w, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from neuralasr_b200 import _lib, host, towers
    from neuralasr_b200.networks import common

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    saved_stdout = None
    if world > 1:
        # stdout carries the one JSON line only: NCCL prints its "NCCL version ..." banner to fd 1 when the
        # communicator comes up, so fd 1 points at stderr until the warm-up (first collective included) is over
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        # the one collective is 32 bytes: one channel (one CTA) is all it needs, and every further NCCL CTA is a CTA
        # slot the loss kernel of the next step does not get (NASR_NCCL_TUNE=0 leaves NCCL's defaults)
        towers.tune_nccl_for_scalars()
        dist.init_process_group("nccl", device_id=dev)
    from neuralasr_b200 import _build
    if not os.path.exists(_build.LIB_PATH) and world == 1:
        _build.build_library(force=True)   # fresh checkout: built artefacts are git-ignored (N > 1: run build() first)
    lib = _lib.load()
    T, B, C = w["T"], w["B"], w["C"]
    x, vals, offs, seq = synth(w, 1234 + rank)
    lens = np.diff(offs)
    rows = np.repeat(np.arange(B), lens)
    cols = np.arange(offs[-1]) - np.repeat(offs[:-1], lens)
    triple = (np.stack([rows, cols], 1).astype(np.int64), vals, np.asarray([B, int(lens.max())], np.int64))
    lab = common.prepare_labels(triple, dev)
    seq_d = torch.from_numpy(seq).to(dev)
    # rotating input/gradient sets so that a step never finds its logits in L2 (126 MB)
    set_bytes = 2 * x.nbytes
    nsets = max(2, int(np.ceil(3 * L2_BYTES / set_bytes)))
    x0 = torch.from_numpy(x).to(dev)
    logits = [x0] + [x0.clone() for _ in range(nsets - 1)]
    grads = [torch.empty_like(x0) for _ in range(nsets)]
    gl = torch.full((B,), 1.0 / B, dtype=torch.float32, device=dev)
    alg_bytes = algorithmic_bytes(w, seq, vals.size)

    pending = []

    def step(i, ev=None):
        k = i % nsets
        if ev is not None:
            ev[0].record()
        loss_b, _, status = common.ctc_loss_and_grad(logits[k], lab, seq_d, grad_loss=gl, out_grad=grads[k])
        if ev is not None:
            ev[1].record()
        sums = common.batch_sums(loss_b=loss_b)
        # the path's one collective: 4 float64 scalars over NCCL, asynchronous like the logging it feeds
        # (the next step's kernels do not wait for it; every reduction is waited for before the clock stops)
        if args.reduce == "step":
            _, work = towers.all_reduce_sums(sums, async_op=True)
            if work is not None:
                pending.append(work)
        return sums

    for i in range(args.warmup):
        step(i)
    for wk in pending:
        wk.wait()
    pending.clear()
    torch.cuda.synchronize()
    if saved_stdout is not None:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    launches0 = _lib.launch_count()
    t_start, t_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    host_t0 = time.perf_counter()
    for i in range(args.steps):
        sums = step(args.warmup + i, evs[i])
    host_loop_s = time.perf_counter() - host_t0     # host time to enqueue the K steps (launch-bound if ~ the device time)
    for wk in pending:
        wk.wait()              # the compute stream now waits for every step's reduction
    pending.clear()
    t_stop.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop()
    ms_total = t_start.elapsed_time(t_stop)
    kern_ms = [a.elapsed_time(b) for a, b in evs]
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        # every rank's own numbers, for the record: its timed region, its mean kernel time, its host loop time per step
        mine = torch.tensor([ms_total / args.steps, statistics.mean(kern_ms), host_loop_s / args.steps * 1e3],
                            dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"ms_per_step": [round(float(a[0]), 4) for a in allr],
                    "kernel_ms": [round(float(a[1]), 4) for a in allr],
                    "host_loop_ms_per_step": [round(float(a[2]), 4) for a in allr]}
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    mean_loss = float(sums[0].item() / sums[3].item())
    frames = float(seq.sum())
    value = frames * world * args.steps / (ms_total * 1e-3)

    # ---- strong scaling beside it: the SAME global batch of B utterances split over the ranks, B/N each
    # (the reference's towers are weak -- config.py:35-36 multiplies the batch by num_gpus -- so weak is the headline)
    strong = None
    if world > 1 and B % world == 0 and B // world >= 1:
        Bs = B // world
        xs_ = x0[:, :Bs, :].contiguous()
        so_ = offs[: Bs + 1].copy()
        sv_ = vals[: so_[-1]]
        lens_s = np.diff(so_)
        tr_s = (np.stack([np.repeat(np.arange(Bs), lens_s), np.arange(so_[-1]) - np.repeat(so_[:-1], lens_s)], 1).astype(np.int64),
                sv_, np.asarray([Bs, max(int(lens_s.max()), 1)], np.int64))
        lab_s = common.prepare_labels(tr_s, dev)
        seq_s = seq_d[:Bs].contiguous()
        gl_s = torch.full((Bs,), 1.0 / B, dtype=torch.float32, device=dev)
        ns = max(2, min(nsets * world, 24))
        lg_s = [xs_] + [xs_.clone() for _ in range(ns - 1)]
        gr_s = [torch.empty_like(xs_) for _ in range(ns)]

        def sstep(i):
            k = i % ns
            loss_b, _, _ = common.ctc_loss_and_grad(lg_s[k], lab_s, seq_s, grad_loss=gl_s, out_grad=gr_s[k])
            sm = common.batch_sums(loss_b=loss_b)
            _, work = towers.all_reduce_sums(sm, async_op=True)
            if work is not None:
                pending.append(work)

        for i in range(max(3, args.warmup)):
            sstep(i)
        for wk in pending:
            wk.wait()
        pending.clear()
        dist.barrier()
        torch.cuda.synchronize()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        for i in range(args.steps):
            sstep(i)
        for wk in pending:
            wk.wait()
        pending.clear()
        b_.record()
        torch.cuda.synchronize()
        dist.barrier()
        ts_ = torch.tensor([a_.elapsed_time(b_)], dtype=torch.float64, device=dev)
        dist.all_reduce(ts_, op=dist.ReduceOp.MAX)
        ms_s = float(ts_.item())
        strong = {"scaling": "strong", "global_batch": B, "per_gpu_batch": Bs, "ms_per_step": ms_s / args.steps,
                  "value": float(seq[:Bs].sum()) * world * args.steps / (ms_s * 1e-3), "unit": UNIT,
                  "l2": "%d rotating sets per rank" % ns}
        del lg_s, gr_s

    # ---- e2e through the HOST-buffer C-ABI call (pinned H2D + kernel + D2H per step) ----
    numa_cpus = host.bind_to_gpu_numa_node(local_rank) if world > 1 else None   # pinned staging on the GPU's node
    ctx = host.HostContext(local_rank, T, B, C, w["Lmax"])
    pin = ctx.pinned_logits[: x.size].reshape(T, B, C)
    pin[...] = x
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        out = ctx.step(pin, vals, offs, seq, grad_loss=np.full(B, 1.0 / B, np.float32), want_decode=False)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = ctx.step(pin, vals, offs, seq, grad_loss=np.full(B, 1.0 / B, np.float32), want_decode=False)
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e_value = frames * world * e2e_steps / e2e_s
    h2d = x.nbytes + vals.nbytes + offs.nbytes + seq.nbytes + 4 * B
    d2h = x.nbytes + 4 * B + 4 * B
    # the snapshot's whole train() fetch (tfnetwork.py:183-190): loss, gradient AND mean label error rate of the
    # beam search decoder, host buffers in and out — reported beside the headline, rank 0 only
    train_step = None
    if rank == 0:
        ctx.set_decoder("beam", 100)
        for _ in range(2):
            out = ctx.step(pin, vals, offs, seq, grad_loss=np.full(B, 1.0 / B, np.float32), want_decode=True)
        t0 = time.perf_counter()
        for _ in range(5):
            out = ctx.step(pin, vals, offs, seq, grad_loss=np.full(B, 1.0 / B, np.float32), want_decode=True)
        ts = (time.perf_counter() - t0) / 5
        train_step = {"ms_per_step": ts * 1e3, "frames_per_s": frames / ts, "mean_ler": float(out["ler"].mean()),
                      "what": "nasr_host_ctc_step with the beam decoder (width 100): H2D, loss+grad, beam search, "
                              "label error rate, D2H of grad/loss/ler"}
    ctx.close()

    # ---- decode + LER, reported beside the headline (not folded into it) ----
    dec_ms = None
    if rank == 0:
        for _ in range(3):
            d, _ = common.decoding(logits[0], seq_d)
            common.edit_distance(d, lab)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(10):
            d, _ = common.decoding(logits[i % nsets], seq_d)
            common.edit_distance(d, lab)
        b.record()
        torch.cuda.synchronize()
        dec_ms = a.elapsed_time(b) / 10

    # ---- beam search decode (what create_model runs today, tfnetwork.py:62: width 100, top path) + LER ----
    beam = None
    if rank == 0:
        for _ in range(2):
            dm, _ = common.create_model_beam(logits[0], seq_d)
            common.edit_distance(dm, lab)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(5):
            dm, _ = common.create_model_beam(logits[i % nsets], seq_d)
            common.edit_distance(dm, lab)
        b.record()
        torch.cuda.synchronize()
        beam_ms = a.elapsed_time(b) / 5
        from oracle import c_oracle
        n_utt = min(B, 2 * c_oracle.num_threads())       # bounded CPU sample: two utterances per host thread
        t0 = time.perf_counter()
        c_oracle.beam_search(x[:, :n_utt], seq[:n_utt], 100, 1, True)
        cpu_s = time.perf_counter() - t0
        beam = {"beam_width": 100, "top_paths": 1, "ms_per_batch": beam_ms, "frames_per_s": frames / (beam_ms * 1e-3),
                "includes": "beam search kernel + label error rate kernel",
                "cpu_port": {"frames_per_s": float(seq[:n_utt].sum()) / cpu_s, "cores": c_oracle.num_threads(),
                             "sample": "%d utterances of the same batch, oracle/beam_oracle.c" % n_utt}}

    # ---- library baseline on the same GPU: torch's CUDA log_softmax + ctc_loss + backward (a number for context,
    #      and an independent check of the mean loss), rank 0 only
    lib_base = None
    if rank == 0:
        try:
            xt = logits[0].detach().clone().requires_grad_(True)
            tgt = torch.from_numpy(vals.astype(np.int64)).to(dev)
            tl = torch.from_numpy(lens.astype(np.int64)).to(dev)
            il = torch.from_numpy(seq.astype(np.int64)).to(dev)

            def torch_step():
                xt.grad = None
                lsm = torch.log_softmax(xt, dim=2)
                l_b = torch.nn.functional.ctc_loss(lsm, tgt, il, tl, blank=C - 1, reduction="none", zero_infinity=False)
                l_b.mean().backward()
                return l_b

            for _ in range(3):
                l_b = torch_step()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                l_b = torch_step()
            b.record()
            torch.cuda.synchronize()
            lib_base = {"what": "torch log_softmax + nn.functional.ctc_loss (CUDA) + backward, same batch",
                        "ms_per_step": a.elapsed_time(b) / 10, "mean_loss": float(l_b.mean().item())}
            del xt, l_b
        except Exception as e:  # a context number, never a reason to fail the bench
            lib_base = {"unavailable": repr(e)[:200]}

    # ---- the model tail's affine projection in front of the loss (bilstm_ctc_net.py:33-45: K = 500):
    #      forward, dH, dW + db on B*T rows, two rotating H sets (2 x 4*rows*K bytes > L2), against the HBM roofline
    proj = None
    if rank == 0:
        try:
            proj = projection_record(common, dev, B * T, 500, C)
        except Exception as e:  # a secondary record: never a reason to fail the headline
            proj = {"unavailable": repr(e)[:200]}

    if rank == 0:
        peak, peak_kind = hbm_peak()
        k_ms = statistics.mean(kern_ms)
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[args.workload]["dram_bytes_per_launch"]
        except Exception:
            pass
        cpu = None
        if world == 1 or True:
            fps, cores, dt = cpu_port_rate(w, x, vals, offs, seq, n_utt=B, reps=5)
            cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "whole %s batch (B=%d,T=%d), mean of <=5 passes after one warm-up, %.3f s per pass" % (args.workload, B, T, dt)}
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, w, world),
            "notes": {"l2": "%d rotating logits/grad sets (%.0f MB) > 126 MB L2" % (nsets, nsets * set_bytes / 1e6),
                      "mean_loss": mean_loss},
            "strong_scaling": strong,
            "per_rank": per_rank,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_kind": peak_kind,
                         "kernel": "ctc_fast_kernel + ctc_robust_kernel retry pass (both launches of nasr_ctc_loss_grad_f32)",
                         "kernel_ms": k_ms,
                         "algorithmic_bytes_per_launch": alg_bytes},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s / e2e_steps * 1e3, "steps": e2e_steps,
                    "numa_bound_cpus": len(numa_cpus) if numa_cpus else None},
            "gpu_launches": launches,
            "clocks": clocks,
            "decode_ler_ms": dec_ms,
            "beam_search": beam,
            "e2e_train_step": train_step,
            "library_baseline": lib_base,
            "projection": proj,
        }))
    if world > 1:
        dist.destroy_process_group()


def _forward_mma_sync(common, H, W, b, out):
    """The forward on the mma.sync kernel (what shapes outside the tcgen05 kernel's rules get), for comparison."""
    prev = os.environ.get("NASR_AFFINE_TC")
    os.environ["NASR_AFFINE_TC"] = "0"
    try:
        return common.affine_logits(H, W, b, out=out)
    finally:
        if prev is None:
            del os.environ["NASR_AFFINE_TC"]
        else:
            os.environ["NASR_AFFINE_TC"] = prev


def _backward_mma_sync(common, H, W, dL):
    prev = os.environ.get("NASR_AFFINE_TC")
    os.environ["NASR_AFFINE_TC"] = "0"
    try:
        return common.affine_backward(H, W, dL, True, True, True, dense_dH=True)
    finally:
        if prev is None:
            del os.environ["NASR_AFFINE_TC"]
        else:
            os.environ["NASR_AFFINE_TC"] = prev


def projection_record(common, dev, rows, K, C):
    """ms per call and fraction of the HBM peak of nasr_affine_logits_f32 / nasr_affine_backward_f32."""
    import torch
    peak, _ = hbm_peak()
    rows = min(rows, 512000)                       # two H sets of at most 1 GB each
    g = torch.Generator(device=dev).manual_seed(7)
    Hs = [torch.randn((rows, K), device=dev, generator=g) for _ in range(2)]
    W = torch.randn((K, C), device=dev, generator=g) / K ** 0.5
    b = torch.zeros((C,), device=dev)
    dL = torch.randn((rows, C), device=dev, generator=g)
    out = torch.empty((rows, C), device=dev)

    def timed(fn, n=10):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            fn(i)
        e.record()
        torch.cuda.synchronize()
        return a.elapsed_time(e) / n

    rec = {"rows": rows, "K": K, "C": C,
           "arithmetic": "3xTF32, float32 accumulation, all on tcgen05 (csrc/affine_tc.cu, affine_tc_dh.cu, affine_tc_dw.cu); *_mma_sync = the mma.sync kernels that take shapes outside their rules",
           "dH_pitch": "backward_dH writes rows pitched at a multiple of 32 floats (128-byte lines); _dense_pitch is a contiguous [rows, K]",
           "l2": "two rotating H sets (%.0f MB) > 126 MB L2" % (2 * 4 * rows * K / 1e6)}
    for name, fn, nbytes in (
            ("forward", lambda i: common.affine_logits(Hs[i & 1], W, b, out=out), 4 * (rows * K + rows * C + K * C)),
            ("forward_mma_sync", lambda i: _forward_mma_sync(common, Hs[i & 1], W, b, out), 4 * (rows * K + rows * C + K * C)),
            ("backward_dH", lambda i: common.affine_backward(Hs[i & 1], W, dL, True, False, False), 4 * (rows * K + rows * C)),
            ("backward_dH_dense_pitch", lambda i: common.affine_backward(Hs[i & 1], W, dL, True, False, False, dense_dH=True),
             4 * (rows * K + rows * C)),
            ("backward_dW_db", lambda i: common.affine_backward(Hs[i & 1], W, dL, False, True, True), 4 * (rows * K + rows * C)),
            ("backward_mma_sync", lambda i: _backward_mma_sync(common, Hs[i & 1], W, dL), 4 * 2 * (rows * K + rows * C))):
        ms = timed(fn)
        rec[name] = {"ms": ms, "algorithmic_gbs": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak}
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        rec["library_fp32"] = {"forward_ms": timed(lambda i: torch.addmm(b, Hs[i & 1], W, out=out)),
                               "backward_ms": timed(lambda i: (torch.mm(dL, W.t()), torch.mm(Hs[i & 1].t(), dL), dL.sum(0))),
                               "what": "torch.addmm / torch.mm float32 (TF32 off) on the same tensors"}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reduce", default="step", choices=["step", "none"],
                    help="experiment: 'none' leaves the per-step scalar all-reduce out (upper bound of weak scaling)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w, rank, world)
    else:
        run_ours(args, w, rank, world, local_rank)


if __name__ == "__main__":
    main()
