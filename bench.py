#!/usr/bin/env python
"""Benchmark of the MCEM VAE-NMF enhancement hot path (BASELINE.json metric: enhanced audio-seconds per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One *step* = one pass of the hot path (STFT -> 100 EM iterations of MH E-step + NMF M-step -> Wiener -> ISTFT) over
one batch of synthetic 3 s utterances.  Workload at every N: BASELINE.json configs[1] per GPU ("M1 batch of 512
synthetic utterances on 1 B200"), i.e. weak scaling with 512 utterances per rank and no data-path collective;
NCCL only gathers per-utterance metrics and the max-over-ranks time.

Printed JSON line (rank 0): `value` times the device-resident path with CUDA events; `e2e` times the public host API
(`dvae_b200.engine.Enhancer.enhance`: pinned host buffers in, host arrays out, copies inside the timed region);
`roofline` describes the dominant kernel (the Metropolis-Hastings sampler) from CUDA events recorded around it
inside the timed region; `cpu_baseline` is the CPU oracle port (the reference's algorithm, torch CPU ops, all host
threads) on one utterance of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from dvae_b200 import synth  # noqa: E402

METRIC = "enhanced_audio_seconds_per_second"
UNIT = "audio-s/s"
SECONDS = 3.0
DEC_FLOP_PER_ROW = {("M1", 16): 168192, ("M2", 16): 168448, ("M2v3", 16): 168448}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="utterances per GPU")
    ap.add_argument("--niter", type=int, default=100)
    ap.add_argument("--variant", default="M1", choices=["M1", "M2", "M2v3"])
    ap.add_argument("--sampler", default=os.environ.get("DVAE_SAMPLER", "tc"), choices=["fp32", "tc"])
    ap.add_argument("--chains", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def measured_traffic(B, variant):
    """DRAM bytes of one sampler launch from the committed ncu capture, when one exists for this workload."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_tc_traffic.json")
    try:
        t = json.load(open(path))
    except (OSError, ValueError):
        return None
    if t.get("utterances_per_gpu") == B and t.get("variant") == variant:
        return t.get("dram_bytes_per_launch")
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tensor=1590.0, src="fallback")


def model_weights(variant, prior_bias):
    y_dim = 0 if variant == "M1" else 1
    return synth.xavier_state_dict(variant, 513, 16, [128, 128], y_dim, seed=1234, out_bias=prior_bias)


def schedule(variant, niter):
    from dvae_b200.engine import McemConfig
    if variant == "M1":      # the reference's effective M1 schedule: 60 iterations keep 30; 105 keep 75
        return dict(niter=niter, keep_E=30, burn_E=30, keep_WF=75, burn_WF=30)
    return dict(niter=niter, keep_E=10, burn_E=30, keep_WF=25, burn_WF=75)


# ------------------------------------------------------------------------------------------------ reference arm / CPU baseline
def cpu_reference_run(variant, niter, u):
    """One utterance through the CPU oracle port exactly like ``process_utt`` (scripts/evaluate_ntcd_M1.py:81-188)."""
    from oracle import mcem_port, stft_np
    kw = dict(fs=16000, wlen_sec=64e-3, win="hann", hop_percent=0.25, center=False, pad_at_end=True)
    x, s, _ = synth.synth_utterance(u, SECONDS)
    t0 = time.perf_counter()
    X = stft_np.stft(x, **kw)
    S = stft_np.stft(s, **kw)
    sd = model_weights(variant, reference_power())
    y = synth.energy_vad(s) if variant != "M1" else None
    o = mcem_port.MCEMOracle(variant, niter, 10, 30, 25, 75, 0.01)
    o.init_parameters(X, S, sd, 10, 1e-8, y=y)
    o.run()
    ikw = dict(fs=16000, wlen_sec=64e-3, win="hann", hop_percent=0.25, center=False, max_len=len(x))
    stft_np.istft(o.S_hat, **ikw)
    stft_np.istft(o.N_hat, **ikw)
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    for i in range(args.warmup):
        cpu_reference_run(args.variant, max(1, args.niter // 20), i)       # warm-up: short runs (thread pools, caches)
    times = [cpu_reference_run(args.variant, args.niter, 100 + i) for i in range(args.steps)]
    t = statistics.median(times)
    value = SECONDS / t
    sample = "1 utterance (3 s, %d EM iterations) per step, sequential like the reference's process_utt" % args.niter
    out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
               ms_per_step=t * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
               data="synthetic", impl="reference",
               config=dict(workload="%s MCEM enhancement, synthetic 3 s 16 kHz utterances, STFT 1024/256, NMF rank 10, "
                                    "%d EM iterations (BASELINE.json configs[1] workload, one utterance per step)" % (args.variant, args.niter)),
               cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample),
               e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
               gpu_launches=0)
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [c.strip() for c in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=(statistics.median(sm) if sm else None), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch.distributed as dist
    from dvae_b200.engine import Enhancer, McemConfig, RaggedBatch
    from dvae_b200.shard import gather_metrics, max_over_ranks

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its debug output (the version banner at NCCL_DEBUG=VERSION / WARN / INFO) to STDOUT by default: send it
        # to stderr so that stdout carries exactly the one JSON line
        # (NCCL honours NCCL_DEBUG_FILE only above the VERSION level, so VERSION / unset becomes WARN)
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = args.batch
    u0 = rank * B                                     # global utterance ids of this rank's shard
    T = int(SECONDS * synth.FS)

    # synthetic inputs (host); a few distinct waveforms are tiled over the batch to keep set-up time short,
    # every utterance still gets its own Philox stream (global id) and its own NMF / latent state
    n_distinct = min(B, 32)
    base_x, base_s = synth.synth_batch(1000, n_distinct, SECONDS)
    idx = (np.arange(B) + u0) % n_distinct
    x_host = [base_x[i] for i in idx]
    y_host = [synth.energy_vad(base_s[i]) for i in idx] if args.variant != "M1" else None
    sd = model_weights(args.variant, reference_power())
    cfg = McemConfig(var_rw=0.01, nmf_rank=10, eps=1e-8, n_chains=args.chains, seed=2024, sampler=args.sampler,
                     **schedule(args.variant, args.niter))
    enh = Enhancer(sd, args.variant, cfg, device=local)
    eng = enh.engine
    utt_ids = list(range(u0, u0 + B))

    # device-resident copy of the inputs for the `value` measurement
    nfr = [synth.num_frames(T)] * B
    batch = RaggedBatch(nfr, dev, utt_ids)
    x_dev = torch.from_numpy(np.stack(x_host)).to(dev).reshape(-1)
    x_off = (torch.arange(B, dtype=torch.int64) * T).to(dev)
    x_len = torch.full((B,), T, dtype=torch.int32, device=dev)
    y_dev = None
    if y_host is not None:
        y_dev = torch.from_numpy(np.ascontiguousarray(np.concatenate([y.T for y in y_host], 0))).to(dev)

    def step_device():
        return enh.run_device(x_dev, x_off, x_len, batch, y_dev, B * T, T)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.timing = True
    eng._events = []
    eng.kernel_launches = 0
    eng.mh_rows = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        s_dev, n_dev, cost = step_device()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(e0.elapsed_time(e1), dev)
    stages = eng.stage_times_ms()
    launches = eng.kernel_launches
    mh_rows = eng.mh_rows
    eng.timing = False
    value = world * B * SECONDS * args.steps / (ms_total * 1e-3)

    # end-to-end through the public host API: pinned host buffers in, host arrays out, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        for _ in range(2):                                           # warm the pinned staging / result buffers (two result sets alternate)
            s_list, n_list, cost_h = enh.enhance(x_host, y_host, utt_ids=utt_ids)
        barrier()
        t0 = time.perf_counter()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        n_e2e = max(1, min(args.steps, 2))
        for _ in range(n_e2e):
            s_list, n_list, cost_h = enh.enhance(x_host, y_host, utt_ids=utt_ids)
        g1.record()
        barrier()
        wall = max_over_ranks(max(time.perf_counter() - t0, g0.elapsed_time(g1) * 1e-3), dev)
        e2e = dict(value=world * B * SECONDS * n_e2e / wall, unit=UNIT, h2d_bytes_per_step=int(enh.h2d_bytes),
                   d2h_bytes_per_step=int(enh.d2h_bytes), steps=n_e2e)

    # per-utterance metric gathered over ranks (the path's only collective): final cost of each utterance
    final_cost = cost[-1].to(torch.float64).reshape(B, 1)
    all_cost = gather_metrics(final_cost, world * B)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    flop_row = DEC_FLOP_PER_ROW[(args.variant, 16)]
    # the sampler kernel itself (events around the dvae_mh_chain_* launch); "mh" additionally holds the draw / pack kernels
    mh_ms, mh_calls = stages.get("mh_kernel", stages.get("mh", (0.0, 0)))
    roof = None
    if mh_calls:
        flops_per_call = flop_row * mh_rows / mh_calls                       # algorithmic: one decoder row per proposal
        achieved = flops_per_call / (mh_ms / mh_calls * 1e-3) / 1e12
        roof = dict(kernel="mh_chain_%s" % args.sampler, bound="tensor", achieved=achieved, peak=pk["tensor"], unit="TFLOP/s",
                    frac=achieved / pk["tensor"], traffic=measured_traffic(B, args.variant), peak_source=pk["src"],
                    share_of_step=mh_ms / ms_total, launches_per_step=mh_calls / args.steps,
                    note="sampler time = CUDA events around dvae_mh_chain_* inside the timed region; FLOPs = %d per decoder row x rows; "
                         "traffic = dram read+write bytes of one E-step launch from ncu (profiles/r01_tc_traffic.json), null if "
                         "no capture matches this batch" % flop_row)
    cfg_name = "BASELINE.json configs[1]" if (args.variant == "M1" and B == 512) else \
        "BASELINE.json configs[%d] shape (%s, %d utterances per GPU)" % (2 if args.variant == "M2" else (3 if args.variant == "M2v3" else 1), args.variant, B)
    stage_share = {k: round(v[0] / ms_total, 4) for k, v in stages.items()}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cpu_reference_run(args.variant, max(1, args.niter // 20), 0)
        t_cpu = cpu_reference_run(args.variant, args.niter, 1000)
        cpu = dict(value=SECONDS / t_cpu, unit=UNIT, cores=cores, kind="port",
                   sample="1 of the %d utterances (3 s, %d EM iterations) through oracle.mcem_port on %d host threads: %.1f s" %
                          (B, args.niter, cores, t_cpu))

    out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
               ms_per_step=ms_total / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
               dtype="f32" if args.sampler == "fp32" else "bf16", data="synthetic",
               config=dict(workload=cfg_name + ": %s batch of %d synthetic 3 s 16 kHz utterances per GPU, STFT 1024/256, "
                                    "NMF rank 10, %d EM iterations, MH schedule %s" % (args.variant, B, args.niter, schedule(args.variant, args.niter)),
                           utterances_per_gpu=B, sampler=args.sampler, chains=args.chains,
                           l2="working set per EM iteration (kept-sample variances %.1f GB) exceeds the 126 MB L2; no explicit flush" %
                              ((eng.VsT.numel() if getattr(eng, "VsT", None) is not None else eng.Vs_flat.numel() * 4) / 1e9),
                           parallelism="utterance shards, %d rank(s), no data-path collective" % world),
               clocks=clocks, e2e=e2e, gpu_launches=int(launches), roofline=roof, cpu_baseline=cpu,
               stage_share=stage_share, mean_final_cost=float(all_cost.mean().item()))
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def reference_power():
    """Decoder output bias shared by both arms: log of the mean clean-speech power spectrum of synthetic utterance 1000
    (random weights around it act as a crude speech prior, so the Wiener filter does real work)."""
    _, s, _ = synth.synth_utterance(1000, SECONDS)
    return synth.speech_prior_bias(s)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
