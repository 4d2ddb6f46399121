#!/usr/bin/env python
"""BASELINE.json configs[4]: STFT -> Wiener mask -> ISTFT round trip over 1k..64k utterances (HBM-bound microbench).

Per utterance (T = 48 000, N = 185, F = 513, ld = 520): algorithmic bytes
  STFT   read 4T, write 8FN (+4FN for |X|^2)      = 1 331 160 B
  Wiener read 8FN + 2*4FN, write 2*8FN             = 3 036 960 B   (masks WFs, WFn -> S_hat, N_hat)
  ISTFT  read 8FN, write 4T                        =   951 240 B   (x2: s_hat and n_hat)
Prints one JSON line per batch size with the device time of each kernel (CUDA events, 5 repetitions after 2 warm-ups,
inputs larger than L2) and the achieved GB/s against the measured HBM peak.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dvae_b200 import _lib                                                          # noqa: E402
from dvae_b200.engine import RaggedBatch, _p, _stream, istft_batch, istft_masked_batch, stft_batch        # noqa: E402

DEV = torch.device("cuda:0")
T, N, F, LD = 48000, 185, 513, 520
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run(B_total, chunk=8192):
    Bc = min(B_total, chunk)
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.rand((Bc, T), device=DEV, generator=g) - 0.5
    off = torch.arange(Bc, device=DEV, dtype=torch.int64) * T
    lens = torch.full((Bc,), T, dtype=torch.int32, device=DEV)
    batch = RaggedBatch([N] * Bc, DEV)
    X, P = stft_batch(x.view(-1), off, lens, batch)
    mask = torch.rand((Bc * N, LD), device=DEV, generator=g)
    inv = 1.0 - mask
    S = torch.empty_like(X)
    Nn = torch.empty_like(X)
    y = torch.empty(Bc * T, device=DEV)
    n_chunks = (B_total + Bc - 1) // Bc

    t_stft = timed(lambda: _lib.call("dvae_stft_f32", _p(x), _p(off), _p(lens), Bc, _p(X), _p(P), _p(batch.fr_off), batch.NT, 1024, 256, LD, _stream()))
    t_wf = timed(lambda: _lib.call("dvae_wiener_apply", _p(X), _p(mask), _p(inv), 1, batch.NT, F, LD, _p(S), _p(Nn), _stream()))
    t_istft = timed(lambda: istft_batch(S, batch, off, lens, Bc * T, T, out=y))
    t_fused = timed(lambda: istft_masked_batch(X, mask, batch, off, lens, Bc * T, T, out=y))      # Wiener mask applied in the ISTFT load
    yy = istft_batch(X, batch, off, lens, Bc * T, T).view(Bc, T)
    err = (yy[:, 800:-800] - x[:, 800:-800]).abs().max().item()
    b_stft, b_wf, b_istft = 4 * T + 12 * F * N, 24 * F * N + 8 * F * N, 8 * F * N + 4 * T
    tot_ms = (t_stft + t_wf + 2 * t_istft) * n_chunks
    b_fused = 12 * F * N + 4 * T
    fused_ms = (t_stft + 2 * t_fused) * n_chunks
    out = dict(workload="configs[4] STFT->Wiener->ISTFT", utterances=B_total, chunk=Bc, round_trip_max_err=err,
               ms=dict(stft=t_stft, wiener=t_wf, istft=t_istft), total_ms=tot_ms,
               audio_s_per_s=B_total * 3.0 / (tot_ms * 1e-3),
               fused=dict(istft_masked_ms=t_fused, total_ms=fused_ms, audio_s_per_s=B_total * 3.0 / (fused_ms * 1e-3),
                          gbs=Bc * b_fused / t_fused / 1e6, frac_of_hbm_peak=Bc * b_fused / t_fused / 1e6 / PEAK),
               gbs=dict(stft=Bc * b_stft / t_stft / 1e6, wiener=Bc * b_wf / t_wf / 1e6, istft=Bc * b_istft / t_istft / 1e6),
               frac_of_hbm_peak=dict(stft=Bc * b_stft / t_stft / 1e6 / PEAK, wiener=Bc * b_wf / t_wf / 1e6 / PEAK,
                                     istft=Bc * b_istft / t_istft / 1e6 / PEAK), hbm_peak_gbs=PEAK)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [1024, 4096, 16384, 65536]
    for b in sizes:
        run(b)
