"""Generate ``tests/golden/mcem_*.npz`` from the UNMODIFIED reference.  Oracle tooling: test infrastructure only.

Run in the build container only (``/root/reference`` does not exist on the GPU box)::

    python oracle/make_golden.py

For each case it
  1. builds the reference's own VAE container (``packages/models/models.py``) and loads seeded weights
     produced by ``dvae_b200.synth.xavier_state_dict`` (which also pins the state_dict key layout),
  2. runs the reference's ``MCEM_*`` class (``packages/models/mcem.py``) on a seeded synthetic spectrogram
     while recording every ``torch.rand`` / ``torch.randn`` draw in consumption order,
  3. replays the same draws into ``oracle.mcem_port.MCEMOracle`` and requires bit-identical outputs,
  4. stores inputs, draws and the reference's outputs as a compressed ``.npz``.

The fixtures are what pins the oracle (and, through it, the CUDA path) to the reference.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, ROOT)

from dvae_b200 import synth  # noqa: E402
from oracle import mcem_port, stft_np  # noqa: E402

CASES = [
    # name, variant, x_dim(n_fft), N frames, K, L, h_dim, y_dim, niter, (nsE, bE, nsWF, bWF)
    dict(name="tiny_M1", variant="M1", n_fft=64, T=64 * 5, K=3, L=4, h=[8, 8], y_dim=0, niter=3, sched=(2, 3, 3, 4)),
    dict(name="tiny_M2", variant="M2", n_fft=64, T=64 * 5, K=3, L=4, h=[8, 8], y_dim=1, niter=3, sched=(2, 3, 3, 4)),
    dict(name="tiny_M2v2", variant="M2v2", n_fft=64, T=64 * 5, K=3, L=4, h=[8, 6], y_dim=1, niter=3, sched=(2, 3, 3, 4)),
    dict(name="tiny_M2v3", variant="M2v3", n_fft=64, T=64 * 5, K=3, L=4, h=[8, 6], y_dim=1, niter=3, sched=(2, 3, 3, 4)),
    dict(name="full_M1", variant="M1", n_fft=1024, T=1024 + 256 * 23, K=10, L=16, h=[128, 128], y_dim=0, niter=2,
         sched=(10, 6, 25, 8)),
    dict(name="full_M2", variant="M2", n_fft=1024, T=1024 + 256 * 23, K=10, L=16, h=[128, 128], y_dim=1, niter=2,
         sched=(4, 6, 5, 7)),
    dict(name="full_M2v3", variant="M2v3", n_fft=1024, T=1024 + 256 * 23, K=10, L=16, h=[128, 128], y_dim=1, niter=2,
         sched=(4, 6, 5, 7)),
    dict(name="cfg1_M1", variant="M1", n_fft=1024, T=1024 + 256 * 15, K=10, L=32, h=[128], y_dim=0, niter=2,
         sched=(10, 5, 25, 6)),
]


def synth_spectra(case, seed):
    """Seeded noisy / clean STFTs (F, N) complex64 and labels for one case (any n_fft)."""
    rng = np.random.default_rng(seed)
    T, n_fft = case["T"], case["n_fft"]
    t = np.arange(T) / 16000.0
    s = np.sin(2 * np.pi * 180.0 * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 5 * t)) ** 2 * 0.3
    n = 0.05 * rng.standard_normal(T)
    fs, wlen = 16000, n_fft / 16000.0
    kw = dict(fs=fs, wlen_sec=wlen, win="hann", hop_percent=0.25, center=False, pad_at_end=True)
    X = stft_np.stft(s + n, **kw)
    S = stft_np.stft(s, **kw)
    y = None
    if case["y_dim"]:
        y = (rng.uniform(size=(case["y_dim"], X.shape[1])) > 0.4).astype(np.float32)
    return X, S, y


def build_reference(case, sd):
    sys.path.insert(0, REF)
    from packages.models import mcem as ref_mcem          # noqa: E402  (reseeds numpy/torch to 0 on import)
    from packages.models import models as ref_models      # noqa: E402
    F = case["n_fft"] // 2 + 1
    v = case["variant"]
    if v == "M1":
        model = ref_models.VariationalAutoencoder([F, case["L"], case["h"]])
        algo_cls = ref_mcem.MCEM_M1
    elif v == "M2":
        model = ref_models.DeepGenerativeModel([F, case["y_dim"], case["L"], case["h"]], None)
        algo_cls = ref_mcem.MCEM_M2
    else:
        model = ref_models.DeepGenerativeModel_v5([F, case["y_dim"], case["L"], case["h"]]).enc_dec_clf
        algo_cls = ref_mcem.MCEM_M2v2 if v == "M2v2" else ref_mcem.MCEM_M2v3
    missing, unexpected = model.load_state_dict({k: torch.tensor(a) for k, a in sd.items()}, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("classifier.") for k in missing), missing
    model.eval()
    for p in model.parameters():
        p.requires_grad = False
    return model, algo_cls


class Recorder:
    """Patch torch.rand / torch.randn to log draws in consumption order."""

    def __enter__(self):
        self.log = []
        self._rand, self._randn = torch.rand, torch.randn

        def rand(*a, **k):
            t = self._rand(*a, **k)
            self.log.append(("rand", t.detach().clone().numpy()))
            return t

        def randn(*a, **k):
            t = self._randn(*a, **k)
            self.log.append(("randn", t.detach().clone().numpy()))
            return t

        torch.rand, torch.randn = rand, randn
        return self

    def __exit__(self, *exc):
        torch.rand, torch.randn = self._rand, self._randn


def run_case(case, seed=7):
    F = case["n_fft"] // 2 + 1
    X, S, y = synth_spectra(case, seed)
    out_bias = float(np.log(np.mean(np.abs(X) ** 2)))
    sd = synth.xavier_state_dict(case["variant"] if case["variant"] != "M2v2" else "M2v3", F, case["L"], case["h"],
                                 case["y_dim"], seed=1234 + seed, out_bias=out_bias)
    model, algo_cls = build_reference(case, sd)
    # container forward (the reconstruct_* scripts' path): models.py:172-180 / 201-204 / 280-287
    rows = torch.tensor(np.abs(S.T[:5]) ** 2)
    yrows = None if y is None else torch.tensor(y.T[:5])
    torch.manual_seed(55 + seed)
    with Recorder() as frec:
        out = model(rows) if case["variant"] == "M1" else model(rows, yrows)
    fwd = dict(fwd_xmu=out[0].numpy(), fwd_mu=out[-2].numpy(), fwd_lv=out[-1].numpy(), fwd_eps=frec.log[0][1])
    pv = "M2v3" if case["variant"] == "M2v2" else case["variant"]
    px, _, pmu, plv = mcem_port.vae_forward(sd, pv, rows, yrows, mcem_port.ReplayDraws(frec.log))
    assert np.array_equal(px.numpy(), fwd["fwd_xmu"]) and np.array_equal(pmu.numpy(), fwd["fwd_mu"])
    assert np.array_equal(plv.numpy(), fwd["fwd_lv"])

    nsE, bE, nsWF, bWF = case["sched"]
    algo = algo_cls(niter=case["niter"], nsamples_E_step=nsE, burnin_E_step=bE, nsamples_WF=nsWF, burnin_WF=bWF,
                    var_RW=0.01)
    torch.manual_seed(100 + seed)
    with Recorder() as rec:
        if case["variant"] == "M1":
            algo.init_parameters(X=X, S=S, vae=model, nmf_rank=case["K"], eps=1e-8, device="cpu")
        else:
            algo.init_parameters(X=X, S=S, y=torch.tensor(y), vae=model, nmf_rank=case["K"], eps=1e-8, device="cpu")
        Z0 = algo.Z.clone().numpy()
        cost = algo.run()
    ref = dict(cost=cost, S_hat=algo.S_hat, N_hat=algo.N_hat, W=algo.W.numpy(), H=algo.H.numpy(), g=algo.g.numpy(),
               Z=algo.Z.numpy(), Z0=Z0, Vs_shape=np.array(algo.Vs.shape), **fwd)

    # the port must reproduce the reference bit-for-bit from the recorded draws
    port = mcem_port.MCEMOracle(case["variant"], case["niter"], nsE, bE, nsWF, bWF, 0.01,
                                draws=mcem_port.ReplayDraws(rec.log))
    port.init_parameters(X, S, sd, case["K"], 1e-8, y=y)
    pcost = port.run()
    checks = dict(cost=pcost, S_hat=port.S_hat, N_hat=port.N_hat, W=port.W.numpy(), H=port.H.numpy(),
                  g=port.g.numpy(), Z=port.Z.numpy())
    for k, v in checks.items():
        if not np.array_equal(np.asarray(v), np.asarray(ref[k])):
            d = np.max(np.abs(np.asarray(v) - np.asarray(ref[k])))
            raise SystemExit("port != reference for %s in %s (max abs diff %g)" % (k, case["name"], d))
    assert port.draws.pos == len(rec.log)
    assert tuple(ref["Vs_shape"]) == tuple(port.Vs.shape)

    kinds = np.array([0 if k == "rand" else 1 for k, _ in rec.log], np.uint8)
    sizes = np.array([v.size for _, v in rec.log], np.int64)
    shapes = np.array([list(v.shape) + [0] * (2 - v.ndim) for _, v in rec.log], np.int32)
    flat = np.concatenate([v.ravel() for _, v in rec.log]).astype(np.float32)
    meta = dict(variant=case["variant"], n_fft=case["n_fft"], K=case["K"], L=case["L"], h=np.array(case["h"]),
                y_dim=case["y_dim"], niter=case["niter"], sched=np.array(case["sched"]), weight_seed=1234 + seed,
                out_bias=out_bias, eps=1e-8, var_RW=0.01)
    path = os.path.join(ROOT, "tests", "golden", "mcem_%s.npz" % case["name"])
    np.savez_compressed(path, X=X, S=S, y=(y if y is not None else np.zeros((0, X.shape[1]), np.float32)),
                        draw_kinds=kinds, draw_sizes=sizes, draw_shapes=shapes, draw_flat=flat,
                        **{"ref_" + k: v for k, v in ref.items()}, **{"meta_" + k: np.asarray(v) for k, v in meta.items()})
    print("%-10s ok: %d draws (%.0f KB), bit-identical port, file %.0f KB" %
          (case["name"], len(rec.log), flat.nbytes / 1024, os.path.getsize(path) / 1024))


def run_metrics():
    """``tests/golden/metrics.npz``: the reference's ``energy_ratios`` / ``si_sdr_leroux`` (packages/metrics.py:39-82) on
    seeded signals; the port must agree to 1e-10 dB."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_metrics", os.path.join(REF, "packages", "metrics.py"))
    ref_metrics = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_metrics)
    rng = np.random.default_rng(77)
    T, n_case = 2048, 3
    t = np.arange(T) / 16000.0
    s_all, n_all, sh_all, ref = [], [], [], []
    for c in range(n_case):
        s = (np.sin(2 * np.pi * (150.0 + 40 * c) * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 4 * t)) ** 2 * 0.3).astype(np.float32)
        n = (0.03 * (c + 1) * rng.standard_normal(T)).astype(np.float32)
        s_hat = (0.9 * s + 0.2 * n + 0.01 * rng.standard_normal(T)).astype(np.float32)
        r = ref_metrics.energy_ratios(s_hat.astype(np.float64), s.astype(np.float64), n.astype(np.float64))
        leroux = ref_metrics.si_sdr_leroux(s_hat.astype(np.float64), s.astype(np.float64))
        port = mcem_port.energy_ratios(s_hat, s, n)
        assert max(abs(a - b) for a, b in zip(r, port)) < 1e-10 and abs(mcem_port.si_sdr(s_hat, s) - leroux) < 1e-10
        assert abs(r[0] - leroux) < 1e-9
        s_all.append(s); n_all.append(n); sh_all.append(s_hat); ref.append(list(r) + [leroux])
    path = os.path.join(ROOT, "tests", "golden", "metrics.npz")
    np.savez_compressed(path, s=np.stack(s_all), n=np.stack(n_all), s_hat=np.stack(sh_all), ref=np.array(ref, np.float64))
    print("metrics    ok: %d cases, port within 1e-10 dB, file %.0f KB" % (n_case, os.path.getsize(path) / 1024))


def run_labels():
    """``tests/golden/labels.npz``: the reference's ``clean_speech_VAD`` / ``clean_speech_IBM`` /
    ``noise_robust_clean_speech_IBM`` (packages/processing/target.py:5-105) on a seeded utterance.  target.py imports
    ``librosa.util`` (absent here) for ``util.frame`` only; a stub module provides that one function (documented librosa
    semantics: column j = y[j*hop : j*hop + frame_length]), everything else is the reference's own code."""
    import importlib.util
    import types
    stub = types.ModuleType("librosa")
    stub.util = types.ModuleType("librosa.util")

    def frame(y, frame_length, hop_length):
        n = 1 + (len(y) - frame_length) // hop_length
        return y[np.arange(frame_length)[:, None] + hop_length * np.arange(n)[None, :]]

    stub.util.frame = frame
    sys.modules["librosa"], sys.modules["librosa.util"] = stub, stub.util
    spec = importlib.util.spec_from_file_location("ref_target", os.path.join(REF, "packages", "processing", "target.py"))
    ref_target = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_target)
    del sys.modules["librosa"], sys.modules["librosa.util"]
    rng = np.random.default_rng(11)
    T = 9000
    t = np.arange(T) / 16000.0
    env = np.clip(np.sin(2 * np.pi * 2.2 * t), 0, None) ** 2                    # speech bursts with silent gaps
    s = (env * sum(np.sin(2 * np.pi * 210.0 * k * t) / k for k in range(1, 30)) * 0.2 + 1e-4 * rng.standard_normal(T)).astype(np.float32)
    kw = dict(fs=16000, wlen_sec=64e-3, hop_percent=0.25, center=False, pad_at_end=True)
    S = stft_np.stft(s, win="hann", **kw)
    vad = ref_target.clean_speech_VAD(s.astype(np.float64), pad_mode="reflect", vad_threshold=1.70, **kw)
    ibm = ref_target.clean_speech_IBM(S, eps=1e-8, ibm_threshold=50)
    nr = ref_target.noise_robust_clean_speech_IBM(s.astype(np.float64), S, pad_mode="reflect", vad_threshold=1.70, eps=1e-8,
                                                  ibm_threshold=50, **kw)
    assert np.array_equal(vad, stft_np.clean_speech_vad(s)) and np.array_equal(ibm, stft_np.clean_speech_ibm(S))
    assert 0.1 < vad.mean() < 0.9 and 0.05 < ibm.mean() < 0.95, (vad.mean(), ibm.mean())
    path = os.path.join(ROOT, "tests", "golden", "labels.npz")
    np.savez_compressed(path, s=s, vad=vad, ibm=ibm.astype(np.uint8), nr=nr.astype(np.uint8))      # S = stft_np.stft(s) is recomputed by the tests
    print("labels     ok: VAD %d/%d frames active, IBM %.1f %% set, port identical, file %.0f KB" %
          (int(vad.sum()), vad.shape[1], 100 * ibm.mean(), os.path.getsize(path) / 1024))


def stats_inputs():
    """Seeded per-utterance metric rows and grouping labels for the ``compute_stats`` fixture (also used by the test)."""
    rng = np.random.default_rng(5)
    rows = [tuple(float(v) for v in rng.normal(loc=(8.0, 15.0, 9.0), scale=2.0)) for _ in range(48)]
    snr = np.array([(-5, 0, 5, 10)[i % 4] for i in range(48)])
    noise = [("Babble", "Cafe", "Car")[i % 3] for i in range(48)]
    return ["si_sdr", "si_sir", "si_sar"], rows, snr, noise


def run_stats():
    """``tests/golden/compute_stats.txt``: what the reference's ``compute_stats`` (packages/metrics.py:84-167) prints for seeded
    metric rows, grouped by input SNR and noise type."""
    import contextlib
    import importlib.util
    import io
    spec = importlib.util.spec_from_file_location("ref_metrics", os.path.join(REF, "packages", "metrics.py"))
    ref_metrics = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_metrics)
    keys, rows, snr, noise = stats_inputs()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ref_metrics.compute_stats(keys, rows, "", 0.95, all_snr_db=snr, all_noise_types=noise)
    path = os.path.join(ROOT, "tests", "golden", "compute_stats.txt")
    with open(path, "w") as f:
        f.write(buf.getvalue())
    print("stats      ok: %d lines" % buf.getvalue().count("\n"))


if __name__ == "__main__":
    if not os.path.isdir(REF):
        raise SystemExit("needs the reference at %s (build container only)" % REF)
    torch.set_num_threads(1)
    run_metrics()
    run_labels()
    run_stats()
    if "--metrics-only" not in sys.argv:
        for c in CASES:
            run_case(c)
