"""``packages/metrics.py`` of the reference (packages/metrics.py:5-82) on the device.

Same names, arguments and return values as the reference for the per-utterance functions (numpy vectors in, Python floats /
numpy arrays out); the arithmetic runs in ``dvae_energy_ratios`` (six inner products in double, closed forms for the
energies) instead of forming ``s_target``, ``e_noise`` and ``e_art``.  ``energy_ratios_batch`` is the batched entry the
enhancement path uses: device tensors in the layout of ``Enhancer.run_device``, nothing leaves the GPU but 24 bytes per
utterance.  CUDA only: without the library or a device the calls raise ``DvaeError``.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from ..engine import _p, _stream


def energy_ratios_batch(s_hat, s, n, off, lens):
    """SI-SDR, SI-SIR, SI-SAR of a ragged batch.

    ``s_hat``, ``s``, ``n``: 1-D float32 CUDA tensors (``n`` may be None); utterance ``u`` is ``[off[u], off[u]+lens[u])``.
    Returns a float64 CUDA tensor ``[B][3]`` (SI-SIR / SI-SAR are NaN without ``n``).
    """
    if not (s_hat.is_cuda and s.is_cuda and off.is_cuda and lens.is_cuda):
        raise _lib.DvaeError("energy_ratios_batch needs CUDA tensors")
    B = int(off.numel())
    out = torch.empty((B, 3), dtype=torch.float64, device=s_hat.device)
    with torch.cuda.device(s_hat.device):
        _lib.call("dvae_energy_ratios", _p(s_hat.contiguous()), _p(s.contiguous()), _p(n.contiguous()) if n is not None else None,
                  _p(off.to(torch.int64)), _p(lens.to(torch.int32)), B, _p(out), _stream())
    return out


def _one(s_hat, s, n, device):
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda" or not torch.cuda.is_available():
        raise _lib.DvaeError("dvae_b200 metrics run on a CUDA device only")
    T = len(s_hat)
    if len(s) != T or (n is not None and len(n) != T):
        raise ValueError("signals must have the same length")
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)      # noqa: E731
    off = torch.zeros(1, dtype=torch.int64, device=dev)
    lens = torch.full((1,), T, dtype=torch.int32, device=dev)
    return energy_ratios_batch(to(s_hat), to(s), None if n is None else to(n), off, lens)[0].cpu().numpy()


def energy_ratios(s_hat, s, n, device=None):
    """``(si_sdr, si_sir, si_sar)`` in dB (packages/metrics.py:39-60)."""
    r = _one(s_hat, s, n, device)
    return float(r[0]), float(r[1]), float(r[2])


def si_sdr_leroux(s_hat, s, device=None):
    """SI-SDR in dB (packages/metrics.py:62-82)."""
    return float(_one(s_hat, s, None, device)[0])


def si_sdr_components(s_hat, s, n):
    """``(s_target, e_noise, e_art)`` (packages/metrics.py:12-37); three axpy's on the host, kept for API completeness."""
    s_hat, s, n = (np.asarray(a) for a in (s_hat, s, n))
    s_target = (np.dot(s_hat, s) / np.dot(s, s)) * s
    e_noise = (np.dot(s_hat, n) / np.dot(n, n)) * n
    return s_target, e_noise, s_hat - s_target - e_noise


# ---------------------------------------------------------------------------------------------- statistics table (host)
def mean_confidence_interval(data, confidence=0.95, round=3):
    """Mean and half-width of the Student-t confidence interval, both rounded to 3 decimals (packages/metrics.py:5-10)."""
    import scipy.stats
    a = np.asarray(data, dtype=np.float64)
    half = scipy.stats.sem(a) * scipy.stats.t.ppf((1.0 + confidence) / 2.0, len(a) - 1)
    return np.round(np.mean(a), 3), np.round(half, 3)


def compute_stats(metrics_keys, all_metrics, model_data_dir, confidence, all_snr_db=None, all_noise_types=None, all_speakers=None,
                  all_noise_stationarities=None):
    """The evaluation table of the reference (packages/metrics.py:84-167): average and confidence interval of every metric
    over all utterances, then per input SNR / noise type / noise stationarity / speaker when those lists are given.  Prints
    the same tables (same headings and column format) and additionally RETURNS them:
    ``{"all": {key: {"avg", "+/-"}}, "snr": {value: {...}}, "noise_type": {...}, "noise_stationarity": {...}, "speaker": {...}}``.
    ``all_metrics`` is a list of per-utterance tuples in the order of ``metrics_keys`` (``scripts/run_metrics.py:117-129``)."""
    cols = {key: np.asarray([row[i] for row in all_metrics], dtype=np.float64) for i, key in enumerate(metrics_keys)}

    def table(select):
        print("{:<10} {:<10} {:<10}".format('METRIC', 'AVERAGE', 'CONF. INT.'))
        out = {}
        for key, col in cols.items():
            m, h = mean_confidence_interval(col[select], confidence=confidence)
            out[key] = {'avg': m, '+/-': h}
            print("{:<10} {:<10} {:<10}".format(key, m, h))
        print('\n')
        return out

    n = len(all_metrics)
    result = {"all": table(np.ones(n, dtype=bool))}
    groups = (("snr", all_snr_db, lambda v: 'Input SNR = {:.2f}'.format(v)),
              ("noise_type", all_noise_types, lambda v: 'Noise type = {}'.format(v)),
              ("noise_stationarity", all_noise_stationarities, lambda v: 'Noise type = {}'.format(v)),      # heading as in the reference
              ("speaker", all_speakers, lambda v: 'Speaker = {}'.format(v)))
    for name, labels, heading in groups:
        if labels is None:
            continue
        labels = np.asarray(labels)
        result[name] = {}
        for v in (np.unique(labels) if name == "snr" else sorted(set(labels.tolist()))):
            print(heading(v))
            result[name][v] = table(labels == v)
    return result
