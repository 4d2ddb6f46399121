"""Seeded synthetic inputs for the enhancement path.

No dataset or checkpoint can be fetched in this environment, so every test and
benchmark runs on synthetic utterances of the reference's shape:

* utterance ``u``: 3 s at 16 kHz; a harmonic "speech-like" source with a 3 Hz
  syllabic envelope plus 1-pole low-passed Gaussian noise mixed at a given SNR
  with the recipe of ``scripts/create_test_set.py:101-116`` (power-matched noise
  gain, then a joint peak normalisation of speech / noise / mixture);
* labels ``y``: energy VAD with the rule of ``packages/processing/target.py:51-53``;
* weights: the reference containers' own initialisation (xavier-normal weights,
  zero biases, ``packages/models/models.py:137-141``) under a fixed seed, with the
  decoder's output bias shifted to ``log(mean |X|^2)`` so Metropolis-Hastings
  acceptance rates are non-degenerate.

Pure numpy / CPU torch; nothing here touches CUDA.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.signal import lfilter

FS = 16000
N_FFT = 1024
HOP = 256
F_BINS = N_FFT // 2 + 1


def synth_utterance(u: int, seconds: float = 3.0, snr_db: float | None = None, fs: int = FS):
    """Return ``(x, s, n)`` float32 arrays of ``round(seconds*fs)`` samples for utterance index ``u``.

    ``x = s + n`` exactly (before float32 rounding), as for the reference's ``_s/_n/_x.wav`` triplets.
    """
    rng = np.random.default_rng(1000 + int(u))
    T = int(round(seconds * fs))
    t = np.arange(T, dtype=np.float64) / fs
    f0 = rng.uniform(90.0, 250.0)
    phase = rng.uniform(0.0, 2 * np.pi, size=11)
    speech = np.zeros(T)
    for k in range(1, 12):
        speech += np.sin(2 * np.pi * k * f0 * t + phase[k - 1]) / k
    env = (0.5 + 0.5 * np.sin(2 * np.pi * 3.0 * t + rng.uniform(0, 2 * np.pi))) ** 2
    speech *= env
    speech = speech / np.max(np.abs(speech))

    white = rng.standard_normal(T)
    noise = lfilter([0.1], [1.0, -0.9], white)  # 1-pole low-pass colouring

    if snr_db is None:
        snr_db = float((-5, 0, 5, 10)[int(u) % 4])
    sp = np.sum(speech ** 2)
    npow = np.sum(noise ** 2)
    noise = noise * np.sqrt(sp * 10.0 ** (-snr_db / 10.0) / npow)
    norm = np.max(np.abs(np.concatenate([speech, noise, speech + noise])))
    x = (speech + noise) / norm
    return x.astype(np.float32), (speech / norm).astype(np.float32), (noise / norm).astype(np.float32)


def synth_batch(u0: int, count: int, seconds: float = 3.0):
    """Stack ``count`` utterances starting at global index ``u0`` -> ``(x, s)`` each ``(count, T)`` float32."""
    xs, ss = [], []
    for u in range(u0, u0 + count):
        x, s, _ = synth_utterance(u, seconds)
        xs.append(x)
        ss.append(s)
    return np.stack(xs), np.stack(ss)


def padded_length(T: int, fs: int = FS, wlen_sec: float = 64e-3, hop_percent: float = 0.25) -> int:
    """Length after the end-padding rule of ``packages/processing/stft.py:45-50``."""
    n_fft = int(wlen_sec * fs)
    hop = int(hop_percent * n_fft)
    utt_len = T / fs
    if math.ceil(utt_len / wlen_sec / hop_percent) != int(utt_len / wlen_sec / hop_percent):
        return T + hop
    return T


def num_frames(T: int, n_fft: int = N_FFT, hop: int = HOP, fs: int = FS) -> int:
    """Frames of the ``center=False`` STFT of a ``T``-sample signal after end padding."""
    Tp = padded_length(T, fs, n_fft / fs, hop / n_fft)
    return 1 + (Tp - n_fft) // hop


def energy_vad(s: np.ndarray, n_fft: int = N_FFT, hop: int = HOP, threshold: float = 1.70, fs: int = FS):
    """``(1, N)`` float32 speech-activity labels, rule of ``target.py:51-53`` with ``center=False``."""
    Tp = padded_length(len(s), fs, n_fft / fs, hop / n_fft)
    y = np.zeros(Tp, dtype=np.float64)
    y[: len(s)] = s
    N = 1 + (Tp - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(N)[:, None]
    power = np.sum(y[idx] ** 2, axis=1)
    vad = power > (10.0 ** threshold) * np.min(power)
    return vad.astype(np.float32)[None]


def xavier_state_dict(variant: str, x_dim: int, z_dim: int, h_dim, y_dim: int = 0, seed: int = 1234,
                      out_bias: float = 0.0):
    """Random weights with the reference's key layout and init (``models.py:137-141``), as numpy float32.

    ``variant``: ``"M1"`` (``VariationalAutoencoder``), ``"M2"`` (``DeepGenerativeModel``: encoder sees
    ``[x;y]``), ``"M2v3"`` (``DeepGenerativeModel_v3``: encoder sees ``x`` only). The decoder sees ``[z;y]``
    for both M2 flavours. Returned keys: ``encoder.hidden.{i}.{weight,bias}``,
    ``encoder.sample.{mu,log_var}.{weight,bias}``, ``decoder.hidden.{i}.{weight,bias}``,
    ``decoder.reconstruction.{weight,bias}``.
    """
    rng = np.random.default_rng(seed)
    h_dim = list(h_dim)
    enc_in = x_dim + (y_dim if variant == "M2" else 0)
    dec_in = z_dim + (y_dim if variant != "M1" else 0)

    def lin(fan_in, fan_out):
        std = math.sqrt(2.0 / (fan_in + fan_out))
        return (rng.standard_normal((fan_out, fan_in)) * std).astype(np.float32), np.zeros(fan_out, np.float32)

    sd = {}
    dims = [enc_in] + h_dim
    for i in range(len(h_dim)):
        sd[f"encoder.hidden.{i}.weight"], sd[f"encoder.hidden.{i}.bias"] = lin(dims[i], dims[i + 1])
    for head in ("mu", "log_var"):
        sd[f"encoder.sample.{head}.weight"], sd[f"encoder.sample.{head}.bias"] = lin(h_dim[-1], z_dim)
    rh = list(reversed(h_dim))
    dims = [dec_in] + rh
    for i in range(len(rh)):
        sd[f"decoder.hidden.{i}.weight"], sd[f"decoder.hidden.{i}.bias"] = lin(dims[i], dims[i + 1])
    w, b = lin(rh[-1], x_dim)
    sd["decoder.reconstruction.weight"] = w
    sd["decoder.reconstruction.bias"] = (b + np.asarray(out_bias, np.float32)).astype(np.float32)   # scalar or per-bin
    return sd


def speech_prior_bias(s, n_fft: int = N_FFT, hop: int = HOP):
    """Per-bin log of the mean clean-speech power spectrum of ``s`` (float64 numpy STFT, Hann, no centring).

    Used as the decoder's output bias so that randomly initialised weights behave like a (crude) speech model:
    the Wiener filter then separates speech from noise and SI-SDR comparisons are well conditioned.
    """
    s = np.asarray(s, np.float64)
    Tp = padded_length(len(s), FS, n_fft / FS, hop / n_fft)
    y = np.zeros(Tp)
    y[: len(s)] = s
    N = 1 + (Tp - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(N)[:, None]
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n_fft) / n_fft)
    P = np.abs(np.fft.rfft(y[idx] * w, axis=1)) ** 2
    return np.log(P.mean(axis=0) + 1e-10).astype(np.float32)

