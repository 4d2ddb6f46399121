"""Device label front end against the oracle and the reference's golden labels."""
import os

import numpy as np
import pytest
import torch

from dvae_b200 import _lib, synth
from dvae_b200.engine import RaggedBatch, stft_batch
from dvae_b200.packages.processing import target as dtarget
from oracle import stft_np
from tests.gpu_util import DEV

pytestmark = pytest.mark.gpu
KW = dict(fs=16000, wlen_sec=64e-3, hop_percent=0.25, center=False, pad_at_end=True)


def test_labels_match_reference_golden():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "labels.npz"))
    s = g["s"]
    S = stft_np.stft(s, win="hann", **KW)
    vad = dtarget.clean_speech_VAD(s, device=DEV, **KW)
    assert vad.shape == g["vad"].shape and vad.dtype == np.float32 and np.array_equal(vad, g["vad"])
    ibm = dtarget.clean_speech_IBM(S, device=DEV)
    assert ibm.shape == S.shape and ibm.dtype == np.float32
    # the reference compares float32 logarithms, the kernel the equivalent magnitudes in double: bins within one float32 ulp
    # of the threshold may differ
    assert np.mean(ibm != g["ibm"]) < 2e-4
    nr = dtarget.noise_robust_clean_speech_IBM(s, S, device=DEV, **KW)
    assert np.mean(nr != g["nr"]) < 2e-4 and np.all(nr[:, vad[0] == 0] == 0)


def test_vad_ragged_batch_matches_oracle():
    lens = [48000, 16384, 1024, 30001]
    sigs = [synth.synth_utterance(u, seconds=3.0, snr_db=0.0)[1][:t].astype(np.float32) for u, t in enumerate(lens)]
    off = np.concatenate([[0], np.cumsum([(t + 1) // 2 * 2 for t in lens])]).astype(np.int64)
    flat = np.zeros(int(off[-1]), np.float32)
    for u, x in enumerate(sigs):
        flat[off[u]:off[u] + lens[u]] = x
    nfr = [synth.num_frames(t) for t in lens]
    batch = RaggedBatch(nfr, DEV)
    x_dev = torch.from_numpy(flat).to(DEV)
    x_off = torch.from_numpy(off[:-1].copy()).to(DEV)
    x_len = torch.tensor(lens, dtype=torch.int32, device=DEV)
    vad = dtarget.vad_batch(x_dev, x_off, x_len, batch).cpu().numpy()
    X, _ = stft_batch(x_dev, x_off, x_len, batch)
    ibm = dtarget.ibm_batch(X, batch, 513).cpu().numpy()
    for u, x in enumerate(sigs):
        ref = stft_np.clean_speech_vad(x)[0]
        got = vad[batch.fr_off_host[u]:batch.fr_off_host[u + 1]]
        assert got.shape == ref.shape and np.array_equal(got, ref), u
        Sref = stft_np.stft(x, win="hann", **KW)
        mref = stft_np.clean_speech_ibm(Sref).T
        mgot = ibm[batch.fr_off_host[u]:batch.fr_off_host[u + 1], :513]
        assert np.mean(mgot != mref) < 1e-3, (u, np.mean(mgot != mref))      # the device STFT differs from the float64 one by 1e-7
        assert np.all(ibm[:, 513:] == 0)


def test_label_shim_errors():
    with pytest.raises(NotImplementedError):
        dtarget.clean_speech_VAD(np.zeros(4096, np.float32), fs=16000, wlen_sec=64e-3, center=True, device=DEV)
    with pytest.raises(ValueError):
        dtarget.clean_speech_VAD(np.zeros(4096, np.float32), fs=16000, wlen_sec=50.01e-3, center=False, device=DEV)
    with pytest.raises(_lib.DvaeError):
        dtarget.clean_speech_IBM(np.ones((513, 3), np.complex64), device="cpu")
