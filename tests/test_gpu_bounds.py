"""Out-of-bounds guards: compute-sanitizer is not available on the GPU pool, so the kernels that write ragged / padded
buffers are checked with sentinels instead - nothing outside the documented output region may change."""
import numpy as np
import pytest
import torch

from dvae_b200 import synth
from dvae_b200.engine import McemConfig, McemEngine, RaggedBatch, VaeWeights, istft_batch, stft_batch
from tests.gpu_util import DEV

pytestmark = pytest.mark.gpu
SENT = -12345.5


def test_stft_istft_write_only_their_region():
    lens = [48000, 5000, 1024, 30001, 777 + 1024]
    gap = 64                                                    # sentinel samples between utterances and at both ends
    off = np.zeros(len(lens), np.int64)
    pos = gap
    for u, t in enumerate(lens):
        off[u] = pos
        pos += (t + 1) // 2 * 2 + gap
    total = pos
    x = torch.full((total,), 0.25, device=DEV)
    for u, t in enumerate(lens):
        x[off[u]:off[u] + t] = torch.from_numpy(synth.synth_utterance(u, 3.0)[0][:t].astype(np.float32)).to(DEV)
    nfr = [synth.num_frames(t) for t in lens]
    batch = RaggedBatch(nfr, DEV)
    x_off = torch.from_numpy(off).to(DEV)
    x_len = torch.tensor(lens, dtype=torch.int32, device=DEV)
    X, P = stft_batch(x, x_off, x_len, batch)
    # the kernel must not write the pad columns 513..519: fill them, run again into the same buffers, compare
    Xs = torch.view_as_real(X)
    Xs[:, 513:, :] = SENT
    P[:, 513:] = SENT
    from dvae_b200 import _lib
    from dvae_b200.engine import _p, _stream
    _lib.call("dvae_stft_f32", _p(x), _p(x_off), _p(x_len), batch.B, _p(X), _p(P), _p(batch.fr_off), batch.NT, 1024, 256, X.shape[1],
              _stream())
    assert bool((Xs[:, 513:, :] == SENT).all()) and bool((P[:, 513:] == SENT).all())
    Xs[:, 513:, :] = 0.0
    y = torch.full((total,), SENT, device=DEV)
    istft_batch(X, batch, x_off, x_len, total, max(lens), out=y)
    mask = torch.zeros(total, dtype=torch.bool, device=DEV)
    for u, t in enumerate(lens):
        mask[off[u]:off[u] + t] = True
    assert bool((y[~mask] == SENT).all()), "istft wrote outside an utterance"
    assert bool((y[mask] != SENT).all()), "istft left output samples unwritten"
    for u, t in enumerate(lens):                                  # and the round trip holds away from the first hop
        a, b = y[off[u] + 800:off[u] + t - 800], x[off[u] + 800:off[u] + t - 800]
        if a.numel():
            assert float((a - b).abs().max()) < 1e-5


@pytest.mark.parametrize("variant,keep", [("M1", 30), ("M2v3", 10)])
def test_estep_decode_keeps_pad_columns_and_tail_rows(variant, keep):
    """decode_stats writes Vs[:, :, :513], A1 / A2[:, :513] of the batch's frames and nothing else."""
    N = [37, 5, 21]                                               # 63 frames: not a multiple of the 4- or 12-frame tile
    rng = np.random.default_rng(1)
    NT = sum(N)
    P = torch.tensor(rng.gamma(1.0, 0.05, size=(NT, 520)).astype(np.float32)).to(DEV)
    X = torch.zeros((NT, 520), dtype=torch.complex64, device=DEV)
    y_dim = 0 if variant == "M1" else 1
    sd = synth.xavier_state_dict(variant, 513, 16, [128, 128], y_dim, seed=3, out_bias=float(np.log(0.05)))
    w = VaeWeights(sd, variant, DEV)
    eng = McemEngine(w, McemConfig(niter=1, keep_E=keep, burn_E=4, keep_WF=keep, burn_WF=4, sampler="tc"), DEV)
    y = None if y_dim == 0 else torch.tensor(rng.integers(0, 2, size=(NT, 1)).astype(np.float32)).to(DEV)
    eng.init_parameters(X, P, RaggedBatch(N, DEV), y)
    from dvae_b200 import tc
    Zs = eng.sample_posterior(keep, 4)
    flat = torch.full(((NT + 3) * keep, 520), SENT, device=DEV)
    Vs = flat[: NT * keep].view(NT, keep, 520)
    tc.decode_stats_tc(eng, Zs, Vs)
    tc.check_status(eng)
    assert bool((Vs[:, :, 513:] == SENT).all()), "pad columns of Vs were written"
    assert bool((flat[NT * keep:] == SENT).all()), "rows beyond the batch were written"
    assert bool(torch.isfinite(Vs[:, :, :513]).all()) and bool((Vs[:, :, :513] > 0).all())
