"""Helpers for the GPU parity tests: layout conversion between the reference's (F,N) matrices and the frame-major
device layout of libdvae_b200, and fixture plumbing."""
import ctypes as C

import numpy as np
import torch

from dvae_b200 import _lib
from dvae_b200.engine import InjectedDraws, McemConfig, McemEngine, RaggedBatch, VaeWeights, _ld_for, _p, _stream

DEV = "cuda:0"


def fm(a, ld=None, dtype=torch.float32):
    """(F, N) host matrix -> [N][ld] device tensor (zero padded)."""
    a = np.asarray(a)
    F, N = a.shape
    ld = ld or _ld_for(F)
    t = torch.zeros((N, ld), dtype=dtype, device=DEV)
    t[:, :F] = torch.from_numpy(np.ascontiguousarray(a.T)).to(DEV).to(dtype)
    return t


def unfm(t, F):
    """[N][ld] device tensor -> (F, N) numpy."""
    return np.ascontiguousarray(t[:, :F].t().cpu().numpy())


def relerr(a, b):
    dt = np.complex128 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else np.float64
    a, b = np.asarray(a, dt), np.asarray(b, dt)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


def golden_draws(g, variant_schedule, n_chains=1):
    """Split a fixture's recorded draw list into the engine's InjectedDraws (batch of one utterance)."""
    (kE, bE), (kW, bW) = variant_schedule
    d = g.draws
    W0 = d[0][1].T[None]                       # rand(F,K) -> [1][K][F]
    H0 = d[1][1].T                             # rand(K,N) -> [N][K]
    pos = 4                                    # two encoder draws are consumed and dropped
    eps, u = [], []
    for call in range(g.niter + 1):
        n_it = (kE + bE) if call < g.niter else (kW + bW)
        e = np.stack([d[pos + 2 * i][1].T for i in range(n_it)])          # randn(L,N) -> [it][N][L]
        uu = np.stack([d[pos + 2 * i + 1][1] for i in range(n_it)])       # rand(N)
        pos += 2 * n_it
        eps.append(torch.from_numpy(np.ascontiguousarray(e)))
        u.append(torch.from_numpy(np.ascontiguousarray(uu)))
    assert pos == len(d)
    return InjectedDraws(torch.from_numpy(np.ascontiguousarray(W0)), torch.from_numpy(np.ascontiguousarray(H0)), eps, u)


def engine_for(g, schedule, sampler="fp32", niter=None):
    (kE, bE), (kW, bW) = schedule
    w = VaeWeights(g.sd, g.variant, torch.device(DEV))
    cfg = McemConfig(niter=g.niter if niter is None else niter, keep_E=kE, burn_E=bE, keep_WF=kW, burn_WF=bW,
                     var_rw=g.var_RW, nmf_rank=g.K, eps=g.eps, sampler=sampler)
    eng = McemEngine(w, cfg, DEV)
    X = fm(g.X, dtype=torch.complex64)
    P = torch.zeros((X.shape[0], X.shape[1]), dtype=torch.float32, device=DEV)
    _lib.call("dvae_power", _p(X), _p(P), X.numel(), _stream())
    y = None if g.y is None else torch.from_numpy(np.ascontiguousarray(g.y.T)).to(DEV)
    batch = RaggedBatch([g.X.shape[1]], DEV)
    return eng, X, P, y, batch
