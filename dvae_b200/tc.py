"""Host glue of the tcgen05 (BF16 tensor-core) sampler and decode: packed operands and the C ABI calls.

``mh_chain_tc`` / ``decode_tc`` are what ``McemEngine`` runs when ``McemConfig.sampler == "tc"``.  Everything here is
bookkeeping: building the decoder's shared-memory image once per model, re-tiling P once per batch and Vb once per
``sample_posterior`` call, and checking the kernel's status word.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


POLY_EX2 = 1            # DVAE_TC_POLY_EX2
POLY_EX2_ALL = 2        # DVAE_TC_POLY_EX2_ALL
POLY_EX2_LIMIT = 120.0  # DVAE_TC_POLY_EX2_LIMIT


def decoder_image(weights):
    """The UMMA-ready image of ``weights.dec`` (cached on the weights object)."""
    img = getattr(weights, "_tc_image", None)
    if img is None:
        lib = _lib.load()
        n = lib.dvae_tc_image_bytes(weights.dec.ref, weights.z_dim, weights.y_dim)
        if n < 0:
            _lib.check(-1, "dvae_tc_image_bytes")
        img = torch.empty(int(n), dtype=torch.uint8, device=weights.device)
        _lib.call("dvae_tc_pack_decoder", weights.dec.ref, weights.z_dim, weights.y_dim, _p(img), _stream())
        weights._tc_image = img
        # one-off range query: may the sampler use the polynomial 2^x (see DVAE_TC_POLY_EX2 in include/dvae_b200.h)?
        bound = C.c_float(0.0)
        _lib.call("dvae_tc_decoder_exponent_bound", weights.dec.ref, weights.z_dim, weights.y_dim, C.byref(bound), _stream())
        mode = os.environ.get("DVAE_TC_POLY", "1")          # 0: MUFU only, 1: half of the exponentials, 2: all of them (tc2 sampler)
        weights._tc_flags = ({"0": 0, "2": POLY_EX2 | POLY_EX2_ALL}.get(mode, POLY_EX2)) if bound.value < POLY_EX2_LIMIT else 0
    return img


def pack_rows(eng, name, src):
    """Frame-major [NT][ld] -> [tile][quad][128 chains][4] (see dvae_tc_pack_rows)."""
    b, C_ = eng.batch, eng.cfg.n_chains
    n = int(_lib.load().dvae_tc_packed_floats(b.NT * C_))
    dst = eng._get(name, (max(n, 4),))
    _lib.call("dvae_tc_pack_rows", _p(src), b.NT, C_, eng.F, eng.ld, _p(dst), _stream())
    eng.kernel_launches += 1
    return dst


def _status(eng):
    st = eng._buf.get("tc_status")
    if st is None:
        st = torch.zeros(1, dtype=torch.int32, device=eng.dev)
        eng._buf["tc_status"] = st
    return st


def check_status(eng):
    """Raise if any tensor-core kernel of this engine reported a pipeline timeout (synchronises)."""
    st = eng._buf.get("tc_status")
    if st is not None and int(st.item()) != 0:
        raise _lib.DvaeError("tcgen05 kernel reported a pipeline timeout (status %d): results are invalid" % int(st.item()))


def mh_chain_tc(eng, Zs, keep, burn, rng, a_trace):
    w, b, cfg = eng.w, eng.batch, eng.cfg
    img = decoder_image(w)
    v2 = w.z_dim in (16, 32) and w.y_dim <= 3 and os.environ.get("DVAE_TC_SAMPLER", "v2") != "v1"
    if not v2:
        if getattr(eng, "_Ppk_for", None) is not eng.P:
            eng._Ppk = pack_rows(eng, "Ppk", eng.P)
            eng._Ppk_for = eng.P
        Vbpk = pack_rows(eng, "Vbpk", eng.Vb)
        _lib.call("dvae_mh_chain_tc", w.dec.ref, _p(img), _p(eng._Ppk), _p(Vbpk), _p(eng.g), _p(eng.y), w.y_dim,
                  _p(b.frame_gid), _p(b.frame_idx), _p(eng.Z), _p(Zs), b.NT, w.z_dim, cfg.n_chains, burn, keep,
                  float(cfg.var_rw), C.byref(rng), _p(eng.n_accept), _p(a_trace), _p(_status(eng)), _stream())
        eng.kernel_launches += 1
        return
    # v2 reads its draws from global memory: injected ones as they are, Philox ones dumped by the generator kernel
    n_iter, chains = keep + burn, b.NT * cfg.n_chains
    if rng.eps:
        eps_ptr, u_ptr = C.c_void_p(rng.eps), C.c_void_p(rng.u)
    else:
        eps = eng._get("tc_eps", (n_iter * chains * w.z_dim,))
        u = eng._get("tc_u", (n_iter * chains,))
        _lib.call("dvae_rng_dump", C.byref(rng), _p(b.frame_gid), _p(b.frame_idx), b.NT, cfg.n_chains, w.z_dim, n_iter,
                  _p(eps), _p(u), _stream())
        eng.kernel_launches += 1
        eps_ptr, u_ptr = _p(eps), _p(u)
    gen = os.environ.get("DVAE_TC_SAMPLER", "v2")
    fn = "dvae_mh_chain_tc3" if (gen == "v3" and w.z_dim == 16) else ("dvae_mh_chain_tc4" if gen == "v4" else "dvae_mh_chain_tc2")
    nb = int(_lib.load().dvae_tc_packed_pv_bytes(chains))
    pv = eng._get("PVpk", (max(nb, 16),), torch.uint8)
    _lib.call("dvae_tc_pack_pv", w.dec.ref, _p(img), w.z_dim, w.y_dim, _p(eng.P), _p(eng.Vb), b.NT, cfg.n_chains, eng.F, eng.ld,
              _p(pv), _stream())
    eng.kernel_launches += 1
    with eng.stage("mh_kernel"):
        args = (w.dec.ref, _p(img), _p(pv), _p(eng.g), _p(eng.y), w.y_dim, _p(eng.Z), _p(Zs), b.NT, w.z_dim, cfg.n_chains,
                burn, keep, float(cfg.var_rw), eps_ptr, u_ptr, _p(eng.n_accept), _p(a_trace))
        if fn != "dvae_mh_chain_tc3":
            _lib.call(fn, *args, int(w._tc_flags), _p(_status(eng)), _stream())
        else:
            _lib.call(fn, *args, _p(_status(eng)), _stream())
    eng.kernel_launches += 1


def decode_tc(eng, x, x2, x2_row_div, out):
    """Vs rows for the kept samples ``x [rows][L]`` (+ labels) through the tensor-core decoder."""
    w = eng.w
    img = decoder_image(w)
    _lib.call("dvae_decode_tc", w.dec.ref, _p(img), _p(x), x.shape[0], w.z_dim, _p(x2), w.y_dim, max(1, x2_row_div), _p(out),
              out.stride(0), _p(_status(eng)), _stream())
    eng.kernel_launches += 1


def decode_wstat_tc(eng, Zs, Vs):
    """Kept-sample decode fused with the W-update statistics (dvae_decode_ws_tc); returns the statistics buffer."""
    from .engine import WS_PARTS
    w, b, cfg = eng.w, eng.batch, eng.cfg
    img = decoder_image(w)
    n = int(_lib.load().dvae_decode_ws_workspace_floats(b.B, WS_PARTS, eng.ld))
    ws = eng._get("wstat", (n,))
    _lib.call("dvae_decode_ws_tc", w.dec.ref, _p(img), _p(Zs), Zs.shape[1], w.z_dim, _p(eng.y), w.y_dim, _p(eng.P), _p(eng.Vb),
              _p(eng.g), _p(eng.H), cfg.nmf_rank, _p(b.fr_off), b.B, b.NT, eng.ld, _p(Vs), _p(ws), WS_PARTS, _p(_status(eng)),
              _stream())
    eng.kernel_launches += 1
    return ws


def decode_stats_tc(eng, Zs, Vs):
    """Warp-specialised decode + per-frame reciprocal sums (dvae_decode_stats_tc); returns the [2][NT][ld] statistics.

    Frames with more than 30 samples (multi-chain runs) are decoded in sample windows of 30 or 10 that accumulate into the
    same statistics (dvae_decode_stats_win_tc)."""
    w, b = eng.w, eng.batch
    img = decoder_image(w)
    st = eng._get("fstat", (2 * b.NT * eng.ld,))
    A1, A2 = st[: b.NT * eng.ld], st[b.NT * eng.ld:]
    R = Zs.shape[1]
    if R in (10, 30):
        _lib.call("dvae_decode_stats_tc", w.dec.ref, _p(img), _p(Zs), R, w.z_dim, _p(eng.y), w.y_dim, _p(eng.Vb),
                  _p(eng.g), b.NT, eng.ld, _p(Vs), _p(A1), _p(A2), _p(_status(eng)), _stream())
        eng.kernel_launches += 1
        return st
    win = 30 if R % 30 == 0 else 10
    for r0 in range(0, R, win):
        _lib.call("dvae_decode_stats_win_tc", w.dec.ref, _p(img), _p(Zs), R, r0, win, w.z_dim, _p(eng.y), w.y_dim, _p(eng.Vb),
                  _p(eng.g), b.NT, eng.ld, _p(Vs), _p(A1), _p(A2), 0 if r0 == 0 else 1, _p(_status(eng)), _stream())
        eng.kernel_launches += 1
    return st


def decode_a1_tc(eng, Zs, r0, R):
    """A1 = sum over the samples r0 .. r0+R of 1 / (g Vs + Vb), without storing Vs (dvae_decode_a1_tc); ``[NT][ld]``."""
    w, b = eng.w, eng.batch
    img = decoder_image(w)
    A1 = eng._get("wf_a1", (b.NT * eng.ld,))
    _lib.call("dvae_decode_a1_tc", w.dec.ref, _p(img), _p(Zs), Zs.shape[1], int(r0), int(R), w.z_dim, _p(eng.y), w.y_dim, _p(eng.Vb),
              _p(eng.g), b.NT, eng.ld, _p(A1), _p(_status(eng)), _stream())
    eng.kernel_launches += 1
    return A1

