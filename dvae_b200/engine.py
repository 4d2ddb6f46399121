"""Batched MCEM VAE-NMF enhancement engine: Python/PyTorch host code over libdvae_b200's C ABI.

PyTorch is plumbing here (device memory, streams); every arithmetic step of the path is a CUDA kernel of
``libdvae_b200.so`` reached through ``dvae_b200._lib``.  The engine processes a *ragged batch* of utterances at once
(the reference handles one utterance per process, ``scripts/evaluate_ntcd_M1.py:190-214``):

    STFT -> |X|^2 -> encoder mean (chain start) -> niter x [ MH E-step -> decode kept samples -> M-step (W,H,g,cost) ]
         -> long MH chain -> Wiener masks -> S_hat, N_hat -> ISTFT

Reference behaviour reproduced: chain start = encoder posterior MEAN (``_, Z, _ = encoder(..)``,
mcem.py:200); Vb kept un-normalised after the W/H renormalisation (Q4); per-variant MH schedules incl. the M1
argument shift (Q1, applied by the ``MCEM_M1`` shim, the engine takes explicit ``(keep, burn)`` pairs).
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import os
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from .synth import FS, HOP, N_FFT, num_frames

LD_ALIGN = 8


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _on_device(fn):
    """Run a method with the object's CUDA device current: the kernels of libdvae_b200 launch on the calling thread's current
    device and ``_stream()`` is that device's current stream, while the reference's scripts hand a device INDEX to every
    worker and never call ``set_device`` (scripts/evaluate_ntcd_M1.py:69,252)."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *a, **kw):
        with torch.cuda.device(self.dev):
            return fn(self, *a, **kw)
    return wrapper


def _require_cuda(device):
    if not torch.cuda.is_available():
        raise _lib.DvaeError("dvae_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    _lib.load()
    dev = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
    if dev.type != "cuda":
        raise _lib.DvaeError("dvae_b200 runs on CUDA devices only (got %r)" % (device,))
    return dev


# --------------------------------------------------------------------------------------------- weights
class PackedMlp:
    """Transposed FP32 weights of a tanh MLP on the device + the host-side ``DvaeMlp`` struct that points at them."""

    def __init__(self, layers: Sequence[tuple], device):
        if not 1 <= len(layers) <= _lib.MAX_LAYERS:
            raise ValueError("an MLP needs 1..%d linear layers" % _lib.MAX_LAYERS)
        self.wt, self.bias, dims = [], [], []
        for w, b in layers:
            w = torch.as_tensor(np.asarray(w) if not torch.is_tensor(w) else w).detach().to(torch.float32)
            b = torch.as_tensor(np.asarray(b) if not torch.is_tensor(b) else b).detach().to(torch.float32)
            if dims and dims[-1] != w.shape[1]:
                raise ValueError("layer input %d does not match previous output %d" % (w.shape[1], dims[-1]))
            if not dims:
                dims.append(int(w.shape[1]))
            dims.append(int(w.shape[0]))
            self.wt.append(w.t().contiguous().to(device))
            self.bias.append(b.contiguous().to(device))
        self.dims = dims
        self.struct = _lib.DvaeMlp()
        self.struct.n_layers = len(layers)
        for i, d in enumerate(dims):
            self.struct.dims[i] = d
        for i in range(len(layers)):
            self.struct.wt[i] = self.wt[i].data_ptr()
            self.struct.bias[i] = self.bias[i].data_ptr()

    @property
    def ref(self):
        return C.byref(self.struct)

    @property
    def in_dim(self):
        return self.dims[0]

    @property
    def out_dim(self):
        return self.dims[-1]


def _sd_get(sd, key):
    v = sd[key]
    return v.detach().cpu() if torch.is_tensor(v) else torch.as_tensor(np.asarray(v))


class VaeWeights:
    """Device copies of one VAE's encoder / decoder in the layout the kernels want.

    ``sd`` uses the reference's state_dict keys: ``encoder.hidden.{i}``, ``encoder.sample.{mu,log_var}``,
    ``decoder.hidden.{i}``, ``decoder.reconstruction``.  ``variant``: "M1" | "M2" | "M2v2" | "M2v3".
    """

    def __init__(self, sd, variant: str, device, prefix: str = ""):
        if variant not in ("M1", "M2", "M2v2", "M2v3"):
            raise ValueError("unknown variant %r" % (variant,))
        self.variant = variant
        self.device = device

        def layers(stem):
            out, i = [], 0
            while f"{prefix}{stem}.hidden.{i}.weight" in sd:
                out.append((_sd_get(sd, f"{prefix}{stem}.hidden.{i}.weight"), _sd_get(sd, f"{prefix}{stem}.hidden.{i}.bias")))
                i += 1
            return out

        enc_h = layers("encoder")
        mu = (_sd_get(sd, f"{prefix}encoder.sample.mu.weight"), _sd_get(sd, f"{prefix}encoder.sample.mu.bias"))
        lv = (_sd_get(sd, f"{prefix}encoder.sample.log_var.weight"), _sd_get(sd, f"{prefix}encoder.sample.log_var.bias"))
        self.enc_mu = PackedMlp(enc_h + [mu], device)
        self.enc_lv = PackedMlp(enc_h + [lv], device)
        dec_h = layers("decoder")
        rec = (_sd_get(sd, f"{prefix}decoder.reconstruction.weight"), _sd_get(sd, f"{prefix}decoder.reconstruction.bias"))
        self.dec = PackedMlp(dec_h + [rec], device)
        self.z_dim = int(mu[0].shape[0])
        self.x_dim = self.dec.out_dim
        self.y_dim = self.dec.in_dim - self.z_dim
        # Tensor-core kernels carry at most three labels in their layer-1 operand.  Wider label vectors (the IBM-conditioned
        # M2 model: y_dim = 513, scripts/evaluate_ntcd_M2.py:66-73) are folded into a per-frame layer-1 bias: the decoder is
        # packed without its label columns and without b1, and W1[:, L:] y + b1 is evaluated once per batch.
        self.tc_label_bias = self.y_dim > 3
        if self.tc_label_bias:
            w1, b1 = dec_h[0]
            self.dec_tc = PackedMlp([(w1[:, :self.z_dim], torch.zeros_like(b1))] + dec_h[1:] + [rec], device)
            self.ybias_mlp = PackedMlp([(w1[:, self.z_dim:], b1)], device)
        else:
            self.dec_tc = self.dec
        self.tc_y_dim = 0 if self.tc_label_bias else self.y_dim
        self.enc_takes_y = self.enc_mu.in_dim == self.x_dim + self.y_dim and self.y_dim > 0
        if (variant == "M1") != (self.y_dim == 0):
            raise ValueError("variant %s does not match a decoder with %d label inputs" % (variant, self.y_dim))
        if variant == "M2" and not self.enc_takes_y:
            raise ValueError("M2 expects an encoder over [x; y]")
        if variant in ("M2v2", "M2v3") and self.enc_mu.in_dim != self.x_dim:
            raise ValueError("%s expects an encoder over x only" % variant)


# --------------------------------------------------------------------------------------------- ragged batch
class RaggedBatch:
    """Frame bookkeeping of a batch: prefix offsets and the per-frame index maps the kernels consume."""

    def __init__(self, n_frames: Sequence[int], device, utt_ids: Optional[Sequence[int]] = None):
        n = np.asarray(n_frames, dtype=np.int64)
        if n.ndim != 1 or len(n) == 0 or np.any(n < 0):
            raise ValueError("n_frames must be a non-empty list of non-negative counts")
        self.B = len(n)
        self.n_frames = n
        off = np.zeros(self.B + 1, np.int64)
        np.cumsum(n, out=off[1:])
        self.NT = int(off[-1])
        self.max_frames = int(n.max())
        ids = np.arange(self.B, dtype=np.int32) if utt_ids is None else np.asarray(utt_ids, np.int32)
        if len(ids) != self.B:
            raise ValueError("utt_ids length mismatch")
        self.fr_off_host = off
        self.fr_off = torch.from_numpy(off).to(device)
        self.utt_ids = torch.from_numpy(ids).to(device)
        local = np.repeat(np.arange(self.B, dtype=np.int32), n)
        self.frame_utt = torch.from_numpy(local).to(device)
        self.frame_gid = torch.from_numpy(ids[local] if self.NT else np.zeros(0, np.int32)).to(device)
        idx = (np.arange(self.NT, dtype=np.int64) - off[:-1][local]).astype(np.int32) if self.NT else np.zeros(0, np.int32)
        self.frame_idx = torch.from_numpy(idx).to(device)
        self._segments = None
        self._device = device

    def segments(self, tile: int = 128, n_chains: int = 1):
        """Segment tables of the fused W-update reduction (dvae_vst_w_partials): a segment is a maximal run of frames inside
        one ``tile``-frame tile AND one utterance.  Returns device tensors ``(seg_start [S+1] int64, tile_seg [n_tiles+1]
        int32, utt_seg [B+1] int32)`` and ``S``; utterance ``u`` owns the segments ``utt_seg[u] .. utt_seg[u+1]``.  With
        ``n_chains`` chains per frame the axis is the chain row ``frame * n_chains + chain`` (``tile`` rows per tile)."""
        key = (tile, n_chains)
        if self._segments is None or self._segments[0] != key:
            rows, off = self.NT * n_chains, self.fr_off_host * n_chains
            n_tiles = (rows + tile - 1) // tile
            cuts = np.unique(np.concatenate([np.arange(0, n_tiles + 1, dtype=np.int64) * tile, off]))
            cuts = cuts[cuts <= rows]
            if len(cuts) == 0 or cuts[-1] != rows:
                cuts = np.append(cuts, rows)
            S = len(cuts) - 1
            tile_seg = np.searchsorted(cuts[:-1], np.arange(0, n_tiles + 1, dtype=np.int64) * tile, side="left").astype(np.int32)
            utt_seg = np.searchsorted(cuts[:-1], off, side="left").astype(np.int32)
            dev = self._device
            self._segments = (key, (torch.from_numpy(cuts.astype(np.int64)).to(dev), torch.from_numpy(tile_seg).to(dev),
                                    torch.from_numpy(utt_seg).to(dev), S))
        return self._segments[1]


# --------------------------------------------------------------------------------------------- STFT / ISTFT
def _ld_for(F: int) -> int:
    return (F + LD_ALIGN - 1) // LD_ALIGN * LD_ALIGN


def stft_batch(x_flat: torch.Tensor, x_off: torch.Tensor, x_len: torch.Tensor, batch: RaggedBatch, n_fft=N_FFT, hop=HOP,
               want_power=True):
    """STFT of a concatenated float32 device signal -> ``(X [NT][ld] complex64, P [NT][ld] float32 | None)``."""
    ld = _ld_for(n_fft // 2 + 1)
    X = torch.empty((batch.NT, ld), dtype=torch.complex64, device=x_flat.device)
    P = torch.empty((batch.NT, ld), dtype=torch.float32, device=x_flat.device) if want_power else None
    with torch.cuda.device(x_flat.device):
        _lib.call("dvae_stft_f32", _p(x_flat), _p(x_off), _p(x_len), batch.B, _p(X), _p(P), _p(batch.fr_off), batch.NT,
                  n_fft, hop, ld, _stream())
    return X, P


def istft_batch(X: torch.Tensor, batch: RaggedBatch, y_off: torch.Tensor, y_len: torch.Tensor, total_len: int,
                max_len: int, n_fft=N_FFT, hop=HOP, out: Optional[torch.Tensor] = None):
    """ISTFT of ``X [NT][ld]`` into a concatenated float32 signal of ``total_len`` samples."""
    y = out if out is not None else torch.empty(total_len, dtype=torch.float32, device=X.device)
    with torch.cuda.device(X.device):
        _lib.call("dvae_istft_f32", _p(X), _p(batch.fr_off), batch.B, _p(y), _p(y_off), _p(y_len), int(max_len), n_fft, hop,
                  X.shape[1], _stream())
    return y


def istft_masked_batch(X: torch.Tensor, mask: torch.Tensor, batch: RaggedBatch, y_off: torch.Tensor, y_len: torch.Tensor,
                       total_len: int, max_len: int, mask_scale: float = 1.0, n_fft=N_FFT, hop=HOP,
                       out: Optional[torch.Tensor] = None):
    """ISTFT of ``(mask * mask_scale) * X`` (real ``mask [NT][ld]``): the Wiener filter of mcem.py:176-177 applied while the
    frames are loaded, bit-identical to ``dvae_wiener_apply`` + ``istft_batch`` without materialising the product."""
    if mask.shape != X.shape or mask.dtype != torch.float32 or not mask.is_contiguous():
        raise ValueError("mask must be a contiguous float32 tensor of the spectrum's shape")
    y = out if out is not None else torch.empty(total_len, dtype=torch.float32, device=X.device)
    with torch.cuda.device(X.device):
        _lib.call("dvae_istft_masked_f32", _p(X), _p(mask), float(mask_scale), _p(batch.fr_off), batch.B, _p(y), _p(y_off), _p(y_len),
                  int(max_len), n_fft, hop, X.shape[1], _stream())
    return y


def mlp_forward(mlp: PackedMlp, x: torch.Tensor, act_last: int, x2: Optional[torch.Tensor] = None, x2_row_div: int = 1,
                out: Optional[torch.Tensor] = None, ws: Optional[torch.Tensor] = None):
    """Run a packed tanh MLP on the rows of ``x`` (2-D, last dim contiguous) [+ label columns ``x2``]."""
    rows, k1 = x.shape
    k2 = 0 if x2 is None else x2.shape[1]
    if out is None:
        out = torch.empty((rows, mlp.out_dim), dtype=torch.float32, device=x.device)
    need = _lib.load().dvae_mlp_workspace_floats(mlp.ref, rows)          # a host-side size formula, no device work
    if ws is None or ws.numel() < need:
        ws = torch.empty(max(need, 1), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("dvae_mlp_fwd", mlp.ref, _p(x), x.stride(0), k1, _p(x2), 0 if x2 is None else x2.stride(0), k2,
                  max(1, x2_row_div), rows, act_last, _p(out), out.stride(0), _p(ws), _stream())
    return out


# --------------------------------------------------------------------------------------------- MCEM
@dataclasses.dataclass
class McemConfig:
    niter: int = 100
    keep_E: int = 10
    burn_E: int = 30
    keep_WF: int = 25
    burn_WF: int = 75
    var_rw: float = 0.01
    nmf_rank: int = 10
    eps: float = 1e-8
    n_chains: int = 1
    seed: int = 0
    sampler: str = "fp32"        # "fp32": CUDA-core exact mode; "tc": tcgen05 BF16 kernels; "auto": tc when supported
    fuse_wstat: bool = True      # tc only: per-frame statistics of the W update instead of a pass over Vs
    emit_vs: bool = True         # tc only: the sampler writes the kept samples' variances itself (BF16); False: FP32 decode
    w_partials: bool = True      # with emit_vs: reduce the W-update sums inside the statistics kernel (no A1 / A2 round trip)


class InjectedDraws:
    """Random numbers supplied by the caller (parity runs), in the batch layout.

    ``W0 [B][K][F]``, ``H0 [NT][K]`` and, per ``sample_posterior`` call c (0..niter-1 for the E-steps, niter for the
    final filter), ``eps[c]: [n_iter][NT*C][L]`` and ``u[c]: [n_iter][NT*C]``.
    """

    def __init__(self, W0, H0, eps, u):
        self.W0, self.H0, self.eps, self.u = W0, H0, eps, u

    def mh_draws(self, call, n_iter, chains, L):
        return self.eps[call], self.u[call]


class TorchCpuDraws(InjectedDraws):
    """Draws taken from torch's global CPU generator in the reference's consumption order, for ONE
    utterance: ``rand(F,K)``, ``rand(K,N)``, ``randn(N,L)`` x2 (the two encoder passes, unused), then per MH iteration
    ``randn(L,N)`` and ``rand(N)``.  With the same ``torch.manual_seed`` the reference on ``device='cpu'`` sees the
    very same numbers."""

    def __init__(self, F, K, N, L):
        W0 = torch.rand(F, K)
        H0 = torch.rand(K, N)
        torch.randn(N, L)
        torch.randn(N, L)
        super().__init__(W0.t().contiguous()[None], H0.t().contiguous(), None, None)

    def mh_draws(self, call, n_iter, chains, L):
        eps = torch.empty(n_iter, chains, L)
        u = torch.empty(n_iter, chains)
        for it in range(n_iter):
            eps[it] = torch.randn(L, chains).t()
            u[it] = torch.rand(chains)
        return eps, u


def tc_supported(w: "VaeWeights") -> bool:
    """True when the decoder fits the tcgen05 kernels: 513 bins, 1-2 hidden layers of 128 units, L in {16, 32}; any number of
    label inputs (more than three are folded into a per-frame bias, which needs a hidden layer to add it to)."""
    dims = w.dec.dims
    return dims[-1] == 513 and len(dims) in (3, 4) and all(d == 128 for d in dims[1:-1]) and w.z_dim in (16, 32)


class McemEngine:
    """EM driver for one ragged batch (replaces ``EM.run`` + ``MCEM_*`` of packages/models/mcem.py)."""

    def __init__(self, weights: VaeWeights, cfg: McemConfig, device):
        self.dev = _require_cuda(device)
        self.w = weights
        self.cfg = cfg
        if cfg.sampler not in ("fp32", "tc", "auto"):
            raise ValueError("sampler must be 'fp32', 'tc' or 'auto'")
        if cfg.sampler == "auto":            # tensor cores whenever the decoder has the shape the tcgen05 kernels serve
            cfg = dataclasses.replace(cfg, sampler="tc" if tc_supported(weights) else "fp32")
            self.cfg = cfg
        if not 1 <= cfg.nmf_rank <= _lib.MAX_K:
            raise ValueError("nmf_rank must be in 1..%d" % _lib.MAX_K)
        if weights.z_dim > _lib.MAX_L:
            raise ValueError("latent dimension above %d" % _lib.MAX_L)
        self.F = weights.x_dim
        self.ld = _ld_for(self.F)
        self._buf = {}
        self.mh_calls = 0
        self.kernel_launches = 0
        self.timing = False          # when True, CUDA events bracket every stage (read with stage_times_ms())
        self._events = []

    class _Stage:
        def __init__(self, eng, name):
            self.eng, self.name = eng, name

        def __enter__(self):
            if self.eng.timing:
                self.t0 = torch.cuda.Event(enable_timing=True)
                self.t1 = torch.cuda.Event(enable_timing=True)
                self.t0.record()

        def __exit__(self, *exc):
            if self.eng.timing:
                self.t1.record()
                self.eng._events.append((self.name, self.t0, self.t1))

    def stage(self, name):
        return McemEngine._Stage(self, name)

    def stage_times_ms(self, reset=True):
        """Sum of device time per stage (ms) and number of bracketed calls, from the recorded CUDA events."""
        torch.cuda.current_stream(self.dev).synchronize()
        out = {}
        for name, t0, t1 in self._events:
            ms, n = out.get(name, (0.0, 0))
            out[name] = (ms + t0.elapsed_time(t1), n + 1)
        if reset:
            self._events = []
        return out

    # -- buffers are cached by (name, shape) so repeated batches of the same geometry do not reallocate
    def _get(self, name, shape, dtype=torch.float32):
        key = (name, tuple(shape), dtype)
        t = self._buf.get(key)
        if t is None:
            for k in [k for k in self._buf if k[0] == name]:
                del self._buf[k]
            t = torch.empty(shape, dtype=dtype, device=self.dev)
            self._buf[key] = t
        return t

    @_on_device
    def init_parameters(self, X: torch.Tensor, P: torch.Tensor, batch: RaggedBatch, y: Optional[torch.Tensor] = None,
                        draws: Optional[InjectedDraws] = None):
        """``EM.init_parameters`` + ``MCEM_*.init_parameters`` (mcem.py:36-58, 195-205, 358-370) for the batch."""
        cfg, w = self.cfg, self.w
        if X.shape != (batch.NT, self.ld) or P.shape != (batch.NT, self.ld):
            raise ValueError("X / P must be [NT][ld=%d]" % self.ld)
        if (w.y_dim > 0) != (y is not None):
            raise ValueError("labels y are required exactly for the M2 variants")
        if y is not None and tuple(y.shape) != (batch.NT, w.y_dim):
            raise ValueError("y must be [NT][y_dim=%d]" % w.y_dim)
        self.batch, self.X, self.P, self.y = batch, X, P, (None if y is None else y.contiguous().float())
        # what the tensor-core kernels see of the labels: the labels themselves (<= 3) or a per-frame layer-1 bias
        self.tc_y, self.ybias, self.kscale = self.y, None, None
        if cfg.sampler == "tc":
            from . import tc
            self.kscale = tc.row_scale(self)
        if cfg.sampler == "tc" and w.tc_label_bias:
            self.tc_y = None
            self.ybias = mlp_forward(w.ybias_mlp, self.y, _lib.ACT_NONE, out=self._get("ybias", (batch.NT, 128)))
        B, NT, K, ld, C_ = batch.B, batch.NT, cfg.nmf_rank, self.ld, cfg.n_chains
        self.W = self._get("W", (B, K, ld))
        self.H = self._get("H", (NT, K))
        self.g = self._get("g", (NT,))
        self.Vb = self._get("Vb", (NT, ld))
        if draws is None:
            _lib.call("dvae_nmf_init", cfg.seed, _p(batch.utt_ids), _p(batch.fr_off), B, NT, self.F, K, ld, cfg.eps,
                      _p(self.W), _p(self.H), _p(self.g), _stream())
        else:
            self.W.zero_()
            self.W[:, :, :self.F] = torch.clamp_min(torch.as_tensor(draws.W0).to(self.dev, torch.float32), cfg.eps)
            self.H.copy_(torch.clamp_min(torch.as_tensor(draws.H0).to(self.dev, torch.float32), cfg.eps))
            self.g.fill_(1.0)
        _lib.call("dvae_nmf_vb", _p(self.W), _p(self.H), _p(batch.frame_utt), NT, self.F, K, ld, _p(self.Vb), _stream())
        # chain start = encoder posterior mean of the noisy power spectrogram (the clean-speech encoding of
        # mcem.py:201 is dead code whose only effect is consuming one randn draw)
        mu = mlp_forward(w.enc_mu, P[:, :self.F], _lib.ACT_NONE, x2=self.y if w.enc_takes_y else None)
        self.Z = self._get("Z", (NT * C_, w.z_dim))
        self.Z.copy_(mu.repeat_interleave(C_, dim=0) if C_ > 1 else mu)
        self.n_accept = self._get("n_accept", (NT * C_,), torch.int32)
        self.n_accept.zero_()
        self.mh_calls = 0
        self.mh_iter0 = 0
        self.vst_R = 0                    # > 0: the last E-step's variances live in the sampler's emission (VsT / vs_idx)
        self._Vs = None
        self._last_Zs = None
        self.cost = self._get("cost", (cfg.niter, B), torch.float64)
        self.kernel_launches = 0

    @property
    def Vs_flat(self):
        """Dense FP32 buffer for materialised speech variances: NT * R_E rows, and at least one frame's worth of Wiener
        samples.  Allocated on first use (the tensor-core path with one chain per frame never needs it)."""
        b, cfg = self.batch, self.cfg
        return self._get("Vs", (max(b.NT * cfg.n_chains * cfg.keep_E, cfg.n_chains * cfg.keep_WF), self.ld))

    @property
    def Vs(self):
        """Speech variances of the last E-step's kept samples, ``[NT][R][ld]`` FP32 (the reference's ``self.Vs`` (R,F,N) up
        to layout).  When they only exist in the sampler's BF16 emission they are unpacked on demand."""
        if self._Vs is None and self.vst_R:
            with torch.cuda.device(self.dev):
                from . import tc
                self._Vs = tc.vst_unpack(self, self.vst_R)
        elif self._Vs is None and getattr(self, "_last_Zs", None) is not None:
            Zs = self._last_Zs                     # the final filter never materialises its samples: decode them on demand
            out = torch.zeros((self.batch.NT, Zs.shape[1], self.ld), dtype=torch.float32, device=self.dev)
            self.decode_samples(Zs, 0, self.batch.NT, out)
            self._Vs = out
        return self._Vs

    # -- one sample_posterior call (mcem.py:207-277): returns kept samples [NT][C*keep][L]
    @_on_device
    def sample_posterior(self, keep: int, burn: int, draws: Optional[InjectedDraws] = None, a_trace=None, emit: bool = False):
        cfg, w, b = self.cfg, self.w, self.batch
        C_, L = cfg.n_chains, w.z_dim
        chains = b.NT * C_
        Zs = self._get("Zs%d" % keep, (b.NT, C_ * keep, L))
        rng = _lib.DvaeRng()
        rng.seed = cfg.seed & 0xFFFFFFFFFFFFFFFF
        rng.iter0 = self.mh_iter0
        eps = u = None
        if draws is not None:
            eps, u = draws.mh_draws(self.mh_calls, keep + burn, chains, L)
            eps = torch.as_tensor(eps).to(self.dev, torch.float32).contiguous()
            u = torch.as_tensor(u).to(self.dev, torch.float32).contiguous()
            if tuple(eps.shape) != (keep + burn, chains, L) or tuple(u.shape) != (keep + burn, chains):
                raise ValueError("injected draws for call %d have the wrong shape" % self.mh_calls)
            rng.eps, rng.u = eps.data_ptr(), u.data_ptr()
        with self.stage("mh"):
            if cfg.sampler == "tc":
                from . import tc
                tc.mh_chain_tc(self, Zs, keep, burn, rng, a_trace, emit)
            else:
                need = _lib.load().dvae_mh_workspace_floats(w.dec.ref, chains, self.F)
                ws = self._get("mh_ws", (max(int(need), 1),))
                _lib.call("dvae_mh_chain_f32", w.dec.ref, _p(self.P), _p(self.Vb), _p(self.g), _p(self.y), w.y_dim,
                          _p(b.frame_gid), _p(b.frame_idx), _p(self.Z), _p(Zs), b.NT, self.F, self.ld, L, C_, burn, keep,
                          float(cfg.var_rw), C.byref(rng), _p(self.n_accept), _p(a_trace), _p(ws), _stream())
                self.kernel_launches += (keep + burn + 1) * (len(w.dec.dims) - 1 + 1)
        self.mh_rows = getattr(self, "mh_rows", 0) + chains * (keep + burn + 1)
        self.mh_calls += 1
        self.mh_iter0 += keep + burn
        return Zs

    @_on_device
    def decode_samples(self, Zs: torch.Tensor, n0: int, n1: int, Vs: torch.Tensor):
        """``compute_Vs`` (mcem.py:280-290) for frames [n0, n1): Vs[(n-n0)][r][ld] = decoder([Zs[n][r]; y[n]])."""
        R, L = Zs.shape[1], Zs.shape[2]
        rows = (n1 - n0) * R
        x = Zs[n0:n1].reshape(rows, L)
        x2 = None if self.y is None else self.y[n0:n1]
        out = Vs.view(-1, self.ld)[:rows]
        with self.stage("decode"):
            if self.cfg.sampler == "tc":
                from . import tc
                tc.decode_tc(self, x, x2, R, out, None if self.ybias is None else self.ybias[n0:n1])
            else:
                ws = self._get("mlp_ws", (max(int(_lib.load().dvae_mlp_workspace_floats(self.w.dec.ref, rows)), 1),))
                mlp_forward(self.w.dec, x, _lib.ACT_EXP, x2=x2, x2_row_div=R, out=out, ws=ws)
                self.kernel_launches += len(self.w.dec.dims) - 1

    @_on_device
    def e_step(self, draws=None):
        """E-step (mcem.py:292-308): sample, keep the last state, decode the kept samples.  On the tensor-core path (one chain
        per frame, or a power-of-two number of them) the sampler emits the variances itself (BF16) and only the statistics
        of the W update are computed here; otherwise the kept samples are decoded into FP32 ``Vs``."""
        cfg = self.cfg
        self.wstat, self.wpart, self._Vs, self.vst_R = None, None, None, 0
        if cfg.sampler == "tc":
            from . import tc
            if tc.vst_supported(self, cfg.keep_E):
                self.sample_posterior(cfg.keep_E, cfg.burn_E, draws, emit=True)
                self.vst_R, self.R = cfg.keep_E, cfg.n_chains * cfg.keep_E      # kept per chain, samples per frame
                with self.stage("decode"):
                    if cfg.w_partials:
                        self.wpart = tc.vst_w_partials(self, self.vst_R)
                    else:
                        self.wstat = tc.vst_frame_stats(self, self.vst_R)
                return
        Zs = self.sample_posterior(cfg.keep_E, cfg.burn_E, draws)
        self.R = Zs.shape[1]
        self._Vs = self.Vs_flat[: self.batch.NT * self.R].view(self.batch.NT, self.R, self.ld)
        if cfg.sampler == "tc" and self.R % 10 == 0 and cfg.nmf_rank <= 10 and cfg.fuse_wstat:
            from . import tc
            with self.stage("decode"):
                self.wstat = tc.decode_stats_tc(self, Zs, self._Vs)
        else:
            self.decode_samples(Zs, 0, self.batch.NT, self._Vs)

    @_on_device
    def m_step(self, it: int):
        b, cfg = self.batch, self.cfg
        need = _lib.load().dvae_nmf_workspace_floats(b.B, cfg.nmf_rank, self.ld, b.max_frames)
        ws = self._get("nmf_ws", (int(need),))
        st = self._buf.get("tc_status")
        with self.stage("mstep"):
            if self.vst_R:
                from . import tc
                w = self.w
                _lib.call("dvae_nmf_mstep_vst", w.dec_tc.ref, _p(tc.decoder_image(w)), w.z_dim, w.tc_y_dim, _p(self.P), _p(self.VsT),
                          _p(self.vs_idx), self.vst_R, _p(self.W), _p(self.H), _p(self.g), _p(self.Vb),
                          C.c_void_p(self.cost[it].data_ptr()), _p(b.fr_off), b.B, b.NT, cfg.nmf_rank, self.ld, b.max_frames,
                          _p(ws), _p(self.wstat), _p(self.wpart),
                          _p(b.segments(n_chains=cfg.n_chains)[2]) if self.wpart is not None else None, cfg.n_chains, _p(st), _stream())
            else:
                _lib.call("dvae_nmf_mstep", _p(self.P), _p(self._Vs), self.R, _p(self.W), _p(self.H), _p(self.g), _p(self.Vb),
                          C.c_void_p(self.cost[it].data_ptr()), _p(b.fr_off), _p(b.frame_utt), b.B, b.NT, self.F, cfg.nmf_rank,
                          self.ld, b.max_frames, _p(ws), _p(getattr(self, "wstat", None)), _p(st), _stream())
        self.kernel_launches += 4

    @_on_device
    def wiener(self, draws=None):
        """``compute_WF(sample=True)`` + the mask application of ``EM.run`` (mcem.py:310-329, 176-177)."""
        cfg, b = self.cfg, self.batch
        Zs = self.sample_posterior(cfg.keep_WF, cfg.burn_WF, draws)
        self._Vs, self.vst_R, self._last_Zs = None, 0, Zs
        R = Zs.shape[1]
        WFs = self._get("WFs", (b.NT, self.ld))
        WFn = self._get("WFn", (b.NT, self.ld))
        chunk = next((c for c in (30, 25, 10) if R % c == 0), 0)
        if cfg.sampler == "tc" and chunk:
            # fused: the filter samples are decoded chunk by chunk straight into A1 = sum_r 1 / Vx; the masks follow from
            # mean_r g Vs / Vx = 1 - Vb A1 / R, so none of the R x F x N variances is written to memory
            from . import tc
            for r0 in range(0, R, chunk):
                A1 = tc.decode_a1_tc(self, Zs, r0, chunk)
                _lib.call("dvae_wiener_from_a1", _p(A1), _p(self.Vb), chunk, b.NT, self.F, self.ld, _p(WFs), _p(WFn),
                          1 if r0 == 0 else 0, _stream())
                self.kernel_launches += 1
        else:
            # decode in frame chunks so the Vs buffer of the E-step (NT * R_E rows) is reused
            flat = self.Vs_flat
            step = max(1, min(b.NT, flat.shape[0] // R))
            for n0 in range(0, b.NT, step):
                n1 = min(b.NT, n0 + step)
                self.decode_samples(Zs, n0, n1, flat)
                _lib.call("dvae_wiener_accum", _p(flat), R, _p(self.Vb[n0:n1]), _p(self.g[n0:n1]), n1 - n0, self.F, self.ld,
                          _p(WFs[n0:n1]), _p(WFn[n0:n1]), 1, _stream())
                self.kernel_launches += 1
        self.S_hat = self._get("S_hat", (b.NT, self.ld), torch.complex64)
        self.N_hat = self._get("N_hat", (b.NT, self.ld), torch.complex64)
        _lib.call("dvae_wiener_apply", _p(self.X), _p(WFs), _p(WFn), R, b.NT, self.F, self.ld, _p(self.S_hat),
                  _p(self.N_hat), _stream())
        self.kernel_launches += 1
        self.WFs, self.WFn, self.R_wf = WFs, WFn, R

    @_on_device
    def run(self, draws: Optional[InjectedDraws] = None):
        """``EM.run`` (mcem.py:156-179): niter x (E-step, M-step, cost), then the Wiener estimates.

        Returns the device cost matrix ``[niter][B]`` (float64); nothing synchronises with the host.
        """
        for it in range(self.cfg.niter):
            self.e_step(draws)
            self.m_step(it)
        self.wiener(draws)
        return self.cost


# --------------------------------------------------------------------------------------------- end-to-end helper
class Enhancer:
    """``process_sublist`` replacement (scripts/evaluate_ntcd_M1.py:190-214): whole batches, host buffers in and out."""

    def __init__(self, sd, variant: str, cfg: McemConfig, device=0, fs=FS, n_fft=N_FFT, hop=HOP):
        self.dev = _require_cuda(device)
        self.weights = sd if isinstance(sd, VaeWeights) else VaeWeights(sd, variant, self.dev)
        self.cfg = cfg
        self.engine = McemEngine(self.weights, cfg, self.dev)
        self.fs, self.n_fft, self.hop = fs, n_fft, hop
        self._pinned = {}
        self._out_free = []              # pinned result buffers not leased to any caller (see _pin_out)
        self._h2d_done = None            # event after the last H2D copy out of the pinned staging buffers
        self._pool = None
        self._stage_threads = max(1, min(8, (os.cpu_count() or 1) // 2))

    def _pin(self, name, n, dtype):
        t = self._pinned.get(name)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty(n, dtype=dtype).pin_memory()
            self._pinned[name] = t
        return t[:n]

    def _pin_out(self, n):
        """A pinned float32 result buffer leased to the caller: ``(tensor, ndarray)``.

        Results are returned as views of pinned memory (no extra host copy).  Ownership is explicit: every lease gets a
        FRESH ndarray object over the pinned storage, the result views handed out keep that object alive through their
        ``.base``, and the storage goes back to the free list only when the object is collected (``weakref.finalize``),
        i.e. when the caller has dropped every array of that call.  Results never change under the caller, and no
        reference counts are inspected.
        """
        import weakref
        k = next((i for i, t in enumerate(self._out_free) if t.numel() >= n), None)
        t = self._out_free.pop(k) if k is not None else torch.empty(n, dtype=torch.float32).pin_memory()
        a = t.numpy()                                              # new ndarray object per lease
        free = self._out_free

        def give_back(tt=t):
            if len(free) < 4:                                      # keep a few buffers for reuse, let the rest go
                free.append(tt)

        weakref.finalize(a, give_back)
        return t, a

    def _stage(self, host_np, x_list, off, lens):
        """Copy the utterances into the pinned staging buffer (a few threads: numpy releases the GIL while copying)."""
        B = len(x_list)

        def work(lo, hi):
            for u in range(lo, hi):
                host_np[off[u]:off[u] + lens[u]] = x_list[u]

        nthr = min(self._stage_threads, max(1, B // 16))
        if nthr <= 1:
            work(0, B)
            return
        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(self._stage_threads)
        step = (B + nthr - 1) // nthr
        futs = [self._pool.submit(work, lo, min(lo + step, B)) for lo in range(0, B, step)]
        for f in futs:
            f.result()

    @_on_device
    def run_device(self, x_dev, x_off, x_len, batch, y, total, max_len, draws=None):
        """The whole path on device-resident inputs: STFT -> MCEM -> Wiener -> ISTFT.  Nothing touches the host.

        ``x_dev``: concatenated float32 signals; ``x_off`` / ``x_len``: per-utterance offsets / lengths (device);
        returns ``(s_hat, n_hat)`` in the layout of ``x_dev`` and the cost matrix ``[niter][B]``.
        """
        eng = self.engine
        X, P = stft_batch(x_dev, x_off, x_len, batch, self.n_fft, self.hop)
        eng.init_parameters(X, P, batch, y, draws)
        cost = eng.run(draws)
        s_dev = istft_batch(eng.S_hat, batch, x_off, x_len, total, max_len, self.n_fft, self.hop,
                            out=eng._get("s_out", (total,)))
        n_dev = istft_batch(eng.N_hat, batch, x_off, x_len, total, max_len, self.n_fft, self.hop,
                            out=eng._get("n_out", (total,)))
        eng.kernel_launches += 4          # stft, power-free init kernels are counted there; 2 istft + stft + vb
        return s_dev, n_dev, cost

    @_on_device
    def enhance(self, x_list, y_list=None, utt_ids=None, max_frames_list=None, draws=None, return_device=False, s_list=None):
        """Enhance a list of 1-D float32 host signals.  Returns ``(s_hat_list, n_hat_list, cost [B][niter])``.

        ``y_list``: per-utterance label arrays ``(y_dim, N_u)`` as the reference passes them (M2 variants).
        ``max_frames_list``: optional per-utterance frame caps (the reference truncates to the video length).
        ``s_list``: clean signals; for a model with one label input and no ``y_list`` the labels are the time-domain VAD
        of the clean speech (``clean_speech_VAD``, scripts/evaluate_ntcd_M2.py), computed on the device.
        ``utt_ids``: global utterance ids = Philox counter word 0; callers that enhance several batches must pass distinct
        ids per utterance (``batch_io.process_sublist`` does), or every batch replays the same random draws.
        ``return_device=True`` returns device tensors that alias the engine's output buffers: they are valid until the
        next call on this Enhancer (clone them to keep them).
        """
        eng, dev = self.engine, self.dev
        B = len(x_list)
        lens = np.array([len(x) for x in x_list], np.int32)
        if np.any(lens < 1):
            raise ValueError("empty utterance")
        nfr = np.array([num_frames(int(t), self.n_fft, self.hop, self.fs) for t in lens], np.int64)
        if np.any(nfr < 1):
            raise ValueError("utterance shorter than one STFT frame")
        if max_frames_list is not None:
            nfr = np.minimum(nfr, np.asarray(max_frames_list, np.int64))
        off = np.zeros(B + 1, np.int64)
        np.cumsum((lens + 1) // 2 * 2, out=off[1:])            # even offsets
        total = int(off[-1])
        if self._h2d_done is not None:          # the previous call's H2D copies must have left the staging buffers
            self._h2d_done.synchronize()
        host = self._pin("x", total, torch.float32)
        self._stage(host.numpy(), x_list, off, lens)
        x_dev = host.to(dev, non_blocking=True)
        x_off = torch.from_numpy(off[:-1].copy()).to(dev)
        x_len = torch.from_numpy(lens).to(dev)
        batch = RaggedBatch(nfr, dev, utt_ids)
        y = None
        if self.weights.y_dim and y_list is None and s_list is not None:
            if self.weights.y_dim != 1 or len(s_list) != B or any(len(a) != t for a, t in zip(s_list, lens)):
                raise ValueError("device VAD labels need one clean signal per utterance and a model with one label input")
            from .packages.processing.target import vad_batch
            hs_in = self._pin("s_in", total, torch.float32)
            self._stage(hs_in.numpy(), s_list, off, lens)
            y = vad_batch(hs_in.to(dev, non_blocking=True), x_off, x_len, batch, self.n_fft, self.hop).view(batch.NT, 1)
        elif self.weights.y_dim:
            if y_list is None:
                raise ValueError("this model needs labels y")
            yc = np.concatenate([np.asarray(yy, np.float32)[:, :n].T for yy, n in zip(y_list, nfr)], axis=0)
            if yc.shape != (batch.NT, self.weights.y_dim):
                raise ValueError("labels do not cover the frames")
            y = torch.from_numpy(np.ascontiguousarray(yc)).to(dev)
        self._h2d_done = torch.cuda.Event()
        self._h2d_done.record()
        s_dev, n_dev, cost = self.run_device(x_dev, x_off, x_len, batch, y, total, int(lens.max()), draws)
        self.h2d_bytes = total * 4 + (0 if y is None else y.numel() * 4)
        self.d2h_bytes = 2 * total * 4 + cost.numel() * 8
        from . import tc
        if return_device:
            tc.check_status(eng)                               # synchronises
            return s_dev, n_dev, cost
        hs, s_np = self._pin_out(total)
        hn, n_np = self._pin_out(total)
        hs[:total].copy_(s_dev, non_blocking=True)
        hn[:total].copy_(n_dev, non_blocking=True)
        cost_h = cost.t().contiguous().cpu()                   # synchronises the stream
        torch.cuda.current_stream().synchronize()
        tc.check_status(eng)
        s_list = [s_np[off[u]:off[u] + lens[u]] for u in range(B)]      # views of pinned memory, see _pin_out
        n_list = [n_np[off[u]:off[u] + lens[u]] for u in range(B)]
        return s_list, n_list, cost_h.numpy()
