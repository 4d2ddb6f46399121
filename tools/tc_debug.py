#!/usr/bin/env python
"""Diagnostics of the tcgen05 decoder kernels on a GPU box: crafted weights isolate layer 1 / 2 / 3, then random
weights against the FP32 kernels, then the sampler's log-acceptance values.  Prints error maps per column / row block."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from dvae_b200 import _lib, synth, tc                                    # noqa: E402
from dvae_b200.engine import (InjectedDraws, McemConfig, McemEngine, RaggedBatch, VaeWeights, _p, _stream,  # noqa: E402
                              mlp_forward)

DEV = torch.device("cuda:0")
torch.manual_seed(0)


def block_report(name, got, ref, rel=True):
    err = (got - ref).abs()
    if rel:
        err = err / ref.abs().clamp_min(1e-30)
    print("%s: max %.3e mean %.3e" % (name, err.max().item(), err.mean().item()))
    cols = [(0, 64), (64, 128), (128, 256), (256, 384), (384, 512), (512, 513)]
    print("   by columns:", " ".join("[%d:%d] %.2e" % (a, b, err[:, a:b].max().item()) for a, b in cols if a < got.shape[1]))
    rows = [(0, 8), (8, 32), (32, 64), (64, 96), (96, 128), (128, 256)]
    print("   by rows   :", " ".join("[%d:%d] %.2e" % (a, b, err[a:b].max().item()) for a, b in rows if a < got.shape[0]))
    return err.max().item()


def status(st):
    torch.cuda.synchronize()
    print("   status word:", int(st.item()))


def decode(w, x, y=None, div=1):
    img = tc.decoder_image(w)
    out = torch.zeros((x.shape[0], 520), device=DEV)
    st = torch.zeros(1, dtype=torch.int32, device=DEV)
    _lib.call("dvae_decode_tc", w.dec.ref, _p(img), _p(x), x.shape[0], w.z_dim, _p(y), w.y_dim, None, div, _p(out), 520, _p(st), _stream())
    status(st)
    return out


def crafted(n_hidden):
    L = 16
    sd = synth.xavier_state_dict("M1", 513, L, [128] * n_hidden, 0, seed=1)
    rng = np.random.default_rng(0)
    sd["decoder.hidden.0.weight"] = (rng.standard_normal((128, L)) * 0.3).astype(np.float32)
    sd["decoder.hidden.0.bias"] = (rng.standard_normal(128) * 0.1).astype(np.float32)
    if n_hidden == 2:
        sd["decoder.hidden.1.weight"] = np.eye(128, dtype=np.float32)
        sd["decoder.hidden.1.bias"] = np.zeros(128, np.float32)
    w3 = np.zeros((513, 128), np.float32)
    w3[:128] = np.eye(128)
    w3[128:256] = 0.5 * np.eye(128)
    w3[512, 5] = 1.0
    sd["decoder.reconstruction.weight"] = w3
    sd["decoder.reconstruction.bias"] = np.zeros(513, np.float32)
    w = VaeWeights(sd, "M1", DEV)
    x = torch.randn((300, L), device=DEV)
    got = decode(w, x)[:, :513]
    h1 = torch.tanh(x @ torch.tensor(sd["decoder.hidden.0.weight"]).to(DEV).t() + torch.tensor(sd["decoder.hidden.0.bias"]).to(DEV))
    h = torch.tanh(h1) if n_hidden == 2 else h1
    ref = torch.zeros((300, 513), device=DEV)
    ref[:, :128] = h
    ref[:, 128:256] = 0.5 * h
    ref[:, 512] = h[:, 5]
    print("crafted weights, %d hidden layer(s): log(Vs) should equal the hidden activations" % n_hidden)
    block_report("  log Vs", torch.log(got), ref, rel=False)


def random_weights(variant, L, h):
    y_dim = 0 if variant == "M1" else 1
    sd = synth.xavier_state_dict(variant, 513, L, h, y_dim, seed=7, out_bias=-3.0)
    w = VaeWeights(sd, variant, DEV)
    rows, div = 1000, 10
    x = torch.randn((rows, L), device=DEV)
    y = (torch.rand((rows // div, 1), device=DEV) > 0.5).float() if y_dim else None
    ref = mlp_forward(w.dec, x, _lib.ACT_EXP, x2=y, x2_row_div=div)
    got = decode(w, x, y, div)[:, :513]
    print("random weights %s L=%d h=%s: tensor-core decode vs FP32 kernel (relative)" % (variant, L, h))
    return block_report("  Vs", got, ref)


def sampler(variant):
    y_dim = 0 if variant == "M1" else 1
    F, N, L = 513, 300, 16
    rng = np.random.default_rng(3)
    P = torch.tensor(rng.gamma(1.0, 0.05, size=(N, 520)).astype(np.float32)).to(DEV)
    sd = synth.xavier_state_dict(variant, F, L, [128, 128], y_dim, seed=9, out_bias=float(np.log(0.05)))
    w = VaeWeights(sd, variant, DEV)
    res = {}
    keep, burn = 3, 5
    eps = torch.randn((keep + burn, N, L))
    u = torch.rand((keep + burn, N))
    for sampler_name in ("fp32", "tc"):
        cfg = McemConfig(niter=1, keep_E=keep, burn_E=burn, keep_WF=2, burn_WF=2, sampler=sampler_name, seed=1)
        eng = McemEngine(w, cfg, DEV)
        batch = RaggedBatch([N], DEV)
        X = torch.zeros((N, 520), dtype=torch.complex64, device=DEV)
        y = (torch.rand((N, 1), device=DEV, generator=torch.Generator(device=DEV).manual_seed(1)) > 0.5).float() if y_dim else None
        W0 = torch.rand((1, 10, F), generator=torch.Generator().manual_seed(2))
        H0 = torch.rand((N, 10), generator=torch.Generator().manual_seed(3)) * 0.01
        draws = InjectedDraws(W0, H0, [eps], [u])
        eng.init_parameters(X, P, batch, y, draws)
        a = torch.zeros((keep + burn, N), device=DEV)
        Zs = eng.sample_posterior(keep, burn, draws, a_trace=a)
        torch.cuda.synchronize()
        if sampler_name == "tc":
            print("   status word:", int(eng._buf["tc_status"].item()))
        res[sampler_name] = (a.cpu(), Zs.cpu().clone(), eng.n_accept.cpu().clone())
    a0, a1 = res["fp32"][0], res["tc"][0]
    print("sampler %s: log-acceptance of iteration 0 (same state, same proposal)" % variant)
    d = (a0[0] - a1[0]).abs()
    print("   |a_fp32 - a_tc|: max %.3e mean %.3e ; |a_fp32| mean %.3e ; corr %.5f" %
          (d.max().item(), d.mean().item(), a0[0].abs().mean().item(), torch.corrcoef(torch.stack([a0[0], a1[0]]))[0, 1].item()))
    print("   accepted proposals: fp32 %d, tc %d of %d" % (res["fp32"][2].sum().item(), res["tc"][2].sum().item(), N * (keep + burn)))
    same = (res["fp32"][1] - res["tc"][1]).abs().amax(dim=(1, 2)) < 1e-6
    print("   chains with identical kept samples: %d of %d" % (same.sum().item(), N))


if __name__ == "__main__":
    which = sys.argv[1:] or ["crafted1", "crafted2", "random", "sampler"]
    if "crafted1" in which:
        crafted(1)
    if "crafted2" in which:
        crafted(2)
    if "random" in which:
        random_weights("M1", 16, [128, 128])
        random_weights("M2", 16, [128, 128])
        random_weights("M1", 32, [128])
    if "sampler" in which:
        sampler("M1")
        sampler("M2")
