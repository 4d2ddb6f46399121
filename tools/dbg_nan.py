#!/usr/bin/env python
"""Status probe: the bench workload (M1, 512 utterances, 100 EM iterations) stage by stage, reporting the first place where the
device status word (DVAE_STATUS_*) is set or a tensor stops being finite; then six full runs.  python tools/dbg_nan.py [B]"""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import bench
from dvae_b200 import synth, tc
from dvae_b200.engine import Enhancer, McemConfig, RaggedBatch, stft_batch
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T = int(3.0 * synth.FS)
xs, ss = bench.synth_set(1000, B)
sd = bench.model_weights("M1", bench.reference_power())
cfg = McemConfig(var_rw=0.01, nmf_rank=10, eps=1e-8, seed=2024, sampler="tc", **bench.schedule("M1", 100))
enh = Enhancer(sd, "M1", cfg, device=0); eng = enh.engine
batch = RaggedBatch([synth.num_frames(T)] * B, dev, list(range(B)))
x_dev = torch.from_numpy(np.stack(xs)).to(dev).reshape(-1)
x_off = (torch.arange(B, dtype=torch.int64) * T).to(dev); x_len = torch.full((B,), T, dtype=torch.int32, device=dev)
X, P = stft_batch(x_dev, x_off, x_len, batch)
print("stft finite", bool(torch.isfinite(torch.view_as_real(X)[:, :513]).all()), bool(torch.isfinite(P[:, :513]).all()), "P min/max", P[:, :513].min().item(), P[:, :513].max().item())
eng.init_parameters(X, P, batch, None, None)
st = tc._status(eng)
for it in range(100):
    eng.e_step()
    s1 = int(st.item())
    eng.m_step(it)
    s2 = int(st.item())
    if s1 or s2:
        print("iteration", it, "status after e_step", s1, "after m_step", s2)
        for name in ("W", "H", "g", "Vb", "Z"):
            t = getattr(eng, name)
            print("   ", name, "finite", bool(torch.isfinite(t).all()), "min", t.min().item(), "max", t.max().item())
        c = eng.cost[it]
        bad = (~torch.isfinite(c)).nonzero().flatten()
        print("    cost non-finite utterances", bad[:10].tolist(), "n", bad.numel())
        break
else:
    print("100 iterations clean; cost mean", eng.cost[99].mean().item())
print("--- full runs")
for rep in range(6):
    X, P = stft_batch(x_dev, x_off, x_len, batch)
    ok_stft = bool(torch.isfinite(P[:, :513]).all())
    eng.init_parameters(X, P, batch, None, None)
    s0 = int(st.item())
    for it in range(cfg.niter):
        eng.e_step(); eng.m_step(it)
    s1 = int(st.item())
    eng.wiener()
    s2 = int(st.item())
    print("rep", rep, "stft finite", ok_stft, "status after init", s0, "after EM", s1, "after wiener", s2,
          "S_hat finite", bool(torch.isfinite(torch.view_as_real(eng.S_hat)[:, :513]).all()), flush=True)
    if s0 or s1 or s2:
        c = eng.cost
        bad = (~torch.isfinite(c)).nonzero()
        print("   non-finite cost entries", bad[:5].tolist(), "n", bad.shape[0])
        print("   Z finite", bool(torch.isfinite(eng.Z).all()), "g finite", bool(torch.isfinite(eng.g).all()), "Vb min", eng.Vb[:, :513].min().item())
        st.zero_()
