#!/usr/bin/env python
"""Phase timing of the tc2 sampler: clock64() stamps of CTA 0, iteration 4 of its first tile (full-size batch)."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from dvae_b200 import _lib, synth                                                     # noqa: E402
from dvae_b200.engine import McemConfig, McemEngine, RaggedBatch, VaeWeights, _p       # noqa: E402

DEV = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = 185
rng = np.random.default_rng(0)
P = torch.tensor(rng.gamma(1.0, 0.05, size=(B * N, 520)).astype(np.float32)).to(DEV)
X = torch.zeros((B * N, 520), dtype=torch.complex64, device=DEV)
sd = synth.xavier_state_dict("M1", 513, 16, [128, 128], 0, seed=9, out_bias=float(np.log(0.05)))
w = VaeWeights(sd, "M1", DEV)
eng = McemEngine(w, McemConfig(niter=1, keep_E=30, burn_E=30, sampler="tc"), DEV)
eng.init_parameters(X, P, RaggedBatch([N] * B, DEV))
buf = torch.zeros(64, dtype=torch.int64, device=DEV)
buf[56] = 2 ** 62
_lib.call("dvae_debug_set_clock_buffer", _p(buf)); _lib.call("dvae_debug_set_clock_buffer3", _p(buf)); _lib.call("dvae_debug_set_clock_buffer4", _p(buf))
eng.timing = True
for _ in range(2):
    buf[56] = 2 ** 62; buf[57] = 0; buf[58] = 0
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); eng.sample_posterior(30, 30); t1.record(); torch.cuda.synchronize()
print("stage events:", eng.stage_times_ms())
print("sample_posterior(60 iters, %d chains): %.2f ms -> %.0f clk per tile-eval at 1.965 GHz" %
      (B * N, t0.elapsed_time(t1), t0.elapsed_time(t1) * 1e-3 * 1.965e9 / 61 / max(1, (B * N / 128) / 148)))
c = buf.cpu().numpy()
print("per-CTA tile-loop clocks (v2 sampler): min %d max %d mean %d" % (c[56], c[57], c[58] / 148))
_lib.call("dvae_debug_set_clock_buffer", None); _lib.call("dvae_debug_set_clock_buffer3", None); _lib.call("dvae_debug_set_clock_buffer4", None)
names = {0: "iter start", 1: "after S1 (A1 written)", 26: "proposal written (thread 0)", 27: "bias preload complete", 24: "layer-1 accumulator ready", 25: "layer-2 accumulator ready", 20: "warp0 done hidden-1", 2: "after S2", 21: "warp0 done hidden-2", 3: "after S3",
         15: "chunk0 ready", 4: "warp0 done chunk0", 16: "chunk1 ready", 5: "warp0 done chunk1", 17: "chunk2 ready",
         6: "warp0 done chunk2 (v3)", 18: "chunk3 ready (v3)", 7: "warp0 done chunk3 (v3)", 19: "chunk4 ready (v3)",
         8: "warp0 done last chunk", 9: "after S4", 23: "accept done (end of iteration)"}
import os
if os.environ.get("DVAE_TC_SAMPLER") == "v4":       # phase 9 of CTA 0: front thread 0 and back warp 4 lane 0
    names = {0: "front: phase start", 1: "front: finalised, operand written", 2: "front: h1 written", 3: "front: h2 stored to TMEM",
             10: "back: chunk0 ready", 20: "back: chunk0 done", 11: "back: chunk1 ready", 21: "back: chunk1 done",
             12: "back: chunk2 ready", 22: "back: chunk2 done", 13: "back: chunk3 ready", 23: "back: chunk3 done"}
base = c[0]
if os.environ.get("DVAE_TC_SAMPLER") == "v4":
    print("back done, phases 60..69:", [int(c[30 + i] - c[30]) for i in range(10)])
    print("front h2 stored, phases 60..69:", [int(c[14 + i] - c[30]) for i in range(10)])
    for b0 in (40, 50):
        print("CTA %s rounds (clk from round 0 start):" % ("0" if b0 == 40 else "100"), [int(c[b0 + i] - c[b0]) for i in range(6) if c[b0 + i]])
for k, v in sorted(((k, c[k] - base) for k in names if c[k]), key=lambda kv: kv[1]):
    print("%8d  %s" % (v, names[k]))
