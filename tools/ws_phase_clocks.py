#!/usr/bin/env python
"""Phase timing of the fused decode + W-statistics kernel: clock64() stamps of CTA 0, third tile of its first item."""
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
eng = McemEngine(w, McemConfig(niter=1, keep_E=30, burn_E=2, sampler="tc"), DEV)
eng.init_parameters(X, P, RaggedBatch([N] * B, DEV))
buf = torch.zeros(64, dtype=torch.int64, device=DEV)
_lib.call("dvae_debug_set_clock_buffer_ws", _p(buf)); _lib.call("dvae_debug_set_clock_buffer_ds", _p(buf))
eng.timing = True
for _ in range(2):
    eng.e_step()
    eng.m_step(0)
torch.cuda.synchronize()
print("stage events (2 EM iterations):", eng.stage_times_ms())
c = buf.cpu().numpy()
_lib.call("dvae_debug_set_clock_buffer_ws", None); _lib.call("dvae_debug_set_clock_buffer_ds", None)
names = {0: "tile start", 1: "after S1 (A1, g, H staged)", 2: "after S2 (h1)", 3: "after S3 (h2)", 4: "tile done",
         10: "group0 chunk0 ready", 20: "group0 chunk0 done", 11: "group1 chunk1 ready", 21: "group1 chunk1 done",
         12: "group0 chunk2 ready", 22: "group0 chunk2 done", 13: "group1 chunk3 ready", 23: "group1 chunk3 done",
         14: "group0 chunk4 ready", 24: "group0 chunk4 done"}
import os
if os.environ.get("DVAE_TC_DECODE", "v3") != "v2":
    names = {0: "front: tile start", 1: "front: z loaded, A free", 2: "front: A1 written", 3: "front: h1 done", 4: "front: h2 done"}
    for j in range(5):
        names[10 + j] = "back w4: chunk %d ready" % j
        names[20 + j] = "back w4: chunk %d done" % j
        names[30 + j] = "back w19: chunk %d ready" % j
        names[40 + j] = "back w19: chunk %d done" % j
        names[50 + j] = "issuer: chunk %d+2 issued" % j
base = c[0]
for k, v in sorted(((k, c[k] - base) for k in names if c[k]), key=lambda kv: kv[1]):
    print("%8d  %s" % (v, names[k]))
