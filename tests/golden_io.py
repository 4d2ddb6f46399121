"""Load the reference-generated fixtures of tests/golden (written by oracle/make_golden.py)."""
import glob
import os

import numpy as np

from dvae_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "mcem_*.npz")))


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, "mcem_%s.npz" % name))
        self.name = name
        self.X, self.S = z["X"], z["S"]
        self.y = z["y"] if z["y"].shape[0] else None
        self.variant = str(z["meta_variant"])
        self.n_fft, self.K, self.L = int(z["meta_n_fft"]), int(z["meta_K"]), int(z["meta_L"])
        self.h = [int(v) for v in z["meta_h"]]
        self.y_dim, self.niter = int(z["meta_y_dim"]), int(z["meta_niter"])
        self.sched = tuple(int(v) for v in z["meta_sched"])
        self.eps, self.var_RW = float(z["meta_eps"]), float(z["meta_var_RW"])
        self.F = self.n_fft // 2 + 1
        wv = "M2v3" if self.variant == "M2v2" else self.variant
        self.sd = synth.xavier_state_dict(wv, self.F, self.L, self.h, self.y_dim, seed=int(z["meta_weight_seed"]),
                                          out_bias=float(z["meta_out_bias"]))
        self.ref = {k[4:]: z[k] for k in z.files if k.startswith("ref_")}
        kinds, sizes, shapes, flat = z["draw_kinds"], z["draw_sizes"], z["draw_shapes"], z["draw_flat"]
        self.draws = []
        pos = 0
        for k, n, sh in zip(kinds, sizes, shapes):
            shape = tuple(int(v) for v in sh if v > 0)
            self.draws.append(("rand" if k == 0 else "randn", flat[pos:pos + n].reshape(shape)))
            pos += int(n)
