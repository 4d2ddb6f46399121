"""The IBM-conditioned M2 model (``y_dim = 513``, scripts/evaluate_ntcd_M2.py:66-73): 513 label inputs per frame on the
encoder and the decoder.  FP32 path: the labels are extra GEMM columns.  Tensor-core path: they are folded into a per-frame
layer-1 bias (include/dvae_b200.h, "Label inputs").  Both against the oracle on the same draws."""
import numpy as np
import pytest
import torch

from dvae_b200 import _lib, synth, tc
from dvae_b200.engine import McemConfig, McemEngine, RaggedBatch, VaeWeights, mlp_forward
from oracle import mcem_port, stft_np
from tests.gpu_util import DEV

pytestmark = pytest.mark.gpu
KW = dict(fs=16000, wlen_sec=64e-3, win="hann", hop_percent=0.25, center=False, pad_at_end=True)
IKW = dict(fs=16000, wlen_sec=64e-3, win="hann", hop_percent=0.25, center=False)


def _case(seconds=1.0):
    x, s, _ = synth.synth_utterance(33, seconds)
    X, S = stft_np.stft(x, **KW), stft_np.stft(s, **KW)
    y = stft_np.clean_speech_ibm(S)                                           # (513, N) in {0, 1}: target.py:58-70
    sd = synth.xavier_state_dict("M2", 513, 16, [128, 128], 513, seed=8, out_bias=synth.speech_prior_bias(s))
    return x, s, X, S, y, sd


def test_decode_with_label_bias_matches_fp32_decoder():
    x, s, X, S, y, sd = _case()
    w = VaeWeights(sd, "M2", torch.device(DEV))
    assert w.y_dim == 513 and w.tc_label_bias and w.tc_y_dim == 0
    N, R = X.shape[1], 10
    yd = torch.from_numpy(np.ascontiguousarray(y.T)).to(DEV)
    z = torch.randn((N * R, 16), device=DEV, generator=torch.Generator(device=DEV).manual_seed(0))
    ref = mlp_forward(w.dec, z, _lib.ACT_EXP, x2=yd, x2_row_div=R)                 # FP32: [z; y] through the full first layer
    eng = McemEngine(w, McemConfig(niter=1, sampler="tc"), DEV)
    P = torch.zeros((N, 520), device=DEV)
    P[:, :513] = torch.from_numpy(np.ascontiguousarray((np.abs(X) ** 2).T)).to(DEV)
    eng.init_parameters(torch.zeros((N, 520), dtype=torch.complex64, device=DEV), P, RaggedBatch([N], DEV), yd)
    assert eng.ybias is not None and tuple(eng.ybias.shape) == (N, 128)
    out = torch.zeros((N * R, 520), device=DEV)
    tc.decode_tc(eng, z, None, R, out, eng.ybias)
    tc.check_status(eng)
    rel = ((out[:, :513] - ref) / ref).abs()
    assert rel.max().item() <= 1e-2 and rel.mean().item() <= 2e-3


@pytest.mark.parametrize("sampler", ["fp32", "tc"])
def test_ibm_conditioned_run_against_oracle(sampler):
    from dvae_b200.packages.models import mcem as shim_mcem
    from dvae_b200.packages.models import models as shim_models
    from dvae_b200.packages.processing.stft import istft
    x, s, X, S, y, sd = _case()
    niter = 12
    torch.manual_seed(5)
    o = mcem_port.MCEMOracle("M2", niter, 10, 30, 25, 75, 0.01)
    o.init_parameters(X, S, sd, 10, 1e-8, y=y)
    cost_ref = o.run()
    s_ref = stft_np.istft(o.S_hat, max_len=len(x), **IKW)
    ref = mcem_port.si_sdr(s_ref[800:-800], s[800:-800])
    model = shim_models.DeepGenerativeModel([513, 513, 16, [128, 128]], None)
    model.load_state_dict({k: torch.tensor(v) for k, v in sd.items()}, strict=False)
    model.to(DEV).eval()
    algo = shim_mcem.MCEM_M2(niter, 10, 30, 25, 75, 0.01, rng="torch", sampler=sampler)
    torch.manual_seed(5)
    algo.init_parameters(X=X, S=S, y=torch.tensor(y, device=DEV), vae=model, nmf_rank=10, eps=1e-8, device=0)
    cost = algo.run()
    assert algo._engine.cfg.sampler == sampler
    s_hat = istft(algo.S_hat, max_len=len(x), **IKW)
    got = mcem_port.si_sdr(s_hat[800:-800], s[800:-800])
    assert abs(got - ref) <= 0.05, (got, ref)
    np.testing.assert_allclose(cost, cost_ref, rtol=2e-2)
    rel = np.linalg.norm(algo.S_hat - o.S_hat) / np.linalg.norm(o.S_hat)
    assert rel <= (2e-2 if sampler == "fp32" else 8e-2), rel


def test_auto_sampler_serves_the_ibm_model_on_tensor_cores():
    _, _, _, _, _, sd = _case(0.5)
    from dvae_b200.engine import tc_supported
    assert tc_supported(VaeWeights(sd, "M2", torch.device(DEV)))
