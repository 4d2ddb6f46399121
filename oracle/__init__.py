"""CPU oracle for the MCEM VAE-NMF enhancement path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain numpy / CPU torch, the algorithm of the reference's
hot path (``packages/processing/stft.py``, ``packages/models/models.py``,
``packages/models/mcem.py``).  It exists so that the CUDA path can be checked
against it; it is never the thing measured or shipped.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``dvae_b200/`` does.

Pinning status
--------------
* ``oracle.mcem_port`` (EM / Metropolis-Hastings / NMF / Wiener / VAE MLPs): **pinned**.
  ``oracle/make_golden.py`` imports the unmodified reference classes from
  ``/root/reference`` (they need only torch + numpy), runs them on seeded inputs with
  recorded random draws, and the port reproduces their outputs bit-for-bit on CPU.  The
  vectors live in ``tests/golden/mcem_*.npz`` and are re-checked by the CPU test-suite.
* ``oracle.stft_np`` (STFT / ISTFT): **parity unpinned by any reference-held vector**.  The
  arithmetic lives in the third-party dependency ``librosa`` (unpinned by the reference;
  0.7 <= version < 0.10 from its call sites), which is not installed here
  and cannot be, and the reference's tests hold no vector at this boundary.  The restatement
  follows librosa's published ``stft``/``istft`` semantics for the one configuration every
  reference caller uses (``center=False``, periodic Hann) and is cross-checked against three
  implementations that share no code with it: ``torch.stft`` in float64 and a direct O(n^2)
  DFT (forward), and ``scipy.signal.stft`` / ``scipy.signal.istft`` with the reference's
  framing -- forward bit-identical after the complex64 cast, inverse equal to float32
  rounding on plain and on Wiener-masked spectrograms wherever librosa and scipy share
  semantics (``tests/test_oracle_stft.py``; the CUDA kernels are also checked against scipy
  directly, ``tests/test_gpu_stft.py``).  librosa-specific edge rules (division by the window
  sum above float32 ``tiny``, ``fix_length``) remain restated from its source only.
"""
