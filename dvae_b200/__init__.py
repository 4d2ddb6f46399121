"""dvae_b200 -- B200-native (sm_100a) MCEM VAE-NMF speech enhancement behind the reference's call signatures.

Layout:
  csrc/                       CUDA kernels + the C ABI (libdvae_b200.so, declared in include/dvae_b200.h)
  _lib.py                     ctypes binding (lazy; raises when the library or a CUDA device is missing)
  engine.py                   batched host driver: ragged batches, EM loop, end-to-end Enhancer
  packages/models/*.py        VAE classes and MCEM_* with the reference's names and signatures
  packages/processing/stft.py stft / istft with the reference's signatures
  synth.py                    seeded synthetic utterances / labels / weights
"""
__version__ = "0.1.0"
