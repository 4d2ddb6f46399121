"""Build libdvae_b200.so in-tree with nvcc for sm_100a (no other architecture, no fallback).

    python -m dvae_b200.build [--force]

The shared library lands next to the sources (dvae_b200/libdvae_b200.so); it is git-ignored but travels to the
GPU box with the repository snapshot.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdvae_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xptxas=-v",
    "-Xcompiler", "-fPIC",
]
OBJ = os.path.join(HERE, "_obj")               # per-source objects (git-ignored); the sources are compiled in parallel


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


MANIFEST = os.path.join(HERE, "libdvae_b200.manifest.json")


def _sha256(path: str) -> str:
    import hashlib
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def _deps():
    return sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(os.path.dirname(HERE), "include", "dvae_b200.h")]


def source_hashes() -> dict:
    return {os.path.relpath(d, os.path.dirname(HERE)): _sha256(d) for d in _deps()}


def verify() -> dict:
    """Check that the shared library on disk is the one built from the sources on disk: the manifest written at link time
    records the SHA-256 of every source / header and of the library itself (the library travels to the GPU box as a binary).
    Returns the manifest; raises RuntimeError on any mismatch."""
    import json
    if not (os.path.exists(LIB) and os.path.exists(MANIFEST)):
        raise RuntimeError("libdvae_b200.so or its manifest is missing: run `python -m dvae_b200.build`")
    with open(MANIFEST) as f:
        man = json.load(f)
    if man.get("library_sha256") != _sha256(LIB):
        raise RuntimeError("libdvae_b200.so does not match its build manifest")
    if man.get("sources") != source_hashes():
        raise RuntimeError("libdvae_b200.so was built from different sources than the ones on disk")
    return man


def needs_build() -> bool:
    """True unless the library's manifest matches the sources on disk (content hashes, not modification times: the tree is
    copied to the GPU box, which does not preserve them)."""
    try:
        verify()
        return False
    except RuntimeError:
        return True


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libdvae_b200.so cannot be built (there is no CPU fallback)")
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = max(os.path.getmtime(d) for d in glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(os.path.dirname(HERE), "include", "dvae_b200.h")])

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return obj, 0, ""                      # object cache of the build container (modification times are reliable here)
        proc = subprocess.run([nvcc] + NVCC_FLAGS + ["-c", "-o", obj, src], capture_output=True, text=True)
        return obj, proc.returncode, proc.stdout + proc.stderr

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max(1, min(8, os.cpu_count() or 1))) as pool:
        results = list(pool.map(compile_one, sources()))
    log = "".join(r[2] for r in results)
    failed = [r for r in results if r[1] != 0]
    if not failed:
        proc = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + [r[0] for r in results],
                              capture_output=True, text=True)
        log += proc.stdout + proc.stderr
        if proc.returncode != 0:
            failed = [("link", proc.returncode, "")]
    if verbose or failed:
        sys.stderr.write(log)
    if failed:
        raise RuntimeError("nvcc failed (%s)" % ", ".join("%s: exit %d" % (os.path.basename(r[0]), r[1]) for r in failed))
    import json
    with open(MANIFEST, "w") as f:
        json.dump({"library_sha256": _sha256(LIB), "sources": source_hashes(), "nvcc_flags": NVCC_FLAGS}, f, indent=1, sort_keys=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
