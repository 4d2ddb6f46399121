"""Batched file front / back end of the enhancement path.

The reference's ``process_utt`` (scripts/evaluate_ntcd_M1.py:81-188) reads one noisy wav, enhances it and writes
``*_s_est.wav`` / ``*_n_est.wav`` before touching the next file; ``process_sublist`` (190-214) loops over a file list in
one process per half GPU.  Here a sublist is cut into ragged batches for ``Enhancer.enhance``: wav decoding and encoding
run on a few host threads while the GPU works on the previous batch.

wav I/O is self-contained (RIFF PCM 8/16/24/32-bit and IEEE float 32/64, mono or first channel): ``soundfile`` and
``h5py``, which the reference uses, are not required.  ``sf.read`` semantics are kept: float64 samples in [-1, 1);
``write_wav`` writes 16-bit PCM like ``sf.write``'s default for ``.wav``.
"""
from __future__ import annotations

import os
import struct
from concurrent.futures import ThreadPoolExecutor

import numpy as np


def read_wav(path, channel=0):
    """``(samples float64 in [-1, 1), fs)`` of a RIFF/WAVE file (``sf.read`` for the formats listed in the module doc)."""
    with open(path, "rb") as f:
        data = f.read()
    if data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError("%s is not a RIFF/WAVE file" % path)
    pos, fmt, raw = 12, None, None
    while pos + 8 <= len(data):
        tag, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if tag == b"fmt ":
            code, nch, fs, _, _, bits = struct.unpack("<HHIIHH", body[:16])
            if code == 0xFFFE and size >= 26:                     # WAVE_FORMAT_EXTENSIBLE: the sub-format's first two bytes
                code = struct.unpack("<H", body[24:26])[0]
            fmt = (code, nch, fs, bits)
        elif tag == b"data":
            raw = body
        pos += 8 + size + (size & 1)
    if fmt is None or raw is None:
        raise ValueError("%s: missing fmt or data chunk" % path)
    code, nch, fs, bits = fmt
    if code == 1:                                                  # integer PCM
        if bits == 8:
            x = (np.frombuffer(raw, np.uint8).astype(np.float64) - 128.0) / 128.0
        elif bits == 16:
            x = np.frombuffer(raw, "<i2").astype(np.float64) / 32768.0
        elif bits == 24:
            b = np.frombuffer(raw[:len(raw) // 3 * 3], np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            x = np.where(v >= 1 << 23, v - (1 << 24), v).astype(np.float64) / float(1 << 23)
        elif bits == 32:
            x = np.frombuffer(raw, "<i4").astype(np.float64) / 2147483648.0
        else:
            raise ValueError("%s: unsupported PCM width %d" % (path, bits))
    elif code == 3:                                                # IEEE float
        x = np.frombuffer(raw, "<f4" if bits == 32 else "<f8").astype(np.float64)
    else:
        raise ValueError("%s: unsupported wav format tag %d" % (path, code))
    if nch > 1:
        x = x[:len(x) // nch * nch].reshape(-1, nch)[:, channel]
    return np.ascontiguousarray(x), int(fs)


def write_wav(path, x, fs):
    """16-bit PCM mono wav (``sf.write(path, x, fs)`` for float input: scale by 2^15, round, clip)."""
    x = np.asarray(x, np.float64)
    pcm = np.clip(np.rint(x * 32768.0), -32768, 32767).astype("<i2").tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(pcm)) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, int(fs), int(fs) * 2, 2, 16) \
        + b"data" + struct.pack("<I", len(pcm))
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        f.write(hdr + pcm)


def output_stem(output_data_dir, proc_noisy_file_path):
    """``output_data_dir + proc_noisy_file_path`` without extension (scripts/evaluate_ntcd_M1.py:85-86, 172-173)."""
    return os.path.splitext(output_data_dir + proc_noisy_file_path)[0]


def utterance_id(path: str) -> int:
    """A stable 31-bit id of an utterance (CRC-32 of its path): word 0 of the Philox counters, so an utterance's random draws do
    not depend on the batch it lands in, its position in the batch or the rank that processes it."""
    import zlib
    return zlib.crc32(path.encode("utf-8")) & 0x7FFFFFFF


def process_sublist(sublist, enhancer, processed_wav_dir, output_data_dir, batch_size=512, max_frames=None, labels=None,
                    io_threads=8, fs=16000, utt_ids=None):
    """Batched replacement of ``process_sublist`` / ``process_utt`` (scripts/evaluate_ntcd_M1.py:81-214).

    ``sublist``: ``(proc_noisy_file_path, clean_file_path)`` pairs as the reference builds them.  Files whose
    ``*_s_est.wav`` exists are skipped, like the reference.  ``max_frames(noisy_path, clean_path) -> int | None`` supplies the
    video-length frame cap the reference reads from HDF5; ``labels(noisy_path, clean_path) -> (y_dim, N) array`` supplies the
    labels of the M2 models.  ``utt_ids(noisy_path) -> int`` names the utterances for the random number generator
    (default: ``utterance_id``, a hash of the path), so every utterance gets its own draws whatever ``batch_size`` and
    sharding are - the reference draws fresh numbers per utterance too.  File I/O overlaps the GPU work: the next batch is read and
    the previous one written by ``io_threads`` workers while the current one is enhanced.  Returns the list of written stems.
    """
    utt_ids = utt_ids or utterance_id
    todo = [(a, b) for a, b in sublist if not os.path.exists(output_stem(output_data_dir, a) + "_s_est.wav")]
    written = []
    with ThreadPoolExecutor(max(1, io_threads)) as pool:
        pending = None                                               # futures of the previous batch's wav writes
        submit_reads = lambda part: [pool.submit(read_wav, processed_wav_dir + ab[0]) for ab in part]
        reads = submit_reads(todo[:batch_size])
        for lo in range(0, len(todo), batch_size):
            part = todo[lo:lo + batch_size]
            sigs = [f.result() for f in reads]
            reads = submit_reads(todo[lo + batch_size:lo + 2 * batch_size])      # read ahead: the next batch loads while this one runs
            for (x, f), ab in zip(sigs, part):
                if f != fs:
                    raise ValueError("%s: sampling rate %d, expected %d" % (ab[0], f, fs))
            xs = [np.asarray(x, np.float32) for x, _ in sigs]
            caps = [max_frames(*ab) for ab in part] if max_frames else None
            if caps is not None and any(c is None for c in caps):
                big = np.iinfo(np.int64).max
                caps = [big if c is None else c for c in caps]
            ys = [labels(*ab) for ab in part] if labels else None
            s_hat, n_hat, _ = enhancer.enhance(xs, y_list=ys, max_frames_list=caps, utt_ids=[utt_ids(ab[0]) for ab in part])
            if pending:
                for fut in pending:
                    fut.result()
            pending = []
            for ab, s, n in zip(part, s_hat, n_hat):
                stem = output_stem(output_data_dir, ab[0])
                # the result arrays lease the enhancer's pinned buffers (Enhancer._pin_out): the writers hold them until done
                pending.append(pool.submit(write_wav, stem + "_s_est.wav", s, fs))
                pending.append(pool.submit(write_wav, stem + "_n_est.wav", n, fs))
                written.append(stem)
        if pending:
            for fut in pending:
                fut.result()
    return written
