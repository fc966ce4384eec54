"""Batched WAV ingest: native header scan + payload reads into one
packed, 16-byte aligned staging buffer (pinned when a CUDA device is present) -- the layout
`batch.frontend_batch(..., lengths=...)` uploads as is.  Stands in for the per-file
`wave.open(...).readframes(...)` + `np.frombuffer` of load_wav (src/audio_processing.py:21-40);
the reference's refusal rules are kept: files `wave` would reject are reported with a reason,
sample widths other than 1 and 2 raise the same ValueError (:39-40) when a single file is asked for
and are skipped in a batch, as the reference's callers do (run_experiments.py:109-111)."""
import ctypes as C
import os
import weakref

import numpy as np

from . import _capi

WAV_ERRORS = {
    1: "cannot open file", 2: "file does not start with RIFF id", 3: "not a WAVE file", 4: "unknown format",
    5: "bad sample width", 6: "bad # of channels", 7: "data chunk before fmt chunk",
    8: "fmt chunk and/or data chunk missing", 9: "truncated file",
}
ALIGN = 16


class HostBuffer:
    """Page-locked staging memory from dsp_host_alloc, exposed as a uint8 ndarray; pageable NumPy
    memory when no CUDA device is present (header-only use on a CPU box)."""

    def __init__(self, nbytes):
        lib = _capi.load_library()
        self.nbytes = int(nbytes)
        self.pinned = False
        self.array = None
        if self.nbytes:
            p = C.c_void_p()
            if lib.dsp_host_alloc(self.nbytes, C.byref(p)) == 0 and p.value:
                self.pinned = True
                self.array = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(self.nbytes,))
                weakref.finalize(self, lib.dsp_host_free, p)
        if self.array is None:
            self.array = np.zeros(self.nbytes, dtype=np.uint8)


def _path_array(paths):
    enc = [os.fsencode(p) for p in paths]
    arr = (C.c_char_p * len(enc))(*enc)
    return arr, enc


def scan(paths, threads=None):
    """Header fields of every file: a (n,) structured view of dsp_wav_info."""
    lib = _capi.load_library()
    n = len(paths)
    info = (_capi.WavInfo * max(n, 1))()
    arr, _keep = _path_array(paths)
    _capi.check(lib.dsp_wav_scan(arr, n, int(threads or min(os.cpu_count() or 1, 16)), info))
    return info


class PackedGroup:
    """Files of one encoding packed back to back: `samples` (int16 or uint8, interleaved channels as stored),
    `offsets` / `lengths` in array elements (16-byte aligned starts), `index` = positions in the file list."""

    def __init__(self, dtype, channels, index, samples, offsets, lengths, buffer):
        self.dtype, self.channels, self.index = dtype, channels, index
        self.samples, self.offsets, self.lengths, self._buffer = samples, offsets, lengths, buffer

    def clip(self, j):
        return self.samples[self.offsets[j]: self.offsets[j] + self.lengths[j]]


def read_packed(paths, threads=None):
    """-> (groups, info): every decodable file of `paths` read into one staging buffer per
    (sample width, channels) group.  `info[i].status != 0` or an unsupported width means file i is in no group."""
    lib = _capi.load_library()
    threads = int(threads or min(os.cpu_count() or 1, 16))
    info = scan(paths, threads)
    n = len(paths)
    by_enc = {}
    for i in range(n):
        w = info[i]
        if w.status == 0 and w.sample_width in (1, 2) and w.data_bytes % w.sample_width == 0:
            # load_wav only down-mixes n_channels == 2 (src/audio_processing.py:43-44); any other count is used as the
            # flat interleaved array it is, i.e. as one channel
            by_enc.setdefault((w.sample_width, 2 if w.channels == 2 else 1), []).append(i)
    groups = []
    arr, _keep = _path_array(paths)
    for (width, ch), idx in sorted(by_enc.items()):
        nbytes = np.array([info[i].data_bytes for i in idx], dtype=np.int64)
        starts = np.zeros(len(idx) + 1, dtype=np.int64)
        np.cumsum((nbytes + ALIGN - 1) // ALIGN * ALIGN, out=starts[1:])
        buf = HostBuffer(int(starts[-1]) + ALIGN)
        dst_off = np.full(n, -1, dtype=np.int64)
        dst_off[idx] = starts[:-1]
        sub = (_capi.WavInfo * n)()
        C.memmove(sub, info, C.sizeof(_capi.WavInfo) * n)
        for i in range(n):
            if dst_off[i] < 0:
                sub[i].status = -1          # not in this group: skipped by dsp_wav_read
        _capi.check(lib.dsp_wav_read(arr, n, threads, sub, dst_off.ctypes.data, buf.array.ctypes.data, buf.nbytes))
        ok = [j for j, i in enumerate(idx) if sub[i].status == 0]
        for j, i in enumerate(idx):
            if sub[i].status != 0:
                info[i].status = sub[i].status
        dt = np.int16 if width == 2 else np.uint8
        samples = buf.array.view(dt)
        offsets = np.append(starts[:-1][ok], starts[-1]) // width      # B+1 entries; explicit lengths travel beside them
        groups.append(PackedGroup(dt, ch, np.array([idx[j] for j in ok], dtype=np.int64), samples,
                                  offsets.astype(np.int64), (nbytes[ok] // width).astype(np.int32), buf))
    return groups, info


def read_wav_pcm(filepath):
    """One file -> (pcm as stored, sample_rate, channels) with load_wav's errors (src/audio_processing.py:21-40)."""
    groups, info = read_packed([filepath], threads=1)
    w = info[0]
    if w.status != 0:
        if w.status == 1:
            raise FileNotFoundError(filepath)
        raise ValueError(f"{filepath}: {WAV_ERRORS.get(w.status, 'invalid WAV file')}")
    if w.sample_width not in (1, 2):
        raise ValueError(f"不支持的采样位数: {w.sample_width}")
    if not groups:
        raise ValueError("buffer size must be a multiple of element size")
    g = groups[0]
    return np.array(g.clip(0)), w.sample_rate, w.channels
