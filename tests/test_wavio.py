"""CPU: the native batched WAV ingest (csrc/wavio.cpp) accepts, rejects and returns exactly what
CPython's `wave` module does -- the decoder load_wav uses (src/audio_processing.py:21-30).  Host-only I/O: no GPU."""
import os
import struct
import wave

import numpy as np
import pytest


def chunk(name, body):
    return name + struct.pack("<L", len(body)) + body + (b"\0" if len(body) & 1 else b"")


def riff(*chunks, size=None):
    body = b"WAVE" + b"".join(chunks)
    return b"RIFF" + struct.pack("<L", len(body) if size is None else size) + body


def fmt(tag=1, ch=1, rate=44100, bits=16, extra=b""):
    width = (bits + 7) // 8
    return chunk(b"fmt ", struct.pack("<HHLLHH", tag, ch, rate, rate * ch * width, ch * width, bits) + extra)


PCM_GUID = bytes([0x01, 0, 0, 0, 0, 0, 0x10, 0, 0x80, 0, 0, 0xaa, 0, 0x38, 0x9b, 0x71])


def cases():
    rng = np.random.default_rng(5)
    pay = lambda n: rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    ext = struct.pack("<HHL", 22, 16, 4) + PCM_GUID
    bad_ext = struct.pack("<HHL", 22, 16, 4) + bytes(16)
    return {
        "s16_mono": riff(fmt(), chunk(b"data", pay(2000))),
        "s16_stereo": riff(fmt(ch=2), chunk(b"data", pay(2004))),
        "u8_mono": riff(fmt(bits=8), chunk(b"data", pay(777))),
        "u8_stereo": riff(fmt(ch=2, bits=8), chunk(b"data", pay(600))),
        "s24": riff(fmt(bits=24), chunk(b"data", pay(300))),
        "bits12": riff(fmt(bits=12), chunk(b"data", pay(400))),
        "list_first": riff(chunk(b"LIST", pay(37)), fmt(), chunk(b"JUNK", pay(5)), chunk(b"data", pay(1000)), chunk(b"LIST", pay(8))),
        "extensible": riff(fmt(tag=0xFFFE, extra=ext), chunk(b"data", pay(512))),
        "extensible_bad_guid": riff(fmt(tag=0xFFFE, extra=bad_ext), chunk(b"data", pay(512))),
        "float_format": riff(fmt(tag=3, bits=32), chunk(b"data", pay(512))),
        "data_before_fmt": riff(chunk(b"data", pay(100)), fmt()),
        "no_data": riff(fmt()),
        "no_fmt": riff(chunk(b"LIST", pay(10))),
        "zero_channels": riff(fmt(ch=0), chunk(b"data", pay(100))),
        "zero_bits": riff(fmt(bits=0), chunk(b"data", pay(100))),
        "empty_data": riff(fmt(), chunk(b"data", b"")),
        "partial_frame": riff(fmt(ch=2), chunk(b"data", pay(1003))),
        "truncated_payload": riff(fmt(), chunk(b"data", pay(1000)))[:-300],
        "riff_size_short": riff(fmt(), chunk(b"data", pay(1000)), size=4 + 24 + 8 + 500),
        "short_fmt": riff(chunk(b"fmt ", struct.pack("<HHL", 1, 1, 44100)), chunk(b"data", pay(100))),
        "not_riff": b"RIFX" + bytes(40),
        "not_wave": b"RIFF" + struct.pack("<L", 36) + b"AVI " + bytes(32),
        "tiny": b"RIF",
        "empty": b"",
    }


def wave_says(path):
    try:
        with wave.open(path, "rb") as w:
            return (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes(), w.readframes(w.getnframes()))
    except (wave.Error, EOFError):
        return None


@pytest.fixture(scope="module")
def wavio():
    import __graft_entry__ as g
    from dsp_audioreclabs_b200 import _capi
    if not os.path.exists(_capi.LIB_PATH):
        g.build()
    from dsp_audioreclabs_b200 import wavio
    return wavio


def test_scan_and_read_agree_with_the_wave_module(wavio, tmp_path):
    names, paths = [], []
    for name, blob in cases().items():
        p = tmp_path / f"{name}.wav"
        p.write_bytes(blob)
        names.append(name)
        paths.append(str(p))
    paths.append(str(tmp_path / "does_not_exist.wav"))
    names.append("missing")
    groups, info = wavio.read_packed(paths, threads=4)
    got = {}
    for g in groups:
        for j, i in enumerate(g.index):
            assert (g.offsets[j] * g.samples.itemsize) % 16 == 0          # the layout the aligned kernels want
            got[int(i)] = g.clip(j).tobytes()
    n_ok = 0
    for i, (name, path) in enumerate(zip(names, paths)):
        ref = wave_says(path) if name != "missing" else None
        w = info[i]
        assert (w.status == 0) == (ref is not None), (name, w.status)
        if ref is None:
            assert i not in got
            continue
        n_ok += 1
        assert (w.channels, w.sample_width, w.sample_rate, w.n_frames) == ref[:4], name
        assert w.data_bytes == len(ref[4]), name
        if w.sample_width in (1, 2) and len(ref[4]) % w.sample_width == 0:
            assert got[i] == ref[4], name
        else:
            assert i not in got, name                                       # load_wav raises for these (:39-40)
    assert n_ok >= 12 and info[len(paths) - 1].status == 1


def test_single_file_errors_match_load_wav(wavio, tmp_path):
    c = cases()
    for name, exc in (("s24", ValueError), ("not_riff", ValueError), ("float_format", ValueError)):
        p = tmp_path / f"{name}.wav"
        p.write_bytes(c[name])
        with pytest.raises(exc):
            wavio.read_wav_pcm(str(p))
    with pytest.raises(FileNotFoundError):
        wavio.read_wav_pcm(str(tmp_path / "nope.wav"))
    p = tmp_path / "ok.wav"
    p.write_bytes(c["s16_stereo"])
    pcm, sr, ch = wavio.read_wav_pcm(str(p))
    ref = wave_says(str(p))
    assert (sr, ch) == (44100, 2) and pcm.dtype == np.int16 and pcm.tobytes() == ref[4]
