"""NumPy restatement of the reference front end (TEST INFRASTRUCTURE ONLY).

Every function names the reference lines it restates (paths are relative to
/root/reference).  The arithmetic is float64 NumPy with the same operation
order as the reference, so that the integer-valued results (zero-crossing
counts, endpoint indices, frame counts) are bit-identical and the float
results agree to the last ulp; tests/test_oracle_golden.py pins this against
fixtures generated from the reference itself (oracle/gen_golden.py).

The per-frame Python loops are kept on purpose: this module is also the
"port" timed as the CPU baseline, and the reference's cost *is* those loops
(SURVEY.md section 3.5).
"""
import numpy as np

WINDOW_IDS = {"rectangular": 0, "hamming": 1, "hanning": 2}
STAT_ORDER = ("mean", "std", "max", "min", "median")
FEATURE_ORDER = ("energy", "magnitude", "zcr")


# ----------------------------------------------------------------------------
# pre-processing                                 src/audio_processing.py:49-90
# ----------------------------------------------------------------------------
def pcm_to_float(pcm):
    """16-bit PCM -> float64 in [-1, 1)  (load_wav, src/audio_processing.py:35-38);
    8-bit unsigned -> (u - 128) / 128  (:31-34)."""
    pcm = np.asarray(pcm)
    if pcm.dtype == np.uint8:
        return (pcm - 128) / 128.0
    return pcm / 32768.0


def stereo_to_mono(x):
    """Interleaved 2-channel float64 -> per-frame mean (src/audio_processing.py:43-44)."""
    return x.reshape(-1, 2).mean(axis=1)


def remove_dc(x):
    """x - mean(x)  (src/audio_processing.py:49-59)."""
    return x - np.mean(x)


def normalize_audio(x):
    """x / max|x| when that maximum is positive, else x (src/audio_processing.py:62-75)."""
    peak = np.max(np.abs(x))
    return x / peak if peak > 0 else x


def preprocess(x):
    """DC removal then peak normalisation (src/audio_processing.py:78-90)."""
    return normalize_audio(remove_dc(x))


# ----------------------------------------------------------------------------
# per-frame primitives                          src/audio_processing.py:93-132
# ----------------------------------------------------------------------------
def short_time_energy(frame):
    """sum(frame**2)  (src/audio_processing.py:93-103)."""
    return np.sum(frame ** 2)


def short_time_magnitude(frame):
    """sum(|frame|) -- a sum, not a mean  (src/audio_processing.py:106-116)."""
    return np.sum(np.abs(frame))


def zero_crossings(frame):
    """Sign changes with zero (and -0.0) counted as negative (src/audio_processing.py:119-132)."""
    s = np.sign(frame)
    s[s == 0] = -1
    return np.sum(np.abs(np.diff(s))) / 2


# ----------------------------------------------------------------------------
# double-threshold endpoint detection          src/audio_processing.py:135-275
# ----------------------------------------------------------------------------
def endpoint_detection(x, frame_length, frame_shift,
                       energy_high_ratio=0.5, energy_low_ratio=0.1,
                       zcr_threshold_ratio=1.5):
    """Returns (start, end, energy_list, zcr_list) exactly as the reference.

    Frame rule (:162-166): no padding, F1 = (L - fl)//fs + 1, unwindowed.
    Thresholds (:186-217, :239-247): noise floor from the first/last
    min(5, F1//10) frames, 90th percentile as the speech level,
    T1 = p90*hr, T2 = noise + (p90 - noise)*lr, T3 = noise_zcr*zr.
    Searches (:205-265): first/last E > T1, outward to the first E <= T2,
    then outward to the first Z <= T3.
    """
    n = len(x)
    if n < frame_length:                                   # :162-163
        return 0, n, np.array([]), np.array([])
    f1 = (n - frame_length) // frame_shift + 1             # :166
    e = np.empty(f1)
    z = np.empty(f1)
    for i in range(f1):                                    # :172-181
        seg = x[i * frame_shift: i * frame_shift + frame_length]
        e[i] = short_time_energy(seg)
        z[i] = zero_crossings(seg)

    nf = min(5, f1 // 10)                                  # :188
    if nf > 0:
        noise_e = np.mean(np.concatenate([e[:nf], e[-nf:]]))   # :190-193
    else:
        noise_e = np.min(e)                                # :195
    speech_e = np.percentile(e, 90)                        # :198
    t1 = speech_e * energy_high_ratio                      # :202
    hot = np.where(e > t1)[0]                              # :205
    if len(hot) == 0:                                      # :207-209
        return 0, n, e, z
    n3, n4 = hot[0], hot[-1]                               # :212-213
    t2 = noise_e + (speech_e - noise_e) * energy_low_ratio  # :217

    n2 = 0                                                 # :220-227
    for i in range(n3 - 1, -1, -1):
        if e[i] <= t2:
            n2 = i + 1
            break
    n5 = f1 - 1                                            # :230-237
    for i in range(n4 + 1, f1):
        if e[i] <= t2:
            n5 = i - 1
            break

    if nf > 0:                                             # :241-247
        noise_z = np.mean(np.concatenate([z[:nf], z[-nf:]]))
    else:
        noise_z = np.min(z)
    t3 = noise_z * zcr_threshold_ratio                     # :249

    n1 = 0                                                 # :252-259
    for i in range(n2 - 1, -1, -1):
        if z[i] <= t3:
            n1 = i + 1
            break
    n6 = f1 - 1                                            # :262-269
    for i in range(n5 + 1, f1):
        if z[i] <= t3:
            n6 = i - 1
            break

    start = int(n1) * frame_shift                          # :272
    end = min(int(n6) * frame_shift + frame_length, n)     # :273
    return start, end, e, z


# ----------------------------------------------------------------------------
# windows and framing                          src/audio_processing.py:278-333
# ----------------------------------------------------------------------------
def make_window(window_type, length):
    """ones / np.hamming / np.hanning; anything else raises (src/audio_processing.py:278-296)."""
    if window_type == "rectangular":
        return np.ones(length)
    if window_type == "hamming":
        return np.hamming(length)
    if window_type == "hanning":
        return np.hanning(length)
    raise ValueError(f"unsupported window type: {window_type}")


def feature_frame_count(n, frame_length, frame_shift):
    """Closed form of the framing loop (:320-331); SURVEY.md A.2."""
    if n <= 0:
        return 0
    a = -(-n // frame_shift)
    b = -(-max(n - frame_length, 0) // frame_shift) + 1
    return min(a, b)


def frame_signal(x, frame_length, frame_shift, window_type="hamming"):
    """Frames at k*fs while k*fs < L; last frame zero padded; each times the
    window; stop after the first frame reaching L (src/audio_processing.py:299-333)."""
    n = len(x)
    if n == 0:
        return np.zeros((0, frame_length))
    w = make_window(window_type, frame_length)
    rows = []
    pos = 0
    while pos < n:
        seg = x[pos: pos + frame_length]
        if len(seg) < frame_length:
            seg = np.pad(seg, (0, frame_length - len(seg)), mode="constant")
        rows.append(seg * w)
        if pos + frame_length >= n:
            break
        pos += frame_shift
    return np.array(rows)


# ----------------------------------------------------------------------------
# per-frame features and statistics              src/feature_extraction.py:12-181
# ----------------------------------------------------------------------------
def frame_features(frames):
    """energy / magnitude / zcr per windowed frame (src/feature_extraction.py:12-43)."""
    nfr = len(frames)
    if nfr == 0:
        raise ValueError("No frames provided for feature extraction.")
    out = {k: np.zeros(nfr) for k in FEATURE_ORDER}
    for i, fr in enumerate(frames):
        out["energy"][i] = short_time_energy(fr)
        out["magnitude"][i] = short_time_magnitude(fr)
        out["zcr"][i] = zero_crossings(fr)
    return out


def sequence_statistics(seq):
    """mean, population std, max, min, median (src/feature_extraction.py:46-62)."""
    return np.array([np.mean(seq), np.std(seq), np.max(seq), np.min(seq), np.median(seq)])


def statistical_vector(ff):
    """15-vector ordered energy_*, magnitude_*, zcr_* (src/feature_extraction.py:65-88)."""
    return np.concatenate([sequence_statistics(ff[k]) for k in FEATURE_ORDER])


def feature_names():
    return [f"{f}_{s}" for f in FEATURE_ORDER for s in STAT_ORDER]


def sequence_matrix(ff, use_only_energy_zcr=False):
    """(F,3) or (F,2) stacked sequences (src/feature_extraction.py:114-129)."""
    keys = ("energy", "zcr") if use_only_energy_zcr else FEATURE_ORDER
    return np.stack([ff[k] for k in keys], axis=1)


def pad_or_truncate(seq, target):
    """Zero-pad rows or cut to `target` rows (src/feature_extraction.py:135-154)."""
    cur = len(seq)
    if cur < target:
        return np.vstack([seq, np.zeros((target - cur, seq.shape[1]))])
    return seq[:target]


def zscore(features, mean=None, std=None):
    """(x - mean)/std over axis 0, std==0 -> 1 (src/feature_extraction.py:157-181)."""
    if mean is None:
        mean = np.mean(features, axis=0)
    if std is None:
        std = np.std(features, axis=0)
    std = np.where(std == 0, 1, std)
    return (features - mean) / std, mean, std


# ----------------------------------------------------------------------------
# whole-utterance pipeline                    src/audio_processing.py:336-396
# ----------------------------------------------------------------------------
def frontend_utterance(audio, frame_length, frame_shift, window_type="hamming",
                       do_endpoint_detection=True, energy_high_ratio=0.5,
                       energy_low_ratio=0.1, zcr_threshold_ratio=1.5):
    """process_audio_file minus the WAV decode, plus extract_frame_features and
    the 15 statistics.  `audio` is int16/uint8 PCM or float64 samples.

    Returns a dict: start, end, n_epd_frames, energy_list, zcr_list, n_frames,
    energy, magnitude, zcr, stats(15).  Raises ValueError exactly where the
    reference does (:388-389, feature_extraction.py:27-28)."""
    a = np.asarray(audio)
    x = pcm_to_float(a) if a.dtype in (np.int16, np.uint8) else a.astype(np.float64, copy=False)
    x = preprocess(x)
    res = {"original_length": len(x)}
    if do_endpoint_detection:
        s, e, el, zl = endpoint_detection(x, frame_length, frame_shift, energy_high_ratio,
                                          energy_low_ratio, zcr_threshold_ratio)
        x = x[s:e]
        res.update(start=s, end=e, energy_list=el, zcr_list=zl, n_epd_frames=len(el))
    else:
        res.update(start=0, end=len(x), energy_list=np.array([]), zcr_list=np.array([]),
                   n_epd_frames=0)
    if len(x) == 0:
        raise ValueError("No audio remaining after preprocessing and endpoint detection.")
    frames = frame_signal(x, frame_length, frame_shift, window_type)
    ff = frame_features(frames)
    res.update(n_frames=len(frames), energy=ff["energy"], magnitude=ff["magnitude"],
               zcr=ff["zcr"], stats=statistical_vector(ff))
    return res


def frontend_batch(samples, offsets, frame_length, frame_shift, window_type="hamming", **kw):
    """Loop `frontend_utterance` over a packed ragged batch; failed utterances
    (ValueError in the reference) yield None, mirroring the callers' per-file
    try/except (experiments/run_experiments.py:88-111)."""
    out = []
    for b in range(len(offsets) - 1):
        try:
            out.append(frontend_utterance(samples[offsets[b]:offsets[b + 1]], frame_length,
                                          frame_shift, window_type, **kw))
        except ValueError:
            out.append(None)
    return out
