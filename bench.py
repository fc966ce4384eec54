#!/usr/bin/env python3
"""Headline benchmark: audio-seconds processed per second by the batched front end.

Workload (BASELINE.json configs[1]): 100,000 synthetic 1 s utterances per GPU (44.1 kHz int16
PCM, SURVEY.md 8(d) recipe generated on the device), frame 256 / shift 128, the three window
types cycled across steps.  One step = one pass of the fused front end (DC removal, peak
normalisation, endpoint detection, framing + window, energy / magnitude / ZCR, 15 statistics)
over the whole batch.

  python bench.py [--gpus N] [--steps K] [--warmup W]           # our CUDA path
  python bench.py --impl reference ...                          # reference algorithm on host cores
  torchrun --nproc-per-node N ... bench.py --gpus N ...         # N>1: one rank per GPU, weak scaling

Rank 0 prints ONE JSON line.  `value` is device-resident throughput (CUDA events, max over
ranks); `e2e` is the same metric through the host-buffer C-ABI call with the H2D / D2H copies
inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 44100
FL, FS = 256, 128
WINDOWS = ("rectangular", "hamming", "hanning")
UTT_LEN = 44104        # samples per utterance: 1 s rounded up to a multiple of 8 samples, so every
                       # utterance of the packed batch starts 16-byte aligned (TMA bulk copies)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--utts", type=int, default=100000, help="utterances per GPU")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-knn", action="store_true", help="skip the secondary KNN numbers")
    ap.add_argument("--cpu-utts", type=int, default=600, help="utterances in the cpu_baseline sample")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------
def synth_batch_device(n_utts, device, seed, chunk=2048):
    """SURVEY.md 8(d) generator on the GPU: noise floor, 50 ms unvoiced onset, Hann-enveloped
    two-partial burst, DC offset, truncation to int16.  Returns (samples, CSR offsets[n+1])."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    n = UTT_LEN
    out = torch.zeros(n_utts * n + 64, dtype=torch.int16, device=device)
    view = out[: n_utts * n].view(n_utts, n)
    pos = torch.arange(n, device=device, dtype=torch.float32)[None, :]
    t = pos / SR
    on = int(0.050 * SR)
    for c0 in range(0, n_utts, chunk):
        c = min(chunk, n_utts - c0)
        cls = (torch.arange(c0, c0 + c, device=device) % 10).float()[:, None]
        u = torch.rand(c, 3, device=device, generator=gen)
        b0 = torch.floor((0.15 + 0.15 * u[:, 0:1]) * n)
        b1 = torch.floor((0.60 + 0.25 * u[:, 1:2]) * n)
        f0 = 150.0 + 90.0 * cls + (20.0 * u[:, 2:3] - 10.0)
        x = torch.randn(c, n, device=device, generator=gen) * 0.005
        inb = (pos >= b0) & (pos < b1)
        env = 0.5 - 0.5 * torch.cos(2 * np.pi * (pos - b0) / (b1 - b0 - 1).clamp(min=1))
        burst = 0.6 * (torch.sin(2 * np.pi * f0 * t) + 0.3 * torch.sin(2 * np.pi * (2 * f0 + 5 * cls) * t))
        x += inb * env * burst
        onset = (pos >= b0 - on) & (pos < b0)
        x += onset * torch.randn(c, n, device=device, generator=gen) * 0.03
        x += 0.01
        x.clamp_(-1.0, 32767.0 / 32768.0)
        view[c0:c0 + c] = torch.trunc(x * 32768.0).to(torch.int16)
        del x, inb, env, burst, onset
    return out, np.arange(n_utts + 1, dtype=np.int64) * n


class ClockSampler:
    """nvidia-smi SM clock / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """Reference algorithm (oracle port, per-frame NumPy loops) over a slice of utterances."""
    first, count, seed0 = args
    from oracle import frontend_oracle as fo, synth
    utts = [synth.utterance_pcm(first + i, UTT_LEN, seed0) for i in range(count)]
    t0 = time.perf_counter()
    for i, pcm in enumerate(utts):
        fo.frontend_utterance(pcm, FL, FS, WINDOWS[i % 3])
    return time.perf_counter() - t0


def cpu_baseline_single(n_utts):
    dt = _cpu_worker((0, n_utts, 777))
    return {"value": n_utts * (UTT_LEN / SR) / dt, "unit": "audio-s/s", "cores": 1, "kind": "port",
            "sample": f"{n_utts} utterances of the same generator/config (1 s, 256/128, windows cycled), "
                      f"oracle/frontend_oracle.py (NumPy float64, per-frame loops as the reference), {dt:.1f} s"}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm on all host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    for v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[v] = "1"
    cores = os.cpu_count() or 1
    per_worker = 24
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        def step(s):
            jobs = [(s * cores * per_worker + w * per_worker, per_worker, 777) for w in range(cores)]
            t0 = time.perf_counter()
            pool.map(_cpu_worker, jobs)
            return time.perf_counter() - t0
        for s in range(args.warmup):
            step(s)
        total = sum(step(args.warmup + s) for s in range(args.steps))
    n = cores * per_worker * args.steps
    val = n * (UTT_LEN / SR) / total
    line = {
        "impl": "reference", "metric": "audio-seconds processed/sec (features+endpoints)", "value": val,
        "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: batched front end, 1 s utterances, frame 256 / shift 128, three windows cycled",
                   "sample_per_step": f"{cores * per_worker} utterances (bounded sample of the 100k-utterance batch)"},
        "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port",
                         "sample": f"{cores * per_worker} utterances per step x {args.steps} steps, multiprocessing.Pool({cores}), "
                                   "oracle/frontend_oracle.py (the reference is pure Python and does not travel to the GPU box)"},
        "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def knn_section(dev, ctx):
    """Secondary numbers of the same run (not part of `value`): the KNN classify step of the full pipeline
    (BASELINE configs[2] shape on one GPU: 1 M queries x 100 k train rows, D = 15, k = 3 -- fp32 tiled scan + float64
    certificate) and the sequence-feature variant (D = 1024: tcgen05 tensor-core scan), device-resident, CUDA events."""
    import torch
    from dsp_audioreclabs_b200 import device as devapi
    out = {}
    try:
        for tag, (m, n, d) in {"statistical_d15": (1000000, 100000, 15), "sequence_d1024": (131072, 50000, 1024)}.items():
            g = torch.Generator(device=dev).manual_seed(5)
            centers = torch.randn(10, d, device=dev, generator=g, dtype=torch.float64) * 1.5
            ytr = torch.randint(0, 10, (n,), device=dev, generator=g)
            xtr = centers[ytr] + torch.randn(n, d, device=dev, generator=g, dtype=torch.float64)
            yq = torch.randint(0, 10, (m,), device=dev, generator=g)
            xq = centers[yq] + torch.randn(m, d, device=dev, generator=g, dtype=torch.float64)
            knn = devapi.DeviceKNN(3, ctx=ctx, device=dev).fit(xtr.contiguous(), ytr.to(torch.int32))
            knn.predict(xq)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            knn.predict(xq)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            rescanned, kind = knn.last_stats()
            out[tag] = {"queries": m, "train_rows": n, "dim": d, "ms": ms, "queries_per_s": m / (ms / 1e3),
                        "algorithmic_TFLOPs": 2.0 * d * m * n / (ms / 1e3) / 1e12,
                        "scan": {1: "fp32 tiled", 2: "tcgen05 split-fp16 (3 MMA passes)"}.get(kind, "float64"),
                        "rescanned_in_float64": rescanned}
            del xtr, xq, knn
            torch.cuda.empty_cache()
    except Exception as exc:      # secondary numbers must never break the bench line
        out["error"] = repr(exc)
    return out


# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from dsp_audioreclabs_b200 import batch, device as devapi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_utts = args.utts
    ctx = batch.default_context(local)
    # utterance shards: rank r owns utterances [r*n_utts, (r+1)*n_utts) of the global batch; the
    # front end needs no collective
    samples, row_offsets = synth_batch_device(n_utts, dev, seed=1234 + rank)
    stream = torch.cuda.Stream(device=dev)
    frontends = {w: devapi.DeviceFrontend(row_offsets, FL, FS, w, ctx=ctx, device=dev) for w in WINDOWS}
    audio_s_per_step = n_utts * (UTT_LEN / SR)

    def step(i):
        frontends[WINDOWS[i % 3]].run(samples, stream=stream)

    with torch.cuda.stream(stream):
        for i in range(max(args.warmup, 3)):
            step(i)
    barrier()
    # ---- timed region: K steps, CUDA events on the launching stream -------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = ctx.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    torch.cuda.cudart().cudaProfilerStart()      # `ncu --profile-from-start off` lists exactly the timed region's launches
    ev[0].record(stream)
    for i in range(args.steps):
        step(i)
        ev[i + 1].record(stream)
    stream.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max = float(tmax.item())
    value = world * audio_s_per_step * args.steps / (total_ms_max / 1000.0)

    # ---- roofline of the dominant kernel (frontend_pipe_kernel, csrc/frontend_pipe.cu) --------
    per_window = {}
    for wi, w in enumerate(WINDOWS):
        ms = [step_ms[i] for i in range(args.steps) if i % 3 == wi]
        if ms:
            per_window[w] = {"ms": float(np.mean(ms)), "algorithmic_bytes": frontends[w].algorithmic_bytes()}
    alg_bytes = float(np.mean([v["algorithmic_bytes"] for v in per_window.values()]))
    avg_ms = float(np.mean(step_ms))
    peak, peak_src = measured_hbm_peak()
    achieved = alg_bytes / (avg_ms / 1000.0) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "frontend_pipe_kernel<true>", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": avg_ms}

    # ---- parity spot check against the oracle on a few of the benchmarked utterances --------
    parity = None
    if rank == 0:
        from oracle import frontend_oracle as fo
        f = frontends["hamming"]
        f.run(samples, stream=stream)
        stream.synchronize()
        nchk = 24
        host = samples[: nchk * UTT_LEN].cpu().numpy()
        st, en, nf = f.start[:nchk].cpu().numpy(), f.end[:nchk].cpu().numpy(), f.n_frames[:nchk].cpu().numpy()
        zc = f.zcr.cpu().numpy()
        bad = 0
        for b in range(nchk):
            r = fo.frontend_utterance(host[b * UTT_LEN:(b + 1) * UTT_LEN], FL, FS, "hamming")
            o = int(f.h_feat_offsets[b])
            bad += not (r["start"] == st[b] and r["end"] == en[b] and r["n_frames"] == nf[b]
                        and np.array_equal(r["zcr"], zc[o:o + nf[b]].astype(np.float64)))
        parity = {"utterances_checked": nchk, "endpoint_or_zcr_mismatches": int(bad),
                  "replayed_in_float64": int((f.status >= 0x100).sum().item())}

    # ---- e2e: host buffers through dsp_frontend_batch_host (H2D + D2H inside) ----------------
    e2e = None
    if not args.no_e2e:
        h_samples = torch.empty(samples.numel(), dtype=torch.int16, pin_memory=True)
        h_samples.copy_(samples)
        torch.cuda.synchronize()
        hs = h_samples.numpy()
        res = None
        for i in range(1):
            res = batch.frontend_batch(hs, row_offsets, FL, FS, WINDOWS[i % 3], emit_frames=False, ctx=ctx)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.e2e_steps):
            res = batch.frontend_batch(hs, row_offsets, FL, FS, WINDOWS[i % 3], emit_frames=False, ctx=ctx)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        d2h = sum(a.nbytes for a in (res.start, res.end, res.n_epd_frames, res.n_frames, res.status, res.stats))
        e2e = {"value": world * audio_s_per_step * args.e2e_steps / dt, "unit": "audio-s/s",
               "h2d_bytes_per_step": int(hs.nbytes + 3 * row_offsets.nbytes), "d2h_bytes_per_step": int(d2h),
               "steps": args.e2e_steps, "call": "batch.frontend_batch -> dsp_frontend_batch_host (pinned host samples, chunked H2D/compute/D2H overlap; "
                       "result read back = endpoints + frame counts + status + the 15 statistics per utterance)"}
        del h_samples, hs

    if rank == 0:
        line = {
            "metric": "audio-seconds processed/sec (features+endpoints)", "value": value, "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i16->f32/f64", "data": "synthetic",
            "config": {"workload": "configs[1]: batched front end only, 100k synthetic 1 s utterances per GPU, frame 256 / shift 128, three windows cycled over steps",
                       "utterances_per_gpu": n_utts, "samples_per_utterance": UTT_LEN, "frame_length": FL, "frame_shift": FS,
                       "windows": list(WINDOWS), "parallelism": f"utterance shards x{world}, no collective",
                       "l2_policy": f"input {samples.numel() * 2 / 1e9:.2f} GB per pass >> 126 MB L2 (no flush needed)"},
            "roofline": roofline, "per_window": per_window, "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(launches), "parity": parity,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_single(args.cpu_utts)
        if world == 1 and not args.no_knn:
            line["knn"] = knn_section(dev, ctx)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
