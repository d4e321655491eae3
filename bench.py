#!/usr/bin/env python
"""Headline benchmark: 1-s 16 kHz clips/sec through the fused MFCC front end.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3], SURVEY.md 8d "cfg4"): MFCC with the BASELINE
parameters ("C-MFCC": 40 coefficients, 25 ms / 10 ms, n_fft 512, 128 Slaney mels)
over a 262,144-clip synthetic corpus, sharded by clip across the N GPUs of one box
(strong scaling: the corpus is fixed, each rank owns 262144/N clips; no collective on
the data path).  A step = one pass of the fused kernel over the rank's whole shard.

One JSON line on rank 0:  value = device-resident throughput (clips/s, whole job);
e2e = the same metric through the public API with pinned HOST buffers (H2D and D2H
inside the timed region); roofline = algorithmic bytes / CUDA-event launch time
against the measured HBM peak; cpu_baseline = the CPU oracle (restated librosa path)
on this box's host cores, bounded sample (reported, not a target).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CORPUS_CLIPS = 262144
N_SAMPLES = 16000
METRIC = "1-s 16kHz clips/sec, MFCC+log-mel front end"
WORKLOAD = ("cfg4: MFCC (40 coeffs, win 400 / hop 160 / n_fft 512, 128 slaney mels) over a 262144-clip "
            "synthetic 16 kHz corpus sharded by clip across the GPUs")


def make_config(args, world: int) -> dict:
    """the `config` object of the JSON line -- the same keys and values for both arms (the driver compares them)"""
    per_gpu = (args.clips + world - 1) // world
    return {"workload": WORKLOAD, "preset": args.preset, "corpus_clips": args.clips, "clips_per_gpu": per_gpu,
            "n_samples": N_SAMPLES, "parallelism": f"clip-sharded x{world}, no collective",
            "l2": "inputs (>= 2 GB per GPU) larger than the 126 MB L2; every step re-reads the whole shard",
            "e2e_clips_per_gpu_per_step": min(args.e2e_clips, per_gpu)}


# ---------------------------------------------------------------------------------------
# clocks sampler (recipe: /opt/skills/guides/B200_PROFILING.md)
# ---------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc, self.thr = gpu_index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append((time.time(), line.strip()))
        self.thr = threading.Thread(target=pump, daemon=True)
        self.thr.start()
        t_wait = time.time()
        while not self.rows and time.time() - t_wait < 5.0:      # nvidia-smi takes a moment to print its first row;
            time.sleep(0.02)                                     # short timed regions must still see samples under load

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.rows:
            f = [c.strip() for c in line.split(",")]
            if len(f) < 8:
                continue
            try:
                clk, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            smax = mx
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm and self.rows:          # timed region shorter than one sample: take the nearest
            try:
                sm = [float(self.rows[-1][1].split(",")[1])]
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------
# reference arm: the reference's CPU feature path (restated librosa) on all host cores
# ---------------------------------------------------------------------------------------
def _host_info() -> dict:
    """CPU model of the box and the thread settings the CPU legs ran under (SURVEY 8d asks for them next to the number)."""
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return {"cpu_model": model, "host_cores": os.cpu_count(), "omp_num_threads": os.environ.get("OMP_NUM_THREADS", "unset")}


_WORKER_CLIPS = {}


def _cpu_prepare(args):
    """untimed: synthesise this worker's clips for one step (the native arm builds its corpus outside its timer too)"""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    key, start, n = args
    import oracle
    _WORKER_CLIPS[key] = oracle.synthetic_corpus(n, config_index=3, start=start)
    oracle.mfcc_ref(_WORKER_CLIPS[key][0], oracle.C_MFCC)        # import / table warm-up
    return os.getpid()


def _cpu_worker(key):
    """timed: the reference's per-clip feature call, one clip per call like Network.forward's loop"""
    import oracle
    x = _WORKER_CLIPS.pop(key)
    t0 = time.perf_counter()
    for i in range(x.shape[0]):
        oracle.mfcc_ref(x[i], oracle.C_MFCC)
    return time.perf_counter() - t0


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    per_worker = 128
    ctx = mp.get_context("fork")
    # one single-worker pool per core: a job prepared on worker w is also computed on worker w
    pools = [ctx.Pool(1) for _ in range(cores)]
    try:
        def step(k):
            prep = [pl.apply_async(_cpu_prepare, ((k, (k * cores + w) * per_worker, per_worker),)) for w, pl in enumerate(pools)]
            for r in prep:
                r.get()
            t0 = time.perf_counter()                                 # ---- timed region: feature calls only ----
            res = [pl.apply_async(_cpu_worker, (k,)) for pl in pools]
            worker_s = [r.get() for r in res]
            return time.perf_counter() - t0, max(worker_s)
        for k in range(args.warmup):
            step(k)
        times = [step(args.warmup + k) for k in range(args.steps)]
    finally:
        for pl in pools:
            pl.terminate()
    total = sum(t for t, _ in times)
    clips = cores * per_worker * args.steps
    value = clips / total
    sample = (f"{per_worker} clips x {cores} worker processes per step, oracle.mfcc_ref (restated librosa-0.6 chain), C-MFCC; "
              "timed region = the feature calls only (clip synthesis and worker start-up outside), wall clock around all workers")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(args, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample, **_host_info(),
                         "omp_num_threads": "1 per worker process",
                         "slowest_worker_compute_s_per_step": sum(m for _, m in times) / args.steps},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def cpu_baseline_single(budget_s: float = 12.0) -> dict:
    """Oracle MFCC (C-MFCC) one clip per call, single thread -- how the reference runs
    (DataLoader num_workers=0, training.py:77).  Bounded by wall time."""
    import oracle
    x = oracle.synthetic_corpus(64, config_index=3)
    oracle.mfcc_ref(x[0], oracle.C_MFCC)
    n, t0 = 0, time.perf_counter()
    while True:
        oracle.mfcc_ref(x[n % 64], oracle.C_MFCC)
        n += 1
        el = time.perf_counter() - t0
        if el > budget_s or n >= 32768:
            break
    return {"value": n / el, "unit": "clips/s", "cores": 1, "kind": "port",
            "sample": f"{n} clips of the seeded corpus (config 3), oracle.mfcc_ref one clip per call, {el:.1f} s, 1 thread",
            **_host_info()}


# ---------------------------------------------------------------------------------------
# native arm
# ---------------------------------------------------------------------------------------
def synth_shard(n_clips: int, device, seed: int):
    """int16-valued Gaussian clips, sigma log-uniform in [30, 8000], built on the device."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    x = torch.empty((n_clips, N_SAMPLES), dtype=torch.float32, device=device)
    chunk = 8192
    for i in range(0, n_clips, chunk):
        m = min(chunk, n_clips - i)
        sigma = torch.exp(torch.empty((m, 1), device=device).uniform_(3.4012, 8.9872, generator=g))
        x[i:i + m] = (torch.randn((m, N_SAMPLES), device=device, generator=g) * sigma).clamp_(-32768, 32767).round_()
    return x


def preset_table(S, x, peak: float) -> dict:
    """Every parameter set at the batch sizes BASELINE.json's configs quote (cfg2: C-FBANK B=1024; cfg3: C-SPEC B=4096)
    and at 16,384 clips: CUDA events around every launch; the L2 (126 MB) is flushed before each timed launch by
    rewriting a 256 MB buffer, so small batches cannot be served from cache."""
    import torch
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=x.device)
    cases = [(name, p, None, 16384) for name, p in S.PRESETS.items()]
    cases += [("R-SPEC", S.R_SPEC, "tf", 16384), ("C-SPEC", S.C_SPEC, "tf", 16384),
              ("C-FBANK", S.C_FBANK, None, 1024), ("R-FBANK", S.R_FBANK, None, 1024),
              ("C-SPEC", S.C_SPEC, None, 4096), ("C-SPEC", S.C_SPEC, "tf", 4096),
              ("R-SPEC", S.R_SPEC, None, 4096), ("R-SPEC", S.R_SPEC, "tf", 4096),
              ("R-MFCC", S.R_MFCC, "tf", 8192), ("C-MFCC", S.C_MFCC, None, 64)]
    rows = {}
    for name, p, layout, nclips in cases:
        fam = type(p).__name__
        f2 = {"SpecParams": S.spec, "FbankParams": S.fbank, "MfccParams": S.mfcc}[fam]
        kw = {"layout": layout} if layout else {}
        xs = x[:nclips]
        for _ in range(3):
            f2(xs, p, **kw)
        reps = 10 if nclips >= 4096 else 30
        evs = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            f2(xs, p, **kw)
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in evs)
        med = ms[len(ms) // 2]
        cps = xs.size(0) / (med * 1e-3)
        gbs = cps * S.bytes_per_clip(p, N_SAMPLES) / 1e9
        lay = layout or getattr(p, "layout", "tf")
        rows[f"{name}/{lay}/B{nclips}"] = {"clips_per_s": cps, "GBps": gbs, "hbm_frac": gbs / peak, "ms_median": med,
                                           "ms_best": ms[0], "launches": reps}
    return rows


def run_native(args) -> None:
    import torch
    import torch.distributed as dist
    import speechrecognitionproject_b200 as S
    from speechrecognitionproject_b200.sharding import shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the native arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    preset = S.PRESETS[args.preset]
    fn = {"SpecParams": S.spec, "FbankParams": S.fbank, "MfccParams": S.mfcc}[type(preset).__name__]
    b0, b1 = shard_range(args.clips, rank, world)
    x = synth_shard(b1 - b0, dev, 20260003 + rank)
    bpc = S.bytes_per_clip(preset, N_SAMPLES)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        y = fn(x, preset)
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    # ---- timed region: K steps, device-resident input (16.8 GB corpus >> 126 MB L2) ----
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    n0 = S.launch_count()
    barrier()
    w0 = time.time()
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_begin.record()
    for a, b in ev:
        a.record()
        y = fn(x, preset)
        b.record()
    e_end.record()
    barrier()
    w1 = time.time()
    launches = S.launch_count() - n0
    total_ms = e_begin.elapsed_time(e_end)
    kern_ms = [a.elapsed_time(b) for a, b in ev]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    clocks = sampler.stop(w0, w1) if rank == 0 else None

    # ---- e2e: public API with pinned host buffers, H2D + kernel + D2H inside the timed region ----
    e2e_clips = min(args.e2e_clips, b1 - b0)
    xh = torch.empty((e2e_clips, N_SAMPLES), dtype=torch.float32).pin_memory()
    xh.copy_(x[:e2e_clips])
    for _ in range(2):
        yh = fn(xh, preset)
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        yh = fn(xh, preset)                       # srfe_mfcc_host_f32: chunked H2D -> kernel -> D2H, synchronous
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_clips * e2e_steps / float(te.item())
    d2h = int(yh.numel() * 4)
    # same, with the wav's native int16 samples as the host buffer (srfe_mfcc_host_i16):
    # reported as an extra -- the contract's `e2e` stays the float32 handoff of dataset.py:117
    xh16 = torch.empty((e2e_clips, N_SAMPLES), dtype=torch.int16).pin_memory()
    xh16.copy_(x[:e2e_clips].to(torch.int16))
    for _ in range(2):
        yh = fn(xh16, preset)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        yh = fn(xh16, preset)
    torch.cuda.synchronize()
    te16 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te16, op=dist.ReduceOp.MAX)
    e2e16_value = world * e2e_clips * e2e_steps / float(te16.item())
    # same float32 handoff from PAGEABLE host memory -- what the reference's DataLoader hands over (training.py:77: no
    # pin_memory); libsrfe stages it through its own pinned ring (srfe_abi.cu: run_host)
    xpg = torch.empty((e2e_clips, N_SAMPLES), dtype=torch.float32)
    xpg.copy_(xh)
    assert not xpg.is_pinned()
    for _ in range(2):
        yh = fn(xpg, preset)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        yh = fn(xpg, preset)
    torch.cuda.synchronize()
    tep = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tep, op=dist.ReduceOp.MAX)
    e2e_pageable_value = world * e2e_clips * e2e_steps / float(tep.item())
    # the box's host->device ceiling: the same pinned bytes copied with NO kernel, all ranks at once (e2e is bound by
    # this, not by a collective: aggregate H2D through one host's memory system / PCIe root complexes)
    xdev = torch.empty_like(xh, device=dev)
    for _ in range(2):
        xdev.copy_(xh, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        xdev.copy_(xh, non_blocking=True)
    torch.cuda.synchronize()
    th = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(th, op=dist.ReduceOp.MAX)
    h2d_gbs = world * e2e_clips * N_SAMPLES * 4 * e2e_steps / float(th.item()) / 1e9
    del xdev, xpg

    if rank == 0:
        value = args.clips * args.steps / (total_ms_max * 1e-3)
        avg_kern_ms = sum(kern_ms) / len(kern_ms)
        achieved = (b1 - b0) * bpc / (avg_kern_ms * 1e-3) / 1e9
        peak, peak_src = 6650.0, "fallback"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured"
        except Exception:
            pass
        traffic, traffic_src = None, None
        try:                       # not measurable inside a timed run: taken from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            tr = tj.get(args.preset)
            if tr:
                traffic = tr["dram_bytes_per_clip"] * (b1 - b0)
                traffic_src = ("profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per clip from the ncu "
                               f"--set full capture {tr.get('source', tj.get('source', '?'))}, scaled to this launch; not measured in this run")
        except Exception:
            pass
        out = {
            "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(args, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "bytes_per_clip": bpc,
                         "kernel_ms_avg": avg_kern_ms, "kernel": "srfe_mfcc_tc_kernel<512, ...> (tcgen05 DCT, TMEM accumulators)",
                         "note": "bound by the shared-memory pipe (74 % of peak, ncu) and FP32 issue, not by HBM: DESIGN.md section 5"},
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": int(e2e_clips * N_SAMPLES * 4),
                    "d2h_bytes_per_step": d2h},
            "e2e_int16_ingest": {"value": e2e16_value, "unit": "clips/s", "h2d_bytes_per_step": int(e2e_clips * N_SAMPLES * 2),
                                 "d2h_bytes_per_step": d2h, "note": "extra: int16 host PCM (wav native type), converted in-kernel"},
            "e2e_pageable_host": {"value": e2e_pageable_value, "unit": "clips/s",
                                  "note": "extra: float32 PCM in pageable host memory (the reference's DataLoader handoff, "
                                          "training.py:77), staged through libsrfe's pinned ring"},
            "e2e_h2d_ceiling": {"GBps": h2d_gbs, "clips_per_s": h2d_gbs * 1e9 / (N_SAMPLES * 4), "e2e_frac": e2e_value / (h2d_gbs * 1e9 / (N_SAMPLES * 4)),
                                "note": "pinned float32 H2D copies of the e2e batch with no kernel, all ranks concurrently, "
                                        "max over ranks: the ceiling e2e (float32 ingest) can reach on this box"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_single()
        if world == 1 and not args.no_presets:
            out["presets"] = preset_table(S, x, peak)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--preset", default="C-MFCC")
    ap.add_argument("--clips", type=int, default=CORPUS_CLIPS, help="total corpus size (all GPUs)")
    ap.add_argument("--e2e-clips", type=int, default=16384, help="host-buffer batch per GPU per e2e step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-presets", action="store_true", help="skip the per-preset table (extra key `presets`, N=1 only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
