#!/usr/bin/env python
"""bench.py -- throughput of the B200 Zernike hot path (one JSON line on stdout, rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload patches|map|map4k|c3|c5] [--impl reference]

BASELINE.json metric: "Zernike patches/sec (n_max=12, 64px) and symmetry-map Mpix/sec at 1/2/4/8 B200".
Workloads (one step = one pass of the hot path over one batch of synthetic input resident in HBM):
  patches  metric shape: ZPs(12, 64) real moments of 262 144 lattice patches per GPU (4.3 GB >> L2)     [default]
  map      BASELINE configs[1]: fused symmetry map (folds 2,3,4,6) of a 2048^2 frame, 48-px window, one frame per GPU
  map4k    BASELINE configs[3]: ONE 4096^2 frame with defects, 64-px window, output row bands over the GPUs
  c3       BASELINE configs[2]: complex ZPs n_max=20 of 1 Mi 64x64 patches, patch ranges over the GPUs
  c5       BASELINE configs[4]: 256 frames 2048^2: local_max -> KeyPoints -> |Zc| features, frames over the GPUs
The line of the chosen workload carries every other one under "also" (a list), so one run records all five.

Multi-GPU (torchrun, one rank per GPU): every rank works on its own shard; the timed step INCLUDES the final
feature / score gather to every rank over NVLink -- `value` is with the gather, the
`gather` object keeps the no-gather figure beside it.  Timing: CUDA events on the launching stream, barrier +
synchronize on both sides, max over ranks.  Every workload asserts parity of a sample of the timed batch's
results against the CPU oracle before it reports a number.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_MAX = 12
PATCH = 64
FOLDS = [2, 3, 4, 6]
C3_NMAX, C3_TOTAL = 20, 1 << 20
C5_FRAMES, C5_SIZE = 256, 2048
PEAK_MIN_DISTANCE, PEAK_THRESHOLD = 5.0, 0.3
WORKLOADS = ["patches", "map", "map4k", "c3", "c5"]
PREC_NAMES = {0: "fp32", 1: "tf32", 2: "tf32x3", 3: "f16", 4: "f16x3"}
DTYPES = {"fp32": "f32", "tf32": "tf32", "tf32x3": "tf32x3(f32-grade)", "f16": "f16", "f16x3": "f16x3(f32-grade)"}
# DRAM bytes per launch from the committed ncu --set full captures (dram__bytes_read.sum + dram__bytes_write.sum)
NCU_TRAFFIC = {"tf32x3": 4.198750e9 + 43.819e6, "tf32": 4.181859e9 + 53.549e6,            # 262 144 patches
               "f16x3": 4.163302e9 + 56.479e6}       # project_fold_kernel<0,0>, profiles/r02_prof_fold_raw.csv
NCU_TRAFFIC_C3_262144 = 4.175737e9 + 215.212e6     # project_fold_kernel<0,1> on 262 144 patches, profiles/r02_prof_c3_raw.csv
NCU_TRAFFIC_MAP = {"tf32x3": 136.456e6 + 48.684e6, "f16x3": 136.000e6 + 47.324e6}        # 2048^2, k=48


def peaks():
    out = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        out = {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
               "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    # tf32 / fp32-SIMT peaks measured by this repo the same way (scripts/measure_peaks.py -> profiles/peaks_tf32.json)
    out["tf32_tflops"], out["tf32_source"] = out["bf16_tflops"] / 2.0, "bf16 / 2 (assumed)"
    mine = os.path.join(ROOT, "profiles", "peaks_tf32.json")
    if os.path.exists(mine):
        with open(mine) as f:
            q = json.load(f)
        if q.get("tf32_tflops"):
            out["tf32_tflops"], out["tf32_source"] = float(q["tf32_tflops"]), "profiles/peaks_tf32.json (measured)"
        if q.get("fp32_simt_tflops"):
            out["fp32_simt_tflops"] = float(q["fp32_simt_tflops"])
    return out


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.002):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._ready = threading.Event()
        self._thread = None

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            names = {
                getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(pynvml, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
            }
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons",
                                  getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons", None))
            self._ready.set()
            tick = 0
            while not self._stop.is_set():
                self.samples.append(int(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                if get_reasons is not None and tick % 4 == 0:          # the reasons query is the slow NVML call on some boxes
                    mask = int(get_reasons(h))
                    for bit, name in names.items():
                        if mask & bit:
                            self.reasons.add(name)
                tick += 1
                time.sleep(self.period)
        except Exception as exc:  # NVML missing: report that rather than fail the bench
            self.reasons.add(f"nvml_unavailable:{type(exc).__name__}")
        finally:
            self._ready.set()

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        self._ready.wait(timeout=5)          # NVML is initialised before the timed region starts
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thread.join(timeout=2)

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_mhz_min": min(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except (ValueError, IndexError):
            pass
    return local


# --------------------------------------------------------------------------- shared configs (both arms print the same dict)
def workload_config(name: str, world: int):
    if name == "patches":
        return {"workload": f"patches n_max={N_MAX} size={PATCH} (metric shape)", "modes": 91, "batch_per_gpu": 262144,
                "l2": "input batch larger than L2 (no flush needed)",
                "parallelism": f"patch shards x{world}" + (", final all-gather of the (N,91) moments over NVLink in the timed step" if world > 1 else "")}
    if name == "map":
        return {"workload": f"symmetry map 2048x2048 n_max={N_MAX} window=48 folds={FOLDS} (BASELINE configs[1])",
                "l2": "256 MiB scratch write between steps", "parallelism": f"one frame per GPU x{world}, no collective"}
    if name == "map4k":
        return {"workload": f"symmetry map 4096x4096 with defects n_max={N_MAX} window=64 folds={FOLDS} (BASELINE configs[3])",
                "l2": "256 MiB scratch write between steps",
                "parallelism": f"one frame, {world} row bands with halo rows from the replicated frame"
                               + (", score bands all-gathered over NVLink in the timed step" if world > 1 else "")}
    if name == "c3":
        return {"workload": f"complex ZPs n_max={C3_NMAX} size={PATCH}, {C3_TOTAL} patches (BASELINE configs[2])", "modes": 231,
                "complex_modes": 121, "l2": "input larger than L2",
                "parallelism": f"{C3_TOTAL} patches in {world} contiguous shards"
                               + (", final all-gather of the complex64 (N,121) features in the timed step" if world > 1 else "")}
    if name == "c5":
        return {"workload": f"{C5_FRAMES} frames {C5_SIZE}x{C5_SIZE}: local_max -> KeyPoints -> ZPs({N_MAX},{PATCH}) |Zc| "
                            f"(BASELINE configs[4])", "l2": "frame series larger than L2",
                "parallelism": f"{C5_FRAMES} frames in {world} shards"
                               + (", ragged all-gather of the (P,49) features in the timed step" if world > 1 else "")}
    raise ValueError(name)


METRICS = {"patches": ("zernike_patches_per_sec", "patches/s"), "map": ("symmetry_map_mpix_per_sec", "Mpix/s"),
           "map4k": ("symmetry_map_mpix_per_sec", "Mpix/s"), "c3": ("zernike_patches_per_sec", "patches/s"),
           "c5": ("zernike_patches_per_sec", "patches/s")}


# --------------------------------------------------------------------------- CPU arm (oracle port of the reference algorithm)
def cpu_threads():
    """Give the BLAS behind numpy.dot every host core (torchrun exports OMP_NUM_THREADS=1) and
    return the thread count actually in use."""
    try:
        import numpy  # noqa: F401  (loads the BLAS whose pool is configured)
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return 1


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import zernike_oracle as zo
    return zo


_cpu_cache = {}


def cpu_patch_sample(n_sample: int, size: int = PATCH):
    """n_sample float32 lattice patches (the bench batch's generator at a smaller frame)."""
    key = ("patches", n_sample, size)
    if key not in _cpu_cache:
        import numpy as np
        zo = _oracle()
        from motif_learn_b200.datasets import honeycomb_image
        img, pts = honeycomb_image(1024, bond=12.0, seed=0)
        base = zo.extract_patches(img, zo.clear_border(pts, img.shape, size), size)
        reps = -(-n_sample // len(base))
        _cpu_cache[key] = np.concatenate([base] * reps)[:n_sample]
    return _cpu_cache[key]


def cpu_patches(n_sample: int, repeats: int, n_max: int = N_MAX, kind: str = "real"):
    """The reference's CPU algorithm for the patch path (numpy.dot in float64 incl. the float32->float64 cast and
    the zmoments constructor copy, _zps.py:146-157, _zmoments.py:269-277; kind='complex': + to_complex,
    _zmoments.py:300-316), through the oracle port.  Returns (patches/s best, [seconds])."""
    import numpy as np
    zo = _oracle()
    sample = cpu_patch_sample(n_sample)
    key = ("basis", n_max, PATCH)
    if key not in _cpu_cache:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            _cpu_cache[key] = zo.zernike_basis(n_max, PATCH)
    n, m, basis = _cpu_cache[key]
    zo.project_patches(sample[:256], basis)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        z = zo.project_patches(sample, basis)
        z = z[:, np.lexsort((m, n))]                      # the reference's constructor always copies
        if kind == "complex":
            zo.to_complex(z, n, m)
        times.append(time.perf_counter() - t0)
    return n_sample / min(times), times


def cpu_map(size: int, repeats: int = 1, window: int = 48):
    """The reference's CPU algorithm for the map path (scipy fftconvolve per mode + rot_maps,
    _zps.py:159-193, _zmoments.py:420-462) through the oracle port on a size x size frame."""
    import numpy as np
    zo = _oracle()
    from motif_learn_b200.datasets import honeycomb_image
    img, _ = honeycomb_image(size, bond=12.0, seed=0)
    n, m, basis = zo.zernike_basis(N_MAX, window)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        z = zo.moment_map_fft(img.astype(np.float64), basis, n)
        zo.rot_maps(z, n, m, FOLDS)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return size * size / 1e6 / best, best


def cpu_c5(size: int = 1024):
    """Reference pipeline of one frame on the host: local_max -> clear_border -> extract_patches -> numpy.dot ->
    to_complex -> abs (oracle port; _local_max_v2.py:46-66, _keypoint.py:44-78, _zps.py:146-157)."""
    import numpy as np
    zo = _oracle()
    from motif_learn_b200.datasets import honeycomb_image
    img, _ = honeycomb_image(size, bond=12.0, seed=0, jitter=0.3, noise=0.01)
    n, m, basis = zo.zernike_basis(N_MAX, PATCH)
    t0 = time.perf_counter()
    pk = zo.local_max(img, PEAK_MIN_DISTANCE, PEAK_THRESHOLD)
    kept = zo.clear_border(pk, img.shape, PATCH)
    patches = zo.extract_patches(img, kept, PATCH)
    z = zo.project_patches(patches, basis)
    np.abs(zo.to_complex(z, n, m)[0])
    dt = time.perf_counter() - t0
    return len(kept) / dt, dt, len(kept)


def cpu_baseline_for(name: str, cores: int):
    if name == "patches":
        v, times = cpu_patches(20000, 8)
        return {"value": v, "unit": "patches/s", "cores": cores, "kind": "port",
                "sample": "20000 float32 64x64 lattice patches, numpy.dot float64 (reference algorithm _zps.py:146-157 "
                          f"+ zmoments ctor copy, via the oracle port), best of {len(times)}"}
    if name == "c3":
        v, times = cpu_patches(20000, 4, n_max=C3_NMAX, kind="complex")
        return {"value": v, "unit": "patches/s", "cores": cores, "kind": "port",
                "sample": f"20000 float32 64x64 patches, n_max={C3_NMAX}: numpy.dot float64 + to_complex, best of {len(times)}"}
    if name == "c5":
        v, dt, cnt = cpu_c5(1024)
        return {"value": v, "unit": "patches/s", "cores": cores, "kind": "port",
                "sample": f"one 1024x1024 frame ({cnt} kept peaks): local_max + extract_patches + numpy.dot + to_complex + abs, {dt:.2f} s"}
    window = 64 if name == "map4k" else 48
    size = 384 if name == "map" else 512
    mv, dt = cpu_map(size, window=window)
    return {"value": mv, "unit": "Mpix/s", "cores": 1, "kind": "port",
            "sample": f"{size}x{size} frame, window {window}: 91 scipy.fftconvolve (1 thread, reference algorithm "
                      f"_zps.py:159-193 via the oracle port) + rot_maps, {dt:.1f} s"}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm of the chosen workload on the host cores of rank 0, each step a
    bounded sample of the workload (the oracle port; the Python reference itself does not exist on the GPU box)."""
    if rank != 0:
        return
    cores = cpu_threads()
    name = args.workload
    metric, unit = METRICS[name]
    per_step, units = [], 0
    if name in ("patches", "c3"):
        n_sample = 20000
        nm, kind = (N_MAX, "real") if name == "patches" else (C3_NMAX, "complex")
        for _ in range(args.warmup):
            cpu_patches(n_sample, 1, nm, kind)
        for _ in range(args.steps):
            _, t = cpu_patches(n_sample, 1, nm, kind)
            per_step.append(t[0])
        units = n_sample
        sample = f"{n_sample} float32 64x64 lattice patches per step, numpy.dot float64" + (" + to_complex" if kind == "complex" else "")
    elif name == "c5":
        for _ in range(min(args.warmup, 1)):
            cpu_c5(512)
        steps = max(1, min(args.steps, 5))
        for _ in range(steps):
            _, dt, cnt = cpu_c5(1024)
            per_step.append(dt)
            units = cnt
        sample = f"one 1024x1024 frame per step ({units} peaks): local_max + extract_patches + numpy.dot + |to_complex|"
    else:
        size, window = 512, (64 if name == "map4k" else 48)
        for _ in range(min(args.warmup, 1)):
            cpu_map(256, window=window)
        for _ in range(args.steps):
            _, dt = cpu_map(size, window=window)
            per_step.append(dt)
        units = size * size / 1e6
        sample = f"{size}x{size} frame per step: 91 scipy.fftconvolve (1 thread) + rot_maps"
        cores = 1
    total = sum(per_step)
    value = units * len(per_step) / total
    line = {"metric": metric, "value": value, "unit": unit, "impl": "reference", "n_gpus": world, "steps": len(per_step),
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(per_step), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(name, world),
            "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm: timing
class Ctx:
    def __init__(self, torch, dist, rank, world, local, args, pk):
        self.torch, self.dist, self.rank, self.world, self.local, self.args, self.pk = torch, dist, rank, world, local, args, pk
        self.gpu_index = physical_gpu_index(local)
        self.frame_max = 1.0

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())


def timed(ctx, fn, steps, warmup, between=None):
    """W warm-up steps, then exactly K timed steps with CUDA events on the launching stream, barrier +
    synchronize on both sides, max over ranks.  Returns (ms_total, clocks, launches)."""
    torch, dist, world = ctx.torch, ctx.dist, ctx.world
    from motif_learn_b200 import _lib
    for _ in range(warmup):
        fn()
        if between:
            between()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    _lib.reset_launch_count()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    with ClockSampler(ctx.gpu_index) as clk:
        for i in range(steps):
            ev0[i].record()
            fn()
            ev1[i].record()
            if between:
                between()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    launches = _lib.launch_count()
    # no L2 flush between steps: one interval over all K steps (launch gaps included);
    # with a flush between steps: sum of the per-step intervals (the flush is not the workload)
    ms = ev0[0].elapsed_time(ev1[-1]) if between is None else sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    return ctx.max_over_ranks(ms), clk.summary(), launches


def fp32_close(got, ref, what, rtol=1e-4, scale=1e-6):
    """The parity gate of SURVEY.md 8d: allclose(rtol=1e-4, atol=1e-6 * max|ref|)."""
    import numpy as np
    tol = scale * np.abs(ref).max()
    bad = np.abs(got - ref) > tol + rtol * np.abs(ref)
    if bad.any():
        raise AssertionError(f"parity FAILED for {what}: {int(bad.sum())} of {bad.size} values outside rtol={rtol} "
                             f"atol={scale}*max; max abs err {np.abs(got - ref).max():.3e}, max|ref| {np.abs(ref).max():.3e}")
    return float(np.abs(got - ref).max() / np.abs(ref).max())


def sample_rows(n: int, count: int = 4096):
    """Row indices of a parity sample: first / middle (shard seam of a 2-GPU run) / last chunks."""
    import numpy as np
    c = min(count // 3, n // 3)
    mid = n // 2 - c // 2
    return np.concatenate([np.arange(0, c), np.arange(mid, mid + c), np.arange(n - c, n)])


def value_bound(ctx, frame_max: float):
    """--value-max: None -> no range hint ('auto' = tf32x3); 'frame' -> the maximum of the frame the patches were cut
    from (what a K2 -> K3 pipeline knows for free); a number -> that bound."""
    v = ctx.args.value_max
    if v in (None, "none", "None"):
        return None
    return float(frame_max) if v == "frame" else float(v)


def make_patch_batch(ctx, n: int, seed: int):
    """n device-resident 64x64 lattice patches: gathered (K2) at the atom sites of a 2048^2 frame, tiled to n."""
    torch = ctx.torch
    from motif_learn_b200.datasets import honeycomb_frame_gpu
    from motif_learn_b200.features import KeyPoints
    img, pts = honeycomb_frame_gpu(2048, bond=12.0, seed=seed)
    ctx.frame_max = float(img.abs().max())
    base = KeyPoints(pts, img, PATCH).extract_patches()            # ~21 k patches, device-resident
    reps = -(-n // base.shape[0])
    out = torch.empty((n, PATCH, PATCH), dtype=torch.float32, device="cuda")
    for r in range(reps):
        lo = r * base.shape[0]
        hi = min(n, lo + base.shape[0])
        out[lo:hi] = base[: hi - lo]
    return out


def gather_report(ctx, how, ms_full, ms_nogather, ms_nccl, steps, units_per_step, nbytes_in):
    """What the K5 gather cost, whole job: the timed step (`value`) uses `how`; beside it the same step without any
    gather and with a plain NCCL all-gather after the kernel (the baseline the fused push is measured against)."""
    out = {"how": how, "in_timed_step": True,
           "ms_per_step_with_gather": ms_full / steps, "ms_per_step_no_gather": ms_nogather / steps,
           "ms_gather": (ms_full - ms_nogather) / steps, "value_no_gather": units_per_step * steps / (ms_nogather / 1e3),
           "bytes_received_per_rank_per_step": nbytes_in,
           "receive_gbs_per_rank": nbytes_in * steps / (ms_full / 1e3) / 1e9}
    if ms_nccl is not None:
        out["nccl_after_kernel"] = {"ms_per_step": ms_nccl / steps, "value": units_per_step * steps / (ms_nccl / 1e3),
                                    "what": "same kernel, then all_gather_into_tensor (NCCL) on the same stream"}
    return out


# --------------------------------------------------------------------------- patches (metric shape)
def bench_patches(ctx):
    import numpy as np
    torch, args, pk, world = ctx.torch, ctx.args, ctx.pk, ctx.world
    from motif_learn_b200.features import ZPs
    from motif_learn_b200.parallel import PeerArray, gather_rows
    zo = _oracle()
    batch = args.batch
    patches = make_patch_batch(ctx, batch, seed=ctx.rank)          # batch * 16 KiB >> 126 MB L2
    vmax = value_bound(ctx, ctx.frame_max)
    zp = ZPs(N_MAX, PATCH, precision=args.precision, value_max=vmax)
    prec = PREC_NAMES[zp._precision_code(device_stack=True)]
    n_modes = len(zp.n)
    gathered = torch.empty((world * batch, n_modes), dtype=torch.float32, device="cuda") if world > 1 else None
    push = world > 1 and args.gather == "push" and prec in ("tf32", "tf32x3", "f16x3")
    peers = PeerArray(world * batch, n_modes) if push else None
    hold = {}

    def compute():
        hold["z"] = zp.transform(patches).data

    def step_nccl():
        compute()
        gather_rows(hold["z"], out=gathered, sizes=[batch] * world)

    def step_push():
        peers.begin()                                   # nobody still reads the previous round's rows
        hold["z"] = zp.transform_allgather(patches, peers, ctx.rank * batch)
        peers.fence()                                   # every rank's rows have landed in every copy

    step = compute if world == 1 else (step_push if push else step_nccl)

    # parity of the timed batch itself: a 4096-patch sample against the oracle (float64 numpy.dot)
    compute()
    rows = sample_rows(batch)
    idx = torch.from_numpy(rows).cuda()
    ref = zo.project_patches(patches[idx].cpu().numpy().astype(np.float64), zp.polynomials)
    got = hold["z"][idx].double().cpu().numpy()
    if prec == "tf32":
        err = float(np.abs(got - ref).max() / np.abs(ref).max())
        assert err <= 1e-3, f"parity FAILED (tf32 stated bound 1e-3*max): {err:.2e}"
    else:
        err = fp32_close(got, ref, "patches (metric shape)")
    parity = {"sample": len(rows), "max_err_over_max": err,
              "gate": "abs err <= 1e-3*max|ref| (tf32 fast mode)" if prec == "tf32" else "allclose(rtol=1e-4, atol=1e-6*max|ref|)"}

    ms, clocks, launches = timed(ctx, step, args.steps, args.warmup)
    total = batch * world
    value = total * args.steps / (ms / 1e3)
    gather = None
    if world > 1:
        ms_ng, _, _ = timed(ctx, compute, args.steps, 3)
        ms_nccl, _, _ = timed(ctx, step_nccl, args.steps, 3)
        if push:                                       # the pushed copies equal the NCCL gather bit for bit, on every rank
            step_push()
            torch.cuda.synchronize()
            assert torch.equal(peers.local, gathered), "peer-pushed rows differ from the NCCL all-gather"
        gather = gather_report(ctx, "fused: P2P stores over NVLink from a warp of the projection kernel" if push
                               else "all_gather_into_tensor (NCCL) after the kernel", ms, ms_ng, ms_nccl if push else None,
                               args.steps, total, (world - 1) * batch * n_modes * 4)
        compute()
        assert torch.equal(gathered[ctx.rank * batch:(ctx.rank + 1) * batch], hold["z"]), "gathered rows differ from the local shard"
    else:
        ms_ng = ms
    alg_bytes = batch * (PATCH * PATCH * 4 + n_modes * 4)
    achieved = alg_bytes * args.steps / (ms_ng / 1e3) / 1e9
    flops = 2.0 * batch * PATCH * PATCH * n_modes
    folded = prec == "f16x3" and PATCH % 64 == 0
    kernel = {"fp32": "project_simt_kernel", "tf32": "project_tc_kernel<plain>", "tf32x3": "project_tc3_kernel<plain,pair>",
              "f16x3": ("project_fold_kernel<plain,0> (mirror-folded fp16 split" if folded else "project_tc3_kernel<plain,pair,f16> (fp16 split")
                       + (", value_max hint)" if vmax else ", auto-range: range_sample_kernel + conditional tf32x3 launch behind it)")}[prec]
    roof = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
            "traffic": NCU_TRAFFIC.get(prec) if batch == 262144 else None,
            "traffic_source": "profiles/ ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch",
            "peak_source": pk["source"] + " burst copy bandwidth", "kernel": kernel, "algorithmic_bytes_per_launch": alg_bytes,
            "tflops": flops * args.steps / (ms_ng / 1e3) / 1e12, "timed": "projection kernel only (the no-gather loop when N>1)"}

    # sustained leg: >= 2.5 s of back-to-back launches with the 2 ms clock sampler -- what a long job sees
    sus_steps = args.steps if args.no_sustained else max(args.steps, int(2500.0 / max(ms_ng / args.steps, 1e-3)))
    ms_s, clocks_s, _ = timed(ctx, compute, sus_steps, 3)
    roof["sustained"] = {"steps": sus_steps, "seconds": ms_s / 1e3, "ms_per_step": ms_s / sus_steps,
                         "value": total * sus_steps / (ms_s / 1e3), "achieved": alg_bytes * sus_steps / (ms_s / 1e3) / 1e9,
                         "frac": alg_bytes * sus_steps / (ms_s / 1e3) / 1e9 / pk["hbm_gbs"], "clocks": clocks_s}

    if prec == "f16x3" and vmax:                       # the same batch without the range hint, same run
        zp_nohint = ZPs(N_MAX, PATCH, precision=args.precision)
        ms_t, _, _ = timed(ctx, lambda: zp_nohint.transform(patches), args.steps, 3)
        p_nh = PREC_NAMES[zp_nohint._precision_code(device_stack=True)]
        roof["without_value_max"] = {"precision": p_nh,
                                     "kernel": "range_sample_kernel + project_fold_kernel + conditional project_tc3_kernel (auto-range)"
                                               if p_nh == "f16x3" else "project_tc3_kernel<plain,pair>",
                                     "ms_per_step": ms_t / args.steps, "value": total * args.steps / (ms_t / 1e3),
                                     "frac": alg_bytes * args.steps / (ms_t / 1e3) / 1e9 / pk["hbm_gbs"]}
        zp_x3 = ZPs(N_MAX, PATCH, precision="tf32x3")   # the unfolded range-free kernel (what the fallback runs)
        ms_x, _, _ = timed(ctx, lambda: zp_x3.transform(patches), args.steps, 3)
        roof["tf32x3_unfolded"] = {"kernel": "project_tc3_kernel<plain,pair>", "ms_per_step": ms_x / args.steps,
                                   "value": total * args.steps / (ms_x / 1e3),
                                   "frac": alg_bytes * args.steps / (ms_x / 1e3) / 1e9 / pk["hbm_gbs"]}
    if peers is not None:
        peers.close()
    e2e = e2e_patches(ctx, zp, n_modes)
    cfg = workload_config("patches", world)
    cfg["batch_per_gpu"] = batch
    return {"metric": "zernike_patches_per_sec", "value": value, "unit": "patches/s", "ms_per_step": ms / args.steps,
            "steps": args.steps, "dtype": DTYPES[prec], "precision": prec, "value_max": vmax, "scaling": "weak", "config": cfg, "roofline": roof, "gather": gather,
            "parity": parity, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}


def e2e_patches(ctx, zp, n_modes):
    """End to end through the public API with HOST buffers, copies inside the timed region.  Headline route = the
    reference's own pipeline: frame + peak coordinates in, float64 moments out (KeyPoints.extract_patches ->
    ZPs.transform; ZPs.transform_peaks_batch -> zb200_project_peaks_host); beside it the patch-stack route
    ZPs.transform(numpy) with pageable and with pinned input (16 KB per patch cross the bus: PCIe-bound)."""
    import numpy as np
    torch, args, world = ctx.torch, ctx.args, ctx.world
    from motif_learn_b200.datasets import honeycomb_image
    from motif_learn_b200.features import ZPs, clear_border
    zo = _oracle()
    zp_host = ZPs(N_MAX, PATCH, precision=args.precision, output="numpy")
    # --- frame route: 16 pinned 2048^2 frames per step (a short in-situ series), ~21 k peaks each
    n_frames = 16
    img, pts = honeycomb_image(1024, bond=12.0, seed=100 + ctx.rank)
    tile = np.tile(img, (2, 2))                                             # 2048^2 without the slow host renderer
    kept = np.concatenate([clear_border(pts + np.array([dx, dy]), tile.shape, PATCH) for dx in (0, 1024) for dy in (0, 1024)])
    pinned = torch.empty((n_frames, 2048, 2048), dtype=torch.float32, pin_memory=True)
    for f in range(n_frames):
        pinned[f].copy_(torch.from_numpy(np.roll(tile, 7 * f, axis=1)))     # distinct frames, same lattice statistics
    frames = [pinned[f].numpy() for f in range(n_frames)]
    # (rolled columns move the atoms too: roll the peak columns and re-apply the border rule)
    pts_list = []
    for f in range(n_frames):
        q = kept.copy()
        q[:, 0] = (q[:, 0] + 7 * f) % 2048
        pts_list.append(clear_border(q, tile.shape, PATCH))
    res = zp_host.transform_peaks_batch(frames[:4], pts_list[:4])           # warm-up: every staging slot, pool threads
    ref = zo.project_patches(zo.extract_patches(frames[1], pts_list[1][:512], PATCH).astype(np.float64), zp.polynomials)
    if args.precision == "tf32":                                            # fast mode: stated bound 1e-3 * max
        assert np.abs(res[1].data[:512] - ref).max() <= 1e-3 * np.abs(ref).max(), "parity FAILED for the e2e frame route (tf32)"
    else:
        fp32_close(res[1].data[:512], ref, "e2e frame route")
    steps = max(2, min(args.steps, 5))
    n_patches = sum(len(q) for q in pts_list)
    result = np.zeros((n_patches, n_modes))                                 # a frame loop keeps ONE result buffer (out=)
    zp_host.transform_peaks_batch(frames, pts_list, out=result)             # full warm-up: the buffer's pages exist
    # the oracle check above ran numpy.dot: OpenBLAS' worker threads keep spinning for ~0.1 s after a call and steal
    # the cores of the library's own host threads (measured: 8 ms -> 18 ms per call) -- let the CHECKER go idle first
    time.sleep(1.0)
    rates = {}
    for label, kw in (("reused_out", {"out": result}), ("fresh_out", {})):
        if world > 1:
            ctx.dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            res = zp_host.transform_peaks_batch(frames, pts_list, **kw)
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        assert sum(r.data.shape[0] for r in res) == n_patches and res[0].data.dtype == np.float64
        rates[label] = ctx.sum_over_ranks(n_patches) * steps / dt
    out = {"value": rates["reused_out"], "unit": "patches/s",
           "h2d_bytes_per_step": n_frames * 2048 * 2048 * 4 + n_patches * 16, "d2h_bytes_per_step": n_patches * n_modes * 4,
           "steps": steps, "patches_per_step": n_patches,
           "api": "ZPs.transform_peaks_batch(16 pinned numpy frames, peak lists, out=buffer) -> zb200_project_peaks_host -> float64 numpy",
           "fresh_result_array_per_call": {"value": rates["fresh_out"], "unit": "patches/s",
                                           "what": "same call without out=: a new 240 MB float64 array per call pays its page faults"}}
    # --- patch-stack route, pageable (what a reference user holds) and pinned
    n_e2e = min(args.batch, args.e2e_batch)
    base = zo.extract_patches(frames[0], pts_list[0][:8192], PATCH)
    pageable = np.concatenate([base] * (-(-n_e2e // len(base))))[:n_e2e]
    pin = torch.empty((n_e2e, PATCH, PATCH), dtype=torch.float32, pin_memory=True)
    pin.copy_(torch.from_numpy(pageable))
    zp_host.transform(pageable[:1024])
    time.sleep(0.5)
    for label, arr in (("pageable", pageable), ("pinned", pin.numpy())):
        if world > 1:
            ctx.dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            r = zp_host.transform(arr)
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        assert r.data.shape == (n_e2e, n_modes)
        out[f"patch_stack_{label}"] = {"value": n_e2e * world * steps / dt, "unit": "patches/s",
                                        "h2d_bytes_per_step": n_e2e * PATCH * PATCH * 4, "d2h_bytes_per_step": n_e2e * n_modes * 4,
                                        "api": f"ZPs.transform({label} numpy stack) -> zb200_project_patches_host"}
    return out


# --------------------------------------------------------------------------- dense symmetry map (C2 and C4)
def parity_map_sample(zp, dimg, scores, row0, rows, count=384):
    """Scores of `count` random pixels of the band against the oracle: window gather (zero extension) + float64
    numpy.dot + rot_maps (the map equals the projection of the window centred at the pixel)."""
    import numpy as np
    zo = _oracle()
    k = zp.size
    h, w = int(dimg.shape[0]), int(dimg.shape[1])
    rng = np.random.default_rng(7)
    ys = np.concatenate([rng.integers(row0, row0 + rows, count - 8), [row0] * 4, [row0 + rows - 1] * 4])
    xs = np.concatenate([rng.integers(0, w, count - 8), [0, 1, w // 2, w - 1] * 2])
    pad = dimg.new_zeros((h + 2 * k, w + 2 * k))
    pad[k:k + h, k:k + w] = dimg
    wins = np.stack([pad[y + k - k // 2: y + k - k // 2 + k, x + k - k // 2: x + k - k // 2 + k].cpu().numpy() for y, x in zip(ys, xs)])
    z = zo.project_patches(wins.astype(np.float64), zp.polynomials)
    ref = zo.rot_maps(z, zp.n, zp.m, FOLDS)
    tix = lambda a: scores.new_tensor(a, dtype=scores.dtype).long()      # noqa: E731
    got = scores[:, tix(ys - row0), tix(xs)].double().cpu().numpy().T
    ok = np.isfinite(ref)
    assert np.array_equal(ok, np.isfinite(got)), "parity FAILED: NaN pattern of the symmetry map differs from the oracle"
    err = float(np.abs(got[ok] - ref[ok]).max())
    return err, count


def bench_map(ctx, tiled=False):
    """tiled=False: BASELINE configs[1], one 2048^2 frame per GPU (weak scaling over frames).
    tiled=True: BASELINE configs[3], ONE 4096^2 frame with defects, 64-px window, output row bands spread over
    the GPUs (strong scaling; halo rows come from the replicated frame), score bands gathered to every rank."""
    import numpy as np
    torch, args, pk, world, rank = ctx.torch, ctx.args, ctx.pk, ctx.world, ctx.rank
    from motif_learn_b200.datasets import honeycomb_frame_gpu
    from motif_learn_b200.features import ZPs
    from motif_learn_b200.parallel import PeerArray, gather_rows, row_band, shard_sizes, symmetry_map_allgather
    name = "map4k" if tiled else "map"
    size, window = (4096, 64) if tiled else (2048, 48)
    zp = ZPs(N_MAX, window, precision=args.precision)
    prec = PREC_NAMES[zp._precision_code(for_map=True)]
    if tiled:
        dimg, _ = honeycomb_frame_gpu(size, bond=12.0, seed=0, vacancy_frac=0.01, dopant_frac=0.005)
        row0, rows = row_band(size, rank, world)
    else:
        dimg, _ = honeycomb_frame_gpu(size, bond=12.0, seed=rank)
        row0, rows = 0, size
    scratch = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    hold = {}
    do_gather = tiled and world > 1
    # the band is (F, rows, W); gathering along rows = all-gather of (rows, F, W) blocks in rank order
    full = torch.empty((size, len(FOLDS), size), dtype=torch.float32, device="cuda") if do_gather else None

    push = do_gather and args.gather == "push"
    peers = PeerArray(len(FOLDS) * size, size) if push else None       # every rank's copy of the (F, H, W) score map

    def compute():
        hold["s"] = zp.symmetry_map(dimg, FOLDS, row0=row0, rows=rows)

    def step_nccl():
        compute()
        gather_rows(hold["s"].permute(1, 0, 2), out=full, sizes=shard_sizes(size, world))

    def step_push():
        # four sub-bands: the copy engines forward sub-band i (one 2-D copy per peer) while i+1 is being computed
        peers.begin()
        hold["parts"] = symmetry_map_allgather(zp, dimg, FOLDS, peers, row0, rows, size, n_sub=4)
        peers.fence()

    step = step_push if push else (step_nccl if do_gather else compute)

    compute()
    err, n_px = parity_map_sample(zp, dimg, hold["s"], row0, rows)
    bound = 1e-3 if prec in ("tf32", "f16") else 1e-5
    assert err <= bound, f"parity FAILED for {name}: score error {err:.2e} > {bound}"
    parity = {"sample": n_px, "max_abs_err": err, "gate": f"n-fold scores abs <= {bound}, NaN pattern equal"}

    steps = max(2, min(args.steps, args.map_steps))
    flush = lambda: scratch.zero_()                                      # noqa: E731  (> L2 capacity)
    ms, clocks, launches = timed(ctx, step, steps, max(3, min(args.warmup, 5)), between=flush)
    mpix = size * size / 1e6
    units = mpix * (1 if tiled else world)
    value = units * steps / (ms / 1e3)
    gather = None
    ms_ng = ms
    if do_gather:
        ms_ng, _, _ = timed(ctx, compute, steps, 3, between=flush)
        ms_nccl, _, _ = timed(ctx, step_nccl, steps, 3, between=flush)
        if push:
            step_push()
            torch.cuda.synchronize()
            assert torch.equal(peers.local.view(len(FOLDS), size, size), full.permute(1, 0, 2)), "pushed score map differs from the NCCL gather"
        gather = gather_report(ctx, "4 sub-bands per rank, each forwarded to every peer by the copy engines (cudaMemcpy2DAsync over NVLink) while the next is computed" if push
                               else "all_gather_into_tensor (NCCL) after the kernel", ms, ms_ng, ms_nccl if push else None,
                               steps, units, (size - rows) * len(FOLDS) * size * 4)
        assert torch.equal(full[row0:row0 + rows].permute(1, 0, 2), hold["s"]), "gathered band differs from the local one"
    if peers is not None:
        peers.close()
    flops = 2.0 * rows * size * window * window * len(zp.n)
    ach = flops * steps / (ms_ng / 1e3) / 1e12
    # denominator: the measured dense bf16 rate for the kind::f16 kernels (same tensor-pipe rate), the tf32 rate for
    # the kind::tf32 kernels.  `achieved` counts ALGORITHMIC flops only; `executed_tflops` adds what the kernel
    # really issues (three split terms, modes padded to 96, window rows padded to 16-tap groups).
    f16_kernel = prec in ("f16", "f16x3")
    peak = pk["bf16_tflops"] if f16_kernel else pk["tf32_tflops"]
    pad_modes = 96.0 / len(zp.n) if prec != "fp32" else 1.0
    pad_taps = 1.0
    if f16_kernel:
        xs = -1.0 + 2.0 * np.arange(window) / (window - 1)
        rows_on = int(((xs[None, :] ** 2 + xs[:, None] ** 2) <= 1.0 + 1e-9).any(axis=1).sum())
        pad_taps = rows_on * (-(-window // 16)) * 16.0 / (window * window)
    executed = ach * {"fp32": 1, "tf32": 1, "tf32x3": 3, "f16": 1, "f16x3": 3}[prec] * pad_modes * pad_taps
    roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "traffic": NCU_TRAFFIC_MAP.get(prec) if not tiled else None,
            "peak_source": pk["source"] + (" bf16 dense (kind::f16 MMAs)" if f16_kernel else " tf32: " + pk["tf32_source"]),
            "kernel": {"fp32": "map_simt_kernel<scores>", "tf32": "map_tc_kernel<scores>", "tf32x3": "map_tc_kernel<scores>",
                       "f16": "map_h_kernel<scores,x1>", "f16x3": "map_h_kernel<scores,x3>"}[prec],
            "algorithmic_flops_per_launch": flops, "executed_tflops": executed, "executed_frac": executed / peak,
            "note": "fp32-grade accuracy costs 3 split MMAs per tap on fp16/tf32 tensor cores: executed_frac is the pipe's "
                    "utilisation, frac the algorithmic one (DESIGN.md section 4)"}
    # end to end: numpy frame in, numpy scores out (float64, like the reference); over the bus: fp32 frame up, fp32 scores down
    zp_host = ZPs(N_MAX, window, precision=args.precision, output="numpy")
    host = torch.empty((size, size), dtype=torch.float32, pin_memory=True)
    host.copy_(dimg)
    zp_host.symmetry_map(host.numpy(), FOLDS, row0=row0, rows=rows)      # warm-up: pinned staging, mempool growth
    torch.cuda.synchronize()
    if world > 1:
        ctx.dist.barrier()
    t0 = time.perf_counter()
    e2e_steps = 3
    for _ in range(e2e_steps):
        res = zp_host.symmetry_map(host.numpy(), FOLDS, row0=row0, rows=rows)
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": units * e2e_steps / dt, "unit": "Mpix/s", "h2d_bytes_per_step": size * size * 4,
           "d2h_bytes_per_step": int(res.size * 4), "steps": e2e_steps, "api": "ZPs.symmetry_map(numpy) -> float64 numpy"}
    cfg = workload_config(name, world)
    return {"metric": "symmetry_map_mpix_per_sec", "value": value, "unit": "Mpix/s", "ms_per_step": ms / steps,
            "steps": steps, "dtype": DTYPES[prec], "precision": prec, "scaling": "strong" if tiled else "weak", "config": cfg, "roofline": roof,
            "gather": gather, "parity": parity, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}


# --------------------------------------------------------------------------- C3: complex n_max=20 on 1 Mi patches
def bench_c3(ctx):
    import warnings
    import numpy as np
    torch, args, pk, world, rank = ctx.torch, ctx.args, ctx.pk, ctx.world, ctx.rank
    from motif_learn_b200.features import ZPs
    from motif_learn_b200.parallel import PeerArray, gather_rows, shard_range, shard_sizes
    zo = _oracle()
    total = args.c3_total
    lo, hi = shard_range(total, rank, world)
    n = hi - lo
    n_c = 121
    patches = make_patch_batch(ctx, n, seed=1000 + rank)                 # 16 KiB per patch: 16.4 GiB at N=1
    vmax = value_bound(ctx, ctx.frame_max)
    zp = ZPs(C3_NMAX, PATCH, precision=args.precision, value_max=vmax)
    prec = PREC_NAMES[zp._precision_code(device_stack=True)]
    gathered = torch.empty((total, n_c), dtype=torch.complex64, device="cuda") if world > 1 else None
    hold = {}

    push = world > 1 and args.gather == "push" and prec in ("tf32", "tf32x3", "f16x3")
    peers = PeerArray(total, 2 * n_c) if push else None

    def compute(kind="complex"):
        hold["z"] = zp.transform_features(patches, kind)

    def step_nccl():
        compute()
        gather_rows(torch.view_as_real(hold["z"]).reshape(n, 2 * n_c), out=torch.view_as_real(gathered).reshape(total, 2 * n_c),
                    sizes=shard_sizes(total, world))

    def step_push():
        peers.begin()
        hold["z"] = torch.view_as_complex(zp.transform_allgather(patches, peers, lo, "complex").view(n, n_c, 2))
        peers.fence()

    step = compute if world == 1 else (step_push if push else step_nccl)

    compute()
    rows = sample_rows(n)
    idx = torch.from_numpy(rows).cuda()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        nn, mm, basis = zo.zernike_basis(C3_NMAX, PATCH)
    refz = zo.to_complex(zo.project_patches(patches[idx].cpu().numpy().astype(np.float64), basis), nn, mm)[0]
    got = hold["z"][idx].cpu().numpy().astype(np.complex128)
    if prec == "tf32":
        err = float(np.abs(got - refz).max() / np.abs(refz).max())
        assert err <= 1e-3, f"parity FAILED (tf32 stated bound): {err:.2e}"
    else:
        # one array [Re | Im]: atol is 1e-6 * max over the whole moment array, as for the real representation
        err = fp32_close(np.concatenate([got.real, got.imag], axis=1), np.concatenate([refz.real, refz.imag], axis=1), "c3 complex moments")
    parity = {"sample": len(rows), "max_err_over_max": err, "gate": "allclose(rtol=1e-4, atol=1e-6*max|ref|) on [Re | Im]"}

    steps = max(2, min(args.steps, 10))
    ms, clocks, launches = timed(ctx, step, steps, 3)
    value = total * steps / (ms / 1e3)
    gather = None
    ms_ng = ms
    if world > 1:
        ms_ng, _, _ = timed(ctx, compute, steps, 3)
        ms_nccl, _, _ = timed(ctx, step_nccl, steps, 3)
        if push:
            step_push()
            torch.cuda.synchronize()
            assert torch.equal(peers.local, torch.view_as_real(gathered).reshape(total, 2 * n_c)), "pushed rows differ from the NCCL gather"
        gather = gather_report(ctx, "fused: P2P stores over NVLink from a warp of the projection kernel" if push
                               else "all_gather_into_tensor (NCCL) after the kernel", ms, ms_ng, ms_nccl if push else None,
                               steps, total, (total - n) * n_c * 8)
        compute()
        assert torch.equal(gathered[lo:hi], hold["z"]), "gathered rows differ from the local shard"
    if peers is not None:
        peers.close()
    ms_abs, _, _ = timed(ctx, lambda: compute("abs"), steps, 3)
    alg_bytes = n * (PATCH * PATCH * 4 + n_c * 8)
    flops = 2.0 * n * PATCH * PATCH * 231
    gbs = alg_bytes * steps / (ms_ng / 1e3) / 1e9
    tfl = flops * steps / (ms_ng / 1e3) / 1e12
    f_h, f_t = gbs / pk["hbm_gbs"], tfl / pk["tf32_tflops"]
    roof = {"bound": "hbm" if f_h >= f_t else "tensor", "achieved": gbs if f_h >= f_t else tfl,
            "peak": pk["hbm_gbs"] if f_h >= f_t else pk["tf32_tflops"], "unit": "GB/s" if f_h >= f_t else "TFLOP/s",
            "frac": max(f_h, f_t), "hbm": {"achieved": gbs, "peak": pk["hbm_gbs"], "frac": f_h},
            "tensor": {"achieved": tfl, "peak": pk["tf32_tflops"], "frac": f_t, "peak_source": pk["tf32_source"]},
            "traffic": NCU_TRAFFIC_C3_262144 * (n / 262144.0) if prec == "f16x3" and PATCH % 64 == 0 else None,
            "traffic_source": "profiles/r02_prof_c3_raw.csv: one ncu --set full capture at 262 144 patches, scaled by the launch's patch count",
            "kernel": ("project_fold_kernel<plain,1> (mirror-folded fp16 split, class widths 80/64/64/64)" if prec == "f16x3" and PATCH % 64 == 0
                       else "project_tc3_kernel<plain> on the complex-interleaved operand" if prec in ("tf32x3", "f16x3") else "project_tc_kernel"),
            "executed_flops_note": "the folded kernel executes 2*N*(k^2/4)*272*3 fp16 flops; frac is on ALGORITHMIC flops 2*N*k^2*231",
            "algorithmic_bytes_per_launch": alg_bytes, "algorithmic_flops_per_launch": flops,
            "note": "SURVEY.md 8d: report both fractions, the larger one binds"}
    cfg = workload_config("c3", world)
    return {"metric": "zernike_patches_per_sec", "value": value, "unit": "patches/s", "ms_per_step": ms / steps, "steps": steps,
            "dtype": DTYPES[prec], "precision": prec, "value_max": vmax, "scaling": "strong", "config": cfg,
            "details": {"total_patches": total, "patches_per_gpu": n}, "roofline": roof, "gather": gather, "parity": parity,
            "abs_features": {"ms_per_step": ms_abs / steps, "value": total * steps / (ms_abs / 1e3), "unit": "patches/s",
                             "what": "|Zc| epilogue instead of complex (no gather)"},
            "e2e": None, "gpu_launches": launches, "clocks": clocks}


# --------------------------------------------------------------------------- C5: frame series -> peaks -> gather -> |Zc|
def bench_c5(ctx):
    import numpy as np
    torch, args, pk, world, rank = ctx.torch, ctx.args, ctx.pk, ctx.world, ctx.rank
    from motif_learn_b200.datasets import honeycomb_frame_gpu
    from motif_learn_b200.features import ZPs, clear_border, local_max, series_features
    from motif_learn_b200.parallel import gather_rows, shard_range
    zo = _oracle()
    n_frames_total = args.c5_frames
    lo, hi = shard_range(n_frames_total, rank, world)
    zp = ZPs(N_MAX, PATCH, precision=args.precision)
    prec = PREC_NAMES[zp._precision_code(device_stack=True)]     # the projection inside series_features runs on CUDA stacks
    n_c = 49
    rng = np.random.default_rng(12345)
    angles = rng.uniform(0.0, 60.0, n_frames_total)
    frames = torch.empty((hi - lo, C5_SIZE, C5_SIZE), dtype=torch.float32, device="cuda")
    for f in range(lo, hi):                                             # in-situ series: seeds 0..255, random angle, jitter 0.3 px
        img, _ = honeycomb_frame_gpu(C5_SIZE, bond=12.0, seed=f, angle=float(angles[f]), jitter=0.3, noise=0.01)
        frames[f - lo] = img
    hold = {}

    def compute():
        # per frame: local_max -> clear_border -> gather kernel -> projection with the |Zc| epilogue; the frames are
        # spread over `workers` host threads with one CUDA stream each so their small kernels interleave
        feats, pts = series_features(zp, frames, PEAK_MIN_DISTANCE, PEAK_THRESHOLD, kind="abs", workers=args.c5_workers,
                                     fused={"auto": None, "off": False}[args.c5_fused])
        hold["feats"], hold["counts"] = feats, [len(q) for q in pts]

    def step():
        compute()
        if world > 1:
            hold["all"] = gather_rows(torch.cat(hold["feats"]))                      # ragged: sizes exchanged first

    compute()
    # parity: frame 0 of this rank -- peaks vs the oracle's local_max, features of 512 of them vs numpy
    f0 = frames[0].cpu().numpy()
    pk_ref = zo.clear_border(zo.local_max(f0, PEAK_MIN_DISTANCE, PEAK_THRESHOLD), f0.shape, PATCH)
    pk_got = clear_border(local_max(frames[0], PEAK_MIN_DISTANCE, PEAK_THRESHOLD), f0.shape, PATCH)
    assert np.array_equal(pk_ref, pk_got), "parity FAILED: peaks of frame 0 differ from the oracle"
    sub = pk_ref[:: max(1, len(pk_ref) // 512)][:512]
    ref = np.abs(zo.to_complex(zo.project_patches(zo.extract_patches(f0, sub, PATCH).astype(np.float64), zp.polynomials), zp.n, zp.m)[0])
    got = zp.transform_peaks(frames[0], sub, "abs").double().cpu().numpy()
    err = float(np.abs(got - ref).max() / ref.max())
    assert err <= 3e-6, f"parity FAILED for c5 |Zc| features: {err:.2e} * max"
    parity = {"peaks_frame0": int(len(pk_ref)), "peaks_equal_oracle": True, "feature_sample": int(len(sub)),
              "max_err_over_max": err, "gate": "identical peak list; |Zc| abs err <= 3e-6*max|ref|"}

    steps = max(2, min(args.steps, 3))
    ms, clocks, launches = timed(ctx, step, steps, 3)
    n_patches = ctx.sum_over_ranks(float(sum(hold["counts"])))
    value = n_patches * steps / (ms / 1e3)
    gather = None
    if world > 1:
        ms_ng, _, _ = timed(ctx, compute, steps, 3)
        mine = float(sum(hold["counts"]))
        gather = gather_report(ctx, "ragged all_gather_into_tensor (NCCL) of the per-rank feature blocks, sizes exchanged first",
                               ms, ms_ng, None, steps, n_patches, int((n_patches - mine) * n_c * 4))
        assert hold["all"].shape[0] == int(n_patches)
    cfg = workload_config("c5", world)
    details = {"frames_total": n_frames_total, "frames_per_gpu": hi - lo, "patches_total": int(n_patches),
               "peaks": f"local_max(min_distance={PEAK_MIN_DISTANCE}, threshold={PEAK_THRESHOLD}) on the GPU",
               "pipeline": f"features.series_features, {args.c5_workers} worker threads / streams"}
    per_frame_bytes = C5_SIZE * C5_SIZE * 4
    return {"metric": "zernike_patches_per_sec", "value": value, "unit": "patches/s", "ms_per_step": ms / steps, "steps": steps,
            "frames_per_sec": n_frames_total * steps / (ms / 1e3), "dtype": DTYPES[prec], "precision": prec, "scaling": "strong",
            "config": cfg, "details": details,
            "roofline": {"bound": "latency", "note": "per frame: peak detection (sort + suppression with host round trips), "
                         "gather kernel, projection kernel; ~1 ms per frame, none of the three near its roofline at 21 k patches",
                         "frame_bytes_per_sec": n_frames_total * per_frame_bytes * steps / (ms / 1e3)},
            "gather": gather, "parity": parity, "e2e": None, "gpu_launches": launches, "clocks": clocks}


RUNNERS = {"patches": bench_patches, "map": lambda c: bench_map(c, tiled=False), "map4k": lambda c: bench_map(c, tiled=True),
           "c3": bench_c3, "c5": bench_c5}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="patches", choices=WORKLOADS)
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "tf32", "tf32x3", "f16", "f16x3"])
    ap.add_argument("--batch", type=int, default=262144, help="patches per GPU per step (metric shape)")
    ap.add_argument("--e2e-batch", type=int, default=32768)
    ap.add_argument("--map-steps", type=int, default=20)
    ap.add_argument("--c3-total", type=int, default=C3_TOTAL)
    ap.add_argument("--c5-frames", type=int, default=C5_FRAMES)
    ap.add_argument("--c5-workers", type=int, default=6, help="host threads / CUDA streams of the frame-series pipeline")
    ap.add_argument("--c5-fused", default="auto", choices=["auto", "off"], help="off: gather kernel + projection instead of the gathering folded kernel")
    ap.add_argument("--also", default="all", help="comma list of the other workloads to carry under 'also' (all | none | names)")
    ap.add_argument("--no-also", action="store_true", help="same as --also none")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--value-max", default="frame",
                    help="range hint for the fp16-split projection: 'frame' (max of the frame the patches come from), a number, or none")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 2.5 s sustained leg of the patch workload")
    ap.add_argument("--gather", default="push", choices=["push", "nccl"],
                    help="N>1: how the timed step gathers the features (push = fused P2P stores / copy engines; nccl = all_gather after the kernel)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)                                   # timing rule: at least 3 warm-up steps

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the communicator is created; stdout must carry
        # exactly one JSON line, so route fd 1 to stderr until the first collective has run
        sys.stdout.flush()
        keep = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(keep, 1)
            os.close(keep)
    ctx = Ctx(torch, dist, rank, world, local, args, peaks())
    line = RUNNERS[args.workload](ctx)
    line.update({"n_gpus": world, "warmup": args.warmup, "higher_is_better": True, "vs_baseline": None, "data": "synthetic"})
    others = [] if (args.no_also or args.also == "none") else (
        [w for w in WORKLOADS if w != args.workload] if args.also == "all" else [w for w in args.also.split(",") if w in WORKLOADS])
    also = []
    for name in others:
        torch.cuda.empty_cache()
        try:
            other = RUNNERS[name](ctx)
            other["workload"] = name
            also.append(other)
        except Exception as exc:  # a secondary number must never sink the primary line (all ranks fail alike: same code path)
            also.append({"workload": name, "error": f"{type(exc).__name__}: {exc}"})
    if also:
        line["also"] = also
    if rank == 0 and not args.no_cpu:
        cores = cpu_threads()
        line["cpu_baseline"] = cpu_baseline_for(args.workload, cores)
        for other in also:
            if "error" not in other:
                try:
                    other["cpu_baseline"] = cpu_baseline_for(other["workload"], cores)
                except Exception as exc:
                    other["cpu_baseline"] = {"error": f"{type(exc).__name__}: {exc}"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
