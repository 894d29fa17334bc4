#!/usr/bin/env python
"""bench.py -- throughput of the B200 Zernike hot path (one JSON line on stdout, rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload patches|map] [--impl reference]

Workloads (BASELINE.json metric: "Zernike patches/sec (n_max=12, 64px) and symmetry-map
Mpix/sec"):
  patches  one step = ZPs(n_max=12, size=64) moments of a batch of synthetic 64x64 lattice
           patches resident in HBM (batch >> L2, so every step streams from HBM)  [default]
  map      one step = fused symmetry map (n-folds 2,3,4,6) of a 2048x2048 synthetic lattice
           frame, n_max=12, 48-px window (BASELINE configs[1]); L2 is flushed between steps
The default line carries the other workload under "also" so both headline numbers appear.
Multi-GPU (torchrun): every rank runs the same step on its own shard (weak scaling), timed
as max over ranks between barriers.
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
MAP_SIZE = 2048
MAP_WINDOW = 48
FOLDS = [2, 3, 4, 6]
# DRAM bytes per launch of the projection kernel at the default batch, from the committed ncu
# captures (profiles/r01_prof_tc3_raw.csv, r01_prof_tc1_raw.csv): read + write
NCU_TRAFFIC = {"tf32x3": 4.198750e9 + 43.819e6, "tf32": 4.181859e9 + 53.549e6}
# dense-map kernels at 2048^2 (profiles/r01_prof_map_raw.csv, r01_prof_maph_raw.csv): read + write
NCU_TRAFFIC_MAP = {"tf32x3": 136.456e6 + 48.684e6, "f16x3": 136.000e6 + 47.324e6}
PREC_NAMES = {0: "fp32", 1: "tf32", 2: "tf32x3", 3: "f16", 4: "f16x3"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
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
            while not self._stop.is_set():
                self.samples.append(int(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                if get_reasons is not None:
                    mask = int(get_reasons(h))
                    for bit, name in names.items():
                        if mask & bit:
                            self.reasons.add(name)
                time.sleep(self.period)
        except Exception as exc:  # NVML missing: report that rather than fail the bench
            self.reasons.add(f"nvml_unavailable:{type(exc).__name__}")

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thread.join(timeout=2)

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------- reference arm
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


def cpu_patches(n_sample: int, repeats: int):
    """The reference's CPU algorithm for the patch path (numpy.dot in float64, _zps.py:146-157),
    through the oracle port, on n_sample patches of the bench batch.  Returns patches/s (best)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import zernike_oracle as zo
    from motif_learn_b200.datasets import honeycomb_image
    img, pts = honeycomb_image(1024, bond=12.0, seed=0)
    kept = zo.clear_border(pts, img.shape, PATCH)
    base = zo.extract_patches(img, kept, PATCH)
    reps = -(-n_sample // len(base))
    sample = np.concatenate([base] * reps)[:n_sample]
    _, _, basis = zo.zernike_basis(N_MAX, PATCH)
    zo.project_patches(sample[:256], basis)
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        zo.project_patches(sample, basis)
        times.append(time.perf_counter() - t0)
    return n_sample / min(times), times


def cpu_map(size: int, repeats: int = 1, window: int = MAP_WINDOW):
    """The reference's CPU algorithm for the map path (scipy fftconvolve per mode + rot_maps,
    _zps.py:159-193, _zmoments.py:420-462) through the oracle port on a size x size crop."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import zernike_oracle as zo
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


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = cpu_threads()
    if args.workload == "patches":
        n_sample = 20000
        per_step = []
        for _ in range(args.warmup):
            cpu_patches(n_sample, 1)
        for _ in range(args.steps):
            v, t = cpu_patches(n_sample, 1)
            per_step.append(t[0])
        total = sum(per_step)
        value = n_sample * args.steps / total
        line = {"metric": "zernike_patches_per_sec", "unit": "patches/s",
                "config": {"workload": f"patches n_max={N_MAX} size={PATCH} (metric shape)", "modes": 91,
                           "batch_per_step": n_sample, "precision": "f64 (numpy.dot)",
                           "parallelism": "host cores of rank 0 (bounded sample of the same workload)"},
                "sample": f"{n_sample} float32 64x64 lattice patches per step, numpy.dot float64 (reference algorithm)"}
    else:
        size = 512
        window = 64 if args.workload == "map4k" else MAP_WINDOW
        per_step = []
        for _ in range(min(args.warmup, 1)):
            cpu_map(256, window=window)
        for _ in range(args.steps):
            _, dt = cpu_map(size, window=window)
            per_step.append(dt)
        total = sum(per_step)
        value = size * size * args.steps / 1e6 / total
        line = {"metric": "symmetry_map_mpix_per_sec", "unit": "Mpix/s",
                "config": {"workload": f"symmetry map n_max={N_MAX} window={window} folds={FOLDS}",
                           "image": f"{size}x{size} crop of the {MAP_SIZE}x{MAP_SIZE} frame"},
                "sample": f"{size}x{size} crop per step: 91 scipy.fftconvolve (1 thread) + rot_maps"}
    line.update({"impl": "reference", "value": value, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                 "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
                 "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                 "cpu_baseline": {"value": value, "unit": line["unit"], "cores": cores, "kind": "port",
                                  "sample": line.pop("sample")},
                 "e2e": {"value": value, "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm
def flush_l2(torch, scratch):
    scratch.zero_()            # > L2 capacity: evicts the previous step's lines


def timed(torch, dist, world, fn, steps, warmup, device_index, between=None):
    """W warm-up steps, then exactly K timed steps with CUDA events on the launching stream,
    barrier + synchronize on both sides, max over ranks.  Returns (ms_total, clocks, launches)."""
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
    with ClockSampler(device_index) as clk:
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
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, clk.summary(), launches


def bench_patches(torch, dist, rank, world, args, pk):
    import numpy as np
    from motif_learn_b200 import _lib
    from motif_learn_b200.datasets import honeycomb_image
    from motif_learn_b200.features import ZPs, KeyPoints
    dev = torch.cuda.current_device()
    zp = ZPs(N_MAX, PATCH, precision=args.precision)
    prec = PREC_NAMES[zp._precision_code()]
    n_modes = len(zp.n)

    # synthetic input: patches gathered at the atom sites of lattice frames (K2), tiled to the batch
    img, pts = honeycomb_image(2048, bond=12.0, seed=rank)
    kp = KeyPoints(pts, torch.from_numpy(img).cuda(), PATCH)
    base = kp.extract_patches()                                   # ~21 k patches, device-resident
    batch = args.batch
    reps = -(-batch // base.shape[0])
    patches = base.repeat(reps, 1, 1)[:batch].contiguous()        # batch*16 KiB >> 126 MB L2
    del base
    out_holder = {}

    def step():
        out_holder["z"] = zp.transform(patches).data

    ms, clocks, launches = timed(torch, dist, world, step, args.steps, args.warmup, dev)
    value = batch * world * args.steps / (ms / 1e3)
    alg_bytes = batch * (PATCH * PATCH * 4 + n_modes * 4)
    achieved = alg_bytes * args.steps / (ms / 1e3) / 1e9
    flops = 2.0 * batch * PATCH * PATCH * n_modes
    roof = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / pk["hbm_gbs"], "traffic": NCU_TRAFFIC.get(prec) if batch == 262144 else None,
            "traffic_source": "profiles/r01 ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch",
            "peak_source": pk["source"],
            "kernel": {"fp32": "project_simt_kernel", "tf32": "project_tc_kernel<plain>",
                       "tf32x3": "project_tc3_kernel<plain,pair>"}[prec],
            "algorithmic_bytes_per_launch": alg_bytes,
            "tflops": flops * args.steps / (ms / 1e3) / 1e12}

    # end to end: the numpy user's call, pinned host buffers, H2D + D2H inside the timed region
    n_e2e = min(batch, args.e2e_batch)
    host = torch.empty((n_e2e, PATCH, PATCH), dtype=torch.float32, pin_memory=True)
    host.copy_(patches[:n_e2e])
    host_np = host.numpy()
    zp_host = ZPs(N_MAX, PATCH, precision=args.precision, output="numpy")
    zp_host.transform(host_np[:1024])
    e2e_steps = max(2, min(args.steps, 5))
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = zp_host.transform(host_np)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    assert res.data.shape == (n_e2e, n_modes)
    e2e = {"value": n_e2e * world * e2e_steps / dt, "unit": "patches/s",
           "h2d_bytes_per_step": n_e2e * PATCH * PATCH * 4, "d2h_bytes_per_step": n_e2e * n_modes * 4,
           "steps": e2e_steps, "api": "ZPs.transform(numpy pinned) -> zb200_project_patches_host"}
    line = {"metric": "zernike_patches_per_sec", "value": value, "unit": "patches/s",
            "ms_per_step": ms / args.steps, "dtype": {"fp32": "f32", "tf32": "tf32", "tf32x3": "tf32x3(f32-grade)"}[prec],
            "config": {"workload": f"patches n_max={N_MAX} size={PATCH} (metric shape)", "batch_per_gpu": batch,
                       "modes": n_modes, "precision": prec, "l2": "input batch larger than L2 (no flush needed)",
                       "parallelism": f"patch shards x{world}, no collective"},
            "roofline": roof, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
    return line


def bench_map(torch, dist, rank, world, args, pk, tiled=False):
    """tiled=False: BASELINE configs[1], one 2048^2 frame per GPU (weak scaling over frames).
    tiled=True: BASELINE configs[3], ONE 4096^2 frame with defects, 64-px window, output row bands
    spread over the GPUs (strong scaling; halo rows come from the replicated frame, no exchange)."""
    import numpy as np
    from motif_learn_b200.datasets import honeycomb_image
    from motif_learn_b200.features import ZPs
    from motif_learn_b200.parallel import row_band
    MAP_SIZE, MAP_WINDOW = (4096, 64) if tiled else (2048, 48)
    dev = torch.cuda.current_device()
    zp = ZPs(N_MAX, MAP_WINDOW, precision=args.precision)
    prec = PREC_NAMES[zp._precision_code(for_map=True)]
    if tiled:
        img, _ = honeycomb_image(MAP_SIZE, bond=12.0, seed=0, vacancy_frac=0.01, dopant_frac=0.005)
        row0, rows = row_band(MAP_SIZE, rank, world)
    else:
        img, _ = honeycomb_image(MAP_SIZE, bond=12.0, seed=rank)
        row0, rows = 0, MAP_SIZE
    dimg = torch.from_numpy(img).cuda()
    scratch = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    holder = {}

    def step():
        holder["s"] = zp.symmetry_map(dimg, FOLDS, row0=row0, rows=rows)

    steps = max(2, min(args.steps, args.map_steps))
    ms, clocks, launches = timed(torch, dist, world, step, steps, min(args.warmup, 3) if args.warmup >= 3 else 3, dev,
                                 between=lambda: flush_l2(torch, scratch))
    mpix = MAP_SIZE * MAP_SIZE / 1e6
    value = mpix * (1 if tiled else world) * steps / (ms / 1e3)
    flops = 2.0 * rows * MAP_SIZE * MAP_WINDOW * MAP_WINDOW * len(zp.n)
    ach = flops * steps / (ms / 1e3) / 1e12
    # denominator: the measured dense bf16 rate for the kind::f16 kernels (same tensor-pipe rate), half of it
    # for the kind::tf32 kernels.  `achieved` counts ALGORITHMIC flops only; `executed_tflops` adds what the
    # kernel really issues (three split terms, modes padded to 96, window rows padded to 16-tap groups).
    f16_kernel = prec in ("f16", "f16x3")
    peak = pk["bf16_tflops"] if f16_kernel else pk["bf16_tflops"] / 2.0
    pad_modes = 96.0 / len(zp.n) if prec != "fp32" else 1.0
    if f16_kernel:
        xs = -1.0 + 2.0 * np.arange(MAP_WINDOW) / (MAP_WINDOW - 1)
        rows_on = int(((xs[None, :] ** 2 + xs[:, None] ** 2) <= 1.0 + 1e-9).any(axis=1).sum())
        pad_taps = rows_on * (-(-MAP_WINDOW // 16)) * 16.0 / (MAP_WINDOW * MAP_WINDOW)
    else:
        pad_taps = 1.0
    executed = ach * {"fp32": 1, "tf32": 1, "tf32x3": 3, "f16": 1, "f16x3": 3}[prec] * pad_modes * pad_taps
    roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "traffic": NCU_TRAFFIC_MAP.get(prec) if not tiled else None,
            "peak_source": pk["source"] + (" bf16 dense (kind::f16 MMAs)" if f16_kernel else " bf16 / 2 (tf32 dense rate)"),
            "kernel": {"fp32": "map_simt_kernel<scores>", "tf32": "map_tc_kernel<scores>", "tf32x3": "map_tc_kernel<scores>",
                       "f16": "map_h_kernel<scores,x1>", "f16x3": "map_h_kernel<scores,x3>"}[prec],
            "algorithmic_flops_per_launch": flops,
            "executed_tflops": executed, "executed_frac": executed / peak}
    # end to end: numpy frame in, numpy scores out
    zp_host = ZPs(N_MAX, MAP_WINDOW, precision=args.precision, output="numpy")
    host = torch.empty((MAP_SIZE, MAP_SIZE), dtype=torch.float32, pin_memory=True)
    host.copy_(dimg)
    zp_host.symmetry_map(host.numpy(), FOLDS, row0=row0, rows=rows)      # warm-up: pinned staging, mempool growth
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_steps = 3
    for _ in range(e2e_steps):
        res = zp_host.symmetry_map(host.numpy(), FOLDS, row0=row0, rows=rows)
    dt = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
    e2e = {"value": mpix * (1 if tiled else world) * e2e_steps / dt, "unit": "Mpix/s", "h2d_bytes_per_step": MAP_SIZE * MAP_SIZE * 4,
           "d2h_bytes_per_step": int(res.size * 8), "steps": e2e_steps, "api": "ZPs.symmetry_map(numpy) -> numpy"}
    return {"metric": "symmetry_map_mpix_per_sec", "value": value, "unit": "Mpix/s", "ms_per_step": ms / steps,
            "steps": steps, "dtype": {"fp32": "f32", "tf32": "tf32", "tf32x3": "tf32x3(f32-grade)", "f16": "f16",
                                      "f16x3": "f16x3(f32-grade)"}[prec],
            "scaling": "strong" if tiled else "weak",
            "config": {"workload": f"symmetry map {MAP_SIZE}x{MAP_SIZE} n_max={N_MAX} window={MAP_WINDOW} "
                                   f"folds={FOLDS} (BASELINE configs[{3 if tiled else 1}])", "precision": prec,
                       "l2": "256 MiB scratch write between steps",
                       "parallelism": (f"one frame, {world} row bands with halo rows from the replicated frame, no collective"
                                       if tiled else f"one frame per GPU x{world}, no collective")},
            "roofline": roof, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="patches", choices=["patches", "map", "map4k"])
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "tf32", "tf32x3", "f16", "f16x3"])
    ap.add_argument("--batch", type=int, default=262144, help="patches per GPU per step")
    ap.add_argument("--e2e-batch", type=int, default=65536)
    ap.add_argument("--map-steps", type=int, default=20)
    ap.add_argument("--no-also", action="store_true", help="skip the secondary workload")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()

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
    pk = peaks()
    if args.workload == "map4k":
        primary = lambda *a: bench_map(*a, tiled=True)          # noqa: E731
        secondary = bench_patches
    else:
        primary = bench_patches if args.workload == "patches" else bench_map
        secondary = bench_map if args.workload == "patches" else bench_patches
    line = primary(torch, dist, rank, world, args, pk)
    line.setdefault("steps", args.steps)
    line.setdefault("scaling", "weak")
    line.update({"n_gpus": world, "warmup": args.warmup, "higher_is_better": True,
                 "vs_baseline": None, "data": "synthetic"})
    if not args.no_also:
        try:
            other = secondary(torch, dist, rank, world, args, pk)
            line["also"] = {k: other[k] for k in ("metric", "value", "unit", "ms_per_step", "dtype", "config",
                                                  "roofline", "e2e", "gpu_launches")}
        except Exception as exc:  # the secondary number must never sink the primary line
            line["also"] = {"error": f"{type(exc).__name__}: {exc}"}
    if rank == 0 and not args.no_cpu:
        cores = cpu_threads()
        if args.workload == "patches":
            v, times = cpu_patches(20000, 8)
            line["cpu_baseline"] = {"value": v, "unit": "patches/s", "cores": cores, "kind": "port",
                                    "sample": "20000 float32 64x64 patches, numpy.dot float64 (reference algorithm "
                                              f"_zps.py:146-157 via oracle port), best of {len(times)}"}
            if "also" in line and "error" not in line["also"]:
                mv, dt = cpu_map(384)
                line["also"]["cpu_baseline"] = {"value": mv, "unit": "Mpix/s", "cores": 1, "kind": "port",
                                                "sample": f"384x384 crop, 91 scipy.fftconvolve + rot_maps, {dt:.1f} s"}
        else:
            mv, dt = cpu_map(512, window=64 if args.workload == "map4k" else MAP_WINDOW)
            line["cpu_baseline"] = {"value": mv, "unit": "Mpix/s", "cores": 1, "kind": "port",
                                    "sample": f"512x512 crop, 91 scipy.fftconvolve (1 thread, reference algorithm "
                                              f"_zps.py:159-193 via oracle port) + rot_maps, {dt:.1f} s"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
