"""``series_features`` -- the reference's patch route over a series of frames (BASELINE config 5).

What a user of the reference writes per frame (notebook "2 How to use ZPs": cells 3-13):

    pts = local_max(img, min_distance, threshold)                 # mtflearn/features/_local_max_v2.py:46-66
    ps  = KeyPoints(pts, img, size).extract_patches(size)         # mtflearn/features/_keypoint.py:53-78
    X   = np.abs(ZPs(n_max, size).fit_transform(ps).to_complex().data)   # _zps.py:146-157, _zmoments.py:300-316

Here the same chain runs on the GPU for every frame of an in-situ series.  One frame is a handful of small
kernels (peak detection needs two host round trips for its counts), so a single in-order stream leaves the GPU
idle between them; the frames are therefore spread over a few worker threads, each with its own CUDA stream
(the C ABI takes the stream per call and ctypes releases the GIL), and their kernels interleave on the device.
Results come back in frame order.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .. import _lib
from ._keypoint import clear_border
from ._local_max import local_max


_streams: dict = {}


def _worker_streams(torch, count: int):
    """Persistent side streams per device: torch's caching allocator keeps its free blocks per stream, so fresh
    streams on every call would turn every patch-stack allocation into a cudaMalloc."""
    key = torch.cuda.current_device()
    have = _streams.setdefault(key, [])
    while len(have) < count:
        have.append(torch.cuda.Stream())
    return have[:count]


def series_features(zps, frames, min_distance, threshold=None, kind: str = "abs", workers: int = 3, fused=None):
    """Peaks -> patches -> Zernike features for every frame of ``frames`` (a CUDA tensor (F, H, W) or a sequence
    of 2-D CUDA tensors / numpy arrays).  ``zps`` is the ``ZPs`` transformer; ``kind`` as in
    ``ZPs.transform_peaks`` ('real' returns the (P, M) moment tensors, 'abs' the rotation-invariant |Zc|, ...).

    Returns ``(features, points)``: two lists with one entry per frame -- the feature tensor (CUDA, float32 /
    complex64; rows in the order of ``points``) and the kept peak coordinates ((P, 2) int64 numpy, (x, y),
    brightest first, already filtered by ``clear_border``).  ``fused`` is passed to ``ZPs.transform_peaks`` (None: the
    mirror-folded kernel gathers the windows itself where the plan has it; False: gather kernel + projection)."""
    torch = _lib.require_cuda()
    n_frames = len(frames)
    if n_frames == 0:
        return [], []
    caller = torch.cuda.current_stream()
    ready = torch.cuda.Event()
    ready.record(caller)
    workers = max(1, min(int(workers), n_frames))
    streams = _worker_streams(torch, workers)

    def one(index: int, stream):
        with torch.cuda.stream(stream):
            frame = frames[index]
            if not (_lib_is_cuda(frame)):
                frame = torch.from_numpy(np.ascontiguousarray(frame, dtype=np.float32)).cuda(non_blocking=True)
            pts = local_max(frame, min_distance, threshold)
            kept = clear_border(pts, tuple(frame.shape), zps.size)
            feats = zps.transform_peaks(frame, kept, kind, fused=fused)
            data = feats.data if kind == "real" else feats
            for t in (data if isinstance(data, tuple) else (data,)):
                t.record_stream(caller)                 # allocated on the worker stream, consumed on the caller's
            return feats, kept

    def run(w: int):
        stream = streams[w]
        stream.wait_event(ready)                        # the frames were produced on the caller's stream
        out = [(i, one(i, stream)) for i in range(w, n_frames, workers)]
        done = torch.cuda.Event()
        done.record(stream)
        return out, done

    if workers == 1:
        parts = [run(0)]
    else:
        with ThreadPoolExecutor(max_workers=workers) as pool:
            parts = list(pool.map(run, range(workers)))
    features, points = [None] * n_frames, [None] * n_frames
    for out, done in parts:
        caller.wait_event(done)                         # later work on the caller's stream sees every result
        for i, (f, k) in out:
            features[i], points[i] = f, k
    return features, points


def _lib_is_cuda(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch" and x.is_cuda
