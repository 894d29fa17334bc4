"""Multi-GPU sharding of the Zernike path: one process per GPU, torch.distributed plumbing.

The path has no mid-computation exchange: patch stacks shard by contiguous
patch ranges, frame batches by frames, a single large frame by row bands whose halo rows are
read from the (replicated, <= 67 MB) frame itself -- no halo exchange.  The only collective is
the optional final gather of the feature / score shards (NCCL over NVLink on GPUs, gloo in the
CPU tests).  The reference is single-process; nothing here mirrors reference code.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) of ``n_items`` for ``rank``: sizes differ by at most one,
    the first ``n_items % world`` ranks get the extra item, every item is owned exactly once."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(int(n_items), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_items: int, world: int) -> List[int]:
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def row_band(height: int, rank: int, world: int) -> Tuple[int, int]:
    """(row0, rows) of the output row band owned by ``rank`` for image-tile sharding.  The band's
    windows reach size//2 rows above and size-1-size//2 below; those halo rows are read from the
    full frame every rank holds (zero beyond the image), so bands need no exchange."""
    lo, hi = shard_range(height, rank, world)
    return lo, hi - lo


def dist_info():
    """(rank, world) of the default process group, (0, 1) when torch.distributed is not initialised."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def gather_ragged(local, dim: int = 0, sizes: Sequence[int] | None = None, group=None):
    """All-gather tensors whose extent along ``dim`` differs per rank; returns the concatenation
    in rank order on every rank.  Sizes are exchanged first unless given; shards are padded to
    the largest one so a single fixed-size all_gather moves the payload."""
    import torch
    import torch.distributed as dist
    rank, world = dist_info()
    if world == 1:
        return local
    if sizes is None:
        mine = torch.tensor([local.shape[dim]], dtype=torch.int64, device=local.device)
        all_sizes = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(all_sizes, mine, group=group)
        sizes = [int(s.item()) for s in all_sizes]
    biggest = max(sizes)
    moved = local.movedim(dim, 0).contiguous()
    if moved.shape[0] < biggest:
        pad = torch.zeros((biggest - moved.shape[0],) + tuple(moved.shape[1:]), dtype=moved.dtype, device=moved.device)
        moved = torch.cat([moved, pad], dim=0)
    parts = [torch.empty_like(moved) for _ in range(world)]
    dist.all_gather(parts, moved, group=group)
    out = torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)
    return out.movedim(0, dim)


def gather_rows(local, out=None, sizes: Sequence[int] | None = None, group=None):
    """All-gather of row blocks ``(n_r, ...)`` into ``out`` ``(sum n_r, ...)`` (allocated when None), in rank
    order, on every rank -- the final feature gather of the path.  Equal blocks move with ONE
    ``all_gather_into_tensor`` straight into ``out`` (no padding, no concatenation); ragged blocks are padded
    to the largest one in a scratch buffer first."""
    import torch
    import torch.distributed as dist
    rank, world = dist_info()
    if world == 1:
        if out is None:
            return local
        out.copy_(local)
        return out
    if sizes is None:
        mine = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        every = torch.empty(world, dtype=torch.int64, device=local.device)
        dist.all_gather_into_tensor(every, mine, group=group)
        sizes = [int(v) for v in every.tolist()]
    total = int(sum(sizes))
    if out is None:
        out = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    local = local.contiguous()
    if len(set(sizes)) == 1:
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    biggest = max(sizes)
    padded = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    scratch = torch.empty((world * biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(scratch, padded, group=group)
    at = 0
    for r, s in enumerate(sizes):
        out[at:at + s] = scratch[r * biggest:r * biggest + s]
        at += s
    return out


def transform_patches_sharded(transform: Callable, patches, gather: bool = True):
    """Project this rank's contiguous shard of a replicated patch stack and (optionally) gather the
    feature rows of all ranks.  ``transform`` maps a patch tensor (n, k, k) to a tensor (n, F) --
    e.g. ``lambda x: zps.transform(x).data``."""
    rank, world = dist_info()
    lo, hi = shard_range(int(patches.shape[0]), rank, world)
    local = transform(patches[lo:hi])
    if not gather or world == 1:
        return local
    return gather_ragged(local, dim=0, sizes=shard_sizes(int(patches.shape[0]), world))


def symmetry_map_sharded(band_fn: Callable, height: int, gather: bool = True):
    """Image-tile sharding of one frame: ``band_fn(row0, rows)`` returns this rank's (F, rows, W)
    score band (e.g. ``lambda r0, r: zps.symmetry_map(img, folds, row0=r0, rows=r)``); bands are
    concatenated along the row axis in rank order."""
    rank, world = dist_info()
    row0, rows = row_band(height, rank, world)
    band = band_fn(row0, rows)
    if not gather or world == 1:
        return band
    return gather_ragged(band, dim=1, sizes=shard_sizes(height, world))


# ---------------------------------------------------------------------------------------------------------------
# K5 fused into the projection kernel: peer-memory result arrays
# ---------------------------------------------------------------------------------------------------------------
class _CudaBlock:
    """A library-owned device allocation seen by torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(int(v) for v in shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class PeerArray:
    """A ``(rows, cols)`` float32 array that exists once per rank and is writable by EVERY rank of the node: each
    process allocates its copy through the C ABI (``zb200_peer_buffer_alloc``), the CUDA IPC handles travel over
    ``torch.distributed`` (plumbing), and every process maps the others' copies (``zb200_peer_buffer_open``).
    ``ZPs.transform_allgather`` then lets the projection kernel write its output rows into all copies over NVLink
    while it is still computing -- the final feature gather of SURVEY.md 8e without a separate collective.

    ``local`` is this rank's copy as a CUDA tensor; ``fence()`` orders the copies: after it (on the current
    stream) every rank's rows are complete in ``local``; call it once more before the NEXT round of writes if a
    consumer on another rank may still be reading (``begin()``)."""

    def __init__(self, rows: int, cols: int, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib
        self.rows, self.cols, self.group = int(rows), int(cols), group
        self.rank, self.world = dist_info()
        lib = _lib.load()
        nbytes = max(16, self.rows * self.cols * 4)
        ptr, handle = C.c_void_p(), (C.c_ubyte * 64)()
        _lib.check(lib.zb200_peer_buffer_alloc(nbytes, C.byref(ptr), handle), "peer_buffer_alloc")
        self._lib, self._own = lib, ptr
        self.local = torch.as_tensor(_CudaBlock(ptr.value, (self.rows, self.cols), "<f4"), device="cuda")
        self.ptrs = [None] * self.world
        self.ptrs[self.rank] = int(ptr.value)
        self._opened = []
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            for r, hb in enumerate(handles):
                if r == self.rank:
                    continue
                other = C.c_void_p()
                buf = (C.c_ubyte * 64).from_buffer_copy(hb)
                _lib.check(lib.zb200_peer_buffer_open(buf, C.byref(other)), f"peer_buffer_open(rank {r})")
                self.ptrs[r] = int(other.value)
                self._opened.append(other)
            self._flag = torch.zeros(1, dtype=torch.int32, device="cuda")

    def row_ptr(self, rank: int, row0: int) -> int:
        return self.ptrs[rank] + int(row0) * self.cols * 4

    def fence(self):
        """Stream-ordered: returns (on the stream) when every rank's kernels queued before it have finished."""
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self._flag, group=self.group)

    begin = fence

    def close(self):
        """Unmap the peers' copies and free the own one (collective: every rank must call it)."""
        import torch
        import torch.distributed as dist
        from . import _lib
        torch.cuda.synchronize()
        for p in self._opened:
            _lib.check(self._lib.zb200_peer_buffer_close(p), "peer_buffer_close")
        self._opened = []
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)            # nobody still maps the allocation that is freed next
        self.local = None
        if self._own is not None:
            _lib.check(self._lib.zb200_peer_buffer_free(self._own), "peer_buffer_free")
            self._own = None


def push_band(arr: PeerArray, band, row0: int, stream=None):
    """Forward ``band`` (rows, cols) -- already stored at rows [row0, row0+rows) of ``arr.local`` or any other
    CUDA tensor -- into the same rows of every peer's copy with the copy engines (no SM work; used for the score
    bands of the tiled dense map, whose kernel stores 4-byte words at a 16-byte stride -- a poor NVLink pattern)."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    band = band.contiguous()
    rows = int(band.shape[0])
    st = C.c_void_p(_lib.current_stream_ptr() if stream is None else int(stream))
    for r in range(arr.world):
        if r == arr.rank:
            if band.data_ptr() != arr.row_ptr(r, row0):
                arr.local[row0:row0 + rows].copy_(band.reshape(rows, arr.cols))
            continue
        _lib.check(lib.zb200_peer_copy_2d(arr.row_ptr(r, row0), arr.cols * 4, int(band.data_ptr()), arr.cols * 4,
                                          arr.cols * 4, rows, st), "peer_copy_2d")


def push_score_bands(arr: PeerArray, band, row0: int, height: int, stream=None):
    """Image-tile sharding of ONE frame (BASELINE config 4): ``arr`` is every rank's copy of the full ``(F, H, W)``
    score map (a PeerArray of ``F*H`` rows), ``band`` this rank's ``(F, rows, W)`` result for rows ``[row0, row0+rows)``.
    One 2-D copy per peer moves all F planes (source pitch rows*W, destination pitch H*W), driven by the copy
    engines; the local copy is filled the same way."""
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    band = band.contiguous()
    n_f, rows, width = (int(v) for v in band.shape)
    st = C.c_void_p(_lib.current_stream_ptr() if stream is None else int(stream))
    for r in range(arr.world):
        _lib.check(lib.zb200_peer_copy_2d(arr.row_ptr(r, row0), height * width * 4, int(band.data_ptr()), rows * width * 4,
                                          rows * width * 4, n_f, st), "peer_copy_2d")


_side_streams: dict = {}


def symmetry_map_allgather(zps, image, n_folds, arr: PeerArray, row0: int, rows: int, height: int, n_sub: int = 4,
                           p=2, m_unselect=None):
    """This rank's row band ``[row0, row0 + rows)`` of the fused symmetry map of one frame, computed in ``n_sub``
    sub-bands whose scores are forwarded to every rank's copy of the ``(F, H, W)`` map (``arr``: a PeerArray of
    ``F*H`` rows) by the copy engines WHILE the next sub-band is being computed -- only the last sub-band's copies are
    exposed.  Sub-bands start on even rows (the map kernel pairs rows by absolute parity), so the result is the
    single-call map bit for bit.  Call ``arr.begin()`` before and ``arr.fence()`` after."""
    import torch
    dev = torch.cuda.current_device()
    side = _side_streams.get(dev)
    if side is None:
        side = _side_streams[dev] = torch.cuda.Stream()
    main = torch.cuda.current_stream()
    n_sub = max(1, min(int(n_sub), max(1, rows // 2)))
    step = -(-rows // n_sub)
    step += step & 1
    keep = []
    for r in range(row0, row0 + rows, step):
        n = min(step, row0 + rows - r)
        band = zps.symmetry_map(image, n_folds, p=p, m_unselect=m_unselect, row0=r, rows=n)
        ready = torch.cuda.Event()
        ready.record(main)
        side.wait_event(ready)
        push_score_bands(arr, band, r, height, stream=side.cuda_stream)
        band.record_stream(side)
        keep.append(band)
    done = torch.cuda.Event()
    done.record(side)
    main.wait_event(done)
    return keep
