"""Host logic of the multi-GPU path on CPU: world_size-2 gloo process groups exercise the
sharding arithmetic and the ragged feature / row-band gathers of motif_learn_b200.parallel."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from motif_learn_b200 import parallel as par


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 91, 10000, 262144):
        for world in (1, 2, 3, 4, 8):
            spans = [par.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == par.shard_sizes(n, world)
    assert par.row_band(4096, 3, 8) == (1536, 512)
    with pytest.raises(ValueError):
        par.shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_patches, height, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)                       # same replicated input on every rank
        patches = torch.from_numpy(rng.random((n_patches, 4, 4)))
        calls = []

        def fake_transform(x):                               # stands in for the CUDA projection
            calls.append(int(x.shape[0]))
            return x.reshape(x.shape[0], -1)[:, :5] * 2.0

        feats = par.transform_patches_sharded(fake_transform, patches)
        lo, hi = par.shard_range(n_patches, rank, world)
        assert calls == [hi - lo]
        local_only = par.transform_patches_sharded(fake_transform, patches, gather=False)
        assert local_only.shape[0] == hi - lo

        image = torch.from_numpy(rng.random((height, 6)))

        def fake_band(row0, rows):
            return torch.stack([image[row0:row0 + rows] * (f + 1) for f in range(3)])

        smap = par.symmetry_map_sharded(fake_band, height)
        ragged = par.gather_ragged(torch.full((rank + 1, 2), float(rank)), dim=0)
        # gather_rows: equal blocks land straight in a preallocated buffer, ragged blocks via padding
        even_out = torch.full((2 * 3, 4), -1.0)
        got = par.gather_rows(torch.full((3, 4), float(rank + 10)), out=even_out, sizes=[3, 3])
        assert got is even_out
        rows = par.gather_rows(torch.arange((rank + 2) * 5, dtype=torch.float64).reshape(rank + 2, 5) + 100 * rank)
        torch.save({"feats": feats, "smap": smap, "ragged": ragged, "even": even_out, "rows": rows},
                   os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_patches,height", [(11, 9), (8, 16)])
def test_world2_gloo_gathers(tmp_path, n_patches, height):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_patches, height, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(0)
    patches = torch.from_numpy(rng.random((n_patches, 4, 4)))
    image = torch.from_numpy(rng.random((height, 6)))
    want_feats = patches.reshape(n_patches, -1)[:, :5] * 2.0
    want_map = torch.stack([image * (f + 1) for f in range(3)])
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        assert torch.equal(got["feats"], want_feats)
        assert torch.equal(got["smap"], want_map)
        assert torch.equal(got["ragged"], torch.tensor([[0.0, 0.0], [1.0, 1.0], [1.0, 1.0]]))
        assert torch.equal(got["even"], torch.cat([torch.full((3, 4), 10.0), torch.full((3, 4), 11.0)]))
        want_rows = torch.cat([torch.arange(10, dtype=torch.float64).reshape(2, 5),
                               torch.arange(15, dtype=torch.float64).reshape(3, 5) + 100])
        assert torch.equal(got["rows"], want_rows)
