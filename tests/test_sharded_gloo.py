"""world_size-2 gloo test of the slice-sharded exchange (host logic of the N>1 path), on CPU.
The per-shard compute is stood in by the oracle; what is tested is the partition, the gather
layout, the min/max reduction and the index sharing."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from eitsynthai_b200 import sharded, synth
from oracle import imaging as O


def test_shard_ranges_cover_and_balance():
    for n in (320, 321, 7, 40):
        for w in (1, 2, 3, 8):
            r = [sharded.shard_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    assert [sharded.owner_of_series(s, 4) for s in range(6)] == [0, 1, 2, 3, 0, 1]


def _worker(rank, world, port, n_slices, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        S = 2
        z0, z1 = sharded.shard_range(n_slices, world, rank)
        rows_local, mm_local = [], []
        for s in range(S):
            vol, inst = synth.phantom_series(n_slices, seed=s, shuffle_seed=5 + s, size=64, z_range=(z0, z1))
            srt = vol[np.argsort(inst, kind="stable")]
            r = O.front_rows(srt)                                  # stand-in for eitb_front_rows on this shard
            rows_local.append(r)
            mm_local.append([int(r.min()), int(r.max())])
        # series 1 is FFS: its coronal image is z-reversed, which must happen on the gathered rows
        rows, mm = sharded.gather_rows(torch.from_numpy(np.stack(rows_local)), torch.tensor(mm_local, dtype=torch.int32), n_slices,
                                       flip_z=[1])
        sel = torch.zeros((S, 4), dtype=torch.int32)
        for s in range(S):
            if sharded.owner_of_series(s, world) == rank:
                sel[s] = torch.tensor([10 + s, 20 + s, 15 + s, 1], dtype=torch.int32)
        sel = sharded.share_selected(sel)
        np.savez(os.path.join(tmp, f"r{rank}.npz"), rows=rows.numpy(), mm=mm.numpy(), sel=sel.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_slices", [24, 25])
def test_exchange_world2(tmp_path, n_slices):
    port = 29500 + (os.getpid() + n_slices) % 2000
    mp.spawn(_worker, args=(2, port, n_slices, str(tmp_path)), nprocs=2, join=True)
    got = [np.load(tmp_path / f"r{r}.npz") for r in range(2)]
    for s in range(2):
        vol, inst = synth.phantom_series(n_slices, seed=s, shuffle_seed=None, size=64)
        want = O.front_rows(vol, "FFS" if s == 1 else "HFS")
        for g in got:
            assert np.array_equal(g["rows"][s], want)
            assert list(g["mm"][s]) == [int(want.min()), int(want.max())]
            assert np.array_equal(O.minmax_u8(g["rows"][s]), O.front_slice_norm(vol, "FFS" if s == 1 else "HFS"))
    assert np.array_equal(got[0]["sel"], got[1]["sel"])
    assert got[0]["sel"].tolist() == [[10, 20, 15, 1], [11, 21, 16, 1]]
