"""N > 1 path on CPU: world_size-2 gloo processes shard frame pairs / reference rows, compute their block (the oracle stands in
for the per-GPU kernel here -- the host-side partition + gather logic is what is under test) and gather on rank 0; the
result must equal the single-process result."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from feature_tracker_b200 import sharding  # noqa: E402


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 1000, 1001):
        for world in (1, 2, 3, 8):
            blocks = [sharding.shard_bounds(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, tmpdir):
    import torch.distributed as dist
    from feature_tracker_b200 import synthetic as S
    from oracle import pyoracle as po
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = po.OracleLib()
    counts = [30, 0, 12, 25, 7]
    pairs = [S.make_pair(96, 128, max(c, 1), pair_id=60 + p, border=8) for p, c in enumerate(counts)]
    offsets = np.concatenate([[0], np.cumsum(counts)])
    ref_uv = np.concatenate([pairs[p][2][:c] for p, c in enumerate(counts)])
    params = po.make_params("basic", "fast", half=4)

    def track_block(lo, hi, local_offsets, local_uv):
        uvs, sts = [], []
        for k, p in enumerate(range(lo, hi)):
            n = local_offsets[k + 1] - local_offsets[k]
            if n == 0:
                continue
            ref, cur = pairs[p][0], pairs[p][1]
            _, uv, st = oracle.klt_track(params, oracle.pyramid_build(ref, 2), oracle.pyramid_build(cur, 2), local_uv[local_offsets[k]:local_offsets[k + 1]])
            uvs.append(uv), sts.append(st)
        if not uvs:
            return np.zeros((0, 2), np.float32), np.zeros(0, np.uint8)
        return np.concatenate(uvs), np.concatenate(sts)

    cur_uv, status = sharding.track_sharded(track_block, offsets, ref_uv, world, rank)
    rb, cb, _, _, _ = S.make_brief_sets(101, 90, seed=4)
    idx = sharding.match_sharded(lambda lo, hi: oracle.match_brief_force(rb[lo:hi], cb, 60.0)[1], rb.shape[0], world, rank)
    if rank == 0:
        full_uv, full_st = track_block(0, len(counts), offsets.astype(np.int32), ref_uv)
        assert (cur_uv.view(np.uint32) == full_uv.view(np.uint32)).all() and (status == full_st).all()
        assert (idx == oracle.match_brief_force(rb, cb, 60.0)[1]).all()
        open(os.path.join(tmpdir, "ok"), "w").write("ok")
    else:
        assert cur_uv is None and status is None and idx is None
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok")
