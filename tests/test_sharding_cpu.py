"""CPU, world_size-2 gloo: the row-band sharding plan (band + recompute halo per stage, band exchange of final output and
temporal history after every frame) reproduces the unsharded result exactly.

Each rank runs the CPU oracle but trashes, after every stage, the rows a sharded GPU rank would NOT have computed
(oracle._poison with the same halos the CUDA host code uses: ray trace +-32, temporal / a-trous +-16, AO vertical blur
+-8); the bands are then all-gathered with torch.distributed (gloo) exactly like hr_shard_exchange does with NCCL.
If a halo were too small, garbage would leak into a band and the comparison with the unsharded oracle would fail.
"""
import os
import sys

import numpy as np
import pytest

import oracle as O
import pyhr

W, H = 128, 96


def shard_rows(h, rank, world):
    tiles = (h + 7) // 8
    base, rem = divmod(tiles, world)
    t0 = rank * base + min(rank, rem)
    t1 = t0 + base + (1 if rank < rem else 0)
    return min(t0 * 8, h), min(t1 * 8, h)


def test_partition_matches_library():
    for h in (96, 48, 2160, 540):
        for world in (2, 3, 8):
            for r in range(world):
                assert shard_rows(h, r, world) == pyhr.shard_rows(h, r, world)


def _gather_bands(dist, arr, pass_h, rank, world, shift=0):
    """all-gather with unequal bands = one broadcast per owner (what hr_shard_exchange does inside an NCCL group)"""
    import torch
    out = arr.copy()
    for r in range(world):
        b, e = shard_rows(pass_h, r, world)
        b2, e2 = b << shift, (arr.shape[0] if e >= pass_h else e << shift)
        if e2 <= b2:
            continue
        t = torch.from_numpy(np.ascontiguousarray(arr[b2:e2]).view(np.uint8).copy())
        dist.broadcast(t, src=r)
        out[b2:e2] = t.numpy().view(arr.dtype).reshape(arr[b2:e2].shape)
    return out


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = pyhr.SynthScene(pyhr.SCENE_SHADOWS_TEST)
        tri, _ = sc.world_triangles()
        osc = O.Scene(tri, brute=True)
        bn = pyhr.blue_noise()
        ref_sh, ref_ao = O.ShadowsOracle(W, H, 0), O.AOOracle(W, H, 1)
        sh, ao = O.ShadowsOracle(W, H, 0), O.AOOracle(W, H, 1)
        sh.band = shard_rows(sh.H, rank, world)
        ao.band = shard_rows(ao.H, rank, world)
        f, prev = None, O.zero_gbuf_mips(W, H)
        for i in range(5):
            dx = 0.0 if i < 2 else 0.08 * (i - 1)
            f = pyhr.make_frame((dx, 14, 34), (dx, 3, 0), W, H, prev=f, num_frames=i)
            cur = O.GBufMips(pyhr.write_gbuffer(sc, f, W, H))
            ref_sh.render(osc, cur, prev, f, bn)
            ref_ao.render(osc, cur, prev, f, bn)
            sh.render(osc, cur, prev, f, bn)
            ao.render(osc, cur, prev, f, bn)
            prev = cur
            pp = f.ping_pong
            # exchange (shadows: prev_image, moments, final; ao: colour, history length, final)
            sh.prev_image[:] = _gather_bands(dist, sh.prev_image, sh.H, rank, world)
            sh.moments[pp][:] = _gather_bands(dist, sh.moments[pp], sh.H, rank, world)
            final_sh = _gather_bands(dist, sh.final, sh.H, rank, world)
            ao.color[pp][:] = _gather_bands(dist, ao.color[pp], ao.H, rank, world)
            ao.length[pp][:] = _gather_bands(dist, ao.length[pp], ao.H, rank, world)
            final_ao = _gather_bands(dist, ao.final, ao.H, rank, world, shift=1)
            for name, a, b in (("shadows final", final_sh, ref_sh.final), ("prev_image", sh.prev_image, ref_sh.prev_image),
                               ("moments", sh.moments[pp], ref_sh.moments[pp]), ("ao final", final_ao, ref_ao.final),
                               ("ao colour", ao.color[pp], ref_ao.color[pp]), ("ao length", ao.length[pp], ref_ao.length[pp])):
                if not np.array_equal(a, b):
                    q.put(f"rank {rank} frame {i}: {name} differs from the unsharded result")
                    return
        q.put("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_oracle_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p_ in procs:
        p_.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p_ in procs:
        p_.join(60)
    assert res == ["ok"] * world, res
