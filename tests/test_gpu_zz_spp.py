"""spp > 1 (NOT in the reference): `spp` rays per pixel with sample index num_frames * spp + s, an
8-bit image of unoccluded-ray counts instead of the 1-bit mask, visibility = count / spp in the temporal stage.

The CUDA kernels of this mode (k_ray_trace_count, k_temporal_count) mirror the validated 1-spp kernels; the ray counts must be
exact against the oracle and the denoised images within the same tolerances as the 1-spp chains (confirmed on a B200 with the
last GPU seconds of round 1: all three tests passed).
"""
import numpy as np
import pytest

import oracle as O
import pyhr

pytestmark = pytest.mark.gpu

W, H = 256, 144


def f16(a):
    return np.ascontiguousarray(a).view(np.float16).astype(np.float32)


def rmse(a, b):
    return float(np.sqrt(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)))


@pytest.mark.parametrize("spp", [2, 4])
def test_spp_counts_exact_and_denoise_within_tolerance(spp):
    sc = pyhr.SynthScene(pyhr.SCENE_SHADOWS_TEST)
    tri, _ = sc.world_triangles()
    osc = O.Scene(tri, brute=True)
    bn = pyhr.blue_noise()
    ctx = pyhr.Context(0)
    try:
        ctx.set_bluenoise(*bn)
        ctx.build_scene(sc)
        ctx.gbuffer_create(W, H)
        sh, ao = pyhr.Pass(ctx, "shadows", W, H, 0), pyhr.Pass(ctx, "ao", W, H, 1)
        sh.params.spp = ao.params.spp = spp
        osh, oao = O.ShadowsOracle(W, H, 0, spp=spp), O.AOOracle(W, H, 1, spp=spp)
        f, prev = None, O.zero_gbuf_mips(W, H)
        for i in range(5):
            dx = 0.0 if i < 3 else 0.05 * (i - 2)
            f = pyhr.make_frame((dx, 14.0, 34.0), (dx, 3.0, 0.0), W, H, prev=f, num_frames=i)
            g = pyhr.write_gbuffer(sc, f, W, H)
            ctx.gbuffer_upload(f.ping_pong, g)
            sh.render(f)
            ao.render(f)
            cur = O.GBufMips(g)
            osh.render(osc, cur, prev, f, bn)
            oao.render(osc, cur, prev, f, bn)
            prev = cur
            # counts are decisions of the deterministic chain: exact
            assert np.array_equal(sh.download(0), osh.count), f"frame {i}: shadows ray counts not exact"
            assert np.array_equal(ao.download(0), oao.count), f"frame {i}: AO ray counts not exact"
            assert np.array_equal(sh.download(6), osh.tile_flags) and np.array_equal(ao.download(6), oao.tile_flags)
            assert np.array_equal(f16(sh.download(4))[..., 2], O.h2f(osh.cur_moments)[..., 2]), f"frame {i}: history length"
            for which, ref, mx in ((1, osh.temporal, 2e-3), (2, osh.atrous_out, 4e-3), (100, osh.final, 4e-3)):
                c, o = f16(sh.download(which)), O.h2f(ref)
                assert rmse(c, o) <= 1e-3 and np.abs(c - o).max() <= mx, f"frame {i} shadows image {which}: rmse {rmse(c, o)} max {np.abs(c - o).max()}"
            for which, ref, mx in ((1, oao.temporal, 2e-3), (2, oao.blur[1], 4e-3), (100, oao.final, 4e-3)):
                c, o = f16(ao.download(which)), O.h2f(ref)
                assert rmse(c, o) <= 1e-3 and np.abs(c - o).max() <= mx, f"frame {i} AO image {which}: rmse {rmse(c, o)} max {np.abs(c - o).max()}"
        assert osh.count.max() == spp and oao.count.max() == spp
        sh.destroy()
        ao.destroy()
    finally:
        ctx.close()


def _disabled_spp_rejected_when_sharded():  # sharded spp > 1 is supported since round 2 (band-local count images)
    ctx = pyhr.Context(0)
    try:
        sc = pyhr.SynthScene(pyhr.SCENE_SHADOWS_TEST)
        ctx.set_bluenoise(*pyhr.blue_noise())
        ctx.build_scene(sc)
        ctx.gbuffer_create(64, 48)
        ctx.shard_config(0, 2)
        sh = pyhr.Pass(ctx, "shadows", 64, 48, 0)
        sh.params.spp = 2
        f = pyhr.make_frame((0.0, 14.0, 34.0), (0.0, 3.0, 0.0), 64, 48)
        ctx.gbuffer_upload(f.ping_pong, pyhr.write_gbuffer(sc, f, 64, 48))
        with pytest.raises(pyhr.HrError):
            sh.render(f)
        sh.destroy()
    finally:
        ctx.close()
