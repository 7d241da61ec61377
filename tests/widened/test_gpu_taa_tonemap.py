"""GPU parity of the post-processing passes through the C ABI: hr_taa_render / hr_tonemap_render vs the oracle
(oracle/orc_post.cpp) on the very images the kernels read (the inputs are downloaded from the device, so no upstream tolerance
leaks in).  Bars: TAA bit-exact (it is a bit-specified stage, csrc/post_px.cuh); tone map within 1 UNORM8 step (powf of the device
vs the host's libm), >= 99.5 % of the bytes identical."""
import numpy as np
import pytest

import oracle as O
import pyhr

pytestmark = pytest.mark.gpu

W, H = 256, 144


def u16(a):
    return np.ascontiguousarray(a).view(np.uint16)


def same_halves(a, b):
    """equal as binary16 VALUES: bit-identical except that -0 == +0 (fmax(-0, +0) may return either zero, IEEE 754 leaves it open)"""
    return np.array_equal(np.ascontiguousarray(a).view(np.float16), np.ascontiguousarray(b).view(np.float16), equal_nan=True)


def test_taa_and_tonemap_match_the_oracle_over_a_jittered_pan():
    sc = pyhr.SynthScene(pyhr.SCENE_SHADOWS_TEST)
    ctx = pyhr.Context(0)
    ctx.set_bluenoise(*pyhr.blue_noise())
    ctx.build_scene(sc)
    ctx.gbuffer_create(W, H)
    sh, ao, de = pyhr.Pass(ctx, "shadows", W, H, 0), pyhr.Pass(ctx, "ao", W, H, 1), pyhr.DeferredPass(ctx, W, H)
    de.params.env_color[0], de.params.env_color[1], de.params.env_color[2] = 0.3, 0.4, 0.6
    # the reference's visualisation modes: TAA over the deferred image (RGBA16F), the shadows output (RG16F at full resolution) and the
    # upsampled AO (R16F); `as written` = reset_every_frame 1, `intended` = 0 (hr_taa_params)
    cases = {"final": (de, 1), "final_intended": (de, 0), "shadows": (sh, 1), "ao_intended": (ao, 0)}
    taa = {k: pyhr.TAAPass(ctx, W, H) for k in cases}
    orc = {}
    for k, (_, reset) in cases.items():
        taa[k].params.reset_every_frame = reset
        orc[k] = O.TAAOracle(W, H, reset_every_frame=reset)
    taa["shadows"].params.sharpen = 0
    orc["shadows"].sharpen = 0
    tm, tm1 = pyhr.TonemapPass(ctx, W, H), pyhr.TonemapPass(ctx, W, H)
    tm.params.exposure = 1.5
    tm1.params.single_channel = 1
    f, prev_j = None, np.zeros(2, np.float32)
    launches0 = ctx.launch_count()
    for i in range(5):
        f = pyhr.make_frame((0.05 * max(0, i - 2), 14.0, 34.0), (0.0, 3.0, 0.0), W, H, prev=f, num_frames=i)
        j = pyhr.taa_jitter(i, W, H)
        assert np.array_equal(j, O.taa_jitter(i, W, H))
        pyhr.apply_jitter(f, j, prev_j)  # update_uniforms: the projection carries the jitter, ubo.current_prev_jitter = (current, previous)
        prev_j = j
        ctx.gbuffer_render(f.ping_pong, f)
        sh.render(f)
        ao.render(f)
        de.render(f, sh, ao, None, None)
        depth = ctx.gbuffer_download(f.ping_pong, 0, 0, W, H)
        gb2 = u16(ctx.gbuffer_download(f.ping_pong, 0, 2, W, H))
        for k, (src, _) in cases.items():
            cur = u16(src.download(100))
            assert cur.shape[:2] == (H, W)
            taa[k].render(f, src)
            want = orc[k].render(f, cur, depth, gb2)
            got = u16(taa[k].download(100))
            assert got.shape == (H, W, 4)
            assert same_halves(got, want), f"frame {i}, TAA over {k}: {np.count_nonzero(got != want)} of {got.size} halves differ"
            assert same_halves(u16(taa[k].download(0)), orc[k].img[1 - f.ping_pong]), f"frame {i}, {k}: history image"
        # tone map of the resolved frame (and the grey-scale mode over the AO resolve)
        for tone, src_k, kw in ((tm, "final", dict(exposure=1.5)), (tm1, "ao_intended", dict(single_channel=1))):
            tone.render(taa[src_k])
            got = tone.download(100)
            want = O.tonemap(u16(taa[src_k].download(100)), **kw)
            assert got.shape == want.shape == (H, W, 4) and got.dtype == np.uint8
            d = np.abs(got.astype(np.int16) - want.astype(np.int16))
            assert d.max() <= 1 and np.mean(d == 0) >= 0.995, f"frame {i}: tone map max diff {d.max()}, identical {np.mean(d == 0):.4f}"
            assert np.all(got[..., 3] == 255)
    out = u16(taa["final"].download(100)).view(np.float16).astype(np.float32)
    assert out[..., :3].max() <= 1.0 and out[..., :3].min() >= 0.0 and out[..., :3].mean() > 0.01 and np.all(out[..., 3] == 1.0)
    assert ctx.launch_count() > launches0
    # error behaviour: a tone map cannot read itself, TAA needs a pass that has rendered at full resolution
    with pytest.raises(pyhr.HrError):
        tm.render(tm)
    fresh = pyhr.DeferredPass(ctx, W, H)
    with pytest.raises(pyhr.HrError, match="no full-resolution final output"):
        taa["final"].render(f, fresh)
    for p in list(taa.values()) + [tm, tm1, fresh, sh, ao, de]:
        p.destroy()
    ctx.close()
