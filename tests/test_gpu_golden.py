"""The CUDA path against the COMMITTED golden fixture (tests/golden/shadows_ao_256x144_seq12.npz, generated from the oracle
by tests/golden/make_golden.py): the 8-static + 4-panning-frame sequence of test_gpu_parity.py is driven through the C ABI
and the state after the last frame is compared with the stored images — no oracle code runs in this test.
Visibility masks, tile classification and history lengths bit-exact; tolerance-checked images within 1e-3 RMSE."""
import os
import sys

import numpy as np
import pytest

import oracle as O  # noqa: F401  (half <-> float helpers only)
import pyhr

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden import seq12_frames  # noqa: E402

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "shadows_ao_256x144_seq12.npz")


def f16(a):
    return np.ascontiguousarray(a).view(np.float16).astype(np.float32)


def rmse(a, b):
    return float(np.sqrt(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)))


def test_sequence_matches_committed_golden():
    W, H = 256, 144
    g = np.load(GOLDEN)
    sc = pyhr.SynthScene(pyhr.SCENE_SHADOWS_TEST)
    ctx = pyhr.Context(0)
    try:
        ctx.set_bluenoise(*pyhr.blue_noise())
        ctx.build_scene(sc)
        ctx.gbuffer_create(W, H)
        sh, ao = pyhr.Pass(ctx, "shadows", W, H, 0), pyhr.Pass(ctx, "ao", W, H, 1)
        for f in seq12_frames(W, H):
            ctx.gbuffer_upload(f.ping_pong, pyhr.write_gbuffer(sc, f, W, H))
            sh.render(f)
            ao.render(f)
        # ---- shadows (same tolerances as tests/test_gpu_parity.py::check_all) ----
        assert np.array_equal(sh.download(0), g["sh_mask"]), "shadow mask not bit-exact against the golden fixture"
        assert np.array_equal(sh.download(6), g["sh_tiles"]), "shadow tile classification differs from the golden fixture"
        m_c, m_o = f16(sh.download(4)), O.h2f(g["sh_moments"])
        assert np.array_equal(m_c[..., 2], m_o[..., 2]), "history length differs from the golden fixture"
        assert np.abs(m_c - m_o)[..., :2].max() <= 2e-3
        for which, key, mx in ((1, "sh_temporal", 2e-3), (2, "sh_atrous", 4e-3), (5, "sh_prev_image", 4e-3), (100, "sh_final", 4e-3)):
            c, o = f16(sh.download(which)), O.h2f(g[key])
            assert c.shape == o.shape and rmse(c, o) <= 1e-3 and np.abs(c - o).max() <= mx, f"{key}: rmse {rmse(c, o)} max {np.abs(c - o).max()}"
        # ---- ambient occlusion ----
        assert np.array_equal(ao.download(0), g["ao_mask"]), "AO mask not bit-exact against the golden fixture"
        assert np.array_equal(ao.download(6), g["ao_tiles"]), "AO tile classification differs from the golden fixture"
        assert np.array_equal(f16(ao.download(4)), O.h2f(g["ao_length"])), "AO history length differs from the golden fixture"
        for which, key, mx in ((1, "ao_temporal", 2e-3), (2, "ao_blur", 4e-3), (100, "ao_final", 4e-3)):
            c, o = f16(ao.download(which)), O.h2f(g[key])
            assert c.shape == o.shape and rmse(c, o) <= 1e-3 and np.abs(c - o).max() <= mx, f"{key}: rmse {rmse(c, o)} max {np.abs(c - o).max()}"
        sh.destroy()
        ao.destroy()
    finally:
        ctx.close()


def close(a, b, name, r=1e-3, m=6e-3):
    """rmse and max-abs bounds relative to max(1, |reference|), as in tests/test_gpu_gi_refl.py"""
    assert a.shape == b.shape, name
    d = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    e = float(np.sqrt(np.mean(d.astype(np.float64) ** 2)))
    assert e <= r, f"{name}: rmse {e}"
    assert d.max() <= m, f"{name}: max abs {d.max()}"


def test_ddgi_reflections_sequence_matches_committed_golden():
    """DDGI (K18-K21) + half-res reflections (K12, K14-K17) after 3 static + 2 panning frames against
    tests/golden/ddgi_reflections_192x112_seq5.npz (screen-space images; ray lengths, tile flags and history length exact)."""
    from make_golden import GI_H, GI_W, gi_frames, gi_params
    g = np.load(os.path.join(os.path.dirname(GOLDEN), "ddgi_reflections_192x112_seq5.npz"))
    sc = pyhr.SynthScene(pyhr.SCENE_SHADOWS_TEST)
    ctx = pyhr.Context(0)
    try:
        ctx.set_bluenoise(*pyhr.blue_noise())
        ctx.build_scene(sc)
        ctx.gbuffer_create(GI_W, GI_H)
        dd, rf = pyhr.DDGIPass(ctx, GI_W, GI_H, 0), pyhr.ReflectionsPass(ctx, GI_W, GI_H, 1)
        gi_params(dd.params, rf.params)
        for i, f, rot in gi_frames():
            ctx.gbuffer_upload(f.ping_pong, pyhr.write_gbuffer(sc, f, GI_W, GI_H))
            dd.render(f, rot)
            rf.render(f, dd)
        close(f16(dd.download(4)), O.h2f(g["ddgi_sample"]), "ddgi sample")
        rt_c, rt_o = f16(rf.download(0)), O.h2f(g["refl_rt"])
        assert np.array_equal(rt_c[..., 3], rt_o[..., 3]), "reflection ray length (hit / miss / t) not exact against the golden fixture"
        close(rt_c[..., :3], rt_o[..., :3], "reflections ray trace", 1e-3, 0.02)
        assert np.array_equal(rf.download(6), g["refl_tiles"]), "reflections tile classification differs from the golden fixture"
        close(f16(rf.download(1)), O.h2f(g["refl_temporal"]), "reflections temporal")
        mo_c, mo_o = f16(rf.download(4)), O.h2f(g["refl_moments"])
        assert np.array_equal(mo_c[..., 2], mo_o[..., 2]), "reflections history length differs from the golden fixture"
        close(mo_c, mo_o, "reflections moments")
        close(f16(rf.download(2)), O.h2f(g["refl_atrous"]), "reflections a-trous")
        close(f16(rf.download(100)), O.h2f(g["refl_final"]), "reflections final")
        rf.destroy()
        dd.destroy()
    finally:
        ctx.close()
