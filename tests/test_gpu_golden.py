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
