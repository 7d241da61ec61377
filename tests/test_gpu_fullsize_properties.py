"""(Added late in the round: everything up to the history-length check ran green on the box; that check then failed because the
test — not the library — fed the G-buffer of frame 0, whose motion vectors are meaningless; fixed, not re-run.)
BASELINE.json's full size (3840x2160, the 262 144-triangle arcade of bench.py) through size-independent properties — the
oracle needs minutes per 4K frame, so nothing here runs it:

* determinism: the same three frames rendered by two fresh contexts give bit-identical images (masks AND fp16 outputs);
* the visibility masks do not depend on the BVH topology (PLOC vs the Karras radix tree) — bit-exact;
* visibility / AO stay in [0, 1], sky pixels carry 0 shadow visibility and AO 1, every mask bit of a sky pixel is 0;
* with a static camera the history length is min(32, frames rendered) on (practically) every surface pixel — the
  reprojection of a pixel onto itself must be accepted;
* the a-trous filter is an averaging filter: the denoised visibility lies within the range of the temporal output.
"""
import numpy as np
import pytest

import pyhr

pytestmark = pytest.mark.gpu

W, H, TRIS = 3840, 2160, 262144
CAM_POS, CAM_TGT, LIGHT_ROT_X = (0.0, 9.0, -4.0), (2.0, 7.0, 60.0), 25.0
N_FRAMES = 3


def f16(a):
    return np.ascontiguousarray(a).view(np.float16).astype(np.float32)


@pytest.fixture(scope="module")
def inputs():
    sc = pyhr.SynthScene(pyhr.SCENE_ARCADE, TRIS)
    light = pyhr.default_light(rot_x_deg=LIGHT_ROT_X)
    frames, f = [], None
    for i in range(N_FRAMES):
        f = pyhr.make_frame(CAM_POS, CAM_TGT, W, H, prev=f, num_frames=i, light=light)
        frames.append(f)
    # static camera: one G-buffer serves every frame; written from frame 1, whose previous-frame matrices equal the current ones
    # (zero motion vectors) — frame 0 has no previous frame and would encode garbage motion
    g = pyhr.write_gbuffer(sc, frames[1], W, H)
    return sc, frames, g


def render(inputs, bvh_quality=1):
    sc, frames, g = inputs
    ctx = pyhr.Context(0)
    try:
        ctx.lib.hr_debug_set(3, bvh_quality)
        ctx.set_bluenoise(*pyhr.blue_noise())
        ctx.build_scene(sc)
        ctx.gbuffer_create(W, H)
        sh, ao = pyhr.Pass(ctx, "shadows", W, H, 0), pyhr.Pass(ctx, "ao", W, H, 1)
        for f in frames:
            ctx.gbuffer_upload(f.ping_pong, g)
            sh.render(f)
            ao.render(f)
        out = dict(sh_mask=sh.download(0).copy(), sh_temporal=sh.download(1).copy(), sh_moments=sh.download(4).copy(), sh_final=sh.download(100).copy(),
                   ao_mask=ao.download(0).copy(), ao_length=ao.download(4).copy(), ao_final=ao.download(100).copy())
        sh.destroy()
        ao.destroy()
    finally:
        ctx.lib.hr_debug_set(3, 1)
        ctx.close()
    return out


@pytest.fixture(scope="module")
def first_run(inputs):
    return render(inputs)


def test_4k_properties(inputs, first_run):
    sc, frames, g = inputs
    a = first_run
    b = render(inputs)
    for k in a:
        assert np.array_equal(a[k], b[k]), f"{k}: two runs of the same frames differ (non-deterministic kernel)"
    c = render(inputs, bvh_quality=0)
    assert np.array_equal(a["sh_mask"], c["sh_mask"]) and np.array_equal(a["ao_mask"], c["ao_mask"]), "visibility mask depends on the BVH topology"
    assert np.array_equal(a["sh_final"], c["sh_final"]) and np.array_equal(a["ao_final"], c["ao_final"])

    sky = g.depth == 1.0
    surf = ~sky
    assert 0.02 < sky.mean() < 0.98, "the test view should contain both sky and geometry"
    vis, ao = f16(a["sh_final"])[..., 0], f16(a["ao_final"])
    assert np.isfinite(vis).all() and np.isfinite(ao).all()
    assert vis.min() >= 0.0 and vis.max() <= 1.0 + 1e-3 and ao.min() >= 0.0 and ao.max() <= 1.0 + 1e-3
    assert (vis[sky] == 0.0).all() and (ao[sky] == 1.0).all()
    assert 0.005 < vis[surf].mean() < 0.995, "shadows should be neither all lit nor all dark"
    # mask bit (y&3)*8 + (x&7) of word (x>>3, y>>2): sky pixels never set a bit
    m = a["sh_mask"]
    yy, xx = np.nonzero(sky[: (H // 4) * 4, : (W // 8) * 8])
    bits = (m[yy >> 2, xx >> 3] >> ((yy & 3) * 8 + (xx & 7)).astype(np.uint32)) & 1
    assert not bits.any(), "a sky pixel has its shadow-mask bit set"
    assert a["sh_mask"].any() and a["ao_mask"].any()


def test_4k_static_camera_history(inputs, first_run):
    """static camera: history accepted everywhere => history length = frames rendered; the filter averages."""
    sc, frames, g = inputs
    a = first_run
    sky = g.depth == 1.0
    surf = ~sky
    hl = f16(a["sh_moments"])[..., 2]
    assert (hl[sky] == 0.0).all()
    assert (hl[surf] == float(N_FRAMES)).mean() > 0.99, "reprojection of a static pixel onto itself was rejected"
    al = f16(a["ao_length"])
    assert set(np.unique(al)).issubset({0.0, 1.0, 2.0, float(N_FRAMES)})
    vis, t = f16(a["sh_final"])[..., 0], f16(a["sh_temporal"])[..., 0]
    assert vis[surf].min() >= t.min() - 1e-3 and vis[surf].max() <= max(t.max(), 1.0) + 1e-3
