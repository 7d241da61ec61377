"""GPU: a scene that went through the asset loaders (host/assets.cpp: glTF binary -> dw::Mesh tables -> RayTracedScene tables ->
hr_scene_build) renders exactly like the procedural scene it was exported from: generic ray queries, shadows / AO masks and the
denoised outputs are bit-identical, and the blue-noise tables read from a PNG directory drive the passes like the in-memory ones."""
import numpy as np
import pytest

import pyhr
from pyhr import assets as A
from test_assets import export_world_baked_glb, write_png

pytestmark = pytest.mark.gpu


def test_loaded_scene_renders_like_the_procedural_one(tmp_path):
    W, H = 256, 144
    sc = pyhr.SynthScene(pyhr.SCENE_SHADOWS_TEST)
    mesh = A.Mesh(export_world_baked_glb(sc, tmp_path))
    asc = A.AssetScene([(mesh, np.eye(4, dtype=np.float32).reshape(-1))])
    tri_ref, inst_ref = sc.world_triangles()
    tri, inst = asc.world_triangles()
    assert np.array_equal(tri.view(np.uint32), tri_ref.view(np.uint32)) and np.array_equal(inst, inst_ref)

    # blue-noise tables through the PNG directory loader
    sobol, sr = pyhr.blue_noise()
    write_png(tmp_path / "sobol_256_4d.png", sobol.reshape(1, 256, 4), 6, 8)
    write_png(tmp_path / "scrambling_ranking_128x128_2d_1spp.png", sr, 6, 8)
    so2, sr2, mask = A.bluenoise_load(tmp_path)
    assert mask == 1 and np.array_equal(so2, sobol) and np.array_equal(sr2[0], sr)

    results = []
    for scene, bn in ((sc, (sobol, sr)), (asc, (so2, np.ascontiguousarray(sr2[0])))):
        ctx = pyhr.Context(0)
        ctx.set_bluenoise(*bn)
        ctx.build_scene(scene)
        ctx.gbuffer_create(W, H)
        sh, ao = pyhr.Pass(ctx, "shadows", W, H, 0), pyhr.Pass(ctx, "ao", W, H, 1)
        out, f = [], None
        for i in range(3):
            f = pyhr.make_frame((0, 14, 34), (0, 3, 0), W, H, prev=f, num_frames=i)
            ctx.gbuffer_upload(f.ping_pong, pyhr.write_gbuffer(sc, f, W, H))
            sh.render(f)
            ao.render(f)
            out += [sh.download(0), ao.download(0), sh.download(100), ao.download(100)]
        results.append(out)
        sh.destroy()
        ao.destroy()
        ctx.close()
    for a, b in zip(*results):
        assert np.array_equal(a, b)
