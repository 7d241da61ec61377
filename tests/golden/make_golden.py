"""Regenerates tests/golden/shadows_ao_64x48.npz and tests/golden/shadows_ao_256x144_seq12.npz from the CPU oracle (run from the
repo root: python tests/golden/make_golden.py).
The reference itself cannot be run (Vulkan RT + GLSL, no tests/golden vectors of its own — SURVEY.md §8c), so these
vectors pin the oracle, not the reference."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from test_cpu_oracle import small_sequence  # noqa: E402

if __name__ == "__main__":
    sh, ao = small_sequence()
    np.savez_compressed(os.path.join(HERE, "shadows_ao_64x48.npz"), sh_mask=sh.mask, ao_mask=ao.mask, sh_temporal=sh.temporal, sh_moments=sh.cur_moments,
                    sh_final=sh.final, sh_prev_image=sh.prev_image, sh_tiles=sh.tile_flags, ao_temporal=ao.temporal, ao_blur=ao.blur[1],
                    ao_final=ao.final, ao_tiles=ao.tile_flags)
    print("wrote golden")


# ---- 256x144, 8 static + 4 panning frames (the sequence of tests/test_gpu_parity.py::test_shadows_ao_static_then_pan):
# the state after the last frame; tests/test_gpu_golden.py drives the same inputs through the C ABI and compares with it.
import oracle as O  # noqa: E402
import pyhr  # noqa: E402


def seq12_frames(W=256, H=144, n=12, pan_from=8):
    f = None
    for i in range(n):
        dx = 0.0 if i < pan_from else 0.05 * (i - pan_from + 1)
        f = pyhr.make_frame((dx, 14.0, 34.0), (dx, 3.0, 0.0), W, H, prev=f, num_frames=i)
        yield f


def seq12_oracle(W=256, H=144):
    sc = pyhr.SynthScene(pyhr.SCENE_SHADOWS_TEST)
    tri, _ = sc.world_triangles()
    osc = O.Scene(tri, brute=True)
    bn = pyhr.blue_noise()
    sh, ao = O.ShadowsOracle(W, H, 0), O.AOOracle(W, H, 1)
    prev = O.zero_gbuf_mips(W, H)
    for f in seq12_frames(W, H):
        cur = O.GBufMips(pyhr.write_gbuffer(sc, f, W, H))
        sh.render(osc, cur, prev, f, bn)
        ao.render(osc, cur, prev, f, bn)
        prev = cur
    return sh, ao


if __name__ == "__main__":
    sh, ao = seq12_oracle()
    np.savez_compressed(os.path.join(HERE, "shadows_ao_256x144_seq12.npz"), sh_mask=sh.mask, sh_tiles=sh.tile_flags, sh_temporal=sh.temporal,
                        sh_moments=sh.cur_moments, sh_atrous=sh.atrous_out, sh_prev_image=sh.prev_image, sh_final=sh.final, ao_mask=ao.mask,
                        ao_tiles=ao.tile_flags, ao_temporal=ao.temporal, ao_length=ao.cur_length, ao_blur=ao.blur[1], ao_final=ao.final)
    print("wrote golden 256x144 seq12")


# ---- 192x112 DDGI + half-res reflections, 3 static + 2 panning frames (tests/test_gpu_gi_refl.py::test_ddgi_and_reflections_
# static_then_pan): screen-space images after the last frame (the probe atlases / ray buffers are too big for a fixture).
GI_W, GI_H, GI_SKY = 192, 112, (0.3, 0.4, 0.6)


def gi_frames(n=5, pan_from=3):
    f = None
    for i in range(n):
        dx = 0.0 if i < pan_from else 0.05 * (i - pan_from + 1)
        f = pyhr.make_frame((dx, 14.0, 34.0), (dx, 3.0, 0.0), GI_W, GI_H, prev=f, num_frames=i)
        yield i, f, pyhr.rotation_matrix(0.7 + 1.3 * i, (0.3, 1.0, -0.5))


def gi_params(dd_params, rf_params):
    for P in (dd_params, rf_params):
        P.sky_color[0], P.sky_color[1], P.sky_color[2] = GI_SKY
    dd_params.probe_distance = 4.0
    dd_params.normal_bias = 1.0


def gi_oracle():
    sc = pyhr.SynthScene(pyhr.SCENE_SHADOWS_TEST)
    ss = O.ShadingScene(sc, brute=True)
    bn = pyhr.blue_noise()
    mn, mx = sc.bounds()
    ddp, rfp = pyhr.hr_ddgi_params(), pyhr.hr_reflections_params()
    pyhr.load_product().hr_ddgi_default_params(C.byref(ddp))
    pyhr.load_product().hr_reflections_default_params(C.byref(rfp))
    gi_params(ddp, rfp)
    odd = O.DDGIOracle(GI_W, GI_H, 0, ddp, mn, mx)
    orf = O.ReflectionsOracle(GI_W, GI_H, 1, rfp)
    prev = O.zero_gbuf_mips(GI_W, GI_H)
    for i, f, rot in gi_frames():
        cur = O.GBufMips(pyhr.write_gbuffer(sc, f, GI_W, GI_H))
        odd.render(ss, cur, f, rot)
        orf.render(ss, cur, prev, f, bn, odd)
        prev = cur
    return odd, orf


import ctypes as C  # noqa: E402

if __name__ == "__main__":
    odd, orf = gi_oracle()
    np.savez_compressed(os.path.join(HERE, "ddgi_reflections_192x112_seq5.npz"), ddgi_sample=odd.sample, refl_rt=orf.rt, refl_tiles=orf.tile_flags,
                        refl_temporal=orf.cur_temporal, refl_moments=orf.cur_moments, refl_atrous=orf.atrous_out, refl_final=orf.final)
    print("wrote golden ddgi + reflections")
