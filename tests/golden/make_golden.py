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
