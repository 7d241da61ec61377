"""Regenerates tests/golden/shadows_ao_64x48.npz from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).
The reference itself cannot be run (Vulkan RT + GLSL, no tests/golden vectors of its own — SURVEY.md §8c), so these
vectors pin the oracle, not the reference."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from test_cpu_oracle import small_sequence  # noqa: E402

sh, ao = small_sequence()
np.savez_compressed(os.path.join(HERE, "shadows_ao_64x48.npz"), sh_mask=sh.mask, ao_mask=ao.mask, sh_temporal=sh.temporal, sh_moments=sh.cur_moments,
                    sh_final=sh.final, sh_prev_image=sh.prev_image, sh_tiles=sh.tile_flags, ao_temporal=ao.temporal, ao_blur=ao.blur[1],
                    ao_final=ao.final, ao_tiles=ao.tile_flags)
print("wrote golden")
