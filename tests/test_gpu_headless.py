"""The C++ host classes (hybrid-rendering_b200/host/hybrid_rendering.h, the reference-named pass interface) driven by the headless
frame loop build/hr_headless: device G-buffer -> shadows -> AO -> DDGI -> reflections -> deferred combine, several frames
(the --post variant with the Halton jitter, TAA and the tone map: tests/widened/test_gpu_whole_frame_post.py)."""
import os
import subprocess

import pytest

import pyhr

pytestmark = pytest.mark.gpu


def test_hr_headless_runs_the_whole_frame():
    exe = os.path.join(pyhr.BUILD_DIR, "hr_headless")
    assert os.path.exists(exe), "build/hr_headless missing (make -C hybrid-rendering_b200)"
    r = subprocess.run([exe, "256", "144", "4", "5000"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    assert "frames=4" in r.stdout and "finite 1" in r.stdout and "output 256x144 fmt 4" in r.stdout, r.stdout
