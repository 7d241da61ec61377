"""The headless C++ frame loop with the post chain (build/hr_headless --post): Halton-jittered projection (hr::TemporalAA::update /
apply_jitter) -> device G-buffer -> four passes -> deferred combine + sky box -> TAA -> tone map (RGBA8), on the host classes of
hybrid-rendering_b200/host/hybrid_rendering.h."""
import os
import subprocess

import pytest

import pyhr

pytestmark = pytest.mark.gpu


def test_hr_headless_post_chain(tmp_path):
    exe = os.path.join(pyhr.BUILD_DIR, "hr_headless")
    assert os.path.exists(exe), "build/hr_headless missing (make -C hybrid-rendering_b200)"
    png = tmp_path / "frame.png"
    r = subprocess.run([exe, "--post", "--png", str(png), "256", "144", "4", "5000"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    assert "frames=4" in r.stdout and "finite 1" in r.stdout and "output 256x144 fmt 4" in r.stdout, r.stdout
    assert "tone-mapped TAA output 256x144 fmt 6" in r.stdout and "alpha opaque 1" in r.stdout, r.stdout
    # the PNG it wrote is the tone-mapped frame: our decoder reads it back, opaque and not black
    from pyhr import assets as A
    img = A.image_load(png)
    assert img.shape == (144, 256, 4) and (img[..., 3] == 255).all() and img[..., :3].mean() > 1.0
