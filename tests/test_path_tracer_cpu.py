"""Ground-truth path tracer on the CPU: the oracle (oracle/orc_path_trace.cpp) against
  * an independent numpy statement of the ray generation (ground_truth_path_trace.rgen:56-75) driven by the RNG restated from the
    reference's parsed constants (tests/test_ref_constants.py::_ref_rng_floats): the primitives hit must be identical;
  * closed-form properties of the accumulation the reference writes (rgen:94-111) and of the lighting (sky pixels, black light,
    energy bounds)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle as O
import pyhr
from test_ref_constants import REF, _ref_rng_floats

W, H = 64, 36
SKY = (0.3, 0.4, 0.6)


def f16(a):
    return np.ascontiguousarray(a).view(np.float16).astype(np.float32)


@pytest.fixture(scope="module")
def scene():
    sc = pyhr.SynthScene(pyhr.SCENE_SHADOWS_TEST)
    return sc, O.ShadingScene(sc, brute=True)


def test_literals_equal_reference():
    pt = REF["path_tracer"]
    assert pt["primary_tmin_tmax"]["value"] == [0.001, 10000.0] and pt["RADIANCE_CLAMP_COLOR"]["value"] == 1.0
    assert pt["shadow_ray_origin_offset"]["value"] == 0.1 and pt["query_distance_t_min"]["value"] == 0.01
    assert pt["indirect_trace_is_commented_out"]["value"] is True  # the restatement's "indirect = 0" rests on this
    assert pt["rchit_defines"]["value"] == ["RAY_TRACING", "SOFT_SHADOWS", "RAY_THROUGHPUT", "SAMPLE_SKY_LIGHT"]
    p = pyhr.hr_path_tracer_params()
    pyhr.load_product().hr_path_tracer_default_params(C.byref(p))
    assert p.max_ray_bounces == REF["defaults"]["path_tracer"]["max_ray_bounces"]["value"] and p.roughness_multiplier == 1.0


def _numpy_primary_rays(f, frame_idx):
    """rgen:56-75 in numpy float32, same operation order (mat4 * vec4 rows as ((m0 x + m1 y) + m2 z) + m3 w)"""
    vi = np.array(f.ubo.view_inverse[:], np.float32)
    pi = np.array(f.ubo.proj_inverse[:], np.float32)

    def mv(M, x, y, z, w):
        return [((M[r] * x + M[4 + r] * y) + M[8 + r] * z) + M[12 + r] * w for r in range(4)]

    rays = np.zeros((H, W, 8), np.float32)
    one, zero = np.float32(1), np.float32(0)
    for y in range(H):
        for x in range(W):
            j = _ref_rng_floats(x, y, frame_idx, 2)
            jx, jy = (np.float32(x) + np.float32(0.5)) + j[0], (np.float32(y) + np.float32(0.5)) + j[1]
            nx, ny = (jx / np.float32(W)) * np.float32(2) - one, (jy / np.float32(H)) * np.float32(2) - one
            o = mv(vi, zero, zero, zero, one)
            t = mv(pi, nx, ny, one, one)
            inv = one / np.sqrt((t[0] * t[0] + t[1] * t[1]) + t[2] * t[2])
            d = mv(vi, t[0] * inv, t[1] * inv, t[2] * inv, zero)
            rays[y, x] = [o[0], o[1], o[2], 0.001, d[0], d[1], d[2], 10000.0]
    return rays


def test_primary_rays_hit_the_primitives_an_independent_numpy_ray_generator_hits(scene):
    sc, ss = scene
    f = pyhr.make_frame((0, 14, 34), (0, 3, 0), W, H)
    for frame_idx in (0, 3):
        o = O.PathTracerOracle(W, H, sky=SKY)
        o.frame_idx = frame_idx
        o.render(ss, f)
        t, prim, _ = ss.scene.trace_closest(_numpy_primary_rays(f, frame_idx).reshape(-1, 8))
        assert np.array_equal(o.prim.reshape(-1), prim)
        assert np.count_nonzero(prim != 0xFFFFFFFF) > 0.3 * W * H and np.count_nonzero(prim == 0xFFFFFFFF) > 0


def test_accumulation_as_the_reference_writes_it(scene):
    """o_0 = c_0; o_n = o_{n-1} + (c_n - o_{n-1}) / n: frame 1 REPLACES frame 0 (weight 1 / 1), after that a running mean of samples 1..n"""
    sc, ss = scene
    f = pyhr.make_frame((0, 14, 34), (0, 3, 0), W, H)
    samples = []
    for n in range(5):  # the individual samples c_n: render each frame index as if it were the first of a fresh accumulation image
        o = O.PathTracerOracle(W, H, sky=SKY)
        o.render(ss, f)           # consumes frame index 0
        if n:
            o.restart_accumulation()
            o2 = O.PathTracerOracle(W, H, sky=SKY)
            o2.frame_idx = n
            o2.img[0][:] = 0      # prev = 0, n = frame index: out = c_n / n
            samples.append(f16(o2.render(ss, f))[..., :3] * n)
        else:
            samples.append(f16(o.final)[..., :3])
    acc = O.PathTracerOracle(W, H, sky=SKY)
    outs = [f16(acc.render(ss, f).copy())[..., :3] for _ in range(5)]
    assert np.array_equal(outs[0], samples[0])
    assert np.allclose(outs[1], samples[1], atol=2e-3)                          # frame 0 is forgotten
    assert np.allclose(outs[4], np.mean(samples[1:5], axis=0), atol=6e-3)       # mean of samples 1..4 (fp16 storage between frames)
    assert not np.allclose(outs[4], np.mean(samples[0:5], axis=0), atol=1e-3)   # ... not of 0..4
    acc.restart_accumulation()
    assert np.array_equal(f16(acc.render(ss, f))[..., :3], samples[0])          # restart_accumulation(): frame index 0 again, image replaced


def test_lighting_properties(scene):
    sc, ss = scene
    f = pyhr.make_frame((0, 14, 34), (0, 3, 0), W, H)
    o = O.PathTracerOracle(W, H, sky=SKY)
    img = f16(o.render(ss, f))
    sky_px = o.prim == 0xFFFFFFFF
    assert np.all(img[..., 3] == 1.0) and np.isfinite(img).all()
    assert np.allclose(img[sky_px][:, :3], np.float16(SKY).astype(np.float32))   # rmiss at depth 0: L = the environment sample
    assert img[..., :3].max() <= 1.0 and img[..., :3].min() >= 0.0               # RADIANCE_CLAMP_COLOR
    # black sky + light switched off: every surface pixel is black; light on: lit pixels exist and some are in shadow
    dark = pyhr.make_frame((0, 14, 34), (0, 3, 0), W, H, light=pyhr.default_light(intensity=0.0))
    o0 = O.PathTracerOracle(W, H)
    assert f16(o0.render(ss, dark))[..., :3].max() == 0.0
    o1 = O.PathTracerOracle(W, H)
    lit = f16(o1.render(ss, f))[..., :3]
    surf = ~sky_px
    assert lit[surf].max() > 0.05 and np.mean(lit[surf].sum(-1) == 0.0) > 0.02
    # the sky term adds energy on top of the punctual light (same RNG draws, same shadow rays)
    assert np.all(img[..., :3][surf] >= lit[surf] - 1e-3) and img[..., :3][surf].mean() > lit[surf].mean()
    # roughness_multiplier reaches the BRDF (rchit:122)
    o2 = O.PathTracerOracle(W, H, roughness_multiplier=0.5)
    assert not np.array_equal(f16(o2.render(ss, f))[..., :3], lit)
