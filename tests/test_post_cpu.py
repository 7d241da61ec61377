"""Post-processing passes (TAA, tone map) on the CPU:
  * the oracle (oracle/orc_post.cpp) against closed-form properties and known answers computed in numpy from the reference's
    parsed literals (tests/golden/ref_constants.json);
  * the PRODUCT's per-pixel device functions (hybrid-rendering_b200/csrc/post_px.cuh — the code k_taa / k_tonemap execute) compiled
    for the host (tests/hostemu) against the oracle: bit-exact for TAA, exact for the tone map (same libm), so the kernels'
    arithmetic is checked without a GPU; tests/widened/test_gpu_taa_tonemap.py repeats the comparison through the C ABI on the device;
  * hr_taa_jitter (a pure host function of the product library) against the oracle and the Halton sequence's closed form."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import oracle as O
import pyhr

HERE = os.path.dirname(os.path.abspath(__file__))
REF = json.load(open(os.path.join(HERE, "golden", "ref_constants.json")))
_emu = None


def emu():
    global _emu
    if _emu is None:
        so = os.path.join(HERE, "hostemu", "_build", "libhostemu.so")
        if not os.path.exists(so):
            subprocess.run(["make", "-C", os.path.join(HERE, "hostemu")], check=True, stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        P, I, F = C.c_void_p, C.c_int, C.c_float
        L.emu_taa.argtypes = [I, I, P, I, P, P, P, P, F, F, I, P]
        L.emu_blit_rgba16f.argtypes = [I, I, P, I, P]
        L.emu_tonemap.argtypes = [I, I, P, I, F, I, P]
        L.emu_half_to_float.restype = F
        L.emu_half_to_float.argtypes = [C.c_uint16]
        L.emu_float_to_half.restype = C.c_uint16
        L.emu_float_to_half.argtypes = [F]
        _emu = L
    return _emu


def h(a):
    return np.asarray(a, np.float32).astype(np.float16).view(np.uint16)


def emu_taa(cur, prev, depth, gb2, jitter, fmin=0.88, fmax=0.97, sharpen=1):
    cur, prev, gb2 = (np.ascontiguousarray(a, np.uint16) for a in (cur, prev, gb2))
    depth = np.ascontiguousarray(depth, np.float32)
    H, W = cur.shape[:2]
    j = np.ascontiguousarray(jitter, np.float32)
    out = np.empty((H, W, 4), np.uint16)
    emu().emu_taa(W, H, O.p(cur), 1 if cur.ndim == 2 else cur.shape[2], O.p(prev), O.p(depth), O.p(gb2), O.p(j), fmin, fmax, int(sharpen), O.p(out))
    return out


def synthetic_inputs(W, H, seed, channels=4, motion_texels=2.0):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    base = 0.5 + 0.4 * np.sin(xx * 0.37 + seed) * np.cos(yy * 0.23)
    cur = np.stack([base * (0.5 + c * 0.3) + rng.random((H, W)) * 0.6 for c in range(4)], -1).astype(np.float32)
    cur[..., 3] = 1.0
    cur[H // 3, W // 2:] *= 6.0  # HDR highlights > 1
    prev = np.clip(cur + rng.normal(0, 0.2, cur.shape), 0, 1).astype(np.float32)
    depth = (0.9 + 0.1 * rng.random((H, W))).astype(np.float32)
    depth[:, W // 2] = 0.5  # a near edge: dilation picks its velocity in the neighbouring columns
    gb2 = np.zeros((H, W, 4), np.float32)
    gb2[..., 0:2] = rng.uniform(-1, 1, (H, W, 2))
    gb2[..., 2] = rng.uniform(-motion_texels, motion_texels, (H, W)) / W
    gb2[..., 3] = rng.uniform(-motion_texels, motion_texels, (H, W)) / H
    cur_h = h(cur if channels == 4 else (cur[..., 0] if channels == 1 else cur[..., :2]))
    return cur_h, h(prev), depth, h(gb2)


# ---------------------------------------------------------------------------------------------- exact half conversions of the host build
def test_hostemu_half_conversions_are_the_ieee_ones():
    L = emu()
    allh = np.arange(65536, dtype=np.uint16)
    want = allh.view(np.float16).astype(np.float32)
    got = np.array([L.emu_half_to_float(int(v)) for v in allh], np.float32)
    nan = np.isnan(want)
    assert np.array_equal(got[~nan].view(np.uint32), want[~nan].view(np.uint32)) and np.all(np.isnan(got[nan]))
    rng = np.random.default_rng(0)
    f = np.concatenate([rng.normal(0, 1, 20000), rng.uniform(-70000, 70000, 5000), 2.0 ** rng.uniform(-30, 17, 20000) * rng.choice([-1, 1], 20000),
                        want[~nan], np.nextafter(want[~nan], np.float32(np.inf)), np.nextafter(want[~nan], np.float32(-np.inf)),
                        (want[~nan][:-1].astype(np.float64) + want[~nan][1:]) .astype(np.float32) / 2]).astype(np.float32)
    f = f[np.isfinite(f)]
    with np.errstate(over="ignore"):
        want_h = f.astype(np.float16).view(np.uint16)
    got_h = np.array([L.emu_float_to_half(float(v)) for v in f], np.uint16)
    assert np.array_equal(got_h, want_h)


# ---------------------------------------------------------------------------------------------- Halton jitter
def _halton_exact(base, index):
    from fractions import Fraction
    r, f = Fraction(0), Fraction(1)
    while index > 0:
        f /= base
        r += f * (index % base)
        index //= base
    return r


def test_taa_jitter_is_the_halton_2_3_sequence():
    n = REF["taa"]["HALTON_SAMPLES"]["value"]
    assert n == 16
    for W, H in ((1920, 1080), (3840, 2160), (256, 144)):
        for frame in range(0, 40):
            i = frame % n + 1  # m_jitter_samples[k] holds sample k + 1 (temporal_aa.cpp:54-55)
            want = np.array([float(2 * _halton_exact(2, i) - 1) / W, float(2 * _halton_exact(3, i) - 1) / H])
            jo, jp = O.taa_jitter(frame, W, H), pyhr.taa_jitter(frame, W, H)
            assert np.array_equal(jo, jp), "product host function differs from the oracle"
            assert np.allclose(jo, want, rtol=1e-6, atol=1e-12)
    # the sample that makes bit-exact texel selection matter: Halton(2, 2) = 1/4 -> jitter.x = -0.5 / W exactly
    assert O.taa_jitter(1, 256, 144)[0] == np.float32(-0.5) / np.float32(256)


def test_apply_jitter_moves_the_image_by_the_jitter_in_ndc():
    """update_uniforms (main.cpp:941-957): projection' = translate(jitter) * projection.  A world point's NDC position moves by exactly the
    jitter; the inverse matrices stay inverses; the history matrix picks up the CURRENT jitter (not on the first frame)"""
    W, H = 256, 144
    f0 = pyhr.make_frame((0, 14, 34), (0, 3, 0), W, H)
    f1 = pyhr.make_frame((0.1, 14, 34), (0, 3, 0), W, H, prev=f0, num_frames=1)
    ref = pyhr.make_frame((0.1, 14, 34), (0, 3, 0), W, H, prev=f0, num_frames=1)
    j1, j0 = pyhr.taa_jitter(1, W, H), pyhr.taa_jitter(0, W, H)
    pyhr.apply_jitter(f1, j1, j0)

    def M(a):
        return np.array(a[:], np.float64).reshape(4, 4).T

    P = np.array([1.0, 2.0, -3.0, 1.0])
    a, b = M(ref.ubo.view_proj) @ P, M(f1.ubo.view_proj) @ P
    assert np.allclose(b[:2] / b[3] - a[:2] / a[3], j1, atol=1e-6) and np.isclose(b[2] / b[3], a[2] / a[3], atol=1e-6)
    assert np.allclose(M(f1.ubo.view_proj) @ M(f1.ubo.view_proj_inverse), np.eye(4), atol=2e-3)
    a, b = M(ref.ubo.prev_view_proj) @ P, M(f1.ubo.prev_view_proj) @ P
    assert np.allclose(b[:2] / b[3] - a[:2] / a[3], j1, atol=1e-6)
    assert np.allclose(f1.ubo.current_prev_jitter[:], [j1[0], j1[1], j0[0], j0[1]])
    # proj_inverse maps the jittered NDC back to the same view-space direction
    nd = np.array([0.3, -0.2, 1.0, 1.0])
    v0 = M(ref.ubo.proj_inverse) @ nd
    v1 = M(f1.ubo.proj_inverse) @ (nd + np.array([j1[0], j1[1], 0, 0]))
    assert np.allclose(v0[:3] / v0[3], v1[:3] / v1[3], rtol=1e-4, atol=1e-5)
    first = pyhr.make_frame((0, 14, 34), (0, 3, 0), W, H)
    keep = list(first.ubo.prev_view_proj[:])
    pyhr.apply_jitter(first, j0, (0.0, 0.0))
    assert list(first.ubo.prev_view_proj[:]) == keep  # main.cpp:955


# ---------------------------------------------------------------------------------------------- TAA
def test_taa_literals_equal_reference():
    t = REF["taa"]
    assert t["FLT_EPS"]["value"] == 1e-8
    assert sorted(t["defines"]["value"]) == sorted(["NUM_THREADS", "USE_DILATION", "MINMAX_3X3_ROUNDED", "USE_CLIPPING", "UNJITTER_REPROJECTION", "UNJITTER_COLORSAMPLES",
                                                    "UNJITTER_NEIGHBORHOOD", "HDR_CORRECTION"])  # no USE_YCOCG, no USE_OPTIMIZATIONS: the branches restated
    p = pyhr.hr_taa_params()
    pyhr.load_product().hr_taa_default_params(C.byref(p))
    d = REF["defaults"]["taa"]
    assert np.float32(p.feedback_min) == np.float32(d["m_feedback_min"]["value"]) and np.float32(p.feedback_max) == np.float32(d["m_feedback_max"]["value"])
    assert p.sharpen == int(d["m_sharpen"]["value"]) and p.reset_every_frame == int(d["m_reset"]["value"]) == 1


@pytest.mark.parametrize("W,H,seed,channels,sharpen", [(64, 36, 1, 4, 1), (61, 37, 2, 4, 0), (40, 24, 3, 1, 1), (33, 17, 4, 2, 1)])
def test_taa_device_functions_equal_the_oracle_bit_for_bit(W, H, seed, channels, sharpen):
    cur, prev, depth, gb2 = synthetic_inputs(W, H, seed, channels)
    for frame in (0, 1, 2, 5, 11):  # frame 1: jitter.x = -0.5 / W (texel borders), frame 0: jitter.x = 0
        j = O.taa_jitter(frame, W, H)
        a = O.taa(cur, prev, depth, gb2, j, sharpen=sharpen)
        b = emu_taa(cur, prev, depth, gb2, j, sharpen=sharpen)
        assert np.array_equal(a, b), f"frame {frame}: {np.count_nonzero(a != b)} halves differ"
    a = O.taa(cur, prev, depth, gb2, (0.0, 0.0), 0.5, 0.75, sharpen)
    assert np.array_equal(a, emu_taa(cur, prev, depth, gb2, (0.0, 0.0), 0.5, 0.75, sharpen))


def test_taa_device_functions_equal_the_oracle_on_arbitrary_finite_inputs():
    """hypothesis: ANY finite binary16 images (negative values, subnormals, +-65504, zeros), any depths in [0, 1], motion vectors of up to +-4
    texels, jitters of up to +-1.5 texels, any feedback pair in [0, 1]: the kernel's per-pixel function and the oracle agree on every bit"""
    from hypothesis import given, settings, strategies as st
    from hypothesis.extra import numpy as hnp
    W, H = 12, 9
    finite_half = st.integers(0, 0xFFFF).filter(lambda v: (v & 0x7C00) != 0x7C00)
    halves = lambda shape: hnp.arrays(np.uint16, shape, elements=finite_half)

    @settings(max_examples=60, deadline=None)
    @given(halves((H, W, 4)), halves((H, W, 4)), hnp.arrays(np.float32, (H, W), elements=st.floats(0.0, 1.0, width=32)),
           hnp.arrays(np.float32, (H, W, 2), elements=st.floats(-4.0, 4.0, width=32)), st.tuples(st.floats(-1.5, 1.5, width=32), st.floats(-1.5, 1.5, width=32)),
           st.floats(0.0, 1.0, width=32), st.floats(0.0, 1.0, width=32), st.integers(0, 1), st.sampled_from([1, 2, 4]))
    def check(cur, prev, depth, mv, jit, fmin, fmax, sharpen, channels):
        gb2 = np.zeros((H, W, 4), np.float32)
        gb2[..., 2], gb2[..., 3] = mv[..., 0] / W, mv[..., 1] / H
        c = cur if channels == 4 else (cur[..., 0].copy() if channels == 1 else cur[..., :2].copy())
        j = (np.float32(jit[0]) / np.float32(W), np.float32(jit[1]) / np.float32(H))
        a = O.taa(c, prev, depth, h(gb2), j, fmin, fmax, sharpen)
        b = emu_taa(c, prev, depth, h(gb2), j, fmin, fmax, sharpen)
        assert np.array_equal(a, b)
        out = a.view(np.float16).astype(np.float32)
        assert np.all(out[..., 3] == 1.0) and np.all((out[..., :3] >= 0.0) & (out[..., :3] <= 1.0))  # NaNs are clamped away too

    check()


def test_taa_closed_form_properties():
    W, H = 32, 16
    gb2 = np.zeros((H, W, 4), np.uint16)
    depth = np.full((H, W), 0.7, np.float32)
    # (1) a constant image with an equal history is a fixed point: every tap = c, clip box = {c}, sharpen 5c - 4c = c
    c = np.array([0.25, 0.5, 0.75, 1.0], np.float32)
    const = h(np.broadcast_to(c, (H, W, 4)))
    for j in ((0.0, 0.0), tuple(O.taa_jitter(1, W, H)), tuple(O.taa_jitter(7, W, H))):
        out = O.taa(const, const, depth, gb2, j).view(np.float16).astype(np.float32)
        assert np.allclose(out[..., :3], c[:3], atol=1e-3) and np.all(out[..., 3] == 1.0)
    # (2) the history is clipped into the neighbourhood box of the current frame: a wildly different history cannot leave it
    hist = h(np.broadcast_to(np.array([0.9, 0.0, 0.1, 1.0], np.float32), (H, W, 4)))
    out = O.taa(const, hist, depth, gb2, (0.0, 0.0)).view(np.float16).astype(np.float32)
    assert np.allclose(out[..., :3], c[:3], atol=2e-3)
    # (3) output is clamped to [0, 1] (imageStore of clamp(to_buffer, 0, 1)) even for HDR input
    bright = h(np.broadcast_to(np.array([8.0, 4.0, 0.5, 1.0], np.float32), (H, W, 4)))
    out = O.taa(bright, bright, depth, gb2, (0.0, 0.0)).view(np.float16).astype(np.float32)
    assert np.allclose(out[..., 0], 1.0) and np.allclose(out[..., 1], 1.0) and np.allclose(out[..., 2], 0.5, atol=1e-3)
    # (4) dilation: the velocity comes from the nearest-depth texel of the 3x3 neighbourhood; with a FLAT depth the strict '>' never
    # fires and the first tap wins — the TOP-LEFT neighbour (dmin = dtl, taa.comp:181).  History = a horizontal ramp, feedback 1 (output =
    # clipped history), only column 10 carries a motion vector (4 texels to the right):
    #   flat depth     column x reads the velocity of column x - 1: only column 11 moves
    #   column 10 near columns 9, 10, 11 all read column 10's velocity: 9 and 10 move as well
    ramp = np.zeros((H, W, 4), np.float32)
    ramp[..., 0] = np.linspace(0.2, 0.8, W)[None, :]
    ramp[..., 3] = 1.0
    d2 = depth.copy()
    d2[:, 10] = 0.1
    g = np.zeros((H, W, 4), np.float32)
    g[:, 10, 2] = 4.0 / W
    still = O.taa(h(ramp), h(ramp), depth, np.zeros((H, W, 4), np.uint16), (0.0, 0.0), 1.0, 1.0, 0).view(np.float16).astype(np.float32)
    flat = O.taa(h(ramp), h(ramp), depth, h(g), (0.0, 0.0), 1.0, 1.0, 0).view(np.float16).astype(np.float32)
    near = O.taa(h(ramp), h(ramp), d2, h(g), (0.0, 0.0), 1.0, 1.0, 0).view(np.float16).astype(np.float32)
    moved_flat = np.any(flat[..., 0] != still[..., 0], axis=0)
    moved_near = np.any(near[..., 0] != still[..., 0], axis=0)
    assert np.flatnonzero(moved_flat).tolist() == [11] and np.flatnonzero(moved_near).tolist() == [9, 10, 11]
    assert np.all(near[:, 9:12, 0] > still[:, 9:12, 0])  # the history tap moved up the ramp (and was clipped to the neighbourhood's maximum)


def test_taa_host_sequencing_reset_quirk():
    """m_reset is never cleared in the reference (temporal_aa.h:57, temporal_aa.cpp:112): as written, every frame blits the current
    input over the history, so two consecutive frames with the same input give the same output whatever happened before;
    with the flag cleared after the first frame the history accumulates"""
    W, H = 48, 32
    cur, _, depth, gb2 = synthetic_inputs(W, H, 9, motion_texels=0.0)
    cur2 = synthetic_inputs(W, H, 10, motion_texels=0.0)[0]
    f0 = pyhr.make_frame((0, 14, 34), (0, 3, 0), W, H, prev=None, num_frames=0)
    f1 = pyhr.make_frame((0, 14, 34), (0, 3, 0), W, H, prev=f0, num_frames=1)
    f2 = pyhr.make_frame((0, 14, 34), (0, 3, 0), W, H, prev=f1, num_frames=2)
    as_written, intended = O.TAAOracle(W, H), O.TAAOracle(W, H, reset_every_frame=0)
    outs = [[t.render(f, c, depth, gb2).copy() for f, c in ((f0, cur), (f1, cur2), (f2, cur2))] for t in (as_written, intended)]
    assert np.array_equal(outs[0][0], outs[1][0])          # first frame: both blit
    assert np.array_equal(outs[0][1], outs[0][2])          # as written: frame 2 == frame 1 (same input, history overwritten)
    assert not np.array_equal(outs[1][1], outs[1][2])      # intended: the history moves the result
    assert not np.array_equal(outs[0][1], outs[1][1])


def test_blit_component_fill():
    rng = np.random.default_rng(1)
    for ch in (1, 2, 4):
        src = h(rng.random((9, 7) if ch == 1 else (9, 7, ch)))
        a = O.blit_rgba16f(src)
        b = np.empty_like(a)
        emu().emu_blit_rgba16f(7, 9, O.p(src), ch, O.p(b))
        assert np.array_equal(a, b)
        s3 = src.reshape(9, 7, ch)
        assert np.array_equal(a[..., :ch], s3) and np.all(a[..., 3] == (0x3C00 if ch < 4 else s3[..., 3])) and (ch > 2 or np.all(a[..., 2] == 0))


# ---------------------------------------------------------------------------------------------- tone map
def test_tonemap_known_answer_from_reference_literals():
    t = REF["tone_map"]
    gamma = t["gamma_exponent"]["value"]
    assert abs(gamma - 1 / 2.2) < 1e-12
    rng = np.random.default_rng(4)
    src = np.concatenate([rng.random((16, 32, 4)) * 4.0, np.zeros((1, 32, 4)), np.full((1, 32, 4), 100.0)]).astype(np.float32)
    for exposure in (1.0, 0.25, 3.0):
        x = src.astype(np.float16).astype(np.float64)[..., :3] * exposure
        want = np.clip((x * (t["aces"]["value"]["a"] * x + t["aces"]["value"]["b"])) / (x * (t["aces"]["value"]["c"] * x + t["aces"]["value"]["d"]) + t["aces"]["value"]["e"]), 0, 1) ** gamma
        got = O.tonemap(h(src), exposure)
        assert np.all(got[..., 3] == 255)
        assert np.max(np.abs(got[..., :3].astype(np.float64) - want * 255.0)) <= 0.5 + 1e-3  # correctly rounded UNORM8 of the fp64 value, up to fp32 noise
        e32 = np.empty((src.shape[0], src.shape[1]), np.uint32)
        emu().emu_tonemap(src.shape[1], src.shape[0], O.p(h(src)), 4, exposure, 0, O.p(e32))
        assert np.array_equal(e32.view(np.uint8).reshape(got.shape), got)
    # single_channel: .rrr without curve (tone_map.frag:54-55)
    g1 = h(rng.random((5, 6)))
    got = O.tonemap(g1, 1.0, 1)
    want = np.rint(g1.view(np.float16).astype(np.float32) * 255.0).astype(np.uint8)
    assert np.array_equal(got[..., 0], want) and np.array_equal(got[..., 1], want) and np.array_equal(got[..., 2], want)
    p = pyhr.hr_tonemap_params()
    pyhr.load_product().hr_tonemap_default_params(C.byref(p))
    assert p.exposure == REF["defaults"]["tone_map"]["m_exposure"]["value"] and p.single_channel == 0
