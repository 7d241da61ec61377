"""CPU tests (no GPU): the oracle against closed-form properties, its own brute force, and the committed golden vectors.

The reference ships no tests / golden vectors for this path ("parity unpinned"); the goldens under
tests/golden/ are outputs of this oracle (script: tests/golden/make_golden.py) and pin it against regressions.
"""
import os
import sys

import numpy as np
import pytest

import oracle as O
import pyhr

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_det_sincos_accuracy_and_quadrants():
    x = np.linspace(0.0, 2 * np.pi, 20001).astype(np.float32)
    s, c = np.empty_like(x), np.empty_like(x)
    O.lib().orc_det_sincos(O.p(x), x.size, O.p(s), O.p(c))
    assert np.abs(s - np.sin(x.astype(np.float64))).max() < 3e-7
    assert np.abs(c - np.cos(x.astype(np.float64))).max() < 3e-7
    # known answers at the quadrant points
    q = np.array([0.0, np.pi / 2, np.pi, 3 * np.pi / 2], np.float32)
    s4, c4 = np.empty_like(q), np.empty_like(q)
    O.lib().orc_det_sincos(O.p(q), 4, O.p(s4), O.p(c4))
    assert np.allclose(s4, [0, 1, 0, -1], atol=2e-7) and np.allclose(c4, [1, 0, -1, 0], atol=2e-7)


def oct_encode(n):
    p = n[:, :2] / np.abs(n).sum(1, keepdims=True)
    neg = n[:, 2] <= 0
    q = (1.0 - np.abs(p[:, ::-1])) * np.where(p >= 0, 1.0, -1.0)
    return np.where(neg[:, None], q, p)


def test_octahedral_round_trip():
    rng = np.random.default_rng(1)
    n = rng.normal(size=(5000, 3))
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    e = oct_encode(n).astype(np.float32)
    out = np.empty((5000, 3), np.float32)
    O.lib().orc_oct_decode(O.p(np.ascontiguousarray(e)), 5000, O.p(out))
    assert np.abs(out - n).max() < 1e-6
    assert np.abs(np.linalg.norm(out, axis=1) - 1).max() < 1e-6


def test_blue_noise_sampler_integer_path():
    sobol, sr = pyhr.blue_noise(1234)
    # dimension 0 of Sobol' is the van der Corput sequence: byte i = bit-reverse(i)
    assert all(int(sobol[i, 0]) == int(f"{i:08b}"[::-1], 2) for i in range(256))
    # each Sobol' dimension is a permutation of 0..255 (a (0,8,1)-net in base 2)
    for d in range(4):
        assert sorted(sobol[:, d].tolist()) == list(range(256))
    x, y, idx = 37, 101, 300
    for dim in range(4):
        rank = int(sr[y % 128, x % 128, 2])
        v = int(sobol[(idx % 256) ^ rank, dim]) ^ int(sr[y % 128, x % 128, dim % 2])
        got = O.lib().orc_sample_blue_noise(x, y, idx, dim, O.p(sobol), O.p(sr))
        assert got == np.float32((0.5 + v) / 256.0)
    # wrap-around of coordinates and index (bnd_sampler.glsl:7-10)
    assert O.lib().orc_sample_blue_noise(5 + 128, 9 + 256, 3 + 512, 1 + 4, O.p(sobol), O.p(sr)) == O.lib().orc_sample_blue_noise(5, 9, 3, 1, O.p(sobol), O.p(sr))


def random_rays(sc, n, seed):
    rng = np.random.default_rng(seed)
    mn, mx = sc.bounds()
    o = rng.uniform(mn - 1, mx + 5, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[: n // 8, 1] = 0.0
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3], rays[:, 3], rays[:, 4:7], rays[:, 7] = o, 0.01, d, rng.choice([7.0, 1e4], n)
    return rays


def test_reference_bvh_equals_bruteforce():
    sc = pyhr.SynthScene(pyhr.SCENE_SHADOWS_TEST)
    tri, _ = sc.world_triangles()
    rays = random_rays(sc, 4000, 3)
    fast, brute = O.Scene(tri, brute=False), O.Scene(tri, brute=True)
    assert np.array_equal(fast.trace_any(rays), brute.trace_any(rays))
    tf, pf, uf = fast.trace_closest(rays)
    tb, pb, ub = brute.trace_closest(rays)
    assert np.array_equal(pf, pb) and np.array_equal(tf, tb) and np.array_equal(uf, ub)
    assert 0.05 < brute.trace_any(rays).mean() < 0.95


def test_single_triangle_known_answers():
    sc = pyhr.SynthScene(pyhr.SCENE_SINGLE_TRIANGLE)
    tri, _ = sc.world_triangles()
    s = O.Scene(tri, brute=True)
    rays = np.array([[0, 0, 0, 0.01, 0, 1, 0, 1e4],      # straight up through the triangle at y = 3
                     [0, 0, 0, 0.01, 0, 1, 0, 2.9],      # t_max short of the triangle
                     [0, 3.005, 0, 0.01, 0, -1, 0, 1e4],  # hit at t = 0.005 < t_min
                     [10, 0, 0, 0.01, 0, 1, 0, 1e4],     # misses sideways
                     [0, 5, 0, 0.01, 0, -1, 0, 1e4]],    # back face: no culling
                    np.float32)
    assert s.trace_any(rays).tolist() == [1, 0, 0, 0, 1]
    t, prim, uv = s.trace_closest(rays)
    assert t[0] == 3.0 and prim[0] == 0 and prim[3] == 0xFFFFFFFF and t[4] == 2.0


def small_sequence(W=64, H=48, frames=4, sh_scale=0, ao_scale=1):
    sc = pyhr.SynthScene(pyhr.SCENE_SHADOWS_TEST)
    tri, _ = sc.world_triangles()
    osc = O.Scene(tri, brute=True)
    bn = pyhr.blue_noise()
    sh, ao = O.ShadowsOracle(W, H, sh_scale), O.AOOracle(W, H, ao_scale)
    f, prev = None, O.zero_gbuf_mips(W, H)
    for i in range(frames):
        dx = 0.0 if i < 2 else 0.05 * (i - 1)
        f = pyhr.make_frame((dx, 14, 34), (dx, 3, 0), W, H, prev=f, num_frames=i)
        cur = O.GBufMips(pyhr.write_gbuffer(sc, f, W, H))
        sh.render(osc, cur, prev, f, bn)
        ao.render(osc, cur, prev, f, bn)
        prev = cur
    return sh, ao


def test_golden_vectors():
    """The committed golden outputs (generated by tests/golden/make_golden.py from this oracle) are reproduced bit for bit."""
    g = np.load(os.path.join(GOLDEN, "shadows_ao_64x48.npz"))
    sh, ao = small_sequence()
    assert np.array_equal(sh.mask, g["sh_mask"])
    assert np.array_equal(ao.mask, g["ao_mask"])
    for name, arr in (("sh_temporal", sh.temporal), ("sh_moments", sh.cur_moments), ("sh_final", sh.final), ("sh_prev_image", sh.prev_image),
                      ("sh_tiles", sh.tile_flags), ("ao_temporal", ao.temporal), ("ao_blur", ao.blur[1]), ("ao_final", ao.final), ("ao_tiles", ao.tile_flags)):
        a, b = np.asarray(arr), g[name]
        if a.dtype == np.uint16:  # fp16 images: allow 1 ulp for libm differences (expf/powf) between build hosts
            d = np.abs(O.h2f(a) - O.h2f(b))
            assert d.max() <= 1e-3, name
        else:
            assert np.array_equal(a, b), name


def test_golden_sequence_256x144():
    """The larger fixture (state after the 12-frame static + pan sequence, the one tests/test_gpu_golden.py compares the CUDA path
    with) is reproduced by the oracle: integer images bit for bit, fp16 images within one ulp of libm noise."""
    sys.path.insert(0, GOLDEN)
    from make_golden import seq12_oracle
    g = np.load(os.path.join(GOLDEN, "shadows_ao_256x144_seq12.npz"))
    sh, ao = seq12_oracle()
    for name, arr in (("sh_mask", sh.mask), ("sh_tiles", sh.tile_flags), ("ao_mask", ao.mask), ("ao_tiles", ao.tile_flags)):
        assert np.array_equal(arr, g[name]), name
    for name, arr in (("sh_temporal", sh.temporal), ("sh_moments", sh.cur_moments), ("sh_atrous", sh.atrous_out), ("sh_prev_image", sh.prev_image),
                      ("sh_final", sh.final), ("ao_temporal", ao.temporal), ("ao_length", ao.cur_length), ("ao_blur", ao.blur[1]), ("ao_final", ao.final)):
        assert np.abs(O.h2f(np.asarray(arr)) - O.h2f(g[name])).max() <= 1e-3, name


def test_golden_ddgi_reflections_192x112():
    """DDGI + reflections fixture (tests/test_gpu_golden.py compares the CUDA path with it): reproduced by the oracle; exact where
    the values are decisions (ray lengths, tile flags, history length), within fp16 / libm noise elsewhere."""
    sys.path.insert(0, GOLDEN)
    from make_golden import gi_oracle
    g = np.load(os.path.join(GOLDEN, "ddgi_reflections_192x112_seq5.npz"))
    odd, orf = gi_oracle()
    assert np.array_equal(orf.tile_flags, g["refl_tiles"])
    assert np.array_equal(O.h2f(orf.rt)[..., 3], O.h2f(g["refl_rt"])[..., 3])
    assert np.array_equal(O.h2f(orf.cur_moments)[..., 2], O.h2f(g["refl_moments"])[..., 2])
    for name, arr in (("ddgi_sample", odd.sample), ("refl_rt", orf.rt), ("refl_temporal", orf.cur_temporal), ("refl_moments", orf.cur_moments),
                      ("refl_atrous", orf.atrous_out), ("refl_final", orf.final)):
        a, b = O.h2f(np.asarray(arr)), O.h2f(g[name])
        assert (np.abs(a - b) / np.maximum(1.0, np.abs(b))).max() <= 2e-3, name


def _unpack_mask(mask, W, H):
    y, x = np.mgrid[0:H, 0:W]
    return ((mask[y >> 2, x >> 3] >> ((y & 3) * 8 + (x & 7)).astype(np.uint32)) & 1).astype(np.uint8)


def test_spp_count_images():
    """spp > 1 (spp rays per pixel, sample index num_frames * spp + s, 8-bit count image instead of the bit mask):
    * spp = 1 through the count path is the mask path: count == unpacked mask, every temporal output bit-identical;
    * with a static camera the 2-spp count of frame n is the sum of the 1-spp masks of frames 2n and 2n+1 (same rays);
    * the denoised 2-spp result stays a visibility in [0, 1] close to the 1-spp mean."""
    W, H = 64, 48
    sc = pyhr.SynthScene(pyhr.SCENE_SHADOWS_TEST)
    tri, _ = sc.world_triangles()
    osc = O.Scene(tri, brute=True)
    bn = pyhr.blue_noise()
    L = O.lib()
    import ctypes as C
    f0 = pyhr.make_frame((0, 14, 34), (0, 3, 0), W, H)
    g = O.GBufMips(pyhr.write_gbuffer(sc, f0, W, H))
    gc = g.c(0)
    # (1) spp = 1: count path == mask path, including the temporal stage
    sh1 = O.ShadowsOracle(W, H, 0)
    zero = O.zero_gbuf_mips(W, H)
    sh1.render(osc, g, zero, f0, bn)
    cnt = np.zeros((H, W), np.uint8)
    L.orc_shadows_ray_trace_spp(osc.h, C.byref(gc), C.byref(f0), sh1.params.bias, 1, O.p(bn[0]), O.p(bn[1]), O.p(cnt))
    assert np.array_equal(cnt, _unpack_mask(sh1.mask, W, H))
    t2, m2, tf2 = np.zeros_like(sh1.temporal), np.zeros_like(sh1.moments[0]), np.zeros_like(sh1.tile_flags)
    zimg, zmom = np.zeros((H, W, 2), np.uint16), np.zeros((H, W, 4), np.uint16)
    gz = zero.c(0)
    L.orc_shadows_temporal_spp(C.byref(gc), C.byref(gz), O.p(cnt), 1, O.p(zimg), O.p(zmom), C.byref(f0), sh1.params.alpha, sh1.params.moments_alpha,
                               O.p(t2), O.p(m2), O.p(tf2))
    assert np.array_equal(t2, sh1.temporal) and np.array_equal(m2, sh1.cur_moments) and np.array_equal(tf2, sh1.tile_flags)
    # (2) 2 spp at frame n == 1 spp at frames 2n and 2n+1 (static camera: only the sample index changes)
    for n in (0, 3):
        c2 = np.zeros((H, W), np.uint8)
        fn = pyhr.make_frame((0, 14, 34), (0, 3, 0), W, H, num_frames=n)
        L.orc_shadows_ray_trace_spp(osc.h, C.byref(gc), C.byref(fn), sh1.params.bias, 2, O.p(bn[0]), O.p(bn[1]), O.p(c2))
        acc = np.zeros((H, W), np.uint8)
        for k in (2 * n, 2 * n + 1):
            fk = pyhr.make_frame((0, 14, 34), (0, 3, 0), W, H, num_frames=k)
            mk = np.zeros_like(sh1.mask)
            L.orc_shadows_ray_trace(osc.h, C.byref(gc), C.byref(fk), sh1.params.bias, O.p(bn[0]), O.p(bn[1]), O.p(mk))
            acc += _unpack_mask(mk, W, H)
        assert np.array_equal(c2, acc)
    # (3) full chains at 2 spp
    sh2, ao2 = O.ShadowsOracle(W, H, 0, spp=2), O.AOOracle(W, H, 1, spp=2)
    sh1b, ao1b = O.ShadowsOracle(W, H, 0), O.AOOracle(W, H, 1)
    f, prev = None, zero
    for i in range(4):
        f = pyhr.make_frame((0, 14, 34), (0, 3, 0), W, H, prev=f, num_frames=i)
        for o in (sh2, ao2, sh1b, ao1b):
            o.render(osc, g, prev, f, bn)
        prev = g
    v2, v1 = O.h2f(sh2.final)[..., 0], O.h2f(sh1b.final)[..., 0]
    assert v2.min() >= 0.0 and v2.max() <= 1.0 + 1e-3 and abs(float(v2.mean()) - float(v1.mean())) < 0.05
    a2, a1 = O.h2f(ao2.final), O.h2f(ao1b.final)
    assert a2.min() >= 0.0 and a2.max() <= 1.0 + 1e-3 and abs(float(a2.mean()) - float(a1.mean())) < 0.05
    assert sh2.count.max() <= 2 and ao2.count.max() <= 2 and sh2.count.any()


def test_mask_bit_order_and_tile_partition():
    sh, ao = small_sequence(frames=2)
    W, H = 64, 48
    # bit (y&3)*8 + (x&7) of word (x>>3, y>>2): a sky pixel must be 0 in the shadow mask and in the temporal output
    vis = O.h2f(sh.temporal)[..., 0]
    m = sh.mask
    bits = np.zeros((H, W), np.uint8)
    for y in range(H):
        for x in range(W):
            bits[y, x] = (int(m[y >> 2, x >> 3]) >> ((y & 3) * 8 + (x & 7))) & 1
    # every tile is on exactly one of the two lists, and shadow-list tiles have no visible pixel after temporal accumulation
    tf = sh.tile_flags.astype(bool)
    for ty in range(H // 8):
        for tx in range(W // 8):
            blk = vis[ty * 8:(ty + 1) * 8, tx * 8:(tx + 1) * 8]
            assert tf[ty, tx] == bool((blk > 0).any())
    assert bits.sum() > 0 and bits.sum() < W * H


def test_history_length_saturates_at_32():
    W, H = 64, 48
    sc = pyhr.SynthScene(pyhr.SCENE_SHADOWS_TEST)
    tri, _ = sc.world_triangles()
    osc = O.Scene(tri, brute=True)
    bn = pyhr.blue_noise()
    sh = O.ShadowsOracle(W, H, 0)
    f, prev, cur = None, O.zero_gbuf_mips(W, H), None
    lens = []
    for i in range(36):
        f = pyhr.make_frame((0, 14, 34), (0, 3, 0), W, H, prev=f, num_frames=i)
        if i < 2:
            cur = O.GBufMips(pyhr.write_gbuffer(sc, f, W, H))
        sh.render(osc, cur, prev, f, bn)
        prev = cur
        lens.append(O.h2f(sh.cur_moments)[..., 2].max())
    assert lens[0] == 1.0 and lens[5] == 6.0 and lens[-1] == 32.0 and max(lens) == 32.0


def test_atrous_constant_input_is_identity():
    """With constant visibility, zero variance, flat geometry the filter is a normalised blur: output == input (power 0)."""
    W, H = 32, 32
    g = pyhr.GBufferHost(W, H)
    g.gb2[..., :] = np.array([0, 0, 0, 0], np.float16).view(np.uint16)          # normal (0,0,1)
    g.gb3[..., :] = np.array([0.5, 0, 1, 10.0], np.float16).view(np.uint16)     # linear z = 10
    g.depth[:] = 0.5
    gm = O.GBufMips(g, 1)
    inp = np.zeros((H, W, 2), np.float16)
    inp[..., 0] = 0.625
    out = np.zeros((H, W, 2), np.uint16)
    tiles = np.ones((H // 8, W // 8), np.uint8)
    gc = gm.c(0)
    import ctypes as C
    for step in (1, 2, 4, 8):
        O.lib().orc_shadows_atrous(C.byref(gc), O.p(inp.view(np.uint16)), O.p(tiles), 1, step, 10.0, 32.0, 1.0, 0.0, O.p(out))
        assert np.array_equal(O.h2f(out)[..., 0], np.full((H, W), 0.625, np.float32))
    tiles[:] = 0  # shadow-list tiles are zero-filled (copy_shadow_tiles)
    O.lib().orc_shadows_atrous(C.byref(gc), O.p(inp.view(np.uint16)), O.p(tiles), 1, 1, 10.0, 32.0, 1.0, 0.0, O.p(out))
    assert not out.any()


def test_gbuffer_writer_conventions():
    W, H = 96, 64
    sc = pyhr.SynthScene(pyhr.SCENE_SHADOWS_TEST)
    f = pyhr.make_frame((0, 14, 34), (0, 3, 0), W, H)
    g = pyhr.write_gbuffer(sc, f, W, H)
    sky = g.depth == 1.0
    assert 0.05 < sky.mean() < 0.95
    assert np.all(O.h2f(g.gb3[..., 3])[sky] == -1.0) and np.all(O.h2f(g.gb3[..., 3])[~sky] > 0)   # g_buffer.cpp:88,96
    assert np.all((g.depth[~sky] >= 0) & (g.depth[~sky] < 1))
    # mip chain: level k texel (x,y) = level k-1 texel (2x+1, 2y+1)
    m = O.GBufMips(g)
    assert np.array_equal(m.levels[1][4], g.depth[1::2, 1::2])
    assert np.array_equal(m.levels[2][2], m.levels[1][2][1::2, 1::2])
