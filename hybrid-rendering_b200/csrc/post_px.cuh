// post_px.cuh — per-pixel functions of the two post-processing passes:
//   temporal anti-aliasing   src/shaders/taa.comp:60-420 (Playdead temporal reprojection as the reference configures it: USE_DILATION,
//                            MINMAX_3X3_ROUNDED, USE_CLIPPING, UNJITTER_*, HDR_CORRECTION; no USE_YCOCG / USE_OPTIMIZATIONS)
//   tone map                 src/shaders/tone_map.frag:38-66 (exposure, ACES film curve, gamma 1 / 2.2)
// Written as __host__ __device__ functions: post.cu wraps them in kernels (one thread per pixel); tests/hostemu compiles THE SAME
// functions for the CPU and compares them with the independently written oracle (oracle/orc_post.cpp), so the arithmetic of the
// kernels is checked in the CPU suite, not only on the GPU box.
//
// TAA is specified like the visibility-mask chain: + - * / min max only, in the order the shader writes them, no
// FMA contraction (post.cu is compiled with -fmad=false, the oracle and the host emulation with -ffp-contract=off).  The reason is
// the texel SELECTION of the nearest-filtered depth / velocity taps and the bilinear footprints: the reference's Halton jitter of
// -0.5 / W puts sample coordinates exactly on texel borders, where one ulp decides which texel is read.  Bit-exact parity.
#pragma once
#include "../../include/hr_api.h"
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef __CUDACC__
#include <cuda_fp16.h>
#define HR_HD __host__ __device__ __forceinline__
#else
#define HR_HD inline
#endif

namespace post {

struct F4 { float x, y, z, w; };
HR_HD F4 f4(float x, float y, float z, float w) { F4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
HR_HD F4 add4(F4 a, F4 b) { return f4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
HR_HD F4 sub4(F4 a, F4 b) { return f4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
HR_HD F4 mul4(F4 a, float s) { return f4(a.x * s, a.y * s, a.z * s, a.w * s); }
HR_HD F4 div4(F4 a, float s) { return f4(a.x / s, a.y / s, a.z / s, a.w / s); }
HR_HD F4 min4(F4 a, F4 b) { return f4(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z), fminf(a.w, b.w)); }
HR_HD F4 max4(F4 a, F4 b) { return f4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w)); }

// ---- binary16 <-> binary32 (exact conversions; the device uses the hardware instructions, the host the same mapping in integer code) ----
HR_HD float half_bits_to_float(uint16_t h)
{
#ifdef __CUDA_ARCH__
    return __half2float(__ushort_as_half(h));
#else
    const uint32_t s = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 31u, m = h & 1023u;
    uint32_t u;
    if (e == 0)
    {
        if (m == 0) u = s;
        else
        {
            int      k = 0;
            uint32_t mm = m;
            while (!(mm & 1024u)) { mm <<= 1; k++; }
            u = s | ((uint32_t)(113 - k) << 23) | ((mm & 1023u) << 13);
        }
    }
    else if (e == 31) u = s | 0x7F800000u | (m << 13);
    else u = s | ((e + 112u) << 23) | (m << 13);
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
HR_HD uint16_t float_to_half_bits(float f)
{ // round to nearest even, like __float2half_rn
#ifdef __CUDA_ARCH__
    return __half_as_ushort(__float2half_rn(f));
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    const uint32_t s = (u >> 16) & 0x8000u, a = u & 0x7FFFFFFFu;
    if (a >= 0x7F800000u) return (uint16_t)(s | (a > 0x7F800000u ? 0x7FFFu : 0x7C00u)); // NaN (canonical, as the hardware returns) / infinity
    if (a >= 0x477FF000u) return (uint16_t)(s | 0x7C00u);                                  // rounds to infinity
    if (a < 0x33000001u) return (uint16_t)s;                                               // rounds to zero (<= 2^-25)
    const int e = (int)(a >> 23) - 127;
    uint32_t  m = (a & 0x7FFFFFu) | 0x800000u;
    int       shift = e < -14 ? 13 + (-14 - e) : 13; // subnormal halves lose more bits
    uint32_t  q = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) q++;
    const uint32_t r = e < -14 ? q : (((uint32_t)(e + 15) << 10) + (q - 1024u)); // a mantissa carry propagates into the exponent
    return (uint16_t)(s | r);
#endif
}

// ---- images ---------------------------------------------------------------------------------------------------------------------
// a pass output as the samplers see it: R16F -> (r, 0, 0, 1), RG16F -> (r, g, 0, 1), RGBA16F (Vulkan's component fill rule)
struct ImgView { const uint16_t* p; int W, H, channels; };
struct alignas(8) H4 { uint16_t v[4]; }; // one RGBA16F texel = one 64-bit load (images are allocated 256-byte aligned, pitch = width * 8)
HR_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
// float -> int like cvt.rzi.s32.f32: toward zero, saturating, NaN -> 0 (a plain C++ cast is undefined outside the int range)
HR_HD int f2i(float f)
{
#ifdef __CUDA_ARCH__
    return __float2int_rz(f);
#else
    if (!(f == f)) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (int)0x80000000;
    return (int)f;
#endif
}
HR_HD F4 fetch_texel(const ImgView& im, int x, int y)
{ // CLAMP_TO_EDGE (vk.cpp:3453-3484)
    x = clampi(x, 0, im.W - 1);
    y = clampi(y, 0, im.H - 1);
    const size_t t = (size_t)y * im.W + x;
    if (im.channels == 4)
    {
        const H4 h = *reinterpret_cast<const H4*>(im.p + 4 * t);
        return f4(half_bits_to_float(h.v[0]), half_bits_to_float(h.v[1]), half_bits_to_float(h.v[2]), half_bits_to_float(h.v[3]));
    }
    const uint16_t* q = im.p + t * im.channels;
    F4 r = f4(half_bits_to_float(q[0]), 0.0f, 0.0f, 1.0f);
    if (im.channels >= 2) r.y = half_bits_to_float(q[1]);
    return r;
}
// textureLod(tex, uv, 0) with the bilinear sampler: unnormalised coordinate u * W - 0.5, weights from its fraction
HR_HD F4 sample_bilinear(const ImgView& im, float u, float v)
{
    const float x = u * (float)im.W - 0.5f, y = v * (float)im.H - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y);
    const float fx = x - fx0, fy = y - fy0;
    const int   x0 = f2i(fx0), y0 = f2i(fy0);
    const F4    a = fetch_texel(im, x0, y0), b = fetch_texel(im, x0 + 1, y0), c = fetch_texel(im, x0, y0 + 1), d = fetch_texel(im, x0 + 1, y0 + 1);
    const float wx0 = 1.0f - fx, wy0 = 1.0f - fy;
    return f4((a.x * wx0 + b.x * fx) * wy0 + (c.x * wx0 + d.x * fx) * fy, (a.y * wx0 + b.y * fx) * wy0 + (c.y * wx0 + d.y * fx) * fy,
              (a.z * wx0 + b.z * fx) * wy0 + (c.z * wx0 + d.z * fx) * fy, (a.w * wx0 + b.w * fx) * wy0 + (c.w * wx0 + d.w * fx) * fy);
}
// nearest sampler: texel floor(u * W), CLAMP_TO_EDGE
HR_HD int nearest_index(float u, int W) { return clampi(f2i(floorf(u * (float)W)), 0, W - 1); }

// ---- TAA ------------------------------------------------------------------------------------------------------------------------
struct TaaArgs {
    ImgView         cur, prev;   // s_Current (the visualised pass's output), s_Prev (this pass's previous output)
    const float*    depth;       // G-buffer depth, mip 0 (s_Depth)
    const uint16_t* gb2;         // G-buffer 2, RGBA16F: .zw = motion vector (s_Velocity)
    int             W, H;
    float           texel_x, texel_y;   // u_TexelSize.xy = 1 / W, 1 / H (temporal_aa.cpp:123)
    float           jitter_x, jitter_y; // u_CurrentPrevJitter.xy
    float           feedback_min, feedback_max;
    int             sharpen;
};
#define TAA_FLT_EPS 0.00000001f // taa.comp:62

// clip_aabb, taa.comp:121-156 (the #else branch: USE_OPTIMIZATIONS is not defined)
HR_HD F4 clip_aabb(F4 aabb_min, F4 aabb_max, F4 p, F4 q)
{
    F4          r = sub4(q, p);
    const float rmaxx = aabb_max.x - p.x, rmaxy = aabb_max.y - p.y, rmaxz = aabb_max.z - p.z;
    const float rminx = aabb_min.x - p.x, rminy = aabb_min.y - p.y, rminz = aabb_min.z - p.z;
    const float eps = TAA_FLT_EPS;
    if (r.x > rmaxx + eps) r = mul4(r, rmaxx / r.x);
    if (r.y > rmaxy + eps) r = mul4(r, rmaxy / r.y);
    if (r.z > rmaxz + eps) r = mul4(r, rmaxz / r.z);
    if (r.x < rminx - eps) r = mul4(r, rminx / r.x);
    if (r.y < rminy - eps) r = mul4(r, rminy / r.y);
    if (r.z < rminz - eps) r = mul4(r, rminz / r.z);
    return add4(p, r);
}
HR_HD float taa_luminance(F4 c) { return fmaxf((c.x * 0.299f + c.y * 0.587f) + c.z * 0.114f, 0.0001f); } // common.glsl:143-146, dot in the chain's order
HR_HD float taa_tonemap(float x) { return x / (x + 1.0f); }                                                 // taa.comp:247-250
HR_HD float taa_inverse_tonemap(float x) { return x / fmaxf(1.0f - x, TAA_FLT_EPS); }                      // :254-257

// main + find_closest_fragment_3x3 + temporal_reprojection, taa.comp:158-198, 261-418.  Returns the texel stored by imageStore (before fp16 rounding).
HR_HD F4 taa_pixel(const TaaArgs& A, int px, int py)
{
    const float tcx = ((float)px + 0.5f) * A.texel_x, tcy = ((float)py + 0.5f) * A.texel_y; // tex_coord
    const float uvx = tcx + A.jitter_x, uvy = tcy + A.jitter_y;                             // UNJITTER_REPROJECTION
    // find_closest_fragment_3x3: dd = abs(u_TexelSize.xy); the nine depth taps in the shader's order, strict '>' comparisons
    const float ddx = fabsf(A.texel_x), ddy = fabsf(A.texel_y);
    float       best_dx = -1.0f, best_dy = -1.0f, best_z = 0.0f;
    {
        const float us[3] = { uvx - ddx, uvx, uvx + ddx }; // uv - du, uv, uv + du
        const float vs[3] = { uvy - ddy, uvy, uvy + ddy };
        bool first = true;
        for (int j = 0; j < 3; j++)
            for (int i = 0; i < 3; i++)
            {
                const float z = A.depth[(size_t)nearest_index(vs[j], A.H) * A.W + nearest_index(us[i], A.W)];
                if (first || best_z > z) { best_dx = (float)(i - 1); best_dy = (float)(j - 1); best_z = z; first = false; }
            }
    }
    const float cfx = uvx + ddx * best_dx, cfy = uvy + ddy * best_dy; // c_frag.xy
    const H4    vel = *reinterpret_cast<const H4*>(A.gb2 + ((size_t)nearest_index(cfy, A.H) * A.W + nearest_index(cfx, A.W)) * 4);
    const float ss_vel_x = half_bits_to_float(vel.v[2]), ss_vel_y = half_bits_to_float(vel.v[3]); // texture(s_Velocity, c_frag.xy).zw

    // temporal_reprojection(tex_coord, ss_vel, vs_dist)
    F4       texel0 = sample_bilinear(A.cur, tcx + A.jitter_x, tcy + A.jitter_y); // UNJITTER_COLORSAMPLES
    F4       texel1 = sample_bilinear(A.prev, tcx + ss_vel_x, tcy + ss_vel_y);
    const float u = tcx + A.jitter_x, v = tcy + A.jitter_y;                     // UNJITTER_NEIGHBORHOOD
    const float dux = A.texel_x, dvy = A.texel_y;
    const F4 ctl = sample_bilinear(A.cur, u - dux, v - dvy), ctc = sample_bilinear(A.cur, u, v - dvy), ctr = sample_bilinear(A.cur, u + dux, v - dvy);
    const F4 cml = sample_bilinear(A.cur, u - dux, v), cmc = sample_bilinear(A.cur, u, v), cmr = sample_bilinear(A.cur, u + dux, v);
    const F4 cbl = sample_bilinear(A.cur, u - dux, v + dvy), cbc = sample_bilinear(A.cur, u, v + dvy), cbr = sample_bilinear(A.cur, u + dux, v + dvy);
    F4 cmin = min4(ctl, min4(ctc, min4(ctr, min4(cml, min4(cmc, min4(cmr, min4(cbl, min4(cbc, cbr))))))));
    F4 cmax = max4(ctl, max4(ctc, max4(ctr, max4(cml, max4(cmc, max4(cmr, max4(cbl, max4(cbc, cbr))))))));
    F4 cavg = div4(add4(add4(add4(add4(add4(add4(add4(add4(ctl, ctc), ctr), cml), cmc), cmr), cbl), cbc), cbr), 9.0f);
    { // MINMAX_3X3_ROUNDED
        const F4 cmin5 = min4(ctc, min4(cml, min4(cmc, min4(cmr, cbc))));
        const F4 cmax5 = max4(ctc, max4(cml, max4(cmc, max4(cmr, cbc))));
        const F4 cavg5 = div4(add4(add4(add4(add4(ctc, cml), cmc), cmr), cbc), 5.0f);
        cmin = mul4(add4(cmin, cmin5), 0.5f);
        cmax = mul4(add4(cmax, cmax5), 0.5f);
        cavg = mul4(add4(cavg, cavg5), 0.5f);
    }
    texel1 = clip_aabb(cmin, cmax, min4(max4(cavg, cmin), cmax), texel1); // USE_CLIPPING: clamp(cavg, cmin, cmax) = min(max(x, lo), hi)
    const float lum0 = taa_luminance(texel0), lum1 = taa_luminance(texel1);
    const float unbiased_diff   = fabsf(lum0 - lum1) / fmaxf(lum0, fmaxf(lum1, 0.2f));
    const float unbiased_weight = 1.0f - unbiased_diff;
    const float w2 = unbiased_weight * unbiased_weight;
    const float k_feedback = A.feedback_min * (1.0f - w2) + A.feedback_max * w2; // mix(min, max, w2)
    if (A.sharpen == 1)
    { // sum += -1 * cml; += -1 * ctc; += 5 * texel0; += -1 * cbc; += -1 * cmr  (taa.comp:361-371)
        F4 sum = f4(0.0f, 0.0f, 0.0f, 0.0f);
        sum = add4(sum, mul4(cml, -1.0f));
        sum = add4(sum, mul4(ctc, -1.0f));
        sum = add4(sum, mul4(texel0, 5.0f));
        sum = add4(sum, mul4(cbc, -1.0f));
        sum = add4(sum, mul4(cmr, -1.0f));
        texel0 = sum;
    }
    // HDR_CORRECTION
    const float t0x = taa_tonemap(texel0.x), t0y = taa_tonemap(texel0.y), t0z = taa_tonemap(texel0.z);
    const float t1x = taa_tonemap(texel1.x), t1y = taa_tonemap(texel1.y), t1z = taa_tonemap(texel1.z);
    const float omk = 1.0f - k_feedback;
    const float bx = taa_inverse_tonemap(t0x * omk + t1x * k_feedback), by = taa_inverse_tonemap(t0y * omk + t1y * k_feedback),
                bz = taa_inverse_tonemap(t0z * omk + t1z * k_feedback);
    // imageStore(i_Color, ..., vec4(clamp(to_buffer, 0.0, 1.0).xyz, 1.0f))
    return f4(fminf(fmaxf(bx, 0.0f), 1.0f), fminf(fmaxf(by, 0.0f), 1.0f), fminf(fmaxf(bz, 0.0f), 1.0f), 1.0f);
}

// ---- tone map -------------------------------------------------------------------------------------------------------------------
// aces_film, tone_map.frag:38-46
HR_HD float aces_film(float x)
{
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
    return fminf(fmaxf((x * (a * x + b)) / (x * (c * x + d) + e), 0.0f), 1.0f);
}
HR_HD uint32_t unorm8(float v) { return (uint32_t)(int)rintf(fminf(fmaxf(v, 0.0f), 1.0f) * 255.0f); } // float -> UNORM8 of the colour attachment
// main, tone_map.frag:52-66; c = the input texel (a full-screen pass at the input's resolution samples texel centres)
HR_HD uint32_t tonemap_pixel(F4 c, float exposure, int single_channel)
{
    float r, g, b;
    if (single_channel == 1) r = g = b = c.x;
    else
    {
        r = powf(aces_film(c.x * exposure), 1.0f / 2.2f);
        g = powf(aces_film(c.y * exposure), 1.0f / 2.2f);
        b = powf(aces_film(c.z * exposure), 1.0f / 2.2f);
    }
    return unorm8(r) | (unorm8(g) << 8) | (unorm8(b) << 16) | (255u << 24);
}

} // namespace post
