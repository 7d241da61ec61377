// oracle/orc_post.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle). Never linked into the product library.
//
// CPU restatement of the post-processing shaders, one C++ function per GLSL function, same names:
//   src/shaders/taa.comp:60-420      temporal anti-aliasing (Playdead), with the defines the reference sets (:16-23): USE_DILATION,
//                                    MINMAX_3X3_ROUNDED, USE_CLIPPING, UNJITTER_REPROJECTION / COLORSAMPLES / NEIGHBORHOOD, HDR_CORRECTION
//   src/shaders/tone_map.frag:38-66  exposure, ACES film, gamma
//   src/temporal_aa.cpp:30-42,66-81  Halton jitter
// Samplers (vk.cpp:3453-3484): s_Current / s_Prev bilinear, G-buffer nearest, all CLAMP_TO_EDGE, textureLod(.., 0).
// Parity unpinned in the sense of DESIGN.md §2 (the reference cannot run here); the literals are pinned by tests/test_ref_constants.py.
// TAA arithmetic: plain IEEE binary32 + - * / min max in the shader's order, no contraction (-ffp-contract=off) — the statement the
// CUDA kernel (csrc/post_px.cuh, built with -fmad=false) must reproduce bit for bit.
#include "orc_glsl.h"
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

namespace {

inline vec4 operator+(vec4 a, vec4 b) { return { a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w }; }
inline vec4 operator-(vec4 a, vec4 b) { return { a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w }; }
inline vec4 operator*(vec4 a, float s) { return { a.x * s, a.y * s, a.z * s, a.w * s }; }
inline vec4 operator*(float s, vec4 a) { return { s * a.x, s * a.y, s * a.z, s * a.w }; }
inline vec4 operator/(vec4 a, float s) { return { a.x / s, a.y / s, a.z / s, a.w / s }; }
inline vec4 vmin(vec4 a, vec4 b) { return { fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z), fminf(a.w, b.w) }; }
inline vec4 vmax(vec4 a, vec4 b) { return { fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w) }; }
inline vec4 vclamp(vec4 x, vec4 lo, vec4 hi) { return vmin(vmax(x, lo), hi); } // GLSL clamp = min(max(x, minVal), maxVal)

// sampler2D over a half-float image with 1, 2 or 4 channels: missing components read (0, 0, 1)
struct Sampler2D {
    int W, H, C;
    const uint16_t* d;
    vec4 texel(int x, int y) const
    { // CLAMP_TO_EDGE
        x = std::min(std::max(x, 0), W - 1);
        y = std::min(std::max(y, 0), H - 1);
        const uint16_t* t = d + (size_t)C * ((size_t)y * W + x);
        vec4 r = { h2f(t[0]), 0.0f, 0.0f, 1.0f };
        if (C >= 2) r.y = h2f(t[1]);
        if (C == 4) { r.z = h2f(t[2]); r.w = h2f(t[3]); }
        return r;
    }
    vec4 bilinear(vec2 uv) const
    { // VK_FILTER_LINEAR: texel-space coordinate u * W - 0.5; i0 = floor, weights = fraction
        float x = uv.x * (float)W - 0.5f, y = uv.y * (float)H - 0.5f;
        float i0 = floorf(x), j0 = floorf(y);
        float a = x - i0, b = y - j0;
        int   i = f2i(i0), j = f2i(j0);
        vec4  t00 = texel(i, j), t10 = texel(i + 1, j), t01 = texel(i, j + 1), t11 = texel(i + 1, j + 1);
        return (t00 * (1.0f - a) + t10 * a) * (1.0f - b) + (t01 * (1.0f - a) + t11 * a) * b;
    }
};
// nearest sampler over float / RGBA16F G-buffer images: texel floor(u * W), CLAMP_TO_EDGE
inline int nearest(float u, int W) { return std::min(std::max(f2i(floorf(u * (float)W)), 0), W - 1); }

struct TaaUniforms { // PushConstants, taa.comp:48-56
    vec4  u_TexelSize;
    vec4  u_CurrentPrevJitter;
    float u_FeedbackMin, u_FeedbackMax;
    int   u_Sharpen;
};
struct TaaInputs {
    Sampler2D    s_Current, s_Prev;
    const float* depth;    // s_Depth
    const uint16_t* gb2;   // s_Velocity (.zw)
    int W, H;
    float s_depth(vec2 uv) const { return depth[(size_t)nearest(uv.y, H) * W + nearest(uv.x, W)]; }
    vec2  s_velocity_zw(vec2 uv) const
    {
        const uint16_t* t = gb2 + 4 * ((size_t)nearest(uv.y, H) * W + nearest(uv.x, W));
        return { h2f(t[2]), h2f(t[3]) };
    }
};

const float FLT_EPS = 0.00000001f; // taa.comp:62

// taa.comp:121-156, #else branch (USE_OPTIMIZATIONS undefined)
vec4 clip_aabb(vec3 aabb_min, vec3 aabb_max, vec4 p, vec4 q)
{
    vec4  r    = q - p;
    vec3  rmax = aabb_max - vec3{ p.x, p.y, p.z };
    vec3  rmin = aabb_min - vec3{ p.x, p.y, p.z };
    const float eps = FLT_EPS;
    if (r.x > rmax.x + eps) r = r * (rmax.x / r.x);
    if (r.y > rmax.y + eps) r = r * (rmax.y / r.y);
    if (r.z > rmax.z + eps) r = r * (rmax.z / r.z);
    if (r.x < rmin.x - eps) r = r * (rmin.x / r.x);
    if (r.y < rmin.y - eps) r = r * (rmin.y / r.y);
    if (r.z < rmin.z - eps) r = r * (rmin.z / r.z);
    return p + r;
}

// taa.comp:160-196
vec3 find_closest_fragment_3x3(const TaaInputs& in, const TaaUniforms& u, vec2 uv)
{
    vec2 dd = { fabsf(u.u_TexelSize.x), fabsf(u.u_TexelSize.y) };
    vec2 du = { dd.x, 0.0f }, dv = { 0.0f, dd.y };
    vec3 dtl = { -1, -1, in.s_depth(uv - dv - du) }, dtc = { 0, -1, in.s_depth(uv - dv) }, dtr = { 1, -1, in.s_depth(uv - dv + du) };
    vec3 dml = { -1, 0, in.s_depth(uv - du) }, dmc = { 0, 0, in.s_depth(uv) }, dmr = { 1, 0, in.s_depth(uv + du) };
    vec3 dbl = { -1, 1, in.s_depth(uv + dv - du) }, dbc = { 0, 1, in.s_depth(uv + dv) }, dbr = { 1, 1, in.s_depth(uv + dv + du) };
    vec3 dmin = dtl;
    if (dmin.z > dtc.z) dmin = dtc;
    if (dmin.z > dtr.z) dmin = dtr;
    if (dmin.z > dml.z) dmin = dml;
    if (dmin.z > dmc.z) dmin = dmc;
    if (dmin.z > dmr.z) dmin = dmr;
    if (dmin.z > dbl.z) dmin = dbl;
    if (dmin.z > dbc.z) dmin = dbc;
    if (dmin.z > dbr.z) dmin = dbr;
    return { uv.x + dd.x * dmin.x, uv.y + dd.y * dmin.y, dmin.z };
}

vec3 tonemap(vec3 x) { return { x.x / (x.x + 1.0f), x.y / (x.y + 1.0f), x.z / (x.z + 1.0f) }; } // :247-250
vec3 inverse_tonemap(vec3 x)                                                                      // :254-257
{
    return { x.x / fmaxf(1.0f - x.x, FLT_EPS), x.y / fmaxf(1.0f - x.y, FLT_EPS), x.z / fmaxf(1.0f - x.z, FLT_EPS) };
}

// taa.comp:261-389
vec3 temporal_reprojection(const TaaInputs& in, const TaaUniforms& u, vec2 ss_txc, vec2 ss_vel, float /*vs_dist*/)
{
    const vec2 jitter = { u.u_CurrentPrevJitter.x, u.u_CurrentPrevJitter.y };
    vec4 texel0 = in.s_Current.bilinear(ss_txc + jitter); // UNJITTER_COLORSAMPLES
    vec4 texel1 = in.s_Prev.bilinear(ss_txc + ss_vel);
    vec2 uv = ss_txc + jitter; // UNJITTER_NEIGHBORHOOD
    vec2 du = { u.u_TexelSize.x, 0.0f }, dv = { 0.0f, u.u_TexelSize.y };
    vec4 ctl = in.s_Current.bilinear(uv - dv - du), ctc = in.s_Current.bilinear(uv - dv), ctr = in.s_Current.bilinear(uv - dv + du);
    vec4 cml = in.s_Current.bilinear(uv - du), cmc = in.s_Current.bilinear(uv), cmr = in.s_Current.bilinear(uv + du);
    vec4 cbl = in.s_Current.bilinear(uv + dv - du), cbc = in.s_Current.bilinear(uv + dv), cbr = in.s_Current.bilinear(uv + dv + du);
    vec4 cmin = vmin(ctl, vmin(ctc, vmin(ctr, vmin(cml, vmin(cmc, vmin(cmr, vmin(cbl, vmin(cbc, cbr))))))));
    vec4 cmax = vmax(ctl, vmax(ctc, vmax(ctr, vmax(cml, vmax(cmc, vmax(cmr, vmax(cbl, vmax(cbc, cbr))))))));
    vec4 cavg = (ctl + ctc + ctr + cml + cmc + cmr + cbl + cbc + cbr) / 9.0f;
    // MINMAX_3X3_ROUNDED
    vec4 cmin5 = vmin(ctc, vmin(cml, vmin(cmc, vmin(cmr, cbc))));
    vec4 cmax5 = vmax(ctc, vmax(cml, vmax(cmc, vmax(cmr, cbc))));
    vec4 cavg5 = (ctc + cml + cmc + cmr + cbc) / 5.0f;
    cmin = 0.5f * (cmin + cmin5);
    cmax = 0.5f * (cmax + cmax5);
    cavg = 0.5f * (cavg + cavg5);
    // USE_CLIPPING
    texel1 = clip_aabb({ cmin.x, cmin.y, cmin.z }, { cmax.x, cmax.y, cmax.z }, vclamp(cavg, cmin, cmax), texel1);
    float lum0 = luminance({ texel0.x, texel0.y, texel0.z });
    float lum1 = luminance({ texel1.x, texel1.y, texel1.z });
    float unbiased_diff       = fabsf(lum0 - lum1) / fmaxf(lum0, fmaxf(lum1, 0.2f));
    float unbiased_weight     = 1.0f - unbiased_diff;
    float unbiased_weight_sqr = unbiased_weight * unbiased_weight;
    float k_feedback          = mixf(u.u_FeedbackMin, u.u_FeedbackMax, unbiased_weight_sqr);
    if (u.u_Sharpen == 1)
    {
        vec4 sum = { 0, 0, 0, 0 };
        sum = sum + -1.0f * cml;
        sum = sum + -1.0f * ctc;
        sum = sum + 5.0f * texel0;
        sum = sum + -1.0f * cbc;
        sum = sum + -1.0f * cmr;
        texel0 = sum;
    }
    // HDR_CORRECTION
    vec3 t0 = tonemap({ texel0.x, texel0.y, texel0.z }), t1 = tonemap({ texel1.x, texel1.y, texel1.z });
    vec3 blended = mix3(t0, t1, k_feedback);
    return inverse_tonemap(blended);
}

} // namespace

// taa.comp main (:395-418) for every pixel.  cur: the visualised pass's final output (cur_channels = 1, 2 or 4 halves per texel);
// prev: this pass's previous output, RGBA16F; out: RGBA16F.
extern "C" void orc_taa(int W, int H, const uint16_t* cur, int cur_channels, const uint16_t* prev, const float* depth, const uint16_t* gb2, const float* jitter_xy,
                        float feedback_min, float feedback_max, int sharpen, uint16_t* out)
{
    TaaInputs in { Sampler2D { W, H, cur_channels, cur }, Sampler2D { W, H, 4, prev }, depth, gb2, W, H };
    TaaUniforms u;
    u.u_TexelSize         = { 1.0f / (float)W, 1.0f / (float)H, (float)W, (float)H }; // temporal_aa.cpp:123
    u.u_CurrentPrevJitter = { jitter_xy[0], jitter_xy[1], 0.0f, 0.0f };
    u.u_FeedbackMin = feedback_min; u.u_FeedbackMax = feedback_max; u.u_Sharpen = sharpen;
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
        {
            vec2 tex_coord = { ((float)x + 0.5f) * u.u_TexelSize.x, ((float)y + 0.5f) * u.u_TexelSize.y };
            vec2 uv = tex_coord + vec2{ u.u_CurrentPrevJitter.x, u.u_CurrentPrevJitter.y }; // UNJITTER_REPROJECTION
            vec3 c_frag  = find_closest_fragment_3x3(in, u, uv);                             // USE_DILATION
            vec2 ss_vel  = in.s_velocity_zw({ c_frag.x, c_frag.y });
            float vs_dist = c_frag.z;
            vec3 to_buffer = temporal_reprojection(in, u, tex_coord, ss_vel, vs_dist);       // resolve_color: identity without USE_YCOCG
            uint16_t* o = out + 4 * ((size_t)y * W + x);
            o[0] = f2h(clampf(to_buffer.x, 0.0f, 1.0f)); o[1] = f2h(clampf(to_buffer.y, 0.0f, 1.0f)); o[2] = f2h(clampf(to_buffer.z, 0.0f, 1.0f)); o[3] = f2h(1.0f);
        }
}

// vkCmdBlitImage into the RGBA16F history image (temporal_aa.cpp:112-121): format conversion with component fill
extern "C" void orc_blit_rgba16f(int W, int H, const uint16_t* src, int channels, uint16_t* out)
{
    Sampler2D s { W, H, channels, src };
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
        {
            vec4 t = s.texel(x, y);
            uint16_t* o = out + 4 * ((size_t)y * W + x);
            o[0] = f2h(t.x); o[1] = f2h(t.y); o[2] = f2h(t.z); o[3] = f2h(t.w);
        }
}

// halton_sequence + TemporalAA::update, temporal_aa.cpp:30-42, 54-55, 66-81
static float halton_sequence(int base, int index)
{
    float result = 0, f = 1;
    while (index > 0)
    {
        f /= base;
        result += f * (index % base);
        index = (int)floor(index / base);
    }
    return result;
}
extern "C" void orc_taa_jitter(uint32_t num_frames, int width, int height, float* out_xy)
{
    const int HALTON_SAMPLES = 16;
    std::vector<vec2> samples;
    for (int i = 1; i <= HALTON_SAMPLES; i++) samples.push_back({ 2.0f * halton_sequence(2, i) - 1.0f, 2.0f * halton_sequence(3, i) - 1.0f });
    const vec2 h = samples[num_frames % samples.size()];
    out_xy[0] = h.x / float(width);
    out_xy[1] = h.y / float(height);
}

// tone_map.frag:38-66
static vec3 aces_film(vec3 x)
{
    float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
    auto  f = [&](float v) { return clampf((v * (a * v + b)) / (v * (c * v + d) + e), 0.0f, 1.0f); };
    return { f(x.x), f(x.y), f(x.z) };
}
extern "C" void orc_tonemap(int W, int H, const uint16_t* src, int channels, float exposure, int single_channel, uint8_t* out_rgba8)
{
    Sampler2D s { W, H, channels, src };
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
        {
            vec4 t = s.texel(x, y);
            vec3 color;
            if (single_channel == 1) color = { t.x, t.x, t.x };
            else
            {
                color = vec3{ t.x, t.y, t.z } * exposure;
                color = aces_film(color);
                color = { powf(color.x, 1.0f / 2.2f), powf(color.y, 1.0f / 2.2f), powf(color.z, 1.0f / 2.2f) };
            }
            uint8_t* o = out_rgba8 + 4 * ((size_t)y * W + x);
            o[0] = (uint8_t)lrintf(clampf(color.x, 0.0f, 1.0f) * 255.0f);
            o[1] = (uint8_t)lrintf(clampf(color.y, 0.0f, 1.0f) * 255.0f);
            o[2] = (uint8_t)lrintf(clampf(color.z, 0.0f, 1.0f) * 255.0f);
            o[3] = 255;
        }
}
