// deferred.cu — deferred shading combine + sky box: src/shaders/deferred.frag:146-205, skybox.{vert,frag}, host
// src/deferred_shading.cpp:56-75 (render = render_shading + render_skybox), :646-731, :734-818.
// Consumes the four pass outputs (shadows .r :187, AO :188, reflections rgb :166, DDGI irradiance :162) + the G-buffer and writes
// Lo = direct_lighting(...) * visibility + indirect_lighting(...) as RGBA16F.  direct_lighting is the raster variant (no
// RAY_TRACING / SOFT_SHADOWS defines in deferred.frag: the visibility is the shadows pass's output).  Environment = constant
// colour (sky / prefiltered cubemaps and the irradiance SH are assets): evaluate_sh9_irradiance(N) = c, every prefiltered mip = c.
// Tolerance-checked colour stage: fast intrinsics allowed.
#include "glsl_fast.cuh"
#include "hr_internal.h"

namespace {

using namespace gf;

__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return make_float3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 norm3(float3 a) { const float i = rsqrtf(dot3(a, a)); return a * i; }

__device__ __forceinline__ float D_ggx(float ndoth, float alpha)
{
    const float a2 = alpha * alpha, denom = (ndoth * ndoth) * (a2 - 1.0f) + 1.0f;
    return a2 / fmaxf(0.0001f, 3.14159265359f * denom * denom);
}
__device__ __forceinline__ float G1_schlick_ggx(float roughness, float ndotv)
{
    const float k = ((roughness + 1.0f) * (roughness + 1.0f)) / 8.0f;
    return ndotv / fmaxf(0.0001f, ndotv * (1.0f - k) + k);
}
// brdf.glsl:130-142
__device__ __forceinline__ float3 evaluate_uber_brdf(float3 cd, float roughness, float3 N, float3 F0, float3 Wo, float3 Wh, float3 Wi)
{
    const float NdotL = fmaxf(dot3(N, Wi), 0.0f), NdotV = fmaxf(dot3(N, Wo), 0.0f), NdotH = fmaxf(dot3(N, Wh), 0.0f), VdotH = fmaxf(dot3(Wi, Wh), 0.0f);
    const float p5 = powf(1.0f - VdotH, 5.0f);
    const float3 F = F0 + (make_float3(1, 1, 1) - F0) * p5;
    const float spec = D_ggx(NdotH, roughness * roughness) * (G1_schlick_ggx(roughness, NdotL) * G1_schlick_ggx(roughness, NdotV)) / fmaxf(0.0001f, 4.0f * NdotL * NdotV);
    return (make_float3(1, 1, 1) - F) * (cd * (1.0f / 3.14159265359f)) + F * spec;
}

__device__ __forceinline__ float2 brdf_lut_fetch(const uint32_t* __restrict__ lut, float u, float v)
{ // bilinear CLAMP_TO_EDGE sampler on the 512 x 512 RG16F LUT
    const float x = u * 512.0f - 0.5f, y = v * 512.0f - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y), fx = x - fx0, fy = y - fy0;
    const int   x0 = min(max((int)fx0, 0), 511), x1 = min(max((int)fx0 + 1, 0), 511), y0 = min(max((int)fy0, 0), 511), y1 = min(max((int)fy0 + 1, 0), 511);
    const float2 a = h2_to_f2(__ldg(lut + y0 * 512 + x0)), b = h2_to_f2(__ldg(lut + y0 * 512 + x1)), c = h2_to_f2(__ldg(lut + y1 * 512 + x0)), e = h2_to_f2(__ldg(lut + y1 * 512 + x1));
    return make_float2((a.x * (1.0f - fx) + b.x * fx) * (1.0f - fy) + (c.x * (1.0f - fx) + e.x * fx) * fy,
                       (a.y * (1.0f - fx) + b.y * fx) * (1.0f - fy) + (c.y * (1.0f - fx) + e.y * fx) * fy);
}

struct DeferredParams { const void* shadow; int shadow_channels; const void* ao; const void* refl; const void* gi; float env[3]; const uint32_t* lut; int row0, row1; };

__global__ void __launch_bounds__(256) k_deferred(GBufLevelDev g, FrameConsts fc, DeferredParams P, uint2* __restrict__ out)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = P.row0 + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= g.W || y >= g.H || y >= P.row1) return;
    const size_t   pi  = (size_t)y * g.W + x;
    const float    ndc_depth = __ldg(g.depth + pi);
    if (ndc_depth == 1.0f)
    { // render_skybox (deferred_shading.cpp:69, 734-818): the sky cube is drawn at depth 1 over the pixels the G-buffer left at its clear depth;
      // skybox.frag:18-22 writes the environment cubemap's colour (a constant here), alpha 1
        out[pi] = make_uint2(f2_to_h2(P.env[0], P.env[1]), f2_to_h2(P.env[2], 1.0f));
        return;
    }
    const float    tu = ((float)x + 0.5f) / (float)g.W, tv = ((float)y + 0.5f) / (float)g.H;
    const float3   Pw = world_position_from_depth(tu, tv, ndc_depth, fc.view_proj_inverse);
    const uint32_t a8 = g.gb1 ? __ldg(g.gb1 + pi) : 0u;
    const float3   albedo   = make_float3((float)(a8 & 255u) / 255.0f, (float)((a8 >> 8) & 255u) / 255.0f, (float)((a8 >> 16) & 255u) / 255.0f);
    const float    metallic = (float)(a8 >> 24) / 255.0f;
    const uint2    g2 = __ldg(g.gb2 + pi), g3 = __ldg(g.gb3 + pi);
    const float    roughness = h2_to_f2(g3.x).x;
    const float2   oct = h2_to_f2(g2.x);
    const float3   N  = octohedral_to_direction(oct.x, oct.y);
    const float3   Wo = norm3(make_float3(fc.cam_pos[0], fc.cam_pos[1], fc.cam_pos[2]) - Pw);
    float visibility = 1.0f, aov = 1.0f;
    if (P.shadow)
        visibility = P.shadow_channels == 2 ? h2_to_f2(__ldg(reinterpret_cast<const uint32_t*>(P.shadow) + pi)).x
                                            : __half2float(__ushort_as_half(__ldg(reinterpret_cast<const unsigned short*>(P.shadow) + pi)));
    if (P.ao) aov = __half2float(__ushort_as_half(__ldg(reinterpret_cast<const unsigned short*>(P.ao) + pi)));
    const float3 F0 = make_float3(0.04f, 0.04f, 0.04f) * (1.0f - metallic) + albedo * metallic;
    const float3 cd = (albedo * (make_float3(1, 1, 1) - F0)) * (1.0f - metallic);
    float3 Lo = make_float3(0, 0, 0);
    { // direct_lighting (raster variant: no shadow ray) * visibility
        const hr_light& L = fc.light;
        const int    type = (int)L.data3[0];
        const float3 ldir = make_float3(L.data0[0], L.data0[1], L.data0[2]);
        const float3 Li = make_float3(L.data2[0] * L.data0[3], L.data2[1] * L.data0[3], L.data2[2] * L.data0[3]);
        float3 Wi = ldir;
        float  att = 1.0f;
        if (type != 0)
        {
            const float3 tl = make_float3(L.data1[0], L.data1[1], L.data1[2]) - Pw;
            const float  d2 = dot3(tl, tl), dist = sqrtf(d2);
            Wi  = tl * (1.0f / dist);
            att = 1.0f / d2;
            if (type == 2)
            {
                const float e0 = L.data3[1], e1 = L.data3[2];
                const float t  = fminf(fmaxf((dot3(Wi, ldir) - e0) / (e1 - e0), 0.0f), 1.0f);
                att *= t * t * (3.0f - 2.0f * t);
            }
        }
        att *= fminf(fmaxf(dot3(N, Wi), 0.0f), 1.0f);
        const float3 Wh = norm3(Wo + Wi);
        Lo = Lo + ((evaluate_uber_brdf(cd, roughness, N, F0, Wo, Wh, Wi) * att) * Li) * visibility;
    }
    { // indirect_lighting, deferred.frag:153-173
        const float  ndv = fmaxf(dot3(N, Wo), 0.0f), p5 = powf(fmaxf(1.0f - ndv, 0.0f), 5.0f), omr = 1.0f - roughness;
        const float3 F   = F0 + (make_float3(fmaxf(omr, F0.x), fmaxf(omr, F0.y), fmaxf(omr, F0.z)) - F0) * p5;
        const float3 kD  = (make_float3(1, 1, 1) - F) * (1.0f - metallic);
        const float3 env = make_float3(P.env[0], P.env[1], P.env[2]);
        float3 irr = env, pre = env;
        if (P.gi) { const float4 v = h4_to_f4(__ldg(reinterpret_cast<const uint2*>(P.gi) + pi)); irr = make_float3(v.x, v.y, v.z); }
        if (P.refl) { const float4 v = h4_to_f4(__ldg(reinterpret_cast<const uint2*>(P.refl) + pi)); pre = make_float3(v.x, v.y, v.z); }
        float3 spec = make_float3(0, 0, 0);
        if (P.lut)
        {
            const float2 b = brdf_lut_fetch(P.lut, ndv, roughness);
            spec = (pre * (F * b.x + make_float3(b.y, b.y, b.y))) * 2.0f; // IndirectSpecularStrength, deferred.frag:20
        }
        Lo = Lo + (kD * (irr * cd) + spec) * aov;
    }
    out[pi] = make_uint2(f2_to_h2(Lo.x, Lo.y), f2_to_h2(Lo.z, 1.0f));
}

} // namespace

void launch_deferred(const GBufLevelDev& g, const FrameConsts& fc, const void* shadow, int shadow_channels, const void* ao, const void* reflections, const void* gi,
                     const float* env3, const void* brdf_lut, void* out, int row0, int row1, cudaStream_t st)
{
    if (row1 <= row0) return;
    DeferredParams P { shadow, shadow_channels, ao, reflections, gi, { env3[0], env3[1], env3[2] }, (const uint32_t*)brdf_lut, row0, row1 };
    dim3 grid((g.W + 31) / 32, (row1 - row0 + 7) / 8);
    k_deferred<<<grid, 256, 0, st>>>(g, fc, P, (uint2*)out);
}
