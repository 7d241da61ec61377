// glsl_fast.cuh — device restatements of the shared GLSL includes used by the floating-point (tolerance-checked)
// denoise kernels.  Unlike det_math.cuh these may be FMA-contracted and use fast intrinsics.
//   common.glsl:143-191, edge_stopping.glsl:10-62, reprojection.glsl:11-67
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gf {

__device__ __forceinline__ float2 h2_to_f2(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }
__device__ __forceinline__ uint32_t f2_to_h2(float a, float b)
{
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float4 h4_to_f4(uint2 w)
{
    const float2 a = h2_to_f2(w.x), b = h2_to_f2(w.y);
    return make_float4(a.x, a.y, b.x, b.y);
}

// common.glsl:150-156
__device__ __forceinline__ float3 octohedral_to_direction(float ex, float ey)
{
    float x = ex, y = ey, z = 1.0f - fabsf(ex) - fabsf(ey);
    if (z < 0.0f)
    {
        const float nx = (1.0f - fabsf(y)) * (x < 0.0f ? -1.0f : 1.0f);
        const float ny = (1.0f - fabsf(x)) * (y < 0.0f ? -1.0f : 1.0f);
        x = nx;
        y = ny;
    }
    const float inv = rsqrtf(x * x + y * y + z * z);
    return make_float3(x * inv, y * inv, z * inv);
}

__device__ __forceinline__ float fast_exp2(float x)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_rcp(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// common.glsl:169-184
__device__ __forceinline__ float3 world_position_from_depth(float u, float v, float d, const float* __restrict__ M)
{
    const float sx = u * 2.0f - 1.0f, sy = v * 2.0f - 1.0f;
    const float wx = M[0] * sx + M[4] * sy + M[8] * d + M[12];
    const float wy = M[1] * sx + M[5] * sy + M[9] * d + M[13];
    const float wz = M[2] * sx + M[6] * sy + M[10] * d + M[14];
    const float ww = M[3] * sx + M[7] * sy + M[11] * d + M[15];
    const float iw = 1.0f / ww;
    return make_float3(wx * iw, wy * iw, wz * iw);
}

// x^32 for x in [0,1] by repeated squaring (normal_edge_stopping_weight with phi_normal = 32, edge_stopping.glsl:10-13)
__device__ __forceinline__ float pow32(float x)
{
    x *= x; x *= x; x *= x; x *= x; x *= x;
    return x;
}
__device__ __forceinline__ float pow_pos(float x, float p) { return x <= 0.0f ? 0.0f : __expf(p * __logf(x)); }
__device__ __forceinline__ float normal_weight(float ndot, float phi_normal)
{
    const float c = fminf(fmaxf(ndot, 0.0f), 1.0f);
    return phi_normal == 32.0f ? pow32(c) : pow_pos(c, phi_normal);
}

} // namespace gf
