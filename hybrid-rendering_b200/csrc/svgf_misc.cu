// svgf_misc.cu — the remaining HBM-streaming stages of the shadows / AO chains and the G-buffer mip chain.
//   K10 ao/ao_denoise_bilateral_blur.comp:75-139   (separable 9-tap bilateral gaussian, run twice)
//   K6  shadows/shadows_upsample.comp:62-109       K11 ao/ao_upsample.comp:63-112
//   mips: GBuffer NEAREST blit chain, src/g_buffer.cpp:236-244 -> vk.cpp:332-407
#include "glsl_fast.cuh"
#include "hr_internal.h"

namespace {

using namespace gf;

struct BlurParams { int W, H, dirx, diry, radius, row0, row1; float zbp_z, zbp_w; };

// One warp per pixel row (32 px): all taps of a row are contiguous spans => coalesced for the horizontal pass,
// row-strided full-line accesses for the vertical pass (neighbouring rows are L1/L2 hits).
__global__ void __launch_bounds__(256) k_ao_blur(GBufLevelDev g, const __half* __restrict__ in, const uint8_t* __restrict__ tile_flags, BlurParams P,
                                                  __half* __restrict__ out)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = P.row0 + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= P.W || y >= P.H || y >= P.row1) return;
    const int    TW  = (P.W + 7) >> 3;
    const size_t idx = (size_t)y * P.W + x;
    const __half one = __float2half_rn(1.0f);
    if (!tile_flags[(size_t)(y >> 3) * TW + (x >> 3)]) { out[idx] = one; return; } // image cleared to 1.0, tile not dispatched
    const float depth = __ldg(g.depth + idx);
    if (depth == 1.0f) { out[idx] = one; return; }
    float        total_ao = __half2float(__ldg(in + idx)), total_w = 1.0f;
    const float  cz = 1.0f / (P.zbp_z * depth + P.zbp_w); // linear_eye_depth, common.glsl:188-191
    const float2 ce = h2_to_f2(__ldg(reinterpret_cast<const uint32_t*>(g.gb2 + idx)));
    const float3 cn = octohedral_to_direction(ce.x, ce.y);
    const float  deviation = (float)P.radius / 1.5f;
    const float  gnorm = 1.0f / sqrtf(2.0f * 3.14159265359f * deviation * deviation);
    const float  ginv  = 1.0f / (2.0f * deviation * deviation);
    for (int i = -P.radius; i <= P.radius; i++)
    {
        if (i == 0) continue;
        const int sx = x + P.dirx * i, sy = y + P.diry * i;
        float     sd = 0.0f, sao = 0.0f;
        float2    se = make_float2(0.0f, 0.0f);
        if (sx >= 0 && sy >= 0 && sx < P.W && sy < P.H)
        { // texelFetch out of bounds => zeros (depth 0, ao 0, normal oct(0,0))
            const size_t si = (size_t)sy * P.W + sx;
            sd  = __ldg(g.depth + si);
            sao = __half2float(__ldg(in + si));
            se  = h2_to_f2(__ldg(reinterpret_cast<const uint32_t*>(g.gb2 + si)));
        }
        const float  sz = 1.0f / (P.zbp_z * sd + P.zbp_w);
        const float3 sn = octohedral_to_direction(se.x, se.y);
        float        w  = gnorm * __expf(-((float)(i * i)) * ginv);                       // gaussian_weight, common.glsl:160-165
        w *= __expf(-1.0f - __expf(-fabsf(cz - sz))) * pow32(fminf(fmaxf(dot3(cn, sn), 0.0f), 1.0f)); // edge_stopping.glsl:31-62, wL = 1
        total_ao += w * sao;
        total_w += w;
    }
    out[idx] = __float2half_rn(total_ao / fmaxf(total_w, 0.0001f));
}

struct UpParams { int W0, H0, Wm, Hm, in_channels, row0, row1; float sky_value, power; };

// textureLod NEAREST + CLAMP_TO_EDGE: texel = clamp(floor(uv * size), 0, size-1)
__device__ __forceinline__ int nearest(float uv, int size) { return min(max((int)floorf(uv * (float)size), 0), size - 1); }

__global__ void __launch_bounds__(256) k_upsample_scalar(GBufLevelDev g0, GBufLevelDev gm, const __half* __restrict__ in, UpParams P, __half* __restrict__ out)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = P.row0 + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= P.W0 || y >= P.H0 || y >= P.row1) return;
    const size_t idx = (size_t)y * P.W0 + x;
    const float  hz  = h2_to_f2(__ldg(reinterpret_cast<const uint32_t*>(g0.gb3 + idx) + 1)).y;
    if (hz == -1.0f) { out[idx] = __float2half_rn(P.sky_value); return; }
    const float2 he = h2_to_f2(__ldg(reinterpret_cast<const uint32_t*>(g0.gb2 + idx)));
    const float3 hn = octohedral_to_direction(he.x, he.y);
    const float  tu = ((float)x + 0.5f) / (float)P.W0, tv = ((float)y + 0.5f) / (float)P.H0;
    const float  tsx = 1.0f / (float)P.Wm, tsy = 1.0f / (float)P.Hm;
    const float  kx[4] = { 0.0f, 1.0f, -1.0f, 0.0f }, ky[4] = { 1.0f, 0.0f, 0.0f, -1.0f };
    float        up = 0.0f, tw = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        const int    cx = nearest(tu + kx[i] * tsx, P.Wm), cy = nearest(tv + ky[i] * tsy, P.Hm);
        const size_t ci = (size_t)cy * P.Wm + cx;
        const float  cz = h2_to_f2(__ldg(reinterpret_cast<const uint32_t*>(gm.gb3 + ci) + 1)).y;
        if (cz == -1.0f) continue;
        const float2 ce = h2_to_f2(__ldg(reinterpret_cast<const uint32_t*>(gm.gb2 + ci)));
        const float3 cn = octohedral_to_direction(ce.x, ce.y);
        const float  w  = __expf(-1.0f - __expf(-fabsf(hz - cz))) * pow32(fminf(fmaxf(dot3(hn, cn), 0.0f), 1.0f));
        up += __half2float(__ldg(in + ci * P.in_channels)) * w;
        tw += w;
    }
    up = up / fmaxf(tw, 0.00000001f);
    if (P.power != 0.0f) up = pow_pos(up, P.power);
    out[idx] = __float2half_rn(up);
}

// dst(x,y) = src(min(2x+1, W-1), min(2y+1, H-1)) for gb2, gb3, depth (and gb1 when present)
__global__ void k_build_mip(int W, int H, const uint2* __restrict__ gb2, const uint2* __restrict__ gb3, const float* __restrict__ depth,
                            const uint32_t* __restrict__ gb1, uint2* __restrict__ o2, uint2* __restrict__ o3, float* __restrict__ od, uint32_t* __restrict__ o1)
{
    const int w = max(W / 2, 1), h = max(H / 2, 1);
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const size_t s = (size_t)min(2 * y + 1, H - 1) * W + min(2 * x + 1, W - 1), d = (size_t)y * w + x;
    o2[d] = gb2[s];
    o3[d] = gb3[s];
    od[d] = depth[s];
    if (gb1 && o1) o1[d] = gb1[s];
}

// Levels 1 and 2 in one launch: level 2 is a NEAREST blit of level 1, i.e. of level 0 at the composed coordinates
// (min(2*min(2X+1, w1-1)+1, W-1), ...), so it can be read straight from level 0 by the threads (x < w2, y < h2).
__global__ void k_build_mips12(int W, int H, const uint2* __restrict__ gb2, const uint2* __restrict__ gb3, const float* __restrict__ depth,
                               const uint32_t* __restrict__ gb1, uint2* __restrict__ o2a, uint2* __restrict__ o3a, float* __restrict__ oda, uint32_t* __restrict__ o1a,
                               uint2* __restrict__ o2b, uint2* __restrict__ o3b, float* __restrict__ odb, uint32_t* __restrict__ o1b)
{
    const int w1 = max(W / 2, 1), h1 = max(H / 2, 1), w2 = max(w1 / 2, 1), h2 = max(h1 / 2, 1);
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w1 || y >= h1) return;
    {
        const size_t s = (size_t)min(2 * y + 1, H - 1) * W + min(2 * x + 1, W - 1), d = (size_t)y * w1 + x;
        o2a[d] = gb2[s];
        o3a[d] = gb3[s];
        oda[d] = depth[s];
        if (gb1 && o1a) o1a[d] = gb1[s];
    }
    if (x < w2 && y < h2)
    {
        const int    x1 = min(2 * x + 1, w1 - 1), y1 = min(2 * y + 1, h1 - 1);
        const size_t s = (size_t)min(2 * y1 + 1, H - 1) * W + min(2 * x1 + 1, W - 1), d = (size_t)y * w2 + x;
        o2b[d] = gb2[s];
        o3b[d] = gb3[s];
        odb[d] = depth[s];
        if (gb1 && o1b) o1b[d] = gb1[s];
    }
}

} // namespace

void launch_ao_blur_v1(const GBufLevelDev& g, const __half* in, const uint8_t* tile_flags, const float* zbp, int dirx, int diry, int radius, __half* out,
                    int row0, int row1, cudaStream_t st)
{
    if (row1 <= row0) return;
    BlurParams P { g.W, g.H, dirx, diry, radius, row0, row1, zbp[2], zbp[3] };
    dim3       grid((g.W + 31) / 32, (row1 - row0 + 7) / 8);
    k_ao_blur<<<grid, 256, 0, st>>>(g, in, tile_flags, P, out);
}

void launch_upsample_scalar_v1(const GBufLevelDev& g0, const GBufLevelDev& gm, const void* in, int in_channels, float sky_value, float power, __half* out,
                            int row0, int row1, cudaStream_t st)
{
    if (row1 <= row0) return;
    UpParams P { g0.W, g0.H, gm.W, gm.H, in_channels, row0, row1, sky_value, power };
    dim3     grid((g0.W + 31) / 32, (row1 - row0 + 7) / 8);
    k_upsample_scalar<<<grid, 256, 0, st>>>(g0, gm, reinterpret_cast<const __half*>(in), P, out);
}

int hr_launch_build_mips(hr_ctx* ctx, GBufSlot& s, int W, int H, cudaStream_t st)
{
    if (HR_MAX_MIPS == 3)
    {
        const int w1 = W / 2 > 0 ? W / 2 : 1, h1 = H / 2 > 0 ? H / 2 : 1;
        dim3      b(32, 8), g((w1 + 31) / 32, (h1 + 7) / 8);
        k_build_mips12<<<g, b, 0, st>>>(W, H, (const uint2*)s.gb2[0], (const uint2*)s.gb3[0], s.depth[0], (const uint32_t*)s.gb1[0], (uint2*)s.gb2[1], (uint2*)s.gb3[1],
                                        s.depth[1], (uint32_t*)s.gb1[1], (uint2*)s.gb2[2], (uint2*)s.gb3[2], s.depth[2], (uint32_t*)s.gb1[2]);
        ctx->launches++;
        HR_CHECK_LAUNCH(ctx);
        return HR_OK;
    }
    int w = W, h = H;
    for (int m = 1; m < HR_MAX_MIPS; m++)
    {
        const int nw = w / 2 > 0 ? w / 2 : 1, nh = h / 2 > 0 ? h / 2 : 1;
        dim3      b(32, 8), g((nw + 31) / 32, (nh + 7) / 8);
        k_build_mip<<<g, b, 0, st>>>(w, h, (const uint2*)s.gb2[m - 1], (const uint2*)s.gb3[m - 1], s.depth[m - 1], (const uint32_t*)s.gb1[m - 1],
                                     (uint2*)s.gb2[m], (uint2*)s.gb3[m], s.depth[m], (uint32_t*)s.gb1[m]);
        ctx->launches++;
        w = nw;
        h = nh;
    }
    HR_CHECK_LAUNCH(ctx);
    return HR_OK;
}
