// trace.cu — software BVH traversal + the two mask-producing ray-trace kernels.
//   K1  shadows/shadows_ray_trace.comp:89-132   (1 soft-shadow ray / pixel -> 1 bit)
//   K7  ao/ao_ray_trace.comp:90-126             (1 cosine-lobe AO ray / pixel -> 1 bit)
// Ray queries replace rayQueryEXT (ray_query.glsl:6-59): any-hit, t in (t_min, t_max), opaque, no culling.
//
// BUILD NOTE: this file is compiled with -fmad=false: the mask chain (det_math.cuh + ray_triangle below) is specified
// without implicit FMA contraction so that the packed visibility mask is bit-exact against the CPU oracle.  The slab
// test uses explicit fmaf (it only has to be conservative, the boxes are padded at build time).
//
// Mapping: one warp = one 8x4 pixel block = one mask word (bit = lane = (y&3)*8 + (x&7), exactly
// gl_LocalInvocationIndex of the reference's 8x4 workgroup), so the word is a single __ballot_sync.

#include "traverse.cuh"

namespace {

using det::V3;
using namespace trv;

__device__ __forceinline__ float2 load_oct_normal(const uint2* gb2, size_t idx)
{
    const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(gb2 + idx)); // first two halves = oct normal
    const __half2  h = *reinterpret_cast<const __half2*>(&w);
    return __half22float2(h);
}

// MODE 0 = shadows (K1), 1 = AO (K7).  Block = 256 threads = 8 warps; warp w covers mask word (bx*4 + (w&3), by*2 + (w>>2)).
template <int MODE>
__global__ void __launch_bounds__(256) k_ray_trace_mask(GBufLevelDev g, BvhDev bvh, FrameConsts fc, float p0, float p1, const uint8_t* __restrict__ sobol,
                                                         const uint8_t* __restrict__ sr, uint32_t* __restrict__ mask, int mrow0, int mrow1)
{
    const int MW   = (g.W + 7) >> 3;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mx = blockIdx.x * 4 + (warp & 3), my = mrow0 + blockIdx.y * 2 + (warp >> 2);
    if (mx >= MW || my >= mrow1) return; // whole warp exits together
    const int x = mx * 8 + (lane & 7), y = my * 4 + (lane >> 3);
    uint32_t  result = 0;
    if (x < g.W && y < g.H)
    {
        const size_t idx   = (size_t)y * g.W + x;
        const float  depth = __ldg(g.depth + idx);
        if (depth != 1.0f)
        {
            const float  u = ((float)x + 0.5f) / (float)g.W, v = ((float)y + 0.5f) / (float)g.H;
            const V3     P  = det::world_position_from_depth(u, v, depth, fc.view_proj_inverse);
            const float2 e  = load_oct_normal(g.gb2, idx);
            const V3     N  = det::octohedral_to_direction(e.x, e.y);
            const float  r0 = det::sample_blue_noise(x, y, (int)fc.num_frames, 0, sobol, sr);
            const float  r1 = det::sample_blue_noise(x, y, (int)fc.num_frames, 1, sobol, sr);
            Ray          r;
            r.tmin = 0.01f;
            if (MODE == 0)
            {
                r.o = det::add(P, det::scale(N, p0)); // bias
                float att;
                det::fetch_light_properties_shadow(fc.light, P, N, r0, r1, r.d, r.tmax, att);
                if (att > 0.0f) result = trace_any(bvh, r) ? 0u : 1u;
            }
            else
            {
                r.o    = det::add(P, det::scale(N, p1)); // bias
                r.d    = det::sample_cosine_lobe(N, r0, r1);
                r.tmax = p0; // ray_length
                result = trace_any(bvh, r) ? 0u : 1u;
            }
        }
    }
    const uint32_t word = __ballot_sync(0xFFFFFFFFu, result != 0);
    if (lane == 0) mask[(size_t)my * MW + mx] = word;
}

__global__ void k_trace_any(BvhDev bvh, const float* __restrict__ rays, size_t n, uint32_t* __restrict__ out)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = rays + 8 * i;
    Ray          r;
    r.o = det::mk(p[0], p[1], p[2]); r.tmin = p[3];
    r.d = det::mk(p[4], p[5], p[6]); r.tmax = p[7];
    out[i] = trace_any(bvh, r) ? 1u : 0u;
}

__global__ void k_trace_closest(BvhDev bvh, const float* __restrict__ rays, size_t n, float* __restrict__ out_t, uint32_t* __restrict__ out_prim,
                                float* __restrict__ out_uv)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = rays + 8 * i;
    Ray          r;
    r.o = det::mk(p[0], p[1], p[2]); r.tmin = p[3];
    r.d = det::mk(p[4], p[5], p[6]); r.tmax = p[7];
    float    t, u, v;
    uint32_t prim;
    trace_closest(bvh, r, t, prim, u, v);
    out_t[i]    = t;
    out_prim[i] = prim;
    if (out_uv) { out_uv[2 * i] = u; out_uv[2 * i + 1] = v; }
}

} // namespace

static inline dim3 mask_grid(int W, int mrow0, int mrow1) { return dim3(((W + 7) / 8 + 3) / 4, (mrow1 - mrow0 + 1) / 2, 1); }

void launch_shadows_ray_trace(const GBufLevelDev& g, const BvhDev& bvh, const FrameConsts& fc, float bias, const uint8_t* sobol, const uint8_t* sr,
                              uint32_t* mask, int row0, int row1, cudaStream_t st)
{
    const int mrow0 = row0 / 4, mrow1 = (row1 + 3) / 4;
    if (mrow1 <= mrow0) return;
    k_ray_trace_mask<0><<<mask_grid(g.W, mrow0, mrow1), 256, 0, st>>>(g, bvh, fc, bias, 0.0f, sobol, sr, mask, mrow0, mrow1);
}

void launch_ao_ray_trace(const GBufLevelDev& g, const BvhDev& bvh, const FrameConsts& fc, float ray_length, float bias, const uint8_t* sobol,
                         const uint8_t* sr, uint32_t* mask, int row0, int row1, cudaStream_t st)
{
    const int mrow0 = row0 / 4, mrow1 = (row1 + 3) / 4;
    if (mrow1 <= mrow0) return;
    k_ray_trace_mask<1><<<mask_grid(g.W, mrow0, mrow1), 256, 0, st>>>(g, bvh, fc, ray_length, bias, sobol, sr, mask, mrow0, mrow1);
}

void launch_trace_any(const BvhDev& bvh, const float* rays, size_t n, uint32_t* out, cudaStream_t st)
{
    if (!n) return;
    k_trace_any<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(bvh, rays, n, out);
}

void launch_trace_closest(const BvhDev& bvh, const float* rays, size_t n, float* out_t, uint32_t* out_prim, float* out_uv, cudaStream_t st)
{
    if (!n) return;
    k_trace_closest<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(bvh, rays, n, out_t, out_prim, out_uv);
}
