// oracle/orc_gbuffer.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle). Never linked into the product library.
//
// CPU statement of the G-buffer producer: what the reference's raster pass writes
//   src/g_buffer.cpp:100-263 (attachments RGBA8 / RGBA16F / RGBA16F / D32, clears :72-96)
//   src/shaders/g_buffer.vert, g_buffer.frag:47-51 (direction_to_octohedral), :55-67 (compute_motion_vector),
//   :71-80 (compute_curvature), :87-111 (main)
// restated as a primary-visibility ray cast over the same triangles (the raster pipeline itself cannot be reproduced; the
// visible surface per pixel centre is the same thing).  Everything is a fixed sequence of binary32 operations (orc_math.h
// rules: no FMA contraction, correctly rounded + - * / sqrt) so that csrc/gbuffer.cu reproduces every output bit:
//   pixel centre (x+.5, y+.5) -> NDC -> world at ndc z = 0 (the Vulkan near clip of the reference's GL-convention projection,
//   SURVEY.md A.0) and ndc z = 1 through view_proj_inverse (world_position_from_depth, common.glsl:169-184);
//   ray o -> e, closest hit with t in (0, |e - o|), ties -> lowest primitive (orc_scene.h);
//   P = o + d t; clip = view_proj P; depth = clip.z / clip.w (sky unless 0 <= depth < 1); linear z = clip.z (g_buffer.frag:107);
//   N = normalize(barycentric vertex normals); motion = prev_uv - cur_uv from prev_view_proj P (frag :55-67);
//   curvature = sqrt(max(|dNdx|^2, |dNdy|^2)) from the fine 2x2-quad differences of the per-pixel normals, 0 across mesh-id
//   boundaries and below 1e-4 (flat interpolants are exactly 0 in a rasteriser; the helper-lane extrapolation of dFdx at
//   triangle edges has no ray-cast counterpart — documented deviation).
#include "orc_shading.h"
#ifdef _OPENMP
#include <omp.h>
#endif
#include <vector>

using namespace orc;

namespace {

// g_buffer.frag:47-51
inline void direction_to_octohedral(vec3 n, float out[2])
{
    const float inv = 1.0f / ((fabsf(n.x) + fabsf(n.y)) + fabsf(n.z));
    const float px = n.x * inv, py = n.y * inv;
    if (n.z > 0.0f) { out[0] = px; out[1] = py; }
    else
    {
        out[0] = (1.0f - fabsf(py)) * (px >= 0.0f ? 1.0f : -1.0f);
        out[1] = (1.0f - fabsf(px)) * (py >= 0.0f ? 1.0f : -1.0f);
    }
}

inline uint8_t unorm8(float v)
{
    const float s = v * 255.0f + 0.5f;
    return (uint8_t)(s < 0.0f ? 0.0f : (s > 255.0f ? 255.0f : s));
}

} // namespace

extern "C" void orc_gbuffer_render(void* shading_scene, const uint32_t* prim_inst, const hr_frame* f, int W, int H, uint8_t* gb1, uint16_t* gb2, uint16_t* gb3,
                                   float* depth)
{
    const ShadingScene& ss  = *(const ShadingScene*)shading_scene;
    const mat4          vpi = load_mat4(f->ubo.view_proj_inverse), vp = load_mat4(f->ubo.view_proj), pvp = load_mat4(f->ubo.prev_view_proj);
    std::vector<vec3>     nrm((size_t)W * H);
    std::vector<uint32_t> mid((size_t)W * H, 0xFFFFFFFFu);
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
        {
            const size_t pi = (size_t)y * W + x;
            const vec2   tc = { ((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H };
            const vec3   o = world_position_from_depth(tc, 0.0f, vpi), e = world_position_from_depth(tc, 1.0f, vpi);
            const vec3   dv  = e - o;
            const float  len = length(dv);
            const vec3   d   = dv * (1.0f / len);
            Hit          h;
            bool         hit = ss.scene->closest(o, d, 0.0f, len, h);
            uint16_t*    g2  = gb2 + 4 * pi;
            uint16_t*    g3  = gb3 + 4 * pi;
            if (hit)
            {
                const vec3  P  = o + d * h.t;
                const vec4  c  = mul(vp, vec4{ P.x, P.y, P.z, 1.0f }), pc = mul(pvp, vec4{ P.x, P.y, P.z, 1.0f });
                const float dz = c.z / c.w;
                if (!(dz >= 0.0f && dz < 1.0f)) hit = false;
                else
                {
                    const float* n  = ss.vnormals.data() + 9ull * h.prim;
                    const float  b0 = 1.0f - h.u - h.v, b1 = h.u, b2 = h.v;
                    const vec3   N  = normalize((vec3{ n[0], n[1], n[2] } * b0 + vec3{ n[3], n[4], n[5] } * b1) + vec3{ n[6], n[7], n[8] } * b2);
                    float        oct[2];
                    direction_to_octohedral(fetch_normal(ss, h.prim, b0, b1, b2, false, N), oct); // g_buffer.frag:100; compute_curvature keeps the interpolated normal
                    const float cu = (c.x / c.w) * 0.5f + 0.5f, cv = (c.y / c.w) * 0.5f + 0.5f;
                    const float pu = (pc.x / pc.w) * 0.5f + 0.5f, pv = (pc.y / pc.w) * 0.5f + 0.5f;
                    const hr_material& m = ss.materials[ss.prim_mat[h.prim]];
                    vec3  albedo = { m.albedo[0], m.albedo[1], m.albedo[2] };
                    float roughness = m.roughness, metallic = m.metallic;
                    fetch_material(ss, h.prim, b0, b1, b2, albedo, roughness, metallic); // g_buffer.frag:90-105 (textures bound: fetch_* at the hit, mip 0)
                    g2[0] = f2h(oct[0]); g2[1] = f2h(oct[1]); g2[2] = f2h(pu - cu); g2[3] = f2h(pv - cv);
                    g3[0] = f2h(roughness); g3[1] = 0; g3[2] = f2h((float)prim_inst[h.prim]); g3[3] = f2h(c.z);
                    depth[pi] = dz;
                    if (gb1)
                    {
                        uint8_t* g1 = gb1 + 4 * pi;
                        g1[0] = unorm8(albedo.x); g1[1] = unorm8(albedo.y); g1[2] = unorm8(albedo.z); g1[3] = unorm8(metallic);
                    }
                    nrm[pi] = N;
                    mid[pi] = prim_inst[h.prim];
                }
            }
            if (!hit)
            { // clear values, g_buffer.cpp:72-96: GB3 = (0,0,0,-1), depth = 1
                g2[0] = g2[1] = g2[2] = g2[3] = 0;
                g3[0] = g3[1] = g3[2] = 0;
                g3[3] = f2h(-1.0f);
                depth[pi] = 1.0f;
                if (gb1) memset(gb1 + 4 * pi, 0, 4);
            }
        }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
        {
            const size_t pi = (size_t)y * W + x;
            if (mid[pi] == 0xFFFFFFFFu) continue;
            const int    x0 = x & ~1, y0 = y & ~1, x1 = x0 + 1 < W ? x0 + 1 : W - 1, y1 = y0 + 1 < H ? y0 + 1 : H - 1;
            const size_t ax = (size_t)y * W + x0, bx = (size_t)y * W + x1, ay = (size_t)y0 * W + x, by = (size_t)y1 * W + x;
            float        cx = 0.0f, cy = 0.0f;
            if (mid[ax] == mid[pi] && mid[bx] == mid[pi]) { const vec3 dd = nrm[bx] - nrm[ax]; cx = dot(dd, dd); }
            if (mid[ay] == mid[pi] && mid[by] == mid[pi]) { const vec3 dd = nrm[by] - nrm[ay]; cy = dot(dd, dd); }
            float curv = sqrtf(fmaxf(cx, cy));
            if (curv < 1e-4f) curv = 0.0f;
            gb3[4 * pi + 1] = f2h(curv);
        }
}
