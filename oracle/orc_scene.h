// oracle/orc_scene.h — TEST INFRASTRUCTURE ONLY (CPU oracle). Never linked into the product library.
//
// Ray queries of the oracle.  The reference traverses a driver-built Vulkan TLAS/BLAS with
// rayQueryEXT / traceRayEXT (src/shaders/ray_query.glsl:6-59; build at
// external/dwSampleFramework/extras/ray_traced_scene.cpp:196-248) — a third-party dependency with no
// source; only hit/no-hit and hit-t semantics can be pinned:
//   any-hit:  true iff some triangle has t_min < t < t_max   (flags Opaque|TerminateOnFirstHit, no culling,
//             instances have cull disabled ray_traced_scene.cpp:606; extra `t < t_max` of ray_query.glsl:56)
//   closest:  minimum such t; ties broken by lowest primitive index (our definition).
// Ground truth here is the brute-force loop over all triangles with the ray/triangle routine below
// (Moeller-Trumbore, fixed fp32 operation order, no contraction).  `RefBVH` is an independent
// median-split BVH used only to make large scenes tractable; tests check it against brute force.
#pragma once
#include "orc_math.h"
#include <algorithm>
#include <cfloat>
#include <numeric>
#include <vector>

namespace orc {

struct Tri { vec3 v0, e1, e2; }; // e1 = v1 - v0, e2 = v2 - v0 (fp32 subtraction)

// Returns true and t/u/v if the ray hits with t_min < t < t_max.
inline bool ray_triangle(const Tri& tri, vec3 o, vec3 d, float t_min, float t_max, float& t, float& u, float& v)
{
    vec3  p   = cross(d, tri.e2);
    float det = dot(tri.e1, p);
    if (det == 0.0f) return false;
    float inv = 1.0f / det;
    vec3  tv  = o - tri.v0;
    u         = dot(tv, p) * inv;
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    vec3 q = cross(tv, tri.e1);
    v      = dot(d, q) * inv;
    if (!(v >= 0.0f && u + v <= 1.0f)) return false;
    t = dot(tri.e2, q) * inv;
    return t > t_min && t < t_max;
}

struct Hit { float t; uint32_t prim; float u, v; };

struct RefBVH {
    struct Node { float lo[3], hi[3]; int left, right; int first, count; }; // leaf if count > 0
    std::vector<Tri>      tris;
    std::vector<uint32_t> order; // leaf slot -> primitive index
    std::vector<Node>     nodes;

    void build(const float* verts9, size_t n)
    {
        tris.resize(n);
        std::vector<vec3> cmin(n), cmax(n), cen(n);
        for (size_t i = 0; i < n; i++)
        {
            const float* p = verts9 + 9 * i;
            vec3 a = { p[0], p[1], p[2] }, b = { p[3], p[4], p[5] }, c = { p[6], p[7], p[8] };
            tris[i] = { a, b - a, c - a };
            cmin[i] = { std::min({ a.x, b.x, c.x }), std::min({ a.y, b.y, c.y }), std::min({ a.z, b.z, c.z }) };
            cmax[i] = { std::max({ a.x, b.x, c.x }), std::max({ a.y, b.y, c.y }), std::max({ a.z, b.z, c.z }) };
            cen[i]  = (cmin[i] + cmax[i]) * 0.5f;
        }
        order.resize(n);
        std::iota(order.begin(), order.end(), 0u);
        nodes.clear();
        nodes.reserve(2 * n / 2 + 16);
        float ext = 0;
        {
            vec3 lo = { FLT_MAX, FLT_MAX, FLT_MAX }, hi = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
            for (size_t i = 0; i < n; i++)
            {
                lo = { std::min(lo.x, cmin[i].x), std::min(lo.y, cmin[i].y), std::min(lo.z, cmin[i].z) };
                hi = { std::max(hi.x, cmax[i].x), std::max(hi.y, cmax[i].y), std::max(hi.z, cmax[i].z) };
            }
            if (n) ext = std::max({ hi.x - lo.x, hi.y - lo.y, hi.z - lo.z });
        }
        pad = 1e-4f * ext + 1e-6f;
        if (n) build_rec(0, (int)n, cmin, cmax, cen);
    }

    float pad = 0;

    int build_rec(int first, int last, const std::vector<vec3>& cmin, const std::vector<vec3>& cmax, const std::vector<vec3>& cen)
    {
        int  idx = (int)nodes.size();
        nodes.push_back({});
        vec3 lo = { FLT_MAX, FLT_MAX, FLT_MAX }, hi = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
        vec3 clo = lo, chi = hi;
        for (int i = first; i < last; i++)
        {
            uint32_t t = order[i];
            lo  = { std::min(lo.x, cmin[t].x), std::min(lo.y, cmin[t].y), std::min(lo.z, cmin[t].z) };
            hi  = { std::max(hi.x, cmax[t].x), std::max(hi.y, cmax[t].y), std::max(hi.z, cmax[t].z) };
            clo = { std::min(clo.x, cen[t].x), std::min(clo.y, cen[t].y), std::min(clo.z, cen[t].z) };
            chi = { std::max(chi.x, cen[t].x), std::max(chi.y, cen[t].y), std::max(chi.z, cen[t].z) };
        }
        Node nd;
        nd.lo[0] = lo.x - pad; nd.lo[1] = lo.y - pad; nd.lo[2] = lo.z - pad;
        nd.hi[0] = hi.x + pad; nd.hi[1] = hi.y + pad; nd.hi[2] = hi.z + pad;
        nd.left = nd.right = -1;
        nd.first = first;
        nd.count = 0;
        int n = last - first;
        if (n <= 4) { nd.count = n; nodes[idx] = nd; return idx; }
        float ex = chi.x - clo.x, ey = chi.y - clo.y, ez = chi.z - clo.z;
        int   ax = (ex >= ey && ex >= ez) ? 0 : (ey >= ez ? 1 : 2);
        int   mid = (first + last) / 2;
        auto  key = [&](uint32_t t) { return ax == 0 ? cen[t].x : (ax == 1 ? cen[t].y : cen[t].z); };
        std::nth_element(order.begin() + first, order.begin() + mid, order.begin() + last, [&](uint32_t a, uint32_t b) { return key(a) < key(b); });
        nodes[idx] = nd;
        int l = build_rec(first, mid, cmin, cmax, cen);
        int r = build_rec(mid, last, cmin, cmax, cen);
        nodes[idx].left  = l;
        nodes[idx].right = r;
        return idx;
    }

    static bool slab(const Node& n, vec3 o, vec3 inv, float t_min, float t_max)
    {
        float t0 = t_min, t1 = t_max;
        const float oo[3] = { o.x, o.y, o.z }, ii[3] = { inv.x, inv.y, inv.z };
        for (int a = 0; a < 3; a++)
        {
            float ta = (n.lo[a] - oo[a]) * ii[a], tb = (n.hi[a] - oo[a]) * ii[a];
            float tn = fminf(ta, tb), tf = fmaxf(ta, tb); // NaN (0*inf) dropped by fmin/fmax => no constraint
            t0 = fmaxf(t0, tn);
            t1 = fminf(t1, tf);
        }
        return t0 <= t1 * 1.0000005f + 1e-30f;
    }

    bool any_hit(vec3 o, vec3 d, float t_min, float t_max) const
    {
        if (nodes.empty()) return false;
        vec3 inv = { 1.0f / d.x, 1.0f / d.y, 1.0f / d.z };
        int  stack[128], sp = 0;
        stack[sp++] = 0;
        while (sp)
        {
            const Node& n = nodes[stack[--sp]];
            if (!slab(n, o, inv, t_min, t_max)) continue;
            if (n.count)
            {
                for (int i = 0; i < n.count; i++)
                {
                    float t, u, v;
                    if (ray_triangle(tris[order[n.first + i]], o, d, t_min, t_max, t, u, v)) return true;
                }
            }
            else { stack[sp++] = n.left; stack[sp++] = n.right; }
        }
        return false;
    }

    bool closest_hit(vec3 o, vec3 d, float t_min, float t_max, Hit& best) const
    {
        best = { t_max, 0xFFFFFFFFu, 0, 0 };
        if (nodes.empty()) return false;
        vec3 inv = { 1.0f / d.x, 1.0f / d.y, 1.0f / d.z };
        int  stack[128], sp = 0;
        stack[sp++] = 0;
        while (sp)
        {
            const Node& n = nodes[stack[--sp]];
            // prune with <= best.t (ties must still be visited to apply the lowest-primitive rule)
            if (!slab(n, o, inv, t_min, best.t)) continue;
            if (n.count)
            {
                for (int i = 0; i < n.count; i++)
                {
                    float    t, u, v;
                    uint32_t prim = order[n.first + i];
                    if (ray_triangle(tris[prim], o, d, t_min, t_max, t, u, v))
                        if (t < best.t || (t == best.t && prim < best.prim)) best = { t, prim, u, v };
                }
            }
            else { stack[sp++] = n.left; stack[sp++] = n.right; }
        }
        return best.prim != 0xFFFFFFFFu;
    }

    bool any_hit_brute(vec3 o, vec3 d, float t_min, float t_max) const
    {
        for (size_t i = 0; i < tris.size(); i++)
        {
            float t, u, v;
            if (ray_triangle(tris[i], o, d, t_min, t_max, t, u, v)) return true;
        }
        return false;
    }
    bool closest_hit_brute(vec3 o, vec3 d, float t_min, float t_max, Hit& best) const
    {
        best = { t_max, 0xFFFFFFFFu, 0, 0 };
        for (size_t i = 0; i < tris.size(); i++)
        {
            float t, u, v;
            if (ray_triangle(tris[i], o, d, t_min, t_max, t, u, v))
                if (t < best.t) best = { t, (uint32_t)i, u, v }; // ascending i => lowest primitive wins ties
        }
        return best.prim != 0xFFFFFFFFu;
    }
};

// ray_query.glsl:34-59 query_distance / :6-30 query_visibility — both reduce to "1 if no hit in (0.01, t_max)".
struct Scene {
    RefBVH bvh;
    bool   brute = false;
    float  query_visibility(vec3 world_pos, vec3 direction, float t_max) const
    {
        const float t_min = 0.01f;
        bool        hit   = brute ? bvh.any_hit_brute(world_pos, direction, t_min, t_max) : bvh.any_hit(world_pos, direction, t_min, t_max);
        return hit ? 0.0f : 1.0f;
    }
    bool closest(vec3 o, vec3 d, float t_min, float t_max, Hit& h) const
    {
        return brute ? bvh.closest_hit_brute(o, d, t_min, t_max, h) : bvh.closest_hit(o, d, t_min, t_max, h);
    }
};

} // namespace orc
