// oracle/orc_path_trace.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle). Never linked into the product library.
//
// CPU restatement of the ground-truth progressive path tracer:
//   src/shaders/ground_truth/ground_truth_path_trace.rgen:52-111   ray generation (RNG-jittered pixel), accumulation
//   src/shaders/ground_truth/ground_truth_path_trace.rchit:64-141  hit shading; indirect_lighting's traceRayEXT is COMMENTED OUT (:92-104),
//                                                                   so it returns p_IndirectPayload.L = vec3(0) whatever the RNG draws
//   src/shaders/ground_truth/ground_truth_path_trace.rmiss:27-35   sky (constant colour here: the cubemap is an asset)
//   src/shaders/lighting.glsl:117-196 with SOFT_SHADOWS, RAY_THROUGHPUT, SAMPLE_SKY_LIGHT (rchit:13-16)
// Geometry-deciding arithmetic (primary ray, hit point, normal, the two shadow rays) follows the deterministic rules of orc_math.h;
// colours are ordinary fp32 (tolerance-checked).  Parity unpinned in the sense of DESIGN.md §2; literals pinned by tests/test_ref_constants.py.
#include "orc_shading.h"
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

namespace {

inline vec2 next_vec2(RNG& rng) { vec2 r; r.x = next_float(rng); r.y = next_float(rng); return r; } // random.glsl:64-67, left to right

// direct_lighting, lighting.glsl:117-196: SOFT_SHADOWS (rng1), RAY_THROUGHPUT (T), SAMPLE_SKY_LIGHT (rng2, sky)
vec3 direct_lighting_path(const Scene& scene, const hr_light& light, vec3 Wo, vec3 N, vec3 P, vec3 F0, vec3 diffuse_color, float roughness, vec3 T, vec2 rng1, vec2 rng2,
                          vec3 sky)
{
    vec3 Lo         = { 0, 0, 0 };
    vec3 ray_origin = P + N * 0.1f; // :143
    { // punctual light
        vec3  Li = light_color(light) * light_intensity(light); // fetch_light_properties :35
        vec3  Wi;
        float t_max, attenuation;
        fetch_light_properties_shadow(light, P, N, rng1, Wi, t_max, attenuation); // the SOFT_SHADOWS branches (:37-105) + :110
        vec3 Wh = normalize(Wo + Wi);                                              // :108
        if (attenuation > 0.0f) attenuation *= scene.query_visibility(ray_origin, Wi, t_max); // query_distance :172
        vec3 brdf = evaluate_uber_brdf(diffuse_color, roughness, N, F0, Wo, Wh, Wi);
        Lo = Lo + ((T * brdf) * attenuation) * Li; // :176
    }
    { // sky light :180-192
        vec3 Wi = sample_cosine_lobe(N, rng2);
        vec3 Li = sky;
        vec3 Wh = normalize(Wo + Wi);
        Li      = Li * scene.query_visibility(ray_origin, Wi, 10000.0f);
        vec3 brdf = evaluate_uber_brdf(diffuse_color, roughness, N, F0, Wo, Wh, Wi);
        Lo = Lo + (T * brdf) * Li;
    }
    return Lo;
}

} // namespace

// One render of GroundTruthPathTracer (one sample per pixel).  prev: the image written by the previous render (ignored when num_frames == 0);
// out: RGBA16F; out_prim (may be NULL): primitive hit by the primary ray, 0xFFFFFFFF = miss.
extern "C" void orc_path_trace(void* shading_scene, const hr_frame* f, int W, int H, uint32_t num_frames, uint32_t max_ray_bounces, float roughness_multiplier,
                               const float* sky3, const uint16_t* prev, uint16_t* out, uint32_t* out_prim)
{
    const ShadingScene& ss  = *(const ShadingScene*)shading_scene;
    const mat4 view_inverse = load_mat4(f->ubo.view_inverse), proj_inverse = load_mat4(f->ubo.proj_inverse);
    const vec3 sky = { sky3[0], sky3[1], sky3[2] };
    const float RADIANCE_CLAMP_COLOR = 1.0f; // common.glsl:19
#pragma omp parallel for schedule(dynamic, 2)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
        {
            // rgen:56-75
            vec3 L = { 0, 0, 0 }, T = { 1, 1, 1 };
            uint32_t depth = 0;
            RNG  rng = rng_init((uint32_t)x, (uint32_t)y, num_frames);
            const vec2 pixel_coord = { (float)x + 0.5f, (float)y + 0.5f };
            const vec2 jitter = next_vec2(rng); // vec2(next_float, next_float): arguments evaluate left to right
            const vec2 jittered_coord = pixel_coord + jitter;
            const vec2 tex_coord = { jittered_coord.x / (float)W, jittered_coord.y / (float)H };
            const vec2 ndc = { tex_coord.x * 2.0f - 1.0f, tex_coord.y * 2.0f - 1.0f };
            const vec4 origin = mul(view_inverse, vec4{ 0.0f, 0.0f, 0.0f, 1.0f });
            const vec4 target = mul(proj_inverse, vec4{ ndc.x, ndc.y, 1.0f, 1.0f });
            const vec3 tn = normalize(vec3{ target.x, target.y, target.z });
            const vec4 direction = mul(view_inverse, vec4{ tn.x, tn.y, tn.z, 0.0f });
            const vec3 o = { origin.x, origin.y, origin.z }, d = { direction.x, direction.y, direction.z };
            Hit h;
            uint32_t prim = 0xFFFFFFFFu;
            if (ss.scene->closest(o, d, 0.001f, 10000.0f, h))
            { // rchit main :112-141
                prim = h.prim;
                Surface s = fetch_surface(ss, h);
                const float roughness = s.roughness * roughness_multiplier; // fetch_roughness (max(., MIN_ROUGHNESS)) * u_PushConstants.roughness_multiplier
                const vec3 N = s.N, Wo = -d;
                const vec3 F0 = mix3(vec3{ 0.04f, 0.04f, 0.04f }, s.albedo, s.metallic);
                const vec3 c_diffuse = mix3(s.albedo * (vec3{ 1, 1, 1 } - F0), vec3{ 0, 0, 0 }, s.metallic);
                const vec2 rng1 = next_vec2(rng), rng2 = next_vec2(rng);
                L = L + direct_lighting_path(*ss.scene, f->ubo.light, Wo, N, s.P, F0, c_diffuse, roughness, T, rng1, rng2, sky);
                if ((depth + 1) < max_ray_bounces) L = L + vec3{ 0, 0, 0 }; // indirect_lighting: vec3(0) (Russian roulette) or p_IndirectPayload.L, which nothing writes
            }
            else L = sky; // rmiss: depth == 0 -> L = environment sample
            // rgen:94-111
            const vec3 clamped = { fminf(L.x, RADIANCE_CLAMP_COLOR), fminf(L.y, RADIANCE_CLAMP_COLOR), fminf(L.z, RADIANCE_CLAMP_COLOR) };
            vec3 final_color = clamped;
            if (num_frames != 0)
            {
                const uint16_t* p = prev + 4 * ((size_t)y * W + x);
                const vec3 prev_color = { h2f(p[0]), h2f(p[1]), h2f(p[2]) };
                const float n = (float)num_frames;
                final_color = { prev_color.x + (clamped.x - prev_color.x) / n, prev_color.y + (clamped.y - prev_color.y) / n, prev_color.z + (clamped.z - prev_color.z) / n };
            }
            uint16_t* q = out + 4 * ((size_t)y * W + x);
            q[0] = f2h(final_color.x); q[1] = f2h(final_color.y); q[2] = f2h(final_color.z); q[3] = f2h(1.0f);
            if (out_prim) out_prim[(size_t)y * W + x] = prim;
        }
}
