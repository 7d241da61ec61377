// oracle/orc_deferred.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle). Never linked into the product library.
//
// CPU restatement of the deferred shading combine and the sky box drawn over it (DeferredShading::render = render_shading +
// render_skybox, src/deferred_shading.cpp:56-75): src/shaders/deferred.frag:146-205 (fresnel_schlick_roughness, indirect_lighting, main)
// and skybox.frag:18-22, with direct_lighting of lighting.glsl:117-196 in its raster variant (deferred.frag defines neither
// RAY_TRACING nor SOFT_SHADOWS: no shadow ray, the visibility comes from the shadows pass) and evaluate_uber_brdf (brdf.glsl:130-142).
// It consumes exactly the four pass outputs (:162 GI, :166 reflections, :187 shadows .r, :188 AO) plus the G-buffer.
// Environment: sky / prefiltered cubemaps and the irradiance SH are a constant colour c (assets absent), for which
// evaluate_sh9_irradiance(N) = c (band 0 only: c * 0.282095 * 4 pi * 0.282095 * pi / pi) and every prefiltered mip = c.
#include "orc_shading.h"
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

namespace {

// fetch_light_properties without RAY_TRACING / SOFT_SHADOWS (lighting.glsl:6-111)
inline void fetch_light_properties_raster(const hr_light& light, vec3 Wo, vec3 P, vec3 N, vec3& Li, vec3& Wi, vec3& Wh, float& attenuation)
{
    float t_max;
    fetch_light_properties_hard(light, P, N, Li, Wi, t_max, attenuation);
    Wh = normalize(Wo + Wi);
}

} // namespace

struct orc_gbuf_full { int32_t W, H; const uint8_t* gb1; const uint16_t* gb2; const uint16_t* gb3; const float* depth; };

// inputs may be NULL (push constant = 0): shadow R16F or RG16F (.r), ao R16F, reflections RGBA16F, gi RGBA16F — all full resolution
extern "C" void orc_deferred(const orc_gbuf_full* g, const hr_frame* f, const uint16_t* shadow, int shadow_channels, const uint16_t* ao, const uint16_t* reflections,
                             const uint16_t* gi, const float* env3, const uint16_t* brdf_lut, uint16_t* out)
{
    const int  W = g->W, H = g->H;
    const mat4 vpi = load_mat4(f->ubo.view_proj_inverse);
    const vec3 cam = { f->ubo.cam_pos[0], f->ubo.cam_pos[1], f->ubo.cam_pos[2] };
    const vec3 env = { env3[0], env3[1], env3[2] };
    BrdfLut    lut;
    lut.rg = brdf_lut;
    const float IndirectSpecularStrength = 2.0f; // deferred.frag:20
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
        {
            const size_t pi = (size_t)y * W + x;
            if (g->depth[pi] == 1.0f)
            { // DeferredShading::render_skybox (deferred_shading.cpp:69, 734-818): the cube drawn at gl_Position = clipPos.xyww (skybox.vert:33, depth 1)
              // covers exactly the pixels left at the G-buffer's clear depth; skybox.frag:18-22: vec4(texture(s_Cubemap, dir).rgb, 1) = the constant environment
                uint16_t* o = out + 4 * pi;
                o[0] = f2h(env.x); o[1] = f2h(env.y); o[2] = f2h(env.z); o[3] = f2h(1.0f);
                continue;
            }
            const vec2   tc = { ((float)x + 0.5f) / (float)W, ((float)y + 0.5f) / (float)H };
            const vec3   P  = world_position_from_depth(tc, g->depth[pi], vpi);
            const vec3   albedo   = { g->gb1[4 * pi] / 255.0f, g->gb1[4 * pi + 1] / 255.0f, g->gb1[4 * pi + 2] / 255.0f };
            const float  metallic = g->gb1[4 * pi + 3] / 255.0f;
            const float  roughness = h2f(g->gb3[4 * pi]);
            const float  visibility = shadow ? h2f(shadow[(size_t)shadow_channels * pi]) : 1.0f;
            const float  aov = ao ? h2f(ao[pi]) : 1.0f;
            const vec3   N  = octohedral_to_direction({ h2f(g->gb2[4 * pi]), h2f(g->gb2[4 * pi + 1]) });
            const vec3   Wo = normalize(cam - P);
            const vec3   F0 = mix3(vec3{ 0.04f, 0.04f, 0.04f }, albedo, metallic);
            const vec3   cd = mix3(albedo * (vec3{ 1, 1, 1 } - F0), vec3{ 0, 0, 0 }, metallic);
            vec3 Lo = { 0, 0, 0 };
            { // direct_lighting * visibility, :198
                vec3  Li, Wi, Wh;
                float att;
                fetch_light_properties_raster(f->ubo.light, Wo, P, N, Li, Wi, Wh, att);
                const vec3 brdf = evaluate_uber_brdf(cd, roughness, N, F0, Wo, Wh, Wi);
                Lo = Lo + ((brdf * att) * Li) * visibility;
            }
            { // indirect_lighting, :153-173
                const float ndv = fmaxf(dot(N, Wo), 0.0f);
                const vec3  F   = fresnel_schlick_roughness(ndv, F0, roughness);
                const vec3  kD  = (vec3{ 1, 1, 1 } - F) * (1.0f - metallic);
                const vec3  irradiance = gi ? vec3{ h2f(gi[4 * pi]), h2f(gi[4 * pi + 1]), h2f(gi[4 * pi + 2]) } : env;
                const vec3  diffuse    = irradiance * cd;
                const vec3  prefiltered = reflections ? vec3{ h2f(reflections[4 * pi]), h2f(reflections[4 * pi + 1]), h2f(reflections[4 * pi + 2]) } : env;
                vec3 specular = { 0, 0, 0 };
                if (lut.rg) specular = ibl_specular(lut, prefiltered, F, ndv, roughness, IndirectSpecularStrength);
                Lo = Lo + (kD * diffuse + specular) * aov;
            }
            uint16_t* o = out + 4 * pi;
            o[0] = f2h(Lo.x); o[1] = f2h(Lo.y); o[2] = f2h(Lo.z); o[3] = f2h(1.0f);
        }
}
