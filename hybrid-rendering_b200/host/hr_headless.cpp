// hr_headless.cpp — headless frame loop in C++ on the host classes (the analogue of HybridRendering::update,
// src/main.cpp:49-129): update_uniforms -> build_tlas -> GBuffer -> Shadows -> AO -> DDGI -> Reflections -> DeferredShading ->
// TemporalAA -> ToneMap -> end_frame, no window / swapchain / Vulkan: the G-buffer is ray cast on the device, the host only sends the per-frame constants.
// Usage: hr_headless [--post [--png out.png]] [width height frames tris [mesh.gltf|mesh.glb|mesh.obj]]   --post: Halton-jittered camera + TemporalAA + ToneMap behind
// the deferred combine;   (a mesh file replaces the procedural arcade; it is drawn once
// with an identity transform, like the single-instance scenes of src/common.cpp:340-534)
#include "hybrid_rendering.h"
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <memory>
#include <string>
#include <vector>

int main(int argc, char** argv)
{
    bool        post = false;
    std::string png_path; // --png FILE (with --post): write the tone-mapped frame as a PNG (hra_image_save_png)
    for (int i = 1; i < argc; i++)
    {
        const std::string a = argv[i];
        const int drop = a == "--post" ? 1 : (a == "--png" && i + 1 < argc) ? 2 : 0;
        if (!drop) continue;
        if (drop == 1) post = true;
        else png_path = argv[i + 1];
        for (int k = i; k + drop < argc; k++) argv[k] = argv[k + drop];
        argc -= drop;
        i--;
    }
    const int W = argc > 1 ? atoi(argv[1]) : 1920, H = argc > 2 ? atoi(argv[2]) : 1080, frames = argc > 3 ? atoi(argv[3]) : 40;
    const int tris = argc > 4 ? atoi(argv[4]) : 262144;
    try
    {
        hr::CommonResources common(0, W, H);
        std::vector<uint8_t> sobol(256 * 4), sr(128 * 128 * 4);
        hrs_blue_noise(1234, sobol.data(), sr.data());
        common.set_blue_noise(sobol.data(), sr.data());
        std::vector<uint16_t> lut(512 * 512 * 2);
        hrs_brdf_lut(64, lut.data());
        hr::check(common.ctx, hr_brdf_lut_set(common.ctx, lut.data()), "hr_brdf_lut_set");
        hrs_scene* scene = nullptr;
        if (argc > 5)
        {
            hra_mesh* mesh = nullptr;
            if (hra_mesh_load(argv[5], &mesh) != HRA_OK) throw std::runtime_error(hra_last_error());
            hra_scene*  as = hra_scene_create();
            const float ident[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
            hra_scene_add_instance(as, mesh, ident);
            hra_scene_finalize(as);
            common.load_scene(as);
            hra_scene_destroy(as);
            hra_mesh_destroy(mesh);
        }
        else
        {
            scene = hrs_scene_create(HRS_SCENE_ARCADE, tris, 7);
            common.load_scene(scene);
        }
        hr::GBuffer              g_buffer(&common);
        hr::RayTracedShadows     shadows(&common, &g_buffer, hr::RAY_TRACE_SCALE_FULL_RES);
        hr::RayTracedAO          ao(&common, &g_buffer, hr::RAY_TRACE_SCALE_HALF_RES);
        hr::DDGI                 ddgi(&common, &g_buffer, hr::RAY_TRACE_SCALE_FULL_RES);
        hr::RayTracedReflections reflections(&common, &g_buffer, hr::RAY_TRACE_SCALE_HALF_RES);
        hr::DeferredShading      deferred(&common, &g_buffer);
        std::unique_ptr<hr::TemporalAA> temporal_aa_p(post ? new hr::TemporalAA(&common, &g_buffer) : nullptr); // only with --post
        std::unique_ptr<hr::ToneMap>    tone_map_p(post ? new hr::ToneMap(&common) : nullptr);
        ddgi.set_probe_distance(8.0f);
        ddgi.set_normal_bias(0.5f);
        const float sky[3] = { 0.3f, 0.4f, 0.6f };
        for (int k = 0; k < 3; k++) ddgi.params.sky_color[k] = reflections.params.sky_color[k] = deferred.params.env_color[k] = sky[k];

        hrs_light_desc light;
        hrs_default_light(&light);
        light.rot_x_deg = 25.0f;
        const float tgt[3] = { 2.0f, 7.0f, 60.0f };
        const float rot[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
        cudaStream_t st;
        cudaStreamCreate(&st);
        double gpu_ms = 0.0;
        for (int i = 0; i < frames; i++)
        {
            const float pos[3] = { 0.02f * (float)i, 9.0f, -4.0f }; // slow lateral pan
            if (post) temporal_aa_p->update(); // main.cpp:1025, before update_uniforms
            common.update_uniforms(pos, tgt, &light);
            if (post) temporal_aa_p->apply_jitter();
            hr_scene_rebuild(common.current_scene(), st); // build_tlas every frame (main.cpp:74)
            auto t0 = std::chrono::steady_clock::now();
            g_buffer.render(st);
            shadows.render(st);
            ao.render(st);
            ddgi.render(rot, st);
            reflections.render(st, &ddgi);
            deferred.render(st, shadows.handle(), ao.handle(), reflections.handle(), ddgi.handle());
            if (post)
            {
                temporal_aa_p->render(st, deferred.handle());
                tone_map_p->render(st, temporal_aa_p.get(), deferred.handle());
            }
            cudaStreamSynchronize(st);
            gpu_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            common.end_frame();
        }
        hr_image o = deferred.output_ds();
        std::vector<__half> px((size_t)o.width * o.height * 4);
        cudaMemcpy(px.data(), o.data, px.size() * sizeof(__half), cudaMemcpyDeviceToHost);
        double sum = 0.0;
        bool   finite = true;
        for (size_t k = 0; k < px.size(); k += 4)
            for (int c = 0; c < 3; c++) { const float v = __half2float(px[k + c]); finite = finite && std::isfinite(v); sum += v; }
        printf("frames=%d last frame %.3f ms (g-buffer + shadows + ao + ddgi + reflections + deferred); output %dx%d fmt %d mean %.5f finite %d\n", frames, gpu_ms, o.width,
               o.height, o.format, sum / (3.0 * o.width * o.height), finite ? 1 : 0);
        bool post_ok = true;
        if (post)
        { // the presented image: RGBA8 from the tone map over the TAA resolve
            hr_image t = tone_map_p->output();
            std::vector<uint8_t> ldr((size_t)t.width * t.height * 4);
            cudaMemcpy(ldr.data(), t.data, ldr.size(), cudaMemcpyDeviceToHost);
            double lsum = 0.0;
            bool   opaque = true;
            for (size_t k = 0; k < ldr.size(); k += 4) { lsum += ldr[k] + ldr[k + 1] + ldr[k + 2]; opaque = opaque && ldr[k + 3] == 255; }
            printf("tone-mapped TAA output %dx%d fmt %d mean %.3f / 255, alpha opaque %d\n", t.width, t.height, t.format, lsum / (3.0 * t.width * t.height), opaque ? 1 : 0);
            post_ok = opaque && lsum > 0.0 && t.format == HR_FMT_RGBA8;
            if (!png_path.empty())
            {
                if (hra_image_save_png(png_path.c_str(), t.width, t.height, 4, ldr.data()) != HRA_OK) throw std::runtime_error(hra_last_error());
                printf("wrote %s\n", png_path.c_str());
            }
        }
        if (scene) hrs_scene_destroy(scene);
        if (!finite || !(sum > 0.0)) return 2;
        if (!post_ok) return 3;
    }
    catch (const std::exception& e)
    {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
