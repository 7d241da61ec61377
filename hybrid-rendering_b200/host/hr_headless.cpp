// hr_headless.cpp — headless frame loop in C++ on the host classes (the analogue of HybridRendering::update,
// src/main.cpp:49-129): update_uniforms -> build_tlas -> GBuffer -> Shadows -> AO -> end_frame, no window / swapchain.
// Usage: hr_headless [width height frames tris]
#include "hybrid_rendering.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

int main(int argc, char** argv)
{
    const int W = argc > 1 ? atoi(argv[1]) : 1920, H = argc > 2 ? atoi(argv[2]) : 1080, frames = argc > 3 ? atoi(argv[3]) : 40;
    const int tris = argc > 4 ? atoi(argv[4]) : 262144;
    try
    {
        hr::CommonResources common(0, W, H);
        std::vector<uint8_t> sobol(256 * 4), sr(128 * 128 * 4);
        hrs_blue_noise(1234, sobol.data(), sr.data());
        common.set_blue_noise(sobol.data(), sr.data());
        hrs_scene* scene = hrs_scene_create(HRS_SCENE_ARCADE, tris, 7);
        common.load_scene(scene);
        hr::GBuffer          g_buffer(&common);
        hr::RayTracedShadows shadows(&common, &g_buffer, hr::RAY_TRACE_SCALE_FULL_RES);
        hr::RayTracedAO      ao(&common, &g_buffer, hr::RAY_TRACE_SCALE_HALF_RES);

        const size_t px = (size_t)W * H;
        uint8_t*  gb1;
        uint16_t *gb2, *gb3;
        float*    depth;
        cudaMallocHost((void**)&gb1, px * 4);
        cudaMallocHost((void**)&gb2, px * 8);
        cudaMallocHost((void**)&gb3, px * 8);
        cudaMallocHost((void**)&depth, px * 4);
        hrs_light_desc light;
        hrs_default_light(&light);
        light.rot_x_deg = 25.0f;
        const float pos[3] = { 0.0f, 9.0f, -4.0f }, tgt[3] = { 2.0f, 7.0f, 60.0f };
        cudaStream_t st;
        cudaStreamCreate(&st);
        hr_gbuffer_desc desc { W, H, gb1, gb2, gb3, depth };
        double gpu_ms = 0.0;
        for (int i = 0; i < frames; i++)
        {
            common.update_uniforms(pos, tgt, &light);
            if (i < 2) hrs_write_gbuffer(scene, &common.frame, W, H, gb1, gb2, gb3, depth); // static camera: frames >= 1 share the G-buffer
            hr_scene_rebuild(common.current_scene(), st);                                  // build_tlas every frame (main.cpp:74)
            auto t0 = std::chrono::steady_clock::now();
            g_buffer.render(&desc, st);
            shadows.render(st);
            ao.render(st);
            cudaStreamSynchronize(st);
            gpu_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            common.end_frame();
        }
        hr_image s = shadows.output_ds(), a = ao.output_ds();
        printf("frames=%d last frame %.3f ms (upload + shadows + ao); shadows out %dx%d fmt %d, ao out %dx%d fmt %d\n", frames, gpu_ms, s.width, s.height,
               s.format, a.width, a.height, a.format);
        hrs_scene_destroy(scene);
        cudaFreeHost(gb1); cudaFreeHost(gb2); cudaFreeHost(gb3); cudaFreeHost(depth);
    }
    catch (const std::exception& e)
    {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
