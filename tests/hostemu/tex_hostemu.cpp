// tests/hostemu/tex_hostemu.cpp — TEST INFRASTRUCTURE ONLY.  Host build of the product's material-texture functions
// (hybrid-rendering_b200/csrc/tex_px.cuh: what the TEX instantiations of the hit-shading / G-buffer kernels call), compared with the
// oracle's independent Texture2D / fetch_material (oracle/orc_shading.h) by tests/test_textures_cpu.py.  Not a fallback: nothing in the
// product library, pyhr or bench.py links or loads it.
#include "../../hybrid-rendering_b200/csrc/tex_px.cuh"
#include <cstddef>

extern "C" {

// sample n uv pairs of texture `ti`: out = 4 floats each
__attribute__((visibility("default"))) void emu_tex_sample(const uint32_t* texels, const tex::TexDesc* desc, int n_textures, const float* srgb_lut, int ti, const float* uv,
                                                           size_t n, float* out4)
{
    tex::TexDev T { texels, desc, nullptr, nullptr, srgb_lut, n_textures };
    for (size_t i = 0; i < n; i++)
    {
        const tex::RGBA c = tex::sample(T, ti, uv[2 * i], uv[2 * i + 1]);
        out4[4 * i] = c.r; out4[4 * i + 1] = c.g; out4[4 * i + 2] = c.b; out4[4 * i + 3] = c.a;
    }
}
// material_at_hit for n hits (prim, u, v): in/out arrays of albedo rgb (3 floats), roughness, metallic initialised with the constants
__attribute__((visibility("default"))) void emu_material_at_hit(const uint32_t* texels, const tex::TexDesc* desc, int n_textures, const float* srgb_lut, const tex::MatTex* mat,
                                                                const float* vuv, const uint32_t* prim_mat, const uint32_t* prim, const float* bary_uv, size_t n,
                                                                float* albedo3, float* roughness, float* metallic)
{
    tex::TexDev T { texels, desc, mat, vuv, srgb_lut, n_textures };
    for (size_t i = 0; i < n; i++)
    {
        const float u = bary_uv[2 * i], v = bary_uv[2 * i + 1];
        tex::material_at_hit(T, prim_mat[prim[i]], prim[i], 1.0f - u - v, u, v, albedo3[3 * i], albedo3[3 * i + 1], albedo3[3 * i + 2], roughness[i], metallic[i]);
    }
}

// normal_at_hit for n hits: in/out arrays of the interpolated unit normal (3 floats each)
__attribute__((visibility("default"))) void emu_normal_at_hit(const uint32_t* texels, const tex::TexDesc* desc, int n_textures, const float* srgb_lut, const tex::MatTex* mat,
                                                              const float* vuv, const float* vtb, const uint32_t* prim_mat, const uint32_t* prim, const float* bary_uv, size_t n,
                                                              int tangent_as_bitangent, float* normal3)
{
    tex::TexDev T { texels, desc, mat, vuv, srgb_lut, n_textures, vtb };
    for (size_t i = 0; i < n; i++)
    {
        const float u = bary_uv[2 * i], v = bary_uv[2 * i + 1];
        tex::normal_at_hit(T, prim_mat[prim[i]], prim[i], 1.0f - u - v, u, v, tangent_as_bitangent != 0, normal3[3 * i], normal3[3 * i + 1], normal3[3 * i + 2]);
    }
}

} // extern "C"
