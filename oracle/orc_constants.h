// orc_constants.h — TEST INFRASTRUCTURE (CPU oracle).  Every literal the reference keeps in source for the hot path, as a
// named constant used by the oracle's code, so that tests/test_ref_constants.py can compare them one by one with the values
// parsed out of /root/reference by tests/golden/make_ref_constants.py (tests/golden/ref_constants.json).
#pragma once
#include <cstdint>
namespace orc_const {
constexpr float    M_PI_REF                               = 3.14159265359f; // common.glsl:16
constexpr float    EPSILON                                = 0.0001f;        // common.glsl:17
constexpr float    MIRROR_REFLECTIONS_ROUGHNESS_THRESHOLD = 0.05f;          // common.glsl:27
constexpr float    DDGI_REFLECTIONS_ROUGHNESS_THRESHOLD   = 0.75f;          // common.glsl:28
constexpr float    NORMAL_DISTANCE                        = 0.1f;           // reprojection.glsl:6
constexpr float    PLANE_DISTANCE                         = 5.0f;           // reprojection.glsl:7
constexpr float    MIN_ROUGHNESS                          = 0.1f;           // scene_descriptor_set.glsl:202
constexpr float    ATROUS_EPS_VARIANCE                    = 1e-10f;         // shadows_denoise_atrous.comp:99, reflections twin :99
constexpr float    ATROUS_KERNEL_WEIGHTS[3]               = { 1.0f, 2.0f / 3.0f, 1.0f / 6.0f };                             // :100
constexpr float    ATROUS_VARIANCE_KERNEL[2][2]           = { { 1.0f / 4.0f, 1.0f / 8.0f }, { 1.0f / 8.0f, 1.0f / 16.0f } }; // :69-72
// random.glsl:17-56
constexpr uint32_t RNG_STAR_MULTIPLIER = 0x9e3779bbu;
constexpr uint32_t RNG_ROTL_A = 26, RNG_SHIFT_B = 9, RNG_ROTL_C = 13;
constexpr uint32_t RNG_HASH_XOR0 = 61, RNG_HASH_SHR0 = 16, RNG_HASH_MUL0 = 9, RNG_HASH_SHR1 = 4, RNG_HASH_MUL1 = 0x27d4eb2du, RNG_HASH_SHR2 = 15;
constexpr uint32_t RNG_SEED_SHIFT = 16, RNG_FLOAT_ONE = 0x3f800000u, RNG_FLOAT_SHIFT = 9;
} // namespace orc_const
