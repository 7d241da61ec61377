// oracle/orc_math.h — TEST INFRASTRUCTURE ONLY (CPU oracle). Never linked into the product library.
//
// Small GLSL-like vector layer + the *deterministic* fp32 primitives the visibility-mask chain is
// specified in.  The reference's arithmetic runs in GLSL on a GPU driver (parity unpinned: no golden
// vectors exist); to make "bit-exact visibility mask" a testable statement the mask
// chain (ray generation + ray/triangle test) is specified as a fixed sequence of IEEE-754 binary32
// operations with no implicit contraction:
//   * + - * / sqrt are correctly rounded (compile with -ffp-contract=off; CUDA side -fmad=false)
//   * fmaf() only where written explicitly
//   * sin/cos = det_sincos() below (Cody-Waite reduction + fixed minimax polynomials, explicit fma)
// The CUDA kernels implement the same sequence independently (hybrid-rendering_b200/csrc/det_math.cuh).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <immintrin.h>

namespace orc {

struct vec2 { float x, y; };
struct vec3 { float x, y, z; };
struct vec4 { float x, y, z, w; };
struct ivec2 { int x, y; };
struct mat4 { float m[16]; }; // column-major: m[c*4+r]

inline vec2 operator+(vec2 a, vec2 b) { return { a.x + b.x, a.y + b.y }; }
inline vec2 operator-(vec2 a, vec2 b) { return { a.x - b.x, a.y - b.y }; }
inline vec2 operator*(vec2 a, float s) { return { a.x * s, a.y * s }; }
inline vec3 operator+(vec3 a, vec3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline vec3 operator-(vec3 a, vec3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline vec3 operator*(vec3 a, float s) { return { a.x * s, a.y * s, a.z * s }; }
inline vec3 operator*(vec3 a, vec3 b) { return { a.x * b.x, a.y * b.y, a.z * b.z }; }
inline vec3 operator-(vec3 a) { return { -a.x, -a.y, -a.z }; }

// dot = (x*x' + y*y') + z*z', no contraction
inline float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
inline vec3  cross(vec3 a, vec3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline float length(vec3 a) { return sqrtf(dot(a, a)); }
// normalize(v) = v * (1 / sqrt(dot(v,v)))
inline vec3  normalize(vec3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; } // GLSL mix
inline vec3  mix3(vec3 a, vec3 b, float t) { return a * (1.0f - t) + b * t; }
inline float fractf(float x) { return x - floorf(x); }
inline float stepf(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
// float -> int conversion, truncation toward zero, saturating, NaN -> 0 (what a GPU cvt.rzi does; plain C++ casts are UB there)
inline int f2i(float f)
{
    if (!(f == f)) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (int)0x80000000;
    return (int)f;
}

// mat4 * vec4, row r = ((m0r*x + m1r*y) + m2r*z) + m3r*w
inline vec4 mul(const mat4& M, vec4 v)
{
    vec4 r;
    r.x = ((M.m[0] * v.x + M.m[4] * v.y) + M.m[8] * v.z) + M.m[12] * v.w;
    r.y = ((M.m[1] * v.x + M.m[5] * v.y) + M.m[9] * v.z) + M.m[13] * v.w;
    r.z = ((M.m[2] * v.x + M.m[6] * v.y) + M.m[10] * v.z) + M.m[14] * v.w;
    r.w = ((M.m[3] * v.x + M.m[7] * v.y) + M.m[11] * v.z) + M.m[15] * v.w;
    return r;
}

// ---- fp16 storage emulation (RG16F / RGBA16F images; round-to-nearest-even like __float2half_rn) ----
inline uint16_t f2h(float f) { return (uint16_t)_cvtss_sh(f, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC); }
inline float    h2f(uint16_t h) { return _cvtsh_ss(h); }
inline float    round_h(float f) { return h2f(f2h(f)); }

// ---- deterministic sin/cos for x >= 0 (angles in [0, 2*pi] on this path) -------------------------
// k = floor(x*(2/pi) + 0.5); r = fma(-k, PIO2_HI, x); r = fma(-k, PIO2_LO, r);
// sin(r) ~ r + r*s*(S1 + s*(S2 + s*S3)),  cos(r) ~ 1 - s/2 + s*s*(C1 + s*(C2 + s*C3)),  s = r*r
inline void det_sincos(float x, float* sn, float* cs)
{
    const float TWO_OVER_PI = 0.636619772367581343f;
    const float PIO2_HI     = 1.57079625129699707031f;   // fp32(pi/2)
    const float PIO2_LO     = 7.54978941586159635335e-08f; // pi/2 - PIO2_HI
    float       kf          = floorf(x * TWO_OVER_PI + 0.5f);
    float       r           = fmaf(-kf, PIO2_HI, x);
    r                       = fmaf(-kf, PIO2_LO, r);
    float s                 = r * r;
    float ps                = fmaf(s, -1.9515295891e-4f, 8.3321608736e-3f);
    ps                      = fmaf(ps, s, -1.6666654611e-1f);
    float sr                = fmaf(r * s, ps, r);
    float pc                = fmaf(s, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc                      = fmaf(pc, s, 4.166664568298827e-2f);
    float cr                = fmaf(s * s, pc, fmaf(s, -0.5f, 1.0f));
    int   q                 = ((int)kf) & 3;
    float sv = (q & 1) ? cr : sr;
    float cv = (q & 1) ? sr : cr;
    if (q == 1 || q == 2) cv = -cv;
    if (q >= 2) sv = -sv;
    *sn = sv;
    *cs = cv;
}

} // namespace orc
