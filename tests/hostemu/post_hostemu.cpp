// tests/hostemu/post_hostemu.cpp — TEST INFRASTRUCTURE ONLY.  Compiles the product's per-pixel device functions
// (hybrid-rendering_b200/csrc/post_px.cuh, the code the CUDA kernels of post.cu execute) for the CPU and runs them over an image,
// so the `-m "not gpu"` suite can compare the kernels' arithmetic with the independently written oracle (oracle/orc_post.cpp).
// It is NOT a CPU fallback: nothing in the product library, bench.py or pyhr links or loads it — only tests/test_post_cpu.py does.
#include "../../hybrid-rendering_b200/csrc/post_px.cuh"

using namespace post;

static ImgView view(const uint16_t* p, int W, int H, int channels)
{
    ImgView v;
    v.p = p; v.W = W; v.H = H; v.channels = channels;
    return v;
}
static void store_h4(uint16_t* o, F4 c)
{
    o[0] = float_to_half_bits(c.x); o[1] = float_to_half_bits(c.y); o[2] = float_to_half_bits(c.z); o[3] = float_to_half_bits(c.w);
}

extern "C" {

// what launch_taa (post.cu) sets up, then k_taa's body for every pixel
__attribute__((visibility("default"))) void emu_taa(int W, int H, const uint16_t* cur, int cur_channels, const uint16_t* prev, const float* depth, const uint16_t* gb2,
                                                   const float* jitter_xy, float feedback_min, float feedback_max, int sharpen, uint16_t* out)
{
    TaaArgs A;
    A.cur = view(cur, W, H, cur_channels);
    A.prev = view(prev, W, H, 4);
    A.depth = depth; A.gb2 = gb2; A.W = W; A.H = H;
    A.texel_x = 1.0f / (float)W; A.texel_y = 1.0f / (float)H;
    A.jitter_x = jitter_xy[0]; A.jitter_y = jitter_xy[1];
    A.feedback_min = feedback_min; A.feedback_max = feedback_max; A.sharpen = sharpen;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) store_h4(out + 4 * ((size_t)y * W + x), taa_pixel(A, x, y));
}
__attribute__((visibility("default"))) void emu_blit_rgba16f(int W, int H, const uint16_t* src, int channels, uint16_t* out)
{
    const ImgView s = view(src, W, H, channels);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) store_h4(out + 4 * ((size_t)y * W + x), fetch_texel(s, x, y));
}
__attribute__((visibility("default"))) void emu_tonemap(int W, int H, const uint16_t* src, int channels, float exposure, int single_channel, uint32_t* out)
{
    const ImgView s = view(src, W, H, channels);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) out[(size_t)y * W + x] = tonemap_pixel(fetch_texel(s, x, y), exposure, single_channel);
}
// the two exact conversions the host build substitutes for the hardware instructions: exhaustively testable
__attribute__((visibility("default"))) float    emu_half_to_float(uint16_t h) { return half_bits_to_float(h); }
__attribute__((visibility("default"))) uint16_t emu_float_to_half(float f) { return float_to_half_bits(f); }

} // extern "C"
