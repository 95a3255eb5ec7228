// ref_shim_new.cpp -- extern "C" wrapper around the reference's NEW-generation core (TEST INFRASTRUCTURE):
// include/ternary_image_codec_v6_min.hpp + src/ternary_image_codec_v6_min.cpp, compiled where they lie under
// /root/reference by oracle/Makefile into oracle/_ref/libt3ref_new.so (a separate library: the NEW Word27 is a
// uint32_t and clashes with the OLD generation's names).  SURVEY.md 8(f).3.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ternary_image_codec_v6_min.hpp" // -I<reference>/include
#include "io_image.hpp"                   // -I<reference>/include: image bridge; its stb entry points are the stubs below

static_assert(sizeof(Word27) == 4 && sizeof(PixelYCbCrQuant) == 6, "NEW-generation layouts");

extern "C" {
int t3n_pack_pixels(const void* px6, size_t n_px, uint32_t* words, int subword)
{
    std::vector<PixelYCbCrQuant> in(n_px);
    if (n_px) std::memcpy(in.data(), px6, 6 * n_px);
    std::vector<Word27> out;
    const bool ok = subword ? encode_raw_pixels_to_words_subword(in, (SubwordMode)subword, out) : encode_raw_pixels_to_words(in, out);
    if (ok && !out.empty()) std::memcpy(words, out.data(), 4 * out.size());
    return ok ? 1 : 0;
}
int t3n_unpack_pixels(const uint32_t* words, size_t n_words, void* px6, int subword)
{
    std::vector<Word27> in(n_words);
    if (n_words) std::memcpy(static_cast<void*>(in.data()), words, 4 * n_words);
    std::vector<PixelYCbCrQuant> out;
    const bool ok = subword ? decode_raw_words_to_pixels_subword(in, (SubwordMode)subword, out) : decode_raw_words_to_pixels(in, out);
    if (ok && !out.empty()) std::memcpy(px6, out.data(), 6 * out.size());
    return ok ? 1 : 0;
}
}

// ---- the image bridge.  image_to_words_subword / words_to_image_subword go through stb for file I/O; the four stb
// entry points the header declares are defined here as in-memory stubs, so the reference's own pipeline code runs unchanged.
static const uint8_t* g_img = nullptr;
static int g_w = 0, g_h = 0;
static std::vector<uint8_t> g_saved;
static int g_sw = 0, g_sh = 0;
extern "C" {
unsigned char* stbi_load(const char*, int* x, int* y, int* comp, int)
{
    if (!g_img || g_w <= 0 || g_h <= 0) return nullptr;
    unsigned char* p = (unsigned char*)std::malloc((size_t)g_w * g_h * 3);
    std::memcpy(p, g_img, (size_t)g_w * g_h * 3);
    *x = g_w; *y = g_h; *comp = 3;
    return p;
}
void stbi_image_free(void* p) { std::free(p); }
int stbi_write_png(const char*, int w, int h, int comp, const void* data, int)
{
    g_sw = w; g_sh = h;
    g_saved.assign((const uint8_t*)data, (const uint8_t*)data + (size_t)w * h * comp);
    return 1;
}
int stbi_write_jpg(const char*, int, int, int, const void*, int) { return 0; }

void t3n_resize_rgb_nn(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh)
{
    ImageU8 a, b;
    a.w = sw; a.h = sh; a.c = 3; a.data.assign(src, src + (size_t)(sw > 0 && sh > 0 ? sw : 0) * (sh > 0 ? sh : 0) * 3);
    resize_rgb_nn(a, dw, dh, b);
    if (!b.data.empty()) std::memcpy(dst, b.data.data(), b.data.size());
}
void t3n_blit_center_rgb(const uint8_t* src, int sw, int sh, uint8_t* dst, int cw, int ch)
{
    ImageU8 a, b;
    a.w = sw; a.h = sh; a.c = 3; a.data.assign(src, src + (size_t)sw * sh * 3);
    blit_center_rgb(a, cw, ch, b);
    if (!b.data.empty()) std::memcpy(dst, b.data.data(), b.data.size());
}
size_t t3n_extract_center_q(const void* full6, int fw, int fh, int sw, int sh, void* sub6)
{
    std::vector<PixelYCbCrQuant> f((size_t)fw * fh), s;
    if (!f.empty()) std::memcpy(f.data(), full6, 6 * f.size());
    extract_center_q(f, fw, fh, sw, sh, s);
    if (!s.empty()) std::memcpy(sub6, s.data(), 6 * s.size());
    return s.size();
}
long long t3n_image_to_words_subword(const uint8_t* rgb, int w, int h, int sub, int centered, uint32_t* words, size_t cap)
{
    g_img = rgb; g_w = w; g_h = h;
    std::vector<Word27> out;
    const bool ok = image_to_words_subword("in-memory", (SubwordMode)sub, centered != 0, out);
    g_img = nullptr;
    if (!ok) return -1;
    if (out.size() > cap) return -2;
    if (!out.empty()) std::memcpy(words, out.data(), 4 * out.size());
    return (long long)out.size();
}
int t3n_words_to_image_subword(const uint32_t* words, size_t n, int sub, int w, int h, uint8_t* rgb)
{
    std::vector<Word27> in(n);
    if (n) std::memcpy(static_cast<void*>(in.data()), words, 4 * n);
    g_saved.clear(); g_sw = g_sh = 0;
    const bool ok = words_to_image_subword(in, (SubwordMode)sub, w, h, "in-memory.png");
    if (ok && g_sw == w && g_sh == h && !g_saved.empty()) std::memcpy(rgb, g_saved.data(), g_saved.size());
    return ok ? 1 : 0;
}
}
