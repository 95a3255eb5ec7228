// io_image.hpp -- drop-in for the reference's old/include/io_image.hpp ("IMG"): the per-pixel bridge (rgb_to_quant_stream,
// quant_stream_to_rgb, IMG:156-192), the image geometry (resize_rgb_nn, blit_center_rgb, IMG:86-127; extract_center_q of the later
// include/io_image.hpp:215-235) and the file -> words conveniences (image_to_words27, words27_to_image, IMG:193-251) behind the same
// names.  Pixels, geometry and words are computed on the device (t3c_* calls of include/t3c.h); only the file decoding itself stays
// with stb_image, exactly as in the reference: the four stbi_* prototypes are declared here and one translation unit defines
// TERNARY_IO_IMAGE_IMPLEMENTATION with stb_image.h / stb_image_write.h on its include path (the reference's third-party/).
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "ternary_image_codec_v6_min.hpp"

struct ImageU8 {
    int w = 0, h = 0, c = 0;
    std::vector<uint8_t> data;
};

extern "C" {
unsigned char* stbi_load(const char* filename, int* x, int* y, int* comp, int req_comp);
void stbi_image_free(void* retval_from_stbi_load);
int stbi_write_png(const char* filename, int w, int h, int comp, const void* data, int stride_in_bytes);
int stbi_write_jpg(const char* filename, int w, int h, int comp, const void* data, int quality);
}
#ifdef TERNARY_IO_IMAGE_IMPLEMENTATION
#define STB_IMAGE_IMPLEMENTATION
#define STB_IMAGE_WRITE_IMPLEMENTATION
extern "C" {
#include "stb_image.h"
#include "stb_image_write.h"
}
#endif

inline void rgb_to_quant_stream(const ImageU8& rgb, std::vector<PixelYCbCrQuant>& out)
{
    const size_t n = (size_t)rgb.w * (size_t)rgb.h;
    t3c_shim::size_for_output(out, n);
    if (n) t3c_rgb_to_quant(t3c_shim::context(), rgb.data.data(), n, reinterpret_cast<t3c_pixel*>(out.data()));
}

inline void quant_stream_to_rgb(const std::vector<PixelYCbCrQuant>& q, int w, int h, ImageU8& out)
{
    out.w = w; out.h = h; out.c = 3;
    const size_t n = (size_t)w * (size_t)h;
    t3c_shim::size_for_output(out.data, n * 3);
    if (n) t3c_quant_to_rgb(t3c_shim::context(), reinterpret_cast<const t3c_pixel*>(q.data()), n, out.data.data());
}

// ---- image geometry (IMG:86-131)
inline void resize_rgb_nn(const ImageU8& src, int dstW, int dstH, ImageU8& dst)
{
    dst.w = dstW; dst.h = dstH; dst.c = 3;
    dst.data.assign((size_t)dstW * dstH * 3, 0);
    if (src.w <= 0 || src.h <= 0 || dstW <= 0 || dstH <= 0) return;
    t3c_resize_rgb_nn(t3c_shim::context(), src.data.data(), src.w, src.h, dst.data.data(), dstW, dstH);
}
inline void blit_center_rgb(const ImageU8& src, int canvasW, int canvasH, ImageU8& dst)
{
    dst.w = canvasW; dst.h = canvasH; dst.c = 3;
    dst.data.assign((size_t)canvasW * canvasH * 3, 0);
    if (src.w <= 0 || src.h <= 0 || canvasW <= 0 || canvasH <= 0) return;
    t3c_blit_center_rgb(t3c_shim::context(), src.data.data(), src.w, src.h, dst.data.data(), canvasW, canvasH);
}
inline int pad_even(int w) { return (w % 2 == 0) ? w : (w + 1); }
// centre window of a quantised frame (include/io_image.hpp:215-235 of the later generation)
inline void extract_center_q(const std::vector<PixelYCbCrQuant>& q_full, int fullW, int fullH, int subW, int subH, std::vector<PixelYCbCrQuant>& q_sub)
{
    q_sub.assign((size_t)std::max(subW, 0) * (size_t)std::max(subH, 0), PixelYCbCrQuant{});
    if (q_sub.empty() || fullW <= 0 || fullH <= 0) return;
    t3c_extract_center_q(t3c_shim::context(), reinterpret_cast<const t3c_pixel*>(q_full.data()), fullW, fullH, subW, subH, reinterpret_cast<t3c_pixel*>(q_sub.data()));
}

// ---- disk I/O: stb, as in the reference (IMG:133-154)
inline bool load_image_rgb8(const std::string& path, ImageU8& out)
{
    int x = 0, y = 0, n = 0;
    unsigned char* pix = stbi_load(path.c_str(), &x, &y, &n, 3);
    if (!pix) return false;
    out.w = x; out.h = y; out.c = 3;
    out.data.assign(pix, pix + (size_t)x * y * 3);
    stbi_image_free(pix);
    return true;
}
inline bool save_image_png(const std::string& path, const ImageU8& img) { return stbi_write_png(path.c_str(), img.w, img.h, 3, img.data.data(), img.w * 3) != 0; }
inline bool save_image_jpg(const std::string& path, const ImageU8& img, int quality = 90) { return stbi_write_jpg(path.c_str(), img.w, img.h, 3, img.data.data(), quality) != 0; }

// ---- file -> words and back (IMG:193-251): resize to the sub-word format's resolution, optional centring in the 8K canvas, odd widths
// padded by repeating the last column, then the bridge and the two-pixel packing -- every step a device call
inline bool image_to_words27(const std::string& path, std::vector<Word27>& out_words, SubwordMode sub = SubwordMode::S27, bool centered = true)
{
    ImageU8 src;
    if (!load_image_rgb8(path, src)) return false;
    const StdRes tgt = std_res_for(sub);
    ImageU8 work;
    if (src.w != tgt.w || src.h != tgt.h) resize_rgb_nn(src, tgt.w, tgt.h, work); else work = src;
    ImageU8 canvas = work;
    if (sub != SubwordMode::S27 && centered) {
        ImageU8 tmp;
        blit_center_rgb(work, std_res_for(SubwordMode::S27).w, std_res_for(SubwordMode::S27).h, tmp);
        std::swap(canvas, tmp);
    }
    const int evenW = pad_even(canvas.w);
    if (evenW != canvas.w) {   // never taken for the standard resolutions (all even): plain row copies
        ImageU8 pad = canvas;
        pad.w = evenW;
        pad.data.resize((size_t)evenW * pad.h * 3);
        for (int y = 0; y < canvas.h; ++y) {
            const uint8_t* srcp = &canvas.data[(size_t)y * canvas.w * 3];
            uint8_t* dstp = &pad.data[(size_t)y * evenW * 3];
            std::copy(srcp, srcp + (size_t)canvas.w * 3, dstp);
            std::copy(srcp + (canvas.w - 1) * 3, srcp + canvas.w * 3, dstp + (evenW - 1) * 3);
        }
        std::swap(canvas, pad);
    }
    std::vector<PixelYCbCrQuant> q;
    rgb_to_quant_stream(canvas, q);
    return encode_raw_pixels_to_words(q, out_words);
}
inline bool words27_to_image(const std::vector<Word27>& words, int w, int h, const std::string& out_path_png)
{
    std::vector<PixelYCbCrQuant> q;
    if (!decode_raw_words_to_pixels(words, q)) return false;
    if ((int)q.size() < w * h) return false;
    ImageU8 img;
    quant_stream_to_rgb(q, w, h, img);
    return save_image_png(out_path_png, img);
}
