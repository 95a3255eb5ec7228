// io_image.hpp -- the per-pixel bridge of the reference's old/include/io_image.hpp (rgb_to_quant_stream,
// quant_stream_to_rgb, :156-192) behind the same names, running on the device.  File I/O (stb), resizing
// and centring are outside the hot path and are not provided here.
#pragma once
#include <cstdint>
#include <vector>

#include "ternary_image_codec_v6_min.hpp"

struct ImageU8 {
    int w = 0, h = 0, c = 0;
    std::vector<uint8_t> data;
};

inline void rgb_to_quant_stream(const ImageU8& rgb, std::vector<PixelYCbCrQuant>& out)
{
    const size_t n = (size_t)rgb.w * (size_t)rgb.h;
    out.assign(n, PixelYCbCrQuant{});
    if (n) t3c_rgb_to_quant(t3c_shim::context(), rgb.data.data(), n, reinterpret_cast<t3c_pixel*>(out.data()));
}

inline void quant_stream_to_rgb(const std::vector<PixelYCbCrQuant>& q, int w, int h, ImageU8& out)
{
    out.w = w; out.h = h; out.c = 3;
    const size_t n = (size_t)w * (size_t)h;
    out.data.assign(n * 3, 0);
    if (n) t3c_quant_to_rgb(t3c_shim::context(), reinterpret_cast<const t3c_pixel*>(q.data()), n, out.data.data());
}
