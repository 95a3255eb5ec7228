// t3v_io.hpp -- drop-in for old/include/t3v_io.hpp of the reference: the .t3v container (54-byte header + frame records
// n | 9n symbol bytes | CRC).  Same names and signatures; the record bytes and both checksums are produced on the device
// (t3c_t3v_frame_record / t3c_t3v_read_frame / t3c_t3v_header, include/t3c.h), the host only moves them through FILE*.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "ternary_image_codec_v6_min.hpp"

namespace t3v_detail {
inline uint32_t crc32(const void* data, size_t len)
{
    uint32_t c = 0;
    t3c_crc32(t3c_shim::context(), static_cast<const uint8_t*>(data), len, &c);
    return c;
}
} // namespace t3v_detail

#pragma pack(push, 1)
struct T3VHeaderBin {
    char magic[4];
    uint8_t version, file_type, profile, subword_code, centered, coset;
    uint32_t width, height;
    uint32_t aw_x0, aw_y0, aw_w, aw_h;
    uint32_t fps_num, fps_den;
    uint32_t frame_count;
    uint32_t reserved0;
    uint32_t header_crc32;
};
#pragma pack(pop)
static_assert(sizeof(T3VHeaderBin) == 54, "packed .t3v header");

inline uint8_t subword_to_code(SubwordMode m)
{
    switch (m) {
    case SubwordMode::S27: return 0;
    case SubwordMode::S24: return 1;
    case SubwordMode::S21: return 2;
    case SubwordMode::S18: return 3;
    case SubwordMode::S15: return 4;
    }
    return 0;
}
inline SubwordMode code_to_subword(uint8_t c)
{
    switch (c) {
    case 1: return SubwordMode::S24;
    case 2: return SubwordMode::S21;
    case 3: return SubwordMode::S18;
    case 4: return SubwordMode::S15;
    default: return SubwordMode::S27;
    }
}

inline bool t3v_write_header(FILE* f, ProfileID prof, SubwordMode sub, bool centered, CosetID coset, uint32_t width, uint32_t height, const ActiveWindow& aw,
                             uint32_t fps_num = 0, uint32_t fps_den = 1, uint32_t frame_count = 1, uint8_t file_type = 0)
{
    uint8_t h[54];
    const uint32_t a[4] = {aw.x0, aw.y0, aw.w, aw.h};
    if (t3c_t3v_header(t3c_shim::context(), h, (int)prof, subword_to_code(sub), centered ? 1 : 0, (int)coset, width, height, a, fps_num, fps_den, frame_count,
                       file_type) != T3C_OK)
        return false;
    return std::fwrite(h, sizeof h, 1, f) == 1;
}
inline bool t3v_read_header(FILE* f, T3VHeaderBin& h)
{
    if (std::fread(&h, sizeof(h), 1, f) != 1) return false;
    if (std::memcmp(h.magic, "T3V1", 4) != 0) return false;
    return t3v_detail::crc32(&h, sizeof(T3VHeaderBin) - sizeof(uint32_t)) == h.header_crc32;
}
inline bool t3v_write_frame(FILE* f, const std::vector<Word27>& words)
{
    std::vector<uint8_t> rec(8 + 9 * words.size());
    size_t n = 0;
    if (t3c_t3v_frame_record(t3c_shim::context(), reinterpret_cast<const uint8_t*>(words.data()), words.size(), rec.data(), &n) != T3C_OK) return false;
    return std::fwrite(rec.data(), n, 1, f) == 1;
}
inline bool t3v_read_frame(FILE* f, std::vector<Word27>& words)
{
    uint32_t n = 0;
    if (std::fread(&n, sizeof(n), 1, f) != 1) return false;
    std::vector<uint8_t> rec(8 + 9 * (size_t)n);
    std::memcpy(rec.data(), &n, 4);
    if (std::fread(rec.data() + 4, rec.size() - 4, 1, f) != 1) return false;
    std::vector<Word27> out(n);
    size_t got = 0;
    int ok = 0;
    if (t3c_t3v_read_frame(t3c_shim::context(), rec.data(), rec.size(), reinterpret_cast<uint8_t*>(out.data()), out.size(), &got, &ok) != T3C_OK || !ok) return false;
    words.swap(out);
    return true;
}
inline SubwordMode t3v_header_subword(const T3VHeaderBin& h) { return code_to_subword(h.subword_code); }
inline ActiveWindow t3v_header_aw(const T3VHeaderBin& h) { return {h.aw_x0, h.aw_y0, h.aw_w, h.aw_h}; }
inline FILE* t3v_fopen(const std::string& path, const char* mode) { return std::fopen(path.c_str(), mode); }
inline void t3v_fclose(FILE* f) { if (f) std::fclose(f); }
