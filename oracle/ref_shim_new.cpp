// ref_shim_new.cpp -- extern "C" wrapper around the reference's NEW-generation core (TEST INFRASTRUCTURE):
// include/ternary_image_codec_v6_min.hpp + src/ternary_image_codec_v6_min.cpp, compiled where they lie under
// /root/reference by oracle/Makefile into oracle/_ref/libt3ref_new.so (a separate library: the NEW Word27 is a
// uint32_t and clashes with the OLD generation's names).  SURVEY.md 8(f).3.
#include <cstdint>
#include <cstring>
#include <vector>

#include "ternary_image_codec_v6_min.hpp" // -I<reference>/include

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
    if (n_words) std::memcpy(in.data(), words, 4 * n_words);
    std::vector<PixelYCbCrQuant> out;
    const bool ok = subword ? decode_raw_words_to_pixels_subword(in, (SubwordMode)subword, out) : decode_raw_words_to_pixels(in, out);
    if (ok && !out.empty()) std::memcpy(px6, out.data(), 6 * out.size());
    return ok ? 1 : 0;
}
}
