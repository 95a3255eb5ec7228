// ternary_packing.hpp -- drop-in for include/ternary_packing.hpp:53-65 of the reference (namespace tpack):
// Word27 <-> 9 bytes per word, every symbol reduced mod 27.  words_to_bytes runs on the device
// (t3c_words_to_bytes); bytes_to_words is the same map in the other direction, so it uses the same call.
#pragma once
#include <cstdint>
#include <vector>

#include "ternary_image_codec_v6_min.hpp"

namespace tpack {

inline void words_to_bytes(const std::vector<Word27>& words, std::vector<uint8_t>& out)
{
    out.assign(words.size() * 9, 0);
    if (!words.empty()) t3c_words_to_bytes(t3c_shim::context(), reinterpret_cast<const uint8_t*>(words.data()), words.size(), out.data());
}

inline void bytes_to_words(const std::vector<uint8_t>& in, std::vector<Word27>& out)
{
    out.clear();
    if (in.size() % 9 != 0) return; // the reference silently returns an empty vector
    out.resize(in.size() / 9);
    if (!out.empty()) t3c_words_to_bytes(t3c_shim::context(), in.data(), out.size(), reinterpret_cast<uint8_t*>(out.data()));
}

} // namespace tpack
