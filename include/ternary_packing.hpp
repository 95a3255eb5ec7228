// ternary_packing.hpp -- drop-in for include/ternary_packing.hpp:18-65 of the reference (namespace tpack):
// base-243 packing of trit streams (5 trits per byte behind a uint32 trit count) and
// Word27 <-> 9 bytes per word, every symbol reduced mod 27.  words_to_bytes runs on the device
// (t3c_words_to_bytes); bytes_to_words is the same map in the other direction, so it uses the same call.
#pragma once
#include <cstdint>
#include <vector>

#include "ternary_image_codec_v6_min.hpp"

namespace tpack {

// ut_to_base243, :29-39
inline void ut_to_base243(const std::vector<UTrit>& in, std::vector<uint8_t>& out)
{
    out.assign(4 + (in.size() + 4) / 5, 0);
    size_t n = 0;
    t3c_base243_pack(t3c_shim::context(), in.data(), in.size(), out.data(), &n);
}
// base243_to_ut, :41-50: false when the payload is shorter than the count it announces (out then holds what was there)
inline bool base243_to_ut(const std::vector<uint8_t>& in, std::vector<UTrit>& out)
{
    out.clear();
    if (in.size() < 4) return false;
    out.assign(5 * (in.size() - 4), 0);
    size_t n = 0;
    int ok = 0;
    t3c_base243_unpack(t3c_shim::context(), in.data(), in.size(), out.data(), out.size(), &n, &ok);
    out.resize(n);
    return ok != 0;
}

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
