// e2e_vector_api.cpp -- the call chain of old/src/main.cpp:15-26 (image -> quant -> raw words -> profile words, and back) through the
// drop-in std::vector API on one synthetic 8K frame, timed end to end on the host clock: every call moves its std::vector (pageable
// host memory) across PCIe.  Prints one JSON line.  Build: g++ -std=c++17 -O2 -Iinclude tools/e2e_vector_api.cpp -L<pkg> -lt3c
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "io_image.hpp"

int main(int argc, char** argv)
{
    const int iters = argc > 1 ? std::atoi(argv[1]) : 6;
    const int W = 7680, H = 4320;
    ImageU8 img;
    img.w = W; img.h = H; img.c = 3;
    img.data.resize((size_t)W * H * 3);
    uint64_t x = 88172645463325252ull;
    for (auto& b : img.data) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; b = (uint8_t)(x >> 32); }
    EncoderContext e;
    e.arith = T3C_FIXED;
    e.cfg.profile = ProfileID::P3_RS26_20;
    uep_uniform(e.cfg.uep, 2);
    DecoderContext d;
    d.arith = T3C_FIXED;
    d.fixed_cfg = &e.cfg;
    std::vector<PixelYCbCrQuant> q, q2;
    std::vector<Word27> raw, prof, raw2;
    ImageU8 back;
    double best = 1e30, first = 0, sum = 0;
    bool ok = true;
    for (int it = 0; it < iters; ++it) {
        const auto t0 = std::chrono::steady_clock::now();
        rgb_to_quant_stream(img, q);
        ok = encode_raw_pixels_to_words(q, raw) && ok;
        ok = encode_profile_from_raw(raw, prof, e) && ok;
        d.expected_raw_words = raw.size();
        ok = decode_profile_to_raw(prof, raw2, d) && ok;
        ok = decode_raw_words_to_pixels(raw2, q2) && ok;
        quant_stream_to_rgb(q2, W, H, back);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (it == 0) first = ms;
        if (it >= 2) { sum += ms; if (ms < best) best = ms; }   // iterations 0 and 1: buffers are noted, then page-locked
    }
    // the decoder recovers a prefix of the raw words (the encoder drops < k symbols per band, bug B8): compare what came back
    ImageU8 want;
    std::vector<PixelYCbCrQuant> qq(q.begin(), q.begin() + (std::ptrdiff_t)q2.size());
    quant_stream_to_rgb(qq, W, (int)(q2.size() / W), want);
    size_t same = 0;
    const size_t n_cmp = std::min(want.data.size(), back.data.size());
    for (size_t i = 0; i < n_cmp; ++i) same += want.data[i] == back.data[i];
    const double mean = sum / (iters - 2);
    std::printf("{\"workload\": \"old/src/main.cpp chain through the std::vector drop-in API, 8K RGB8, RS(26,20): rgb_to_quant_stream + encode_raw_pixels_to_words + "
                "encode_profile_from_raw + decode_profile_to_raw (FIXED) + decode_raw_words_to_pixels + quant_stream_to_rgb\", \"ok\": %s, \"pixels_back\": %zu, "
                "\"bytes_equal\": %zu, \"bytes_compared\": %zu, \"first_call_ms\": %.2f, \"ms_per_frame\": %.2f, \"best_ms\": %.2f, \"mpix_per_s\": %.1f, "
                "\"pcie_bytes_per_frame\": %.0f}\n",
                ok ? "true" : "false", q2.size(), same, n_cmp, first, mean, best, (double)W * H / mean / 1e3,
                3.0 * W * H + 6.0 * W * H * 2 + 9.0 * raw.size() * 2 + 9.0 * prof.size() * 2 + 9.0 * raw2.size() * 2 + 6.0 * q2.size() * 2 + 3.0 * W * H);
    return ok && same == n_cmp ? 0 : 1;
}
