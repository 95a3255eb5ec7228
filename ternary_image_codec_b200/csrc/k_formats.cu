// k_formats.cu -- SURVEY 8(f) "next" rows: the data formats either side of the coded path.  Pure bandwidth kernels.
//   8(f).2  sub-word trit streams (OLD:816-859) and base-243 packing (include/ternary_packing.hpp:18-50)
//   8(f).3  NEW-generation RAW path: one pixel <-> one 32-bit word (src/ternary_image_codec_v6_min.cpp:62-126)
// One thread produces 16 (or 4/12) output bytes so that global stores are 128-bit; inputs are gathered through L1
// (every input byte is read by one or two neighbouring threads).
#include <cuda_runtime.h>

#include <cstdint>

#include "launch.h"

namespace t3c {
namespace {

constexpr int TPB = 256;
inline unsigned blocks_for(size_t n, int per) { return (unsigned)((n + per - 1) / per); }

// trit t (0..26) of a word: digit t%3 of symbol t/3 reduced mod 27 (unpack3, OLD:28-31, on an arbitrary byte)
__device__ __forceinline__ uint32_t word_trit(const uint8_t* __restrict__ w9, uint32_t t)
{
    const uint32_t s = w9[t / 3], c = t % 3;
    return c == 0 ? s % 3 : (c == 1 ? (s / 3) % 3 : (s / 9) % 3);
}

// extract_subword_stream_from_words: out[N*w + i] = trit i of word w.  Thread = 16 consecutive output bytes.
__global__ void __launch_bounds__(TPB) k_subword_stream(const uint8_t* __restrict__ words, uint64_t n_out, uint32_t N, uint8_t* __restrict__ out)
{
    const uint64_t o0 = 16 * ((uint64_t)blockIdx.x * TPB + threadIdx.x);
    if (o0 >= n_out) return;
    uint64_t w = o0 / N;
    uint32_t i = (uint32_t)(o0 - w * N);
    uint32_t v[4] = {0, 0, 0, 0};
    const int cnt = n_out - o0 < 16 ? (int)(n_out - o0) : 16;
    for (int q = 0; q < cnt; ++q) {
        v[q >> 2] |= word_trit(words + 9 * w, i) << (8 * (q & 3));
        if (++i == N) { i = 0; ++w; }
    }
    if (cnt == 16 && !(reinterpret_cast<uintptr_t>(out) & 15)) *reinterpret_cast<uint4*>(out + o0) = make_uint4(v[0], v[1], v[2], v[3]);
    else for (int q = 0; q < cnt; ++q) out[o0 + q] = (uint8_t)(v[q >> 2] >> (8 * (q & 3)));
}
// build_words_from_subword_stream: word w takes trits [N*w, N*w+N) (zeros past the end of the stream), trits N..26 = fill.
// pack3 narrows to a byte (GF27 is uint8_t), so out-of-range "trits" wrap exactly like the reference.
__global__ void __launch_bounds__(TPB) k_words_from_subword_stream(const uint8_t* __restrict__ trits, uint64_t n_trits, uint32_t N, uint32_t fill,
                                                                    uint64_t n_words, uint8_t* __restrict__ words)
{
    const uint64_t w = (uint64_t)blockIdx.x * TPB + threadIdx.x;
    if (w >= n_words) return;
    for (int s = 0; s < 9; ++s) {
        uint32_t sym = 0, mul = 1;
        for (int c = 0; c < 3; ++c, mul *= 3) {
            const uint32_t t = 3 * s + c;
            const uint64_t gi = N * w + t;
            sym += mul * (t < N ? (gi < n_trits ? trits[gi] : 0u) : fill);
        }
        words[9 * w + s] = (uint8_t)sym;
    }
}
// ut_to_base243: out[0..3] = count (LE), out[4 + j] = (uint8)(t[5j] + 3 t[5j+1] + 9 t[5j+2] + 27 t[5j+3] + 81 t[5j+4]), zero padded
__global__ void __launch_bounds__(TPB) k_base243_pack(const uint8_t* __restrict__ trits, uint64_t n_trits, uint8_t* __restrict__ out)
{
    const uint64_t j = (uint64_t)blockIdx.x * TPB + threadIdx.x, nb = (n_trits + 4) / 5;
    if (j == 0) { const uint32_t total = (uint32_t)n_trits; for (int i = 0; i < 4; ++i) out[i] = (uint8_t)(total >> (8 * i)); }
    if (j >= nb) return;
    uint32_t v = 0, mul = 1;
    for (int c = 0; c < 5; ++c, mul *= 3) { const uint64_t ti = 5 * j + c; v += mul * (ti < n_trits ? trits[ti] : 0u); }
    out[4 + j] = (uint8_t)v;
}
// base243_to_ut: trit i = digit i%5 of payload[i/5] (the byte as an int: v%3, then v/=3)
__global__ void __launch_bounds__(TPB) k_base243_unpack(const uint8_t* __restrict__ payload, uint64_t n_trits, uint8_t* __restrict__ trits)
{
    const uint64_t j = (uint64_t)blockIdx.x * TPB + threadIdx.x; // one payload byte -> five trits
    if (5 * j >= n_trits) return;
    uint32_t v = payload[j];
    for (int c = 0; c < 5; ++c) { if (5 * j + c < n_trits) trits[5 * j + c] = (uint8_t)(v % 3); v /= 3; }
}
// fused extract + base-243: byte j packs stream trits 5j..5j+4, stream trit t = trit t%N of word t/N
__global__ void __launch_bounds__(TPB) k_words_to_base243(const uint8_t* __restrict__ words, uint64_t n_trits, uint32_t N, uint8_t* __restrict__ out)
{
    const uint64_t j = (uint64_t)blockIdx.x * TPB + threadIdx.x, nb = (n_trits + 4) / 5;
    if (j == 0) { const uint32_t total = (uint32_t)n_trits; for (int i = 0; i < 4; ++i) out[i] = (uint8_t)(total >> (8 * i)); }
    if (j >= nb) return;
    uint64_t w = (5 * j) / N;
    uint32_t i = (uint32_t)(5 * j - w * N), v = 0, mul = 1;
    for (int c = 0; c < 5; ++c, mul *= 3) {
        if (5 * j + c < n_trits) v += mul * word_trit(words + 9 * w, i);
        if (++i == N) { i = 0; ++w; }
    }
    out[4 + j] = (uint8_t)v;
}

// NEW-generation RAW path.  pack13_from_quant (:62-78): clamp(Yq,0,242) + 243 (clamp(Cbq+40,0,80) + 81 clamp(Crq+40,0,80)).
// Thread = 4 pixels: 24 bytes in (three 64-bit loads), one 128-bit store.
__device__ __forceinline__ uint32_t v6new_code(uint32_t yq, int cb, int cr)
{
    const uint32_t Y = min(yq, 242u), Cb = (uint32_t)min(max(cb + 40, 0), 80), Cr = (uint32_t)min(max(cr + 40, 0), 80);
    return Y + 243u * (Cb + 81u * Cr);
}
__global__ void __launch_bounds__(TPB) k_v6new_pack(const uint16_t* __restrict__ px, uint64_t n_px, uint32_t* __restrict__ words)
{
    const uint64_t p0 = 4 * ((uint64_t)blockIdx.x * TPB + threadIdx.x);
    if (p0 >= n_px) return;
    if (p0 + 4 <= n_px && !((reinterpret_cast<uintptr_t>(px) & 7) | (reinterpret_cast<uintptr_t>(words) & 15))) {
        const uint2* src = reinterpret_cast<const uint2*>(px + 3 * p0); // 24 bytes, 8-byte aligned
        const uint2 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
        const uint32_t h[6] = {a.x, a.y, b.x, b.y, c.x, c.y};
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            auto hw = [&](int k) { return (h[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu; };
            o[i] = v6new_code(hw(3 * i), (int)(int16_t)hw(3 * i + 1), (int)(int16_t)hw(3 * i + 2));
        }
        *reinterpret_cast<uint4*>(words + p0) = make_uint4(o[0], o[1], o[2], o[3]);
    } else
        for (uint64_t p = p0; p < n_px && p < p0 + 4; ++p) words[p] = v6new_code(px[3 * p], (int)(int16_t)px[3 * p + 1], (int)(int16_t)px[3 * p + 2]);
}
// unpack13_to_quant (:81-95): Y = min(code % 243, 242), Cb = clamp((code/243) % 81 - 40), Cr = clamp((code/243)/81 - 40, -40, 40)
__device__ __forceinline__ void v6new_fields(uint32_t code, uint32_t& Y, uint32_t& Cb, uint32_t& Cr)
{
    const uint32_t block = code / 243u, cr = block / 81u;
    Y = min(code - 243u * block, 242u);
    Cb = (block - 81u * cr) - 40u;                    // already in [-40, 40]
    Cr = (uint32_t)(min((int)cr, 80) - 40);            // arbitrary 32-bit words can exceed 80
}
__global__ void __launch_bounds__(TPB) k_v6new_unpack(const uint32_t* __restrict__ words, uint64_t n_words, uint16_t* __restrict__ px)
{
    const uint64_t p0 = 4 * ((uint64_t)blockIdx.x * TPB + threadIdx.x);
    if (p0 >= n_words) return;
    if (p0 + 4 <= n_words && !((reinterpret_cast<uintptr_t>(px) & 7) | (reinterpret_cast<uintptr_t>(words) & 15))) {
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(words + p0));
        const uint32_t c[4] = {w.x, w.y, w.z, w.w};
        uint32_t hws[12];
#pragma unroll
        for (int i = 0; i < 4; ++i) v6new_fields(c[i], hws[3 * i], hws[3 * i + 1], hws[3 * i + 2]);
        uint2* dst = reinterpret_cast<uint2*>(px + 3 * p0);
#pragma unroll
        for (int q = 0; q < 3; ++q)
            dst[q] = make_uint2((hws[4 * q] & 0xFFFFu) | (hws[4 * q + 1] << 16), (hws[4 * q + 2] & 0xFFFFu) | (hws[4 * q + 3] << 16));
    } else
        for (uint64_t p = p0; p < n_words && p < p0 + 4; ++p) {
            uint32_t Y, Cb, Cr;
            v6new_fields(words[p], Y, Cb, Cr);
            px[3 * p] = (uint16_t)Y; px[3 * p + 1] = (uint16_t)Cb; px[3 * p + 2] = (uint16_t)Cr;
        }
}

} // namespace

int launch_subword_stream(const uint8_t* words9, size_t n_words, int N, uint8_t* trits, cudaStream_t st)
{
    const uint64_t n_out = (uint64_t)n_words * (uint64_t)N;
    if (!n_out) return 0;
    k_subword_stream<<<blocks_for((n_out + 15) / 16, TPB), TPB, 0, st>>>(words9, n_out, (uint32_t)N, trits);
    return 1;
}
int launch_words_from_subword_stream(const uint8_t* trits, size_t n_trits, int N, uint8_t fill, uint8_t* words9, cudaStream_t st)
{
    const uint64_t nw = (n_trits + (size_t)N - 1) / (size_t)N;
    if (!nw) return 0;
    k_words_from_subword_stream<<<blocks_for(nw, TPB), TPB, 0, st>>>(trits, n_trits, (uint32_t)N, fill, nw, words9);
    return 1;
}
int launch_base243_pack(const uint8_t* trits, size_t n_trits, uint8_t* out, cudaStream_t st)
{
    k_base243_pack<<<blocks_for((n_trits + 4) / 5 + 1, TPB), TPB, 0, st>>>(trits, n_trits, out);
    return 1;
}
int launch_base243_unpack(const uint8_t* payload, size_t n_trits, uint8_t* trits, cudaStream_t st)
{
    if (!n_trits) return 0;
    k_base243_unpack<<<blocks_for((n_trits + 4) / 5, TPB), TPB, 0, st>>>(payload, n_trits, trits);
    return 1;
}
int launch_words_to_base243(const uint8_t* words9, size_t n_words, int N, uint8_t* out, cudaStream_t st)
{
    const uint64_t n_trits = (uint64_t)n_words * (uint64_t)N;
    k_words_to_base243<<<blocks_for((n_trits + 4) / 5 + 1, TPB), TPB, 0, st>>>(words9, n_trits, (uint32_t)N, out);
    return 1;
}
int launch_v6new_pack_pixels(const t3c_pixel* px, size_t n_px, uint32_t* words, cudaStream_t st)
{
    if (!n_px) return 0;
    k_v6new_pack<<<blocks_for((n_px + 3) / 4, TPB), TPB, 0, st>>>(reinterpret_cast<const uint16_t*>(px), n_px, words);
    return 1;
}
int launch_v6new_unpack_pixels(const uint32_t* words, size_t n_words, t3c_pixel* px, cudaStream_t st)
{
    if (!n_words) return 0;
    k_v6new_unpack<<<blocks_for((n_words + 3) / 4, TPB), TPB, 0, st>>>(words, n_words, reinterpret_cast<uint16_t*>(px));
    return 1;
}

} // namespace t3c
