// k_formats.cu -- SURVEY 8(f) "next" rows: the data formats either side of the coded path.  Pure bandwidth kernels.
//   8(f).2  sub-word trit streams (OLD:816-859) and base-243 packing (include/ternary_packing.hpp:18-50)
//   8(f).3  NEW-generation RAW path: one pixel <-> one 32-bit word (src/ternary_image_codec_v6_min.cpp:62-126)
// One thread produces 16 (or 4/12) output bytes so that global stores are 128-bit; inputs are gathered through L1
// (every input byte is read by one or two neighbouring threads).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

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
__global__ void __launch_bounds__(TPB) k_words_to_base243(const uint8_t* __restrict__ words, uint64_t n_trits, uint32_t N, uint8_t* __restrict__ out, uint64_t j0 = 0)
{
    const uint64_t j = j0 + (uint64_t)blockIdx.x * TPB + threadIdx.x, nb = (n_trits + 4) / 5;
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

// ---- tiled versions of the two kernels that feed the .t3p / .t3b writers: a CTA takes 320 consecutive words (2880 bytes in, 320 N trits out;
// 320 N is a multiple of 5, so base-243 bytes never straddle CTAs): coalesced 32-bit loads into shared memory, thread = word: nine symbols ->
// 27 trits through a 27-entry table (three trit bytes per symbol), the first N laid down in shared memory, then coalesced stores.
constexpr int SW_WORDS = 320, SW_TPB = 320;
__device__ __forceinline__ void subword_tile_trits(const uint8_t* __restrict__ words, uint64_t n_words, uint32_t N, uint64_t w0, uint8_t* s_in, uint8_t* s_tr,
                                                   const uint32_t* s_lut)
{
    const uint32_t tid = threadIdx.x;
    const uint64_t left = n_words - w0;
    const uint32_t nw = (uint32_t)(left < SW_WORDS ? left : SW_WORDS), nb = 9u * nw;
    const uint8_t* src = words + 9ull * w0;
    if ((reinterpret_cast<uintptr_t>(src) & 3) == 0) {
        for (uint32_t i = tid; i < nb / 4; i += SW_TPB) reinterpret_cast<uint32_t*>(s_in)[i] = __ldg(reinterpret_cast<const uint32_t*>(src) + i);
        for (uint32_t i = (nb & ~3u) + tid; i < nb; i += SW_TPB) s_in[i] = src[i];
    } else
        for (uint32_t i = tid; i < nb; i += SW_TPB) s_in[i] = src[i];
    __syncthreads();
    if (tid < nw) {
        uint8_t* o = s_tr + N * tid;
        uint32_t e[9];
#pragma unroll
        for (int sy = 0; sy < 9; ++sy) e[sy] = s_lut[s_in[9 * tid + sy] % 27u]; // trit bytes of the symbol reduced mod 27 (unpack3, OLD:28-31)
        // 27 trit bytes as seven words
        const uint32_t w[7] = {e[0] | e[1] << 24, e[1] >> 8 | e[2] << 16, e[2] >> 16 | e[3] << 8, e[4] | e[5] << 24, e[5] >> 8 | e[6] << 16, e[6] >> 16 | e[7] << 8, e[8]};
        if ((N & 3u) == 0) {
#pragma unroll
            for (int j = 0; j < 6; ++j) if (4u * j < N) reinterpret_cast<uint32_t*>(o)[j] = w[j];
        } else {
#pragma unroll
            for (int t = 0; t < 27; ++t) if ((uint32_t)t < N) o[t] = (uint8_t)(w[t >> 2] >> (8 * (t & 3)));
        }
    } else if (tid < SW_WORDS) {
        for (uint32_t t = 0; t < N; ++t) s_tr[N * tid + t] = 0;       // past the end of the stream: zero trits (the base-243 padding)
    }
    __syncthreads();
}
__global__ void __launch_bounds__(SW_TPB) k_subword_stream_tiled(const uint8_t* __restrict__ words, uint64_t n_words, uint32_t N, uint8_t* __restrict__ out)
{
    __shared__ __align__(16) uint8_t s_in[9 * SW_WORDS];
    __shared__ __align__(16) uint8_t s_tr[27 * SW_WORDS];
    __shared__ uint32_t s_lut[27];
    const uint32_t tid = threadIdx.x;
    if (tid < 27) s_lut[tid] = (tid % 3u) | ((tid / 3u) % 3u) << 8 | (tid / 9u) << 16;
    const uint64_t w0 = (uint64_t)blockIdx.x * SW_WORDS;
    subword_tile_trits(words, n_words, N, w0, s_in, s_tr, s_lut);
    const uint64_t left = n_words - w0;
    const uint32_t nw = (uint32_t)(left < SW_WORDS ? left : SW_WORDS), nb = N * nw;
    uint8_t* dst = out + (uint64_t)N * w0;
    if ((reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
        for (uint32_t i = tid; i < nb / 4; i += SW_TPB) reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(s_tr)[i];
        for (uint32_t i = (nb & ~3u) + tid; i < nb; i += SW_TPB) dst[i] = s_tr[i];
    } else
        for (uint32_t i = tid; i < nb; i += SW_TPB) dst[i] = s_tr[i];
}
__global__ void __launch_bounds__(SW_TPB) k_words_to_base243_tiled(const uint8_t* __restrict__ words, uint64_t n_words, uint32_t N, uint8_t* __restrict__ out)
{
    __shared__ __align__(16) uint8_t s_in[9 * SW_WORDS];
    __shared__ __align__(16) uint8_t s_tr[27 * SW_WORDS];
    __shared__ __align__(16) uint8_t s_out[27 * SW_WORDS / 5 + 4];
    __shared__ uint32_t s_lut[27];
    const uint32_t tid = threadIdx.x;
    if (tid < 27) s_lut[tid] = (tid % 3u) | ((tid / 3u) % 3u) << 8 | (tid / 9u) << 16;
    const uint64_t w0 = (uint64_t)blockIdx.x * SW_WORDS, n_trits = n_words * N;
    if (blockIdx.x == 0 && tid < 4) out[tid] = (uint8_t)((uint32_t)n_trits >> (8 * tid));
    subword_tile_trits(words, n_words, N, w0, s_in, s_tr, s_lut);
    const uint64_t byte0 = (uint64_t)N * w0 / 5, nbytes_all = (n_trits + 4) / 5;      // N * w0 is a multiple of 5
    const uint64_t left = nbytes_all - byte0;
    const uint32_t nb = (uint32_t)(left < (uint64_t)N * SW_WORDS / 5 ? left : (uint64_t)N * SW_WORDS / 5);
    for (uint32_t j = tid; j < nb; j += SW_TPB) {
        const uint8_t* t = s_tr + 5 * j;
        s_out[j] = (uint8_t)(t[0] + 3u * t[1] + 9u * t[2] + 27u * t[3] + 81u * t[4]);
    }
    __syncthreads();
    uint8_t* dst = out + 4 + byte0;
    if ((reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
        for (uint32_t i = tid; i < nb / 4; i += SW_TPB) reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(s_out)[i];
        for (uint32_t i = (nb & ~3u) + tid; i < nb; i += SW_TPB) dst[i] = s_out[i];
    } else
        for (uint32_t i = tid; i < nb; i += SW_TPB) dst[i] = s_out[i];
}

// ---- v2 of the fused kernel.  The tiled version above goes through shared memory one byte at a time (~100 wavefronts per word: it is
// bound by the shared-memory pipe at 0.15 of the HBM roofline).  Here a thread owns 20 consecutive words -- 180 bytes in, 20 N trits =
// exactly 4 N payload bytes out -- held in 45 registers; every payload byte is a compile-time sum of at most three pieces
// (symbol / 3^a) % 3^len * 3^c of the symbols it straddles (the pieces come from one 27-entry table look-up per symbol), so N is a
// template parameter.  Global traffic is coalesced through shared
// memory with 32-bit accesses at odd word pitches (45 in, N or N + 1 out): ~4 wavefronts per word.
constexpr int B243_TPB = 128, B243_WPT = 20, B243_INW = 9 * B243_WPT / 4;
__host__ __device__ constexpr uint32_t pow3(int e) { return e <= 0 ? 1u : 3u * pow3(e - 1); }
// piece (symbol / 3^d) % 3^len of symbol byte bi of the thread's 180 bytes: the symbol itself, or a byte of its look-up word
// lut[s] = s % 9 | (s / 9) << 8 | (s % 3) << 16 | (s / 3) << 24 (one conflict-free shared load per symbol, shared by the bytes it straddles)
template <int BI, int D, int LEN>
__device__ __forceinline__ uint32_t b243_piece(const uint32_t (&x)[B243_INW], const uint32_t* lut)
{
    const uint32_t sym = (x[BI >> 2] >> (8 * (BI & 3))) & 0xFFu;
    if constexpr (D == 0 && LEN == 3) return sym;
    else {
        const uint32_t l = lut[sym];
        if constexpr (D == 0 && LEN == 2) return l & 0xFFu;
        else if constexpr (D == 0 && LEN == 1) return (l >> 16) & 0xFFu;
        else if constexpr (D == 1 && LEN == 2) return l >> 24;
        else if constexpr (D == 2) return (l >> 8) & 0xFFu;
        else return (l >> 24) - 3u * ((l >> 8) & 0xFFu);                       // D == 1, LEN == 1: only where a word's N trits end inside a symbol
    }
}
template <int N, int J, int C>
__device__ __forceinline__ uint32_t b243_terms(const uint32_t (&x)[B243_INW], const uint32_t* lut)
{
    if constexpr (C >= 5) return 0u;
    else {
        constexpr int T = 5 * J + C, wd = T / N, tr = T % N, sy = tr / 3, d = tr % 3;
        constexpr int l0 = 3 - d, l1 = 5 - C, l2 = N - tr;                     // to the end of the symbol / of the byte / of the word's N trits
        constexpr int len = l0 < l1 ? (l0 < l2 ? l0 : l2) : (l1 < l2 ? l1 : l2);
        return b243_piece<9 * wd + sy, d, len>(x, lut) * pow3(C) + b243_terms<N, J, C + len>(x, lut);
    }
}
template <int N, int JW>
__device__ __forceinline__ void b243_words(const uint32_t (&x)[B243_INW], const uint32_t* lut, uint32_t (&o)[N])
{
    if constexpr (JW < N) {
        o[JW] = b243_terms<N, 4 * JW, 0>(x, lut) | (b243_terms<N, 4 * JW + 1, 0>(x, lut) << 8) | (b243_terms<N, 4 * JW + 2, 0>(x, lut) << 16) |
                (b243_terms<N, 4 * JW + 3, 0>(x, lut) << 24);
        b243_words<N, JW + 1>(x, lut, o);
    }
}
// bytes >= 27 of a word reduced mod 27 (unpack3, OLD:28-31); bit 7 of (b & 0x7F) + 101 or of b itself is set exactly for b >= 27
__device__ __forceinline__ uint32_t mod27x4(uint32_t w)
{
    if (((((w & 0x7F7F7F7Fu) + 0x65656565u) | w) & 0x80808080u) == 0) return w;
    uint32_t r = 0;
    for (int q = 0; q < 4; ++q) r |= (((w >> (8 * q)) & 0xFFu) % 27u) << (8 * q);
    return r;
}
template <int N>
__global__ void __launch_bounds__(B243_TPB) k_words_to_base243_v2(const uint8_t* __restrict__ words, uint32_t n_groups, uint32_t total_trits, uint8_t* __restrict__ out)
{
    constexpr int PITCH = (N & 1) ? N : N + 1;
    __shared__ __align__(16) uint32_t s_in[B243_TPB * B243_INW];
    uint32_t* s_out = s_in;                                      // the payload words take the input's place once every thread holds its 45 words
    static_assert(PITCH <= B243_INW, "output tile fits the input tile");
    __shared__ uint32_t s_lut[32];
    const uint32_t tid = threadIdx.x, g0 = blockIdx.x * B243_TPB, ng = min((uint32_t)B243_TPB, n_groups - g0);
    if (tid < 32) s_lut[tid] = (tid % 9u) | (tid / 9u) << 8 | (tid % 3u) << 16 | (tid / 3u) << 24;
    if (blockIdx.x == 0 && tid == 0) *reinterpret_cast<uint32_t*>(out) = total_trits;
    const uint8_t* src = words + (size_t)(9 * B243_WPT) * g0;
    const uint32_t nw = ng * B243_INW;
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
        for (uint32_t i = tid; i < nw / 4; i += B243_TPB) {
            uint4 v = __ldcs(reinterpret_cast<const uint4*>(src) + i);
            v.x = mod27x4(v.x); v.y = mod27x4(v.y); v.z = mod27x4(v.z); v.w = mod27x4(v.w);
            reinterpret_cast<uint4*>(s_in)[i] = v;
        }
        for (uint32_t i = (nw & ~3u) + tid; i < nw; i += B243_TPB) s_in[i] = mod27x4(__ldcs(reinterpret_cast<const uint32_t*>(src) + i));
    } else
        for (uint32_t i = tid; i < nw; i += B243_TPB) s_in[i] = mod27x4(__ldcs(reinterpret_cast<const uint32_t*>(src) + i));
    __syncthreads();
    uint32_t x[B243_INW];
#pragma unroll
    for (int i = 0; i < B243_INW; ++i) x[i] = tid < ng ? s_in[B243_INW * tid + i] : 0u;
    __syncthreads();
    if (tid < ng) {
        uint32_t o[N];
        b243_words<N, 0>(x, s_lut, o);
#pragma unroll
        for (int j = 0; j < N; ++j) s_out[PITCH * tid + j] = o[j];
    }
    __syncthreads();
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + 4 + (size_t)(4 * N) * g0);
    for (uint32_t i = tid; i < ng * N; i += B243_TPB) {
        const uint32_t row = i / N;
        __stcs(dst + i, s_out[row * PITCH + (i - row * N)]);
    }
}
template <int N>
void launch_b243_v2(const uint8_t* words, uint32_t n_groups, uint32_t total_trits, uint8_t* out, cudaStream_t st)
{
    k_words_to_base243_v2<N><<<(n_groups + B243_TPB - 1) / B243_TPB, B243_TPB, 0, st>>>(words, n_groups, total_trits, out);
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
    k_subword_stream_tiled<<<blocks_for(n_words, SW_WORDS), SW_TPB, 0, st>>>(words9, n_words, (uint32_t)N, trits);
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
    if (!n_words) { k_words_to_base243<<<1, TPB, 0, st>>>(words9, n_trits, (uint32_t)N, out); return 1; } // just the count
    const size_t groups = n_words / B243_WPT;
    if (!groups || groups > 0xFFFFFFFFull || ((reinterpret_cast<uintptr_t>(words9) | reinterpret_cast<uintptr_t>(out)) & 3u) || N < 1 || N > 27) {
        k_words_to_base243_tiled<<<blocks_for(n_words, SW_WORDS), SW_TPB, 0, st>>>(words9, n_words, (uint32_t)N, out);
        return 1;
    }
    // 20 words per thread; the last n_words mod 20 words (and the zero padding of the last byte) by the one-byte-per-thread kernel
    using Fn = void (*)(const uint8_t*, uint32_t, uint32_t, uint8_t*, cudaStream_t);
    static const Fn fn[27] = {launch_b243_v2<1>, launch_b243_v2<2>, launch_b243_v2<3>, launch_b243_v2<4>, launch_b243_v2<5>, launch_b243_v2<6>, launch_b243_v2<7>,
                              launch_b243_v2<8>, launch_b243_v2<9>, launch_b243_v2<10>, launch_b243_v2<11>, launch_b243_v2<12>, launch_b243_v2<13>, launch_b243_v2<14>,
                              launch_b243_v2<15>, launch_b243_v2<16>, launch_b243_v2<17>, launch_b243_v2<18>, launch_b243_v2<19>, launch_b243_v2<20>, launch_b243_v2<21>,
                              launch_b243_v2<22>, launch_b243_v2<23>, launch_b243_v2<24>, launch_b243_v2<25>, launch_b243_v2<26>, launch_b243_v2<27>};
    fn[N - 1](words9, (uint32_t)groups, (uint32_t)n_trits, out, st);
    const uint64_t j0 = (uint64_t)(4 * N) * groups, nb = (n_trits + 4) / 5;
    if (j0 < nb) { k_words_to_base243<<<blocks_for(nb - j0, TPB), TPB, 0, st>>>(words9, n_trits, (uint32_t)N, out, j0); return 2; }
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

// =============================================================================================
// the .t3v container's frame records (old/include/t3v_io.hpp:128-160) and its CRC-32 (:14-40).
//   record = n (uint32 LE) | 9n symbol bytes, each % 27 | crc32(payload) ^ (crc32(&n, 4) * 16777619)
// CRC-32 is parallelised through crc(A | B) = x^(8|B|) * crc(A) + crc(B) over GF(2)[x] / P (the identity behind zlib's crc32_combine) and
// through its linearity over GF(2).
// k_t3v_tiles_strided: the payload in tiles of 32 KiB.  Full tiles: one warp per tile, lane-strided 16-byte accesses -> [% 27] -> record, the CRC
//               kept per lane and joined by shuffles (see the kernel).  The partial last tile of a frame: one CTA of the same launch stages it in
//               shared memory (t3v_tile_body: thread = 128-byte segment, slice-by-4).  One CRC per tile -> tile_crc[f * tiles + t]
// k_t3v_finish: one CTA per frame joins the tile CRCs: Horner walks with the constant multiplier x^(8 * 32 KiB) as four 256-entry tables, then
//               pairwise joins with per-level multipliers.  Writes n and the record's CRC (or checks them).
// =============================================================================================
namespace t3c {
namespace {

constexpr uint32_t CRC_POLY = 0xEDB88320u;
constexpr int T3V_SEG = 128, T3V_TPB = 256, T3V_TILE = T3V_SEG * T3V_TPB; // bytes

// reflected polynomials: bit 31 is x^0.  a * b mod P (zlib's multmodp)
__host__ __device__ inline uint32_t crc_mul(uint32_t a, uint32_t b)
{
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1)) == 0) break; }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ CRC_POLY : b >> 1;
    }
    return p;
}
// x^(8 n) mod P by square and multiply
__host__ __device__ inline uint32_t crc_xpow8(uint64_t n_bytes)
{
    uint32_t p = 1u << 31, sq = 1u << 23; // x^0, x^8
    for (uint64_t n = n_bytes; n; n >>= 1) { if (n & 1) p = crc_mul(sq, p); sq = crc_mul(sq, sq); }
    return p;
}
__device__ __forceinline__ uint32_t crc_byte_table(uint32_t i)
{
    uint32_t c = i;
    for (int j = 0; j < 8; ++j) c = (c & 1u) ? (CRC_POLY ^ (c >> 1)) : (c >> 1);
    return c;
}
__device__ __forceinline__ uint32_t mod27_word(uint32_t w)
{
    if ((((w & 0x7F7F7F7Fu) + 0x65656565u) | w) & 0x80808080u) { // some byte >= 27
        uint32_t r = 0;
        for (int q = 0; q < 4; ++q) r |= (((w >> (8 * q)) & 0xFFu) % 27u) << (8 * q);
        return r;
    }
    return w;
}
// Device table buffer (built once per context by build_crc_tables): [0, 1024) slice-by-4 tables T0..T3; then nine shift tables
// of 4 x 256 entries, table j multiplying a 32-bit remainder by x^(8 * 128 * 2^j) (j = 0: one segment ... j = 8: one tile); then
// x^(8 * 2^i), i = 0..31, for the generic shift.
constexpr int CRC_SHIFT0 = 1024, CRC_NSHIFT = 9, CRC_POW0 = CRC_SHIFT0 + CRC_NSHIFT * 1024;
// ... then, for the lane-strided tile kernel: sixteen 256-entry tables X_j[v] = (CRC state after byte v and 15 - j zero bytes, from state 0),
// three shift tables for 16, 32 and 64 bytes, and the state after one tile of zero bytes from 0xFFFFFFFF
constexpr int CRC_NPOW = 48, CRC_X0 = CRC_POW0 + CRC_NPOW, CRC_LANE0 = CRC_X0 + 16 * 256, CRC_K0 = CRC_LANE0 + 3 * 1024;
// ... and the 32-entry heads of the X tables for the four 512-byte pieces of a 2048-byte step: P[p][j][v] = X_j[v] * x^(8 * 512 * (3 - p)), v < 32
constexpr int CRC_P0 = CRC_K0 + 1, CRC_WORDS = CRC_P0 + 4 * 16 * 32;
__device__ __forceinline__ uint32_t crc_shift(const uint32_t* __restrict__ t, uint32_t v) // t = one shift table (shared or global)
{
    return t[v & 0xFFu] ^ t[256 + ((v >> 8) & 0xFFu)] ^ t[512 + ((v >> 16) & 0xFFu)] ^ t[768 + (v >> 24)];
}
__device__ inline uint32_t crc_shift_bytes(const uint32_t* __restrict__ tabs, uint32_t v, uint64_t n_bytes) // v * x^(8 n)
{
    for (int i = 0; n_bytes; ++i, n_bytes >>= 1) if (n_bytes & 1) v = crc_mul(__ldg(tabs + CRC_POW0 + i), v);
    return v;
}
// src / dst: frame f at + f * pitch; *_off = where the 9n payload bytes start inside a frame of src / dst (0 or 4), all 4-byte aligned.
// out: one CRC per tile (the partial last one included) at tile_crc[f * tiles + t]
// Shared-memory layout of the tile body below (uint32 words): slice tables, padded tile, join scratch
constexpr int T3V_SW = 8; // warps (tiles in flight) per CTA of the lane-strided kernel
constexpr int T3V_BODY_WORDS = 4 * 256 + T3V_TPB * 33 + T3V_TPB;
__device__ inline uint32_t crc_mul_bf(uint32_t a, uint32_t b)   // crc_mul without data-dependent control flow
{
    uint32_t p = 0;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) { p ^= b & (0u - (a >> 31)); a <<= 1; b = (b >> 1) ^ (CRC_POLY & (0u - (b & 1u))); }
    return p;
}
// One CTA, one tile (tile t of frame f, possibly the partial last one): stage it in shared memory, one 128-byte segment per thread by
// slice-by-4, then join: the complete segments sit right-aligned in red[] (empty strings in front: their CRC 0 joins as nothing), pairwise
// crc(A | B) = crc(A) x^(8|B|) + crc(B) over 1, 2, 4 ... 128 segments with the shift tables 0..7, and thread 0 appends the last (1..128 byte) one.
__device__ __forceinline__ void t3v_tile_body(uint32_t* __restrict__ sm, uint32_t f, uint32_t t, const uint8_t* __restrict__ src, uint64_t src_pitch, uint32_t src_off,
                                              uint8_t* __restrict__ dst, uint64_t dst_pitch, uint32_t dst_off, uint64_t n_bytes, uint32_t tiles_per_frame, int reduce,
                                              const uint32_t* __restrict__ tabs, uint32_t* __restrict__ tile_crc)
{
    uint32_t* tab = sm;
    uint32_t* tile = sm + 4 * 256;
    uint32_t* red = tile + T3V_TPB * 33;
    const uint32_t tid = threadIdx.x;
    for (int k = 0; k < 4; ++k) tab[256 * k + tid] = __ldg(tabs + 256 * k + tid);
    const uint64_t b0 = (uint64_t)t * T3V_TILE, left = n_bytes - b0, nb = left < T3V_TILE ? left : T3V_TILE; // bytes of this tile
    const uint8_t* s = src + f * src_pitch + src_off + b0;
    uint8_t* d = dst ? dst + f * dst_pitch + dst_off + b0 : nullptr;
    const uint32_t nw = (uint32_t)(nb >> 2);
    for (uint32_t i = tid; i < nw; i += T3V_TPB) {
        uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(s) + i);
        if (reduce) w = mod27_word(w);
        if (d) reinterpret_cast<uint32_t*>(d)[i] = w;
        tile[(i >> 5) * 33 + (i & 31)] = w;
    }
    if (tid < (nb & 3)) { // the last 1..3 bytes of the frame
        uint32_t b = s[4 * nw + tid];
        if (reduce) b %= 27u;
        if (d) d[4 * nw + tid] = (uint8_t)b;
        reinterpret_cast<uint8_t*>(tile)[4 * ((nw >> 5) * 33 + (nw & 31)) + tid] = (uint8_t)b;
    }
    red[tid] = 0;
    __syncthreads();
    const uint32_t nseg = (uint32_t)((nb + T3V_SEG - 1) / T3V_SEG), last_len = (uint32_t)(nb - (uint64_t)(nseg - 1) * T3V_SEG);   // nb >= 1
    uint32_t crc = 0;
    if (tid < nseg) {
        const uint32_t len = tid + 1 < nseg ? T3V_SEG : last_len;
        const uint32_t* p = tile + tid * 33;
        uint32_t c = 0xFFFFFFFFu;
        uint32_t i = 0;
        for (; i + 4 <= len; i += 4) {
            c ^= p[i >> 2];
            c = tab[768 + (c & 0xFFu)] ^ tab[512 + ((c >> 8) & 0xFFu)] ^ tab[256 + ((c >> 16) & 0xFFu)] ^ tab[c >> 24];
        }
        for (; i < len; ++i) c = tab[(c ^ reinterpret_cast<const uint8_t*>(p)[i]) & 0xFFu] ^ (c >> 8);
        crc = c ^ 0xFFFFFFFFu;
        if (tid + 1 < nseg) red[tid + T3V_TPB - (nseg - 1)] = crc;
    }
    __syncthreads();
    for (int j = 0; j < 8; ++j) {
        const uint32_t st = 1u << j;
        if ((tid & (2 * st - 1)) == 0) red[tid] = crc_shift(tabs + CRC_SHIFT0 + 1024 * j, red[tid]) ^ red[tid + st];
        __syncthreads();
    }
    if (tid == nseg - 1) tile_crc[(uint64_t)f * tiles_per_frame + t] = crc_shift_bytes(tabs, red[0], last_len) ^ crc;
}
// the sixteen steps of one tile for one lane (see k_t3v_tiles_strided): sm = P[4][16][32] | per-lane nibble tables of the 2048-byte shift
template <bool S16, bool D16, bool REDUCE, bool COPY>
__device__ __forceinline__ uint32_t t3v_tile_steps(const uint32_t* sm, const uint32_t* __restrict__ tabs, const uint8_t* __restrict__ sp, uint8_t* __restrict__ dp, uint32_t lane)
{
    constexpr int G = 4;   // pieces per step: the step's loads are issued together
    const uint32_t* s2k = sm + 4 * 16 * 32 + lane;
    uint32_t c = 0;
    auto load = [&](uint32_t (&w)[G][4], int k0g) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const uint8_t* a = sp + 512 * (k0g + g);
            if constexpr (S16) { const uint4 q = __ldcs(reinterpret_cast<const uint4*>(a)); w[g][0] = q.x; w[g][1] = q.y; w[g][2] = q.z; w[g][3] = q.w; }
            else {
#pragma unroll
                for (int i = 0; i < 4; ++i) w[g][i] = __ldcs(reinterpret_cast<const uint32_t*>(a) + i);
            }
        }
    };
    auto step = [&](uint32_t (&w)[G][4], int k0g) {
        uint32_t x = 0;                // the state moves 2048 bytes on, then every piece adds its bytes' share at the step's end
#pragma unroll
        for (int n = 0; n < 8; ++n) x ^= s2k[512 * n + 32 * ((c >> (4 * n)) & 15u)];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const uint32_t any = w[g][0] | w[g][1] | w[g][2] | w[g][3];
            bool small = (any & 0xE0E0E0E0u) == 0;                                     // all sixteen bytes < 32
            if constexpr (REDUCE) {   // symbols >= 27 are stored % 27 (rare).  With all bytes < 32, b + 5 reaches bit 5 exactly for b >= 27
                const uint32_t over = ((w[g][0] + 0x05050505u) | (w[g][1] + 0x05050505u) | (w[g][2] + 0x05050505u) | (w[g][3] + 0x05050505u)) & 0x20202020u;
                if (!small || over) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) w[g][i] = mod27_word(w[g][i]);
                    small = true;
                }
            }
            if constexpr (COPY) {
                uint8_t* a = dp + 512 * (k0g + g);
                if constexpr (D16) __stcs(reinterpret_cast<uint4*>(a), make_uint4(w[g][0], w[g][1], w[g][2], w[g][3]));
                else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) __stcs(reinterpret_cast<uint32_t*>(a) + i, w[g][i]);
                }
            }
            if (small) {              // the 32-entry table heads: the lanes of a look-up stay inside 32 consecutive words (no bank conflicts)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t w4 = w[g][i] << 2;                                  // byte offsets into a 32-entry table; no carries between bytes
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        x ^= *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(sm) + 4 * (16 * 32 * g + 32 * (4 * i + q)) + __byte_perm(w4, 0u, 0x4440u | (uint32_t)q));
                }
            } else {                  // a stored byte >= 32 (checking a foreign file): the full tables from global memory
                uint32_t y = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int q = 0; q < 4; ++q) y ^= __ldg(tabs + CRC_X0 + 256 * (4 * i + q) + ((w[g][i] >> (8 * q)) & 0xFFu));
                if (g < 3 && ((3 - g) & 1)) y = crc_shift(tabs + CRC_SHIFT0 + 1024 * 2, y);   // x^(8 * 512)
                if (g < 2) y = crc_shift(tabs + CRC_SHIFT0 + 1024 * 3, y);                     // x^(8 * 1024)
                x ^= y;
            }
        }
        c = x;
    };
    uint32_t w[G][4];
#pragma unroll 1
    for (int k0g = 0; k0g < T3V_TILE / 512; k0g += G) {
        load(w, k0g);
        step(w, k0g);
    }
    return c;
}
// Full tiles, lane-strided: a warp takes one 32 KiB tile in 16 steps of 2048 bytes, four pieces of 512 bytes each, lane l the 16 bytes at
// 512 (4 k + p) + 16 l: perfectly coalesced loads and stores, no shared-memory tile.  A lane keeps the CRC state of "its" bytes as if the other
// lanes' bytes were zero:
//   c <- c * x^(8*2048) + sum_p P_p(16 bytes of piece p)
// CRC is linear over GF(2): P_p = sum_j P[p][j][byte j], the byte's share of the state at the END of the step (its own 15 - j trailing bytes
// and the 3 - p pieces behind it folded into the table), so one shift serves 64 bytes.  Stored symbols are < 27: the 32 lanes of one look-up
// hit at most 27 consecutive words of a 32-entry table head -- no bank conflicts; a piece with a byte >= 32 (only when checking a foreign
// file) takes the 256-entry tables from global memory.  The shift by 2048 bytes is eight nibble look-ups in a per-lane copy of its table.
// 72 shared look-ups per 64 bytes (the 512-byte-step version took 96 and sat on the shared-memory pipe at 0.49 of the HBM roofline).
// The 32 lane states are joined at the end of the tile: crc = sum_l c_l * x^(8*16*(31-l)), pairwise by shuffles.
__global__ void __launch_bounds__(32 * T3V_SW) k_t3v_tiles_strided(const uint8_t* __restrict__ src, uint64_t src_pitch, uint32_t src_off, uint8_t* __restrict__ dst,
                                                                 uint64_t dst_pitch, uint32_t dst_off, uint32_t full_tiles, uint32_t tiles_per_frame, uint32_t n_frames,
                                                                 int reduce, const uint32_t* __restrict__ tabs, uint32_t* __restrict__ tile_crc, uint64_t n_bytes,
                                                                 uint32_t body_ctas)
{
    static_assert(T3V_BODY_WORDS >= 4 * 16 * 32 + 8 * 16 * 32 && 32 * T3V_SW == T3V_TPB, "one shared buffer, one CTA shape for both roles");
    __shared__ __align__(16) uint32_t sm[T3V_BODY_WORDS];
    if (blockIdx.x < body_ctas) {   // the partial last tile of frame blockIdx.x (first in the grid: it is the longest serial piece)
        asm volatile("griddepcontrol.launch_dependents;");
        t3v_tile_body(sm, blockIdx.x, full_tiles, src, src_pitch, src_off, dst, dst_pitch, dst_off, n_bytes, tiles_per_frame, reduce, tabs, tile_crc);
        return;
    }
    uint32_t* sp4 = sm;                    // P[4][16][32]
    uint32_t* s2k = sm + 4 * 16 * 32;      // the shift by 2048 bytes by nibbles, one copy per lane
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, cta = blockIdx.x - body_ctas, n_cta = gridDim.x - body_ctas;
    for (uint32_t i = tid; i < 4 * 16 * 32; i += 32 * T3V_SW) sp4[i] = __ldg(tabs + CRC_P0 + i);
    // the shift by 2048 bytes (table 4: 128 * 2^4 bytes) by nibbles, every entry once per lane (word 32 * (16 n + v) + lane): the eight look-ups of a
    // step, with state-dependent (random) indices, stay inside the lane's own bank
    for (uint32_t i = tid; i < 8 * 16 * 32; i += 32 * T3V_SW) {
        const uint32_t n = i >> 9, v = (i >> 5) & 15u;
        s2k[i] = __ldg(tabs + CRC_SHIFT0 + 1024 * 4 + 256 * (n >> 1) + (v << (4 * (n & 1))));
    }
    __syncthreads();
    // programmatic dependent launch: k_t3v_finish may start its prologue (tables, level multipliers) now; it waits for this grid's tile CRCs
    asm volatile("griddepcontrol.launch_dependents;");
    const uint32_t k0 = __ldg(tabs + CRC_K0);
    const uint64_t total = (uint64_t)full_tiles * n_frames;
    for (uint64_t gw = (uint64_t)cta * T3V_SW + warp; gw < total; gw += (uint64_t)n_cta * T3V_SW) {
        const uint32_t f = (uint32_t)(gw / full_tiles), t = (uint32_t)(gw - (uint64_t)f * full_tiles);
        const uint8_t* sp = src + f * src_pitch + src_off + (uint64_t)t * T3V_TILE + 16u * lane;
        uint8_t* dp = dst ? dst + f * dst_pitch + dst_off + (uint64_t)t * T3V_TILE + 16u * lane : nullptr;
        // 16-byte accesses where the side is 16-byte aligned, else 4-byte ones (a record's payload sits 4 bytes in: t3c.h tells callers to place
        // records at 12 mod 16); the variants are compile-time so that the step loop carries no branches
        const bool s16 = (reinterpret_cast<uintptr_t>(sp) & 15) == 0, d16 = (reinterpret_cast<uintptr_t>(dp) & 15) == 0;
        uint32_t c;
        if (!dp) c = s16 ? t3v_tile_steps<true, true, false, false>(sm, tabs, sp, dp, lane) : t3v_tile_steps<false, true, false, false>(sm, tabs, sp, dp, lane);
        else if (reduce) c = (s16 && d16) ? t3v_tile_steps<true, true, true, true>(sm, tabs, sp, dp, lane) : s16 ? t3v_tile_steps<true, false, true, true>(sm, tabs, sp, dp, lane)
                           : d16 ? t3v_tile_steps<false, true, true, true>(sm, tabs, sp, dp, lane) : t3v_tile_steps<false, false, true, true>(sm, tabs, sp, dp, lane);
        else c = (s16 && d16) ? t3v_tile_steps<true, true, false, true>(sm, tabs, sp, dp, lane) : s16 ? t3v_tile_steps<true, false, false, true>(sm, tabs, sp, dp, lane)
                 : d16 ? t3v_tile_steps<false, true, false, true>(sm, tabs, sp, dp, lane) : t3v_tile_steps<false, false, false, true>(sm, tabs, sp, dp, lane);
        // join the lanes: lane l stands 16 (31 - l) bytes before the end of a step
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const uint32_t other = __shfl_down_sync(0xFFFFFFFFu, c, 1u << j);
            const uint32_t* tb = j < 3 ? tabs + CRC_LANE0 + 1024 * j : tabs + CRC_SHIFT0 + 1024 * (j - 3);  // 16, 32, 64 | 128, 256 bytes
            c = (__ldg(tb + (c & 0xFFu)) ^ __ldg(tb + 256 + ((c >> 8) & 0xFFu)) ^ __ldg(tb + 512 + ((c >> 16) & 0xFFu)) ^ __ldg(tb + 768 + (c >> 24))) ^ other;
        }
        if (lane == 0) tile_crc[(uint64_t)f * tiles_per_frame + t] = c ^ k0 ^ 0xFFFFFFFFu;   // + the initial state carried through the tile, final inversion
    }
}
// One CTA per frame joins the tile CRCs.  check = 0: write n and the record's CRC into rec (record f at rec + f * pitch); check = 1: compare
// them with what the record holds -> ok[f]; crc_out: the payload's plain CRC-32.
// Thread i < 128 walks `per` consecutive full tiles (Horner with the 32 KiB shift table: four look-ups a tile), the tiles right-aligned over the
// 128 threads so that every pairwise join on level j is the same multiplication by x^(8 * 32 KiB * per * 2^j); thread 0 appends the partial last
// tile.  The multipliers come from the other warps meanwhile.  (Bit-serial multiplications cost ~200 dependent instructions each: few threads
// walking far and seven join levels beat 1024 threads and ten levels, 20 -> ~8 us.)
constexpr int FIN_T = 128, FIN_LV = 7, FIN_TPB = FIN_T + 32 * (FIN_LV + 1);
__global__ void __launch_bounds__(FIN_TPB) k_t3v_finish(const uint32_t* __restrict__ tabs, const uint32_t* __restrict__ tile_crc, uint32_t tiles_per_frame, uint64_t n_bytes,
                                                      uint32_t n_words, uint8_t* __restrict__ rec, uint64_t pitch, int check, uint8_t* __restrict__ ok,
                                                      uint32_t* __restrict__ crc_out)
{
    __shared__ uint32_t red[FIN_T];
    __shared__ uint32_t mlev[FIN_LV + 1];
    const uint32_t tid = threadIdx.x, f = blockIdx.x;
    const uint64_t n_full = n_bytes / T3V_TILE, tail = n_bytes - n_full * T3V_TILE;     // full tiles, bytes of the partial last tile
    const uint64_t per = n_full ? (n_full + FIN_T - 1) / FIN_T : 1, pad = FIN_T * per - n_full; // empty places in front
    const uint32_t* p = tile_crc + (uint64_t)f * tiles_per_frame;
    __shared__ uint32_t sh32k[4 * 256];
    __shared__ uint32_t lt[(FIN_LV + 1) * 1024];
    __shared__ uint32_t s_cn;
    // ---- prologue: nothing here depends on the tile kernel; as a programmatic dependent launch it runs while that kernel is still busy
    for (uint32_t i = tid; i < 4 * 256; i += FIN_TPB) sh32k[i] = __ldg(tabs + CRC_SHIFT0 + 1024 * 8 + i);
    if (tid >= FIN_T) {
        // warp j < 7: mlev[j] = x^(8 * 32 KiB * per * 2^j) = product over the set bits b of per of x^(8 * 2^(15 + j + b)); warp 7: x^(8 * tail).
        // One table entry per lane, multiplied up by shuffles.  per < 2^14 (n_bytes < 2^36): 15 + 6 + 13 < CRC_NPOW
        const uint32_t wj = (tid - FIN_T) >> 5, lb = tid & 31u;
        const uint64_t bits = wj < FIN_LV ? per : tail;
        uint32_t m = 1u << 31;                                              // x^0
        if (lb < 16 && ((bits >> lb) & 1)) m = __ldg(tabs + CRC_POW0 + (wj < FIN_LV ? 15 + wj : 0) + lb);
        for (int sh = 8; sh; sh >>= 1) m = crc_mul_bf(m, __shfl_xor_sync(0xFFFFFFFFu, m, sh));
        if (lb == 0) mlev[wj] = m;
    } else if (tid == 0) {
        uint32_t cn = 0xFFFFFFFFu;                                        // crc32 of the four bytes of n
        for (int i = 0; i < 4; ++i) { cn ^= (n_words >> (8 * i)) & 0xFFu; for (int j = 0; j < 8; ++j) cn = (cn & 1u) ? (CRC_POLY ^ (cn >> 1)) : (cn >> 1); }
        s_cn = cn ^ 0xFFFFFFFFu;
    }
    __syncthreads();
    // big frames: the level multipliers as byte tables (a bit-serial product is ~200 dependent instructions, eight of them in a row on the
    // critical path behind the tile kernel; the tables cost 19 products per thread HERE, where they are hidden, and a join becomes four look-ups)
    const bool tables = n_full >= 256;
    if (tables)
        for (uint32_t e = tid; e < (FIN_LV + 1) * 1024; e += FIN_TPB) lt[e] = crc_mul_bf(mlev[e >> 10], (e & 255u) << (8 * ((e >> 8) & 3u)));
    __syncthreads();
    // ---- behind the tile kernel
    asm volatile("griddepcontrol.wait;" ::: "memory");                    // the tile kernel's CRCs (no-op when launched without the attribute)
    if (tid < FIN_T) {
        uint32_t acc = 0;
        const uint64_t v0 = (uint64_t)tid * per;
#pragma unroll 8
        for (uint64_t i = 0; i < per; ++i) {     // the loads do not depend on acc: issued ahead; an empty place adds 0 to a still-zero acc
            const uint32_t val = v0 + i >= pad ? __ldcg(p + (v0 + i - pad)) : 0u;   // L2: written by the grid before this one while this CTA was already resident
            acc = (sh32k[acc & 0xFFu] ^ sh32k[256 + ((acc >> 8) & 0xFFu)] ^ sh32k[512 + ((acc >> 16) & 0xFFu)] ^ sh32k[768 + (acc >> 24)]) ^ val;
        }
        red[tid] = acc;
    }
    __syncthreads();
    for (int j = 0; j < FIN_LV; ++j) {
        const uint32_t st = 1u << j;
        if (tid < FIN_T && (tid & (2 * st - 1)) == 0) red[tid] = (tables ? crc_shift(lt + 1024 * j, red[tid]) : crc_mul_bf(mlev[j], red[tid])) ^ red[tid + st];
        __syncthreads();
    }
    if (tid == 0 && tail) red[0] = (tables ? crc_shift(lt + 1024 * FIN_LV, red[0]) : crc_mul_bf(mlev[FIN_LV], red[0])) ^ __ldcg(p + n_full);
    if (tid == 0 && crc_out) crc_out[f] = red[0];                         // plain crc32 of the payload
    if (tid == 0 && rec) {
        const uint32_t crc = red[0] ^ (s_cn * 16777619u);                   // crc32 of no bytes is 0: red[0] = 0 for an empty frame
        uint8_t* r = rec + f * pitch;
        if (!check) {
            for (int i = 0; i < 4; ++i) { r[i] = (uint8_t)(n_words >> (8 * i)); r[4 + n_bytes + i] = (uint8_t)(crc >> (8 * i)); }
        } else {
            uint32_t n_in = 0, c_in = 0;
            for (int i = 0; i < 4; ++i) { n_in |= (uint32_t)r[i] << (8 * i); c_in |= (uint32_t)r[4 + n_bytes + i] << (8 * i); }
            ok[f] = (n_in == n_words && c_in == crc) ? 1 : 0;
        }
    }
}

} // namespace

// host: the table buffer described above (T3V_CRC_TABLE_WORDS uint32)
void build_crc_tables(uint32_t* h)
{
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int j = 0; j < 8; ++j) c = (c & 1u) ? (CRC_POLY ^ (c >> 1)) : (c >> 1);
        h[i] = c;
    }
    for (int k = 1; k < 4; ++k) for (uint32_t i = 0; i < 256; ++i) h[256 * k + i] = h[h[256 * (k - 1) + i] & 0xFFu] ^ (h[256 * (k - 1) + i] >> 8);
    for (int j = 0; j < CRC_NSHIFT; ++j) {
        const uint32_t m = crc_xpow8((uint64_t)T3V_SEG << j);
        for (int k = 0; k < 4; ++k) for (uint32_t u = 0; u < 256; ++u) h[CRC_SHIFT0 + 1024 * j + 256 * k + u] = crc_mul(m, u << (8 * k));
    }
    uint32_t sq = 1u << 23; // x^8
    for (int i = 0; i < CRC_NPOW; ++i) { h[CRC_POW0 + i] = sq; sq = crc_mul(sq, sq); }
    for (int j = 0; j < 16; ++j) {
        const uint32_t m = crc_xpow8((uint64_t)(15 - j));
        for (uint32_t v = 0; v < 256; ++v) h[CRC_X0 + 256 * j + v] = crc_mul(m, h[v]);
    }
    for (int j = 0; j < 3; ++j) {
        const uint32_t m = crc_xpow8(16ull << j);
        for (int k = 0; k < 4; ++k) for (uint32_t u = 0; u < 256; ++u) h[CRC_LANE0 + 1024 * j + 256 * k + u] = crc_mul(m, u << (8 * k));
    }
    h[CRC_K0] = crc_mul(crc_xpow8(T3V_TILE), 0xFFFFFFFFu);
    for (int p = 0; p < 4; ++p)
        for (int j = 0; j < 16; ++j) {
            const uint32_t m = crc_xpow8((uint64_t)(15 - j) + 512ull * (3 - p));
            for (uint32_t v = 0; v < 32; ++v) h[CRC_P0 + (16 * p + j) * 32 + v] = crc_mul(m, h[v]);
        }
}
size_t crc_table_words() { return CRC_WORDS; }

// k_t3v_finish behind the tile kernel as a programmatic dependent launch: its prologue overlaps the tile kernel's tail, and the
// launch latency disappears from the critical path (the join of one frame is a 13 us serial tail otherwise)
static void launch_finish(unsigned n_frames, cudaStream_t st, const uint32_t* tabs, const uint32_t* tile_crc, uint32_t tiles_per_frame, uint64_t n_bytes, uint32_t n_words,
                          uint8_t* rec, uint64_t pitch, int check, uint8_t* ok, uint32_t* crc_out)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_frames);
    cfg.blockDim = dim3(FIN_TPB);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, k_t3v_finish, tabs, tile_crc, tiles_per_frame, n_bytes, n_words, rec, pitch, check, ok, crc_out) != cudaSuccess) {
        (void)cudaGetLastError();
        k_t3v_finish<<<n_frames, FIN_TPB, 0, st>>>(tabs, tile_crc, tiles_per_frame, n_bytes, n_words, rec, pitch, check, ok, crc_out);
    }
}
// one launch: the full tiles lane-strided, the partial last tile of every frame by a CTA of its own
static int t3v_tiles(const uint8_t* src, uint64_t src_pitch, uint32_t src_off, uint8_t* dst, uint64_t dst_pitch, uint32_t dst_off, uint64_t nb, uint64_t tiles,
                     size_t n_frames, int reduce, const uint32_t* tabs, uint32_t* tile_crc, cudaStream_t st)
{
    if (!tiles) return 0;
    const uint64_t full = nb / T3V_TILE, total = full * n_frames, body = tiles > full ? n_frames : 0;
    uint64_t grid = (total + T3V_SW - 1) / T3V_SW;
    if (grid > 148 * 8) grid = 148 * 8;   // B200: 148 SMs, up to 8 CTAs each; the warps stride over the tiles beyond that
    k_t3v_tiles_strided<<<(unsigned)(grid + body), 32 * T3V_SW, 0, st>>>(src, src_pitch, src_off, dst, dst_pitch, dst_off, (uint32_t)full, (uint32_t)tiles, (uint32_t)n_frames,
                                                                       reduce, tabs, tile_crc, nb, (uint32_t)body);
    return 1;
}
// scratch (uint32): one CRC per tile
size_t t3v_partial_words(size_t n_words, size_t n_frames)
{
    const uint64_t nb = 9ull * n_words, tiles = (nb + T3V_TILE - 1) / T3V_TILE;
    return (size_t)((tiles ? tiles : 1) * n_frames);
}
// words9 (frame f at + f * 9 * stride_words, 4-byte aligned) -> records (record f at + f * record_pitch, 4-byte aligned)
int launch_t3v_records(const uint32_t* tabs, const uint8_t* words9, size_t n_words, size_t stride_words, size_t n_frames, uint8_t* records, size_t record_pitch,
                       uint32_t* partial, cudaStream_t st)
{
    if (!n_frames) return 0;
    const uint64_t nb = 9ull * n_words, tiles = (nb + T3V_TILE - 1) / T3V_TILE, tpf = tiles ? tiles : 1;
    int n = 0;
    n += t3v_tiles(words9, 9ull * stride_words, 0, records, record_pitch, 4, nb, tiles, n_frames, 1, tabs, partial, st);
    launch_finish((unsigned)n_frames, st, tabs, partial, (uint32_t)tpf, nb, (uint32_t)n_words, records, record_pitch, 0, nullptr, nullptr);
    return n + 1;
}
// records -> words9 (may be null: check only) and ok[f] = the record announces n_words and its CRC matches (t3v_read_frame)
int launch_t3v_read(const uint32_t* tabs, const uint8_t* records, size_t record_pitch, size_t n_frames, size_t n_words, uint8_t* words9, size_t stride_words,
                    uint32_t* partial, uint8_t* ok, cudaStream_t st)
{
    if (!n_frames) return 0;
    const uint64_t nb = 9ull * n_words, tiles = (nb + T3V_TILE - 1) / T3V_TILE, tpf = tiles ? tiles : 1;
    int n = 0;
    n += t3v_tiles(records, record_pitch, 4, words9, 9ull * stride_words, 0, nb, tiles, n_frames, 0, tabs, partial, st);
    launch_finish((unsigned)n_frames, st, tabs, partial, (uint32_t)tpf, nb, (uint32_t)n_words, const_cast<uint8_t*>(records), record_pitch, 1, ok, nullptr);
    return n + 1;
}
// plain CRC-32 of n bytes (4-byte aligned) with the same two kernels
int launch_crc32(const uint32_t* tabs, const uint8_t* data, size_t n, uint32_t* partial, uint32_t* out, cudaStream_t st)
{
    const uint64_t tiles = ((uint64_t)n + T3V_TILE - 1) / T3V_TILE, tpf = tiles ? tiles : 1;
    int k = 0;
    k += t3v_tiles(data, 0, 0, nullptr, 0, 0, n, tiles, 1, 0, tabs, partial, st);
    launch_finish(1u, st, tabs, partial, (uint32_t)tpf, n, 0u, nullptr, 0, 0, nullptr, out);
    return k + 1;
}

} // namespace t3c

// =============================================================================================
// image-bridge geometry of the NEW generation (include/io_image.hpp:102-140, 215-235): nearest-neighbour resize,
// centre blit into the S27 canvas, centre-window extraction.  Pure copies: one thread per destination pixel.
// =============================================================================================
namespace t3c {
namespace {
// resize_rgb_nn, :102-124: sx = clamp((int)((x + 0.5) * (double)src_w / dst_w), 0, src_w - 1) -- the same two IEEE double operations
__global__ void k_resize_rgb_nn(const uint8_t* __restrict__ src, int sw, int sh, uint8_t* __restrict__ dst, int dw, int dh)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw) return;
    int sy = (int)__ddiv_rn(__dmul_rn((double)y + 0.5, (double)sh), (double)dh);
    int sx = (int)__ddiv_rn(__dmul_rn((double)x + 0.5, (double)sw), (double)dw);
    sy = min(max(sy, 0), sh - 1);
    sx = min(max(sx, 0), sw - 1);
    const uint8_t* sp = src + ((size_t)sy * sw + sx) * 3;
    uint8_t* dp = dst + ((size_t)y * dw + x) * 3;
    dp[0] = sp[0]; dp[1] = sp[1]; dp[2] = sp[2];
}
// blit_center_rgb, :125-140 (src no wider than the canvas): black canvas, src rows at (x0, y0) = ((cw - sw) / 2, (ch - sh) / 2) clamped at 0,
// rows that fall below the canvas dropped.  One thread per 4 canvas bytes.
__global__ void k_blit_center_rgb(const uint8_t* __restrict__ src, int sw, int sh, uint8_t* __restrict__ dst, int cw, int ch)
{
    const size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4, total = (size_t)cw * ch * 3;
    if (i4 >= total) return;
    const int x0 = max(0, (cw - sw) / 2), y0 = max(0, (ch - sh) / 2);
    const size_t row_bytes = (size_t)cw * 3;
    uint32_t v = 0;
    for (int k = 0; k < 4 && i4 + k < total; ++k) {
        const size_t i = i4 + k, y = i / row_bytes, xb = i - y * row_bytes;
        const long long ys = (long long)y - y0, xs = (long long)xb - 3ll * x0;
        uint32_t b = 0;
        if (ys >= 0 && ys < sh && xs >= 0 && xs < 3ll * sw) b = src[(size_t)ys * sw * 3 + (size_t)xs];
        v |= b << (8 * k);
    }
    if (i4 + 4 <= total && !(reinterpret_cast<uintptr_t>(dst) & 3)) *reinterpret_cast<uint32_t*>(dst + i4) = v;
    else for (int k = 0; k < 4 && i4 + k < total; ++k) dst[i4 + k] = (uint8_t)(v >> (8 * k));
}
// extract_center_q, :215-235 (window no wider than the frame): sub_w x sub_h pixels from (x0, y0), rows below the frame zero
__global__ void k_extract_center_q(const uint16_t* __restrict__ full, int fw, int fh, int sw, int sh, uint16_t* __restrict__ sub)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= sw) return;
    const int x0 = max(0, (fw - sw) / 2), y0 = max(0, (fh - sh) / 2), fy = y + y0;
    uint16_t* d = sub + ((size_t)y * sw + x) * 3;
    if (fy >= fh) { d[0] = d[1] = d[2] = 0; return; }
    const uint16_t* sp = full + ((size_t)fy * fw + x0 + x) * 3;
    d[0] = sp[0]; d[1] = sp[1]; d[2] = sp[2];
}
} // namespace

int launch_resize_rgb_nn(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh, cudaStream_t st)
{
    if (dw <= 0 || dh <= 0) return 0;
    if (sw <= 0 || sh <= 0) { cudaMemsetAsync(dst, 0, (size_t)dw * dh * 3, st); return 0; } // the reference leaves the black image
    k_resize_rgb_nn<<<dim3(blocks_for((size_t)dw, 256), (unsigned)dh), 256, 0, st>>>(src, sw, sh, dst, dw, dh);
    return 1;
}
int launch_blit_center_rgb(const uint8_t* src, int sw, int sh, uint8_t* dst, int cw, int ch, cudaStream_t st)
{
    if (cw <= 0 || ch <= 0) return 0;
    k_blit_center_rgb<<<blocks_for(((size_t)cw * ch * 3 + 3) / 4, 256), 256, 0, st>>>(src, sw, sh, dst, cw, ch);
    return 1;
}
int launch_extract_center_q(const t3c_pixel* full, int fw, int fh, int sw, int sh, t3c_pixel* sub, cudaStream_t st)
{
    if (sw <= 0 || sh <= 0) return 0;
    k_extract_center_q<<<dim3(blocks_for((size_t)sw, 256), (unsigned)sh), 256, 0, st>>>(reinterpret_cast<const uint16_t*>(full), fw, fh, sw, sh, reinterpret_cast<uint16_t*>(sub));
    return 1;
}

} // namespace t3c
