// k_general.cu -- stage-level kernels and the general (any-config) profile codec.
//
// These kernels are correct for every configuration the reference accepts (mixed-k UEP, 2D
// interleave with any tile, beacons, odd sizes).  The tiled fused kernels in k_fast.cu cover the
// headline family (uniform k, 1D, no beacon) at memory-system speed.
#include <cstring>

#include "dev.cuh"
#include "launch.h"

namespace t3c {
namespace {

constexpr int TPB = 128;
inline unsigned blocks_for(size_t n, int per) { return (unsigned)((n + per - 1) / per); }

// ------------------------------------------------------------------------------------------
// K1: RGB8 <-> quant (IMG:47-84,156-192) and 2 px <-> Word27 (OLD:693-747)
// ------------------------------------------------------------------------------------------
__global__ void k_rgb_to_quant(const uint8_t* __restrict__ rgb, size_t n_px, uint16_t* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_px) return;
    int Y, Cb, Cr;
    rgb_to_ycbcr8(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], Y, Cb, Cr);
    out[3 * i] = (uint16_t)quant_y(Y);
    out[3 * i + 1] = (uint16_t)(int16_t)(quant_c_off(Cb) - 40);
    out[3 * i + 2] = (uint16_t)(int16_t)(quant_c_off(Cr) - 40);
}
__global__ void k_quant_to_rgb(const uint16_t* __restrict__ px, size_t n_px, uint8_t* __restrict__ rgb)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_px) return;
    const int Yq = px[3 * i], Cbq = (int16_t)px[3 * i + 1], Crq = (int16_t)px[3 * i + 2];
    int R, G, B;
    ycbcr8_to_rgb(dequant_y(Yq), dequant_c(Cbq), dequant_c(Crq), R, G, B);
    rgb[3 * i] = (uint8_t)R;
    rgb[3 * i + 1] = (uint8_t)G;
    rgb[3 * i + 2] = (uint8_t)B;
}

// word = base-27 digits of A_a + 3^13 A_b (26 trits, T[26]=0), split so that everything stays 32-bit:
//   s0..s3 = digits of A_a;  s4 = trit12(A_a) + 3*(A_b % 9);  s5..s7 = digits of A_b/9;  s8 = A_b / 177147
__device__ __forceinline__ void word_from_values(uint32_t Aa, uint32_t Ab, uint32_t& w0, uint32_t& w1, uint32_t& w2)
{
    const uint32_t q1 = Aa / 27, q2 = Aa / 729, q3 = Aa / 19683, q4 = Aa / 531441;
    w0 = (Aa - 27 * q1) | ((q1 - 27 * q2) << 8) | ((q2 - 27 * q3) << 16) | ((q3 - 27 * q4) << 24);
    const uint32_t lo = Ab % 9, h = Ab / 9;
    const uint32_t h1 = h / 27, h2 = h / 729, h3 = h / 19683;
    w1 = (q4 + 3 * lo) | ((h - 27 * h1) << 8) | ((h1 - 27 * h2) << 16) | ((h2 - 27 * h3) << 24);
    w2 = h3;
}
// inverse on arbitrary bytes: every symbol contributes its low three trits (unpack3, OLD:28-31)
__device__ __forceinline__ void values_from_word(const uint8_t* s, uint32_t& Ya, uint32_t& Cba, uint32_t& Cra, uint32_t& Yb, uint32_t& Cbb, uint32_t& Crb)
{
    uint32_t v[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) v[i] = s[i] % 27u;
    Ya = v[0] + 27 * (v[1] % 9);
    Cba = v[1] / 9 + 3 * v[2];
    Cra = v[3] + 27 * (v[4] % 3);
    Yb = v[4] / 3 + 9 * v[5];
    Cbb = v[6] + 27 * (v[7] % 3);
    Crb = v[7] / 3 + 9 * (v[8] % 9);
}

constexpr int PACK_PX_PER_THREAD = 8, PACK_TPB = 256, PACK_PX_PER_BLOCK = PACK_PX_PER_THREAD * PACK_TPB;
// Full tiles: 2048 pixels (12288 B) -> 1024 words (9216 B) staged through shared memory so that both
// the global loads and stores are 128-bit and fully coalesced.
__global__ void __launch_bounds__(PACK_TPB) k_pack_pixels(const uint16_t* __restrict__ px, size_t n_px, uint8_t* __restrict__ words)
{
    __shared__ __align__(16) uint32_t sin[PACK_PX_PER_BLOCK * 6 / 4];
    __shared__ __align__(16) uint32_t sout[PACK_PX_PER_BLOCK / 2 * 9 / 4];
    const size_t p0 = (size_t)blockIdx.x * PACK_PX_PER_BLOCK;
    const int t = threadIdx.x;
    if (p0 + PACK_PX_PER_BLOCK <= n_px) {
        const uint4* src = reinterpret_cast<const uint4*>(px + 3 * p0);
        uint4* s4 = reinterpret_cast<uint4*>(sin);
#pragma unroll
        for (int j = 0; j < 3; ++j) s4[t + PACK_TPB * j] = __ldg(src + t + PACK_TPB * j);
        __syncthreads();
        const uint16_t* me = reinterpret_cast<const uint16_t*>(sin) + 24 * t;
        uint4 a = reinterpret_cast<const uint4*>(me)[0], b = reinterpret_cast<const uint4*>(me)[1], c = reinterpret_cast<const uint4*>(me)[2];
        const uint32_t h[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
        uint32_t A[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            // pixel i = halfwords 3i, 3i+1, 3i+2
            auto hw = [&](int k) { return (uint32_t)((h[k >> 1] >> ((k & 1) * 16)) & 0xFFFF); };
            A[i] = pixel_value(hw(3 * i), (int16_t)hw(3 * i + 1), (int16_t)hw(3 * i + 2));
        }
        uint32_t o[9];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            uint32_t w0, w1, w2;
            word_from_values(A[2 * w], A[2 * w + 1], w0, w1, w2);
            // 9 bytes at byte offset 9w of the thread's 36-byte output
            const int off = 9 * w;
            uint64_t lo = (uint64_t)w0 | ((uint64_t)w1 << 32);
#pragma unroll
            for (int bq = 0; bq < 9; ++bq) {
                const uint32_t byte = bq < 8 ? (uint32_t)((lo >> (8 * bq)) & 0xFF) : w2;
                const int pos = off + bq;
                if (bq == 0 && (pos & 3) == 0) o[pos >> 2] = 0;
                if ((pos & 3) == 0) o[pos >> 2] = byte; else o[pos >> 2] |= byte << (8 * (pos & 3));
            }
        }
#pragma unroll
        for (int j = 0; j < 9; ++j) sout[9 * t + j] = o[j];
        __syncthreads();
        uint4* dst = reinterpret_cast<uint4*>(words + 9 * (p0 / 2));
        const uint4* so4 = reinterpret_cast<const uint4*>(sout);
        for (int j = t; j < PACK_PX_PER_BLOCK / 2 * 9 / 16; j += PACK_TPB) dst[j] = so4[j];
    } else { // ragged tail: one word per thread, scalar
        const size_t n_words = (n_px + 1) / 2;
        for (size_t w = p0 / 2 + t; w < n_words && w < p0 / 2 + PACK_PX_PER_BLOCK / 2; w += PACK_TPB) {
            const size_t a = 2 * w, b = 2 * w + 1;
            const uint32_t Aa = pixel_value(px[3 * a], (int16_t)px[3 * a + 1], (int16_t)px[3 * a + 2]);
            const uint32_t Ab = b < n_px ? pixel_value(px[3 * b], (int16_t)px[3 * b + 1], (int16_t)px[3 * b + 2])
                                         : pixel_value(0, 0, 0); // pairs with PixelYCbCrQuant{}, OLD:730
            uint32_t w0, w1, w2;
            word_from_values(Aa, Ab, w0, w1, w2);
            uint8_t* o = words + 9 * w;
#pragma unroll
            for (int i = 0; i < 4; ++i) { o[i] = (uint8_t)(w0 >> (8 * i)); o[4 + i] = (uint8_t)(w1 >> (8 * i)); }
            o[8] = (uint8_t)w2;
        }
    }
}

__global__ void __launch_bounds__(PACK_TPB) k_unpack_pixels(const uint8_t* __restrict__ words, size_t n_words, uint16_t* __restrict__ px)
{
    __shared__ __align__(16) uint32_t sin[PACK_PX_PER_BLOCK / 2 * 9 / 4];
    __shared__ __align__(16) uint32_t sout[PACK_PX_PER_BLOCK * 6 / 4];
    constexpr int WPB = PACK_PX_PER_BLOCK / 2;
    const size_t w0 = (size_t)blockIdx.x * WPB;
    const int t = threadIdx.x;
    if (w0 + WPB <= n_words) {
        const uint4* src = reinterpret_cast<const uint4*>(words + 9 * w0);
        uint4* s4 = reinterpret_cast<uint4*>(sin);
        for (int j = t; j < WPB * 9 / 16; j += PACK_TPB) s4[j] = __ldg(src + j);
        __syncthreads();
        uint32_t in[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) in[j] = sin[9 * t + j];
        uint16_t* me = reinterpret_cast<uint16_t*>(sout) + 24 * t;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            uint8_t s[9];
#pragma unroll
            for (int bq = 0; bq < 9; ++bq) { const int pos = 9 * w + bq; s[bq] = (uint8_t)(in[pos >> 2] >> (8 * (pos & 3))); }
            uint32_t Ya, Cba, Cra, Yb, Cbb, Crb;
            values_from_word(s, Ya, Cba, Cra, Yb, Cbb, Crb);
            me[6 * w + 0] = (uint16_t)Ya; me[6 * w + 1] = (uint16_t)(int16_t)((int)Cba - 40); me[6 * w + 2] = (uint16_t)(int16_t)((int)Cra - 40);
            me[6 * w + 3] = (uint16_t)Yb; me[6 * w + 4] = (uint16_t)(int16_t)((int)Cbb - 40); me[6 * w + 5] = (uint16_t)(int16_t)((int)Crb - 40);
        }
        __syncthreads();
        uint4* dst = reinterpret_cast<uint4*>(px + 6 * w0);
        const uint4* so4 = reinterpret_cast<const uint4*>(sout);
#pragma unroll
        for (int j = 0; j < 3; ++j) dst[t + PACK_TPB * j] = so4[t + PACK_TPB * j];
    } else {
        for (size_t w = w0 + t; w < n_words && w < w0 + WPB; w += PACK_TPB) {
            uint8_t s[9];
            for (int i = 0; i < 9; ++i) s[i] = words[9 * w + i];
            uint32_t Ya, Cba, Cra, Yb, Cbb, Crb;
            values_from_word(s, Ya, Cba, Cra, Yb, Cbb, Crb);
            uint16_t* o = px + 6 * w;
            o[0] = (uint16_t)Ya; o[1] = (uint16_t)(int16_t)((int)Cba - 40); o[2] = (uint16_t)(int16_t)((int)Cra - 40);
            o[3] = (uint16_t)Yb; o[4] = (uint16_t)(int16_t)((int)Cbb - 40); o[5] = (uint16_t)(int16_t)((int)Crb - 40);
        }
    }
}

// =============================================================================================
// RAW mode v2 (BASELINE config 3): persistent warps, 512-pixel tiles (3072 B of PixelYCbCrQuant <-> 2304 B of Word27:
// both multiples of 16, so no edge handling), tile I/O on the bulk-async copy engine with the next tile prefetched,
// and the per-word arithmetic cut from ~57 / ~81 to ~25 instructions per pixel (the v1 kernels were issue-bound:
// profiles/r01c_raw_v1_ncu_summary.txt).  The ragged tail (< 512 pixels) keeps the v1 kernels.
// =============================================================================================
namespace raw2 {
constexpr int TILE_PX = 512, TILE_IN = 6 * TILE_PX, TILE_OUT = 9 * TILE_PX / 2, WARPS = 32, TPB = 32 * WARPS;
constexpr int WARP_BYTES = TILE_IN + TILE_OUT + 16; // quantised pixels | words | mbarrier
constexpr int SMEM = WARPS * WARP_BYTES;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) { asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// two pixels (six halfwords: Ya Cba Cra Yb Cbb Crb) -> nine symbols, pack_two_pixels OLD:693-705 with i2tr's wrap
// (every field keeps its low 5 / 4 trits, OLD:675-682)
__device__ __forceinline__ void word_from_pixels(uint32_t h0, uint32_t h1, uint32_t h2, uint32_t& o0, uint32_t& o1, uint32_t& o2)
{
    // fields as 16-bit values, chroma offset by 40; in range (Y < 243, chroma + 40 < 81) nothing wraps
    uint32_t Ya = h0 & 0xFFFFu, Cba = ((h0 >> 16) + 40u) & 0xFFFFu, Cra = (h1 + 40u) & 0xFFFFu;
    uint32_t Yb = h1 >> 16, Cbb = (h2 + 40u) & 0xFFFFu, Crb = ((h2 >> 16) + 40u) & 0xFFFFu;
    if (((Ya + 13u) | (Yb + 13u) | (Cba + 175u) | (Cra + 175u) | (Cbb + 175u) | (Crb + 175u)) >> 8) { // rare: i2tr keeps the low trits of the 32-bit value
        Ya %= 243u; Yb %= 243u;
        Cba = (uint32_t)((int)(int16_t)(h0 >> 16) + 40) % 81u; Cra = (uint32_t)((int)(int16_t)(h1 & 0xFFFFu) + 40) % 81u;
        Cbb = (uint32_t)((int)(int16_t)(h2 & 0xFFFFu) + 40) % 81u; Crb = (uint32_t)((int)(int16_t)(h2 >> 16) + 40) % 81u;
    }
    // all fields < 243: one IMAD.HI per quotient (ceil(2^32/d) is exact far beyond this range)
    const uint32_t ya1 = __umulhi(Ya, 159072863u), cba1 = __umulhi(Cba, 1431655766u), cra1 = __umulhi(Cra, 159072863u);
    const uint32_t yb1 = __umulhi(Yb, 477218589u), cbb1 = __umulhi(Cbb, 159072863u), crb1 = __umulhi(Crb, 477218589u);
    const uint32_t s0 = Ya - 27u * ya1, s1 = ya1 + 9u * (Cba - 3u * cba1), s2 = cba1, s3 = Cra - 27u * cra1;
    const uint32_t s4 = cra1 + 3u * (Yb - 9u * yb1), s5 = yb1, s6 = Cbb - 27u * cbb1, s7 = cbb1 + 3u * (Crb - 9u * crb1), s8 = crb1;
    o0 = s0 + (s1 << 8) + (s2 << 16) + (s3 << 24);
    o1 = s4 + (s5 << 8) + (s6 << 16) + (s7 << 24);
    o2 = s8;
}
// nine symbols -> two pixels as three words of halfwords; every symbol contributes its low three trits (unpack3, OLD:28-31)
__device__ __forceinline__ void pixels_from_word(uint32_t w0, uint32_t w1, uint32_t s8, uint32_t& h0, uint32_t& h1, uint32_t& h2)
{
    if (((w0 | w1 | s8) & 0xE0E0E0E0u) || (((w0 & 0x1F1F1F1Fu) + 0x05050505u) | ((w1 & 0x1F1F1F1Fu) + 0x05050505u) | (s8 + 5u)) & 0x20202020u) {
        // some byte is >= 27: reduce every symbol mod 27 (rare)
        uint32_t a = 0, b = 0;
        for (int i = 0; i < 4; ++i) { a |= (((w0 >> (8 * i)) & 0xFFu) % 27u) << (8 * i); b |= (((w1 >> (8 * i)) & 0xFFu) % 27u) << (8 * i); }
        w0 = a; w1 = b; s8 = (s8 & 0xFFu) % 27u;
    }
    const uint32_t v0 = w0 & 0xFFu, v1 = (w0 >> 8) & 0xFFu, v2 = (w0 >> 16) & 0xFFu, v3 = w0 >> 24;
    const uint32_t v4 = w1 & 0xFFu, v5 = (w1 >> 8) & 0xFFu, v6 = (w1 >> 16) & 0xFFu, v7 = w1 >> 24;
    const uint32_t q1 = __umulhi(v1, 477218589u), q4 = __umulhi(v4, 1431655766u), q7 = __umulhi(v7, 1431655766u); // v < 27
    const uint32_t Ya = v0 + 27u * (v1 - 9u * q1), Cba = q1 + 3u * v2 - 40u, Cra = v3 + 27u * (v4 - 3u * q4) - 40u;
    const uint32_t Yb = q4 + 9u * v5, Cbb = v6 + 27u * (v7 - 3u * q7) - 40u, Crb = q7 + 9u * (s8 - 9u * __umulhi(s8, 477218589u)) - 40u;
    h0 = Ya | (Cba << 16);
    h1 = (Cra & 0xFFFFu) | (Yb << 16);
    h2 = (Cbb & 0xFFFFu) | (Crb << 16);
}

__global__ void __launch_bounds__(TPB, 1) k_pack_pixels_v2(const uint8_t* __restrict__ px, uint32_t n_tiles, uint8_t* __restrict__ words)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* IN = smem + warp * WARP_BYTES;
    uint8_t* OUT = IN + TILE_IN;
    const uint32_t bar = smem_u32(OUT + TILE_OUT);
    if (lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    // tiles are dealt round-robin over all warps of the grid: at any time the machine works on one moving window of
    // consecutive tiles (DRAM-friendly); tile boundaries are 16-byte aligned, so nothing is carried between tiles
    const uint32_t lo = blockIdx.x + gridDim.x * warp, hi = n_tiles, step = gridDim.x * WARPS;
    auto fetch = [&](uint32_t t) { fence_async_smem(); mbar_expect_tx(bar, TILE_IN); bulk_g2s(smem_u32(IN), px + (size_t)TILE_IN * t, TILE_IN, bar); };
    if (lo < hi && lane == 0) fetch(lo);
    uint32_t phase = 0;
    for (uint32_t t = lo; t < hi; t += step) {
        mbar_wait(bar, phase);
        phase ^= 1;
        uint4 q[6]; // two passes: this lane's 2 x 4 words = 2 x 48 bytes of pixels
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int j = 0; j < 3; ++j) q[3 * p + j] = *reinterpret_cast<const uint4*>(IN + 1536 * p + 48 * lane + 16 * j);
        __syncwarp();
        if (t + step < hi && lane == 0) fetch(t + step);           // IN is in registers: next tile on its way
        if (lane == 0) bulk_wait_read();                           // the previous tile's store has read OUT
        __syncwarp();
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const uint32_t h[12] = {q[3 * p].x, q[3 * p].y, q[3 * p].z, q[3 * p].w, q[3 * p + 1].x, q[3 * p + 1].y, q[3 * p + 1].z, q[3 * p + 1].w,
                                    q[3 * p + 2].x, q[3 * p + 2].y, q[3 * p + 2].z, q[3 * p + 2].w};
            uint32_t o[9];
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                uint32_t a, b, c;
                word_from_pixels(h[3 * w], h[3 * w + 1], h[3 * w + 2], a, b, c);
                // nine bytes at byte offset 9w of the lane's 36 output bytes
                if (w == 0) { o[0] = a; o[1] = b; o[2] = c; }
                if (w == 1) { o[2] |= a << 8; o[3] = (a >> 24) | (b << 8); o[4] = (b >> 24) | (c << 8); }
                if (w == 2) { o[4] |= a << 16; o[5] = (a >> 16) | (b << 16); o[6] = (b >> 16) | (c << 16); }
                if (w == 3) { o[6] |= a << 24; o[7] = (a >> 8) | (b << 24); o[8] = (b >> 8) | (c << 24); }
            }
            uint32_t* d = reinterpret_cast<uint32_t*>(OUT + 1152 * p + 36 * lane);
#pragma unroll
            for (int j = 0; j < 9; ++j) d[j] = o[j];
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) { bulk_s2g(words + (size_t)TILE_OUT * t, smem_u32(OUT), TILE_OUT); bulk_commit(); }
    }
    if (lane == 0) bulk_wait_all();
}

__global__ void __launch_bounds__(TPB, 1) k_unpack_pixels_v2(const uint8_t* __restrict__ words, uint32_t n_tiles, uint8_t* __restrict__ px)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* OUT = smem + warp * WARP_BYTES;       // pixels
    uint8_t* IN = OUT + TILE_IN;                   // words
    const uint32_t bar = smem_u32(IN + TILE_OUT);
    if (lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    const uint32_t lo = blockIdx.x + gridDim.x * warp, hi = n_tiles, step = gridDim.x * WARPS; // round-robin, see k_pack_pixels_v2
    auto fetch = [&](uint32_t t) { fence_async_smem(); mbar_expect_tx(bar, TILE_OUT); bulk_g2s(smem_u32(IN), words + (size_t)TILE_OUT * t, TILE_OUT, bar); };
    if (lo < hi && lane == 0) fetch(lo);
    uint32_t phase = 0;
    for (uint32_t t = lo; t < hi; t += step) {
        mbar_wait(bar, phase);
        phase ^= 1;
        uint32_t in[18]; // two passes: this lane's 2 x 4 words = 2 x 36 bytes of symbols
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int j = 0; j < 9; ++j) in[9 * p + j] = *reinterpret_cast<const uint32_t*>(IN + 1152 * p + 36 * lane + 4 * j);
        __syncwarp();
        if (t + step < hi && lane == 0) fetch(t + step);
        if (lane == 0) bulk_wait_read();
        __syncwarp();
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const uint32_t* s = in + 9 * p;
            uint32_t h[12];
            // word w = bytes 9w .. 9w+8 of the lane's 36 bytes
            pixels_from_word(s[0], s[1], s[2] & 0xFFu, h[0], h[1], h[2]);
            pixels_from_word(__funnelshift_r(s[2], s[3], 8), __funnelshift_r(s[3], s[4], 8), (s[4] >> 8) & 0xFFu, h[3], h[4], h[5]);
            pixels_from_word(__funnelshift_r(s[4], s[5], 16), __funnelshift_r(s[5], s[6], 16), (s[6] >> 16) & 0xFFu, h[6], h[7], h[8]);
            pixels_from_word(__funnelshift_r(s[6], s[7], 24), __funnelshift_r(s[7], s[8], 24), s[8] >> 24, h[9], h[10], h[11]);
            uint4* d = reinterpret_cast<uint4*>(OUT + 1536 * p + 48 * lane);
            d[0] = make_uint4(h[0], h[1], h[2], h[3]); d[1] = make_uint4(h[4], h[5], h[6], h[7]); d[2] = make_uint4(h[8], h[9], h[10], h[11]);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) { bulk_s2g(px + (size_t)TILE_IN * t, smem_u32(OUT), TILE_IN); bulk_commit(); }
    }
    if (lane == 0) bulk_wait_all();
}
} // namespace raw2

__global__ void k_init_status(uint32_t* st, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 2 * n) st[i] = (i & 1) ? 0u : 1u;
}
__global__ void k_mod27(const uint8_t* __restrict__ in, size_t n, uint8_t* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] % 27;
}

// ------------------------------------------------------------------------------------------
// Block-level RS (RSCodec::encode_block / decode_block), one thread per block
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_rs_encode_blocks(const RowTable* __restrict__ tab, int k, const uint8_t* __restrict__ data,
                                                          size_t n, uint8_t* __restrict__ out)
{
    __shared__ uint64_t row[24 * kVals];
    for (int i = threadIdx.x; i < k * kVals; i += TPB) row[i] = tab->e[i / kVals][i % kVals];
    __syncthreads();
    const size_t blk = (size_t)blockIdx.x * TPB + threadIdx.x;
    if (blk >= n) return;
    const int r = 26 - k;
    Planes acc{0, 0};
    for (int i = 0; i < k; ++i) {
        const uint32_t raw = data[blk * k + i];
        out[blk * 26 + i] = (uint8_t)raw;      // systematic part is copied verbatim (OLD:532)
        gf3_add(acc, row[i * kVals + raw % 27]);
    }
    const uint32_t lo = planes_to_sym4_lo(acc), hi = planes_to_sym4_hi(acc);
    for (int j = 0; j < r; ++j) out[blk * 26 + k + j] = (uint8_t)((j < 4 ? lo >> (8 * j) : hi >> (8 * (j - 4))) & 0xFF);
}

__global__ void __launch_bounds__(TPB) k_rs_decode_blocks(const GfTables* __restrict__ gf, int k, int fixed, uint8_t* __restrict__ inout,
                                                          size_t n, uint8_t* __restrict__ out_k, uint8_t* __restrict__ ok)
{
    __shared__ GfTables sg;
    load_gf(sg, gf);
    __syncthreads();
    const size_t blk = (size_t)blockIdx.x * TPB + threadIdx.x;
    if (blk >= n) return;
    uint8_t c[26];
    for (int i = 0; i < 26; ++i) c[i] = inout[blk * 26 + i] % 27;
    const bool good = rs_decode_thread(sg, c, k, fixed != 0);
    for (int i = 0; i < 26; ++i) inout[blk * 26 + i] = c[i];
    for (int i = 0; i < k; ++i) out_k[blk * k + i] = good ? c[i] : 0;
    ok[blk] = good ? 1 : 0;
}

template <typename I>
__global__ void k_perm2d(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, I n, I area, uint32_t w)
{
    const I i = (I)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[perm2d<I>(i, n, area, w)];
}

// 32-bit index arithmetic is enough when every symbol / trit / byte index of the super-frame stays below 2^31
static inline bool small_geom(const Geom& g) { return 27 * g.n_words + 64 < (1ull << 31) && 9 * g.n_out + 64 < (1ull << 31) && g.l_exp + 64 < (1ull << 31); }

// ------------------------------------------------------------------------------------------
// K5: super-frame header (OLD:155-380, 1142-1158) -- one warp
// ------------------------------------------------------------------------------------------
__device__ void header_symbols(const t3c_config& c, uint32_t frame_seq, uint32_t hash, uint8_t* p, uint32_t magic = 0x0A2, uint32_t version = 1)
{
    // HeaderCodec::pack, OLD:208-289.  at(i,v) narrows v to uint8 before %27.
    for (int i = 0; i < 27; ++i) p[i] = 0;
    p[0] = magic % 27; p[1] = (magic / 27) % 27; p[2] = (uint8_t)(version % 27) % 27; p[3] = c.profile % 27;
    for (int g = 0; g < 3; ++g) p[4 + g] = (uint8_t)(9 * (c.uep[3 * g] % 3) + 3 * (c.uep[3 * g + 1] % 3) + (c.uep[3 * g + 2] % 3));
    p[7] = c.tile_w % 27; p[8] = c.tile_h % 27;
    p[9] = c.seed_a % 27; p[10] = c.seed_b % 27; p[11] = c.seed_s0 % 27;
    const uint32_t sub = c.subword == 24 ? 1 : c.subword == 21 ? 2 : c.subword == 18 ? 3 : c.subword == 15 ? 4 : 0;
    p[12] = (uint8_t)((sub + 9 * (c.centered ? 1 : 0)) % 27);
    p[13] = hash % 27; p[14] = (hash / 27) % 27; p[15] = (hash / 729) % 27;
    p[16] = c.coset % 3;
    p[17] = frame_seq % 27; p[18] = (frame_seq / 27) % 27; p[19] = (frame_seq / 729) % 27;
    p[23] = c.beacon_enabled ? 1 : 0; p[24] = c.beacon_slot % 27;
    p[25] = (uint8_t)(c.beacon_period < 26 ? c.beacon_period : 26);
}
__device__ void header_crc(const uint8_t* p, uint8_t* r) // CRC3::rem12 over 69 trits + 12 zeros, OLD:176-205
{
    for (int i = 0; i < 12; ++i) r[i] = 0;
    for (int q = 0; q < 27 + 4; ++q) {          // 23 symbols, then 4 all-zero "symbols" = the 12 flush steps
        uint32_t s = 0;
        if (q < 27) { if (q == 20 || q == 21 || q == 22 || q == 26) continue; s = p[q]; }
        for (int c = 0; c < 3; ++c) {
            const uint32_t in = c == 0 ? s % 3 : (c == 1 ? (s / 3) % 3 : (s / 9) % 3);
            const uint32_t fb = (in + r[11]) % 3;
            uint8_t nx[12] = {(uint8_t)fb, r[0], r[1], (uint8_t)((r[2] + fb) % 3), (uint8_t)((r[3] + fb) % 3), r[4], r[5],
                              (uint8_t)((r[6] + fb) % 3), r[7], r[8], r[9], r[10]};
            for (int i = 0; i < 12; ++i) r[i] = nx[i];
        }
    }
}
// warp-cooperative: lane 0 builds the 27 symbols + CRC, lanes 0..15 each produce one parity symbol
// of the two RS(26,18) blocks.  hdr27/coded52 point to shared or global memory.
__device__ void header_emit_warp(const t3c_config& c, int arith, const GfTables* gf, const RsTables* rs, uint8_t* hdr27, uint8_t* coded52)
{
    const int lane = threadIdx.x & 31;
    if (lane == 0) {
        uint8_t p[27], r[12];
        header_symbols(c, 0, 0, p);
        header_crc(p, r);
        p[20] = r[0] + 3 * r[1] + 9 * r[2]; p[21] = r[3] + 3 * r[4] + 9 * r[5];
        p[22] = r[6] + 3 * r[7] + 9 * r[8]; p[26] = r[9] + 3 * r[10] + 9 * r[11];
        for (int i = 0; i < 27; ++i) hdr27[i] = p[i];
        for (int i = 0; i < 18; ++i) coded52[i] = p[i];                       // A = sym[0..17]
        for (int i = 0; i < 18; ++i) coded52[26 + i] = i < 9 ? p[18 + i] : 0; // B = sym[18..26] | 0^9
    }
    __syncwarp();
    if (lane < 16) {
        const int blk = lane >> 3, j = lane & 7;
        uint32_t acc = 0;
        for (int i = 0; i < 18; ++i) acc = gf->add[acc * 27 + gf->mul[coded52[26 * blk + i] * 27 + rs->par[arith][3][i][j]]];
        coded52[26 * blk + 18 + j] = (uint8_t)acc;
    }
    __syncwarp();
}
__global__ void k_header_emit(t3c_config cfg, int arith, const GfTables* gf, const RsTables* rs, uint8_t* hdr27, uint8_t* coded52)
{
    header_emit_warp(cfg, arith, gf, rs, hdr27, coded52);
}

// HeaderCodec::pack / check / unpack on one 27-symbol header (OLD:208-379), CRC3::rem12 on a trit string (OLD:176-205): the L1 names
// of the reference's public surface, one thread each (they exist for drop-in callers and tests, not for throughput)
__global__ void k_header_pack(t3c_config cfg, uint32_t magic, uint32_t version, uint32_t hash, uint32_t seq, uint8_t* hdr27)
{
    uint8_t p[27], r[12];
    header_symbols(cfg, seq, hash, p, magic, version);
    header_crc(p, r);
    p[20] = r[0] + 3 * r[1] + 9 * r[2]; p[21] = r[3] + 3 * r[4] + 9 * r[5];
    p[22] = r[6] + 3 * r[7] + 9 * r[8]; p[26] = r[9] + 3 * r[10] + 9 * r[11];
    for (int i = 0; i < 27; ++i) hdr27[i] = p[i];
}
// out4 = {magic, version, band_map_hash, frame_seq}; *ok = HeaderCodec::check; unpack fills cfg whether or not the CRC holds (as the reference's unpack does)
__global__ void k_header_check_unpack(const uint8_t* __restrict__ sym27, t3c_config* cfg, uint32_t* out4, int* ok)
{
    uint8_t p[27], r[12];
    for (int i = 0; i < 27; ++i) p[i] = sym27[i];
    header_crc(p, r);   // unpack3 on the symbols as they are (OLD:296): digits of p[i] mod 27
    bool good = true;
    const int slots[4] = {20, 21, 22, 26};
    for (int i = 0; i < 4; ++i) {
        const uint32_t s = p[slots[i]];
        good = good && s % 3 == r[3 * i] && (s / 3) % 3 == r[3 * i + 1] && (s / 9) % 3 == r[3 * i + 2];
    }
    *ok = good ? 1 : 0;
    for (int i = 0; i < 27; ++i) p[i] %= 27;  // rd(i) = symbols[i] % 27, OLD:324
    t3c_config h = *cfg;
    h.profile = p[3] % 5;
    for (int g = 0; g < 3; ++g) { const uint32_t v = p[4 + g]; h.uep[3 * g] = v % 3; h.uep[3 * g + 1] = (v / 3) % 3; h.uep[3 * g + 2] = (v / 9) % 3; }
    h.tile_w = p[7]; h.tile_h = p[8];
    h.seed_a = p[9]; h.seed_b = p[10]; h.seed_s0 = p[11];
    const uint32_t sub = p[12] % 9, cen = (p[12] / 9) % 3;
    h.subword = sub == 1 ? 24 : sub == 2 ? 21 : sub == 3 ? 18 : sub == 4 ? 15 : 27;
    h.centered = cen != 0;
    h.coset = p[16] % 3;
    h.beacon_enabled = p[23] != 0; h.beacon_slot = p[24] % 9; h.beacon_period = p[25];
    *cfg = h;
    out4[0] = p[0] + 27u * p[1]; out4[1] = p[2]; out4[2] = p[13] + 27u * p[14] + 729u * p[15]; out4[3] = p[17] + 27u * p[18] + 729u * p[19];
}
__global__ void k_crc3_rem12(const uint8_t* __restrict__ trits, size_t n, uint8_t* out12)
{
    uint8_t r[12];
    for (int i = 0; i < 12; ++i) r[i] = 0;
    for (size_t q = 0; q < n + 12; ++q) {
        const uint32_t in = q < n ? trits[q] : 0u;   // the reference adds the trit as it is and reduces mod 3 (OLD:184)
        const uint32_t fb = (in + r[11]) % 3;
        uint8_t nx[12] = {(uint8_t)fb, r[0], r[1], (uint8_t)((r[2] + fb) % 3), (uint8_t)((r[3] + fb) % 3), r[4], r[5],
                          (uint8_t)((r[6] + fb) % 3), r[7], r[8], r[9], r[10]};
        for (int i = 0; i < 12; ++i) r[i] = nx[i];
    }
    for (int i = 0; i < 12; ++i) out12[i] = r[i];
}
// scramble_symbol / descramble_symbol (OLD:81-94) applied to a sequence: symbol i sees the state after i + 1 steps of st <- (a*st + b) % 3
// (st8: the eventually periodic state sequence, scrambler_states)
struct ScrStates { uint8_t st[8]; };
__global__ void k_scramble(uint8_t* __restrict__ syms, size_t n, ScrStates S, const GfTables* __restrict__ gf, int inverse)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t st = i < 2 ? S.st[i] : S.st[2 + (i - 2) % 6];
    const uint32_t s = syms[i] % 27u;
    syms[i] = inverse ? gf->dsc[st][s] : gf->scr[st][s];
}

// read_and_decode_header_from_words (OLD:918-937) + HeaderCodec::check/unpack (OLD:290-379)
__global__ void k_header_parse(const GfTables* __restrict__ gf, int fixed, const uint8_t* __restrict__ words, size_t n_words,
                               t3c_config* out, int* ok)
{
    __shared__ GfTables sg;
    __shared__ uint8_t blk[2][26];
    __shared__ int good[2];
    load_gf(sg, gf);
    __syncthreads();
    if (n_words < 6) { if (threadIdx.x == 0) *ok = 0; return; }
    if (threadIdx.x < 2) {
        uint8_t c[26];
        for (int i = 0; i < 26; ++i) c[i] = words[26 * threadIdx.x + i] % 27; // A = sy[0..25], B = sy[26..51]
        good[threadIdx.x] = rs_decode_thread(sg, c, 18, fixed != 0) ? 1 : 0;
        for (int i = 0; i < 26; ++i) blk[threadIdx.x][i] = c[i];
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    if (!good[0] || !good[1]) { *ok = 0; return; }
    uint8_t p[27], r[12];
    for (int i = 0; i < 18; ++i) p[i] = blk[0][i];
    for (int i = 0; i < 9; ++i) p[18 + i] = blk[1][i];
    header_crc(p, r);
    const uint8_t want[4] = {p[20], p[21], p[22], p[26]};
    for (int i = 0; i < 4; ++i)
        if (want[i] != r[3 * i] + 3 * r[3 * i + 1] + 9 * r[3 * i + 2]) { *ok = 0; return; }
    t3c_config h = *out; // fields the header does not carry (superframe_words) are kept
    h.profile = p[3] % 5;
    for (int g = 0; g < 3; ++g) { // LSB-first triples (bug B7)
        uint32_t v = p[4 + g];
        h.uep[3 * g] = v % 3; h.uep[3 * g + 1] = (v / 3) % 3; h.uep[3 * g + 2] = (v / 9) % 3;
    }
    h.tile_w = p[7]; h.tile_h = p[8];
    h.seed_a = p[9]; h.seed_b = p[10]; h.seed_s0 = p[11];
    const uint32_t sub = p[12] % 9, cen = (p[12] / 9) % 3;
    h.subword = sub == 1 ? 24 : sub == 2 ? 21 : sub == 3 ? 18 : sub == 4 ? 15 : 27;
    h.centered = cen != 0;
    h.coset = p[16] % 3;
    h.beacon_enabled = p[23] != 0; h.beacon_slot = p[24] % 9; h.beacon_period = p[25];
    *out = h;
    *ok = 1;
}

// ------------------------------------------------------------------------------------------
// General profile encoder (A.1-A.6): one CTA = TPB codewords of one band
// ------------------------------------------------------------------------------------------
// cwb = codewords per CTA (<= TPB): TPB for whole frames (thread per codeword), small for the ragged end the tiled kernels leave,
// where the symbol gather -- index maps with divisions -- is then spread over all threads
template <typename I>
__global__ void __launch_bounds__(TPB) k_encode_general(const uint8_t* __restrict__ raw, uint8_t* __restrict__ out, Geom g,
                                                        const GfTables* __restrict__ gf, const RsTables* __restrict__ rs, CwStart cs, uint32_t cwb)
{
    __shared__ uint64_t row[24 * kVals];
    __shared__ uint8_t stage[TPB * 26];
    const int b = blockIdx.y, k = g.k[b], r = 26 - k;
    const I c0 = (I)cs.c[b] + (I)blockIdx.x * cwb; // codewords before cs.c[b] of band b were coded by the tiled kernels
    if (c0 >= (I)g.ncw[b]) return;
    const RowTable& tab = rs->row[g.arith][(24 - k) / 2];
    for (int i = threadIdx.x; i < k * kVals; i += TPB) row[i] = tab.e[i / kVals][i % kVals];
    const uint32_t ncta = (uint32_t)(((I)g.ncw[b] - c0) < cwb ? ((I)g.ncw[b] - c0) : cwb);
    for (uint32_t idx = threadIdx.x; idx < ncta * (uint32_t)k; idx += TPB) {
        const uint32_t cl = idx / (uint32_t)k, i = idx - cl * (uint32_t)k;
        const I is = 9 * ((I)k * (c0 + cl) + i) + b;                             // band split, A.3
        const I j = perm2d<I>(is, (I)g.n_s, (I)g.tile_area, g.tile_w);           // 2D interleave, A.2
        stage[26 * cl + i] = (uint8_t)raw_symbol<I>(raw, (I)g.n_words, j);       // regroup, A.1
    }
    __syncthreads();
    if (threadIdx.x < ncta) {
        Planes acc{0, 0};
        uint8_t* my = stage + 26 * threadIdx.x;
        for (int i = 0; i < k; ++i) gf3_add(acc, row[i * kVals + my[i]]);
        const uint32_t lo = planes_to_sym4_lo(acc), hi = planes_to_sym4_hi(acc);
        for (int j = 0; j < r; ++j) my[k + j] = (uint8_t)((j < 4 ? lo >> (8 * j) : hi >> (8 * (j - 4))) & 0xFF);
    }
    __syncthreads();
    // scramble (A.4) and beacon-aware scatter (A.5): one division per thread, then the expanded index advances with the body index
    const I p0 = 26 * ((I)g.cw_base[b] + c0);
    for (uint32_t idx = threadIdx.x; idx < 26 * ncta; idx += TPB) {
        const I p = p0 + idx;
        out[52 + beacon_expand<I>(g, p)] = gf->scr[scr_state<I>(g, p)][stage[idx]];
    }
}
// header, beacon symbols and zero padding of one super-frame
__global__ void k_frame_misc(uint8_t* __restrict__ out_base, size_t stride_bytes, Geom g, const uint8_t* __restrict__ hdr52)
{
    uint8_t* __restrict__ out = out_base + stride_bytes * blockIdx.y;
    if (blockIdx.x == 0 && threadIdx.x < 52) out[threadIdx.x] = hdr52[threadIdx.x]; // the coded header depends on the config only (cached_header)
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t W = g.l_exp / 9;
    if (g.period && g.slot >= 0)
        for (uint64_t wd = tid * g.period; wd < W; wd += nth * g.period) out[52 + 9 * wd + g.slot] = g.bsym;
    // zeros after the last body symbol up to the end of the last word (never on a beacon position)
    const uint64_t q_end = g.l_body ? beacon_expand(g, g.l_body - 1) + 1 : 0;
    for (uint64_t q = q_end + tid; 52 + q < 9 * g.n_out; q += nth) {
        const bool is_beacon = g.period && g.slot >= 0 && q < g.l_exp && (q / 9) % g.period == 0 && (int)(q % 9) == g.slot;
        if (!is_beacon) out[52 + q] = 0;
    }
}

// The same for frames whose full super-tiles were coded by the super-tile kernels (k_super.cuh): those write the beacon slots that lie
// inside their runs, so only three kinds of slots are left: the one right before a run's first symbol (run boundaries), the ones inside the
// ragged end of every band (coded by k_encode_general, which writes body symbols only) and the ones after the last body symbol.
struct SparseMisc { uint32_t n_tiles; uint32_t ncw[9]; };
__global__ void k_frame_misc_sparse(uint8_t* __restrict__ out_base, size_t stride_bytes, Geom g, const uint8_t* __restrict__ hdr52, SparseMisc sm)
{
    uint8_t* __restrict__ out = out_base + stride_bytes * blockIdx.y;
    if (blockIdx.x == 0 && threadIdx.x < 52) out[threadIdx.x] = hdr52[threadIdx.x];
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t G = 9ull * g.period;
    auto is_slot = [&](uint64_t o) { return (o / 9) % g.period == 0 && (int)(o % 9) == g.slot; };
    // run boundaries: band b, super-tile T = 0 .. n_tiles (T = n_tiles: where the ragged end starts)
    for (uint64_t i = tid; i < 9ull * (sm.n_tiles + 1); i += nth) {
        const uint32_t b = (uint32_t)(i / (sm.n_tiles + 1)), T = (uint32_t)(i - (uint64_t)b * (sm.n_tiles + 1));
        const uint64_t p = 26 * (g.cw_base[b] + (uint64_t)sm.ncw[b] * T);
        if (p >= g.l_body) continue;
        const uint64_t o = beacon_expand<uint64_t>(g, p);
        if (o >= 1 && is_slot(o - 1)) out[52 + o - 1] = g.bsym;
    }
    // the ragged end of every band: slots inside [o(first symbol), o(last symbol)]
    if (tid < 9) {
        const uint32_t b = (uint32_t)tid;
        const uint64_t c0 = (uint64_t)sm.ncw[b] * sm.n_tiles;
        if (g.ncw[b] > c0) {
            const uint64_t o_a = beacon_expand<uint64_t>(g, 26 * (g.cw_base[b] + c0)), o_b = beacon_expand<uint64_t>(g, 26 * (g.cw_base[b] + g.ncw[b]) - 1);
            uint64_t s = o_a <= (uint64_t)g.slot ? (uint64_t)g.slot : ((o_a - g.slot + G - 1) / G) * G + g.slot; // first slot at or after o_a
            for (; s <= o_b; s += G) out[52 + s] = g.bsym;
        }
    }
    // after the last body symbol: slots of the remaining words, zeros elsewhere up to the end of the last word
    const uint64_t q_end = g.l_body ? beacon_expand<uint64_t>(g, g.l_body - 1) + 1 : 0;
    for (uint64_t q = q_end + tid; 52 + q < 9 * g.n_out; q += nth) out[52 + q] = (q < g.l_exp && is_slot(q)) ? g.bsym : 0;
}

// ------------------------------------------------------------------------------------------
// General consistent decoder (A.8): thread per codeword -> symbol stream sy' in scratch
// ------------------------------------------------------------------------------------------
// cwb = codewords per CTA (<= TPB): TPB for whole frames, small for the ragged end the tiled kernels leave, where fetching the 26 symbols of
// a codeword (beacon expansion and scrambler phase: divisions) is then spread over all threads
template <typename I>
__global__ void __launch_bounds__(TPB) k_decode_fixed_general(const uint8_t* __restrict__ in, uint8_t* __restrict__ sy, Geom g,
                                                              const GfTables* __restrict__ gf, uint32_t* status, CwStart cs, uint64_t pitch, uint32_t cwb)
{
    __shared__ GfTables sg;
    __shared__ uint8_t stage[TPB * 26];
    load_gf(sg, gf);
    const int b = blockIdx.y, k = g.k[b];
    const I c0 = (I)cs.c[b] + (I)blockIdx.x * cwb;
    if (c0 >= (I)g.ncw[b]) return;
    const uint32_t ncta = (uint32_t)(((I)g.ncw[b] - c0) < cwb ? ((I)g.ncw[b] - c0) : cwb);
    __syncthreads();
    const I p0 = 26 * ((I)g.cw_base[b] + c0);
    for (uint32_t idx = threadIdx.x; idx < 26 * ncta; idx += TPB) {
        const I p = p0 + idx;
        stage[idx] = sg.dsc[scr_state<I>(g, p)][in[52 + beacon_expand<I>(g, p)] % 27];
    }
    __syncthreads();
    if (threadIdx.x >= ncta) return;
    const I c = c0 + threadIdx.x;
    uint8_t cw[26], orig[26];
    for (int i = 0; i < 26; ++i) cw[i] = orig[i] = stage[26 * threadIdx.x + i];
    if (!rs_decode_thread(sg, cw, k, true, true)) { atomicExch(&status[0], 0u); return; }
    uint32_t nfix = 0;
    for (int i = 0; i < 26; ++i) nfix += cw[i] != orig[i];
    if (nfix) atomicAdd(&status[1], nfix);
    // scratch is band-major (band b at b*pitch): a thread's k symbols are contiguous, the regroup kernels gather through stream_trit
    for (int i = 0; i < k; ++i) sy[(I)b * (I)pitch + (I)k * c + i] = cw[i];
}
// symbols -> trits -> groups of 26 -> Word27 (OLD:1022-1039), optional de-interleave (involution)
template <typename I>
__device__ __forceinline__ uint32_t stream_trit(const uint8_t* __restrict__ sy, I n_sy, I area, uint32_t w, I ti, I pitch)
{
    const I j = ti / 3, jp = perm2d<I>(j, n_sy, area, w);
    const I q = jp / 9;
    const uint32_t s = sy[pitch ? (jp - 9 * q) * pitch + q : jp], c = (uint32_t)(ti - 3 * j); // pitch != 0: band-major scratch
    return c == 0 ? s % 3 : (c == 1 ? (s / 3) % 3 : (s / 9) % 3);
}
template <typename I>
__global__ void k_regroup_words(const uint8_t* __restrict__ sy, I n_sy, I area, uint32_t tw, uint8_t* __restrict__ out, I n_words, I w_start, I pitch)
{
    const I w = w_start + (I)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    for (int s = 0; s < 9; ++s) {
        uint32_t v = 0, mul = 1;
        for (int c = 0; c < 3; ++c, mul *= 3) {
            const int t = 3 * s + c;
            if (t < 26) v += mul * stream_trit<I>(sy, n_sy, area, tw, 26 * w + t, pitch);
        }
        out[9 * w + s] = (uint8_t)v;
    }
}
template <typename I>
__global__ void k_regroup_rgb(const uint8_t* __restrict__ sy, I n_sy, I area, uint32_t tw, uint8_t* __restrict__ rgb, I n_px, I pitch, I p_start)
{
    const I p = p_start + (I)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_px) return;
    uint32_t t[13];
    for (int i = 0; i < 13; ++i) t[i] = stream_trit<I>(sy, n_sy, area, tw, 13 * p + i, pitch);
    const int Yq = t[0] + 3 * t[1] + 9 * t[2] + 27 * t[3] + 81 * t[4];
    const int Cb = (int)(t[5] + 3 * t[6] + 9 * t[7] + 27 * t[8]) - 40, Cr = (int)(t[9] + 3 * t[10] + 9 * t[11] + 27 * t[12]) - 40;
    int R, G, B;
    ycbcr8_to_rgb(dequant_y(Yq), dequant_c(Cb), dequant_c(Cr), R, G, B);
    rgb[3 * p] = (uint8_t)R; rgb[3 * p + 1] = (uint8_t)G; rgb[3 * p + 2] = (uint8_t)B;
}

// ------------------------------------------------------------------------------------------
// Reference decoder as shipped (A.7): slot-major demap of words 6.., descramble by absolute
// body-symbol position, band-major `use`
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_decode_ref_general(const uint8_t* __restrict__ in, uint8_t* __restrict__ use, RefDecGeom g,
                                                            const GfTables* __restrict__ gf, uint32_t* status)
{
    __shared__ GfTables sg;
    load_gf(sg, gf);
    __syncthreads();
    const int b = blockIdx.y, k = g.k[b];
    const uint64_t c = (uint64_t)blockIdx.x * TPB + threadIdx.x;
    if (c >= g.ncw[b]) return;
    uint8_t cw[26];
    const bool skipping = g.period && b == g.slot;
    for (int i = 0; i < 26; ++i) {
        const uint64_t n = 26 * c + i;                                            // n-th kept word of this band
        const uint64_t wi = skipping ? n + n / (g.period - 1) + 1 : n;            // words with wi%P==0 are skipped (OLD:957)
        const uint64_t pos = 9 * wi + b;                                          // descrambler runs over all 9 slots (OLD:938-947)
        const uint32_t st = pos < 2 ? g.st[pos] : g.st[2 + (uint32_t)((pos - 2) % 6)];
        cw[i] = sg.dsc[st][in[54 + pos] % 27];
    }
    if (!rs_decode_thread(sg, cw, k, false)) { atomicExch(&status[0], 0u); return; }
    for (int i = 0; i < k; ++i) use[g.use_base[b] + (uint64_t)k * c + i] = cw[i];
}

} // namespace

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
int launch_init_status(uint32_t* d_status, size_t n_frames, cudaStream_t st)
{
    if (!n_frames) return 0;
    k_init_status<<<blocks_for(2 * n_frames, 256), 256, 0, st>>>(d_status, n_frames);
    return 1;
}
int launch_rgb_to_quant(const uint8_t* rgb, size_t n, t3c_pixel* out, cudaStream_t st)
{
    if (!n) return 0;
    const size_t done = launch_rgb_to_quant8(rgb, n, out, st);
    if (done < n) k_rgb_to_quant<<<blocks_for(n - done, 256), 256, 0, st>>>(rgb + 3 * done, n - done, reinterpret_cast<uint16_t*>(out + done));
    return (done ? 1 : 0) + (done < n ? 1 : 0);
}
int launch_quant_to_rgb(const t3c_pixel* px, size_t n, uint8_t* rgb, cudaStream_t st)
{
    if (!n) return 0;
    const size_t done = launch_quant_to_rgb8(px, n, rgb, st);
    if (done < n) k_quant_to_rgb<<<blocks_for(n - done, 256), 256, 0, st>>>(reinterpret_cast<const uint16_t*>(px + done), n - done, rgb + 3 * done);
    return (done ? 1 : 0) + (done < n ? 1 : 0);
}
static int raw2_grid(uint32_t n_tiles)
{
    static int sms[64] = {}; // per device: SM count, and the opt-in to > 48 KB of dynamic shared memory
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!sms[dev]) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(raw2::k_pack_pixels_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, raw2::SMEM);
        cudaFuncSetAttribute(raw2::k_unpack_pixels_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, raw2::SMEM);
        sms[dev] = n;
    }
    const uint32_t need = (n_tiles + raw2::WARPS - 1) / raw2::WARPS;
    return (int)(need < (uint32_t)sms[dev] ? need : (uint32_t)sms[dev]);
}
int launch_pack_pixels(const t3c_pixel* px, size_t n, uint8_t* words, cudaStream_t st)
{
    if (!n) return 0;
    int k = 0;
    size_t done = 0;
    if (n >= 64 * raw2::TILE_PX && !(((uintptr_t)px | (uintptr_t)words) & 15)) { // whole 512-pixel tiles: bulk-async persistent kernel
        const uint32_t n_tiles = (uint32_t)(n / raw2::TILE_PX);
        raw2::k_pack_pixels_v2<<<raw2_grid(n_tiles), raw2::TPB, raw2::SMEM, st>>>(reinterpret_cast<const uint8_t*>(px), n_tiles, words);
        done = (size_t)n_tiles * raw2::TILE_PX;
        ++k;
    }
    if (done < n) {
        k_pack_pixels<<<blocks_for(n - done, PACK_PX_PER_BLOCK), PACK_TPB, 0, st>>>(reinterpret_cast<const uint16_t*>(px) + 3 * done, n - done, words + 9 * (done / 2));
        ++k;
    }
    return k;
}
int launch_unpack_pixels(const uint8_t* words, size_t n, t3c_pixel* px, cudaStream_t st)
{
    if (!n) return 0;
    int k = 0;
    size_t done = 0; // words
    if (n >= 32 * raw2::TILE_PX && !(((uintptr_t)px | (uintptr_t)words) & 15)) {
        const uint32_t n_tiles = (uint32_t)(n / (raw2::TILE_PX / 2));
        raw2::k_unpack_pixels_v2<<<raw2_grid(n_tiles), raw2::TPB, raw2::SMEM, st>>>(words, n_tiles, reinterpret_cast<uint8_t*>(px));
        done = (size_t)n_tiles * (raw2::TILE_PX / 2);
        ++k;
    }
    if (done < n) {
        k_unpack_pixels<<<blocks_for(n - done, PACK_PX_PER_BLOCK / 2), PACK_TPB, 0, st>>>(words + 9 * done, n - done, reinterpret_cast<uint16_t*>(px) + 6 * done);
        ++k;
    }
    return k;
}
int launch_mod27(const uint8_t* in, size_t n, uint8_t* out, cudaStream_t st)
{
    if (!n) return 0;
    k_mod27<<<blocks_for(n, 256), 256, 0, st>>>(in, n, out);
    return 1;
}
int launch_rs_encode_blocks(const DevTables& T, int k, int arith, const uint8_t* data, size_t n, uint8_t* out, cudaStream_t st)
{
    if (!n) return 0;
    k_rs_encode_blocks<<<blocks_for(n, TPB), TPB, 0, st>>>(&T.rs->row[arith ? 1 : 0][kidx_of(k)], k, data, n, out);
    return 1;
}
int launch_rs_decode_blocks(const DevTables& T, int k, int arith, uint8_t* inout, size_t n, uint8_t* out_k, uint8_t* ok, cudaStream_t st)
{
    if (!n) return 0;
    k_rs_decode_blocks<<<blocks_for(n, TPB), TPB, 0, st>>>(T.gf, k, arith, inout, n, out_k, ok);
    return 1;
}
int launch_perm2d(const uint8_t* in, uint8_t* out, size_t n, uint32_t w, uint32_t h, cudaStream_t st)
{
    if (!n) return 0;
    if (n < (1ull << 31)) k_perm2d<uint32_t><<<blocks_for(n, 256), 256, 0, st>>>(in, out, (uint32_t)n, (uint32_t)w * h, w);
    else k_perm2d<uint64_t><<<blocks_for(n, 256), 256, 0, st>>>(in, out, (uint64_t)n, (uint64_t)w * h, w);
    return 1;
}
int launch_header_emit(const DevTables& T, const t3c_config& cfg, int arith, uint8_t* hdr27, uint8_t* coded52, cudaStream_t st)
{
    k_header_emit<<<1, 32, 0, st>>>(cfg, arith ? 1 : 0, T.gf, T.rs, hdr27, coded52);
    return 1;
}
int launch_header_pack(const t3c_config& cfg, uint32_t magic, uint32_t version, uint32_t hash, uint32_t seq, uint8_t* d_hdr27, cudaStream_t st)
{
    k_header_pack<<<1, 1, 0, st>>>(cfg, magic, version, hash, seq, d_hdr27);
    return 1;
}
int launch_header_check_unpack(const uint8_t* d_sym27, t3c_config* d_cfg, uint32_t* d_out4, int* d_ok, cudaStream_t st)
{
    k_header_check_unpack<<<1, 1, 0, st>>>(d_sym27, d_cfg, d_out4, d_ok);
    return 1;
}
int launch_crc3_rem12(const uint8_t* d_trits, size_t n, uint8_t* d_out12, cudaStream_t st)
{
    k_crc3_rem12<<<1, 1, 0, st>>>(d_trits, n, d_out12);
    return 1;
}
int launch_scramble(const DevTables& T, uint8_t* d_syms, size_t n, const uint8_t st8[8], int inverse, cudaStream_t st)
{
    if (!n) return 0;
    ScrStates S;
    for (int i = 0; i < 8; ++i) S.st[i] = st8[i];
    k_scramble<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_syms, n, S, T.gf, inverse);
    return 1;
}
int launch_header_parse(const DevTables& T, int arith, const uint8_t* words, size_t n_words, t3c_config* d_cfg, int* d_ok, cudaStream_t st)
{
    k_header_parse<<<1, 64, 0, st>>>(T.gf, arith, words, n_words, d_cfg, d_ok);
    return 1;
}
// most codewords any band still has to code after its start
static uint64_t cw_left(const Geom& g, const CwStart& cs)
{
    uint64_t mx = 0;
    for (int b = 0; b < 9; ++b) if (g.ncw[b] > cs.c[b] && g.ncw[b] - cs.c[b] > mx) mx = g.ncw[b] - cs.c[b];
    return mx;
}
int launch_encode_general(const DevTables& T, const t3c_config& cfg, const Geom& g, const uint8_t* raw, uint8_t* out, cudaStream_t st, uint64_t cw_start)
{
    CwStart cs;
    for (int b = 0; b < 9; ++b) cs.c[b] = cw_start;
    return launch_encode_general_from(T, cfg, g, raw, out, st, cs);
}
int launch_encode_general_from(const DevTables& T, const t3c_config& cfg, const Geom& g, const uint8_t* raw, uint8_t* out, cudaStream_t st, const CwStart& cs, bool finish)
{
    int n = 0;
    const uint64_t mx = cw_left(g, cs);
    if (mx) {
        const uint32_t cwb = mx >= 4096 ? (uint32_t)TPB : 8u; // the ragged end of a frame: few codewords, many threads per codeword
        if (small_geom(g)) k_encode_general<uint32_t><<<dim3(blocks_for(mx, cwb), 9), TPB, 0, st>>>(raw, out, g, T.gf, T.rs, cs, cwb);
        else k_encode_general<uint64_t><<<dim3(blocks_for(mx, cwb), 9), TPB, 0, st>>>(raw, out, g, T.gf, T.rs, cs, cwb);
        ++n;
    }
    if (!finish) return n;
    // without a beacon the rest of the frame is the (cached) coded header and the zero padding
    return n + (use_beacon(cfg) ? launch_frame_misc(T, cfg, g, out, 1, 0, st) : launch_frame_finish(T, cfg, g, out, 1, 0, st));
}
const uint8_t* cached_header(const DevTables& T, const t3c_config& cfg, int arith, cudaStream_t st, int& launches)
{
    HeaderCache& H = *T.hdr;
    const int a = arith ? 1 : 0;
    for (int i = 0; i < HeaderCache::N; ++i)
        if (H.e[i].valid && H.e[i].arith == a && std::memcmp(&H.e[i].cfg, &cfg, sizeof cfg) == 0) return H.e[i].d52;
    int slot = -1;
    for (int i = 0; i < HeaderCache::N; ++i) if (!H.e[i].valid) { slot = i; break; }
    if (slot < 0) {   // all entries taken: nothing in flight may still read the one that is reused
        cudaDeviceSynchronize();
        slot = H.next;
        H.next = (H.next + 1) % HeaderCache::N;
    }
    HeaderCache::Entry& E = H.e[slot];
    E.valid = false;
    launches += launch_header_emit(T, cfg, arith, E.d27, E.d52, st);
    cudaStreamSynchronize(st); // once per new config: later calls may come on other streams
    E.cfg = cfg; E.arith = a; E.valid = true;
    return E.d52;
}
__global__ void k_frame_finish(uint8_t* __restrict__ out_base, size_t stride_bytes, const uint8_t* __restrict__ hdr52, uint64_t body_end, uint64_t frame_bytes)
{
    uint8_t* __restrict__ out = out_base + stride_bytes * blockIdx.x;
    const uint32_t t = threadIdx.x;
    if (t < 52) out[t] = hdr52[t];
    if (body_end + t < frame_bytes && t < 16) out[body_end + t] = 0; // zeros up to the end of the last word
}
int launch_frame_finish(const DevTables& T, const t3c_config& cfg, const Geom& g, uint8_t* out, size_t n_frames, size_t stride_bytes, cudaStream_t st)
{
    if (!n_frames) return 0;
    int n = 0;
    const uint8_t* hdr = cached_header(T, cfg, g.arith, st, n);
    k_frame_finish<<<(unsigned)n_frames, 64, 0, st>>>(out, stride_bytes, hdr, 52 + g.l_body, 9 * g.n_out);
    return n + 1;
}
int launch_frame_misc(const DevTables& T, const t3c_config& cfg, const Geom& g, uint8_t* out, size_t n_frames, size_t stride_bytes, cudaStream_t st)
{
    if (!n_frames) return 0;
    const uint64_t nb = g.period ? g.l_exp / 9 / g.period + 1 : 1;
    unsigned blocks = (unsigned)((nb + 255) / 256);
    if (blocks > 1024) blocks = 1024;
    int n = 0;
    const uint8_t* hdr = cached_header(T, cfg, g.arith, st, n);
    k_frame_misc<<<dim3(blocks, (unsigned)n_frames), 256, 0, st>>>(out, stride_bytes, g, hdr);
    return n + 1;
}
// header, padding and the beacon slots the super-tile kernels and the general tail encoder leave (see k_frame_misc_sparse)
int launch_frame_misc_sparse(const DevTables& T, const t3c_config& cfg, const Geom& g, uint8_t* out, size_t n_frames, size_t stride_bytes, cudaStream_t st,
                             uint32_t n_tiles, const uint32_t ncw_tile[9])
{
    if (!n_frames) return 0;
    if (!(g.period && g.slot >= 0)) return use_beacon(cfg) ? launch_frame_misc(T, cfg, g, out, n_frames, stride_bytes, st) : launch_frame_finish(T, cfg, g, out, n_frames, stride_bytes, st);
    int n = 0;
    const uint8_t* hdr = cached_header(T, cfg, g.arith, st, n);
    SparseMisc sm;
    sm.n_tiles = n_tiles;
    for (int b = 0; b < 9; ++b) sm.ncw[b] = ncw_tile[b];
    const uint64_t items = 9ull * (n_tiles + 1);
    unsigned blocks = (unsigned)((items + 255) / 256);
    if (blocks > 256) blocks = 256;
    k_frame_misc_sparse<<<dim3(blocks ? blocks : 1, (unsigned)n_frames), 256, 0, st>>>(out, stride_bytes, g, hdr, sm);
    return n + 1;
}
int launch_decode_fixed_general(const DevTables& T, const Geom& g, const uint8_t* in, uint8_t* sy, uint64_t pitch, uint32_t* status, cudaStream_t st, uint64_t cw_start)
{
    CwStart cs;
    for (int b = 0; b < 9; ++b) cs.c[b] = cw_start;
    return launch_decode_fixed_general_from(T, g, in, sy, pitch, status, st, cs);
}
int launch_decode_fixed_general_from(const DevTables& T, const Geom& g, const uint8_t* in, uint8_t* sy, uint64_t pitch, uint32_t* status, cudaStream_t st, const CwStart& cs)
{
    const uint64_t mx = cw_left(g, cs);
    if (!mx) return 0;
    const uint32_t cwb = mx >= 4096 ? (uint32_t)TPB : 8u; // the ragged end of a frame: few codewords, many threads per codeword
    if (small_geom(g)) k_decode_fixed_general<uint32_t><<<dim3(blocks_for(mx, cwb), 9), TPB, 0, st>>>(in, sy, g, T.gf, status, cs, pitch, cwb);
    else k_decode_fixed_general<uint64_t><<<dim3(blocks_for(mx, cwb), 9), TPB, 0, st>>>(in, sy, g, T.gf, status, cs, pitch, cwb);
    return 1;
}
int launch_regroup_words(const uint8_t* sy, uint64_t n_sy, uint64_t area, uint32_t tw, uint8_t* out, size_t n_words, cudaStream_t st, size_t w_start, uint64_t pitch)
{
    if (n_words <= w_start) return 0;
    if (27 * (uint64_t)n_words + 64 < (1ull << 31) && 3 * n_sy + 64 < (1ull << 31))
        k_regroup_words<uint32_t><<<blocks_for(n_words - w_start, 256), 256, 0, st>>>(sy, (uint32_t)n_sy, (uint32_t)area, tw, out, (uint32_t)n_words, (uint32_t)w_start, (uint32_t)pitch);
    else k_regroup_words<uint64_t><<<blocks_for(n_words - w_start, 256), 256, 0, st>>>(sy, n_sy, area, tw, out, (uint64_t)n_words, (uint64_t)w_start, pitch);
    return 1;
}
int launch_regroup_rgb(const uint8_t* sy, uint64_t n_sy, uint64_t area, uint32_t tw, uint8_t* rgb, size_t n_px, cudaStream_t st, uint64_t pitch, size_t p_start)
{
    if (n_px <= p_start) return 0;
    if (13 * (uint64_t)n_px + 64 < (1ull << 31) && 3 * n_sy + 64 < (1ull << 31))
        k_regroup_rgb<uint32_t><<<blocks_for(n_px - p_start, 256), 256, 0, st>>>(sy, (uint32_t)n_sy, (uint32_t)area, tw, rgb, (uint32_t)n_px, (uint32_t)pitch, (uint32_t)p_start);
    else k_regroup_rgb<uint64_t><<<blocks_for(n_px - p_start, 256), 256, 0, st>>>(sy, n_sy, area, tw, rgb, (uint64_t)n_px, pitch, (uint64_t)p_start);
    return 1;
}
int launch_decode_ref_general(const DevTables& T, const RefDecGeom& g, const uint8_t* in, uint8_t* use, uint32_t* status, cudaStream_t st)
{
    uint64_t mx = 0;
    for (int b = 0; b < 9; ++b) mx = g.ncw[b] > mx ? g.ncw[b] : mx;
    if (!mx) return 0;
    k_decode_ref_general<<<dim3(blocks_for(mx, TPB), 9), TPB, 0, st>>>(in, use, g, T.gf, status);
    return 1;
}

} // namespace t3c
