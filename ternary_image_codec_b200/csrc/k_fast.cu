// k_fast.cu -- tiled, fused kernels for the headline family: uniform k, 1D 9-band interleave, no beacon.
//
//   encode: RGB8 --(bridge, 13 trits/px)--> regrouped symbol stream --(9-band transpose in shared memory)-->
//           RS(26,k) parity (bit-plane row tables) --> scramble --> 9 band-major runs of the body
//   decode: 9 runs --> syndrome screen (same row tables) --> [dirty codewords: BM/Chien/Forney] --> descramble
//           --> 9-band transpose back --> 13 trits/px --> dequant --> RGB8
//
// One CTA works on a tile of C=52 codewords per band (9*k*52 stream symbols = 108*k pixels): every global
// access is a 128-bit coalesced transfer of a contiguous run (one RGB run, nine body runs), the band
// transpose and all table look-ups stay in shared memory, and the grid is persistent (a multiple of the SM
// count) so the row table is staged once per CTA.  HBM traffic is exactly the algorithmic 3 B/px + 9 B/word.
#include "dev.cuh"
#include "launch.h"

namespace t3c {
namespace {

constexpr int C_TILE = 52;      // codewords per band per tile (multiple of 26: tile = whole 12-pixel units)
constexpr int FAST_TPB = 256;
constexpr uint32_t M27 = 159072863u; // ceil(2^32/27): exact quotient for x < 2^26

struct FastParams {
    const uint8_t* in;      // encode: rgb frames   decode: profile words
    uint8_t* out;           // encode: profile words decode: rgb frames
    uint64_t in_stride;     // bytes between frames
    uint64_t out_stride;
    uint64_t n_px;          // pixels per frame
    uint64_t px_out;        // decode: pixels to write per frame
    uint32_t n_tiles;       // tiles per frame
    uint32_t n_frames;
    uint32_t* status;       // decode: {ok, n_corrected} per frame
    uint32_t chk_nz[7], chk_two[7]; // decode: sum of T_i[13*st_i] per scrambler phase (6) and for p0==0
};

template <int K> struct Cfg {
    static constexpr int R = 26 - K;
    static constexpr int UNITS = 9 * K;                 // 12-pixel units per tile
    static constexpr int PX = 12 * UNITS;               // pixels per tile
    static constexpr int RGB_BYTES = 3 * PX;
    static constexpr int SYM = 52 * UNITS;              // stream symbols per tile = 9*K*C_TILE
    static constexpr int RUN = 26 * C_TILE;             // bytes per band run
    static constexpr int RUN_PITCH = RUN + 24;          // + alignment slack, multiple of 8
    static constexpr int NCW = 9 * C_TILE;
    // shared memory carve-up (bytes)
    static constexpr int OFF_TAB = 0;
    static constexpr int TAB_BYTES = 26 * kVals * 8;
    static constexpr int OFF_RGB = OFF_TAB + TAB_BYTES;
    static constexpr int RGB_PITCH = ((RGB_BYTES + 15) / 16) * 16 + 32;
    static constexpr int OFF_S = OFF_RGB + RGB_PITCH;
    static constexpr int S_PITCH = ((SYM + 15) / 16) * 16;
    static constexpr int OFF_O = OFF_S + S_PITCH;
    static constexpr int OFF_GF = OFF_O + 9 * RUN_PITCH;
    static constexpr int TOTAL = OFF_GF + ((int)sizeof(GfTables) + 15) / 16 * 16 + 64;
};

// 13 base-27 digits of three pixel values (39 trits): packed 4+4+4 symbols and the 13th
__device__ __forceinline__ uint32_t digits4(uint32_t x, uint32_t& q4)
{
    const uint32_t q1 = __umulhi(x, M27), q2 = __umulhi(q1, M27), q3 = __umulhi(q2, M27);
    q4 = __umulhi(q3, M27);
    return x + 229u * q1 + 58624u * q2 + 15007744u * q3 - 452984832u * q4; // s0 | s1<<8 | s2<<16 | s3<<24
}
__device__ __forceinline__ void triple_to_symbols(uint32_t A0, uint32_t A1, uint32_t A2, uint32_t& w0, uint32_t& w1, uint32_t& w2, uint32_t& s12)
{
    uint32_t t, u;
    w0 = digits4(A0, t);            // trits 0..11, t = trit 12
    w1 = digits4(t + 3u * A1, u);   // trits 12..23, u = trits 24,25
    w2 = digits4(u + 9u * A2, s12); // trits 24..35, s12 = trits 36..38
}
// inverse: 13 symbols -> three 13-trit pixel values
__device__ __forceinline__ void symbols_to_triple(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t s12, uint32_t& A0, uint32_t& A1, uint32_t& A2)
{
    // b0 + 27 b1 + 729 (b2 + 27 b3): two 4-way byte dot products and one multiply-add
    auto val4 = [](uint32_t w) { return __dp4a(w, 0x00001B01u, 0u) + 729u * __dp4a(w, 0x1B010000u, 0u); };
    const uint32_t v0 = val4(w0), v1 = val4(w1), v2 = val4(w2) + 531441u * s12; // 12, 12 and 15 trits
    const uint32_t t = v1 % 3u;                         // trit 12 belongs to pixel 0
    A0 = v0 + 531441u * t;
    const uint32_t u = v2 % 9u;                         // trits 24,25 belong to pixel 1
    A1 = v1 / 3u + 177147u * u;
    A2 = v2 / 9u;
}

// byte j of w as float, exact (magic-number conversion: full-rate PRMT + FADD instead of I2F)
__device__ __forceinline__ float byte_to_float(uint32_t w, int j)
{
    return __fadd_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u | (uint32_t)j)), -8388608.0f);
}
// rgb_to_ycbcr + quantize_ycbcr (IMG:47-56,69-78) -> 13-trit pixel value; no clamps are needed for 8-bit
// inputs: y in [0,255.0001), cb,cr in [0.5,255.5] and round(255.5)=256 quantises like 255.
// 0.5f*x is exact, so fma(0.5,x,t) == fl(t + fl(0.5*x)): two multiplies are folded without changing a bit.
// Rounding: FADD.RM against 2^22+0.5 leaves B = 0x25400000 + floor(v+0.5) after >>1 (see round_pos); the
// offset is folded into the quantiser constants (0x25400000 has its low 7 bits clear).
__device__ __forceinline__ uint32_t rgb_to_value(float r, float g, float b)
{
    const float y = __fadd_rn(__fadd_rn(__fmul_rn(0.299f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b));
    const float cb = __fadd_rn(__fmaf_rn(0.5f, b, __fsub_rn(__fmul_rn(-0.168736f, r), __fmul_rn(0.331264f, g))), 128.0f);
    const float cr = __fadd_rn(__fsub_rn(__fmaf_rn(0.5f, r, -__fmul_rn(0.418688f, g)), __fmul_rn(0.081312f, b)), 128.0f);
    constexpr uint32_t KB = 0x25400000u;
    const uint32_t By = (uint32_t)__float_as_int(__fadd_rd(y, 4194304.5f)) >> 1;
    const uint32_t Bb = (uint32_t)__float_as_int(__fadd_rd(cb, 4194304.5f)) >> 1;
    const uint32_t Br = (uint32_t)__float_as_int(__fadd_rd(cr, 4194304.5f)) >> 1;
    const uint32_t yq = (By * 484u + (255u - KB * 484u)) / 510u;                                   // quant_y
    const uint32_t ub = (5u * Bb + (Bb >> 7) + (7u - 5u * KB - (KB >> 7))) >> 4;                   // quant_c_off
    const uint32_t ur = (5u * Br + (Br >> 7) + (7u - 5u * KB - (KB >> 7))) >> 4;
    return yq + 243u * ub + 19683u * ur;
}
// pixel value -> RGB8 bytes (decode_raw_words_to_pixels + quant_stream_to_rgb, OLD:706-722, IMG:57-84)
__device__ __forceinline__ uint32_t value_to_rgb(uint32_t A)
{
    const uint32_t q = A / 243u, Yq = A - 243u * q, ur = q / 81u, ub = q - 81u * ur;
    int R, G, B;
    ycbcr8_to_rgb(dequant_y((int)Yq), dequant_c((int)ub - 40), dequant_c((int)ur - 40), R, G, B);
    return (uint32_t)R | ((uint32_t)G << 8) | ((uint32_t)B << 16);
}

// coalesced copy of a contiguous global byte range into shared memory; smem byte i <-> global byte
// (g_lo - pad + i) with pad = g_lo % 16, so 16-byte global chunks stay 16-byte aligned in shared memory.
__device__ __forceinline__ void load_run(uint8_t* s, const uint8_t* __restrict__ gbase, uint64_t g_lo, uint64_t g_hi, uint64_t g_limit)
{
    const uint64_t a0 = g_lo & ~15ull;
    const uint32_t nchunk = (uint32_t)((g_hi - a0 + 15) >> 4);
    for (uint32_t c = threadIdx.x; c < nchunk; c += FAST_TPB) {
        const uint64_t ga = a0 + 16ull * c;
        if (ga + 16 <= g_limit) {
            *reinterpret_cast<uint4*>(s + 16 * c) = __ldg(reinterpret_cast<const uint4*>(gbase + ga));
        } else {
            for (int i = 0; i < 16; ++i) s[16 * c + i] = ga + i < g_limit ? gbase[ga + i] : 0;
        }
    }
}
// coalesced copy shared -> global of bytes [g_lo, g_hi); same alignment convention as load_run
__device__ __forceinline__ void store_run(const uint8_t* s, uint8_t* __restrict__ gbase, uint64_t g_lo, uint64_t g_hi)
{
    if (g_hi <= g_lo) return;
    const uint64_t a0 = g_lo & ~15ull;
    const uint32_t nchunk = (uint32_t)((g_hi - a0 + 15) >> 4);
    for (uint32_t c = threadIdx.x; c < nchunk; c += FAST_TPB) {
        const uint64_t ga = a0 + 16ull * c;
        if (ga >= g_lo && ga + 16 <= g_hi) {
            *reinterpret_cast<uint4*>(gbase + ga) = *reinterpret_cast<const uint4*>(s + 16 * c);
        } else {
            for (int i = 0; i < 16; ++i) if (ga + i >= g_lo && ga + i < g_hi) gbase[ga + i] = s[16 * c + i];
        }
    }
}

// the nine band runs of a tile, shared -> global, flattened over (run, 16-byte chunk); 32-bit index math,
// 128-bit stores for interior chunks, 32-bit (or byte) stores only in the two boundary chunks of a run
template <int PITCH>
__device__ __forceinline__ void store_runs9(const uint8_t* O, uint8_t* __restrict__ gbase, const uint64_t* run_lo, const uint32_t* run_n)
{
    constexpr int CH = PITCH / 16;
    for (int j = threadIdx.x; j < 9 * CH; j += FAST_TPB) {
        const int b = j / CH, c = j - b * CH;
        const int len = 26 * (int)run_n[b];
        const uint64_t lo = run_lo[b];
        const int pad = (int)(lo & 15), r0 = 16 * c - pad, r1 = r0 + 16;
        if (r1 <= 0 || r0 >= len) continue;
        uint8_t* gp = gbase + (lo - pad) + 16 * c;
        const uint8_t* sp = O + PITCH * b + 16 * c;
        if (r0 >= 0 && r1 <= len) {
            *reinterpret_cast<uint4*>(gp) = *reinterpret_cast<const uint4*>(sp);
        } else {
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int a = r0 + 4 * w;
                if (a >= 0 && a + 4 <= len) *reinterpret_cast<uint32_t*>(gp + 4 * w) = *reinterpret_cast<const uint32_t*>(sp + 4 * w);
                else
                    for (int i = 0; i < 4; ++i) if (a + i >= 0 && a + i < len) gp[4 * w + i] = sp[4 * w + i];
            }
        }
    }
}
template <int PITCH>
__device__ __forceinline__ void load_runs9(uint8_t* O, const uint8_t* __restrict__ gbase, const uint64_t* run_lo, const uint32_t* run_n, uint64_t g_limit)
{
    constexpr int CH = PITCH / 16;
    for (int j = threadIdx.x; j < 9 * CH; j += FAST_TPB) {
        const int b = j / CH, c = j - b * CH;
        const int len = 26 * (int)run_n[b];
        const uint64_t lo = run_lo[b];
        const int pad = (int)(lo & 15), r0 = 16 * c - pad;
        if (r0 + 16 <= 0 || r0 >= len) continue;
        const uint64_t ga = (lo - pad) + 16 * c;
        uint8_t* sp = O + PITCH * b + 16 * c;
        if (ga + 16 <= g_limit) *reinterpret_cast<uint4*>(sp) = __ldg(reinterpret_cast<const uint4*>(gbase + ga));
        else
            for (int i = 0; i < 16; ++i) sp[i] = ga + i < g_limit ? gbase[ga + i] : 0;
    }
}

// body runs start at even byte offsets (the launchers only take the fast path when every frame does)
__device__ __forceinline__ void store2(uint8_t* p, uint32_t lo, uint32_t hi) { *reinterpret_cast<uint16_t*>(p) = (uint16_t)(lo | (hi << 8)); }
__device__ __forceinline__ uint32_t load2(const uint8_t* p) { return *reinterpret_cast<const uint16_t*>(p); }

// =============================================================================================
// encode
// =============================================================================================
template <int K>
__global__ void __launch_bounds__(FAST_TPB, 4) k_encode_rgb_fast(FastParams P, Geom g, const GfTables* __restrict__ gf, const RsTables* __restrict__ rs)
{
    using L = Cfg<K>;
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t* tab = reinterpret_cast<uint64_t*>(smem + L::OFF_TAB);
    uint8_t* s_rgb = smem + L::OFF_RGB;
    uint8_t* S = smem + L::OFF_S;
    uint8_t* O = smem + L::OFF_O;
    uint8_t* scr = smem + L::OFF_GF; // 3 x 32 scramble look-up
    __shared__ uint64_t run_lo[9];
    __shared__ uint32_t run_n[9], run_ph[9];
    const int tid = threadIdx.x;
    __shared__ uint32_t stoff[14]; // 32*st for phase index 0..11 (two periods) and for body indices 0,1
    {
        const RowTable& T = rs->row[g.arith][(24 - K) / 2];
        for (int i = tid; i < K * kVals; i += FAST_TPB) tab[i] = T.e[i / kVals][i % kVals];
        if (tid < 96) scr[tid] = gf->scr[tid / 32][tid % 32];
        if (tid < 14) stoff[tid] = 32u * (tid < 12 ? g.st[2 + tid % 6] : g.st[tid - 12]);
    }
    const uint64_t total = (uint64_t)P.n_tiles * P.n_frames;
    for (uint64_t tile_id = blockIdx.x; tile_id < total; tile_id += gridDim.x) {
        const uint32_t f = (uint32_t)(tile_id / P.n_tiles), tile = (uint32_t)(tile_id - (uint64_t)f * P.n_tiles);
        const uint8_t* rgb = P.in + P.in_stride * f;
        uint8_t* out = P.out + P.out_stride * f;
        // ---- phase 0: the tile's RGB run -> shared
        const uint64_t px0 = (uint64_t)L::PX * tile;
        const uint64_t g_lo = (uint64_t)(rgb - P.in) + 3 * px0;
        uint64_t g_hi = (uint64_t)(rgb - P.in) + 3 * (px0 + L::PX < P.n_px ? px0 + L::PX : P.n_px);
        if (g_hi < g_lo) g_hi = g_lo;
        const uint32_t pad = (uint32_t)(g_lo & 15);
        __syncthreads(); // previous tile's phase C is done with O; tab/scr visible
        if (tid < 9) { // the nine output runs of this tile, as byte offsets from P.out (A.6: 52 + 26*(cw_base_b + c))
            const uint64_t c0 = (uint64_t)C_TILE * tile;
            const uint64_t n = c0 >= g.ncw[tid] ? 0 : ((g.ncw[tid] - c0) < C_TILE ? (g.ncw[tid] - c0) : C_TILE);
            run_lo[tid] = (uint64_t)(out - P.out) + 52 + 26 * (g.cw_base[tid] + c0);
            run_n[tid] = (uint32_t)n;
            run_ph[tid] = (uint32_t)((26 * (g.cw_base[tid] + c0) + 4) % 6) | (g.cw_base[tid] + c0 == 0 ? 8u : 0u);
        }
        load_run(s_rgb, P.in, g_lo, g_hi, P.in_stride * P.n_frames);
        __syncthreads();
        // ---- phase A: 12 pixels -> 52 stream symbols per thread (A.1 regroup fused with the bridge)
        for (int u = tid; u < L::UNITS; u += FAST_TPB) {
            const uint8_t* me = s_rgb + pad + 36 * u;
            const uint32_t sh = (uint32_t)((uintptr_t)me & 3) * 8;
            const uint32_t* mw = reinterpret_cast<const uint32_t*>((uintptr_t)me & ~(uintptr_t)3);
            uint32_t w[10];
#pragma unroll
            for (int j = 0; j < 10; ++j) w[j] = mw[j];
#pragma unroll
            for (int j = 0; j < 9; ++j) w[j] = __funnelshift_r(w[j], w[j + 1], sh); // the 36 bytes, now word aligned
            uint32_t A[12];
#pragma unroll
            for (int p = 0; p < 12; ++p) {
                // bytes 3p,3p+1,3p+2 of the (unaligned) 36-byte group
                float ch[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int bi = 3 * p + c;
                    ch[c] = byte_to_float(w[bi >> 2], bi & 3);
                }
                const uint64_t pix = px0 + 12ull * u + p;
                uint32_t v = rgb_to_value(ch[0], ch[1], ch[2]);
                if (pix >= P.n_px) v = pix < 2 * g.n_words ? 797040u : 0u; // odd tail pairs with PixelYCbCrQuant{} (OLD:730); beyond: zero trits
                A[p] = v;
            }
            uint32_t o[13];
            {
                uint32_t w0, w1, w2, s12;
                triple_to_symbols(A[0], A[1], A[2], w0, w1, w2, s12);       // bytes 0..12
                o[0] = w0; o[1] = w1; o[2] = w2; o[3] = s12;
                triple_to_symbols(A[3], A[4], A[5], w0, w1, w2, s12);       // bytes 13..25
                o[3] |= w0 << 8; o[4] = __funnelshift_l(w0, w1, 8); o[5] = __funnelshift_l(w1, w2, 8); o[6] = __funnelshift_l(w2, s12, 8);
                triple_to_symbols(A[6], A[7], A[8], w0, w1, w2, s12);       // bytes 26..38
                o[6] |= w0 << 16; o[7] = __funnelshift_l(w0, w1, 16); o[8] = __funnelshift_l(w1, w2, 16); o[9] = __funnelshift_l(w2, s12, 16);
                triple_to_symbols(A[9], A[10], A[11], w0, w1, w2, s12);     // bytes 39..51
                o[9] |= w0 << 24; o[10] = __funnelshift_l(w0, w1, 24); o[11] = __funnelshift_l(w1, w2, 24); o[12] = __funnelshift_l(w2, s12, 24);
            }
            uint32_t* dst = reinterpret_cast<uint32_t*>(S) + 13 * u;
#pragma unroll
            for (int j = 0; j < 13; ++j) dst[j] = o[j];
        }
        __syncthreads();
        // ---- phase B: one codeword per thread iteration: 9-band gather (A.3), RS parity, scramble (A.4)
        for (int cw = tid; cw < L::NCW; cw += FAST_TPB) {
            const int b = cw / C_TILE, cl = cw - b * C_TILE;
            if ((uint32_t)cl >= run_n[b]) continue;
            const uint32_t rp = run_ph[b];
            const uint32_t ph = ((rp & 7) + 2u * (uint32_t)cl) % 6u;   // (p0 + 4) % 6 with p0 = 26*(cw_base_b + c)
            const bool first = (rp & 8) && cl == 0;                    // the block at body index 0 (LCG transient)
            uint32_t so[6]; // scramble row offset for symbol i: so[i%6]
#pragma unroll
            for (int j = 0; j < 6; ++j) so[j] = stoff[ph + j];
            const uint32_t so0 = first ? stoff[12] : so[0], so1 = first ? stoff[13] : so[1];
            const uint8_t* src = S + 9 * K * cl + b;
            uint8_t* dst = O + L::RUN_PITCH * b + (uint32_t)(run_lo[b] & 15) + 26 * cl;
            Planes acc{0, 0};
            uint32_t pair = 0;
#pragma unroll
            for (int i = 0; i < K; ++i) {
                const uint32_t d = src[9 * i];
                gf3_add(acc, tab[i * kVals + d]);
                const uint32_t sc = scr[(i == 0 ? so0 : i == 1 ? so1 : so[i % 6]) + d];
                if (i & 1) store2(dst + i - 1, pair, sc); else pair = sc;
            }
            const uint32_t lo = planes_to_sym4_lo(acc), hi = planes_to_sym4_hi(acc);
#pragma unroll
            for (int j = 0; j < L::R; ++j) {
                const uint32_t pj = (j < 4 ? lo >> (8 * j) : hi >> (8 * (j - 4))) & 0xFF;
                const uint32_t sc = scr[so[(K + j) % 6] + pj];
                if (j & 1) store2(dst + K + j - 1, pair, sc); else pair = sc;
            }
        }
        __syncthreads();
        // ---- phase C: nine band-major runs -> global (A.6 assembly: offset 52 + 26*(cw_base_b + c))
        store_runs9<L::RUN_PITCH>(O, P.out, run_lo, run_n);
    }
}

// =============================================================================================
// decode
// =============================================================================================
template <int K>
__global__ void __launch_bounds__(FAST_TPB, 4) k_decode_rgb_fast(FastParams P, Geom g, const GfTables* __restrict__ gf, const RsTables* __restrict__ rs)
{
    using L = Cfg<K>;
    extern __shared__ __align__(16) uint8_t smem[];
    uint64_t* tab = reinterpret_cast<uint64_t*>(smem + L::OFF_TAB);
    uint8_t* s_rgb = smem + L::OFF_RGB;
    uint8_t* S = smem + L::OFF_S;
    uint8_t* O = smem + L::OFF_O;
    GfTables& sg = *reinterpret_cast<GfTables*>(smem + L::OFF_GF);
    __shared__ uint64_t run_lo[9];
    __shared__ uint32_t run_n[9], run_ph[9];
    const int tid = threadIdx.x;
    __shared__ uint32_t stoff[14];
    {
        const RowTable& T = rs->row[1][(24 - K) / 2]; // the consistent decoder always uses the repaired code
        for (int i = tid; i < 26 * kVals; i += FAST_TPB) tab[i] = T.e[i / kVals][i % kVals];
        load_gf(sg, gf);
        if (tid < 14) stoff[tid] = 32u * (tid < 12 ? g.st[2 + tid % 6] : g.st[tid - 12]);
    }
    const uint64_t total = (uint64_t)P.n_tiles * P.n_frames;
    for (uint64_t tile_id = blockIdx.x; tile_id < total; tile_id += gridDim.x) {
        const uint32_t f = (uint32_t)(tile_id / P.n_tiles), tile = (uint32_t)(tile_id - (uint64_t)f * P.n_tiles);
        const uint8_t* in = P.in + P.in_stride * f;
        uint8_t* rgb = P.out + P.out_stride * f;
        const uint64_t c0 = (uint64_t)C_TILE * tile;
        __syncthreads();
        if (tid < 9) {
            const uint64_t n = c0 >= g.ncw[tid] ? 0 : ((g.ncw[tid] - c0) < C_TILE ? (g.ncw[tid] - c0) : C_TILE);
            run_lo[tid] = (uint64_t)(in - P.in) + 52 + 26 * (g.cw_base[tid] + c0);
            run_n[tid] = (uint32_t)n;
            run_ph[tid] = (uint32_t)((26 * (g.cw_base[tid] + c0) + 4) % 6) | (g.cw_base[tid] + c0 == 0 ? 8u : 0u);
        }
        __syncthreads();
        // ---- phase 0: nine runs -> shared
        load_runs9<L::RUN_PITCH>(O, P.in, run_lo, run_n, P.in_stride * (P.n_frames - 1) + 9 * g.n_out);
        __syncthreads();
        // ---- phase B: syndrome screen per codeword; dirty ones take BM/Chien/Forney; descramble; 9-band scatter
        for (int cw = tid; cw < L::NCW; cw += FAST_TPB) {
            const int b = cw / C_TILE, cl = cw - b * C_TILE;
            if ((uint32_t)cl >= run_n[b]) continue;
            const uint32_t rp = run_ph[b];
            const uint32_t ph = ((rp & 7) + 2u * (uint32_t)cl) % 6u;
            const bool first = (rp & 8) && cl == 0;
            const uint64_t p0 = 26 * (g.cw_base[b] + c0 + cl);
            uint32_t so[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) so[j] = stoff[ph + j];
            const uint32_t so0 = first ? stoff[12] : so[0], so1 = first ? stoff[13] : so[1];
            const uint8_t* src = O + L::RUN_PITCH * b + (uint32_t)(run_lo[b] & 15) + 26 * cl;
            uint8_t* dst = S + 9 * K * cl + b;
            Planes acc{0, 0};
#pragma unroll
            for (int i = 0; i < 26; i += 2) {
                const uint32_t two = load2(src + i);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int ii = i + h;
                    uint32_t s = (two >> (8 * h)) & 0xFF;
                    if (s >= 27) s %= 27; // out-of-alphabet bytes read as their low three trits, like unpack3 (OLD:28-31)
                    gf3_add(acc, tab[ii * kVals + s]);
                    if (ii < K) dst[9 * ii] = sg.dsc[0][(ii == 0 ? so0 : ii == 1 ? so1 : so[ii % 6]) + s];
                }
            }
            const int ci = first ? 6 : (int)ph;
            if (acc.nz != P.chk_nz[ci] || acc.two != P.chk_two[ci]) {
                // slow path: full decode of this codeword (descrambled), then rewrite its data symbols
                uint8_t cwd[26], orig[26];
                for (int i = 0; i < 26; ++i) {
                    const uint32_t st = scr_state(g, p0 + i);
                    cwd[i] = orig[i] = sg.dsc[st][src[i] % 27];
                }
                if (!rs_decode_thread(sg, cwd, K, true)) {
                    atomicExch(&P.status[2 * f], 0u);
                } else {
                    uint32_t nfix = 0;
                    for (int i = 0; i < 26; ++i) nfix += cwd[i] != orig[i];
                    if (nfix) atomicAdd(&P.status[2 * f + 1], nfix);
                    for (int i = 0; i < K; ++i) dst[9 * i] = cwd[i];
                }
            }
        }
        __syncthreads();
        // ---- phase A: 52 stream symbols -> 12 pixels -> RGB8 into the staging run
        const uint64_t px0 = (uint64_t)L::PX * tile;
        const uint64_t g_lo = (uint64_t)(rgb - P.out) + 3 * px0;
        const uint64_t px_hi = px0 + L::PX < P.px_out ? px0 + L::PX : P.px_out;
        const uint64_t g_hi = (uint64_t)(rgb - P.out) + 3 * (px_hi > px0 ? px_hi : px0);
        const uint32_t pad = (uint32_t)(g_lo & 15);
        for (int u = tid; u < L::UNITS; u += FAST_TPB) {
            const uint32_t* srcw = reinterpret_cast<const uint32_t*>(S) + 13 * u;
            uint32_t w[13];
#pragma unroll
            for (int j = 0; j < 13; ++j) w[j] = srcw[j];
            uint32_t pixrgb[12];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                // 13 bytes at byte offset 13t = word 3t + t/4.. : realign with funnel shifts by 8t bits
                uint32_t w0, w1, w2, s12;
                if (t == 0) { w0 = w[0]; w1 = w[1]; w2 = w[2]; s12 = w[3] & 0xFF; }
                else {
                    const int q = 3 * t, sh = 8 * t;
                    w0 = __funnelshift_r(w[q], w[q + 1], sh); w1 = __funnelshift_r(w[q + 1], w[q + 2], sh);
                    w2 = __funnelshift_r(w[q + 2], w[q + 3], sh); s12 = (w[q + 3] >> sh) & 0xFF;
                }
                uint32_t A0, A1, A2;
                symbols_to_triple(w0, w1, w2, s12, A0, A1, A2);
                pixrgb[3 * t] = value_to_rgb(A0);
                pixrgb[3 * t + 1] = value_to_rgb(A1);
                pixrgb[3 * t + 2] = value_to_rgb(A2);
            }
            uint8_t* me = s_rgb + pad + 36 * u;
            if ((pad & 3) == 0) {
                uint32_t* mw = reinterpret_cast<uint32_t*>(me);
#pragma unroll
                for (int q = 0; q < 3; ++q) { // 4 pixels = 12 bytes = 3 words
                    const uint32_t a = pixrgb[4 * q], b2 = pixrgb[4 * q + 1], c2 = pixrgb[4 * q + 2], d = pixrgb[4 * q + 3];
                    mw[3 * q] = a | (b2 << 24);
                    mw[3 * q + 1] = (b2 >> 8) | (c2 << 16);
                    mw[3 * q + 2] = (c2 >> 16) | (d << 8);
                }
            } else {
#pragma unroll
                for (int p = 0; p < 12; ++p) { me[3 * p] = (uint8_t)pixrgb[p]; me[3 * p + 1] = (uint8_t)(pixrgb[p] >> 8); me[3 * p + 2] = (uint8_t)(pixrgb[p] >> 16); }
            }
        }
        __syncthreads();
        store_run(s_rgb, P.out, g_lo, g_hi);
    }
}

template <int K>
int launch_enc(const DevTables& T, const FastParams& P, const Geom& g, cudaStream_t st)
{
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_encode_rgb_fast<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<K>::TOTAL); attr = true; }
    const uint64_t total = (uint64_t)P.n_tiles * P.n_frames;
    uint64_t grid = (uint64_t)T.sm_count * 4;
    if (grid > total) grid = total;
    if (!grid) return 0;
    k_encode_rgb_fast<K><<<(unsigned)grid, FAST_TPB, Cfg<K>::TOTAL, st>>>(P, g, T.gf, T.rs);
    return 1;
}
template <int K>
int launch_dec(const DevTables& T, const FastParams& P, const Geom& g, cudaStream_t st)
{
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(k_decode_rgb_fast<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<K>::TOTAL); attr = true; }
    const uint64_t total = (uint64_t)P.n_tiles * P.n_frames;
    uint64_t grid = (uint64_t)T.sm_count * 4;
    if (grid > total) grid = total;
    if (!grid) return 0;
    k_decode_rgb_fast<K><<<(unsigned)grid, FAST_TPB, Cfg<K>::TOTAL, st>>>(P, g, T.gf, T.rs);
    return 1;
}

} // namespace

bool fast_path_ok(const t3c_config& cfg)
{
    if (cfg.profile == T3C_PROFILE_RAW) return false;
    if (use_2d(cfg) || use_beacon(cfg)) return false;
    for (int b = 1; b < 9; ++b) if (cfg.uep[b] % 4 != cfg.uep[0] % 4) return false;
    return true;
}

int launch_encode_rgb_fast(const DevTables& T, const t3c_config& cfg, const Geom& g, const uint8_t* rgb, size_t n_px, size_t n_frames,
                           uint8_t* out, size_t stride_words, cudaStream_t st)
{
    if (((uintptr_t)rgb | (uintptr_t)out) & 15) return -1; // 128-bit transfers need 16-byte aligned buffer bases
    if (n_frames > 1 && (stride_words & 1)) return -1;     // and every frame's body must start on an even byte
    FastParams P{};
    P.in = rgb; P.out = out;
    P.in_stride = 3ull * n_px; P.out_stride = 9ull * stride_words;
    P.n_px = n_px; P.n_frames = (uint32_t)n_frames;
    uint64_t mx = 0;
    for (int b = 0; b < 9; ++b) mx = g.ncw[b] > mx ? g.ncw[b] : mx;
    P.n_tiles = (uint32_t)((mx + C_TILE - 1) / C_TILE);
    int n = 0;
    switch (g.uniform_k) {
    case 24: n = launch_enc<24>(T, P, g, st); break;
    case 22: n = launch_enc<22>(T, P, g, st); break;
    case 20: n = launch_enc<20>(T, P, g, st); break;
    case 18: n = launch_enc<18>(T, P, g, st); break;
    default: return 0;
    }
    return n + launch_frame_misc(T, cfg, g, out, n_frames, 9ull * stride_words, st);
}

int launch_decode_rgb_fast(const DevTables& T, const t3c_config& cfg, const Geom& g, const uint8_t* in, size_t stride_words,
                           size_t n_frames, size_t n_px, size_t n_px_out, uint8_t* rgb, uint32_t* d_status, cudaStream_t st,
                           const uint32_t* chk_nz, const uint32_t* chk_two)
{
    (void)cfg;
    if (((uintptr_t)rgb | (uintptr_t)in) & 15) return -1;
    if (n_frames > 1 && (stride_words & 1)) return -1;
    FastParams P{};
    P.in = in; P.out = rgb;
    P.in_stride = 9ull * stride_words; P.out_stride = 3ull * n_px;
    P.n_px = n_px; P.px_out = n_px_out; P.n_frames = (uint32_t)n_frames;
    P.status = d_status;
    for (int i = 0; i < 7; ++i) { P.chk_nz[i] = chk_nz[i]; P.chk_two[i] = chk_two[i]; }
    uint64_t mx = 0;
    for (int b = 0; b < 9; ++b) mx = g.ncw[b] > mx ? g.ncw[b] : mx;
    P.n_tiles = (uint32_t)((mx + C_TILE - 1) / C_TILE);
    switch (g.uniform_k) {
    case 24: return launch_dec<24>(T, P, g, st);
    case 22: return launch_dec<22>(T, P, g, st);
    case 20: return launch_dec<20>(T, P, g, st);
    case 18: return launch_dec<18>(T, P, g, st);
    }
    return 0;
}

} // namespace t3c
