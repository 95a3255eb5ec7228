// k_fast.cu -- tiled, fused kernels for the headline family: uniform k, 1D 9-band interleave, no beacon.
//
//   encode: RGB8 --(bridge, 13 trits/px)--> regrouped symbol stream --(9-band transpose in shared memory)-->
//           RS(26,k) parity (bit-plane row tables) --> scramble --> 9 band-major runs of the body
//   decode: 9 runs --> syndrome screen (same row tables) --> [dirty codewords: BM/Chien/Forney] --> descramble
//           --> 9-band transpose back --> 13 trits/px --> dequant --> RGB8
//
// Work unit = a mini-tile of 13 codewords per band (9*k*13 stream symbols = 27*k pixels = 9*k pixel triples),
// owned by ONE WARP from load to store: the three phases (bridge+regroup, RS+scramble, run copy) are separated
// by __syncwarp only, so there is no block-level barrier in the steady state and every warp always has work
// (the first CTA-tiled version spent 47 % of its warp-stall samples in bar.sync, profiles/r01_*).  Every global
// access is a 128-bit transfer inside a contiguous run (one RGB run, nine body runs of 338 B), the 9-band
// transpose and all table look-ups stay in shared memory, the grid is persistent (SM count x resident CTAs)
// so the row table is staged once per CTA, and HBM traffic is exactly the algorithmic 3 B/px + 9 B/word.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <type_traits>
#include <utility>
#include <vector>

#include "dev.cuh"
#include "launch.h"

namespace t3c {
namespace {

constexpr int C_MINI = 13;      // codewords per band per mini-tile: 9*k*13 symbols = 9*k whole pixel triples
constexpr int FAST_TPB = 256, FAST_WARPS = FAST_TPB / 32;
constexpr uint32_t M27 = 159072863u; // ceil(2^32/27): exact quotient for x < 2^26

struct FastParams {
    const uint8_t* in;      // encode: rgb frames   decode: profile words
    uint8_t* out;           // encode: profile words decode: rgb frames
    uint64_t in_stride;     // bytes between frames
    uint64_t out_stride;
    uint64_t n_px;          // pixels per frame
    uint64_t px_out;        // decode: pixels to write per frame
    uint32_t n_tiles;       // mini-tiles per frame handled by this launch
    uint32_t tile0;         // first mini-tile of each frame handled by this launch
    uint32_t n_frames;
    uint32_t* status;       // decode: {ok, n_corrected} per frame
    uint32_t chk_nz[7], chk_two[7]; // decode: sum of T_i[13*st_i] per scrambler phase (6) and for p0==0
    uint32_t flags;         // v5: 1 | delay << 8 = staggered start of every other warp (k_fast5.cuh, T3C_V5_FLAGS); 2 = decode: runs arrive by one 3-D tensor copy
    uint32_t ts_tiles;      // v5 encode with the tensor store: mini-tiles of a frame whose 23-chunk box stays inside a band's row (0: no tensor store)
    uint64_t band_stride;   // v5 decode with the tensor copy: bytes between the runs of neighbouring bands (26 * codewords per band, a multiple of 16)
};

template <int K> struct Cfg {
    static constexpr int R = 26 - K;
    static constexpr int TRIPLES = 9 * K;               // pixel triples (13 symbols each) per mini-tile
    static constexpr int PX = 3 * TRIPLES;
    static constexpr int RGB_BYTES = 3 * PX;
    static constexpr int SYM = 13 * TRIPLES;            // stream symbols per mini-tile = 9*K*C_MINI
    static constexpr int RUN = 26 * C_MINI;             // 338 bytes per band run
    static constexpr int RUN_PITCH = 368;               // >= 15 + RUN, multiple of 16
    static constexpr int NCW = 9 * C_MINI;
    // per-warp shared memory: S (symbol stream) | U (RGB run, later reused for the nine body runs) | meta
    static constexpr int S_BYTES = (SYM + 15) / 16 * 16;
    static constexpr int RGB_PITCH = (RGB_BYTES + 15 + 15) / 16 * 16 + 16;
    static constexpr int U_BYTES = RGB_PITCH > 9 * RUN_PITCH ? RGB_PITCH : 9 * RUN_PITCH;
    static constexpr int META_BYTES = 160;              // run_lo[9] (u64) | run_n[9] | run_ph[9] (u32)
    static constexpr int WARP_BYTES = S_BYTES + U_BYTES + META_BYTES;
    // CTA-shared
    static constexpr int OFF_TAB = 0;
    static constexpr int TAB_BYTES = 26 * kVals * 8;
    static constexpr int OFF_LUT = OFF_TAB + TAB_BYTES; // 3x32 scramble/descramble LUT + stoff[14]
    static constexpr int LUT_BYTES = 96 + 64;
    static constexpr int OFF_GF = OFF_LUT + LUT_BYTES;  // decode only: GF(27) tables of the slow path
    static constexpr int GF_BYTES = ((int)sizeof(GfTables) + 15) / 16 * 16;
    static constexpr int OFF_PXLUT = OFF_GF + GF_BYTES; // decode only: dequant + colour-matrix products per quantised value
    static constexpr int PXLUT_BYTES = 256 * 4 + 2 * 88 * 8;
    static constexpr int OFF_WARP_ENC = OFF_GF;
    static constexpr int OFF_WARP_DEC = OFF_PXLUT + PXLUT_BYTES;
    static constexpr int TOTAL_ENC = OFF_WARP_ENC + FAST_WARPS * WARP_BYTES;
    static constexpr int TOTAL_DEC = OFF_WARP_DEC + FAST_WARPS * WARP_BYTES;
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

// body runs start at even byte offsets (the launchers only take the fast path when every frame does)
__device__ __forceinline__ void store2(uint8_t* p, uint32_t lo, uint32_t hi) { *reinterpret_cast<uint16_t*>(p) = (uint16_t)(lo | (hi << 8)); }
__device__ __forceinline__ uint32_t load2(const uint8_t* p) { return *reinterpret_cast<const uint16_t*>(p); }

// pixel value -> RGB8 through per-component tables holding exactly the float32 terms of ycbcr_to_rgb
// (IMG:57-66): y, (0.344136f*cb, 1.772f*cb), (1.402f*cr, 0.714136f*cr); the sums keep the reference's order.
struct PxLut { float y[256]; float2 b[88]; float2 r[88]; };
__device__ __forceinline__ uint32_t value_to_rgb_lut(const PxLut& T, uint32_t A)
{
    const uint32_t q = A / 243u, Yq = A - 243u * q, ur = min(q / 81u, 87u), ub = q - 81u * (q / 81u);
    const float y = T.y[Yq];
    const float2 tb = T.b[ub], tr = T.r[ur];
    const float r = __fadd_rn(y, tr.x);
    const float g = __fsub_rn(__fsub_rn(y, tb.x), tr.y);
    const float b = __fadd_rn(y, tb.y);
    // clamp(round-half-away(x),0,255) == round(clamp(x,0,255)): FADD.RM against 2^22+0.5 leaves 2*value in the low bits
    const uint32_t R = (uint32_t)__float_as_int(__fadd_rd(fminf(fmaxf(r, 0.0f), 255.0f), 4194304.5f)) >> 1;
    const uint32_t G = (uint32_t)__float_as_int(__fadd_rd(fminf(fmaxf(g, 0.0f), 255.0f), 4194304.5f)) >> 1;
    const uint32_t B = (uint32_t)__float_as_int(__fadd_rd(fminf(fmaxf(b, 0.0f), 255.0f), 4194304.5f)) >> 1;
    return __byte_perm(__byte_perm(R, G, 0x0040), B, 0x7410) & 0x00FFFFFFu; // R | G<<8 | B<<16
}

struct WarpMeta { uint64_t run_lo[9]; uint32_t run_n[9]; uint32_t run_ph[9]; };
static_assert(sizeof(WarpMeta) <= 160, "meta");

// per-warp: where the nine body runs of mini-tile `tile` of a frame live (A.6: 52 + 26*(cw_base_b + c))
__device__ __forceinline__ void setup_runs(WarpMeta& m, const Geom& g, uint64_t frame_off, uint32_t tile, int lane)
{
    if (lane < 9) {
        const uint64_t c0 = (uint64_t)C_MINI * tile;
        const uint64_t n = c0 >= g.ncw[lane] ? 0 : ((g.ncw[lane] - c0) < C_MINI ? (g.ncw[lane] - c0) : C_MINI);
        const uint64_t cwi = g.cw_base[lane] + c0;
        m.run_lo[lane] = frame_off + 52 + 26 * cwi;
        m.run_n[lane] = (uint32_t)n;
        m.run_ph[lane] = (uint32_t)((26 * cwi + 4) % 6) | (cwi == 0 ? 8u : 0u); // scrambler phase of the run's first symbol
    }
}
// a contiguous global byte range -> shared; smem byte i <-> global byte (g_lo - g_lo%16 + i)
__device__ __forceinline__ void warp_load_run(uint8_t* s, const uint8_t* __restrict__ gbase, uint64_t g_lo, uint64_t g_hi, uint64_t g_limit, int lane)
{
    const uint64_t a0 = g_lo & ~15ull;
    const int nchunk = (int)((g_hi - a0 + 15) >> 4);
    for (int c = lane; c < nchunk; c += 32) {
        const uint64_t ga = a0 + 16ull * c;
        if (ga + 16 <= g_limit) *reinterpret_cast<uint4*>(s + 16 * c) = __ldg(reinterpret_cast<const uint4*>(gbase + ga));
        else
            for (int i = 0; i < 16; ++i) s[16 * c + i] = ga + i < g_limit ? gbase[ga + i] : 0;
    }
}
// shared -> global of one run [g_lo, g_lo+len): 128-bit stores for the interior 16-byte chunks, and the (at most
// 15+15) edge bytes written one byte per lane, so that no lane ever walks an edge chunk alone
__device__ __forceinline__ void warp_store_run(const uint8_t* s, uint8_t* __restrict__ gbase, uint64_t g_lo, int len, int lane)
{
    if (len <= 0) return;
    const int pad = (int)(g_lo & 15);
    uint8_t* g0 = gbase + (g_lo - pad);               // 16-byte aligned; smem byte i <-> g0[i]
    const int first_full = pad ? 1 : 0, end = pad + len, last_full = end >> 4; // chunks [first_full, last_full) are interior
    for (int c = first_full + lane; c < last_full; c += 32) *reinterpret_cast<uint4*>(g0 + 16 * c) = *reinterpret_cast<const uint4*>(s + 16 * c);
    {   // head bytes [pad, min(16,end)) and tail bytes [max(16*last_full, pad), end)
        const int hi = lane < 16 ? lane : 16 * last_full + (lane - 16);
        const bool in_head = lane < 16 && pad && hi >= pad && hi < end && hi < 16;
        const bool in_tail = lane >= 16 && hi >= pad && hi < end && (last_full >= first_full) && !(last_full == 0 && pad);
        if (in_head || in_tail) g0[hi] = s[hi];
    }
}
// nine runs, one per iteration: lanes 0..CH-1 move one interior 16-byte chunk each, and in the same pass
// lanes 0-15 / 16-31 write the head / tail edge bytes, so the whole run costs ~20 warp instructions
template <int PITCH>
__device__ __forceinline__ void warp_store_runs9(const uint8_t* O, uint8_t* __restrict__ gbase, const WarpMeta& m, int lane)
{
    static_assert(PITCH / 16 <= 32, "one chunk per lane");
#pragma unroll 1
    for (int b = 0; b < 9; ++b) {
        const int len = 26 * (int)m.run_n[b];
        if (len == 0) continue;
        const uint64_t lo = m.run_lo[b];
        const int pad = (int)(lo & 15), end = pad + len;
        uint8_t* g0 = gbase + (lo - pad);
        const uint8_t* s0 = O + PITCH * b;
        const int c16 = 16 * lane;
        if (c16 >= pad && c16 + 16 <= end) *reinterpret_cast<uint4*>(g0 + c16) = *reinterpret_cast<const uint4*>(s0 + c16);
        const int pos = lane < 16 ? lane : (end & ~15) + (lane - 16);
        const bool head = lane < 16 && pad != 0;
        const bool tail = lane >= 16 && (end & 15) != 0 && !(pad != 0 && (end >> 4) == 0);
        if ((head || tail) && pos >= pad && pos < end) g0[pos] = s0[pos];
    }
}
template <int PITCH>
__device__ __forceinline__ void warp_load_runs9(uint8_t* O, const uint8_t* __restrict__ gbase, const WarpMeta& m, uint64_t g_limit, int lane)
{
    static_assert(PITCH / 16 <= 32, "one chunk per lane");
#pragma unroll 1
    for (int b = 0; b < 9; ++b) {
        const int len = 26 * (int)m.run_n[b];
        if (len == 0) continue;
        const uint64_t lo = m.run_lo[b];
        const int pad = (int)(lo & 15), c16 = 16 * lane;
        if (c16 >= pad + len) continue;
        const uint64_t ga = (lo - pad) + c16;
        uint8_t* sp = O + PITCH * b + c16;
        if (ga + 16 <= g_limit) *reinterpret_cast<uint4*>(sp) = __ldg(reinterpret_cast<const uint4*>(gbase + ga));
        else
            for (int i = 0; i < 16; ++i) sp[i] = ga + i < g_limit ? gbase[ga + i] : 0;
    }
}

// =============================================================================================
// encode
// =============================================================================================
template <int K>
__global__ void __launch_bounds__(FAST_TPB, 4) k_encode_rgb_fast(FastParams P, Geom g, const GfTables* __restrict__ gf, const RsTables* __restrict__ rs)
{
    using L = Cfg<K>;
    extern __shared__ __align__(16) uint8_t smem[];
    const uint64_t* tab = reinterpret_cast<const uint64_t*>(smem + L::OFF_TAB);
    const uint8_t* scr = smem + L::OFF_LUT;                               // 3 x 32 scramble look-up
    const uint32_t* stoff = reinterpret_cast<const uint32_t*>(smem + L::OFF_LUT + 96); // 32*st: phase 0..11, body index 0,1
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* S = smem + L::OFF_WARP_ENC + warp * L::WARP_BYTES;
    uint8_t* U = S + L::S_BYTES;
    WarpMeta& meta = *reinterpret_cast<WarpMeta*>(U + L::U_BYTES);
    {
        const RowTable& T = rs->row[g.arith][(24 - K) / 2];
        uint64_t* t = reinterpret_cast<uint64_t*>(smem + L::OFF_TAB);
        for (int i = tid; i < K * kVals; i += FAST_TPB) t[i] = T.e[i / kVals][i % kVals];
        if (tid < 96) smem[L::OFF_LUT + tid] = gf->scr[tid / 32][tid % 32];
        if (tid < 14) reinterpret_cast<uint32_t*>(smem + L::OFF_LUT + 96)[tid] = 32u * (tid < 12 ? g.st[2 + tid % 6] : g.st[tid - 12]);
    }
    __syncthreads(); // the only block-level barrier: tables staged
    const uint64_t total = (uint64_t)P.n_tiles * P.n_frames;
    const uint64_t nwarps = (uint64_t)gridDim.x * FAST_WARPS;
    for (uint64_t mt = (uint64_t)blockIdx.x * FAST_WARPS + warp; mt < total; mt += nwarps) {
        const uint32_t f = (uint32_t)(mt / P.n_tiles), tile = P.tile0 + (uint32_t)(mt - (uint64_t)f * P.n_tiles);
        const uint64_t px0 = (uint64_t)L::PX * tile;
        const uint64_t in_off = P.in_stride * f;
        const uint64_t g_lo = in_off + 3 * (px0 < P.n_px ? px0 : P.n_px);
        const uint64_t g_hi = in_off + 3 * (px0 + L::PX < P.n_px ? px0 + L::PX : P.n_px);
        const int pad = (int)(g_lo & 15);
        // pixels of this mini-tile that exist / that are the odd tail's default partner (OLD:730)
        const int lim = px0 >= P.n_px ? 0 : (P.n_px - px0 < (uint64_t)L::PX ? (int)(P.n_px - px0) : L::PX);
        const int lim2 = px0 >= 2 * g.n_words ? 0 : (2 * g.n_words - px0 < (uint64_t)L::PX ? (int)(2 * g.n_words - px0) : L::PX);
        setup_runs(meta, g, P.out_stride * f, tile, lane);
        warp_load_run(U, P.in, g_lo, g_hi, P.in_stride * P.n_frames, lane);
        __syncwarp();
        // ---- phase A: two pixel triples (18 bytes) -> 26 stream symbols per lane iteration (bridge + A.1 regroup)
        for (int u = lane; u < L::TRIPLES / 2; u += 32) {
            const uint32_t a = (uint32_t)pad + 18u * (uint32_t)u;
            const uint32_t* mw = reinterpret_cast<const uint32_t*>(U + (a & ~3u));
            const uint32_t sh = (a & 3u) * 8u;
            uint32_t x[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) x[j] = mw[j];
            uint32_t y[5]; // the 18 bytes, word aligned
#pragma unroll
            for (int j = 0; j < 4; ++j) y[j] = __funnelshift_r(x[j], x[j + 1], sh);
            y[4] = __funnelshift_r(x[4], sh == 24 ? mw[5] : 0u, sh); // 18 bytes from byte offset 3 reach into a sixth word (odd frame offsets)
            uint32_t A[6];
#pragma unroll
            for (int p = 0; p < 6; ++p) {
                const int q = 3 * p;
                A[p] = rgb_to_value(byte_to_float(y[q >> 2], q & 3), byte_to_float(y[(q + 1) >> 2], (q + 1) & 3), byte_to_float(y[(q + 2) >> 2], (q + 2) & 3));
            }
            if (lim < L::PX) { // ragged end of the frame (warp-uniform)
#pragma unroll
                for (int p = 0; p < 6; ++p) { const int lp = 6 * u + p; if (lp >= lim) A[p] = lp < lim2 ? 797040u : 0u; }
            }
            uint32_t w0, w1, w2, s12, v0, v1, v2, t12;
            triple_to_symbols(A[0], A[1], A[2], w0, w1, w2, s12);
            triple_to_symbols(A[3], A[4], A[5], v0, v1, v2, t12);
            uint16_t* d = reinterpret_cast<uint16_t*>(S + 26 * u); // 13 halfwords (STS.U16 keeps the low 16 bits)
            d[0] = (uint16_t)w0; d[1] = (uint16_t)(w0 >> 16); d[2] = (uint16_t)w1; d[3] = (uint16_t)(w1 >> 16);
            d[4] = (uint16_t)w2; d[5] = (uint16_t)(w2 >> 16); d[6] = (uint16_t)(s12 | (v0 << 8));
            d[7] = (uint16_t)(v0 >> 8); d[8] = (uint16_t)__funnelshift_r(v0, v1, 24); d[9] = (uint16_t)(v1 >> 8);
            d[10] = (uint16_t)__funnelshift_r(v1, v2, 24); d[11] = (uint16_t)(v2 >> 8); d[12] = (uint16_t)((v2 >> 24) | (t12 << 8));
        }
        __syncwarp();
        // ---- phase B: one codeword per lane iteration: 9-band gather (A.3), RS parity, scramble (A.4)
        for (int cw = lane; cw < L::NCW; cw += 32) {
            const int b = cw / C_MINI, cl = cw - b * C_MINI;
            if ((uint32_t)cl >= meta.run_n[b]) continue;
            const uint32_t rp = meta.run_ph[b];
            const uint32_t ph = ((rp & 7) + 2u * (uint32_t)cl) % 6u;   // (p0 + 4) % 6 with p0 = 26*(cw_base_b + c)
            const bool first = (rp & 8) && cl == 0;                    // the block at body index 0 (LCG transient)
            uint32_t so[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) so[j] = stoff[ph + j];
            const uint32_t so0 = first ? stoff[12] : so[0], so1 = first ? stoff[13] : so[1];
            const uint8_t* src = S + 9 * K * cl + b;
            uint8_t* dst = U + L::RUN_PITCH * b + (uint32_t)(meta.run_lo[b] & 15) + 26 * cl;
            uint32_t d[K];
#pragma unroll
            for (int i = 0; i < K; ++i) d[i] = src[9 * i];
            Planes acc{0, 0}, acc2{0, 0};
            uint32_t pair = 0;
#pragma unroll
            for (int i = 0; i < K; ++i) {
                const uint64_t e = tab[i * kVals + d[i]];
                if (i & 1) gf3_add(acc2, e); else gf3_add(acc, e);
                const uint32_t sc = scr[(i == 0 ? so0 : i == 1 ? so1 : so[i % 6]) + d[i]];
                if (i & 1) store2(dst + i - 1, pair, sc); else pair = sc;
            }
            gf3_add(acc, acc2.nz, acc2.two);
            const uint32_t lo = planes_to_sym4_lo(acc), hi = planes_to_sym4_hi(acc);
#pragma unroll
            for (int j = 0; j < L::R; ++j) {
                const uint32_t pj = (j < 4 ? lo >> (8 * j) : hi >> (8 * (j - 4))) & 0xFF;
                const uint32_t sc = scr[so[(K + j) % 6] + pj];
                if (j & 1) store2(dst + K + j - 1, pair, sc); else pair = sc;
            }
        }
        __syncwarp();
        // ---- phase C: nine band-major runs -> global
        warp_store_runs9<L::RUN_PITCH>(U, P.out, meta, lane);
        __syncwarp();
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
    const uint64_t* tab = reinterpret_cast<const uint64_t*>(smem + L::OFF_TAB);
    const uint8_t* dsc = smem + L::OFF_LUT;
    const uint32_t* stoff = reinterpret_cast<const uint32_t*>(smem + L::OFF_LUT + 96);
    GfTables& sg = *reinterpret_cast<GfTables*>(smem + L::OFF_GF);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* S = smem + L::OFF_WARP_DEC + warp * L::WARP_BYTES;
    uint8_t* U = S + L::S_BYTES;
    WarpMeta& meta = *reinterpret_cast<WarpMeta*>(U + L::U_BYTES);
    {
        const RowTable& T = rs->row[1][(24 - K) / 2]; // the consistent decoder always uses the repaired code
        uint64_t* t = reinterpret_cast<uint64_t*>(smem + L::OFF_TAB);
        for (int i = tid; i < 26 * kVals; i += FAST_TPB) t[i] = T.e[i / kVals][i % kVals];
        if (tid < 96) smem[L::OFF_LUT + tid] = gf->dsc[tid / 32][tid % 32];
        if (tid < 14) reinterpret_cast<uint32_t*>(smem + L::OFF_LUT + 96)[tid] = 32u * (tid < 12 ? g.st[2 + tid % 6] : g.st[tid - 12]);
        load_gf(sg, gf);
        PxLut& W = *reinterpret_cast<PxLut*>(smem + L::OFF_PXLUT);
        if (tid < 256) W.y[tid] = (float)dequant_y(tid);
        if (tid < 88) {
            const float c = __fsub_rn((float)dequant_c(tid - 40), 128.0f);
            W.b[tid] = make_float2(__fmul_rn(0.344136f, c), __fmul_rn(1.772f, c));
            W.r[tid] = make_float2(__fmul_rn(1.402f, c), __fmul_rn(0.714136f, c));
        }
    }
    __syncthreads();
    const PxLut& pxlut = *reinterpret_cast<const PxLut*>(smem + L::OFF_PXLUT);
    const uint64_t total = (uint64_t)P.n_tiles * P.n_frames;
    const uint64_t nwarps = (uint64_t)gridDim.x * FAST_WARPS;
    const uint64_t in_limit = P.in_stride * (P.n_frames - 1) + 9 * g.n_out;
    for (uint64_t mt = (uint64_t)blockIdx.x * FAST_WARPS + warp; mt < total; mt += nwarps) {
        const uint32_t f = (uint32_t)(mt / P.n_tiles), tile = P.tile0 + (uint32_t)(mt - (uint64_t)f * P.n_tiles);
        setup_runs(meta, g, P.in_stride * f, tile, lane);
        __syncwarp();
        warp_load_runs9<L::RUN_PITCH>(U, P.in, meta, in_limit, lane);
        __syncwarp();
        // ---- phase B: syndrome screen per codeword; dirty ones take BM/Chien/Forney; descramble; 9-band scatter
        for (int cw = lane; cw < L::NCW; cw += 32) {
            const int b = cw / C_MINI, cl = cw - b * C_MINI;
            if ((uint32_t)cl >= meta.run_n[b]) continue;
            const uint32_t rp = meta.run_ph[b];
            const uint32_t ph = ((rp & 7) + 2u * (uint32_t)cl) % 6u;
            const bool first = (rp & 8) && cl == 0;
            uint32_t so[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) so[j] = stoff[ph + j];
            const uint32_t so0 = first ? stoff[12] : so[0], so1 = first ? stoff[13] : so[1];
            const uint8_t* src = U + L::RUN_PITCH * b + (uint32_t)(meta.run_lo[b] & 15) + 26 * cl;
            uint8_t* dst = S + 9 * K * cl + b;
            Planes acc{0, 0}, acc2{0, 0};
#pragma unroll
            for (int i = 0; i < 26; i += 2) {
                const uint32_t two = load2(src + i);
                uint32_t s0 = two & 0xFF, s1 = two >> 8;
                if (s0 >= 27) s0 %= 27; // out-of-alphabet bytes read as their low three trits, like unpack3 (OLD:28-31)
                if (s1 >= 27) s1 %= 27;
                gf3_add(acc, tab[i * kVals + s0]);
                gf3_add(acc2, tab[(i + 1) * kVals + s1]);
                if (i < K) dst[9 * i] = dsc[(i == 0 ? so0 : so[i % 6]) + s0];
                if (i + 1 < K) dst[9 * (i + 1)] = dsc[(i == 0 ? so1 : so[(i + 1) % 6]) + s1];
            }
            gf3_add(acc, acc2.nz, acc2.two);
            const int ci = first ? 6 : (int)ph;
            if (acc.nz != P.chk_nz[ci] || acc.two != P.chk_two[ci]) {
                // slow path: full decode of this codeword (descrambled), then rewrite its data symbols
                const uint64_t p0 = 26 * (g.cw_base[b] + (uint64_t)C_MINI * tile + cl);
                uint8_t cwd[26], orig[26];
                for (int i = 0; i < 26; ++i) cwd[i] = orig[i] = sg.dsc[scr_state(g, p0 + i)][src[i] % 27];
                if (!rs_decode_thread(sg, cwd, K, true, true)) {
                    atomicExch(&P.status[2 * f], 0u);
                } else {
                    uint32_t nfix = 0;
                    for (int i = 0; i < 26; ++i) nfix += cwd[i] != orig[i];
                    if (nfix) atomicAdd(&P.status[2 * f + 1], nfix);
                    for (int i = 0; i < K; ++i) dst[9 * i] = cwd[i];
                }
            }
        }
        __syncwarp();
        // ---- phase A: 13 stream symbols -> one pixel triple -> 9 RGB bytes per lane iteration
        const uint64_t px0 = (uint64_t)L::PX * tile;
        const uint64_t out_off = P.out_stride * f;
        const uint64_t g_lo = out_off + 3 * (px0 < P.px_out ? px0 : P.px_out);
        const int len = px0 >= P.px_out ? 0 : 3 * (int)(P.px_out - px0 < (uint64_t)L::PX ? P.px_out - px0 : L::PX);
        const uint32_t pad = (uint32_t)(g_lo & 15);
        for (int u = lane; u < L::TRIPLES / 2; u += 32) {
            const uint32_t a = 26u * (uint32_t)u;
            const uint32_t* mw = reinterpret_cast<const uint32_t*>(S + (a & ~3u));
            const uint32_t sh = (a & 2u) * 8u;
            uint32_t x[7];
#pragma unroll
            for (int j = 0; j < 7; ++j) x[j] = mw[j];
            uint32_t y[7]; // the 26 symbols, word aligned
#pragma unroll
            for (int j = 0; j < 6; ++j) y[j] = __funnelshift_r(x[j], x[j + 1], sh);
            y[6] = x[6] >> sh;
            uint32_t A[6];
            symbols_to_triple(y[0], y[1], y[2], y[3] & 0xFF, A[0], A[1], A[2]);
            symbols_to_triple(__funnelshift_r(y[3], y[4], 8), __funnelshift_r(y[4], y[5], 8), __funnelshift_r(y[5], y[6], 8), (y[6] >> 8) & 0xFF, A[3], A[4], A[5]);
            uint32_t p[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) p[q] = value_to_rgb_lut(pxlut, A[q]);
            uint8_t* d = U + pad + 18 * u;
            if ((pad & 1) == 0) {
                uint16_t* dh = reinterpret_cast<uint16_t*>(d);
#pragma unroll
                for (int q = 0; q < 3; ++q) { // two pixels = three halfwords
                    dh[3 * q] = (uint16_t)p[2 * q];
                    dh[3 * q + 1] = (uint16_t)((p[2 * q] >> 16) | (p[2 * q + 1] << 8));
                    dh[3 * q + 2] = (uint16_t)(p[2 * q + 1] >> 8);
                }
            } else {
#pragma unroll
                for (int q = 0; q < 6; ++q) { d[3 * q] = (uint8_t)p[q]; d[3 * q + 1] = (uint8_t)(p[q] >> 8); d[3 * q + 2] = (uint8_t)(p[q] >> 16); }
            }
        }
        __syncwarp();
        warp_store_run(U, P.out, g_lo, len, lane);
        __syncwarp();
    }
}

// =============================================================================================
// v3: full mini-tiles (all 117 codewords and all 27k pixels exist; the ragged last tiles of a frame keep the
// kernels above).  Same warp-autonomous structure, rebuilt around what the v2 profile showed
// (profiles/r01a_fast_v2_ncu_summary.txt): the kernels were bound by shared-memory wavefronts (bank conflicts on
// 64-bit row-table look-ups and on 2-byte accesses at 18/26-byte lane strides) and by the half-rate ALU pipe.
//   * row tables split into two conflict-free 32-bit plane tables of 27 (encode) / 32 (decode) entries per
//     position; the low byte of a plane entry carries the scrambled (descrambled) symbol for that position, one
//     table variant per scrambler phase of the codeword (26 = 2 mod 6: three variants), so the separate
//     scramble look-up and its address arithmetic disappear;
//   * symbols travel pre-scaled by 4 (a table byte offset), so a look-up is LDS [reg + immediate];
//   * parity leaves the bit planes through PRMT used as an 8-entry byte table (nibble = 3 plane bits), the
//     scrambler is added to the parity in the plane domain (3 LOP3 for all parity symbols);
//   * codewords are dealt to lanes row-major (9 bands of one codeword row are adjacent lanes) and 6-pixel units
//     even/odd, which makes the 9-byte-stride gathers and the 18/26-byte-stride accesses conflict-free;
//   * the bridge uses PRMT-built 2^23+x floats inside FFMAs (exactly fl(c*x)), two FADD.RM for the rounding and
//     one IMAD.HI per quantiser (constants found by exhaustive search, checked by the 2^24 colour test).
// =============================================================================================
constexpr uint32_t QY_M = 4076008176u, QY_Z = 114u, QY_C0 = 1194143128u;  // hi32((0x4B000000+Z+Y)*M) - C0 == (484Y+255)/510, Y in [0,255]
constexpr uint32_t QC_M = 1344274432u, QC_Z = 126u, QC_C0 = 393830439u;   // ... == (5C+7+(C>=128))>>4, C in [0,256]
constexpr uint32_t Q_CONST = 0u - (QY_C0 + 243u * QC_C0 + 19683u * QC_C0);
__device__ __forceinline__ uint32_t mad_hi(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// byte j of w as the float 2^23 + byte
__device__ __forceinline__ float byte_magic(uint32_t w, int j) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u | (uint32_t)j)); }
// rgb_to_ycbcr + quantize_ycbcr (IMG:47-56,69-78) on magic-form bytes -> 13-trit pixel value.
// fma(c, 2^23+x, -c*2^23) = fl(c*x) exactly (c*2^23 is representable, the fma rounds once), 0.5*x is exact.
__device__ __forceinline__ uint32_t rgb_to_value3(float R, float G, float B)
{
    constexpr float T23 = 8388608.0f;
    const float y = __fadd_rn(__fadd_rn(__fmaf_rn(0.299f, R, -0.299f * T23), __fmaf_rn(0.587f, G, -0.587f * T23)), __fmaf_rn(0.114f, B, -0.114f * T23));
    const float cb = __fadd_rn(__fadd_rn(__fsub_rn(__fmaf_rn(-0.168736f, R, 0.168736f * T23), __fmaf_rn(0.331264f, G, -0.331264f * T23)),
                                         __fmaf_rn(0.5f, B, -0.5f * T23)), 128.0f);
    const float cr = __fadd_rn(__fsub_rn(__fsub_rn(__fmaf_rn(0.5f, R, -0.5f * T23), __fmaf_rn(0.418688f, G, -0.418688f * T23)),
                                         __fmaf_rn(0.081312f, B, -0.081312f * T23)), 128.0f);
    // floor(v + 0.5) into the low mantissa bits, offset by Z: both adds round down, the second one is exact on integers
    const uint32_t by = (uint32_t)__float_as_int(__fadd_rd(__fadd_rd(y, 0.5f), T23 + (float)QY_Z));
    const uint32_t bb = (uint32_t)__float_as_int(__fadd_rd(__fadd_rd(cb, 0.5f), T23 + (float)QC_Z));
    const uint32_t br = (uint32_t)__float_as_int(__fadd_rd(__fadd_rd(cr, 0.5f), T23 + (float)QC_Z));
    const uint32_t fy = mad_hi(by, QY_M, Q_CONST);
    return fy + 243u * __umulhi(bb, QC_M) + 19683u * __umulhi(br, QC_M);
}
__device__ __forceinline__ uint32_t floor_sat_u8(float x)
{
    uint32_t r;
    asm("cvt.rmi.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// pixel value -> RGB8 (decode_raw_words_to_pixels + dequantize_ycbcr + ycbcr_to_rgb, OLD:706-722, IMG:57-84), arithmetic only
__device__ __forceinline__ uint32_t value_to_rgb3(uint32_t A)
{
    const uint32_t q = __umulhi(A, 17674763u);                 // A / 243, exact for A < 3^13
    const uint32_t Yq = A - 243u * q;
    const uint32_t ur = __umulhi(q, 53024288u);                // q / 81, exact for q < 6561
    const uint32_t ub = q - 81u * ur;
    // Y = (510 Yq + 241) / 484 (dev.cuh dequant_y; <= 255 for Yq <= 242), C = min((64 u + 10) / 20, 255)
    const float y = __fadd_rn(__uint_as_float(mad_hi(Yq * 510u + 241u, 8873899u, 0x4B000000u)), -8388608.0f);
    const float cb = __fadd_rn(__uint_as_float(min(mad_hi(32u * ub + 5u, 429496730u, 0x4B000000u), 0x4B0000FFu)), -8388736.0f);
    const float cr = __fadd_rn(__uint_as_float(min(mad_hi(32u * ur + 5u, 429496730u, 0x4B000000u), 0x4B0000FFu)), -8388736.0f);
    const float r = __fadd_rn(y, __fmul_rn(1.402f, cr));
    const float g = __fsub_rn(__fsub_rn(y, __fmul_rn(0.344136f, cb)), __fmul_rn(0.714136f, cr));
    const float b = __fadd_rn(y, __fmul_rn(1.772f, cb));
    // clamp(round-half-away(x), 0, 255) = sat_u8(floor(x + 0.5)) with the add rounded down (a round-to-nearest add could reach the next
    // integer from just below a half): one FADD.RM and one saturating conversion per component instead of two FMNMX and two FADDs, and
    // the bytes are packed by multiply-adds: the clamps and PRMTs left the half-rate ALU pipe, which bounds this kernel
    const uint32_t Rb = floor_sat_u8(__fadd_rd(r, 0.5f)), Gb = floor_sat_u8(__fadd_rd(g, 0.5f)), Bb = floor_sat_u8(__fadd_rd(b, 0.5f));
    return Rb + 256u * Gb + 65536u * Bb;                                       // R | G<<8 | B<<16
}
// 8-entry byte table through PRMT: nibble n of sel (3 plane bits b0 b1 b2) -> b0 + 3 b1 + 9 b2
// (PRMT itself, not __byte_perm: the intrinsic masks bit 3 of every selector nibble first; plane nibbles hold three bits, bit 3 is never set)
__device__ __forceinline__ uint32_t planes4_to_sym(uint32_t sel)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(0x04030100u), "r"(0x0D0C0A09u), "r"(sel));
    return r;
}
// first plane bit of parity symbol j: nibbles while they fit (r <= 6), dense 3-bit groups for RS(26,18) (8 x 3 trits = bits 8..31)
template <int K> __host__ __device__ constexpr int plane_shift(int j) { return 26 - K <= 6 ? 8 + 4 * j : 8 + 3 * j; }
// eight 3-bit groups (24 bits) -> eight nibbles
__device__ __forceinline__ uint32_t spread3to4(uint32_t x)
{
    x = (x & 0x00000FFFu) | ((x & 0x00FFF000u) << 4);
    x = (x & 0x003F003Fu) | ((x & 0x0FC00FC0u) << 2);
    return (x & 0x07070707u) | ((x & 0x38383838u) << 1);
}
// the parity symbols held by a pair of planes, as bytes: lo = symbols 0..3, hi = symbols 4..7
template <int K>
__device__ __forceinline__ void planes_to_parity(uint32_t nz, uint32_t two, uint32_t& lo, uint32_t& hi)
{
    uint32_t nzp = nz >> 8, twp = two >> 8;
    if constexpr (26 - K > 6) { nzp = spread3to4(nzp); twp = spread3to4(twp); }
    lo = planes4_to_sym(nzp) + planes4_to_sym(twp);
    hi = planes4_to_sym(nzp >> 16) + planes4_to_sym(twp >> 16);
}

template <int K> struct Cfg3 {
    static_assert(K == 18 || K == 20 || K == 22 || K == 24, "RS(26,k) profiles");
    static constexpr int R = 26 - K;
    static constexpr int TRIPLES = 9 * K, PX = 27 * K, RGB_BYTES = 81 * K, SYM = 117 * K, UNITS = TRIPLES / 2;
    static constexpr int NCW = 9 * C_MINI, RUN = 26 * C_MINI;
    static constexpr int PASS_A = (UNITS + 31) / 32, PASS_B = (NCW + 31) / 32;
    static constexpr int RUN_PITCH = 368;
    static constexpr int S_BYTES = (SYM + 15) / 16 * 16;
    static constexpr int RGB_PITCH = (RGB_BYTES + 15 + 15) / 16 * 16 + 16;
    static constexpr int U_BYTES = RGB_PITCH > 9 * RUN_PITCH ? RGB_PITCH : 9 * RUN_PITCH;
    static constexpr int META_BYTES = 96;               // run_lo[9] (u64) | vb[9] (u8)
    static constexpr int CARRY_BYTES = 9 * 16;          // the partial last 16-byte chunk of every output stream, kept for the next tile
    static constexpr int WARP_BYTES = S_BYTES + U_BYTES + META_BYTES + CARRY_BYTES;
    // encode, CTA-shared: per variant {A[K][27] | B[K][27]} | pat[3][2]  (B is the same in every variant: one address register serves both planes)
    static_assert(RUN_PITCH == 16 * 23, "chunk slots per run");
    static constexpr int ENC_A = 0, ENC_PLANE = 4 * K * 27, ENC_VAR = 2 * ENC_PLANE, ENC_PAT = ENC_A + 3 * ENC_VAR, ENC_MAP = (ENC_PAT + 24 + 15) / 16 * 16, ENC_WARP = ENC_MAP + 3 * 128;
    static constexpr int TOTAL_ENC = ENC_WARP + FAST_WARPS * WARP_BYTES;
    // decode, CTA-shared (offsets from a 256-byte aligned base): per variant {A[26][32], B[26][32]} | chk[3][2] | GF(27) tables
    // of the slow path.  A variant block is 26*256 bytes, so the low byte of a row's address is zero and PRMT can drop a
    // received symbol (x4, < 128) straight into it.
    static constexpr int DEC_A = 0, DEC_PLANE = 4 * 26 * 32, DEC_VAR = 2 * DEC_PLANE, DEC_CHK = 3 * DEC_VAR, DEC_GF = (DEC_CHK + 24 + 15) / 16 * 16;
    static constexpr int DEC_MAP = DEC_GF + ((int)sizeof(GfTables) + 15) / 16 * 16;
    static constexpr int DEC_WARP = DEC_MAP + 3 * 128;
    static constexpr int TOTAL_DEC = 256 + DEC_WARP + FAST_WARPS * WARP_BYTES;
};
struct WarpMeta3 { uint64_t run_lo[9]; uint8_t vb[16]; };
static_assert(sizeof(WarpMeta3) <= 96, "meta3");

// the six (periodic) scrambler states seen by a codeword of variant v (= codeword index mod 3) at position i
__device__ __forceinline__ uint32_t st_of(const Geom& g, int v, int i) { return g.st[2 + (2 * v + i + 4) % 6]; }

__device__ __forceinline__ void setup_runs3(WarpMeta3& m, const Geom& g, uint64_t frame_off, uint32_t tile, int lane)
{
    if (lane < 9) {
        const uint64_t cwi = g.cw_base[lane] + (uint64_t)C_MINI * tile;
        m.run_lo[lane] = frame_off + 52 + 26 * cwi;
        m.vb[lane] = (uint8_t)(cwi % 3);
    }
}
// Output streams are written as whole 16-byte chunks only.  A warp owns a contiguous range of mini-tiles, so the
// partial chunk at the end of a tile's run is completed by the same warp's next tile: it is parked in `carry` and
// restored at offset 0 of the next staging buffer.  Only the first / last tile of a range (or of a frame) writes
// edge bytes one by one.  len = bytes of this tile's piece, piece at staging offset pad = lo & 15.
__device__ __forceinline__ void stream_store(const uint8_t* s0, uint8_t* __restrict__ gbase, uint64_t lo, int len, bool first, bool last, uint4* carry, int lane)
{
    const int pad = (int)(lo & 15), end = pad + len, cend = end >> 4;
    uint8_t* g0 = gbase + (lo - pad);
    const int c0 = (first && pad) ? 1 : 0;
    for (int c = c0 + lane; c < cend; c += 32) *reinterpret_cast<uint4*>(g0 + 16 * c) = *reinterpret_cast<const uint4*>(s0 + 16 * c);
    if (first && pad) { if (lane >= pad && lane < 16 && lane < end) g0[lane] = s0[lane]; }
    if (last) { const int pos = 16 * cend + lane; if (lane < 16 && pos < end && !(c0 && cend == 0)) g0[pos] = s0[pos]; }
    else if (lane == 0) *carry = *reinterpret_cast<const uint4*>(s0 + 16 * cend);
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// read-only table word at shared address a + OFF.  Deliberately not volatile: the tables never change after the
// kernel's first barrier, so the compiler may schedule these loads freely.
template <int OFF>
__device__ __forceinline__ uint32_t lds_tab(uint32_t a)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
    return v;
}
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f)
{
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}
// Phase-B lane -> codeword map.  A codeword's table variant is (band offset + tile + row) mod 3; passes 0..2 take 32
// codewords of one variant each (all lanes then read the same 27-entry block of table A: no bank conflicts), pass 3
// the remaining 21.  One 128-byte map per tile mod 3; 255 = idle lane.
// Built by 3 x 128 threads (one per tile-mod-3 and slot candidate): a codeword's rank inside its variant's row-major list is
// a closed form of how many bands share each band offset mod 3.
__device__ __forceinline__ void build_pass_map(uint8_t* maps, const Geom& g, int t)
{
    if (t >= 3 * 128) return;
    const int tm = t >> 7, cw = t & 127;
    uint8_t* map = maps + 128 * tm;
    if (cw >= 9 * C_MINI) { map[cw] = 255; return; }
    uint32_t cwb[9];
    int nb[3] = {0, 0, 0}; // bands per offset class
    for (int b = 0; b < 9; ++b) { cwb[b] = (uint32_t)(g.cw_base[b] % 3); ++nb[cwb[b]]; }
    const int cl = cw / 9, b = cw - 9 * cl;
    const int v = (int)((cwb[b] + (uint32_t)tm + (uint32_t)cl) % 3);
    // variant of (band class x, row r) is (x + tm + r) % 3
    auto row_count = [&](int vv, int r) { return nb[((vv - tm - r) % 3 + 3) % 3]; };
    int rank = 0;
    for (int r = 0; r < cl; ++r) rank += row_count(v, r);
    for (int bb = 0; bb < b; ++bb) rank += ((int)((cwb[bb] + (uint32_t)tm + (uint32_t)cl) % 3) == v);
    if (rank < 32) { map[32 * v + rank] = (uint8_t)cw; return; }
    int off = 96; // leftovers of variants below v come first in the mixed pass
    for (int vv = 0; vv < v; ++vv) {
        int n = 0;
        for (int r = 0; r < C_MINI; ++r) n += row_count(vv, r);
        off += n - 32;
    }
    map[off + rank - 32] = (uint8_t)cw;
}
// the nine staged runs <-> global in whole 16-byte chunks, flattened over (band, chunk): 9 x 23 slots in 7 steps
constexpr int RUN_SLOTS = 23;
// contiguous share of [0, total) for warp gw of nw
__device__ __forceinline__ void warp_range(uint64_t total, uint32_t gw, uint32_t nw, uint32_t& lo, uint32_t& hi)
{
    lo = (uint32_t)(total * gw / nw);
    hi = (uint32_t)(total * (gw + 1) / nw);
}

// 26 stream symbols of one unit (x4, as produced by two triple_to_symbols) -> shared memory at dst (even address).  REV: units of an
// odd row of a 26-wide 2D tile leave in reverse order (boustrophedon, OLD:750-780), decided per lane by `rev`
template <bool REV>
__device__ __forceinline__ void store_unit26(uint8_t* dst, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t s12, uint32_t v0, uint32_t v1, uint32_t v2,
                                             uint32_t t12, bool rev)
{
    uint16_t* d = reinterpret_cast<uint16_t*>(dst); // 13 halfwords (STS.U16 keeps the low 16 bits)
    if constexpr (!REV) {
        d[0] = (uint16_t)w0; d[1] = (uint16_t)(w0 >> 16); d[2] = (uint16_t)w1; d[3] = (uint16_t)(w1 >> 16);
        d[4] = (uint16_t)w2; d[5] = (uint16_t)(w2 >> 16); d[6] = (uint16_t)(s12 | (v0 << 8));
        d[7] = (uint16_t)(v0 >> 8); d[8] = (uint16_t)__funnelshift_r(v0, v1, 24); d[9] = (uint16_t)(v1 >> 8);
        d[10] = (uint16_t)__funnelshift_r(v1, v2, 24); d[11] = (uint16_t)(v2 >> 8); d[12] = (uint16_t)((v2 >> 24) | (t12 << 8));
    } else {
        uint32_t x[7] = {w0, w1, w2, s12 | (v0 << 8), __funnelshift_r(v0, v1, 24), __funnelshift_r(v1, v2, 24), (v2 >> 24) | (t12 << 8)};
        uint32_t z[7];
#pragma unroll
        for (int j = 0; j < 6; ++j) z[j] = rev ? __byte_perm(x[5 - j], x[6 - j], 0x2345) : x[j]; // bytes 25-4j .. 22-4j
        z[6] = rev ? __byte_perm(x[0], 0u, 0x4401) : x[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) { d[2 * j] = (uint16_t)z[j]; d[2 * j + 1] = (uint16_t)(z[j] >> 16); }
        d[12] = (uint16_t)z[6];
    }
}
// the inverse on the way back: 26 symbols at S + a (even) as seven word-aligned registers (y[6]: two symbols)
template <bool REV>
__device__ __forceinline__ void load_unit26(const uint8_t* S, uint32_t a, uint32_t (&y)[7], bool rev)
{
    const uint32_t* mw = reinterpret_cast<const uint32_t*>(S + (a & ~3u));
    const uint32_t sh = (a & 2u) * 8u;
    uint32_t x[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) x[j] = mw[j];
#pragma unroll
    for (int j = 0; j < 6; ++j) y[j] = __funnelshift_r(x[j], x[j + 1], sh);
    y[6] = x[6] >> sh;
    if constexpr (REV) {
        uint32_t z[7];
#pragma unroll
        for (int j = 0; j < 6; ++j) z[j] = rev ? __byte_perm(y[5 - j], y[6 - j], 0x2345) : y[j];
        z[6] = rev ? __byte_perm(y[0], 0u, 0x4401) : y[6];
#pragma unroll
        for (int j = 0; j < 7; ++j) y[j] = z[j];
    }
}
// ---- encode phase A: six pixels (18 bytes at U + a) -> 26 stream symbols (x4) at d
template <bool ODD, bool REV = false>
__device__ __forceinline__ void enc_unit_rgb(const uint8_t* U, uint32_t a, uint8_t* dst, bool rev = false)
{
    const uint32_t* mw = reinterpret_cast<const uint32_t*>(U + (a & ~3u));
    const uint32_t sh = (a & 3u) * 8u;
    uint32_t x[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) x[j] = mw[j];
    uint32_t y[5]; // the 18 bytes, word aligned
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = __funnelshift_r(x[j], x[j + 1], sh);
    if constexpr (ODD) y[4] = __funnelshift_r(x[4], sh == 24 ? mw[5] : 0u, sh); // 18 bytes from byte offset 3 reach into a sixth word
    else y[4] = x[4] >> sh;
    uint32_t A[6];
#pragma unroll
    for (int p = 0; p < 6; ++p) {
        const int q = 3 * p;
        A[p] = rgb_to_value3(byte_magic(y[q >> 2], q & 3), byte_magic(y[(q + 1) >> 2], (q + 1) & 3), byte_magic(y[(q + 2) >> 2], (q + 2) & 3));
    }
    uint32_t w0, w1, w2, s12, v0, v1, v2, t12;
    triple_to_symbols(A[0], A[1], A[2], w0, w1, w2, s12);
    triple_to_symbols(A[3], A[4], A[5], v0, v1, v2, t12);
    w0 <<= 2; w1 <<= 2; w2 <<= 2; s12 <<= 2; v0 <<= 2; v1 <<= 2; v2 <<= 2; t12 <<= 2; // symbols <= 26: no carry between bytes
    store_unit26<REV>(dst, w0, w1, w2, s12, v0, v1, v2, t12, rev);
}
template <int K, bool ODD>
__device__ __forceinline__ void enc_phase_a_impl(const uint8_t* U, uint32_t pad, uint8_t* S, int lane)
{
    using L = Cfg3<K>;
    // units dealt even / odd so that the 18- and 26-byte lane strides become 36 and 52 bytes: 9 and 13 words, conflict-free
#pragma unroll 1
    for (int pass = 0; pass < L::PASS_A; ++pass) {
        const int u = pass < 2 ? 2 * lane + pass : 32 * pass + lane;
        if (u >= L::UNITS) continue;
        enc_unit_rgb<ODD>(U, pad + 18u * (uint32_t)u, S + 26 * u);
    }
}
// frames that start on an odd byte (odd pixel counts, several frames per call) put pixel units at byte offset 3 of a word
template <int K>
__device__ __forceinline__ void enc_phase_a(const uint8_t* U, uint32_t pad, uint8_t* S, int lane)
{
    if (pad & 1u) enc_phase_a_impl<K, true>(U, pad, S, lane); else enc_phase_a_impl<K, false>(U, pad, S, lane);
}
// ---- one codeword of encode phase B: K data symbols (x4) gathered at byte stride 9 from src -> 26 scrambled symbols at dst (even address).
// pa = shared address of the codeword's table variant {A[K][27] | B[K][27]}; pat = the scrambler on the parity symbols, plane domain
template <int K>
__device__ __forceinline__ void enc_cw(const uint8_t* src, uint8_t* dst, uint32_t pa, uint32_t pat_nz, uint32_t pat_two)
{
    constexpr int R = 26 - K, PLANE = 4 * K * 27;
    Planes acc{0, 0}, acc2{0, 0};
    uint32_t prev = 0, pk[K / 2];
    static_for<0, K>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        const uint32_t d4 = src[9 * i];
        const uint32_t ra = pa + d4;
        const uint32_t ea = lds_tab<108 * i>(ra), eb = lds_tab<108 * i + PLANE>(ra);
        if (i & 1) { if (i == 1) acc2 = Planes{ea, eb}; else gf3_add(acc2, ea, eb); pk[i / 2] = __byte_perm(prev, ea, 0x0040); }
        else { if (i == 0) acc = Planes{ea, eb}; else gf3_add(acc, ea, eb); prev = ea; }
    });
    gf3_add(acc, acc2.nz, acc2.two);
    gf3_add(acc, pat_nz, pat_two);
    uint32_t lo, hi;
    planes_to_parity<K>(acc.nz, acc.two, lo, hi);
#pragma unroll
    for (int j = 0; j < K / 2; ++j) *reinterpret_cast<uint16_t*>(dst + 2 * j) = (uint16_t)pk[j]; // stores after all loads: nothing to order
    *reinterpret_cast<uint16_t*>(dst + K) = (uint16_t)lo;
    if (R > 2) *reinterpret_cast<uint16_t*>(dst + K + 2) = (uint16_t)(lo >> 16);
    if (R > 4) *reinterpret_cast<uint16_t*>(dst + K + 4) = (uint16_t)hi;
    if (R > 6) *reinterpret_cast<uint16_t*>(dst + K + 6) = (uint16_t)(hi >> 16);
}
// ---- encode phase B: stream symbols -> nine staged runs (data scrambled through the table bytes, parity through the planes)
template <int K>
__device__ __forceinline__ void enc_phase_b(const uint8_t* S, uint8_t* U, const WarpMeta3& meta, const uint8_t* pmap, uint32_t tabA32,
                                            const uint32_t* pat, int lane)
{
    using L = Cfg3<K>;
    // ---- phase B: one codeword per lane (lane -> codeword through the variant-sorted pass map)
#pragma unroll 1
    for (int pass = 0; pass < L::PASS_B; ++pass) {
        const uint32_t cw = pmap[32 * pass + lane];
        if (cw == 255) continue;
        const uint32_t cl = (cw * 57u) >> 9, b = cw - 9u * cl;          // cw / 9 for cw < 128
        const uint32_t v = ((uint32_t)meta.vb[b] + cl) % 3u;
        uint32_t pa = tabA32 + v * L::ENC_VAR;
        asm volatile("" : "+r"(pa));                                   // keep the variant base in a register
        const uint8_t* src = S + cw + (9 * K - 9) * cl;                 // 9K*cl + b
        uint8_t* dst = U + L::RUN_PITCH * b + ((uint32_t)meta.run_lo[b] & 15u) + 26 * cl;
        enc_cw<K>(src, dst, pa, pat[2 * v], pat[2 * v + 1]);
    }
}
// ---- one codeword of decode phase B: 26 received symbols (x4) at src (even address) -> syndrome screen / slow path ->
// K descrambled data symbols scattered at byte stride 9 from dst.  pa = shared address of the variant block {A[26][32] | B[26][32]}
// (256-byte aligned), tab_v = the same block as a pointer, chk = the clean-codeword constant of the variant
// PRESCALED = false: src holds the symbols as they lie in the frame; they are scaled here, four at a time, by multiply-shifts on the
// full-rate pipe (bytes >= 32 -- out-of-alphabet symbols, OLD:28-31 -- are reduced mod 27 first; 27..31 alias 0..4 inside the tables)
template <int K, bool PRESCALED = true>
__device__ __forceinline__ void dec_cw(const uint8_t* src, uint8_t* dst, uint32_t pa, const uint8_t* tab_v, uint32_t chk_nz, uint32_t chk_two,
                                       const GfTables& sg, const uint32_t* chien, uint32_t* status, bool count = true)
{
    constexpr int PLANE = 4 * 26 * 32;
    // the codeword's 26 symbols as 7 words (it starts on an even byte), then one PRMT per symbol builds the
    // table address: byte 0 = symbol x4, bytes 1..3 = the variant block's address
    const uint32_t sa = smem_u32(src), sh = (sa & 2u) * 8u;      // explicit shared-window loads (an integer round trip of the pointer made these generic LD.E)
    uint32_t xw[7];
    static_for<0, 7>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(xw[j]) : "r"(sa & ~3u), "n"(4 * j) : "memory");
    });
#pragma unroll
    for (int j = 0; j < 6; ++j) xw[j] = __funnelshift_r(xw[j], xw[j + 1], sh);
    xw[6] >>= sh;
    if constexpr (!PRESCALED) {
        xw[6] &= 0xFFFFu;                                          // symbols 24, 25 only
        if ((xw[0] | xw[1] | xw[2] | xw[3] | xw[4] | xw[5] | xw[6]) & 0xE0E0E0E0u) {
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                uint32_t r = 0;
                for (int q = 0; q < 4; ++q) r |= (((xw[j] >> (8 * q)) & 0xFFu) % 27u) << (8 * q);
                xw[j] = r;
            }
        }
#pragma unroll
        for (int j = 0; j < 7; ++j) xw[j] *= 4u;                   // < 128 per byte: no carry between symbols
    }
    Planes acc{0, 0}, acc2{0, 0};
    uint32_t ev[K];
    static_for<0, 26>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        const uint32_t ra = __byte_perm(xw[i >> 2], pa, 0x7650u | (uint32_t)(i & 3));
        const uint32_t ea = lds_tab<128 * i>(ra);
        const uint32_t eb = lds_tab<128 * i + PLANE>(ra);
        if (i == 0) acc = Planes{ea, eb}; else if (i == 1) acc2 = Planes{ea, eb};   // 0 + x = x: the first entry of an accumulator is a move
        else if (i & 1) gf3_add(acc2, ea, eb); else gf3_add(acc, ea, eb);
        if (i < K) ev[i < K ? i : 0] = ea;
    });
#pragma unroll
    for (int i = 0; i < K; ++i) dst[9 * i] = (uint8_t)ev[i];       // stores after all loads: nothing to order
    gf3_add(acc, acc2.nz, acc2.two);
    if (((acc.nz ^ chk_nz) | (acc.two ^ chk_two)) & ~0xFFu) { // the low bytes carry the embedded symbols
        // slow path: the screen's sum minus the clean-codeword constant is the parity residual, from which the bounded-distance
        // decoder (dev.cuh) finds the error values and repairs the data symbols stored above
        Planes d{acc.nz, acc.two};
        gf3_add(d, chk_nz, chk_nz ^ chk_two);                      // minus the constant: -x keeps nz and flips two where nz is set
        uint32_t lo, hi;
        planes_to_parity<K>(d.nz, d.two, lo, hi);
        rs_bd_fix<K>(sg, chien, dst, lo, hi, status, count);
    }
}
// ---- decode phase B: nine staged runs (x4) -> syndrome screen / slow path -> descrambled stream symbols
template <int K, bool PRESCALED = true>
__device__ __forceinline__ void dec_phase_b(const uint8_t* U, uint8_t* S, const WarpMeta3& meta, const uint8_t* pmap, uint32_t tabA32,
                                            const uint8_t* tabA, const uint32_t* chk, const GfTables& sg, const uint32_t* chien, uint32_t* status, int lane)
{
    using L = Cfg3<K>;
    // ---- phase B: syndrome screen per codeword (variant-sorted lanes); descrambled data symbols -> stream order
#pragma unroll 1
    for (int pass = 0; pass < L::PASS_B; ++pass) {
        const uint32_t cw = pmap[32 * pass + lane];
        if (cw == 255) continue;
        const uint32_t cl = (cw * 57u) >> 9, b = cw - 9u * cl;          // cw / 9 for cw < 128
        const uint32_t v = ((uint32_t)meta.vb[b] + cl) % 3u;
        uint32_t pa = tabA32 + v * L::DEC_VAR;                         // low byte 0
        asm volatile("" : "+r"(pa));
        const uint8_t* src = U + L::RUN_PITCH * b + ((uint32_t)meta.run_lo[b] & 15u) + 26 * cl;
        uint8_t* dst = S + cw + (9 * K - 9) * cl;                       // 9K*cl + b
        dec_cw<K, PRESCALED>(src, dst, pa, tabA + v * L::DEC_VAR, chk[2 * v], chk[2 * v + 1], sg, chien, status);
    }
}
// ---- decode phase A: 26 stream symbols at src (even address) -> six pixels -> 18 RGB bytes at dst (even address)
template <bool REV = false>
__device__ __forceinline__ void dec_unit_rgb(const uint8_t* S, uint32_t a, uint8_t* dst, bool rev = false)
{
    uint32_t y[7]; // the 26 symbols, word aligned
    load_unit26<REV>(S, a, y, rev);
    uint32_t A[6];
    symbols_to_triple(y[0], y[1], y[2], y[3] & 0xFF, A[0], A[1], A[2]);
    symbols_to_triple(__funnelshift_r(y[3], y[4], 8), __funnelshift_r(y[4], y[5], 8), __funnelshift_r(y[5], y[6], 8), (y[6] >> 8) & 0xFF, A[3], A[4], A[5]);
    uint32_t p[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) p[q] = value_to_rgb3(A[q]);
    uint16_t* dh = reinterpret_cast<uint16_t*>(dst);
#pragma unroll
    for (int q = 0; q < 3; ++q) { // two pixels = three halfwords
        dh[3 * q] = (uint16_t)p[2 * q];
        dh[3 * q + 1] = (uint16_t)((p[2 * q] >> 16) | (p[2 * q + 1] << 8));
        dh[3 * q + 2] = (uint16_t)(p[2 * q + 1] >> 8);
    }
}
template <int K>
__device__ __forceinline__ void dec_phase_a(const uint8_t* S, uint8_t* U, uint32_t pad, int lane)
{
    using L = Cfg3<K>;
#pragma unroll 1
    for (int pass = 0; pass < L::PASS_A; ++pass) {
        const int u = pass < 2 ? 2 * lane + pass : 32 * pass + lane;
        if (u >= L::UNITS) continue;
        dec_unit_rgb(S, 26u * (uint32_t)u, U + pad + 18 * u); // pad is even: every frame starts on a 16-byte boundary and 3*PX*tile is even
    }
}

// ---- raw-word front end (encode_profile_from_raw's own input, OLD:1051-1082): three Word27 (27 bytes at IN + pad + 27u) ->
// six 13-trit pixel values -> 26 stream symbols (x4) at S + 26u.  The regroup keeps the first 26 trits of every word, which
// are exactly the two 13-trit halves; bytes >= 27 read as their low three trits (unpack3, OLD:28-31).
template <bool REV = false>
__device__ __forceinline__ void enc_unit_words(const uint8_t* U, uint32_t a, uint8_t* dst, bool rev = false)
{
    const uint32_t* mw = reinterpret_cast<const uint32_t*>(U + (a & ~3u));
    const uint32_t sh = (a & 3u) * 8u;
    uint32_t x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = mw[j];
    uint32_t y[7]; // the 27 bytes, word aligned
#pragma unroll
    for (int j = 0; j < 7; ++j) y[j] = __funnelshift_r(x[j], x[j + 1], sh);
    y[6] &= 0x00FFFFFFu;
    uint32_t wild = 0;
#pragma unroll
    for (int j = 0; j < 7; ++j) wild |= ((y[j] & 0x7F7F7F7Fu) + 0x65656565u) | y[j]; // bit 7 of a byte set <=> byte >= 27
    if (wild & 0x80808080u) {
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            uint32_t r = 0;
            for (int q = 0; q < 4; ++q) r |= (((y[j] >> (8 * q)) & 0xFFu) % 27u) << (8 * q);
            y[j] = r;
        }
    }
    auto val4 = [](uint32_t w) { return __dp4a(w, 0x00001B01u, 0u) + 729u * __dp4a(w, 0x1B010000u, 0u); };
    uint32_t A[6];
#pragma unroll
    for (int w = 0; w < 3; ++w) { // word w = bytes 9w .. 9w+8
        const int o = 9 * w;
        auto word_at = [&](int b) { return (b & 3) ? __funnelshift_r(y[b >> 2], y[(b >> 2) + 1 > 6 ? 6 : (b >> 2) + 1], 8 * (b & 3)) : y[b >> 2]; };
        const uint32_t lo4 = word_at(o), s4 = word_at(o + 4) & 0xFFu;
        uint32_t hi4 = word_at(o + 5);
        if (w == 2) hi4 = (y[5] >> 24) | (y[6] << 8);                   // bytes 23..26 (word_at would read past y[6])
        const uint32_t s8 = hi4 >> 24, q8 = __umulhi(s8, 477218589u);      // the 27th trit is dropped: s8 % 9
        hi4 = (hi4 & 0x00FFFFFFu) | ((s8 - 9u * q8) << 24);
        const uint32_t q4 = __umulhi(s4, 1431655766u);                     // s4 / 3
        A[2 * w] = val4(lo4) + 531441u * (s4 - 3u * q4);
        A[2 * w + 1] = q4 + 9u * val4(hi4);
    }
    uint32_t w0, w1, w2, s12, v0, v1, v2, t12;
    triple_to_symbols(A[0], A[1], A[2], w0, w1, w2, s12);
    triple_to_symbols(A[3], A[4], A[5], v0, v1, v2, t12);
    w0 <<= 2; w1 <<= 2; w2 <<= 2; s12 <<= 2; v0 <<= 2; v1 <<= 2; v2 <<= 2; t12 <<= 2;
    store_unit26<REV>(dst, w0, w1, w2, s12, v0, v1, v2, t12, rev);
}
template <int K>
__device__ __forceinline__ void enc_phase_a_words(const uint8_t* U, uint32_t pad, uint8_t* S, int lane)
{
    using L = Cfg3<K>;
#pragma unroll 1
    for (int pass = 0; pass < L::PASS_A; ++pass) {
        const int u = pass < 2 ? 2 * lane + pass : 32 * pass + lane;
        if (u >= L::UNITS) continue;
        enc_unit_words(U, pad + 27u * (uint32_t)u, S + 26 * u);
    }
}
// ---- raw-word back end (the regroup of decode_profile_to_raw, OLD:1022-1039): 26 stream symbols at S + a -> six pixel
// values -> three Word27 (27 bytes, T[26] = 0) at d
template <bool REV = false>
__device__ __forceinline__ void dec_unit_words(const uint8_t* S, uint32_t a, uint8_t* d, bool rev = false)
{
    uint32_t y[7];
    load_unit26<REV>(S, a, y, rev);
    uint32_t A[6];
    symbols_to_triple(y[0], y[1], y[2], y[3] & 0xFF, A[0], A[1], A[2]);
    symbols_to_triple(__funnelshift_r(y[3], y[4], 8), __funnelshift_r(y[4], y[5], 8), __funnelshift_r(y[5], y[6], 8), (y[6] >> 8) & 0xFF, A[3], A[4], A[5]);
#pragma unroll
    for (int w = 0; w < 3; ++w) { // pack_two_pixels on values: s0..s3 | trit 12 + 3 (Ab % 9) | digits of Ab / 9
        uint32_t t, q;
        const uint32_t lo4 = digits4(A[2 * w], t);
        const uint32_t h = __umulhi(A[2 * w + 1], 477218589u), s4 = t + 3u * (A[2 * w + 1] - 9u * h);
        const uint32_t hi4 = digits4(h, q);
        uint8_t* o = d + 9 * w; // arbitrary alignment: bytes
        o[0] = (uint8_t)lo4; o[1] = (uint8_t)(lo4 >> 8); o[2] = (uint8_t)(lo4 >> 16); o[3] = (uint8_t)(lo4 >> 24); o[4] = (uint8_t)s4;
        o[5] = (uint8_t)hi4; o[6] = (uint8_t)(hi4 >> 8); o[7] = (uint8_t)(hi4 >> 16); o[8] = (uint8_t)(hi4 >> 24);
    }
}
template <int K>
__device__ __forceinline__ void dec_phase_a_words(const uint8_t* S, uint8_t* U, uint32_t pad, int lane)
{
    using L = Cfg3<K>;
#pragma unroll 1
    for (int pass = 0; pass < L::PASS_A; ++pass) {
        const int u = pass < 2 ? 2 * lane + pass : 32 * pass + lane;
        if (u >= L::UNITS) continue;
        dec_unit_words(S, 26u * (uint32_t)u, U + pad + 27 * u);
    }
}

template <int K>
__global__ void __launch_bounds__(FAST_TPB, 4) k_encode_rgb_v3(FastParams P, Geom g, const GfTables* __restrict__ gf, const RsTables* __restrict__ rs)
{
    using L = Cfg3<K>;
    extern __shared__ __align__(16) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* S = smem + L::ENC_WARP + warp * L::WARP_BYTES;     // stream symbols, pre-scaled by 4
    uint8_t* U = S + L::S_BYTES;                                // RGB run, later the nine body runs
    WarpMeta3& meta = *reinterpret_cast<WarpMeta3*>(U + L::U_BYTES);
    {
        const uint32_t(*pl)[kVals][2] = rs->pl[g.arith][(24 - K) / 2];
        for (int idx = tid; idx < 3 * K * 27; idx += FAST_TPB) {
            const int v = idx / (K * 27), rem = idx - v * (K * 27), i = rem / 27, d = rem - 27 * i;
            uint32_t* blk = reinterpret_cast<uint32_t*>(smem + L::ENC_A + v * L::ENC_VAR);
            blk[rem] = pl[i][d][0] | gf->scr[st_of(g, v, i)][d];
            blk[K * 27 + rem] = pl[i][d][1];
        }
        if (tid < 3) { // the scrambler as seen by the parity symbols of a variant-tid codeword, in the plane domain
            uint32_t nz = 0, two = 0;
            for (int j = 0; j < L::R; ++j) {
                const uint32_t st = st_of(g, tid, K + j);
                if (st) nz |= 7u << plane_shift<K>(j);
                if (st == 2) two |= 7u << plane_shift<K>(j);
            }
            reinterpret_cast<uint32_t*>(smem + L::ENC_PAT)[2 * tid] = nz;
            reinterpret_cast<uint32_t*>(smem + L::ENC_PAT)[2 * tid + 1] = two;
        }
        for (int t = tid; t < 3 * 128; t += (int)blockDim.x) build_pass_map(smem + L::ENC_MAP, g, t);
    }
    __syncthreads(); // the only block-level barrier: tables staged
    const uint32_t tabA32 = smem_u32(smem + L::ENC_A);
    const uint32_t* pat = reinterpret_cast<const uint32_t*>(smem + L::ENC_PAT);
    uint4* carry = reinterpret_cast<uint4*>(U + L::U_BYTES + L::META_BYTES);
    uint32_t mt_lo, mt_hi;
    warp_range((uint64_t)P.n_tiles * P.n_frames, blockIdx.x * FAST_WARPS + warp, gridDim.x * FAST_WARPS, mt_lo, mt_hi);
    for (uint32_t mt = mt_lo; mt < mt_hi; ++mt) {
        const uint32_t f = mt / P.n_tiles, tile = P.tile0 + (mt - f * P.n_tiles);
        const bool first = mt == mt_lo || tile == P.tile0, last = mt + 1 == mt_hi || tile + 1 == P.tile0 + P.n_tiles; // of a contiguous stretch
        const uint64_t g_lo = P.in_stride * f + 3ull * L::PX * tile;
        const uint32_t pad = (uint32_t)(g_lo & 15);
        setup_runs3(meta, g, P.out_stride * f, tile, lane);
        warp_load_run(U, P.in, g_lo, g_lo + L::RGB_BYTES, P.in_stride * P.n_frames, lane);
        if (!last && lane < (L::RGB_BYTES + 127) / 128 && g_lo + L::RGB_BYTES + 128 * lane < P.in_stride * P.n_frames) prefetch_l2(P.in + g_lo + L::RGB_BYTES + 128 * lane); // the next tile's pixels
        __syncwarp();
        enc_phase_a<K>(U, pad, S, lane);
        __syncwarp();
        if (!first && lane < 9) *reinterpret_cast<uint4*>(U + L::RUN_PITCH * lane) = carry[lane]; // bytes [0, pad) of each run: the previous tile's tail
        __syncwarp();
        enc_phase_b<K>(S, U, meta, smem + L::ENC_MAP + 128 * (tile % 3u), tabA32, pat, lane);
        __syncwarp();
        if (tile == 0 && lane == 0 && g.cw_base[0] == 0) { // body symbols 0 and 1 may still see the scrambler's transient (A.4)
            uint8_t* dst = U + ((uint32_t)meta.run_lo[0] & 15u);
            dst[0] = gf->scr[g.st[0]][S[0] >> 2];
            dst[1] = gf->scr[g.st[1]][S[9] >> 2];
        }
        __syncwarp();
        // ---- phase C: nine band-major runs -> global, whole chunks (see stream_store for the carry scheme)
#pragma unroll 1
        for (int it = 0; it < (9 * RUN_SLOTS + 31) / 32; ++it) {
            const uint32_t sl = 32u * it + lane, b = (sl * 2850u) >> 16, c = sl - RUN_SLOTS * b;   // sl / 23 for sl < 256
            if (b < 9) {
                const uint64_t lo = meta.run_lo[b];
                const uint32_t padb = (uint32_t)lo & 15u, cend = (padb + L::RUN) >> 4;
                if (c < cend && !(first && padb && c == 0))
                    *reinterpret_cast<uint4*>(P.out + (lo - padb) + 16 * c) = *reinterpret_cast<const uint4*>(U + 16 * sl);
            }
        }
        if (!last) {
            if (lane < 9) carry[lane] = *reinterpret_cast<const uint4*>(U + L::RUN_PITCH * lane + 16 * ((((uint32_t)meta.run_lo[lane] & 15u) + L::RUN) >> 4));
        }
        if (first || last) { // edge bytes of a contiguous stretch, one by one
#pragma unroll 1
            for (int b = 0; b < 9; ++b) {
                const uint64_t lo = meta.run_lo[b];
                const int padb = (int)(lo & 15), end = padb + L::RUN, cend = end >> 4;
                const uint8_t* s0 = U + L::RUN_PITCH * b;
                uint8_t* g0 = P.out + (lo - padb);
                if (first && padb && lane >= padb && lane < 16) g0[lane] = s0[lane];
                if (last && lane < 16 && 16 * cend + lane < end) g0[16 * cend + lane] = s0[16 * cend + lane];
            }
        }
        __syncwarp();
    }
}

template <int K>
__global__ void __launch_bounds__(FAST_TPB, 3) k_decode_rgb_v3(FastParams P, Geom g, const GfTables* __restrict__ gf, const RsTables* __restrict__ rs)
{
    using L = Cfg3<K>;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((256u - (smem_u32(smem_raw) & 255u)) & 255u); // 256-byte aligned (see Cfg3)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* S = smem + L::DEC_WARP + warp * L::WARP_BYTES;     // descrambled stream symbols (plain)
    uint8_t* U = S + L::S_BYTES;                                // the nine body runs (x4), later the RGB run
    WarpMeta3& meta = *reinterpret_cast<WarpMeta3*>(U + L::U_BYTES);
    GfTables& sg = *reinterpret_cast<GfTables*>(smem + L::DEC_GF);
    {
        const uint32_t(*pl)[kVals][2] = rs->pl[1][(24 - K) / 2]; // the consistent decoder always uses the repaired code
        for (int idx = tid; idx < 3 * 26 * 32; idx += FAST_TPB) {
            const int v = idx / (26 * 32), rem = idx - v * (26 * 32), i = rem / 32, x = rem - 32 * i, xm = x >= 27 ? x - 27 : x;
            uint32_t* blk = reinterpret_cast<uint32_t*>(smem + L::DEC_A + v * L::DEC_VAR);
            blk[rem] = pl[i][xm][0] | gf->dsc[st_of(g, v, i)][xm];
            blk[26 * 32 + rem] = pl[i][xm][1];
        }
        load_gf(sg, gf);
        for (int t = tid; t < 3 * 128; t += (int)blockDim.x) build_pass_map(smem + L::DEC_MAP, g, t);
    }
    __syncthreads();
    if (tid < 3) { // a received block r = c (+) 13*st is a codeword iff sum_i T_i[r_i] == sum_i T_i[13*st_i] (GF(3)-linear tables)
        Planes c{0, 0};
        const uint32_t* blk = reinterpret_cast<const uint32_t*>(smem + L::DEC_A + tid * L::DEC_VAR);
        for (int i = 0; i < 26; ++i) {
            const int idx = i * 32 + 13 * (int)st_of(g, tid, i);
            gf3_add(c, blk[idx] & ~0xFFu, blk[26 * 32 + idx]);
        }
        reinterpret_cast<uint32_t*>(smem + L::DEC_CHK)[2 * tid] = c.nz;
        reinterpret_cast<uint32_t*>(smem + L::DEC_CHK)[2 * tid + 1] = c.two;
    }
    __syncthreads();
    const uint32_t tabA32 = smem_u32(smem + L::DEC_A);
    const uint32_t* chk = reinterpret_cast<const uint32_t*>(smem + L::DEC_CHK);
    const uint64_t in_limit = P.in_stride * (P.n_frames - 1) + 9 * g.n_out;
    uint4* carry = reinterpret_cast<uint4*>(U + L::U_BYTES + L::META_BYTES);
    uint32_t mt_lo, mt_hi;
    warp_range((uint64_t)P.n_tiles * P.n_frames, blockIdx.x * FAST_WARPS + warp, gridDim.x * FAST_WARPS, mt_lo, mt_hi);
    for (uint32_t mt = mt_lo; mt < mt_hi; ++mt) {
        const uint32_t f = mt / P.n_tiles, tile = P.tile0 + (mt - f * P.n_tiles);
        const bool first = mt == mt_lo || tile == P.tile0, last = mt + 1 == mt_hi || tile + 1 == P.tile0 + P.n_tiles; // of a contiguous stretch
        setup_runs3(meta, g, P.in_stride * f, tile, lane);
        __syncwarp();
        if (!last) { // the next tile's nine runs: 3 lines each, towards L2
            const int b = lane / 3;
            if (b < 9 && meta.run_lo[b] + L::RUN + 128 * (lane - 3 * b) < in_limit) prefetch_l2(P.in + meta.run_lo[b] + L::RUN + 128 * (lane - 3 * b));
        }
        // ---- nine runs -> shared, every byte scaled by 4 (table byte offset).  Bytes >= 32 would leave the 32-entry
        // rows: they are reduced mod 27 first (out-of-alphabet symbols read as their low three trits, OLD:28-31)
        bool wild = false;
#pragma unroll 1
        for (int it = 0; it < (9 * RUN_SLOTS + 31) / 32; ++it) {
            const uint32_t sl = 32u * it + lane, b = (sl * 2850u) >> 16, c = sl - RUN_SLOTS * b;   // sl / 23 for sl < 256
            if (b < 9) {
                const uint64_t lo = meta.run_lo[b];
                const uint32_t padb = (uint32_t)lo & 15u;
                const uint64_t ga = (lo - padb) + 16 * c;
                if (16 * c < padb + L::RUN) {
                    uint4 q;
                    if (ga + 16 <= in_limit) q = __ldg(reinterpret_cast<const uint4*>(P.in + ga));
                    else { // the last chunk of the last frame may poke past the buffer
                        uint32_t t[4] = {0, 0, 0, 0};
                        for (int i = 0; i < 16; ++i) if (ga + i < in_limit) t[i >> 2] |= (uint32_t)P.in[ga + i] << (8 * (i & 3));
                        q = make_uint4(t[0], t[1], t[2], t[3]);
                    }
                    wild |= ((q.x | q.y | q.z | q.w) & 0xE0E0E0E0u) != 0;
                    q.x <<= 2; q.y <<= 2; q.z <<= 2; q.w <<= 2;
                    *reinterpret_cast<uint4*>(U + 16 * sl) = q;
                }
            }
        }
        if (__any_sync(0xFFFFFFFFu, wild)) { // rare: reload those chunks byte by byte
#pragma unroll 1
            for (int b = 0; b < 9; ++b) {
                const uint64_t lo = meta.run_lo[b];
                const int padb = (int)(lo & 15), c16 = 16 * lane;
                if (c16 < padb + L::RUN)
                    for (int i = 0; i < 16; ++i) {
                        const uint64_t ga = (lo - padb) + c16 + i;
                        U[L::RUN_PITCH * b + c16 + i] = (uint8_t)(ga < in_limit ? 4u * (P.in[ga] % 27u) : 0u);
                    }
            }
        }
        __syncwarp();
        if (tile == 0 && lane == 0 && g.cw_base[0] == 0) { // body symbols 0,1: move them from the transient states to the periodic ones
            uint8_t* r0 = U + ((uint32_t)meta.run_lo[0] & 15u);
            r0[0] = (uint8_t)(4u * sg.scr[st_of(g, 0, 0)][sg.dsc[g.st[0]][(r0[0] >> 2) % 27u]]);
            r0[1] = (uint8_t)(4u * sg.scr[st_of(g, 0, 1)][sg.dsc[g.st[1]][(r0[1] >> 2) % 27u]]);
        }
        __syncwarp();
        dec_phase_b<K>(U, S, meta, smem + L::DEC_MAP + 128 * (tile % 3u), tabA32, smem + L::DEC_A, chk, sg, chien_of(gf), P.status + 2 * f, lane);
        __syncwarp();
        // ---- phase A: 26 stream symbols -> six pixels -> 18 RGB bytes per lane (units dealt even / odd)
        const uint64_t g_lo = P.out_stride * f + 3ull * L::PX * tile;
        const uint32_t pad = (uint32_t)(g_lo & 15);
        if (!first && lane == 0) *reinterpret_cast<uint4*>(U) = carry[0]; // bytes [0, pad): the previous tile's tail
        __syncwarp();
        dec_phase_a<K>(S, U, pad, lane);
        __syncwarp();
        stream_store(U, P.out, g_lo, L::RGB_BYTES, first, last, carry, lane);
        __syncwarp();
    }
}

// =============================================================================================
// v4: the v3 phases with all tile I/O on the bulk-async copy engine (cp.async.bulk, SASS UBLKCP) and one CTA per SM.
//   * the next tile's input is fetched into its own shared buffer by cp.async.bulk + mbarrier while the current tile is
//     processed (the v3 profile showed 12 % of warp time waiting on the tile's first global load);
//   * finished runs leave shared memory as bulk stores of whole 16-byte chunks (one instruction per run, issued by one
//     lane per band) instead of LDS.128/STG.128 loops; the partial last chunk is carried to the next tile as in v3;
//   * 28 (encode) / 27 (decode) warps share one copy of the tables.
// =============================================================================================
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
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
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// v4 tile ranges: CTAs get equal contiguous shares; inside a CTA the four SM sub-partitions (warp % 4) get equal shares
// (a sub-partition's time is the sum of its warps' work, so per-warp rounding must not pile up on one of them), and
// the warps of a sub-partition split its share.
__device__ __forceinline__ void warp_range_smsp(uint64_t total, uint32_t cta, uint32_t n_cta, uint32_t warp, uint32_t n_warps, uint32_t& lo, uint32_t& hi)
{
    const uint64_t c_lo = total * cta / n_cta, c_hi = total * (cta + 1) / n_cta;
    const uint32_t q = warp & 3u, j = warp >> 2, nq = (n_warps - q + 3u) >> 2; // warps on sub-partition q
    const uint64_t q_lo = c_lo + (c_hi - c_lo) * q / 4, q_hi = c_lo + (c_hi - c_lo) * (q + 1) / 4;
    lo = (uint32_t)(q_lo + (q_hi - q_lo) * j / nq);
    hi = (uint32_t)(q_lo + (q_hi - q_lo) * (j + 1) / nq);
}

// WORDS: the pixel side of the tile is raw Word27 words (9 bytes per two pixels) instead of RGB8
template <int K, bool WORDS = false> struct Cfg4 {
    using L = Cfg3<K>;
    static constexpr int PIX_BYTES = WORDS ? 9 * (L::PX / 2) : L::RGB_BYTES; // pixel-side bytes of one mini-tile
    static constexpr int IN_BYTES = (PIX_BYTES + 15 + 15) / 16 * 16 + (WORDS ? 16 : 0); // with alignment slack (+ the word front end reads 32 bytes per unit)
    static constexpr int RUNS_BYTES = 9 * L::RUN_PITCH;
    static constexpr int TAIL_BYTES = L::META_BYTES + L::CARRY_BYTES + 16;   // meta | carry | mbarrier
    static constexpr int WARP_BYTES = IN_BYTES + L::S_BYTES + RUNS_BYTES + TAIL_BYTES; // encode: IN | S | U(runs);  decode: OUT | S | R(runs)
    static constexpr int SMEM_MAX = 227 * 1024;
    static constexpr int ENC_WARPS = (SMEM_MAX - L::ENC_WARP) / WARP_BYTES < 32 ? (SMEM_MAX - L::ENC_WARP) / WARP_BYTES : 32; // K=20: 28
    static constexpr int DEC_WARPS = (SMEM_MAX - 256 - L::DEC_WARP) / WARP_BYTES < 32 ? (SMEM_MAX - 256 - L::DEC_WARP) / WARP_BYTES : 32; // K=20: 27
    static constexpr int TOTAL_ENC = L::ENC_WARP + ENC_WARPS * WARP_BYTES;
    static constexpr int TOTAL_DEC = 256 + L::DEC_WARP + DEC_WARPS * WARP_BYTES;
};

template <int K, bool WORDS>
__global__ void __launch_bounds__(32 * Cfg4<K, WORDS>::ENC_WARPS, 1) k_encode_rgb_v4(FastParams P, Geom g, const GfTables* __restrict__ gf, const RsTables* __restrict__ rs)
{
    using L = Cfg3<K>;
    using L4 = Cfg4<K, WORDS>;
    constexpr int PIX = L4::PIX_BYTES;
    constexpr int V4_ENC_WARPS = L4::ENC_WARPS, TPB = 32 * V4_ENC_WARPS;
    extern __shared__ __align__(16) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* IN = smem + L::ENC_WARP + warp * L4::WARP_BYTES;   // the RGB tile (bulk-loaded one tile ahead)
    uint8_t* S = IN + L4::IN_BYTES;                             // stream symbols, pre-scaled by 4
    uint8_t* U = S + L::S_BYTES;                                // the nine body runs
    WarpMeta3& meta = *reinterpret_cast<WarpMeta3*>(U + L4::RUNS_BYTES);
    uint4* carry = reinterpret_cast<uint4*>(U + L4::RUNS_BYTES + L::META_BYTES);
    const uint32_t bar = smem_u32(U + L4::RUNS_BYTES + L::META_BYTES + L::CARRY_BYTES);
    {
        const uint32_t(*pl)[kVals][2] = rs->pl[g.arith][(24 - K) / 2];
        for (int idx = tid; idx < 3 * K * 27; idx += TPB) {
            const int v = idx / (K * 27), rem = idx - v * (K * 27), i = rem / 27, d = rem - 27 * i;
            uint32_t* blk = reinterpret_cast<uint32_t*>(smem + L::ENC_A + v * L::ENC_VAR);
            blk[rem] = pl[i][d][0] | gf->scr[st_of(g, v, i)][d];
            blk[K * 27 + rem] = pl[i][d][1];
        }
        if (tid < 3) {
            uint32_t nz = 0, two = 0;
            for (int j = 0; j < L::R; ++j) {
                const uint32_t st = st_of(g, tid, K + j);
                if (st) nz |= 7u << plane_shift<K>(j);
                if (st == 2) two |= 7u << plane_shift<K>(j);
            }
            reinterpret_cast<uint32_t*>(smem + L::ENC_PAT)[2 * tid] = nz;
            reinterpret_cast<uint32_t*>(smem + L::ENC_PAT)[2 * tid + 1] = two;
        }
        for (int t = tid; t < 3 * 128; t += (int)blockDim.x) build_pass_map(smem + L::ENC_MAP, g, t);
        if (lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    }
    __syncthreads(); // the only block-level barrier: tables and barriers ready
    const uint32_t tabA32 = smem_u32(smem + L::ENC_A);
    const uint32_t* pat = reinterpret_cast<const uint32_t*>(smem + L::ENC_PAT);
    const uint64_t in_limit = P.in_stride * P.n_frames;
    uint32_t mt_lo, mt_hi;
    warp_range_smsp((uint64_t)P.n_tiles * P.n_frames, blockIdx.x, gridDim.x, warp, V4_ENC_WARPS, mt_lo, mt_hi);
    // bulk load of tile mt's pixels: the 16-byte aligned superset of [g_lo, g_lo + RGB_BYTES), clipped to the buffer
    auto fetch = [&](uint32_t mt) {
        const uint32_t f = mt / P.n_tiles, tile = P.tile0 + (mt - f * P.n_tiles);
        const uint64_t g_lo = P.in_stride * f + (uint64_t)PIX * tile, a0 = g_lo & ~15ull;
        uint32_t bytes = (uint32_t)((g_lo - a0) + PIX + 15) & ~15u;
        if (a0 + bytes > in_limit) bytes = (uint32_t)(in_limit - a0) & ~15u;
        fence_async_smem();
        mbar_expect_tx(bar, bytes);
        bulk_g2s(smem_u32(IN), P.in + a0, bytes, bar);
    };
    if (mt_lo < mt_hi && lane == 0) fetch(mt_lo);
    uint32_t phase = 0;
    for (uint32_t mt = mt_lo; mt < mt_hi; ++mt) {
        const uint32_t f = mt / P.n_tiles, tile = P.tile0 + (mt - f * P.n_tiles);
        const bool first = mt == mt_lo || tile == P.tile0, last = mt + 1 == mt_hi || tile + 1 == P.tile0 + P.n_tiles; // of a contiguous stretch
        const uint64_t g_lo = P.in_stride * f + (uint64_t)PIX * tile;
        const uint32_t pad = (uint32_t)(g_lo & 15);
        setup_runs3(meta, g, P.out_stride * f, tile, lane);
        mbar_wait(bar, phase);
        phase ^= 1;
        {   // the last bytes of the buffer that a clipped bulk copy left out (only the final tile of the final frame)
            const uint64_t a0 = g_lo - pad;
            const uint32_t want = pad + PIX;
            if (a0 + ((want + 15) & ~15u) > in_limit)
                for (uint32_t i = ((uint32_t)(in_limit - a0) & ~15u) + lane; i < want; i += 32) IN[i] = a0 + i < in_limit ? P.in[a0 + i] : 0;
        }
        __syncwarp();
        if constexpr (WORDS) enc_phase_a_words<K>(IN, pad, S, lane); else enc_phase_a<K>(IN, pad, S, lane);
        __syncwarp();
        if (mt + 1 < mt_hi && lane == 0) fetch(mt + 1);                      // IN is free again: next tile's pixels on their way
        if (lane < 9) {
            bulk_wait_read();                                                // the previous tile's bulk stores have read U
            if (!first) *reinterpret_cast<uint4*>(U + L::RUN_PITCH * lane) = carry[lane]; // bytes [0, pad) of each run: the previous tile's tail
        }
        __syncwarp();
        enc_phase_b<K>(S, U, meta, smem + L::ENC_MAP + 128 * (tile % 3u), tabA32, pat, lane);
        __syncwarp();
        if (tile == 0 && lane == 0 && g.cw_base[0] == 0) { // body symbols 0 and 1 may still see the scrambler's transient (A.4)
            uint8_t* dst = U + ((uint32_t)meta.run_lo[0] & 15u);
            dst[0] = gf->scr[g.st[0]][S[0] >> 2];
            dst[1] = gf->scr[g.st[1]][S[9] >> 2];
        }
        fence_async_smem();                                                  // generic-proxy writes to U before the bulk engine reads it
        __syncwarp();
        // ---- phase C: nine band-major runs -> global as bulk stores of whole chunks (carry scheme of stream_store)
        if (lane < 9) {
            const uint64_t lo = meta.run_lo[lane];
            const uint32_t padb = (uint32_t)lo & 15u, cend = (padb + L::RUN) >> 4, c0 = (first && padb) ? 1u : 0u;
            if (!last) carry[lane] = *reinterpret_cast<const uint4*>(U + L::RUN_PITCH * lane + 16 * cend);
            if (cend > c0) bulk_s2g(P.out + (lo - padb) + 16 * c0, smem_u32(U + L::RUN_PITCH * lane + 16 * c0), 16 * (cend - c0));
            bulk_commit();
        }
        if (first || last) { // edge bytes of a contiguous stretch, one by one
#pragma unroll 1
            for (int b = 0; b < 9; ++b) {
                const uint64_t lo = meta.run_lo[b];
                const int padb = (int)(lo & 15), end = padb + L::RUN, cend = end >> 4;
                const uint8_t* s0 = U + L::RUN_PITCH * b;
                uint8_t* g0 = P.out + (lo - padb);
                if (first && padb && lane >= padb && lane < 16) g0[lane] = s0[lane];
                if (last && lane < 16 && 16 * cend + lane < end) g0[16 * cend + lane] = s0[16 * cend + lane];
            }
        }
        __syncwarp();
    }
    if (lane < 9) bulk_wait_all(); // shared memory must outlive the copies that read it
}

template <int K, bool WORDS>
__global__ void __launch_bounds__(32 * Cfg4<K, WORDS>::DEC_WARPS, 1) k_decode_rgb_v4(FastParams P, Geom g, const GfTables* __restrict__ gf, const RsTables* __restrict__ rs)
{
    using L = Cfg3<K>;
    using L4 = Cfg4<K, WORDS>;
    constexpr int PIX = L4::PIX_BYTES;
    constexpr int V4_DEC_WARPS = L4::DEC_WARPS, TPB = 32 * V4_DEC_WARPS;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((256u - (smem_u32(smem_raw) & 255u)) & 255u); // 256-byte aligned (see Cfg3)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* OUT = smem + L::DEC_WARP + warp * L4::WARP_BYTES;  // the RGB tile on its way out
    uint8_t* S = OUT + L4::IN_BYTES;                            // descrambled stream symbols (plain)
    uint8_t* R = S + L::S_BYTES;                                // the nine body runs (bulk-loaded one tile ahead, then x4 in place)
    WarpMeta3& meta = *reinterpret_cast<WarpMeta3*>(R + L4::RUNS_BYTES);
    uint4* carry = reinterpret_cast<uint4*>(R + L4::RUNS_BYTES + L::META_BYTES);
    const uint32_t bar = smem_u32(R + L4::RUNS_BYTES + L::META_BYTES + L::CARRY_BYTES);
    GfTables& sg = *reinterpret_cast<GfTables*>(smem + L::DEC_GF);
    {
        const uint32_t(*pl)[kVals][2] = rs->pl[1][(24 - K) / 2]; // the consistent decoder always uses the repaired code
        for (int idx = tid; idx < 3 * 26 * 32; idx += TPB) {
            const int v = idx / (26 * 32), rem = idx - v * (26 * 32), i = rem / 32, x = rem - 32 * i, xm = x >= 27 ? x - 27 : x;
            uint32_t* blk = reinterpret_cast<uint32_t*>(smem + L::DEC_A + v * L::DEC_VAR);
            blk[rem] = pl[i][xm][0] | gf->dsc[st_of(g, v, i)][xm];
            blk[26 * 32 + rem] = pl[i][xm][1];
        }
        load_gf(sg, gf);
        for (int t = tid; t < 3 * 128; t += (int)blockDim.x) build_pass_map(smem + L::DEC_MAP, g, t);
        if (lane == 0) { mbar_init(bar, 9); fence_mbar_init(); }
    }
    __syncthreads();
    if (tid < 3) { // a received block r = c (+) 13*st is a codeword iff sum_i T_i[r_i] == sum_i T_i[13*st_i] (GF(3)-linear tables)
        Planes c{0, 0};
        const uint32_t* blk = reinterpret_cast<const uint32_t*>(smem + L::DEC_A + tid * L::DEC_VAR);
        for (int i = 0; i < 26; ++i) {
            const int idx = i * 32 + 13 * (int)st_of(g, tid, i);
            gf3_add(c, blk[idx] & ~0xFFu, blk[26 * 32 + idx]);
        }
        reinterpret_cast<uint32_t*>(smem + L::DEC_CHK)[2 * tid] = c.nz;
        reinterpret_cast<uint32_t*>(smem + L::DEC_CHK)[2 * tid + 1] = c.two;
    }
    __syncthreads();
    const uint32_t tabA32 = smem_u32(smem + L::DEC_A);
    const uint32_t* chk = reinterpret_cast<const uint32_t*>(smem + L::DEC_CHK);
    const uint64_t in_limit = P.in_stride * (P.n_frames - 1) + 9 * g.n_out;
    uint32_t mt_lo, mt_hi;
    warp_range_smsp((uint64_t)P.n_tiles * P.n_frames, blockIdx.x, gridDim.x, warp, V4_DEC_WARPS, mt_lo, mt_hi);
    // lane b < 9 bulk-loads band b's run of tile mt: the 16-byte aligned superset of the run, clipped to the buffer
    auto fetch = [&](uint32_t mt) {
        const uint32_t f = mt / P.n_tiles, tile = P.tile0 + (mt - f * P.n_tiles);
        const uint64_t lo = P.in_stride * f + 52 + 26 * (g.cw_base[lane] + (uint64_t)C_MINI * tile), a0 = lo & ~15ull;
        uint32_t bytes = (uint32_t)((lo - a0) + L::RUN + 15) & ~15u;
        if (a0 + bytes > in_limit) bytes = (uint32_t)(in_limit - a0) & ~15u;
        fence_async_smem();
        mbar_expect_tx(bar, bytes);
        if (bytes) bulk_g2s(smem_u32(R + L::RUN_PITCH * lane), P.in + a0, bytes, bar);
    };
    if (mt_lo < mt_hi && lane < 9) fetch(mt_lo);
    uint32_t phase = 0;
    for (uint32_t mt = mt_lo; mt < mt_hi; ++mt) {
        const uint32_t f = mt / P.n_tiles, tile = P.tile0 + (mt - f * P.n_tiles);
        const bool first = mt == mt_lo || tile == P.tile0, last = mt + 1 == mt_hi || tile + 1 == P.tile0 + P.n_tiles; // of a contiguous stretch
        setup_runs3(meta, g, P.in_stride * f, tile, lane);
        mbar_wait(bar, phase);
        phase ^= 1;
        __syncwarp();
        // ---- the nine runs are in R as they lie in the frame (phase B scales the symbols itself).  Only the clipped end of the buffer
        // (the last chunks of the last frame, which the bulk copy left out) is fetched here, byte by byte
        if (lane < 9) {
            const uint64_t lo = meta.run_lo[lane];
            const uint32_t padb = (uint32_t)lo & 15u;
            const uint64_t a0 = lo - padb, a1 = a0 + ((padb + L::RUN + 15u) & ~15u);
            if (a1 > in_limit)
                for (uint64_t ga = (in_limit > a0 ? (in_limit - a0) & ~15ull : 0) + a0; ga < a1; ++ga) R[L::RUN_PITCH * lane + (ga - a0)] = ga < in_limit ? P.in[ga] : 0;
        }
        __syncwarp();
        if (tile == 0 && lane == 0 && g.cw_base[0] == 0) { // body symbols 0,1: move them from the transient states to the periodic ones
            uint8_t* r0 = R + ((uint32_t)meta.run_lo[0] & 15u);
            r0[0] = sg.scr[st_of(g, 0, 0)][sg.dsc[g.st[0]][r0[0] % 27u]];
            r0[1] = sg.scr[st_of(g, 0, 1)][sg.dsc[g.st[1]][r0[1] % 27u]];
        }
        __syncwarp();
        dec_phase_b<K, false>(R, S, meta, smem + L::DEC_MAP + 128 * (tile % 3u), tabA32, smem + L::DEC_A, chk, sg, chien_of(gf), P.status + 2 * f, lane);
        __syncwarp();
        if (mt + 1 < mt_hi && lane < 9) fetch(mt + 1);                       // R is free again: next tile's runs on their way
        const uint64_t g_lo = P.out_stride * f + (uint64_t)PIX * tile;
        const uint32_t pad = (uint32_t)(g_lo & 15);
        if (lane == 0) {
            bulk_wait_read();                                                // the previous tile's bulk store has read OUT
            if (!first) *reinterpret_cast<uint4*>(OUT) = carry[0];           // bytes [0, pad): the previous tile's tail
        }
        __syncwarp();
        if constexpr (WORDS) dec_phase_a_words<K>(S, OUT, pad, lane); else dec_phase_a<K>(S, OUT, pad, lane);
        fence_async_smem();
        __syncwarp();
        {   // the RGB tile -> global: whole chunks by one bulk store, edge bytes of a stretch one by one
            const uint32_t end = pad + PIX, cend = end >> 4, c0 = (first && pad) ? 1u : 0u;
            uint8_t* g0 = P.out + (g_lo - pad);
            if (lane == 0) {
                if (!last) carry[0] = *reinterpret_cast<const uint4*>(OUT + 16 * cend);
                bulk_s2g(g0 + 16 * c0, smem_u32(OUT + 16 * c0), 16 * (cend - c0));
                bulk_commit();
            }
            if (first && pad && lane >= (int)pad && lane < 16) g0[lane] = OUT[lane];
            if (last && lane < 16 && 16 * cend + lane < end) g0[16 * cend + lane] = OUT[16 * cend + lane];
        }
        __syncwarp();
    }
    if (lane == 0) bulk_wait_all();
}

#include "k_fast5.cuh"

// occupancy (and the opt-in to > 48 KB of dynamic shared memory) per kernel AND per device: a process may drive several GPUs
static int persistent_ctas_per_sm(const void* kern, int tpb, int smem_bytes)
{
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, int> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find({kern, dev});
    if (it != cache.end()) return it->second;
    int n = 0;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, tpb, smem_bytes);
    if (n < 1) n = 1;
    cache[{kern, dev}] = n;
    return n;
}
// the config's CTA-shared image of the v5 kernels (FastImageCache, launch.h), built on `st` the first time a config is seen
template <int K>
static const uint8_t* v5_image(const DevTables& T, const Geom& g, bool dec, cudaStream_t st, int& launches)
{
    static_assert(Cfg5<K>::ENC_WARP <= FastImageCache::BYTES && Cfg5<K>::DEC_WARP <= FastImageCache::BYTES && Cfg5<K>::ENC_WARP % 16 == 0 && Cfg5<K>::DEC_WARP % 16 == 0, "image size");
    FastImageCache& C = *T.img;
    uint8_t key[40] = {};
    key[0] = (uint8_t)K; key[1] = dec ? 1 : g.arith; key[2] = dec ? 1 : 0;      // the decoder's tables are always the repaired code's
    for (int i = 0; i < 8; ++i) key[3 + i] = g.st[i];
    for (int b = 0; b < 9; ++b) key[11 + b] = (uint8_t)(g.cw_base[b] % 3);
    for (int i = 0; i < FastImageCache::N; ++i)
        if (C.e[i].valid && std::memcmp(C.e[i].key, key, sizeof key) == 0) return C.base + (size_t)i * FastImageCache::BYTES;
    int slot = -1;
    for (int i = 0; i < FastImageCache::N; ++i) if (!C.e[i].valid) { slot = i; break; }
    if (slot < 0) { cudaDeviceSynchronize(); slot = C.next; C.next = (C.next + 1) % FastImageCache::N; }
    C.e[slot].valid = false;
    uint8_t* img = C.base + (size_t)slot * FastImageCache::BYTES;
    if (dec) {
        cudaFuncSetAttribute(k_v5_image_dec<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg5<K>::DEC_WARP);
        k_v5_image_dec<K><<<1, 256, Cfg5<K>::DEC_WARP, st>>>(g, T.gf, T.rs, img);
    } else {
        cudaFuncSetAttribute(k_v5_image_enc<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg5<K>::ENC_WARP);
        k_v5_image_enc<K><<<1, 256, Cfg5<K>::ENC_WARP, st>>>(g, T.gf, T.rs, img);
    }
    ++launches;
    cudaStreamSynchronize(st);   // once per new config: later calls may come on other streams
    std::memcpy(C.e[slot].key, key, sizeof key);
    C.e[slot].valid = true;
    return img;
}
typedef CUresult (*tensor_map_encode_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tensor_map_encode_t tensor_map_encode_fn()   // cuTensorMapEncodeTiled through the runtime (no link against libcuda)
{
    static tensor_map_encode_t enc = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess) enc = (tensor_map_encode_t)fn;
        else cudaGetLastError();
    }
    return enc;
}
// T3C_V5_FLAGS overrides both defaults.  The decoder runs without the staggered start since its screen sums the K data positions only
// (r02q: 142.4 us against 144.4 us with it; the encoder still gains: 147.5 against 153.5 us)
static uint32_t v5_flags(bool dec = false)
{
    static int fl = -1, env = 0;
    if (fl < 0) { const char* e = getenv("T3C_V5_FLAGS"); env = e != nullptr; fl = e ? atoi(e) : (int)V5_FLAGS_DEFAULT; }
    return dec && !env ? 0u : (uint32_t)fl;
}
template <int K, bool WORDS>
static int launch_v5_enc(const DevTables& T, FastParams P, const Geom& g, cudaStream_t st)
{
    static int occ = 0;
    using L5 = Cfg5<K, WORDS>;
    int n = 0;
    const uint8_t* img = v5_image<K>(T, g, false, st, n);
    occ = persistent_ctas_per_sm(reinterpret_cast<const void*>(k_encode_v5<K, WORDS>), 32 * L5::ENC_WARPS, L5::TOTAL_ENC);
    const uint64_t total = (uint64_t)P.n_tiles * P.n_frames, need = (total + L5::ENC_WARPS - 1) / L5::ENC_WARPS;
    uint64_t grid = (uint64_t)T.sm_count * occ;
    if (grid > need) grid = need;
    if (!grid) return n;
    P.flags = v5_flags() & ~16u;
    // One 3-D tensor store (TMA, SASS UTMASTG) per mini-tile instead of nine bulk stores when the nine runs of a tile form a regular box
    // (the condition of launch_v5_dec's tensor copy, on the output buffer).  T3C_V5_FLAGS |= 32 turns it off.
    CUtensorMap tmap;
    std::memset(&tmap, 0, sizeof tmap);
    if (!(v5_flags() & 32u)) {
        bool regular = P.out_stride % 16 == 0 || P.n_frames == 1;
        for (int b = 1; b < 9; ++b) regular = regular && g.ncw[b] == g.ncw[0];
        const uint64_t pitch = 26 * g.ncw[0];
        regular = regular && pitch % 16 == 0 && pitch >= 4096 && ((uintptr_t)P.out & 15) == 0 && g.cw_base[0] == 0;
        if (regular) {
            const cuuint64_t frame_pitch = P.n_frames > 1 ? P.out_stride : (9 * pitch + 4096 + 15) / 16 * 16;
            const cuuint64_t dims[3] = {pitch / 2, 9, P.n_frames}, strides[2] = {pitch, frame_pitch};
            const cuuint32_t box[3] = {(cuuint32_t)L5::RUN_PITCH / 2, 9, 1}, estr[3] = {1, 1, 1};
            if (tensor_map_encode_fn() && tensor_map_encode_fn()(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, P.out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
                P.flags |= 16u;
                P.band_stride = pitch;
                // ((52 + 338 t) & ~15) + RUN_PITCH <= pitch  for every t < ts_tiles
                uint64_t t = pitch >= (uint64_t)L5::RUN_PITCH + 52 ? (pitch - L5::RUN_PITCH - 52) / 338 + 1 : 0;
                while (t && ((52ull + 338ull * (t - 1)) & ~15ull) + L5::RUN_PITCH > pitch) --t;
                while (((52ull + 338ull * t) & ~15ull) + L5::RUN_PITCH <= pitch) ++t;
                P.ts_tiles = (uint32_t)(t > 0xFFFFFFFFull ? 0xFFFFFFFFull : t);
            }
        }
    }
    if (getenv("T3C_DEBUG")) std::fprintf(stderr, "t3c: k_encode_v5<%d,%d> grid %u, flags %#x (tensor store %s)\n", K, (int)WORDS, (unsigned)grid, P.flags, (P.flags & 16u) ? "on" : "off");
    k_encode_v5<K, WORDS><<<(unsigned)grid, 32 * L5::ENC_WARPS, L5::TOTAL_ENC, st>>>(P, g, T.gf, img, tmap);
    return n + 1;
}
template <int K, bool WORDS>
static int launch_v5_dec(const DevTables& T, FastParams P, const Geom& g, cudaStream_t st)
{
    static int occ = 0;
    using L5 = Cfg5<K, WORDS>;
    int n = 0;
    const uint8_t* img = v5_image<K>(T, g, true, st, n);
    occ = persistent_ctas_per_sm(reinterpret_cast<const void*>(k_decode_v5<K, WORDS>), 32 * L5::DEC_WARPS, L5::TOTAL_DEC);
    const uint64_t total = (uint64_t)P.n_tiles * P.n_frames, need = (total + L5::DEC_WARPS - 1) / L5::DEC_WARPS;
    uint64_t grid = (uint64_t)T.sm_count * occ;
    if (grid > need) grid = need;
    if (!grid) return n;
    P.flags = v5_flags(true) & ~2u;
    // One 3-D tensor copy (TMA, SASS UTMALDG) per mini-tile instead of nine bulk copies when the nine runs of a tile form a regular
    // box: every band holds the same number of codewords and the band pitch 26 * ncw is a multiple of 16 bytes (8K / 4K / 1080p frames
    // at k = 20 are) and frames start on 16-byte boundaries.  Tensor {band pitch, 9 bands, frames} of 16-bit elements over the input
    // buffer, box {RUN_PITCH, 9, 1} at the 16-byte aligned column below the runs' first byte 52 + 338 * tile.  Tiles whose box would
    // reach past the band's row (the last one or two of a frame) keep the nine bulk copies.
    CUtensorMap tmap;
    std::memset(&tmap, 0, sizeof tmap);
    if (!(v5_flags(true) & 4u)) {
        bool regular = P.in_stride % 16 == 0 || P.n_frames == 1;
        for (int b = 1; b < 9; ++b) regular = regular && g.ncw[b] == g.ncw[0];
        const uint64_t pitch = 26 * g.ncw[0];
        regular = regular && pitch % 16 == 0 && pitch >= 4096 && ((uintptr_t)P.in & 15) == 0 && g.cw_base[0] == 0;
        if (regular) {
            typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
            static encode_fn enc = nullptr;
            static bool tried = false;
            if (!tried) {
                tried = true;
                void* fn = nullptr;
                cudaDriverEntryPointQueryResult qr;
                if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess) enc = (encode_fn)fn;
                else cudaGetLastError();
            }
            if (enc) {
                const cuuint64_t frame_pitch = P.n_frames > 1 ? P.in_stride : (9 * pitch + 4096 + 15) / 16 * 16;
                // 16-bit elements: a box side holds at most 256 elements, a run needs 368 bytes (runs start on even bytes: x = 2 + 169 * tile)
                const cuuint64_t dims[3] = {pitch / 2, 9, P.n_frames}, strides[2] = {pitch, frame_pitch};
                const cuuint32_t box[3] = {(cuuint32_t)L5::RUN_PITCH / 2, 9, 1}, estr[3] = {1, 1, 1};
                if (enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<uint8_t*>(P.in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
                    P.flags |= 2u;
                    P.band_stride = pitch;
                }
            }
        }
    }
    if (getenv("T3C_DEBUG")) std::fprintf(stderr, "t3c: k_decode_v5<%d,%d> grid %u, flags %#x, band pitch %llu (tensor copy %s)\n", K, (int)WORDS, (unsigned)grid, P.flags,
                                          (unsigned long long)P.band_stride, (P.flags & 2u) ? "on" : "off");
    k_decode_v5<K, WORDS><<<(unsigned)grid, 32 * L5::DEC_WARPS, L5::TOTAL_DEC, st>>>(P, g, img, tmap);
    return n + 1;
}
template <class Kern>
int launch_persistent(Kern kern, int smem_bytes, const DevTables& T, const FastParams& P, const Geom& g, cudaStream_t st, int& ctas_per_sm,
                      int warps = FAST_WARPS)
{
    ctas_per_sm = persistent_ctas_per_sm(reinterpret_cast<const void*>(kern), 32 * warps, smem_bytes);
    const uint64_t total = (uint64_t)P.n_tiles * P.n_frames;
    uint64_t grid = (uint64_t)T.sm_count * ctas_per_sm;
    const uint64_t need = (total + warps - 1) / warps;
    if (grid > need) grid = need;
    if (!grid) return 0;
    kern<<<(unsigned)grid, 32 * warps, smem_bytes, st>>>(P, g, T.gf, T.rs);
    return 1;
}
// T3C_FAST=3 / 4 keep the v3 kernels (plain loads/stores, 4 CTAs per SM) / the v4 kernels (bulk-async I/O) for A/B comparison;
// default is v5 (k_fast5.cuh)
static bool smem_window_ok();
static int fast_version()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("T3C_FAST"); v = (e && e[0] == '3') ? 3 : (e && e[0] == '4') ? 4 : 5; }
    return v == 5 && !smem_window_ok() ? 4 : v;
}
static bool use_v4() { return fast_version() >= 4; }
// v5 addresses its phase-B tables absolutely (k_fast5.cuh, SMEM_WINDOW_BASE): check the assumption once per device
static bool smem_window_ok()
{
    static std::mutex mu;
    static std::map<int, bool> ok;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    auto it = ok.find(dev);
    if (it != ok.end()) return it->second;
    uint32_t* d = nullptr;
    uint32_t h = 0;
    bool r = false;
    if (cudaMalloc(&d, 4) == cudaSuccess) {
        k_smem_window_probe<<<1, 32, 64>>>(d);
        r = cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost) == cudaSuccess && h == SMEM_WINDOW_BASE;
        cudaFree(d);
    }
    ok[dev] = r;
    return r;
}
// tiles [t0, t1) of [0, n_full) of every frame go to the v3 kernels; with `tail` the ragged rest [n_full, n_all) goes to
// the general-tile kernels
template <int K>
int launch_enc(const DevTables& T, FastParams P, const Geom& g, cudaStream_t st, uint32_t n_full, uint32_t t0, uint32_t t1, bool tail, bool words = false)
{
    static int occ4 = 0, occ4w = 0, occ3 = 0, occ2 = 0;
    const uint32_t n_all = P.n_tiles;
    int n = 0;
    if constexpr (K >= 18) {
        if (t1 > n_full) t1 = n_full;
        if (t1 > t0) {
            P.tile0 = t0; P.n_tiles = t1 - t0;
            if (fast_version() == 5 && words) n += launch_v5_enc<K, true>(T, P, g, st);
            else if (fast_version() == 5) n += launch_v5_enc<K, false>(T, P, g, st);
            else if (words) n += launch_persistent(k_encode_rgb_v4<K, true>, Cfg4<K, true>::TOTAL_ENC, T, P, g, st, occ4w, Cfg4<K, true>::ENC_WARPS);
            else if (use_v4()) n += launch_persistent(k_encode_rgb_v4<K, false>, Cfg4<K>::TOTAL_ENC, T, P, g, st, occ4, Cfg4<K>::ENC_WARPS);
            else n += launch_persistent(k_encode_rgb_v3<K>, Cfg3<K>::TOTAL_ENC, T, P, g, st, occ3);
        }
    } else n_full = 0;
    if (tail && n_all > n_full) { P.tile0 = n_full; P.n_tiles = n_all - n_full; n += launch_persistent(k_encode_rgb_fast<K>, Cfg<K>::TOTAL_ENC, T, P, g, st, occ2); }
    return n;
}
template <int K>
int launch_dec(const DevTables& T, FastParams P, const Geom& g, cudaStream_t st, uint32_t n_full, uint32_t t0, uint32_t t1, bool tail, bool words = false)
{
    static int occ4 = 0, occ4w = 0, occ3 = 0, occ2 = 0;
    const uint32_t n_all = P.n_tiles;
    int n = 0;
    if constexpr (K >= 18) {
        if (t1 > n_full) t1 = n_full;
        if (t1 > t0) {
            P.tile0 = t0; P.n_tiles = t1 - t0;
            if (fast_version() == 5 && words) n += launch_v5_dec<K, true>(T, P, g, st);
            else if (fast_version() == 5) n += launch_v5_dec<K, false>(T, P, g, st);
            else if (words) n += launch_persistent(k_decode_rgb_v4<K, true>, Cfg4<K, true>::TOTAL_DEC, T, P, g, st, occ4w, Cfg4<K, true>::DEC_WARPS);
            else if (use_v4()) n += launch_persistent(k_decode_rgb_v4<K, false>, Cfg4<K>::TOTAL_DEC, T, P, g, st, occ4, Cfg4<K>::DEC_WARPS);
            else n += launch_persistent(k_decode_rgb_v3<K>, Cfg3<K>::TOTAL_DEC, T, P, g, st, occ3);
        }
    } else n_full = 0;
    if (tail && n_all > n_full) { P.tile0 = n_full; P.n_tiles = n_all - n_full; n += launch_persistent(k_decode_rgb_fast<K>, Cfg<K>::TOTAL_DEC, T, P, g, st, occ2); }
    return n;
}
// mini-tiles of a frame whose 117 codewords all exist and whose 27k pixels lie inside [0, px_limit)
uint32_t full_tiles(const Geom& g, uint64_t px_limit)
{
    uint64_t mn = ~0ull;
    for (int b = 0; b < 9; ++b) mn = g.ncw[b] < mn ? g.ncw[b] : mn;
    uint64_t n = mn / C_MINI;
    const uint64_t by_px = px_limit / (27ull * (uint64_t)g.uniform_k);
    if (by_px < n) n = by_px;
    return (uint32_t)n;
}

#include "k_super.cuh"

} // namespace

// ---- super-tile kernels: launchers ------------------------------------------------------------
bool super_path_ok(const t3c_config& cfg) { return super_config_ok(cfg); }
// host-only view of the plan the super-tile kernels would use for one super-frame of n_words raw words (tests, tools):
// out16 = {M, UN, n_tiles, nk, kk[4], ncw[4], npass[3], smem_bytes}; map / kv as uploaded to the device (may be null)
int super_plan_describe(const t3c_config& cfg, size_t n_words, int decode, int words, uint32_t* out16, uint16_t* map, uint8_t* kv)
{
    Geom g;
    make_geom(cfg, n_words, 1, g);
    SuperPlan P;
    if (!make_super_plan(cfg, g, decode != 0, words != 0, 2 * (uint64_t)n_words, P)) return 0;
    static thread_local uint16_t h_map[3 * SUP_MAX_PASS * 32];
    static thread_local uint8_t h_kv[3 * SUP_MAX_PASS];
    uint32_t npass[3];
    if (!build_super_maps(P, g, h_map, h_kv, npass)) return 0;
    const uint32_t v[16] = {P.M, P.UN, P.n_tiles, P.nk, P.kk[0], P.kk[1], P.kk[2], P.kk[3], P.ncw[0], P.ncw[1], P.ncw[2], P.ncw[3],
                            npass[0], npass[1], npass[2], P.smem_bytes};
    std::memcpy(out16, v, sizeof v);
    if (map) std::memcpy(map, h_map, sizeof h_map);
    if (kv) std::memcpy(kv, h_kv, sizeof h_kv);
    return 1;
}
// debug builds (-DT3C_SUPER_DEBUG): per-phase cycle counters of CTA 0, read and reset
int super_debug_counters(uint32_t* out32)
{
#ifdef T3C_SUPER_DEBUG
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out32, g_sup_dbg, 32 * sizeof(uint32_t));
    uint32_t z[32] = {};
    cudaMemcpyToSymbol(g_sup_dbg, z, sizeof z);
    return 1;
#else
    (void)out32;
    return 0;
#endif
}

static void super_tail_all(SuperTail* tail)
{
    for (int b = 0; b < 9; ++b) tail->cs.c[b] = 0;
    tail->m_start = 0;
    tail->unit_start = 0;
    tail->n_tiles = 0;
    for (int b = 0; b < 9; ++b) tail->ncw_tile[b] = 0;
}
// plan + pass maps (cached on the device per kernel flavour); false = the super-tile kernels do not apply
static bool super_prepare(const DevTables& T, const t3c_config& cfg, const Geom& g, bool decode, bool words, uint64_t px_limit, cudaStream_t st, SuperPlan& P,
                          SuperTail* tail)
{
    super_tail_all(tail);
    if (!T.sup) return false;
    if (!make_super_plan(cfg, g, decode, words, px_limit, P)) return false;
    SuperCache::Slot& C = T.sup->slot[(decode ? 2 : 0) + (words ? 1 : 0)];
    if (!C.d_map) return false;
    uint8_t key[48] = {};
    for (int b = 0; b < 9; ++b) { key[b] = (uint8_t)g.k[b]; key[9 + b] = (uint8_t)(g.cw_base[b] % 3); }
    std::memcpy(key + 20, &P.M, 4);
    if (!C.valid || std::memcmp(C.key, key, sizeof key) != 0) {
        static thread_local uint16_t h_map[3 * SUP_MAX_PASS * 32];
        static thread_local uint8_t h_kv[3 * SUP_MAX_PASS];
        C.valid = false;
        if (!build_super_maps(P, g, h_map, h_kv, C.npass)) return false;
        // once per config change: kernels in flight on any stream may still read the old maps, and the copies (pageable memory, legacy
        // stream) must have landed before a kernel on a non-blocking stream reads the new ones
        (void)st;
        if (cudaDeviceSynchronize() != cudaSuccess) return false;
        if (cudaMemcpy(C.d_map, h_map, sizeof h_map, cudaMemcpyHostToDevice) != cudaSuccess) return false;
        if (cudaMemcpy(C.d_kv, h_kv, sizeof h_kv, cudaMemcpyHostToDevice) != cudaSuccess) return false;
        if (cudaDeviceSynchronize() != cudaSuccess) return false;
        std::memcpy(C.key, key, sizeof key);
        C.valid = true;
    }
    for (int i = 0; i < 3; ++i) P.npass[i] = C.npass[i];
    P.map = C.d_map;
    P.pass_kv = C.d_kv;
    for (int b = 0; b < 9; ++b) { tail->cs.c[b] = (uint64_t)P.ncw[P.kslot[b]] * P.n_tiles; tail->ncw_tile[b] = P.ncw[P.kslot[b]]; }
    tail->n_tiles = P.n_tiles;
    tail->m_start = (uint64_t)P.M * P.n_tiles;
    tail->unit_start = (uint64_t)P.UN * P.n_tiles;
    return true;
}
template <class Kern>
static int super_launch(Kern kern, const DevTables& T, const FastParams& Q, const SuperPlan& P, const Geom& g, cudaStream_t st)
{
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, uint32_t> opted; // largest dynamic shared memory opted in per (kernel, device)
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lock(mu);
        uint32_t& cur = opted[{reinterpret_cast<const void*>(kern), dev}];
        if (cur < P.smem_bytes) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem_bytes) != cudaSuccess) return -1;
            cur = P.smem_bytes;
        }
    }
    const uint64_t total = (uint64_t)P.n_tiles * Q.n_frames;
    uint64_t grid = (uint64_t)P.ctas_per_sm * (uint64_t)T.sm_count;
    if (const char* e = getenv("T3C_SUPER_GRID")) { const int v = atoi(e); if (v > 0) grid = (uint64_t)v; } // experiments only
    if (grid > total) grid = total;
    kern<<<(unsigned)grid, SUP_TPB, P.smem_bytes, st>>>(Q, P, g, T.gf, T.rs);
    return 1;
}
// the 8-pixels-per-thread bridge kernels (k_fast5.cuh) for the leading multiple of eight pixels of 16-byte aligned buffers; returns the
// number of pixels covered (the caller's one-pixel kernels take the rest)
size_t launch_rgb_to_quant8(const uint8_t* rgb, size_t n_px, t3c_pixel* out, cudaStream_t st)
{
    const size_t groups = n_px / 8;
    if (!groups || (reinterpret_cast<uintptr_t>(rgb) & 7u) || (reinterpret_cast<uintptr_t>(out) & 15u)) return 0;
    k_rgb_to_quant8<<<(unsigned)((groups + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint2*>(rgb), groups, reinterpret_cast<uint4*>(out));
    return 8 * groups;
}
size_t launch_quant_to_rgb8(const t3c_pixel* px, size_t n_px, uint8_t* rgb, cudaStream_t st)
{
    const size_t groups = n_px / 8;
    if (!groups || (reinterpret_cast<uintptr_t>(rgb) & 7u) || (reinterpret_cast<uintptr_t>(px) & 15u)) return 0;
    k_quant_to_rgb8<<<(unsigned)((groups + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(px), groups, reinterpret_cast<uint2*>(rgb));
    return 8 * groups;
}
int launch_encode_super(const DevTables& T, const t3c_config& cfg, const Geom& g, const uint8_t* in, size_t in_pitch, bool words, size_t n_px,
                        size_t n_frames, uint8_t* out9, size_t stride_words, cudaStream_t st, SuperTail* tail)
{
    super_tail_all(tail);
    if ((((uintptr_t)in | (uintptr_t)out9) & 15) || !n_frames) return 0;      // 128-bit transfers need 16-byte aligned buffer bases
    if (n_frames > 1 && (stride_words & 1)) return 0;                           // every frame's body must start on an even byte
    SuperPlan P;
    if (!super_prepare(T, cfg, g, false, words, n_px < 2 * g.n_words ? n_px : 2 * g.n_words, st, P, tail)) return 0;
    FastParams Q{};
    Q.in = in; Q.out = out9;
    Q.in_stride = in_pitch; Q.out_stride = 9ull * stride_words;
    Q.n_frames = (uint32_t)n_frames;
    const int n = words ? super_launch(k_encode_super<true>, T, Q, P, g, st) : super_launch(k_encode_super<false>, T, Q, P, g, st);
    if (n <= 0) { super_tail_all(tail); return 0; }
    return n;
}
int launch_decode_super(const DevTables& T, const t3c_config& cfg, const Geom& g, const uint8_t* in9, size_t stride_words, size_t n_frames, uint8_t* out,
                        size_t out_pitch, bool words, size_t n_px_out, uint32_t* d_status, cudaStream_t st, SuperTail* tail)
{
    super_tail_all(tail);
    if ((((uintptr_t)in9 | (uintptr_t)out) & 15) || !n_frames) return 0;
    if (n_frames > 1 && ((stride_words & 1) || (out_pitch & 1))) return 0;      // codewords and pixel units are read / written in 2-byte pieces
    SuperPlan P;
    if (!super_prepare(T, cfg, g, true, words, n_px_out, st, P, tail)) return 0;
    FastParams Q{};
    Q.in = in9; Q.out = out;
    Q.in_stride = 9ull * stride_words; Q.out_stride = out_pitch;
    Q.n_frames = (uint32_t)n_frames;
    Q.status = d_status;
    const int n = words ? super_launch(k_decode_super<true>, T, Q, P, g, st) : super_launch(k_decode_super<false>, T, Q, P, g, st);
    if (n <= 0) { super_tail_all(tail); return 0; }
    return n;
}

bool fast_path_ok(const t3c_config& cfg)
{
    if (cfg.profile == T3C_PROFILE_RAW) return false;
    if (use_2d(cfg) || use_beacon(cfg)) return false;
    for (int b = 1; b < 9; ++b) if (cfg.uep[b] % 4 != cfg.uep[0] % 4) return false;
    return true;
}

static uint32_t all_tiles(const Geom& g)
{
    uint64_t mx = 0;
    for (int b = 0; b < 9; ++b) mx = g.ncw[b] > mx ? g.ncw[b] : mx;
    return (uint32_t)((mx + C_MINI - 1) / C_MINI);
}
uint32_t fast_full_tiles_encode(const Geom& g, size_t n_px)
{
    if (g.uniform_k < 18) return 0;
    return full_tiles(g, n_px < 2 * g.n_words ? n_px : 2 * g.n_words);
}
uint32_t fast_full_tiles_decode(const Geom& g, size_t n_px_out, size_t out_pitch, size_t n_frames)
{
    if (g.uniform_k < 18) return 0;
    return (n_frames > 1 && (out_pitch & 1)) ? 0 : full_tiles(g, n_px_out); // v3 writes RGB with 2-byte stores: frames start on even bytes
}

int launch_encode_rgb_fast_part(const DevTables& T, const t3c_config& cfg, const Geom& g, const uint8_t* rgb, size_t n_px, size_t in_pitch,
                                size_t n_frames, uint8_t* out, size_t stride_words, cudaStream_t st, uint32_t t0, uint32_t t1, bool tail)
{
    if (((uintptr_t)rgb | (uintptr_t)out) & 15) return -1; // 128-bit transfers need 16-byte aligned buffer bases
    if (n_frames > 1 && (stride_words & 1)) return -1;     // and every frame's body must start on an even byte
    FastParams P{};
    P.in = rgb; P.out = out;
    P.in_stride = in_pitch; P.out_stride = 9ull * stride_words;
    P.n_px = n_px; P.n_frames = (uint32_t)n_frames;
    P.n_tiles = all_tiles(g);
    int n = 0;
    const uint32_t n_full = fast_full_tiles_encode(g, n_px);
    switch (g.uniform_k) {
#ifndef T3C_DEV_K20
    case 24: n = launch_enc<24>(T, P, g, st, n_full, t0, t1, tail); break;
#endif
#ifndef T3C_DEV_K20
    case 22: n = launch_enc<22>(T, P, g, st, n_full, t0, t1, tail); break;
#endif
    case 20: n = launch_enc<20>(T, P, g, st, n_full, t0, t1, tail); break;
#ifndef T3C_DEV_K20
    case 18: n = launch_enc<18>(T, P, g, st, n_full, t0, t1, tail); break;
#endif
    default: return 0;
    }
    if (tail) n += launch_frame_finish(T, cfg, g, out, n_frames, 9ull * stride_words, st); // fast path: no beacon, only header + padding
    return n;
}
int launch_encode_rgb_fast(const DevTables& T, const t3c_config& cfg, const Geom& g, const uint8_t* rgb, size_t n_px, size_t n_frames,
                           uint8_t* out, size_t stride_words, cudaStream_t st)
{
    return launch_encode_rgb_fast_part(T, cfg, g, rgb, n_px, 3 * n_px, n_frames, out, stride_words, st, 0, ~0u, true);
}

int launch_decode_rgb_fast_part(const DevTables& T, const Geom& g, const uint8_t* in, size_t stride_words, size_t n_frames, size_t n_px,
                                size_t out_pitch, size_t n_px_out, uint8_t* rgb, uint32_t* d_status, cudaStream_t st, const uint32_t* chk_nz,
                                const uint32_t* chk_two, uint32_t t0, uint32_t t1, bool tail)
{
    if (((uintptr_t)rgb | (uintptr_t)in) & 15) return -1;
    if (n_frames > 1 && (stride_words & 1)) return -1;
    FastParams P{};
    P.in = in; P.out = rgb;
    P.in_stride = 9ull * stride_words; P.out_stride = out_pitch;
    P.n_px = n_px; P.px_out = n_px_out; P.n_frames = (uint32_t)n_frames;
    P.status = d_status;
    for (int i = 0; i < 7; ++i) { P.chk_nz[i] = chk_nz[i]; P.chk_two[i] = chk_two[i]; }
    P.n_tiles = all_tiles(g);
    const uint32_t n_full = fast_full_tiles_decode(g, n_px_out, out_pitch, n_frames);
    switch (g.uniform_k) {
#ifndef T3C_DEV_K20
    case 24: return launch_dec<24>(T, P, g, st, n_full, t0, t1, tail);
#endif
#ifndef T3C_DEV_K20
    case 22: return launch_dec<22>(T, P, g, st, n_full, t0, t1, tail);
#endif
    case 20: return launch_dec<20>(T, P, g, st, n_full, t0, t1, tail);
#ifndef T3C_DEV_K20
    case 18: return launch_dec<18>(T, P, g, st, n_full, t0, t1, tail);
#endif
    }
    return 0;
}
int launch_decode_rgb_fast(const DevTables& T, const t3c_config& cfg, const Geom& g, const uint8_t* in, size_t stride_words,
                           size_t n_frames, size_t n_px, size_t n_px_out, uint8_t* rgb, uint32_t* d_status, cudaStream_t st,
                           const uint32_t* chk_nz, const uint32_t* chk_two)
{
    (void)cfg;
    return launch_decode_rgb_fast_part(T, g, in, stride_words, n_frames, n_px, 3 * n_px, n_px_out, rgb, d_status, st, chk_nz, chk_two, 0, ~0u, true);
}

// ---- raw-word variants (encode_profile_from_raw / the consistent decoder on Word27 streams): the v4 kernels on the full
// mini-tiles [0, *n_full) of one super-frame; the caller runs the general kernels on the codewords / words after them.
// Return -1 when the fast path does not apply (K = 18, unaligned buffers).
int launch_encode_words_fast(const DevTables& T, const Geom& g, const uint8_t* raw9, uint8_t* out9, cudaStream_t st, uint32_t* n_full_out)
{
    *n_full_out = 0;
    if (g.uniform_k < 18 || (((uintptr_t)raw9 | (uintptr_t)out9) & 15)) return -1;
    const uint32_t n_full = full_tiles(g, 2 * g.n_words);
    if (!n_full) return 0;
    FastParams P{};
    P.in = raw9; P.out = out9;
    P.in_stride = 9ull * g.n_words; P.out_stride = 9ull * g.n_out;
    P.n_px = 2 * g.n_words; P.n_frames = 1;
    P.n_tiles = all_tiles(g);
    int n = 0;
    switch (g.uniform_k) {
#ifndef T3C_DEV_K20
    case 24: n = launch_enc<24>(T, P, g, st, n_full, 0, n_full, false, true); break;
#endif
#ifndef T3C_DEV_K20
    case 22: n = launch_enc<22>(T, P, g, st, n_full, 0, n_full, false, true); break;
#endif
    case 20: n = launch_enc<20>(T, P, g, st, n_full, 0, n_full, false, true); break;
#ifndef T3C_DEV_K20
    case 18: n = launch_enc<18>(T, P, g, st, n_full, 0, n_full, false, true); break;
#endif
    default: return -1;
    }
    *n_full_out = n_full;
    return n;
}
int launch_decode_words_fast(const DevTables& T, const Geom& g, const uint8_t* in9, uint8_t* raw9, size_t n_words_out, uint32_t* d_status,
                             cudaStream_t st, uint32_t* n_full_out)
{
    *n_full_out = 0;
    if (g.uniform_k < 18 || (((uintptr_t)raw9 | (uintptr_t)in9) & 15)) return -1;
    const uint32_t n_full = full_tiles(g, 2 * n_words_out);
    if (!n_full) return 0;
    FastParams P{};
    P.in = in9; P.out = raw9;
    P.in_stride = 9ull * g.n_out; P.out_stride = 9ull * g.n_words;
    P.n_px = 2 * g.n_words; P.px_out = 2 * n_words_out; P.n_frames = 1;
    P.status = d_status;
    P.n_tiles = all_tiles(g);
    int n = 0;
    switch (g.uniform_k) {
#ifndef T3C_DEV_K20
    case 24: n = launch_dec<24>(T, P, g, st, n_full, 0, n_full, false, true); break;
#endif
#ifndef T3C_DEV_K20
    case 22: n = launch_dec<22>(T, P, g, st, n_full, 0, n_full, false, true); break;
#endif
    case 20: n = launch_dec<20>(T, P, g, st, n_full, 0, n_full, false, true); break;
#ifndef T3C_DEV_K20
    case 18: n = launch_dec<18>(T, P, g, st, n_full, 0, n_full, false, true); break;
#endif
    default: return -1;
    }
    *n_full_out = n_full;
    return n;
}

} // namespace t3c
