// k_fast5.cuh -- v5 of the warp-tile kernels of k_fast.cu (uniform k, 1D 9-band interleave, no beacon: BASELINE configs 0, 1, 4).
// Included by k_fast.cu inside its anonymous namespace; it reuses the digit / plane helpers defined there.
//
// Same work decomposition as v4 (a mini-tile of 13 codewords per band is owned by one warp from its bulk load to its bulk
// stores), rebuilt around the per-phase instruction budget that the v4 SASS gave (round-2 notes in DESIGN.md section 4.1):
// per mini-tile v4 issued ~2300 warp instructions -- 3 x 311 (phase A), 4 x 227 (phase B) and ~460 of per-tile bookkeeping.
//   * bookkeeping: a warp walks its tile range with incremental state (global offsets, scrambler variant class, 16-byte phase of
//     the nine runs live in registers and advance by constants); no divisions, no per-tile metadata in shared memory, the
//     lane -> codeword assignment of phase B is a per-CTA table of ready-made (source, destination, band) records;
//   * phase B: passes 0..2 hold codewords of ONE scrambler variant each, so their table block is a compile-time offset
//     from the shared window: a look-up is LDS [symbol*4 + UR + imm] with no address arithmetic at all (3 LDS + 3 LOP3 per
//     data symbol); only the mixed last pass adds a per-lane block offset;
//   * phase A (RGB): the two chroma channels are computed in exact integer arithmetic straight from the packed pixel words --
//     Cb = round(128 - 0.168736 R - 0.331264 G + 0.5 B) is floor((X + 15625) / 31250) + 128 with X = -5273 R - 10352 G + 15625 B
//     (two 16x8-bit dot products, IDP.2A), and it agrees with the reference's float32 evaluation (IMG:47-56) for all 2^24 colours
//     because no colour comes closer than 32e-6 to a rounding boundary other than exact ties, which float32 also hits exactly;
//     the quantiser (IMG:69-78) follows as one mask-or and one multiply-high.  Luma is computed the same way; 824 of its 16782
//     exact ties round down in float32, so tie candidates are flagged (low nine bits of the scaled sum) and redone in float32;
//   * decode: the nine band runs of a tile arrive by one 3-D tensor copy, YCbCr -> RGB is exact fixed-point arithmetic, dirty
//     codewords go to an out-of-line bounded-distance decoder with uniform control flow (dev.cuh rs_bd_fix);
//   * the CTA-shared tables are built once per configuration (FastImageCache) instead of by every CTA of every launch.
// Later steps of round 2 (DESIGN.md sections 4.1, 7b and 9):
//   * decode phase B screens a codeword by comparing the parity its K data symbols imply with the received parity bytes (K look-ups
//     instead of 26); only codewords that differ finish the syndrome sum, out of line, through .shared addresses (no pointers);
//   * decode phase A: chroma dequantisation by an 81-byte table (21 words: no bank conflict possible), the 18 bytes of a unit
//     gathered by 13 PRMT, compile-time symbol alignment in the first two passes; no staggered start for the decoder;
//   * encode: inside a stretch of a regular frame the nine runs of a tile leave by one 3-D tensor store (UTMASTG); the first two
//     phase-A passes know the alignment of their 26 symbols; both directions: an accumulator's first table entry is a move.
// What bounds these kernels is the register file's read ports (a three-source LOP3 / IDP / funnel shift takes them for two cycles
// whichever pipe executes it: tools/probe/rf_ports.cu, tools/rf_model.py), then the multiply pipe (IMAD.HI: four cycles) in phase A.
#pragma once
#ifndef T3C_ENC_WARPS_CAP
#define T3C_ENC_WARPS_CAP 32   // experiments: fewer encoder warps per CTA
#endif
#ifndef T3C_ENC_WAIT_READ
#define T3C_ENC_WAIT_READ 0      // experiment only: does not order the tensor store's trailing chunks
#endif
#ifndef T3C_ENC_ONE_LOOP
#define T3C_ENC_ONE_LOOP 0       // experiment: one loop body for all phase-A passes of the RGB encoder too (146.6 us against 142.4)
#endif

// ---- exact integer chroma (checked against the float path over all 2^24 colours by tests/test_gpu_parity.py) ------------------------
constexpr uint32_t C5_M = 2251799814u;                         // ceil(2^46 / 31250): hi32(X * M) >> 14 == X / 31250 for X < 2^32
constexpr uint32_t C5_OFF = 128u * 31250u + 15625u + 31250u;   // +128, +0.5 (round half up), +1 (the quantiser below takes C + 1)
constexpr uint32_t C5_Z = 6518u, C5_M2 = 82048u;               // hi32(((C+1) << 14 | Z) * M2) == (641 C + 896) >> 11 == (5C + 7 + (C >= 128)) >> 4
constexpr int pk16(int lo, int hi) { return (int)(((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16)); }
__device__ __forceinline__ int dp2a_lo(int a, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi(int a, uint32_t b, int c)
{
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// coefficient pairs of the dot products (kept in registers: IDP takes no constant-bank operand):
// row c: {pk(c0,c1), pk(c2,0)};  rows: Cb, Cr (units of 1/31250), luma (units of 1/1000)
__constant__ int c5_coef[6] = {pk16(-5273, -10352), pk16(15625, 0), pk16(15625, -13084), pk16(-2541, 0), pk16(299, 587), pk16(114, 0)};
// X = c0*R + c1*G + c2*B + off for the pixel word px = {R, G, B, any}
template <int ROW>
__device__ __forceinline__ uint32_t dot_rgb(uint32_t px, int off)
{
    return (uint32_t)dp2a_hi(c5_coef[2 * ROW + 1], px, dp2a_lo(c5_coef[2 * ROW], px, off));
}
// quantised chroma + 40 (0..80) of the pixel word px: CB selects the Cb / Cr row of BT.601 (IMG:50-53)
template <bool CB>
__device__ __forceinline__ uint32_t chroma_q(uint32_t px)
{
    const uint32_t X = dot_rgb<CB ? 0 : 1>(px, (int)C5_OFF);
    const uint32_t t = __umulhi(X, C5_M);
    return __umulhi(lop3<0xEA>(t, 0xFFFFC000u, C5_Z), C5_M2);     // (t & mask) | Z
}
// luma in float32 exactly as rgb_to_value3 (IMG:47-49), quantised: Yq + QY_C0 (the caller subtracts the constant)
__device__ __forceinline__ uint32_t luma_q(float R, float G, float B)
{
    constexpr float T23 = 8388608.0f;
    const float y = __fadd_rn(__fadd_rn(__fmaf_rn(0.299f, R, -0.299f * T23), __fmaf_rn(0.587f, G, -0.587f * T23)), __fmaf_rn(0.114f, B, -0.114f * T23));
    const uint32_t by = (uint32_t)__float_as_int(__fadd_rd(__fadd_rd(y, 0.5f), T23 + (float)QY_Z));
    return __umulhi(by, QY_M);
}
// 13-trit value of the pixel word px = {R, G, B, any}, luma in float32 (exact for every colour).  Out of line: it runs for two
// pixels in a thousand, and the hot loop should stay small (instruction cache)
static __device__ __noinline__ uint32_t rgb_to_value5f(uint32_t px)
{
    return (luma_q(byte_magic(px, 0), byte_magic(px, 1), byte_magic(px, 2)) + (0u - QY_C0)) + 243u * chroma_q<true>(px) + 19683u * chroma_q<false>(px);
}
// The same with integer luma: Y = floor((299 R + 587 G + 114 B + 500) / 1000) equals the reference's float32 round(0.299f*R + 0.587f*G +
// 0.114f*B) (IMG:47-49) for every colour that is not an exact tie (the float32 error stays below 5e-5, the nearest non-tie is 1e-3 away);
// of the 16782 ties, 824 round down in float32.  t = floor(X * 2^9 / 1000): its low nine bits vanish for ties (and for X = 1 mod 1000);
// those pixels are flagged and recomputed by rgb_to_value5f.  Yq = (484 Y + 255) / 510 (IMG:71) as one multiply-high.
constexpr uint32_t Y5_M = 2199023256u;                          // ceil(2^41 / 1000)
constexpr uint32_t Y5_Z = 269u, Y5_M2 = 7961011u;               // hi32(((Y << 9) | Z) * M2) == (484 Y + 255) / 510, Y in [0, 255]
__device__ __forceinline__ uint32_t rgb_to_value5(uint32_t px, uint32_t& t)
{
    t = __umulhi(dot_rgb<2>(px, 500), Y5_M);
    const uint32_t yq = __umulhi(lop3<0xEA>(t, 0xFFFFFE00u, Y5_Z), Y5_M2);
    return yq + 243u * chroma_q<true>(px) + 19683u * chroma_q<false>(px);
}
// 26 bytes held in q[0..6] (q[6]: two bytes) -> shared memory at dst, which is 2-byte aligned: six 32-bit stores and one 16-bit store
// instead of thirteen 16-bit ones (a 16-bit store costs a full wavefront, and lanes 26 bytes apart collide on banks either way).
// PAR = 0 / 1: dst is known to be 0 / 2 mod 4;  PAR = 2: decided per lane (funnel shifts by 0 or 16 bits)
template <int PAR>
__device__ __forceinline__ void store26(uint8_t* dst, const uint32_t (&q)[7])
{
    if constexpr (PAR == 0) {
        uint32_t* d = reinterpret_cast<uint32_t*>(dst);
#pragma unroll
        for (int j = 0; j < 6; ++j) d[j] = q[j];
        *reinterpret_cast<uint16_t*>(dst + 24) = (uint16_t)q[6];
    } else if constexpr (PAR == 1) {
        *reinterpret_cast<uint16_t*>(dst) = (uint16_t)q[0];
        uint32_t* d = reinterpret_cast<uint32_t*>(dst + 2);
#pragma unroll
        for (int j = 0; j < 6; ++j) d[j] = __funnelshift_r(q[j], q[j + 1], 16);
    } else {
        const uint32_t a = smem_u32(dst), odd = a & 2u, sh = odd << 3;
        uint32_t* d = reinterpret_cast<uint32_t*>(dst + odd);
#pragma unroll
        for (int j = 0; j < 6; ++j) d[j] = __funnelshift_r(q[j], q[j + 1], sh);
        *reinterpret_cast<uint16_t*>(dst + (odd ? 0 : 24)) = (uint16_t)(odd ? q[0] : q[6]);
    }
}
// ---- encode phase A: six pixels (18 bytes at U + a) -> 26 stream symbols (x4) at dst
template <int PAR>
__device__ __forceinline__ void enc_unit_rgb5(const uint8_t* U, uint32_t a, uint8_t* dst)
{
    const uint32_t* mw = reinterpret_cast<const uint32_t*>(U + (a & ~3u));
    const uint32_t sh = (a & 3u) * 8u;
    uint32_t x[6];  // 18 bytes from any byte offset touch at most six words (the buffer has the slack)
#pragma unroll
    for (int j = 0; j < 6; ++j) x[j] = mw[j];
    uint32_t y[5]; // the 18 bytes, word aligned
#pragma unroll
    for (int j = 0; j < 5; ++j) y[j] = __funnelshift_r(x[j], x[j + 1], sh);
    // one word per pixel (fourth byte: don't care), so that every dot product uses the same two coefficient pairs
    const uint32_t px[6] = {y[0], __byte_perm(y[0], y[1], 0x6543), __byte_perm(y[1], y[2], 0x5432), __byte_perm(y[2], y[2], 0x0321), y[3], __byte_perm(y[3], y[4], 0x6543)};
    uint32_t A[6], t[6];
#pragma unroll
    for (int p = 0; p < 6; ++p) A[p] = rgb_to_value5(px[p], t[p]);
    if (!(t[0] & 511u) | !(t[1] & 511u) | !(t[2] & 511u) | !(t[3] & 511u) | !(t[4] & 511u) | !(t[5] & 511u)) { // rare: a luma tie among the six
#pragma unroll
        for (int p = 0; p < 6; ++p) if (!(t[p] & 511u)) A[p] = rgb_to_value5f(px[p]);
    }
    uint32_t w0, w1, w2, s12, v0, v1, v2, t12;
    triple_to_symbols(A[0], A[1], A[2], w0, w1, w2, s12);
    triple_to_symbols(A[3], A[4], A[5], v0, v1, v2, t12);
    // x4 (table byte offsets; symbols <= 26: no carry between bytes) as funnel shifts: phase A is bound by the multiply pipe, which
    // the compiler would otherwise also use for these shifts (IMAD.SHL)
    w0 = __funnelshift_l(0u, w0, 2); w1 = __funnelshift_l(0u, w1, 2); w2 = __funnelshift_l(0u, w2, 2); s12 = __funnelshift_l(0u, s12, 2);
    v0 = __funnelshift_l(0u, v0, 2); v1 = __funnelshift_l(0u, v1, 2); v2 = __funnelshift_l(0u, v2, 2); t12 = __funnelshift_l(0u, t12, 2);
    const uint32_t q[7] = {w0, w1, w2, s12 | (v0 << 8), __funnelshift_r(v0, v1, 24), __funnelshift_r(v1, v2, 24), (v2 >> 24) | (t12 << 8)};
    store26<PAR>(dst, q);
}
// one pass of phase A: unit u of the mini-tile; PAR as in store26 (S is 4-byte aligned, a unit is 26 bytes: PAR = u mod 2)
template <int K, bool WORDS, int PAR>
__device__ __forceinline__ void enc_unit5(const uint8_t* IN, uint32_t pad, uint8_t* S, int u)
{
    if constexpr (WORDS) enc_unit_words(IN, pad + 27u * (uint32_t)u, S + 26 * u);
    else enc_unit_rgb5<PAR>(IN, pad + 18u * (uint32_t)u, S + 26 * u);
}
template <int K, bool WORDS>
__device__ __forceinline__ void enc_phase_a5(const uint8_t* IN, uint32_t pad, uint8_t* S, int lane)
{
    using L = Cfg3<K>;
    static_assert(L::UNITS >= 64, "two full passes at least");
    // units dealt even / odd over the first two passes: 36- and 52-byte lane strides (9 and 13 words, conflict-free), and the pass is the
    // alignment of the unit's 26 symbols in S (compile-time store26)
    if constexpr (WORDS || T3C_ENC_ONE_LOOP) {   // the raw-word front end stores halfwords (no alignment to know) and is large: one copy of it
#pragma unroll 1
        for (int pass = 0; pass < L::PASS_A; ++pass) {
            const int u = pass < 2 ? 2 * lane + pass : 32 * pass + lane;
            if (u < L::UNITS) enc_unit5<K, WORDS, 2>(IN, pad, S, u);
        }
    } else {
        enc_unit5<K, WORDS, 0>(IN, pad, S, 2 * lane);
        enc_unit5<K, WORDS, 1>(IN, pad, S, 2 * lane + 1);
#pragma unroll 1
        for (int pass = 2; pass < L::PASS_A; ++pass) {
            const int u = 32 * pass + lane;
            if (u < L::UNITS) enc_unit5<K, WORDS, 2>(IN, pad, S, u);
        }
    }
}

// ---- shared-memory plan -----------------------------------------------------------------------------------------------------------
template <int K, bool WORDS = false> struct Cfg5 {
    using L = Cfg3<K>;
    static constexpr int PIX_BYTES = WORDS ? 9 * (L::PX / 2) : L::RGB_BYTES;          // pixel-side bytes of one mini-tile
    static constexpr int IN_BYTES = (PIX_BYTES + 15 + 15) / 16 * 16 + (WORDS ? 16 : 0); // with alignment slack (the word front end reads 32 bytes per unit)
    static constexpr int RUN_PITCH = 368;                                              // >= 15 + 338, multiple of 16
    static constexpr int RUNS_BYTES = 9 * RUN_PITCH;
    static constexpr int CARRY_BYTES = 9 * 16, BAR_BYTES = 16;
    static constexpr int WARP_BYTES = IN_BYTES + L::S_BYTES + RUNS_BYTES + CARRY_BYTES + BAR_BYTES;
    static constexpr int SMEM_MAX = 227 * 1024;
    // lane records of phase B: [tile mod 3][pass][lane] -> {source offset | destination offset << 16, band | variant << 8}
    static constexpr int REC_BYTES = 3 * 128 * 8;
    // encode, CTA-shared: per variant {A[K][27] | B[K][27]} | pat[3][2] (kept for reference: the kernel finds the pattern in position 0's entries) | records
    static constexpr int ENC_PLANE = 4 * K * 27, ENC_VAR = 2 * ENC_PLANE, ENC_PAT = 3 * ENC_VAR, ENC_REC = (ENC_PAT + 24 + 15) / 16 * 16;
    static constexpr int ENC_IMAGE = ENC_REC + REC_BYTES;                       // what the image holds
    static constexpr int ENC_WARP = (ENC_IMAGE + 127) / 128 * 128;              // per-warp blocks start 128-byte aligned (the tensor store reads U from there)
    static constexpr int ENC_WARP_BYTES = (WARP_BYTES + 127) / 128 * 128;       // encode: U | IN | S | carry | barrier
    static constexpr int ENC_WARPS = (SMEM_MAX - ENC_WARP) / ENC_WARP_BYTES < T3C_ENC_WARPS_CAP ? (SMEM_MAX - ENC_WARP) / ENC_WARP_BYTES : T3C_ENC_WARPS_CAP;
    static constexpr int TOTAL_ENC = ENC_WARP + ENC_WARPS * ENC_WARP_BYTES;
    // decode, CTA-shared (from a 256-byte aligned base): per variant {A[26][32] | B[26][32]} | chk[3][2] | par[3][2] | GF(27) + Chien tables | records
    static constexpr int DEC_PLANE = 4 * 26 * 32, DEC_VAR = 2 * DEC_PLANE, DEC_CHK = 3 * DEC_VAR, DEC_GF = (DEC_CHK + 48 + 15) / 16 * 16;   // chk[3][2] | par[3][2]
    static constexpr int CHIEN_BYTES = ((26 - K) / 2) * 27 * 24;                // the locator has at most t coefficients besides sigma_0
    static constexpr int DEC_REC = DEC_GF + ((int)sizeof(GfTables) + CHIEN_BYTES + 15) / 16 * 16;
    static constexpr int DEC_CLUT = DEC_REC + REC_BYTES;                        // dequantised chroma of the 81 quantised values, one byte each (dec_unit_rgb5)
    static constexpr int DEC_IMAGE = DEC_CLUT + 96;                             // what the image holds
    static constexpr int DEC_WARP = (DEC_IMAGE + 127) / 128 * 128;              // per-warp blocks start 128-byte aligned (tensor copies land there)
    static constexpr int DEC_WARP_BYTES = (WARP_BYTES + 127) / 128 * 128;       // decode: R | OUT | S | carry | barrier
    static constexpr int DEC_WARPS = (SMEM_MAX - 256 - DEC_WARP) / DEC_WARP_BYTES < 32 ? (SMEM_MAX - 256 - DEC_WARP) / DEC_WARP_BYTES : 32;
    static constexpr int TOTAL_DEC = 256 + DEC_WARP + DEC_WARPS * DEC_WARP_BYTES;
};
constexpr uint32_t REC_IDLE = 0xFFFFFFFFu;
// records from the variant-sorted pass maps (build_pass_map): after the maps are in `maps` (3 x 128 bytes) and a barrier
template <int K, int PITCH>
__device__ __forceinline__ void build_records(uint2* rec, const uint8_t* maps, const Geom& g, int t)
{
    if (t >= 3 * 128) return;
    const int tm = t >> 7;
    const uint32_t cw = maps[t];
    if (cw == 255) { rec[t] = make_uint2(0u, REC_IDLE); return; }
    const uint32_t cl = cw / 9u, b = cw - 9u * cl;
    const uint32_t v = ((uint32_t)(g.cw_base[b] % 3) + (uint32_t)tm + cl) % 3u;
    rec[t] = make_uint2((9u * K * cl + b) | ((PITCH * b + 26u * cl) << 16), b | (v << 8));
}

// Phase A is bound by the multiply pipe (IMAD / IDP), phase B by the logic pipe (LOP3 / PRMT).  Warps that start together stay in step
// and leave one of the two pipes idle most of the time (sm__pipe_fmaheavy 47 % + sm__pipe_alu 53 % of the elapsed cycles, back to back:
// profiles/r02b_*); every other warp of a sub-partition therefore starts about half a tile period late, which keeps the two groups
// in opposite phases (8K encode 160.8 -> 152.6 us).  flags = 1 | delay_cycles << 8 (T3C_V5_FLAGS overrides the default)
constexpr int V5_CR = 91881, V5_CB = 116130;         // 1.402, 1.772 x 2^16 (value_to_rgb5)
constexpr int V5_G1 = 1443411, V5_G2 = 2995303;      // 0.344136, 0.714136 x 2^22
constexpr uint32_t V5_FLAGS_DEFAULT = 1u | (9000u << 8);
__device__ __forceinline__ void stagger_start(uint32_t flags, int warp, uint32_t n_tiles)
{
    // flags & 8: four groups (warp / 4 mod 4) a quarter period apart instead of two (experiment)
    const uint32_t grp = (flags & 8u) ? ((uint32_t)warp >> 2) & 3u : ((uint32_t)warp >> 2) & 1u;
    if ((flags & 1u) && grp && n_tiles >= 8) {   // short ranges (chunked host pipelines) would only lose the delay
        const long long t0 = clock64(), d = (long long)grp * (long long)(flags >> 8);
        while (clock64() - t0 < d) { }
    }
}
// position of a warp inside its contiguous tile range: everything phase A / B / C need, advanced by constants
struct TileCursor {
    uint32_t f, tile, tm;      // frame, mini-tile inside the frame, tile mod 3
    uint64_t pix;              // global byte offset of the tile's pixel-side bytes
    uint64_t run;              // lanes 0..8: global byte offset of the band's run in the profile words
};
template <int PIX>
__device__ __forceinline__ void cursor_seek(TileCursor& c, const FastParams& P, const Geom& g, uint64_t pix_stride, uint64_t run_stride, uint32_t mt, int lane)
{
    c.f = mt / P.n_tiles;
    c.tile = P.tile0 + (mt - c.f * P.n_tiles);
    c.tm = c.tile % 3u;
    c.pix = pix_stride * c.f + (uint64_t)PIX * c.tile;
    c.run = run_stride * c.f + 52 + 26 * (g.cw_base[lane < 9 ? lane : 0] + (uint64_t)C_MINI * c.tile);
}
template <int PIX>
__device__ __forceinline__ void cursor_next(TileCursor& c, const FastParams& P, const Geom& g, uint64_t pix_stride, uint64_t run_stride, int lane)
{
    if (c.tile + 1 == P.tile0 + P.n_tiles) {   // next frame (rare)
        c.f += 1;
        c.tile = P.tile0;
        c.tm = c.tile % 3u;
        c.pix = pix_stride * c.f + (uint64_t)PIX * c.tile;
        c.run = run_stride * c.f + 52 + 26 * (g.cw_base[lane < 9 ? lane : 0] + (uint64_t)C_MINI * c.tile);
    } else {
        c.tile += 1;
        c.tm = c.tm == 2 ? 0 : c.tm + 1;
        c.pix += PIX;
        c.run += 26 * C_MINI;
    }
}

// .shared address of the first byte of dynamic shared memory for a kernel without static shared memory that is not launched in a
// cluster (sm_100a reserves the first KiB of the window).  The phase-B table look-ups below use ABSOLUTE addresses -- symbol*4 in
// the register, table block + position as the immediate -- because inside the (to the compiler possibly divergent) tile loop the
// window base would otherwise live in a vector register and cost one add per look-up.  smem_window_probe() checks the value on
// the device before the first v5 launch; the v4 kernels are used if it ever differs.
constexpr uint32_t SMEM_WINDOW_BASE = 0x400u;
template <uint32_t ABS>
__device__ __forceinline__ uint32_t lds_abs(uint32_t r)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(r), "n"(ABS));   // read-only tables: not volatile, the compiler may schedule these freely
    return v;
}
__global__ void k_smem_window_probe(uint32_t* out)
{
    extern __shared__ __align__(16) uint8_t smem[];
    if (threadIdx.x == 0) *out = smem_u32(smem);
}

// ---- one codeword of encode phase B (see enc_cw).  VAR >= 0: the codeword's scrambler variant is the compile-time VAR (variant-uniform
// pass, vb unused); VAR < 0: vb = variant * ENC_VAR, per lane.  TAB0 = offset of the variant-0 table block {A[K][27] | B[K][27]} from
// the start of dynamic shared memory
template <int K, int VAR, uint32_t TAB0, uint32_t ENC_VAR>
__device__ __forceinline__ void enc_cw5(const uint8_t* src, uint8_t* dst, uint32_t vb)
{
    constexpr int R = 26 - K;
    constexpr uint32_t PLANE = 4 * K * 27, BASE = SMEM_WINDOW_BASE + TAB0 + (VAR >= 0 ? VAR * ENC_VAR : 0u);
    Planes acc{0, 0}, acc2{0, 0};
    uint32_t e[K];
    static_for<0, K>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        const uint32_t d4 = src[9 * i];
        const uint32_t ra = VAR >= 0 ? d4 : d4 + vb;
        // plane B is the same in every variant but for position 0 (which carries the variant's parity scramble pattern): the mixed pass
        // reads the others from block 0, where all its lanes meet in one 27-word row
        const uint32_t ea = lds_abs<BASE + 108 * i>(ra), eb = lds_abs<(VAR >= 0 ? BASE : SMEM_WINDOW_BASE + TAB0) + 108 * i + PLANE>(i == 0 ? ra : d4);
        if (i == 0) acc = Planes{ea, eb}; else if (i == 1) acc2 = Planes{ea, eb};   // 0 + x = x: the first entry of an accumulator is a move
        else if (i & 1) gf3_add(acc2, ea, eb); else gf3_add(acc, ea, eb);
        e[i] = ea;
    });
    gf3_add(acc, acc2.nz, acc2.two);          // (the scramble pattern of the parity symbols came in with position 0's table entry)
    uint32_t lo, hi;
    planes_to_parity<K>(acc.nz, acc.two, lo, hi);
    // the codeword's 26 bytes as words: data symbols are the low bytes of the A entries
    uint32_t q[7];
#pragma unroll
    for (int j = 0; j < K / 4; ++j) q[j] = __byte_perm(__byte_perm(e[4 * j], e[4 * j + 1], 0x0040), __byte_perm(e[4 * j + 2], e[4 * j + 3], 0x0040), 0x5410);
    if constexpr (K % 4 == 0) {          // K = 20 / 24: parity starts on a word
        q[K / 4] = lo;
        if (K / 4 + 1 < 7) q[K / 4 + 1] = hi;
    } else {                              // K = 18 / 22: two data symbols, then parity
        const uint32_t t = __byte_perm(e[K - 2], e[K - 1], 0x0040);
        q[K / 4] = __byte_perm(t, lo, 0x5410);
        q[K / 4 + 1] = __byte_perm(lo, hi, 0x5432);
        if (K / 4 + 2 < 7) q[K / 4 + 2] = hi >> 16;
    }
    store26<2>(dst, q);
}

// ---- the CTA-shared image of the encoder (Cfg5: per variant {A[K][27] | B[K][27]} | pat[3][2] | lane records), built in the shared
// memory of a one-CTA setup kernel and copied out to the config's FastImageCache entry
template <int K>
__global__ void __launch_bounds__(256, 1) k_v5_image_enc(Geom g, const GfTables* __restrict__ gf, const RsTables* __restrict__ rs, uint8_t* __restrict__ image)
{
    using L = Cfg3<K>;
    using L5 = Cfg5<K, false>;
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint8_t maps[3 * 128];
    const int tid = threadIdx.x, TPB = blockDim.x;
    const uint32_t(*pl)[kVals][2] = rs->pl[g.arith][(24 - K) / 2];
    for (int idx = tid; idx < 3 * K * 27; idx += TPB) {
        const int v = idx / (K * 27), rem = idx - v * (K * 27), i = rem / 27, d = rem - 27 * i;
        uint32_t* blk = reinterpret_cast<uint32_t*>(smem + v * L5::ENC_VAR);
        Planes e{pl[i][d][0], pl[i][d][1]};
        if (i == 0) {   // the scrambler as seen by the parity symbols of a variant-v codeword rides on position 0: every codeword adds it exactly once
            uint32_t nz = 0, two = 0;
            for (int j = 0; j < L::R; ++j) {
                const uint32_t st = st_of(g, v, K + j);
                if (st) nz |= 7u << plane_shift<K>(j);
                if (st == 2) two |= 7u << plane_shift<K>(j);
            }
            gf3_add(e, nz, two);
        }
        blk[rem] = e.nz | gf->scr[st_of(g, v, i)][d];
        blk[K * 27 + rem] = e.two;
    }
    if (tid < 3) { // the scrambler as seen by the parity symbols of a variant-tid codeword, in the plane domain
        uint32_t nz = 0, two = 0;
        for (int j = 0; j < L::R; ++j) {
            const uint32_t st = st_of(g, tid, K + j);
            if (st) nz |= 7u << plane_shift<K>(j);
            if (st == 2) two |= 7u << plane_shift<K>(j);
        }
        reinterpret_cast<uint32_t*>(smem + L5::ENC_PAT)[2 * tid] = nz;
        reinterpret_cast<uint32_t*>(smem + L5::ENC_PAT)[2 * tid + 1] = two;
    }
    for (int t = tid; t < 3 * 128; t += TPB) build_pass_map(maps, g, t);
    __syncthreads();
    for (int t = tid; t < 3 * 128; t += TPB) build_records<K, L5::RUN_PITCH>(reinterpret_cast<uint2*>(smem + L5::ENC_REC), maps, g, t);
    __syncthreads();
    for (int i = tid; i < L5::ENC_IMAGE / 16; i += TPB) reinterpret_cast<uint4*>(image)[i] = reinterpret_cast<const uint4*>(smem)[i];
}
template <int K>
__global__ void __launch_bounds__(256, 1) k_v5_image_dec(Geom g, const GfTables* __restrict__ gf, const RsTables* __restrict__ rs, uint8_t* __restrict__ image)
{
    using L5 = Cfg5<K, false>;
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint8_t maps[3 * 128];
    const int tid = threadIdx.x, TPB = blockDim.x;
    GfTables& sg = *reinterpret_cast<GfTables*>(smem + L5::DEC_GF);
    const uint32_t(*pl)[kVals][2] = rs->pl[1][(24 - K) / 2]; // the consistent decoder always uses the repaired code
    for (int idx = tid; idx < 3 * 26 * 32; idx += TPB) {
        const int v = idx / (26 * 32), rem = idx - v * (26 * 32), i = rem / 32, x = rem - 32 * i, xm = x >= 27 ? x - 27 : x;
        uint32_t* blk = reinterpret_cast<uint32_t*>(smem + v * L5::DEC_VAR);
        blk[rem] = pl[i][xm][0] | gf->dsc[st_of(g, v, i)][xm];
        blk[26 * 32 + rem] = pl[i][xm][1];
    }
    load_gf(sg, gf);
    for (int i = tid; i < L5::CHIEN_BYTES / 4; i += TPB) reinterpret_cast<uint32_t*>(&sg + 1)[i] = chien_of(gf)[i];
    for (int t = tid; t < 3 * 128; t += TPB) build_pass_map(maps, g, t);
    if (tid < 96) smem[L5::DEC_CLUT + tid] = (uint8_t)min((32u * (uint32_t)tid + 5u) / 10u, 255u);   // dequantize_ycbcr's chroma (IMG:79-84) of Cq + 40 = tid
    __syncthreads();
    for (int t = tid; t < 3 * 128; t += TPB) build_records<K, L5::RUN_PITCH>(reinterpret_cast<uint2*>(smem + L5::DEC_REC), maps, g, t);
    if (tid < 3) { // a received block r = c (+) 13*st is a codeword iff sum_i T_i[r_i] == sum_i T_i[13*st_i] (GF(3)-linear tables)
        Planes c{0, 0};
        const uint32_t* blk = reinterpret_cast<const uint32_t*>(smem + tid * L5::DEC_VAR);
        for (int i = 0; i < 26; ++i) {
            const int idx = i * 32 + 13 * (int)st_of(g, tid, i);
            gf3_add(c, blk[idx] & ~0xFFu, blk[26 * 32 + idx]);
        }
        reinterpret_cast<uint32_t*>(smem + L5::DEC_CHK)[2 * tid] = c.nz;
        reinterpret_cast<uint32_t*>(smem + L5::DEC_CHK)[2 * tid + 1] = c.two;
        // the hot path sums the K data positions only and compares the parity it implies with the received parity symbols as bytes:
        // sum_{i<K} T_i[r_i] = parity(c) + D with D = sum_{i<K} T_i[13*st_i], the received parity symbol j of a codeword is parity(c)_j (+)
        // 13*st_{K+j}, so the constant to add is par = (scrambler pattern of the parity positions) - D
        Planes e{0, 0};
        for (int i = 0; i < K; ++i) {
            const int idx = i * 32 + 13 * (int)st_of(g, tid, i);
            gf3_add(e, blk[idx] & ~0xFFu, blk[26 * 32 + idx]);
        }
        e.two ^= e.nz;                                          // -D
        uint32_t pn = 0, pt = 0;
        for (int j = 0; j < 26 - K; ++j) {
            const uint32_t st = st_of(g, tid, K + j);
            if (st) pn |= 7u << plane_shift<K>(j);
            if (st == 2) pt |= 7u << plane_shift<K>(j);
        }
        gf3_add(e, pn, pt);
        reinterpret_cast<uint32_t*>(smem + L5::DEC_CHK)[6 + 2 * tid] = e.nz;
        reinterpret_cast<uint32_t*>(smem + L5::DEC_CHK)[6 + 2 * tid + 1] = e.two;
    }
    __syncthreads();
    // par rides on position 0's table entries: every codeword's data sum then already holds it (no add in the hot path, six registers
    // free); the out-of-line path, which needs the plain 26-position sum, takes it off again (dec_cw_dirty5)
    if (tid < 3 * 32) {
        const int v = tid >> 5, x = tid & 31;
        uint32_t* blk = reinterpret_cast<uint32_t*>(smem + v * L5::DEC_VAR);
        const uint32_t* par = reinterpret_cast<const uint32_t*>(smem + L5::DEC_CHK) + 6 + 2 * v;
        Planes e{blk[x], blk[26 * 32 + x]};
        gf3_add(e, par[0], par[1]);
        blk[x] = e.nz;
        blk[26 * 32 + x] = e.two;
    }
    __syncthreads();
    for (int i = tid; i < L5::DEC_IMAGE / 16; i += TPB) reinterpret_cast<uint4*>(image)[i] = reinterpret_cast<const uint4*>(smem)[i];
}

template <int K, bool WORDS>
__global__ void __launch_bounds__(32 * Cfg5<K, WORDS>::ENC_WARPS, 1) k_encode_v5(FastParams P, Geom g, const GfTables* __restrict__ gf, const uint8_t* __restrict__ image,
                                                                                 const __grid_constant__ CUtensorMap tmap)
{
    using L = Cfg3<K>;
    using L5 = Cfg5<K, WORDS>;
    constexpr int PIX = L5::PIX_BYTES, NW = L5::ENC_WARPS, TPB = 32 * NW, PITCH = L5::RUN_PITCH;
    extern __shared__ __align__(16) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* U = smem + L5::ENC_WARP + warp * L5::ENC_WARP_BYTES; // the nine body runs (128-byte aligned: the tensor store reads them from here)
    uint8_t* IN = U + L5::RUNS_BYTES;                           // the pixel side of the tile (bulk-loaded one tile ahead)
    uint8_t* S = IN + L5::IN_BYTES;                             // stream symbols, pre-scaled by 4
    uint4* carry = reinterpret_cast<uint4*>(S + L::S_BYTES);
    const uint32_t bar = smem_u32(S + L::S_BYTES + L5::CARRY_BYTES);
    const uint2* rec = reinterpret_cast<const uint2*>(smem + L5::ENC_REC);
    {   // the CTA-shared tables: a copy of the image v5_build_enc made once for this config (FastImageCache)
        const uint4* src = reinterpret_cast<const uint4*>(image);
        uint4* dst = reinterpret_cast<uint4*>(smem);
        for (int i = tid; i < L5::ENC_IMAGE / 16; i += TPB) dst[i] = __ldg(src + i);
        if (lane == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    }
    __syncthreads(); // tables, records and barriers ready; no block-level barrier after this one
    const uint64_t in_limit = P.in_stride * P.n_frames;
    uint32_t mt_lo, mt_hi;
    warp_range_smsp((uint64_t)P.n_tiles * P.n_frames, blockIdx.x, gridDim.x, warp, NW, mt_lo, mt_hi);
    if (mt_lo >= mt_hi) return;
    // bulk load of a tile's pixel side: the 16-byte aligned superset of [pix, pix + PIX), clipped to the buffer
    auto fetch = [&](uint64_t pix) {
        const uint64_t a0 = pix & ~15ull;
        uint32_t bytes = (uint32_t)((pix - a0) + PIX + 15) & ~15u;
        if (a0 + bytes > in_limit) bytes = (uint32_t)(in_limit - a0) & ~15u;
        fence_async_smem();
        mbar_expect_tx(bar, bytes);
        bulk_g2s(smem_u32(IN), P.in + a0, bytes, bar);
    };
    TileCursor c;
    cursor_seek<PIX>(c, P, g, P.in_stride, P.out_stride, mt_lo, lane);
    if (lane == 0) fetch(c.pix);
    stagger_start(P.flags, warp, mt_hi - mt_lo);
    uint32_t phase = 0;
    bool first = true;
    for (uint32_t left = mt_hi - mt_lo; left; --left) {
        const bool last = left == 1 || c.tile + 1 == P.tile0 + P.n_tiles;           // of a contiguous stretch
        const uint32_t pad = (uint32_t)c.pix & 15u, padb = (uint32_t)c.run & 15u;     // 16-byte phase of the pixel run / of lane b's band run
        mbar_wait(bar, phase);
        phase ^= 1;
        if (last) {   // the last bytes of the buffer that a clipped bulk copy left out (only the final tile of the final frame)
            const uint64_t a0 = c.pix - pad;
            const uint32_t want = pad + PIX;
            if (a0 + ((want + 15) & ~15u) > in_limit)
                for (uint32_t i = ((uint32_t)(in_limit - a0) & ~15u) + lane; i < want; i += 32) IN[i] = a0 + i < in_limit ? P.in[a0 + i] : 0;
            __syncwarp();
        }
        enc_phase_a5<K, WORDS>(IN, pad, S, lane);
        __syncwarp();
        TileCursor nx = c;
        cursor_next<PIX>(nx, P, g, P.in_stride, P.out_stride, lane);
        if (left > 1 && lane == 0) fetch(nx.pix);                                  // IN is free again: next tile's pixels on their way
        // (advancing the cursor in place at the end of the loop, as the decoder does, frees seven registers across phase B -- the kernel
        // then compiles to 64 registers -- and was measured slower: 140.2 us against 138.3 us)
        if (lane < 9) {
#if T3C_ENC_WAIT_READ
            bulk_wait_read();
#else
            bulk_wait_all();                                                         // the previous tile's stores have read U and have landed (see phase C)
#endif
            if (!first) *reinterpret_cast<uint4*>(U + PITCH * lane) = carry[lane];   // bytes [0, padb) of each run: the previous tile's tail
        }
        __syncwarp();
        // ---- phase B: one codeword per lane and pass; passes 0..2 are variant-uniform, pass 3 takes the leftovers of all three
        {
            const uint2* rt = rec + 128 * c.tm;
            static_for<0, 3>([&](auto pc) {
                constexpr int p = decltype(pc)::value;
                const uint2 r = rt[32 * p + lane];
                const uint32_t pb = __shfl_sync(0xFFFFFFFFu, padb, (int)(r.y & 0xFFu));
                enc_cw5<K, p, 0u, (uint32_t)L5::ENC_VAR>(S + (r.x & 0xFFFFu), U + (r.x >> 16) + pb, 0u);
            });
            const uint2 r = rt[96 + lane];
            const uint32_t pb = __shfl_sync(0xFFFFFFFFu, padb, (int)(r.y & 0xFu));
            if (r.y != REC_IDLE) {
                const uint32_t v = r.y >> 8;
                enc_cw5<K, -1, 0u, (uint32_t)L5::ENC_VAR>(S + (r.x & 0xFFFFu), U + (r.x >> 16) + pb, v * L5::ENC_VAR);
            }
        }
        __syncwarp();
        if (c.tile == 0 && lane == 0 && g.cw_base[0] == 0) { // body symbols 0 and 1 may still see the scrambler's transient (A.4)
            uint8_t* dst = U + padb;
            dst[0] = gf->scr[g.st[0]][S[0] >> 2];
            dst[1] = gf->scr[g.st[1]][S[9] >> 2];
        }
        fence_async_smem();                                                          // generic-proxy writes to U before the bulk engine reads it
        __syncwarp();
        // ---- phase C: nine band-major runs -> global as bulk stores of whole chunks; the partial last chunk is carried to the next tile
        // Inside a stretch of a regular frame (launch_v5_enc: every band's run has the same 16-byte phase) the nine rows of U leave by ONE
        // 3-D tensor store (SASS UTMASTG) of all 23 chunks per row: the chunks behind this tile's last whole one hold its tail and stale
        // bytes, and the same warp's next tile stores over them (its row starts with this tile's tail, carried) -- after this store has
        // landed: bulk_wait_all above.  The first and the last tile of a stretch store exactly their own chunks, run by run.
        const bool ts = !first && !last && c.tile < P.ts_tiles;     // ts_tiles: 0 without the tensor store, else the tiles whose box stays inside a band's row
        if (lane < 9) {
            const uint32_t cend = (padb + L::RUN) >> 4, c0 = (first && padb) ? 1u : 0u;
            if (!last) carry[lane] = *reinterpret_cast<const uint4*>(U + PITCH * lane + 16 * cend);
            if (!ts) {
                if (cend > c0) bulk_s2g(P.out + (c.run - padb) + 16 * c0, smem_u32(U + PITCH * lane + 16 * c0), 16 * (cend - c0));
                bulk_commit();
            }
        }
        if (ts && lane == 0) {
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                         ::"l"(&tmap), "r"((int)(((52u + 338u * c.tile) & ~15u) >> 1)), "r"(0), "r"((int)c.f), "r"(smem_u32(U)) : "memory");   // x in 16-bit elements
            bulk_commit();
        }
        if (first || last) { // edge bytes of a contiguous stretch, one by one
#pragma unroll 1
            for (int b = 0; b < 9; ++b) {
                const uint64_t lo = __shfl_sync(0xFFFFFFFFu, c.run, b);
                const int pb = (int)(lo & 15), end = pb + L::RUN, cend = end >> 4;
                const uint8_t* s0 = U + PITCH * b;
                uint8_t* g0 = P.out + (lo - pb);
                if (first && pb && lane >= pb && lane < 16) g0[lane] = s0[lane];
                if (last && lane < 16 && 16 * cend + lane < end) g0[16 * cend + lane] = s0[16 * cend + lane];
            }
        }
        __syncwarp();
        first = last;   // a stretch ends at a frame boundary: the next tile starts a new one
        c = nx;
    }
    if (lane < 9) bulk_wait_all(); // shared memory must outlive the copies that read it
}

// =============================================================================================
// decode
// =============================================================================================
// The rare branches of a codeword's screen live out of line: inlined into the four unrolled passes they made the hot loop 100 KB of
// code (6350 instructions), more than the instruction caches hold for 27 warps at different places of it.
// bytes >= 27 in the received words (out-of-alphabet symbols read as their low three trits, unpack3, OLD:28-31).
// The out-of-line paths take .shared addresses (32 bits), not pointers: the hot path then never forms the 64-bit generic addresses
// that pointer arguments of a call it rarely makes would need before every branch (ten instructions per codeword pass).
__device__ __forceinline__ uint8_t* smem_ptr(uint32_t a) { return reinterpret_cast<uint8_t*>(__cvta_shared_to_generic(a)); }
static __device__ __noinline__ void dec_cw_mod27(uint32_t src_s)
{
    uint8_t* src = smem_ptr(src_s);
    for (int i = 0; i < 26; ++i) src[i] = (uint8_t)(src[i] % 27u);   // in place in the staged run: the codeword belongs to this lane alone
}
// a codeword whose received parity differs from the parity its data symbols imply: finish the syndrome screen with the R parity
// positions (table rows K..25 hold -x in the planes), subtract the clean-codeword constant -- that is the parity residual, from which
// the bounded-distance decoder (dev.cuh rs_bd_fix) repairs the data symbols the hot path has already stored
template <int K>
static __device__ __noinline__ void dec_cw_dirty5(uint32_t src_s, uint32_t dst_s, uint32_t acc_nz, uint32_t acc_two, uint32_t tab_s, uint32_t chk_s, uint32_t sg_s, uint32_t* status)
{
    constexpr int PLANE = 26 * 32;
    const uint8_t* src = smem_ptr(src_s);
    const uint32_t* tab_v = reinterpret_cast<const uint32_t*>(smem_ptr(tab_s));
    const uint32_t* chk_v = reinterpret_cast<const uint32_t*>(smem_ptr(chk_s));
    const GfTables& sg = *reinterpret_cast<const GfTables*>(smem_ptr(sg_s));
    Planes d{acc_nz, acc_two};
#pragma unroll 1
    for (int i = K; i < 26; ++i) {
        const uint32_t* row = tab_v + 32 * i + src[i];             // src[i] < 32 here (dec_cw_mod27 ran if any byte was larger)
        gf3_add(d, row[0], row[PLANE]);
    }
    const uint32_t cn = chk_v[0], ct = chk_v[1], pn = chk_v[6], pt = chk_v[7];   // par of this variant lies 24 bytes behind its chk
    gf3_add(d, cn, cn ^ ct);                                   // minus the constant: -x keeps nz and flips two where nz is set
    gf3_add(d, pn, pn ^ pt);                                   // minus par, which the data sum brought along from position 0's entries
    if (!(d.nz >> 8)) return;                                  // a parity byte 27..31 (alias of 0..4) in an otherwise clean codeword
    uint32_t lo, hi;
    planes_to_parity<K>(d.nz, d.two, lo, hi);
    rs_bd_fix<K>(sg, chien_of(&sg), smem_ptr(dst_s), lo, hi, status, true);   // the image keeps the Chien tables behind the GF(27) tables, as HostTables does
}
// ---- one codeword of decode phase B: 26 received symbols at .shared address sa (even) -> K descrambled data symbols scattered at byte
// stride 9 from .shared address da, and the screen: the parity the K data symbols imply (K table look-ups, plane sums), scrambled in the
// plane domain (the constant par of k_v5_image_dec, which rides on position 0's table entries), converted to bytes and compared with the R received parity symbols as they lie in the run --
// 6 look-ups, 18 LOP3 and 6 PRMT fewer per codeword than the full 26-position syndrome sum, which only dirty codewords finish
// (dec_cw_dirty5).  pa = .shared address of the variant's table block, chk_s / sg_s = .shared addresses of its clean-codeword constant
// and of the GF(27) tables
template <int K>
__device__ __forceinline__ void dec_cw5(uint32_t sa, uint32_t da, uint32_t pa, uint32_t chk_s, uint32_t sg_s, uint32_t* status)
{
    constexpr int PLANE = 4 * 26 * 32, W = K / 4, NW = (K + 3) / 4;
    asm volatile("" : "+r"(pa));   // the block address in a vector register: with a uniform one PRMT would need its selector in a register (a move per symbol)
    const uint32_t sh = (sa & 2u) * 8u;
    uint32_t xw[7];
    auto load = [&]() {
        static_for<0, 7>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(xw[j]) : "r"(sa & ~3u), "n"(4 * j) : "memory");
        });
#pragma unroll
        for (int j = 0; j < 6; ++j) xw[j] = __funnelshift_r(xw[j], xw[j + 1], sh);
        xw[6] = (xw[6] >> sh) & 0xFFFFu;                           // symbols 24, 25 only
    };
    load();
    if ((xw[0] | xw[1] | xw[2] | xw[3] | xw[4] | xw[5] | xw[6]) & 0xE0E0E0E0u) { dec_cw_mod27(sa); load(); }
    // the received parity symbols K..25 as bytes of two words
    uint32_t rx_lo, rx_hi = 0;
    if constexpr (K % 4 == 0) {
        rx_lo = xw[W];
        if constexpr (W + 1 < 7) rx_hi = xw[W + 1];
    } else {
        rx_lo = __funnelshift_r(xw[W], xw[W + 1], 16);
        if constexpr (W + 2 < 7) rx_hi = __funnelshift_r(xw[W + 1], xw[W + 2], 16);
    }
#pragma unroll
    for (int j = 0; j < NW; ++j) xw[j] *= 4u;                      // table byte offsets; < 128 per byte: no carry between symbols
    Planes acc{0, 0}, acc2{0, 0};
    uint32_t ev[K];
    static_for<0, K>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        const uint32_t ra = __byte_perm(xw[i >> 2], pa, 0x7650u | (uint32_t)(i & 3));
        const uint32_t ea = lds_tab<128 * i>(ra);
        const uint32_t eb = lds_tab<128 * i + PLANE>(ra);
        if (i == 0) acc = Planes{ea, eb}; else if (i == 1) acc2 = Planes{ea, eb};   // 0 + x = x: the first entry of an accumulator is a move
        else if (i & 1) gf3_add(acc2, ea, eb); else gf3_add(acc, ea, eb);
        ev[i] = ea;
    });
    static_for<0, K>([&](auto ic) {                                // stores after all loads: nothing to order
        constexpr int i = decltype(ic)::value;
        asm volatile("st.shared.u8 [%0+%1], %2;" ::"r"(da), "n"(9 * i), "r"(ev[i]) : "memory");
    });
    gf3_add(acc, acc2.nz, acc2.two);          // (par came in with position 0's table entry)
    uint32_t lo, hi;
    planes_to_parity<K>(acc.nz, acc.two, lo, hi);
    if ((26 - K > 4) ? (((lo ^ rx_lo) | (hi ^ rx_hi)) != 0u) : (lo != rx_lo))
        dec_cw_dirty5<K>(sa, da, acc.nz, acc.two, pa, chk_s, sg_s, status);
}

// pixel value -> RGB8 (decode_raw_words_to_pixels + dequantize_ycbcr + ycbcr_to_rgb, OLD:706-722, IMG:57-84) in integers.  The
// reference's float32 chain  r = y + 1.402 cr,  g = (y - 0.344136 cb) - 0.714136 cr,  b = y + 1.772 cb  followed by round-half-away
// and the clamp is reproduced for every one of the 243 x 81 x 81 dequantised (Y, Cb, Cr) by fixed-point sums: 16 fraction bits for r and
// b (their exact fractions are multiples of 0.002 / 0.004, never .5 for r; b has one exact tie, cb = -125, which the +32 keeps on the
// round-half-away side), 22 bits for g (closest approach to a tie 5.6e-5).  tests/test_host_logic.py checks all 1 594 323 values
// against the float32 chain.  No conversions, no F2I (quarter-rate pipe), and the clamped r / b leave through byte 2 of their sums.
__device__ __forceinline__ uint32_t yuvq_to_rgb5(uint32_t Yq, uint32_t ub, uint32_t ur)   // Yq <= 242, ub = Cbq + 40, ur = Crq + 40 <= 80
{
    const int Y = (int)__umulhi(Yq * 510u + 241u, 8873899u);
    // (96 u + 15) / 30 rather than (32 u + 5) / 10: a multiplier that is not a power of two stays on the multiply pipe (the logic pipe is the busy one)
    const int Cb = (int)min(__umulhi(96u * ub + 15u, 143165577u), 255u);
    const int Cr = (int)min(__umulhi(96u * ur + 15u, 143165577u), 255u);
    const int Y16 = Y << 16, Y22 = Y << 22;
    // add the constant, clamp to [0, max] in one instruction each (VIADDMNMX.RELU)
    const uint32_t r = (uint32_t)__viaddmin_s32_relu(Cr * V5_CR + Y16, 32768 - 128 * V5_CR, 0xFFFFFF);
    const uint32_t b = (uint32_t)__viaddmin_s32_relu(Cb * V5_CB + Y16, 32768 + 32 - 128 * V5_CB, 0xFFFFFF);
    const uint32_t g = (uint32_t)__viaddmin_s32_relu(Cr * -V5_G2 + (Cb * -V5_G1 + Y22), 2097152 + 128 * (V5_G1 + V5_G2), 0x3FFFFFFF) >> 22;
    return __byte_perm(__byte_perm(r, g, 0x3042), b, 0x3610);                  // R | G<<8 | B<<16 (byte 3 of the clamped r is 0)
}
__device__ __forceinline__ uint32_t value_to_rgb5(uint32_t A)
{
    const uint32_t q = __umulhi(A, 17674763u);                 // A / 243, exact for A < 3^13
    const uint32_t Yq = A - 243u * q;
    const uint32_t ur = __umulhi(q, 53024288u);                // q / 81, exact for q < 6561
    return yuvq_to_rgb5(Yq, q - 81u * ur, ur);
}
// six 13-trit pixel values -> their 18 RGB bytes as words (w[4]: two bytes): dequantised chroma from the 81-byte table, exact fixed-point
// colour sums (yuvq_to_rgb5), 13 PRMT gather the bytes from the clamped sums (r and b in byte 2, g in byte 0)
__device__ __forceinline__ void values_to_rgb18(const uint32_t (&A)[6], const uint8_t* __restrict__ clut, uint32_t (&w)[5])
{
    uint32_t r[6], g[6], b[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const uint32_t qq = __umulhi(A[q], 17674763u);             // A / 243, exact for A < 3^13
        const uint32_t Yq = A[q] - 243u * qq;
        const uint32_t ur = __umulhi(qq, 53024288u);               // / 81, exact below 6561
        const int Cb = clut[qq - 81u * ur], Cr = clut[ur];
        const int Y = (int)__umulhi(Yq * 510u + 241u, 8873899u);
        const int Y16 = Y << 16, Y22 = Y << 22;
        r[q] = (uint32_t)__viaddmin_s32_relu(Cr * V5_CR + Y16, 32768 - 128 * V5_CR, 0xFFFFFF);
        b[q] = (uint32_t)__viaddmin_s32_relu(Cb * V5_CB + Y16, 32768 + 32 - 128 * V5_CB, 0xFFFFFF);
        g[q] = (uint32_t)__viaddmin_s32_relu(Cr * -V5_G2 + (Cb * -V5_G1 + Y22), 2097152 + 128 * (V5_G1 + V5_G2), 0x3FFFFFFF) >> 22;
    }
    auto rg = [&](int q) { return __byte_perm(r[q], g[q], 0x0042); };           // R | G << 8
    auto gb = [&](int q) { return __byte_perm(g[q], b[q], 0x0060); };           // G | B << 8
    auto br = [&](int q) { return __byte_perm(b[q], r[q + 1], 0x0062); };       // B | R' << 8
    w[0] = __byte_perm(rg(0), br(0), 0x5410); w[1] = __byte_perm(gb(1), rg(2), 0x5410); w[2] = __byte_perm(br(2), gb(3), 0x5410);
    w[3] = __byte_perm(rg(4), br(4), 0x5410); w[4] = gb(5);
}
// ---- decode phase A: 26 stream symbols at S + a (even) -> six pixels -> 18 RGB bytes at dst (even address): four 32-bit stores and one
// 16-bit store, aligned per lane by funnel shifts (see store26).  PARU = 0 / 1: a is 0 / 2 mod 4, known at compile time (the alignment
// shifts of the loads vanish or become immediates); PARU = 2: decided per lane.  clut: the 81 dequantised chroma values as bytes -- 21
// consecutive words, so a look-up never meets a bank conflict -- instead of a multiply, a multiply-high and a min per component (the
// multiply pipe bounds this phase: a multiply-high takes it for four cycles, tools/probe/pipe_rates.cu).  The 18 bytes are gathered
// from the clamped sums (r and b in byte 2, g in byte 0) by 13 PRMT.
template <int PARU>
__device__ __forceinline__ void dec_unit_rgb5(const uint8_t* S, uint32_t a, uint8_t* dst, const uint8_t* __restrict__ clut)
{
    const uint32_t* mw = reinterpret_cast<const uint32_t*>(S + (a & ~3u));
    uint32_t x[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) x[j] = mw[j];
    uint32_t y0, y1, y2, s12, z0, z1, z2, t12;   // symbols 0..11, 12, 13..24, 25
    if constexpr (PARU == 0) {
        y0 = x[0]; y1 = x[1]; y2 = x[2]; s12 = x[3] & 0xFFu;
        z0 = __funnelshift_r(x[3], x[4], 8); z1 = __funnelshift_r(x[4], x[5], 8); z2 = __funnelshift_r(x[5], x[6], 8); t12 = (x[6] >> 8) & 0xFFu;
    } else if constexpr (PARU == 1) {
        y0 = __funnelshift_r(x[0], x[1], 16); y1 = __funnelshift_r(x[1], x[2], 16); y2 = __funnelshift_r(x[2], x[3], 16); s12 = (x[3] >> 16) & 0xFFu;
        z0 = __funnelshift_r(x[3], x[4], 24); z1 = __funnelshift_r(x[4], x[5], 24); z2 = __funnelshift_r(x[5], x[6], 24); t12 = x[6] >> 24;
    } else {
        const uint32_t sh = (a & 2u) * 8u, sh8 = sh + 8u;
        y0 = __funnelshift_r(x[0], x[1], sh); y1 = __funnelshift_r(x[1], x[2], sh); y2 = __funnelshift_r(x[2], x[3], sh); s12 = (x[3] >> sh) & 0xFFu;
        z0 = __funnelshift_r(x[3], x[4], sh8); z1 = __funnelshift_r(x[4], x[5], sh8); z2 = __funnelshift_r(x[5], x[6], sh8); t12 = (x[6] >> sh8) & 0xFFu;
    }
    uint32_t A[6], w[5];
    symbols_to_triple(y0, y1, y2, s12, A[0], A[1], A[2]);
    symbols_to_triple(z0, z1, z2, t12, A[3], A[4], A[5]);
    values_to_rgb18(A, clut, w);
    const uint32_t da = smem_u32(dst), odd = da & 2u, sh = odd << 3;
    uint32_t* d = reinterpret_cast<uint32_t*>(dst + odd);
#pragma unroll
    for (int j = 0; j < 4; ++j) d[j] = __funnelshift_r(w[j], w[j + 1], sh);
    *reinterpret_cast<uint16_t*>(dst + (odd ? 0 : 16)) = (uint16_t)(odd ? w[0] : w[4]);
}
template <int K>
__device__ __forceinline__ void dec_phase_a5(const uint8_t* S, uint8_t* OUT, uint32_t pad, int lane, const uint8_t* __restrict__ clut)
{
    using L = Cfg3<K>;
    static_assert(L::PASS_A >= 3, "two full passes at least");
    // units dealt even / odd over the first two passes (S is 4-byte aligned and a unit has 26 bytes: the pass is the alignment)
    dec_unit_rgb5<0>(S, 52u * (uint32_t)lane, OUT + pad + 36 * lane, clut);               // pad is even: every frame starts on a 16-byte boundary and 3*PX*tile is even
    dec_unit_rgb5<1>(S, 52u * (uint32_t)lane + 26u, OUT + pad + 36 * lane + 18, clut);
#pragma unroll 1
    for (int pass = 2; pass < L::PASS_A; ++pass) {
        const int u = 32 * pass + lane;
        if (u < L::UNITS) dec_unit_rgb5<2>(S, 26u * (uint32_t)u, OUT + pad + 18 * u, clut);
    }
}

template <int K, bool WORDS>
__global__ void __launch_bounds__(32 * Cfg5<K, WORDS>::DEC_WARPS, 1) k_decode_v5(FastParams P, Geom g, const uint8_t* __restrict__ image, const __grid_constant__ CUtensorMap tmap)
{
    using L = Cfg3<K>;
    using L5 = Cfg5<K, WORDS>;
    constexpr int PIX = L5::PIX_BYTES, NW = L5::DEC_WARPS, TPB = 32 * NW, PITCH = L5::RUN_PITCH;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((256u - (smem_u32(smem_raw) & 255u)) & 255u); // 256-byte aligned: PRMT drops a symbol (x4) into the low address byte
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* R = smem + L5::DEC_WARP + warp * L5::DEC_WARP_BYTES; // the nine body runs as they lie in the frame (loaded one tile ahead); 128-byte aligned
    uint8_t* OUT = R + L5::RUNS_BYTES;                          // the pixel side of the tile on its way out
    uint8_t* S = OUT + L5::IN_BYTES;                            // descrambled stream symbols (plain)
    uint4* carry = reinterpret_cast<uint4*>(S + L::S_BYTES);
    const uint32_t bar = smem_u32(S + L::S_BYTES + L5::CARRY_BYTES);
    const GfTables& sg = *reinterpret_cast<const GfTables*>(smem + L5::DEC_GF);
    const uint2* rec = reinterpret_cast<const uint2*>(smem + L5::DEC_REC);
    {   // the CTA-shared tables: a copy of the image k_v5_image_dec made once for this config (FastImageCache)
        const uint4* src = reinterpret_cast<const uint4*>(image);
        uint4* dst = reinterpret_cast<uint4*>(smem);
        for (int i = tid; i < L5::DEC_IMAGE / 16; i += TPB) dst[i] = __ldg(src + i);
        if (lane == 0) { mbar_init(bar, 9); fence_mbar_init(); }
    }
    __syncthreads();
    const uint32_t tabA32 = smem_u32(smem), R32 = smem_u32(R), S32 = smem_u32(S), chk32 = tabA32 + L5::DEC_CHK, sg32 = tabA32 + L5::DEC_GF;
    const uint32_t* chk = reinterpret_cast<const uint32_t*>(smem + L5::DEC_CHK);
    const uint64_t in_limit = P.in_stride * (P.n_frames - 1) + 9 * g.n_out;
    uint32_t mt_lo, mt_hi;
    warp_range_smsp((uint64_t)P.n_tiles * P.n_frames, blockIdx.x, gridDim.x, warp, NW, mt_lo, mt_hi);
    if (mt_lo >= mt_hi) return;
    // lane b < 9 bulk-loads band b's run: the 16-byte aligned superset of the run, clipped to the buffer
    auto fetch = [&](uint64_t lo) {
        const uint64_t a0 = lo & ~15ull;
        uint32_t bytes = (uint32_t)((lo - a0) + L::RUN + 15) & ~15u;
        if (a0 + bytes > in_limit) bytes = (uint32_t)(in_limit - a0) & ~15u;
        fence_async_smem();
        mbar_expect_tx(bar, bytes);
        if (bytes) bulk_g2s(smem_u32(R + PITCH * lane), P.in + a0, bytes, bar);
    };
    // the same nine (16-byte aligned supersets of the) runs by ONE 3-D tensor copy when the frame's runs form a regular box
    // (launch_v5_dec): same layout in R as the bulk copies give; all nine lanes arrive on the barrier, lane 0 carries the byte count
    // (a box must start on a 16-byte boundary -- a tensor copy from an unaligned column raises an illegal-instruction fault,
    // tools/probe/tma3d_probe.cu -- so it is the aligned superset of the runs, exactly what the bulk copies fetch)
    auto tensor_tile = [&](uint32_t tile) { return (P.flags & 2u) && ((52ull + 338ull * tile) & ~15ull) + PITCH <= P.band_stride; };
    auto fetch_tensor = [&](uint32_t f, uint32_t tile) {
        fence_async_smem();
        mbar_expect_tx(bar, lane == 0 ? 9u * PITCH : 0u);
        if (lane == 0)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(R)), "l"(&tmap), "r"((int)(((52u + 338u * tile) & ~15u) >> 1)), "r"(0), "r"((int)f), "r"(bar) : "memory");   // x in 16-bit elements
    };
    TileCursor c;
    cursor_seek<PIX>(c, P, g, P.out_stride, P.in_stride, mt_lo, lane);
    if (lane < 9) { if (tensor_tile(c.tile)) fetch_tensor(c.f, c.tile); else fetch(c.run); }
    stagger_start(P.flags, warp, mt_hi - mt_lo);
    uint32_t phase = 0;
    bool first = true;
    for (uint32_t left = mt_hi - mt_lo; left; --left) {
        const bool last = left == 1 || c.tile + 1 == P.tile0 + P.n_tiles;           // of a contiguous stretch
        const uint32_t pad = (uint32_t)c.pix & 15u, padb = (uint32_t)c.run & 15u;
        mbar_wait(bar, phase);
        phase ^= 1;
        if (last && !tensor_tile(c.tile)) { // only the clipped end of the buffer (the last chunks of the last frame, which the bulk copy left out) is fetched here
            if (lane < 9) {
                const uint64_t a0 = c.run - padb, a1 = a0 + ((padb + L::RUN + 15u) & ~15u);
                if (a1 > in_limit)
                    for (uint64_t ga = (in_limit > a0 ? (in_limit - a0) & ~15ull : 0) + a0; ga < a1; ++ga) R[PITCH * lane + (ga - a0)] = ga < in_limit ? P.in[ga] : 0;
            }
            __syncwarp();
        }
        if (c.tile == 0) {
            if (lane == 0 && g.cw_base[0] == 0) { // body symbols 0,1: move them from the transient states to the periodic ones
                uint8_t* r0 = R + padb;
                r0[0] = sg.scr[st_of(g, 0, 0)][sg.dsc[g.st[0]][r0[0] % 27u]];
                r0[1] = sg.scr[st_of(g, 0, 1)][sg.dsc[g.st[1]][r0[1] % 27u]];
            }
            __syncwarp();
        }
        // ---- phase B: syndrome screen per codeword (passes 0..2 variant-uniform); descrambled data symbols -> stream order
        {
            uint32_t* status = P.status + 2 * c.f;
            const uint2* rt = rec + 128 * c.tm;
            static_for<0, 3>([&](auto pc) {
                constexpr int p = decltype(pc)::value;
                const uint2 r = rt[32 * p + lane];
                const uint32_t pb = __shfl_sync(0xFFFFFFFFu, padb, (int)(r.y & 0xFFu));
                dec_cw5<K>(R32 + (r.x >> 16) + pb, S32 + (r.x & 0xFFFFu), tabA32 + p * L5::DEC_VAR, chk32 + 8 * p, sg32, status);
            });
            const uint2 r = rt[96 + lane];
            const uint32_t pb = __shfl_sync(0xFFFFFFFFu, padb, (int)(r.y & 0xFu));
            if (r.y != REC_IDLE) {
                const uint32_t v = r.y >> 8;
                dec_cw5<K>(R32 + (r.x >> 16) + pb, S32 + (r.x & 0xFFFFu), tabA32 + v * L5::DEC_VAR, chk32 + 8 * v, sg32, status);
            }
        }
        __syncwarp();
        if (left > 1 && lane < 9) {   // R is free again: next tile's runs on their way (the cursor itself advances at the end of the loop)
            const bool wrap = c.tile + 1 == P.tile0 + P.n_tiles;
            const uint32_t nf = wrap ? c.f + 1 : c.f, nt = wrap ? P.tile0 : c.tile + 1;
            if (tensor_tile(nt)) fetch_tensor(nf, nt);
            else fetch(wrap ? P.in_stride * nf + 52 + 26 * (g.cw_base[lane] + (uint64_t)C_MINI * nt) : c.run + 26 * C_MINI);
        }
        if (lane == 0) {
            bulk_wait_read();                                                        // the previous tile's bulk store has read OUT
            if (!first) *reinterpret_cast<uint4*>(OUT) = carry[0];                   // bytes [0, pad): the previous tile's tail
        }
        __syncwarp();
        if constexpr (WORDS) dec_phase_a_words<K>(S, OUT, pad, lane); else dec_phase_a5<K>(S, OUT, pad, lane, smem + L5::DEC_CLUT);
        fence_async_smem();
        __syncwarp();
        {   // the pixel side of the tile -> global: whole chunks by one bulk store, edge bytes of a stretch one by one
            const uint32_t end = pad + PIX, cend = end >> 4, c0 = (first && pad) ? 1u : 0u;
            uint8_t* g0 = P.out + (c.pix - pad);
            if (lane == 0) {
                if (!last) carry[0] = *reinterpret_cast<const uint4*>(OUT + 16 * cend);
                bulk_s2g(g0 + 16 * c0, smem_u32(OUT + 16 * c0), 16 * (cend - c0));
                bulk_commit();
            }
            if (first && pad && lane >= (int)pad && lane < 16) g0[lane] = OUT[lane];
            if (last && lane < 16 && 16 * cend + lane < end) g0[16 * cend + lane] = OUT[16 * cend + lane];
        }
        __syncwarp();
        first = last;
        cursor_next<PIX>(c, P, g, P.out_stride, P.in_stride, lane);
    }
    if (lane == 0) bulk_wait_all();
}

// ---- the bridge of the general chain at streaming speed (rgb_to_quant_stream / quant_stream_to_rgb, IMG:156-192): eight pixels per
// thread -- 24 bytes of RGB8 as three 8-byte accesses, 48 bytes of PixelYCbCrQuant {u16 Yq, i16 Cbq, i16 Crq} as three 16-byte ones --
// with the integer arithmetic of the fused kernels (luma ties and out-of-range quantised values take the float32 path of dev.cuh).
// 9 bytes per pixel either way; the one-pixel-per-thread kernels of k_general.cu (byte loads, 2-byte stores) stay for unaligned
// buffers and the last n mod 8 pixels.
__device__ __forceinline__ void px8_from_words(const uint32_t (&y)[6], uint32_t (&px)[8])
{
    px[0] = y[0]; px[1] = __byte_perm(y[0], y[1], 0x6543); px[2] = __byte_perm(y[1], y[2], 0x5432); px[3] = __byte_perm(y[2], y[2], 0x0321);
    px[4] = y[3]; px[5] = __byte_perm(y[3], y[4], 0x6543); px[6] = __byte_perm(y[4], y[5], 0x5432); px[7] = __byte_perm(y[5], y[5], 0x0321);
}
__global__ void __launch_bounds__(256) k_rgb_to_quant8(const uint2* __restrict__ rgb, size_t n_groups, uint4* __restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n_groups) return;
    const uint2 a = __ldcs(rgb + 3 * i), b = __ldcs(rgb + 3 * i + 1), c = __ldcs(rgb + 3 * i + 2);
    const uint32_t y[6] = {a.x, a.y, b.x, b.y, c.x, c.y};
    uint32_t px[8], h[24];
    px8_from_words(y, px);
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        const uint32_t t = __umulhi(dot_rgb<2>(px[p], 500), Y5_M);
        uint32_t yq = __umulhi(lop3<0xEA>(t, 0xFFFFFE00u, Y5_Z), Y5_M2);
        if (!(t & 511u)) yq = luma_q(byte_magic(px[p], 0), byte_magic(px[p], 1), byte_magic(px[p], 2)) + (0u - QY_C0);   // possible tie: float32 decides
        h[3 * p] = yq;
        h[3 * p + 1] = (chroma_q<true>(px[p]) - 40u) & 0xFFFFu;
        h[3 * p + 2] = (chroma_q<false>(px[p]) - 40u) & 0xFFFFu;
    }
    uint32_t w[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) w[j] = h[2 * j] | (h[2 * j + 1] << 16);
    __stcs(out + 3 * i, make_uint4(w[0], w[1], w[2], w[3]));
    __stcs(out + 3 * i + 1, make_uint4(w[4], w[5], w[6], w[7]));
    __stcs(out + 3 * i + 2, make_uint4(w[8], w[9], w[10], w[11]));
}
__global__ void __launch_bounds__(256) k_quant_to_rgb8(const uint4* __restrict__ px, size_t n_groups, uint2* __restrict__ rgb)
{
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n_groups) return;
    const uint4 a = __ldcs(px + 3 * i), b = __ldcs(px + 3 * i + 1), c = __ldcs(px + 3 * i + 2);
    const uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
    uint32_t p[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const uint32_t h0 = (w[(3 * q) >> 1] >> (16 * ((3 * q) & 1))) & 0xFFFFu, h1 = (w[(3 * q + 1) >> 1] >> (16 * ((3 * q + 1) & 1))) & 0xFFFFu,
                       h2 = (w[(3 * q + 2) >> 1] >> (16 * ((3 * q + 2) & 1))) & 0xFFFFu;
        const uint32_t ub = (h1 + 40u) & 0xFFFFu, ur = (h2 + 40u) & 0xFFFFu;
        if (h0 <= 242u && ub <= 80u && ur <= 80u) p[q] = yuvq_to_rgb5(h0, ub, ur);
        else {                                                       // anything a PixelYCbCrQuant can hold: the clamping float32 path (IMG:57-66,79-84)
            int R, G, B;
            ycbcr8_to_rgb(dequant_y((int)h0), dequant_c((int)(int16_t)h1), dequant_c((int)(int16_t)h2), R, G, B);
            p[q] = (uint32_t)R | ((uint32_t)G << 8) | ((uint32_t)B << 16);
        }
    }
    __stcs(rgb + 3 * i, make_uint2(p[0] | (p[1] << 24), (p[1] >> 8) | (p[2] << 16)));
    __stcs(rgb + 3 * i + 1, make_uint2((p[2] >> 16) | (p[3] << 8), p[4] | (p[5] << 24)));
    __stcs(rgb + 3 * i + 2, make_uint2((p[5] >> 8) | (p[6] << 16), (p[6] >> 16) | (p[7] << 8)));
}
