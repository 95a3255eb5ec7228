// dev.cuh -- device-side building blocks shared by the kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "t3c_internal.h"

namespace t3c {

// ---------------------------------------------------------------------------------------------
// GF(3)^24 arithmetic on bit planes.  A vector of up to 24 trits lives in two 32-bit registers:
//   nz  : bit set where the trit is 1 or 2        two : bit set where the trit is 2
// With this encoding (0->00, 1->10, 2->11) trit-wise addition mod 3 is exactly three LOP3s
// (found by exhaustive search over 2-level LOP3 networks; checked in tests/test_host_logic.py).
// ---------------------------------------------------------------------------------------------
template <int IMM>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(IMM));
    return r;
}
struct Planes { uint32_t nz, two; };
__device__ __forceinline__ void gf3_add(Planes& a, uint32_t bnz, uint32_t btwo)
{
    const uint32_t t = lop3<0x92>(a.nz, a.two, btwo);
    const uint32_t s0 = lop3<0xE6>(t, a.nz, bnz);
    const uint32_t s1 = lop3<0x24>(t, a.two, bnz);
    a.nz = s0;
    a.two = s1;
}
__device__ __forceinline__ void gf3_add(Planes& a, uint64_t e) { gf3_add(a, (uint32_t)e, (uint32_t)(e >> 32)); }
// four symbols (bytes) from plane bits b0,b1,b2 of each byte: b0 + 3 b1 + 9 b2, trit = nz + two
__device__ __forceinline__ uint32_t trits_to_sym4(uint32_t y)
{
    return y + ((y >> 1) & 0x01010101u) + 5u * ((y >> 2) & 0x01010101u);
}
__device__ __forceinline__ uint32_t planes_to_sym4_lo(const Planes& p) // parity symbols 0..3
{
    return trits_to_sym4(p.nz & 0x07070707u) + trits_to_sym4(p.two & 0x07070707u);
}
__device__ __forceinline__ uint32_t planes_to_sym4_hi(const Planes& p) // parity symbols 4..7
{
    return trits_to_sym4((p.nz >> 4) & 0x07070707u) + trits_to_sym4((p.two >> 4) & 0x07070707u);
}

// ---------------------------------------------------------------------------------------------
// Index maps of the wire format
// ---------------------------------------------------------------------------------------------
// scrambler state for pre-beacon body index p (A.4)
// (the index maps are templates on the index type: super-frames below 2^31 symbols -- an 8K frame has 1.9e8 -- use 32-bit
// arithmetic, where a division is ~5x cheaper than in 64 bits)
template <typename I>
__device__ __forceinline__ uint32_t scr_state(const Geom& g, I p)
{
    return p < 2 ? g.st[p] : g.st[2 + (uint32_t)((p - 2) % 6)];
}
// pre-beacon body index p -> index in the beacon-expanded body (A.5)
template <typename I>
__device__ __forceinline__ I beacon_expand(const Geom& g, I p)
{
    if (g.period == 0 || g.slot < 0) return p;
    const I blk = p / (I)g.beacon_per;
    const uint32_t rem = (uint32_t)(p - blk * g.beacon_per);
    if (rem < 8) return 9 * (blk * g.period) + (rem < (uint32_t)g.slot ? rem : rem + 1);
    return 9 * (blk * g.period + 1 + (rem - 8) / 9) + (rem - 8) % 9;
}
// 2D boustrophedon: position i of the permuted stream reads position perm2d(i) of the source;
// the map is an involution, so the same function de-interleaves (A.2, OLD:750-813).
template <typename I>
__device__ __forceinline__ I perm2d(I i, I n, I area, uint32_t w)
{
    if (area == 0) return i;
    const I base = (i / area) * area;
    const I take = (n - base) < area ? (n - base) : area;
    const I off = i - base, r = off / w;
    if ((r & 1) == 0) return i;
    const I rs = r * w;
    const I cnt = (take - rs) < w ? (take - rs) : (I)w;
    return base + rs + (cnt - 1 - (off - rs));
}
// symbol j of the regrouped stream (A.1): trits 3j..3j+2 of the 26-trits-per-word stream of `raw`
template <typename I>
__device__ __forceinline__ uint32_t raw_trit(const uint8_t* __restrict__ raw, I n_words, I ti)
{
    const I w = ti / 26;
    if (w >= n_words) return 0;
    const uint32_t o = (uint32_t)(ti - w * 26);
    const uint32_t s = raw[9 * w + o / 3];
    const uint32_t c = o % 3;
    return c == 0 ? s % 3 : (c == 1 ? (s / 3) % 3 : (s / 9) % 3); // unpack3, OLD:28-31
}
template <typename I>
__device__ __forceinline__ uint32_t raw_symbol(const uint8_t* __restrict__ raw, I n_words, I j)
{
    return raw_trit(raw, n_words, 3 * j) + 3 * raw_trit(raw, n_words, 3 * j + 1) + 9 * raw_trit(raw, n_words, 3 * j + 2);
}

// ---------------------------------------------------------------------------------------------
// GF(27) through shared-memory look-ups (general / slow paths)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_gf(GfTables& dst, const GfTables* __restrict__ src)
{
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d = reinterpret_cast<uint32_t*>(&dst);
    for (uint32_t i = threadIdx.x; i < sizeof(GfTables) / 4; i += blockDim.x) d[i] = s[i];
}
__device__ __forceinline__ uint8_t gmul(const GfTables& g, uint32_t a, uint32_t b) { return g.mul[a * 27 + b]; }
__device__ __forceinline__ uint8_t gadd(const GfTables& g, uint32_t a, uint32_t b) { return g.add[a * 27 + b]; }
__device__ __forceinline__ uint8_t gsub(const GfTables& g, uint32_t a, uint32_t b) { return g.add[a * 27 + g.neg[b]]; }

// One thread decodes one RS(26,k) block: syndromes, Berlekamp-Massey, Chien, Forney, exactly as
// RSCodec::decode_block (OLD:546-662) with the vectors restated as zero-padded fixed arrays (their
// logical sizes never exceed r+1 <= 9).  fixed selects the Forney sign (bug B2 / Appendix B).
// c[] is corrected in place in ascending position order, including the partial corrections the
// reference leaves behind when it bails out on a zero derivative.
// Two things are done differently from the reference without changing any result:
//  * polynomial evaluations start at the highest non-zero coefficient (leading zeros contribute nothing to a Horner sum);
//  * with the repaired arithmetic, syndromes of the form S_j = e X^(j+1) (X != 0) are the one-error case: Berlekamp-Massey
//    would find sigma = 1 - X x (the shortest recurrence is unique while 2L <= r), Chien its single root at position log X and
//    Forney the magnitude e, so the correction c[log X] -= e is applied directly.
// strict (the frame-level consistent decoder only; the block-level API stays bit-exact with the reference's decode_block): reject
// unless Berlekamp-Massey's L <= t, deg(sigma) == L and sigma has L distinct roots -- with more than t errors the reference accepts
// locators that are not error locators and "corrects" 0..t symbols of a block it cannot decode (OLD:611-624 only counts roots).
static __device__ __noinline__ bool rs_decode_core(const GfTables& g, uint8_t* c, int k, bool fixed, const uint8_t* S, bool strict = false)
{
    const int r = 26 - k, t = r >> 1;
    if (fixed && S[0]) { // one error?
        const uint32_t X = gmul(g, S[1], g.inv[S[0]]);
        bool one = X != 0;
        for (int j = 1; j + 1 < r; ++j) one = one && S[j + 1] == gmul(g, S[j], X);
        if (one) {
            const uint32_t pos = g.lg[X], e = gmul(g, S[0], g.inv[X]);
            c[pos] = gsub(g, c[pos], e);
            return true;
        }
    }
    uint8_t sg[10], B[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) sg[i] = B[i] = 0;
    sg[0] = B[0] = 1;
    int L = 0, m = 1;
    for (int n = 0; n < r; ++n) {
        uint32_t delta = S[n];
        for (int i = 1; i <= L; ++i) delta = gadd(g, delta, gmul(g, sg[i], S[n - i]));
        if (delta) {
            uint8_t T[10];
#pragma unroll
            for (int i = 0; i < 10; ++i) T[i] = sg[i];
            for (int i = m; i < 10; ++i) sg[i] = gsub(g, sg[i], gmul(g, delta, B[i - m]));
            if (2 * L <= n) {
                const uint32_t invd = g.inv[delta];
#pragma unroll
                for (int i = 0; i < 10; ++i) B[i] = gmul(g, T[i], invd);
                L = n + 1 - L;
                m = 1;
            } else ++m;
        } else ++m;
    }
    uint8_t Om[8];
    for (int i = 0; i < r; ++i) {
        uint32_t acc = 0;
        for (int a = 0; a <= i; ++a) acc = gadd(g, acc, gmul(g, S[a], sg[i - a])); // i-a <= 7 < 10
        Om[i] = (uint8_t)acc;
    }
    int top = 9;
    while (top > 0 && sg[top] == 0) --top; // sg[0] = 1 always
    if (strict && (L > t || top != L)) return false;
    // Chien search: positions i with sigma(alpha^-i) = 0.  Degrees 1 and 2 are solved instead of searched (the same root set):
    // in characteristic 3, x^2 + b x + c = (x - b)^2 - (b^2 - c), so the roots are b +- sqrt(b^2 - c)
    uint32_t roots = 0;
    int nroots = 0;
    if (top == 1) {
        const uint32_t x = g.neg[g.inv[sg[1]]];                 // 1 + s1 x = 0
        roots = 1u << ((26 - g.lg[x]) % 26);
        nroots = 1;
    } else if (top == 2) {
        const uint32_t i2 = g.inv[sg[2]], b = gmul(g, sg[1], i2), q = g.sqr[gsub(g, gmul(g, b, b), i2)]; // x^2 + b x + 1/s2
        if (q != 255) {
            const uint32_t x1 = gadd(g, b, q), x2 = gsub(g, b, q);  // never 0: their product is 1/s2
            roots = 1u << ((26 - g.lg[x1]) % 26);
            roots |= 1u << ((26 - g.lg[x2]) % 26);
            nroots = x1 == x2 ? 1 : 2;
        }
    } else {
        for (int i = 0; i < 26; ++i) {
            const uint32_t x = g.exp[(26 - i) % 26];
            uint32_t acc = sg[top];
            for (int d = top - 1; d >= 0; --d) acc = gadd(g, gmul(g, acc, x), sg[d]);
            if (acc == 0) { roots |= 1u << i; ++nroots; }
        }
    }
    if (nroots > t || (strict && nroots != L)) return false;
    uint8_t sp[9];
    for (int i = 1; i < 10; ++i) {
        const int im = i % 3;
        sp[i - 1] = im == 0 ? 0 : (im == 1 ? sg[i] : g.neg[sg[i]]); // 2a = -a in characteristic 3
    }
    int tsp = top > 0 ? top - 1 : 0, tom = r - 1;
    while (tsp > 0 && sp[tsp] == 0) --tsp;
    while (tom > 0 && Om[tom] == 0) --tom;
    for (uint32_t left = roots; left; left &= left - 1) { // ascending positions
        const int pos = __ffs((int)left) - 1;
        const uint32_t x = g.exp[(26 - pos) % 26];
        uint32_t num = Om[tom], den = sp[tsp];
        for (int d = tom - 1; d >= 0; --d) num = gadd(g, gmul(g, num, x), Om[d]);
        for (int d = tsp - 1; d >= 0; --d) den = gadd(g, gmul(g, den, x), sp[d]);
        if (den == 0) return false;
        const uint32_t mag = gmul(g, g.neg[num], g.inv[den]);
        c[pos] = fixed ? gsub(g, c[pos], mag) : gadd(g, c[pos], mag);
    }
    return true;
}
// power-sum syndromes straight from the block (any arithmetic)
static __device__ __noinline__ bool rs_decode_thread(const GfTables& g, uint8_t* c, int k, bool fixed, bool strict = false)
{
    const int r = 26 - k;
    uint8_t S[8];
    bool all0 = true;
    for (int j = 0; j < r; ++j) {
        uint32_t acc = 0, e = 0;
        for (int i = 0; i < 26; ++i) {
            acc = gadd(g, acc, gmul(g, c[i], g.exp[e]));
            e += j + 1;
            if (e >= 26) e -= 26;
        }
        S[j] = (uint8_t)acc;
        all0 = all0 && acc == 0;
    }
    if (all0) return true;
    return rs_decode_core(g, c, k, fixed, S, strict);
}
// The Chien tables lie directly behind the GF(27) tables (HostTables; the tiled kernels copy both into their shared image)
__device__ __forceinline__ const uint32_t* chien_of(const GfTables* gf) { return reinterpret_cast<const uint32_t*>(gf + 1); }

// Slow path of the tiled decoders (repaired code, strict acceptance): a bounded-distance decoder in registers with uniform control
// flow, so that a warp whose lanes hold codewords with different error counts does not serialise.  Under strict acceptance
// (rs_decode_core above: L <= t, deg sigma == L, L distinct roots) a block is corrected iff a codeword lies within distance t, and
// then to that codeword -- which any bounded-distance decoder finds -- so this one returns what rs_decode_core(strict) returns on the same block:
//  * syndromes from the parity residual p = parity(received data) - received parity, which the syndrome screen has already
//    computed (res_lo/res_hi: symbols 0..3 / 4..7 as bytes): S_j = sum_m syn[j][m] p_m, r*r products instead of 26*r;
//  * Berlekamp-Massey (Massey's form, OLD:572-600) on T+1 coefficients with x^m B kept pre-shifted: while L <= T neither sigma nor
//    x^m B has a term above x^T (deg x^m B <= n+1-L), and L > T is final, so dropping higher terms changes no accepted result;
//  * roots of sigma for all 26 positions at once on GF(3) bit planes (ChienTables); fewer than L roots covers deg sigma < L;
//  * Forney with Omega = S sigma mod x^T (the key equation makes the higher coefficients vanish for an accepted locator) and the
//    formal derivative in characteristic 3; the L corrections go straight to the data symbols already stored at dst (stride 9).
// status[0] = 0 when a block is rejected, status[1] += corrected symbols (count), aggregated over the lanes that are here together.
template <int K>
static __device__ __noinline__ void rs_bd_fix(const GfTables& g, const uint32_t* __restrict__ chien, uint8_t* dst, uint32_t res_lo, uint32_t res_hi,
                                              uint32_t* status, bool count)
{
    constexpr int R = 26 - K, T = R / 2, KI = (24 - K) / 2;
    uint32_t S[R];
    {
        uint32_t p[R];
#pragma unroll
        for (int m = 0; m < R; ++m) p[m] = ((m < 4 ? res_lo : res_hi) >> (8 * (m & 3))) & 0xFFu;
#pragma unroll
        for (int j = 0; j < R; ++j) {
            uint32_t acc = gmul(g, p[0], g.syn[KI][j][0]);
#pragma unroll
            for (int m = 1; m < R; ++m) acc = gadd(g, acc, gmul(g, p[m], g.syn[KI][j][m]));
            S[j] = acc;
        }
    }
    uint32_t sig[T + 1], Bs[T + 1];                       // sigma and x^m B
#pragma unroll
    for (int i = 0; i <= T; ++i) sig[i] = Bs[i] = 0;
    sig[0] = 1;
    Bs[1] = 1;
    uint32_t L = 0;
    bool bad = false;
#pragma unroll
    for (int n = 0; n < R; ++n) {
        uint32_t delta = S[n];
#pragma unroll
        for (int i = 1; i <= T; ++i)
            if (i <= n) delta = gadd(g, delta, gmul(g, sig[i], S[n - i]));
        const bool grow = delta != 0 && 2 * L <= (uint32_t)n;
        const uint32_t nd = g.neg[delta], invd = g.inv[delta];
        uint32_t old[T + 1];
#pragma unroll
        for (int i = 0; i <= T; ++i) old[i] = sig[i];
#pragma unroll
        for (int i = 1; i <= T; ++i) sig[i] = gadd(g, sig[i], gmul(g, nd, Bs[i]));   // delta = 0 adds nothing; x^m B has no constant term
#pragma unroll
        for (int i = T; i >= 1; --i) Bs[i] = grow ? gmul(g, old[i - 1], invd) : Bs[i - 1];
        if (grow) L = (uint32_t)n + 1 - L;
        bad = bad || L > (uint32_t)T;
    }
    uint32_t z[3];
    {
        Planes a[3] = {{0x09249249u, 0}, {0x09249249u, 0}, {0x00009249u, 0}};                  // sigma_0 = 1 at the 10 + 10 + 6 positions
#pragma unroll
        for (int j = 1; j <= T; ++j) {
            const uint2* e = reinterpret_cast<const uint2*>(chien + ((j - 1) * 27 + sig[j]) * 6);
            const uint2 e0 = e[0], e1 = e[1], e2 = e[2];
            gf3_add(a[0], e0.x, e1.y);
            gf3_add(a[1], e0.y, e2.x);
            gf3_add(a[2], e1.x, e2.y);
        }
#pragma unroll
        for (int w = 0; w < 3; ++w) z[w] = ~(a[w].nz | (a[w].nz >> 1) | (a[w].nz >> 2)) & (w < 2 ? 0x09249249u : 0x00009249u);
    }
    const bool ok = !bad && (uint32_t)(__popc(z[0]) + __popc(z[1]) + __popc(z[2])) == L;
    uint32_t Om[T], sp[T];
#pragma unroll
    for (int i = 0; i < T; ++i) {
        uint32_t acc = S[i];
#pragma unroll
        for (int a = 0; a < i; ++a) acc = gadd(g, acc, gmul(g, S[a], sig[i - a]));
        Om[i] = acc;
        sp[i] = (i + 1) % 3 == 0 ? 0u : ((i + 1) % 3 == 1 ? sig[i + 1] : (uint32_t)g.neg[sig[i + 1]]);   // 2a = -a
    }
#pragma unroll
    for (int e = 0; e < T; ++e) {
        if (ok && (uint32_t)e < L) {
            const uint32_t w = z[0] ? 0u : (z[1] ? 1u : 2u), zw = z[0] ? z[0] : (z[1] ? z[1] : z[2]);
            const uint32_t b = (uint32_t)__ffs((int)zw) - 1u, pos = 10u * w + (b * 11u >> 5);    // b / 3 for b < 32
            const uint32_t low = zw & (0u - zw);
            if (w == 0) z[0] ^= low; else if (w == 1) z[1] ^= low; else z[2] ^= low;
            const uint32_t x = g.exp[pos ? 26u - pos : 0u];
            uint32_t num = Om[T - 1], den = sp[T - 1];
#pragma unroll
            for (int d = T - 2; d >= 0; --d) {
                num = gadd(g, gmul(g, num, x), Om[d]);
                den = gadd(g, gmul(g, den, x), sp[d]);
            }
            const uint32_t mag = gmul(g, g.neg[num], g.inv[den]);
            if (pos < (uint32_t)K) dst[9 * pos] = gsub(g, dst[9 * pos], mag);
        }
    }
    // one status update per frame and warp: with every codeword of an 8K frame dirty, 6.8 M same-address atomics are 4 ms on their own
    const uint32_t am = __activemask();
    const uint32_t grp = __match_any_sync(am, reinterpret_cast<uintptr_t>(status));
    const uint32_t fixed = __reduce_add_sync(grp, ok && count ? L : 0u);
    const uint32_t failed = __ballot_sync(am, !ok) & grp;
    if (((uint32_t)threadIdx.x & 31u) == (uint32_t)__ffs((int)grp) - 1u) {
        if (fixed) atomicAdd(&status[1], fixed);
        if (failed) atomicExch(&status[0], 0u);
    }
}

// ---------------------------------------------------------------------------------------------
// Per-pixel arithmetic
// ---------------------------------------------------------------------------------------------
// round-half-away for x in [0, 2^21): floor(x + 0.5).  FADD.RM against 2^22+0.5 leaves
// 2*floor(2x+1)/2 in the mantissa, one rounding only, so there is no double-rounding hazard.
__device__ __forceinline__ int round_pos(float x)
{
    return (__float_as_int(__fadd_rd(x, 4194304.5f)) >> 1) & 0xFFFFF;
}
// rgb_to_ycbcr (IMG:47-56): float32, left-to-right, no FMA contraction, round half away, clamp
__device__ __forceinline__ void rgb_to_ycbcr8(uint32_t R, uint32_t G, uint32_t B, int& Y, int& Cb, int& Cr)
{
    const float r = (float)R, g = (float)G, b = (float)B;
    const float y = __fadd_rn(__fadd_rn(__fmul_rn(0.299f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b));
    const float cb = __fadd_rn(__fadd_rn(__fsub_rn(__fmul_rn(-0.168736f, r), __fmul_rn(0.331264f, g)), __fmul_rn(0.5f, b)), 128.0f);
    const float cr = __fadd_rn(__fsub_rn(__fsub_rn(__fmul_rn(0.5f, r), __fmul_rn(0.418688f, g)), __fmul_rn(0.081312f, b)), 128.0f);
    // all three are >= 0 for 8-bit inputs (min 128-127.5), so round-half-away == floor(x+0.5)
    Y = min(round_pos(y), 255);
    Cb = min(round_pos(fmaxf(cb, 0.0f)), 255);
    Cr = min(round_pos(fmaxf(cr, 0.0f)), 255);
}
// quantize_ycbcr (IMG:69-78) in exact integer form:
//   Yq = round(Y*242/255): no ties exist (484Y = 510m+255 has no solution), so (484Y+255)/510
//   Cq = round-half-away((C-128)*5/16) = ((5C + 7 + (C>=128)) >> 4) - 40
__device__ __forceinline__ int quant_y(int Y) { return (Y * 484 + 255) / 510; }
__device__ __forceinline__ int quant_c_off(int C) { return (5 * C + 7 + (C >> 7)) >> 4; } // Cq + 40, 0..80
// dequantize_ycbcr (IMG:79-84) in exact integer form.
//   Y = clamp(round(Yq*(255.0/242.0))): the only exact tie below the clamp is Yq=121 (127.5), where
//   the reference's double product is 127.49999999999999 and rounds DOWN; "+241" reproduces that and
//   is otherwise identical to floor(x+0.5).  C = clamp(round(128 + Cq*3.2)) has no ties.
__device__ __forceinline__ int dequant_y(int Yq) { return min((Yq * 510 + 241) / 484, 255); }
__device__ __forceinline__ int dequant_c(int Cq)
{
    const int v = 2570 + 64 * Cq; // (128 + 3.2 Cq + 0.5) * 20
    return v <= 0 ? 0 : min(v / 20, 255);
}
// ycbcr_to_rgb (IMG:57-66)
__device__ __forceinline__ int round_clamp255(float x)
{
    // std::round (half away from zero) then clamp to [0,255]; negatives all clamp to 0
    return x <= 0.0f ? 0 : min(round_pos(fminf(x, 300.0f)), 255);
}
__device__ __forceinline__ void ycbcr8_to_rgb(int Y, int Cb, int Cr, int& R, int& G, int& B)
{
    const float y = (float)Y, cb = __fsub_rn((float)Cb, 128.0f), cr = __fsub_rn((float)Cr, 128.0f);
    const float r = __fadd_rn(y, __fmul_rn(1.402f, cr));
    const float g = __fsub_rn(__fsub_rn(y, __fmul_rn(0.344136f, cb)), __fmul_rn(0.714136f, cr));
    const float b = __fadd_rn(y, __fmul_rn(1.772f, cb));
    R = round_clamp255(r);
    G = round_clamp255(g);
    B = round_clamp255(b);
}
// 13 trits of one pixel as an integer < 3^13: A = Yq%243 + 243*((Cbq+40)%81) + 19683*((Crq+40)%81),
// the digits i2tr keeps (OLD:675-682,697-702; out-of-range values wrap through the uint32 cast)
__device__ __forceinline__ uint32_t pixel_value(uint32_t yq, int cbq, int crq)
{
    return yq % 243u + 243u * ((uint32_t)(cbq + 40) % 81u) + 19683u * ((uint32_t)(crq + 40) % 81u);
}

} // namespace t3c
