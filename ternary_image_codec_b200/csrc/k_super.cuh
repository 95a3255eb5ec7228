// k_super.cuh -- tiled, fused kernels for the configs the warp-tile kernels of k_fast.cu leave out: per-band k (UEP,
// OLD:59-72), the 2D boustrophedon interleave (OLD:750-813) and the sparse beacon (OLD:95-113,1118-1141) -- BASELINE
// config 2.  Included by k_fast.cu inside its anonymous namespace (it reuses the unit / codeword bodies).
//
// Work unit = a SUPER-TILE of M symbols per band, M a common multiple of 26 and of every k in use: it starts on a
// 6-pixel unit (26 stream symbols), on a row of the 2D interleave (tile widths that divide 26) and on a codeword of every
// band, so super-tiles are independent and need no halo (luma-priority UEP: M = 2860, 990 units = 5940 pixels,
// 3 x 143 + 6 x 130 codewords).  One CTA works on one super-tile; two CTAs per SM cover each other's barriers.
//   encode: pixels -> (phase A, thread per unit) stream symbols in S -> [row reversal in place] -> (phase B, lane per
//           codeword, one (k, scrambler variant) per warp pass so a pass reads one conflict-free table block) nine
//           pre-beacon runs in U -> (phase C) beacon-expanded runs to global in 128-bit chunks
//   decode: the same phases backwards; the beacon slots are squeezed out while the runs are scaled by 4 for phase B.
// Codewords / pixels after the last full super-tile of a frame are left to the general kernels.
#pragma once

constexpr int SUP_TPB = 512, SUP_WARPS = SUP_TPB / 32, SUP_MAX_PASS = 64;
constexpr uint32_t SUP_IDLE = 0xFFFFu;
// T3C_SUPER_DEBUG builds: CTA 0 accumulates the clock cycles between its barriers into g_sup_dbg[phase] (read with t3c_debug_counters)
#ifdef T3C_SUPER_DEBUG
__device__ uint32_t g_sup_dbg[32];
#define SUP_TICK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_sup_dbg[(i)], (uint32_t)(t_ - sup_t0)); sup_t0 = t_; } } while (0)
#define SUP_TICK0() long long sup_t0 = clock64()
#else
#define SUP_TICK(i) do { } while (0)
#define SUP_TICK0() do { } while (0)
#endif

struct SuperPlan {
    uint32_t M, UN, n_tiles, nk;
    uint32_t kk[4];         // the distinct k values in use
    uint32_t ncw[4];        // codewords per band and super-tile for k slot s: M / kk[s]
    uint8_t kslot[12];      // band -> k slot
    uint32_t npass[3];      // phase-B passes per class of the super-tile index mod 3
    uint32_t run_base[9];   // staging slot (16-byte aligned, + 32 bytes of slack) of band b's pre-beacon run in U / R
    uint32_t raw_base[9];   // decode: slot of band b's run as it lies in the frame (beacon slots included)
    uint32_t off_tab[4], off_aux, off_gf, off_meta, off_S, off_U, off_IN, smem_bytes;
    uint32_t G, G_magic, Gm1_magic; // beacon: 9 * period (0: none), floor(2^32 / G) + 1, floor(2^32 / (G - 1)) + 1
    uint32_t slot, bsym;
    uint32_t tile_w, tile_area, tile_h26, h26_magic; // tile_h26: tile height when the width is 26 (rows = units: reversed in registers), else 0; floor(2^32 / h) + 1
    uint32_t ch_shift;      // log2 of the per-band slot count of the flattened 16-byte chunk loops
    uint32_t ctas_per_sm;   // 2 when two CTAs' shared memory fits one SM, else 1
    const uint16_t* map;    // [3][SUP_MAX_PASS * 32]: b | cl << 4, SUP_IDLE = idle lane
    const uint8_t* pass_kv; // [3][SUP_MAX_PASS]: k slot | variant << 2
};
struct SuperMeta {          // per super-tile, written by threads 0..8
    uint64_t g_lo[9];       // global byte offset of the first byte of band b's run in the frame
    uint32_t len[9];        // bytes of the run in the frame (beacon slots inside it included)
    uint32_t o_first[9];    // index, from the run's first byte, of the first beacon slot at or after it
    uint32_t stage[9];      // offset of the run's pre-beacon byte 0 inside U / R
    uint32_t nch[9];        // 16-byte chunks of the aligned superset of the run in the frame
};

// 16 bytes at an arbitrary shared-memory byte address (the buffer has >= 4 bytes of slack after the last byte read)
__device__ __forceinline__ uint4 lds_gather16(uint32_t a)
{
    const uint32_t aw = a & ~3u, sh = (a & 3u) * 8u;
    uint32_t x0, x1, x2, x3, x4;
    asm volatile("ld.shared.u32 %0, [%5];\n\tld.shared.u32 %1, [%5+4];\n\tld.shared.u32 %2, [%5+8];\n\tld.shared.u32 %3, [%5+12];\n\tld.shared.u32 %4, [%5+16];"
                 : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3), "=r"(x4) : "r"(aw) : "memory");
    return make_uint4(__funnelshift_r(x0, x1, sh), __funnelshift_r(x1, x2, sh), __funnelshift_r(x2, x3, sh), __funnelshift_r(x3, x4, sh));
}
// the 16-byte mask whose bytes [0, e) are 0xFF (e in 0..16), as four words: two clamped 64-bit shifts
__device__ __forceinline__ void below16(uint32_t e, uint32_t (&m)[4])
{
    const uint32_t s0 = min(8u * e, 64u), s1 = 8u * e > 64u ? 8u * e - 64u : 0u;
    const uint64_t lo = s0 == 64u ? ~0ull : ((1ull << s0) - 1ull), hi = s1 == 64u ? ~0ull : ((1ull << s1) - 1ull);
    m[0] = (uint32_t)lo; m[1] = (uint32_t)(lo >> 32); m[2] = (uint32_t)hi; m[3] = (uint32_t)(hi >> 32);
}
// bytes [0, e) of a, bytes [e, 16) of b
__device__ __forceinline__ uint4 merge16(uint4 a, uint4 b, uint32_t e)
{
    uint32_t m[4];
    below16(e, m);
    return make_uint4((a.x & m[0]) | (b.x & ~m[0]), (a.y & m[1]) | (b.y & ~m[1]), (a.z & m[2]) | (b.z & ~m[2]), (a.w & m[3]) | (b.w & ~m[3]));
}
// bytes [0, e) of a, byte e = v (v4 = v in every byte), bytes (e, 16) of b; e in 0..15
__device__ __forceinline__ uint4 merge16_put(uint4 a, uint4 b, uint32_t e, uint32_t v4)
{
    uint32_t m[4];
    below16(e, m);
    const uint32_t bm = 0xFFu << (8u * (e & 3u)), w = e >> 2;              // the slot's byte inside word w
    const uint32_t x[4] = {a.x, a.y, a.z, a.w}, y[4] = {b.x, b.y, b.z, b.w};
    uint32_t r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t bj = w == (uint32_t)j ? bm : 0u;
        r[j] = (x[j] & m[j]) | (v4 & bj) | (y[j] & ~(m[j] | bj));
    }
    return make_uint4(r[0], r[1], r[2], r[3]);
}
// frame-local indices fit 32 bits (checked by the plan)
template <bool DECODE>
__device__ __forceinline__ void super_run_meta(SuperMeta& m, const SuperPlan& P, const Geom& g, uint64_t frame_off, uint32_t T, int b)
{
    const uint32_t n = P.ncw[P.kslot[b]], L = 26u * n;
    const uint32_t p_lo = 26u * ((uint32_t)g.cw_base[b] + n * T), p_hi = p_lo + L;
    uint32_t o_lo = p_lo, o_last = p_hi - 1u, o_first = 0x7FFFFFFFu;
    if (P.G) {
        o_lo = beacon_expand<uint32_t>(g, p_lo);
        o_last = beacon_expand<uint32_t>(g, p_hi - 1u);
        const uint32_t j0 = o_lo <= P.slot ? 0u : (o_lo - P.slot + P.G - 1u) / P.G;
        o_first = j0 * P.G + P.slot - o_lo;
    }
    const uint64_t lo = frame_off + 52 + o_lo;
    const uint32_t len = o_last + 1u - o_lo;
    m.g_lo[b] = lo;
    m.len[b] = len;
    m.o_first[b] = o_first;
    m.nch[b] = (((uint32_t)lo & 15u) + len + 15u) >> 4;
    m.stage[b] = P.run_base[b] + ((DECODE || P.G) ? 0u : (uint32_t)lo & 15u); // encode without a beacon: the staged run keeps the frame's 16-byte phase
}
// boustrophedon rows of the super-tile (A.2) for tile widths 2 and 13: reverse, in place, every row whose index inside its
// w x h tile is odd.  Super-tiles start on a row (9M is a multiple of 26, w divides 26) and hold whole rows only.  (Width 26:
// rows are units, reversed in registers by phase A; width 1: nothing to do.)
__device__ __forceinline__ void super_reverse_rows(uint8_t* S, const SuperPlan& P, uint32_t T, int tid)
{
    const uint32_t w = P.tile_w, n_rows = 9u * P.M / w;
    const uint64_t s0 = 9ull * P.M * T;
    for (uint32_t row = tid; row < n_rows; row += SUP_TPB) {
        const uint64_t pos = s0 + (uint64_t)row * w;
        const uint32_t r = (uint32_t)((pos % P.tile_area) / w);
        if (!(r & 1u)) continue;
        uint8_t* p = S + row * w;
        for (uint32_t i = 0; i < w / 2; ++i) { const uint8_t a = p[i], c = p[w - 1 - i]; p[i] = c; p[w - 1 - i] = a; }
    }
}
// 26-wide tiles: is unit x of the frame an odd row of its tile?  x mod h through the reciprocal (exact: x * h < 2^32, plan)
__device__ __forceinline__ bool super_row_odd(const SuperPlan& P, uint32_t x)
{
    if (!P.tile_h26) return false;
    return ((x - __umulhi(x, P.h26_magic) * P.tile_h26) & 1u) != 0;
}
__device__ __forceinline__ bool super_rows_in_smem(const SuperPlan& P) { return P.tile_area && !P.tile_h26 && P.tile_w > 1; }
// the lines of a byte range, towards L2 (the next super-tile's input while this one is being coded)
__device__ __forceinline__ void super_prefetch(const uint8_t* base, uint64_t lo, uint32_t bytes, uint64_t limit, int tid)
{
    const uint64_t a0 = lo & ~127ull;
    const uint32_t n = (uint32_t)((lo + bytes - a0 + 127) >> 7);
    for (uint32_t i = tid; i < n; i += SUP_TPB) if (a0 + 128ull * i < limit) prefetch_l2(base + a0 + 128ull * i);
}

// =============================================================================================
// encode
// =============================================================================================
template <bool WORDS>
__global__ void __launch_bounds__(SUP_TPB, 2) k_encode_super(FastParams Q, SuperPlan P, Geom g, const GfTables* __restrict__ gf, const RsTables* __restrict__ rs)
{
    constexpr uint32_t PIXB = WORDS ? 27u : 18u;   // pixel-side bytes of one unit
    extern __shared__ __align__(16) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* S = smem + P.off_S;                   // stream symbols, pre-scaled by 4
    uint8_t* U = smem + P.off_U;                   // the nine pre-beacon runs
    uint32_t* pat = reinterpret_cast<uint32_t*>(smem + P.off_aux); // [k slot][variant][2]
    for (uint32_t ks = 0; ks < P.nk; ++ks) {
        const int K = (int)P.kk[ks];
        const uint32_t(*pl)[kVals][2] = rs->pl[g.arith][(24 - K) / 2];
        for (int idx = tid; idx < 3 * K * 27; idx += SUP_TPB) {
            const int v = idx / (K * 27), rem = idx - v * (K * 27), i = rem / 27, d = rem - 27 * i;
            uint32_t* blk = reinterpret_cast<uint32_t*>(smem + P.off_tab[ks] + v * (8 * K * 27));
            blk[rem] = pl[i][d][0] | gf->scr[st_of(g, v, i)][d];
            blk[K * 27 + rem] = pl[i][d][1];
        }
        if (tid < 3) { // the scrambler as seen by the parity symbols of a variant-tid codeword, in the plane domain
            uint32_t nz = 0, two = 0;
            for (int j = 0; j < 26 - K; ++j) {
                const uint32_t st = st_of(g, tid, K + j);
                const int sh = K == 18 ? plane_shift<18>(j) : plane_shift<20>(j);
                if (st) nz |= 7u << sh;
                if (st == 2) two |= 7u << sh;
            }
            pat[6 * ks + 2 * tid] = nz;
            pat[6 * ks + 2 * tid + 1] = two;
        }
    }
    const uint32_t bar = smem_u32(smem + P.off_aux + 96);
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    __syncthreads();
    const uint32_t smem32 = smem_u32(smem), U32 = smem_u32(U), bsym4 = P.bsym * 0x01010101u;
    const uint64_t in_limit = Q.in_stride * Q.n_frames;
    const uint32_t total = P.n_tiles * Q.n_frames;
    uint8_t* IN = smem + P.off_IN;                 // the pixel-side bytes of a super-tile, fetched by the bulk copy engine one super-tile ahead
    SuperMeta* metas = reinterpret_cast<SuperMeta*>(smem + P.off_meta); // two copies: phase C of one super-tile runs beside phase A of the next
    // bulk load of super-tile st2: the 16-byte aligned superset of its pixel bytes, clipped to the buffer
    auto fetch = [&](uint32_t st2) {
        const uint32_t f2 = st2 / P.n_tiles, T2 = st2 - f2 * P.n_tiles;
        const uint64_t lo = Q.in_stride * f2 + (uint64_t)PIXB * P.UN * T2, a0 = lo & ~15ull;
        uint32_t bytes = ((uint32_t)(lo - a0) + PIXB * P.UN + 15u) & ~15u;
        if (a0 + bytes > in_limit) bytes = (uint32_t)(in_limit - a0) & ~15u;
        fence_async_smem();
        mbar_expect_tx(bar, bytes);
        bulk_g2s(smem_u32(IN), Q.in + a0, bytes, bar);
    };
    uint32_t phase = 0;
    // ---- phase A of super-tile st2: its pixels (in IN once the mbarrier flips) -> stream symbols in S; thread per unit, dealt even / odd
    // inside chunks of 64 units (conflict-free 36- and 52-byte lane strides); also its run geometry into metas[mb]
    auto phase_a_wait = [&](uint32_t st2, uint32_t mb) { // all threads: geometry, arrival of the pixels, the clipped end of the buffer
        const uint32_t f2 = st2 / P.n_tiles, T2 = st2 - f2 * P.n_tiles;
        const uint64_t g_lo = Q.in_stride * f2 + (uint64_t)PIXB * P.UN * T2, a0 = g_lo & ~15ull;
        const uint32_t want = (uint32_t)(g_lo & 15u) + PIXB * P.UN;
        if (tid >= SUP_TPB - 9) super_run_meta<false>(metas[mb], P, g, Q.out_stride * f2, T2, SUP_TPB - 1 - tid); // (lanes of the last warp: it has the fewest units)
        mbar_wait(bar, phase);
        phase ^= 1;
        if (a0 + ((want + 15u) & ~15u) > in_limit) { // the last bytes of the buffer that a clipped bulk copy left out (final super-tile of the final frame)
            for (uint32_t i = ((uint32_t)(in_limit - a0) & ~15u) + tid; i < want; i += SUP_TPB) IN[i] = a0 + i < in_limit ? Q.in[a0 + i] : 0;
            __syncthreads();
        }
    };
    auto phase_a = [&](uint32_t st2) {
        const uint32_t f2 = st2 / P.n_tiles, T2 = st2 - f2 * P.n_tiles;
        const uint32_t pad = (uint32_t)((Q.in_stride * f2 + (uint64_t)PIXB * P.UN * T2) & 15u);
#pragma unroll 1
        for (uint32_t hc = warp; 64u * (hc >> 1) < P.UN; hc += SUP_WARPS) { // half-chunk: the units of one parity of a 64-unit chunk
            const uint32_t u = 64u * (hc >> 1) + 2u * lane + (hc & 1u);
            if (u >= P.UN) continue;
            const bool rev = super_row_odd(P, P.UN * T2 + u);               // 26-wide tiles: unit = row
            if constexpr (WORDS) enc_unit_words<true>(IN, pad + 27u * u, S + 26u * u, rev);
            else enc_unit_rgb<true, true>(IN, pad + 18u * u, S + 26u * u, rev); // (the variant that also takes units at byte offset 3 of a word)
        }
    };
    if (blockIdx.x < total) {
        if (tid == 0) fetch(blockIdx.x);
        phase_a_wait(blockIdx.x, 0);
        phase_a(blockIdx.x);
    }
    __syncthreads();                               // S complete, IN free again
    if (blockIdx.x + gridDim.x < total && tid == 0) fetch(blockIdx.x + gridDim.x);
    uint32_t mb = 0;
    for (uint32_t st = blockIdx.x; st < total; st += gridDim.x, mb ^= 1u) {
        const uint32_t f = st / P.n_tiles, T = st - f * P.n_tiles, tm = T % 3u;
        const SuperMeta& meta = metas[mb];
        SUP_TICK0();
        if (super_rows_in_smem(P)) { super_reverse_rows(S, P, T, tid); __syncthreads(); }
        // ---- phase B: one codeword per lane; a pass holds codewords of one k and one scrambler variant
        const uint16_t* map = P.map + tm * (SUP_MAX_PASS * 32);
        const uint8_t* pkv = P.pass_kv + tm * SUP_MAX_PASS;
        const uint32_t n_pass = P.npass[tm];
        uint32_t e_nx = 0, kv_nx = 0; // the map entry of a warp's next pass is fetched while it codes the current one
        if ((uint32_t)warp < n_pass) { e_nx = __ldg(map + 32 * warp + lane); kv_nx = __ldg(pkv + warp); }
#pragma unroll 1
        for (uint32_t pass = warp; pass < n_pass; pass += SUP_WARPS) {
            const uint32_t e = e_nx, kv = kv_nx;
            if (pass + SUP_WARPS < n_pass) { e_nx = __ldg(map + 32 * (pass + SUP_WARPS) + lane); kv_nx = __ldg(pkv + pass + SUP_WARPS); }
            if (e == SUP_IDLE) continue;
            const uint32_t ks = kv & 3u, v = kv >> 2, b = e & 15u, cl = e >> 4, K = P.kk[ks];
            const uint8_t* src = S + 9u * K * cl + b;
            uint8_t* dst = U + meta.stage[b] + 26u * cl;
            uint32_t pa = smem32 + P.off_tab[ks] + v * (8u * K * 27u);
            const uint32_t pnz = pat[6 * ks + 2 * v], ptw = pat[6 * ks + 2 * v + 1];
            if (K == 20) enc_cw<20>(src, dst, pa, pnz, ptw);
            else if (K == 22) enc_cw<22>(src, dst, pa, pnz, ptw);
            else if (K == 24) enc_cw<24>(src, dst, pa, pnz, ptw);
            else enc_cw<18>(src, dst, pa, pnz, ptw);
        }
        __syncthreads();                           // the nine runs complete, S free again
        SUP_TICK(2);
        if (T == 0) { // body symbols 0 and 1 may still see the scrambler's transient (A.4)
            if (tid == 0) {
                uint8_t* dst = U + meta.stage[0];
                dst[0] = gf->scr[g.st[0]][S[0] >> 2];
                dst[1] = gf->scr[g.st[1]][S[9] >> 2];
            }
            __syncthreads();
        }
        // ---- phase C: the nine runs -> global.  Chunk c of a run = 16 bytes at the aligned address a0 + 16c: bytes of the
        // frame at index x = 16c - pad from the run's first byte.  Interior chunks are one gather + one 128-bit store (two
        // gathers merged around the beacon symbol when the chunk holds a beacon slot); the loop is flattened over the nine
        // bands (2^ch_shift slots per band).  It shares its barrier interval with phase A of the next super-tile.
        auto phase_c = [&]() {
            for (uint32_t idx = tid; idx < (9u << P.ch_shift); idx += SUP_TPB) {
                const uint32_t b = idx >> P.ch_shift, c = idx & ((1u << P.ch_shift) - 1u);
                if (c == 0 || c + 1 >= meta.nch[b]) continue;
                const uint64_t lo = meta.g_lo[b];
                const uint32_t padb = (uint32_t)lo & 15u, x0 = 16u * c - padb;
                uint4 q;
                if (!P.G) q = *reinterpret_cast<const uint4*>(U + meta.stage[b] + x0); // staged in the frame's 16-byte phase
                else {
                    const uint32_t t = x0 + P.G - 1u - meta.o_first[b], nb = __umulhi(t, P.G_magic);
                    const uint32_t e = P.G - 1u - (t - nb * P.G);    // distance to the next beacon slot at or after x0
                    const uint32_t sp = U32 + meta.stage[b] + (x0 - nb);
                    q = lds_gather16(sp);
                    if (e < 16u) q = merge16_put(q, lds_gather16(sp - 1u), e, bsym4); // bytes after the slot lag by one
                }
                *reinterpret_cast<uint4*>(Q.out + (lo - padb) + 16u * c) = q;
            }
            // the first and the last chunk of every run may be partial: one byte per thread
            if (tid < 18 * 16) {
                const uint32_t b = (uint32_t)tid >> 5, last = ((uint32_t)tid >> 4) & 1u, i = (uint32_t)tid & 15u;
                const uint64_t lo = meta.g_lo[b];
                const uint32_t padb = (uint32_t)lo & 15u, len = meta.len[b], nch = meta.nch[b], of = meta.o_first[b];
                const uint32_t c = last ? nch - 1u : 0u;
                const int32_t x = (int32_t)(16u * c + i) - (int32_t)padb;
                if (!(last && nch < 2u) && x >= 0 && (uint32_t)x < len) {
                    uint8_t* gout = Q.out + (lo - padb) + 16u * c + i;
                    uint32_t nb = 0;
                    bool slot = false;
                    if (P.G) {
                        const uint32_t t = (uint32_t)x + P.G - 1u - of;
                        nb = __umulhi(t, P.G_magic);
                        slot = t - nb * P.G == P.G - 1u;         // x is a beacon slot
                    }
                    *gout = slot ? (uint8_t)P.bsym : U[meta.stage[b] + (uint32_t)x - nb];
                }
            }
        };
        // half of the warps store first and compute second, the other half the other way round: store latency and arithmetic interleave
        const bool more = st + gridDim.x < total;
        if (more) phase_a_wait(st + gridDim.x, mb ^ 1u);
        if (warp & 1) { if (more) phase_a(st + gridDim.x); phase_c(); }
        else { phase_c(); if (more) phase_a(st + gridDim.x); }
        SUP_TICK(3);
        __syncthreads();                           // S complete, IN and U free again
        SUP_TICK(1);
        if (st + 2 * gridDim.x < total && tid == 0) fetch(st + 2 * gridDim.x); // pixels of the super-tile after the next on their way
    }
}

// ---- one codeword of decode phase B, parity-compare screen (k_fast5.cuh dec_cw5 has the reasoning): the K data symbols' look-ups give the
// parity they imply; scrambled in the plane domain and converted to bytes it is compared with the R received parity symbols (x4 here: the
// squeeze pass has scaled the runs).  Only codewords that differ finish the 26-position syndrome sum and go to the bounded-distance decoder.
// c4 = {chk_nz, chk_two, par_nz, par_two} of the codeword's (k, variant)
template <int K>
static __device__ __noinline__ void dec_cw_dirty_s(uint32_t src_s, uint32_t dst_s, uint32_t acc_nz, uint32_t acc_two, uint32_t tab_s, uint32_t chk_nz, uint32_t chk_two,
                                                   uint32_t sg_s, const uint32_t* chien, uint32_t* status)
{
    const uint8_t* src = smem_ptr(src_s);
    const uint8_t* tab_v = smem_ptr(tab_s);
    Planes d{acc_nz, acc_two};
#pragma unroll 1
    for (int i = K; i < 26; ++i) {
        const uint32_t* row = reinterpret_cast<const uint32_t*>(tab_v + 128 * i + src[i]);   // src holds symbols x4 (< 128)
        gf3_add(d, row[0], row[26 * 32]);
    }
    gf3_add(d, chk_nz, chk_nz ^ chk_two);                      // minus the clean-codeword constant
    if (!(d.nz >> 8)) return;
    uint32_t lo, hi;
    planes_to_parity<K>(d.nz, d.two, lo, hi);
    rs_bd_fix<K>(*reinterpret_cast<const GfTables*>(smem_ptr(sg_s)), chien, smem_ptr(dst_s), lo, hi, status, true);
}
// sa / da / pa / sg_s: .shared addresses (the out-of-line path takes no pointers: k_fast5.cuh dec_cw_mod27)
template <int K>
__device__ __forceinline__ void dec_cw_s(uint32_t sa, uint32_t da, uint32_t pa, const uint4 c4, uint32_t sg_s, const uint32_t* chien, uint32_t* status)
{
    constexpr int PLANE = 4 * 26 * 32, W = K / 4;
    const uint32_t sh = (sa & 2u) * 8u;
    uint32_t xw[7];
    static_for<0, 7>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(xw[j]) : "r"(sa & ~3u), "n"(4 * j) : "memory");
    });
#pragma unroll
    for (int j = 0; j < 6; ++j) xw[j] = __funnelshift_r(xw[j], xw[j + 1], sh);
    xw[6] = (xw[6] >> sh) & 0xFFFFu;
    uint32_t rx_lo, rx_hi = 0;
    if constexpr (K % 4 == 0) {
        rx_lo = xw[W];
        if constexpr (W + 1 < 7) rx_hi = xw[W + 1];
    } else {
        rx_lo = __funnelshift_r(xw[W], xw[W + 1], 16);
        if constexpr (W + 2 < 7) rx_hi = __funnelshift_r(xw[W + 1], xw[W + 2], 16);
    }
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
    static_for<0, K>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        asm volatile("st.shared.u8 [%0+%1], %2;" ::"r"(da), "n"(9 * i), "r"(ev[i]) : "memory");
    });
    gf3_add(acc, acc2.nz, acc2.two);
    Planes s = acc;
    gf3_add(s, c4.z, c4.w);
    uint32_t lo, hi;
    planes_to_parity<K>(s.nz, s.two, lo, hi);
    if ((26 - K > 4) ? ((((lo << 2) ^ rx_lo) | ((hi << 2) ^ rx_hi)) != 0u) : ((lo << 2) != rx_lo))
        dec_cw_dirty_s<K>(sa, da, acc.nz, acc.two, pa, c4.x, c4.y, sg_s, chien, status);
}

// ---- decode phase A of the super-tile kernel: a unit (a row of a 26-wide 2D tile when rev) -> six pixels -> 18 RGB bytes, with the pixel
// arithmetic of the v5 warp-tile decoder (values_to_rgb18: table chroma, fixed-point sums, PRMT gather)
__device__ __forceinline__ void dec_unit_rgb_s(const uint8_t* S, uint32_t a, uint8_t* dst, bool rev, const uint8_t* __restrict__ clut)
{
    uint32_t y[7];
    load_unit26<true>(S, a, y, rev);
    uint32_t A[6], w[5];
    symbols_to_triple(y[0], y[1], y[2], y[3] & 0xFF, A[0], A[1], A[2]);
    symbols_to_triple(__funnelshift_r(y[3], y[4], 8), __funnelshift_r(y[4], y[5], 8), __funnelshift_r(y[5], y[6], 8), (y[6] >> 8) & 0xFF, A[3], A[4], A[5]);
    values_to_rgb18(A, clut, w);
    const uint32_t da = smem_u32(dst), odd = da & 2u, sh = odd << 3;
    uint32_t* d = reinterpret_cast<uint32_t*>(dst + odd);
#pragma unroll
    for (int j = 0; j < 4; ++j) d[j] = __funnelshift_r(w[j], w[j + 1], sh);
    *reinterpret_cast<uint16_t*>(dst + (odd ? 0 : 16)) = (uint16_t)(odd ? w[0] : w[4]);
}
// =============================================================================================
// decode
// =============================================================================================
template <bool WORDS>
__global__ void __launch_bounds__(SUP_TPB, 2) k_decode_super(FastParams Q, SuperPlan P, Geom g, const GfTables* __restrict__ gf, const RsTables* __restrict__ rs)
{
    constexpr uint32_t PIXB = WORDS ? 27u : 18u;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((256u - (smem_u32(smem_raw) & 255u)) & 255u); // the variant blocks are 256-byte aligned (dec_cw)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* S = smem + P.off_S;                   // the runs as they lie in the frame, later the descrambled stream symbols
    uint8_t* R = smem + P.off_U;                   // the nine pre-beacon runs (x4), later the pixel-side bytes on their way out
    SuperMeta& meta = *reinterpret_cast<SuperMeta*>(smem + P.off_meta);
    uint32_t* chk = reinterpret_cast<uint32_t*>(smem + P.off_aux); // [k slot][variant]{chk_nz, chk_two, par_nz, par_two}
    uint8_t* clut = smem + P.off_aux + 192;        // dequantised chroma of the 81 quantised values (values_to_rgb18)
    if (tid < 96) clut[tid] = (uint8_t)min((32u * (uint32_t)tid + 5u) / 10u, 255u);
    GfTables& sg = *reinterpret_cast<GfTables*>(smem + P.off_gf);
    for (uint32_t ks = 0; ks < P.nk; ++ks) {
        const int K = (int)P.kk[ks];
        const uint32_t(*pl)[kVals][2] = rs->pl[1][(24 - K) / 2]; // the consistent decoder always uses the repaired code
        for (int idx = tid; idx < 3 * 26 * 32; idx += SUP_TPB) {
            const int v = idx / (26 * 32), rem = idx - v * (26 * 32), i = rem / 32, x = rem - 32 * i, xm = x >= 27 ? x - 27 : x;
            uint32_t* blk = reinterpret_cast<uint32_t*>(smem + P.off_tab[ks] + v * 6656);
            blk[rem] = pl[i][xm][0] | gf->dsc[st_of(g, v, i)][xm];
            blk[26 * 32 + rem] = pl[i][xm][1];
        }
    }
    load_gf(sg, gf);
    __syncthreads();
    if (tid < 3 * (int)P.nk) { // a received block r = c (+) 13*st is a codeword iff sum_i T_i[r_i] == sum_i T_i[13*st_i]
        const int ks = tid / 3, v = tid - 3 * ks;
        Planes c{0, 0};
        const uint32_t* blk = reinterpret_cast<const uint32_t*>(smem + P.off_tab[ks] + v * 6656);
        for (int i = 0; i < 26; ++i) {
            const int idx = i * 32 + 13 * (int)st_of(g, v, i);
            gf3_add(c, blk[idx] & ~0xFFu, blk[26 * 32 + idx]);
        }
        chk[4 * tid] = c.nz;
        chk[4 * tid + 1] = c.two;
        // the parity-compare constant of dec_cw_s (k_v5_image_dec derives it): (scrambler pattern of the parity positions) - sum_{i<K} T_i[13*st_i]
        const int K = (int)P.kk[ks];
        Planes e{0, 0};
        for (int i = 0; i < K; ++i) {
            const int idx = i * 32 + 13 * (int)st_of(g, v, i);
            gf3_add(e, blk[idx] & ~0xFFu, blk[26 * 32 + idx]);
        }
        e.two ^= e.nz;
        uint32_t pn = 0, pt = 0;
        for (int j = 0; j < 26 - K; ++j) {
            const uint32_t st = st_of(g, v, K + j);
            const int sh = 26 - K <= 6 ? 8 + 4 * j : 8 + 3 * j;   // plane_shift<K>(j)
            if (st) pn |= 7u << sh;
            if (st == 2) pt |= 7u << sh;
        }
        gf3_add(e, pn, pt);
        chk[4 * tid + 2] = e.nz;
        chk[4 * tid + 3] = e.two;
    }
    const uint32_t smem32 = smem_u32(smem);
    const uint64_t in_limit = Q.in_stride * (Q.n_frames - 1) + 9 * g.n_out;
    const uint32_t total = P.n_tiles * Q.n_frames;
    const uint32_t ch_mask = (1u << P.ch_shift) - 1u, S32 = smem_u32(S);
    // the nine runs of the super-tile described by `meta`, as they lie in the frame -> S region (slot raw_base[b], byte i <-> global a0 + i):
    // 16-byte asynchronous copies (LDGSTS), issued as soon as S is free so that they land while the previous super-tile is stored
    auto load_runs = [&]() {
        for (uint32_t idx = tid; idx < (9u << P.ch_shift); idx += SUP_TPB) {
            const uint32_t b = idx >> P.ch_shift, c = idx & ch_mask;
            if (c >= meta.nch[b]) continue;
            const uint64_t ga = (meta.g_lo[b] & ~15ull) + 16ull * c;
            uint8_t* dstp = S + P.raw_base[b] + 16u * c;
            if (ga + 16 <= in_limit) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dstp)), "l"(Q.in + ga) : "memory");
            else for (int i = 0; i < 16; ++i) dstp[i] = ga + i < in_limit ? Q.in[ga + i] : 0;   // the end of the buffer
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (blockIdx.x < total && tid < 9) super_run_meta<true>(meta, P, g, Q.in_stride * (blockIdx.x / P.n_tiles), blockIdx.x % P.n_tiles, tid);
    __syncthreads();
    if (blockIdx.x < total) load_runs();
    for (uint32_t st = blockIdx.x; st < total; st += gridDim.x) {
        const uint32_t f = st / P.n_tiles, T = st - f * P.n_tiles, tm = T % 3u, n_pass = P.npass[tm];
        SUP_TICK0();
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                           // this super-tile's runs are in S; the previous one's pixels have left R
        SUP_TICK(4);
        const uint16_t* map = P.map + tm * (SUP_MAX_PASS * 32);
        const uint8_t* pkv = P.pass_kv + tm * SUP_MAX_PASS;
        uint32_t e_nx = 0, kv_nx = 0;              // the map entry of a warp's next phase-B pass is always fetched one pass (here: one phase) ahead
        if ((uint32_t)warp < n_pass) { e_nx = __ldg(map + 32 * warp + lane); kv_nx = __ldg(pkv + warp); }
        // ---- squeeze the beacon slots out and scale by 4 (table byte offset): R_b[q] = 4 * frame byte (q + beacon slots before
        // body symbol q).  Bytes >= 27 are reduced mod 27 first (out-of-alphabet symbols read as their low three trits, OLD:28-31)
        for (uint32_t idx = tid; idx < (9u << P.ch_shift); idx += SUP_TPB) {
            const uint32_t b = idx >> P.ch_shift, d = idx & ch_mask, L = 26u * P.ncw[P.kslot[b]];
            if (16u * d >= L) continue;
            const uint32_t q0 = 16u * d;
            const uint32_t srcb = S32 + P.raw_base[b] + ((uint32_t)meta.g_lo[b] & 15u);
            uint4 q;
            if (!P.G) q = lds_gather16(srcb + q0);
            else {
                const uint32_t t = q0 + P.G - 1u - meta.o_first[b], nb = __umulhi(t, P.Gm1_magic);
                const uint32_t e = P.G - 1u - (t - nb * (P.G - 1u)); // distance to the next body symbol with one more slot before it
                q = lds_gather16(srcb + q0 + nb);
                if (e < 16u) q = merge16(q, lds_gather16(srcb + q0 + nb + 1u), e);
            }
            if (((q.x | q.y | q.z | q.w) & 0xE0E0E0E0u) != 0) {
                uint32_t t[4] = {q.x, q.y, q.z, q.w};
                for (int i = 0; i < 4; ++i) {
                    uint32_t r = 0;
                    for (int j = 0; j < 4; ++j) r |= (((t[i] >> (8 * j)) & 0xFFu) % 27u) << (8 * j);
                    t[i] = r;
                }
                q = make_uint4(t[0], t[1], t[2], t[3]);
            }
            q.x <<= 2; q.y <<= 2; q.z <<= 2; q.w <<= 2;
            *reinterpret_cast<uint4*>(R + P.run_base[b] + q0) = q;
        }
        __syncthreads();                           // R complete, the raw runs in S and this tile's meta dead
        SUP_TICK(5);
        if (st + gridDim.x < total) { // this CTA's next super-tile: its run geometry now, its runs towards L2
            const uint32_t st2 = st + gridDim.x, f2 = st2 / P.n_tiles, T2 = st2 - f2 * P.n_tiles;
            if (tid < 9) super_run_meta<true>(meta, P, g, Q.in_stride * f2, T2, tid);
            if (tid >= 32 && tid < 32 + 9 * 32) { // warp 1 + b fetches band b's lines
                const uint32_t b = (uint32_t)(tid - 32) >> 5, n = P.ncw[P.kslot[b]];
                const uint64_t p_lo = 26ull * (g.cw_base[b] + (uint64_t)n * T2);
                const uint64_t o_lo = P.G ? p_lo + p_lo / (P.G - 1u) : p_lo; // within a few bytes of the run's start: good enough for a prefetch
                const uint64_t lo = Q.in_stride * f2 + 52 + o_lo, a0 = lo & ~127ull;
                const uint32_t nl = (uint32_t)((lo + 26u * n + 26u * n / 26u + 255u - a0) >> 7);
                for (uint32_t i = lane; i < nl; i += 32) if (a0 + 128ull * i < in_limit) prefetch_l2(Q.in + a0 + 128ull * i);
            }
        }
        if (T == 0) { // body symbols 0,1: move them from the transient states to the periodic ones
            if (tid == 0) {
                uint8_t* r0 = R + P.run_base[0];
                r0[0] = (uint8_t)(4u * sg.scr[st_of(g, 0, 0)][sg.dsc[g.st[0]][(r0[0] >> 2) % 27u]]);
                r0[1] = (uint8_t)(4u * sg.scr[st_of(g, 0, 1)][sg.dsc[g.st[1]][(r0[1] >> 2) % 27u]]);
            }
            __syncthreads();
        }
        // ---- phase B: syndrome screen per codeword (BM / Chien / Forney in-thread for the dirty ones), descrambled data -> stream order
#pragma unroll 1
        for (uint32_t pass = warp; pass < n_pass; pass += SUP_WARPS) {
            const uint32_t e = e_nx, kv = kv_nx;
            if (pass + SUP_WARPS < n_pass) { e_nx = __ldg(map + 32 * (pass + SUP_WARPS) + lane); kv_nx = __ldg(pkv + pass + SUP_WARPS); }
            if (e == SUP_IDLE) continue;
            const uint32_t ks = kv & 3u, v = kv >> 2, b = e & 15u, cl = e >> 4, K = P.kk[ks];
            const uint32_t src = smem32 + P.off_U + P.run_base[b] + 26u * cl;
            const uint32_t dst = S32 + 9u * K * cl + b;
            const uint32_t toff = P.off_tab[ks] + v * 6656u, sg32 = smem32 + P.off_gf;
            const uint4 c4 = *reinterpret_cast<const uint4*>(chk + 4 * (3 * ks + v));
            if (K == 20) dec_cw_s<20>(src, dst, smem32 + toff, c4, sg32, chien_of(gf), Q.status + 2 * f);
            else if (K == 22) dec_cw_s<22>(src, dst, smem32 + toff, c4, sg32, chien_of(gf), Q.status + 2 * f);
            else if (K == 24) dec_cw_s<24>(src, dst, smem32 + toff, c4, sg32, chien_of(gf), Q.status + 2 * f);
            else dec_cw_s<18>(src, dst, smem32 + toff, c4, sg32, chien_of(gf), Q.status + 2 * f);
        }
        __syncthreads();                           // S complete, R dead
        SUP_TICK(6);
        if (super_rows_in_smem(P)) { super_reverse_rows(S, P, T, tid); __syncthreads(); }
        // ---- phase A: 26 stream symbols -> six pixels per thread -> pixel-side bytes in R
        const uint64_t g_lo = Q.out_stride * f + (uint64_t)PIXB * P.UN * T;
        const uint32_t pad = (uint32_t)(g_lo & 15u);
#pragma unroll 1
        for (uint32_t hc = warp; 64u * (hc >> 1) < P.UN; hc += SUP_WARPS) { // half-chunk: the units of one parity of a 64-unit chunk
            {
                const uint32_t u = 64u * (hc >> 1) + 2u * lane + (hc & 1u);
                if (u >= P.UN) continue;
                const bool rev = super_row_odd(P, P.UN * T + u);                // 26-wide tiles: unit = row
                if constexpr (WORDS) dec_unit_words<true>(S, 26u * u, R + pad + 27u * u, rev);
                else dec_unit_rgb_s(S, 26u * u, R + pad + 18u * u, rev, clut);   // pad is even (frames start on even bytes, 18 UN T is even)
            }
        }
        __syncthreads();
        SUP_TICK(7);
        if (st + gridDim.x < total) load_runs();   // S is free since phase A: the next super-tile's runs (geometry computed after the squeeze)
        {   // pixel-side bytes -> global: whole 16-byte chunks, the (at most 15 + 15) edge bytes one by one
            const uint32_t n_out = PIXB * P.UN, end = pad + n_out, c_lo = pad ? 1u : 0u, c_hi = end >> 4;
            uint8_t* gout = Q.out + (g_lo - pad);
            for (uint32_t c = c_lo + tid; c < c_hi; c += SUP_TPB) *reinterpret_cast<uint4*>(gout + 16u * c) = *reinterpret_cast<const uint4*>(R + 16u * c);
            if (tid < 16 && pad && (uint32_t)tid >= pad && (uint32_t)tid < end) gout[tid] = R[tid];
            if (tid >= 32 && tid < 48) { const uint32_t pos = 16u * c_hi + (uint32_t)(tid - 32); if (pos < end && (c_hi >= c_lo) && !(c_hi == 0 && pad)) gout[pos] = R[pos]; }
        }
        SUP_TICK(8);                               // (the barrier at the top of the loop separates this store from the next squeeze)
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// =============================================================================================
// host side: plan (geometry, shared-memory layout, pass maps) and launchers
// =============================================================================================
static uint32_t gcd_u32(uint32_t a, uint32_t b) { while (b) { const uint32_t t = a % b; a = b; b = t; } return a; }
static uint32_t lcm_u32(uint32_t a, uint32_t b) { return a / gcd_u32(a, b) * b; }
static uint32_t up16(uint32_t x) { return (x + 15u) & ~15u; }
static uint32_t up256(uint32_t x) { return (x + 255u) & ~255u; }

// the configs the super-tile kernels take: any per-band k, 2D tile widths that divide 26,
// beacon periods that leave at most one slot per 16-byte chunk and fit the 32-bit reciprocal
bool super_config_ok(const t3c_config& cfg)
{
    if (cfg.profile == T3C_PROFILE_RAW) return false;
    static const int ks[4] = {24, 22, 20, 18};
    if (use_2d(cfg) && (cfg.tile_w > 26 || 26 % cfg.tile_w != 0)) return false;
    if (use_beacon(cfg) && cfg.beacon_slot < 9 && (cfg.beacon_period < 3 || cfg.beacon_period > 255)) return false;
    return true;
}
// false when no super-tile shape fits (shared memory, pass table) or the frame holds no full super-tile
static bool make_super_plan(const t3c_config& cfg, const Geom& g, bool decode, bool words, uint64_t px_limit, SuperPlan& P)
{
    if (!super_config_ok(cfg)) return false;
    if (g.n_s >= (1ull << 31) || g.l_exp >= (1ull << 31)) return false;
    std::memset(&P, 0, sizeof P);
    uint32_t l = 26;
    for (int b = 0; b < 9; ++b) {
        uint32_t s = 0;
        while (s < P.nk && P.kk[s] != (uint32_t)g.k[b]) ++s;
        if (s == P.nk) P.kk[P.nk++] = (uint32_t)g.k[b];
        P.kslot[b] = (uint8_t)s;
        l = lcm_u32(l, (uint32_t)g.k[b]);
    }
    if (g.period && g.slot >= 0) {
        P.G = 9u * g.period; P.slot = (uint32_t)g.slot; P.bsym = g.bsym;
        P.G_magic = (uint32_t)((1ull << 32) / P.G) + 1u;
        P.Gm1_magic = (uint32_t)((1ull << 32) / (P.G - 1u)) + 1u;
    }
    if (g.tile_area) { P.tile_w = g.tile_w; P.tile_area = (uint32_t)g.tile_area; P.tile_h26 = (g.tile_w == 26 && g.tile_area > 26 && (g.n_s / 26 + 1) * (g.tile_area / 26) < (1ull << 32)) ? (uint32_t)(g.tile_area / 26) : 0u; if (P.tile_h26) P.h26_magic = (uint32_t)((1ull << 32) / P.tile_h26) + 1u; }
    const uint32_t pixb = words ? 27u : 18u;
    // the largest multiple of l that fits two CTAs per SM; one CTA per SM when even l itself does not (k = 24/22: M = 3432)
    bool found = false;
    for (uint32_t per_sm = 2; per_sm >= 1 && !found; --per_sm)
    for (uint32_t mult = 16; mult >= 1 && !found; --mult) {
        const uint32_t budget = (227u * 1024u - 1024u * per_sm) / per_sm - (decode ? 256u : 0u);
        P.ctas_per_sm = per_sm;
        const uint32_t M = l * mult;
        if (9u * M % 26u) continue;
        uint32_t n_cw = 0, off = 0, runs = 0, raws = 0;
        for (uint32_t s = 0; s < P.nk; ++s) P.ncw[s] = M / P.kk[s];
        for (int b = 0; b < 9; ++b) {
            const uint32_t n = P.ncw[P.kslot[b]], L = 26u * n;
            n_cw += n;
            P.run_base[b] = runs; runs += up16(L + 32u);
            P.raw_base[b] = raws; raws += up16(L + (P.G ? L / (P.G - 1u) + 2u : 0u) + 48u);
        }
        if (M / 20u >= 4095u || n_cw / 32u + 3u * P.nk > SUP_MAX_PASS) continue;
        for (uint32_t s = 0; s < P.nk; ++s) { P.off_tab[s] = off; off += decode ? 3u * 6656u : up256(3u * 8u * P.kk[s] * 27u); }
        P.off_aux = off; off += up16(decode ? 4u * 3u * 16u + 96u : 4u * 6u * 4u + 16u);   // encoder: pat[4][3][2] + the mbarrier of its input buffer; decoder: [4][3]{chk, par}
        P.off_gf = off; off += decode ? up16((uint32_t)sizeof(GfTables)) : 0u;
        P.off_meta = off; off += 2u * up16((uint32_t)sizeof(SuperMeta));
        const uint32_t UN = 9u * M / 26u, pix = up16(pixb * UN + 48u);
        P.off_S = off; off += up16((decode && raws > 9u * M ? raws : 9u * M) + 32u);
        if (decode) { P.off_U = off; off += runs > pix ? runs : pix; P.off_IN = P.off_U; }   // R, later the pixel-side bytes on their way out
        else { P.off_U = off; off += runs; P.off_IN = off; off += pix; }                    // the runs | the pixels (bulk-loaded one super-tile ahead)
        if (off > budget) continue;
        P.smem_bytes = off + (decode ? 256u : 0u);
        P.M = M; P.UN = UN;
        uint32_t max_len = 0;
        for (uint32_t s = 0; s < P.nk; ++s) max_len = 26u * P.ncw[s] > max_len ? 26u * P.ncw[s] : max_len;
        const uint32_t max_exp = max_len + (P.G ? max_len / (P.G - 1u) + 2u : 0u);           // run length in the frame, beacon slots included
        const uint32_t max_ch = (max_exp + 30u) / 16u + 1u;
        P.ch_shift = 0; while ((1u << P.ch_shift) < max_ch) ++P.ch_shift;
        found = true;
    }
    if (!found) return false;
    uint64_t nt = ~0ull;
    for (int b = 0; b < 9; ++b) { const uint64_t t = g.ncw[b] / P.ncw[P.kslot[b]]; nt = t < nt ? t : nt; }
    const uint64_t by_px = px_limit / (6ull * P.UN);
    nt = by_px < nt ? by_px : nt;
    if (!nt || nt > 0x7FFFFFFFull) return false;
    P.n_tiles = (uint32_t)nt;
    return true;
}
// Pass maps (built when the cached ones do not match): for super-tile index T = tm (mod 3), codeword cl of band b has scrambler
// variant (cw_base_b + n_b T + cl) mod 3; a pass takes up to 32 codewords of one (k, variant).  Phase B gathers a codeword's
// symbols at byte addresses a + 9i, a = 9 k cl + b: two lanes collide on a bank for some i exactly when their a differ by more
// than 3 and by 0, +-1, +-2, +-3 modulo 128, so codewords are dealt greedily into the first pass where they collide with nobody.
static bool build_super_maps(const SuperPlan& P, const Geom& g, uint16_t* h_map, uint8_t* h_kv, uint32_t npass[3])
{
    struct Pass { int n; int32_t occ[128]; };
    std::vector<Pass> passes;
    for (int tm = 0; tm < 3; ++tm) {
        uint32_t pass0 = 0;
        for (uint32_t s = 0; s < P.nk; ++s)
            for (uint32_t v = 0; v < 3; ++v) {
                std::vector<std::pair<uint32_t, int32_t>> list; // (map entry, gather address)
                for (uint32_t cl = 0; cl < P.ncw[s]; ++cl)
                    for (uint32_t b = 0; b < 9; ++b)
                        if (P.kslot[b] == s && (g.cw_base[b] + (uint64_t)P.ncw[s] * tm + cl) % 3 == v) list.push_back({b | cl << 4, (int32_t)(9u * P.kk[s] * cl + b)});
                const uint32_t np = ((uint32_t)list.size() + 31u) / 32u;
                if (pass0 + np > SUP_MAX_PASS) return false;
                passes.assign(np, Pass{});
                for (auto& p : passes) { p.n = 0; for (int i = 0; i < 128; ++i) p.occ[i] = -1; }
                for (uint32_t i = 0; i < np * 32u; ++i) h_map[(tm * SUP_MAX_PASS + pass0) * 32 + i] = (uint16_t)SUP_IDLE;
                for (uint32_t p = 0; p < np; ++p) h_kv[tm * SUP_MAX_PASS + pass0 + p] = (uint8_t)(s | v << 2);
                auto clash = [](const Pass& p, int32_t a) {
                    for (int d = -3; d <= 3; ++d) {
                        const int32_t o = p.occ[((a + d) % 128 + 128) % 128];
                        if (o >= 0 && (o - a > 3 || a - o > 3)) return true;
                    }
                    return false;
                };
                for (const auto& e : list) {
                    int best = -1;
                    for (uint32_t p = 0; p < np && best < 0; ++p) if (passes[p].n < 32 && !clash(passes[p], e.second)) best = (int)p;
                    for (uint32_t p = 0; p < np && best < 0; ++p) if (passes[p].n < 32) best = (int)p; // no clean place left: take the clash
                    Pass& q = passes[best];
                    h_map[(tm * SUP_MAX_PASS + pass0 + best) * 32 + q.n++] = (uint16_t)e.first;
                    q.occ[(e.second % 128 + 128) % 128] = e.second;
                }
                pass0 += np;
            }
        npass[tm] = pass0;
    }
    return true;
}
