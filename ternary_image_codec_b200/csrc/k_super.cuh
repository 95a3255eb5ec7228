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

struct SuperPlan {
    uint32_t M, UN, n_tiles, nk;
    uint32_t kk[4];         // the distinct k values in use
    uint32_t ncw[4];        // codewords per band and super-tile for k slot s: M / kk[s]
    uint8_t kslot[12];      // band -> k slot
    uint32_t npass[3];      // phase-B passes per class of the super-tile index mod 3
    uint32_t run_base[9];   // staging slot (16-byte aligned, + 32 bytes of slack) of band b's pre-beacon run in U / R
    uint32_t raw_base[9];   // decode: slot of band b's run as it lies in the frame (beacon slots included)
    uint32_t off_tab[4], off_aux, off_gf, off_meta, off_S, off_U, smem_bytes;
    uint32_t G, G_magic, Gm1_magic; // beacon: 9 * period (0: none), floor(2^32 / G) + 1, floor(2^32 / (G - 1)) + 1
    uint32_t slot, bsym;
    uint32_t tile_w, tile_area, tile_h26; // tile_h26: tile height when the width is 26 (rows = units: reversed in registers), else 0
    uint32_t ch_shift, sl_shift;          // log2 of the per-band slot counts of the flattened chunk loops (16-byte chunks / byte-wise chunks)
    const uint16_t* map;    // [3][SUP_MAX_PASS * 32]: b | cl << 4, SUP_IDLE = idle lane
    const uint8_t* pass_kv; // [3][SUP_MAX_PASS]: k slot | variant << 2
};
struct SuperMeta {          // per super-tile, written by threads 0..8
    uint64_t g_lo[9];       // global byte offset of the first byte of band b's run in the frame
    uint32_t len[9];        // bytes of the run in the frame (beacon slots inside it included)
    uint32_t o_first[9];    // index, from the run's first byte, of the first beacon slot at or after it
    uint32_t stage[9];      // offset of the run's pre-beacon byte 0 inside U / R
    uint32_t nch[9];        // 16-byte chunks of the aligned superset of the run in the frame
    uint32_t n_slow[9];     // encode: 2 + beacon slots inside the run; decode: slot-count boundaries inside the padded pre-beacon run
};

// 16 bytes at an arbitrary byte offset of a shared buffer (the buffer has >= 4 bytes of slack after the last byte read)
__device__ __forceinline__ uint4 lds_gather16(const uint8_t* base, uint32_t off)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(base) + off;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const uint32_t sh = ((uint32_t)a & 3u) * 8u;
    const uint32_t x0 = w[0], x1 = w[1], x2 = w[2], x3 = w[3], x4 = w[4];
    return make_uint4(__funnelshift_r(x0, x1, sh), __funnelshift_r(x1, x2, sh), __funnelshift_r(x2, x3, sh), __funnelshift_r(x3, x4, sh));
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
    if (DECODE) {
        m.stage[b] = P.run_base[b];
        const uint32_t lp = (L + 15u) & ~15u; // boundaries are counted up to the end of the last 16-byte chunk
        m.n_slow[b] = (P.G && lp > o_first) ? (lp - 1u - o_first) / (P.G - 1u) + 1u : 0u;
    } else {
        m.stage[b] = P.run_base[b] + (P.G ? 0u : (uint32_t)lo & 15u); // without a beacon the staged run keeps the frame's 16-byte phase
        m.n_slow[b] = 2u + ((P.G && len > o_first) ? (len - 1u - o_first) / P.G + 1u : 0u);
    }
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
    uint8_t* U = smem + P.off_U;                   // the pixel-side bytes of the super-tile, later the nine pre-beacon runs
    SuperMeta& meta = *reinterpret_cast<SuperMeta*>(smem + P.off_meta);
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
                if (st) nz |= 7u << (8 + 4 * j);
                if (st == 2) two |= 7u << (8 + 4 * j);
            }
            pat[6 * ks + 2 * tid] = nz;
            pat[6 * ks + 2 * tid + 1] = two;
        }
    }
    __syncthreads();
    const uint32_t smem32 = smem_u32(smem);
    const uint64_t in_limit = Q.in_stride * Q.n_frames;
    const uint32_t total = P.n_tiles * Q.n_frames;
    for (uint32_t st = blockIdx.x; st < total; st += gridDim.x) {
        const uint32_t f = st / P.n_tiles, T = st - f * P.n_tiles, tm = T % 3u;
        // ---- the super-tile's pixels -> U (128-bit loads of the 16-byte aligned superset)
        const uint64_t g_lo = Q.in_stride * f + (uint64_t)PIXB * P.UN * T, a0 = g_lo & ~15ull;
        const uint32_t pad = (uint32_t)(g_lo & 15u), n_in = (pad + PIXB * P.UN + 15u) >> 4;
        for (uint32_t c = tid; c < n_in; c += SUP_TPB) {
            const uint64_t ga = a0 + 16ull * c;
            uint4 q;
            if (ga + 16 <= in_limit) q = __ldg(reinterpret_cast<const uint4*>(Q.in + ga));
            else {
                uint32_t t[4] = {0, 0, 0, 0};
                for (int i = 0; i < 16; ++i) if (ga + i < in_limit) t[i >> 2] |= (uint32_t)Q.in[ga + i] << (8 * (i & 3));
                q = make_uint4(t[0], t[1], t[2], t[3]);
            }
            *reinterpret_cast<uint4*>(U + 16 * c) = q;
        }
        if (st + gridDim.x < total) { // this CTA's next super-tile: its pixels towards L2
            const uint32_t st2 = st + gridDim.x, f2 = st2 / P.n_tiles, T2 = st2 - f2 * P.n_tiles;
            super_prefetch(Q.in, Q.in_stride * f2 + (uint64_t)PIXB * P.UN * T2, PIXB * P.UN, in_limit, tid);
        }
        if (tid < 9) super_run_meta<false>(meta, P, g, Q.out_stride * f, T, tid);
        __syncthreads();
        // ---- phase A: thread per unit, dealt even / odd inside chunks of 64 units (conflict-free 36- and 52-byte lane strides)
        for (uint32_t ch = warp; 64u * ch < P.UN; ch += SUP_WARPS) {
#pragma unroll 1
            for (uint32_t h = 0; h < 2; ++h) {
                const uint32_t u = 64u * ch + 2u * lane + h;
                if (u >= P.UN) continue;
                const bool rev = P.tile_h26 && ((P.UN * T + u) % P.tile_h26 & 1u); // 26-wide tiles: unit = row
                if constexpr (WORDS) enc_unit_words<true>(U, pad + 27u * u, S + 26u * u, rev);
                else if (pad & 1u) enc_unit_rgb<true, true>(U, pad + 18u * u, S + 26u * u, rev);
                else enc_unit_rgb<false, true>(U, pad + 18u * u, S + 26u * u, rev);
            }
        }
        __syncthreads();                           // S complete, U (pixels) dead
        if (super_rows_in_smem(P)) { super_reverse_rows(S, P, T, tid); __syncthreads(); }
        // ---- phase B: one codeword per lane; a pass holds codewords of one k and one scrambler variant
        const uint16_t* map = P.map + tm * (SUP_MAX_PASS * 32);
        const uint8_t* pkv = P.pass_kv + tm * SUP_MAX_PASS;
#pragma unroll 1
        for (uint32_t pass = warp; pass < P.npass[tm]; pass += SUP_WARPS) {
            const uint32_t e = __ldg(map + 32 * pass + lane), kv = __ldg(pkv + pass);
            if (e == SUP_IDLE) continue;
            const uint32_t ks = kv & 3u, v = kv >> 2, b = e & 15u, cl = e >> 4, K = P.kk[ks];
            const uint8_t* src = S + 9u * K * cl + b;
            uint8_t* dst = U + meta.stage[b] + 26u * cl;
            uint32_t pa = smem32 + P.off_tab[ks] + v * (8u * K * 27u);
            const uint32_t pnz = pat[6 * ks + 2 * v], ptw = pat[6 * ks + 2 * v + 1];
            if (K == 20) enc_cw<20>(src, dst, pa, pnz, ptw);
            else if (K == 22) enc_cw<22>(src, dst, pa, pnz, ptw);
            else enc_cw<24>(src, dst, pa, pnz, ptw);
        }
        __syncthreads();
        if (T == 0) { // body symbols 0 and 1 may still see the scrambler's transient (A.4)
            if (tid == 0) {
                uint8_t* dst = U + meta.stage[0];
                dst[0] = gf->scr[g.st[0]][S[0] >> 2];
                dst[1] = gf->scr[g.st[1]][S[9] >> 2];
            }
            __syncthreads();
        }
        // ---- phase C: the nine runs -> global.  Chunk c of a run = 16 bytes at the aligned address a0 + 16c: bytes of the
        // frame at index x = 16c - pad from the run's first byte.  Interior chunks without a beacon slot are one gather +
        // one 128-bit store; the first / last chunk of a run and the chunks holding a beacon slot go byte by byte.
        // Both loops are flattened over the nine bands (2^ch_shift / 2^sl_shift slots per band).
        for (uint32_t idx = tid; idx < (9u << P.ch_shift); idx += SUP_TPB) {
            const uint32_t b = idx >> P.ch_shift, c = idx & ((1u << P.ch_shift) - 1u);
            if (c == 0 || c + 1 >= meta.nch[b]) continue;
            const uint64_t lo = meta.g_lo[b];
            const uint32_t padb = (uint32_t)lo & 15u, x0 = 16u * c - padb;
            uint32_t nb = 0;
            if (P.G) {
                const uint32_t t = x0 + P.G - 1u - meta.o_first[b];
                nb = __umulhi(t, P.G_magic);
                if (P.G - 1u - (t - nb * P.G) < 16u) continue; // a beacon slot inside: second loop
            }
            const uint8_t* sp = U + meta.stage[b] + (x0 - nb);
            uint4 q;
            if (!P.G) q = *reinterpret_cast<const uint4*>(sp); // staged in the frame's 16-byte phase
            else q = lds_gather16(sp, 0);
            *reinterpret_cast<uint4*>(Q.out + (lo - padb) + 16u * c) = q;
        }
        // slow items per band: 0 = first chunk, 1 = last chunk, 2 + j = the chunk of beacon slot j (unless it is the first / last one)
        for (uint32_t idx = tid; idx < (9u << P.sl_shift); idx += SUP_TPB) {
            const uint32_t b = idx >> P.sl_shift, it = idx & ((1u << P.sl_shift) - 1u);
            if (it >= meta.n_slow[b]) continue;
            const uint64_t lo = meta.g_lo[b];
            const uint32_t padb = (uint32_t)lo & 15u, len = meta.len[b], nch = meta.nch[b], of = meta.o_first[b];
            uint32_t c;
            if (it == 0) c = 0;
            else if (it == 1) { if (nch < 2) continue; c = nch - 1; }
            else { c = (of + P.G * (it - 2u) + padb) >> 4; if (c == 0 || c + 1 == nch) continue; }
            uint8_t* gout = Q.out + (lo - padb) + 16u * c;
            const uint8_t* sb = U + meta.stage[b];
            for (uint32_t i = 0; i < 16; ++i) {
                const int32_t x = (int32_t)(16u * c + i) - (int32_t)padb;
                if (x < 0 || (uint32_t)x >= len) continue;
                uint32_t nb = 0;
                if (P.G) {
                    const uint32_t t = (uint32_t)x + P.G - 1u - of;
                    nb = __umulhi(t, P.G_magic);
                    if (t - nb * P.G == P.G - 1u) { gout[i] = (uint8_t)P.bsym; continue; } // x is a beacon slot
                }
                gout[i] = sb[(uint32_t)x - nb];
            }
        }
        __syncthreads();                           // U and S are reused by the next super-tile
    }
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
    uint32_t* chk = reinterpret_cast<uint32_t*>(smem + P.off_aux); // [k slot][variant][2]
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
        chk[2 * tid] = c.nz;
        chk[2 * tid + 1] = c.two;
    }
    const uint32_t smem32 = smem_u32(smem);
    const uint64_t in_limit = Q.in_stride * (Q.n_frames - 1) + 9 * g.n_out;
    const uint32_t total = P.n_tiles * Q.n_frames;
    const uint32_t ch_mask = (1u << P.ch_shift) - 1u, sl_mask = (1u << P.sl_shift) - 1u;
    if (blockIdx.x < total && tid < 9) super_run_meta<true>(meta, P, g, Q.in_stride * (blockIdx.x / P.n_tiles), blockIdx.x % P.n_tiles, tid);
    __syncthreads();
    for (uint32_t st = blockIdx.x; st < total; st += gridDim.x) {
        const uint32_t f = st / P.n_tiles, T = st - f * P.n_tiles, tm = T % 3u;
        // ---- the nine runs as they lie in the frame -> S region (slot raw_base[b], byte i <-> global a0 + i)
        for (uint32_t idx = tid; idx < (9u << P.ch_shift); idx += SUP_TPB) {
            const uint32_t b = idx >> P.ch_shift, c = idx & ch_mask;
            if (c >= meta.nch[b]) continue;
            const uint64_t ga = (meta.g_lo[b] & ~15ull) + 16ull * c;
            uint4 q;
            if (ga + 16 <= in_limit) q = __ldg(reinterpret_cast<const uint4*>(Q.in + ga));
            else {
                uint32_t t[4] = {0, 0, 0, 0};
                for (int i = 0; i < 16; ++i) if (ga + i < in_limit) t[i >> 2] |= (uint32_t)Q.in[ga + i] << (8 * (i & 3));
                q = make_uint4(t[0], t[1], t[2], t[3]);
            }
            *reinterpret_cast<uint4*>(S + P.raw_base[b] + 16u * c) = q;
        }
        __syncthreads();
        // ---- squeeze the beacon slots out and scale by 4 (table byte offset): R_b[q] = 4 * frame byte (q + beacon slots before
        // body symbol q).  Bytes >= 27 are reduced mod 27 first (out-of-alphabet symbols read as their low three trits, OLD:28-31)
        for (uint32_t idx = tid; idx < (9u << P.ch_shift); idx += SUP_TPB) {
            const uint32_t b = idx >> P.ch_shift, d = idx & ch_mask, L = 26u * P.ncw[P.kslot[b]];
            if (16u * d >= L) continue;
            const uint32_t q0 = 16u * d;
            uint32_t nb = 0;
            if (P.G) {
                const uint32_t t = q0 + P.G - 1u - meta.o_first[b];
                nb = __umulhi(t, P.Gm1_magic);
                if (P.G - 2u - (t - nb * (P.G - 1u)) < 15u) continue; // the count changes inside this chunk: second loop
            }
            uint4 q = lds_gather16(S + P.raw_base[b] + ((uint32_t)meta.g_lo[b] & 15u), q0 + nb);
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
        // chunks in which the number of beacon slots passed changes, byte by byte: boundary j is the first q with j + 1 slots before it
        for (uint32_t idx = tid; idx < (9u << P.sl_shift); idx += SUP_TPB) {
            const uint32_t b = idx >> P.sl_shift, j = idx & sl_mask;
            if (j >= meta.n_slow[b]) continue;
            const uint32_t qj = meta.o_first[b] + (P.G - 1u) * j;
            if (!(qj & 15u)) continue;
            const uint32_t q0 = qj & ~15u;
            const uint8_t* srcb = S + P.raw_base[b] + ((uint32_t)meta.g_lo[b] & 15u);
            uint8_t* dstb = R + P.run_base[b];
            for (uint32_t i = 0; i < 16; ++i) {
                const uint32_t q = q0 + i, v = srcb[q + j + (q >= qj ? 1u : 0u)];
                dstb[q] = (uint8_t)(4u * (v % 27u));
            }
        }
        __syncthreads();                           // R complete, the raw runs in S and this tile's meta dead
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
        const uint16_t* map = P.map + tm * (SUP_MAX_PASS * 32);
        const uint8_t* pkv = P.pass_kv + tm * SUP_MAX_PASS;
#pragma unroll 1
        for (uint32_t pass = warp; pass < P.npass[tm]; pass += SUP_WARPS) {
            const uint32_t e = __ldg(map + 32 * pass + lane), kv = __ldg(pkv + pass);
            if (e == SUP_IDLE) continue;
            const uint32_t ks = kv & 3u, v = kv >> 2, b = e & 15u, cl = e >> 4, K = P.kk[ks];
            const uint8_t* src = R + P.run_base[b] + 26u * cl;
            uint8_t* dst = S + 9u * K * cl + b;
            const uint32_t toff = P.off_tab[ks] + v * 6656u;
            const uint32_t cnz = chk[6 * ks + 2 * v], ctw = chk[6 * ks + 2 * v + 1];
            if (K == 20) dec_cw<20>(src, dst, smem32 + toff, smem + toff, cnz, ctw, sg, Q.status + 2 * f);
            else if (K == 22) dec_cw<22>(src, dst, smem32 + toff, smem + toff, cnz, ctw, sg, Q.status + 2 * f);
            else dec_cw<24>(src, dst, smem32 + toff, smem + toff, cnz, ctw, sg, Q.status + 2 * f);
        }
        __syncthreads();                           // S complete, R dead
        if (super_rows_in_smem(P)) { super_reverse_rows(S, P, T, tid); __syncthreads(); }
        // ---- phase A: 26 stream symbols -> six pixels per thread -> pixel-side bytes in R
        const uint64_t g_lo = Q.out_stride * f + (uint64_t)PIXB * P.UN * T;
        const uint32_t pad = (uint32_t)(g_lo & 15u);
        for (uint32_t ch = warp; 64u * ch < P.UN; ch += SUP_WARPS) {
#pragma unroll 1
            for (uint32_t h = 0; h < 2; ++h) {
                const uint32_t u = 64u * ch + 2u * lane + h;
                if (u >= P.UN) continue;
                const bool rev = P.tile_h26 && ((P.UN * T + u) % P.tile_h26 & 1u); // 26-wide tiles: unit = row
                if constexpr (WORDS) dec_unit_words<true>(S, 26u * u, R + pad + 27u * u, rev);
                else dec_unit_rgb<true>(S, 26u * u, R + pad + 18u * u, rev);   // pad is even (frames start on even bytes, 18 UN T is even)
            }
        }
        __syncthreads();
        {   // pixel-side bytes -> global: whole 16-byte chunks, the (at most 15 + 15) edge bytes one by one
            const uint32_t n_out = PIXB * P.UN, end = pad + n_out, c_lo = pad ? 1u : 0u, c_hi = end >> 4;
            uint8_t* gout = Q.out + (g_lo - pad);
            for (uint32_t c = c_lo + tid; c < c_hi; c += SUP_TPB) *reinterpret_cast<uint4*>(gout + 16u * c) = *reinterpret_cast<const uint4*>(R + 16u * c);
            if (tid < 16 && pad && (uint32_t)tid >= pad && (uint32_t)tid < end) gout[tid] = R[tid];
            if (tid >= 32 && tid < 48) { const uint32_t pos = 16u * c_hi + (uint32_t)(tid - 32); if (pos < end && (c_hi >= c_lo) && !(c_hi == 0 && pad)) gout[pos] = R[pos]; }
        }
        __syncthreads();
    }
}

// =============================================================================================
// host side: plan (geometry, shared-memory layout, pass maps) and launchers
// =============================================================================================
static uint32_t gcd_u32(uint32_t a, uint32_t b) { while (b) { const uint32_t t = a % b; a = b; b = t; } return a; }
static uint32_t lcm_u32(uint32_t a, uint32_t b) { return a / gcd_u32(a, b) * b; }
static uint32_t up16(uint32_t x) { return (x + 15u) & ~15u; }
static uint32_t up256(uint32_t x) { return (x + 255u) & ~255u; }

// the configs the super-tile kernels take: every k >= 20 (plane tables), at most the 2D tile widths that divide 26,
// beacon periods that leave at most one slot per 16-byte chunk and fit the 32-bit reciprocal
bool super_config_ok(const t3c_config& cfg)
{
    if (cfg.profile == T3C_PROFILE_RAW) return false;
    static const int ks[4] = {24, 22, 20, 18};
    for (int b = 0; b < 9; ++b) if (ks[cfg.uep[b] % 4] < 20) return false;
    if (use_2d(cfg) && (cfg.tile_w > 26 || 26 % cfg.tile_w != 0)) return false;
    if (use_beacon(cfg) && cfg.beacon_slot < 9 && (cfg.beacon_period < 3 || cfg.beacon_period > 255)) return false;
    return true;
}
// false when no super-tile shape fits (shared memory, pass table) or the frame holds no full super-tile
static bool make_super_plan(const t3c_config& cfg, const Geom& g, bool decode, bool words, uint64_t px_limit, SuperPlan& P, uint16_t* h_map, uint8_t* h_kv)
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
    if (g.tile_area) { P.tile_w = g.tile_w; P.tile_area = (uint32_t)g.tile_area; P.tile_h26 = g.tile_w == 26 ? (uint32_t)(g.tile_area / 26) : 0u; }
    const uint32_t pixb = words ? 27u : 18u;
    const uint32_t budget = (227u * 1024u - 2048u) / 2u - (decode ? 256u : 0u); // two CTAs per SM
    // the largest multiple of l that fits
    bool found = false;
    for (uint32_t mult = 16; mult >= 1 && !found; --mult) {
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
        P.off_aux = off; off += up16(4u * 6u * 4u);
        P.off_gf = off; off += decode ? up16((uint32_t)sizeof(GfTables)) : 0u;
        P.off_meta = off; off += up16((uint32_t)sizeof(SuperMeta));
        const uint32_t UN = 9u * M / 26u, pix = up16(pixb * UN + 48u);
        P.off_S = off; off += up16((decode && raws > 9u * M ? raws : 9u * M) + 32u);
        P.off_U = off; off += runs > pix ? runs : pix;
        if (off > budget) continue;
        P.smem_bytes = off + (decode ? 256u : 0u);
        P.M = M; P.UN = UN;
        uint32_t max_len = 0;
        for (uint32_t s = 0; s < P.nk; ++s) max_len = 26u * P.ncw[s] > max_len ? 26u * P.ncw[s] : max_len;
        const uint32_t max_exp = max_len + (P.G ? max_len / (P.G - 1u) + 2u : 0u);           // run length in the frame, beacon slots included
        const uint32_t max_ch = (max_exp + 30u) / 16u + 1u, max_sl = 3u + (P.G ? (max_exp + 15u) / (P.G - 1u) + 1u : 0u);
        P.ch_shift = 0; while ((1u << P.ch_shift) < max_ch) ++P.ch_shift;
        P.sl_shift = 0; while ((1u << P.sl_shift) < max_sl) ++P.sl_shift;
        found = true;
    }
    if (!found) return false;
    uint64_t nt = ~0ull;
    for (int b = 0; b < 9; ++b) { const uint64_t t = g.ncw[b] / P.ncw[P.kslot[b]]; nt = t < nt ? t : nt; }
    const uint64_t by_px = px_limit / (6ull * P.UN);
    nt = by_px < nt ? by_px : nt;
    if (!nt || nt > 0x7FFFFFFFull) return false;
    P.n_tiles = (uint32_t)nt;
    // pass maps: for super-tile index T = tm (mod 3), codeword cl of band b has scrambler variant (cw_base_b + n_b T + cl) mod 3;
    // a pass takes 32 codewords of one (k, variant) in (row, band) order
    for (int tm = 0; tm < 3; ++tm) {
        uint32_t pass = 0;
        for (uint32_t s = 0; s < P.nk; ++s)
            for (uint32_t v = 0; v < 3; ++v) {
                uint32_t fill = 32;
                for (uint32_t cl = 0; cl < P.ncw[s]; ++cl)
                    for (uint32_t b = 0; b < 9; ++b) {
                        if (P.kslot[b] != s || (g.cw_base[b] + (uint64_t)P.ncw[s] * tm + cl) % 3 != v) continue;
                        if (fill == 32) {
                            if (pass == SUP_MAX_PASS) return false;
                            for (int i = 0; i < 32; ++i) h_map[(tm * SUP_MAX_PASS + pass) * 32 + i] = (uint16_t)SUP_IDLE;
                            h_kv[tm * SUP_MAX_PASS + pass] = (uint8_t)(s | v << 2);
                            ++pass; fill = 0;
                        }
                        h_map[(tm * SUP_MAX_PASS + pass - 1) * 32 + fill++] = (uint16_t)(b | cl << 4);
                    }
            }
        P.npass[tm] = pass;
    }
    return true;
}
