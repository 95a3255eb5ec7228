// mma_gf3_probe.cu -- measurement behind DESIGN.md section 7b: is the RS parity of a mini-tile cheaper as a GF(3) matrix product on the
// tensor cores than as the scalar bit-plane sums of the fused kernels?
//
// Problem instance = encode phase B of one k = 20 mini-tile: 117 codewords x 20 data symbols of GF(27) in the 9-band stream order of
// the product kernel (symbol i of codeword c at S[180 (c / 9) + c % 9 + 9 i]) -> 6 parity symbols per codeword.  The parity trits are
// (60 data trits) x P over GF(3) with P a 60 x 18 matrix (a random one here: the arithmetic is the same for the real generator).
//
//   scalar : lane = codeword, four passes; per data symbol one byte gather, two table loads (nz / two planes of that symbol's
//            contribution), three LOP3 (trit-wise add mod 3); planes -> symbols through PRMT.  This is enc_cw5 of k_fast5.cuh.
//   mma    : eight m16 row blocks; A = data trits as int8 (one 4-byte slot per symbol: t0 t1 t2 0, from a 27-entry table), B = P in
//            fragment layout (18 registers, loaded once), 3 k-steps x 3 n-tiles of mma.sync.m16n8k32.s8 per block, then every int32
//            sum reduced mod 3 and three trits recombined into a symbol.
//
// Both write the same 117 x 6 parity bytes (checked).  Each CTA = 28 warps, one per SM, every warp loops over `tiles` mini-tiles
// (its own shared-memory copy of the data, as in the product); cycles per tile per warp and per SM are printed.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/probe/mma_gf3_probe tools/probe/mma_gf3_probe.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int K = 20, R = 6, CW = 117, WARPS = 28, S_BYTES = 13 * 9 * K;   // 2340 bytes of stream symbols per mini-tile

template <int IMM>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(IMM));
    return r;
}
__device__ __forceinline__ void gf3_add(uint32_t& nz, uint32_t& two, uint32_t bnz, uint32_t btwo)   // dev.cuh
{
    const uint32_t t = lop3<0x92>(nz, two, btwo);
    const uint32_t s0 = lop3<0xE6>(t, nz, bnz);
    const uint32_t s1 = lop3<0x24>(t, two, bnz);
    nz = s0;
    two = s1;
}
__device__ __forceinline__ uint32_t planes4_to_sym(uint32_t sel)   // nibble (3 plane bits) -> b0 + 3 b1 + 9 b2, four at a time
{
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(0x04030100u), "r"(0x0D0C0A09u), "r"(sel));
    return r;
}

// ---- scalar: tabs = [2][K][27] words (plane nz, plane two), trit c of parity symbol j at bit 4 j + c
__global__ void __launch_bounds__(32 * WARPS, 1) k_scalar(const uint8_t* __restrict__ data, const uint32_t* __restrict__ tabs, uint8_t* __restrict__ out, int tiles,
                                                            long long* __restrict__ cycles)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t* tab = reinterpret_cast<uint32_t*>(smem);                   // 2 * 20 * 27 words
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t* S = smem + 4 * 2 * K * 27 + warp * (S_BYTES + 12 + 8 * 128);
    uint8_t* O = S + S_BYTES + 12;                                        // 117 x 8 bytes of parity
    for (int i = tid; i < 2 * K * 27; i += blockDim.x) tab[i] = tabs[i];
    for (int i = lane; i < S_BYTES; i += 32) S[i] = data[i] * 4;          // x4: table byte offsets, as phase A leaves them
    __syncthreads();
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
#pragma unroll 1
        for (int pass = 0; pass < 4; ++pass) {
            const int c = 32 * pass + lane;
            if (c >= CW) continue;
            const uint8_t* src = S + 180 * (c / 9) + c % 9;
            uint32_t nz = 0, two = 0, nz2 = 0, two2 = 0;
#pragma unroll
            for (int i = 0; i < K; ++i) {
                const uint32_t d4 = src[9 * i];
                const uint32_t ea = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(tab) + 108 * i + d4);
                const uint32_t eb = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(tab) + 4 * K * 27 + 108 * i + d4);
                if (i & 1) gf3_add(nz2, two2, ea, eb); else gf3_add(nz, two, ea, eb);
            }
            gf3_add(nz, two, nz2, two2);
            const uint32_t tw = two;                                      // trit = nz + two per bit
            const uint32_t lo = planes4_to_sym(nz) + planes4_to_sym(tw), hi = planes4_to_sym(nz >> 16) + planes4_to_sym(tw >> 16);
            *reinterpret_cast<uint2*>(O + 8 * c) = make_uint2(lo, hi & 0xFFFFu);
        }
        __syncwarp();
    }
    const long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x * WARPS + warp] = t1 - t0;
    if (blockIdx.x == 0 && warp == 0)
        for (int i = lane; i < CW * R; i += 32) out[i] = O[8 * (i / R) + i % R];
}

// ---- mma: bfrag = [3 k-steps][3 n-tiles][2][32 lanes] words; lut[27] = trit bytes of a symbol
__global__ void __launch_bounds__(32 * WARPS, 1) k_mma(const uint8_t* __restrict__ data, const uint32_t* __restrict__ bfrag, uint8_t* __restrict__ out, int tiles,
                                                         long long* __restrict__ cycles)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t* lut = reinterpret_cast<uint32_t*>(smem);                   // 32 words
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tig = lane & 3;
    uint8_t* S = smem + 128 + warp * (S_BYTES + 12 + 8 * 128);
    uint8_t* O = S + S_BYTES + 12;
    if (tid < 27) lut[tid] = (tid % 3) | ((tid / 3) % 3) << 8 | (tid / 9) << 16;
    for (int i = lane; i < S_BYTES; i += 32) S[i] = data[i] * 4;          // x4 as well: the table is indexed by byte offset
    uint32_t b[3][3][2];
#pragma unroll
    for (int ks = 0; ks < 3; ++ks)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt) { b[ks][nt][0] = bfrag[((ks * 3 + nt) * 2 + 0) * 32 + lane]; b[ks][nt][1] = bfrag[((ks * 3 + nt) * 2 + 1) * 32 + lane]; }
    __syncthreads();
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
#pragma unroll 1
        for (int mb = 0; mb < 8; ++mb) {
            const int r0 = 16 * mb + g, r1 = r0 + 8;
            const int q0 = r0 < CW ? r0 : CW - 1, q1 = r1 < CW ? r1 : CW - 1;   // rows past the tile: computed, not stored
            const uint8_t* s0 = S + 180 * (q0 / 9) + q0 % 9 + 9 * tig;
            const uint8_t* s1 = S + 180 * (q1 / 9) + q1 % 9 + 9 * tig;
            int acc[3][4];
#pragma unroll
            for (int nt = 0; nt < 3; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0;
#pragma unroll
            for (int ks = 0; ks < 3; ++ks) {
                // symbols 8 ks + tig (a0 / a1: rows r0 / r1) and 8 ks + 4 + tig (a2 / a3); slots 20..23 are empty
                const uint32_t a0 = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(lut) + s0[72 * ks]);
                const uint32_t a1 = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(lut) + s1[72 * ks]);
                uint32_t a2 = 0, a3 = 0;
                if (ks < 2) {
                    a2 = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(lut) + s0[72 * ks + 36]);
                    a3 = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(lut) + s1[72 * ks + 36]);
                }
#pragma unroll
                for (int nt = 0; nt < 3; ++nt)
                    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+r"(acc[nt][0]), "+r"(acc[nt][1]), "+r"(acc[nt][2]), "+r"(acc[nt][3])
                                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b[ks][nt][0]), "r"(b[ks][nt][1]));
            }
            // thread tig holds parity symbols 2 tig and 2 tig + 1 (tig 3: padding columns) of rows r0 (acc[.][0..1]) and r1 (acc[.][2..3]); n-tile = trit
            if (tig < 3) {
                uint32_t sym[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    uint32_t v = 0, m = 1;
#pragma unroll
                    for (int nt = 0; nt < 3; ++nt, m *= 3) {
                        const uint32_t x = (uint32_t)acc[nt][e];
                        v += m * (x - 3u * ((x * 171u) >> 9));             // x mod 3, x <= 120
                    }
                    sym[e] = v;
                }
                if (r0 < CW) *reinterpret_cast<uint16_t*>(O + 8 * r0 + 2 * tig) = (uint16_t)(sym[0] | sym[1] << 8);
                if (r1 < CW) *reinterpret_cast<uint16_t*>(O + 8 * r1 + 2 * tig) = (uint16_t)(sym[2] | sym[3] << 8);
            }
        }
        __syncwarp();
    }
    const long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x * WARPS + warp] = t1 - t0;
    if (blockIdx.x == 0 && warp == 0)
        for (int i = lane; i < CW * R; i += 32) out[i] = O[8 * (i / R) + i % R];
}

int main(int argc, char** argv)
{
    const int tiles = argc > 1 ? atoi(argv[1]) : 400;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    srand(12345);
    // P[3 i + c][3 j + c2]: contribution of data trit c of symbol i to parity trit c2 of symbol j
    std::vector<uint8_t> P(60 * 18), data(S_BYTES), want(CW * R);
    for (auto& v : P) v = rand() % 3;
    for (auto& v : data) v = rand() % 27;
    for (int c = 0; c < CW; ++c)
        for (int j = 0; j < R; ++j) {
            int sym = 0, mul = 1;
            for (int c2 = 0; c2 < 3; ++c2, mul *= 3) {
                int acc = 0;
                for (int i = 0; i < K; ++i) {
                    const int d = data[180 * (c / 9) + c % 9 + 9 * i], tr[3] = {d % 3, (d / 3) % 3, d / 9};
                    for (int cc = 0; cc < 3; ++cc) acc += tr[cc] * P[(3 * i + cc) * 18 + 3 * j + c2];
                }
                sym += mul * (acc % 3);
            }
            want[c * R + j] = (uint8_t)sym;
        }
    // scalar tables
    std::vector<uint32_t> tabs(2 * K * 27, 0);
    for (int i = 0; i < K; ++i)
        for (int d = 0; d < 27; ++d) {
            const int tr[3] = {d % 3, (d / 3) % 3, d / 9};
            for (int j = 0; j < R; ++j)
                for (int c2 = 0; c2 < 3; ++c2) {
                    int acc = 0;
                    for (int cc = 0; cc < 3; ++cc) acc += tr[cc] * P[(3 * i + cc) * 18 + 3 * j + c2];
                    acc %= 3;
                    if (acc) tabs[i * 27 + d] |= 1u << (4 * j + c2);
                    if (acc == 2) tabs[K * 27 + i * 27 + d] |= 1u << (4 * j + c2);
                }
        }
    // B fragments: k = 32 ks + 16 h + 4 tig + byte -> symbol slot 8 ks + 4 h + tig, trit `byte`; n = 8 nt + g -> parity symbol g, trit nt
    std::vector<uint32_t> bfrag(3 * 3 * 2 * 32, 0);
    for (int ks = 0; ks < 3; ++ks)
        for (int nt = 0; nt < 3; ++nt)
            for (int h = 0; h < 2; ++h)
                for (int lane = 0; lane < 32; ++lane) {
                    const int g = lane >> 2, tig = lane & 3, i = 8 * ks + 4 * h + tig;
                    uint32_t w = 0;
                    if (i < K && g < R)
                        for (int cc = 0; cc < 3; ++cc) w |= (uint32_t)P[(3 * i + cc) * 18 + 3 * g + nt] << (8 * cc);
                    bfrag[((ks * 3 + nt) * 2 + h) * 32 + lane] = w;
                }
    uint8_t *d_data, *d_out;
    uint32_t *d_tabs, *d_b;
    long long* d_cyc;
    cudaMalloc(&d_data, S_BYTES); cudaMalloc(&d_out, CW * R); cudaMalloc(&d_tabs, tabs.size() * 4); cudaMalloc(&d_b, bfrag.size() * 4);
    cudaMalloc(&d_cyc, sizeof(long long) * sms * WARPS);
    cudaMemcpy(d_data, data.data(), S_BYTES, cudaMemcpyHostToDevice);
    cudaMemcpy(d_tabs, tabs.data(), tabs.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_b, bfrag.data(), bfrag.size() * 4, cudaMemcpyHostToDevice);
    const int smem_a = 4 * 2 * K * 27 + WARPS * (S_BYTES + 12 + 8 * 128), smem_b = 128 + WARPS * (S_BYTES + 12 + 8 * 128);
    cudaFuncSetAttribute(k_scalar, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_a);
    cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_b);
    std::vector<uint8_t> got(CW * R);
    std::vector<long long> cyc(sms * WARPS);
    for (int variant = 0; variant < 2; ++variant) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaMemset(d_out, 0xFF, CW * R);
            cudaEventRecord(e0);
            if (variant == 0) k_scalar<<<sms, 32 * WARPS, smem_a>>>(d_data, d_tabs, d_out, tiles, d_cyc);
            else k_mma<<<sms, 32 * WARPS, smem_b>>>(d_data, d_b, d_out, tiles, d_cyc);
            cudaEventRecord(e1);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        cudaMemcpy(got.data(), d_out, CW * R, cudaMemcpyDeviceToHost);
        cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * sms * WARPS, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < CW * R; ++i) bad += got[i] != want[i];
        long long mx = 0;
        for (auto c : cyc) mx = c > mx ? c : mx;
        const double tiles_total = (double)tiles * WARPS * sms;
        printf("%-6s parity mismatches %d / %d | %d tiles per warp, %d warps per SM: %.1f us, %.0f cycles per tile per warp, %.1f cycles per tile per SM, %.2f ns per tile (whole GPU)\n",
               variant ? "mma" : "scalar", bad, CW * R, tiles, WARPS, best * 1e3, (double)mx / tiles, (double)mx / tiles / WARPS, best * 1e6 / tiles_total);
    }
    return 0;
}
