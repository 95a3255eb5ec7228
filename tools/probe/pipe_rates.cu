// pipe_rates.cu -- issue rate of the integer instructions the fused kernels are made of, per SM sub-partition (sm_100a).
// Each test runs N independent dependency chains per thread, 8 warps per sub-partition (1024 threads on one SM), and reports
// cycles per warp instruction per sub-partition.  nvcc -arch=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 2048
#define CH 8
template <int OP>
__global__ void k(uint32_t* out, uint32_t seed, long long* cyc)
{
    uint32_t x[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = seed * (i + 1) + threadIdx.x;
    uint32_t c1 = seed | 1u, c2 = seed ^ 0x9E3779B9u;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c1), "r"(c2));
            if (OP == 1) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(c1));
            if (OP == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(c1), "r"(c2));
            if (OP == 3) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c1), "r"(c2));
            if (OP == 4) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c1), "r"(c2));
            if (OP == 5) asm volatile("dp2a.lo.s32.u32 %0, %1, %0, %2;" : "+r"(x[i]) : "r"(c1), "r"(c2));
            if (OP == 6) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c1), "r"(c2));
            if (OP == 7) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(c1));
            if (OP == 8) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c1), "r"(c2));
            if (OP == 9) asm volatile("min.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(c1));
            if (OP == 10) asm volatile("mad.lo.u32 %0, %0, 229, %1;" : "+r"(x[i]) : "r"(c2));
            if (OP == 11) asm volatile("shl.b32 %0, %0, 2;" : "+r"(x[i]));
            if (OP == 12) asm volatile("mul.hi.u32 %0, %0, 2251799814;" : "+r"(x[i]));
            if (OP == 13) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(c1), "r"(c2)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[(i + 4) % CH]) : "r"(c1), "r"(c2)); }
            if (OP == 14) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c1), "r"(c2));
            if (OP == 15) { int d; asm volatile("vadd.s32.s32.s32.min %0, %1, %2, %3;" : "=r"(d) : "r"(x[i]), "r"(c1), "r"(c2)); x[i] = d; }
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// shared-memory look-ups: 32-bit loads from a 27-entry row (conflict-free), byte loads at stride 9, 64-bit loads
template <int OP>
__global__ void ks(uint32_t* out, uint32_t seed, long long* cyc)
{
    __shared__ __align__(16) uint32_t tab[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) tab[i] = (i * 2654435761u) % 27u * 4u;
    __syncthreads();
    uint32_t x[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = ((seed * (i + 1) + threadIdx.x) % 27u) * 4u;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (OP == 0) x[i] = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(tab) + x[i] + 128 * i);
            if (OP == 1) x[i] = (uint32_t)(reinterpret_cast<const uint8_t*>(tab)[x[i] * 9 + i] & 0x7Cu);
            if (OP == 2) { const uint2 v = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(tab) + 2 * x[i] + 256 * i); x[i] = (v.x ^ v.y) & 0x7Cu; }
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <class F>
static void run(const char* name, F launch, int per_iter, int warps_per_smsp)
{
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 4 * 1024 * 4); cudaMalloc(&cyc, 64);
    launch(out, cyc); launch(out, cyc);
    cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double inst = (double)ITER * CH * per_iter * warps_per_smsp;
    printf("%-28s %6.2f cycles per warp instruction per sub-partition (%d warps)\n", name, (double)h / inst, warps_per_smsp);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    const char* names[16] = {"IMAD r,r,r", "IMAD.HI (mul.hi) r,r", "LOP3", "PRMT r,r,r", "SHF.R.W", "IDP.2A", "IDP.4A", "IADD", "IMAD.HI (mad.hi) r,r,r", "VIMNMX (min)",
                             "IMAD r,imm,r", "SHL imm", "IMAD.HI r,imm", "LOP3 + IMAD pair", "FFMA r,r,r", "vadd.min (PTX video op: emulated)"};
#define RUN(OP, PER) run(names[OP], [](uint32_t* o, long long* c) { k<OP><<<1, 1024>>>(o, 12345u, c); }, PER, 8);
    RUN(0, 1) RUN(10, 1) RUN(1, 1) RUN(12, 1) RUN(8, 1) RUN(2, 1) RUN(3, 1) RUN(4, 1) RUN(11, 1) RUN(5, 1) RUN(6, 1) RUN(7, 1) RUN(9, 1) RUN(15, 1) RUN(14, 1) RUN(13, 2)
    run("LDS.32 27-entry rows", [](uint32_t* o, long long* c) { ks<0><<<1, 1024>>>(o, 12345u, c); }, 1, 8);
    run("LDS.U8 stride 9 (+LOP)", [](uint32_t* o, long long* c) { ks<1><<<1, 1024>>>(o, 12345u, c); }, 1, 8);
    run("LDS.64 27-entry rows (+2 LOP)", [](uint32_t* o, long long* c) { ks<2><<<1, 1024>>>(o, 12345u, c); }, 1, 8);
    return 0;
}
