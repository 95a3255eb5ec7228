// rf_ports.cu -- does register-file read bandwidth (two banks: even / odd registers) bound the issue rate of an ALU + FMA instruction mix?
// Every test alternates one LOP3 (alu pipe) and one IMAD (fma pipe), both rt = 2 on their own pipes, so a pair could issue in 2 cycles;
// the variants differ only in how many distinct registers an instruction reads and in which banks they lie.  The register numbers
// that ptxas chose are in the SASS (cuobjdump -sass rf_ports | grep -A40 Li<N>E); 8 warps per sub-partition, 8 chains per thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 2048
#define CH 8
template <int OP>
__global__ void __launch_bounds__(1024, 1) k(uint32_t* out, uint32_t seed, long long* cyc)
{
    uint32_t x[CH], y[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { x[i] = seed * (i + 1) + threadIdx.x; y[i] = seed * (i + 77) ^ threadIdx.x; }
    uint32_t c1 = seed | 1u, c2 = seed ^ 0x9E3779B9u, c3 = seed * 3u, c4 = seed * 5u;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (OP == 0) { // one register source each
                asm volatile("xor.b32 %0, %0, 0x12345;" : "+r"(x[i]));
                asm volatile("mul.lo.u32 %0, %0, 229;" : "+r"(y[i]));
            }
            if (OP == 1) { // two register sources each (x[i], x[i+1])
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(x[(i + 1) % CH]));
                asm volatile("mad.lo.u32 %0, %0, 229, %1;" : "+r"(y[i]) : "r"(y[(i + 1) % CH]));
            }
            if (OP == 2) { // two register sources each (x[i], x[i+2])
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(x[(i + 2) % CH]));
                asm volatile("mad.lo.u32 %0, %0, 229, %1;" : "+r"(y[i]) : "r"(y[(i + 2) % CH]));
            }
            if (OP == 3) { // three register sources each, two of them shared by all instructions (reuse cache)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(c1), "r"(c2));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(c3), "r"(c4));
            }
            if (OP == 4) { // three distinct register sources each, none shared
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(x[(i + 1) % CH]), "r"(x[(i + 2) % CH]));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(y[(i + 1) % CH]), "r"(y[(i + 2) % CH]));
            }
            if (OP == 5) { // LOP3 alone, three distinct sources
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(x[(i + 1) % CH]), "r"(x[(i + 2) % CH]));
            }
            if (OP == 6) { // the trit adder: three LOP3 (nine register reads) + one IMAD r,imm,r
                uint32_t t;
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x92;" : "=r"(t) : "r"(x[i]), "r"(y[i]), "r"(c2));
                asm volatile("lop3.b32 %0, %1, %0, %2, 0xE6;" : "+r"(x[i]) : "r"(t), "r"(c1));
                asm volatile("lop3.b32 %0, %1, %0, %2, 0x24;" : "+r"(y[i]) : "r"(t), "r"(c1));
            }
            if (OP == 7) { // LOP3 r,r,imm + IMAD.HI r,imm
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(x[(i + 1) % CH]));
                asm volatile("mul.hi.u32 %0, %0, 2251799814;" : "+r"(y[i]));
            }
            if (OP == 9) { x[i] = (uint32_t)__viaddmin_s32_relu((int)x[i], (int)c1, (int)c2); asm volatile("" : "+r"(x[i])); }
            if (OP == 8) { // three pipes: LOP3 r,imm + IMAD r,imm + shared load (conflict-free)
                asm volatile("xor.b32 %0, %0, 0x12345;" : "+r"(x[i]));
                asm volatile("mul.lo.u32 %0, %0, 229;" : "+r"(y[i]));
            }
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s ^= x[i] ^ y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
static void run(const char* name, int per_iter)
{
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 4 * 1024 * 4); cudaMalloc(&cyc, 64);
    k<OP><<<1, 1024>>>(out, 12345u, cyc); k<OP><<<1, 1024>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-72s %5.2f cycles per warp instruction per sub-partition\n", name, (double)h / ((double)ITER * CH * per_iter * 8));
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    run<0>("LOP3 r,imm + IMAD r,imm (1 register read each)", 2);
    run<1>("LOP3 r,r' + IMAD r,imm,r' (2 reads, neighbours i, i+1)", 2);
    run<2>("LOP3 r,r' + IMAD r,imm,r' (2 reads, i, i+2)", 2);
    run<3>("LOP3 r,c1,c2 + IMAD r,c3,c4 (3 reads, 2 shared by all)", 2);
    run<4>("LOP3 r,r',r'' + IMAD r,r',r'' (3 distinct reads)", 2);
    run<5>("LOP3 r,r',r'' alone", 1);
    run<6>("gf3_add: 3 LOP3 (t; nz'; two')", 3);
    run<7>("LOP3 r,r' + IMAD.HI r,imm", 2);
    run<9>("VIADDMNMX.RELU r,r,r", 1);
    return 0;
}
