// tma3d_probe.cu -- stand-alone check of the 3-D tensor copy the v5 decoder uses for its nine band runs (u16 elements, box {184, 9, 1}).
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma3d_probe tma3d_probe.cu && ./tma3d_probe
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__global__ void k2(const __grid_constant__ CUtensorMap tmap, uint8_t* out, int x, int nbytes)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t bar = dst + 8192;
    const int lane = threadIdx.x;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nbytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(dst), "l"(&tmap), "r"(x), "r"(0), "r"(bar) : "memory");
    }
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(0) : "memory");
    } while (!ok);
    __syncwarp();
    for (int i = lane; i < nbytes; i += 32) out[i] = smem[i];
}
__global__ void k(const __grid_constant__ CUtensorMap tmap, uint8_t* out, int x, int z, int arrivals)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t bar = dst + 9 * 368;
    const int lane = threadIdx.x;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(arrivals) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane < arrivals) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(lane == 0 ? 9u * 368u : 0u) : "memory");
        if (lane == 0)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(dst), "l"(&tmap), "r"(x), "r"(0), "r"(z), "r"(bar) : "memory");
    }
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(0) : "memory");
    } while (!ok);
    __syncwarp();
    for (int i = lane; i < 9 * 368; i += 32) out[i] = smem[i];
}

int main(int argc, char** argv)
{
    const uint64_t pitch = 26ull * 798720;     // bytes between band runs (8K, k = 20)
    const size_t bytes = 9 * pitch + 64 + 4096;
    std::vector<uint8_t> h(bytes);
    for (size_t i = 0; i < bytes; ++i) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *d, *o;
    cudaMalloc(&d, bytes); cudaMalloc(&o, 9 * 368);
    cudaMemcpy(d, h.data(), bytes, cudaMemcpyHostToDevice);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) { std::puts("no entry point"); return 2; }
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    if (argc > 1) {   // 2-D variants: elem size, box width (elements), box height, L2 promotion
        const int es = atoi(argv[1]), bw = atoi(argv[2]), bh = atoi(argv[3]), l2 = atoi(argv[4]);
        CUtensorMap t2;
        const cuuint64_t d2[2] = {pitch / es, 9}, s2[1] = {pitch};
        const cuuint32_t b2[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, e2[2] = {1, 1};
        const CUresult r2 = ((encode_fn)fn)(&t2, es == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : es == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d + 48, d2, s2, b2, e2,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        std::printf("2d es=%d box=%dx%d l2=%d encode rc=%d\n", es, bw, bh, l2, (int)r2);
        if (r2 != CUDA_SUCCESS) return 3;
        const int nb = es * bw * bh;
        k2<<<1, 32, 8192 + 64>>>(t2, o, 0, nb);
        const cudaError_t e = cudaDeviceSynchronize();
        std::vector<uint8_t> got(nb);
        cudaMemcpy(got.data(), o, nb, cudaMemcpyDeviceToHost);
        size_t bad = 0;
        for (int b = 0; b < bh; ++b)
            for (int i = 0; i < es * bw; ++i) bad += got[es * bw * b + i] != h[48 + pitch * b + i];
        std::printf("  -> %s, mismatches %zu\n", cudaGetErrorString(e), bad);
        return e == cudaSuccess ? 0 : 1;
    }
    CUtensorMap tm;
    const cuuint64_t nfr = getenv("NFR") ? atoi(getenv("NFR")) : 1;
    const cuuint64_t dims[3] = {pitch / 2, 9, nfr}, strides[2] = {pitch, (9 * pitch + 15) / 16 * 16 + (getenv("PADF") ? 4096 : 0)};
    const cuuint32_t box[3] = {184, 9, 1}, estr[3] = {1, 1, 1};
    const CUresult r = ((encode_fn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    std::printf("encode rc=%d\n", (int)r);
    if (r != CUDA_SUCCESS) return 3;
    for (int arrivals : {1, 9}) {
        for (int tile : {-1, 0, 1, 7, 61000}) {
            const int x = tile < 0 ? 0 : (int)(((52u + 338u * (unsigned)tile) & ~15u) >> 1);   // 16-byte aligned column (an unaligned one faults: illegal instruction)
            cudaMemset(o, 0, 9 * 368);
            k<<<1, 32, 9 * 368 + 64>>>(tm, o, x, 0, arrivals);
            const cudaError_t e = cudaDeviceSynchronize();
            std::vector<uint8_t> got(9 * 368);
            cudaMemcpy(got.data(), o, got.size(), cudaMemcpyDeviceToHost);
            size_t bad = 0;
            for (int b = 0; b < 9; ++b)
                for (int i = 0; i < 368; ++i) bad += got[368 * b + i] != h[pitch * b + 2 * (size_t)x + i];
            std::printf("arrivals=%d tile=%d: %s, mismatches %zu\n", arrivals, tile, cudaGetErrorString(e), bad);
            if (e != cudaSuccess) return 1;
        }
    }
    return 0;
}
