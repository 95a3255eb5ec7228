// k_fast.cu -- tiled fused kernels (placeholder until the general path is parity-green on the GPU)
#include "dev.cuh"
#include "launch.h"
namespace t3c {
bool fast_path_ok(const t3c_config&) { return false; }
int launch_encode_rgb_fast(const DevTables&, const t3c_config&, const Geom&, const uint8_t*, size_t, size_t, uint8_t*, size_t, cudaStream_t) { return 0; }
int launch_decode_rgb_fast(const DevTables&, const t3c_config&, const Geom&, const uint8_t*, size_t, size_t, size_t, size_t, uint8_t*, uint32_t*, cudaStream_t) { return 0; }
}
