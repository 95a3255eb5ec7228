// api.cu -- the C ABI of libt3c.so (include/t3c.h): context, device buffers, host<->device staging
// and dispatch to the kernels.  No CPU implementation of any codec stage lives here: without a CUDA
// device every entry point fails (T3C_ERR_NODEVICE), loudly.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <string>
#include <vector>

#include "launch.h"
#include "t3c_internal.h"

using namespace t3c;
static_assert(sizeof(t3c_config) == 44 && sizeof(t3c_pixel) == 6, "ABI struct layout");

struct t3c_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr; // copy streams of the chunked host pipelines
    cudaEvent_t ev[32] = {};
    int ev_next = 0;
    DevTables tabs{};
    HeaderCache hdr_cache{};
    SuperCache sup_cache{};
    FastImageCache img_cache{};
    void* d_tables = nullptr;
    uint32_t* d_crc = nullptr; // CRC-32 tables of the .t3v record kernels
    HostTables* host = nullptr; // host copy of the constant tables (decoder screen constants are derived per call)
    // grow-only device scratch
    struct Buf { void* p = nullptr; size_t cap = 0; cudaStream_t last = nullptr; bool used = false; };  // last: the stream of the latest user
    Buf buf[6];
    // small results: device mailbox + pinned host mirror (copied explicitly, MAIL_DOWN)
    struct Mail { uint32_t status[64]; t3c_config cfg; int ok; uint8_t hdr27[27]; uint8_t coded52[52]; };
    Mail* h_mail = nullptr;
    Mail* d_mail = nullptr;
    uint64_t launches = 0;
    std::string err;
    // pageable host buffers of the host-buffer calls (t3c_set_host_registration): ranges seen before are page-locked in place
    struct HostReg { const void* p = nullptr; size_t n = 0; bool registered = false; uint64_t last_use = 0; };
    HostReg regs[16];
    int host_reg_mode = 0;
    uint64_t reg_clock = 0;
};

namespace {

enum { B_IN = 0, B_OUT = 1, B_TMP = 2, B_TMP2 = 3, B_AUX = 4, B_AUX2 = 5 };

t3c_status fail(t3c_ctx* c, t3c_status s, const char* what, cudaError_t e = cudaSuccess)
{
    if (c) {
        c->err = what;
        if (e != cudaSuccess) { c->err += ": "; c->err += cudaGetErrorString(e); }
    }
    return s;
}
#define CU(call)                                                                   \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess) return fail(ctx, T3C_ERR_CUDA, #call, e_);          \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int d) { cudaGetDevice(&prev); if (prev != d) cudaSetDevice(d); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

t3c_status chain(t3c_ctx* ctx, cudaStream_t from, cudaStream_t to);
// Scratch is per context, not per stream: a caller that moves to another stream is ordered after the work the previous stream
// already holds (an event recorded on that stream now covers every earlier use of the buffer).  Growth synchronises the device.
t3c_status reserve_on(t3c_ctx* ctx, int slot, size_t bytes, void** out, cudaStream_t s)
{
    t3c_ctx::Buf& b = ctx->buf[slot];
    if (bytes + 64 > b.cap) {
        CU(cudaDeviceSynchronize());
        if (b.p) CU(cudaFree(b.p));
        b.p = nullptr; b.cap = 0; b.used = false;
        size_t want = bytes + 64 + bytes / 8;
        CU(cudaMalloc(&b.p, want));
        b.cap = want;
    }
    if (b.used && b.last != s) { t3c_status r = chain(ctx, b.last, s); if (r != T3C_OK) return r; }
    b.last = s; b.used = true;
    *out = b.p;
    return T3C_OK;
}
t3c_status reserve(t3c_ctx* ctx, int slot, size_t bytes, void** out) { return reserve_on(ctx, slot, bytes, out, ctx->stream); }
template <class T>
t3c_status reserve_t(t3c_ctx* ctx, int slot, size_t bytes, T** out)
{
    void* p = nullptr;
    t3c_status s = reserve(ctx, slot, bytes, &p);
    *out = static_cast<T*>(p);
    return s;
}
template <class T>
t3c_status reserve_t(t3c_ctx* ctx, int slot, size_t bytes, T** out, cudaStream_t st)
{
    void* p = nullptr;
    t3c_status s = reserve_on(ctx, slot, bytes, &p, st);
    *out = static_cast<T*>(p);
    return s;
}
// Pageable host memory makes every cudaMemcpyAsync a staged, blocking copy: the chunked pipelines degrade to ~1/5 of their pinned
// throughput (8K encode + decode: 43.8 ms against 7.9 ms per frame, tools/e2e_pageable.py).  Page-locking a range costs ~170 us per
// MB, so it only pays for buffers that come back: a large pageable range is noted the first time it is seen and page-locked
// (cudaHostRegister) when the same range is passed again; the registrations are kept (least recently used out) and dropped with
// the context.  The caller must not free such a buffer while another thread is inside a call that uses it -- nothing new -- and a
// range that was freed and re-allocated elsewhere simply stops matching.  Off unless t3c_set_host_registration(ctx, 1).
constexpr size_t kHostRegMinBytes = 8u << 20;
void host_buffer(t3c_ctx* ctx, const void* p, size_t bytes)
{
    if (!ctx->host_reg_mode || !p || bytes < kHostRegMinBytes) return;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return; }
    if (at.type != cudaMemoryTypeUnregistered) return;                       // already pinned (or device / managed memory)
    ++ctx->reg_clock;
    t3c_ctx::HostReg* slot = nullptr;
    for (auto& r : ctx->regs) if (r.p == p && r.n == bytes) { slot = &r; break; }
    if (!slot) {                                                              // first sight: remember, evicting the least recently used entry
        for (auto& r : ctx->regs) if (!slot || r.last_use < slot->last_use) slot = &r;
        if (slot->registered) { cudaHostUnregister(const_cast<void*>(slot->p)); cudaGetLastError(); }
        *slot = t3c_ctx::HostReg{p, bytes, false, ctx->reg_clock};
        return;
    }
    slot->last_use = ctx->reg_clock;
    if (slot->registered) return;                                             // (the attribute query would have said so; a stale entry)
    cudaError_t e = cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) {                          // overlaps a range registered earlier at another size: drop those and retry
        cudaGetLastError();
        for (auto& r : ctx->regs)
            if (r.registered && (const char*)r.p < (const char*)p + bytes && (const char*)p < (const char*)r.p + r.n) { cudaHostUnregister(const_cast<void*>(r.p)); r = t3c_ctx::HostReg{}; }
        e = cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault);
    }
    if (e == cudaSuccess) slot->registered = true; else cudaGetLastError();   // could not pin: the call proceeds with pageable copies
}
t3c_status check_launch(t3c_ctx* ctx, int n)
{
    ctx->launches += (uint64_t)n;
    CU(cudaGetLastError());
    return T3C_OK;
}
#define TRY(expr) do { t3c_status s_ = (expr); if (s_ != T3C_OK) return s_; } while (0)

// everything enqueued on `to` after this call runs after everything enqueued on `from` so far
t3c_status chain(t3c_ctx* ctx, cudaStream_t from, cudaStream_t to)
{
    cudaEvent_t e = ctx->ev[ctx->ev_next];
    ctx->ev_next = (ctx->ev_next + 1) % 32;
    CU(cudaEventRecord(e, from));
    CU(cudaStreamWaitEvent(to, e, 0));
    return T3C_OK;
}
static uint32_t kPipeChunks = 8;          // chunks per frame in the host-buffer pipelines (T3C_PIPE_CHUNKS overrides)
constexpr uint32_t kPipeMinTiles = 512;   // below this a frame is copied and coded in one piece
// Tiles before chunk c of a frame's host pipeline.  Chunk weights 1, 2, 4, then 5s: the encoder's first chunks are small (its device-to-host
// leg, the long one, starts after a short fill), the decoder's last ones (a short drain after its host-to-device leg).  >= 1 tile per chunk
// for n_full >= kPipeMinTiles and up to 64 chunks.
static uint32_t pipe_edge(uint32_t n_full, uint32_t c, bool small_first)
{
    uint64_t tot = 0, before = 0;
    for (uint32_t i = 0; i < kPipeChunks; ++i) {
        const uint32_t j = small_first ? i : kPipeChunks - 1 - i, w = j < 3 ? (1u << j) : 5u;
        tot += w;
        if (i < c) before += w;
    }
    return (uint32_t)((uint64_t)n_full * before / tot);
}

void ref_dec_geom(const t3c_config& h, size_t n_words, RefDecGeom& g)
{
    static const int ks[4] = {24, 22, 20, 18};
    std::memset(&g, 0, sizeof g);
    g.n_body_words = n_words - 6;
    const bool skip = h.beacon_enabled && h.beacon_period > 0; // OLD:952
    g.period = skip ? h.beacon_period : 0;
    g.slot = skip ? h.beacon_slot : -1;
    uint64_t use = 0;
    for (int b = 0; b < 9; ++b) {
        g.k[b] = ks[h.uep[b] % 4];
        uint64_t L = g.n_body_words;
        if (skip && b == h.beacon_slot) L -= (g.n_body_words + h.beacon_period - 1) / h.beacon_period;
        g.ncw[b] = L / 26;
        g.use_base[b] = use;
        use += g.ncw[b] * (uint64_t)g.k[b];
    }
    g.n_use = use;
    if (h.profile == 4 && h.tile_w && h.tile_h) { g.tile_w = h.tile_w; g.tile_area = (uint64_t)h.tile_w * h.tile_h; } // OLD:1018
    scrambler_states(h.seed_a, h.seed_b, h.seed_s0, g.st);
}

// recovered prefix of the regrouped stream for the consistent decoder (A.8): first position of sy
// (pre-interleave order) that is not covered by a decoded codeword
uint64_t known_prefix(const t3c_config& c, const Geom& g)
{
    uint64_t pfx = g.n_s;
    for (int b = 0; b < 9; ++b) {
        const uint64_t first_unknown = 9 * ((uint64_t)g.k[b] * g.ncw[b]) + b;
        if (first_unknown < pfx) pfx = first_unknown;
    }
    if (use_2d(c) && pfx < g.n_s) {
        const uint64_t A = g.tile_area, base = (pfx / A) * A, off = pfx - base, row = off / c.tile_w;
        if (row & 1) pfx = base + row * c.tile_w; // a reversed row loses its low end first
    }
    return pfx;
}

} // namespace

extern "C" {

int t3c_version(void) { return 100; }

void t3c_config_default(t3c_config* c)
{
    std::memset(c, 0, sizeof *c);
    c->profile = 1;
    for (int i = 0; i < 9; ++i) c->uep[i] = 1;
    c->seed_a = c->seed_b = c->seed_s0 = 1;
    c->superframe_words = 8192;
    c->subword = 27;
    c->centered = 1;
}

t3c_status t3c_create(int device, t3c_ctx** out)
{
    if (!out) return T3C_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return T3C_ERR_NODEVICE;
    t3c_ctx* ctx = new t3c_ctx();
    ctx->device = device;
    if (const char* e = getenv("T3C_PIPE_CHUNKS")) { const int v = atoi(e); if (v >= 1 && v <= 64) kPipeChunks = (uint32_t)v; }
    DeviceGuard guard(device);
    HostTables* ht = new HostTables();
    build_tables(*ht);
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking);
    for (auto& ev : ctx->ev) if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_tables, sizeof(HostTables));
    if (e == cudaSuccess) e = cudaMemcpy(ctx->d_tables, ht, sizeof(HostTables), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&ctx->h_mail, sizeof(t3c_ctx::Mail));
    if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->d_mail, sizeof(t3c_ctx::Mail));
    if (e == cudaSuccess) e = cudaMemset(ctx->d_mail, 0, sizeof(t3c_ctx::Mail));
    ctx->host = ht;
    if (e != cudaSuccess) { t3c_destroy(ctx); return T3C_ERR_CUDA; }
    ctx->tabs.gf = &static_cast<HostTables*>(ctx->d_tables)->gf;
    ctx->tabs.rs = &static_cast<HostTables*>(ctx->d_tables)->rs;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    ctx->tabs.sm_count = prop.multiProcessorCount;
    ctx->tabs.hdr = &ctx->hdr_cache;
    if (cudaMalloc((void**)&ctx->hdr_cache.base, 128 * HeaderCache::N) != cudaSuccess) { t3c_destroy(ctx); return T3C_ERR_CUDA; }
    for (int i = 0; i < HeaderCache::N; ++i) { ctx->hdr_cache.e[i].d52 = ctx->hdr_cache.base + 128 * i; ctx->hdr_cache.e[i].d27 = ctx->hdr_cache.base + 128 * i + 64; }
    ctx->tabs.sup = &ctx->sup_cache;
    ctx->tabs.img = &ctx->img_cache;
    if (cudaMalloc((void**)&ctx->img_cache.base, (size_t)FastImageCache::N * FastImageCache::BYTES) != cudaSuccess) { t3c_destroy(ctx); return T3C_ERR_CUDA; }
    {
        std::vector<uint32_t> h(crc_table_words());
        build_crc_tables(h.data());
        if (cudaMalloc((void**)&ctx->d_crc, 4 * h.size()) != cudaSuccess || cudaMemcpy(ctx->d_crc, h.data(), 4 * h.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
            t3c_destroy(ctx);
            return T3C_ERR_CUDA;
        }
        ctx->tabs.crc = ctx->d_crc;
    }
    for (auto& sl : ctx->sup_cache.slot) {
        if (cudaMalloc((void**)&sl.d_map, 3 * 64 * 32 * sizeof(uint16_t) + 256) != cudaSuccess) { t3c_destroy(ctx); return T3C_ERR_CUDA; }
        sl.d_kv = reinterpret_cast<uint8_t*>(sl.d_map + 3 * 64 * 32);
    }
    // the table uploads above come from pageable memory on the legacy stream: make sure they have landed before any kernel on the
    // context's non-blocking streams can read them
    if (cudaDeviceSynchronize() != cudaSuccess) { t3c_destroy(ctx); return T3C_ERR_CUDA; }
    *out = ctx;
    return T3C_OK;
}

void t3c_destroy(t3c_ctx* ctx)
{
    if (!ctx) return;
    DeviceGuard guard(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto& r : ctx->regs) if (r.registered) cudaHostUnregister(const_cast<void*>(r.p));
    for (auto& b : ctx->buf) if (b.p) cudaFree(b.p);
    if (ctx->d_tables) cudaFree(ctx->d_tables);
    if (ctx->d_crc) cudaFree(ctx->d_crc);
    if (ctx->hdr_cache.base) cudaFree(ctx->hdr_cache.base);
    if (ctx->img_cache.base) cudaFree(ctx->img_cache.base);
    for (auto& sl : ctx->sup_cache.slot) if (sl.d_map) cudaFree(sl.d_map);
    if (ctx->h_mail) cudaFreeHost(ctx->h_mail);
    if (ctx->d_mail) cudaFree(ctx->d_mail);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    delete ctx->host;
    delete ctx;
}

t3c_status t3c_set_host_registration(t3c_ctx* ctx, int mode)
{
    if (!ctx) return T3C_ERR_ARG;
    ctx->host_reg_mode = mode ? 1 : 0;
    return T3C_OK;
}
const char* t3c_last_error(const t3c_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context (no CUDA device?)"; }
void* t3c_stream(t3c_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t t3c_kernel_launches(const t3c_ctx* ctx) { return ctx ? ctx->launches : 0; }
t3c_status t3c_sync(t3c_ctx* ctx)
{
    if (!ctx) return T3C_ERR_ARG;
    DeviceGuard guard(ctx->device);
    CU(cudaStreamSynchronize(ctx->stream));
    return T3C_OK;
}
size_t t3c_profile_words(const t3c_config* cfg, size_t n_raw_words) { return cfg ? profile_words(*cfg, n_raw_words) : 0; }
int t3c_fast_path_available(const t3c_config* cfg) { return cfg && fast_path_ok(*cfg) ? 1 : 0; }
int t3c_super_path_available(const t3c_config* cfg) { return cfg && super_path_ok(*cfg) ? 1 : 0; }
int t3c_debug_counters(uint32_t* out32) { return out32 ? super_debug_counters(out32) : 0; }
int t3c_super_plan_describe(const t3c_config* cfg, size_t n_raw_words, int decode, int words, uint32_t* out16, uint16_t* map, uint8_t* pass_kv)
{
    return cfg && out16 ? super_plan_describe(*cfg, n_raw_words, decode, words, out16, map, pass_kv) : 0;
}

// =============================================================================================
// device-pointer API
// =============================================================================================
t3c_status t3c_rgb_to_quant_dev(t3c_ctx* ctx, const uint8_t* d_rgb, size_t n_px, t3c_pixel* d_out, void* st)
{
    if (!ctx || (n_px && (!d_rgb || !d_out))) return fail(ctx, T3C_ERR_ARG, "rgb_to_quant: null");
    DeviceGuard guard(ctx->device);
    return check_launch(ctx, launch_rgb_to_quant(d_rgb, n_px, d_out, (cudaStream_t)st));
}
t3c_status t3c_quant_to_rgb_dev(t3c_ctx* ctx, const t3c_pixel* d_px, size_t n_px, uint8_t* d_rgb, void* st)
{
    if (!ctx || (n_px && (!d_rgb || !d_px))) return fail(ctx, T3C_ERR_ARG, "quant_to_rgb: null");
    DeviceGuard guard(ctx->device);
    return check_launch(ctx, launch_quant_to_rgb(d_px, n_px, d_rgb, (cudaStream_t)st));
}
t3c_status t3c_pack_pixels_dev(t3c_ctx* ctx, const t3c_pixel* d_px, size_t n_px, uint8_t* d_words, void* st)
{
    if (!ctx || (n_px && (!d_px || !d_words))) return fail(ctx, T3C_ERR_ARG, "pack_pixels: null");
    DeviceGuard guard(ctx->device);
    return check_launch(ctx, launch_pack_pixels(d_px, n_px, d_words, (cudaStream_t)st));
}
t3c_status t3c_unpack_pixels_dev(t3c_ctx* ctx, const uint8_t* d_words, size_t n_words, t3c_pixel* d_px, void* st)
{
    if (!ctx || (n_words && (!d_px || !d_words))) return fail(ctx, T3C_ERR_ARG, "unpack_pixels: null");
    DeviceGuard guard(ctx->device);
    return check_launch(ctx, launch_unpack_pixels(d_words, n_words, d_px, (cudaStream_t)st));
}
t3c_status t3c_rs_encode_blocks_dev(t3c_ctx* ctx, int k, int arith, const uint8_t* d_data, size_t n, uint8_t* d_out, void* st)
{
    if (!ctx || !k_valid(k) || (n && (!d_data || !d_out))) return fail(ctx, T3C_ERR_ARG, "rs_encode_blocks: bad k or null");
    DeviceGuard guard(ctx->device);
    return check_launch(ctx, launch_rs_encode_blocks(ctx->tabs, k, arith, d_data, n, d_out, (cudaStream_t)st));
}
t3c_status t3c_rs_decode_blocks_dev(t3c_ctx* ctx, int k, int arith, uint8_t* d_inout, size_t n, uint8_t* d_out_k, uint8_t* d_ok, void* st)
{
    if (!ctx || !k_valid(k) || (n && (!d_inout || !d_out_k || !d_ok))) return fail(ctx, T3C_ERR_ARG, "rs_decode_blocks: bad k or null");
    DeviceGuard guard(ctx->device);
    return check_launch(ctx, launch_rs_decode_blocks(ctx->tabs, k, arith, d_inout, n, d_out_k, d_ok, (cudaStream_t)st));
}

t3c_status t3c_encode_profile_dev(t3c_ctx* ctx, const t3c_config* cfg, int arith, const uint8_t* d_raw, size_t n_words,
                                  uint8_t* d_out, size_t cap_words, void* st)
{
    if (!ctx || !cfg || !d_out || (n_words && !d_raw)) return fail(ctx, T3C_ERR_ARG, "encode_profile: null");
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)st;
    if (cfg->profile == T3C_PROFILE_RAW) { // OLD:1046-1050
        if (cap_words < n_words) return fail(ctx, T3C_ERR_CAPACITY, "encode_profile: capacity");
        CU(cudaMemcpyAsync(d_out, d_raw, 9 * n_words, cudaMemcpyDeviceToDevice, s));
        return T3C_OK;
    }
    Geom g;
    make_geom(*cfg, n_words, arith, g);
    if (cap_words < g.n_out) return fail(ctx, T3C_ERR_CAPACITY, "encode_profile: capacity");
    uint32_t n_full = 0;
    int n = 0;
    if (fast_path_ok(*cfg)) { // uniform k, 1D, no beacon: the tiled kernels code the full mini-tiles straight from the raw words
        n = launch_encode_words_fast(ctx->tabs, g, d_raw, d_out, s, &n_full);
        if (n < 0) { n = 0; n_full = 0; }
    }
    else if (super_path_ok(*cfg)) { // per-band k / 2D / beacon: the super-tile kernels code the full super-tiles, the general kernel the rest
        SuperTail tail;
        n = launch_encode_super(ctx->tabs, *cfg, g, d_raw, 9 * n_words, true, 2 * n_words, 1, d_out, g.n_out, s, &tail);
        if (n > 0) {
            n += launch_encode_general_from(ctx->tabs, *cfg, g, d_raw, d_out, s, tail.cs, false);
            n += launch_frame_misc_sparse(ctx->tabs, *cfg, g, d_out, 1, 0, s, tail.n_tiles, tail.ncw_tile);
        } else n = launch_encode_general_from(ctx->tabs, *cfg, g, d_raw, d_out, s, tail.cs);
        return check_launch(ctx, n);
    }
    n += launch_encode_general(ctx->tabs, *cfg, g, d_raw, d_out, s, 13ull * n_full); // the ragged rest (or everything), header, padding
    return check_launch(ctx, n);
}

t3c_status t3c_decode_profile_fixed_dev(t3c_ctx* ctx, const t3c_config* cfg, size_t n_raw_words, const uint8_t* d_in, size_t n_words,
                                        uint8_t* d_out, size_t cap_words, uint32_t* d_status, void* st)
{
    if (!ctx || !cfg || !d_in || !d_out || !d_status) return fail(ctx, T3C_ERR_ARG, "decode_profile_fixed: null");
    if (cfg->profile == T3C_PROFILE_RAW || !n_raw_words) return fail(ctx, T3C_ERR_UNSUPPORTED, "decode_profile_fixed_dev: needs n_raw_words and a coded profile");
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)st;
    Geom g;
    make_geom(*cfg, n_raw_words, 1, g);
    if (g.n_out != n_words) return fail(ctx, T3C_ERR_ARG, "decode_profile_fixed: n_words does not match the config");
    uint64_t nw = 3 * known_prefix(*cfg, g) / 26;
    if (nw > n_raw_words) nw = n_raw_words;
    if (cap_words < nw) return fail(ctx, T3C_ERR_CAPACITY, "decode_profile_fixed: capacity");
    uint8_t* sy = nullptr;
    const uint64_t pitch = (g.n_s + 8) / 9; // band-major scratch: band b's decoded symbols at b*pitch
    TRY(reserve_t(ctx, B_TMP, 9 * pitch + 16, &sy, s));
    int n = launch_init_status(d_status, 1, s);
    uint32_t n_full = 0;
    if (fast_path_ok(*cfg)) { // tiled kernels: profile words -> raw words for the full mini-tiles
        const int k = launch_decode_words_fast(ctx->tabs, g, d_in, d_out, (size_t)nw, d_status, s, &n_full);
        if (k < 0) n_full = 0; else n += k;
    }
    else if (super_path_ok(*cfg)) {
        SuperTail tail;
        n += launch_decode_super(ctx->tabs, *cfg, g, d_in, g.n_out, 1, d_out, 9 * (size_t)nw, true, 2 * (size_t)nw, d_status, s, &tail);
        for (int b = 0; b < 9; ++b) if (pitch > tail.m_start) CU(cudaMemsetAsync(sy + b * pitch + tail.m_start, 0, pitch - tail.m_start, s));
        n += launch_decode_fixed_general_from(ctx->tabs, g, d_in, sy, pitch, d_status, s, tail.cs);
        n += launch_regroup_words(sy, g.n_s, g.tile_area, g.tile_w, d_out, (size_t)nw, s, (size_t)(3 * tail.unit_start), pitch);
        return check_launch(ctx, n);
    }
    const uint64_t m_done = 13ull * n_full * (uint64_t)g.uniform_k; // symbols per band already turned into words by the tiled kernels
    for (int b = 0; b < 9; ++b) if (pitch > m_done) CU(cudaMemsetAsync(sy + b * pitch + m_done, 0, pitch - m_done, s));
    n += launch_decode_fixed_general(ctx->tabs, g, d_in, sy, pitch, d_status, s, 13ull * n_full);
    n += launch_regroup_words(sy, g.n_s, g.tile_area, g.tile_w, d_out, (size_t)nw, s, (size_t)n_full * (27 * (size_t)g.uniform_k / 2), pitch);
    return check_launch(ctx, n);
}

t3c_status t3c_encode_frames_rgb8_dev(t3c_ctx* ctx, const t3c_config* cfg, int arith, const uint8_t* d_rgb, size_t n_px, size_t n_frames,
                                      uint8_t* d_out, size_t stride_words, void* st)
{
    if (!ctx || !cfg || !d_rgb || !d_out) return fail(ctx, T3C_ERR_ARG, "encode_frames: null");
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)st;
    const size_t n_words = (n_px + 1) / 2;
    const size_t n_out = profile_words(*cfg, n_words);
    if (stride_words < n_out) return fail(ctx, T3C_ERR_CAPACITY, "encode_frames: stride < words per frame");
    if (cfg->profile != T3C_PROFILE_RAW && fast_path_ok(*cfg)) {
        Geom g;
        make_geom(*cfg, n_words, arith, g);
        const int n = launch_encode_rgb_fast(ctx->tabs, *cfg, g, d_rgb, n_px, n_frames, d_out, stride_words, s);
        if (n >= 0) return check_launch(ctx, n);
        // unaligned buffers: fall through to the general kernels
    }
    if (cfg->profile != T3C_PROFILE_RAW && super_path_ok(*cfg)) { // per-band k / 2D / beacon: super-tile kernels + general kernels on the ragged rest
        Geom g;
        make_geom(*cfg, n_words, arith, g);
        SuperTail tail;
        const int ns = launch_encode_super(ctx->tabs, *cfg, g, d_rgb, 3 * n_px, false, n_px, n_frames, d_out, stride_words, s, &tail);
        if (ns > 0) {
            TRY(check_launch(ctx, ns));
            const size_t px0 = 6 * (size_t)tail.unit_start, n_tail = n_px - px0, w_tail = (n_tail + 1) / 2; // px0 is even: whole words
            t3c_pixel* q = nullptr;
            uint8_t* raw = nullptr;
            TRY(reserve_t(ctx, B_AUX, 6 * n_tail + 64, &q, s));
            TRY(reserve_t(ctx, B_AUX2, 9 * w_tail + 64, &raw, s));
            for (size_t f = 0; f < n_frames; ++f) {
                int n = launch_rgb_to_quant(d_rgb + 3 * n_px * f + 3 * px0, n_tail, q, s);
                n += launch_pack_pixels(q, n_tail, raw, s);
                // the general encoder indexes raw words from the start of the frame: hand it the base the tail words would have there
                const uint8_t* raw0 = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(raw) - 9 * (px0 / 2));
                n += launch_encode_general_from(ctx->tabs, *cfg, g, raw0, d_out + 9 * stride_words * f, s, tail.cs, false);
                TRY(check_launch(ctx, n));
            }
            // header, zero padding and the beacon slots outside the super-tile runs, all frames at once
            TRY(check_launch(ctx, launch_frame_misc_sparse(ctx->tabs, *cfg, g, d_out, n_frames, 9 * stride_words, s, tail.n_tiles, tail.ncw_tile)));
            return T3C_OK;
        }
    }
    // general path: per frame K1 (bridge, pack) into scratch, then the general profile encoder
    t3c_pixel* q = nullptr;
    uint8_t* raw = nullptr;
    TRY(reserve_t(ctx, B_AUX, 6 * n_px, &q, s));
    TRY(reserve_t(ctx, B_AUX2, 9 * n_words, &raw, s));
    for (size_t f = 0; f < n_frames; ++f) {
        int n = launch_rgb_to_quant(d_rgb + 3 * n_px * f, n_px, q, s);
        n += launch_pack_pixels(q, n_px, raw, s);
        TRY(check_launch(ctx, n));
        TRY(t3c_encode_profile_dev(ctx, cfg, arith, raw, n_words, d_out + 9 * stride_words * f, stride_words, st));
    }
    return T3C_OK;
}

t3c_status t3c_decode_frames_rgb8_dev(t3c_ctx* ctx, const t3c_config* cfg, const uint8_t* d_in, size_t words_per_frame, size_t stride_words,
                                      size_t n_frames, size_t n_px, uint8_t* d_rgb, uint32_t* d_status, void* st)
{
    if (!ctx || !cfg || !d_in || !d_rgb || !d_status) return fail(ctx, T3C_ERR_ARG, "decode_frames: null");
    if (cfg->profile == T3C_PROFILE_RAW) return fail(ctx, T3C_ERR_UNSUPPORTED, "decode_frames: RAW profile carries raw words, use unpack_pixels");
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)st;
    const size_t n_words = (n_px + 1) / 2;
    Geom g;
    make_geom(*cfg, n_words, 1, g);
    if (g.n_out != words_per_frame) return fail(ctx, T3C_ERR_ARG, "decode_frames: words_per_frame does not match config and n_px");
    uint64_t nw = 3 * known_prefix(*cfg, g) / 26;
    if (nw > n_words) nw = n_words;
    size_t px_out = 2 * (size_t)nw < n_px ? 2 * (size_t)nw : n_px;
    TRY(check_launch(ctx, launch_init_status(d_status, n_frames, s)));
    if (fast_path_ok(*cfg)) {
        uint32_t chk_nz[7], chk_two[7];
        fast_check_constants(*ctx->host, g, chk_nz, chk_two);
        const int n = launch_decode_rgb_fast(ctx->tabs, *cfg, g, d_in, stride_words, n_frames, n_px, px_out, d_rgb, d_status, s, chk_nz, chk_two);
        if (n >= 0) return check_launch(ctx, n);
    }
    if (super_path_ok(*cfg)) {
        SuperTail tail;
        const int ns = launch_decode_super(ctx->tabs, *cfg, g, d_in, stride_words, n_frames, d_rgb, 3 * n_px, false, px_out, d_status, s, &tail);
        if (ns > 0) {
            TRY(check_launch(ctx, ns));
            // the ragged rest: band symbols from m_start on, in a band-major scratch addressed as if it started at symbol 0
            const uint64_t pitch_all = (g.n_s + 8) / 9, pitch_t = pitch_all - tail.m_start;
            uint8_t* syt = nullptr;
            TRY(reserve_t(ctx, B_TMP, 9 * pitch_t + 16, &syt, s));
            uint8_t* sy0 = reinterpret_cast<uint8_t*>(reinterpret_cast<uintptr_t>(syt) - tail.m_start);
            for (size_t f = 0; f < n_frames; ++f) {
                CU(cudaMemsetAsync(syt, 0, 9 * pitch_t + 16, s));
                int n = launch_decode_fixed_general_from(ctx->tabs, g, d_in + 9 * stride_words * f, sy0, pitch_t, d_status + 2 * f, s, tail.cs);
                n += launch_regroup_rgb(sy0, g.n_s, g.tile_area, g.tile_w, d_rgb + 3 * n_px * f, px_out, s, pitch_t, 6 * (size_t)tail.unit_start);
                TRY(check_launch(ctx, n));
            }
            return T3C_OK;
        }
    }
    uint8_t* sy = nullptr;
    const uint64_t pitch = (g.n_s + 8) / 9;
    TRY(reserve_t(ctx, B_TMP, 9 * pitch + 16, &sy, s));
    for (size_t f = 0; f < n_frames; ++f) {
        CU(cudaMemsetAsync(sy, 0, 9 * pitch + 16, s));
        int n = launch_decode_fixed_general(ctx->tabs, g, d_in + 9 * stride_words * f, sy, pitch, d_status + 2 * f, s);
        n += launch_regroup_rgb(sy, g.n_s, g.tile_area, g.tile_w, d_rgb + 3 * n_px * f, px_out, s, pitch);
        TRY(check_launch(ctx, n));
    }
    return T3C_OK;
}

// =============================================================================================
// host-buffer API: stage in, run, stage out
// =============================================================================================
#define H2D(dst, src, bytes) CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream))
#define D2H(dst, src, bytes) CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream))
#define SYNC() CU(cudaStreamSynchronize(ctx->stream))
#define MAIL_UP() CU(cudaMemcpyAsync(ctx->d_mail, ctx->h_mail, sizeof(t3c_ctx::Mail), cudaMemcpyHostToDevice, ctx->stream))
#define MAIL_DOWN()                                                                                               \
    do {                                                                                                          \
        CU(cudaMemcpyAsync(ctx->h_mail, ctx->d_mail, sizeof(t3c_ctx::Mail), cudaMemcpyDeviceToHost, ctx->stream)); \
        SYNC();                                                                                                   \
    } while (0)

t3c_status t3c_rgb_to_quant(t3c_ctx* ctx, const uint8_t* rgb, size_t n_px, t3c_pixel* out)
{
    if (!ctx || (n_px && (!rgb || !out))) return fail(ctx, T3C_ERR_ARG, "rgb_to_quant: null");
    if (!n_px) return T3C_OK;
    DeviceGuard guard(ctx->device);
    host_buffer(ctx, rgb, 3 * n_px); host_buffer(ctx, out, 6 * n_px);
    uint8_t* d_in; t3c_pixel* d_out;
    TRY(reserve_t(ctx, B_IN, 3 * n_px, &d_in)); TRY(reserve_t(ctx, B_OUT, 6 * n_px, &d_out));
    H2D(d_in, rgb, 3 * n_px);
    TRY(t3c_rgb_to_quant_dev(ctx, d_in, n_px, d_out, ctx->stream));
    D2H(out, d_out, 6 * n_px);
    SYNC();
    return T3C_OK;
}
t3c_status t3c_quant_to_rgb(t3c_ctx* ctx, const t3c_pixel* px, size_t n_px, uint8_t* rgb)
{
    if (!ctx || (n_px && (!rgb || !px))) return fail(ctx, T3C_ERR_ARG, "quant_to_rgb: null");
    if (!n_px) return T3C_OK;
    DeviceGuard guard(ctx->device);
    host_buffer(ctx, px, 6 * n_px); host_buffer(ctx, rgb, 3 * n_px);
    t3c_pixel* d_in; uint8_t* d_out;
    TRY(reserve_t(ctx, B_IN, 6 * n_px, &d_in)); TRY(reserve_t(ctx, B_OUT, 3 * n_px, &d_out));
    H2D(d_in, px, 6 * n_px);
    TRY(t3c_quant_to_rgb_dev(ctx, d_in, n_px, d_out, ctx->stream));
    D2H(rgb, d_out, 3 * n_px);
    SYNC();
    return T3C_OK;
}
t3c_status t3c_pack_pixels(t3c_ctx* ctx, const t3c_pixel* px, size_t n_px, uint8_t* words, size_t* n_words)
{
    if (!ctx || (n_px && (!px || !words))) return fail(ctx, T3C_ERR_ARG, "pack_pixels: null");
    const size_t nw = (n_px + 1) / 2;
    if (n_words) *n_words = nw;
    if (!n_px) return T3C_OK;
    DeviceGuard guard(ctx->device);
    host_buffer(ctx, px, 6 * n_px); host_buffer(ctx, words, 9 * ((n_px + 1) / 2));
    t3c_pixel* d_in; uint8_t* d_out;
    TRY(reserve_t(ctx, B_IN, 6 * n_px, &d_in)); TRY(reserve_t(ctx, B_OUT, 9 * nw, &d_out));
    H2D(d_in, px, 6 * n_px);
    TRY(t3c_pack_pixels_dev(ctx, d_in, n_px, d_out, ctx->stream));
    D2H(words, d_out, 9 * nw);
    SYNC();
    return T3C_OK;
}
t3c_status t3c_unpack_pixels(t3c_ctx* ctx, const uint8_t* words, size_t n_words, t3c_pixel* px)
{
    if (!ctx || (n_words && (!px || !words))) return fail(ctx, T3C_ERR_ARG, "unpack_pixels: null");
    if (!n_words) return T3C_OK;
    DeviceGuard guard(ctx->device);
    host_buffer(ctx, words, 9 * n_words); host_buffer(ctx, px, 12 * n_words);
    uint8_t* d_in; t3c_pixel* d_out;
    TRY(reserve_t(ctx, B_IN, 9 * n_words, &d_in)); TRY(reserve_t(ctx, B_OUT, 12 * n_words, &d_out));
    H2D(d_in, words, 9 * n_words);
    TRY(t3c_unpack_pixels_dev(ctx, d_in, n_words, d_out, ctx->stream));
    D2H(px, d_out, 12 * n_words);
    SYNC();
    return T3C_OK;
}
t3c_status t3c_words_to_bytes(t3c_ctx* ctx, const uint8_t* words, size_t n_words, uint8_t* bytes)
{
    if (!ctx || (n_words && (!bytes || !words))) return fail(ctx, T3C_ERR_ARG, "words_to_bytes: null");
    if (!n_words) return T3C_OK;
    DeviceGuard guard(ctx->device);
    uint8_t *d_in, *d_out;
    TRY(reserve_t(ctx, B_IN, 9 * n_words, &d_in)); TRY(reserve_t(ctx, B_OUT, 9 * n_words, &d_out));
    H2D(d_in, words, 9 * n_words);
    TRY(check_launch(ctx, launch_mod27(d_in, 9 * n_words, d_out, ctx->stream)));
    D2H(bytes, d_out, 9 * n_words);
    SYNC();
    return T3C_OK;
}
t3c_status t3c_rs_encode_blocks(t3c_ctx* ctx, int k, int arith, const uint8_t* data, size_t n, uint8_t* out26)
{
    if (!ctx || !k_valid(k) || (n && (!data || !out26))) return fail(ctx, T3C_ERR_ARG, "rs_encode_blocks: bad k or null");
    if (!n) return T3C_OK;
    DeviceGuard guard(ctx->device);
    uint8_t *d_in, *d_out;
    TRY(reserve_t(ctx, B_IN, n * k, &d_in)); TRY(reserve_t(ctx, B_OUT, n * 26, &d_out));
    H2D(d_in, data, n * k);
    TRY(t3c_rs_encode_blocks_dev(ctx, k, arith, d_in, n, d_out, ctx->stream));
    D2H(out26, d_out, n * 26);
    SYNC();
    return T3C_OK;
}
t3c_status t3c_rs_decode_blocks(t3c_ctx* ctx, int k, int arith, uint8_t* inout26, size_t n, uint8_t* out_k, uint8_t* ok)
{
    if (!ctx || !k_valid(k) || (n && (!inout26 || !out_k || !ok))) return fail(ctx, T3C_ERR_ARG, "rs_decode_blocks: bad k or null");
    if (!n) return T3C_OK;
    DeviceGuard guard(ctx->device);
    uint8_t *d_io, *d_out, *d_ok;
    TRY(reserve_t(ctx, B_IN, n * 26, &d_io)); TRY(reserve_t(ctx, B_OUT, n * k, &d_out)); TRY(reserve_t(ctx, B_TMP, n, &d_ok));
    H2D(d_io, inout26, n * 26);
    TRY(t3c_rs_decode_blocks_dev(ctx, k, arith, d_io, n, d_out, d_ok, ctx->stream));
    D2H(inout26, d_io, n * 26);
    D2H(out_k, d_out, n * k);
    D2H(ok, d_ok, n);
    SYNC();
    return T3C_OK;
}
t3c_status t3c_interleave2d(t3c_ctx* ctx, uint8_t* syms, size_t n, uint16_t w, uint16_t h, int inverse)
{
    (void)inverse; // the boustrophedon permutation is an involution: one kernel serves both directions
    if (!ctx || (n && !syms)) return fail(ctx, T3C_ERR_ARG, "interleave2d: null");
    if (!n || !w || !h) return T3C_OK; // OLD:752
    DeviceGuard guard(ctx->device);
    uint8_t *d_in, *d_out;
    TRY(reserve_t(ctx, B_IN, n, &d_in)); TRY(reserve_t(ctx, B_OUT, n, &d_out));
    H2D(d_in, syms, n);
    TRY(check_launch(ctx, launch_perm2d(d_in, d_out, n, w, h, ctx->stream)));
    D2H(syms, d_out, n);
    SYNC();
    return T3C_OK;
}
t3c_status t3c_header_emit(t3c_ctx* ctx, const t3c_config* cfg, int arith, uint8_t hdr27[27], uint8_t coded52[52])
{
    if (!ctx || !cfg) return fail(ctx, T3C_ERR_ARG, "header_emit: null");
    DeviceGuard guard(ctx->device);
    TRY(check_launch(ctx, launch_header_emit(ctx->tabs, *cfg, arith, ctx->d_mail->hdr27, ctx->d_mail->coded52, ctx->stream)));
    MAIL_DOWN();
    if (hdr27) std::memcpy(hdr27, ctx->h_mail->hdr27, 27);
    if (coded52) std::memcpy(coded52, ctx->h_mail->coded52, 52);
    return T3C_OK;
}
t3c_status t3c_header_parse(t3c_ctx* ctx, int arith, const uint8_t* words, size_t n_words, t3c_config* out, int* ok)
{
    if (!ctx || !out || !ok || (n_words && !words)) return fail(ctx, T3C_ERR_ARG, "header_parse: null");
    DeviceGuard guard(ctx->device);
    *ok = 0;
    if (n_words < 6) return T3C_OK; // OLD:920
    uint8_t* d_in;
    TRY(reserve_t(ctx, B_IN, 54, &d_in));
    H2D(d_in, words, 54);
    ctx->h_mail->cfg = *out;
    ctx->h_mail->ok = 0;
    MAIL_UP();
    TRY(check_launch(ctx, launch_header_parse(ctx->tabs, arith, d_in, n_words, &ctx->d_mail->cfg, &ctx->d_mail->ok, ctx->stream)));
    MAIL_DOWN();
    *ok = ctx->h_mail->ok;
    if (*ok) *out = ctx->h_mail->cfg;
    return T3C_OK;
}

// ---- L0 / L1 names of the reference's public surface (single items)
t3c_status t3c_header_pack(t3c_ctx* ctx, const t3c_header* h, uint8_t sym27[27])
{
    if (!ctx || !h || !sym27) return fail(ctx, T3C_ERR_ARG, "header_pack: null");
    DeviceGuard guard(ctx->device);
    TRY(check_launch(ctx, launch_header_pack(h->cfg, h->magic, h->version, h->band_map_hash, h->frame_seq, ctx->d_mail->hdr27, ctx->stream)));
    MAIL_DOWN();
    std::memcpy(sym27, ctx->h_mail->hdr27, 27);
    return T3C_OK;
}
static t3c_status header_check_unpack(t3c_ctx* ctx, const uint8_t sym27[27], t3c_header* out, int* ok)
{
    DeviceGuard guard(ctx->device);
    std::memcpy(ctx->h_mail->hdr27, sym27, 27);
    std::memset(&ctx->h_mail->cfg, 0, sizeof(t3c_config));
    ctx->h_mail->cfg.superframe_words = 8192;
    ctx->h_mail->ok = 0;
    MAIL_UP();
    TRY(check_launch(ctx, launch_header_check_unpack(ctx->d_mail->hdr27, &ctx->d_mail->cfg, ctx->d_mail->status, &ctx->d_mail->ok, ctx->stream)));
    MAIL_DOWN();
    if (ok) *ok = ctx->h_mail->ok;
    if (out) {
        out->magic = (uint16_t)ctx->h_mail->status[0]; out->version = (uint8_t)ctx->h_mail->status[1]; out->pad_ = 0;
        out->band_map_hash = ctx->h_mail->status[2]; out->frame_seq = ctx->h_mail->status[3];
        out->cfg = ctx->h_mail->cfg;
    }
    return T3C_OK;
}
t3c_status t3c_header_check(t3c_ctx* ctx, const uint8_t sym27[27], int* ok)
{
    if (!ctx || !sym27 || !ok) return fail(ctx, T3C_ERR_ARG, "header_check: null");
    return header_check_unpack(ctx, sym27, nullptr, ok);
}
t3c_status t3c_header_unpack(t3c_ctx* ctx, const uint8_t sym27[27], t3c_header* out)
{
    if (!ctx || !sym27 || !out) return fail(ctx, T3C_ERR_ARG, "header_unpack: null");
    return header_check_unpack(ctx, sym27, out, nullptr);
}
t3c_status t3c_crc3_rem12(t3c_ctx* ctx, const uint8_t* trits, size_t n, uint8_t out12[12])
{
    if (!ctx || !out12 || (n && !trits)) return fail(ctx, T3C_ERR_ARG, "crc3_rem12: null");
    DeviceGuard guard(ctx->device);
    uint8_t* d_in;
    TRY(reserve_t(ctx, B_IN, n + 16, &d_in));
    if (n) H2D(d_in, trits, n);
    TRY(check_launch(ctx, launch_crc3_rem12(d_in, n, ctx->d_mail->hdr27, ctx->stream)));
    MAIL_DOWN();
    std::memcpy(out12, ctx->h_mail->hdr27, 12);
    return T3C_OK;
}
t3c_status t3c_scramble_symbols(t3c_ctx* ctx, uint8_t* syms, size_t n, uint32_t a, uint32_t b, uint32_t* st, int inverse)
{
    if (!ctx || !st || (n && !syms)) return fail(ctx, T3C_ERR_ARG, "scramble_symbols: null");
    if (!n) return T3C_OK;
    DeviceGuard guard(ctx->device);
    // the first step takes the caller's state as it is (uint32 arithmetic, OLD:83), the following ones the reduced states
    uint8_t st8[8];
    uint32_t s = (a * *st + b) % 3;
    st8[0] = (uint8_t)s;
    for (int p = 1; p < 8; ++p) { s = (a * s + b) % 3; st8[p] = (uint8_t)s; }
    uint8_t* d_io;
    TRY(reserve_t(ctx, B_IN, n, &d_io));
    H2D(d_io, syms, n);
    TRY(check_launch(ctx, launch_scramble(ctx->tabs, d_io, n, st8, inverse, ctx->stream)));
    D2H(syms, d_io, n);
    SYNC();
    *st = n <= 2 ? st8[n - 1] : st8[2 + (n - 3) % 6];
    return T3C_OK;
}
t3c_status t3c_beacon_symbol(t3c_ctx* ctx, int profile, uint32_t frame_seq_mod, uint32_t health_flags, uint8_t* sym)
{
    if (!ctx || !sym) return fail(ctx, T3C_ERR_ARG, "beacon_symbol: null");
    const unsigned p = (uint8_t)profile, s = (uint8_t)((uint16_t)frame_seq_mod % 5), h = (uint8_t)((uint8_t)health_flags % 3);
    *sym = (uint8_t)((p + 5 * s + 15 * h) % 27);   // metadata arithmetic, OLD:107-113 (the encoder's own beacon symbol is Geom::bsym)
    return T3C_OK;
}
t3c_status t3c_gf27_tables(t3c_ctx* ctx, t3c_gf27* out)
{
    if (!ctx || !out || !ctx->host) return fail(ctx, T3C_ERR_ARG, "gf27_tables: null");
    const GfTables& g = ctx->host->gf;
    std::memset(out, 0, sizeof *out);
    for (int i = 0; i < 78; ++i) out->exp[i] = g.exp[i % 26];
    for (int a = 0; a < 27; ++a) { out->log[a] = g.lg[a] == 255 ? (int16_t)-1 : (int16_t)g.lg[a]; out->inv[a] = g.inv[a]; }
    for (int a = 0; a < 27; ++a)
        for (int b = 0; b < 27; ++b) {
            out->mul[a * 27 + b] = g.mul[a * 27 + b];
            out->add[a * 27 + b] = g.add[a * 27 + b];
            out->sub[a * 27 + b] = g.add[a * 27 + g.neg[b]];
        }
    out->primitive = g.exp[1];
    return T3C_OK;
}

t3c_status t3c_encode_profile(t3c_ctx* ctx, const t3c_config* cfg, int arith, const uint8_t* raw, size_t n_words, uint8_t* out,
                              size_t cap_words, size_t* n_out)
{
    if (!ctx || !cfg || !out || !n_out || (n_words && !raw)) return fail(ctx, T3C_ERR_ARG, "encode_profile: null");
    DeviceGuard guard(ctx->device);
    host_buffer(ctx, raw, 9 * n_words); host_buffer(ctx, out, 9 * cap_words);
    const size_t no = profile_words(*cfg, n_words);
    *n_out = 0;
    if (cap_words < no) return fail(ctx, T3C_ERR_CAPACITY, "encode_profile: capacity");
    uint8_t *d_in, *d_out;
    TRY(reserve_t(ctx, B_IN, 9 * n_words, &d_in)); TRY(reserve_t(ctx, B_OUT, 9 * no, &d_out));
    if (n_words) H2D(d_in, raw, 9 * n_words);
    TRY(t3c_encode_profile_dev(ctx, cfg, arith, d_in, n_words, d_out, no, ctx->stream));
    if (no) D2H(out, d_out, 9 * no);
    SYNC();
    *n_out = no;
    return T3C_OK;
}

t3c_status t3c_decode_profile(t3c_ctx* ctx, t3c_config* seen, const uint8_t* in, size_t n_words, uint8_t* out, size_t cap_words,
                              size_t* n_out, int* ok)
{
    if (!ctx || !seen || !n_out || !ok || (n_words && !in) || (cap_words && !out)) return fail(ctx, T3C_ERR_ARG, "decode_profile: null");
    DeviceGuard guard(ctx->device);
    host_buffer(ctx, in, 9 * n_words); host_buffer(ctx, out, 9 * cap_words);
    *n_out = 0; *ok = 0;
    if (seen->profile == T3C_PROFILE_RAW) { // stateful passthrough keyed on the PREVIOUS header, OLD:998-1002
        if (cap_words < n_words) return fail(ctx, T3C_ERR_CAPACITY, "decode_profile: capacity");
        if (n_words) std::memcpy(out, in, 9 * n_words); // out=in: a host-to-host copy of the caller's own buffers, no codec work
        *n_out = n_words; *ok = 1;
        return T3C_OK;
    }
    if (n_words < 6) return T3C_OK;
    uint8_t* d_in;
    TRY(reserve_t(ctx, B_IN, 9 * n_words, &d_in));
    H2D(d_in, in, 9 * n_words);
    ctx->h_mail->cfg = *seen;
    ctx->h_mail->ok = 0;
    MAIL_UP();
    TRY(check_launch(ctx, launch_header_parse(ctx->tabs, 0, d_in, n_words, &ctx->d_mail->cfg, &ctx->d_mail->ok, ctx->stream)));
    MAIL_DOWN();
    if (!ctx->h_mail->ok) return T3C_OK;            // header rejected: cfg_last_seen untouched (OLD:1005)
    *seen = ctx->h_mail->cfg;                       // OLD:1006-1013, before the body is looked at
    RefDecGeom g;
    ref_dec_geom(*seen, n_words, g);
    const size_t nw = (size_t)(3 * g.n_use / 26);
    uint8_t *use, *d_out;
    TRY(reserve_t(ctx, B_TMP, g.n_use + 16, &use)); TRY(reserve_t(ctx, B_OUT, 9 * nw + 16, &d_out));
    int n = launch_init_status(ctx->d_mail->status, 1, ctx->stream);
    n += launch_decode_ref_general(ctx->tabs, g, d_in, use, ctx->d_mail->status, ctx->stream);
    TRY(check_launch(ctx, n));
    MAIL_DOWN();
    if (!ctx->h_mail->status[0]) return T3C_OK;     // some block failed: out stays empty (OLD:987,997)
    if (cap_words < nw) return fail(ctx, T3C_ERR_CAPACITY, "decode_profile: capacity");
    TRY(check_launch(ctx, launch_regroup_words(use, g.n_use, g.tile_area, g.tile_w, d_out, nw, ctx->stream)));
    if (nw) D2H(out, d_out, 9 * nw);
    SYNC();
    *n_out = nw; *ok = 1;
    return T3C_OK;
}

t3c_status t3c_decode_profile_fixed(t3c_ctx* ctx, const t3c_config* cfg, size_t n_raw_words, const uint8_t* in, size_t n_words,
                                    uint8_t* out, size_t cap_words, size_t* n_out, int* ok, size_t* n_corrected)
{
    if (!ctx || !cfg || !n_out || !ok || (n_words && !in)) return fail(ctx, T3C_ERR_ARG, "decode_profile_fixed: null");
    DeviceGuard guard(ctx->device);
    host_buffer(ctx, in, 9 * n_words); host_buffer(ctx, out, 9 * cap_words);
    *n_out = 0; *ok = 0;
    if (n_corrected) *n_corrected = 0;
    if (cfg->profile == T3C_PROFILE_RAW) {
        if (cap_words < n_words) return fail(ctx, T3C_ERR_CAPACITY, "decode_profile_fixed: capacity");
        if (n_words) std::memcpy(out, in, 9 * n_words);
        *n_out = n_words; *ok = 1;
        return T3C_OK;
    }
    Geom g;
    if (n_raw_words) { make_geom(*cfg, n_raw_words, 1, g); if (g.n_out != n_words) return T3C_OK; }
    else {
        if (!geom_from_nout(*cfg, n_words, 1, g)) return T3C_OK;
        if (use_2d(*cfg) && g.l_body) return T3C_OK; // the last partial tile's permutation depends on N_w
        n_raw_words = (size_t)g.n_words;
    }
    uint64_t nw = 3 * known_prefix(*cfg, g) / 26;
    if (nw > n_raw_words) nw = n_raw_words;
    if (cap_words < nw) return fail(ctx, T3C_ERR_CAPACITY, "decode_profile_fixed: capacity");
    if (!g.n_cw || !nw) { *ok = 1; return T3C_OK; }
    uint8_t *d_in, *d_out;
    TRY(reserve_t(ctx, B_IN, 9 * n_words, &d_in)); TRY(reserve_t(ctx, B_OUT, 9 * nw + 16, &d_out));
    H2D(d_in, in, 9 * n_words);
    TRY(t3c_decode_profile_fixed_dev(ctx, cfg, n_raw_words, d_in, n_words, d_out, (size_t)nw, ctx->d_mail->status, ctx->stream));
    D2H(out, d_out, 9 * nw);
    MAIL_DOWN();
    if (!ctx->h_mail->status[0]) return T3C_OK;
    *n_out = (size_t)nw; *ok = 1;
    if (n_corrected) *n_corrected = ctx->h_mail->status[1];
    return T3C_OK;
}

t3c_status t3c_encode_frames_rgb8(t3c_ctx* ctx, const t3c_config* cfg, int arith, const uint8_t* rgb, size_t n_px, size_t n_frames,
                                  uint8_t* out, size_t stride_words, size_t* words_per_frame)
{
    if (!ctx || !cfg || !rgb || !out || !words_per_frame) return fail(ctx, T3C_ERR_ARG, "encode_frames: null");
    DeviceGuard guard(ctx->device);
    host_buffer(ctx, rgb, 3 * n_px * n_frames); host_buffer(ctx, out, 9 * stride_words * n_frames);
    const size_t n_words = (n_px + 1) / 2;
    const size_t no = profile_words(*cfg, n_words);
    *words_per_frame = no;
    if (stride_words < no) return fail(ctx, T3C_ERR_CAPACITY, "encode_frames: stride");
    if (!n_frames) return T3C_OK;
    uint8_t *d_in, *d_out;
    const size_t dstride = (no + 15) & ~(size_t)15; // device-side frame pitch: keeps every frame 16-byte aligned
    Geom g;
    uint32_t n_full = 0;
    if (cfg->profile != T3C_PROFILE_RAW && fast_path_ok(*cfg)) { make_geom(*cfg, n_words, arith, g); n_full = fast_full_tiles_encode(g, n_px); }
    if (n_full < kPipeMinTiles) { // one piece: stage in, run, stage out
        TRY(reserve_t(ctx, B_IN, 3 * n_px * n_frames, &d_in)); TRY(reserve_t(ctx, B_OUT, 9 * dstride * n_frames, &d_out));
        H2D(d_in, rgb, 3 * n_px * n_frames);
        TRY(t3c_encode_frames_rgb8_dev(ctx, cfg, arith, d_in, n_px, n_frames, d_out, dstride, ctx->stream));
        CU(cudaMemcpy2DAsync(out, 9 * stride_words, d_out, 9 * dstride, 9 * no, n_frames, cudaMemcpyDeviceToHost, ctx->stream));
        SYNC();
        return T3C_OK;
    }
    // Chunked pipeline over the tiled fast path: the H2D copy of chunk c+1, the kernel of chunk c and the D2H copies of
    // chunk c-1 (nine band segments each) overlap on three streams; both PCIe directions stay busy.
    const size_t in_pitch = (3 * n_px + 15) & ~(size_t)15;
    TRY(reserve_t(ctx, B_IN, in_pitch * n_frames, &d_in)); TRY(reserve_t(ctx, B_OUT, 9 * dstride * n_frames, &d_out));
    const size_t tile_bytes = 3 * 27 * (size_t)g.uniform_k;
    bool uniform_bands = true;
    for (int b = 1; b < 9; ++b) uniform_bands = uniform_bands && g.ncw[b] == g.ncw[0];
    for (size_t f = 0; f < n_frames; ++f) {
        const uint8_t* h_in = rgb + 3 * n_px * f;
        uint8_t* h_out = out + 9 * stride_words * f;
        uint8_t* df_in = d_in + in_pitch * f;
        uint8_t* df_out = d_out + 9 * dstride * f;
        for (uint32_t c = 0; c <= kPipeChunks; ++c) { // the last round is the ragged tail + header
            const bool tail = c == kPipeChunks;
            const uint32_t t0 = tail ? n_full : pipe_edge(n_full, c, true), t1 = tail ? n_full : pipe_edge(n_full, c + 1, true);
            const size_t b0 = tile_bytes * t0, b1 = tail ? 3 * n_px : tile_bytes * t1;
            if (b1 > b0) CU(cudaMemcpyAsync(df_in + b0, h_in + b0, b1 - b0, cudaMemcpyHostToDevice, ctx->s_h2d));
            TRY(chain(ctx, ctx->s_h2d, ctx->stream));
            const int n = launch_encode_rgb_fast_part(ctx->tabs, *cfg, g, df_in, n_px, in_pitch, 1, df_out, dstride, ctx->stream, t0, t1, tail);
            if (n < 0) return fail(ctx, T3C_ERR_CUDA, "encode_frames: unaligned staging");
            TRY(check_launch(ctx, n));
            TRY(chain(ctx, ctx->stream, ctx->s_d2h));
            if (uniform_bands) { // nine equal segments, equally spaced: one strided copy (a copy command costs ~10 us of the copy engine's time)
                const uint64_t seg = tail ? 26 * (g.ncw[0] - 13ull * t0) : 26 * 13ull * (t1 - t0);
                if (seg) CU(cudaMemcpy2DAsync(h_out + 52 + 26 * 13ull * t0, 26 * g.ncw[0], df_out + 52 + 26 * 13ull * t0, 26 * g.ncw[0], seg, 9, cudaMemcpyDeviceToHost,
                                              ctx->s_d2h));
            } else
                for (int b = 0; b < 9; ++b) {
                    const uint64_t c0 = g.cw_base[b] + 13ull * t0, c1 = tail ? g.cw_base[b] + g.ncw[b] : g.cw_base[b] + 13ull * t1;
                    if (c1 > c0) CU(cudaMemcpyAsync(h_out + 52 + 26 * c0, df_out + 52 + 26 * c0, 26 * (c1 - c0), cudaMemcpyDeviceToHost, ctx->s_d2h));
                }
            if (tail) {
                CU(cudaMemcpyAsync(h_out, df_out, 52, cudaMemcpyDeviceToHost, ctx->s_d2h));
                const size_t body_end = 52 + (size_t)g.l_body;
                if (9 * no > body_end) CU(cudaMemcpyAsync(h_out + body_end, df_out + body_end, 9 * no - body_end, cudaMemcpyDeviceToHost, ctx->s_d2h));
            }
        }
    }
    CU(cudaStreamSynchronize(ctx->s_d2h));
    SYNC();
    return T3C_OK;
}

t3c_status t3c_decode_frames_rgb8(t3c_ctx* ctx, const t3c_config* cfg, const uint8_t* in, size_t words_per_frame, size_t stride_words,
                                  size_t n_frames, size_t n_px, uint8_t* rgb, uint8_t* ok, size_t* px_recovered, size_t* n_corrected)
{
    if (!ctx || !cfg || !in || !rgb || !ok) return fail(ctx, T3C_ERR_ARG, "decode_frames: null");
    if (cfg->profile == T3C_PROFILE_RAW) return fail(ctx, T3C_ERR_UNSUPPORTED, "decode_frames: RAW profile carries raw words, use unpack_pixels");
    if (n_frames > 32) {   // the status mailbox holds 32 frames: longer batches go through in pieces (same result, same order)
        size_t corrected = 0, part = 0;
        for (size_t f = 0; f < n_frames; f += 32) {
            const size_t nf = n_frames - f < 32 ? n_frames - f : 32;
            TRY(t3c_decode_frames_rgb8(ctx, cfg, in + 9 * stride_words * f, words_per_frame, stride_words, nf, n_px, rgb + 3 * n_px * f, ok + f, px_recovered, &part));
            corrected += part;
        }
        if (n_corrected) *n_corrected = corrected;
        return T3C_OK;
    }
    DeviceGuard guard(ctx->device);
    host_buffer(ctx, in, 9 * stride_words * n_frames); host_buffer(ctx, rgb, 3 * n_px * n_frames);
    if (n_corrected) *n_corrected = 0;
    if (px_recovered) *px_recovered = 0;
    if (!n_frames) return T3C_OK;
    const size_t n_words = (n_px + 1) / 2;
    Geom g;
    make_geom(*cfg, n_words, 1, g);
    if (g.n_out != words_per_frame) return fail(ctx, T3C_ERR_ARG, "decode_frames: words_per_frame does not match config and n_px");
    uint64_t nw = 3 * known_prefix(*cfg, g) / 26;
    if (nw > n_words) nw = n_words;
    const size_t px_out = 2 * (size_t)nw < n_px ? 2 * (size_t)nw : n_px;
    uint8_t *d_in, *d_out;
    const size_t dstride = (words_per_frame + 15) & ~(size_t)15;
    const size_t out_pitch = (3 * n_px + 15) & ~(size_t)15;
    const uint32_t n_full = fast_path_ok(*cfg) ? fast_full_tiles_decode(g, px_out, out_pitch, 1) : 0;
    if (n_full < kPipeMinTiles) { // one piece
        TRY(reserve_t(ctx, B_IN, 9 * dstride * n_frames, &d_in)); TRY(reserve_t(ctx, B_OUT, 3 * n_px * n_frames, &d_out));
        CU(cudaMemcpy2DAsync(d_in, 9 * dstride, in, 9 * stride_words, 9 * words_per_frame, n_frames, cudaMemcpyHostToDevice, ctx->stream));
        TRY(t3c_decode_frames_rgb8_dev(ctx, cfg, d_in, words_per_frame, dstride, n_frames, n_px, d_out, ctx->d_mail->status, ctx->stream));
        for (size_t f = 0; f < n_frames; ++f) if (px_out) D2H(rgb + 3 * n_px * f, d_out + 3 * n_px * f, 3 * px_out);
    } else { // chunked pipeline, see t3c_encode_frames_rgb8
        TRY(reserve_t(ctx, B_IN, 9 * dstride * n_frames, &d_in)); TRY(reserve_t(ctx, B_OUT, out_pitch * n_frames, &d_out));
        TRY(check_launch(ctx, launch_init_status(ctx->d_mail->status, n_frames, ctx->stream)));
        uint32_t chk_nz[7], chk_two[7];
        fast_check_constants(*ctx->host, g, chk_nz, chk_two);
        const size_t tile_bytes = 3 * 27 * (size_t)g.uniform_k, frame_bytes = 9 * words_per_frame;
        bool uniform_bands = true;
        for (int b = 1; b < 9; ++b) uniform_bands = uniform_bands && g.ncw[b] == g.ncw[0];
        // (the header's 52 bytes are not needed: the consistent decoder takes its config out of band)
        for (size_t f = 0; f < n_frames; ++f) {
            const uint8_t* h_in = in + 9 * stride_words * f;
            uint8_t* h_out = rgb + 3 * n_px * f;
            uint8_t* df_in = d_in + 9 * dstride * f;
            uint8_t* df_out = d_out + out_pitch * f;
            for (uint32_t c = 0; c <= kPipeChunks; ++c) {
                const bool tail = c == kPipeChunks;
                const uint32_t t0 = tail ? n_full : pipe_edge(n_full, c, false), t1 = tail ? n_full : pipe_edge(n_full, c + 1, false);
                // band segments, widened by 16 bytes: the kernels load whole 16-byte chunks around every run
                const uint64_t seg = tail ? 26 * (g.ncw[0] - 13ull * t0) : 26 * 13ull * (t1 - t0);
                if (uniform_bands && 52 + 26 * (g.cw_base[8] + 13ull * t0) + seg + 16 <= frame_bytes) { // equal, equally spaced: one strided copy
                    if (seg) CU(cudaMemcpy2DAsync(df_in + 52 + 26 * 13ull * t0, 26 * g.ncw[0], h_in + 52 + 26 * 13ull * t0, 26 * g.ncw[0], seg + 16, 9, cudaMemcpyHostToDevice,
                                                  ctx->s_h2d));
                } else
                    for (int b = 0; b < 9; ++b) {
                        const uint64_t c0 = g.cw_base[b] + 13ull * t0, c1 = tail ? g.cw_base[b] + g.ncw[b] : g.cw_base[b] + 13ull * t1;
                        if (c1 <= c0) continue;
                        size_t lo = (52 + 26 * c0) & ~(size_t)15, hi = (52 + 26 * c1 + 15) & ~(size_t)15;
                        if (hi > frame_bytes) hi = frame_bytes;
                        CU(cudaMemcpyAsync(df_in + lo, h_in + lo, hi - lo, cudaMemcpyHostToDevice, ctx->s_h2d));
                    }
                TRY(chain(ctx, ctx->s_h2d, ctx->stream));
                const int n = launch_decode_rgb_fast_part(ctx->tabs, g, df_in, dstride, 1, n_px, out_pitch, px_out, df_out, ctx->d_mail->status + 2 * f,
                                                          ctx->stream, chk_nz, chk_two, t0, t1, tail);
                if (n < 0) return fail(ctx, T3C_ERR_CUDA, "decode_frames: unaligned staging");
                TRY(check_launch(ctx, n));
                TRY(chain(ctx, ctx->stream, ctx->s_d2h));
                const size_t b0 = tile_bytes * t0, b1 = tail ? 3 * px_out : tile_bytes * t1;
                if (b1 > b0) CU(cudaMemcpyAsync(h_out + b0, df_out + b0, b1 - b0, cudaMemcpyDeviceToHost, ctx->s_d2h));
            }
        }
        CU(cudaStreamSynchronize(ctx->s_d2h));
    }
    MAIL_DOWN();
    for (size_t f = 0; f < n_frames; ++f) {
        ok[f] = ctx->h_mail->status[2 * f] ? 1 : 0;
        if (n_corrected) *n_corrected += ctx->h_mail->status[2 * f + 1];
    }
    if (px_recovered) *px_recovered = px_out;
    return T3C_OK;
}

// =============================================================================================
// sub-word streams + base-243 (8(f).2), NEW-generation RAW path (8(f).3)
// =============================================================================================
static bool subword_ok(int sub) { return sub == 0 || sub == 27 || sub == 24 || sub == 21 || sub == 18 || sub == 15; } // is_valid_subword, NEW:118-123

t3c_status t3c_subword_stream_dev(t3c_ctx* ctx, const uint8_t* d_words, size_t n_words, int N, uint8_t* d_trits, void* st)
{
    if (!ctx || N < 0 || N > 27 || (n_words && N && (!d_words || !d_trits))) return fail(ctx, T3C_ERR_ARG, "subword_stream: bad N or null");
    DeviceGuard guard(ctx->device);
    return check_launch(ctx, launch_subword_stream(d_words, n_words, N, d_trits, (cudaStream_t)st));
}
t3c_status t3c_words_from_subword_stream_dev(t3c_ctx* ctx, const uint8_t* d_trits, size_t n_trits, int N, uint8_t fill, uint8_t* d_words, void* st)
{
    if (!ctx || N < 1 || N > 27 || (n_trits && (!d_words || !d_trits))) return fail(ctx, T3C_ERR_ARG, "words_from_subword_stream: bad N or null");
    DeviceGuard guard(ctx->device);
    return check_launch(ctx, launch_words_from_subword_stream(d_trits, n_trits, N, fill, d_words, (cudaStream_t)st));
}
t3c_status t3c_base243_pack_dev(t3c_ctx* ctx, const uint8_t* d_trits, size_t n_trits, uint8_t* d_out, void* st)
{
    if (!ctx || !d_out || (n_trits && !d_trits)) return fail(ctx, T3C_ERR_ARG, "base243_pack: null");
    DeviceGuard guard(ctx->device);
    return check_launch(ctx, launch_base243_pack(d_trits, n_trits, d_out, (cudaStream_t)st));
}
t3c_status t3c_base243_unpack_dev(t3c_ctx* ctx, const uint8_t* d_payload, size_t n_trits, uint8_t* d_trits, void* st)
{
    if (!ctx || (n_trits && (!d_payload || !d_trits))) return fail(ctx, T3C_ERR_ARG, "base243_unpack: null");
    DeviceGuard guard(ctx->device);
    return check_launch(ctx, launch_base243_unpack(d_payload, n_trits, d_trits, (cudaStream_t)st));
}
t3c_status t3c_words_to_base243_dev(t3c_ctx* ctx, const uint8_t* d_words, size_t n_words, int N, uint8_t* d_out, void* st)
{
    if (!ctx || N < 1 || N > 27 || !d_out || (n_words && !d_words)) return fail(ctx, T3C_ERR_ARG, "words_to_base243: bad N or null");
    DeviceGuard guard(ctx->device);
    return check_launch(ctx, launch_words_to_base243(d_words, n_words, N, d_out, (cudaStream_t)st));
}
t3c_status t3c_v6new_pack_pixels_dev(t3c_ctx* ctx, const t3c_pixel* d_px, size_t n_px, uint32_t* d_words, void* st)
{
    if (!ctx || (n_px && (!d_px || !d_words))) return fail(ctx, T3C_ERR_ARG, "v6new_pack_pixels: null");
    DeviceGuard guard(ctx->device);
    return check_launch(ctx, launch_v6new_pack_pixels(d_px, n_px, d_words, (cudaStream_t)st));
}
t3c_status t3c_v6new_unpack_pixels_dev(t3c_ctx* ctx, const uint32_t* d_words, size_t n_words, t3c_pixel* d_px, void* st)
{
    if (!ctx || (n_words && (!d_px || !d_words))) return fail(ctx, T3C_ERR_ARG, "v6new_unpack_pixels: null");
    DeviceGuard guard(ctx->device);
    return check_launch(ctx, launch_v6new_unpack_pixels(d_words, n_words, d_px, (cudaStream_t)st));
}

t3c_status t3c_subword_stream(t3c_ctx* ctx, const uint8_t* words, size_t n_words, int N, uint8_t* trits)
{
    if (!ctx || N < 0 || N > 27 || (n_words && N && (!words || !trits))) return fail(ctx, T3C_ERR_ARG, "subword_stream: bad N or null");
    if (!n_words || !N) return T3C_OK;
    DeviceGuard guard(ctx->device);
    uint8_t *d_in, *d_out;
    TRY(reserve_t(ctx, B_IN, 9 * n_words, &d_in)); TRY(reserve_t(ctx, B_OUT, n_words * (size_t)N, &d_out));
    H2D(d_in, words, 9 * n_words);
    TRY(t3c_subword_stream_dev(ctx, d_in, n_words, N, d_out, ctx->stream));
    D2H(trits, d_out, n_words * (size_t)N);
    SYNC();
    return T3C_OK;
}
t3c_status t3c_words_from_subword_stream(t3c_ctx* ctx, const uint8_t* trits, size_t n_trits, int N, uint8_t fill, uint8_t* words, size_t* n_words)
{
    if (!ctx || !n_words || N < 1 || N > 27 || (n_trits && (!words || !trits))) return fail(ctx, T3C_ERR_ARG, "words_from_subword_stream: bad N or null");
    const size_t nw = (n_trits + (size_t)N - 1) / (size_t)N;
    *n_words = nw;
    if (!nw) return T3C_OK;
    DeviceGuard guard(ctx->device);
    uint8_t *d_in, *d_out;
    TRY(reserve_t(ctx, B_IN, n_trits, &d_in)); TRY(reserve_t(ctx, B_OUT, 9 * nw, &d_out));
    H2D(d_in, trits, n_trits);
    TRY(t3c_words_from_subword_stream_dev(ctx, d_in, n_trits, N, fill, d_out, ctx->stream));
    D2H(words, d_out, 9 * nw);
    SYNC();
    return T3C_OK;
}
t3c_status t3c_base243_pack(t3c_ctx* ctx, const uint8_t* trits, size_t n_trits, uint8_t* out, size_t* n_bytes)
{
    if (!ctx || !out || !n_bytes || (n_trits && !trits)) return fail(ctx, T3C_ERR_ARG, "base243_pack: null");
    const size_t nb = 4 + (n_trits + 4) / 5;
    *n_bytes = nb;
    DeviceGuard guard(ctx->device);
    uint8_t *d_in, *d_out;
    TRY(reserve_t(ctx, B_IN, n_trits, &d_in)); TRY(reserve_t(ctx, B_OUT, nb, &d_out));
    if (n_trits) H2D(d_in, trits, n_trits);
    TRY(t3c_base243_pack_dev(ctx, d_in, n_trits, d_out, ctx->stream));
    D2H(out, d_out, nb);
    SYNC();
    return T3C_OK;
}
t3c_status t3c_base243_unpack(t3c_ctx* ctx, const uint8_t* in, size_t n_bytes, uint8_t* trits, size_t cap, size_t* n_trits, int* ok)
{
    if (!ctx || !n_trits || !ok || (n_bytes && !in)) return fail(ctx, T3C_ERR_ARG, "base243_unpack: null");
    *n_trits = 0; *ok = 0;
    if (n_bytes < 4) return T3C_OK;                                   // base243_to_ut: false
    uint32_t total = 0;
    std::memcpy(&total, in, 4);                                       // the count is container metadata, read on the host like a header
    const size_t avail = 5 * (n_bytes - 4), n = total < avail ? total : avail;
    *n_trits = n; *ok = n == total;
    const size_t nw = n < cap ? n : cap;
    if (!nw) return T3C_OK;
    if (!trits) return fail(ctx, T3C_ERR_ARG, "base243_unpack: null output");
    DeviceGuard guard(ctx->device);
    uint8_t *d_in, *d_out;
    TRY(reserve_t(ctx, B_IN, n_bytes, &d_in)); TRY(reserve_t(ctx, B_OUT, nw, &d_out));
    H2D(d_in, in + 4, n_bytes - 4);
    TRY(t3c_base243_unpack_dev(ctx, d_in, nw, d_out, ctx->stream));
    D2H(trits, d_out, nw);
    SYNC();
    return T3C_OK;
}
t3c_status t3c_words_to_base243(t3c_ctx* ctx, const uint8_t* words, size_t n_words, int N, uint8_t* out, size_t* n_bytes)
{
    if (!ctx || !out || !n_bytes || N < 1 || N > 27 || (n_words && !words)) return fail(ctx, T3C_ERR_ARG, "words_to_base243: bad N or null");
    const size_t nb = 4 + (n_words * (size_t)N + 4) / 5;
    *n_bytes = nb;
    DeviceGuard guard(ctx->device);
    uint8_t *d_in, *d_out;
    TRY(reserve_t(ctx, B_IN, 9 * n_words, &d_in)); TRY(reserve_t(ctx, B_OUT, nb, &d_out));
    if (n_words) H2D(d_in, words, 9 * n_words);
    TRY(t3c_words_to_base243_dev(ctx, d_in, n_words, N, d_out, ctx->stream));
    D2H(out, d_out, nb);
    SYNC();
    return T3C_OK;
}
t3c_status t3c_v6new_pack_pixels(t3c_ctx* ctx, const t3c_pixel* px, size_t n_px, uint32_t* words, int subword)
{
    if (!ctx || !subword_ok(subword) || (n_px && (!px || !words))) return fail(ctx, T3C_ERR_ARG, "v6new_pack_pixels: invalid subword mode or null");
    if (!n_px) return T3C_OK;
    DeviceGuard guard(ctx->device);
    t3c_pixel* d_in; uint32_t* d_out;
    TRY(reserve_t(ctx, B_IN, 6 * n_px, &d_in)); TRY(reserve_t(ctx, B_OUT, 4 * n_px, &d_out));
    H2D(d_in, px, 6 * n_px);
    TRY(t3c_v6new_pack_pixels_dev(ctx, d_in, n_px, d_out, ctx->stream));
    D2H(words, d_out, 4 * n_px);
    SYNC();
    return T3C_OK;
}
t3c_status t3c_v6new_unpack_pixels(t3c_ctx* ctx, const uint32_t* words, size_t n_words, t3c_pixel* px, int subword)
{
    if (!ctx || !subword_ok(subword) || (n_words && (!px || !words))) return fail(ctx, T3C_ERR_ARG, "v6new_unpack_pixels: invalid subword mode or null");
    if (!n_words) return T3C_OK;
    DeviceGuard guard(ctx->device);
    uint32_t* d_in; t3c_pixel* d_out;
    TRY(reserve_t(ctx, B_IN, 4 * n_words, &d_in)); TRY(reserve_t(ctx, B_OUT, 6 * n_words, &d_out));
    H2D(d_in, words, 4 * n_words);
    TRY(t3c_v6new_unpack_pixels_dev(ctx, d_in, n_words, d_out, ctx->stream));
    D2H(px, d_out, 6 * n_words);
    SYNC();
    return T3C_OK;
}


// ---- .t3v container records -----------------------------------------------------
t3c_status t3c_t3v_frame_records_dev(t3c_ctx* ctx, const uint8_t* d_words, size_t n_words, size_t stride_words, size_t n_frames, uint8_t* d_rec,
                                     size_t record_pitch, void* st)
{
    if (!ctx || !d_rec || (n_words && !d_words)) return fail(ctx, T3C_ERR_ARG, "t3v_frame_records: null");
    if (n_words > 0xFFFFFFFFull) return fail(ctx, T3C_ERR_ARG, "t3v_frame_records: the count is a uint32");
    if (record_pitch < 8 + 9 * n_words || (record_pitch & 3) || ((uintptr_t)d_rec & 3) || ((uintptr_t)d_words & 3) || (n_frames > 1 && ((9 * stride_words) & 3)))
        return fail(ctx, T3C_ERR_ARG, "t3v_frame_records: pitch / alignment");
    DeviceGuard guard(ctx->device);
    uint32_t* part = nullptr;
    TRY(reserve_t(ctx, B_TMP2, 4 * t3v_partial_words(n_words, n_frames), &part, (cudaStream_t)st));
    return check_launch(ctx, launch_t3v_records(ctx->tabs.crc, d_words, n_words, stride_words, n_frames, d_rec, record_pitch, part, (cudaStream_t)st));
}
t3c_status t3c_t3v_read_frames_dev(t3c_ctx* ctx, const uint8_t* d_rec, size_t record_pitch, size_t n_frames, size_t n_words, uint8_t* d_words,
                                   size_t stride_words, uint8_t* d_ok, void* st)
{
    if (!ctx || !d_rec || !d_ok) return fail(ctx, T3C_ERR_ARG, "t3v_read_frames: null");
    if (n_words > 0xFFFFFFFFull) return fail(ctx, T3C_ERR_ARG, "t3v_read_frames: the count is a uint32");
    if (record_pitch < 8 + 9 * n_words || (record_pitch & 3) || ((uintptr_t)d_rec & 3) || ((uintptr_t)d_words & 3) || (n_frames > 1 && ((9 * stride_words) & 3)))
        return fail(ctx, T3C_ERR_ARG, "t3v_read_frames: pitch / alignment");
    DeviceGuard guard(ctx->device);
    uint32_t* part = nullptr;
    TRY(reserve_t(ctx, B_TMP2, 4 * t3v_partial_words(n_words, n_frames), &part, (cudaStream_t)st));
    return check_launch(ctx, launch_t3v_read(ctx->tabs.crc, d_rec, record_pitch, n_frames, n_words, d_words, stride_words, part, d_ok, (cudaStream_t)st));
}
t3c_status t3c_crc32(t3c_ctx* ctx, const uint8_t* data, size_t n, uint32_t* crc)
{
    if (!ctx || !crc || (n && !data)) return fail(ctx, T3C_ERR_ARG, "crc32: null");
    DeviceGuard guard(ctx->device);
    uint8_t* d_in;
    uint32_t* part;
    TRY(reserve_t(ctx, B_IN, n + 16, &d_in));
    TRY(reserve_t(ctx, B_TMP2, 4 * t3v_partial_words((n + 8) / 9, 1) + 64, &part));
    if (n) H2D(d_in, data, n);
    TRY(check_launch(ctx, launch_crc32(ctx->tabs.crc, d_in, n, part + 16, part, ctx->stream)));
    D2H(&ctx->h_mail->status[0], part, 4);
    SYNC();
    *crc = ctx->h_mail->status[0];
    return T3C_OK;
}
t3c_status t3c_t3v_index_build(t3c_ctx* ctx, const uint64_t* n_words, size_t n_frames, uint64_t first_offset, uint8_t* out, size_t* n_bytes)
{
    if (!ctx || !out || !n_bytes || (n_frames && !n_words)) return fail(ctx, T3C_ERR_ARG, "t3v_index_build: null");
    if (n_frames > 0xFFFFFFFFull) return fail(ctx, T3C_ERR_ARG, "t3v_index_build: the frame count is a uint32");
    uint8_t h[17] = {'T', '3', 'V', 'I', 1};
    const uint32_t cnt = (uint32_t)n_frames, zero = 0;
    std::memcpy(h + 5, &cnt, 4); std::memcpy(h + 9, &zero, 4);
    uint32_t crc = 0;
    TRY(t3c_crc32(ctx, h, 13, &crc));          // device CRC-32, as for the .t3v header
    std::memcpy(h + 13, &crc, 4);
    std::memcpy(out, h, 17);
    uint64_t off = first_offset;                // record i = 4 + 9 n_i + 4 bytes (t3v_write_frame, old/include/t3v_io.hpp:128-142)
    for (size_t i = 0; i < n_frames; ++i) {
        if (n_words[i] > 0xFFFFFFFFull) return fail(ctx, T3C_ERR_ARG, "t3v_index_build: a record's count is a uint32");
        std::memcpy(out + 17 + 8 * i, &off, 8);
        off += 8 + 9 * n_words[i];
    }
    *n_bytes = 17 + 8 * n_frames;
    return T3C_OK;
}
t3c_status t3c_t3v_frame_record(t3c_ctx* ctx, const uint8_t* words, size_t n_words, uint8_t* record, size_t* n_bytes)
{
    if (!ctx || !record || !n_bytes || (n_words && !words)) return fail(ctx, T3C_ERR_ARG, "t3v_frame_record: null");
    if (n_words > 0xFFFFFFFFull) return fail(ctx, T3C_ERR_ARG, "t3v_frame_record: the count is a uint32");
    DeviceGuard guard(ctx->device);
    uint8_t *d_in, *d_out;
    const size_t pitch = (8 + 9 * n_words + 3) & ~(size_t)3;
    TRY(reserve_t(ctx, B_IN, 9 * n_words + 16, &d_in)); TRY(reserve_t(ctx, B_OUT, pitch + 32, &d_out));
    d_out += 12;                                                      // the record at 12 mod 16: its payload (at + 4) moves in 16-byte accesses
    if (n_words) H2D(d_in, words, 9 * n_words);
    TRY(t3c_t3v_frame_records_dev(ctx, d_in, n_words, n_words, 1, d_out, pitch, ctx->stream));
    D2H(record, d_out, 8 + 9 * n_words);
    SYNC();
    *n_bytes = 8 + 9 * n_words;
    return T3C_OK;
}
t3c_status t3c_t3v_read_frame(t3c_ctx* ctx, const uint8_t* record, size_t n_bytes, uint8_t* words, size_t cap_words, size_t* n_words, int* ok)
{
    if (!ctx || !n_words || !ok || (n_bytes && !record)) return fail(ctx, T3C_ERR_ARG, "t3v_read_frame: null");
    *n_words = 0; *ok = 0;
    if (n_bytes < 4) return T3C_OK;                                   // fread of the count fails
    uint32_t n = 0;
    std::memcpy(&n, record, 4);                                       // the count is container metadata, read on the host like a header
    if (n_bytes < 8 + 9 * (size_t)n) return T3C_OK;                   // fread of the payload / CRC fails
    if (cap_words < n) return fail(ctx, T3C_ERR_CAPACITY, "t3v_read_frame: capacity");
    DeviceGuard guard(ctx->device);
    uint8_t *d_in, *d_out;
    const size_t pitch = (8 + 9 * (size_t)n + 3) & ~(size_t)3;
    TRY(reserve_t(ctx, B_IN, pitch + 32, &d_in)); TRY(reserve_t(ctx, B_OUT, 9 * (size_t)n + 32, &d_out));
    d_in += 12;                                                       // as in t3c_t3v_frame_record
    H2D(d_in, record, 8 + 9 * (size_t)n);
    uint8_t* d_ok = d_out + ((9 * (size_t)n + 15) & ~(size_t)15);
    TRY(t3c_t3v_read_frames_dev(ctx, d_in, pitch, 1, n, d_out, n, d_ok, ctx->stream));
    D2H(&ctx->h_mail->status[0], d_ok, 1);
    SYNC();
    if (!(ctx->h_mail->status[0] & 0xFF)) return T3C_OK;              // CRC mismatch: the reference returns false and leaves `words` alone
    if (n) { D2H(words, d_out, 9 * (size_t)n); SYNC(); }
    *n_words = n; *ok = 1;
    return T3C_OK;
}
t3c_status t3c_t3v_header(t3c_ctx* ctx, uint8_t out[54], int profile, int subword_code, int centered, int coset, uint32_t width, uint32_t height,
                          const uint32_t aw[4], uint32_t fps_num, uint32_t fps_den, uint32_t frame_count, int file_type)
{
    if (!ctx || !out || !aw) return fail(ctx, T3C_ERR_ARG, "t3v_header: null");
    auto put32 = [](uint8_t* p, uint32_t v) { for (int i = 0; i < 4; ++i) p[i] = (uint8_t)(v >> (8 * i)); };
    std::memset(out, 0, 54);
    std::memcpy(out, "T3V1", 4);
    out[4] = 1; out[5] = (uint8_t)file_type; out[6] = (uint8_t)profile; out[7] = (uint8_t)subword_code; out[8] = centered ? 1 : 0; out[9] = (uint8_t)coset;
    put32(out + 10, width); put32(out + 14, height);
    for (int i = 0; i < 4; ++i) put32(out + 18 + 4 * i, aw[i]);
    put32(out + 34, fps_num); put32(out + 38, fps_den); put32(out + 42, frame_count);
    uint32_t crc = 0;
    TRY(t3c_crc32(ctx, out, 50, &crc));                               // the checksum itself comes from the device, like every other one
    put32(out + 50, crc);
    return T3C_OK;
}

// ---- image-bridge geometry ------------------------------------------------------
t3c_status t3c_resize_rgb_nn_dev(t3c_ctx* ctx, const uint8_t* d_src, int sw, int sh, uint8_t* d_dst, int dw, int dh, void* st)
{
    if (!ctx || dw < 0 || dh < 0 || (dw && dh && !d_dst) || (sw > 0 && sh > 0 && !d_src)) return fail(ctx, T3C_ERR_ARG, "resize_rgb_nn: null or negative size");
    DeviceGuard guard(ctx->device);
    return check_launch(ctx, launch_resize_rgb_nn(d_src, sw, sh, d_dst, dw, dh, (cudaStream_t)st));
}
t3c_status t3c_blit_center_rgb_dev(t3c_ctx* ctx, const uint8_t* d_src, int sw, int sh, uint8_t* d_dst, int cw, int ch, void* st)
{
    if (!ctx || cw < 0 || ch < 0 || sw < 0 || sh < 0 || (cw && ch && !d_dst) || (sw && sh && !d_src)) return fail(ctx, T3C_ERR_ARG, "blit_center_rgb: null or negative size");
    if (sw > cw) return fail(ctx, T3C_ERR_ARG, "blit_center_rgb: source wider than the canvas (the reference writes past the row)");
    DeviceGuard guard(ctx->device);
    return check_launch(ctx, launch_blit_center_rgb(d_src, sw, sh, d_dst, cw, ch, (cudaStream_t)st));
}
t3c_status t3c_extract_center_q_dev(t3c_ctx* ctx, const t3c_pixel* d_full, int fw, int fh, int sw, int sh, t3c_pixel* d_sub, void* st)
{
    if (!ctx || fw < 0 || fh < 0 || sw < 0 || sh < 0 || (sw && sh && (!d_sub || !d_full))) return fail(ctx, T3C_ERR_ARG, "extract_center_q: null or negative size");
    if (sw > fw) return fail(ctx, T3C_ERR_ARG, "extract_center_q: window wider than the frame (the reference reads past the row)");
    DeviceGuard guard(ctx->device);
    return check_launch(ctx, launch_extract_center_q(d_full, fw, fh, sw, sh, d_sub, (cudaStream_t)st));
}
t3c_status t3c_resize_rgb_nn(t3c_ctx* ctx, const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh)
{
    if (!ctx || sw < 0 || sh < 0 || dw < 0 || dh < 0) return fail(ctx, T3C_ERR_ARG, "resize_rgb_nn: negative size");
    DeviceGuard guard(ctx->device);
    uint8_t *d_in, *d_out;
    const size_t ni = (size_t)sw * sh * 3, no = (size_t)dw * dh * 3;
    if (!no) return T3C_OK;
    TRY(reserve_t(ctx, B_IN, ni + 16, &d_in)); TRY(reserve_t(ctx, B_OUT, no + 16, &d_out));
    if (ni) H2D(d_in, src, ni);
    TRY(t3c_resize_rgb_nn_dev(ctx, d_in, sw, sh, d_out, dw, dh, ctx->stream));
    D2H(dst, d_out, no);
    SYNC();
    return T3C_OK;
}
t3c_status t3c_blit_center_rgb(t3c_ctx* ctx, const uint8_t* src, int sw, int sh, uint8_t* dst, int cw, int ch)
{
    if (!ctx || sw < 0 || sh < 0 || cw < 0 || ch < 0) return fail(ctx, T3C_ERR_ARG, "blit_center_rgb: negative size");
    DeviceGuard guard(ctx->device);
    uint8_t *d_in, *d_out;
    const size_t ni = (size_t)sw * sh * 3, no = (size_t)cw * ch * 3;
    if (!no) return T3C_OK;
    TRY(reserve_t(ctx, B_IN, ni + 16, &d_in)); TRY(reserve_t(ctx, B_OUT, no + 16, &d_out));
    if (ni) H2D(d_in, src, ni);
    TRY(t3c_blit_center_rgb_dev(ctx, d_in, sw, sh, d_out, cw, ch, ctx->stream));
    D2H(dst, d_out, no);
    SYNC();
    return T3C_OK;
}
t3c_status t3c_extract_center_q(t3c_ctx* ctx, const t3c_pixel* full, int fw, int fh, int sw, int sh, t3c_pixel* sub)
{
    if (!ctx || fw < 0 || fh < 0 || sw < 0 || sh < 0) return fail(ctx, T3C_ERR_ARG, "extract_center_q: negative size");
    DeviceGuard guard(ctx->device);
    t3c_pixel *d_in, *d_out;
    const size_t ni = (size_t)fw * fh * 6, no = (size_t)sw * sh * 6;
    if (!no) return T3C_OK;
    TRY(reserve_t(ctx, B_IN, ni + 16, &d_in)); TRY(reserve_t(ctx, B_OUT, no + 16, &d_out));
    if (ni) H2D(d_in, full, ni);
    TRY(t3c_extract_center_q_dev(ctx, d_in, fw, fh, sw, sh, d_out, ctx->stream));
    D2H(sub, d_out, no);
    SYNC();
    return T3C_OK;
}
static bool v6new_std_res(int sub, int& w, int& h) // std_res_for, NEW:55-64
{
    switch (sub) {
    case 27: w = 7680; h = 4320; return true;
    case 24: w = 3840; h = 2160; return true;
    case 21: w = 1920; h = 1080; return true;
    case 18: w = 1280; h = 720; return true;
    case 15: w = 960; h = 540; return true;
    }
    w = h = 0;
    return false;
}
t3c_status t3c_v6new_image_to_words(t3c_ctx* ctx, const uint8_t* rgb, int w, int h, int subword, int centered, uint32_t* words, size_t cap_words,
                                    size_t* n_words, int* ok)
{
    if (!ctx || !n_words || !ok) return fail(ctx, T3C_ERR_ARG, "v6new_image_to_words: null");
    *n_words = 0; *ok = 0;
    int tw, th;
    if (!v6new_std_res(subword, tw, th) || w <= 0 || h <= 0 || !rgb) return T3C_OK;     // load fails / encode_*_subword rejects the mode
    const bool embed = centered && subword != 27;
    const int ow = embed ? 7680 : tw, oh = embed ? 4320 : th;
    const size_t n_px = (size_t)ow * oh;
    if (cap_words < n_px || !words) return fail(ctx, T3C_ERR_CAPACITY, "v6new_image_to_words: capacity");
    DeviceGuard guard(ctx->device);
    uint8_t *d_src, *d_work, *d_canvas;
    t3c_pixel* d_q;
    uint32_t* d_words;
    TRY(reserve_t(ctx, B_IN, (size_t)w * h * 3 + 16, &d_src));
    TRY(reserve_t(ctx, B_TMP, (size_t)tw * th * 3 + 16, &d_work));
    TRY(reserve_t(ctx, B_TMP2, embed ? n_px * 3 + 16 : 16, &d_canvas));
    TRY(reserve_t(ctx, B_AUX, n_px * 6 + 16, &d_q));
    TRY(reserve_t(ctx, B_OUT, n_px * 4 + 16, &d_words));
    H2D(d_src, rgb, (size_t)w * h * 3);
    const uint8_t* img = d_src;
    if (w != tw || h != th) { TRY(t3c_resize_rgb_nn_dev(ctx, d_src, w, h, d_work, tw, th, ctx->stream)); img = d_work; }
    if (embed) { TRY(t3c_blit_center_rgb_dev(ctx, img, tw, th, d_canvas, ow, oh, ctx->stream)); img = d_canvas; }
    TRY(check_launch(ctx, launch_rgb_to_quant(img, n_px, d_q, ctx->stream)));
    TRY(check_launch(ctx, launch_v6new_pack_pixels(d_q, n_px, d_words, ctx->stream)));
    D2H(words, d_words, 4 * n_px);
    SYNC();
    *n_words = n_px; *ok = 1;
    return T3C_OK;
}
t3c_status t3c_v6new_words_to_image(t3c_ctx* ctx, const uint32_t* words, size_t n_words, int subword, int w, int h, uint8_t* rgb, int* ok)
{
    if (!ctx || !ok || w < 0 || h < 0) return fail(ctx, T3C_ERR_ARG, "v6new_words_to_image: null");
    *ok = 0;
    int tw, th;
    if (!v6new_std_res(subword, tw, th)) return T3C_OK;                                   // decode_*_subword rejects the mode
    const size_t need = (size_t)w * h, full = (size_t)7680 * 4320;
    *ok = 1;
    if (!need) return T3C_OK;
    if (!rgb || (n_words && !words)) return fail(ctx, T3C_ERR_ARG, "v6new_words_to_image: null");
    DeviceGuard guard(ctx->device);
    uint32_t* d_words;
    t3c_pixel *d_q, *d_sub;
    uint8_t* d_rgb;
    TRY(reserve_t(ctx, B_IN, 4 * n_words + 16, &d_words));
    TRY(reserve_t(ctx, B_AUX, 6 * n_words + 16, &d_q));
    TRY(reserve_t(ctx, B_AUX2, 6 * (size_t)tw * th + 16, &d_sub));
    TRY(reserve_t(ctx, B_OUT, 3 * need + 16, &d_rgb));
    if (n_words) H2D(d_words, words, 4 * n_words);
    TRY(check_launch(ctx, launch_v6new_unpack_pixels(d_words, n_words, d_q, ctx->stream)));
    const t3c_pixel* q = d_q;
    size_t nq = n_words;
    if (n_words != need && n_words == full && subword != 27) { // the core returned an S27 canvas: its centre window
        TRY(t3c_extract_center_q_dev(ctx, d_q, 7680, 4320, tw, th, d_sub, ctx->stream));
        q = d_sub; nq = (size_t)tw * th;
    }
    const size_t fill = nq < need ? nq : need;                  // quant_stream_to_rgb stops when the pixels run out (:195)
    CU(cudaMemsetAsync(d_rgb, 0, 3 * need, ctx->stream));
    TRY(check_launch(ctx, launch_quant_to_rgb(q, fill, d_rgb, ctx->stream)));
    D2H(rgb, d_rgb, 3 * need);
    SYNC();
    return T3C_OK;
}
} // extern "C"


// ---- multi-device streams: one context and one host thread per lane, frame f -> lane f mod n (include/t3c.h) -----------------------
struct t3c_streamset {
    std::vector<t3c_ctx*> lane;
};
t3c_status t3c_stream_create(const int* devices, int n_lanes, t3c_streamset** out)
{
    if (!devices || n_lanes <= 0 || n_lanes > 64 || !out) return T3C_ERR_ARG;
    *out = nullptr;
    t3c_streamset* s = new t3c_streamset;
    for (int i = 0; i < n_lanes; ++i) {
        t3c_ctx* c = nullptr;
        const t3c_status st = t3c_create(devices[i], &c);
        if (st != T3C_OK) { t3c_stream_destroy(s); return st; }
        s->lane.push_back(c);
    }
    *out = s;
    return T3C_OK;
}
void t3c_stream_destroy(t3c_streamset* s)
{
    if (!s) return;
    for (t3c_ctx* c : s->lane) t3c_destroy(c);
    delete s;
}
int t3c_stream_lanes(const t3c_streamset* s) { return s ? (int)s->lane.size() : 0; }
namespace {
// lane l codes frames f with (first_frame + f) % n == l, one call per frame, on its own thread; the first failure is reported
template <class F>
t3c_status stream_run(t3c_streamset* s, size_t n_frames, size_t first_frame, F&& per_frame)
{
    const size_t n = s->lane.size();
    std::vector<t3c_status> res(n, T3C_OK);
    std::vector<std::thread> th;
    for (size_t l = 0; l < n; ++l)
        th.emplace_back([&, l] {
            for (size_t f = 0; f < n_frames; ++f) {
                if ((first_frame + f) % n != l) continue;
                const t3c_status st = per_frame(s->lane[l], f);
                if (st != T3C_OK) { res[l] = st; return; }
            }
        });
    for (auto& t : th) t.join();
    for (t3c_status r : res) if (r != T3C_OK) return r;
    return T3C_OK;
}
} // namespace
t3c_status t3c_stream_encode_rgb8(t3c_streamset* s, const t3c_config* cfg, int arith, const uint8_t* rgb, size_t n_px, size_t n_frames, size_t first_frame,
                                  uint8_t* out9, size_t stride_words, size_t* words_per_frame)
{
    if (!s || s->lane.empty() || !cfg || !words_per_frame || (n_frames && (!rgb || !out9))) return T3C_ERR_ARG;
    *words_per_frame = profile_words(*cfg, (n_px + 1) / 2);
    if (stride_words < *words_per_frame) return T3C_ERR_CAPACITY;
    return stream_run(s, n_frames, first_frame, [&](t3c_ctx* c, size_t f) {
        size_t w = 0;
        return t3c_encode_frames_rgb8(c, cfg, arith, rgb + 3 * n_px * f, n_px, 1, out9 + 9 * stride_words * f, stride_words, &w);
    });
}
t3c_status t3c_stream_decode_rgb8(t3c_streamset* s, const t3c_config* cfg, const uint8_t* in9, size_t words_per_frame, size_t stride_words, size_t n_frames,
                                  size_t first_frame, size_t n_px, uint8_t* rgb, uint8_t* ok, size_t* n_corrected)
{
    if (!s || s->lane.empty() || !cfg || (n_frames && (!in9 || !rgb || !ok))) return T3C_ERR_ARG;
    std::vector<size_t> fixed(n_frames, 0);
    const t3c_status st = stream_run(s, n_frames, first_frame, [&](t3c_ctx* c, size_t f) {
        size_t rec = 0;
        return t3c_decode_frames_rgb8(c, cfg, in9 + 9 * stride_words * f, words_per_frame, stride_words, 1, n_px, rgb + 3 * n_px * f, ok + f, &rec, &fixed[f]);
    });
    if (n_corrected) { *n_corrected = 0; for (size_t v : fixed) *n_corrected += v; }
    return st;
}
