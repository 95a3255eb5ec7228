// launch.h -- host-callable launchers implemented in the .cu files.  All take device pointers and
// enqueue on `st`; none synchronises.  Each returns the number of kernels it launched.
#pragma once
#include <cuda_runtime.h>

#include "t3c_internal.h"

namespace t3c {

// The coded super-frame header depends on the config only: it is emitted on the device (k_header_emit) when the config
// changes and kept in `d52`; every later frame copies the 52 symbols.
// One entry per (config, arithmetic) seen; an entry is written once (emit + stream synchronise) and never again while other entries are
// free, so kernels in flight on any stream that copy an entry's 52 symbols can never see another config's header.  When all entries
// are taken the device is synchronised before the oldest one is reused.
struct HeaderCache {
    static constexpr int N = 8;
    struct Entry { uint8_t* d52 = nullptr; uint8_t* d27 = nullptr; t3c_config cfg{}; int arith = -1; bool valid = false; } e[N];
    uint8_t* base = nullptr;   // N x 128 bytes
    int next = 0;
};
// Phase-B pass maps of the super-tile kernels (k_super.cuh): they depend on the per-band k and on cw_base mod 3 only; one slot
// per kernel flavour (encode / decode x RGB / raw words), re-uploaded when the key changes.
struct SuperCache {
    struct Slot { uint16_t* d_map = nullptr; uint8_t* d_kv = nullptr; uint8_t key[48] = {}; uint32_t npass[3] = {}; bool valid = false; } slot[4];
};
// The CTA-shared part of the v5 kernels' shared memory (plane tables with the embedded (de)scrambled symbols, parity / screen constants,
// GF(27) tables, phase-B lane records: 16-26 KB) depends on the config only.  It is built once by a one-CTA setup kernel and kept on
// the device; the kernels' prologue is then a plain copy (building it in every CTA cost ~8 us of a 150 us launch).  Entries are
// written once (kernel + stream synchronise) and reused; when all are taken the device is synchronised before one is replaced.
struct FastImageCache {
    static constexpr int N = 8, BYTES = 32 * 1024;
    struct Entry { uint8_t key[40] = {}; bool valid = false; } e[N];
    uint8_t* base = nullptr;   // N x BYTES
    int next = 0;
};
struct DevTables { const GfTables* gf; const RsTables* rs; int sm_count; HeaderCache* hdr; SuperCache* sup; const uint32_t* crc; FastImageCache* img; };
// first codeword of every band that the general kernels still have to code (the tiled kernels did the ones before)
struct CwStart { uint64_t c[9]; };
// what a super-tile launch leaves to the general kernels: codewords from cs.c[b] on, band symbols from m_start, 6-pixel units from unit_start
struct SuperTail { CwStart cs; uint64_t m_start; uint64_t unit_start; uint32_t n_tiles; uint32_t ncw_tile[9]; };

// geometry of the reference decoder as shipped (A.7): slot-major demap of words 6.. of the input
struct RefDecGeom {
    uint64_t n_body_words;
    uint64_t ncw[9], use_base[9], n_use;
    uint64_t tile_area;
    uint32_t tile_w;
    uint32_t period;   // 0 = no skipping
    int32_t  slot;
    int32_t  k[9];
    uint8_t  st[8];
};

int launch_init_status(uint32_t* d_status, size_t n_frames, cudaStream_t st); // {ok=1, n_corrected=0} per frame
// K1
int launch_rgb_to_quant(const uint8_t* rgb, size_t n_px, t3c_pixel* out, cudaStream_t st);
int launch_quant_to_rgb(const t3c_pixel* px, size_t n_px, uint8_t* rgb, cudaStream_t st);
size_t launch_rgb_to_quant8(const uint8_t* rgb, size_t n_px, t3c_pixel* out, cudaStream_t st);   // k_fast.cu: pixels covered (0: unaligned)
size_t launch_quant_to_rgb8(const t3c_pixel* px, size_t n_px, uint8_t* rgb, cudaStream_t st);
int launch_pack_pixels(const t3c_pixel* px, size_t n_px, uint8_t* words9, cudaStream_t st);
int launch_unpack_pixels(const uint8_t* words9, size_t n_words, t3c_pixel* px, cudaStream_t st);
int launch_mod27(const uint8_t* in, size_t n, uint8_t* out, cudaStream_t st);
// block codecs
int launch_rs_encode_blocks(const DevTables& T, int k, int arith, const uint8_t* data, size_t n, uint8_t* out26, cudaStream_t st);
int launch_rs_decode_blocks(const DevTables& T, int k, int arith, uint8_t* inout26, size_t n, uint8_t* out_k, uint8_t* ok, cudaStream_t st);
int launch_perm2d(const uint8_t* in, uint8_t* out, size_t n, uint32_t w, uint32_t h, cudaStream_t st);
int launch_header_emit(const DevTables& T, const t3c_config& cfg, int arith, uint8_t* d_hdr27, uint8_t* d_coded52, cudaStream_t st);
int launch_header_parse(const DevTables& T, int arith, const uint8_t* d_words9, size_t n_words, t3c_config* d_cfg, int* d_ok, cudaStream_t st);
// L1 names of the reference's public surface (OLD:81-94, 176-205, 208-379), single items
int launch_header_pack(const t3c_config& cfg, uint32_t magic, uint32_t version, uint32_t hash, uint32_t seq, uint8_t* d_hdr27, cudaStream_t st);
int launch_header_check_unpack(const uint8_t* d_sym27, t3c_config* d_cfg, uint32_t* d_out4, int* d_ok, cudaStream_t st);
int launch_crc3_rem12(const uint8_t* d_trits, size_t n, uint8_t* d_out12, cudaStream_t st);
int launch_scramble(const DevTables& T, uint8_t* d_syms, size_t n, const uint8_t st8[8], int inverse, cudaStream_t st);
// general profile codec (any config)
int launch_encode_general(const DevTables& T, const t3c_config& cfg, const Geom& g, const uint8_t* raw9, uint8_t* out9, cudaStream_t st, uint64_t cw_start = 0);
// scratch_sy is band-major: decoded data symbol m of band b at b*pitch + m (pitch >= ceil(n_s / 9))
int launch_decode_fixed_general(const DevTables& T, const Geom& g, const uint8_t* in9, uint8_t* scratch_sy, uint64_t pitch, uint32_t* d_status, cudaStream_t st,
                                uint64_t cw_start = 0);
// pitch = 0: sy in stream order; pitch != 0: band-major scratch of launch_decode_fixed_general
int launch_regroup_words(const uint8_t* sy, uint64_t n_sy, uint64_t tile_area, uint32_t tile_w, uint8_t* out9, size_t n_words, cudaStream_t st, size_t w_start = 0,
                         uint64_t pitch = 0);
int launch_regroup_rgb(const uint8_t* sy, uint64_t n_sy, uint64_t tile_area, uint32_t tile_w, uint8_t* rgb, size_t n_px, cudaStream_t st, uint64_t pitch = 0,
                       size_t p_start = 0);
int launch_encode_general_from(const DevTables& T, const t3c_config& cfg, const Geom& g, const uint8_t* raw9, uint8_t* out9, cudaStream_t st, const CwStart& cs,
                               bool finish = true);
int launch_frame_misc_sparse(const DevTables& T, const t3c_config& cfg, const Geom& g, uint8_t* out9, size_t n_frames, size_t stride_bytes, cudaStream_t st,
                             uint32_t n_tiles, const uint32_t ncw_tile[9]);
int launch_decode_fixed_general_from(const DevTables& T, const Geom& g, const uint8_t* in9, uint8_t* scratch_sy, uint64_t pitch, uint32_t* d_status, cudaStream_t st,
                                     const CwStart& cs);
// super-tile kernels (k_super.cuh): per-band k, 2D tiles whose width divides 26, beacon periods 3..255.  They code the full
// super-tiles of every frame and report what is left in *tail; 0 = not applicable (nothing launched, *tail = everything)
bool super_path_ok(const t3c_config& cfg);
int super_debug_counters(uint32_t* out32);
int super_plan_describe(const t3c_config& cfg, size_t n_words, int decode, int words, uint32_t* out16, uint16_t* map, uint8_t* kv);
int launch_encode_super(const DevTables& T, const t3c_config& cfg, const Geom& g, const uint8_t* in, size_t in_pitch, bool words, size_t n_px,
                        size_t n_frames, uint8_t* out9, size_t stride_words, cudaStream_t st, SuperTail* tail);
int launch_decode_super(const DevTables& T, const t3c_config& cfg, const Geom& g, const uint8_t* in9, size_t stride_words, size_t n_frames, uint8_t* out,
                        size_t out_pitch, bool words, size_t n_px_out, uint32_t* d_status, cudaStream_t st, SuperTail* tail);
int launch_decode_ref_general(const DevTables& T, const RefDecGeom& g, const uint8_t* in9, uint8_t* use, uint32_t* d_status, cudaStream_t st);
// fused fast path (uniform k, 1D, no beacon): frames batched
bool fast_path_ok(const t3c_config& cfg);
int launch_encode_rgb_fast(const DevTables& T, const t3c_config& cfg, const Geom& g, const uint8_t* rgb, size_t n_px, size_t n_frames,
                           uint8_t* out9, size_t stride_words, cudaStream_t st);
// chk_*: bit planes of sum_i T_i[13*st_i] for the 6 scrambler phases and for the codeword at body index 0
int launch_decode_rgb_fast(const DevTables& T, const t3c_config& cfg, const Geom& g, const uint8_t* in9, size_t stride_words,
                           size_t n_frames, size_t n_px, size_t n_px_out, uint8_t* rgb, uint32_t* d_status, cudaStream_t st,
                           const uint32_t* chk_nz, const uint32_t* chk_two);
// the same for tiles [t0, t1) of the full mini-tiles only (chunked host pipelines); `tail` adds the ragged last tiles
// (and, for encode, header / padding).  in_pitch / out_pitch = bytes between RGB frames on the device.
int launch_encode_rgb_fast_part(const DevTables& T, const t3c_config& cfg, const Geom& g, const uint8_t* rgb, size_t n_px, size_t in_pitch,
                                size_t n_frames, uint8_t* out9, size_t stride_words, cudaStream_t st, uint32_t t0, uint32_t t1, bool tail);
int launch_decode_rgb_fast_part(const DevTables& T, const Geom& g, const uint8_t* in9, size_t stride_words, size_t n_frames, size_t n_px,
                                size_t out_pitch, size_t n_px_out, uint8_t* rgb, uint32_t* d_status, cudaStream_t st, const uint32_t* chk_nz,
                                const uint32_t* chk_two, uint32_t t0, uint32_t t1, bool tail);
// raw-word variants of the tiled kernels for one super-frame: full mini-tiles [0, *n_full) only; -1 = not applicable
int launch_encode_words_fast(const DevTables& T, const Geom& g, const uint8_t* raw9, uint8_t* out9, cudaStream_t st, uint32_t* n_full);
int launch_decode_words_fast(const DevTables& T, const Geom& g, const uint8_t* in9, uint8_t* raw9, size_t n_words_out, uint32_t* d_status, cudaStream_t st,
                             uint32_t* n_full);
uint32_t fast_full_tiles_encode(const Geom& g, size_t n_px);
uint32_t fast_full_tiles_decode(const Geom& g, size_t n_px_out, size_t out_pitch, size_t n_frames);
// both return -1 when the buffers are not 16-byte aligned (caller falls back to the general kernels)
// header + beacons + zero padding for n_frames super-frames laid out every stride_bytes
// coded header for cfg (device pointer, 52 symbols), re-emitted on `st` only when cfg/arith differ from the cached one
const uint8_t* cached_header(const DevTables& T, const t3c_config& cfg, int arith, cudaStream_t st, int& launches);
// header copy + zero padding for fast-path frames (no beacon) from the cached header
int launch_frame_finish(const DevTables& T, const t3c_config& cfg, const Geom& g, uint8_t* out9, size_t n_frames, size_t stride_bytes, cudaStream_t st);
int launch_frame_misc(const DevTables& T, const t3c_config& cfg, const Geom& g, uint8_t* out9, size_t n_frames, size_t stride_bytes, cudaStream_t st);

// SURVEY 8(f).2 / 8(f).3 (k_formats.cu)
int launch_subword_stream(const uint8_t* words9, size_t n_words, int N, uint8_t* trits, cudaStream_t st);
int launch_words_from_subword_stream(const uint8_t* trits, size_t n_trits, int N, uint8_t fill, uint8_t* words9, cudaStream_t st);
int launch_base243_pack(const uint8_t* trits, size_t n_trits, uint8_t* out, cudaStream_t st);           // writes the 4-byte count too
int launch_base243_unpack(const uint8_t* payload, size_t n_trits, uint8_t* trits, cudaStream_t st);
int launch_words_to_base243(const uint8_t* words9, size_t n_words, int N, uint8_t* out, cudaStream_t st);
int launch_v6new_pack_pixels(const t3c_pixel* px, size_t n_px, uint32_t* words, cudaStream_t st);
int launch_v6new_unpack_pixels(const uint32_t* words, size_t n_words, t3c_pixel* px, cudaStream_t st);

// image-bridge geometry (k_formats.cu)
int launch_resize_rgb_nn(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh, cudaStream_t st);
int launch_blit_center_rgb(const uint8_t* src, int sw, int sh, uint8_t* dst, int cw, int ch, cudaStream_t st);
int launch_extract_center_q(const t3c_pixel* full, int fw, int fh, int sw, int sh, t3c_pixel* sub, cudaStream_t st);
// .t3v frame records and CRC-32 (k_formats.cu); partial = scratch of t3v_partial_words(...) uint32
size_t t3v_partial_words(size_t n_words, size_t n_frames);
void build_crc_tables(uint32_t* h);   // host: crc_table_words() entries (slice-by-4 tables, shift tables, x^(8 2^i))
size_t crc_table_words();
int launch_t3v_records(const uint32_t* tabs, const uint8_t* words9, size_t n_words, size_t stride_words, size_t n_frames, uint8_t* records, size_t record_pitch, uint32_t* partial,
                       cudaStream_t st);
int launch_t3v_read(const uint32_t* tabs, const uint8_t* records, size_t record_pitch, size_t n_frames, size_t n_words, uint8_t* words9, size_t stride_words, uint32_t* partial,
                    uint8_t* ok, cudaStream_t st);
int launch_crc32(const uint32_t* tabs, const uint8_t* data, size_t n, uint32_t* partial, uint32_t* out, cudaStream_t st);

} // namespace t3c
