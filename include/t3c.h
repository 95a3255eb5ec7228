/*
 * t3c.h -- C ABI of libt3c.so: the B200-native (sm_100a) implementation of the Ternary Image
 * Codec v6 data-parallel encode/decode path.
 *
 * The reference has no FFI layer: the path is a header-only set of inline C++ functions
 * (old/include/ternary_image_codec_v6_min.hpp, "OLD").  Each entry point below replaces one of
 * them (cited as OLD:line / IMG:line = old/include/io_image.hpp:line); the C++ shim headers in
 * this directory (ternary_image_codec_v6_min.hpp, ternary_packing.hpp, io_image_bridge.hpp) keep
 * the reference's names and types and forward here.  See INTEGRATION.md.
 *
 * Conventions
 *   - Word27 = 9 bytes, symbol s at byte s (OLD:666-669); PixelYCbCrQuant = 6 bytes (OLD:670-674);
 *     RGB8 = 3 bytes interleaved, row-major.  No torch/C++ types cross this boundary.
 *   - Host-buffer calls copy in/out on the context's stream and return when the result is in
 *     the caller's buffer.  `_dev` calls take device pointers and a cudaStream_t (as void*), are
 *     asynchronous and never synchronise, except where a result count must be returned.
 *   - One t3c_ctx per device; calls on one ctx are serialised by the caller.
 *   - There is NO CPU fallback: without a CUDA device t3c_create fails with T3C_ERR_NODEVICE.
 *   - arith: T3C_REF_EXACT reproduces the reference bit for bit, bugs included;
 *     T3C_FIXED applies the 3-line RS repair and, for decode, inverts the
 *     encoder's framing.
 */
#ifndef T3C_H
#define T3C_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define T3C_API
#else
#define T3C_API __attribute__((visibility("default")))
#endif

typedef struct t3c_ctx t3c_ctx;

typedef enum {
    T3C_OK = 0,
    T3C_ERR_ARG = 1,        /* bad argument (null pointer, invalid k, ...) */
    T3C_ERR_CUDA = 2,       /* CUDA runtime error, see t3c_last_error */
    T3C_ERR_CAPACITY = 3,   /* output buffer too small */
    T3C_ERR_NODEVICE = 4,   /* no CUDA device / driver: there is no CPU fallback */
    T3C_ERR_UNSUPPORTED = 5
} t3c_status;

typedef enum { T3C_REF_EXACT = 0, T3C_FIXED = 1 } t3c_arith;

#define T3C_PROFILE_RAW 0xFF /* ProfileID::RAW_MODE, OLD:34 */

/* EncoderConfig (OLD:862-873) / DecoderConfigSeen (OLD:874-884) as plain integers. */
typedef struct {
    uint8_t  profile;           /* ProfileID 0..4 = P1..P5, 0xFF = RAW_MODE */
    uint8_t  uep[9];            /* UEPLayout::band_profile; k = {24,22,20,18}[uep%4]  (OLD:1089-1100) */
    uint16_t tile_w, tile_h;    /* Tile2D; used only when profile==P5 and both non-zero (OLD:1083) */
    uint32_t seed_a, seed_b, seed_s0; /* ScramblerSeed */
    uint32_t beacon_period;     /* SparseBeaconCfg::words_period */
    uint8_t  beacon_slot;
    uint8_t  beacon_enabled;
    uint8_t  subword;           /* SubwordMode 27/24/21/18/15 (header slot 12 only) */
    uint8_t  centered;
    uint8_t  coset;             /* CosetID 0..2 (header slot 16 only) */
    uint8_t  pad_[3];
    uint32_t superframe_words;  /* only %5 reaches the beacon payload (OLD:1130) */
} t3c_config;

typedef struct { uint16_t Yq; int16_t Cbq, Crq; } t3c_pixel; /* PixelYCbCrQuant */

/* ---- context ---------------------------------------------------------------------------- */
T3C_API t3c_status  t3c_create(int device, t3c_ctx** out);      /* EncoderContext()/DecoderContext() ctor work, OLD:885-916 */
T3C_API void        t3c_destroy(t3c_ctx* ctx);
T3C_API const char* t3c_last_error(const t3c_ctx* ctx);
/* Host-buffer calls copy through PCIe in chunks that overlap with the kernels; with PAGEABLE buffers every copy is staged by the driver
 * and blocks (an 8K encode + decode drops from 7.9 to 43.8 ms).  mode = 1: a pageable buffer of 8 MiB or more that is passed a second
 * time (same address and size) is page-locked in place (cudaHostRegister, ~170 us per MB once) and stays so until it falls out of a
 * 16-entry LRU or the context is destroyed; do not free such a buffer while a call that uses it is running.  mode = 0 (default): never.
 * The C++ drop-in headers switch it on for their shared context (std::vector storage is pageable). */
T3C_API t3c_status  t3c_set_host_registration(t3c_ctx* ctx, int mode);
T3C_API int         t3c_version(void);
T3C_API void        t3c_config_default(t3c_config* cfg);        /* EncoderConfig defaults + uep_uniform(1), OLD:862-873,898 */
T3C_API void*       t3c_stream(t3c_ctx* ctx);                   /* the context's cudaStream_t */
T3C_API t3c_status  t3c_sync(t3c_ctx* ctx);
T3C_API uint64_t    t3c_kernel_launches(const t3c_ctx* ctx);    /* kernels launched by this ctx so far */

/* ---- sizes ------------------------------------------------------------------------------ */
/* exact number of profile words encode_profile_from_raw emits for n_raw_words */
T3C_API size_t t3c_profile_words(const t3c_config* cfg, size_t n_raw_words);

/* ---- K1: RGB8 <-> quant <-> Word27 -------------------------------------------------------- */
/* rgb_to_quant_stream, IMG:156-170 (rgb_to_ycbcr IMG:47-56 + quantize_ycbcr IMG:69-78) */
T3C_API t3c_status t3c_rgb_to_quant(t3c_ctx*, const uint8_t* rgb, size_t n_px, t3c_pixel* out);
/* quant_stream_to_rgb, IMG:171-192 */
T3C_API t3c_status t3c_quant_to_rgb(t3c_ctx*, const t3c_pixel* px, size_t n_px, uint8_t* rgb);
/* encode_raw_pixels_to_words, OLD:723-734: writes ceil(n_px/2) words */
T3C_API t3c_status t3c_pack_pixels(t3c_ctx*, const t3c_pixel* px, size_t n_px, uint8_t* words9, size_t* n_words);
/* decode_raw_words_to_pixels, OLD:735-747: writes 2*n_words pixels */
T3C_API t3c_status t3c_unpack_pixels(t3c_ctx*, const uint8_t* words9, size_t n_words, t3c_pixel* px);
/* tpack::words_to_bytes / bytes_to_words, include/ternary_packing.hpp:53-65 (each symbol %27) */
T3C_API t3c_status t3c_words_to_bytes(t3c_ctx*, const uint8_t* words9, size_t n_words, uint8_t* bytes);

/* ---- stage-level block codecs (parity hooks for OLD:490-663, 749-813, 155-380) ------------- */
/* RSCodec::encode_block over n_blocks blocks of k symbols -> 26 symbols each, OLD:517-535 */
T3C_API t3c_status t3c_rs_encode_blocks(t3c_ctx*, int k, int arith, const uint8_t* data, size_t n_blocks, uint8_t* out26);
/* RSCodec::decode_block, OLD:546-662: inout26 corrected in place, out_k = first k symbols (zeros when !ok) */
T3C_API t3c_status t3c_rs_decode_blocks(t3c_ctx*, int k, int arith, uint8_t* inout26, size_t n_blocks, uint8_t* out_k, uint8_t* ok);
/* interleave2D_boustrophedon / deinterleave2D_boustrophedon, OLD:750-813 (in place) */
T3C_API t3c_status t3c_interleave2d(t3c_ctx*, uint8_t* syms, size_t n, uint16_t w, uint16_t h, int inverse);
/* HeaderCodec::pack + 2x RS(26,18): the 52 coded header symbols, OLD:1142-1158 (device kernel) */
T3C_API t3c_status t3c_header_emit(t3c_ctx*, const t3c_config*, int arith, uint8_t hdr27[27], uint8_t coded52[52]);
/* read_and_decode_header_from_words, OLD:918-937: *ok = RS+CRC verdict */
T3C_API t3c_status t3c_header_parse(t3c_ctx*, int arith, const uint8_t* words9, size_t n_words, t3c_config* out, int* ok);

/* ---- L0 / L1 names of the reference's public surface, single items (for drop-in callers and tests; bulk work goes through
 * the profile codec).  SuperframeHeader (OLD:155-171) = the config fields plus magic, version, band_map_hash, frame_seq. */
typedef struct {
    uint16_t   magic;            /* 0x0A2 */
    uint8_t    version;          /* 1 */
    uint8_t    pad_;
    uint32_t   band_map_hash;
    uint32_t   frame_seq;
    t3c_config cfg;              /* profile, uep, tile, seed, beacon, subword, centered, coset (superframe_words unused) */
} t3c_header;
/* HeaderCodec::pack, OLD:208-289: 27 symbols with the ternary CRC-12 in slots 20, 21, 22, 26 */
T3C_API t3c_status t3c_header_pack(t3c_ctx*, const t3c_header*, uint8_t sym27[27]);
/* HeaderCodec::check, OLD:290-319 */
T3C_API t3c_status t3c_header_check(t3c_ctx*, const uint8_t sym27[27], int* ok);
/* HeaderCodec::unpack, OLD:320-379 (profile % 5, UEP triples LSB-first, beacon slot % 9: the reference's reading, bug B7 included) */
T3C_API t3c_status t3c_header_unpack(t3c_ctx*, const uint8_t sym27[27], t3c_header* out);
/* CRC3::rem12, OLD:176-205: n message trits, then twelve zero trits */
T3C_API t3c_status t3c_crc3_rem12(t3c_ctx*, const uint8_t* trits, size_t n, uint8_t out12[12]);
/* scramble_symbol / descramble_symbol (OLD:81-94) applied to n symbols in sequence, in place: *st is the running LCG state, read
 * (as is, like the reference's uint32_t& st) and left at the state the last symbol saw */
T3C_API t3c_status t3c_scramble_symbols(t3c_ctx*, uint8_t* syms, size_t n, uint32_t a, uint32_t b, uint32_t* st, int inverse);
/* encode_beacon_symbol, OLD:107-113 */
T3C_API t3c_status t3c_beacon_symbol(t3c_ctx*, int profile, uint32_t frame_seq_mod, uint32_t health_flags, uint8_t* sym);
/* GF27Tables as GF27Context::init builds them (OLD:414-466) plus the digit-wise add / sub of OLD:383-401 as tables */
typedef struct {
    uint8_t exp[78];
    int16_t log[27];
    uint8_t mul[729];
    uint8_t inv[27];
    uint8_t primitive;
    uint8_t add[729];
    uint8_t sub[729];
} t3c_gf27;
T3C_API t3c_status t3c_gf27_tables(t3c_ctx*, t3c_gf27* out);

/* ---- K2..K5: profile codec ----------------------------------------------------------------- */
/* encode_profile_from_raw, OLD:1043-1169 */
T3C_API t3c_status t3c_encode_profile(t3c_ctx*, const t3c_config*, int arith, const uint8_t* raw9, size_t n_words,
                                      uint8_t* out9, size_t cap_words, size_t* n_out);
/* decode_profile_to_raw AS SHIPPED, OLD:995-1041: `seen` is DecoderContext::cfg_last_seen (read, then
 * overwritten from the header); *ok is the reference's bool; on !ok *n_out = 0. */
T3C_API t3c_status t3c_decode_profile(t3c_ctx*, t3c_config* seen, const uint8_t* in9, size_t n_words,
                                      uint8_t* out9, size_t cap_words, size_t* n_out, int* ok);
/* consistent decoder for T3C_FIXED streams.  cfg = the encoder's config (carried out of
 * band: the 27-symbol header cannot hold tiles >= 27, UEP index 3 or periods > 26, bug B7);
 * n_raw_words = N_w given to the encoder (0 = infer from n_words; not possible with the 2D interleave).
 * *n_out = recovered prefix of raw words (the encoder drops < k symbols per band, bug B8). */
T3C_API t3c_status t3c_decode_profile_fixed(t3c_ctx*, const t3c_config* cfg, size_t n_raw_words, const uint8_t* in9,
                                            size_t n_words, uint8_t* out9, size_t cap_words, size_t* n_out, int* ok,
                                            size_t* n_corrected);

/* ---- fused, batched frames (old/src/main.cpp:15-26 as one call) ----------------------------- */
/* n_frames RGB8 frames of n_px pixels each -> n_frames profile-word streams of *words_per_frame words,
 * frame f at out9 + f*stride_words*9.  Equal to rgb_to_quant_stream -> encode_raw_pixels_to_words ->
 * encode_profile_from_raw per frame. */
T3C_API t3c_status t3c_encode_frames_rgb8(t3c_ctx*, const t3c_config*, int arith, const uint8_t* rgb, size_t n_px,
                                          size_t n_frames, uint8_t* out9, size_t stride_words, size_t* words_per_frame);
/* inverse (T3C_FIXED framing): per frame decode -> decode_raw_words_to_pixels -> quant_stream_to_rgb.
 * ok[f] per frame; pixels past the recovered prefix are left untouched; *px_recovered per frame; *n_corrected summed over the frames.
 * Any n_frames (batches above 32 frames are decoded in pieces of 32 inside the call).  Limits of the whole ABI: a frame's body below 2^32
 * symbols on the tiled paths (an 8K frame has 1.9e8), .t3v record counts are uint32 (T3C_ERR_ARG beyond), sub-word N in 1..27. */
T3C_API t3c_status t3c_decode_frames_rgb8(t3c_ctx*, const t3c_config*, const uint8_t* in9, size_t words_per_frame,
                                          size_t stride_words, size_t n_frames, size_t n_px, uint8_t* rgb,
                                          uint8_t* ok, size_t* px_recovered, size_t* n_corrected);

/* ---- device-pointer variants (bench / pipelines: PCIe outside the timed region) ------------- */
/* The _dev calls enqueue on `stream` and return; they never wait for the device EXCEPT when they have to change context-wide state: the
 * first use of a config (the coded header / the pass maps of the super-tile kernels are built once and cached: one stream or device
 * synchronisation) and the growth of the context's scratch buffers.  They are therefore not safe under stream capture.  Scratch is per
 * context: a call on another stream than the previous one is ordered after it by an event. */
T3C_API t3c_status t3c_rgb_to_quant_dev(t3c_ctx*, const uint8_t* d_rgb, size_t n_px, t3c_pixel* d_out, void* stream);
T3C_API t3c_status t3c_quant_to_rgb_dev(t3c_ctx*, const t3c_pixel* d_px, size_t n_px, uint8_t* d_rgb, void* stream);
T3C_API t3c_status t3c_pack_pixels_dev(t3c_ctx*, const t3c_pixel* d_px, size_t n_px, uint8_t* d_words9, void* stream);
T3C_API t3c_status t3c_unpack_pixels_dev(t3c_ctx*, const uint8_t* d_words9, size_t n_words, t3c_pixel* d_px, void* stream);
T3C_API t3c_status t3c_rs_encode_blocks_dev(t3c_ctx*, int k, int arith, const uint8_t* d_data, size_t n_blocks, uint8_t* d_out26, void* stream);
T3C_API t3c_status t3c_rs_decode_blocks_dev(t3c_ctx*, int k, int arith, uint8_t* d_inout26, size_t n_blocks, uint8_t* d_out_k, uint8_t* d_ok, void* stream);
T3C_API t3c_status t3c_encode_profile_dev(t3c_ctx*, const t3c_config*, int arith, const uint8_t* d_raw9, size_t n_words,
                                          uint8_t* d_out9, size_t cap_words, void* stream);
/* d_status[0] = ok flag (1/0), d_status[1] = symbols corrected; both written asynchronously */
T3C_API t3c_status t3c_decode_profile_fixed_dev(t3c_ctx*, const t3c_config*, size_t n_raw_words, const uint8_t* d_in9,
                                                size_t n_words, uint8_t* d_out9, size_t cap_words, uint32_t* d_status, void* stream);
T3C_API t3c_status t3c_encode_frames_rgb8_dev(t3c_ctx*, const t3c_config*, int arith, const uint8_t* d_rgb, size_t n_px,
                                              size_t n_frames, uint8_t* d_out9, size_t stride_words, void* stream);
/* d_status: 2 uint32 per frame {ok, n_corrected} */
T3C_API t3c_status t3c_decode_frames_rgb8_dev(t3c_ctx*, const t3c_config*, const uint8_t* d_in9, size_t words_per_frame,
                                              size_t stride_words, size_t n_frames, size_t n_px, uint8_t* d_rgb,
                                              uint32_t* d_status, void* stream);
/* which kernel family the fused calls will use for this config: 1 = tiled fast path, 0 = general path */
T3C_API int t3c_fast_path_available(const t3c_config* cfg);
/* 1 = the super-tile kernels take this config (any per-band k, 2D tile widths dividing 26, beacon periods 3..255); they run when
 * t3c_fast_path_available is 0, the general kernels otherwise keep only the ragged end of a frame */
T3C_API int t3c_super_path_available(const t3c_config* cfg);
/* host-only (no device needed): the plan of the super-tile kernels for one super-frame of n_raw_words: out16 = {M band symbols per
 * super-tile, units (6 pixels) per super-tile, full super-tiles, number of distinct k, k[4], codewords per band and super-tile [4],
 * phase-B passes per super-tile index mod 3 [3], shared memory bytes}; map (3 x 64 x 32 entries band | codeword << 4, 0xFFFF = idle lane)
 * and pass_kv (3 x 64: k slot | scrambler variant << 2) may be NULL.  Returns 0 when the super-tile kernels do not take the config. */
T3C_API int t3c_super_plan_describe(const t3c_config* cfg, size_t n_raw_words, int decode, int words, uint32_t* out16, uint16_t* map, uint8_t* pass_kv);
/* development aid: per-phase cycle counters of the super-tile kernels (32 values, read and reset); 0 unless built with -DT3C_SUPER_DEBUG */
T3C_API int t3c_debug_counters(uint32_t* out32);

/* ---- SURVEY 8(f) "next" rows: the data formats either side of the path ------------------------ */
/* 8(f).2 sub-word streams (OLD:816-859) and base-243 packing (include/ternary_packing.hpp:18-50), N in 1..27.
 * extract_subword_stream_from_words, OLD:835-845: the first N trits of every word, one trit (0..2) per byte */
T3C_API t3c_status t3c_subword_stream(t3c_ctx*, const uint8_t* words9, size_t n_words, int N, uint8_t* trits);
/* build_words_from_subword_stream, OLD:846-859: ceil(n_trits/N) words, trits N..26 of each word = fill */
T3C_API t3c_status t3c_words_from_subword_stream(t3c_ctx*, const uint8_t* trits, size_t n_trits, int N, uint8_t fill, uint8_t* words9, size_t* n_words);
/* tpack::ut_to_base243, include/ternary_packing.hpp:29-39: uint32 LE trit count, then 5 trits per byte: 4 + ceil(n/5) bytes */
T3C_API t3c_status t3c_base243_pack(t3c_ctx*, const uint8_t* trits, size_t n_trits, uint8_t* out, size_t* n_bytes);
/* tpack::base243_to_ut, include/ternary_packing.hpp:41-50: *ok = the reference's bool; writes min(count, cap) trits */
T3C_API t3c_status t3c_base243_unpack(t3c_ctx*, const uint8_t* in, size_t n_bytes, uint8_t* trits, size_t cap, size_t* n_trits, int* ok);
/* fused extract_subword_stream_from_words + ut_to_base243 (what old/include/t3p_io.hpp:19 writes), no 1-byte-per-trit stream */
T3C_API t3c_status t3c_words_to_base243(t3c_ctx*, const uint8_t* words9, size_t n_words, int N, uint8_t* out, size_t* n_bytes);
/* 8(f).3 NEW-generation RAW path, src/ternary_image_codec_v6_min.cpp:62-126: one pixel <-> one 32-bit word
 * Y + 243 (Cb+40 + 81 (Cr+40)) with clamps; subword = 0 or a SubwordMode value (27/24/21/18/15), anything else
 * makes the reference's *_subword variants return false: T3C_ERR_ARG here */
T3C_API t3c_status t3c_v6new_pack_pixels(t3c_ctx*, const t3c_pixel* px, size_t n_px, uint32_t* words, int subword);
T3C_API t3c_status t3c_v6new_unpack_pixels(t3c_ctx*, const uint32_t* words, size_t n_words, t3c_pixel* px, int subword);
T3C_API t3c_status t3c_subword_stream_dev(t3c_ctx*, const uint8_t* d_words9, size_t n_words, int N, uint8_t* d_trits, void* stream);
T3C_API t3c_status t3c_words_from_subword_stream_dev(t3c_ctx*, const uint8_t* d_trits, size_t n_trits, int N, uint8_t fill, uint8_t* d_words9, void* stream);
T3C_API t3c_status t3c_base243_pack_dev(t3c_ctx*, const uint8_t* d_trits, size_t n_trits, uint8_t* d_out, void* stream);
T3C_API t3c_status t3c_base243_unpack_dev(t3c_ctx*, const uint8_t* d_payload /* after the 4-byte count */, size_t n_trits, uint8_t* d_trits, void* stream);
T3C_API t3c_status t3c_words_to_base243_dev(t3c_ctx*, const uint8_t* d_words9, size_t n_words, int N, uint8_t* d_out, void* stream);
T3C_API t3c_status t3c_v6new_pack_pixels_dev(t3c_ctx*, const t3c_pixel* d_px, size_t n_px, uint32_t* d_words, void* stream);
T3C_API t3c_status t3c_v6new_unpack_pixels_dev(t3c_ctx*, const uint32_t* d_words, size_t n_words, t3c_pixel* d_px, void* stream);

/* 8(f).1 the .t3v container's records (old/include/t3v_io.hpp).  A frame record is n (uint32 LE) | 9n symbol bytes, each % 27 |
 * crc32(payload) ^ (crc32(&n, 4) * 16777619) (t3v_write_frame, :128-142); CRC-32 is the reflected 0xEDB88320 one (:14-40). */
T3C_API t3c_status t3c_crc32(t3c_ctx*, const uint8_t* data, size_t n, uint32_t* crc);
T3C_API t3c_status t3c_t3v_frame_record(t3c_ctx*, const uint8_t* words9, size_t n_words, uint8_t* record /* 8 + 9n bytes */, size_t* n_bytes);
/* t3v_read_frame, :143-160: *ok = 0 when the record is short or its CRC does not match (nothing written); symbols are returned as stored */
T3C_API t3c_status t3c_t3v_read_frame(t3c_ctx*, const uint8_t* record, size_t n_bytes, uint8_t* words9, size_t cap_words, size_t* n_words, int* ok);
/* t3v_write_header, :97-119: the 54-byte packed T3VHeaderBin, its last field the CRC-32 of the 50 bytes before it; aw = {x0, y0, w, h} */
T3C_API t3c_status t3c_t3v_header(t3c_ctx*, uint8_t out54[54], int profile, int subword_code, int centered, int coset, uint32_t width, uint32_t height,
                                  const uint32_t aw[4], uint32_t fps_num, uint32_t fps_den, uint32_t frame_count, int file_type);
/* the .t3vi sidecar (old/include/t3v_indexed_io.hpp:14-44) of n_frames records that lie back to back from byte first_offset of a .t3v
 * file, frame i holding n_words[i] words: 17-byte header ("T3VI", 1, n_frames, 0, CRC-32 of those 13 bytes) + n_frames uint64 offsets */
T3C_API t3c_status t3c_t3v_index_build(t3c_ctx*, const uint64_t* n_words, size_t n_frames, uint64_t first_offset, uint8_t* out /* 17 + 8 n */, size_t* n_bytes);
/* batched, device-resident: frame f of n_words words at d_words9 + f * 9 * stride_words <-> record f at d_records + f * record_pitch; all frame
 * starts 4-byte aligned, record_pitch >= 8 + 9 n_words.  d_ok[f] = the record announces n_words and carries the right CRC.
 * A record is n | payload | CRC, so its payload starts 4 bytes in: with d_records at 12 mod 16 (and record_pitch, 9 * stride_words multiples
 * of 16) both sides of the copy move in 16-byte accesses; any 4-byte aligned placement is correct, only slower (4-byte accesses on one side) */
T3C_API t3c_status t3c_t3v_frame_records_dev(t3c_ctx*, const uint8_t* d_words9, size_t n_words, size_t stride_words, size_t n_frames, uint8_t* d_records,
                                             size_t record_pitch, void* stream);
T3C_API t3c_status t3c_t3v_read_frames_dev(t3c_ctx*, const uint8_t* d_records, size_t record_pitch, size_t n_frames, size_t n_words, uint8_t* d_words9,
                                           size_t stride_words, uint8_t* d_ok, void* stream);

/* 8(f).4 image-bridge geometry of the NEW generation (include/io_image.hpp:102-140, 215-235) and the pipelines built from it (:238-338).
 * RGB8 images are w*h*3 bytes, row-major.  resize: nearest neighbour with the reference's double-precision index; an empty source leaves
 * the destination black.  blit: black canvas with src centred (needs src_w <= canvas_w; rows below the canvas are dropped).
 * extract: the centred sub_w x sub_h window of a quantised frame (needs sub_w <= full_w; rows below the frame are zero). */
T3C_API t3c_status t3c_resize_rgb_nn(t3c_ctx*, const uint8_t* src, int src_w, int src_h, uint8_t* dst, int dst_w, int dst_h);
T3C_API t3c_status t3c_blit_center_rgb(t3c_ctx*, const uint8_t* src, int src_w, int src_h, uint8_t* canvas, int canvas_w, int canvas_h);
T3C_API t3c_status t3c_extract_center_q(t3c_ctx*, const t3c_pixel* full, int full_w, int full_h, int sub_w, int sub_h, t3c_pixel* sub);
T3C_API t3c_status t3c_resize_rgb_nn_dev(t3c_ctx*, const uint8_t* d_src, int src_w, int src_h, uint8_t* d_dst, int dst_w, int dst_h, void* stream);
T3C_API t3c_status t3c_blit_center_rgb_dev(t3c_ctx*, const uint8_t* d_src, int src_w, int src_h, uint8_t* d_canvas, int canvas_w, int canvas_h, void* stream);
T3C_API t3c_status t3c_extract_center_q_dev(t3c_ctx*, const t3c_pixel* d_full, int full_w, int full_h, int sub_w, int sub_h, t3c_pixel* d_sub, void* stream);
/* image_to_words_subword after the file load, :238-301: resize to std_res_for(subword) (NEW: 27 -> 7680x4320, 24 -> 3840x2160, 21 -> 1920x1080,
 * 18 -> 1280x720, 15 -> 960x540), centred in the 7680x4320 canvas when centered && subword != 27, quantise, one 32-bit word per pixel.
 * *ok = the reference's bool (0 for an invalid subword or an empty image). */
T3C_API t3c_status t3c_v6new_image_to_words(t3c_ctx*, const uint8_t* rgb, int w, int h, int subword, int centered, uint32_t* words, size_t cap_words,
                                            size_t* n_words, int* ok);
/* words_to_image_subword before the file write, :304-338: words -> pixels; as many as w*h: that image; a full 7680x4320 canvas (subword != 27):
 * its centre window of std_res_for(subword), poured row-major into w x h; anything else: poured into w x h as far as it goes (the rest black) */
T3C_API t3c_status t3c_v6new_words_to_image(t3c_ctx*, const uint32_t* words, size_t n_words, int subword, int w, int h, uint8_t* rgb, int* ok);

/* ---- multi-device streams (old/src/main_video_t3v.cpp:19-26 for a whole stream; BASELINE config 4) ---------------------------------
 * A stream owns one context per listed device (a device may be listed more than once: several lanes on one GPU).  Frame f of a call
 * is coded on lane (first_frame + f) % n_lanes by that lane's own host thread through the chunked host-buffer pipeline; frames are
 * independent (one super-frame per frame, no state between them), every lane writes its frames at their place in the caller's
 * output, so the result is in frame order when the call returns.  No collective, no device-to-device traffic. */
typedef struct t3c_streamset t3c_streamset;
T3C_API t3c_status t3c_stream_create(const int* devices, int n_lanes, t3c_streamset** out);
T3C_API void       t3c_stream_destroy(t3c_streamset*);
T3C_API int        t3c_stream_lanes(const t3c_streamset*);
/* n_frames RGB8 frames of n_px pixels, contiguous -> profile words, frame f at out + 9 * stride_words * f */
T3C_API t3c_status t3c_stream_encode_rgb8(t3c_streamset*, const t3c_config*, int arith, const uint8_t* rgb, size_t n_px, size_t n_frames, size_t first_frame,
                                          uint8_t* out9, size_t stride_words, size_t* words_per_frame);
/* the inverse (consistent decoder); ok[f] per frame, *n_corrected summed over the frames */
T3C_API t3c_status t3c_stream_decode_rgb8(t3c_streamset*, const t3c_config*, const uint8_t* in9, size_t words_per_frame, size_t stride_words, size_t n_frames,
                                          size_t first_frame, size_t n_px, uint8_t* rgb, uint8_t* ok, size_t* n_corrected);

#ifdef __cplusplus
}
#endif
#endif /* T3C_H */
