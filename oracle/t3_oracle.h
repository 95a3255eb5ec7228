/*
 * t3_oracle.h -- CPU oracle for the ternary image codec hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * algorithm (old/include/ternary_image_codec_v6_min.hpp, "OLD" below, and the
 * per-pixel bridge in old/include/io_image.hpp).  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load it.
 * The product (libt3c.so) never links or calls anything in oracle/.
 *
 * Parity pinning: the reference holds no golden vectors;
 * this restatement is pinned differentially against the reference itself,
 * compiled from /root/reference into oracle/_ref/ (see oracle/Makefile), by
 * tests/test_oracle_vs_reference.py, and against fixtures generated from that
 * reference build (tests/golden/, generators tests/golden/make_golden.py, make_golden_formats.py).
 *
 * Two arithmetic modes:
 *   fixed=0  REF-EXACT: the reference as shipped, bugs included.
 *   fixed=1  FIXED: RS arithmetic repaired by the 3-line patch of SURVEY
 *            Appendix B, plus a decoder that actually inverts the encoder
 *            (t3o_decode_profile_fixed).
 */
#ifndef T3_ORACLE_H
#define T3_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define T3O_PROFILE_RAW 0xFF

/* Mirrors EncoderConfig (OLD:862-873) field for field, as plain integers. */
typedef struct {
    uint8_t  profile;          /* ProfileID: 0..4, 0xFF = RAW_MODE (OLD:34) */
    uint8_t  uep[9];           /* UEPLayout::band_profile (OLD:60-63) */
    uint16_t tile_w, tile_h;   /* Tile2D (OLD:73-76) */
    uint32_t seed_a, seed_b, seed_s0; /* ScramblerSeed (OLD:77-80) */
    uint32_t beacon_period;    /* SparseBeaconCfg (OLD:95-100) */
    uint8_t  beacon_slot;
    uint8_t  beacon_enabled;
    uint8_t  subword;          /* SubwordMode: 27,24,21,18,15 (OLD:117) */
    uint8_t  centered;
    uint8_t  coset;            /* CosetID 0..2 (OLD:114) */
    uint8_t  pad_[3];
    uint32_t superframe_words; /* OLD:869 */
} t3o_cfg;

/* PixelYCbCrQuant (OLD:670-674): 6 bytes. */
typedef struct { uint16_t Yq; int16_t Cbq, Crq; } t3o_pixel;

void t3o_cfg_default(t3o_cfg* c);       /* EncoderContext() defaults, OLD:862-873,898 */

/* GF(27), OLD:383-487 */
uint8_t t3o_gf_add(uint8_t a, uint8_t b);
uint8_t t3o_gf_sub(uint8_t a, uint8_t b);
uint8_t t3o_gf_mul(uint8_t a, uint8_t b);
uint8_t t3o_gf_inv(uint8_t a);
uint8_t t3o_gf_pow_alpha(int e);
int     t3o_gf_log(uint8_t a);

/* RS(26,k), OLD:490-663.  k in {24,22,20,18}. */
int  t3o_rs_gen(int k, uint8_t* g_out /* r+1 */);
void t3o_rs_encode(int k, int fixed, const uint8_t* data_k, uint8_t* out26);
int  t3o_rs_decode(int k, int fixed, uint8_t* inout26, uint8_t* out_k);
void t3o_rs_encode_blocks(int k, int fixed, const uint8_t* data, size_t nblk, uint8_t* out);
void t3o_rs_decode_blocks(int k, int fixed, uint8_t* inout, size_t nblk, uint8_t* out, uint8_t* ok);

/* Header + CRC, OLD:155-380. */
void t3o_crc12(const uint8_t* trits, size_t n, uint8_t out12[12]);
void t3o_header_pack(const t3o_cfg* c, uint32_t frame_seq, uint32_t band_map_hash, uint8_t sym27[27]);
int  t3o_header_check(const uint8_t sym27[27]);
void t3o_header_unpack(const uint8_t sym27[27], t3o_cfg* out, uint32_t* frame_seq, uint32_t* band_map_hash,
                       uint16_t* magic, uint8_t* version);

/* 2 px <-> Word27, OLD:665-747. words are 9 bytes each. */
size_t t3o_pack_pixels(const t3o_pixel* px, size_t n_px, uint8_t* words9);
void   t3o_unpack_pixels(const uint8_t* words9, size_t n_words, t3o_pixel* px);

/* 2D boustrophedon, OLD:749-813 (in place). */
void t3o_interleave2d(uint8_t* sy, size_t n, unsigned w, unsigned h);
void t3o_deinterleave2d(uint8_t* sy, size_t n, unsigned w, unsigned h);

/* Scrambler / beacon, OLD:77-113. */
uint8_t t3o_scramble_symbol(uint8_t s, uint32_t a, uint32_t b, uint32_t* st);
uint8_t t3o_descramble_symbol(uint8_t s, uint32_t a, uint32_t b, uint32_t* st);
uint8_t t3o_beacon_symbol(uint8_t profile, uint16_t frame_seq_mod, uint8_t health);

/* Profile codec, OLD:918-1169. */
size_t t3o_profile_words_bound(const t3o_cfg* c, size_t n_raw_words);  /* exact N_out */
size_t t3o_encode_profile(const t3o_cfg* c, int fixed, const uint8_t* raw9, size_t n_words,
                          uint8_t* out9, size_t cap_words);
/* Reference decoder as shipped (A.7). `seen` is DecoderContext::cfg_last_seen, read then mutated.
 * Returns the reference's bool; *n_out = words written (0 when false). */
int t3o_decode_profile_ref(t3o_cfg* seen, const uint8_t* in9, size_t n_words,
                           uint8_t* out9, size_t cap_words, size_t* n_out);
/* Consistent decoder for FIXED mode (A.8).  cfg = the encoder's true config,
 * n_raw_words = N_w the encoder was given (0 = infer; not allowed with 2D interleave).
 * Returns 1 ok / 0 failure; *n_out = recovered prefix of raw words; *n_corrected = symbols corrected. */
int t3o_decode_profile_fixed(const t3o_cfg* cfg, size_t n_raw_words, const uint8_t* in9, size_t n_words,
                             uint8_t* out9, size_t cap_words, size_t* n_out, size_t* n_corrected);

/* RGB8 <-> quant bridge, old/include/io_image.hpp:47-84,156-192. */
void t3o_rgb_to_quant(const uint8_t* rgb, size_t n_px, t3o_pixel* out);
void t3o_quant_to_rgb(const t3o_pixel* px, size_t n_px, uint8_t* rgb);

/* Fused conveniences used by bench / tests: RGB8 -> profile words and back. */
size_t t3o_encode_rgb(const t3o_cfg* c, int fixed, const uint8_t* rgb, size_t n_px, uint8_t* out9, size_t cap_words);
int    t3o_decode_rgb_fixed(const t3o_cfg* c, size_t n_px, const uint8_t* in9, size_t n_words, uint8_t* rgb,
                            size_t* n_px_out, size_t* n_corrected);

/* ---- sub-word streams + base-243 (OLD:816-859, include/ternary_packing.hpp:18-50) and the
 * NEW-generation RAW path (src/ternary_image_codec_v6_min.cpp:62-126: one pixel -> one 32-bit word, clamped) ---- */
/* ---- image-bridge geometry of the NEW generation (include/io_image.hpp:102-140, 215-235) */
void t3o_resize_rgb_nn(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh);
void t3o_blit_center_rgb(const uint8_t* src, int sw, int sh, uint8_t* dst, int cw, int ch);             /* needs sw <= cw */
void t3o_extract_center_q(const t3o_pixel* full, int fw, int fh, int sw, int sh, t3o_pixel* sub);       /* needs sw <= fw */
/* ---- .t3v container records (old/include/t3v_io.hpp) */
uint32_t t3o_crc32(const uint8_t* data, size_t n);                                            /* t3v_detail::crc32, :14-40 */
size_t   t3o_t3v_frame_record(const uint8_t* words9, uint32_t n_words, uint8_t* out);       /* t3v_write_frame, :128-142: 8 + 9n bytes */
int      t3o_t3v_read_frame(const uint8_t* rec, size_t n_bytes, uint8_t* words9, uint32_t* n_words); /* t3v_read_frame, :143-160 */
void     t3o_t3v_header(uint8_t out54[54], int profile, int subword_code, int centered, int coset, uint32_t w, uint32_t h, const uint32_t aw[4],
                        uint32_t fps_num, uint32_t fps_den, uint32_t frame_count, int file_type); /* t3v_write_header, :97-119 */
void   t3o_subword_stream(const uint8_t* words9, size_t n_words, int N, uint8_t* trits /* N*n_words */);
size_t t3o_words_from_subword_stream(const uint8_t* trits, size_t n_trits, int N, uint8_t fill, uint8_t* words9 /* ceil(n/N) */);
size_t t3o_base243_pack(const uint8_t* trits, size_t n_trits, uint8_t* out /* 4 + ceil(n/5) */);
int    t3o_base243_unpack(const uint8_t* in, size_t n_bytes, uint8_t* trits, size_t cap, size_t* n_trits);
void   t3o_v6new_pack_pixels(const t3o_pixel* px, size_t n_px, uint32_t* words);
void   t3o_v6new_unpack_pixels(const uint32_t* words, size_t n_words, t3o_pixel* px);

#ifdef __cplusplus
}
#endif
#endif
