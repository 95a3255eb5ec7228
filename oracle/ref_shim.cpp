// ref_shim.cpp -- thin extern "C" wrapper around the UNMODIFIED reference (TEST INFRASTRUCTURE).
//
// Compiled by oracle/Makefile against the reference sources where they lie under
// /root/reference (never copied into this repo); the outputs go to oracle/_ref/:
//   libt3ref.so        reference as shipped                      (REF-EXACT oracle)
//   libt3ref_fixed.so  same sources with the 3-line arithmetic repair of SURVEY.md
//                      Appendix B applied by sed into a temp dir  (FIXED RS oracle)
// The per-pixel RGB<->quant functions are lines 47-84 of old/include/io_image.hpp, extracted
// at build time into $(TMP)/ref_bridge_extract.inc because the whole header does not compile
// (SURVEY.md 0.1).  Every function below only marshals plain arrays to the reference's types.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "ternary_image_codec_v6_min.hpp" // -I<reference>/old/include or the patched temp copy
#include "ref_bridge_extract.inc"          // rgb_to_ycbcr, ycbcr_to_rgb, quantize_ycbcr, dequantize_ycbcr
#include REF_TPACK_HEADER                  // <reference>/include/ternary_packing.hpp (tpack::), staged in the temp dir so that its own #include resolves to OLD

#include "t3v_io.hpp" // <reference>/old/include: .t3v container

#include "t3_oracle.h" // t3o_cfg / t3o_pixel layouts only

static_assert(sizeof(Word27) == 9, "Word27 must be 9 bytes");
static_assert(sizeof(PixelYCbCrQuant) == 6, "PixelYCbCrQuant must be 6 bytes");

namespace {
EncoderContext& ectx() { static EncoderContext e; return e; }
RSCodec* codec_for(int k)
{
    EncoderContext& e = ectx();
    switch (k) { case 24: return &e.rs_p1; case 22: return &e.rs_p2; case 20: return &e.rs_p3; case 18: return &e.rs_p4; }
    return nullptr;
}
void to_cfg(const t3o_cfg* c, EncoderConfig& o)
{
    o.profile = (ProfileID)c->profile;
    for (int i = 0; i < 9; ++i) o.uep.band_profile[i] = c->uep[i];
    o.tile = Tile2D{c->tile_w, c->tile_h};
    o.seed = ScramblerSeed{c->seed_a, c->seed_b, c->seed_s0};
    o.beacon.words_period = c->beacon_period; o.beacon.band_slot = c->beacon_slot; o.beacon.enabled = c->beacon_enabled != 0;
    o.superframe_words = c->superframe_words;
    o.subword = (SubwordMode)c->subword; o.centered = c->centered != 0; o.coset = (CosetID)c->coset;
}
std::vector<Word27> to_words(const uint8_t* p, size_t n)
{
    std::vector<Word27> v(n);
    if (n) std::memcpy(v.data(), p, 9 * n);
    return v;
}
} // namespace

extern "C" {

void t3r_gf_tables(uint8_t* exp78, int16_t* log27, uint8_t* mul729, uint8_t* inv27, uint8_t* prim)
{
    const GF27Tables& t = ectx().gf.tab;
    std::memcpy(exp78, t.exp.data(), 78); std::memcpy(log27, t.log.data(), 27 * sizeof(int16_t));
    std::memcpy(mul729, t.mul.data(), 729); std::memcpy(inv27, t.inv.data(), 27); *prim = t.primitive;
}
uint8_t t3r_gf_add(uint8_t a, uint8_t b) { return gf27_add(a, b); }
uint8_t t3r_gf_sub(uint8_t a, uint8_t b) { return gf27_sub(a, b); }
uint8_t t3r_gf_mul(uint8_t a, uint8_t b) { return gf27_mul_poly(a, b); }

int t3r_rs_gen(int k, uint8_t* g)
{
    RSCodec* r = codec_for(k); if (!r) return -1;
    std::memcpy(g, r->g.data(), r->g.size());
    return (int)r->g.size();
}
void t3r_rs_encode(int k, const uint8_t* d, uint8_t* out26) { codec_for(k)->encode_block(d, out26); }
int t3r_rs_decode(int k, uint8_t* inout26, uint8_t* outk) { return codec_for(k)->decode_block(inout26, outk) ? 1 : 0; }
void t3r_rs_encode_blocks(int k, const uint8_t* d, size_t n, uint8_t* out)
{
    RSCodec* r = codec_for(k);
    for (size_t i = 0; i < n; ++i) r->encode_block(d + i * (size_t)k, out + 26 * i);
}
void t3r_rs_decode_blocks(int k, uint8_t* inout, size_t n, uint8_t* out, uint8_t* ok)
{
    RSCodec* r = codec_for(k);
    for (size_t i = 0; i < n; ++i) {
        bool f = r->decode_block(inout + 26 * i, out + i * (size_t)k);
        if (!f) std::memset(out + i * (size_t)k, 0, (size_t)k);
        ok[i] = f ? 1 : 0;
    }
}

void t3r_crc12(const uint8_t* tr, size_t n, uint8_t* out12)
{
    std::vector<UTrit> m(tr, tr + n); std::array<UTrit, 12> r{};
    CRC3::rem12(m, r); std::memcpy(out12, r.data(), 12);
}
void t3r_header_pack(const t3o_cfg* c, uint32_t frame_seq, uint32_t hash, uint8_t* sym27)
{
    EncoderConfig e; to_cfg(c, e);
    SuperframeHeader h{};
    h.profile = e.profile; h.uep = e.uep; h.tile = e.tile; h.seed = e.seed; h.beacon = e.beacon;
    h.subword = e.subword; h.centered = e.centered; h.coset = e.coset; h.frame_seq = frame_seq; h.band_map_hash = hash;
    HeaderPack p = HeaderCodec::pack(h);
    std::memcpy(sym27, p.symbols.data(), 27);
}
int t3r_header_check(const uint8_t* sym27)
{
    HeaderPack p{}; std::memcpy(p.symbols.data(), sym27, 27);
    return HeaderCodec::check(p) ? 1 : 0;
}
void t3r_header_unpack(const uint8_t* sym27, t3o_cfg* o, uint32_t* frame_seq, uint32_t* hash, uint16_t* magic, uint8_t* version)
{
    HeaderPack p{}; std::memcpy(p.symbols.data(), sym27, 27);
    SuperframeHeader h = HeaderCodec::unpack(p);
    o->profile = (uint8_t)h.profile;
    for (int i = 0; i < 9; ++i) o->uep[i] = h.uep.band_profile[i];
    o->tile_w = h.tile.w; o->tile_h = h.tile.h;
    o->seed_a = h.seed.a; o->seed_b = h.seed.b; o->seed_s0 = h.seed.s0;
    o->beacon_period = h.beacon.words_period; o->beacon_slot = h.beacon.band_slot; o->beacon_enabled = h.beacon.enabled;
    o->subword = (uint8_t)h.subword; o->centered = h.centered; o->coset = (uint8_t)h.coset;
    if (frame_seq) *frame_seq = h.frame_seq;
    if (hash) *hash = h.band_map_hash;
    if (magic) *magic = h.magic;
    if (version) *version = h.version;
}

size_t t3r_pack_pixels(const t3o_pixel* px, size_t n, uint8_t* words9)
{
    std::vector<PixelYCbCrQuant> v(n);
    if (n) std::memcpy(static_cast<void*>(v.data()), px, 6 * n);
    std::vector<Word27> w; encode_raw_pixels_to_words(v, w);
    if (!w.empty()) std::memcpy(words9, w.data(), 9 * w.size());
    return w.size();
}
void t3r_unpack_pixels(const uint8_t* words9, size_t nw, t3o_pixel* px)
{
    std::vector<PixelYCbCrQuant> v; decode_raw_words_to_pixels(to_words(words9, nw), v);
    if (!v.empty()) std::memcpy(px, v.data(), 6 * v.size());
}
void t3r_interleave2d(uint8_t* sy, size_t n, unsigned w, unsigned h)
{
    std::vector<GF27> v(sy, sy + n); interleave2D_boustrophedon(v, Tile2D{(uint16_t)w, (uint16_t)h});
    if (n) std::memcpy(sy, v.data(), n);
}
void t3r_deinterleave2d(uint8_t* sy, size_t n, unsigned w, unsigned h)
{
    std::vector<GF27> v(sy, sy + n); deinterleave2D_boustrophedon(v, Tile2D{(uint16_t)w, (uint16_t)h});
    if (n) std::memcpy(sy, v.data(), n);
}
uint8_t t3r_scramble_symbol(uint8_t s, uint32_t a, uint32_t b, uint32_t* st) { ScramblerSeed sd{a, b, 0}; return scramble_symbol(s, sd, *st); }
uint8_t t3r_descramble_symbol(uint8_t s, uint32_t a, uint32_t b, uint32_t* st) { ScramblerSeed sd{a, b, 0}; return descramble_symbol(s, sd, *st); }
uint8_t t3r_beacon_symbol(uint8_t profile, uint16_t seq, uint8_t health) { return encode_beacon_symbol(BeaconPayload{(ProfileID)profile, seq, health}); }

size_t t3r_encode_profile(const t3o_cfg* c, const uint8_t* raw9, size_t n, uint8_t* out9, size_t cap)
{
    EncoderContext e; to_cfg(c, e.cfg);
    std::vector<Word27> out;
    encode_profile_from_raw(to_words(raw9, n), out, e);
    if (out.size() > cap) return (size_t)-1;
    if (!out.empty()) std::memcpy(out9, out.data(), 9 * out.size());
    return out.size();
}
int t3r_decode_profile(t3o_cfg* seen, const uint8_t* in9, size_t n, uint8_t* out9, size_t cap, size_t* n_out)
{
    DecoderContext d;
    DecoderConfigSeen& s = d.cfg_last_seen;
    s.profile = (ProfileID)seen->profile;
    for (int i = 0; i < 9; ++i) s.uep.band_profile[i] = seen->uep[i];
    s.tile = Tile2D{seen->tile_w, seen->tile_h}; s.seed = ScramblerSeed{seen->seed_a, seen->seed_b, seen->seed_s0};
    s.beacon.words_period = seen->beacon_period; s.beacon.band_slot = seen->beacon_slot; s.beacon.enabled = seen->beacon_enabled != 0;
    s.subword = (SubwordMode)seen->subword; s.centered = seen->centered != 0; s.coset = (CosetID)seen->coset;
    std::vector<Word27> out;
    bool ok = decode_profile_to_raw(to_words(in9, n), out, d);
    seen->profile = (uint8_t)s.profile;
    for (int i = 0; i < 9; ++i) seen->uep[i] = s.uep.band_profile[i];
    seen->tile_w = s.tile.w; seen->tile_h = s.tile.h;
    seen->seed_a = s.seed.a; seen->seed_b = s.seed.b; seen->seed_s0 = s.seed.s0;
    seen->beacon_period = s.beacon.words_period; seen->beacon_slot = s.beacon.band_slot; seen->beacon_enabled = s.beacon.enabled;
    seen->subword = (uint8_t)s.subword; seen->centered = s.centered; seen->coset = (uint8_t)s.coset;
    *n_out = 0;
    if (!ok) return 0;
    if (out.size() > cap) return 0;
    if (!out.empty()) std::memcpy(out9, out.data(), 9 * out.size());
    *n_out = out.size();
    return 1;
}

void t3r_rgb_to_quant(const uint8_t* rgb, size_t n, t3o_pixel* out) // loop body of rgb_to_quant_stream, io_image.hpp:156-170
{
    for (size_t i = 0; i < n; ++i) {
        uint8_t Y, Cb, Cr;
        rgb_to_ycbcr(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], Y, Cb, Cr);
        PixelYCbCrQuant q = quantize_ycbcr(Y, Cb, Cr);
        std::memcpy(&out[i], &q, 6);
    }
}
void t3r_quant_to_rgb(const t3o_pixel* px, size_t n, uint8_t* rgb) // loop body of quant_stream_to_rgb, io_image.hpp:171-192
{
    for (size_t i = 0; i < n; ++i) {
        PixelYCbCrQuant q; std::memcpy(static_cast<void*>(&q), &px[i], 6);
        uint8_t Y, Cb, Cr; dequantize_ycbcr(q, Y, Cb, Cr);
        ycbcr_to_rgb(Y, Cb, Cr, rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
    }
}

void t3r_words_to_bytes(const uint8_t* words9, size_t n, uint8_t* out)
{
    std::vector<uint8_t> b; tpack::words_to_bytes(to_words(words9, n), b);
    if (!b.empty()) std::memcpy(out, b.data(), b.size());
}
size_t t3r_bytes_to_words(const uint8_t* bytes, size_t nbytes, uint8_t* words9)
{
    std::vector<uint8_t> b(bytes, bytes + nbytes); std::vector<Word27> w; tpack::bytes_to_words(b, w);
    if (!w.empty()) std::memcpy(words9, w.data(), 9 * w.size());
    return w.size();
}
void t3r_selftests(int* rs_ok, int* api_ok) { *rs_ok = selftest_rs_unit() ? 1 : 0; *api_ok = selftest_api_roundtrip() ? 1 : 0; }

/* fused convenience: the chain of old/src/main.cpp:15-19 on an in-memory RGB8 buffer */
size_t t3r_encode_rgb(const t3o_cfg* c, const uint8_t* rgb, size_t n_px, uint8_t* out9, size_t cap)
{
    std::vector<PixelYCbCrQuant> q(n_px);
    for (size_t i = 0; i < n_px; ++i) {
        uint8_t Y, Cb, Cr; rgb_to_ycbcr(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], Y, Cb, Cr);
        q[i] = quantize_ycbcr(Y, Cb, Cr);
    }
    std::vector<Word27> raw; encode_raw_pixels_to_words(q, raw);
    EncoderContext e; to_cfg(c, e.cfg);
    std::vector<Word27> out; encode_profile_from_raw(raw, out, e);
    if (out.size() > cap) return (size_t)-1;
    if (!out.empty()) std::memcpy(out9, out.data(), 9 * out.size());
    return out.size();
}

// ---- sub-word streams (OLD:816-859) and base-243 (include/ternary_packing.hpp:18-50)
void t3r_subword_stream(const uint8_t* words9, size_t n_words, int N, uint8_t* trits)
{
    std::vector<UTrit> out;
    extract_subword_stream_from_words(to_words(words9, n_words), N, out);
    if (!out.empty()) std::memcpy(trits, out.data(), out.size());
}
size_t t3r_words_from_subword_stream(const uint8_t* trits, size_t n_trits, int N, uint8_t fill, uint8_t* words9)
{
    std::vector<UTrit> in(trits, trits + n_trits);
    std::vector<Word27> out;
    build_words_from_subword_stream(in, N, out, fill);
    if (!out.empty()) std::memcpy(words9, out.data(), 9 * out.size());
    return out.size();
}
size_t t3r_base243_pack(const uint8_t* trits, size_t n_trits, uint8_t* out_bytes)
{
    std::vector<UTrit> in(trits, trits + n_trits);
    std::vector<uint8_t> out;
    tpack::ut_to_base243(in, out);
    std::memcpy(out_bytes, out.data(), out.size());
    return out.size();
}
int t3r_base243_unpack(const uint8_t* in_bytes, size_t n_bytes, uint8_t* trits, size_t cap, size_t* n_trits)
{
    std::vector<uint8_t> in(in_bytes, in_bytes + n_bytes);
    std::vector<UTrit> out;
    const bool ok = tpack::base243_to_ut(in, out);
    *n_trits = out.size();
    std::memcpy(trits, out.data(), std::min(cap, out.size()));
    return ok ? 1 : 0;
}

} // extern "C"

// ---- .t3v container records through the reference's own FILE* functions (old/include/t3v_io.hpp), on a tmpfile
extern "C" {
uint32_t t3r_crc32(const uint8_t* data, size_t n) { return t3v_detail::crc32(data, n); }
size_t t3r_t3v_frame_record(const uint8_t* words9, uint32_t n_words, uint8_t* out)
{
    std::vector<Word27> w(n_words);
    if (n_words) std::memcpy(w.data(), words9, 9 * (size_t)n_words);
    FILE* f = std::tmpfile();
    if (!f) return 0;
    const bool ok = t3v_write_frame(f, w);
    const long len = std::ftell(f);
    std::rewind(f);
    size_t got = ok && len > 0 ? std::fread(out, 1, (size_t)len, f) : 0;
    std::fclose(f);
    return got;
}
int t3r_t3v_read_frame(const uint8_t* rec, size_t n_bytes, uint8_t* words9, uint32_t* n_words)
{
    FILE* f = std::tmpfile();
    if (!f) return 0;
    if (n_bytes) std::fwrite(rec, 1, n_bytes, f);
    std::rewind(f);
    std::vector<Word27> w;
    const bool ok = t3v_read_frame(f, w);
    std::fclose(f);
    *n_words = ok ? (uint32_t)w.size() : 0;
    if (ok && !w.empty()) std::memcpy(words9, w.data(), 9 * w.size());
    return ok ? 1 : 0;
}
size_t t3r_t3v_header(uint8_t* out, int profile, int subword, int centered, int coset, uint32_t w, uint32_t h, const uint32_t aw[4],
                      uint32_t fps_num, uint32_t fps_den, uint32_t frame_count, int file_type, int* reads_back)
{
    FILE* f = std::tmpfile();
    if (!f) return 0;
    ActiveWindow a{};
    a.x0 = (decltype(a.x0))aw[0]; a.y0 = (decltype(a.y0))aw[1]; a.w = (decltype(a.w))aw[2]; a.h = (decltype(a.h))aw[3];
    const bool ok = t3v_write_header(f, (ProfileID)profile, (SubwordMode)subword, centered != 0, (CosetID)coset, w, h, a, fps_num, fps_den, frame_count, (uint8_t)file_type);
    const long len = std::ftell(f);
    std::rewind(f);
    size_t got = ok && len > 0 ? std::fread(out, 1, (size_t)len, f) : 0;
    std::rewind(f);
    T3VHeaderBin hb{};
    *reads_back = t3v_read_header(f, hb) ? 1 : 0;
    std::fclose(f);
    return got;
}
}
