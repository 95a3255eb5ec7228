// dropin_main.cpp -- a caller written ONLY against the reference's public names
// (old/include/ternary_image_codec_v6_min.hpp + include/ternary_packing.hpp), in the style of
// old/src/main.cpp:15-26 and old/src/main_bare.cpp.  It is compiled twice, unchanged:
//   * against the reference headers            -> oracle/_ref/dropin_ref   (CPU, reference code)
//   * against this repo's include/ + libt3c.so -> tests/cpp/_dropin_ours   (B200)
// and tests/test_dropin_cpp.py requires the two programs to print identical lines.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <unistd.h>

#include "ternary_image_codec_v6_min.hpp"
#include "ternary_packing.hpp"
#include "t3v_io.hpp"
#include "t3v_indexed_io.hpp"

static uint64_t fnv(const void* p, size_t n, uint64_t h = 1469598103934665603ull)
{
    const uint8_t* b = static_cast<const uint8_t*>(p);
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
static uint32_t lcg_state = 12345;
static uint32_t lcg() { lcg_state = lcg_state * 1664525u + 1013904223u; return lcg_state >> 8; }

int main()
{
    // pixels -> raw words
    std::vector<PixelYCbCrQuant> px(20001);
    for (auto& p : px) { p.Yq = (uint16_t)(lcg() % 243); p.Cbq = (int16_t)((int)(lcg() % 81) - 40); p.Crq = (int16_t)((int)(lcg() % 81) - 40); }
    std::vector<Word27> raw;
    encode_raw_pixels_to_words(px, raw);
    std::printf("raw %zu %016llx\n", raw.size(), (unsigned long long)fnv(raw.data(), raw.size() * 9));
    std::vector<PixelYCbCrQuant> back;
    decode_raw_words_to_pixels(raw, back);
    std::printf("unpack %zu %016llx\n", back.size(), (unsigned long long)fnv(back.data(), back.size() * sizeof(PixelYCbCrQuant)));

    // profile encode over a small config matrix (old/src/main.cpp:17 uses P2 + tile + beacon)
    for (int variant = 0; variant < 5; ++variant) {
        EncoderContext e;
        switch (variant) {
        case 0: break;
        case 1: e.cfg.profile = ProfileID::P2_RS26_22; e.cfg.tile = {64, 64}; e.cfg.beacon = {83, 2, true}; break;
        case 2: e.cfg.profile = ProfileID::P3_RS26_20; uep_uniform(e.cfg.uep, 2); break;
        case 3: e.cfg.profile = ProfileID::P5_RS26_22_2D; e.cfg.tile = {26, 26}; uep_luma_priority(e.cfg.uep); e.cfg.beacon = {26, 2, true};
                e.cfg.seed = {2, 1, 1}; e.cfg.coset = CosetID::C1; break;
        case 4: e.cfg.profile = ProfileID::RAW_MODE; break;
        }
        std::vector<Word27> prof;
        const bool ok = encode_profile_from_raw(raw, prof, e);
        std::printf("encode[%d] %d %zu %016llx\n", variant, (int)ok, prof.size(), (unsigned long long)fnv(prof.data(), prof.size() * 9));
        DecoderContext d;
        std::vector<Word27> dec;
        const bool dok = decode_profile_to_raw(prof, dec, d); // as shipped: rejects its own encoder's output (bug B1)
        std::printf("decode[%d] %d %zu seen_profile=%d\n", variant, (int)dok, dec.size(), (int)d.cfg_last_seen.profile);
        std::vector<uint8_t> bytes;
        tpack::words_to_bytes(prof, bytes);
        std::printf("bytes[%d] %zu %016llx\n", variant, bytes.size(), (unsigned long long)fnv(bytes.data(), bytes.size()));
    }

    // sub-word trit streams and base-243 packing (what the .t3p writer does: old/include/t3p_io.hpp:19)
    for (int N : {27, 24, 15}) {
        std::vector<UTrit> trits, back_t;
        extract_subword_stream_from_words(raw, N, trits);
        std::vector<uint8_t> packed;
        tpack::ut_to_base243(trits, packed);
        const bool uok = tpack::base243_to_ut(packed, back_t);
        std::vector<Word27> rebuilt;
        build_words_from_subword_stream(back_t, N, rebuilt, (UTrit)1);
        std::printf("subword[%d] %zu %016llx packed %zu %016llx back %d %zu words %zu %016llx\n", N, trits.size(),
                    (unsigned long long)fnv(trits.data(), trits.size()), packed.size(), (unsigned long long)fnv(packed.data(), packed.size()), (int)uok,
                    back_t.size(), rebuilt.size(), (unsigned long long)fnv(rebuilt.data(), rebuilt.size() * 9));
        packed.pop_back();
        const bool sok = tpack::base243_to_ut(packed, back_t); // one payload byte short of the count it announces
        std::printf("subword_short[%d] %d %zu %016llx\n", N, (int)sok, back_t.size(), (unsigned long long)fnv(back_t.data(), back_t.size()));
    }

    // the .t3v container as old/src/main.cpp:21-25 and main_video_t3v.cpp:19-26 use it: header + frame records, then read back
    {
        FILE* f = std::tmpfile();
        const ActiveWindow aw = centered_window(SubwordMode::S27);
        const bool hok = t3v_write_header(f, ProfileID::P2_RS26_22, SubwordMode::S27, true, CosetID::C0, std_res_for(SubwordMode::S27).w,
                                          std_res_for(SubwordMode::S27).h, aw, 30000, 1001, 2, 1);
        std::vector<Word27> wild = raw;
        for (size_t i = 0; i < wild.size(); i += 7) wild[i].sym[i % 9] = (GF27)(200 + i % 50); // stored % 27
        const bool f1 = t3v_write_frame(f, raw), f2 = t3v_write_frame(f, wild);
        const long len = std::ftell(f);
        std::rewind(f);
        std::vector<uint8_t> all((size_t)len);
        const size_t got = std::fread(all.data(), 1, all.size(), f);
        std::printf("t3v write %d %d %d %ld %016llx\n", (int)hok, (int)f1, (int)f2, len, (unsigned long long)fnv(all.data(), got));
        std::rewind(f);
        T3VHeaderBin hb{};
        const bool rh = t3v_read_header(f, hb);
        std::vector<Word27> r1, r2, r3;
        const bool b1 = t3v_read_frame(f, r1), b2 = t3v_read_frame(f, r2), b3 = t3v_read_frame(f, r3);
        std::printf("t3v read %d frames=%u sub=%d aw=%u,%u,%u,%u %d %zu %016llx %d %zu %016llx %d %zu\n", (int)rh, hb.frame_count, (int)t3v_header_subword(hb),
                    t3v_header_aw(hb).x0, t3v_header_aw(hb).y0, t3v_header_aw(hb).w, t3v_header_aw(hb).h, (int)b1, r1.size(),
                    (unsigned long long)fnv(r1.data(), r1.size() * 9), (int)b2, r2.size(), (unsigned long long)fnv(r2.data(), r2.size() * 9), (int)b3, r3.size());
        std::fclose(f);
        // the index sidecar (old/include/t3v_indexed_io.hpp): scan the file just written, read the index back
        {
            char tn[] = "/tmp/t3c_dropin_XXXXXX";
            const int fd = mkstemp(tn);
            FILE* tf = fdopen(fd, "wb");
            std::fwrite(all.data(), 1, got, tf);
            std::fclose(tf);
            const std::string t3v = tn, idx = t3v + ".t3vi";
            const bool sok = t3v_scan_and_index(t3v, idx);
            T3VIndexBin ib{};
            std::vector<uint64_t> offs;
            const bool rok = t3v_index_read(idx, ib, offs);
            std::printf("t3vi %d %d count=%u crc=%08x offs=%zu %016llx\n", (int)sok, (int)rok, ib.frame_count, ib.header_crc32, offs.size(),
                        (unsigned long long)fnv(offs.data(), offs.size() * 8));
            std::remove(t3v.c_str());
            std::remove(idx.c_str());
        }
        // a damaged record is rejected
        all[54 + 4 + 100] ^= 1;
        FILE* g = std::tmpfile();
        std::fwrite(all.data(), 1, all.size(), g);
        std::rewind(g);
        const bool rh2 = t3v_read_header(g, hb);
        const bool d1 = t3v_read_frame(g, r1);
        std::printf("t3v damaged %d %d\n", (int)rh2, (int)d1);
        std::fclose(g);
    }

    // block-level RS with the reference's selftest data
    GF27Context gf;
    gf.init();
    for (ProfileID pid : {ProfileID::P1_RS26_24, ProfileID::P2_RS26_22, ProfileID::P3_RS26_20, ProfileID::P4_RS26_18}) {
        RSCodec rs;
        rs.init(&gf, rs_params_for(pid));
        const int k = rs.params.k;
        std::vector<GF27> data(k), code(26), outk(k, 0);
        for (int i = 0; i < k; ++i) data[i] = (GF27)((i * 5 + 7) % 27);
        rs.encode_block(data.data(), code.data());
        std::printf("rs_enc k=%d %016llx\n", k, (unsigned long long)fnv(code.data(), 26));
        code[3] = (GF27)((code[3] + 1) % 27);
        const bool ok = rs.decode_block(code.data(), outk.data());
        std::printf("rs_dec k=%d %d %016llx %016llx\n", k, (int)ok, (unsigned long long)fnv(code.data(), 26), (unsigned long long)fnv(outk.data(), k));
    }

    // 2D boustrophedon
    std::vector<GF27> sy(1000);
    for (auto& s : sy) s = (GF27)(lcg() % 27);
    interleave2D_boustrophedon(sy, Tile2D{7, 5});
    std::printf("il2d %016llx\n", (unsigned long long)fnv(sy.data(), sy.size()));
    deinterleave2D_boustrophedon(sy, Tile2D{7, 5});
    std::printf("dil2d %016llx\n", (unsigned long long)fnv(sy.data(), sy.size()));

    // ---- L0 / L1 public names (OLD:81-113, 155-380, 383-487, 693-722, 816-833)
    {
        uint64_t h = 1469598103934665603ull;
        for (int a = 0; a < 27; ++a)
            for (int b = 0; b < 27; ++b) {
                const uint8_t v[5] = {gf27_add((GF27)a, (GF27)b), gf27_sub((GF27)a, (GF27)b), gf27_mul_poly((GF27)a, (GF27)b), gf.mul((GF27)a, (GF27)b), gf.add((GF27)a, (GF27)b)};
                h = fnv(v, 5, h);
            }
        for (int a = 0; a < 27; ++a) { const int16_t v[3] = {(int16_t)gf.inv((GF27)a), (int16_t)gf.log((GF27)a), (int16_t)gf.pow_alpha(a * 7 - 40)}; h = fnv(v, sizeof v, h); }
        std::printf("gf27 prim=%d order3=%d %016llx\n", (int)gf.tab.primitive, gf.order_of(3), (unsigned long long)h);
    }
    {
        const ScramblerSeed seeds[3] = {{1, 1, 1}, {2, 1, 1}, {4000000007u, 3000000001u, 5}};
        for (const auto& sd : seeds) {
            uint32_t st = sd.s0, st2 = sd.s0;
            std::vector<GF27> a(40), b(40);
            for (int i = 0; i < 40; ++i) { a[i] = scramble_symbol((GF27)((i * 11) % 27), sd, st); b[i] = descramble_symbol(a[i], sd, st2); }
            std::printf("scramble %016llx %016llx st=%u,%u\n", (unsigned long long)fnv(a.data(), 40), (unsigned long long)fnv(b.data(), 40), st, st2);
        }
        const BeaconPayload bp{ProfileID::P5_RS26_22_2D, 8192 % 5 + 10, 7};
        std::printf("beacon %d %d\n", (int)encode_beacon_symbol(bp), (int)encode_beacon_symbol(BeaconPayload{ProfileID::P2_RS26_22, 2, 0}));
    }
    {
        SuperframeHeader hd;
        hd.profile = ProfileID::P5_RS26_22_2D; uep_luma_priority(hd.uep); hd.tile = {26, 7}; hd.seed = {2, 30, 55}; hd.band_map_hash = 12345; hd.frame_seq = 6789;
        hd.beacon = {83, 11, true}; hd.subword = SubwordMode::S21; hd.centered = false; hd.coset = CosetID::C2;
        HeaderPack hp = HeaderCodec::pack(hd);
        const SuperframeHeader u = HeaderCodec::unpack(hp);
        std::printf("header %016llx check=%d magic=%u ver=%u prof=%d uep=%d%d%d%d%d%d%d%d%d tile=%u,%u seed=%u,%u,%u hash=%u seq=%u beacon=%d,%u,%u sub=%d cen=%d coset=%d\n",
                    (unsigned long long)fnv(hp.symbols.data(), 27), (int)HeaderCodec::check(hp), u.magic, u.version, (int)u.profile, u.uep.band_profile[0], u.uep.band_profile[1],
                    u.uep.band_profile[2], u.uep.band_profile[3], u.uep.band_profile[4], u.uep.band_profile[5], u.uep.band_profile[6], u.uep.band_profile[7], u.uep.band_profile[8],
                    u.tile.w, u.tile.h, u.seed.a, u.seed.b, u.seed.s0, u.band_map_hash, u.frame_seq, (int)u.beacon.enabled, (unsigned)u.beacon.band_slot, u.beacon.words_period,
                    (int)u.subword, (int)u.centered, (int)u.coset);
        hp.symbols[5] = (GF27)((hp.symbols[5] + 1) % 27);
        std::printf("header damaged check=%d default %016llx\n", (int)HeaderCodec::check(hp), (unsigned long long)fnv(HeaderCodec::pack(SuperframeHeader{}).symbols.data(), 27));
        std::vector<UTrit> msg(69);
        for (size_t i = 0; i < msg.size(); ++i) msg[i] = (UTrit)(lcg() % 3);
        std::array<UTrit, CRC3::L> rem{};
        CRC3::rem12(msg, rem);
        std::printf("crc3 %016llx\n", (unsigned long long)fnv(rem.data(), rem.size()));
    }
    {
        Word27 w{};
        PixelYCbCrQuant a{200, -17, 33}, b{7, 40, -40}, c, d;
        pack_two_pixels(a, b, w);
        unpack_two_pixels(w, c, d);
        std::array<UTrit, 27> tr{};
        extract_subword_trits_from_word(w, 15, tr);
        Word27 w2{};
        inject_subword_trits_into_word(tr.data(), 15, w2, (UTrit)2);
        std::printf("pair %016llx %u,%d,%d %u,%d,%d trits %016llx inject %016llx\n", (unsigned long long)fnv(w.sym.data(), 9), c.Yq, c.Cbq, c.Crq, d.Yq, d.Cbq, d.Crq,
                    (unsigned long long)fnv(tr.data(), 27), (unsigned long long)fnv(w2.sym.data(), 9));
    }

    std::printf("selftests RS:%s API:%s\n", selftest_rs_unit() ? "OK" : "FAIL", selftest_api_roundtrip() ? "OK" : "FAIL");
    return 0;
}
