// t3v_indexed_io.hpp -- drop-in for old/include/t3v_indexed_io.hpp of the reference: the .t3vi index sidecar of a .t3v file
// (17-byte packed header "T3VI" | version | frame_count | reserved | CRC-32 of the 13 bytes before it, then one uint64 byte offset
// per frame record).  Same names and signatures; the checksum is the device CRC-32 behind t3v_detail::crc32 (t3v_io.hpp), the
// offset table of frames emitted on the device comes from t3c_t3v_index_build (every record's offset is known there without a scan).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "t3v_io.hpp"

#pragma pack(push, 1)
struct T3VIndexBin {
    char magic[4];
    uint8_t version;
    uint32_t frame_count;
    uint32_t reserved0;
    uint32_t header_crc32;
};
#pragma pack(pop)
static_assert(sizeof(T3VIndexBin) == 17, "packed .t3vi header");

inline bool t3v_index_write(const std::string& idx_path, uint32_t frame_count, const std::vector<uint64_t>& offsets)
{
    FILE* f = std::fopen(idx_path.c_str(), "wb");
    if (!f) return false;
    T3VIndexBin h{};
    std::memcpy(h.magic, "T3VI", 4);
    h.version = 1;
    h.frame_count = frame_count;
    h.reserved0 = 0;
    h.header_crc32 = t3v_detail::crc32(&h, sizeof(T3VIndexBin) - sizeof(uint32_t));
    bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1;
    ok = ok && (!offsets.empty() ? std::fwrite(offsets.data(), offsets.size() * sizeof(uint64_t), 1, f) == 1 : true);
    std::fclose(f);
    return ok;
}
inline bool t3v_index_read(const std::string& idx_path, T3VIndexBin& h, std::vector<uint64_t>& offsets)
{
    FILE* f = std::fopen(idx_path.c_str(), "rb");
    if (!f) return false;
    bool ok = std::fread(&h, sizeof(h), 1, f) == 1 && std::memcmp(h.magic, "T3VI", 4) == 0 &&
              t3v_detail::crc32(&h, sizeof(T3VIndexBin) - sizeof(uint32_t)) == h.header_crc32;
    if (ok) {
        offsets.resize(h.frame_count);
        ok = !h.frame_count || std::fread(offsets.data(), offsets.size() * sizeof(uint64_t), 1, f) == 1;
    }
    std::fclose(f);
    return ok;
}
// walk the records of a .t3v file (count, 9 x count payload bytes, CRC) and write their offsets
inline bool t3v_scan_and_index(const std::string& t3v_path, const std::string& idx_path)
{
    FILE* f = std::fopen(t3v_path.c_str(), "rb");
    if (!f) return false;
    T3VHeaderBin th{};
    if (!t3v_read_header(f, th)) { std::fclose(f); return false; }
    std::vector<uint64_t> offs;
    for (;;) {
        const long here = std::ftell(f);
        if (here < 0) break;
        uint32_t n = 0, crc = 0;
        if (std::fread(&n, sizeof(n), 1, f) != 1) break;
        if (n && std::fseek(f, (long)((size_t)n * 9), SEEK_CUR) != 0) break;
        if (std::fread(&crc, sizeof(crc), 1, f) != 1) break;
        offs.push_back((uint64_t)here);
    }
    std::fclose(f);
    return t3v_index_write(idx_path, (uint32_t)offs.size(), offs);
}
// the same sidecar for frames whose records were emitted by t3c_t3v_frame_records_dev (n_words[i] words in frame i, the first record
// at byte `first_offset` of the .t3v file, normally 54): no scan, the offsets follow from the record sizes
inline bool t3v_index_write_for_records(const std::string& idx_path, const std::vector<uint64_t>& n_words, uint64_t first_offset = sizeof(T3VHeaderBin))
{
    std::vector<uint8_t> buf(17 + 8 * n_words.size());
    size_t nb = 0;
    if (t3c_t3v_index_build(t3c_shim::context(), n_words.data(), n_words.size(), first_offset, buf.data(), &nb) != T3C_OK) return false;
    FILE* f = std::fopen(idx_path.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(buf.data(), nb, 1, f) == 1;
    std::fclose(f);
    return ok;
}
