// t3c_internal.h -- structures shared by the host side (tables, geometry, context) and the kernels.
// Not part of the public ABI (that is include/t3c.h).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

#include "../../include/t3c.h"

namespace t3c {

// ----------------------------------------------------------------------------------------------
// Device-resident tables (built once per context by tables.cpp)
// ----------------------------------------------------------------------------------------------
// GF(27) look-ups used by the thread-per-codeword decoder and the header kernels.
struct GfTables {
    uint8_t mul[736];   // mul[a*27+b]               (OLD:458)
    uint8_t add[736];   // add[a*27+b], trit-wise     (OLD:383-388)
    uint8_t inv[32];    // inv[a], inv[0]=0           (OLD:459-465)
    uint8_t exp[32];    // alpha^e, e=0..25           (OLD:450-457), alpha=3
    uint8_t neg[32];    // 0-a
    uint8_t scr[3][32]; // scr[st][s] = s (+) 13*st   (scramble_symbol, OLD:81-87)
    uint8_t dsc[3][32]; // dsc[st][s] = s (-) 13*st   (descramble_symbol, OLD:88-94)
    uint8_t lg[32];     // lg[alpha^e] = e, lg[0] = 255
    uint8_t sqr[32];    // sqr[a] = a square root of a (the other one is its negative), sqr[0] = 0, 255 when a is not a square
    // syn[kidx][j][m] = -(alpha^((j+1)(k+m))): the power-sum syndromes of a received block from its parity residual
    // p = parity(received data) - received parity (repaired code):  S_j = sum_m syn[j][m] * p_m
    uint8_t syn[4][8][8];
};

// Bit-plane row tables: the RS encoder as shipped (B1) and repaired are both GF(27)-linear maps
// parity = data * P, hence GF(3)-linear on trits.  Row i, symbol value d holds the 3r
// parity trits of d*P[i][*] in two bit planes:  lo32 bit = (trit != 0), hi32 bit = (trit == 2),
// trit c of parity symbol j at bit 8*(j&3) + 4*(j>>2) + c.  Rows k..25 hold -(e_j * d) so that the
// sum over all 26 received symbols is zero iff the block is a codeword of that encoder.
constexpr int kRows = 26, kVals = 27;
struct RowTable { uint64_t e[kRows][kVals]; };      // 5616 B
// index: [arith][kidx] with kidx = (24-k)/2
// Root search of the bounded-distance decoder on bit planes: e[j-1][v] holds v * alpha^(-i*j) for the 26 positions i as GF(3) trits,
// ten positions per word, trit c of position i at bit 3*(i%10)+c of word i/10; words 0..2 are the nz plane, 3..5 the two plane.
// 1 + sum_j sigma_j * alpha^(-i*j) over all positions at once is then one plane addition per coefficient; position i is a root where
// its three nz bits are clear.  Lies directly behind GfTables in the device copy (chien_of below).
struct ChienTables { uint32_t e[4][27][6]; };
struct RsTables {
    RowTable row[2][4];
    uint8_t  gen[4][12];      // generator polynomials, low-first (OLD:501-516)
    uint8_t  par[2][4][24][8]; // P[i][j] as symbols (general/per-symbol kernels)
    // the same rows for the tiled kernels: {nz, two} planes with trit c of parity symbol j at bit 8 + 4j + c (k >= 20; k = 18: its 24
    // parity trits fill bits 8..31 densely, bit 8 + 3j + c, and are spread to nibbles when they leave the planes), so that the low byte of an entry is free for an embedded (de)scrambled symbol and each
    // nibble of (plane >> 8) is a PRMT selector (planes -> symbol bytes without shifts or tables)
    uint32_t pl[2][4][kRows][kVals][2];
};

inline int kidx_of(int k) { return (24 - k) / 2; }
inline bool k_valid(int k) { return k == 24 || k == 22 || k == 20 || k == 18; }

// ----------------------------------------------------------------------------------------------
// Per-call geometry of one super-frame, passed to kernels by value.
// ----------------------------------------------------------------------------------------------
struct Geom {
    uint64_t n_words;     // N_w raw words
    uint64_t n_s;         // regrouped symbols ceil(26 N_w / 3)               (A.1)
    uint64_t s_b[9];      // band lengths                                     (A.3)
    uint64_t ncw[9];      // codewords per band
    uint64_t cw_base[9];  // codewords before band b (band-major body)
    uint64_t use_base[9]; // decoded symbols before band b (reference decoder's `use`)
    uint64_t n_cw;        // total codewords
    uint64_t l_body;      // 26 * n_cw
    uint64_t l_exp;       // after beacon expansion                           (A.5)
    uint64_t n_out;       // profile words
    uint64_t tile_area;   // w*h when the 2D interleave is active, else 0      (A.2)
    uint32_t tile_w;
    uint32_t period;      // beacon period, 0 = no beacon
    uint32_t beacon_per;  // 9*period-1 body symbols per period block (when the slot exists)
    int32_t  slot;        // beacon slot, -1 when no slot is ever replaced (band_slot > 8)
    int32_t  k[9];
    int32_t  uniform_k;   // k when all nine bands share it, else 0
    uint8_t  bsym;        // beacon symbol                                    (OLD:107-113,1130)
    uint8_t  st[8];       // scrambler state for body index p: p<2 -> st[p], else st[2+(p-2)%6]  (A.4)
    uint8_t  arith;
    uint8_t  hdr[52];     // coded header (filled by the header kernel for device paths)
};

struct HostTables { GfTables gf; ChienTables ch; RsTables rs; };
static_assert(offsetof(HostTables, ch) == sizeof(GfTables) && sizeof(GfTables) % 16 == 0, "ChienTables must follow GfTables directly (chien_of)");
void build_tables(HostTables& t);
void fast_check_constants(const HostTables& H, const Geom& g, uint32_t chk_nz[7], uint32_t chk_two[7]);

// host-side geometry; returns false on invalid config values
void make_geom(const t3c_config& c, size_t n_words, int arith, Geom& g);
// smallest N_w whose encode has n_out words, and its geometry; false if no N_w matches
bool geom_from_nout(const t3c_config& c, size_t n_out, int arith, Geom& g);
size_t profile_words(const t3c_config& c, size_t n_words);
void scrambler_states(uint32_t a, uint32_t b, uint32_t s0, uint8_t st[8]);
inline bool use_2d(const t3c_config& c) { return c.profile == 4 && c.tile_w && c.tile_h; }
inline bool use_beacon(const t3c_config& c) { return c.beacon_enabled && c.beacon_period > 0; }

} // namespace t3c
