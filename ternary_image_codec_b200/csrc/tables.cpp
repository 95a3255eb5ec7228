// tables.cpp -- host-side construction of the constant tables and the per-call geometry.
//
// Everything here runs once per context (tables) or once per call (geometry, O(1)); none of it
// touches pixel or symbol data.  Field: GF(27) = GF(3)[x]/(x^3+2x+1), alpha = x = symbol 3
// (the reference searches the first element of order 26, OLD:436-447, which is 3).
#include "t3c_internal.h"

#include <cstring>

namespace t3c {
namespace {

struct Tr { int t[3]; };
Tr split(int s) { return Tr{{s % 3, (s / 3) % 3, (s / 9) % 3}}; }
int join(const Tr& a) { return a.t[0] + 3 * a.t[1] + 9 * a.t[2]; }
int gadd(int a, int b) { Tr x = split(a), y = split(b); return join(Tr{{(x.t[0] + y.t[0]) % 3, (x.t[1] + y.t[1]) % 3, (x.t[2] + y.t[2]) % 3}}); }
int gneg(int a) { Tr x = split(a); return join(Tr{{(3 - x.t[0]) % 3, (3 - x.t[1]) % 3, (3 - x.t[2]) % 3}}); }
int gsub(int a, int b) { return gadd(a, gneg(b)); }
// multiply by x: (t0 + t1 x + t2 x^2) x = t0 x + t1 x^2 + t2 (x + 2)
int mulx(int a) { Tr v = split(a); return join(Tr{{(2 * v.t[2]) % 3, (v.t[0] + v.t[2]) % 3, v.t[1]}}); }
int gmul_slow(int a, int b)
{
    Tr y = split(b);
    int acc = 0, ax = a;
    for (int i = 0; i < 3; ++i) {
        for (int c = 0; c < y.t[i]; ++c) acc = gadd(acc, ax);
        ax = mulx(ax);
    }
    return acc;
}

// plane bit of trit c of parity symbol j
inline int plane_bit(int j, int c) { return 8 * (j & 3) + 4 * (j >> 2) + c; }
uint64_t planes_of(const int* sym, int r)
{
    uint32_t nz = 0, two = 0;
    for (int j = 0; j < r; ++j) {
        Tr v = split(sym[j]);
        for (int c = 0; c < 3; ++c) {
            if (v.t[c]) nz |= 1u << plane_bit(j, c);
            if (v.t[c] == 2) two |= 1u << plane_bit(j, c);
        }
    }
    return (uint64_t)nz | ((uint64_t)two << 32);
}

} // namespace

void build_tables(HostTables& T)
{
    std::memset(&T, 0, sizeof T);
    GfTables& g = T.gf;
    int ex[26], lg[27];
    ex[0] = 1;
    for (int i = 1; i < 26; ++i) ex[i] = mulx(ex[i - 1]);
    for (int i = 0; i < 27; ++i) lg[i] = -1;
    for (int i = 0; i < 26; ++i) lg[ex[i]] = i;
    for (int a = 0; a < 27; ++a)
        for (int b = 0; b < 27; ++b) {
            g.mul[a * 27 + b] = (uint8_t)((a && b) ? ex[(lg[a] + lg[b]) % 26] : 0);
            g.add[a * 27 + b] = (uint8_t)gadd(a, b);
        }
    for (int a = 0; a < 27; ++a) {
        g.inv[a] = (uint8_t)(a ? ex[(26 - lg[a]) % 26] : 0);
        g.neg[a] = (uint8_t)gneg(a);
        for (int st = 0; st < 3; ++st) {
            g.scr[st][a] = (uint8_t)gadd(a, 13 * st);
            g.dsc[st][a] = (uint8_t)gsub(a, 13 * st);
        }
    }
    for (int i = 0; i < 26; ++i) g.exp[i] = (uint8_t)ex[i];
    for (int a = 0; a < 32; ++a) g.lg[a] = (uint8_t)((a > 0 && a < 27) ? lg[a] : 255);
    for (int a = 0; a < 32; ++a) g.sqr[a] = 255;
    g.sqr[0] = 0;
    for (int e = 0; e < 26; e += 2) g.sqr[ex[e]] = (uint8_t)ex[e / 2]; // alpha^e is a square iff e is even
    for (int ki = 0; ki < 4; ++ki)
        for (int j = 0; j < 8; ++j)
            for (int m = 0; m < 8; ++m) g.syn[ki][j][m] = (uint8_t)gneg(ex[((j + 1) * (24 - 2 * ki + m)) % 26]);
    for (int j = 1; j <= 4; ++j)
        for (int v = 0; v < 27; ++v)
            for (int i = 0; i < 26; ++i) {
                const int s = (int)g.mul[v * 27 + ex[((26 - i) * j) % 26]];
                Tr t = split(s);
                for (int c = 0; c < 3; ++c) {
                    if (t.t[c]) T.ch.e[j - 1][v][i / 10] |= 1u << (3 * (i % 10) + c);
                    if (t.t[c] == 2) T.ch.e[j - 1][v][3 + i / 10] |= 1u << (3 * (i % 10) + c);
                }
            }
    // sanity: table multiply agrees with the polynomial product
    for (int a = 0; a < 27; ++a)
        for (int b = 0; b < 27; ++b)
            if (g.mul[a * 27 + b] != gmul_slow(a, b)) std::memset(&g, 0xFF, sizeof g);

    auto mul = [&](int a, int b) { return (int)g.mul[a * 27 + b]; };
    for (int ki = 0; ki < 4; ++ki) {
        const int k = 24 - 2 * ki, r = 26 - k;
        // g(x) = prod_{i=1..r} (x - alpha^i), low-first
        int gen[12] = {1}, n = 1;
        for (int i = 1; i <= r; ++i) {
            int nx[12] = {0};
            for (int j = 0; j < n; ++j) {
                nx[j] = gsub(nx[j], mul(gen[j], ex[i % 26]));
                nx[j + 1] = gadd(nx[j + 1], gen[j]);
            }
            ++n;
            std::memcpy(gen, nx, sizeof gen);
        }
        for (int j = 0; j <= r; ++j) T.rs.gen[ki][j] = (uint8_t)gen[j];
        // P[i][*] = parity of the unit vector e_i under each encoder variant (both are GF(27)-linear):
        //   as shipped (B1, OLD:522-533): low-first LFSR with coef=T[i], parity = +T[k+j]
        //   repaired (Appendix B):        coef=T[i]/g[0],                 parity = -T[k+j]
        for (int arith = 0; arith < 2; ++arith) {
            const int ig0 = g.inv[gen[0]];
            for (int i = 0; i < k; ++i) {
                int Tv[34] = {0};
                Tv[i] = 1;
                for (int q = 0; q < k; ++q) {
                    int coef = arith ? mul(Tv[q], ig0) : Tv[q];
                    if (!coef) continue;
                    for (int j = 0; j <= r; ++j) Tv[q + j] = gsub(Tv[q + j], mul(gen[j], coef));
                }
                for (int j = 0; j < r; ++j) T.rs.par[arith][ki][i][j] = (uint8_t)(arith ? gneg(Tv[k + j]) : Tv[k + j]);
            }
            RowTable& R = T.rs.row[arith][ki];
            for (int i = 0; i < 26; ++i)
                for (int d = 0; d < 27; ++d) {
                    int v[8] = {0};
                    if (i < k) for (int j = 0; j < r; ++j) v[j] = mul(d, T.rs.par[arith][ki][i][j]);
                    else v[i - k] = gneg(d);
                    R.e[i][d] = planes_of(v, r);
                    uint32_t nz = 0, two = 0;
                    for (int j = 0; j < r; ++j) { // nibbles for r <= 6 (PRMT selectors), dense 3-bit groups for r = 8 (24 trits fill bits 8..31)
                        Tr t = split(v[j]);
                        const int sh = r <= 6 ? 8 + 4 * j : 8 + 3 * j;
                        for (int c = 0; c < 3; ++c) {
                            if (t.t[c]) nz |= 1u << (sh + c);
                            if (t.t[c] == 2) two |= 1u << (sh + c);
                        }
                    }
                    T.rs.pl[arith][ki][i][d][0] = nz;
                    T.rs.pl[arith][ki][i][d][1] = two;
                }
        }
    }
}

// scrambler: st <- (a*st + b) % 3 evaluated in uint32 like the reference (OLD:83), so a,b >= 2^30
// wrap exactly as they do there.  The map on {0,1,2} need not be a bijection: up to 2 transient
// states, then a cycle whose length divides 6.
void scrambler_states(uint32_t a, uint32_t b, uint32_t s0, uint8_t st[8])
{
    uint32_t s = s0 % 3;
    for (int p = 0; p < 8; ++p) { s = (a * s + b) % 3; st[p] = (uint8_t)s; }
}

static uint8_t beacon_symbol(const t3c_config& c) // encode_beacon_symbol, OLD:107-113, payload of OLD:1130
{
    unsigned p = c.profile, s = (uint16_t)(c.superframe_words % 5) % 5;
    return (uint8_t)((p + 5 * s) % 27);
}

void make_geom(const t3c_config& c, size_t n_words, int arith, Geom& g)
{
    static const int ks[4] = {24, 22, 20, 18};
    std::memset(&g, 0, sizeof g);
    g.n_words = n_words;
    g.n_s = (26 * (uint64_t)n_words + 2) / 3;
    uint64_t tot = 0, use = 0;
    g.uniform_k = ks[c.uep[0] % 4];
    for (int b = 0; b < 9; ++b) {
        g.k[b] = ks[c.uep[b] % 4];
        if (g.k[b] != g.uniform_k) g.uniform_k = 0;
        g.s_b[b] = g.n_s > (uint64_t)b ? (g.n_s - b + 8) / 9 : 0;
        g.ncw[b] = g.s_b[b] / (uint64_t)g.k[b];
        g.cw_base[b] = tot;
        g.use_base[b] = use;
        tot += g.ncw[b];
        use += g.ncw[b] * (uint64_t)g.k[b];
    }
    g.n_cw = tot;
    g.l_body = 26 * tot;
    g.l_exp = g.l_body;
    g.slot = -1;
    if (use_beacon(c)) {
        const uint64_t P = c.beacon_period;
        const bool has = c.beacon_slot < 9;
        g.period = c.beacon_period;
        g.slot = has ? c.beacon_slot : -1;
        g.beacon_per = (uint32_t)(9 * P - 1);
        uint64_t W = 0;
        if (g.l_body) { // smallest W with 9W - ceil(W/P) >= l_body (A.5)
            W = has ? (g.l_body * P) / (9 * P - 1) : g.l_body / 9;
            W = W > 2 ? W - 2 : 0;
            while (9 * W - (has ? (W + P - 1) / P : 0) < g.l_body) ++W;
        }
        g.l_exp = 9 * W;
        g.bsym = beacon_symbol(c);
    }
    g.n_out = (52 + g.l_exp + 8) / 9;
    if (use_2d(c)) { g.tile_w = c.tile_w; g.tile_area = (uint64_t)c.tile_w * c.tile_h; }
    scrambler_states(c.seed_a, c.seed_b, c.seed_s0, g.st);
    g.arith = (uint8_t)(arith ? 1 : 0);
}

// decoder screen: a received block r = c (+) 13*st is a codeword iff sum_i T_i[r_i] == sum_i T_i[13*st_i]
// (the row tables are GF(3)-linear).  One constant per scrambler phase, plus the block at body index 0
// whose first two symbols may still be in the LCG's transient.
void fast_check_constants(const HostTables& H, const Geom& g, uint32_t chk_nz[7], uint32_t chk_two[7])
{
    if (!g.uniform_k) { for (int i = 0; i < 7; ++i) chk_nz[i] = chk_two[i] = 0; return; }
    const RowTable& R = H.rs.row[1][kidx_of(g.uniform_k)];
    for (int ph = 0; ph < 7; ++ph) {
        int trit[32] = {0};
        for (int i = 0; i < 26; ++i) {
            int st;
            if (ph == 6) st = i < 2 ? g.st[i] : g.st[2 + (i - 2) % 6];
            else st = g.st[2 + (ph + i) % 6];
            const uint64_t e = R.e[i][13 * st];
            for (int bit = 0; bit < 32; ++bit) {
                const int v = ((e >> bit) & 1) + ((e >> (32 + bit)) & 1); // 0,1,2
                trit[bit] = (trit[bit] + v) % 3;
            }
        }
        uint32_t nz = 0, two = 0;
        for (int bit = 0; bit < 32; ++bit) { if (trit[bit]) nz |= 1u << bit; if (trit[bit] == 2) two |= 1u << bit; }
        chk_nz[ph] = nz; chk_two[ph] = two;
    }
}

size_t profile_words(const t3c_config& c, size_t n_words)
{
    if (c.profile == T3C_PROFILE_RAW) return n_words;
    Geom g;
    make_geom(c, n_words, 0, g);
    return (size_t)g.n_out;
}

bool geom_from_nout(const t3c_config& c, size_t n_out, int arith, Geom& g)
{
    size_t lo = 0, hi = n_out + 16;
    while (profile_words(c, hi) < n_out) hi *= 2;
    while (lo < hi) {
        size_t mid = lo + (hi - lo) / 2;
        if (profile_words(c, mid) < n_out) lo = mid + 1; else hi = mid;
    }
    make_geom(c, lo, arith, g);
    return g.n_out == n_out;
}

} // namespace t3c
