/*
 * t3_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, see t3_oracle.h).
 *
 * Plain-C restatement of the reference hot path.  "OLD:n" cites
 * /root/reference/old/include/ternary_image_codec_v6_min.hpp line n;
 * "IMG:n" cites /root/reference/old/include/io_image.hpp line n.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (no -march=native: FMA
 * contraction changes 859 of 2^24*3 bridge outputs).
 */
#include "t3_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* Trit / symbol primitives, OLD:20-31                                 */
/* ------------------------------------------------------------------ */
static inline uint8_t pack3(unsigned a, unsigned b, unsigned c) { return (uint8_t)(a + 3 * b + 9 * c); } /* OLD:24-27 */
static inline void unpack3(uint8_t s, uint8_t d[3]) { d[0] = s % 3; d[1] = (s / 3) % 3; d[2] = (s / 9) % 3; } /* OLD:28-31 */

/* ------------------------------------------------------------------ */
/* GF(27) = GF(3)[x]/(x^3+2x+1), OLD:383-487                           */
/* ------------------------------------------------------------------ */
uint8_t t3o_gf_add(uint8_t a, uint8_t b) /* OLD:383-388 */
{
    unsigned a0 = a % 3, a1 = (a / 3) % 3, a2 = (a / 9) % 3;
    unsigned b0 = b % 3, b1 = (b / 3) % 3, b2 = (b / 9) % 3;
    return (uint8_t)(((a0 + b0) % 3) + 3 * ((a1 + b1) % 3) + 9 * ((a2 + b2) % 3));
}
uint8_t t3o_gf_sub(uint8_t a, uint8_t b) /* OLD:389-401 */
{
    int a0 = a % 3, a1 = (a / 3) % 3, a2 = (a / 9) % 3;
    int b0 = b % 3, b1 = (b / 3) % 3, b2 = (b / 9) % 3;
    return (uint8_t)(((a0 - b0 + 3) % 3) + 3 * ((a1 - b1 + 3) % 3) + 9 * ((a2 - b2 + 3) % 3));
}
static uint8_t gf_mul_poly(uint8_t a, uint8_t b) /* OLD:402-413 */
{
    if (a == 0 || b == 0) return 0;
    int a0 = a % 3, a1 = (a / 3) % 3, a2 = (a / 9) % 3;
    int b0 = b % 3, b1 = (b / 3) % 3, b2 = (b / 9) % 3;
    int r0 = (a0 * b0) % 3, r1 = (a0 * b1 + a1 * b0) % 3, r2 = (a0 * b2 + a1 * b1 + a2 * b0) % 3;
    int r3 = (a1 * b2 + a2 * b1) % 3, r4 = (a2 * b2) % 3;
    /* x^3 = x + 2 (i.e. -2x-1 mod 3), x^4 = x^2 + 2x */
    r1 = (r1 + r3) % 3;
    r0 = (r0 + 2 * r3) % 3;
    r2 = (r2 + r4) % 3;
    r1 = (r1 + 2 * r4) % 3;
    return (uint8_t)(r0 + 3 * r1 + 9 * r2);
}

static struct {
    int     ready;
    uint8_t exp[78];      /* OLD:416 */
    int16_t log[27];      /* OLD:417 */
    uint8_t mul[729];     /* OLD:418 */
    uint8_t inv[27];      /* OLD:419 */
    uint8_t prim;
    uint8_t g[4][9];      /* generator polys for k=24,22,20,18 (index (24-k)/2), low-first */
} G;

static int order_of(uint8_t g) /* OLD:425-435 */
{
    if (g == 0 || g == 1) return -1;
    uint8_t x = 1;
    for (int i = 1; i <= 26; ++i) { x = gf_mul_poly(x, g); if (x == 1) return i; }
    return -1;
}
static void build_gen(int k, uint8_t* g); /* fwd */
static void gf_init(void) /* OLD:436-466 */
{
    if (G.ready) return;
    uint8_t prim = 0;
    for (uint8_t c = 2; c < 27; ++c) if (order_of(c) == 26) { prim = c; break; }
    if (prim == 0) prim = 3;
    G.prim = prim;
    for (int i = 0; i < 27; ++i) G.log[i] = -1;
    G.exp[0] = 1; G.log[1] = 0;
    for (int i = 1; i < 26; ++i) { G.exp[i] = gf_mul_poly(G.exp[i - 1], prim); G.log[G.exp[i]] = (int16_t)i; }
    for (int i = 26; i < 78; ++i) G.exp[i] = G.exp[i - 26];
    for (int a = 0; a < 27; ++a) for (int b = 0; b < 27; ++b) G.mul[a * 27 + b] = gf_mul_poly((uint8_t)a, (uint8_t)b);
    G.inv[0] = 0;
    for (int a = 1; a < 27; ++a) G.inv[a] = G.exp[(26 - G.log[a]) % 26];
    G.ready = 1;
    for (int i = 0; i < 4; ++i) build_gen(24 - 2 * i, G.g[i]);
}
/* Symbols reaching the tables are always < 27 on valid input; the reference
 * indexes tab.mul[a*27+b] unchecked (OLD:475).  We reduce mod 27 so that
 * out-of-contract bytes are at least deterministic. */
uint8_t t3o_gf_mul(uint8_t a, uint8_t b) { gf_init(); return G.mul[(a % 27) * 27 + (b % 27)]; }
uint8_t t3o_gf_inv(uint8_t a) { gf_init(); return G.inv[a % 27]; }
uint8_t t3o_gf_pow_alpha(int e) { gf_init(); int m = (e % 26 + 26) % 26; return G.exp[m]; } /* OLD:479-482 */
int t3o_gf_log(uint8_t a) { gf_init(); return G.log[a % 27]; }
#define MUL(a, b) (G.mul[(a) * 27 + (b)])

/* ------------------------------------------------------------------ */
/* RS(26,k), OLD:490-663                                               */
/* ------------------------------------------------------------------ */
static void build_gen(int k, uint8_t* g) /* OLD:501-516: g(x)=prod_{i=1..r}(x-alpha^i), low-first */
{
    int r = 26 - k, n = 1;
    uint8_t cur[10] = {1}, nx[10];
    for (int i = 1; i <= r; ++i) {
        memset(nx, 0, sizeof nx);
        uint8_t root = G.exp[i % 26];
        for (int j = 0; j < n; ++j) {
            nx[j] = t3o_gf_sub(nx[j], MUL(cur[j], root));
            nx[j + 1] = t3o_gf_add(nx[j + 1], cur[j]);
        }
        ++n;
        memcpy(cur, nx, sizeof cur);
    }
    memcpy(g, cur, (size_t)(r + 1));
}
static const uint8_t* gen_for(int k) { gf_init(); return G.g[(24 - k) / 2]; }
static int k_valid(int k) { return k == 24 || k == 22 || k == 20 || k == 18; }

int t3o_rs_gen(int k, uint8_t* g_out)
{
    if (!k_valid(k)) return -1;
    memcpy(g_out, gen_for(k), (size_t)(26 - k + 1));
    return 26 - k + 1;
}

void t3o_rs_encode(int k, int fixed, const uint8_t* data_k, uint8_t* out26) /* OLD:517-535 */
{
    const uint8_t* g = gen_for(k);
    int r = 26 - k;
    uint8_t T[26];
    memset(T, 0, sizeof T);
    for (int i = 0; i < k; ++i) T[i] = data_k[i];
    for (int i = 0; i < k; ++i) {
        /* OLD:524 as shipped: coef=T[i]; Appendix-B repair: coef=T[i]*inv(g[0]) */
        uint8_t coef = fixed ? MUL(T[i], G.inv[g[0]]) : T[i];
        if (coef == 0) continue;
        for (int j = 0; j <= r; ++j) T[i + j] = t3o_gf_sub(T[i + j], MUL(g[j], coef));
    }
    for (int i = 0; i < k; ++i) out26[i] = data_k[i];
    /* OLD:533 as shipped: +T[k+i]; repaired: 0-T[k+i] */
    for (int i = 0; i < r; ++i) out26[k + i] = fixed ? t3o_gf_sub(0, T[k + i]) : T[k + i];
}

#define VMAX 40
/* strict: the acceptance rule of the consistent (FIXED) frame decoder, which has no reference to match: reject unless L <= t,
 * deg(sigma) == L and sigma has L distinct roots (with more than t errors OLD:611-624 accepts locators that locate nothing and
 * "corrects" 0..t symbols of a block it cannot decode).  strict = 0 is decode_block as shipped / as repaired, bit for bit. */
static int rs_decode_ex(int k, int fixed, int strict, uint8_t* c, uint8_t* out_k) /* OLD:546-662 */
{
    gf_init();
    const int n = 26, r = n - k, t = r / 2;
    uint8_t S[8];
    int all0 = 1;
    for (int j = 0; j < r; ++j) { /* OLD:551-561 */
        uint8_t acc = 0;
        for (int i = 0; i < n; ++i) acc = t3o_gf_add(acc, MUL(c[i], G.exp[((j + 1) * i) % 26]));
        S[j] = acc;
        if (acc) all0 = 0;
    }
    if (all0) { for (int i = 0; i < k; ++i) out_k[i] = c[i]; return 1; } /* OLD:562-566 */
    /* Berlekamp-Massey with explicit vector sizes, OLD:567-605 */
    uint8_t sigma[VMAX] = {1}, B[VMAX] = {1};
    int ns = 1, nb = 1, L = 0, m = 1;
    for (int nS = 0; nS < r; ++nS) {
        uint8_t delta = S[nS];
        for (int i = 1; i <= L; ++i) if (i < ns) delta = t3o_gf_add(delta, MUL(sigma[i], S[nS - i]));
        if (delta != 0) {
            uint8_t T[VMAX]; int nt = ns;
            memcpy(T, sigma, sizeof T);
            uint8_t x[VMAX]; int nx = m + nb;
            memset(x, 0, sizeof x);
            for (int i = 0; i < nb; ++i) x[m + i] = MUL(delta, B[i]);
            int nd = ns > nx ? ns : nx;
            uint8_t s2[VMAX];
            memset(s2, 0, sizeof s2);
            for (int i = 0; i < nd; ++i) {
                uint8_t a = i < ns ? sigma[i] : 0, b = i < nx ? x[i] : 0;
                s2[i] = t3o_gf_sub(a, b);
            }
            memcpy(sigma, s2, sizeof sigma); ns = nd;
            if (2 * L <= nS) {
                uint8_t invd = G.inv[delta];
                memset(B, 0, sizeof B);
                for (int i = 0; i < nt; ++i) B[i] = MUL(T[i], invd);
                nb = nt; L = nS + 1 - L; m = 1;
            } else m += 1;
        } else m += 1;
    }
    if (strict) {
        int deg = ns - 1;
        while (deg > 0 && sigma[deg] == 0) --deg;
        if (L > t || deg != L) return 0;
    }
    /* Omega = (S * sigma) mod x^r, OLD:606-610 */
    uint8_t Om[VMAX + 9];
    memset(Om, 0, sizeof Om);
    for (int i = 0; i < r; ++i) for (int j = 0; j < ns; ++j) Om[i + j] = t3o_gf_add(Om[i + j], MUL(S[i], sigma[j]));
    int nom = (r + 1) + ns - 1; if (nom > r) nom = r;
    /* Chien, OLD:611-624 */
    int pos[26], np = 0;
    for (int i = 0; i < n; ++i) {
        uint8_t x = t3o_gf_pow_alpha((-i) % 26), acc = 0;
        for (int d = ns - 1; d >= 0; --d) acc = t3o_gf_add(MUL(acc, x), sigma[d]);
        if (acc == 0) pos[np++] = i;
    }
    if (np > t || (strict && np != L)) return 0;
    /* formal derivative in char 3, OLD:625-641 */
    uint8_t sp[VMAX]; int nsp = ns > 1 ? ns - 1 : 1;
    memset(sp, 0, sizeof sp);
    for (int i = 1; i < ns; ++i) {
        int im = i % 3;
        if (im == 0) sp[i - 1] = 0;
        else if (im == 1) sp[i - 1] = sigma[i];
        else { uint8_t a = sigma[i]; sp[i - 1] = (uint8_t)(((2 * (a % 3)) % 3) + 3 * ((2 * ((a / 3) % 3)) % 3) + 9 * ((2 * ((a / 9) % 3)) % 3)); }
    }
    /* Forney, OLD:642-659 */
    for (int e = 0; e < np; ++e) {
        uint8_t Xin = t3o_gf_pow_alpha((-pos[e]) % 26), num = 0, den = 0;
        for (int d = nom - 1; d >= 0; --d) num = t3o_gf_add(MUL(num, Xin), Om[d]);
        for (int d = nsp - 1; d >= 0; --d) den = t3o_gf_add(MUL(den, Xin), sp[d]);
        if (den == 0) return 0;
        uint8_t mag = MUL(t3o_gf_sub(0, num), G.inv[den]);
        /* OLD:658 as shipped: add; Appendix-B repair: sub */
        c[pos[e]] = fixed ? t3o_gf_sub(c[pos[e]], mag) : t3o_gf_add(c[pos[e]], mag);
    }
    for (int i = 0; i < k; ++i) out_k[i] = c[i];
    return 1;
}

int t3o_rs_decode(int k, int fixed, uint8_t* c, uint8_t* out_k) { return rs_decode_ex(k, fixed, 0, c, out_k); }

void t3o_rs_encode_blocks(int k, int fixed, const uint8_t* data, size_t nblk, uint8_t* out)
{
    for (size_t i = 0; i < nblk; ++i) t3o_rs_encode(k, fixed, data + i * (size_t)k, out + i * 26);
}
void t3o_rs_decode_blocks(int k, int fixed, uint8_t* inout, size_t nblk, uint8_t* out, uint8_t* ok)
{
    for (size_t i = 0; i < nblk; ++i) {
        int f = t3o_rs_decode(k, fixed, inout + i * 26, out + i * (size_t)k);
        if (!f) memset(out + i * (size_t)k, 0, (size_t)k);
        ok[i] = (uint8_t)f;
    }
}

/* ------------------------------------------------------------------ */
/* Config defaults, profile -> k                                        */
/* ------------------------------------------------------------------ */
void t3o_cfg_default(t3o_cfg* c) /* OLD:862-873, 898 */
{
    memset(c, 0, sizeof *c);
    c->profile = 1;
    for (int i = 0; i < 9; ++i) c->uep[i] = 1;
    c->seed_a = c->seed_b = c->seed_s0 = 1;
    c->superframe_words = 8192;
    c->subword = 27;
    c->centered = 1;
}
static int k_for_band(const t3o_cfg* c, int b) /* OLD:1089-1100, 966-980 */
{
    static const int ks[4] = {24, 22, 20, 18};
    return ks[c->uep[b] % 4];
}

/* ------------------------------------------------------------------ */
/* Header + ternary CRC-12, OLD:155-380                                 */
/* ------------------------------------------------------------------ */
void t3o_crc12(const uint8_t* msg, size_t n, uint8_t out[12]) /* OLD:176-205 */
{
    uint8_t r[12] = {0};
    for (size_t q = 0; q < n + 12; ++q) {
        uint8_t in = q < n ? msg[q] : 0;
        uint8_t fb = (uint8_t)((in + r[11]) % 3), nx[12];
        nx[0] = fb; nx[1] = r[0]; nx[2] = r[1];
        nx[3] = (uint8_t)((r[2] + fb) % 3); nx[4] = (uint8_t)((r[3] + fb) % 3);
        nx[5] = r[4]; nx[6] = r[5]; nx[7] = (uint8_t)((r[6] + fb) % 3);
        nx[8] = r[7]; nx[9] = r[8]; nx[10] = r[9]; nx[11] = r[10];
        memcpy(r, nx, 12);
    }
    memcpy(out, r, 12);
}
static void header_crc(const uint8_t sym[27], uint8_t r[12]) /* OLD:269-279, 292-302 */
{
    uint8_t tr[81]; size_t n = 0;
    for (int i = 0; i < 27; ++i) {
        if (i == 20 || i == 21 || i == 22 || i == 26) continue;
        unpack3(sym[i], tr + n); n += 3;
    }
    t3o_crc12(tr, n, r);
}
void t3o_header_pack(const t3o_cfg* c, uint32_t frame_seq, uint32_t band_map_hash, uint8_t p[27]) /* OLD:208-289 */
{
    const uint16_t magic = 0x0A2; const uint8_t version = 1;
    memset(p, 0, 27);
#define AT(i, v) p[i] = (uint8_t)(((uint8_t)(v)) % 27) /* at(): arg narrows to GF27 first, OLD:211-214 */
    AT(0, magic % 27); AT(1, (magic / 27) % 27); AT(2, version % 27); AT(3, c->profile);
    uint32_t u0 = 0, u1 = 0, u2 = 0;
    for (int i = 0; i < 3; ++i) u0 = u0 * 3 + (c->uep[i] % 3);
    for (int i = 3; i < 6; ++i) u1 = u1 * 3 + (c->uep[i] % 3);
    for (int i = 6; i < 9; ++i) u2 = u2 * 3 + (c->uep[i] % 3);
    AT(4, u0); AT(5, u1); AT(6, u2);
    AT(7, c->tile_w % 27); AT(8, c->tile_h % 27);
    AT(9, c->seed_a % 27); AT(10, c->seed_b % 27); AT(11, c->seed_s0 % 27);
    uint8_t sub = 0;
    switch (c->subword) { case 27: sub = 0; break; case 24: sub = 1; break; case 21: sub = 2; break;
                          case 18: sub = 3; break; case 15: sub = 4; break; default: sub = 0; }
    AT(12, (sub + 9 * (c->centered ? 1 : 0)) % 27);
    AT(13, band_map_hash % 27); AT(14, (band_map_hash / 27) % 27); AT(15, (band_map_hash / 729) % 27);
    AT(16, c->coset % 3);
    AT(17, frame_seq % 27); AT(18, (frame_seq / 27) % 27); AT(19, (frame_seq / 729) % 27);
    AT(23, c->beacon_enabled ? 1 : 0); AT(24, c->beacon_slot % 27);
    AT(25, c->beacon_period < 26 ? c->beacon_period : 26);
    uint8_t r[12];
    header_crc(p, r);
    AT(20, pack3(r[0], r[1], r[2])); AT(21, pack3(r[3], r[4], r[5]));
    AT(22, pack3(r[6], r[7], r[8])); AT(26, pack3(r[9], r[10], r[11]));
#undef AT
}
int t3o_header_check(const uint8_t p[27]) /* OLD:290-316 */
{
    uint8_t r[12], h[12];
    header_crc(p, r);
    unpack3(p[20], h); unpack3(p[21], h + 3); unpack3(p[22], h + 6); unpack3(p[26], h + 9);
    return memcmp(r, h, 12) == 0;
}
void t3o_header_unpack(const uint8_t p[27], t3o_cfg* h, uint32_t* frame_seq, uint32_t* band_map_hash,
                       uint16_t* magic, uint8_t* version) /* OLD:317-379 */
{
#define RD(i) ((uint32_t)(p[i] % 27))
    if (magic) *magic = (uint16_t)(RD(0) + 27 * RD(1));
    if (version) *version = (uint8_t)RD(2);
    h->profile = (uint8_t)(RD(3) % 5);
    for (int g = 0; g < 3; ++g) { /* dec3: LSB-first (bug B7), OLD:327-340 */
        uint32_t v = RD(4 + g);
        h->uep[3 * g + 0] = (uint8_t)(v % 3); v /= 3;
        h->uep[3 * g + 1] = (uint8_t)(v % 3); v /= 3;
        h->uep[3 * g + 2] = (uint8_t)(v % 3);
    }
    h->tile_w = (uint16_t)RD(7); h->tile_h = (uint16_t)RD(8);
    h->seed_a = RD(9); h->seed_b = RD(10); h->seed_s0 = RD(11);
    {
        uint32_t v = RD(12) % 27; uint8_t cen = (uint8_t)((v / 9) % 3), sub = (uint8_t)(v % 9);
        static const uint8_t modes[5] = {27, 24, 21, 18, 15};
        h->subword = sub < 5 ? modes[sub] : 27;
        h->centered = cen != 0;
    }
    if (band_map_hash) *band_map_hash = RD(13) + 27 * RD(14) + 729 * RD(15);
    h->coset = (uint8_t)(RD(16) % 3);
    if (frame_seq) *frame_seq = RD(17) + 27 * RD(18) + 729 * RD(19);
    h->beacon_enabled = RD(23) != 0;
    h->beacon_slot = (uint8_t)(RD(24) % 9);
    h->beacon_period = RD(25);
#undef RD
}

/* ------------------------------------------------------------------ */
/* 2 px <-> Word27, OLD:665-747                                         */
/* ------------------------------------------------------------------ */
static void i2tr(uint32_t v, int w, uint8_t* d, int s) { for (int i = 0; i < w; ++i) { d[s + i] = (uint8_t)(v % 3); v /= 3; } } /* OLD:675-682 */
static uint32_t tr2i(const uint8_t* d, int w, int s) { uint32_t val = 0, p = 1; for (int i = 0; i < w; ++i) { val += p * d[s + i]; p *= 3; } return val; } /* OLD:683-692 */
static void pack_two(const t3o_pixel* a, const t3o_pixel* b, uint8_t* w) /* OLD:693-705 */
{
    uint8_t T[27];
    memset(T, 0, sizeof T);
    i2tr(a->Yq, 5, T, 0); i2tr((uint32_t)(a->Cbq + 40), 4, T, 5); i2tr((uint32_t)(a->Crq + 40), 4, T, 9);
    i2tr(b->Yq, 5, T, 13); i2tr((uint32_t)(b->Cbq + 40), 4, T, 18); i2tr((uint32_t)(b->Crq + 40), 4, T, 22);
    T[26] = 0;
    for (int s = 0; s < 9; ++s) w[s] = pack3(T[3 * s], T[3 * s + 1], T[3 * s + 2]);
}
static void unpack_two(const uint8_t* w, t3o_pixel* a, t3o_pixel* b) /* OLD:706-722 */
{
    uint8_t T[27];
    for (int s = 0; s < 9; ++s) unpack3(w[s], T + 3 * s);
    a->Yq = (uint16_t)tr2i(T, 5, 0); a->Cbq = (int16_t)((int16_t)tr2i(T, 4, 5) - 40); a->Crq = (int16_t)((int16_t)tr2i(T, 4, 9) - 40);
    b->Yq = (uint16_t)tr2i(T, 5, 13); b->Cbq = (int16_t)((int16_t)tr2i(T, 4, 18) - 40); b->Crq = (int16_t)((int16_t)tr2i(T, 4, 22) - 40);
}
size_t t3o_pack_pixels(const t3o_pixel* px, size_t n, uint8_t* words) /* OLD:723-734 */
{
    size_t nw = 0;
    const t3o_pixel zero = {0, 0, 0};
    for (size_t i = 0; i < n; i += 2) pack_two(&px[i], i + 1 < n ? &px[i + 1] : &zero, words + 9 * nw++);
    return nw;
}
void t3o_unpack_pixels(const uint8_t* words, size_t nw, t3o_pixel* px) /* OLD:735-747 */
{
    for (size_t i = 0; i < nw; ++i) unpack_two(words + 9 * i, &px[2 * i], &px[2 * i + 1]);
}

/* ------------------------------------------------------------------ */
/* 2D boustrophedon interleave, OLD:749-813                             */
/* ------------------------------------------------------------------ */
static void boustro(uint8_t* sy, size_t n, unsigned w, unsigned h, int inverse)
{
    if (!w || !h) return;
    size_t A = (size_t)w * h;
    uint8_t* out = (uint8_t*)malloc(n ? n : 1);
    size_t i = 0;
    while (i < n) {
        size_t take = A < n - i ? A : n - i, k = 0;
        for (unsigned r = 0; r < h; ++r) {
            if (r % 2 == 0) {
                for (unsigned c = 0; c < w && (size_t)r * w + c < take; ++c) {
                    size_t idx = (size_t)r * w + c;
                    if (!inverse) out[i + k++] = sy[i + idx]; else out[i + idx] = sy[i + k++];
                }
            } else {
                for (int c = (int)w - 1; c >= 0; --c) {
                    size_t idx = (size_t)r * w + (size_t)c;
                    if (idx < take) { if (!inverse) out[i + k++] = sy[i + idx]; else out[i + idx] = sy[i + k++]; }
                }
            }
        }
        i += take;
    }
    memcpy(sy, out, n);
    free(out);
}
void t3o_interleave2d(uint8_t* sy, size_t n, unsigned w, unsigned h) { boustro(sy, n, w, h, 0); }   /* OLD:750-780 */
void t3o_deinterleave2d(uint8_t* sy, size_t n, unsigned w, unsigned h) { boustro(sy, n, w, h, 1); } /* OLD:781-813 */

/* ------------------------------------------------------------------ */
/* Scrambler / beacon, OLD:77-113                                       */
/* ------------------------------------------------------------------ */
uint8_t t3o_scramble_symbol(uint8_t s, uint32_t a, uint32_t b, uint32_t* st) /* OLD:81-87 */
{
    *st = ((a * *st) + b) % 3; /* uint32 wrap-around, as the reference */
    uint8_t d[3]; unpack3(s, d);
    for (int i = 0; i < 3; ++i) d[i] = (uint8_t)((d[i] + *st) % 3);
    return pack3(d[0], d[1], d[2]);
}
uint8_t t3o_descramble_symbol(uint8_t s, uint32_t a, uint32_t b, uint32_t* st) /* OLD:88-94 */
{
    *st = ((a * *st) + b) % 3;
    uint8_t d[3]; unpack3(s, d);
    for (int i = 0; i < 3; ++i) d[i] = (uint8_t)((3 + d[i] - (*st % 3)) % 3);
    return pack3(d[0], d[1], d[2]);
}
uint8_t t3o_beacon_symbol(uint8_t profile, uint16_t frame_seq_mod, uint8_t health) /* OLD:107-113 */
{
    uint8_t p = profile, s = (uint8_t)(frame_seq_mod % 5), h = (uint8_t)(health % 3);
    return (uint8_t)((p + 5 * s + 15 * h) % 27);
}

/* ------------------------------------------------------------------ */
/* Profile encoder, OLD:1043-1169                                       */
/* ------------------------------------------------------------------ */
static int use_2d(const t3o_cfg* c) { return c->profile == 4 && c->tile_w && c->tile_h; } /* OLD:1083 */
static int use_beacon(const t3o_cfg* c) { return c->beacon_enabled && c->beacon_period > 0; } /* OLD:1118 */

typedef struct {
    size_t n_s;          /* regrouped symbols = ceil(26*N_w/3) */
    size_t s_b[9];       /* band lengths */
    int    k_b[9];
    size_t ncw_b[9];     /* codewords per band */
    size_t cw_base[9];   /* codewords before band b */
    size_t l_body;       /* 26 * sum ncw */
    size_t l_exp;        /* after beacon expansion */
    size_t n_out;        /* output words */
} geom_t;

static void geometry(const t3o_cfg* c, size_t n_words, geom_t* g)
{
    memset(g, 0, sizeof *g);
    g->n_s = (26 * n_words + 2) / 3;
    size_t tot = 0;
    for (int b = 0; b < 9; ++b) {
        g->s_b[b] = g->n_s > (size_t)b ? (g->n_s - (size_t)b + 8) / 9 : 0;
        g->k_b[b] = k_for_band(c, b);
        g->ncw_b[b] = g->s_b[b] / (size_t)g->k_b[b];
        g->cw_base[b] = tot;
        tot += g->ncw_b[b];
    }
    g->l_body = 26 * tot;
    g->l_exp = g->l_body;
    if (use_beacon(c)) {
        /* emission loop OLD:1122-1139: words are emitted while body symbols remain; a word with
         * wd%P==0 carries 8 body symbols (if band_slot<9) else 9.  W = smallest count with
         * 9W - beacons(W) >= l_body, beacons(W) = ceil(W/P). */
        size_t P = c->beacon_period, W = 0;
        int has = c->beacon_slot < 9;
        if (g->l_body) {
            W = has ? (g->l_body * P) / (9 * P - 1) : g->l_body / 9;
            if (W > 2) W -= 2; else W = 0;
            while (9 * W - (has ? (W + P - 1) / P : 0) < g->l_body) ++W;
        }
        g->l_exp = 9 * W;
    }
    g->n_out = (52 + g->l_exp + 8) / 9;
}
size_t t3o_profile_words_bound(const t3o_cfg* c, size_t n_words)
{
    if (c->profile == T3O_PROFILE_RAW) return n_words;
    geom_t g; geometry(c, n_words, &g);
    return g.n_out;
}

/* regroup: first 26 trits of every word -> symbols of 3 trits, OLD:1051-1082 */
static uint8_t* regroup(const uint8_t* raw9, size_t n_words, size_t* n_s)
{
    size_t ntr = 26 * n_words, ns = (ntr + 2) / 3;
    uint8_t* tr = (uint8_t*)calloc(3 * ns + 3, 1);
    for (size_t w = 0; w < n_words; ++w) {
        uint8_t T[27];
        for (int s = 0; s < 9; ++s) unpack3(raw9[9 * w + (size_t)s], T + 3 * s);
        memcpy(tr + 26 * w, T, 26);
    }
    uint8_t* sy = (uint8_t*)malloc(ns ? ns : 1);
    for (size_t j = 0; j < ns; ++j) sy[j] = pack3(tr[3 * j], tr[3 * j + 1], tr[3 * j + 2]);
    free(tr);
    *n_s = ns;
    return sy;
}

static void header_encode(const t3o_cfg* c, int fixed, uint8_t out52[52]) /* OLD:1142-1158 */
{
    uint8_t hp[27], A[18], B[18];
    t3o_header_pack(c, 0, 0, hp);
    memcpy(A, hp, 18);
    memset(B, 0, 18); memcpy(B, hp + 18, 9);
    t3o_rs_encode(18, fixed, A, out52);
    t3o_rs_encode(18, fixed, B, out52 + 26);
}

size_t t3o_encode_profile(const t3o_cfg* c, int fixed, const uint8_t* raw9, size_t n_words, uint8_t* out9, size_t cap)
{
    gf_init();
    if (c->profile == T3O_PROFILE_RAW) { /* OLD:1046-1050 */
        if (cap < n_words) return (size_t)-1;
        memcpy(out9, raw9, 9 * n_words);
        return n_words;
    }
    geom_t g; geometry(c, n_words, &g);
    if (cap < g.n_out) return (size_t)-1;
    size_t ns; uint8_t* sy = regroup(raw9, n_words, &ns);
    if (use_2d(c)) t3o_interleave2d(sy, ns, c->tile_w, c->tile_h); /* OLD:1083-1086 */
    /* band split + RS, OLD:1087-1115 */
    uint8_t* body = (uint8_t*)malloc(g.l_body ? g.l_body : 1);
    size_t o = 0;
    uint8_t* band = (uint8_t*)malloc(g.s_b[0] ? g.s_b[0] : 1);
    for (int b = 0; b < 9; ++b) {
        size_t L = 0;
        for (size_t i = (size_t)b; i < ns; i += 9) band[L++] = sy[i];
        int k = g.k_b[b];
        for (size_t j = 0; j + (size_t)k <= L; j += (size_t)k) { t3o_rs_encode(k, fixed, band + j, body + o); o += 26; }
    }
    free(band); free(sy);
    /* scramble, OLD:1116-1117 */
    uint32_t st = c->seed_s0 % 3;
    for (size_t p = 0; p < g.l_body; ++p) body[p] = t3o_scramble_symbol(body[p], c->seed_a, c->seed_b, &st);
    /* beacon, OLD:1118-1141 */
    uint8_t* all = (uint8_t*)calloc(9 * g.n_out + 9, 1);
    header_encode(c, fixed, all);
    if (use_beacon(c)) {
        uint8_t bs = t3o_beacon_symbol(c->profile, (uint16_t)(c->superframe_words % 5), 0);
        size_t k = 0, wd = 0, q = 52;
        while (k < g.l_body) {
            int ins = (wd % c->beacon_period) == 0;
            for (int slot = 0; slot < 9; ++slot) {
                if (ins && slot == c->beacon_slot) all[q++] = bs;
                else all[q++] = k < g.l_body ? body[k++] : 0;
            }
            ++wd;
        }
    } else memcpy(all + 52, body, g.l_body);
    free(body);
    memcpy(out9, all, 9 * g.n_out); /* words of 9 symbols, zero padded, OLD:1164-1167 */
    free(all);
    return g.n_out;
}

/* ------------------------------------------------------------------ */
/* Reference decoder as shipped, OLD:918-1041              */
/* ------------------------------------------------------------------ */
int t3o_decode_profile_ref(t3o_cfg* seen, const uint8_t* in9, size_t n, uint8_t* out9, size_t cap, size_t* n_out)
{
    gf_init();
    *n_out = 0;
    if (seen->profile == T3O_PROFILE_RAW) { /* OLD:998-1002 */
        if (cap < n) return 0;
        memcpy(out9, in9, 9 * n); *n_out = n; return 1;
    }
    if (n < 6) return 0; /* OLD:920 */
    uint8_t A[26], B[26], a18[18], b18[18], hp[27];
    memcpy(A, in9, 26); memcpy(B, in9 + 26, 26); /* OLD:925-927: symbols 0..51 of the first 54 */
    if (!t3o_rs_decode(18, 0, A, a18)) return 0;
    if (!t3o_rs_decode(18, 0, B, b18)) return 0;
    memcpy(hp, a18, 18); memcpy(hp + 18, b18, 9);
    if (!t3o_header_check(hp)) return 0;
    t3o_cfg h = *seen;
    t3o_header_unpack(hp, &h, NULL, NULL, NULL, NULL);
    *seen = h; /* OLD:1006-1013 (mutated before the body is decoded) */
    size_t nb = n - 6, nsym = 9 * nb;
    uint8_t* sy = (uint8_t*)malloc(nsym ? nsym : 1);
    uint32_t st = h.seed_s0 % 3; /* OLD:943-944 */
    for (size_t i = 0; i < nsym; ++i) sy[i] = t3o_descramble_symbol(in9[54 + i], h.seed_a, h.seed_b, &st);
    /* demap slot-major, OLD:950-961 */
    int skip = h.beacon_enabled && h.beacon_period > 0;
    uint8_t* use = (uint8_t*)malloc(nsym ? nsym : 1);
    size_t nuse = 0;
    uint8_t* band = (uint8_t*)malloc(nb ? nb : 1);
    int ok = 1;
    for (int b = 0; b < 9 && ok; ++b) {
        size_t L = 0;
        for (size_t wi = 0; wi < nb; ++wi) {
            if (skip && (wi % h.beacon_period) == 0 && b == h.beacon_slot) continue;
            band[L++] = sy[9 * wi + (size_t)b];
        }
        int k = k_for_band(&h, b);
        for (size_t j = 0; j + 26 <= L; j += 26) { /* OLD:983-990 */
            uint8_t nbuf[26], kbuf[26];
            memcpy(nbuf, band + j, 26);
            if (!t3o_rs_decode(k, 0, nbuf, kbuf)) { ok = 0; break; }
            memcpy(use + nuse, kbuf, (size_t)k); nuse += (size_t)k;
        }
    }
    free(band); free(sy);
    if (!ok) { free(use); return 0; }
    if (h.profile == 4 && h.tile_w && h.tile_h) t3o_deinterleave2d(use, nuse, h.tile_w, h.tile_h); /* OLD:1018-1021 */
    /* symbols -> trits -> groups of 26 -> Word27, OLD:1022-1039 */
    size_t nwords = (3 * nuse) / 26;
    if (cap < nwords) { free(use); return 0; }
    for (size_t w = 0; w < nwords; ++w) {
        uint8_t T[27];
        for (int i = 0; i < 26; ++i) { size_t ti = 26 * w + (size_t)i; uint8_t d[3]; unpack3(use[ti / 3], d); T[i] = d[ti % 3]; }
        T[26] = 0;
        for (int s = 0; s < 9; ++s) out9[9 * w + (size_t)s] = pack3(T[3 * s], T[3 * s + 1], T[3 * s + 2]);
    }
    free(use);
    *n_out = nwords;
    return 1;
}

/* ------------------------------------------------------------------ */
/* Consistent decoder for FIXED mode (ours).                */
/* Inverts t3o_encode_profile(fixed=1): skip 52 header symbols,         */
/* un-beacon by body-word index, descramble by pre-beacon index,        */
/* band-major regions -> RS decode -> re-multiplex, de-interleave,      */
/* regroup 26 trits/word.  Tail loss (B8) remains.                      */
/* ------------------------------------------------------------------ */
static int geometry_from_nout(const t3o_cfg* c, size_t n_out_words, geom_t* g, size_t* n_words_min)
{
    /* N_out is monotone in N_w: binary search the smallest N_w that yields n_out_words. */
    size_t lo = 0, hi = n_out_words + 16; /* N_w < N_out*... body is larger than input */
    while (t3o_profile_words_bound(c, hi) < n_out_words) hi *= 2;
    while (lo < hi) { size_t mid = lo + (hi - lo) / 2; if (t3o_profile_words_bound(c, mid) < n_out_words) lo = mid + 1; else hi = mid; }
    geometry(c, lo, g);
    if (g->n_out != n_out_words) return 0;
    *n_words_min = lo;
    return 1;
}

int t3o_decode_profile_fixed(const t3o_cfg* c, size_t n_raw_words, const uint8_t* in9, size_t n, uint8_t* out9,
                             size_t cap, size_t* n_out, size_t* n_corrected)
{
    gf_init();
    *n_out = 0; if (n_corrected) *n_corrected = 0;
    if (c->profile == T3O_PROFILE_RAW) { if (cap < n) return 0; memcpy(out9, in9, 9 * n); *n_out = n; return 1; }
    geom_t g;
    if (n_raw_words) { geometry(c, n_raw_words, &g); if (g.n_out != n) return 0; }
    else {
        if (!geometry_from_nout(c, n, &g, &n_raw_words)) return 0;
        if (use_2d(c) && g.l_body) return 0; /* the last partial tile's permutation depends on N_w */
    }
    /* un-beacon */
    uint8_t* body = (uint8_t*)malloc(g.l_body ? g.l_body : 1);
    if (use_beacon(c)) {
        size_t k = 0, wd = 0, q = 52;
        while (k < g.l_body) {
            int ins = (wd % c->beacon_period) == 0;
            for (int slot = 0; slot < 9; ++slot) {
                if (ins && slot == c->beacon_slot) { ++q; continue; }
                if (k < g.l_body) body[k++] = in9[q];
                ++q;
            }
            ++wd;
        }
    } else memcpy(body, in9 + 52, g.l_body);
    /* descramble by pre-beacon index */
    uint32_t st = c->seed_s0 % 3;
    for (size_t p = 0; p < g.l_body; ++p) body[p] = t3o_descramble_symbol(body[p], c->seed_a, c->seed_b, &st);
    /* RS decode per band, re-multiplex sy'[9m+b] */
    size_t ns = g.n_s, ncorr = 0;
    uint8_t* sy = (uint8_t*)calloc(ns ? ns : 1, 1);
    size_t pfx = ns; /* contiguous known prefix of sy' */
    int ok = 1;
    for (int b = 0; b < 9; ++b) {
        int k = g.k_b[b];
        for (size_t cw = 0; cw < g.ncw_b[b]; ++cw) {
            uint8_t nbuf[26], kbuf[26], orig[26];
            memcpy(nbuf, body + 26 * (g.cw_base[b] + cw), 26); memcpy(orig, nbuf, 26);
            if (!rs_decode_ex(k, 1, 1, nbuf, kbuf)) { ok = 0; break; }
            for (int i = 0; i < 26; ++i) if (nbuf[i] != orig[i]) ++ncorr;
            for (int i = 0; i < k; ++i) sy[9 * ((size_t)k * cw + (size_t)i) + (size_t)b] = kbuf[i];
        }
        if (!ok) break;
        size_t first_unknown = 9 * ((size_t)k * g.ncw_b[b]) + (size_t)b;
        if (first_unknown < pfx) pfx = first_unknown;
    }
    free(body);
    if (!ok) { free(sy); return 0; }
    size_t known = pfx;
    if (use_2d(c)) {
        /* sy'[i]=sy[src(i)] is an involution per row; unknown tail i>=pfx maps into its own rows */
        t3o_deinterleave2d(sy, ns, c->tile_w, c->tile_h);
        if (pfx < ns) {
            size_t A = (size_t)c->tile_w * c->tile_h, base = (pfx / A) * A, off = pfx - base;
            size_t row = off / c->tile_w, rs = base + row * c->tile_w;
            known = (row % 2 == 1) ? rs : pfx; /* a reversed row loses its low end first */
        }
    }
    size_t nwords = (3 * known) / 26;
    if (nwords > n_raw_words) nwords = n_raw_words;
    if (cap < nwords) { free(sy); return 0; }
    for (size_t w = 0; w < nwords; ++w) {
        uint8_t T[27];
        for (int i = 0; i < 26; ++i) { size_t ti = 26 * w + (size_t)i; uint8_t d[3]; unpack3(sy[ti / 3], d); T[i] = d[ti % 3]; }
        T[26] = 0;
        for (int s = 0; s < 9; ++s) out9[9 * w + (size_t)s] = pack3(T[3 * s], T[3 * s + 1], T[3 * s + 2]);
    }
    free(sy);
    *n_out = nwords;
    if (n_corrected) *n_corrected = ncorr;
    return 1;
}

/* ------------------------------------------------------------------ */
/* RGB8 <-> YCbCr8 <-> quant bridge, IMG:47-84,156-192                  */
/* ------------------------------------------------------------------ */
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
void t3o_rgb_to_quant(const uint8_t* rgb, size_t n, t3o_pixel* out)
{
    for (size_t i = 0; i < n; ++i) {
        float r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];            /* IMG:49 */
        float y = 0.299f * r + 0.587f * g + 0.114f * b;                          /* IMG:50 */
        float cb = -0.168736f * r - 0.331264f * g + 0.5f * b + 128.0f;           /* IMG:51 */
        float cr = 0.5f * r - 0.418688f * g - 0.081312f * b + 128.0f;            /* IMG:52 */
        int Y = clampi((int)roundf(y), 0, 255), Cb = clampi((int)roundf(cb), 0, 255), Cr = clampi((int)roundf(cr), 0, 255);
        out[i].Yq = (uint16_t)clampi((int)round(Y * (242.0 / 255.0)), 0, 242);   /* IMG:72 */
        out[i].Cbq = (int16_t)clampi((int)round((Cb - 128) * (40.0 / 128.0)), -40, 40); /* IMG:73,75 */
        out[i].Crq = (int16_t)clampi((int)round((Cr - 128) * (40.0 / 128.0)), -40, 40); /* IMG:74,76 */
    }
}
void t3o_quant_to_rgb(const t3o_pixel* q, size_t n, uint8_t* rgb)
{
    for (size_t i = 0; i < n; ++i) {
        int Y = clampi((int)round(q[i].Yq * (255.0 / 242.0)), 0, 255);           /* IMG:81 */
        int Cb = clampi((int)round(128 + q[i].Cbq * (128.0 / 40.0)), 0, 255);    /* IMG:82 */
        int Cr = clampi((int)round(128 + q[i].Crq * (128.0 / 40.0)), 0, 255);    /* IMG:83 */
        float y = (float)Y, cb = (float)Cb - 128.0f, cr = (float)Cr - 128.0f;    /* IMG:59 */
        float r = y + 1.402f * cr;                                               /* IMG:60 */
        float g = y - 0.344136f * cb - 0.714136f * cr;                           /* IMG:61 */
        float b = y + 1.772f * cb;                                               /* IMG:62 */
        rgb[3 * i] = (uint8_t)clampi((int)roundf(r), 0, 255);
        rgb[3 * i + 1] = (uint8_t)clampi((int)roundf(g), 0, 255);
        rgb[3 * i + 2] = (uint8_t)clampi((int)roundf(b), 0, 255);
    }
}

/* ------------------------------------------------------------------ */
/* Fused conveniences (chain of the stages above, old/src/main.cpp:14-26) */
/* ------------------------------------------------------------------ */
size_t t3o_encode_rgb(const t3o_cfg* c, int fixed, const uint8_t* rgb, size_t n_px, uint8_t* out9, size_t cap)
{
    t3o_pixel* q = (t3o_pixel*)malloc((n_px ? n_px : 1) * sizeof *q);
    size_t nw = (n_px + 1) / 2;
    uint8_t* raw = (uint8_t*)malloc(nw ? 9 * nw : 1);
    t3o_rgb_to_quant(rgb, n_px, q);
    t3o_pack_pixels(q, n_px, raw);
    size_t r = t3o_encode_profile(c, fixed, raw, nw, out9, cap);
    free(raw); free(q);
    return r;
}
int t3o_decode_rgb_fixed(const t3o_cfg* c, size_t n_px, const uint8_t* in9, size_t n_words, uint8_t* rgb,
                         size_t* n_px_out, size_t* n_corrected)
{
    size_t nw = (n_px + 1) / 2, got = 0;
    uint8_t* raw = (uint8_t*)malloc(nw ? 9 * nw : 1);
    int ok = t3o_decode_profile_fixed(c, nw, in9, n_words, raw, nw, &got, n_corrected);
    *n_px_out = 0;
    if (ok) {
        t3o_pixel* q = (t3o_pixel*)malloc((got ? 2 * got : 1) * sizeof *q);
        t3o_unpack_pixels(raw, got, q);
        size_t np = 2 * got < n_px ? 2 * got : n_px;
        t3o_quant_to_rgb(q, np, rgb);
        *n_px_out = np;
        free(q);
    }
    free(raw);
    return ok;
}

/* ------------------------------------------------------------------ */
/* sub-word streams, OLD:816-859                        */
/* ------------------------------------------------------------------ */
void t3o_subword_stream(const uint8_t* words9, size_t n_words, int N, uint8_t* trits) /* extract_subword_stream_from_words, OLD:835-845 */
{
    for (size_t w = 0; w < n_words; ++w) {
        uint8_t T[27];
        for (int s = 0; s < 9; ++s) unpack3(words9[9 * w + s], T + 3 * s); /* extract_subword_trits_from_word, OLD:817-827 */
        for (int i = 0; i < N; ++i) trits[(size_t)N * w + i] = T[i];
    }
}
size_t t3o_words_from_subword_stream(const uint8_t* trits, size_t n_trits, int N, uint8_t fill, uint8_t* words9) /* OLD:846-859 */
{
    size_t idx = 0, nw = 0;
    while (idx < n_trits) {
        uint8_t buf[27] = {0}, T[27];
        int take = (int)((n_trits - idx) < (size_t)N ? (n_trits - idx) : (size_t)N);
        for (int i = 0; i < take; ++i) buf[i] = trits[idx + i];
        for (int i = 0; i < N; ++i) T[i] = buf[i];            /* inject_subword_trits_into_word, OLD:828-834 */
        for (int i = N; i < 27; ++i) T[i] = fill;
        for (int s = 0; s < 9; ++s) words9[9 * nw + s] = pack3(T[3 * s], T[3 * s + 1], T[3 * s + 2]);
        ++nw;
        idx += (size_t)take;
    }
    return nw;
}
/* base-243, include/ternary_packing.hpp:18-50: uint32 LE trit count, then 5 trits per byte (LSD first, zero padded) */
size_t t3o_base243_pack(const uint8_t* trits, size_t n_trits, uint8_t* out)
{
    uint32_t total = (uint32_t)n_trits;
    memcpy(out, &total, 4);
    size_t o = 4, i = 0;
    while (i < n_trits) {
        uint8_t buf[5] = {0, 0, 0, 0, 0};
        for (int t = 0; t < 5 && i < n_trits; ++t) buf[t] = trits[i++];
        out[o++] = (uint8_t)(buf[0] + 3 * buf[1] + 9 * buf[2] + 27 * buf[3] + 81 * buf[4]);
    }
    return o;
}
int t3o_base243_unpack(const uint8_t* in, size_t n_bytes, uint8_t* trits, size_t cap, size_t* n_trits)
{
    *n_trits = 0;
    if (n_bytes < 4) return 0;
    uint32_t total = 0;
    memcpy(&total, in, 4);
    size_t idx = 4, n = 0;
    while (idx < n_bytes && n < total) {
        int v = in[idx++];
        for (int k = 0; k < 5; ++k) { uint8_t t = (uint8_t)(v % 3); v /= 3; if (n < total) { if (n < cap) trits[n] = t; ++n; } }
    }
    *n_trits = n;
    return n == total;
}
/* ------------------------------------------------------------------ */
/* NEW-generation RAW path, src/ternary_image_codec_v6_min.cpp:62-99 */
/* ------------------------------------------------------------------ */
void t3o_v6new_pack_pixels(const t3o_pixel* px, size_t n_px, uint32_t* words) /* pack13_from_quant, :62-78 */
{
    for (size_t i = 0; i < n_px; ++i) {
        uint32_t Y = (uint32_t)clampi(px[i].Yq, 0, 242), Cb = (uint32_t)clampi(px[i].Cbq + 40, 0, 80), Cr = (uint32_t)clampi(px[i].Crq + 40, 0, 80);
        words[i] = Y + 243u * (Cb + 81u * Cr);
    }
}
void t3o_v6new_unpack_pixels(const uint32_t* words, size_t n_words, t3o_pixel* px) /* unpack13_to_quant, :81-95 */
{
    for (size_t i = 0; i < n_words; ++i) {
        uint32_t code = words[i], block = code / 243u, Y = code % 243u, Cr = block / 81u, Cb = block % 81u;
        px[i].Yq = (uint16_t)(Y < 242u ? Y : 242u);
        px[i].Cbq = (int16_t)clampi((int32_t)Cb - 40, -40, 40);
        px[i].Crq = (int16_t)clampi((int32_t)Cr - 40, -40, 40);
    }
}


/* ------------------------------------------------------------------------------------------------------------------
 * the .t3v container's records (old/include/t3v_io.hpp).  CRC-32 (reflected 0xEDB88320, init and final
 * xor 0xFFFFFFFF), :14-40; a frame record is n (uint32 LE) | 9n symbol bytes, each % 27 | crc, with
 * crc = crc32(payload) ^ (crc32(&n, 4) * 16777619), :128-142; the header is the 54-byte packed T3VHeaderBin whose last
 * field is the CRC of the 50 bytes before it, :42-60,97-119. */
uint32_t t3o_crc32(const uint8_t* data, size_t n)
{
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; ++i) {
        c ^= data[i];
        for (int j = 0; j < 8; ++j) c = (c & 1u) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
    }
    return c ^ 0xFFFFFFFFu;
}
static void put32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
static uint32_t get32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
size_t t3o_t3v_frame_record(const uint8_t* words9, uint32_t n_words, uint8_t* out)
{
    const size_t nb = 9 * (size_t)n_words;
    put32(out, n_words);
    for (size_t i = 0; i < nb; ++i) out[4 + i] = (uint8_t)(words9[i] % 27);
    put32(out + 4 + nb, t3o_crc32(out + 4, nb) ^ (t3o_crc32(out, 4) * 16777619u));
    return 8 + nb;
}
int t3o_t3v_read_frame(const uint8_t* rec, size_t n_bytes, uint8_t* words9, uint32_t* n_words)
{
    *n_words = 0;
    if (n_bytes < 4) return 0;
    const uint32_t n = get32(rec);
    const size_t nb = 9 * (size_t)n;
    if (n_bytes < 8 + nb) return 0;                                  /* fread fails */
    if ((t3o_crc32(rec + 4, nb) ^ (t3o_crc32(rec, 4) * 16777619u)) != get32(rec + 4 + nb)) return 0;
    memcpy(words9, rec + 4, nb);                                     /* symbols are stored as read, not reduced again */
    *n_words = n;
    return 1;
}
void t3o_t3v_header(uint8_t out[54], int profile, int subword_code, int centered, int coset, uint32_t w, uint32_t h, const uint32_t aw[4],
                    uint32_t fps_num, uint32_t fps_den, uint32_t frame_count, int file_type)
{
    memset(out, 0, 54);
    memcpy(out, "T3V1", 4);
    out[4] = 1; out[5] = (uint8_t)file_type; out[6] = (uint8_t)profile; out[7] = (uint8_t)subword_code; out[8] = centered ? 1 : 0; out[9] = (uint8_t)coset;
    put32(out + 10, w); put32(out + 14, h);
    for (int i = 0; i < 4; ++i) put32(out + 18 + 4 * i, aw[i]);
    put32(out + 34, fps_num); put32(out + 38, fps_den); put32(out + 42, frame_count); put32(out + 46, 0);
    put32(out + 50, t3o_crc32(out, 50));
}


/* ------------------------------------------------------------------------------------------------------------------
 * image-bridge geometry of the NEW generation (include/io_image.hpp). */
void t3o_resize_rgb_nn(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) /* :102-124 */
{
    memset(dst, 0, (size_t)dw * dh * 3);
    if (sw <= 0 || sh <= 0) return;
    for (int y = 0; y < dh; ++y) {
        int sy = (int)((y + 0.5) * (double)sh / dh);
        sy = sy < 0 ? 0 : (sy > sh - 1 ? sh - 1 : sy);
        for (int x = 0; x < dw; ++x) {
            int sx = (int)((x + 0.5) * (double)sw / dw);
            sx = sx < 0 ? 0 : (sx > sw - 1 ? sw - 1 : sx);
            memcpy(dst + ((size_t)y * dw + x) * 3, src + ((size_t)sy * sw + sx) * 3, 3);
        }
    }
}
void t3o_blit_center_rgb(const uint8_t* src, int sw, int sh, uint8_t* dst, int cw, int ch) /* :125-140 */
{
    memset(dst, 0, (size_t)cw * ch * 3);
    const int x0 = (cw - sw) / 2 > 0 ? (cw - sw) / 2 : 0, y0 = (ch - sh) / 2 > 0 ? (ch - sh) / 2 : 0;
    for (int y = 0; y < sh; ++y) {
        if (y + y0 < 0 || y + y0 >= ch) continue;
        memcpy(dst + (size_t)(y + y0) * cw * 3 + (size_t)x0 * 3, src + (size_t)y * sw * 3, (size_t)sw * 3);
    }
}
void t3o_extract_center_q(const t3o_pixel* full, int fw, int fh, int sw, int sh, t3o_pixel* sub) /* :215-235 */
{
    const int x0 = (fw - sw) / 2 > 0 ? (fw - sw) / 2 : 0, y0 = (fh - sh) / 2 > 0 ? (fh - sh) / 2 : 0;
    for (int y = 0; y < sh; ++y) {
        const int fy = y + y0;
        if (fy < 0 || fy >= fh) { memset(sub + (size_t)y * sw, 0, (size_t)sw * sizeof(t3o_pixel)); continue; } /* resize() value-initialises */
        memcpy(sub + (size_t)y * sw, full + (size_t)fy * fw + x0, (size_t)sw * sizeof(t3o_pixel));
    }
}
