/*
 * bncorpus.c -- synthetic PCM + from-scratch FLAC encoder + frame-level tiling (see bncorpus.h).
 * Test/bench infrastructure: produces the inputs; never linked into libbnflac.so.
 */
#include "bncorpus.h"
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ CRC + MD5 */
static uint8_t c8[256];
static uint16_t c16[8][256];
static pthread_once_t tab_once = PTHREAD_ONCE_INIT;
static void mk_tabs(void) {
    for (int i = 0; i < 256; i++) {
        uint8_t c = (uint8_t)i;
        for (int k = 0; k < 8; k++) c = (c & 0x80) ? (uint8_t)((c << 1) ^ 7) : (uint8_t)(c << 1);
        c8[i] = c;
        uint16_t d = (uint16_t)(i << 8);
        for (int k = 0; k < 8; k++) d = (d & 0x8000) ? (uint16_t)((d << 1) ^ 0x8005) : (uint16_t)(d << 1);
        c16[0][i] = d;
    }
    for (int t = 1; t < 8; t++)
        for (int i = 0; i < 256; i++) c16[t][i] = (uint16_t)((c16[t-1][i] << 8) ^ c16[0][c16[t-1][i] >> 8]);
}
static uint8_t crc8(const uint8_t* p, size_t n) { uint8_t c = 0; while (n--) c = c8[c ^ *p++]; return c; }
static uint16_t crc16(const uint8_t* p, size_t n) {
    uint16_t c = 0;
    while (n >= 8) {
        c = (uint16_t)(c16[7][p[0] ^ (c >> 8)] ^ c16[6][p[1] ^ (c & 0xff)] ^ c16[5][p[2]] ^ c16[4][p[3]] ^
                       c16[3][p[4]] ^ c16[2][p[5]] ^ c16[1][p[6]] ^ c16[0][p[7]]);
        p += 8; n -= 8;
    }
    while (n--) c = (uint16_t)((c << 8) ^ c16[0][(c >> 8) ^ *p++]);
    return c;
}

typedef struct { uint32_t s[4]; uint64_t n; uint8_t buf[64]; size_t fill; } md5_t;
static const uint32_t md5_k[64] = {
    0xd76aa478,0xe8c7b756,0x242070db,0xc1bdceee,0xf57c0faf,0x4787c62a,0xa8304613,0xfd469501,0x698098d8,0x8b44f7af,0xffff5bb1,0x895cd7be,0x6b901122,0xfd987193,0xa679438e,0x49b40821,
    0xf61e2562,0xc040b340,0x265e5a51,0xe9b6c7aa,0xd62f105d,0x02441453,0xd8a1e681,0xe7d3fbc8,0x21e1cde6,0xc33707d6,0xf4d50d87,0x455a14ed,0xa9e3e905,0xfcefa3f8,0x676f02d9,0x8d2a4c8a,
    0xfffa3942,0x8771f681,0x6d9d6122,0xfde5380c,0xa4beea44,0x4bdecfa9,0xf6bb4b60,0xbebfbc70,0x289b7ec6,0xeaa127fa,0xd4ef3085,0x04881d05,0xd9d4d039,0xe6db99e5,0x1fa27cf8,0xc4ac5665,
    0xf4292244,0x432aff97,0xab9423a7,0xfc93a039,0x655b59c3,0x8f0ccc92,0xffeff47d,0x85845dd1,0x6fa87e4f,0xfe2ce6e0,0xa3014314,0x4e0811a1,0xf7537e82,0xbd3af235,0x2ad7d2bb,0xeb86d391};
static const uint8_t md5_r[64] = {7,12,17,22,7,12,17,22,7,12,17,22,7,12,17,22,5,9,14,20,5,9,14,20,5,9,14,20,5,9,14,20,
                                  4,11,16,23,4,11,16,23,4,11,16,23,4,11,16,23,6,10,15,21,6,10,15,21,6,10,15,21,6,10,15,21};
static void md5_block(md5_t* m, const uint8_t* p) {
    uint32_t w[16], a = m->s[0], b = m->s[1], c = m->s[2], d = m->s[3];
    memcpy(w, p, 64); /* little-endian host */
    for (int i = 0; i < 64; i++) {
        uint32_t f; int g;
        if (i < 16) { f = (b & c) | (~b & d); g = i; }
        else if (i < 32) { f = (d & b) | (~d & c); g = (5*i + 1) & 15; }
        else if (i < 48) { f = b ^ c ^ d; g = (3*i + 5) & 15; }
        else { f = c ^ (b | ~d); g = (7*i) & 15; }
        uint32_t t = a + f + md5_k[i] + w[g];
        a = d; d = c; c = b; b = b + ((t << md5_r[i]) | (t >> (32 - md5_r[i])));
    }
    m->s[0] += a; m->s[1] += b; m->s[2] += c; m->s[3] += d;
}
static void md5_init(md5_t* m) { m->s[0] = 0x67452301; m->s[1] = 0xefcdab89; m->s[2] = 0x98badcfe; m->s[3] = 0x10325476; m->n = 0; m->fill = 0; }
static void md5_update(md5_t* m, const uint8_t* p, size_t n) {
    m->n += n;
    if (m->fill) {
        size_t t = 64 - m->fill; if (t > n) t = n;
        memcpy(m->buf + m->fill, p, t); m->fill += t; p += t; n -= t;
        if (m->fill == 64) { md5_block(m, m->buf); m->fill = 0; }
    }
    for (; n >= 64; p += 64, n -= 64) md5_block(m, p);
    if (n) { memcpy(m->buf, p, n); m->fill = n; }
}
static void md5_final(md5_t* m, uint8_t out[16]) {
    uint64_t bits = m->n * 8;
    uint8_t pad[72] = {0x80};
    size_t padlen = (m->fill < 56) ? 56 - m->fill : 120 - m->fill;
    md5_update(m, pad, padlen);
    uint8_t l[8]; for (int k = 0; k < 8; k++) l[k] = (uint8_t)(bits >> (8*k));
    md5_update(m, l, 8);
    for (int k = 0; k < 4; k++) for (int j = 0; j < 4; j++) out[4*k + j] = (uint8_t)(m->s[k] >> (8*j));
}

/* ------------------------------------------------------------------ PCM synthesis */
void bnc_synth(int32_t* pcm, uint64_t n, uint32_t ch, uint32_t bps, uint32_t sr, uint32_t noise_bits, uint32_t kind,
               uint32_t special_period, uint32_t seed) {
    enum { TB = 4096 };
    static int32_t* tab = NULL;
    if (!tab) {
        /* round(32767 sin(2 pi i / 4096)) by an integer rotation in Q62 (no libm: the corpus must not depend on the host's
         * sin()); the constants are cos / sin of 2 pi / 4096 scaled by 2^62, the error after 4096 steps is below 2^-48 */
        const __int128 cd = 4611680592556051597LL, sd = 7074234977634094LL, half = (__int128)1 << 61;
        __int128 c = (__int128)1 << 62, sn = 0;
        tab = (int32_t*)malloc(sizeof(int32_t) * TB);
        for (int i = 0; i < TB; i++) {
            tab[i] = (int32_t)((sn * 32767 + half) >> 62);
            const __int128 c2 = (c * cd - sn * sd + half) >> 62, s2 = (sn * cd + c * sd + half) >> 62;
            c = c2; sn = s2;
        }
    }
    uint32_t s = seed * 2654435761u + 12345u;
    /* 4 partials per channel, phase increments in 1/2^32 cycles per sample */
    uint32_t inc[8][4], ph[8][4]; int32_t amp_shift[8][4];
    static const double base_hz[4] = {220.0, 554.37, 1318.5, 3520.0};
    for (uint32_t c = 0; c < ch; c++)
        for (int k = 0; k < 4; k++) {
            double hz = base_hz[k] * (1.0 + 0.003 * (double)(c / 2)); /* channel pairs share partials -> inter-channel correlation */
            inc[c][k] = (uint32_t)(hz / (double)sr * 4294967296.0);
            s = s * 1664525u + 1013904223u; ph[c][k] = (c & 1) ? ph[c-1][k] + (s >> 6) : s;
            amp_shift[c][k] = k; /* 1, 1/2, 1/4, 1/8 */
        }
    int scale_shift = (int)bps - 3 - 15; /* table peak 2^15 -> 2^(bps-3) */
    int32_t lim = (int32_t)((1u << (bps - 1)) - 1);
    uint32_t nb = noise_bits > bps ? bps : noise_bits;
    for (uint64_t i = 0; i < n; i++) {
        uint32_t seg = special_period ? (uint32_t)((i / special_period) % 8) : 0;
        /* slow amplitude envelope so Rice parameters drift between partitions */
        int32_t env = 256 + (tab[(uint32_t)((i * 3) >> 8) & (TB - 1)] >> 8); /* 128..384 */
        for (uint32_t c = 0; c < ch; c++) {
            int64_t v = 0;
            for (int k = 0; k < 4; k++) { ph[c][k] += inc[c][k]; v += tab[ph[c][k] >> 20] >> amp_shift[c][k]; }
            v = (v * env) >> 8;
            v = scale_shift >= 0 ? v << scale_shift : v >> (-scale_shift);
            s = s * 1664525u + 1013904223u;
            int32_t noise = nb ? (int32_t)(s >> (32 - nb)) - (int32_t)(1u << (nb - 1)) : 0;
            /* second LCG draw shapes the noise to a rough triangular pdf */
            s = s * 1664525u + 1013904223u;
            if (nb) noise = (noise + ((int32_t)(s >> (32 - nb)) - (int32_t)(1u << (nb - 1)))) / 2;
            v += noise;
            if (kind == 2) v = 0;                                                      /* whole stream digital silence */
            if (kind == 1) {
                switch (seg) {
                case 1: v = 0; break;                                                  /* digital silence -> CONSTANT */
                case 3: s = s * 1664525u + 1013904223u; v = (int32_t)s >> (32 - bps); break; /* full-scale noise -> VERBATIM */
                case 5: v &= ~(int64_t)0xF; break;                                     /* 4 wasted bits */
                case 6: if (c & 1) v = 1234 % (lim + 1); break;                        /* DC on odd channels */
                default: break;
                }
            }
            if (v > lim) v = lim; if (v < -lim - 1) v = -lim - 1;
            pcm[i * ch + c] = (int32_t)v;
        }
    }
}

void bnc_pack_pcm(const int32_t* pcm, uint64_t n, uint32_t bps, uint8_t* out) {
    unsigned B = (bps + 7) / 8;
    for (uint64_t i = 0; i < n; i++) { uint32_t v = (uint32_t)pcm[i]; for (unsigned k = 0; k < B; k++) *out++ = (uint8_t)(v >> (8*k)); }
}

/* ------------------------------------------------------------------ bit writer */
typedef struct { uint8_t* p; size_t cap; uint64_t acc; unsigned nacc; size_t len; } bw_t;
static void bw_init(bw_t* b) { b->cap = 1 << 14; b->p = (uint8_t*)malloc(b->cap); b->acc = 0; b->nacc = 0; b->len = 0; }
static void bw_reset(bw_t* b) { b->acc = 0; b->nacc = 0; b->len = 0; }
static void bw_free(bw_t* b) { free(b->p); b->p = NULL; }
static inline void bw_room(bw_t* b, size_t extra) { if (b->len + extra + 16 > b->cap) { while (b->len + extra + 16 > b->cap) b->cap *= 2; b->p = (uint8_t*)realloc(b->p, b->cap); } }
static inline void bw_put(bw_t* b, uint32_t v, unsigned n) { /* n <= 32 */
    if (!n) return;
    if (n < 32) v &= (1u << n) - 1;
    b->acc = (b->acc << n) | v; b->nacc += n;
    if (b->nacc >= 32) {
        bw_room(b, 8);
        while (b->nacc >= 8) { b->p[b->len++] = (uint8_t)(b->acc >> (b->nacc - 8)); b->nacc -= 8; }
    }
}
static inline void bw_unary(bw_t* b, uint32_t q) { while (q >= 32) { bw_put(b, 0, 32); q -= 32; } bw_put(b, 1, q + 1); }
static inline uint64_t bw_bits(const bw_t* b) { return (uint64_t)b->len * 8 + b->nacc; }
static void bw_flush_partial(bw_t* b) { bw_room(b, 8); while (b->nacc >= 8) { b->p[b->len++] = (uint8_t)(b->acc >> (b->nacc - 8)); b->nacc -= 8; } }
static void bw_align(bw_t* b) { if (b->nacc & 7) bw_put(b, 0, 8 - (b->nacc & 7)); bw_flush_partial(b); }
static void bw_append(bw_t* dst, bw_t* src) {
    bw_flush_partial(src);
    bw_room(dst, src->len + 8);
    for (size_t i = 0; i < src->len; i++) bw_put(dst, src->p[i], 8);
    if (src->nacc) bw_put(dst, (uint32_t)(src->acc & ((1u << src->nacc) - 1)), src->nacc);
}

/* ------------------------------------------------------------------ residual coding */
typedef struct { uint8_t k[1 << 8 | 1]; uint8_t esc[1 << 8 | 1]; uint8_t rawbits[1 << 8 | 1]; uint32_t po, method; uint64_t bits; } rice_plan;
#define MAX_PO 8

static inline uint32_t zz(int32_t r) { return ((uint32_t)r << 1) ^ (uint32_t)(r >> 31); }
static unsigned sbits(int32_t v) { /* bits for signed raw */ uint32_t a = v < 0 ? ~(uint32_t)v : (uint32_t)v; unsigned n = 1; while (a) { n++; a >>= 1; } return n; }

static void plan_rice(const int32_t* r, uint32_t bs, uint32_t order, const bnc_params* p, uint32_t part_counter, rice_plan* best) {
    best->bits = ~0ull;
    uint32_t maxpo = p->max_part_order > MAX_PO ? MAX_PO : p->max_part_order;
    uint32_t minpo = p->min_part_order > maxpo ? maxpo : p->min_part_order;
    for (int po = (int)maxpo; po >= (int)minpo; po--) {
        if (po > 0) {
            if (bs & ((1u << po) - 1)) continue;
            uint32_t ps = bs >> po;
            if (p->allow_zero_part ? ps < order : ps <= order) continue;
        } else if (bs < order) continue;
        rice_plan cur; cur.po = (uint32_t)po; cur.method = 0; cur.bits = 0;
        uint32_t idx = 0, nparts = 1u << po;
        for (uint32_t q = 0; q < nparts; q++) {
            uint32_t n = po == 0 ? bs - order : (q == 0 ? (bs >> po) - order : (bs >> po));
            uint64_t sum = 0; unsigned rb = 0;
            for (uint32_t i = 0; i < n; i++) { sum += zz(r[idx + i]); }
            int force_esc = p->escape_every && ((part_counter + q) % p->escape_every == p->escape_every - 1);
            uint32_t k = 0;
            if (n) { uint64_t mean = sum / n; while ((mean >> k) > 1 && k < 30) k++; if (mean > 0 && k == 0 && mean >= 2) k = 1; }
            uint64_t bestb = ~0ull; uint32_t bestk = k;
            for (uint32_t kk = k; kk <= k + 1 && kk <= 30; kk++) {
                uint64_t bits = (uint64_t)n * (kk + 1);
                for (uint32_t i = 0; i < n; i++) bits += zz(r[idx + i]) >> kk;
                if (bits < bestb) { bestb = bits; bestk = kk; }
            }
            cur.esc[q] = 0;
            if (force_esc) {
                for (uint32_t i = 0; i < n; i++) { unsigned s = sbits(r[idx + i]); if (s > rb) rb = s; }
                int allzero = 1; for (uint32_t i = 0; i < n; i++) if (r[idx + i]) { allzero = 0; break; }
                if (allzero) rb = 0;
                if (rb <= 31) { cur.esc[q] = 1; cur.rawbits[q] = (uint8_t)rb; bestb = 5 + (uint64_t)n * rb; }
            }
            cur.k[q] = (uint8_t)bestk;
            if (!cur.esc[q] && bestk > 14) cur.method = 1;
            cur.bits += bestb;
            idx += n;
        }
        cur.bits += (uint64_t)nparts * (cur.method ? 5 : 4);
        if (cur.bits < best->bits) *best = cur;
    }
    if (best->bits != ~0ull) best->bits += 6;
}

static void write_residual(bw_t* b, const int32_t* r, uint32_t bs, uint32_t order, const rice_plan* pl) {
    bw_put(b, pl->method, 2); bw_put(b, pl->po, 4);
    unsigned plen = pl->method ? 5 : 4;
    uint32_t idx = 0, nparts = 1u << pl->po;
    for (uint32_t q = 0; q < nparts; q++) {
        uint32_t n = pl->po == 0 ? bs - order : (q == 0 ? (bs >> pl->po) - order : (bs >> pl->po));
        if (pl->esc[q]) {
            bw_put(b, pl->method ? 31 : 15, plen); bw_put(b, pl->rawbits[q], 5);
            for (uint32_t i = 0; i < n; i++) bw_put(b, (uint32_t)r[idx + i], pl->rawbits[q]);
        } else {
            uint32_t k = pl->k[q];
            bw_put(b, k, plen);
            for (uint32_t i = 0; i < n; i++) { uint32_t u = zz(r[idx + i]); bw_unary(b, u >> k); bw_put(b, u, k); }
        }
        idx += n;
    }
}

/* ------------------------------------------------------------------ prediction */
static int ilog2u(uint32_t v) { int l = 0; while (v >>= 1) l++; return l; }
/* log2 from IEEE +,-,*,/ in a fixed order (frexp is exact): the same double on every host, unlike libm's log() */
static double det_log2(double x) {
    int e; double m = frexp(x, &e);
    if (m < 0.70710678118654752) { m *= 2.0; e--; }
    const double t = (m - 1.0) / (m + 1.0), t2 = t * t;
    double s = 0.0;
    for (int k = 25; k >= 1; k -= 2) s = s * t2 + 1.0 / (double)k;      /* ln m = 2 atanh t */
    return (double)e + 2.0 * t * s * 1.4426950408889634;
}

/* returns 0 if usable */
static int lpc_analyse(const int32_t* x, uint32_t bs, uint32_t maxorder, double lpc[33][32], double err[33]) {
    double ac[33];
    double* w = (double*)malloc(sizeof(double) * bs);
    double half = (bs - 1) / 2.0;
    for (uint32_t i = 0; i < bs; i++) { double t = (i - half) / (half + 1.0); w[i] = x[i] * (1.0 - t * t); }
    for (uint32_t l = 0; l <= maxorder; l++) { double s = 0; for (uint32_t i = l; i < bs; i++) s += w[i] * w[i - l]; ac[l] = s; }
    free(w);
    if (ac[0] <= 0.0) return -1;
    double e = ac[0], a[32];
    err[0] = e;
    for (uint32_t i = 0; i < maxorder; i++) {
        double r = -ac[i + 1];
        for (uint32_t j = 0; j < i; j++) r -= a[j] * ac[i - j];
        r /= e;
        a[i] = r;
        for (uint32_t j = 0; j < i / 2; j++) { double t = a[j]; a[j] += r * a[i - 1 - j]; a[i - 1 - j] += r * t; }
        if (i & 1) a[i / 2] += a[i / 2] * r;
        e *= (1.0 - r * r);
        if (!(e > 0.0)) e = 1e-9;
        for (uint32_t j = 0; j <= i; j++) lpc[i + 1][j] = -a[j];
        err[i + 1] = e;
    }
    return 0;
}

static int quantize(const double* lp, uint32_t order, uint32_t prec, int32_t* q, int* shift_out) {
    double cmax = 0; for (uint32_t i = 0; i < order; i++) { double a = fabs(lp[i]); if (a > cmax) cmax = a; }
    if (cmax <= 0) return -1;
    int l2; (void)frexp(cmax, &l2); l2--;
    int shift = (int)prec - l2 - 1;
    if (shift > 15) shift = 15;
    if (shift < 0) return -1;
    int32_t qmax = (1 << (prec - 1)) - 1, qmin = -qmax - 1;
    double e = 0;
    for (uint32_t i = 0; i < order; i++) {
        e += lp[i] * (double)(1 << shift);
        long v = lround(e);
        if (v > qmax) v = qmax; if (v < qmin) v = qmin;
        e -= (double)v; q[i] = (int32_t)v;
    }
    *shift_out = shift;
    return 0;
}

typedef struct { int32_t* r_fixed; int32_t* r_lpc; int32_t* tmp; bw_t cand[4]; bw_t frame; uint32_t cap; uint32_t part_counter, sub_counter; } enc_work;

/* encode one subframe of effective width bps into b; returns bits written */
static uint64_t encode_subframe(bw_t* b, const int32_t* xin, uint32_t bs, uint32_t bps, const bnc_params* p, enc_work* wk) {
    uint64_t start = bw_bits(b);
    int32_t* x = wk->tmp;
    /* wasted bits */
    uint32_t orv = 0; int allsame = 1;
    for (uint32_t i = 0; i < bs; i++) { orv |= (uint32_t)xin[i]; if (xin[i] != xin[0]) allsame = 0; }
    uint32_t w = 0;
    if (orv) { while (!((orv >> w) & 1)) w++; }
    if (w >= bps) w = 0;
    for (uint32_t i = 0; i < bs; i++) x[i] = xin[i] >> w;
    uint32_t ebps = bps - w;
    uint32_t subno = wk->sub_counter++;
    int force_verbatim = p->verbatim_every && (subno % p->verbatim_every == p->verbatim_every - 1);
#define HDR(type) do { bw_put(b, 0, 1); bw_put(b, (type), 6); bw_put(b, w ? 1 : 0, 1); if (w) bw_unary(b, w - 1); } while (0)
    if (allsame && !force_verbatim) { HDR(0); bw_put(b, (uint32_t)x[0], ebps); return bw_bits(b) - start; }
    uint64_t verb_bits = (uint64_t)bs * ebps;
    /* FIXED: choose order by sum |residual| */
    uint32_t fo = 0; rice_plan fplan; fplan.bits = ~0ull;
    uint64_t fixed_bits = ~0ull;
    if (!force_verbatim) {
        uint64_t bestsum = ~0ull;
        for (uint32_t o = 0; o <= 4 && o < bs; o++) {
            uint64_t sum = 0; int ok = 1;
            for (uint32_t i = o; i < bs; i++) {
                int64_t pr = 0;
                switch (o) { case 1: pr = x[i-1]; break; case 2: pr = 2*(int64_t)x[i-1] - x[i-2]; break;
                    case 3: pr = 3*(int64_t)x[i-1] - 3*(int64_t)x[i-2] + x[i-3]; break;
                    case 4: pr = 4*(int64_t)x[i-1] - 6*(int64_t)x[i-2] + 4*(int64_t)x[i-3] - x[i-4]; break; default: break; }
                int64_t r = x[i] - pr;
                if (r > 0x3fffffff || r < -0x3fffffff) { ok = 0; break; }
                sum += (uint64_t)(r < 0 ? -r : r);
            }
            if (ok && sum < bestsum) { bestsum = sum; fo = o; }
        }
        if (bestsum != ~0ull) {
            for (uint32_t i = fo; i < bs; i++) {
                int64_t pr = 0;
                switch (fo) { case 1: pr = x[i-1]; break; case 2: pr = 2*(int64_t)x[i-1] - x[i-2]; break;
                    case 3: pr = 3*(int64_t)x[i-1] - 3*(int64_t)x[i-2] + x[i-3]; break;
                    case 4: pr = 4*(int64_t)x[i-1] - 6*(int64_t)x[i-2] + 4*(int64_t)x[i-3] - x[i-4]; break; default: break; }
                wk->r_fixed[i - fo] = (int32_t)(x[i] - pr);
            }
            plan_rice(wk->r_fixed, bs, fo, p, wk->part_counter, &fplan);
            if (fplan.bits != ~0ull) fixed_bits = fplan.bits + (uint64_t)fo * ebps;
        }
    }
    /* LPC */
    uint64_t lpc_bits = ~0ull; uint32_t lo = 0, prec = 0; int shift = 0; int32_t qc[32]; rice_plan lplan; lplan.bits = ~0ull;
    uint32_t maxo = p->max_lpc_order; if (maxo > 32) maxo = 32; if (maxo >= bs) maxo = bs - 1;
    if (maxo && !force_verbatim) {
        static __thread double lpc[33][32]; double err[33];
        if (lpc_analyse(x, bs, maxo, lpc, err) == 0) {
            prec = p->qlp_precision ? p->qlp_precision : (ebps > 16 ? (bs > 1152 ? 15 : 14) : (bs > 4608 ? 13 : 12));
            if (prec > 15) prec = 15;
            lo = maxo;
            if (p->search_order) {
                double bestest = 1e300;
                for (uint32_t o = 1; o <= maxo; o++) {
                    double e = err[o] * (0.5 * M_LN2 * M_LN2 / (double)bs);
                    double bpr = e > 0 ? 0.5 * det_log2(e) : 0; if (bpr < 0) bpr = 0;
                    double est = bpr * (bs - o) + (double)o * (ebps + prec);
                    if (est < bestest) { bestest = est; lo = o; }
                }
            }
            if (quantize(lpc[lo], lo, prec, qc, &shift) == 0) {
                int ok = 1;
                for (uint32_t i = lo; i < bs && ok; i++) {
                    int64_t s = 0; for (uint32_t j = 0; j < lo; j++) s += (int64_t)qc[j] * x[i - 1 - j];
                    int64_t r = (int64_t)x[i] - (s >> shift);
                    if (r > 0x3fffffff || r < -0x3fffffff) ok = 0;
                    wk->r_lpc[i - lo] = (int32_t)r;
                }
                if (ok) { plan_rice(wk->r_lpc, bs, lo, p, wk->part_counter, &lplan);
                          if (lplan.bits != ~0ull) lpc_bits = lplan.bits + (uint64_t)lo * (ebps + prec) + 9; }
            }
        }
    }
    if (force_verbatim || (verb_bits <= fixed_bits && verb_bits <= lpc_bits)) {
        HDR(1); for (uint32_t i = 0; i < bs; i++) bw_put(b, (uint32_t)x[i], ebps);
    } else if (lpc_bits < fixed_bits) {
        HDR(32 + lo - 1);
        for (uint32_t i = 0; i < lo; i++) bw_put(b, (uint32_t)x[i], ebps);
        bw_put(b, prec - 1, 4); bw_put(b, (uint32_t)shift, 5);
        for (uint32_t i = 0; i < lo; i++) bw_put(b, (uint32_t)qc[i], prec);
        write_residual(b, wk->r_lpc, bs, lo, &lplan);
        wk->part_counter += 1u << lplan.po;
    } else {
        HDR(8 + fo);
        for (uint32_t i = 0; i < fo; i++) bw_put(b, (uint32_t)x[i], ebps);
        write_residual(b, wk->r_fixed, bs, fo, &fplan);
        wk->part_counter += 1u << fplan.po;
    }
#undef HDR
    return bw_bits(b) - start;
}

static unsigned put_utf8(uint8_t* o, uint64_t v) {
    if (v < 0x80) { o[0] = (uint8_t)v; return 1; }
    unsigned n = v < 0x800 ? 2 : v < 0x10000 ? 3 : v < 0x200000 ? 4 : v < 0x4000000 ? 5 : v < 0x80000000ull ? 6 : 7;
    for (unsigned i = n - 1; i >= 1; i--) { o[i] = (uint8_t)(0x80 | (v & 0x3f)); v >>= 6; }
    o[0] = (uint8_t)((0xFF << (8 - n)) | (n == 7 ? 0 : v));
    return n;
}

static unsigned frame_header(uint8_t* h, const bnc_params* p, uint32_t bs, uint32_t assignment, int variable, uint64_t number) {
    static const uint32_t bst[16] = {0,192,576,1152,2304,4608,0,0,256,512,1024,2048,4096,8192,16384,32768};
    static const uint32_t srt[12] = {0,88200,176400,192000,8000,16000,22050,24000,32000,44100,48000,96000};
    unsigned bsc = 0, src = 0, ssc = 0;
    for (unsigned i = 1; i < 16; i++) if (bst[i] == bs) bsc = i;
    if (!bsc) bsc = (bs <= 256) ? 6 : 7;
    if (!p->streaminfo_in_frames) {
        for (unsigned i = 1; i < 12; i++) if (srt[i] == p->sample_rate) src = i;
        if (!src) { if (p->sample_rate % 1000 == 0 && p->sample_rate / 1000 < 256) src = 12; else if (p->sample_rate < 65536) src = 13; else if (p->sample_rate % 10 == 0 && p->sample_rate / 10 < 65536) src = 14; }
        switch (p->bps) { case 8: ssc = 1; break; case 12: ssc = 2; break; case 16: ssc = 4; break; case 20: ssc = 5; break; case 24: ssc = 6; break; default: ssc = 0; }
    }
    unsigned q = 0;
    h[q++] = 0xFF; h[q++] = (uint8_t)(0xF8 | (variable ? 1 : 0));
    h[q++] = (uint8_t)(bsc << 4 | src);
    h[q++] = (uint8_t)(assignment << 4 | ssc << 1);
    q += put_utf8(h + q, number);
    if (bsc == 6) h[q++] = (uint8_t)(bs - 1);
    else if (bsc == 7) { h[q++] = (uint8_t)((bs - 1) >> 8); h[q++] = (uint8_t)(bs - 1); }
    if (src == 12) h[q++] = (uint8_t)(p->sample_rate / 1000);
    else if (src == 13) { h[q++] = (uint8_t)(p->sample_rate >> 8); h[q++] = (uint8_t)p->sample_rate; }
    else if (src == 14) { h[q++] = (uint8_t)((p->sample_rate / 10) >> 8); h[q++] = (uint8_t)(p->sample_rate / 10); }
    h[q] = crc8(h, q); q++;
    return q;
}

static void work_init(enc_work* w, uint32_t cap) {
    w->cap = cap; w->r_fixed = (int32_t*)malloc(4ull * cap); w->r_lpc = (int32_t*)malloc(4ull * cap); w->tmp = (int32_t*)malloc(4ull * cap);
    for (int i = 0; i < 4; i++) bw_init(&w->cand[i]);
    bw_init(&w->frame); w->part_counter = 0; w->sub_counter = 0;
}
static void work_done(enc_work* w) { free(w->r_fixed); free(w->r_lpc); free(w->tmp); for (int i = 0; i < 4; i++) bw_free(&w->cand[i]); bw_free(&w->frame); }

/* encode one frame; result in wk->frame */
static void encode_frame(const int32_t* pcm, uint32_t bs, const bnc_params* p, int variable, uint64_t number, enc_work* wk, int32_t** chbuf) {
    uint32_t ch = p->channels;
    for (uint32_t c = 0; c < ch; c++) for (uint32_t i = 0; i < bs; i++) chbuf[c][i] = pcm[(uint64_t)i * ch + c];
    bw_t* f = &wk->frame; bw_reset(f);
    uint32_t assignment = ch - 1;
    bw_t body; bw_init(&body);
    if (ch == 2 && p->stereo_mode) {
        int32_t* L = chbuf[0]; int32_t* R = chbuf[1]; int32_t* M = chbuf[2]; int32_t* S = chbuf[3];
        for (uint32_t i = 0; i < bs; i++) { M[i] = (int32_t)(((int64_t)L[i] + R[i]) >> 1); S[i] = L[i] - R[i]; }
        uint64_t bits[4] = {~0ull, ~0ull, ~0ull, ~0ull};
        int need[4] = {0,0,0,0};
        switch (p->stereo_mode) { case 1: need[0]=need[1]=need[2]=need[3]=1; break; case 2: need[0]=need[3]=1; break; case 3: need[1]=need[3]=1; break; default: need[2]=need[3]=1; break; }
        uint32_t pc0 = wk->part_counter, sc0 = wk->sub_counter;
        const int32_t* src[4] = {L, R, M, S};
        for (int k = 0; k < 4; k++) if (need[k]) { bw_reset(&wk->cand[k]); wk->part_counter = pc0 + (uint32_t)k; wk->sub_counter = sc0 + (k == 3 || k == 1 ? 1u : 0u); bits[k] = encode_subframe(&wk->cand[k], src[k], bs, p->bps + (k == 3), p, wk); }
        wk->part_counter = pc0 + 7; wk->sub_counter = sc0 + 2;
        int mode = p->stereo_mode;
        if (mode == 1) {
            uint64_t lr = bits[0] + bits[1], ls = bits[0] + bits[3], sr = bits[3] + bits[1], ms = bits[2] + bits[3];
            mode = 0; uint64_t bb = lr;
            if (ls < bb) { bb = ls; mode = 2; } if (sr < bb) { bb = sr; mode = 3; } if (ms < bb) { bb = ms; mode = 4; }
        }
        switch (mode) {
        case 0: assignment = 1; bw_append(&body, &wk->cand[0]); bw_append(&body, &wk->cand[1]); break;
        case 2: assignment = 8; bw_append(&body, &wk->cand[0]); bw_append(&body, &wk->cand[3]); break;
        case 3: assignment = 9; bw_append(&body, &wk->cand[3]); bw_append(&body, &wk->cand[1]); break;
        default: assignment = 10; bw_append(&body, &wk->cand[2]); bw_append(&body, &wk->cand[3]); break;
        }
    } else {
        for (uint32_t c = 0; c < ch; c++) encode_subframe(&body, chbuf[c], bs, p->bps, p, wk);
    }
    uint8_t h[20]; unsigned hl = frame_header(h, p, bs, assignment, variable, number);
    for (unsigned i = 0; i < hl; i++) bw_put(f, h[i], 8);
    bw_append(f, &body); bw_free(&body);
    bw_align(f);
    uint16_t c = crc16(f->p, f->len);
    bw_put(f, c, 16); bw_flush_partial(f);
}

typedef struct {
    const int32_t* pcm; const bnc_params* p; uint64_t nsamples;
    uint32_t nframes; const uint64_t* fstart; const uint32_t* fbs; int variable;
    uint8_t** out; size_t* outlen; volatile uint32_t* next;
} job_t;
static void* enc_thread(void* arg) {
    job_t* j = (job_t*)arg;
    uint32_t maxbs = 0; for (uint32_t i = 0; i < j->nframes; i++) if (j->fbs[i] > maxbs) maxbs = j->fbs[i];
    enc_work wk; work_init(&wk, maxbs + 64);
    int32_t* chbuf[8]; for (int c = 0; c < 8; c++) chbuf[c] = (int32_t*)malloc(4ull * (maxbs + 64));
    for (;;) {
        uint32_t f = __sync_fetch_and_add(j->next, 1);
        if (f >= j->nframes) break;
        /* deterministic per-frame counters so output does not depend on thread scheduling */
        wk.part_counter = f * 131u; wk.sub_counter = f * j->p->channels;
        encode_frame(j->pcm + j->fstart[f] * j->p->channels, j->fbs[f], j->p, j->variable, j->variable ? j->fstart[f] : f, &wk, chbuf);
        j->out[f] = (uint8_t*)malloc(wk.frame.len); memcpy(j->out[f], wk.frame.p, wk.frame.len); j->outlen[f] = wk.frame.len;
    }
    for (int c = 0; c < 8; c++) free(chbuf[c]);
    work_done(&wk);
    return NULL;
}

static size_t write_metadata(uint8_t* o, const bnc_params* p, uint32_t minbs, uint32_t maxbs, uint32_t minfs, uint32_t maxfs, uint64_t total, const uint8_t md5[16]) {
    size_t q = 0;
    memcpy(o, "fLaC", 4); q = 4;
    int more = p->padding_bytes > 0;
    o[q++] = (uint8_t)(more ? 0x00 : 0x80); o[q++] = 0; o[q++] = 0; o[q++] = 34;
    o[q++] = (uint8_t)(minbs >> 8); o[q++] = (uint8_t)minbs; o[q++] = (uint8_t)(maxbs >> 8); o[q++] = (uint8_t)maxbs;
    o[q++] = (uint8_t)(minfs >> 16); o[q++] = (uint8_t)(minfs >> 8); o[q++] = (uint8_t)minfs;
    o[q++] = (uint8_t)(maxfs >> 16); o[q++] = (uint8_t)(maxfs >> 8); o[q++] = (uint8_t)maxfs;
    uint64_t x = ((uint64_t)p->sample_rate << 44) | ((uint64_t)(p->channels - 1) << 41) | ((uint64_t)(p->bps - 1) << 36) | (total & 0xFFFFFFFFFull);
    for (int i = 7; i >= 0; i--) o[q++] = (uint8_t)(x >> (8*i));
    memcpy(o + q, md5, 16); q += 16;
    if (more) {
        /* VORBIS_COMMENT (vendor only) then PADDING as last block */
        static const char vendor[] = "bncorpus synthetic";
        uint32_t vl = (uint32_t)strlen(vendor), bl = 4 + vl + 4;
        o[q++] = 0x04; o[q++] = 0; o[q++] = (uint8_t)(bl >> 8); o[q++] = (uint8_t)bl;
        o[q++] = (uint8_t)vl; o[q++] = 0; o[q++] = 0; o[q++] = 0; memcpy(o + q, vendor, vl); q += vl;
        o[q++] = 0; o[q++] = 0; o[q++] = 0; o[q++] = 0;
        uint32_t pl = p->padding_bytes;
        o[q++] = 0x81; o[q++] = (uint8_t)(pl >> 16); o[q++] = (uint8_t)(pl >> 8); o[q++] = (uint8_t)pl;
        memset(o + q, 0, pl); q += pl;
    }
    return q;
}

int bnc_encode(const int32_t* pcm, uint64_t nsamples, const bnc_params* p, bnc_stream* out) {
    pthread_once(&tab_once, mk_tabs);
    memset(out, 0, sizeof *out);
    if (!p->channels || p->channels > 8 || p->bps < 4 || p->bps > 24 || !p->blocksize) return -1;
    /* frame plan */
    uint32_t cap = (uint32_t)(nsamples / (p->nvar ? 16 : p->blocksize) + 2);
    uint64_t* fstart = (uint64_t*)malloc(sizeof(uint64_t) * (cap + 1)); uint32_t* fbs = (uint32_t*)malloc(sizeof(uint32_t) * (cap + 1));
    uint32_t nf = 0; uint64_t pos = 0;
    while (pos < nsamples) {
        uint32_t bs = p->nvar ? p->var_bs[nf % p->nvar] : p->blocksize;
        if (bs > nsamples - pos) bs = (uint32_t)(nsamples - pos);
        if (nf >= cap) { cap *= 2; fstart = (uint64_t*)realloc(fstart, sizeof(uint64_t) * (cap + 1)); fbs = (uint32_t*)realloc(fbs, sizeof(uint32_t) * (cap + 1)); }
        fstart[nf] = pos; fbs[nf] = bs; nf++; pos += bs;
    }
    uint8_t** fo = (uint8_t**)calloc(nf ? nf : 1, sizeof(uint8_t*)); size_t* fl = (size_t*)calloc(nf ? nf : 1, sizeof(size_t));
    volatile uint32_t next = 0;
    job_t j = { pcm, p, nsamples, nf, fstart, fbs, p->nvar ? 1 : 0, fo, fl, &next };
    int nt = 8; const char* e = getenv("BNC_THREADS"); if (e) nt = atoi(e); if (nt < 1) nt = 1; if (nt > 64) nt = 64;
    if (nf < 8) nt = 1;
    pthread_t th[64];
    for (int t = 1; t < nt; t++) pthread_create(&th[t], NULL, enc_thread, &j);
    enc_thread(&j);
    for (int t = 1; t < nt; t++) pthread_join(th[t], NULL);
    /* md5 + sizes */
    uint8_t md5[16] = {0};
    if (!p->no_md5) {
        md5_t m; md5_init(&m);
        unsigned B = (p->bps + 7) / 8; uint8_t tmp[4096 * 3]; uint64_t n = nsamples * p->channels, i = 0;
        while (i < n) { uint64_t c = n - i; if (c > 4096) c = 4096; bnc_pack_pcm(pcm + i, c, p->bps, tmp); md5_update(&m, tmp, (size_t)c * B); i += c; }
        md5_final(&m, md5);
    }
    uint32_t minbs = 0xffffffff, maxbs = 0, minfs = 0xffffffff, maxfs = 0; size_t total = 0;
    for (uint32_t f = 0; f < nf; f++) {
        /* STREAMINFO min_blocksize excludes the short last frame of a fixed-blocksize stream (so min==max there) */
        if (!(f == nf - 1 && !p->nvar && nf > 1)) { if (fbs[f] < minbs) minbs = fbs[f]; }
        if (fbs[f] > maxbs) maxbs = fbs[f];
        if (fl[f] < minfs) minfs = (uint32_t)fl[f]; if (fl[f] > maxfs) maxfs = (uint32_t)fl[f]; total += fl[f];
    }
    if (!nf) { minbs = maxbs = p->blocksize; minfs = maxfs = 0; }
    if (!p->nvar) minbs = maxbs = p->blocksize;
    out->data = (uint8_t*)malloc(total + 256 + p->padding_bytes + 64);
    size_t q = write_metadata(out->data, p, minbs, maxbs, minfs, maxfs, nsamples, md5);
    out->first_frame = q;
    out->frame_off = (uint64_t*)malloc(sizeof(uint64_t) * (nf + 1)); out->frame_bs = fbs; out->nframes = nf;
    for (uint32_t f = 0; f < nf; f++) { out->frame_off[f] = q; memcpy(out->data + q, fo[f], fl[f]); q += fl[f]; free(fo[f]); }
    out->frame_off[nf] = q; out->len = q; out->total_samples = nsamples; memcpy(out->md5, md5, 16);
    memset(out->data + q, 0, 32);
    free(fo); free(fl); free(fstart);
    return 0;
}

/* ------------------------------------------------------------------ tiling / remux */
typedef struct { unsigned hl; unsigned num_off, num_len; int variable; uint64_t number; } hinfo;
static void parse_hdr(const uint8_t* p, hinfo* h) {
    h->variable = p[1] & 1;
    unsigned q = 4, n = 0; uint8_t x = p[q];
    if (x >= 0x80) { while (x & (0x80 >> n)) n++; } else n = 1;
    uint64_t num = (x < 0x80) ? x : (n == 7 ? 0 : (x & ((1u << (7 - n)) - 1)));
    for (unsigned i = 1; i < n; i++) num = num << 6 | (p[q + i] & 0x3f);
    h->num_off = 4; h->num_len = n; h->number = num;
    q += n;
    unsigned bsc = p[2] >> 4, src = p[2] & 15;
    if (bsc == 6) q += 1; else if (bsc == 7) q += 2;
    if (src == 12) q += 1; else if (src == 13 || src == 14) q += 2;
    h->hl = q + 1;
}
/* rewrite one frame with a new number (and optionally new blocking-strategy bit); returns new length */
static size_t rewrite_frame(const uint8_t* in, size_t inlen, uint8_t* o, int variable, uint64_t number) {
    hinfo h; parse_hdr(in, &h);
    unsigned q = 0;
    o[q++] = 0xFF; o[q++] = (uint8_t)(0xF8 | (variable ? 1 : 0)); o[q++] = in[2]; o[q++] = in[3];
    q += put_utf8(o + q, number);
    unsigned rest = h.hl - 1 - (h.num_off + h.num_len);
    memcpy(o + q, in + h.num_off + h.num_len, rest); q += rest;
    o[q] = crc8(o, q); q++;
    size_t body = inlen - h.hl - 2;
    memcpy(o + q, in + h.hl, body); q += (unsigned)body;
    uint16_t c = crc16(o, q);
    o[q++] = (uint8_t)(c >> 8); o[q++] = (uint8_t)c;
    return q;
}

typedef struct { const bnc_stream* in; uint32_t times; int variable; uint8_t** out; size_t* outlen; volatile uint32_t* next; uint64_t tile_samples; const uint64_t* fpos; } tjob;
static void* tile_thread(void* arg) {
    tjob* j = (tjob*)arg;
    for (;;) {
        uint32_t t = __sync_fetch_and_add(j->next, 1);
        if (t >= j->times) break;
        size_t cap = j->in->len - j->in->first_frame + 8 * j->in->nframes + 64;
        uint8_t* o = (uint8_t*)malloc(cap); size_t q = 0;
        for (size_t f = 0; f < j->in->nframes; f++) {
            uint64_t num = j->variable ? (uint64_t)t * j->tile_samples + j->fpos[f] : (uint64_t)t * j->in->nframes + f;
            q += rewrite_frame(j->in->data + j->in->frame_off[f], (size_t)(j->in->frame_off[f+1] - j->in->frame_off[f]), o + q, j->variable, num);
        }
        j->out[t] = o; j->outlen[t] = q;
    }
    return NULL;
}

static int tile_impl(const bnc_stream* in, uint32_t times, int to_variable, const uint8_t* pcm, size_t pcm_len, int nthreads, bnc_stream* out) {
    pthread_once(&tab_once, mk_tabs);
    memset(out, 0, sizeof *out);
    if (!in->nframes) return -1;
    hinfo h0; parse_hdr(in->data + in->frame_off[0], &h0);
    int variable = h0.variable || to_variable;
    uint64_t* fpos = (uint64_t*)malloc(sizeof(uint64_t) * (in->nframes + 1));
    uint64_t acc = 0; uint32_t bs0 = in->frame_bs[0];
    for (size_t f = 0; f < in->nframes; f++) { fpos[f] = acc; acc += in->frame_bs[f]; if (!variable && times > 1 && in->frame_bs[f] != bs0) { free(fpos); return -2; } }
    uint8_t** to = (uint8_t**)calloc(times, sizeof(uint8_t*)); size_t* tl = (size_t*)calloc(times, sizeof(size_t));
    volatile uint32_t next = 0;
    tjob j = { in, times, variable, to, tl, &next, acc, fpos };
    if (nthreads < 1) nthreads = 1; if (nthreads > 64) nthreads = 64; if ((uint32_t)nthreads > times) nthreads = (int)times;
    pthread_t th[64];
    for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, tile_thread, &j);
    tile_thread(&j);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    size_t total = 0; for (uint32_t t = 0; t < times; t++) total += tl[t];
    out->data = (uint8_t*)malloc(in->first_frame + total + 64);
    memcpy(out->data, in->data, in->first_frame);
    out->first_frame = in->first_frame;
    out->nframes = in->nframes * times;
    out->frame_off = (uint64_t*)malloc(sizeof(uint64_t) * (out->nframes + 1)); out->frame_bs = (uint32_t*)malloc(sizeof(uint32_t) * (out->nframes + 1));
    size_t q = in->first_frame, fi = 0; uint32_t minfs = 0xffffffff, maxfs = 0;
    for (uint32_t t = 0; t < times; t++) {
        /* re-derive frame offsets inside the tile by walking headers (lengths change with the number field) */
        size_t o = 0;
        for (size_t f = 0; f < in->nframes; f++) {
            hinfo hi, ho; parse_hdr(in->data + in->frame_off[f], &hi); parse_hdr(to[t] + o, &ho);
            size_t inl = (size_t)(in->frame_off[f+1] - in->frame_off[f]); size_t ol = inl - hi.hl + ho.hl;
            out->frame_off[fi] = q + o; out->frame_bs[fi] = in->frame_bs[f]; fi++;
            if (ol < minfs) minfs = (uint32_t)ol; if (ol > maxfs) maxfs = (uint32_t)ol;
            o += ol;
        }
        memcpy(out->data + q, to[t], tl[t]); q += tl[t]; free(to[t]);
    }
    out->frame_off[out->nframes] = q; out->len = q; out->total_samples = acc * times;
    memset(out->data + q, 0, 32);
    /* patch STREAMINFO: block sizes (if now variable), frame sizes, total samples, md5 */
    uint8_t* s = out->data + 8;
    if (variable) { uint32_t mn = 0xffffffff, mx = 0; for (size_t f = 0; f < in->nframes; f++) { if (in->frame_bs[f] < mn) mn = in->frame_bs[f]; if (in->frame_bs[f] > mx) mx = in->frame_bs[f]; }
        s[0] = (uint8_t)(mn >> 8); s[1] = (uint8_t)mn; s[2] = (uint8_t)(mx >> 8); s[3] = (uint8_t)mx; }
    s[4] = (uint8_t)(minfs >> 16); s[5] = (uint8_t)(minfs >> 8); s[6] = (uint8_t)minfs; s[7] = (uint8_t)(maxfs >> 16); s[8] = (uint8_t)(maxfs >> 8); s[9] = (uint8_t)maxfs;
    uint64_t tot = out->total_samples & 0xFFFFFFFFFull;
    s[13] = (uint8_t)((s[13] & 0xF0) | (tot >> 32)); s[14] = (uint8_t)(tot >> 24); s[15] = (uint8_t)(tot >> 16); s[16] = (uint8_t)(tot >> 8); s[17] = (uint8_t)tot;
    if (pcm) { md5_t m; md5_init(&m); for (uint32_t t = 0; t < times; t++) md5_update(&m, pcm, pcm_len); md5_final(&m, out->md5); }
    else memset(out->md5, 0, 16);
    memcpy(s + 18, out->md5, 16);
    free(to); free(tl); free(fpos);
    return 0;
}
int bnc_tile(const bnc_stream* in, uint32_t times, const uint8_t* pcm, size_t pcm_len, int nthreads, bnc_stream* out) { return tile_impl(in, times, 0, pcm, pcm_len, nthreads, out); }
int bnc_to_variable(const bnc_stream* in, bnc_stream* out) {
    int rc = tile_impl(in, 1, 1, NULL, 0, 1, out);
    if (!rc) { memcpy(out->md5, in->md5, 16); memcpy(out->data + 8 + 18, in->md5, 16); }
    return rc;
}

void bnc_free(bnc_stream* s) { free(s->data); free(s->frame_off); free(s->frame_bs); memset(s, 0, sizeof *s); }

#ifdef BNC_MAIN
/* bncorpus gen out.flac [out.pcm] key=value...   (see usage) */
static uint32_t kv(int argc, char** argv, const char* key, uint32_t def) {
    size_t kl = strlen(key);
    for (int i = 1; i < argc; i++) if (!strncmp(argv[i], key, kl) && argv[i][kl] == '=') return (uint32_t)strtoul(argv[i] + kl + 1, NULL, 0);
    return def;
}
int main(int argc, char** argv) {
    if (argc < 3 || strcmp(argv[1], "gen")) {
        fprintf(stderr, "usage: %s gen out.flac [pcm=out.pcm] ch= bps= sr= seconds= samples= bs= lpc= prec= minpo= maxpo= stereo= search= noise= kind= period= seed= esc= verb= var=bs1,bs2.. tile= tovar=\n", argv[0]);
        return 2;
    }
    bnc_params p; memset(&p, 0, sizeof p);
    p.channels = kv(argc, argv, "ch", 2); p.bps = kv(argc, argv, "bps", 16); p.sample_rate = kv(argc, argv, "sr", 44100);
    p.blocksize = kv(argc, argv, "bs", 4096); p.max_lpc_order = kv(argc, argv, "lpc", 8); p.qlp_precision = kv(argc, argv, "prec", 0);
    p.min_part_order = kv(argc, argv, "minpo", 0); p.max_part_order = kv(argc, argv, "maxpo", 5); p.stereo_mode = kv(argc, argv, "stereo", 1);
    p.search_order = kv(argc, argv, "search", 1); p.escape_every = kv(argc, argv, "esc", 0); p.verbatim_every = kv(argc, argv, "verb", 0);
    p.allow_zero_part = kv(argc, argv, "zeropart", 0); p.streaminfo_in_frames = kv(argc, argv, "sihdr", 0); p.padding_bytes = kv(argc, argv, "pad", 0);
    for (int i = 1; i < argc; i++) if (!strncmp(argv[i], "var=", 4)) { char* s = argv[i] + 4; while (*s && p.nvar < 16) { p.var_bs[p.nvar++] = (uint32_t)strtoul(s, &s, 10); if (*s == ',') s++; } }
    uint64_t n = kv(argc, argv, "samples", 0); if (!n) n = (uint64_t)kv(argc, argv, "seconds", 1) * p.sample_rate;
    uint32_t tile = kv(argc, argv, "tile", 1), tovar = kv(argc, argv, "tovar", 0);
    if (tile > 1 && !p.nvar) n -= n % p.blocksize;
    int32_t* pcm = (int32_t*)malloc(sizeof(int32_t) * n * p.channels);
    bnc_synth(pcm, n, p.channels, p.bps, p.sample_rate, kv(argc, argv, "noise", p.bps > 16 ? 12 : 6), kv(argc, argv, "kind", 0), kv(argc, argv, "period", p.blocksize), kv(argc, argv, "seed", 2026));
    bnc_stream s; if (bnc_encode(pcm, n, &p, &s)) { fprintf(stderr, "encode failed\n"); return 1; }
    size_t pl = (size_t)n * p.channels * ((p.bps + 7) / 8); uint8_t* packed = (uint8_t*)malloc(pl + 1); bnc_pack_pcm(pcm, n * p.channels, p.bps, packed);
    if (tovar) { bnc_stream v; if (bnc_to_variable(&s, &v)) { fprintf(stderr, "tovar failed\n"); return 1; } bnc_free(&s); s = v; }
    if (tile > 1) { bnc_stream t; int rc = bnc_tile(&s, tile, packed, pl, 8, &t); if (rc) { fprintf(stderr, "tile failed %d\n", rc); return 1; } bnc_free(&s); s = t; }
    FILE* f = fopen(argv[2], "wb"); fwrite(s.data, 1, s.len, f); fclose(f);
    for (int i = 1; i < argc; i++) if (!strncmp(argv[i], "pcm=", 4)) { f = fopen(argv[i] + 4, "wb"); for (uint32_t t = 0; t < tile; t++) fwrite(packed, 1, pl, f); fclose(f); }
    printf("frames=%zu bytes=%zu samples=%llu ratio=%.3f md5=", s.nframes, s.len, (unsigned long long)s.total_samples, (double)s.len / ((double)pl * tile));
    for (int i = 0; i < 16; i++) printf("%02x", s.md5[i]);
    printf("\n");
    return 0;
}
#endif
