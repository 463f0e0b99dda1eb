/*
 * flac_oracle.c -- CPU restatement of the reference's FLAC decode path (see flac_oracle.h for the
 * provenance statement; TEST INFRASTRUCTURE ONLY, never linked into the product).
 *
 * Follows, step by step, what the reference does per frame inside
 *   FLAC__stream_decoder_process_single (LibFLACSharp.cs:54-55, call site FLACDecoder.cs:215)
 * as published for libFLAC 1.2.1 / the FLAC format, then the byte layout of
 *   FLACDecoder.WriteCallback (FLACDecoder.cs:552-562,571-576) and
 *   FLACFileReader.CopyFlacBufferToNAudioBuffer (FLACFileReader.cs:214-243).
 */
#include "flac_oracle.h"
#include <stdlib.h>
#include <string.h>

/* Damaged frames: five rules of libFLAC 1.2.1 that were found by fuzzing this file against the reference DLL
 * (oracle/fuzz_vs_ref.py) and are pinned by tests/golden/golden_damage.json (224 reference-DLL records):
 *   1. the zero padding before the CRC-16 must be zero, else LOST_SYNC and the frame is not delivered;
 *   2. after a failure inside a frame the sync search goes on where the bit reader stands (not two bytes after the sync
 *      code) and reports LOST_SYNC again when it then skips bytes;
 *   3. qlp precision 1111 is LOST_SYNC, not UNPARSEABLE;   4. a negative qlp shift is not an error;
 *   5. a residual partition smaller than the predictor order is LOST_SYNC.
 * LOST marks a subframe failure that libFLAC reports as LOST_SYNC (the others are UNPARSEABLE_STREAM). */
#define LOST 0x100

/* ---------------------------------------------------------------- CRC (App. A.2 / A.5) */
static uint8_t crc8_tab[256];
static uint16_t crc16_tab[256];
static int tabs_ready = 0;
static void init_tabs(void) {
    if (tabs_ready) return;
    for (int i = 0; i < 256; i++) {
        uint8_t c = (uint8_t)i;
        for (int k = 0; k < 8; k++) c = (c & 0x80) ? (uint8_t)((c << 1) ^ 0x07) : (uint8_t)(c << 1);
        crc8_tab[i] = c;
        uint16_t d = (uint16_t)(i << 8);
        for (int k = 0; k < 8; k++) d = (d & 0x8000) ? (uint16_t)((d << 1) ^ 0x8005) : (uint16_t)(d << 1);
        crc16_tab[i] = d;
    }
    tabs_ready = 1;
}
uint8_t fo_crc8(const uint8_t* p, size_t n) {
    init_tabs();
    uint8_t c = 0;
    for (size_t i = 0; i < n; i++) c = crc8_tab[c ^ p[i]];
    return c;
}
uint16_t fo_crc16(const uint8_t* p, size_t n) {
    init_tabs();
    uint16_t c = 0;
    for (size_t i = 0; i < n; i++) c = (uint16_t)((c << 8) ^ crc16_tab[(c >> 8) ^ p[i]]);
    return c;
}

/* ---------------------------------------------------------------- MD5 (App. A.7; RFC 1321) */
typedef struct { uint32_t s[4]; uint64_t n; uint8_t buf[64]; } md5_t;
static const uint32_t md5_k[64] = {
    0xd76aa478,0xe8c7b756,0x242070db,0xc1bdceee,0xf57c0faf,0x4787c62a,0xa8304613,0xfd469501,0x698098d8,0x8b44f7af,0xffff5bb1,0x895cd7be,0x6b901122,0xfd987193,0xa679438e,0x49b40821,
    0xf61e2562,0xc040b340,0x265e5a51,0xe9b6c7aa,0xd62f105d,0x02441453,0xd8a1e681,0xe7d3fbc8,0x21e1cde6,0xc33707d6,0xf4d50d87,0x455a14ed,0xa9e3e905,0xfcefa3f8,0x676f02d9,0x8d2a4c8a,
    0xfffa3942,0x8771f681,0x6d9d6122,0xfde5380c,0xa4beea44,0x4bdecfa9,0xf6bb4b60,0xbebfbc70,0x289b7ec6,0xeaa127fa,0xd4ef3085,0x04881d05,0xd9d4d039,0xe6db99e5,0x1fa27cf8,0xc4ac5665,
    0xf4292244,0x432aff97,0xab9423a7,0xfc93a039,0x655b59c3,0x8f0ccc92,0xffeff47d,0x85845dd1,0x6fa87e4f,0xfe2ce6e0,0xa3014314,0x4e0811a1,0xf7537e82,0xbd3af235,0x2ad7d2bb,0xeb86d391};
static const uint8_t md5_r[64] = {7,12,17,22,7,12,17,22,7,12,17,22,7,12,17,22,5,9,14,20,5,9,14,20,5,9,14,20,5,9,14,20,
                                  4,11,16,23,4,11,16,23,4,11,16,23,4,11,16,23,6,10,15,21,6,10,15,21,6,10,15,21,6,10,15,21};
static void md5_block(md5_t* m, const uint8_t* p) {
    uint32_t w[16], a = m->s[0], b = m->s[1], c = m->s[2], d = m->s[3];
    for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4*i] | (uint32_t)p[4*i+1] << 8 | (uint32_t)p[4*i+2] << 16 | (uint32_t)p[4*i+3] << 24;
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
void fo_md5(const uint8_t* p, size_t n, uint8_t out[16]) {
    md5_t m = {{0x67452301, 0xefcdab89, 0x98badcfe, 0x10325476}, 0, {0}};
    size_t i = 0;
    for (; i + 64 <= n; i += 64) md5_block(&m, p + i);
    size_t r = n - i;
    memcpy(m.buf, p + i, r);
    m.buf[r++] = 0x80;
    if (r > 56) { memset(m.buf + r, 0, 64 - r); md5_block(&m, m.buf); r = 0; }
    memset(m.buf + r, 0, 56 - r);
    uint64_t bits = (uint64_t)n * 8;
    for (int k = 0; k < 8; k++) m.buf[56 + k] = (uint8_t)(bits >> (8*k));
    md5_block(&m, m.buf);
    for (int k = 0; k < 4; k++) for (int j = 0; j < 4; j++) out[4*k + j] = (uint8_t)(m.s[k] >> (8*j));
}

/* ---------------------------------------------------------------- bit reader (MSB first) */
typedef struct { const uint8_t* p; size_t len; uint64_t bit; int overrun; } br_t;
static inline uint32_t br_peek32(br_t* b) { /* next 32 bits, zero-extended past the end */
    size_t byte = (size_t)(b->bit >> 3);
    uint64_t v = 0;
    if (byte + 8 <= b->len) {
        const uint8_t* q = b->p + byte;
        v = (uint64_t)q[0] << 56 | (uint64_t)q[1] << 48 | (uint64_t)q[2] << 40 | (uint64_t)q[3] << 32 |
            (uint64_t)q[4] << 24 | (uint64_t)q[5] << 16 | (uint64_t)q[6] << 8 | (uint64_t)q[7];
    } else {
        for (int i = 0; i < 8; i++) v = v << 8 | (byte + i < b->len ? b->p[byte + i] : 0);
    }
    return (uint32_t)((v << (b->bit & 7)) >> 32);
}
static inline uint32_t br_u(br_t* b, unsigned n) { /* n <= 32 */
    if (n == 0) return 0;
    uint32_t v = br_peek32(b) >> (32 - n);
    b->bit += n;
    if (b->bit > (uint64_t)b->len * 8) b->overrun = 1;
    return v;
}
static inline int32_t br_s(br_t* b, unsigned n) {
    if (n == 0) return 0;
    uint32_t v = br_u(b, n);
    return (int32_t)(v << (32 - n)) >> (32 - n);
}
static inline uint32_t br_unary(br_t* b) { /* zeros before the terminating 1 */
    uint32_t q = 0;
    for (;;) {
        uint32_t w = br_peek32(b);
        if (w) { int z = __builtin_clz(w); b->bit += (unsigned)z + 1; q += (uint32_t)z; break; }
        q += 32; b->bit += 32;
        if (b->bit > (uint64_t)b->len * 8) { b->overrun = 1; return q; }
    }
    if (b->bit > (uint64_t)b->len * 8) b->overrun = 1;
    return q;
}

/* ---------------------------------------------------------------- stream header (App. A.1) */
int fo_read_streaminfo(const uint8_t* d, size_t len, fo_streaminfo* si) {
    size_t pos = 0;
    memset(si, 0, sizeof *si);
    /* libFLAC skips a leading ID3v2 tag */
    if (len >= 10 && d[0] == 'I' && d[1] == 'D' && d[2] == '3') {
        size_t sz = ((size_t)(d[6] & 0x7f) << 21) | ((size_t)(d[7] & 0x7f) << 14) | ((size_t)(d[8] & 0x7f) << 7) | (d[9] & 0x7f);
        pos = 10 + sz;
    }
    if (pos + 4 > len || memcmp(d + pos, "fLaC", 4)) return -1;
    pos += 4;
    int have = 0;
    for (;;) {
        if (pos + 4 > len) return -2;
        int last = d[pos] >> 7, type = d[pos] & 0x7f;
        size_t l = (size_t)d[pos+1] << 16 | (size_t)d[pos+2] << 8 | d[pos+3];
        pos += 4;
        if (pos + l > len) return -2;
        if (type == 0 && l >= 34 && !have) {
            const uint8_t* s = d + pos;
            si->min_blocksize = s[0] << 8 | s[1];
            si->max_blocksize = s[2] << 8 | s[3];
            si->min_framesize = s[4] << 16 | s[5] << 8 | s[6];
            si->max_framesize = s[7] << 16 | s[8] << 8 | s[9];
            uint64_t x = 0;
            for (int i = 10; i < 18; i++) x = x << 8 | s[i];
            si->sample_rate = (uint32_t)(x >> 44);
            si->channels = (uint32_t)((x >> 41) & 7) + 1;
            si->bits_per_sample = (uint32_t)((x >> 36) & 31) + 1;
            si->total_samples = x & 0xFFFFFFFFFull;
            memcpy(si->md5, s + 18, 16);
            have = 1;
        }
        pos += l;
        if (last) break;
    }
    si->first_frame_offset = pos;
    return have ? 0 : -3;
}

/* ---------------------------------------------------------------- frame decode (App. A.2-A.6) */
static const int bs_tab[16] = {0,192,576,1152,2304,4608,0,0,256,512,1024,2048,4096,8192,16384,32768};
static const int sr_tab[12] = {0,88200,176400,192000,8000,16000,22050,24000,32000,44100,48000,96000};
static const int ss_tab[8] = {0,8,12,-1,16,20,24,-1};

typedef struct {
    uint32_t blocksize, sample_rate, channels, assignment, bps, variable, header_len;
    uint64_t number;
} fhdr_t;

/* returns 0 ok, 1 bad header (crc8 / syntax), 3 unparseable (reserved), -1 truncated */
static int parse_frame_header(const uint8_t* d, size_t len, size_t pos, const fo_streaminfo* si, fhdr_t* h) {
    if (pos + 5 > len) return -1;
    const uint8_t* p = d + pos;
    int unparse = 0;
    if (p[1] & 0x02) unparse = 1;
    h->variable = p[1] & 1;
    int bsc = p[2] >> 4, src = p[2] & 15, ca = p[3] >> 4, ssc = (p[3] >> 1) & 7;
    if (p[3] & 1) unparse = 1;
    if (bsc == 0) unparse = 1;
    if (src == 15) return 1; /* libFLAC: invalid sample-rate code => bad header + resync */
    if (ca > 10) unparse = 1;
    if (ss_tab[ssc] < 0) unparse = 1;
    size_t q = 4;
    /* UTF-8 style number */
    uint32_t x = p[q++];
    uint64_t num;
    int n = 0;
    if (x < 0x80) num = x;
    else {
        while (x & (0x80u >> n)) n++;
        if (n == 1 || n > 7 || (!h->variable && n == 7)) return 1;
        num = (n == 7) ? 0 : (x & ((1u << (7 - n)) - 1));
        for (int i = 1; i < n; i++) {
            if (pos + q >= len) return -1;
            uint32_t y = p[q++];
            if ((y >> 6) != 2) return 1;
            num = num << 6 | (y & 0x3f);
        }
    }
    h->number = num;
    if (pos + q + 5 > len + 0 && pos + q + 1 > len) return -1;
    uint32_t bs = bs_tab[bsc];
    if (bsc == 6) { if (pos + q + 1 > len) return -1; bs = p[q] + 1u; q += 1; }
    else if (bsc == 7) { if (pos + q + 2 > len) return -1; bs = ((uint32_t)p[q] << 8 | p[q+1]) + 1u; q += 2; }
    uint32_t sr = src < 12 ? (uint32_t)sr_tab[src] : 0;
    if (src == 0) sr = si->sample_rate;
    else if (src == 12) { if (pos + q + 1 > len) return -1; sr = p[q] * 1000u; q += 1; }
    else if (src == 13) { if (pos + q + 2 > len) return -1; sr = (uint32_t)p[q] << 8 | p[q+1]; q += 2; }
    else if (src == 14) { if (pos + q + 2 > len) return -1; sr = ((uint32_t)p[q] << 8 | p[q+1]) * 10u; q += 2; }
    if (pos + q + 1 > len) return -1;
    if (fo_crc8(p, q) != p[q]) return 1;
    q++;
    if (unparse) return 3;
    h->blocksize = bs; h->sample_rate = sr;
    h->channels = ca < 8 ? (uint32_t)ca + 1 : 2; h->assignment = (uint32_t)ca;
    h->bps = ssc == 0 ? si->bits_per_sample : (uint32_t)ss_tab[ssc];
    h->header_len = (uint32_t)q;
    return 0;
}

static int ilog2u(uint32_t v) { int l = 0; while (v >>= 1) l++; return l; }

/* residual (App. A.4). returns 0 ok, 3 unparseable */
static int read_residual(br_t* b, int32_t* r, uint32_t bs, uint32_t order, fo_subframe* sf) {
    uint32_t method = br_u(b, 2);
    if (method > 1) return 3;
    unsigned plen = method ? 5 : 4, esc = method ? 31 : 15;
    uint32_t po = br_u(b, 4);
    if (sf) { sf->rice_method = (uint8_t)method; sf->partition_order = (uint8_t)po; }
    uint32_t nparts = 1u << po;
    if (po > 0 ? ((bs >> po) < order) : (bs < order)) return 3 + LOST; /* libFLAC: partition smaller than predictor order is LOST_SYNC (rule 5) */
    if (po > 0 && (bs & (nparts - 1))) { /* libFLAC 1.2.1 does not reject this; samples = bs>>po each */ }
    uint32_t idx = 0;
    for (uint32_t p = 0; p < nparts; p++) {
        uint32_t n = (po == 0) ? bs - order : (p == 0 ? (bs >> po) - order : (bs >> po));
        uint32_t k = br_u(b, plen);
        if (k == esc) {
            uint32_t nb = br_u(b, 5);
            for (uint32_t i = 0; i < n; i++) r[idx++] = br_s(b, nb);
        } else {
            for (uint32_t i = 0; i < n; i++) {
                uint32_t q = br_unary(b);
                uint32_t u = (q << k) | br_u(b, k);
                r[idx++] = (int32_t)(u >> 1) ^ -(int32_t)(u & 1);
                if (b->overrun) return -1;
            }
        }
        if (b->overrun) return -1;
    }
    return 0;
}

/* one subframe into out[0..bs) (App. A.3).  returns 0 ok, 3 unparseable, -1 ran off the buffer */
static int read_subframe(br_t* b, int32_t* out, int32_t* resid, uint32_t bs, uint32_t bps, fo_subframe* sf) {
    if (sf) { memset(sf, 0, sizeof *sf); sf->bit_offset = b->bit; }
    uint32_t x = br_u(b, 8);
    if (x & 0x80) return 3 + LOST; /* lost sync in libFLAC terms: pad bit set */
    uint32_t type = (x >> 1) & 0x3f, w = 0;
    if (x & 1) { w = br_unary(b) + 1; if (w >= bps) return 3; bps -= w; }
    if (sf) sf->wasted = (uint8_t)w;
    if (type == 0) {
        int32_t v = br_s(b, bps);
        for (uint32_t i = 0; i < bs; i++) out[i] = v;
        if (sf) sf->type = 0;
    } else if (type == 1) {
        for (uint32_t i = 0; i < bs; i++) out[i] = br_s(b, bps);
        if (sf) sf->type = 1;
    } else if (type >= 8 && type <= 12) {
        uint32_t o = type - 8;
        if (sf) { sf->type = 2; sf->order = (uint8_t)o; }
        if (o > bs) return 3;
        for (uint32_t i = 0; i < o; i++) out[i] = br_s(b, bps);
        int rc = read_residual(b, resid, bs, o, sf);
        if (rc) return rc;
        /* FLAC__fixed_restore_signal: 32-bit wrap-around arithmetic */
        uint32_t* s = (uint32_t*)out; const uint32_t* r = (const uint32_t*)resid;
        switch (o) {
        case 0: for (uint32_t i = 0; i < bs; i++) s[i] = r[i]; break;
        case 1: for (uint32_t i = 1; i < bs; i++) s[i] = r[i-1] + s[i-1]; break;
        case 2: for (uint32_t i = 2; i < bs; i++) s[i] = r[i-2] + 2u*s[i-1] - s[i-2]; break;
        case 3: for (uint32_t i = 3; i < bs; i++) s[i] = r[i-3] + 3u*s[i-1] - 3u*s[i-2] + s[i-3]; break;
        case 4: for (uint32_t i = 4; i < bs; i++) s[i] = r[i-4] + 4u*s[i-1] - 6u*s[i-2] + 4u*s[i-3] - s[i-4]; break;
        }
    } else if (type >= 32) {
        uint32_t o = type - 31;
        if (sf) { sf->type = 3; sf->order = (uint8_t)o; }
        if (o > bs) return 3;
        for (uint32_t i = 0; i < o; i++) out[i] = br_s(b, bps);
        uint32_t prec = br_u(b, 4) + 1;
        if (prec == 16) return 3 + LOST;      /* LOST_SYNC in libFLAC 1.2.1 (rule 3) */
        int32_t shift = br_s(b, 5);
        if (sf) { sf->precision = (uint8_t)prec; sf->shift = (uint8_t)shift; }
        int32_t c[32];
        for (uint32_t i = 0; i < o; i++) c[i] = br_s(b, prec);
        int rc = read_residual(b, resid, bs, o, sf);
        if (rc) return rc;
        if (shift < 0) shift = 0; /* rule 4: not an error in libFLAC 1.2.1 (it shifts by a negative amount); the frame then fails its CRC */
        /* width rule of libFLAC 1.2.1 (App. A.9): 32-bit accumulate iff bps+precision+ilog2(order) <= 32 */
        if (bps + prec + (uint32_t)ilog2u(o) <= 32) {
            for (uint32_t i = o; i < bs; i++) {
                uint32_t sum = 0;
                for (uint32_t j = 0; j < o; j++) sum += (uint32_t)c[j] * (uint32_t)out[i-1-j];
                out[i] = (int32_t)((uint32_t)resid[i-o] + (uint32_t)((int32_t)sum >> shift));
            }
        } else {
            for (uint32_t i = o; i < bs; i++) {
                int64_t sum = 0;
                for (uint32_t j = 0; j < o; j++) sum += (int64_t)c[j] * (int64_t)out[i-1-j];
                out[i] = (int32_t)((uint32_t)resid[i-o] + (uint32_t)(int32_t)(sum >> shift));
            }
        }
    } else return 3;
    if (w) for (uint32_t i = 0; i < bs; i++) out[i] = (int32_t)((uint32_t)out[i] << w);
    return b->overrun ? -1 : 0;
}

static inline void put_sample(uint8_t* dst, int32_t v, unsigned bytes) {
    for (unsigned k = 0; k < bytes; k++) dst[k] = (uint8_t)((uint32_t)v >> (8*k));
}

typedef struct {
    int32_t* ch[8]; int32_t* resid; uint32_t cap;
} work_t;
static int work_reserve(work_t* w, uint32_t bs) {
    if (bs <= w->cap) return 0;
    for (int c = 0; c < 8; c++) { free(w->ch[c]); w->ch[c] = (int32_t*)malloc((size_t)bs * 4 + 64); if (!w->ch[c]) return -1; }
    free(w->resid); w->resid = (int32_t*)malloc((size_t)bs * 4 + 64);
    w->cap = bs;
    return w->resid ? 0 : -1;
}
static void work_free(work_t* w) { for (int c = 0; c < 8; c++) free(w->ch[c]); free(w->resid); }

static int64_t decode_span(const uint8_t* d, size_t len, const fo_streaminfo* si, size_t pos, size_t end,
                           uint8_t* pcm, size_t pcm_cap,
                           fo_frame* frames, size_t frames_cap, size_t* nframes_out, fo_subframe* subframes,
                           uint32_t* errors, size_t errors_cap, size_t* nerrors_out) {
    init_tabs();
    work_t w; memset(&w, 0, sizeof w);
    size_t nframes = 0, nerr = 0;
    uint64_t out = 0;
    int in_sync = 1;
#define ERR(code) do { if (errors && nerr < errors_cap) errors[nerr] = (code); nerr++; } while (0)
    while (pos + 2 <= len && pos < end) {
        /* frame sync (App. A.2) */
        if (!(d[pos] == 0xFF && (d[pos+1] & 0xFC) == 0xF8)) {
            if (in_sync) { ERR(FO_ERR_LOST_SYNC); in_sync = 0; }
            pos++;
            continue;
        }
        fhdr_t h;
        int rc = parse_frame_header(d, len, pos, si, &h);
        if (rc < 0) break; /* truncated: end of stream */
        if (rc == 1) { ERR(FO_ERR_BAD_HEADER); in_sync = 1; pos += 2; continue; }   /* libFLAC reports LOST_SYNC again when it then has to skip bytes */
        if (rc == 3) { ERR(FO_ERR_UNPARSEABLE); in_sync = 1; pos += 2; continue; }
        if (work_reserve(&w, h.blocksize)) { work_free(&w); return -100; }
        br_t b = { d, len, (uint64_t)(pos + h.header_len) * 8, 0 };
        int bad = 0;
        fo_subframe sfl[8];
        for (uint32_t c = 0; c < h.channels && !bad; c++) {
            uint32_t bps = h.bps;
            if ((h.assignment == 8 && c == 1) || (h.assignment == 9 && c == 0) || (h.assignment == 10 && c == 1)) bps++;
            int r = read_subframe(&b, w.ch[c], w.resid, h.blocksize, bps, &sfl[c]);
            if (r) bad = r;
        }
        if (bad == -1 || b.overrun) break; /* ran off the end of the stream: END_OF_STREAM, frame not delivered */
        /* rule 2: the sync search goes on where the bit reader stands, and reports LOST_SYNC again when it then skips bytes */
        if (bad) { ERR(bad > 0x100 ? FO_ERR_LOST_SYNC : FO_ERR_UNPARSEABLE); in_sync = 1; pos = (size_t)((b.bit + 7) >> 3); continue; }
        /* rule 1: zero padding to the byte boundary (read_zero_padding_): bits that are not zero mean LOST_SYNC, the frame is
         * not delivered and the search goes on from the byte boundary */
        if (b.bit & 7) {
            const uint32_t nb = 8 - (uint32_t)(b.bit & 7);
            const size_t byte = (size_t)(b.bit >> 3);
            if (byte >= len) break;
            const uint32_t padv = d[byte] & ((1u << nb) - 1u);
            b.bit += nb;
            if (padv) { ERR(FO_ERR_LOST_SYNC); in_sync = 1; pos = (size_t)(b.bit >> 3); continue; }
        }
        size_t fend = (size_t)(b.bit >> 3);
        if (fend + 2 > len) break;
        uint16_t want = (uint16_t)(d[fend] << 8 | d[fend+1]);
        uint32_t status = 0;
        if (fo_crc16(d + pos, fend - pos) != want) { ERR(FO_ERR_CRC_MISMATCH); status = FO_ERR_CRC_MISMATCH; }
        /* undo inter-channel decorrelation (App. A.6), 32-bit like libFLAC */
        uint32_t bs = h.blocksize;
        if (status) { for (uint32_t c = 0; c < h.channels; c++) memset(w.ch[c], 0, (size_t)bs * 4); }
        else if (h.assignment == 8) { for (uint32_t i = 0; i < bs; i++) w.ch[1][i] = (int32_t)((uint32_t)w.ch[0][i] - (uint32_t)w.ch[1][i]); }
        else if (h.assignment == 9) { for (uint32_t i = 0; i < bs; i++) w.ch[0][i] = (int32_t)((uint32_t)w.ch[0][i] + (uint32_t)w.ch[1][i]); }
        else if (h.assignment == 10) {
            for (uint32_t i = 0; i < bs; i++) {
                int32_t side = w.ch[1][i];
                int32_t mid = (int32_t)(((uint32_t)w.ch[0][i] << 1) | ((uint32_t)side & 1));
                w.ch[0][i] = (int32_t)((uint32_t)mid + (uint32_t)side) >> 1;
                w.ch[1][i] = (int32_t)((uint32_t)mid - (uint32_t)side) >> 1;
            }
        }
        /* interleave + pack (FLACDecoder.cs:552-562; FLACFileReader.cs:214-243): LE, ceil(bps/8) bytes */
        unsigned B = (h.bps + 7) / 8;
        uint64_t need = (uint64_t)bs * h.channels * B;
        if (pcm) {
            if (out + need > pcm_cap) { work_free(&w); return -101; }
            uint8_t* o = pcm + out;
            if (B == 2 && h.channels == 2) {
                for (uint32_t i = 0; i < bs; i++) { put_sample(o, w.ch[0][i], 2); put_sample(o + 2, w.ch[1][i], 2); o += 4; }
            } else {
                for (uint32_t i = 0; i < bs; i++) for (uint32_t c = 0; c < h.channels; c++) { put_sample(o, w.ch[c][i], B); o += B; }
            }
        }
        if (frames && nframes < frames_cap) {
            fo_frame* f = &frames[nframes];
            f->offset = pos; f->length = (uint32_t)(fend + 2 - pos); f->blocksize = bs; f->channels = h.channels;
            f->bits_per_sample = h.bps; f->channel_assignment = h.assignment; f->sample_rate = h.sample_rate;
            f->variable = h.variable; f->number = h.number; f->status = status;
            if (subframes) memcpy(subframes + 8 * nframes, sfl, sizeof(fo_subframe) * h.channels);
        }
        nframes++;
        out += need;
        pos = fend + 2;
        in_sync = 1;
    }
#undef ERR
    work_free(&w);
    if (nframes_out) *nframes_out = nframes;
    if (nerrors_out) *nerrors_out = nerr;
    return (int64_t)out;
}

int64_t fo_decode(const uint8_t* data, size_t len, uint8_t* pcm, size_t pcm_cap,
                  fo_frame* frames, size_t frames_cap, size_t* nframes, fo_subframe* subframes,
                  uint32_t* errors, size_t errors_cap, size_t* nerrors) {
    fo_streaminfo si;
    int rc = fo_read_streaminfo(data, len, &si);
    if (rc) return rc;
    return decode_span(data, len, &si, (size_t)si.first_frame_offset, len, pcm, pcm_cap, frames, frames_cap, nframes, subframes,
                       errors, errors_cap, nerrors);
}

int64_t fo_decode_range(const uint8_t* data, size_t len, const fo_streaminfo* si, size_t begin, size_t end,
                        uint8_t* pcm, size_t pcm_cap, size_t* nframes) {
    return decode_span(data, len, si, begin, end, pcm, pcm_cap, NULL, 0, nframes, NULL, NULL, 0, NULL);
}

#ifdef FO_MAIN
#include <stdio.h>
static uint8_t* slurp(const char* path, size_t* n) {
    FILE* f = fopen(path, "rb"); if (!f) return NULL;
    fseek(f, 0, SEEK_END); long l = ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t* b = (uint8_t*)malloc((size_t)l + 16); if (fread(b, 1, (size_t)l, f) != (size_t)l) { fclose(f); return NULL; }
    memset(b + l, 0, 16); fclose(f); *n = (size_t)l; return b;
}
/* usage: flac_oracle dec in.flac [out.pcm]  -> prints streaminfo, frame count, md5(pcm), errors */
int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s dec in.flac [out.pcm]\n", argv[0]); return 2; }
    size_t n; uint8_t* d = slurp(argv[2], &n); if (!d) { perror("read"); return 1; }
    fo_streaminfo si; int rc = fo_read_streaminfo(d, n, &si);
    if (rc) { fprintf(stderr, "not a FLAC stream (%d)\n", rc); return 1; }
    int64_t need = fo_decode(d, n, NULL, 0, NULL, 0, NULL, NULL, NULL, 0, NULL);
    if (need < 0) { fprintf(stderr, "decode failed %lld\n", (long long)need); return 1; }
    uint8_t* pcm = (uint8_t*)malloc((size_t)need + 1);
    size_t nf = 0, ne = 0; uint32_t errs[64];
    int64_t got = fo_decode(d, n, pcm, (size_t)need, NULL, 0, &nf, NULL, errs, 64, &ne);
    uint8_t md[16]; fo_md5(pcm, (size_t)got, md);
    printf("sr=%u ch=%u bps=%u total=%llu minbs=%u maxbs=%u frames=%zu bytes=%lld errors=%zu md5=", si.sample_rate, si.channels,
           si.bits_per_sample, (unsigned long long)si.total_samples, si.min_blocksize, si.max_blocksize, nf, (long long)got, ne);
    for (int i = 0; i < 16; i++) printf("%02x", md[i]);
    printf(" si_md5="); for (int i = 0; i < 16; i++) printf("%02x", si.md5[i]);
    printf("\n");
    for (size_t i = 0; i < ne && i < 64; i++) printf("error[%zu]=%u\n", i, errs[i]);
    if (argc > 3) { FILE* f = fopen(argv[3], "wb"); fwrite(pcm, 1, (size_t)got, f); fclose(f); }
    return 0;
}
#endif
