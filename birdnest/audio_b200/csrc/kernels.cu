// kernels.cu -- sm_100a kernels of the bnflac decode pipeline (no tensor cores: nothing here is a contraction).
//
//   K1  k_scan      frame-sync scan (0xFFF8/0xFFF9) + header syntax + CRC-8  -> per-chunk sorted candidate lists
//       k_order     orders the per-tile candidate lists into one frame table (prefix of tile counts + gather)
//       k_crc       CRC-16 residue of every inter-candidate span, joined from the tile-prefix residues k_scan leaves
//       k_link      span validation (CRC residue 0 + frame/sample-number continuity), false-sync elimination
//   K2  k_parse     subframe header parse (CONSTANT / VERBATIM / FIXED 0-4 / LPC 1-32, wasted bits) and residual
//                   skip: finds where every subframe starts, so K3-5 can run one thread per (frame, channel)
//       k_prefix    accepted-frame compaction + PCM byte offsets
//   K3-5 k_decode   Rice/Rice2/escape residual decode + FIXED/LPC restore (coefficients + history in registers)
//                   + stereo decorrelation (warp shuffle between the two channel lanes) + interleave + 8/16/24-bit
//                   little-endian pack, staged through shared memory into coalesced word stores
//
// Replaces what the reference reaches through FLAC__stream_decoder_process_single (LibFLACSharp.cs:54-55,
// FLACDecoder.cs:215) and the interleave loops FLACDecoder.cs:552-576 / FLACFileReader.cs:214-243.
#include "kernels_common.cuh"

namespace bnf {

static std::atomic<int> g_launches{0};
int kernel_launch_count() { return g_launches.load(std::memory_order_relaxed); }
void count_launch() { g_launches++; }
uint32_t next_decode_epoch() {
    static std::atomic<uint32_t> ctr{0};
    uint32_t e;
    do e = ctr.fetch_add(1, std::memory_order_relaxed) + 1; while (e == 0);
    return e;
}

// ------------------------------------------------------------------------------------------------ frame header
struct Hdr {
    uint64_t number;
    uint32_t bs, sample_rate;
    uint32_t hdr_len, bps, assign, variable, channels;
};

// Frame header syntax + CRC-8.  at(i) returns byte i of the candidate; 17 bytes must be readable.
struct Crc8Table { uint8_t v[256]; };
constexpr Crc8Table make_crc8_table() {
    Crc8Table t{};
    for (int i = 0; i < 256; i++) { uint32_t c = (uint32_t)i; for (int k = 0; k < 8; k++) c = (c & 0x80) ? ((c << 1) ^ 0x07) & 0xFF : (c << 1) & 0xFF; t.v[i] = (uint8_t)c; }
    return t;
}
__constant__ Crc8Table c_crc8 = make_crc8_table();

template <class F>
__device__ __forceinline__ bool parse_header_t(F at, const SegInfo& s, Hdr& h) {
    const uint32_t b1 = at(1);
    if (at(0) != 0xFF || (b1 & 0xFE) != 0xF8) return false;
    const uint32_t b2 = at(2), b3 = at(3);
    const uint32_t bsc = b2 >> 4, src = b2 & 15, ca = b3 >> 4, ssc = (b3 >> 1) & 7;
    if ((b3 & 1) || bsc == 0 || src == 15 || ca > 10 || ssc == 3 || ssc == 7) return false;
    // the engine decodes streams whose frames agree with STREAMINFO (every real encoder; per-frame channel / bps changes
    // are treated as false syncs).  Checked before the CRC-8: it rejects ~97 % of the random sync patterns
    h.channels = ca < 8 ? ca + 1 : 2;
    h.bps = ssc == 0 ? s.bps : ssc == 1 ? 8u : ssc == 2 ? 12u : ssc == 4 ? 16u : ssc == 5 ? 20u : 24u;
    if (h.channels != s.channels || h.bps != s.bps) return false;
    h.variable = b1 & 1;
    uint32_t c = c_crc8.v[c_crc8.v[c_crc8.v[c_crc8.v[0xFF] ^ b1] ^ b2] ^ b3];
    uint32_t q = 4;
    uint32_t x = at(q++);
    c = c_crc8.v[c ^ x];
    uint64_t num;
    if (x < 0x80) num = x;
    else {
        int n = __clz(~(x << 24));   // leading ones
        if (n == 1 || n > 7 || (!h.variable && n == 7)) return false;
        num = (n == 7) ? 0 : (x & ((1u << (7 - n)) - 1));
        for (int i = 1; i < n; i++) {
            uint32_t y = at(q++);
            c = c_crc8.v[c ^ y];
            if ((y >> 6) != 2) return false;
            num = (num << 6) | (y & 0x3f);
        }
    }
    h.number = num;
    uint32_t bs;
    if (bsc == 1) bs = 192;
    else if (bsc <= 5) bs = 576u << (bsc - 2);
    else if (bsc == 6) { uint32_t v = at(q++); c = c_crc8.v[c ^ v]; bs = v + 1u; }
    else if (bsc == 7) { uint32_t v0 = at(q++), v1 = at(q++); c = c_crc8.v[c_crc8.v[c ^ v0] ^ v1]; bs = (v0 << 8 | v1) + 1u; }
    else bs = 256u << (bsc - 8);
    uint32_t sr;
    switch (src) {
    case 0: sr = s.sample_rate; break;
    case 1: sr = 88200; break; case 2: sr = 176400; break; case 3: sr = 192000; break; case 4: sr = 8000; break;
    case 5: sr = 16000; break; case 6: sr = 22050; break; case 7: sr = 24000; break; case 8: sr = 32000; break;
    case 9: sr = 44100; break; case 10: sr = 48000; break; case 11: sr = 96000; break;
    case 12: { uint32_t v = at(q++); c = c_crc8.v[c ^ v]; sr = v * 1000u; break; }
    default: { uint32_t v0 = at(q++), v1 = at(q++); c = c_crc8.v[c_crc8.v[c ^ v0] ^ v1]; sr = (v0 << 8 | v1) * (src == 13 ? 1u : 10u); break; }
    }
    if (c != at(q)) return false;
    q++;
    h.bs = bs; h.sample_rate = sr; h.hdr_len = q; h.assign = ca;
    return true;
}

// ------------------------------------------------------------------------------------------------ CRC-16 without tables
// FLAC's frame CRC polynomial factors: x^16 + x^15 + x^2 + 1 = (x + 1)(x^15 + x + 1).  A frame is intact iff its bytes,
// read as one polynomial F over GF(2), are divisible by both factors (F = M*x^16 + crc(M), and x is invertible):
//   F mod (x + 1)          is the parity of all the bits;
//   F mod Q, Q = x^15+x+1  is a trinomial residue: x^15 == x + 1, hence (Frobenius) x^(15*2^j) == x^(2^j) + 1 and in
//                          particular x^480 == x^32 + 1 -- a 32-bit word that lies 15 words above the end of a message
//                          can be XORed into the words 1 and 0 positions further down.  Folding a message down to its
//                          last 15 words therefore costs ONE three-input XOR per word and no table.
// Everything the pipeline keeps is a 16-bit "residue" r = parity << 15 | (F mod Q); r == 0 <=> CRC-16 matches.
// Residues combine like CRCs: res(A || B) = res(A) * x^(8 |B|) + res(B), the parity bit unaffected by the shift.
constexpr uint32_t QPOLY = 0x8003u;
__host__ __device__ constexpr uint32_t q_reduce(uint32_t v) { return (v & 0x7FFFu) ^ (v >> 15) ^ ((v >> 15) << 1); }   // one round: v < 2^29 -> < 2^15 if v < 2^23
__host__ __device__ constexpr uint32_t q_mul(uint32_t a, uint32_t b) {                  // a, b < 2^15
    uint32_t r = 0;
    for (int i = 14; i >= 0; --i) {
        r <<= 1;
        if (r & 0x8000u) r ^= QPOLY;
        if ((b >> i) & 1u) r ^= a;
    }
    return r;
}
__host__ __device__ constexpr uint32_t q_xpow8(uint64_t nbytes) {                       // x^(8 nbytes) mod Q
    uint32_t result = 1, base = 0x100;
    while (nbytes) {
        if (nbytes & 1u) result = q_mul(result, base);
        base = q_mul(base, base);
        nbytes >>= 1;
    }
    return result;
}
// x^(8 n) mod Q for n <= 256 (the distances inside one 256-byte piece), built at compile time
struct XPow8Table { uint16_t v[257]; };
constexpr XPow8Table make_xpow8_table() { XPow8Table t{}; for (int i = 0; i <= 256; i++) t.v[i] = (uint16_t)q_xpow8((uint64_t)i); return t; }
__constant__ XPow8Table c_xpow8 = make_xpow8_table();
__device__ __forceinline__ uint32_t q_xpow8_dev(uint64_t nbytes) { return nbytes <= 256 ? (uint32_t)c_xpow8.v[nbytes] : q_xpow8(nbytes); }
// residue of A || B from res(A), |B| in bytes and res(B)
__device__ __forceinline__ uint32_t res_append(uint32_t ra, uint64_t nbytes_b, uint32_t rb) {
    return ((ra ^ rb) & 0x8000u) | (q_mul(ra & 0x7FFFu, q_xpow8_dev(nbytes_b)) ^ (rb & 0x7FFFu));
}
// residue of up to a few hundred bytes read from global memory (frame-table join only; never on the streaming path)
__device__ uint32_t res_bytes(const uint8_t* p, uint32_t n) {
    uint32_t q = 0, par = 0, i = 0;
    for (; i < n && ((uintptr_t)(p + i) & 15u); i++) { const uint32_t b = p[i]; par ^= b; q = q_reduce((q << 8) ^ b); }
    for (; i + 16 <= n; i += 16) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p + i));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) { par ^= w[k]; q = q_reduce(q_reduce((q << 4) ^ (q << 2) ^ __byte_perm(w[k], 0, 0x0123))); }
    }
    for (; i < n; i++) { const uint32_t b = p[i]; par ^= b; q = q_reduce((q << 8) ^ b); }
    return ((__popc(par) & 1u) << 15) | q;
}

// ------------------------------------------------------------------------------------------------ K1 scan + CRC-16
// One streaming pass over the compressed bytes does both jobs of K1: it finds the frame-sync candidates (0xFFF8/0xFFF9 +
// header syntax + CRC-8) and it leaves behind what is needed to get the CRC-16 residue of ANY byte range in O(1):
// the residue of the bytes from the start of each 8 KiB tile to the end of each 256-byte piece of it.
//   * persistent, warp-autonomous: a warp owns a tile at a time (no CTA-wide barrier anywhere), double-buffered through
//     shared memory with coalesced 16-byte cp.async (units XOR-swizzled so that the per-lane reads are conflict free);
//   * a lane owns one 256-byte piece.  Per 32-bit word: one LDS.128 per four words, a four-instruction sync filter
//     (byte == 0xFF followed by a byte >= 0xF8), one XOR for the parity and one three-input XOR for the fold described
//     above; then 15 Horner steps with x^32 == x^4 + x^2 bring the piece down to 15 bits;
//   * a Kogge-Stone scan across the lanes with the constants x^(2048 d) mod Q (all of degree <= 8, so a multiplication
//     is a few shifts and one reduction round) turns piece residues into tile-prefix residues;
//   * lanes whose filter fired (about one per tile) re-read their words, check the sync codes exactly and validate the
//     header (syntax, UTF-8 number, CRC-8, agreement with STREAMINFO); the tile's candidates are appended in order.
// k_crc turns prefix residues into the residue of every span between consecutive candidates.
#ifndef SC_WARPS_N
#define SC_WARPS_N 14
#endif
constexpr int SC_WARPS = SC_WARPS_N;
#ifndef SC_CTAS_PER_SM
#define SC_CTAS_PER_SM (14 / SC_WARPS_N)      // persistent grid: this many CTAs per SM (14 warps of 16 KB double-buffered tiles fill an SM)
#endif
constexpr int SC_THREADS = 32 * SC_WARPS;
constexpr int SC_PIECE = SCAN_CHUNK / 32;                  // bytes per lane and tile
constexpr int SC_PWORDS = SC_PIECE / 4;
constexpr int SC_STAGES = 2;
constexpr uint32_t SC_SMEM = SC_WARPS * SC_STAGES * SCAN_CHUNK + SC_WARPS * 128;      // tile buffers + per-warp candidate lists (SC_LIST u16)
static_assert(SC_PIECE == 256 && SCAN_CHUNK == 8192, "scan constants below are for 256-byte pieces of 8 KiB tiles");
// x^(8 * SC_PIECE * d) mod Q for d = 1, 2, 4, 8, 16 and x^(8 * SCAN_CHUNK) mod Q
constexpr uint32_t SC_XD1 = q_xpow8(256), SC_XD2 = q_xpow8(512), SC_XD4 = q_xpow8(1024), SC_XD8 = q_xpow8(2048), SC_XD16 = q_xpow8(4096);
static_assert(SC_XD1 == 0x114 && SC_XD2 == 0x116 && SC_XD4 == 0x112 && SC_XD8 == 0x102 && SC_XD16 == 0x2 && q_xpow8(8192) == 0x4, "sparse multipliers");

__device__ __forceinline__ uint32_t sc_swz(uint32_t unit) { return (unit ^ ((unit >> 4) & 7u)) << 4; }    // 16-byte unit -> byte offset in the tile buffer
template <uint32_t C> __device__ __forceinline__ uint32_t q_mulc(uint32_t q) {           // q * C mod Q for a constant C of degree <= 8
    uint32_t v = 0;
#pragma unroll
    for (int t = 0; t <= 8; t++) if ((C >> t) & 1u) v ^= q << t;
    return q_reduce(v);
}

// byte o of a tile buffer; bytes past the buffer come from global memory (headers that straddle the tile end)
struct TileBytes {
    uint32_t dbase; uint64_t ab; const uint8_t* in; uint64_t in_len;
    __device__ __forceinline__ uint32_t operator()(uint32_t o) const {
        if (o < (uint32_t)SCAN_CHUNK) { const uint32_t wd = lds32(dbase + sc_swz(o >> 4) + (o & 12u)); return (wd >> (8 * (o & 3u))) & 0xFFu; }
        const uint64_t g = ab + o;
        return g < in_len ? in[g] : 0u;
    }
};

// Exact check of the flagged 16-byte units of a tile + header validation, by the whole warp: the flagged (lane, unit)
// pairs are visited in stream order; lanes 0..15 each look at one byte position of the unit, the (rare) lanes that see a
// sync code validate the header, and the valid ones append the tile offset to the warp's list in order.
// Returns the number of candidates found; at most SC_LIST are stored (the caller falls back to counting only).
constexpr uint32_t SC_LIST = 64;
__device__ __forceinline__ void sc_emit(const PassArgs& a, const SegInfo& seg, uint32_t seg_id, uint64_t go, const Hdr& h, uint32_t slot) {
    if (slot >= a.cand_cap) return;
    Cand cd;
    cd.off = go; cd.number = h.number; cd.bs = h.bs; cd.seg = seg_id;
    cd.hdr_len = (uint8_t)h.hdr_len; cd.bps = (uint8_t)h.bps; cd.assign = (uint8_t)h.assign;
    cd.flags = (uint8_t)(h.variable | ((go < seg.own_begin || go >= seg.own_end) ? 2u : 0u));
    cd.sample_rate = h.sample_rate;
    a.cand_tmp[slot] = cd;
}
// DIRECT: second pass for tiles with more candidates than the list holds (streams of tiny frames): the frame table slot
// of the tile is known, every valid lane writes its entry itself.
template <bool DIRECT>
__device__ __forceinline__ uint32_t sc_collect(const PassArgs& a, const TileBytes tb, const SegInfo& seg, uint32_t seg_id, uint32_t lo, uint32_t hi, uint32_t hits,
                                               uint32_t lane, uint32_t s_list, uint32_t gbase) {
    uint32_t n = 0;
    uint32_t lanes = __ballot_sync(FULL, hits != 0);
#pragma unroll 1
    while (lanes) {
        const uint32_t L = (uint32_t)__ffs(lanes) - 1u;
        lanes &= lanes - 1;
        uint32_t hl = __shfl_sync(FULL, hits, L);
#pragma unroll 1
        while (hl) {
            const uint32_t u = (uint32_t)__ffs(hl) - 1u;
            hl &= hl - 1;
            const uint32_t o = L * SC_PIECE + 16u * u + (lane & 15u);     // tile offset this lane looks at
            bool ok = false;
            Hdr h;
            if (lane < 16 && o >= lo && o < hi && tb(o) == 0xFFu && (tb(o + 1) & 0xFEu) == 0xF8u)
                ok = parse_header_t([&](uint32_t i) { return tb(o + i); }, seg, h) && tb.ab + o + h.hdr_len + 2 <= seg.end;
            const uint32_t m = __ballot_sync(FULL, ok);
            if (ok) {
                const uint32_t slot = n + __popc(m & ((1u << lane) - 1u));
                if (DIRECT) sc_emit(a, seg, seg_id, tb.ab + o, h, gbase + slot);
                else if (slot < SC_LIST) asm volatile("st.shared.u16 [%0], %1;" ::"r"(s_list + 2u * slot), "h"((uint16_t)o) : "memory");
            }
            n += __popc(m);
        }
    }
    __syncwarp();
    return n;
}

__global__ void __launch_bounds__(SC_THREADS, 1) k_scan(PassArgs a) {
    extern __shared__ __align__(128) uint8_t s_sc[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t wbase = smem_u32(s_sc) + warp * (SC_STAGES * SCAN_CHUNK);
    const uint32_t gw = blockIdx.x * SC_WARPS + warp, nwarps = gridDim.x * SC_WARPS;
    const uint32_t my_swz = sc_swz(lane);                                       // unit lane + 32 k sits at my_swz ^ (k << 5) + 512 k
    auto prefetch = [&](const Chunk& c, uint32_t buf) {
        const uint64_t ab = c.begin & ~15ull;
        const uint32_t sdst = wbase + buf * SCAN_CHUNK;
        const uint8_t* src = a.in + ab + 16u * lane;
        if (ab + SCAN_CHUNK <= a.in_len) {
#pragma unroll
            for (int k = 0; k < SCAN_CHUNK / 512; k++)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst + sc_swz(lane + 32u * k)), "l"(src + 512 * k) : "memory");
        } else {
#pragma unroll 1
            for (int k = 0; k < SCAN_CHUNK / 512; k++) {
                const uint64_t g = ab + 16ull * (lane + 32u * k);
                const uint32_t n = g + 16 <= a.in_len ? 16u : 0u;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sdst + sc_swz(lane + 32u * k)), "l"(a.in + (n ? g : 0)), "r"(n) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    (void)my_swz;

    // descriptors run two tiles ahead of the data so that no global round trip is exposed inside the loop
    uint32_t buf = 0;
    Chunk c_cur{}, c_next{};
    if (gw < a.nchunks) { c_cur = a.chunks[gw]; prefetch(c_cur, 0); }
    if (gw + nwarps < a.nchunks) c_next = a.chunks[gw + nwarps];
#pragma unroll 1
    for (uint32_t chunk = gw; chunk < a.nchunks; chunk += nwarps, buf ^= 1) {
        const Chunk c = c_cur;
        const uint64_t ab = c.begin & ~15ull;
        const uint32_t lo = (uint32_t)(c.begin - ab), hi = lo + c.len;          // valid bytes of the tile, buffer-relative
        if (chunk + nwarps < a.nchunks) { prefetch(c_next, buf ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        Chunk c_nn{};
        if (chunk + 2 * nwarps < a.nchunks) c_nn = a.chunks[chunk + 2 * nwarps];
        const uint32_t dbase = wbase + buf * SCAN_CHUNK;
        if (lo > 0 || hi < (uint32_t)SCAN_CHUNK) {
            // first / last tile of a segment: bytes outside [lo, hi) belong to something else and count as zero
#pragma unroll 1
            for (uint32_t wi = lane; wi < (uint32_t)(SCAN_CHUNK / 4); wi += 32) {
                const uint32_t b0 = 4 * wi;
                if (b0 >= lo && b0 + 4 <= hi) continue;
                uint32_t msk = 0;
                if (b0 + 4 > lo && b0 < hi) {
                    msk = 0xFFFFFFFFu;
                    if (b0 < lo) msk &= 0xFFFFFFFFu << (8 * (lo - b0));
                    if (b0 + 4 > hi) msk &= 0xFFFFFFFFu >> (8 * (b0 + 4 - hi));
                }
                const uint32_t ad = dbase + sc_swz(b0 >> 4) + (b0 & 12u);
                sts32(ad, lds32(ad) & msk);
            }
            __syncwarp();
        }

        // ---- streaming part: parity, fold, sync filter
        uint32_t f[16];
        uint32_t par = 0, hits = 0;
#pragma unroll
        for (int g = 0; g < SC_PWORDS / 16; g++) {
            uint32_t m[17];
#pragma unroll
            for (int k = 0; k < 4; k++) { const uint4 v = lds128(dbase + sc_swz(16 * lane + 4 * g + k)); m[4 * k] = v.x; m[4 * k + 1] = v.y; m[4 * k + 2] = v.z; m[4 * k + 3] = v.w; }
            // the word after the group (the byte after the tile is not in the buffer: assume a sync continuation there)
            const uint32_t un = 16 * lane + 4 * g + 4;
            m[16] = un < (uint32_t)(SCAN_CHUNK / 16) ? lds32(dbase + sc_swz(un)) : 0xF8F8F8F8u;
#pragma unroll
            for (int k4 = 0; k4 < 4; k4++) {
                uint32_t acc = 0;
#pragma unroll
                for (int kk = 0; kk < 4; kk++) {
                    const int k = 4 * k4 + kk, j = 16 * g + k;
                    const uint32_t w = m[k];
                    // sync filter: byte in {0xFE, 0xFF} followed by a byte in {0xF8, 0xF9}  <=>  z has a byte >= 0xFE
                    const uint32_t z = w & (__funnelshift_r(w, m[k + 1], 8) ^ 0x07070707u);
                    acc |= (0xFDFDFDFDu - z) & z;                             // bit 7 of a byte survives only if that byte (or a lower one) is >= 0xFE
                    par ^= w;
                    uint32_t v = w;
                    if (j >= 15) v ^= f[(j - 15) & 15];
                    if (j >= 14 && j <= SC_PWORDS - 2) v ^= f[(j - 14) & 15];
                    f[j & 15] = v;
                }
                if (acc & 0x80808080u) hits |= 1u << (4 * g + k4);
            }
        }
        // ---- the last 15 words -> 15 bits (Horner, x^32 == x^4 + x^2), then the prefix over the lanes
        uint32_t q = 0;
#pragma unroll
        for (int j = SC_PWORDS - 15; j < SC_PWORDS; j++) q = q_reduce((q << 4) ^ (q << 2) ^ __byte_perm(f[j & 15], 0, 0x0123));
        q = q_reduce(q);
        {
            uint32_t t;
            t = __shfl_up_sync(FULL, q, 1);  if (lane >= 1)  q ^= q_mulc<SC_XD1>(t);
            t = __shfl_up_sync(FULL, q, 2);  if (lane >= 2)  q ^= q_mulc<SC_XD2>(t);
            t = __shfl_up_sync(FULL, q, 4);  if (lane >= 4)  q ^= q_mulc<SC_XD4>(t);
            t = __shfl_up_sync(FULL, q, 8);  if (lane >= 8)  q ^= q_mulc<SC_XD8>(t);
            t = __shfl_up_sync(FULL, q, 16); if (lane >= 16) q ^= q_mulc<SC_XD16>(t);
        }
        const uint32_t pbits = __ballot_sync(FULL, __popc(par) & 1u);
        const uint32_t ppar = __popc(pbits & (0xFFFFFFFFu >> (31 - lane))) & 1u;
        a.pref[(uint64_t)chunk * 32 + lane] = (uint16_t)((ppar << 15) | q);

        // ---- candidates (rare): exact sync check + header validation by the whole warp, appended in order
        uint32_t n_tile = 0, gbase = 0;
        if (__any_sync(FULL, hits != 0)) {
            const SegInfo seg = a.segs[c.seg];
            const TileBytes tb{dbase, ab, a.in, a.in_len};
            const uint32_t s_list = smem_u32(s_sc) + SC_WARPS * SC_STAGES * SCAN_CHUNK + warp * (2u * SC_LIST);
            n_tile = sc_collect<false>(a, tb, seg, c.seg, lo, hi, hits, lane, s_list, 0);
            if (n_tile) {
                if (lane == 0) gbase = atomicAdd(&a.counters[0], n_tile);
                gbase = __shfl_sync(FULL, gbase, 0);
                if (n_tile > SC_LIST) sc_collect<true>(a, tb, seg, c.seg, lo, hi, hits, lane, s_list, gbase);
                else {
                    for (uint32_t e = lane; e < n_tile; e += 32) {       // one candidate per lane: header fields again, table entry
                        uint16_t o16; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(o16) : "r"(s_list + 2u * e));
                        const uint32_t o = o16;
                        Hdr h;
                        parse_header_t([&](uint32_t i) { return tb(o + i); }, seg, h);
                        sc_emit(a, seg, c.seg, ab + o, h, gbase + e);
                    }
                }
            }
        }
        if (lane == 0) {
            a.chunk_base[chunk] = gbase;
            a.chunk_count[chunk] = (gbase + n_tile > a.cand_cap) ? (gbase < a.cand_cap ? a.cand_cap - gbase : 0) : n_tile;
        }
        __syncwarp();                      // every lane is done with this buffer before the next iteration refills it
        c_cur = c_next; c_next = c_nn;
    }
}

// Orders the frame table: exclusive prefix of the per-tile candidate counts, then every tile's (already sorted) list
// moves to its place.  One launch: each CTA scans a contiguous run of tiles with coalesced loads, publishes its total,
// waits for the totals of the CTAs before it (the grid is at most one CTA per SM, so all of them are resident) and
// copies its candidates.
constexpr int ORD_THREADS = 1024;
__global__ void __launch_bounds__(ORD_THREADS) k_order(PassArgs a, uint32_t per_cta, volatile uint32_t* blk_tot) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry, s_base;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t t0 = blockIdx.x * per_cta, t1 = min(t0 + per_cta, a.nchunks);
    if (tid == 0) s_carry = 0;
    __syncthreads();
    // pass 1: local exclusive prefix (relative to the CTA's first tile) into chunk_scan
    for (uint32_t b = t0; b < t1; b += ORD_THREADS) {
        const uint32_t i = b + tid;
        const uint32_t c = i < t1 ? a.chunk_count[i] : 0u;
        uint32_t inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(FULL, inc, d); if (lane >= (uint32_t)d) inc += o; }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(FULL, wi, d); if (lane >= (uint32_t)d) wi += o; }
            s_warp[lane] = wi - w;                                   // exclusive over warps
            if (lane == 31) s_base = wi;                             // total of this batch
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        if (i < t1) a.chunk_scan[i] = carry + s_warp[warp] + inc - c;
        __syncthreads();
        if (tid == 0) s_carry = carry + s_base;
        __syncthreads();
    }
    // publish the CTA total (+1 so that 0 means "not yet"), then sum the totals of the CTAs in front
    if (tid == 0) { __threadfence(); blk_tot[blockIdx.x] = s_carry + 1u; }
    uint32_t before = 0;
    if (warp == 0) {
        for (uint32_t b = lane; b < blockIdx.x; b += 32) { uint32_t v; while ((v = blk_tot[b]) == 0u) { } before += v - 1u; }
#pragma unroll
        for (int d = 16; d; d >>= 1) before += __shfl_xor_sync(FULL, before, d);
        if (lane == 0) s_base = before;
    }
    __syncthreads();
    before = s_base;
    // pass 2: final offsets + gather
    for (uint32_t i = t0 + tid; i < t1; i += ORD_THREADS) {
        const uint32_t dst = a.chunk_scan[i] + before;
        a.chunk_scan[i] = dst;
        const uint32_t n = a.chunk_count[i];
        if (n) { const uint32_t src = a.chunk_base[i]; for (uint32_t r = 0; r < n; r++) a.cand[dst + r] = a.cand_tmp[src + r]; }
    }
}

// ------------------------------------------------------------------------------------------------ K1c CRC-16 of spans
__device__ __forceinline__ uint32_t ncand(const PassArgs& a) { return min(a.counters[0], a.cand_cap); }

__device__ __forceinline__ uint64_t span_end(const PassArgs& a, uint32_t i, uint32_t n, const Cand& ci) {
    if (i + 1 < n) { const Cand& cn = a.cand[i + 1]; if (cn.seg == ci.seg) return cn.off; }
    return a.segs[ci.seg].end;
}

// residue of the bytes from the start of p's tile up to p (exclusive).  `as_end`: p closes a range, so a p on a tile
// boundary belongs to the tile before it.  Also returns the tile index and p's offset in it.
__device__ uint32_t tile_prefix(const PassArgs& a, const SegInfo& sg, uint64_t p, bool as_end, uint32_t& tile, uint32_t& off) {
    const uint64_t ab0 = sg.begin & ~15ull;
    const uint64_t rel = p - ab0 - (as_end ? 1u : 0u);
    tile = sg.first_chunk + (uint32_t)(rel / SCAN_CHUNK);
    const uint64_t tstart = ab0 + (rel / SCAN_CHUNK) * SCAN_CHUNK;
    off = (uint32_t)(p - tstart);                                            // 0 .. SCAN_CHUNK (SCAN_CHUNK only as_end)
    const uint32_t l = off / SC_PIECE, o = off % SC_PIECE;
    uint32_t r = l ? a.pref[(uint64_t)tile * 32 + l - 1] : 0u;
    if (o) {
        uint64_t from = tstart + (uint64_t)l * SC_PIECE;
        if (from < sg.begin) from = sg.begin;                                // bytes in front of the segment count as zero
        r = res_append(r, o, from < p ? res_bytes(a.in + from, (uint32_t)(p - from)) : 0u);
    }
    return r;
}

// The CRC-16 residue of the bytes [b, e) of a segment from the tile-prefix residues k_scan left behind:
// res[b, e) = G(e) - G(b) x^(8 (e - b)) inside one tile; across tiles the tail of the first tile, whole tiles in between
// (each one multiplication by x^(8 * 8192) == x^2) and the head of the last.
__device__ uint32_t span_residue(const PassArgs& a, const SegInfo& sg, uint64_t b, uint64_t e) {
    uint32_t ta, oa, te, oe;
    const uint32_t ga = tile_prefix(a, sg, b, false, ta, oa);
    const uint32_t ge = tile_prefix(a, sg, e, true, te, oe);
    if (ta == te) return res_append(ga, oe - oa, ge);                       // GF(2): subtraction is XOR
    uint32_t r = res_append(ga, SCAN_CHUNK - oa, a.pref[(uint64_t)ta * 32 + 31]);
    for (uint32_t t = ta + 1; t < te; t++) {
        const uint32_t w = a.pref[(uint64_t)t * 32 + 31];
        r = ((r ^ w) & 0x8000u) | (q_mulc<0x4>(r & 0x7FFFu) ^ (w & 0x7FFFu));
    }
    return res_append(r, oe, ge);
}

// residue of every span between consecutive candidates
__global__ void __launch_bounds__(256) k_crc(PassArgs a) {
    const uint32_t n = ncand(a);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Cand ci = a.cand[i];
    const SegInfo sg = a.segs[ci.seg];
    a.seg_crc[i] = (uint16_t)span_residue(a, sg, ci.off, span_end(a, i, n, ci));
}

// ------------------------------------------------------------------------------------------------ K1d link / validate
__device__ __forceinline__ bool continuous(const Cand& a, const Cand& b) {
    if ((a.flags ^ b.flags) & 1) return false;
    return (a.flags & 1) ? (b.number == a.number + a.bs) : (b.number == a.number + 1);
}

__global__ void __launch_bounds__(256) k_link(PassArgs a) {
    const uint32_t n = ncand(a);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Cand ci = a.cand[i];
    const SegInfo& sg = a.segs[ci.seg];
    const uint64_t seg_end = sg.end;
    const uint32_t max_frame = sg.max_frame_bytes;
    uint32_t crc = a.seg_crc[i];
    uint32_t j = i + 1;
    uint8_t st = ST_DROP; uint32_t nx = i + 1; uint64_t fend = 0;
    for (int tries = 0; tries < 64; tries++) {
        const bool at_seg_end = !(j < n && a.cand[j].seg == ci.seg);
        const uint64_t e = at_seg_end ? seg_end : a.cand[j].off;
        if (crc == 0 && (at_seg_end || continuous(ci, a.cand[j]))) { st = ST_OK; nx = j; fend = e; break; }
        if (at_seg_end || e - ci.off > max_frame) break;
        // extend the span over candidate j (a false sync inside this frame)
        const uint64_t e2 = (j + 1 < n && a.cand[j + 1].seg == ci.seg) ? a.cand[j + 1].off : seg_end;
        crc = res_append(crc, e2 - e, a.seg_crc[j]);
        j++;
    }
    if (st != ST_OK) {
        // CRC failed over every plausible span.  If a later candidate continues the numbering, this frame is damaged
        // (the reference delivers it zero-filled).  Otherwise its successor's header is what is damaged (or this is the
        // last frame before trailing bytes): K2 validates it the way the reference does, by parsing to its end.
        uint32_t k = i + 1;
        bool found = false;
        for (;; k++) {
            const bool at_seg_end = !(k < n && a.cand[k].seg == ci.seg);
            if (at_seg_end || a.cand[k].off - ci.off > max_frame) break;
            if (continuous(ci, a.cand[k])) { st = ST_CRC; nx = k; fend = a.cand[k].off; found = true; break; }
        }
        if (!found) {
            st = ST_CHECK; nx = k;
            fend = seg_end - ci.off > (uint64_t)max_frame ? ci.off + max_frame : seg_end;
        }
    }
    a.status[i] = st;
    a.next[i] = nx;
    a.flen[i] = (uint32_t)(st == ST_DROP ? 0 : fend - ci.off);
}

// candidates that lie inside a validated frame are false syncs
__global__ void __launch_bounds__(256) k_cover(PassArgs a) {
    const uint32_t n = ncand(a);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (a.status[i] != ST_OK) return;
    const uint32_t nx = a.next[i];
    for (uint32_t j = i + 1; j < nx && j < n; j++) a.status[j] = ST_DROP;
}

// ------------------------------------------------------------------------------------------------ K2 parse
// One lane per frame: walks the subframes, records where each starts and what it is, skips the residual.  Lanes of a warp
// advance channel by channel and, inside a subframe, one refill period (16 sample indices) at a time in lockstep.  A
// period is ONE branch-free group of 16 codewords (window, bfind, add; an overflow flag instead of a branch) when every
// walking lane has at least 16 codewords left in its partition, else two half periods that each take the 8-codeword
// group or the careful per-sample path.  The loop overhead (votes, partition checks, checkpoint) is paid per period:
// it was 40 % of the instructions with 8-codeword steps (ncu, cfg3: 6.4 -> 5.3 ms; cfg2: 0.42 -> 0.39 ms).  The residual of the LAST subframe of a
// CRC-validated frame is not walked: nothing starts after it.
constexpr int PARSE_THREADS = 64;
// longest run of zero bits worth following from ring position `pos` in a frame that ends at `end_pos` (damaged frames: the
// reference follows a unary run as far as it goes; capped at 8 MiB of zeros, beyond which the frame counts as running off the stream)
__device__ __forceinline__ uint32_t unary_limit(uint32_t pos, uint32_t end_pos) { return pos < end_pos ? min(end_pos - pos, 1u << 26) : 0u; }

struct ParseSub {            // per-lane state of the residual being skipped
    uint32_t order, psize, plen, left, k, kp32, parts;
    bool first, raw;
};

// Next partition parameter(s): 0 = a partition with samples is open, 1 = no partition is left (the residual ends here:
// libFLAC 1.2.1 reads 2^order partitions of blocksize >> order samples even when that is not the whole block), 2 = an escape
// partition runs past the end of the stream.  A zero-sample partition 0 (order == partition size) is followed immediately
// by partition 1.
template <class BR>
__device__ __forceinline__ int parse_param(BR& br, ParseSub& p, const uint8_t* in, uint64_t end_bit) {
#pragma unroll 1
    for (;;) {
        if (p.parts == 0) return 1;
        p.parts--;
        const uint32_t cnt = p.psize - (p.first ? p.order : 0);
        p.first = false;
        const uint32_t k = br.get(p.plen);
        p.raw = false;
        if (k == (p.plen == 5 ? 31u : 15u)) {
            const uint32_t nb = br.get(5);
            if (br.abs_pos(in) + (uint64_t)cnt * nb > end_bit) return 2;
            br.jump(cnt * nb);
            p.raw = true;
        }
        p.k = k; p.kp32 = k + 32u; p.left = cnt;
        if (cnt) return 0;
        br.ensure_now();
    }
}

// N codewords of every walking lane's partition, branch-free: window, bfind, add; a codeword that does not fit one
// 32-bit window raises a flag and the lane redoes the group one careful codeword at a time.
// GIVE_UP (speculative jobs): a codeword longer than the window ends the job instead -- in a stream that is what it seems to be
// this (almost) never happens, while a wrong guess walks noise and would drag its warp through the careful path
template <int N, bool GIVE_UP, class BR>
__device__ __forceinline__ void parse_group(BR& br, ParseSub& ps, bool& walk, bool& bad, uint32_t end_pos) {
    uint32_t pos = br.pos;
    const uint32_t k = ps.k, kp32 = ps.kp32;
    bool ovf = false;
    typename BR::Win3 wn = br.win_init(pos);
#pragma unroll
    for (int j = 0; j < N; j++) {
        const uint32_t nxt = j < N - 1 ? br.win_next(pos) : 0u;
        const uint32_t f = bfind_fast(BR::win_peek(wn, pos));
        ovf |= (int32_t)(f - k) < 0;
        const uint32_t np = pos + kp32 - f;
        if (j < N - 1) BR::win_advance(wn, pos, np, nxt);
        pos = np;
    }
    if (walk) {
        ps.left -= N;
        if (!ps.raw) {
            if (!ovf) br.pos = pos;
            else if (GIVE_UP) { bad = true; walk = false; }
            else {
#pragma unroll 1
                for (int j = 0; j < N; j++) if (!br.rice_skip_careful(k, unary_limit(br.pos, end_pos))) { bad = true; walk = false; break; }
            }
        }
    }
}

// How the parse of a frame ended, in the reference's terms (oracle/flac_oracle.c decode_span; libFLAC 1.2.1 rules pinned by
// tests/golden/golden_damage.json): the bit reader stands where the failure was noticed, and that is where the sync search
// resumes.
enum : uint32_t { PF_NONE = 0, PF_UNPARSE = 1, PF_LOST = 2, PF_EOS = 3 };

// LEAN: branch-free ring refill (streams of few, large frames: the kernel is then a handful of warps, see RingBitsT::request)
// SPEC: the speculative form (few, large frames -- see k_spec_find): a lane is one JOB, one subframe walked from a guessed
// start, instead of one frame; it records where that subframe ends, and k_spec_resolve keeps the guesses that chain up.
template <bool LEAN, bool SPEC>
__global__ void __launch_bounds__(PARSE_THREADS) k_parse(PassArgs a) {
    extern __shared__ __align__(256) uint8_t s_ring[];
    const uint32_t n = ncand(a);
    const uint32_t lane_id = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t i = lane_id;
    uint32_t my_ch = 0, my_start = 0;
    bool have_job = !SPEC;
    if (SPEC) {
        const uint32_t njobs = min(a.counters[CNT_SPEC], a.spec_cap);
        i = n;
        if (lane_id < njobs) { const SpecJob jb = a.spec_jobs[lane_id]; i = jb.frame; my_ch = jb.ch; my_start = jb.start_bit; have_job = true; }
    }
    uint8_t st = ST_DROP;
    Cand c;
    c.bs = 0; c.assign = 0; c.flags = 0; c.off = 0; c.hdr_len = 0; c.bps = 0;
    if (i < n) { st = a.status[i]; c = a.cand[i]; }
    // CRC-validated frames are walked up to their last subframe; frames whose CRC failed (ST_CRC: a later frame continues
    // the numbering; ST_CHECK: nothing does) are parsed to the end the way the reference would, because what the
    // reference does next depends on where that parse stops (flag 2: frame of the neighbouring shard, end marker only)
    bool live = (st == ST_OK || st == ST_CHECK || st == ST_CRC) && !(c.flags & 2);
    if (SPEC) live = have_job && st == ST_OK && !(c.flags & 2);                        // guesses are only made for clean frames
    else if (live && st == ST_OK && a.spec_done && a.spec_done[i]) live = false;       // every subframe start already confirmed
    const bool clean = st == ST_OK;
    const uint32_t channels = c.assign < 8 ? c.assign + 1u : 2u;
    const uint64_t frame_bit0 = c.off * 8;
    uint64_t end_bit = 0, seg_end = 0;
    if (live) {
        const SegInfo& sg = a.segs[c.seg];
        seg_end = sg.end;
        // a damaged frame is parsed as far as its (damaged) fields say, like the reference does: only the stream ends it
        end_bit = (clean ? c.off + a.flen[i] : sg.end) * 8;
    }
    ParseBitsT<LEAN> br;
    if (live) br.init(smem_u32(s_ring) + threadIdx.x * ParseBits::STRIDE, a.in, a.in_len, frame_bit0 + (SPEC ? (uint64_t)my_start : 8ull * c.hdr_len));
    else br.init_idle(smem_u32(s_ring) + threadIdx.x * ParseBits::STRIDE, a.in);
    // end of the frame as a ring-relative bit position (saturated: a frame is far below 2^32 bits)
    const uint32_t end_pos = live ? (uint32_t)min((uint64_t)0xFFFFFFFFu, end_bit - (uint64_t)(br.g0 - a.in) * 8) : 0xFFFFFFFFu;
    uint32_t fail = PF_NONE;
    uint32_t max_order = 0, any_wide = 0;
    const uint32_t wmax_ch = __reduce_max_sync(FULL, live ? channels : 0u);
    const uint32_t wmax_bs = __reduce_max_sync(FULL, live ? c.bs : 0u);
    SubInfo spec_si; spec_si.bit_offset = 0; spec_si.type = 0; spec_si.order = 0; spec_si.wasted = 0; spec_si.flags = 0;
    for (uint32_t it = 0; it < (SPEC ? 1u : wmax_ch); it++) {
        const uint32_t ch = SPEC ? my_ch : it;          // a job is one subframe: every lane of the warp walks its own channel
        bool walk = false;            // this lane walks a Rice-coded residual in this phase
        bool bad = false;             // ... and ran off the stream doing so
        ParseSub ps;
        ps.order = 0; ps.psize = 0; ps.plen = 4; ps.left = 0; ps.k = 0; ps.kp32 = 32; ps.parts = 0; ps.first = true; ps.raw = false;
        if (live && fail == PF_NONE && ch < channels) {
            uint32_t bps = c.bps + (((c.assign == 8 && ch == 1) || (c.assign == 9 && ch == 0) || (c.assign == 10 && ch == 1)) ? 1u : 0u);
            SubInfo si;
            si.bit_offset = (uint32_t)(br.abs_pos(a.in) - frame_bit0);
            si.type = 0; si.order = 0; si.flags = 0; si.wasted = 0;
            uint32_t x = br.get(8);
            uint32_t type = (x >> 1) & 0x3f, w = 0;
            if (x & 0x80) fail = PF_LOST;                          // pad bit set: the reference reports LOST_SYNC
            else if ((x & 1) && (w = br.unary(clean ? 64u : unary_limit(br.pos, end_pos)) + 1) >= bps) fail = PF_UNPARSE;
            else {
                br.ensure_now();
                bps -= w;
                si.wasted = (uint8_t)w;
                bool has_resid = false;
                uint32_t order = 0;
                if (type == 0) br.jump(bps);
                else if (type == 1) {
                    si.type = 1;
                    if (br.abs_pos(a.in) + (uint64_t)c.bs * bps > end_bit) fail = PF_EOS; else br.jump(c.bs * bps);
                } else if (type >= 8 && type <= 12) { si.type = 2; order = type - 8; has_resid = true; }
                else if (type >= 32) { si.type = 3; order = type - 31; has_resid = true; }
                else fail = PF_UNPARSE;
                if (has_resid) {
                    si.order = (uint8_t)order;
                    if (order > c.bs) fail = PF_UNPARSE;
                    else if (br.abs_pos(a.in) + (uint64_t)order * bps > end_bit) fail = PF_EOS;
                    else {
                        if (order > max_order) max_order = order;
                        br.jump(order * bps);
                        if (si.type == 3) {
                            const uint32_t prec = br.get(4) + 1;
                            if (prec == 16) fail = PF_LOST;        // libFLAC 1.2.1: LOST_SYNC (not UNPARSEABLE), noticed before the shift is read
                            else {
                                br.skip(5);                        // quantisation shift; a negative one is not an error in libFLAC 1.2.1
                                br.jump(order * prec);
                                if (bps + prec + (uint32_t)ilog2u(order) <= 32) si.flags |= 1; else any_wide = 1;
                            }
                        } else si.flags |= 1;
                        if (fail == PF_NONE) {
                            const uint32_t method = br.get(2);
                            if (method > 1) fail = PF_UNPARSE;
                            else {
                                if (method) { si.flags |= 2; ps.plen = 5; }
                                const uint32_t po = br.get(4);
                                ps.psize = po ? c.bs >> po : c.bs;
                                ps.order = order;
                                ps.parts = 1u << po;
                                if (ps.psize < order) fail = PF_LOST;      // libFLAC 1.2.1: partition smaller than the predictor order is LOST_SYNC
                                else if (ps.psize == 0) {
                                    // blocksize < 2^po (never written by an encoder): 2^po partitions without samples, only their parameters
#pragma unroll 1
                                    for (uint32_t q = 0; q < ps.parts && fail == PF_NONE; q++) {
                                        if (br.get(ps.plen) == (ps.plen == 5 ? 31u : 15u)) br.skip(5);
                                        br.ensure_now();
                                        if (br.pos > end_pos) fail = PF_EOS;
                                    }
                                } else walk = !(clean && ch + 1 == channels);
                                br.ensure_now();
                            }
                        }
                    }
                }
            }
            if (fail != PF_NONE) walk = false;
            else if (SPEC) spec_si = si;
            else a.sub[(uint64_t)i * MAX_CH + ch] = si;
        }
        if (!__any_sync(FULL, walk)) continue;
        auto next_param = [&]() {
            const int pr = parse_param(br, ps, a.in, end_bit);
            if (pr == 2 || (SPEC && ps.raw)) bad = true;          // (a speculative job also gives up on an escape partition)
            if (pr || bad) walk = false;
        };
#pragma unroll 1
        for (uint32_t s0 = 0; s0 < wmax_bs; s0 += 16) {
            if (walk) br.checkpoint();                            // one refill checkpoint per 16 samples (ParseBits::PERIOD_REACH)
            if (walk && ps.left == 0 && s0 >= ps.order && s0 < c.bs) next_param();
            if (__all_sync(FULL, !walk || ps.left >= 16)) {
                parse_group<16, SPEC>(br, ps, walk, bad, end_pos);      // the whole period in one branch-free group
                if (walk && br.pos > end_pos) { bad = true; walk = false; }     // ran past any possible end of the frame
                if (walk && s0 + 16 >= c.bs) walk = false;
                continue;
            }
#pragma unroll 1
            for (uint32_t s1 = s0; s1 < s0 + 16; s1 += 8) {
                if (s1 != s0 && walk && ps.left == 0 && s1 >= ps.order && s1 < c.bs) next_param();
                if (__all_sync(FULL, !walk || ps.left >= 8)) parse_group<8, SPEC>(br, ps, walk, bad, end_pos);
                else {
#pragma unroll 1
                    for (uint32_t j = 0; j < 8; j++) {
                        const uint32_t s = s1 + j;
                        if (walk && s >= ps.order && s < c.bs) {
                            if (ps.left == 0) next_param();
                            if (walk) {
                                ps.left--;
                                if (!ps.raw && !br.rice_skip_careful(ps.k, unary_limit(br.pos, end_pos))) { bad = true; walk = false; }
                            }
                        }
                    }
                }
                if (walk && br.pos > end_pos) { bad = true; walk = false; }
                if (walk && s1 + 8 >= c.bs) walk = false;
            }
        }
        if (bad) fail = PF_EOS;
    }
    if (SPEC) {
        // where the subframe ended (0: the guess did not parse, or ran out of the frame)
        if (have_job) {
            const uint64_t stop_bit = live ? br.abs_pos(a.in) - frame_bit0 : 0;
            SpecJob& jb = a.spec_jobs[lane_id];
            jb.end_bit = (live && fail == PF_NONE && br.abs_pos(a.in) <= end_bit) ? (uint32_t)stop_bit : 0u;
            jb.si = spec_si;
        }
        return;
    }
    if (!live) return;
    // Outcome, in the reference's terms (oracle/flac_oracle.c decode_span).  The parse failed: nothing is delivered and the
    // sync search resumes at the byte where the bit reader stands (flen = that distance).  The parse completed: the zero
    // padding to the byte boundary must be zero (else LOST_SYNC, nothing delivered, search resumes after the padding); the
    // frame is then delivered -- zero-filled when the CRC-16 read where the parse stopped does not match -- and the search
    // resumes after that CRC.  Whatever reads past the end of the stream ends the stream (END_OF_STREAM, nothing delivered).
    const uint8_t st_in = st;
    if (fail != PF_NONE || !clean) {
        const uint64_t stop_bit = br.abs_pos(a.in);
        if (fail == PF_EOS || stop_bit > seg_end * 8) st = ST_EOS;
        else if (fail != PF_NONE) {
            st = fail == PF_UNPARSE ? ST_UNPARSEABLE : ST_LOSTSYNC;
            a.flen[i] = (uint32_t)(((stop_bit + 7) >> 3) - c.off);
        } else {
            uint64_t p = stop_bit >> 3;
            bool pad_ok = true, eos = false;
            if (stop_bit & 7) {
                if (p >= seg_end) eos = true;
                else { pad_ok = (a.in[p] & (0xFFu >> (stop_bit & 7))) == 0; p++; }
            }
            if (eos) st = ST_EOS;
            else if (!pad_ok) { st = ST_LOSTSYNC; a.flen[i] = (uint32_t)(p - c.off); }
            else if (p + 2 > seg_end) st = ST_EOS;
            else {
                a.flen[i] = (uint32_t)(p + 2 - c.off);
                st = span_residue(a, a.segs[c.seg], c.off, p + 2) == 0 ? ST_OK : ST_CRC;
            }
        }
    }
    a.status[i] = st;
    if (st == ST_OK) {
        if (max_order) atomicMax(&a.totals->max_order, max_order);
        if (any_wide) atomicOr(&a.totals->any_wide, 1u);
    }
    if (st != ST_OK || st_in != ST_OK) {                               // off the clean chain: k_resync decides what follows it
        const uint32_t slot = atomicAdd(&a.counters[CNT_ANOM], 1u);
        if (slot < ANOM_CAP) a.anom[slot] = i;
    }
}

// ------------------------------------------------------------------------------------------------ speculative parse
// Streams of few, large frames (BASELINE cfg3: 7,000 frames of 8 x 16,384 samples).  k_parse finds where the subframes of
// a frame start by walking them one after the other -- a serial chain of channels x blocksize codewords per frame that no
// number of SMs shortens (4.5 ms for cfg3, with 97 % of the issue slots empty).  The walk of subframe c+1 could run beside
// the walk of subframe c if only its start were known: so it is GUESSED, all guesses are walked in parallel (k_parse in job
// mode), and a guess is kept only when the walk of the subframe before it ends exactly there -- channel 0's start is known,
// so by induction every kept start is the true one, and the result is the serial parse's, bit for bit.  Frames whose
// guesses do not chain up are simply left to the serial kernel.
// The guess: an encoder that does not search per subframe (libFLAC without exhaustive model search, the corpus generator)
// writes every subframe of a frame with the same header byte, quantisation precision, Rice method and partition order;
// those 14 (FIXED) / 19 (LPC) bits, taken from subframe 0, are looked for around the place where subframe c would start
// if all subframes were equally long.
constexpr int SPEC_THREADS = 128;
__device__ __forceinline__ uint32_t load_be32(const uint8_t* in, uint64_t word) { return __byte_perm(__ldg(reinterpret_cast<const uint32_t*>(in) + word), 0, 0x0123); }
// n <= 25 bits at absolute bit position `bit`, MSB first
__device__ __forceinline__ uint32_t gbits(const uint8_t* in, uint64_t bit, uint32_t n) {
    const uint64_t w = bit >> 5; const uint32_t sh = (uint32_t)bit & 31u;
    return __funnelshift_l(load_be32(in, w + 1), load_be32(in, w), sh) >> (32u - n);
}
__device__ __forceinline__ uint32_t side_bit(uint32_t assign, uint32_t ch) { return ((assign == 8 && ch == 1) || (assign == 9 && ch == 0) || (assign == 10 && ch == 1)) ? 1u : 0u; }

__global__ void __launch_bounds__(SPEC_THREADS) k_spec_find(PassArgs a) {
    constexpr uint32_t SPEC_QUEUE = 2048;
    __shared__ uint32_t s_n, s_seen, s_nq;
    __shared__ uint32_t s_list[SPEC_MAX_PER_FRAME];
    __shared__ uint32_t s_queue[SPEC_QUEUE];
    const uint32_t n = ncand(a), i = blockIdx.x, tid = threadIdx.x;
    if (i >= n) return;
    if (tid == 0) { a.spec_count[i] = 0; a.spec_done[i] = 0; s_n = 0; s_seen = 0; s_nq = 0; }
    const Cand c = a.cand[i];
    const uint32_t C = c.assign < 8 ? c.assign + 1u : 2u;
    if (a.status[i] != ST_OK || (c.flags & 2) || C < 3) return;          // (two channels: the only start to find is the last one, which needs no guess)
    // subframe 0: the signature (every thread reads the same few words: broadcast from L2 / L1)
    const uint64_t fb0 = c.off * 8;
    const uint32_t start0 = 8u * c.hdr_len, end = a.flen[i] * 8u - 16u;          // frame-relative bits; `end`: where the CRC-16 starts
    if (end <= start0 + 64u) return;
    const uint32_t x = gbits(a.in, fb0 + start0, 8);
    if (x & 0x81u) return;                                                      // pad bit / wasted bits: not guessed
    const uint32_t type = (x >> 1) & 0x3fu;
    const bool lpc = type >= 32;
    if (!lpc && !(type >= 8 && type <= 12)) return;                             // CONSTANT / VERBATIM / reserved in front: serial
    const uint32_t order = lpc ? type - 31u : type - 8u;
    if (order > c.bs) return;
    uint32_t q = start0 + 8u + order * (c.bps + side_bit(c.assign, 0));
    uint32_t prec = 0;
    if (lpc) { prec = gbits(a.in, fb0 + q, 4) + 1u; if (prec == 16) return; q += 9u + order * prec; }
    if (q + 6u >= end) return;
    const uint32_t tail0 = gbits(a.in, fb0 + q, 6);                             // Rice method (2) + partition order (4)
    if (tail0 >= 32u) return;                                                   // reserved method
    __syncthreads();
    // Phase 1: every place in the windows where the header byte of subframe 0 shows up again (one in 256 positions) goes into a
    // queue, four byte-aligned positions per test.  Phase 2: the queue is checked against the rest of the signature with all
    // lanes busy -- done inside the scan loop, the rare hit made its whole warp wait.
    const uint32_t avg = (end - start0) / C;
    const uint32_t wd = 2048u + avg / 64u;
    const uint32_t pat = x * 0x01010101u;
    for (uint32_t ch = 1; ch + 1 < C; ch++) {          // the last subframe starts where the one before it ends and is not walked: no guess needed
        const uint32_t f2 = 8u + order * (c.bps + side_bit(c.assign, ch)) + (lpc ? 9u + order * prec : 0u);
        const uint32_t centre = start0 + (uint32_t)(((uint64_t)(end - start0) * ch) / C);
        const uint32_t lo = max(start0 + 8u, centre > wd ? centre - wd : 0u), hi = min(end > f2 + 8u ? end - f2 - 8u : 0u, centre + wd);   // guesses in [lo, hi)
        if (lo >= hi) continue;
        const uint64_t A0 = fb0 + lo, A1 = fb0 + hi;                            // absolute bit positions
        for (uint64_t w = (A0 >> 5) + tid; w <= ((A1 - 1) >> 5); w += SPEC_THREADS) {
            const uint32_t w0 = load_be32(a.in, w), w1 = load_be32(a.in, w + 1);
            uint32_t any = 0, zz[8];
#pragma unroll
            for (uint32_t sft = 0; sft < 8; sft++) {
                const uint32_t v = __funnelshift_l(w1, w0, sft) ^ pat;          // bytes at bit positions 32 w + sft + 8 j
                zz[sft] = (v - 0x01010101u) & ~v & 0x80808080u;                // a zero byte (may also flag the byte above one)
                any |= zz[sft];
            }
            if (!any) continue;
#pragma unroll
            for (uint32_t sft = 0; sft < 8; sft++) {
                uint32_t z = zz[sft];
                while (z) {
                    const uint32_t b = 31u - (uint32_t)__clz(z);               // bit 8 (3 - j) + 7
                    z &= ~(1u << b);
                    const uint64_t pa = (w << 5) + sft + 8u * (3u - (b >> 3));
                    if (pa < A0 || pa >= A1) continue;
                    const uint32_t slot = atomicAdd(&s_nq, 1u);
                    if (slot < SPEC_QUEUE) s_queue[slot] = (ch << 28) | (uint32_t)(pa - fb0);
                }
            }
        }
    }
    __syncthreads();
    const uint32_t nq = s_nq;
    if (nq > SPEC_QUEUE) return;                                                // a window full of look-alikes: serial
    for (uint32_t e = tid; e < nq; e += SPEC_THREADS) {
        const uint32_t v = s_queue[e], ch = v >> 28;
        const uint64_t pa = fb0 + (v & 0x0FFFFFFFu);
        const uint32_t f1 = 8u + order * (c.bps + side_bit(c.assign, ch)), f2 = f1 + (lpc ? 9u + order * prec : 0u);
        if (gbits(a.in, pa, 8) != x) continue;                                  // (the zero-byte test over-reports)
        if (lpc && gbits(a.in, pa + f1, 5) != ((prec - 1u) << 1)) continue;     // precision, and a shift that is not negative
        if (gbits(a.in, pa + f2, 6) != tail0) continue;
        const uint32_t slot = atomicAdd(&s_n, 1u);
        if (slot < SPEC_MAX_PER_FRAME - 1u) s_list[slot] = v;
        atomicOr(&s_seen, 1u << ch);
    }
    __syncthreads();
    __shared__ uint32_t s_base;
    const uint32_t cnt = s_n;
    if (cnt > SPEC_MAX_PER_FRAME - 1u || s_seen != ((1u << (C - 1u)) - 2u)) return;      // too many look-alikes, or a channel without any guess
    if (tid == 0) s_base = atomicAdd(&a.counters[CNT_SPEC], cnt + 1u);
    __syncthreads();
    const uint32_t base = s_base;
    if (base + cnt + 1u > a.spec_cap) return;
    for (uint32_t e = tid; e <= cnt; e += SPEC_THREADS) {
        SpecJob jb;
        jb.frame = i; jb.end_bit = 0; jb.si.bit_offset = 0; jb.si.type = 0; jb.si.order = 0; jb.si.wasted = 0; jb.si.flags = 0;
        if (e == 0) { jb.start_bit = start0; jb.ch = 0; }
        else { const uint32_t v = s_list[e - 1]; jb.start_bit = v & 0x0FFFFFFFu; jb.ch = v >> 28; }
        a.spec_jobs[base + e] = jb;
    }
    if (tid == 0) { a.spec_base[i] = base; a.spec_count[i] = (uint8_t)(cnt + 1u); }
}

// What k_parse records for a subframe it does not walk (the last one of a CRC-validated frame), read straight from the
// stream; false for anything the serial walk should see for itself (wasted bits, reserved or inconsistent fields).
__device__ bool peek_subframe(const uint8_t* in, uint64_t frame_bit0, uint32_t start, uint32_t end, uint32_t bps, uint32_t bs, SubInfo& si) {
    si.bit_offset = start; si.type = 0; si.order = 0; si.wasted = 0; si.flags = 0;
    if (start + 8u + bps > end) return false;
    const uint32_t x = gbits(in, frame_bit0 + start, 8);
    if (x & 0x81u) return false;
    const uint32_t type = (x >> 1) & 0x3fu;
    if (type == 0) return true;
    if (type == 1) { si.type = 1; return (uint64_t)start + 8u + (uint64_t)bs * bps <= end; }
    uint32_t order, q;
    if (type >= 8 && type <= 12) { si.type = 2; order = type - 8u; si.flags = 1; q = start + 8u + order * bps; }
    else if (type >= 32) {
        si.type = 3; order = type - 31u;
        q = start + 8u + order * bps;
        if (order > bs || q + 9u >= end) return false;
        const uint32_t prec = gbits(in, frame_bit0 + q, 4) + 1u;
        if (prec == 16) return false;
        if (bps + prec + (uint32_t)ilog2u(order) <= 32) si.flags = 1;
        q += 9u + order * prec;
    } else return false;
    si.order = (uint8_t)order;
    if (order > bs || q + 6u > end) return false;
    const uint32_t tail = gbits(in, frame_bit0 + q, 6), po = tail & 15u;
    if (tail >= 32u) return false;
    if (tail & 16u) si.flags |= 2;
    const uint32_t psize = po ? bs >> po : bs;
    return psize != 0 && psize >= order;
}

// keeps the guesses that chain up: subframe c + 1 starts where the walk of subframe c ended
__global__ void __launch_bounds__(128) k_spec_resolve(PassArgs a) {
    const uint32_t n = ncand(a);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t cnt = a.spec_count[i];
    if (!cnt) return;
    const uint32_t base = a.spec_base[i];
    const Cand c = a.cand[i];
    const uint32_t C = c.assign < 8 ? c.assign + 1u : 2u;
    SubInfo sub[MAX_CH];
    uint32_t e = a.spec_jobs[base].end_bit;
    if (!e) return;
    sub[0] = a.spec_jobs[base].si;
    for (uint32_t ch = 1; ch + 1 < C; ch++) {
        uint32_t hit = 0;
        for (uint32_t k = 1; k < cnt; k++) { const SpecJob& jb = a.spec_jobs[base + k]; if (jb.ch == ch && jb.start_bit == e && jb.end_bit) { hit = k; break; } }
        if (!hit) return;                                                          // not guessed (or the guess did not parse): the serial walk takes the frame
        sub[ch] = a.spec_jobs[base + hit].si;
        e = a.spec_jobs[base + hit].end_bit;
    }
    if (!peek_subframe(a.in, c.off * 8, e, a.flen[i] * 8u - 16u, c.bps + side_bit(c.assign, C - 1u), c.bs, sub[C - 1u])) return;
    uint32_t max_order = 0, any_wide = 0;
    for (uint32_t ch = 0; ch < C; ch++) {
        a.sub[(uint64_t)i * MAX_CH + ch] = sub[ch];
        if (sub[ch].type >= 2) max_order = max(max_order, (uint32_t)sub[ch].order);
        if (sub[ch].type == 3 && !(sub[ch].flags & 1)) any_wide = 1;
    }
    if (max_order) atomicMax(&a.totals->max_order, max_order);
    if (any_wide) atomicOr(&a.totals->any_wide, 1u);
    a.spec_done[i] = 1;
    atomicAdd(&a.counters[CNT_SPEC_DONE], 1u);
}

// ------------------------------------------------------------------------------------------------ resync
// The reference is a sequential decoder: after a damaged frame it resumes its sync search wherever its parse of the
// damaged bits stopped, so a frame can be lost because the one before it was damaged.  Intact frames chain exactly (each
// ends where the next starts) and need no attention; k_parse lists the few candidates that are off that chain.  One
// thread walks that list in stream order: it computes where the reference's cursor lands after each of them and marks
// the candidates the cursor jumps over as never reached.  Candidates inside CRC-validated frames (ST_DROP) are not
// reconsidered.  Nothing to do (and nothing done) for intact streams.
__global__ void __launch_bounds__(256) k_resync(PassArgs a) {
    __shared__ uint32_t s_sorted[1024];
    const uint32_t n = ncand(a);
    const uint32_t m_all = a.counters[CNT_ANOM];
    if (m_all == 0) return;
    const uint32_t tid = threadIdx.x;
    auto step = [&](uint32_t ai) {
        const uint8_t st = a.status[ai];
        if (st == ST_SKIP || st == ST_DROP) return;
        const Cand ca = a.cand[ai];
        if (ca.flags & 2) return;
        const uint64_t landing = st == ST_EOS ? ~0ull : ca.off + (uint64_t)a.flen[ai];      // where the reference's cursor stands after this candidate
        for (uint32_t j = ai + 1; j < n; j++) {
            const Cand& cj = a.cand[j];
            if (cj.seg != ca.seg) break;
            const uint8_t sj = a.status[j];
            if (sj == ST_DROP) continue;
            if (cj.off >= landing) break;
            a.status[j] = ST_SKIP;
        }
    };
    if (m_all > ANOM_CAP) {                    // heavily damaged input: walk the whole table
        if (tid == 0) for (uint32_t i = 0; i < n; i++) step(i);
        return;
    }
    // rank sort of the (unordered) list, 1024 entries at a time is not needed: ranks are global
    for (uint32_t base = 0; base < m_all; base += 1024) {
        __syncthreads();
        for (uint32_t e = tid; e < m_all; e += blockDim.x) {
            const uint32_t v = a.anom[e];
            uint32_t rank = 0;
            for (uint32_t o = 0; o < m_all; o++) rank += a.anom[o] < v;
            if (rank >= base && rank < base + 1024) s_sorted[rank - base] = v;
        }
        __syncthreads();
        if (tid == 0) { const uint32_t cnt = min(1024u, m_all - base); for (uint32_t e = 0; e < cnt; e++) step(s_sorted[e]); }
    }
}

// ------------------------------------------------------------------------------------------------ prefix
// Accepted-frame compaction + PCM byte offsets.  Same shape as k_order: every CTA reduces a contiguous run of the frame
// table, publishes (bytes, count), sums what the CTAs in front published, then scans its run and writes the offsets.
__device__ __forceinline__ bool frame_delivered(const PassArgs& a, uint32_t i, uint32_t& bs, uint32_t& ch) {
    const uint8_t st = a.status[i];
    const Cand& c = a.cand[i];
    bs = c.bs; ch = c.assign < 8 ? c.assign + 1u : 2u;
    return (st == ST_OK || st == ST_CRC) && !(c.flags & 2);
}
__global__ void __launch_bounds__(ORD_THREADS) k_prefix(PassArgs a, uint32_t bytes_per_sample, uint32_t per_cta, volatile uint32_t* blk_cnt, volatile unsigned long long* blk_bytes) {
    __shared__ unsigned long long s_wb[32];
    __shared__ uint32_t s_wc[32];
    __shared__ unsigned long long s_cb, s_tb;
    __shared__ uint32_t s_cc, s_tc;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n = ncand(a);
    const uint32_t t0 = min(blockIdx.x * per_cta, n), t1 = min(t0 + per_cta, n);
    // pass 1: totals of this CTA's run
    unsigned long long bytes = 0; uint32_t cnt = 0, maxbs = 0;
    for (uint32_t i = t0 + tid; i < t1; i += ORD_THREADS) {
        uint32_t bs, ch;
        if (frame_delivered(a, i, bs, ch)) { bytes += (unsigned long long)bs * ch * bytes_per_sample; cnt++; maxbs = max(maxbs, bs); }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) { bytes += __shfl_xor_sync(FULL, bytes, d); cnt += __shfl_xor_sync(FULL, cnt, d); maxbs = max(maxbs, __shfl_xor_sync(FULL, maxbs, d)); }
    if (lane == 0) { s_wb[warp] = bytes; s_wc[warp] = cnt; if (maxbs) atomicMax(&a.totals->max_bs, maxbs); }
    __syncthreads();
    if (warp == 0) {
        bytes = s_wb[lane]; cnt = s_wc[lane];
#pragma unroll
        for (int d = 16; d; d >>= 1) { bytes += __shfl_xor_sync(FULL, bytes, d); cnt += __shfl_xor_sync(FULL, cnt, d); }
        if (lane == 0) { s_tb = bytes; s_tc = cnt; blk_bytes[blockIdx.x] = bytes; __threadfence(); blk_cnt[blockIdx.x] = cnt + 1u; }
        unsigned long long bb = 0; uint32_t bc = 0;
        for (uint32_t b = lane; b < blockIdx.x; b += 32) { uint32_t v; while ((v = blk_cnt[b]) == 0u) { } __threadfence(); bc += v - 1u; bb += blk_bytes[b]; }
#pragma unroll
        for (int d = 16; d; d >>= 1) { bb += __shfl_xor_sync(FULL, bb, d); bc += __shfl_xor_sync(FULL, bc, d); }
        if (lane == 0) { s_cb = bb; s_cc = bc; }
    }
    __syncthreads();
    if (blockIdx.x == gridDim.x - 1 && tid == 0) {
        a.totals->pcm_bytes = s_cb + s_tb; a.totals->n_accepted = s_cc + s_tc; a.totals->n_cand = n; a.totals->overflow = a.counters[1];
    }
    // pass 2: offsets
    for (uint32_t b = t0; b < t1; b += ORD_THREADS) {
        const uint32_t i = b + tid;
        uint32_t bs = 0, ch = 0;
        const bool ok = i < t1 && frame_delivered(a, i, bs, ch);
        const unsigned long long v = ok ? (unsigned long long)bs * ch * bytes_per_sample : 0ull;
        unsigned long long ib = v; uint32_t ic = ok ? 1u : 0u;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned long long ob = __shfl_up_sync(FULL, ib, d); const uint32_t oc = __shfl_up_sync(FULL, ic, d); if (lane >= (uint32_t)d) { ib += ob; ic += oc; } }
        __syncthreads();                               // (s_wb / s_wc of the previous batch have been read)
        if (lane == 31) { s_wb[warp] = ib; s_wc[warp] = ic; }
        __syncthreads();
        unsigned long long wb = 0; uint32_t wc = 0;
        for (uint32_t w = 0; w < warp; w++) { wb += s_wb[w]; wc += s_wc[w]; }
        const unsigned long long cb = s_cb; const uint32_t cc = s_cc;
        if (i < t1) {
            a.pcm_off[i] = cb + wb + ib - v;
            if (ok) a.acc_idx[cc + wc + ic - 1u] = i;
        }
        __syncthreads();
        if (tid == ORD_THREADS - 1) { s_cb = cb + wb + ib; s_cc = cc + wc + ic; }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ blocksize classes
// k_decode gives a warp 32 / channels consecutive frames and runs until the longest of them is done.  In a stream of
// variable blocksize (0xFFF9: 4096, 1152, 4080, 720, 16, 192, 2304 samples in turn) that is the 4096-sample frame every
// time, and the lanes of the short frames idle: 122 G samples/s against 286 for the same audio at a fixed blocksize.  The
// accepted frames are therefore regrouped by half-octave blocksize class, largest first (where a frame's PCM goes is
// pcm_off's business, not the order of decoding): per-CTA class counts, one warp turns them into start positions, scatter.
__device__ __forceinline__ uint32_t bs_class(uint32_t bs) {          // 0 = largest
    const uint32_t l = (uint32_t)ilog2u(bs | 1u);
    return 31u - min(31u, 2u * l + (l ? (bs >> (l - 1)) & 1u : 0u));
}
__global__ void __launch_bounds__(256) k_bucket_hist(PassArgs a) {
    __shared__ uint32_t s_h[32];
    const uint32_t n_acc = a.totals->n_accepted;
    const uint32_t k0 = blockIdx.x * BUCKET_CHUNK;
    if (threadIdx.x < 32) s_h[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t k = k0 + threadIdx.x; k < min(k0 + BUCKET_CHUNK, n_acc); k += blockDim.x) atomicAdd(&s_h[bs_class(a.cand[a.acc_idx[k]].bs)], 1u);
    __syncthreads();
    if (threadIdx.x < 32) a.bucket_hist[blockIdx.x * 32 + threadIdx.x] = s_h[threadIdx.x];
}
__global__ void __launch_bounds__(32) k_bucket_scan(PassArgs a, uint32_t nblocks) {
    const uint32_t c = threadIdx.x;
    uint32_t tot = 0;
    for (uint32_t b = 0; b < nblocks; b++) tot += a.bucket_hist[b * 32 + c];
    uint32_t inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(FULL, inc, d); if (c >= (uint32_t)d) inc += o; }
    uint32_t run = inc - tot;                                         // where class c starts
    for (uint32_t b = 0; b < nblocks; b++) { const uint32_t h = a.bucket_hist[b * 32 + c]; a.bucket_hist[b * 32 + c] = run; run += h; }
}
__global__ void __launch_bounds__(256) k_bucket_scatter(PassArgs a) {
    __shared__ uint32_t s_cur[32];
    const uint32_t n_acc = a.totals->n_accepted;
    const uint32_t k0 = blockIdx.x * BUCKET_CHUNK;
    if (threadIdx.x < 32) s_cur[threadIdx.x] = a.bucket_hist[blockIdx.x * 32 + threadIdx.x];
    __syncthreads();
    for (uint32_t k = k0 + threadIdx.x; k < min(k0 + BUCKET_CHUNK, n_acc); k += blockDim.x) {
        const uint32_t i = a.acc_idx[k];
        a.acc_sorted[atomicAdd(&s_cur[bs_class(a.cand[i].bs)], 1u)] = i;
    }
}

// start-of-pass reset and end-of-stage hand-off to the host.  The host learns the candidate count and the totals through
// a few words of MAPPED pinned memory written by k_publish, not through cudaMemcpy: a small copy would queue on the copy
// engines behind the multi-megabyte uploads/downloads of the other sub-shards of a pipelined decode and stall the pass.
// Scan-tile descriptors of a one-segment pass, generated where they are used: tile k of the segment covers
// [max(begin, A + k * SCAN_CHUNK), min(end, A + (k + 1) * SCAN_CHUNK)), A = begin & ~15.  The segment record travels as a kernel
// argument, so opening a stream costs no host-built table, no small upload queued behind bulk copies and no synchronisation.
__global__ void __launch_bounds__(256) k_make_chunks(SegInfo seg, SegInfo* d_seg, Chunk* chunks, uint32_t nchunks) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) d_seg[0] = seg;
    if (k >= nchunks) return;
    const uint64_t A = seg.begin & ~15ull;
    const uint64_t b = max(seg.begin, A + (uint64_t)k * SCAN_CHUNK), e = min(seg.end, A + (uint64_t)(k + 1) * SCAN_CHUNK);
    chunks[k] = Chunk{b, (uint32_t)(e - b), 0u};
}

__global__ void __launch_bounds__(256) k_clear(uint32_t* counters, Totals* totals) {
    for (uint32_t i = threadIdx.x; i < CNT_WORDS; i += blockDim.x) counters[i] = 0;     // counters, k_order / k_prefix CTA totals
    if (threadIdx.x < sizeof(Totals) / 4) reinterpret_cast<uint32_t*>(totals)[threadIdx.x] = 0;
}
__global__ void k_publish(const uint32_t* src, uint32_t* dst_mapped, uint32_t nwords) {
    if (threadIdx.x < nwords) dst_mapped[threadIdx.x] = src[threadIdx.x];
    __threadfence_system();
}

// per-segment (clip) summary for batched passes: where the segment's PCM starts and whether any of its frames is damaged
__global__ void __launch_bounds__(256) k_seg_summary(PassArgs a, uint64_t* seg_pcm, uint32_t* seg_flags) {
    const uint32_t n = ncand(a);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t sg = a.cand[i].seg;
    if (i == 0 || a.cand[i - 1].seg != sg) seg_pcm[sg] = a.pcm_off[i];
    const uint8_t st = a.status[i];
    if (st == ST_CRC) atomicOr(&seg_flags[sg], 1u);
    else if (st == ST_UNPARSEABLE) atomicOr(&seg_flags[sg], 8u);
}

// ------------------------------------------------------------------------------------------------ launchers

// Function attributes (opt-in shared memory) and the SM count belong to a DEVICE, and one process may hold handles on
// several (bnflac_opts.device): both are looked up / set per current device, once per (kernel, device).
constexpr int MAX_DEVICES = 64;
static int current_device() { int dev = 0; if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) dev = 0; return dev; }
int sm_count() {
    static std::atomic<int> n_sm[MAX_DEVICES];
    const int dev = current_device();
    int n = n_sm[dev].load(std::memory_order_relaxed);
    if (!n) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n < 1) n = 148;
        if (n > 256) n = 256;          // CNT_ORDER / CNT_PFX_* hold one slot per CTA of the look-back kernels
        n_sm[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}
// Once-per-(kernel, device) work such as opting in to large dynamic shared memory.  first_use_on_device() says whether the work is still to be
// done on the current device, used_on_device() records it AFTER it has been done: two threads that start their first pass at the same moment
// (handles on different host threads) both do the idempotent work -- neither can launch before its own attribute call has returned.
// (Setting the flag first let the second thread launch k_scan with 226 KB of dynamic shared memory before the first had opted in:
// "invalid argument" on the very first concurrent decode of a process.)
bool first_use_on_device(std::atomic<uint64_t>& done) { return !(done.load(std::memory_order_acquire) & (1ull << current_device())); }
void used_on_device(std::atomic<uint64_t>& done) { done.fetch_or(1ull << current_device(), std::memory_order_release); }

void launch_scan(const PassArgs& a, void* stream) {
    static std::atomic<uint64_t> attr_done{0};
    if (first_use_on_device(attr_done)) { cudaFuncSetAttribute(k_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SC_SMEM); used_on_device(attr_done); }
    const int n_sm = sm_count();
    uint32_t grid = (a.nchunks + SC_WARPS - 1) / SC_WARPS;
    if (grid > (uint32_t)n_sm * SC_CTAS_PER_SM) grid = (uint32_t)n_sm * SC_CTAS_PER_SM;
    if (!grid) grid = 1;
    k_scan<<<grid, SC_THREADS, SC_SMEM, S(stream)>>>(a); count_launch();
}
void launch_order(const PassArgs& a, void* stream) {
    const int n_sm = sm_count();
    uint32_t* blk_tot = a.counters + CNT_ORDER;
    uint32_t per = (a.nchunks + n_sm - 1) / n_sm;
    per = (per + ORD_THREADS - 1) / ORD_THREADS * ORD_THREADS;
    if (!per) per = ORD_THREADS;
    const uint32_t grid = a.nchunks ? (a.nchunks + per - 1) / per : 1;
    k_order<<<grid, ORD_THREADS, 0, S(stream)>>>(a, per, blk_tot); count_launch();
}
void launch_crc(const PassArgs& a, uint32_t nb, void* stream) {
    k_crc<<<blocks_for(nb, 256), 256, 0, S(stream)>>>(a); count_launch();
}
void launch_link(const PassArgs& a, uint32_t nb, void* stream) {
    k_link<<<blocks_for(nb, 256), 256, 0, S(stream)>>>(a); count_launch();
    k_cover<<<blocks_for(nb, 256), 256, 0, S(stream)>>>(a); count_launch();
}
// fewer than two warps of frame lanes per scheduler: the walk is latency-bound
static bool few_frames(uint32_t nb) { return nb < (uint32_t)sm_count() * 4u * 2u * 32u; }
bool parse_wants_speculation(uint32_t nb, uint32_t channels) {
    const char* force = getenv("BNFLAC_PARSE_SPEC");            // 0 / 1 forces it off / on (the tests run both on the same streams)
    if (force) return force[0] == '1' && channels >= 3;
    // (guessing does not reduce the work, it spreads it: with more than two warps of frame lanes per scheduler the serial walk
    // already fills the machine -- 1 h of 8-channel 192 kHz audio, 42,000 frames: 6.8 ms serial, 7.9 ms guessed)
    return channels >= 3 && few_frames(nb);
}
void launch_parse(const PassArgs& a, uint32_t nb, void* stream) {
    // few frames: branches and divergent loops are what a lone warp waits for (BNFLAC_PARSE_LEAN=0/1 forces one variant)
    const char* force = getenv("BNFLAC_PARSE_LEAN");
    const bool lean = force ? force[0] == '1' : few_frames(nb);
    const size_t smem = PARSE_THREADS * ParseBits::STRIDE;
    if (a.spec_jobs) {
        // guessed subframe starts, all walked at once; what chains up is final, the rest falls through to the serial walk below
        k_spec_find<<<nb ? nb : 1, SPEC_THREADS, 0, S(stream)>>>(a); count_launch();
        const uint32_t jb = blocks_for(a.spec_cap, PARSE_THREADS);
        if (lean) k_parse<true, true><<<jb, PARSE_THREADS, smem, S(stream)>>>(a); else k_parse<false, true><<<jb, PARSE_THREADS, smem, S(stream)>>>(a);
        count_launch();
        k_spec_resolve<<<blocks_for(nb, 128), 128, 0, S(stream)>>>(a); count_launch();
    }
    if (lean) k_parse<true, false><<<blocks_for(nb, PARSE_THREADS), PARSE_THREADS, smem, S(stream)>>>(a);
    else k_parse<false, false><<<blocks_for(nb, PARSE_THREADS), PARSE_THREADS, smem, S(stream)>>>(a);
    count_launch();
}
void launch_resync(const PassArgs& a, void* stream) { k_resync<<<1, 256, 0, S(stream)>>>(a); count_launch(); }
void launch_prefix(const PassArgs& a, uint32_t ncand_bound, uint32_t bytes_per_sample, void* stream) {
    const int n_sm = sm_count();
    uint32_t per = (ncand_bound + n_sm - 1) / n_sm;
    per = (per + ORD_THREADS - 1) / ORD_THREADS * ORD_THREADS;
    if (!per) per = ORD_THREADS;
    const uint32_t grid = ncand_bound ? (ncand_bound + per - 1) / per : 1;
    k_prefix<<<grid, ORD_THREADS, 0, S(stream)>>>(a, bytes_per_sample, per, a.counters + CNT_PFX_CNT, reinterpret_cast<unsigned long long*>(a.counters + CNT_PFX_BYTES)); count_launch();
}
void launch_bucket(const PassArgs& a, uint32_t nacc_bound, void* stream) {
    const uint32_t nb = blocks_for(nacc_bound, BUCKET_CHUNK);
    k_bucket_hist<<<nb, 256, 0, S(stream)>>>(a); count_launch();
    k_bucket_scan<<<1, 32, 0, S(stream)>>>(a, nb); count_launch();
    k_bucket_scatter<<<nb, 256, 0, S(stream)>>>(a); count_launch();
}
void launch_make_chunks(const SegInfo& seg, SegInfo* d_seg, Chunk* chunks, uint32_t nchunks, void* stream) {
    k_make_chunks<<<blocks_for(std::max<uint32_t>(nchunks, 1u), 256), 256, 0, S(stream)>>>(seg, d_seg, chunks, nchunks); count_launch();
}
void launch_clear(const PassArgs& a, void* stream) { k_clear<<<1, 256, 0, S(stream)>>>(a.counters, a.totals); count_launch(); }
void launch_publish(const void* src, void* dst_mapped, uint32_t nwords, void* stream) {
    k_publish<<<1, 32, 0, S(stream)>>>((const uint32_t*)src, (uint32_t*)dst_mapped, nwords); count_launch();
}
void launch_seg_summary(const PassArgs& a, uint32_t nb, uint64_t* seg_pcm, uint32_t* seg_flags, void* stream) {
    k_seg_summary<<<blocks_for(nb, 256), 256, 0, S(stream)>>>(a, seg_pcm, seg_flags); count_launch();
}

} // namespace bnf
