// kernels.cu -- sm_100a kernels of the bnflac decode pipeline (no tensor cores: nothing here is a contraction).
//
//   K1  k_scan      frame-sync scan (0xFFF8/0xFFF9) + header syntax + CRC-8  -> per-chunk sorted candidate lists
//       k_chunk_scan / k_gather   order the per-chunk lists into one frame table
//       k_crc       CRC-16 of every inter-candidate span (warp per span, GF(2) combine of lane pieces)
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
#include "bnflac_dev.h"
#include <cuda_runtime.h>

namespace bnf {

static int g_launches = 0;
int kernel_launch_count() { return g_launches; }

#define FULL 0xffffffffu

// ------------------------------------------------------------------------------------------------ CRC helpers
__device__ __forceinline__ uint32_t crc8_update(uint32_t c, uint32_t byte) {
    c ^= byte;
#pragma unroll
    for (int k = 0; k < 8; k++) c = (c & 0x80) ? ((c << 1) ^ 0x07) & 0xFF : (c << 1) & 0xFF;
    return c;
}
__device__ __forceinline__ uint32_t crc16_update_bitwise(uint32_t c, uint32_t byte) {
    c ^= byte << 8;
#pragma unroll
    for (int k = 0; k < 8; k++) c = (c & 0x8000) ? ((c << 1) ^ 0x8005) & 0xFFFF : (c << 1) & 0xFFFF;
    return c;
}
// product of two residues modulo x^16 + x^15 + x^2 + 1
__device__ __forceinline__ uint32_t gf_mul(uint32_t a, uint32_t b) {
    uint32_t r = 0;
#pragma unroll
    for (int i = 15; i >= 0; --i) {
        r <<= 1;
        if (r & 0x10000u) r ^= 0x18005u;
        if ((b >> i) & 1u) r ^= a;
    }
    return r & 0xFFFFu;
}
// x^(8*nbytes) mod P : appending nbytes bytes to a message multiplies its CRC by this
__device__ uint32_t gf_xpow8(uint32_t nbytes) {
    uint32_t result = 1, base = 0x100;
    while (nbytes) {
        if (nbytes & 1u) result = gf_mul(result, base);
        base = gf_mul(base, base);
        nbytes >>= 1;
    }
    return result;
}

// ------------------------------------------------------------------------------------------------ frame header
struct Hdr {
    uint64_t number;
    uint32_t bs, sample_rate;
    uint32_t hdr_len, bps, assign, variable, channels;
};

// Frame header syntax + CRC-8.  p must have 17 readable bytes.
__device__ bool parse_header(const uint8_t* __restrict__ p, const SegInfo& s, Hdr& h) {
    if (p[0] != 0xFF || (p[1] & 0xFE) != 0xF8) return false;
    uint32_t b2 = p[2], b3 = p[3];
    uint32_t bsc = b2 >> 4, src = b2 & 15, ca = b3 >> 4, ssc = (b3 >> 1) & 7;
    if ((b3 & 1) || bsc == 0 || src == 15 || ca > 10 || ssc == 3 || ssc == 7) return false;
    h.variable = p[1] & 1;
    uint32_t q = 4;
    uint32_t x = p[q++];
    uint64_t num;
    if (x < 0x80) num = x;
    else {
        int n = __clz(~(x << 24));   // leading ones
        if (n == 1 || n > 7 || (!h.variable && n == 7)) return false;
        num = (n == 7) ? 0 : (x & ((1u << (7 - n)) - 1));
        for (int i = 1; i < n; i++) {
            uint32_t y = p[q++];
            if ((y >> 6) != 2) return false;
            num = (num << 6) | (y & 0x3f);
        }
    }
    h.number = num;
    uint32_t bs;
    if (bsc == 1) bs = 192;
    else if (bsc <= 5) bs = 576u << (bsc - 2);
    else if (bsc == 6) { bs = p[q] + 1u; q += 1; }
    else if (bsc == 7) { bs = ((uint32_t)p[q] << 8 | p[q + 1]) + 1u; q += 2; }
    else bs = 256u << (bsc - 8);
    uint32_t sr;
    switch (src) {
    case 0: sr = s.sample_rate; break;
    case 1: sr = 88200; break; case 2: sr = 176400; break; case 3: sr = 192000; break; case 4: sr = 8000; break;
    case 5: sr = 16000; break; case 6: sr = 22050; break; case 7: sr = 24000; break; case 8: sr = 32000; break;
    case 9: sr = 44100; break; case 10: sr = 48000; break; case 11: sr = 96000; break;
    case 12: sr = p[q] * 1000u; q += 1; break;
    case 13: sr = (uint32_t)p[q] << 8 | p[q + 1]; q += 2; break;
    default: sr = ((uint32_t)p[q] << 8 | p[q + 1]) * 10u; q += 2; break;
    }
    uint32_t c = 0;
    for (uint32_t i = 0; i < q; i++) c = crc8_update(c, p[i]);
    if (c != p[q]) return false;
    q++;
    h.bs = bs; h.sample_rate = sr; h.hdr_len = q; h.assign = ca;
    h.channels = ca < 8 ? ca + 1 : 2;
    static const uint8_t sst[8] = {0, 8, 12, 0, 16, 20, 24, 0};
    h.bps = ssc == 0 ? s.bps : sst[ssc];
    // the engine decodes streams whose frames agree with STREAMINFO (every real encoder; per-frame channel/bps
    // changes are treated as false syncs)
    if (h.channels != s.channels || h.bps != s.bps) return false;
    return true;
}

// ------------------------------------------------------------------------------------------------ K1a scan
__device__ __forceinline__ uint32_t bytemask4(uint32_t eq) {   // 0xFF/0x00 per byte -> 4-bit mask
    uint32_t x = (eq & 0x80808080u) >> 7;
    return (x | (x >> 7) | (x >> 14) | (x >> 21)) & 0xFu;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan(PassArgs a) {
    __shared__ uint32_t s_list[SCAN_SCAP];
    __shared__ uint32_t s_sorted[SCAN_SCAP];
    __shared__ uint32_t s_n, s_base;
    const int tid = threadIdx.x, lane = tid & 31;
    for (uint32_t chunk = blockIdx.x; chunk < a.nchunks; chunk += gridDim.x) {
        const Chunk c = a.chunks[chunk];
        const SegInfo seg = a.segs[c.seg];
        if (tid == 0) s_n = 0;
        __syncthreads();
        const uint64_t abeg = c.begin & ~15ull;
        const uint32_t nunits = (uint32_t)((c.begin + c.len - abeg + 15) >> 4);
        const uint32_t nloop = (nunits + 31) & ~31u;
        for (uint32_t u = tid; u < nloop; u += SCAN_THREADS) {
            const bool valid = u < nunits;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (valid) v = __ldg(reinterpret_cast<const uint4*>(a.in + abeg) + u);
            uint32_t nextb = __shfl_down_sync(FULL, v.x & 0xFFu, 1);
            if (valid && (lane == 31 || u + 1 >= nunits)) nextb = a.in[abeg + 16ull * (u + 1)];   // padded input: always readable
            uint32_t ff = bytemask4(__vcmpeq4(v.x, 0xFFFFFFFFu)) | bytemask4(__vcmpeq4(v.y, 0xFFFFFFFFu)) << 4 |
                          bytemask4(__vcmpeq4(v.z, 0xFFFFFFFFu)) << 8 | bytemask4(__vcmpeq4(v.w, 0xFFFFFFFFu)) << 12;
            if (ff) {
                const uint32_t m = 0xFEFEFEFEu, k = 0xF8F8F8F8u;
                uint32_t f8 = bytemask4(__vcmpeq4(v.x & m, k)) | bytemask4(__vcmpeq4(v.y & m, k)) << 4 |
                              bytemask4(__vcmpeq4(v.z & m, k)) << 8 | bytemask4(__vcmpeq4(v.w & m, k)) << 12;
                if ((nextb & 0xFE) == 0xF8) f8 |= 1u << 16;
                uint32_t hits = ff & (f8 >> 1);
                while (hits) {
                    int p = __ffs(hits) - 1;
                    hits &= hits - 1;
                    uint64_t o = abeg + 16ull * u + p;
                    if (o < c.begin || o >= c.begin + c.len) continue;
                    Hdr h;
                    if (!parse_header(a.in + o, seg, h)) continue;
                    if (o + h.hdr_len + 2 > seg.end) continue;
                    uint32_t slot = atomicAdd(&s_n, 1u);
                    if (slot < SCAN_SCAP) s_list[slot] = (uint32_t)(o - c.begin);
                }
            }
        }
        __syncthreads();
        uint32_t n = s_n;
        if (n > SCAN_SCAP) { n = SCAN_SCAP; if (tid == 0) atomicOr(&a.counters[1], 1u); }
        // order the chunk's list by offset (n is ~chunk/frame size; rank sort)
        for (uint32_t e = tid; e < n; e += SCAN_THREADS) {
            uint32_t key = s_list[e], rank = 0;
            for (uint32_t j = 0; j < n; j++) rank += (s_list[j] < key);
            s_sorted[rank] = key;
        }
        if (tid == 0) {
            uint32_t base = atomicAdd(&a.counters[0], n);
            if (base + n > a.cand_cap) { atomicOr(&a.counters[1], 2u); n = base < a.cand_cap ? a.cand_cap - base : 0; }
            s_base = base; s_n = n;
            a.chunk_base[chunk] = base;
            a.chunk_count[chunk] = n;
        }
        __syncthreads();
        n = s_n;
        for (uint32_t e = tid; e < n; e += SCAN_THREADS) {
            uint64_t o = c.begin + s_sorted[e];
            Hdr h;
            parse_header(a.in + o, seg, h);
            Cand cd;
            cd.off = o; cd.number = h.number; cd.bs = h.bs; cd.seg = c.seg;
            cd.hdr_len = (uint8_t)h.hdr_len; cd.bps = (uint8_t)h.bps; cd.assign = (uint8_t)h.assign;
            cd.flags = (uint8_t)(h.variable | ((o < seg.own_begin || o >= seg.own_end) ? 2u : 0u));
            cd.sample_rate = h.sample_rate;
            a.cand_tmp[s_base + e] = cd;
        }
        __syncthreads();
    }
}

// single-CTA exclusive scan of chunk_count -> chunk_scan
__global__ void __launch_bounds__(1024) k_chunk_scan(PassArgs a) {
    __shared__ uint32_t s_part[1024];
    const uint32_t tid = threadIdx.x, n = a.nchunks;
    const uint32_t per = (n + 1023) / 1024;
    uint32_t b = tid * per, e = min(b + per, n), sum = 0;
    for (uint32_t i = b; i < e; i++) sum += a.chunk_count[i];
    s_part[tid] = sum;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {
        uint32_t v = tid >= d ? s_part[tid - d] : 0;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    uint32_t run = s_part[tid] - sum;
    for (uint32_t i = b; i < e; i++) { a.chunk_scan[i] = run; run += a.chunk_count[i]; }
}

// move each chunk's (already sorted) list to its place in the global frame table
__global__ void __launch_bounds__(256) k_gather(PassArgs a) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t chunk = warp; chunk < a.nchunks; chunk += nwarps) {
        uint32_t n = a.chunk_count[chunk], src = a.chunk_base[chunk], dst = a.chunk_scan[chunk];
        for (uint32_t r = lane; r < n; r += 32) a.cand[dst + r] = a.cand_tmp[src + r];
    }
}

// ------------------------------------------------------------------------------------------------ K1c CRC-16 of spans
__device__ __forceinline__ uint32_t ncand(const PassArgs& a) { return min(a.counters[0], a.cand_cap); }

__device__ __forceinline__ uint64_t span_end(const PassArgs& a, uint32_t i, uint32_t n, const Cand& ci) {
    if (i + 1 < n) { const Cand& cn = a.cand[i + 1]; if (cn.seg == ci.seg) return cn.off; }
    return a.segs[ci.seg].end;
}

__global__ void __launch_bounds__(256) k_crc(PassArgs a) {
    __shared__ uint16_t s_tab[256];
    {
        uint32_t d = (uint32_t)threadIdx.x << 8;
#pragma unroll
        for (int k = 0; k < 8; k++) d = (d & 0x8000) ? ((d << 1) ^ 0x8005) & 0xFFFF : (d << 1) & 0xFFFF;
        s_tab[threadIdx.x] = (uint16_t)d;
    }
    __syncthreads();
    const uint32_t n = ncand(a);
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = warp; i < n; i += nwarps) {
        const Cand ci = a.cand[i];
        const uint64_t e = span_end(a, i, n, ci);
        const int64_t L = (int64_t)(e - ci.off);
        const int64_t m = (L + 31) >> 5;
        // zero bytes in front of a message do not change its CRC: lane pieces are laid out from the END so that all
        // (virtual) pieces have the same length m and one factor x^(8m) serves the whole combine tree
        int64_t pb = L - (int64_t)(32 - lane) * m, pe = pb + m;
        if (pb < 0) pb = 0;
        uint32_t crc = 0;
        const uint8_t* p = a.in + ci.off;
        for (int64_t b = pb; b < pe; b++) crc = ((crc << 8) ^ s_tab[(crc >> 8) ^ p[b]]) & 0xFFFF;
        uint32_t f = gf_xpow8((uint32_t)m);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o = __shfl_down_sync(FULL, crc, d);
            crc = gf_mul(crc, f) ^ o;
            f = gf_mul(f, f);
        }
        if (lane == 0) a.seg_crc[i] = (uint16_t)crc;
    }
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
        crc = gf_mul(crc, gf_xpow8((uint32_t)(e2 - e))) ^ a.seg_crc[j];
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

// ------------------------------------------------------------------------------------------------ bit reader
struct BitReader {
    const uint32_t* wp;   // next 32-bit word to load
    const uint32_t* wend; // first word past the (padded) input: loads beyond it read as zero
    uint64_t buf;         // MSB-aligned window, at least 32 valid bits after every operation
    int cnt;              // valid bits in buf

    __device__ __forceinline__ uint32_t load() {
        uint32_t w = wp < wend ? __ldg(wp) : 0u;
        wp++;
        return __byte_perm(w, 0, 0x0123);
    }
    __device__ __forceinline__ void set_limit(const uint8_t* base, uint64_t in_len) {
        wend = reinterpret_cast<const uint32_t*>(base) + ((in_len + 3) >> 2);
    }
    __device__ __forceinline__ void init(const uint8_t* base, uint64_t bitpos) {
        wp = reinterpret_cast<const uint32_t*>(base) + (bitpos >> 5);
        uint32_t w0 = load(), w1 = load();
        buf = ((uint64_t)w0 << 32) | w1;
        cnt = 64;
        int s = (int)(bitpos & 31);
        if (s) skip(s);
    }
    __device__ __forceinline__ void skip(int n) {   // 0 <= n <= 32
        buf <<= n;
        cnt -= n;
        if (cnt < 32) { buf |= (uint64_t)load() << (32 - cnt); cnt += 32; }
    }
    __device__ __forceinline__ uint32_t peek32() const { return (uint32_t)(buf >> 32); }
    __device__ __forceinline__ uint32_t get(int n) {   // 0 <= n <= 32
        uint32_t v = n ? (uint32_t)(buf >> (64 - n)) : 0u;
        skip(n);
        return v;
    }
    __device__ __forceinline__ int32_t gets(int n) {
        int32_t v = n ? (int32_t)((int64_t)buf >> (64 - n)) : 0;
        skip(n);
        return v;
    }
    __device__ __forceinline__ uint64_t pos(const uint8_t* base) const {
        return (uint64_t)(wp - reinterpret_cast<const uint32_t*>(base)) * 32 - (uint64_t)cnt;
    }
    // zeros before the terminating 1; `limit` bounds a runaway on damaged data
    __device__ __forceinline__ uint32_t unary(uint32_t limit) {
        uint32_t q = 0;
        for (;;) {
            uint32_t top = peek32();
            if (top) { int z = __clz(top); skip(z + 1); return q + (uint32_t)z; }
            q += 32; skip(32);
            if (q > limit) return q;
        }
    }
};

// ------------------------------------------------------------------------------------------------ ring bit reader
// Every lane walks its own serial bitstream.  The bytes are staged through shared memory by per-lane cp.async
// (LDGSTS, 16 B each) into a private ring that runs several blocks ahead of the read position, so the serial parse
// never waits on HBM.  Reads are position based: two LDS + byte swaps + one funnel shift give a 32-bit window.
// Refill is CHECKPOINTED: all lanes of a warp top up their rings at the same loop iterations (every CK samples), so the
// divergent "my ring ran low" branch is not taken on almost every iteration by some lane.  Between checkpoints a lane
// may consume at most CK*32 bits + one window, which the ring always holds ahead (see checkpoint()).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t shr_c(uint32_t v, uint32_t n) { uint32_t r; asm("shr.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n)); return r; }   // n >= 32 -> 0
__device__ __forceinline__ uint32_t shl_c(uint32_t v, uint32_t n) { uint32_t r; asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n)); return r; }
__device__ __forceinline__ int32_t sar_c(int32_t v, uint32_t n) { int32_t r; asm("shr.s32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n)); return r; }

template <int NBLK>                              // ring = NBLK blocks of 16 bytes per lane (NBLK a power of two)
struct RingBits {
    static constexpr int BLK = 16;
    static constexpr int RB_BYTES = BLK * NBLK;
    static constexpr int STRIDE = RB_BYTES + 16;           // lane stride: 16 B skew spreads lanes over banks, keeps 16 B alignment
    // bytes guaranteed readable ahead of pos right after checkpoint(): blocks up to (curblk_prev + NBLK - 1) are complete
    uint32_t sring;        // shared-space address of this lane's ring
    uint32_t pos;          // bit position relative to g0
    uint32_t filled;       // blocks [.., filled) have been requested
    const uint8_t* g0;     // global address of ring byte 0 (16 B aligned)
    const uint8_t* gend;   // end of the padded input: blocks past it are zero-filled

    __device__ __forceinline__ void fetch(uint32_t b) {
        const uint8_t* s = g0 + (uint64_t)b * BLK;
        uint32_t n = (s + 16 <= gend) ? 16u : 0u;
        if (!n) s = gend - 16;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sring + (b & (NBLK - 1)) * BLK), "l"(s), "r"(n) : "memory");
    }
    __device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
    __device__ __forceinline__ void wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
    __device__ __forceinline__ void request() {     // request every block the ring has room for (never the slot being read)
        const uint32_t curblk = pos / (BLK * 8);
        if (filled < curblk) filled = curblk;       // jumped over unrequested blocks
        const uint32_t lim = curblk + NBLK;
        if (filled < lim) {
#pragma unroll 1
            do { fetch(filled); filled++; } while (filled < lim);
            commit();
        }
    }
    // Periodic top-up, executed by all lanes at the same iteration: first make what was requested one period ago
    // visible, then request more.  Requests therefore have a whole period to land.
    __device__ __forceinline__ void checkpoint() { wait_all(); request(); }
    // Synchronous variant for rare big moves (VERBATIM skip, long unary runs, init).
    __device__ __forceinline__ void ensure_now() { request(); wait_all(); }
    __device__ __forceinline__ void init(uint32_t sring_, const uint8_t* in, uint64_t in_len, uint64_t abs_bit) {
        sring = sring_;
        const uint64_t b0 = (abs_bit >> 3) & ~(uint64_t)(BLK - 1);
        g0 = in + b0; gend = in + (in_len & ~15ull);
        pos = (uint32_t)(abs_bit - b0 * 8);
        filled = 0;
        ensure_now();
    }
    __device__ __forceinline__ uint64_t abs_pos(const uint8_t* in) const { return (uint64_t)(g0 - in) * 8 + pos; }
    __device__ __forceinline__ uint32_t window() const {      // next 32 bits, MSB first
        const uint32_t bo = (pos >> 3) & (RB_BYTES - 4);
        const uint32_t a = lds32(sring + bo), b = lds32(sring + ((bo + 4) & (RB_BYTES - 4)));
        return __funnelshift_l(__byte_perm(b, 0, 0x0123), __byte_perm(a, 0, 0x0123), pos & 31);
    }
    __device__ __forceinline__ void skip(uint32_t n) { pos += n; }                        // n <= 32, covered by the checkpoint budget
    __device__ __forceinline__ void jump(uint32_t n) { pos += n; ensure_now(); }          // any n
    __device__ __forceinline__ uint32_t get(uint32_t n) { uint32_t v = shr_c(window(), 32 - n); pos += n; return v; }          // n <= 32
    __device__ __forceinline__ int32_t gets(uint32_t n) { int32_t v = n ? sar_c((int32_t)window(), 32 - n) : 0; pos += n; return v; }
    // unary run that did not terminate inside one window (rare): walks 32 zero bits at a time with synchronous refills
    __device__ __forceinline__ uint32_t unary_slow(uint32_t limit) {
        uint32_t q = 0;
#pragma unroll 1
        for (;;) {
            ensure_now();
            uint32_t w = window();
            if (w) { uint32_t z = __clz(w); pos += z + 1; ensure_now(); return q + z; }
            q += 32; pos += 32;
            if (q > limit) { ensure_now(); return q; }
        }
    }
    __device__ __forceinline__ uint32_t unary(uint32_t limit) {
        uint32_t w = window();
        if (w) { uint32_t z = __clz(w); pos += z + 1; return z; }
        return unary_slow(limit);
    }
};
// budget: between two checkpoints a lane may advance by at most (NBLK*16 - 16 - 8 - 16) bytes:
//   ring - (partially consumed current block) - (window over-read) - (blocks requested this checkpoint are not yet waited for)
// NBLK = 8: 128 B ring, blocks requested one checkpoint ago are complete -> >= 64 B ahead; CK = 8 samples * 4 B + slack fits.

// ------------------------------------------------------------------------------------------------ K2 parse
// One thread per frame: walks the subframes, records where each starts and what it is, skips the residual.
__device__ __forceinline__ int ilog2u(uint32_t v) { return 31 - __clz(v); }
constexpr int PARSE_THREADS = 64;
using ParseBits = RingBits<16>;   // 256 B ring: a Rice codeword walked here may be up to 62 bits

__global__ void __launch_bounds__(PARSE_THREADS) k_parse(PassArgs a) {
    extern __shared__ __align__(16) uint8_t s_ring[];
    const uint32_t n = ncand(a);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint8_t st = ST_DROP;
    Cand c;
    c.bs = 0; c.assign = 0; c.flags = 0; c.off = 0; c.hdr_len = 0; c.bps = 0;
    if (i < n) { st = a.status[i]; c = a.cand[i]; }
    bool live = (st == ST_OK || st == ST_CHECK) && !(c.flags & 2);   // flag 2: frame of the neighbouring shard (end marker only)
    const uint32_t channels = c.assign < 8 ? c.assign + 1u : 2u;
    const uint64_t frame_bit0 = c.off * 8;
    const uint64_t end_bit = live ? (c.off + a.flen[i]) * 8 : 0;
    ParseBits br;
    if (live) br.init(smem_u32(s_ring) + threadIdx.x * ParseBits::STRIDE, a.in, a.in_len, frame_bit0 + 8ull * c.hdr_len);
    bool bad = false, unparse = false;
    uint32_t max_order = 0, any_wide = 0;
    // lanes of a warp are different frames; they walk channel by channel and, inside a subframe, sample index by
    // sample index in lockstep, so that partition boundaries (multiples of blocksize >> order) fall on the same iteration
    const uint32_t wmax_ch = __reduce_max_sync(FULL, live ? channels : 0u);
    const uint32_t wmax_bs = __reduce_max_sync(FULL, live ? c.bs : 0u);
    for (uint32_t ch = 0; ch < wmax_ch; ch++) {
        bool walk = false;            // this lane walks a Rice-coded residual in this phase
        uint32_t order = 0, plen = 4, esc = 15, psize = 0, left = 0, k1 = 1, rawskip = 0;
        bool first = true;
        if (live && !bad && !unparse && ch < channels) {
            uint32_t bps = c.bps + (((c.assign == 8 && ch == 1) || (c.assign == 9 && ch == 0) || (c.assign == 10 && ch == 1)) ? 1u : 0u);
            SubInfo si;
            si.bit_offset = (uint32_t)(br.abs_pos(a.in) - frame_bit0);
            si.type = 0; si.order = 0; si.flags = 0;
            uint32_t x = br.get(8);
            uint32_t type = (x >> 1) & 0x3f, w = 0;
            if (x & 0x80) bad = true;
            else if ((x & 1) && (w = br.unary(64) + 1) >= bps) unparse = true;
            else {
                bps -= w;
                si.wasted = (uint8_t)w;
                bool has_resid = false;
                if (type == 0) br.skip(bps);
                else if (type == 1) {
                    si.type = 1;
                    if (br.abs_pos(a.in) + (uint64_t)c.bs * bps > end_bit) bad = true; else br.jump(c.bs * bps);
                } else if (type >= 8 && type <= 12) { si.type = 2; order = type - 8; has_resid = true; }
                else if (type >= 32) { si.type = 3; order = type - 31; has_resid = true; }
                else unparse = true;
                if (has_resid) {
                    si.order = (uint8_t)order;
                    if (order > c.bs) unparse = true;
                    else if (br.abs_pos(a.in) + (uint64_t)order * bps > end_bit) bad = true;
                    else {
                        if (order > max_order) max_order = order;
                        br.jump(order * bps);
                        if (si.type == 3) {
                            uint32_t prec = br.get(4) + 1;
                            int32_t shift = br.gets(5);
                            if (prec == 16 || shift < 0) unparse = true;
                            else {
                                br.jump(order * prec);
                                if (bps + prec + (uint32_t)ilog2u(order) <= 32) si.flags |= 1; else any_wide = 1;
                            }
                        } else si.flags |= 1;
                        if (!unparse) {
                            uint32_t method = br.get(2);
                            if (method > 1) unparse = true;
                            else {
                                if (method) { si.flags |= 2; plen = 5; esc = 31; }
                                uint32_t po = br.get(4);
                                psize = po ? c.bs >> po : c.bs;
                                if (psize < order) unparse = true;
                                // the residual of the LAST subframe need not be walked when the frame span is already
                                // CRC-validated: nothing starts after it
                                else walk = !(st == ST_OK && ch + 1 == channels);
                            }
                        }
                    }
                }
            }
            if (!bad && !unparse) a.sub[(uint64_t)i * MAX_CH + ch] = si;
            else walk = false;
        }
        if (!__any_sync(FULL, walk)) continue;
        for (uint32_t s = 0; s < wmax_bs; s++) {
            if ((s & 7) == 0 && walk) br.checkpoint();
            if (walk && s >= order && s < c.bs) {
                if (left == 0) {   // partition boundary (s is a multiple of psize, or s == order)
                    do {               // twice only when partition 0 holds zero samples (its parameter is still coded)
                        left = psize - (s == order && first ? order : 0);
                        first = false;
                        uint32_t k = br.get(plen);
                        rawskip = 0;
                        if (k == esc) {
                            uint32_t nb = br.get(5);
                            if (br.abs_pos(a.in) + (uint64_t)left * nb > end_bit) { bad = true; walk = false; left = 1; }
                            else { br.jump(left * nb); rawskip = 1; }
                        }
                        k1 = k + 1;
                    } while (left == 0);
                }
                left--;
                if (!rawskip && walk) {
                    uint32_t wd = br.window();
                    if (wd) br.skip(__clz(wd) + k1);
                    else {
                        br.unary_slow(1u << 16);
                        br.jump(k1 - 1);
                        if (br.abs_pos(a.in) > end_bit) { bad = true; walk = false; }
                    }
                }
            }
        }
        if (walk && br.abs_pos(a.in) > end_bit) bad = true;
    }
    if (!live) return;
    if (!bad && !unparse && st == ST_CHECK) {
        uint64_t p = (br.abs_pos(a.in) + 7) & ~7ull;
        // validate the way the reference does: CRC-16 over the bytes the parse consumed
        if (p + 16 > end_bit) bad = true;
        else {
            const uint8_t* q = a.in + c.off;
            uint32_t nbytes = (uint32_t)(p / 8 - c.off), crc = 0;
            for (uint32_t b = 0; b < nbytes; b++) crc = crc16_update_bitwise(crc, q[b]);
            uint32_t want = (uint32_t)q[nbytes] << 8 | q[nbytes + 1];
            a.flen[i] = nbytes + 2;
            st = (crc == want) ? ST_OK : ST_CRC;
        }
    }
    if (unparse) st = ST_UNPARSEABLE;
    else if (bad) st = (st == ST_CHECK) ? ST_DROP : ST_CRC;
    a.status[i] = st;
    if (st == ST_OK) {
        if (max_order) atomicMax(&a.totals->max_order, max_order);
        if (any_wide) atomicOr(&a.totals->any_wide, 1u);
    }
}

// ------------------------------------------------------------------------------------------------ prefix
__global__ void __launch_bounds__(1024) k_prefix(PassArgs a, uint32_t bytes_per_sample) {
    __shared__ uint64_t s_bytes[1024];
    __shared__ uint32_t s_cnt[1024];
    __shared__ uint32_t s_maxbs[32];
    const uint32_t tid = threadIdx.x, n = ncand(a);
    const uint32_t per = (n + 1023) / 1024;
    const uint32_t b = min(tid * per, n), e = min(b + per, n);
    uint64_t bytes = 0; uint32_t cnt = 0, maxbs = 0;
    for (uint32_t i = b; i < e; i++) {
        uint8_t st = a.status[i];
        const Cand& c = a.cand[i];
        if ((st == ST_OK || st == ST_CRC) && !(c.flags & 2)) {
            uint32_t ch = c.assign < 8 ? c.assign + 1u : 2u;
            bytes += (uint64_t)c.bs * ch * bytes_per_sample; cnt++;
            maxbs = max(maxbs, c.bs);
        }
    }
    s_bytes[tid] = bytes; s_cnt[tid] = cnt;
    maxbs = __reduce_max_sync(FULL, maxbs);
    if ((tid & 31) == 0) s_maxbs[tid >> 5] = maxbs;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {
        uint64_t vb = tid >= d ? s_bytes[tid - d] : 0; uint32_t vc = tid >= d ? s_cnt[tid - d] : 0;
        __syncthreads();
        s_bytes[tid] += vb; s_cnt[tid] += vc;
        __syncthreads();
    }
    uint64_t rb = s_bytes[tid] - bytes; uint32_t rc = s_cnt[tid] - cnt;
    for (uint32_t i = b; i < e; i++) {
        uint8_t st = a.status[i];
        const Cand& c = a.cand[i];
        if ((st == ST_OK || st == ST_CRC) && !(c.flags & 2)) {
            uint32_t ch = c.assign < 8 ? c.assign + 1u : 2u;
            a.pcm_off[i] = rb; a.acc_idx[rc] = i;
            rb += (uint64_t)c.bs * ch * bytes_per_sample; rc++;
        } else a.pcm_off[i] = rb;
    }
    if (tid == 1023) {
        a.totals->pcm_bytes = s_bytes[1023]; a.totals->n_accepted = s_cnt[1023]; a.totals->n_cand = n;
        a.totals->overflow = a.counters[1];
    }
    if (tid == 0) { uint32_t m = 0; for (int w = 0; w < 32; w++) m = max(m, s_maxbs[w]); a.totals->max_bs = m; }
}

// ------------------------------------------------------------------------------------------------ K3-5 decode
// One thread per (frame, channel).  Threads of a CTA advance sample index by sample index in lockstep; each keeps its
// bit position, Rice state, quantised LPC coefficients and the last ORD samples in registers (the history is a
// register ring: the sample loop is unrolled ORD times so every tap has a fixed register).  Stereo decorrelation is a
// warp shuffle between the two channel lanes; samples are staged interleaved in shared memory and written out packed.
constexpr int DEC_THREADS = 128;
using DecBits = RingBits<8>;                     // 128 B ring per lane

template <int ORD> struct DecCfg {
    static constexpr int T = 64;                          // samples per channel staged per tile
    static constexpr int U = (ORD >= 16) ? 8 : 4;         // sample-loop unroll; the history shifts by U registers every U samples
    static constexpr int CK = 8;                          // samples between ring checkpoints
};

enum : int { M_IDLE = 0, M_CONST = 1, M_VERBATIM = 2, M_PRED = 3 };

// bytes [0, nbytes) of the interleaved packed PCM of `ts` (flat int32 samples, B bytes each) -> dst, cooperatively by a warp
__device__ __forceinline__ uint32_t sample_byte(const int32_t* ts, uint32_t b, uint32_t B) {
    uint32_t q = b / B, r = b - q * B;
    return ((uint32_t)ts[q] >> (8 * r)) & 0xFFu;
}
__device__ void store_packed(uint8_t* __restrict__ dst, uint32_t nbytes, const int32_t* __restrict__ ts, uint32_t B, int lane) {
    uint32_t head = (4u - (uint32_t)((uintptr_t)dst & 3u)) & 3u;
    if (head > nbytes) head = nbytes;
    if ((uint32_t)lane < head) dst[lane] = (uint8_t)sample_byte(ts, lane, B);
    const uint32_t nw = (nbytes - head) >> 2;
    uint32_t* __restrict__ dw = reinterpret_cast<uint32_t*>(dst + head);
    for (uint32_t j = lane; j < nw; j += 32) {
        const uint32_t b0 = head + 4 * j;
        uint32_t w;
        if (B == 3) {
            uint32_t q = (b0 * 0xAAABu) >> 17, r = b0 - 3 * q;
            uint32_t s0 = (uint32_t)ts[q], s1 = (uint32_t)ts[q + 1];
            w = r == 0 ? __byte_perm(s0, s1, 0x4210) : r == 1 ? __byte_perm(s0, s1, 0x5421) : __byte_perm(s0, s1, 0x6542);
        } else if (B == 2 && !(b0 & 1)) {
            uint32_t q = b0 >> 1;
            w = __byte_perm((uint32_t)ts[q], (uint32_t)ts[q + 1], 0x5410);
        } else if (B == 1) {
            w = __byte_perm(__byte_perm((uint32_t)ts[b0], (uint32_t)ts[b0 + 1], 0x0040), __byte_perm((uint32_t)ts[b0 + 2], (uint32_t)ts[b0 + 3], 0x0040), 0x5410);
        } else {
            w = sample_byte(ts, b0, B) | sample_byte(ts, b0 + 1, B) << 8 | sample_byte(ts, b0 + 2, B) << 16 | sample_byte(ts, b0 + 3, B) << 24;
        }
        dw[j] = w;
    }
    const uint32_t tb = head + 4 * nw;
    if (tb + lane < nbytes) dst[tb + lane] = (uint8_t)sample_byte(ts, tb + lane, B);
}

template <int ORD, bool WIDE>
__global__ void __launch_bounds__(DEC_THREADS) k_decode(PassArgs a, uint32_t C, uint32_t B) {
    constexpr int T = DecCfg<ORD>::T, U = DecCfg<ORD>::U, CK = DecCfg<ORD>::CK;
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ uint32_t s_bs[DEC_THREADS];
    __shared__ uint64_t s_po[DEC_THREADS];
    __shared__ uint32_t s_maxbs;
    int32_t* s_tile = reinterpret_cast<int32_t*>(s_dyn + DEC_THREADS * DecBits::STRIDE);
    const uint32_t F = DEC_THREADS / C;
    const uint32_t stride = T * C + C;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n_acc = a.totals->n_accepted;
    const uint32_t fl = tid / C, ch = tid - fl * C;
    // idle threads (C does not divide the CTA size) write to a scratch word behind the tiles
    const uint32_t my_base = smem_u32(fl < F ? s_tile + fl * stride + ch : s_tile + F * stride);
    const uint32_t c4 = fl < F ? C * 4 : 0;
    for (uint32_t g = blockIdx.x; (uint64_t)g * F < n_acc; g += gridDim.x) {
        const uint32_t kf = g * F + fl;
        const bool active = fl < F && kf < n_acc;
        if (tid == 0) s_maxbs = 0;
        __syncthreads();
        // ---- per-subframe state (registers)
        DecBits br;
        int32_t coef[ORD], hist[ORD];          // hist[0] = most recent sample (before the wasted-bits shift)
#pragma unroll
        for (int j = 0; j < ORD; j++) { coef[j] = 0; hist[j] = 0; }
        uint32_t bs = 0, assign = 0, order = 0, wasted = 0, shift = 0, bps = 0;
        uint32_t psize = 0, plen = 4, fastleft = 0, rawleft = 0, rawbits = 0, k = 0;
        int32_t cval = 0;
        int mode = M_IDLE;
        bool narrow = true, first_part = true;
        if (active) {
            const uint32_t i = a.acc_idx[kf];
            const Cand c = a.cand[i];
            bs = c.bs; assign = c.assign;
            bool ok = a.status[i] == ST_OK;
            const uint64_t po = a.pcm_off[i];
            if (po + (uint64_t)bs * C * B > a.out_cap) { ok = false; bs = 0; }   // never write past the caller's buffer
            if (ch == 0) { s_bs[fl] = bs; s_po[fl] = po; }
            atomicMax(&s_maxbs, bs);
            if (ok) {
                const SubInfo si = a.sub[(uint64_t)i * MAX_CH + ch];
                br.init(smem_u32(s_dyn) + tid * DecBits::STRIDE, a.in, a.in_len, c.off * 8 + si.bit_offset);
                uint32_t x = br.get(8);
                if (x & 1) br.unary(64);
                order = si.order; wasted = si.wasted;
                bps = (uint32_t)c.bps + (((assign == 8 && ch == 1) || (assign == 9 && ch == 0) || (assign == 10 && ch == 1)) ? 1u : 0u) - wasted;
                if (si.type == 0) { mode = M_CONST; cval = br.gets(bps); }
                else if (si.type == 1) mode = M_VERBATIM;
                else {
                    mode = M_PRED;
                    // warm-up samples are parked in this thread's tile slots 0..order-1 (order <= 32 <= T) and picked up
                    // again, in order, by the sample loop
#pragma unroll 1
                    for (uint32_t j = 0; j < order; j++) { sts32(my_base + j * c4, (uint32_t)br.gets(bps)); if ((j & 7) == 7) br.ensure_now(); }
                    br.ensure_now();
                    if (si.type == 3) {
                        const uint32_t prec = br.get(4) + 1;
                        shift = (uint32_t)br.gets(5);
                        narrow = (bps + prec + (uint32_t)ilog2u(order)) <= 32;
#pragma unroll
                        for (int j = 0; j < ORD; j++) if (j < (int)order) { coef[j] = br.gets(prec); if ((j & 7) == 7) br.ensure_now(); }
                        br.ensure_now();
                    } else {   // FIXED predictors as LPC coefficient sets, 32-bit wrap-around arithmetic
                        const int o = (int)order;
                        if (ORD >= 1 && o >= 1) coef[0] = o == 1 ? 1 : o == 2 ? 2 : o == 3 ? 3 : 4;
                        if (ORD >= 2 && o >= 2) coef[1] = o == 2 ? -1 : o == 3 ? -3 : -6;
                        if (ORD >= 3 && o >= 3) coef[2] = o == 3 ? 1 : 4;
                        if (ORD >= 4 && o >= 4) coef[3] = -1;
                    }
                    const uint32_t method = br.get(2);
                    plen = method ? 5 : 4;
                    const uint32_t porder = br.get(4);
                    psize = porder ? bs >> porder : bs;
                }
            }
        } else if (fl < F && ch == 0) s_bs[fl] = 0;
        __syncthreads();
        const uint32_t maxbs = s_maxbs;
        const bool reads = mode >= M_VERBATIM;
        for (uint32_t i0 = 0; i0 < maxbs; i0 += T) {
#pragma unroll 1
            for (uint32_t t0 = 0; t0 < (uint32_t)T; t0 += U) {
                const uint32_t my_row = my_base + t0 * c4;
                if ((t0 % CK) == 0 && reads) br.checkpoint();
                int32_t nw[U];
#pragma unroll
                for (int j = 0; j < U; j++) {
                    const uint32_t idx = i0 + t0 + j;
                    int32_t s = 0;
                    bool pred = false;
                    int32_t r = 0;
                    if (fastleft) {          // the common case: next Rice codeword of the current partition
                        fastleft--;
                        pred = true;
                        const uint32_t w = br.window();
                        const uint32_t z = __clz(w);
                        if (z + 1 + k <= 32) {
                            const uint32_t u = (z << k) | shr_c(shl_c(w, z + 1), 32 - k);
                            r = (int32_t)(u >> 1) ^ -(int32_t)(u & 1);
                            br.skip(z + 1 + k);
                        } else {                 // codeword longer than one window (rare): resynchronise the ring afterwards
                            const uint32_t q = br.unary(1u << 24);
                            const uint32_t u = (q << k) | br.get(k);
                            r = (int32_t)(u >> 1) ^ -(int32_t)(u & 1);
                            br.ensure_now();
                        }
                    } else if (idx < bs && mode != M_IDLE) {
                        if (mode == M_CONST) s = cval;
                        else if (mode == M_VERBATIM) s = br.gets(bps);
                        else if (idx < order) s = (int32_t)lds32(my_row + (uint32_t)j * c4);      // parked warm-up sample
                        else {
                            pred = true;
                            if (rawleft == 0) {                  // partition boundary
                                uint32_t cnt;
#pragma unroll 1
                                do {                             // twice only if partition 0 holds zero samples
                                    cnt = psize - (first_part ? order : 0);
                                    first_part = false;
                                    k = br.get(plen);
                                    if (k == (plen == 5 ? 31u : 15u)) { rawbits = br.get(5); rawleft = cnt; fastleft = 0; }
                                    else { fastleft = cnt; rawleft = 0; }
                                } while (cnt == 0);
                            }
                            if (rawleft) { rawleft--; r = br.gets(rawbits); }
                            else {
                                fastleft--;
                                const uint32_t q = br.unary(1u << 24);
                                const uint32_t u = (q << k) | br.get(k);
                                r = (int32_t)(u >> 1) ^ -(int32_t)(u & 1);
                                br.ensure_now();
                            }
                        }
                    }
                    if (pred) {
                        if (WIDE) {
                            int64_t sum = 0;
#pragma unroll
                            for (int m = 0; m < ORD; m++) {
                                const int32_t h = (m < j) ? nw[j - 1 - m < 0 ? 0 : j - 1 - m] : hist[m - j < 0 ? 0 : m - j];
                                asm("mad.wide.s32 %0, %1, %2, %0;" : "+l"(sum) : "r"(coef[m]), "r"(h));
                            }
                            if (narrow) sum = (int64_t)(int32_t)sum;      // libFLAC 1.2.1 accumulates in 32 bits here
                            s = (int32_t)((uint32_t)r + (uint32_t)(int32_t)(sum >> shift));
                        } else {
                            uint32_t sum = 0;
#pragma unroll
                            for (int m = 0; m < ORD; m++) {
                                const int32_t h = (m < j) ? nw[j - 1 - m < 0 ? 0 : j - 1 - m] : hist[m - j < 0 ? 0 : m - j];
                                sum += (uint32_t)coef[m] * (uint32_t)h;
                            }
                            s = (int32_t)((uint32_t)r + (uint32_t)((int32_t)sum >> shift));
                        }
                    }
                    nw[j] = s;
                    int32_t v = (int32_t)((uint32_t)s << wasted);
                    if (C == 2) {
                        const int32_t o = __shfl_xor_sync(FULL, v, 1);
                        if (assign == 8) { if (ch == 1) v = (int32_t)((uint32_t)o - (uint32_t)v); }
                        else if (assign == 9) { if (ch == 0) v = (int32_t)((uint32_t)v + (uint32_t)o); }
                        else if (assign == 10) {
                            const int32_t mid = ch == 0 ? v : o, side = ch == 0 ? o : v;
                            const uint32_t m2 = ((uint32_t)mid << 1) | ((uint32_t)side & 1u);
                            v = ch == 0 ? ((int32_t)(m2 + (uint32_t)side) >> 1) : ((int32_t)(m2 - (uint32_t)side) >> 1);
                        }
                    }
                    sts32(my_row + (uint32_t)j * c4, (uint32_t)v);
                }
                // slide the history window by U samples
#pragma unroll
                for (int m = ORD - 1; m >= U; m--) hist[m] = hist[m - U];
#pragma unroll
                for (int m = 0; m < U && m < ORD; m++) hist[m] = nw[U - 1 - m];
            }
            __syncthreads();
            for (uint32_t f = warp; f < F; f += DEC_THREADS / 32) {
                const uint32_t fbs = s_bs[f];
                if (i0 >= fbs) continue;
                const uint32_t nt = min((uint32_t)T, fbs - i0);
                store_packed(a.out + s_po[f] + (uint64_t)i0 * C * B, nt * C * B, s_tile + f * stride, B, lane);
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------ launchers
static inline cudaStream_t S(void* s) { return (cudaStream_t)s; }
static inline uint32_t blocks_for(uint64_t n, uint32_t per) { uint64_t b = (n + per - 1) / per; return (uint32_t)(b ? b : 1); }

void launch_scan(const PassArgs& a, void* stream) {
    uint32_t grid = a.nchunks < 148u * 8 ? a.nchunks : 148u * 8;
    if (!grid) grid = 1;
    k_scan<<<grid, SCAN_THREADS, 0, S(stream)>>>(a); g_launches++;
}
void launch_order(const PassArgs& a, void* stream) {
    k_chunk_scan<<<1, 1024, 0, S(stream)>>>(a); g_launches++;
    k_gather<<<148 * 2, 256, 0, S(stream)>>>(a); g_launches++;
}
void launch_crc(const PassArgs& a, uint32_t nb, void* stream) {
    uint32_t grid = blocks_for(nb, 8); if (grid > 148 * 8) grid = 148 * 8;
    k_crc<<<grid, 256, 0, S(stream)>>>(a); g_launches++;
}
void launch_link(const PassArgs& a, uint32_t nb, void* stream) {
    k_link<<<blocks_for(nb, 256), 256, 0, S(stream)>>>(a); g_launches++;
    k_cover<<<blocks_for(nb, 256), 256, 0, S(stream)>>>(a); g_launches++;
}
void launch_parse(const PassArgs& a, uint32_t nb, void* stream) {
    k_parse<<<blocks_for(nb, PARSE_THREADS), PARSE_THREADS, PARSE_THREADS * ParseBits::STRIDE, S(stream)>>>(a); g_launches++;
}
void launch_prefix(const PassArgs& a, uint32_t bytes_per_sample, void* stream) {
    k_prefix<<<1, 1024, 0, S(stream)>>>(a, bytes_per_sample); g_launches++;
}
template <int ORD, bool WIDE>
static void launch_decode_t(const PassArgs& a, uint32_t nacc, uint32_t C, uint32_t B, cudaStream_t st) {
    const uint32_t F = DEC_THREADS / C;
    uint32_t grid = blocks_for(nacc, F);
    if (grid > 148 * 32) grid = 148 * 32;
    const size_t smem = (size_t)DEC_THREADS * DecBits::STRIDE + ((size_t)F * (DecCfg<ORD>::T * C + C) + 4) * sizeof(int32_t);
    static bool attr_done = false;
    if (!attr_done) { cudaFuncSetAttribute(k_decode<ORD, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); attr_done = true; }
    k_decode<ORD, WIDE><<<grid, DEC_THREADS, smem, st>>>(a, C, B);
    g_launches++;
}
void launch_decode(const PassArgs& a, uint32_t nacc, uint32_t C, uint32_t B, uint32_t max_order, bool wide, void* stream) {
    cudaStream_t st = S(stream);
    if (wide) {
        if (max_order <= 4) launch_decode_t<4, true>(a, nacc, C, B, st);
        else if (max_order <= 8) launch_decode_t<8, true>(a, nacc, C, B, st);
        else if (max_order <= 12) launch_decode_t<12, true>(a, nacc, C, B, st);
        else if (max_order <= 16) launch_decode_t<16, true>(a, nacc, C, B, st);
        else launch_decode_t<32, true>(a, nacc, C, B, st);
    } else {
        if (max_order <= 4) launch_decode_t<4, false>(a, nacc, C, B, st);
        else if (max_order <= 8) launch_decode_t<8, false>(a, nacc, C, B, st);
        else if (max_order <= 12) launch_decode_t<12, false>(a, nacc, C, B, st);
        else if (max_order <= 16) launch_decode_t<16, false>(a, nacc, C, B, st);
        else launch_decode_t<32, false>(a, nacc, C, B, st);
    }
}

} // namespace bnf
