// kernels_common.cuh -- what the translation units of the device code share: bit / shared-memory helpers and the ring bit
// reader both serial walkers (k_parse in kernels.cu, k_decode in kernels_decode.cuh) are built on.
#pragma once
#include "bnflac_dev.h"
#include <cuda_runtime.h>
#include <type_traits>
#include <cstdlib>
#include <algorithm>
#include <cstdio>
#include <atomic>

namespace bnf {

uint32_t next_decode_epoch();  // kernels.cu: a number no earlier balanced k_decode launch of this process has used (never 0)
void count_launch();          // kernels.cu: one more kernel launched (bench "gpu_launches")

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
// ------------------------------------------------------------------------------------------------ shared-memory / bit helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ uint2 lds64(uint32_t addr) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr)); return v; }
__device__ __forceinline__ uint4 lds128(uint32_t addr) { uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)); return v; }
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) { asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ uint32_t get4(const uint4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
__device__ __forceinline__ void set4(uint4& v, int i, uint32_t x) { if (i == 0) v.x = x; else if (i == 1) v.y = x; else if (i == 2) v.z = x; else v.w = x; }
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t shr_c(uint32_t v, uint32_t n) { uint32_t r; asm("shr.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n)); return r; }   // n >= 32 -> 0
__device__ __forceinline__ uint32_t shl_c(uint32_t v, uint32_t n) { uint32_t r; asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n)); return r; }
__device__ __forceinline__ int32_t sar_c(int32_t v, uint32_t n) { int32_t r; asm("shr.s32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n)); return r; }
__device__ __forceinline__ uint32_t bfind(uint32_t v) { uint32_t r; asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(v)); return r; }   // index of the leading one; 0xffffffff for 0
// index of the leading one for the branch-free codeword groups.  BFIND_I2F = 1: through the integer-to-float converter (round toward zero:
// the exponent is floor(log2 v)); 0 gives -127, negative like bfind's -1, which is all the callers test
#ifndef BFIND_I2F
#define BFIND_I2F 0
#endif
__device__ __forceinline__ uint32_t bfind_fast(uint32_t v) {
#if BFIND_I2F
    float f; asm("cvt.rz.f32.u32 %0, %1;" : "=f"(f) : "r"(v));
    return (__float_as_uint(f) >> 23) - 127u;
#else
    return bfind(v);
#endif
}
__device__ __forceinline__ int ilog2u(uint32_t v) { return 31 - __clz(v); }



// ------------------------------------------------------------------------------------------------ ring bit reader
// Every lane walks its own serial bitstream.  The bytes are staged through shared memory by per-lane cp.async
// (LDGSTS, 16 B each) into a private 128-byte ring that runs several blocks ahead of the read position, so the serial
// walk never waits on HBM.  Reads are position based: two LDS + byte swaps + one funnel shift give a 32-bit window; block
// 0 of the ring is duplicated behind block 7, so the second word of a window never needs a wrap-around address.
// Refill is CHECKPOINTED: all lanes of a warp top up their rings at the same loop iterations (every 8 samples): first
// wait for what was requested one period ago, then request more.  Budget: blocks up to (block of pos at the previous
// checkpoint) + 7 have landed, i.e. >= 113 bytes past that position; a period may therefore advance by A bytes with
// 2A + 8 <= 113 (this period's reads reach pos_prev + 2A + 8).  The walkers keep A <= 42: eight samples of at most 32
// bits each plus partition parameters; anything longer takes a synchronous path (ensure_now).
// DEPTH = how many of the most recent refill groups may still be in flight after a checkpoint (prefetch distance in
// checkpoint periods).  With DEPTH = 0 a checkpoint waits for what was requested one period earlier, which exposes the
// HBM latency whenever too few warps are resident to hide it (streams with few, large frames).  With DEPTH > 0 the
// checkpoint only waits for older groups, provided what those covered (mark[DEPTH]) reaches past everything the coming
// period can read; a lane that consumed unusually many bits falls back to a full wait.
// DUP = false (k_decode): no duplicate of block 0 -- every reader wraps its own addresses (three masks more per group of 8 codewords)
// and a refill is one cp.async per block instead of a second, almost always predicated-off one plus its branch.
#ifndef RING_LEAN_FETCH
#define RING_LEAN_FETCH 0
#endif
template <int NBLK_, int DEPTH, int STEADY_ = 0, bool DUP = true>
struct RingBitsT {
    static constexpr int NBLK = NBLK_, BLK = 16, RB_BYTES = NBLK * BLK;
    static constexpr int STRIDE = DUP ? RB_BYTES + BLK : RB_BYTES;    // per-lane footprint: the ring (+ the duplicate of block 0)
    // DUP = false: the lane's ring is RB_BYTES-aligned and `sring` holds its address OR-ed with a per-lane swizzle of the block index
    // (bits 4..): a byte offset x of the ring lives at (x & (RB_BYTES - 4)) ^ sring -- ONE three-input logic instruction where base + masked
    // offset are two, on every codeword of the branch-free groups (the word after the window is fetched on each step) -- and lanes that
    // read or fill the same ring offset spread over NBLK groups of four banks as the 144-byte stride of the DUP layout did.
    __device__ __forceinline__ uint32_t ad(uint32_t x) const { return DUP ? sring + (x & (RB_BYTES - 4)) : ((x & (RB_BYTES - 4)) ^ sring); }
    __device__ __forceinline__ static uint32_t ring_handle(uint32_t lane_base) {
        return DUP ? lane_base : (lane_base | (((lane_base / RB_BYTES) & (NBLK - 1)) * BLK));
    }
    // bytes a refill period can advance + window look-ahead: 8 samples of <= 32 bits plus parameters per 8-sample group
    // (longer codewords take synchronous refills); rings of 16 blocks are checkpointed every 16 samples
    static constexpr uint32_t PERIOD_REACH = (NBLK >= 16 ? 84 : 42) + 16;
    uint32_t sring;        // shared-space address of this lane's ring
    uint32_t pos;          // bit position relative to g0
    uint32_t filled;       // blocks [.., filled) have been requested
    uint32_t navail;       // whole 16-byte blocks readable from g0 (blocks past the padded input read as zero)
    uint32_t mark[DEPTH + 1];   // `filled` after each of the last DEPTH + 1 checkpoints (mark[0] most recent)
    const uint8_t* g0;     // global address of ring byte 0 (16 B aligned)

    __device__ __forceinline__ void fetch(uint32_t b) {
        const bool in = b < navail;
        const uint32_t n = in ? 16u : 0u;
        // k_decode's ring: one select on the block index and one wide multiply-add (request loop 26 -> 13 instructions per block together
        // with DUP = false; cfg2 decode 1.70 -> 1.67 ms).  k_parse keeps the other form: the same change made the guessed walks of cfg3 3 % slower.
        const uint8_t* s = (!DUP || RING_LEAN_FETCH) ? g0 + (uint64_t)(in ? b : 0u) * BLK : g0 + (n ? (uint64_t)b * BLK : 0ull);
        const uint32_t slot = b & (NBLK - 1);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(ad(slot * BLK)), "l"(s), "r"(n) : "memory");
        if (DUP && slot == 0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sring + RB_BYTES), "l"(s), "r"(n) : "memory");
    }
    __device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
    __device__ __forceinline__ void wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
    // fetch(b) under a predicate, without a branch (a cp.async of size 0 would zero the slot, so the instruction itself is
    // predicated)
    __device__ __forceinline__ void fetch_if(uint32_t b, bool p) {
        const uint32_t n = b < navail ? 16u : 0u;
        const uint8_t* s = g0 + (n ? (uint64_t)b * BLK : 0ull);
        const uint32_t slot = b & (NBLK - 1);
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q cp.async.cg.shared.global [%0], [%1], 16, %2;\n\t}" ::"r"(ad(slot * BLK)), "l"(s), "r"(n), "r"((uint32_t)p) : "memory");
        if (DUP) asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q cp.async.cg.shared.global [%0], [%1], 16, %2;\n\t}" ::"r"(sring + RB_BYTES), "l"(s), "r"(n), "r"((uint32_t)(p && slot == 0)) : "memory");
    }
    // Request every block the ring has room for (never the slot being read).  STEADY = how many blocks a refill period
    // normally frees: that many are requested branch-free under predicates (the lanes of a warp free different numbers of
    // blocks, so a loop here diverges on nearly every call); whatever is left after a big move goes through the loop.
    // STEADY = 0 keeps the plain loop: with many resident warps (streams of many frames) the predicated slots that turn out
    // empty cost more issue slots than the divergence they avoid (cfg2: parse +3 %), with few warps the branches dominate
    // (cfg3: parse 5.30 -> 4.56 ms, decode 4.65 -> 4.50 ms).
    static constexpr int STEADY = STEADY_;
    __device__ __forceinline__ void request() {
        const uint32_t curblk = pos / (BLK * 8);
        if (filled < curblk) filled = curblk;       // jumped over unrequested blocks
        const uint32_t lim = curblk + NBLK;
#pragma unroll
        for (int i = 0; i < STEADY; i++) { const bool p = filled < lim; fetch_if(filled, p); filled += p ? 1u : 0u; }
        if (filled < lim) {
#pragma unroll 1
            do { fetch(filled); filled++; } while (filled < lim);
        }
        commit();                                   // one group per call, possibly empty: wait_group counts calls
    }
    __device__ __forceinline__ void checkpoint() {
        if constexpr (DEPTH == 0) wait_all();
        else {
            if ((pos >> 3) + PERIOD_REACH <= mark[DEPTH] * (uint32_t)BLK) asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH) : "memory");
            else {
                wait_all();
#pragma unroll
                for (int d = 0; d <= DEPTH; d++) mark[d] = filled;
            }
        }
        request();
#pragma unroll
        for (int d = DEPTH; d > 0; d--) mark[d] = mark[d - 1];
        mark[0] = filled;
    }
    // k_decode's schedule for an 8-block ring: ONE full checkpoint per 16 samples (wait for what the previous one requested, request
    // again) and, after the first 8 of them, only a test whether what has landed still covers the next 8 samples (PERIOD_REACH bytes
    // past the position).  mark[0] = blocks known to have landed.  A step of 8 samples advances by A <= 42 bytes, so after a wait
    // everything up to (block of the position at the last request + NBLK) is there: >= 113 bytes past that position, and
    // 42 + PERIOD_REACH <= 113.  The request loop costs ~24 instructions per block whatever the period; what this halves is the
    // fixed part (wait, compare, branch, commit) -- ncu: the 8-sample checkpoint was 11.5 % of k_decode's instructions and 15.6 %
    // of its stall samples on the 24-bit stereo stream.
    __device__ __forceinline__ void ckpt_full() {
        wait_all();
        const uint32_t landed = filled;
        request();
        if ((pos >> 3) + PERIOD_REACH > landed * (uint32_t)BLK) { wait_all(); mark[0] = filled; }     // > 55 bytes in 16 samples: rare
        else mark[0] = landed;
    }
    __device__ __forceinline__ bool ckpt_mid_needed() const { return (pos >> 3) + PERIOD_REACH > mark[0] * (uint32_t)BLK; }
    __device__ __forceinline__ void ckpt_mid_wait() { wait_all(); mark[0] = filled; }
    __device__ __forceinline__ void ensure_now() {                               // synchronous: rare big moves, init
        request(); wait_all();
#pragma unroll
        for (int d = 0; d <= DEPTH; d++) mark[d] = filled;
    }
    __device__ __forceinline__ void init(uint32_t sring_, const uint8_t* in, uint64_t in_len, uint64_t abs_bit) {
        sring = ring_handle(sring_);
        const uint64_t b0 = (abs_bit >> 3) & ~(uint64_t)(BLK - 1);
        g0 = in + b0;
        const uint64_t nb = in_len > b0 ? (in_len - b0) >> 4 : 0;
        navail = (uint32_t)(nb > 0xffffffffull ? 0xffffffffull : nb);
        pos = (uint32_t)(abs_bit - b0 * 8);
        filled = 0;
        ensure_now();
    }
    __device__ __forceinline__ void init_idle(uint32_t sring_, const uint8_t* in) {
        sring = ring_handle(sring_); g0 = in; navail = 0; pos = 0; filled = NBLK;
#pragma unroll
        for (int d = 0; d <= DEPTH; d++) mark[d] = NBLK;
    }
    __device__ __forceinline__ uint64_t abs_pos(const uint8_t* in) const { return (uint64_t)(g0 - in) * 8 + pos; }
    __device__ __forceinline__ uint32_t window_at(uint32_t p) const {      // 32 bits starting at bit p, MSB first
        const uint32_t bo = (p >> 3) & (RB_BYTES - 4);
        const uint32_t a = lds32(ad(bo)), b = lds32(DUP ? sring + bo + 4 : ad(bo + 4));
        return __funnelshift_l(__byte_perm(b, 0, 0x0123), __byte_perm(a, 0, 0x0123), p);
    }
    __device__ __forceinline__ uint32_t window() const { return window_at(pos); }
    // Register-cached window for the branch-free groups: w0:w1 are the two big-endian words under the read position, w2
    // the word after them.  A codeword advances the position by at most 32 bits, so at most one word is crossed per
    // step.  The word that a crossing shifts in (position word + 3) is loaded at the START of every step, from the
    // position the step starts at, so the shared-memory latency is entirely off the serial
    // position -> window -> length -> position chain (which is then SHF, FLO, IADD3, LOP3, SEL).
    struct Win3 { uint32_t w0, w1, w2; };
    __device__ __forceinline__ Win3 win_init(uint32_t p) const {
        Win3 w;
        if (DUP) {
            const uint32_t ad = sring + ((p >> 3) & (RB_BYTES - 4));
            w.w0 = __byte_perm(lds32(ad), 0, 0x0123); w.w1 = __byte_perm(lds32(ad + 4), 0, 0x0123); w.w2 = __byte_perm(lds32(ad + 8), 0, 0x0123);
        } else {
            const uint32_t by = p >> 3;
            w.w0 = __byte_perm(lds32(ad(by)), 0, 0x0123);
            w.w1 = __byte_perm(lds32(ad(by + 4)), 0, 0x0123);
            w.w2 = __byte_perm(lds32(ad(by + 8)), 0, 0x0123);
        }
        return w;
    }
    __device__ __forceinline__ uint32_t win_next(uint32_t p) const { return __byte_perm(lds32(ad((p >> 3) + 12)), 0, 0x0123); }   // word of p, + 3
    __device__ __forceinline__ static uint32_t win_peek(const Win3& w, uint32_t p) { return __funnelshift_l(w.w1, w.w0, p); }
    __device__ __forceinline__ static void win_advance(Win3& w, uint32_t p, uint32_t np, uint32_t nxt) {
        if ((p ^ np) & 32u) { w.w0 = w.w1; w.w1 = w.w2; w.w2 = nxt; }
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
    // one Rice codeword with parameter k, any length; leaves the ring synchronised when the codeword was long
    __device__ __forceinline__ int32_t rice_careful(uint32_t k) {
        const uint32_t w = window();
        const uint32_t f = bfind(w);
        uint32_t u;
        if ((int32_t)(f - k) >= 0) { u = (31u - f) << k | (shr_c(w, f - k) & ((1u << k) - 1u)); pos += k + 32u - f; }
        else { const uint32_t q = unary(1u << 24); ensure_now(); u = (q << k) | get(k); ensure_now(); }
        return (int32_t)(u >> 1) ^ -(int32_t)(u & 1);
    }
    // skip one Rice codeword; returns false when the unary run is implausibly long (damaged data)
    __device__ __forceinline__ bool rice_skip_careful(uint32_t k, uint32_t limit = 1u << 16) {
        const uint32_t w = window();
        const uint32_t f = bfind(w);
        if ((int32_t)(f - k) >= 0) { pos += k + 32u - f; return true; }
        const uint32_t q = unary(limit);
        jump(k);
        return q <= limit;
    }
};

#ifndef PARSE_RING_BLOCKS
#define PARSE_RING_BLOCKS 16
#endif
#ifndef PARSE_RING_DEPTH
#define PARSE_RING_DEPTH 2
#endif
#ifndef DEC_RING_DEPTH
#define DEC_RING_DEPTH 1
#endif
#ifndef PARSE_RING_NODUP
#define PARSE_RING_NODUP 1
#endif
using ParseBits = RingBitsT<PARSE_RING_BLOCKS, PARSE_RING_DEPTH, 0, !PARSE_RING_NODUP>;   // no sample tile in k_parse: room for a longer ring and a deeper prefetch
template <bool LEAN> using ParseBitsT = RingBitsT<PARSE_RING_BLOCKS, PARSE_RING_DEPTH, LEAN ? 3 : 0, !PARSE_RING_NODUP>;
#ifndef DEC_RING_BLOCKS
#define DEC_RING_BLOCKS 8
#endif
#ifndef DEC_RING_STEADY_SMALL
#define DEC_RING_STEADY_SMALL 0
#endif
#ifndef DEC_RING_NODUP
#define DEC_RING_NODUP 1
#endif
using RingBits = RingBitsT<DEC_RING_BLOCKS, DEC_RING_DEPTH, 0, !DEC_RING_NODUP>;         // k_decode: geometry (STRIDE) of the ring every DecRing<ORD> shares
template <int ORD> using DecRing = RingBitsT<DEC_RING_BLOCKS, DEC_RING_DEPTH, (ORD > 12 ? 2 : DEC_RING_STEADY_SMALL), !DEC_RING_NODUP>;   // orders > 12: few, long subframes


// ------------------------------------------------------------------------------------------------ launch helpers
static inline cudaStream_t S(void* s) { return (cudaStream_t)s; }
static inline uint32_t blocks_for(uint64_t n, uint32_t per) { uint64_t b = (n + per - 1) / per; return (uint32_t)(b ? b : 1); }
int sm_count();                                             // of the current device (kernels.cu)
bool first_use_on_device(std::atomic<uint64_t>& done);     // per-device once-only work still to be done for the caller's bitmask? (kernels.cu)
void used_on_device(std::atomic<uint64_t>& done);          // ... it has been done

} // namespace bnf
