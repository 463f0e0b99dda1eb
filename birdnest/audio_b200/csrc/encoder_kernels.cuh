// encoder_kernels.cuh -- device side of the FLAC encoder (see encoder.cu for the design).  Kept free of host runtime calls so that
// tools/enc_emu.cpp can compile the very same kernels for the CPU (one pthread per CUDA thread) and run them against the oracle and
// the reference decoder where no GPU exists.
#pragma once
#include <cstdint>
#ifdef BNFLAC_EMU
#define ENC_DYN_SMEM(name) uint8_t* name = emu_dyn_smem
#else
#define ENC_DYN_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#endif

namespace bnfe {

#define FULL 0xffffffffu
// threads per CTA of both frame kernels: a template parameter NT (256 for blocks up to 4608 samples, 1024 above: a CTA of 256 threads with
// 130 KB of staged samples would be alone on its SM with two warps per scheduler)
constexpr int NT_SMALL = 256, NT_BIG = 1024, NW_MAX = NT_BIG / 32;
constexpr uint32_t NT_BIG_FROM_BS = 4609;
constexpr int MAX_PO = 8;
constexpr uint32_t MAX_BS = 16384;
constexpr int MAX_LPC = 32;

struct EncSub {                     // decision record of one subframe, 336 bytes
    uint8_t type;                   // 0 CONSTANT 1 VERBATIM 2 FIXED 3 LPC
    uint8_t order, wasted, prec, shift, po, method;
    uint8_t variant;                // which signal: channel index, or for stereo 0 L, 1 R, 2 M, 3 S
    uint32_t bits;                  // exact size of the subframe in bits
    int32_t cval;                   // CONSTANT: the value (already shifted by `wasted`)
    int16_t qc[MAX_LPC];
    uint8_t k[1 << MAX_PO];         // Rice parameter per partition
};
static_assert(sizeof(EncSub) == 336, "EncSub layout");

struct EncFrame {                   // 80 bytes
    uint64_t byte_off;              // where the frame starts in the output stream (k_enc_scan)
    uint32_t nbytes;                // frame size, CRC-16 included
    uint32_t bs;
    uint8_t assignment, hdr_len, pad[2];
    uint8_t hdr[16];
    uint32_t sub_bit[8];            // bit offset of every subframe from the frame's first byte
    uint32_t pad2[3];
};
static_assert(sizeof(EncFrame) == 80, "EncFrame layout");

struct EncTotals { uint64_t total_bytes; uint32_t min_fs, max_fs; };

struct EncArgs {
    const uint8_t* pcm;             // interleaved little-endian PCM, `bin` bytes per sample
    uint64_t total_samples;         // per channel
    uint32_t ch, bps, bin, bs, sample_rate;
    uint32_t max_lpc, prec, min_po, max_po, stereo, search_order;
    uint32_t nframes;
    uint64_t first_frame;           // bytes of metadata before the first frame
    uint64_t first_number;          // coded number of the first frame (a stream encoded in several calls)
    EncSub* sub;                    // [nframes][8]
    EncFrame* frm;
    EncTotals* totals;
    uint8_t* out;
};

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ int32_t load_sample(const uint8_t* __restrict__ pcm, uint64_t idx, uint32_t bin) {
    const uint8_t* p = pcm + idx * bin;
    if (bin == 4) return (int32_t)__ldg(reinterpret_cast<const uint32_t*>(p));
    if (bin == 2) return (int32_t)(int16_t)__ldg(reinterpret_cast<const uint16_t*>(p));
    uint32_t v = __ldg(p);
    if (bin >= 2) v |= (uint32_t)__ldg(p + 1) << 8;
    if (bin >= 3) v |= (uint32_t)__ldg(p + 2) << 16;
    const uint32_t sh = 32u - 8u * bin;
    return (int32_t)(v << sh) >> sh;
}
// sample i of signal `variant` of the frame whose first interleaved sample is `base`
__device__ __forceinline__ int32_t load_variant(const EncArgs& a, uint64_t base, uint32_t i, uint32_t variant) {
    if (!a.stereo) return load_sample(a.pcm, base + (uint64_t)i * a.ch + variant, a.bin);
    const int32_t L = load_sample(a.pcm, base + 2ull * i, a.bin), R = load_sample(a.pcm, base + 2ull * i + 1, a.bin);
    return variant == 0 ? L : variant == 1 ? R : variant == 2 ? (L + R) >> 1 : L - R;
}
// residuals are written with consecutive lanes on consecutive samples and read back by threads that own RUNS of consecutive samples
// (stride = run length between lanes): one padding word per 32 keeps both patterns (nearly) free of bank conflicts
__device__ __forceinline__ uint32_t ridx(uint32_t i) { return i + (i >> 5); }
// dynamic shared memory of both frame kernels: pad[XPAD] | x[bs] | r[bs + bs/32 + 1].  The padding lets the unrolled predictor loops
// read x[i - 1 - j] for taps beyond the order (zero coefficients) without a bounds test.
constexpr uint32_t XPAD = 32;
static inline size_t enc_smem_bytes(uint32_t bs) { return ((size_t)XPAD + (size_t)bs * 2 + bs / 32 + 1) * 4; }
__device__ __forceinline__ uint32_t zigzag(int32_t r) { return ((uint32_t)r << 1) ^ (uint32_t)(r >> 31); }

// block-wide reductions; every thread gets the result.  `scratch` has NW entries and is reused call after call.
template <int NT>
__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v, unsigned long long* scratch) {
    constexpr int NW = NT / 32;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long t = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) t += scratch[w];
    return t;
}
template <int NT>
__device__ __forceinline__ uint32_t block_or_u32(uint32_t v, unsigned long long* scratch) {
    constexpr int NW = NT / 32;
    v = __reduce_or_sync(FULL, v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) t |= (uint32_t)scratch[w];
    return t;
}
// exclusive prefix sum over the block's threads; *total = sum of all
template <int NT>
__device__ __forceinline__ uint32_t block_excl_scan_u32(uint32_t v, unsigned long long* scratch, uint32_t* total) {
    constexpr int NW = NT / 32;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, inc, o); if (lane >= (uint32_t)o) inc += t; }
    __syncthreads();
    if (lane == 31) scratch[wid] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) { const uint32_t s = (uint32_t)scratch[w]; if ((uint32_t)w < wid) base += s; tot += s; }
    *total = tot;
    return base + inc - v;
}

// FLAC frame-header pieces
__device__ __forceinline__ uint32_t crc8_bytes(const uint8_t* p, uint32_t n) {
    uint32_t c = 0;
    for (uint32_t i = 0; i < n; i++) {
        c ^= p[i];
        for (int k = 0; k < 8; k++) c = (c & 0x80) ? ((c << 1) ^ 0x07) & 0xFF : (c << 1) & 0xFF;
    }
    return c;
}
__device__ __forceinline__ uint32_t put_utf8(uint8_t* o, uint64_t v) {
    if (v < 0x80) { o[0] = (uint8_t)v; return 1; }
    const uint32_t n = v < 0x800 ? 2 : v < 0x10000 ? 3 : v < 0x200000 ? 4 : v < 0x4000000 ? 5 : v < 0x80000000ull ? 6 : 7;
    for (uint32_t i = n - 1; i >= 1; i--) { o[i] = (uint8_t)(0x80 | (v & 0x3f)); v >>= 6; }
    o[0] = (uint8_t)((0xFF << (8 - n)) | (n == 7 ? 0 : v));
    return n;
}
__device__ uint32_t make_frame_header(uint8_t* h, const EncArgs& a, uint32_t bs, uint32_t assignment, uint64_t number) {
    const uint32_t bst[16] = {0, 192, 576, 1152, 2304, 4608, 0, 0, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768};
    const uint32_t srt[12] = {0, 88200, 176400, 192000, 8000, 16000, 22050, 24000, 32000, 44100, 48000, 96000};
    uint32_t bsc = 0, src = 0, ssc = 0;
    for (uint32_t i = 1; i < 16; i++) if (bst[i] == bs) bsc = i;
    if (!bsc) bsc = (bs <= 256) ? 6 : 7;
    for (uint32_t i = 1; i < 12; i++) if (srt[i] == a.sample_rate) src = i;
    if (!src) {
        if (a.sample_rate % 1000 == 0 && a.sample_rate / 1000 < 256) src = 12;
        else if (a.sample_rate < 65536) src = 13;
        else if (a.sample_rate % 10 == 0 && a.sample_rate / 10 < 65536) src = 14;
    }
    switch (a.bps) { case 8: ssc = 1; break; case 12: ssc = 2; break; case 16: ssc = 4; break; case 20: ssc = 5; break; case 24: ssc = 6; break; default: ssc = 0; }
    uint32_t q = 0;
    h[q++] = 0xFF; h[q++] = 0xF8;
    h[q++] = (uint8_t)(bsc << 4 | src);
    h[q++] = (uint8_t)(assignment << 4 | ssc << 1);
    q += put_utf8(h + q, number);
    if (bsc == 6) h[q++] = (uint8_t)(bs - 1);
    else if (bsc == 7) { h[q++] = (uint8_t)((bs - 1) >> 8); h[q++] = (uint8_t)(bs - 1); }
    if (src == 12) h[q++] = (uint8_t)(a.sample_rate / 1000);
    else if (src == 13) { h[q++] = (uint8_t)(a.sample_rate >> 8); h[q++] = (uint8_t)a.sample_rate; }
    else if (src == 14) { h[q++] = (uint8_t)((a.sample_rate / 10) >> 8); h[q++] = (uint8_t)(a.sample_rate / 10); }
    h[q] = (uint8_t)crc8_bytes(h, q); q++;
    return q;
}

// residuals of an LPC predictor, coefficients in registers, taps unrolled (taps beyond the order have zero coefficients and read the
// padding in front of x).  Returns 1 in *big if a residual leaves the range the Rice coder is given.
template <int NO, int NT>
__device__ __forceinline__ uint32_t lpc_residuals(const int32_t* __restrict__ x, int32_t* __restrict__ r, uint32_t bs, uint32_t lo, const int32_t* qc_sh, uint32_t shift, bool narrow) {
    int32_t c[NO];
#pragma unroll
    for (int j = 0; j < NO; j++) c[j] = (uint32_t)j < lo ? qc_sh[j] : 0;
    uint32_t big = 0;
    for (uint32_t i = threadIdx.x; i < bs; i += NT) {
        if (i < lo) continue;
        long long s = 0;
#pragma unroll
        for (int j = 0; j < NO; j++) s += (long long)c[j] * (long long)x[(int)i - 1 - j];
        const int32_t p = narrow ? ((int32_t)(uint32_t)s >> shift) : (int32_t)(s >> shift);
        const long long rr = (long long)x[i] - (long long)p;
        if (rr > 0x3fffffffll || rr < -0x3fffffffll) big = 1;
        r[ridx(i)] = (int32_t)rr;
    }
    return big;
}
template <int NT>
__device__ __forceinline__ uint32_t lpc_residuals_any(const int32_t* x, int32_t* r, uint32_t bs, uint32_t lo, const int32_t* qc_sh, uint32_t shift, bool narrow) {
    if (lo <= 4) return lpc_residuals<4, NT>(x, r, bs, lo, qc_sh, shift, narrow);
    if (lo <= 8) return lpc_residuals<8, NT>(x, r, bs, lo, qc_sh, shift, narrow);
    if (lo <= 12) return lpc_residuals<12, NT>(x, r, bs, lo, qc_sh, shift, narrow);
    if (lo <= 16) return lpc_residuals<16, NT>(x, r, bs, lo, qc_sh, shift, narrow);
    return lpc_residuals<32, NT>(x, r, bs, lo, qc_sh, shift, narrow);
}

// ------------------------------------------------------------------------------------------------ plan kernel
constexpr int NPART = (2 << MAX_PO) - 1;         // partitions of all orders 0..MAX_PO; order L starts at (1 << L) - 1
struct PlanShared {
    double ac[MAX_LPC + 1];
    double err[MAX_LPC + 1];
    double est[MAX_LPC + 1];
    double lpcs[MAX_LPC * (MAX_LPC + 1) / 2];  // predictor of order o (Levinson-Durbin stage o) at o (o - 1) / 2 ... + o
    unsigned long long psum[NPART];           // sum of zig-zagged residuals per partition, every partition order
    uint32_t pb[3][NPART];                    // sum of (u >> k) per partition for the three candidate parameters
    uint8_t k0[NPART];
    uint8_t kb[NPART];                        // best parameter per partition
    uint32_t lvl_bits[MAX_PO + 1], lvl_big[MAX_PO + 1];
    unsigned long long lvl_est[MAX_PO + 1];
    unsigned long long scratch[NW_MAX];
    unsigned long long red5[5][NW_MAX];
    EncSub dec[8];
    EncSub tmp;
    int32_t qc[MAX_LPC];
    uint32_t lpc_order, lpc_prec, lpc_shift, lpc_ok;
    uint8_t hdr[16];
};

// The cheapest partitioned-Rice plan for residuals r[order..bs) (indexed by sample position).  Result in d->po / method / k[];
// returns the size of the residual section in bits (method + partition order + parameters + codewords).
// Two passes over the residuals: sums of u per partition of the finest order -> every coarser order by addition -> the partition
// order with the smallest ESTIMATED size (libFLAC's estimate from the sums; sizing every order exactly cost 7 x the instructions
// for 0.1 % of the bytes) -> three candidate parameters around log2(mean) per partition of that order, sized EXACTLY.
template <int NT>
__device__ uint32_t plan_rice(const int32_t* __restrict__ r, uint32_t bs, uint32_t order, uint32_t min_po, uint32_t max_po, PlanShared& sh, EncSub* d) {
    const uint32_t tid = threadIdx.x;
    // valid partition orders: the block divides evenly and partition 0 keeps at least one residual
    uint32_t hi = 0;
    for (uint32_t L = 1; L <= max_po && L <= (uint32_t)MAX_PO; L++) { if ((bs & ((1u << L) - 1)) == 0 && (bs >> L) > order) hi = L; else break; }
    const uint32_t lo = min_po < hi ? min_po : hi;
    const uint32_t nper = (bs + NT - 1) / NT;
    const uint32_t i_begin = min(bs, max(order, tid * nper)), i_end = min(bs, (tid + 1) * nper);
    const uint32_t e_lo = (1u << lo) - 1u, e_hi = (2u << hi) - 1u;        // entries [e_lo, e_hi) of the per-partition tables
    for (uint32_t e = e_lo + tid; e < e_hi; e += NT) { sh.psum[e] = 0; sh.pb[0][e] = 0; sh.pb[1][e] = 0; sh.pb[2][e] = 0; }
    if (tid <= (uint32_t)MAX_PO) { sh.lvl_bits[tid] = 0; sh.lvl_big[tid] = 0; sh.lvl_est[tid] = 0; }
    __syncthreads();
    {   // finest order: sums of u
        const uint32_t psz = bs >> hi, base = (1u << hi) - 1u;
        uint32_t i = i_begin;
        while (i < i_end) {
            const uint32_t q = i / psz, stop = min(i_end, (q + 1) * psz);
            unsigned long long s = 0;
            for (; i < stop; i++) s += zigzag(r[ridx(i)]);
            atomicAdd(&sh.psum[base + q], s);
        }
    }
    __syncthreads();
    // Every coarser order straight from the finest sums (no barrier per order); per partition the parameter near log2(mean) and
    // libFLAC's estimate of its size from the sum alone: n (k + 1) + (sum >> k) - n / 2 (each floor loses half a bit on average).
    for (uint32_t e = e_lo + tid; e < e_hi; e += NT) {
        const uint32_t L = 31u - (uint32_t)__clz((int)(e + 1u)), q = e + 1u - (1u << L);
        unsigned long long sum;
        if (L == hi) sum = sh.psum[e];
        else {
            const uint32_t span = 1u << (hi - L), fb = (1u << hi) - 1u + q * span;
            sum = 0;
            for (uint32_t j = 0; j < span; j++) sum += sh.psum[fb + j];
        }
        const uint32_t n = (bs >> L) - (q == 0 ? order : 0);
        const unsigned long long mean = sum / n;
        uint32_t k = mean ? 63u - (uint32_t)__clzll((long long)mean) : 0u;
        if (k > 29) k = 29;
        if (k < 1) k = 1;           // candidates k-1, k, k+1
        sh.k0[e] = (uint8_t)k;
        unsigned long long est = ~0ull;
        for (uint32_t kk = k - 1; kk <= k + 1; kk++) {
            const unsigned long long body = sum >> kk, half = kk ? n / 2 : 0;
            const unsigned long long t = (unsigned long long)n * (kk + 1) + (body > half ? body - half : 0);
            if (t < est) est = t;
        }
        atomicAdd(&sh.lvl_est[L], est + (k >= 14 ? 5u : 4u));
    }
    __syncthreads();
    uint32_t best_L = hi;
    {
        unsigned long long be = ~0ull;
        for (int L = (int)hi; L >= (int)lo; L--) if (sh.lvl_est[L] < be) { be = sh.lvl_est[L]; best_L = (uint32_t)L; }
    }
    // exact sizes of the three candidate parameters of every partition of the chosen order
    const uint32_t psz = bs >> best_L, base = (1u << best_L) - 1u, np = 1u << best_L;
    {
        uint32_t i = i_begin;
        while (i < i_end) {
            const uint32_t q = i / psz, stop = min(i_end, (q + 1) * psz);
            const uint32_t k = sh.k0[base + q];
            uint32_t s0 = 0, s1 = 0, s2 = 0;
            for (; i < stop; i++) { const uint32_t u = zigzag(r[ridx(i)]); s0 += u >> (k - 1); s1 += u >> k; s2 += u >> (k + 1); }
            atomicAdd(&sh.pb[0][base + q], s0); atomicAdd(&sh.pb[1][base + q], s1); atomicAdd(&sh.pb[2][base + q], s2);
        }
    }
    __syncthreads();
    for (uint32_t q = tid; q < np; q += NT) {
        const uint32_t e = base + q;
        const uint32_t n = psz - (q == 0 ? order : 0), k = sh.k0[e];
        uint32_t bb = n * k + sh.pb[0][e], kk = k - 1;                      // n * (kk + 1) + sum(u >> kk)
        const uint32_t b1 = n * (k + 1) + sh.pb[1][e], b2 = n * (k + 2) + sh.pb[2][e];
        if (b1 < bb) { bb = b1; kk = k; }
        if (b2 < bb) { bb = b2; kk = k + 1; }
        d->k[q] = (uint8_t)kk;
        atomicAdd(&sh.lvl_bits[best_L], bb);
        if (kk > 14) atomicOr(&sh.lvl_big[best_L], 1u);
    }
    __syncthreads();
    const uint32_t best_bits = sh.lvl_bits[best_L] + np * (sh.lvl_big[best_L] ? 5u : 4u) + 6u;
    if (tid == 0) { d->po = (uint8_t)best_L; d->method = (uint8_t)(sh.lvl_big[best_L] ? 1 : 0); }
    __syncthreads();
    return best_bits;
}

template <int NT>
__global__ void __launch_bounds__(NT) k_enc_plan(EncArgs a) {
    constexpr int NW = NT / 32;
    ENC_DYN_SMEM(dyn);
    __shared__ PlanShared sh;
    const uint32_t tid = threadIdx.x, f = blockIdx.x;
    const uint64_t s0 = (uint64_t)f * a.bs;
    const uint32_t bs = (uint32_t)min((uint64_t)a.bs, a.total_samples - s0);
    int32_t* x = reinterpret_cast<int32_t*>(dyn) + XPAD;
    if (threadIdx.x < XPAD) x[(int)threadIdx.x - (int)XPAD] = 0;
    int32_t* r = x + a.bs;
    int32_t* wi = r;                                  // the windowed signal lives where the residuals go later
    const uint64_t base = s0 * a.ch;
    const uint32_t nvar = a.stereo ? 4u : a.ch;

    for (uint32_t v = 0; v < nvar; v++) {
        EncSub* d = &sh.dec[v];
        const uint32_t vbps = a.bps + ((a.stereo && v == 3) ? 1u : 0u);
        // ---- stage the signal, wasted bits, CONSTANT
        const int32_t x0 = load_variant(a, base, 0, v);
        uint32_t orv = 0, differs = 0;
        {   // four samples per trip: the global loads of a trip are issued together (the staging loop is what waits for memory)
            uint32_t i = tid;
            for (; i + 3 * NT < bs; i += 4 * NT) {
                int32_t sv[4];
#pragma unroll
                for (int q = 0; q < 4; q++) sv[q] = load_variant(a, base, i + q * NT, v);
#pragma unroll
                for (int q = 0; q < 4; q++) { x[i + q * NT] = sv[q]; orv |= (uint32_t)sv[q]; differs |= (uint32_t)(sv[q] != x0); }
            }
            for (; i < bs; i += NT) { const int32_t s = load_variant(a, base, i, v); x[i] = s; orv |= (uint32_t)s; differs |= (uint32_t)(s != x0); }
        }
        {   // one exchange for both
            orv = __reduce_or_sync(FULL, orv); differs = __reduce_or_sync(FULL, differs);
            __syncthreads();
            if ((tid & 31) == 0) sh.scratch[tid >> 5] = (unsigned long long)differs << 32 | orv;
            __syncthreads();
            unsigned long long t = 0;
#pragma unroll
            for (int wv = 0; wv < NW; wv++) t |= sh.scratch[wv];
            orv = (uint32_t)t; differs = (uint32_t)(t >> 32);
        }
        uint32_t w = orv ? (uint32_t)__ffs((int)orv) - 1u : 0u;
        if (w >= vbps) w = 0;
        if (w) { for (uint32_t i = tid; i < bs; i += NT) x[i] >>= w; }
        const uint32_t ebps = vbps - w;
        __syncthreads();
        if (tid == 0) { d->variant = (uint8_t)v; d->wasted = (uint8_t)w; d->order = 0; d->prec = 0; d->shift = 0; d->po = 0; d->method = 0; d->cval = x[0]; }
        if (!differs) {
            if (tid == 0) { d->type = 0; d->bits = 8 + w + ebps; }
            __syncthreads();
            continue;
        }
        const uint32_t verb_bits = 8 + w + bs * ebps;
        // ---- FIXED: order with the smallest sum of |residual|
        // samples have at most 25 bits (24 + the side channel's extra one): every difference up to order 4 fits 29 bits, so the
        // arithmetic is 32-bit and no residual can leave the range the Rice coder takes
        unsigned long long fs[5] = {0, 0, 0, 0, 0};
        const uint32_t fbig = 0;
        if (tid < bs) {         // the thread's first sample may lie among the first four: orders above its index do not count it
            const uint32_t i = tid;
            const int32_t v0 = x[i], v1 = x[(int)i - 1], v2 = x[(int)i - 2], v3 = x[(int)i - 3], v4 = x[(int)i - 4];   // padding reads 0
            const int32_t p1 = v1 - v2, p2 = p1 - (v2 - v3), p3 = p2 - ((v2 - v3) - (v3 - v4));
            int32_t dd[5];
            dd[0] = v0; dd[1] = v0 - v1; dd[2] = dd[1] - p1; dd[3] = dd[2] - p2; dd[4] = dd[3] - p3;
#pragma unroll
            for (int o = 0; o < 5; o++) if (i >= (uint32_t)o) fs[o] += (uint32_t)(dd[o] < 0 ? -dd[o] : dd[o]);
        }
        for (uint32_t i = tid + NT; i < bs; i += NT) {
            const int32_t v0 = x[i], v1 = x[i - 1], v2 = x[i - 2], v3 = x[i - 3], v4 = x[i - 4];
            const int32_t p1 = v1 - v2, p2 = p1 - (v2 - v3), p3 = p2 - ((v2 - v3) - (v3 - v4));       // differences ending at i - 1
            const int32_t d1 = v0 - v1, d2 = d1 - p1, d3 = d2 - p2, d4 = d3 - p3;
            fs[0] += (uint32_t)(v0 < 0 ? -v0 : v0); fs[1] += (uint32_t)(d1 < 0 ? -d1 : d1); fs[2] += (uint32_t)(d2 < 0 ? -d2 : d2);
            fs[3] += (uint32_t)(d3 < 0 ? -d3 : d3); fs[4] += (uint32_t)(d4 < 0 ? -d4 : d4);
        }
        // one exchange for the five sums (the overflow flags ride in bit 63 of each)
#pragma unroll
        for (int o = 0; o < 5; o++) {
            unsigned long long t = fs[o] | ((fbig >> o) & 1u ? 1ull << 63 : 0ull);
            const unsigned long long flag = __reduce_or_sync(FULL, (uint32_t)(t >> 63)) ? 1ull << 63 : 0ull;
            t &= ~(1ull << 63);
#pragma unroll
            for (int sft = 16; sft; sft >>= 1) t += __shfl_xor_sync(FULL, t, sft);
            if ((tid & 31) == 0) sh.red5[o][tid >> 5] = t | flag;
        }
        __syncthreads();
        uint32_t fo = 0xffffffffu; unsigned long long fbest = ~0ull;
#pragma unroll
        for (int o = 0; o < 5; o++) {
            unsigned long long t = 0, flag = 0;
#pragma unroll
            for (int wv = 0; wv < NW; wv++) { const unsigned long long q = sh.red5[o][wv]; t += q & ~(1ull << 63); flag |= q >> 63; }
            if ((uint32_t)o < bs && !flag && t < fbest) { fbest = t; fo = (uint32_t)o; }
        }
        // ---- LPC analysis (before r[] is used: wi aliases it)
        uint32_t maxo = min(a.max_lpc, (uint32_t)MAX_LPC);
        if (maxo >= bs) maxo = bs - 1;
        if (tid == 0) sh.lpc_ok = 0;
        if (maxo) {
            // Welch window in Q15 on the integer samples, autocorrelation in exact 64-bit integer arithmetic (25-bit samples: a product
            // has 48 bits, 16384 of them 62): a warp takes a PAIR of lags (one load of sample i serves both), its lanes stride over the
            // samples, one shuffle reduction per lag.  (FP32 running sums, libFLAC's own choice, were measured first: 2 instructions per
            // lag-sample like this, but tonal signals lost up to 20 % of their compression to the rounding of the sums.)
            const float half = 0.5f * (float)(bs - 1), inv = 1.0f / (half + 1.0f);
            for (uint32_t i = tid; i < bs; i += NT) {
                const float t = ((float)i - half) * inv;
                const int32_t wq = (int32_t)(32767.0f * (1.0f - t * t));
                wi[i] = (int32_t)(((long long)x[i] * wq) >> 15);
            }
            __syncthreads();
            for (uint32_t l = 2 * (tid >> 5); l <= maxo; l += 2 * NW) {
                const uint32_t lane = tid & 31;
                long long sa = 0, sb = 0;
                for (uint32_t i = l + 1 + lane; i < bs; i += 32) { const long long c = wi[i]; sa += c * wi[i - l]; sb += c * wi[i - l - 1]; }
                if (lane == 0) sa += (long long)wi[l] * wi[0];
#pragma unroll
                for (int o = 16; o; o >>= 1) { sa += __shfl_xor_sync(FULL, sa, o); sb += __shfl_xor_sync(FULL, sb, o); }
                if (lane == 0) { sh.ac[l] = (double)sa; if (l + 1 <= maxo) sh.ac[l + 1] = (double)sb; }
            }
            __syncthreads();
            if (tid == 0 && sh.ac[0] > 0.0) {
                // Levinson-Durbin; err[o] = prediction error energy of order o
                double e = sh.ac[0], aa[MAX_LPC];
                sh.err[0] = e;
                for (uint32_t i = 0; i < maxo; i++) {
                    double rr = -sh.ac[i + 1];
                    for (uint32_t j = 0; j < i; j++) rr -= aa[j] * sh.ac[i - j];
                    rr /= e;
                    aa[i] = rr;
                    for (uint32_t j = 0; j < i / 2; j++) { const double t = aa[j]; aa[j] += rr * aa[i - 1 - j]; aa[i - 1 - j] += rr * t; }
                    if (i & 1) aa[i / 2] += aa[i / 2] * rr;
                    e *= (1.0 - rr * rr);
                    if (!(e > 0.0)) e = 1e-9;
                    for (uint32_t j = 0; j <= i; j++) sh.lpcs[(i + 1) * i / 2 + j] = -aa[j];
                    sh.err[i + 1] = e;
                }
                sh.lpc_ok = 2;          // analysis done, order not chosen yet
            }
            __syncthreads();
            const uint32_t prec_v = min(15u, max(5u, a.prec ? a.prec : (ebps > 16 ? (bs > 1152 ? 15u : 14u) : (bs > 4608 ? 13u : 12u))));
            if (sh.lpc_ok == 2 && a.search_order && tid >= 1 && tid <= maxo) {      // libFLAC's estimate: bits per residual sample from the error energy, one order per thread
                const double ee = sh.err[tid] * (0.5 * 0.4804530139182014 / (double)bs);
                double bpr = ee > 0 ? 0.5 * log2(ee) : 0.0;
                if (bpr < 0) bpr = 0;
                sh.est[tid] = bpr * (double)(bs - tid) + (double)tid * (double)(ebps + prec_v);
            }
            __syncthreads();
            if (tid == 0 && sh.lpc_ok == 2) {
                sh.lpc_ok = 0;
                const uint32_t prec = prec_v;
                uint32_t lo = maxo;
                if (a.search_order) {
                    double bestest = 1e300;
                    for (uint32_t o = 1; o <= maxo; o++) if (sh.est[o] < bestest) { bestest = sh.est[o]; lo = o; }
                }
                // quantise with error feedback
                double cmax = 0;
                for (uint32_t i = 0; i < lo; i++) { const double m = fabs(sh.lpcs[lo * (lo - 1) / 2 + i]); if (m > cmax) cmax = m; }
                if (cmax > 0) {
                    int l2; (void)frexp(cmax, &l2); l2--;
                    int shift = (int)prec - l2 - 2;
                    if (shift > 15) shift = 15;
                    if (shift >= 0) {
                        const int32_t qmax = (1 << (prec - 1)) - 1, qmin = -qmax - 1;
                        double ef = 0;
                        for (uint32_t i = 0; i < lo; i++) {
                            ef += sh.lpcs[lo * (lo - 1) / 2 + i] * (double)(1 << shift);
                            long long vq = llrint(ef);
                            if (vq > qmax) vq = qmax;
                            if (vq < qmin) vq = qmin;
                            ef -= (double)vq; sh.qc[i] = (int32_t)vq;
                        }
                        sh.lpc_order = lo; sh.lpc_prec = prec; sh.lpc_shift = (uint32_t)shift; sh.lpc_ok = 1;
                    }
                }
            }
            __syncthreads();
        }
        // ---- FIXED candidate
        uint32_t best_bits = verb_bits;
        if (tid == 0) d->type = 1;
        if (fo != 0xffffffffu) {
            for (uint32_t i = tid; i < bs; i += NT) {
                if (i < fo) continue;
                int32_t p = 0;
                switch (fo) {
                case 1: p = x[i - 1]; break;
                case 2: p = 2 * x[i - 1] - x[i - 2]; break;
                case 3: p = 3 * x[i - 1] - 3 * x[i - 2] + x[i - 3]; break;
                case 4: p = 4 * x[i - 1] - 6 * x[i - 2] + 4 * x[i - 3] - x[i - 4]; break;
                default: break;
                }
                r[ridx(i)] = x[i] - p;
            }
            __syncthreads();
            const uint32_t rb = plan_rice<NT>(r, bs, fo, a.min_po, a.max_po, sh, d);
            const uint32_t fb = 8 + w + fo * ebps + rb;
            if (fb < best_bits) { best_bits = fb; if (tid == 0) { d->type = 2; d->order = (uint8_t)fo; } }
        }
        __syncthreads();
        // ---- LPC candidate
        if (sh.lpc_ok) {
            const uint32_t lo = sh.lpc_order, prec = sh.lpc_prec, shift = sh.lpc_shift;
            const bool narrow = ebps + prec + (31u - (uint32_t)__clz(lo)) <= 32u;     // libFLAC 1.2.1's width rule: what the decoder will do
            uint32_t big = lpc_residuals_any<NT>(x, r, bs, lo, sh.qc, shift, narrow);
            big = block_or_u32<NT>(big, sh.scratch);
            if (!big) {
                EncSub* t = &sh.tmp;
                const uint32_t rb = plan_rice<NT>(r, bs, lo, a.min_po, a.max_po, sh, t);
                const uint32_t lb = 8 + w + lo * ebps + 9 + lo * prec + rb;
                if (lb < best_bits) {
                    best_bits = lb;
                    for (uint32_t q = tid; q < (1u << MAX_PO); q += NT) d->k[q] = t->k[q];
                    if (tid < lo) d->qc[tid] = (int16_t)sh.qc[tid];
                    if (tid == 0) { d->type = 3; d->order = (uint8_t)lo; d->prec = (uint8_t)prec; d->shift = (uint8_t)shift; d->po = t->po; d->method = t->method; }
                }
            }
        }
        __syncthreads();
        if (tid == 0) d->bits = best_bits;
        __syncthreads();
    }
    // ---- stereo decision, frame layout
    __shared__ uint32_t s_pick[8];
    if (tid == 0) {
        uint32_t assign = a.ch - 1;
        if (a.stereo) {
            const uint32_t bL = sh.dec[0].bits, bR = sh.dec[1].bits, bM = sh.dec[2].bits, bS = sh.dec[3].bits;
            uint32_t bb = bL + bR; assign = 1; s_pick[0] = 0; s_pick[1] = 1;
            if (bL + bS < bb) { bb = bL + bS; assign = 8; s_pick[0] = 0; s_pick[1] = 3; }
            if (bS + bR < bb) { bb = bS + bR; assign = 9; s_pick[0] = 3; s_pick[1] = 1; }
            if (bM + bS < bb) { bb = bM + bS; assign = 10; s_pick[0] = 2; s_pick[1] = 3; }
        } else for (uint32_t c = 0; c < a.ch; c++) s_pick[c] = c;
        EncFrame* F = &a.frm[f];
        const uint32_t hl = make_frame_header(sh.hdr, a, bs, assign, a.first_number + f);
        uint32_t bit = hl * 8;
        for (uint32_t c = 0; c < 8; c++) { F->sub_bit[c] = bit; if (c < a.ch) bit += sh.dec[s_pick[c]].bits; }
        F->byte_off = 0; F->nbytes = (bit + 7) / 8 + 2; F->bs = bs; F->assignment = (uint8_t)assign; F->hdr_len = (uint8_t)hl;
        for (uint32_t i = 0; i < 16; i++) F->hdr[i] = i < hl ? sh.hdr[i] : 0;
    }
    __syncthreads();
    for (uint32_t c = 0; c < a.ch; c++) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&sh.dec[s_pick[c]]);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&a.sub[(uint64_t)f * 8 + c]);
        for (uint32_t q = tid; q < sizeof(EncSub) / 4; q += NT) dst[q] = src[q];
    }
}

// ------------------------------------------------------------------------------------------------ scan kernel (one CTA)
__global__ void __launch_bounds__(1024) k_enc_scan(EncArgs a) {
    __shared__ unsigned long long s_sum[1024];
    __shared__ uint32_t s_min[32], s_max[32];
    const uint32_t tid = threadIdx.x, n = a.nframes;
    const uint32_t per = (n + 1023) / 1024;
    const uint32_t b = min(n, tid * per), e = min(n, b + per);
    unsigned long long s = 0; uint32_t mn = 0xffffffffu, mx = 0;
    for (uint32_t i = b; i < e; i++) { const uint32_t nb = a.frm[i].nbytes; s += nb; mn = min(mn, nb); mx = max(mx, nb); }
    s_sum[tid] = s;
    mn = __reduce_min_sync(FULL, mn); mx = __reduce_max_sync(FULL, mx);
    if ((tid & 31) == 0) { s_min[tid >> 5] = mn; s_max[tid >> 5] = mx; }
    __syncthreads();
    if (tid == 0) {
        unsigned long long run = 0;
        for (uint32_t i = 0; i < 1024; i++) { const unsigned long long t = s_sum[i]; s_sum[i] = run; run += t; }
        uint32_t m0 = 0xffffffffu, m1 = 0;
        for (uint32_t i = 0; i < 32; i++) { m0 = min(m0, s_min[i]); m1 = max(m1, s_max[i]); }
        a.totals->total_bytes = a.first_frame + run; a.totals->min_fs = n ? m0 : 0; a.totals->max_fs = m1;
    }
    __syncthreads();
    unsigned long long off = a.first_frame + s_sum[tid];
    for (uint32_t i = b; i < e; i++) { a.frm[i].byte_off = off; off += a.frm[i].nbytes; }
}

// ------------------------------------------------------------------------------------------------ write kernel
// MSB-first bit writer over a zeroed buffer of 32-bit words (stored big-endian).  A run of bits owned by one thread: whole words
// are stored, the first and the last word of the run (shared with the neighbouring runs) are ORed in atomically.
struct BitW {
    uint32_t* base; unsigned long long pos; unsigned long long acc; uint32_t nacc; bool partial;
    __device__ __forceinline__ void init(uint8_t* out, unsigned long long bitpos) {
        base = reinterpret_cast<uint32_t*>(out); pos = bitpos; acc = 0; nacc = (uint32_t)(bitpos & 31); partial = nacc != 0;
    }
    __device__ __forceinline__ void put(uint32_t v, uint32_t n) {        // n <= 32, v < 2^n
        if (n == 0) return;
        acc = (acc << n) | v; nacc += n; pos += n;
        if (nacc >= 32) {
            const uint32_t w = (uint32_t)(acc >> (nacc - 32));
            uint32_t* p = base + ((pos - nacc) >> 5);
            const uint32_t be = __byte_perm(w, 0, 0x0123);
            if (partial) { atomicOr(p, be); partial = false; } else *p = be;
            nacc -= 32;
            acc &= (1ull << nacc) - 1ull;
        }
    }
    __device__ __forceinline__ void zeros(uint32_t n) { while (n >= 32) { put(0, 32); n -= 32; } put(0, n); }
    __device__ __forceinline__ void finish() {
        if (nacc) { const uint32_t w = (uint32_t)(acc << (32 - nacc)); atomicOr(base + ((pos - nacc) >> 5), __byte_perm(w, 0, 0x0123)); nacc = 0; }
    }
};

__device__ __forceinline__ uint32_t gf_mulmod(uint32_t x, uint32_t y) {      // in GF(2)[x] / (x^16 + x^15 + x^2 + 1)
    uint32_t r = 0;
#pragma unroll
    for (int i = 15; i >= 0; i--) { r <<= 1; if (r & 0x10000u) r ^= 0x18005u; if ((y >> i) & 1u) r ^= x; }
    return r;
}
__device__ __forceinline__ uint32_t gf_xpow8(uint32_t nbytes) {              // x^(8 * nbytes) mod P
    uint32_t res = 1, b = 0x100, e = nbytes;
    while (e) { if (e & 1u) res = gf_mulmod(res, b); b = gf_mulmod(b, b); e >>= 1; }
    return res;
}

template <int NT>
__global__ void __launch_bounds__(NT) k_enc_write(EncArgs a) {
    constexpr int NW = NT / 32;
    ENC_DYN_SMEM(dyn);
    __shared__ EncSub D;
    __shared__ int32_t qc32[MAX_LPC];
    __shared__ unsigned long long scratch[NW_MAX];
    __shared__ uint16_t crctab[256];
    const uint32_t tid = threadIdx.x, f = blockIdx.x;
    const EncFrame F = a.frm[f];
    const uint32_t bs = F.bs;
    int32_t* x = reinterpret_cast<int32_t*>(dyn) + XPAD;
    if (threadIdx.x < XPAD) x[(int)threadIdx.x - (int)XPAD] = 0;
    int32_t* r = x + a.bs;
    const uint64_t base = (uint64_t)f * a.bs * a.ch;
    const unsigned long long fbit = F.byte_off * 8ull;
    if (tid < 256) {
        uint32_t c = tid << 8;
        for (int k = 0; k < 8; k++) c = (c & 0x8000u) ? ((c << 1) ^ 0x8005u) & 0xffffu : (c << 1) & 0xffffu;
        crctab[tid] = (uint16_t)c;
    }
    if (tid == 0) {
        BitW bw; bw.init(a.out, fbit);
        for (uint32_t i = 0; i < F.hdr_len; i++) bw.put(F.hdr[i], 8);
        bw.finish();
    }
    const uint32_t nper = (bs + NT - 1) / NT;
    for (uint32_t c = 0; c < a.ch; c++) {
        __syncthreads();
        {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(&a.sub[(uint64_t)f * 8 + c]);
            uint32_t* dst = reinterpret_cast<uint32_t*>(&D);
            for (uint32_t q = tid; q < sizeof(EncSub) / 4; q += NT) dst[q] = src[q];
        }
        __syncthreads();
        if (tid < MAX_LPC) qc32[tid] = D.qc[tid];          // visible after the barrier that follows the sample load
        const uint32_t type = D.type, order = D.order, w = D.wasted, variant = D.variant;
        const uint32_t vbps = a.bps + ((a.stereo && variant == 3) ? 1u : 0u), ebps = vbps - w;
        const uint32_t emask = ebps >= 32 ? 0xffffffffu : (1u << ebps) - 1u;
        const unsigned long long sbit = fbit + F.sub_bit[c];
        if (type != 0) {
            uint32_t i = tid;
            for (; i + 3 * NT < bs; i += 4 * NT) {
                int32_t sv[4];
#pragma unroll
                for (int q = 0; q < 4; q++) sv[q] = load_variant(a, base, i + q * NT, variant);
#pragma unroll
                for (int q = 0; q < 4; q++) x[i + q * NT] = sv[q] >> w;
            }
            for (; i < bs; i += NT) x[i] = load_variant(a, base, i, variant) >> w;
        }
        __syncthreads();
        uint32_t head_bits = 8 + w;                      // bits before the per-sample part, all written by thread 0
        if (type >= 2) head_bits += order * ebps + (type == 3 ? 9u + order * D.prec : 0u) + 6u;
        if (type == 0) head_bits += ebps;
        if (tid == 0) {
            BitW bw; bw.init(a.out, sbit);
            const uint32_t code = type == 0 ? 0u : type == 1 ? 1u : type == 2 ? 8u + order : 32u + order - 1u;
            bw.put(code << 1 | (w ? 1u : 0u), 8);
            if (w) bw.put(1, w);                          // unary: w - 1 zeros, then a one
            if (type == 0) bw.put((uint32_t)D.cval & emask, ebps);
            if (type >= 2) {
                for (uint32_t i = 0; i < order; i++) bw.put((uint32_t)x[i] & emask, ebps);
                if (type == 3) {
                    bw.put(D.prec - 1u, 4); bw.put(D.shift & 31u, 5);
                    const uint32_t pm = (1u << D.prec) - 1u;
                    for (uint32_t i = 0; i < order; i++) bw.put((uint32_t)(int32_t)D.qc[i] & pm, D.prec);
                }
                bw.put(D.method, 2); bw.put(D.po, 4);
            }
            bw.finish();
        }
        if (type == 1) {
            const uint32_t ib = min(bs, tid * nper), ie = min(bs, (tid + 1) * nper);
            if (ib < ie) {
                BitW bw; bw.init(a.out, sbit + head_bits + (unsigned long long)ib * ebps);
                for (uint32_t i = ib; i < ie; i++) bw.put((uint32_t)x[i] & emask, ebps);
                bw.finish();
            }
        } else if (type >= 2) {
            // residuals
            if (type == 2) {
                for (uint32_t i = tid; i < bs; i += NT) {
                    if (i < order) continue;
                    int32_t p = 0;
                    switch (order) {
                    case 1: p = x[i - 1]; break;
                    case 2: p = 2 * x[i - 1] - x[i - 2]; break;
                    case 3: p = 3 * x[i - 1] - 3 * x[i - 2] + x[i - 3]; break;
                    case 4: p = 4 * x[i - 1] - 6 * x[i - 2] + 4 * x[i - 3] - x[i - 4]; break;
                    default: break;
                    }
                    r[ridx(i)] = x[i] - p;
                }
            } else {
                const uint32_t prec = D.prec, shift = D.shift;
                const bool narrow = ebps + prec + (31u - (uint32_t)__clz(order)) <= 32u;
                (void)lpc_residuals_any<NT>(x, r, bs, order, qc32, shift, narrow);
            }
            __syncthreads();
            const uint32_t po = D.po, psz = bs >> po, plen = D.method ? 5u : 4u;
            const uint32_t ib = min(bs, max(order, tid * nper)), ie = min(bs, (tid + 1) * nper);
            uint32_t mybits = 0;
            for (uint32_t i = ib; i < ie; i++) {
                const uint32_t q = i / psz;
                if (i == order || i == q * psz) mybits += plen;
                const uint32_t k = D.k[q];
                mybits += (zigzag(r[ridx(i)]) >> k) + 1u + k;
            }
            uint32_t total;
            const uint32_t start = block_excl_scan_u32<NT>(mybits, scratch, &total);
            if (ib < ie) {
                BitW bw; bw.init(a.out, sbit + head_bits + start);
                for (uint32_t i = ib; i < ie; i++) {
                    const uint32_t q = i / psz, k = D.k[q];
                    if (i == order || i == q * psz) bw.put(k, plen);
                    const uint32_t u = zigzag(r[ridx(i)]), msb = u >> k;
                    const uint32_t low = (1u << k) | (u & ((1u << k) - 1u));          // stop bit + k low bits
                    if (msb + k + 1u <= 32u) bw.put(low, msb + k + 1u);
                    else { bw.zeros(msb); bw.put(low, k + 1u); }
                }
                bw.finish();
            }
        }
    }
    // ---- CRC-16 of the frame (everything before the last two bytes), then the footer
    __threadfence();
    __syncthreads();
    const uint32_t nb = F.nbytes - 2;
    const uint32_t per = (nb + NT - 1) / NT;
    const uint32_t b0 = min(nb, tid * per), b1 = min(nb, (tid + 1) * per);
    uint32_t crc = 0;
    const uint8_t* fp = a.out + F.byte_off;
    for (uint32_t i = b0; i < b1; i++) crc = ((crc << 8) & 0xffffu) ^ crctab[(crc >> 8) ^ __ldcg(fp + i)];
    if (b0 < b1 && nb - b1) crc = gf_mulmod(crc, gf_xpow8(nb - b1));
    if (b0 >= b1) crc = 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) crc ^= __shfl_xor_sync(FULL, crc, o);
    __syncthreads();
    if ((tid & 31) == 0) scratch[tid >> 5] = crc;
    __syncthreads();
    if (tid == 0) {
        uint32_t cc = 0;
        for (int wv = 0; wv < NW; wv++) cc ^= (uint32_t)scratch[wv];
        BitW bw; bw.init(a.out, fbit + (unsigned long long)nb * 8ull);
        bw.put(cc, 16);
        bw.finish();
    }
}

} // namespace bnfe
