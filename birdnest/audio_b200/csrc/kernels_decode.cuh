// kernels_decode.cuh -- K3-5: the decode kernel and its launcher templates.  Instantiated by kernels_decode_wide.cu (64-bit
// accumulation) and kernels_decode_narrow.cu (32-bit) so that the two halves of the variants compile in parallel.
#pragma once
#include "kernels_common.cuh"

// build knobs of the kernel experiments (tools/build_variants.sh); the defaults are what was measured fastest
#ifndef DEC_PACK_UNIFORM
#define DEC_PACK_UNIFORM 0
#endif
#ifndef DEC_I2D_CVT
#define DEC_I2D_CVT 1
#endif
#ifndef DEC_DEPHASE_NS
#define DEC_DEPHASE_NS 0
#endif
// DEC_LANE_MAJOR: sample t of lane l at word l * (T + 4) + t (a lane's samples are contiguous) instead of t * S + l: the Rice phase
// stores and the restore phase loads and stores four samples per instruction (conflict free: T + 4 = 4 mod 8, so the 16-byte
// accesses of a quarter warp fall into eight different groups of four banks); the pack phase reads 16-byte runs per channel
#ifndef DEC_LANE_MAJOR
#define DEC_LANE_MAJOR 1
#endif
#ifndef DEC_CKPT16
#define DEC_CKPT16 1
#endif
#ifndef DEC_ZZ_BFE
#define DEC_ZZ_BFE 1
#endif

namespace bnf {

// ------------------------------------------------------------------------------------------------ K3-5 decode
// One lane per (frame, channel); a warp (= one CTA) owns 32/C frames and works tile by tile (T samples per channel):
//   Rice phase     each lane decodes its next T residuals into its own column of the shared-memory tile.  Groups of 8
//                  codewords take a branch-free path (window, bfind, shift, one IMAD, zig-zag; an overflow flag instead
//                  of a branch) whenever every lane has 8 codewords left in its partition; warm-up samples, partition
//                  tails, escape partitions and VERBATIM subframes take the careful per-sample path.
//   restore phase  each lane runs the FIXED/LPC recurrence over its column in place.  Coefficients and the last ORD
//                  samples stay in registers; the loop is unrolled ORD times so every tap has a fixed register.
//                  16-bit streams accumulate in 32-bit IMADs.  Streams that need libFLAC's 64-bit accumulator use the FP64
//                  pipe instead: every product and partial sum is an integer below 2^53, so DFMA is exact, the
//                  quantisation shift is folded into the coefficients (a power of two), floor() is one round-down add of
//                  1.5*2^52, and the FP64 pipe runs beside the integer pipes the Rice phase of the other warps keeps busy.
//   pack phase     the warp re-reads the tile in interleaved order (a lane per 16 consecutive output samples of one frame), applies
//                  left/side, side/right, mid/side decorrelation to stereo pairs (branch free) and writes packed little-endian
//                  8/16/24-bit PCM with 16-byte stores.
// Tile layout (DEC_LANE_MAJOR, the default): sample t of lane l at word l*(T+4) + t -- a lane's samples are contiguous, so the Rice
// phase stores and the restore phase loads and stores FOUR samples per instruction and the pack phase reads 16-byte runs per channel;
// T + 4 = 4 (mod 8) keeps the 16-byte accesses of a quarter warp in eight different groups of four banks.  (The column-major layout
// it replaced, sample t of lane l at word t*S + l with S = 32 + pad, is still selectable: one 4-byte access per sample.)
// A launch is either plain (one job = 32/C frames per warp) or balanced (template parameter BAL, see k_decode).
enum : int { M_IDLE = 0, M_CONST = 1, M_VERBATIM = 2, M_PRED = 3 };

#ifndef DEC_MAXNREG
#define DEC_MAXNREG 120
#endif
// orders > 16: 168 registers = three warps per scheduler, one wave for the few-large-frames streams these variants exist for
// (ptxas: 108 bytes of spill stores for ORD = 32, none for 16)
#ifndef DEC_MAXNREG_BIG
#define DEC_MAXNREG_BIG 168
#endif
#ifndef DEC_TILE
#define DEC_TILE 48
#endif

// orders > 16: one ring period per tile (T = 32 against 64: cfg3 decode 4.32 -> 4.13 ms, the phases of the three warps a scheduler
// holds interleave more finely)
#ifndef DEC_TILE_BIG
#define DEC_TILE_BIG 32
#endif
// FT: rows of the per-warp frame table (bs | assign | PCM offset); formats with two or more channels hold at most 16 frames per warp
template <int ORD, int SPEC = 0> struct DecCfg {
    static constexpr int T = (ORD > 16) ? DEC_TILE_BIG : DEC_TILE;
    static constexpr int FT = (SPEC && (SPEC >> 2) >= 2) ? 16 : 32;
    static_assert(T % 8 == 0 && T % ORD == 0 && (T * (SPEC ? (SPEC >> 2) : 1)) % 16 == 0, "a tile is whole Rice groups, whole restore blocks and whole pack units");
};

struct RiceSt {
    uint32_t fastleft, rawleft, rawbits, k, kp32, negP, c30, psize, plen, order;
    bool first;
};

template <class BR>
__device__ __forceinline__ void rice_param(BR& br, RiceSt& rs) {
#pragma unroll 1
    for (int guard = 0; guard < 2; guard++) {
        const uint32_t cnt = rs.psize - (rs.first ? rs.order : 0);
        rs.first = false;
        const uint32_t k = br.get(rs.plen);
        if (k == (rs.plen == 5 ? 31u : 15u)) { rs.rawbits = br.get(5); rs.rawleft = cnt; rs.fastleft = 0; }
        else { rs.fastleft = cnt; rs.rawleft = 0; rs.k = k; rs.kp32 = k + 32u; rs.negP = 0u - (1u << k); rs.c30 = 30u << k; }
        if (cnt) break;
    }
}

// ---- restore: one block of ORD samples of this lane's column, in place
// `h` is a ring: before a block that starts at sample t0, h[j] holds sample t0 - ORD + j; step j reads every tap from a
// fixed register and then overwrites h[j] (whose old value, the oldest sample, was used for the last time in that step).
template <int ORD, bool FIRST, bool EXTRA>
__device__ __forceinline__ void restore_block_i32(uint32_t addr, uint32_t rs4, const int32_t (&cf)[ORD], int32_t (&h)[ORD],
                                                  uint32_t order, uint32_t shift, uint32_t wasted) {
    uint4 v4 = make_uint4(0, 0, 0, 0);      // lane-major tile: four residuals in, four samples out per access
#pragma unroll
    for (int j = 0; j < ORD; j++) {
        if (DEC_LANE_MAJOR && j % 4 == 0) v4 = lds128(addr + 4 * j);
        const int32_t r = DEC_LANE_MAJOR ? (int32_t)get4(v4, j % 4) : (int32_t)lds32(addr + j * rs4);
        uint32_t sum = 0;
#pragma unroll
        for (int m = ORD - 1; m >= 0; m--) sum += (uint32_t)cf[m] * (uint32_t)h[(j - 1 - m + 2 * ORD) % ORD];
        int32_t s = (int32_t)((uint32_t)r + (uint32_t)((int32_t)sum >> shift));
        if (FIRST) { if (j < (int)order) s = r; }
        h[j] = s;
        {
            const uint32_t o = EXTRA ? (uint32_t)s << wasted : (uint32_t)s;
            if (DEC_LANE_MAJOR) { set4(v4, j % 4, o); if (j % 4 == 3) sts128(addr + 4 * (j - 3), v4); }
            else sts32(addr + j * rs4, o);
        }
    }
}

template <int ORD, bool FIRST, bool EXTRA>
__device__ __forceinline__ void restore_block_f64(uint32_t addr, uint32_t rs4, const double (&cf)[ORD], double (&h)[ORD],
                                                  uint32_t order, uint32_t sh_n, uint32_t wasted) {
    uint4 v4 = make_uint4(0, 0, 0, 0);      // lane-major tile: four residuals in, four samples out per access
#pragma unroll
    for (int j = 0; j < ORD; j++) {
        if (DEC_LANE_MAJOR && j % 4 == 0) v4 = lds128(addr + 4 * j);
        const int32_t r = DEC_LANE_MAJOR ? (int32_t)get4(v4, j % 4) : (int32_t)lds32(addr + j * rs4);
        // NACC independent accumulation chains (every product and partial sum is an integer, scaled by 2^-shift, below 2^53: any
        // association is exact).  Orders > 16 run as a few warps per scheduler, where one chain of ORD dependent DFMAs (8.7 cycles
        // each, tools/ulat.cu) is what the warp waits for: 4 chains, cfg3 decode 4.38 -> 4.30 ms.  For orders <= 12 ptxas already
        // keeps three samples in flight; 2 or 3 chains there were measured slower (cfg2 decode 1.90 -> 1.96 / 1.98 ms).
        constexpr int NACC = ORD > 16 ? 4 : ORD > 12 ? 2 : 1;
        double acc[NACC];
#pragma unroll
        for (int q = 0; q < NACC; q++) acc[q] = 0.0;
#pragma unroll
        for (int m = ORD - 1; m >= 0; m--) acc[m % NACC] = fma(cf[m], h[(j - 1 - m + 2 * ORD) % ORD], acc[m % NACC]);
#pragma unroll
        for (int w = NACC; w > 1; w /= 2)
#pragma unroll
            for (int q = 0; q < w / 2; q++) acc[q] += acc[q + w / 2];
        const double acc0 = acc[0];
        const double y = __dadd_rd(acc0, 6755399441055744.0);            // + 1.5*2^52, rounded down: low word = floor(acc) mod 2^32
        int32_t p = __double2loint(y);
        if (EXTRA) p >>= sh_n;
        int32_t s = (int32_t)((uint32_t)r + (uint32_t)p);
        if (FIRST) { if (j < (int)order) s = r; }
#if DEC_I2D_CVT
        h[j] = (double)s;                                                            // I2F.F64.S32: one instruction instead of three
#else
        h[j] = __hiloint2double(0x43300000, s ^ 0x80000000) - 4503601774854144.0;   // (double)s, exact: 2^52 + 2^31 bias
#endif
        {
            const uint32_t o = EXTRA ? (uint32_t)s << wasted : (uint32_t)s;
            if (DEC_LANE_MAJOR) { set4(v4, j % 4, o); if (j % 4 == 3) sts128(addr + 4 * (j - 3), v4); }
            else sts32(addr + j * rs4, o);
        }
    }
}

// 64-bit integer accumulation (mad.wide.s32): int32 coefficients and history, half the registers of the FP64 form.
// `wasted` carries the per-lane "narrow" flag in bit 31 (EXTRA only): narrow LPC subframes wrap at 32 bits before the shift.
template <int ORD, bool FIRST, bool EXTRA>
__device__ __forceinline__ void restore_block_i64(uint32_t addr, uint32_t rs4, const int32_t (&cf)[ORD], int32_t (&h)[ORD],
                                                  uint32_t order, uint32_t shift, uint32_t wasted) {
    uint4 v4 = make_uint4(0, 0, 0, 0);      // lane-major tile: four residuals in, four samples out per access
#pragma unroll
    for (int j = 0; j < ORD; j++) {
        if (DEC_LANE_MAJOR && j % 4 == 0) v4 = lds128(addr + 4 * j);
        const int32_t r = DEC_LANE_MAJOR ? (int32_t)get4(v4, j % 4) : (int32_t)lds32(addr + j * rs4);
        long long acc = 0;
#pragma unroll
        for (int m = ORD - 1; m >= 0; m--) asm("mad.wide.s32 %0, %1, %2, %0;" : "+l"(acc) : "r"(cf[m]), "r"(h[(j - 1 - m + 2 * ORD) % ORD]));
        const uint32_t lo = (uint32_t)acc, hi = (uint32_t)((unsigned long long)acc >> 32);
        int32_t p = (int32_t)__funnelshift_r(lo, hi, shift);               // shift < 32
        if (EXTRA) { if (wasted & 0x80000000u) p = (int32_t)lo >> shift; }
        int32_t s = (int32_t)((uint32_t)r + (uint32_t)p);
        if (FIRST) { if (j < (int)order) s = r; }
        h[j] = s;
        {
            const uint32_t o = EXTRA ? (uint32_t)s << (wasted & 31u) : (uint32_t)s;
            if (DEC_LANE_MAJOR) { set4(v4, j % 4, o); if (j % 4 == 3) sts128(addr + 4 * (j - 3), v4); }
            else sts32(addr + j * rs4, o);
        }
    }
}

#ifndef DEC_WIDE_I64
#define DEC_WIDE_I64 0
#endif

// Orders > 16 (streams of few, long subframes: a scheduler holds two or three warps, each of which would issue its dot product to ONE
// pipe): the 16 most recent taps accumulate on the FP64 pipe (DFMA, history as doubles), the 16 older ones on the integer multiply
// pipe (mad.wide.s32, history as int32) -- a warp that alternates two pipes issues 0.55 instructions per clock where one pipe gives
// it 0.29 (tools/ulat.cu).  Both sums are exact integers: the DFMA chain starts at 1.5 * 2^52, so the bits of its result are
// 0x4338000000000000 + sum, and the integer chain starts at -0x4338000000000000: one 64-bit add gives the whole dot product, the
// quantisation shift is a funnel shift.  `wasted` carries the per-lane "narrow" flag in bit 31 as in restore_block_i64.
// h: the last 32 samples (int32, slot = sample index mod 32); hD: the last 16 as doubles (slot = index mod 16).
// MEASURED AND NOT USED (bit-exact, cfg3 decode 4.12 -> 4.86 ms): mad.wide.s32 with a 64-bit accumulator costs the warp more issue
// time than the DFMA it replaces, and the 64-bit recombination adds six instructions per sample.  Kept as a build variant (-DDEC_MIX=1).
#ifndef DEC_MIX
#define DEC_MIX 0
#endif
template <int ORD, bool WIDE> __host__ __device__ constexpr bool dec_mix() { return WIDE && ORD > 16 && DEC_MIX && !DEC_WIDE_I64; }
template <int ORD, bool FIRST, bool EXTRA>
__device__ __forceinline__ void restore_block_mix(uint32_t addr, uint32_t rs4, const int32_t (&cf)[ORD], int32_t (&h)[ORD], const double (&cfD)[16], double (&hD)[16],
                                                  uint32_t order, uint32_t shift, uint32_t wasted) {
    static_assert(ORD == 32, "two halves of 16 taps");
    uint4 v4 = make_uint4(0, 0, 0, 0);      // lane-major tile: four residuals in, four samples out per access
#pragma unroll
    for (int j = 0; j < ORD; j++) {
        if (DEC_LANE_MAJOR && j % 4 == 0) v4 = lds128(addr + 4 * j);
        const int32_t r = DEC_LANE_MAJOR ? (int32_t)get4(v4, j % 4) : (int32_t)lds32(addr + j * rs4);
        long long ai0 = -0x4338000000000000ll, ai1 = 0;
#pragma unroll
        for (int m = ORD - 1; m >= 16; m--) {
            if (m & 1) asm("mad.wide.s32 %0, %1, %2, %0;" : "+l"(ai1) : "r"(cf[m]), "r"(h[(j - 1 - m + 2 * ORD) % ORD]));
            else asm("mad.wide.s32 %0, %1, %2, %0;" : "+l"(ai0) : "r"(cf[m]), "r"(h[(j - 1 - m + 2 * ORD) % ORD]));
        }
        double ad[4] = {6755399441055744.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int m = 15; m >= 0; m--) ad[m & 3] = fma(cfD[m], hD[(j - 1 - m + 32) % 16], ad[m & 3]);
        const double y = (ad[0] + ad[1]) + (ad[2] + ad[3]);
        const long long tot = __double_as_longlong(y) + ai0 + ai1;
        const uint32_t lo = (uint32_t)tot, hi = (uint32_t)((unsigned long long)tot >> 32);
        int32_t p = (int32_t)__funnelshift_r(lo, hi, shift);               // shift < 32
        if (EXTRA) { if (wasted & 0x80000000u) p = (int32_t)lo >> shift; }
        int32_t s = (int32_t)((uint32_t)r + (uint32_t)p);
        if (FIRST) { if (j < (int)order) s = r; }
        h[j] = s;
        hD[j % 16] = __hiloint2double(0x43300000, s ^ 0x80000000) - 4503601774854144.0;   // (double)s, exact: 2^52 + 2^31 bias
        {
            const uint32_t o = EXTRA ? (uint32_t)s << (wasted & 31u) : (uint32_t)s;
            if (DEC_LANE_MAJOR) { set4(v4, j % 4, o); if (j % 4 == 3) sts128(addr + 4 * (j - 3), v4); }
            else sts32(addr + j * rs4, o);
        }
    }
}
template <int ORD, bool WIDE, bool FIRST, bool EXTRA, class TT>
__device__ __forceinline__ void restore_block(uint32_t addr, uint32_t rs4, const TT (&cf)[ORD], TT (&h)[ORD], const double (&cfD)[16], double (&hD)[16], uint32_t order, uint32_t shift, uint32_t wasted) {
    if constexpr (dec_mix<ORD, WIDE>()) restore_block_mix<ORD, FIRST, EXTRA>(addr, rs4, cf, h, cfD, hD, order, shift, wasted);
    else if constexpr (WIDE && DEC_WIDE_I64) restore_block_i64<ORD, FIRST, EXTRA>(addr, rs4, cf, h, order, shift, wasted);
    else if constexpr (WIDE) restore_block_f64<ORD, FIRST, EXTRA>(addr, rs4, cf, h, order, shift, wasted);
    else restore_block_i32<ORD, FIRST, EXTRA>(addr, rs4, cf, h, order, shift, wasted);
}

// ---- pack: 16 consecutive samples (interleaved order) of one frame per lane -> B 16-byte stores
__device__ __forceinline__ void pack4(uint32_t* dw, uint32_t B, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3) {
    if (B == 3) { dw[0] = __byte_perm(v0, v1, 0x4210); dw[1] = __byte_perm(v1, v2, 0x5421); dw[2] = __byte_perm(v2, v3, 0x6542); }
    else if (B == 2) { dw[0] = __byte_perm(v0, v1, 0x5410); dw[1] = __byte_perm(v2, v3, 0x5410); }
    else dw[0] = __byte_perm(__byte_perm(v0, v1, 0x0040), __byte_perm(v2, v3, 0x0040), 0x5410);
}
// stereo decorrelation of one (ch0, ch1) pair, branch free: the lanes of a pack step hold different frames, so a
// branch per assignment diverges on streams with adaptive stereo (ncu: the branchy form was 10 instructions per pair and 13 % of
// the kernel's stall samples).  With d = x - (y >> sh), sh = 1 for mid/side and 0 otherwise:
//   left/side  (8): ch1 = x - y = d,            ch0 = x = d + y
//   mid/side  (10): ch1 = M - (S >> 1) = d,     ch0 = ch1 + S = d + y      (m' = 2M + (S&1): L = (m'+S)>>1 = R + S, R = (m'-S)>>1;
//                                                                            identical in every bit that reaches the output, also when int32 wraps)
//   side/right (9): ch0 = x + y,                ch1 = y
// i.e. ch0 = (useD ? d : x) + (y & addm), ch1 = useD ? d : y.
struct DecorrSel { uint32_t sh, addm; bool useD; };
__device__ __forceinline__ DecorrSel decorr_sel(uint32_t assign) {
    DecorrSel k; k.sh = assign == 10 ? 1u : 0u; k.addm = assign >= 8 ? ~0u : 0u; k.useD = assign == 8 || assign == 10;
    return k;
}
__device__ __forceinline__ void decorr(const DecorrSel& k, uint32_t& x, uint32_t& y) {
    const uint32_t d = x - (uint32_t)((int32_t)y >> k.sh);
    const uint32_t b = k.useD ? d : x;
    x = b + (y & k.addm);
    y = k.useD ? d : y;
}

__device__ __forceinline__ void pack_tile(uint32_t tile_base, uint32_t S, uint32_t C, uint32_t B, uint32_t F, uint32_t i0, uint32_t T,
                                          uint32_t ftab, uint32_t FT, uint8_t* __restrict__ out, uint32_t lane) {
    const uint32_t upf = (T * C) >> 4;                 // units of 16 samples per frame-tile (T is a multiple of 16)
    const uint32_t total = F * upf;
    const uint32_t rcp_upf = 65536u / upf + 1u;        // g / upf for g < 2^9
    const uint32_t rcp_c = 65536u / C + 1u;
    for (uint32_t g = lane; g < total; g += 32) {
        const uint32_t f = (g * rcp_upf) >> 16, u = g - f * upf;
        const uint32_t bs = lds32(ftab + 4 * f);
        const uint32_t nt = bs > i0 ? min(T, bs - i0) : 0u, nsamp = nt * C, q0 = 16 * u;
        if (q0 >= nsamp) continue;
        const uint32_t assign = lds32(ftab + 4 * FT + 4 * f);
        const uint2 pol = lds64(ftab + 8 * FT + 8 * f);
        uint8_t* dst = out + ((((uint64_t)pol.y << 32) | pol.x) + (i0 * C + q0) * B);      // 32-bit offset inside the frame: blocksize * channels * bytes < 2^22
        uint32_t v[16];
        const uint32_t LS4 = 4 * S;                                     // lane-major tile: bytes between the channels (lanes) of a frame
        const uint32_t fbase = DEC_LANE_MAJOR ? tile_base + f * C * LS4 : tile_base + 4 * f * C;
        if (C == 2) {
            if (DEC_LANE_MAJOR) {                                      // 8 time steps: two 16-byte runs per channel
                const uint32_t ad = fbase + 4 * (q0 >> 1);
                const uint4 a0 = lds128(ad), a1 = lds128(ad + 16), b0 = lds128(ad + LS4), b1 = lds128(ad + LS4 + 16);
                v[0] = a0.x; v[2] = a0.y; v[4] = a0.z; v[6] = a0.w; v[8] = a1.x; v[10] = a1.y; v[12] = a1.z; v[14] = a1.w;
                v[1] = b0.x; v[3] = b0.y; v[5] = b0.z; v[7] = b0.w; v[9] = b1.x; v[11] = b1.y; v[13] = b1.z; v[15] = b1.w;
            } else {
                const uint32_t ad = fbase + 4 * (q0 >> 1) * S;
#pragma unroll
                for (int e = 0; e < 8; e++) { const uint2 p = lds64(ad + 4 * e * S); v[2 * e] = p.x; v[2 * e + 1] = p.y; }
            }
            const DecorrSel ds = decorr_sel(assign);
#if DEC_PACK_UNIFORM
            if (__all_sync(__activemask(), ds.useD)) {       // every unit of this step is left/side or mid/side (the usual choices of an adaptive encoder): 3 instead of 6 per pair
#pragma unroll
                for (int e = 0; e < 8; e++) { const uint32_t d = v[2 * e] - (uint32_t)((int32_t)v[2 * e + 1] >> ds.sh); v[2 * e] = d + v[2 * e + 1]; v[2 * e + 1] = d; }
            } else
#endif
            {
#pragma unroll
                for (int e = 0; e < 8; e++) decorr(ds, v[2 * e], v[2 * e + 1]);
            }
        } else if (DEC_LANE_MAJOR) {
            if (C == 1) {                                              // 16 time steps of one lane
#pragma unroll
                for (int e = 0; e < 4; e++) { const uint4 p = lds128(fbase + 4 * q0 + 16 * e); v[4 * e] = p.x; v[4 * e + 1] = p.y; v[4 * e + 2] = p.z; v[4 * e + 3] = p.w; }
            } else if (C == 4) {                                       // 4 time steps x 4 channels
#pragma unroll
                for (int c = 0; c < 4; c++) { const uint4 p = lds128(fbase + c * LS4 + 4 * (q0 >> 2)); v[c] = p.x; v[4 + c] = p.y; v[8 + c] = p.z; v[12 + c] = p.w; }
            } else if (C == 8) {                                       // 2 time steps x 8 channels
#pragma unroll
                for (int c = 0; c < 8; c++) { const uint2 p = lds64(fbase + c * LS4 + 4 * (q0 >> 3)); v[c] = p.x; v[8 + c] = p.y; }
            } else {
                uint32_t t = (q0 * rcp_c) >> 16, c = q0 - t * C;
                uint32_t ad = fbase + c * LS4 + 4 * t;
                const uint32_t wrap = (C - 1) * LS4 - 4;                // from the last channel of step t back to channel 0 of step t + 1
#pragma unroll
                for (int e = 0; e < 16; e++) { v[e] = lds32(ad); ad += LS4; if (++c == C) { c = 0; ad -= wrap + LS4; } }
            }
        } else if (C == 1) {
#pragma unroll
            for (int e = 0; e < 16; e++) v[e] = lds32(fbase + 4 * (q0 + e) * S);
        } else if (C == 4) {
#pragma unroll
            for (int e = 0; e < 4; e++) { const uint4 p = lds128(fbase + 4 * ((q0 >> 2) + e) * S); v[4 * e] = p.x; v[4 * e + 1] = p.y; v[4 * e + 2] = p.z; v[4 * e + 3] = p.w; }
        } else if (C == 8) {
#pragma unroll
            for (int e = 0; e < 4; e++) { const uint4 p = lds128(fbase + 4 * ((q0 >> 3) + (e >> 1)) * S + 16 * (e & 1)); v[4 * e] = p.x; v[4 * e + 1] = p.y; v[4 * e + 2] = p.z; v[4 * e + 3] = p.w; }
        } else {
            uint32_t t = (q0 * rcp_c) >> 16, c = q0 - t * C;
            uint32_t ad = fbase + 4 * (t * S + c);
            const uint32_t wrap = 4 * (S - C);
#pragma unroll
            for (int e = 0; e < 16; e++) { v[e] = lds32(ad); ad += 4; if (++c == C) { c = 0; ad += wrap; } }
        }
        if (q0 + 16 <= nsamp && (((uintptr_t)dst) & 15u) == 0) {
            uint4* d4 = reinterpret_cast<uint4*>(dst);
            if (B == 3) {
#pragma unroll
                for (int e = 0; e < 3; e++) {       // 16 samples x 3 bytes = 12 words: word j holds bytes 4j..4j+3
                    uint32_t w[4];
#pragma unroll
                    for (int x = 0; x < 4; x++) {
                        const int j = 4 * e + x, q = (4 * j) / 3, r = (4 * j) % 3;
                        w[x] = r == 0 ? __byte_perm(v[q], v[q + 1 > 15 ? 15 : q + 1], 0x4210) : r == 1 ? __byte_perm(v[q], v[q + 1 > 15 ? 15 : q + 1], 0x5421) : __byte_perm(v[q], v[q + 1 > 15 ? 15 : q + 1], 0x6542);
                    }
                    d4[e] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            } else if (B == 2) {
#pragma unroll
                for (int e = 0; e < 2; e++)
                    d4[e] = make_uint4(__byte_perm(v[8 * e], v[8 * e + 1], 0x5410), __byte_perm(v[8 * e + 2], v[8 * e + 3], 0x5410),
                                       __byte_perm(v[8 * e + 4], v[8 * e + 5], 0x5410), __byte_perm(v[8 * e + 6], v[8 * e + 7], 0x5410));
            } else {
                uint32_t w[4];
#pragma unroll
                for (int x = 0; x < 4; x++) w[x] = __byte_perm(__byte_perm(v[4 * x], v[4 * x + 1], 0x0040), __byte_perm(v[4 * x + 2], v[4 * x + 3], 0x0040), 0x5410);
                d4[0] = make_uint4(w[0], w[1], w[2], w[3]);
            }
        } else {
            const uint32_t nv = min(16u, nsamp - q0);
#pragma unroll
            for (uint32_t e = 0; e < 16; e += 4) {
                if (e + 4 <= nv && (((uintptr_t)dst) & 3u) == 0) pack4(reinterpret_cast<uint32_t*>(dst + e * B), B, v[e], v[e + 1], v[e + 2], v[e + 3]);
                else {
#pragma unroll
                    for (uint32_t x = 0; x < 4; x++)
                        if (e + x < nv) for (uint32_t b = 0; b < B; b++) dst[(e + x) * B + b] = (uint8_t)(v[e + x] >> (8 * b));
                }
            }
        }
    }
}

#ifndef DEC_WARPS_N
#define DEC_WARPS_N 2
#endif
constexpr int DEC_WARPS = DEC_WARPS_N;        // independent warps per CTA (no CTA-wide barrier anywhere)
// S: column-major tile: words between consecutive samples of a lane (32 + pad); lane-major tile: words between lanes (T + 4)
__host__ __device__ constexpr uint32_t dec_tile_words(int T, uint32_t S) { return DEC_LANE_MAJOR ? 32u * S : (uint32_t)T * S; }
__host__ __device__ constexpr uint32_t dec_warp_smem(int T, uint32_t S, int FT) { return 32u * RingBits::STRIDE + dec_tile_words(T, S) * 4u + 16u * (uint32_t)FT; }

// SPEC: 0 = any channel count / sample width (run-time C, B, S); else 4 C + B: the common formats get C, B and the tile
// stride S as compile-time constants (tile addresses become immediates, lane -> (frame, channel) is a shift, the pack
// phase loses its format dispatch): measured 2.10 -> 1.93 ms on the 1 h 24-bit stereo stream
#ifndef DEC_SPECIALISE
#define DEC_SPECIALISE 1
#endif
__host__ __device__ constexpr uint32_t dec_tile_stride(uint32_t C, int T) { return DEC_LANE_MAJOR ? (uint32_t)T + 4u : 32u + ((C & 3u) == 0 ? 4u : (C & 1u) == 0 ? 2u : 1u); }
// Balanced schedule (epoch != 0): the launch is ONE resident wave of warps ("slots") and the job list -- job j = the 32/C frames a
// warp decodes side by side, NT nominal tiles each -- is cut into equal runs of tiles, one run per slot, wherever the cut falls.
// A plain launch of J jobs on M slots takes ceil(J / M) subframe chains however small the last wave is (cfg2: 2.23 waves cost 3,
// the 16-bit hour 1.02 waves cost 2); cut this way every slot is busy for J / M chains.  A job cut in two is handed over through
// HBM: the slot that decodes its first tiles does so FIRST (slots walk their run from its last job to its first), saves per lane
// the bit position, the Rice partition state and the last ORD samples, and raises flag[slot]; the slot that owns the rest of the job
// reaches it LAST, re-reads the subframe header (coefficients, shift) and resumes from the saved state.  The producer is the
// lower-numbered slot and never waits itself, so CTAs dispatched in index order cannot deadlock even when they are not all resident.
// the cut schedule exists for the format-specialised variants of orders <= 16 (streams of many frames; few-large-frame streams are
// a single wave or many full ones): keeps the number of kernels to compile in bounds
template <int ORD, int SPEC> constexpr bool dec_has_balanced() { return ORD <= 16 && SPEC != 0; }
inline uint32_t dec_force_slots() { static const uint32_t v = getenv("BNFLAC_BALANCE_SLOTS") ? (uint32_t)atoi(getenv("BNFLAC_BALANCE_SLOTS")) : 0u; return v; }
constexpr uint32_t DEC_STATE_WORDS = 8 + 32;               // per lane: position (2), fastleft, rawleft, rawbits, k | first, 2 spare, history
constexpr uint32_t DEC_SLOT_WORDS = DEC_STATE_WORDS * 32;  // word w of lane l at [w * 32 + l]
// BAL = false is the plain launch (one job per warp: the job loop below runs once and the hand-over code is not compiled); the cut
// schedule is a second instantiation because the loop-carried state costs the plain kernel 2 - 4 % (cfg2 decode 1.72 -> 1.79 ms).
template <int ORD, bool WIDE, int SPEC, bool BAL>
__global__ void __launch_bounds__(32 * DEC_WARPS) __maxnreg__(ORD <= 12 ? DEC_MAXNREG : ORD <= 16 ? 128 : DEC_MAXNREG_BIG) k_decode(PassArgs a, uint32_t C_, uint32_t B_, uint32_t S_, uint32_t epoch) {
    constexpr int T = DecCfg<ORD, SPEC>::T, FT = DecCfg<ORD, SPEC>::FT;
    const uint32_t C = SPEC ? (uint32_t)(SPEC >> 2) : C_, B = SPEC ? (uint32_t)(SPEC & 3) : B_, S = SPEC ? dec_tile_stride(SPEC >> 2, T) : S_;
    extern __shared__ __align__(128) uint8_t s_dyn[];
    const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // through a shuffle: ptxas otherwise rematerialises this address from SR_CgaCtaId / SR_TID in every pack step (two S2R + four more)
    const uint32_t ring_base = __shfl_sync(FULL, smem_u32(s_dyn) + wib * dec_warp_smem(T, S, FT), 0);
    const uint32_t tile_base = ring_base + 32 * RingBits::STRIDE;
    const uint32_t ftab = tile_base + dec_tile_words(T, S) * 4;   // bs[FT] | assign[FT] | pcm offset[FT] (u64)
    const uint32_t F = 32 / C;
    const uint32_t n_acc = a.totals->n_accepted;
    const uint32_t fl = lane / C, ch = lane - fl * C;
    const uint32_t rs4 = DEC_LANE_MAJOR ? 4u : S * 4;                                   // bytes between consecutive samples of a lane
    const uint32_t col = tile_base + (DEC_LANE_MAJOR ? lane * S * 4 : lane * 4);      // this lane's sample 0
    const uint32_t slot = blockIdx.x * DEC_WARPS + wib;
#if DEC_DEPHASE_NS
    // Warps launched together do identical work and stay in step: all the warps of a scheduler are in the restore phase (FP64 pipe)
    // or in the Rice phase (integer pipe) at the same time and queue for one pipe while the others idle.  A pseudo-random start
    // delay of up to one tile period spreads the phases for the rest of the launch.
    __nanosleep(((slot * 2654435761u) >> 19) % (uint32_t)DEC_DEPHASE_NS);
#endif
    // this warp's run: jobs j_lo .. j_hi, tiles [tb, ..) of j_lo and [.., te) of j_hi
    uint32_t j_lo = slot, j_hi = slot, tb = 0, te = 0xffffu;
    if constexpr (BAL) {
        const uint32_t NT = max(1u, (a.totals->max_bs + T - 1) / T), J = (n_acc + F - 1) / F, M = gridDim.x * DEC_WARPS;
        const uint32_t W = J * NT, q = (W + M - 1) / M;               // the host takes this path only when J * NT < 2^31
        const uint32_t g0 = min(W, slot * q), g1 = min(W, g0 + q);
        if (g0 >= g1) return;
        j_lo = g0 / NT; tb = g0 - j_lo * NT;
        j_hi = (g1 - 1) / NT; te = g1 - j_hi * NT;
    }
    uint32_t job = j_hi;
#pragma unroll 1
    for (;;) {
    const uint32_t t_begin = job == j_lo ? tb : 0u, t_end = job == j_hi ? te : 0xffffu;
    const uint32_t kf = job * F + fl;
    const bool active = fl < F && kf < n_acc;

    // ---- per-subframe state (registers)
    DecRing<ORD> br;
    br.init_idle(ring_base + lane * RingBits::STRIDE, a.in);
    constexpr bool MIX = dec_mix<ORD, WIDE>();         // orders > 16: half the taps on the FP64 pipe, half on the integer pipe (restore_block_mix)
    constexpr bool F64 = WIDE && !DEC_WIDE_I64 && !MIX;       // FP64-pipe accumulation (coefficients scaled by 2^-shift) vs mad.wide.s32
    typename std::conditional<F64, double, int32_t>::type cf[ORD], hist[ORD];
#pragma unroll
    for (int j = 0; j < ORD; j++) { cf[j] = 0; hist[j] = 0; }
    RiceSt rs;
    rs.fastleft = 0; rs.rawleft = 0; rs.rawbits = 0; rs.k = 0; rs.kp32 = 32; rs.negP = ~0u; rs.c30 = 30; rs.psize = 0; rs.plen = 4; rs.order = 0; rs.first = true;
    uint32_t bs = 0, assign = 0, wasted = 0, shift = 0, bps = 0;
    int mode = M_IDLE;
    if (fl < F && ch == 0) { sts32(ftab + 4 * fl, 0); sts32(ftab + 4 * FT + 4 * fl, 0); }
    if (active) {
        const uint32_t i = (a.acc_sorted ? a.acc_sorted : a.acc_idx)[kf];
        const Cand c = a.cand[i];
        bs = c.bs; assign = c.assign;
        bool ok = a.status[i] == ST_OK;
        const uint64_t po = a.pcm_off[i];
        if (po + (uint64_t)bs * C * B > a.out_cap) { ok = false; bs = 0; }   // never write past the caller's buffer
        if (ch == 0) { sts32(ftab + 4 * fl, bs); sts32(ftab + 4 * FT + 4 * fl, assign); sts32(ftab + 8 * FT + 8 * fl, (uint32_t)po); sts32(ftab + 8 * FT + 4 + 8 * fl, (uint32_t)(po >> 32)); }
        int32_t cval = 0;
        mode = M_CONST;                              // damaged frames (CRC mismatch) are delivered zero-filled
        if (ok) {
            const SubInfo si = a.sub[(uint64_t)i * MAX_CH + ch];
            br.init(ring_base + lane * RingBits::STRIDE, a.in, a.in_len, c.off * 8 + si.bit_offset);
            uint32_t x = br.get(8);
            if (x & 1) { br.unary(64); br.ensure_now(); }
            const uint32_t order = si.order;
            wasted = si.wasted;
            bps = (uint32_t)c.bps + (((assign == 8 && ch == 1) || (assign == 9 && ch == 0) || (assign == 10 && ch == 1)) ? 1u : 0u) - wasted;
            if (si.type == 0) cval = br.gets(bps);
            else if (si.type == 1) mode = M_VERBATIM;
            else {
                mode = M_PRED;
                rs.order = order;
                // warm-up samples are parked in this lane's tile rows 0..order-1 (order <= 32 <= T)
#pragma unroll 1
                for (uint32_t j = 0; j < order; j++) { sts32(col + j * rs4, (uint32_t)br.gets(bps)); if ((j & 7) == 7) br.ensure_now(); }
                br.ensure_now();
                bool narrow = true;
                if (si.type == 3) {
                    const uint32_t prec = br.get(4) + 1;
                    { const int32_t sh = br.gets(5); shift = sh < 0 ? 0u : (uint32_t)sh; }   // negative: not an error in libFLAC 1.2.1 (never emitted)
                    narrow = (bps + prec + (uint32_t)ilog2u(order)) <= 32;
                    // libFLAC 1.2.1 width rule: narrow subframes accumulate in 32 bits (wrap), the others in 64
                    const double scale = (F64 && !narrow) ? __hiloint2double((int)((1023u - shift) << 20), 0) : 1.0;
#pragma unroll
                    for (int j = 0; j < ORD; j++) if (j < (int)order) {
                        const int32_t q = br.gets(prec);
                        if constexpr (F64) cf[j] = (double)q * scale; else cf[j] = q;
                        if ((j & 7) == 7) br.ensure_now();
                    }
                    br.ensure_now();
                    if (F64 && !narrow) shift = 0;           // folded into the coefficients
                    if (WIDE && !F64 && narrow && shift) wasted |= 0x80000000u;    // 32-bit wrap before the shift (restore_block_i64)
                } else {   // FIXED predictors as coefficient sets, 32-bit wrap-around arithmetic
                    const int o = (int)order;
                    if (ORD >= 1 && o >= 1) cf[0] = o == 1 ? 1 : o == 2 ? 2 : o == 3 ? 3 : 4;
                    if (ORD >= 2 && o >= 2) cf[1] = o == 2 ? -1 : o == 3 ? -3 : -6;
                    if (ORD >= 3 && o >= 3) cf[2] = o == 3 ? 1 : 4;
                    if (ORD >= 4 && o >= 4) cf[3] = -1;
                }
                const uint32_t method = br.get(2);
                rs.plen = method ? 5 : 4;
                const uint32_t porder = br.get(4);
                rs.psize = porder ? bs >> porder : bs;
                br.ensure_now();
            }
        }
        if (mode == M_CONST) {                        // column holds the final value once and for all (restore leaves it unchanged)
            const uint32_t v = (uint32_t)cval << wasted;
            wasted = 0;
#pragma unroll 1
            for (uint32_t t = 0; t < (uint32_t)T; t++) sts32(col + t * rs4, v);        // (once per subframe)
        }
    }
    double cfD[16], hD[16];                             // MIX only: the 16 most recent taps as doubles (dead code otherwise)
    if constexpr (MIX) {
#pragma unroll
        for (int j = 0; j < 16; j++) { cfD[j] = (double)cf[j]; hD[j] = 0.0; }
    }
    __syncwarp();
    const uint32_t maxbs = __reduce_max_sync(FULL, bs);
    const bool reads = mode >= M_VERBATIM;
    const bool extra = __any_sync(FULL, wasted != 0 || (F64 && shift != 0));
    const uint32_t order = rs.order;
    const uint32_t i0_begin = t_begin * T, i0_end = min(maxbs, t_end * (uint32_t)T);
    if (BAL && i0_begin && i0_begin < maxbs) {      // the rest of a job another slot began: wait for its state
        const uint32_t* st = a.dec_state + (uint64_t)(slot - 1) * DEC_SLOT_WORDS + lane;
        const uint32_t* flag = a.dec_flags + (slot - 1);
        while (ld_acquire_u32(flag) != epoch) __nanosleep(200);
        if (reads) {
            const uint64_t abs_bit = (uint64_t)__ldcg(st) | ((uint64_t)__ldcg(st + 32) << 32);
            br.init(ring_base + lane * RingBits::STRIDE, a.in, a.in_len, abs_bit);
            rs.fastleft = __ldcg(st + 64); rs.rawleft = __ldcg(st + 96); rs.rawbits = __ldcg(st + 128);
            const uint32_t kw = __ldcg(st + 160);
            rs.first = (kw >> 31) != 0;
            const uint32_t k = kw & 31u;
            rs.k = k; rs.kp32 = k + 32u; rs.negP = 0u - (1u << k); rs.c30 = 30u << k;
#pragma unroll
            for (int j = 0; j < ORD; j++) { const int32_t v = (int32_t)__ldcg(st + 256 + 32 * j); hist[j] = v; if constexpr (MIX) hD[j % 16] = (double)v; }
        }
        __syncwarp();
    }
#pragma unroll 1
    for (uint32_t i0 = i0_begin; i0 < i0_end; i0 += T) {
        // ---- Rice phase
#pragma unroll 1
        for (uint32_t t0 = 0, row = col; t0 < (uint32_t)T; t0 += 8, row += 8 * rs4) {      // row: loop-carried, or it is rematerialised from SR_TID every step
            const uint32_t idx0 = i0 + t0;
#if DEC_CKPT16
            if (!(idx0 & 8u)) { if (reads) br.ckpt_full(); }
            else if (__any_sync(FULL, reads && br.ckpt_mid_needed())) br.ckpt_mid_wait();
#else
            if (reads && (DEC_RING_BLOCKS < 16 || !(t0 & 8u))) br.checkpoint();
#endif
            const bool inert = mode <= M_CONST || idx0 >= bs;
            if (mode == M_PRED && !inert && rs.fastleft == 0 && rs.rawleft == 0 && idx0 >= order) rice_param(br, rs);
            const bool fast_ok = inert || (mode == M_PRED && rs.fastleft >= 8);
            // branch-free group of 8 Rice codewords; `commit`: this lane is really in a Rice partition with >= 8 left
            auto rice_group = [&](const bool commit) {
                uint32_t pos = br.pos;
                const uint32_t k = rs.k, kp32 = rs.kp32, negP = rs.negP, c30 = rs.c30;
                bool ovf = false;
                int32_t r[8];
                typename DecRing<ORD>::Win3 wn = br.win_init(pos);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t nxt = j < 7 ? br.win_next(pos) : 0u;
                    const uint32_t w = DecRing<ORD>::win_peek(wn, pos);
                    const uint32_t f = bfind_fast(w);
                    const uint32_t d = f - k;
                    ovf |= (int32_t)d < 0;
                    const uint32_t np = pos + kp32 - f;
                    if (j < 7) DecRing<ORD>::win_advance(wn, pos, np, nxt);
                    pos = np;
                    const uint32_t u = f * negP + shr_c(w, d) + c30;      // (31-f) << k | low bits, stop bit cancelled
#if DEC_ZZ_BFE
                    { int32_t m; asm("bfe.s32 %0, %1, 0, 1;" : "=r"(m) : "r"(u)); r[j] = (int32_t)(u >> 1) ^ m; }      // -(u & 1) as one sign-extending field extract
#else
                    r[j] = (int32_t)(u >> 1) ^ -(int32_t)(u & 1);
#endif
                }
                if (commit) {
                    rs.fastleft -= 8;
                    if (!ovf) {
                        br.pos = pos;
#pragma unroll
                        if (DEC_LANE_MAJOR) { sts128(row, make_uint4(r[0], r[1], r[2], r[3])); sts128(row + 16, make_uint4(r[4], r[5], r[6], r[7])); }
                        else for (int j = 0; j < 8; j++) sts32(row + j * rs4, (uint32_t)r[j]);
                    } else {             // a codeword longer than one window (rare): redo the group carefully
#pragma unroll 1
                        for (int j = 0; j < 8; j++) sts32(row + j * rs4, (uint32_t)br.rice_careful(k));
                    }
                }
            };
            if (__all_sync(FULL, fast_ok)) rice_group(!inert);
            else {
                // lanes that read fixed-width samples for the whole step (VERBATIM subframes, escape partitions): when they are
                // the only reason the step is not uniform, the Rice lanes keep their branch-free group and these lanes read
                // their 8 samples in a short loop of their own -- instead of a careful walk of every lane
                const bool rawlane = !inert && ((mode == M_VERBATIM && idx0 + 8 <= bs) || (mode == M_PRED && rs.rawleft >= 8));
                if (__all_sync(FULL, fast_ok || rawlane)) {
                    rice_group(!inert && !rawlane);
                    if (rawlane) {
                        const uint32_t nb = mode == M_VERBATIM ? bps : rs.rawbits;
                        if (mode == M_PRED) rs.rawleft -= 8;
#pragma unroll 1
                        for (uint32_t j = 0; j < 8; j++) sts32(row + j * rs4, (uint32_t)br.gets(nb));
                    }
                } else {
#pragma unroll 1
                    for (uint32_t j = 0; j < 8; j++) {
                        const uint32_t idx = idx0 + j;
                        if (mode <= M_CONST || idx >= bs) continue;
                        int32_t v;
                        if (mode == M_VERBATIM) v = br.gets(bps);
                        else {
                            if (idx < order) continue;               // parked warm-up sample
                            if (rs.fastleft == 0 && rs.rawleft == 0) rice_param(br, rs);
                            if (rs.rawleft) { rs.rawleft--; v = br.gets(rs.rawbits); }
                            else { rs.fastleft--; v = br.rice_careful(rs.k); }
                        }
                        sts32(row + j * rs4, (uint32_t)v);
                    }
                }
            }
        }
        // ---- restore phase (own column only: no warp synchronisation needed before it)
        if (i0 == 0) {
            if (extra) restore_block<ORD, WIDE, true, true>(col, rs4, cf, hist, cfD, hD, order, shift, wasted);
            else restore_block<ORD, WIDE, true, false>(col, rs4, cf, hist, cfD, hD, order, shift, wasted);
        }
        if (extra) {
#pragma unroll 1
            for (uint32_t t = (i0 == 0 ? ORD : 0); t < (uint32_t)T; t += ORD) restore_block<ORD, WIDE, false, true>(col + t * rs4, rs4, cf, hist, cfD, hD, order, shift, wasted);
        } else {
#pragma unroll 1
            for (uint32_t t = (i0 == 0 ? ORD : 0); t < (uint32_t)T; t += ORD) restore_block<ORD, WIDE, false, false>(col + t * rs4, rs4, cf, hist, cfD, hD, order, shift, wasted);
        }
        __syncwarp();
        // ---- pack phase
        pack_tile(tile_base, S, C, B, F, i0, T, ftab, FT, a.out, lane);
        __syncwarp();
    }
    if (BAL && i0_end < maxbs) {                     // the job goes on in the next slot: hand over
        uint32_t* st = a.dec_state + (uint64_t)slot * DEC_SLOT_WORDS + lane;
        if (reads) {
            const uint64_t abs_bit = br.abs_pos(a.in);
            __stcg(st, (uint32_t)abs_bit); __stcg(st + 32, (uint32_t)(abs_bit >> 32));
            __stcg(st + 64, rs.fastleft); __stcg(st + 96, rs.rawleft); __stcg(st + 128, rs.rawbits); __stcg(st + 160, rs.k | (rs.first ? 0x80000000u : 0u));
#pragma unroll
            for (int j = 0; j < ORD; j++) __stcg(st + 256 + 32 * j, (uint32_t)(int32_t)hist[j]);
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) st_release_u32(a.dec_flags + slot, epoch);
    }
    if (!BAL || job == j_lo) break;
    --job;
    __syncwarp();
    }
}

template <int ORD, bool WIDE, int SPEC>
static void launch_decode_s(const PassArgs& a, uint32_t nacc, uint32_t C, uint32_t B, cudaStream_t st);
template <int ORD, bool WIDE>
static void launch_decode_t(const PassArgs& a, uint32_t nacc, uint32_t C, uint32_t B, cudaStream_t st) {
    const uint32_t key = DEC_SPECIALISE ? 4 * C + B : 0;
#ifdef DEC_EXPERIMENT      // kernel experiments (tools/build_variants.sh): only the variants the cfg2 / cfg3 benchmark streams use are compiled (40 s instead of 3 min)
    if constexpr (WIDE && ORD == 12) { if (key == 11) launch_decode_s<12, true, 11>(a, nacc, C, B, st); }
    else if constexpr (WIDE && ORD == 32) { if (key == 35) launch_decode_s<32, true, 35>(a, nacc, C, B, st); }
    else if constexpr (WIDE && ORD == 8) { if (key == 10) launch_decode_s<8, true, 10>(a, nacc, C, B, st); }
    return;
#else
    switch (key) {
#if DEC_SPECIALISE
    case 4 * 1 + 2: launch_decode_s<ORD, WIDE, 4 * 1 + 2>(a, nacc, C, B, st); break;     // mono 16-bit
    case 4 * 2 + 2: launch_decode_s<ORD, WIDE, 4 * 2 + 2>(a, nacc, C, B, st); break;     // stereo 16-bit
    case 4 * 2 + 3: launch_decode_s<ORD, WIDE, 4 * 2 + 3>(a, nacc, C, B, st); break;     // stereo 24-bit
    case 4 * 6 + 3: launch_decode_s<ORD, WIDE, 4 * 6 + 3>(a, nacc, C, B, st); break;     // 5.1 24-bit
    case 4 * 8 + 3: launch_decode_s<ORD, WIDE, 4 * 8 + 3>(a, nacc, C, B, st); break;     // 7.1 24-bit
#endif
    default: launch_decode_s<ORD, WIDE, 0>(a, nacc, C, B, st); break;
    }
#endif
}
template <int ORD, bool WIDE, int SPEC>
static void launch_decode_s(const PassArgs& a, uint32_t nacc, uint32_t C, uint32_t B, cudaStream_t st) {
    constexpr int T = DecCfg<ORD, SPEC>::T, FT = DecCfg<ORD, SPEC>::FT;
    const uint32_t F = 32 / C;
    const uint32_t S = dec_tile_stride(C, T);
    const uint32_t grid = blocks_for(nacc, F * DEC_WARPS);
    size_t smem = (size_t)DEC_WARPS * dec_warp_smem(T, S, FT);
    static std::atomic<uint64_t> attr_done{0};
    if (first_use_on_device(attr_done)) { cudaFuncSetAttribute(k_decode<ORD, WIDE, SPEC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); used_on_device(attr_done); }
    const int n_sm = sm_count();
    static const bool trace = getenv("BNFLAC_TRACE") != nullptr;
    static const bool balance = getenv("BNFLAC_DEC_BALANCE") && getenv("BNFLAC_DEC_BALANCE")[0] == '1';
    if (trace || balance) {
        int max_resident = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_resident, k_decode<ORD, WIDE, SPEC, false>, 32 * DEC_WARPS, smem);
        if (max_resident < 1) max_resident = 1;
        if (trace) fprintf(stderr, "[bnflac] k_decode<%d,%d,%d>: %d CTAs of %d warps resident per SM, grid %u (%.2f waves), %zu B smem/CTA\n", ORD, (int)WIDE, SPEC, max_resident, DEC_WARPS,
                           grid, (double)grid / ((double)n_sm * max_resident), smem);
        // Every warp runs for about the same time (one frame per lane), so the launch proceeds in waves.  BNFLAC_DEC_BALANCE=1 caps
        // the residency (by asking for more shared memory) so that the waves are equally full (measured: no gain, off by default).
        const uint64_t per_wave = (uint64_t)n_sm * max_resident;
        const uint32_t waves = (uint32_t)((grid + per_wave - 1) / per_wave);
        uint32_t resident = (uint32_t)((grid + (uint64_t)n_sm * waves - 1) / ((uint64_t)n_sm * waves));
        if (resident < 1) resident = 1;
        if (balance && resident < (uint32_t)max_resident) {
            size_t want = ((size_t)227 * 1024 / resident - 1024) & ~(size_t)127;
            if (want > 200 * 1024) want = 200 * 1024;
            if (want > smem) smem = want;
        }
    }
    // balanced schedule: one resident wave of slots, the job list cut into equal runs of tiles (see k_decode); only when there is
    // more than one wave of jobs -- a single wave cannot be shortened by cutting it
    if constexpr (dec_has_balanced<ORD, SPEC>()) if (a.dec_state && a.dec_flags && !a.acc_sorted) {
        static std::atomic<uint64_t> attr_done_b{0};
        if (first_use_on_device(attr_done_b)) { cudaFuncSetAttribute(k_decode<ORD, WIDE, SPEC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); used_on_device(attr_done_b); }
        static std::atomic<int> resident_per_sm{0};         // of this variant (same on every device of one kind)
        int mr = resident_per_sm.load(std::memory_order_relaxed);
        if (!mr) {
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&mr, k_decode<ORD, WIDE, SPEC, true>, 32 * DEC_WARPS, smem);
            if (mr < 1) mr = 1;
            resident_per_sm.store(mr, std::memory_order_relaxed);
        }
        uint32_t slots = (uint32_t)n_sm * (uint32_t)mr * DEC_WARPS;
        if (const uint32_t cap = dec_force_slots()) slots = std::max<uint32_t>(DEC_WARPS, std::min(slots, cap) / DEC_WARPS * DEC_WARPS);   // tests: hand-overs on small streams
        const uint32_t J = blocks_for(nacc, F);
        // A plain launch of w = J / slots waves takes floor(w) chains at full residency plus one chain of the last, partial wave; that
        // last chain is short when few warps share the SM (measured on the 16-bit hour: 0.26 ms alone against 0.73 ms at full
        // residency), so cutting only pays when the partial wave is small: cfg1, 1.02 waves, 0.99 -> 0.75 ms; cfg2, 2.23 waves, no gain.
        const uint32_t Jd = blocks_for(a.dec_expect ? std::min(a.dec_expect, nacc) : nacc, F), rem = Jd % slots;
        const bool pays = dec_force_slots() || (rem != 0 && rem * 5u <= slots);
        if (Jd > slots && pays && slots <= a.dec_slots && (uint64_t)J * (65536u / T + 1u) < (1ull << 31)) {
            const uint32_t epoch = next_decode_epoch();       // process-wide and never 0: flag words left behind by any earlier launch do not match
            if (trace) fprintf(stderr, "[bnflac] k_decode<%d,%d,%d>: balanced over %u slots, %u jobs (%.2f waves)\n", ORD, (int)WIDE, SPEC, slots, J, (double)J / slots);
            k_decode<ORD, WIDE, SPEC, true><<<slots / DEC_WARPS, 32 * DEC_WARPS, smem, st>>>(a, C, B, S, epoch);
            count_launch();
            return;
        }
    }
    k_decode<ORD, WIDE, SPEC, false><<<grid, 32 * DEC_WARPS, smem, st>>>(a, C, B, S, 0u);
    count_launch();
}
template <bool WIDE>
static void launch_decode_w(const PassArgs& a, uint32_t nacc, uint32_t C, uint32_t B, uint32_t max_order, cudaStream_t st) {
#ifdef DEC_EXPERIMENT
    if (WIDE && max_order <= 8) launch_decode_t<8, WIDE>(a, nacc, C, B, st);
    else if (max_order <= 12) launch_decode_t<12, WIDE>(a, nacc, C, B, st); else launch_decode_t<32, WIDE>(a, nacc, C, B, st);
    return;
#endif
    if (max_order <= 4) launch_decode_t<4, WIDE>(a, nacc, C, B, st);
    else if (max_order <= 8) launch_decode_t<8, WIDE>(a, nacc, C, B, st);
    else if (max_order <= 12) launch_decode_t<12, WIDE>(a, nacc, C, B, st);
    else if (max_order <= 16) launch_decode_t<16, WIDE>(a, nacc, C, B, st);
    else launch_decode_t<32, WIDE>(a, nacc, C, B, st);
}

} // namespace bnf
