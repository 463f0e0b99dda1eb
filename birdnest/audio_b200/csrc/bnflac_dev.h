// bnflac_dev.h -- device-side tables shared by the kernels (kernels.cu) and the host engine (engine.cu).
//
// Data layout in HBM for one pipeline pass (one stream, one shard of a stream, or a batch of clips):
//   in        : compressed bytes of every segment back to back, each segment start 16-byte aligned,
//               64 zero bytes of padding after the last byte (bit readers may over-read, never fault)
//   SegInfo[] : one per segment (= one FLAC stream or one shard of it)
//   Chunk[]   : scan work units (<= SCAN_CHUNK bytes of one segment), built on the host
//   Cand[]    : frame table, sorted by byte offset (K1 output)
//   pref[]    : per scan tile, 32 tile-prefix CRC residues (K1 output, 64 B per 8 KiB of input)
//   per-candidate arrays: seg_crc, next, flen, status, pcm_off, acc_idx, sub_bitoff[8]
//   out       : interleaved little-endian packed PCM, frames back to back in stream order
#pragma once
#include <stdint.h>

namespace bnf {

constexpr int SCAN_CHUNK = 8192;       // bytes per scan tile (one warp); tile k of a segment covers [max(begin, A + k*8192), A + (k+1)*8192), A = begin & ~15
constexpr int MAX_CH = 8;
// layout of PassArgs::counters (words): [0] candidates appended, [1] overflow flags, then the per-CTA totals of the two
// look-back kernels (at most 256 CTAs each)
constexpr int CNT_ANOM = 2;            // candidates off the clean frame chain (k_parse -> k_resync)
constexpr uint32_t ANOM_CAP = 8192;
constexpr int CNT_SPEC = 3;            // speculative subframe walks queued (k_spec_find -> k_parse<.., true>)
constexpr int CNT_SPEC_DONE = 4;       // frames whose guessed subframe starts all chained up (k_spec_resolve)
constexpr int CNT_ORDER = 16, CNT_PFX_CNT = CNT_ORDER + 256, CNT_PFX_BYTES = CNT_PFX_CNT + 256, CNT_WORDS = CNT_PFX_BYTES + 512;

// per-candidate status
enum : uint8_t {
    ST_OK = 0,
    ST_LOST = 1,         // (host side only) gap before this frame
    ST_CRC = 3,          // CRC-16 mismatch: frame is delivered zero-filled (reference behaviour)
    ST_UNPARSEABLE = 4,  // reserved field used: frame is not delivered, the reference reports UNPARSEABLE_STREAM
    ST_LOSTSYNC = 5,     // parse failed (pad bit, ran past any possible frame end): not delivered, the reference reports LOST_SYNC
    ST_EOS = 7,          // parse ran off the end of the stream: the reference stops decoding here (END_OF_STREAM)
    ST_SKIP = 6,         // never reached by the reference's cursor (it resumed past this candidate after a damaged frame)
    ST_CHECK = 0xFE,     // span check inconclusive: K2 validates by parsing
    ST_DROP = 0xFF       // false sync / covered by another frame
};

struct SegInfo {
    uint64_t begin, end;          // byte range of frame data inside `in`
    uint64_t own_begin, own_end;  // frames whose sync lies in [own_begin, own_end) belong to this shard
    uint32_t bps, channels, sample_rate, min_bs, max_bs;
    uint32_t max_frame_bytes;     // upper bound used when extending a CRC span over false syncs
    uint32_t first_chunk;         // index of the segment's first scan chunk
    uint32_t pad;
};

struct Chunk {
    uint64_t begin;
    uint32_t len;
    uint32_t seg;
};

struct Cand {               // 32 bytes
    uint64_t off;           // absolute byte offset of the sync code inside `in`
    uint64_t number;        // coded frame number (fixed blocksize) or sample number (variable)
    uint32_t bs;            // blocksize
    uint32_t seg;
    uint8_t hdr_len, bps, assign, flags;   // flags bit0: variable blocksize
    uint32_t sample_rate;
};

struct SubInfo {            // K2 output, 8 bytes per (frame, channel)
    uint32_t bit_offset;    // bit offset of the subframe header from the frame's first byte
    uint8_t type;           // 0 CONSTANT 1 VERBATIM 2 FIXED 3 LPC
    uint8_t order;
    uint8_t wasted;
    uint8_t flags;          // bit0: narrow (32-bit) accumulate per libFLAC's width rule; bit1: Rice2
};

// Speculative parse (streams of few, large frames): one job = one subframe walked from a GUESSED start.  The jobs of a frame
// are contiguous; job 0 of a frame is channel 0 at its true start.
struct SpecJob {            // 24 bytes
    uint32_t frame;         // candidate index
    uint32_t start_bit;     // bit offset of the (guessed) subframe header from the frame's first byte
    uint32_t end_bit;       // where the walk of this subframe ended (k_parse<.., true>); 0 = the walk failed
    SubInfo si;             // what the walk found there
    uint32_t ch;            // channel the guess is for
};
constexpr uint32_t SPEC_MAX_PER_FRAME = 48;     // guesses kept per frame (all channels)

struct Totals {             // written by the prefix-sum kernel
    uint64_t pcm_bytes;     // bytes the pass produces
    uint32_t n_accepted;    // frames delivered (OK or zero-filled)
    uint32_t n_cand;        // candidates in the table
    uint32_t overflow;      // bit0 scan-chunk overflow, bit1 candidate-table overflow
    uint32_t max_order;     // largest predictor order among accepted frames
    uint32_t any_wide;      // some LPC subframe needs 64-bit accumulation
    uint32_t max_bs;
};

struct PassArgs {
    const uint8_t* in;
    uint64_t in_len;
    const SegInfo* segs;
    uint32_t nsegs;
    const Chunk* chunks;
    uint32_t nchunks;
    // K1
    Cand* cand_tmp;
    Cand* cand;
    uint32_t cand_cap;
    uint32_t* chunk_base;
    uint32_t* chunk_count;
    uint32_t* chunk_scan;
    uint32_t* counters;     // CNT_WORDS words, see above
    uint16_t* pref;         // [tile][32] CRC-16 residue of the bytes from the start of the tile to the end of each 256-byte piece
    uint16_t* seg_crc;      // CRC-16 residue of each span between consecutive candidates (0 <=> the span is a frame whose CRC matches)
    uint32_t* next;
    uint32_t* flen;
    uint8_t* status;
    uint32_t* anom;         // ANOM_CAP candidate indices that are off the clean chain (unordered)
    // K2
    SubInfo* sub;
    SpecJob* spec_jobs;     // speculative parse: job list (spec_cap entries), per-candidate base / count / done flag
    uint32_t spec_cap;
    uint32_t* spec_base;
    uint8_t* spec_count;
    uint8_t* spec_done;     // 1: every subframe start of this (CRC-validated) frame was confirmed, k_parse has nothing to do
    // prefix
    uint64_t* pcm_off;
    uint32_t* acc_idx;
    uint32_t* acc_sorted;   // variable-blocksize passes: acc_idx regrouped by blocksize class (k_bucket_*); nullptr = stream order is kept
    uint32_t* bucket_hist;  // ... per-CTA class counts, then start positions
    Totals* totals;
    // K3-5
    uint8_t* out;
    uint64_t out_cap;
    // balanced decode schedule (kernels_decode.cuh): per slot one flag word and the state of a job handed over to the next slot;
    // nullptr = plain launch (one job per warp)
    uint32_t* dec_flags;
    uint32_t* dec_state;
    uint32_t dec_slots;     // slots the two arrays hold
    uint32_t dec_expect;    // frames the pass is expected to deliver (the previous pass's count; the launch bound is ~6 % above it): what the choice of schedule goes by
};

// launchers (kernels.cu); all asynchronous on `stream`
void launch_scan(const PassArgs& a, void* stream);
void launch_order(const PassArgs& a, void* stream);
void launch_crc(const PassArgs& a, uint32_t ncand_bound, void* stream);
void launch_link(const PassArgs& a, uint32_t ncand_bound, void* stream);
void launch_parse(const PassArgs& a, uint32_t ncand_bound, void* stream);
bool parse_wants_speculation(uint32_t ncand_bound, uint32_t channels);   // few, large frames: the engine then provides PassArgs::spec_*
void launch_resync(const PassArgs& a, void* stream);
void launch_prefix(const PassArgs& a, uint32_t ncand_bound, uint32_t bytes_per_sample, void* stream);
constexpr uint32_t BUCKET_CHUNK = 4096;  // accepted frames per CTA of the regrouping kernels
void launch_bucket(const PassArgs& a, uint32_t nacc_bound, void* stream);   // after launch_prefix; k_decode then reads a.acc_sorted
void launch_make_chunks(const SegInfo& seg, SegInfo* d_seg, Chunk* chunks, uint32_t nchunks, void* stream);   // scan tiles of a one-segment pass
void launch_clear(const PassArgs& a, void* stream);                                   // counters + totals = 0
void launch_publish(const void* src, void* dst_mapped, uint32_t nwords, void* stream);  // <= 32 words to mapped host memory
void launch_seg_summary(const PassArgs& a, uint32_t ncand_bound, uint64_t* seg_pcm, uint32_t* seg_flags, void* stream);
void launch_decode(const PassArgs& a, uint32_t nacc_bound, uint32_t channels, uint32_t bytes_per_sample, uint32_t max_order, bool wide, void* stream);
// Slots (resident decode warps) a pass of `nacc_bound` frames could be cut over, 0 when the plain launch is used anyway (at most one
// wave of jobs, or switched off): the engine then provides PassArgs::dec_flags (4 bytes per slot, zeroed once) and dec_state
// (decode_sched_state_bytes per slot).
uint32_t decode_sched_slots(uint32_t nacc_bound, uint32_t channels);
uint64_t decode_sched_state_bytes();
int kernel_launch_count();   // kernels launched so far by this process (bench "gpu_launches")

} // namespace bnf
