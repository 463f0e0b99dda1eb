/*
 * bncorpus.h -- synthetic FLAC corpus generator (test/bench infrastructure, not the product).
 *
 * The reference ships no .flac fixture (Library/OpenALDemo/01_Ghosts_I.flac is absent, see
 * /root/reference/.MISSING_LARGE_BLOBS) and no encoder is ever called by its code, so the streams
 * BASELINE.json names ("synthetic FLAC streams of the named shapes") are produced here:
 *   - deterministic integer PCM synthesis (seeded LCG noise + sinusoid sums),
 *   - a small from-scratch FLAC encoder able to emit every construct the decoder must handle
 *     (CONSTANT / VERBATIM / FIXED 0-4 / LPC 1-32, wasted bits, Rice + Rice2, escape partitions,
 *     all four stereo assignments, 1-8 channels, fixed and variable blocksize),
 *   - frame-level tiling (renumber + re-CRC) to reach 1 h / 10 h stream lengths.
 * Every stream it writes is cross-checked by the reference decoder binary (oracle/refdll) in the
 * CPU test-suite before any GPU parity claim is made on it.
 */
#ifndef BNCORPUS_H
#define BNCORPUS_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint32_t channels, bps, sample_rate;
    uint32_t blocksize;          /* nominal blocksize */
    uint32_t nvar;               /* >0: variable-blocksize stream (0xFFF9) cycling var_bs[] */
    uint32_t var_bs[16];
    uint32_t max_lpc_order;      /* 0 = FIXED predictors only */
    uint32_t qlp_precision;      /* 0 = auto (15 for bps>16, else 12..14 like libFLAC) */
    uint32_t min_part_order, max_part_order;
    uint32_t stereo_mode;        /* 0 independent, 1 adaptive (best of 4), 2 L/S, 3 S/R, 4 M/S */
    uint32_t search_order;       /* 1: pick the LPC order with the fewest estimated bits, 0: always max order */
    uint32_t escape_every;       /* >0: every n-th residual partition is written as an escape (raw) partition */
    uint32_t verbatim_every;     /* >0: every n-th subframe forced VERBATIM */
    uint32_t allow_zero_part;    /* allow partition 0 to hold zero samples ((bs>>po)==order) */
    uint32_t streaminfo_in_frames; /* 1: use "get from STREAMINFO" codes for sample rate / bps in frame headers */
    uint32_t no_md5;             /* leave STREAMINFO md5 zero */
    uint32_t padding_bytes;      /* add a PADDING block of this length (>0) and a small VORBIS_COMMENT */
} bnc_params;

typedef struct {
    uint8_t* data; size_t len;        /* whole stream */
    uint64_t* frame_off; size_t nframes; /* byte offset of every frame, plus frame_off[nframes] = len */
    uint32_t* frame_bs;               /* blocksize of every frame */
    uint64_t total_samples;           /* per channel */
    uint8_t md5[16];
    size_t first_frame;
} bnc_stream;

/* synthetic interleaved int32 PCM (deterministic in seed).  kind: 0 music-like (sinusoids + noise of
 * noise_bits, inter-channel correlation), 1 adds "special" segments (silence, full-scale noise, wasted
 * low bits, constant DC) cycling every `blocksize` samples. */
void bnc_synth(int32_t* pcm, uint64_t nsamples, uint32_t channels, uint32_t bps, uint32_t sample_rate,
               uint32_t noise_bits, uint32_t kind, uint32_t special_period, uint32_t seed);

/* encode interleaved int32 PCM.  Returns 0 on success; caller frees with bnc_free. */
int bnc_encode(const int32_t* pcm, uint64_t nsamples, const bnc_params* p, bnc_stream* out);

/* tile a fixed- or variable-blocksize stream `times` times (renumber frames, recompute CRC-8/CRC-16,
 * patch STREAMINFO total_samples/md5).  Requires every frame of `in` to be full-size unless variable.
 * pcm_md5_src: interleaved LE packed PCM of ONE tile for the md5 (may be NULL -> md5 zero). */
int bnc_tile(const bnc_stream* in, uint32_t times, const uint8_t* pcm_one_tile, size_t pcm_len, int nthreads, bnc_stream* out);

/* convert a fixed-blocksize stream to variable-blocksize framing (0xFFF9 + sample numbers) in place of numbers */
int bnc_to_variable(const bnc_stream* in, bnc_stream* out);

void bnc_free(bnc_stream* s);

/* pack interleaved int32 -> LE ceil(bps/8)-byte PCM (the MD5 layout) */
void bnc_pack_pcm(const int32_t* pcm, uint64_t n, uint32_t bps, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif
