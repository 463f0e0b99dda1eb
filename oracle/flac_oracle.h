/*
 * flac_oracle.h -- CPU restatement of the reference's FLAC decode path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, link,
 * import or execute it, and there only as the checker or the timed CPU baseline.  The product
 * path (birdnest/audio_b200/csrc, libbnflac.so) never includes or links this file.
 *
 * What it restates: the decode algorithm that BirdNest.Audio reaches through
 *   LibFLAC.FLAC__stream_decoder_process_single   (Library/LibFLACSharp/LibFLACSharp.cs:54-55,
 *                                                  called at Library/BirdNest.Audio/FLACDecoder.cs:215)
 * i.e. the third-party module libFLAC 1.2.1 (20070917) that the reference ships only as the
 * binary Library/BirdNest.Audio/LibFLACDLL/LibFlac.dll (no source under /root/reference), plus
 * the C# interleave/pack that follows it:
 *   FLACDecoder.WriteCallback                     (Library/BirdNest.Audio/FLACDecoder.cs:520-580, 16-bit)
 *   FLACFileReader.CopyFlacBufferToNAudioBuffer   (Library/BirdNest.Audio.UnitTests/FLACFileReader.cs:208-254, 16/24-bit N-ch)
 * The published FLAC format rules followed are the ones listed in SURVEY.md Appendix A, each of
 * which was checked against the DLL.  Parity pinning: the reference's own tests hold no vectors
 * for this path (Library/BirdNest.Audio.UnitTests/Test.cs:9-268 is all ArrayPool), so this
 * restatement is pinned against outputs of the reference binary itself, run in this container
 * through oracle/refdll (tests/golden/ holds the committed streams + PCM hashes it produced).
 */
#ifndef FLAC_ORACLE_H
#define FLAC_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint32_t min_blocksize, max_blocksize, min_framesize, max_framesize;
    uint32_t sample_rate, channels, bits_per_sample;
    uint64_t total_samples;
    uint8_t md5[16];
    uint64_t first_frame_offset; /* byte offset of the first audio frame */
} fo_streaminfo;

/* one record per frame delivered to the (virtual) write callback, in stream order */
typedef struct {
    uint64_t offset;      /* byte offset of the sync code */
    uint32_t length;      /* bytes incl. CRC-16 */
    uint32_t blocksize;
    uint32_t channels, bits_per_sample, channel_assignment /* 0..7 indep, 8 L/S, 9 S/R, 10 M/S */;
    uint32_t sample_rate;
    uint32_t variable;    /* blocking strategy bit */
    uint64_t number;      /* coded frame- or sample-number */
    uint32_t status;      /* 0 ok, 2 = FRAME_CRC_MISMATCH (frame delivered zero-filled) */
} fo_frame;

/* per-subframe detail (tests of K2): 8 per frame max */
typedef struct {
    uint8_t type;       /* 0 CONSTANT 1 VERBATIM 2 FIXED 3 LPC */
    uint8_t order, wasted, precision, shift, rice_method, partition_order, pad;
    uint64_t bit_offset; /* absolute bit offset of the subframe header within the stream */
} fo_subframe;

/* error callback statuses (LibFLACSharp.cs:262-268) */
enum { FO_ERR_LOST_SYNC = 0, FO_ERR_BAD_HEADER = 1, FO_ERR_CRC_MISMATCH = 2, FO_ERR_UNPARSEABLE = 3 };

/* Parse "fLaC" + metadata; returns 0 or negative error. */
int fo_read_streaminfo(const uint8_t* data, size_t len, fo_streaminfo* si);

/* Decode the whole stream.  pcm receives interleaved little-endian samples, ceil(bps/8) bytes each
 * (the FLAC MD5 layout == FLACFileReader layout == FLACDecoder layout for 16-bit mono/stereo).
 * frames/subframes/errors may be NULL.  Returns bytes written (or needed if pcm==NULL), <0 on error. */
int64_t fo_decode(const uint8_t* data, size_t len, uint8_t* pcm, size_t pcm_cap,
                  fo_frame* frames, size_t frames_cap, size_t* nframes,
                  fo_subframe* subframes /* 8*frames_cap */,
                  uint32_t* errors, size_t errors_cap, size_t* nerrors);

/* decode a byte range of frames only (multi-threaded CPU baseline helper): starts at a frame
 * boundary `begin`, stops at `end`. */
int64_t fo_decode_range(const uint8_t* data, size_t len, const fo_streaminfo* si, size_t begin, size_t end,
                        uint8_t* pcm, size_t pcm_cap, size_t* nframes);

uint8_t fo_crc8(const uint8_t* p, size_t n);
uint16_t fo_crc16(const uint8_t* p, size_t n);
void fo_md5(const uint8_t* p, size_t n, uint8_t out[16]);

#ifdef __cplusplus
}
#endif
#endif
