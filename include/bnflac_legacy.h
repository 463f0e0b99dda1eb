/*
 * bnflac_legacy.h -- the libFLAC 1.2.1 stream-decoder symbols BirdNest.Audio P/Invokes, served by the bnflac engine
 * (SURVEY 8b tier A / 8f-1).  libLibFlac.so exports them, so the UNMODIFIED reference sources
 *   Library/LibFLACSharp/LibFLACSharp.cs:42-85,175-185   (DllImport "LibFlac", cdecl)
 *   Library/BirdNest.Audio/FLACDecoder.cs:49-70,207-224  (new -> init_stream -> process_until_end_of_metadata -> process_single*)
 *   Library/BirdNest.Audio.UnitTests/FLACFileReader.cs:53-77,160-180,298,357-361 (init_file, get_total_samples, seek_absolute)
 * run against the GPU engine through a Mono dllmap / a renamed library.  It is a replay layer: the first
 * process_single decodes the whole stream on the GPU, every call then hands one frame to the write callback as planar
 * int32 (what libFLAC hands out, FLACDecoder.cs:493-494); it cannot be the fast path and is not meant to be.
 *
 * Struct views follow the native x86-64 / i386-MSVC layouts the C# marshals against (LibFLACSharp.cs:216-234,282-319):
 * FLAC__FrameHeader 40 bytes {blocksize@0, sample_rate@4, channels@8, channel_assignment@12, bits_per_sample@16,
 * number_type@20, number@24, crc@32}; FLAC__StreamMetadata {type@0, is_last@4, length@8, stream_info@16:
 * min/max_blocksize, min/max_framesize, sample_rate@32, channels@36, bits_per_sample@40, total_samples@48, md5@56}.
 */
#ifndef BNFLAC_LEGACY_H
#define BNFLAC_LEGACY_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct FLAC__StreamDecoder FLAC__StreamDecoder;
typedef int32_t FLAC__bool;

typedef struct {
    uint32_t blocksize, sample_rate, channels, channel_assignment /* 0 independent 1 left/side 2 right/side 3 mid/side */, bits_per_sample;
    uint32_t number_type; /* 0 frame number, 1 sample number */
    union { uint32_t frame_number; uint64_t sample_number; } number;
    uint8_t crc;
} FLAC__FrameHeader;
typedef struct { FLAC__FrameHeader header; uint8_t subframes_and_footer[8 * 512 + 16]; /* not filled in: the reference reads only the header */ } FLAC__Frame;

typedef struct {
    uint32_t type /* 0 = STREAMINFO */, is_last, length, pad_;
    struct { uint32_t min_blocksize, max_blocksize, min_framesize, max_framesize, sample_rate, channels, bits_per_sample, pad_; uint64_t total_samples; uint8_t md5sum[16]; } stream_info;
    uint8_t rest_[128];
} FLAC__StreamMetadata;

/* callback shapes: LibFLACSharp.cs:187-212 */
typedef int (*FLAC__StreamDecoderReadCallback)(const FLAC__StreamDecoder*, uint8_t buffer[], size_t* bytes, void* client);      /* 0 continue, 1 end of stream, 2 abort */
typedef int (*FLAC__StreamDecoderSeekCallback)(const FLAC__StreamDecoder*, uint64_t absolute_byte_offset, void* client);
typedef int (*FLAC__StreamDecoderTellCallback)(const FLAC__StreamDecoder*, uint64_t* absolute_byte_offset, void* client);
typedef int (*FLAC__StreamDecoderLengthCallback)(const FLAC__StreamDecoder*, uint64_t* stream_length, void* client);
typedef FLAC__bool (*FLAC__StreamDecoderEofCallback)(const FLAC__StreamDecoder*, void* client);
typedef int (*FLAC__StreamDecoderWriteCallback)(const FLAC__StreamDecoder*, const FLAC__Frame*, const int32_t* const buffer[], void* client); /* 0 continue, 1 abort */
typedef void (*FLAC__StreamDecoderMetadataCallback)(const FLAC__StreamDecoder*, const FLAC__StreamMetadata*, void* client);
typedef void (*FLAC__StreamDecoderErrorCallback)(const FLAC__StreamDecoder*, int status, void* client);   /* 0 LostSync 1 BadHeader 2 FrameCrcMismatch 3 Unparseable */

FLAC__StreamDecoder* FLAC__stream_decoder_new(void);                                             /* LibFLACSharp.cs:43 */
void FLAC__stream_decoder_delete(FLAC__StreamDecoder*);                                          /* :49 */
FLAC__bool FLAC__stream_decoder_finish(FLAC__StreamDecoder*);                                    /* :46 */
int FLAC__stream_decoder_init_file(FLAC__StreamDecoder*, const char* filename, FLAC__StreamDecoderWriteCallback, FLAC__StreamDecoderMetadataCallback,
                                   FLAC__StreamDecoderErrorCallback, void* client);             /* :52 */
int FLAC__stream_decoder_init_stream(FLAC__StreamDecoder*, FLAC__StreamDecoderReadCallback, FLAC__StreamDecoderSeekCallback, FLAC__StreamDecoderTellCallback,
                                     FLAC__StreamDecoderLengthCallback, FLAC__StreamDecoderEofCallback, FLAC__StreamDecoderWriteCallback,
                                     FLAC__StreamDecoderMetadataCallback, FLAC__StreamDecoderErrorCallback, void* client);   /* :175-185 */
FLAC__bool FLAC__stream_decoder_process_single(FLAC__StreamDecoder*);                            /* :55 */
FLAC__bool FLAC__stream_decoder_process_until_end_of_metadata(FLAC__StreamDecoder*);             /* :58 */
FLAC__bool FLAC__stream_decoder_process_until_end_of_stream(FLAC__StreamDecoder*);               /* :61 */
FLAC__bool FLAC__stream_decoder_seek_absolute(FLAC__StreamDecoder*, uint64_t sample);            /* :64 */
FLAC__bool FLAC__stream_decoder_get_decode_position(const FLAC__StreamDecoder*, uint64_t* position);   /* :67 */
uint64_t FLAC__stream_decoder_get_total_samples(const FLAC__StreamDecoder*);                     /* :70 */
unsigned FLAC__stream_decoder_get_channels(const FLAC__StreamDecoder*);                          /* :73 */
unsigned FLAC__stream_decoder_get_bits_per_sample(const FLAC__StreamDecoder*);                   /* :76 */
unsigned FLAC__stream_decoder_get_sample_rate(const FLAC__StreamDecoder*);                       /* :79 */
int FLAC__stream_decoder_get_state(const FLAC__StreamDecoder*);                                  /* :82, numbering of LibFLACSharp.cs:24-36 */
FLAC__bool FLAC__stream_decoder_reset(FLAC__StreamDecoder*);                                     /* :85 */

/* ---- the encoder half (LibFLACSharp.cs:322-387), SURVEY 8f-4 ---------------------------------------------------------------------
 * Same symbols, same argument meaning; the work is done by bnflac_encode (one CTA per frame on the GPU).  A replay layer like the
 * decoder half: init writes "fLaC" and a provisional STREAMINFO block through the write callback, samples handed to process /
 * process_interleaved are collected and, every 1024 blocks, encoded in one GPU call and handed out frame by frame (samples = blocksize,
 * current_frame counting from 0) -- output flows during process(), as with libFLAC; finish() encodes the rest and then puts the final
 * STREAMINFO (frame sizes, total samples, MD5 kept as a running digest) where the provisional one is when the sink can seek (the seek
 * callback, or the file of init_file), and reports it through the metadata callback.  State numbering: libFLAC's FLAC__StreamEncoderState (0 OK,
 * 1 UNINITIALIZED, 5 CLIENT_ERROR, 6 IO_ERROR, 8 MEMORY_ALLOCATION_ERROR); init status: 0 OK, 1 ENCODER_ERROR, 3 INVALID_CALLBACKS,
 * 4 INVALID_NUMBER_OF_CHANNELS, 5 INVALID_BITS_PER_SAMPLE, 6 INVALID_SAMPLE_RATE, 7 INVALID_BLOCK_SIZE, 13 ALREADY_INITIALIZED. */
typedef struct FLAC__StreamEncoder FLAC__StreamEncoder;
typedef int (*FLAC__StreamEncoderWriteCallback)(const FLAC__StreamEncoder*, const uint8_t buffer[], size_t bytes, unsigned samples, unsigned current_frame, void* client);  /* :377; 0 ok, 1 fatal */
typedef int (*FLAC__StreamEncoderSeekCallback)(const FLAC__StreamEncoder*, uint64_t absolute_byte_offset, void* client);      /* :380: used once, by finish(), to rewrite STREAMINFO */
typedef int (*FLAC__StreamEncoderTellCallback)(const FLAC__StreamEncoder*, uint64_t* absolute_byte_offset, void* client);     /* :383 (accepted, never called) */
typedef void (*FLAC__StreamEncoderMetadataCallback)(const FLAC__StreamEncoder*, const FLAC__StreamMetadata*, void* client);    /* :386 */
FLAC__StreamEncoder* FLAC__stream_encoder_new(void);                                              /* :325 */
FLAC__bool FLAC__stream_encoder_finish(FLAC__StreamEncoder*);                                     /* :328 */
void FLAC__stream_encoder_delete(FLAC__StreamEncoder*);                                           /* :331 */
FLAC__bool FLAC__stream_encoder_set_channels(FLAC__StreamEncoder*, unsigned);                     /* :334 */
FLAC__bool FLAC__stream_encoder_set_bits_per_sample(FLAC__StreamEncoder*, unsigned);              /* :337 */
FLAC__bool FLAC__stream_encoder_set_sample_rate(FLAC__StreamEncoder*, unsigned);                  /* :340 */
FLAC__bool FLAC__stream_encoder_set_compression_level(FLAC__StreamEncoder*, unsigned);            /* :343 */
FLAC__bool FLAC__stream_encoder_set_blocksize(FLAC__StreamEncoder*, unsigned);                    /* :346 */
int FLAC__stream_encoder_init_stream(FLAC__StreamEncoder*, FLAC__StreamEncoderWriteCallback, FLAC__StreamEncoderSeekCallback, FLAC__StreamEncoderTellCallback,
                                     FLAC__StreamEncoderMetadataCallback, void* client);          /* :349 */
int FLAC__stream_encoder_init_file(FLAC__StreamEncoder*, const char* filename, void* progress_callback, void* client);   /* :352 */
FLAC__bool FLAC__stream_encoder_process_interleaved(FLAC__StreamEncoder*, const int32_t buffer[], unsigned samples);      /* :355 */
FLAC__bool FLAC__stream_encoder_process(FLAC__StreamEncoder*, const int32_t* const buffer[], unsigned samples);          /* :358 */
FLAC__bool FLAC__stream_encoder_set_verify(FLAC__StreamEncoder*, FLAC__bool);                     /* :361: finish() then decodes the stream again on the GPU and compares */
FLAC__bool FLAC__stream_encoder_set_streamable_subset(FLAC__StreamEncoder*, FLAC__bool);          /* :364 (accepted; what is written is always within the subset's limits for blocksize <= 4608 / 16384) */
FLAC__bool FLAC__stream_encoder_set_do_mid_side_stereo(FLAC__StreamEncoder*, FLAC__bool);         /* :367 */
FLAC__bool FLAC__stream_encoder_set_loose_mid_side_stereo(FLAC__StreamEncoder*, FLAC__bool);      /* :370 (accepted: the choice is made per frame from exact sizes) */
int FLAC__stream_encoder_get_state(const FLAC__StreamEncoder*);                                   /* :373 */

#ifdef __cplusplus
}
#endif
#endif
