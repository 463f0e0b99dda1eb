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

#ifdef __cplusplus
}
#endif
#endif
