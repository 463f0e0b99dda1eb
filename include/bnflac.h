/*
 * bnflac.h -- C ABI of libbnflac.so, the B200-native FLAC decode engine that replaces the native
 * codec behind BirdNest.Audio's `FLACDecoder : Stream`.
 *
 * Boundary being replaced (reference file:line, all under /root/reference/Library):
 *   LibFLACSharp/LibFLACSharp.cs:42-85,175-185   the 16 FLAC__stream_decoder_* DllImports (cdecl, DLLName "LibFlac" :22)
 *   LibFLACSharp/LibFLACSharp.cs:187-212          the callback delegate types (read/seek/tell/length/eof/write/metadata/error)
 *   LibFLACSharp/LibFLACSharp.cs:24-36,262-268    StreamDecoderState / DecodeError enums (error vocabulary kept 1:1 below)
 *   LibFLACSharp/LibFLACSharp.cs:282-319          FLACMetaData / FLACStreamInfo marshalled views  -> bnflac_info_t
 *   BirdNest.Audio/FLACDecoder.cs:23,72           ctor: new -> init_stream -> process_until_end_of_metadata -> bnflac_open_*
 *   BirdNest.Audio/FLACDecoder.cs:124-224         Read / RequestAnotherFLACPacket / process_single           -> bnflac_read
 *   BirdNest.Audio/FLACDecoder.cs:325-363         ReadCallback (pull model, <=16 KiB per call)                -> bnflac_read_cb
 *   BirdNest.Audio/FLACDecoder.cs:431-473         MetadataCallback (STREAMINFO -> properties, ALFormat)       -> bnflac_info
 *   BirdNest.Audio/FLACDecoder.cs:520-580         WriteCallback (planar int32 -> interleaved LE 16-bit)       -> fused on the GPU
 *   BirdNest.Audio.UnitTests/FLACFileReader.cs:208-254  16/24-bit N-channel interleave (layout followed for bps != 16, ch > 2)
 *   BirdNest.Audio/FLACDecoder.cs:285-319,590-594 Dispose (finish+delete) / ErrorCallback                    -> bnflac_close / frame status
 *
 * Conventions: plain C, cdecl, opaque handles, caller-owned buffers are never retained past a call
 * unless stated, int return 0 = OK / negative = bnflac_err.  A handle is single-threaded (like the
 * reference); different handles may be used from different threads.  There is NO CPU decode path:
 * every decode entry point fails with BNFLAC_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef BNFLAC_H
#define BNFLAC_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BNFLAC_ABI_VERSION 1

typedef struct bnflac bnflac_t;

/* error codes; names mirror libFLAC's StreamDecoderState / DecodeError so the C# wrapper can keep the
 * reference's exception texts ("FLAC: Could not {op} - {state}!", FLACDecoder.cs:98-105,590-594) */
typedef enum {
    BNFLAC_OK = 0,
    BNFLAC_ERR_ARG = -1,            /* NULL / bad argument */
    BNFLAC_ERR_NOT_FLAC = -2,       /* no "fLaC" marker / no STREAMINFO  (reference: init/metadata failure) */
    BNFLAC_ERR_TRUNCATED = -3,      /* metadata runs past the end of the data */
    BNFLAC_ERR_NO_DEVICE = -4,      /* no usable CUDA device: there is no CPU fallback */
    BNFLAC_ERR_CUDA = -5,           /* CUDA runtime error (see bnflac_last_cuda_error) */
    BNFLAC_ERR_MEMORY = -6,         /* StreamDecoderState.MemoryAllocationError */
    BNFLAC_ERR_CAPACITY = -7,       /* destination too small */
    BNFLAC_ERR_ABORTED = -8,        /* StreamDecoderState.Aborted (read callback returned abort) */
    BNFLAC_ERR_UNSUPPORTED = -9,    /* stream shape outside engine limits (e.g. > 8 channels, blocksize > 65535) */
    BNFLAC_ERR_STATE = -10          /* call not valid in this state (e.g. read after close) */
} bnflac_err;

/* decoder state, same numbering as LibFLACSharp.cs:24-36 */
typedef enum {
    BNFLAC_STATE_SEARCH_FOR_METADATA = 0, BNFLAC_STATE_READ_METADATA = 1, BNFLAC_STATE_SEARCH_FOR_FRAME_SYNC = 2,
    BNFLAC_STATE_READ_FRAME = 3, BNFLAC_STATE_END_OF_STREAM = 4, BNFLAC_STATE_OGG_ERROR = 5, BNFLAC_STATE_SEEK_ERROR = 6,
    BNFLAC_STATE_ABORTED = 7, BNFLAC_STATE_MEMORY_ALLOCATION_ERROR = 8, BNFLAC_STATE_UNINITIALIZED = 9
} bnflac_state_t;

/* per-frame status, same numbering as DecodeError (LibFLACSharp.cs:262-268) shifted by one so 0 = OK */
enum { BNFLAC_FRAME_OK = 0, BNFLAC_FRAME_LOST_SYNC = 1, BNFLAC_FRAME_BAD_HEADER = 2, BNFLAC_FRAME_CRC_MISMATCH = 3, BNFLAC_FRAME_UNPARSEABLE = 4 };

/* OpenAL formats FLACDecoder.Format can take (FLACDecoder.cs:454-465); 0 = unmapped + logger warning */
enum { BNFLAC_AL_NONE = 0, BNFLAC_AL_MONO8 = 0x1100, BNFLAC_AL_MONO16 = 0x1101, BNFLAC_AL_STEREO8 = 0x1102, BNFLAC_AL_STEREO16 = 0x1103 };

typedef struct {
    uint32_t struct_size;       /* = sizeof(bnflac_opts); 0-initialise the rest for defaults */
    int32_t  device;            /* CUDA device ordinal, -1 = current device */
    void*    stream;            /* cudaStream_t to launch on; NULL = library-owned stream */
    uint32_t shard_index;       /* frame-range sharding: this handle owns shard_index of shard_count */
    uint32_t shard_count;       /* 0 or 1 = whole stream */
    uint32_t flags;             /* BNFLAC_OPT_* */
    uint32_t read_chunk_frames; /* frames decoded per look-ahead batch by bnflac_read; 0 = default */
} bnflac_opts;
#define BNFLAC_OPT_VERIFY_MD5 1u   /* decode_all also checks md5(PCM) against STREAMINFO (host side, not timed) */
#define BNFLAC_OPT_BORROW_INPUT 2u /* open_memory does not copy `data`: the caller keeps it valid (and ideally pinned) until close */
#define BNFLAC_OPT_PACKED_INPUT 8u /* decode_batch: the clips lie in ascending address order inside ONE host buffer that the caller owns from the
                                    * first clip's first byte to the last clip's last byte (gaps included): the range is uploaded in place */
#define BNFLAC_OPT_LAZY_PULL 4u    /* open_callbacks pulls only the metadata; bnflac_read pulls the rest as the reader advances (the
                                    * callback and its `user` must then stay valid until close).  What FLACDecoder(Stream) does:
                                    * metadata in the constructor, stream bytes on demand (FLACDecoder.cs:72-88,207-224,325-363) */

/* what MetadataCallback derives (FLACDecoder.cs:431-473) plus the raw STREAMINFO fields */
typedef struct {
    uint32_t sample_rate, channels, bits_per_sample;
    uint32_t min_blocksize, max_blocksize, min_framesize, max_framesize;
    uint32_t block_align;       /* channels * (bits_per_sample/8)  (FLACDecoder.cs:448) */
    uint32_t al_format;         /* BNFLAC_AL_* */
    uint32_t bytes_per_sample;  /* ceil(bits_per_sample/8): width of one packed PCM sample in the output */
    uint64_t total_samples;     /* per channel, 0 = unknown */
    uint64_t pcm_bytes;         /* total_samples * channels * bytes_per_sample */
    uint64_t length_reference;  /* FLACDecoder.Length as the reference computes it (block_align * total_samples, low 32 bits of total) */
    double   duration_seconds;  /* FLACDecoder.Duration */
    uint8_t  md5[16];
    uint64_t first_frame_offset;
} bnflac_info_t;

/* pull-model source, the shape of FLACDecoder.ReadCallback (FLACDecoder.cs:325-363):
 * fill buf with up to *bytes bytes, store the count in *bytes; return 0 continue, 1 end of stream, 2 abort */
typedef int (*bnflac_read_cb)(void* user, uint8_t* buf, size_t* bytes);

/* ---- open / info / close ------------------------------------------------------------------ */
/* Opens a stream held in host memory.  The library copies what it needs to the device; `data` must stay
 * valid until bnflac_close only with BNFLAC_OPT_BORROW_INPUT (otherwise it is copied: the caller keeps ownership).
 * A stream that starts with an Ogg page ("OggS") is taken as Ogg FLAC (mapping 1.x): its pages are taken apart on the host
 * (page CRC-32 and sequence checked; damaged or missing pages cost the packets they carry) and the native stream inside is
 * decoded; frame offsets in the diagnostics then refer to that native stream.  Also through bnflac_open_callbacks (the
 * whole stream is pulled at open); not through bnflac_open_device / bnflac_decode_batch. */
int bnflac_open_memory(const uint8_t* data, size_t len, const bnflac_opts* opts, bnflac_t** out);
/* Opens from a pull callback (a C# Stream): pulls the whole stream in large requests, then as open_memory;
 * with BNFLAC_OPT_LAZY_PULL only the metadata is pulled here and the rest on demand by bnflac_read. */
int bnflac_open_callbacks(bnflac_read_cb read, void* user, const bnflac_opts* opts, bnflac_t** out);
/* Opens a stream whose bytes are ALREADY resident in device memory (d_data stays owned by the caller and must
 * outlive the handle).  `header` = at least the first header_len bytes of the same stream in host memory
 * (metadata is parsed on the host).  Benchmark / pipeline path.
 * The kernels fetch the input in aligned 16-byte units: the allocation behind d_data must be READABLE for 64
 * bytes past d_data + len (what those bytes hold does not matter; nothing past len is ever interpreted). */
int bnflac_open_device(const void* d_data, size_t len, const uint8_t* header, size_t header_len, const bnflac_opts* opts, bnflac_t** out);
int bnflac_info(bnflac_t* h, bnflac_info_t* info);
/* Host-only (no device needed): the metadata of a stream held in host memory -- what bnflac_open_* + bnflac_info would
 * report, i.e. what SetupStreamInfo + MetadataCallback derive (FLACDecoder.cs:66-70,431-473).  `data` needs to hold the
 * metadata blocks only (BNFLAC_ERR_TRUNCATED if they are cut short); ID3v2 prefixes and Ogg FLAC are accepted. */
int bnflac_probe(const uint8_t* data, size_t len, bnflac_info_t* info);
/* Host-only: the native FLAC stream inside an Ogg FLAC stream (what the decoder is fed after de-paging).  *written = its
 * size; BNFLAC_ERR_CAPACITY (nothing copied) if cap is smaller, BNFLAC_ERR_NOT_FLAC if `data` is not Ogg FLAC. */
int bnflac_ogg_to_native(const uint8_t* data, size_t len, uint8_t* dst, size_t cap, size_t* written);
/* Host-only: the byte range [*own_begin, *own_end) of frame data that shard `index` of `count` owns in a stream of `len` bytes
 * whose first frame starts at `first_frame_offset` (a shard owns the frames whose sync code lies in its range).
 * The one place this arithmetic lives: bnflac_opts.shard_index / shard_count use it, so does any host-side planner. */
int bnflac_shard_range(uint64_t len, uint64_t first_frame_offset, uint32_t index, uint32_t count, uint64_t* own_begin, uint64_t* own_end);
int bnflac_state(bnflac_t* h);                       /* bnflac_state_t */
void bnflac_close(bnflac_t* h);                      /* finish + delete (FLACDecoder.cs:296-300) */

/* ---- decode -------------------------------------------------------------------------------- */
/* Stream read (FLACDecoder.Read, FLACDecoder.cs:124-205): interleaved little-endian PCM, bytes_per_sample each.
 * Returns bytes written (== count unless the stream ended), 0 at end of stream, <0 = bnflac_err. */
int64_t bnflac_read(bnflac_t* h, uint8_t* dst, size_t count);
/* One shot to a host buffer (what OpenALDemo does with CopyTo(MemoryStream), Program.cs:33-38). */
int bnflac_decode_all(bnflac_t* h, uint8_t* dst, size_t cap, uint64_t* written);
/* One shot, result left in device memory.  If d_dst is NULL the library allocates (owned by the handle, valid
 * until the next decode or close) and returns it in *d_out; else writes into caller memory of capacity cap. */
int bnflac_decode_device(bnflac_t* h, void* d_dst, size_t cap, void** d_out, uint64_t* written);
/* Size in bytes the decode of this handle (its shard) will produce, known after the frame scan.  For a large host-resident
 * stream this is what STREAMINFO promises (no scan); if bnflac_decode_all then reports BNFLAC_ERR_CAPACITY with a buffer of
 * that size (STREAMINFO understated the stream), the next call here scans and returns the exact figure: size again, retry. */
int bnflac_decoded_size(bnflac_t* h, uint64_t* bytes);

/* ---- batch of independent clips (BASELINE cfg4: sharded by file) --------------------------- */
typedef struct { const uint8_t* data; size_t len; } bnflac_span;
typedef struct { uint64_t pcm_offset, pcm_bytes; uint32_t sample_rate, channels, bits_per_sample, status; uint64_t total_samples; } bnflac_clip_result;
/* Decodes n clips in one pipeline pass into one PCM buffer (clip i at results[i].pcm_offset).
 * dst==NULL: only sizes are computed.  dst_is_device: dst is a device pointer.
 * With BNFLAC_OPT_PACKED_INPUT in opts->flags (clips in ascending address order inside ONE host buffer owned by the caller,
 * e.g. an archive read in one piece) the whole range is uploaded in place with a single copy; otherwise only the clips'
 * own bytes are read: they are gathered into pinned staging memory first. */
int bnflac_decode_batch(const bnflac_span* clips, size_t n, const bnflac_opts* opts, uint8_t* dst, size_t cap, int dst_is_device,
                        bnflac_clip_result* results, uint64_t* written);

/* ---- encode ------------------------------------------------------------------ */
/* The encoder half of the native codec: LibFLACSharp.cs:322-387 declares libFLAC's stream encoder (new, set_channels :333,
 * set_bits_per_sample :336, set_sample_rate :339, set_compression_level :342, set_blocksize :345, init_stream :348,
 * process_interleaved :354, process :357, set_do_mid_side_stereo :366, finish :327, delete :330) and its write callback :375.
 * One call here encodes a whole PCM buffer on the GPU (frames are independent: one CTA per frame); libLibFlac.so replays the
 * result frame by frame behind those legacy symbols.  Fixed blocksize, 1-8 channels, 4-24 bits, blocksize 16..16384, FIXED 0-4 and
 * LPC 1-32 predictors, partitioned Rice / Rice2 (partition order 0-8), all four stereo assignments, wasted bits, CONSTANT and
 * VERBATIM subframes.  PCM layout = what the decoder returns: interleaved, little-endian, ceil(bits/8) bytes per sample
 * (or int32 per sample with BNFLAC_ENC_INPUT_INT32, the layout of FLAC__stream_encoder_process_interleaved). */
typedef struct {
    uint32_t struct_size;           /* = sizeof(bnflac_enc_opts); 0-initialise the rest for defaults */
    int32_t  device;                /* CUDA device ordinal, -1 = current device */
    void*    stream;                /* cudaStream_t to launch on; NULL = default stream */
    uint32_t sample_rate, channels, bits_per_sample;
    uint32_t blocksize;             /* 0 = 4096 (or the preset's) */
    uint32_t max_lpc_order;         /* 0 = FIXED predictors only */
    uint32_t qlp_precision;         /* 0 = libFLAC's choice by sample width and blocksize */
    uint32_t min_partition_order, max_partition_order;
    uint32_t mid_side;              /* stereo only: 1 = cheapest of L/R, L/S, S/R, M/S per frame; 0 = independent channels */
    uint32_t compression_level;     /* used with BNFLAC_ENC_USE_LEVEL: libFLAC's presets 0-8 (set_compression_level) */
    uint32_t flags;                 /* BNFLAC_ENC_* */
    uint64_t first_frame_number;    /* coded number of the first frame: a stream encoded in several calls (frames are independent; the caller
                                     * keeps only the frames of the later calls and writes STREAMINFO itself -- what libLibFlac.so does) */
} bnflac_enc_opts;
#define BNFLAC_ENC_NO_MD5 1u        /* leave STREAMINFO's MD5 zero (the MD5 is serial host work: it runs beside the GPU for host input, and
                                     * costs a device-to-host copy of the PCM for bnflac_encode_device) */
#define BNFLAC_ENC_INPUT_INT32 2u   /* one int32 per sample instead of packed bytes */
#define BNFLAC_ENC_USE_LEVEL 4u     /* blocksize / max_lpc_order / partition orders / mid_side come from compression_level */
#define BNFLAC_ENC_FIXED_ORDER 8u   /* always use max_lpc_order instead of the order with the fewest estimated bits */
typedef struct {
    float plan_ms, write_ms, total_ms;   /* CUDA-event times of the analysis kernel (+ prefix sum), the bitstream kernel, and both + the memset */
    uint32_t frames;
    uint64_t bytes;                      /* stream size */
    uint32_t min_framesize, max_framesize;
    uint32_t* frame_sizes;               /* in: NULL, or room for frame_sizes_cap entries; out: the size in bytes of every frame, in order */
    uint64_t frame_sizes_cap;
} bnflac_enc_stats;
/* Host-only: an upper bound of the stream size for `pcm_bytes` of input. */
int bnflac_encode_bound(size_t pcm_bytes, const bnflac_enc_opts* opts, uint64_t* bound);
/* Host PCM -> host FLAC stream ("fLaC" + STREAMINFO + frames).  *written = stream size; BNFLAC_ERR_CAPACITY if cap is smaller
 * (dst == NULL: size only).  stats may be NULL; zero-initialise it otherwise (frame_sizes is an input). */
int bnflac_encode(const uint8_t* pcm, size_t pcm_bytes, const bnflac_enc_opts* opts, uint8_t* dst, size_t cap, uint64_t* written, bnflac_enc_stats* stats);
/* Device PCM -> device FLAC stream (d_dst 4-byte aligned, cap bytes; bnflac_encode_bound is always enough; d_pcm aligned to the
 * sample container when that is 2 or 4 bytes). */
int bnflac_encode_device(const void* d_pcm, size_t pcm_bytes, const bnflac_enc_opts* opts, void* d_dst, size_t cap, uint64_t* written, bnflac_enc_stats* stats);

/* ---- diagnostics --------------------------------------------------------------------------- */
/* per-frame table of the last decode (host copies, owned by the handle) */
typedef struct { uint64_t offset; uint32_t length, blocksize; uint8_t channels, bits_per_sample, assignment, status; uint32_t pad; uint64_t number; uint64_t pcm_offset; } bnflac_frame_t;
int bnflac_frames(bnflac_t* h, const bnflac_frame_t** frames, size_t* n);
/* per-subframe table (K2 output) of the last decode: 8 entries per frame */
typedef struct { uint32_t bit_offset; uint8_t type, order, wasted, flags; } bnflac_subframe_t;
int bnflac_subframes(bnflac_t* h, const bnflac_subframe_t** sub, size_t* n);
/* number of error-callback events the reference would have raised (ErrorCallback, FLACDecoder.cs:590-594) and their codes */
int bnflac_errors(bnflac_t* h, const uint32_t** codes, size_t* n);
/* the same list, but during a streamed bnflac_read session only the events of the sub-shards issued so far: nothing is
 * decoded (or, with BNFLAC_OPT_LAZY_PULL, pulled) ahead for it, so a host can poll it after every Read and raise an error
 * no later than the Read that reaches the damaged frame, like ErrorCallback does (FLACDecoder.cs:590-594).  The list
 * only grows within a session; outside of one it equals bnflac_errors. */
int bnflac_errors_so_far(bnflac_t* h, const uint32_t** codes, size_t* n);
/* for each of those events, how many frames had been delivered before it was raised (lets a frame-at-a-time host such as
 * the libFLAC-symbol shim raise its error callbacks at the right moment) */
int bnflac_error_frames(bnflac_t* h, const uint32_t** at, size_t* n);
/* CUDA-event timings (ms) of the stages of the last decode on this handle */
typedef struct { float total, scan, crc, link, parse, decode; uint32_t launches; uint32_t pad; } bnflac_timing;
int bnflac_last_timing(bnflac_t* h, bnflac_timing* t);
const char* bnflac_strerror(int err);
const char* bnflac_state_name(int state);           /* "EndOfStream", ... as C# prints the enum */
const char* bnflac_frame_status_name(int status);   /* "LostSync", "BadHeader", "FrameCrcMismatch", "UnparsableStream" */
const char* bnflac_last_cuda_error(void);
int bnflac_abi_version(void);
int bnflac_device_count(void);
/* kernels launched by this process so far (bench.py's gpu_launches is a difference of two reads) */
uint64_t bnflac_kernel_launches(void);
/* Device and pinned-host blocks are cached process-wide between handles (cudaMalloc/cudaFree cost milliseconds per
 * gigabyte and cudaFree stalls the device); this returns every cached block to the driver. */
void bnflac_trim_pools(void);

#ifdef __cplusplus
}
#endif
#endif
