// libflac_shim.cpp -- libLibFlac.so: the FLAC__stream_decoder_* symbols of include/bnflac_legacy.h over a bnflac handle.
// A replay layer (SURVEY 8b tier A): the stream is pulled through the caller's read callback (or read from the file),
// decoded in one go by the CUDA engine at the first process_single, and handed out frame by frame as planar int32 with
// the error callbacks raised where the reference raises them.  No decoding happens here.
#include <bnflac.h>
#include <bnflac_legacy.h>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

struct FLAC__StreamDecoder {
    int state = BNFLAC_STATE_UNINITIALIZED;
    FLAC__StreamDecoderReadCallback read = nullptr;
    FLAC__StreamDecoderEofCallback eof = nullptr;
    FLAC__StreamDecoderWriteCallback write = nullptr;
    FLAC__StreamDecoderMetadataCallback metadata = nullptr;
    FLAC__StreamDecoderErrorCallback error = nullptr;
    void* client = nullptr;
    std::vector<uint8_t> bytes;            // the compressed stream
    bool have_bytes = false, metadata_done = false, decoded = false;
    bnflac_t* h = nullptr;
    bnflac_info_t info{};
    std::vector<uint8_t> pcm;
    const bnflac_frame_t* frames = nullptr; size_t nframes = 0;
    const uint32_t* errs = nullptr; const uint32_t* errs_at = nullptr; size_t nerrs = 0;
    size_t next_frame = 0, next_err = 0;
    uint64_t skip_samples = 0;             // after seek_absolute: samples of the next frame that lie before the target
    std::vector<uint64_t> first_sample;    // per delivered frame
    std::vector<int32_t> planes[8];

    void close_engine() { if (h) bnflac_close(h); h = nullptr; frames = nullptr; nframes = 0; errs = errs_at = nullptr; nerrs = 0; decoded = false; }
};

static bool pull_stream(FLAC__StreamDecoder* d) {
    if (d->have_bytes) return true;
    if (!d->read) return false;
    // The C# callback caps each read at its own byte[] (16 KiB, FLACDecoder.cs:336).  The vector is kept at its high-water
    // size and grown geometrically: `fill` counts what has been read, so a 16 KiB read never value-initialises a megabyte.
    const size_t req = 1u << 20;
    size_t fill = 0;
    try {
        for (;;) {
            if (d->bytes.size() < fill + req) d->bytes.resize(std::max(fill + req, d->bytes.size() + d->bytes.size() / 2));
            size_t got = req;
            const int st = d->read(d, d->bytes.data() + fill, &got, d->client);
            if (st == 2 || got > req) { d->bytes.clear(); d->state = BNFLAC_STATE_ABORTED; return false; }
            fill += got;
            if (st == 1 || got == 0) break;
        }
        d->bytes.resize(fill);
    } catch (const std::bad_alloc&) { d->bytes.clear(); d->state = BNFLAC_STATE_MEMORY_ALLOCATION_ERROR; return false; }
    d->have_bytes = true;
    return true;
}

static bool do_metadata(FLAC__StreamDecoder* d) {
    if (d->metadata_done) return true;
    if (!pull_stream(d)) return false;
    bnflac_opts o{}; o.struct_size = sizeof o; o.device = -1; o.flags = BNFLAC_OPT_BORROW_INPUT;
    const int rc = bnflac_open_memory(d->bytes.data(), d->bytes.size(), &o, &d->h);
    if (rc) { d->state = (rc == BNFLAC_ERR_MEMORY) ? BNFLAC_STATE_MEMORY_ALLOCATION_ERROR : BNFLAC_STATE_END_OF_STREAM; return false; }
    bnflac_info(d->h, &d->info);
    if (d->metadata) {
        FLAC__StreamMetadata m; memset(&m, 0, sizeof m);
        m.type = 0; m.is_last = 1; m.length = 34;
        m.stream_info.min_blocksize = d->info.min_blocksize; m.stream_info.max_blocksize = d->info.max_blocksize;
        m.stream_info.min_framesize = d->info.min_framesize; m.stream_info.max_framesize = d->info.max_framesize;
        m.stream_info.sample_rate = d->info.sample_rate; m.stream_info.channels = d->info.channels; m.stream_info.bits_per_sample = d->info.bits_per_sample;
        m.stream_info.total_samples = d->info.total_samples; memcpy(m.stream_info.md5sum, d->info.md5, 16);
        d->metadata(d, &m, d->client);
    }
    d->metadata_done = true;
    d->state = BNFLAC_STATE_SEARCH_FOR_FRAME_SYNC;
    return true;
}

static bool do_decode(FLAC__StreamDecoder* d) {
    if (d->decoded) return true;
    uint64_t w = 0;
    int rc = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        uint64_t need = 0;
        if (bnflac_decoded_size(d->h, &need)) { d->state = BNFLAC_STATE_ABORTED; return false; }
        // a STREAMINFO that overstates the stream grossly must not take the process down through an extern "C" frame
        try { d->pcm.resize((size_t)need + 64); } catch (const std::bad_alloc&) { d->state = BNFLAC_STATE_MEMORY_ALLOCATION_ERROR; return false; }
        rc = bnflac_decode_all(d->h, d->pcm.data(), d->pcm.size(), &w);
        if (rc != BNFLAC_ERR_CAPACITY) break;          // STREAMINFO understated the stream: decoded_size scans the second time
    }
    if (rc) { d->state = (rc == BNFLAC_ERR_MEMORY) ? BNFLAC_STATE_MEMORY_ALLOCATION_ERROR : BNFLAC_STATE_ABORTED; return false; }
    d->pcm.resize((size_t)w);
    size_t ne2 = 0;
    if (bnflac_frames(d->h, &d->frames, &d->nframes) || bnflac_errors(d->h, &d->errs, &d->nerrs) || bnflac_error_frames(d->h, &d->errs_at, &ne2)) { d->state = BNFLAC_STATE_ABORTED; return false; }
    d->first_sample.assign(d->nframes + 1, 0);
    for (size_t i = 0; i < d->nframes; i++) d->first_sample[i + 1] = d->first_sample[i] + d->frames[i].blocksize;
    d->decoded = true;
    return true;
}

// hands frame `i` (minus its first `skip` samples) to the write callback
static bool deliver(FLAC__StreamDecoder* d, size_t i, uint64_t skip) {
    const bnflac_frame_t& f = d->frames[i];
    const uint32_t C = f.channels, B = d->info.bytes_per_sample, bs = f.blocksize;
    const uint8_t* src = d->pcm.data() + f.pcm_offset;
    const int sh = 32 - 8 * (int)B;
    for (uint32_t c = 0; c < C; c++) d->planes[c].resize(bs);
    for (uint32_t t = 0; t < bs; t++)
        for (uint32_t c = 0; c < C; c++) {
            const uint8_t* p = src + ((size_t)t * C + c) * B;
            uint32_t v = 0;
            for (uint32_t k = 0; k < B; k++) v |= (uint32_t)p[k] << (8 * k);
            d->planes[c][t] = (int32_t)(v << sh) >> sh;                   // sign-extend the packed sample
        }
    FLAC__Frame fr; memset(&fr, 0, sizeof fr);
    fr.header.blocksize = bs - (uint32_t)skip; fr.header.sample_rate = d->info.sample_rate; fr.header.channels = C;
    fr.header.channel_assignment = f.assignment < 8 ? 0u : (uint32_t)f.assignment - 7u;
    fr.header.bits_per_sample = f.bits_per_sample;
    if (skip) { fr.header.number_type = 1; fr.header.number.sample_number = d->first_sample[i] + skip; }
    else if (d->info.min_blocksize != d->info.max_blocksize && f.number == d->first_sample[i] && i) { fr.header.number_type = 1; fr.header.number.sample_number = f.number; }
    else { fr.header.number_type = 0; fr.header.number.frame_number = (uint32_t)f.number; }
    const int32_t* bufs[8] = {nullptr};
    for (uint32_t c = 0; c < C; c++) bufs[c] = d->planes[c].data() + skip;
    d->state = BNFLAC_STATE_READ_FRAME;
    if (d->write && d->write(d, &fr, bufs, d->client) != 0) { d->state = BNFLAC_STATE_ABORTED; return false; }
    d->state = BNFLAC_STATE_SEARCH_FOR_FRAME_SYNC;
    return true;
}

static void raise_errors_before(FLAC__StreamDecoder* d, size_t frame_index_inclusive) {
    // events the reference raises before (and, for a CRC mismatch, together with) the frame at this index
    while (d->next_err < d->nerrs && d->errs_at[d->next_err] <= frame_index_inclusive) {
        if (d->error) d->error(d, (int)d->errs[d->next_err], d->client);
        d->next_err++;
    }
}

extern "C" {

FLAC__StreamDecoder* FLAC__stream_decoder_new(void) { return new (std::nothrow) FLAC__StreamDecoder; }
void FLAC__stream_decoder_delete(FLAC__StreamDecoder* d) { if (d) { d->close_engine(); delete d; } }

FLAC__bool FLAC__stream_decoder_finish(FLAC__StreamDecoder* d) {
    if (!d) return 0;
    d->close_engine();
    d->bytes.clear(); d->have_bytes = false; d->metadata_done = false; d->next_frame = d->next_err = 0; d->skip_samples = 0;
    d->state = BNFLAC_STATE_UNINITIALIZED;
    return 1;
}

int FLAC__stream_decoder_init_stream(FLAC__StreamDecoder* d, FLAC__StreamDecoderReadCallback read, FLAC__StreamDecoderSeekCallback, FLAC__StreamDecoderTellCallback,
                                     FLAC__StreamDecoderLengthCallback, FLAC__StreamDecoderEofCallback eof, FLAC__StreamDecoderWriteCallback write,
                                     FLAC__StreamDecoderMetadataCallback metadata, FLAC__StreamDecoderErrorCallback error, void* client) {
    if (!d || d->state != BNFLAC_STATE_UNINITIALIZED) return 5;          /* FLAC__STREAM_DECODER_INIT_STATUS_ALREADY_INITIALIZED */
    if (!read || !write || !error) return 2;                              /* ..._INVALID_CALLBACKS */
    d->read = read; d->eof = eof; d->write = write; d->metadata = metadata; d->error = error; d->client = client;
    d->state = BNFLAC_STATE_SEARCH_FOR_METADATA;
    return 0;
}

int FLAC__stream_decoder_init_file(FLAC__StreamDecoder* d, const char* filename, FLAC__StreamDecoderWriteCallback write, FLAC__StreamDecoderMetadataCallback metadata,
                                   FLAC__StreamDecoderErrorCallback error, void* client) {
    if (!d || d->state != BNFLAC_STATE_UNINITIALIZED) return 5;
    if (!write || !error) return 2;
    FILE* f = filename ? fopen(filename, "rb") : nullptr;
    if (!f) return 4;                                                     /* ..._ERROR_OPENING_FILE */
    fseek(f, 0, SEEK_END); const long n = ftell(f); fseek(f, 0, SEEK_SET);
    d->bytes.resize(n > 0 ? (size_t)n : 0);
    const size_t got = d->bytes.empty() ? 0 : fread(d->bytes.data(), 1, d->bytes.size(), f);
    fclose(f);
    d->bytes.resize(got); d->have_bytes = true;
    d->write = write; d->metadata = metadata; d->error = error; d->client = client;
    d->state = BNFLAC_STATE_SEARCH_FOR_METADATA;
    return 0;
}

FLAC__bool FLAC__stream_decoder_process_until_end_of_metadata(FLAC__StreamDecoder* d) { return d && d->state != BNFLAC_STATE_UNINITIALIZED && do_metadata(d) ? 1 : 0; }

FLAC__bool FLAC__stream_decoder_process_single(FLAC__StreamDecoder* d) {
    if (!d || d->state == BNFLAC_STATE_UNINITIALIZED || d->state == BNFLAC_STATE_ABORTED) return 0;
    if (!d->metadata_done) return do_metadata(d) ? 1 : 0;                 // libFLAC: one call consumes the metadata
    if (d->state == BNFLAC_STATE_END_OF_STREAM) return 1;
    if (!do_decode(d)) return 0;
    if (d->next_frame >= d->nframes) {
        raise_errors_before(d, d->nframes);
        d->state = BNFLAC_STATE_END_OF_STREAM;
        return 1;
    }
    const size_t i = d->next_frame++;
    raise_errors_before(d, i);
    const uint64_t skip = d->skip_samples; d->skip_samples = 0;
    return deliver(d, i, skip) ? 1 : 0;
}

FLAC__bool FLAC__stream_decoder_process_until_end_of_stream(FLAC__StreamDecoder* d) {
    if (!d) return 0;
    while (d->state != BNFLAC_STATE_END_OF_STREAM) if (!FLAC__stream_decoder_process_single(d)) return 0;
    return 1;
}

// libFLAC delivers the (partial) frame that holds the target sample from inside seek_absolute; so does this
FLAC__bool FLAC__stream_decoder_seek_absolute(FLAC__StreamDecoder* d, uint64_t sample) {
    if (!d || d->state == BNFLAC_STATE_UNINITIALIZED) return 0;
    if (!do_metadata(d) || !do_decode(d)) return 0;
    const uint64_t total = d->first_sample.empty() ? 0 : d->first_sample.back();
    if (sample >= total) { d->state = BNFLAC_STATE_SEEK_ERROR; return 0; }
    size_t lo = 0, hi = d->nframes;                                       // last frame whose first sample <= target
    while (hi - lo > 1) { const size_t mid = (lo + hi) / 2; if (d->first_sample[mid] <= sample) lo = mid; else hi = mid; }
    d->next_frame = lo + 1;
    d->next_err = 0;
    while (d->next_err < d->nerrs && d->errs_at[d->next_err] <= lo) d->next_err++;    // events of the skipped part are not replayed
    d->state = BNFLAC_STATE_SEARCH_FOR_FRAME_SYNC;
    return deliver(d, lo, sample - d->first_sample[lo]) ? 1 : 0;
}

FLAC__bool FLAC__stream_decoder_get_decode_position(const FLAC__StreamDecoder* d, uint64_t* position) {
    if (!d || !position || !d->decoded) return 0;
    *position = d->next_frame < d->nframes ? d->frames[d->next_frame].offset : d->bytes.size();
    return 1;
}
uint64_t FLAC__stream_decoder_get_total_samples(const FLAC__StreamDecoder* d) { return d && d->metadata_done ? d->info.total_samples : 0; }
unsigned FLAC__stream_decoder_get_channels(const FLAC__StreamDecoder* d) { return d && d->metadata_done ? d->info.channels : 0; }
unsigned FLAC__stream_decoder_get_bits_per_sample(const FLAC__StreamDecoder* d) { return d && d->metadata_done ? d->info.bits_per_sample : 0; }
unsigned FLAC__stream_decoder_get_sample_rate(const FLAC__StreamDecoder* d) { return d && d->metadata_done ? d->info.sample_rate : 0; }
int FLAC__stream_decoder_get_state(const FLAC__StreamDecoder* d) { return d ? d->state : BNFLAC_STATE_UNINITIALIZED; }

FLAC__bool FLAC__stream_decoder_reset(FLAC__StreamDecoder* d) {
    if (!d || d->state == BNFLAC_STATE_UNINITIALIZED) return 0;
    d->next_frame = d->next_err = 0; d->skip_samples = 0;
    d->state = d->metadata_done ? BNFLAC_STATE_SEARCH_FOR_FRAME_SYNC : BNFLAC_STATE_SEARCH_FOR_METADATA;
    return 1;
}

} // extern "C"


// ------------------------------------------------------------------------------------------------ encoder half
// The FLAC__stream_encoder_* symbols of LibFLACSharp.cs:322-387 over bnflac_encode.  Samples are collected (packed, the engine's input
// layout); whenever ENC_CHUNK_FRAMES whole blocks are in hand they are encoded on the GPU in one call (frames are independent: the call is
// told the number of its first frame) and handed to the write callback frame by frame, as libFLAC does during process(); finish()
// encodes the rest, then puts the final STREAMINFO in place -- through the seek callback (or fseek) when there is one, like libFLAC,
// otherwise it is only reported through the metadata callback.  No encoding happens here.
#include "md5.hpp"
static constexpr size_t ENC_CHUNK_FRAMES = 1024;

struct FLAC__StreamEncoder {
    int state = 1;                        // FLAC__STREAM_ENCODER_UNINITIALIZED
    unsigned channels = 2, bps = 16, sample_rate = 44100, level = 5, blocksize = 0;
    int mid_side = -1;                    // -1: the preset's
    bool verify = false;
    FLAC__StreamEncoderWriteCallback write = nullptr;
    FLAC__StreamEncoderSeekCallback seek = nullptr;
    FLAC__StreamEncoderMetadataCallback metadata = nullptr;
    void* client = nullptr;
    FILE* file = nullptr;
    std::vector<uint8_t> pcm;             // interleaved, little-endian, ceil(bps/8) bytes per sample: not yet encoded
    std::vector<uint8_t> out, back;
    std::vector<uint32_t> sizes;
    bnflac_enc_opts o{};
    bnfe::Md5 md5;
    uint64_t frames = 0, samples = 0;     // emitted so far
    uint32_t min_fs = 0xffffffffu, max_fs = 0;
    size_t frame_bytes() const { return (size_t)o.blocksize * channels * ((bps + 7) / 8); }
};

static void enc_streaminfo(const FLAC__StreamEncoder* e, uint8_t o[38], const uint8_t md5[16]) {      // block header + 34 bytes
    const uint32_t bs = e->o.blocksize, mn = e->frames ? e->min_fs : 0, mx = e->max_fs;
    size_t q = 0;
    o[q++] = 0x80; o[q++] = 0; o[q++] = 0; o[q++] = 34;
    o[q++] = (uint8_t)(bs >> 8); o[q++] = (uint8_t)bs; o[q++] = (uint8_t)(bs >> 8); o[q++] = (uint8_t)bs;
    o[q++] = (uint8_t)(mn >> 16); o[q++] = (uint8_t)(mn >> 8); o[q++] = (uint8_t)mn;
    o[q++] = (uint8_t)(mx >> 16); o[q++] = (uint8_t)(mx >> 8); o[q++] = (uint8_t)mx;
    const uint64_t x = ((uint64_t)e->sample_rate << 44) | ((uint64_t)(e->channels - 1) << 41) | ((uint64_t)(e->bps - 1) << 36) | (e->samples & 0xFFFFFFFFFull);
    for (int i = 7; i >= 0; i--) o[q++] = (uint8_t)(x >> (8 * i));
    memcpy(o + q, md5, 16);
}
static bool enc_emit(FLAC__StreamEncoder* e, const uint8_t* p, size_t n, unsigned samples, unsigned frame) {
    if (e->file) { if (fwrite(p, 1, n, e->file) != n) { e->state = 6; return false; } return true; }
    if (e->write(e, p, n, samples, frame, e->client) != 0) { e->state = 5; return false; }
    return true;
}
static int enc_init_common(FLAC__StreamEncoder* e) {
    if (e->state != 1) return 13;
    if (e->channels < 1 || e->channels > 8) return 4;
    if (e->bps < 4 || e->bps > 24) return 5;
    if (e->sample_rate < 1 || e->sample_rate > 655350) return 6;
    if (e->blocksize && (e->blocksize < 16 || e->blocksize > 16384)) return 7;
    // the preset spelled out (an explicit set_do_mid_side_stereo overrides its stereo setting)
    static const struct { uint32_t bs, lpc, ms, po; } lv[9] = {{1152, 0, 0, 3}, {1152, 0, 1, 3}, {1152, 0, 1, 3}, {4096, 6, 0, 4}, {4096, 8, 1, 4}, {4096, 8, 1, 5}, {4096, 8, 1, 6}, {4096, 8, 1, 6}, {4096, 12, 1, 6}};
    bnflac_enc_opts& o = e->o;
    o = bnflac_enc_opts{}; o.struct_size = sizeof o; o.device = -1;
    o.sample_rate = e->sample_rate; o.channels = e->channels; o.bits_per_sample = e->bps;
    o.blocksize = e->blocksize ? e->blocksize : lv[e->level].bs;
    o.max_lpc_order = lv[e->level].lpc; o.max_partition_order = lv[e->level].po;
    o.mid_side = e->channels == 2 ? (e->mid_side >= 0 ? (uint32_t)e->mid_side : lv[e->level].ms) : 0u;
    o.flags = BNFLAC_ENC_NO_MD5;          // the running digest over all chunks is kept here
    e->pcm.clear(); e->md5 = bnfe::Md5(); e->frames = e->samples = 0; e->min_fs = 0xffffffffu; e->max_fs = 0;
    e->state = 0;
    return 0;
}
// encodes and emits the first `nbytes` of the collected PCM (whole blocks, or everything at finish)
static bool enc_flush(FLAC__StreamEncoder* e, size_t nbytes) {
    if (!nbytes) return true;
    const size_t spb = (size_t)e->channels * ((e->bps + 7) / 8), nsamp = nbytes / spb, bs = e->o.blocksize, nf = (nsamp + bs - 1) / bs;
    uint64_t bound = 0, n = 0;
    e->o.first_frame_number = e->frames;
    int rc = bnflac_encode_bound(nbytes, &e->o, &bound);
    if (!rc) {
        try { if (e->out.size() < bound) e->out.resize((size_t)bound); if (e->sizes.size() < nf) e->sizes.resize(nf); } catch (const std::bad_alloc&) { rc = BNFLAC_ERR_MEMORY; }
    }
    bnflac_enc_stats st{};
    if (!rc) { st.frame_sizes = e->sizes.data(); st.frame_sizes_cap = nf; rc = bnflac_encode(e->pcm.data(), nbytes, &e->o, e->out.data(), e->out.size(), &n, &st); }
    if (rc) { e->state = rc == BNFLAC_ERR_MEMORY ? 8 : 5; return false; }
    if (e->verify) {                       // set_verify: what was written must decode back to what was handed in
        bnflac_t* h = nullptr; bnflac_opts d{}; d.struct_size = sizeof d; d.device = -1; d.flags = BNFLAC_OPT_BORROW_INPUT;
        uint64_t w = 0;
        bool same = false;
        try {
            e->back.resize(nbytes + 64);
            same = !bnflac_open_memory(e->out.data(), (size_t)n, &d, &h) && !bnflac_decode_all(h, e->back.data(), e->back.size(), &w) && w == nbytes && !memcmp(e->back.data(), e->pcm.data(), nbytes);
        } catch (const std::bad_alloc&) {}
        if (h) bnflac_close(h);
        if (!same) { e->state = 4; return false; }      // FLAC__STREAM_ENCODER_VERIFY_MISMATCH_IN_AUDIO_DATA
    }
    e->md5.update(e->pcm.data(), nbytes);
    size_t at = 42;
    for (size_t f = 0; f < st.frames; f++) {
        const unsigned smp = (unsigned)std::min<size_t>(bs, nsamp - f * bs);
        if (!enc_emit(e, e->out.data() + at, e->sizes[f], smp, (unsigned)(e->frames + f))) return false;
        at += e->sizes[f];
    }
    e->frames += st.frames; e->samples += nsamp;
    if (st.frames) { e->min_fs = std::min(e->min_fs, st.min_framesize); e->max_fs = std::max(e->max_fs, st.max_framesize); }
    e->pcm.erase(e->pcm.begin(), e->pcm.begin() + (ptrdiff_t)nbytes);
    return true;
}
static bool enc_maybe_flush(FLAC__StreamEncoder* e) {
    const size_t fb = e->frame_bytes();
    if (e->pcm.size() < ENC_CHUNK_FRAMES * fb) return true;
    return enc_flush(e, e->pcm.size() / fb * fb);
}
static bool enc_begin_stream(FLAC__StreamEncoder* e) {           // libFLAC writes the marker and the (provisional) STREAMINFO at init
    uint8_t si[38]; const uint8_t zero[16] = {0};
    enc_streaminfo(e, si, zero);
    return enc_emit(e, (const uint8_t*)"fLaC", 4, 0, 0) && enc_emit(e, si, 38, 0, 0);
}

extern "C" {
FLAC__StreamEncoder* FLAC__stream_encoder_new(void) { return new (std::nothrow) FLAC__StreamEncoder(); }
void FLAC__stream_encoder_delete(FLAC__StreamEncoder* e) { if (e) { if (e->file) fclose(e->file); delete e; } }
FLAC__bool FLAC__stream_encoder_set_channels(FLAC__StreamEncoder* e, unsigned v) { if (!e || e->state != 1) return 0; e->channels = v; return 1; }
FLAC__bool FLAC__stream_encoder_set_bits_per_sample(FLAC__StreamEncoder* e, unsigned v) { if (!e || e->state != 1) return 0; e->bps = v; return 1; }
FLAC__bool FLAC__stream_encoder_set_sample_rate(FLAC__StreamEncoder* e, unsigned v) { if (!e || e->state != 1) return 0; e->sample_rate = v; return 1; }
FLAC__bool FLAC__stream_encoder_set_compression_level(FLAC__StreamEncoder* e, unsigned v) { if (!e || e->state != 1) return 0; e->level = v > 8 ? 8 : v; return 1; }
FLAC__bool FLAC__stream_encoder_set_blocksize(FLAC__StreamEncoder* e, unsigned v) { if (!e || e->state != 1) return 0; e->blocksize = v; return 1; }
FLAC__bool FLAC__stream_encoder_set_verify(FLAC__StreamEncoder* e, FLAC__bool v) { if (!e || e->state != 1) return 0; e->verify = v != 0; return 1; }
FLAC__bool FLAC__stream_encoder_set_streamable_subset(FLAC__StreamEncoder* e, FLAC__bool) { return e && e->state == 1; }
FLAC__bool FLAC__stream_encoder_set_do_mid_side_stereo(FLAC__StreamEncoder* e, FLAC__bool v) { if (!e || e->state != 1) return 0; e->mid_side = v ? 1 : 0; return 1; }
FLAC__bool FLAC__stream_encoder_set_loose_mid_side_stereo(FLAC__StreamEncoder* e, FLAC__bool) { return e && e->state == 1; }
int FLAC__stream_encoder_get_state(const FLAC__StreamEncoder* e) { return e ? e->state : 1; }

int FLAC__stream_encoder_init_stream(FLAC__StreamEncoder* e, FLAC__StreamEncoderWriteCallback write, FLAC__StreamEncoderSeekCallback seek, FLAC__StreamEncoderTellCallback,
                                     FLAC__StreamEncoderMetadataCallback metadata, void* client) {
    if (!e) return 1;
    if (!write) return 3;
    const int rc = enc_init_common(e);
    if (rc) return rc;
    e->write = write; e->seek = seek; e->metadata = metadata; e->client = client;
    return enc_begin_stream(e) ? 0 : 1;
}
int FLAC__stream_encoder_init_file(FLAC__StreamEncoder* e, const char* filename, void*, void* client) {
    if (!e) return 1;
    const int rc = enc_init_common(e);
    if (rc) return rc;
    e->file = filename ? fopen(filename, "wb") : nullptr;
    if (!e->file) { e->state = 6; return 1; }
    e->client = client;
    return enc_begin_stream(e) ? 0 : 1;
}
FLAC__bool FLAC__stream_encoder_process_interleaved(FLAC__StreamEncoder* e, const int32_t buffer[], unsigned samples) {
    if (!e || e->state != 0 || (!buffer && samples)) return 0;
    const unsigned B = (e->bps + 7) / 8;
    const size_t n = (size_t)samples * e->channels, at = e->pcm.size();
    try { e->pcm.resize(at + n * B); } catch (const std::bad_alloc&) { e->state = 8; return 0; }
    uint8_t* o = e->pcm.data() + at;
    for (size_t i = 0; i < n; i++) for (unsigned b = 0; b < B; b++) *o++ = (uint8_t)((uint32_t)buffer[i] >> (8 * b));
    return enc_maybe_flush(e) ? 1 : 0;
}
FLAC__bool FLAC__stream_encoder_process(FLAC__StreamEncoder* e, const int32_t* const buffer[], unsigned samples) {
    if (!e || e->state != 0 || (!buffer && samples)) return 0;
    const unsigned B = (e->bps + 7) / 8, C = e->channels;
    const size_t at = e->pcm.size();
    try { e->pcm.resize(at + (size_t)samples * C * B); } catch (const std::bad_alloc&) { e->state = 8; return 0; }
    uint8_t* o = e->pcm.data() + at;
    for (unsigned t = 0; t < samples; t++) for (unsigned c = 0; c < C; c++) for (unsigned b = 0; b < B; b++) *o++ = (uint8_t)((uint32_t)buffer[c][t] >> (8 * b));
    return enc_maybe_flush(e) ? 1 : 0;
}
FLAC__bool FLAC__stream_encoder_finish(FLAC__StreamEncoder* e) {
    if (!e) return 0;
    if (e->state == 1) return 1;                      // libFLAC: finishing an uninitialised encoder is a no-op
    bool ok = e->state == 0 && enc_flush(e, e->pcm.size());
    if (ok) {
        uint8_t digest[16], si[38];
        e->md5.final(digest);
        enc_streaminfo(e, si, digest);
        // the final STREAMINFO goes where the provisional one is, when the sink can seek (libFLAC does the same); the metadata
        // callback reports it either way
        if (e->file) { ok = fseek(e->file, 4, SEEK_SET) == 0 && fwrite(si, 1, 38, e->file) == 38 && fseek(e->file, 0, SEEK_END) == 0; if (!ok) e->state = 6; }
        else if (e->seek) { if (e->seek(e, 4, e->client) == 0) ok = enc_emit(e, si, 38, 0, 0); }
        if (ok && e->metadata) {
            FLAC__StreamMetadata m; memset(&m, 0, sizeof m);
            m.type = 0; m.is_last = 1; m.length = 34;
            m.stream_info.min_blocksize = m.stream_info.max_blocksize = e->o.blocksize;
            m.stream_info.min_framesize = e->frames ? e->min_fs : 0; m.stream_info.max_framesize = e->max_fs;
            m.stream_info.sample_rate = e->sample_rate; m.stream_info.channels = e->channels; m.stream_info.bits_per_sample = e->bps;
            m.stream_info.total_samples = e->samples; memcpy(m.stream_info.md5sum, digest, 16);
            e->metadata(e, &m, e->client);
        }
    }
    if (e->file) { if (fclose(e->file) != 0) ok = false; e->file = nullptr; }
    e->pcm.clear(); e->pcm.shrink_to_fit(); e->out.clear(); e->out.shrink_to_fit(); e->back.clear(); e->back.shrink_to_fit();
    if (ok) e->state = 1;                               // back to UNINITIALIZED, ready for another init (libFLAC)
    return ok ? 1 : 0;
}
}
