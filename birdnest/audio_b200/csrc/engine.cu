// engine.cu -- host runtime + C ABI of libbnflac.so (include/bnflac.h).
//
// Owns: metadata parse (host), device buffers, the kernel pipeline of kernels.cu, the Stream-style
// read buffering (FLACDecoder.Read, FLACDecoder.cs:124-205), frame-range sharding and the
// per-frame status / error vocabulary of the reference (LibFLACSharp.cs:24-36,262-268).
// There is no CPU decode path in this library: without a CUDA device every decode call fails.
#include "../../../include/bnflac.h"
#include "bnflac_dev.h"
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

using namespace bnf;

static thread_local std::string g_cuda_err;
#define CK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { g_cuda_err = std::string(#expr) + ": " + cudaGetErrorString(e_); return BNFLAC_ERR_CUDA; } } while (0)

// ------------------------------------------------------------------------------------------------ small helpers
namespace {

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { g_cuda_err = std::string("cudaMalloc: ") + cudaGetErrorString(e); p = nullptr; return BNFLAC_ERR_MEMORY; }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinBuf {
    void* p = nullptr; size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaHostAlloc(&p, n + 64, cudaHostAllocDefault);
        if (e != cudaSuccess) { g_cuda_err = std::string("cudaHostAlloc: ") + cudaGetErrorString(e); p = nullptr; return BNFLAC_ERR_MEMORY; }
        cap = n + 64;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// RFC 1321, used only by the optional BNFLAC_OPT_VERIFY_MD5 host check
struct Md5 {
    uint32_t s[4]; uint64_t n; uint8_t buf[64]; size_t fill;
    Md5() { s[0] = 0x67452301; s[1] = 0xefcdab89; s[2] = 0x98badcfe; s[3] = 0x10325476; n = 0; fill = 0; }
    static uint32_t rol(uint32_t v, int r) { return (v << r) | (v >> (32 - r)); }
    void block(const uint8_t* p) {
        static const uint32_t K[64] = {
            0xd76aa478,0xe8c7b756,0x242070db,0xc1bdceee,0xf57c0faf,0x4787c62a,0xa8304613,0xfd469501,0x698098d8,0x8b44f7af,0xffff5bb1,0x895cd7be,0x6b901122,0xfd987193,0xa679438e,0x49b40821,
            0xf61e2562,0xc040b340,0x265e5a51,0xe9b6c7aa,0xd62f105d,0x02441453,0xd8a1e681,0xe7d3fbc8,0x21e1cde6,0xc33707d6,0xf4d50d87,0x455a14ed,0xa9e3e905,0xfcefa3f8,0x676f02d9,0x8d2a4c8a,
            0xfffa3942,0x8771f681,0x6d9d6122,0xfde5380c,0xa4beea44,0x4bdecfa9,0xf6bb4b60,0xbebfbc70,0x289b7ec6,0xeaa127fa,0xd4ef3085,0x04881d05,0xd9d4d039,0xe6db99e5,0x1fa27cf8,0xc4ac5665,
            0xf4292244,0x432aff97,0xab9423a7,0xfc93a039,0x655b59c3,0x8f0ccc92,0xffeff47d,0x85845dd1,0x6fa87e4f,0xfe2ce6e0,0xa3014314,0x4e0811a1,0xf7537e82,0xbd3af235,0x2ad7d2bb,0xeb86d391};
        static const int R[64] = {7,12,17,22,7,12,17,22,7,12,17,22,7,12,17,22,5,9,14,20,5,9,14,20,5,9,14,20,5,9,14,20,
                                  4,11,16,23,4,11,16,23,4,11,16,23,4,11,16,23,6,10,15,21,6,10,15,21,6,10,15,21,6,10,15,21};
        uint32_t w[16]; memcpy(w, p, 64);
        uint32_t a = s[0], b = s[1], c = s[2], d = s[3];
        for (int i = 0; i < 64; i++) {
            uint32_t f; int g;
            if (i < 16) { f = (b & c) | (~b & d); g = i; }
            else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
            else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
            else { f = c ^ (b | ~d); g = (7 * i) & 15; }
            uint32_t t = a + f + K[i] + w[g];
            a = d; d = c; c = b; b = b + rol(t, R[i]);
        }
        s[0] += a; s[1] += b; s[2] += c; s[3] += d;
    }
    void update(const uint8_t* p, size_t len) {
        n += len;
        if (fill) { size_t t = std::min(len, 64 - fill); memcpy(buf + fill, p, t); fill += t; p += t; len -= t; if (fill == 64) { block(buf); fill = 0; } }
        for (; len >= 64; p += 64, len -= 64) block(p);
        if (len) { memcpy(buf, p, len); fill = len; }
    }
    void final(uint8_t out[16]) {
        uint64_t bits = n * 8; uint8_t pad[72] = {0x80};
        update(pad, fill < 56 ? 56 - fill : 120 - fill);
        uint8_t l[8]; for (int k = 0; k < 8; k++) l[k] = (uint8_t)(bits >> (8 * k));
        update(l, 8);
        for (int k = 0; k < 4; k++) for (int j = 0; j < 4; j++) out[4 * k + j] = (uint8_t)(s[k] >> (8 * j));
    }
};

// "fLaC" + metadata blocks; returns 0 or bnflac_err
int parse_metadata(const uint8_t* d, size_t len, bnflac_info_t* si) {
    size_t pos = 0;
    memset(si, 0, sizeof *si);
    if (len >= 10 && d[0] == 'I' && d[1] == 'D' && d[2] == '3')   // libFLAC skips a leading ID3v2 tag
        pos = 10 + (((size_t)(d[6] & 0x7f) << 21) | ((size_t)(d[7] & 0x7f) << 14) | ((size_t)(d[8] & 0x7f) << 7) | (d[9] & 0x7f));
    if (pos + 4 > len) return len >= 4 ? BNFLAC_ERR_NOT_FLAC : BNFLAC_ERR_TRUNCATED;
    if (memcmp(d + pos, "fLaC", 4)) return BNFLAC_ERR_NOT_FLAC;
    pos += 4;
    bool have = false;
    for (;;) {
        if (pos + 4 > len) return BNFLAC_ERR_TRUNCATED;
        int last = d[pos] >> 7, type = d[pos] & 0x7f;
        size_t l = (size_t)d[pos + 1] << 16 | (size_t)d[pos + 2] << 8 | d[pos + 3];
        pos += 4;
        if (pos + l > len) return BNFLAC_ERR_TRUNCATED;
        if (type == 0 && l >= 34 && !have) {
            const uint8_t* s = d + pos;
            si->min_blocksize = s[0] << 8 | s[1]; si->max_blocksize = s[2] << 8 | s[3];
            si->min_framesize = s[4] << 16 | s[5] << 8 | s[6]; si->max_framesize = s[7] << 16 | s[8] << 8 | s[9];
            uint64_t x = 0; for (int i = 10; i < 18; i++) x = x << 8 | s[i];
            si->sample_rate = (uint32_t)(x >> 44); si->channels = (uint32_t)((x >> 41) & 7) + 1;
            si->bits_per_sample = (uint32_t)((x >> 36) & 31) + 1; si->total_samples = x & 0xFFFFFFFFFull;
            memcpy(si->md5, s + 18, 16);
            have = true;
        }
        pos += l;
        if (last) break;
    }
    if (!have) return BNFLAC_ERR_NOT_FLAC;
    si->first_frame_offset = pos;
    si->bytes_per_sample = (si->bits_per_sample + 7) / 8;
    si->block_align = si->channels * (si->bits_per_sample / 8);                      // FLACDecoder.cs:448
    si->pcm_bytes = si->total_samples * si->channels * si->bytes_per_sample;
    // FLACDecoder.cs:449-450: (Hi << 32) on a 32-bit int is a no-op in C#, so only the low 32 bits of total_samples count
    si->length_reference = (uint64_t)si->block_align * (uint32_t)si->total_samples;
    si->duration_seconds = si->sample_rate ? (double)(uint32_t)si->total_samples / (double)si->sample_rate : 0.0;  // :452
    if (si->bits_per_sample == 16) si->al_format = si->channels == 2 ? BNFLAC_AL_STEREO16 : BNFLAC_AL_MONO16;  // :454-465
    else if (si->bits_per_sample == 8) si->al_format = si->channels == 2 ? BNFLAC_AL_STEREO8 : BNFLAC_AL_MONO8;
    else si->al_format = BNFLAC_AL_NONE;
    return 0;
}

uint32_t frame_bound(const bnflac_info_t& si) {
    uint32_t bs = si.max_blocksize ? si.max_blocksize : 65535;
    uint64_t b = 64 + (uint64_t)si.channels * ((uint64_t)bs * (si.bits_per_sample + 2) / 8 + 64);
    if (si.max_framesize > b) b = si.max_framesize;
    return (uint32_t)std::min<uint64_t>(b, 0x7fffffffu);
}

} // namespace

// ------------------------------------------------------------------------------------------------ the handle
struct bnflac {
    int device = 0;
    cudaStream_t stream = nullptr; bool own_stream = false;
    bnflac_opts opts{};
    bnflac_info_t info{};
    int state = BNFLAC_STATE_UNINITIALIZED;

    // input
    std::vector<uint8_t> host;          // copy of the stream when opened from host memory
    const uint8_t* host_ptr = nullptr;  // the bytes to upload (== host.data() unless BNFLAC_OPT_BORROW_INPUT)
    const uint8_t* d_ext = nullptr;     // caller-owned device copy (open_device)
    size_t len = 0;
    uint64_t slice_begin = 0, slice_end = 0;   // byte range of the whole stream this handle (shard) keeps on the device
    uint64_t own_begin = 0, own_end = 0;
    bool uploaded = false;

    // device state
    DevBuf d_in, d_segs, d_chunks, d_cand_tmp, d_cand, d_chunk_base, d_chunk_count, d_chunk_scan, d_counters, d_seg_crc, d_next, d_crc_tmp, d_chunk_head,
        d_flen, d_status, d_sub, d_pcm_off, d_acc_idx, d_totals, d_out;
    uint32_t nchunks = 0, cand_cap = 0;
    bool tables_ready = false;
    PassArgs args{};
    Totals totals{};
    uint32_t ncand = 0;
    cudaEvent_t ev[8] = {};
    bnflac_timing timing{};

    // results
    std::vector<bnflac_frame_t> frames; std::vector<bnflac_subframe_t> subs; std::vector<uint32_t> errors;
    bool diag_valid = false;

    // Stream-style read buffering
    PinBuf pcm_host; uint64_t pcm_len = 0, read_pos = 0; bool decoded = false;

    ~bnflac() {
        cudaSetDevice(device);
        DevBuf* all[] = {&d_in, &d_segs, &d_chunks, &d_cand_tmp, &d_cand, &d_chunk_base, &d_chunk_count, &d_chunk_scan, &d_counters, &d_seg_crc, &d_crc_tmp, &d_chunk_head,
                         &d_next, &d_flen, &d_status, &d_sub, &d_pcm_off, &d_acc_idx, &d_totals, &d_out};
        for (DevBuf* b : all) b->release();
        pcm_host.release();
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        if (own_stream && stream) cudaStreamDestroy(stream);
    }
};

static int setup_device(bnflac* h) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { g_cuda_err = "no CUDA device"; return BNFLAC_ERR_NO_DEVICE; }
    if (h->opts.device >= 0) h->device = h->opts.device; else if (cudaGetDevice(&h->device) != cudaSuccess) h->device = 0;
    if (h->device >= n) return BNFLAC_ERR_ARG;
    CK(cudaSetDevice(h->device));
    if (h->opts.stream) { h->stream = (cudaStream_t)h->opts.stream; h->own_stream = false; }
    else { CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }
    for (auto& e : h->ev) CK(cudaEventCreate(&e));
    return 0;
}

static void compute_shard(bnflac* h) {
    const uint64_t first = h->info.first_frame_offset, D = h->len - first;
    uint32_t n = h->opts.shard_count ? h->opts.shard_count : 1, i = std::min(h->opts.shard_index, n - 1);
    h->own_begin = first + D * i / n;
    h->own_end = (i + 1 == n) ? h->len : first + D * (i + 1) / n;
    h->slice_begin = h->own_begin & ~15ull;
    h->slice_end = (i + 1 == n) ? h->len : std::min<uint64_t>(h->len, h->own_end + frame_bound(h->info) + 32);
}

static int common_open(bnflac* h, const uint8_t* header, size_t header_len) {
    int rc = parse_metadata(header, header_len, &h->info);
    if (rc) return rc;
    if (h->info.channels > 8 || h->info.bits_per_sample > 24 || h->info.bits_per_sample < 4) return BNFLAC_ERR_UNSUPPORTED;
    compute_shard(h);
    h->state = BNFLAC_STATE_SEARCH_FOR_FRAME_SYNC;   // what libFLAC reports after process_until_end_of_metadata
    return 0;
}

static bnflac_opts default_opts(const bnflac_opts* o) {
    bnflac_opts d{}; d.struct_size = sizeof d; d.device = -1;
    if (o) { size_t n = std::min<size_t>(o->struct_size ? o->struct_size : sizeof d, sizeof d); memcpy(&d, o, n); d.struct_size = sizeof d; }
    return d;
}

// ------------------------------------------------------------------------------------------------ pipeline
static int ensure_input(bnflac* h) {
    if (h->uploaded) return 0;
    if (h->d_ext) { h->uploaded = true; return 0; }
    const size_t n = (size_t)(h->slice_end - h->slice_begin);
    int rc = h->d_in.reserve(n + 128); if (rc) return rc;
    CK(cudaMemcpyAsync(h->d_in.p, h->host_ptr + h->slice_begin, n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync((uint8_t*)h->d_in.p + n, 0, 128, h->stream));
    h->uploaded = true;
    return 0;
}

static int ensure_tables(bnflac* h) {
    if (h->tables_ready) return 0;
    // `in` addresses are relative to the start of what is on the device
    const uint64_t base = h->d_ext ? 0 : h->slice_begin;
    SegInfo seg{};
    seg.begin = std::max<uint64_t>(h->own_begin, h->info.first_frame_offset) - base;
    seg.end = h->slice_end - base;
    seg.own_begin = h->own_begin - base; seg.own_end = h->own_end - base;
    seg.bps = h->info.bits_per_sample; seg.channels = h->info.channels; seg.sample_rate = h->info.sample_rate;
    seg.min_bs = h->info.min_blocksize; seg.max_bs = h->info.max_blocksize; seg.max_frame_bytes = frame_bound(h->info);
    std::vector<Chunk> chunks;
    seg.first_chunk = 0;
    for (uint64_t p = seg.begin; p < seg.end;) {          // chunk boundaries at multiples of 32 KiB from the segment's aligned base
        const uint64_t stop = std::min<uint64_t>(seg.end, ((p & ~15ull) - ((p & ~15ull) - (seg.begin & ~15ull)) % SCAN_CHUNK) + SCAN_CHUNK);
        chunks.push_back(Chunk{p, (uint32_t)(stop - p), 0});
        p = stop;
    }
    h->nchunks = (uint32_t)chunks.size();
    int rc;
    if ((rc = h->d_segs.reserve(sizeof seg))) return rc;
    if ((rc = h->d_chunks.reserve(sizeof(Chunk) * std::max<size_t>(1, chunks.size())))) return rc;
    if ((rc = h->d_chunk_base.reserve(4ull * (h->nchunks + 1)))) return rc;
    if ((rc = h->d_chunk_count.reserve(4ull * (h->nchunks + 1)))) return rc;
    if ((rc = h->d_chunk_scan.reserve(4ull * (h->nchunks + 1)))) return rc;
    if ((rc = h->d_chunk_head.reserve(2ull * (h->nchunks + 1)))) return rc;
    if ((rc = h->d_counters.reserve(64))) return rc;
    if ((rc = h->d_totals.reserve(sizeof(Totals)))) return rc;
    CK(cudaMemcpyAsync(h->d_segs.p, &seg, sizeof seg, cudaMemcpyHostToDevice, h->stream));
    if (!chunks.empty()) CK(cudaMemcpyAsync(h->d_chunks.p, chunks.data(), sizeof(Chunk) * chunks.size(), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));   // chunks is a local
    h->args.in = h->d_ext ? h->d_ext : h->d_in.as<uint8_t>();
    h->args.in_len = (h->d_ext ? h->len : (h->slice_end - h->slice_begin)) + 64;
    h->args.segs = h->d_segs.as<SegInfo>(); h->args.nsegs = 1;
    h->args.chunks = h->d_chunks.as<Chunk>(); h->args.nchunks = h->nchunks;
    h->args.chunk_base = h->d_chunk_base.as<uint32_t>(); h->args.chunk_count = h->d_chunk_count.as<uint32_t>();
    h->args.chunk_scan = h->d_chunk_scan.as<uint32_t>(); h->args.counters = h->d_counters.as<uint32_t>();
    h->args.chunk_head = h->d_chunk_head.as<uint16_t>();
    h->args.totals = h->d_totals.as<Totals>();
    h->tables_ready = true;
    return 0;
}

static int reserve_cand(bnflac* h, uint32_t cap) {
    int rc;
    if ((rc = h->d_cand_tmp.reserve(sizeof(Cand) * (size_t)cap))) return rc;
    if ((rc = h->d_cand.reserve(sizeof(Cand) * (size_t)cap))) return rc;
    if ((rc = h->d_seg_crc.reserve(2ull * cap))) return rc;
    if ((rc = h->d_crc_tmp.reserve(2ull * cap))) return rc;
    if ((rc = h->d_next.reserve(4ull * cap))) return rc;
    if ((rc = h->d_flen.reserve(4ull * cap))) return rc;
    if ((rc = h->d_status.reserve(cap))) return rc;
    if ((rc = h->d_sub.reserve(sizeof(SubInfo) * MAX_CH * (size_t)cap))) return rc;
    if ((rc = h->d_pcm_off.reserve(8ull * cap))) return rc;
    if ((rc = h->d_acc_idx.reserve(4ull * cap))) return rc;
    h->cand_cap = cap;
    h->args.cand_tmp = h->d_cand_tmp.as<Cand>(); h->args.cand = h->d_cand.as<Cand>(); h->args.cand_cap = cap;
    h->args.seg_crc = h->d_seg_crc.as<uint16_t>(); h->args.crc_tmp = h->d_crc_tmp.as<uint16_t>(); h->args.next = h->d_next.as<uint32_t>(); h->args.flen = h->d_flen.as<uint32_t>();
    h->args.status = h->d_status.as<uint8_t>(); h->args.sub = h->d_sub.as<SubInfo>(); h->args.pcm_off = h->d_pcm_off.as<uint64_t>();
    h->args.acc_idx = h->d_acc_idx.as<uint32_t>();
    return 0;
}

// Runs K1..K2 + prefix (everything up to knowing the output size).
static int run_front(bnflac* h) {
    int rc;
    CK(cudaSetDevice(h->device));
    CK(cudaEventRecord(h->ev[0], h->stream));
    if ((rc = ensure_input(h))) return rc;
    if ((rc = ensure_tables(h))) return rc;
    if (!h->cand_cap) {
        uint64_t est = (h->slice_end - h->slice_begin) / 512 + 4096;
        if ((rc = reserve_cand(h, (uint32_t)std::min<uint64_t>(est, 0x7fffffff)))) return rc;
    }
    for (int attempt = 0;; attempt++) {
        CK(cudaMemsetAsync(h->d_counters.p, 0, 64, h->stream));
        CK(cudaMemsetAsync(h->d_totals.p, 0, sizeof(Totals), h->stream));
        launch_scan(h->args, h->stream);
        uint32_t counters[2];
        CK(cudaMemcpyAsync(counters, h->d_counters.p, 8, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (counters[1] & 1u) return BNFLAC_ERR_UNSUPPORTED;       // > SCAN_SCAP frame headers inside 32 KiB
        if (counters[0] > h->cand_cap) {
            if (attempt > 2) return BNFLAC_ERR_MEMORY;
            if ((rc = reserve_cand(h, counters[0] + counters[0] / 8 + 1024))) return rc;
            continue;
        }
        h->ncand = counters[0];
        break;
    }
    CK(cudaEventRecord(h->ev[1], h->stream));
    launch_order(h->args, h->stream);
    launch_crc(h->args, h->ncand, h->stream);
    CK(cudaEventRecord(h->ev[2], h->stream));
    launch_link(h->args, h->ncand, h->stream);
    CK(cudaEventRecord(h->ev[3], h->stream));
    launch_parse(h->args, h->ncand, h->stream);
    launch_prefix(h->args, h->info.bytes_per_sample, h->stream);
    CK(cudaEventRecord(h->ev[4], h->stream));
    CK(cudaMemcpyAsync(&h->totals, h->d_totals.p, sizeof(Totals), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    h->diag_valid = false;
    return 0;
}

static int run_back(bnflac* h, uint8_t* d_out, uint64_t cap) {
    if (h->totals.pcm_bytes > cap) return BNFLAC_ERR_CAPACITY;
    h->args.out = d_out; h->args.out_cap = cap;
    if (h->totals.n_accepted)
        launch_decode(h->args, h->totals.n_accepted, h->info.channels, h->info.bytes_per_sample, h->totals.max_order, h->totals.any_wide != 0, h->stream);
    CK(cudaEventRecord(h->ev[5], h->stream));
    CK(cudaGetLastError());
    return 0;
}

static int finish_timing(bnflac* h) {
    CK(cudaEventSynchronize(h->ev[5]));
    auto ms = [&](int a, int b) { float t = 0; cudaEventElapsedTime(&t, h->ev[a], h->ev[b]); return t; };
    h->timing.scan = ms(0, 1); h->timing.crc = ms(1, 2); h->timing.link = ms(2, 3); h->timing.parse = ms(3, 4);
    h->timing.decode = ms(4, 5); h->timing.total = ms(0, 5);
    h->timing.launches = 10;
    return 0;
}

static int decode_to_device(bnflac* h, void* d_dst, size_t cap, void** d_out, uint64_t* written) {
    int rc = run_front(h); if (rc) return rc;
    uint8_t* out = (uint8_t*)d_dst;
    if (!out) {
        if ((rc = h->d_out.reserve((size_t)h->totals.pcm_bytes + 64))) return rc;
        out = h->d_out.as<uint8_t>(); cap = h->d_out.cap;
    }
    if ((rc = run_back(h, out, cap))) return rc;
    if ((rc = finish_timing(h))) return rc;
    if (d_out) *d_out = out;
    if (written) *written = h->totals.pcm_bytes;
    h->state = BNFLAC_STATE_END_OF_STREAM;
    return 0;
}

static int fetch_diag(bnflac* h) {
    if (h->diag_valid) return 0;
    CK(cudaSetDevice(h->device));
    const uint32_t n = h->ncand;
    std::vector<Cand> cand(n); std::vector<uint8_t> st(n); std::vector<uint32_t> fl(n); std::vector<uint64_t> po(n); std::vector<SubInfo> sub((size_t)n * MAX_CH);
    if (n) {
        CK(cudaMemcpy(cand.data(), h->d_cand.p, sizeof(Cand) * n, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(st.data(), h->d_status.p, n, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(fl.data(), h->d_flen.p, 4ull * n, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(po.data(), h->d_pcm_off.p, 8ull * n, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(sub.data(), h->d_sub.p, sizeof(SubInfo) * MAX_CH * (size_t)n, cudaMemcpyDeviceToHost));
    }
    h->frames.clear(); h->subs.clear(); h->errors.clear();
    const uint64_t base = h->d_ext ? 0 : h->slice_begin;
    uint64_t expect = std::max<uint64_t>(h->own_begin, h->info.first_frame_offset);
    for (uint32_t i = 0; i < n; i++) {
        if (cand[i].flags & 2) continue;
        if (st[i] == ST_UNPARSEABLE) { h->errors.push_back(3); continue; }
        if (st[i] != ST_OK && st[i] != ST_CRC) continue;
        bnflac_frame_t f{};
        f.offset = cand[i].off + base; f.length = fl[i]; f.blocksize = cand[i].bs;
        f.channels = (uint8_t)(cand[i].assign < 8 ? cand[i].assign + 1 : 2); f.bits_per_sample = cand[i].bps; f.assignment = cand[i].assign;
        f.status = st[i] == ST_OK ? BNFLAC_FRAME_OK : BNFLAC_FRAME_CRC_MISMATCH;
        f.number = cand[i].number; f.pcm_offset = po[i];
        if (f.offset != expect) {
            // bytes were skipped before this frame.  The reference reports BAD_HEADER first when the skipped bytes begin
            // with a sync code whose header did not validate, then LOST_SYNC (observed on the DLL, tests/golden faults)
            uint8_t two[2] = {0, 0};
            if (expect + 2 <= h->len) {
                if (h->host_ptr) memcpy(two, h->host_ptr + expect, 2);
                else cudaMemcpy(two, h->d_ext + expect, 2, cudaMemcpyDeviceToHost);
            }
            if (two[0] == 0xFF && (two[1] & 0xFC) == 0xF8) h->errors.push_back(1);
            h->errors.push_back(0);
        }
        if (st[i] == ST_CRC) h->errors.push_back(2);              // FRAME_CRC_MISMATCH
        expect = f.offset + f.length;
        h->frames.push_back(f);
        for (int c = 0; c < MAX_CH; c++) {
            bnflac_subframe_t s{};
            if (c < f.channels && st[i] == ST_OK) { const SubInfo& si = sub[(size_t)i * MAX_CH + c]; s.bit_offset = si.bit_offset; s.type = si.type; s.order = si.order; s.wasted = si.wasted; s.flags = si.flags; }
            h->subs.push_back(s);
        }
    }
    h->diag_valid = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" {

int bnflac_abi_version(void) { return BNFLAC_ABI_VERSION; }
int bnflac_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }
const char* bnflac_last_cuda_error(void) { return g_cuda_err.c_str(); }
uint64_t bnflac_kernel_launches(void) { return (uint64_t)kernel_launch_count(); }

const char* bnflac_strerror(int err) {
    switch (err) {
    case BNFLAC_OK: return "OK";
    case BNFLAC_ERR_ARG: return "bad argument";
    case BNFLAC_ERR_NOT_FLAC: return "not a FLAC stream (no fLaC marker / STREAMINFO)";
    case BNFLAC_ERR_TRUNCATED: return "stream truncated inside the metadata";
    case BNFLAC_ERR_NO_DEVICE: return "no CUDA device (this engine has no CPU fallback)";
    case BNFLAC_ERR_CUDA: return "CUDA runtime error";
    case BNFLAC_ERR_MEMORY: return "MemoryAllocationError";
    case BNFLAC_ERR_CAPACITY: return "destination buffer too small";
    case BNFLAC_ERR_ABORTED: return "Aborted";
    case BNFLAC_ERR_UNSUPPORTED: return "stream shape outside engine limits";
    case BNFLAC_ERR_STATE: return "call not valid in this state";
    default: return "unknown error";
    }
}
const char* bnflac_state_name(int s) {
    static const char* n[] = {"SearchForMetadata", "ReadMetadata", "SearchForFrameSync", "ReadFrame", "EndOfStream", "OggError", "SeekError", "Aborted", "MemoryAllocationError", "Uninitialized"};
    return (s >= 0 && s < 10) ? n[s] : "Unknown";
}
const char* bnflac_frame_status_name(int s) {
    static const char* n[] = {"Ok", "LostSync", "BadHeader", "FrameCrcMismatch", "UnparsableStream"};
    return (s >= 0 && s < 5) ? n[s] : "Unknown";
}

int bnflac_open_memory(const uint8_t* data, size_t len, const bnflac_opts* opts, bnflac_t** out) {
    if (!data || !out) return BNFLAC_ERR_ARG;
    *out = nullptr;
    bnflac* h = new (std::nothrow) bnflac; if (!h) return BNFLAC_ERR_MEMORY;
    h->opts = default_opts(opts); h->len = len;
    bnflac_info_t probe; int rc = parse_metadata(data, len, &probe);      // fail before touching the device
    if (rc) { delete h; return rc; }
    if (h->opts.flags & BNFLAC_OPT_BORROW_INPUT) h->host_ptr = data;
    else {
        try { h->host.assign(data, data + len); } catch (...) { delete h; return BNFLAC_ERR_MEMORY; }
        h->host_ptr = h->host.data();
    }
    if ((rc = common_open(h, h->host_ptr, len)) || (rc = setup_device(h))) { delete h; return rc; }
    *out = h;
    return 0;
}

int bnflac_open_callbacks(bnflac_read_cb read, void* user, const bnflac_opts* opts, bnflac_t** out) {
    if (!read || !out) return BNFLAC_ERR_ARG;
    *out = nullptr;
    std::vector<uint8_t> buf;
    const size_t req = 1u << 20;   // the reference pulls <= 16 KiB per callback (FLACDecoder.cs:21,336); we ask for 1 MiB
    try {
        for (;;) {
            size_t old = buf.size(); buf.resize(old + req);
            size_t got = req;
            int st = read(user, buf.data() + old, &got);
            if (st == 2) return BNFLAC_ERR_ABORTED;
            if (got > req) return BNFLAC_ERR_ARG;
            buf.resize(old + got);
            if (st == 1 || got == 0) break;
        }
    } catch (...) { return BNFLAC_ERR_MEMORY; }
    return bnflac_open_memory(buf.data(), buf.size(), opts, out);
}

int bnflac_open_device(const void* d_data, size_t len, const uint8_t* header, size_t header_len, const bnflac_opts* opts, bnflac_t** out) {
    if (!d_data || !header || !out) return BNFLAC_ERR_ARG;
    *out = nullptr;
    bnflac* h = new (std::nothrow) bnflac; if (!h) return BNFLAC_ERR_MEMORY;
    h->opts = default_opts(opts); h->len = len; h->d_ext = (const uint8_t*)d_data;
    int rc;
    if ((rc = common_open(h, header, std::min(header_len, len))) || (rc = setup_device(h))) { delete h; return rc; }
    *out = h;
    return 0;
}

int bnflac_info(bnflac_t* h, bnflac_info_t* info) { if (!h || !info) return BNFLAC_ERR_ARG; *info = h->info; return 0; }
int bnflac_state(bnflac_t* h) { return h ? h->state : BNFLAC_STATE_UNINITIALIZED; }
void bnflac_close(bnflac_t* h) { delete h; }

int bnflac_decode_device(bnflac_t* h, void* d_dst, size_t cap, void** d_out, uint64_t* written) {
    if (!h) return BNFLAC_ERR_ARG;
    return decode_to_device(h, d_dst, cap, d_out, written);
}

int bnflac_decoded_size(bnflac_t* h, uint64_t* bytes) {
    if (!h || !bytes) return BNFLAC_ERR_ARG;
    int rc = run_front(h); if (rc) return rc;
    *bytes = h->totals.pcm_bytes;
    return 0;
}

int bnflac_decode_all(bnflac_t* h, uint8_t* dst, size_t cap, uint64_t* written) {
    if (!h || !dst) return BNFLAC_ERR_ARG;
    int rc = run_front(h); if (rc) return rc;
    if (h->totals.pcm_bytes > cap) return BNFLAC_ERR_CAPACITY;
    if ((rc = h->d_out.reserve((size_t)h->totals.pcm_bytes + 64))) return rc;
    if ((rc = run_back(h, h->d_out.as<uint8_t>(), h->d_out.cap))) return rc;
    CK(cudaMemcpyAsync(dst, h->d_out.p, (size_t)h->totals.pcm_bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if ((rc = finish_timing(h))) return rc;
    if (written) *written = h->totals.pcm_bytes;
    h->state = BNFLAC_STATE_END_OF_STREAM;
    if (h->opts.flags & BNFLAC_OPT_VERIFY_MD5) {
        static const uint8_t zero[16] = {0};
        if (memcmp(h->info.md5, zero, 16) && (h->opts.shard_count <= 1)) {
            Md5 m; m.update(dst, (size_t)h->totals.pcm_bytes); uint8_t d[16]; m.final(d);
            if (memcmp(d, h->info.md5, 16)) return BNFLAC_ERR_STATE;
        }
    }
    return 0;
}

int64_t bnflac_read(bnflac_t* h, uint8_t* dst, size_t count) {
    if (!h || (!dst && count)) return BNFLAC_ERR_ARG;
    if (!h->decoded) {
        int rc = run_front(h); if (rc) return rc;
        if ((rc = h->d_out.reserve((size_t)h->totals.pcm_bytes + 64))) return rc;
        if ((rc = h->pcm_host.reserve((size_t)h->totals.pcm_bytes))) return rc;
        if ((rc = run_back(h, h->d_out.as<uint8_t>(), h->d_out.cap))) return rc;
        CK(cudaMemcpyAsync(h->pcm_host.p, h->d_out.p, (size_t)h->totals.pcm_bytes, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if ((rc = finish_timing(h))) return rc;
        h->pcm_len = h->totals.pcm_bytes; h->read_pos = 0; h->decoded = true;
        h->state = BNFLAC_STATE_READ_FRAME;
    }
    uint64_t left = h->pcm_len - h->read_pos;
    size_t n = (size_t)std::min<uint64_t>(left, count);
    if (n) memcpy(dst, (const uint8_t*)h->pcm_host.p + h->read_pos, n);
    h->read_pos += n;
    if (h->read_pos == h->pcm_len) h->state = BNFLAC_STATE_END_OF_STREAM;
    return (int64_t)n;
}

int bnflac_decode_batch(const bnflac_span* clips, size_t n, const bnflac_opts* opts, uint8_t* dst, size_t cap, int dst_is_device,
                        bnflac_clip_result* results, uint64_t* written) {
    (void)clips; (void)n; (void)opts; (void)dst; (void)cap; (void)dst_is_device; (void)results; (void)written;
    return BNFLAC_ERR_UNSUPPORTED;
}

int bnflac_frames(bnflac_t* h, const bnflac_frame_t** frames, size_t* n) {
    if (!h || !frames || !n) return BNFLAC_ERR_ARG;
    int rc = fetch_diag(h); if (rc) return rc;
    *frames = h->frames.data(); *n = h->frames.size();
    return 0;
}
int bnflac_subframes(bnflac_t* h, const bnflac_subframe_t** sub, size_t* n) {
    if (!h || !sub || !n) return BNFLAC_ERR_ARG;
    int rc = fetch_diag(h); if (rc) return rc;
    *sub = h->subs.data(); *n = h->subs.size();
    return 0;
}
int bnflac_errors(bnflac_t* h, const uint32_t** codes, size_t* n) {
    if (!h || !codes || !n) return BNFLAC_ERR_ARG;
    int rc = fetch_diag(h); if (rc) return rc;
    *codes = h->errors.data(); *n = h->errors.size();
    return 0;
}
int bnflac_last_timing(bnflac_t* h, bnflac_timing* t) { if (!h || !t) return BNFLAC_ERR_ARG; *t = h->timing; return 0; }

} // extern "C"
