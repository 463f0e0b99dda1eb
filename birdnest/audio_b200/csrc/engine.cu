// engine.cu -- host runtime + C ABI of libbnflac.so (include/bnflac.h).
//
// Owns: metadata parse (host), device buffers, the kernel pipeline of kernels.cu, the Stream-style
// read buffering (FLACDecoder.Read, FLACDecoder.cs:124-205), frame-range sharding and the
// per-frame status / error vocabulary of the reference (LibFLACSharp.cs:24-36,262-268).
// There is no CPU decode path in this library: without a CUDA device every decode call fails.
#if defined(__x86_64__)
#include <emmintrin.h>
#endif
#include "../../../include/bnflac.h"
#include "bnflac_dev.h"
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <utility>
#include <vector>
#include <sys/mman.h>

using namespace bnf;

static thread_local std::string g_cuda_err;
namespace bnf { void set_cuda_error(const char* what) { g_cuda_err = what; } }   // for encoder.cu (same thread-local text behind bnflac_last_cuda_error)
#define CK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { g_cuda_err = std::string(#expr) + ": " + cudaGetErrorString(e_); return BNFLAC_ERR_CUDA; } } while (0)

// ------------------------------------------------------------------------------------------------ small helpers
namespace {

// Every entry point of the C ABI works on the handle's device and puts the caller's current device back when it returns
// (a host that drives several GPUs from one thread must not find its device switched under it).
struct DeviceScope {
    int prev = -1; bool changed = false;
    explicit DeviceScope(int dev) {
        if (dev < 0 || cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); return; }
        if (prev != dev && cudaSetDevice(dev) == cudaSuccess) changed = true;
    }
    ~DeviceScope() { if (changed) cudaSetDevice(prev); }
    DeviceScope(const DeviceScope&) = delete; DeviceScope& operator=(const DeviceScope&) = delete;
};

// Process-wide cache of device and pinned-host blocks (the only global mutable state, mutex-guarded).  cudaMalloc /
// cudaFree / cudaHostAlloc cost milliseconds per gigabyte and cudaFree synchronises the device, which would serialise
// the pipelined host decode; handles therefore take their buffers from here and give them back on close.
struct BlockPool {
    struct Blk { void* p; size_t n; int dev; };     // dev < 0: pinned host block
    std::mutex m;
    std::vector<Blk> free_;
    static constexpr size_t MAX_CACHED = 2048;
    void* take(size_t n, int dev, size_t* got) {
        std::lock_guard<std::mutex> g(m);
        size_t best = (size_t)-1;
        for (size_t i = 0; i < free_.size(); i++)
            if (free_[i].dev == dev && free_[i].n >= n && free_[i].n <= 2 * n + (1u << 20) && (best == (size_t)-1 || free_[i].n < free_[best].n)) best = i;
        if (best == (size_t)-1) return nullptr;
        void* p = free_[best].p; *got = free_[best].n;
        free_[best] = free_.back(); free_.pop_back();
        return p;
    }
    static void destroy(const Blk& b) {
        if (b.dev < 0) cudaFreeHost(b.p);
        else { int cur = 0; cudaGetDevice(&cur); if (cur != b.dev) cudaSetDevice(b.dev); cudaFree(b.p); if (cur != b.dev) cudaSetDevice(cur); }
    }
    void give(void* p, size_t n, int dev) {
        Blk drop{nullptr, 0, 0};
        {
            std::lock_guard<std::mutex> g(m);
            free_.push_back(Blk{p, n, dev});
            if (free_.size() > MAX_CACHED) {       // drop the smallest block
                size_t k = 0;
                for (size_t i = 1; i < free_.size(); i++) if (free_[i].n < free_[k].n) k = i;
                drop = free_[k]; free_[k] = free_.back(); free_.pop_back();
            }
        }
        if (drop.p) destroy(drop);
    }
    void trim() {
        std::vector<Blk> all;
        { std::lock_guard<std::mutex> g(m); all.swap(free_); }
        for (const Blk& b : all) destroy(b);
    }
};
BlockPool g_pool;

struct DevBuf {
    void* p = nullptr; size_t cap = 0; int dev = 0;
    int reserve(size_t n) {
        if (n <= cap) return 0;
        release();
        cudaGetDevice(&dev);
        size_t want = ((n + n / 8 + 256) + 255) & ~(size_t)255;
        if ((p = g_pool.take(want, dev, &cap))) return 0;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { cudaGetLastError(); g_pool.trim(); e = cudaMalloc(&p, want); }
        if (e != cudaSuccess) { g_cuda_err = std::string("cudaMalloc: ") + cudaGetErrorString(e); cudaGetLastError(); p = nullptr; return BNFLAC_ERR_MEMORY; }
        cap = want;
        return 0;
    }
    void release() { if (p) g_pool.give(p, cap, dev); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinBuf {
    void* p = nullptr; size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return 0;
        release();
        size_t want = ((n + 64) + 4095) & ~(size_t)4095;
        if ((p = g_pool.take(want, -1, &cap))) return 0;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocMapped | cudaHostAllocPortable);
        if (e != cudaSuccess) { g_cuda_err = std::string("cudaHostAlloc: ") + cudaGetErrorString(e); cudaGetLastError(); p = nullptr; return BNFLAC_ERR_MEMORY; }
        cap = want;
        return 0;
    }
    void release() { if (p) g_pool.give(p, cap, -1); p = nullptr; cap = 0; }
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
// One `bnflac` is one pipeline pass over one device-resident run of bytes: a whole stream, one frame-range shard of it,
// or (batch) the clips of one format group laid back to back.  A handle that decodes a large host-resident stream to
// host memory owns `kids`: sub-shards of its own range that are uploaded, decoded and downloaded on their own streams
// so that the PCIe transfers of one overlap the kernels of another (decode_host_pipelined).
struct SegDesc {            // host description of one segment of a pass, in coordinates of the uploaded bytes
    uint64_t begin, end, own_begin, own_end;
    uint32_t sample_rate, min_bs, max_bs, max_frame_bytes;
};

// Bytes pulled from a read callback: one contiguous buffer that grows IN PLACE.  Address space is reserved up front
// (anonymous, MAP_NORESERVE: pages exist once they are written) and extended with mremap when it runs out, so appending
// never copies or zero-fills what is already there -- a std::vector grown 1 MiB at a time spent 25-55 ms on page faults
// and copies for the first 19 MiB, ten times the decode of those bytes.
struct GrowBuf {
    uint8_t* p = nullptr; size_t len = 0, cap = 0;
    GrowBuf() = default;
    GrowBuf(const GrowBuf&) = delete; GrowBuf& operator=(const GrowBuf&) = delete;
    ~GrowBuf() { release(); }
    void release() { if (p) munmap(p, cap); p = nullptr; len = cap = 0; }
    bool room(size_t more) {                              // make [len, len + more) writable
        if (len + more <= cap) return true;
        size_t want = cap ? cap : (size_t)1 << 32;        // 4 GiB of address space to start with
        while (want < len + more) want <<= 1;
        void* q;
        if (!p) {
            q = mmap(nullptr, want, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            while (q == MAP_FAILED && want / 2 >= len + more && want > ((size_t)1 << 24)) {      // address space limited (ulimit -v)
                want >>= 1;
                q = mmap(nullptr, want, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            }
        } else q = mremap(p, cap, want, MREMAP_MAYMOVE);  // moves page tables, not bytes
        if (q == MAP_FAILED) return false;
        p = (uint8_t*)q; cap = want;
#ifdef MADV_HUGEPAGE
        madvise(p, cap, MADV_HUGEPAGE);                   // first-touch faults per 2 MiB instead of per 4 KiB where the kernel allows it
#endif
        return true;
    }
};

struct DiagCursor {          // the reference's sync-search state, carried across the sub-shards of a pipelined decode
    uint64_t expect = 0;     // where its cursor stands
    bool in_sync = true;     // LOST_SYNC is reported once per excursion
    bool ended = false;      // a truncated header ended the stream
    bool have_expect = true; // false: (shard > 0) start at the first frame found, whatever lies before it
    std::vector<uint32_t>* at = nullptr;                 // per event: how many frames had been delivered before it
    const std::vector<bnflac_frame_t>* frames = nullptr;
    void event(std::vector<uint32_t>& errors, uint32_t code) { errors.push_back(code); if (at) at->push_back(frames ? (uint32_t)frames->size() : 0u); }
};

struct bnflac {
    int device = 0;
    cudaStream_t stream = nullptr; bool own_stream = false;
    bnflac_opts opts{};
    bnflac_info_t info{};
    int state = BNFLAC_STATE_UNINITIALIZED;

    // input
    std::vector<uint8_t> host;          // copy of the stream when opened from host memory
    GrowBuf pulled;                     // ... or what has been pulled from the read callback so far
    const uint8_t* host_ptr = nullptr;  // the bytes to upload (== host.data() unless BNFLAC_OPT_BORROW_INPUT)
    const uint8_t* d_ext = nullptr;     // caller-owned device copy (open_device)
    size_t len = 0;
    uint64_t slice_begin = 0, slice_end = 0;   // byte range of the whole stream this handle (shard) keeps on the device
    uint64_t own_begin = 0, own_end = 0;
    uint64_t sub_begin = 0, sub_end = 0;       // (kid) explicit sub-range of the parent's own range; 0,0 = none
    bool uploaded = false;
    std::vector<SegDesc> batch_segs;           // batch passes: segments given explicitly (coordinates of d_in)

    // device state
    DevBuf d_in, d_segs, d_chunks, d_cand_tmp, d_cand, d_chunk_base, d_chunk_count, d_chunk_scan, d_counters, d_seg_crc, d_next, d_pref, d_anom,
        d_flen, d_status, d_sub, d_pcm_off, d_acc_idx, d_totals, d_out, d_seg_pcm, d_seg_flags, d_spec_jobs, d_spec_base, d_spec_count, d_spec_done, d_acc_sorted, d_bucket_hist, d_dec_sched;
    uint32_t nchunks = 0, cand_cap = 0, nsegs = 0;
    bool tables_ready = false;
    PassArgs args{};
    Totals totals{};
    uint32_t ncand = 0;
    struct Predict { bool valid = false; uint32_t ncand = 0, n_accepted = 0, max_order = 0, any_wide = 0; uint64_t pcm_bytes = 0; } pred;   // what the last pass found
    cudaEvent_t ev[10] = {};             // 0..5 stage boundaries, 6/7 pipelined start/end, 8 upload done
    cudaStream_t up_stream = nullptr;    // (pipelined parent) all sub-shard uploads, in order
    bnflac_timing timing{};
    int launches0 = 0;

    PinBuf mailbox;                     // mapped pinned words the kernels publish counters / totals into
    // pipelined host decode
    std::vector<bnflac*> kids;
    uint64_t pcm_base = 0;              // (kid) where this sub-shard's PCM starts in the parent's output

    // results
    std::vector<bnflac_frame_t> frames; std::vector<bnflac_subframe_t> subs; std::vector<uint32_t> errors, errors_at;
    bool diag_valid = false;
    DiagCursor diag_cur; uint32_t diag_kids = 0; bool diag_started = false;   // streamed Read: sub-shards already folded into the tables

    // Stream-style read buffering
    PinBuf pcm_host; uint64_t pcm_len = 0, read_pos = 0; bool decoded = false;
    // lazily pulled source (BNFLAC_OPT_LAZY_PULL): the read callback is called as the reader advances, see pull_more()
    bnflac_read_cb pull_cb = nullptr; void* pull_user = nullptr; bool pull_eof = true;
    uint64_t pl_next = 0, pl_size = 0;   // start of the next sub-shard to issue, its size
    bool pl_session = false;             // the current streamed Read session cuts its sub-shards as the bytes arrive
    bool pl_ahead = false;               // ... and has started to issue sub-shards ahead of the reader
    bool front_ran = false;              // a decode has been started on this handle (the diagnostics have something to describe)
    bool mixed_blocksizes = false;       // variable-blocksize stream, or a batch of clips of unlike blocksizes (set at open)
    bool size_understated = false;       // a decode ran out of room although the buffer held what STREAMINFO promises: size by scanning
    // streaming Read session: sub-shards decoded ahead of the reader, see stream_read()
    bool rd_active = false; uint32_t rd_issued = 0, rd_cur = 0; uint64_t rd_off = 0, rd_total = 0;
    double tr_pull = 0, tr_issue = 0, tr_wait = 0, tr_copy = 0; uint64_t tr_reads = 0;     // BNFLAC_TRACE: where a streamed Read session spent its host time (ms)

    ~bnflac() {
        if (tr_reads && getenv("BNFLAC_TRACE"))
            fprintf(stderr, "[bnflac] streamed Read session: %llu reads; host time: pulling the source %.1f ms, issuing sub-shards %.1f ms (of which pulls), waiting for downloads %.1f ms, copying PCM out %.1f ms\n",
                    (unsigned long long)tr_reads, tr_pull, tr_issue, tr_wait, tr_copy);
        for (bnflac* k : kids) delete k;
        DeviceScope on(device);
        if (stream) cudaStreamSynchronize(stream);     // buffers go back to the shared pool: nothing may still be using them
        DevBuf* all[] = {&d_in, &d_segs, &d_chunks, &d_cand_tmp, &d_cand, &d_chunk_base, &d_chunk_count, &d_chunk_scan, &d_counters, &d_seg_crc, &d_pref, &d_anom,
                         &d_next, &d_flen, &d_status, &d_sub, &d_pcm_off, &d_acc_idx, &d_totals, &d_out, &d_seg_pcm, &d_seg_flags, &d_spec_jobs, &d_spec_base, &d_spec_count, &d_spec_done, &d_acc_sorted, &d_bucket_hist, &d_dec_sched};
        for (DevBuf* b : all) b->release();
        pcm_host.release(); mailbox.release();
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        if (own_stream && stream) cudaStreamDestroy(stream);
        if (up_stream) cudaStreamDestroy(up_stream);
    }
};

static int setup_device(bnflac* h) {       // leaves the handle's device current: callers hold a DeviceScope
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); g_cuda_err = "no CUDA device"; return BNFLAC_ERR_NO_DEVICE; }
    if (h->opts.device >= 0) h->device = h->opts.device; else if (cudaGetDevice(&h->device) != cudaSuccess) h->device = 0;
    if (h->device >= n) return BNFLAC_ERR_ARG;
    CK(cudaSetDevice(h->device));
    if (h->opts.stream) { h->stream = (cudaStream_t)h->opts.stream; h->own_stream = false; }
    else { CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)); h->own_stream = true; }
    for (auto& e : h->ev) CK(cudaEventCreate(&e));
    return 0;
}

// byte range of frame data owned by shard i of n; exported as bnflac_shard_range
static void shard_range(uint64_t len, uint64_t first, uint32_t i, uint32_t n, uint64_t* b, uint64_t* e) {
    if (!n) n = 1;
    if (i >= n) i = n - 1;
    if (first > len) first = len;
    const uint64_t D = len - first;
    // D * i can pass 2^64 only for streams beyond 2^32 shards x bytes: split the product
    *b = first + (D / n) * i + (D % n) * i / n;
    *e = (i + 1 == n) ? len : first + (D / n) * (i + 1) + (D % n) * (i + 1) / n;
}
static void compute_shard(bnflac* h) {
    uint64_t b, e;
    shard_range(h->len, h->info.first_frame_offset, h->opts.shard_index, h->opts.shard_count, &b, &e);
    if (h->sub_end > h->sub_begin) { b = h->sub_begin; e = h->sub_end; }      // sub-shard of the own range
    h->own_begin = b; h->own_end = e;
    h->slice_begin = b & ~15ull;
    h->slice_end = (e >= h->len) ? h->len : std::min<uint64_t>(h->len, e + frame_bound(h->info) + 32);
}

static int common_open(bnflac* h, const uint8_t* header, size_t header_len) {
    int rc = parse_metadata(header, header_len, &h->info);
    if (rc) return rc;
    if (h->info.channels > 8 || h->info.bits_per_sample > 24 || h->info.bits_per_sample < 4) return BNFLAC_ERR_UNSUPPORTED;
    compute_shard(h);
    h->mixed_blocksizes = h->info.min_blocksize != h->info.max_blocksize;
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
    int rc = h->d_in.reserve(n + 128); if (rc) return rc;       // (no-op when already reserved)
    CK(cudaMemcpyAsync(h->d_in.p, h->host_ptr + h->slice_begin, n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync((uint8_t*)h->d_in.p + n, 0, 128, h->stream));
    h->uploaded = true;
    return 0;
}

static int ensure_tables(bnflac* h) {
    if (h->tables_ready) return 0;
    std::vector<SegDesc> descs;
    uint64_t in_len;
    if (!h->batch_segs.empty()) { descs = h->batch_segs; in_len = h->len; }
    else {
        // `in` addresses are relative to the start of what is on the device
        const uint64_t base = h->d_ext ? 0 : h->slice_begin;
        SegDesc d{};
        d.begin = std::max<uint64_t>(h->own_begin, h->info.first_frame_offset) - base;
        d.end = h->slice_end - base;
        d.own_begin = h->own_begin - base; d.own_end = h->own_end - base;
        d.sample_rate = h->info.sample_rate; d.min_bs = h->info.min_blocksize; d.max_bs = h->info.max_blocksize; d.max_frame_bytes = frame_bound(h->info);
        descs.push_back(d);
        in_len = h->d_ext ? h->len : (h->slice_end - h->slice_begin);
    }
    if (h->batch_segs.empty()) {
        // one segment: the tile descriptors are generated on the device (k_make_chunks), nothing is built or uploaded here
        const SegDesc& d = descs[0];
        SegInfo seg{};
        seg.begin = d.begin; seg.end = d.end; seg.own_begin = d.own_begin; seg.own_end = d.own_end;
        seg.bps = h->info.bits_per_sample; seg.channels = h->info.channels; seg.sample_rate = d.sample_rate;
        seg.min_bs = d.min_bs; seg.max_bs = d.max_bs; seg.max_frame_bytes = d.max_frame_bytes; seg.first_chunk = 0; seg.pad = 0;
        const uint64_t A = seg.begin & ~15ull;
        h->nchunks = seg.end > seg.begin ? (uint32_t)((seg.end - A + SCAN_CHUNK - 1) / SCAN_CHUNK) : 0u;
        h->nsegs = 1;
        int rc;
        if ((rc = h->d_segs.reserve(sizeof(SegInfo)))) return rc;
        if ((rc = h->d_chunks.reserve(sizeof(Chunk) * std::max<size_t>(1, h->nchunks)))) return rc;
        if ((rc = h->d_chunk_base.reserve(4ull * (h->nchunks + 1)))) return rc;
        if ((rc = h->d_chunk_count.reserve(4ull * (h->nchunks + 1)))) return rc;
        if ((rc = h->d_chunk_scan.reserve(4ull * (h->nchunks + 1)))) return rc;
        if ((rc = h->d_pref.reserve(64ull * (h->nchunks + 1)))) return rc;
        if ((rc = h->d_counters.reserve(4 * CNT_WORDS))) return rc;
        if ((rc = h->d_totals.reserve(sizeof(Totals)))) return rc;
        if ((rc = h->mailbox.reserve(256))) return rc;
        if ((rc = h->d_anom.reserve(4ull * ANOM_CAP))) return rc;
        launch_make_chunks(seg, h->d_segs.as<SegInfo>(), h->d_chunks.as<Chunk>(), h->nchunks, h->stream);
        CK(cudaGetLastError());
    } else {
    std::vector<SegInfo> segs(descs.size());
    std::vector<Chunk> chunks;
    for (size_t k = 0; k < descs.size(); k++) {
        const SegDesc& d = descs[k];
        SegInfo& seg = segs[k];
        seg.begin = d.begin; seg.end = d.end; seg.own_begin = d.own_begin; seg.own_end = d.own_end;
        seg.bps = h->info.bits_per_sample; seg.channels = h->info.channels; seg.sample_rate = d.sample_rate;
        seg.min_bs = d.min_bs; seg.max_bs = d.max_bs; seg.max_frame_bytes = d.max_frame_bytes;
        seg.first_chunk = (uint32_t)chunks.size(); seg.pad = 0;
        for (uint64_t p = seg.begin; p < seg.end;) {          // chunk boundaries at multiples of SCAN_CHUNK from the segment's aligned base
            const uint64_t stop = std::min<uint64_t>(seg.end, ((p & ~15ull) - ((p & ~15ull) - (seg.begin & ~15ull)) % SCAN_CHUNK) + SCAN_CHUNK);
            chunks.push_back(Chunk{p, (uint32_t)(stop - p), (uint32_t)k});
            p = stop;
        }
    }
    h->nchunks = (uint32_t)chunks.size();
    h->nsegs = (uint32_t)segs.size();
    int rc;
    if ((rc = h->d_segs.reserve(sizeof(SegInfo) * segs.size()))) return rc;
    if ((rc = h->d_chunks.reserve(sizeof(Chunk) * std::max<size_t>(1, chunks.size())))) return rc;
    if ((rc = h->d_chunk_base.reserve(4ull * (h->nchunks + 1)))) return rc;
    if ((rc = h->d_chunk_count.reserve(4ull * (h->nchunks + 1)))) return rc;
    if ((rc = h->d_chunk_scan.reserve(4ull * (h->nchunks + 1)))) return rc;
    if ((rc = h->d_pref.reserve(64ull * (h->nchunks + 1)))) return rc;
    if ((rc = h->d_counters.reserve(4 * CNT_WORDS))) return rc;
    if ((rc = h->d_totals.reserve(sizeof(Totals)))) return rc;
    if ((rc = h->mailbox.reserve(256))) return rc;
    if ((rc = h->d_anom.reserve(4ull * ANOM_CAP))) return rc;
    CK(cudaMemcpyAsync(h->d_segs.p, segs.data(), sizeof(SegInfo) * segs.size(), cudaMemcpyHostToDevice, h->stream));
    if (!chunks.empty()) CK(cudaMemcpyAsync(h->d_chunks.p, chunks.data(), sizeof(Chunk) * chunks.size(), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));   // segs / chunks are locals
    }
    h->args.in = h->d_ext ? h->d_ext : h->d_in.as<uint8_t>();
    h->args.in_len = in_len + 64;
    h->args.segs = h->d_segs.as<SegInfo>(); h->args.nsegs = h->nsegs;
    h->args.chunks = h->d_chunks.as<Chunk>(); h->args.nchunks = h->nchunks;
    h->args.chunk_base = h->d_chunk_base.as<uint32_t>(); h->args.chunk_count = h->d_chunk_count.as<uint32_t>();
    h->args.chunk_scan = h->d_chunk_scan.as<uint32_t>(); h->args.counters = h->d_counters.as<uint32_t>();
    h->args.pref = h->d_pref.as<uint16_t>();
    h->args.anom = h->d_anom.as<uint32_t>();
    h->args.totals = h->d_totals.as<Totals>();
    h->tables_ready = true;
    return 0;
}

static int reserve_cand(bnflac* h, uint32_t cap) {
    int rc;
    if ((rc = h->d_cand_tmp.reserve(sizeof(Cand) * (size_t)cap))) return rc;
    if ((rc = h->d_cand.reserve(sizeof(Cand) * (size_t)cap))) return rc;
    if ((rc = h->d_seg_crc.reserve(2ull * cap))) return rc;
    if ((rc = h->d_next.reserve(4ull * cap))) return rc;
    if ((rc = h->d_flen.reserve(4ull * cap))) return rc;
    if ((rc = h->d_status.reserve(cap))) return rc;
    if ((rc = h->d_sub.reserve(sizeof(SubInfo) * MAX_CH * (size_t)cap))) return rc;
    if ((rc = h->d_pcm_off.reserve(8ull * cap))) return rc;
    if ((rc = h->d_acc_idx.reserve(4ull * cap))) return rc;
    h->cand_cap = cap;
    h->args.cand_tmp = h->d_cand_tmp.as<Cand>(); h->args.cand = h->d_cand.as<Cand>(); h->args.cand_cap = cap;
    h->args.seg_crc = h->d_seg_crc.as<uint16_t>(); h->args.next = h->d_next.as<uint32_t>(); h->args.flen = h->d_flen.as<uint32_t>();
    h->args.status = h->d_status.as<uint8_t>(); h->args.sub = h->d_sub.as<SubInfo>(); h->args.pcm_off = h->d_pcm_off.as<uint64_t>();
    h->args.acc_idx = h->d_acc_idx.as<uint32_t>();
    return 0;
}

// K1b..K2 + prefix for at most `nb` candidates: every kernel takes the real count from the device and leaves early
static int launch_front_tail(bnflac* h, uint32_t nb) {
    CK(cudaEventRecord(h->ev[1], h->stream));
    launch_order(h->args, h->stream);
    launch_crc(h->args, nb, h->stream);
    CK(cudaEventRecord(h->ev[2], h->stream));
    launch_link(h->args, nb, h->stream);
    CK(cudaEventRecord(h->ev[3], h->stream));
    // few, large frames: subframe starts are guessed and walked in parallel (kernels.cu, "speculative parse")
    if (parse_wants_speculation(nb, h->info.channels)) {
        const uint64_t cap = (uint64_t)std::max<uint32_t>(nb, 1u) * h->info.channels * 2u + 64u;
        int rc;
        if ((rc = h->d_spec_jobs.reserve(sizeof(SpecJob) * cap)) || (rc = h->d_spec_base.reserve(4ull * h->cand_cap)) ||
            (rc = h->d_spec_count.reserve(h->cand_cap)) || (rc = h->d_spec_done.reserve(h->cand_cap))) return rc;
        h->args.spec_jobs = h->d_spec_jobs.as<SpecJob>(); h->args.spec_cap = (uint32_t)std::min<uint64_t>(cap, h->d_spec_jobs.cap / sizeof(SpecJob));
        h->args.spec_base = h->d_spec_base.as<uint32_t>(); h->args.spec_count = h->d_spec_count.as<uint8_t>(); h->args.spec_done = h->d_spec_done.as<uint8_t>();
    } else { h->args.spec_jobs = nullptr; h->args.spec_cap = 0; h->args.spec_base = nullptr; h->args.spec_count = nullptr; h->args.spec_done = nullptr; }
    launch_parse(h->args, nb, h->stream);
    launch_resync(h->args, h->stream);
    launch_prefix(h->args, nb, h->info.bytes_per_sample, h->stream);
    // frames of unlike blocksizes: regroup by blocksize class so that the frames a decode warp holds take equally long
    h->args.acc_sorted = nullptr; h->args.bucket_hist = nullptr;
    if (h->mixed_blocksizes && !getenv("BNFLAC_NO_BUCKETS")) {
        int rc;
        if ((rc = h->d_acc_sorted.reserve(4ull * h->cand_cap)) || (rc = h->d_bucket_hist.reserve(128ull * ((uint64_t)h->cand_cap / BUCKET_CHUNK + 2)))) return rc;
        h->args.acc_sorted = h->d_acc_sorted.as<uint32_t>(); h->args.bucket_hist = h->d_bucket_hist.as<uint32_t>();
        launch_bucket(h->args, nb, h->stream);
    }
    CK(cudaEventRecord(h->ev[4], h->stream));
    return 0;
}
static_assert(sizeof(Totals) % 4 == 0 && sizeof(Totals) / 4 <= 32 - 16, "totals fit the mailbox");

static void remember_pass(bnflac* h) {          // what the next pass over the same bytes can be launched on without asking
    h->pred.valid = true; h->pred.ncand = h->ncand; h->pred.n_accepted = h->totals.n_accepted;
    h->pred.max_order = h->totals.max_order; h->pred.any_wide = h->totals.any_wide; h->pred.pcm_bytes = h->totals.pcm_bytes;
}

// Runs K1..K2 + prefix (everything up to knowing the output size).  Two host hand-offs: the candidate count after the
// scan (table capacity, grid sizes) and the totals after the prefix (output size, decode variant).
static int run_front(bnflac* h) {
    h->front_ran = true;
    int rc;
    CK(cudaSetDevice(h->device));
    h->launches0 = kernel_launch_count();
    CK(cudaEventRecord(h->ev[0], h->stream));
    if ((rc = ensure_input(h))) return rc;
    if ((rc = ensure_tables(h))) return rc;
    if (!h->cand_cap) {
        const uint64_t nbytes = h->batch_segs.empty() ? (h->slice_end - h->slice_begin) : h->len;
        uint64_t est = nbytes / 512 + 4096;
        if ((rc = reserve_cand(h, (uint32_t)std::min<uint64_t>(est, 0x7fffffff)))) return rc;
    }
    for (int attempt = 0;; attempt++) {
        launch_clear(h->args, h->stream);
        launch_scan(h->args, h->stream);
        volatile uint32_t* counters = (volatile uint32_t*)h->mailbox.p;
        launch_publish(h->d_counters.p, h->mailbox.p, 2, h->stream);
        CK(cudaStreamSynchronize(h->stream));
        if (counters[0] > h->cand_cap) {
            if (attempt > 2) return BNFLAC_ERR_MEMORY;
            if ((rc = reserve_cand(h, counters[0] + counters[0] / 8 + 1024))) return rc;
            continue;
        }
        h->ncand = counters[0];
        break;
    }
    if ((rc = launch_front_tail(h, h->ncand))) return rc;
    launch_publish(h->d_totals.p, (uint8_t*)h->mailbox.p + 64, sizeof(Totals) / 4, h->stream);
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    memcpy(&h->totals, (const uint8_t*)h->mailbox.p + 64, sizeof(Totals));
    h->diag_valid = false;
    return 0;
}

// Passes of more than one wave of decode warps get the scratch of the balanced schedule (kernels_decode.cuh): a flag word and one
// job state per slot.  Without it (small passes, no memory) launch_decode uses the plain launch.
static int prepare_decode_sched(bnflac* h, uint32_t nacc, uint32_t expect) {
    h->args.dec_flags = nullptr; h->args.dec_state = nullptr; h->args.dec_slots = 0; h->args.dec_expect = expect;
    const uint32_t slots = decode_sched_slots(nacc, h->info.channels);
    if (!slots) return 0;
    const uint64_t flag_bytes = ((uint64_t)slots * 4 + 255) & ~255ull;
    const void* before = h->d_dec_sched.p;
    if (h->d_dec_sched.reserve(flag_bytes + (uint64_t)slots * decode_sched_state_bytes())) { cudaGetLastError(); return 0; }
    if (h->d_dec_sched.p != before) CK(cudaMemsetAsync(h->d_dec_sched.p, 0, flag_bytes, h->stream));   // whatever the block held before is not an epoch
    h->args.dec_flags = h->d_dec_sched.as<uint32_t>();
    h->args.dec_state = reinterpret_cast<uint32_t*>(h->d_dec_sched.as<uint8_t>() + flag_bytes);
    h->args.dec_slots = slots;
    return 0;
}

static int run_back(bnflac* h, uint8_t* d_out, uint64_t cap) {
    if (h->totals.pcm_bytes > cap) return BNFLAC_ERR_CAPACITY;
    h->args.out = d_out; h->args.out_cap = cap;
    int rc_s;
    if ((rc_s = prepare_decode_sched(h, h->totals.n_accepted, h->totals.n_accepted))) return rc_s;
    if (h->totals.n_accepted)
        launch_decode(h->args, h->totals.n_accepted, h->info.channels, h->info.bytes_per_sample, h->totals.max_order, h->totals.any_wide != 0, h->stream);
    CK(cudaEventRecord(h->ev[5], h->stream));
    CK(cudaGetLastError());
    remember_pass(h);
    return 0;
}

// A pass over bytes that have been decoded before on this handle (a device-resident stream decoded again: the benchmark
// loop, a frame-range shard re-run) is launched in one go on what the previous pass found -- candidate count, frames
// delivered, decode variant, output size -- WITHOUT the two host hand-offs of run_front; the kernels take the real counts
// from the device, and the host checks afterwards that the pass stayed inside what it was launched for.  Returns 1 when it
// did not (the caller then runs the synchronous path), 0 when the result stands.
static uint32_t order_class(uint32_t o) { return o <= 4 ? 4u : o <= 8 ? 8u : o <= 12 ? 12u : o <= 16 ? 16u : 32u; }
static int run_pass_predicted(bnflac* h, uint8_t* d_out, uint64_t cap) {
    const bnflac::Predict p = h->pred;
    if (!p.valid || !h->cand_cap || !h->tables_ready || !h->uploaded || p.pcm_bytes > cap || getenv("BNFLAC_NO_PREDICT")) return 1;
    h->front_ran = true;
    CK(cudaSetDevice(h->device));
    h->launches0 = kernel_launch_count();
    CK(cudaEventRecord(h->ev[0], h->stream));
    const uint32_t nb = (uint32_t)std::min<uint64_t>(h->cand_cap, (uint64_t)p.ncand + p.ncand / 16 + 64);
    const uint32_t nacc = (uint32_t)std::min<uint64_t>(nb, (uint64_t)p.n_accepted + p.n_accepted / 16 + 64);
    launch_clear(h->args, h->stream);
    launch_scan(h->args, h->stream);
    int rc;
    if ((rc = launch_front_tail(h, nb))) return rc;
    h->args.out = d_out; h->args.out_cap = cap;
    if ((rc = prepare_decode_sched(h, nacc, p.n_accepted))) return rc;
    if (p.n_accepted) launch_decode(h->args, nacc, h->info.channels, h->info.bytes_per_sample, p.max_order, p.any_wide != 0, h->stream);
    CK(cudaEventRecord(h->ev[5], h->stream));
    launch_publish(h->d_counters.p, h->mailbox.p, 2, h->stream);
    launch_publish(h->d_totals.p, (uint8_t*)h->mailbox.p + 64, sizeof(Totals) / 4, h->stream);
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    const uint32_t found = ((volatile uint32_t*)h->mailbox.p)[0];
    Totals t; memcpy(&t, (const uint8_t*)h->mailbox.p + 64, sizeof t);
    const bool stands = found <= nb && t.n_accepted <= nacc && t.pcm_bytes <= cap && (t.n_accepted == 0) == (p.n_accepted == 0) &&
                        order_class(t.max_order) == order_class(p.max_order) && (t.any_wide != 0) == (p.any_wide != 0);
    if (!stands) { h->pred.valid = false; return 1; }
    h->ncand = found; h->totals = t;
    h->diag_valid = false;
    remember_pass(h);
    return 0;
}

static int finish_timing(bnflac* h) {
    CK(cudaEventSynchronize(h->ev[5]));
    auto ms = [&](int a, int b) { float t = 0; cudaEventElapsedTime(&t, h->ev[a], h->ev[b]); return t; };
    h->timing.scan = ms(0, 1); h->timing.crc = ms(1, 2); h->timing.link = ms(2, 3); h->timing.parse = ms(3, 4);
    h->timing.decode = ms(4, 5); h->timing.total = ms(0, 5);
    h->timing.launches = (uint32_t)(kernel_launch_count() - h->launches0);
    static const bool trace = getenv("BNFLAC_TRACE") != nullptr;
    if (trace && h->args.spec_jobs) {
        uint32_t cnt[8] = {0};
        cudaMemcpy(cnt, h->d_counters.p, sizeof cnt, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[bnflac] speculative parse: %u candidates, %u subframe walks queued (cap %u), %u frames confirmed, %u left to the serial walk of %u off-chain\n",
                cnt[0], cnt[CNT_SPEC], h->args.spec_cap, cnt[CNT_SPEC_DONE], h->totals.n_accepted > cnt[CNT_SPEC_DONE] ? h->totals.n_accepted - cnt[CNT_SPEC_DONE] : 0u, cnt[CNT_ANOM]);
    }
    return 0;
}

static int decode_to_device(bnflac* h, void* d_dst, size_t cap, void** d_out, uint64_t* written) {
    int rc;
    uint8_t* out = (uint8_t*)d_dst;
    if (h->pred.valid) {                                      // second and later passes over the same bytes: no host hand-off
        if (!out) {
            if ((rc = h->d_out.reserve((size_t)h->pred.pcm_bytes + 64))) return rc;
            out = h->d_out.as<uint8_t>(); cap = h->d_out.cap;
        }
        rc = run_pass_predicted(h, out, cap);
        if (rc < 0) return rc;
        if (rc == 0) {
            if ((rc = finish_timing(h))) return rc;
            if (d_out) *d_out = out;
            if (written) *written = h->totals.pcm_bytes;
            h->state = BNFLAC_STATE_END_OF_STREAM;
            return 0;
        }
        out = (uint8_t*)d_dst;
    }
    if ((rc = run_front(h))) return rc;
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

// ---- host-destination decode: one pass, or a software pipeline of sub-shards for large host-resident streams
// Sub-shard schedule of the pipelined host decode: the first sub-shard is small so that the download engine starts
// early, sizes then double up to a cap (large sub-shards keep the kernels efficient and the per-pass host work small).
static size_t env_mb(const char* name, long dflt) { const char* e = getenv(name); long mb = e ? atol(e) : dflt; if (mb < 1) mb = 1; return (size_t)mb << 20; }
static size_t pipe_shard_bytes() { return env_mb("BNFLAC_PIPE_MB", 256); }                                   // cap (MiB of compressed bytes)
static size_t pipe_first_bytes() { return std::min(env_mb("BNFLAC_PIPE_FIRST_MB", 16), pipe_shard_bytes()); }
static std::vector<uint64_t> pipe_cuts(uint64_t b, uint64_t e) {       // boundaries b = c0 < c1 < ... < cK = e; K = 1: not worth pipelining
    std::vector<uint64_t> cuts{b};
    const uint64_t first = pipe_first_bytes(), cap = pipe_shard_bytes();
    if (e - b >= 3 * first) {
        const char* g = getenv("BNFLAC_PIPE_GROWTH");                       // per cent per sub-shard
        const uint64_t growth = g ? std::max<long>(100, std::min<long>(400, atol(g))) : 200;
        uint64_t pos = b, sz = first;
        while (e - pos > sz + sz / 2 && cuts.size() < 256) { pos += sz; cuts.push_back(pos); sz = std::min<uint64_t>((sz * growth / 100) & ~4095ull, cap); }
    }
    cuts.push_back(e);
    return cuts;
}

static int decode_host_single(bnflac* h, uint8_t* dst, size_t cap, uint64_t* written) {
    int rc = run_front(h); if (rc) return rc;
    if (h->totals.pcm_bytes > cap) return BNFLAC_ERR_CAPACITY;
    if ((rc = h->d_out.reserve((size_t)h->totals.pcm_bytes + 64))) return rc;
    if ((rc = run_back(h, h->d_out.as<uint8_t>(), h->d_out.cap))) return rc;
    if (h->totals.pcm_bytes) CK(cudaMemcpyAsync(dst, h->d_out.p, (size_t)h->totals.pcm_bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if ((rc = finish_timing(h))) return rc;
    *written = h->totals.pcm_bytes;
    return 0;
}

// Sub-shards k = 0..K-1 of the handle's range, each on its own stream: all uploads are queued first (the copy engine
// runs them back to back), then for each sub-shard in order: front kernels, host learns its PCM size (and thereby
// where the next one starts), decode kernel, download.  The host-side waits inside run_front block only the issuing
// thread: the other streams' copies and kernels keep running, so the H2D engine, the SMs and the D2H engine overlap.
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// one child handle per sub-shard [cuts[i], cuts[i+1]) of the parent's own range (kept when the cuts did not change)
static int make_kids(bnflac* h, const std::vector<uint64_t>& cuts) {
    const uint32_t K = (uint32_t)cuts.size() - 1;
    bool same = h->kids.size() == K;
    for (uint32_t i = 0; same && i < K; i++) same = h->kids[i]->sub_begin == cuts[i] && h->kids[i]->sub_end == cuts[i + 1];
    if (same) return 0;
    for (bnflac* k : h->kids) delete k;
    h->kids.clear();
    for (uint32_t i = 0; i < K; i++) {
        bnflac* c = new (std::nothrow) bnflac; if (!c) return BNFLAC_ERR_MEMORY;
        h->kids.push_back(c);
        c->opts = h->opts; c->opts.stream = nullptr; c->opts.device = h->device; c->opts.flags &= ~BNFLAC_OPT_VERIFY_MD5;
        c->info = h->info; c->len = h->len; c->host_ptr = h->host_ptr; c->sub_begin = cuts[i]; c->sub_end = cuts[i + 1]; c->mixed_blocksizes = h->mixed_blocksizes;
        compute_shard(c);
        c->state = BNFLAC_STATE_SEARCH_FOR_FRAME_SYNC;
        int rc = setup_device(c); if (rc) return rc;
    }
    return 0;
}

static int decode_host_pipelined(bnflac* h, const std::vector<uint64_t>& cuts, uint8_t* dst, size_t cap, uint64_t* written) {
    int rc = 0;
    const uint32_t K = (uint32_t)cuts.size() - 1;
    CK(cudaSetDevice(h->device));
    const bool trace = getenv("BNFLAC_TRACE") != nullptr;
    const double t_begin = now_ms();
    double t_kids = 0, t_tables = 0, t_upq = 0, t_loop = 0;
    if ((rc = make_kids(h, cuts))) return rc;
    auto drain = [&]() { if (h->up_stream) cudaStreamSynchronize(h->up_stream); for (bnflac* c : h->kids) if (c->stream) cudaStreamSynchronize(c->stream); };
    t_kids = now_ms();
    // tables first: their (small) uploads must not queue behind the bulk uploads on the copy engine
    for (bnflac* c : h->kids) {
        if (!c->tables_ready && ((rc = c->d_in.reserve((size_t)(c->slice_end - c->slice_begin) + 128)) || (rc = ensure_tables(c)))) { drain(); return rc; }
        if (!c->cand_cap && (rc = reserve_cand(c, (uint32_t)((c->slice_end - c->slice_begin) / 512 + 4096)))) { drain(); return rc; }
    }
    t_tables = now_ms();
    // All uploads go through ONE stream, in sub-shard order.  Spread over the sub-shards' own streams the copy engines
    // would serve them concurrently, every upload would finish late and no download could start early.
    if (!h->up_stream) CK(cudaStreamCreateWithFlags(&h->up_stream, cudaStreamNonBlocking));
    cudaEventRecord(h->kids.front()->ev[6], h->up_stream);
    for (bnflac* c : h->kids) {
        if (c != h->kids.front()) cudaEventRecord(c->ev[6], h->up_stream);
        if (!c->uploaded) {
            const size_t n = (size_t)(c->slice_end - c->slice_begin);
            if (cudaMemcpyAsync(c->d_in.p, c->host_ptr + c->slice_begin, n, cudaMemcpyHostToDevice, h->up_stream) != cudaSuccess ||
                cudaMemsetAsync((uint8_t*)c->d_in.p + n, 0, 128, h->up_stream) != cudaSuccess) { drain(); cudaStreamSynchronize(h->up_stream); g_cuda_err = "pipelined upload"; return BNFLAC_ERR_CUDA; }
            c->uploaded = true;
        }
        cudaEventRecord(c->ev[8], h->up_stream);
        cudaStreamWaitEvent(c->stream, c->ev[8], 0);
    }
    t_upq = now_ms();
    uint64_t off = 0;
    for (bnflac* c : h->kids) {
        if ((rc = run_front(c))) break;
        c->pcm_base = off;
        if (off + c->totals.pcm_bytes > cap) { rc = BNFLAC_ERR_CAPACITY; break; }
        if ((rc = c->d_out.reserve((size_t)c->totals.pcm_bytes + 64))) break;
        if ((rc = run_back(c, c->d_out.as<uint8_t>(), c->d_out.cap))) break;
        if (c->totals.pcm_bytes && cudaMemcpyAsync(dst + off, c->d_out.p, (size_t)c->totals.pcm_bytes, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { rc = BNFLAC_ERR_CUDA; g_cuda_err = "cudaMemcpyAsync (D2H)"; break; }
        cudaEventRecord(c->ev[7], c->stream);
        off += c->totals.pcm_bytes;
    }
    t_loop = now_ms();
    drain();
    if (trace) fprintf(stderr, "[bnflac] pipelined K=%u: kids %.2f ms, tables %.2f, upload queue %.2f, issue loop %.2f, drain %.2f\n", K,
                       t_kids - t_begin, t_tables - t_kids, t_upq - t_tables, t_loop - t_upq, now_ms() - t_loop);
    if (rc) return rc;
    CK(cudaGetLastError());
    bnflac_timing t{};
    for (bnflac* c : h->kids) {
        if ((rc = finish_timing(c))) return rc;
        t.scan += c->timing.scan; t.crc += c->timing.crc; t.link += c->timing.link; t.parse += c->timing.parse; t.decode += c->timing.decode;
        t.launches += c->timing.launches;
    }
    if (trace) {
        for (bnflac* c : h->kids) {
            float a = 0, b = 0, d = 0, e = 0;
            cudaEventElapsedTime(&a, h->kids.front()->ev[6], c->ev[6]); cudaEventElapsedTime(&b, h->kids.front()->ev[6], c->ev[0]);
            cudaEventElapsedTime(&d, h->kids.front()->ev[6], c->ev[5]); cudaEventElapsedTime(&e, h->kids.front()->ev[6], c->ev[7]);
            fprintf(stderr, "[bnflac]   sub-shard %3.0f MiB: upload %6.2f..%6.2f  kernels ..%6.2f  download ..%6.2f ms (%.0f MiB)\n",
                    (c->own_end - c->own_begin) / 1048576.0, a, b, d, e, c->totals.pcm_bytes / 1048576.0);
        }
    }
    float wall = 0;
    cudaEventElapsedTime(&wall, h->kids.front()->ev[6], h->kids.back()->ev[7]);
    t.total = wall;                     // first upload queued .. last download finished
    h->timing = t;
    h->diag_valid = false;
    *written = off;
    return 0;
}

static int pull_all(bnflac* h);
static int decode_host(bnflac* h, uint8_t* dst, size_t cap, uint64_t* written) {
    if (h->rd_active && h->pull_cb) { for (bnflac* k : h->kids) delete k; h->kids.clear(); }
    h->rd_active = false;                     // a one-shot decode ends any streaming Read session on this handle
    h->pl_session = false;
    { int rc = pull_all(h); if (rc) return rc; }
    if (h->host_ptr && !h->d_ext && h->batch_segs.empty()) {
        const std::vector<uint64_t> cuts = pipe_cuts(std::max<uint64_t>(h->own_begin, h->info.first_frame_offset), h->own_end);
        if (cuts.size() > 2) return decode_host_pipelined(h, cuts, dst, cap, written);
    }
    for (bnflac* k : h->kids) delete k;
    h->kids.clear();
    return decode_host_single(h, dst, cap, written);
}

// ---- streaming Read (the consumer the reference sketches in StreamingPlayer.cs:8-19,424-464: a ring of
// small buffers refilled from FLACDecoder.Read while the previous ones play).  FLACDecoder.Read decodes one more frame
// whenever its queue runs dry (FLACDecoder.cs:124-224); here the stream is cut into sub-shards by frame ranges -- a small
// first one so that the first Read returns after one short upload + pass + download, then doubling -- and the
// sub-shards after the one being read are decoded AHEAD of the reader on their own streams into their own pinned
// buffers.  A Read only ever waits for the download event of the sub-shard it is copying from; consumed sub-shards give
// their device and pinned blocks back to the pool (their small frame tables stay for the diagnostics).
static size_t env_kb(const char* name, long dflt) { const char* e = getenv(name); long kb = e ? atol(e) : dflt; if (kb < 16) kb = 16; return (size_t)kb << 10; }
static std::vector<uint64_t> read_cuts(const bnflac* h, uint64_t b, uint64_t e) {
    uint64_t first = env_kb("BNFLAC_READ_FIRST_KB", 4096), cap = env_mb("BNFLAC_READ_MB", 64);
    if (h->opts.read_chunk_frames) {       // caller-chosen look-ahead batch: that many frames of average compressed size
        const uint64_t bs = h->info.max_blocksize ? h->info.max_blocksize : 4096;
        const uint64_t nframes = h->info.total_samples ? std::max<uint64_t>(1, h->info.total_samples / bs) : 0;
        const uint64_t per = nframes ? std::max<uint64_t>(64, (e - b) / nframes) : std::max<uint64_t>(64, frame_bound(h->info) / 2);
        first = cap = std::max<uint64_t>(16384, per * h->opts.read_chunk_frames);
    }
    if (cap < first) cap = first;
    std::vector<uint64_t> cuts{b};
    if (e - b >= 3 * first) {
        uint64_t pos = b, sz = first;
        while (e - pos > sz + sz / 2 && cuts.size() < 4096) { pos += sz; cuts.push_back(pos); sz = std::min<uint64_t>(2 * sz, cap); }
    }
    cuts.push_back(e);
    return cuts;
}

static constexpr uint32_t READ_LOOKAHEAD = 2;      // sub-shards in flight beyond the one being read

// ---- lazily pulled source: FLACDecoder(Stream) reads only the metadata in its constructor and one frame's worth of bytes per
// Read (FLACDecoder.cs:72-88,207-224,325-363).  With BNFLAC_OPT_LAZY_PULL bnflac_open_callbacks does the same -- it pulls
// until STREAMINFO and the end of the metadata are in hand -- and bnflac_read pulls what the next sub-shard needs (its
// byte range plus one maximum frame of overlap) right before issuing it.  Everything pulled is kept (the diagnostics replay
// the reference's sync search over the bytes between frames).
static double now_ms();
static int pull_more(bnflac* h, uint64_t need_len) {         // until need_len bytes are in hand or the stream ends
    const size_t req = env_kb("BNFLAC_PULL_KB", 1024);        // bytes asked of the callback per call
    GrowBuf& b = h->pulled;
    struct Tick { bnflac* h; double t0; ~Tick() { h->tr_pull += now_ms() - t0; } } tick{h, now_ms()};
    while (!h->pull_eof && b.len < need_len) {
        if (!b.room(req)) return BNFLAC_ERR_MEMORY;
        size_t got = req;
        const int st = h->pull_cb(h->pull_user, b.p + b.len, &got);
        if (st == 2) return BNFLAC_ERR_ABORTED;
        if (got > req) return BNFLAC_ERR_ARG;
        b.len += got;
        if (st == 1 || got == 0) h->pull_eof = true;
    }
    h->host_ptr = b.p; h->len = b.len;                        // (the buffer may have moved: kids take the pointer when they are issued)
    return 0;
}
static int pull_all(bnflac* h) {                             // entry points that need the whole stream
    if (!h->pull_cb) return 0;
    if (!h->pull_eof) { int rc = pull_more(h, ~0ull); if (rc) return rc; }
    compute_shard(h);                                        // the handle's own range was cut when only the metadata was in hand
    return 0;
}

// next sub-shard of a lazily pulled stream: pull its bytes, create its handle, issue it
static int stream_issue(bnflac* h, uint32_t k);
static int lazy_issue_next(bnflac* h) {
    uint64_t first = env_kb("BNFLAC_READ_FIRST_KB", 4096), cap = env_mb("BNFLAC_READ_MB", 64);
    if (h->opts.read_chunk_frames) {       // caller-chosen look-ahead batch (as read_cuts; the stream's length is not known here)
        const uint64_t mn = h->info.min_framesize, mx = h->info.max_framesize;
        const uint64_t per = (mn && mx) ? (mn + mx) / 2 : std::max<uint64_t>(64, frame_bound(h->info) / 2);
        first = cap = std::max<uint64_t>(16384, per * h->opts.read_chunk_frames);
    }
    if (cap < first) cap = first;
    if (!h->pl_size) { h->pl_size = first; h->pl_next = h->info.first_frame_offset; }
    const uint64_t b = h->pl_next;
    uint64_t e = b + h->pl_size;
    int rc = pull_more(h, e + h->pl_size / 2 + frame_bound(h->info) + 64); if (rc) return rc;
    if (h->pull_eof && h->len <= e + h->pl_size / 2) e = h->len;          // what is left is not worth a sub-shard of its own
    if (e > h->len) e = h->len;
    bnflac* c = new (std::nothrow) bnflac; if (!c) return BNFLAC_ERR_MEMORY;
    h->kids.push_back(c);
    c->opts = h->opts; c->opts.stream = nullptr; c->opts.device = h->device; c->opts.flags &= ~(BNFLAC_OPT_VERIFY_MD5 | BNFLAC_OPT_LAZY_PULL);
    c->info = h->info; c->len = h->len; c->host_ptr = h->host_ptr; c->sub_begin = b; c->sub_end = std::max(e, b + 1); c->mixed_blocksizes = h->mixed_blocksizes;
    if (b >= h->len) { c->sub_begin = h->len ? h->len - 1 : 0; c->sub_end = h->len; }     // empty tail
    compute_shard(c);
    c->state = BNFLAC_STATE_SEARCH_FOR_FRAME_SYNC;
    if ((rc = setup_device(c))) return rc;
    h->pl_next = e; h->pl_size = std::min<uint64_t>(2 * h->pl_size, cap);
    if ((rc = stream_issue(h, (uint32_t)h->kids.size() - 1))) return rc;
    h->rd_issued = (uint32_t)h->kids.size();
    return 0;
}
static bool lazy_more(const bnflac* h) { return h->pl_session && !(h->pull_eof && h->pl_size && h->pl_next >= h->len); }

// upload + front kernels + decode + download of sub-shard k, all asynchronous but for the two size hand-offs in run_front
static int stream_issue(bnflac* h, uint32_t k) {
    bnflac* c = h->kids[k];
    int rc = run_front(c); if (rc) return rc;
    c->pcm_base = h->rd_total;
    h->rd_total += c->totals.pcm_bytes;
    if ((rc = c->pcm_host.reserve((size_t)c->totals.pcm_bytes + 64))) return rc;
    if ((rc = c->d_out.reserve((size_t)c->totals.pcm_bytes + 64))) return rc;
    if ((rc = run_back(c, c->d_out.as<uint8_t>(), c->d_out.cap))) return rc;
    if (c->totals.pcm_bytes) CK(cudaMemcpyAsync(c->pcm_host.p, c->d_out.p, (size_t)c->totals.pcm_bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaEventRecord(c->ev[7], c->stream));
    c->pcm_len = c->totals.pcm_bytes;
    return 0;
}

static int stream_issue_all(bnflac* h) {
    while (h->rd_issued < h->kids.size()) { int rc = stream_issue(h, h->rd_issued); if (rc) return rc; h->rd_issued++; }
    while (lazy_more(h)) { int rc = lazy_issue_next(h); if (rc) return rc; }
    return 0;
}

static int64_t stream_read(bnflac* h, uint8_t* dst, size_t count) {
    CK(cudaSetDevice(h->device));
    size_t done = 0;
    int rc;
    while (done < count && (h->rd_cur < h->kids.size() || lazy_more(h))) {
        if (h->rd_issued <= h->rd_cur) {
            const double t0 = now_ms();
            if (h->rd_issued < h->kids.size()) { if ((rc = stream_issue(h, h->rd_issued))) return rc; h->rd_issued++; }
            else if ((rc = lazy_issue_next(h))) return rc;
            h->tr_issue += now_ms() - t0;
        }
        bnflac* c = h->kids[h->rd_cur];
        if (cudaEventQuery(c->ev[7]) != cudaSuccess) { const double t0 = now_ms(); CK(cudaEventSynchronize(c->ev[7])); h->tr_wait += now_ms() - t0; }
        const size_t n = (size_t)std::min<uint64_t>(c->pcm_len - h->rd_off, count - done);
        if (n) { const bool big = n >= (1u << 20); const double t0 = big ? now_ms() : 0; memcpy(dst + done, (const uint8_t*)c->pcm_host.p + h->rd_off, n); if (big) h->tr_copy += now_ms() - t0; }
        done += n; h->rd_off += n;
        if (h->rd_off == c->pcm_len) {               // sub-shard consumed: its blocks go back to the pool
            finish_timing(c);
            h->timing.scan += c->timing.scan; h->timing.crc += c->timing.crc; h->timing.link += c->timing.link; h->timing.parse += c->timing.parse;
            h->timing.decode += c->timing.decode; h->timing.total += c->timing.total; h->timing.launches += c->timing.launches;
            c->pcm_host.release(); c->d_out.release();
            if (!c->d_ext) { c->d_in.release(); c->uploaded = false; c->tables_ready = false; }
            h->rd_cur++; h->rd_off = 0;
        }
    }
    // keep the decode ahead of the reader: at most one more sub-shard per call, so no single Read pays for several -- and
    // none in the first call of a lazily pulled session, which has just paid for pulling its first sub-shard
    const bool first_lazy_call = h->pl_session && h->kids.size() == 1 && h->rd_cur == 0 && h->rd_issued == 1 && !h->pl_ahead;
    h->pl_ahead = true;
    h->tr_reads++;
    if (!first_lazy_call && h->rd_issued <= h->rd_cur + READ_LOOKAHEAD) {
        const double t0 = now_ms();
        if (h->rd_issued < h->kids.size()) { if ((rc = stream_issue(h, h->rd_issued))) return rc; h->rd_issued++; }
        else if (lazy_more(h) && (rc = lazy_issue_next(h))) return rc;
        h->tr_issue += now_ms() - t0;
    }
    h->state = (h->rd_cur == h->kids.size() && !lazy_more(h)) ? BNFLAC_STATE_END_OF_STREAM : BNFLAC_STATE_READ_FRAME;
    return (int64_t)done;
}

// ---- diagnostics: frame / subframe tables and the error-callback events the reference would have raised
namespace {
// Frame header check of the reference's sync search: 0 valid, 1 BAD_HEADER (CRC-8 / syntax), 3 UNPARSEABLE
// (reserved values), -1 truncated.  Host side, only ever run on the few bytes the decoder skipped in a damaged stream.
int header_rc(const uint8_t* p, size_t avail) {
    if (avail < 5) return -1;
    bool unparse = (p[1] & 0x02) != 0;
    const bool variable = p[1] & 1;
    const int bsc = p[2] >> 4, src = p[2] & 15, ca = p[3] >> 4, ssc = (p[3] >> 1) & 7;
    if ((p[3] & 1) || bsc == 0 || ca > 10 || ssc == 3 || ssc == 7) unparse = true;
    if (src == 15) return 1;
    size_t q = 4;
    uint32_t x = p[q++];
    if (x >= 0x80) {
        int n = 0;
        while (x & (0x80u >> n)) n++;
        if (n == 1 || n > 7 || (!variable && n == 7)) return 1;
        for (int i = 1; i < n; i++) { if (q >= avail) return -1; if ((p[q++] >> 6) != 2) return 1; }
    }
    q += (bsc == 6) ? 1 : (bsc == 7) ? 2 : 0;
    q += (src == 12) ? 1 : (src == 13 || src == 14) ? 2 : 0;
    if (q + 1 > avail) return -1;
    uint32_t c = 0;
    for (size_t i = 0; i < q; i++) { c ^= p[i]; for (int k = 0; k < 8; k++) c = (c & 0x80) ? ((c << 1) ^ 0x07) & 0xFF : (c << 1) & 0xFF; }
    if (c != p[q]) return 1;
    return unparse ? 3 : 0;
}

} // namespace

// bytes [from, to) of the stream as the reference sees them while hunting for a sync code
static void scan_gap(bnflac* h, const uint8_t* host_ptr, uint64_t stream_len, DiagCursor& cur, uint64_t to, std::vector<uint32_t>& errors) {
    if (cur.ended || cur.expect >= to) return;
    const uint64_t from = cur.expect;
    const uint64_t hi = std::min<uint64_t>(stream_len, to + 32);          // a header that starts in the gap may end after it
    std::vector<uint8_t> tmp;
    const uint8_t* d;
    if (host_ptr) d = host_ptr + from;
    else { tmp.resize((size_t)(hi - from)); cudaMemcpy(tmp.data(), h->d_ext + from, tmp.size(), cudaMemcpyDeviceToHost); d = tmp.data(); }
    uint64_t pos = from;
    while (pos < to && pos + 2 <= stream_len) {
        const uint8_t* p = d + (pos - from);
        if (!(p[0] == 0xFF && (p[1] & 0xFC) == 0xF8)) { if (cur.in_sync) { cur.event(errors, 0); cur.in_sync = false; } pos++; continue; }
        const int rc = header_rc(p, (size_t)std::min<uint64_t>(hi - pos, stream_len - pos));
        if (rc < 0) { cur.ended = true; break; }
        if (rc == 1) { cur.event(errors, 1); cur.in_sync = true; pos += 2; continue; }
        if (rc == 3) { cur.event(errors, 3); cur.in_sync = true; pos += 2; continue; }
        // a valid header the engine did not consider (frame of another format, cut off by the end of the stream)
        if (cur.in_sync) { cur.event(errors, 0); cur.in_sync = false; }
        pos++;
    }
    cur.expect = std::max(cur.expect, to);
}

// Builds the host-side frame / subframe / error tables of one pass, replaying the reference's cursor over the frame table.
static int collect_diag(bnflac* h, DiagCursor& cur, uint64_t pcm_base, const uint8_t* host_ptr, uint64_t stream_len,
                        std::vector<bnflac_frame_t>& frames, std::vector<bnflac_subframe_t>& subs, std::vector<uint32_t>& errors) {
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
    const uint64_t base = h->d_ext ? 0 : h->slice_begin;
    for (uint32_t i = 0; i < n && !cur.ended; i++) {
        if (cand[i].flags & 2) continue;
        const bool is_frame = st[i] == ST_OK || st[i] == ST_CRC;
        if (!is_frame && st[i] != ST_LOSTSYNC && st[i] != ST_UNPARSEABLE && st[i] != ST_EOS) continue;      // covered by a frame / never reached
        const uint64_t off = cand[i].off + base;
        if (!cur.have_expect) { cur.expect = off; cur.have_expect = true; }
        scan_gap(h, host_ptr, stream_len, cur, off, errors);
        if (cur.ended) break;
        if (st[i] == ST_EOS) { cur.ended = true; break; }         // the stream ended inside this frame
        if (!is_frame) {                                          // the parse failed: reported, the search resumes where the bit reader stood
            cur.event(errors, st[i] == ST_LOSTSYNC ? 0u : 3u);   // (libFLAC 1.2.1; it reports LOST_SYNC again when it then skips bytes)
            cur.in_sync = true;
            cur.expect = off + fl[i];
            continue;
        }
        bnflac_frame_t f{};
        f.offset = off; f.length = fl[i]; f.blocksize = cand[i].bs;
        f.channels = (uint8_t)(cand[i].assign < 8 ? cand[i].assign + 1 : 2); f.bits_per_sample = cand[i].bps; f.assignment = cand[i].assign;
        f.status = st[i] == ST_OK ? BNFLAC_FRAME_OK : BNFLAC_FRAME_CRC_MISMATCH;
        f.number = cand[i].number; f.pcm_offset = po[i] + pcm_base;
        if (st[i] == ST_CRC) cur.event(errors, 2);              // FRAME_CRC_MISMATCH
        cur.expect = f.offset + f.length; cur.in_sync = true;
        frames.push_back(f);
        for (int c = 0; c < MAX_CH; c++) {
            bnflac_subframe_t s{};
            if (c < f.channels && st[i] == ST_OK) { const SubInfo& si = sub[(size_t)i * MAX_CH + c]; s.bit_offset = si.bit_offset; s.type = si.type; s.order = si.order; s.wasted = si.wasted; s.flags = si.flags; }
            subs.push_back(s);
        }
    }
    return 0;
}

static int fetch_diag(bnflac* h, bool all = true) {
    if (h->diag_valid) return 0;
    int rc;
    if (h->rd_active) {
        // streamed Read session: the tables grow sub-shard by sub-shard (each is folded in once, after its front kernels);
        // `all` decodes what is left first, otherwise only what has been issued so far is described
        if (all && (rc = stream_issue_all(h))) return rc;
        if (!h->diag_started) {
            h->frames.clear(); h->subs.clear(); h->errors.clear(); h->errors_at.clear();
            h->diag_cur = DiagCursor{};
            h->diag_cur.at = &h->errors_at; h->diag_cur.frames = &h->frames;
            h->diag_cur.expect = std::max<uint64_t>(h->own_begin, h->info.first_frame_offset);
            h->diag_cur.have_expect = h->own_begin <= h->info.first_frame_offset;
            h->diag_kids = 0; h->diag_started = true;
        }
        for (; h->diag_kids < h->rd_issued; h->diag_kids++) {
            bnflac* c = h->kids[h->diag_kids];
            if ((rc = collect_diag(c, h->diag_cur, c->pcm_base, h->host_ptr, h->len, h->frames, h->subs, h->errors))) return rc;
        }
        if (h->rd_issued == h->kids.size() && !lazy_more(h)) {
            if (h->pull_cb && h->pull_eof) h->own_end = std::max<uint64_t>(h->own_end, h->len);
            if (h->own_end >= h->len && h->diag_cur.have_expect) scan_gap(h, h->host_ptr, h->len, h->diag_cur, h->len, h->errors);   // what follows the last frame
            h->diag_valid = true;
        }
        return 0;
    }
    h->diag_started = false;
    h->frames.clear(); h->subs.clear(); h->errors.clear(); h->errors_at.clear();
    if (h->kids.empty() && !h->front_ran) return 0;                    // nothing decoded yet: empty tables
    DiagCursor cur;
    cur.at = &h->errors_at; cur.frames = &h->frames;
    if ((rc = pull_all(h))) return rc;
    cur.expect = std::max<uint64_t>(h->own_begin, h->info.first_frame_offset);
    cur.have_expect = h->own_begin <= h->info.first_frame_offset;      // a later shard starts wherever its first frame starts
    if (!h->kids.empty()) {
        for (bnflac* c : h->kids)
            if ((rc = collect_diag(c, cur, c->pcm_base, h->host_ptr, h->len, h->frames, h->subs, h->errors))) return rc;
    } else if ((rc = collect_diag(h, cur, 0, h->host_ptr, h->len, h->frames, h->subs, h->errors))) return rc;
    if (h->own_end >= h->len && cur.have_expect) scan_gap(h, h->host_ptr, h->len, cur, h->len, h->errors);   // what follows the last frame
    h->diag_valid = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------ batch of clips
namespace {
struct ClipMeta { bnflac_info_t info; int rc; size_t group; uint64_t seg_begin; };

// copies the clips of one group into a pinned staging buffer with several host threads (memcpy of gigabytes on one
// thread would dominate the batch)
// One clip into the staging buffer.  The destination (16-byte aligned: clip starts are) is written once and next read by the DMA
// engine, never by this CPU: non-temporal stores skip the read-for-ownership of every destination line, a third of the gather's
// host-memory traffic (clips of ~100 KB are below the size at which memcpy switches to them by itself).
static inline void copy_clip(uint8_t* d, const uint8_t* s, size_t n) {
#if defined(__x86_64__) && !defined(BNFLAC_NO_NT_COPY)
    if (n >= 4096 && !((uintptr_t)d & 15u)) {
        size_t k = 0;
        for (; k + 64 <= n; k += 64) {
            const __m128i a = _mm_loadu_si128((const __m128i*)(s + k)), b = _mm_loadu_si128((const __m128i*)(s + k + 16)),
                          c = _mm_loadu_si128((const __m128i*)(s + k + 32)), e = _mm_loadu_si128((const __m128i*)(s + k + 48));
            _mm_stream_si128((__m128i*)(d + k), a); _mm_stream_si128((__m128i*)(d + k + 16), b);
            _mm_stream_si128((__m128i*)(d + k + 32), c); _mm_stream_si128((__m128i*)(d + k + 48), e);
        }
        if (k < n) memcpy(d + k, s + k, n - k);
        return;
    }
#endif
    memcpy(d, s, n);
}
void parallel_gather(uint8_t* dst, const std::vector<std::pair<const uint8_t*, size_t>>& src, const std::vector<uint64_t>& at, size_t first = 0, size_t last = (size_t)-1) {
    const size_t n = std::min(last, src.size());
    size_t total = 0; for (size_t i = first; i < n; i++) total += src[i].second;
    unsigned nt = (unsigned)std::min<size_t>(std::max<size_t>(1, total >> 24), std::min<unsigned>(16, std::max(1u, std::thread::hardware_concurrency())));
    auto work = [&](unsigned t, unsigned step) {
        for (size_t i = first + t; i < n; i += step) copy_clip(dst + at[i], src[i].first, src[i].second);
#if defined(__x86_64__) && !defined(BNFLAC_NO_NT_COPY)
        _mm_sfence();                                  // the streamed lines are globally visible before the upload is issued
#endif
    };
    if (nt <= 1) { work(0, 1); return; }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++) th.emplace_back(work, t, nt);
    for (auto& x : th) x.join();
}
} // namespace

static int decode_batch_impl(const bnflac_span* clips, size_t n, const bnflac_opts* opts_in, uint8_t* dst, size_t cap, int dst_is_device,
                             bnflac_clip_result* results, uint64_t* written) {
    bnflac_opts opts = default_opts(opts_in);
    const bool trace = getenv("BNFLAC_TRACE") != nullptr;
    const double t_begin = now_ms();
    std::vector<ClipMeta> meta(n);
    struct Group { uint32_t ch, bps; std::vector<size_t> clips; };
    std::vector<Group> groups;
    for (size_t i = 0; i < n; i++) {
        ClipMeta& m = meta[i];
        m.rc = (clips[i].data && clips[i].len) ? parse_metadata(clips[i].data, clips[i].len, &m.info) : BNFLAC_ERR_ARG;
        if (!m.rc && (m.info.channels > 8 || m.info.bits_per_sample > 24 || m.info.bits_per_sample < 4)) m.rc = BNFLAC_ERR_UNSUPPORTED;
        if (results) { memset(&results[i], 0, sizeof results[i]); results[i].status = m.rc ? 4u : 0u; }
        if (m.rc) continue;
        size_t g = 0;
        for (; g < groups.size(); g++) if (groups[g].ch == m.info.channels && groups[g].bps == m.info.bits_per_sample) break;
        if (g == groups.size()) groups.push_back(Group{m.info.channels, m.info.bits_per_sample, {}});
        groups[g].clips.push_back(i);
        m.group = g;
    }
    uint64_t out_off = 0;
    int rc = 0;
    // Packed input (BNFLAC_OPT_PACKED_INPUT): the clips lie in ascending order inside ONE host buffer the caller owns from the
    // first clip's first byte to the last clip's last byte (a shard / archive file read in one piece).  The whole range is then
    // uploaded ONCE, in place, and every format group's pass addresses its clips inside it: no gather into staging memory (40
    // of the 104 ms of a 2.5 GB batch), no second copy of the input in host memory.  Without the flag the bytes between clips
    // are never touched: every clip is gathered.
    const uint8_t* pk_lo = nullptr; uint64_t pk_span = 0;
    if (opts.flags & BNFLAC_OPT_PACKED_INPUT) {        // the caller vouches for the whole range: never inferred from the addresses
        uint64_t total = 0, nvalid = 0; const uint8_t* prev_end = nullptr; bool ascending = true;
        for (size_t i = 0; i < n && ascending; i++) {
            if (meta[i].rc) continue;
            if (prev_end && clips[i].data < prev_end) ascending = false;
            if (!pk_lo) pk_lo = clips[i].data;
            prev_end = clips[i].data + clips[i].len; total += clips[i].len; nvalid++;
        }
        if (ascending && nvalid >= 2 && (uint64_t)(prev_end - pk_lo) <= total + total / 4 + nvalid * 1024) pk_span = (uint64_t)(prev_end - pk_lo);
    }
    DevBuf packed;
    if (pk_span) {
        int dev = opts.device; if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) dev = 0;
        CK(cudaSetDevice(dev));
        const double t_u0 = now_ms();
        if ((rc = packed.reserve((size_t)pk_span + 128))) return rc;
        if (cudaMemcpy(packed.p, pk_lo, (size_t)pk_span, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemset((uint8_t*)packed.p + pk_span, 0, 128) != cudaSuccess) {
            cudaGetLastError(); packed.release(); g_cuda_err = "batch upload (packed)"; return BNFLAC_ERR_CUDA;
        }
        if (trace) fprintf(stderr, "[bnflac] batch: packed input, %.1f MB uploaded in place in %.2f ms\n", pk_span / 1e6, now_ms() - t_u0);
    }
    struct PackedGuard { DevBuf& b; ~PackedGuard() { b.release(); } } packed_guard{packed};
    for (Group& G : groups) {
        bnflac h;
        h.opts = opts; h.info = meta[G.clips[0]].info;
        for (size_t ci : G.clips) if (meta[ci].info.min_blocksize != meta[ci].info.max_blocksize || meta[ci].info.max_blocksize != h.info.max_blocksize) { h.mixed_blocksizes = true; break; }
        // layout of the group's bytes on the device: each clip's frame data (metadata stripped) at a 16-byte aligned offset
        std::vector<std::pair<const uint8_t*, size_t>> src; std::vector<uint64_t> at;
        uint64_t pos = 0;
        for (size_t ci : G.clips) {
            const ClipMeta& m = meta[ci];
            const uint64_t first = m.info.first_frame_offset, nbytes = clips[ci].len - first;
            SegDesc d{};
            if (pk_span) pos = (uint64_t)(clips[ci].data + first - pk_lo);          // where the clip's frames already are
            d.begin = pos; d.end = pos + nbytes; d.own_begin = d.begin; d.own_end = d.end;
            d.sample_rate = m.info.sample_rate; d.min_bs = m.info.min_blocksize; d.max_bs = m.info.max_blocksize; d.max_frame_bytes = frame_bound(m.info);
            h.batch_segs.push_back(d);
            src.emplace_back(clips[ci].data + first, (size_t)nbytes); at.push_back(pos);
            pos = (pos + nbytes + 15) & ~15ull;
        }
        h.len = pk_span ? pk_span : pos;
        const double t_g0 = now_ms();
        if ((rc = setup_device(&h))) return rc;
        double t_g1 = t_g0, t_g2 = t_g0;
        if (pk_span) { h.d_ext = packed.as<uint8_t>(); rc = run_front(&h); if (rc) return rc; }
        else {
        // The clips are gathered into pinned staging memory by host threads (non-temporal stores) and uploaded from there.
        PinBuf stage;
        if ((rc = stage.reserve((size_t)pos + 64))) return rc;
        t_g1 = now_ms();
        if ((rc = h.d_in.reserve((size_t)pos + 128))) { stage.release(); return rc; }
        // The clips are gathered in runs of ~128 MB and each run is uploaded while the next is gathered (BNFLAC_BATCH_RUNS forces the
        // number of runs).  With the plain memcpy gather this overlap had been measured slower (104 -> 148 ms for 2.5 GB: the gather fell
        // from 62 to 19 GB/s while the DMA engine read host memory); with non-temporal stores the gather moves a third less data and
        // the overlap pays: 12,500 clips / 1.07 GB 34.3 ms in one run, 32.4 in four, 32.0 in eight.
        static const unsigned runs_env = getenv("BNFLAC_BATCH_RUNS") ? (unsigned)atoi(getenv("BNFLAC_BATCH_RUNS")) : 0u;
        const size_t nclip = src.size();
        const unsigned runs_auto = (unsigned)std::min<uint64_t>(8, std::max<uint64_t>(1, pos >> 27));
        const unsigned runs = (unsigned)std::max<size_t>(1, std::min<size_t>(runs_env ? runs_env : runs_auto, nclip / 64 + 1));
        bool up_ok = true;
        for (unsigned r = 0; r < runs && up_ok; r++) {
            const size_t c0 = nclip * r / runs, c1 = nclip * (r + 1) / runs;
            if (c0 == c1) continue;
            parallel_gather((uint8_t*)stage.p, src, at, c0, c1);
            const uint64_t b0 = at[c0], b1 = (c1 < nclip) ? at[c1] : pos;
            up_ok = cudaMemcpyAsync((uint8_t*)h.d_in.p + b0, (uint8_t*)stage.p + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, h.stream) == cudaSuccess;
        }
        t_g2 = now_ms();
        if (!up_ok || cudaMemsetAsync((uint8_t*)h.d_in.p + pos, 0, 128, h.stream) != cudaSuccess) { stage.release(); g_cuda_err = "batch upload"; return BNFLAC_ERR_CUDA; }
        h.uploaded = true;
        rc = run_front(&h);
        stage.release();                            // run_front synchronised the stream: the upload is done
        if (rc) return rc;
        }
        const double t_g3 = now_ms();
        const uint32_t ns = (uint32_t)h.batch_segs.size();
        if ((rc = h.d_seg_pcm.reserve(8ull * ns)) || (rc = h.d_seg_flags.reserve(4ull * ns))) return rc;
        CK(cudaMemsetAsync(h.d_seg_pcm.p, 0xFF, 8ull * ns, h.stream));
        CK(cudaMemsetAsync(h.d_seg_flags.p, 0, 4ull * ns, h.stream));
        if (h.ncand) launch_seg_summary(h.args, h.ncand, h.d_seg_pcm.as<uint64_t>(), h.d_seg_flags.as<uint32_t>(), h.stream);
        std::vector<uint64_t> seg_pcm(ns); std::vector<uint32_t> seg_flags(ns);
        CK(cudaMemcpyAsync(seg_pcm.data(), h.d_seg_pcm.p, 8ull * ns, cudaMemcpyDeviceToHost, h.stream));
        CK(cudaMemcpyAsync(seg_flags.data(), h.d_seg_flags.p, 4ull * ns, cudaMemcpyDeviceToHost, h.stream));
        CK(cudaStreamSynchronize(h.stream));
        // a segment without any candidate never wrote its slot: it starts where the next one does
        uint64_t nextb = h.totals.pcm_bytes;
        for (uint32_t k = ns; k-- > 0;) { if (seg_pcm[k] == ~0ull) seg_pcm[k] = nextb; nextb = seg_pcm[k]; }
        if (results) {
            for (uint32_t k = 0; k < ns; k++) {
                const size_t ci = G.clips[k];
                const ClipMeta& m = meta[ci];
                bnflac_clip_result& r = results[ci];
                r.pcm_offset = out_off + seg_pcm[k];
                r.pcm_bytes = (k + 1 < ns ? seg_pcm[k + 1] : h.totals.pcm_bytes) - seg_pcm[k];
                r.sample_rate = m.info.sample_rate; r.channels = m.info.channels; r.bits_per_sample = m.info.bits_per_sample;
                r.total_samples = m.info.total_samples;
                r.status = seg_flags[k];
                if (m.info.total_samples && r.pcm_bytes != m.info.pcm_bytes) r.status |= 2u;     // frames lost or extra
            }
        }
        if (dst) {
            if (out_off + h.totals.pcm_bytes > cap) return BNFLAC_ERR_CAPACITY;
            if (dst_is_device) {
                if ((rc = run_back(&h, dst + out_off, cap - out_off))) return rc;
            } else {
                if ((rc = h.d_out.reserve((size_t)h.totals.pcm_bytes + 64))) return rc;
                if ((rc = run_back(&h, h.d_out.as<uint8_t>(), h.d_out.cap))) return rc;
                if (h.totals.pcm_bytes) CK(cudaMemcpyAsync(dst + out_off, h.d_out.p, (size_t)h.totals.pcm_bytes, cudaMemcpyDeviceToHost, h.stream));
            }
            CK(cudaStreamSynchronize(h.stream));
            CK(cudaGetLastError());
        }
        out_off += h.totals.pcm_bytes;
        if (trace) fprintf(stderr, "[bnflac] batch group %u ch %u bit: %zu clips, %.1f MB: layout+setup %.2f ms, gather %.2f, upload+front %.2f, summary+decode %.2f (since call %.2f)\n",
                           G.ch, G.bps, G.clips.size(), pos / 1e6, t_g1 - t_g0, t_g2 - t_g1, t_g3 - t_g2, now_ms() - t_g3, now_ms() - t_begin);
    }
    if (written) *written = out_off;
    return 0;
}

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" {

int bnflac_abi_version(void) { return BNFLAC_ABI_VERSION; }
int bnflac_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }
const char* bnflac_last_cuda_error(void) { return g_cuda_err.c_str(); }
uint64_t bnflac_kernel_launches(void) { return (uint64_t)kernel_launch_count(); }
void bnflac_trim_pools(void) { g_pool.trim(); }

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

// ---- Ogg FLAC.  The reference's DLL exports FLAC__stream_decoder_init_ogg_* but its C# never binds them
// (LibFLACSharp.cs:42-85 lists the native-FLAC entry points only), so this is container breadth, not a parity surface: the
// pages are taken apart on the host and what is left -- "fLaC", the metadata blocks, the frames -- is a native FLAC stream
// that goes down the same pipeline.  Mapping (Ogg FLAC 1.0): page = "OggS", version 0, flags (1 continued packet, 2 first,
// 4 last page), granule position, serial, page number, CRC-32 (polynomial 0x04C11DB7, no reflection, initial value 0, over
// the page with the CRC field zeroed), segment count, lacing values; a lacing value < 255 ends a packet.  First packet:
// 0x7F "FLAC" major minor, number of header packets (BE16), then "fLaC" + the STREAMINFO block; every further header packet
// is one metadata block, every audio packet one frame.  Pages that fail their CRC or are out of sequence are dropped with the
// packet they interrupt (the frame chain then sees a jump in the frame numbers, like any other missing frame).
namespace {
struct OggCrcTable {
    uint32_t t[4][256];
    OggCrcTable() {
        for (uint32_t i = 0; i < 256; i++) { uint32_t r = i << 24; for (int k = 0; k < 8; k++) r = (r << 1) ^ ((r & 0x80000000u) ? 0x04C11DB7u : 0u); t[0][i] = r; }
        for (int k = 1; k < 4; k++) for (uint32_t i = 0; i < 256; i++) t[k][i] = (t[k - 1][i] << 8) ^ t[0][t[k - 1][i] >> 24];
    }
};
uint32_t ogg_crc(uint32_t crc, const uint8_t* p, size_t n) {
    static const OggCrcTable T;
    for (; n >= 4; p += 4, n -= 4) {
        crc ^= (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3];
        crc = T.t[3][crc >> 24] ^ T.t[2][(crc >> 16) & 255] ^ T.t[1][(crc >> 8) & 255] ^ T.t[0][crc & 255];
    }
    for (; n; p++, n--) crc = (crc << 8) ^ T.t[0][(crc >> 24) ^ *p];
    return crc;
}
inline uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
inline bool is_ogg(const uint8_t* d, size_t n) { return n >= 4 && !memcmp(d, "OggS", 4); }

// 0 and the native stream in `out`; NOT_FLAC when no FLAC logical stream starts in the data, UNSUPPORTED for another mapping version
int ogg_depage(const uint8_t* d, size_t n, std::vector<uint8_t>& out) {
    static const uint8_t zero4[4] = {0, 0, 0, 0};
    out.clear();
    size_t pos = 0;
    bool have_serial = false, in_pkt = false, meta_done = false;
    uint32_t serial = 0, expect_seq = 0;
    uint64_t npackets = 0;
    std::vector<uint8_t> pkt;
    auto resync = [&](size_t from) {                     // next capture pattern at or after `from`
        while (from + 4 <= n) {
            const uint8_t* q = (const uint8_t*)memchr(d + from, 'O', n - from - 3);
            if (!q) return n;
            if (!memcmp(q, "OggS", 4)) return (size_t)(q - d);
            from = (size_t)(q - d) + 1;
        }
        return n;
    };
    while (pos + 27 <= n) {
        if (memcmp(d + pos, "OggS", 4)) { pos = resync(pos + 1); in_pkt = false; pkt.clear(); continue; }
        const uint32_t nseg = d[pos + 26];
        if (pos + 27 + nseg > n) break;
        size_t body = 0; for (uint32_t i = 0; i < nseg; i++) body += d[pos + 27 + i];
        const size_t total = 27 + nseg + body;
        if (pos + total > n) {                             // the stream ends inside this page -- or the segment table is garbage
            const size_t nx = resync(pos + 1);
            if (nx >= n) break;
            pos = nx; in_pkt = false; pkt.clear(); continue;
        }
        uint32_t crc = ogg_crc(0, d + pos, 22);
        crc = ogg_crc(crc, zero4, 4);
        crc = ogg_crc(crc, d + pos + 26, total - 26);
        if (d[pos + 4] != 0 || crc != le32(d + pos + 22)) { pos = resync(pos + 1); in_pkt = false; pkt.clear(); continue; }
        const uint8_t flags = d[pos + 5];
        const uint32_t ser = le32(d + pos + 14), seq = le32(d + pos + 18);
        const uint8_t* p = d + pos + 27 + nseg;
        if (!have_serial) {                                // the first logical stream whose first packet is an Ogg FLAC header
            if (!(flags & 2) || body < 13 || p[0] != 0x7F || memcmp(p + 1, "FLAC", 4)) { pos += total; continue; }
            if (p[5] != 1) return BNFLAC_ERR_UNSUPPORTED;  // mapping major version
            have_serial = true; serial = ser; expect_seq = seq;
        }
        if (ser != serial) { pos += total; continue; }     // another logical stream multiplexed in
        if (seq != expect_seq) { in_pkt = false; pkt.clear(); }       // pages are missing: the packet they carried is lost
        expect_seq = seq + 1;
        bool skipping = (flags & 1) && !in_pkt;            // the rest of a packet whose beginning is lost
        if (!(flags & 1) && in_pkt) { in_pkt = false; pkt.clear(); }
        try {
            for (uint32_t i = 0; i < nseg; i++) {
                const uint32_t L = d[pos + 27 + i];
                if (skipping) { p += L; if (L < 255) skipping = false; continue; }
                pkt.insert(pkt.end(), p, p + L); p += L; in_pkt = true;
                if (L == 255) continue;
                if (npackets == 0) {                       // 0x7F "FLAC" major minor nheaders(2) "fLaC" STREAMINFO block
                    if (pkt.size() < 13 + 4 + 34 || memcmp(pkt.data() + 9, "fLaC", 4)) return BNFLAC_ERR_NOT_FLAC;
                    out.insert(out.end(), pkt.begin() + 9, pkt.end());
                    meta_done = (pkt[13] & 0x80) != 0;
                } else if (!pkt.empty()) {
                    if (!meta_done) meta_done = (pkt[0] & 0x80) != 0;      // one metadata block per header packet
                    out.insert(out.end(), pkt.begin(), pkt.end());        // ... one frame per audio packet
                }
                npackets++;
                pkt.clear(); in_pkt = false;
            }
        } catch (...) { return BNFLAC_ERR_MEMORY; }
        pos += total;
        if (flags & 4) break;                              // last page of the logical stream
    }
    return have_serial && npackets ? 0 : BNFLAC_ERR_NOT_FLAC;
}
} // namespace

int bnflac_open_memory(const uint8_t* data, size_t len, const bnflac_opts* opts, bnflac_t** out) {
    if (!data || !out) return BNFLAC_ERR_ARG;
    int caller_dev = -1; if (cudaGetDevice(&caller_dev) != cudaSuccess) { cudaGetLastError(); caller_dev = -1; }
    struct Restore { int d; ~Restore() { if (d >= 0) { int cur = -1; if (cudaGetDevice(&cur) == cudaSuccess && cur != d) cudaSetDevice(d); } } } restore{caller_dev};
    *out = nullptr;
    bnflac* h = new (std::nothrow) bnflac; if (!h) return BNFLAC_ERR_MEMORY;
    h->opts = default_opts(opts); h->len = len;
    int rc;
    bool in_place = (h->opts.flags & BNFLAC_OPT_BORROW_INPUT) != 0;       // `data` is used where it lies
    if (is_ogg(data, len)) {                                              // Ogg FLAC: the de-paged copy is what is decoded
        if ((rc = ogg_depage(data, len, h->host))) { delete h; return rc; }
        data = h->host.data(); len = h->len = h->host.size();
        in_place = true;
    }
    bnflac_info_t probe; rc = parse_metadata(data, len, &probe);          // fail before touching the device
    if (rc) { delete h; return rc; }
    if (in_place) h->host_ptr = data;
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
    int caller_dev = -1; if (cudaGetDevice(&caller_dev) != cudaSuccess) { cudaGetLastError(); caller_dev = -1; }
    struct Restore { int d; ~Restore() { if (d >= 0) { int cur = -1; if (cudaGetDevice(&cur) == cudaSuccess && cur != d) cudaSetDevice(d); } } } restore{caller_dev};
    *out = nullptr;
    const bnflac_opts o = default_opts(opts);
    const bool lazy = (o.flags & BNFLAC_OPT_LAZY_PULL) != 0;
    if (lazy && o.shard_count > 1) return BNFLAC_ERR_ARG;        // a shard needs the stream's length
    bnflac* h = new (std::nothrow) bnflac; if (!h) return BNFLAC_ERR_MEMORY;
    h->opts = o; h->pull_cb = read; h->pull_user = user; h->pull_eof = false;
    // the reference pulls <= 16 KiB per callback (FLACDecoder.cs:21,336); pull_more asks for 1 MiB at a time
    int rc = 0;
    if (!lazy) rc = pull_more(h, ~0ull);                         // the whole stream now
    else for (uint64_t want = 64 << 10;; want *= 2) {            // only until the metadata is complete
        if ((rc = pull_more(h, want))) break;
        if (is_ogg(h->pulled.p, h->pulled.len)) { rc = pull_more(h, ~0ull); break; }      // Ogg FLAC is de-paged as a whole
        bnflac_info_t probe;
        rc = parse_metadata(h->pulled.p, h->pulled.len, &probe);
        if (rc != BNFLAC_ERR_TRUNCATED || h->pull_eof) { rc = 0; break; }                  // (common_open reports what is wrong)
    }
    if (!rc && is_ogg(h->pulled.p, h->pulled.len)) {
        if (!(rc = ogg_depage(h->pulled.p, h->pulled.len, h->host))) { h->pulled.release(); h->host_ptr = h->host.data(); h->len = h->host.size(); }
    }
    if (!rc) rc = common_open(h, h->host_ptr, h->len);
    if (!rc) rc = setup_device(h);
    if (rc) { delete h; return rc; }
    *out = h;
    return 0;
}

int bnflac_open_device(const void* d_data, size_t len, const uint8_t* header, size_t header_len, const bnflac_opts* opts, bnflac_t** out) {
    if (!d_data || !header || !out) return BNFLAC_ERR_ARG;
    int caller_dev = -1; if (cudaGetDevice(&caller_dev) != cudaSuccess) { cudaGetLastError(); caller_dev = -1; }
    struct Restore { int d; ~Restore() { if (d >= 0) { int cur = -1; if (cudaGetDevice(&cur) == cudaSuccess && cur != d) cudaSetDevice(d); } } } restore{caller_dev};
    *out = nullptr;
    bnflac* h = new (std::nothrow) bnflac; if (!h) return BNFLAC_ERR_MEMORY;
    h->opts = default_opts(opts); h->len = len; h->d_ext = (const uint8_t*)d_data;
    int rc;
    if ((rc = common_open(h, header, std::min(header_len, len))) || (rc = setup_device(h))) { delete h; return rc; }
    *out = h;
    return 0;
}

int bnflac_ogg_to_native(const uint8_t* data, size_t len, uint8_t* dst, size_t cap, size_t* written) {
    if (!data || !written || (!dst && cap)) return BNFLAC_ERR_ARG;
    *written = 0;
    if (!is_ogg(data, len)) return BNFLAC_ERR_NOT_FLAC;
    std::vector<uint8_t> native;
    int rc = ogg_depage(data, len, native); if (rc) return rc;
    *written = native.size();
    if (native.size() > cap) return BNFLAC_ERR_CAPACITY;
    if (!native.empty()) memcpy(dst, native.data(), native.size());
    return 0;
}

int bnflac_probe(const uint8_t* data, size_t len, bnflac_info_t* info) {
    if (!data || !info) return BNFLAC_ERR_ARG;
    if (is_ogg(data, len)) {
        std::vector<uint8_t> native;
        int rc = ogg_depage(data, len, native); if (rc) return rc;
        rc = parse_metadata(native.data(), native.size(), info);
        return rc;
    }
    return parse_metadata(data, len, info);
}

int bnflac_shard_range(uint64_t len, uint64_t first_frame_offset, uint32_t index, uint32_t count, uint64_t* own_begin, uint64_t* own_end) {
    if (!own_begin || !own_end || first_frame_offset > len || (count && index >= count)) return BNFLAC_ERR_ARG;
    shard_range(len, first_frame_offset, index, count, own_begin, own_end);
    return 0;
}

int bnflac_info(bnflac_t* h, bnflac_info_t* info) { if (!h || !info) return BNFLAC_ERR_ARG; *info = h->info; return 0; }
int bnflac_state(bnflac_t* h) { return h ? h->state : BNFLAC_STATE_UNINITIALIZED; }
void bnflac_close(bnflac_t* h) { if (!h) return; DeviceScope on(h->device); delete h; }

int bnflac_decode_device(bnflac_t* h, void* d_dst, size_t cap, void** d_out, uint64_t* written) {
    if (!h) return BNFLAC_ERR_ARG;
    DeviceScope on(h->device);
    for (bnflac* k : h->kids) delete k;
    h->kids.clear();
    h->rd_active = false; h->pl_session = false;
    { int rc = pull_all(h); if (rc) return rc; }
    return decode_to_device(h, d_dst, cap, d_out, written);
}

int bnflac_decoded_size(bnflac_t* h, uint64_t* bytes) {
    if (!h || !bytes) return BNFLAC_ERR_ARG;
    DeviceScope on(h->device);
    { int rc = pull_all(h); if (rc) return rc; }
    if (h->info.total_samples && (h->opts.shard_count <= 1)) {
        // STREAMINFO states it; a damaged stream may decode to less, never to more than its frame count allows.  Large
        // host streams take the pipelined path, where the exact figure is only known at the end.
        if (!h->size_understated && h->host_ptr && !h->d_ext && pipe_cuts(std::max<uint64_t>(h->own_begin, h->info.first_frame_offset), h->own_end).size() > 2) { *bytes = h->info.pcm_bytes; return 0; }
    }
    int rc = run_front(h); if (rc) return rc;
    *bytes = h->totals.pcm_bytes;
    return 0;
}

int bnflac_decode_all(bnflac_t* h, uint8_t* dst, size_t cap, uint64_t* written) {
    if (!h || !dst) return BNFLAC_ERR_ARG;
    DeviceScope on(h->device);
    uint64_t w = 0;
    int rc = decode_host(h, dst, cap, &w);
    // libFLAC decodes every frame whatever STREAMINFO.total_samples says: when a buffer of the promised size is too small, the
    // next bnflac_decoded_size scans (exact figure) instead of quoting STREAMINFO, so the caller can size and retry
    if (rc == BNFLAC_ERR_CAPACITY && cap >= h->info.pcm_bytes) h->size_understated = true;
    if (rc) return rc;
    if (written) *written = w;
    h->state = BNFLAC_STATE_END_OF_STREAM;
    if (h->opts.flags & BNFLAC_OPT_VERIFY_MD5) {
        static const uint8_t zero[16] = {0};
        if (memcmp(h->info.md5, zero, 16) && (h->opts.shard_count <= 1)) {
            Md5 m; m.update(dst, (size_t)w); uint8_t d[16]; m.final(d);
            if (memcmp(d, h->info.md5, 16)) return BNFLAC_ERR_STATE;
        }
    }
    return 0;
}

int64_t bnflac_read(bnflac_t* h, uint8_t* dst, size_t count) {
    if (!h || (!dst && count)) return BNFLAC_ERR_ARG;
    DeviceScope on(h->device);
    if (h->rd_active) return stream_read(h, dst, count);
    if (!h->decoded && h->pull_cb && !h->pull_eof) {       // lazily pulled source: sub-shards are cut as the bytes arrive
        CK(cudaSetDevice(h->device));
        for (bnflac* k : h->kids) delete k;
        h->kids.clear();
        h->rd_active = true; h->rd_issued = h->rd_cur = 0; h->rd_off = h->rd_total = 0; h->pl_size = 0; h->pl_session = true; h->pl_ahead = false; h->timing = bnflac_timing{}; h->diag_valid = false; h->diag_started = false;
        return stream_read(h, dst, count);
    }
    if (!h->decoded && h->host_ptr && !h->d_ext && h->batch_segs.empty()) {
        // large host-resident stream: decode ahead of the reader in sub-shards instead of all at once
        const std::vector<uint64_t> cuts = read_cuts(h, std::max<uint64_t>(h->own_begin, h->info.first_frame_offset), h->own_end);
        if (cuts.size() > 2) {
            CK(cudaSetDevice(h->device));
            int rc = make_kids(h, cuts); if (rc) return rc;
            h->rd_active = true; h->rd_issued = h->rd_cur = 0; h->rd_off = h->rd_total = 0; h->pl_session = false; h->timing = bnflac_timing{}; h->diag_valid = false; h->diag_started = false;
            return stream_read(h, dst, count);
        }
    }
    if (!h->decoded) {
        // the whole (shard of the) stream is decoded at the first Read into pinned host memory; later Reads are memcpy.
        // Capacity: what STREAMINFO promises, or (unknown length / damaged stream) what the frame scan finds.
        CK(cudaSetDevice(h->device));
        uint64_t need = 0; int rc;
        if ((rc = bnflac_decoded_size(h, &need))) return rc;
        if ((rc = h->pcm_host.reserve((size_t)need + 64))) return rc;
        uint64_t w = 0;
        rc = decode_host(h, (uint8_t*)h->pcm_host.p, h->pcm_host.cap, &w);
        if (rc == BNFLAC_ERR_CAPACITY) {           // STREAMINFO understated the stream: size it by scanning
            for (bnflac* k : h->kids) delete k;
            h->kids.clear();
            if ((rc = run_front(h))) return rc;
            if ((rc = h->pcm_host.reserve((size_t)h->totals.pcm_bytes + 64))) return rc;
            rc = decode_host_single(h, (uint8_t*)h->pcm_host.p, h->pcm_host.cap, &w);
        }
        if (rc) return rc;
        h->pcm_len = w; h->read_pos = 0; h->decoded = true;
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
    if ((!clips && n) || (!dst && cap)) return BNFLAC_ERR_ARG;
    if (written) *written = 0;
    if (!n) return 0;
    int caller_dev = -1; if (cudaGetDevice(&caller_dev) != cudaSuccess) { cudaGetLastError(); caller_dev = -1; }
    const int rc = decode_batch_impl(clips, n, opts, dst, cap, dst_is_device, results, written);
    if (caller_dev >= 0) { int cur = -1; if (cudaGetDevice(&cur) == cudaSuccess && cur != caller_dev) cudaSetDevice(caller_dev); }
    return rc;
}

int bnflac_frames(bnflac_t* h, const bnflac_frame_t** frames, size_t* n) {
    if (!h || !frames || !n) return BNFLAC_ERR_ARG;
    DeviceScope on(h->device);
    int rc = fetch_diag(h); if (rc) return rc;
    *frames = h->frames.data(); *n = h->frames.size();
    return 0;
}
int bnflac_subframes(bnflac_t* h, const bnflac_subframe_t** sub, size_t* n) {
    if (!h || !sub || !n) return BNFLAC_ERR_ARG;
    DeviceScope on(h->device);
    int rc = fetch_diag(h); if (rc) return rc;
    *sub = h->subs.data(); *n = h->subs.size();
    return 0;
}
int bnflac_errors(bnflac_t* h, const uint32_t** codes, size_t* n) {
    if (!h || !codes || !n) return BNFLAC_ERR_ARG;
    DeviceScope on(h->device);
    int rc = fetch_diag(h); if (rc) return rc;
    *codes = h->errors.data(); *n = h->errors.size();
    return 0;
}
int bnflac_errors_so_far(bnflac_t* h, const uint32_t** codes, size_t* n) {
    if (!h || !codes || !n) return BNFLAC_ERR_ARG;
    DeviceScope on(h->device);
    int rc = fetch_diag(h, false); if (rc) return rc;
    *codes = h->errors.data(); *n = h->errors.size();
    return 0;
}
int bnflac_error_frames(bnflac_t* h, const uint32_t** at, size_t* n) {
    if (!h || !at || !n) return BNFLAC_ERR_ARG;
    DeviceScope on(h->device);
    int rc = fetch_diag(h); if (rc) return rc;
    *at = h->errors_at.data(); *n = h->errors_at.size();
    return 0;
}
int bnflac_last_timing(bnflac_t* h, bnflac_timing* t) { if (!h || !t) return BNFLAC_ERR_ARG; *t = h->timing; return 0; }

} // extern "C"
