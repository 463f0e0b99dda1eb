// encoder.cu -- SURVEY 8f-4: the FLAC *encoder* of the native codec, B200-native.
//
// Reference surface: Library/LibFLACSharp/LibFLACSharp.cs:322-387 declares libFLAC's stream encoder (new / set_channels /
// set_bits_per_sample / set_sample_rate / set_compression_level / set_blocksize / set_do_mid_side_stereo / init_stream /
// process_interleaved / process / finish / delete) and the write callback it drives.  The C# never calls it, the DLL carries it;
// SURVEY 8f ranks it last ("lets corpus generation run on the GPU").  This file is the engine behind bnflac_encode* (include/bnflac.h);
// csrc/libflac_shim.cpp replays it frame by frame behind the legacy FLAC__stream_encoder_* symbols.
//
// A frame is the parallel unit (frames are independent: own header, own warm-up samples, own CRC).  Three launches per stream:
//   k_enc_plan   CTA per frame.  For every channel (stereo: L, R, M=(L+R)>>1, S=L-R) the samples are staged in shared memory and the
//                CTA works out wasted bits / CONSTANT, the best FIXED order (sum of |residual| of orders 0..4), one LPC candidate
//                (Welch window, autocorrelation in FP64 spread over the threads, Levinson-Durbin + order choice + quantisation by one
//                thread), the exact residuals of both, and for each the cheapest partitioned-Rice plan (every partition order of
//                the allowed range, three Rice parameters around log2(mean) per partition, exact bit counts) -- then keeps the
//                smallest of VERBATIM / FIXED / LPC.  Stereo frames take the cheapest of the four assignments.  Output: one 336-byte
//                decision record per (frame, channel), the frame header bytes (CRC-8 included) and the frame's exact size.
//   k_enc_scan   one CTA: exclusive prefix sum of the frame sizes -> byte offset of every frame, min / max frame size, total.
//   k_enc_write  CTA per frame.  Residuals are recomputed from the decision record; every thread owns a run of consecutive samples,
//                a block-wide prefix sum of the runs' bit counts gives each run its bit offset, and the thread writes its codewords
//                with plain 32-bit stores (atomicOr only for the first and last word of a run, which neighbours share).  The frame's
//                CRC-16 is computed by all threads (chunk CRCs combined with x^(8n) mod P) once the body is in place.
// Nothing is staged through a per-frame scratch and nothing is copied afterwards: the second kernel writes every frame where it
// finally lives.  HBM traffic: the PCM is read once per candidate channel in the plan and once in the write (L2 serves most of
// the second), the stream is written once (plus the memset that zeroes it).
//
// The encoder is free in its choices (any valid stream that decodes to the same PCM is correct), so parity here means: the
// reference DECODER (LibFlac.dll under oracle/refdll), the oracle and the GPU decoder all return the input PCM bit for bit and the
// STREAMINFO MD5 matches -- tests/test_encode_gpu.py.
#include "../../../include/bnflac.h"
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <cstdio>
#include <new>
#include <string>
#include <thread>
#include <vector>
#include <algorithm>

namespace bnf { void count_launch(); void set_cuda_error(const char* what); }      // kernels.cu, engine.cu

#include "encoder_kernels.cuh"
#include "md5.hpp"

namespace bnfe {

// ------------------------------------------------------------------------------------------------ host side
static int cuda_fail(const char* what, cudaError_t e) {
    bnf::set_cuda_error((std::string(what) + ": " + cudaGetErrorString(e)).c_str());
    cudaGetLastError();
    return BNFLAC_ERR_CUDA;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = cuda_fail(#call, e_); goto done; } } while (0)

struct Resolved { uint32_t ch, bps, bin, bs, sr, max_lpc, prec, min_po, max_po, stereo, search, flags; uint64_t first_number; };

// libFLAC's presets (the reference binds set_compression_level, LibFLACSharp.cs:342-343)
static const struct { uint32_t bs, lpc, ms, po; } kLevels[9] = {
    {1152, 0, 0, 3}, {1152, 0, 1, 3}, {1152, 0, 1, 3}, {4096, 6, 0, 4}, {4096, 8, 1, 4}, {4096, 8, 1, 5}, {4096, 8, 1, 6}, {4096, 8, 1, 6}, {4096, 12, 1, 6}};

static int resolve(const bnflac_enc_opts* o, Resolved* r) {
    if (!o || o->struct_size < sizeof(bnflac_enc_opts)) return BNFLAC_ERR_ARG;
    r->ch = o->channels; r->bps = o->bits_per_sample; r->sr = o->sample_rate;
    if (r->ch < 1 || r->ch > 8 || r->bps < 4 || r->bps > 24 || r->sr < 1 || r->sr > 655350) return BNFLAC_ERR_UNSUPPORTED;
    r->flags = o->flags;
    r->bin = (o->flags & BNFLAC_ENC_INPUT_INT32) ? 4u : (r->bps + 7) / 8;
    const uint32_t lvl = o->compression_level > 8 ? 8 : o->compression_level;
    const bool preset = (o->flags & BNFLAC_ENC_USE_LEVEL) != 0;
    r->bs = o->blocksize ? o->blocksize : (preset ? kLevels[lvl].bs : 4096);
    r->max_lpc = preset ? kLevels[lvl].lpc : o->max_lpc_order;
    r->max_po = preset ? kLevels[lvl].po : o->max_partition_order;
    r->min_po = preset ? 0 : o->min_partition_order;
    r->stereo = (r->ch == 2 && (preset ? kLevels[lvl].ms : o->mid_side)) ? 1u : 0u;
    r->prec = o->qlp_precision;
    r->search = (o->flags & BNFLAC_ENC_FIXED_ORDER) ? 0u : 1u;
    r->first_number = o->first_frame_number;
    if (r->bs < 16 || r->bs > MAX_BS || r->max_lpc > 32 || r->max_po > (uint32_t)MAX_PO || r->min_po > r->max_po || r->prec > 15) return BNFLAC_ERR_UNSUPPORTED;
    return 0;
}

static uint64_t bound_bytes(const Resolved& r, uint64_t total_samples) {
    const uint64_t nframes = (total_samples + r.bs - 1) / r.bs;
    const uint64_t per = 16 + ((uint64_t)r.ch * (8 + 32 + (uint64_t)r.bs * (r.bps + 1)) + 7) / 8 + 2;
    return 42 + nframes * per + 64;
}

static void write_streaminfo(uint8_t* o, const Resolved& r, uint32_t minfs, uint32_t maxfs, uint64_t total, const uint8_t md5[16]) {
    size_t q = 0;
    memcpy(o, "fLaC", 4); q = 4;
    o[q++] = 0x80; o[q++] = 0; o[q++] = 0; o[q++] = 34;
    o[q++] = (uint8_t)(r.bs >> 8); o[q++] = (uint8_t)r.bs; o[q++] = (uint8_t)(r.bs >> 8); o[q++] = (uint8_t)r.bs;
    o[q++] = (uint8_t)(minfs >> 16); o[q++] = (uint8_t)(minfs >> 8); o[q++] = (uint8_t)minfs;
    o[q++] = (uint8_t)(maxfs >> 16); o[q++] = (uint8_t)(maxfs >> 8); o[q++] = (uint8_t)maxfs;
    const uint64_t x = ((uint64_t)r.sr << 44) | ((uint64_t)(r.ch - 1) << 41) | ((uint64_t)(r.bps - 1) << 36) | (total & 0xFFFFFFFFFull);
    for (int i = 7; i >= 0; i--) o[q++] = (uint8_t)(x >> (8 * i));
    memcpy(o + q, md5, 16);
}

// MD5 of the PCM in FLAC's layout (little-endian, ceil(bps/8) bytes per sample); int32 input is packed on the fly
static void md5_pcm(const uint8_t* pcm, uint64_t nsamp_all, const Resolved& r, uint8_t out[16]) {
    Md5 m;
    const uint32_t B = (r.bps + 7) / 8;
    if (r.bin == B) m.update(pcm, (size_t)(nsamp_all * B));
    else {
        std::vector<uint8_t> tmp(4096 * 3);
        for (uint64_t i = 0; i < nsamp_all;) {
            const uint64_t c = std::min<uint64_t>(4096, nsamp_all - i);
            for (uint64_t j = 0; j < c; j++) for (uint32_t b = 0; b < B; b++) tmp[j * B + b] = pcm[(i + j) * 4 + b];
            m.update(tmp.data(), (size_t)(c * B)); i += c;
        }
    }
    m.final(out);
}

// d_pcm: device PCM.  d_out: device buffer of out_cap bytes (4-byte aligned).  header42: receives "fLaC" + STREAMINFO (md5 zero).
static int encode_on_device(const Resolved& r, const uint8_t* d_pcm, uint64_t total_samples, uint8_t* d_out, uint64_t out_cap, uint64_t* written,
                            uint32_t* minfs, uint32_t* maxfs, bnflac_enc_stats* st, cudaStream_t stream) {
    int rc = 0;
    const uint32_t nframes = (uint32_t)((total_samples + r.bs - 1) / r.bs);
    EncSub* d_sub = nullptr; EncFrame* d_frm = nullptr; EncTotals* d_tot = nullptr;
    EncTotals tot{};
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    EncArgs a{};
    size_t smem_plan = 0, smem_write = 0;
    const bool big = r.bs >= NT_BIG_FROM_BS;       // large blocks: CTAs of 1024 threads (the staged samples leave room for one CTA per SM)
    if (((uintptr_t)d_out & 3u) != 0) return BNFLAC_ERR_ARG;
    if (!nframes) { *written = 42; *minfs = 0; *maxfs = 0; if (st) { st->frames = 0; st->bytes = 42; } return out_cap >= 42 ? 0 : BNFLAC_ERR_CAPACITY; }
    for (auto& e : ev) CK(cudaEventCreate(&e));
    CK(cudaMalloc(&d_sub, (size_t)nframes * 8 * sizeof(EncSub)));
    CK(cudaMalloc(&d_frm, (size_t)nframes * sizeof(EncFrame)));
    CK(cudaMalloc(&d_tot, sizeof(EncTotals)));
    a.pcm = d_pcm; a.total_samples = total_samples; a.ch = r.ch; a.bps = r.bps; a.bin = r.bin; a.bs = r.bs; a.sample_rate = r.sr;
    a.max_lpc = r.max_lpc; a.prec = r.prec; a.min_po = r.min_po; a.max_po = r.max_po; a.stereo = r.stereo; a.search_order = r.search;
    a.nframes = nframes; a.first_frame = 42; a.first_number = r.first_number; a.sub = d_sub; a.frm = d_frm; a.totals = d_tot; a.out = d_out;
    smem_plan = smem_write = enc_smem_bytes(r.bs);
    if (big) {
        CK(cudaFuncSetAttribute(k_enc_plan<NT_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_plan));
        CK(cudaFuncSetAttribute(k_enc_write<NT_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_write));
    } else {
        CK(cudaFuncSetAttribute(k_enc_plan<NT_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_plan));
        CK(cudaFuncSetAttribute(k_enc_write<NT_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_write));
    }
    CK(cudaEventRecord(ev[0], stream));
    if (big) k_enc_plan<NT_BIG><<<nframes, NT_BIG, smem_plan, stream>>>(a); else k_enc_plan<NT_SMALL><<<nframes, NT_SMALL, smem_plan, stream>>>(a);
    bnf::count_launch();
    k_enc_scan<<<1, 1024, 0, stream>>>(a); bnf::count_launch();
    CK(cudaGetLastError());
    CK(cudaEventRecord(ev[1], stream));
    CK(cudaMemcpyAsync(&tot, d_tot, sizeof tot, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (tot.total_bytes + 4 > out_cap) { rc = BNFLAC_ERR_CAPACITY; *written = tot.total_bytes; goto done; }
    CK(cudaMemsetAsync(d_out, 0, (size_t)((tot.total_bytes + 3) & ~3ull), stream));
    CK(cudaEventRecord(ev[2], stream));
    if (big) k_enc_write<NT_BIG><<<nframes, NT_BIG, smem_write, stream>>>(a); else k_enc_write<NT_SMALL><<<nframes, NT_SMALL, smem_write, stream>>>(a);
    bnf::count_launch();
    CK(cudaGetLastError());
    CK(cudaEventRecord(ev[3], stream));
    CK(cudaStreamSynchronize(stream));
    *written = tot.total_bytes; *minfs = tot.min_fs; *maxfs = tot.max_fs;
    if (st) {
        float t = 0;
        cudaEventElapsedTime(&t, ev[0], ev[1]); st->plan_ms = t;
        cudaEventElapsedTime(&t, ev[2], ev[3]); st->write_ms = t;
        cudaEventElapsedTime(&t, ev[0], ev[3]); st->total_ms = t;
        st->frames = nframes; st->bytes = tot.total_bytes; st->min_framesize = tot.min_fs; st->max_framesize = tot.max_fs;
        if (st->frame_sizes && st->frame_sizes_cap) {
            const size_t n = (size_t)std::min<uint64_t>(nframes, st->frame_sizes_cap);
            CK(cudaMemcpy2D(st->frame_sizes, 4, reinterpret_cast<const uint8_t*>(d_frm) + offsetof(EncFrame, nbytes), sizeof(EncFrame), 4, n, cudaMemcpyDeviceToHost));
        }
    }
done:
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    if (d_sub) cudaFree(d_sub);
    if (d_frm) cudaFree(d_frm);
    if (d_tot) cudaFree(d_tot);
    return rc;
}

struct DeviceScope {     // every entry point puts the caller's current device back
    int prev = -1; bool ok = true;
    explicit DeviceScope(int want) {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) { cudaGetLastError(); ok = false; return; }
        cudaGetDevice(&prev);
        if (want >= 0 && want != prev) { if (want >= n || cudaSetDevice(want) != cudaSuccess) { cudaGetLastError(); ok = false; } }
    }
    ~DeviceScope() { if (prev >= 0) { int cur = -1; cudaGetDevice(&cur); if (cur != prev) cudaSetDevice(prev); } }
};

} // namespace bnfe

using namespace bnfe;

extern "C" int bnflac_encode_bound(size_t pcm_bytes, const bnflac_enc_opts* opts, uint64_t* bound) {
    Resolved r;
    if (!bound) return BNFLAC_ERR_ARG;
    if (int e = resolve(opts, &r)) return e;
    if (pcm_bytes % ((size_t)r.bin * r.ch)) return BNFLAC_ERR_ARG;
    *bound = bound_bytes(r, pcm_bytes / ((size_t)r.bin * r.ch));
    return 0;
}

extern "C" int bnflac_encode_device(const void* d_pcm, size_t pcm_bytes, const bnflac_enc_opts* opts, void* d_dst, size_t cap, uint64_t* written, bnflac_enc_stats* stats) {
    Resolved r;
    if (!d_pcm || !d_dst || !written) return BNFLAC_ERR_ARG;
    if (int e = resolve(opts, &r)) return e;
    if (pcm_bytes % ((size_t)r.bin * r.ch)) return BNFLAC_ERR_ARG;
    if ((r.bin == 2 && ((uintptr_t)d_pcm & 1u)) || (r.bin == 4 && ((uintptr_t)d_pcm & 3u))) return BNFLAC_ERR_ARG;   // 16-bit / int32 samples are fetched as such
    DeviceScope ds(opts->device);
    if (!ds.ok) return BNFLAC_ERR_NO_DEVICE;
    const uint64_t total = pcm_bytes / ((size_t)r.bin * r.ch);
    uint32_t minfs = 0, maxfs = 0;
    int rc = encode_on_device(r, (const uint8_t*)d_pcm, total, (uint8_t*)d_dst, cap, written, &minfs, &maxfs, stats, (cudaStream_t)opts->stream);
    if (rc) return rc;
    uint8_t hdr[42], md5[16] = {0};
    if (!(r.flags & BNFLAC_ENC_NO_MD5)) {        // device-resident PCM: fetched for the (serial) MD5 only when asked for
        std::vector<uint8_t> host;
        try { host.resize(pcm_bytes); } catch (const std::bad_alloc&) { return BNFLAC_ERR_MEMORY; }      // never through the extern "C" frame
        if (cudaMemcpy(host.data(), d_pcm, pcm_bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return cuda_fail("cudaMemcpy(md5)", cudaGetLastError());
        md5_pcm(host.data(), total * r.ch, r, md5);
    }
    write_streaminfo(hdr, r, minfs, maxfs, total, md5);
    if (cudaMemcpy(d_dst, hdr, 42, cudaMemcpyHostToDevice) != cudaSuccess) return cuda_fail("cudaMemcpy(header)", cudaGetLastError());
    return 0;
}

extern "C" int bnflac_encode(const uint8_t* pcm, size_t pcm_bytes, const bnflac_enc_opts* opts, uint8_t* dst, size_t cap, uint64_t* written, bnflac_enc_stats* stats) {
    Resolved r;
    if ((!pcm && pcm_bytes) || !written) return BNFLAC_ERR_ARG;
    if (int e = resolve(opts, &r)) return e;
    if (pcm_bytes % ((size_t)r.bin * r.ch)) return BNFLAC_ERR_ARG;
    DeviceScope ds(opts->device);
    if (!ds.ok) return BNFLAC_ERR_NO_DEVICE;
    const uint64_t total = pcm_bytes / ((size_t)r.bin * r.ch);
    const uint64_t bound = bound_bytes(r, total);
    int rc = 0;
    uint8_t* d_pcm = nullptr; uint8_t* d_out = nullptr;
    uint8_t md5[16] = {0};
    uint32_t minfs = 0, maxfs = 0;
    uint64_t n = 0;
    std::thread hasher;
    if (!(r.flags & BNFLAC_ENC_NO_MD5)) hasher = std::thread([&] { md5_pcm(pcm, total * r.ch, r, md5); });   // serial by nature: runs beside the GPU
    CK(cudaMalloc(&d_pcm, pcm_bytes + 64));
    CK(cudaMalloc(&d_out, bound));
    CK(cudaMemcpy(d_pcm, pcm, pcm_bytes, cudaMemcpyHostToDevice));
    rc = encode_on_device(r, d_pcm, total, d_out, bound, &n, &minfs, &maxfs, stats, (cudaStream_t)opts->stream);
    if (rc) goto done;
    *written = n;
    if (!dst || n > cap) { rc = dst ? BNFLAC_ERR_CAPACITY : 0; goto done; }
    CK(cudaMemcpy(dst, d_out, n, cudaMemcpyDeviceToHost));
done:
    if (hasher.joinable()) hasher.join();
    if (!rc && dst && n <= cap && n >= 42) write_streaminfo(dst, r, minfs, maxfs, total, md5);
    if (d_pcm) cudaFree(d_pcm);
    if (d_out) cudaFree(d_out);
    return rc;
}
