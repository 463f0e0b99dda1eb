// enc_emu.cpp -- runs the encoder kernels (csrc/encoder_kernels.cuh) on the CPU through tools/emu/cuda_emu.h.  TEST TOOL: built by
// `make emu` into tools/emu/_build/libencemu.so and used by tests/test_encode_emu_cpu.py to check the kernels' logic against the oracle
// and the reference decoder without a GPU.  It mirrors encode_on_device() of csrc/encoder.cu (plan -> scan -> memset -> write).
#include "cuda_emu.h"
#include "../../birdnest/audio_b200/csrc/encoder_kernels.cuh"
#include <cstdlib>
using namespace bnfe;

extern "C" int enc_emu(const uint8_t* pcm, uint64_t total_samples, uint32_t ch, uint32_t bps, uint32_t bin, uint32_t bs, uint32_t sr,
                       uint32_t max_lpc, uint32_t prec, uint32_t min_po, uint32_t max_po, uint32_t stereo, uint32_t search,
                       uint8_t* out, uint64_t cap, uint64_t* written, uint32_t* minfs, uint32_t* maxfs) {
    const uint32_t nframes = (uint32_t)((total_samples + bs - 1) / bs);
    std::vector<EncSub> sub((size_t)nframes * 8);
    std::vector<EncFrame> frm(nframes);
    EncTotals tot{};
    EncArgs a{};
    a.pcm = pcm; a.total_samples = total_samples; a.ch = ch; a.bps = bps; a.bin = bin; a.bs = bs; a.sample_rate = sr;
    a.max_lpc = max_lpc; a.prec = prec; a.min_po = min_po; a.max_po = max_po; a.stereo = stereo; a.search_order = search;
    a.nframes = nframes; a.first_frame = 42; a.first_number = 0; a.sub = sub.data(); a.frm = frm.data(); a.totals = &tot; a.out = out;
    const bool big = bs >= NT_BIG_FROM_BS;
    if (big) emu_launch(k_enc_plan<NT_BIG>, nframes, NT_BIG, enc_smem_bytes(bs), a); else emu_launch(k_enc_plan<NT_SMALL>, nframes, NT_SMALL, enc_smem_bytes(bs), a);
    emu_launch(k_enc_scan, 1, 1024, 0, a);
    if (tot.total_bytes + 4 > cap) return -7;
    memset(out, 0, (size_t)((tot.total_bytes + 3) & ~3ull));
    if (big) emu_launch(k_enc_write<NT_BIG>, nframes, NT_BIG, enc_smem_bytes(bs), a); else emu_launch(k_enc_write<NT_SMALL>, nframes, NT_SMALL, enc_smem_bytes(bs), a);
    *written = tot.total_bytes; *minfs = tot.min_fs; *maxfs = tot.max_fs;
    return 0;
}
