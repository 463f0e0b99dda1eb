// kernels_decode_wide.cu -- k_decode variants with 64-bit accumulation (libFLAC's width rule) + the public launcher.
#include "kernels_decode.cuh"

namespace bnf {

void launch_decode_narrow(const PassArgs& a, uint32_t nacc, uint32_t C, uint32_t B, uint32_t max_order, void* stream);

// at least 8 warps are resident per SM whatever the variant (168 registers: 12), at most 32
uint32_t decode_sched_slots(uint32_t nacc_bound, uint32_t channels) {
    static const bool off = getenv("BNFLAC_NO_BALANCE") != nullptr;
    if (off || channels == 0 || channels > 32) return 0;
    const uint32_t F = 32 / channels;
    const uint32_t n_sm = (uint32_t)sm_count();
    if (const uint32_t cap = dec_force_slots()) { if (blocks_for(nacc_bound, F) <= std::max<uint32_t>(cap, 1u)) return 0; }
    else if (blocks_for(nacc_bound, F) <= n_sm * 8u) return 0;
    return n_sm * 32u;
}
uint64_t decode_sched_state_bytes() { return (uint64_t)DEC_SLOT_WORDS * 4u; }

void launch_decode(const PassArgs& a, uint32_t nacc, uint32_t C, uint32_t B, uint32_t max_order, bool wide, void* stream) {
    if (wide) launch_decode_w<true>(a, nacc, C, B, max_order, S(stream));
    else launch_decode_narrow(a, nacc, C, B, max_order, stream);
}

} // namespace bnf
