// kernels_decode_wide.cu -- k_decode variants with 64-bit accumulation (libFLAC's width rule) + the public launcher.
#include "kernels_decode.cuh"

namespace bnf {

void launch_decode_narrow(const PassArgs& a, uint32_t nacc, uint32_t C, uint32_t B, uint32_t max_order, void* stream);

void launch_decode(const PassArgs& a, uint32_t nacc, uint32_t C, uint32_t B, uint32_t max_order, bool wide, void* stream) {
    if (wide) launch_decode_w<true>(a, nacc, C, B, max_order, S(stream));
    else launch_decode_narrow(a, nacc, C, B, max_order, stream);
}

} // namespace bnf
