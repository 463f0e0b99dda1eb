// kernels_decode_narrow.cu -- k_decode variants with 32-bit accumulation (16-bit streams, FIXED predictors).
#include "kernels_decode.cuh"

namespace bnf {

void launch_decode_narrow(const PassArgs& a, uint32_t nacc, uint32_t C, uint32_t B, uint32_t max_order, void* stream) {
    launch_decode_w<false>(a, nacc, C, B, max_order, S(stream));
}

} // namespace bnf
