// ulat.cu -- dependent-issue latency of the ops the serial bit walkers are made of (B200, sm_100a): one warp, one chain.
// Prints cycles per op in a chain where every instruction consumes the previous result.
// Build: nvcc -arch=sm_100a -O3 -o tools/ulat tools/ulat.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N_ITER 4096
#define UNROLL 16

template <int OP>
__global__ void __launch_bounds__(32) k(uint32_t* out, uint32_t seed, long long* cyc) {
    __shared__ uint32_t sm[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) sm[i] = (i * 4 + 4) & 4095;     // pointer-chase table: next byte offset
    uint32_t a = seed + threadIdx.x, c = seed | 1, s = (seed & 7) + 1;
    double d = (double)a, dc = 1.0000001;
    float f = (float)a, fc = 1.0001f;
    uint64_t w = a;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
    __syncwarp();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < N_ITER; it++) {
#pragma unroll
        for (int i = 0; i < UNROLL; i++) {
            if (OP == 0) asm volatile("shf.l.wrap.b32 %0, %0, %1, %0;" : "+r"(a) : "r"(c));
            if (OP == 1) asm volatile("bfind.u32 %0, %0;" : "+r"(a));
            if (OP == 2) asm volatile("add.s32 %0, %0, %1;" : "+r"(a) : "r"(c));
            if (OP == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(c), "r"(s));
            if (OP == 4) asm volatile("prmt.b32 %0, %0, %1, 0x0123;" : "+r"(a) : "r"(c));
            if (OP == 5) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a) : "r"(c), "r"(s));
            if (OP == 6) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d) : "d"(dc));
            if (OP == 7) asm volatile("add.rm.f64 %0, %0, %1;" : "+d"(d) : "d"(dc));
            if (OP == 8) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a) : "r"(sbase + (a & 4092u)));
            if (OP == 9) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f) : "f"(fc));
            if (OP == 10) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w) : "r"(c), "r"(s));
            if (OP == 11) { asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\tselp.b32 %0, %1, %2, p;\n\t}" : "+r"(a) : "r"(c), "r"(s)); }
            if (OP == 12) asm volatile("popc.b32 %0, %0;" : "+r"(a));
            if (OP == 13) asm volatile("shl.b32 %0, %0, 1;" : "+r"(a));
            if (OP == 14) asm volatile("cvt.rn.f32.u32 %0, %1;\n\tmov.b32 %1, %0;" : "+f"(f), "+r"(a));
            if (OP == 15) asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(a) : "r"(s));
            if (OP == 16) { asm volatile("shf.l.wrap.b32 %0, %0, %1, %0;" : "+r"(a) : "r"(c)); asm volatile("bfind.u32 %0, %0;" : "+r"(a)); asm volatile("sub.s32 %0, %1, %0;" : "+r"(a) : "r"(c)); }   // the walker's chain: SHF, FLO, IADD
            if (OP == 17) asm volatile("brev.b32 %0, %0;" : "+r"(a));
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = a ^ (uint32_t)d ^ (uint32_t)f ^ (uint32_t)w ^ (uint32_t)(w >> 32);
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int OP> void run(const char* name, int per) {
    uint32_t* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    k<OP><<<1, 32>>>(out, 12345, cyc); cudaDeviceSynchronize();
    k<OP><<<1, 32>>>(out, 12345, cyc); cudaDeviceSynchronize();
    long long c = 0; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s %.2f cycles per dependent op\n", name, (double)c / ((double)N_ITER * UNROLL * per));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    asm volatile("" ::: "memory");
    run<0>("SHF (funnel shift)", 1); run<1>("FLO (bfind)", 1); run<2>("IADD", 1); run<3>("LOP3", 1); run<4>("PRMT", 1); run<5>("IMAD.lo", 1);
    run<6>("DFMA", 1); run<7>("DADD.RM", 1); run<8>("LDS (pointer chase)", 1); run<9>("FFMA", 1); run<10>("IMAD.WIDE (acc chain)", 1);
    run<11>("ISETP+SEL", 1); run<12>("POPC", 1); run<13>("SHL", 1); run<14>("I2F.U32", 1); run<15>("SHFL.IDX", 1); run<16>("SHF+FLO+IADD (walker chain)", 3); run<17>("BREV", 1);
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
