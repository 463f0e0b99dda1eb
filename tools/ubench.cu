// ubench.cu -- issue-rate microbenchmarks for the integer ops the decode kernels are made of (B200, sm_100a).
// Prints warp-instructions per clock per SM sub-partition for each op / mix.  Build: nvcc -arch=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N_ITER 16384
#define CHAINS 8

template <int OP>
__global__ void __launch_bounds__(128) k(uint32_t* out, uint32_t seed, long long* cyc) {
    uint32_t a[CHAINS], b[CHAINS]; uint64_t w[CHAINS]; double d[CHAINS];
    for (int i = 0; i < CHAINS; i++) { b[i] = seed * 3 + i; a[i] = seed + i * 77 + threadIdx.x; w[i] = a[i]; d[i] = (double)a[i]; }
    uint32_t c = seed | 1, s = (seed & 7) + 1;
    double dc = (double)c;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < N_ITER; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (OP == 0) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(s));
            if (OP == 1) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(c));
            if (OP == 2) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(dc));
            if (OP == 3) asm volatile("clz.b32 %0, %0;" : "+r"(a[i]));
            if (OP == 4) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(s));
            if (OP == 5) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(c), "r"(s));
            if (OP == 6) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(s));
            if (OP == 7) asm volatile("bfe.s32 %0, %0, 0, 1;" : "+r"(a[i]));
            if (OP == 8) asm volatile("add.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(c));
            if (OP == 9) { asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(c)); asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(c), "r"(s)); }
            if (OP == 10) { asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(s)); asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(c), "r"(s)); }
            if (OP == 11) { asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(d[i]) : "r"(a[i])); asm volatile("cvt.rmi.s32.f64 %0, %1;" : "=r"(a[i]) : "d"(d[i])); }
            if (OP == 12) { asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(dc)); asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(c), "r"(s)); }
            if (OP == 13) { asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(dc)); asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(s)); }
            if (OP == 14) { asm volatile("mul.hi.s32 %0, %0, %1;" : "+r"(a[i]) : "r"(c)); }
            if (OP == 15) { asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(c), "r"(s)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(c), "r"(s)); asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(s)); }
            if (OP == 16) { asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(c)); asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(c), "r"(s)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(c), "r"(s)); }
            if (OP == 17) { asm volatile("clz.b32 %0, %0;" : "+r"(a[i])); asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(c), "r"(s)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(c), "r"(s)); }
        }
    }
    long long t1 = clock64();
    uint32_t r = 0;
    for (int i = 0; i < CHAINS; i++) r ^= a[i] ^ b[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ (uint32_t)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP> void run(const char* name, int per_iter) {
    uint32_t* out; long long* cyc; cudaMalloc(&out, 148 * 8 * 256 * 4); cudaMalloc(&cyc, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps_per_smsp : {1, 2, 4, 8, 12}) {
        int threads = 128;                       // one warp per SMSP per CTA
        int ctas_per_sm = warps_per_smsp;
        k<OP><<<148 * ctas_per_sm, threads>>>(out, 12345, cyc);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        k<OP><<<148 * ctas_per_sm, threads>>>(out, 12345, cyc);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double instr = (double)N_ITER * CHAINS * per_iter * warps_per_smsp;   // per SMSP
        double clk = ms * 1e-3 * 1.965e9;
        printf("%-28s warps/SMSP=%-2d  ms=%.3f  warp-instr/clk/SMSP=%.3f\n", name, warps_per_smsp, ms, instr / clk);
    }
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("IMAD.lo", 1); run<1>("IMAD.WIDE", 1); run<2>("DFMA", 1); run<3>("FLO(clz)", 1); run<4>("SHF", 1); run<5>("LOP3", 1);
    run<6>("PRMT", 1); run<7>("SGXT(bfe.s32 0,1)", 1); run<8>("IADD", 1); run<9>("IMAD.WIDE+SHF", 2); run<10>("IMAD.lo+SHF", 2);
    run<11>("I2F.F64+F2I.F64", 2); run<12>("DFMA+SHF", 2); run<13>("DFMA+IMAD", 2); run<14>("IMUL.HI", 1); run<15>("SHF+LOP3+IMAD", 3); run<16>("IMAD.WIDE+SHF+LOP3", 3); run<17>("FLO+SHF+LOP3", 3);
    cudaError_t e = cudaGetLastError(); printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
