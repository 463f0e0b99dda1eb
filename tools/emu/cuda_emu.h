// cuda_emu.h -- just enough of the CUDA execution model to run csrc/encoder_kernels.cuh on the CPU: one pthread per CUDA thread,
// a pthread barrier per CTA (__syncthreads) and per warp (shuffles / warp reductions exchange through a per-warp buffer), CTAs run
// one after the other, `__shared__` variables are function-local statics.  TEST TOOL (tools/emu/enc_emu.cpp): lets the encoder
// kernels be checked against the oracle and the reference decoder in a container without a GPU; never part of the product.
#pragma once
#include <pthread.h>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <vector>

#define BNFLAC_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)

struct EmuIdx { uint32_t x, y, z; };
static thread_local EmuIdx threadIdx, blockIdx;
static uint8_t* emu_dyn_smem;
static pthread_barrier_t emu_block_bar;
struct EmuWarp { pthread_barrier_t bar; uint64_t xch[32]; };
static EmuWarp* emu_warps;

using std::min;
using std::max;

static inline void __syncthreads() { pthread_barrier_wait(&emu_block_bar); }
static inline void __threadfence() { __sync_synchronize(); }
static inline EmuWarp& emu_w() { return emu_warps[threadIdx.x >> 5]; }
template <class T> static inline uint64_t emu_bits(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof v); return b; }
template <class T> static inline T emu_val(uint64_t b) { T v; memcpy(&v, &b, sizeof v); return v; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int o) {
    EmuWarp& w = emu_w(); const uint32_t l = threadIdx.x & 31;
    w.xch[l] = emu_bits(v); pthread_barrier_wait(&w.bar);
    const T r = emu_val<T>(w.xch[l ^ (uint32_t)o]); pthread_barrier_wait(&w.bar);
    return r;
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, int o) {
    EmuWarp& w = emu_w(); const uint32_t l = threadIdx.x & 31;
    w.xch[l] = emu_bits(v); pthread_barrier_wait(&w.bar);
    const T r = l >= (uint32_t)o ? emu_val<T>(w.xch[l - (uint32_t)o]) : v; pthread_barrier_wait(&w.bar);
    return r;
}
template <class F> static inline uint32_t emu_reduce(uint32_t v, F f) {
    EmuWarp& w = emu_w(); const uint32_t l = threadIdx.x & 31;
    w.xch[l] = v; pthread_barrier_wait(&w.bar);
    uint32_t r = (uint32_t)w.xch[0];
    for (int i = 1; i < 32; i++) r = f(r, (uint32_t)w.xch[i]);
    pthread_barrier_wait(&w.bar);
    return r;
}
static inline uint32_t __reduce_or_sync(unsigned, uint32_t v) { return emu_reduce(v, [](uint32_t a, uint32_t b) { return a | b; }); }
static inline uint32_t __reduce_min_sync(unsigned, uint32_t v) { return emu_reduce(v, [](uint32_t a, uint32_t b) { return a < b ? a : b; }); }
static inline uint32_t __reduce_max_sync(unsigned, uint32_t v) { return emu_reduce(v, [](uint32_t a, uint32_t b) { return a > b ? a : b; }); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline uint32_t atomicOr(uint32_t* p, uint32_t v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *(const volatile T*)p; }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
    const uint64_t t = (uint64_t)y << 32 | x; uint32_t r = 0;
    for (int i = 0; i < 4; i++) { const uint32_t sel = (s >> (4 * i)) & 7; r |= (uint32_t)((t >> (8 * sel)) & 0xff) << (8 * i); }
    return r;
}

template <class Args> struct EmuLaunch { void (*k)(Args); Args a; uint32_t tid, bid; };
template <class Args> static void* emu_thread(void* p) {
    auto* l = static_cast<EmuLaunch<Args>*>(p);
    threadIdx = EmuIdx{l->tid, 0, 0}; blockIdx = EmuIdx{l->bid, 0, 0};
    l->k(l->a);
    return nullptr;
}
template <class Args> static void emu_launch(void (*k)(Args), uint32_t grid, uint32_t block, size_t smem, const Args& a) {
    std::vector<uint8_t> dyn(smem + 64);
    emu_dyn_smem = reinterpret_cast<uint8_t*>(((uintptr_t)dyn.data() + 15) & ~(uintptr_t)15);
    const uint32_t nw = (block + 31) / 32;
    std::vector<EmuWarp> warps(nw);
    emu_warps = warps.data();
    std::vector<EmuLaunch<Args>> ls(block);
    std::vector<pthread_t> th(block);
    pthread_attr_t at; pthread_attr_init(&at); pthread_attr_setstacksize(&at, 256 * 1024);
    for (uint32_t b = 0; b < grid; b++) {
        pthread_barrier_init(&emu_block_bar, nullptr, block);
        for (uint32_t w = 0; w < nw; w++) pthread_barrier_init(&warps[w].bar, nullptr, std::min(32u, block - 32 * w));
        for (uint32_t t = 0; t < block; t++) { ls[t] = EmuLaunch<Args>{k, a, t, b}; pthread_create(&th[t], &at, emu_thread<Args>, &ls[t]); }
        for (uint32_t t = 0; t < block; t++) pthread_join(th[t], nullptr);
        pthread_barrier_destroy(&emu_block_bar);
        for (uint32_t w = 0; w < nw; w++) pthread_barrier_destroy(&warps[w].bar);
    }
    pthread_attr_destroy(&at);
}
