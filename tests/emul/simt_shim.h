// simt_shim.h -- runs the warp-cooperative CUDA kernels of the product ON THE CPU, unchanged: one warp = 32 fibers
// (ucontext) scheduled round-robin, every warp intrinsic is a rendezvous. TEST INFRASTRUCTURE (tests/emul), never part
// of the product: it exists so that the lockstep search kernel (knn_kernels.cuh) -- whose control flow lives in votes
// and shuffles and therefore cannot be restated lane by lane -- is checked by `-m "not gpu"` tests too.
//
// Model: all kernels here call every *_sync intrinsic with the full mask from warp-uniform control flow, so every live
// lane executes the same sequence of intrinsics. Rendezvous k: a lane deposits its operand in buffer k & 1 and yields to
// the scheduler; the scheduler resumes the lanes in turn, so when a lane runs again all 32 deposits of rendezvous k are
// there. Two buffers suffice: rendezvous k + 2 cannot start before every lane has read k. Code between two rendezvous
// simply runs lane after lane. Arithmetic: the host is compiled with -ffp-contract=off, the library with --fmad=false,
// both IEEE, so *_rn intrinsics are the plain operators.
#pragma once
#define LI_SIMT_EMUL 1
#include <ucontext.h>
#include <vector_functions.h>
#include <vector_types.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

using std::isfinite;

#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif

struct SimtDim { unsigned x, y, z; };
inline SimtDim threadIdx{0, 0, 0}, blockIdx{0, 0, 0}, blockDim{32, 1, 1}, gridDim{1, 1, 1};

struct SimtWarp {
    ucontext_t sched;
    ucontext_t lane[32];
    char* stack[32];
    bool fin[32];
    uint64_t buf[2][32];
    unsigned long nsync[32];
    int cur;
    unsigned long long rendezvous;   // statistics
};
inline SimtWarp g_w;

inline uint64_t simt_exchange_begin(uint64_t v) {   // deposit + yield; returns the buffer index to read from
    const int l = g_w.cur;
    const unsigned long k = g_w.nsync[l]++;
    g_w.buf[k & 1][l] = v;
    g_w.rendezvous++;
    swapcontext(&g_w.lane[l], &g_w.sched);
    return k & 1;
}
template <class T>
inline uint64_t simt_bits(T v) {
    static_assert(sizeof(T) <= 8, "shuffle operand too wide");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <class T>
inline T simt_unbits(uint64_t b) {
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}

template <class T>
inline T __shfl_sync(unsigned, T v, int src) {
    const uint64_t k = simt_exchange_begin(simt_bits(v));
    return simt_unbits<T>(g_w.buf[k][src & 31]);
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int o) {
    const int l = g_w.cur;
    const uint64_t k = simt_exchange_begin(simt_bits(v));
    return simt_unbits<T>(g_w.buf[k][(l ^ o) & 31]);
}
template <class T>
inline T __shfl_up_sync(unsigned, T v, int o) {
    const int l = g_w.cur;
    const uint64_t k = simt_exchange_begin(simt_bits(v));
    return l >= o ? simt_unbits<T>(g_w.buf[k][l - o]) : v;
}
inline unsigned __ballot_sync(unsigned, bool p) {
    const uint64_t k = simt_exchange_begin(p ? 1u : 0u);
    unsigned r = 0;
    for (int i = 0; i < 32; i++)
        if (!g_w.fin[i] && g_w.buf[k][i]) r |= 1u << i;
    return r;
}
inline int __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0u; }
inline int __all_sync(unsigned m, bool p) { return __ballot_sync(m, !p) == 0u; }
inline unsigned __reduce_min_sync(unsigned, unsigned v) {
    const uint64_t k = simt_exchange_begin(v);
    unsigned r = 0xffffffffu;
    for (int i = 0; i < 32; i++)
        if (!g_w.fin[i] && (unsigned)g_w.buf[k][i] < r) r = (unsigned)g_w.buf[k][i];
    return r;
}
inline void __syncwarp(unsigned = 0xffffffffu) { (void)simt_exchange_begin(0); }

// ---- scalar device intrinsics -----------------------------------------------------------------------
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline unsigned __float_as_uint(float f) { return simt_unbits<unsigned>(simt_bits(f)); }
inline float __uint_as_float(unsigned u) { return simt_unbits<float>(simt_bits(u)); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
template <class T>
inline T __ldg(const T* p) { return *p; }
template <class T>
inline T __ldcg(const T* p) { return *p; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
// one OS thread: atomics are plain read-modify-writes
template <class T>
inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T>
inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T>
inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T>
inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }

// ---- running one warp ---------------------------------------------------------------------------------
typedef void (*SimtLaneFn)(void*);
struct SimtLaunch { SimtLaneFn fn; void* arg; };
inline SimtLaunch g_simt_launch;
inline void simt_trampoline() {
    g_simt_launch.fn(g_simt_launch.arg);
    g_w.fin[g_w.cur] = true;   // uc_link brings us back to the scheduler
}
// Runs fn(arg) once per lane of ONE warp (threadIdx.x = 0..31, blockDim.x = 32, the given block of a grid of `nblocks`).
inline void simt_run_warp(SimtLaneFn fn, void* arg, unsigned block, unsigned nblocks) {
    const size_t STACK = 256 * 1024;
    g_simt_launch.fn = fn;
    g_simt_launch.arg = arg;
    blockDim = SimtDim{32, 1, 1};
    gridDim = SimtDim{nblocks, 1, 1};
    blockIdx = SimtDim{block, 0, 0};
    for (int l = 0; l < 32; l++) {
        if (!g_w.stack[l]) g_w.stack[l] = (char*)malloc(STACK);
        g_w.fin[l] = false;
        g_w.nsync[l] = 0;
        getcontext(&g_w.lane[l]);
        g_w.lane[l].uc_stack.ss_sp = g_w.stack[l];
        g_w.lane[l].uc_stack.ss_size = STACK;
        g_w.lane[l].uc_link = &g_w.sched;
        makecontext(&g_w.lane[l], simt_trampoline, 0);
    }
    for (bool alive = true; alive;) {
        alive = false;
        for (int l = 0; l < 32; l++) {
            if (g_w.fin[l]) continue;
            g_w.cur = l;
            threadIdx = SimtDim{(unsigned)l, 0, 0};
            swapcontext(&g_w.sched, &g_w.lane[l]);
            if (!g_w.fin[l]) alive = true;
        }
    }
    // lockstep check: every lane must have gone through the same number of rendezvous
    for (int l = 1; l < 32; l++)
        if (g_w.nsync[l] != g_w.nsync[0]) {
            fprintf(stderr, "simt_shim: lane %d made %lu rendezvous, lane 0 %lu -- control flow around a warp intrinsic is not warp-uniform\n", l,
                    g_w.nsync[l], g_w.nsync[0]);
            abort();
        }
}
