// simt_shim.h -- runs the CUDA kernels of the product ON THE CPU, unchanged: one thread = one fiber (ucontext), every warp
// intrinsic and every __syncthreads is a rendezvous. TEST INFRASTRUCTURE (tests/emul), never part
// of the product: it exists so that the lockstep search kernel (knn_kernels.cuh) -- whose control flow lives in votes
// and shuffles and therefore cannot be restated lane by lane -- is checked by `-m "not gpu"` tests too.
//
// Model: the kernels call every *_sync intrinsic with the full mask from warp-uniform control flow, so every live lane of
// a warp executes the same sequence of intrinsics. An intrinsic is a counting barrier over the live lanes of the warp with
// two exchange buffers (generation parity): a lane deposits its operand, the last one to arrive opens the generation, the
// others yield to the scheduler until it is open; nobody can start generation k + 2 before everybody has left k + 1, i.e.
// has read k. __syncthreads is the same barrier over the block. Blocks run one after the other, the threads of a block
// are resumed round-robin; code between two rendezvous simply runs thread after thread, atomics are plain
// read-modify-writes (one OS thread), __shared__ is a static. Threads that return from the kernel stop counting.
// Arithmetic: the host is compiled with -ffp-contract=off, the library with --fmad=false, both IEEE, so *_rn intrinsics
// are the plain operators.
#pragma once
#define LI_SIMT_EMUL 1
#include <ucontext.h>
#include <vector_types.h>   // the vector structs only (host_defines.h qualifiers come with it); no CUDA API header

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

using std::isfinite;

// the constructors of vector_functions.h that the kernels use
inline float4 make_float4(float x, float y, float z, float w) { float4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
inline int4 make_int4(int x, int y, int z, int w) { int4 v; v.x = x; v.y = y; v.z = z; v.w = w; return v; }
inline float2 make_float2(float x, float y) { float2 v; v.x = x; v.y = y; return v; }
inline uint2 make_uint2(unsigned x, unsigned y) { uint2 v; v.x = x; v.y = y; return v; }
inline ushort2 make_ushort2(unsigned short x, unsigned short y) { ushort2 v; v.x = x; v.y = y; return v; }
inline float3 make_float3(float x, float y, float z) { float3 v; v.x = x; v.y = y; v.z = z; return v; }
inline int3 make_int3(int x, int y, int z) { int3 v; v.x = x; v.y = y; v.z = z; return v; }

struct SimtDim { unsigned x, y, z; };
inline SimtDim threadIdx{0, 0, 0}, blockIdx{0, 0, 0}, blockDim{32, 1, 1}, gridDim{1, 1, 1};

#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#undef __shared__
#define __shared__ static   // one block runs at a time and all its fibers live on one OS thread

constexpr int SIMT_MAX_THREADS = 1024;
struct SimtBlock {
    ucontext_t sched;
    ucontext_t th[SIMT_MAX_THREADS];
    char* stack[SIMT_MAX_THREADS];
    bool fin[SIMT_MAX_THREADS];
    int nthreads;
    int cur;   // thread of the block that is running
    // warp rendezvous: a counting barrier per warp, two exchange buffers (generation parity)
    uint64_t buf[SIMT_MAX_THREADS / 32][2][32];
    int warr[SIMT_MAX_THREADS / 32];
    unsigned long wgen[SIMT_MAX_THREADS / 32];
    int wlive[SIMT_MAX_THREADS / 32];
    // block barrier
    int barr;
    unsigned long bgen;
    int blive;
    unsigned long long rendezvous;   // statistics
};
inline SimtBlock g_b;

inline void simt_yield() { swapcontext(&g_b.th[g_b.cur], &g_b.sched); }

// deposit v, wait until every live lane of this thread's warp has deposited; returns the buffer to read from
inline const uint64_t* simt_exchange(uint64_t v) {
    const int t = g_b.cur, w = t >> 5, l = t & 31;
    const unsigned long gen = g_b.wgen[w];
    g_b.buf[w][gen & 1][l] = v;
    g_b.rendezvous++;
    if (++g_b.warr[w] >= g_b.wlive[w]) {
        g_b.warr[w] = 0;
        g_b.wgen[w]++;
    } else {
        while (g_b.wgen[w] == gen) simt_yield();
    }
    return g_b.buf[w][gen & 1];
}
inline bool simt_lane_live(int l) { return !g_b.fin[(g_b.cur & ~31) + l]; }

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
    return simt_unbits<T>(simt_exchange(simt_bits(v))[src & 31]);
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int o) {
    const int l = g_b.cur & 31;
    return simt_unbits<T>(simt_exchange(simt_bits(v))[(l ^ o) & 31]);
}
template <class T>
inline T __shfl_up_sync(unsigned, T v, int o) {
    const int l = g_b.cur & 31;
    const uint64_t* b = simt_exchange(simt_bits(v));
    return l >= o ? simt_unbits<T>(b[l - o]) : v;
}
template <class T>
inline T __shfl_down_sync(unsigned, T v, int o) {
    const int l = g_b.cur & 31;
    const uint64_t* b = simt_exchange(simt_bits(v));
    return l + o < 32 ? simt_unbits<T>(b[l + o]) : v;
}
inline unsigned __ballot_sync(unsigned, bool p) {
    const uint64_t* b = simt_exchange(p ? 1u : 0u);
    unsigned r = 0;
    for (int i = 0; i < 32; i++)
        if (simt_lane_live(i) && b[i]) r |= 1u << i;
    return r;
}
inline int __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0u; }
inline int __all_sync(unsigned m, bool p) { return __ballot_sync(m, !p) == 0u; }
inline unsigned __reduce_min_sync(unsigned, unsigned v) {
    const uint64_t* b = simt_exchange(v);
    unsigned r = 0xffffffffu;
    for (int i = 0; i < 32; i++)
        if (simt_lane_live(i) && (unsigned)b[i] < r) r = (unsigned)b[i];
    return r;
}
inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    const uint64_t* b = simt_exchange(v);
    unsigned r = 0u;
    for (int i = 0; i < 32; i++)
        if (simt_lane_live(i) && (unsigned)b[i] > r) r = (unsigned)b[i];
    return r;
}
inline unsigned __reduce_add_sync(unsigned, unsigned v) {
    const uint64_t* b = simt_exchange(v);
    unsigned r = 0u;
    for (int i = 0; i < 32; i++)
        if (simt_lane_live(i)) r += (unsigned)b[i];
    return r;
}
inline unsigned __match_any_sync(unsigned, unsigned v) {
    const uint64_t* b = simt_exchange(v);
    unsigned r = 0;
    for (int i = 0; i < 32; i++)
        if (simt_lane_live(i) && (unsigned)b[i] == v) r |= 1u << i;
    return r;
}
inline unsigned __activemask() { unsigned r = 0; for (int i = 0; i < 32; i++) if (simt_lane_live(i)) r |= 1u << i; return r; }
inline void __syncwarp(unsigned = 0xffffffffu) { (void)simt_exchange(0); }
inline void __syncthreads() {
    const unsigned long gen = g_b.bgen;
    if (++g_b.barr >= g_b.blive) {
        g_b.barr = 0;
        g_b.bgen++;
    } else {
        while (g_b.bgen == gen) simt_yield();
    }
}
inline void __threadfence() {}
inline void __threadfence_system() {}

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
inline int __float_as_int(float f) { return simt_unbits<int>(simt_bits(f)); }
inline float __int_as_float(int i) { return simt_unbits<float>(simt_bits(i)); }
inline long long __double_as_longlong(double d) { return simt_unbits<long long>(simt_bits(d)); }
inline double __longlong_as_double(long long v) { return simt_unbits<double>(simt_bits(v)); }
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
template <class T, class U>
inline T atomicAdd(T* p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U>
inline T atomicSub(T* p, U v) { T o = *p; *p = (T)(o - (T)v); return o; }
template <class T, class U>
inline T atomicOr(T* p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <class T, class U>
inline T atomicAnd(T* p, U v) { T o = *p; *p = (T)(o & (T)v); return o; }
template <class T, class U>
inline T atomicExch(T* p, U v) { T o = *p; *p = (T)v; return o; }
template <class T, class U>
inline T atomicMin(T* p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class U>
inline T atomicMax(T* p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U, class V>
inline T atomicCAS(T* p, U c, V v) { T o = *p; if (o == (T)c) *p = (T)v; return o; }

// ---- running a grid: blocks one after the other, the threads of a block as fibers ------------------------------------
struct SimtLaunch { void (*fn)(void*); void* arg; };
inline SimtLaunch g_simt_launch;
inline void simt_trampoline() {
    g_simt_launch.fn(g_simt_launch.arg);
    // a thread that leaves the kernel no longer takes part in rendezvous / barriers (as on the device: exited threads do not block)
    const int t = g_b.cur, w = t >> 5;
    g_b.fin[t] = true;
    g_b.wlive[w]--;
    g_b.blive--;
    if (g_b.wlive[w] > 0 && g_b.warr[w] >= g_b.wlive[w]) { g_b.warr[w] = 0; g_b.wgen[w]++; }
    if (g_b.blive > 0 && g_b.barr >= g_b.blive) { g_b.barr = 0; g_b.bgen++; }
}
// kernel<<<grid, block>>>: fn(arg) is the kernel call with its arguments bound
inline void simt_launch(unsigned grid, unsigned block, void (*fn)(void*), void* arg) {
    const size_t STACK = 192 * 1024;
    if (block == 0 || block > SIMT_MAX_THREADS || (block & 31)) {
        fprintf(stderr, "simt_shim: block size %u not supported (multiple of 32, <= %d)\n", block, SIMT_MAX_THREADS);
        abort();
    }
    g_simt_launch.fn = fn;
    g_simt_launch.arg = arg;
    blockDim = SimtDim{block, 1, 1};
    gridDim = SimtDim{grid, 1, 1};
    g_b.nthreads = (int)block;
    // LIINIT_EMUL_SHUFFLE=<seed>: the blocks of every launch run in a pseudo-random order and the threads of every other block are
    // scheduled last-to-first -- the device promises no order either; a result that depends on it fails the oracle comparisons here
    static const unsigned shuffle_seed = [] { const char* e = getenv("LIINIT_EMUL_SHUFFLE"); return e ? (unsigned)atoi(e) * 2654435761u + 1u : 0u; }();
    static unsigned shuffle_state = shuffle_seed;
    static std::vector<unsigned> order;
    order.resize(grid);
    for (unsigned b = 0; b < grid; b++) order[b] = b;
    if (shuffle_seed)
        for (unsigned b = grid; b > 1; b--) {
            shuffle_state = shuffle_state * 1664525u + 1013904223u;
            const unsigned j = (shuffle_state >> 8) % b;
            const unsigned t = order[b - 1]; order[b - 1] = order[j]; order[j] = t;
        }
    for (unsigned bi = 0; bi < grid; bi++) {
        const unsigned b = order[bi];
        const bool reverse = shuffle_seed && (bi & 1u);
        blockIdx = SimtDim{b, 0, 0};
        g_b.barr = 0; g_b.bgen = 0; g_b.blive = (int)block;
        for (unsigned w = 0; w < block / 32; w++) { g_b.warr[w] = 0; g_b.wgen[w] = 0; g_b.wlive[w] = 32; }
        for (unsigned t = 0; t < block; t++) {
            if (!g_b.stack[t]) g_b.stack[t] = (char*)malloc(STACK);
            g_b.fin[t] = false;
            getcontext(&g_b.th[t]);
            g_b.th[t].uc_stack.ss_sp = g_b.stack[t];
            g_b.th[t].uc_stack.ss_size = STACK;
            g_b.th[t].uc_link = &g_b.sched;
            makecontext(&g_b.th[t], simt_trampoline, 0);
        }
        for (bool alive = true; alive;) {
            alive = false;
            for (unsigned tt = 0; tt < block; tt++) {
                const unsigned t = reverse ? block - 1 - tt : tt;
                if (g_b.fin[t]) continue;
                g_b.cur = (int)t;
                threadIdx = SimtDim{t, 0, 0};
                swapcontext(&g_b.sched, &g_b.th[t]);
                if (!g_b.fin[t]) alive = true;
            }
        }
    }
}
// bind a kernel call: SIMT_LAUNCH(grid, block, kernel<...>, args...)
template <class F>
inline void simt_launch_fn(unsigned grid, unsigned block, F&& f) {
    auto thunk = [](void* p) { (*static_cast<typename std::remove_reference<F>::type*>(p))(); };
    simt_launch(grid, block, thunk, (void*)&f);
}
