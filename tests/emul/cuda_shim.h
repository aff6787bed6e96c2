// cuda_shim.h -- the handful of CUDA runtime calls liinit_gpu.cu makes, on host memory, for the CPU checker (tests/emul).
// "Device" pointers are host pointers, streams and events do nothing (kernels run synchronously through simt_shim.h).
// TEST INFRASTRUCTURE: lets the C-ABI layer (argument checks, call order, staging, lazy initialisation, launch sequences)
// be compiled from its own source and exercised without a GPU. Never part of the product.
#pragma once
#include <cstdlib>
#include <cstring>

typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaHostAllocMapped = 2 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes { cudaMemoryType type; void* devicePointer; void* hostPointer; };
struct cudaDeviceProp { int multiProcessorCount; };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->multiProcessorCount = 2; return cudaSuccess; }   // small fixed grids
// cudaMalloc does not zero device memory. LIINIT_EMUL_POISON=1: every allocation is filled with 0xCD, so that code which leans on fresh
// pages being zero shows up as a failing test here instead of as a rare failure on a GPU whose memory was used before.
inline bool li_emul_poison() {
    static const bool on = [] { const char* e = getenv("LIINIT_EMUL_POISON"); return e && e[0] == '1'; }();
    return on;
}
template <class T>
inline cudaError_t cudaMalloc(T** p, size_t n) {
    *p = (T*)malloc(n ? n : 1);
    if (*p && li_emul_poison()) memset((void*)*p, 0xCD, n ? n : 1);
    return *p ? cudaSuccess : 2;
}
template <class T>
inline cudaError_t cudaMallocHost(T** p, size_t n) {
    *p = (T*)calloc(n ? n : 1, 1);
    if (*p && li_emul_poison()) memset((void*)*p, 0xCD, n ? n : 1);
    return *p ? cudaSuccess : 2;
}
template <class T>
inline cudaError_t cudaHostAlloc(T** p, size_t n, unsigned) {
    *p = (T*)calloc(n ? n : 1, 1);
    if (*p && li_emul_poison()) memset((void*)*p, 0xCD, n ? n : 1);
    return *p ? cudaSuccess : 2;
}
inline cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) { *d = h; return cudaSuccess; }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)1; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (void*)1; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
    a->type = cudaMemoryTypeHost;   // every buffer is "page-locked and mapped" here
    a->devicePointer = const_cast<void*>(p);
    a->hostPointer = const_cast<void*>(p);
    return cudaSuccess;
}
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
