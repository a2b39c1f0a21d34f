// cuda_runtime.h — TEST INFRASTRUCTURE ONLY (tests/emu): a SIMT emulation of the subset of CUDA that
// fugu_b200/csrc/fg_kernels.cu and fg_api.cu use, so that the SAME kernel source can be compiled by
// g++ into tests/emu/libfugu_emu.so and its logic checked against the oracle on a machine without a
// GPU. It is never built into, loaded by or shipped with the product (fugu_b200/libfugu_gpu.so has no
// CPU path: fg_ctx_create fails with FG_ERR_NO_DEVICE); nothing measured or reported comes from it.
//
// Execution model: a launch runs the grid's CTAs one after the other on the calling thread. The
// threads of a CTA are fibers (one stack each, hand-written x86-64 context switch); a fiber runs
// until it reaches a warp collective or a CTA barrier and the scheduler resumes the next one, i.e.
// warps are NOT executed in lock step — any interleaving the CUDA memory model allows between
// synchronisation points is a legal execution, and this is one of them. Collectives require the full
// mask (the kernels use nothing else). A barrier or collective that can never complete aborts with a
// diagnostic instead of hanging. "Device memory" is host memory; streams and events are ordering
// no-ops because every operation completes before its call returns.
#pragma once
#define FG_EMULATE 1

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

// every standard header the sources use comes in BEFORE the qualifier macros below: libstdc++ spells
// attributes as __attribute__((__noinline__)), which a `#define __noinline__ ...` would break
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <unordered_map>
#include <vector>

// ---- qualifiers ---------------------------------------------------------------------------------
#define __device__
#define __host__
#define __global__ static
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static

// ---- vector types -------------------------------------------------------------------------------
struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct __attribute__((aligned(8))) uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(8))) float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint3 make_uint3(unsigned x, unsigned y, unsigned z) { return uint3{x, y, z}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

// ---- the emulator -------------------------------------------------------------------------------
namespace fgemu {
struct Fiber {
    void* sp;
    uint3 tid;
    int lane, warp;
    bool done;
};
struct Warp {
    uint64_t vals[32];
    uint32_t arrived, drained;
    int draining;
};
struct Cta {
    uint3 bid, bdim, gdim;
    unsigned live, bar_arrived, bar_gen;
    int orv[2];
    unsigned char* dyn;
    Warp warps[32];
};
extern Fiber* g_cur;
extern Cta g_cta;
extern unsigned long long g_progress;
void yield();
extern const char* g_kernel_name;
void launch(unsigned grid, unsigned block, size_t smem, const std::function<void()>& body);
void die(const char* what);
inline unsigned char* dyn_smem() { return g_cta.dyn; }

// every lane deposits one 64-bit value; returns once all 32 have; out[] = the 32 values
inline void warp_gather(unsigned mask, uint64_t v, uint64_t out[32]) {
    if (mask != 0xFFFFFFFFu) die("warp collective with a partial mask (not emulated)");
    Warp& W = g_cta.warps[g_cur->warp];
    const int l = g_cur->lane;
    while (W.draining) yield();  // the previous collective of this warp is still being read
    W.vals[l] = v;
    W.arrived |= 1u << l;
    g_progress++;
    if (W.arrived == 0xFFFFFFFFu) W.draining = 1;
    else while (!W.draining) yield();
    for (int i = 0; i < 32; i++) out[i] = W.vals[i];
    W.drained |= 1u << l;
    g_progress++;
    if (W.drained == 0xFFFFFFFFu) { W.arrived = 0; W.drained = 0; W.draining = 0; }
}
template <class T> inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

inline int cta_barrier(int pred) {
    Cta& C = g_cta;
    const unsigned g = C.bar_gen;
    const int slot = g & 1;
    if (pred) C.orv[slot] = 1;
    g_progress++;
    if (++C.bar_arrived >= C.live) {
        C.bar_arrived = 0;
        C.orv[(g + 1) & 1] = 0;
        C.bar_gen = g + 1;
    } else {
        while (C.bar_gen == g) yield();
        g_progress++;
    }
    return C.orv[slot];
}
}  // namespace fgemu

#define threadIdx (fgemu::g_cur->tid)
#define blockIdx (fgemu::g_cta.bid)
#define blockDim (fgemu::g_cta.bdim)
#define gridDim (fgemu::g_cta.gdim)

static inline void __syncthreads() { fgemu::cta_barrier(0); }
static inline int __syncthreads_or(int p) { return fgemu::cta_barrier(p != 0); }
static inline void __syncwarp(unsigned mask = 0xFFFFFFFFu) { uint64_t o[32]; fgemu::warp_gather(mask, 0, o); }

template <class T> static inline T __shfl_sync(unsigned m, T v, int src) {
    uint64_t o[32]; fgemu::warp_gather(m, fgemu::to_bits(v), o); return fgemu::from_bits<T>(o[src & 31]);
}
template <class T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d) {
    uint64_t o[32]; fgemu::warp_gather(m, fgemu::to_bits(v), o);
    const int l = fgemu::g_cur->lane; return (int)d <= l ? fgemu::from_bits<T>(o[l - (int)d]) : v;
}
template <class T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d) {
    uint64_t o[32]; fgemu::warp_gather(m, fgemu::to_bits(v), o);
    const int l = fgemu::g_cur->lane; return l + (int)d < 32 ? fgemu::from_bits<T>(o[l + (int)d]) : v;
}
template <class T> static inline T __shfl_xor_sync(unsigned m, T v, int x) {
    uint64_t o[32]; fgemu::warp_gather(m, fgemu::to_bits(v), o); return fgemu::from_bits<T>(o[(fgemu::g_cur->lane ^ x) & 31]);
}
static inline unsigned __ballot_sync(unsigned m, int p) {
    uint64_t o[32]; fgemu::warp_gather(m, p ? 1u : 0u, o);
    unsigned r = 0; for (int i = 0; i < 32; i++) r |= (unsigned)(o[i] & 1u) << i; return r;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, p) == 0xFFFFFFFFu; }

// ---- intrinsics -----------------------------------------------------------------------------------
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { return (unsigned)((((uint64_t)hi << 32) | lo) >> (s & 31)); }
static inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned s) { return (unsigned)((((uint64_t)hi << 32) | lo) >> (s > 32 ? 32 : s)); }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { return (unsigned)(((((uint64_t)hi << 32) | lo) << (s & 31)) >> 32); }
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    const uint64_t src = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned sel = (s >> (4 * i)) & 0xF;
        unsigned b = (unsigned)(src >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) b = (b & 0x80) ? 0xFF : 0x00;
        r |= b << (8 * i);
    }
    return r;
}
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
static inline long long clock64() { return 0; }
using std::max;
using std::min;

template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <class T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }

// ---- runtime API ------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
typedef struct fgemu_stream* cudaStream_t;
typedef struct fgemu_event { double ms; }* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int major, minor, multiProcessorCount; char name[64]; };

static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : e == cudaErrorMemoryAllocation ? "out of memory" : "invalid value"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 8; return cudaSuccess; }  /* one per rank of a multi-process test */
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof(*p)); p->major = 10; p->minor = 0; p->multiProcessorCount = 148; strcpy(p->name, "SIMT emulator (tests/emu)");
    return cudaSuccess;
}
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) & ~(size_t)255); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t = nullptr) {
    for (size_t r = 0; r < height; r++) memcpy((char*)d + r * dpitch, (const char*)s + r * spitch, width);
    return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)calloc(1, sizeof(fgemu_event)); return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { return cudaEventCreateWithFlags(e, 0); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) {
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); e->ms = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; return cudaSuccess;
}
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->ms - a->ms); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
