// TEST INFRASTRUCTURE ONLY.  A functional emulation of the CUDA execution model on host threads, just wide enough for
// the plain kernels of csrc/fmbn.cu and csrc/sgns.cu: blocks run one after the other, every thread of a block is an
// OS thread, __syncthreads is a block barrier, warp shuffles exchange through a per-warp buffer between two warp
// barriers, atomics are host atomics, a thread that returns drops out of its barriers.  tests/emu/build_emu.py rewrites
// `kernel<<<grid, block, smem, stream>>>(args);` into emu::launch(...) and compiles the unit with g++ -std=c++20.
// It checks indexing and arithmetic of kernels that could not be run on a GPU yet; it says nothing about performance
// or about memory-model subtleties of the real hardware.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
typedef void *cudaGraphExec_t;
#define cudaSuccess 0
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n); return *p ? cudaSuccess : 1; }
inline cudaError_t cudaMallocAsync(void **p, size_t n, cudaStream_t) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : 1; }
inline cudaError_t cudaFreeAsync(void *p, cudaStream_t) { free(p); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : 1; }
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, int) { memcpy(d, s, n); return cudaSuccess; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
typedef void *cudaMemPool_t;
enum { cudaMemPoolAttrReleaseThreshold = 0 };
enum { cudaMemAllocationTypePinned = 1, cudaMemHandleTypeNone = 0, cudaMemLocationTypeDevice = 1 };
struct cudaMemPoolProps { int allocType, handleTypes; struct { int type, id; } location; };
inline cudaError_t cudaMemPoolCreate(cudaMemPool_t *, const cudaMemPoolProps *) { return 1; }  // "no private pool": plain cudaMallocAsync
inline cudaError_t cudaMemPoolDestroy(cudaMemPool_t) { return cudaSuccess; }
inline cudaError_t cudaMallocFromPoolAsync(void **p, size_t n, cudaMemPool_t, cudaStream_t) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : 1; }
inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t *, int) { return 1; }  // "no pool": the caller skips its tuning
inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, int, void *) { return cudaSuccess; }

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

namespace emu {
struct Block {
    std::unique_ptr<std::barrier<>> bar;
    std::vector<std::unique_ptr<std::barrier<>>> wbar;
    std::vector<uint64_t> xbuf;  // [warps][32]
    std::vector<uint64_t> dyn;   // dynamic shared memory of the launch (8-byte aligned)
};
inline Block *&cur() { static Block *b = nullptr; return b; }
inline thread_local int t_lin = 0;
inline void *dyn_smem() { return cur()->dyn.data(); }  // `extern __shared__ T name[];` is rewritten to use this
}  // namespace emu
inline thread_local dim3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

inline long long clock64() { return 0; }   // timing instrumentation compiles; the emulation measures nothing
inline void __syncthreads() { emu::cur()->bar->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::cur()->wbar[emu::t_lin >> 5]->arrive_and_wait(); }

template <class T>
inline T emu_exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    emu::Block *b = emu::cur();
    const int w = emu::t_lin >> 5, lane = emu::t_lin & 31;
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    b->xbuf[(size_t)w * 32 + lane] = bits;
    b->wbar[w]->arrive_and_wait();
    const uint64_t got = b->xbuf[(size_t)w * 32 + (src_lane & 31)];
    b->wbar[w]->arrive_and_wait();
    T r;
    memcpy(&r, &got, sizeof(T));
    return r;
}
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_exchange(v, (emu::t_lin & 31) ^ m); }

inline int atomicOr(int *p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline int atomicMin(int *p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
template <class T> inline T min(T a, T b) { return a < b ? a : b; }

// ctx.cuh keeps its device helpers behind __CUDACC__ (some are inline PTX): plain equivalents of the ones the emulated
// units use (cache hints have no meaning here)
struct float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float4 ld_row(const float *base, size_t i) { return reinterpret_cast<const float4 *>(base)[i]; }
inline void st_row(float *base, size_t i, float4 v) { reinterpret_cast<float4 *>(base)[i] = v; }
inline float4 ld_stream(const float *base, size_t i) { return ld_row(base, i); }
inline void st_stream(float *base, size_t i, float4 v) { st_row(base, i, v); }
inline float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
inline float4 f4_sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline float4 f4_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline float f4_dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= (unsigned)(emu_exchange(pred ? 1 : 0, l) != 0) << l;
    return m;
}
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T max(T a, T b) { return a < b ? b : a; }
// the two warp reductions the plain kernels use:
inline float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
inline double warp_sum_d(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

namespace emu {
struct Cfg {
    dim3 grid, block;
    size_t smem;
    Cfg(dim3 g, dim3 b, size_t sm = 0, cudaStream_t = nullptr) : grid(g), block(b), smem(sm) {}
};
inline void launch(const Cfg &c, const std::function<void()> &body) {
    const int T = (int)(c.block.x * c.block.y * c.block.z);
    if (T % 32) { fprintf(stderr, "emu: block of %d threads is not a multiple of 32\n", T); abort(); }
    blockDim = c.block;
    gridDim = c.grid;
    for (unsigned bz = 0; bz < c.grid.z; ++bz)
        for (unsigned by = 0; by < c.grid.y; ++by)
            for (unsigned bx = 0; bx < c.grid.x; ++bx) {
                Block blk;
                blk.bar = std::make_unique<std::barrier<>>(T);
                for (int w = 0; w < T / 32; ++w) blk.wbar.push_back(std::make_unique<std::barrier<>>(32));
                blk.xbuf.assign((size_t)T, 0);
                blk.dyn.assign(c.smem / 8 + 1, 0);
                cur() = &blk;
                std::vector<std::thread> th;
                th.reserve(T);
                for (int t = 0; t < T; ++t)
                    th.emplace_back([&, t] {
                        t_lin = t;
                        threadIdx = dim3(t % c.block.x, (t / c.block.x) % c.block.y, t / (c.block.x * c.block.y));
                        blockIdx = dim3(bx, by, bz);
                        body();
                        blk.wbar[t >> 5]->arrive_and_drop();  // a finished thread no longer takes part in barriers
                        blk.bar->arrive_and_drop();
                    });
                for (auto &x : th) x.join();
                cur() = nullptr;
            }
}
}  // namespace emu
