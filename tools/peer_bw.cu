// Micro-benchmark (2 GPUs, one process): random 512-byte row gather FROM / scatter TO a peer GPU over NVLink.
// Decides how the row-sharded step moves item rows (pull by peer loads, TMA bulk pull, or push by peer stores).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peer_bw tools/peer_bw.cu && ./peer_bw
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>
#include <random>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ float4 ld_na(const float4 *p) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// pull: dst[c] = src[idx[c]]  (src may be a peer pointer), R rows in flight per warp
template <int R, bool NA>
__global__ void __launch_bounds__(256) k_pull(const float4 *__restrict__ src, const uint32_t *__restrict__ idx, int n, float4 *__restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t c0 = warp * R; c0 < (uint32_t)n; c0 += nw * R) {
        float4 r[R];
#pragma unroll
        for (int j = 0; j < R; ++j)
            if (c0 + j < n) { const float4 *p = src + (size_t)idx[c0 + j] * 32 + lane; r[j] = NA ? ld_na(p) : *p; }
#pragma unroll
        for (int j = 0; j < R; ++j)
            if (c0 + j < n) dst[(size_t)(c0 + j) * 32 + lane] = r[j];
    }
}

// push: dst[idx[c]] = src[c]  (dst may be a peer pointer)
template <int R>
__global__ void __launch_bounds__(256) k_push(const float4 *__restrict__ src, const uint32_t *__restrict__ idx, int n, float4 *__restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t c0 = warp * R; c0 < (uint32_t)n; c0 += nw * R) {
        float4 r[R];
#pragma unroll
        for (int j = 0; j < R; ++j)
            if (c0 + j < n) r[j] = ld_na(src + (size_t)(c0 + j) * 32 + lane);
#pragma unroll
        for (int j = 0; j < R; ++j)
            if (c0 + j < n) dst[(size_t)idx[c0 + j] * 32 + lane] = r[j];
    }
}

// serve: dst[c] = src[idx[c]] with LOCAL random reads and sequential PEER writes (owner pushes requested rows)
template <int R>
__global__ void __launch_bounds__(256) k_serve(const float4 *__restrict__ src, const uint32_t *__restrict__ idx, int n, float4 *__restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t c0 = warp * R; c0 < (uint32_t)n; c0 += nw * R) {
        float4 r[R];
#pragma unroll
        for (int j = 0; j < R; ++j)
            if (c0 + j < n) r[j] = src[(size_t)idx[c0 + j] * 32 + lane];
#pragma unroll
        for (int j = 0; j < R; ++j)
            if (c0 + j < n) dst[(size_t)(c0 + j) * 32 + lane] = r[j];
    }
}

// TMA bulk pull: each warp's lane 0 issues 512-byte cp.async.bulk copies global(peer) -> shared, NS rows per stage,
// waits on an mbarrier, then bulk-stores shared -> local global.
template <int NS>
__global__ void __launch_bounds__(128) k_pull_tma(const float4 *__restrict__ src, const uint32_t *__restrict__ idx, int n, float4 *__restrict__ dst) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *buf = smem + (size_t)wid * NS * 512;
    __shared__ __align__(8) unsigned long long bars[4];
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bars[wid]);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncwarp();
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    uint32_t phase = 0;
    for (uint32_t c0 = warp * NS; c0 < (uint32_t)n; c0 += nw * NS) {
        const int cnt = min(NS, n - (int)c0);
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(cnt * 512));
        __syncwarp();
        for (int j = lane; j < cnt; j += 32) {
            const void *g = src + (size_t)idx[c0 + j] * 32;
            const uint32_t s = (uint32_t)__cvta_generic_to_shared(buf + (size_t)j * 512);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s), "l"(g), "r"(512), "r"(bar) : "memory");
        }
        // wait
        uint32_t done = 0;
        while (!done) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(phase) : "memory");
        }
        phase ^= 1;
        // rows c0 .. c0+cnt are contiguous in dst: one bulk store of cnt*512 bytes
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;");
            const uint32_t s = (uint32_t)__cvta_generic_to_shared(buf);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (size_t)c0 * 32), "r"(s), "r"(cnt * 512) : "memory");
            asm volatile("cp.async.bulk.commit_group;");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncwarp();
    }
}

int main() {
    int nd = 0;
    CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
    const size_t ROWS = 10u << 20;   // 10M rows x 512 B = 5.1 GB table on each GPU
    const int N = 700000;            // rows moved per launch (~ the remote rows of one step at G=2)
    float4 *tab[2], *buf[2];
    uint32_t *idx[2];
    std::mt19937 rng(1);
    std::vector<uint32_t> h(N);
    for (auto &x : h) x = rng() % ROWS;
    std::sort(h.begin(), h.end());
    for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d));
        CK(cudaDeviceEnablePeerAccess(1 - d, 0));
        CK(cudaMalloc(&tab[d], ROWS * 512));
        CK(cudaMalloc(&buf[d], (size_t)N * 512));
        CK(cudaMalloc(&idx[d], N * 4));
        CK(cudaMemset(tab[d], 1, ROWS * 512));
        CK(cudaMemset(buf[d], 2, (size_t)N * 512));
        CK(cudaMemcpy(idx[d], h.data(), N * 4, cudaMemcpyHostToDevice));
    }
    CK(cudaSetDevice(0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    auto time = [&](const char *name, auto launch) {
        for (int i = 0; i < 2; ++i) launch();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        const int it = 5;
        for (int i = 0; i < it; ++i) launch();
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= it;
        printf("%-44s %8.3f ms  %8.1f GB/s\n", name, ms, (double)N * 512 / ms * 1e-6);
    };
    const int G = 148;
    time("local gather (baseline)            R=4", [&] { k_pull<4, true><<<G * 4, 256>>>(tab[0], idx[0], N, buf[0]); });
    time("peer pull  ld.na                   R=4", [&] { k_pull<4, true><<<G * 4, 256>>>(tab[1], idx[0], N, buf[0]); });
    time("peer pull  ld.na                   R=8", [&] { k_pull<8, true><<<G * 4, 256>>>(tab[1], idx[0], N, buf[0]); });
    time("peer pull  ld.na  R=8, 8 blocks/SM    ", [&] { k_pull<8, true><<<G * 8, 256>>>(tab[1], idx[0], N, buf[0]); });
    time("peer pull  ld (cached)             R=4", [&] { k_pull<4, false><<<G * 4, 256>>>(tab[1], idx[0], N, buf[0]); });
    time("peer pull  ld.na  R=4, 1 block/SM     ", [&] { k_pull<4, true><<<G, 256>>>(tab[1], idx[0], N, buf[0]); });
    for (int d = 1; d >= 0; --d) { CK(cudaSetDevice(d)); CK(cudaFuncSetAttribute(k_pull_tma<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32 * 512)); }
    time("peer pull  TMA bulk 512B, 32 rows/warp", [&] { k_pull_tma<32><<<G * 2, 128, 4 * 32 * 512>>>(tab[1], idx[0], N, buf[0]); });
    CK(cudaFuncSetAttribute(k_pull_tma<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16 * 512));
    time("peer pull  TMA bulk 512B, 16 rows/warp", [&] { k_pull_tma<16><<<G * 4, 128, 4 * 16 * 512>>>(tab[1], idx[0], N, buf[0]); });
    time("local TMA bulk gather  32 rows/warp   ", [&] { k_pull_tma<32><<<G * 2, 128, 4 * 32 * 512>>>(tab[0], idx[0], N, buf[0]); });
    time("peer push  random rows (scatter)   R=4", [&] { k_push<4><<<G * 4, 256>>>(buf[0], idx[0], N, tab[1]); });
    time("peer serve local gather->peer seq  R=4", [&] { k_serve<4><<<G * 4, 256>>>(tab[0], idx[0], N, buf[1]); });
    time("peer serve local gather->peer seq  R=8", [&] { k_serve<8><<<G * 8, 256>>>(tab[0], idx[0], N, buf[1]); });
    time("cudaMemcpyPeerAsync 358 MB            ", [&] { cudaMemcpyPeerAsync(buf[0], 0, buf[1], 1, (size_t)N * 512, 0); });
    // both directions at once: each GPU pulls from / serves to the other simultaneously
    {
        cudaEvent_t a0, a1, b0, b1;
        CK(cudaSetDevice(1)); CK(cudaEventCreate(&b0)); CK(cudaEventCreate(&b1));
        CK(cudaSetDevice(0)); CK(cudaEventCreate(&a0)); CK(cudaEventCreate(&a1));
        auto both = [&](const char *name, auto launch) {
            for (int w = 0; w < 2; ++w) { for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); launch(d); } }
            for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
            const int it = 5;
            CK(cudaSetDevice(0)); CK(cudaEventRecord(a0));
            CK(cudaSetDevice(1)); CK(cudaEventRecord(b0));
            for (int i = 0; i < it; ++i) for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); launch(d); }
            CK(cudaSetDevice(0)); CK(cudaEventRecord(a1));
            CK(cudaSetDevice(1)); CK(cudaEventRecord(b1));
            for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
            float ma, mb;
            CK(cudaEventElapsedTime(&ma, a0, a1)); CK(cudaEventElapsedTime(&mb, b0, b1));
            printf("%-44s gpu0 %7.3f ms %7.1f GB/s | gpu1 %7.3f ms %7.1f GB/s\n", name, ma / it, (double)N * 512 / (ma / it) * 1e-6,
                   mb / it, (double)N * 512 / (mb / it) * 1e-6);
        };
        both("BIDIR peer pull ld.na R=4", [&](int d) { k_pull<4, true><<<G * 4, 256>>>(tab[1 - d], idx[d], N, buf[d]); });
        both("BIDIR peer pull TMA 32 rows/warp", [&](int d) { k_pull_tma<32><<<G * 2, 128, 4 * 32 * 512>>>(tab[1 - d], idx[d], N, buf[d]); });
        both("BIDIR peer serve (local gather, peer st)", [&](int d) { k_serve<4><<<G * 4, 256>>>(tab[d], idx[d], N, buf[1 - d]); });
        both("BIDIR peer push random rows", [&](int d) { k_push<4><<<G * 4, 256>>>(buf[d], idx[d], N, tab[1 - d]); });
        both("BIDIR memcpy peer", [&](int d) { cudaMemcpyPeerAsync(buf[d], d, buf[1 - d], 1 - d, (size_t)N * 512, 0); });
        CK(cudaSetDevice(0));
    }
    // verify the TMA path moved the right bytes
    CK(cudaMemset(buf[0], 0, (size_t)N * 512));
    k_pull_tma<32><<<G * 2, 128, 4 * 32 * 512>>>(tab[1], idx[0], N, buf[0]);
    CK(cudaDeviceSynchronize());
    std::vector<unsigned char> chk(4096);
    CK(cudaMemcpy(chk.data(), (char *)buf[0] + (size_t)(N - 8) * 512, 4096, cudaMemcpyDeviceToHost));
    bool ok = true;
    for (auto b : chk) ok = ok && b == 1;
    printf("TMA pull verify: %s\n", ok ? "ok" : "MISMATCH");
    return 0;
}
