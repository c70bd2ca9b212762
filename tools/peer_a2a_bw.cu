// Micro-benchmark (one process, all GPUs of the box): what the NVLink / NVSwitch fabric gives the access pattern of the
// row-sharded step -- every GPU pulls random 512-byte rows from ALL its peers and pushes 512-byte rows into ALL its
// peers at the same time.  The 2-GPU figures (tools/peer_bw.cu: 625-700 GB/s per direction) do not say what 8 GPUs
// pulling from 7 peers each sustain; this does.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peer_a2a_bw tools/peer_a2a_bw.cu && ./tools/peer_a2a_bw [G]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <random>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ float4 ld_na(const float4 *p) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
// dst[c] = *src[c] (512-byte rows; src[c] may point into any peer), R rows in flight per warp
template <int R>
__global__ void __launch_bounds__(256) k_pull(const float4 *const *__restrict__ src, int n, float4 *__restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t c0 = warp * R; c0 < (uint32_t)n; c0 += nw * R) {
        float4 r[R];
#pragma unroll
        for (int j = 0; j < R; ++j) if (c0 + j < n) r[j] = ld_na(src[c0 + j] + lane);
#pragma unroll
        for (int j = 0; j < R; ++j) if (c0 + j < n) dst[(size_t)(c0 + j) * 32 + lane] = r[j];
    }
}
// *dstp[c] = local[c]
template <int R>
__global__ void __launch_bounds__(256) k_push(float4 *const *__restrict__ dstp, int n, const float4 *__restrict__ local) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t c0 = warp * R; c0 < (uint32_t)n; c0 += nw * R) {
        float4 r[R];
#pragma unroll
        for (int j = 0; j < R; ++j) if (c0 + j < n) r[j] = local[(size_t)(c0 + j) * 32 + lane];
#pragma unroll
        for (int j = 0; j < R; ++j) if (c0 + j < n) dstp[c0 + j][lane] = r[j];
    }
}
// the step's pattern: pull a row, push a row derived from it
template <int R>
__global__ void __launch_bounds__(256) k_pull_push(const float4 *const *__restrict__ src, float4 *const *__restrict__ dstp, int n) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t c0 = warp * R; c0 < (uint32_t)n; c0 += nw * R) {
        float4 r[R];
#pragma unroll
        for (int j = 0; j < R; ++j) if (c0 + j < n) r[j] = ld_na(src[c0 + j] + lane);
#pragma unroll
        for (int j = 0; j < R; ++j) if (c0 + j < n) { r[j].x += 1.f; dstp[c0 + j][lane] = r[j]; }
    }
}

int main(int argc, char **argv) {
    int G = 0;
    CK(cudaGetDeviceCount(&G));
    if (argc > 1) G = std::min(G, atoi(argv[1]));
    const size_t rows = 2500000;   // rows of 512 B per GPU (config 5 at 8 ranks: 20 M items / 8)
    const int n = 1150000;         // remote rows per GPU per step at 8 ranks
    printf("%d GPUs, %zu rows of 512 B per GPU, %d remote rows per GPU per launch (%.0f MB each way)\n", G, rows, n, n * 512.0 / 1e6);
    std::vector<float4 *> table(G), pulled(G), recv(G);
    std::vector<const float4 **> src(G);
    std::vector<float4 **> dstp(G);
    std::vector<const float4 **> src_i(G);
    std::vector<float4 **> dstp_i(G);
    std::vector<cudaStream_t> s1(G), s2(G);
    std::vector<cudaEvent_t> e0(G), e1(G);
    for (int d = 0; d < G; ++d) {
        CK(cudaSetDevice(d));
        for (int p = 0; p < G; ++p) if (p != d) { cudaError_t e = cudaDeviceEnablePeerAccess(p, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { printf("no peer access %d -> %d\n", d, p); return 1; } cudaGetLastError(); }
        CK(cudaMalloc(&table[d], rows * 512)); CK(cudaMemset(table[d], 0, rows * 512));
        CK(cudaMalloc(&pulled[d], (size_t)n * 512)); CK(cudaMalloc(&recv[d], (size_t)G * n * 512 / std::max(1, G - 1) + 512 * 1024));
        CK(cudaMalloc(&src[d], (size_t)n * 8)); CK(cudaMalloc(&dstp[d], (size_t)n * 8));
        CK(cudaStreamCreateWithFlags(&s1[d], cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2[d], cudaStreamNonBlocking));
        CK(cudaEventCreate(&e0[d])); CK(cudaEventCreate(&e1[d]));
    }
    std::mt19937_64 rng(7);
    for (int d = 0; d < G; ++d) {
        std::vector<std::pair<int, uint32_t>> refs(n);
        for (auto &r : refs) { int o = (int)(rng() % (G - 1)); if (o >= d) ++o; r = {o, (uint32_t)(rng() % rows)}; }
        std::sort(refs.begin(), refs.end());  // grouped by owner, ascending rows: the step's cache-row order
        std::vector<const float4 *> hs(n); std::vector<float4 *> hd(n);
        std::vector<size_t> cnt(G, 0);
        const size_t region = (size_t)n / std::max(1, G - 1) + 4096;  // sender d's region in every owner's recv buffer
        for (int c = 0; c < n; ++c) {
            const int o = refs[c].first;
            hs[c] = table[o] + (size_t)refs[c].second * 32;
            hd[c] = recv[o] + ((size_t)(d < o ? d : d - 1) * region + std::min(cnt[o]++, region - 1)) * 32;
        }
        CK(cudaSetDevice(d));
        CK(cudaMemcpy(src[d], hs.data(), (size_t)n * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dstp[d], hd.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
        // the same rows, dealt round-robin over G ranges of the owner-grouped list in groups of 32 rows (what
        // MainArgs::ilv / k_shard_fetch do): at any moment a GPU's accesses are spread over all owners
        const int grp = 32, ng = (n + grp - 1) / grp, per = (ng + G - 1) / G;
        std::vector<const float4 *> is; std::vector<float4 *> id;
        for (int raw = 0; raw < G * per; ++raw) {
            const int c = (raw % G) * per + raw / G;
            if (c >= ng) continue;
            for (int x = c * grp; x < std::min(n, (c + 1) * grp); ++x) { is.push_back(hs[x]); id.push_back(hd[x]); }
        }
        CK(cudaMalloc(&src_i[d], (size_t)n * 8)); CK(cudaMalloc(&dstp_i[d], (size_t)n * 8));
        CK(cudaMemcpy(src_i[d], is.data(), (size_t)n * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dstp_i[d], id.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
    }
    auto run = [&](const char *name, int mode, double bytes_each_way) {
        for (int it = 0; it < 4; ++it) {
            for (int d = 0; d < G; ++d) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
            for (int d = 0; d < G; ++d) {
                CK(cudaSetDevice(d));
                CK(cudaEventRecord(e0[d], s1[d]));
                const float4 **sp = mode >= 4 ? src_i[d] : src[d];
                float4 **dp = mode >= 4 ? dstp_i[d] : dstp[d];
                if (mode % 4 == 0) k_pull<4><<<148 * 8, 256, 0, s1[d]>>>(sp, n, pulled[d]);
                if (mode % 4 == 1) k_push<4><<<148 * 8, 256, 0, s1[d]>>>(dp, n, pulled[d]);
                if (mode % 4 == 2) k_pull_push<4><<<148 * 8, 256, 0, s1[d]>>>(sp, dp, n);
                if (mode % 4 == 3) k_pull_push<8><<<148 * 8, 256, 0, s1[d]>>>(sp, dp, n);
                CK(cudaEventRecord(e1[d], s1[d]));
            }
            if (it == 3) {
                float lo = 1e9f, hi = 0.f;
                for (int d = 0; d < G; ++d) { CK(cudaSetDevice(d)); CK(cudaEventSynchronize(e1[d])); float ms; CK(cudaEventElapsedTime(&ms, e0[d], e1[d])); lo = std::min(lo, ms); hi = std::max(hi, ms); }
                printf("%-44s %.3f - %.3f ms per GPU  => %.0f - %.0f GB/s per direction per GPU\n", name, lo, hi, bytes_each_way / hi / 1e6, bytes_each_way / lo / 1e6);
            }
        }
    };
    const double b = n * 512.0;
    run("all GPUs pull from all peers", 0, b);
    run("all GPUs push to all peers", 1, b);
    run("pull + push per row (4 rows in flight/warp)", 2, 2 * b);  // every GPU also serves the same volume: 2 x b each way
    run("pull + push per row (8 rows in flight/warp)", 3, 2 * b);
    printf("-- the same rows, interleaved over the owners --\n");
    run("all GPUs pull from all peers", 4, b);
    run("all GPUs push to all peers", 5, b);
    run("pull + push per row (4 rows in flight/warp)", 6, 2 * b);
    run("pull + push per row (8 rows in flight/warp)", 7, 2 * b);
    return 0;
}
