// Micro-benchmark: what one warp pays for gpu-scope ordering primitives on B200 (single warp alone on the GPU, and with
// 1000 other warps polling).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fence_cost tools/fence_cost.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ unsigned ld_acq(const unsigned *p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_rlx(const unsigned *p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_rel(unsigned *p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_rlx(unsigned *p, unsigned v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// mode: 0 stores only; 1 stores + fence.acq_rel.gpu; 2 stores + st.release (lane 0); 3 ld.acquire only; 4 ld.relaxed + dependent ld.cg;
//       5 stores + fence.sc.gpu; 6 stores + read-back (ld.cg same address) ; 7 stores + __threadfence_block
__global__ void k(int mode, int iters, double *rows, unsigned *ver, long long *out, int nrows) {
    const int lane = threadIdx.x & 31;
    if (blockIdx.x != 0 || threadIdx.x >= 32) {  // background pollers (if launched with more than one warp)
        unsigned spins = 0;
        while (ld_acq(&ver[1 + (blockIdx.x % 1000)]) != 0xFFFFFFFFu && spins < (1u << 30)) { __nanosleep(20); ++spins; if (*(volatile unsigned *)&ver[0] == 0xDEADu) break; }
        return;
    }
    double acc = 0.0;
    long long t0 = clock64();
    unsigned r = 1;
    for (int it = 0; it < iters; ++it) {
        r = r * 1664525u + 1013904223u;
        double *row = rows + (size_t)(r % nrows) * 128;
        if (mode == 0 || mode == 1 || mode == 2 || mode == 5 || mode == 6 || mode == 7) {
            for (int v = 0; v < 4; ++v) __stcg(row + lane + 32 * v, (double)it);
        }
        if (mode == 1) asm volatile("fence.acq_rel.gpu;" ::: "memory");
        if (mode == 5) asm volatile("fence.sc.gpu;" ::: "memory");
        if (mode == 7) __threadfence_block();
        if (mode == 2) { __syncwarp(); if (lane == 0) st_rel(&ver[2000 + (r % 1000)], it); }
        if (mode == 3) acc += ld_acq(&ver[2000 + (r % 1000)]);
        if (mode == 4) { unsigned v = ld_rlx(&ver[2000 + (r % 1000)]); acc += __ldcg(row + lane + (v & 0)); }
        if (mode == 6) { for (int v = 0; v < 4; ++v) acc += __ldcg(row + lane + 32 * v); }
    }
    long long t1 = clock64();
    if (lane == 0) { out[0] = t1 - t0; out[1] = (long long)acc; ver[0] = 0xDEADu; }
}

int main() {
    const int nrows = 50000, iters = 20000;
    double *rows; unsigned *ver; long long *out;
    CK(cudaMalloc(&rows, (size_t)nrows * 128 * 8)); CK(cudaMalloc(&ver, 4000 * 4)); CK(cudaMalloc(&out, 16));
    CK(cudaMemset(rows, 0, (size_t)nrows * 128 * 8));
    int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    const char *names[] = {"4 x st.cg (1 KB row) only", "stores + fence.acq_rel.gpu", "stores + syncwarp + st.release.gpu (lane 0)", "ld.acquire.gpu only",
                           "ld.relaxed + address-dependent ld.cg", "stores + fence.sc.gpu", "stores + read-back of the row (ld.cg)", "stores + fence.acq_rel.cta"};
    for (int bg = 0; bg < 2; ++bg)
        for (int mode = 0; mode < 8; ++mode) {
            CK(cudaMemset(ver, 0, 4000 * 4));
            k<<<bg ? 1000 : 1, bg ? 64 : 32>>>(mode, iters, rows, ver, out, nrows);
            CK(cudaDeviceSynchronize());
            long long h[2]; CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
            printf("%-28s %-48s %8.1f cycles = %6.3f us per iteration\n", bg ? "[+~2000 polling warps]" : "[warp alone]", names[mode], (double)h[0] / iters, (double)h[0] / iters / (clk * 1e-3));
        }
    return 0;
}
