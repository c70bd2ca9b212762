// One-warp micro-probe of what bounds the funk-SVD chain (tools/; not part of the library):
// DFMA dependent-chain latency, DFMA issue rate with 4 / 8 independent chains, 64-bit shuffle latency, st.global.cg rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_probe tools/fp64_probe.cu && ./tools/fp64_probe
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(double *out, long long *cyc, double *buf) {
    const int lane = threadIdx.x;
    double x = 1.0 + lane * 1e-9, a = 1.0000001, b = 1e-9;
    long long t0, t1;
    // (1) dependent DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; ++i) x = fma(x, a, b);
    t1 = clock64();
    cyc[0] = t1 - t0;
    // (2) 4 independent chains
    double y0 = x, y1 = x + 1, y2 = x + 2, y3 = x + 3;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) { y0 = fma(y0, a, b); y1 = fma(y1, a, b); y2 = fma(y2, a, b); y3 = fma(y3, a, b); }
    t1 = clock64();
    cyc[1] = t1 - t0;
    // (3) 8 independent chains
    double z[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = y0 + j;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) z[j] = fma(z[j], a, b);
    t1 = clock64();
    cyc[2] = t1 - t0;
    // (4) dependent 64-bit shuffle + add (one butterfly step), 64 of them
    double s = z[0] + y1 + y2 + y3;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) s += __shfl_xor_sync(0xffffffffu, s, 1 + (i & 15));
    t1 = clock64();
    cyc[3] = t1 - t0;
    // (5) 64 st.global.cg of 8 B per lane (256 B per instruction), rows of 1 KB
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; ++i) __stcg(buf + (size_t)i * 128 + lane, s + i);
    t1 = clock64();
    cyc[4] = t1 - t0;
    // (6) 32 independent ld.global.cg (L2 hits after the stores) consumed at the end
    double acc = 0;
    t0 = clock64();
    double r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = __ldcg(buf + (size_t)i * 128 + lane);
#pragma unroll
    for (int i = 0; i < 32; ++i) acc += r[i];
    t1 = clock64();
    cyc[5] = t1 - t0;
    // (7) fence.acq_rel.gpu after 32 stores
#pragma unroll
    for (int i = 0; i < 32; ++i) __stcg(buf + (size_t)i * 128 + lane, acc + i);
    t0 = clock64();
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    t1 = clock64();
    cyc[6] = t1 - t0;
    out[lane] = s + acc + z[1] + z[2] + z[3] + z[4] + z[5] + z[6] + z[7];
}
int main() {
    double *out, *buf; long long *cyc, h[8];
    cudaMalloc(&out, 32 * 8); cudaMalloc(&buf, 64 * 128 * 8 + 4096); cudaMalloc(&cyc, 64);
    for (int rep = 0; rep < 3; ++rep) {
        probe<<<1, 32>>>(out, cyc, buf);
        cudaMemcpy(h, cyc, 56, cudaMemcpyDeviceToHost);
        printf("dep DFMA %.1f cyc/op | 4 chains %.1f cyc/op | 8 chains %.1f cyc/op | shfl64+add %.1f cyc/step | stcg %.1f cyc/inst | 32 ldcg+sum %lld cyc | fence after 32 stores %lld cyc\n",
               h[0] / 256.0, h[1] / 256.0, h[2] / 256.0, h[3] / 64.0, h[4] / 64.0, h[5], h[6]);
    }
    return 0;
}
