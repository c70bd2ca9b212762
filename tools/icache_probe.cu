// One warp, the same number of DFMAs per outer iteration, straight-line bodies of different code size (tools/; not part
// of the library): does a single warp slow down when its loop body outgrows the instruction caches?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/icache_probe tools/icache_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int BODY>  // BODY DFMAs per chain, 8 chains, fully unrolled => 8*BODY instructions of straight-line code
__global__ void probe(double *out, long long *cyc, int iters, double a, double b) {
    double z[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = 1.0 + j + threadIdx.x * 1e-9;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < BODY; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) z[j] = fma(z[j], a, b);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += z[j];
    out[threadIdx.x] = s;
}
template <int BODY>
void run(double *out, long long *cyc) {
    const int total = 1 << 18;  // DFMAs per chain overall
    for (int rep = 0; rep < 2; ++rep) probe<BODY><<<1, 32>>>(out, cyc, total / BODY, 1.0000001, 1e-9);
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("body %5d instructions (%4d KB of SASS): %.2f cycles per DFMA\n", 8 * BODY, 8 * BODY * 16 / 1024, (double)h / (8.0 * total));
}
int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 64);
    run<16>(out, cyc); run<64>(out, cyc); run<128>(out, cyc); run<256>(out, cyc); run<512>(out, cyc); run<1024>(out, cyc); run<2048>(out, cyc);
    return 0;
}
