// TEST INFRASTRUCTURE ONLY: cub::DeviceRadixSort::SortPairs as a stable host sort (tests/emu/emu.h).
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>
#include "../emu.h"
namespace cub {
struct DeviceRadixSort {
    template <class K, class V>
    static cudaError_t SortPairs(void *tmp, size_t &tmp_bytes, const K *kin, K *kout, const V *vin, V *vout, int n,
                                 int begin_bit, int end_bit, cudaStream_t = nullptr) {
        if (!tmp) { tmp_bytes = 1; return cudaSuccess; }
        const K mask = end_bit >= (int)(8 * sizeof(K)) ? ~K(0) : (K)((K(1) << end_bit) - 1);
        std::vector<int> order(n);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ((kin[a] & mask) >> begin_bit) < ((kin[b] & mask) >> begin_bit); });
        for (int i = 0; i < n; ++i) { kout[i] = kin[order[i]]; vout[i] = vin[order[i]]; }
        return cudaSuccess;
    }
};
struct DeviceScan {
    template <class In, class Out>
    static cudaError_t ExclusiveSum(void *tmp, size_t &tmp_bytes, const In *in, Out *out, int n, cudaStream_t = nullptr) {
        if (!tmp) { tmp_bytes = 1; return cudaSuccess; }
        Out acc = 0;
        for (int i = 0; i < n; ++i) { const Out v = (Out)in[i]; out[i] = acc; acc += v; }
        return cudaSuccess;
    }
};
}  // namespace cub
