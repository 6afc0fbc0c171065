// tests/emu/include/cub/cub.cuh -- TEST INFRASTRUCTURE (see ../cuda_runtime.h): the two CUB device algorithms the ingest
// uses, restated with the standard library (same two-phase "size query, then run" calling convention).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <numeric>
#include <vector>

namespace cub {

struct DeviceRadixSort {
    // stable LSD sort on key bits [begin_bit, end_bit)
    template <typename K, typename V, typename N>
    static cudaError_t SortPairs(void* tmp, size_t& tmp_bytes, const K* keys_in, K* keys_out, const V* vals_in, V* vals_out, N n, int begin_bit = 0,
                                 int end_bit = (int)sizeof(K) * 8, cudaStream_t = nullptr) {
        if (!tmp) { tmp_bytes = 1; return cudaSuccess; }
        const K mask = (end_bit - begin_bit >= (int)sizeof(K) * 8) ? ~(K)0 : (K)((((K)1) << (end_bit - begin_bit)) - 1);
        std::vector<size_t> order((size_t)n);
        std::iota(order.begin(), order.end(), (size_t)0);
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return ((keys_in[a] >> begin_bit) & mask) < ((keys_in[b] >> begin_bit) & mask); });
        std::vector<K> k((size_t)n);
        std::vector<V> v((size_t)n);
        for (size_t i = 0; i < (size_t)n; i++) { k[i] = keys_in[order[i]]; v[i] = vals_in[order[i]]; }
        std::copy(k.begin(), k.end(), keys_out);
        std::copy(v.begin(), v.end(), vals_out);
        return cudaSuccess;
    }
};

struct DeviceScan {
    template <typename In, typename Out, typename N>
    static cudaError_t ExclusiveSum(void* tmp, size_t& tmp_bytes, In in, Out out, N n, cudaStream_t = nullptr) {
        if (!tmp) { tmp_bytes = 1; return cudaSuccess; }
        typename std::remove_reference<decltype(out[0])>::type run = 0;
        for (N i = 0; i < n; i++) { auto x = in[i]; out[i] = run; run += x; }
        return cudaSuccess;
    }
};

}  // namespace cub
