// topk.cuh -- warp-held sorted candidate lists and block-level bitonic merge.
#pragma once
#include "common.cuh"

namespace evdb {

// `list` is a warp-private ascending array of KP keys in shared memory (padded
// with kKeyMax).  Insert `key` (same value in every lane, key < list[KP-1]) and
// drop the last element.  Returns the new last element (the warp's threshold).
__device__ __forceinline__ uint64_t warp_list_insert(uint64_t *list, int KP, uint64_t key,
                                                     int lane) {
    for (int base = (KP > kWarp ? KP - kWarp : 0); base >= 0; base -= kWarp) {
        int p = base + lane;
        bool in = p < KP;
        uint64_t cur = in ? list[p] : kKeyMax;
        uint64_t prev = (in && p > 0) ? list[p - 1] : 0ull;
        __syncwarp();
        uint64_t nv = cur < key ? cur : ((p == 0 || prev < key) ? key : prev);
        if (in) list[p] = nv;
        // lowest slot of this chunk already below the key: nothing further down moves
        bool low_below = __shfl_sync(0xffffffffu, (int)(cur < key), 0) != 0;
        __syncwarp();
        if (low_below) break;
    }
    return list[KP - 1];
}

// In-place ascending bitonic sort of n (power of two) u64 keys in shared memory
// by the whole CTA.  Ends with a __syncthreads().
__device__ __forceinline__ void block_bitonic_sort(uint64_t *buf, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    uint64_t a = buf[i], b = buf[ixj];
                    bool asc = (i & k) == 0;
                    if ((a > b) == asc) {
                        buf[i] = b;
                        buf[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// Same for (key, payload) pairs ordered by (key, payload).
__device__ __forceinline__ void block_bitonic_sort_pairs(uint64_t *keys, uint64_t *vals, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    uint64_t a = keys[i], b = keys[ixj];
                    uint64_t va = vals[i], vb = vals[ixj];
                    bool gt = a > b || (a == b && va > vb);
                    bool asc = (i & k) == 0;
                    if (gt == asc) {
                        keys[i] = b; keys[ixj] = a;
                        vals[i] = vb; vals[ixj] = va;
                    }
                }
            }
            __syncthreads();
        }
    }
}

}  // namespace evdb
