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

// Same result as block_bitonic_sort for n <= blockDim.x keys, one key per thread in a register:
// compare-exchange distances below 32 are warp shuffles, only the wider ones go through shared
// memory (ping-pong between buf and tmp: ONE barrier per such stage).  n = 1024 takes 15 barriers
// instead of 55.  Larger n fall back to the in-place version.  tmp: n keys of scratch.
__device__ __forceinline__ void block_bitonic_sort_fast(uint64_t *buf, int n, uint64_t *tmp) {
    if (n > (int)blockDim.x) {
        block_bitonic_sort(buf, n);
        return;
    }
    const int i = threadIdx.x;
    const bool live = (i & ~31) < n;   // warps without a key only keep the barriers company: the network
                                       // is issue-bound, 32 warps stepping through it cost 4x what 8 do
    uint64_t v = i < n ? buf[i] : ~0ull;
    uint64_t *cur = tmp;
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            uint64_t o = ~0ull;
            if (j >= 32) {
                if (i < n) cur[i] = v;
                __syncthreads();
                if (i < n) o = cur[i ^ j];
                cur = cur == tmp ? buf : tmp;
            } else if (live) {
                o = __shfl_xor_sync(0xffffffffu, v, j);
            }
            if (live) {
                const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
                v = keep_min ? (v < o ? v : o) : (v > o ? v : o);
            }
        }
    }
    __syncthreads();
    if (i < n) buf[i] = v;
    __syncthreads();
}

// (key, payload) pairs ordered by (key, payload), n <= blockDim.x; tmp: 2n words of scratch.
__device__ __forceinline__ void block_bitonic_sort_pairs_fast(uint64_t *keys, uint64_t *vals, int n, uint64_t *tmp) {
    if (n > (int)blockDim.x) {
        block_bitonic_sort_pairs(keys, vals, n);
        return;
    }
    const int i = threadIdx.x;
    const bool live = (i & ~31) < n;
    uint64_t a = i < n ? keys[i] : ~0ull, va = i < n ? vals[i] : ~0ull;
    uint64_t *ck = tmp, *cv = tmp + n;
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            uint64_t b = ~0ull, vb = ~0ull;
            if (j >= 32) {
                if (i < n) { ck[i] = a; cv[i] = va; }
                __syncthreads();
                if (i < n) { b = ck[i ^ j]; vb = cv[i ^ j]; }
                const bool first = ck == tmp;
                ck = first ? keys : tmp;
                cv = first ? vals : tmp + n;
            } else if (live) {
                b = __shfl_xor_sync(0xffffffffu, a, j);
                vb = __shfl_xor_sync(0xffffffffu, va, j);
            }
            if (live) {
                const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
                const bool mine_less = a < b || (a == b && va < vb);
                if (keep_min != mine_less && !(a == b && va == vb)) { a = b; va = vb; }
            }
        }
    }
    __syncthreads();
    if (i < n) { keys[i] = a; vals[i] = va; }
    __syncthreads();
}

// Selection without sorting: tau = the `need`-th smallest 32-bit score (orderable encoding, high
// word of the key) among up to 256 keys held 8 per lane (kKeyMax pads), by 4-way search on the
// value with warp-wide population counts.  Returns tau; *n_less = number of keys with score < tau.
__device__ __forceinline__ uint32_t warp_select_score(const uint64_t (&x)[8], int need, int *n_less) {
    uint32_t sc[8];
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        sc[r] = (uint32_t)(x[r] >> 32);
        mn = min(mn, sc[r]);
        if (x[r] != kKeyMax) mx = max(mx, sc[r]);
    }
    uint32_t lo = __reduce_min_sync(0xffffffffu, mn);
    uint32_t hi = __reduce_max_sync(0xffffffffu, mx);  // invariant: count(score <= hi) >= need
#pragma unroll 1
    while (lo < hi) {
        const uint32_t span = hi - lo;
        const uint32_t q = span >> 2;
        const uint32_t p2 = lo + (span >> 1);
        const uint32_t p1 = q ? lo + q : p2;
        const uint32_t p3 = q ? p2 + q : p2;
        int c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            c1 += sc[r] <= p1 ? 1 : 0;
            c2 += sc[r] <= p2 ? 1 : 0;
            c3 += sc[r] <= p3 ? 1 : 0;
        }
        c1 = __reduce_add_sync(0xffffffffu, c1);
        c2 = __reduce_add_sync(0xffffffffu, c2);
        c3 = __reduce_add_sync(0xffffffffu, c3);
        if (c1 >= need) hi = p1;
        else if (c2 >= need) { lo = p1 + 1; hi = p2; }
        else if (c3 >= need) { lo = p2 + 1; hi = p3; }
        else lo = p3 + 1;
    }
    int c = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) c += sc[r] < lo ? 1 : 0;
    *n_less = __reduce_add_sync(0xffffffffu, c);
    return lo;
}

// Keep the `need` best keys of x (8 per lane): those below tau first, then ties at tau, written
// densely to dst[i * stride].  Returns tau (orderable score of the need-th best).
__device__ __forceinline__ uint32_t warp_compact(const uint64_t (&x)[8], int need, uint64_t *dst,
                                                 size_t stride, int lane) {
    int n_less;
    const uint32_t tau = warp_select_score(x, need, &n_less);
    int base_less = 0, base_tie = n_less, ties_left = need - n_less;
    const unsigned below = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const uint32_t scr = (uint32_t)(x[r] >> 32);
        const bool is_less = scr < tau;
        const bool is_tie = scr == tau && x[r] != kKeyMax;
        const unsigned ml = __ballot_sync(0xffffffffu, is_less);
        const unsigned mt = __ballot_sync(0xffffffffu, is_tie);
        if (is_less) dst[(size_t)(base_less + __popc(ml & below)) * stride] = x[r];
        const int trank = __popc(mt & below);
        if (is_tie && trank < ties_left) dst[(size_t)(base_tie + trank) * stride] = x[r];
        base_less += __popc(ml);
        const int used = min(__popc(mt), ties_left);
        base_tie += used;
        ties_left -= used;
    }
    return tau;
}

// ----------------------------------------------------------------------------
// Append-and-prune candidate buffer of one warp (scans, KP <= 128).  Keys that beat the warp's
// threshold are appended to an unsorted shared-memory buffer; when it nears capacity the warp
// keeps its KP best by value bisection (no sort) and the KP-th best score becomes the new
// threshold.  A key whose score equals the threshold is dropped: every dropped key has
// score >= the final window's last score, which is all the completeness proof in select.cu needs.
// ----------------------------------------------------------------------------
constexpr int kAppendMaxKP = 128;
__host__ __device__ __forceinline__ int append_cap(int KP) { return KP <= 32 ? 96 : (KP <= 64 ? 160 : 256); }

__device__ __forceinline__ void warp_buf_prune(uint64_t *buf, int &cnt, uint64_t &thr, const int KP, const int lane) {
    __syncwarp();
    uint64_t x[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int e = r * 32 + lane;
        x[r] = e < cnt ? buf[e] : kKeyMax;
    }
    __syncwarp();
    const uint32_t tau = warp_compact(x, KP, buf, 1, lane);
    cnt = KP;
    const uint64_t t2 = (uint64_t)tau << 32;
    thr = t2 < thr ? t2 : thr;
    __syncwarp();
}

__device__ __forceinline__ void offer_append(const uint64_t key, uint64_t &thr, uint64_t *buf, int &cnt,
                                             const int cap, const int KP, const int lane) {
    const unsigned m = __ballot_sync(0xffffffffu, key < thr);
    if (m) {
        if (key < thr) buf[cnt + __popc(m & ((1u << lane) - 1u))] = key;
        cnt += __popc(m);
        if (cnt > cap - 32) warp_buf_prune(buf, cnt, thr, KP, lane);
    }
}

}  // namespace evdb
