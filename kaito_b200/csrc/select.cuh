// select.cuh -- "P smallest 64-bit keys" selection in shared memory.
//
// All candidate keys in this library are unique (the low 32 bits carry the ordinal), so
// the P smallest keys are a well-defined set and the result is independent of the order
// in which threads push candidates.  A group of NT threads (a whole CTA, or the epilogue
// warps of the tensor-core kernel) shares one buffer:
//
//   push   : key < thr  ->  buf[atomicAdd(count)] = key          (any thread, any time)
//   prune  : bitonic sort, keep the P smallest, thr = buf[P-1]    (all NT threads)
//
// Callers guarantee that at most (cap - P) pushes happen between two prunes ("epochs"),
// so the buffer never overflows and the admitted set at each prune is deterministic.
#pragma once
#include "common.cuh"

namespace krag {

struct SelectBuf {
    uint64_t* buf;   // [cap] shared
    int* count;      // shared
    uint64_t* thr;   // shared: admission threshold (exclusive upper bound)
    int cap;         // power of two
};

__device__ __forceinline__ void select_init(const SelectBuf& s, int tid)
{
    if (tid == 0) { *s.count = 0; *s.thr = KEY_PAD; }
}

__device__ __forceinline__ void select_push(const SelectBuf& s, uint64_t key, uint64_t thr_reg)
{
    if (key < thr_reg) {
        int pos = atomicAdd(s.count, 1);
        if (pos < s.cap) s.buf[pos] = key;  // never false when the epoch contract holds
    }
}

// In-place ascending bitonic sort of buf[0..n2) (n2 power of two) by NT threads.
template <int NT>
__device__ __forceinline__ void bitonic_sort_smem(uint64_t* buf, int n2, int tid, int bar_id)
{
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (n2 >> 1); t += NT) {
                // t-th compare-exchange pair of this step
                int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int p = i | j;
                bool up = ((i & k) == 0);
                uint64_t a = buf[i], b = buf[p];
                if ((a > b) == up) { buf[i] = b; buf[p] = a; }
            }
            bar_sync(bar_id, NT);
        }
    }
}

// Sort the admitted candidates, keep the P smallest, publish the new threshold.
// Must be called by all NT threads of the group; contains barriers.
template <int NT>
__device__ __forceinline__ void select_prune(const SelectBuf& s, int P, int tid, int bar_id)
{
    bar_sync(bar_id, NT);  // all pushes of the epoch are visible
    int cnt = min(*s.count, s.cap);
    int n2 = next_pow2(max(cnt, 2));
    for (int i = cnt + tid; i < n2; i += NT) s.buf[i] = KEY_PAD;
    bar_sync(bar_id, NT);
    if (cnt > 1) bitonic_sort_smem<NT>(s.buf, n2, tid, bar_id);
    if (tid == 0) {
        int keep = min(cnt, P);
        *s.count = keep;
        *s.thr = (keep == P) ? s.buf[P - 1] : KEY_PAD;
    }
    bar_sync(bar_id, NT);
}

// Write the current (pruned, sorted) list to global memory, padded to P entries.
template <int NT>
__device__ __forceinline__ void select_store(const SelectBuf& s, int P, uint64_t* out, int tid)
{
    int cnt = *s.count;
    for (int i = tid; i < P; i += NT) out[i] = (i < cnt) ? s.buf[i] : KEY_PAD;
}

}  // namespace krag
