// merge_fuse.cu -- K4: candidate-list merge, BM25 zero-score fill, and the reference's
// weighted fusion (HybridRetriever._fuse, presets/ragengine/vector_store/retriever/
// hybrid_retriever.py:132-166).  Tiny, latency-bound kernels: one CTA per query.
#include <math_constants.h>
#include <stdio.h>

#include "engine.h"
#include "select.cuh"

namespace krag {

// ------------------------------------------------------------------------------ merge
constexpr int MG_THREADS = 512;
constexpr int MG_CAP = 2048;  // >= KRAG_MAX_POOL + MG_THREADS

__global__ void __launch_bounds__(MG_THREADS)
merge_kernel(const uint64_t* __restrict__ in, int n_lists, int list_len, int P, int64_t list_stride,
             int64_t batch_stride, uint64_t* __restrict__ out, const uint64_t* __restrict__ thr_hint,
             const uint32_t* __restrict__ only_flag)
{
    if (only_flag != nullptr && only_flag[blockIdx.x] == 0) return;   // uniform per block
    __shared__ uint64_t s_buf[MG_CAP];
    __shared__ int s_count;
    __shared__ uint64_t s_thr;
    const int tid = threadIdx.x;
    SelectBuf sel{s_buf, &s_count, &s_thr, MG_CAP};
    select_init(sel, tid);
    // thr_hint[q] (optional) is an upper bound of the P-th smallest key: only keys <= hint can be in the result
    if (tid == 0 && thr_hint != nullptr && thr_hint[blockIdx.x] != KEY_PAD) s_thr = thr_hint[blockIdx.x] + 1;
    __syncthreads();
    const uint64_t* base = in + (int64_t)blockIdx.x * batch_stride;
    const int64_t total = (int64_t)n_lists * list_len;
    const int epoch = (MG_CAP - P) / MG_THREADS;  // >= 2 for P <= 1024
    uint64_t thr = s_thr;
    int it = 0;
    for (int64_t i0 = 0; i0 < total; i0 += MG_THREADS, ++it) {
        int64_t i = i0 + tid;
        if (i < total) {
            int l = (int)(i / list_len), j = (int)(i - (int64_t)l * list_len);
            uint64_t key = base[(int64_t)l * list_stride + j];
            if (key != KEY_PAD) select_push(sel, key, thr);
        }
        if ((it + 1) % epoch == 0) {
            __syncthreads();
            if (s_count + epoch * MG_THREADS > MG_CAP) select_prune<MG_THREADS>(sel, P, tid, 0);
            thr = s_thr;
        }
    }
    select_prune<MG_THREADS>(sel, P, tid, 0);
    select_store<MG_THREADS>(sel, P, out + (int64_t)blockIdx.x * P, tid);
}

void launch_merge(const uint64_t* keys_in, int n_lists, int list_len, int batch, int P, int64_t list_stride,
                  int64_t batch_stride, uint64_t* keys_out, cudaStream_t st, const uint64_t* thr_hint, const uint32_t* only_flag)
{
    merge_kernel<<<batch, MG_THREADS, 0, st>>>(keys_in, n_lists, list_len, P, list_stride, batch_stride, keys_out, thr_hint, only_flag);
    KRAG_CUDA(cudaGetLastError());
    count_launch();
}

// ------------------------------------------------------ peer-memory exchange (multi-GPU, one process per GPU)
// Replaces "NCCL all-gather, then merge" by two kernels of our own over NVLink peer memory:
//   p2p_push_kernel : block r copies this rank's candidate lists straight into rank r's mailbox (P2P stores through
//                     the IPC-mapped pointer), then publishes a sequence number with a system-scope release;
//   p2p_merge_kernel: spins (system-scope acquire) until every rank's sequence number has arrived in the LOCAL
//                     mailbox, then merges G*P -> P per (list, query) out of local memory.
// Mailbox layout (u64 words): flags[world] padded to 32 words, then data[2 slots][world][nl*B*P].  Two slots
// alternate with the sequence parity: a rank can run at most one exchange ahead of its slowest peer.
constexpr int P2P_HDR = 32;
__global__ void __launch_bounds__(1024)
p2p_push_kernel(const uint64_t* __restrict__ local, int64_t n_words, uint64_t* const* __restrict__ mailboxes, int my_rank,
                int world, int64_t slot_words, unsigned long long seq)
{
    uint64_t* mb = mailboxes[blockIdx.x];                                   // destination rank = blockIdx.x (peer or self)
    uint64_t* dst = mb + P2P_HDR + ((seq & 1ull) * world + my_rank) * slot_words;
    const uint4* s4 = reinterpret_cast<const uint4*>(local);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (int64_t i = threadIdx.x; i < n_words / 2; i += blockDim.x) d4[i] = s4[i];
    if ((n_words & 1) && threadIdx.x == 0) dst[n_words - 1] = local[n_words - 1];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(mb + my_rank), "l"(seq) : "memory");
    }
}

__global__ void __launch_bounds__(MG_THREADS)
p2p_merge_kernel(uint64_t* __restrict__ mailbox, int world, int64_t slot_words, unsigned long long seq, int nl, int batch, int P,
                 uint64_t* __restrict__ out /*[nl][batch][P]*/)
{
    __shared__ uint64_t s_buf[MG_CAP];
    __shared__ int s_count;
    __shared__ uint64_t s_thr;
    const int tid = threadIdx.x;
    const int l = blockIdx.y, q = blockIdx.x;
    if (tid < world) {                                                       // wait for every rank's lists of this exchange
        const long long t0 = clock64();
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mailbox + tid) : "memory");
            if (v < seq && clock64() - t0 > 240000000000ll)   /* ~2 minutes: a peer that is this late is gone */ { printf("p2p_merge: rank %d never arrived (seq %llu)\n", tid, seq); __trap(); }
        } while (v < seq);
    }
    SelectBuf sel{s_buf, &s_count, &s_thr, MG_CAP};
    select_init(sel, tid);
    __syncthreads();
    const uint64_t* data = mailbox + P2P_HDR + (seq & 1ull) * world * slot_words;
    const int64_t total = (int64_t)world * P;
    const int epoch = (MG_CAP - P) / MG_THREADS;
    uint64_t thr = KEY_PAD;
    int it = 0;
    for (int64_t i0 = 0; i0 < total; i0 += MG_THREADS, ++it) {
        const int64_t i = i0 + tid;
        if (i < total) {
            const int r = (int)(i / P), j = (int)(i - (int64_t)r * P);
            const uint64_t key = data[(int64_t)r * slot_words + ((int64_t)l * batch + q) * P + j];
            if (key != KEY_PAD) select_push(sel, key, thr);
        }
        if ((it + 1) % epoch == 0) {
            __syncthreads();
            if (s_count + epoch * MG_THREADS > MG_CAP) select_prune<MG_THREADS>(sel, P, tid, 0);
            thr = s_thr;
        }
    }
    select_prune<MG_THREADS>(sel, P, tid, 0);
    select_store<MG_THREADS>(sel, P, out + ((int64_t)l * batch + q) * P, tid);
}

size_t p2p_mailbox_words(int world, int nl, int max_batch, int max_P)
{
    return (size_t)P2P_HDR + (size_t)2 * world * ((size_t)nl * max_batch * max_P);
}

void launch_p2p_exchange_merge(uint64_t* const* d_mailboxes, uint64_t* own_mailbox, int rank, int world, int64_t slot_words,
                               unsigned long long seq, int nl, int batch, int P, const uint64_t* local, uint64_t* merged,
                               cudaStream_t st)
{
    const int64_t n_words = (int64_t)nl * batch * P;
    p2p_push_kernel<<<world, 1024, 0, st>>>(local, n_words, d_mailboxes, rank, world, slot_words, seq);
    KRAG_CUDA(cudaGetLastError());
    count_launch();
    p2p_merge_kernel<<<dim3((unsigned)batch, (unsigned)nl), MG_THREADS, 0, st>>>(own_mailbox, world, slot_words, seq, nl, batch, P, merged);
    KRAG_CUDA(cudaGetLastError());
    count_launch();
}

// ------------------------------------------------------------------- filter pushdown: eligible = allow & alive
__global__ void __launch_bounds__(256) bitmap_and_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ other, int64_t words)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < words) dst[i] &= other[i];
}
void launch_bitmap_and(uint32_t* dst, const uint32_t* other, int64_t words, cudaStream_t st)
{
    bitmap_and_kernel<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(dst, other, words);
    KRAG_CUDA(cudaGetLastError());
    count_launch();
}

// ------------------------------------------------------------------- BM25 zero fill
// bm25s ranks all N documents (argpartition over the full score vector), so a query that
// matches fewer than P documents still returns P entries: the rest score 0.  With the
// (score desc, ordinal asc) rule those are the lowest-ordinal live documents that are
// not already in the list (every positive-score document is in it when it is short).
__global__ void __launch_bounds__(256)
bm25_fill_kernel(uint64_t* __restrict__ keys, int P, const uint32_t* __restrict__ alive, int64_t n_rows,
                 OrdMap ord_base)
{
    __shared__ int s_m, s_found;
    uint64_t* k = keys + (int64_t)blockIdx.x * P;
    const int tid = threadIdx.x;
    if (tid == 0) {
        int m = 0;
        while (m < P && k[m] != KEY_PAD) ++m;
        s_m = m;
        s_found = 0;
    }
    __syncthreads();
    const int m = s_m;
    if (m >= P) return;
    const int need = P - m;
    const uint32_t zero_bits = ~f32_ordered_bits(0.0f);
    // candidates in ascending local row order, 256 at a time; deterministic because each
    // round's admitted set is a prefix decided by an in-order scan
    for (int64_t r0 = 0; r0 < n_rows; r0 += 256) {
        int64_t r = r0 + tid;
        bool ok = r < n_rows && (alive == nullptr || bit_test(alive, (uint32_t)r));
        if (ok) {
            uint32_t ord = ord_base + (uint32_t)r;
            for (int j = 0; j < m; ++j) if (key_ordinal(k[j]) == ord) { ok = false; break; }
        }
        // in-order compaction: position = number of ok threads before me
        __shared__ int s_scan[256];
        s_scan[tid] = ok ? 1 : 0;
        __syncthreads();
        if (tid == 0) {
            int acc = s_found;
            for (int t = 0; t < 256; ++t) { int v = s_scan[t]; s_scan[t] = acc; acc += v; }
            s_found = acc;
        }
        __syncthreads();
        if (ok && s_scan[tid] < need) k[m + s_scan[tid]] = ((uint64_t)zero_bits << 32) | (ord_base + (uint32_t)r);
        __syncthreads();
        if (s_found >= need) break;
    }
}

void launch_bm25_fill(uint64_t* keys, int batch, int P, const uint32_t* alive, int64_t n_rows, OrdMap ord_base,
                      cudaStream_t st)
{
    bm25_fill_kernel<<<batch, 256, 0, st>>>(keys, P, alive, n_rows, ord_base);
    KRAG_CUDA(cudaGetLastError());
    count_launch();
}

// ------------------------------------------------------------------------------- fuse
// final = w_v * vec + w_t * 1/(1+rank), IEEE double, no contraction (Python floats in the
// reference); sort by (final desc, ordinal asc); cut to k.
constexpr int FU_THREADS = 256;
constexpr int FU_CAP = 2 * 1024;  // >= 2 * KRAG_MAX_POOL

struct FuseEntry { uint64_t fkey; uint32_t ord; int32_t rank; float dense; float sparse; };

__device__ __forceinline__ bool fuse_less(const FuseEntry& a, const FuseEntry& b)
{
    // fkey = ~bits(final) for final >= 0 -> ascending fkey == descending final
    return a.fkey < b.fkey || (a.fkey == b.fkey && a.ord < b.ord);
}

__global__ void __launch_bounds__(FU_THREADS)
fuse_kernel(int P, int k, const uint64_t* __restrict__ dense_keys, const uint64_t* __restrict__ bm25_keys, double w_v,
            double w_t, int mode, const uint32_t* __restrict__ allow, double* __restrict__ out_final,
            float* __restrict__ out_dense, float* __restrict__ out_sparse, int32_t* __restrict__ out_rank,
            int64_t* __restrict__ out_ord, int32_t* __restrict__ out_count)
{
    extern __shared__ __align__(16) unsigned char fsm[];
    FuseEntry* e = reinterpret_cast<FuseEntry*>(fsm);          // [n2]
    __shared__ int s_nd, s_nb, s_m;
    __shared__ int s_rank[1024 + 1];                           // filtered keyword ranks (prefix sums)
    const int tid = threadIdx.x, q = blockIdx.x;
    const uint64_t* dk = dense_keys + (int64_t)q * P;
    const uint64_t* bk = bm25_keys ? bm25_keys + (int64_t)q * P : nullptr;

    if (tid == 0) {
        int nd = 0;
        while (nd < P && dk[nd] != KEY_PAD) ++nd;
        s_nd = nd;
        int nb = 0;
        if (bk) {
            // keyword-side metadata post-filter (hybrid_retriever.py:227-235): drop entries
            // that are not allowed, ranks are positions in the filtered list
            int r = 0;
            while (nb < P && bk[nb] != KEY_PAD) {
                bool ok = allow == nullptr || bit_test(allow, key_ordinal(bk[nb]));
                s_rank[nb] = ok ? r++ : -1;
                ++nb;
            }
        }
        s_nb = nb;
        s_m = nd;
    }
    __syncthreads();
    const int nd = s_nd, nb = s_nb;
    for (int i = tid; i < nd; i += FU_THREADS) {
        FuseEntry x;
        x.ord = key_ordinal(dk[i]); x.dense = key_value_asc(dk[i]); x.sparse = CUDART_NAN_F; x.rank = -1; x.fkey = 0;
        e[i] = x;
    }
    __syncthreads();
    for (int j = tid; j < nb; j += FU_THREADS) {
        int r = s_rank[j];
        if (r < 0) continue;
        uint32_t ord = key_ordinal(bk[j]);
        int hit = -1;
        for (int i = 0; i < nd; ++i) if (e[i].ord == ord) { hit = i; break; }
        if (hit < 0) {
            hit = atomicAdd(&s_m, 1);
            e[hit].ord = ord; e[hit].dense = CUDART_NAN_F;
        }
        e[hit].sparse = key_value_desc(bk[j]);
        e[hit].rank = r;
    }
    __syncthreads();
    const int m = s_m;
    const int n2 = next_pow2(max(m, 2));
    for (int i = tid; i < n2; i += FU_THREADS) {
        if (i < m) {
            double vec = 0.0;
            if (!isnan(e[i].dense)) vec = (mode == 1) ? __dsub_rn(1.0, __ddiv_rn((double)e[i].dense, 2.0)) : (double)e[i].dense;
            double txt = e[i].rank >= 0 ? __ddiv_rn(1.0, __dadd_rn(1.0, (double)e[i].rank)) : 0.0;
            double fin = __dadd_rn(__dmul_rn(w_v, vec), __dmul_rn(w_t, txt));
            // order-preserving bits of a double, complemented for descending order
            uint64_t u = (uint64_t)__double_as_longlong(fin);
            u = (u >> 63) ? ~u : (u | 0x8000000000000000ull);
            e[i].fkey = ~u;
        } else {
            e[i].fkey = ~0ull; e[i].ord = 0xffffffffu; e[i].rank = -1; e[i].dense = CUDART_NAN_F; e[i].sparse = CUDART_NAN_F;
        }
    }
    __syncthreads();
    for (int kk = 2; kk <= n2; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (n2 >> 1); t += FU_THREADS) {
                int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int p = i | j;
                bool up = ((i & kk) == 0);
                FuseEntry a = e[i], b = e[p];
                if (fuse_less(b, a) == up) { e[i] = b; e[p] = a; }
            }
            __syncthreads();
        }
    }
    const int cnt = min(m, k);
    for (int i = tid; i < k; i += FU_THREADS) {
        int64_t o = (int64_t)q * k + i;
        if (i < cnt) {
            uint64_t u = ~e[i].fkey;
            u = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
            out_final[o] = __longlong_as_double((long long)u);
            out_dense[o] = e[i].dense; out_sparse[o] = e[i].sparse; out_rank[o] = e[i].rank; out_ord[o] = (int64_t)e[i].ord;
        } else {
            out_final[o] = 0.0; out_dense[o] = CUDART_NAN_F; out_sparse[o] = CUDART_NAN_F; out_rank[o] = -1; out_ord[o] = -1;
        }
    }
    if (tid == 0) out_count[q] = cnt;
}

// vector-only fallback (hybrid_retriever.py:216-218): vector_nodes[:max_results]
__global__ void __launch_bounds__(FU_THREADS)
dense_only_kernel(int P, int k, const uint64_t* __restrict__ dense_keys, double* __restrict__ out_final,
                  float* __restrict__ out_dense, float* __restrict__ out_sparse, int32_t* __restrict__ out_rank,
                  int64_t* __restrict__ out_ord, int32_t* __restrict__ out_count)
{
    const int q = blockIdx.x, tid = threadIdx.x;
    const uint64_t* dk = dense_keys + (int64_t)q * P;
    __shared__ int s_cnt;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    int local = 0;
    for (int i = tid; i < k; i += FU_THREADS) {
        int64_t o = (int64_t)q * k + i;
        bool ok = i < P && dk[i] != KEY_PAD;
        if (ok) {
            float d = key_value_asc(dk[i]);
            out_final[o] = (double)d; out_dense[o] = d; out_ord[o] = (int64_t)key_ordinal(dk[i]);
            ++local;
        } else {
            out_final[o] = 0.0; out_dense[o] = CUDART_NAN_F; out_ord[o] = -1;
        }
        out_sparse[o] = CUDART_NAN_F; out_rank[o] = -1;
    }
    atomicAdd(&s_cnt, local);
    __syncthreads();
    if (tid == 0) out_count[q] = s_cnt;
}

void launch_fuse(int batch, int P, int k, const uint64_t* dense_keys, const uint64_t* bm25_keys, double w_v, double w_t,
                 int mode, const uint32_t* allow, double* out_final, float* out_dense, float* out_sparse,
                 int32_t* out_rank, int64_t* out_ord, int32_t* out_count, cudaStream_t st)
{
    if (bm25_keys == nullptr) {
        dense_only_kernel<<<batch, FU_THREADS, 0, st>>>(P, k, dense_keys, out_final, out_dense, out_sparse, out_rank,
                                                         out_ord, out_count);
    } else {
        const size_t smem = sizeof(FuseEntry) * (size_t)next_pow2(2 * P > 2 ? 2 * P : 2);
        static bool attr_set = false;
        if (!attr_set) {
            KRAG_CUDA(cudaFuncSetAttribute(fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(sizeof(FuseEntry) * FU_CAP)));
            attr_set = true;
        }
        fuse_kernel<<<batch, FU_THREADS, smem, st>>>(P, k, dense_keys, bm25_keys, w_v, w_t, mode, allow, out_final,
                                                      out_dense, out_sparse, out_rank, out_ord, out_count);
    }
    KRAG_CUDA(cudaGetLastError());
    count_launch();
}

}  // namespace krag
