// bm25.cu -- K3: BM25 postings build (index time) and score/accumulate/select (query time).
//
// Replaces, with identical arithmetic, what the reference does on EVERY query at
// presets/ragengine/vector_store/retriever/hybrid_retriever.py:104-130 (BM25Retriever.
// from_defaults -> bm25s "lucene" score matrix) and :220 (bm25_retriever.aretrieve):
//   score(t,d) = (f32)( (f64) idf32[t] * tf / (k1*((1-b) + b*dl/avgdl) + tf) ),  k1=1.5, b=0.75
//   acc[d]    += score(t,d)  for the query's term ids in query order (fp32, duplicates kept)
//   result     = top-P by (score desc, ordinal asc)
// idf32 is computed on the host with glibc log() (the library the reference's math.log
// uses); every other operation is a correctly rounded IEEE op issued without contraction.
//
// Query kernel: one CTA per (doc-range tile, query).  The tile's fp32 accumulators live in
// shared memory, postings are read exactly once with coalesced loads (8 bytes per posting
// = the algorithmic traffic), terms are applied one after another (doc ids are unique
// inside a term, so no atomics and a fixed fp32 summation order), and the top-P select
// runs over shared memory.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "engine.h"
#include "select.cuh"

namespace krag {

// -------------------------------------------------------------------------- index time
__global__ void expand_entry_doc_kernel(const int64_t* __restrict__ off, int64_t n_docs, uint32_t* __restrict__ entry_doc)
{
    // one warp per document
    int64_t d = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (d >= n_docs) return;
    for (int64_t i = off[d] + lane; i < off[d + 1]; i += 32) entry_doc[i] = (uint32_t)d;
}
void launch_expand_entry_doc(const int64_t* term_offsets, int64_t n_docs, uint32_t* entry_doc, cudaStream_t st)
{
    if (n_docs == 0) return;
    int64_t threads = n_docs * 32;
    expand_entry_doc_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(term_offsets, n_docs, entry_doc);
    KRAG_CUDA(cudaGetLastError());
    count_launch();
}

__global__ void df_hist_kernel(const uint32_t* __restrict__ term_ids, const uint32_t* __restrict__ entry_doc,
                               const uint32_t* __restrict__ alive, int64_t nnz, uint32_t* __restrict__ df)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x)
        if (alive == nullptr || bit_test(alive, entry_doc[i])) atomicAdd(&df[term_ids[i]], 1u);
}
void launch_df_histogram(const uint32_t* term_ids, const uint32_t* entry_doc, const uint32_t* alive, int64_t nnz,
                         uint32_t* df, cudaStream_t st)
{
    if (nnz == 0) return;
    df_hist_kernel<<<148 * 8, 256, 0, st>>>(term_ids, entry_doc, alive, nnz, df);
    KRAG_CUDA(cudaGetLastError());
    count_launch();
}

__global__ void pack_sort_values_kernel(const uint32_t* __restrict__ term_ids, const uint16_t* __restrict__ term_tf,
                                        const uint32_t* __restrict__ entry_doc, const uint32_t* __restrict__ alive,
                                        int64_t nnz, uint32_t dead_key, uint32_t* __restrict__ keys,
                                        uint64_t* __restrict__ vals)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t d = entry_doc[i];
        bool ok = alive == nullptr || bit_test(alive, d);
        keys[i] = ok ? term_ids[i] : dead_key;  // tombstoned docs sort past the last term
        vals[i] = ((uint64_t)d << 16) | term_tf[i];
    }
}

__global__ void score_postings_kernel(const uint32_t* __restrict__ sorted_terms, const uint64_t* __restrict__ sorted_vals,
                                      const uint32_t* __restrict__ doc_len, const float* __restrict__ idf, double avgdl,
                                      int64_t nnz_live, uint32_t* __restrict__ post_doc, float* __restrict__ post_score)
{
    const double k1 = 1.5, b = 0.75;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz_live; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t v = sorted_vals[i];
        uint32_t d = (uint32_t)(v >> 16);
        double tf = (double)(uint32_t)(v & 0xffffu);
        // tf / (k1 * ((1 - b) + b * dl / avgdl) + tf), left-to-right as Python evaluates it
        double t1 = __dmul_rn(b, (double)doc_len[d]);
        double t2 = __ddiv_rn(t1, avgdl);
        double t3 = __dadd_rn(__dsub_rn(1.0, b), t2);
        double t4 = __dmul_rn(k1, t3);
        double t5 = __dadd_rn(t4, tf);
        double tfc = __ddiv_rn(tf, t5);
        post_doc[i] = d;
        post_score[i] = __double2float_rn(__dmul_rn((double)idf[sorted_terms[i]], tfc));
    }
}

__global__ void df_to_i64_kernel(const uint32_t* __restrict__ df_local, int64_t vocab, int64_t* __restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= vocab; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = i < vocab ? (int64_t)df_local[i] : 0;
}

void build_postings(const uint32_t* term_ids, const uint16_t* term_tf, const uint32_t* entry_doc,
                    const uint32_t* doc_len, const uint32_t* alive, int64_t nnz, int64_t vocab, const float* idf,
                    double avgdl, Postings& out, cudaStream_t st)
{
    // local df (live docs only) -> exclusive scan -> offsets
    uint32_t* df_local = nullptr;
    int64_t* off = nullptr;
    KRAG_CUDA(cudaMalloc(&df_local, sizeof(uint32_t) * (size_t)(vocab + 1)));
    KRAG_CUDA(cudaMemsetAsync(df_local, 0, sizeof(uint32_t) * (size_t)(vocab + 1), st));
    launch_df_histogram(term_ids, entry_doc, alive, nnz, df_local, st);
    KRAG_CUDA(cudaMalloc(&off, sizeof(int64_t) * (size_t)(vocab + 1)));
    int64_t* tmp64 = nullptr;
    KRAG_CUDA(cudaMalloc(&tmp64, sizeof(int64_t) * (size_t)(vocab + 1)));
    df_to_i64_kernel<<<256, 256, 0, st>>>(df_local, vocab, tmp64);
    count_launch();
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, tmp64, off, (int)(vocab + 1), st);
    void* scan_tmp = nullptr;
    KRAG_CUDA(cudaMalloc(&scan_tmp, scan_bytes ? scan_bytes : 16));
    cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, tmp64, off, (int)(vocab + 1), st);
    count_launch();
    int64_t nnz_live = 0;
    KRAG_CUDA(cudaMemcpyAsync(&nnz_live, off + vocab, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    KRAG_CUDA(cudaStreamSynchronize(st));
    KRAG_CUDA(cudaFree(scan_tmp));
    KRAG_CUDA(cudaFree(tmp64));
    KRAG_CUDA(cudaFree(df_local));

    uint32_t* post_doc = nullptr;
    float* post_score = nullptr;
    KRAG_CUDA(cudaMalloc(&post_doc, sizeof(uint32_t) * (size_t)(nnz_live > 0 ? nnz_live : 1)));
    KRAG_CUDA(cudaMalloc(&post_score, sizeof(float) * (size_t)(nnz_live > 0 ? nnz_live : 1)));

    if (nnz > 0) {
        // stable LSD radix sort by term id keeps documents ascending inside every term
        uint32_t *k_in = nullptr, *k_out = nullptr;
        uint64_t *v_in = nullptr, *v_out = nullptr;
        KRAG_CUDA(cudaMalloc(&k_in, sizeof(uint32_t) * (size_t)nnz));
        KRAG_CUDA(cudaMalloc(&k_out, sizeof(uint32_t) * (size_t)nnz));
        KRAG_CUDA(cudaMalloc(&v_in, sizeof(uint64_t) * (size_t)nnz));
        KRAG_CUDA(cudaMalloc(&v_out, sizeof(uint64_t) * (size_t)nnz));
        pack_sort_values_kernel<<<148 * 8, 256, 0, st>>>(term_ids, term_tf, entry_doc, alive, nnz, (uint32_t)vocab,
                                                        k_in, v_in);
        count_launch();
        int end_bit = 1;
        while (((int64_t)1 << end_bit) <= vocab) ++end_bit;  // keys range over [0, vocab]
        size_t sort_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, k_in, k_out, v_in, v_out, nnz, 0, end_bit, st);
        void* sort_tmp = nullptr;
        KRAG_CUDA(cudaMalloc(&sort_tmp, sort_bytes ? sort_bytes : 16));
        cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, k_in, k_out, v_in, v_out, nnz, 0, end_bit, st);
        count_launch();
        if (nnz_live > 0) {
            score_postings_kernel<<<148 * 8, 256, 0, st>>>(k_out, v_out, doc_len, idf, avgdl, nnz_live, post_doc,
                                                          post_score);
            count_launch();
        }
        KRAG_CUDA(cudaStreamSynchronize(st));
        KRAG_CUDA(cudaFree(sort_tmp));
        KRAG_CUDA(cudaFree(k_in)); KRAG_CUDA(cudaFree(k_out));
        KRAG_CUDA(cudaFree(v_in)); KRAG_CUDA(cudaFree(v_out));
    }
    if (out.off) cudaFree(out.off);
    if (out.doc) cudaFree(out.doc);
    if (out.score) cudaFree(out.score);
    out.off = off; out.doc = post_doc; out.score = post_score; out.vocab = vocab; out.nnz = nnz_live;
}

// -------------------------------------------------------------------------- query time
constexpr int BQ_THREADS = 512;
constexpr int BQ_WARPS = BQ_THREADS / 32;
constexpr int BQ_MAX_TERMS = 64;  // terms applied per pass; longer queries loop

// warp-cooperative lower_bound over a sorted u32 range: first index in [lo,hi) with a[i] >= target
__device__ __forceinline__ int64_t warp_lower_bound(const uint32_t* __restrict__ a, int64_t lo, int64_t hi, uint32_t target,
                                                    int lane)
{
    while (hi - lo > 32) {
        int64_t step = (hi - lo + 31) / 32;
        int64_t idx = lo + (int64_t)lane * step;
        bool less = idx < hi && a[idx] < target;
        unsigned m = __ballot_sync(0xffffffffu, less);
        int c = __popc(m);  // probes below target form a prefix
        if (c == 0) return lo;
        int64_t nlo = lo + (int64_t)(c - 1) * step + 1;
        int64_t nhi = lo + (int64_t)c * step;
        lo = nlo;
        hi = nhi < hi ? nhi : hi;
    }
    bool less = lo + lane < hi && a[lo + lane] < target;
    return lo + __popc(__ballot_sync(0xffffffffu, less));
}

__global__ void __launch_bounds__(BQ_THREADS)
bm25_tile_kernel(const int64_t* __restrict__ post_off, const uint32_t* __restrict__ post_doc,
                 const float* __restrict__ post_score, int64_t vocab, const uint32_t* __restrict__ q_terms,
                 const int32_t* __restrict__ q_term_offsets, int64_t n_rows, const uint32_t* __restrict__ alive, int P,
                 int cap, uint32_t ord_base, uint64_t* __restrict__ part /*[batch][n_tiles][P]*/)
{
    extern __shared__ __align__(16) unsigned char bsm[];
    float* acc = reinterpret_cast<float*>(bsm);                                          // [BM25_TILE_DOCS]
    uint64_t* sbuf = reinterpret_cast<uint64_t*>(bsm + (size_t)BM25_TILE_DOCS * 4);      // [cap]
    __shared__ int s_count;
    __shared__ uint64_t s_thr;
    __shared__ int64_t s_lo[BQ_MAX_TERMS], s_hi[BQ_MAX_TERMS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, qi = blockIdx.y;
    const int64_t t0 = (int64_t)tile * BM25_TILE_DOCS;
    const int tile_n = (int)min((int64_t)BM25_TILE_DOCS, n_rows - t0);
    const int tb = q_term_offsets[qi], te = q_term_offsets[qi + 1];

    for (int i = tid; i < BM25_TILE_DOCS / 4; i += BQ_THREADS) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    SelectBuf sel{sbuf, &s_count, &s_thr, cap};
    select_init(sel, tid);

    for (int c0 = tb; c0 < te; c0 += BQ_MAX_TERMS) {
        const int nt = min(BQ_MAX_TERMS, te - c0);
        __syncthreads();
        // posting sub-ranges of this tile, one warp per term
        for (int j = warp; j < nt; j += BQ_WARPS) {
            uint32_t t = q_terms[c0 + j];
            int64_t lo = 0, hi = 0;
            if ((int64_t)t < vocab) {
                int64_t b = post_off[t], e = post_off[t + 1];
                lo = warp_lower_bound(post_doc, b, e, (uint32_t)t0, lane);
                hi = warp_lower_bound(post_doc, lo, e, (uint32_t)(t0 + tile_n), lane);
            }
            if (lane == 0) { s_lo[j] = lo; s_hi[j] = hi; }
        }
        __syncthreads();
        // apply the terms in query order; a barrier between terms fixes the fp32 order
        for (int j = 0; j < nt; ++j) {
            const int64_t lo = s_lo[j], hi = s_hi[j];
            for (int64_t p = lo + tid; p < hi; p += BQ_THREADS) acc[post_doc[p] - (uint32_t)t0] += post_score[p];
            __syncthreads();
        }
    }
    __syncthreads();
    // select over the tile: only matched (score > 0), live documents are candidates
    const int epoch = (cap - P) / BQ_THREADS > 0 ? (cap - P) / BQ_THREADS : 1;
    uint64_t thr = KEY_PAD;
    int it = 0;
    for (int i0 = 0; i0 < tile_n; i0 += BQ_THREADS, ++it) {
        int i = i0 + tid;
        if (i < tile_n) {
            float s = acc[i];
            if (s > 0.f && (alive == nullptr || bit_test(alive, (uint32_t)(t0 + i))))
                select_push(sel, make_key_desc(s, ord_base + (uint32_t)(t0 + i)), thr);
        }
        if ((it + 1) % epoch == 0) {
            __syncthreads();
            if (s_count + epoch * BQ_THREADS > cap) select_prune<BQ_THREADS>(sel, P, tid, 0);
            thr = s_thr;
        }
    }
    select_prune<BQ_THREADS>(sel, P, tid, 0);
    select_store<BQ_THREADS>(sel, P, part + ((size_t)qi * gridDim.x + tile) * P, tid);
}

static int bq_cap(int P) { return P <= 512 ? 1024 : 2048; }

size_t bm25_part_elems(int64_t n_rows, int batch, int P)
{
    int64_t n_tiles = (n_rows + BM25_TILE_DOCS - 1) / BM25_TILE_DOCS;
    if (n_tiles < 1) n_tiles = 1;
    return (size_t)n_tiles * batch * P;
}

void launch_bm25(const DeviceInfo& di, const Postings& post, int64_t n_rows, const uint32_t* alive,
                 const uint32_t* q_terms, const int32_t* q_term_offsets, int max_terms, int batch, int P,
                 uint32_t ord_base, uint64_t* part, uint64_t* keys_out, cudaStream_t st)
{
    (void)di; (void)max_terms;
    int64_t n_tiles = (n_rows + BM25_TILE_DOCS - 1) / BM25_TILE_DOCS;
    if (n_tiles < 1) n_tiles = 1;
    const int cap = bq_cap(P);
    const size_t smem = (size_t)BM25_TILE_DOCS * 4 + (size_t)cap * 8;
    static bool attr_set = false;
    if (!attr_set) {
        KRAG_CUDA(cudaFuncSetAttribute(bm25_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr_set = true;
    }
    // grid.y is limited to 65535: split very large batches
    for (int b0 = 0; b0 < batch; b0 += 32768) {
        int nb = batch - b0 < 32768 ? batch - b0 : 32768;
        dim3 grid((unsigned)n_tiles, (unsigned)nb);
        bm25_tile_kernel<<<grid, BQ_THREADS, smem, st>>>(post.off, post.doc, post.score, post.vocab, q_terms,
                                                         q_term_offsets + b0, n_rows, alive, P, cap, ord_base,
                                                         part + (size_t)b0 * n_tiles * P);
        KRAG_CUDA(cudaGetLastError());
        count_launch();
    }
    launch_merge(part, (int)n_tiles, P, batch, P, /*list_stride=*/P, /*batch_stride=*/n_tiles * P, keys_out, st);
    launch_bm25_fill(keys_out, batch, P, alive, n_rows, ord_base, st);
}

}  // namespace krag
