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

#include <stdlib.h>
#include <vector>

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

// ---- tile index (frequent terms only)
__global__ void tile_flag_kernel(const int64_t* __restrict__ off, int64_t vocab, int32_t* __restrict__ flag)
{
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t <= vocab; t += (int64_t)gridDim.x * blockDim.x)
        flag[t] = (t < vocab && off[t + 1] - off[t] > BM25_RARE_MAX) ? 1 : 0;
}
__global__ void tile_slot_kernel(const int32_t* __restrict__ flag, const int32_t* __restrict__ scan, int64_t vocab,
                                 int32_t* __restrict__ slot)
{
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < vocab; t += (int64_t)gridDim.x * blockDim.x)
        slot[t] = flag[t] ? scan[t] : -1;
}
__global__ void tile_index_kernel(const uint32_t* __restrict__ sorted_terms, const uint32_t* __restrict__ post_doc,
                                  const int64_t* __restrict__ off, const int32_t* __restrict__ slot, int64_t nnz_live,
                                  int64_t n_tiles, uint32_t* __restrict__ tile_off)
{
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz_live; p += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t t = sorted_terms[p];
        const int32_t sl = slot[t];
        if (sl < 0) continue;
        const int64_t b = off[t], e = off[t + 1];
        const int64_t i = p - b;
        const int64_t tile_d = post_doc[p] / BM25_TILE_DOCS;
        const int64_t prev = (i == 0) ? -1 : (int64_t)(post_doc[p - 1] / BM25_TILE_DOCS);
        uint32_t* row = tile_off + (int64_t)sl * (n_tiles + 1);
        for (int64_t tl = prev + 1; tl <= tile_d; ++tl) row[tl] = (uint32_t)i;
        if (p == e - 1) for (int64_t tl = tile_d + 1; tl <= n_tiles; ++tl) row[tl] = (uint32_t)(e - b);
    }
}
// device scratch of the index build: everything allocated through it is freed on scope exit (also when a CUDA call
// throws half-way), except what keep() hands over to the Postings
struct DevScratch {
    std::vector<void*> ptrs;
    template <typename T> T* alloc(size_t n)
    {
        void* p = nullptr;
        KRAG_CUDA(cudaMalloc(&p, sizeof(T) * (n ? n : 1)));
        ptrs.push_back(p);
        return static_cast<T*>(p);
    }
    void free_now(void* p)
    {
        for (auto& q : ptrs) if (q == p && q) { cudaFree(q); q = nullptr; }
    }
    template <typename T> T* keep(T* p)
    {
        for (auto& q : ptrs) if (q == static_cast<void*>(p)) q = nullptr;
        return p;
    }
    ~DevScratch() { for (void* p : ptrs) if (p) cudaFree(p); }
};

static void build_tile_index(const uint32_t* sorted_terms, const uint32_t* post_doc, const int64_t* off, int64_t nnz_live,
                             int64_t vocab, int64_t n_rows, Postings& out, cudaStream_t st)
{
    DevScratch sc;
    int32_t* flag = sc.alloc<int32_t>((size_t)(vocab + 1));
    int32_t* scan = sc.alloc<int32_t>((size_t)(vocab + 1));
    int32_t* slot = sc.alloc<int32_t>((size_t)vocab);
    tile_flag_kernel<<<256, 256, 0, st>>>(off, vocab, flag);
    count_launch();
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, flag, scan, (int)(vocab + 1), st);
    void* tmp = sc.alloc<unsigned char>(bytes);
    cub::DeviceScan::ExclusiveSum(tmp, bytes, flag, scan, (int)(vocab + 1), st);
    count_launch();
    tile_slot_kernel<<<256, 256, 0, st>>>(flag, scan, vocab, slot);
    count_launch();
    int32_t n_slots = 0;
    KRAG_CUDA(cudaMemcpyAsync(&n_slots, scan + vocab, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    KRAG_CUDA(cudaStreamSynchronize(st));
    sc.free_now(tmp); sc.free_now(flag); sc.free_now(scan);
    int64_t n_tiles = (n_rows + BM25_TILE_DOCS - 1) / BM25_TILE_DOCS;
    if (n_tiles < 1) n_tiles = 1;
    uint32_t* tile_off = sc.alloc<uint32_t>((size_t)((int64_t)(n_slots > 0 ? n_slots : 1) * (n_tiles + 1)));
    if (n_slots > 0) {
        tile_index_kernel<<<148 * 8, 256, 0, st>>>(sorted_terms, post_doc, off, slot, nnz_live, n_tiles, tile_off);
        KRAG_CUDA(cudaGetLastError());
        count_launch();
    }
    if (out.tile_slot) cudaFree(out.tile_slot);
    if (out.tile_off) cudaFree(out.tile_off);
    sc.keep(slot); sc.keep(tile_off);
    out.tile_slot = slot; out.tile_off = tile_off; out.n_slots = n_slots; out.n_tiles = n_tiles;
}

void build_postings(const uint32_t* term_ids, const uint16_t* term_tf, const uint32_t* entry_doc,
                    const uint32_t* doc_len, const uint32_t* alive, int64_t nnz, int64_t vocab, const float* idf,
                    double avgdl, int64_t n_docs_rows, Postings& out, cudaStream_t st)
{
    // local df (live docs only) -> exclusive scan -> offsets
    DevScratch sc;
    uint32_t* df_local = sc.alloc<uint32_t>((size_t)(vocab + 1));
    KRAG_CUDA(cudaMemsetAsync(df_local, 0, sizeof(uint32_t) * (size_t)(vocab + 1), st));
    launch_df_histogram(term_ids, entry_doc, alive, nnz, df_local, st);
    int64_t* off = sc.alloc<int64_t>((size_t)(vocab + 1));
    int64_t* tmp64 = sc.alloc<int64_t>((size_t)(vocab + 1));
    df_to_i64_kernel<<<256, 256, 0, st>>>(df_local, vocab, tmp64);
    count_launch();
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, tmp64, off, (int)(vocab + 1), st);
    void* scan_tmp = sc.alloc<unsigned char>(scan_bytes);
    cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, tmp64, off, (int)(vocab + 1), st);
    count_launch();
    int64_t nnz_live = 0;
    KRAG_CUDA(cudaMemcpyAsync(&nnz_live, off + vocab, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    KRAG_CUDA(cudaStreamSynchronize(st));
    sc.free_now(scan_tmp); sc.free_now(tmp64); sc.free_now(df_local);

    uint32_t* post_doc = sc.alloc<uint32_t>((size_t)(nnz_live > 0 ? nnz_live : 1));
    float* post_score = sc.alloc<float>((size_t)(nnz_live > 0 ? nnz_live : 1));

    if (nnz > 0) {
        // stable LSD radix sort by term id keeps documents ascending inside every term
        uint32_t* k_in = sc.alloc<uint32_t>((size_t)nnz);
        uint32_t* k_out = sc.alloc<uint32_t>((size_t)nnz);
        uint64_t* v_in = sc.alloc<uint64_t>((size_t)nnz);
        uint64_t* v_out = sc.alloc<uint64_t>((size_t)nnz);
        pack_sort_values_kernel<<<148 * 8, 256, 0, st>>>(term_ids, term_tf, entry_doc, alive, nnz, (uint32_t)vocab,
                                                        k_in, v_in);
        count_launch();
        int end_bit = 1;
        while (((int64_t)1 << end_bit) <= vocab) ++end_bit;  // keys range over [0, vocab]
        size_t sort_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, k_in, k_out, v_in, v_out, nnz, 0, end_bit, st);
        void* sort_tmp = sc.alloc<unsigned char>(sort_bytes);
        cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, k_in, k_out, v_in, v_out, nnz, 0, end_bit, st);
        count_launch();
        if (nnz_live > 0) {
            score_postings_kernel<<<148 * 8, 256, 0, st>>>(k_out, v_out, doc_len, idf, avgdl, nnz_live, post_doc,
                                                          post_score);
            count_launch();
            build_tile_index(k_out, post_doc, off, nnz_live, vocab, n_docs_rows, out, st);
        }
        KRAG_CUDA(cudaStreamSynchronize(st));
        sc.free_now(sort_tmp);
        sc.free_now(k_in); sc.free_now(k_out); sc.free_now(v_in); sc.free_now(v_out);
    }
    if (nnz_live == 0 || nnz == 0) {   // no postings: every term is "rare" with an empty list
        int32_t* slot = sc.alloc<int32_t>((size_t)vocab);
        KRAG_CUDA(cudaMemset(slot, 0xFF, sizeof(int32_t) * (size_t)vocab));
        uint32_t* toff = sc.alloc<uint32_t>(4);
        if (out.tile_slot) cudaFree(out.tile_slot);
        if (out.tile_off) cudaFree(out.tile_off);
        out.tile_slot = sc.keep(slot); out.tile_off = sc.keep(toff); out.n_slots = 0;
        out.n_tiles = (n_docs_rows + BM25_TILE_DOCS - 1) / BM25_TILE_DOCS < 1 ? 1 : (n_docs_rows + BM25_TILE_DOCS - 1) / BM25_TILE_DOCS;
    }
    if (out.off) cudaFree(out.off);
    if (out.doc) cudaFree(out.doc);
    if (out.score) cudaFree(out.score);
    sc.keep(off); sc.keep(post_doc); sc.keep(post_score);
    out.off = off; out.doc = post_doc; out.score = post_score; out.vocab = vocab; out.nnz = nnz_live;
}

// -------------------------------------------------------------------------- query time
constexpr int BQ_THREADS = 256;   // a term contributes ~200 postings to a 16384-doc tile: 256-wide slabs keep the lanes busy
constexpr int BQ_MAX_TERMS = 32;    // query terms resolved per pass; longer queries loop
constexpr int BQ_MAX_SLABS = 96;    // 512-posting slabs per pass
constexpr int BQ_PREFETCH = 8;      // slabs held in registers at a time
constexpr int BQ_GROUP_MAX = 32;    // upper bound of the group size (shared-memory table)
// consecutive doc tiles handled by one work item (same query): one resolve, one select, one store.  Runtime value
// (KRAG_BM25_GROUP, default 8): larger groups mean fewer per-item sorts/stores and tighter own thresholds, smaller
// ones more items to balance over the 444 resident CTAs
static int bq_group()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("KRAG_BM25_GROUP"); int g = e ? atoi(e) : 8; v = g < 1 ? 1 : (g > BQ_GROUP_MAX ? BQ_GROUP_MAX : g); }
    return v;
}

// A "slab" is up to 512 consecutive postings of one query term that fall into this CTA's doc
// range (frequent terms: looked up in the tile index; rare terms: the whole <= 256-entry list,
// filtered by doc).  Slabs are applied in query-term order with a barrier in between, which
// fixes the per-document fp32 summation order to the oracle's.
__global__ void __launch_bounds__(BQ_THREADS)
bm25_tile_kernel(const int64_t* __restrict__ post_off, const uint32_t* __restrict__ post_doc,
                 const float* __restrict__ post_score, const int32_t* __restrict__ tile_slot,
                 const uint32_t* __restrict__ tile_off, int64_t n_tiles_idx, int64_t vocab,
                 const uint32_t* __restrict__ q_terms, const int32_t* __restrict__ q_term_offsets,
                 const int32_t* __restrict__ q_slot, const int64_t* __restrict__ q_base, const int32_t* __restrict__ q_rare_len,
                 int64_t n_rows, const uint32_t* __restrict__ alive, int P, int cap, uint32_t ord_base, int batch, int n_tiles, int group,
                 uint64_t* __restrict__ part /*[batch][n_groups][P]*/, unsigned long long* __restrict__ g_thr /*[batch]*/)
{
    extern __shared__ __align__(16) unsigned char bsm[];
    float* acc = reinterpret_cast<float*>(bsm);                                          // [BM25_TILE_DOCS]
    uint64_t* sbuf = reinterpret_cast<uint64_t*>(bsm + (size_t)BM25_TILE_DOCS * 4);      // [cap]
    __shared__ int s_count;
    __shared__ uint64_t s_thr;
    __shared__ int64_t s_lo[BQ_MAX_TERMS];
    __shared__ int s_len[BQ_MAX_TERMS];
    __shared__ int64_t s_slab_lo[BQ_MAX_SLABS];
    __shared__ int s_slab_n[BQ_MAX_SLABS];
    __shared__ int s_nslab, s_next_term, s_next_off;
    __shared__ int32_t s_slot[BQ_MAX_TERMS], s_rare[BQ_MAX_TERMS];
    __shared__ int64_t s_base[BQ_MAX_TERMS];
    __shared__ uint32_t s_toff[BQ_MAX_TERMS][BQ_GROUP_MAX + 1];

    const int tid = threadIdx.x;
    // the accumulators are zeroed once: every touched entry is reset by the claim step
    for (int i = tid; i < BM25_TILE_DOCS / 4; i += BQ_THREADS) reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    SelectBuf sel{sbuf, &s_count, &s_thr, cap};

    // persistent loop over (tile group, query) work items, query fastest.  A work item covers BQ_GROUP consecutive
    // doc tiles of one query: the term -> posting-range lookups are fetched once for the whole group, the select
    // buffer (and its threshold) carries over from tile to tile, and one top-P list is stored per item.
    // g_thr[q] (min over finished items of their P-th best key -- an upper bound of the global P-th best) prunes
    // almost every candidate of later items before it reaches the select buffer.
    const int n_groups = (n_tiles + group - 1) / group;
    const int64_t n_items = (int64_t)n_groups * batch;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int tg = (int)(item / batch), qi = (int)(item - (int64_t)tg * batch);
    const int tile0 = tg * group, gcount = min(group, n_tiles - tile0);
    const int tb = q_term_offsets[qi], te = q_term_offsets[qi + 1];
    const bool single_chunk = (te - tb <= BQ_MAX_TERMS);
    __syncthreads();   // previous item fully stored
    if (tid == 0) {
        const unsigned long long h = __ldcg(&g_thr[qi]);
        s_count = 0;
        s_thr = (h == KEY_PAD) ? KEY_PAD : h + 1;   // admit keys <= hint
    }
    if (single_chunk && tid < te - tb) {
        // one fetch per item: slot/base of the term and its posting offsets at the group's tile boundaries
        const int32_t sl = q_slot[tb + tid];
        s_slot[tid] = sl; s_base[tid] = q_base[tb + tid]; s_rare[tid] = q_rare_len[tb + tid];
        if (sl >= 0) {
            const uint32_t* row = tile_off + (int64_t)sl * (n_tiles_idx + 1) + tile0;
            for (int g = 0; g <= gcount; ++g) s_toff[tid][g] = row[g];
        }
    }
    __syncthreads();
    uint64_t thr = s_thr;

    for (int gi = 0; gi < gcount; ++gi) {
    const int tile = tile0 + gi;
    const int64_t t0 = (int64_t)tile * BM25_TILE_DOCS;
    const int tile_n = (int)min((int64_t)BM25_TILE_DOCS, n_rows - t0);

    // posting range of each term of the chunk [c0, c0+nt) inside this tile
    auto resolve = [&](int c0, int nt) {
        __syncthreads();
        if (tid < nt) {
            // (slot, base, rare length) of every query term were resolved once per batch by bm25_resolve_kernel
            const int32_t sl = q_slot[c0 + tid];
            int64_t lo = 0; int len = 0;
            if (sl >= 0) {
                const uint32_t* row = tile_off + (int64_t)sl * (n_tiles_idx + 1);
                const uint32_t o0 = row[tile], o1 = row[tile + 1];
                lo = q_base[c0 + tid] + o0; len = (int)(o1 - o0);
            } else if (sl == -1) {
                lo = q_base[c0 + tid]; len = q_rare_len[c0 + tid];   // rare: whole list, filtered by doc range below
            }
            s_lo[tid] = lo; s_len[tid] = len;
        }
        if (tid == 0) { s_next_term = 0; s_next_off = 0; }
        __syncthreads();
    };
    // next pass of at most BQ_MAX_SLABS slabs of the resolved chunk; returns the slab count
    auto next_pass = [&](int nt) -> int {
        if (tid == 0) {
            int ns = 0, j = s_next_term, o = s_next_off;
            while (j < nt && ns < BQ_MAX_SLABS) {
                const int rem = s_len[j] - o;
                if (rem <= 0) { ++j; o = 0; continue; }
                s_slab_lo[ns] = s_lo[j] + o; s_slab_n[ns] = rem < BQ_THREADS ? rem : BQ_THREADS; ++ns;
                o += BQ_THREADS;
            }
            s_nslab = ns; s_next_term = j; s_next_off = o;
        }
        __syncthreads();
        return s_nslab;
    };
    // acc[doc] += score, slab after slab (query-term order), BQ_PREFETCH slabs in registers at a time
    auto accumulate = [&](int ns) {
        for (int s0 = 0; s0 < ns; s0 += BQ_PREFETCH) {
            uint32_t d[BQ_PREFETCH]; float sc[BQ_PREFETCH];
#pragma unroll
            for (int u = 0; u < BQ_PREFETCH; ++u) {
                d[u] = 0xFFFFFFFFu; sc[u] = 0.f;
                if (s0 + u < ns && tid < s_slab_n[s0 + u]) {
                    const int64_t p = s_slab_lo[s0 + u] + tid;
                    d[u] = post_doc[p]; sc[u] = post_score[p];
                }
            }
#pragma unroll
            for (int u = 0; u < BQ_PREFETCH; ++u) {
                if (s0 + u < ns) {
                    const uint32_t rel = d[u] - (uint32_t)t0;     // wraps to a huge value for docs below t0
                    if (d[u] != 0xFFFFFFFFu && rel < (uint32_t)tile_n) acc[rel] += sc[u];
                    __syncthreads();
                }
            }
        }
    };
    // every touched document is pushed exactly once with its final score (first claimer takes it)
    auto claim = [&](int ns) {
        for (int s0 = 0; s0 < ns; s0 += BQ_PREFETCH) {
            uint32_t d[BQ_PREFETCH];
#pragma unroll
            for (int u = 0; u < BQ_PREFETCH; ++u) {
                d[u] = 0xFFFFFFFFu;
                if (s0 + u < ns && tid < s_slab_n[s0 + u]) d[u] = post_doc[s_slab_lo[s0 + u] + tid];
            }
#pragma unroll
            for (int u = 0; u < BQ_PREFETCH; ++u) {
                if (s0 + u < ns) {
                    const uint32_t rel = d[u] - (uint32_t)t0;
                    if (d[u] != 0xFFFFFFFFu && rel < (uint32_t)tile_n) {
                        const float sum = atomicExch(&acc[rel], 0.f);
                        if (sum > 0.f && (alive == nullptr || bit_test(alive, d[u])))
                            select_push(sel, make_key_desc(sum, ord_base + d[u]), thr);
                    }
                    __syncthreads();
                    if (s_count + BQ_THREADS > cap) select_prune<BQ_THREADS>(sel, P, tid, 0);
                    thr = s_thr;
                }
            }
        }
    };

    bool done = false;
    if (single_chunk) {
        // common case: the whole query resolves in one chunk (ranges come from the per-item shared-memory copy);
        // if it also fits one pass, accumulate and claim without resolving twice
        __syncthreads();
        if (tid < te - tb) {
            const int32_t sl = s_slot[tid];
            int64_t lo = 0; int len = 0;
            if (sl >= 0) { lo = s_base[tid] + s_toff[tid][gi]; len = (int)(s_toff[tid][gi + 1] - s_toff[tid][gi]); }
            else if (sl == -1) { lo = s_base[tid]; len = s_rare[tid]; }
            s_lo[tid] = lo; s_len[tid] = len;
        }
        if (tid == 0) { s_next_term = 0; s_next_off = 0; }
        __syncthreads();
        const int ns = next_pass(te - tb);
        const bool more = (s_next_term < te - tb);   // uniform: written before the barrier inside next_pass
        if (!more && ns <= BQ_PREFETCH) {
            // all slabs of the item fit the register window: accumulate and claim from the same registers
            uint32_t d[BQ_PREFETCH]; float sc[BQ_PREFETCH];
#pragma unroll
            for (int u = 0; u < BQ_PREFETCH; ++u) {
                d[u] = 0xFFFFFFFFu; sc[u] = 0.f;
                if (u < ns && tid < s_slab_n[u]) {
                    const int64_t p = s_slab_lo[u] + tid;
                    d[u] = post_doc[p]; sc[u] = post_score[p];
                }
            }
#pragma unroll
            for (int u = 0; u < BQ_PREFETCH; ++u) {
                if (u < ns) {
                    const uint32_t rel = d[u] - (uint32_t)t0;
                    if (d[u] != 0xFFFFFFFFu && rel < (uint32_t)tile_n) acc[rel] += sc[u];
                    __syncthreads();
                }
            }
#pragma unroll
            for (int u = 0; u < BQ_PREFETCH; ++u) {
                if (u < ns) {
                    const uint32_t rel = d[u] - (uint32_t)t0;
                    if (d[u] != 0xFFFFFFFFu && rel < (uint32_t)tile_n) {
                        const float sum = atomicExch(&acc[rel], 0.f);
                        if (sum > 0.f && (alive == nullptr || bit_test(alive, d[u])))
                            select_push(sel, make_key_desc(sum, ord_base + d[u]), thr);
                    }
                    __syncthreads();
                    if (s_count + BQ_THREADS > cap) select_prune<BQ_THREADS>(sel, P, tid, 0);
                    thr = s_thr;
                }
            }
            done = true;
        } else if (!more) {
            accumulate(ns);
            claim(ns);
            done = true;
        }
    }
    if (!done) {
        // general case: ALL terms are accumulated before any document is claimed
        for (int sweep = 0; sweep < 2; ++sweep) {
            for (int c0 = tb; c0 < te; c0 += BQ_MAX_TERMS) {
                const int nt = min(BQ_MAX_TERMS, te - c0);
                resolve(c0, nt);
                for (int ns = next_pass(nt); ns > 0; ns = next_pass(nt)) {
                    if (sweep == 0) accumulate(ns); else claim(ns);
                    __syncthreads();
                }
            }
        }
    }
    }   // tiles of the group
    select_prune<BQ_THREADS>(sel, P, tid, 0);
    select_store<BQ_THREADS>(sel, P, part + ((size_t)qi * n_groups + tg) * P, tid);
    if (tid == 0 && s_count == P) atomicMin(&g_thr[qi], (unsigned long long)sbuf[P - 1]);
    }   // work items
}

// once per batch: (tile-index slot, posting base, rare length) of every query term position
__global__ void bm25_resolve_kernel(const uint32_t* __restrict__ q_terms, int n_terms, const int64_t* __restrict__ post_off,
                                    const int32_t* __restrict__ tile_slot, int64_t vocab, int32_t* __restrict__ q_slot,
                                    int64_t* __restrict__ q_base, int32_t* __restrict__ q_rare_len)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_terms) return;
    const uint32_t t = q_terms[i];
    int32_t sl = -2; int64_t b = 0; int32_t rl = 0;
    if ((int64_t)t < vocab) {
        b = post_off[t];
        sl = tile_slot[t];
        if (sl < 0) { sl = -1; rl = (int32_t)(post_off[t + 1] - b); }
    }
    q_slot[i] = sl; q_base[i] = b; q_rare_len[i] = rl;
}

static int bq_cap(int P) { return P <= 256 ? 512 : (P <= 512 ? 1024 : 2048); }   // cap - P >= BQ_THREADS; small caps keep 3 CTAs per SM

size_t bm25_part_elems(int64_t n_rows, int batch, int P)
{
    int64_t n_tiles = (n_rows + BM25_TILE_DOCS - 1) / BM25_TILE_DOCS;
    if (n_tiles < 1) n_tiles = 1;
    const int64_t n_groups = (n_tiles + bq_group() - 1) / bq_group();
    return (size_t)n_groups * batch * P + (size_t)batch;   // + per-query threshold hints
}

void launch_bm25(const DeviceInfo& di, const Postings& post, int64_t n_rows, const uint32_t* alive,
                 const uint32_t* q_terms, const int32_t* q_term_offsets, int n_terms_total, void* resolve_ws, int batch, int P,
                 uint32_t ord_base, uint64_t* part, uint64_t* keys_out, cudaStream_t st)
{
    // resolve_ws: >= n_terms_total * 16 bytes
    int64_t* q_base = reinterpret_cast<int64_t*>(resolve_ws);
    int32_t* q_slot = reinterpret_cast<int32_t*>(q_base + (n_terms_total > 0 ? n_terms_total : 1));
    int32_t* q_rare = q_slot + (n_terms_total > 0 ? n_terms_total : 1);
    if (n_terms_total > 0) {
        bm25_resolve_kernel<<<(n_terms_total + 255) / 256, 256, 0, st>>>(q_terms, n_terms_total, post.off, post.tile_slot, post.vocab,
                                                                         q_slot, q_base, q_rare);
        KRAG_CUDA(cudaGetLastError());
        count_launch();
    }
    int64_t n_tiles = (n_rows + BM25_TILE_DOCS - 1) / BM25_TILE_DOCS;
    if (n_tiles < 1) n_tiles = 1;
    const int cap = bq_cap(P);
    const size_t smem = (size_t)BM25_TILE_DOCS * 4 + (size_t)cap * 8;
    static bool attr_set = false;
    if (!attr_set) {
        KRAG_CUDA(cudaFuncSetAttribute(bm25_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr_set = true;
    }
    const int64_t n_groups = (n_tiles + bq_group() - 1) / bq_group();
    unsigned long long* g_thr = reinterpret_cast<unsigned long long*>(part + (size_t)n_groups * batch * P);
    KRAG_CUDA(cudaMemsetAsync(g_thr, 0xFF, sizeof(unsigned long long) * (size_t)batch, st));
    const int64_t n_items = n_groups * batch;
    const int per_sm = (smem <= 74 * 1024) ? 3 : 2;
    const int64_t max_grid = (int64_t)per_sm * di.sm_count;
    const int grid = (int)(n_items < max_grid ? n_items : max_grid);
    bm25_tile_kernel<<<grid, BQ_THREADS, smem, st>>>(post.off, post.doc, post.score, post.tile_slot, post.tile_off,
                                                     post.n_tiles, post.vocab, q_terms, q_term_offsets, q_slot, q_base, q_rare, n_rows,
                                                     alive, P, cap, ord_base, batch, (int)n_tiles, bq_group(), part, g_thr);
    KRAG_CUDA(cudaGetLastError());
    count_launch();
    launch_merge(part, (int)n_groups, P, batch, P, /*list_stride=*/P, /*batch_stride=*/n_groups * P, keys_out, st,
                 reinterpret_cast<const uint64_t*>(g_thr));
    launch_bm25_fill(keys_out, batch, P, alive, n_rows, ord_base, st);
}

}  // namespace krag
