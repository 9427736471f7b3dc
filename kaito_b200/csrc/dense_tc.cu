// dense_tc.cu -- K2: batched dense search on the 5th-gen tensor cores.
//
// Replaces faiss IndexFlatL2.search for query batches (its nq >= 20 sgemm path; reference
// call site presets/ragengine/vector_store/faiss_store.py:44-49 via hybrid_retriever.py:209-213).
//
//   1. dense_tc_kernel<NQ, DUMP>  (tcgen05.mma kind::tf32, TMA-staged, accumulators in TMEM)
//        D[128 corpus rows x NQ queries] = X_tile . Q^T, fp32 corpus fed as TF32 straight
//        from HBM (no conversion pass; the corpus is streamed exactly once per pass).
//        epilogue: a = |x|^2 - 2 x.q  per (row, query), compared with a per-query threshold;
//        rows below it are appended to that query's candidate list (rare: ~C of N rows).
//        DUMP mode writes `a` for a strided 1/64 sample of the tiles instead.
//   2. sample_threshold_kernel: per query, the m-th smallest sampled `a` -> threshold that
//        admits ~C = max(512, 4P) rows of the full corpus.
//   3. rescore_kernel: EXACT fp32 squared L2 of every candidate in K1's operation order
//        (bit-identical to dense_scan.cu / the oracle), then the ordinary top-P merge.
//   4. certify_kernel: a query's result is exact if every non-candidate row is provably
//        farther than the P-th exact distance:  D_P <= thr + |q|^2 - 2 eps,  eps = worst-case
//        TF32 dot-product error (2^-9 + 2^-12) |x|_max |q|.  Queries that fail (candidate
//        overflow, adversarial near-duplicates) are re-run on the exact scan kernel K1.
//
// So the tensor cores only PRUNE; every returned score is an exact fp32 distance.
//
// Roofline: bytes = n_rows*dpad*4 per pass (HBM), flops = 2*NQ*n_rows*dpad (TF32 pipe).
// At NQ = 256 the kernel sits on the HBM/TF32 ridge (128 flop/byte).
#include <cuda.h>
#include <cuda_bf16.h>
#include <math_constants.h>

#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "engine.h"
#include "select.cuh"
#include "tc_ptx.cuh"

namespace krag {

// ------------------------------------------------------------------------- the kernel
constexpr int TC_TILE_M = 128;          // corpus rows per tile (UMMA M)
constexpr int TC_KBLOCK = 32;           // fp32 elements per k-block = one 128-byte swizzle row
constexpr int TC_THREADS = 352;         // warp 0: corpus TMA, warp 1: MMA issuer, warp 2: query TMA, warps 3-10: epilogue
constexpr int TC_SAMPLE_STRIDE = 64;    // DUMP mode visits every 64th tile
constexpr int TC_A_BYTES = TC_TILE_M * TC_KBLOCK * 4;  // 16 KB

// Two independent shared-memory rings.  The corpus ring is deep because its slabs come from
// HBM (latency ~2 us under load: ~100 KB must be in flight per SM to sustain 6.5 TB/s); the
// query ring is shallow because its slabs are L2 hits re-read for every tile.
template <int NQ> struct TcCfg {
    static constexpr int Q_BYTES = NQ * TC_KBLOCK * 4;
    static constexpr int Q_STAGES = (NQ == 256) ? 3 : 4;
    static constexpr int A_STAGES = (NQ == 256) ? 7 : (NQ == 128 ? 9 : 11);
    static constexpr int TMEM_COLS = (2 * NQ < 32) ? 32 : 2 * NQ;  // double-buffered accumulators
    static constexpr int BAR_BYTES = 512;
    static constexpr size_t SMEM = (size_t)A_STAGES * TC_A_BYTES + (size_t)Q_STAGES * Q_BYTES + 1024 /*align*/ + BAR_BYTES + NQ * 4;
};

template <int NQ, bool DUMP>
__global__ void __launch_bounds__(TC_THREADS, 1)
dense_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmQ, int64_t n_rows,
                int kblocks, int64_t n_ltiles, int tile_stride, const float* __restrict__ xnorm,
                const uint32_t* __restrict__ alive, const float* __restrict__ thr_g, uint32_t* __restrict__ cand_count,
                uint32_t* __restrict__ cand_rows, int cap, float* __restrict__ dump, int64_t S, int dump_min)
{
    using Cfg = TcCfg<NQ>;
    extern __shared__ unsigned char tc_smem_raw[];
    unsigned char* smem_a = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* smem_q = smem_a + (size_t)Cfg::A_STAGES * TC_A_BYTES;
    unsigned char* tail = smem_q + (size_t)Cfg::Q_STAGES * Cfg::Q_BYTES;
    uint64_t* afull = reinterpret_cast<uint64_t*>(tail);               // [A_STAGES]
    uint64_t* aempty = afull + Cfg::A_STAGES;                           // [A_STAGES]
    uint64_t* qfull = aempty + Cfg::A_STAGES;                           // [Q_STAGES]
    uint64_t* qempty = qfull + Cfg::Q_STAGES;                           // [Q_STAGES]
    uint64_t* tfull_bar = qempty + Cfg::Q_STAGES;                       // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                               // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* s_thr = reinterpret_cast<float*>(tail + Cfg::BAR_BYTES);     // [NQ]
    static_assert((2 * (Cfg::A_STAGES + Cfg::Q_STAGES) + 4) * 8 + 8 <= Cfg::BAR_BYTES, "barrier area too small");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmQ);
        for (int s = 0; s < Cfg::A_STAGES; ++s) { mbar_init(&afull[s], 1); mbar_init(&aempty[s], 1); }
        for (int s = 0; s < Cfg::Q_STAGES; ++s) { mbar_init(&qfull[s], 1); mbar_init(&qempty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 8); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // s_thr[j] = thr_j / 2:  a = |x|^2 - 2 x.q < thr_j   <=>   x.q + thr_j/2 > |x|^2/2
    if (!DUMP) for (int j = threadIdx.x; j < NQ; j += TC_THREADS) s_thr[j] = 0.5f * thr_g[j];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== corpus producer (HBM stream, read once) =====================
        int stage = 0; uint32_t phase = 0;
        for (int64_t lt = blockIdx.x; lt < n_ltiles; lt += gridDim.x) {
            const int row0 = (int)(lt * tile_stride * TC_TILE_M);
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&aempty[stage], phase ^ 1);
                if (lane == 0) {
                    mbar_expect_tx(&afull[stage], TC_A_BYTES);
                    tma_load_2d(smem_a + (size_t)stage * TC_A_BYTES, &tmX, &afull[stage], kb * TC_KBLOCK, row0, TMA_EVICT_FIRST);
                }
                __syncwarp();
                if (++stage == Cfg::A_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 2) {
        // ===================== query producer (L2-resident, re-read per tile) =====================
        int stage = 0; uint32_t phase = 0;
        for (int64_t lt = blockIdx.x; lt < n_ltiles; lt += gridDim.x) {
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&qempty[stage], phase ^ 1);
                if (lane == 0) {
                    mbar_expect_tx(&qfull[stage], Cfg::Q_BYTES);
                    tma_load_2d(smem_q + (size_t)stage * Cfg::Q_BYTES, &tmQ, &qfull[stage], kb * TC_KBLOCK, 0, TMA_EVICT_LAST);
                }
                __syncwarp();
                if (++stage == Cfg::Q_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = umma_idesc_tf32(TC_TILE_M, NQ);
        int sa = 0, sq = 0; uint32_t pa = 0, pq = 0; int acc = 0; uint32_t acc_phase = 0;
        for (int64_t lt = blockIdx.x; lt < n_ltiles; lt += gridDim.x) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NQ);
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&qfull[sq], pq);
                mbar_wait(&afull[sa], pa);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_base = smem_u32(smem_a + (size_t)sa * TC_A_BYTES);
                    const uint32_t b_base = smem_u32(smem_q + (size_t)sq * Cfg::Q_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_KBLOCK / 8; ++k) {   // UMMA K = 8 for tf32 (32 bytes)
                        umma_tf32(d_tmem, umma_desc_sw128(a_base + k * 32), umma_desc_sw128(b_base + k * 32), idesc,
                                  (uint32_t)((kb | k) != 0));
                    }
                    umma_commit(&aempty[sa]);                // slots are free once these MMAs retire
                    umma_commit(&qempty[sq]);
                    if (kb == kblocks - 1) umma_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                if (++sa == Cfg::A_STAGES) { sa = 0; pa ^= 1; }
                if (++sq == Cfg::Q_STAGES) { sq = 0; pq ^= 1; }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    } else {
        // ===================== epilogue: 8 warps, two per TMEM lane group, half of the columns each =====
        // (the epilogue paces the kernel when it is slower than the 96 MMAs of a tile, so it is kept to
        //  ~2 instructions per (row, query) pair: one add and one 3-input max; hits are ~1 in 20000)
        const int lg = warp & 3;                         // TMEM lane group this warp may access
        const int col_half = (warp - 3) >> 2;            // 0: columns [0, NQ/2), 1: [NQ/2, NQ)
        const int row_in_tile = lg * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        for (int64_t lt = blockIdx.x; lt < n_ltiles; lt += gridDim.x) {
            const int64_t row = lt * tile_stride * TC_TILE_M + row_in_tile;
            float xn = CUDART_INF_F;
            if (row < n_rows && (alive == nullptr || bit_test(alive, (uint32_t)row))) xn = xnorm[row];
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * NQ);
            const float half_xn = 0.5f * xn;
#pragma unroll 1
            for (int c0 = col_half * (NQ / 2); c0 < (col_half + 1) * (NQ / 2); c0 += 32) {
                uint32_t v[32];
                tmem_ld_x32(taddr + c0, v);
                tmem_wait_ld();
                if (DUMP) {
                    if (dump_min) {
                        // threshold sampling: per query only the MINIMUM over this warp's 32 rows is kept (one REDUX per
                        // column) -- the few smallest values of the sample, which is all the threshold needs, survive
                        // (two of them share a 32-row group with probability ~m^2 / (2 groups)); 40x less dump traffic
                        uint32_t mine = 0xFFFFFFFFu;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const uint32_t mn = __reduce_min_sync(0xffffffffu, f32_ordered_bits(fmaf(-2.f, __uint_as_float(v[j]), xn)));
                            if (lane == j) mine = mn;
                        }
                        dump[(int64_t)(c0 + lane) * S + lt * (TC_TILE_M / 32) + lg] = f32_from_ordered_bits(mine);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            dump[(int64_t)(c0 + j) * S + lt * TC_TILE_M + row_in_tile] = fmaf(-2.f, __uint_as_float(v[j]), xn);
                    }
                } else {
                    float t[32];
                    float mx = -CUDART_INF_F;
#pragma unroll
                    for (int j4 = 0; j4 < 32; j4 += 4) {
                        const float4 h = *reinterpret_cast<const float4*>(s_thr + c0 + j4);   // warp-uniform: broadcast
                        t[j4 + 0] = __uint_as_float(v[j4 + 0]) + h.x;
                        t[j4 + 1] = __uint_as_float(v[j4 + 1]) + h.y;
                        t[j4 + 2] = __uint_as_float(v[j4 + 2]) + h.z;
                        t[j4 + 3] = __uint_as_float(v[j4 + 3]) + h.w;
                        mx = fmaxf(mx, fmaxf(fmaxf(t[j4 + 0], t[j4 + 1]), fmaxf(t[j4 + 2], t[j4 + 3])));
                    }
                    if (mx > half_xn) {   // rare: some pair in this chunk passes  x.q + thr/2 > |x|^2/2
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (t[j] > half_xn) {
                                const uint32_t pos = atomicAdd(&cand_count[c0 + j], 1u);
                                if (pos < (uint32_t)cap) cand_rows[(size_t)(c0 + j) * cap + pos] = (uint32_t)row;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------- 2-CTA main pass (cta_group::2)
// A CTA pair (one cluster of 2, one CTA per SM of a TPC) computes D[256 rows x NQ] per pair-tile with
// tcgen05.mma.cta_group::2 (M = 256): each CTA stages ITS 128 corpus rows and HALF of the query slab,
// the pair's tensor cores share the operands.  Per 128-cycle MMA each SM now reads 8 KB of operands
// (A 4 KB + B half 4 KB) and TMA writes 32 KB per 512 cycles: 64 + 64 B/clk against the 128 B/clk
// shared-memory port, versus 96 + 96 for the 1-CTA kernel (which caps its tensor pipe at 67%).
// The halved query slab also doubles the query ring depth (6 stages) and halves L2->SM traffic.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N)
{
    return (1u << 4) /* D f32 */ | (1u << 7) /* A bf16 */ | (1u << 10) /* B bf16 */ | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <bool BF16>
__device__ __forceinline__ void umma_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    if (BF16) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        umma_tf32_2sm(d_tmem, adesc, bdesc, idesc, accumulate);
    }
}
// fp32 -> bf16 (round to nearest even) shadow rows; n4 = number of float4 groups
__global__ void f32_to_bf16_kernel(const float4* __restrict__ in, uint2* __restrict__ out, int64_t n4)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = in[i];
        const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&a); o.y = *reinterpret_cast<const uint32_t*>(&b);
        out[i] = o;
    }
}
void launch_f32_to_bf16(const float* in, uint16_t* out, int64_t n_elems, cudaStream_t st)
{
    if (n_elems == 0) return;
    const int64_t n4 = n_elems / 4;
    int64_t blocks = (n4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    f32_to_bf16_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<uint2*>(out), n4);
    KRAG_CUDA(cudaGetLastError());
    count_launch();
}

// One ring; a stage holds TC2_KB_PER_STAGE k-blocks of this CTA's corpus rows and of its half of the
// query rows, so the single MMA-issuing thread pays one barrier wait and one commit per 8 MMAs (it was
// the bottleneck at 2 waits + 2 commits per 4 MMAs: ~770 cycles per k-block against 512 cycles of MMA work).
constexpr int TC2_KB_PER_STAGE = 2;
template <int NQ> struct Tc2Cfg {
    static constexpr int QH_BYTES = (NQ / 2) * TC_KBLOCK * 4;          // this CTA's half of one query k-block
    static constexpr int STAGE_BYTES = TC2_KB_PER_STAGE * (TC_A_BYTES + QH_BYTES);
    static constexpr int STAGES = (200 * 1024) / STAGE_BYTES;          // 3 at NQ = 256, 4 at NQ = 128, 5 at NQ = 64
    static constexpr int TMEM_COLS = (2 * NQ < 32) ? 32 : 2 * NQ;
    static constexpr int BAR_BYTES = 256;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 + BAR_BYTES + NQ * 4;
};

__device__ __forceinline__ uint64_t desc_with_lo(uint64_t base_desc, uint32_t lo)
{
    return (base_desc & 0xFFFFFFFF00000000ull) | (uint64_t)lo;
}

// BF16 = true: the operands are a bf16 SHADOW of the corpus (and of the queries): k-blocks are still 128 bytes per row
// (64 elements), the MMA is kind::f16 with K = 16 -- half the HBM bytes and half the tensor time of the TF32 pass.
// Only the pruning changes; the rescoring that produces the returned distances always reads the fp32 corpus.
template <int NQ, bool BF16>
__global__ void __launch_bounds__(TC_THREADS, 1)
dense_tc2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmQh, int64_t n_rows,
                 int kblocks, int64_t n_ptiles, const float* __restrict__ xnorm, const uint32_t* __restrict__ alive,
                 const float* __restrict__ thr_g, uint32_t* __restrict__ cand_count, uint32_t* __restrict__ cand_rows, int cap)
{
    using Cfg = Tc2Cfg<NQ>;
    extern __shared__ unsigned char tc_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* tail = smem + (size_t)Cfg::STAGES * Cfg::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);             // used in the leader only (2 arrivals + tx of both CTAs)
    uint64_t* empty_bar = full_bar + Cfg::STAGES;                        // per CTA
    uint64_t* tfull_bar = empty_bar + Cfg::STAGES;                       // per CTA   [2]
    uint64_t* tempty_bar = tfull_bar + 2;                                // leader only [2], 16 arrivals (8 warps x 2 CTAs)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* s_thr = reinterpret_cast<float*>(tail + Cfg::BAR_BYTES);
    static_assert((2 * Cfg::STAGES + 4) * 8 + 8 <= Cfg::BAR_BYTES, "barrier area too small");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();          // 0 = pair leader
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int n_stages_per_tile = (kblocks + TC2_KB_PER_STAGE - 1) / TC2_KB_PER_STAGE;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmQh);
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 16); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int j = threadIdx.x; j < NQ; j += TC_THREADS) s_thr[j] = 0.5f * thr_g[j];
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                // both CTAs' barriers are initialised before any remote signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== producer: this CTA's 128 corpus rows + its half of the query rows =====================
        int stage = 0; uint32_t phase = 0;
        for (int64_t pt = pair; pt < n_ptiles; pt += n_pairs) {
            const int row0 = (int)(pt * (2 * TC_TILE_M) + rank * TC_TILE_M);
            for (int sk = 0; sk < n_stages_per_tile; ++sk) {
                const int nkb = min(TC2_KB_PER_STAGE, kblocks - sk * TC2_KB_PER_STAGE);
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (lane == 0) {
                    unsigned char* sa = smem + (size_t)stage * Cfg::STAGE_BYTES;
                    unsigned char* sq = sa + TC2_KB_PER_STAGE * TC_A_BYTES;
                    if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * nkb * (TC_A_BYTES + Cfg::QH_BYTES));   // both CTAs' bytes
                    else mbar_arrive_remote(&full_bar[stage], 0);
                    for (int u = 0; u < nkb; ++u) {
                        const int kb = sk * TC2_KB_PER_STAGE + u;
                        constexpr int KBE = BF16 ? 2 * TC_KBLOCK : TC_KBLOCK;   // elements per 128-byte k-block
                        tma_load_2d_2sm(sa + (size_t)u * TC_A_BYTES, &tmX, &full_bar[stage], kb * KBE, row0, TMA_EVICT_FIRST);
                        tma_load_2d_2sm(sq + (size_t)u * Cfg::QH_BYTES, &tmQh, &full_bar[stage], kb * KBE, (int)rank * (NQ / 2),
                                        TMA_EVICT_LAST);
                    }
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // ===================== MMA issuer (pair leader only) =====================
            constexpr uint32_t idesc = BF16 ? umma_idesc_bf16(2 * TC_TILE_M, NQ) : umma_idesc_tf32(2 * TC_TILE_M, NQ);
            const uint64_t desc0 = umma_desc_sw128(smem_u32(smem));      // descriptor of ring byte 0; only the low word moves
            const uint32_t lo0 = (uint32_t)desc0;
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
            for (int64_t pt = pair; pt < n_ptiles; pt += n_pairs) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NQ);
                for (int sk = 0; sk < n_stages_per_tile; ++sk) {
                    const int nkb = min(TC2_KB_PER_STAGE, kblocks - sk * TC2_KB_PER_STAGE);
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t a_lo = lo0 + (uint32_t)((stage * Cfg::STAGE_BYTES) >> 4);
                        const uint32_t q_lo = a_lo + (uint32_t)((TC2_KB_PER_STAGE * TC_A_BYTES) >> 4);
#pragma unroll
                        for (int u = 0; u < TC2_KB_PER_STAGE; ++u) {
                            if (u < nkb) {
#pragma unroll
                                for (int k = 0; k < TC_KBLOCK / 8; ++k)   // UMMA K = 8 (32 bytes = 2 descriptor units)
                                    umma_2sm<BF16>(d_tmem, desc_with_lo(desc0, a_lo + (uint32_t)((u * TC_A_BYTES) >> 4) + 2 * k),
                                                  desc_with_lo(desc0, q_lo + (uint32_t)((u * Cfg::QH_BYTES) >> 4) + 2 * k), idesc,
                                                  (uint32_t)((sk | u | k) != 0));
                            }
                        }
                        umma_commit_2sm(&empty_bar[stage]);            // frees the stage in both CTAs
                        if (sk == n_stages_per_tile - 1) umma_commit_2sm(&tfull_bar[acc]);
                    }
                    __syncwarp();
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (warp >= 3) {
        // ===================== epilogue: 8 warps per CTA over this CTA's 128 TMEM lanes =====================
        const int lg = warp & 3;
        const int col_half = (warp - 3) >> 2;
        const int row_in_tile = lg * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        for (int64_t pt = pair; pt < n_ptiles; pt += n_pairs) {
            const int64_t row = pt * (2 * TC_TILE_M) + rank * TC_TILE_M + row_in_tile;
            float xn = CUDART_INF_F;
            if (row < n_rows && (alive == nullptr || bit_test(alive, (uint32_t)row))) xn = xnorm[row];
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * NQ);
            const float half_xn = 0.5f * xn;
#pragma unroll 1
            for (int c0 = col_half * (NQ / 2); c0 < (col_half + 1) * (NQ / 2); c0 += 32) {
                uint32_t v[32];
                tmem_ld_x32(taddr + c0, v);
                tmem_wait_ld();
                float t[32];
                float mx = -CUDART_INF_F;
#pragma unroll
                for (int j4 = 0; j4 < 32; j4 += 4) {
                    const float4 h = *reinterpret_cast<const float4*>(s_thr + c0 + j4);
                    t[j4 + 0] = __uint_as_float(v[j4 + 0]) + h.x;
                    t[j4 + 1] = __uint_as_float(v[j4 + 1]) + h.y;
                    t[j4 + 2] = __uint_as_float(v[j4 + 2]) + h.z;
                    t[j4 + 3] = __uint_as_float(v[j4 + 3]) + h.w;
                    mx = fmaxf(mx, fmaxf(fmaxf(t[j4 + 0], t[j4 + 1]), fmaxf(t[j4 + 2], t[j4 + 3])));
                }
                if (mx > half_xn) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (t[j] > half_xn) {
                            const uint32_t pos = atomicAdd(&cand_count[c0 + j], 1u);
                            if (pos < (uint32_t)cap) cand_rows[(size_t)(c0 + j) * cap + pos] = (uint32_t)row;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) mbar_arrive(&tempty_bar[acc]);
                else mbar_arrive_remote(&tempty_bar[acc], 0);
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                // the peer may still be reading operands from this CTA's smem
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------- K2, default: fp32 rows -> bf16 IN SHARED MEMORY -> kind::f16
// The TF32 pass above runs at 97-99% tensor-pipe activity with the SM clock held at 1.2-1.3 GHz by the power cap:
// it is bound by tensor energy, not by HBM.  This variant keeps the corpus fp32 in HBM (no shadow copy, same
// 4 bytes/element of traffic) but halves the tensor work: TMA lands the fp32 k-blocks in a staging ring, four
// converter warps round them to bf16 (cvt.rn.bf16x2, the same rounding as the shadow mode) straight into the
// 128B-swizzled operand ring, and the pair's MMAs are kind::f16 (K = 16).  The pass becomes HBM-bound.
// Pruning only, as before: candidates are re-scored in exact fp32 and certified with the bf16 error bound.
//   warp 0: corpus TMA (fp32 staging ring, per CTA)      warp 1: MMA issuer (pair leader)
//   warp 2: query TMA (bf16 queries, operand ring)       warps 3-10: epilogue      warps 11-13: converters, warp w owns
//   operand stage w: the three conversions in flight hide each other's shared-memory and barrier latency
constexpr int TCV_THREADS = 448;
constexpr int TCV_A_STAGES = 8;            // fp32 staging k-blocks (16 KB each): 128 KB of corpus in flight per SM covers the HBM latency
constexpr int TCV_B_STAGES = 3;            // bf16 operand k-blocks: 16 KB corpus + this CTA's half of the queries
template <int NQ> struct TcvCfg {
    static constexpr int QH_BYTES = (NQ / 2) * 128;                    // 64 bf16 per row
    static constexpr int B_STAGE = TC_A_BYTES + QH_BYTES;
    static constexpr int TMEM_COLS = (2 * NQ < 32) ? 32 : 2 * NQ;
    static constexpr int BAR_BYTES = 256;
    static constexpr size_t SMEM = (size_t)TCV_A_STAGES * TC_A_BYTES + (size_t)TCV_B_STAGES * B_STAGE + 1024 + BAR_BYTES + NQ * 4;
};

template <int NQ>
__global__ void __launch_bounds__(TCV_THREADS, 1)
dense_tc2cvt_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmQh, int64_t n_rows,
                    int kblocks /*fp32 k-blocks, even*/, int64_t n_ptiles, const float* __restrict__ xnorm,
                    const uint32_t* __restrict__ alive, const float* __restrict__ thr_g, uint32_t* __restrict__ cand_count,
                    uint32_t* __restrict__ cand_rows, int cap)
{
    using Cfg = TcvCfg<NQ>;
    extern __shared__ unsigned char tc_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* ring_a = smem;                                            // fp32 staging
    unsigned char* ring_b = smem + (size_t)TCV_A_STAGES * TC_A_BYTES;        // bf16 operands
    unsigned char* tail = ring_b + (size_t)TCV_B_STAGES * Cfg::B_STAGE;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(tail);                    // per CTA
    uint64_t* a_empty = a_full + TCV_A_STAGES;                               // per CTA, the converter warp that read the slot
    uint64_t* b_full = a_empty + TCV_A_STAGES;                               // leader only: 2 query producers + 2 converter warps + tx
    uint64_t* b_empty = b_full + TCV_B_STAGES;                               // per CTA (multicast commit)
    uint64_t* tfull_bar = b_empty + TCV_B_STAGES;                            // per CTA [2]
    uint64_t* tempty_bar = tfull_bar + 2;                                    // leader only [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* s_thr = reinterpret_cast<float*>(tail + Cfg::BAR_BYTES);
    static_assert((2 * TCV_A_STAGES + 2 * TCV_B_STAGES + 4) * 8 + 8 <= Cfg::BAR_BYTES, "barrier area too small");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int kb2 = kblocks >> 1;                                            // bf16 k-blocks per tile

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmQh);
        for (int s = 0; s < TCV_A_STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < TCV_B_STAGES; ++s) { mbar_init(&b_full[s], 4); mbar_init(&b_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 16); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int j = threadIdx.x; j < NQ; j += TCV_THREADS) s_thr[j] = 0.5f * thr_g[j];
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== corpus producer: fp32 k-blocks of this CTA's 128 rows =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t pt = pair; pt < n_ptiles; pt += n_pairs) {
                const int row0 = (int)(pt * (2 * TC_TILE_M) + rank * TC_TILE_M);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&a_empty[stage], phase ^ 1);
                    mbar_expect_tx(&a_full[stage], TC_A_BYTES);
                    tma_load_2d(ring_a + (size_t)stage * TC_A_BYTES, &tmX, &a_full[stage], kb * TC_KBLOCK, row0, TMA_EVICT_FIRST);
                    if (++stage == TCV_A_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 2) {
        // ===================== query producer: this CTA's half of the bf16 query rows =====================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t pt = pair; pt < n_ptiles; pt += n_pairs) {
                for (int jb = 0; jb < kb2; ++jb) {
                    mbar_wait(&b_empty[stage], phase ^ 1);
                    if (rank == 0) mbar_expect_tx(&b_full[stage], 2 * Cfg::QH_BYTES);
                    else mbar_arrive_remote(&b_full[stage], 0);
                    tma_load_2d_2sm(ring_b + (size_t)stage * Cfg::B_STAGE + TC_A_BYTES, &tmQh, &b_full[stage], jb * 64, (int)rank * (NQ / 2),
                                    TMA_EVICT_LAST);
                    if (++stage == TCV_B_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // ===================== MMA issuer (pair leader) =====================
            constexpr uint32_t idesc = umma_idesc_bf16(2 * TC_TILE_M, NQ);
            const uint64_t desc0 = umma_desc_sw128(smem_u32(ring_b));
            const uint32_t lo0 = (uint32_t)desc0;
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
            for (int64_t pt = pair; pt < n_ptiles; pt += n_pairs) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NQ);
                for (int jb = 0; jb < kb2; ++jb) {
                    mbar_wait(&b_full[stage], phase);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t a_lo = lo0 + (uint32_t)((stage * Cfg::B_STAGE) >> 4);
                        const uint32_t q_lo = a_lo + (uint32_t)(TC_A_BYTES >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k)      // UMMA K = 16 bf16 (32 bytes = 2 descriptor units)
                            umma_2sm<true>(d_tmem, desc_with_lo(desc0, a_lo + 2 * k), desc_with_lo(desc0, q_lo + 2 * k), idesc,
                                           (uint32_t)((jb | k) != 0));
                        umma_commit_2sm(&b_empty[stage]);
                        if (jb == kb2 - 1) umma_commit_2sm(&tfull_bar[acc]);
                    }
                    __syncwarp();
                    if (++stage == TCV_B_STAGES) { stage = 0; phase ^= 1; }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (warp >= 11) {
        // ===================== converters: two fp32 staging k-blocks -> one bf16 operand k-block =====================
        static_assert(TCV_B_STAGES == 3 && TCV_A_STAGES == 8, "converter warp w owns operand stage w; staging ring of 8");
        const int w = warp - 11;                                // operand stage owned by this warp
        const int c = lane & 7;                                 // 16-byte chunk of the fp32 row
        // rows handled by a lane: base {0,4,1,5}[lane >> 3] + 2 (it & 1) + 8 (it >> 1).  The two rows of a half-warp differ in
        // bit 2, so their swizzled 64-byte bf16 halves fall into opposite bank halves (conflict-free 64-bit stores)
        const int rr = (((lane >> 3) & 1) << 2) | (lane >> 4);
        const int64_t my_tiles = (n_ptiles - pair + n_pairs - 1) / n_pairs;
        const int64_t total = my_tiles * kb2;                   // bf16 k-blocks this CTA converts
        unsigned char* dst = ring_b + (size_t)w * Cfg::B_STAGE;
        uint32_t pbe = 1;                                       // parity to wait for on b_empty[w]
        for (int64_t g = w; g < total; g += TCV_B_STAGES) {
            mbar_wait(&b_empty[w], pbe);                        // the MMAs that read this operand stage have retired
            pbe ^= 1;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int64_t i = 2 * g + h;                    // fp32 k-block counter of this CTA
                const int sa = (int)(i & 7);
                mbar_wait(&a_full[sa], (uint32_t)((i >> 3) & 1));
                const unsigned char* src = ring_a + (size_t)sa * TC_A_BYTES;
                const int cd = (h << 2) | (c >> 1);             // 16-byte chunk of the bf16 row
#pragma unroll
                for (int b8 = 0; b8 < 4; ++b8) {
                    float4 v[8];
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int r = rr + 2 * (it & 1) + 8 * (b8 * 4 + (it >> 1));
                        v[it] = *reinterpret_cast<const float4*>(src + r * 128 + ((c ^ (r & 7)) << 4));
                    }
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int r = rr + 2 * (it & 1) + 8 * (b8 * 4 + (it >> 1));
                        const __nv_bfloat162 lo = __floats2bfloat162_rn(v[it].x, v[it].y), hi = __floats2bfloat162_rn(v[it].z, v[it].w);
                        uint2 o;
                        o.x = *reinterpret_cast<const uint32_t*>(&lo); o.y = *reinterpret_cast<const uint32_t*>(&hi);
                        *reinterpret_cast<uint2*>(dst + r * 128 + ((cd ^ (r & 7)) << 4) + ((c & 1) << 3)) = o;
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_empty[sa]);        // staging slot read: TMA may refill it
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA's async proxy
            __syncwarp();
            if (lane == 0) { if (rank == 0) mbar_arrive(&b_full[w]); else mbar_arrive_remote(&b_full[w], 0); }
        }
    } else {
        // ===================== epilogue (warps 3-10): as dense_tc2_kernel =====================
        const int lg = warp & 3;
        const int col_half = (warp - 3) >> 2;
        const int row_in_tile = lg * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        for (int64_t pt = pair; pt < n_ptiles; pt += n_pairs) {
            const int64_t row = pt * (2 * TC_TILE_M) + rank * TC_TILE_M + row_in_tile;
            float xn = CUDART_INF_F;
            if (row < n_rows && (alive == nullptr || bit_test(alive, (uint32_t)row))) xn = xnorm[row];
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * NQ);
            const float half_xn = 0.5f * xn;
#pragma unroll 1
            for (int c0 = col_half * (NQ / 2); c0 < (col_half + 1) * (NQ / 2); c0 += 32) {
                uint32_t v[32];
                tmem_ld_x32(taddr + c0, v);
                tmem_wait_ld();
                float t[32];
                float mx = -CUDART_INF_F;
#pragma unroll
                for (int j4 = 0; j4 < 32; j4 += 4) {
                    const float4 h = *reinterpret_cast<const float4*>(s_thr + c0 + j4);
                    t[j4 + 0] = __uint_as_float(v[j4 + 0]) + h.x;
                    t[j4 + 1] = __uint_as_float(v[j4 + 1]) + h.y;
                    t[j4 + 2] = __uint_as_float(v[j4 + 2]) + h.z;
                    t[j4 + 3] = __uint_as_float(v[j4 + 3]) + h.w;
                    mx = fmaxf(mx, fmaxf(fmaxf(t[j4 + 0], t[j4 + 1]), fmaxf(t[j4 + 2], t[j4 + 3])));
                }
                if (mx > half_xn) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (t[j] > half_xn) {
                            const uint32_t pos = atomicAdd(&cand_count[c0 + j], 1u);
                            if (pos < (uint32_t)cap) cand_rows[(size_t)(c0 + j) * cap + pos] = (uint32_t)row;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) mbar_arrive(&tempty_bar[acc]);
                else mbar_arrive_remote(&tempty_bar[acc], 0);
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------- row norms (index time)
__global__ void __launch_bounds__(256)
row_norms_kernel(const float* __restrict__ X, int64_t row0, int64_t n, int dpad, float* __restrict__ xnorm,
                 uint32_t* __restrict__ xn_max_bits)
{
    const int lane = threadIdx.x & 31;
    const int64_t r = row0 + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (r >= row0 + n) return;
    const float4* p = reinterpret_cast<const float4*>(X + r * dpad);
    float s = 0.f;
    for (int i = lane; i < dpad / 4; i += 32) { float4 v = p[i]; s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) { xnorm[r] = s; atomicMax(xn_max_bits, __float_as_uint(s)); }
}
void launch_row_norms(const float* X, int64_t row0, int64_t n, int dpad, float* xnorm, uint32_t* xn_max_bits, cudaStream_t st)
{
    if (n == 0) return;
    row_norms_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(X, row0, n, dpad, xnorm, xn_max_bits);
    KRAG_CUDA(cudaGetLastError());
    count_launch();
}

// -------------------------------------------------------- sample -> per-query threshold
constexpr int ST_THREADS = 512;
// CAP = 1024 when m <= 256: the first prune (a bitonic sort of CAP keys, the dominant cost of this kernel) comes after two
// iterations instead of three and sorts half as many keys
template <int ST_CAP>
__global__ void __launch_bounds__(ST_THREADS)
sample_threshold_kernel(const float* __restrict__ dump, int64_t S, int m, int nq_valid, float* __restrict__ thr,
                        float* __restrict__ part_vals /* != null: grid (NQ, parts), write this part's m smallest values */)
{
    __shared__ uint64_t s_buf[ST_CAP];
    __shared__ int s_count;
    __shared__ uint64_t s_thr;
    const int tid = threadIdx.x, j = blockIdx.x;
    const int parts = gridDim.y, part = blockIdx.y;
    if (j >= nq_valid) {                                                    // padded query: admits nothing
        if (part_vals) { for (int i = tid; i < m; i += ST_THREADS) part_vals[((int64_t)j * parts + part) * m + i] = CUDART_INF_F; }
        else if (tid == 0) thr[j] = -CUDART_INF_F;
        return;
    }
    SelectBuf sel{s_buf, &s_count, &s_thr, ST_CAP};
    select_init(sel, tid);
    __syncthreads();
    // this block's slice of the query's samples (two-level selection: `parts` blocks per query, then one more pass
    // of this kernel over their parts * m survivors)
    const int64_t lo = S * part / parts, hi = S * (part + 1) / parts;
    const float* col = dump + (int64_t)j * S + lo;
    S = hi - lo;
    const int epoch = (ST_CAP - m) / ST_THREADS;
    uint64_t t = KEY_PAD;
    int it = 0;
    for (int64_t i0 = 0; i0 < S; i0 += ST_THREADS, ++it) {
        int64_t i = i0 + tid;
        if (i < S) select_push(sel, make_key_asc(col[i], (uint32_t)i), t);
        if ((it + 1) % epoch == 0) {
            __syncthreads();
            if (s_count + epoch * ST_THREADS > ST_CAP) select_prune<ST_THREADS>(sel, m, tid, 0);
            t = s_thr;
        }
    }
    select_prune<ST_THREADS>(sel, m, tid, 0);
    if (part_vals) {
        for (int i = tid; i < m; i += ST_THREADS)
            part_vals[((int64_t)j * parts + part) * m + i] = (i < s_count) ? key_value_asc(s_buf[i]) : CUDART_INF_F;
    } else if (tid == 0) {
        thr[j] = (s_count >= m) ? key_value_asc(s_buf[m - 1]) : CUDART_INF_F;
    }
}

// ------------------------------------------------------------------ exact rescoring
// one 8-lane group per candidate; same arithmetic as dense_scan_kernel / oracle l2sq_row
__global__ void __launch_bounds__(256)
rescore_kernel(const float* __restrict__ X, int dpad, const float* __restrict__ Q, const uint32_t* __restrict__ cand_count,
               const uint32_t* __restrict__ cand_rows, int cap, OrdMap ord_base, uint64_t* __restrict__ exact_keys)
{
    extern __shared__ __align__(16) float rs_q[];   // [dpad]
    const int j = blockIdx.y, tid = threadIdx.x, lane = tid & 31, l8 = lane & 7;
    const int cnt = (int)min(cand_count[j], (uint32_t)cap);
    const int c = blockIdx.x * 32 + (tid >> 3);
    if (blockIdx.x * 32 >= cnt) {   // whole block past the end: pad and leave
        if (l8 == 0 && c < cap) exact_keys[(size_t)j * cap + c] = KEY_PAD;
        return;
    }
    for (int i = tid; i < dpad; i += 256) rs_q[i] = Q[(size_t)j * dpad + i];
    __syncthreads();
    const bool valid = c < cnt;
    const uint32_t row = valid ? cand_rows[(size_t)j * cap + c] : 0u;
    const float4* p = reinterpret_cast<const float4*>(X + (size_t)row * dpad) + l8;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int steps = dpad / KRAG_LANES;
#pragma unroll 4
    for (int i = 0; i < steps; ++i) {
        const float4 a = __ldg(p + i * 8);
        const float4 qv = *reinterpret_cast<const float4*>(rs_q + i * KRAG_LANES + l8 * 4);
        float t;
        t = a.x - qv.x; acc.x = fmaf(t, t, acc.x);
        t = a.y - qv.y; acc.y = fmaf(t, t, acc.y);
        t = a.z - qv.z; acc.z = fmaf(t, t, acc.z);
        t = a.w - qv.w; acc.w = fmaf(t, t, acc.w);
    }
    float s = (acc.x + acc.y) + (acc.z + acc.w);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (l8 == 0 && c < cap) exact_keys[(size_t)j * cap + c] = valid ? make_key_asc(s, ord_base + row) : KEY_PAD;
}

// ------------------------------------------------------------------------ certificate
__global__ void certify_kernel(const float* __restrict__ Q, int dpad, int nq, int P, const float* __restrict__ thr,
                               const uint32_t* __restrict__ cand_count, int cap, const uint64_t* __restrict__ keys_out,
                               const uint32_t* __restrict__ xn_max_bits, float eps_rel, int32_t* __restrict__ flags)
{
    const int j = blockIdx.x * blockDim.x / 32 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= nq) return;
    float qn = 0.f;
    for (int i = lane; i < dpad; i += 32) { float v = Q[(size_t)j * dpad + i]; qn = fmaf(v, v, qn); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qn += __shfl_xor_sync(0xffffffffu, qn, o);
    if (lane == 0) {
        const uint64_t kP = keys_out[(size_t)j * P + (P - 1)];
        const float xmax = __uint_as_float(*xn_max_bits);
        bool ok = cand_count[j] <= (uint32_t)cap && kP != KEY_PAD;
        if (ok) {
            const float dP = key_value_asc(kP);
            // |tf32 dot - exact dot| <= (2^-9 + 2^-12) |x| |q|  (operand truncation 2^-10 each + fp32 accumulation)
            const float eps = eps_rel * sqrtf(xmax) * sqrtf(qn);   // eps_rel: 2^-9+2^-12 (TF32 truncation), 2^-8+2^-12 (bf16 RN)
            const float slack = 4e-6f * (1.f + qn + xmax);   // fp32 rounding of |x|^2, |q|^2 and the epilogue fma
            ok = dP <= thr[j] + qn - 2.f * eps - slack;
        }
        flags[j] = ok ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}
static bool make_map(CUtensorMap* tm, const void* base, int64_t rows, int dpad, int box_rows, bool bf16 = false)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)dpad, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)dpad * (bf16 ? 2 : 4)};
    cuuint32_t box[2] = {(cuuint32_t)(bf16 ? 2 * TC_KBLOCK : TC_KBLOCK), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

constexpr int64_t TC_MIN_ROWS = 262144;   // below this the exact scan is cheaper than sample + prune + rescore

bool dense_tc_supported(const DeviceInfo& di, int dpad)
{
    return di.cc_major == 10 && dpad % TC_KBLOCK == 0 && di.smem_optin >= TcCfg<256>::SMEM && get_encode() != nullptr;
}

static int tc_target(int P, bool bf16 = false)
{
    int c = (bf16 ? 6 : 4) * P, lo = bf16 ? 1024 : 512;   // bf16 pruning needs a wider margin for its certificate
    return c < lo ? lo : (c > 3072 ? 3072 : c);
}
// Gamma(m)-distributed sample estimate (m = C/64 >= 8): 4x the target is a < 1e-8 overflow tail
static int tc_cap(int P) { return 4 * tc_target(P, true); }   // sized for either mode

struct TcWorkspace {   // carved out of the caller's byte buffer
    float* thr; uint32_t* cand_count; int32_t* flags; uint32_t* cand_rows; uint64_t* exact_keys; float* dump; uint16_t* q_bf16;
};
static size_t tc_carve(TcWorkspace* w, unsigned char* base, int cap, int64_t S)
{
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = (o + bytes + 255) & ~(size_t)255; return base ? base + at : nullptr; };
    unsigned char* p;
    p = take(256 * 4); if (w) w->thr = (float*)p;
    p = take(256 * 4); if (w) w->cand_count = (uint32_t*)p;
    p = take(256 * 4); if (w) w->flags = (int32_t*)p;
    p = take((size_t)256 * cap * 4); if (w) w->cand_rows = (uint32_t*)p;
    p = take((size_t)256 * cap * 8); if (w) w->exact_keys = (uint64_t*)p;
    p = take((size_t)256 * S * 4); if (w) w->dump = (float*)p;
    p = take((size_t)256 * 16384 * 2); if (w) w->q_bf16 = (uint16_t*)p;   // queries as bf16 (dim <= 16384)
    return o;
}
static int64_t tc_sample_tiles(int64_t n_rows)
{
    int64_t n_tiles = (n_rows + TC_TILE_M - 1) / TC_TILE_M;
    return (n_tiles + TC_SAMPLE_STRIDE - 1) / TC_SAMPLE_STRIDE;
}
size_t dense_tc_workspace_bytes(const DeviceInfo&, int64_t n_rows, int P)
{
    return tc_carve(nullptr, nullptr, tc_cap(P), tc_sample_tiles(n_rows) * TC_TILE_M);
}
bool dense_tc_wants(int64_t n_rows, int batch) { return n_rows >= TC_MIN_ROWS && batch >= 16; }

static bool tc_use_2cta()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("KRAG_TC_2CTA"); v = (e == nullptr || e[0] != '0') ? 1 : 0; }
    return v == 1;
}

template <int NQ>
static void tc_pass(const DeviceInfo& di, const CUtensorMap& tmX, const CUtensorMap& tmQ, const float* X, int64_t n_rows,
                    int dpad, const float* xnorm, const uint32_t* xn_max_bits, const uint32_t* alive, const float* q,
                    int nq, int P, OrdMap ord_base, const TcWorkspace& w, int cap, int64_t S, uint64_t* keys_out,
                    const uint16_t* Xh, bool cvt, cudaStream_t st)
{
    using Cfg = TcCfg<NQ>;
    const bool bf16 = Xh != nullptr || cvt;      // bf16 operands in the prune pass (shadow copy, or converted on chip)
    static bool attr_set = false;
    if (!attr_set) {
        KRAG_CUDA(cudaFuncSetAttribute(dense_tc_kernel<NQ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        KRAG_CUDA(cudaFuncSetAttribute(dense_tc_kernel<NQ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr_set = true;
    }
    const int kblocks = dpad / TC_KBLOCK;
    const int64_t n_tiles = (n_rows + TC_TILE_M - 1) / TC_TILE_M;
    const int64_t s_tiles = tc_sample_tiles(n_rows);
    const int grid_s = (int)(s_tiles < di.sm_count ? s_tiles : di.sm_count);
    const int grid_m = (int)(n_tiles < di.sm_count ? n_tiles : di.sm_count);
    // 1. sample pass (1/64 of the tiles), dump a = |x|^2 - 2 x.q
    // the dump holds one value per (32-row group of a sampled tile, query): S_min per query
    const int64_t S_min = s_tiles * (TC_TILE_M / 32);
    (void)S;
    dense_tc_kernel<NQ, true><<<grid_s, TC_THREADS, Cfg::SMEM, st>>>(tmX, tmQ, n_rows, kblocks, s_tiles, TC_SAMPLE_STRIDE, xnorm,
                                                                     alive, nullptr, nullptr, nullptr, 0, w.dump, S_min, 1);
    KRAG_CUDA(cudaGetLastError());
    count_launch();
    // 2. threshold admitting ~C rows: the m-th smallest of the sample, m = C * sample_fraction
    int64_t sampled_rows = 0;
    for (int64_t t = 0; t < s_tiles; ++t) {
        int64_t r0 = t * TC_SAMPLE_STRIDE * TC_TILE_M;
        sampled_rows += (r0 + TC_TILE_M <= n_rows) ? TC_TILE_M : (n_rows > r0 ? n_rows - r0 : 0);
    }
    int m = (int)((double)tc_target(P, bf16) * (double)sampled_rows / (double)n_rows + 0.5);
    if (m < 4) m = 4;
    if (m > 1024) m = 1024;
    // m-th smallest of the S_min group minima per query (a few thousand values: one block per query, one launch)
    if (m > (int)S_min) m = (int)S_min;
    if (m <= 256) sample_threshold_kernel<1024><<<dim3(NQ, 1), ST_THREADS, 0, st>>>(w.dump, S_min, m, nq, w.thr, nullptr);
    else sample_threshold_kernel<2048><<<dim3(NQ, 1), ST_THREADS, 0, st>>>(w.dump, S_min, m, nq, w.thr, nullptr);
    KRAG_CUDA(cudaGetLastError());
    count_launch();
    KRAG_CUDA(cudaMemsetAsync(w.cand_count, 0, 256 * 4, st));
    // 3. main pass: stream the corpus once, prune on the tensor cores
    if (bf16) launch_f32_to_bf16(q, w.q_bf16, (int64_t)nq * dpad, st);
    dense_timer_begin(st, cvt ? 5 : (bf16 ? 4 : (tc_use_2cta() ? 3 : 2)), n_rows * (int64_t)dpad * ((bf16 && !cvt) ? 2 : 4),
                      2 * (int64_t)NQ * n_rows * dpad);
    if (cvt) {
        using CfgV = TcvCfg<NQ>;
        static bool attrv_set = false;
        if (!attrv_set) {
            KRAG_CUDA(cudaFuncSetAttribute(dense_tc2cvt_kernel<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CfgV::SMEM));
            attrv_set = true;
        }
        CUtensorMap tmQh;
        if (!make_map(&tmQh, (const void*)w.q_bf16, nq, dpad, NQ / 2, true))
            throw CudaError{cudaErrorUnknown, "cuTensorMapEncodeTiled(Q half, bf16)", __FILE__, __LINE__};
        const int64_t n_ptiles = (n_rows + 2 * TC_TILE_M - 1) / (2 * TC_TILE_M);
        const int64_t max_pairs = di.sm_count / 2;
        const int n_pairs = (int)(n_ptiles < max_pairs ? n_ptiles : max_pairs);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(2 * n_pairs));
        cfg.blockDim = dim3(TCV_THREADS);
        cfg.dynamicSmemBytes = CfgV::SMEM;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        KRAG_CUDA(cudaLaunchKernelEx(&cfg, dense_tc2cvt_kernel<NQ>, tmX, tmQh, n_rows, kblocks, n_ptiles, xnorm, alive,
                                     (const float*)w.thr, w.cand_count, w.cand_rows, cap));
    } else if (tc_use_2cta() || bf16) {
        using Cfg2 = Tc2Cfg<NQ>;
        static bool attr2_set = false;
        if (!attr2_set) {
            KRAG_CUDA(cudaFuncSetAttribute(dense_tc2_kernel<NQ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg2::SMEM));
            KRAG_CUDA(cudaFuncSetAttribute(dense_tc2_kernel<NQ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg2::SMEM));
            attr2_set = true;
        }
        CUtensorMap tmQh, tmXh;
        if (!make_map(&tmQh, bf16 ? (const void*)w.q_bf16 : (const void*)q, nq, dpad, NQ / 2, bf16))
            throw CudaError{cudaErrorUnknown, "cuTensorMapEncodeTiled(Q half)", __FILE__, __LINE__};
        if (bf16 && !make_map(&tmXh, Xh, n_rows, dpad, TC_TILE_M, true))
            throw CudaError{cudaErrorUnknown, "cuTensorMapEncodeTiled(bf16 shadow)", __FILE__, __LINE__};
        const int kb2 = bf16 ? dpad / (2 * TC_KBLOCK) : kblocks;
        const int64_t n_ptiles = (n_rows + 2 * TC_TILE_M - 1) / (2 * TC_TILE_M);
        const int64_t max_pairs = di.sm_count / 2;
        const int n_pairs = (int)(n_ptiles < max_pairs ? n_ptiles : max_pairs);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(2 * n_pairs));
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = Cfg2::SMEM;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        if (bf16)
            KRAG_CUDA(cudaLaunchKernelEx(&cfg, dense_tc2_kernel<NQ, true>, tmXh, tmQh, n_rows, kb2, n_ptiles, xnorm, alive,
                                         (const float*)w.thr, w.cand_count, w.cand_rows, cap));
        else
            KRAG_CUDA(cudaLaunchKernelEx(&cfg, dense_tc2_kernel<NQ, false>, tmX, tmQh, n_rows, kb2, n_ptiles, xnorm, alive,
                                         (const float*)w.thr, w.cand_count, w.cand_rows, cap));
    } else {
        dense_tc_kernel<NQ, false><<<grid_m, TC_THREADS, Cfg::SMEM, st>>>(tmX, tmQ, n_rows, kblocks, n_tiles, 1, xnorm, alive, w.thr,
                                                                          w.cand_count, w.cand_rows, cap, nullptr, 0, 0);
    }
    dense_timer_end(st);
    KRAG_CUDA(cudaGetLastError());
    count_launch();
    // 4. exact fp32 rescoring of the survivors + top-P
    dim3 rg((unsigned)((cap + 31) / 32), (unsigned)nq);
    rescore_kernel<<<rg, 256, (size_t)dpad * 4, st>>>(X, dpad, q, w.cand_count, w.cand_rows, cap, ord_base, w.exact_keys);
    KRAG_CUDA(cudaGetLastError());
    count_launch();
    launch_merge(w.exact_keys, 1, cap, nq, P, cap, cap, keys_out, st);
    // 5. certificate
    // the threshold came from TF32 values and the pruning from bf16 ones: the certificate uses the sum of both error bounds
    const float eps_rel = bf16 ? (0.00390625f + 0.001953125f + 0.00048828125f) : (0.001953125f + 0.000244140625f);
    certify_kernel<<<(nq * 32 + 255) / 256, 256, 0, st>>>(q, dpad, nq, P, w.thr, w.cand_count, cap, keys_out, xn_max_bits, eps_rel, w.flags);
    KRAG_CUDA(cudaGetLastError());
    count_launch();
}

static bool tc_use_cvt()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("KRAG_TC_CVT"); v = (e == nullptr || e[0] != '0') ? 1 : 0; }
    return v == 1;
}

static std::atomic<int64_t> g_tc_fallback_queries{0};   // searches run concurrently (reader lock)
int64_t dense_tc_fallback_queries() { return g_tc_fallback_queries.load(); }

bool launch_dense_tc(const DeviceInfo& di, const float* X, int64_t n_rows, int dpad, const uint32_t* alive,
                     const float* xnorm, const uint32_t* xn_max_bits, const float* q, int batch, int P, OrdMap ord_base,
                     void* workspace, size_t workspace_bytes, uint64_t* part, uint64_t* keys_out, cudaStream_t st,
                     const uint16_t* Xh, bool allow_cvt)
{
    if (Xh != nullptr && dpad % (2 * TC_KBLOCK) != 0) Xh = nullptr;   // bf16 k-blocks are 64 elements wide
    // default prune pass: fp32 rows converted to bf16 in shared memory (no shadow); KRAG_TC_CVT=0 keeps the TF32 pass
    const bool cvt = Xh == nullptr && allow_cvt && tc_use_cvt() && tc_use_2cta() && dpad % (2 * TC_KBLOCK) == 0 &&
                     di.smem_optin >= TcvCfg<256>::SMEM;
    if (n_rows < TC_MIN_ROWS || n_rows >= (1ll << 31) || xnorm == nullptr) return false;
    const int cap = tc_cap(P);
    const int64_t S = tc_sample_tiles(n_rows) * TC_TILE_M;
    TcWorkspace w;
    if (tc_carve(&w, (unsigned char*)workspace, cap, S) > workspace_bytes) return false;
    CUtensorMap tmX;
    if (!make_map(&tmX, X, n_rows, dpad, TC_TILE_M)) return false;
    std::vector<int32_t> flags(256);
    for (int b0 = 0; b0 < batch; b0 += 256) {
        const int nq = batch - b0 < 256 ? batch - b0 : 256;
        const int NQ = nq <= 64 ? 64 : (nq <= 128 ? 128 : 256);
        const float* qb = q + (size_t)b0 * dpad;
        CUtensorMap tmQ;
        if (!make_map(&tmQ, qb, nq, dpad, NQ)) return false;
        uint64_t* ko = keys_out + (size_t)b0 * P;
        if (NQ == 64) tc_pass<64>(di, tmX, tmQ, X, n_rows, dpad, xnorm, xn_max_bits, alive, qb, nq, P, ord_base, w, cap, S, ko, Xh, cvt, st);
        else if (NQ == 128) tc_pass<128>(di, tmX, tmQ, X, n_rows, dpad, xnorm, xn_max_bits, alive, qb, nq, P, ord_base, w, cap, S, ko, Xh, cvt, st);
        else tc_pass<256>(di, tmX, tmQ, X, n_rows, dpad, xnorm, xn_max_bits, alive, qb, nq, P, ord_base, w, cap, S, ko, Xh, cvt, st);
        // uncertified queries (rare) are re-run on the exact scan kernel -- still on the GPU
        KRAG_CUDA(cudaMemcpyAsync(flags.data(), w.flags, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, st));
        KRAG_CUDA(cudaStreamSynchronize(st));
        for (int j = 0; j < nq; ++j) {
            if (flags[(size_t)j]) continue;
            ++g_tc_fallback_queries;
            launch_dense_scan(di, X, n_rows, dpad, alive, qb + (size_t)j * dpad, 1, P, ord_base, part, ko + (size_t)j * P, st);
        }
    }
    return true;
}

bool dense_tc_debug_dump(const DeviceInfo& di, const float* X, int64_t n_rows, int dpad, const float* xnorm,
                         const float* q, int nq, float* dump_out, int64_t* S_out, int* nq_pad_out, cudaStream_t st)
{
    if (!dense_tc_supported(di, dpad) || nq > 256) return false;
    const int NQ = nq <= 64 ? 64 : (nq <= 128 ? 128 : 256);
    const int64_t n_tiles = (n_rows + TC_TILE_M - 1) / TC_TILE_M;
    const int64_t S = n_tiles * TC_TILE_M;
    *S_out = S;
    *nq_pad_out = NQ;
    if (dump_out == nullptr) return true;   // size query
    CUtensorMap tmX, tmQ;
    if (!make_map(&tmX, X, n_rows, dpad, TC_TILE_M) || !make_map(&tmQ, q, nq, dpad, NQ)) return false;
    const int kblocks = dpad / TC_KBLOCK;
    const int grid = (int)(n_tiles < di.sm_count ? n_tiles : di.sm_count);
#define KRAG_TC_DUMP(N)                                                                                                    \
    do {                                                                                                                   \
        KRAG_CUDA(cudaFuncSetAttribute(dense_tc_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcCfg<N>::SMEM)); \
        dense_tc_kernel<N, true><<<grid, TC_THREADS, TcCfg<N>::SMEM, st>>>(tmX, tmQ, n_rows, kblocks, n_tiles, 1, xnorm, nullptr,   \
                                                                           nullptr, nullptr, nullptr, 0, dump_out, S, 0);  \
    } while (0)
    if (NQ == 64) KRAG_TC_DUMP(64); else if (NQ == 128) KRAG_TC_DUMP(128); else KRAG_TC_DUMP(256);
#undef KRAG_TC_DUMP
    KRAG_CUDA(cudaGetLastError());
    count_launch();
    return true;
}

}  // namespace krag
