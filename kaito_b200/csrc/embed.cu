// embed.cu -- K5: BERT-family (bge-small / bge-base / bge-large) encoder forward on the GPU.
//
// Replaces the torch / sentence-transformers forward behind
// presets/ragengine/embedding/huggingface_local_embedding.py:34-53 (LlamaIndex HuggingFaceEmbedding ->
// BertModel, CLS pooling, L2 normalisation; reached from embedding/base.py:25-26 on /retrieve and from
// VectorStoreIndex.from_documents on /index, vector_store/base.py:155-166).
//
//   tokens (packed, variable length: no padding rows)  ->  word+position+type embeddings -> LayerNorm
//   L x [ QKV GEMM -> attention -> O GEMM (+bias +residual) -> LayerNorm -> FFN1 GEMM (+bias, erf-GELU)
//         -> FFN2 GEMM (+bias +residual) -> LayerNorm ]  ->  CLS row  ->  L2 normalise
//
// The four GEMMs per layer (24 S d^2 of the 24 S d^2 + 4 S^2 d flops) run on tcgen05.mma kind::tf32 with
// fp32 accumulation in TMEM, operands staged by TMA (fp32 weights/activations are consumed as TF32, no
// conversion pass); bias / GELU / residual are fused into the TMEM epilogue.  LayerNorm, softmax and the
// attention products are fp32 CUDA-core code (attention is 4 S^2 d: 2% of the flops at S = 32 queries).
#include <math_constants.h>
#include <stdlib.h>

#include <map>
#include <stdexcept>
#include <tuple>
#include <string>
#include <vector>

#include "engine.h"
#include "embed_config.h"
#include "tc_ptx.cuh"
#include "common.cuh"

namespace krag {

// ------------------------------------------------------------------ GEMM: C = A . B^T (+bias)(gelu)(+res)
// A [M, K] row-major (activations), B [N, K] row-major (nn.Linear weight), C [M, N] row-major, all fp32.
// Persistent CTAs over 128 x 128 output tiles (n fastest so neighbouring CTAs share the A tile in L2);
// one TMA warp, one MMA-issuing thread, four epilogue warps; stage = 2 k-blocks of 32 floats for A and B.
constexpr int GM_TILE = 128;
constexpr int GM_KB = 32;                 // floats per k-block (128-byte swizzle row)
constexpr int GM_KB_PER_STAGE = 2;
constexpr int GM_SLAB = GM_TILE * GM_KB * 4;                     // 16 KB
constexpr int GM_STAGE_BYTES = GM_KB_PER_STAGE * 2 * GM_SLAB;    // 64 KB
constexpr int GM_STAGES = 3;
constexpr int GM_THREADS = 192;
constexpr size_t GM_SMEM = (size_t)GM_STAGES * GM_STAGE_BYTES + 1024 + 256;

// 256-bit global accesses (sm_100): one full 32-byte sector per thread and instruction
__device__ __forceinline__ void st_global_v8(float* p, const float* o)
{
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7]) : "memory");
}
__device__ __forceinline__ void ld_global_v8(const float* p, float* r)
{
    asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "l"(p));
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

__global__ void __launch_bounds__(GM_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                 const float* __restrict__ bias, const float* __restrict__ residual, int act_gelu, float* __restrict__ C)
{
    extern __shared__ unsigned char gm_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gm_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* tail = smem + (size_t)GM_STAGES * GM_STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
    uint64_t* empty_bar = full_bar + GM_STAGES;
    uint64_t* tfull_bar = empty_bar + GM_STAGES;      // [2]
    uint64_t* tempty_bar = tfull_bar + 2;             // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (M + GM_TILE - 1) / GM_TILE, n_tiles = N / GM_TILE;
    const int total = m_tiles * n_tiles;
    const int kblocks = K / GM_KB;
    const int stages_per_tile = (kblocks + GM_KB_PER_STAGE - 1) / GM_KB_PER_STAGE;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < GM_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 4); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            const int m0 = (t / n_tiles) * GM_TILE, n0 = (t % n_tiles) * GM_TILE;
            for (int sk = 0; sk < stages_per_tile; ++sk) {
                const int nkb = min(GM_KB_PER_STAGE, kblocks - sk * GM_KB_PER_STAGE);
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (lane == 0) {
                    unsigned char* sa = smem + (size_t)stage * GM_STAGE_BYTES;
                    unsigned char* sb = sa + GM_KB_PER_STAGE * GM_SLAB;
                    mbar_expect_tx(&full_bar[stage], nkb * 2 * GM_SLAB);
                    for (int u = 0; u < nkb; ++u) {
                        const int k0 = (sk * GM_KB_PER_STAGE + u) * GM_KB;
                        tma_load_2d(sa + (size_t)u * GM_SLAB, &tmA, &full_bar[stage], k0, m0, TMA_EVICT_FIRST);
                        tma_load_2d(sb + (size_t)u * GM_SLAB, &tmB, &full_bar[stage], k0, n0, TMA_EVICT_LAST);
                    }
                }
                __syncwarp();
                if (++stage == GM_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_tf32(GM_TILE, GM_TILE);
        const uint64_t desc0 = umma_desc_sw128(smem_u32(smem));
        const uint32_t lo0 = (uint32_t)desc0;
        int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * GM_TILE);
            for (int sk = 0; sk < stages_per_tile; ++sk) {
                const int nkb = min(GM_KB_PER_STAGE, kblocks - sk * GM_KB_PER_STAGE);
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_lo = lo0 + (uint32_t)((stage * GM_STAGE_BYTES) >> 4);
                    const uint32_t b_lo = a_lo + (uint32_t)((GM_KB_PER_STAGE * GM_SLAB) >> 4);
#pragma unroll
                    for (int u = 0; u < GM_KB_PER_STAGE; ++u) {
                        if (u < nkb) {
#pragma unroll
                            for (int k = 0; k < GM_KB / 8; ++k) {
                                const uint64_t ad = (desc0 & 0xFFFFFFFF00000000ull) | (uint64_t)(a_lo + (uint32_t)((u * GM_SLAB) >> 4) + 2 * k);
                                const uint64_t bd = (desc0 & 0xFFFFFFFF00000000ull) | (uint64_t)(b_lo + (uint32_t)((u * GM_SLAB) >> 4) + 2 * k);
                                umma_tf32(d_tmem, ad, bd, idesc, (uint32_t)((sk | u | k) != 0));
                            }
                        }
                    }
                    umma_commit(&empty_bar[stage]);
                    if (sk == stages_per_tile - 1) umma_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                if (++stage == GM_STAGES) { stage = 0; phase ^= 1; }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    } else {
        const int lg = warp & 3;
        const int r_in = lg * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            const int m0 = (t / n_tiles) * GM_TILE, n0 = (t % n_tiles) * GM_TILE;
            const int row = m0 + r_in;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * GM_TILE);
#pragma unroll 1
            for (int c0 = 0; c0 < GM_TILE; c0 += 32) {
                uint32_t v[32];
                tmem_ld_x32(taddr + c0, v);
                tmem_wait_ld();
                if (row < M) {
                    float* crow = C + (size_t)row * N + n0 + c0;
                    const float* rrow = residual ? residual + (size_t)row * N + n0 + c0 : nullptr;
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {   // 32-byte (full-sector) vector accesses: a thread owns a row segment
                        float o[8];
#pragma unroll
                        for (int t = 0; t < 8; t += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c0 + j + t));
                            o[t + 0] = __uint_as_float(v[j + t + 0]) + b4.x; o[t + 1] = __uint_as_float(v[j + t + 1]) + b4.y;
                            o[t + 2] = __uint_as_float(v[j + t + 2]) + b4.z; o[t + 3] = __uint_as_float(v[j + t + 3]) + b4.w;
                        }
                        if (act_gelu) {
#pragma unroll
                            for (int t = 0; t < 8; ++t) o[t] = gelu_erf(o[t]);
                        }
                        if (rrow) {
                            float r[8];
                            ld_global_v8(rrow + j, r);
#pragma unroll
                            for (int t = 0; t < 8; ++t) o[t] += r[t];
                        }
                        st_global_v8(crow + j, o);
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
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

// ---------------------------------------------------------------- 2-CTA GEMM (cta_group::2, M = 256 per pair)
// Same structure as dense_tc2_kernel (dense_tc.cu): a CTA pair owns a 256 x BN output tile; each CTA stages its
// 128 activation rows and HALF of the BN weight rows, the pair's tensor cores share the operands
// (per SM and 128-cycle MMA: 8 KB operand reads + 64 B/clk of TMA writes instead of 12 KB + 96 B/clk, which capped
// the 1-CTA kernel at ~40% tensor-pipe activity on these shapes); one ring, 2 k-blocks per stage, one issuer thread.
template <int BN> struct Gm2Cfg {
    static constexpr int BH_BYTES = (BN / 2) * GM_KB * 4;              // this CTA's half of one weight k-block
    static constexpr int STAGE_BYTES = GM_KB_PER_STAGE * (GM_SLAB + BH_BYTES);
    static constexpr int STAGES = (200 * 1024) / STAGE_BYTES;          // 3 at BN = 256, 4 at BN = 128
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 + 256;
};
constexpr int GM2_THREADS = 320;   // warp 0: TMA, warp 1: MMA issuer (leader), warps 2-9: epilogue (2 per TMEM lane group)

template <int BN>
__global__ void __launch_bounds__(GM2_THREADS, 1)
gemm2_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh, int M, int N, int K,
                  const float* __restrict__ bias, const float* __restrict__ residual, int act_gelu, float* __restrict__ C)
{
    using Cfg = Gm2Cfg<BN>;
    extern __shared__ unsigned char gm_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gm_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* tail = smem + (size_t)Cfg::STAGES * Cfg::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);             // leader only: 2 arrivals + tx bytes of both CTAs
    uint64_t* empty_bar = full_bar + Cfg::STAGES;                        // per CTA
    uint64_t* tfull_bar = empty_bar + Cfg::STAGES;                       // per CTA [2]
    uint64_t* tempty_bar = tfull_bar + 2;                                // leader only [2]: 16 arrivals
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int m_ptiles = (M + 2 * GM_TILE - 1) / (2 * GM_TILE), n_tiles = N / BN;
    const int total = m_ptiles * n_tiles;
    const int kblocks = K / GM_KB;
    const int stages_per_tile = (kblocks + GM_KB_PER_STAGE - 1) / GM_KB_PER_STAGE;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmBh);
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 16); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int t = pair; t < total; t += n_pairs) {
            const int m0 = (t / n_tiles) * (2 * GM_TILE) + (int)rank * GM_TILE, n0 = (t % n_tiles) * BN + (int)rank * (BN / 2);
            for (int sk = 0; sk < stages_per_tile; ++sk) {
                const int nkb = min(GM_KB_PER_STAGE, kblocks - sk * GM_KB_PER_STAGE);
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (lane == 0) {
                    unsigned char* sa = smem + (size_t)stage * Cfg::STAGE_BYTES;
                    unsigned char* sb = sa + GM_KB_PER_STAGE * GM_SLAB;
                    if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * nkb * (GM_SLAB + Cfg::BH_BYTES));
                    else mbar_arrive_remote(&full_bar[stage], 0);
                    for (int u = 0; u < nkb; ++u) {
                        const int k0 = (sk * GM_KB_PER_STAGE + u) * GM_KB;
                        tma_load_2d_2sm(sa + (size_t)u * GM_SLAB, &tmA, &full_bar[stage], k0, m0, TMA_EVICT_FIRST);
                        tma_load_2d_2sm(sb + (size_t)u * Cfg::BH_BYTES, &tmBh, &full_bar[stage], k0, n0, TMA_EVICT_LAST);
                    }
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(2 * GM_TILE, BN);
            const uint64_t desc0 = umma_desc_sw128(smem_u32(smem));
            const uint32_t lo0 = (uint32_t)desc0;
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
            for (int t = pair; t < total; t += n_pairs) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int sk = 0; sk < stages_per_tile; ++sk) {
                    const int nkb = min(GM_KB_PER_STAGE, kblocks - sk * GM_KB_PER_STAGE);
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t a_lo = lo0 + (uint32_t)((stage * Cfg::STAGE_BYTES) >> 4);
                        const uint32_t b_lo = a_lo + (uint32_t)((GM_KB_PER_STAGE * GM_SLAB) >> 4);
#pragma unroll
                        for (int u = 0; u < GM_KB_PER_STAGE; ++u) {
                            if (u < nkb) {
#pragma unroll
                                for (int k = 0; k < GM_KB / 8; ++k) {
                                    const uint64_t ad = (desc0 & 0xFFFFFFFF00000000ull) | (uint64_t)(a_lo + (uint32_t)((u * GM_SLAB) >> 4) + 2 * k);
                                    const uint64_t bd = (desc0 & 0xFFFFFFFF00000000ull) | (uint64_t)(b_lo + (uint32_t)((u * Cfg::BH_BYTES) >> 4) + 2 * k);
                                    umma_tf32_2sm(d_tmem, ad, bd, idesc, (uint32_t)((sk | u | k) != 0));
                                }
                            }
                        }
                        umma_commit_2sm(&empty_bar[stage]);
                        if (sk == stages_per_tile - 1) umma_commit_2sm(&tfull_bar[acc]);
                    }
                    __syncwarp();
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        // 8 epilogue warps: lane group = warp & 3, column half = (warp - 2) >> 2
        const int lg = warp & 3;
        const int col_half = (warp - 2) >> 2;
        const int r_in = lg * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        for (int t = pair; t < total; t += n_pairs) {
            const int m0 = (t / n_tiles) * (2 * GM_TILE) + (int)rank * GM_TILE, n0 = (t % n_tiles) * BN;
            const int row = m0 + r_in;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
            for (int c0 = col_half * (BN / 2); c0 < (col_half + 1) * (BN / 2); c0 += 32) {
                uint32_t v[32];
                tmem_ld_x32(taddr + c0, v);
                tmem_wait_ld();
                if (row < M) {
                    float* crow = C + (size_t)row * N + n0 + c0;
                    const float* rrow = residual ? residual + (size_t)row * N + n0 + c0 : nullptr;
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {   // 32-byte (full-sector) vector accesses: a thread owns a row segment
                        float o[8];
#pragma unroll
                        for (int t = 0; t < 8; t += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c0 + j + t));
                            o[t + 0] = __uint_as_float(v[j + t + 0]) + b4.x; o[t + 1] = __uint_as_float(v[j + t + 1]) + b4.y;
                            o[t + 2] = __uint_as_float(v[j + t + 2]) + b4.z; o[t + 3] = __uint_as_float(v[j + t + 3]) + b4.w;
                        }
                        if (act_gelu) {
#pragma unroll
                            for (int t = 0; t < 8; ++t) o[t] = gelu_erf(o[t]);
                        }
                        if (rrow) {
                            float r[8];
                            ld_global_v8(rrow + j, r);
#pragma unroll
                            for (int t = 0; t < 8; ++t) o[t] += r[t];
                        }
                        st_global_v8(crow + j, o);
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

// ---------------------------------------------------------------- small-M GEMM: 128 x BN tiles, split-K
// A handful of query tokens (batch-1 latency; the per-rank slice of a batch on 8 GPUs) gives the persistent kernels
// above 6..24 tiles for 148 SMs, each CTA streaming up to 1.5 MB of weights through a serial K loop.  This kernel
// cuts the output into 128 x BN tiles (BN = 32 or 64) AND the K range into `splits` slices, one CTA each, so that
// ~148 CTAs pull the weight matrix concurrently; when M < 128 only round8(M) activation rows are staged (the other
// MMA rows compute on stale shared memory and are never stored).  splits == 1: bias / GELU / residual epilogue
// straight to C; splits > 1: raw fp32 partials to ws[split][M][N], summed in split order by splitk_reduce_kernel
// (deterministic), which also applies bias / residual / LayerNorm.
// Stages are packed: round8(M) activation rows (a_bytes) + BN weight rows per k-block, so that for a handful of
// tokens the whole K slice of a CTA (16..24 k-blocks of ~6 KB) is in flight at once and two CTAs fit on an SM.  The
// MMA still reads a 128-row A operand from each stage base; the bytes past a_bytes belong to later stages (or the
// 16 KB slack after the last one) and only feed accumulator rows that are never stored.
constexpr int GSK_MAX_STAGES = 32;
template <int BN> struct GskCfg {
    static constexpr int B_BYTES = BN * GM_KB * 4;
};
static inline size_t gsk_smem_bytes(int n_stages, int stage_bytes) { return 2048 + (size_t)n_stages * stage_bytes + GM_SLAB; }

template <int BN>
__global__ void __launch_bounds__(GM_THREADS, 1)
gemm_sk_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int kb_per_split,
                    int a_bytes, int n_stages, int b_evict_first, const float* __restrict__ bias, const float* __restrict__ residual,
                    int act_gelu, float* __restrict__ C, float* __restrict__ ws)
{
    using Cfg = GskCfg<BN>;
    extern __shared__ unsigned char gm_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gm_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(base);
    uint64_t* empty_bar = full_bar + GSK_MAX_STAGES;
    uint64_t* tfull_bar = empty_bar + GSK_MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
    unsigned char* smem = base + 1024;
    const int stage_bytes = a_bytes + Cfg::B_BYTES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * GM_TILE, split = blockIdx.z;
    const int kb0 = split * kb_per_split;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < n_stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tfull_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    pdl_launch_dependents();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                        // everything above overlapped the previous kernel's tail

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t tx = (uint32_t)stage_bytes;
            int stage = 0; uint32_t phase = 0;
            for (int i = 0; i < kb_per_split; ++i) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                unsigned char* sa = smem + (size_t)stage * stage_bytes;
                mbar_expect_tx(&full_bar[stage], tx);
                const int k0 = (kb0 + i) * GM_KB;
                tma_load_2d(sa, &tmA, &full_bar[stage], k0, m0, TMA_EVICT_LAST);
                tma_load_2d(sa + a_bytes, &tmB, &full_bar[stage], k0, n0, b_evict_first ? TMA_EVICT_FIRST : TMA_EVICT_LAST);
                if (++stage == n_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_tf32(GM_TILE, BN);
        const uint64_t desc0 = umma_desc_sw128(smem_u32(smem));
        const uint32_t lo0 = (uint32_t)desc0;
        int stage = 0; uint32_t phase = 0;
        for (int i = 0; i < kb_per_split; ++i) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a_lo = lo0 + (uint32_t)((stage * stage_bytes) >> 4);
                const uint32_t b_lo = a_lo + (uint32_t)(a_bytes >> 4);
#pragma unroll
                for (int k = 0; k < GM_KB / 8; ++k) {
                    const uint64_t ad = (desc0 & 0xFFFFFFFF00000000ull) | (uint64_t)(a_lo + 2 * k);
                    const uint64_t bd = (desc0 & 0xFFFFFFFF00000000ull) | (uint64_t)(b_lo + 2 * k);
                    umma_tf32(tmem_base, ad, bd, idesc, (uint32_t)((i | k) != 0));
                }
                umma_commit(&empty_bar[stage]);
                if (i == kb_per_split - 1) umma_commit(tfull_bar);
            }
            __syncwarp();
            if (++stage == n_stages) { stage = 0; phase ^= 1; }
        }
    } else {
        const int lg = warp & 3;
        const int row = m0 + lg * 32 + lane;
        mbar_wait(tfull_bar, 0);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16);
        const bool direct = gridDim.z == 1;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld_x32(taddr + c0, v);
            tmem_wait_ld();
            if (row < M) {
                if (direct) {
                    float* crow = C + (size_t)row * N + n0 + c0;
                    const float* rrow = residual ? residual + (size_t)row * N + n0 + c0 : nullptr;
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        float o[8];
#pragma unroll
                        for (int t = 0; t < 8; t += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c0 + j + t));
                            o[t + 0] = __uint_as_float(v[j + t + 0]) + b4.x; o[t + 1] = __uint_as_float(v[j + t + 1]) + b4.y;
                            o[t + 2] = __uint_as_float(v[j + t + 2]) + b4.z; o[t + 3] = __uint_as_float(v[j + t + 3]) + b4.w;
                        }
                        if (act_gelu) {
#pragma unroll
                            for (int t = 0; t < 8; ++t) o[t] = gelu_erf(o[t]);
                        }
                        if (rrow) {
                            float r[8];
                            ld_global_v8(rrow + j, r);
#pragma unroll
                            for (int t = 0; t < 8; ++t) o[t] += r[t];
                        }
                        st_global_v8(crow + j, o);
                    }
                } else {
                    float* wrow = ws + ((size_t)split * M + row) * N + n0 + c0;
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        float o[8];
#pragma unroll
                        for (int t = 0; t < 8; ++t) o[t] = __uint_as_float(v[j + t]);
                        st_global_v8(wrow + j, o);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
    }
}

// --------------------------------------------------------------------------- fp32 row kernels
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per row: y = LayerNorm(x) * g + b   (two-pass variance like torch)
__device__ __forceinline__ void warp_layernorm_row(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ g,
                                                   const float* __restrict__ b, int d, float eps, int lane)
{
    float s = 0.f;
    for (int i = lane; i < d; i += 32) s += x[i];
    const float mean = warp_sum(s) / (float)d;
    float v = 0.f;
    for (int i = lane; i < d; i += 32) { float t = x[i] - mean; v += t * t; }
    const float rstd = rsqrtf(warp_sum(v) / (float)d + eps);
    for (int i = lane; i < d; i += 32) y[i] = (x[i] - mean) * rstd * g[i] + b[i];
}

__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ g, const float* __restrict__ b,
                 int rows, int d, float eps)
{
    pdl_launch_dependents();
    pdl_wait();
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= rows) return;
    warp_layernorm_row(x + (size_t)r * d, y + (size_t)r * d, g, b, d, eps, lane);
}

// embeddings: x[t] = LN(word[id] + pos[p] + type[0]); one warp per token; scratch row in shared memory
__global__ void __launch_bounds__(256)
embed_ln_kernel(const int32_t* __restrict__ tok, const int32_t* __restrict__ pos, const float* __restrict__ word,
                const float* __restrict__ pemb, const float* __restrict__ temb, const float* __restrict__ g,
                const float* __restrict__ b, float* __restrict__ x, int n_tok, int d, int vocab, float eps)
{
    extern __shared__ float e_sm[];   // [8][d]
    pdl_launch_dependents();
    pdl_wait();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * 8 + w;
    if (t >= n_tok) return;
    float* row = e_sm + (size_t)w * d;
    int id = tok[t];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const float* wr = word + (size_t)id * d;
    const float* pr = pemb + (size_t)pos[t] * d;
    for (int i = lane; i < d; i += 32) row[i] = (wr[i] + temb[i]) + pr[i];   // torch: inputs_embeds + token_type + position
    __syncwarp();
    warp_layernorm_row(row, x + (size_t)t * d, g, b, d, eps, lane);
}

// ------------------------------------------------------------------------------- attention
// packed variable-length sequences: qkv [n_tok, 3d] (Q | K | V, heads contiguous inside each), out [n_tok, d].
// grid (ceil(max_len / 16), heads, batch); 8 warps, 2 query rows per warp; keys processed in chunks of 64.
// split-K epilogue: out[row] = f(sum_s ws[s][row] + bias) (+GELU) (+residual), optionally followed by LayerNorm over
// the row (N <= 1024).  One 256-thread block per row, one float4 column group per thread: the `splits` partial loads
// of a thread are independent (all in flight at once) -- with a few rows this kernel is pure load latency.
// `residual` and `out` may alias (a thread reads its residual elements before it writes them).
constexpr int SK_MAX_SPLITS = 8;
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, int splits, int M, int N, const float* __restrict__ bias, const float* residual,
                     int act_gelu, const float* __restrict__ ln_g, const float* __restrict__ ln_b, float eps, float* out)
{
    __shared__ float s_red[2][8];
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t plane = (size_t)M * N;
    const float* w0 = ws + (size_t)row * N;
    auto element = [&](int c) {
        float4 p[SK_MAX_SPLITS];
#pragma unroll
        for (int sp = 0; sp < SK_MAX_SPLITS; ++sp)
            if (sp < splits) p[sp] = *reinterpret_cast<const float4*>(w0 + (size_t)sp * plane + c);
        float4 s = p[0];
#pragma unroll
        for (int sp = 1; sp < SK_MAX_SPLITS; ++sp)
            if (sp < splits) { s.x += p[sp].x; s.y += p[sp].y; s.z += p[sp].z; s.w += p[sp].w; }
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c));
        s.x += b4.x; s.y += b4.y; s.z += b4.z; s.w += b4.w;
        if (act_gelu) { s.x = gelu_erf(s.x); s.y = gelu_erf(s.y); s.z = gelu_erf(s.z); s.w = gelu_erf(s.w); }
        if (residual) {
            const float4 r = *reinterpret_cast<const float4*>(residual + (size_t)row * N + c);
            s.x += r.x; s.y += r.y; s.z += r.z; s.w += r.w;
        }
        return s;
    };
    if (ln_g == nullptr) {
        for (int c = tid * 4; c < N; c += 1024) *reinterpret_cast<float4*>(out + (size_t)row * N + c) = element(c);
        return;
    }
    const int c = tid * 4;                         // N <= 1024: at most one column group per thread
    const bool on = c < N;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (on) v = element(c);
    float sum = warp_sum((v.x + v.y) + (v.z + v.w));
    if (lane == 0) s_red[0][warp] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += s_red[0][w];
    const float mean = sum / (float)N;
    const float a = v.x - mean, b = v.y - mean, cc = v.z - mean, dd = v.w - mean;
    float var = on ? (a * a + b * b) + (cc * cc + dd * dd) : 0.f;
    var = warp_sum(var);
    if (lane == 0) s_red[1][warp] = var;
    __syncthreads();
    var = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) var += s_red[1][w];
    const float rstd = rsqrtf(var / (float)N + eps);
    if (on) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(ln_g + c)), be = __ldg(reinterpret_cast<const float4*>(ln_b + c));
        float4 o;
        o.x = a * rstd * g.x + be.x; o.y = b * rstd * g.y + be.y; o.z = cc * rstd * g.z + be.z; o.w = dd * rstd * g.w + be.w;
        *reinterpret_cast<float4*>(out + (size_t)row * N + c) = o;
    }
}

constexpr int AT_ROWS = 16;
constexpr int AT_CHUNK = 64;
__global__ void __launch_bounds__(256)
attention_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ seq_off, float* __restrict__ out, int d, int heads,
                 int max_len)
{
    extern __shared__ float a_sm[];
    const int dh = d / heads;
    const int b = blockIdx.z, h = blockIdx.y;
    const int t0 = seq_off[b], len = seq_off[b + 1] - t0;
    const int r0 = blockIdx.x * AT_ROWS;
    if (r0 >= len) return;
    float* s_q = a_sm;                                  // [AT_ROWS][dh]
    float* s_kT = s_q + AT_ROWS * dh;                   // [dh][AT_CHUNK + 1]
    float* s_v = s_kT + dh * (AT_CHUNK + 1);            // [AT_CHUNK][dh]
    float* s_p = s_v + AT_CHUNK * dh;                   // [AT_ROWS][max_len] scores / probabilities
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float scale = rsqrtf((float)dh);
    const int nrows = min(AT_ROWS, len - r0);
    for (int i = tid; i < nrows * dh; i += 256) {
        const int r = i / dh, c = i % dh;
        s_q[i] = qkv[(size_t)(t0 + r0 + r) * 3 * d + h * dh + c];
    }
    // pass 1: scores
    for (int k0 = 0; k0 < len; k0 += AT_CHUNK) {
        const int nk = min(AT_CHUNK, len - k0);
        __syncthreads();
        for (int i = tid; i < nk * dh; i += 256) {
            const int kk = i / dh, c = i % dh;
            s_kT[c * (AT_CHUNK + 1) + kk] = qkv[(size_t)(t0 + k0 + kk) * 3 * d + d + h * dh + c];
        }
        __syncthreads();
        for (int rr = 0; rr < 2; ++rr) {
            const int r = warp * 2 + rr;
            if (r >= nrows) break;
            for (int kk = lane; kk < nk; kk += 32) {
                float acc = 0.f;
                for (int c = 0; c < dh; ++c) acc = fmaf(s_q[r * dh + c], s_kT[c * (AT_CHUNK + 1) + kk], acc);
                s_p[r * max_len + k0 + kk] = acc * scale;
            }
        }
    }
    __syncthreads();
    // softmax per row (fp32, max-subtracted like torch)
    for (int rr = 0; rr < 2; ++rr) {
        const int r = warp * 2 + rr;
        if (r >= nrows) break;
        float* p = s_p + r * max_len;
        float m = -CUDART_INF_F;
        for (int k = lane; k < len; k += 32) m = fmaxf(m, p[k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = 0.f;
        for (int k = lane; k < len; k += 32) { float e = expf(p[k] - m); p[k] = e; s += e; }
        s = warp_sum(s);
        const float inv = 1.f / s;
        for (int k = lane; k < len; k += 32) p[k] *= inv;
    }
    // pass 2: out = P . V
    float o_acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};      // [row of the warp][dh / 32 columns per lane, dh <= 64]
    for (int k0 = 0; k0 < len; k0 += AT_CHUNK) {
        const int nk = min(AT_CHUNK, len - k0);
        __syncthreads();
        for (int i = tid; i < nk * dh; i += 256) {
            const int kk = i / dh, c = i % dh;
            s_v[kk * dh + c] = qkv[(size_t)(t0 + k0 + kk) * 3 * d + 2 * d + h * dh + c];
        }
        __syncthreads();
        for (int rr = 0; rr < 2; ++rr) {
            const int r = warp * 2 + rr;
            if (r >= nrows) break;
            for (int kk = 0; kk < nk; ++kk) {
                const float p = s_p[r * max_len + k0 + kk];
                if (lane < dh) o_acc[rr][0] = fmaf(p, s_v[kk * dh + lane], o_acc[rr][0]);
                if (lane + 32 < dh) o_acc[rr][1] = fmaf(p, s_v[kk * dh + lane + 32], o_acc[rr][1]);
            }
        }
    }
    for (int rr = 0; rr < 2; ++rr) {
        const int r = warp * 2 + rr;
        if (r >= nrows) break;
        float* orow = out + (size_t)(t0 + r0 + r) * d + h * dh;
        if (lane < dh) orow[lane] = o_acc[rr][0];
        if (lane + 32 < dh) orow[lane + 32] = o_acc[rr][1];
    }
}

// short sequences (queries; len <= S2 in {32, 64}): one CTA of 128 threads per (sequence, head).  Both products
// are register-tiled out of shared memory: thread (ty, tx) owns S2/8 rows x S2/16 score columns, then S2/8 rows x 4
// output columns; operands are stored transposed where needed so every shared-memory read is a conflict-free
// vector load.  fp32 throughout (scores, max-subtracted softmax, P.V) like torch's math path.
constexpr int ATS_MAX = 64;
template <int S2>
__global__ void __launch_bounds__(128)
attention_tiled_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ seq_off, float* __restrict__ out, int d, int heads)
{
    constexpr int RPT = S2 / 8;        // rows per thread
    constexpr int CPT = S2 / 16;       // score columns per thread
    constexpr int LD = S2 + 4;         // padded leading dimension (keeps 16-byte alignment, staggers banks)
    extern __shared__ __align__(16) float as_sm[];
    const int dh = d / heads;          // 32 or 64
    float* s_qT = as_sm;               // [dh][LD]   q transposed: s_qT[c][r]
    float* s_kT = s_qT + 64 * LD;      // [dh][LD]   k transposed: s_kT[c][key]
    float* s_v = s_kT + 64 * LD;       // [S2][64]   v: s_v[key][c]
    float* s_pT = s_v + S2 * 64;       // [S2][LD]   probabilities transposed: s_pT[key][r]
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.x, h = blockIdx.y;
    const int t0 = seq_off[b], len = seq_off[b + 1] - t0;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    {   // a warp moves 4 rows x 8 columns per step: 32-byte global segments, and with LD = S2 + 4 the
        // transposed stores (bank = 4 c + r) touch 32 distinct banks
        const int warp = tid >> 5, lane = tid & 31, cl = lane & 7, rl = lane >> 3;
        const int cblocks = dh / 8;
        for (int blk = warp; blk < cblocks * (S2 / 4); blk += 4) {
            const int c = (blk % cblocks) * 8 + cl, r = (blk / cblocks) * 4 + rl;
            float q = 0.f, k = 0.f, v = 0.f;
            if (r < len) {
                const float* base = qkv + (size_t)(t0 + r) * 3 * d + h * dh + c;
                q = base[0]; k = base[d]; v = base[2 * d];
            }
            s_qT[c * LD + r] = q; s_kT[c * LD + r] = k; s_v[r * 64 + c] = v;
        }
    }
    __syncthreads();
    // ---- scores
    float acc[RPT][CPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int j = 0; j < CPT; ++j) acc[i][j] = 0.f;
    for (int c = 0; c < dh; ++c) {
        float qv[RPT], kv[CPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) qv[i] = s_qT[c * LD + ty * RPT + i];
#pragma unroll
        for (int j = 0; j < CPT; ++j) kv[j] = s_kT[c * LD + tx * CPT + j];
#pragma unroll
        for (int i = 0; i < RPT; ++i)
#pragma unroll
            for (int j = 0; j < CPT; ++j) acc[i][j] = fmaf(qv[i], kv[j], acc[i][j]);
    }
    const float scale = rsqrtf((float)dh);
    // ---- softmax over each row: its S2 columns live in the 16 threads tx = 0..15 of one half-warp
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        float m = -CUDART_INF_F;
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            acc[i][j] = (tx * CPT + j < len) ? acc[i][j] * scale : -CUDART_INF_F;
            m = fmaxf(m, acc[i][j]);
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < CPT; ++j) { acc[i][j] = (tx * CPT + j < len) ? expf(acc[i][j] - m) : 0.f; sum += acc[i][j]; }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.f / sum;
#pragma unroll
        for (int j = 0; j < CPT; ++j) s_pT[(tx * CPT + j) * LD + ty * RPT + i] = acc[i][j] * inv;
    }
    __syncthreads();
    // ---- out = P . V : thread (ty, tx) owns rows ty*RPT.. and columns tx*4.. (dh = 64) / tx*2.. (dh = 32)
    const int cw = dh / 16;            // 4 or 2 output columns per thread
    float o_acc[RPT][4];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o_acc[i][j] = 0.f;
    for (int k = 0; k < len; ++k) {
        float pv[RPT], vv[4];
#pragma unroll
        for (int i = 0; i < RPT; ++i) pv[i] = s_pT[k * LD + ty * RPT + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) vv[j] = j < cw ? s_v[k * 64 + tx * cw + j] : 0.f;
#pragma unroll
        for (int i = 0; i < RPT; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) o_acc[i][j] = fmaf(pv[i], vv[j], o_acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int r = ty * RPT + i;
        if (r < len) {
            float* orow = out + (size_t)(t0 + r) * d + h * dh + tx * cw;
#pragma unroll
            for (int j = 0; j < 4; ++j) if (j < cw) orow[j] = o_acc[i][j];
        }
    }
}

// long sequences (chunks on the /index path; len up to max_position): flash-style fp32 attention.
// One CTA of 128 threads per (sequence, head, block of 64 query rows); keys are visited in chunks of 64 with an
// online (running max / running sum) softmax, so no S x S score matrix is materialised.  Both products are
// register-tiled exactly like attention_tiled_kernel<64>: thread (ty, tx) owns 8 rows x 4 score columns and
// 8 rows x (dh/16) output columns.
__global__ void __launch_bounds__(128)
attention_flash_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ seq_off, float* __restrict__ out, int d, int heads)
{
    constexpr int S2 = 64, RPT = 8, CPT = 4, LD = S2 + 4;
    extern __shared__ __align__(16) float af_sm[];
    const int dh = d / heads;
    float* s_qT = af_sm;               // [dh][LD]
    float* s_kT = s_qT + 64 * LD;      // [dh][LD]
    float* s_v = s_kT + 64 * LD;       // [64][64]
    float* s_pT = s_v + S2 * 64;       // [64 keys][LD rows]
    const int b = blockIdx.z, h = blockIdx.y;
    const int t0 = seq_off[b], len = seq_off[b + 1] - t0;
    const int r0 = blockIdx.x * S2;
    if (r0 >= len) return;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int warp = tid >> 5, lane = tid & 31, cl = lane & 7, rl = lane >> 3;
    const int cblocks = dh / 8;
    for (int blk = warp; blk < cblocks * (S2 / 4); blk += 4) {       // Q rows of this block, transposed
        const int c = (blk % cblocks) * 8 + cl, r = (blk / cblocks) * 4 + rl;
        s_qT[c * LD + r] = (r0 + r < len) ? qkv[(size_t)(t0 + r0 + r) * 3 * d + h * dh + c] : 0.f;
    }
    const float scale = rsqrtf((float)dh);
    const int cw = dh / 16;
    float m_run[RPT], l_run[RPT], o_acc[RPT][4];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        m_run[i] = -CUDART_INF_F; l_run[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) o_acc[i][j] = 0.f;
    }
    for (int k0 = 0; k0 < len; k0 += S2) {
        __syncthreads();                                             // previous chunk fully consumed (s_kT, s_v, s_pT)
        for (int blk = warp; blk < cblocks * (S2 / 4); blk += 4) {
            const int c = (blk % cblocks) * 8 + cl, r = (blk / cblocks) * 4 + rl;
            float k = 0.f, v = 0.f;
            if (k0 + r < len) {
                const float* base = qkv + (size_t)(t0 + k0 + r) * 3 * d + d + h * dh + c;
                k = base[0]; v = base[d];
            }
            s_kT[c * LD + r] = k; s_v[r * 64 + c] = v;
        }
        __syncthreads();
        float acc[RPT][CPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i)
#pragma unroll
            for (int j = 0; j < CPT; ++j) acc[i][j] = 0.f;
        for (int c = 0; c < dh; ++c) {
            float qv[RPT], kv[CPT];
#pragma unroll
            for (int i = 0; i < RPT; ++i) qv[i] = s_qT[c * LD + ty * RPT + i];
#pragma unroll
            for (int j = 0; j < CPT; ++j) kv[j] = s_kT[c * LD + tx * CPT + j];
#pragma unroll
            for (int i = 0; i < RPT; ++i)
#pragma unroll
                for (int j = 0; j < CPT; ++j) acc[i][j] = fmaf(qv[i], kv[j], acc[i][j]);
        }
        // online softmax: a row's 64 chunk columns live in the 16 threads tx = 0..15 of one half-warp
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            float m = -CUDART_INF_F;
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                acc[i][j] = (k0 + tx * CPT + j < len) ? acc[i][j] * scale : -CUDART_INF_F;
                m = fmaxf(m, acc[i][j]);
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            const float m_new = fmaxf(m_run[i], m);                  // finite: every chunk holds at least one valid key
            const float corr = expf(m_run[i] - m_new);               // exp(-inf) = 0 on the first chunk
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const float p = (k0 + tx * CPT + j < len) ? expf(acc[i][j] - m_new) : 0.f;
                sum += p;
                s_pT[(tx * CPT + j) * LD + ty * RPT + i] = p;
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            l_run[i] = l_run[i] * corr + sum;
            m_run[i] = m_new;
#pragma unroll
            for (int j = 0; j < 4; ++j) o_acc[i][j] *= corr;
        }
        __syncthreads();
        const int nk = min(S2, len - k0);
        for (int k = 0; k < nk; ++k) {
            float pv[RPT], vv[4];
#pragma unroll
            for (int i = 0; i < RPT; ++i) pv[i] = s_pT[k * LD + ty * RPT + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) vv[j] = j < cw ? s_v[k * 64 + tx * cw + j] : 0.f;
#pragma unroll
            for (int i = 0; i < RPT; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) o_acc[i][j] = fmaf(pv[i], vv[j], o_acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int r = r0 + ty * RPT + i;
        if (r < len) {
            const float inv = 1.f / l_run[i];
            float* orow = out + (size_t)(t0 + r) * d + h * dh + tx * cw;
#pragma unroll
            for (int j = 0; j < 4; ++j) if (j < cw) orow[j] = o_acc[i][j] * inv;
        }
    }
}

// CLS pooling + L2 normalisation (F.normalize, eps 1e-12): one warp per sequence
__global__ void __launch_bounds__(256)
cls_normalize_kernel(const float* __restrict__ x, const int32_t* __restrict__ seq_off, float* __restrict__ out, int batch, int d,
                     int ld_out)
{
    pdl_launch_dependents();
    pdl_wait();
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= batch) return;
    const float* row = x + (size_t)seq_off[b] * d;
    float s = 0.f;
    for (int i = lane; i < d; i += 32) s = fmaf(row[i], row[i], s);
    const float n = fmaxf(sqrtf(warp_sum(s)), 1e-12f);
    for (int i = lane; i < d; i += 32) out[(size_t)b * ld_out + i] = row[i] / n;
}

// ------------------------------------------------------------------------------------ host
static bool pdl_enabled()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("KRAG_PDL"); v = (e == nullptr || e[0] != '0') ? 1 : 0; }
    return v == 1;
}
// launch with programmatic stream serialization (see pdl_wait in tc_ptx.cuh); only for kernels that call pdl_wait()
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    KRAG_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
    count_launch();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn emb_encode()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
static void emb_map(CUtensorMap* tm, const float* base, int rows, int cols, int box_rows = GM_TILE)
{
    EncodeTiledFn enc = emb_encode();
    if (!enc) throw CudaError{cudaErrorUnknown, "cuTensorMapEncodeTiled unavailable", __FILE__, __LINE__};
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {(cuuint32_t)GM_KB, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    if (enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        throw CudaError{cudaErrorUnknown, "cuTensorMapEncodeTiled failed", __FILE__, __LINE__};
}

void launch_gemm_tf32(const DeviceInfo& di, const float* A, const float* B, int M, int N, int K, const float* bias,
                      const float* residual, bool gelu, float* C, cudaStream_t st)
{
    static bool attr_set = false;
    if (!attr_set) {
        KRAG_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GM_SMEM));
        attr_set = true;
    }
    CUtensorMap tmA, tmB;
    emb_map(&tmA, A, M, K);
    static int use2 = -1;
    if (use2 < 0) { const char* ev = getenv("KRAG_GEMM_2CTA"); use2 = (ev == nullptr || ev[0] != '0') ? 1 : 0; }
    // CTA pairs (256 x BN tiles) win when there are enough tiles to fill the 74 pairs; small problems (few query
    // tokens per rank) are latency-bound by the K loop of a single tile, where 128 x 128 tiles on single CTAs
    // halve the per-tile MMA time and quadruple the number of CTAs
    const int pair_tiles = ((M + 2 * GM_TILE - 1) / (2 * GM_TILE)) * (N / ((N % 256 == 0) ? 256 : 128));
    if (use2 && M > GM_TILE && pair_tiles >= di.sm_count / 4) {
        const int BN = (N % 256 == 0) ? 256 : 128;
        emb_map(&tmB, B, N, K, BN / 2);
        const int total2 = ((M + 2 * GM_TILE - 1) / (2 * GM_TILE)) * (N / BN);
        const int max_pairs = di.sm_count / 2;
        const int n_pairs = total2 < max_pairs ? total2 : max_pairs;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(2 * n_pairs));
        cfg.blockDim = dim3(GM2_THREADS);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        const int gelu_i = gelu ? 1 : 0;
        if (BN == 256) {
            static bool a256 = false;
            if (!a256) { KRAG_CUDA(cudaFuncSetAttribute(gemm2_tf32_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Gm2Cfg<256>::SMEM)); a256 = true; }
            cfg.dynamicSmemBytes = Gm2Cfg<256>::SMEM;
            KRAG_CUDA(cudaLaunchKernelEx(&cfg, gemm2_tf32_kernel<256>, tmA, tmB, M, N, K, bias, residual, gelu_i, C));
        } else {
            static bool a128 = false;
            if (!a128) { KRAG_CUDA(cudaFuncSetAttribute(gemm2_tf32_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Gm2Cfg<128>::SMEM)); a128 = true; }
            cfg.dynamicSmemBytes = Gm2Cfg<128>::SMEM;
            KRAG_CUDA(cudaLaunchKernelEx(&cfg, gemm2_tf32_kernel<128>, tmA, tmB, M, N, K, bias, residual, gelu_i, C));
        }
        count_launch();
        return;
    }
    emb_map(&tmB, B, N, K);
    const int total = ((M + GM_TILE - 1) / GM_TILE) * (N / GM_TILE);
    const int grid = total < di.sm_count ? total : di.sm_count;
    gemm_tf32_kernel<<<grid, GM_THREADS, GM_SMEM, st>>>(tmA, tmB, M, N, K, bias, residual, gelu ? 1 : 0, C);
    KRAG_CUDA(cudaGetLastError());
    count_launch();
}


// small-M path: 128 x BN tiles x split-K slices, ~one CTA per SM (see gemm_sk_tf32_kernel)
template <int BN>
static void launch_gemm_sk(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int m_tiles, int splits, int kb_per_split, int a_rows,
                           const float* bias, const float* residual, bool gelu, float* C, float* ws, cudaStream_t st)
{
    static bool attr = false;
    if (!attr) { KRAG_CUDA(cudaFuncSetAttribute(gemm_sk_tf32_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); attr = true; }
    const int a_bytes = a_rows * GM_KB * 4, stage_bytes = a_bytes + GskCfg<BN>::B_BYTES;
    // a few activation rows: keep a CTA under half an SM's shared memory (two CTAs per SM, the next kernel's CTAs
    // can become resident while this one drains); full 128-row tiles take the whole SM
    const int budget = (a_rows <= 32 ? 96 : 196) * 1024;
    int n_stages = budget / stage_bytes;
    n_stages = n_stages > GSK_MAX_STAGES ? GSK_MAX_STAGES : n_stages;
    n_stages = n_stages > kb_per_split ? kb_per_split : n_stages;
    launch_pdl(gemm_sk_tf32_kernel<BN>, dim3((unsigned)(N / BN), (unsigned)m_tiles, (unsigned)splits), dim3(GM_THREADS),
               gsk_smem_bytes(n_stages, stage_bytes), st, tmA, tmB, M, N, kb_per_split, a_bytes, n_stages, m_tiles == 1 ? 1 : 0, bias, residual,
               gelu ? 1 : 0, C, ws);
}

void launch_linear(const DeviceInfo& di, const float* A, const float* B, int M, int N, int K, const float* bias, const float* residual,
                   bool gelu, float* C, const float* ln_g, const float* ln_b, float eps, float* Y, float* ws, size_t ws_floats,
                   cudaStream_t st)
{
    static int use_sk = -1;
    if (use_sk < 0) { const char* ev = getenv("KRAG_GEMM_SPLITK"); use_sk = (ev == nullptr || ev[0] != '0') ? 1 : 0; }
    const int m_tiles = (M + GM_TILE - 1) / GM_TILE;
    const int tiles128 = m_tiles * (N / GM_TILE);
    if (use_sk && ws && tiles128 < di.sm_count / 2 && N % 64 == 0 && K % GM_KB == 0 && (!ln_g || N <= 1024)) {
        // long-K layers over several activation tiles (FFN2 for tens of queries): 128 x 128 tiles halve the operand
        // re-reads through L2 (the binding resource with 4-byte operands); split-K restores the CTA count
        static int use_wide = -1;
        if (use_wide < 0) { const char* ev = getenv("KRAG_SK_WIDE"); use_wide = (ev == nullptr || ev[0] != '0') ? 1 : 0; }
        const int BN = m_tiles == 1 ? 32 : ((use_wide && K >= 2048 && N % 128 == 0) ? 128 : 64);
        const int ctas = m_tiles * (N / BN), kblocks = K / GM_KB;
        int splits = 1;
        if (ctas <= di.sm_count / 3) {              // 72 CTAs with the whole K range beat 144 + a reduce kernel
            const int want = di.sm_count / ctas < SK_MAX_SPLITS ? di.sm_count / ctas : SK_MAX_SPLITS;
            for (int s = want; s >= 2; --s)
                if (kblocks % s == 0 && kblocks / s >= 4 && (size_t)s * M * N <= ws_floats) { splits = s; break; }
        }
        const int a_rows = m_tiles == 1 ? ((M + 7) / 8) * 8 : GM_TILE;
        CUtensorMap tmA, tmB;
        emb_map(&tmA, A, M, K, a_rows);
        emb_map(&tmB, B, N, K, BN);
        float* direct_out = C;
        if (BN == 32) launch_gemm_sk<32>(tmA, tmB, M, N, m_tiles, splits, kblocks / splits, a_rows, bias, residual, gelu, direct_out, ws, st);
        else if (BN == 128) launch_gemm_sk<128>(tmA, tmB, M, N, m_tiles, splits, kblocks / splits, a_rows, bias, residual, gelu, direct_out, ws, st);
        else launch_gemm_sk<64>(tmA, tmB, M, N, m_tiles, splits, kblocks / splits, a_rows, bias, residual, gelu, direct_out, ws, st);
        if (splits > 1) {
            launch_pdl(splitk_reduce_kernel, dim3((unsigned)M), dim3(256), 0, st, ws, splits, M, N, bias, residual, gelu ? 1 : 0, ln_g, ln_b, eps,
                       ln_g ? Y : C);
            return;
        }
    } else {
        launch_gemm_tf32(di, A, B, M, N, K, bias, residual, gelu, C, st);
    }
    if (ln_g) {
        launch_pdl(layernorm_kernel, dim3((unsigned)((M * 32 + 255) / 256)), dim3(256), 0, st, C, Y, ln_g, ln_b, M, N, eps);
    }
}


struct Embedder {
    DeviceInfo di;
    BertConfig cfg;
    std::map<std::string, float*> t;      // HF tensor name -> device copy
    std::vector<float*> wqkv, bqkv;       // fused per layer at finalize
    bool finalized = false;
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    std::map<std::tuple<int, int, int>, cudaGraphExec_t> graphs;   // captured forward per (n_tok, batch, max_len)
    std::map<std::tuple<int, int, int>, int> seen;
    // workspaces, grown on demand
    int cap_tok = 0, cap_batch = 0;
    int32_t *d_tok = nullptr, *d_pos = nullptr, *d_off = nullptr;
    float *x = nullptr, *x2 = nullptr, *qkv = nullptr, *ctx = nullptr, *ffn = nullptr, *out = nullptr;
    float* ws = nullptr;                  // split-K partials (small token counts)
    static constexpr size_t WS_FLOATS = (size_t)4 << 20;
};

static float* emb_get(Embedder* e, const std::string& name, int64_t n)
{
    auto it = e->t.find(name);
    if (it == e->t.end()) throw std::runtime_error("embedder: tensor not loaded: " + name);
    (void)n;
    return it->second;
}
static std::string lname(int l, const char* s) { return "encoder.layer." + std::to_string(l) + "." + s; }

Embedder* embedder_create(const DeviceInfo& di, const BertConfig& cfg)
{
    if (cfg.hidden % 128 || cfg.inter % 128 || cfg.hidden % cfg.heads || cfg.hidden / cfg.heads > 64 || (cfg.hidden / cfg.heads) % 32)
        throw std::runtime_error("embedder: hidden/intermediate must be multiples of 128 and head_dim 32 or 64");
    Embedder* e = new Embedder();
    e->di = di; e->cfg = cfg;
    KRAG_CUDA(cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking));
    return e;
}

void embedder_load(Embedder* e, const char* name, const float* data, int64_t n)
{
    float* p = nullptr;
    KRAG_CUDA(cudaMalloc(&p, sizeof(float) * (size_t)n));
    KRAG_CUDA(cudaMemcpy(p, data, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice));
    auto it = e->t.find(name);
    if (it != e->t.end()) cudaFree(it->second);
    e->t[name] = p;
    e->finalized = false;
}

void embedder_finalize(Embedder* e)
{
    const int d = e->cfg.hidden;
    for (float* p : e->wqkv) cudaFree(p);
    for (float* p : e->bqkv) cudaFree(p);
    e->wqkv.clear(); e->bqkv.clear();
    emb_get(e, "embeddings.word_embeddings.weight", 0); emb_get(e, "embeddings.position_embeddings.weight", 0);
    emb_get(e, "embeddings.token_type_embeddings.weight", 0); emb_get(e, "embeddings.LayerNorm.weight", 0);
    emb_get(e, "embeddings.LayerNorm.bias", 0);
    for (int l = 0; l < e->cfg.layers; ++l) {
        float *w = nullptr, *b = nullptr;
        KRAG_CUDA(cudaMalloc(&w, sizeof(float) * (size_t)3 * d * d));
        KRAG_CUDA(cudaMalloc(&b, sizeof(float) * (size_t)3 * d));
        const char* names[3] = {"attention.self.query", "attention.self.key", "attention.self.value"};
        for (int j = 0; j < 3; ++j) {
            KRAG_CUDA(cudaMemcpy(w + (size_t)j * d * d, emb_get(e, lname(l, names[j]) + ".weight", 0), sizeof(float) * (size_t)d * d, cudaMemcpyDeviceToDevice));
            KRAG_CUDA(cudaMemcpy(b + (size_t)j * d, emb_get(e, lname(l, names[j]) + ".bias", 0), sizeof(float) * (size_t)d, cudaMemcpyDeviceToDevice));
        }
        e->wqkv.push_back(w); e->bqkv.push_back(b);
        for (const char* s : {"attention.output.dense.weight", "attention.output.dense.bias", "attention.output.LayerNorm.weight",
                              "attention.output.LayerNorm.bias", "intermediate.dense.weight", "intermediate.dense.bias",
                              "output.dense.weight", "output.dense.bias", "output.LayerNorm.weight", "output.LayerNorm.bias"})
            emb_get(e, lname(l, s), 0);
    }
    if (!e->ws) KRAG_CUDA(cudaMalloc(&e->ws, sizeof(float) * Embedder::WS_FLOATS));
    e->finalized = true;
}

void embedder_destroy(Embedder* e)
{
    for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second);
    for (auto& kv : e->t) cudaFree(kv.second);
    for (float* p : e->wqkv) cudaFree(p);
    for (float* p : e->bqkv) cudaFree(p);
    for (void* p : {(void*)e->d_tok, (void*)e->d_pos, (void*)e->d_off, (void*)e->x, (void*)e->x2, (void*)e->qkv, (void*)e->ctx, (void*)e->ffn, (void*)e->out})
        if (p) cudaFree(p);
    if (e->ws) cudaFree(e->ws);
    if (e->st) cudaStreamDestroy(e->st);
    delete e;
}

int embedder_hidden(const Embedder* e) { return e->cfg.hidden; }

// tok_ids: packed tokens of all sequences; tok_offsets [batch+1]; out_host [batch, hidden]
// out_host != nullptr: embeddings are copied to the host and the call returns synchronised.
// out_dev  != nullptr: embeddings are written to device rows of stride ld_out floats and `consumer` (a stream of
//                      the caller) is made to wait for them -- no host round trip.
void embedder_forward(Embedder* e, int batch, const int32_t* tok_ids, const int32_t* tok_offsets, float* out_host,
                      float* out_dev, int ld_out, cudaStream_t consumer)
{
    if (!e->finalized) throw std::runtime_error("embedder: call finalize after loading the weights");
    const BertConfig& c = e->cfg;
    const int d = c.hidden, n_tok = tok_offsets[batch];
    cudaStream_t st = e->st;
    int max_len = 0;
    std::vector<int32_t> pos((size_t)n_tok);
    for (int b = 0; b < batch; ++b) {
        const int len = tok_offsets[b + 1] - tok_offsets[b];
        if (len < 1 || len > c.max_pos) throw std::runtime_error("embedder: sequence length must be in [1, max_position]");
        max_len = len > max_len ? len : max_len;
        for (int i = 0; i < len; ++i) pos[(size_t)tok_offsets[b] + i] = i;
    }
    if (n_tok > e->cap_tok || batch > e->cap_batch) {
        for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second);   // captured pointers are about to change
        e->graphs.clear(); e->seen.clear();
        KRAG_CUDA(cudaStreamSynchronize(st));
        for (void* p : {(void*)e->d_tok, (void*)e->d_pos, (void*)e->d_off, (void*)e->x, (void*)e->x2, (void*)e->qkv, (void*)e->ctx, (void*)e->ffn, (void*)e->out})
            if (p) cudaFree(p);
        const size_t T = (size_t)(n_tok > e->cap_tok ? n_tok : e->cap_tok), Bc = (size_t)(batch > e->cap_batch ? batch : e->cap_batch);
        KRAG_CUDA(cudaMalloc(&e->d_tok, 4 * T)); KRAG_CUDA(cudaMalloc(&e->d_pos, 4 * T)); KRAG_CUDA(cudaMalloc(&e->d_off, 4 * (Bc + 1)));
        KRAG_CUDA(cudaMalloc(&e->x, 4 * T * d)); KRAG_CUDA(cudaMalloc(&e->x2, 4 * T * d)); KRAG_CUDA(cudaMalloc(&e->qkv, 4 * T * 3 * d));
        KRAG_CUDA(cudaMalloc(&e->ctx, 4 * T * d)); KRAG_CUDA(cudaMalloc(&e->ffn, 4 * T * c.inter)); KRAG_CUDA(cudaMalloc(&e->out, 4 * Bc * d));
        e->cap_tok = (int)T; e->cap_batch = (int)Bc;
    }
    KRAG_CUDA(cudaMemcpyAsync(e->d_tok, tok_ids, 4 * (size_t)n_tok, cudaMemcpyHostToDevice, st));
    KRAG_CUDA(cudaMemcpyAsync(e->d_pos, pos.data(), 4 * (size_t)n_tok, cudaMemcpyHostToDevice, st));
    KRAG_CUDA(cudaMemcpyAsync(e->d_off, tok_offsets, 4 * (size_t)(batch + 1), cudaMemcpyHostToDevice, st));

    // The forward is ~7 kernels per layer; for a handful of query tokens it is bound by launch gaps and kernel
    // prologues, so each (n_tok, batch, max_len) shape is captured into a CUDA graph the second time it is seen.
    auto run_layers = [&]() {
    launch_pdl(embed_ln_kernel, dim3((unsigned)((n_tok + 7) / 8)), dim3(256), (size_t)8 * d * 4, st, e->d_tok, e->d_pos,
               e->t["embeddings.word_embeddings.weight"], e->t["embeddings.position_embeddings.weight"],
               e->t["embeddings.token_type_embeddings.weight"], e->t["embeddings.LayerNorm.weight"], e->t["embeddings.LayerNorm.bias"], e->x,
               n_tok, d, c.vocab, c.eps);
    const int dh = d / c.heads;
    const size_t at_smem = sizeof(float) * ((size_t)AT_ROWS * dh + (size_t)dh * (AT_CHUNK + 1) + (size_t)AT_CHUNK * dh + (size_t)AT_ROWS * max_len);
    const size_t ats_smem32 = sizeof(float) * ((size_t)2 * 64 * (32 + 4) + (size_t)32 * 64 + (size_t)32 * (32 + 4));
    const size_t ats_smem64 = sizeof(float) * ((size_t)2 * 64 * (64 + 4) + (size_t)64 * 64 + (size_t)64 * (64 + 4));
    static bool at_attr = false;
    if (!at_attr) {
        KRAG_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        KRAG_CUDA(cudaFuncSetAttribute(attention_tiled_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
        KRAG_CUDA(cudaFuncSetAttribute(attention_tiled_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
        KRAG_CUDA(cudaFuncSetAttribute(attention_flash_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
        at_attr = true;
    }
    static int use_flash = -1;
    if (use_flash < 0) { const char* ev = getenv("KRAG_ATTN_FLASH"); use_flash = (ev == nullptr || ev[0] != '0') ? 1 : 0; }
    for (int l = 0; l < c.layers; ++l) {
        launch_linear(e->di, e->x, e->wqkv[(size_t)l], n_tok, 3 * d, d, e->bqkv[(size_t)l], nullptr, false, e->qkv, nullptr, nullptr, 0.f,
                      nullptr, e->ws, Embedder::WS_FLOATS, st);
        if (max_len <= 32) {
            launch_pdl(attention_tiled_kernel<32>, dim3((unsigned)batch, (unsigned)c.heads), dim3(128), ats_smem32, st, e->qkv, e->d_off, e->ctx, d, c.heads);
        } else if (max_len <= ATS_MAX) {
            launch_pdl(attention_tiled_kernel<64>, dim3((unsigned)batch, (unsigned)c.heads), dim3(128), ats_smem64, st, e->qkv, e->d_off, e->ctx, d, c.heads);
        } else {
            if (use_flash)
                attention_flash_kernel<<<dim3((unsigned)((max_len + 63) / 64), (unsigned)c.heads, (unsigned)batch), 128, ats_smem64, st>>>(
                    e->qkv, e->d_off, e->ctx, d, c.heads);
            else
                attention_kernel<<<dim3((unsigned)((max_len + AT_ROWS - 1) / AT_ROWS), (unsigned)c.heads, (unsigned)batch), 256, at_smem, st>>>(
                    e->qkv, e->d_off, e->ctx, d, c.heads, max_len);
            KRAG_CUDA(cudaGetLastError());
            count_launch();
        }
        // O projection (+bias +residual) -> LayerNorm; FFN1 (+bias, GELU); FFN2 (+bias +residual) -> LayerNorm
        launch_linear(e->di, e->ctx, e->t[lname(l, "attention.output.dense.weight")], n_tok, d, d, e->t[lname(l, "attention.output.dense.bias")],
                      e->x, false, e->x2, e->t[lname(l, "attention.output.LayerNorm.weight")], e->t[lname(l, "attention.output.LayerNorm.bias")],
                      c.eps, e->x, e->ws, Embedder::WS_FLOATS, st);
        launch_linear(e->di, e->x, e->t[lname(l, "intermediate.dense.weight")], n_tok, c.inter, d, e->t[lname(l, "intermediate.dense.bias")],
                      nullptr, true, e->ffn, nullptr, nullptr, 0.f, nullptr, e->ws, Embedder::WS_FLOATS, st);
        launch_linear(e->di, e->ffn, e->t[lname(l, "output.dense.weight")], n_tok, d, c.inter, e->t[lname(l, "output.dense.bias")], e->x, false,
                      e->x2, e->t[lname(l, "output.LayerNorm.weight")], e->t[lname(l, "output.LayerNorm.bias")], c.eps, e->x, e->ws,
                      Embedder::WS_FLOATS, st);
    }
    };   // run_layers
    static int use_graph = -1;
    if (use_graph < 0) { const char* ev = getenv("KRAG_EMBED_GRAPH"); use_graph = (ev == nullptr || ev[0] != '0') ? 1 : 0; }
    const std::tuple<int, int, int> key(n_tok, batch, max_len);
    auto git = e->graphs.find(key);
    if (use_graph && git != e->graphs.end()) {
        KRAG_CUDA(cudaGraphLaunch(git->second, st));
        count_launch(1 + 7 * c.layers);
    } else if (use_graph && e->seen[key]++ >= 1) {
        if (e->graphs.size() >= 64) { for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second); e->graphs.clear(); }
        cudaGraph_t g = nullptr;
        cudaGraphExec_t ex = nullptr;
        KRAG_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        try { run_layers(); } catch (...) { cudaStreamEndCapture(st, &g); if (g) cudaGraphDestroy(g); throw; }
        KRAG_CUDA(cudaStreamEndCapture(st, &g));
        KRAG_CUDA(cudaGraphInstantiate(&ex, g, 0));
        KRAG_CUDA(cudaGraphDestroy(g));
        e->graphs[key] = ex;
        KRAG_CUDA(cudaGraphLaunch(ex, st));
    } else {
        run_layers();
    }
    if (out_dev) {
        cls_normalize_kernel<<<(batch * 32 + 255) / 256, 256, 0, st>>>(e->x, e->d_off, out_dev, batch, d, ld_out);
        KRAG_CUDA(cudaGetLastError());
        count_launch();
        if (!e->done) KRAG_CUDA(cudaEventCreateWithFlags(&e->done, cudaEventDisableTiming));
        KRAG_CUDA(cudaEventRecord(e->done, st));
        KRAG_CUDA(cudaStreamWaitEvent(consumer, e->done, 0));
    }
    if (out_host) {
        cls_normalize_kernel<<<(batch * 32 + 255) / 256, 256, 0, st>>>(e->x, e->d_off, e->out, batch, d, d);
        KRAG_CUDA(cudaGetLastError());
        count_launch();
        KRAG_CUDA(cudaMemcpyAsync(out_host, e->out, 4 * (size_t)batch * d, cudaMemcpyDeviceToHost, st));
        KRAG_CUDA(cudaStreamSynchronize(st));
    }
}

}  // namespace krag
