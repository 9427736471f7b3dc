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
#include <cuda_fp16.h>
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
// A [M, K] (activations), B [N, K] (nn.Linear weight), C [M, N], row-major.  The tensor cores see fp16 operands, the result
// is fp32-accurate: every fp32 value v is carried as TWO fp16 numbers
//     hi = fp16(v),   lo = fp16((v - hi) * 2^11)          v = hi + lo * 2^-11  to ~2^-22 relative
// ("split" operands: producers write them next to / instead of the fp32 tensor, weights are split once at load time), and
//     A . B^T  =  hi_A . hi_B^T  +  2^-11 (hi_A . lo_B^T + lo_A . hi_B^T)          (lo . lo term: 2^-22, dropped)
// is three kind::f16 MMAs per K = 16 step into two TMEM accumulators (main term / correction terms, combined in the
// epilogue; the scaling keeps the lo parts out of the fp16 subnormal range and the small terms out of the big
// accumulator's rounding).  fp16 products are exact in fp32, so what is left is the accumulation order -- the same kind of
// error an fp32 SIMT GEMM has.  Operand bytes per element are those of fp32 (2 + 2), the tensor work is 1.5x a TF32 GEMM's:
// with these layer shapes the kernels stay bound by L2 -> shared-memory operand traffic, as the TF32 kernels they replace
// were, but now meet the reference's fp32 arithmetic (TF32 carried 10 mantissa bits: ~5e-3 on the unit-norm embedding).
// Persistent CTAs over 128 x 128 output tiles (n fastest so neighbouring CTAs share the A tile in L2); one TMA warp,
// one MMA-issuing thread, four epilogue warps; stage = one k-block of 64 halfs of hi_A, lo_A, hi_B, lo_B.
constexpr int GM_TILE = 128;
constexpr int GH_KB = 64;                   // halfs per k-block (128-byte swizzle row)
constexpr int GM_SLAB = GM_TILE * 128;      // 16 KB: 128 rows x 128 B
constexpr int GM_STAGE_BYTES = 4 * GM_SLAB; // 64 KB
constexpr int GM_STAGES = 3;
constexpr int GM_THREADS = 192;
constexpr size_t GM_SMEM = (size_t)GM_STAGES * GM_STAGE_BYTES + 1024 + 256;
constexpr float GH_LO_SCALE = 2048.f, GH_LO_INV = 1.f / 2048.f;

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

// ---- split fp16 operands
__device__ __forceinline__ uint16_t f2h_sat(float v)      // round to nearest even, finite saturation
{
    uint16_t h;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
    return h;
}
__device__ __forceinline__ float h2f(uint16_t h)
{
    float f;
    asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"(h));
    return f;
}
__device__ __forceinline__ void split_f16(float v, uint16_t& hi, uint16_t& lo)
{
    hi = f2h_sat(v);
    lo = f2h_sat((v - h2f(hi)) * GH_LO_SCALE);
}
// 8 consecutive elements -> one 16-byte store into each plane
__device__ __forceinline__ void split_store8(const float* o, uint16_t* p_hi, uint16_t* p_lo)
{
    uint32_t a[4], b[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        uint16_t h0, l0, h1, l1;
        split_f16(o[2 * t], h0, l0); split_f16(o[2 * t + 1], h1, l1);
        a[t] = (uint32_t)h0 | ((uint32_t)h1 << 16); b[t] = (uint32_t)l0 | ((uint32_t)l1 << 16);
    }
    *reinterpret_cast<uint4*>(p_hi) = make_uint4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<uint4*>(p_lo) = make_uint4(b[0], b[1], b[2], b[3]);
}
__global__ void __launch_bounds__(256)
split_f16_kernel(const float* __restrict__ in, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int64_t n8)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        float o[8];
        const float4 x = reinterpret_cast<const float4*>(in)[2 * i], y = reinterpret_cast<const float4*>(in)[2 * i + 1];
        o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w; o[4] = y.x; o[5] = y.y; o[6] = y.z; o[7] = y.w;
        split_store8(o, hi + 8 * i, lo + 8 * i);
    }
}
void launch_split_f16(const float* in, uint16_t* hi, uint16_t* lo, int64_t n, cudaStream_t st)   // n % 8 == 0
{
    const int64_t n8 = n / 8;
    const int64_t want = (n8 + 255) / 256;
    split_f16_kernel<<<(unsigned)(want < 148 * 16 ? (want > 0 ? want : 1) : 148 * 16), 256, 0, st>>>(in, hi, lo, n8);
    KRAG_CUDA(cudaGetLastError());
    count_launch();
}

// instruction descriptor for kind::f16 with fp16 operands (format 0), fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N)
{
    return (1u << 4) /* D f32 */ | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// what a GEMM epilogue writes: fp32 C and / or the split planes of C (operand of the next GEMM)
struct GemmOut {
    float* c;            // [M, N] fp32 or null
    uint16_t* hi;        // [M, N] split planes or null
    uint16_t* lo;
};

// 32 accumulator columns of one row: v = main + 2^-11 corr (+bias)(gelu)(+residual) -> outputs
__device__ __forceinline__ void gemm_epilogue_row32(const uint32_t* v0, const uint32_t* v1, const float* __restrict__ bias_c,
                                                    const float* rrow, int act_gelu, const GemmOut& out, size_t off)
{
#pragma unroll
    for (int j = 0; j < 32; j += 8) {   // 32-byte (full-sector) vector accesses: a thread owns a row segment
        float o[8];
#pragma unroll
        for (int t = 0; t < 8; t += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias_c + j + t));
            o[t + 0] = fmaf(__uint_as_float(v1[j + t + 0]), GH_LO_INV, __uint_as_float(v0[j + t + 0])) + b4.x;
            o[t + 1] = fmaf(__uint_as_float(v1[j + t + 1]), GH_LO_INV, __uint_as_float(v0[j + t + 1])) + b4.y;
            o[t + 2] = fmaf(__uint_as_float(v1[j + t + 2]), GH_LO_INV, __uint_as_float(v0[j + t + 2])) + b4.z;
            o[t + 3] = fmaf(__uint_as_float(v1[j + t + 3]), GH_LO_INV, __uint_as_float(v0[j + t + 3])) + b4.w;
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
        if (out.c) st_global_v8(out.c + off + j, o);
        if (out.hi) split_store8(o, out.hi + off + j, out.lo + off + j);
    }
}

__global__ void __launch_bounds__(GM_THREADS, 1)
gemm_f16s_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2, int M, int N, int K,
                 const float* __restrict__ bias, const float* __restrict__ residual, int act_gelu, GemmOut out)
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
    const int kblocks = K / GH_KB;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB1); tma_prefetch_desc(&tmB2);
        for (int s = 0; s < GM_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 4); }
        fence_barrier_init();
    }
    if (warp == 1) {   // 2 buffers x (main 128 + correction 128) columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
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
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (lane == 0) {
                    unsigned char* s0 = smem + (size_t)stage * GM_STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], GM_STAGE_BYTES);
                    const int k0 = kb * GH_KB;
                    tma_load_2d(s0, &tmA1, &full_bar[stage], k0, m0, TMA_EVICT_FIRST);
                    tma_load_2d(s0 + GM_SLAB, &tmA2, &full_bar[stage], k0, m0, TMA_EVICT_FIRST);
                    tma_load_2d(s0 + 2 * GM_SLAB, &tmB1, &full_bar[stage], k0, n0, TMA_EVICT_LAST);
                    tma_load_2d(s0 + 3 * GM_SLAB, &tmB2, &full_bar[stage], k0, n0, TMA_EVICT_LAST);
                }
                __syncwarp();
                if (++stage == GM_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_f16(GM_TILE, GM_TILE);
        const uint64_t desc0 = umma_desc_sw128(smem_u32(smem));
        const uint64_t dhi = desc0 & 0xFFFFFFFF00000000ull;
        const uint32_t lo0 = (uint32_t)desc0;
        int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_main = tmem_base + (uint32_t)(acc * 2 * GM_TILE), d_corr = d_main + GM_TILE;
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a1 = lo0 + (uint32_t)((stage * GM_STAGE_BYTES) >> 4);
                    const uint32_t a2 = a1 + (GM_SLAB >> 4), b1 = a1 + (2 * GM_SLAB >> 4), b2 = a1 + (3 * GM_SLAB >> 4);
#pragma unroll
                    for (int k = 0; k < GH_KB / 16; ++k) {
                        const uint32_t first = (uint32_t)((kb | k) != 0);
                        umma_f16(d_main, dhi | (uint64_t)(a1 + 2 * k), dhi | (uint64_t)(b1 + 2 * k), idesc, first);
                        umma_f16(d_corr, dhi | (uint64_t)(a1 + 2 * k), dhi | (uint64_t)(b2 + 2 * k), idesc, first);
                        umma_f16(d_corr, dhi | (uint64_t)(a2 + 2 * k), dhi | (uint64_t)(b1 + 2 * k), idesc, 1u);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (kb == kblocks - 1) umma_commit(&tfull_bar[acc]);
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
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * 2 * GM_TILE);
#pragma unroll 1
            for (int c0 = 0; c0 < GM_TILE; c0 += 32) {
                uint32_t v0[32], v1[32];
                tmem_ld_x32(taddr + c0, v0);
                tmem_ld_x32(taddr + GM_TILE + c0, v1);
                tmem_wait_ld();
                if (row < M) {
                    const size_t off = (size_t)row * N + n0 + c0;
                    gemm_epilogue_row32(v0, v1, bias + n0 + c0, residual ? residual + off : nullptr, act_gelu, out, off);
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
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------- 2-CTA GEMM (cta_group::2, M = 256 per pair)
// Same structure as dense_tc2_kernel (dense_tc.cu): a CTA pair owns a 256 x BN output tile; each CTA stages its
// 128 activation rows (hi and lo planes) and HALF of the BN weight rows (hi and lo), the pair's tensor cores share the
// operands.  Both accumulators of the tile fill the pair's TMEM (2 x BN columns: BN = 256 -> all 512), so a tile's
// epilogue is not overlapped with the next tile's MMAs; with 12..48 k-blocks per tile that costs a few percent.
template <int BN> struct Gm2Cfg {
    static constexpr int BH_BYTES = (BN / 2) * 128;                    // this CTA's half of one weight k-block (one plane)
    static constexpr int STAGE_BYTES = 2 * GM_SLAB + 2 * BH_BYTES;     // 64 KB at BN = 256, 48 KB at BN = 128
    static constexpr int STAGES = (200 * 1024) / STAGE_BYTES;          // 3 / 4
    static constexpr int NBUF = BN <= 128 ? 2 : 1;                     // accumulator sets in TMEM: (main + correction) x NBUF = 512 columns
    static constexpr int TMEM_COLS = 2 * BN * NBUF;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024 + 256;
};
constexpr int GM2_THREADS = 320;   // warp 0: TMA, warp 1: MMA issuer (leader), warps 2-9: epilogue (2 per TMEM lane group)

template <int BN>
__global__ void __launch_bounds__(GM2_THREADS, 1)
gemm2_f16s_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                  const __grid_constant__ CUtensorMap tmB1h, const __grid_constant__ CUtensorMap tmB2h, int M, int N, int K,
                  const float* __restrict__ bias, const float* __restrict__ residual, int act_gelu, GemmOut out)
{
    using Cfg = Gm2Cfg<BN>;
    extern __shared__ unsigned char gm_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gm_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* tail = smem + (size_t)Cfg::STAGES * Cfg::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);             // leader only: 2 arrivals + tx bytes of both CTAs
    uint64_t* empty_bar = full_bar + Cfg::STAGES;                        // per CTA
    uint64_t* tfull_bar = empty_bar + Cfg::STAGES;                       // per CTA [1]
    uint64_t* tempty_bar = tfull_bar + 2;                                // leader only [1]: 16 arrivals
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int m_ptiles = (M + 2 * GM_TILE - 1) / (2 * GM_TILE), n_tiles = N / BN;
    const int total = m_ptiles * n_tiles;
    const int kblocks = K / GH_KB;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB1h); tma_prefetch_desc(&tmB2h);
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
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (lane == 0) {
                    unsigned char* s0 = smem + (size_t)stage * Cfg::STAGE_BYTES;
                    if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                    else mbar_arrive_remote(&full_bar[stage], 0);
                    const int k0 = kb * GH_KB;
                    tma_load_2d_2sm(s0, &tmA1, &full_bar[stage], k0, m0, TMA_EVICT_FIRST);
                    tma_load_2d_2sm(s0 + GM_SLAB, &tmA2, &full_bar[stage], k0, m0, TMA_EVICT_FIRST);
                    tma_load_2d_2sm(s0 + 2 * GM_SLAB, &tmB1h, &full_bar[stage], k0, n0, TMA_EVICT_LAST);
                    tma_load_2d_2sm(s0 + 2 * GM_SLAB + Cfg::BH_BYTES, &tmB2h, &full_bar[stage], k0, n0, TMA_EVICT_LAST);
                }
                __syncwarp();
                if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(2 * GM_TILE, BN);
            const uint64_t desc0 = umma_desc_sw128(smem_u32(smem));
            const uint64_t dhi = desc0 & 0xFFFFFFFF00000000ull;
            const uint32_t lo0 = (uint32_t)desc0;
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
            for (int t = pair; t < total; t += n_pairs) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_main = tmem_base + (uint32_t)(acc * 2 * BN), d_corr = d_main + (uint32_t)BN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t a1 = lo0 + (uint32_t)((stage * Cfg::STAGE_BYTES) >> 4);
                        const uint32_t a2 = a1 + (GM_SLAB >> 4), b1 = a1 + (2 * GM_SLAB >> 4), b2 = b1 + (Cfg::BH_BYTES >> 4);
#pragma unroll
                        for (int k = 0; k < GH_KB / 16; ++k) {
                            const uint32_t first = (uint32_t)((kb | k) != 0);
                            umma_f16_2sm(d_main, dhi | (uint64_t)(a1 + 2 * k), dhi | (uint64_t)(b1 + 2 * k), idesc, first);
                            umma_f16_2sm(d_corr, dhi | (uint64_t)(a1 + 2 * k), dhi | (uint64_t)(b2 + 2 * k), idesc, first);
                            umma_f16_2sm(d_corr, dhi | (uint64_t)(a2 + 2 * k), dhi | (uint64_t)(b1 + 2 * k), idesc, 1u);
                        }
                        umma_commit_2sm(&empty_bar[stage]);
                        if (kb == kblocks - 1) umma_commit_2sm(&tfull_bar[acc]);
                    }
                    __syncwarp();
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == Cfg::NBUF) { acc = 0; acc_phase ^= 1; }
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
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * 2 * BN);
#pragma unroll 1
            for (int c0 = col_half * (BN / 2); c0 < (col_half + 1) * (BN / 2); c0 += 32) {
                uint32_t v0[32], v1[32];
                tmem_ld_x32(taddr + c0, v0);
                tmem_ld_x32(taddr + BN + c0, v1);
                tmem_wait_ld();
                if (row < M) {
                    const size_t off = (size_t)row * N + n0 + c0;
                    gemm_epilogue_row32(v0, v1, bias + n0 + c0, residual ? residual + off : nullptr, act_gelu, out, off);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) mbar_arrive(&tempty_bar[acc]);
                else mbar_arrive_remote(&tempty_bar[acc], 0);
            }
            if (++acc == Cfg::NBUF) { acc = 0; acc_phase ^= 1; }
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
// cuts the output into 128 x BN tiles (BN = 32, 64 or 128) AND the K range into `splits` slices, one CTA each, so that
// ~148 CTAs pull the weight matrix concurrently; when M < 128 only round8(M) activation rows are staged (the other
// MMA rows compute on stale shared memory and are never stored).  splits == 1: bias / GELU / residual epilogue
// straight to the outputs; splits > 1: fp32 partials (main + 2^-11 correction) to ws[split][M][N], summed in split order
// by splitk_reduce_kernel (deterministic), which also applies bias / residual / LayerNorm and writes the split planes.
// Stages are packed: round8(M) activation rows (a_bytes, hi then lo plane) + BN weight rows (hi, lo) per k-block, so that
// for a handful of tokens the whole K slice of a CTA is in flight at once and two CTAs fit on an SM.  The MMA still
// reads a 128-row A operand from each plane's base; the bytes past a_bytes belong to the following planes / stages (or
// the 16 KB slack after the last one) and only feed accumulator rows that are never stored.
constexpr int GSK_MAX_STAGES = 32;
template <int BN> struct GskCfg {
    static constexpr int B_BYTES = BN * 128;      // one plane of one weight k-block
};
static inline size_t gsk_smem_bytes(int n_stages, int stage_bytes) { return 2048 + (size_t)n_stages * stage_bytes + GM_SLAB; }

template <int BN>
__global__ void __launch_bounds__(GM_THREADS, 1)
gemm_sk_f16s_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                    const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2, int M, int N, int kb_per_split,
                    int a_bytes, int n_stages, int b_evict_first, const float* __restrict__ bias, const float* __restrict__ residual,
                    int act_gelu, GemmOut out, float* __restrict__ ws)
{
    using Cfg = GskCfg<BN>;
    extern __shared__ unsigned char gm_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gm_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(base);
    uint64_t* empty_bar = full_bar + GSK_MAX_STAGES;
    uint64_t* tfull_bar = empty_bar + GSK_MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
    unsigned char* smem = base + 1024;
    const int stage_bytes = 2 * a_bytes + 2 * Cfg::B_BYTES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * GM_TILE, split = blockIdx.z;
    const int kb0 = split * kb_per_split;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB1); tma_prefetch_desc(&tmB2);
        for (int s = 0; s < n_stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tfull_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(2 * BN)) : "memory");
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
                const int k0 = (kb0 + i) * GH_KB;
                const uint64_t pol = b_evict_first ? TMA_EVICT_FIRST : TMA_EVICT_LAST;
                tma_load_2d(sa, &tmA1, &full_bar[stage], k0, m0, TMA_EVICT_LAST);
                tma_load_2d(sa + a_bytes, &tmA2, &full_bar[stage], k0, m0, TMA_EVICT_LAST);
                tma_load_2d(sa + 2 * a_bytes, &tmB1, &full_bar[stage], k0, n0, pol);
                tma_load_2d(sa + 2 * a_bytes + Cfg::B_BYTES, &tmB2, &full_bar[stage], k0, n0, pol);
                if (++stage == n_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_f16(GM_TILE, BN);
        const uint64_t desc0 = umma_desc_sw128(smem_u32(smem));
        const uint64_t dhi = desc0 & 0xFFFFFFFF00000000ull;
        const uint32_t lo0 = (uint32_t)desc0;
        const uint32_t d_main = tmem_base, d_corr = tmem_base + (uint32_t)BN;
        int stage = 0; uint32_t phase = 0;
        for (int i = 0; i < kb_per_split; ++i) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t a1 = lo0 + (uint32_t)((stage * stage_bytes) >> 4);
                const uint32_t a2 = a1 + (uint32_t)(a_bytes >> 4), b1 = a2 + (uint32_t)(a_bytes >> 4), b2 = b1 + (Cfg::B_BYTES >> 4);
#pragma unroll
                for (int k = 0; k < GH_KB / 16; ++k) {
                    const uint32_t first = (uint32_t)((i | k) != 0);
                    umma_f16(d_main, dhi | (uint64_t)(a1 + 2 * k), dhi | (uint64_t)(b1 + 2 * k), idesc, first);
                    umma_f16(d_corr, dhi | (uint64_t)(a1 + 2 * k), dhi | (uint64_t)(b2 + 2 * k), idesc, first);
                    umma_f16(d_corr, dhi | (uint64_t)(a2 + 2 * k), dhi | (uint64_t)(b1 + 2 * k), idesc, 1u);
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
            uint32_t v0[32], v1[32];
            tmem_ld_x32(taddr + c0, v0);
            tmem_ld_x32(taddr + BN + c0, v1);
            tmem_wait_ld();
            if (row < M) {
                if (direct) {
                    const size_t off = (size_t)row * N + n0 + c0;
                    gemm_epilogue_row32(v0, v1, bias + n0 + c0, residual ? residual + off : nullptr, act_gelu, out, off);
                } else {
                    float* wrow = ws + ((size_t)split * M + row) * N + n0 + c0;
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        float o[8];
#pragma unroll
                        for (int t = 0; t < 8; ++t) o[t] = fmaf(__uint_as_float(v1[j + t]), GH_LO_INV, __uint_as_float(v0[j + t]));
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
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN)) : "memory");
    }
}

// --------------------------------------------------------------------------- fp32 row kernels
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per row: y = LayerNorm(x) * g + b   (two-pass variance like torch); a lane owns groups of 8 consecutive
// elements (d % 8 == 0, d <= 1024) that it reads ONCE into registers; the fp32 row and the split fp16 planes are written
// with 32- / 16-byte stores
constexpr int LN_MAX_GROUPS = 4;      // 8-element groups per lane: d <= 32 * 4 * 8 = 1024
__device__ __forceinline__ void warp_layernorm_row(const float* __restrict__ x, float* __restrict__ y, uint16_t* __restrict__ y_hi,
                                                   uint16_t* __restrict__ y_lo, const float* __restrict__ g,
                                                   const float* __restrict__ b, int d, float eps, int lane)
{
    const int groups = d >> 3;
    float v[LN_MAX_GROUPS][8];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAX_GROUPS; ++k) {
        const int gi = lane + 32 * k;
        if (gi < groups) {
            const float4 a = *reinterpret_cast<const float4*>(x + gi * 8), c = *reinterpret_cast<const float4*>(x + gi * 8 + 4);
            v[k][0] = a.x; v[k][1] = a.y; v[k][2] = a.z; v[k][3] = a.w; v[k][4] = c.x; v[k][5] = c.y; v[k][6] = c.z; v[k][7] = c.w;
            s += ((a.x + a.y) + (a.z + a.w)) + ((c.x + c.y) + (c.z + c.w));
        }
    }
    const float mean = warp_sum(s) / (float)d;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAX_GROUPS; ++k) {
        if (lane + 32 * k < groups) {
#pragma unroll
            for (int t = 0; t < 8; ++t) { v[k][t] -= mean; }
            q += ((v[k][0] * v[k][0] + v[k][1] * v[k][1]) + (v[k][2] * v[k][2] + v[k][3] * v[k][3])) +
                 ((v[k][4] * v[k][4] + v[k][5] * v[k][5]) + (v[k][6] * v[k][6] + v[k][7] * v[k][7]));
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
#pragma unroll
    for (int k = 0; k < LN_MAX_GROUPS; ++k) {
        const int gi = lane + 32 * k;
        if (gi < groups) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + gi * 8)), g1 = __ldg(reinterpret_cast<const float4*>(g + gi * 8 + 4));
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + gi * 8)), b1 = __ldg(reinterpret_cast<const float4*>(b + gi * 8 + 4));
            float o[8];
            o[0] = v[k][0] * rstd * g0.x + b0.x; o[1] = v[k][1] * rstd * g0.y + b0.y; o[2] = v[k][2] * rstd * g0.z + b0.z; o[3] = v[k][3] * rstd * g0.w + b0.w;
            o[4] = v[k][4] * rstd * g1.x + b1.x; o[5] = v[k][5] * rstd * g1.y + b1.y; o[6] = v[k][6] * rstd * g1.z + b1.z; o[7] = v[k][7] * rstd * g1.w + b1.w;
            if (y) {
                *reinterpret_cast<float4*>(y + gi * 8) = make_float4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<float4*>(y + gi * 8 + 4) = make_float4(o[4], o[5], o[6], o[7]);
            }
            if (y_hi) split_store8(o, y_hi + gi * 8, y_lo + gi * 8);
        }
    }
}

__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, float* __restrict__ y, uint16_t* __restrict__ y_hi, uint16_t* __restrict__ y_lo,
                 const float* __restrict__ g, const float* __restrict__ b, int rows, int d, float eps)
{
    pdl_launch_dependents();
    pdl_wait();
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= rows) return;
    warp_layernorm_row(x + (size_t)r * d, y ? y + (size_t)r * d : nullptr, y_hi ? y_hi + (size_t)r * d : nullptr,
                       y_lo ? y_lo + (size_t)r * d : nullptr, g, b, d, eps, lane);
}

// embeddings: x[t] = LN(word[id] + pos[p] + type[0]); one warp per token; scratch row in shared memory
__global__ void __launch_bounds__(256)
embed_ln_kernel(const int32_t* __restrict__ tok, const int32_t* __restrict__ pos, const float* __restrict__ word,
                const float* __restrict__ pemb, const float* __restrict__ temb, const float* __restrict__ g,
                const float* __restrict__ b, float* __restrict__ x, uint16_t* __restrict__ x_hi, uint16_t* __restrict__ x_lo,
                int n_tok, int d, int vocab, float eps)
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
    warp_layernorm_row(row, x + (size_t)t * d, x_hi + (size_t)t * d, x_lo + (size_t)t * d, g, b, d, eps, lane);
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
                     int act_gelu, const float* __restrict__ ln_g, const float* __restrict__ ln_b, float eps, float* out,
                     uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo)
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
    auto store4 = [&](int c, const float4& o) {
        if (out) *reinterpret_cast<float4*>(out + (size_t)row * N + c) = o;
        if (out_hi) {
            uint16_t h0, l0, h1, l1, h2, l2, h3, l3;
            split_f16(o.x, h0, l0); split_f16(o.y, h1, l1); split_f16(o.z, h2, l2); split_f16(o.w, h3, l3);
            *reinterpret_cast<uint2*>(out_hi + (size_t)row * N + c) = make_uint2((uint32_t)h0 | ((uint32_t)h1 << 16), (uint32_t)h2 | ((uint32_t)h3 << 16));
            *reinterpret_cast<uint2*>(out_lo + (size_t)row * N + c) = make_uint2((uint32_t)l0 | ((uint32_t)l1 << 16), (uint32_t)l2 | ((uint32_t)l3 << 16));
        }
    };
    if (ln_g == nullptr) {
        for (int c = tid * 4; c < N; c += 1024) store4(c, element(c));
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
        store4(c, o);
    }
}

constexpr int AT_ROWS = 16;
constexpr int AT_CHUNK = 64;
__global__ void __launch_bounds__(256)
attention_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ seq_off, uint16_t* __restrict__ out_hi,
                 uint16_t* __restrict__ out_lo, int d, int heads, int max_len)
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
        const size_t ob = (size_t)(t0 + r0 + r) * d + h * dh;
        if (lane < dh) split_f16(o_acc[rr][0], out_hi[ob + lane], out_lo[ob + lane]);
        if (lane + 32 < dh) split_f16(o_acc[rr][1], out_hi[ob + lane + 32], out_lo[ob + lane + 32]);
    }
}

// short sequences (queries; len <= S2 in {32, 64}): one CTA of 128 threads per (sequence, head).  Both products
// are register-tiled out of shared memory: thread (ty, tx) owns S2/8 rows x S2/16 score columns, then S2/8 rows x 4
// output columns; operands are stored transposed where needed so every shared-memory read is a conflict-free
// vector load.  fp32 throughout (scores, max-subtracted softmax, P.V) like torch's math path.
constexpr int ATS_MAX = 64;
template <int S2>
__global__ void __launch_bounds__(128)
attention_tiled_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ seq_off, uint16_t* __restrict__ out_hi,
                       uint16_t* __restrict__ out_lo, int d, int heads)
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
            const size_t ob = (size_t)(t0 + r) * d + h * dh + tx * cw;
#pragma unroll
            for (int j = 0; j < 4; ++j) if (j < cw) split_f16(o_acc[i][j], out_hi[ob + j], out_lo[ob + j]);
        }
    }
}

// long sequences (chunks on the /index path; len up to max_position): flash-style fp32 attention.
// One CTA of 128 threads per (sequence, head, block of 64 query rows); keys are visited in chunks of 64 with an
// online (running max / running sum) softmax, so no S x S score matrix is materialised.  Both products are
// register-tiled exactly like attention_tiled_kernel<64>: thread (ty, tx) owns 8 rows x 4 score columns and
// 8 rows x (dh/16) output columns.
__global__ void __launch_bounds__(128)
attention_flash_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ seq_off, uint16_t* __restrict__ out_hi,
                       uint16_t* __restrict__ out_lo, int d, int heads)
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
            const size_t ob = (size_t)(t0 + r) * d + h * dh + tx * cw;
#pragma unroll
            for (int j = 0; j < 4; ++j) if (j < cw) split_f16(o_acc[i][j] * inv, out_hi[ob + j], out_lo[ob + j]);
        }
    }
}

// ---- tensor-core attention: mma.sync m16n8k16 on the split fp16 planes the QKV projection writes, fp32 accumulators and
// fp32 max-subtracted softmax.  Every product is the error-compensated 3-MMA form of the GEMMs above:
//   A . B^T = hiA.hiB + 2^-11 (hiA.loB + loA.hiB).
// One CTA per (sequence, head, block of 16 x WARPS query rows), one warp per 16 rows (flash-style: online softmax over key
// blocks of KB keys; probabilities go from the score accumulators straight into the A fragments of P.V).  K and V planes
// of a key block are staged in shared memory by cp.async (STAGES = 2: the next block flies during the current one) and read
// with ldmatrix (V transposed on the fly), rows padded by 16 bytes so the 8 x 16-byte row reads hit 8 distinct bank groups.
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_f16(float* c, const uint32_t* a, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t smem_dst, const void* gsrc, bool valid)
{
    const int sz = valid ? 16 : 0;                        // 0 source bytes: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void split_pack2(float a, float b, uint32_t& hi, uint32_t& lo)
{
    uint16_t h0, l0, h1, l1;
    split_f16(a, h0, l0); split_f16(b, h1, l1);
    hi = (uint32_t)h0 | ((uint32_t)h1 << 16); lo = (uint32_t)l0 | ((uint32_t)l1 << 16);
}
__device__ __forceinline__ void am_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void am_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
template <int DH, int WARPS, int KB, int STAGES> struct AmCfg {
    static constexpr int LDH = DH + 8;                                   // row stride in halfs (16 bytes of padding)
    static constexpr int Q_HALFS = 2 * 16 * WARPS * LDH;                 // hi | lo
    static constexpr int KV_HALFS = 2 * KB * LDH;                        // hi | lo of one operand of one stage
    static constexpr size_t SMEM = 2 * (size_t)(Q_HALFS + 2 * STAGES * KV_HALFS);
};
template <int DH, int WARPS, int KB, int STAGES>
__global__ void __launch_bounds__(WARPS * 32)
attention_mma_kernel(const uint16_t* __restrict__ qkv_hi, const uint16_t* __restrict__ qkv_lo, const int32_t* __restrict__ seq_off,
                     uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_lo, int d, int heads)
{
    using Cfg = AmCfg<DH, WARPS, KB, STAGES>;
    constexpr int LDH = Cfg::LDH, QROWS = 16 * WARPS, THREADS = WARPS * 32, CH = DH / 8 /* 16-byte chunks per row */;
    extern __shared__ __align__(16) uint16_t am_sm[];
    uint16_t* sQ = am_sm;                                  // [2][QROWS][LDH]
    uint16_t* sK = sQ + Cfg::Q_HALFS;                      // [STAGES][2][KB][LDH]
    uint16_t* sV = sK + STAGES * Cfg::KV_HALFS;            // [STAGES][2][KB][LDH]
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.z, h = blockIdx.y;
    const int t0 = seq_off[b], len = seq_off[b + 1] - t0;
    const int r0 = blockIdx.x * QROWS;
    if (r0 >= len) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const size_t ld = (size_t)3 * d;
    // rows [row0, row0 + nrows) of one of Q | K | V (column offset `col`) of this head, both planes -> dst[2][nrows][LDH]
    auto stage_rows = [&](uint16_t* dst, int nrows, int row0, int col) {
        for (int i = tid; i < 2 * nrows * CH; i += THREADS) {
            const int pl = i / (nrows * CH), rem = i % (nrows * CH), r = rem / CH, ch = rem % CH;
            const bool valid = row0 + r < len;
            const uint16_t* src = (pl ? qkv_lo : qkv_hi) + (valid ? (size_t)(t0 + row0 + r) * ld + col + h * DH + ch * 8 : 0);
            cp_async16_zfill(smem_u32(dst + ((size_t)pl * nrows + r) * LDH + ch * 8), src, valid);
        }
    };
    auto stage_kv = [&](int kb, int stage) {
        stage_rows(sK + stage * Cfg::KV_HALFS, KB, kb * KB, d);
        stage_rows(sV + stage * Cfg::KV_HALFS, KB, kb * KB, 2 * d);
        am_commit();
    };
    stage_rows(sQ, QROWS, r0, 0);
    stage_kv(0, 0);
    const int nkb = (len + KB - 1) / KB;
    uint32_t qh[DH / 16][4], ql[DH / 16][4];
    float om[DH / 8][4], oc[DH / 8][4];                    // output accumulators: main and correction products
#pragma unroll
    for (int j = 0; j < DH / 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) { om[j][e] = 0.f; oc[j][e] = 0.f; }
    float m_run[2] = {-CUDART_INF_F, -CUDART_INF_F}, l_run[2] = {0.f, 0.f};     // rows g and g + 8 of the warp's 16
    const float scale = rsqrtf((float)DH);
    for (int kb = 0; kb < nkb; ++kb) {
        const int stage = STAGES == 2 ? (kb & 1) : 0;
        if (STAGES == 2 && kb + 1 < nkb) { stage_kv(kb + 1, (kb + 1) & 1); am_wait<1>(); } else am_wait<0>();
        __syncthreads();
        if (kb == 0) {
#pragma unroll
            for (int kk = 0; kk < DH / 16; ++kk) {
                const uint32_t a = smem_u32(sQ + (size_t)(warp * 16 + (lane & 15)) * LDH + kk * 16 + (lane >> 4) * 8);
                ldsm_x4(a, qh[kk][0], qh[kk][1], qh[kk][2], qh[kk][3]);
                ldsm_x4(a + 2 * QROWS * LDH, ql[kk][0], ql[kk][1], ql[kk][2], ql[kk][3]);
            }
        }
        const uint16_t* kH = sK + stage * Cfg::KV_HALFS;
        const uint16_t* vH = sV + stage * Cfg::KV_HALFS;
        // ---- scores: the two correction products first, scaled by 2^-11, then the main product on top (one accumulator)
        float s[KB / 8][4];
#pragma unroll
        for (int j = 0; j < KB / 8; ++j) {
            s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
            for (int k2 = 0; k2 < DH / 32; ++k2) {
                const uint32_t a = smem_u32(kH + (size_t)(j * 8 + (lane & 7)) * LDH + k2 * 32 + (lane >> 3) * 8);
                uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
                ldsm_x4(a, h0, h1, h2, h3);
                ldsm_x4(a + 2 * KB * LDH, l0, l1, l2, l3);
                mma_f16(s[j], qh[2 * k2], l0, l1); mma_f16(s[j], ql[2 * k2], h0, h1);
                mma_f16(s[j], qh[2 * k2 + 1], l2, l3); mma_f16(s[j], ql[2 * k2 + 1], h2, h3);
            }
            s[j][0] *= GH_LO_INV; s[j][1] *= GH_LO_INV; s[j][2] *= GH_LO_INV; s[j][3] *= GH_LO_INV;
#pragma unroll
            for (int k2 = 0; k2 < DH / 32; ++k2) {
                const uint32_t a = smem_u32(kH + (size_t)(j * 8 + (lane & 7)) * LDH + k2 * 32 + (lane >> 3) * 8);
                uint32_t h0, h1, h2, h3;
                ldsm_x4(a, h0, h1, h2, h3);
                mma_f16(s[j], qh[2 * k2], h0, h1); mma_f16(s[j], qh[2 * k2 + 1], h2, h3);
            }
        }
        // ---- online softmax (fp32): thread (g, t) holds columns 2t, 2t+1 of every 8-key tile for rows g (e = 0, 1) and g + 8 (e = 2, 3)
        const int k0 = kb * KB;
        float mx[2] = {-CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
        for (int j = 0; j < KB / 8; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool in = k0 + j * 8 + 2 * t + (e & 1) < len;
                s[j][e] = in ? s[j][e] * scale : -CUDART_INF_F;
                mx[e >> 1] = fmaxf(mx[e >> 1], s[j][e]);
            }
        float corr[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);            // finite: every key block holds at least one valid key
            corr[r] = expf(m_run[r] - m_new);                      // exp(-inf) = 0 on the first block
            m_run[r] = m_new;
        }
        float sum[2] = {0.f, 0.f};
#pragma unroll
        for (int j = 0; j < KB / 8; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float p = expf(s[j][e] - m_run[e >> 1]);     // masked columns: exp(-inf) = 0
                s[j][e] = p;
                sum[e >> 1] += p;
            }
        l_run[0] = l_run[0] * corr[0] + sum[0];                    // per-thread partial row sums (reduced over t at the end)
        l_run[1] = l_run[1] * corr[1] + sum[1];
#pragma unroll
        for (int j = 0; j < DH / 8; ++j) {
            om[j][0] *= corr[0]; om[j][1] *= corr[0]; om[j][2] *= corr[1]; om[j][3] *= corr[1];
            oc[j][0] *= corr[0]; oc[j][1] *= corr[0]; oc[j][2] *= corr[1]; oc[j][3] *= corr[1];
        }
        // ---- out += P . V : the score accumulators of key tiles 2u, 2u+1 ARE the A fragment of 16-key step u
#pragma unroll
        for (int u = 0; u < KB / 16; ++u) {
            uint32_t ph[4], pl[4];
            split_pack2(s[2 * u][0], s[2 * u][1], ph[0], pl[0]);
            split_pack2(s[2 * u][2], s[2 * u][3], ph[1], pl[1]);
            split_pack2(s[2 * u + 1][0], s[2 * u + 1][1], ph[2], pl[2]);
            split_pack2(s[2 * u + 1][2], s[2 * u + 1][3], ph[3], pl[3]);
#pragma unroll
            for (int jp = 0; jp < DH / 16; ++jp) {
                const uint32_t a = smem_u32(vH + (size_t)(u * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDH + jp * 16 + (lane >> 4) * 8);
                uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
                ldsm_x4_trans(a, h0, h1, h2, h3);
                ldsm_x4_trans(a + 2 * KB * LDH, l0, l1, l2, l3);
                mma_f16(om[2 * jp], ph, h0, h1); mma_f16(oc[2 * jp], ph, l0, l1); mma_f16(oc[2 * jp], pl, h0, h1);
                mma_f16(om[2 * jp + 1], ph, h2, h3); mma_f16(oc[2 * jp + 1], ph, l2, l3); mma_f16(oc[2 * jp + 1], pl, h2, h3);
            }
        }
        __syncthreads();                                           // the stage is free again
        if (STAGES == 1 && kb + 1 < nkb) stage_kv(kb + 1, 0);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv[2] = {1.f / l_run[0], 1.f / l_run[1]};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = r0 + warp * 16 + g + 8 * r;
        if (row >= len) continue;
        const size_t ob = (size_t)(t0 + row) * d + h * DH + 2 * t;
#pragma unroll
        for (int j = 0; j < DH / 8; ++j) {
            uint32_t hi, lo;
            split_pack2(fmaf(oc[j][2 * r], GH_LO_INV, om[j][2 * r]) * inv[r], fmaf(oc[j][2 * r + 1], GH_LO_INV, om[j][2 * r + 1]) * inv[r], hi, lo);
            *reinterpret_cast<uint32_t*>(out_hi + ob + j * 8) = hi;
            *reinterpret_cast<uint32_t*>(out_lo + ob + j * 8) = lo;
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

template <int DH, int WARPS, int KB, int STAGES>
static void launch_attention_mma(const uint16_t* q_hi, const uint16_t* q_lo, const int32_t* seq_off, uint16_t* out_hi, uint16_t* out_lo, int d,
                                 int heads, int batch, int max_len, cudaStream_t st)
{
    using Cfg = AmCfg<DH, WARPS, KB, STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
        KRAG_CUDA(cudaFuncSetAttribute(attention_mma_kernel<DH, WARPS, KB, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        attr_set = true;
    }
    launch_pdl(attention_mma_kernel<DH, WARPS, KB, STAGES>, dim3((unsigned)((max_len + 16 * WARPS - 1) / (16 * WARPS)), (unsigned)heads, (unsigned)batch),
               dim3(WARPS * 32), Cfg::SMEM, st, q_hi, q_lo, seq_off, out_hi, out_lo, d, heads);
}
// queries (<= 32 tokens): 2 warps, one key block; <= 64: 4 warps, one block; chunks: 4 warps, 64-key blocks, double-buffered
template <int DH>
static void attention_mma(const uint16_t* q_hi, const uint16_t* q_lo, const int32_t* seq_off, uint16_t* out_hi, uint16_t* out_lo, int d, int heads,
                          int batch, int max_len, cudaStream_t st)
{
    if (max_len <= 32) launch_attention_mma<DH, 2, 32, 1>(q_hi, q_lo, seq_off, out_hi, out_lo, d, heads, batch, max_len, st);
    else if (max_len <= 64) launch_attention_mma<DH, 4, 64, 1>(q_hi, q_lo, seq_off, out_hi, out_lo, d, heads, batch, max_len, st);
    else launch_attention_mma<DH, 4, 64, 2>(q_hi, q_lo, seq_off, out_hi, out_lo, d, heads, batch, max_len, st);
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
// 2D map over one split plane [rows, cols] of fp16, box = 64 halfs (one 128-byte swizzle row) x box_rows
static void emb_map(CUtensorMap* tm, const uint16_t* base, int rows, int cols, int box_rows = GM_TILE)
{
    EncodeTiledFn enc = emb_encode();
    if (!enc) throw CudaError{cudaErrorUnknown, "cuTensorMapEncodeTiled unavailable", __FILE__, __LINE__};
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)GH_KB, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    if (enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<uint16_t*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        throw CudaError{cudaErrorUnknown, "cuTensorMapEncodeTiled failed", __FILE__, __LINE__};
}

// large-M GEMM on split operands: CTA pairs (256 x BN tiles) or single CTAs (128 x 128 tiles)
void launch_gemm_f16s(const DeviceInfo& di, const SplitMat& A, const SplitMat& B, int M, int N, int K, const float* bias,
                      const float* residual, bool gelu, float* C, uint16_t* C_hi, uint16_t* C_lo, cudaStream_t st)
{
    static bool attr_set = false;
    if (!attr_set) {
        KRAG_CUDA(cudaFuncSetAttribute(gemm_f16s_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GM_SMEM));
        attr_set = true;
    }
    const GemmOut out{C, C_hi, C_lo};
    CUtensorMap tmA1, tmA2, tmB1, tmB2;
    emb_map(&tmA1, A.hi, M, K);
    emb_map(&tmA2, A.lo, M, K);
    static int use2 = -1;
    if (use2 < 0) { const char* ev = getenv("KRAG_GEMM_2CTA"); use2 = (ev == nullptr || ev[0] != '0') ? 1 : 0; }
    // CTA pairs (256 x BN tiles) win when there are enough tiles to fill the 74 pairs; small problems (few query
    // tokens per rank) are latency-bound by the K loop of a single tile, where 128 x 128 tiles on single CTAs
    // halve the per-tile MMA time and quadruple the number of CTAs
    static int want_bn = -1;
    if (want_bn < 0) { const char* ev = getenv("KRAG_GEMM_BN"); want_bn = (ev && atoi(ev) == 256) ? 256 : 128; }
    const int BN = (want_bn == 256 && N % 256 == 0) ? 256 : 128;
    const int pair_tiles = ((M + 2 * GM_TILE - 1) / (2 * GM_TILE)) * (N / BN);
    if (use2 && M > GM_TILE && pair_tiles >= di.sm_count / 4) {
        emb_map(&tmB1, B.hi, N, K, BN / 2);
        emb_map(&tmB2, B.lo, N, K, BN / 2);
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
            if (!a256) { KRAG_CUDA(cudaFuncSetAttribute(gemm2_f16s_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Gm2Cfg<256>::SMEM)); a256 = true; }
            cfg.dynamicSmemBytes = Gm2Cfg<256>::SMEM;
            KRAG_CUDA(cudaLaunchKernelEx(&cfg, gemm2_f16s_kernel<256>, tmA1, tmA2, tmB1, tmB2, M, N, K, bias, residual, gelu_i, out));
        } else {
            static bool a128 = false;
            if (!a128) { KRAG_CUDA(cudaFuncSetAttribute(gemm2_f16s_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Gm2Cfg<128>::SMEM)); a128 = true; }
            cfg.dynamicSmemBytes = Gm2Cfg<128>::SMEM;
            KRAG_CUDA(cudaLaunchKernelEx(&cfg, gemm2_f16s_kernel<128>, tmA1, tmA2, tmB1, tmB2, M, N, K, bias, residual, gelu_i, out));
        }
        count_launch();
        return;
    }
    emb_map(&tmB1, B.hi, N, K);
    emb_map(&tmB2, B.lo, N, K);
    const int total = ((M + GM_TILE - 1) / GM_TILE) * (N / GM_TILE);
    const int grid = total < di.sm_count ? total : di.sm_count;
    gemm_f16s_kernel<<<grid, GM_THREADS, GM_SMEM, st>>>(tmA1, tmA2, tmB1, tmB2, M, N, K, bias, residual, gelu ? 1 : 0, out);
    KRAG_CUDA(cudaGetLastError());
    count_launch();
}


// small-M path: 128 x BN tiles x split-K slices, ~one CTA per SM (see gemm_sk_f16s_kernel)
template <int BN>
static void launch_gemm_sk(const CUtensorMap& tmA1, const CUtensorMap& tmA2, const CUtensorMap& tmB1, const CUtensorMap& tmB2, int M, int N,
                           int m_tiles, int splits, int kb_per_split, int a_rows, const float* bias, const float* residual, bool gelu,
                           const GemmOut& out, float* ws, cudaStream_t st)
{
    static bool attr = false;
    if (!attr) { KRAG_CUDA(cudaFuncSetAttribute(gemm_sk_f16s_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); attr = true; }
    const int a_bytes = a_rows * 128, stage_bytes = 2 * a_bytes + 2 * GskCfg<BN>::B_BYTES;
    // a few activation rows: keep a CTA under half an SM's shared memory (two CTAs per SM, the next kernel's CTAs
    // can become resident while this one drains); full 128-row tiles take the whole SM
    const int budget = (a_rows <= 32 ? 96 : 196) * 1024;
    int n_stages = budget / stage_bytes;
    n_stages = n_stages > GSK_MAX_STAGES ? GSK_MAX_STAGES : n_stages;
    n_stages = n_stages > kb_per_split ? kb_per_split : n_stages;
    n_stages = n_stages < 1 ? 1 : n_stages;
    launch_pdl(gemm_sk_f16s_kernel<BN>, dim3((unsigned)(N / BN), (unsigned)m_tiles, (unsigned)splits), dim3(GM_THREADS),
               gsk_smem_bytes(n_stages, stage_bytes), st, tmA1, tmA2, tmB1, tmB2, M, N, kb_per_split, a_bytes, n_stages, m_tiles == 1 ? 1 : 0, bias,
               residual, gelu ? 1 : 0, out, ws);
}

// Linear layer on split operands: C = A . B^T + bias (+GELU) (+residual) -> fp32 C and / or split planes of C; with ln_g the
// row LayerNorm of that goes to Y (fp32) and its split planes (C is then scratch).
void launch_linear(const DeviceInfo& di, const SplitMat& A, const SplitMat& B, int M, int N, int K, const float* bias, const float* residual,
                   bool gelu, float* C, uint16_t* C_hi, uint16_t* C_lo, const float* ln_g, const float* ln_b, float eps, float* Y,
                   uint16_t* Y_hi, uint16_t* Y_lo, float* ws, size_t ws_floats, cudaStream_t st)
{
    static int use_sk = -1;
    if (use_sk < 0) { const char* ev = getenv("KRAG_GEMM_SPLITK"); use_sk = (ev == nullptr || ev[0] != '0') ? 1 : 0; }
    if (ln_g && N > 32 * LN_MAX_GROUPS * 8) throw std::runtime_error("launch_linear: fused LayerNorm rows hold at most 1024 elements");
    const int m_tiles = (M + GM_TILE - 1) / GM_TILE;
    const int tiles128 = m_tiles * (N / GM_TILE);
    if (use_sk && ws && tiles128 < di.sm_count / 2 && N % 64 == 0 && K % GH_KB == 0 && (!ln_g || N <= 1024)) {
        // long-K layers over several activation tiles (FFN2 for tens of queries): 128 x 128 tiles halve the operand
        // re-reads through L2 (the binding resource); split-K restores the CTA count
        static int use_wide = -1;
        if (use_wide < 0) { const char* ev = getenv("KRAG_SK_WIDE"); use_wide = (ev == nullptr || ev[0] != '0') ? 1 : 0; }
        const int BN = m_tiles == 1 ? 32 : ((use_wide && K >= 2048 && N % 128 == 0) ? 128 : 64);
        const int ctas = m_tiles * (N / BN), kblocks = K / GH_KB;
        int splits = 1;
        if (ctas <= di.sm_count / 3) {              // 72 CTAs with the whole K range beat 144 + a reduce kernel
            const int want = di.sm_count / ctas < SK_MAX_SPLITS ? di.sm_count / ctas : SK_MAX_SPLITS;
            for (int sp = want; sp >= 2; --sp)
                if (kblocks % sp == 0 && kblocks / sp >= 2 && (size_t)sp * M * N <= ws_floats) { splits = sp; break; }
        }
        const int a_rows = m_tiles == 1 ? ((M + 7) / 8) * 8 : GM_TILE;
        CUtensorMap tmA1, tmA2, tmB1, tmB2;
        emb_map(&tmA1, A.hi, M, K, a_rows); emb_map(&tmA2, A.lo, M, K, a_rows);
        emb_map(&tmB1, B.hi, N, K, BN); emb_map(&tmB2, B.lo, N, K, BN);
        // splits == 1 and a LayerNorm to follow: the GEMM writes fp32 C (scratch), the LayerNorm kernel writes Y + planes
        const GemmOut direct = ln_g ? GemmOut{C, nullptr, nullptr} : GemmOut{C, C_hi, C_lo};
        if (BN == 32) launch_gemm_sk<32>(tmA1, tmA2, tmB1, tmB2, M, N, m_tiles, splits, kblocks / splits, a_rows, bias, residual, gelu, direct, ws, st);
        else if (BN == 128) launch_gemm_sk<128>(tmA1, tmA2, tmB1, tmB2, M, N, m_tiles, splits, kblocks / splits, a_rows, bias, residual, gelu, direct, ws, st);
        else launch_gemm_sk<64>(tmA1, tmA2, tmB1, tmB2, M, N, m_tiles, splits, kblocks / splits, a_rows, bias, residual, gelu, direct, ws, st);
        if (splits > 1) {
            launch_pdl(splitk_reduce_kernel, dim3((unsigned)M), dim3(256), 0, st, ws, splits, M, N, bias, residual, gelu ? 1 : 0, ln_g, ln_b, eps,
                       ln_g ? Y : C, ln_g ? Y_hi : C_hi, ln_g ? Y_lo : C_lo);
            return;
        }
    } else {
        if (ln_g) launch_gemm_f16s(di, A, B, M, N, K, bias, residual, gelu, C, nullptr, nullptr, st);
        else launch_gemm_f16s(di, A, B, M, N, K, bias, residual, gelu, C, C_hi, C_lo, st);
    }
    if (ln_g) {
        launch_pdl(layernorm_kernel, dim3((unsigned)((M * 32 + 255) / 256)), dim3(256), 0, st, C, Y, Y_hi, Y_lo, ln_g, ln_b, M, N, eps);
    }
}

// test hook (krag_debug_gemm_tf32 / krag_debug_linear_ln): fp32 device operands are split here, then take the product path
void launch_linear_f32(const DeviceInfo& di, const float* A, const float* B, int M, int N, int K, const float* bias, const float* residual,
                       bool gelu, float* C, const float* ln_g, const float* ln_b, float eps, float* Y, float* ws, size_t ws_floats,
                       cudaStream_t st)
{
    uint16_t* planes = nullptr;
    const size_t na = (size_t)M * K, nb = (size_t)N * K;
    KRAG_CUDA(cudaMalloc(&planes, 2 * (2 * na + 2 * nb)));
    SplitMat As{planes, planes + na}, Bs{planes + 2 * na, planes + 2 * na + nb};
    try {
        launch_split_f16(A, As.hi, As.lo, (int64_t)na, st);
        launch_split_f16(B, Bs.hi, Bs.lo, (int64_t)nb, st);
        launch_linear(di, As, Bs, M, N, K, bias, residual, gelu, C, nullptr, nullptr, ln_g, ln_b, eps, Y, nullptr, nullptr, ws, ws_floats, st);
        KRAG_CUDA(cudaStreamSynchronize(st));
    } catch (...) { cudaFree(planes); throw; }
    cudaFree(planes);
}


struct Embedder {
    DeviceInfo di;
    BertConfig cfg;
    std::map<std::string, float*> t;      // HF tensor name -> device copy
    std::vector<float*> bqkv;             // fused Q|K|V bias per layer (finalize)
    // split fp16 planes of the GEMM weights (finalize): per layer QKV (fused), attention output, FFN1, FFN2
    std::vector<SplitMat> s_qkv, s_o, s_f1, s_f2;
    std::vector<uint16_t*> planes;        // owning pointers of all weight planes
    bool finalized = false;
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr, in_ev = nullptr;
    std::map<std::tuple<int, int, int>, cudaGraphExec_t> graphs;   // captured forward per (n_tok, batch, max_len)
    std::map<std::tuple<int, int, int>, int> seen;
    // workspaces, grown on demand
    int cap_tok = 0, cap_batch = 0;
    int32_t *d_tok = nullptr, *d_pos = nullptr, *d_off = nullptr;
    float *x = nullptr, *x2 = nullptr, *qkv = nullptr, *out = nullptr;
    uint16_t *xs = nullptr, *cs = nullptr, *fs = nullptr;   // split planes [2][T][*] of x (LayerNorm outputs), attention context, FFN1 output
    float* ws = nullptr;                  // split-K partials (small token counts)
    static constexpr size_t WS_FLOATS = (size_t)4 << 20;
};

static float* emb_get(Embedder* e, const std::string& name, int64_t n)
{
    auto it = e->t.find(name);
    if (it == e->t.end() || it->second == nullptr) throw std::runtime_error("embedder: tensor not loaded: " + name);
    (void)n;
    return it->second;
}
static std::string lname(int l, const char* s) { return "encoder.layer." + std::to_string(l) + "." + s; }

Embedder* embedder_create(const DeviceInfo& di, const BertConfig& cfg)
{
    if (cfg.hidden % 128 || cfg.inter % 128 || cfg.hidden % cfg.heads || cfg.hidden / cfg.heads > 64 || (cfg.hidden / cfg.heads) % 32 || cfg.hidden > 1024)
        throw std::runtime_error("embedder: hidden/intermediate must be multiples of 128, hidden <= 1024 and head_dim 32 or 64");
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

static SplitMat emb_split_weight(Embedder* e, const float* w, size_t n)
{
    uint16_t* p = nullptr;
    KRAG_CUDA(cudaMalloc(&p, 2 * 2 * n));
    e->planes.push_back(p);
    SplitMat m{p, p + n};
    launch_split_f16(w, m.hi, m.lo, (int64_t)n, e->st);
    return m;
}

void embedder_finalize(Embedder* e)
{
    const int d = e->cfg.hidden, inter = e->cfg.inter;
    for (float* p : e->bqkv) cudaFree(p);
    for (uint16_t* p : e->planes) cudaFree(p);
    e->bqkv.clear(); e->planes.clear(); e->s_qkv.clear(); e->s_o.clear(); e->s_f1.clear(); e->s_f2.clear();
    for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second);   // captured weight pointers are about to change
    e->graphs.clear(); e->seen.clear();
    emb_get(e, "embeddings.word_embeddings.weight", 0); emb_get(e, "embeddings.position_embeddings.weight", 0);
    emb_get(e, "embeddings.token_type_embeddings.weight", 0); emb_get(e, "embeddings.LayerNorm.weight", 0);
    emb_get(e, "embeddings.LayerNorm.bias", 0);
    float* wtmp = nullptr;
    KRAG_CUDA(cudaMalloc(&wtmp, sizeof(float) * (size_t)3 * d * d));
    for (int l = 0; l < e->cfg.layers; ++l) {
        float* b = nullptr;
        KRAG_CUDA(cudaMalloc(&b, sizeof(float) * (size_t)3 * d));
        const char* names[3] = {"attention.self.query", "attention.self.key", "attention.self.value"};
        for (int j = 0; j < 3; ++j) {
            KRAG_CUDA(cudaMemcpyAsync(wtmp + (size_t)j * d * d, emb_get(e, lname(l, names[j]) + ".weight", 0), sizeof(float) * (size_t)d * d, cudaMemcpyDeviceToDevice, e->st));
            KRAG_CUDA(cudaMemcpyAsync(b + (size_t)j * d, emb_get(e, lname(l, names[j]) + ".bias", 0), sizeof(float) * (size_t)d, cudaMemcpyDeviceToDevice, e->st));
        }
        e->bqkv.push_back(b);
        e->s_qkv.push_back(emb_split_weight(e, wtmp, (size_t)3 * d * d));       // stream order: the split reads wtmp before the next layer overwrites it
        for (const char* nm : {"attention.output.dense.bias", "attention.output.LayerNorm.weight", "attention.output.LayerNorm.bias",
                               "intermediate.dense.bias", "output.dense.bias", "output.LayerNorm.weight", "output.LayerNorm.bias"})
            emb_get(e, lname(l, nm), 0);
        e->s_o.push_back(emb_split_weight(e, emb_get(e, lname(l, "attention.output.dense.weight"), 0), (size_t)d * d));
        e->s_f1.push_back(emb_split_weight(e, emb_get(e, lname(l, "intermediate.dense.weight"), 0), (size_t)inter * d));
        e->s_f2.push_back(emb_split_weight(e, emb_get(e, lname(l, "output.dense.weight"), 0), (size_t)d * inter));
    }
    KRAG_CUDA(cudaStreamSynchronize(e->st));
    cudaFree(wtmp);
    // the fp32 copies of the GEMM weights are not read again: free them (the split planes hold the same bytes per element)
    for (int l = 0; l < e->cfg.layers; ++l)
        for (const char* nm : {"attention.self.query.weight", "attention.self.key.weight", "attention.self.value.weight",
                               "attention.output.dense.weight", "intermediate.dense.weight", "output.dense.weight"}) {
            auto it = e->t.find(lname(l, nm));
            if (it != e->t.end() && it->second) { cudaFree(it->second); it->second = nullptr; }
        }
    if (!e->ws) KRAG_CUDA(cudaMalloc(&e->ws, sizeof(float) * Embedder::WS_FLOATS));
    e->finalized = true;
}

void embedder_destroy(Embedder* e)
{
    for (auto& kv : e->graphs) cudaGraphExecDestroy(kv.second);
    for (auto& kv : e->t) if (kv.second) cudaFree(kv.second);
    for (float* p : e->bqkv) cudaFree(p);
    for (uint16_t* p : e->planes) cudaFree(p);
    for (void* p : {(void*)e->d_tok, (void*)e->d_pos, (void*)e->d_off, (void*)e->x, (void*)e->x2, (void*)e->qkv, (void*)e->xs, (void*)e->cs, (void*)e->fs, (void*)e->out})
        if (p) cudaFree(p);
    if (e->ws) cudaFree(e->ws);
    if (e->done) cudaEventDestroy(e->done);
    if (e->in_ev) cudaEventDestroy(e->in_ev);
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
        for (void* p : {(void*)e->d_tok, (void*)e->d_pos, (void*)e->d_off, (void*)e->x, (void*)e->x2, (void*)e->qkv, (void*)e->xs, (void*)e->cs, (void*)e->fs, (void*)e->out})
            if (p) cudaFree(p);
        const size_t T = (size_t)(n_tok > e->cap_tok ? n_tok : e->cap_tok), Bc = (size_t)(batch > e->cap_batch ? batch : e->cap_batch);
        KRAG_CUDA(cudaMalloc(&e->d_tok, 4 * T)); KRAG_CUDA(cudaMalloc(&e->d_pos, 4 * T)); KRAG_CUDA(cudaMalloc(&e->d_off, 4 * (Bc + 1)));
        KRAG_CUDA(cudaMalloc(&e->x, 4 * T * d)); KRAG_CUDA(cudaMalloc(&e->x2, 4 * T * d)); KRAG_CUDA(cudaMalloc(&e->qkv, 4 * T * 3 * d));
        KRAG_CUDA(cudaMalloc(&e->xs, 2 * 2 * T * d)); KRAG_CUDA(cudaMalloc(&e->cs, 2 * 2 * T * d)); KRAG_CUDA(cudaMalloc(&e->fs, 2 * 2 * T * c.inter));
        KRAG_CUDA(cudaMalloc(&e->out, 4 * Bc * d));
        e->cap_tok = (int)T; e->cap_batch = (int)Bc;
    }
    // split planes: [hi plane | lo plane], each cap_tok rows
    const SplitMat xs{e->xs, e->xs + (size_t)e->cap_tok * d}, cs{e->cs, e->cs + (size_t)e->cap_tok * d},
                   fs{e->fs, e->fs + (size_t)e->cap_tok * c.inter};
    KRAG_CUDA(cudaMemcpyAsync(e->d_tok, tok_ids, 4 * (size_t)n_tok, cudaMemcpyHostToDevice, st));
    KRAG_CUDA(cudaMemcpyAsync(e->d_pos, pos.data(), 4 * (size_t)n_tok, cudaMemcpyHostToDevice, st));
    KRAG_CUDA(cudaMemcpyAsync(e->d_off, tok_offsets, 4 * (size_t)(batch + 1), cudaMemcpyHostToDevice, st));

    // The forward is ~7 kernels per layer; for a handful of query tokens it is bound by launch gaps and kernel
    // prologues, so each (n_tok, batch, max_len) shape is captured into a CUDA graph the second time it is seen.
    auto run_layers = [&]() {
    launch_pdl(embed_ln_kernel, dim3((unsigned)((n_tok + 7) / 8)), dim3(256), (size_t)8 * d * 4, st, e->d_tok, e->d_pos,
               e->t["embeddings.word_embeddings.weight"], e->t["embeddings.position_embeddings.weight"],
               e->t["embeddings.token_type_embeddings.weight"], e->t["embeddings.LayerNorm.weight"], e->t["embeddings.LayerNorm.bias"], e->x,
               xs.hi, xs.lo, n_tok, d, c.vocab, c.eps);
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
    static int use_mma = -1;
    if (use_mma < 0) { const char* ev = getenv("KRAG_ATTN_MMA"); use_mma = (ev == nullptr || ev[0] != '0') ? 1 : 0; }
    static int use_flash = -1;
    if (use_flash < 0) { const char* ev = getenv("KRAG_ATTN_FLASH"); use_flash = (ev == nullptr || ev[0] != '0') ? 1 : 0; }
    for (int l = 0; l < c.layers; ++l) {
        const size_t L = (size_t)l;
        if (use_mma) {
            // QKV projection writes the split planes only (same bytes as the fp32 matrix they replace); attention on the tensor cores
            uint16_t* q_hi = reinterpret_cast<uint16_t*>(e->qkv);
            uint16_t* q_lo = q_hi + (size_t)n_tok * 3 * d;
            launch_linear(e->di, xs, e->s_qkv[L], n_tok, 3 * d, d, e->bqkv[L], nullptr, false, nullptr, q_hi, q_lo, nullptr, nullptr, 0.f,
                          nullptr, nullptr, nullptr, e->ws, Embedder::WS_FLOATS, st);
            if (dh == 64) attention_mma<64>(q_hi, q_lo, e->d_off, cs.hi, cs.lo, d, c.heads, batch, max_len, st);
            else attention_mma<32>(q_hi, q_lo, e->d_off, cs.hi, cs.lo, d, c.heads, batch, max_len, st);
        } else {
        // KRAG_ATTN_MMA=0: fp32 qkv and the fp32 CUDA-core attention kernels
        launch_linear(e->di, xs, e->s_qkv[L], n_tok, 3 * d, d, e->bqkv[L], nullptr, false, e->qkv, nullptr, nullptr, nullptr, nullptr, 0.f,
                      nullptr, nullptr, nullptr, e->ws, Embedder::WS_FLOATS, st);
        if (max_len <= 32) {
            launch_pdl(attention_tiled_kernel<32>, dim3((unsigned)batch, (unsigned)c.heads), dim3(128), ats_smem32, st, e->qkv, e->d_off, cs.hi, cs.lo, d, c.heads);
        } else if (max_len <= ATS_MAX) {
            launch_pdl(attention_tiled_kernel<64>, dim3((unsigned)batch, (unsigned)c.heads), dim3(128), ats_smem64, st, e->qkv, e->d_off, cs.hi, cs.lo, d, c.heads);
        } else {
            if (use_flash)
                attention_flash_kernel<<<dim3((unsigned)((max_len + 63) / 64), (unsigned)c.heads, (unsigned)batch), 128, ats_smem64, st>>>(
                    e->qkv, e->d_off, cs.hi, cs.lo, d, c.heads);
            else
                attention_kernel<<<dim3((unsigned)((max_len + AT_ROWS - 1) / AT_ROWS), (unsigned)c.heads, (unsigned)batch), 256, at_smem, st>>>(
                    e->qkv, e->d_off, cs.hi, cs.lo, d, c.heads, max_len);
            KRAG_CUDA(cudaGetLastError());
            count_launch();
        }
        }
        // O projection (+bias +residual) -> LayerNorm -> x (fp32 residual stream) + its split planes
        launch_linear(e->di, cs, e->s_o[L], n_tok, d, d, e->t[lname(l, "attention.output.dense.bias")], e->x, false, e->x2, nullptr, nullptr,
                      e->t[lname(l, "attention.output.LayerNorm.weight")], e->t[lname(l, "attention.output.LayerNorm.bias")], c.eps, e->x, xs.hi, xs.lo,
                      e->ws, Embedder::WS_FLOATS, st);
        // FFN1 (+bias, GELU): only the split planes are written (FFN2 is the sole consumer)
        launch_linear(e->di, xs, e->s_f1[L], n_tok, c.inter, d, e->t[lname(l, "intermediate.dense.bias")], nullptr, true, nullptr, fs.hi, fs.lo,
                      nullptr, nullptr, 0.f, nullptr, nullptr, nullptr, e->ws, Embedder::WS_FLOATS, st);
        // FFN2 (+bias +residual) -> LayerNorm -> x + split planes
        launch_linear(e->di, fs, e->s_f2[L], n_tok, d, c.inter, e->t[lname(l, "output.dense.bias")], e->x, false, e->x2, nullptr, nullptr,
                      e->t[lname(l, "output.LayerNorm.weight")], e->t[lname(l, "output.LayerNorm.bias")], c.eps, e->x, xs.hi, xs.lo, e->ws,
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
    } else if (use_graph && (e->seen.size() > 4096 ? (e->seen.clear(), false) : e->seen[key]++ >= 1)) {
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
        // the caller's buffer may still be read by work it enqueued earlier on `consumer` (the previous batch's scan): only
        // the final writer waits for that; the forward itself overlaps with it
        if (!e->in_ev) KRAG_CUDA(cudaEventCreateWithFlags(&e->in_ev, cudaEventDisableTiming));
        KRAG_CUDA(cudaEventRecord(e->in_ev, consumer));
        KRAG_CUDA(cudaStreamWaitEvent(st, e->in_ev, 0));
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
