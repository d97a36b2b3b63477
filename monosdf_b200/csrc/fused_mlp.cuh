// The SDF network evaluated as ONE persistent tcgen05 kernel: activations never leave the SM.
//
// Replaces, for the sdf-only queries (reference code/model/network.py:131-137 get_sdf_vals, :307-309 for the grid
// class; callers: the sampler ray_sampler.py:131, marching cubes utils/plots.py:145-151), the per-layer launches of
// mlp.cu's forward sweep.  In: 12 B / point (x), out: 4 B / point (sdf); everything between lives in shared memory
// and tensor memory:
//
//   * a CTA owns 256 rows = two 128-row sub-tiles X and Y that ping-pong: while the tensor core runs layer l of one
//     sub-tile (tcgen05.mma 128 x 256 x 16, fp16 operands, fp32 accumulator in TMEM: 256 columns per sub-tile = all 512),
//     the 8 epilogue warps turn the other sub-tile's accumulator into the next layer's A operand;
//   * the A operand of layer l + 1 is written by the epilogue straight into shared memory in the canonical UMMA
//     SWIZZLE_128B K-major layout (the layout TMA would have produced), 64 KB per sub-tile;
//   * the positional encoding is computed in the prologue of a tile (and once more for the skip concat's columns);
//   * weights stream per 64-column k-block (32 KB) through a TMA ring from one packed fp16 matrix (L2 resident, 1 MB);
//   * the last linear layer (one output row: the sdf) is a dot product in the last hidden layer's epilogue.
//
// Arithmetic of the epilogue (the part that bounds the kernel: 32768 activations per 2048-clock MMA pass):
//   * the network runs in a SCALED domain: activations are stored as c h with c = 100 log2(e), so that the accumulator
//     holds v' = c (W h + b) and softplus(beta = 100) becomes  c h = max(v', 0) + log2(1 + 2^-|v'|):  one MUFU (ex2 with
//     free |.| / negate modifiers), a degree-4 polynomial for log2(1 + t) on the FMA pipe, one FMNMX -- no multiply by
//     100 log2(e), no lg2 (the per-layer kernels' ex2 + lg2 version is bound by the XU pipe at 16 results / clk / SM).
//     The scale is folded into the packed weights (layer 0 and the skip concat's input columns x c, biases x c, the
//     last linear layer / c): fp16 rounding is relative, so nothing is lost;
//   * the bias is not added per element: whoever drains an accumulator chunk writes the NEXT layer's bias into it
//     (tcgen05.st) and every MMA accumulates.
//
// Warp roles (640 threads, setmaxnreg 40 / 104): warp 0 TMA producer, warp 1 MMA issuer, warps 4-19 epilogue (TMEM lane
// quadrant = warp % 4; the four warps of a quadrant take 32-column chunks g and g + 4).
#pragma once
#include "tc_gemm.cuh"

namespace msdf_fused {

using namespace msdf_tc;

constexpr int kMaxL = MSDF_MAX_LAYERS;
constexpr int kWStages = 3;
constexpr uint32_t kWStageBytes = 256u * 128u;      // 256 weight rows x 64 columns of fp16
constexpr uint32_t kKBlockBytes = 128u * 128u;      // 128 rows x 64 columns of one sub-tile's A operand
constexpr uint32_t kActBytes = 4u * kKBlockBytes;   // 256 columns
constexpr int kTileRows = 256;
constexpr int kFusedEpiWarps = 16;
constexpr int kFusedThreads = (4 + kFusedEpiWarps) * 32;
constexpr int kFusedRegsLight = 40, kFusedRegsEpi = 104;     // 128 x 40 + 512 x 104 = 58368 <= 640 x 96 (the launch allocation: the pool setmaxnreg draws from)

struct Plan {
    int L;                    // hidden layers = layers run on the tensor core
    int kb[kMaxL];            // 64-column k-blocks of layer l's input
    int skip_after;           // the encoded input is appended to the output row of this layer, from column 256 - d0 (-1: no skip)
    const float* bias;        // [L][256], zero padded
    const float* w_last;      // [256]: the sdf row of the last linear layer (zero padded)
    const float* b_last;      // its bias (device scalar)
    float clamp_radius, sphere_scale;   // get_sdf_vals' bounding-sphere clamp (network.py:134-136); radius <= 0: none
    // training forward (kTrain): the stored activation of layer l is scaled by oscale[l] (1/sqrt2 for the skip concat,
    // network.py:88-89: the convention of the per-layer sweeps that read the saved matrices back)
    float oscale[kMaxL];
    __half* feat; int64_t ldfeat;       // feature head output rows (the colour net's input, fp16), may be null
};

struct StoreMaps { CUtensorMap m[kMaxL + 1]; };     // kTrain: H[l], the saved input of layer l (TMA stores of the operand tiles)


struct FusedBarriers {
    uint64_t wfull[kWStages], wempty[kWStages], actready[2], accfull[2], storedone[2];
    uint32_t tmem_base, pad;
};
static_assert(sizeof(FusedBarriers) <= 128, "the partial-sum slots start 128 bytes behind the barriers");
constexpr uint32_t kFusedTailBytes = 128u + 3u * 128u * 4u;          // barriers + partial sdf sums

constexpr float kHalfPi = 1.57079632679489662f;
constexpr float kScale = 144.26950408889634f;               // c = 100 log2(e)
// log2(1 + t) ~ t (c1 + c2 t + ...) on [0, 1]: minimax fits of log1p x log2(e).  Degree 4: max |error| 7.1e-5, degree 3:
// 5.3e-4 -- in h = (.) / c that is an absolute error of 7.1e-7 / 5.3e-6 against an fp16 storage resolution of 6e-5 at
// |h| = 0.1 (MSDF_FUSED_POLY selects; one FFMA per element apart)
#ifndef MSDF_FUSED_POLY
#define MSDF_FUSED_POLY 3
#endif
constexpr float kLog2e = 1.4426950408889634f;
#if MSDF_FUSED_POLY == 4
constexpr float kC1 = 0.9974489686439758f * kLog2e, kC2 = -0.47130128814472494f * kLog2e, kC3 = 0.225685683948268f * kLog2e,
                kC4 = -0.05875711578778469f * kLog2e;
#else
constexpr float kC1 = 0.987453191724069f * kLog2e, kC2 = -0.4084068241486968f * kLog2e, kC3 = 0.11463518098588148f * kLog2e;
#endif

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// c softplus100(v' / c) for the scaled pre-activation v'
__device__ __forceinline__ float softplus_scaled(float v) {
    const float t = ex2f(-fabsf(v));
#if MSDF_FUSED_POLY == 4
    float p = fmaf(kC4, t, kC3);
    p = fmaf(p, t, kC2);
#else
    float p = fmaf(kC3, t, kC2);
#endif
    p = fmaf(p, t, kC1);
    return fmaf(p, t, fmaxf(v, 0.f));
}

// softplus100(v) for the UNSCALED pre-activation (training forward: the saved activations keep the per-layer sweeps' units)
__device__ __forceinline__ float softplus_plain(float v) {
    constexpr float k = 1.0f / kScale;
    const float t = ex2f(-fabsf(v) * kScale);
#if MSDF_FUSED_POLY == 4
    float p = fmaf(kC4 * k, t, kC3 * k);
    p = fmaf(p, t, kC2 * k);
#else
    float p = fmaf(kC3 * k, t, kC2 * k);
#endif
    p = fmaf(p, t, kC1 * k);
    return fmaf(p, t, fmaxf(v, 0.f));
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// EXPERIMENT (off): the scaled softplus on packed fp16 pairs -- HMNMX2 + three HFMA2 per pair, 4 issue slots per activation
// instead of 6.5.  Measured: no faster (8.45 vs 8.25 ms for 8.4 M points) and less accurate (one more fp16 rounding per
// layer): ex2.approx.f16x2 is TWO MUFU.EX2.F16 + a PRMT, so the XU pipe -- 32768 results per 2048-clock MMA pass at
// 16 / clk / SM, i.e. exactly the MMA time -- stays the limiter, not the issue slots.
#ifndef MSDF_FUSED_HALF2
#define MSDF_FUSED_HALF2 0
#endif
__device__ __forceinline__ uint32_t softplus_scaled_h2(float lo, float hi) {
    const uint32_t vp = WarpIO::pack2<kF16>(lo, hi);                       // cvt.rn.satfinite.f16x2.f32
    const __half2 v = *reinterpret_cast<const __half2*>(&vp);
    // (h2exp2() goes through fp32; the PTX instruction maps to two MUFU.EX2.F16, one per half, no conversions)
    const __half2 na = __hneg2(__habs2(v));
    uint32_t tp;
    asm("ex2.approx.f16x2 %0, %1;" : "=r"(tp) : "r"(*reinterpret_cast<const uint32_t*>(&na)));
    const __half2 t = *reinterpret_cast<const __half2*>(&tp);
    const __half2 c1 = __float2half2_rn(kC1), c2 = __float2half2_rn(kC2), c3 = __float2half2_rn(kC3);
    __half2 p = __hfma2(c3, t, c2);
    p = __hfma2(p, t, c1);
    const __half2 h = __hfma2(p, t, __hmax2(v, __float2half2_rn(0.f)));
    return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t r[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
          "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
          "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// column idx (compile time after unrolling) of the encoded input row of the point x:
// [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...] (embedder.py:5-36), then the hash features hf[0 .. kD0 - kPeW) (zeros
// when the network has none), zero beyond kD0.  cos(a) = sin(a + pi/2).
template <int kPeW, int kD0>
__device__ __forceinline__ float enc_value(const float x[3], const float* hf, int idx) {
    if (idx < 0 || idx >= kD0) return 0.f;
    if (idx < 3) return x[idx];
    if (idx < kPeW) {
        const int t = idx - 3, k = t / 6, r = t - 6 * k, d = r >= 3 ? r - 3 : r;
        return __sinf(fmaf(x[d], (float)(1 << k), r >= 3 ? kHalfPi : 0.f));
    }
    return hf[idx - kPeW];
}

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) { return WarpIO::pack2<kF16>(lo, hi); }

// 32 consecutive fp16 columns (chunk c of the row) of sub-tile `act` into the SWIZZLE_128B K-major operand layout:
// k-block c / 2, row r: 128 bytes = 8 pieces of 16 bytes, piece p stored at position p ^ (r & 7)
__device__ __forceinline__ void store_chunk(uint32_t act, int r, int c, const uint32_t w[16]) {
    const uint32_t rowbase = act + (uint32_t)(c >> 1) * kKBlockBytes + (uint32_t)r * 128u;
    const uint32_t sw = (uint32_t)(r & 7);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const uint32_t piece = (uint32_t)((c & 1) * 4 + p);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowbase + ((piece ^ sw) << 4)), "r"(w[4 * p]), "r"(w[4 * p + 1]),
                     "r"(w[4 * p + 2]), "r"(w[4 * p + 3]) : "memory");
    }
}

// 32 columns [kCol0, kCol0 + 32) of the encoded input row, packed
template <int kPeW, int kD0, int kCol0>
__device__ __forceinline__ void enc_chunk(const float x[3], const float* hf, uint32_t w[16]) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) w[j >> 1] = pack_h2(enc_value<kPeW, kD0>(x, hf, kCol0 + j), enc_value<kPeW, kD0>(x, hf, kCol0 + j + 1));
}
// the skip concat (network.py:88-89): columns >= 256 - kD0 of chunk kC hold the encoded input
template <int kPeW, int kD0, int kC>
__device__ __forceinline__ void skip_fix(const float x[3], const float* hf, float h[32]) {
    constexpr int kSkipCol = 256 - kD0;
    if constexpr (kC * 32 + 32 > kSkipCol) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (kC * 32 + j >= kSkipCol) h[j] = enc_value<kPeW, kD0>(x, hf, kC * 32 + j - kSkipCol);
    }
}
// the point's hash features (fp32 row of hashf) into registers; zeros without a grid or beyond M
template <int kNf>
__device__ __forceinline__ void load_hf(const float* __restrict__ hashf, int64_t grow, int64_t M, float* hf) {
    if constexpr (kNf > 0) {
        if (hashf != nullptr && grow < M) {
#pragma unroll
            for (int j = 0; j < kNf / 4; ++j) {
                const float4 t4 = __ldg(reinterpret_cast<const float4*>(hashf + grow * kNf) + j);
                hf[4 * j] = t4.x; hf[4 * j + 1] = t4.y; hf[4 * j + 2] = t4.z; hf[4 * j + 3] = t4.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < kNf; ++j) hf[j] = 0.f;
        }
    }
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// kTrain = false: sdf-only queries (scaled domain, nothing stored but the sdf).
// kTrain = true : the forward sweep of a training / rendering step: plain units, every layer's input H[l] leaves the SM
//                 ONCE, as TMA stores of the operand tiles the MMAs read anyway (write-only saved activations), plus one
//                 more tensor-core layer for the feature head (rows 1.. of the last linear layer) whose rows go straight
//                 into the colour net's input matrix; sdf_out receives the raw sdf (the decode kernel clamps).
template <int kPeW, int kD0, bool kTrain>
__global__ void __launch_bounds__(kFusedThreads, 1)
k_fused_sdf(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ Plan P, const __grid_constant__ StoreMaps SM,
            const float* __restrict__ x, const float* __restrict__ hashf, int64_t M, float* __restrict__ sdf_out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sAct = base;                                  // 2 x 64 KB
    const uint32_t sW = sAct + 2u * kActBytes;                   // kWStages x 32 KB
    FusedBarriers* bars = reinterpret_cast<FusedBarriers*>(gen_base + 2u * kActBytes + kWStages * kWStageBytes);
    const uint32_t sPart = sW + kWStages * kWStageBytes + 128u;  // [3 groups][128 rows] partial sdf sums
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t num_tiles = (M + kTileRows - 1) / kTileRows;
    const int L = P.L;
    const int LT = kTrain ? L + 1 : L;                           // tensor-core layers: + the feature head when training

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapW);
        for (int s = 0; s < kWStages; ++s) { mbar_init(smem_u32(&bars->wfull[s]), 1); mbar_init(smem_u32(&bars->wempty[s]), 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&bars->actready[s]), kFusedEpiWarps); mbar_init(smem_u32(&bars->accfull[s]), 1);
            mbar_init(smem_u32(&bars->storedone[s]), 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(&bars->tmem_base), kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kFusedRegsLight));
        if (warp == 0) {
            if (lane == 0) {
                // ---- TMA producer: the k-blocks of every (layer, sub-tile) pass, in the order the MMA warp consumes them
                int s = 0; uint32_t ph = 0;
                for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
                    for (int l = 0; l < LT; ++l)
                        for (int sub = 0; sub < 2; ++sub)
                            for (int kb = 0; kb < P.kb[l]; ++kb) {
                                mbar_wait(smem_u32(&bars->wempty[s]), ph ^ 1u);
                                mbar_expect_tx(smem_u32(&bars->wfull[s]), kWStageBytes);
                                tma_load_2d(sW + (uint32_t)s * kWStageBytes, &mapW, smem_u32(&bars->wfull[s]), kb * BK, l * 256);
                                if (++s == kWStages) { s = 0; ph ^= 1u; }
                            }
            }
            __syncwarp();
        } else if (warp == 1) {
            if (lane == 0) {
                // ---- MMA issuer.  Every MMA accumulates: the epilogue warps preset the accumulator to the layer's bias.
                const uint32_t idesc = instr_desc(BM, 256, 0, 0, kF16, kF16);
                int s = 0; uint32_t ph = 0;
                uint32_t par0 = 0u, par1 = 0u;
                for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
                    for (int l = 0; l < LT; ++l)
#pragma unroll
                        for (int sub = 0; sub < 2; ++sub) {
                            // the epilogue has drained this sub-tile's accumulator and written layer l's operand
                            mbar_wait(smem_u32(&bars->actready[sub]), sub == 0 ? par0 : par1);
                            if (sub == 0) par0 ^= 1u; else par1 ^= 1u;
                            tc_fence_after();
                            const uint32_t tmem_d = tmem_base + (uint32_t)sub * 256u;
                            const uint32_t act = sAct + (uint32_t)sub * kActBytes;
                            const int nkb = P.kb[l];

                            for (int kb = 0; kb < nkb; ++kb) {
                                mbar_wait(smem_u32(&bars->wfull[s]), ph);
                                tc_fence_after();
#pragma unroll
                                for (int k = 0; k < BK / UMMA_K; ++k) {
                                    const uint64_t da = smem_desc(act + (uint32_t)kb * kKBlockBytes + k * (UMMA_K * 2), 16, 1024);
                                    const uint64_t db = smem_desc(sW + (uint32_t)s * kWStageBytes + k * (UMMA_K * 2), 16, 1024);
                                    umma_bf16(tmem_d, da, db, idesc, 1u);
                                }
                                umma_commit(smem_u32(&bars->wempty[s]));
                                if (++s == kWStages) { s = 0; ph ^= 1u; }
                            }
                            umma_commit(smem_u32(&bars->accfull[sub]));
                        }
            }
            __syncwarp();
        } else if (warp == 2) {
            if constexpr (kTrain) {
                if (lane == 0) {
                    // ---- store warp: H[l] <- the operand tile of every (layer, sub-tile) pass, as soon as the epilogue has
                    // written it (its writes are fenced for the async proxy); storedone tells the epilogue warps that the
                    // tile has been read and may be overwritten.  (In the MMA warp these waits stalled the tensor pipe.)
                    uint32_t par0 = 0u, par1 = 0u;
                    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
                        for (int l = 0; l < LT; ++l)
#pragma unroll
                            for (int sub = 0; sub < 2; ++sub) {
                                mbar_wait(smem_u32(&bars->actready[sub]), sub == 0 ? par0 : par1);
                                if (sub == 0) par0 ^= 1u; else par1 ^= 1u;
                                const uint32_t act = sAct + (uint32_t)sub * kActBytes;
                                const int row = (int)(tile * kTileRows + sub * 128);
                                for (int kb = 0; kb < P.kb[l]; ++kb) tma_store_2d(&SM.m[l], act + (uint32_t)kb * kKBlockBytes, kb * BK, row);
                                tma_store_commit();
                                tma_store_wait_read();
                                mbar_arrive(smem_u32(&bars->storedone[sub]));
                            }
                    tma_store_wait_all();
                }
                __syncwarp();
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kFusedRegsEpi));
        // ---- epilogue warps: quadrant q = TMEM lanes 32 q .. 32 q + 31 = rows of the sub-tile; group g takes chunks g, g + 4
        const int q = warp & 3, g = (warp - 4) >> 2;
        const int r = q * 32 + lane;
        const uint32_t tlane = (uint32_t)(q * 32) << 16;
        constexpr int kNf = kD0 - kPeW;                             // hash features per point (0 or 32)
        constexpr int kKb0 = (kD0 + 63) / 64;
        uint32_t par0 = 0u, par1 = 0u;
        // this thread's row of the SWIZZLE_128B operand layout: 16-byte piece p of a 128-byte row sits at p ^ (r & 7);
        // split into the part that depends on the chunk parity (bit 2 of the piece: swhi) and the four low offsets
        const uint32_t rowoff = (uint32_t)r * 128u, swhi = (uint32_t)(r & 4) << 4;
        uint32_t swlo[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) swlo[p] = (uint32_t)(p ^ (r & 3)) << 4;

        // accumulator chunk `c` of a sub-tile <- the bias of layer l (same 32 values in every row)
        auto preset_bias = [&](uint32_t tacc, int l, int c) {
            uint32_t bb[32];
            const float4* src = reinterpret_cast<const float4*>(P.bias + l * 256 + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 t4 = __ldg(src + j);
                bb[4 * j] = __float_as_uint(t4.x); bb[4 * j + 1] = __float_as_uint(t4.y);
                bb[4 * j + 2] = __float_as_uint(t4.z); bb[4 * j + 3] = __float_as_uint(t4.w);
            }
            tmem_st32(tacc + (uint32_t)c * 32u, bb);
        };
        // encoded input of a sub-tile -> k-blocks 0 .. kKb0-1 of its operand buffer; accumulator <- bias of layer 0
        auto produce_input = [&](int sub, const float xv[3], int64_t grow) {
            const uint32_t act = sAct + (uint32_t)sub * kActBytes;
            const uint32_t tacc = tmem_base + tlane + (uint32_t)sub * 256u;
            preset_bias(tacc, 0, g);
            preset_bias(tacc, 0, g + 4);
            if (g < 2 * kKb0) {
                float hf[kNf > 0 ? kNf : 1];
                load_hf<kNf>(hashf, grow, M, hf);
                uint32_t w[16];
                if (g == 0) enc_chunk<kPeW, kD0, 0>(xv, hf, w);
                else if (g == 1) enc_chunk<kPeW, kD0, 32>(xv, hf, w);
                else if (g == 2) enc_chunk<kPeW, kD0, 64>(xv, hf, w);
                else enc_chunk<kPeW, kD0, 96>(xv, hf, w);
                store_chunk(act, r, g, w);
            }
            tmem_st_wait();
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars->actready[sub]));
        };
        auto load_x = [&](int64_t grow, float xv[3]) {
            if (grow < M) { xv[0] = __ldg(x + 3 * grow); xv[1] = __ldg(x + 3 * grow + 1); xv[2] = __ldg(x + 3 * grow + 2); }
            else { xv[0] = xv[1] = xv[2] = 0.f; }
        };

        float xc[2][3];                                            // the points of the current tile's two sub-tiles
        if ((int64_t)blockIdx.x < num_tiles) {
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
                const int64_t grow = (int64_t)blockIdx.x * kTileRows + sub * 128 + r;
                load_x(grow, xc[sub]);
                produce_input(sub, xc[sub], grow);
            }
        }
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int64_t next = tile + gridDim.x;
            float xn[2][3];
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) load_x(next < num_tiles ? next * kTileRows + sub * 128 + r : M, xn[sub]);
#pragma unroll 1
            for (int l = 0; l < LT; ++l) {
                const bool last = l == L - 1;                 // last hidden layer: its output also feeds the sdf head
                const bool head = kTrain && l == L;           // feature head (training only): linear, rows leave for the colour net
                const bool skip = l == P.skip_after;
                const bool writes_act = kTrain ? !head : !last;
                const float os = kTrain ? P.oscale[l < L ? l : 0] : 1.0f;
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    const uint32_t act = sAct + (uint32_t)sub * kActBytes;
                    const uint32_t tacc = tmem_base + tlane + (uint32_t)sub * 256u;
                    const int64_t grow = tile * kTileRows + sub * 128 + r;
                    mbar_wait(smem_u32(&bars->accfull[sub]), sub == 0 ? par0 : par1);
                    if constexpr (kTrain) mbar_wait(smem_u32(&bars->storedone[sub]), sub == 0 ? par0 : par1);   // same sequence of phases
                    if (sub == 0) par0 ^= 1u; else par1 ^= 1u;
                    tc_fence_after();
                    float dot = 0.f;
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int c = g + 4 * i;
                        // one accumulator chunk at a time (the other three warps of the scheduler hide the TMEM latency);
                        // the next layer's bias for this chunk is fetched BEFORE the math so that its store never waits
                        uint32_t acc[32], bb[32];
                        tmem_ld32_issue(tacc + (uint32_t)c * 32u, acc);
                        if (writes_act) {
                            const float4* src = reinterpret_cast<const float4*>(P.bias + (l + 1) * 256 + c * 32);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 t4 = __ldg(src + j);
                                bb[4 * j] = __float_as_uint(t4.x); bb[4 * j + 1] = __float_as_uint(t4.y);
                                bb[4 * j + 2] = __float_as_uint(t4.z); bb[4 * j + 3] = __float_as_uint(t4.w);
                            }
                        }
                        tmem_ld32_wait(acc);
                        if (head) {
                            // feature head: acc + bias (preset) -> fp16 -> the colour net's input row, 64 contiguous bytes per lane
                            if (P.feat != nullptr && grow < M) {
                                uint4* dst = reinterpret_cast<uint4*>(P.feat + grow * P.ldfeat + c * 32);
#pragma unroll
                                for (int p = 0; p < 4; ++p) {
                                    uint4 q;
                                    q.x = pack_h2(__uint_as_float(acc[8 * p]), __uint_as_float(acc[8 * p + 1]));
                                    q.y = pack_h2(__uint_as_float(acc[8 * p + 2]), __uint_as_float(acc[8 * p + 3]));
                                    q.z = pack_h2(__uint_as_float(acc[8 * p + 4]), __uint_as_float(acc[8 * p + 5]));
                                    q.w = pack_h2(__uint_as_float(acc[8 * p + 6]), __uint_as_float(acc[8 * p + 7]));
                                    dst[p] = q;
                                }
                            }
                            continue;
                        }
                        if (last) {                                         // sdf head: dot with the fp32 activations
                            const float4* wl4 = reinterpret_cast<const float4*>(P.w_last + c * 32);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 t4 = __ldg(wl4 + j);
                                if constexpr (kTrain) {
                                    // (keep the activations: they are stored below)
                                    float h0 = softplus_plain(__uint_as_float(acc[4 * j])), h1 = softplus_plain(__uint_as_float(acc[4 * j + 1]));
                                    float h2 = softplus_plain(__uint_as_float(acc[4 * j + 2])), h3 = softplus_plain(__uint_as_float(acc[4 * j + 3]));
                                    dot = fmaf(h0, t4.x, dot); dot = fmaf(h1, t4.y, dot); dot = fmaf(h2, t4.z, dot); dot = fmaf(h3, t4.w, dot);
                                    acc[4 * j] = __float_as_uint(h0); acc[4 * j + 1] = __float_as_uint(h1);
                                    acc[4 * j + 2] = __float_as_uint(h2); acc[4 * j + 3] = __float_as_uint(h3);
                                } else {
                                    dot = fmaf(softplus_scaled(__uint_as_float(acc[4 * j])), t4.x, dot);
                                    dot = fmaf(softplus_scaled(__uint_as_float(acc[4 * j + 1])), t4.y, dot);
                                    dot = fmaf(softplus_scaled(__uint_as_float(acc[4 * j + 2])), t4.z, dot);
                                    dot = fmaf(softplus_scaled(__uint_as_float(acc[4 * j + 3])), t4.w, dot);
                                }
                            }
                            if constexpr (!kTrain) continue;
                        }
                        // hidden layer (and, when training, the last one: its activations are already in acc)
                        if (skip && i == 1) {                            // the skip concat's columns live in chunks 5 .. 7
                            float h[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) h[j] = kTrain ? softplus_plain(__uint_as_float(acc[j])) : softplus_scaled(__uint_as_float(acc[j]));
                            float hf[kNf > 0 ? kNf : 1];
                            load_hf<kNf>(hashf, grow, M, hf);
                            if (g == 1) skip_fix<kPeW, kD0, 5>(xc[sub], hf, h);
                            else if (g == 2) skip_fix<kPeW, kD0, 6>(xc[sub], hf, h);
                            else if (g == 3) skip_fix<kPeW, kD0, 7>(xc[sub], hf, h);
                            uint32_t w[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) w[j] = pack_h2(h[2 * j] * os, h[2 * j + 1] * os);
                            store_chunk(act, r, c, w);
                        } else {
                            // 8 columns at a time: activation, pack, one 16-byte store into the swizzled operand row
                            const uint32_t cbase = act + (uint32_t)(c >> 1) * kKBlockBytes + rowoff + ((uint32_t)((c & 1) << 6) ^ swhi);
#pragma unroll
                            for (int p = 0; p < 4; ++p) {
                                uint32_t w[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    float h0 = __uint_as_float(acc[8 * p + 2 * j]), h1 = __uint_as_float(acc[8 * p + 2 * j + 1]);
                                    if constexpr (kTrain) {
                                        if (!last) { h0 = softplus_plain(h0); h1 = softplus_plain(h1); }
                                        h0 *= os; h1 *= os;
                                    } else {
#if MSDF_FUSED_HALF2 && MSDF_FUSED_POLY == 3
                                        w[j] = softplus_scaled_h2(h0, h1);
                                        continue;
#else
                                        h0 = softplus_scaled(h0); h1 = softplus_scaled(h1);
#endif
                                    }
                                    w[j] = pack_h2(h0, h1);
                                }
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cbase + swlo[p]), "r"(w[0]), "r"(w[1]),
                                             "r"(w[2]), "r"(w[3]) : "memory");
                            }
                        }
                        tmem_st32(tacc + (uint32_t)c * 32u, bb);
                    }
                    if (writes_act) {
                        tmem_st_wait();
                        fence_proxy_async();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&bars->actready[sub]));
                    }
                    if (last) {
                        // sdf = dot over all 256 columns: the four warps of a quadrant combine through three 512-byte slots
                        // behind the barriers (fixed order: deterministic); the second barrier lets the slots be reused
                        if (g != 0) asm volatile("st.shared.f32 [%0], %1;" ::"r"(sPart + (uint32_t)((g - 1) * 128 + r) * 4u), "f"(dot) : "memory");
                        named_bar_sync(1 + q, 128);
                        if (g == 0) {
                            float o1, o2, o3;
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o1) : "r"(sPart + (uint32_t)r * 4u) : "memory");
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o2) : "r"(sPart + (uint32_t)(128 + r) * 4u) : "memory");
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o3) : "r"(sPart + (uint32_t)(256 + r) * 4u) : "memory");
                            float sdf = ((dot + o1) + (o2 + o3)) + __ldg(P.b_last);
                            if (P.clamp_radius > 0.f) {
                                const float* p3 = xc[sub];
                                sdf = fminf(sdf, P.sphere_scale * (P.clamp_radius - sqrtf(p3[0] * p3[0] + p3[1] * p3[1] + p3[2] * p3[2])));
                            }
                            if (grow < M) sdf_out[grow] = sdf;
                        }
                        named_bar_sync(1 + q, 128);
                    }
                    if (l == LT - 1) {
                        // this sub-tile's accumulator and operand buffer are free: the next tile's input goes in
                        xc[sub][0] = xn[sub][0]; xc[sub][1] = xn[sub][1]; xc[sub][2] = xn[sub][2];
                        if (next < num_tiles) produce_input(sub, xc[sub], next * kTileRows + sub * 128 + r);
                        else tc_fence_before();
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

// Packed parameters.  Scaled domain (sdf-only queries, c = kScale):
//   Wp[(l * 256 + r) * 256 + k] = fp16(W_l[r, k] * s_l(k)), zero padded; s = c for layer 0 (unscaled input), 1 otherwise;
//        the skip layer: 1/sqrt2 on the columns of the previous layer's (scaled) output, c/sqrt2 on the input's columns
//   bp[l * 256 + r] = c b_l[r];   wl[k] = W_last[0, k] / c
// Plain domain (training forward): the weights as they are (the stored activations carry the 1/sqrt2 of the skip concat),
//   plus one more layer: rows 1 .. 256 of the last linear layer (the feature head) with their biases.
struct PackArgs {
    int L;
    const float* W[kMaxL]; const float* b[kMaxL];
    int out[kMaxL], in[kMaxL]; int64_t ldw[kMaxL];
    int skip, skip_col;                 // skip layer (-1: none) and its first input column
    const float* w_last; const float* b_last; int in_last; int64_t ldw_last;
    int plain, feat_rows;               // plain domain; rows of the feature head (0: none)
};
__global__ void k_pack_fused(const __grid_constant__ PackArgs a, __half* __restrict__ Wp, float* __restrict__ bp, float* __restrict__ wl) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int LT = a.L + (a.feat_rows > 0 ? 1 : 0);
    const int64_t nW = (int64_t)LT * 65536;
    if (i < nW) {
        const int l = (int)(i >> 16), r = (int)((i >> 8) & 255), k = (int)(i & 255);
        float v = 0.f;
        if (l == a.L) {                 // feature head: rows 1 .. of the last layer
            if (r < a.feat_rows && k < a.in_last) v = a.w_last[(int64_t)(r + 1) * a.ldw_last + k];
        } else {
            float sc = 1.0f;
            if (!a.plain) {
                sc = l == 0 ? kScale : 1.0f;
                if (l == a.skip) sc = (k < a.skip_col ? 1.0f : kScale) * 0.70710678118654752440f;
            }
            if (r < a.out[l] && k < a.in[l]) v = a.W[l][(int64_t)r * a.ldw[l] + k] * sc;
        }
        Wp[i] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    } else if (i < nW + (int64_t)LT * 256) {
        const int64_t t = i - nW;
        const int l = (int)(t >> 8), r = (int)(t & 255);
        float v = 0.f;
        if (l == a.L) { if (r < a.feat_rows) v = a.b_last[r + 1]; }
        else if (r < a.out[l]) v = a.b[l][r] * (a.plain ? 1.0f : kScale);
        bp[t] = v;
    } else if (i < nW + (int64_t)LT * 256 + 256) {
        const int k = (int)(i - nW - (int64_t)LT * 256);
        wl[k] = k < a.in_last ? a.w_last[k] * (a.plain ? 1.0f : 1.0f / kScale) : 0.f;
    }
}

inline size_t fused_workspace_bytes(int L) { return (size_t)(L + 1) * 131072 + (size_t)(L + 1) * 1024 + 1024; }   // (+ the feature head)

// can the network run fused?  hidden widths 256 (the layer before the skip: 256 - d0), at most 128 encoded inputs
inline bool fused_supported(int nl, const int* in, const int* out, int skip, int d0, int pe_w) {
    if (nl < 2 || nl - 1 > kMaxL || pe_w != 39 || (d0 != 39 && d0 != 71)) return false;
    for (int l = 0; l < nl - 1; ++l) {
        const int want = (l + 1 == skip) ? 256 - d0 : 256;
        if (out[l] != want) return false;
        if (l > 0 && in[l] != 256) return false;
    }
    return in[nl - 1] == 256 && skip != 1;     // (a skip into layer 1 would need the input next to itself)
}

}  // namespace msdf_fused
