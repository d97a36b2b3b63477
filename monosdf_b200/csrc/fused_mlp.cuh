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
// Softplus(beta = 100): max(v, 0) + log1p(t) / 100 with t = exp(-100 |v|): ONE MUFU (ex2) per element, log1p as a
// degree-4 polynomial on the FMA pipe (|error| < 7.1e-7 absolute in h, far below fp16 storage resolution) -- the
// per-layer kernels' ex2 + lg2 version is bound by the XU pipe at 16 results / clk / SM.
//
// Warp roles (384 threads, setmaxnreg 40 / 232 like k_tc_gemm): warp 0 TMA producer, warp 1 MMA issuer, warps 4-11
// epilogue (TMEM lane quadrant = warp % 4, the two warps of a quadrant split the 32-column chunks).
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

struct Plan {
    int L;                    // hidden layers = layers run on the tensor core
    int kb[kMaxL];            // 64-column k-blocks of layer l's input
    int skip_after;           // the encoded input is appended to the output row of this layer, from column 256 - d0 (-1: no skip)
    const float* bias;        // [L][256], zero padded
    const float* w_last;      // [256]: the sdf row of the last linear layer (zero padded)
    const float* b_last;      // its bias (device scalar)
    float clamp_radius, sphere_scale;   // get_sdf_vals' bounding-sphere clamp (network.py:134-136); radius <= 0: none
};

struct FusedBarriers {
    uint64_t wfull[kWStages], wempty[kWStages], actready[2], accfull[2];
    uint32_t tmem_base, pad;
};

constexpr float kHalfPi = 1.57079632679489662f;
constexpr float k100Log2e = 144.26950408889634f;
// log1p(t) / 100 ~ t (c1 + c2 t + c3 t^2 + c4 t^3) on [0, 1]  (minimax fit, max |error| 7.1e-5 before the / 100)
constexpr float kC1 = 0.9974489686439758e-2f, kC2 = -0.47130128814472494e-2f, kC3 = 0.225685683948268e-2f, kC4 = -0.05875711578778469e-2f;

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ __forceinline__ float softplus100_poly(float v) {
    const float t = ex2f(-fabsf(v) * k100Log2e);
    float p = fmaf(kC4, t, kC3);
    p = fmaf(p, t, kC2);
    p = fmaf(p, t, kC1);
    return fmaf(p, t, fmaxf(v, 0.f));
}

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

template <int kPeW, int kD0>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_fused_sdf(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ Plan P, const float* __restrict__ x,
            const float* __restrict__ hashf, int64_t M, float* __restrict__ sdf_out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sAct = base;                                  // 2 x 64 KB
    const uint32_t sW = sAct + 2u * kActBytes;                   // kWStages x 32 KB
    const uint32_t sPart = sW + kWStages * kWStageBytes;         // [2][128] floats: half-1 partial dot products
    FusedBarriers* bars = reinterpret_cast<FusedBarriers*>(gen_base + 2u * kActBytes + kWStages * kWStageBytes + 1024u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t num_tiles = (M + kTileRows - 1) / kTileRows;
    const int L = P.L;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapW);
        for (int s = 0; s < kWStages; ++s) { mbar_init(smem_u32(&bars->wfull[s]), 1); mbar_init(smem_u32(&bars->wempty[s]), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&bars->actready[s]), kEpiWarps); mbar_init(smem_u32(&bars->accfull[s]), 1); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(&bars->tmem_base), kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsLight));
        if (warp == 0) {
            if (lane == 0) {
                // ---- TMA producer: the k-blocks of every (layer, sub-tile) pass, in the order the MMA warp consumes them
                int s = 0; uint32_t ph = 0;
                for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
                    for (int l = 0; l < L; ++l)
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
                // ---- MMA issuer
                const uint32_t idesc = instr_desc(BM, 256, 0, 0, kF16, kF16);
                int s = 0; uint32_t ph = 0;
                uint32_t par[2] = {0u, 0u};
                for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
                    for (int l = 0; l < L; ++l)
#pragma unroll
                        for (int sub = 0; sub < 2; ++sub) {
                            // the epilogue has read this sub-tile's accumulator and written layer l's operand
                            mbar_wait(smem_u32(&bars->actready[sub]), par[sub]);
                            par[sub] ^= 1u;
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
                                    umma_bf16(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                                }
                                umma_commit(smem_u32(&bars->wempty[s]));
                                if (++s == kWStages) { s = 0; ph ^= 1u; }
                            }
                            umma_commit(smem_u32(&bars->accfull[sub]));
                        }
            }
            __syncwarp();
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpi));
        // ---- epilogue warps
        const int q = warp & 3, half = (warp - 4) >> 2;
        const int r = q * 32 + lane;                               // row inside a sub-tile = TMEM lane
        const uint32_t tlane = (uint32_t)(q * 32) << 16;
        constexpr int kNf = kD0 - kPeW;                             // hash features per point (0 or 32)
        constexpr int kKb0 = (kD0 + 63) / 64;
        uint32_t par0 = 0u, par1 = 0u;

        // encoded input of a sub-tile -> k-blocks 0 .. kKb0-1 of its operand buffer
        auto produce_input = [&](int sub, const float xv[3], int64_t grow) {
            const uint32_t act = sAct + (uint32_t)sub * kActBytes;
            float hf[kNf > 0 ? kNf : 1];
            load_hf<kNf>(hashf, grow, M, hf);
            uint32_t w[16];
            if (half == 0) enc_chunk<kPeW, kD0, 0>(xv, hf, w); else enc_chunk<kPeW, kD0, 32>(xv, hf, w);
            store_chunk(act, r, half, w);
            if constexpr (kKb0 > 1) {
                if (half == 0) enc_chunk<kPeW, kD0, 64>(xv, hf, w); else enc_chunk<kPeW, kD0, 96>(xv, hf, w);
                store_chunk(act, r, 2 + half, w);
            }
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
            for (int l = 0; l < L; ++l) {
                const bool last = l == L - 1;
                const bool skip = l == P.skip_after;
                const float* __restrict__ bias = P.bias + l * 256 + half * 32;
#pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    const uint32_t act = sAct + (uint32_t)sub * kActBytes;
                    const uint32_t tacc = tmem_base + tlane + (uint32_t)sub * 256u + (uint32_t)half * 32u;
                    mbar_wait(smem_u32(&bars->accfull[sub]), sub == 0 ? par0 : par1);
                    if (sub == 0) par0 ^= 1u; else par1 ^= 1u;
                    tc_fence_after();
                    uint32_t acc[2][32];
                    float dot = 0.f;
                    tmem_ld32_issue(tacc, acc[0]);
                    // chunks c = half + 2 i, two per iteration (TMEM loads double buffered in registers)
#pragma unroll 1
                    for (int ii = 0; ii < 2; ++ii) {
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const int i = 2 * ii + k;
                            const int c = half + 2 * i;
                            float b[32];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 t4 = __ldg(reinterpret_cast<const float4*>(bias + i * 64) + j);
                                b[4 * j] = t4.x; b[4 * j + 1] = t4.y; b[4 * j + 2] = t4.z; b[4 * j + 3] = t4.w;
                            }
                            tmem_ld32_wait(acc[k]);
                            if (i + 1 < 4) tmem_ld32_issue(tacc + (uint32_t)(i + 1) * 64u, acc[k ^ 1]);
                            float h[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) h[j] = softplus100_poly(__uint_as_float(acc[k][j]) + b[j]);
                            if (!last) {
                                if (skip && i == 3) {                        // the skip concat's columns: chunks 6 / 7 (kD0 <= 64)
                                    float hf[kNf > 0 ? kNf : 1];
                                    load_hf<kNf>(hashf, tile * kTileRows + sub * 128 + r, M, hf);
                                    if (half == 0) skip_fix<kPeW, kD0, 6>(xc[sub], hf, h); else skip_fix<kPeW, kD0, 7>(xc[sub], hf, h);
                                }
                                if constexpr (kD0 > 64) {
                                    if (skip && i == 2) {                    // ... and 4 / 5 for wider inputs
                                        float hf[kNf > 0 ? kNf : 1];
                                        load_hf<kNf>(hashf, tile * kTileRows + sub * 128 + r, M, hf);
                                        if (half == 0) skip_fix<kPeW, kD0, 4>(xc[sub], hf, h); else skip_fix<kPeW, kD0, 5>(xc[sub], hf, h);
                                    }
                                }
                                uint32_t w[16];
#pragma unroll
                                for (int j = 0; j < 16; ++j) w[j] = pack_h2(h[2 * j], h[2 * j + 1]);
                                store_chunk(act, r, c, w);
                            } else {
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float4 t4 = __ldg(reinterpret_cast<const float4*>(P.w_last + c * 32) + j);
                                    dot = fmaf(h[4 * j], t4.x, dot); dot = fmaf(h[4 * j + 1], t4.y, dot);
                                    dot = fmaf(h[4 * j + 2], t4.z, dot); dot = fmaf(h[4 * j + 3], t4.w, dot);
                                }
                            }
                        }
                    }
                    if (!last) {
                        fence_proxy_async();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&bars->actready[sub]));
                    } else {
                        // sdf = dot over all 256 columns: the two warps of the quadrant combine through shared memory
                        const uint32_t slot = sPart + (uint32_t)(sub * 128 + r) * 4u;
                        if (half == 1) asm volatile("st.shared.f32 [%0], %1;" ::"r"(slot), "f"(dot) : "memory");
                        named_bar_sync(1 + q, 64);
                        if (half == 0) {
                            float other;
                            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(other) : "r"(slot) : "memory");
                            const int64_t grow = tile * kTileRows + sub * 128 + r;
                            float sdf = dot + other + __ldg(P.b_last);
                            if (P.clamp_radius > 0.f) {
                                const float* p3 = xc[sub];
                                sdf = fminf(sdf, P.sphere_scale * (P.clamp_radius - sqrtf(p3[0] * p3[0] + p3[1] * p3[1] + p3[2] * p3[2])));
                            }
                            if (grow < M) sdf_out[grow] = sdf;
                        }
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

// Wp[(l * 256 + r) * 256 + k] = fp16(W_l[r, k] * scale_l), zero padded; bp[l * 256 + r] = b_l[r]; wl[k] = W_last[0, k]
struct PackArgs {
    int L;
    const float* W[kMaxL]; const float* b[kMaxL];
    int out[kMaxL], in[kMaxL]; int64_t ldw[kMaxL]; float scale[kMaxL];
    const float* w_last; int in_last;
};
__global__ void k_pack_fused(const __grid_constant__ PackArgs a, __half* __restrict__ Wp, float* __restrict__ bp, float* __restrict__ wl) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nW = (int64_t)a.L * 65536;
    if (i < nW) {
        const int l = (int)(i >> 16), r = (int)((i >> 8) & 255), k = (int)(i & 255);
        const float v = (r < a.out[l] && k < a.in[l]) ? a.W[l][(int64_t)r * a.ldw[l] + k] * a.scale[l] : 0.f;
        Wp[i] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    } else if (i < nW + (int64_t)a.L * 256) {
        const int64_t t = i - nW;
        const int l = (int)(t >> 8), r = (int)(t & 255);
        bp[t] = r < a.out[l] ? a.b[l][r] : 0.f;
    } else if (i < nW + (int64_t)a.L * 256 + 256) {
        const int k = (int)(i - nW - (int64_t)a.L * 256);
        wl[k] = k < a.in_last ? a.w_last[k] : 0.f;
    }
}

inline size_t fused_workspace_bytes(int L) { return (size_t)L * 131072 + (size_t)L * 1024 + 1024; }

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
