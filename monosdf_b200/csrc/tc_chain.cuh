// k_tc_chain: a whole backward-side SWEEP of the SDF network over a chunk in one launch, the chain variable on chip.
//
// The per-layer kernels (k_tc_stream) read the sweep's state back from HBM as the A operand of every layer although the
// previous launch has just written it: 1 of the 3 matrices a reverse / tangent layer moves.  Here a CTA owns a 128-row
// tile through ALL layers of the sweep (the structure of the fused forward kernel, fused_mlp.cuh): the epilogue warps
// write the layer's output straight into shared memory in the UMMA SWIZZLE_128B K-major operand layout (the "chain"
// buffer, 4 k-blocks x 16 KB), the next layer's tcgen05.mma reads it from there, and the same bytes leave for HBM as TMA
// stores (the weight gradients and the backward sweep need every layer's state).  Per layer and point that is one
// read-back operand in and one state matrix out: 2 instead of 3 matrices, and the reverse sweep's start
// (a_{L-2} = w_last * sigma) no longer is a kernel of its own.
//
// 640 threads: warp 0 = weight producer (32-column SWIZZLE_64B k-blocks of W_l^T, three 16 KB stages, L2 resident), warp 1
// = MMA issuer (tcgen05.mma 128 x N x 16, A = chain, one 256-column TMEM accumulator per sub-tile), warp 2 = read-back
// operand producer ([128 x 64] boxes, ring of two), warp 3 = TMA store issuer, warps 4-19 = epilogue (thread = row, as in
// k_tc_stream).  A CTA works on TWO 128-row sub-tiles X / Y that ping-pong through the layers like in the fused forward
// kernel: while the tensor core runs layer l of one, the epilogue warps turn the other's accumulator into its next chain.
// (First version, one tile per CTA, MMA and epilogue alternating: 720 us per 262 144-point reverse sweep against 576 us
// of per-layer launches -- the serial schedule loses what the saved traffic wins.)
#pragma once
#include "tc_stream.cuh"

namespace msdf_tc {

constexpr int kChainMaxSteps = 9;
constexpr uint32_t kChainBytes = 4u * kBoxBytes;          // 128 rows x 256 columns x 2 bytes
constexpr int kChainBK = 32;                              // columns per weight k-block (SWIZZLE_64B rows)
constexpr uint32_t kChainWStage = 256u * 64u;             // a k-block of up to 256 weight rows
constexpr int kChainWStages = 3;
constexpr int kChainBoxes = 2;                            // (measured: 3 stages + 2 boxes 546 us, 2 stages + 3 boxes 555 us per 262 144 points)
constexpr uint32_t kChainSlot = 1024;                     // per epilogue warp: staging of the fp32 g0 columns

struct ChainStep {
    int BN;            // MMA N of the step (multiple of 16); 0: no MMA, the epilogue starts the chain from the column vector
    int KB;            // 64-column k-blocks of the chain the MMA reads
    int nb;            // read-back operand boxes of the step (0: none)
    int ob;            // 64-column boxes of the chain stored afterwards (0: the step does not write the chain)
    int N;             // valid output columns
    int dh;            // columns >= dh leave as fp32 rows of g0
    float hscale, qscale;
    float* g0;
};
struct ChainPlan { int S; int pad; int64_t ldg; ChainStep st[kChainMaxSteps]; };
struct ChainMaps { CUtensorMap w[kChainMaxSteps], op[kChainMaxSteps], out[kChainMaxSteps]; };
struct ChainBarriers {
    uint64_t wfull[kChainWStages], wempty[kChainWStages], ofull[kChainBoxes], oempty[kChainBoxes], chainready[2], accfull[2], storedone[2];
    uint32_t tmem_base, pad;
};

__device__ __forceinline__ void chain_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}

__device__ __forceinline__ float chain_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Reverse sweep (network.py:98-109 restated as an analytic sweep): a_{l-1} = (a_l W_l) sigma(h_l), fp16 chain.
//   step 0 (BN = 0):  a_{L-2}[m,n] = w_last[n] sigma(h_{L-1}[m,n])            (colvec = the sdf row of the last layer)
//   step s:           acc = chain W_l^T ;  columns < dh: acc qscale sigma(h_l hscale) -> chain, A[l-1];  columns >= dh: acc qscale -> g0
struct RevChainPolicy {
    static constexpr int kFmt = kF16;                      // chain, weights and read-back operand: forward-like fp16
    const float* colvec;
    // v: accumulators (or the column vector) of columns n0 .. n0 + 31 of this thread's row; on return the chain values
    __device__ __forceinline__ void chunk(const ChainStep& S, int64_t ldg, const WarpIO& io, int n0, float v[32], const OpRow& r) const {
        const int nv = S.N - n0 < 32 ? S.N - n0 : 32;
        const int nh = S.dh - n0 < nv ? S.dh - n0 : nv;     // columns that stay in the chain
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= S.qscale;
        if (nh < nv) io.store_f32_narrow8(S.g0, ldg, n0 - S.dh, v, nh > 0 ? nh : 0, nv);
        if (nh <= 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
            return;
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            uint32_t hw[4];
            r.piece(0, p, hw);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float s = 1.0f - chain_ex2(-144.26950408889634f * (WarpIO::unpack<kF16>(hw, j) * S.hscale));   // sigmoid(100 p) from h
                v[8 * p + j] = 8 * p + j < nh ? v[8 * p + j] * s : 0.f;
            }
        }
    }
};

template <class Pol>
__global__ void __launch_bounds__(kStreamThreads, 1)
k_tc_chain(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainPlan P, int64_t M, Pol pol) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sC = base;                                        // chains of sub-tiles X, Y: 4 k-blocks of [128 x 64] each
    const uint32_t sW = sC + 2u * kChainBytes;                       // weight stages
    const uint32_t sO = sW + (uint32_t)kChainWStages * kChainWStage; // operand ring
    const uint32_t sE = sO + (uint32_t)kChainBoxes * kBoxBytes;      // 1 KB slots (fp32 g0 columns)
    const uint32_t sV = sE + kStreamEpiWarps * kChainSlot;           // column vector, 256 floats
    ChainBarriers* bars = reinterpret_cast<ChainBarriers*>(gen_base + 2u * kChainBytes + kChainWStages * kChainWStage + kChainBoxes * kBoxBytes +
                                                           kStreamEpiWarps * kChainSlot + kColVecBytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t num_pairs = (M + 2 * BM - 1) / (2 * BM);
    const int S = P.S;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < S; ++s) { tma_prefetch_desc(&maps.w[s]); tma_prefetch_desc(&maps.op[s]); tma_prefetch_desc(&maps.out[s]); }
        for (int s = 0; s < kChainWStages; ++s) { mbar_init(smem_u32(&bars->wfull[s]), 1); mbar_init(smem_u32(&bars->wempty[s]), 1); }
        for (int b = 0; b < kChainBoxes; ++b) { mbar_init(smem_u32(&bars->ofull[b]), 1); mbar_init(smem_u32(&bars->oempty[b]), kStreamEpiWarps); }
        for (int u = 0; u < 2; ++u) {
            mbar_init(smem_u32(&bars->chainready[u]), kStreamEpiWarps);
            mbar_init(smem_u32(&bars->accfull[u]), 1);
            mbar_init(smem_u32(&bars->storedone[u]), 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(&bars->tmem_base), kTmemCols);
    if (threadIdx.x >= 128) {
        const int j = (int)threadIdx.x - 128;
        if (j < 256) {
            const float x = pol.colvec != nullptr ? __ldg(pol.colvec + j) : 0.f;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(sV + (uint32_t)j * 4u), "f"(x) : "memory");
        }
        // the chains' K padding must be finite from the first tile on
        for (uint32_t o = (uint32_t)j * 16u; o < 2u * kChainBytes; o += 512u * 16u)
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sC + o), "r"(0u) : "memory");
        fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kStreamRegsLight));
        if (warp == 0) {
            if (lane == 0) {
                // ---- weight producer: the k-blocks of a step once per sub-tile
                int st = 0; uint32_t ph = 0;
                for (int64_t pair = blockIdx.x; pair < num_pairs; pair += gridDim.x)
                    for (int s = 0; s < S; ++s) {
                        const int BN = P.st[s].BN, KB = P.st[s].KB * (BK / kChainBK);
                        if (BN <= 0) continue;
                        for (int u = 0; u < 2; ++u)
                            for (int kb = 0; kb < KB; ++kb) {
                                mbar_wait(smem_u32(&bars->wempty[st]), ph ^ 1u);
                                mbar_expect_tx(smem_u32(&bars->wfull[st]), (uint32_t)BN * (kChainBK * 2u));
                                tma_load_2d(sW + (uint32_t)st * kChainWStage, &maps.w[s], smem_u32(&bars->wfull[st]), kb * kChainBK, 0);
                                if (++st == kChainWStages) { st = 0; ph ^= 1u; }
                            }
                    }
            }
            __syncwarp();
        } else if (warp == 1) {
            if (lane == 0) {
                // ---- MMA issuer: one wait on chainready[u] per MMA step (the phase that produced its A operand)
                int st = 0; uint32_t ph = 0, cph[2] = {0u, 0u};
                constexpr uint64_t kSw64 = ((uint64_t)4 << 61) ^ ((uint64_t)2 << 61);      // layout_type_ 2 -> 4 (SWIZZLE_64B)
                for (int64_t pair = blockIdx.x; pair < num_pairs; pair += gridDim.x)
                    for (int s = 0; s < S; ++s) {
                        const int BN = P.st[s].BN, KB = P.st[s].KB * (BK / kChainBK);
                        if (BN <= 0) continue;
                        const uint32_t idesc = instr_desc(BM, BN, 0, 0, Pol::kFmt, Pol::kFmt);
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            mbar_wait(smem_u32(&bars->chainready[u]), cph[u]);
                            cph[u] ^= 1u;
                            tc_fence_after();
                            const uint32_t chain = sC + (uint32_t)u * kChainBytes, tmem_d = tmem_base + (uint32_t)u * 256u;
                            for (int kb = 0; kb < KB; ++kb) {
                                mbar_wait(smem_u32(&bars->wfull[st]), ph);
                                tc_fence_after();
#pragma unroll
                                for (int k = 0; k < kChainBK / UMMA_K; ++k) {
                                    // A: 32-column half (kb & 1) of the chain's 64-column SWIZZLE_128B k-block kb / 2
                                    const uint64_t da = smem_desc(chain + (uint32_t)(kb >> 1) * kBoxBytes + ((kb & 1) * 2 + k) * (UMMA_K * 2), 16, 1024);
                                    const uint64_t db = smem_desc(sW + (uint32_t)st * kChainWStage + k * (UMMA_K * 2), 16, 512) ^ kSw64;
                                    umma_bf16(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                                }
                                umma_commit(smem_u32(&bars->wempty[st]));
                                if (++st == kChainWStages) { st = 0; ph ^= 1u; }
                            }
                            umma_commit(smem_u32(&bars->accfull[u]));
                        }
                    }
            }
            __syncwarp();
        } else if (warp == 2) {
            if (lane == 0) {
                // ---- read-back operand producer, in the order the epilogue consumes the boxes
                uint32_t g = 0;
                for (int64_t pair = blockIdx.x; pair < num_pairs; pair += gridDim.x)
                    for (int s = 0; s < S; ++s)
                        for (int u = 0; u < 2; ++u)
                            for (int b = 0; b < P.st[s].nb; ++b, ++g) {
                                const uint32_t slot = g % (uint32_t)kChainBoxes, ph = (g / (uint32_t)kChainBoxes) & 1u;
                                mbar_wait(smem_u32(&bars->oempty[slot]), ph ^ 1u);
                                mbar_expect_tx(smem_u32(&bars->ofull[slot]), kBoxBytes);
                                tma_load_2d(sO + slot * kBoxBytes, &maps.op[s], smem_u32(&bars->ofull[slot]), b * BK, (int)((pair * 2 + u) * BM));
                            }
            }
            __syncwarp();
        } else {
            if (lane == 0) {
                // ---- store warp: every phase that wrote a chain is followed by its TMA stores; storedone[u] tells the
                // epilogue warps that the chain has been read and may be overwritten
                uint32_t cph[2] = {0u, 0u};
                for (int64_t pair = blockIdx.x; pair < num_pairs; pair += gridDim.x)
                    for (int s = 0; s < S; ++s) {
                        const int ob = P.st[s].ob;
                        if (ob <= 0) continue;
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            mbar_wait(smem_u32(&bars->chainready[u]), cph[u]);
                            cph[u] ^= 1u;
                            for (int b = 0; b < ob; ++b)
                                chain_store_2d(&maps.out[s], sC + (uint32_t)u * kChainBytes + (uint32_t)b * kBoxBytes, b * BK, (int)((pair * 2 + u) * BM));
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                            mbar_arrive(smem_u32(&bars->storedone[u]));
                        }
                    }
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            }
            __syncwarp();
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kStreamRegsEpi));
        // ---- epilogue: TMEM lane quadrant = warp % 4 = 32 rows of a sub-tile; group g4 takes chunk 2 b + (g4 & 1) of the boxes
        // b with (b & 1) == g4 >> 1; every warp follows every operand box so that the ring's phases stay in step
        const int q = warp & 3, g4 = (warp - 4) >> 2;
        WarpIO io{sE + (uint32_t)(warp - 4) * kChainSlot, lane, (int64_t)blockIdx.x * BM + q * 32, M, sV};
        io.init();
        io.abuf = 0u;
        io.flip_mask = 0u;
        OpRow orow;
        orow.sw = (uint32_t)(lane & 7);
        orow.pc0 = (uint32_t)(g4 & 1) * 4u;
        const uint32_t rowoff = (uint32_t)(q * 32 + lane) * 128u;
        uint32_t g = 0, aph0 = 0, aph1 = 0, sph0 = 0, sph1 = 0;
        bool first0 = true, first1 = true;
        for (int64_t pair = blockIdx.x; pair < num_pairs; pair += gridDim.x) {
#pragma unroll 1
            for (int s = 0; s < S; ++s) {
                const ChainStep& St = P.st[s];
                const int BN = St.BN;
                const bool writes = St.ob > 0;
                const int cols = BN > 0 ? BN : St.N;
                const int chunks = (cols + 31) / 32, nbx = (chunks + 1) / 2;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    io.retile((pair * 2 + u) * BM + q * 32);
                    const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)u * 256u;
                    const uint32_t chain = sC + (uint32_t)u * kChainBytes;
                    if (BN > 0) {
                        mbar_wait(smem_u32(&bars->accfull[u]), u == 0 ? aph0 : aph1);
                        if (u == 0) aph0 ^= 1u; else aph1 ^= 1u;
                        tc_fence_after();
                    }
                    if (writes) {
                        // the previous contents of this chain have been stored (and, before that, read by their MMA: accfull)
                        const bool first = u == 0 ? first0 : first1;
                        if (!first) {
                            mbar_wait(smem_u32(&bars->storedone[u]), u == 0 ? sph0 : sph1);
                            if (u == 0) sph0 ^= 1u; else sph1 ^= 1u;
                        }
                        if (u == 0) first0 = false; else first1 = false;
                    }
#pragma unroll 1
                    for (int b = 0; b < nbx; ++b) {
                        const bool mine = (b & 1) == (g4 >> 1);
                        const int c = 2 * b + (g4 & 1);
                        uint32_t slot = 0;
                        if (St.nb > 0) {
                            slot = g % (uint32_t)kChainBoxes;
                            mbar_wait(smem_u32(&bars->ofull[slot]), (g / (uint32_t)kChainBoxes) & 1u);
                            orow.row[0] = sO + slot * kBoxBytes + rowoff;
                            ++g;
                        }
                        if (mine && c < chunks) {
                            float v[32];
                            if (BN > 0) tmem_ld32(tacc + (uint32_t)c * 32u, v);
                            else io.colvec(c * 32, v);
                            pol.chunk(St, P.ldg, io, c * 32, v, orow);
                            if (writes) {
                                const uint32_t dst = chain + (uint32_t)(c >> 1) * kBoxBytes + rowoff;
#pragma unroll
                                for (int p = 0; p < 4; ++p) {
                                    uint32_t w[4];
#pragma unroll
                                    for (int j = 0; j < 4; ++j) w[j] = WarpIO::pack2<Pol::kFmt>(v[8 * p + 2 * j], v[8 * p + 2 * j + 1]);
                                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((orow.pc0 + (uint32_t)p) ^ orow.sw) * 16u),
                                                 "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
                                }
                            }
                        }
                        __syncwarp();
                        if (St.nb > 0 && lane == 0) mbar_arrive(smem_u32(&bars->oempty[slot]));
                    }
                    tc_fence_before();
                    if (writes) {
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&bars->chainready[u]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

constexpr size_t kChainSmem = 1024 + 2 * kChainBytes + kChainWStages * kChainWStage + kChainBoxes * kBoxBytes + kStreamEpiWarps * kChainSlot +
                              kColVecBytes + sizeof(ChainBarriers);

template <class Pol>
int launch_chain(const ChainMaps& maps, const ChainPlan& plan, int64_t M, const Pol& pol, double flops, double bytes, cudaStream_t st,
                 const char* what) {
    if (M <= 0) return MSDF_OK;
    static_assert(kChainSmem <= 227 * 1024, "chain kernel shared memory");
    static bool attr_set = false;   // per instantiation
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_tc_chain<Pol>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmem);
        if (e != cudaSuccess) { msdf_set_error("%s: cannot opt in to %zu bytes of shared memory: %s", what, kChainSmem, cudaGetErrorString(e)); return MSDF_ERR_CUDA; }
        attr_set = true;
    }
    const int64_t pairs = (M + 2 * BM - 1) / (2 * BM);
    const int grid = (int)(pairs < sm_count() ? pairs : sm_count());
    const int prof = msdf_prof_begin(MSDF_PROF_TC_CHAIN, flops, st, bytes);
    k_tc_chain<Pol><<<grid, kStreamThreads, kChainSmem, st>>>(maps, plan, M, pol);
    msdf_prof_end(prof, st);
    MSDF_COUNT_LAUNCH();
    MSDF_CHECK_LAUNCH(what);
    return MSDF_OK;
}

}  // namespace msdf_tc
