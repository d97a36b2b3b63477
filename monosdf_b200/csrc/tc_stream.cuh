// k_tc_stream: one MLP layer of a backward-side sweep (reverse / tangent / backward) with EVERY operand fed by TMA.
//
//   C[m, n] = epi( sum_k A[m,k] W[n,k] ; R_0[m,n], R_1[m,n], R_2[m,n] )
//
// Same math and outputs as k_tc_gemm (tc_gemm.cuh) under the epilogues that read matrices back (sigma source h, the
// tangent t, the reverse-sweep state a), but those read-back operands no longer travel through LDG + register / cp.async
// prefetch rings + a shared-memory transpose per chunk (ptxas tracks all register prefetches with one scoreboard, so that
// ring was one chunk deep in effect and the kernels sat at 4.5 TB/s: profiles/README.md, round 1).  Here a second
// producer thread streams them as [128 rows x 64 columns] TMA boxes (SWIZZLE_128B) into a ring of their own, two to four
// boxes = four to eight 32-column chunks ahead of the epilogue, and an epilogue thread -- which owns one ROW of the tile,
// like its TMEM lane -- reads its 64 bytes of a box directly (16-byte pieces, swizzle p ^ (row & 7): conflict free), no
// transpose.  Room for the rings comes from streaming the layer's weights per 64-column k-block with the A tile (they are
// L2 resident: +128 KB of L2 -> SM traffic per 128-row tile) instead of keeping all of them resident in shared memory.
//
// 640 threads: warp 0 = A / W producer, warp 1 = MMA issuer (tcgen05.mma 128 x N x 16, double-buffered TMEM
// accumulators), warp 2 = read-back operand producer, warps 4-19 = epilogue (no prefetch rings to hold: 104 registers do).
#pragma once
#include <stdlib.h>

#include "tc_gemm.cuh"

namespace msdf_tc {

constexpr int kMaxOps = 3;
constexpr int kStreamEpiWarps = 16;                       // 4 per TMEM lane quadrant: group g takes chunks g and g + 4
constexpr int kStreamThreads = (4 + kStreamEpiWarps) * 32;
constexpr int kStreamRegsLight = 40, kStreamRegsEpi = 104;  // 128 x 40 + 512 x 104 <= 640 x 96 (the launch allocation)
constexpr int kMaxBoxes = 4;
constexpr uint32_t kBoxBytes = BM * 128;                 // 128 rows x 64 columns x 2 bytes
constexpr uint32_t kStreamSlot = 2048;                   // per epilogue warp: one 32 x 32 block of 16-bit values

struct StreamBarriers {
    uint64_t sfull[2], sempty[2], tfull[2], tempty[2], ofull[kMaxOps][kMaxBoxes], oempty[kMaxOps][kMaxBoxes];
    uint32_t tmem_base, pad;
};

struct OpMaps { CUtensorMap m[kMaxOps]; };

// what an epilogue thread needs to read its row of the staged operand boxes of the current chunk
struct OpRow {
    uint32_t row[kMaxOps];     // shared-memory address of this thread's 128-byte row in operand o's box
    uint32_t sw;               // row & 7
    uint32_t pc0;              // first 16-byte piece of the chunk inside the row: 0 or 4
    // columns 8 p .. 8 p + 7 of the chunk (p = 0..3), operand o, packed 16-bit pairs
    __device__ __forceinline__ void piece(int o, int p, uint32_t w[4]) const {
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
                     : "r"(row[o] + (((pc0 + (uint32_t)p) ^ sw) << 4)) : "memory");
    }
};

// Epilogue concept (beside N, colvec()): static constexpr int kOps; void chunk_smem(const WarpIO&, int n0, float v[32], const OpRow&) const
template <class Epi>
__global__ void __launch_bounds__(kStreamThreads, 1)
k_tc_stream(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW, const __grid_constant__ OpMaps ops,
            int64_t M, int BN, int KB, int nboxes, int a_fmt, int bk, Epi epi) {
    // bk: columns per A / W k-block: 64 (SWIZZLE_128B rows) or 32 (SWIZZLE_64B rows: half the stage, room for deeper operand rings)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
    constexpr int kOps = Epi::kOps;
    const uint32_t a_block = (uint32_t)BM * (uint32_t)bk * 2u, w_block = (uint32_t)BN * (uint32_t)bk * 2u;
    const uint32_t stage_bytes = a_block + w_block;             // A k-block + W k-block
    const uint32_t sS = base;                                   // 2 stages
    const uint32_t sO = sS + 2u * stage_bytes;                  // kOps x nboxes boxes
    const uint32_t sE = sO + (uint32_t)(kOps * nboxes) * kBoxBytes;   // staging slots of the epilogue warps (output stores): 2 KB each
    const uint32_t sV = sE + kStreamEpiWarps * kStreamSlot;     // per-column vector, 256 floats
    StreamBarriers* bars = reinterpret_cast<StreamBarriers*>(gen_base + 2u * stage_bytes + (size_t)(kOps * nboxes) * kBoxBytes +
                                                             kStreamEpiWarps * kStreamSlot + kColVecBytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t num_tiles = (M + BM - 1) / BM;
    const int chunks = (BN + 31) / 32;
    const int nbx = (chunks + 1) / 2;                           // operand boxes per tile

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapW);
        for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&bars->sfull[s]), 1); mbar_init(smem_u32(&bars->sempty[s]), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(&bars->tfull[a]), 1); mbar_init(smem_u32(&bars->tempty[a]), kStreamEpiWarps * 32); }
        for (int o = 0; o < kMaxOps; ++o)
            for (int b = 0; b < kMaxBoxes; ++b) { mbar_init(smem_u32(&bars->ofull[o][b]), 1); mbar_init(smem_u32(&bars->oempty[o][b]), kStreamEpiWarps); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(&bars->tmem_base), kTmemCols);
    if (threadIdx.x >= 128) {
        const float* cv = epi.colvec();
        const int j = (int)threadIdx.x - 128;
        if (j < 256) {
            const float x = (cv != nullptr && j < epi.N) ? __ldg(cv + j) : 0.f;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(sV + (uint32_t)j * 4u), "f"(x) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kStreamRegsLight));
        if (warp == 0) {
            if (lane == 0) {
                // ---- producer 1: per tile and k-block the A tile's and the weights' 64 columns
                int s = 0; uint32_t ph = 0;
                for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(smem_u32(&bars->sempty[s]), ph ^ 1u);
                        mbar_expect_tx(smem_u32(&bars->sfull[s]), stage_bytes);
                        const uint32_t st = sS + (uint32_t)s * stage_bytes;
                        tma_load_2d(st, &mapA, smem_u32(&bars->sfull[s]), kb * bk, (int)(tile * BM));
                        tma_load_2d(st + a_block, &mapW, smem_u32(&bars->sfull[s]), kb * bk, 0);
                        if (++s == 2) { s = 0; ph ^= 1u; }
                    }
            }
            __syncwarp();
        } else if (warp == 1) {
            if (lane == 0) {
                // ---- MMA issuer
                const uint32_t idesc = instr_desc(BM, BN, 0, 0, a_fmt, a_fmt);
                int s = 0; uint32_t ph = 0; uint32_t it = 0;
                for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                    const uint32_t a = it & 1u, aph = (it >> 1) & 1u;
                    mbar_wait(smem_u32(&bars->tempty[a]), aph ^ 1u);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + a * 256u;
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(smem_u32(&bars->sfull[s]), ph);
                        tc_fence_after();
                        const uint32_t st = sS + (uint32_t)s * stage_bytes;
                        // K-major swizzled atoms of 8 rows: 1024 B apart for 128-byte rows, 512 B for 64-byte rows
                        const uint64_t swz = bk == 64 ? 0ull : (((uint64_t)4 << 61) ^ ((uint64_t)2 << 61));   // layout_type_ 2 -> 4 (SWIZZLE_64B)
                        const uint32_t sbo = bk == 64 ? 1024u : 512u;
                        for (int k = 0; k < bk / UMMA_K; ++k) {
                            const uint64_t da = smem_desc(st + k * (UMMA_K * 2), 16, sbo) ^ swz;
                            const uint64_t db = smem_desc(st + a_block + k * (UMMA_K * 2), 16, sbo) ^ swz;
                            umma_bf16(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        umma_commit(smem_u32(&bars->sempty[s]));
                        if (++s == 2) { s = 0; ph ^= 1u; }
                    }
                    umma_commit(smem_u32(&bars->tfull[a]));
                }
            }
            __syncwarp();
        } else if (warp == 2) {
            if (lane == 0 && kOps > 0) {
                // ---- producer 2: the read-back operands' boxes, in the order the epilogue consumes them
                for (int o = 0; o < kOps; ++o) tma_prefetch_desc(&ops.m[o]);
                uint32_t g = 0;
                for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
                    for (int b = 0; b < nbx; ++b, ++g) {
                        const uint32_t slot = g % (uint32_t)nboxes, ph = (g / (uint32_t)nboxes) & 1u;
                        for (int o = 0; o < kOps; ++o) {
                            mbar_wait(smem_u32(&bars->oempty[o][slot]), ph ^ 1u);
                            mbar_expect_tx(smem_u32(&bars->ofull[o][slot]), kBoxBytes);
                            tma_load_2d(sO + (uint32_t)(o * nboxes + (int)slot) * kBoxBytes, &ops.m[o], smem_u32(&bars->ofull[o][slot]), b * BK,
                                        (int)(tile * BM));
                        }
                    }
            }
            __syncwarp();
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kStreamRegsEpi));
        // ---- epilogue: TMEM lane quadrant = warp % 4; the four warps of a quadrant take chunks g and g + 4 (g = group);
        // every warp follows every operand box (waits for it, releases it) so that the ring's phases stay in step
        const int q = warp & 3, g4 = (warp - 4) >> 2;
        WarpIO io{sE + (uint32_t)(warp - 4) * kStreamSlot, lane, (int64_t)blockIdx.x * BM + q * 32, M, sV};
        io.init();
        io.abuf = 0u;
        io.flip_mask = 0u;                                       // 2 KB slot: no alternation (a __syncwarp follows every chunk)
        OpRow orow;
        orow.sw = (uint32_t)(lane & 7);                          // (q * 32 + lane) & 7
        orow.pc0 = (uint32_t)(g4 & 1) * 4u;
        const uint32_t rowoff = (uint32_t)(q * 32 + lane) * 128u;
        uint32_t it = 0, g = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const uint32_t a = it & 1u, aph = (it >> 1) & 1u;
            io.retile(tile * BM + q * 32);
            const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + a * 256u;
            mbar_wait(smem_u32(&bars->tfull[a]), aph);
            tc_fence_after();
#pragma unroll 1
            for (int bb = 0; bb < nbx; bb += 2) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int b = bb + k;
                    if (b >= nbx) break;
                    // box b holds chunks 2b, 2b+1: they belong to groups (2b) % 4 and (2b+1) % 4, i.e. to this warp iff
                    // (b & 1) == (g4 >> 1); its chunk is then 2b + (g4 & 1) = g4 + 4 (b >> 1)
                    const bool mine = (b & 1) == (g4 >> 1);
                    const int c = 2 * b + (g4 & 1);
                    const uint32_t slot = (g + (uint32_t)k) % (uint32_t)nboxes, ph = ((g + (uint32_t)k) / (uint32_t)nboxes) & 1u;
#pragma unroll
                    for (int o = 0; o < kOps; ++o) {
                        mbar_wait(smem_u32(&bars->ofull[o][slot]), ph);
                        orow.row[o] = sO + (uint32_t)(o * nboxes + (int)slot) * kBoxBytes + rowoff;
                    }
                    if (mine && c < chunks) {
                        // (the other three warps of the scheduler hide the TMEM latency: no register double buffer)
                        float v[32];
                        tmem_ld32(tacc + (uint32_t)c * 32u, v);
                        epi.chunk_smem(io, c * 32, v, orow);
                    }
                    // this warp is done with the boxes (chunk_smem reads its operands before it stores)
                    __syncwarp();
                    if (lane == 0) {
#pragma unroll
                        for (int o = 0; o < kOps; ++o) mbar_arrive(smem_u32(&bars->oempty[o][slot]));
                    }
                }
                g += 2;
            }
            g -= (uint32_t)(nbx & 1);                            // an odd box count advanced g one too far
            tc_fence_before();
            mbar_arrive(smem_u32(&bars->tempty[a]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

// C = epi(A W^T; R_o): A [M, Kp] (ld lda), W [BN, Kp] (ld ldw), read-back operands R_o [M, >= BN] 16-bit (ld ldr[o]), all
// K-major with leading dimensions that are multiples of 64 columns.
template <class Epi>
int launch_stream(const void* A, int a_fmt, int64_t lda, int64_t M, int Kp, const void* W, int64_t ldw, int BN, const void* const* R,
                  const int* r_fmt, const int64_t* ldr, const Epi& epi, cudaStream_t st, const char* what) {
    if (M <= 0) return MSDF_OK;
    if (Kp % 64 != 0 || Kp <= 0 || Kp > 640 || BN % 16 != 0 || BN < 16 || BN > 256) {
        msdf_set_error("%s: streamed tensor-core GEMM needs K %% 64 == 0 (<= 640) and N %% 16 == 0 (<= 256); got K=%d N=%d", what, Kp, BN);
        return MSDF_ERR_ARG;
    }
    constexpr int kOps = Epi::kOps;
    CUtensorMap mA, mW;
    OpMaps ops{};
    // three read-back operands: 32-column k-blocks (2 x 24 KB of stages instead of 2 x 48 KB) make room for rings of 3 boxes
    static const int bk_env = [] { const char* e = getenv("MSDF_STREAM_BK"); return e ? atoi(e) : 0; }();
    const int bk = (bk_env == 32 || bk_env == 64) ? bk_env : (kOps >= 3 ? 32 : 64);
    int rc = make_map(&mA, A, a_fmt, M, Kp, lda, BM, what, bk); if (rc) return rc;
    rc = make_map(&mW, W, a_fmt, BN, Kp, ldw, BN, what, bk); if (rc) return rc;
    const int cols = (BN + 63) / 64 * 64;
    for (int o = 0; o < kOps; ++o) { rc = make_map(&ops.m[o], R[o], r_fmt[o], M, cols, ldr[o], BM, what); if (rc) return rc; }
    const int KB = Kp / bk;
    const size_t stage_bytes = (size_t)(BM + BN) * bk * 2;
    const size_t fixed = 1024 + sizeof(StreamBarriers) + kStreamEpiWarps * kStreamSlot + kColVecBytes + 2 * stage_bytes;
    int nboxes = kOps > 0 ? (int)((227 * 1024 - fixed) / ((size_t)kOps * kBoxBytes)) : 0;
    if (nboxes > kMaxBoxes) nboxes = kMaxBoxes;
    { static const int cap = [] { const char* e = getenv("MSDF_STREAM_BOXES"); return e ? atoi(e) : kMaxBoxes; }(); if (nboxes > cap && cap >= 2) nboxes = cap; }   // experiment knob: ring depth
    if (kOps > 0 && nboxes < 2) { msdf_set_error("%s: no room for the operand rings", what); return MSDF_ERR_ARG; }
    const size_t smem = fixed + (size_t)kOps * nboxes * kBoxBytes;
    static bool attr_set = false;   // per instantiation
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_tc_stream<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { msdf_set_error("%s: cannot opt in to 227 KB shared memory: %s", what, cudaGetErrorString(e)); return MSDF_ERR_CUDA; }
        attr_set = true;
    }
    const int64_t tiles = (M + BM - 1) / BM;
    const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    const int prof = msdf_prof_begin(MSDF_PROF_TC_STREAM, 2.0 * (double)M * (double)BN * (double)Kp, st,
                                     (double)M * 2.0 * ((double)Kp + (double)epi.N * (double)(kOps + Epi::kStores)));
    k_tc_stream<Epi><<<grid, kStreamThreads, smem, st>>>(mA, mW, ops, M, BN, KB, nboxes, a_fmt, bk, epi);
    msdf_prof_end(prof, st);
    MSDF_COUNT_LAUNCH();
    MSDF_CHECK_LAUNCH(what);
    return MSDF_OK;
}

}  // namespace msdf_tc
