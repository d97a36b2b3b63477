// tcgen05 / TMEM / TMA GEMM engine for the MLP sweeps (the tensor-core mode of the field kernels), sm_100a only.
//
// Two persistent, warp-specialised kernels, 1 CTA per SM:
//
//   k_tc_gemm   C[m, n] = epi( sum_k A[m,k] * W[n,k] )        A [M, K] 16-bit K-major (activations: fp16, adjoints:
//               bf16), W [N<=256, K<=320] in A's format, K-major: the layer's weights, loaded ONCE per CTA and kept
//               resident in shared memory while the CTA walks over its 128-row tiles of A (TMA ring).  Accumulators live
//               in TMEM, double buffered (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//               384 threads = 3 warpgroups: {TMA producer warp, MMA issuer warp, 2 idle} at 40 registers and 8 epilogue
//               warps at 232 (setmaxnreg).  Used by the forward / tangent sweeps (W = W_l) and the reverse / backward
//               sweeps (W = W_l^T, transposed once per call by the weight-prep kernel).
//
//   k_tc_wgrad  dW[i, j] += sum_m X[m,i] * Y[m,j]             X [M, <=256 per tile], Y [M, <=256], both MN-major
//               operands (the contraction runs over the rows = points), split over m across CTAs, fp32 atomics into
//               dW; 320 threads (producer, MMA issuer, 8 warps that convert the fp16 operand of a mixed-format product
//               to bf16 in shared memory, sum bias columns and run the epilogue).
//
// Shared-memory operand layout is the canonical UMMA SWIZZLE_128B layout written by TMA boxes of 64 16-bit elements
// (128 B) inner extent; smem/instruction descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp.
// Every mbarrier wait carries a clock watchdog that traps instead of hanging the GPU.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace msdf_tc {

constexpr int BM = 128;            // rows per tile == TMEM lanes
constexpr int BK = 64;             // bf16 per 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kEpiWarps = 8;
constexpr int kWgRows = 64;                           // k-rows (points) per k_tc_wgrad stage (measured per 256x256 pair: 32 rows 187 us, 48 rows 150, 64 rows 114, 96 rows 118)
constexpr int kWgEpiWarps = 16;                       // k_tc_wgrad: producer warp, MMA warp, 16 converter / epilogue warps (the
constexpr int kWgThreads = (2 + kWgEpiWarps) * 32;    // conversion and the bias sums are issue-latency bound: 2 warps per scheduler were not enough)
constexpr int kGemmThreads = (4 + kEpiWarps) * 32;    // k_tc_gemm: warpgroup 0 = {producer, MMA, 2 idle warps}, then 8 epilogue warps
constexpr int kRegsLight = 40, kRegsEpi = 232;        // setmaxnreg split of k_tc_gemm: 128 x 40 + 256 x 232 = 384 x 168
constexpr int kMaxStages = 8;
constexpr uint32_t kStageBytesA = BM * BK * 2;   // 16 KB
constexpr uint32_t kTmemCols = 512;

// ----------------------------------------------------------------------------------------------------------
// PTX wrappers
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(20000u)      // suspend-time hint (ns): the waiting warp sleeps instead of spinning
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();   // ~2 s: a protocol bug, fail loudly instead of hanging
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = lane = row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float v[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// split form: issue the load, later wait for it.  The wait names the registers as read-write operands so that every
// consumer of r[] is ordered after it.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t r[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32_wait(uint32_t r[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                   "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                   "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// shared-memory matrix descriptors (cute::UMMA::SmemDescriptor): SWIZZLE_128B, Blackwell version bit
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;     // version_ = 1
    d |= (uint64_t)2 << 61;     // layout_type_ = SWIZZLE_128B
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): 16-bit x 16-bit -> fp32, M x N, operand majors
// 16-bit operand formats of tcgen05.mma.kind::f16 (the a_format / b_format fields of the instruction descriptor).
// The field sweeps keep "forward-like" matrices (activations h, reverse-sweep state a, colour rows) in fp16
// (11 significant bits: 8x smaller sdf error than bf16) and the wide-range adjoint / tangent
// matrices in bf16.  Both operands of one MMA must have the SAME format (a mixed descriptor raises an illegal
// instruction on B200, measured), so the weights exist in both formats and the weight-gradient kernel converts its
// fp16 operand to bf16 in shared memory.
enum Fmt : int { kF16 = 0, kBF16 = 1 };
template <class E> struct FmtOf;
template <> struct FmtOf<__half> { static constexpr int value = kF16; };
template <> struct FmtOf<__nv_bfloat16> { static constexpr int value = kBF16; };

__host__ __device__ inline uint32_t instr_desc(int M, int N, int a_mn_major, int b_mn_major, int a_fmt, int b_fmt) {
    return (1u << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}


// ----------------------------------------------------------------------------------------------------------
// Epilogue I/O: TMEM hands every thread one ROW of the tile (lane = row), but global memory wants a warp to touch
// whole rows.  WarpIO transposes 32-row x 32-column blocks through a private, swizzled 4 KB shared-memory slot so
// that every global load / store / atomic of the epilogue is coalesced (8 rows x 64 B per bf16 request, one full
// 128-byte row per fp32 atomic request).
// ----------------------------------------------------------------------------------------------------------
struct WarpIO {
    uint32_t slot;      // shared-memory address of this warp's 4 KB staging slot
    int lane;
    int64_t row0;       // global row of lane 0
    int64_t M;          // rows of the problem (rows >= M are masked)
    uint32_t cvec;      // shared-memory address of the CTA's per-column vector (bias), zero padded to 256 entries
    // derived once per warp / tile (init()): everything the hot paths need without integer arithmetic per element
    uint32_t own[4];    // slot offsets of the 16-byte pieces 0..3 of this lane's own row (row-major side of the transpose)
    uint32_t trn;       // slot offset of piece lane&3 of row lane/4 (global side); rows +8 are 512 bytes further
    int rows_left;      // M - row0, clamped to [0, 32]
    mutable uint32_t flip;   // bf16 stagings alternate between the two 2 KB halves of the slot: one __syncwarp each
    uint32_t flip_mask;      // 2048; 0 for a 2 KB slot (k_tc_stream: a __syncwarp separates its consecutive stagings anyway)
    int64_t pf_row_shift;    // prefetch() addresses rows row0 + pf_row_shift (set while prefetching for the next tile)
    // Second prefetch path.  ptxas puts EVERY register prefetch (ld.global -> uint4) of the epilogue on one scoreboard
    // (tools/sass_scoreboards.py), so the staging store of a chunk waits for ALL outstanding loads, the refills issued
    // one chunk ago included: a register ring is one chunk deep whatever its size.  Odd ring slots therefore go through
    // cp.async into a per-warp landing buffer (abuf: 2 KB per operand, the layout unstage() writes), tracked by
    // cp.async.wait_all instead of that scoreboard: each mechanism then only ever waits for loads issued two chunks ago.
    uint32_t abuf;           // shared-memory address of the landing buffer, 0 = not available (wide-K launches)
    mutable int amode;       // 1 while the engine prefetches / consumes an odd slot
    mutable const uint4* qbase;   // register array of the slot in flight: (q - qbase) / 4 = operand index

    __device__ __forceinline__ void init() {
        const int x = (lane >> 1) & 3, r = lane >> 2, pc = lane & 3;
#pragma unroll
        for (int p2 = 0; p2 < 4; ++p2) own[p2] = (uint32_t)(lane * 64 + ((p2 ^ x) << 4));
        trn = (uint32_t)(r * 64 + ((pc ^ ((r >> 1) & 3)) << 4));
        flip = 0u;
        flip_mask = 2048u;
        pf_row_shift = 0;
        amode = 0;
        qbase = nullptr;
        retile(row0);
    }
    __device__ __forceinline__ void retile(int64_t new_row0) {
        row0 = new_row0;
        const int64_t left = M - row0;
        rows_left = left < 0 ? 0 : (left > 32 ? 32 : (int)left);
    }

    // b[j] = column vector entry n0 + j: 8 broadcast 16-byte shared loads (every lane reads the same address)
    __device__ __forceinline__ void colvec(int n0, float b[32]) const {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b[4 * i]), "=f"(b[4 * i + 1]), "=f"(b[4 * i + 2]), "=f"(b[4 * i + 3])
                         : "r"(cvec + (uint32_t)(n0 + 4 * i) * 4u));
    }

    __device__ __forceinline__ int64_t row() const { return row0 + lane; }
    __device__ __forceinline__ bool valid() const { return lane < rows_left; }

    // Coalesced read of the 32 x 32 bf16 block P[row0.., n0..n0+31] in two steps so that the global latency can be
    // hidden: prefetch() issues the loads (lane t fetches 16-byte piece t&3 of rows i*8 + t/4; predicated, rows >= M
    // are left undefined and must not be stored), unstage() transposes them through the slot so that
    // out[j] = P[row(), n0 + j].  Requires 16-byte aligned P + n0 and ld % 8 == 0.
    __device__ __forceinline__ uint32_t abuf_of(const uint4* q) const { return abuf + (uint32_t)(q - qbase) * 512u; }   // 4 uint4 = 2 KB
    __device__ __forceinline__ void async_ready() const {
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
    }
    __device__ __forceinline__ void prefetch(const void* Pv, int64_t ld, int n0, uint4 q[4]) const {
        const uint16_t* P = reinterpret_cast<const uint16_t*>(Pv);     // any 16-bit element type
        const int64_t base_row = row0 + pf_row_shift;        // pf_row_shift != 0: the same rows of a later tile
        const int64_t left64 = M - base_row;
        const int left = left64 < 0 ? 0 : (left64 > 32 ? 32 : (int)left64);
        const char* base = reinterpret_cast<const char*>(P + (base_row + (lane >> 2)) * ld + n0 + (lane & 3) * 8);
        const int64_t step = ld * 16;            // 8 rows, in bytes
        const int r = lane >> 2;
        if (amode) {
            // rows >= M: zero fill (source size 0; the address is kept inside the matrix)
            const uint32_t dst = abuf_of(q) + trn;
            __syncwarp();                         // every lane is done reading the buffer's previous contents
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool on = (i * 8 + r) < left;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)i * 512u), "l"(on ? base + i * step : reinterpret_cast<const char*>(P)),
                             "r"(on ? 16u : 0u) : "memory");
            }
            return;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t on = (i * 8 + r) < left ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];\n\t}"
                         : "=r"(q[i].x), "=r"(q[i].y), "=r"(q[i].z), "=r"(q[i].w)
                         : "l"(base + i * step), "r"(on));
        }
    }
    template <int F>
    static __device__ __forceinline__ float lo_of(uint32_t w) {
        return F == kBF16 ? __uint_as_float(w << 16) : __half2float(__ushort_as_half((unsigned short)(w & 0xffffu)));
    }
    template <int F>
    static __device__ __forceinline__ float hi_of(uint32_t w) {
        return F == kBF16 ? __uint_as_float(w & 0xffff0000u) : __half2float(__ushort_as_half((unsigned short)(w >> 16)));
    }
    template <int F>
    __device__ __forceinline__ void unstage(const uint4 q[4], float out[32]) const {
        uint32_t h;
        if (amode) {
            async_ready();
            h = abuf_of(q);
        } else {
            h = slot + flip;
            flip ^= flip_mask;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(h + trn + (uint32_t)i * 512u), "r"(q[i].x), "r"(q[i].y), "r"(q[i].z), "r"(q[i].w) : "memory");
            __syncwarp();
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            uint32_t w[4];
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(h + own[p]) : "memory");
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                out[p * 8 + 2 * j] = lo_of<F>(w[j]);
                out[p * 8 + 2 * j + 1] = hi_of<F>(w[j]);
            }
        }
    }
    // the two steps of unstage_packed() separately, for epilogues with several read-back operands that want to walk
    // them 8 columns at a time instead of holding every operand's 16 packed words: stage() moves a prefetched block into
    // the slot (or returns the cp.async landing buffer) and returns its address, piece() reads columns 8p .. 8p+7 of
    // this lane's row.  Consecutive stage() calls alternate between the slot's two halves: at most two staged blocks
    // are live; a __syncwarp() must separate the last piece() of a block from the second stage() after it.
    __device__ __forceinline__ uint32_t stage(const uint4 q[4]) const {
        if (amode) { async_ready(); return abuf_of(q); }
        const uint32_t h = slot + flip;
        flip ^= flip_mask;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(h + trn + (uint32_t)i * 512u), "r"(q[i].x), "r"(q[i].y), "r"(q[i].z), "r"(q[i].w) : "memory");
        __syncwarp();
        return h;
    }
    __device__ __forceinline__ void piece(uint32_t h, int p, uint32_t w[4]) const {
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(h + own[p]) : "memory");
    }
    // same, but the row stays packed: w[p * 4 + j] holds columns p*8 + 2j (low half) and p*8 + 2j + 1 (high half)
    __device__ __forceinline__ void unstage_packed(const uint4 q[4], uint32_t w[16]) const {
        uint32_t h;
        if (amode) {
            async_ready();
            h = abuf_of(q);
        } else {
            h = slot + flip;
            flip ^= flip_mask;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(h + trn + (uint32_t)i * 512u), "r"(q[i].x), "r"(q[i].y), "r"(q[i].z), "r"(q[i].w) : "memory");
            __syncwarp();
        }
#pragma unroll
        for (int p = 0; p < 4; ++p)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[4 * p]), "=r"(w[4 * p + 1]), "=r"(w[4 * p + 2]), "=r"(w[4 * p + 3]) : "r"(h + own[p]) : "memory");
    }
    template <int F>
    static __device__ __forceinline__ float unpack(const uint32_t w[16], int j) {      // column j of a packed row
        return (j & 1) ? hi_of<F>(w[j >> 1]) : lo_of<F>(w[j >> 1]);
    }
    template <int F>
    __device__ __forceinline__ void load(const void* P, int64_t ld, int n0, float out[32]) const {
        uint4 q[4];
        prefetch(P, ld, n0, q);
        unstage<F>(q, out);
    }

    // P[row(), n0 + j] = v[j] for j < nvalid (<= 32); rows >= M are skipped
    template <int F>
    static __device__ __forceinline__ uint32_t pack2(float lo, float hi) {
        if (F == kBF16) {
            const __nv_bfloat162 hh = __floats2bfloat162_rn(lo, hi);
            return *reinterpret_cast<const uint32_t*>(&hh);
        }
        uint32_t w;                           // fp16 saturates to +-65504 instead of overflowing to inf
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(hi), "f"(lo));
        return w;
    }
    template <int F>
    __device__ __forceinline__ void store(void* P, int64_t ld, int n0, const float v[32], int nvalid) const {
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) w[j] = pack2<F>(v[2 * j], v[2 * j + 1]);
        store_packed(P, ld, n0, w, nvalid);
    }
    // same with the row already packed (w[j] = columns 2j, 2j+1): lets an epilogue with two outputs pack one of them
    // pair by pair while it computes, instead of holding 32 more floats
    __device__ __forceinline__ void store_packed(void* Pv, int64_t ld, int n0, const uint32_t w[16], int nvalid) const {
        uint16_t* P = reinterpret_cast<uint16_t*>(Pv);
        const uint32_t h = slot + flip;
        flip ^= flip_mask;
#pragma unroll
        for (int p = 0; p < 4; ++p)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(h + own[p]), "r"(w[4 * p]), "r"(w[4 * p + 1]), "r"(w[4 * p + 2]), "r"(w[4 * p + 3]) : "memory");
        __syncwarp();
        char* base = reinterpret_cast<char*>(P + (row0 + (lane >> 2)) * ld + n0 + (lane & 3) * 8);
        const int64_t step = ld * 16;
        const int r = lane >> 2;
        const int pv = nvalid - (lane & 3) * 8;      // valid columns of this lane's piece
        if (nvalid == 32) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint4 q;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(h + trn + (uint32_t)i * 512u) : "memory");
                const uint32_t on = (i * 8 + r) < rows_left ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p st.global.v4.b32 [%0], {%1, %2, %3, %4};\n\t}"
                             ::"l"(base + i * step), "r"(q.x), "r"(q.y), "r"(q.z), "r"(q.w), "r"(on) : "memory");
            }
        } else {
            store_ragged(h + trn, base, step, r, rows_left, pv);
        }
    }
    // ragged last chunk of a layer (e.g. the 217-column layer before the skip concat): compact and out of line
    static __device__ __noinline__ void store_ragged(uint32_t src, char* base, int64_t step, int r, int rows_left, int pv) {
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
            uint4 q;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(src + (uint32_t)i * 512u) : "memory");
            if ((i * 8 + r) >= rows_left || pv <= 0) continue;
            uint16_t* dst = reinterpret_cast<uint16_t*>(base + i * step);
            if (pv >= 8) { *reinterpret_cast<uint4*>(dst) = q; continue; }
            uint32_t w = q.x;
#pragma unroll 1
            for (int j = 0; j < pv; ++j) {
                if (j == 2) w = q.y; else if (j == 4) w = q.z; else if (j == 6) w = q.w;
                dst[j] = (uint16_t)((j & 1) ? (w >> 16) : (w & 0xffffu));
            }
        }
    }

    // fp32 block [32][32]: element (r, c) at r*128 + ((c ^ r) & 31)*4  (row-wise and column-wise conflict free)
    __device__ __forceinline__ void stage_f32(const float v[32]) const {
        __syncwarp();          // the bf16 stagings do not end with a barrier; this one uses the whole slot
        flip = 0u;
#pragma unroll
        for (int c = 0; c < 32; ++c)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(slot + (uint32_t)(lane * 128 + (((c ^ lane) & 31) << 2))), "f"(v[c]) : "memory");
        __syncwarp();
    }
    __device__ __forceinline__ float staged(int r) const {   // element (r, lane)
        float x;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(slot + (uint32_t)(r * 128 + (((lane ^ r) & 31) << 2))) : "memory");
        return x;
    }

    // atomicAdd(addr(row0 + r, c), v_r[c]) for every row below row_limit; addr returns nullptr for masked elements.
    // One request per row: 32 lanes = 32 consecutive columns (coalesced when addr is contiguous in c).
    template <class AddrFn>
    __device__ __forceinline__ void atomic_add(const float v[32], int64_t row_limit, AddrFn addr) const {
        stage_f32(v);
        for (int r = 0; r < 32; ++r) {
            const int64_t gr = row0 + r;
            if (gr >= row_limit) break;
            const float x = staged(r);
            float* p = addr(gr, lane);
            if (p != nullptr) atomicAdd(p, x);
        }
        __syncwarp();
    }

    // Same for a full 32-column chunk whose rows are 16-byte aligned: vector reductions (red.global.add.v4.f32), 4 rows x
    // 128 contiguous bytes per warp instruction -- a quarter of the lane operations of the scalar form, which is what
    // the weight-gradient kernel's flush is bound by (65536 reductions per CTA: about half of a launch, measured).
    // rowptr(gr) returns the address of column n0 of global row gr (nullptr: skip the row).
    template <class RowFn>
    __device__ __forceinline__ void atomic_add_v4(const float v[32], int64_t row_limit, RowFn rowptr) const {
        __syncwarp();          // the 16-bit stagings do not end with a barrier; this one uses the whole slot
        flip = 0u;
        const uint32_t sw = (uint32_t)(lane & 7);
#pragma unroll
        for (int p = 0; p < 8; ++p)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(slot + (uint32_t)lane * 128u + (((uint32_t)p ^ sw) << 4)),
                         "f"(v[4 * p]), "f"(v[4 * p + 1]), "f"(v[4 * p + 2]), "f"(v[4 * p + 3]) : "memory");
        __syncwarp();
        const int piece = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + (lane >> 3);
            const int64_t gr = row0 + r;
            float a, b, c, d;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d)
                         : "r"(slot + (uint32_t)r * 128u + (((uint32_t)piece ^ (uint32_t)(r & 7)) << 4)) : "memory");
            float* ptr = gr < row_limit ? rowptr(gr) : nullptr;
            if (ptr != nullptr)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ptr + piece * 4), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
        }
        __syncwarp();
    }

    // C[row, c0 + j] = (C[row, c0 + j] +) v[j] for jlo <= j < jhi and rows < M: fp32, one contiguous row per request.
    // Rare path (first layer / skip columns): kept out of line so that it does not bloat the unrolled chunk loop.
    __device__ __forceinline__ void store_f32(float* C, int64_t ldc, int c0, const float v[32], int jlo, int jhi, bool accum) const {
        float tmp[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) tmp[j] = v[j];
        warp_store_f32(slot, lane, row0, rows_left, C, ldc, c0, tmp, jlo, jhi, accum);
        flip = 0u;
    }
    // The same for a 2 KB slot (k_tc_stream): two passes of 16 columns; every global store covers two rows x 64 B.
    __device__ __forceinline__ void store_f32_narrow(float* C, int64_t ldc, int c0, const float v[32], int jlo, int jhi) const {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (jhi <= 16 * h || jlo >= 16 * h + 16) continue;      // warp-uniform
            __syncwarp();
            const uint32_t sw = (uint32_t)((lane >> 1) & 3);
#pragma unroll
            for (int p = 0; p < 4; ++p)
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(slot + (uint32_t)lane * 64u + (((uint32_t)p ^ sw) << 4)),
                             "f"(v[16 * h + 4 * p]), "f"(v[16 * h + 4 * p + 1]), "f"(v[16 * h + 4 * p + 2]), "f"(v[16 * h + 4 * p + 3]) : "memory");
            __syncwarp();
            const int c = lane & 15, j = 16 * h + c;
            const bool on = j >= jlo && j < jhi;
#pragma unroll 4
            for (int it = 0; it < 16; ++it) {
                const int r = it * 2 + (lane >> 4);
                float x;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x)
                             : "r"(slot + (uint32_t)r * 64u + ((((uint32_t)c >> 2) ^ (((uint32_t)r >> 1) & 3u)) << 4) + ((uint32_t)c & 3u) * 4u) : "memory");
                if (on && r < rows_left) C[(row0 + r) * ldc + c0 + j] = x;
            }
        }
        __syncwarp();
        flip = 0u;
    }
    // The same for a 1 KB slot (k_tc_chain): four passes of 8 columns; every global store covers four rows x 32 B.
    __device__ __forceinline__ void store_f32_narrow8(float* C, int64_t ldc, int c0, const float v[32], int jlo, int jhi) const {
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            if (jhi <= 8 * h || jlo >= 8 * h + 8) continue;         // warp-uniform
            __syncwarp();
#pragma unroll
            for (int p = 0; p < 2; ++p)
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(slot + (uint32_t)lane * 32u + (uint32_t)p * 16u),
                             "f"(v[8 * h + 4 * p]), "f"(v[8 * h + 4 * p + 1]), "f"(v[8 * h + 4 * p + 2]), "f"(v[8 * h + 4 * p + 3]) : "memory");
            __syncwarp();
            const int c = lane & 7, j = 8 * h + c;
            const bool on = j >= jlo && j < jhi;
#pragma unroll 4
            for (int it = 0; it < 8; ++it) {
                const int r = it * 4 + (lane >> 3);
                float x;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(slot + (uint32_t)r * 32u + (uint32_t)c * 4u) : "memory");
                if (on && r < rows_left) C[(row0 + r) * ldc + c0 + j] = x;
            }
        }
        __syncwarp();
        flip = 0u;
    }
    static __device__ __noinline__ void warp_store_f32(uint32_t slot, int lane, int64_t row0, int rows_left, float* C, int64_t ldc,
                                                       int c0, const float* v, int jlo, int jhi, bool accum) {
        __syncwarp();
#pragma unroll 4
        for (int c = 0; c < 32; ++c)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(slot + (uint32_t)(lane * 128 + (((c ^ lane) & 31) << 2))), "f"(v[c]) : "memory");
        __syncwarp();
        const bool on = lane >= jlo && lane < jhi;
#pragma unroll 4
        for (int r = 0; r < rows_left; ++r) {
            float x;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(slot + (uint32_t)(r * 128 + (((lane ^ r) & 31) << 2))) : "memory");
            if (on) { float* p = C + (row0 + r) * ldc + c0 + lane; *p = accum ? *p + x : x; }
        }
        __syncwarp();
    }
};

// ----------------------------------------------------------------------------------------------------------
// shared-memory carve-up
// ----------------------------------------------------------------------------------------------------------
struct Barriers {
    uint64_t full[kMaxStages], empty[kMaxStages], conv[kMaxStages], bfull, tfull[2], tempty[2];
    uint32_t tmem_base, pad;
};

constexpr uint32_t kSlotBytes = 4096;   // per epilogue warp
constexpr uint32_t kColVecBytes = 1024;

// Epilogue concept:
//   static constexpr int kPre                                   number of bf16 operands the epilogue reads back (0..2)
//   void prefetch(const WarpIO& io, int n0, uint4* q) const     issue their global loads into q[4 * kPre] (no waiting);
//                                                               called for all of the warp's chunks BEFORE the
//                                                               accumulator is awaited, so the loads overlap the MMAs
//   const float* colvec() const                                 per-column vector (bias, N entries) or nullptr; the kernel
//                                                               stages it in shared memory once, read with io.colvec()
//   void chunk(const WarpIO& io, int n0, float v[32], const uint4* q) const
//        v[j] = accumulator of row io.row(), column n0 + j (n0 % 32 == 0); the functor masks columns beyond its own N
//        and rows beyond io.M, and uses io.unstage / io.store / io.atomic_add for coalesced traffic.
template <class Epi> constexpr int pre_regs() { return Epi::kPre > 0 ? 4 * Epi::kPre : 1; }
constexpr int kMaxChunksPerWarp = 4;   // 256 accumulator columns / 32 / 2 warps per TMEM lane quadrant

// ==========================================================================================================
// C = epi(A W^T), weights resident
// ==========================================================================================================
// 384 threads = 3 warpgroups, launched with 168 registers per thread (65536 / 384).  Warpgroup 0 (TMA producer, MMA
// issuer: single-thread roles) shrinks to 40 registers with setmaxnreg and the two epilogue warpgroups grow to 232: room
// for a deeper rolling prefetch of the operands the epilogue reads back (ncu: 27 % of the EpiRev samples sat on the
// long-scoreboard stall of the staging store that consumes a prefetched vector -- not enough bytes in flight).
template <class Epi>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_tc_gemm(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW, int64_t M, int BN, int KB,
          int stages, int abytes, int a_fmt, int w_fmt, Epi epi) {    // abytes: cp.async landing buffer per epilogue warp (0: none)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sW = base;                                   // KB blocks of [BN rows x 128 B]
    const uint32_t w_block = (uint32_t)BN * 128u;
    const uint32_t sA = sW + (uint32_t)KB * w_block;            // stages x 16 KB
    const uint32_t sE = sA + (uint32_t)stages * kStageBytesA;   // kEpiWarps staging slots
    const uint32_t sV = sE + kEpiWarps * kSlotBytes;            // per-column vector (bias), 256 floats
    const uint32_t sX = sV + kColVecBytes;                      // cp.async landing buffers, kEpiWarps x abytes
    Barriers* bars = reinterpret_cast<Barriers*>(gen_base + (size_t)KB * w_block + (size_t)stages * kStageBytesA + kEpiWarps * kSlotBytes + kColVecBytes +
                                                 (size_t)kEpiWarps * abytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t num_tiles = (M + BM - 1) / BM;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapW);
        for (int s = 0; s < stages; ++s) { mbar_init(smem_u32(&bars->full[s]), 1); mbar_init(smem_u32(&bars->empty[s]), 1); }
        mbar_init(smem_u32(&bars->bfull), 1);
        for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(&bars->tfull[a]), 1); mbar_init(smem_u32(&bars->tempty[a]), kEpiWarps * 32); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(&bars->tmem_base), kTmemCols);
    if (threadIdx.x >= 128) {                                   // epilogue warps stage the column vector
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

    // (each setmaxnreg sits INSIDE its role's branch: after a merge point ptxas budgets for the smaller of the two counts)
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsLight));
    if (warp == 0) {
        if (lane == 0) {
            // ---- TMA producer: the layer's weights once, then the ring of A tiles
            mbar_expect_tx(smem_u32(&bars->bfull), (uint32_t)KB * w_block);
            for (int kb = 0; kb < KB; ++kb) tma_load_2d(sW + kb * w_block, &mapW, smem_u32(&bars->bfull), kb * BK, 0);
            int s = 0; uint32_t ph = 0;
            for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(smem_u32(&bars->empty[s]), ph ^ 1u);
                    mbar_expect_tx(smem_u32(&bars->full[s]), kStageBytesA);
                    tma_load_2d(sA + s * kStageBytesA, &mapA, smem_u32(&bars->full[s]), kb * BK, (int)(tile * BM));
                    if (++s == stages) { s = 0; ph ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // ---- MMA issuer
            const uint32_t idesc = instr_desc(BM, BN, 0, 0, a_fmt, w_fmt);
            mbar_wait(smem_u32(&bars->bfull), 0);
            int s = 0; uint32_t ph = 0; uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const uint32_t a = it & 1u, aph = (it >> 1) & 1u;
                mbar_wait(smem_u32(&bars->tempty[a]), aph ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + a * 256u;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(smem_u32(&bars->full[s]), ph);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t da = smem_desc(sA + s * kStageBytesA + k * (UMMA_K * 2), 16, 1024);
                        const uint64_t db = smem_desc(sW + kb * w_block + k * (UMMA_K * 2), 16, 1024);
                        umma_bf16(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(smem_u32(&bars->empty[s]));
                    if (++s == stages) { s = 0; ph ^= 1u; }
                }
                umma_commit(smem_u32(&bars->tfull[a]));
            }
        }
        __syncwarp();
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpi));
        // ---- epilogue: TMEM lane quadrant = warp % 4; the two warps of a quadrant interleave 32-column chunks
        const int q = warp & 3, half = (warp - 4) >> 2;
        const int chunks = (BN + 31) / 32;
        uint32_t it = 0;
        constexpr int kDepth = 2;                           // ring slots per warp: slot 0 in registers, slot 1 via cp.async
        uint4 pre[kDepth][pre_regs<Epi>()];
        WarpIO io{sE + (uint32_t)(warp - 4) * kSlotBytes, lane, (int64_t)blockIdx.x * BM + q * 32, M, sV};
        io.init();
        io.abuf = abytes > 0 ? sX + (uint32_t)(warp - 4) * (uint32_t)abytes : 0u;
        const bool use_async = abytes > 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const uint32_t a = it & 1u, aph = (it >> 1) & 1u;
            io.retile(tile * BM + q * 32);
            const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + a * 256u;
            if (Epi::kPre == 0) {
                mbar_wait(smem_u32(&bars->tfull[a]), aph);
                tc_fence_after();
                // accumulator chunks double buffered in registers: the TMEM load of chunk i+1 overlaps the math of chunk i
                uint32_t r[2][32];
                if (half < chunks) tmem_ld32_issue(tacc + (uint32_t)half * 32u, r[0]);
#pragma unroll
                for (int i = 0; i < kMaxChunksPerWarp; ++i) {
                    const int c = half + 2 * i;
                    if (c < chunks) {
                        tmem_ld32_wait(r[i & 1]);
                        if (c + 2 < chunks && i + 1 < kMaxChunksPerWarp) tmem_ld32_issue(tacc + (uint32_t)(c + 2) * 32u, r[(i + 1) & 1]);
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[i & 1][j]);
                        epi.chunk(io, c * 32, v, nullptr);
                    }
                }
            } else {
                // Rolling prefetch of the epilogue's bf16 operands, kDepth chunks ahead (across tile boundaries): the
                // registers of a consumed chunk are refilled at once with the loads of chunk + kDepth, so global latency
                // is covered by the math / stores of the chunks in between and never by an idle warp.
                if (it == 0) {
#pragma unroll
                    for (int i = 0; i < kDepth; ++i)
                        if (half + 2 * i < chunks) {
                            io.amode = (i & 1) && use_async; io.qbase = pre[i];
                            epi.prefetch(io, (half + 2 * i) * 32, pre[i]);
                            io.amode = 0;
                        }
                }
                const bool has_next = tile + gridDim.x < num_tiles;
                mbar_wait(smem_u32(&bars->tfull[a]), aph);
                tc_fence_after();
                // accumulator chunks double buffered in registers like above (the 232-register budget pays for it) -- except
                // under an epilogue with three read-back operands, whose prefetch ring alone holds 96 registers
                constexpr bool kAcc2 = Epi::kPre < 3;
                uint32_t r[kAcc2 ? 2 : 1][32];
                if (kAcc2 && half < chunks) tmem_ld32_issue(tacc + (uint32_t)half * 32u, r[0]);
#pragma unroll 1
                for (int ii = 0; ii < kMaxChunksPerWarp / kDepth; ++ii) {
#pragma unroll
                    for (int k = 0; k < kDepth; ++k) {
                        const int c = half + 2 * (ii * kDepth + k);
                        if (c < chunks) {
                            constexpr int kb0 = kAcc2 ? 1 : 0;
                            if (!kAcc2) tmem_ld32_issue(tacc + (uint32_t)c * 32u, r[0]);
                            tmem_ld32_wait(r[k & kb0]);
                            if (kAcc2 && c + 2 < chunks) tmem_ld32_issue(tacc + (uint32_t)(c + 2) * 32u, r[(k + 1) & kb0]);
                            float v[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[k & kb0][j]);
                            io.amode = (k & 1) && use_async; io.qbase = pre[k];
                            epi.chunk(io, c * 32, v, pre[k]);
                        }
                        io.amode = (k & 1) && use_async; io.qbase = pre[k];
                        if (ii + 1 < kMaxChunksPerWarp / kDepth) {             // refill: chunk + kDepth of this tile ...
                            if (c + 2 * kDepth < chunks) epi.prefetch(io, (c + 2 * kDepth) * 32, pre[k]);
                        } else if (has_next && half + 2 * k < chunks) {         // ... or chunk k of the next tile
                            io.pf_row_shift = (int64_t)gridDim.x * BM;
                            epi.prefetch(io, (half + 2 * k) * 32, pre[k]);
                            io.pf_row_shift = 0;
                        }
                        io.amode = 0;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&bars->tempty[a]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

// ==========================================================================================================
// dW += X^T Y  (both operands MN-major), split over the rows
// ==========================================================================================================
// grid (i_tiles, j_tiles, splits).  X tile: BI = 64 nbx columns of X (nbx = 2 or 4 boxes of [64 rows(k) x 64 cols]:
// one or two 128-row accumulators); Y tile: BJ columns (multiple of 64, <= 256) = BJ/64 boxes.  One k-block = 64 rows.
// The kernel is bound by shared-memory bandwidth (per stage: TMA writes + MMA operand reads + the format conversion
// below), so the 256-column X tile matters: the Y boxes are written / converted once for two accumulators' worth of
// MMAs (per 128x256 unit of work 112 KB instead of 160 KB of shared-memory traffic in the mixed-format products).
// Output tiles of one weight-gradient launch and the CTAs (row splits) each of them gets: a remainder tile along Y (the 64
// columns a 320-wide operand has beyond 256: the 289-wide colour input layer) loads and multiplies only its own boxes and
// receives CTAs in proportion to its shared-memory traffic instead of half of the grid.
struct WgTiles { int n; int first[5]; int i0[4], j0[4], nbx[4], bj[4]; long long rps[4]; };

template <class Epi>
__global__ void __launch_bounds__(kWgThreads, 1)
k_tc_wgrad(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapY, int64_t Mrows, int nbx_, int BJ_,
           const __grid_constant__ WgTiles tiles, int stages, int x_fmt, int y_fmt, Epi epi, float* __restrict__ colsum, int colsum_n,
           int colsum_perm, int colsum_shift, const __grid_constant__ CUtensorMap mapX2, const __grid_constant__ CUtensorMap mapY2,
           int x_fmt2, int y_fmt2, int phases) {
    // phases == 2: dW += X^T Y + X2^T Y2 in ONE launch (the two weight-gradient products of a layer, pbar^T h and a^T t:
    // same tile geometry, one prologue, one flush of the accumulator); the second product's k-blocks follow the first's
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
    constexpr uint32_t kBox = kWgRows * 128;                     // kWgRows k-rows x 128 B
    int tile = 0;
    while (tile + 1 < tiles.n && (int)blockIdx.x >= tiles.first[tile + 1]) ++tile;
    const int BJ = tiles.bj[tile];
    const uint32_t nbx = (uint32_t)tiles.nbx[tile], nby = (uint32_t)BJ / 64u;
    const uint32_t tmem_cols = nbx > 2 ? 512u : 256u;
    const uint32_t stage_bytes = ((uint32_t)nbx_ + (uint32_t)BJ_ / 64u) * kBox;   // stride of the ring: the launch's largest tile
    const uint32_t stage_tx = (nbx + nby) * kBox;                                  // bytes this CTA's tile loads per stage
    const uint32_t sE = base;   // the flush's staging slots reuse the stage ring (idle by then; host: ring >= 16 slots)
    Barriers* bars = reinterpret_cast<Barriers*>(gen_base + (size_t)stages * stage_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i0 = tiles.i0[tile], j0 = tiles.j0[tile];
    const int64_t rows_per_split = tiles.rps[tile];
    const int64_t r0 = (int64_t)((int)blockIdx.x - tiles.first[tile]) * rows_per_split;
    const int64_t r1 = (r0 + rows_per_split < Mrows) ? r0 + rows_per_split : Mrows;
    const bool do_colsum = colsum != nullptr && j0 == 0;   // out[i] += sum_m X[m, i]: bias gradients for free
    // mixed formats (pbar^T h, a^T t: one adjoint-like bf16 operand, one forward-like fp16 operand): the epilogue
    // warps, idle during the main loop, round the fp16 boxes of every stage to bf16 in place before the MMAs read them.
    // The kernel is bound by shared-memory bandwidth (TMA write + MMA read = 96 KB per stage; the conversion adds a
    // read and a write of the fp16 boxes: 128 / 160 KB, measured 1.7x per launch).  Tried and rejected: the epilogue
    // warps loading the fp16 operand with LDG two stages ahead and writing the swizzled boxes themselves (96 KB of
    // shared-memory traffic again, but 64 KB of loads in flight per SM is more than the LSU path sustains: 2.2x).
    // Also tried: the instruction descriptor has separate format fields for A and B, but a kind::f16 MMA with one fp16
    // and one bf16 operand traps on sm_100a ("an illegal instruction was encountered"), so the conversion stays.
    const bool do_conv = x_fmt != y_fmt;
    const int nkb = r1 > r0 ? (int)((r1 - r0 + kWgRows - 1) / kWgRows) : 0;   // the last block of a split may run past r1: the host
                                                                 // makes rows_per_split a multiple of 64, so only the
                                                                 // global tail is ragged and TMA zero-fills it
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapX);
        tma_prefetch_desc(&mapY);
        // a stage is released by the MMA commit and, when the column sums of X ride along, by the 8 reducing warps
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(&bars->full[s]), 1);
            mbar_init(smem_u32(&bars->empty[s]), do_colsum ? 1 + kWgEpiWarps : 1);
            mbar_init(smem_u32(&bars->conv[s]), kWgEpiWarps);
        }
        mbar_init(smem_u32(&bars->tfull[0]), 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(&bars->tmem_base), tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            if (phases > 1) { tma_prefetch_desc(&mapX2); tma_prefetch_desc(&mapY2); }
            for (int kk = 0; kk < nkb * phases; ++kk) {
                const int kb = kk >= nkb ? kk - nkb : kk;
                mbar_wait(smem_u32(&bars->empty[s]), ph ^ 1u);
                mbar_expect_tx(smem_u32(&bars->full[s]), stage_tx);
                const uint32_t st = base + s * stage_bytes;
                const int row = (int)(r0 + (int64_t)kb * kWgRows);
                const CUtensorMap* mx = kk >= nkb ? &mapX2 : &mapX;
                const CUtensorMap* my = kk >= nkb ? &mapY2 : &mapY;
                for (uint32_t b = 0; b < nbx; ++b) tma_load_2d(st + b * kBox, mx, smem_u32(&bars->full[s]), i0 + (int)b * 64, row);
                for (uint32_t b = 0; b < nby; ++b) tma_load_2d(st + (nbx + b) * kBox, my, smem_u32(&bars->full[s]), j0 + (int)b * 64, row);
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0 && nkb > 0) {
            const int mma_fmt = do_conv ? (int)kBF16 : x_fmt;
            const uint32_t idesc = instr_desc(128, BJ, 1, 1, mma_fmt, mma_fmt);
            int s = 0; uint32_t ph = 0;
            for (int kb = 0; kb < nkb * phases; ++kb) {
                mbar_wait(smem_u32(do_conv ? &bars->conv[s] : &bars->full[s]), ph);
                tc_fence_after();
                const uint32_t st = base + s * stage_bytes;
#pragma unroll
                for (int k = 0; k < kWgRows / UMMA_K; ++k) {
                    // MN-major SW128: 64-element groups along M/N are kBox apart (LBO), 8-row k groups 1024 B apart (SBO)
                    const uint64_t db = smem_desc(st + nbx * kBox + k * (UMMA_K * 128), kBox, 1024);
                    for (uint32_t h = 0; h < nbx / 2; ++h) {          // one 128-row accumulator per pair of X boxes
                        const uint64_t da = smem_desc(st + h * 2 * kBox + k * (UMMA_K * 128), kBox, 1024);
                        umma_bf16(tmem_base + h * 256u, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                }
                umma_commit(smem_u32(&bars->empty[s]));
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
            umma_commit(smem_u32(&bars->tfull[0]));
        }
        __syncwarp();
    } else if (nkb > 0) {
        if (do_colsum || do_conv) {
            // While the MMAs run these 8 warps are idle.  Per stage they (1) round the fp16 operand's boxes to bf16 in
            // place (element-wise, so the swizzle does not matter; 16-byte vectors, thread t takes vectors t, t + 256,
            // ...), make the writes visible to the tensor core (async proxy) and release the stage to the MMA warp;
            // (2) read the X boxes out of shared memory (the operand is there anyway) and accumulate its column sums.
            // Thread t: column pair t % (32 nbx) of the X tile, k-rows rpg (t / (32 nbx)) .. + rpg - 1 of the 64-row block
            // (rpg = 16 for the 128-column tile, 32 for the 256-column tile); a warp reads whole 128-byte swizzle rows.
            const int t = (int)threadIdx.x - 64;
            const int pairs = (int)nbx * 32, rpg = kWgRows / (kWgEpiWarps * 32 / pairs);
            const int pi = t % pairs, g = t / pairs;
            const int c = (pi & 31) * 2;
            const uint32_t box_off = (uint32_t)(pi >> 5) * kBox;
            const bool x_bf16 = do_conv || x_fmt == kBF16;          // format of X once the stage is released
            float s0 = 0.f, s1 = 0.f;
            int s = 0; uint32_t ph = 0;
            for (int kk = 0; kk < nkb * phases; ++kk) {
                const int xf = kk >= nkb ? x_fmt2 : x_fmt;            // the fp16 operand of this k-block's product
                const uint32_t conv_off = xf == kF16 ? 0u : nbx * kBox;
                const int conv_vecs = (int)((xf == kF16 ? nbx : nby) * (kBox / 16u));
                if (lane == 0) mbar_wait(smem_u32(&bars->full[s]), ph);   // one poller per warp
                __syncwarp();
                if (do_conv) {
                    const uint32_t cb = base + s * stage_bytes + conv_off;
                    // all of a thread's vectors are loaded before the first is converted: one shared-memory latency per
                    // stage instead of one per vector (ncu source page, one vector per trip: 32 % of the kernel's samples
                    // on the first conversion after the LDS.128, short scoreboard)
                    constexpr int kConvU = (kWgRows * 32 + kWgEpiWarps * 32 - 1) / (kWgEpiWarps * 32);   // 4 boxes x 8 kWgRows vectors / 512 threads
                    uint32_t w[kConvU][4];
#pragma unroll
                    for (int u = 0; u < kConvU; ++u) {
                        const int i = t + u * kWgEpiWarps * 32;
                        if (i < conv_vecs)
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[u][0]), "=r"(w[u][1]), "=r"(w[u][2]), "=r"(w[u][3]) : "r"(cb + (uint32_t)i * 16u) : "memory");
                    }
#pragma unroll
                    for (int u = 0; u < kConvU; ++u) {
                        const int i = t + u * kWgEpiWarps * 32;
                        if (i < conv_vecs) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) w[u][j] = WarpIO::pack2<kBF16>(WarpIO::lo_of<kF16>(w[u][j]), WarpIO::hi_of<kF16>(w[u][j]));
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cb + (uint32_t)i * 16u), "r"(w[u][0]), "r"(w[u][1]), "r"(w[u][2]), "r"(w[u][3]) : "memory");
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bars->conv[s]));
                    // X converted by OTHER warps is only read below when X is the fp16 operand, and column sums are
                    // never requested for that call (the sweeps sum adjoint-like, i.e. bf16, matrices only)
                }
                if (do_colsum) {
                    const uint32_t st = base + s * stage_bytes + box_off;
                    if (kk < nkb) {                                 // (the second product's X has no bias attached)
#pragma unroll 16
                    for (int kr = 0; kr < rpg; ++kr) {
                        const int k = g * rpg + kr;
                        uint32_t w;
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(st + (uint32_t)(k * 128 + ((((c >> 3) ^ (k & 7)) << 4) | ((c & 7) * 2)))) : "memory");
                        s0 += x_bf16 ? WarpIO::lo_of<kBF16>(w) : WarpIO::lo_of<kF16>(w);
                        s1 += x_bf16 ? WarpIO::hi_of<kBF16>(w) : WarpIO::hi_of<kF16>(w);
                    }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&bars->empty[s]));
                }
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
            if (do_colsum) {
                const int col = i0 + (pi >> 5) * 64 + c;
                // colsum_perm > 0: X's columns are a layer's rows with the first colsum_shift moved to the end
                if (col < colsum_n) atomicAdd(colsum + (colsum_perm > 0 ? (col + colsum_shift) % colsum_perm : col), s0);
                if (col + 1 < colsum_n) atomicAdd(colsum + (colsum_perm > 0 ? (col + 1 + colsum_shift) % colsum_perm : col + 1), s1);
            }
        }
        const int q = warp & 3, grp = (warp - 2) >> 2;
        const int chunks = (BJ + 31) / 32;
        mbar_wait(smem_u32(&bars->tfull[0]), 0);
        tc_fence_after();
        // every MMA has read its stage; the barrier also ends the other warps' bias sums out of the last stage
        asm volatile("bar.sync 1, %0;" ::"n"(kWgEpiWarps * 32) : "memory");
        WarpIO io{sE + (uint32_t)(warp - 2) * kSlotBytes, lane, (int64_t)i0 + q * 32, (int64_t)1 << 40, 0u};
        io.init();
        for (uint32_t h = 0; h < nbx / 2; ++h) {
            io.retile((int64_t)i0 + h * 128 + q * 32);
            for (int c = grp; c < chunks; c += kWgEpiWarps / 4) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + h * 256u + (uint32_t)c * 32u, v);
                epi.chunk(io, j0 + c * 32, v, nullptr);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

// ----------------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// bf16 row-major [rows, cols] with leading dimension ld (elements); box = [box_rows x 64 cols], 128-byte swizzle
inline int make_map(CUtensorMap* map, const void* p, int fmt, int64_t rows, int64_t cols, int64_t ld, int box_rows, const char* who,
                    int box_cols = 64) {   // 64: SWIZZLE_128B rows, 32: SWIZZLE_64B rows
    EncodeTiledFn fn = encode_fn();
    if (!fn) { msdf_set_error("%s: cuTensorMapEncodeTiled is unavailable", who); return MSDF_ERR_CUDA; }
    if ((((uintptr_t)p) & 15) != 0 || (ld % 8) != 0) { msdf_set_error("%s: TMA operand must be 16-byte aligned (ld %% 8 == 0)", who); return MSDF_ERR_ARG; }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(map, fmt == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(p), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { msdf_set_error("%s: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box_rows=%d", who, (int)r,
                                            (long long)rows, (long long)cols, (long long)ld, box_rows); return MSDF_ERR_CUDA; }
    return MSDF_OK;
}

inline int sm_count() {
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
    return n;
}

// C = epi(A W^T): A [M, Kp] (ld lda), W [BN, Kp] (ld ldw); Kp multiple of 64 (<= 320), BN multiple of 16 (<= 256)
template <class Epi>
int launch_gemm(const void* A, int a_fmt, int64_t lda, int64_t M, int Kp, const void* W, int w_fmt, int64_t ldw, int BN, const Epi& epi,
                cudaStream_t st, const char* what) {
    if (M <= 0) return MSDF_OK;
    if (a_fmt != w_fmt) { msdf_set_error("%s: both MMA operands must have the same 16-bit format", what); return MSDF_ERR_ARG; }
    if (Kp % 64 != 0 || Kp <= 0 || Kp > 640 || BN % 16 != 0 || BN < 16 || BN > 256) {
        msdf_set_error("%s: tensor-core GEMM needs K %% 64 == 0 (<= 640) and N %% 16 == 0 (<= 256); got K=%d N=%d", what, Kp, BN);
        return MSDF_ERR_ARG;
    }
    CUtensorMap mA, mW;
    int rc = make_map(&mA, A, a_fmt, M, Kp, lda, BM, what); if (rc) return rc;
    rc = make_map(&mW, W, w_fmt, BN, Kp, ldw, BN, what); if (rc) return rc;
    const int KB = Kp / 64;
    const size_t wbytes = (size_t)KB * BN * 128;
    size_t fixed = 1024 + sizeof(Barriers) + kEpiWarps * kSlotBytes + kColVecBytes;
    // cp.async landing buffers of the epilogue's second prefetch path (2 KB per read-back operand and warp), when the
    // weights leave room for them next to two A stages (the depth of the A ring does not matter: measured 2 = 3 = 4)
    int abytes = Epi::kPre * 2048;
    if (227 * 1024 < fixed + (size_t)kEpiWarps * abytes + wbytes + 2 * kStageBytesA) abytes = 0;
    fixed += (size_t)kEpiWarps * abytes;
    int stages = (int)((227 * 1024 - fixed - wbytes) / kStageBytesA);
    if (stages > 6) stages = 6;
    if (stages < 2) { msdf_set_error("%s: weights do not leave room for the A ring", what); return MSDF_ERR_ARG; }
    const size_t smem = fixed + wbytes + (size_t)stages * kStageBytesA;
    static bool attr_set = false;   // per instantiation
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_tc_gemm<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { msdf_set_error("%s: cannot opt in to 227 KB shared memory: %s", what, cudaGetErrorString(e)); return MSDF_ERR_CUDA; }
        attr_set = true;
    }
    const int64_t tiles = (M + BM - 1) / BM;
    const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    // algorithmic bytes: the A tile once, plus every bf16 operand the epilogue reads back and writes (epi.N real columns)
    const int prof = msdf_prof_begin(MSDF_PROF_TC_GEMM, 2.0 * (double)M * (double)BN * (double)Kp, st,
                                     (double)M * 2.0 * ((double)Kp + (double)epi.N * (double)(Epi::kPre + Epi::kStores)));
    k_tc_gemm<Epi><<<grid, kGemmThreads, smem, st>>>(mA, mW, M, BN, KB, stages, abytes, a_fmt, w_fmt, epi);
    msdf_prof_end(prof, st);
    MSDF_COUNT_LAUNCH();
    MSDF_CHECK_LAUNCH(what);
    return MSDF_OK;
}

// dW[i,j] (+)= sum_m X[m,i] Y[m,j]: X [M, Ci] (ld ldx), Y [M, Cj] (ld ldy), Ci / Cj = padded column counts (multiples
// of 64); the functor masks i / j beyond the real sizes and accumulates atomically.
template <class Epi>
int launch_wgrad(const void* X, int x_fmt, int64_t ldx, int Ci, const void* Y, int y_fmt, int64_t ldy, int Cj, int64_t M, const Epi& epi,
                 cudaStream_t st, const char* what, float* colsum = nullptr, int colsum_n = 0, int colsum_perm = 0, int colsum_shift = 1,
                 const void* X2 = nullptr, int x_fmt2 = 0, int64_t ldx2 = 0, const void* Y2 = nullptr, int y_fmt2 = 0, int64_t ldy2 = 0) {
    if (M <= 0) return MSDF_OK;
    if (X2 != nullptr && ((x_fmt != y_fmt) != (x_fmt2 != y_fmt2) || x_fmt == y_fmt)) {
        msdf_set_error("%s: a two-product weight gradient needs two mixed-format products", what);
        return MSDF_ERR_ARG;
    }
    if (Ci % 64 != 0 || Cj % 64 != 0 || Ci <= 0 || Cj <= 0) { msdf_set_error("%s: wgrad needs column counts %% 64 == 0", what); return MSDF_ERR_ARG; }
    if (colsum != nullptr && x_fmt != y_fmt && x_fmt == kF16) {
        msdf_set_error("%s: column sums of the fp16 operand of a mixed-format weight gradient are not supported", what);
        return MSDF_ERR_UNSUPPORTED;
    }
    CUtensorMap mX, mY, mX2, mY2;
    int rc = make_map(&mX, X, x_fmt, M, Ci, ldx, kWgRows, what); if (rc) return rc;
    rc = make_map(&mY, Y, y_fmt, M, Cj, ldy, kWgRows, what); if (rc) return rc;
    mX2 = mX; mY2 = mY;
    if (X2 != nullptr) {
        rc = make_map(&mX2, X2, x_fmt2, M, Ci, ldx2, kWgRows, what); if (rc) return rc;
        rc = make_map(&mY2, Y2, y_fmt2, M, Cj, ldy2, kWgRows, what); if (rc) return rc;
    }
    const int phases = X2 != nullptr ? 2 : 1;
    const int BJ = Cj < 256 ? Cj : 256;
    const int nbx = Ci > 128 ? 4 : 2;                       // X tile of 256 (two accumulators) or 128 columns
    const int it = (Ci + nbx * 64 - 1) / (nbx * 64), jt = (Cj + BJ - 1) / BJ;
    if (it * jt > 4) { msdf_set_error("%s: more than 4 output tiles", what); return MSDF_ERR_ARG; }
    WgTiles tiles{};
    double cost[4], total_cost = 0.0;
    for (int ti = 0; ti < it; ++ti)
        for (int tj = 0; tj < jt; ++tj) {
            const int t = tiles.n++;
            const int xb = (Ci - ti * nbx * 64) / 64, yc = Cj - tj * BJ;
            tiles.i0[t] = ti * nbx * 64; tiles.j0[t] = tj * BJ;
            tiles.nbx[t] = nbx;      // (a remainder X tile trimmed to 128 columns costs as much per stage as a full one: measured)
            (void)xb;
            tiles.bj[t] = yc < BJ ? yc : BJ;
            const int bx = tiles.nbx[t], by = tiles.bj[t] / 64;
            // shared-memory traffic of a stage in KB-ish units: TMA writes, MMA reads (B once per accumulator), the
            // conversion's read + write of the fp16 operand, the bias sums
            cost[t] = 8.0 * (bx + by) + (bx / 2) * (16.0 + 8.0 * by) + (x_fmt != y_fmt ? 16.0 * (x_fmt == kF16 ? bx : by) : 0.0) +
                      (colsum != nullptr && tj == 0 ? 8.0 * bx : 0.0);
            // measured (1 060 864 rows): a [256 x 64] tile's stage takes 0.45-0.5 of a [256 x 256] tile's, as modelled
            total_cost += cost[t];
        }
    int grid_ctas = 0;
    for (int t = 0; t < tiles.n; ++t) {
        int splits = (int)(sm_count() * cost[t] / total_cost + 0.5);
        if (splits < 1) splits = 1;
        int64_t rps = (M + splits - 1) / splits;
        rps = (rps + kWgRows - 1) / kWgRows * kWgRows;
        if (rps < 4 * kWgRows) rps = 4 * kWgRows;
        splits = (int)((M + rps - 1) / rps);
        tiles.rps[t] = rps; tiles.first[t] = grid_ctas;
        grid_ctas += splits;
    }
    tiles.first[tiles.n] = grid_ctas;
    const uint32_t stage_bytes = (nbx + BJ / 64) * kWgRows * 128;
    const size_t fixed = 1024 + sizeof(Barriers);
    int stages = (int)((227 * 1024 - fixed) / stage_bytes);
    if (stages > kMaxStages - 1) stages = kMaxStages - 1;
    if ((size_t)stages * stage_bytes < (size_t)kWgEpiWarps * kSlotBytes) { msdf_set_error("%s: stage ring smaller than the flush slots", what); return MSDF_ERR_ARG; }
    const size_t smem = fixed + (size_t)stages * stage_bytes;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_tc_wgrad<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { msdf_set_error("%s: cannot opt in to 227 KB shared memory: %s", what, cudaGetErrorString(e)); return MSDF_ERR_CUDA; }
        attr_set = true;
    }
    dim3 grid((unsigned)grid_ctas, 1, 1);
    const int prof = msdf_prof_begin(MSDF_PROF_TC_WGRAD, 2.0 * phases * (double)M * (double)Ci * (double)Cj, st, phases * (double)M * 2.0 * (double)(Ci + Cj));
    k_tc_wgrad<Epi><<<grid, kWgThreads, smem, st>>>(mX, mY, M, nbx, BJ, tiles, stages, x_fmt, y_fmt, epi, colsum, colsum_n, colsum_perm, colsum_shift,
                                                  mX2, mY2, x_fmt2, y_fmt2, phases);
    msdf_prof_end(prof, st);
    MSDF_COUNT_LAUNCH();
    MSDF_CHECK_LAUNCH(what);
    return MSDF_OK;
}

}  // namespace msdf_tc
