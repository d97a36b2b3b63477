// fp32 SIMT GEMM with pluggable epilogues -- the "fp32 mode" contraction engine of the MLP sweeps
// (mlp.cu).  The reference runs every nn.Linear as a cuBLAS fp32 SGEMM (no TF32 anywhere in its tree),
// so this path reproduces its arithmetic class: fp32 FFMA accumulation, results within 1e-4 relative.
//
//   C[m,n] = sum_k A(m,k) * B(k,n),   m < M, n < N, k < K,  handed to an epilogue functor.
//   Layout NT: A[M,K] row-major (lda), B(k,n) = W[n*ldb + k]          forward / tangent sweeps  (x W^T)
//   Layout NN: A[M,K] row-major (lda), B(k,n) = W[k*ldb + n]          reverse / backward sweeps (a W)
//   Layout TN: A(m,k) = X[k*lda + m],  B(k,n) = Y[k*ldb + n]          weight gradients (X^T Y), split over k
//
// Tile 128x128x16, 256 threads, 8x8 register micro-tile (2x2 groups of 4x4), double-buffered shared
// memory with register prefetch.  Arbitrary M, N, K (guards); 128-bit loads when the operand is 16-byte
// aligned with a leading dimension that is a multiple of 4.
#pragma once
#include "common.cuh"

namespace msdf_gemm {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4, NT_THREADS = 256;
enum Layout { kNT = 0, kNN = 1, kTN = 2 };

struct Operand {
    const float* p;
    int64_t ld;
    int vec;   // 1: 16-byte aligned base and ld % 4 == 0
};

static inline Operand make_operand(const float* p, int64_t ld) {
    Operand o{p, ld, (int)((((uintptr_t)p) & 15) == 0 && (ld & 3) == 0)};
    return o;
}

// tile loaders ---------------------------------------------------------------------------------------------
// K-contiguous source (row index r in [0,R), k contiguous): thread loads 2 x 4 consecutive k of rows t/4 (+64)
__device__ __forceinline__ void ldg_kcontig(const Operand& op, int64_t r0, int64_t R, int k0, int K, float4 v[2]) {
    const int t = threadIdx.x;
    const int kq = (t & 3) * 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int64_t r = r0 + (t >> 2) + 64 * i;
        const int k = k0 + kq;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < R) {
            const float* src = op.p + r * op.ld + k;
            if (op.vec && k + 3 < K) {
                x = *reinterpret_cast<const float4*>(src);
            } else {
                if (k < K) x.x = src[0];
                if (k + 1 < K) x.y = src[1];
                if (k + 2 < K) x.z = src[2];
                if (k + 3 < K) x.w = src[3];
            }
        }
        v[i] = x;
    }
}
__device__ __forceinline__ void sts_kcontig(float (*s)[BM + PAD], const float4 v[2]) {
    const int t = threadIdx.x;
    const int kq = (t & 3) * 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int r = (t >> 2) + 64 * i;
        s[kq + 0][r] = v[i].x; s[kq + 1][r] = v[i].y; s[kq + 2][r] = v[i].z; s[kq + 3][r] = v[i].w;
    }
}
// column-contiguous source (element (k, c) at p[k*ld + c], c in [0,Cn)): thread loads 2 x 4 consecutive c
__device__ __forceinline__ void ldg_ccontig(const Operand& op, int64_t c0, int64_t Cn, int64_t k0, int64_t K, float4 v[2]) {
    const int t = threadIdx.x;
    const int cq = (t & 31) * 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int64_t k = k0 + (t >> 5) + 8 * i;
        const int64_t c = c0 + cq;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < K) {
            const float* src = op.p + k * op.ld + c;
            if (op.vec && c + 3 < Cn) {
                x = *reinterpret_cast<const float4*>(src);
            } else {
                if (c < Cn) x.x = src[0];
                if (c + 1 < Cn) x.y = src[1];
                if (c + 2 < Cn) x.z = src[2];
                if (c + 3 < Cn) x.w = src[3];
            }
        }
        v[i] = x;
    }
}
__device__ __forceinline__ void sts_ccontig(float (*s)[BM + PAD], const float4 v[2]) {
    const int t = threadIdx.x;
    const int cq = (t & 31) * 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) *reinterpret_cast<float4*>(&s[(t >> 5) + 8 * i][cq]) = v[i];
}

// Epilogue concept:  __device__ void operator()(int64_t m, int n, const float v[4], int nvalid) const
//   -- 4 consecutive columns n..n+3 of row m (n % 4 == 0), the first nvalid of them inside N.
template <int LAYOUT, class Epi>
__global__ void __launch_bounds__(NT_THREADS, 2)
k_gemm(Operand A, Operand B, int64_t M, int N, int64_t K, int64_t k_per_split, Epi epi) {
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];
    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int64_t kb = (int64_t)blockIdx.z * k_per_split;
    const int64_t ke = (kb + k_per_split < K) ? kb + k_per_split : K;
    if (kb >= ke) return;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float4 ra[2], rb[2];
    auto load = [&](int64_t k0) {
        if (LAYOUT == kTN) ldg_ccontig(A, m0, M, k0, ke, ra); else ldg_kcontig(A, m0, M, (int)k0, (int)ke, ra);
        if (LAYOUT == kNT) ldg_kcontig(B, n0, N, (int)k0, (int)ke, rb); else ldg_ccontig(B, n0, N, k0, ke, rb);
    };
    auto store = [&](int buf) {
        if (LAYOUT == kTN) sts_ccontig(As[buf], ra); else sts_kcontig(As[buf], ra);
        if (LAYOUT == kNT) sts_kcontig(Bs[buf], rb); else sts_ccontig(Bs[buf], rb);
    };
    load(kb);
    store(0);
    __syncthreads();
    int buf = 0;
    for (int64_t k0 = kb; k0 < ke; k0 += BK) {
        const bool more = k0 + BK < ke;
        if (more) load(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) {
            store(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
#pragma unroll
        for (int jg = 0; jg < 2; ++jg) {
            const int n = n0 + jg * 64 + tx * 4;
            const int nvalid = N - n;
            if (nvalid <= 0) continue;
            const float v[4] = {acc[i][jg * 4 + 0], acc[i][jg * 4 + 1], acc[i][jg * 4 + 2], acc[i][jg * 4 + 3]};
            epi(m, n, v, nvalid < 4 ? nvalid : 4);
        }
    }
}

// splits > 1 only makes sense with an accumulating (atomic) epilogue.
template <int LAYOUT, class Epi>
int launch(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int N, int64_t K, int splits,
           const Epi& epi, cudaStream_t st, const char* what) {
    if (M <= 0 || N <= 0 || K <= 0) return MSDF_OK;
    if (splits < 1) splits = 1;
    int64_t kps = (K + splits - 1) / splits;
    kps = (kps + BK - 1) / BK * BK;
    splits = (int)((K + kps - 1) / kps);
    dim3 grid((unsigned)msdf_div_up(M, BM), (unsigned)msdf_div_up(N, BN), (unsigned)splits);
    const int prof = msdf_prof_begin(MSDF_PROF_GEMM_F32, 2.0 * (double)M * (double)N * (double)K, st);
    k_gemm<LAYOUT, Epi><<<grid, NT_THREADS, 0, st>>>(make_operand(A, lda), make_operand(B, ldb), M, N, K, kps, epi);
    msdf_prof_end(prof, st);
    MSDF_COUNT_LAUNCH();
    MSDF_CHECK_LAUNCH(what);
    return MSDF_OK;
}

}  // namespace msdf_gemm
