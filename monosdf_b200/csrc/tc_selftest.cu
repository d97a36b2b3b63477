// Self-test of the tcgen05 GEMM engine (tc_gemm.cuh) against a naive fp32-accumulate kernel on the same 16-bit
// inputs (bf16 and fp16 operands in every combination the sweeps use).  Test infrastructure reachable through the C ABI (msdf_tc_selftest) so that the -m gpu tests can pin the
// UMMA descriptors / TMA layouts before the sweeps use them.
#include "tc_gemm.cuh"

namespace {
using bf16 = uint16_t;      // raw 16-bit storage; the format travels as msdf_tc::Fmt
using msdf_tc::kBF16;
using msdf_tc::kF16;

__device__ __forceinline__ float dec(uint16_t w, int f) { return f == kBF16 ? __uint_as_float((uint32_t)w << 16) : __half2float(__ushort_as_half(w)); }
__device__ __forceinline__ uint16_t enc(float v, int f) {
    return f == kBF16 ? __bfloat16_as_ushort(__float2bfloat16(v)) : __half_as_ushort(__float2half_rn(v));
}

__global__ void k_fill(bf16* p, int64_t n, uint32_t seed, float scale, int f) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t x = (uint32_t)i * 747796405u + seed * 2891336453u + 12345u;
    x ^= x >> 17; x *= 0xed5ad4bbu; x ^= x >> 11; x *= 0xac4c1b51u; x ^= x >> 15;
    p[i] = enc(((x & 0xffff) / 65536.0f - 0.5f) * scale, f);
}
// C[m,n] = sum_k A[m,k] W[n,k]
__global__ void k_ref_gemm(const bf16* A, int fa, int64_t lda, const bf16* W, int fb, int64_t ldw, int64_t M, int N, int K, float* C, int64_t ldc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * N) return;
    const int64_t m = i / N; const int n = (int)(i - m * N);
    float s = 0.f;
    for (int k = 0; k < K; ++k) s = fmaf(dec(A[m * lda + k], fa), dec(W[(int64_t)n * ldw + k], fb), s);
    C[m * ldc + n] = s;
}
// C[i,j] = sum_m X[m,i] Y[m,j]
__global__ void k_ref_wgrad(const bf16* X, int fa, int64_t ldx, const bf16* Y, int fb, int64_t ldy, int64_t M, int Ni, int Nj, float* C, int64_t ldc) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Ni * Nj) return;
    const int i = idx / Nj, j = idx - i * Nj;
    float s = 0.f;
    for (int64_t m = 0; m < M; ++m) {
        float xv = dec(X[m * ldx + i], fa), yv = dec(Y[m * ldy + j], fb);
        if (fa != fb) { if (fa == kF16) xv = dec(enc(xv, kBF16), kBF16); else yv = dec(enc(yv, kBF16), kBF16); }
        s = fmaf(xv, yv, s);
    }
    C[(int64_t)i * ldc + j] = s;
}
// expected result of the bf16-io epilogue: columns < N: bf16(ref + Cin); columns >= N: the sentinel already in Cb
__global__ void k_ref_bf16(const float* ref, const bf16* Cin, int fin, const bf16* Cb, int fout, int64_t M, int N, int ldc, float* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * ldc) return;
    const int n = (int)(i % ldc);
    out[i] = n < N ? dec(enc(ref[i] + dec(Cin[i], fin), fout), fout) : dec(Cb[i], fout);
}
__global__ void k_bf16_to_f32(const bf16* a, int f, float* b, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = dec(a[i], f);
}
// out[i] = sum_m X[m, i] (reference of the column sums that ride along in the weight-gradient kernel)
__global__ void k_ref_colsum(const bf16* X, int f, int64_t ldx, int64_t M, int N, float* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float s = 0.f;
    for (int64_t m = 0; m < M; ++m) s += dec(X[m * ldx + i], f);
    out[i] = s;
}
__global__ void k_maxerr(const float* a, const float* b, int64_t n, float* out) {   // out[0] = max |a-b|, out[1] = max |b|
    float e = 0.f, r = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float d = fabsf(a[i] - b[i]);
        e = (d > e || d != d) ? d : e;
        r = fmaxf(r, fabsf(b[i]));
    }
    atomicMax(reinterpret_cast<int*>(out), __float_as_int(e != e ? INFINITY : e));
    atomicMax(reinterpret_cast<int*>(out + 1), __float_as_int(r));
}

struct EpiStore {     // C fp32: no coalescing helper for fp32 row stores; plain per-row writes are fine for a test
    float* C; int64_t ldc; int N;
    static constexpr int kPre = 0; static constexpr int kStores = 1;
    __device__ __forceinline__ void prefetch(const msdf_tc::WarpIO&, int, uint4*) const {}
    __device__ __forceinline__ const float* colvec() const { return nullptr; }
    __device__ __forceinline__ void chunk(const msdf_tc::WarpIO& io, int n0, float v[32], const uint4* q) const {
        if (!io.valid()) return;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (n0 + j < N) C[io.row() * ldc + n0 + j] = v[j];
    }
};
template <int FIN, int FOUT>
struct EpiStoreBf16 {  // exercises WarpIO::load / store: C = fmt_out(acc + fmt_in(Cin))
    bf16* C; const bf16* Cin; int64_t ldc; int N;
    static constexpr int kPre = 1; static constexpr int kStores = 1;
    __device__ __forceinline__ void prefetch(const msdf_tc::WarpIO& io, int n0, uint4* q) const { if (n0 < N) io.prefetch(Cin, ldc, n0, q); }
    __device__ __forceinline__ const float* colvec() const { return nullptr; }
    __device__ __forceinline__ void chunk(const msdf_tc::WarpIO& io, int n0, float v[32], const uint4* q) const {
        const int nv = N - n0;
        if (nv <= 0) return;
        float a[32];
        io.unstage<FIN>(q, a);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += a[j];
        io.store<FOUT>(C, ldc, n0, v, nv < 32 ? nv : 32);
    }
};
struct EpiAtomicAdd {
    float* C; int64_t ldc; int Ni, Nj;
    static constexpr int kPre = 0; static constexpr int kStores = 1;
    __device__ __forceinline__ void prefetch(const msdf_tc::WarpIO&, int, uint4*) const {}
    __device__ __forceinline__ const float* colvec() const { return nullptr; }
    __device__ __forceinline__ void chunk(const msdf_tc::WarpIO& io, int n0, float v[32], const uint4* q) const {
        const int nv = Nj - n0;
        if (nv <= 0) return;
        float* Cp = C; const int64_t ld = ldc; const int nn = nv;
        io.atomic_add(v, (int64_t)Ni, [=](int64_t r, int c) -> float* { return c < nn ? Cp + r * ld + n0 + c : nullptr; });
    }
};

// kind 0: gemm (N x K weights), 1: wgrad (Ni = N, Nj = K), 2: GEMM with the 16-bit load / store epilogue;
// fa / fb: formats of the two MMA operands (kind 2: also of the epilogue's read-back operand / its output)
struct Case { int kind; int64_t M; int N, K; int Np, Kp; int fa, fb; };
const Case kCases[] = {
    {0, 1000, 256, 256, 256, 256, kBF16, kBF16},
    {0, 128 * 150 + 77, 217, 39, 224, 64, kF16, kF16},       // forward sweep: fp16 activations x fp16 weights
    {0, 40000, 64, 289, 64, 320, kBF16, kBF16},              // tangent / backward sweeps: bf16 adjoints x bf16 weights
    {0, 128 * 600, 256, 256, 256, 256, kF16, kF16},
    {1, 5000, 256, 256, 256, 256, kBF16, kBF16},
    {1, 70001, 257, 39, 320, 64, kBF16, kF16},               // dW = pbar^T h
    {1, 300, 64, 256, 64, 256, kF16, kBF16},                 // dW = a^T t
    {1, 262144, 256, 256, 256, 256, kBF16, kF16},
    {2, 128 * 40 + 19, 217, 256, 224, 256, kBF16, kBF16},    // ragged N, ragged M
    {2, 70000, 256, 256, 256, 256, kF16, kF16},
    {2, 128 * 33 + 5, 256, 256, 256, 256, kF16, kBF16},      // tangent epilogue: reads fp16, writes bf16
};
}  // namespace

extern "C" int msdf_tc_selftest_count(void) { return (int)(sizeof(kCases) / sizeof(kCases[0])); }

extern "C" int msdf_tc_selftest(int variant, float* result_host, void* stream) {
    const int ncases = (int)(sizeof(kCases) / sizeof(kCases[0]));
    MSDF_CHECK_ARG(variant >= 0 && variant < ncases && result_host, "msdf_tc_selftest: variant must be in [0,%d)", ncases);
    const Case c = kCases[variant];
    cudaStream_t st = (cudaStream_t)stream;
    bf16 *A = nullptr, *B = nullptr; float *C = nullptr, *R = nullptr, *out = nullptr;
    int rc = MSDF_OK;
    if (c.kind == 2) {
        const int64_t lda = c.Kp, ldw = c.Kp, ldc = 256;
        bf16 *Cb = nullptr, *Ci = nullptr;
        MSDF_CUDA_CALL(cudaMalloc(&A, c.M * lda * 2)); MSDF_CUDA_CALL(cudaMalloc(&B, (int64_t)c.Np * ldw * 2));
        MSDF_CUDA_CALL(cudaMalloc(&Cb, c.M * ldc * 2)); MSDF_CUDA_CALL(cudaMalloc(&Ci, c.M * ldc * 2));
        MSDF_CUDA_CALL(cudaMalloc(&C, c.M * ldc * 4)); MSDF_CUDA_CALL(cudaMalloc(&R, c.M * ldc * 4)); MSDF_CUDA_CALL(cudaMalloc(&out, 8));
        // MMA operands: A and the weights in format fa; the epilogue reads Ci in fa and writes Cb in fb
        k_fill<<<(unsigned)msdf_div_up(c.M * lda, 256), 256, 0, st>>>(A, c.M * lda, 1, 2.0f, c.fa);
        k_fill<<<(unsigned)msdf_div_up((int64_t)c.Np * ldw, 256), 256, 0, st>>>(B, (int64_t)c.Np * ldw, 2, 1.0f, c.fa);
        k_fill<<<(unsigned)msdf_div_up(c.M * ldc, 256), 256, 0, st>>>(Ci, c.M * ldc, 5, 4.0f, c.fa);
        k_fill<<<(unsigned)msdf_div_up(c.M * ldc, 256), 256, 0, st>>>(Cb, c.M * ldc, 6, 1.0f, c.fb);   // sentinel beyond N must survive
        MSDF_CUDA_CALL(cudaMemsetAsync(out, 0, 8, st));
        k_ref_gemm<<<(unsigned)msdf_div_up(c.M * c.N, 256), 256, 0, st>>>(A, c.fa, lda, B, c.fa, ldw, c.M, c.N, c.Kp, R, ldc);
        k_ref_bf16<<<(unsigned)msdf_div_up(c.M * ldc, 256), 256, 0, st>>>(R, Ci, c.fa, Cb, c.fb, c.M, c.N, (int)ldc, R);
        if (c.fa == kBF16 && c.fb == kBF16) {
            EpiStoreBf16<kBF16, kBF16> e{Cb, Ci, ldc, c.N};
            rc = msdf_tc::launch_gemm(A, c.fa, lda, c.M, c.Kp, B, c.fa, ldw, c.Np, e, st, "msdf_tc_selftest(gemm 16-bit io)");
        } else if (c.fa == kF16 && c.fb == kF16) {
            EpiStoreBf16<kF16, kF16> e{Cb, Ci, ldc, c.N};
            rc = msdf_tc::launch_gemm(A, c.fa, lda, c.M, c.Kp, B, c.fa, ldw, c.Np, e, st, "msdf_tc_selftest(gemm 16-bit io)");
        } else {
            EpiStoreBf16<kF16, kBF16> e{Cb, Ci, ldc, c.N};
            rc = msdf_tc::launch_gemm(A, c.fa, lda, c.M, c.Kp, B, c.fa, ldw, c.Np, e, st, "msdf_tc_selftest(gemm 16-bit io)");
        }
        if (!rc) {
            k_bf16_to_f32<<<(unsigned)msdf_div_up(c.M * ldc, 256), 256, 0, st>>>(Cb, c.fb, C, c.M * ldc);
            k_maxerr<<<256, 256, 0, st>>>(C, R, c.M * ldc, out);
        }
        cudaError_t e3 = cudaStreamSynchronize(st);
        if (!rc && e3 != cudaSuccess) { msdf_set_error("msdf_tc_selftest: kernel failed: %s", cudaGetErrorString(e3)); rc = MSDF_ERR_CUDA; }
        if (!rc) cudaMemcpy(result_host, out, 8, cudaMemcpyDeviceToHost);
        cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(R); cudaFree(out); cudaFree(Cb); cudaFree(Ci);
        return rc;
    }
    if (c.kind == 0) {
        const int64_t lda = c.Kp, ldw = c.Kp, ldc = c.Np;
        MSDF_CUDA_CALL(cudaMalloc(&A, c.M * lda * 2)); MSDF_CUDA_CALL(cudaMalloc(&B, (int64_t)c.Np * ldw * 2));
        MSDF_CUDA_CALL(cudaMalloc(&C, c.M * ldc * 4)); MSDF_CUDA_CALL(cudaMalloc(&R, c.M * ldc * 4)); MSDF_CUDA_CALL(cudaMalloc(&out, 8));
        k_fill<<<(unsigned)msdf_div_up(c.M * lda, 256), 256, 0, st>>>(A, c.M * lda, 1, 2.0f, c.fa);
        k_fill<<<(unsigned)msdf_div_up((int64_t)c.Np * ldw, 256), 256, 0, st>>>(B, (int64_t)c.Np * ldw, 2, 1.0f, c.fb);
        MSDF_CUDA_CALL(cudaMemsetAsync(C, 0, c.M * ldc * 4, st)); MSDF_CUDA_CALL(cudaMemsetAsync(R, 0, c.M * ldc * 4, st));
        MSDF_CUDA_CALL(cudaMemsetAsync(out, 0, 8, st));
        // the padded K columns of A hold random (finite) data: zero the padded K columns of the weights instead
        k_ref_gemm<<<(unsigned)msdf_div_up(c.M * c.N, 256), 256, 0, st>>>(A, c.fa, lda, B, c.fb, ldw, c.M, c.N, c.Kp, R, ldc);
        EpiStore e{C, ldc, c.N};
        rc = msdf_tc::launch_gemm(A, c.fa, lda, c.M, c.Kp, B, c.fb, ldw, c.Np, e, st, "msdf_tc_selftest(gemm)");
        if (!rc) k_maxerr<<<256, 256, 0, st>>>(C, R, c.M * ldc, out);
    } else {
        const int64_t ldx = c.Np, ldy = c.Kp, ldc = c.K;
        MSDF_CUDA_CALL(cudaMalloc(&A, c.M * ldx * 2)); MSDF_CUDA_CALL(cudaMalloc(&B, c.M * ldy * 2));
        MSDF_CUDA_CALL(cudaMalloc(&C, (int64_t)c.N * ldc * 4)); MSDF_CUDA_CALL(cudaMalloc(&R, (int64_t)c.N * ldc * 4)); MSDF_CUDA_CALL(cudaMalloc(&out, 8));
        k_fill<<<(unsigned)msdf_div_up(c.M * ldx, 256), 256, 0, st>>>(A, c.M * ldx, 3, 1.0f, c.fa);
        k_fill<<<(unsigned)msdf_div_up(c.M * ldy, 256), 256, 0, st>>>(B, c.M * ldy, 4, 1.0f, c.fb);
        MSDF_CUDA_CALL(cudaMemsetAsync(C, 0, (int64_t)c.N * ldc * 4, st)); MSDF_CUDA_CALL(cudaMemsetAsync(out, 0, 8, st));
        k_ref_wgrad<<<(unsigned)msdf_div_up((int64_t)c.N * c.K, 128), 128, 0, st>>>(A, c.fa, ldx, B, c.fb, ldy, c.M, c.N, c.K, R, ldc);
        EpiAtomicAdd e{C, ldc, c.N, c.K};
        const bool with_cs = !(c.fa != c.fb && c.fa == kF16);   // no column sums of the fp16 operand of a mixed call
        float* cs = nullptr;   // [0, N): fused column sums, [N, 2N): reference
        MSDF_CUDA_CALL(cudaMalloc(&cs, 2 * c.N * 4));
        MSDF_CUDA_CALL(cudaMemsetAsync(cs, 0, 2 * c.N * 4, st));
        k_ref_colsum<<<(unsigned)msdf_div_up(c.N, 128), 128, 0, st>>>(A, c.fa, ldx, c.M, c.N, cs + c.N);
        rc = msdf_tc::launch_wgrad(A, c.fa, ldx, c.Np, B, c.fb, ldy, c.Kp, c.M, e, st, "msdf_tc_selftest(wgrad)", with_cs ? cs : nullptr, c.N, 0);
        if (!rc) k_maxerr<<<64, 256, 0, st>>>(C, R, (int64_t)c.N * ldc, out);
        if (!rc && with_cs) k_maxerr<<<1, 256, 0, st>>>(cs, cs + c.N, c.N, out);
        cudaStreamSynchronize(st);
        cudaFree(cs);
    }
    cudaError_t e2 = cudaStreamSynchronize(st);
    if (!rc && e2 != cudaSuccess) { msdf_set_error("msdf_tc_selftest: kernel failed: %s", cudaGetErrorString(e2)); rc = MSDF_ERR_CUDA; }
    if (!rc) cudaMemcpy(result_host, out, 8, cudaMemcpyDeviceToHost);
    cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(R); cudaFree(out);
    return rc;
}
