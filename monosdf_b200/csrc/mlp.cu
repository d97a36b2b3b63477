// The neural field (SDF/feature MLP + colour MLP) forward and analytic backward.
//
// Replaces (reference code/model/network.py): ImplicitNetwork.forward/get_outputs/gradient_sdf/get_sdf_vals
// :79-137, ImplicitNetworkGrid :247-309, RenderingNetwork.forward :389-470, embedder.py:5-50, and the
// autograd double backward the reference obtains from torch.autograd.grad(create_graph=True) (:121-127).
//
// Per chunk of points the library runs explicit sweeps of GEMMs with fused epilogues:
//   forward sweep   h_{l+1} = softplus100(W_l u_l + b_l)                       u_l = h_l or [h_l, h_0]/sqrt2
//   reverse sweep   a_{l-1} = (W_l^T a_l)|_h * sigma_{l-1},  g0 = d sdf/d h_0   (analytic grad_x sdf)
//   tangent sweep   t_{l+1} = (W_l t_l) * sigma_l,  z_l = (W_l t_l) a_l 100 (1 - sigma_l)     (backward only)
//   backward sweep  pbar_{l-1} = (W_l^T pbar_l)|_h * sigma_{l-1} + z_{l-1}                    (backward only)
//   weight grads    dW_l += pbar_l^T u_l + a_l^T t_l,  db_l += colsum(pbar_l)
// where sigma_l = sigmoid(100 p_l) is recovered from the stored post-activation as -expm1(-100 h).
// The library keeps no state: either the caller hands both calls a buffer for the chunk activations (saved mode: the
// forward leaves h_l, a_l, the colour rows there and the backward reads them back) or the backward recomputes the chunk.
//
// The sweeps are written once, generic over the activation element type T:
//   T = float          fp32 mode: SIMT FFMA GEMMs (gemm_f32.cuh), the reference's arithmetic class (1e-4 parity)
//   T = __nv_bfloat16  tensor-core mode: tcgen05 kernels with TMEM accumulators and TMA-fed operands -- the forward sweep
//                      and the sdf-only queries as ONE kernel with the activations on chip (fused_mlp.cuh), the reverse
//                      sweep as one chained kernel (tc_chain.cuh), tangent / backward layers with every operand streamed
//                      by TMA (tc_stream.cuh; z_l is rebuilt in the backward epilogue instead of stored), weight
//                      gradients and the colour net on tc_gemm.cuh; forward-like matrices stored in fp16, adjoint-like
//                      ones in bf16 (Fw<T> below), fp32 accumulation (2e-2 parity)
#include <stdlib.h>
#include <type_traits>

#include "gemm_f32.cuh"
#include "tc_gemm.cuh"
#include "fused_mlp.cuh"
#include "tc_stream.cuh"
#include "tc_chain.cuh"

int msdf_hash_forward_rows(const float* x, const float* table, const int* offsets, float* out, int64_t out_ld,
                           int64_t B, int C, int L, float S, uint32_t H, float divide_factor, float* dy_dx, cudaStream_t st);
int msdf_hash_scatter_rows(const float* x, const int* offsets, int64_t B, int C, int L, float S, uint32_t H, float divide_factor,
                           const float* grad, int64_t grad_ld, const float* grad2, int64_t grad2_ld, const float* gg_x,
                           float gg_scale, float* grad_table, cudaStream_t st);

namespace {

using bf16 = __nv_bfloat16;
using msdf_gemm::kNN;
using msdf_gemm::kNT;
using msdf_gemm::kTN;

constexpr float kInvSqrt2 = 0.70710678118654752440f;
constexpr float kSqrt2 = 1.41421356237309504880f;

using f16 = __half;

// T is the MODE tag of the sweeps: float (fp32 mode) or bf16 (tensor-core mode).  In tensor-core mode the matrices
// come in two 16-bit formats, both legal tcgen05.mma.kind::f16 operands in any combination:
//   Fw<T> = fp16  "forward-like" values of bounded range where resolution matters: activations h_l, the reverse-sweep
//                 state a_l, the colour-net rows and every weight copy (11 significant bits -> 8x smaller sdf /
//                 grad error than bf16; stores saturate at +-65504)
//   T     = bf16  adjoint / tangent values (t_l, z_l, pbar_l, output adjoints): linear in the upstream gradient,
//                 hence of arbitrary scale -> they need fp32's exponent range
template <class T> constexpr bool kIsBf16 = std::is_same<T, bf16>::value;
template <class T> struct FwdOf { using type = T; };
#ifndef MSDF_FWD_BF16
template <> struct FwdOf<bf16> { using type = f16; };
#endif
template <class T> using Fw = typename FwdOf<T>::type;
template <class E> struct Fmt16 { static constexpr int value = -1; };                  // fp32 mode: no 16-bit format
template <> struct Fmt16<f16> { static constexpr int value = msdf_tc::kF16; };
template <> struct Fmt16<bf16> { static constexpr int value = msdf_tc::kBF16; };

// ----------------------------------------------------------------------------------------------------------
// element access
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(bf16* p, float v) { *p = __float2bfloat16(v); }
__device__ __forceinline__ float ldf(const f16* p) { return __half2float(*p); }
__device__ __forceinline__ float sat16(float v) { return fminf(fmaxf(v, -65504.f), 65504.f); }
__device__ __forceinline__ void stf(f16* p, float v) { *p = __float2half_rn(sat16(v)); }

// W consecutive elements of a row; vectorised when the whole group is valid and 16-byte aligned
template <int W>
__device__ __forceinline__ void load_row(const float* p, float* o, int nv) {
#pragma unroll
    for (int j = 0; j < W; ++j) o[j] = j < nv ? p[j] : 0.f;
}
template <int W>
__device__ __forceinline__ void load_row(const bf16* p, float* o, int nv) {
    if (W == 8 && nv == 8 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        const uint4 q = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            o[2 * j] = __uint_as_float(w[j] << 16);
            o[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < W; ++j) o[j] = j < nv ? __bfloat162float(p[j]) : 0.f;
}
template <int W>
__device__ __forceinline__ void load_row(const f16* p, float* o, int nv) {
    if (W == 8 && nv == 8 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        const uint4 q = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            o[2 * j] = msdf_tc::WarpIO::lo_of<msdf_tc::kF16>(w[j]);
            o[2 * j + 1] = msdf_tc::WarpIO::hi_of<msdf_tc::kF16>(w[j]);
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < W; ++j) o[j] = j < nv ? __half2float(p[j]) : 0.f;
}
template <int W>
__device__ __forceinline__ void store_row(float* p, const float* v, int nv) {
#pragma unroll
    for (int j = 0; j < W; ++j)
        if (j < nv) p[j] = v[j];
}
template <int W>
__device__ __forceinline__ void store_row(bf16* p, const float* v, int nv) {
    if (W == 8 && nv == 8 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            w[j] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
        return;
    }
#pragma unroll
    for (int j = 0; j < W; ++j)
        if (j < nv) p[j] = __float2bfloat16(v[j]);
}
template <int W>
__device__ __forceinline__ void store_row(f16* p, const float* v, int nv) {
    if (W == 8 && nv == 8 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = msdf_tc::WarpIO::pack2<msdf_tc::kF16>(v[2 * j], v[2 * j + 1]);
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
        return;
    }
#pragma unroll
    for (int j = 0; j < W; ++j)
        if (j < nv) stf(p + j, v[j]);
}

// activation math.  FAST = bf16 mode: branch-free MUFU ex2/lg2 approximations with flush-to-zero (error far below
// bf16 resolution; no denormal slow paths, no divergent branches in the GEMM epilogues); precise (libdevice) otherwise.
__device__ __forceinline__ float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr float k100Log2e = 144.26950408889634f;       // 100 / ln 2
constexpr float kLn2Over100 = 0.0069314718055994531f;  // ln 2 / 100

template <bool FAST>
__device__ __forceinline__ float softplus100(float p) {   // nn.Softplus(beta=100), threshold 20 (network.py:77)
    // fast: max(p, 0) + log(1 + exp(-|100 p|)) / 100; past the threshold 1 + exp(.) rounds to 1, so the result is p
    if (FAST) return fmaxf(p, 0.f) + fast_lg2(1.0f + fast_ex2(-fabsf(p) * k100Log2e)) * kLn2Over100;
    const float bp = 100.f * p;
    return bp > 20.f ? p : log1pf(expf(bp)) / 100.f;
}
// sigmoid(100 p) from h = softplus100(p):  1 - exp(-100 h)
template <bool FAST>
__device__ __forceinline__ float sig_from_h(float h) { return FAST ? 1.0f - fast_ex2(-k100Log2e * h) : -expm1f(-100.f * h); }
// (sigmoid, 100 (1 - sigmoid)) from h; the second is exactly 0 past the softplus threshold like torch's double backward
template <bool FAST>
__device__ __forceinline__ void sig_dsig_from_h(float h, float& s, float& d) {
    const float t = 100.f * h;
    const float e = FAST ? fast_ex2(-k100Log2e * h) : expf(-t);
    s = FAST ? 1.0f - e : -expm1f(-t);
    d = t > 20.f ? 0.f : 100.f * e;
}

// ----------------------------------------------------------------------------------------------------------
// network geometry
// ----------------------------------------------------------------------------------------------------------
struct Net {
    int L, d0, skip;
    int in[MSDF_MAX_LAYERS], out[MSDF_MAX_LAYERS];
    int64_t ldw[MSDF_MAX_LAYERS];
    const float* W[MSDF_MAX_LAYERS];
    const float* b[MSDF_MAX_LAYERS];
    int maxw;
    // 16-bit copies for the tensor-core path (prepared per call): Wk = W (K-major over the inputs), Wt = W^T, each in
    // the forward format (forward / reverse sweeps) and in bf16 (tangent / backward sweeps): one MMA takes both
    // operands in the same format
    Fw<bf16>* Wk[MSDF_MAX_LAYERS]; Fw<bf16>* Wt[MSDF_MAX_LAYERS];
    bf16* Wkb[MSDF_MAX_LAYERS]; bf16* Wtb[MSDF_MAX_LAYERS];
    const void* wk(int l, int fmt) const { return fmt == msdf_tc::kBF16 ? (const void*)Wkb[l] : (const void*)Wk[l]; }
    const void* wt(int l, int fmt) const { return fmt == msdf_tc::kBF16 ? (const void*)Wtb[l] : (const void*)Wt[l]; }
    int tap;           // colour net, spec variant: the first 3 outputs of layer `tap` leave the chain (diffuse colour) and
                       // layer tap + 1 takes the remaining ones; -1 otherwise
    int perm_last;     // 1: rows of the last layer are ordered [features..., sdf] in Wk/Wt (bf16 SDF net)
    int rot0;          // bf16 copies of layer 0: input column k' holds W[:, (k' + rot0) % in] (rotated colour input)
};

inline int round_up(int x, int a) { return (x + a - 1) / a * a; }
template <class T> inline int padw(int n) { return kIsBf16<T> ? round_up(n, 64) : round_up(n, 4); }

constexpr int kTapN = 3;   // diffuse colour channels tapped off the colour net (network.py:441)

// row order of the 16-bit weight copies / of the matrices those layers write: the first `shift` rows moved to the end
// (1: SDF net's last layer -> [features..., sdf]; 3: tapped colour layer -> [specular features..., diffuse])
inline int row_shift(const Net& n, int l) { return (n.perm_last && l == n.L - 1) ? 1 : ((n.tap >= 0 && l == n.tap) ? kTapN : 0); }

int make_net(const msdf_mlp_desc* d, Net& n, const char* who, int tap = -1) {
    MSDF_CHECK_ARG(d != nullptr, "%s: null network descriptor", who);
    MSDF_CHECK_ARG(d->n_layers >= 2 && d->n_layers <= MSDF_MAX_LAYERS, "%s: n_layers=%d not in [2,%d]", who, d->n_layers,
                   MSDF_MAX_LAYERS);
    n.L = d->n_layers; n.d0 = d->d0; n.skip = d->skip_layer; n.perm_last = 0; n.rot0 = 0; n.tap = tap;
    MSDF_CHECK_ARG(tap < 0 || (tap >= 1 && tap + 1 < n.L), "%s: the diffuse/specular split needs at least 4 layers", who);
    MSDF_CHECK_ARG(n.skip < n.L && n.skip != 0, "%s: skip_layer=%d invalid", who, n.skip);
    int w = 0;
    for (int l = 0; l < n.L; ++l) {
        n.in[l] = d->in_dim[l]; n.out[l] = d->out_dim[l]; n.ldw[l] = d->ldw[l]; n.W[l] = d->W[l]; n.b[l] = d->b[l];
        n.Wk[l] = n.Wt[l] = nullptr; n.Wkb[l] = n.Wtb[l] = nullptr;
        MSDF_CHECK_ARG(n.W[l] && n.b[l], "%s: layer %d has null weights", who, l);
        MSDF_CHECK_ARG(n.ldw[l] >= n.in[l] && n.in[l] > 0 && n.out[l] > 0, "%s: layer %d bad dims", who, l);
        const int expect = (l == 0) ? n.d0 : (l == n.skip ? n.out[l - 1] + n.d0 : (l == tap + 1 && tap >= 0 ? n.out[l - 1] - kTapN : n.out[l - 1]));
        MSDF_CHECK_ARG(n.in[l] == expect, "%s: layer %d in_dim=%d, expected %d", who, l, n.in[l], expect);
        if (l > 0) w = w > n.in[l] ? w : n.in[l];
        if (l < n.L - 1) w = w > n.out[l] ? w : n.out[l];
        if (l == tap) MSDF_CHECK_ARG(n.out[l] > kTapN, "%s: tapped layer too narrow", who);
    }
    n.maxw = w;
    return MSDF_OK;
}

// the tensor-core path keeps whole rows of every operand in 64-column TMA boxes
int check_tc_net(const Net& n, const char* who) {
    for (int l = 1; l < n.L; ++l)
        MSDF_CHECK_ARG(n.in[l] % 64 == 0 && n.in[l] <= 256, "%s: bf16 mode needs hidden widths that are multiples of 64 (<= 256); layer %d has %d",
                       who, l, n.in[l]);
    MSDF_CHECK_ARG(n.in[0] <= 640 && n.out[n.L - 1] <= 320, "%s: bf16 mode supports at most 640 input / 320 output columns", who);
    return MSDF_OK;
}

// ----------------------------------------------------------------------------------------------------------
// encoding kernels  (embedder.py:5-50; hash features come from hashgrid.cu)
// ----------------------------------------------------------------------------------------------------------
// value / derivative of PE column j (< 3 + 6*multires) at x: column order [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...]
__device__ __forceinline__ float pe_value(const float x[3], int j) {
    if (j < 3) return x[j];
    const int k = (j - 3) / 6, r = (j - 3) - 6 * k;
    const float f = (float)(1 << k);
    return r < 3 ? sinf(x[r] * f) : cosf(x[r - 3] * f);
}
// d PE_j / d x_d is non-zero only for d = pe_dim(j)
__device__ __forceinline__ int pe_dim(int j) { return j < 3 ? j : ((j - 3) % 3); }
__device__ __forceinline__ float pe_deriv(const float x[3], int j) {
    if (j < 3) return 1.0f;
    const int k = (j - 3) / 6, r = (j - 3) - 6 * k;
    const float f = (float)(1 << k);
    return r < 3 ? f * cosf(x[r] * f) : -f * sinf(x[r - 3] * f);
}

// H0[m, j] for j < cols: PE(x_m), then hash features (from hashf, or zeros), then zero padding
template <class T>
__global__ void k_encode(const float* __restrict__ x, int64_t M, int pe_w, int grid_w, const float* __restrict__ hashf,
                         T* __restrict__ H0, int64_t ld0, int cols) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * cols) return;
    const int64_t m = i / cols; const int j = (int)(i - m * cols);
    float v = 0.f;
    if (j < pe_w) {
        const float p[3] = {x[3 * m], x[3 * m + 1], x[3 * m + 2]};
        v = pe_value(p, j);
    } else if (j < pe_w + grid_w && hashf != nullptr) {
        v = hashf[m * grid_w + (j - pe_w)];
    }
    stf(H0 + m * ld0 + j, v);
}

// ---- row-per-thread producers of the padded operand rows (one thread computes a whole row and writes it in
// 8-column vectors; the per-column form above costs one thread, one set of loads and one sin/cos call per element)

// PE of a 3-vector: out[0..2] = x, then per frequency k: sin(2^k x) (3), cos(2^k x) (3).  sincosf keeps fp32 parity.
__device__ __forceinline__ void pe_row(const float p[3], int multires, float* out) {
    out[0] = p[0]; out[1] = p[1]; out[2] = p[2];
    for (int k = 0; k < multires; ++k) {
        const float f = (float)(1 << k);
#pragma unroll
        for (int d = 0; d < 3; ++d) sincosf(p[d] * f, &out[3 + 6 * k + d], &out[3 + 6 * k + 3 + d]);
    }
}

constexpr int kMaxPe = 3 + 6 * 16;

// H0[m, 0:cols) = [PE(x_m) | hash features or zeros | zero padding]
template <class T>
__global__ void __launch_bounds__(128)
k_encode_rows(const float* __restrict__ x, int64_t M, int multires, int pe_w, int grid_w, const float* __restrict__ hashf,
              T* __restrict__ H0, int64_t ld0, int cols) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float p[3] = {x[3 * m], x[3 * m + 1], x[3 * m + 2]};
    float pe[kMaxPe];
    pe_row(p, multires, pe);
    T* row = H0 + m * ld0;
    for (int j0 = 0; j0 < cols; j0 += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = j0 + u;
            v[u] = j < pe_w ? pe[j] : ((j < pe_w + grid_w && hashf != nullptr) ? hashf[m * grid_w + (j - pe_w)] : 0.f);
        }
        store_row<8>(row + j0, v, cols - j0 < 8 ? cols - j0 : 8);
    }
}

// The non-feature columns of the colour-net input row (see k_color_input), written as whole 8-column groups: only
// used when those columns form one aligned, contiguous range [c_lo, c_hi) of the (rotated) row, i.e. bf16 mode.
template <class T>
__global__ void __launch_bounds__(128)
k_color_input_rows(const float* __restrict__ x, const float* __restrict__ view, const float* __restrict__ normal,
                   const float* __restrict__ code, int64_t M, int n_samples, int mode_idr, int multires_view, int pe_w, int cd,
                   int code_per_ray, T* __restrict__ X, int64_t ldx, int c_lo, int c_hi) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int64_t ray = m / n_samples;
    const float d[3] = {view[3 * ray], view[3 * ray + 1], view[3 * ray + 2]};
    float pe[kMaxPe];
    pe_row(d, multires_view, pe);
    // rotated order of the range: [code (cd) | x (3, idr) | PE(view) (pe_w) | normal (3, idr) | zero padding]
    const int o_x = cd, o_pe = o_x + (mode_idr ? 3 : 0), o_n = o_pe + pe_w, o_end = o_n + (mode_idr ? 3 : 0);
    T* row = X + m * ldx + c_lo;
    for (int j0 = 0; j0 < c_hi - c_lo; j0 += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = j0 + u;
            float t = 0.f;
            if (j < o_x) t = code[(code_per_ray ? ray : 0) * cd + j];
            else if (j < o_pe) t = x[3 * m + (j - o_x)];
            else if (j < o_n) t = pe[j - o_pe];
            else if (j < o_end) t = normal[3 * m + (j - o_n)];
            v[u] = t;
        }
        store_row<8>(row + j0, v, 8);
    }
}

// Backward prologue, row part: dn[m] = mask (d_grad + dn_color); TG0[m, :] = J_enc dn (zero padded); and the columns
// [tail_lo, out_cols) of Dout: the sdf adjoint at sdf_col, zeros elsewhere (the feature columns are handled by
// k_backward_prologue only when they need work).
template <class T>
__global__ void __launch_bounds__(128)
k_backward_rows(const float* __restrict__ x, int64_t M, int multires, int pe_w, int grid_w, int n_levels, int level_dim,
                const float* __restrict__ dy_dx, float hash_chain, const float* __restrict__ mask,
                const float* __restrict__ d_sdf, const float* __restrict__ d_grad, const float* __restrict__ dn_color,
                int sdf_col, int tail_lo, T* __restrict__ Dout, int64_t ldo, int out_cols, float* __restrict__ dn,
                T* __restrict__ TG0, int64_t ldt, int t_cols) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float mk = mask[m];
    float v[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) v[d] = mk * ((d_grad ? d_grad[3 * m + d] : 0.f) + (dn_color ? dn_color[3 * m + d] : 0.f));
    dn[3 * m] = v[0]; dn[3 * m + 1] = v[1]; dn[3 * m + 2] = v[2];
    const float p[3] = {x[3 * m], x[3 * m + 1], x[3 * m + 2]};
    float pe[kMaxPe];
    pe_row(p, multires, pe);            // d/dx sin(f x) = f cos(f x), d/dx cos(f x) = -f sin(f x)
    T* trow = TG0 + m * ldt;
    const int d0 = pe_w + grid_w;
    // hash columns: J_hash dn from 16-byte loads of the point's dy_dx row (two levels = three loads)
    float hr[32];
    const bool hvec = dy_dx != nullptr && level_dim == 2 && (n_levels & 1) == 0 && n_levels <= 16;
    if (hvec) {
        const float4* d4 = reinterpret_cast<const float4*>(dy_dx + m * (int64_t)(n_levels * 6));
#pragma unroll
        for (int l2 = 0; l2 < 8; ++l2) {
            if (l2 < n_levels / 2) {
                const float4 a = d4[3 * l2], b = d4[3 * l2 + 1], c = d4[3 * l2 + 2];
                hr[4 * l2] = (a.x * v[0] + a.z * v[1] + b.x * v[2]) * hash_chain;
                hr[4 * l2 + 1] = (a.y * v[0] + a.w * v[1] + b.y * v[2]) * hash_chain;
                hr[4 * l2 + 2] = (b.z * v[0] + c.x * v[1] + c.z * v[2]) * hash_chain;
                hr[4 * l2 + 3] = (b.w * v[0] + c.y * v[1] + c.w * v[2]) * hash_chain;
            }
        }
    }
    for (int j0 = 0; j0 < t_cols; j0 += 8) {
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = j0 + u;
            float r = 0.f;
            if (j < 3) r = v[j];
            else if (j < pe_w) {
                const int k = (j - 3) / 6, q = (j - 3) - 6 * k;
                const float f = (float)(1 << k);
                r = q < 3 ? f * pe[j + 3] * v[q] : -f * pe[j - 3] * v[q - 3];
            } else if (j < d0 && dy_dx != nullptr) {
                if (hvec) {
                    r = hr[j - pe_w];
                } else {
                    const int qq = j - pe_w, l = qq / level_dim, c = qq - l * level_dim;
                    const float* dd = dy_dx + m * (int64_t)(n_levels * 3 * level_dim) + (l * 3) * level_dim + c;
                    r = (dd[0] * v[0] + dd[level_dim] * v[1] + dd[2 * level_dim] * v[2]) * hash_chain;
                }
            }
            t[u] = r;
        }
        store_row<8>(trow + j0, t, t_cols - j0 < 8 ? t_cols - j0 : 8);
    }
    if (tail_lo >= 0) {
        T* drow = Dout + m * ldo;
        const float ds = d_sdf ? mk * d_sdf[m] : 0.f;
        for (int j0 = tail_lo; j0 < out_cols; j0 += 8) {
            float t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = (j0 + u == sdf_col) ? ds : 0.f;
            store_row<8>(drow + j0, t, out_cols - j0 < 8 ? out_cols - j0 : 8);
        }
    }
}

template <class T>
__global__ void k_zero_cols(T* __restrict__ dst, int64_t ld, int64_t M, int c0, int w) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * w) return;
    const int64_t m = i / w; const int j = (int)(i - m * w);
    stf(dst + m * ld + c0 + j, 0.f);
}

// dst[m, c0 + j] = src[m, j] * scale   (the [.., h0]/sqrt2 half of the skip concat, network.py:88-89)
template <class T>
__global__ void k_skip_copy(const T* __restrict__ src, int64_t lds, T* __restrict__ dst, int64_t ldd, int64_t M, int w, int c0,
                            float scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * w) return;
    const int64_t m = i / w; const int j = (int)(i - m * w);
    stf(dst + m * ldd + c0 + j, ldf(src + m * lds + j) * scale);
}

// grad_x = J_enc(x)^T g0 (+ hash chain), then the bounding-sphere clamp of get_outputs (network.py:116-118):
//   sdf = min(sdf_raw, sphere_scale (R - |x|)); the gradient follows the selected branch (ties split 1/2, like
//   torch.minimum's backward).  mask[m] = d sdf / d sdf_raw  in {1, 0, 0.5}.
__global__ void k_decode(const float* __restrict__ x, const float* __restrict__ g0, const float* __restrict__ g0b, int64_t ldg, int64_t M, int pe_w,
                         int grid_w, int n_levels, int level_dim, const float* __restrict__ dy_dx, float hash_chain,
                         const float* __restrict__ sdf_raw, float clamp_radius, float sphere_scale,
                         float* __restrict__ sdf_out, float* __restrict__ grad_out, float* __restrict__ mask_out) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float p[3] = {x[3 * m], x[3 * m + 1], x[3 * m + 2]};
    float g[3] = {0.f, 0.f, 0.f};
    if (g0 != nullptr) {
        // A thread walks its own rows: 16-byte loads (ldg % 4 == 0, 16-byte aligned bases) keep the LSU transaction
        // count down -- with scalar loads the kernel was bound by it (165 us per 262 144 points of the grid conf).
        const float* gr = g0 + m * ldg;
        const float4* gr4 = reinterpret_cast<const float4*>(gr);
        const float4* gb4 = g0b != nullptr ? reinterpret_cast<const float4*>(g0b + m * ldg) : nullptr;
        float pe[kMaxPe];
        pe_row(p, (pe_w - 3) / 6, pe);      // d/dx sin(f x) = f cos(f x), d/dx cos(f x) = -f sin(f x): the row's own entries
        for (int c4 = 0; 4 * c4 < pe_w; ++c4) {
            float4 a = gr4[c4];
            if (gb4 != nullptr) { const float4 b = gb4[c4]; a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = 4 * c4 + e;
                if (j >= pe_w) break;
                float dv = 1.0f; int dim = j;
                if (j >= 3) {
                    const int k = (j - 3) / 6, q = (j - 3) - 6 * k;
                    const float f = (float)(1 << k);
                    dv = q < 3 ? f * pe[j + 3] : -f * pe[j - 3];
                    dim = q < 3 ? q : q - 3;
                }
                const float t = av[e] * dv;
                if (dim == 0) g[0] += t; else if (dim == 1) g[1] += t; else g[2] += t;
            }
        }
        if (grid_w > 0 && dy_dx != nullptr) {
            const float* dd = dy_dx + m * (int64_t)(n_levels * 3 * level_dim);
            float h[3] = {0.f, 0.f, 0.f};
            if (level_dim == 2 && (n_levels & 1) == 0) {     // two levels = 12 derivatives = three 16-byte loads
                const float4* d4 = reinterpret_cast<const float4*>(dd);
                for (int l2 = 0; l2 < n_levels / 2; ++l2) {
                    const float4 a = d4[3 * l2], b = d4[3 * l2 + 1], c = d4[3 * l2 + 2];
                    const float* gg = gr + pe_w + 4 * l2;
                    const float g0v = gg[0], g1v = gg[1], g2v = gg[2], g3v = gg[3];
                    h[0] += g0v * a.x; h[0] += g1v * a.y; h[1] += g0v * a.z; h[1] += g1v * a.w; h[2] += g0v * b.x; h[2] += g1v * b.y;
                    h[0] += g2v * b.z; h[0] += g3v * b.w; h[1] += g2v * c.x; h[1] += g3v * c.y; h[2] += g2v * c.z; h[2] += g3v * c.w;
                }
            } else {
                for (int l = 0; l < n_levels; ++l)
                    for (int d = 0; d < 3; ++d)
                        for (int c = 0; c < level_dim; ++c) h[d] += gr[pe_w + l * level_dim + c] * dd[(l * 3 + d) * level_dim + c];
            }
            g[0] += h[0] * hash_chain; g[1] += h[1] * hash_chain; g[2] += h[2] * hash_chain;
        }
    }
    float s = sdf_raw[m], mk = 1.0f;
    if (clamp_radius > 0.f) {
        const float nrm = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
        const float sphere = sphere_scale * (clamp_radius - nrm);
        if (sphere < s || sphere == s) {
            const float w = (sphere == s) ? 0.5f : 0.0f;
            const float inv = nrm > 0.f ? -sphere_scale / nrm : 0.f;
            g[0] = w * g[0] + (1.f - w) * inv * p[0];
            g[1] = w * g[1] + (1.f - w) * inv * p[1];
            g[2] = w * g[2] + (1.f - w) * inv * p[2];
            s = sphere; mk = w;
        }
    }
    if (sdf_out) sdf_out[m] = s;
    if (grad_out) { grad_out[3 * m] = g[0]; grad_out[3 * m + 1] = g[1]; grad_out[3 * m + 2] = g[2]; }
    if (mask_out) mask_out[m] = mk;
}

// Backward prologue: dn[m] = mask * (d_grad[m] + dn_color[m]);  Dout[m, sdf_col] = mask * d_sdf[m];
//   Dout[m, feat_col0 + j] (+)= d_feat[m,j], remaining columns of Dout zero;  TG0[m, :] = J_enc dn (zero padded).
template <class T>
__global__ void k_backward_prologue(const float* __restrict__ x, int64_t M, int pe_w, int grid_w, int n_levels, int level_dim,
                                    const float* __restrict__ dy_dx, float hash_chain, const float* __restrict__ mask,
                                    const float* __restrict__ d_sdf, const float* __restrict__ d_grad,
                                    const float* __restrict__ dn_color, const float* __restrict__ d_feat, int64_t ld_dfeat,
                                    int feat_w, int have_color_feat, int sdf_col, int feat_col0, T* __restrict__ Dout, int64_t ldo,
                                    int out_cols, float* __restrict__ dn, T* __restrict__ TG0, int64_t ldt, int t_cols) {
    const int d0 = pe_w + grid_w;
    const int cols = t_cols > out_cols ? t_cols : out_cols;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * cols) return;
    const int64_t m = i / cols; const int j = (int)(i - m * cols);
    const float mk = mask[m];
    float v[3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
        v[d] = mk * ((d_grad ? d_grad[3 * m + d] : 0.f) + (dn_color ? dn_color[3 * m + d] : 0.f));
    if (j < 3) dn[3 * m + j] = v[j];
    if (j < out_cols) {
        float f = 0.f;
        if (j == sdf_col) f = d_sdf ? mk * d_sdf[m] : 0.f;
        else if (j >= feat_col0 && j < feat_col0 + feat_w) {
            if (have_color_feat) f = ldf(Dout + m * ldo + j);
            if (d_feat) f += d_feat[m * ld_dfeat + (j - feat_col0)];
        }
        stf(Dout + m * ldo + j, f);
    }
    if (j < t_cols) {
        float t = 0.f;
        if (j < pe_w) {
            const float p[3] = {x[3 * m], x[3 * m + 1], x[3 * m + 2]};
            t = pe_deriv(p, j) * v[pe_dim(j)];
        } else if (j < d0 && dy_dx != nullptr) {
            const int q = j - pe_w, l = q / level_dim, c = q - l * level_dim;
            const float* dd = dy_dx + m * (int64_t)(n_levels * 3 * level_dim) + (l * 3) * level_dim + c;
            t = (dd[0] * v[0] + dd[level_dim] * v[1] + dd[2 * level_dim] * v[2]) * hash_chain;
        }
        stf(TG0 + m * ldt + j, t);
    }
}

// out[n] += sum_m w[m*ws] * X[m*ldx + n]   (w == nullptr -> 1).  Bias gradients and the sdf row of the last layer.
template <class T>
__global__ void __launch_bounds__(256)
k_wcolsum(const T* __restrict__ X, int64_t ldx, const T* __restrict__ w, int64_t ws, int64_t M, int N,
          int64_t rows_per_block, float* __restrict__ out) {
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = (r0 + rows_per_block < M) ? r0 + rows_per_block : M;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float s = 0.f;
        for (int64_t m = r0; m < r1; ++m) s += (w ? ldf(w + m * ws) : 1.0f) * ldf(X + m * ldx + n);
        atomicAdd(out + n, s);
    }
}

// ----------------------------------------------------------------------------------------------------------
// small-N layers: out[m,n] = act(sum_k A[m,k] W[n,k] + b[n]), N <= 4, one warp per row
// ----------------------------------------------------------------------------------------------------------
enum Act { kActNone = 0, kActSigmoid = 1, kActRelu = 2 };

template <int NMAX, class T>
__global__ void __launch_bounds__(256)
k_rowdot(const T* __restrict__ A, int64_t lda, const float* __restrict__ W, int64_t ldw, const float* __restrict__ b,
         int64_t M, int N, int K, int act, float* __restrict__ out, int64_t ldo) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (m >= M) return;
    float s[NMAX];
#pragma unroll
    for (int n = 0; n < NMAX; ++n) s[n] = 0.f;
    const T* a = A + m * lda;
    for (int k = lane; k < K; k += 32) {
        const float av = ldf(a + k);
#pragma unroll
        for (int n = 0; n < NMAX; ++n)
            if (n < N) s[n] = fmaf(av, __ldg(W + n * ldw + k), s[n]);
    }
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s[n] += __shfl_xor_sync(0xffffffffu, s[n], off);
    if (lane == 0) {
#pragma unroll
        for (int n = 0; n < NMAX; ++n)
            if (n < N) {
                float v = s[n] + b[n];
                if (act == kActSigmoid) v = 1.0f / (1.0f + expf(-v));
                else if (act == kActRelu) v = fmaxf(v, 0.f);
                out[m * ldo + n] = v;
            }
    }
}

// sdf head of the tensor-core mode: out[m] = sum_k A[m,k] W[k] + b, A 16-bit with K % 8 == 0 and 16-byte aligned rows.
// One warp per row like k_rowdot, but every lane reads 16-byte vectors (a 256-wide row is ONE request per warp instead
// of eight 64-byte ones) and keeps its slice of W in registers across the rows the warp walks over.
template <class TA>
__global__ void __launch_bounds__(256)
k_rowdot1_vec(const TA* __restrict__ A, int64_t lda, const float* __restrict__ W, const float* __restrict__ b, int64_t M, int K,
              float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    constexpr int kMaxVec = 2;                           // K <= 512
    float w[kMaxVec][8];
#pragma unroll
    for (int v = 0; v < kMaxVec; ++v)
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int k = (v * 32 + lane) * 8 + j; w[v][j] = k < K ? __ldg(W + k) : 0.f; }
    const float bias = b[0];
    constexpr int R = 4;                                 // rows in flight per warp: all loads issued before the math
    for (int64_t m0 = warp * R; m0 < M; m0 += nwarps * R) {
        uint4 q[R][kMaxVec];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int v = 0; v < kMaxVec; ++v) {
                const int k0 = (v * 32 + lane) * 8;
                q[r][v] = make_uint4(0u, 0u, 0u, 0u);
                if (m0 + r < M && k0 < K) q[r][v] = *reinterpret_cast<const uint4*>(A + (m0 + r) * lda + k0);
            }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float s = 0.f;
#pragma unroll
            for (int v = 0; v < kMaxVec; ++v) {
                constexpr int F = Fmt16<TA>::value;
                const uint32_t ww[4] = {q[r][v].x, q[r][v].y, q[r][v].z, q[r][v].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s = fmaf(msdf_tc::WarpIO::lo_of<F>(ww[j]), w[v][2 * j], s);
                    s = fmaf(msdf_tc::WarpIO::hi_of<F>(ww[j]), w[v][2 * j + 1], s);
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (lane == 0 && m0 + r < M) out[m0 + r] = s + bias;
        }
    }
}

__device__ __forceinline__ float act_grad(int act, float y) {   // d act / d pre as a function of the output y
    return act == kActSigmoid ? y * (1.0f - y) : (act == kActRelu ? (y > 0.f ? 1.f : 0.f) : 1.f);
}

// dA[m,k] = relu'(A[m,k]) * sum_n dpre[m,n] W[n,k],   dpre = d_out * act'(out)
template <int NMAX, class T>
__global__ void k_rowdot_dgrad(const float* __restrict__ d_out, const float* __restrict__ out, int64_t ldo, int act,
                               const float* __restrict__ W, int64_t ldw, const T* __restrict__ A, int64_t lda,
                               int64_t M, int N, int K, T* __restrict__ dA, int64_t ldda) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * K) return;
    const int64_t m = i / K; const int k = (int)(i - m * K);
    float s = 0.f;
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
        if (n < N) s = fmaf(d_out[m * ldo + n] * act_grad(act, out[m * ldo + n]), __ldg(W + n * ldw + k), s);
    stf(dA + m * ldda + k, ldf(A + m * lda + k) > 0.f ? s : 0.f);
}

// dW[n,k] += sum_m dpre[m,n] A[m,k];  db[n] += sum_m dpre[m,n]
template <int NMAX, class T>
__global__ void __launch_bounds__(256)
k_rowdot_wgrad(const float* __restrict__ d_out, const float* __restrict__ out, int64_t ldo, int act,
               const T* __restrict__ A, int64_t lda, int64_t M, int N, int K, int64_t rows_per_block,
               float* __restrict__ dW, int64_t ldw, float* __restrict__ db) {
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = (r0 + rows_per_block < M) ? r0 + rows_per_block : M;
    for (int k = threadIdx.x; k < K + 1; k += blockDim.x) {
        float s[NMAX];
#pragma unroll
        for (int n = 0; n < NMAX; ++n) s[n] = 0.f;
        for (int64_t m = r0; m < r1; ++m) {
            const float av = (k < K) ? ldf(A + m * lda + k) : 1.0f;
#pragma unroll
            for (int n = 0; n < NMAX; ++n)
                if (n < N) s[n] = fmaf(d_out[m * ldo + n] * act_grad(act, out[m * ldo + n]), av, s[n]);
        }
#pragma unroll
        for (int n = 0; n < NMAX; ++n)
            if (n < N) {
                if (k < K) atomicAdd(dW + n * ldw + k, s[n]); else atomicAdd(db + n, s[n]);
            }
    }
}

// ----------------------------------------------------------------------------------------------------------
// GEMM epilogues.  Each provides run<W>(m, n, v, nv): W consecutive columns n..n+W-1 of row m, the first nv of
// them inside the GEMM's N; operator() overloads adapt to the SIMT engine (W = 4) and the tcgen05 engine (W = 8).
// ----------------------------------------------------------------------------------------------------------
template <class Derived>
struct EpiBase {
    int N;   // real number of output columns (the tcgen05 engine runs on a padded N)
    __device__ __forceinline__ void operator()(int64_t m, int n, const float v[4], int nv) const {
        static_cast<const Derived*>(this)->template run<4>(m, n, v, nv);
    }
    // tcgen05 engine: number of valid columns of the 32-column chunk starting at n0 (<= 0: nothing to do)
    __device__ __forceinline__ int chunk_cols(int n0) const { const int nv = N - n0; return nv < 32 ? nv : 32; }
    static constexpr int kPre = 0;                 // epilogues that read operands back override kPre / prefetch
    static constexpr int kStores = 1;              // bf16 matrices written per element (roofline accounting)
    __device__ __forceinline__ void prefetch(const msdf_tc::WarpIO&, int, uint4*) const {}
    __device__ __forceinline__ const float* colvec() const { return nullptr; }
};

using msdf_tc::WarpIO;

template <class T>
struct EpiFwdAct : EpiBase<EpiFwdAct<T>> {   // out[m,n] = softplus100(acc + b[n]) * oscale
    using TF = Fw<T>;
    static constexpr int FO = Fmt16<TF>::value;
    const float* bias; TF* out; int64_t ldo; float oscale;
    template <int W> __device__ __forceinline__ void run(int64_t m, int n, const float* v, int nv) const {
        float o[W];
#pragma unroll
        for (int j = 0; j < W; ++j) o[j] = j < nv ? softplus100<kIsBf16<T>>(v[j] + __ldg(bias + n + j)) * oscale : 0.f;
        store_row<W>(out + m * ldo + n, o, nv);
    }
    __device__ __forceinline__ void chunk(const WarpIO& io, int n0, float v[32], const uint4* q) const {
        const int nv = this->chunk_cols(n0);
        if (nv <= 0) return;
#pragma unroll
        float b[32];
        io.colvec(n0, b);
        const float os = oscale, c2 = kLn2Over100 * oscale;
#pragma unroll
        for (int j = 0; j < 32; ++j) {          // softplus100(p) * oscale, branch free (see softplus100<true>)
            const float p = v[j] + b[j];
            v[j] = fmaf(fast_lg2(1.0f + fast_ex2(-fabsf(p) * k100Log2e)), c2, fmaxf(p, 0.f) * os);
        }
        io.store<FO>(out, ldo, n0, v, nv);
    }
    __device__ __forceinline__ const float* colvec() const { return bias; }
};
template <class T>
struct EpiBias : EpiBase<EpiBias<T>> {       // out[m,n] = acc + b[n]
    using TF = Fw<T>;
    static constexpr int FO = Fmt16<TF>::value;
    const float* bias; TF* out; int64_t ldo;
    template <int W> __device__ __forceinline__ void run(int64_t m, int n, const float* v, int nv) const {
        float o[W];
#pragma unroll
        for (int j = 0; j < W; ++j) o[j] = j < nv ? v[j] + __ldg(bias + n + j) : 0.f;
        store_row<W>(out + m * ldo + n, o, nv);
    }
    __device__ __forceinline__ void chunk(const WarpIO& io, int n0, float v[32], const uint4* q) const {
        const int nv = this->chunk_cols(n0);
        if (nv <= 0) return;
#pragma unroll
        float b[32];
        io.colvec(n0, b);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += b[j];
        io.store<FO>(out, ldo, n0, v, nv);
    }
    __device__ __forceinline__ const float* colvec() const { return bias; }
};
template <class T>
struct EpiRelu : EpiBase<EpiRelu<T>> {       // out[m,n] = relu(acc + b[n])
    using TF = Fw<T>;
    static constexpr int FO = Fmt16<TF>::value;
    const float* bias; TF* out; int64_t ldo;
    template <int W> __device__ __forceinline__ void run(int64_t m, int n, const float* v, int nv) const {
        float o[W];
#pragma unroll
        for (int j = 0; j < W; ++j) o[j] = j < nv ? fmaxf(v[j] + __ldg(bias + n + j), 0.f) : 0.f;
        store_row<W>(out + m * ldo + n, o, nv);
    }
    __device__ __forceinline__ void chunk(const WarpIO& io, int n0, float v[32], const uint4* q) const {
        const int nv = this->chunk_cols(n0);
        if (nv <= 0) return;
#pragma unroll
        float b[32];
        io.colvec(n0, b);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j] + b[j], 0.f);
        io.store<FO>(out, ldo, n0, v, nv);
    }
    __device__ __forceinline__ const float* colvec() const { return bias; }
};
// C[row(m), n] += acc  (split-K weight gradients).  With perm_rows > 0 the GEMM's row index i addresses a layer whose
// first perm_shift rows were moved to the end (row_shift()): i -> row (i + perm_shift) % perm_rows.
struct EpiAtomic : EpiBase<EpiAtomic> {
    float* C; int64_t ldc; int Mrows; int perm_rows; int perm_shift;
    int col_rot;                                   // GEMM column c addresses column (c + col_rot) % N (rotated colour input)
    template <int W> __device__ __forceinline__ void run(int64_t m, int n, const float* v, int nv) const {
        if (m >= Mrows) return;
        const int64_t r = perm_rows > 0 ? (m + perm_shift) % perm_rows : m;
        float* o = C + r * ldc + n;
#pragma unroll
        for (int j = 0; j < W; ++j)
            if (j < nv) atomicAdd(o + j, v[j]);
    }
    __device__ __forceinline__ void chunk(const WarpIO& io, int n0, float v[32], const uint4* q) const {
        const int nv = this->chunk_cols(n0);
        if (nv <= 0) return;
        const EpiAtomic e = *this;
        if (nv == 32 && col_rot == 0 && (ldc & 3) == 0 && (n0 & 3) == 0 && ((reinterpret_cast<uintptr_t>(C) & 15) == 0)) {
            io.atomic_add_v4(v, (int64_t)Mrows, [=](int64_t m) -> float* {
                const int64_t r = e.perm_rows > 0 ? (m + e.perm_shift) % e.perm_rows : m;
                return e.C + r * e.ldc + n0;
            });
            return;
        }
        io.atomic_add(v, (int64_t)Mrows, [=](int64_t m, int c) -> float* {
            if (c >= nv) return nullptr;
            const int64_t r = e.perm_rows > 0 ? (m + e.perm_shift) % e.perm_rows : m;
            int col = n0 + c + e.col_rot;
            if (col >= e.N) col -= e.N;
            return e.C + r * e.ldc + col;
        });
    }
};
// reverse sweep, layer l: acc = (a_l W_l)[m,n], n over the layer's inputs
template <class T>
struct EpiRev : EpiBase<EpiRev<T>> {
    using TF = Fw<T>;
    static constexpr int FF = Fmt16<TF>::value;
    const TF* Hin; int64_t ldh; float hscale;      // stored input of layer l and the factor that undoes its scaling
    TF* Aout; int64_t lda;                         // a_{l-1}
    float* g0; int64_t ldg;
    int dh; float qscale;                          // skip layer: columns >= dh are the h0 half; both halves scaled 1/sqrt2
    int layer0, g0_accum;
    template <int W> __device__ __forceinline__ void run(int64_t m, int n, const float* v, int nv) const {
        if (!layer0 && n + nv <= dh) {             // fast path: a whole group of hidden columns
            float h[W], o[W];
            load_row<W>(Hin + m * ldh + n, h, nv);
#pragma unroll
            for (int j = 0; j < W; ++j) o[j] = v[j] * qscale * sig_from_h<kIsBf16<T>>(h[j] * hscale);
            store_row<W>(Aout + m * lda + n, o, nv);
            return;
        }
#pragma unroll
        for (int j = 0; j < W; ++j) {
            if (j >= nv) break;
            const int c = n + j;
            const float r = v[j] * qscale;
            if (c >= dh) { g0[m * ldg + (c - dh)] = r; continue; }
            if (layer0) { float* g = g0 + m * ldg + c; *g = g0_accum ? *g + r : r; }
            else stf(Aout + m * lda + c, r * sig_from_h<kIsBf16<T>>(ldf(Hin + m * ldh + c) * hscale));
        }
    }
    __device__ __forceinline__ void chunk(const WarpIO& io, int n0, float v[32], const uint4* q) const {
        const int nv = this->chunk_cols(n0);
        if (nv <= 0) return;
        if (layer0) {                              // everything is d sdf / d h0
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= qscale;
            io.store_f32(g0, ldg, n0, v, 0, nv, g0_accum != 0);
            return;
        }
        const int nh = dh - n0 < nv ? dh - n0 : nv;   // hidden columns of this chunk
        if (nh < nv) {                             // h0 half of the skip concat
            float r[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = v[j] * qscale;
            io.store_f32(g0, ldg, n0 - dh, r, nh > 0 ? nh : 0, nv, false);
        }
        if (nh > 0) {
            float h[32];
            io.unstage<FF>(q, h);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v[j] * qscale * sig_from_h<true>(h[j] * hscale);
            io.store<FF>(Aout, lda, n0, v, nh);
        }
    }
    static constexpr int kPre = 1;
    __device__ __forceinline__ void prefetch(const WarpIO& io, int n0, uint4* q) const {
        if (!layer0 && n0 < dh && n0 < this->N) io.prefetch(Hin, ldh, n0, q);
    }
};
// tangent sweep, layer l: acc = (t_l W_l^T)[m,n], n over the layer's outputs:  t_{l+1} = acc * sigma_l * tscale.
// (z_l = acc a_l 100 (1 - sigma_l) is NOT formed here any more: the backward sweep rebuilds it from t_{l+1}, a_l and the
// same sigma -- this sweep then reads one operand back and writes one matrix, like the reverse sweep, instead of 2 + 2.)
template <class T>
struct EpiTan : EpiBase<EpiTan<T>> {
    using TF = Fw<T>;
    static constexpr int FF = Fmt16<TF>::value, FA = Fmt16<T>::value;
    const TF* Hn; int64_t ldh; float hscale;       // h_{l+1} as stored (input of layer l+1)
    T* Tout; int64_t ldt; float tscale;
    template <int W> __device__ __forceinline__ void run(int64_t m, int n, const float* v, int nv) const {
        float h[W], t[W];
        load_row<W>(Hn + m * ldh + n, h, nv);
#pragma unroll
        for (int j = 0; j < W; ++j) t[j] = v[j] * sig_from_h<kIsBf16<T>>(h[j] * hscale) * tscale;
        store_row<W>(Tout + m * ldt + n, t, nv);
    }
    __device__ __forceinline__ void chunk(const WarpIO& io, int n0, float v[32], const uint4* q) const {
        const int nv = this->chunk_cols(n0);
        if (nv <= 0) return;
        float h[32];
        io.unstage<FF>(q, h);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = v[j] * sig_from_h<true>(h[j] * hscale) * tscale;
        io.store<FA>(Tout, ldt, n0, v, nv);
    }
    static constexpr int kPre = 1;
    __device__ __forceinline__ void prefetch(const WarpIO& io, int n0, uint4* q) const {
        if (n0 < this->N) io.prefetch(Hn, ldh, n0, q);
    }
};
// backward sweep, layer l: acc = (pbar_l W_l)[m,n], n over the layer's inputs:
//   pbar_{l-1} = acc * qscale * sigma_{l-1} + z_{l-1},   z_{l-1} = (W_{l-1} t_{l-1}) a_{l-1} 100 (1 - sigma_{l-1})
// with (W_{l-1} t_{l-1}) recovered from the stored tangent t_l = (W_{l-1} t_{l-1}) sigma_{l-1} tscale_l  (a_{l-1} carries a
// factor sigma_{l-1} itself, so sigma = 0 means z = 0 on both routes).
template <class T>
struct EpiBwd : EpiBase<EpiBwd<T>> {
    using TF = Fw<T>;
    static constexpr int FF = Fmt16<TF>::value, FA = Fmt16<T>::value;
    const TF* Hin; int64_t ldh; float hscale;      // h_l as stored (input of layer l): sigma_{l-1}
    const T* Tin; int64_t ldt; float inv_tscale;    // t_l (adjoint format) and 1 / its storage scale
    const TF* Ain; int64_t lda;                     // a_{l-1}
    T* Pout; int64_t ldp;                           // pbar_{l-1}
    float* bh0; int64_t ldb;                       // adjoint of h_0 (hash-grid nets only), may be null
    int dh; float qscale; int layer0, bh0_accum;
    template <bool FAST>
    static __device__ __forceinline__ float combine(float acc_q, float h, float t, float a, float inv_ts) {
        float sg, d;
        sig_dsig_from_h<FAST>(h, sg, d);
        const float z = sg > 0.f ? (FAST ? __fdividef(t * inv_ts, sg) : (t * inv_ts) / sg) * a * d : 0.f;
        return acc_q * sg + z;
    }
    template <int W> __device__ __forceinline__ void run(int64_t m, int n, const float* v, int nv) const {
        if (!layer0 && n + nv <= dh) {
            float h[W], t[W], a[W], o[W];
            load_row<W>(Hin + m * ldh + n, h, nv);
            load_row<W>(Tin + m * ldt + n, t, nv);
            load_row<W>(Ain + m * lda + n, a, nv);
#pragma unroll
            for (int j = 0; j < W; ++j) o[j] = combine<kIsBf16<T>>(v[j] * qscale, h[j] * hscale, t[j], a[j], inv_tscale);
            store_row<W>(Pout + m * ldp + n, o, nv);
            return;
        }
#pragma unroll
        for (int j = 0; j < W; ++j) {
            if (j >= nv) break;
            const int c = n + j;
            const float r = v[j] * qscale;
            if (c >= dh) { if (bh0) bh0[m * ldb + (c - dh)] = r; continue; }
            if (layer0) { if (bh0) { float* g = bh0 + m * ldb + c; *g = bh0_accum ? *g + r : r; } }
            else stf(Pout + m * ldp + c, combine<kIsBf16<T>>(r, ldf(Hin + m * ldh + c) * hscale, ldf(Tin + m * ldt + c), ldf(Ain + m * lda + c), inv_tscale));
        }
    }
    __device__ __forceinline__ void chunk(const WarpIO& io, int n0, float v[32], const uint4* q) const {
        const int nv = this->chunk_cols(n0);
        if (nv <= 0) return;
        if (layer0) {
            if (bh0 == nullptr) return;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= qscale;
            io.store_f32(bh0, ldb, n0, v, 0, nv, bh0_accum != 0);
            return;
        }
        const int nh = dh - n0 < nv ? dh - n0 : nv;
        if (nh < nv && bh0 != nullptr) {
            float r[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = v[j] * qscale;
            io.store_f32(bh0, ldb, n0 - dh, r, nh > 0 ? nh : 0, nv, false);
        }
        if (nh > 0) {
            // three read-back operands, walked 8 columns at a time (16 packed words per operand would not fit next to the
            // prefetch ring): (h, t) give the sigma part and the z coefficient, then a
            const uint32_t bh = io.stage(q), bt = io.stage(q + 4);
            float w[32];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                uint32_t hw[4], tw[4];
                io.piece(bh, p, hw);
                io.piece(bt, p, tw);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float sg, d;
                    sig_dsig_from_h<true>(WarpIO::unpack<FF>(hw, j) * hscale, sg, d);
                    w[8 * p + j] = sg > 0.f ? __fdividef(WarpIO::unpack<FA>(tw, j) * inv_tscale, sg) * d : 0.f;
                    v[8 * p + j] = v[8 * p + j] * qscale * sg;
                }
            }
            __syncwarp();                                  // every lane is done with h before a is staged over it
            const uint32_t ba = io.stage(q + 8);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                uint32_t aw[4];
                io.piece(ba, p, aw);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[8 * p + j] = fmaf(w[8 * p + j], WarpIO::unpack<FF>(aw, j), v[8 * p + j]);
            }
            io.store<FA>(Pout, ldp, n0, v, nh);
        }
    }
    static constexpr int kPre = 3;
    __device__ __forceinline__ void prefetch(const WarpIO& io, int n0, uint4* q) const {
        if (!layer0 && n0 < dh && n0 < this->N) { io.prefetch(Hin, ldh, n0, q); io.prefetch(Tin, ldt, n0, q + 4); io.prefetch(Ain, lda, n0, q + 8); }
    }
};
// ---- the same three epilogues for k_tc_stream (tc_stream.cuh): read-back operands arrive as TMA boxes, a thread reads
// its own row of them 8 columns at a time.  Regular layers only (no h_0 columns, not layer 0): the callers fall back to
// the k_tc_gemm forms above for the skip layer and layer 0.  Tensor-core mode only.
template <class T>
struct EpiTanS : EpiBase<EpiTanS<T>> {            // operand 0: h_{l+1}
    using TF = Fw<T>;
    static constexpr int FF = Fmt16<TF>::value, FA = Fmt16<T>::value;
    static constexpr int kOps = 1;
    float hscale; T* Tout; int64_t ldt; float tscale;
    template <int W> __device__ __forceinline__ void run(int64_t, int, const float*, int) const {}
    __device__ __forceinline__ void chunk_smem(const WarpIO& io, int n0, float v[32], const msdf_tc::OpRow& r) const {
        const int nv = this->chunk_cols(n0);
        if (nv <= 0) return;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            uint32_t hw[4];
            r.piece(0, p, hw);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[8 * p + j] = v[8 * p + j] * sig_from_h<true>(WarpIO::unpack<FF>(hw, j) * hscale) * tscale;
        }
        io.store<FA>(Tout, ldt, n0, v, nv);
    }
};
template <class T, bool kG0 = false>
struct EpiRevS : EpiBase<EpiRevS<T, kG0>> {       // operand 0: h_l (sigma_{l-1})
    using TF = Fw<T>;
    static constexpr int FF = Fmt16<TF>::value;
    static constexpr int kOps = 1;
    float hscale; TF* Aout; int64_t lda;
    // kG0: columns >= dh are d sdf / d h0 (the h0 half of the skip concat; all of layer 0: dh = 0) and leave as fp32
    // rows of g0; both halves of the skip layer carry its 1/sqrt2
    int dh; float qscale; float* g0; int64_t ldg;
    template <int W> __device__ __forceinline__ void run(int64_t, int, const float*, int) const {}
    __device__ __forceinline__ void chunk_smem(const WarpIO& io, int n0, float v[32], const msdf_tc::OpRow& r) const {
        const int nv = this->chunk_cols(n0);
        if (nv <= 0) return;
        int nh = nv;
        if constexpr (kG0) {
            nh = dh - n0 < nv ? dh - n0 : nv;         // hidden columns of this chunk
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= qscale;
            if (nh < nv) io.store_f32_narrow(g0, ldg, n0 - dh, v, nh > 0 ? nh : 0, nv);
            if (nh <= 0) return;
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            uint32_t hw[4];
            r.piece(0, p, hw);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[8 * p + j] = v[8 * p + j] * sig_from_h<true>(WarpIO::unpack<FF>(hw, j) * hscale);
        }
        io.store<FF>(Aout, lda, n0, v, nh);
    }
};
template <class T>
struct EpiBwdS : EpiBase<EpiBwdS<T>> {            // operands: h_l, t_l, a_{l-1}
    using TF = Fw<T>;
    static constexpr int FF = Fmt16<TF>::value, FA = Fmt16<T>::value;
    static constexpr int kOps = 3;
    float hscale, inv_tscale, qscale; T* Pout; int64_t ldp;
    template <int W> __device__ __forceinline__ void run(int64_t, int, const float*, int) const {}
    __device__ __forceinline__ void chunk_smem(const WarpIO& io, int n0, float v[32], const msdf_tc::OpRow& r) const {
        const int nv = this->chunk_cols(n0);
        if (nv <= 0) return;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            uint32_t hw[4], tw[4], aw[4];
            r.piece(0, p, hw);
            r.piece(1, p, tw);
            r.piece(2, p, aw);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                v[8 * p + j] = EpiBwd<T>::template combine<true>(v[8 * p + j] * qscale, WarpIO::unpack<FF>(hw, j) * hscale, WarpIO::unpack<FA>(tw, j),
                                                                 WarpIO::unpack<FF>(aw, j), inv_tscale);
        }
        io.store<FA>(Pout, ldp, n0, v, nv);
    }
};
template <class T>
struct EpiBwdRelu : EpiBase<EpiBwdRelu<T>> {  // colour net dgrad: out[m,n] = acc * [Hin[m,n] > 0]
    using TF = Fw<T>;
    static constexpr int FF = Fmt16<TF>::value, FA = Fmt16<T>::value;
    const TF* Hin; int64_t ldh; T* out; int64_t ldo;
    template <int W> __device__ __forceinline__ void run(int64_t m, int n, const float* v, int nv) const {
        float h[W], o[W];
        load_row<W>(Hin + m * ldh + n, h, nv);
#pragma unroll
        for (int j = 0; j < W; ++j) o[j] = h[j] > 0.f ? v[j] : 0.f;
        store_row<W>(out + m * ldo + n, o, nv);
    }
    __device__ __forceinline__ void chunk(const WarpIO& io, int n0, float v[32], const uint4* q) const {
        const int nv = this->chunk_cols(n0);
        if (nv <= 0) return;
        float h[32];
        io.unstage<FF>(q, h);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = h[j] > 0.f ? v[j] : 0.f;
        io.store<FA>(out, ldo, n0, v, nv);
    }
    static constexpr int kPre = 1;
    __device__ __forceinline__ void prefetch(const WarpIO& io, int n0, uint4* q) const {
        if (n0 < this->N) io.prefetch(Hin, ldh, n0, q);
    }
};
// colour net layer-0 dgrad: route d(input) columns to the SDF net's adjoints
template <class T>
struct EpiColorIn : EpiBase<EpiColorIn<T>> {
    int n_off;                                     // column offset of this launch (the tcgen05 engine splits N > 256)
    int nc, fc, F, cc, cd;                         // column of the normal (-1 = none), of feat, of the code
    static constexpr int FA = Fmt16<T>::value;
    float* dn; T* Dout; int64_t ldo; int feat_col0; float* dcode; int64_t ldc;
    int rot, in0;                                  // GEMM column c is input column (c + rot) % in0 (bf16 mode: rot = fc)
    template <int W> __device__ __forceinline__ void run(int64_t m, int n, const float* v, int nv) const {
        const int c0 = n + n_off;
#pragma unroll
        for (int j = 0; j < W; ++j) {
            if (j >= nv) break;
            const int c = c0 + j;
            if (nc >= 0 && c >= nc && c < nc + 3) dn[3 * m + (c - nc)] = v[j];
            else if (c >= fc && c < fc + F) stf(Dout + m * ldo + feat_col0 + (c - fc), v[j]);
            else if (cd > 0 && c >= cc && c < cc + cd) dcode[m * ldc + (c - cc)] = v[j];
        }
    }
    // bf16 mode: the input is stored rotated so that feat starts at column 0 -> whole chunks of feat columns
    __device__ __forceinline__ void chunk(const WarpIO& io, int n0, float v[32], const uint4* q) const {
        const int nv = this->chunk_cols(n0);
        if (nv <= 0) return;
        const int g0c = n0 + n_off;                // first rotated column of the chunk
        if (g0c + 32 <= F) { io.store<FA>(Dout, ldo, feat_col0 + g0c, v, 32); return; }
        if (g0c < F) io.store<FA>(Dout, ldo, feat_col0 + g0c, v, F - g0c);
        if (cd > 0) {                              // code columns follow feat in the original order
            const int jlo = cc - rot - g0c;        // rotated column of code[0] is cc - rot
            io.store_f32(dcode, ldc, -jlo, v, jlo > 0 ? jlo : 0, jlo + cd < nv ? jlo + cd : nv, false);
        }
        if (nc >= 0 && io.valid()) {
            const int jn = in0 - rot + nc - g0c;   // nc < rot: the normal sits after the wrap-around
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j >= jn && j < jn + 3 && j < nv) dn[3 * io.row() + (j - jn)] = v[j];
        }
    }
};
// colour head on the tensor cores (bf16 mode): out[m, n] = act(acc + b[n]), n < N <= 4, fp32 output
struct EpiHead : EpiBase<EpiHead> {
    const float* bias; float* out; int64_t ldo; int act;
    template <int W> __device__ __forceinline__ void run(int64_t, int, const float*, int) const {}
    __device__ __forceinline__ void chunk(const WarpIO& io, int n0, float v[32], const uint4*) const {
        const int nv = this->chunk_cols(n0);
        if (nv <= 0) return;
        float b[32];
        io.colvec(n0, b);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float p = v[j] + b[j];
            v[j] = act == kActSigmoid ? 1.0f / (1.0f + __expf(-p)) : (act == kActRelu ? fmaxf(p, 0.f) : p);
        }
        io.store_f32(out, ldo, n0, v, 0, nv, false);
    }
    __device__ __forceinline__ const float* colvec() const { return bias; }
};

// dpre[m, n] = d_out[m, n] * act'(out[m, n]) for n < N, zero padded to `cols` columns (operand of the head's
// dgrad / wgrad GEMMs in bf16 mode)
template <class T>
__global__ void __launch_bounds__(128)
k_head_dpre(const float* __restrict__ d_out, const float* __restrict__ out, int64_t M, int N, int act, T* __restrict__ D, int64_t ldd, int cols) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    T* row = D + m * ldd;
    for (int j0 = 0; j0 < cols; j0 += 8) {
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = (j0 + u < N) ? d_out[m * N + j0 + u] * act_grad(act, out[m * N + j0 + u]) : 0.f;
        store_row<8>(row + j0, t, cols - j0 < 8 ? cols - j0 : 8);
    }
}

// Diffuse/specular split (network.py:441-453).  Forward: the diffuse colour is the tapped layer's (ReLU'd) columns
// [tap0, tap0 + 3); rgb6[m] = [diffuse + specular | specular].
template <class TF>
__global__ void k_spec_combine(const TF* __restrict__ Ctap, int64_t ldc, int tap0, const float* __restrict__ spec, int64_t M,
                               float* __restrict__ rgb6) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * kTapN) return;
    const int64_t m = i / kTapN; const int j = (int)(i - m * kTapN);
    const float sp = spec[i];
    rgb6[m * 6 + j] = ldf(Ctap + m * ldc + tap0 + j) + sp;
    rgb6[m * 6 + 3 + j] = sp;
}
// Backward, head side: the specular head sees dL/drgb + dL/drgb_spec; its output drives act'.
__global__ void k_spec_head_adjoint(const float* __restrict__ rgb6, const float* __restrict__ d_rgb6, int64_t M,
                                    float* __restrict__ spec, float* __restrict__ d_spec) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * kTapN) return;
    const int64_t m = i / kTapN; const int j = (int)(i - m * kTapN);
    spec[i] = rgb6[m * 6 + 3 + j];
    d_spec[i] = d_rgb6[m * 6 + j] + d_rgb6[m * 6 + 3 + j];
}
// Backward, tap side: adjoint of the tapped layer's pre-activation in the diffuse columns, P[m, tap0 + j] =
// dL/drgb[m, j] * [diffuse > 0]; columns [zero_lo, zero_hi) (the K padding of the tensor-core operand) are zeroed.
template <class TF, class T>
__global__ void k_spec_tail(const TF* __restrict__ Ctap, int64_t ldc, int tap0, const float* __restrict__ d_rgb6, int64_t M,
                            T* __restrict__ P, int64_t ldp, int zero_lo, int zero_hi) {
    const int w = kTapN + (zero_hi > zero_lo ? zero_hi - zero_lo : 0);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * w) return;
    const int64_t m = i / w; const int j = (int)(i - m * w);
    if (j < kTapN) stf(P + m * ldp + tap0 + j, ldf(Ctap + m * ldc + tap0 + j) > 0.f ? d_rgb6[m * 6 + j] : 0.f);
    else stf(P + m * ldp + zero_lo + (j - kTapN), 0.f);
}

// reverse-sweep start: a_{L-1} = e_0, so (a W_{L-1})[m,n] = W_{L-1}[0,n] for every point
template <class T>
__global__ void k_rev_init(const float* __restrict__ w_row, int64_t M, int N, EpiRev<T> epi) {
    constexpr int G = kIsBf16<T> ? 8 : 4;          // columns per thread: one 16-byte vector of T
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int ng = (N + G - 1) / G;
    if (i >= M * ng) return;
    const int64_t m = i / ng; const int n = (int)(i - m * ng) * G;
    float v[G];
#pragma unroll
    for (int j = 0; j < G; ++j) v[j] = (n + j < N) ? __ldg(w_row + n + j) : 0.f;
    epi.template run<G>(m, n, v, (N - n) < G ? (N - n) : G);
}

// colour-net input row (network.py:393-413): idr [x, PE(view), normal, feat, code], nerf [PE(view), feat, code].
// The feat columns are written by the SDF net's last layer; this kernel fills the rest (and the zero padding).
template <class T>
__global__ void k_color_input(const float* __restrict__ x, const float* __restrict__ view, const float* __restrict__ normal,
                              const float* __restrict__ code, int64_t M, int n_samples, int mode_idr, int pe_w, int F, int cd,
                              int code_per_ray, T* __restrict__ X, int64_t ldx, int pad_cols, int rot) {
    const int pre = (mode_idr ? 3 : 0) + pe_w + (mode_idr ? 3 : 0);
    const int cols = pre + cd + pad_cols;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * cols) return;
    const int64_t m = i / cols; int j = (int)(i - m * cols);
    const int64_t ray = m / n_samples;
    float v; int col;
    if (j >= pre + cd) {
        v = 0.f; col = pre + F + cd + (j - pre - cd);
    } else if (j >= pre) {
        v = code[(code_per_ray ? ray : 0) * cd + (j - pre)];
        col = pre + F + (j - pre);
    } else {
        col = j;
        if (mode_idr && j < 3) v = x[3 * m + j];
        else {
            if (mode_idr) j -= 3;
            if (j < pe_w) { const float d[3] = {view[3 * ray], view[3 * ray + 1], view[3 * ray + 2]}; v = pe_value(d, j); }
            else v = normal[3 * m + (j - pe_w)];
        }
    }
    const int in0 = pre + F + cd;
    if (rot > 0 && col < in0) { col -= rot; if (col < 0) col += in0; }
    stf(X + m * ldx + col, v);
}

__global__ void k_ray_points(const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ z,
                             int64_t n_rays, int n, float* __restrict__ pts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays * n * 3) return;
    const int64_t p = i / 3; const int k = (int)(i - 3 * p);
    const int64_t r = p / n;
    pts[i] = o[3 * r + k] + z[p] * d[3 * r + k];
}

// d_code[r, j] += sum_s dX_code[(r*n_samples + s), j]   (network.py:411-412: one code per ray, repeated per sample)
__global__ void k_code_grad(const float* __restrict__ dcode, int64_t n_rays, int n_samples, int cd, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays * cd) return;
    const int64_t r = i / cd; const int j = (int)(i - r * cd);
    float s = 0.f;
    for (int k = 0; k < n_samples; ++k) s += dcode[(r * n_samples + k) * cd + j];
    out[i] += s;
}

// 16-bit weight preparation for the tensor-core path.  Wk[r, k] = W[row(r), k] (zero padded to [rows_p, in_p]),
// Wt[k, r] = W[row(r), k] (zero padded to [in_p16, rows_p64]); row(r) = (r + perm) % out moves the first perm rows to
// the end (row_shift()).
__global__ void k_prep_weights(const float* __restrict__ W, int64_t ldw, int out, int in, int perm, int rot, Fw<bf16>* __restrict__ Wk,
                               int wk_rows, int wk_ld, Fw<bf16>* __restrict__ Wt, int wt_rows, int wt_ld, bf16* __restrict__ Wkb,
                               bf16* __restrict__ Wtb) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nk = (int64_t)wk_rows * wk_ld, nt = (int64_t)wt_rows * wt_ld;
    if (i < nk) {
        const int r = (int)(i / wk_ld), k = (int)(i - (int64_t)r * wk_ld);
        float v = 0.f;
        if (r < out && k < in) { const int src = (r + perm) % out; v = W[(int64_t)src * ldw + (k + rot) % in]; }
        stf(Wk + i, v); stf(Wkb + i, v);
    } else if (i < nk + nt) {
        const int64_t t = i - nk;
        const int k = (int)(t / wt_ld), r = (int)(t - (int64_t)k * wt_ld);
        float v = 0.f;
        if (r < out && k < in) { const int src = (r + perm) % out; v = W[(int64_t)src * ldw + (k + rot) % in]; }
        stf(Wt + t, v); stf(Wtb + t, v);
    }
}

inline unsigned nblk(int64_t n, int t = 256) { return (unsigned)msdf_div_up(n, t); }

#define RUN(expr) do { int rc_ = (expr); if (rc_) return rc_; } while (0)
#define LAUNCHED(name) do { MSDF_COUNT_LAUNCH(); MSDF_CHECK_LAUNCH(name); } while (0)

// ----------------------------------------------------------------------------------------------------------
// workspace layout for one chunk
// ----------------------------------------------------------------------------------------------------------
struct ColorGeom { int pe_w, nc, fc, cc, in0; };

ColorGeom color_geom(const msdf_color_desc* cd) {
    ColorGeom g;
    g.pe_w = cd->multires_view > 0 ? 3 + 6 * cd->multires_view : 3;
    if (cd->mode_idr) { g.nc = 3 + g.pe_w; g.fc = g.nc + 3; } else { g.nc = -1; g.fc = g.pe_w; }
    g.cc = g.fc + cd->feat_dim;
    g.in0 = g.cc + cd->code_dim;
    return g;
}

template <class T>
struct Bufs {
    using TF = Fw<T>;
    TF* H[MSDF_MAX_LAYERS];  // H[0] = encoded input (ld d0p); H[l] = input of layer l (ld ldh)
    TF* A[MSDF_MAX_LAYERS];  // a_l (forward format), later z_l / pbar_l in the adjoint format T, in place  (l < L-1)
    T *TG0, *T2[2], *Dout;     // T2: pbar ping-pong of the backward sweep
    T* TL[MSDF_MAX_LAYERS];    // TL[l] = t_l, the tangent entering layer l (TL[0] = TG0); kept for the backward sweep
    float *G0, *G0b, *BH0, *dydx, *hashf, *sdf_raw, *mask, *dn, *dn_color, *gradc, *sdfc, *dcode;   // G0b: the skip layer's share of d sdf / d h0
    float *spec32, *dspec32;   // spec variant: specular head output / its adjoint, [Mc, 3]
    TF* X; TF* C[MSDF_MAX_LAYERS]; T* dC[2]; T* Hd;   // Hd: colour head dpre, [Mc, 64] (bf16 mode)
    char* fusedW;              // packed parameters of the fused forward kernel (tensor-core mode)
    T* adj(int l) const { return reinterpret_cast<T*>(A[l]); }   // A[l] once it holds z_l / pbar_l
    int64_t d0p, ldh, ldo, ldx, ldc;   // leading dimensions
};

struct Carver {
    char* base; size_t off; bool dry;
    template <class U> U* take(int64_t rows, int64_t cols) {
        const size_t bytes = msdf_align((size_t)rows * (size_t)cols * sizeof(U), 1024);
        U* p = dry ? nullptr : reinterpret_cast<U*>(base + off);
        off += bytes;
        return p;
    }
};

struct Ctx {
    Net sn; const msdf_encoding_desc* enc; bool grid; int pe_w; float hash_chain;
    bool has_color; Net cn; const msdf_color_desc* cd; ColorGeom cg;
    float clamp_radius, sphere_scale;
    cudaStream_t st;
};

// mode: MSDF_MODE_*.  sn_w / cn_w (may be null) receive the bf16 weight-copy pointers.
// With split == true the activations the backward needs (the "persistent" group: H, A, G0, dy_dx, mask, colour-net
// rows) are carved from `saved` instead of the workspace and *saved_size receives their size: the forward of a
// training step writes them there and the backward reads them back instead of recomputing the chunk.
template <class T>
size_t carve(const Ctx& cx, int64_t Mc, int mode, void* ws, Bufs<T>* out, Net* sn_w, Net* cn_w, void* saved = nullptr,
             bool split = false, size_t* saved_size = nullptr) {
    const Net& sn = cx.sn;
    const msdf_encoding_desc* enc = cx.enc;
    Carver c{(char*)ws, 0, ws == nullptr};
    Carver ps{(char*)saved, 0, saved == nullptr};
    Carver& p = split ? ps : c;
    using TF = Fw<T>;
    Bufs<T> b{};
    if (kIsBf16<T>) {   // weight copies first (fixed size, independent of the chunk)
        Net* dst[2] = {sn_w, cn_w};
        const Net* src[2] = {&cx.sn, &cx.cn};
        for (int k = 0; k < (cx.has_color ? 2 : 1); ++k)
            for (int l = 0; l < src[k]->L; ++l) {
                Fw<bf16>* wk = c.take<Fw<bf16>>(round_up(src[k]->out[l], 16), round_up(src[k]->in[l], 64));
                Fw<bf16>* wt = c.take<Fw<bf16>>(round_up(src[k]->in[l], 16), round_up(src[k]->out[l], 64));
                bf16* wkb = c.take<bf16>(round_up(src[k]->out[l], 16), round_up(src[k]->in[l], 64));
                bf16* wtb = c.take<bf16>(round_up(src[k]->in[l], 16), round_up(src[k]->out[l], 64));
                if (dst[k]) { dst[k]->Wk[l] = wk; dst[k]->Wt[l] = wt; dst[k]->Wkb[l] = wkb; dst[k]->Wtb[l] = wtb; }
            }
    }
    if (kIsBf16<T> && mode != MSDF_MODE_SDF_ONLY)
        b.fusedW = c.take<char>(1, (int64_t)msdf_fused::fused_workspace_bytes(sn.L - 1));
    const bool grid_feats = enc->grid_feat_dim > 0 && enc->table != nullptr;
    b.d0p = padw<T>(sn.d0); b.ldh = padw<T>(sn.maxw);
    b.H[0] = p.take<TF>(Mc, b.d0p);
    b.sdf_raw = c.take<float>(Mc, 1);
    if (grid_feats && kIsBf16<T>) b.hashf = c.take<float>(Mc, enc->grid_feat_dim);
    if (mode == MSDF_MODE_SDF_ONLY) {
        TF* pp[2] = {c.take<TF>(Mc, b.ldh), c.take<TF>(Mc, b.ldh)};
        for (int l = 1; l < sn.L; ++l) b.H[l] = pp[(l - 1) & 1];
    } else {
        for (int l = 1; l < sn.L; ++l) b.H[l] = p.take<TF>(Mc, b.ldh);
        if (mode == MSDF_MODE_FORWARD && !split) {
            TF* pp[2] = {c.take<TF>(Mc, b.ldh), c.take<TF>(Mc, b.ldh)};
            for (int l = 0; l < sn.L - 1; ++l) b.A[l] = pp[l & 1];
        } else {
            for (int l = 0; l < sn.L - 1; ++l) b.A[l] = p.take<TF>(Mc, b.ldh);
        }
        b.G0 = p.take<float>(Mc, round_up(sn.d0, 4));
        // d sdf / d h0 arrives twice in a skip net (layer 0 and the skip layer's input half): two buffers summed by the decode
        // kernel instead of a read-modify-write in the layer-0 epilogue (274 us per 262144 rows, measured); nets that also
        // feed G0 to the hash scatter keep the accumulating form
        if (sn.skip > 0 && !grid_feats) b.G0b = c.take<float>(Mc, round_up(sn.d0, 4));
        if (grid_feats) b.dydx = p.take<float>(Mc, enc->n_levels * 3 * enc->level_dim);
        b.mask = p.take<float>(Mc, 1);
        if (mode == MSDF_MODE_BACKWARD) {
            b.TG0 = c.take<T>(Mc, b.d0p);
            if (grid_feats) b.BH0 = c.take<float>(Mc, round_up(sn.d0, 4));
            b.T2[0] = c.take<T>(Mc, b.ldh); b.T2[1] = c.take<T>(Mc, b.ldh);
            b.TL[0] = b.TG0;
            for (int l = 1; l < sn.L; ++l) b.TL[l] = c.take<T>(Mc, b.ldh);
            b.ldo = padw<T>(sn.out[sn.L - 1]);
            b.Dout = c.take<T>(Mc, b.ldo);
            b.dn = c.take<float>(Mc, 3); b.dn_color = c.take<float>(Mc, 3);
            b.gradc = c.take<float>(Mc, 3); b.sdfc = c.take<float>(Mc, 1);
        }
        if (cx.has_color) {
            b.ldx = padw<T>(cx.cg.in0); b.ldc = padw<T>(cx.cn.maxw);
            b.X = p.take<TF>(Mc, b.ldx);
            b.C[0] = b.X;
            for (int l = 1; l < cx.cn.L; ++l) b.C[l] = p.take<TF>(Mc, b.ldc);
            if (cx.cn.tap >= 0) { b.spec32 = c.take<float>(Mc, kTapN); b.dspec32 = c.take<float>(Mc, kTapN); }
            if (mode == MSDF_MODE_BACKWARD) {
                b.dC[0] = c.take<T>(Mc, b.ldc); b.dC[1] = c.take<T>(Mc, b.ldc);
                if (kIsBf16<T>) b.Hd = c.take<T>(Mc, 64);
                if (cx.cd->code_dim > 0) b.dcode = c.take<float>(Mc, cx.cd->code_dim);
            }
        }
    }
    if (out) *out = b;
    if (saved_size) *saved_size = ps.off;
    return c.off;
}

// Points per chunk in bf16 mode (knob MSDF_CHUNK_POINTS).  Default 1 060 864 = 148 SMs x 256 rows x 28: whole waves of the
// persistent kernels' 256-row tiles, and a quarter of the launches of the 262 144-point chunks of round 1 (measured on
// one box, 65 536-ray step: 262 144 -> 454 k rays/s, 795 648 -> 473 k, 1 060 864 -> 475 k, 2 121 728 -> 477 k; the
// per-call workspace grows with the chunk: 7.7 GB here).
constexpr int64_t kChunkPointsBf16 = 1060864;
inline int64_t chunk_cap_bf16() {
    static int64_t v = 0;
    if (!v) { const char* e = getenv("MSDF_CHUNK_POINTS"); v = e ? atoll(e) : kChunkPointsBf16; if (v < 1024) v = kChunkPointsBf16; }
    return v;
}

// Chunking of the saved-activation mode: a function of the problem only, so that forward and backward agree.
template <class T>
int64_t saved_chunk(const Ctx& cx, int64_t M, int n_samples) {
    const int64_t cap = kIsBf16<T> ? chunk_cap_bf16() : 65536;   // the caps of pick_chunk
    int64_t mc = M < cap ? M : cap;
    if (cx.has_color && mc < M) mc = mc / n_samples * n_samples;
    return mc;
}
template <class T>
size_t saved_stride(const Ctx& cx, int64_t chunk) {
    size_t sz = 0;
    carve<T>(cx, (chunk + 127) / 128 * 128, MSDF_MODE_BACKWARD, nullptr, nullptr, nullptr, nullptr, nullptr, true, &sz);
    return sz;
}

template <class T>
int64_t pick_chunk(const Ctx& cx, int64_t M, int mode, size_t ws_bytes) {
    int64_t cap = kIsBf16<T> ? chunk_cap_bf16() : 65536;
    if (mode == MSDF_MODE_SDF_ONLY) cap = 2 * chunk_cap_bf16();
    int64_t mc = M < cap ? M : cap;
    mc = (mc + 127) / 128 * 128;
    while (mc > 128 && carve<T>(cx, mc, mode, nullptr, nullptr, nullptr, nullptr) > ws_bytes) mc = (mc / 2 + 127) / 128 * 128;
    if (carve<T>(cx, mc, mode, nullptr, nullptr, nullptr, nullptr) > ws_bytes) return 0;
    return mc;
}

// ----------------------------------------------------------------------------------------------------------
// GEMM back ends
// ----------------------------------------------------------------------------------------------------------
// C = epi(A W_l^T) over output rows [r0, r0 + nrows) of W_l (in the bf16 copy's row order)
template <class T, class TA, class Epi>
int gemm_nt(const Ctx& c, const Net& n, int l, const TA* A, int64_t lda, int64_t Mc, int r0, int nrows, Epi epi, const char* what) {
    epi.N = nrows;
    if constexpr (kIsBf16<T>) {
        constexpr int fa = Fmt16<TA>::value;
        const uint16_t* wk = reinterpret_cast<const uint16_t*>(n.wk(l, fa));
        const int kp = round_up(n.in[l], 64);
        if (kp > 320 || nrows > 256) {
            // a wide first layer (e.g. 321 colour inputs with the per-image code): its weights only fit in shared
            // memory 128 rows at a time, so the layer runs as column blocks of the output; likewise a layer with more
            // than 256 outputs (the 259-wide tapped layer of the spec colour net) exceeds one accumulator
            const int blk = kp > 320 ? 128 : 256;
            if constexpr (std::is_same<Epi, EpiRelu<T>>::value || std::is_same<Epi, EpiFwdAct<T>>::value || std::is_same<Epi, EpiBias<T>>::value) {
                for (int q0 = 0; q0 < nrows; q0 += blk) {
                    const int nb = nrows - q0 < blk ? nrows - q0 : blk;
                    Epi e = epi;
                    e.N = nb; e.bias = epi.bias + q0; e.out = epi.out + q0;
                    RUN(msdf_tc::launch_gemm(A, fa, lda, Mc, kp, wk + (int64_t)(r0 + q0) * kp, fa, kp, round_up(nb, 16), e, c.st, what));
                }
                return MSDF_OK;
            } else {
                msdf_set_error("%s: more than 320 input / 256 output columns with this epilogue", what);
                return MSDF_ERR_UNSUPPORTED;
            }
        }
        return msdf_tc::launch_gemm(A, fa, lda, Mc, kp, wk + (int64_t)r0 * kp, fa, kp, round_up(nrows, 16), epi, c.st, what);
    } else {
        return msdf_gemm::launch<kNT>(A, lda, n.W[l] + (int64_t)r0 * n.ldw[l], n.ldw[l], Mc, nrows, n.in[l], 1, epi, c.st, what);
    }
}
// C = epi(A W_l): A [Mc, out_l] -> [Mc, in_l]
template <class T, class TA, class Epi>
int gemm_nn(const Ctx& c, const Net& n, int l, const TA* A, int64_t lda, int64_t Mc, Epi epi, const char* what) {
    epi.N = n.in[l];
    if constexpr (kIsBf16<T>) {
        constexpr int fa = Fmt16<TA>::value;
        const uint16_t* wt = reinterpret_cast<const uint16_t*>(n.wt(l, fa));
        const int kp = round_up(n.out[l], 64);
        const int np = round_up(n.in[l], 16);
        for (int c0 = 0; c0 < np; c0 += 256) {     // the accumulator holds at most 256 columns
            const int bn = np - c0 < 256 ? np - c0 : 256;
            Epi e = epi;
            e.N = n.in[l] - c0;
            if constexpr (std::is_same<Epi, EpiColorIn<T>>::value) e.n_off = c0;
            else if (c0 > 0) { msdf_set_error("%s: more than 256 output columns", what); return MSDF_ERR_UNSUPPORTED; }
            RUN(msdf_tc::launch_gemm(A, fa, lda, Mc, kp, wt + (int64_t)c0 * kp, fa, kp, bn, e, c.st, what));
        }
        return MSDF_OK;
    } else {
        return msdf_gemm::launch<kNN>(A, lda, n.W[l], n.ldw[l], Mc, n.in[l], n.out[l], 1, epi, c.st, what);
    }
}
// k_tc_stream launches (tensor-core mode): W_l for an "nt" product, W_l^T for an "nn" product, in A's format
int g_stream_enabled = [] { const char* e = getenv("MSDF_STREAM"); return e ? atoi(e) : 1; }();
template <class TA, class Epi, int kOps>
int stream_gemm(const Ctx& c, const Net& n, int l, bool transposed, const TA* A, int64_t lda, int64_t Mc, const void* const (&R)[kOps],
                const int (&rf)[kOps], const int64_t (&ldr)[kOps], Epi epi, const char* what, int ncols = 0) {
    constexpr int fa = Fmt16<TA>::value;
    const int rows = ncols > 0 ? ncols : (transposed ? n.in[l] : n.out[l]), kdim = transposed ? n.out[l] : n.in[l];
    const uint16_t* w = reinterpret_cast<const uint16_t*>(transposed ? n.wt(l, fa) : n.wk(l, fa));
    const int kp = round_up(kdim, 64);
    epi.N = rows;
    return msdf_tc::launch_stream(A, fa, lda, Mc, kp, w, kp, round_up(rows, 16), R, rf, ldr, epi, c.st, what);
}

template <class T>
int colsum(const Ctx& c, const T* X, int64_t ldx, const T* w, int64_t ws, int64_t Mc, int N, float* out);

// dW[rows, cols] += X^T Y, X [Mc, rows], Y [Mc, cols]; with db != nullptr also db[row(i)] += sum_m X[m, i] (the bias
// gradient: in bf16 mode it rides along in the weight-gradient kernel, which has X in shared memory anyway)
template <class T, class TX, class TY>
int wgrad(const Ctx& c, const TX* X, int64_t ldx, const TY* Y, int64_t ldy, int rows, int cols, int64_t Mc, float* dW, int64_t ldw,
          int perm_rows, int col_rot = 0, float* db = nullptr, int perm_shift = 1) {
    EpiAtomic e{};
    e.N = cols; e.C = dW; e.ldc = ldw; e.Mrows = rows; e.perm_rows = perm_rows; e.perm_shift = perm_shift; e.col_rot = col_rot;
    if constexpr (kIsBf16<T>) {
        return msdf_tc::launch_wgrad(X, Fmt16<TX>::value, ldx, round_up(rows, 64), Y, Fmt16<TY>::value, ldy, round_up(cols, 64), Mc, e, c.st,
                                     "weight gradient", db, rows, perm_rows, perm_shift);
    } else {
        if (db != nullptr) {
            if (perm_rows > 0) { msdf_set_error("weight gradient: permuted bias sums are a bf16-mode feature"); return MSDF_ERR_UNSUPPORTED; }
            RUN(colsum<T>(c, X, ldx, nullptr, 0, Mc, rows, db));
        }
        const int tiles = (int)(msdf_div_up(rows, msdf_gemm::BM) * msdf_div_up(cols, msdf_gemm::BN));
        int splits = (int)((2 * 148 + tiles - 1) / tiles);
        const int64_t max_splits = msdf_div_up(Mc, 512);
        if (splits > max_splits) splits = (int)max_splits;
        return msdf_gemm::launch<kTN>(X, ldx, Y, ldy, rows, cols, Mc, splits, e, c.st, "weight gradient");
    }
}
template <class T>
int colsum(const Ctx& c, const T* X, int64_t ldx, const T* w, int64_t ws, int64_t Mc, int N, float* out) {
    const int64_t rpb = 256;
    k_wcolsum<T><<<nblk(Mc, (int)rpb), 256, 0, c.st>>>(X, ldx, w, ws, Mc, N, rpb, out);
    LAUNCHED("column sum");
    return MSDF_OK;
}
int prep_weights(const Ctx& c, Net& n, int perm_last) {
    n.perm_last = perm_last;
    for (int l = 0; l < n.L; ++l) {
        const int perm = row_shift(n, l);
        const int wk_rows = round_up(n.out[l], 16), wk_ld = round_up(n.in[l], 64);
        const int wt_rows = round_up(n.in[l], 16), wt_ld = round_up(n.out[l], 64);
        const int64_t total = (int64_t)wk_rows * wk_ld + (int64_t)wt_rows * wt_ld;
        k_prep_weights<<<nblk(total), 256, 0, c.st>>>(n.W[l], n.ldw[l], n.out[l], n.in[l], perm, l == 0 ? n.rot0 : 0, n.Wk[l], wk_rows, wk_ld, n.Wt[l],
                                                     wt_rows, wt_ld, n.Wkb[l], n.Wtb[l]);
        LAUNCHED("weight prep");
    }
    return MSDF_OK;
}

// ----------------------------------------------------------------------------------------------------------
// sweeps over one chunk
// ----------------------------------------------------------------------------------------------------------
inline float in_scale(const Net& n, int l) { return l == n.skip ? kSqrt2 : 1.0f; }   // undoes the stored 1/sqrt2

template <class T>
int encode_chunk(const Ctx& c, const Bufs<T>& b, const float* x, int64_t Mc, bool want_dydx) {
    const int gw = c.enc->grid_feat_dim;
    if (c.grid) {
        // fp32 mode: features go straight into H0's columns; bf16 mode: through an fp32 staging buffer
        float* dst = kIsBf16<T> ? b.hashf : reinterpret_cast<float*>(b.H[0]) + c.pe_w;
        const int64_t ld = kIsBf16<T> ? gw : b.d0p;
        RUN(msdf_hash_forward_rows(x, c.enc->table, c.enc->offsets, dst, ld, Mc, c.enc->level_dim, c.enc->n_levels,
                                   c.enc->log2_per_level_scale, (uint32_t)c.enc->base_res, c.enc->divide_factor,
                                   want_dydx ? b.dydx : nullptr, c.st));
    }
    if (c.grid && !kIsBf16<T>) {   // the hash kernel already wrote its columns; PE only
        k_encode<Fw<T>><<<nblk(Mc * c.pe_w), 256, 0, c.st>>>(x, Mc, c.pe_w, 0, nullptr, b.H[0], b.d0p, c.pe_w);
    } else {                       // PE + staged hash features (or zero features) + zero padding
        const int cols = (int)b.d0p;
        k_encode_rows<Fw<T>><<<nblk(Mc, 128), 128, 0, c.st>>>(x, Mc, c.enc->multires, c.pe_w, gw, c.grid ? b.hashf : nullptr, b.H[0], b.d0p, cols);
    }
    LAUNCHED("encode");
    return MSDF_OK;
}

// forward sweep; feat (ld ldf_) may be null
template <class T>
int forward_sweep(const Ctx& c, const Bufs<T>& b, int64_t Mc, Fw<T>* feat, int64_t ldf_) {
    const Net& n = c.sn;
    for (int l = 0; l < n.L - 1; ++l) {
        const int64_t ldin = l == 0 ? b.d0p : b.ldh;
        if (l + 1 == n.skip) {
            k_skip_copy<Fw<T>><<<nblk(Mc * n.d0), 256, 0, c.st>>>(b.H[0], b.d0p, b.H[l + 1], b.ldh, Mc, n.d0, n.out[l], kInvSqrt2);
            LAUNCHED("skip copy");
        }
        EpiFwdAct<T> e{};
        e.bias = n.b[l]; e.out = b.H[l + 1]; e.ldo = b.ldh; e.oscale = l + 1 == n.skip ? kInvSqrt2 : 1.0f;
        RUN((gemm_nt<T>(c, n, l, b.H[l], ldin, Mc, 0, n.out[l], e, "sdf forward layer")));
    }
    const int l = n.L - 1;
    bool vec = false;
    if constexpr (kIsBf16<T>) vec = n.in[l] % 8 == 0 && n.in[l] <= 512;
    if constexpr (kIsBf16<T>) if (vec) {
        const int64_t blocks = msdf_div_up(Mc, 8 * 16);      // 8 warps per block, 16 rows (4 groups of 4) per warp
        k_rowdot1_vec<Fw<T>><<<(unsigned)(blocks > 0 ? blocks : 1), 256, 0, c.st>>>(b.H[l], b.ldh, n.W[l], n.b[l], Mc, n.in[l], b.sdf_raw);
    }
    if (!vec) {
        k_rowdot<1, Fw<T>><<<nblk(Mc, 8), 256, 0, c.st>>>(b.H[l], b.ldh, n.W[l], n.ldw[l], n.b[l], Mc, 1, n.in[l], kActNone, b.sdf_raw, 1);
    }
    LAUNCHED("sdf head");
    if (feat != nullptr && n.out[l] > 1) {
        EpiBias<T> e{};
        e.bias = n.b[l] + 1; e.out = feat; e.ldo = ldf_;
        // rows of the features: 1.. in W_l; 0.. in the permuted bf16 copy
        RUN((gemm_nt<T>(c, n, l, b.H[l], b.ldh, Mc, kIsBf16<T> ? 0 : 1, n.out[l] - 1, e, "feature head")));
    }
    return MSDF_OK;
}

template <class T>
EpiRev<T> make_rev(const Net& n, const Bufs<T>& b, int l) {
    EpiRev<T> e{};
    e.N = n.in[l];
    e.Hin = b.H[l]; e.ldh = l == 0 ? b.d0p : b.ldh; e.hscale = in_scale(n, l);
    e.Aout = l > 0 ? b.A[l - 1] : nullptr; e.lda = b.ldh;
    e.g0 = (l == n.skip && b.G0b != nullptr) ? b.G0b : b.G0; e.ldg = round_up(n.d0, 4);
    e.dh = l == n.skip ? n.in[l] - n.d0 : n.in[l];
    e.qscale = l == n.skip ? kInvSqrt2 : 1.0f;
    e.layer0 = l == 0; e.g0_accum = n.skip > 0 && b.G0b == nullptr;
    return e;
}

int g_chain_enabled = [] { const char* e = getenv("MSDF_CHAIN"); return e ? atoi(e) : 1; }();

// The whole reverse sweep of a chunk as ONE launch (tc_chain.cuh): a_l stays on chip between the layers.
template <class T>
bool rev_chain_applicable(const Net& n, const Bufs<T>& b) {
    if (!kIsBf16<T> || !g_chain_enabled || !g_stream_enabled) return false;
    if (n.L < 2 || n.L > msdf_tc::kChainMaxSteps) return false;
    if (n.skip > 0 && b.G0b == nullptr) return false;              // (layer 0 would have to accumulate into G0)
    for (int l = 0; l < n.L; ++l) {
        if (n.in[l] > 256) return false;
        if (l < n.L - 1 && n.out[l] > 256) return false;
    }
    return true;
}
template <class T>
int reverse_sweep_chain(const Ctx& c, const Bufs<T>& b, int64_t Mc) {
    using namespace msdf_tc;
    const Net& n = c.sn;
    ChainMaps maps{};
    ChainPlan plan{};
    plan.ldg = round_up(n.d0, 4);
    double flops = 0.0, bytes = 0.0;
    int s = 0;
    {
        const int l = n.L - 1;
        ChainStep& st = plan.st[s];
        st.BN = 0; st.KB = 0; st.N = n.in[l]; st.dh = n.in[l];
        st.nb = (round_up(n.in[l], 32) / 32 + 1) / 2; st.ob = st.nb;
        st.hscale = in_scale(n, l); st.qscale = 1.0f; st.g0 = nullptr;
        maps.w[s] = CUtensorMap{};
        RUN(make_map(&maps.op[s], b.H[l], kF16, Mc, st.nb * 64, b.ldh, BM, "reverse sweep (h)"));
        RUN(make_map(&maps.out[s], b.A[l - 1], kF16, Mc, st.ob * 64, b.ldh, BM, "reverse sweep (a)"));
        bytes += (double)Mc * 2.0 * 64.0 * (st.nb + st.ob);
        ++s;
    }
    for (int l = n.L - 2; l >= 0; --l, ++s) {
        ChainStep& st = plan.st[s];
        const int dh = l == 0 ? 0 : (l == n.skip ? n.in[l] - n.d0 : n.in[l]);
        st.BN = round_up(n.in[l], 16); st.KB = round_up(n.out[l], 64) / 64; st.N = n.in[l]; st.dh = dh;
        st.nb = l > 0 ? (round_up(st.BN, 32) / 32 + 1) / 2 : 0;
        st.ob = l > 0 ? round_up(dh, 64) / 64 : 0;
        st.hscale = in_scale(n, l); st.qscale = l == n.skip ? kInvSqrt2 : 1.0f;
        st.g0 = (l == n.skip && b.G0b != nullptr) ? b.G0b : b.G0;
        const int kp = st.KB * 64;
        RUN(make_map(&maps.w[s], n.wt(l, kF16), kF16, st.BN, kp, kp, st.BN, "reverse sweep (W^T)", kChainBK));
        if (st.nb > 0) RUN(make_map(&maps.op[s], b.H[l], kF16, Mc, st.nb * 64, l == 0 ? b.d0p : b.ldh, BM, "reverse sweep (h)"));
        else maps.op[s] = maps.op[0];
        if (st.ob > 0) RUN(make_map(&maps.out[s], b.A[l - 1], kF16, Mc, st.ob * 64, b.ldh, BM, "reverse sweep (a)"));
        else maps.out[s] = maps.out[0];
        flops += 2.0 * (double)Mc * st.BN * kp;
        bytes += (double)Mc * 2.0 * 64.0 * (st.nb + st.ob) + (double)Mc * 4.0 * (st.N - dh);
    }
    plan.S = s;
    RevChainPolicy pol{n.W[n.L - 1]};
    return launch_chain(maps, plan, Mc, pol, flops, bytes, c.st, "sdf reverse sweep (chained)");
}

template <class T>
int reverse_sweep(const Ctx& c, const Bufs<T>& b, int64_t Mc) {
    const Net& n = c.sn;
    if (rev_chain_applicable<T>(n, b)) return reverse_sweep_chain<T>(c, b, Mc);
    {
        const int l = n.L - 1;
        const int ng = (n.in[l] + (kIsBf16<T> ? 7 : 3)) / (kIsBf16<T> ? 8 : 4);
        k_rev_init<T><<<nblk(Mc * ng), 256, 0, c.st>>>(n.W[l], Mc, n.in[l], make_rev<T>(n, b, l));
        LAUNCHED("reverse init");
    }
    for (int l = n.L - 2; l >= 0; --l) {
        if (kIsBf16<T> && n.out[l] % 64 != 0) {   // K padding of the operand must be finite (zero)
            const int w = round_up(n.out[l], 64) - n.out[l];
            k_zero_cols<Fw<T>><<<nblk(Mc * w), 256, 0, c.st>>>(b.A[l], b.ldh, Mc, n.out[l], w);
            LAUNCHED("zero operand padding");
        }
        if constexpr (kIsBf16<T>) {
            const bool plain = l > 0 && l != n.skip;
            if (g_stream_enabled && (plain || b.G0b != nullptr || n.skip <= 0) && n.in[l] <= 256 && round_up(n.out[l], 64) <= 320) {
                const void* const R[1] = {b.H[l]}; const int rf[1] = {Fmt16<Fw<T>>::value}; const int64_t ldr[1] = {l == 0 ? b.d0p : b.ldh};
                if (plain) {
                    EpiRevS<T> es{};
                    es.hscale = in_scale(n, l); es.Aout = b.A[l - 1]; es.lda = b.ldh;
                    RUN((stream_gemm<Fw<T>, EpiRevS<T>, 1>(c, n, l, true, b.A[l], b.ldh, Mc, R, rf, ldr, es, "sdf reverse layer")));
                } else {
                    EpiRevS<T, true> es{};
                    es.hscale = in_scale(n, l); es.Aout = l > 0 ? b.A[l - 1] : nullptr; es.lda = b.ldh;
                    es.dh = l == 0 ? 0 : n.in[l] - n.d0;
                    es.qscale = l == n.skip ? kInvSqrt2 : 1.0f;
                    es.g0 = (l == n.skip && b.G0b != nullptr) ? b.G0b : b.G0; es.ldg = round_up(n.d0, 4);
                    RUN((stream_gemm<Fw<T>, EpiRevS<T, true>, 1>(c, n, l, true, b.A[l], b.ldh, Mc, R, rf, ldr, es, "sdf reverse layer (h0 columns)")));
                }
                continue;
            }
        }
        RUN((gemm_nn<T>(c, n, l, b.A[l], b.ldh, Mc, make_rev<T>(n, b, l), "sdf reverse layer")));
    }
    return MSDF_OK;
}

template <class T>
int decode_chunk(const Ctx& c, const Bufs<T>& b, const float* x, int64_t Mc, bool with_grad, float* sdf, float* grad, float* mask) {
    k_decode<<<nblk(Mc, 128), 128, 0, c.st>>>(x, with_grad ? b.G0 : nullptr, with_grad ? b.G0b : nullptr, round_up(c.sn.d0, 4), Mc, c.pe_w,
                                             c.grid ? c.enc->grid_feat_dim : 0, c.enc->n_levels, c.enc->level_dim, b.dydx, c.hash_chain,
                                             b.sdf_raw, c.clamp_radius, c.sphere_scale, sdf, grad, mask);
    LAUNCHED("decode");
    return MSDF_OK;
}

template <class T>
int color_forward(const Ctx& c, const Bufs<T>& b, const float* x, int64_t Mc, const float* view, int n_samples,
                  const float* code, const float* normal, float* rgb) {
    const Net& n = c.cn;
    const ColorGeom& g = c.cg;
    const int pad = (int)b.ldx - g.in0;
    const int cols = g.in0 - c.cd->feat_dim + pad;
    // chunks start on a ray boundary; view / code pointers are already offset to the chunk's first ray
    if (kIsBf16<T> && n.rot0 == g.fc && c.cd->feat_dim % 8 == 0) {
        // rotated row: [feat | code | x | PE(view) | normal | padding]: everything after feat is one aligned range
        k_color_input_rows<Fw<T>><<<nblk(Mc, 128), 128, 0, c.st>>>(x, view, normal, code, Mc, n_samples, c.cd->mode_idr, c.cd->multires_view,
                                                              g.pe_w, c.cd->code_dim, c.cd->code_per_ray, b.X, b.ldx, c.cd->feat_dim,
                                                              (int)b.ldx);
    } else {
        k_color_input<Fw<T>><<<nblk(Mc * cols), 256, 0, c.st>>>(x, view, normal, code, Mc, n_samples, c.cd->mode_idr, g.pe_w, c.cd->feat_dim,
                                                           c.cd->code_dim, c.cd->code_per_ray, b.X, b.ldx, pad, n.rot0);
    }
    LAUNCHED("colour input");
    // spec variant: the tapped layer's output is stored [specular features | diffuse] in tensor-core mode (rows moved
    // by the weight copy, so the next layer's operand starts 16-byte aligned) and in the reference's order
    // [diffuse | features] in fp32 mode; tap_in = first column of the next layer's input, tap0 = first diffuse column
    const bool spec = n.tap >= 0;
    const int tap_in = spec && !kIsBf16<T> ? kTapN : 0;
    const int tap0 = spec ? (kIsBf16<T> ? n.out[n.tap] - kTapN : 0) : 0;
    for (int l = 0; l < n.L - 1; ++l) {
        EpiRelu<T> e{};
        e.bias = n.b[l]; e.out = b.C[l + 1]; e.ldo = b.ldc;
        const Fw<T>* in = b.C[l] + (spec && l == n.tap + 1 ? tap_in : 0);
        const int64_t ldin = l == 0 ? b.ldx : b.ldc;
        if (kIsBf16<T> && spec && l == n.tap) {
            // rows of the 16-bit copy: [features (out - 3) | diffuse (3)]; the bias vector is in the reference's order
            const int nf = n.out[l] - kTapN;
            e.bias = n.b[l] + kTapN;
            RUN((gemm_nt<T>(c, n, l, in, ldin, Mc, 0, nf, e, "colour layer (specular features)")));
            e.bias = n.b[l]; e.out = b.C[l + 1] + nf;
            RUN((gemm_nt<T>(c, n, l, in, ldin, Mc, nf, kTapN, e, "colour layer (diffuse)")));
            continue;
        }
        RUN((gemm_nt<T>(c, n, l, in, ldin, Mc, 0, n.out[l], e, "colour layer")));
    }
    const int l = n.L - 1;
    MSDF_CHECK_ARG(n.out[l] <= 4, "colour net: d_out=%d > 4 unsupported", n.out[l]);
    MSDF_CHECK_ARG(!spec || n.out[l] == kTapN, "colour net: the diffuse/specular split needs d_out = 3");
    if (rgb == nullptr) return MSDF_OK;   // backward recompute: the saved rgb drives act', the head is not needed
    const int act = (spec || c.cd->final_act != 0) ? kActRelu : kActSigmoid;
    float* head_out = spec ? b.spec32 : rgb;
    if constexpr (kIsBf16<T>) {
        EpiHead e{};
        e.bias = n.b[l]; e.out = head_out; e.ldo = n.out[l]; e.act = act;
        RUN((gemm_nt<T>(c, n, l, b.C[l], b.ldc, Mc, 0, n.out[l], e, "colour head")));
    } else {
        k_rowdot<4, Fw<T>><<<nblk(Mc, 8), 256, 0, c.st>>>(b.C[l], b.ldc, n.W[l], n.ldw[l], n.b[l], Mc, n.out[l], n.in[l], act, head_out, n.out[l]);
        LAUNCHED("colour head");
    }
    if (spec) {
        k_spec_combine<Fw<T>><<<nblk(Mc * kTapN), 256, 0, c.st>>>(b.C[n.tap + 1], b.ldc, tap0, b.spec32, Mc, rgb);
        LAUNCHED("diffuse + specular");
    }
    return MSDF_OK;
}

template <class T>
int color_backward(const Ctx& c, const Bufs<T>& b, int64_t Mc, const float* rgb, const float* d_rgb, const msdf_mlp_grads* gr) {
    const Net& n = c.cn;
    const ColorGeom& g = c.cg;
    int l = n.L - 1;
    const bool spec = n.tap >= 0;
    const int tap_in = spec && !kIsBf16<T> ? kTapN : 0;
    const int tap0 = spec ? (kIsBf16<T> ? n.out[n.tap] - kTapN : 0) : 0;
    const int act = (spec || c.cd->final_act != 0) ? kActRelu : kActSigmoid;
    const float* d_rgb6 = d_rgb;
    if (spec) {   // rgb / d_rgb are [Mc, 6] = [rgb | rgb_spec]: the head is the specular branch
        k_spec_head_adjoint<<<nblk(Mc * kTapN), 256, 0, c.st>>>(rgb, d_rgb, Mc, b.spec32, b.dspec32);
        LAUNCHED("specular head adjoint");
        rgb = b.spec32; d_rgb = b.dspec32;
    }
    T* P = b.dC[0];
    if constexpr (kIsBf16<T>) {
        // the 3-wide head as zero-padded tensor-core GEMMs: dpre [Mc, 64] is the operand of its wgrad and dgrad
        k_head_dpre<T><<<nblk(Mc, 128), 128, 0, c.st>>>(d_rgb, rgb, Mc, n.out[l], act, b.Hd, 64, 64);
        LAUNCHED("colour head dpre");
        RUN(wgrad<T>(c, b.Hd, 64, b.C[l], b.ldc, n.out[l], n.in[l], Mc, gr->dW[l], n.ldw[l], 0, 0, gr->db[l]));
        EpiBwdRelu<T> e{};
        e.Hin = b.C[l]; e.ldh = b.ldc; e.out = P; e.ldo = b.ldc;
        RUN((gemm_nn<T>(c, n, l, b.Hd, 64, Mc, e, "colour head dgrad")));
    } else {
        k_rowdot_wgrad<4, T><<<nblk(Mc, 512), 256, 0, c.st>>>(d_rgb, rgb, n.out[l], act, b.C[l], b.ldc, Mc, n.out[l], n.in[l], 512,
                                                             gr->dW[l], n.ldw[l], gr->db[l]);
        LAUNCHED("colour head wgrad");
        k_rowdot_dgrad<4, T><<<nblk(Mc * n.in[l]), 256, 0, c.st>>>(d_rgb, rgb, n.out[l], act, n.W[l], n.ldw[l], b.C[l], b.ldc, Mc,
                                                                  n.out[l], n.in[l], P, b.ldc);
        LAUNCHED("colour head dgrad");
    }
    int pp = 0;
    for (l = n.L - 2; l >= 0; --l) {
        const int64_t ldin = l == 0 ? b.ldx : b.ldc;
        const int in_off = spec && l == n.tap + 1 ? tap_in : 0;          // this layer reads past the diffuse columns
        const int shift = kIsBf16<T> ? row_shift(n, l) : 0;              // rows of the tapped layer are moved in the 16-bit copies
        RUN(wgrad<T>(c, P, b.ldc, b.C[l] + in_off, ldin, n.out[l], n.in[l], Mc, gr->dW[l], n.ldw[l], shift > 0 ? n.out[l] : 0,
                     l == 0 ? n.rot0 : 0, gr->db[l], shift > 0 ? shift : 1));
        if (l > 0) {
            T* Pn = b.dC[pp ^ 1];
            EpiBwdRelu<T> e{};
            e.Hin = b.C[l] + in_off; e.ldh = b.ldc; e.out = Pn + in_off; e.ldo = b.ldc;
            RUN((gemm_nn<T>(c, n, l, P, b.ldc, Mc, e, "colour dgrad")));
            if (spec && l == n.tap + 1) {
                // Pn is now the adjoint of the tapped layer's pre-activation in the feature columns: add the diffuse
                // columns and, in tensor-core mode, zero the K padding the next GEMMs read
                const int no = n.out[n.tap];
                k_spec_tail<Fw<T>, T><<<nblk(Mc * (kTapN + 64)), 256, 0, c.st>>>(b.C[l], b.ldc, tap0, d_rgb6, Mc, Pn, b.ldc, no,
                                                                             kIsBf16<T> ? round_up(no, 64) : no);
                LAUNCHED("diffuse adjoint");
            }
            P = Pn; pp ^= 1;
        } else {
            EpiColorIn<T> e{};
            e.n_off = 0; e.nc = g.nc; e.fc = g.fc; e.F = c.cd->feat_dim; e.cc = g.cc; e.cd = c.cd->code_dim;
            e.dn = b.dn_color; e.Dout = b.Dout; e.ldo = b.ldo; e.feat_col0 = kIsBf16<T> ? 0 : 1;
            e.dcode = b.dcode; e.ldc = c.cd->code_dim; e.rot = n.rot0; e.in0 = g.in0;
            RUN((gemm_nn<T>(c, n, l, P, b.ldc, Mc, e, "colour input dgrad")));
        }
    }
    return MSDF_OK;
}

template <class T>
int sdf_backward(const Ctx& c, const Bufs<T>& b, const float* x, int64_t Mc, const msdf_mlp_grads* gr, float* grad_table) {
    const Net& n = c.sn;
    const int L1 = n.L - 1;
    // ---- tangent sweep (adjoint of the reverse sweep): t_{l+1} = (W_l t_l) sigma_l, every t_l kept
    for (int l = 0; l < L1; ++l) {
        T* Tin = b.TL[l]; const int64_t ldt = l == 0 ? b.d0p : b.ldh;
        if (l == n.skip) {
            k_skip_copy<T><<<nblk(Mc * n.d0), 256, 0, c.st>>>(b.TG0, b.d0p, Tin, ldt, Mc, n.d0, n.in[l] - n.d0, kInvSqrt2);
            LAUNCHED("tangent skip copy");
        }
        bool streamed = false;
        if constexpr (kIsBf16<T>) {
            if (g_stream_enabled && round_up(n.in[l], 64) <= 320 && n.out[l] <= 256) {
                EpiTanS<T> es{};
                es.hscale = in_scale(n, l + 1); es.Tout = b.TL[l + 1]; es.ldt = b.ldh; es.tscale = l + 1 == n.skip ? kInvSqrt2 : 1.0f;
                const void* const R[1] = {b.H[l + 1]}; const int rf[1] = {Fmt16<Fw<T>>::value}; const int64_t ldr[1] = {b.ldh};
                RUN((stream_gemm<T, EpiTanS<T>, 1>(c, n, l, false, Tin, ldt, Mc, R, rf, ldr, es, "sdf tangent layer")));
                streamed = true;
            }
        }
        if (!streamed) {
            EpiTan<T> e{};
            e.Hn = b.H[l + 1]; e.ldh = b.ldh; e.hscale = in_scale(n, l + 1);
            e.Tout = b.TL[l + 1]; e.ldt = b.ldh; e.tscale = l + 1 == n.skip ? kInvSqrt2 : 1.0f;
            RUN((gemm_nt<T>(c, n, l, Tin, ldt, Mc, 0, n.out[l], e, "sdf tangent layer")));
        }
    }
    const int out_last = n.out[L1];
    {
        T* Tin = b.TL[L1]; const int64_t ldt = b.ldh;
        if (L1 == n.skip) {
            k_skip_copy<T><<<nblk(Mc * n.d0), 256, 0, c.st>>>(b.TG0, b.d0p, Tin, ldt, Mc, n.d0, n.in[L1] - n.d0, kInvSqrt2);
            LAUNCHED("tangent skip copy");
        }
        RUN(colsum<T>(c, Tin, ldt, nullptr, 0, Mc, n.in[L1], gr->dW[L1]));                          // a_{L-1} = e_0 -> sdf row
        // ---- last layer of the backward sweep: pbar_{L-1} = Dout
        if constexpr (kIsBf16<T>) {
            // Dout columns are [features..., sdf]: one weight-gradient GEMM with the row permutation folded in.  (Tried: the
            // sdf row as a weighted column sum of h next to a one-tile GEMM -- 240 + 160 us against 463 us per 1 060 864
            // rows, but the column-sum kernel re-reads h: no net gain.)
            RUN(wgrad<T>(c, b.Dout, b.ldo, b.H[L1], b.ldh, out_last, n.in[L1], Mc, gr->dW[L1], n.ldw[L1], out_last, 0, gr->db[L1]));
        } else {
            RUN(colsum<T>(c, b.H[L1], b.ldh, b.Dout, b.ldo, Mc, n.in[L1], gr->dW[L1]));             // sdf row
            if (out_last > 1)
                RUN(wgrad<T>(c, b.Dout + 1, b.ldo, b.H[L1], b.ldh, out_last - 1, n.in[L1], Mc, gr->dW[L1] + n.ldw[L1], n.ldw[L1], 0));
            RUN(colsum<T>(c, b.Dout, b.ldo, nullptr, 0, Mc, out_last, gr->db[L1]));
        }
    }
    // ---- backward sweep with both weight-gradient products of every layer: dW_l += pbar_l^T u_l + a_l^T t_l
    const T* P = b.Dout; int64_t ldp = b.ldo;
    for (int l = L1; l >= 0; --l) {
        const int64_t ldin = l == 0 ? b.d0p : b.ldh;
        if (l < L1) {
            bool paired = false;
            if constexpr (kIsBf16<T>) {
                if (g_stream_enabled) {     // both products of the layer in one launch (one prologue, one flush of dW_l)
                    EpiAtomic e{};
                    e.N = n.in[l]; e.C = gr->dW[l]; e.ldc = n.ldw[l]; e.Mrows = n.out[l]; e.perm_rows = 0; e.perm_shift = 1; e.col_rot = 0;
                    RUN(msdf_tc::launch_wgrad(P, Fmt16<T>::value, ldp, round_up(n.out[l], 64), b.H[l], Fmt16<Fw<T>>::value, ldin,
                                              round_up(n.in[l], 64), Mc, e, c.st, "weight gradient (both products)", gr->db[l], n.out[l], 0, 1,
                                              b.A[l], Fmt16<Fw<T>>::value, b.ldh, b.TL[l], Fmt16<T>::value, ldin));
                    paired = true;
                }
            }
            if (!paired) {
                RUN(wgrad<T>(c, P, ldp, b.H[l], ldin, n.out[l], n.in[l], Mc, gr->dW[l], n.ldw[l], 0, 0, gr->db[l]));
                RUN(wgrad<T>(c, b.A[l], b.ldh, b.TL[l], ldin, n.out[l], n.in[l], Mc, gr->dW[l], n.ldw[l], 0));
            }
        }
        if (l == 0 && !(c.grid && grad_table)) break;
        if constexpr (kIsBf16<T>) {
            // (the skip layer too when nothing reads the adjoint of its h_0 half: only the hidden columns are computed)
            const bool skip_ok = l != n.skip || !(c.grid && grad_table);
            if (g_stream_enabled && l > 0 && skip_ok && n.in[l] <= 256 && round_up(n.out[l], 64) <= 320) {
                T* Pout = b.T2[l & 1];
                EpiBwdS<T> es{};
                es.hscale = in_scale(n, l); es.inv_tscale = l == n.skip ? kSqrt2 : 1.0f; es.qscale = l == n.skip ? kInvSqrt2 : 1.0f;
                es.Pout = Pout; es.ldp = b.ldh;
                const void* const R[3] = {b.H[l], b.TL[l], b.A[l - 1]};
                const int rf[3] = {Fmt16<Fw<T>>::value, Fmt16<T>::value, Fmt16<Fw<T>>::value};
                const int64_t ldr[3] = {ldin, ldin, b.ldh};
                const int ncols = l == n.skip ? n.in[l] - n.d0 : n.in[l];
                RUN((stream_gemm<T, EpiBwdS<T>, 3>(c, n, l, true, P, ldp, Mc, R, rf, ldr, es, "sdf backward layer", ncols)));
                if (n.out[l - 1] % 64 != 0) {   // K padding of the next launch's operand must be finite (zero)
                    const int w = round_up(n.out[l - 1], 64) - n.out[l - 1];
                    k_zero_cols<T><<<nblk(Mc * w), 256, 0, c.st>>>(Pout, b.ldh, Mc, n.out[l - 1], w);
                    LAUNCHED("zero operand padding");
                }
                P = Pout; ldp = b.ldh;
                continue;
            }
        }
        EpiBwd<T> e{};
        e.Hin = b.H[l]; e.ldh = ldin; e.hscale = in_scale(n, l);
        e.Tin = b.TL[l]; e.ldt = ldin; e.inv_tscale = l == n.skip ? kSqrt2 : 1.0f;
        e.Ain = l > 0 ? b.A[l - 1] : nullptr; e.lda = b.ldh;
        T* Pout = b.T2[l & 1];
        e.Pout = l > 0 ? Pout : nullptr; e.ldp = b.ldh;
        e.bh0 = (c.grid && grad_table) ? b.BH0 : nullptr; e.ldb = round_up(n.d0, 4);
        e.dh = l == n.skip ? n.in[l] - n.d0 : n.in[l];
        e.qscale = l == n.skip ? kInvSqrt2 : 1.0f;
        e.layer0 = l == 0; e.bh0_accum = n.skip > 0;
        RUN((gemm_nn<T>(c, n, l, P, ldp, Mc, e, "sdf backward layer")));
        if (kIsBf16<T> && l > 0 && n.out[l - 1] % 64 != 0) {   // K padding of the next launch's operand must be finite (zero)
            const int w = round_up(n.out[l - 1], 64) - n.out[l - 1];
            k_zero_cols<T><<<nblk(Mc * w), 256, 0, c.st>>>(Pout, b.ldh, Mc, n.out[l - 1], w);
            LAUNCHED("zero operand padding");
        }
        if (l > 0) { P = Pout; ldp = b.ldh; }
    }
    if (c.grid && grad_table) {
        const int64_t ldg = round_up(n.d0, 4);
        RUN(msdf_hash_scatter_rows(x, c.enc->offsets, Mc, c.enc->level_dim, c.enc->n_levels, c.enc->log2_per_level_scale,
                                   (uint32_t)c.enc->base_res, c.enc->divide_factor, b.BH0 + c.pe_w, ldg, b.G0 + c.pe_w, ldg,
                                   b.dn, c.hash_chain, grad_table, c.st));
    }
    return MSDF_OK;
}

int make_ctx(Ctx& c, const msdf_mlp_desc* sdf_net, const msdf_encoding_desc* enc, const msdf_mlp_desc* color_net,
             const msdf_color_desc* cd, float clamp_radius, float sphere_scale, void* stream, unsigned flags, const char* who) {
    MSDF_CHECK_ARG(enc != nullptr, "%s: null encoding descriptor", who);
    RUN(make_net(sdf_net, c.sn, who));
    c.enc = enc;
    c.pe_w = enc->multires > 0 ? 3 + 6 * enc->multires : 3;
    MSDF_CHECK_ARG(enc->multires <= 16, "%s: multires=%d too large", who, enc->multires);
    MSDF_CHECK_ARG(c.pe_w + enc->grid_feat_dim == c.sn.d0, "%s: d0=%d but the encoding yields %d", who, c.sn.d0,
                   c.pe_w + enc->grid_feat_dim);
    c.grid = enc->grid_feat_dim > 0 && enc->table != nullptr;
    if (enc->grid_feat_dim > 0) {
        MSDF_CHECK_ARG(enc->n_levels * enc->level_dim == enc->grid_feat_dim, "%s: grid_feat_dim != n_levels*level_dim", who);
        MSDF_CHECK_ARG(!c.grid || enc->offsets, "%s: hash offsets missing", who);
        MSDF_CHECK_ARG(enc->divide_factor > 0.f, "%s: divide_factor must be > 0", who);
    }
    c.hash_chain = c.grid ? 1.0f / (2.0f * enc->divide_factor) : 0.f;
    c.has_color = color_net != nullptr;
    c.cd = cd;
    if (c.has_color) {
        MSDF_CHECK_ARG(cd != nullptr, "%s: colour descriptor missing", who);
        RUN(make_net(color_net, c.cn, who, cd->spec ? color_net->n_layers - 3 : -1));
        c.cg = color_geom(cd);
        MSDF_CHECK_ARG(c.cn.skip < 0, "%s: colour net has no skip connection", who);
        MSDF_CHECK_ARG(c.cg.in0 == c.cn.d0, "%s: colour net d0=%d but inputs total %d", who, c.cn.d0, c.cg.in0);
        MSDF_CHECK_ARG(cd->feat_dim == c.sn.out[c.sn.L - 1] - 1, "%s: feat_dim mismatch", who);
        MSDF_CHECK_ARG(c.cd->multires_view <= 16, "%s: multires_view too large", who);
    }
    if (flags & MSDF_FLAG_TENSOR_BF16) {
        RUN(check_tc_net(c.sn, who));
        if (c.has_color) RUN(check_tc_net(c.cn, who));
    }
    c.clamp_radius = clamp_radius; c.sphere_scale = sphere_scale;
    c.st = (cudaStream_t)stream;
    return MSDF_OK;
}

// ----------------------------------------------------------------------------------------------------------
// sdf-only queries of the tensor-core mode: the whole network in one persistent kernel (fused_mlp.cuh)
// ----------------------------------------------------------------------------------------------------------
int g_fused_enabled = [] { const char* e = getenv("MSDF_FUSED"); return e ? atoi(e) : 1; }();   // msdf_set_fused() / MSDF_FUSED=0: A/B switch (fused kernel vs per-layer sweep)

bool fused_applicable(const Ctx& c) {
    return g_fused_enabled && msdf_fused::fused_supported(c.sn.L, c.sn.in, c.sn.out, c.sn.skip, c.sn.d0, c.pe_w);
}

// packed parameters + weight tensor map + plan of one fused launch series (scaled domain: sdf-only; plain: training forward)
struct FusedPrep { msdf_fused::Plan P; CUtensorMap mW; };

int fused_prepare(const Ctx& c, char* ws, bool train, FusedPrep& fp, const char* who) {
    namespace mf = msdf_fused;
    const Net& n = c.sn;
    const int L = n.L - 1, LT = L + (train ? 1 : 0);
    __half* Wp = reinterpret_cast<__half*>(ws);
    float* bp = reinterpret_cast<float*>(ws + (size_t)LT * 131072);
    float* wl = bp + (size_t)LT * 256;
    mf::PackArgs pa{};
    mf::Plan& P = fp.P;
    P = mf::Plan{};
    pa.L = L; P.L = L;
    for (int l = 0; l < L; ++l) {
        pa.W[l] = n.W[l]; pa.b[l] = n.b[l]; pa.out[l] = n.out[l]; pa.in[l] = n.in[l]; pa.ldw[l] = n.ldw[l];
        P.kb[l] = round_up(n.in[l], 64) / 64;
        P.oscale[l] = l + 1 == n.skip ? kInvSqrt2 : 1.0f;
    }
    P.kb[L] = 4;
    pa.w_last = n.W[L]; pa.b_last = n.b[L]; pa.in_last = n.in[L]; pa.ldw_last = n.ldw[L];
    pa.skip = n.skip > 0 ? n.skip : -1; pa.skip_col = n.skip > 0 ? n.out[n.skip - 1] : 0;
    pa.plain = train ? 1 : 0; pa.feat_rows = train ? n.out[L] - 1 : 0;
    P.skip_after = n.skip > 0 ? n.skip - 1 : -1;
    P.bias = bp; P.w_last = wl; P.b_last = n.b[L];
    P.clamp_radius = train ? 0.f : c.clamp_radius; P.sphere_scale = c.sphere_scale;
    const int64_t total = (int64_t)LT * 65536 + (int64_t)LT * 256 + 256;
    mf::k_pack_fused<<<nblk(total), 256, 0, c.st>>>(pa, Wp, bp, wl);
    LAUNCHED("fused weight pack");
    return msdf_tc::make_map(&fp.mW, Wp, msdf_tc::kF16, (int64_t)LT * 256, 256, 256, 256, who);
}

// one launch over Mc points.  train: maps of the saved inputs H[0 .. L] and the feature rows (may be null)
int fused_launch(const Ctx& c, const FusedPrep& fp, bool train, const float* x, const float* hashf, int64_t Mc, float* sdf,
                 const msdf_fused::StoreMaps* sm, __half* feat, int64_t ldfeat, const char* who) {
    namespace mf = msdf_fused;
    const Net& n = c.sn;
    const int L = n.L - 1;
    const int variant = (n.d0 == 39 ? 0 : 1) + (train ? 2 : 0);
    using Kern = void (*)(const CUtensorMap, const mf::Plan, const mf::StoreMaps, const float*, const float*, int64_t, float*);
    static const Kern kerns[4] = {mf::k_fused_sdf<39, 39, false>, mf::k_fused_sdf<39, 71, false>, mf::k_fused_sdf<39, 39, true>,
                                  mf::k_fused_sdf<39, 71, true>};
    static bool attr_set[4] = {false, false, false, false};
    if (!attr_set[variant]) {
        cudaError_t e = cudaFuncSetAttribute(kerns[variant], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { msdf_set_error("%s: cannot opt in to 227 KB shared memory: %s", who, cudaGetErrorString(e)); return MSDF_ERR_CUDA; }
        attr_set[variant] = true;
    }
    const size_t smem = 1024 + 2 * mf::kActBytes + mf::kWStages * mf::kWStageBytes + mf::kFusedTailBytes;
    mf::Plan P = fp.P;
    P.feat = feat; P.ldfeat = ldfeat;
    static const mf::StoreMaps no_maps{};
    double kcols = train ? 256.0 : 0.0;
    for (int l = 0; l < L; ++l) kcols += 64.0 * P.kb[l];
    const int64_t tiles = (Mc + mf::kTileRows - 1) / mf::kTileRows;
    const int grid = (int)(tiles < msdf_tc::sm_count() ? tiles : msdf_tc::sm_count());
    const int gw = c.grid ? c.enc->grid_feat_dim : 0;
    const double bytes = train ? (double)Mc * (16.0 + 4.0 * gw + 2.0 * (round_up(n.d0, 64) + 256.0 * L + (feat ? 256.0 : 0.0)))
                               : (double)Mc * (16.0 + 4.0 * gw);
    const int prof = msdf_prof_begin(train ? MSDF_PROF_TC_FUSED_TRAIN : MSDF_PROF_TC_FUSED_SDF, 2.0 * (double)Mc * 256.0 * kcols, c.st, bytes);
    kerns[variant]<<<grid, mf::kFusedThreads, smem, c.st>>>(fp.mW, P, sm ? *sm : no_maps, x, hashf, Mc, sdf);
    msdf_prof_end(prof, c.st);
    LAUNCHED("fused sdf network");
    return MSDF_OK;
}

int fused_sdf_only(const Ctx& c, const float* x, int64_t M, void* workspace, size_t workspace_bytes, float* sdf, const char* who) {
    namespace mf = msdf_fused;
    const int L = c.sn.L - 1;
    const size_t fixed = mf::fused_workspace_bytes(L);
    const int gw = c.grid ? c.enc->grid_feat_dim : 0;
    MSDF_CHECK_ARG(sdf != nullptr, "%s: sdf output missing", who);
    MSDF_CHECK_ARG(workspace_bytes >= fixed + (size_t)gw * 4 * 256, "%s: workspace of %zu bytes is too small for the fused sdf kernel (%zu)",
                   who, workspace_bytes, fixed + (size_t)gw * 4 * 256);
    char* ws = (char*)workspace;
    float* hashf = gw > 0 ? reinterpret_cast<float*>(ws + fixed) : nullptr;
    int64_t chunk = M;
    if (gw > 0) {
        const int64_t cap = (int64_t)((workspace_bytes - fixed) / ((size_t)gw * 4)) / 256 * 256;
        if (chunk > cap) chunk = cap;
    }
    FusedPrep fp;
    RUN(fused_prepare(c, ws, false, fp, who));
    for (int64_t m0 = 0; m0 < M; m0 += chunk) {
        const int64_t Mc = M - m0 < chunk ? M - m0 : chunk;
        if (gw > 0)
            RUN(msdf_hash_forward_rows(x + 3 * m0, c.enc->table, c.enc->offsets, hashf, gw, Mc, c.enc->level_dim, c.enc->n_levels,
                                       c.enc->log2_per_level_scale, (uint32_t)c.enc->base_res, c.enc->divide_factor, nullptr, c.st));
        RUN(fused_launch(c, fp, false, x + 3 * m0, hashf, Mc, sdf + m0, nullptr, nullptr, 0, who));
    }
    return MSDF_OK;
}

// training / rendering forward of one chunk: hash features (+ dy_dx), then ONE launch that leaves H[0 .. L], the raw sdf and
// the feature rows (replaces encode + the per-layer forward sweep + sdf head + feature head)
bool fused_train_applicable(const Ctx& c) {
    return fused_applicable(c) && c.sn.out[c.sn.L - 1] == 257;
}
template <class T>
int fused_forward_chunk(const Ctx& c, const Bufs<T>& b, const FusedPrep& fp, const float* x, int64_t Mc, Fw<T>* feat, int64_t ldf_,
                        bool want_dydx, const char* who) {
    namespace mf = msdf_fused;
    const Net& n = c.sn;
    const int L = n.L - 1;
    if (c.grid)
        RUN(msdf_hash_forward_rows(x, c.enc->table, c.enc->offsets, b.hashf, c.enc->grid_feat_dim, Mc, c.enc->level_dim, c.enc->n_levels,
                                   c.enc->log2_per_level_scale, (uint32_t)c.enc->base_res, c.enc->divide_factor,
                                   want_dydx ? b.dydx : nullptr, c.st));
    mf::StoreMaps sm{};
    for (int l = 0; l <= L; ++l)
        RUN(msdf_tc::make_map(&sm.m[l], b.H[l], msdf_tc::kF16, Mc, l == 0 ? b.d0p : 256, l == 0 ? b.d0p : b.ldh, 128, who));
    return fused_launch(c, fp, true, x, c.grid ? b.hashf : nullptr, Mc, b.sdf_raw, &sm, reinterpret_cast<__half*>(feat), ldf_, who);
}

// ----------------------------------------------------------------------------------------------------------
// forward / backward drivers, generic over T
// ----------------------------------------------------------------------------------------------------------
template <class T>
int field_forward(Ctx& c, const float* x, int64_t M, const float* view_dirs, int n_samples, const float* code, int mode,
                  void* workspace, size_t workspace_bytes, float* sdf, float* grad, float* feat, int64_t ld_feat, float* rgb,
                  void* saved, size_t saved_bytes, const char* who) {
    if constexpr (kIsBf16<T>)
        if (mode == MSDF_MODE_SDF_ONLY && fused_applicable(c)) return fused_sdf_only(c, x, M, workspace, workspace_bytes, sdf, who);
    const bool saving = saved != nullptr && mode == MSDF_MODE_FORWARD;
    int64_t chunk;
    size_t stride = 0;
    if (saving) {
        MSDF_CHECK_ARG(grad != nullptr, "%s: saving activations needs the grad output", who);
        chunk = saved_chunk<T>(c, M, n_samples);
        MSDF_CHECK_ARG(chunk > 0, "%s: n_samples=%d exceeds the chunk size", who, n_samples);
        stride = saved_stride<T>(c, chunk);
        const size_t need = stride * (size_t)((M + chunk - 1) / chunk);
        MSDF_CHECK_ARG(saved_bytes >= need, "%s: saved buffer of %zu bytes, need %zu", who, saved_bytes, need);
        const size_t wneed = carve<T>(c, (chunk + 127) / 128 * 128, mode, nullptr, nullptr, nullptr, nullptr, nullptr, true);
        MSDF_CHECK_ARG(workspace_bytes >= wneed, "%s: workspace of %zu bytes, need %zu with saved activations", who, workspace_bytes, wneed);
    } else {
        chunk = pick_chunk<T>(c, M, mode, workspace_bytes);
        MSDF_CHECK_ARG(chunk > 0, "%s: workspace of %zu bytes is too small (need %zu for 128 points)", who, workspace_bytes,
                       carve<T>(c, 128, mode, nullptr, nullptr, nullptr, nullptr));
        if (c.has_color && chunk < M) {   // keep chunks ray aligned
            chunk = chunk / n_samples * n_samples;
            MSDF_CHECK_ARG(chunk > 0, "%s: workspace too small for one ray of %d samples", who, n_samples);
        }
    }
    Bufs<T> b{};
    carve<T>(c, (chunk + 127) / 128 * 128, mode, workspace, &b, &c.sn, &c.cn, saved, saving);
    if (kIsBf16<T>) {
        RUN(prep_weights(c, c.sn, 1));
        if (c.has_color) { c.cn.rot0 = c.cg.fc; RUN(prep_weights(c, c.cn, 0)); }
    }
    const bool with_grad = mode == MSDF_MODE_FORWARD && (grad != nullptr);
    MSDF_CHECK_ARG(!(kIsBf16<T> && feat != nullptr), "%s: the raw feature output is only available in fp32 mode", who);
    // tensor-core mode: the forward sweep of every chunk is ONE launch of the fused kernel (fused_mlp.cuh, kTrain)
    bool fused_fwd = false;
    FusedPrep fprep;
    if constexpr (kIsBf16<T>) {
        fused_fwd = mode == MSDF_MODE_FORWARD && fused_train_applicable(c) && b.fusedW != nullptr && b.ldh == 256;
        if (fused_fwd) RUN(fused_prepare(c, b.fusedW, true, fprep, who));
    }
    for (int64_t m0 = 0; m0 < M; m0 += chunk) {
        const int64_t Mc = (M - m0 < chunk) ? M - m0 : chunk;
        const float* xc = x + 3 * m0;
        if (saving && m0 > 0)   // this chunk's slice of the saved buffer (the workspace part is reused)
            carve<T>(c, (chunk + 127) / 128 * 128, mode, workspace, &b, nullptr, nullptr, (char*)saved + stride * (size_t)(m0 / chunk), true);
        Fw<T>* featc = nullptr; int64_t ldf_ = 0;
        if (c.has_color) { featc = b.X + (kIsBf16<T> ? 0 : c.cg.fc); ldf_ = b.ldx; }
        else if (feat != nullptr) { featc = reinterpret_cast<Fw<T>*>(feat + m0 * ld_feat); ldf_ = ld_feat; }
        if (fused_fwd) {
            if constexpr (kIsBf16<T>) RUN(fused_forward_chunk<T>(c, b, fprep, xc, Mc, featc, ldf_, with_grad, who));
        } else {
            RUN(encode_chunk<T>(c, b, xc, Mc, with_grad));
            RUN(forward_sweep<T>(c, b, Mc, featc, ldf_));
        }
        if (with_grad) RUN(reverse_sweep<T>(c, b, Mc));
        RUN(decode_chunk<T>(c, b, xc, Mc, with_grad, sdf ? sdf + m0 : nullptr, with_grad ? grad + 3 * m0 : nullptr,
                            saving ? b.mask : nullptr));
        if (c.has_color) {
            const int64_t ray0 = m0 / n_samples;
            RUN(color_forward<T>(c, b, xc, Mc, view_dirs + 3 * ray0, n_samples,
                                 code ? code + (c.cd->code_per_ray ? ray0 * c.cd->code_dim : 0) : nullptr, grad + 3 * m0,
                                 rgb + (int64_t)(c.cn.tap >= 0 ? 2 * kTapN : c.cn.out[c.cn.L - 1]) * m0));
        }
    }
    return MSDF_OK;
}

template <class T>
int field_backward(Ctx& c, const float* x, int64_t M, const float* view_dirs, int n_samples, const float* code, void* workspace,
                   size_t workspace_bytes, const float* d_sdf, const float* d_grad, const float* d_feat, int64_t ld_dfeat,
                   const float* rgb, const float* d_rgb, const msdf_mlp_grads* sdf_grads, const msdf_mlp_grads* color_grads,
                   float* grad_table, float* d_code, void* saved, size_t saved_bytes, const char* who) {
    const bool have_saved = saved != nullptr;
    int64_t chunk;
    size_t stride = 0;
    if (have_saved) {
        chunk = saved_chunk<T>(c, M, n_samples);
        MSDF_CHECK_ARG(chunk > 0, "%s: n_samples=%d exceeds the chunk size", who, n_samples);
        stride = saved_stride<T>(c, chunk);
        const size_t need = stride * (size_t)((M + chunk - 1) / chunk);
        MSDF_CHECK_ARG(saved_bytes >= need, "%s: saved buffer of %zu bytes, need %zu", who, saved_bytes, need);
        const size_t wneed = carve<T>(c, (chunk + 127) / 128 * 128, MSDF_MODE_BACKWARD, nullptr, nullptr, nullptr, nullptr, nullptr, true);
        MSDF_CHECK_ARG(workspace_bytes >= wneed, "%s: workspace of %zu bytes, need %zu with saved activations", who, workspace_bytes, wneed);
    } else {
        chunk = pick_chunk<T>(c, M, MSDF_MODE_BACKWARD, workspace_bytes);
        MSDF_CHECK_ARG(chunk > 0, "%s: workspace of %zu bytes is too small (need %zu for 128 points)", who, workspace_bytes,
                       carve<T>(c, 128, MSDF_MODE_BACKWARD, nullptr, nullptr, nullptr, nullptr));
        if (c.has_color && chunk < M) {   // keep chunks ray aligned
            chunk = chunk / n_samples * n_samples;
            MSDF_CHECK_ARG(chunk > 0, "%s: workspace too small for one ray of %d samples", who, n_samples);
        }
    }
    Bufs<T> b{};
    carve<T>(c, (chunk + 127) / 128 * 128, MSDF_MODE_BACKWARD, workspace, &b, &c.sn, &c.cn, saved, have_saved);
    if (kIsBf16<T>) {
        RUN(prep_weights(c, c.sn, 1));
        if (c.has_color) { c.cn.rot0 = c.cg.fc; RUN(prep_weights(c, c.cn, 0)); }
    }
    bool fused_fwd = false;
    FusedPrep fprep;
    if constexpr (kIsBf16<T>) {
        fused_fwd = !have_saved && fused_train_applicable(c) && b.fusedW != nullptr && b.ldh == 256;
        if (fused_fwd) RUN(fused_prepare(c, b.fusedW, true, fprep, who));
    }
    const int out_last = c.sn.out[c.sn.L - 1];
    const int feat_w = out_last - 1;
    const int sdf_col = kIsBf16<T> ? feat_w : 0, feat_col0 = kIsBf16<T> ? 0 : 1;
    for (int64_t m0 = 0; m0 < M; m0 += chunk) {
        const int64_t Mc = (M - m0 < chunk) ? M - m0 : chunk;
        const float* xc = x + 3 * m0;
        if (have_saved) {
            // ---- the forward of this step left the chunk's activations in `saved`
            if (m0 > 0)
                carve<T>(c, (chunk + 127) / 128 * 128, MSDF_MODE_BACKWARD, workspace, &b, nullptr, nullptr,
                         (char*)saved + stride * (size_t)(m0 / chunk), true);
        } else {
            // ---- recompute the chunk
            if (fused_fwd) {
                if constexpr (kIsBf16<T>)
                    RUN(fused_forward_chunk<T>(c, b, fprep, xc, Mc, c.has_color ? b.X : nullptr, c.has_color ? b.ldx : 0, true, who));
            } else {
                RUN(encode_chunk<T>(c, b, xc, Mc, true));
                RUN(forward_sweep<T>(c, b, Mc, c.has_color ? b.X + (kIsBf16<T> ? 0 : c.cg.fc) : nullptr, c.has_color ? b.ldx : 0));
            }
            RUN(reverse_sweep<T>(c, b, Mc));
            RUN(decode_chunk<T>(c, b, xc, Mc, true, b.sdfc, b.gradc, b.mask));
        }
        if (c.has_color) {
            const int64_t ray0 = m0 / n_samples;
            const int no = c.cn.tap >= 0 ? 2 * kTapN : c.cn.out[c.cn.L - 1];   // spec: [rgb | rgb_spec]
            const msdf_color_desc* cd = c.cd;
            if (!have_saved)
                RUN(color_forward<T>(c, b, xc, Mc, view_dirs + 3 * ray0, n_samples,
                                     code ? code + (cd->code_per_ray ? ray0 * cd->code_dim : 0) : nullptr, b.gradc, nullptr));
            RUN(color_backward<T>(c, b, Mc, rgb + (int64_t)no * m0, d_rgb + (int64_t)no * m0, color_grads));
            if (cd->code_dim > 0 && d_code) {
                if (cd->code_per_ray) {
                    const int64_t nr = Mc / n_samples;
                    k_code_grad<<<nblk(nr * cd->code_dim), 256, 0, c.st>>>(b.dcode, nr, n_samples, cd->code_dim, d_code + ray0 * cd->code_dim);
                    LAUNCHED("per-ray code gradient");
                } else {
                    RUN(colsum<float>(c, b.dcode, cd->code_dim, nullptr, 0, Mc, cd->code_dim, d_code));
                }
            }
        }
        // ---- adjoints of (sdf, feat, grad) -> tangent of the encoded input
        {
            const int t_cols = (int)b.d0p, out_cols = (int)b.ldo;
            // bf16 layout [features..., sdf, padding] with the features already written by the colour dgrad and no
            // external d_feat: only the tail columns need writing, done by the row kernel
            const bool tail_only = kIsBf16<T> && c.has_color && d_feat == nullptr && feat_w % 8 == 0;
            if (!tail_only) {
                k_backward_prologue<T><<<nblk(Mc * out_cols), 256, 0, c.st>>>(
                    xc, Mc, c.pe_w, c.enc->grid_feat_dim, c.enc->n_levels, c.enc->level_dim, c.grid ? b.dydx : nullptr, c.hash_chain,
                    b.mask, d_sdf ? d_sdf + m0 : nullptr, d_grad ? d_grad + 3 * m0 : nullptr, c.has_color ? b.dn_color : nullptr,
                    d_feat ? d_feat + m0 * ld_dfeat : nullptr, ld_dfeat, feat_w, c.has_color ? 1 : 0, sdf_col, feat_col0, b.Dout, b.ldo,
                    out_cols, b.dn, b.TG0, b.d0p, 0);
                LAUNCHED("backward prologue (output adjoints)");
            }
            k_backward_rows<T><<<nblk(Mc, 128), 128, 0, c.st>>>(
                xc, Mc, c.enc->multires, c.pe_w, c.enc->grid_feat_dim, c.enc->n_levels, c.enc->level_dim, c.grid ? b.dydx : nullptr,
                c.hash_chain, b.mask, d_sdf ? d_sdf + m0 : nullptr, d_grad ? d_grad + 3 * m0 : nullptr,
                c.has_color ? b.dn_color : nullptr, sdf_col, tail_only ? feat_w : -1, b.Dout, b.ldo, out_cols, b.dn, b.TG0, b.d0p, t_cols);
            LAUNCHED("backward prologue");
        }
        RUN(sdf_backward<T>(c, b, xc, Mc, sdf_grads, grad_table));
    }
    return MSDF_OK;
}

}  // namespace

// =============================================================================================================
// C ABI
// =============================================================================================================
extern "C" size_t msdf_field_workspace_bytes(const msdf_mlp_desc* sdf_net, const msdf_encoding_desc* enc,
                                             const msdf_mlp_desc* color_net, const msdf_color_desc* cd, int64_t chunk_points,
                                             int mode, unsigned flags) {
    Ctx c{};
    msdf_mlp_desc tmp_s = *sdf_net;
    static const float dummy = 0.f;
    for (int l = 0; l < tmp_s.n_layers && l < MSDF_MAX_LAYERS; ++l) { if (!tmp_s.W[l]) tmp_s.W[l] = &dummy; if (!tmp_s.b[l]) tmp_s.b[l] = &dummy; }
    msdf_mlp_desc tmp_c{};
    if (color_net) { tmp_c = *color_net; for (int l = 0; l < tmp_c.n_layers && l < MSDF_MAX_LAYERS; ++l) { if (!tmp_c.W[l]) tmp_c.W[l] = &dummy; if (!tmp_c.b[l]) tmp_c.b[l] = &dummy; } }
    if (mode == MSDF_MODE_SDF_ONLY) color_net = nullptr;
    if (make_ctx(c, &tmp_s, enc, color_net ? &tmp_c : nullptr, cd, 0.f, 1.f, nullptr, flags, "msdf_field_workspace_bytes")) return 0;
    int64_t mc = (chunk_points + 127) / 128 * 128;
    if (mc < 128) mc = 128;
    if (flags & MSDF_FLAG_TENSOR_BF16) return carve<bf16>(c, mc, mode, nullptr, nullptr, nullptr, nullptr);
    return carve<float>(c, mc, mode, nullptr, nullptr, nullptr, nullptr);
}

extern "C" size_t msdf_field_saved_bytes(const msdf_mlp_desc* sdf_net, const msdf_encoding_desc* enc, const msdf_mlp_desc* color_net,
                                         const msdf_color_desc* cd, int64_t M, int n_samples, unsigned flags) {
    Ctx c{};
    msdf_mlp_desc tmp_s = *sdf_net;
    static const float dummy = 0.f;
    for (int l = 0; l < tmp_s.n_layers && l < MSDF_MAX_LAYERS; ++l) { if (!tmp_s.W[l]) tmp_s.W[l] = &dummy; if (!tmp_s.b[l]) tmp_s.b[l] = &dummy; }
    msdf_mlp_desc tmp_c{};
    if (color_net) { tmp_c = *color_net; for (int l = 0; l < tmp_c.n_layers && l < MSDF_MAX_LAYERS; ++l) { if (!tmp_c.W[l]) tmp_c.W[l] = &dummy; if (!tmp_c.b[l]) tmp_c.b[l] = &dummy; } }
    if (make_ctx(c, &tmp_s, enc, color_net ? &tmp_c : nullptr, cd, 0.f, 1.f, nullptr, flags, "msdf_field_saved_bytes")) return 0;
    if (M <= 0) return 0;
    if (n_samples < 1) n_samples = 1;
    const bool tc = (flags & MSDF_FLAG_TENSOR_BF16) != 0;
    const int64_t chunk = tc ? saved_chunk<bf16>(c, M, n_samples) : saved_chunk<float>(c, M, n_samples);
    if (chunk <= 0) return 0;
    const size_t stride = tc ? saved_stride<bf16>(c, chunk) : saved_stride<float>(c, chunk);
    return stride * (size_t)((M + chunk - 1) / chunk);
}

extern "C" int msdf_field_forward(const msdf_mlp_desc* sdf_net, const msdf_encoding_desc* enc, const msdf_mlp_desc* color_net,
                                  const msdf_color_desc* cd, const float* x, int64_t M, const float* view_dirs, int64_t n_rays,
                                  int n_samples, const float* code, int mode, float clamp_radius, float sphere_scale,
                                  unsigned flags, void* workspace, size_t workspace_bytes, float* sdf, float* grad,
                                  float* feat, int64_t ld_feat, float* rgb, void* saved, size_t saved_bytes, void* stream) {
    const char* who = "msdf_field_forward";
    MSDF_CHECK_ARG(mode == MSDF_MODE_SDF_ONLY || mode == MSDF_MODE_FORWARD, "%s: bad mode %d", who, mode);
    if (mode == MSDF_MODE_SDF_ONLY) color_net = nullptr;
    if (feat != nullptr) flags &= ~MSDF_FLAG_TENSOR_BF16;   // raw feature output: fp32 engine
    Ctx c{};
    RUN(make_ctx(c, sdf_net, enc, color_net, cd, clamp_radius, sphere_scale, stream, flags, who));
    if (M == 0) return MSDF_OK;
    MSDF_CHECK_ARG(x && workspace, "%s: null x / workspace", who);
    MSDF_CHECK_ARG(mode == MSDF_MODE_SDF_ONLY || grad != nullptr || !c.has_color, "%s: grad output required with a colour net", who);
    if (c.has_color) {
        MSDF_CHECK_ARG(view_dirs && rgb && n_samples > 0 && n_rays * (int64_t)n_samples == M, "%s: colour net needs view_dirs, rgb and M == n_rays*n_samples", who);
        MSDF_CHECK_ARG(cd->code_dim == 0 || code, "%s: per-image code missing", who);
        MSDF_CHECK_ARG(feat == nullptr, "%s: feat output and colour net are mutually exclusive", who);
    }
    if (flags & MSDF_FLAG_TENSOR_BF16)
        return field_forward<bf16>(c, x, M, view_dirs, n_samples, code, mode, workspace, workspace_bytes, sdf, grad, feat, ld_feat, rgb, saved, saved_bytes, who);
    return field_forward<float>(c, x, M, view_dirs, n_samples, code, mode, workspace, workspace_bytes, sdf, grad, feat, ld_feat, rgb, saved, saved_bytes, who);
}

extern "C" int msdf_field_backward(const msdf_mlp_desc* sdf_net, const msdf_encoding_desc* enc, const msdf_mlp_desc* color_net,
                                   const msdf_color_desc* cd, const float* x, int64_t M, const float* view_dirs, int64_t n_rays,
                                   int n_samples, const float* code, float clamp_radius, float sphere_scale, unsigned flags,
                                   void* workspace, size_t workspace_bytes, const float* d_sdf, const float* d_grad,
                                   const float* d_feat, int64_t ld_dfeat, const float* rgb, const float* d_rgb,
                                   const msdf_mlp_grads* sdf_grads, const msdf_mlp_grads* color_grads, float* grad_table,
                                   float* d_code, void* saved, size_t saved_bytes, void* stream) {
    const char* who = "msdf_field_backward";
    if (d_rgb == nullptr) color_net = nullptr;
    Ctx c{};
    RUN(make_ctx(c, sdf_net, enc, color_net, cd, clamp_radius, sphere_scale, stream, flags, who));
    if (M == 0) return MSDF_OK;
    MSDF_CHECK_ARG(x && workspace && sdf_grads, "%s: null x / workspace / sdf_grads", who);
    for (int l = 0; l < c.sn.L; ++l) MSDF_CHECK_ARG(sdf_grads->dW[l] && sdf_grads->db[l], "%s: sdf grad buffer %d missing", who, l);
    if (c.has_color) {
        MSDF_CHECK_ARG(view_dirs && rgb && color_grads && n_samples > 0 && n_rays * (int64_t)n_samples == M, "%s: colour net needs view_dirs, rgb, color_grads and M == n_rays*n_samples", who);
        for (int l = 0; l < c.cn.L; ++l) MSDF_CHECK_ARG(color_grads->dW[l] && color_grads->db[l], "%s: colour grad buffer %d missing", who, l);
        MSDF_CHECK_ARG(cd->code_dim == 0 || code, "%s: per-image code missing", who);
    }
    if (flags & MSDF_FLAG_TENSOR_BF16)
        return field_backward<bf16>(c, x, M, view_dirs, n_samples, code, workspace, workspace_bytes, d_sdf, d_grad, d_feat, ld_dfeat,
                                    rgb, d_rgb, sdf_grads, color_grads, grad_table, d_code, saved, saved_bytes, who);
    return field_backward<float>(c, x, M, view_dirs, n_samples, code, workspace, workspace_bytes, d_sdf, d_grad, d_feat, ld_dfeat,
                                 rgb, d_rgb, sdf_grads, color_grads, grad_table, d_code, saved, saved_bytes, who);
}

extern "C" void msdf_set_fused(int on) { g_fused_enabled = on; }
extern "C" void msdf_set_sweeps(int stream, int chain) { g_stream_enabled = stream; g_chain_enabled = chain; }

extern "C" int msdf_ray_points(const float* ray_o, const float* ray_d, const float* z, int64_t n_rays, int n, float* points,
                               void* stream) {
    MSDF_CHECK_ARG(ray_o && ray_d && z && points, "msdf_ray_points: null pointer");
    if (n_rays == 0 || n == 0) return MSDF_OK;
    k_ray_points<<<nblk(n_rays * n * 3), 256, 0, (cudaStream_t)stream>>>(ray_o, ray_d, z, n_rays, n, points);
    LAUNCHED("msdf_ray_points");
    return MSDF_OK;
}
