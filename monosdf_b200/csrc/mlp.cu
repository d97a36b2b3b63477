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
// Nothing is kept between forward and backward: the backward recomputes the chunk.
#include "gemm_f32.cuh"

int msdf_hash_forward_rows(const float* x, const float* table, const int* offsets, float* out, int64_t out_ld,
                           int64_t B, int C, int L, float S, uint32_t H, float divide_factor, float* dy_dx, cudaStream_t st);
int msdf_hash_scatter_rows(const float* x, const int* offsets, int64_t B, int C, int L, float S, uint32_t H, float divide_factor,
                           const float* grad, int64_t grad_ld, const float* grad2, int64_t grad2_ld, const float* gg_x,
                           float gg_scale, float* grad_table, cudaStream_t st);

namespace {

using namespace msdf_gemm;

constexpr float kInvSqrt2 = 0.70710678118654752440f;
constexpr float kSqrt2 = 1.41421356237309504880f;

__device__ __forceinline__ float softplus100(float p) {   // nn.Softplus(beta=100), threshold 20 (network.py:77)
    const float bp = 100.f * p;
    return bp > 20.f ? p : log1pf(expf(bp)) / 100.f;
}
// sigmoid(100 p) from h = softplus100(p):  1 - exp(-100 h)
__device__ __forceinline__ float sig_from_h(float h) { return -expm1f(-100.f * h); }
// 100 * (1 - sigmoid(100 p)) from h;  exactly 0 past the softplus threshold like torch's double backward
__device__ __forceinline__ float dsig_over_sig_from_h(float h) {
    const float t = 100.f * h;
    return t > 20.f ? 0.f : 100.f * expf(-t);
}

// ----------------------------------------------------------------------------------------------------------
// network geometry
// ----------------------------------------------------------------------------------------------------------
struct Net {
    int L, d0, d0p, skip, ldh;
    int in[MSDF_MAX_LAYERS], out[MSDF_MAX_LAYERS];
    int64_t ldw[MSDF_MAX_LAYERS];
    const float* W[MSDF_MAX_LAYERS];
    const float* b[MSDF_MAX_LAYERS];
};

inline int round4(int x) { return (x + 3) / 4 * 4; }

int make_net(const msdf_mlp_desc* d, Net& n, const char* who) {
    MSDF_CHECK_ARG(d != nullptr, "%s: null network descriptor", who);
    MSDF_CHECK_ARG(d->n_layers >= 2 && d->n_layers <= MSDF_MAX_LAYERS, "%s: n_layers=%d not in [2,%d]", who, d->n_layers,
                   MSDF_MAX_LAYERS);
    n.L = d->n_layers; n.d0 = d->d0; n.d0p = round4(d->d0); n.skip = d->skip_layer;
    MSDF_CHECK_ARG(n.skip < n.L && n.skip != 0, "%s: skip_layer=%d invalid", who, n.skip);
    int w = 0;
    for (int l = 0; l < n.L; ++l) {
        n.in[l] = d->in_dim[l]; n.out[l] = d->out_dim[l]; n.ldw[l] = d->ldw[l]; n.W[l] = d->W[l]; n.b[l] = d->b[l];
        MSDF_CHECK_ARG(n.W[l] && n.b[l], "%s: layer %d has null weights", who, l);
        MSDF_CHECK_ARG(n.ldw[l] >= n.in[l] && n.in[l] > 0 && n.out[l] > 0, "%s: layer %d bad dims", who, l);
        const int expect = (l == 0) ? n.d0 : (l == n.skip ? n.out[l - 1] + n.d0 : n.out[l - 1]);
        MSDF_CHECK_ARG(n.in[l] == expect, "%s: layer %d in_dim=%d, expected %d", who, l, n.in[l], expect);
        if (l > 0) w = w > n.in[l] ? w : n.in[l];
        if (l < n.L - 1) w = w > n.out[l] ? w : n.out[l];
    }
    n.ldh = round4(w);
    return MSDF_OK;
}

// ----------------------------------------------------------------------------------------------------------
// encoding kernels  (embedder.py:5-50; hash features are written by hashgrid.cu)
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

// H0[m, j] = PE_j(x_m) for j < pe_w   (hash columns are filled by msdf_hash_forward_rows; padding untouched)
__global__ void k_encode(const float* __restrict__ x, int64_t M, int pe_w, float* __restrict__ H0, int64_t ld0) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * pe_w) return;
    const int64_t m = i / pe_w; const int j = (int)(i - m * pe_w);
    const float p[3] = {x[3 * m], x[3 * m + 1], x[3 * m + 2]};
    H0[m * ld0 + j] = pe_value(p, j);
}

__global__ void k_zero_cols(float* __restrict__ dst, int64_t ld, int64_t M, int c0, int w) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * w) return;
    const int64_t m = i / w; const int j = (int)(i - m * w);
    dst[m * ld + c0 + j] = 0.f;
}

// dst[m, c0 + j] = src[m, j] * scale   (the [.., h0]/sqrt2 half of the skip concat, network.py:88-89)
__global__ void k_skip_copy(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd, int64_t M,
                            int w, int c0, float scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * w) return;
    const int64_t m = i / w; const int j = (int)(i - m * w);
    dst[m * ldd + c0 + j] = src[m * lds + j] * scale;
}

// grad_x = J_enc(x)^T g0 (+ hash chain), then the bounding-sphere clamp of get_outputs (network.py:116-118):
//   sdf = min(sdf_raw, sphere_scale (R - |x|)); the gradient follows the selected branch (ties split 1/2, like
//   torch.minimum's backward).  mask[m] = d sdf / d sdf_raw  in {1, 0, 0.5}.
__global__ void k_decode(const float* __restrict__ x, const float* __restrict__ g0, int64_t ldg, int64_t M, int pe_w,
                         int grid_w, int n_levels, int level_dim, const float* __restrict__ dy_dx, float hash_chain,
                         const float* __restrict__ sdf_raw, float clamp_radius, float sphere_scale,
                         float* __restrict__ sdf_out, float* __restrict__ grad_out, float* __restrict__ mask_out) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float p[3] = {x[3 * m], x[3 * m + 1], x[3 * m + 2]};
    float g[3] = {0.f, 0.f, 0.f};
    if (g0 != nullptr) {
        const float* gr = g0 + m * ldg;
        for (int j = 0; j < pe_w; ++j) g[pe_dim(j)] += gr[j] * pe_deriv(p, j);
        if (grid_w > 0 && dy_dx != nullptr) {
            const float* dd = dy_dx + m * (int64_t)(n_levels * 3 * level_dim);
            float h[3] = {0.f, 0.f, 0.f};
            for (int l = 0; l < n_levels; ++l)
                for (int d = 0; d < 3; ++d)
                    for (int c = 0; c < level_dim; ++c) h[d] += gr[pe_w + l * level_dim + c] * dd[(l * 3 + d) * level_dim + c];
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

// Backward prologue: dn[m] = mask * (d_grad[m] + dn_color[m]);  Dout[m,0] = mask * d_sdf[m];
//   Dout[m,1+j] (+)= d_feat[m,j];   TG0[m, j] = (J_enc dn)[j]  -- the tangent of the encoded input.
__global__ void k_backward_prologue(const float* __restrict__ x, int64_t M, int pe_w, int grid_w, int n_levels, int level_dim,
                                    const float* __restrict__ dy_dx, float hash_chain, const float* __restrict__ mask,
                                    const float* __restrict__ d_sdf, const float* __restrict__ d_grad,
                                    const float* __restrict__ dn_color, const float* __restrict__ d_feat, int64_t ld_dfeat,
                                    int feat_w, int have_color_feat, float* __restrict__ Dout, int64_t ldo,
                                    float* __restrict__ dn, float* __restrict__ TG0, int64_t ldt) {
    const int d0 = pe_w + grid_w;
    const int cols = d0 > feat_w + 1 ? d0 : feat_w + 1;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * cols) return;
    const int64_t m = i / cols; const int j = (int)(i - m * cols);
    const float mk = mask[m];
    float v[3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
        v[d] = mk * ((d_grad ? d_grad[3 * m + d] : 0.f) + (dn_color ? dn_color[3 * m + d] : 0.f));
    if (j < 3) dn[3 * m + j] = v[j];
    if (j == 0) Dout[m * ldo] = d_sdf ? mk * d_sdf[m] : 0.f;
    if (j >= 1 && j <= feat_w) {
        float f = have_color_feat ? Dout[m * ldo + j] : 0.f;
        if (d_feat) f += d_feat[m * ld_dfeat + (j - 1)];
        Dout[m * ldo + j] = f;
    }
    if (j < d0) {
        float t;
        if (j < pe_w) {
            const float p[3] = {x[3 * m], x[3 * m + 1], x[3 * m + 2]};
            t = pe_deriv(p, j) * v[pe_dim(j)];
        } else {
            const int q = j - pe_w, l = q / level_dim, c = q - l * level_dim;
            t = 0.f;
            if (dy_dx != nullptr) {
                const float* dd = dy_dx + m * (int64_t)(n_levels * 3 * level_dim) + (l * 3) * level_dim + c;
                t = (dd[0] * v[0] + dd[level_dim] * v[1] + dd[2 * level_dim] * v[2]) * hash_chain;
            }
        }
        TG0[m * ldt + j] = t;
    }
}

// out[n] += sum_m w[m*ws] * X[m*ldx + n]   (w == nullptr -> 1).  Bias gradients and the sdf row of the last layer.
__global__ void __launch_bounds__(256)
k_wcolsum(const float* __restrict__ X, int64_t ldx, const float* __restrict__ w, int64_t ws, int64_t M, int N,
          int64_t rows_per_block, float* __restrict__ out) {
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = (r0 + rows_per_block < M) ? r0 + rows_per_block : M;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float s = 0.f;
        for (int64_t m = r0; m < r1; ++m) s += (w ? w[m * ws] : 1.0f) * X[m * ldx + n];
        atomicAdd(out + n, s);
    }
}

// ----------------------------------------------------------------------------------------------------------
// small-N layers: out[m,n] = act(sum_k A[m,k] W[n,k] + b[n]), N <= 4, one warp per row
// ----------------------------------------------------------------------------------------------------------
enum Act { kActNone = 0, kActSigmoid = 1, kActRelu = 2 };

template <int NMAX>
__global__ void __launch_bounds__(256)
k_rowdot(const float* __restrict__ A, int64_t lda, const float* __restrict__ W, int64_t ldw, const float* __restrict__ b,
         int64_t M, int N, int K, int act, float* __restrict__ out, int64_t ldo) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (m >= M) return;
    float s[NMAX];
#pragma unroll
    for (int n = 0; n < NMAX; ++n) s[n] = 0.f;
    const float* a = A + m * lda;
    for (int k = lane; k < K; k += 32) {
        const float av = a[k];
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

__device__ __forceinline__ float act_grad(int act, float y) {   // d act / d pre as a function of the output y
    return act == kActSigmoid ? y * (1.0f - y) : (act == kActRelu ? (y > 0.f ? 1.f : 0.f) : 1.f);
}

// dA[m,k] = relu'(A[m,k]) * sum_n dpre[m,n] W[n,k],   dpre = d_out * act'(out)
template <int NMAX>
__global__ void k_rowdot_dgrad(const float* __restrict__ d_out, const float* __restrict__ out, int64_t ldo, int act,
                               const float* __restrict__ W, int64_t ldw, const float* __restrict__ A, int64_t lda,
                               int64_t M, int N, int K, float* __restrict__ dA, int64_t ldda) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * K) return;
    const int64_t m = i / K; const int k = (int)(i - m * K);
    float s = 0.f;
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
        if (n < N) s = fmaf(d_out[m * ldo + n] * act_grad(act, out[m * ldo + n]), __ldg(W + n * ldw + k), s);
    dA[m * ldda + k] = A[m * lda + k] > 0.f ? s : 0.f;
}

// dW[n,k] += sum_m dpre[m,n] A[m,k];  db[n] += sum_m dpre[m,n]
template <int NMAX>
__global__ void __launch_bounds__(256)
k_rowdot_wgrad(const float* __restrict__ d_out, const float* __restrict__ out, int64_t ldo, int act,
               const float* __restrict__ A, int64_t lda, int64_t M, int N, int K, int64_t rows_per_block,
               float* __restrict__ dW, int64_t ldw, float* __restrict__ db) {
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = (r0 + rows_per_block < M) ? r0 + rows_per_block : M;
    for (int k = threadIdx.x; k < K + 1; k += blockDim.x) {
        float s[NMAX];
#pragma unroll
        for (int n = 0; n < NMAX; ++n) s[n] = 0.f;
        for (int64_t m = r0; m < r1; ++m) {
            const float av = (k < K) ? A[m * lda + k] : 1.0f;
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
// GEMM epilogues
// ----------------------------------------------------------------------------------------------------------
struct EpiFwdAct {   // out[m,n] = softplus100(acc + b[n]) * oscale
    const float* bias; float* out; int64_t ldo; float oscale;
    __device__ __forceinline__ void operator()(int64_t m, int n, const float v[4], int nv) const {
        float* o = out + m * ldo + n;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < nv) o[j] = softplus100(v[j] + __ldg(bias + n + j)) * oscale;
    }
};
struct EpiBias {     // out[m,n] = acc + b[n]
    const float* bias; float* out; int64_t ldo;
    __device__ __forceinline__ void operator()(int64_t m, int n, const float v[4], int nv) const {
        float* o = out + m * ldo + n;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < nv) o[j] = v[j] + __ldg(bias + n + j);
    }
};
struct EpiRelu {     // out[m,n] = relu(acc + b[n])
    const float* bias; float* out; int64_t ldo;
    __device__ __forceinline__ void operator()(int64_t m, int n, const float v[4], int nv) const {
        float* o = out + m * ldo + n;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < nv) o[j] = fmaxf(v[j] + __ldg(bias + n + j), 0.f);
    }
};
struct EpiAtomic {   // C[m,n] += acc   (split-K weight gradients)
    float* C; int64_t ldc;
    __device__ __forceinline__ void operator()(int64_t m, int n, const float v[4], int nv) const {
        float* o = C + m * ldc + n;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < nv) atomicAdd(o + j, v[j]);
    }
};
// reverse sweep, layer l: acc = (a_l W_l)[m,n], n over the layer's inputs
struct EpiRev {
    const float* Hin; int64_t ldh; float hscale;   // stored input of layer l and the factor that undoes its scaling
    float* Aout; int64_t lda;                      // a_{l-1}
    float* g0; int64_t ldg;
    int dh; float qscale;                          // skip layer: columns >= dh are the h0 half; both halves scaled 1/sqrt2
    int layer0, g0_accum;
    __device__ __forceinline__ void operator()(int64_t m, int n, const float v[4], int nv) const {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j >= nv) break;
            const int c = n + j;
            const float r = v[j] * qscale;
            if (c >= dh) { g0[m * ldg + (c - dh)] = r; continue; }
            if (layer0) { float* g = g0 + m * ldg + c; *g = g0_accum ? *g + r : r; }
            else Aout[m * lda + c] = r * sig_from_h(Hin[m * ldh + c] * hscale);
        }
    }
};
// tangent sweep, layer l: acc = (t_l W_l^T)[m,n], n over the layer's outputs
struct EpiTan {
    const float* Hn; int64_t ldh; float hscale;    // h_{l+1} as stored (input of layer l+1)
    float* AZ; int64_t lda;                        // in: a_l, out: z_l
    float* Tout; int64_t ldt; float tscale;
    __device__ __forceinline__ void operator()(int64_t m, int n, const float v[4], int nv) const {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j >= nv) break;
            const float h = Hn[m * ldh + n + j] * hscale;
            const float s = sig_from_h(h);
            Tout[m * ldt + n + j] = v[j] * s * tscale;
            float* az = AZ + m * lda + n + j;
            *az = v[j] * (*az) * dsig_over_sig_from_h(h);
        }
    }
};
// backward sweep, layer l: acc = (pbar_l W_l)[m,n], n over the layer's inputs
struct EpiBwd {
    const float* Hin; int64_t ldh; float hscale;
    float* PZ; int64_t ldp;                        // in: z_{l-1}, out: pbar_{l-1}
    float* bh0; int64_t ldb;                       // adjoint of h_0 (hash-grid nets only), may be null
    int dh; float qscale; int layer0, bh0_accum;
    __device__ __forceinline__ void operator()(int64_t m, int n, const float v[4], int nv) const {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j >= nv) break;
            const int c = n + j;
            const float r = v[j] * qscale;
            if (c >= dh) { if (bh0) bh0[m * ldb + (c - dh)] = r; continue; }
            if (layer0) { if (bh0) { float* g = bh0 + m * ldb + c; *g = bh0_accum ? *g + r : r; } }
            else { float* p = PZ + m * ldp + c; *p = r * sig_from_h(Hin[m * ldh + c] * hscale) + *p; }
        }
    }
};
struct EpiBwdRelu {  // colour net dgrad: out[m,n] = acc * [Hin[m,n] > 0]
    const float* Hin; int64_t ldh; float* out; int64_t ldo;
    __device__ __forceinline__ void operator()(int64_t m, int n, const float v[4], int nv) const {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < nv) out[m * ldo + n + j] = Hin[m * ldh + n + j] > 0.f ? v[j] : 0.f;
    }
};
// colour net layer-0 dgrad: route d(input) columns to the SDF net's adjoints
struct EpiColorIn {
    int nc, fc, F, cc, cd;                         // column of the normal (-1 = none), of feat, of the code
    float* dn; float* Dout; int64_t ldo; float* dcode; int64_t ldc;
    __device__ __forceinline__ void operator()(int64_t m, int n, const float v[4], int nv) const {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j >= nv) break;
            const int c = n + j;
            if (nc >= 0 && c >= nc && c < nc + 3) dn[3 * m + (c - nc)] = v[j];
            else if (c >= fc && c < fc + F) Dout[m * ldo + 1 + (c - fc)] = v[j];
            else if (cd > 0 && c >= cc && c < cc + cd) dcode[m * ldc + (c - cc)] = v[j];
        }
    }
};

// reverse-sweep start: a_{L-1} = e_0, so (a W_{L-1})[m,n] = W_{L-1}[0,n] for every point
__global__ void k_rev_init(const float* __restrict__ w_row, int64_t M, int N, EpiRev epi) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int n4 = (N + 3) / 4;
    if (i >= M * n4) return;
    const int64_t m = i / n4; const int n = (int)(i - m * n4) * 4;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (n + j < N) ? __ldg(w_row + n + j) : 0.f;
    epi(m, n, v, (N - n) < 4 ? (N - n) : 4);
}

// colour-net input row (network.py:393-413): idr [x, PE(view), normal, feat, code], nerf [PE(view), feat, code].
// The feat columns are written by the SDF net's last layer; this kernel fills the rest.
__global__ void k_color_input(const float* __restrict__ x, const float* __restrict__ view, const float* __restrict__ normal,
                              const float* __restrict__ code, int64_t M, int n_samples, int mode_idr, int pe_w, int F, int cd,
                              int code_per_ray, float* __restrict__ X, int64_t ldx) {
    const int pre = (mode_idr ? 3 : 0) + pe_w + (mode_idr ? 3 : 0);
    const int cols = pre + cd;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * cols) return;
    const int64_t m = i / cols; int j = (int)(i - m * cols);
    const int64_t ray = m / n_samples;
    float v; int col;
    if (j >= pre) {
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
    X[m * ldx + col] = v;
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

inline unsigned nblk(int64_t n, int t = 256) { return (unsigned)msdf_div_up(n, t); }

#define RUN(expr) do { int rc_ = (expr); if (rc_) return rc_; } while (0)
#define LAUNCHED(name) do { MSDF_COUNT_LAUNCH(); MSDF_CHECK_LAUNCH(name); } while (0)

// ----------------------------------------------------------------------------------------------------------
// workspace layout for one chunk
// ----------------------------------------------------------------------------------------------------------
struct ColorGeom { int pe_w, nc, fc, cc, in0, ldx; };

ColorGeom color_geom(const msdf_color_desc* cd) {
    ColorGeom g;
    g.pe_w = cd->multires_view > 0 ? 3 + 6 * cd->multires_view : 3;
    if (cd->mode_idr) { g.nc = 3 + g.pe_w; g.fc = g.nc + 3; } else { g.nc = -1; g.fc = g.pe_w; }
    g.cc = g.fc + cd->feat_dim;
    g.in0 = g.cc + cd->code_dim;
    g.ldx = round4(g.in0);
    return g;
}

struct Bufs {
    float* H[MSDF_MAX_LAYERS];   // H[0] = encoded input (ld d0p); H[l] = input of layer l (ld ldh)
    float* A[MSDF_MAX_LAYERS];   // a_l, later z_l / pbar_l  (l < L-1)
    float *G0, *TG0, *BH0, *T[2], *Dout, *dydx, *sdf_raw, *mask, *dn, *dn_color, *gradc, *sdfc;
    float* X; float* C[MSDF_MAX_LAYERS]; float* dC[2]; float* dcode; float* rgbc;
    int64_t ldo;
};

struct Carver {
    char* base; size_t off; bool dry;
    float* take(int64_t rows, int64_t cols) {
        const size_t bytes = msdf_align((size_t)rows * (size_t)cols * sizeof(float));
        float* p = dry ? nullptr : reinterpret_cast<float*>(base + off);
        off += bytes;
        return p;
    }
};

// mode: MSDF_MODE_*
size_t carve(const Net& sn, const msdf_encoding_desc* enc, const Net* cn, const msdf_color_desc* cdesc, int64_t Mc, int mode,
             void* ws, Bufs* out) {
    Carver c{(char*)ws, 0, ws == nullptr};
    Bufs b{};
    const bool grid = enc->grid_feat_dim > 0 && enc->table != nullptr;
    b.H[0] = c.take(Mc, sn.d0p);
    b.sdf_raw = c.take(Mc, 1);
    if (mode == MSDF_MODE_SDF_ONLY) {
        float* pp[2] = {c.take(Mc, sn.ldh), c.take(Mc, sn.ldh)};
        for (int l = 1; l < sn.L; ++l) b.H[l] = pp[(l - 1) & 1];
    } else {
        for (int l = 1; l < sn.L; ++l) b.H[l] = c.take(Mc, sn.ldh);
        if (mode == MSDF_MODE_FORWARD) {
            float* pp[2] = {c.take(Mc, sn.ldh), c.take(Mc, sn.ldh)};
            for (int l = 0; l < sn.L - 1; ++l) b.A[l] = pp[l & 1];
        } else {
            for (int l = 0; l < sn.L - 1; ++l) b.A[l] = c.take(Mc, sn.ldh);
        }
        b.G0 = c.take(Mc, sn.d0p);
        if (grid) b.dydx = c.take(Mc, enc->n_levels * 3 * enc->level_dim);
        b.mask = c.take(Mc, 1);
        if (mode == MSDF_MODE_BACKWARD) {
            b.TG0 = c.take(Mc, sn.d0p);
            if (grid) b.BH0 = c.take(Mc, sn.d0p);
            b.T[0] = c.take(Mc, sn.ldh); b.T[1] = c.take(Mc, sn.ldh);
            b.ldo = round4(sn.out[sn.L - 1]);
            b.Dout = c.take(Mc, b.ldo);
            b.dn = c.take(Mc, 3); b.dn_color = c.take(Mc, 3);
            b.gradc = c.take(Mc, 3); b.sdfc = c.take(Mc, 1);
        }
        if (cn != nullptr) {
            const ColorGeom g = color_geom(cdesc);
            b.X = c.take(Mc, g.ldx);
            b.C[0] = b.X;
            for (int l = 1; l < cn->L; ++l) b.C[l] = c.take(Mc, cn->ldh);
            if (mode == MSDF_MODE_BACKWARD) {
                b.dC[0] = c.take(Mc, cn->ldh); b.dC[1] = c.take(Mc, cn->ldh);
                if (cdesc->code_dim > 0) b.dcode = c.take(Mc, cdesc->code_dim);
            }
        }
    }
    if (out) *out = b;
    return c.off;
}

int64_t pick_chunk(const Net& sn, const msdf_encoding_desc* enc, const Net* cn, const msdf_color_desc* cd, int64_t M, int mode,
                   size_t ws_bytes) {
    int64_t cap = 65536;
    if (mode == MSDF_MODE_SDF_ONLY) cap = 262144;
    int64_t mc = M < cap ? M : cap;
    mc = (mc + 127) / 128 * 128;
    while (mc > 128 && carve(sn, enc, cn, cd, mc, mode, nullptr, nullptr) > ws_bytes) mc = (mc / 2 + 127) / 128 * 128;
    if (carve(sn, enc, cn, cd, mc, mode, nullptr, nullptr) > ws_bytes) return 0;
    return mc;
}

// ----------------------------------------------------------------------------------------------------------
// sweeps over one chunk
// ----------------------------------------------------------------------------------------------------------
struct Ctx {
    Net sn; const msdf_encoding_desc* enc; bool grid; int pe_w; float hash_chain;
    bool has_color; Net cn; const msdf_color_desc* cd; ColorGeom cg;
    float clamp_radius, sphere_scale;
    cudaStream_t st;
};

inline float in_scale(const Net& n, int l) { return l == n.skip ? kSqrt2 : 1.0f; }   // undoes the stored 1/sqrt2

int encode_chunk(const Ctx& c, const Bufs& b, const float* x, int64_t Mc, bool want_dydx) {
    k_encode<<<nblk(Mc * c.pe_w), 256, 0, c.st>>>(x, Mc, c.pe_w, b.H[0], c.sn.d0p);
    LAUNCHED("encode");
    if (c.enc->grid_feat_dim > 0) {
        if (c.grid) {
            RUN(msdf_hash_forward_rows(x, c.enc->table, c.enc->offsets, b.H[0] + c.pe_w, c.sn.d0p, Mc, c.enc->level_dim,
                                       c.enc->n_levels, c.enc->log2_per_level_scale, (uint32_t)c.enc->base_res,
                                       c.enc->divide_factor, want_dydx ? b.dydx : nullptr, c.st));
        } else {   // use_grid_feature = False: zero features (network.py:251-252)
            k_zero_cols<<<nblk(Mc * c.enc->grid_feat_dim), 256, 0, c.st>>>(b.H[0], c.sn.d0p, Mc, c.pe_w, c.enc->grid_feat_dim);
            LAUNCHED("zero grid features");
        }
    }
    return MSDF_OK;
}

// forward sweep; feat (ld ldf) may be null
int forward_sweep(const Ctx& c, const Bufs& b, int64_t Mc, float* feat, int64_t ldf) {
    const Net& n = c.sn;
    for (int l = 0; l < n.L - 1; ++l) {
        const int64_t ldin = l == 0 ? n.d0p : n.ldh;
        if (l + 1 == n.skip) {
            k_skip_copy<<<nblk(Mc * n.d0), 256, 0, c.st>>>(b.H[0], n.d0p, b.H[l + 1], n.ldh, Mc, n.d0, n.out[l], kInvSqrt2);
            LAUNCHED("skip copy");
        }
        EpiFwdAct e{n.b[l], b.H[l + 1], n.ldh, l + 1 == n.skip ? kInvSqrt2 : 1.0f};
        RUN((launch<kNT>(b.H[l], ldin, n.W[l], n.ldw[l], Mc, n.out[l], n.in[l], 1, e, c.st, "sdf forward layer")));
    }
    const int l = n.L - 1;
    k_rowdot<1><<<nblk(Mc, 8), 256, 0, c.st>>>(b.H[l], n.ldh, n.W[l], n.ldw[l], n.b[l], Mc, 1, n.in[l], kActNone, b.sdf_raw, 1);
    LAUNCHED("sdf head");
    if (feat != nullptr && n.out[l] > 1) {
        EpiBias e{n.b[l] + 1, feat, ldf};
        RUN((launch<kNT>(b.H[l], n.ldh, n.W[l] + n.ldw[l], n.ldw[l], Mc, n.out[l] - 1, n.in[l], 1, e, c.st, "feature head")));
    }
    return MSDF_OK;
}

EpiRev make_rev(const Net& n, const Bufs& b, int l) {
    EpiRev e{};
    e.Hin = b.H[l]; e.ldh = l == 0 ? n.d0p : n.ldh; e.hscale = in_scale(n, l);
    e.Aout = l > 0 ? b.A[l - 1] : nullptr; e.lda = n.ldh;
    e.g0 = b.G0; e.ldg = n.d0p;
    e.dh = l == n.skip ? n.in[l] - n.d0 : n.in[l];
    e.qscale = l == n.skip ? kInvSqrt2 : 1.0f;
    e.layer0 = l == 0; e.g0_accum = n.skip > 0;
    return e;
}

int reverse_sweep(const Ctx& c, const Bufs& b, int64_t Mc) {
    const Net& n = c.sn;
    {
        const int l = n.L - 1;
        const int n4 = (n.in[l] + 3) / 4;
        k_rev_init<<<nblk(Mc * n4), 256, 0, c.st>>>(n.W[l], Mc, n.in[l], make_rev(n, b, l));
        LAUNCHED("reverse init");
    }
    for (int l = n.L - 2; l >= 0; --l)
        RUN((launch<kNN>(b.A[l], n.ldh, n.W[l], n.ldw[l], Mc, n.in[l], n.out[l], 1, make_rev(n, b, l), c.st, "sdf reverse layer")));
    return MSDF_OK;
}

int decode_chunk(const Ctx& c, const Bufs& b, const float* x, int64_t Mc, bool with_grad, float* sdf, float* grad, float* mask) {
    k_decode<<<nblk(Mc, 128), 128, 0, c.st>>>(x, with_grad ? b.G0 : nullptr, c.sn.d0p, Mc, c.pe_w, c.grid ? c.enc->grid_feat_dim : 0,
                                             c.enc->n_levels, c.enc->level_dim, b.dydx, c.hash_chain, b.sdf_raw, c.clamp_radius,
                                             c.sphere_scale, sdf, grad, mask);
    LAUNCHED("decode");
    return MSDF_OK;
}

int color_forward(const Ctx& c, const Bufs& b, const float* x, int64_t Mc, const float* view, int n_samples,
                  const float* code, const float* normal, float* rgb) {
    const Net& n = c.cn;
    const ColorGeom& g = c.cg;
    const int cols = g.in0 - c.cd->feat_dim;
    // chunks start on a ray boundary; view / code pointers are already offset to the chunk's first ray
    k_color_input<<<nblk(Mc * cols), 256, 0, c.st>>>(x, view, normal, code, Mc, n_samples, c.cd->mode_idr, g.pe_w, c.cd->feat_dim,
                                                    c.cd->code_dim, c.cd->code_per_ray, b.X, g.ldx);
    LAUNCHED("colour input");
    for (int l = 0; l < n.L - 1; ++l) {
        EpiRelu e{n.b[l], b.C[l + 1], n.ldh};
        RUN((launch<kNT>(b.C[l], l == 0 ? g.ldx : n.ldh, n.W[l], n.ldw[l], Mc, n.out[l], n.in[l], 1, e, c.st, "colour layer")));
    }
    const int l = n.L - 1;
    MSDF_CHECK_ARG(n.out[l] <= 4, "colour net: d_out=%d > 4 unsupported", n.out[l]);
    if (rgb == nullptr) return MSDF_OK;   // backward recompute: the saved rgb drives act', the head is not needed
    k_rowdot<4><<<nblk(Mc, 8), 256, 0, c.st>>>(b.C[l], n.ldh, n.W[l], n.ldw[l], n.b[l], Mc, n.out[l], n.in[l],
                                              c.cd->final_act == 0 ? kActSigmoid : kActRelu, rgb, n.out[l]);
    LAUNCHED("colour head");
    return MSDF_OK;
}

int wgrad(const float* X, int64_t ldx, const float* Y, int64_t ldy, int rows, int cols, int64_t Mc, float* dW, int64_t ldw,
          cudaStream_t st) {
    const int tiles = (int)(msdf_div_up(rows, BM) * msdf_div_up(cols, BN));
    int splits = (int)((2 * 148 + tiles - 1) / tiles);
    const int64_t max_splits = msdf_div_up(Mc, 512);
    if (splits > max_splits) splits = (int)max_splits;
    EpiAtomic e{dW, ldw};
    return launch<kTN>(X, ldx, Y, ldy, rows, cols, Mc, splits, e, st, "weight gradient");
}

int colsum(const float* X, int64_t ldx, const float* w, int64_t ws, int64_t Mc, int N, float* out, cudaStream_t st) {
    const int64_t rpb = 256;
    k_wcolsum<<<nblk(Mc, (int)rpb), 256, 0, st>>>(X, ldx, w, ws, Mc, N, rpb, out);
    LAUNCHED("column sum");
    return MSDF_OK;
}

int color_backward(const Ctx& c, const Bufs& b, int64_t Mc, const float* rgb, const float* d_rgb, const msdf_mlp_grads* gr) {
    const Net& n = c.cn;
    const ColorGeom& g = c.cg;
    int l = n.L - 1;
    const int act = c.cd->final_act == 0 ? kActSigmoid : kActRelu;
    k_rowdot_wgrad<4><<<nblk(Mc, 512), 256, 0, c.st>>>(d_rgb, rgb, n.out[l], act, b.C[l], n.ldh, Mc, n.out[l], n.in[l], 512,
                                                      gr->dW[l], n.ldw[l], gr->db[l]);
    LAUNCHED("colour head wgrad");
    float* P = b.dC[0];
    k_rowdot_dgrad<4><<<nblk(Mc * n.in[l]), 256, 0, c.st>>>(d_rgb, rgb, n.out[l], act, n.W[l], n.ldw[l], b.C[l], n.ldh, Mc,
                                                           n.out[l], n.in[l], P, n.ldh);
    LAUNCHED("colour head dgrad");
    int pp = 0;
    for (l = n.L - 2; l >= 0; --l) {
        const int64_t ldin = l == 0 ? g.ldx : n.ldh;
        RUN(wgrad(P, n.ldh, b.C[l], ldin, n.out[l], n.in[l], Mc, gr->dW[l], n.ldw[l], c.st));
        RUN(colsum(P, n.ldh, nullptr, 0, Mc, n.out[l], gr->db[l], c.st));
        if (l > 0) {
            float* Pn = b.dC[pp ^ 1];
            EpiBwdRelu e{b.C[l], n.ldh, Pn, n.ldh};
            RUN((launch<kNN>(P, n.ldh, n.W[l], n.ldw[l], Mc, n.in[l], n.out[l], 1, e, c.st, "colour dgrad")));
            P = Pn; pp ^= 1;
        } else {
            EpiColorIn e{g.nc, g.fc, c.cd->feat_dim, g.cc, c.cd->code_dim, b.dn_color, b.Dout, b.ldo, b.dcode, c.cd->code_dim};
            RUN((launch<kNN>(P, n.ldh, n.W[l], n.ldw[l], Mc, n.in[l], n.out[l], 1, e, c.st, "colour input dgrad")));
        }
    }
    return MSDF_OK;
}

int sdf_backward(const Ctx& c, const Bufs& b, const float* x, int64_t Mc, const msdf_mlp_grads* gr, float* grad_table) {
    const Net& n = c.sn;
    // ---- tangent sweep (adjoint of the reverse sweep) with the a_l^T t_l weight gradients
    const float* Tin = b.TG0; int64_t ldt = n.d0p;
    for (int l = 0; l < n.L - 1; ++l) {
        if (l == n.skip) {
            k_skip_copy<<<nblk(Mc * n.d0), 256, 0, c.st>>>(b.TG0, n.d0p, const_cast<float*>(Tin), ldt, Mc, n.d0, n.in[l] - n.d0, kInvSqrt2);
            LAUNCHED("tangent skip copy");
        }
        RUN(wgrad(b.A[l], n.ldh, Tin, ldt, n.out[l], n.in[l], Mc, gr->dW[l], n.ldw[l], c.st));
        float* Tout = b.T[(l + 1) & 1];
        EpiTan e{b.H[l + 1], n.ldh, in_scale(n, l + 1), b.A[l], n.ldh, Tout, n.ldh, l + 1 == n.skip ? kInvSqrt2 : 1.0f};
        RUN((launch<kNT>(Tin, ldt, n.W[l], n.ldw[l], Mc, n.out[l], n.in[l], 1, e, c.st, "sdf tangent layer")));
        Tin = Tout; ldt = n.ldh;
    }
    {
        const int l = n.L - 1;
        if (l == n.skip) {
            k_skip_copy<<<nblk(Mc * n.d0), 256, 0, c.st>>>(b.TG0, n.d0p, const_cast<float*>(Tin), ldt, Mc, n.d0, n.in[l] - n.d0, kInvSqrt2);
            LAUNCHED("tangent skip copy");
        }
        RUN(colsum(Tin, ldt, nullptr, 0, Mc, n.in[l], gr->dW[l], c.st));                       // a_{L-1} = e_0
        // ---- last layer of the backward sweep: pbar_{L-1} = Dout
        RUN(colsum(b.H[l], n.ldh, b.Dout, b.ldo, Mc, n.in[l], gr->dW[l], c.st));               // sdf row
        if (n.out[l] > 1) RUN(wgrad(b.Dout + 1, b.ldo, b.H[l], n.ldh, n.out[l] - 1, n.in[l], Mc, gr->dW[l] + n.ldw[l], n.ldw[l], c.st));
        RUN(colsum(b.Dout, b.ldo, nullptr, 0, Mc, n.out[l], gr->db[l], c.st));
    }
    const float* P = b.Dout; int64_t ldp = b.ldo;
    for (int l = n.L - 1; l >= 0; --l) {
        if (l < n.L - 1) {
            RUN(wgrad(P, ldp, b.H[l], l == 0 ? n.d0p : n.ldh, n.out[l], n.in[l], Mc, gr->dW[l], n.ldw[l], c.st));
            RUN(colsum(P, ldp, nullptr, 0, Mc, n.out[l], gr->db[l], c.st));
        }
        if (l == 0 && !(c.grid && grad_table)) break;
        EpiBwd e{};
        e.Hin = b.H[l]; e.ldh = l == 0 ? n.d0p : n.ldh; e.hscale = in_scale(n, l);
        e.PZ = l > 0 ? b.A[l - 1] : nullptr; e.ldp = n.ldh;
        e.bh0 = (c.grid && grad_table) ? b.BH0 : nullptr; e.ldb = n.d0p;
        e.dh = l == n.skip ? n.in[l] - n.d0 : n.in[l];
        e.qscale = l == n.skip ? kInvSqrt2 : 1.0f;
        e.layer0 = l == 0; e.bh0_accum = n.skip > 0;
        RUN((launch<kNN>(P, ldp, n.W[l], n.ldw[l], Mc, n.in[l], n.out[l], 1, e, c.st, "sdf backward layer")));
        if (l > 0) { P = b.A[l - 1]; ldp = n.ldh; }
    }
    if (c.grid && grad_table) {
        RUN(msdf_hash_scatter_rows(x, c.enc->offsets, Mc, c.enc->level_dim, c.enc->n_levels, c.enc->log2_per_level_scale,
                                   (uint32_t)c.enc->base_res, c.enc->divide_factor, b.BH0 + c.pe_w, n.d0p, b.G0 + c.pe_w, n.d0p,
                                   b.dn, c.hash_chain, grad_table, c.st));
    }
    return MSDF_OK;
}

int make_ctx(Ctx& c, const msdf_mlp_desc* sdf_net, const msdf_encoding_desc* enc, const msdf_mlp_desc* color_net,
             const msdf_color_desc* cd, float clamp_radius, float sphere_scale, void* stream, const char* who) {
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
        RUN(make_net(color_net, c.cn, who));
        c.cg = color_geom(cd);
        MSDF_CHECK_ARG(c.cn.skip < 0, "%s: colour net has no skip connection", who);
        MSDF_CHECK_ARG(c.cg.in0 == c.cn.d0, "%s: colour net d0=%d but inputs total %d", who, c.cn.d0, c.cg.in0);
        MSDF_CHECK_ARG(cd->feat_dim == c.sn.out[c.sn.L - 1] - 1, "%s: feat_dim mismatch", who);
        MSDF_CHECK_ARG(c.cd->multires_view <= 16, "%s: multires_view too large", who);
    }
    c.clamp_radius = clamp_radius; c.sphere_scale = sphere_scale;
    c.st = (cudaStream_t)stream;
    return MSDF_OK;
}

}  // namespace

// =============================================================================================================
// C ABI
// =============================================================================================================
extern "C" size_t msdf_field_workspace_bytes(const msdf_mlp_desc* sdf_net, const msdf_encoding_desc* enc,
                                             const msdf_mlp_desc* color_net, const msdf_color_desc* cd, int64_t chunk_points,
                                             int mode, unsigned flags) {
    (void)flags;
    Ctx c{};
    msdf_mlp_desc tmp_s = *sdf_net;
    static const float dummy = 0.f;
    for (int l = 0; l < tmp_s.n_layers && l < MSDF_MAX_LAYERS; ++l) { if (!tmp_s.W[l]) tmp_s.W[l] = &dummy; if (!tmp_s.b[l]) tmp_s.b[l] = &dummy; }
    msdf_mlp_desc tmp_c{};
    if (color_net) { tmp_c = *color_net; for (int l = 0; l < tmp_c.n_layers && l < MSDF_MAX_LAYERS; ++l) { if (!tmp_c.W[l]) tmp_c.W[l] = &dummy; if (!tmp_c.b[l]) tmp_c.b[l] = &dummy; } }
    if (make_ctx(c, &tmp_s, enc, color_net ? &tmp_c : nullptr, cd, 0.f, 1.f, nullptr, "msdf_field_workspace_bytes")) return 0;
    int64_t mc = (chunk_points + 127) / 128 * 128;
    if (mc < 128) mc = 128;
    return carve(c.sn, enc, c.has_color ? &c.cn : nullptr, cd, mc, mode, nullptr, nullptr);
}

extern "C" int msdf_field_forward(const msdf_mlp_desc* sdf_net, const msdf_encoding_desc* enc, const msdf_mlp_desc* color_net,
                                  const msdf_color_desc* cd, const float* x, int64_t M, const float* view_dirs, int64_t n_rays,
                                  int n_samples, const float* code, int mode, float clamp_radius, float sphere_scale,
                                  unsigned flags, void* workspace, size_t workspace_bytes, float* sdf, float* grad,
                                  float* feat, int64_t ld_feat, float* rgb, void* stream) {
    const char* who = "msdf_field_forward";
    MSDF_CHECK_ARG(mode == MSDF_MODE_SDF_ONLY || mode == MSDF_MODE_FORWARD, "%s: bad mode %d", who, mode);
    MSDF_CHECK_ARG((flags & MSDF_FLAG_TENSOR_BF16) == 0, "%s: tensor-core path not built into this entry point yet", who);
    if (mode == MSDF_MODE_SDF_ONLY) color_net = nullptr;
    Ctx c{};
    RUN(make_ctx(c, sdf_net, enc, color_net, cd, clamp_radius, sphere_scale, stream, who));
    if (M == 0) return MSDF_OK;
    MSDF_CHECK_ARG(x && workspace, "%s: null x / workspace", who);
    MSDF_CHECK_ARG(mode == MSDF_MODE_SDF_ONLY || grad != nullptr || !c.has_color, "%s: grad output required with a colour net", who);
    if (c.has_color) {
        MSDF_CHECK_ARG(view_dirs && rgb && n_samples > 0 && n_rays * (int64_t)n_samples == M, "%s: colour net needs view_dirs, rgb and M == n_rays*n_samples", who);
        MSDF_CHECK_ARG(cd->code_dim == 0 || code, "%s: per-image code missing", who);
    }
    const Net* cn = c.has_color ? &c.cn : nullptr;
    int64_t chunk = pick_chunk(c.sn, enc, cn, cd, M, mode, workspace_bytes);
    MSDF_CHECK_ARG(chunk > 0, "%s: workspace of %zu bytes is too small (need %zu for 128 points)", who, workspace_bytes,
                   carve(c.sn, enc, cn, cd, 128, mode, nullptr, nullptr));
    if (c.has_color && chunk < M) {   // keep chunks ray aligned
        chunk = chunk / n_samples * n_samples;
        MSDF_CHECK_ARG(chunk > 0, "%s: workspace too small for one ray of %d samples", who, n_samples);
    }
    Bufs b{};
    carve(c.sn, enc, cn, cd, (chunk + 127) / 128 * 128, mode, workspace, &b);
    const bool with_grad = mode == MSDF_MODE_FORWARD && (grad != nullptr);
    for (int64_t m0 = 0; m0 < M; m0 += chunk) {
        const int64_t Mc = (M - m0 < chunk) ? M - m0 : chunk;
        const float* xc = x + 3 * m0;
        RUN(encode_chunk(c, b, xc, Mc, with_grad));
        float* featc = feat ? feat + m0 * ld_feat : nullptr; int64_t ldf = ld_feat;
        if (c.has_color) {
            MSDF_CHECK_ARG(feat == nullptr, "%s: feat output and colour net are mutually exclusive", who);
            featc = b.X + c.cg.fc; ldf = c.cg.ldx;
        }
        RUN(forward_sweep(c, b, Mc, featc, ldf));
        if (with_grad) RUN(reverse_sweep(c, b, Mc));
        RUN(decode_chunk(c, b, xc, Mc, with_grad, sdf ? sdf + m0 : nullptr, with_grad ? grad + 3 * m0 : nullptr, nullptr));
        if (c.has_color) {
            // rays of this chunk: point m0+i belongs to ray (m0+i)/n_samples -> chunk must start on a ray boundary
            const int64_t ray0 = m0 / n_samples;
            MSDF_CHECK_ARG(m0 % n_samples == 0, "%s: internal: chunk not ray aligned", who);
            RUN(color_forward(c, b, xc, Mc, view_dirs + 3 * ray0, n_samples,
                              code ? code + (cd->code_per_ray ? ray0 * cd->code_dim : 0) : nullptr, grad + 3 * m0, rgb + (int64_t)c.cn.out[c.cn.L - 1] * m0));
        }
    }
    return MSDF_OK;
}

extern "C" int msdf_field_backward(const msdf_mlp_desc* sdf_net, const msdf_encoding_desc* enc, const msdf_mlp_desc* color_net,
                                   const msdf_color_desc* cd, const float* x, int64_t M, const float* view_dirs, int64_t n_rays,
                                   int n_samples, const float* code, float clamp_radius, float sphere_scale, unsigned flags,
                                   void* workspace, size_t workspace_bytes, const float* d_sdf, const float* d_grad,
                                   const float* d_feat, int64_t ld_dfeat, const float* rgb, const float* d_rgb,
                                   const msdf_mlp_grads* sdf_grads, const msdf_mlp_grads* color_grads, float* grad_table,
                                   float* d_code, void* stream) {
    const char* who = "msdf_field_backward";
    MSDF_CHECK_ARG((flags & MSDF_FLAG_TENSOR_BF16) == 0, "%s: tensor-core path not built into this entry point yet", who);
    if (d_rgb == nullptr) color_net = nullptr;
    Ctx c{};
    RUN(make_ctx(c, sdf_net, enc, color_net, cd, clamp_radius, sphere_scale, stream, who));
    if (M == 0) return MSDF_OK;
    MSDF_CHECK_ARG(x && workspace && sdf_grads, "%s: null x / workspace / sdf_grads", who);
    for (int l = 0; l < c.sn.L; ++l) MSDF_CHECK_ARG(sdf_grads->dW[l] && sdf_grads->db[l], "%s: sdf grad buffer %d missing", who, l);
    if (c.has_color) {
        MSDF_CHECK_ARG(view_dirs && rgb && color_grads && n_samples > 0 && n_rays * (int64_t)n_samples == M, "%s: colour net needs view_dirs, rgb, color_grads and M == n_rays*n_samples", who);
        for (int l = 0; l < c.cn.L; ++l) MSDF_CHECK_ARG(color_grads->dW[l] && color_grads->db[l], "%s: colour grad buffer %d missing", who, l);
        MSDF_CHECK_ARG(cd->code_dim == 0 || code, "%s: per-image code missing", who);
    }
    const Net* cn = c.has_color ? &c.cn : nullptr;
    int64_t chunk = pick_chunk(c.sn, enc, cn, cd, M, MSDF_MODE_BACKWARD, workspace_bytes);
    MSDF_CHECK_ARG(chunk > 0, "%s: workspace of %zu bytes is too small (need %zu for 128 points)", who, workspace_bytes,
                   carve(c.sn, enc, cn, cd, 128, MSDF_MODE_BACKWARD, nullptr, nullptr));
    if (c.has_color && chunk < M) {   // keep chunks ray aligned
        chunk = chunk / n_samples * n_samples;
        MSDF_CHECK_ARG(chunk > 0, "%s: workspace too small for one ray of %d samples", who, n_samples);
    }
    Bufs b{};
    carve(c.sn, enc, cn, cd, (chunk + 127) / 128 * 128, MSDF_MODE_BACKWARD, workspace, &b);
    for (int64_t m0 = 0; m0 < M; m0 += chunk) {
        const int64_t Mc = (M - m0 < chunk) ? M - m0 : chunk;
        const float* xc = x + 3 * m0;
        // ---- recompute the chunk
        RUN(encode_chunk(c, b, xc, Mc, true));
        RUN(forward_sweep(c, b, Mc, c.has_color ? b.X + c.cg.fc : nullptr, c.has_color ? c.cg.ldx : 0));
        RUN(reverse_sweep(c, b, Mc));
        RUN(decode_chunk(c, b, xc, Mc, true, b.sdfc, b.gradc, b.mask));
        if (c.has_color) {
            const int64_t ray0 = m0 / n_samples;
            const int no = c.cn.out[c.cn.L - 1];
            RUN(color_forward(c, b, xc, Mc, view_dirs + 3 * ray0, n_samples,
                              code ? code + (cd->code_per_ray ? ray0 * cd->code_dim : 0) : nullptr, b.gradc, nullptr));
            RUN(color_backward(c, b, Mc, rgb + (int64_t)no * m0, d_rgb + (int64_t)no * m0, color_grads));
            if (cd->code_dim > 0 && d_code) {
                if (cd->code_per_ray) {
                    const int64_t nr = Mc / n_samples;
                    k_code_grad<<<nblk(nr * cd->code_dim), 256, 0, c.st>>>(b.dcode, nr, n_samples, cd->code_dim, d_code + ray0 * cd->code_dim);
                    LAUNCHED("per-ray code gradient");
                } else {
                    RUN(colsum(b.dcode, cd->code_dim, nullptr, 0, Mc, cd->code_dim, d_code, c.st));
                }
            }
        }
        // ---- adjoints of (sdf, feat, grad) -> tangent of the encoded input
        {
            const int feat_w = c.sn.out[c.sn.L - 1] - 1;
            const int d0 = c.sn.d0;
            const int cols = d0 > feat_w + 1 ? d0 : feat_w + 1;
            k_backward_prologue<<<nblk(Mc * cols), 256, 0, c.st>>>(
                xc, Mc, c.pe_w, c.enc->grid_feat_dim, c.enc->n_levels, c.enc->level_dim, c.grid ? b.dydx : nullptr, c.hash_chain,
                b.mask, d_sdf ? d_sdf + m0 : nullptr, d_grad ? d_grad + 3 * m0 : nullptr, c.has_color ? b.dn_color : nullptr,
                d_feat ? d_feat + m0 * ld_dfeat : nullptr, ld_dfeat, feat_w, c.has_color ? 1 : 0, b.Dout, b.ldo, b.dn, b.TG0, c.sn.d0p);
            LAUNCHED("backward prologue");
        }
        RUN(sdf_backward(c, b, xc, Mc, sdf_grads, grad_table));
    }
    return MSDF_OK;
}

extern "C" int msdf_ray_points(const float* ray_o, const float* ray_d, const float* z, int64_t n_rays, int n, float* points,
                               void* stream) {
    MSDF_CHECK_ARG(ray_o && ray_d && z && points, "msdf_ray_points: null pointer");
    if (n_rays == 0 || n == 0) return MSDF_OK;
    k_ray_points<<<nblk(n_rays * n * 3), 256, 0, (cudaStream_t)stream>>>(ray_o, ray_d, z, n_rays, n, points);
    LAUNCHED("msdf_ray_points");
    return MSDF_OK;
}
