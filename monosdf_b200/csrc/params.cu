// Parameter-side kernels: weight normalisation (forward/backward) and the fused Adam step.
//
// Replaces nn.utils.weight_norm (dim=0) on every Linear of the two MLPs (reference code/model/network.py:72-73,
// 239-240, 383-384) and torch.optim.Adam (code/training/monosdf_train.py:210-221,432).
#include "common.cuh"

namespace {
constexpr unsigned kFull = 0xffffffffu;
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
    return v;
}

// d_table[idx[r], :] += d_code[r, :].  A training batch draws all its rays from one or a few images, so the indices are
// (nearly) all equal and a plain scatter serialises on the atomics of one row (torch's index backward takes 11 ms for
// 32768 rays).  One warp walks a segment of consecutive rays, lane = column, and only flushes when the index changes.
__global__ void k_code_scatter(const float* __restrict__ d_code, const int64_t* __restrict__ idx, int64_t n_rays, int cd,
                               int64_t table_rows, int rays_per_warp, float* __restrict__ d_table) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t r0 = w * rays_per_warp;
    if (r0 >= n_rays) return;
    const int64_t r1 = r0 + rays_per_warp < n_rays ? r0 + rays_per_warp : n_rays;
    for (int j = lane; j < cd; j += 32) {
        int64_t cur = idx[r0];
        float acc = 0.f;
        for (int64_t r = r0; r < r1; ++r) {
            const int64_t i = idx[r];
            if (i != cur) {
                if (cur >= 0 && cur < table_rows) atomicAdd(d_table + cur * cd + j, acc);
                cur = i; acc = 0.f;
            }
            acc += d_code[r * cd + j];
        }
        if (cur >= 0 && cur < table_rows) atomicAdd(d_table + cur * cd + j, acc);
    }
}

// one warp per output row: W[o,k] = g[o] v[o,k] / ||v[o,:]||, padding columns [in, ldw) zeroed
__global__ void k_weightnorm_forward(const float* __restrict__ g, const float* __restrict__ v, int out_dim, int in_dim,
                                     float* __restrict__ W, int ldw) {
    const int lane = threadIdx.x & 31;
    const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (o >= out_dim) return;
    const float* vr = v + (size_t)o * in_dim;
    float ss = 0.f;
    for (int k = lane; k < in_dim; k += 32) ss = fmaf(vr[k], vr[k], ss);
    ss = warp_sum(ss);
    const float s = g[o] / sqrtf(ss);
    for (int k = lane; k < ldw; k += 32) W[(size_t)o * ldw + k] = k < in_dim ? vr[k] * s : 0.f;
}

// dg[o] = <dW[o,:], v[o,:]> / ||v||;  dv[o,k] = g/||v|| (dW[o,k] - v[o,k] <dW,v> / ||v||^2)
__global__ void k_weightnorm_backward(const float* __restrict__ g, const float* __restrict__ v, const float* __restrict__ dW,
                                      int ldw, int out_dim, int in_dim, float* __restrict__ dg, float* __restrict__ dv) {
    const int lane = threadIdx.x & 31;
    const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (o >= out_dim) return;
    const float* vr = v + (size_t)o * in_dim;
    const float* dr = dW + (size_t)o * ldw;
    float ss = 0.f, dot = 0.f;
    for (int k = lane; k < in_dim; k += 32) { ss = fmaf(vr[k], vr[k], ss); dot = fmaf(dr[k], vr[k], dot); }
    ss = warp_sum(ss); dot = warp_sum(dot);
    const float nrm = sqrtf(ss);
    if (lane == 0) dg[o] = dot / nrm;
    const float a = g[o] / nrm, bcoef = dot / ss;
    for (int k = lane; k < in_dim; k += 32) dv[(size_t)o * in_dim + k] = a * (dr[k] - vr[k] * bcoef);
}

__global__ void k_adam(float* __restrict__ p, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                       int64_t n, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt,
                       float grad_scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gi = grad[i] * grad_scale;
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
}
}  // namespace

extern "C" int msdf_weightnorm_forward(const float* g, const float* v, int out_dim, int in_dim, float* W, int ldw, void* stream) {
    MSDF_CHECK_ARG(g && v && W && out_dim > 0 && in_dim > 0 && ldw >= in_dim, "msdf_weightnorm_forward: bad arguments");
    k_weightnorm_forward<<<(unsigned)msdf_div_up(out_dim, 4), 128, 0, (cudaStream_t)stream>>>(g, v, out_dim, in_dim, W, ldw);
    MSDF_COUNT_LAUNCH();
    MSDF_CHECK_LAUNCH("msdf_weightnorm_forward");
    return MSDF_OK;
}

extern "C" int msdf_weightnorm_backward(const float* g, const float* v, const float* dW, int ldw, int out_dim, int in_dim,
                                        float* dg, float* dv, void* stream) {
    MSDF_CHECK_ARG(g && v && dW && dg && dv && out_dim > 0 && in_dim > 0 && ldw >= in_dim, "msdf_weightnorm_backward: bad arguments");
    k_weightnorm_backward<<<(unsigned)msdf_div_up(out_dim, 4), 128, 0, (cudaStream_t)stream>>>(g, v, dW, ldw, out_dim, in_dim, dg, dv);
    MSDF_COUNT_LAUNCH();
    MSDF_CHECK_LAUNCH("msdf_weightnorm_backward");
    return MSDF_OK;
}

extern "C" int msdf_fused_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                               float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                               void* stream) {
    MSDF_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && step >= 1, "msdf_fused_adam: bad arguments");
    if (n == 0) return MSDF_OK;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    k_adam<<<(unsigned)msdf_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                            weight_decay, (float)bc1, (float)sqrt(bc2), grad_scale);
    MSDF_COUNT_LAUNCH();
    MSDF_CHECK_LAUNCH("msdf_fused_adam");
    return MSDF_OK;
}


extern "C" int msdf_code_scatter(const float* d_code, const int64_t* indices, int64_t n_rays, int code_dim, int64_t table_rows,
                                 float* d_table, void* stream) {
    if (n_rays == 0) return MSDF_OK;
    MSDF_CHECK_ARG(d_code && indices && d_table && code_dim > 0 && table_rows > 0, "msdf_code_scatter: bad arguments");
    const int rpw = 64;
    const int64_t warps = msdf_div_up(n_rays, rpw);
    k_code_scatter<<<(unsigned)msdf_div_up(warps, 8), 256, 0, (cudaStream_t)stream>>>(d_code, indices, n_rays, code_dim, table_rows, rpw, d_table);
    MSDF_COUNT_LAUNCH();
    MSDF_CHECK_LAUNCH("msdf_code_scatter");
    return MSDF_OK;
}
