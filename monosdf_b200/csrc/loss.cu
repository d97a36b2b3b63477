// MonoSDFLoss forward + backward in three launches.
//
// Replaces reference code/model/loss.py:180-311 (MonoSDFLoss.forward) with its helpers: compute_scale_and_shift_1D
// (:29-49), get_rgb_loss (:217-220), get_eikonal_loss (:222-224), get_smooth_loss (:226-234), get_depth_loss (:236-243),
// get_normal_loss (:245-250) and ScaleAndShiftInvariantLoss / mse_loss in pixel-batch mode (:52-66, :150-176).
// The reference spends ~40 small kernels, boolean-mask indexing and two host synchronisations here; this file is
//   pass 1   warp per ray: foreground mask from the ray's sdf row (any > 0 and any < 0, :274) and every plain sum
//            (rgb, normal L1 / cosine, the five normal-equation sums of the scale/shift fit); thread per point:
//            eikonal and smoothness sums
//   pass 2   thread per ray / per point: scale and shift from the sums, the depth residual sum, and ALL gradients
//   final    one thread: the seven scalars
// The least-squares (scale, shift) minimise the depth residual, so d num / d depth = 2 m res scale exactly (the terms
// through scale and shift vanish at the optimum) -- what autograd computes for the reference up to rounding.
// HBM-bound: per ray it reads 4 S + 52 bytes and writes 28; per eikonal point 24 in, 24 out.
#include "common.cuh"

namespace {

constexpr unsigned kFull = 0xffffffffu;
enum { A00 = 0, A01, A11, B0, B1, RGB, NL1, NCOS, EIK, SMOOTH, DNUM, NACC };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
    return v;
}

// block-level accumulation of NACC partial sums into global (one atomic per quantity per block)
__device__ __forceinline__ void block_accumulate(float (&part)[NACC], double* __restrict__ acc) {
    __shared__ float sh[NACC][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
        const float s = warp_sum(part[k]);
        if (lane == 0) sh[k][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < NACC) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[threadIdx.x][w];
        if (s != 0.f) atomicAdd(acc + threadIdx.x, (double)s);     // fp64 totals: the scale/shift determinant cancels
    }
}

__device__ __forceinline__ float gamma2(float x) {           // loss.py:209-215
    return x <= 0.0031308f ? 12.92f * x : 1.055f * powf(x, 1.0f / 2.4f) - 0.055f;
}
__device__ __forceinline__ float gamma2_grad(float x) {
    return x <= 0.0031308f ? 12.92f : 1.055f / 2.4f * powf(x, 1.0f / 2.4f - 1.0f);
}

struct Rays {
    const float *rgb, *rgb_gt, *depth, *depth_gt, *gt_mask, *normal, *normal_gt, *sdf;
    int64_t n; int S;
};

// F.normalize(v, p=2, dim=-1): v / max(|v|, 1e-12)
__device__ __forceinline__ float safe_norm(const float v[3]) { return fmaxf(sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), 1e-12f); }

__global__ void __launch_bounds__(256)
k_loss_pass1(Rays r, msdf_loss_desc d, int64_t n_eik, const float* __restrict__ g1, const float* __restrict__ g2,
             float* __restrict__ maskbuf, double* __restrict__ acc) {
    float part[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) part[k] = 0.f;
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); ray < r.n; ray += warps) {
        bool pos = false, neg = false;
        for (int j = lane; j < r.S; j += 32) { const float s = r.sdf[ray * r.S + j]; pos |= s > 0.f; neg |= s < 0.f; }
        const bool fg = __any_sync(kFull, pos) && __any_sync(kFull, neg);
        if (lane == 0) {
            const float m = (fg && r.gt_mask[ray] > 0.5f) ? 1.f : 0.f;
            maskbuf[ray] = m;
            const float p = r.depth[ray];
            const float t = d.scale_invariant_depth ? r.depth_gt[ray] * 50.f + 0.5f : r.depth_gt[ray];
            part[A00] += m * p * p; part[A01] += m * p; part[A11] += m; part[B0] += m * p * t; part[B1] += m * t;
            float e = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float a = r.rgb[3 * ray + c], b = r.rgb_gt[3 * ray + c];
                if (d.gamma) { a = gamma2(a); b = gamma2(b); }
                e += d.rgb_mse ? (a - b) * (a - b) : fabsf(a - b);
            }
            part[RGB] += e;
            float v[3] = {r.normal[3 * ray] * m, r.normal[3 * ray + 1] * m, r.normal[3 * ray + 2] * m};
            float gt[3] = {r.normal_gt[3 * ray], r.normal_gt[3 * ray + 1], r.normal_gt[3 * ray + 2]};
            const float iv = 1.f / safe_norm(v), ig = 1.f / safe_norm(gt);
            float l1 = 0.f, dot = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) { const float a = v[c] * iv, b = gt[c] * ig; l1 += fabsf(a - b); dot += a * b; }
            part[NL1] += l1; part[NCOS] += 1.f - dot;
        }
    }
    const int64_t threads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_eik; i += threads) {
        const float a[3] = {g1[3 * i], g1[3 * i + 1], g1[3 * i + 2]}, b[3] = {g2[3 * i], g2[3 * i + 1], g2[3 * i + 2]};
        const float na = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]), nb = sqrtf(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
        part[EIK] += (na - 1.f) * (na - 1.f);
        float q = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) { const float x = a[c] / (na + 1e-5f) - b[c] / (nb + 1e-5f); q += x * x; }
        part[SMOOTH] += sqrtf(q);
    }
    block_accumulate(part, acc);
}

__device__ __forceinline__ void scale_shift(const double* acc, const msdf_loss_desc& d, float& s, float& sh) {
    s = 1.f; sh = 0.f;
    if (!d.scale_invariant_depth) return;
    const double det = acc[A00] * acc[A11] - acc[A01] * acc[A01];
    if (det != 0.0) { s = (float)((acc[A11] * acc[B0] - acc[A01] * acc[B1]) / det); sh = (float)((-acc[A01] * acc[B0] + acc[A00] * acc[B1]) / det); }
    else { s = 0.f; sh = 0.f; }
}

__global__ void __launch_bounds__(256)
k_loss_pass2(Rays r, msdf_loss_desc d, int64_t n_eik, const float* __restrict__ g1, const float* __restrict__ g2,
             const float* __restrict__ maskbuf, double* __restrict__ acc, float* __restrict__ d_rgb, float* __restrict__ d_depth,
             float* __restrict__ d_normal, float* __restrict__ d_g1, float* __restrict__ d_g2) {
    float part[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) part[k] = 0.f;
    float s, sh;
    scale_shift(acc, d, s, sh);
    const float div = 2.f * (float)acc[A11];
    const float inv_n = 1.f / (float)r.n;
    const int64_t threads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; ray < r.n; ray += threads) {
        const float m = maskbuf[ray];
        const float p = r.depth[ray];
        const float t = d.scale_invariant_depth ? r.depth_gt[ray] * 50.f + 0.5f : r.depth_gt[ray];
        const float res = s * p + sh - t;
        part[DNUM] += m * res * res;
        // depth_loss = num / div  (0 when the mask is empty, loss.py:57-60)
        d_depth[ray] = div > 0.f ? d.decay * d.depth_weight * (2.f * m * res * s) / div : 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float a = r.rgb[3 * ray + c], b = r.rgb_gt[3 * ray + c], ga = 1.f;
            if (d.gamma) { ga = gamma2_grad(a); a = gamma2(a); b = gamma2(b); }
            const float e = a - b;
            const float g = d.rgb_mse ? 2.f * e : (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f));
            d_rgb[3 * ray + c] = g * ga * inv_n * (1.f / 3.f);                   // mean over N x 3 elements
        }
        float v[3] = {r.normal[3 * ray] * m, r.normal[3 * ray + 1] * m, r.normal[3 * ray + 2] * m};
        float gt[3] = {r.normal_gt[3 * ray], r.normal_gt[3 * ray + 1], r.normal_gt[3 * ray + 2]};
        const float nv = safe_norm(v), ig = 1.f / safe_norm(gt);
        float u[3], gn[3], dot = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            u[c] = v[c] / nv;
            const float b = gt[c] * ig, e = u[c] - b;
            // d/du of  w_l1 sum_c |u - b| + w_cos (1 - u.b), both averaged over the rays
            gn[c] = d.decay * inv_n * (d.normal_l1_weight * (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f)) - d.normal_cos_weight * b);
            dot += gn[c] * u[c];
        }
        // through u = v / max(|v|, eps): (g - u (u.g)) / |v| when |v| > eps, g / eps otherwise; then v = normal * m
        const bool tiny = nv <= 1e-12f;
#pragma unroll
        for (int c = 0; c < 3; ++c) d_normal[3 * ray + c] = m * (tiny ? gn[c] / 1e-12f : (gn[c] - u[c] * dot) / nv);
    }
    const float inv_e = n_eik > 0 ? 1.f / (float)n_eik : 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_eik; i += threads) {
        const float a[3] = {g1[3 * i], g1[3 * i + 1], g1[3 * i + 2]}, b[3] = {g2[3 * i], g2[3 * i + 1], g2[3 * i + 2]};
        const float na = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]), nb = sqrtf(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
        float ua[3], ub[3], df[3], q = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) { ua[c] = a[c] / (na + 1e-5f); ub[c] = b[c] / (nb + 1e-5f); df[c] = ua[c] - ub[c]; q += df[c] * df[c]; }
        const float nq = sqrtf(q);
        // smooth = mean |ua - ub|: d/dua = df / |df| (0 at df = 0, like torch.norm's subgradient)
        float gs[3], da = 0.f, db = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) { gs[c] = nq > 0.f ? d.smooth_weight * inv_e * df[c] / nq : 0.f; da += gs[c] * a[c]; db += gs[c] * b[c]; }
        // u = g / (|g| + eps):  du/dg . gs = gs / (|g| + eps) - g (g.gs) / (|g| (|g| + eps)^2)
        const float ea = na + 1e-5f, eb = nb + 1e-5f;
        const float eik = na > 0.f ? d.eikonal_weight * inv_e * 2.f * (na - 1.f) / na : 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            d_g1[3 * i + c] = eik * a[c] + gs[c] / ea - (na > 0.f ? a[c] * da / (na * ea * ea) : 0.f);
            d_g2[3 * i + c] = -(gs[c] / eb - (nb > 0.f ? b[c] * db / (nb * eb * eb) : 0.f));
        }
    }
    block_accumulate(part, acc);
}

__global__ void k_loss_final(const double* __restrict__ acc, msdf_loss_desc d, int64_t n_rays, int64_t n_eik, float* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double inv_n = 1.0 / (double)n_rays;
    const float rgb = (float)(acc[RGB] * inv_n / 3.0);
    const float eik = n_eik > 0 ? (float)(acc[EIK] / (double)n_eik) : 0.f;
    const float smooth = n_eik > 0 ? (float)(acc[SMOOTH] / (double)n_eik) : 0.f;
    const double div = 2.0 * acc[A11];
    const float depth = div > 0.0 ? (float)(acc[DNUM] / div) : 0.f;
    const float nl1 = (float)(acc[NL1] * inv_n), ncos = (float)(acc[NCOS] * inv_n);
    out[0] = rgb + d.eikonal_weight * eik + d.smooth_weight * smooth +
             d.decay * (d.depth_weight * depth + d.normal_l1_weight * nl1 + d.normal_cos_weight * ncos);
    out[1] = rgb; out[2] = eik; out[3] = smooth; out[4] = depth; out[5] = nl1; out[6] = ncos; out[7] = (float)acc[A11];
}

}  // namespace

extern "C" int msdf_loss_forward_backward(const msdf_loss_desc* desc, int64_t n_rays, int n_samples, const float* rgb,
                                          const float* rgb_gt, const float* depth, const float* depth_gt, const float* gt_mask,
                                          const float* normal, const float* normal_gt, const float* sdf, int64_t n_eik,
                                          const float* g1, const float* g2, float* workspace, float* out, float* d_rgb,
                                          float* d_depth, float* d_normal, float* d_g1, float* d_g2, void* stream) {
    MSDF_CHECK_ARG(desc != nullptr && n_rays > 0 && n_samples > 0, "msdf_loss_forward_backward: bad sizes");
    MSDF_CHECK_ARG(rgb && rgb_gt && depth && depth_gt && gt_mask && normal && normal_gt && sdf && workspace && out && d_rgb && d_depth && d_normal,
                   "msdf_loss_forward_backward: null pointer");
    MSDF_CHECK_ARG(n_eik == 0 || (g1 && g2 && d_g1 && d_g2), "msdf_loss_forward_backward: eikonal buffers missing");
    cudaStream_t st = (cudaStream_t)stream;
    MSDF_CHECK_ARG((((uintptr_t)workspace) & 7) == 0, "msdf_loss_forward_backward: workspace must be 8-byte aligned");
    double* acc = reinterpret_cast<double*>(workspace);     // 16 fp64 sums (32 floats), then [n_rays] mask
    float* maskbuf = workspace + 32;
    MSDF_CUDA_CALL(cudaMemsetAsync(acc, 0, 16 * sizeof(double), st));
    Rays r{rgb, rgb_gt, depth, depth_gt, gt_mask, normal, normal_gt, sdf, n_rays, n_samples};
    const int64_t work = n_rays > n_eik / 8 ? n_rays : n_eik / 8;
    int blocks = (int)msdf_div_up(work, 8);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_loss_pass1<<<blocks, 256, 0, st>>>(r, *desc, n_eik, g1, g2, maskbuf, acc);
    MSDF_COUNT_LAUNCH(); MSDF_CHECK_LAUNCH("msdf_loss (pass 1)");
    int blocks2 = (int)msdf_div_up(n_rays > n_eik ? n_rays : n_eik, 256);
    if (blocks2 > 148 * 8) blocks2 = 148 * 8;
    k_loss_pass2<<<blocks2, 256, 0, st>>>(r, *desc, n_eik, g1, g2, maskbuf, acc, d_rgb, d_depth, d_normal, d_g1, d_g2);
    MSDF_COUNT_LAUNCH(); MSDF_CHECK_LAUNCH("msdf_loss (pass 2)");
    k_loss_final<<<1, 32, 0, st>>>(acc, *desc, n_rays, n_eik, out);
    MSDF_COUNT_LAUNCH(); MSDF_CHECK_LAUNCH("msdf_loss (final)");
    return MSDF_OK;
}
