// Error-bounded ray sampler kernels: one warp per ray, warp-scan prefix sums.
//
// Replaces the per-iteration body of ErrorBoundSampler.get_z_vals
// (reference code/model/ray_sampler.py:110-262; UniformSampler :48-83; get_error_bound :264-272)
// by three phases with the SDF evaluation (the MLP) between them:
//   init      -> cube far, 128 initial samples, Lemma-2 beta bound, sample points
//   round     -> merge new samples, d* (Theorem 1), beta bisection, batch-global "not converged" flag
//   upsample  -> inverse-CDF samples proportional to the error bound  (loop continues)
//   finalize  -> N_samples from the opacity pdf + near/far/extra picks, sorted      (loop ends)
//
// Arithmetic contract: this file is compiled with -fmad=false and uses only the IEEE-exact helpers in
// include/msdf_detmath.h; prefix sums use the "row-wise warp order" (Kogge-Stone inside each block of 32
// consecutive elements plus a running carry).  oracle/sampler_oracle.c restates the same order on the CPU,
// so sample positions are bit-identical between the two.
//
// Work per ray-round is ~25 scans over <=640 elements and ~3*11*n software exps: SFU/issue bound, not HBM
// bound; global traffic is one read of z/sdf and one write of the merged row.
#include "common.cuh"
#include "../../include/msdf_detmath.h"

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float ks_scan(float v, int lane) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        float t = __shfl_up_sync(kFull, v, off);
        if (lane >= off) v = v + t;
    }
    return v;
}

// total of n elements held in smem, row-wise warp order
__device__ float sum_rowwise(const float* a, int n, int lane) {
    float carry = 0.0f;
    for (int base = 0; base < n; base += 32) {
        float v = (base + lane < n) ? a[base + lane] : 0.0f;
        v = ks_scan(v, lane);
        carry = carry + __shfl_sync(kFull, v, 31);
    }
    return carry;
}

__device__ __forceinline__ float warp_max_nonan(float m) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        float o = __shfl_xor_sync(kFull, m, off);
        m = (o > m) ? o : m;
    }
    return m;
}

// get_error_bound for one ray (ray_sampler.py:264-272); sdf/dists/dstar in shared memory
__device__ float error_bound(const float* sdf, const float* dists, const float* dstar, int n, float beta, int lane) {
    float carry_i = 0.0f, carry_e = 0.0f, m = -INFINITY;
    for (int base = 0; base < n - 1; base += 32) {
        int i = base + lane;
        bool ok = i < n - 1;
        float sfe = (ok && i > 0) ? dists[i - 1] * msdf_density(sdf[i - 1], beta) : 0.0f;
        float es = ok ? msdf_err_section(dstar[i], dists[i], beta) : 0.0f;
        float integ = carry_i + ks_scan(sfe, lane);
        float eint = carry_e + ks_scan(es, lane);
        carry_i = __shfl_sync(kFull, integ, 31);
        carry_e = __shfl_sync(kFull, eint, 31);
        if (ok) {
            float bo = msdf_bound_opacity(eint, msdf_expf(-integ));
            m = (bo > m) ? bo : m;
        }
    }
    return warp_max_nonan(m);
}

__device__ float cube_far(const float* o, const float* d, float bound, float far_max) {
    float nearv = -INFINITY, farv = INFINITY;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float den = d[k] + 1e-15f;
        float tmin = (-bound - o[k]) / den;
        float tmax = (bound - o[k]) / den;
        float lo = (tmin < tmax) ? tmin : tmax;
        float hi = (tmin > tmax) ? tmin : tmax;
        nearv = (lo > nearv) ? lo : nearv;
        farv = (hi < farv) ? hi : farv;
    }
    if (farv < nearv) farv = 1e9f;
    return (farv > far_max) ? far_max : farv;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_sampler_init(const float* __restrict__ ray_o, const float* __restrict__ ray_d, int64_t n_rays,
               const float* __restrict__ t_vals, const float* __restrict__ t_rand, int n0,
               float bound, float nearv, float far_max, float beta_coef,
               float* __restrict__ z, int cap, float* __restrict__ beta, float* __restrict__ pts) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
    if (r >= n_rays) return;
    float* zs = smem + (size_t)warp * 2 * n0;
    float* tmp = zs + n0;
    float o[3], d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { o[k] = ray_o[3 * r + k]; d[k] = ray_d[3 * r + k]; }
    const float farv = cube_far(o, d, bound, far_max);
    for (int j = lane; j < n0; j += 32) zs[j] = nearv * (1.0f - t_vals[j]) + farv * t_vals[j];
    __syncwarp();
    if (t_rand != nullptr) {
        for (int j = lane; j < n0; j += 32) {
            float lower = (j == 0) ? zs[0] : 0.5f * (zs[j] + zs[j - 1]);
            float upper = (j == n0 - 1) ? zs[n0 - 1] : 0.5f * (zs[j + 1] + zs[j]);
            tmp[j] = lower + (upper - lower) * t_rand[(int64_t)n0 * r + j];
        }
        __syncwarp();
        for (int j = lane; j < n0; j += 32) zs[j] = tmp[j];
        __syncwarp();
    }
    for (int j = lane; j < n0 - 1; j += 32) { float dd = zs[j + 1] - zs[j]; tmp[j] = dd * dd; }
    __syncwarp();
    float total = sum_rowwise(tmp, n0 - 1, lane);
    if (lane == 0) beta[r] = sqrtf(beta_coef * total);
    float* zr = z + (int64_t)cap * r;
    for (int j = lane; j < n0; j += 32) {
        float zj = zs[j];
        zr[j] = zj;
        float* p = pts + ((int64_t)n0 * r + j) * 3;
        p[0] = o[0] + zj * d[0]; p[1] = o[1] + zj * d[1]; p[2] = o[2] + zj * d[2];
    }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_sampler_round(int64_t n_rays, int n_old, int n_new, float* __restrict__ z, float* __restrict__ sdf,
                const float* __restrict__ z_new, const float* __restrict__ sdf_new, int cap,
                const float* __restrict__ beta0_p, float eps, int beta_iters, float* __restrict__ beta, unsigned int* __restrict__ flag) {
    extern __shared__ float smem[];
    const float beta0 = *beta0_p;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
    if (r >= n_rays) return;
    const int n = n_old + n_new;
    float* zs = smem + (size_t)warp * 4 * cap;   // merged z
    float* ss = zs + cap;                        // merged sdf
    float* a0 = ss + cap;                        // scratch: old z   -> dists
    float* a1 = a0 + cap;                        // scratch: old sdf -> dstar
    float* zr = z + (int64_t)cap * r;
    float* sr = sdf + (int64_t)cap * r;
    if (n_old > 0) {
        const float* zn = z_new + (int64_t)n_new * r;
        const float* sn = sdf_new + (int64_t)n_new * r;
        float* nz = a0 + n_old;   // staged new z (cap >= n_old + n_new)
        float* ns = a1 + n_old;
        for (int i = lane; i < n_old; i += 32) { a0[i] = zr[i]; a1[i] = sr[i]; }
        for (int j = lane; j < n_new; j += 32) { nz[j] = zn[j]; ns[j] = sn[j]; }
        __syncwarp();
        for (int i = lane; i < n_old; i += 32) {            // stable merge: old entries first on ties
            float zi = a0[i];
            int c = 0;
            for (int k = 0; k < n_new; ++k) c += (nz[k] < zi);
            zs[i + c] = zi; ss[i + c] = a1[i];
        }
        for (int j = lane; j < n_new; j += 32) {
            float zj = nz[j];
            int lo = 0, hi = n_old;                         // count of old <= zj (old is sorted)
            while (lo < hi) { int mid = (lo + hi) >> 1; if (a0[mid] <= zj) lo = mid + 1; else hi = mid; }
            int c = lo;
            for (int k = 0; k < n_new; ++k) c += (nz[k] < zj) || (nz[k] == zj && k < j);
            zs[c] = zj; ss[c] = ns[j];
        }
        __syncwarp();
        for (int i = lane; i < n; i += 32) { zr[i] = zs[i]; sr[i] = ss[i]; }
    } else {
        const float* sn = sdf_new + (int64_t)n_new * r;
        for (int i = lane; i < n; i += 32) { zs[i] = zr[i]; float s = sn[i]; ss[i] = s; sr[i] = s; }
    }
    __syncwarp();
    float* dists = a0; float* dstar = a1;
    for (int i = lane; i < n - 1; i += 32) {
        float dd = zs[i + 1] - zs[i];
        dists[i] = dd;
        dstar[i] = msdf_dstar(dd, ss[i], ss[i + 1]);
    }
    __syncwarp();
    float b = beta[r];
    float err = error_bound(ss, dists, dstar, n, beta0, lane);
    if (err <= eps) b = beta0;
    float bmin = beta0, bmax = b;
    for (int it = 0; it < beta_iters; ++it) {
        float mid = (bmin + bmax) / 2.0f;
        err = error_bound(ss, dists, dstar, n, mid, lane);
        if (err <= eps) bmax = mid;
        if (err > eps) bmin = mid;
    }
    if (lane == 0) {
        beta[r] = bmax;
        if (bmax > beta0) atomicOr(flag, 1u);
    }
}

// pdf -> cdf -> inverse-CDF samples (ray_sampler.py:168-228). zs/ss: merged row in smem; pdf/cdf scratch (cap+1 each)
__device__ void draw_samples(const float* zs, const float* ss, int n, float beta, bool upsample, float add_tiny,
                             const float* __restrict__ u, int n_u, float* out, float* pdf, float* cdf, int lane) {
    float carry_i = 0.0f, carry_e = 0.0f;
    for (int base = 0; base < n; base += 32) {
        int i = base + lane;
        float dprev = (i > 0 && i < n) ? zs[i] - zs[i - 1] : 0.0f;
        float sfe = (i > 0 && i < n) ? dprev * msdf_density(ss[i - 1], beta) : 0.0f;
        float integ = carry_i + ks_scan(sfe, lane);
        carry_i = __shfl_sync(kFull, integ, 31);
        bool ok = i < n - 1;
        float di = ok ? zs[i + 1] - zs[i] : 0.0f;
        if (upsample) {
            float es = ok ? msdf_err_section(msdf_dstar(di, ss[i], ss[i + 1]), di, beta) : 0.0f;
            float eint = carry_e + ks_scan(es, lane);
            carry_e = __shfl_sync(kFull, eint, 31);
            if (ok) pdf[i] = msdf_bound_opacity(eint, msdf_expf(-integ)) + add_tiny;
        } else if (ok) {
            float fe = di * msdf_density(ss[i], beta);
            float alpha = 1.0f - msdf_expf(-fe);
            pdf[i] = alpha * msdf_expf(-integ) + 1e-5f;
        }
    }
    __syncwarp();
    const float total = sum_rowwise(pdf, n - 1, lane);
    float carry = 0.0f;
    if (lane == 0) cdf[0] = 0.0f;
    for (int base = 0; base < n - 1; base += 32) {
        int i = base + lane;
        float v = (i < n - 1) ? pdf[i] / total : 0.0f;
        float c = carry + ks_scan(v, lane);
        carry = __shfl_sync(kFull, c, 31);
        if (i < n - 1) cdf[i + 1] = c;
    }
    __syncwarp();
    for (int j = lane; j < n_u; j += 32) {
        float uj = u[j];
        int lo = 0, hi = n;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (cdf[mid] <= uj) lo = mid + 1; else hi = mid; }
        int below = (lo - 1 > 0) ? lo - 1 : 0;
        int above = (lo < n - 1) ? lo : n - 1;
        float denom = cdf[above] - cdf[below];
        if (denom < 1e-5f) denom = 1.0f;
        float t = (uj - cdf[below]) / denom;
        out[j] = zs[below] + t * (zs[above] - zs[below]);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_sampler_upsample(int64_t n_rays, int n, const float* __restrict__ z, const float* __restrict__ sdf, int cap,
                   const float* __restrict__ beta, float add_tiny, const float* __restrict__ u, int n_new,
                   const float* __restrict__ ray_o, const float* __restrict__ ray_d,
                   float* __restrict__ z_new, float* __restrict__ pts_new) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
    if (r >= n_rays) return;
    const int stride = 4 * (cap + 1) + n_new;
    float* zs = smem + (size_t)warp * stride;
    float* ss = zs + cap + 1; float* pdf = ss + cap + 1; float* cdf = pdf + cap + 1; float* out = cdf + cap + 1;
    for (int i = lane; i < n; i += 32) { zs[i] = z[(int64_t)cap * r + i]; ss[i] = sdf[(int64_t)cap * r + i]; }
    __syncwarp();
    draw_samples(zs, ss, n, beta[r], true, add_tiny, u, n_new, out, pdf, cdf, lane);
    float o[3], d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { o[k] = ray_o[3 * r + k]; d[k] = ray_d[3 * r + k]; }
    for (int j = lane; j < n_new; j += 32) {
        float zj = out[j];
        z_new[(int64_t)n_new * r + j] = zj;
        float* p = pts_new + ((int64_t)n_new * r + j) * 3;
        p[0] = o[0] + zj * d[0]; p[1] = o[1] + zj * d[1]; p[2] = o[2] + zj * d[2];
    }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_sampler_finalize(int64_t n_rays, int n, const float* __restrict__ z, const float* __restrict__ sdf, int cap,
                   const float* __restrict__ beta, const float* __restrict__ u, int u_per_ray, int n_s,
                   const int32_t* __restrict__ pick, int n_extra, float nearv, float farv,
                   const int64_t* __restrict__ eik_idx, float* __restrict__ z_out, float* __restrict__ z_eik) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
    if (r >= n_rays) return;
    const int n_out = n_s + 2 + n_extra;
    const int stride = 4 * (cap + 1) + 2 * n_out;
    float* zs = smem + (size_t)warp * stride;
    float* ss = zs + cap + 1; float* pdf = ss + cap + 1; float* cdf = pdf + cap + 1;
    float* cat = cdf + cap + 1; float* srt = cat + n_out;
    for (int i = lane; i < n; i += 32) { zs[i] = z[(int64_t)cap * r + i]; ss[i] = sdf[(int64_t)cap * r + i]; }
    __syncwarp();
    draw_samples(zs, ss, n, beta[r], false, 0.0f, u_per_ray ? u + (int64_t)n_s * r : u, n_s, cat, pdf, cdf, lane);
    if (lane == 0) { cat[n_s] = nearv; cat[n_s + 1] = farv; }
    for (int k = lane; k < n_extra; k += 32) cat[n_s + 2 + k] = zs[pick[k]];
    __syncwarp();
    for (int i = lane; i < n_out; i += 32) {   // stable rank sort (torch.sort, ray_sampler.py:251)
        float v = cat[i];
        int c = 0;
        for (int k = 0; k < n_out; ++k) c += (cat[k] < v) || (cat[k] == v && k < i);
        srt[c] = v;
    }
    __syncwarp();
    for (int i = lane; i < n_out; i += 32) z_out[(int64_t)n_out * r + i] = srt[i];
    if (z_eik != nullptr && lane == 0) z_eik[r] = srt[eik_idx[r]];
}

template <typename K>
int set_smem(K kernel, size_t bytes, const char* name) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) { msdf_set_error("%s: cannot opt in to %zu B shared memory: %s", name, bytes, cudaGetErrorString(e)); return MSDF_ERR_CUDA; }
    }
    return MSDF_OK;
}

}  // namespace

extern "C" int msdf_sampler_init(const float* ray_o, const float* ray_d, int64_t n_rays, const float* t_vals,
                                 const float* t_rand, int n0, float bound, float near_, float far_max,
                                 float beta_coef, float* z, int cap, float* beta, float* pts, void* stream) {
    if (n_rays == 0) return MSDF_OK;
    MSDF_CHECK_ARG(ray_o && ray_d && t_vals && z && beta && pts, "msdf_sampler_init: null pointer");
    MSDF_CHECK_ARG(n0 >= 2 && cap >= n0, "msdf_sampler_init: need 2 <= n0 <= cap (n0=%d cap=%d)", n0, cap);
    size_t smem = (size_t)kWarpsPerBlock * 2 * n0 * sizeof(float);
    int rc = set_smem(k_sampler_init, smem, "msdf_sampler_init"); if (rc) return rc;
    k_sampler_init<<<(unsigned)msdf_div_up(n_rays, kWarpsPerBlock), kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
        ray_o, ray_d, n_rays, t_vals, t_rand, n0, bound, near_, far_max, beta_coef, z, cap, beta, pts);
    MSDF_COUNT_LAUNCH();
    MSDF_CHECK_LAUNCH("msdf_sampler_init");
    return MSDF_OK;
}

extern "C" int msdf_sampler_round(int64_t n_rays, int n_old, int n_new, float* z, float* sdf, const float* z_new,
                                  const float* sdf_new, int cap, const float* beta0, float eps, int beta_iters, float* beta,
                                  unsigned int* flag, void* stream) {
    if (n_rays == 0) return MSDF_OK;
    MSDF_CHECK_ARG(z && sdf && sdf_new && beta && flag && beta0, "msdf_sampler_round: null pointer");
    MSDF_CHECK_ARG(n_old >= 0 && n_new >= 1 && n_old + n_new <= cap && n_old + n_new >= 2,
                   "msdf_sampler_round: bad sizes n_old=%d n_new=%d cap=%d", n_old, n_new, cap);
    MSDF_CHECK_ARG(n_old == 0 || z_new, "msdf_sampler_round: z_new required when n_old > 0");
    size_t smem = (size_t)kWarpsPerBlock * 4 * cap * sizeof(float);
    int rc = set_smem(k_sampler_round, smem, "msdf_sampler_round"); if (rc) return rc;
    // algorithmic bytes per ray: the old row (z, sdf) and the new samples in, the merged row out; work = software exps
    const int prof = msdf_prof_begin(MSDF_PROF_SAMPLER, (double)n_rays * 3.0 * (beta_iters + 1) * (n_old + n_new), (cudaStream_t)stream,
                                     (double)n_rays * 8.0 * (n_old + n_new + (n_old + n_new)));
    k_sampler_round<<<(unsigned)msdf_div_up(n_rays, kWarpsPerBlock), kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
        n_rays, n_old, n_new, z, sdf, z_new, sdf_new, cap, beta0, eps, beta_iters, beta, flag);
    msdf_prof_end(prof, (cudaStream_t)stream);
    MSDF_COUNT_LAUNCH();
    MSDF_CHECK_LAUNCH("msdf_sampler_round");
    return MSDF_OK;
}

extern "C" int msdf_sampler_upsample(int64_t n_rays, int n, const float* z, const float* sdf, int cap,
                                     const float* beta, float add_tiny, const float* u, int n_new, const float* ray_o,
                                     const float* ray_d, float* z_new, float* pts_new, void* stream) {
    if (n_rays == 0) return MSDF_OK;
    MSDF_CHECK_ARG(z && sdf && beta && u && ray_o && ray_d && z_new && pts_new, "msdf_sampler_upsample: null pointer");
    MSDF_CHECK_ARG(n >= 2 && n <= cap && n_new >= 1, "msdf_sampler_upsample: bad sizes n=%d cap=%d n_new=%d", n, cap, n_new);
    size_t smem = (size_t)kWarpsPerBlock * (4 * (cap + 1) + n_new) * sizeof(float);
    int rc = set_smem(k_sampler_upsample, smem, "msdf_sampler_upsample"); if (rc) return rc;
    k_sampler_upsample<<<(unsigned)msdf_div_up(n_rays, kWarpsPerBlock), kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
        n_rays, n, z, sdf, cap, beta, add_tiny, u, n_new, ray_o, ray_d, z_new, pts_new);
    MSDF_COUNT_LAUNCH();
    MSDF_CHECK_LAUNCH("msdf_sampler_upsample");
    return MSDF_OK;
}

extern "C" int msdf_sampler_finalize(int64_t n_rays, int n, const float* z, const float* sdf, int cap,
                                     const float* beta, const float* u, int u_per_ray, int n_s, const int32_t* pick,
                                     int n_extra, float near_, float far_, const int64_t* eik_idx, float* z_out,
                                     float* z_eik, void* stream) {
    if (n_rays == 0) return MSDF_OK;
    MSDF_CHECK_ARG(z && sdf && beta && u && z_out, "msdf_sampler_finalize: null pointer");
    MSDF_CHECK_ARG(n >= 2 && n <= cap && n_s >= 1 && n_extra >= 0, "msdf_sampler_finalize: bad sizes");
    MSDF_CHECK_ARG(n_extra == 0 || pick, "msdf_sampler_finalize: pick required when n_extra > 0");
    MSDF_CHECK_ARG(z_eik == nullptr || eik_idx, "msdf_sampler_finalize: eik_idx required with z_eik");
    size_t smem = (size_t)kWarpsPerBlock * (4 * (cap + 1) + 2 * (n_s + 2 + n_extra)) * sizeof(float);
    int rc = set_smem(k_sampler_finalize, smem, "msdf_sampler_finalize"); if (rc) return rc;
    k_sampler_finalize<<<(unsigned)msdf_div_up(n_rays, kWarpsPerBlock), kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
        n_rays, n, z, sdf, cap, beta, u, u_per_ray, n_s, pick, n_extra, near_, far_, eik_idx, z_out, z_eik);
    MSDF_COUNT_LAUNCH();
    MSDF_CHECK_LAUNCH("msdf_sampler_finalize");
    return MSDF_OK;
}
