/*
 * sampler_oracle.c -- CPU oracle for the error-bounded ray sampler.  TEST INFRASTRUCTURE ONLY
 * (loaded by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; never by the product).
 *
 * Restates ErrorBoundSampler.get_z_vals (reference code/model/ray_sampler.py:110-262),
 * UniformSampler.get_z_vals / near_far_from_cube (:48-83) and get_error_bound (:264-272) as
 * three per-ray phases with the SDF values SUPPLIED by the caller, in exactly the arithmetic
 * order the CUDA kernels (monosdf_b200/csrc/sampler.cu) use:
 *   - element math from include/msdf_detmath.h (shared source, IEEE-only primitives),
 *   - prefix sums in "row-wise warp order": blocks of 32 consecutive elements, a Kogge-Stone
 *     inclusive scan inside each block, plus a running carry (carry + block_scan[i]).
 * With that, GPU results are bit-identical to this file.  Parity against the reference's own
 * torch arithmetic (different cumsum/exp rounding) is statistical and is pinned in
 * tests/test_sampler_oracle.py against tests/golden/ fixtures made from the real reference.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "../include/msdf_detmath.h"

#define WARP 32

/* inclusive scan, row-wise warp order; in and out may alias */
static void scan_rowwise(const float* in, float* out, int n) {
    float carry = 0.0f;
    for (int base = 0; base < n; base += WARP) {
        float v[WARP], t[WARP];
        for (int l = 0; l < WARP; ++l) v[l] = (base + l < n) ? in[base + l] : 0.0f;
        for (int off = 1; off < WARP; off <<= 1) {
            for (int l = 0; l < WARP; ++l) t[l] = (l >= off) ? v[l] + v[l - off] : v[l];
            memcpy(v, t, sizeof(v));
        }
        for (int l = 0; l < WARP; ++l) {
            float r = carry + v[l];
            if (base + l < n) out[base + l] = r;
            v[l] = r;
        }
        carry = v[WARP - 1];
    }
}

/* total of n elements in the same order (last element of the inclusive scan over a zero-padded row) */
static float sum_rowwise(const float* in, int n) {
    float carry = 0.0f;
    for (int base = 0; base < n; base += WARP) {
        float v[WARP], t[WARP];
        for (int l = 0; l < WARP; ++l) v[l] = (base + l < n) ? in[base + l] : 0.0f;
        for (int off = 1; off < WARP; off <<= 1) {
            for (int l = 0; l < WARP; ++l) t[l] = (l >= off) ? v[l] + v[l - off] : v[l];
            memcpy(v, t, sizeof(v));
        }
        carry = carry + v[WARP - 1];
    }
    return carry;
}

/* ray_sampler.py:48-60 (far only) */
static float cube_far(const float* o, const float* d, float bound, float far_max) {
    float nearv = -INFINITY, farv = INFINITY;
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

/* Phase 0: initial samples, beta upper bound (Lemma 2) and the sample points. ray_sampler.py:63-83,114-120 */
void msdf_oracle_sampler_init(const float* ray_o, const float* ray_d, int64_t n_rays,
                              const float* t_vals, const float* t_rand, int n0,
                              float bound, float nearv, float far_max, float beta_coef,
                              float* z, int cap, float* beta, float* pts) {
    float* tmp = (float*)malloc(sizeof(float) * n0);
    for (int64_t r = 0; r < n_rays; ++r) {
        const float* o = ray_o + 3 * r; const float* d = ray_d + 3 * r;
        float farv = cube_far(o, d, bound, far_max);
        float* zr = z + (int64_t)cap * r;
        for (int j = 0; j < n0; ++j) zr[j] = nearv * (1.0f - t_vals[j]) + farv * t_vals[j];
        if (t_rand) {
            for (int j = 0; j < n0; ++j) {
                float lower = (j == 0) ? zr[0] : 0.5f * (zr[j] + zr[j - 1]);
                float upper = (j == n0 - 1) ? zr[n0 - 1] : 0.5f * (zr[j + 1] + zr[j]);
                tmp[j] = lower + (upper - lower) * t_rand[(int64_t)n0 * r + j];
            }
            memcpy(zr, tmp, sizeof(float) * n0);
        }
        for (int j = 0; j < n0 - 1; ++j) { float dd = zr[j + 1] - zr[j]; tmp[j] = dd * dd; }
        beta[r] = sqrtf(beta_coef * sum_rowwise(tmp, n0 - 1));
        for (int j = 0; j < n0; ++j)
            for (int k = 0; k < 3; ++k) pts[((int64_t)n0 * r + j) * 3 + k] = o[k] + zr[j] * d[k];
    }
    free(tmp);
}

/* get_error_bound for one ray (ray_sampler.py:264-272) */
static float error_bound(const float* sdf, const float* dists, const float* dstar, int n, float beta,
                         float* w0, float* w1) {
    /* w0: shifted free energy -> integral ; w1: error per section -> error integral */
    w0[0] = 0.0f;
    for (int i = 0; i < n - 2; ++i) w0[i + 1] = dists[i] * msdf_density(sdf[i], beta);
    scan_rowwise(w0, w0, n - 1);
    for (int i = 0; i < n - 1; ++i) w1[i] = msdf_err_section(dstar[i], dists[i], beta);
    scan_rowwise(w1, w1, n - 1);
    float m = -INFINITY;
    for (int i = 0; i < n - 1; ++i) {
        float bo = msdf_bound_opacity(w1[i], msdf_expf(-w0[i]));
        m = (bo > m) ? bo : m;   /* NaN never wins, like a fmaxf-reduction */
    }
    return m;
}

/* Phase A: merge the new samples, d*, beta line search. ray_sampler.py:132-165,179 */
void msdf_oracle_sampler_round(int64_t n_rays, int n_old, int n_new, float* z, float* sdf,
                               const float* z_new, const float* sdf_new, int cap,
                               float beta0, float eps, int beta_iters, float* beta, uint32_t* flag) {
    int n = n_old + n_new;
    float* zb = (float*)malloc(sizeof(float) * n * 6);
    float *sb = zb + n, *dists = sb + n, *dstar = dists + n, *w0 = dstar + n, *w1 = w0 + n;
    for (int64_t r = 0; r < n_rays; ++r) {
        float* zr = z + (int64_t)cap * r; float* sr = sdf + (int64_t)cap * r;
        if (n_old > 0) {
            const float* zn = z_new + (int64_t)n_new * r; const float* sn = sdf_new + (int64_t)n_new * r;
            for (int i = 0; i < n_old; ++i) {           /* stable: old entries first on ties */
                int c = 0; for (int k = 0; k < n_new; ++k) c += (zn[k] < zr[i]);
                zb[i + c] = zr[i]; sb[i + c] = sr[i];
            }
            for (int j = 0; j < n_new; ++j) {
                int c = 0;
                for (int i = 0; i < n_old; ++i) c += (zr[i] <= zn[j]);
                for (int k = 0; k < n_new; ++k) c += (zn[k] < zn[j]) || (zn[k] == zn[j] && k < j);
                zb[c] = zn[j]; sb[c] = sn[j];
            }
            memcpy(zr, zb, sizeof(float) * n); memcpy(sr, sb, sizeof(float) * n);
        } else {
            memcpy(sr, sdf_new + (int64_t)n_new * r, sizeof(float) * n_new);
        }
        for (int i = 0; i < n - 1; ++i) {
            dists[i] = zr[i + 1] - zr[i];
            dstar[i] = msdf_dstar(dists[i], sr[i], sr[i + 1]);
        }
        float b = beta[r];
        float err = error_bound(sr, dists, dstar, n, beta0, w0, w1);
        if (err <= eps) b = beta0;
        float bmin = beta0, bmax = b;
        for (int it = 0; it < beta_iters; ++it) {
            float mid = (bmin + bmax) / 2.0f;
            err = error_bound(sr, dists, dstar, n, mid, w0, w1);
            if (err <= eps) bmax = mid;
            if (err > eps) bmin = mid;
        }
        beta[r] = bmax;
        if (bmax > beta0) *flag = 1u;
    }
    free(zb);
}

/* shared by both Phase-B variants: pdf -> cdf -> inverse-CDF samples. ray_sampler.py:168-228 */
static void draw_samples(const float* zr, const float* sr, int n, float beta, int upsample, float add_tiny,
                         const float* u, int n_u, float* out, float* scratch) {
    float *dists = scratch, *w0 = dists + n, *w1 = w0 + n, *pdf = w1 + n, *cdf = pdf + n;
    for (int i = 0; i < n - 1; ++i) dists[i] = zr[i + 1] - zr[i];
    /* free energy (last interval 1e10), shifted, transmittance */
    w0[0] = 0.0f;
    for (int i = 0; i < n - 1; ++i) w0[i + 1] = dists[i] * msdf_density(sr[i], beta);
    scan_rowwise(w0, w0, n);
    if (upsample) {
        for (int i = 0; i < n - 1; ++i) w1[i] = msdf_err_section(msdf_dstar(dists[i], sr[i], sr[i + 1]), dists[i], beta);
        scan_rowwise(w1, w1, n - 1);
        for (int i = 0; i < n - 1; ++i) pdf[i] = msdf_bound_opacity(w1[i], msdf_expf(-w0[i])) + add_tiny;
    } else {
        for (int i = 0; i < n - 1; ++i) {
            float fe = dists[i] * msdf_density(sr[i], beta);
            float alpha = 1.0f - msdf_expf(-fe);
            pdf[i] = alpha * msdf_expf(-w0[i]) + 1e-5f;
        }
    }
    float total = sum_rowwise(pdf, n - 1);
    for (int i = 0; i < n - 1; ++i) pdf[i] = pdf[i] / total;
    cdf[0] = 0.0f;
    scan_rowwise(pdf, cdf + 1, n - 1);
    for (int j = 0; j < n_u; ++j) {
        float uj = u[j];
        int lo = 0, hi = n;                       /* searchsorted(right=True): count of cdf <= u */
        while (lo < hi) { int mid = (lo + hi) >> 1; if (cdf[mid] <= uj) lo = mid + 1; else hi = mid; }
        int below = (lo - 1 > 0) ? lo - 1 : 0;
        int above = (lo < n - 1) ? lo : n - 1;
        float denom = cdf[above] - cdf[below];
        if (denom < 1e-5f) denom = 1.0f;
        float t = (uj - cdf[below]) / denom;
        out[j] = zr[below] + t * (zr[above] - zr[below]);
    }
}

/* Phase B (continue): n_new samples proportional to the error bound, and their points. :184-194,211-228 */
void msdf_oracle_sampler_upsample(int64_t n_rays, int n, const float* z, const float* sdf, int cap,
                                  const float* beta, float add_tiny, const float* u, int n_new,
                                  const float* ray_o, const float* ray_d, float* z_new, float* pts_new) {
    float* scratch = (float*)malloc(sizeof(float) * (n + 1) * 5);
    for (int64_t r = 0; r < n_rays; ++r) {
        float* out = z_new + (int64_t)n_new * r;
        draw_samples(z + (int64_t)cap * r, sdf + (int64_t)cap * r, n, beta[r], 1, add_tiny, u, n_new, out, scratch);
        for (int j = 0; j < n_new; ++j)
            for (int k = 0; k < 3; ++k)
                pts_new[((int64_t)n_new * r + j) * 3 + k] = ray_o[3 * r + k] + out[j] * ray_d[3 * r + k];
    }
    free(scratch);
}

/* Phase B (final): N_samples from the opacity pdf + near/far/extra picks, sorted. :199-206,236-255 */
void msdf_oracle_sampler_finalize(int64_t n_rays, int n, const float* z, const float* sdf, int cap,
                                  const float* beta, const float* u, int u_per_ray, int n_s,
                                  const int32_t* pick, int n_extra, float nearv, float farv,
                                  const int64_t* eik_idx, float* z_out, float* z_eik) {
    int n_out = n_s + 2 + n_extra;
    float* scratch = (float*)malloc(sizeof(float) * (n + 1) * 5);
    float* cat = (float*)malloc(sizeof(float) * n_out);
    for (int64_t r = 0; r < n_rays; ++r) {
        const float* zr = z + (int64_t)cap * r;
        draw_samples(zr, sdf + (int64_t)cap * r, n, beta[r], 0, 0.0f, u_per_ray ? u + (int64_t)n_s * r : u, n_s, cat, scratch);
        cat[n_s] = nearv; cat[n_s + 1] = farv;
        for (int k = 0; k < n_extra; ++k) cat[n_s + 2 + k] = zr[pick[k]];
        float* zo = z_out + (int64_t)n_out * r;
        for (int i = 0; i < n_out; ++i) {          /* stable rank sort */
            int c = 0;
            for (int k = 0; k < n_out; ++k) c += (cat[k] < cat[i]) || (cat[k] == cat[i] && k < i);
            zo[c] = cat[i];
        }
        if (z_eik) z_eik[r] = zo[eik_idx[r]];
    }
    free(scratch); free(cat);
}
