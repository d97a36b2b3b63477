/*
 * msdf_detmath.h -- bit-reproducible fp32 math shared by the CUDA sampler kernels
 * (monosdf_b200/csrc/sampler.cu, compiled with -fmad=false) and the C oracle
 * (oracle/sampler_oracle.c, compiled with -ffp-contract=off).
 *
 * Every function here is built only from IEEE-754 correctly rounded primitives
 * (+, -, *, /, fmaf, sqrtf, rintf) and integer bit manipulation, so the CPU oracle and
 * the GPU kernel produce identical bits.  libm / libdevice expf/expm1f are NOT
 * used: they differ from each other in the last ulp, which would make the
 * "sample positions bit-exact" check (BASELINE.json north_star) meaningless.
 *
 * The formulas restated are those of the reference's error-bounded sampler:
 *   density.py:21-26 (Laplace density), ray_sampler.py:141-153 (d* bound),
 *   ray_sampler.py:264-272 (opacity error bound).
 */
#ifndef MSDF_DETMATH_H
#define MSDF_DETMATH_H

#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define MSDF_HD __host__ __device__ __forceinline__
#else
#define MSDF_HD static inline
#endif

MSDF_HD float msdf_bits_to_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

/* 2^k for k in [-126, 127] */
MSDF_HD float msdf_pow2i(int k) { return msdf_bits_to_float((uint32_t)(k + 127) << 23); }

/* e^r - 1 - r = r^2 * Q(r) on |r| <= ln2/2 ; Cephes expf coefficients */
MSDF_HD float msdf_exp_q(float r) {
    float q = 1.9875691500e-4f;
    q = fmaf(q, r, 1.3981999507e-3f);
    q = fmaf(q, r, 8.3334519073e-3f);
    q = fmaf(q, r, 4.1665795894e-2f);
    q = fmaf(q, r, 1.6666665459e-1f);
    q = fmaf(q, r, 5.0000001201e-1f);
    return q;
}

MSDF_HD float msdf_expf(float x) {
    if (!(x <= 88.72283f)) return (x != x) ? x : INFINITY;
    if (x < -104.0f) return 0.0f;
    float n = rintf(x * 1.44269504088896341f);
    float r = fmaf(n, -0.693145751953125f, x);
    r = fmaf(n, -1.42860682030941723212e-6f, r);
    float p = fmaf(r * r, msdf_exp_q(r), r) + 1.0f;
    int ni = (int)n;
    int n1 = ni >> 1;
    int n2 = ni - n1;
    return (p * msdf_pow2i(n1)) * msdf_pow2i(n2);
}

MSDF_HD float msdf_expm1f(float x) {
    if (x != x) return x;
    if (x < -17.5f) return -1.0f;
    if (x > 88.72283f) return INFINITY;
    float n = rintf(x * 1.44269504088896341f);
    float r = fmaf(n, -0.693145751953125f, x);
    r = fmaf(n, -1.42860682030941723212e-6f, r);
    float em1 = fmaf(r * r, msdf_exp_q(r), r);
    int ni = (int)n;
    if (ni == 0) return em1;
    if (ni > 127) ni = 127;
    float t = msdf_pow2i(ni);
    return fmaf(em1, t, t - 1.0f);
}

MSDF_HD float msdf_signf(float x) { return (float)((x > 0.0f) - (x < 0.0f)); }

/* LaplaceDensity.density_func, density.py:21-26: alpha*(0.5 + 0.5*sign(s)*expm1(-|s|/beta)) */
MSDF_HD float msdf_density(float sdf, float beta) {
    float alpha = 1.0f / beta;
    float e = msdf_expm1f(-fabsf(sdf) / beta);
    return alpha * (0.5f + (0.5f * msdf_signf(sdf)) * e);
}

/* ray_sampler.py:141-153: d* for one interval; a = z[i+1]-z[i], d0 = sdf[i], d1 = sdf[i+1] */
MSDF_HD float msdf_dstar(float a, float d0, float d1) {
    float b = fabsf(d0), c = fabsf(d1);
    float a2 = a * a, b2 = b * b, c2 = c * c;
    int first = (a2 + b2) <= c2;
    int second = (a2 + c2) <= b2;
    float ds = 0.0f;
    if (first) ds = b;
    if (second) ds = c;
    if (!first && !second && ((b + c) - a > 0.0f)) {
        float s = ((a + b) + c) / 2.0f;
        float area = ((s * (s - a)) * (s - b)) * (s - c);
        ds = (2.0f * sqrtf(area)) / a;
    }
    float sg = msdf_signf(d1) * msdf_signf(d0);
    return (sg == 1.0f) ? ds : 0.0f * ds;   /* bool * float, ray_sampler.py:153 */
}

/* ray_sampler.py:268: exp(-dstar / beta) * dists^2 / (4 beta^2) */
MSDF_HD float msdf_err_section(float d_star, float dist, float beta) {
    return (msdf_expf(-d_star / beta) * (dist * dist)) / (4.0f * (beta * beta));
}

/* ray_sampler.py:270: (clamp(exp(E), max=1e6) - 1) * T */
MSDF_HD float msdf_bound_opacity(float err_integral, float transmittance) {
    float e = msdf_expf(err_integral);
    e = (e > 1.0e6f) ? 1.0e6f : e;
    return (e - 1.0f) * transmittance;
}

#endif /* MSDF_DETMATH_H */
