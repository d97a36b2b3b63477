// Multi-resolution hash-grid encoder: forward gather (+ analytic dy/dx), backward scatter and the
// double-backward terms used by the eikonal / normal losses.
//
// Semantics follow the reference op exactly (code/hashencoder/src/hashencoder.cu):
//   index/hash        :35-72   (dense while stride <= hashmap_size, else x*1 ^ y*2654435761 ^ z*805459861; % size)
//   smoothstep        :87-93   (weights are smoothstep(frac), derivative 6f(1-f))
//   forward + dy_dx   :104-254 (scale = exp2f(l*S)*H - 1, res = ceil(scale)+1, OOB inputs -> zeros)
//   backward          :258-343 (atomic scatter of w*grad), input grad :347-372
//   second backward   :376-428 (wrt incoming grad), :432-595 (wrt table)
// Entry points mirror hashencoder.h:13-15 with raw device pointers instead of at::Tensor.
//
// B200 layout decisions (differences from the reference's launch shape, not its results):
//   * one thread per (POINT, LEVEL), level fastest (the whole 46.5 MB table is L2-resident on B200, so the
//     reference's level-major grid "so one level fits cache" buys nothing): a warp holds two points x 16 levels, all
//     128 gathers of a point are in flight at once (8-byte float2 loads), and the 128 B feature row / 384 B dy_dx row
//     of a point leave the warp as contiguous runs;
//   * scatters use 8-byte vector atomics (red.global.add.v2.f32) into the L2-resident gradient table.
// All three kernels are HBM/L2-gather bound; algorithmic bytes per point are listed in DESIGN.md.
#include "common.cuh"

namespace {

constexpr uint32_t kPrimeY = 2654435761u, kPrimeZ = 805459861u;

struct LevelGeom { uint32_t hashmap_size; uint32_t res; float scale; };

__device__ __forceinline__ LevelGeom level_geom(const int* __restrict__ offsets, uint32_t level, float S, uint32_t H) {
    LevelGeom g;
    g.hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
    g.scale = exp2f(level * S) * H - 1.0f;
    g.res = (uint32_t)ceilf(g.scale) + 1;
    return g;
}

__device__ __forceinline__ uint32_t grid_index(uint32_t x, uint32_t y, uint32_t z, uint32_t hashmap_size, uint32_t res) {
    // per-lookup dense-or-hash decision by stride overflow (hashencoder.cu:54-72)
    uint32_t stride = 1, index = 0;
    if (stride <= hashmap_size) { index += x * stride; stride *= res; }
    if (stride <= hashmap_size) { index += y * stride; stride *= res; }
    if (stride <= hashmap_size) { index += z * stride; stride *= res; }
    if (stride > hashmap_size) index = x ^ (y * kPrimeY) ^ (z * kPrimeZ);
    // index % hashmap_size without the ~20-instruction integer division where it can be avoided: hashed levels have a
    // power-of-two size (a mask), dense levels only wrap for the corner coordinate that equals res (a rare branch)
    if ((hashmap_size & (hashmap_size - 1u)) == 0u) return index & (hashmap_size - 1u);
    return index < hashmap_size ? index : index % hashmap_size;
}

__device__ __forceinline__ bool load_point(const float* __restrict__ x, int64_t b, float divide_factor, float p[3]) {
    bool oob = false;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float v = x[b * 3 + d];
        if (divide_factor > 0.0f) v = (v / divide_factor + 1.0f) / 2.0f;   // network.py:250, hashgrid.py:158
        p[d] = v;
        oob |= (v < 0.0f) || (v > 1.0f);
    }
    return oob;
}

struct Cell { uint32_t g[3]; float w1[3]; float dw[3]; };

__device__ __forceinline__ Cell make_cell(const float p[3], float scale) {
    Cell c;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float pos = p[d] * scale;
        float fl = floorf(pos);
        c.g[d] = (uint32_t)fl;
        float f = pos - fl;
        c.dw[d] = 6 * f * (1.0f - f);
        c.w1[d] = f * f * (3.0f - 2.0f * f);
    }
    return c;
}

template <int C> struct VecT;
template <> struct VecT<1> { using T = float; };
template <> struct VecT<2> { using T = float2; };
template <> struct VecT<4> { using T = float4; };

template <int C>
__device__ __forceinline__ void load_feat(const float* __restrict__ t, uint32_t idx, float v[C]) {
    if constexpr (C == 2) { float2 q = __ldg(reinterpret_cast<const float2*>(t) + idx); v[0] = q.x; v[1] = q.y; }
    else if constexpr (C == 4) { float4 q = __ldg(reinterpret_cast<const float4*>(t) + idx); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
    else {
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = __ldg(t + (size_t)idx * C + c);
    }
}

template <int C>
__device__ __forceinline__ void atomic_add_feat(float* t, uint32_t idx, const float v[C]) {
    if constexpr (C == 2) {
        atomicAdd(reinterpret_cast<float2*>(t) + idx, make_float2(v[0], v[1]));
    } else if constexpr (C == 4) {
        atomicAdd(reinterpret_cast<float4*>(t) + idx, make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) atomicAdd(t + (size_t)idx * C + c, v[c]);
    }
}

// outputs: level_major ? [L,B,C] : row b at out + b*out_ld, feature (l*C + c)
// dy_dx  : [B, L, 3, C] (hashencoder.cu:212)
// One thread per (point, level), level fastest: a warp covers 32 / L points x all L levels, so that ALL the gathers of
// a point (L x 8 corners) are in flight at once instead of 8 at a time (the thread-per-point form walked the levels
// serially: 296 us for 262144 points, 14 % of the issue slots busy), and a point's feature row / dy_dx row leaves the
// warp as one contiguous run (L*C floats: lane l writes features l*C .. l*C+C-1).
template <int C>
__global__ void __launch_bounds__(256)
k_hash_forward(const float* __restrict__ x, const float* __restrict__ table, const int* __restrict__ offsets,
               float* __restrict__ out, int64_t out_ld, int level_major, int64_t B, int L, float S, uint32_t H,
               float divide_factor, float* __restrict__ dy_dx) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t b = gid / L;
    const int l = (int)(gid - b * L);
    if (b >= B) return;
    float p[3];
    const bool oob = load_point(x, b, divide_factor, p);
    float res[C];
    float grad[3][C];
#pragma unroll
    for (int c = 0; c < C; ++c) { res[c] = 0.0f; grad[0][c] = grad[1][c] = grad[2][c] = 0.0f; }
    if (!oob) {
        const LevelGeom g = level_geom(offsets, l, S, H);
        const float* t = table + (size_t)offsets[l] * C;
        const Cell cell = make_cell(p, g.scale);
        float v[8][C];
#pragma unroll
        for (int idx = 0; idx < 8; ++idx) {
            uint32_t gi = grid_index(cell.g[0] + (idx & 1), cell.g[1] + ((idx >> 1) & 1), cell.g[2] + ((idx >> 2) & 1),
                                     g.hashmap_size, g.res);
            load_feat<C>(t, gi, v[idx]);
        }
#pragma unroll
        for (int idx = 0; idx < 8; ++idx) {
            float w = 1.0f;
#pragma unroll
            for (int d = 0; d < 3; ++d) w *= ((idx >> d) & 1) ? cell.w1[d] : 1.0f - cell.w1[d];
#pragma unroll
            for (int c = 0; c < C; ++c) res[c] += w * v[idx][c];
        }
        if (dy_dx != nullptr) {
#pragma unroll
            for (int gd = 0; gd < 3; ++gd) {
#pragma unroll
                for (int sub = 0; sub < 4; ++sub) {
                    float w = g.scale;
                    int base = 0;
#pragma unroll
                    for (int nd = 0; nd < 2; ++nd) {
                        const int d = (nd >= gd) ? nd + 1 : nd;
                        const int bit = (sub >> nd) & 1;
                        w *= bit ? cell.w1[d] : 1.0f - cell.w1[d];
                        base |= bit << d;
                    }
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        grad[gd][c] += w * (v[base | (1 << gd)][c] - v[base][c]) * cell.dw[gd];
                }
            }
        }
    }
    float* o = level_major ? out + ((int64_t)l * B + b) * C : out + b * out_ld + l * C;
    if constexpr (C == 2) {
        if ((reinterpret_cast<uintptr_t>(o) & 7) == 0) *reinterpret_cast<float2*>(o) = make_float2(res[0], res[1]);
        else { o[0] = res[0]; o[1] = res[1]; }
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) o[c] = res[c];
    }
    if (dy_dx != nullptr) {
        float* od = dy_dx + ((b * L + l) * 3) * C;
        if constexpr (C == 2) {       // 24 bytes per (point, level), 8-byte aligned: three vector stores
#pragma unroll
            for (int gd = 0; gd < 3; ++gd) *reinterpret_cast<float2*>(od + gd * 2) = make_float2(grad[gd][0], grad[gd][1]);
        } else {
#pragma unroll
            for (int gd = 0; gd < 3; ++gd)
#pragma unroll
                for (int c = 0; c < C; ++c) od[gd * C + c] = grad[gd][c];
        }
    }
}

// grad: level_major ? [L,B,C] : row b at grad + b*grad_ld.
// Scatters  w*grad  (first backward, :258-343)  and, if gg_x != nullptr, the double-backward table term
//   +-scale * w_other * smoothstep'(f_d) * grad2[l,b,c] * gg_x[b,d]   (:432-595)
// where grad2 is the "grad" operand of the second backward (the first-backward incoming gradient).
// Either source may be null; both go into the same table with one vector atomic per corner.
template <int C>
__global__ void __launch_bounds__(256)
k_hash_scatter(const float* __restrict__ x, const int* __restrict__ offsets, int64_t B, int L, float S, uint32_t H,
               float divide_factor,
               const float* __restrict__ grad, int64_t grad_ld, int grad_level_major,
               const float* __restrict__ grad2, int64_t grad2_ld, int grad2_level_major,
               const float* __restrict__ gg_x, float gg_scale,
               float* __restrict__ grad_table) {
    // one thread per (point, level), level fastest (see k_hash_forward): the gradient row of a point is read as one
    // contiguous run by the warp and all L x 8 vector reductions of a point are issued at once
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t b = gid / L;
    const int l = (int)(gid - b * L);
    if (b >= B) return;
    float p[3];
    if (load_point(x, b, divide_factor, p)) return;   // grad table is pre-zeroed by the caller (:285)
    float ggx[3] = {0.f, 0.f, 0.f};
    if (gg_x != nullptr) {
#pragma unroll
        for (int d = 0; d < 3; ++d) ggx[d] = gg_x[b * 3 + d] * gg_scale;
    }
    const LevelGeom g = level_geom(offsets, l, S, H);
    const Cell cell = make_cell(p, g.scale);
    float g1[C], g2[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        g1[c] = grad ? (grad_level_major ? grad[((int64_t)l * B + b) * C + c] : grad[b * grad_ld + l * C + c]) : 0.0f;
        g2[c] = (gg_x && grad2) ? (grad2_level_major ? grad2[((int64_t)l * B + b) * C + c] : grad2[b * grad2_ld + l * C + c]) : 0.0f;
    }
    float* t = grad_table + (size_t)offsets[l] * C;
#pragma unroll
    for (int idx = 0; idx < 8; ++idx) {
        float w = 1.0f;
#pragma unroll
        for (int d = 0; d < 3; ++d) w *= ((idx >> d) & 1) ? cell.w1[d] : 1.0f - cell.w1[d];
        // d(w)/d(x01_d) * gg_x[d] summed over d
        float wd = 0.0f;
        if (gg_x != nullptr) {
#pragma unroll
            for (int gd = 0; gd < 3; ++gd) {
                float wo = g.scale;
#pragma unroll
                for (int d = 0; d < 3; ++d)
                    if (d != gd) wo *= ((idx >> d) & 1) ? cell.w1[d] : 1.0f - cell.w1[d];
                const float sgn = ((idx >> gd) & 1) ? 1.0f : -1.0f;
                wd += sgn * wo * cell.dw[gd] * ggx[gd];
            }
        }
        float val[C];
#pragma unroll
        for (int c = 0; c < C; ++c) val[c] = w * g1[c] + wd * g2[c];
        const uint32_t gi = grid_index(cell.g[0] + (idx & 1), cell.g[1] + ((idx >> 1) & 1), cell.g[2] + ((idx >> 2) & 1),
                                       g.hashmap_size, g.res);
        atomic_add_feat<C>(t, gi, val);
    }
}

// grad_inputs[b,d] = sum_{l,c} grad[l,b,c] * dy_dx[b,l,d,c]    (:347-372)
__global__ void k_hash_input_backward(const float* __restrict__ grad, int64_t grad_ld, int level_major,
                                      const float* __restrict__ dy_dx, float* __restrict__ grad_inputs,
                                      int64_t B, int L, int C) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * 3) return;
    const int64_t b = t / 3; const int d = (int)(t - b * 3);
    const float* dd = dy_dx + b * L * 3 * C;
    float r = 0.0f;
    for (int l = 0; l < L; ++l)
        for (int c = 0; c < C; ++c) {
            float g = level_major ? grad[((int64_t)l * B + b) * C + c] : grad[b * grad_ld + l * C + c];
            r += g * dd[(l * 3 + d) * C + c];
        }
    grad_inputs[t] = r;
}

// grad_grad[l,b,c] = sum_d gg_x[b,d] * dy_dx[b,l,d,c]   (:376-428); out level_major or row-major with ld
__global__ void k_hash_second_backward_grad(const float* __restrict__ gg_x, float gg_scale, const float* __restrict__ dy_dx,
                                            float* __restrict__ out, int64_t out_ld, int level_major,
                                            int64_t B, int L, int C) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * L) return;
    const int64_t b = t / L; const int l = (int)(t - b * L);
    const float* dd = dy_dx + (b * L + l) * 3 * C;
    for (int c = 0; c < C; ++c) {
        float r = 0.0f;
        for (int d = 0; d < 3; ++d) r += (gg_x[b * 3 + d] * gg_scale) * dd[d * C + c];
        if (level_major) out[((int64_t)l * B + b) * C + c] = r; else out[b * out_ld + l * C + c] = r;
    }
}

template <int C>
int launch_forward(const float* x, const float* table, const int* offsets, float* out, int64_t out_ld, int level_major,
                   int64_t B, int L, float S, uint32_t H, float divide_factor, float* dy_dx, cudaStream_t st) {
    // algorithmic bytes per point: 12 (x) + L*8 corners*C*4 (gathers) + L*C*4 (features) (+ L*3*C*4 when dy_dx is materialised)
    const int prof = msdf_prof_begin(MSDF_PROF_HASH, 0.0, st, (double)B * (12.0 + L * C * 4.0 * 9.0 + (dy_dx ? L * C * 12.0 : 0.0)));
    k_hash_forward<C><<<(unsigned)msdf_div_up(B * L, 256), 256, 0, st>>>(x, table, offsets, out, out_ld, level_major, B, L, S, H,
                                                                    divide_factor, dy_dx);
    msdf_prof_end(prof, st);
    MSDF_COUNT_LAUNCH();
    MSDF_CHECK_LAUNCH("msdf_hashgrid_forward");
    return MSDF_OK;
}

template <int C>
int launch_scatter(const float* x, const int* offsets, int64_t B, int L, float S, uint32_t H, float divide_factor,
                   const float* grad, int64_t grad_ld, int glm, const float* grad2, int64_t grad2_ld, int g2lm,
                   const float* gg_x, float gg_scale, float* grad_table, cudaStream_t st) {
    // per point: 12 (x) + L*C*4 (grad) + L*8*C*4*2 (read-modify-write of the gradient table)
    const int prof = msdf_prof_begin(MSDF_PROF_HASH, 0.0, st, (double)B * (12.0 + L * C * 4.0 * (grad2 ? 2.0 : 1.0) + L * C * 64.0));
    k_hash_scatter<C><<<(unsigned)msdf_div_up(B * L, 256), 256, 0, st>>>(x, offsets, B, L, S, H, divide_factor, grad, grad_ld, glm,
                                                                    grad2, grad2_ld, g2lm, gg_x, gg_scale, grad_table);
    msdf_prof_end(prof, st);
    MSDF_COUNT_LAUNCH();
    MSDF_CHECK_LAUNCH("msdf_hashgrid_scatter");
    return MSDF_OK;
}

}  // namespace

// ---- internal (engine) entry points -------------------------------------------------------------------------
int msdf_hash_forward_rows(const float* x, const float* table, const int* offsets, float* out, int64_t out_ld,
                           int64_t B, int C, int L, float S, uint32_t H, float divide_factor, float* dy_dx, cudaStream_t st) {
    if (B == 0) return MSDF_OK;
    switch (C) {
        case 1: return launch_forward<1>(x, table, offsets, out, out_ld, 0, B, L, S, H, divide_factor, dy_dx, st);
        case 2: return launch_forward<2>(x, table, offsets, out, out_ld, 0, B, L, S, H, divide_factor, dy_dx, st);
        case 4: return launch_forward<4>(x, table, offsets, out, out_ld, 0, B, L, S, H, divide_factor, dy_dx, st);
        default: msdf_set_error("hash grid: level_dim must be 1, 2 or 4 (got %d)", C); return MSDF_ERR_UNSUPPORTED;
    }
}

int msdf_hash_scatter_rows(const float* x, const int* offsets, int64_t B, int C, int L, float S, uint32_t H, float divide_factor,
                           const float* grad, int64_t grad_ld, const float* grad2, int64_t grad2_ld, const float* gg_x,
                           float gg_scale, float* grad_table, cudaStream_t st) {
    if (B == 0) return MSDF_OK;
    switch (C) {
        case 1: return launch_scatter<1>(x, offsets, B, L, S, H, divide_factor, grad, grad_ld, 0, grad2, grad2_ld, 0, gg_x, gg_scale, grad_table, st);
        case 2: return launch_scatter<2>(x, offsets, B, L, S, H, divide_factor, grad, grad_ld, 0, grad2, grad2_ld, 0, gg_x, gg_scale, grad_table, st);
        case 4: return launch_scatter<4>(x, offsets, B, L, S, H, divide_factor, grad, grad_ld, 0, grad2, grad2_ld, 0, gg_x, gg_scale, grad_table, st);
        default: msdf_set_error("hash grid: level_dim must be 1, 2 or 4 (got %d)", C); return MSDF_ERR_UNSUPPORTED;
    }
}

// ---- C ABI, mirroring hashencoder.h:13-15 ----------------------------------------------------------------------
extern "C" int msdf_hash_encode_forward(const float* inputs, const float* embeddings, const int32_t* offsets, float* outputs,
                                        uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                                        int calc_grad_inputs, float* dy_dx, void* stream) {
    if (B == 0) return MSDF_OK;
    MSDF_CHECK_ARG(inputs && embeddings && offsets && outputs, "msdf_hash_encode_forward: null pointer");
    MSDF_CHECK_ARG(D == 3, "msdf_hash_encode_forward: only D=3 is built (got D=%u)", D);
    MSDF_CHECK_ARG(!calc_grad_inputs || dy_dx, "msdf_hash_encode_forward: dy_dx required when calc_grad_inputs");
    cudaStream_t st = (cudaStream_t)stream;
    float* dd = calc_grad_inputs ? dy_dx : nullptr;
    switch (C) {
        case 1: return launch_forward<1>(inputs, embeddings, offsets, outputs, 0, 1, B, L, S, H, 0.0f, dd, st);
        case 2: return launch_forward<2>(inputs, embeddings, offsets, outputs, 0, 1, B, L, S, H, 0.0f, dd, st);
        case 4: return launch_forward<4>(inputs, embeddings, offsets, outputs, 0, 1, B, L, S, H, 0.0f, dd, st);
        default: msdf_set_error("msdf_hash_encode_forward: C must be 1, 2 or 4 (got %u)", C); return MSDF_ERR_UNSUPPORTED;
    }
}

extern "C" int msdf_hash_encode_backward(const float* grad, const float* inputs, const float* embeddings,
                                         const int32_t* offsets, float* grad_embeddings, uint32_t B, uint32_t D, uint32_t C,
                                         uint32_t L, float S, uint32_t H, int calc_grad_inputs, const float* dy_dx,
                                         float* grad_inputs, void* stream) {
    (void)embeddings;
    MSDF_CHECK_ARG(grad && inputs && offsets && grad_embeddings, "msdf_hash_encode_backward: null pointer");
    MSDF_CHECK_ARG(D == 3, "msdf_hash_encode_backward: only D=3 is built (got D=%u)", D);
    MSDF_CHECK_ARG(!calc_grad_inputs || (dy_dx && grad_inputs), "msdf_hash_encode_backward: dy_dx/grad_inputs required");
    if (B == 0) return MSDF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    switch (C) {
        case 1: rc = launch_scatter<1>(inputs, offsets, B, L, S, H, 0.0f, grad, 0, 1, nullptr, 0, 1, nullptr, 0.f, grad_embeddings, st); break;
        case 2: rc = launch_scatter<2>(inputs, offsets, B, L, S, H, 0.0f, grad, 0, 1, nullptr, 0, 1, nullptr, 0.f, grad_embeddings, st); break;
        case 4: rc = launch_scatter<4>(inputs, offsets, B, L, S, H, 0.0f, grad, 0, 1, nullptr, 0, 1, nullptr, 0.f, grad_embeddings, st); break;
        default: msdf_set_error("msdf_hash_encode_backward: C must be 1, 2 or 4 (got %u)", C); return MSDF_ERR_UNSUPPORTED;
    }
    if (rc) return rc;
    if (calc_grad_inputs) {
        k_hash_input_backward<<<(unsigned)msdf_div_up((int64_t)B * 3, 256), 256, 0, st>>>(grad, 0, 1, dy_dx, grad_inputs, B, L, C);
        MSDF_COUNT_LAUNCH();
        MSDF_CHECK_LAUNCH("msdf_hash_encode_backward(input grad)");
    }
    return MSDF_OK;
}

extern "C" int msdf_hash_encode_second_backward(const float* grad, const float* inputs, const float* embeddings,
                                                const int32_t* offsets, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                                                uint32_t H, int calc_grad_inputs, const float* dy_dx,
                                                const float* grad_grad_inputs, float* grad_grad, float* grad2_embeddings,
                                                void* stream) {
    (void)embeddings; (void)calc_grad_inputs;
    MSDF_CHECK_ARG(grad && inputs && offsets && dy_dx && grad_grad_inputs && grad_grad && grad2_embeddings,
                   "msdf_hash_encode_second_backward: null pointer");
    MSDF_CHECK_ARG(D == 3, "msdf_hash_encode_second_backward: only D=3 is built (got D=%u)", D);
    if (B == 0) return MSDF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    k_hash_second_backward_grad<<<(unsigned)msdf_div_up((int64_t)B * L, 256), 256, 0, st>>>(grad_grad_inputs, 1.0f, dy_dx, grad_grad,
                                                                                         0, 1, B, L, C);
    MSDF_COUNT_LAUNCH();
    MSDF_CHECK_LAUNCH("msdf_hash_encode_second_backward(grad)");
    switch (C) {
        case 1: return launch_scatter<1>(inputs, offsets, B, L, S, H, 0.0f, nullptr, 0, 1, grad, 0, 1, grad_grad_inputs, 1.0f, grad2_embeddings, st);
        case 2: return launch_scatter<2>(inputs, offsets, B, L, S, H, 0.0f, nullptr, 0, 1, grad, 0, 1, grad_grad_inputs, 1.0f, grad2_embeddings, st);
        case 4: return launch_scatter<4>(inputs, offsets, B, L, S, H, 0.0f, nullptr, 0, 1, grad, 0, 1, grad_grad_inputs, 1.0f, grad2_embeddings, st);
        default: msdf_set_error("msdf_hash_encode_second_backward: C must be 1, 2 or 4 (got %u)", C); return MSDF_ERR_UNSUPPORTED;
    }
}
