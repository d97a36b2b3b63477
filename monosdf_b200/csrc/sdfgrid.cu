// Coarse-to-fine SDF volume for marching cubes ("next" row f4 of SURVEY.md section 8).
// Replaces the point pyramid / masked evaluation of utils/plots.py:131-194 (get_surface_sliding): the reference builds
// the crop's cropN^3 points on the host, average-pools them three times, evaluates the coarsest level densely and every
// finer level only where the (nearest-upsampled) parent had |sdf| < threshold, copying every 100 000-point chunk of SDF
// values back to the host.  Here a level's points are generated and compacted on the device, the SDF network runs on
// the compacted list, and a second kernel assembles the level (evaluated value, else the parent's) and its mask.
#include "common.cuh"

namespace {

// fine-level coordinate n of a crop axis: float32(np.linspace(lo, hi, cropN)[n]) (plots.py:135-137, 143)
__device__ __forceinline__ float fine_coord(double lo, double hi, int cropN, int n) {
    if (cropN <= 1) return (float)lo;
    const double step = (hi - lo) / (double)(cropN - 1);
    return (float)(n == cropN - 1 ? hi : lo + (double)n * step);
}
// coordinate i of pyramid level s (cells of 2^s fine samples): the mean of the float32 fine coordinates, which is what
// s successive AvgPool3d(2) passes compute up to fp32 rounding (plots.py:152-156)
__device__ __forceinline__ float level_coord(double lo, double hi, int cropN, int s, int i) {
    const int w = 1 << s;
    double acc = 0.0;
    for (int t = 0; t < w; ++t) acc += (double)fine_coord(lo, hi, cropN, i * w + t);
    return (float)(acc / (double)w);
}

// One thread per cell of the level (n^3 cells, n = cropN >> s).  A cell is evaluated when there is no parent mask
// (coarsest level) or its parent cell's mask is set; evaluated cells get a slot in the compacted point list.
__global__ void k_level_points(double lo0, double lo1, double lo2, double hi0, double hi1, double hi2, int cropN, int s,
                               const unsigned char* __restrict__ parent_mask, int* __restrict__ slot, float* __restrict__ points,
                               int* __restrict__ counter) {
    const int n = cropN >> s;
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (int64_t)n * n * n) return;
    const int k = (int)(c % n), j = (int)((c / n) % n), i = (int)(c / ((int64_t)n * n));      // x slowest (meshgrid 'ij')
    bool on = true;
    if (parent_mask != nullptr) {
        const int h = n >> 1;
        on = parent_mask[((int64_t)(i >> 1) * h + (j >> 1)) * h + (k >> 1)] != 0;            // nearest upsample (plots.py:183-184)
    }
    int my = -1;
    if (on) {
        my = atomicAdd(counter, 1);
        points[3 * (int64_t)my] = level_coord(lo0, hi0, cropN, s, i);
        points[3 * (int64_t)my + 1] = level_coord(lo1, hi1, cropN, s, j);
        points[3 * (int64_t)my + 2] = level_coord(lo2, hi2, cropN, s, k);
    }
    slot[c] = my;
}

// level[c] = evaluated value, else the parent's value (plots.py:175-176, 186-188); mask[c] = |level[c]| < threshold (:181)
__global__ void k_level_assemble(int n, const int* __restrict__ slot, const float* __restrict__ values,
                                 const float* __restrict__ parent, float threshold, float* __restrict__ level,
                                 unsigned char* __restrict__ mask) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (int64_t)n * n * n) return;
    const int sl = slot[c];
    float v;
    if (sl >= 0) v = values[sl];
    else {
        const int k = (int)(c % n), j = (int)((c / n) % n), i = (int)(c / ((int64_t)n * n));
        const int h = n >> 1;
        v = parent[((int64_t)(i >> 1) * h + (j >> 1)) * h + (k >> 1)];
    }
    level[c] = v;
    if (mask != nullptr) mask[c] = fabsf(v) < threshold ? 1 : 0;
}
}  // namespace

extern "C" int msdf_sdfgrid_level_points(const double* lo, const double* hi, int crop_n, int level_shift, const unsigned char* parent_mask,
                                         int* slot, float* points, int* counter, void* stream) {
    MSDF_CHECK_ARG(lo && hi && slot && points && counter, "msdf_sdfgrid_level_points: null pointer");
    MSDF_CHECK_ARG(crop_n > 0 && level_shift >= 0 && (crop_n >> level_shift) > 0 && ((crop_n >> level_shift) << level_shift) == crop_n,
                   "msdf_sdfgrid_level_points: crop_n=%d is not divisible by 2^%d", crop_n, level_shift);
    MSDF_CHECK_ARG(parent_mask == nullptr || ((crop_n >> level_shift) % 2) == 0, "msdf_sdfgrid_level_points: odd level size under a parent mask");
    const int64_t n = crop_n >> level_shift;
    k_level_points<<<(unsigned)msdf_div_up(n * n * n, 256), 256, 0, (cudaStream_t)stream>>>(lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], crop_n, level_shift,
                                                                                           parent_mask, slot, points, counter);
    MSDF_COUNT_LAUNCH();
    MSDF_CHECK_LAUNCH("msdf_sdfgrid_level_points");
    return MSDF_OK;
}

extern "C" int msdf_sdfgrid_level_assemble(int n, const int* slot, const float* values, const float* parent, float threshold,
                                           float* level, unsigned char* mask, void* stream) {
    MSDF_CHECK_ARG(slot && level && n > 0, "msdf_sdfgrid_level_assemble: null pointer");
    const int64_t cells = (int64_t)n * n * n;
    k_level_assemble<<<(unsigned)msdf_div_up(cells, 256), 256, 0, (cudaStream_t)stream>>>(n, slot, values, parent, threshold, level, mask);
    MSDF_COUNT_LAUNCH();
    MSDF_CHECK_LAUNCH("msdf_sdfgrid_level_assemble");
    return MSDF_OK;
}
