// Pixel -> world ray generation for the image ("uv") input path.
// Replaces rend_util.get_camera_params + lift (reference code/utils/rend_util.py:63-91,105-118) for 4x4 poses.
#include "common.cuh"

namespace {
__global__ void k_camera_rays(const float* __restrict__ uv, const float* __restrict__ pose, const float* __restrict__ intr,
                              int64_t B, int64_t N, float* __restrict__ dirs, float* __restrict__ cam_loc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * N) return;
    const int64_t b = i / N;
    const float* P = pose + b * 16;
    const float* K = intr + b * 16;
    const float fx = K[0], fy = K[5], cx = K[2], cy = K[6], sk = K[1];
    const float x = uv[2 * i], y = uv[2 * i + 1], z = 1.0f;
    const float xl = (x - cx + cy * sk / fy - sk * y / fy) / fx * z;   // lift(), rend_util.py:105-118
    const float yl = (y - cy) / fy * z;
    float w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) w[k] = P[k * 4 + 0] * xl + P[k * 4 + 1] * yl + P[k * 4 + 2] * z + P[k * 4 + 3] - P[k * 4 + 3];
    float n = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    n = n > 1e-12f ? n : 1e-12f;                                        // F.normalize eps
    dirs[3 * i] = w[0] / n; dirs[3 * i + 1] = w[1] / n; dirs[3 * i + 2] = w[2] / n;
    if (i - b * N == 0) { cam_loc[3 * b] = P[3]; cam_loc[3 * b + 1] = P[7]; cam_loc[3 * b + 2] = P[11]; }
}
}  // namespace

extern "C" int msdf_camera_rays(const float* uv, const float* pose, const float* intrinsics, int64_t batch, int64_t n_pixels,
                                float* ray_dirs, float* cam_loc, void* stream) {
    if (batch * n_pixels == 0) return MSDF_OK;
    MSDF_CHECK_ARG(uv && pose && intrinsics && ray_dirs && cam_loc, "msdf_camera_rays: null pointer");
    k_camera_rays<<<(unsigned)msdf_div_up(batch * n_pixels, 256), 256, 0, (cudaStream_t)stream>>>(uv, pose, intrinsics, batch, n_pixels,
                                                                                                ray_dirs, cam_loc);
    MSDF_COUNT_LAUNCH();
    MSDF_CHECK_LAUNCH("msdf_camera_rays");
    return MSDF_OK;
}
