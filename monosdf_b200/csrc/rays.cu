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

// lift() + pose, shared by the two kernels: unit world direction of pixel (x, y) for intrinsics K and pose P (4x4)
__device__ __forceinline__ void pixel_dir(const float* __restrict__ P, const float* __restrict__ K, float x, float y, float* out) {
    const float fx = K[0], fy = K[5], cx = K[2], cy = K[6], sk = K[1];
    const float z = 1.0f;
    const float xl = (x - cx + cy * sk / fy - sk * y / fy) / fx * z;
    const float yl = (y - cy) / fy * z;
    float w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) w[k] = P[k * 4 + 0] * xl + P[k * 4 + 1] * yl + P[k * 4 + 2] * z + P[k * 4 + 3] - P[k * 4 + 3];
    float n = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    n = n > 1e-12f ? n : 1e-12f;
    out[0] = w[0] / n; out[1] = w[1] / n; out[2] = w[2] / n;
}

// One thread per sampled ray: everything SceneDatasetDN.__getitem__ + collate_fn hand the trainer in pixel mode
// (datasets/scene_dataset.py:374-401, 438-464), computed from the per-frame data instead of gathered from the
// per-ray arrays convert_to_pixels materialises (:269-307; ray_pose alone is 64 B per pixel there).
__global__ void k_pixel_batch(const int64_t* __restrict__ ray_ids, int64_t n, const float* __restrict__ poses,
                              const float* __restrict__ intr, int64_t n_frames, int H, int W, const float* __restrict__ rgb,
                              const float* __restrict__ depth, const float* __restrict__ mask, const float* __restrict__ normal,
                              float* __restrict__ ray_dirs, float* __restrict__ ray_dirs_tmp, float* __restrict__ cam_loc,
                              float* __restrict__ ray_pose, int64_t* __restrict__ frame_idx, float* __restrict__ g_rgb,
                              float* __restrict__ g_depth, float* __restrict__ g_mask, float* __restrict__ g_normal,
                              int* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t hw = (int64_t)H * W;
    int64_t r = ray_ids[i];
    if (r < 0 || r >= n_frames * hw) { atomicExch(bad, 1); r = 0; }       // reported by the host wrapper
    const int64_t f = r / hw, p = r - f * hw;
    const float x = (float)(p % W), y = (float)(p / W);                   // uv = (column, row): scene_dataset.py:258-260
    const float* P = poses + f * 16;
    const float* K = intr + f * 16;
    const float eye[16] = {1.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f, 1.f};
    pixel_dir(P, K, x, y, ray_dirs + 3 * i);
    pixel_dir(eye, K, x, y, ray_dirs_tmp + 3 * i);                        // un-rotated direction: depth scale (:286-288)
    cam_loc[3 * i] = P[3]; cam_loc[3 * i + 1] = P[7]; cam_loc[3 * i + 2] = P[11];
#pragma unroll
    for (int k = 0; k < 16; ++k) ray_pose[16 * i + k] = P[k];
    frame_idx[i] = f;
    if (g_rgb) { g_rgb[3 * i] = rgb[3 * r]; g_rgb[3 * i + 1] = rgb[3 * r + 1]; g_rgb[3 * i + 2] = rgb[3 * r + 2]; }
    if (g_depth) g_depth[i] = depth[r];
    if (g_mask) g_mask[i] = mask[r];
    if (g_normal) { g_normal[3 * i] = normal[3 * r]; g_normal[3 * i + 1] = normal[3 * r + 1]; g_normal[3 * i + 2] = normal[3 * r + 2]; }
}
}  // namespace

extern "C" int msdf_pixel_batch(const int64_t* ray_ids, int64_t n, const float* poses, const float* intrinsics, int64_t n_frames,
                                int height, int width, const float* rgb, const float* depth, const float* mask, const float* normal,
                                float* ray_dirs, float* ray_dirs_tmp, float* ray_cam_loc, float* ray_pose, int64_t* frame_idx,
                                float* gt_rgb, float* gt_depth, float* gt_mask, float* gt_normal, int* bad_flag, void* stream) {
    if (n == 0) return MSDF_OK;
    MSDF_CHECK_ARG(ray_ids && poses && intrinsics && ray_dirs && ray_dirs_tmp && ray_cam_loc && ray_pose && frame_idx && bad_flag,
                   "msdf_pixel_batch: null pointer");
    MSDF_CHECK_ARG(n_frames > 0 && height > 0 && width > 0, "msdf_pixel_batch: empty bank (%lld frames of %dx%d)", (long long)n_frames,
                   height, width);
    MSDF_CHECK_ARG((!gt_rgb || rgb) && (!gt_depth || depth) && (!gt_mask || mask) && (!gt_normal || normal),
                   "msdf_pixel_batch: a ground-truth output without its source image");
    k_pixel_batch<<<(unsigned)msdf_div_up(n, 128), 128, 0, (cudaStream_t)stream>>>(ray_ids, n, poses, intrinsics, n_frames, height, width, rgb,
                                                                                  depth, mask, normal, ray_dirs, ray_dirs_tmp, ray_cam_loc,
                                                                                  ray_pose, frame_idx, gt_rgb, gt_depth, gt_mask, gt_normal,
                                                                                  bad_flag);
    MSDF_COUNT_LAUNCH();
    MSDF_CHECK_LAUNCH("msdf_pixel_batch");
    return MSDF_OK;
}

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
