/*
 * monosdf_b200.h -- C ABI of libmonosdf_b200.so: hand-written sm_100a CUDA kernels for MonoSDF's
 * VolSDF-style volume-rendering hot path (MonoSDFNetwork.forward + backward).
 *
 * Conventions:
 *   - plain C: raw DEVICE pointers, sizes and scalars; no C++/torch types cross the boundary;
 *   - every function returns 0 on success, non-zero on an argument / launch error whose text is
 *     available from msdf_last_error() (thread-local); nothing throws;
 *   - the caller owns every buffer (outputs, gradients, workspaces); gradient buffers are ACCUMULATED
 *     into and must be zeroed by the caller (same contract as the reference op, hashgrid.py:75-76);
 *   - kernels are enqueued on the cudaStream_t passed as `void* stream` (the caller's current stream);
 *     no hidden synchronisation, no host<->device copies;
 *   - all tensors fp32 contiguous row-major unless a leading dimension `ld*` is given.
 *
 * Each entry point cites the reference code it replaces (paths relative to the reference's code/).
 */
#ifndef MONOSDF_B200_H
#define MONOSDF_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSDF_MAX_LAYERS 12
#define MSDF_ABI_VERSION 3

/* ------------------------------------------------------------------ library ------------------------------ */
const char* msdf_last_error(void);
int msdf_abi_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long msdf_launch_count(void);

/* Per-launch device timing of the dominant kernels (CUDA events on the launching stream), for bench.py's roofline
 * line.  class: 0 fp32 GEMM, 1 tensor-core kernels, 2 hash grid, 3 sampler, 4 compositing.  work = FLOPs (GEMMs) or
 * algorithmic bytes (the others) summed over the recorded launches.  Class 1 has sub-classes, read with
 * cls = 1 | k << 8: k = 1 fused SDF network (sdf-only), 2 fused SDF network (forward sweep of a step), 3 chained reverse
 * sweep, 4 per-layer sweeps with TMA-fed operands, 5 per-layer engine of round 1 (colour net), 6 weight gradients. */
int msdf_profile_enable(int on);
int msdf_profile_read(int cls, double* total_ms, double* total_work, long long* count, int reset);
int msdf_profile_read_bytes(int cls, double* total_bytes);   /* algorithmic bytes of the same launches */

/* ------------------------------------------------------------------ ray sampler ---------------------------
 * ErrorBoundSampler.get_z_vals, model/ray_sampler.py:110-262, split at the SDF evaluations.
 * Row buffers z/sdf are [n_rays, cap] with cap >= N_samples_eval * max_total_iters.                      */

/* UniformSampler.get_z_vals + near_far_from_cube (:48-83) and the Lemma-2 beta bound (:118-120).
 * t_vals[n0] = linspace(0,1,n0); t_rand[n_rays,n0] = stratified jitter (NULL in eval mode).
 * beta_coef = 1/(4 log(1+eps)).  Writes z[:, :n0], beta[n_rays], pts[n_rays*n0,3] = o + z*d.              */
int msdf_sampler_init(const float* ray_o, const float* ray_d, int64_t n_rays, const float* t_vals,
                      const float* t_rand, int n0, float bound, float near_, float far_max, float beta_coef,
                      float* z, int cap, float* beta, float* pts, void* stream);

/* One Algorithm-1 iteration up to the convergence test (:132-165,179): merges the n_new samples
 * (z_new, sdf_new) into the sorted row (n_old entries; n_old == 0 on the first call, where z already
 * holds the row and only sdf_new is consumed), computes d*, runs the beta bisection, stores beta and
 * ORs 1 into *flag if any ray still has beta > beta0 (the reference's batch-global beta.max() > beta0).
 * beta0 is a DEVICE pointer to density.get_beta() (one float). */
int msdf_sampler_round(int64_t n_rays, int n_old, int n_new, float* z, float* sdf, const float* z_new,
                       const float* sdf_new, int cap, const float* beta0, float eps, int beta_iters, float* beta,
                       unsigned int* flag, void* stream);

/* Not converged: n_new inverse-CDF samples of the error-bound pdf (:184-194,209-228) and their points. */
int msdf_sampler_upsample(int64_t n_rays, int n, const float* z, const float* sdf, int cap, const float* beta,
                          float add_tiny, const float* u, int n_new, const float* ray_o, const float* ray_d,
                          float* z_new, float* pts_new, void* stream);

/* Converged / out of iterations: n_s samples of the opacity pdf (:199-206), plus near, far and n_extra
 * picks z[:, pick[k]] (:238-249), sorted (:251) into z_out[n_rays, n_s+2+n_extra]; z_eik[r] =
 * z_out[r, eik_idx[r]] (:254-255, may be NULL).  u is [n_s] (u_per_ray = 0) or [n_rays, n_s].         */
int msdf_sampler_finalize(int64_t n_rays, int n, const float* z, const float* sdf, int cap, const float* beta,
                          const float* u, int u_per_ray, int n_s, const int32_t* pick, int n_extra, float near_,
                          float far_, const int64_t* eik_idx, float* z_out, float* z_eik, void* stream);

/* ------------------------------------------------------------------ hash grid -----------------------------
 * Same three operations, argument order and tensor layouts as the reference extension
 * (hashencoder/src/hashencoder.h:13-15, bindings.cpp:5-9): inputs [B,D] in [0,1], embeddings [sO,C],
 * offsets int32[L+1], outputs / grad [L,B,C], dy_dx [B, L*D*C].  fp32, D = 3, C in {1,2,4}.             */
int msdf_hash_encode_forward(const float* inputs, const float* embeddings, const int32_t* offsets, float* outputs,
                             uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                             int calc_grad_inputs, float* dy_dx, void* stream);
int msdf_hash_encode_backward(const float* grad, const float* inputs, const float* embeddings, const int32_t* offsets,
                              float* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                              uint32_t H, int calc_grad_inputs, const float* dy_dx, float* grad_inputs, void* stream);
int msdf_hash_encode_second_backward(const float* grad, const float* inputs, const float* embeddings,
                                     const int32_t* offsets, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                                     uint32_t H, int calc_grad_inputs, const float* dy_dx,
                                     const float* grad_grad_inputs, float* grad_grad, float* grad2_embeddings,
                                     void* stream);

/* ------------------------------------------------------------------ networks ------------------------------ */
/* A stack of Linear layers with EFFECTIVE weights (weight-norm already applied by msdf_weightnorm_forward). */
typedef struct {
    int32_t n_layers;
    int32_t d0;                          /* width of the encoded input h0                                   */
    int32_t skip_layer;                  /* layer whose input is cat([h, h0])/sqrt(2) (network.py:88-89), -1 */
    int32_t in_dim[MSDF_MAX_LAYERS];     /* columns of W_l (after the skip concat)                          */
    int32_t out_dim[MSDF_MAX_LAYERS];    /* rows of W_l                                                     */
    int32_t ldw[MSDF_MAX_LAYERS];        /* leading dimension of W_l (>= in_dim; multiples of 4 vectorise)  */
    const float* W[MSDF_MAX_LAYERS];     /* [out_dim, ldw] row-major                                        */
    const float* b[MSDF_MAX_LAYERS];     /* [out_dim]                                                       */
} msdf_mlp_desc;

typedef struct {
    float* dW[MSDF_MAX_LAYERS];          /* [out_dim, ldw], accumulated into (caller zeroes)                */
    float* db[MSDF_MAX_LAYERS];
} msdf_mlp_grads;

/* Input encoding of the SDF network: embedder.py:5-50 (+ HashEncoder, hashgrid.py:107-166 for Grid nets). */
typedef struct {
    int32_t multires;                    /* PE frequencies for xyz (0 = raw xyz)                            */
    int32_t grid_feat_dim;               /* 0 for ImplicitNetwork; L*C for ImplicitNetworkGrid              */
    int32_t n_levels, level_dim, base_res;
    float log2_per_level_scale;          /* S                                                               */
    float divide_factor;                 /* network.py:250                                                  */
    const float* table;                  /* embeddings; NULL -> zero features (use_grid_feature=False)      */
    const int32_t* offsets;
} msdf_encoding_desc;

/* RenderingNetwork.forward (network.py:389-470), mode 'idr' ([x, PE(view), normal, feat, code]) or 'nerf'. */
typedef struct {
    int32_t mode_idr;                    /* 1 = idr, 0 = nerf ([PE(view), feat, code], network.py:395-396)  */
    int32_t multires_view;
    int32_t feat_dim;                    /* feature_vector_size                                             */
    int32_t code_dim;                    /* 0 or 32 (per_image_code)                                        */
    int32_t code_per_ray;                /* 1: code[n_rays,code_dim] (:411-412); 0: code[1,code_dim] (:409) */
    int32_t final_act;                   /* 0 = sigmoid, 1 = relu (HDR, :465-468)                           */
    int32_t spec;                        /* 1 = diffuse/specular split (spec=True, :427-454; HDR only): every layer
                                          * is followed by ReLU, the first 3 outputs of layer L-3 are the diffuse
                                          * colour, the remaining ones feed layer L-2 (in_dim = out_dim - 3), the last
                                          * layer yields the specular colour; rgb = diffuse + specular.  The rgb
                                          * buffers of msdf_field_forward / _backward are then [M,6] = [rgb | rgb_spec]
                                          * (and d_rgb [M,6] = [dL/drgb | dL/drgb_spec]).                      */
} msdf_color_desc;

#define MSDF_MODE_SDF_ONLY 0   /* get_sdf_vals, network.py:131-137,307-309 (sampler, marching cubes)         */
#define MSDF_MODE_FORWARD 1    /* get_outputs / gradient_sdf (+ RenderingNetwork when color_net != NULL)     */
#define MSDF_MODE_BACKWARD 2   /* workspace query only: what msdf_field_backward needs                      */

#define MSDF_FLAG_TENSOR_BF16 1u   /* run the dense contractions on tcgen05 tensor cores in bf16 (2e-2 mode)  */

/* Bytes of workspace for a chunk of `chunk_points` points (the library processes M points in chunks of the
 * largest size that fits the workspace it is given; any size >= the 128-point figure works). */
size_t msdf_field_workspace_bytes(const msdf_mlp_desc* sdf_net, const msdf_encoding_desc* enc,
                                  const msdf_mlp_desc* color_net, const msdf_color_desc* cd, int64_t chunk_points,
                                  int mode, unsigned flags);

/* The neural field at M points: ImplicitNetwork / ImplicitNetworkGrid forward (network.py:79-96,247-275),
 * get_outputs (:111-129,290-305), gradient_sdf (:98-109,277-288), get_sdf_vals (:131-137,307-309) and, when
 * color_net != NULL, RenderingNetwork.forward (:389-470) on [x, PE(view), grad, feat, code].
 *   x [M,3]; sdf [M] (after the bounding-sphere clamp when clamp_radius > 0, :116-118); grad [M,3] = analytic
 *   d sdf / d x, replacing torch.autograd.grad(create_graph=True) (NULL in SDF_ONLY mode); feat [M, F] with
 *   leading dimension ld_feat (NULL to skip); rgb [M,3] (NULL when color_net == NULL);
 *   view_dirs [n_rays,3], point m belongs to ray m / n_samples; code [n_rays|1, code_dim] or NULL.
 * The library keeps no state.  With saved == NULL the backward recomputes the chunk's activations; with a caller
 * buffer of msdf_field_saved_bytes() bytes passed to BOTH calls (MODE_FORWARD with a grad output), the forward leaves
 * every layer's activations of both sweeps there (~10 KB per point in bf16 mode: B200's 180 GB of HBM hold a 65536-ray
 * step) and the backward reads them back -- and overwrites them: one backward per saved forward. */
size_t msdf_field_saved_bytes(const msdf_mlp_desc* sdf_net, const msdf_encoding_desc* enc, const msdf_mlp_desc* color_net,
                              const msdf_color_desc* cd, int64_t M, int n_samples, unsigned flags);

int msdf_field_forward(const msdf_mlp_desc* sdf_net, const msdf_encoding_desc* enc, const msdf_mlp_desc* color_net,
                       const msdf_color_desc* cd, const float* x, int64_t M, const float* view_dirs, int64_t n_rays,
                       int n_samples, const float* code, int mode, float clamp_radius, float sphere_scale,
                       unsigned flags, void* workspace, size_t workspace_bytes, float* sdf, float* grad,
                       float* feat, int64_t ld_feat, float* rgb, void* saved, size_t saved_bytes, void* stream);

/* Backward of the above INCLUDING the double-backward terms through grad (what loss.backward() does through
 * autograd.grad(create_graph=True) in the reference): given dL/dsdf [M], dL/dgrad [M,3], dL/dfeat [M,F] (ld),
 * dL/drgb [M,3] (any may be NULL = zero) and the forward's rgb [M,3], accumulates dL/dW, dL/db of both
 * networks and, for Grid nets, dL/dtable (grad_table, may be NULL); d_code [n_rays|1, code_dim] or NULL. */
int msdf_field_backward(const msdf_mlp_desc* sdf_net, const msdf_encoding_desc* enc, const msdf_mlp_desc* color_net,
                        const msdf_color_desc* cd, const float* x, int64_t M, const float* view_dirs, int64_t n_rays,
                        int n_samples, const float* code, float clamp_radius, float sphere_scale, unsigned flags,
                        void* workspace, size_t workspace_bytes, const float* d_sdf, const float* d_grad,
                        const float* d_feat, int64_t ld_dfeat, const float* rgb, const float* d_rgb,
                        const msdf_mlp_grads* sdf_grads, const msdf_mlp_grads* color_grads, float* grad_table,
                        float* d_code, void* saved, size_t saved_bytes, void* stream);

/* points[r*n + j] = o[r] + z[r,j] * d[r]   (network.py:532-533, ray_sampler.py:129) */
int msdf_ray_points(const float* ray_o, const float* ray_d, const float* z, int64_t n_rays, int n, float* points,
                    void* stream);

/* rend_util.get_camera_params + lift (utils/rend_util.py:63-91,105-118), 4x4 poses: uv [B,N,2], pose [B,4,4],
 * intrinsics [B,4,4] -> ray_dirs [B,N,3] (unit), cam_loc [B,3]. */
int msdf_camera_rays(const float* uv, const float* pose, const float* intrinsics, int64_t batch, int64_t n_pixels,
                     float* ray_dirs, float* cam_loc, void* stream);

/* Pixel-mode batch assembly on the device ("next" row f3): what SceneDatasetDN.convert_to_pixels / __getitem__ /
 * collate_fn (datasets/scene_dataset.py:269-307, 374-401, 438-464) hand the trainer, computed per sampled ray from the
 * per-frame data.  Ray id r addresses pixel p = r % (H*W) (row p / W, column p % W, uv = (column, row), :258-260) of
 * frame f = r / (H*W).  poses / intrinsics [n_frames,4,4]; rgb / normal [n_frames*H*W,3], depth / mask [n_frames*H*W]
 * (any may be NULL together with its output).  Outputs for the n ids: ray_dirs, ray_dirs_tmp (identity pose), ray_cam_loc
 * [n,3], ray_pose [n,4,4], frame_idx [n] int64 and the gathered ground-truth rows.  bad_flag: device int the caller
 * zeroes; set to 1 when an id is out of range (that ray then reads ray 0). */
int msdf_pixel_batch(const int64_t* ray_ids, int64_t n, const float* poses, const float* intrinsics, int64_t n_frames,
                     int height, int width, const float* rgb, const float* depth, const float* mask, const float* normal,
                     float* ray_dirs, float* ray_dirs_tmp, float* ray_cam_loc, float* ray_pose, int64_t* frame_idx,
                     float* gt_rgb, float* gt_depth, float* gt_mask, float* gt_normal, int* bad_flag, void* stream);

/* Coarse-to-fine SDF volume of one marching-cubes crop ("next" row f4; utils/plots.py:131-194, get_surface_sliding).
 * Level s (0 = finest) of a crop with crop_n samples per axis over [lo, hi] (HOST double[3]) has n = crop_n >> s cells
 * per axis, cell coordinates = mean of the 2^s fine linspace samples they cover.
 * msdf_sdfgrid_level_points: for every cell of the level whose parent cell is masked (parent_mask [(n/2)^3] bytes, NULL =
 *   all cells) append its point to points [<= n^3, 3] (order unspecified), store the list position in slot [n^3] (-1 =
 *   not evaluated) and count them in *counter (device int the caller zeroes).
 * msdf_sdfgrid_level_assemble: level[c] = values[slot[c]] if evaluated else parent[parent cell of c] (nearest upsample);
 *   mask[c] = |level[c]| < threshold (mask may be NULL for the finest level; parent may be NULL when every cell was
 *   evaluated). */
int msdf_sdfgrid_level_points(const double* lo, const double* hi, int crop_n, int level_shift, const unsigned char* parent_mask,
                              int* slot, float* points, int* counter, void* stream);
int msdf_sdfgrid_level_assemble(int n, const int* slot, const float* values, const float* parent, float threshold,
                                float* level, unsigned char* mask, void* stream);

/* ------------------------------------------------------------------ compositing ---------------------------
 * LaplaceDensity (density.py:21-30) + volume_rendering (network.py:626-640) + the weighted sums and the
 * normal-map rotation (network.py:552-562,603-616), one warp per ray.
 * beta: DEVICE pointer to density.get_beta() (one float; no host sync); depth_scale[r*stride] = ray_dirs_tmp z;
 * pose: matrices [n_rays,4,4] (pose_per_ray=1) or [1,4,4]; normal_map = R^T * sum w n; bg_color: device float[3]. */
int msdf_render_forward(const float* z_vals, const float* sdf, const float* rgb, const float* grad, int64_t n_rays,
                        int n_samples, const float* beta, const float* depth_scale, int64_t depth_scale_stride,
                        const float* pose, int pose_per_ray, int white_bkgd, const float* bg_color, float* weights,
                        float* rgb_values, float* depth_values, float* normal_map, void* stream);
/* Adjoint of the above: d_weights / d_rgb_values / d_depth_values / d_normal_map may be NULL (= zero);
 * writes d_sdf [n_rays*n_samples], d_rgb, d_grad [.,3] and accumulates d_beta[1] (caller zeroes). */
int msdf_render_backward(const float* z_vals, const float* sdf, const float* rgb, const float* grad, int64_t n_rays,
                         int n_samples, const float* beta, const float* depth_scale, int64_t depth_scale_stride,
                         const float* pose, int pose_per_ray, int white_bkgd, const float* bg_color,
                         const float* d_weights, const float* d_rgb_values, const float* d_depth_values,
                         const float* d_normal_map, float* d_sdf, float* d_rgb, float* d_grad, float* d_beta,
                         void* stream);

/* ------------------------------------------------------------------ parameters ----------------------------
 * nn.utils.weight_norm (dim=0), network.py:72-73: W[o,:] = g[o] * v[o,:] / ||v[o,:]||.                   */
int msdf_weightnorm_forward(const float* g, const float* v, int out_dim, int in_dim, float* W, int ldw, void* stream);
int msdf_weightnorm_backward(const float* g, const float* v, const float* dW, int ldw, int out_dim, int in_dim,
                             float* dg, float* dv, void* stream);
/* Adjoint of the per-image appearance-code lookup embeddings[indices] (network.py:400-413): d_table[indices[r], :] +=
 * d_code[r, :] with indices int64 [n_rays]; d_table [table_rows, code_dim] is accumulated into (caller zeroes). */
int msdf_code_scatter(const float* d_code, const int64_t* indices, int64_t n_rays, int code_dim, int64_t table_rows,
                      float* d_table, void* stream);
/* torch.optim.Adam step (monosdf_train.py:210-221) over a flat arena; grad_scale folds the 1/world_size of
 * the gradient all-reduce.  step is the 1-based step count. */
int msdf_fused_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                    void* stream);

/* ------------------------------------------------------------------ loss ----------------------------------
 * MonoSDFLoss.forward (model/loss.py:180-311, pixel-batch mode) and its gradient with respect to the renderer's
 * outputs, fused: L1 / MSE colour (optionally through gamma2, :209-215), eikonal (:222-224), smoothness (:226-234),
 * scale-and-shift-invariant depth (:29-49, :52-66, :236-243: target = 50 gt + 0.5, closed-form scale/shift over the
 * masked batch), normal L1 + cosine (:245-250), foreground mask = gt mask & (ray's sdf row changes sign) (:274-276).
 * No host synchronisation.  out[0..6] = loss, rgb, eikonal, smooth, depth, normal_l1, normal_cos; out[7] = number of
 * masked rays.  d_* receive d loss / d input (written, not accumulated).  workspace: 32 + n_rays floats, 8-byte aligned. */
typedef struct {
    float eikonal_weight, smooth_weight, depth_weight, normal_l1_weight, normal_cos_weight;
    float decay;                        /* exp(-step / end_step * 10) or 1 (:287-290) */
    int32_t rgb_mse;                    /* 0: L1Loss (mean), 1: MSELoss (mean) */
    int32_t gamma;                      /* if_gamma_loss */
    int32_t scale_invariant_depth;      /* if_scale_invariant_depth */
} msdf_loss_desc;
int msdf_loss_forward_backward(const msdf_loss_desc* desc, int64_t n_rays, int n_samples, const float* rgb_values,
                               const float* rgb_gt, const float* depth_values, const float* depth_gt, const float* gt_mask,
                               const float* normal_map, const float* normal_gt, const float* sdf, int64_t n_eik,
                               const float* grad_theta, const float* grad_theta_nei, float* workspace, float* out,
                               float* d_rgb_values, float* d_depth_values, float* d_normal_map, float* d_grad_theta,
                               float* d_grad_theta_nei, void* stream);

/* ------------------------------------------------------------------ tensor-core path (16-bit operands) -----
 * Self-test of the tcgen05 / TMEM / TMA GEMM engine (csrc/tc_gemm.cuh) against a naive kernel on the same 16-bit
 * inputs.  variant in [0, msdf_tc_selftest_count()) selects a shape and an operand-format combination (forward-type
 * GEMMs and weight-gradient GEMMs, fp16 / bf16 / mixed operands, ragged sizes included);
 * result_host[0] = max |C - ref|, result_host[1] = max |ref| (HOST pointer, the call synchronises). */
int msdf_tc_selftest(int variant, float* result_host, void* stream);
int msdf_tc_selftest_count(void);

/* Sdf-only queries of the tensor-core mode (MSDF_MODE_SDF_ONLY + MSDF_FLAG_TENSOR_BF16) run the whole SDF network as ONE
 * persistent tcgen05 kernel (csrc/fused_mlp.cuh: activations stay in shared / tensor memory; replaces the per-layer
 * nn.Linear + Softplus chain of network.py:79-96 behind get_sdf_vals :131-137 / :307-309) whenever the geometry allows
 * it (hidden width 256, PE multires 6 with 0 or 32 grid features).  msdf_set_fused(0) switches back to the per-layer
 * sweep (A/B measurements, tests); default 1. */
void msdf_set_fused(int on);

/* Backward-side sweeps of the tensor-core mode (the analytic replacement of autograd's double backward through
 * network.py:79-109): stream != 0 runs the per-layer kernels with every operand fed by TMA (csrc/tc_stream.cuh) instead
 * of the round-1 engine (csrc/tc_gemm.cuh); chain != 0 additionally runs the reverse sweep of a chunk as ONE launch with
 * its state on chip (csrc/tc_chain.cuh).  Same arithmetic on every path (A/B measurements, tests); default 1, 1. */
void msdf_set_sweeps(int stream, int chain);

#ifdef __cplusplus
}
#endif
#endif /* MONOSDF_B200_H */
