/* drosfm_b200 -- C ABI of the B200-native dense depth-pose warping path of dro-sfm.
 *
 * One shared library (dro_sfm_b200/libdrosfm_b200.so, sm_100a) exports exactly the entry points a
 * binding of the reference's hot path needs.  The reference (xyang9527/dro-sfm) is pure Python on
 * top of ATen, so it has no FFI of its own; each function below names the reference operator it
 * replaces (paths relative to the upstream repo root) and INTEGRATION.md shows the ctypes stub a
 * maintainer adds on the reference side.
 *
 * Conventions
 *  - All tensors are contiguous float32 device memory unless stated; images/features are NCHW.
 *  - The library never allocates, frees or synchronises: outputs and workspaces are caller
 *    owned, every call only enqueues kernels on `stream` (CUDA-graph capturable).
 *  - Return value: 0 = ok, <0 = argument error (DROSFM_E*), >0 = cudaError_t of the launch.
 *    drosfm_last_error() returns a thread-local message for the last non-zero return.
 *  - "ws" workspaces must be zero-filled once by the caller (drosfm_ws_bytes gives the size);
 *    kernels leave them zeroed again, so one allocation per (device, stream) can be reused.
 *  - Buffers marked "accumulated" must be zero-filled by the caller on the same stream before the
 *    call (they are scatter targets of red.global.add).
 *  - Gradient outputs may be NULL when not wanted.
 */
#ifndef DROSFM_B200_H_
#define DROSFM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DROSFM_ABI_VERSION 4
#define DROSFM_MAX_VIEWS 8      /* source views per call (forward_context + back_context) */
#define DROSFM_MAX_PREDS 16     /* depth predictions per loss call (GRU iterations seen by the loss) */
#define DROSFM_MAX_COST_JOBS 9  /* cost evaluations per batched launch: one depth cost + one pose cost per view */

enum { DROSFM_OK = 0, DROSFM_EINVAL = -1, DROSFM_ERANGE = -2, DROSFM_EALIGN = -3, DROSFM_ENOTSUP = -4 };
enum { DROSFM_POSE_IDENTITY = 0, DROSFM_POSE_MAT4 = 1, DROSFM_POSE_EULER6 = 2 };
enum { DROSFM_PAD_ZEROS = 0, DROSFM_PAD_BORDER = 1 };
enum { DROSFM_F32 = 0, DROSFM_F64 = 1 };
enum { DROSFM_DEPTH = 0, DROSFM_INV_DEPTH = 1, DROSFM_DISP = 2 };   /* DISP: see drosfm_cost_job_t */
enum { DROSFM_REDUCE_MIN = 0, DROSFM_REDUCE_MEAN = 1 };
enum { DROSFM_NCHW = 0, DROSFM_NHWC = 1 };
enum { DROSFM_ACCUMULATE_FMAP = 1 };   /* feat_cost_bwd flags */

typedef void* drosfm_stream_t;  /* cudaStream_t */

/* The camera pair of one warp: target camera (pose = identity unless Twc given) and source camera.
 * Replaces the per-call host work of the reference: K.float(), Camera.scaled (camera.py:83-107),
 * Camera.Kinv (camera.py:70-79), Pose.from_vec / euler2mat (pose.py:38-45, pose_utils.py:40-85). */
typedef struct {
    const void* K;        /* [B,3,3] target intrinsics, float32 or float64 (k_dtype)            */
    const void* Kref;     /* [B,3,3] source intrinsics, same dtype                              */
    int32_t k_dtype;      /* DROSFM_F32 | DROSFM_F64 (float64 is rounded to float32 = K.float()) */
    float sx, sy;         /* Camera.scaled factors; 1,1 = untouched intrinsics                  */
    const float* Twc;     /* [B,4,4] world<-target-camera, NULL = identity (Camera(K) default)  */
    const float* pose;    /* source camera Tcw: [B,4,4] (MAT4) or [B,6] (EULER6) or NULL        */
    int32_t pose_kind;    /* DROSFM_POSE_*                                                      */
} drosfm_cams_t;

int drosfm_version(void);
const char* drosfm_last_error(void);
/* Number of kernels this library has launched in the process so far (for benchmarks and tests that
 * must prove the CUDA path ran). */
unsigned long long drosfm_launch_count(void);
/* Bytes of zero-initialised workspace for calls that reduce pose gradients or scalar losses over
 * `slots` independent accumulators (one slot = one (sample, view, prediction) pose or one scalar). */
size_t drosfm_ws_bytes(int slots);

/* ---- Pose.from_vec(vec, 'euler') (pose.py:38-45, pose_utils.py:40-85) -----------------------
 * vec [N,6] = (tx,ty,tz,rx,ry,rz) -> mat [N,4,4] with R = Rx Ry Rz; one launch instead of ~40
 * ATen ops.  bwd: g_mat [N,4,4] -> g_vec [N,6].  The kernels that take DROSFM_POSE_EULER6 poses
 * evaluate exactly this conversion in their prologue. */
int drosfm_pose_vec2mat_fwd(const float* vec, float* mat, int N, drosfm_stream_t stream);
int drosfm_pose_vec2mat_bwd(const float* g_mat, const float* vec, float* g_vec, int N, drosfm_stream_t stream);

/* ---- Camera.reconstruct (camera.py:111-147) -------------------------------------------------
 * depth [B,1,H,W] -> points [B,3,H,W].  K: [B,3,3] (k_dtype).  Twc: [B,4,4] or NULL (frame 'c'). */
int drosfm_reconstruct_fwd(const float* depth, const void* K, int k_dtype, const float* Twc,
                           float* points, int B, int H, int W, drosfm_stream_t stream);
int drosfm_reconstruct_bwd(const float* g_points, const void* K, int k_dtype, const float* Twc,
                           float* g_depth, int B, int H, int W, drosfm_stream_t stream);

/* ---- Camera.project (camera.py:149-194) -----------------------------------------------------
 * points [B,3,H,W] -> uv [B,H,W,2].  Tcw: [B,4,4] or NULL (frame 'c').
 * bwd: g_points [B,3,H,W] and/or g_Tcw [B,4,4] (needs ws of drosfm_ws_bytes(B)). */
int drosfm_project_fwd(const float* points, const void* K, int k_dtype, const float* Tcw, float* uv,
                       int B, int H, int W, int normalize, drosfm_stream_t stream);
int drosfm_project_bwd(const float* g_uv, const float* points, const void* K, int k_dtype,
                       const float* Tcw, float* g_points, float* g_Tcw, void* ws,
                       int B, int H, int W, int normalize, drosfm_stream_t stream);

/* ---- fused reconstruct -> project (camera_utils.py:50-52, DepthPoseNet.py:83-90,
 *      supervised_loss.py:279-291) ------------------------------------------------------------
 * depth (or inverse depth, depth_kind) [B,1,H,W] -> uv [B,H,W,2]; optional mask [B,H,W,2] u8 =
 * (uv >= -1) & (uv <= 1).  bwd: g_depth [B,1,H,W]; g_pose [B,4,4] (MAT4) or [B,6] (EULER6). */
int drosfm_warp_coords_fwd(const float* depth, int depth_kind, const drosfm_cams_t* cams, float* uv,
                           uint8_t* mask, int B, int H, int W, int normalize, drosfm_stream_t stream);
int drosfm_warp_coords_bwd(const float* g_uv, const float* depth, int depth_kind, const drosfm_cams_t* cams,
                           float* g_depth, float* g_pose, void* ws, int B, int H, int W, int normalize,
                           drosfm_stream_t stream);

/* Device self-test (used by the parity tests): counts the floats x with 2^-126 <= |x| < 2^126 -- all 4.2e9 of them -- for
 * which the branch-free reciprocal of the fused kernels (MUFU.RCP + one Newton step) differs from rcp.rn.  `mismatches`
 * is one zero-initialised uint64 on the device; expected result 0. */
int drosfm_selftest_rcp(unsigned long long* mismatches, drosfm_stream_t stream);

/* ---- F.grid_sample(bilinear, align_corners=True) (camera_utils.py:55, DepthPoseNet.py:92) ---
 * src [B,C,Hs,Ws], uv [B,H,W,2] -> out [B,C,H,W].  bwd: g_src accumulated; g_uv written. */
int drosfm_grid_gather_fwd(const float* src, const float* uv, float* out, int B, int C, int Hs, int Ws,
                           int H, int W, int padding, drosfm_stream_t stream);
int drosfm_grid_gather_bwd(const float* g_out, const float* src, const float* uv, float* g_src, float* g_uv,
                           int B, int C, int Hs, int Ws, int H, int W, int padding, drosfm_stream_t stream);

/* ---- view_synthesis (camera_utils.py:23-56), coordinates never materialised -----------------
 * src [B,C,Hs,Ws], depth [B,1,H,W] -> out [B,C,H,W].
 * bwd: g_src accumulated (may be NULL), g_depth written, g_pose written (ws of drosfm_ws_bytes(B)). */
int drosfm_view_synthesis_fwd(const float* src, const float* depth, int depth_kind, const drosfm_cams_t* cams,
                              float* out, int B, int C, int Hs, int Ws, int H, int W, int padding,
                              drosfm_stream_t stream);
int drosfm_view_synthesis_bwd(const float* g_out, const float* src, const float* depth, int depth_kind,
                              const drosfm_cams_t* cams, float* g_src, float* g_depth, float* g_pose, void* ws,
                              int B, int C, int Hs, int Ws, int H, int W, int padding, drosfm_stream_t stream);

/* ---- feature-metric cost (DepthPoseNet.get_cost_each :76-96, depth_cost_calc :98-105) --------
 * cost[b,c,y,x] = (1/V) * sum_v (fmap - warp_v(fmap_ref_v))^2, zeros padding, V = n_views
 * (V = 1 is get_cost_each).  fmap, fmap_ref[v], cost: [B,C,h,w] in `layout` (NCHW, or NHWC =
 * torch channels_last storage of the same logical tensor).  poses[v]: per-view source pose in
 * cams->pose_kind encoding ([B,4,4] or [B,6]); cams->pose is ignored.
 * bwd: g_fmap written (added to when flags & DROSFM_ACCUMULATE_FMAP: lets a caller sum the gradients of
 * all cost calls of a step in one buffer); g_fmap_ref[v] and g_depth accumulated (entries may be NULL);
 * g_poses[v] written ([B,4,4] or [B,6]; entries may be NULL; ws of drosfm_ws_bytes(V*B)). */
int drosfm_feat_cost_fwd(const float* fmap, const float* const* fmap_ref, const float* depth, int depth_kind,
                         const drosfm_cams_t* cams, const float* const* poses, int n_views, float* cost,
                         int B, int C, int h, int w, int layout, drosfm_stream_t stream);
int drosfm_feat_cost_bwd(const float* g_cost, const float* fmap, const float* const* fmap_ref,
                         const float* depth, int depth_kind, const drosfm_cams_t* cams,
                         const float* const* poses, int n_views, float* g_fmap, float* const* g_fmap_ref,
                         float* g_depth, float* const* g_poses, void* ws,
                         int B, int C, int h, int w, int layout, int flags, drosfm_stream_t stream);

/* ---- a batch of independent feature-metric cost calls in ONE launch ---------------------------
 * Within one step of the recurrent optimiser the depth cost (depth_cost_calc: V views, mean) and the V per-view pose
 * costs (get_cost_each) depend only on the state at the start of the step: DepthPoseNet.forward builds every cost
 * closure (DepthPoseNet.py:159-167) before either update block runs (:171-192, update.py:161,189).  A caller that
 * advances the update blocks in lock-step (dro_sfm_b200/networks/lockstep.py) hands all of them over at once; at the
 * training shapes one call alone is a one-wave launch bound by latency.  Each job is one drosfm_feat_cost_{fwd,bwd}
 * call with the same B, C, h, w, cams (intrinsics, scale, pose encoding) and the NHWC layout; jobs may share feature
 * maps.  Gradient buffers shared by several jobs must be zero-filled and flagged DROSFM_ACCUMULATE_FMAP (g_fmap) --
 * g_fmap_ref / g_depth follow the rules of drosfm_feat_cost_bwd.  ws: drosfm_ws_bytes(B * sum of n_views). */
typedef struct {
    const float* fmap;               /* [B,C,h,w] target features (NHWC storage)                   */
    const float* const* fmap_ref;    /* n_views source maps                                         */
    const float* depth;              /* [B,1,h,w] depth, inverse depth or raw disparity             */
    int32_t depth_kind;              /* DROSFM_DEPTH | DROSFM_INV_DEPTH | DROSFM_DISP               */
    int32_t n_views;                 /* 1 = get_cost_each, V = depth_cost_calc (mean over views)    */
    const float* const* poses;       /* n_views poses in cams->pose_kind encoding                   */
    float* cost;                     /* [B,C,h,w] output (ignored by the backward)                  */
    /* DROSFM_DISP: `depth` is the network's disparity in [0,1]; the kernel evaluates disp_to_depth as the reference does
     * (networks/layers/resnet/layers.py:11-20, DepthPoseNet.py:38-41: inverse depth = disp_min + disp_range * disp with
     * disp_min = 1/max_depth, disp_range = 1/min_depth - 1/max_depth, both rounded to float32) and then inv2depth;
     * g_depth is the gradient w.r.t. the raw disparity. */
    float disp_min, disp_range;
} drosfm_cost_job_t;
typedef struct {
    const float* g_cost;             /* upstream gradient of the job's cost map                     */
    float* g_fmap;                   /* written, or added to with DROSFM_ACCUMULATE_FMAP; may be NULL */
    float* const* g_fmap_ref;        /* accumulated; entries / the array may be NULL                */
    float* g_depth;                  /* written; may be NULL                                        */
    float* const* g_poses;           /* written; entries / the array may be NULL                    */
    int32_t flags;                   /* DROSFM_ACCUMULATE_FMAP                                      */
} drosfm_cost_job_grads_t;
int drosfm_feat_cost_batch_fwd(const drosfm_cost_job_t* jobs, int n_jobs, const drosfm_cams_t* cams,
                               int B, int C, int h, int w, int layout, drosfm_stream_t stream);
int drosfm_feat_cost_batch_bwd(const drosfm_cost_job_t* jobs, const drosfm_cost_job_grads_t* grads, int n_jobs,
                               const drosfm_cams_t* cams, void* ws, int B, int C, int h, int w, int layout,
                               drosfm_stream_t stream);

/* ---- photometric loss (multiview_photometric_loss_mf.py:15-54,132-269,333-353) ---------------
 * Options shared by the photometric entry points. */
typedef struct {
    float ssim_w;        /* ssim_loss_weight (0.85) */
    float C1, C2;        /* SSIM constants (1e-4, 9e-4) */
    int32_t padding;     /* DROSFM_PAD_* for the warp */
    int32_t reduce_op;   /* DROSFM_REDUCE_MIN | DROSFM_REDUCE_MEAN */
    int32_t automask;    /* 1: un-warped source maps compete in the per-pixel min (min only) */
    float gamma;         /* decay over predictions: weight_i = gamma^(n-1-i) (0.85) */
    /* clip_loss > 0 (calc_photometric_loss, multiview_photometric_loss_mf.py:220-227): every photometric map is clamped at
     * mean + clip_loss * std of that map (unbiased std over B*H*W).  Costs one extra statistics pass per call; runs on the
     * fused path only (warped_save == NULL).  clip_scratch: 8 * (n_views + n_preds * n_views) floats of device memory,
     * 8-byte aligned, ZERO-FILLED by the caller before drosfm_automask_fwd / drosfm_photometric_fwd of a step (the first
     * n_views slots belong to the un-warped maps).  With clip_loss > 0, sel is required for both reduce ops: for 'mean' it
     * receives the bit mask of the views that were NOT clipped at the pixel (their gradient survives). */
    float clip_loss;
    float* clip_scratch;
} drosfm_photo_opts_t;

/* flags of drosfm_photometric_fwd / _bwd (staged path only) */
#define DROSFM_PHOTO_WARPED_READY 1 /* fwd: warped_save already holds drosfm_warp_sources_fwd's output */
#define DROSFM_PHOTO_NO_ADJOINT 2   /* bwd: stop after g_warped; the caller runs drosfm_warp_sources_bwd itself */
#define DROSFM_PHOTO_FUSE_BWD 4     /* fwd (training, 2/4/6/8 views): also write g_warped = d loss / d warped, UNSCALED by the
                                     * loss's upstream gradient -- the backward of the loss is then drosfm_warp_sources_bwd alone
                                     * (with g_scale = that upstream gradient); drosfm_photometric_bwd is not called */

/* Un-warped (auto-mask) pass: automask[b,y,x] = min_v photometric(context_v, image), computed once
 * per step instead of once per prediction (lines 346-351 recompute it n times). */
int drosfm_automask_fwd(const float* image, const float* const* context, int n_views,
                        const drosfm_photo_opts_t* opts, float* automask, int B, int H, int W,
                        drosfm_stream_t stream);
/* loss = sum_i gamma^(n-1-i) * reduce_v,pixels(photometric(warp(context_v; inv_depth_i, pose_{v,i}), image)).
 * inv_depths[i]: [B,1,H,W]; poses[v*n_preds+i]: [B,4,4] or [B,6] per cams->pose_kind.
 * sel [n_preds,B,H,W] u8 receives the arg-min view per pixel (255 = auto-mask won); loss: 1 float.
 * ws of drosfm_ws_bytes(n_preds + 1).
 * warped_save [n_preds,V,B,3,H,W] float, optional: when given, the call runs staged -- one flat kernel warps every
 * source view once into warped_save (no tile halos, the geometry runs once per pixel), a second one evaluates the
 * SSIM / L1 / min maps from it -- and the buffer is what drosfm_photometric_bwd wants back.  NULL: one fused
 * kernel that keeps nothing (inference, or when 12 bytes per pixel, view and prediction are too much). */
int drosfm_photometric_fwd(const float* image, const float* const* context, int n_views,
                           const float* const* inv_depths, int depth_kind, int n_preds,
                           const drosfm_cams_t* cams, const float* const* poses, const float* automask,
                           const drosfm_photo_opts_t* opts, uint8_t* sel, float* loss, void* ws,
                           float* warped_save, float* g_warped, int flags, int B, int H, int W, drosfm_stream_t stream);
/* g_loss: 1 float on the device (upstream gradient).  g_inv_depths[i] [B,1,H,W] written;
 * g_poses[v*n_preds+i] written ([B,4,4] or [B,6]); ws of drosfm_ws_bytes(n_views*n_preds*B).
 * warped_save (the forward's) and g_warped (scratch of the same size, contents undefined on return) go together:
 * given, the backward runs staged (window-gradient kernel -> g_warped -> flat warp adjoint); both NULL: one fused
 * kernel that re-warps every tile. */
int drosfm_photometric_bwd(const float* g_loss, const float* image, const float* const* context, int n_views,
                           const float* const* inv_depths, int depth_kind, int n_preds,
                           const drosfm_cams_t* cams, const float* const* poses, const uint8_t* sel,
                           const drosfm_photo_opts_t* opts, float* const* g_inv_depths, float* const* g_poses,
                           void* ws, const float* warped_save, float* g_warped, int flags, int B, int H, int W,
                           drosfm_stream_t stream);

/* The two warp stages of the staged path on their own, so that a caller can overlap the independent parts of the
 * loss (auto-mask, smoothness) on a second stream (multiview_photometric_loss_mf.py:132-171, warp_ref_image):
 * fwd: warped [n_preds,V,B,3,H,W] = view_synthesis(context_v; inv_depth_i, pose_{v,i}) for every (i, v); hand it to
 *      drosfm_photometric_fwd with DROSFM_PHOTO_WARPED_READY.
 * bwd: adjoint of that warp for the g_warped a DROSFM_PHOTO_NO_ADJOINT backward left behind: g_inv_depths[i] written
 *      (accumulate != 0: added to), g_poses[v*n_preds+i] written; ws of drosfm_ws_bytes(n_views*n_preds*B).
 * g_scale (bwd, optional): one float on the device that multiplies g_warped (the loss's upstream gradient when g_warped
 *      came from a DROSFM_PHOTO_FUSE_BWD forward); NULL = 1.
 * rgbx (optional scratch, [V,B,H,W,4] floats, 16-byte aligned): when given, fwd first packs the source pictures into
 *      RGBx texels there (one more launch) and both directions gather each bilinear tap with ONE 128-bit load instead
 *      of three 32-bit ones; bwd expects the buffer fwd filled.  NULL: gathers from the caller's planes. */
int drosfm_warp_sources_fwd(const float* const* context, int n_views, const float* const* inv_depths, int depth_kind,
                            int n_preds, const drosfm_cams_t* cams, const float* const* poses, int padding, float* rgbx,
                            float* warped, int B, int H, int W, drosfm_stream_t stream);
int drosfm_warp_sources_bwd(const float* g_warped, const float* const* context, int n_views,
                            const float* const* inv_depths, int depth_kind, int n_preds, const drosfm_cams_t* cams,
                            const float* const* poses, int padding, const float* rgbx, const float* g_scale,
                            float* const* g_inv_depths, float* const* g_poses, void* ws, int accumulate, int B, int H, int W,
                            drosfm_stream_t stream);

/* ---- smoothness loss (multiview_photometric_loss_mf.py:273-299, utils/depth.py:147-199) -------
 * loss = weight/n * sum_i (mean|dx(d_i/mean(d_i)) * wx| + mean|dy(..) * wy|) / 2^i.
 * stats [n_preds,B,4] float scratch written by fwd and read by bwd (per-sample mean inverse depth
 * and the two per-sample edge sums); ws of drosfm_ws_bytes(n_preds*B + 1).
 * bwd: g_inv_depths[i] [B,1,H,W] written (accumulate 0), added to (1), or added to with atomic reductions (2: another kernel
 * -- the warp adjoint of the photometric term -- may add into the same maps concurrently on a second stream); entries may be NULL.
 * edge_w (optional, [B,2,H,W] floats): fwd leaves the edge weights exp(-mean_c |dI|) towards the right / lower neighbour
 * there and bwd reads them instead of re-deriving them from the image (4 loads instead of 24 loads + 4 exponentials per pixel). */
int drosfm_smoothness_fwd(const float* image, const float* const* inv_depths, int n_preds, float weight,
                          float* stats, float* loss, void* ws, float* edge_w, int B, int H, int W, drosfm_stream_t stream);
int drosfm_smoothness_bwd(const float* g_loss, const float* image, const float* const* inv_depths, int n_preds,
                          float weight, const float* stats, float* const* g_inv_depths, int accumulate, const float* edge_w,
                          int B, int H, int W, drosfm_stream_t stream);

/* ---- reprojection pose loss (supervised_loss.py:279-325) ------------------------------------
 * loss = sum_i w_i/V * sum_v mean(valid * clamp(|uv(pred_{v,i}) - uv(gt_v)|, -1, 1)) / sum_i w_i,
 * valid = in-range(gt) & in-range(pred) & (min_depth < depth < max_depth/4), depth = GT depth
 * [B,1,H,W] (depth_kind may be DROSFM_INV_DEPTH: inv2depth is fused).  gt_poses[v], pred_poses[v*n+i]
 * in cams->pose_kind encoding.  fwd: ws of drosfm_ws_bytes(n + 1).  bwd: g_pred_poses[v*n+i] written
 * (entries may be NULL); ws of drosfm_ws_bytes(V*n*B). */
int drosfm_reproj_loss_fwd(const float* depth, int depth_kind, const drosfm_cams_t* cams,
                           const float* const* gt_poses, const float* const* pred_poses, int n_views, int n_preds,
                           float min_depth, float max_depth, float gamma, float* loss, void* ws,
                           int B, int H, int W, drosfm_stream_t stream);
int drosfm_reproj_loss_bwd(const float* g_loss, const float* depth, int depth_kind, const drosfm_cams_t* cams,
                           const float* const* gt_poses, const float* const* pred_poses, int n_views, int n_preds,
                           float min_depth, float max_depth, float gamma, float* const* g_pred_poses, void* ws,
                           int B, int H, int W, drosfm_stream_t stream);

/* ---- supervised depth loss (supervised_loss.py:244-277) ----------------------------------------
 * loss = sum_i w_i * mean(valid * |gt_inv - inv_depth_i|) / sum_i w_i, valid = 1/max_depth < gt_inv < 1/min_depth,
 * w_i = gamma^(n-1-i); all maps [B,1,H,W].  fwd: ws of drosfm_ws_bytes(n_preds + 1).
 * bwd: g_inv_depths[i] written (entries may be NULL). */
int drosfm_sup_depth_loss_fwd(const float* gt_inv_depth, const float* const* inv_depths, int n_preds, float min_depth,
                              float max_depth, float gamma, float* loss, void* ws, int B, int H, int W,
                              drosfm_stream_t stream);
int drosfm_sup_depth_loss_bwd(const float* g_loss, const float* gt_inv_depth, const float* const* inv_depths, int n_preds,
                              float min_depth, float max_depth, float gamma, float* const* g_inv_depths, int B, int H, int W,
                              drosfm_stream_t stream);

/* ---- convex up-sampling (DepthPoseNet.upsample_depth, DepthPoseNet.py:63-74) ------
 * depth [N,1,H,W], mask [N,9*ratio*ratio,H,W] -> out [N,1,ratio*H,ratio*W]:
 * out[8y+i,8x+j] = sum_k softmax_k(mask[k*64+i*8+j, y, x]) * depth_zero_padded[y+k/3-1, x+k%3-1].  ratio must be 8.
 * bwd: g_mask written, g_depth accumulated (either may be NULL).
 * disp_min / disp_range: the disp_to_depth scaling that follows every up-sampling in DepthPoseNet.forward
 * (scale_inv_depth, DepthPoseNet.py:38-41,128,180; layers.py:11-20) as an epilogue: out = disp_min + disp_range * up
 * (0, 1 = plain up-sampling). */
int drosfm_upsample_depth_fwd(const float* depth, const float* mask, float* out, int N, int H, int W, int ratio,
                              float disp_min, float disp_range, drosfm_stream_t stream);
int drosfm_upsample_depth_bwd(const float* g_out, const float* depth, const float* mask, float* g_depth, float* g_mask,
                              int N, int H, int W, int ratio, float disp_range, drosfm_stream_t stream);

/* ---- evaluation path (ModelWrapper.evaluate_depth, models/model_wrapper.py:355-399) -------
 * post_process_inv_depth (dro_sfm/utils/depth.py:230-258): out = mask_hat * inv + mask * flip(inv_flipped) +
 * (1 - mask - mask_hat) * fuse(inv, flip(inv_flipped)); inv, inv_flipped, out: [B,1,H,W]; method 0 mean, 1 max, 2 min. */
int drosfm_post_process_inv_depth(const float* inv_depth, const float* inv_depth_flipped, float* out, int B, int H, int W,
                                  int method, drosfm_stream_t stream);
/* compute_depth_metrics (dro_sfm/utils/depth.py:261-340): gt [B,1,H,W], pred [B,1,Hp,Wp] (interpolated to H x W inside:
 * bilinear, align_corners=True; clamped at 1e-6) -> metrics[9] = abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3, SILog,
 * iabs_diff, each the batch average of the per-sample value over its valid pixels (min_depth < gt < max_depth inside the
 * crop: 0 none, 1 'garg', 2 'eigen_nyu').  use_gt_scale: every sample's prediction is scaled by the median of gt / pred
 * over its valid pixels -- the exact lower median torch.median returns, found by a three-pass radix selection.
 * ws: drosfm_eval_ws_bytes(B) bytes, zero-filled once (left zeroed). */
size_t drosfm_eval_ws_bytes(int B);
int drosfm_depth_metrics(const float* gt, const float* pred, int B, int H, int W, int Hp, int Wp, float min_depth, float max_depth,
                         int crop, int use_gt_scale, float* metrics, void* ws, drosfm_stream_t stream);

/* ---- feature-map storage layout (the encoder's maps are NCHW, DepthPoseNet.py:113-115) -----------
 * Copies a [B,C,H,W] tensor from NCHW storage to NHWC storage (to_layout = DROSFM_NHWC) or back (DROSFM_NCHW);
 * what `.contiguous(memory_format=torch.channels_last)` / `.contiguous()` do, as a coalesced tiled transpose. */
int drosfm_relayout(const float* src, float* dst, int B, int C, int H, int W, int to_layout, drosfm_stream_t stream);

/* ---- 8-bit pictures -> float32 (ToTensor, dro_sfm/datasets/augmentations.py:149-152) ---------------
 * dst[i] = src[i] / 255 (IEEE division) for n elements: lets a data loader ship the target and source pictures as
 * uint8 (a quarter of the host-to-device bytes) and convert them in one launch on the device. */
int drosfm_images_u8_to_f32(const uint8_t* src, float* dst, size_t n, drosfm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DROSFM_B200_H_ */
