// Kernel family 1: back-project -> rigid transform -> project.
//
//   drosfm_reconstruct_*  : Camera.reconstruct            (dro_sfm/geometry/camera.py:111-147)
//   drosfm_project_*      : Camera.project                (dro_sfm/geometry/camera.py:149-194)
//   drosfm_warp_coords_*  : the fused composition used by view_synthesis (camera_utils.py:50-52),
//                           get_cost_each (DepthPoseNet.py:83-90) and get_ref_coords
//                           (supervised_loss.py:279-291); points are never materialised and the
//                           pixel grid (utils/image.py:304-332) is implicit.
//
// HBM-bound: fwd 12 B/px (depth 4 + uv 8), bwd 16 B/px (g_uv 8 + depth 4 + g_depth 4).
// One thread handles 4 consecutive pixels with 128-bit loads/stores; grid = (chunks, B).
#include <cstdlib>

#include "common.cuh"

namespace drosfm {

constexpr int kThreads = 256;

__device__ __forceinline__ void pix_xy(int p, int W, int& x, int& y) {
    y = p / W;
    x = p - y * W;
}

// ------------------------------------------------------------------------------------------
// fused forward
// ------------------------------------------------------------------------------------------
// CHAIN 0: the plain div.rn chain (default).  1: the divisions of the projection go through div_by_rcp (supervised
// losses, fused tile kernels).  2: warp_pixel_fast, the branch-free variant of the flat warp / cost kernels.  All give the
// same bits (2: wherever |coordinate| < 1e30); DROSFM_COORDS_SHARED_RCP=1|2 selects them so that the tests can hold
// every chain the fused kernels use to the same bit-exact oracle.
template <bool VEC, int CHAIN>
__global__ void __launch_bounds__(kThreads)
warp_coords_fwd_kernel(const float* __restrict__ depth, int depth_kind, drosfm_cams_t cams,
                       float* __restrict__ uv, uint8_t* __restrict__ mask, int H, int W, int normalize) {
    __shared__ Cam cam;
    const int b = blockIdx.y;
    const int P = H * W;
    const float wm1 = static_cast<float>(W - 1), hm1 = static_cast<float>(H - 1);
    constexpr int PX = VEC ? 4 : 1;
    const int p0 = (blockIdx.x * kThreads + threadIdx.x) * PX;
    float d[PX];
    if (p0 < P) {
        if constexpr (VEC) {
            const float4 v = ldg4(depth + static_cast<size_t>(b) * P + p0);
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        } else {
            d[0] = __ldg(depth + static_cast<size_t>(b) * P + p0);
        }
    }
    if (threadIdx.x == 0) setup_cam(cams, cams.pose, b, cam);
    __syncthreads();
    if (p0 >= P) return;
    float out[2 * PX];
#pragma unroll
    for (int k = 0; k < PX; ++k) {
        int x, y;
        pix_xy(p0 + k, W, x, y);
        Warp w;
        if constexpr (CHAIN == 2) warp_pixel_fast(cam, x, y, to_depth_fast(d[k], depth_kind), make_norm(W, H), normalize != 0, w);
        else warp_pixel<CHAIN == 1>(cam, x, y, to_depth(d[k], depth_kind), wm1, hm1, normalize != 0, w);
        out[2 * k] = w.p.u;
        out[2 * k + 1] = w.p.v;
    }
    float* dst = uv + (static_cast<size_t>(b) * P + p0) * 2;
    if constexpr (VEC) {
        st4_streaming(dst, make_float4(out[0], out[1], out[2], out[3]));
        st4_streaming(dst + 4, make_float4(out[4], out[5], out[6], out[7]));
    } else {
        dst[0] = out[0];
        dst[1] = out[1];
    }
    if (mask != nullptr) {
        uint8_t* m = mask + (static_cast<size_t>(b) * P + p0) * 2;
#pragma unroll
        for (int k = 0; k < 2 * PX; ++k) m[k] = (out[k] >= -1.0f && out[k] <= 1.0f) ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------
// fused backward: persistent blocks per sample, pose gradient reduced in fp64
// ------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(kThreads)
warp_coords_bwd_kernel(const float* __restrict__ g_uv, const float* __restrict__ depth, int depth_kind,
                       drosfm_cams_t cams, float* __restrict__ g_depth, float* __restrict__ g_pose, Slot* ws,
                       int H, int W, int normalize) {
    __shared__ Cam cam;
    __shared__ double red[12 * (kThreads / 32)];
    __shared__ int flag;
    const int b = blockIdx.y;
    const int P = H * W;
    const float wm1 = static_cast<float>(W - 1), hm1 = static_cast<float>(H - 1);
    constexpr int PX = VEC ? 4 : 1;
    if (threadIdx.x == 0) setup_cam(cams, cams.pose, b, cam);
    __syncthreads();
    float gT[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) gT[i] = 0.0f;
    const size_t base = static_cast<size_t>(b) * P;
    for (int p0 = (blockIdx.x * kThreads + threadIdx.x) * PX; p0 < P; p0 += gridDim.x * kThreads * PX) {
        float d[PX], g[2 * PX], gd[PX];
        if constexpr (VEC) {
            const float4 v = ldg4(depth + base + p0);
            const float4 a = ldg4(g_uv + (base + p0) * 2), c = ldg4(g_uv + (base + p0) * 2 + 4);
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
            g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w;
            g[4] = c.x; g[5] = c.y; g[6] = c.z; g[7] = c.w;
        } else {
            d[0] = __ldg(depth + base + p0);
            g[0] = __ldg(g_uv + (base + p0) * 2);
            g[1] = __ldg(g_uv + (base + p0) * 2 + 1);
        }
#pragma unroll
        for (int k = 0; k < PX; ++k) {
            int x, y;
            pix_xy(p0 + k, W, x, y);
            Warp w;
            warp_pixel(cam, x, y, to_depth(d[k], depth_kind), wm1, hm1, normalize != 0, w);
            const float gdk = warp_pixel_adjoint(cam, w, to_depth(d[k], depth_kind), wm1, hm1, normalize != 0, g[2 * k], g[2 * k + 1], gT);
            gd[k] = depth_kind == DROSFM_INV_DEPTH ? inv2depth_grad(d[k], gdk) : gdk;
        }
        if (g_depth != nullptr) {
            if constexpr (VEC) st4(g_depth + base + p0, make_float4(gd[0], gd[1], gd[2], gd[3]));
            else g_depth[base + p0] = gd[0];
        }
    }
    if (g_pose == nullptr) return;
    Slot* slot = slot_at(ws, b);
    block_accumulate12(gT, red, spread_acc(slot));
    if (last_block(slot, gridDim.x, &flag) && threadIdx.x < 32) {
        const int stride = cams.pose_kind == DROSFM_POSE_EULER6 ? 6 : 16;
        finish_pose_grad_warp(slot, cams.pose_kind, cams.pose_kind == DROSFM_POSE_EULER6 ? cams.pose + b * 6 : nullptr,
                         g_pose + b * stride);
    }
}

// ------------------------------------------------------------------------------------------
// Camera.reconstruct
// ------------------------------------------------------------------------------------------
struct RecCam {
    float Ki[9];
    float Rt[12];
};

__device__ __forceinline__ void setup_rec(const void* K, int k_dtype, const float* Twc, int b, RecCam& c) {
    float Kt[9];
    load_scaled_K(K, k_dtype, b, 1.0f, 1.0f, Kt);
    invert_K(Kt, c.Ki);
    if (Twc != nullptr) load_mat34(Twc, b, c.Rt); else identity34(c.Rt);
}

__global__ void __launch_bounds__(kThreads)
reconstruct_fwd_kernel(const float* __restrict__ depth, const void* K, int k_dtype, const float* Twc,
                       float* __restrict__ points, int H, int W) {
    __shared__ RecCam c;
    const int b = blockIdx.y, P = H * W;
    if (threadIdx.x == 0) setup_rec(K, k_dtype, Twc, b, c);
    __syncthreads();
    const int p = blockIdx.x * kThreads + threadIdx.x;
    if (p >= P) return;
    int x, y;
    pix_xy(p, W, x, y);
    const float d = __ldg(depth + static_cast<size_t>(b) * P + p);
    const float fx = static_cast<float>(x), fy = static_cast<float>(y);
    float Xc[3], Xw[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) Xc[k] = __fmul_rn(dot3(c.Ki + 3 * k, fx, fy, 1.0f), d);
    if (Twc != nullptr) rigid(c.Rt, Xc, Xw);
    else { Xw[0] = Xc[0]; Xw[1] = Xc[1]; Xw[2] = Xc[2]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) points[(static_cast<size_t>(b) * 3 + k) * P + p] = Xw[k];
}

__global__ void __launch_bounds__(kThreads)
reconstruct_bwd_kernel(const float* __restrict__ g_points, const void* K, int k_dtype, const float* Twc,
                       float* __restrict__ g_depth, int H, int W) {
    __shared__ RecCam c;
    const int b = blockIdx.y, P = H * W;
    if (threadIdx.x == 0) setup_rec(K, k_dtype, Twc, b, c);
    __syncthreads();
    const int p = blockIdx.x * kThreads + threadIdx.x;
    if (p >= P) return;
    int x, y;
    pix_xy(p, W, x, y);
    const float fx = static_cast<float>(x), fy = static_cast<float>(y);
    float g[3], gc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) g[k] = __ldg(g_points + (static_cast<size_t>(b) * 3 + k) * P + p);
    if (Twc != nullptr) rigid_adjoint(c.Rt, g, gc);
    else { gc[0] = g[0]; gc[1] = g[1]; gc[2] = g[2]; }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) s += gc[k] * dot3(c.Ki + 3 * k, fx, fy, 1.0f);
    g_depth[static_cast<size_t>(b) * P + p] = s;
}

// ------------------------------------------------------------------------------------------
// Camera.project
// ------------------------------------------------------------------------------------------
struct ProjCam {
    float Kr[9];
    float T[12];
};

__device__ __forceinline__ void setup_proj(const void* K, int k_dtype, const float* Tcw, int b, ProjCam& c) {
    load_scaled_K(K, k_dtype, b, 1.0f, 1.0f, c.Kr);
    if (Tcw != nullptr) load_mat34(Tcw, b, c.T); else identity34(c.T);
}

__global__ void __launch_bounds__(kThreads)
project_fwd_kernel(const float* __restrict__ points, const void* K, int k_dtype, const float* Tcw,
                   float* __restrict__ uv, int H, int W, int normalize) {
    __shared__ ProjCam c;
    const int b = blockIdx.y, P = H * W;
    if (threadIdx.x == 0) setup_proj(K, k_dtype, Tcw, b, c);
    __syncthreads();
    const int p = blockIdx.x * kThreads + threadIdx.x;
    if (p >= P) return;
    float X[3], Y[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) X[k] = __ldg(points + (static_cast<size_t>(b) * 3 + k) * P + p);
    if (Tcw != nullptr) rigid(c.T, X, Y);
    else { Y[0] = X[0]; Y[1] = X[1]; Y[2] = X[2]; }
    Proj pr;
    project_cam(c.Kr, Y, static_cast<float>(W - 1), static_cast<float>(H - 1), normalize != 0, pr);
    reinterpret_cast<float2*>(uv)[static_cast<size_t>(b) * P + p] = make_float2(pr.u, pr.v);
}

__global__ void __launch_bounds__(kThreads)
project_bwd_kernel(const float* __restrict__ g_uv, const float* __restrict__ points, const void* K, int k_dtype,
                   const float* Tcw, float* __restrict__ g_points, float* __restrict__ g_Tcw, Slot* ws,
                   int H, int W, int normalize) {
    __shared__ ProjCam c;
    __shared__ double red[12 * (kThreads / 32)];
    __shared__ int flag;
    const int b = blockIdx.y, P = H * W;
    const float wm1 = static_cast<float>(W - 1), hm1 = static_cast<float>(H - 1);
    if (threadIdx.x == 0) setup_proj(K, k_dtype, Tcw, b, c);
    __syncthreads();
    float gT[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) gT[i] = 0.0f;
    for (int p = blockIdx.x * kThreads + threadIdx.x; p < P; p += gridDim.x * kThreads) {
        float X[3], Y[3], gY[3], gX[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) X[k] = __ldg(points + (static_cast<size_t>(b) * 3 + k) * P + p);
        if (Tcw != nullptr) rigid(c.T, X, Y);
        else { Y[0] = X[0]; Y[1] = X[1]; Y[2] = X[2]; }
        Proj pr;
        project_cam(c.Kr, Y, wm1, hm1, normalize != 0, pr);
        const float2 g = __ldg(reinterpret_cast<const float2*>(g_uv) + static_cast<size_t>(b) * P + p);
        project_cam_adjoint(c.Kr, pr, Y, wm1, hm1, normalize != 0, g.x, g.y, gY);
        if (Tcw != nullptr) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                gT[4 * k + 0] += gY[k] * X[0];
                gT[4 * k + 1] += gY[k] * X[1];
                gT[4 * k + 2] += gY[k] * X[2];
                gT[4 * k + 3] += gY[k];
            }
            rigid_adjoint(c.T, gY, gX);
        } else {
            gX[0] = gY[0]; gX[1] = gY[1]; gX[2] = gY[2];
        }
        if (g_points != nullptr) {
#pragma unroll
            for (int k = 0; k < 3; ++k) g_points[(static_cast<size_t>(b) * 3 + k) * P + p] = gX[k];
        }
    }
    if (g_Tcw == nullptr) return;
    Slot* slot = slot_at(ws, b);
    block_accumulate12(gT, red, spread_acc(slot));
    if (last_block(slot, gridDim.x, &flag) && threadIdx.x < 32)
        finish_pose_grad_warp(slot, DROSFM_POSE_MAT4, nullptr, g_Tcw + b * 16);
}

// ------------------------------------------------------------------------------------------
// Pose.from_vec(vec, 'euler')
// ------------------------------------------------------------------------------------------
__global__ void pose_vec2mat_fwd_kernel(const float* __restrict__ vec, float* __restrict__ mat, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float v[6], T[12], trig[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) v[k] = vec[i * 6 + k];
    euler_to_mat34(v, T, trig);
#pragma unroll
    for (int k = 0; k < 12; ++k) mat[i * 16 + k] = T[k];
    mat[i * 16 + 12] = 0.0f; mat[i * 16 + 13] = 0.0f; mat[i * 16 + 14] = 0.0f; mat[i * 16 + 15] = 1.0f;
}

__global__ void pose_vec2mat_bwd_kernel(const float* __restrict__ g_mat, const float* __restrict__ vec,
                                        float* __restrict__ g_vec, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    float v[6], T[12], trig[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) v[k] = vec[i * 6 + k];
    euler_to_mat34(v, T, trig);
    double g[12], gv[6];
#pragma unroll
    for (int k = 0; k < 12; ++k) g[k] = g_mat[i * 16 + k];
    euler_adjoint(g, trig, gv);
#pragma unroll
    for (int k = 0; k < 6; ++k) g_vec[i * 6 + k] = static_cast<float>(gv[k]);
}

// drosfm_selftest_rcp: rcp_rn_normal(x) against __frcp_rn(x) for EVERY float with 2^-126 <= |x| < 2^126
__global__ void __launch_bounds__(256) selftest_rcp_kernel(unsigned long long* mismatches) {
    unsigned long long bad = 0;
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned long long i = blockIdx.x * blockDim.x + threadIdx.x; i < (1ull << 32); i += stride) {
        const unsigned bits = static_cast<unsigned>(i);
        const unsigned e = (bits >> 23) & 0xffu;
        if (e < 1u || e > 252u) continue;
        const float x = __uint_as_float(bits);
        if (__float_as_uint(rcp_rn_normal(x)) != __float_as_uint(__frcp_rn(x))) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

static int check_dims(int B, int H, int W) {
    DROSFM_REQUIRE(B >= 0 && H >= 0 && W >= 0, DROSFM_EINVAL, "negative dimension B=%d H=%d W=%d", B, H, W);
    DROSFM_REQUIRE(static_cast<long long>(H) * W < (1ll << 30) && B <= 65535, DROSFM_ERANGE,
                   "dimensions out of range: B=%d (max 65535) H*W=%lld (max 2^30)", B, static_cast<long long>(H) * W);
    return DROSFM_OK;
}

static int check_cams(const drosfm_cams_t* c) {
    DROSFM_REQUIRE(c != nullptr && c->K != nullptr && c->Kref != nullptr, DROSFM_EINVAL, "cams/K/Kref is NULL");
    DROSFM_REQUIRE(c->k_dtype == DROSFM_F32 || c->k_dtype == DROSFM_F64, DROSFM_EINVAL, "bad k_dtype %d", c->k_dtype);
    DROSFM_REQUIRE(c->pose_kind >= DROSFM_POSE_IDENTITY && c->pose_kind <= DROSFM_POSE_EULER6, DROSFM_EINVAL,
                   "bad pose_kind %d", c->pose_kind);
    return DROSFM_OK;
}

static int persistent_blocks(int work_items_per_sample, int per_block, int B) {
    int need = (work_items_per_sample + per_block - 1) / per_block;
    int cap = (kNumSMs * 8 + B - 1) / (B > 0 ? B : 1);
    if (cap < 1) cap = 1;
    if (need < 1) need = 1;
    return need < cap ? need : cap;
}

}  // namespace drosfm

using namespace drosfm;

extern "C" {

int drosfm_pose_vec2mat_fwd(const float* vec, float* mat, int N, drosfm_stream_t stream) {
    DROSFM_REQUIRE(N >= 0, DROSFM_EINVAL, "pose_vec2mat_fwd: negative N");
    if (N == 0) return DROSFM_OK;
    DROSFM_REQUIRE(vec && mat, DROSFM_EINVAL, "pose_vec2mat_fwd: NULL argument");
    pose_vec2mat_fwd_kernel<<<(N + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(vec, mat, N);
    return launch_status("pose_vec2mat_fwd");
}

int drosfm_pose_vec2mat_bwd(const float* g_mat, const float* vec, float* g_vec, int N, drosfm_stream_t stream) {
    DROSFM_REQUIRE(N >= 0, DROSFM_EINVAL, "pose_vec2mat_bwd: negative N");
    if (N == 0) return DROSFM_OK;
    DROSFM_REQUIRE(g_mat && vec && g_vec, DROSFM_EINVAL, "pose_vec2mat_bwd: NULL argument");
    pose_vec2mat_bwd_kernel<<<(N + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(g_mat, vec, g_vec, N);
    return launch_status("pose_vec2mat_bwd");
}

int drosfm_warp_coords_fwd(const float* depth, int depth_kind, const drosfm_cams_t* cams, float* uv, uint8_t* mask,
                           int B, int H, int W, int normalize, drosfm_stream_t stream) {
    if (int e = check_dims(B, H, W)) return e;
    if (B == 0 || H * W == 0) return DROSFM_OK;
    if (int e = check_cams(cams)) return e;
    DROSFM_REQUIRE(depth && uv, DROSFM_EINVAL, "warp_coords_fwd: NULL depth/uv");
    DROSFM_REQUIRE(cams->pose_kind == DROSFM_POSE_IDENTITY || cams->pose != nullptr, DROSFM_EINVAL,
                   "warp_coords_fwd: pose_kind %d needs cams->pose", cams->pose_kind);
    const int P = H * W;
    const bool vec = (P % 4 == 0) && aligned16(depth) && aligned16(uv);
    const int px = vec ? 4 : 1;
    dim3 grid((P + kThreads * px - 1) / (kThreads * px), B);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const char* env = std::getenv("DROSFM_COORDS_SHARED_RCP");
    const int chain = env == nullptr ? 0 : (env[0] == '1' ? 1 : (env[0] == '2' ? 2 : 0));
#define COORDS(VEC_, CH_) warp_coords_fwd_kernel<VEC_, CH_><<<grid, kThreads, 0, s>>>(depth, depth_kind, *cams, uv, mask, H, W, normalize)
    if (vec) { if (chain == 2) COORDS(true, 2); else if (chain == 1) COORDS(true, 1); else COORDS(true, 0); }
    else { if (chain == 2) COORDS(false, 2); else if (chain == 1) COORDS(false, 1); else COORDS(false, 0); }
#undef COORDS
    return launch_status("warp_coords_fwd");
}

int drosfm_selftest_rcp(unsigned long long* mismatches, drosfm_stream_t stream) {
    DROSFM_REQUIRE(mismatches != nullptr, DROSFM_EINVAL, "selftest_rcp: NULL counter");
    selftest_rcp_kernel<<<kNumSMs * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(mismatches);
    return launch_status("selftest_rcp");
}

int drosfm_warp_coords_bwd(const float* g_uv, const float* depth, int depth_kind, const drosfm_cams_t* cams,
                           float* g_depth, float* g_pose, void* ws, int B, int H, int W, int normalize,
                           drosfm_stream_t stream) {
    if (int e = check_dims(B, H, W)) return e;
    if (B == 0) return DROSFM_OK;
    if (int e = check_cams(cams)) return e;
    DROSFM_REQUIRE((g_uv && depth) || H * W == 0, DROSFM_EINVAL, "warp_coords_bwd: NULL g_uv/depth");
    DROSFM_REQUIRE(g_pose == nullptr || (ws != nullptr && cams->pose != nullptr), DROSFM_EINVAL,
                   "warp_coords_bwd: g_pose needs ws and cams->pose");
    const int P = H * W;
    const bool vec = (P % 4 == 0) && aligned16(depth) && aligned16(g_uv) && (g_depth == nullptr || aligned16(g_depth));
    const int px = vec ? 4 : 1;
    dim3 grid(persistent_blocks(P, kThreads * px, B), B);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (vec) warp_coords_bwd_kernel<true><<<grid, kThreads, 0, s>>>(g_uv, depth, depth_kind, *cams, g_depth, g_pose,
                                                                    static_cast<Slot*>(ws), H, W, normalize);
    else warp_coords_bwd_kernel<false><<<grid, kThreads, 0, s>>>(g_uv, depth, depth_kind, *cams, g_depth, g_pose,
                                                                 static_cast<Slot*>(ws), H, W, normalize);
    return launch_status("warp_coords_bwd");
}

int drosfm_reconstruct_fwd(const float* depth, const void* K, int k_dtype, const float* Twc, float* points,
                           int B, int H, int W, drosfm_stream_t stream) {
    if (int e = check_dims(B, H, W)) return e;
    if (B == 0 || H * W == 0) return DROSFM_OK;
    DROSFM_REQUIRE(depth && K && points, DROSFM_EINVAL, "reconstruct_fwd: NULL argument");
    dim3 grid((H * W + kThreads - 1) / kThreads, B);
    reconstruct_fwd_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(depth, K, k_dtype, Twc, points, H, W);
    return launch_status("reconstruct_fwd");
}

int drosfm_reconstruct_bwd(const float* g_points, const void* K, int k_dtype, const float* Twc, float* g_depth,
                           int B, int H, int W, drosfm_stream_t stream) {
    if (int e = check_dims(B, H, W)) return e;
    if (B == 0 || H * W == 0) return DROSFM_OK;
    DROSFM_REQUIRE(g_points && K && g_depth, DROSFM_EINVAL, "reconstruct_bwd: NULL argument");
    dim3 grid((H * W + kThreads - 1) / kThreads, B);
    reconstruct_bwd_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(g_points, K, k_dtype, Twc, g_depth, H, W);
    return launch_status("reconstruct_bwd");
}

int drosfm_project_fwd(const float* points, const void* K, int k_dtype, const float* Tcw, float* uv,
                       int B, int H, int W, int normalize, drosfm_stream_t stream) {
    if (int e = check_dims(B, H, W)) return e;
    if (B == 0 || H * W == 0) return DROSFM_OK;
    DROSFM_REQUIRE(points && K && uv, DROSFM_EINVAL, "project_fwd: NULL argument");
    dim3 grid((H * W + kThreads - 1) / kThreads, B);
    project_fwd_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(points, K, k_dtype, Tcw, uv, H, W, normalize);
    return launch_status("project_fwd");
}

int drosfm_project_bwd(const float* g_uv, const float* points, const void* K, int k_dtype, const float* Tcw,
                       float* g_points, float* g_Tcw, void* ws, int B, int H, int W, int normalize,
                       drosfm_stream_t stream) {
    if (int e = check_dims(B, H, W)) return e;
    if (B == 0) return DROSFM_OK;
    DROSFM_REQUIRE(g_uv && points && K, DROSFM_EINVAL, "project_bwd: NULL argument");
    DROSFM_REQUIRE(g_Tcw == nullptr || (ws != nullptr && Tcw != nullptr), DROSFM_EINVAL, "project_bwd: g_Tcw needs ws and Tcw");
    dim3 grid(persistent_blocks(H * W, kThreads, B), B);
    project_bwd_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(g_uv, points, K, k_dtype, Tcw, g_points,
                                                                                g_Tcw, static_cast<Slot*>(ws), H, W, normalize);
    return launch_status("project_bwd");
}

}  // extern "C"
