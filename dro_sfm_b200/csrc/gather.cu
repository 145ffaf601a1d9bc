// Kernel family 2: bilinear gather (F.grid_sample, bilinear, align_corners=True, zeros|border).
//
//   drosfm_grid_gather_*    : the grid_sample call itself (camera_utils.py:55, DepthPoseNet.py:92)
//   drosfm_view_synthesis_* : view_synthesis (camera_utils.py:23-56) with the coordinate chain of
//                             coords.cu fused in front, so points and coordinates never touch HBM.
//
// Forward is a 4-tap gather per (pixel, channel); lanes of a warp own adjacent pixels, so the taps of
// a warp fall into one or two 128-byte lines per source row and mostly hit L1/L2.  The gradient
// w.r.t. the source is a scatter: contributions of neighbouring lanes to the same source pixel are
// merged with warp shuffles first (warp-aggregated atomics), the rest goes out as red.global.add.
//
// Algorithmic bytes (C=3): view_synthesis fwd 28 B/px, bwd 32 B/px (+12 with g_src).
#include "common.cuh"

namespace drosfm {

constexpr int kThreads = 256;

__device__ __forceinline__ void pix_xy(int p, int W, int& x, int& y) {
    y = p / W;
    x = p - y * W;
}

__device__ __forceinline__ float tap(const float* __restrict__ plane, int Ws, const Taps& t, const Weights& w) {
    float acc = 0.0f;
    const float* r0 = plane + t.y0 * Ws + t.x0;
    if (t.valid & 1u) acc += __ldg(r0) * w.nw;
    if (t.valid & 2u) acc += __ldg(r0 + 1) * w.ne;
    if (t.valid & 4u) acc += __ldg(r0 + Ws) * w.sw;
    if (t.valid & 8u) acc += __ldg(r0 + Ws + 1) * w.se;
    return acc;
}

// Values of the four taps (0 where out of bounds): needed by the coordinate gradient.
__device__ __forceinline__ void tap_values(const float* __restrict__ plane, int Ws, const Taps& t, float* v) {
    const float* r0 = plane + t.y0 * Ws + t.x0;
    v[0] = (t.valid & 1u) ? __ldg(r0) : 0.0f;
    v[1] = (t.valid & 2u) ? __ldg(r0 + 1) : 0.0f;
    v[2] = (t.valid & 4u) ? __ldg(r0 + Ws) : 0.0f;
    v[3] = (t.valid & 8u) ? __ldg(r0 + Ws + 1) : 0.0f;
}

// d(sample)/d(ix), d(sample)/d(iy) for tap values v
__device__ __forceinline__ void tap_coord_grad(const Taps& t, const float* v, float& dx, float& dy) {
    const float bx = 1.0f - t.ax, by = 1.0f - t.ay;
    dx = (v[1] - v[0]) * by + (v[3] - v[2]) * t.ay;
    dy = (v[2] - v[0]) * bx + (v[3] - v[1]) * t.ax;
}

// Merge plan for the scatter of one warp: lane L hands its east column to lane L+1 when that lane's
// west column is the same source pixel column in the same source row pair.
struct MergePlan {
    bool give;   // my east taps are handed to lane+1
    bool take;   // I receive lane-1's east taps into my west taps
};

__device__ __forceinline__ MergePlan plan_merge(const Taps& t, bool active) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int key_x = active && t.valid ? t.x0 : -0x40000000;
    const int key_y = active && t.valid ? t.y0 : -0x40000000;
    const int nx = __shfl_down_sync(full, key_x, 1), ny = __shfl_down_sync(full, key_y, 1);
    MergePlan m;
    m.give = active && t.valid != 0u && lane < 31 && ny == key_y && nx == key_x + 1;
    const int g = __shfl_up_sync(full, m.give ? 1 : 0, 1);
    m.take = lane > 0 && g != 0;
    return m;
}

// Scatter g * weights of one channel into the source-gradient plane with the merge plan applied.
// All 32 lanes of the warp must call this (inactive lanes pass g = 0 and a plan with give=take=false).
__device__ __forceinline__ void scatter_taps(float* __restrict__ plane, int Ws, const Taps& t, const Weights& w,
                                             const MergePlan& m, float g) {
    const unsigned full = 0xffffffffu;
    float cnw = g * w.nw, cne = g * w.ne, csw = g * w.sw, cse = g * w.se;
    // the east column of lane-1 lands on my west column; validity of the shared taps is identical
    const float rn = __shfl_up_sync(full, cne, 1), rs = __shfl_up_sync(full, cse, 1);
    if (m.take) { cnw += rn; csw += rs; }
    float* r0 = plane + t.y0 * Ws + t.x0;
    if (t.valid & 1u) atomicAdd(r0, cnw);
    if (t.valid & 4u) atomicAdd(r0 + Ws, csw);
    if (!m.give) {
        if (t.valid & 2u) atomicAdd(r0 + 1, cne);
        if (t.valid & 8u) atomicAdd(r0 + Ws + 1, cse);
    }
}

// ------------------------------------------------------------------------------------------
// plain grid gather
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
grid_gather_fwd_kernel(const float* __restrict__ src, const float* __restrict__ uv, float* __restrict__ out,
                       int C, int Hs, int Ws, int H, int W, int padding) {
    const int b = blockIdx.y, P = H * W;
    const int p = blockIdx.x * kThreads + threadIdx.x;
    if (p >= P) return;
    const float2 c = __ldg(reinterpret_cast<const float2*>(uv) + static_cast<size_t>(b) * P + p);
    Taps t;
    make_taps(c.x, c.y, Hs, Ws, padding, t);
    const Weights w = tap_weights(t);
    const size_t sp = static_cast<size_t>(Hs) * Ws;
    const float* s = src + static_cast<size_t>(b) * C * sp;
    float* o = out + static_cast<size_t>(b) * C * P + p;
#pragma unroll 4
    for (int ch = 0; ch < C; ++ch) o[static_cast<size_t>(ch) * P] = t.valid ? tap(s + ch * sp, Ws, t, w) : 0.0f;
}

__global__ void __launch_bounds__(kThreads)
grid_gather_bwd_kernel(const float* __restrict__ g_out, const float* __restrict__ src, const float* __restrict__ uv,
                       float* __restrict__ g_src, float* __restrict__ g_uv, int C, int Hs, int Ws, int H, int W,
                       int padding) {
    const int b = blockIdx.y, P = H * W;
    const int p = blockIdx.x * kThreads + threadIdx.x;
    const bool active = p < P;
    Taps t;
    t.valid = 0u; t.x0 = t.y0 = 0; t.ax = t.ay = 0.0f; t.mx = t.my = 0.0f;
    if (active) {
        const float2 c = __ldg(reinterpret_cast<const float2*>(uv) + static_cast<size_t>(b) * P + p);
        make_taps(c.x, c.y, Hs, Ws, padding, t);
    }
    const Weights w = tap_weights(t);
    const MergePlan m = plan_merge(t, active);
    const size_t sp = static_cast<size_t>(Hs) * Ws;
    const float* s = src + static_cast<size_t>(b) * C * sp;
    float gx = 0.0f, gy = 0.0f;
    for (int ch = 0; ch < C; ++ch) {
        const float g = active ? __ldg(g_out + (static_cast<size_t>(b) * C + ch) * P + p) : 0.0f;
        if (g_uv != nullptr && t.valid) {
            float v[4], dx, dy;
            tap_values(s + ch * sp, Ws, t, v);
            tap_coord_grad(t, v, dx, dy);
            gx += g * dx;
            gy += g * dy;
        }
        if (g_src != nullptr) scatter_taps(g_src + (static_cast<size_t>(b) * C + ch) * sp, Ws, t, w, m, g);
    }
    if (active && g_uv != nullptr)
        reinterpret_cast<float2*>(g_uv)[static_cast<size_t>(b) * P + p] = make_float2(gx * t.mx, gy * t.my);
}

// ------------------------------------------------------------------------------------------
// fused view synthesis
// ------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(kThreads)
view_synthesis_fwd_kernel(const float* __restrict__ src, const float* __restrict__ depth, int depth_kind,
                          drosfm_cams_t cams, float* __restrict__ out, int C, int Hs, int Ws, int H, int W,
                          int padding) {
    __shared__ Cam cam;
    const int b = blockIdx.y, P = H * W;
    const Norm nm = make_norm(W, H);
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
    Taps t[PX];
    Weights w[PX];
#pragma unroll
    for (int k = 0; k < PX; ++k) {
        int x, y;
        pix_xy(p0 + k, W, x, y);
        Warp wp;
        warp_pixel_fast(cam, x, y, to_depth_fast(d[k], depth_kind), nm, true, wp);
        make_taps(wp.p.u, wp.p.v, Hs, Ws, padding, t[k]);
        w[k] = tap_weights(t[k]);
    }
    const size_t sp = static_cast<size_t>(Hs) * Ws;
    const float* s = src + static_cast<size_t>(b) * C * sp;
    for (int ch = 0; ch < C; ++ch) {
        float o[PX];
#pragma unroll
        for (int k = 0; k < PX; ++k) o[k] = t[k].valid ? tap(s + ch * sp, Ws, t[k], w[k]) : 0.0f;
        float* dst = out + (static_cast<size_t>(b) * C + ch) * P + p0;
        if constexpr (VEC) st4_streaming(dst, make_float4(o[0], o[1], o[2], o[3]));
        else dst[0] = o[0];
    }
}

__global__ void __launch_bounds__(kThreads)
view_synthesis_bwd_kernel(const float* __restrict__ g_out, const float* __restrict__ src,
                          const float* __restrict__ depth, int depth_kind, drosfm_cams_t cams,
                          float* __restrict__ g_src, float* __restrict__ g_depth, float* __restrict__ g_pose, Slot* ws,
                          int C, int Hs, int Ws, int H, int W, int padding) {
    __shared__ Cam cam;
    __shared__ double red[12 * (kThreads / 32)];
    __shared__ int flag;
    const int b = blockIdx.y, P = H * W;
    const Norm nm = make_norm(W, H);
    const float wm1 = nm.wm1, hm1 = nm.hm1;
    if (threadIdx.x == 0) setup_cam(cams, cams.pose, b, cam);
    __syncthreads();
    float gT[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) gT[i] = 0.0f;
    const size_t sp = static_cast<size_t>(Hs) * Ws;
    const float* s = src + static_cast<size_t>(b) * C * sp;
    const int stride = gridDim.x * kThreads;
    // every warp runs the same trip count so that the shuffles in scatter_taps stay convergent
    const int trips = (P + stride - 1) / stride;
    for (int it = 0; it < trips; ++it) {
        const int p = it * stride + blockIdx.x * kThreads + threadIdx.x;
        const bool active = p < P;
        Taps t;
        t.valid = 0u; t.x0 = t.y0 = 0; t.ax = t.ay = 0.0f; t.mx = t.my = 0.0f;
        Warp wp;
        float dv = 0.0f;
        if (active) {
            int x, y;
            pix_xy(p, W, x, y);
            dv = __ldg(depth + static_cast<size_t>(b) * P + p);
            warp_pixel_fast(cam, x, y, to_depth_fast(dv, depth_kind), nm, true, wp);
            make_taps(wp.p.u, wp.p.v, Hs, Ws, padding, t);
        }
        const Weights w = tap_weights(t);
        MergePlan m;
        m.give = m.take = false;
        if (g_src != nullptr) m = plan_merge(t, active);
        float gx = 0.0f, gy = 0.0f;
        for (int ch = 0; ch < C; ++ch) {
            const float g = active ? __ldg(g_out + (static_cast<size_t>(b) * C + ch) * P + p) : 0.0f;
            if (t.valid) {
                float v[4], dx, dy;
                tap_values(s + ch * sp, Ws, t, v);
                tap_coord_grad(t, v, dx, dy);
                gx += g * dx;
                gy += g * dy;
            }
            if (g_src != nullptr) scatter_taps(g_src + (static_cast<size_t>(b) * C + ch) * sp, Ws, t, w, m, g);
        }
        if (active) {
            const float gd = warp_pixel_adjoint(cam, wp, to_depth_fast(dv, depth_kind), wm1, hm1, true, gx * t.mx, gy * t.my, gT);
            if (g_depth != nullptr)
                g_depth[static_cast<size_t>(b) * P + p] = depth_kind == DROSFM_INV_DEPTH ? inv2depth_grad(dv, gd) : gd;
        }
    }
    if (g_pose == nullptr) return;
    Slot* slot = slot_at(ws, b);
    block_accumulate12(gT, red, spread_acc(slot));
    if (last_block(slot, gridDim.x, &flag) && threadIdx.x < 32) {
        const int st = cams.pose_kind == DROSFM_POSE_EULER6 ? 6 : 16;
        finish_pose_grad_warp(slot, cams.pose_kind, cams.pose_kind == DROSFM_POSE_EULER6 ? cams.pose + b * 6 : nullptr,
                         g_pose + b * st);
    }
}

static int check_gather(int B, int C, int Hs, int Ws, int H, int W, int padding) {
    DROSFM_REQUIRE(B >= 0 && C >= 0 && Hs >= 0 && Ws >= 0 && H >= 0 && W >= 0, DROSFM_EINVAL, "negative dimension");
    DROSFM_REQUIRE(B <= 65535 && static_cast<long long>(H) * W < (1ll << 30) &&
                   static_cast<long long>(Hs) * Ws < (1ll << 30), DROSFM_ERANGE, "dimension out of range");
    DROSFM_REQUIRE(padding == DROSFM_PAD_ZEROS || padding == DROSFM_PAD_BORDER, DROSFM_EINVAL,
                   "unsupported padding mode %d", padding);
    return DROSFM_OK;
}

}  // namespace drosfm

using namespace drosfm;

extern "C" {

int drosfm_grid_gather_fwd(const float* src, const float* uv, float* out, int B, int C, int Hs, int Ws,
                           int H, int W, int padding, drosfm_stream_t stream) {
    if (int e = check_gather(B, C, Hs, Ws, H, W, padding)) return e;
    if (B == 0 || C == 0 || H * W == 0) return DROSFM_OK;
    DROSFM_REQUIRE(src && uv && out, DROSFM_EINVAL, "grid_gather_fwd: NULL argument");
    DROSFM_REQUIRE(Hs > 0 && Ws > 0, DROSFM_EINVAL, "grid_gather_fwd: empty source");
    dim3 grid((H * W + kThreads - 1) / kThreads, B);
    grid_gather_fwd_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(src, uv, out, C, Hs, Ws, H, W, padding);
    return launch_status("grid_gather_fwd");
}

int drosfm_grid_gather_bwd(const float* g_out, const float* src, const float* uv, float* g_src, float* g_uv,
                           int B, int C, int Hs, int Ws, int H, int W, int padding, drosfm_stream_t stream) {
    if (int e = check_gather(B, C, Hs, Ws, H, W, padding)) return e;
    if (B == 0 || H * W == 0) return DROSFM_OK;
    DROSFM_REQUIRE(g_out && src && uv, DROSFM_EINVAL, "grid_gather_bwd: NULL argument");
    DROSFM_REQUIRE(Hs > 0 && Ws > 0, DROSFM_EINVAL, "grid_gather_bwd: empty source");
    dim3 grid((H * W + kThreads - 1) / kThreads, B);
    grid_gather_bwd_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(g_out, src, uv, g_src, g_uv, C, Hs, Ws,
                                                                                    H, W, padding);
    return launch_status("grid_gather_bwd");
}

int drosfm_view_synthesis_fwd(const float* src, const float* depth, int depth_kind, const drosfm_cams_t* cams,
                              float* out, int B, int C, int Hs, int Ws, int H, int W, int padding,
                              drosfm_stream_t stream) {
    if (int e = check_gather(B, C, Hs, Ws, H, W, padding)) return e;
    if (B == 0 || C == 0 || H * W == 0) return DROSFM_OK;
    DROSFM_REQUIRE(cams && cams->K && cams->Kref, DROSFM_EINVAL, "view_synthesis_fwd: NULL cams");
    DROSFM_REQUIRE(src && depth && out, DROSFM_EINVAL, "view_synthesis_fwd: NULL argument");
    DROSFM_REQUIRE(Hs > 0 && Ws > 0, DROSFM_EINVAL, "view_synthesis_fwd: empty source");
    const int P = H * W;
    const bool vec = (P % 4 == 0) && aligned16(depth) && aligned16(out);
    const int px = vec ? 4 : 1;
    dim3 grid((P + kThreads * px - 1) / (kThreads * px), B);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (vec) view_synthesis_fwd_kernel<true><<<grid, kThreads, 0, s>>>(src, depth, depth_kind, *cams, out, C, Hs, Ws, H, W, padding);
    else view_synthesis_fwd_kernel<false><<<grid, kThreads, 0, s>>>(src, depth, depth_kind, *cams, out, C, Hs, Ws, H, W, padding);
    return launch_status("view_synthesis_fwd");
}

int drosfm_view_synthesis_bwd(const float* g_out, const float* src, const float* depth, int depth_kind,
                              const drosfm_cams_t* cams, float* g_src, float* g_depth, float* g_pose, void* ws,
                              int B, int C, int Hs, int Ws, int H, int W, int padding, drosfm_stream_t stream) {
    if (int e = check_gather(B, C, Hs, Ws, H, W, padding)) return e;
    if (B == 0) return DROSFM_OK;
    DROSFM_REQUIRE(cams && cams->K && cams->Kref, DROSFM_EINVAL, "view_synthesis_bwd: NULL cams");
    DROSFM_REQUIRE(g_out && src && depth, DROSFM_EINVAL, "view_synthesis_bwd: NULL argument");
    DROSFM_REQUIRE(g_pose == nullptr || (ws != nullptr && cams->pose != nullptr), DROSFM_EINVAL,
                   "view_synthesis_bwd: g_pose needs ws and cams->pose");
    DROSFM_REQUIRE(Hs > 0 && Ws > 0, DROSFM_EINVAL, "view_synthesis_bwd: empty source");
    const int P = H * W;
    int blocks = (P + kThreads - 1) / kThreads;
    const int cap = (kNumSMs * 8 + B - 1) / B;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    dim3 grid(blocks, B);
    view_synthesis_bwd_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        g_out, src, depth, depth_kind, *cams, g_src, g_depth, g_pose, static_cast<Slot*>(ws), C, Hs, Ws, H, W, padding);
    return launch_status("view_synthesis_bwd");
}

}  // extern "C"
