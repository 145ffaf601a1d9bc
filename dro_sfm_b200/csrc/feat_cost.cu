// Kernel family 3: fused feature-metric cost of the recurrent optimiser.
//
//   DepthPoseNet.get_cost_each   (dro_sfm/networks/depth_pose/DepthPoseNet.py:76-96)   V = 1
//   DepthPoseNet.depth_cost_calc (dro_sfm/networks/depth_pose/DepthPoseNet.py:98-105)  V views, mean
//
//   cost[b,c,y,x] = (1/V) * sum_v (fmap[b,c,y,x] - bilinear(fmap_ref_v[b,c], uv_v(y,x)))^2
//
// with uv_v from the fused back-project/transform/project chain (common.cuh).  Neither the warped
// features, the differences, the per-view costs nor their stack are materialised: per call the
// kernel reads fmap, the V source maps and depth once and writes the cost map once
// ((V+2)*4*C + 4 bytes per feature pixel; backward (2V+3)*4*C + 8).
//
// Two storage layouts of the same logical [B,C,h,w] tensors:
//   NCHW  lane = pixel, loop over a channel group; 4-byte accesses, 128 B per warp instruction.
//   NHWC  (torch channels_last) warp = 32 pixels, lane = 4 channels; every tap, load and store is a
//         128-bit access of a 512-byte contiguous segment, source gradients are red.global.add.v4.f32.
// The backward pass scatters into the source gradients with atomics; neighbouring lanes that hit the
// same source pixel are merged with shuffles first (NCHW), pose gradients are reduced in fp64.
#include "common.cuh"

namespace drosfm {

struct ViewPtrs {
    const float* ref[DROSFM_MAX_VIEWS];
    const float* pose[DROSFM_MAX_VIEWS];
};
struct ViewGrads {
    float* g_ref[DROSFM_MAX_VIEWS];
    float* g_pose[DROSFM_MAX_VIEWS];
};

__device__ __forceinline__ void pix_xy(int p, int W, int& x, int& y) {
    y = p / W;
    x = p - y * W;
}

__device__ __forceinline__ void tap_values(const float* __restrict__ plane, int Ws, const Taps& t, float* v) {
    const float* r0 = plane + t.y0 * Ws + t.x0;
    v[0] = (t.valid & 1u) ? __ldg(r0) : 0.0f;
    v[1] = (t.valid & 2u) ? __ldg(r0 + 1) : 0.0f;
    v[2] = (t.valid & 4u) ? __ldg(r0 + Ws) : 0.0f;
    v[3] = (t.valid & 8u) ? __ldg(r0 + Ws + 1) : 0.0f;
}

__device__ __forceinline__ float blend(const float* v, const Weights& w) {
    return v[0] * w.nw + v[1] * w.ne + v[2] * w.sw + v[3] * w.se;
}

// ------------------------------------------------------------------------------------------
// NCHW forward
// ------------------------------------------------------------------------------------------
constexpr int kPixThreads = 128;

template <int VT>
__global__ void __launch_bounds__(kPixThreads)
feat_cost_fwd_nchw(const float* __restrict__ fmap, ViewPtrs vp, const float* __restrict__ depth, int depth_kind,
                   drosfm_cams_t cams, int V, float* __restrict__ cost, int C, int h, int w, int cg) {
    __shared__ Cam cam[VT];
    const int b = blockIdx.z, P = h * w;
    const int p = blockIdx.x * kPixThreads + threadIdx.x;
    const bool active = p < P;
    const float d = active ? to_depth(__ldg(depth + static_cast<size_t>(b) * P + p), depth_kind) : 0.0f;
#pragma unroll
    for (int v = 0; v < VT; ++v)
        if (v < V && threadIdx.x == v) setup_cam(cams, vp.pose[v], b, cam[v]);
    __syncthreads();
    if (!active) return;
    const float wm1 = static_cast<float>(w - 1), hm1 = static_cast<float>(h - 1);
    int x, y;
    pix_xy(p, w, x, y);
    Taps t[VT];
    Weights wt[VT];
#pragma unroll
    for (int v = 0; v < VT; ++v) {
        if (v < V) {
            Warp wp;
            warp_pixel(cam[v], x, y, d, wm1, hm1, true, wp);
            make_taps(wp.p.u, wp.p.v, h, w, DROSFM_PAD_ZEROS, t[v]);
            wt[v] = tap_weights(t[v]);
        }
    }
    const int c0 = blockIdx.y * cg, c1 = min(C, c0 + cg);
    const float fV = static_cast<float>(V);
    const size_t sample = static_cast<size_t>(b) * C * P;
#pragma unroll 4
    for (int c = c0; c < c1; ++c) {
        const size_t plane = sample + static_cast<size_t>(c) * P;
        const float f = __ldg(fmap + plane + p);
        float acc = 0.0f;
#pragma unroll
        for (int v = 0; v < VT; ++v) {
            if (v < V) {
                float tv[4];
                tap_values(vp.ref[v] + plane, w, t[v], tv);
                const float df = f - blend(tv, wt[v]);
                acc += df * df;
            }
        }
        cost[plane + p] = V == 1 ? acc : acc / fV;
    }
}

// ------------------------------------------------------------------------------------------
// NCHW backward
// ------------------------------------------------------------------------------------------
struct MergePlan {
    bool give, take;
};

__device__ __forceinline__ MergePlan plan_merge(const Taps& t, bool active) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const bool on = active && t.valid != 0u;
    const int key_x = on ? t.x0 : -0x40000000, key_y = on ? t.y0 : -0x40000000;
    const int nx = __shfl_down_sync(full, key_x, 1), ny = __shfl_down_sync(full, key_y, 1);
    MergePlan m;
    m.give = on && lane < 31 && ny == key_y && nx == key_x + 1;
    m.take = (__shfl_up_sync(full, m.give ? 1 : 0, 1) != 0) && lane > 0;
    return m;
}

__device__ __forceinline__ void scatter_taps(float* __restrict__ plane, int Ws, const Taps& t, const Weights& w,
                                             const MergePlan& m, float g) {
    const unsigned full = 0xffffffffu;
    float cnw = g * w.nw, cne = g * w.ne, csw = g * w.sw, cse = g * w.se;
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

template <int VT>
__global__ void __launch_bounds__(kPixThreads)
feat_cost_bwd_nchw(const float* __restrict__ g_cost, const float* __restrict__ fmap, ViewPtrs vp,
                   const float* __restrict__ depth, int depth_kind, drosfm_cams_t cams, int V,
                   float* __restrict__ g_fmap, ViewGrads vg, float* __restrict__ g_depth, Slot* ws,
                   int B, int C, int h, int w, int cg, int need_coord_grad) {
    __shared__ Cam cam[VT];
    __shared__ double red[12 * (kPixThreads / 32)];
    __shared__ int flag;
    const int b = blockIdx.z, P = h * w;
    const int p = blockIdx.x * kPixThreads + threadIdx.x;
    const bool active = p < P;
    const float draw = active ? __ldg(depth + static_cast<size_t>(b) * P + p) : 0.0f;
    const float d = to_depth(draw, depth_kind);
#pragma unroll
    for (int v = 0; v < VT; ++v)
        if (v < V && threadIdx.x == v) setup_cam(cams, vp.pose[v], b, cam[v]);
    __syncthreads();
    const float wm1 = static_cast<float>(w - 1), hm1 = static_cast<float>(h - 1);
    int x = 0, y = 0;
    if (active) pix_xy(p, w, x, y);
    Taps t[VT];
    Weights wt[VT];
    MergePlan mp[VT];
    float gx[VT], gy[VT];
#pragma unroll
    for (int v = 0; v < VT; ++v) {
        t[v].valid = 0u; t[v].x0 = t[v].y0 = 0; t[v].ax = t[v].ay = 0.0f; t[v].mx = t[v].my = 0.0f;
        gx[v] = gy[v] = 0.0f;
        mp[v].give = mp[v].take = false;
        if (v < V) {
            if (active) {
                Warp wp;
                warp_pixel(cam[v], x, y, d, wm1, hm1, true, wp);
                make_taps(wp.p.u, wp.p.v, h, w, DROSFM_PAD_ZEROS, t[v]);
            }
            if (vg.g_ref[v] != nullptr) mp[v] = plan_merge(t[v], active);
        }
        wt[v] = tap_weights(t[v]);
    }
    const int c0 = blockIdx.y * cg, c1 = min(C, c0 + cg);
    const float scale = 2.0f / static_cast<float>(V);
    const size_t sample = static_cast<size_t>(b) * C * P;
    for (int c = c0; c < c1; ++c) {
        const size_t plane = sample + static_cast<size_t>(c) * P;
        const float g = active ? __ldg(g_cost + plane + p) * scale : 0.0f;
        const float f = active ? __ldg(fmap + plane + p) : 0.0f;
        float gf = 0.0f;
#pragma unroll
        for (int v = 0; v < VT; ++v) {
            if (v < V) {
                float tv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                if (t[v].valid) tap_values(vp.ref[v] + plane, w, t[v], tv);
                const float coef = (f - blend(tv, wt[v])) * g;   // d cost / d fmap  (= -d cost / d warped)
                gf += coef;
                if (need_coord_grad) {
                    const float bx = 1.0f - t[v].ax, by = 1.0f - t[v].ay;
                    gx[v] -= coef * ((tv[1] - tv[0]) * by + (tv[3] - tv[2]) * t[v].ay);
                    gy[v] -= coef * ((tv[2] - tv[0]) * bx + (tv[3] - tv[1]) * t[v].ax);
                }
                if (vg.g_ref[v] != nullptr) scatter_taps(vg.g_ref[v] + plane, w, t[v], wt[v], mp[v], -coef);
            }
        }
        if (active && g_fmap != nullptr) g_fmap[plane + p] = gf;
    }
    if (!need_coord_grad) return;
    float gd = 0.0f;
#pragma unroll
    for (int v = 0; v < VT; ++v) {
        if (v < V) {
            float gT[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) gT[i] = 0.0f;
            if (active) {
                Warp wp;
                warp_pixel(cam[v], x, y, d, wm1, hm1, true, wp);
                gd += warp_pixel_adjoint(cam[v], wp, d, wm1, hm1, true, gx[v] * t[v].mx, gy[v] * t[v].my, gT);
            }
            if (vg.g_pose[v] != nullptr) {
                Slot* slot = ws + (v * B + b);
                block_accumulate<12>(gT, red, slot->acc);
                if (last_block(slot, gridDim.x * gridDim.y, &flag) && threadIdx.x == 0) {
                    const bool eul = cams.pose_kind == DROSFM_POSE_EULER6;
                    finish_pose_grad(slot, cams.pose_kind, eul ? vp.pose[v] + b * 6 : nullptr,
                                     vg.g_pose[v] + b * (eul ? 6 : 16));
                }
            }
        }
    }
    if (active && g_depth != nullptr) {
        if (depth_kind == DROSFM_INV_DEPTH) gd = inv2depth_grad(draw, gd);
        if (gridDim.y == 1) g_depth[static_cast<size_t>(b) * P + p] = gd;
        else atomicAdd(g_depth + static_cast<size_t>(b) * P + p, gd);
    }
}

// ------------------------------------------------------------------------------------------
// NHWC (channels_last) forward: warp = 32 consecutive pixels, lane = 4 channels of a 128-channel slab
// ------------------------------------------------------------------------------------------
constexpr int kWarpsPerBlock = 4;

struct TapB {   // taps of one pixel, broadcast from the lane that computed them
    int off;        // (y0 * w + x0) element offset of the north-west tap (may be negative)
    float ax, ay;
    unsigned valid;
};

__device__ __forceinline__ TapB bcast(const Taps& t, int w, int src) {
    const unsigned full = 0xffffffffu;
    TapB o;
    o.off = __shfl_sync(full, t.y0 * w + t.x0, src);
    o.ax = __shfl_sync(full, t.ax, src);
    o.ay = __shfl_sync(full, t.ay, src);
    o.valid = __shfl_sync(full, t.valid, src);
    return o;
}

__device__ __forceinline__ float4 ld4z(const float* p, bool ok) {
    return ok ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
}

__device__ __forceinline__ float4 blend4(const float4& a, const float4& b, const float4& c, const float4& d,
                                         float wa, float wb, float wc, float wd) {
    return make_float4(a.x * wa + b.x * wb + c.x * wc + d.x * wd, a.y * wa + b.y * wb + c.y * wc + d.y * wd,
                       a.z * wa + b.z * wb + c.z * wc + d.z * wd, a.w * wa + b.w * wb + c.w * wc + d.w * wd);
}

template <int VT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
feat_cost_fwd_nhwc(const float* __restrict__ fmap, ViewPtrs vp, const float* __restrict__ depth, int depth_kind,
                   drosfm_cams_t cams, int V, float* __restrict__ cost, int C, int h, int w) {
    __shared__ Cam cam[VT];
    const int b = blockIdx.y, P = h * w;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int pbase = (blockIdx.x * kWarpsPerBlock + wid) * 32;
    const int p = pbase + lane;
    const bool active = p < P;
    const float d = active ? to_depth(__ldg(depth + static_cast<size_t>(b) * P + p), depth_kind) : 0.0f;
#pragma unroll
    for (int v = 0; v < VT; ++v)
        if (v < V && threadIdx.x == v) setup_cam(cams, vp.pose[v], b, cam[v]);
    __syncthreads();
    if (pbase >= P) return;
    const float wm1 = static_cast<float>(w - 1), hm1 = static_cast<float>(h - 1);
    int x = 0, y = 0;
    if (active) pix_xy(p, w, x, y);
    Taps t[VT];
#pragma unroll
    for (int v = 0; v < VT; ++v) {
        t[v].valid = 0u; t[v].x0 = t[v].y0 = 0; t[v].ax = t[v].ay = 0.0f;
        if (v < V && active) {
            Warp wp;
            warp_pixel(cam[v], x, y, d, wm1, hm1, true, wp);
            make_taps(wp.p.u, wp.p.v, h, w, DROSFM_PAD_ZEROS, t[v]);
        }
    }
    const int npix = min(32, P - pbase);
    const float fV = static_cast<float>(V);
    const size_t sample = static_cast<size_t>(b) * P * C;
    for (int cb0 = 0; cb0 < C; cb0 += 128) {
        const int cb = cb0 + lane * 4;
        const bool chan_ok = cb < C;   // lanes beyond C idle but still take part in the shuffles
#pragma unroll 2
        for (int j = 0; j < npix; ++j) {
            const size_t px = sample + static_cast<size_t>(pbase + j) * C + cb;
            const float4 f = ld4z(fmap + px, chan_ok);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int v = 0; v < VT; ++v) {
                if (v < V) {
                    const TapB tb = bcast(t[v], w, j);
                    const float* r0 = vp.ref[v] + sample + static_cast<ptrdiff_t>(tb.off) * C + cb;
                    const float4 a = ld4z(r0, chan_ok && (tb.valid & 1u)), bq = ld4z(r0 + C, chan_ok && (tb.valid & 2u));
                    const float4 c = ld4z(r0 + static_cast<size_t>(w) * C, chan_ok && (tb.valid & 4u));
                    const float4 e = ld4z(r0 + static_cast<size_t>(w) * C + C, chan_ok && (tb.valid & 8u));
                    const float bx = 1.0f - tb.ax, by = 1.0f - tb.ay;
                    const float4 wv = blend4(a, bq, c, e, bx * by, tb.ax * by, bx * tb.ay, tb.ax * tb.ay);
                    const float dx = f.x - wv.x, dy = f.y - wv.y, dz = f.z - wv.z, dw = f.w - wv.w;
                    acc.x += dx * dx; acc.y += dy * dy; acc.z += dz * dz; acc.w += dw * dw;
                }
            }
            if (V != 1) { acc.x /= fV; acc.y /= fV; acc.z /= fV; acc.w /= fV; }
            if (chan_ok) *reinterpret_cast<float4*>(cost + px) = acc;
        }
    }
}

// red.global.add.v4.f32 (sm_90+): one 16-byte reduction per lane, 512 contiguous bytes per warp
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int VT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
feat_cost_bwd_nhwc(const float* __restrict__ g_cost, const float* __restrict__ fmap, ViewPtrs vp,
                   const float* __restrict__ depth, int depth_kind, drosfm_cams_t cams, int V,
                   float* __restrict__ g_fmap, ViewGrads vg, float* __restrict__ g_depth, Slot* ws,
                   int B, int C, int h, int w, int need_coord_grad) {
    __shared__ Cam cam[VT];
    __shared__ double red[12 * kWarpsPerBlock];
    __shared__ int flag;
    const int b = blockIdx.y, P = h * w;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int pbase = (blockIdx.x * kWarpsPerBlock + wid) * 32;
    const int p = pbase + lane;
    const bool active = p < P;
    const float draw = active ? __ldg(depth + static_cast<size_t>(b) * P + p) : 0.0f;
    const float d = to_depth(draw, depth_kind);
#pragma unroll
    for (int v = 0; v < VT; ++v)
        if (v < V && threadIdx.x == v) setup_cam(cams, vp.pose[v], b, cam[v]);
    __syncthreads();
    const float wm1 = static_cast<float>(w - 1), hm1 = static_cast<float>(h - 1);
    int x = 0, y = 0;
    if (active) pix_xy(p, w, x, y);
    Taps t[VT];
    float gx[VT], gy[VT];
#pragma unroll
    for (int v = 0; v < VT; ++v) {
        t[v].valid = 0u; t[v].x0 = t[v].y0 = 0; t[v].ax = t[v].ay = 0.0f; t[v].mx = t[v].my = 0.0f;
        gx[v] = gy[v] = 0.0f;
        if (v < V && active) {
            Warp wp;
            warp_pixel(cam[v], x, y, d, wm1, hm1, true, wp);
            make_taps(wp.p.u, wp.p.v, h, w, DROSFM_PAD_ZEROS, t[v]);
        }
    }
    const int npix = max(0, min(32, P - pbase));
    const float scale = 2.0f / static_cast<float>(V);
    const size_t sample = static_cast<size_t>(b) * P * C;
    for (int cb0 = 0; cb0 < C; cb0 += 128) {
        const int cb = cb0 + lane * 4;
        const bool chan_ok = cb < C;   // lanes beyond C idle but still take part in the shuffles
        for (int j = 0; j < npix; ++j) {
            const size_t px = sample + static_cast<size_t>(pbase + j) * C + cb;
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f), f = g;
            if (chan_ok) {
                g = __ldg(reinterpret_cast<const float4*>(g_cost + px));
                f = __ldg(reinterpret_cast<const float4*>(fmap + px));
            }
            g.x *= scale; g.y *= scale; g.z *= scale; g.w *= scale;
            float4 gf = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int v = 0; v < VT; ++v) {
                if (v < V) {
                    const TapB tb = bcast(t[v], w, j);
                    const ptrdiff_t o = static_cast<ptrdiff_t>(tb.off) * C + cb;
                    const float* r0 = vp.ref[v] + sample + o;
                    const bool ok = chan_ok;
                    const float4 a = ld4z(r0, ok && (tb.valid & 1u)), bq = ld4z(r0 + C, ok && (tb.valid & 2u));
                    const float4 c = ld4z(r0 + static_cast<size_t>(w) * C, ok && (tb.valid & 4u));
                    const float4 e = ld4z(r0 + static_cast<size_t>(w) * C + C, ok && (tb.valid & 8u));
                    const float bx = 1.0f - tb.ax, by = 1.0f - tb.ay;
                    const float wnw = bx * by, wne = tb.ax * by, wsw = bx * tb.ay, wse = tb.ax * tb.ay;
                    const float4 wv = blend4(a, bq, c, e, wnw, wne, wsw, wse);
                    const float4 co = make_float4((f.x - wv.x) * g.x, (f.y - wv.y) * g.y, (f.z - wv.z) * g.z,
                                                  (f.w - wv.w) * g.w);
                    gf.x += co.x; gf.y += co.y; gf.z += co.z; gf.w += co.w;
                    if (vg.g_ref[v] != nullptr && ok) {
                        float* q0 = vg.g_ref[v] + sample + o;
                        if (tb.valid & 1u) red_add4(q0, -co.x * wnw, -co.y * wnw, -co.z * wnw, -co.w * wnw);
                        if (tb.valid & 2u) red_add4(q0 + C, -co.x * wne, -co.y * wne, -co.z * wne, -co.w * wne);
                        if (tb.valid & 4u) red_add4(q0 + static_cast<size_t>(w) * C, -co.x * wsw, -co.y * wsw, -co.z * wsw, -co.w * wsw);
                        if (tb.valid & 8u) red_add4(q0 + static_cast<size_t>(w) * C + C, -co.x * wse, -co.y * wse, -co.z * wse, -co.w * wse);
                    }
                    if (need_coord_grad) {
                        float sx = co.x * ((bq.x - a.x) * by + (e.x - c.x) * tb.ay) + co.y * ((bq.y - a.y) * by + (e.y - c.y) * tb.ay)
                                 + co.z * ((bq.z - a.z) * by + (e.z - c.z) * tb.ay) + co.w * ((bq.w - a.w) * by + (e.w - c.w) * tb.ay);
                        float sy = co.x * ((c.x - a.x) * bx + (e.x - bq.x) * tb.ax) + co.y * ((c.y - a.y) * bx + (e.y - bq.y) * tb.ax)
                                 + co.z * ((c.z - a.z) * bx + (e.z - bq.z) * tb.ax) + co.w * ((c.w - a.w) * bx + (e.w - bq.w) * tb.ax);
                        sx = warp_sum(sx);
                        sy = warp_sum(sy);
                        if (lane == j) { gx[v] -= sx; gy[v] -= sy; }
                    }
                }
            }
            if (g_fmap != nullptr && chan_ok) *reinterpret_cast<float4*>(g_fmap + px) = gf;
        }
    }
    if (!need_coord_grad) return;
    float gd = 0.0f;
#pragma unroll
    for (int v = 0; v < VT; ++v) {
        if (v < V) {
            float gT[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) gT[i] = 0.0f;
            if (active) {
                Warp wp;
                warp_pixel(cam[v], x, y, d, wm1, hm1, true, wp);
                gd += warp_pixel_adjoint(cam[v], wp, d, wm1, hm1, true, gx[v] * t[v].mx, gy[v] * t[v].my, gT);
            }
            if (vg.g_pose[v] != nullptr) {
                Slot* slot = ws + (v * B + b);
                block_accumulate<12>(gT, red, slot->acc);
                if (last_block(slot, gridDim.x, &flag) && threadIdx.x == 0) {
                    const bool eul = cams.pose_kind == DROSFM_POSE_EULER6;
                    finish_pose_grad(slot, cams.pose_kind, eul ? vp.pose[v] + b * 6 : nullptr,
                                     vg.g_pose[v] + b * (eul ? 6 : 16));
                }
            }
        }
    }
    if (active && g_depth != nullptr)
        g_depth[static_cast<size_t>(b) * P + p] = depth_kind == DROSFM_INV_DEPTH ? inv2depth_grad(draw, gd) : gd;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int check_cost_args(const drosfm_cams_t* cams, const float* const* fmap_ref, const float* const* poses,
                           int n_views, int B, int C, int h, int w, int layout) {
    DROSFM_REQUIRE(cams && cams->K && cams->Kref, DROSFM_EINVAL, "feat_cost: NULL cams");
    DROSFM_REQUIRE(n_views >= 1 && n_views <= DROSFM_MAX_VIEWS, DROSFM_ERANGE, "feat_cost: n_views=%d outside [1,%d]",
                   n_views, DROSFM_MAX_VIEWS);
    DROSFM_REQUIRE(B >= 0 && C >= 0 && h >= 0 && w >= 0, DROSFM_EINVAL, "feat_cost: negative dimension");
    DROSFM_REQUIRE(B <= 65535 && static_cast<long long>(h) * w < (1ll << 26), DROSFM_ERANGE, "feat_cost: dimension out of range");
    DROSFM_REQUIRE(layout == DROSFM_NCHW || layout == DROSFM_NHWC, DROSFM_EINVAL, "feat_cost: bad layout %d", layout);
    DROSFM_REQUIRE(layout == DROSFM_NCHW || C % 4 == 0, DROSFM_ENOTSUP, "feat_cost: NHWC layout needs C %% 4 == 0 (C=%d)", C);
    DROSFM_REQUIRE(fmap_ref != nullptr && poses != nullptr, DROSFM_EINVAL, "feat_cost: NULL view arrays");
    DROSFM_REQUIRE(cams->pose_kind == DROSFM_POSE_MAT4 || cams->pose_kind == DROSFM_POSE_EULER6, DROSFM_EINVAL,
                   "feat_cost: pose_kind must be MAT4 or EULER6");
    for (int v = 0; v < n_views; ++v)
        DROSFM_REQUIRE(fmap_ref[v] != nullptr && poses[v] != nullptr, DROSFM_EINVAL, "feat_cost: view %d has a NULL pointer", v);
    return DROSFM_OK;
}

static int channel_group(int P, int B, int C) {
    // split channels over blocks until the grid covers the chip a few times
    const int pix_blocks = ((P + kPixThreads - 1) / kPixThreads) * (B > 0 ? B : 1);
    int cg = C;
    while (cg > 8 && pix_blocks * ((C + cg - 1) / cg) < kNumSMs * 4) cg = (cg + 1) / 2;
    return cg < 1 ? 1 : cg;
}

#define DISPATCH_VT(V, CALL)            \
    do {                                \
        if ((V) == 1) { CALL(1); }      \
        else if ((V) == 2) { CALL(2); } \
        else if ((V) <= 4) { CALL(4); } \
        else { CALL(8); }               \
    } while (0)

}  // namespace drosfm

using namespace drosfm;

extern "C" {

int drosfm_feat_cost_fwd(const float* fmap, const float* const* fmap_ref, const float* depth, int depth_kind,
                         const drosfm_cams_t* cams, const float* const* poses, int n_views, float* cost,
                         int B, int C, int h, int w, int layout, drosfm_stream_t stream) {
    if (B == 0 || C == 0 || h * w == 0) return DROSFM_OK;
    if (int e = check_cost_args(cams, fmap_ref, poses, n_views, B, C, h, w, layout)) return e;
    DROSFM_REQUIRE(fmap && depth && cost, DROSFM_EINVAL, "feat_cost_fwd: NULL argument");
    ViewPtrs vp{};
    for (int v = 0; v < n_views; ++v) { vp.ref[v] = fmap_ref[v]; vp.pose[v] = poses[v]; }
    const int P = h * w;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (layout == DROSFM_NHWC) {
        DROSFM_REQUIRE(aligned16(fmap) && aligned16(cost), DROSFM_EALIGN, "feat_cost_fwd: NHWC tensors must be 16-byte aligned");
        for (int v = 0; v < n_views; ++v)
            DROSFM_REQUIRE(aligned16(fmap_ref[v]), DROSFM_EALIGN, "feat_cost_fwd: NHWC tensors must be 16-byte aligned");
        dim3 grid((P + kWarpsPerBlock * 32 - 1) / (kWarpsPerBlock * 32), B);
#define CALL(VT) feat_cost_fwd_nhwc<VT><<<grid, kWarpsPerBlock * 32, 0, s>>>(fmap, vp, depth, depth_kind, *cams, n_views, cost, C, h, w)
        DISPATCH_VT(n_views, CALL);
#undef CALL
    } else {
        const int cg = channel_group(P, B, C);
        dim3 grid((P + kPixThreads - 1) / kPixThreads, (C + cg - 1) / cg, B);
#define CALL(VT) feat_cost_fwd_nchw<VT><<<grid, kPixThreads, 0, s>>>(fmap, vp, depth, depth_kind, *cams, n_views, cost, C, h, w, cg)
        DISPATCH_VT(n_views, CALL);
#undef CALL
    }
    return launch_status("feat_cost_fwd");
}

int drosfm_feat_cost_bwd(const float* g_cost, const float* fmap, const float* const* fmap_ref, const float* depth,
                         int depth_kind, const drosfm_cams_t* cams, const float* const* poses, int n_views,
                         float* g_fmap, float* const* g_fmap_ref, float* g_depth, float* const* g_poses, void* ws,
                         int B, int C, int h, int w, int layout, drosfm_stream_t stream) {
    if (B == 0 || h * w == 0 || C == 0) return DROSFM_OK;
    if (int e = check_cost_args(cams, fmap_ref, poses, n_views, B, C, h, w, layout)) return e;
    DROSFM_REQUIRE(g_cost && fmap && depth, DROSFM_EINVAL, "feat_cost_bwd: NULL argument");
    ViewPtrs vp{};
    ViewGrads vg{};
    bool want_pose = false;
    for (int v = 0; v < n_views; ++v) {
        vp.ref[v] = fmap_ref[v];
        vp.pose[v] = poses[v];
        vg.g_ref[v] = g_fmap_ref ? g_fmap_ref[v] : nullptr;
        vg.g_pose[v] = g_poses ? g_poses[v] : nullptr;
        want_pose |= vg.g_pose[v] != nullptr;
    }
    DROSFM_REQUIRE(!want_pose || ws != nullptr, DROSFM_EINVAL, "feat_cost_bwd: pose gradients need ws");
    const int need_coord = (want_pose || g_depth != nullptr) ? 1 : 0;
    const int P = h * w;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (layout == DROSFM_NHWC) {
        DROSFM_REQUIRE(aligned16(fmap) && aligned16(g_cost) && (!g_fmap || aligned16(g_fmap)), DROSFM_EALIGN,
                       "feat_cost_bwd: NHWC tensors must be 16-byte aligned");
        for (int v = 0; v < n_views; ++v)
            DROSFM_REQUIRE(aligned16(fmap_ref[v]) && (!vg.g_ref[v] || aligned16(vg.g_ref[v])), DROSFM_EALIGN,
                           "feat_cost_bwd: NHWC tensors must be 16-byte aligned");
        dim3 grid((P + kWarpsPerBlock * 32 - 1) / (kWarpsPerBlock * 32), B);
#define CALL(VT) feat_cost_bwd_nhwc<VT><<<grid, kWarpsPerBlock * 32, 0, s>>>(g_cost, fmap, vp, depth, depth_kind, *cams, n_views, \
                                                                       g_fmap, vg, g_depth, static_cast<Slot*>(ws), B, C, h, w, need_coord)
        DISPATCH_VT(n_views, CALL);
#undef CALL
    } else {
        const int cg = channel_group(P, B, C);
        dim3 grid((P + kPixThreads - 1) / kPixThreads, (C + cg - 1) / cg, B);
#define CALL(VT) feat_cost_bwd_nchw<VT><<<grid, kPixThreads, 0, s>>>(g_cost, fmap, vp, depth, depth_kind, *cams, n_views, g_fmap, vg, \
                                                                 g_depth, static_cast<Slot*>(ws), B, C, h, w, cg, need_coord)
        DISPATCH_VT(n_views, CALL);
#undef CALL
    }
    return launch_status("feat_cost_bwd");
}

}  // extern "C"
