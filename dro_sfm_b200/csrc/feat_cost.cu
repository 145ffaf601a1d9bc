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
#include <cstdlib>
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
// NCHW: lane = pixel, CG channels per thread with every load of a view issued before the first use
// ------------------------------------------------------------------------------------------
constexpr int kPixThreads = 128;

// Bilinear taps in "clamped" form: the four element offsets are always inside the source plane and an
// out-of-bounds tap carries weight 0, so the gathers need no predicates.
struct CTaps {
    int o00, dx, dy;            // offset of the clamped north-west tap, +dx = east, +dy = south
    float w00, w01, w10, w11;
    float ax, ay;
    unsigned valid;
    int x0, y0;
};

__device__ __forceinline__ void make_ctaps(float u, float v, int h, int w, CTaps& c) {
    Taps t;
    make_taps(u, v, h, w, DROSFM_PAD_ZEROS, t);
    const float bx = 1.0f - t.ax, by = 1.0f - t.ay;
    c.w00 = (t.valid & 1u) ? bx * by : 0.0f;
    c.w01 = (t.valid & 2u) ? t.ax * by : 0.0f;
    c.w10 = (t.valid & 4u) ? bx * t.ay : 0.0f;
    c.w11 = (t.valid & 8u) ? t.ax * t.ay : 0.0f;
    const int x0 = max(t.x0, 0), y0 = max(t.y0, 0);
    const int x1 = min(t.x0 + 1, w - 1), y1 = min(t.y0 + 1, h - 1);
    c.o00 = y0 * w + x0;
    c.dx = x1 - x0;
    c.dy = (y1 - y0) * w;
    c.ax = t.ax; c.ay = t.ay; c.valid = t.valid; c.x0 = t.x0; c.y0 = t.y0;
}

template <int VT, int CG>
__global__ void __launch_bounds__(kPixThreads)
feat_cost_fwd_nchw(const float* __restrict__ fmap, ViewPtrs vp, const float* __restrict__ depth, int depth_kind,
                   drosfm_cams_t cams, int V, float* __restrict__ cost, int C, int h, int w) {
    __shared__ Cam cam[VT];
    const int b = blockIdx.z, P = h * w;
    const int p = blockIdx.x * kPixThreads + threadIdx.x;
    const bool active = p < P;
    const int c0 = blockIdx.y * CG;
    const size_t base = (static_cast<size_t>(b) * C + c0) * P + p;
    // loads that do not depend on the coordinates go out first and overlap the camera set-up
    const float draw = active ? __ldg(depth + static_cast<size_t>(b) * P + p) : 0.0f;
    float f[CG];
#pragma unroll
    for (int k = 0; k < CG; ++k) f[k] = (active && c0 + k < C) ? __ldg(fmap + base + static_cast<size_t>(k) * P) : 0.0f;
#pragma unroll
    for (int v = 0; v < VT; ++v)
        if (v < V && threadIdx.x == v) setup_cam(cams, vp.pose[v], b, cam[v]);
    __syncthreads();
    if (!active) return;
    const float d = to_depth(draw, depth_kind);
    const float wm1 = static_cast<float>(w - 1), hm1 = static_cast<float>(h - 1);
    int x, y;
    pix_xy(p, w, x, y);
    float acc[CG];
#pragma unroll
    for (int k = 0; k < CG; ++k) acc[k] = 0.0f;
#pragma unroll
    for (int v = 0; v < VT; ++v) {
        if (v < V) {
            Warp wp;
            warp_pixel<true>(cam[v], x, y, d, wm1, hm1, true, wp);
            CTaps t;
            make_ctaps(wp.p.u, wp.p.v, h, w, t);
            const float* r = vp.ref[v] + (static_cast<size_t>(b) * C + c0) * P + t.o00;
            float tv[CG][4];
#pragma unroll
            for (int k = 0; k < CG; ++k) {
                const float* rk = r + static_cast<size_t>(c0 + k < C ? k : 0) * P;
                tv[k][0] = __ldg(rk);
                tv[k][1] = __ldg(rk + t.dx);
                tv[k][2] = __ldg(rk + t.dy);
                tv[k][3] = __ldg(rk + t.dy + t.dx);
            }
#pragma unroll
            for (int k = 0; k < CG; ++k) {
                const float df = f[k] - (tv[k][0] * t.w00 + tv[k][1] * t.w01 + tv[k][2] * t.w10 + tv[k][3] * t.w11);
                acc[k] += df * df;
            }
        }
    }
    const float fV = static_cast<float>(V);
#pragma unroll
    for (int k = 0; k < CG; ++k)
        if (c0 + k < C) cost[base + static_cast<size_t>(k) * P] = V == 1 ? acc[k] : acc[k] / fV;
}

// ------------------------------------------------------------------------------------------
// NCHW backward
// ------------------------------------------------------------------------------------------
struct MergePlan {
    bool give, take;
};

// Lane L hands its east taps to lane L+1 when that lane's west taps are the same source pixels.
__device__ __forceinline__ MergePlan plan_merge(const CTaps& t, bool active) {
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

__device__ __forceinline__ void scatter_taps(float* __restrict__ plane, const CTaps& t, const MergePlan& m, float g) {
    const unsigned full = 0xffffffffu;
    float cnw = g * t.w00, cne = g * t.w01, csw = g * t.w10, cse = g * t.w11;
    const float rn = __shfl_up_sync(full, cne, 1), rs = __shfl_up_sync(full, cse, 1);
    if (m.take) { cnw += rn; csw += rs; }
    float* r0 = plane + t.o00;
    if (t.valid & 1u) atomicAdd(r0, cnw);
    if (t.valid & 4u) atomicAdd(r0 + t.dy, csw);
    if (!m.give) {
        if (t.valid & 2u) atomicAdd(r0 + t.dx, cne);
        if (t.valid & 8u) atomicAdd(r0 + t.dy + t.dx, cse);
    }
}

template <int VT, int CG>
__global__ void __launch_bounds__(kPixThreads)
feat_cost_bwd_nchw(const float* __restrict__ g_cost, const float* __restrict__ fmap, ViewPtrs vp,
                   const float* __restrict__ depth, int depth_kind, drosfm_cams_t cams, int V,
                   float* __restrict__ g_fmap, ViewGrads vg, float* __restrict__ g_depth, Slot* ws,
                   int B, int C, int h, int w, int need_coord_grad, int acc_fmap) {
    __shared__ Cam cam[VT];
    __shared__ int flag;
    const int b = blockIdx.z, P = h * w;
    const int p = blockIdx.x * kPixThreads + threadIdx.x;
    const bool active = p < P;
    const int c0 = blockIdx.y * CG;
    const size_t base = (static_cast<size_t>(b) * C + c0) * P + p;
    const float draw = active ? __ldg(depth + static_cast<size_t>(b) * P + p) : 0.0f;
    const float scale = 2.0f / static_cast<float>(V);
    float g[CG], f[CG], gf[CG];
#pragma unroll
    for (int k = 0; k < CG; ++k) {
        const bool ok = active && c0 + k < C;
        g[k] = ok ? __ldg(g_cost + base + static_cast<size_t>(k) * P) * scale : 0.0f;
        f[k] = ok ? __ldg(fmap + base + static_cast<size_t>(k) * P) : 0.0f;
        gf[k] = 0.0f;
    }
#pragma unroll
    for (int v = 0; v < VT; ++v)
        if (v < V && threadIdx.x == v) setup_cam(cams, vp.pose[v], b, cam[v]);
    __syncthreads();
    const float d = to_depth(draw, depth_kind);
    const float wm1 = static_cast<float>(w - 1), hm1 = static_cast<float>(h - 1);
    int x = 0, y = 0;
    if (active) pix_xy(p, w, x, y);
    float gd = 0.0f;
#pragma unroll
    for (int v = 0; v < VT; ++v) {
        if (v < V) {
            CTaps t;
            t.o00 = t.dx = t.dy = 0; t.w00 = t.w01 = t.w10 = t.w11 = 0.0f; t.ax = t.ay = 0.0f; t.valid = 0u; t.x0 = t.y0 = 0;
            Warp wp;
            if (active) {
                warp_pixel<true>(cam[v], x, y, d, wm1, hm1, true, wp);
                make_ctaps(wp.p.u, wp.p.v, h, w, t);
            }
            const float* r = vp.ref[v] + (static_cast<size_t>(b) * C + c0) * P + t.o00;
            float tv[CG][4];
#pragma unroll
            for (int k = 0; k < CG; ++k) {
                const float* rk = r + static_cast<size_t>(c0 + k < C ? k : 0) * P;
                tv[k][0] = __ldg(rk);
                tv[k][1] = __ldg(rk + t.dx);
                tv[k][2] = __ldg(rk + t.dy);
                tv[k][3] = __ldg(rk + t.dy + t.dx);
            }
            MergePlan mp;
            mp.give = mp.take = false;
            float* gref = vg.g_ref[v];
            if (gref != nullptr) mp = plan_merge(t, active);
            float gx = 0.0f, gy = 0.0f;
            const float bx = 1.0f - t.ax, by = 1.0f - t.ay;
#pragma unroll
            for (int k = 0; k < CG; ++k) {
                // d cost / d fmap  (= -d cost / d warped)
                const float coef = (f[k] - (tv[k][0] * t.w00 + tv[k][1] * t.w01 + tv[k][2] * t.w10 + tv[k][3] * t.w11)) * g[k];
                gf[k] += coef;
                // taps outside the source count as zeros in the coordinate gradient
                const float a0 = (t.valid & 1u) ? tv[k][0] : 0.0f, a1 = (t.valid & 2u) ? tv[k][1] : 0.0f;
                const float a2 = (t.valid & 4u) ? tv[k][2] : 0.0f, a3 = (t.valid & 8u) ? tv[k][3] : 0.0f;
                gx -= coef * ((a1 - a0) * by + (a3 - a2) * t.ay);
                gy -= coef * ((a2 - a0) * bx + (a3 - a1) * t.ax);
                if (gref != nullptr && c0 + k < C)
                    scatter_taps(gref + (static_cast<size_t>(b) * C + c0 + k) * P, t, mp, -coef);
            }
            if (need_coord_grad) {
                float gT[12];
#pragma unroll
                for (int i = 0; i < 12; ++i) gT[i] = 0.0f;
                if (active && t.valid) {
                    const float mx = 0.5f * static_cast<float>(w - 1), my = 0.5f * static_cast<float>(h - 1);
                    gd += warp_pixel_adjoint(cam[v], wp, d, wm1, hm1, true, gx * mx, gy * my, gT);
                }
                if (vg.g_pose[v] != nullptr) warp_accumulate12(gT, spread_acc(slot_at(ws, v * B + b)));
            }
        }
    }
    if (active && g_fmap != nullptr) {
#pragma unroll
        for (int k = 0; k < CG; ++k)
            if (c0 + k < C) {
                float* gp = g_fmap + base + static_cast<size_t>(k) * P;
                *gp = acc_fmap ? *gp + gf[k] : gf[k];
            }
    }
    if (need_coord_grad) {
        // one ticket per block and view: the last block converts the fp64 sums into the caller's encoding
#pragma unroll
        for (int v = 0; v < VT; ++v) {
            if (v < V && vg.g_pose[v] != nullptr) {
                Slot* slot = slot_at(ws, v * B + b);
                if (last_block(slot, gridDim.x * gridDim.y, &flag) && threadIdx.x < 32) {
                    const bool eul = cams.pose_kind == DROSFM_POSE_EULER6;
                    finish_pose_grad_warp(slot, cams.pose_kind, eul ? vp.pose[v] + b * 6 : nullptr,
                                          vg.g_pose[v] + b * (eul ? 6 : 16));
                }
            }
        }
    }
    if (active) {
        if (need_coord_grad && g_depth != nullptr) {
            if (depth_kind == DROSFM_INV_DEPTH) gd = inv2depth_grad(draw, gd);
            if (gridDim.y == 1) g_depth[static_cast<size_t>(b) * P + p] = gd;
            else atomicAdd(g_depth + static_cast<size_t>(b) * P + p, gd);
        }
    }
}

// ------------------------------------------------------------------------------------------
// NHWC (torch channels_last): warp = `ppw` consecutive pixels, lane = 4 channels of a 128-channel slab.
// Every tap, load and store is one 128-bit access per lane, i.e. 512 contiguous bytes per warp.
// Lanes 0..ppw-1 first compute the taps of "their" pixel and park them in shared memory; the warp then
// walks over the pixels with all lanes on the channel axis.  ppw is chosen by the host: small for small
// maps (more warps in flight), 32 for large ones (coordinate work amortised over more bytes).
//
// One launch evaluates a BATCH of independent cost calls ("jobs", blockIdx.z): within one step of the recurrent
// optimiser the depth cost (V views, mean) and the V per-view pose costs depend only on the state at the start of the
// step (DepthPoseNet.py:159-167 builds all cost closures before either update block runs), so a caller that advances
// the update blocks in lock-step submits them together.  At the training shapes a single call is a one-wave launch
// bound by the latency of one warp's chain; three or more jobs per launch overlap each other's phases.
// ------------------------------------------------------------------------------------------
#ifndef DROSFM_COST_WARPS
#define DROSFM_COST_WARPS 4
#endif
constexpr int kWarpsPerBlock = DROSFM_COST_WARPS;      // >= 3 (camera set-up of multi-view jobs uses three warps)
constexpr int kMaxPpw = 32;
constexpr int kMaxJobs = DROSFM_MAX_COST_JOBS;

struct CostJob {
    const float* fmap;
    const float* depth;
    float* cost;
    const float* ref[DROSFM_MAX_VIEWS];
    const float* pose[DROSFM_MAX_VIEWS];
    int V, depth_kind;
    float disp_min, disp_range;      // DROSFM_DISP: inverse depth = disp_min + disp_range * raw (disp_to_depth, layers.py:11-20)
};

// depth of one pixel from what the job stores (depth | inverse depth | raw disparity)
__device__ __forceinline__ float job_depth(const CostJob& job, float raw) {
    if (job.depth_kind == DROSFM_DISP) return inv2depth_fast(__fadd_rn(job.disp_min, __fmul_rn(job.disp_range, raw)));
    return to_depth_fast(raw, job.depth_kind);
}
// ... and the gradient back to the stored quantity
__device__ __forceinline__ float job_depth_grad(const CostJob& job, float raw, float g_depth) {
    if (job.depth_kind == DROSFM_DISP)
        return inv2depth_grad(__fadd_rn(job.disp_min, __fmul_rn(job.disp_range, raw)), g_depth) * job.disp_range;
    return job.depth_kind == DROSFM_INV_DEPTH ? inv2depth_grad(raw, g_depth) : g_depth;
}
struct CostJobGrad {
    const float* g_cost;
    float* g_fmap;
    float* g_depth;
    float* g_ref[DROSFM_MAX_VIEWS];
    float* g_pose[DROSFM_MAX_VIEWS];
    int acc_fmap, need_coord, slot0, pad_;      // slot0: first workspace slot of the job (V * B slots)
};
struct CostJobs { CostJob j[kMaxJobs]; };
struct CostJobGrads { CostJobGrad j[kMaxJobs]; };

struct alignas(16) STap {      // 32 bytes: two broadcast LDS.128 per pixel and view
    int o00, dx, dy;           // element offsets (pixel units, to be multiplied by C)
    unsigned valid;
    float w00, w01, w10, w11;
};

__device__ __forceinline__ void store_stap(STap* dst, const CTaps& t) {
    dst->o00 = t.o00; dst->dx = t.dx; dst->dy = t.dy; dst->valid = t.valid;
    dst->w00 = t.w00; dst->w01 = t.w01; dst->w10 = t.w10; dst->w11 = t.w11;
}

__device__ __forceinline__ float4 ldg4f(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float4 blend4(const float4& a, const float4& b, const float4& c, const float4& d,
                                         float wa, float wb, float wc, float wd) {
    return make_float4(a.x * wa + b.x * wb + c.x * wc + d.x * wd, a.y * wa + b.y * wb + c.y * wc + d.y * wd,
                       a.z * wa + b.z * wb + c.z * wc + d.z * wd, a.w * wa + b.w * wb + c.w * wc + d.w * wd);
}

// Camera set-up of a block for the V views of its job: the three parts of every view's set-up run in three warps
// (lane = view), one barrier.
template <int VT>
__device__ __forceinline__ void setup_cams_block(const drosfm_cams_t& cams, const CostJob& job, int b, Cam* cam) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (wid < 3 && lane < job.V) setup_cam_part(cams, job.pose[lane], b, cam[lane], wid);      // kWarpsPerBlock >= 3
    __syncthreads();
}

template <int VT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
feat_cost_fwd_nhwc(const __grid_constant__ CostJobs jobs, drosfm_cams_t cams, int C, int h, int w, int ppw) {
    __shared__ Cam cam[VT];
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    STap* taps = reinterpret_cast<STap*>(dyn_smem);          // [warp][ppw][VT]
    const CostJob& job = jobs.j[blockIdx.z];
    const int V = job.V;
    const float* __restrict__ fmap = job.fmap;
    float* __restrict__ cost = job.cost;
    const int b = blockIdx.y, P = h * w;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int pbase = (blockIdx.x * kWarpsPerBlock + wid) * ppw;
    const int npix = max(0, min(ppw, P - pbase));
    const bool mine = lane < npix;
    const int p = pbase + lane;
    const float draw = mine ? __ldg(job.depth + static_cast<size_t>(b) * P + p) : 0.0f;
    setup_cams_block<VT>(cams, job, b, cam);
    if (npix == 0) return;
    STap* wt = taps + static_cast<size_t>(wid) * ppw * VT;
    if (mine) {
        const float d = job_depth(job, draw);
        const Norm nm = make_norm(w, h);
        int x, y;
        pix_xy(p, w, x, y);
#pragma unroll
        for (int v = 0; v < VT; ++v) {
            if (v < V) {
                Warp wp;
                warp_pixel_fast(cam[v], x, y, d, nm, true, wp);
                CTaps t;
                make_ctaps(wp.p.u, wp.p.v, h, w, t);
                store_stap(wt + lane * VT + v, t);
            }
        }
    }
    __syncwarp();
    const float fV = static_cast<float>(V);
    const size_t sample = static_cast<size_t>(b) * P * C;
    for (int cb0 = 0; cb0 < C; cb0 += 128) {
        const int cb = cb0 + lane * 4;
        if (cb >= C) continue;
#pragma unroll 4
        for (int j = 0; j < npix; ++j) {
            const size_t px = sample + static_cast<size_t>(pbase + j) * C + cb;
            const float4 f = ldg4f(fmap + px);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int v = 0; v < VT; ++v) {
                if (v < V) {
                    const STap t = wt[j * VT + v];
                    const float* r0 = job.ref[v] + sample + static_cast<size_t>(t.o00) * C + cb;
                    const float4 a = ldg4f(r0), bq = ldg4f(r0 + t.dx * C);
                    const float4 c = ldg4f(r0 + t.dy * C), e = ldg4f(r0 + (t.dy + t.dx) * C);
                    const float4 wv = blend4(a, bq, c, e, t.w00, t.w01, t.w10, t.w11);
                    const float dx = f.x - wv.x, dy = f.y - wv.y, dz = f.z - wv.z, dw = f.w - wv.w;
                    acc.x += dx * dx; acc.y += dy * dy; acc.z += dz * dz; acc.w += dw * dw;
                }
            }
            if (V != 1) { acc.x /= fV; acc.y /= fV; acc.z /= fV; acc.w /= fV; }
            *reinterpret_cast<float4*>(cost + px) = acc;
        }
    }
}

// red.global.add.v4.f32 (sm_90+): one 16-byte reduction per lane, 512 contiguous bytes per warp.  Neither this nor the
// store below carries a memory clobber, so the compiler may hoist the read-only loads of the next pixel across them;
// both are `asm volatile`, which keeps them in program order RELATIVE TO EACH OTHER (the first view's plain store of a
// target-map gradient is followed by the other views' reductions into the same address).
__device__ __forceinline__ void red_add4_nc(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d));
}
__device__ __forceinline__ void st_global4_nc(float* p, float a, float b, float c, float d) {
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d));
}

// Backward, NHWC: same pixel-to-warp mapping as the forward.  Per pixel and view the warp loads the upstream
// gradient, the target features and the four taps (six 512-byte accesses), issues four 512-byte
// red.global.add.v4.f32 into the source gradient and reduces the coordinate gradient with shuffles.  The loads
// of pixel j+1 are issued before pixel j is consumed (software pipeline).
struct PixLoad {
    float4 g, f, a, b, c, e;
};

// Register budget per template: four and eight views keep per-view state for every view and get 4 / 3 blocks per SM
// instead of spilling (round 1's 5 blocks/SM spilled 104 / 256 local-memory accesses in the V = 4 / 8 kernels).
// one / two views: 4 blocks per SM (128 registers): cost phase of the benchmark step 360 us; 5 blocks 371, 6 blocks 371
#ifndef DROSFM_COST_BWD_ALLVALID
#define DROSFM_COST_BWD_ALLVALID 1
#endif
#ifndef DROSFM_COST_BWD_BLOCKS
#define DROSFM_COST_BWD_BLOCKS 4
#endif
template <int VT> struct BwdBlocks { static constexpr int value = (VT <= 2 ? DROSFM_COST_BWD_BLOCKS : (VT <= 4 ? 4 : 3)) * 4 / kWarpsPerBlock; };

template <int VT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, BwdBlocks<VT>::value)
feat_cost_bwd_nhwc(const __grid_constant__ CostJobs jobs, const __grid_constant__ CostJobGrads grads, drosfm_cams_t cams,
                   Slot* ws, int B, int C, int h, int w, int ppw) {
    __shared__ Cam cam[VT];
    __shared__ int flag;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    STap* taps = reinterpret_cast<STap*>(dyn_smem);                                                   // [warp][ppw][VT]
    float2* gxy = reinterpret_cast<float2*>(taps + static_cast<size_t>(kWarpsPerBlock) * ppw * VT);   // [warp][ppw][VT]
    const CostJob& job = jobs.j[blockIdx.z];
    const CostJobGrad& jg = grads.j[blockIdx.z];
    const int V = job.V, need_coord_grad = jg.need_coord, acc_fmap = jg.acc_fmap;
    const float* __restrict__ g_cost = jg.g_cost;
    const float* __restrict__ fmap = job.fmap;
    float* __restrict__ g_fmap = jg.g_fmap;
    const int b = blockIdx.y, P = h * w;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int pbase = (blockIdx.x * kWarpsPerBlock + wid) * ppw;
    const int npix = max(0, min(ppw, P - pbase));
    const bool mine = lane < npix;
    const int p = pbase + lane;
    const float draw = mine ? __ldg(job.depth + static_cast<size_t>(b) * P + p) : 0.0f;
    const float d = job_depth(job, draw);
    setup_cams_block<VT>(cams, job, b, cam);
    const Norm nm = make_norm(w, h);
    const float wm1 = nm.wm1, hm1 = nm.hm1;
    STap* wt = taps + static_cast<size_t>(wid) * ppw * VT;
    float2* wg = gxy + static_cast<size_t>(wid) * ppw * VT;
    int x = 0, y = 0;
    if (mine) {
        pix_xy(p, w, x, y);
#pragma unroll
        for (int v = 0; v < VT; ++v) {
            if (v < V) {
                Warp wp;
                warp_pixel_fast(cam[v], x, y, d, nm, true, wp);
                CTaps t;
                make_ctaps(wp.p.u, wp.p.v, h, w, t);
                store_stap(wt + lane * VT + v, t);
                // the fractional offsets (needed once per pixel by the coordinate gradient) wait in shared memory
                wg[lane * VT + v] = make_float2(t.ax, t.ay);
            }
        }
    }
    __syncwarp();
    const float scale = 2.0f / static_cast<float>(V);
    const size_t sample = static_cast<size_t>(b) * P * C;
    float2 acc_g[VT];           // lane j: coordinate gradient of pixel j (d cost / d ix, d cost / d iy) per view
#pragma unroll
    for (int v = 0; v < VT; ++v) acc_g[v] = make_float2(0.f, 0.f);
    for (int cb0 = 0; cb0 < C; cb0 += 128) {
        const int cb = cb0 + lane * 4;
        const bool chan_ok = cb < C;      // lanes beyond C idle but still take part in the shuffles
        const int cbs = chan_ok ? cb : 0;
#pragma unroll
        for (int v = 0; v < VT; ++v) {
            if (v >= V || npix == 0) continue;
            float* gref = jg.g_ref[v];
            const float* ref = job.ref[v] + sample + cbs;
            auto fetch = [&](int j, const STap& t) {
                PixLoad q;
                const size_t px = sample + static_cast<size_t>(pbase + j) * C + cbs;
                q.g = ldg4f(g_cost + px);
                q.f = ldg4f(fmap + px);
                const float* r0 = ref + static_cast<size_t>(t.o00) * C;
                q.a = ldg4f(r0);
                q.b = ldg4f(r0 + t.dx * C);
                q.c = ldg4f(r0 + t.dy * C);
                q.e = ldg4f(r0 + (t.dy + t.dx) * C);
                return q;
            };
            STap t = wt[v];
            PixLoad cur = fetch(0, t);
            for (int j = 0; j < npix; ++j) {
                STap tn = t;
                PixLoad nxt = cur;
                if (j + 1 < npix) {
                    tn = wt[(j + 1) * VT + v];
                    nxt = fetch(j + 1, tn);
                }
                const size_t px = sample + static_cast<size_t>(pbase + j) * C + cbs;
                const float4 wv = blend4(cur.a, cur.b, cur.c, cur.e, t.w00, t.w01, t.w10, t.w11);
                float4 co = make_float4((cur.f.x - wv.x) * cur.g.x * scale, (cur.f.y - wv.y) * cur.g.y * scale,
                                        (cur.f.z - wv.z) * cur.g.z * scale, (cur.f.w - wv.w) * cur.g.w * scale);
                if (!chan_ok) co = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g_fmap != nullptr && chan_ok) {
                    // accumulation as a fire-and-forget reduction: a read-modify-write would put one exposed L2 round
                    // trip per pixel on the warp's critical path (21 % of the stall samples before this change)
                    if (v == 0 && !acc_fmap) st_global4_nc(g_fmap + px, co.x, co.y, co.z, co.w);
                    else red_add4_nc(g_fmap + px, co.x, co.y, co.z, co.w);
                }
                if (gref != nullptr && chan_ok) {
                    float* q0 = gref + sample + static_cast<size_t>(t.o00) * C + cbs;
                    const int odx = t.dx * C, ody = t.dy * C;
                    if (t.valid & 1u) red_add4_nc(q0, -co.x * t.w00, -co.y * t.w00, -co.z * t.w00, -co.w * t.w00);
                    if (t.valid & 2u) red_add4_nc(q0 + odx, -co.x * t.w01, -co.y * t.w01, -co.z * t.w01, -co.w * t.w01);
                    if (t.valid & 4u) red_add4_nc(q0 + ody, -co.x * t.w10, -co.y * t.w10, -co.z * t.w10, -co.w * t.w10);
                    if (t.valid & 8u) red_add4_nc(q0 + ody + odx, -co.x * t.w11, -co.y * t.w11, -co.z * t.w11, -co.w * t.w11);
                }
                if (need_coord_grad) {
                    // taps outside the source count as zeros in the coordinate gradient
                    float4 a = cur.a, bq = cur.b, c4 = cur.c, e = cur.e;
#if DROSFM_COST_BWD_ALLVALID
                    if (t.valid != 15u)      // warp-uniform (the tap record is a shared-memory broadcast); rare: image border
#endif
                    {
                        const float m0 = (t.valid & 1u) ? 1.f : 0.f, m1 = (t.valid & 2u) ? 1.f : 0.f;
                        const float m2 = (t.valid & 4u) ? 1.f : 0.f, m3 = (t.valid & 8u) ? 1.f : 0.f;
                        a = make_float4(a.x * m0, a.y * m0, a.z * m0, a.w * m0);
                        bq = make_float4(bq.x * m1, bq.y * m1, bq.z * m1, bq.w * m1);
                        c4 = make_float4(c4.x * m2, c4.y * m2, c4.z * m2, c4.w * m2);
                        e = make_float4(e.x * m3, e.y * m3, e.z * m3, e.w * m3);
                    }
                    const float s01 = co.x * (bq.x - a.x) + co.y * (bq.y - a.y) + co.z * (bq.z - a.z) + co.w * (bq.w - a.w);
                    const float s23 = co.x * (e.x - c4.x) + co.y * (e.y - c4.y) + co.z * (e.z - c4.z) + co.w * (e.w - c4.w);
                    const float s02 = co.x * (c4.x - a.x) + co.y * (c4.y - a.y) + co.z * (c4.z - a.z) + co.w * (c4.w - a.w);
                    const float s13 = co.x * (e.x - bq.x) + co.y * (e.y - bq.y) + co.z * (e.z - bq.z) + co.w * (e.w - bq.w);
                    // this lane's share of d cost / d (ix, iy) of pixel j (its fractional offsets are broadcast from shared
                    // memory), then ONE reduce-scatter over the warp: 6 shuffles instead of 4 butterflies of 5
                    const float2 fr = wg[j * VT + v];
                    const float px_ = s01 * (1.0f - fr.y) + s23 * fr.y, py_ = s02 * (1.0f - fr.x) + s13 * fr.x;
                    const bool hi = lane & 16;
                    float r = xchg_add(hi ? py_ : px_, hi ? px_ : py_, 16);      // lanes 0-15: x share, lanes 16-31: y share
                    r += __shfl_xor_sync(0xffffffffu, r, 8);
                    r += __shfl_xor_sync(0xffffffffu, r, 4);
                    r += __shfl_xor_sync(0xffffffffu, r, 2);
                    r += __shfl_xor_sync(0xffffffffu, r, 1);
                    const float other = __shfl_xor_sync(0xffffffffu, r, 16);
                    if (lane == j) {
                        acc_g[v].x -= hi ? other : r;
                        acc_g[v].y -= hi ? r : other;
                    }
                }
                cur = nxt;
                t = tn;
            }
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
            if (mine && wt[lane * VT + v].valid) {
                Warp wp;
                warp_pixel_fast(cam[v], x, y, d, nm, true, wp);
                const float mx = 0.5f * static_cast<float>(w - 1), my = 0.5f * static_cast<float>(h - 1);
                gd += warp_pixel_adjoint(cam[v], wp, d, wm1, hm1, true, acc_g[v].x * mx, acc_g[v].y * my, gT);
            }
            if (jg.g_pose[v] != nullptr) warp_accumulate12(gT, spread_acc(slot_at(ws, jg.slot0 + v * B + b)));
        }
    }
    if (mine && jg.g_depth != nullptr)
        jg.g_depth[static_cast<size_t>(b) * P + p] = job_depth_grad(job, draw, gd);
#pragma unroll
    for (int v = 0; v < VT; ++v) {
        if (v < V && jg.g_pose[v] != nullptr) {
            Slot* slot = slot_at(ws, jg.slot0 + v * B + b);
            if (last_block(slot, gridDim.x, &flag) && threadIdx.x < 32) {
                const bool eul = cams.pose_kind == DROSFM_POSE_EULER6;
                finish_pose_grad_warp(slot, cams.pose_kind, eul ? job.pose[v] + b * 6 : nullptr, jg.g_pose[v] + b * (eul ? 6 : 16));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static_assert(sizeof(drosfm_cost_job_t) == 56 && sizeof(drosfm_cost_job_grads_t) == 48, "ABI layout (see dro_sfm_b200/_lib.py)");

static int check_cost_dims(const drosfm_cams_t* cams, int B, int C, int h, int w, int layout) {
    DROSFM_REQUIRE(cams && cams->K && cams->Kref, DROSFM_EINVAL, "feat_cost: NULL cams");
    DROSFM_REQUIRE(B >= 0 && C >= 0 && h >= 0 && w >= 0, DROSFM_EINVAL, "feat_cost: negative dimension");
    DROSFM_REQUIRE(B <= 65535 && static_cast<long long>(h) * w < (1ll << 26), DROSFM_ERANGE, "feat_cost: dimension out of range");
    DROSFM_REQUIRE(layout == DROSFM_NCHW || layout == DROSFM_NHWC, DROSFM_EINVAL, "feat_cost: bad layout %d", layout);
    DROSFM_REQUIRE(layout == DROSFM_NCHW || C % 4 == 0, DROSFM_ENOTSUP, "feat_cost: NHWC layout needs C %% 4 == 0 (C=%d)", C);
    DROSFM_REQUIRE(cams->pose_kind == DROSFM_POSE_MAT4 || cams->pose_kind == DROSFM_POSE_EULER6, DROSFM_EINVAL,
                   "feat_cost: pose_kind must be MAT4 or EULER6");
    return DROSFM_OK;
}

static int check_cost_views(const float* const* fmap_ref, const float* const* poses, int n_views) {
    DROSFM_REQUIRE(n_views >= 1 && n_views <= DROSFM_MAX_VIEWS, DROSFM_ERANGE, "feat_cost: n_views=%d outside [1,%d]",
                   n_views, DROSFM_MAX_VIEWS);
    DROSFM_REQUIRE(fmap_ref != nullptr && poses != nullptr, DROSFM_EINVAL, "feat_cost: NULL view arrays");
    for (int v = 0; v < n_views; ++v)
        DROSFM_REQUIRE(fmap_ref[v] != nullptr && poses[v] != nullptr, DROSFM_EINVAL, "feat_cost: view %d has a NULL pointer", v);
    return DROSFM_OK;
}

// Channels per thread: 8 by default (all 8 * 4 gathers of a view in flight at once); 16 once the grid is
// already several waves deep, which halves the redundant coordinate work.
static int channel_group(int P, int B, int C) {
    const long long pix_blocks = static_cast<long long>((P + kPixThreads - 1) / kPixThreads) * (B > 0 ? B : 1);
    return (pix_blocks * ((C + 15) / 16) >= static_cast<long long>(kNumSMs) * 16) ? 16 : 8;
}

// NHWC: pixels per warp.  Enough warps to fill the chip several times over wins at small maps (each warp
// keeps ~5 * ppw 128-bit loads in flight); large maps amortise the per-warp coordinate work over 32 pixels.
// `units` = pixel-views of the whole launch (all jobs).
static int pixels_per_warp(long long units, bool backward) {
    // the forward holds 44-48 registers (all warps of a training-shape launch are resident at once: the more the better);
    // the backward 90-150 (fewer, longer warps).  Measured on the cost phase of the benchmark steps (fwd / bwd pixels per
    // warp): KITTI 4/8 = 8/8 = 360 us, 8/4 = 414, 16/16 = 410; ScanNet 5 views 4/8 = 567, 8/8 = 629, 16/16 = 889
    const long long want = static_cast<long long>(kNumSMs) * (backward ? 32 : 64);
    int ppw = kMaxPpw;
    while (ppw > 4 && units / ppw < want) ppw /= 2;
    // tuning knobs (DROSFM_PPW for both directions, DROSFM_PPW_FWD / _BWD for one); anything but a supported value is ignored
    for (const char* name : {"DROSFM_PPW", backward ? "DROSFM_PPW_BWD" : "DROSFM_PPW_FWD"}) {
        if (const char* e = std::getenv(name)) {
            const int v = atoi(e);
            if (v == 4 || v == 8 || v == 16 || v == 32) ppw = v;
        }
    }
    return ppw;
}

#define DISPATCH_VT(V, CALL)            \
    do {                                \
        if ((V) == 1) { CALL(1); }      \
        else if ((V) == 2) { CALL(2); } \
        else if ((V) <= 4) { CALL(4); } \
        else { CALL(8); }               \
    } while (0)

static int vt_of(int v) { return v == 1 ? 1 : (v == 2 ? 2 : (v <= 4 ? 4 : 8)); }

// Validates a batch and fills the device-side job tables.  grads == nullptr: forward.
static int fill_jobs(const drosfm_cost_job_t* jobs, const drosfm_cost_job_grads_t* grads, int n_jobs, int B, void* ws,
                     CostJobs& cj, CostJobGrads& cg, int& max_v, long long& units, const char* who) {
    DROSFM_REQUIRE(jobs != nullptr && n_jobs >= 1 && n_jobs <= kMaxJobs, DROSFM_ERANGE, "%s: n_jobs=%d outside [1,%d]", who, n_jobs,
                   kMaxJobs);
    max_v = 0;
    units = 0;
    int slot = 0;
    for (int k = 0; k < n_jobs; ++k) {
        const drosfm_cost_job_t& in = jobs[k];
        if (int e = check_cost_views(in.fmap_ref, in.poses, in.n_views)) return e;
        DROSFM_REQUIRE(in.fmap && in.depth && (grads != nullptr || in.cost), DROSFM_EINVAL, "%s: job %d has a NULL argument", who, k);
        DROSFM_REQUIRE(aligned16(in.fmap) && aligned16(in.cost), DROSFM_EALIGN, "%s: NHWC tensors must be 16-byte aligned", who);
        CostJob& o = cj.j[k];
        DROSFM_REQUIRE(in.depth_kind == DROSFM_DEPTH || in.depth_kind == DROSFM_INV_DEPTH || in.depth_kind == DROSFM_DISP, DROSFM_EINVAL,
                       "%s: job %d: bad depth_kind %d", who, k, in.depth_kind);
        o.fmap = in.fmap; o.depth = in.depth; o.cost = in.cost; o.V = in.n_views; o.depth_kind = in.depth_kind;
        o.disp_min = in.disp_min; o.disp_range = in.disp_range;
        for (int v = 0; v < in.n_views; ++v) {
            DROSFM_REQUIRE(aligned16(in.fmap_ref[v]), DROSFM_EALIGN, "%s: NHWC tensors must be 16-byte aligned", who);
            o.ref[v] = in.fmap_ref[v];
            o.pose[v] = in.poses[v];
        }
        max_v = in.n_views > max_v ? in.n_views : max_v;
        units += in.n_views;
        if (grads != nullptr) {
            const drosfm_cost_job_grads_t& gi = grads[k];
            CostJobGrad& g = cg.j[k];
            DROSFM_REQUIRE(gi.g_cost != nullptr && aligned16(gi.g_cost) && (!gi.g_fmap || aligned16(gi.g_fmap)), DROSFM_EALIGN,
                           "%s: job %d: g_cost is NULL or a gradient tensor is not 16-byte aligned", who, k);
            g.g_cost = gi.g_cost; g.g_fmap = gi.g_fmap; g.g_depth = gi.g_depth;
            g.acc_fmap = (gi.flags & DROSFM_ACCUMULATE_FMAP) ? 1 : 0;
            bool want_pose = false;
            for (int v = 0; v < in.n_views; ++v) {
                g.g_ref[v] = gi.g_fmap_ref ? gi.g_fmap_ref[v] : nullptr;
                g.g_pose[v] = gi.g_poses ? gi.g_poses[v] : nullptr;
                DROSFM_REQUIRE(!g.g_ref[v] || aligned16(g.g_ref[v]), DROSFM_EALIGN, "%s: NHWC tensors must be 16-byte aligned", who);
                want_pose |= g.g_pose[v] != nullptr;
            }
            DROSFM_REQUIRE(!want_pose || ws != nullptr, DROSFM_EINVAL, "%s: pose gradients need ws", who);
            g.need_coord = (want_pose || g.g_depth != nullptr) ? 1 : 0;
            g.slot0 = slot;
            g.pad_ = 0;
            slot += in.n_views * B;
        }
    }
    return DROSFM_OK;
}

static int launch_cost_fwd_nhwc(const CostJobs& cj, int n_jobs, int max_v, long long units, const drosfm_cams_t* cams,
                                int B, int C, int h, int w, cudaStream_t s) {
    const int P = h * w, VT = vt_of(max_v);
    const int ppw = pixels_per_warp(static_cast<long long>(P) * B * units, false);
    const int per_block = kWarpsPerBlock * ppw;
    dim3 grid((P + per_block - 1) / per_block, B, n_jobs);
    const size_t smem = static_cast<size_t>(kWarpsPerBlock) * ppw * VT * sizeof(STap);
#define CALL(VT_) feat_cost_fwd_nhwc<VT_><<<grid, kWarpsPerBlock * 32, smem, s>>>(cj, *cams, C, h, w, ppw)
    DISPATCH_VT(max_v, CALL);
#undef CALL
    return launch_status("feat_cost_fwd");
}

static int launch_cost_bwd_nhwc(const CostJobs& cj, const CostJobGrads& cg, int n_jobs, int max_v, long long units,
                                const drosfm_cams_t* cams, void* ws, int B, int C, int h, int w, cudaStream_t s) {
    const int P = h * w, VT = vt_of(max_v);
    const int ppw = pixels_per_warp(static_cast<long long>(P) * B * units, true);
    const int per_block = kWarpsPerBlock * ppw;
    dim3 grid((P + per_block - 1) / per_block, B, n_jobs);
    const size_t smem = static_cast<size_t>(kWarpsPerBlock) * ppw * VT * (sizeof(STap) + sizeof(float2));
#define CALL(VT_) feat_cost_bwd_nhwc<VT_><<<grid, kWarpsPerBlock * 32, smem, s>>>(cj, cg, *cams, static_cast<Slot*>(ws), B, C, h, w, ppw)
    DISPATCH_VT(max_v, CALL);
#undef CALL
    return launch_status("feat_cost_bwd");
}

}  // namespace drosfm

using namespace drosfm;

extern "C" {

int drosfm_feat_cost_batch_fwd(const drosfm_cost_job_t* jobs, int n_jobs, const drosfm_cams_t* cams,
                               int B, int C, int h, int w, int layout, drosfm_stream_t stream) {
    if (B == 0 || C == 0 || h * w == 0 || n_jobs == 0) return DROSFM_OK;
    if (int e = check_cost_dims(cams, B, C, h, w, layout)) return e;
    DROSFM_REQUIRE(layout == DROSFM_NHWC, DROSFM_ENOTSUP, "feat_cost_batch_fwd: batches run on the NHWC (channels_last) layout only");
    CostJobs cj{};
    CostJobGrads cg{};
    int max_v;
    long long units;
    if (int e = fill_jobs(jobs, nullptr, n_jobs, B, nullptr, cj, cg, max_v, units, "feat_cost_batch_fwd")) return e;
    return launch_cost_fwd_nhwc(cj, n_jobs, max_v, units, cams, B, C, h, w, static_cast<cudaStream_t>(stream));
}

int drosfm_feat_cost_batch_bwd(const drosfm_cost_job_t* jobs, const drosfm_cost_job_grads_t* grads, int n_jobs,
                               const drosfm_cams_t* cams, void* ws, int B, int C, int h, int w, int layout,
                               drosfm_stream_t stream) {
    if (B == 0 || C == 0 || h * w == 0 || n_jobs == 0) return DROSFM_OK;
    if (int e = check_cost_dims(cams, B, C, h, w, layout)) return e;
    DROSFM_REQUIRE(layout == DROSFM_NHWC, DROSFM_ENOTSUP, "feat_cost_batch_bwd: batches run on the NHWC (channels_last) layout only");
    DROSFM_REQUIRE(grads != nullptr, DROSFM_EINVAL, "feat_cost_batch_bwd: NULL grads");
    CostJobs cj{};
    CostJobGrads cg{};
    int max_v;
    long long units;
    if (int e = fill_jobs(jobs, grads, n_jobs, B, ws, cj, cg, max_v, units, "feat_cost_batch_bwd")) return e;
    return launch_cost_bwd_nhwc(cj, cg, n_jobs, max_v, units, cams, ws, B, C, h, w, static_cast<cudaStream_t>(stream));
}

int drosfm_feat_cost_fwd(const float* fmap, const float* const* fmap_ref, const float* depth, int depth_kind,
                         const drosfm_cams_t* cams, const float* const* poses, int n_views, float* cost,
                         int B, int C, int h, int w, int layout, drosfm_stream_t stream) {
    if (B == 0 || C == 0 || h * w == 0) return DROSFM_OK;
    if (int e = check_cost_dims(cams, B, C, h, w, layout)) return e;
    if (int e = check_cost_views(fmap_ref, poses, n_views)) return e;
    DROSFM_REQUIRE(fmap && depth && cost, DROSFM_EINVAL, "feat_cost_fwd: NULL argument");
    if (layout == DROSFM_NHWC) {
        DROSFM_REQUIRE(depth_kind != DROSFM_DISP, DROSFM_ENOTSUP, "feat_cost_fwd: DROSFM_DISP needs the batch entry (drosfm_cost_job_t carries the range)");
        const drosfm_cost_job_t job{fmap, fmap_ref, depth, depth_kind, n_views, poses, cost, 0.0f, 1.0f};
        return drosfm_feat_cost_batch_fwd(&job, 1, cams, B, C, h, w, layout, stream);
    }
    ViewPtrs vp{};
    for (int v = 0; v < n_views; ++v) { vp.ref[v] = fmap_ref[v]; vp.pose[v] = poses[v]; }
    const int P = h * w;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int cg = channel_group(P, B, C);
    dim3 grid((P + kPixThreads - 1) / kPixThreads, (C + cg - 1) / cg, B);
#define CALL(VT)                                                                                                            \
    do {                                                                                                                    \
        if (cg == 16) feat_cost_fwd_nchw<VT, 16><<<grid, kPixThreads, 0, s>>>(fmap, vp, depth, depth_kind, *cams, n_views, cost, C, h, w); \
        else feat_cost_fwd_nchw<VT, 8><<<grid, kPixThreads, 0, s>>>(fmap, vp, depth, depth_kind, *cams, n_views, cost, C, h, w);           \
    } while (0)
    DISPATCH_VT(n_views, CALL);
#undef CALL
    return launch_status("feat_cost_fwd");
}

int drosfm_feat_cost_bwd(const float* g_cost, const float* fmap, const float* const* fmap_ref, const float* depth,
                         int depth_kind, const drosfm_cams_t* cams, const float* const* poses, int n_views,
                         float* g_fmap, float* const* g_fmap_ref, float* g_depth, float* const* g_poses, void* ws,
                         int B, int C, int h, int w, int layout, int flags, drosfm_stream_t stream) {
    const int acc_fmap = (flags & DROSFM_ACCUMULATE_FMAP) ? 1 : 0;
    if (B == 0 || h * w == 0 || C == 0) return DROSFM_OK;
    if (int e = check_cost_dims(cams, B, C, h, w, layout)) return e;
    if (int e = check_cost_views(fmap_ref, poses, n_views)) return e;
    DROSFM_REQUIRE(g_cost && fmap && depth, DROSFM_EINVAL, "feat_cost_bwd: NULL argument");
    if (layout == DROSFM_NHWC) {
        DROSFM_REQUIRE(depth_kind != DROSFM_DISP, DROSFM_ENOTSUP, "feat_cost_bwd: DROSFM_DISP needs the batch entry");
        const drosfm_cost_job_t job{fmap, fmap_ref, depth, depth_kind, n_views, poses, nullptr, 0.0f, 1.0f};
        const drosfm_cost_job_grads_t jg{g_cost, g_fmap, g_fmap_ref, g_depth, g_poses, flags};
        return drosfm_feat_cost_batch_bwd(&job, &jg, 1, cams, ws, B, C, h, w, layout, stream);
    }
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
    // the backward keeps g, f, gf and 4 taps per channel live: 4 channels per thread (8 for deep grids)
    const int cg = channel_group(P, B, C) == 16 ? 8 : 4;
    dim3 grid((P + kPixThreads - 1) / kPixThreads, (C + cg - 1) / cg, B);
#define CALL(VT)                                                                                                            \
    do {                                                                                                                    \
        if (cg == 8)                                                                                                        \
            feat_cost_bwd_nchw<VT, 8><<<grid, kPixThreads, 0, s>>>(g_cost, fmap, vp, depth, depth_kind, *cams, n_views, g_fmap, vg,  \
                                                                   g_depth, static_cast<Slot*>(ws), B, C, h, w, need_coord, acc_fmap);      \
        else                                                                                                                \
            feat_cost_bwd_nchw<VT, 4><<<grid, kPixThreads, 0, s>>>(g_cost, fmap, vp, depth, depth_kind, *cams, n_views, g_fmap, vg,  \
                                                                   g_depth, static_cast<Slot*>(ws), B, C, h, w, need_coord, acc_fmap);      \
    } while (0)
    DISPATCH_VT(n_views, CALL);
#undef CALL
    return launch_status("feat_cost_bwd");
}

}  // extern "C"
