// Kernel family 4: single-pass multi-view photometric loss (SSIM + L1, auto-mask, per-pixel min).
//
//   SSIM                      dro_sfm/losses/multiview_photometric_loss_mf.py:15-54
//   warp_ref_image            :132-171   (view_synthesis of every source view for every prediction)
//   calc_photometric_loss     :194-229   (clamp((1-SSIM)/2,0,1), 0.85*mean_c ssim + 0.15*mean_c |a-b|)
//   auto-mask                 :346-351   (un-warped source maps join the per-pixel min)
//   reduce_photometric_loss   :231-269   (cat(2V maps).min(1).mean(), gamma-decay over predictions)
//
// One block owns a 2-D tile of one (sample, prediction).  For every source view it warps the source
// into shared memory (tile + halo; the coordinate chain of common.cuh is fused in, nothing is
// materialised in HBM), evaluates the 3x3 SSIM statistics with a sliding window over shared memory,
// and keeps the running per-pixel minimum in registers.  Reflection padding of the SSIM window
// (ReflectionPad2d(1)) is an index remap into the same tile.  The un-warped (auto-mask) minimum does
// not depend on the prediction and is computed once per step by drosfm_automask_fwd instead of once per
// prediction.  The loss scalar is reduced in fp64; sel records the arg-min view for the backward pass.
//
// Backward: dSSIM_p/dx_q = A_p + B_p x_q + C_p y_q for every tap q of the window of p, so the gradient
// w.r.t. a warped pixel is a 3x3 box sum (with reflection multiplicities) of three coefficient maps,
// evaluated on tile + halo 1 from warped values on tile + halo 2, then pushed through the bilinear
// taps and the projection adjoint to d/d(inv_depth) and the fp64-reduced pose gradients.
//
// Algorithmic bytes per pixel and prediction: fwd 16 + 12 V (target 12 + depth 4 + V sources 12),
// bwd 20 + 12 V.
#include <type_traits>
#include "common.cuh"

namespace drosfm {


struct PhotoPtrs {
    const float* context[DROSFM_MAX_VIEWS];
    const float* inv_depth[DROSFM_MAX_PREDS];
    const float* pose[DROSFM_MAX_VIEWS * DROSFM_MAX_PREDS];
    float weight[DROSFM_MAX_PREDS];   // gamma^(n-1-i)
};
struct PhotoGrads {
    float* g_inv_depth[DROSFM_MAX_PREDS];
    float* g_pose[DROSFM_MAX_VIEWS * DROSFM_MAX_PREDS];
};

__device__ __forceinline__ int reflect_idx(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ float tap3(const float* __restrict__ plane, int Ws, const Taps& t, const Weights& w) {
    float acc = 0.0f;
    const float* r0 = plane + t.y0 * Ws + t.x0;
    if (t.valid & 1u) acc += __ldg(r0) * w.nw;
    if (t.valid & 2u) acc += __ldg(r0 + 1) * w.ne;
    if (t.valid & 4u) acc += __ldg(r0 + Ws) * w.sw;
    if (t.valid & 8u) acc += __ldg(r0 + Ws + 1) * w.se;
    return acc;
}

// SSIM statistics of one window: sums over the 9 taps
struct Win {
    float sx, sy, sxx, syy, sxy;
};

struct Ssim {
    float mu_x, mu_y, A1, A2, B1, B2, s;
};

__device__ __forceinline__ Ssim ssim_from(const Win& w, float C1, float C2) {
    // Plain fp32 with FMA contraction: the variances E[x^2]-mu^2 cancel 3-4 digits, which makes the SSIM
    // value uncertain at the 1e-4 relative level in ANY fp32 evaluation order (the reference's included),
    // so the exact rounding sequence of the ATen ops is not worth extra instructions here.
    const float inv9 = 1.0f / 9.0f;
    Ssim r;
    r.mu_x = w.sx * inv9;
    r.mu_y = w.sy * inv9;
    const float mu_xy = r.mu_x * r.mu_y, mu_xx = r.mu_x * r.mu_x, mu_yy = r.mu_y * r.mu_y;
    const float sig_x = w.sxx * inv9 - mu_xx;
    const float sig_y = w.syy * inv9 - mu_yy;
    const float sig_xy = w.sxy * inv9 - mu_xy;
    r.A1 = 2.0f * mu_xy + C1;
    r.A2 = 2.0f * sig_xy + C2;
    r.B1 = mu_xx + mu_yy + C1;
    r.B2 = sig_x + sig_y + C2;
    r.s = __fdividef(r.A1 * r.A2, r.B1 * r.B2);
    return r;
}

// horizontal 3-tap sums of one shared-memory row
__device__ __forceinline__ Win row_sums(const float* __restrict__ xr, const float* __restrict__ yr, int cm, int c0, int cp,
                                        float& xc, float& yc) {
    const float x0 = xr[cm], x1 = xr[c0], x2 = xr[cp];
    const float y0 = yr[cm], y1 = yr[c0], y2 = yr[cp];
    xc = x1;
    yc = y1;
    Win w;
    w.sx = x0 + x1 + x2;
    w.sy = y0 + y1 + y2;
    w.sxx = x0 * x0 + x1 * x1 + x2 * x2;
    w.syy = y0 * y0 + y1 * y1 + y2 * y2;
    w.sxy = x0 * y0 + x1 * y1 + x2 * y2;
    return w;
}

__device__ __forceinline__ Win add3(const Win& a, const Win& b, const Win& c) {
    Win w;
    w.sx = a.sx + b.sx + c.sx;
    w.sy = a.sy + b.sy + c.sy;
    w.sxx = a.sxx + b.sxx + c.sxx;
    w.syy = a.syy + b.syy + c.syy;
    w.sxy = a.sxy + b.sxy + c.sxy;
    return w;
}

// ------------------------------------------------------------------------------------------
// warped source tile: shared by the forward and backward kernels
// ------------------------------------------------------------------------------------------
// Fills xs[3][ROWS][COLS] with the source view sampled at every in-image pixel of the region whose
// top-left corner is (oy, ox) (zeros elsewhere).  WARP: through the depth/pose warp; otherwise the
// un-warped source (auto-mask).  Work is batched so that each thread has NB depth loads, then NB*12
// gathers in flight; the camera lives in registers.
template <int ROWS, int COLS, int NT, bool WARP>
__device__ __forceinline__ void fill_source_tile(float* __restrict__ xs, const float* __restrict__ src,
                                                 const float* __restrict__ invd, int depth_kind, const Cam& cam,
                                                 int oy, int ox, int H, int W, int P, float wm1, float hm1, int padding) {
    constexpr int N = ROWS * COLS, NB = 2;
    constexpr int STEP_Y = NT / COLS, STEP_X = NT % COLS;       // (row, column) advance of idx += NT
    const int tid = threadIdx.x;
    int ry = tid / COLS, rx = tid - ry * COLS;
    for (int base = tid; base < N; base += NT * NB) {
        int sidx[NB], sxs[NB], sys[NB];
        bool in[NB];
        float d[NB];
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            const int idx = base + k * NT;
            sidx[k] = idx;
            sys[k] = oy + ry;
            sxs[k] = ox + rx;
            in[k] = idx < N && sys[k] >= 0 && sys[k] < H && sxs[k] >= 0 && sxs[k] < W;
            d[k] = 0.0f;
            if (WARP && in[k]) d[k] = __ldg(invd + sys[k] * W + sxs[k]);
            rx += STEP_X;
            ry += STEP_Y;
            if (rx >= COLS) { rx -= COLS; ++ry; }
        }
        // taps: coordinates are clamped into the image and the weight of an out-of-bounds tap is zeroed, so
        // the gathers below need no predicates
        int o00[NB], dxo[NB], dyo[NB];
        float w00[NB], w01[NB], w10[NB], w11[NB];
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            o00[k] = 0; dxo[k] = 0; dyo[k] = 0;
            w00[k] = w01[k] = w10[k] = w11[k] = 0.0f;
            if (WARP && in[k]) {
                Warp wp;
                warp_pixel<true>(cam, sxs[k], sys[k], to_depth(d[k], depth_kind), wm1, hm1, true, wp);
                Taps t;
                make_taps(wp.p.u, wp.p.v, H, W, padding, t);
                if (t.valid) {
                    const float bx = 1.0f - t.ax, by = 1.0f - t.ay;
                    w00[k] = (t.valid & 1u) ? bx * by : 0.0f;
                    w01[k] = (t.valid & 2u) ? t.ax * by : 0.0f;
                    w10[k] = (t.valid & 4u) ? bx * t.ay : 0.0f;
                    w11[k] = (t.valid & 8u) ? t.ax * t.ay : 0.0f;
                    const int x0 = max(t.x0, 0), y0 = max(t.y0, 0);
                    const int x1 = min(t.x0 + 1, W - 1), y1 = min(t.y0 + 1, H - 1);
                    o00[k] = y0 * W + x0;
                    dxo[k] = x1 - x0;
                    dyo[k] = (y1 - y0) * W;
                }
            } else if (!WARP && in[k]) {
                o00[k] = sys[k] * W + sxs[k];
            }
        }
        float v[NB][12];
#pragma unroll
        for (int k = 0; k < NB; ++k) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float* r0 = src + c * P + o00[k];
                if (WARP) {
                    v[k][4 * c + 0] = __ldg(r0);
                    v[k][4 * c + 1] = __ldg(r0 + dxo[k]);
                    v[k][4 * c + 2] = __ldg(r0 + dyo[k]);
                    v[k][4 * c + 3] = __ldg(r0 + dyo[k] + dxo[k]);
                } else {
                    v[k][c] = in[k] ? __ldg(r0) : 0.0f;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < NB; ++k) {
            if (sidx[k] < N) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    // the evaluation order of the flat warp (warp_sources_kernel): both paths produce the same bits
                    const float o = WARP ? __fmaf_rn(v[k][4 * c + 3], w11[k], __fmaf_rn(v[k][4 * c + 2], w10[k],
                                               __fmaf_rn(v[k][4 * c + 1], w01[k], __fmul_rn(v[k][4 * c], w00[k]))))
                                         : v[k][c];
                    xs[c * N + sidx[k]] = o;
                }
            }
        }
    }
}

// Target image tile (zeros outside the image).
template <int ROWS, int COLS, int NT>
__device__ __forceinline__ void fill_target_tile(float* __restrict__ ys, const float* __restrict__ img, int oy, int ox,
                                                 int H, int W, int P) {
    constexpr int N = ROWS * COLS;
    constexpr int STEP_Y = NT / COLS, STEP_X = NT % COLS;
    int ry = threadIdx.x / COLS, rx = threadIdx.x - ry * COLS;
    for (int idx = threadIdx.x; idx < N; idx += NT) {
        const int gy = oy + ry, gx = ox + rx;
        rx += STEP_X;
        ry += STEP_Y;
        if (rx >= COLS) { rx -= COLS; ++ry; }
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        const int o = gy * W + gx;
#pragma unroll
        for (int c = 0; c < 3; ++c) ys[c * N + idx] = in ? __ldg(img + c * P + o) : 0.0f;
    }
}

// ------------------------------------------------------------------------------------------
// forward (MODE 0) and auto-mask pre-pass (MODE 1)
// ------------------------------------------------------------------------------------------
constexpr int kFwdThreads = 512, kFwdGroups = kFwdThreads / 32;
constexpr int FW = 32, FH = 64, FSW = FW + 2, FSH = FH + 2, FRPT = FH / kFwdGroups;
constexpr int kFwdSmemBytes = 2 * 3 * FSH * FSW * static_cast<int>(sizeof(float));

// clip_loss > 0 (calc_photometric_loss, multiview_photometric_loss_mf.py:220-227): every photometric map -- one per
// (prediction, view), plus the V un-warped maps of the auto-mask -- is clamped at mean + clip * std of that map.  One slot
// of the caller's scratch per map: the sums of the statistics pass, then the threshold.
struct alignas(8) ClipSlot {
    double s1, s2;       // sum and sum of squares of the map (statistics pass)
    float thr, pad_;     // min(map, thr) afterwards
    double pad2_;
};
static_assert(sizeof(ClipSlot) == 32, "ClipSlot is 8 floats of drosfm_photo_opts_t.clip_scratch");

__global__ void clip_threshold_kernel(ClipSlot* slots, int count, double n_elems, float clip) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double mean = slots[i].s1 / n_elems;
    const double var = n_elems > 1.0 ? (slots[i].s2 - n_elems * mean * mean) / (n_elems - 1.0) : 0.0;     // torch.std: unbiased
    // the reference forms mean + clip * std from float32 tensors (and then takes float())
    slots[i].thr = __fadd_rn(static_cast<float>(mean), __fmul_rn(clip, static_cast<float>(sqrt(var > 0.0 ? var : 0.0))));
}

// SAVED: the warped sources come from warp_sources_kernel's output instead of being sampled here.
// STATS: statistics pass of the clipped loss -- only the sums of every map are produced (into `clip`).
template <int MODE, bool SAVED, bool STATS = false>
__global__ void __launch_bounds__(kFwdThreads, (SAVED || MODE == 1) ? 2 : 1)
photometric_fwd_kernel(const float* __restrict__ image, const __grid_constant__ PhotoPtrs pp, int V, int depth_kind, int n_preds,
                       drosfm_cams_t cams, const float* __restrict__ automask_in, drosfm_photo_opts_t opts,
                       float l1_w, uint8_t* __restrict__ sel_out, float* __restrict__ automask_out,
                       float* __restrict__ loss, Slot* ws, const float* __restrict__ warped_save, ClipSlot* clip,
                       int B, int H, int W) {
    extern __shared__ float smem[];
    float* ys = smem;                        // [3][FSH][FSW]
    float* xs = smem + 3 * FSH * FSW;        // [3][FSH][FSW]
    __shared__ Cam cam_s[(MODE == 0 && !SAVED) ? DROSFM_MAX_VIEWS : 1];
    __shared__ double red[2 * kFwdGroups];
    __shared__ int flag;
    constexpr int PLANE = FSH * FSW;
    const int tid = threadIdx.x, lane = tid & 31, grp = tid >> 5;
    const int tx0 = blockIdx.x * FW, ty0 = blockIdx.y * FH;
    const int b = MODE == 0 ? static_cast<int>(blockIdx.z) % B : static_cast<int>(blockIdx.z);
    const int ip = MODE == 0 ? static_cast<int>(blockIdx.z) / B : 0;
    const int P = H * W;
    const float wm1 = static_cast<float>(W - 1), hm1 = static_cast<float>(H - 1);

    if (MODE == 0 && !SAVED && tid < V) setup_cam(cams, pp.pose[tid * n_preds + ip], b, cam_s[tid]);
    fill_target_tile<FSH, FSW, kFwdThreads>(ys, image + static_cast<size_t>(b) * 3 * P, ty0 - 1, tx0 - 1, H, W, P);
    __syncthreads();

    // per-thread stat rows: column gx = tx0 + lane, rows gy0 + k; shared-memory offsets are hoisted
    const int gx = tx0 + lane;
    const int gy0 = ty0 + grp * FRPT;
    const int cm = clampi(reflect_idx(gx - 1, W) - tx0 + 1, 0, FSW - 1);
    const int c0 = lane + 1;
    const int cp = clampi(reflect_idx(gx + 1, W) - tx0 + 1, 0, FSW - 1);
    int roff[FRPT + 2];
#pragma unroll
    for (int j = 0; j < FRPT + 2; ++j) roff[j] = clampi(reflect_idx(gy0 - 1 + j, H) - ty0 + 1, 0, FSH - 1) * FSW;

    float best[FRPT];
    int sel[FRPT];
    const bool use_min = opts.reduce_op == DROSFM_REDUCE_MIN;
    const bool l1_only = !(opts.ssim_w > 0.0f);
#pragma unroll
    for (int k = 0; k < FRPT; ++k) {
        best[k] = use_min ? __int_as_float(0x7f800000) : 0.0f;
        sel[k] = 254;
    }
    const float* invd = MODE == 0 ? pp.inv_depth[ip] + static_cast<size_t>(b) * P : nullptr;

    for (int v = 0; v < V; ++v) {
        // phase A: source view v on tile + halo 1
        if (SAVED) {
            const size_t slot = (static_cast<size_t>(ip) * V + v) * B + b;
            fill_target_tile<FSH, FSW, kFwdThreads>(xs, warped_save + slot * 3 * P, ty0 - 1, tx0 - 1, H, W, P);
        } else {
            const float* src = pp.context[v] + static_cast<size_t>(b) * 3 * P;
            const Cam cam = cam_s[MODE == 0 ? v : 0];       // register copy for the sampling loop
            fill_source_tile<FSH, FSW, kFwdThreads, MODE == 0>(xs, src, invd, depth_kind, cam, ty0 - 1, tx0 - 1, H, W, P, wm1, hm1,
                                                               opts.padding);
        }
        __syncthreads();

        // phase B: SSIM + L1 with a vertical sliding window
        // the map's slot in the clip scratch: the V un-warped maps first, then (prediction, view)
        ClipSlot* cslot = clip != nullptr ? clip + (MODE == 1 ? v : V + ip * V + v) : nullptr;
        const float thr = (!STATS && cslot != nullptr) ? cslot->thr : __int_as_float(0x7f800000);
        float st[2] = {0.0f, 0.0f};
        // one value of one photometric map at row k of this thread's column: statistics, clip, min / sum.  `code` is what
        // the selection records when the value wins (the view; for the per-channel L1 maps view + 8 * channel).
        auto consume = [&](int k, float pm, int code) {
            if (STATS) {
                if (gy0 + k < H && gx < W) { st[0] += pm; st[1] += pm * pm; }
                return;
            }
            const bool clipped = pm > thr;           // torch.clamp(max=thr): the value becomes thr, its gradient zero
            pm = clipped ? thr : pm;
            if (use_min) {
                if (pm < best[k]) { best[k] = pm; sel[k] = clipped ? 254 : code; }
            } else {
                best[k] += pm;
                if (!clipped) sel[k] |= 1 << v;      // 'mean': bit v = view v carries a gradient at this pixel
            }
        };
        if (!use_min && v == 0) {
#pragma unroll
            for (int k = 0; k < FRPT; ++k) sel[k] = 0;
        }
        float ssim_acc[FRPT], l1_acc[FRPT];
#pragma unroll
        for (int k = 0; k < FRPT; ++k) ssim_acc[k] = l1_acc[k] = 0.0f;
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {
            const float* xc_ = xs + c * PLANE;
            const float* yc_ = ys + c * PLANE;
            Win ra, rb;
            float xcb = 0.0f, ycb = 0.0f;
#pragma unroll
            for (int j = 0; j < FRPT + 2; ++j) {
                float xc, yc;
                const Win rc = row_sums(xc_ + roff[j], yc_ + roff[j], cm, c0, cp, xc, yc);
                if (j >= 2) {
                    if (l1_only) {
                        // ssim_loss_weight == 0: the map IS the per-channel |a - b| (three channels per view,
                        // multiview_photometric_loss_mf.py:217-218), every channel competes in the min on its own
                        consume(j - 2, fabsf(xcb - ycb), v + 8 * c);
                    } else {
                        const Ssim s = ssim_from(add3(ra, rb, rc), opts.C1, opts.C2);
                        const float l = (1.0f - s.s) * 0.5f;
                        ssim_acc[j - 2] += fminf(fmaxf(l, 0.0f), 1.0f);
                        l1_acc[j - 2] += fabsf(xcb - ycb);
                    }
                }
                ra = rb; rb = rc;
                xcb = xc; ycb = yc;
            }
        }
        if (!l1_only) {
#pragma unroll
            for (int k = 0; k < FRPT; ++k)
                consume(k, __fadd_rn(__fmul_rn(opts.ssim_w, __fdiv_rn(ssim_acc[k], 3.0f)), __fmul_rn(l1_w, __fdiv_rn(l1_acc[k], 3.0f))), v);
        }
        if (STATS) block_accumulate<2>(st, red, &cslot->s1);
        __syncthreads();
    }
    if (STATS) return;

    // epilogue
    float local = 0.0f;
#pragma unroll
    for (int k = 0; k < FRPT; ++k) {
        const int gy = gy0 + k;
        if (gy < H && gx < W) {
            const size_t o = static_cast<size_t>(b) * P + gy * W + gx;
            if (MODE == 1) {
                automask_out[o] = best[k];
            } else {
                float val = best[k];
                int s = sel[k];
                if (automask_in != nullptr) {
                    const float a = __ldg(automask_in + o);
                    if (a < val) { val = a; s = 255; }
                }
                if (sel_out != nullptr) sel_out[static_cast<size_t>(ip) * B * P + o] = static_cast<uint8_t>(s);
                local += val;
            }
        }
    }
    if (MODE == 1) return;
    __syncthreads();
    double part = warp_sum(static_cast<double>(local));
    if (lane == 0) red[grp] = part;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int k = 0; k < kFwdGroups; ++k) s += red[k];
        atomicAdd(spread_acc(slot_at(ws, ip)), s);
    }
    Slot* ticket = slot_at(ws, n_preds);
    // 'mean': every map is averaged over its own elements -- B*P of them, or B*3*P for the per-channel L1 maps
    if (last_block(ticket, gridDim.x * gridDim.y * gridDim.z, &flag) && tid < 32)
        finish_weighted_means(ws, n_preds, pp.weight,
                              static_cast<double>(B) * P * (use_min ? 1.0 : static_cast<double>(V) * (l1_only ? 3.0 : 1.0)), ticket, loss);
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
constexpr int kBwdThreads = 256, kBwdGroups = kBwdThreads / 32;
constexpr int BRPT = 5;                              // coefficient rows per thread
constexpr int CW = 32, CH = kBwdGroups * BRPT;       // coefficient region (tile + halo 1): 32 x 40
constexpr int IW = CW - 2, IH = CH - 2;              // interior pixels owned by the block: 30 x 38
constexpr int BSW = CW + 2, BSH = CH + 2;            // sample region (tile + halo 2): 34 x 42
constexpr int kBwdSmemFloats = 2 * 3 * BSH * BSW + 3 * CH * CW + 3 * IH * CW;
constexpr int kBwdSmemBytes = kBwdSmemFloats * static_cast<int>(sizeof(float)) + CH * CW;
constexpr int kBwdSavedSmemBytes = (2 * 3 * BSH * BSW + 3 * CH * CW) * static_cast<int>(sizeof(float)) + CH * CW;

// SAVED = false: the whole backward in one kernel (warp recomputed, nothing kept by the forward).
// SAVED = true : stage 3 of the staged path -- warped sources are read back from the forward's copy and the
//                gradient w.r.t. every warped pixel goes to g_warped for warp_sources_adjoint_kernel; no geometry here.
template <bool SAVED>
__global__ void __launch_bounds__(kBwdThreads, SAVED ? 3 : 2)
photometric_bwd_kernel(const float* __restrict__ g_loss, const float* __restrict__ image,
                       const __grid_constant__ PhotoPtrs pp, int V,
                       int depth_kind, int n_preds, drosfm_cams_t cams, const uint8_t* __restrict__ sel_in,
                       drosfm_photo_opts_t opts, float l1_w, const __grid_constant__ PhotoGrads pg, Slot* ws,
                       const float* __restrict__ warped_save, float* __restrict__ g_warped, int B, int H, int W) {
    extern __shared__ float smem[];
    constexpr int SPLANE = BSH * BSW, CPLANE = CH * CW;
    float* ys = smem;                                 // [3][BSH][BSW]
    float* xs = ys + 3 * SPLANE;                      // [3][BSH][BSW]
    float* ca = xs + 3 * SPLANE;                      // [CH][CW] x 3 (a, b, c coefficient maps of one channel)
    float* cb = ca + CPLANE;
    float* cc = cb + CPLANE;
    float* gxs = cc + CPLANE;                         // [3][IH][CW] gradient w.r.t. the warped pixel (fused path only)
    uint8_t* selt = reinterpret_cast<uint8_t*>(SAVED ? gxs : gxs + 3 * IH * CW);   // [CH][CW]
    __shared__ Cam cam_s[SAVED ? 1 : DROSFM_MAX_VIEWS];
    __shared__ unsigned present;                      // bit v: some window of the coefficient region selected view v
    const int tid = threadIdx.x, lane = tid & 31, grp = tid >> 5;
    const int tx0 = blockIdx.x * IW, ty0 = blockIdx.y * IH;          // interior origin
    const int cx0 = tx0 - 1, cy0 = ty0 - 1;                            // coefficient-region origin
    const int sx0 = tx0 - 2, sy0 = ty0 - 2;                            // sample-region origin
    const int b = static_cast<int>(blockIdx.z) % B, ip = static_cast<int>(blockIdx.z) / B;
    const int P = H * W;
    const float wm1 = static_cast<float>(W - 1), hm1 = static_cast<float>(H - 1);
    const bool use_min = opts.reduce_op == DROSFM_REDUCE_MIN;
    const bool mask_mode = !use_min && opts.clip_loss > 0.0f;
    const bool l1_only = !(opts.ssim_w > 0.0f);        // per-channel L1 maps: the selection is view + 8 * channel
    const float G = __ldg(g_loss) * pp.weight[ip] /
                    (static_cast<float>(B) * static_cast<float>(P) * (use_min ? 1.0f : static_cast<float>(V)));
    const float kp = G * opts.ssim_w * (-1.0f / 6.0f) * (2.0f / 9.0f);   // d loss / d ssim  x  2/9 of the window derivative
    // L1 term: l1_w * mean over the channels; a per-channel L1 map that won the min carries the whole gradient
    const float kl1 = (l1_only && use_min) ? G : G * l1_w * (1.0f / 3.0f);

    if (!SAVED && tid < V) setup_cam(cams, pp.pose[tid * n_preds + ip], b, cam_s[tid]);
    if (tid == 0) present = 0u;
    __syncthreads();
    fill_target_tile<BSH, BSW, kBwdThreads>(ys, image + static_cast<size_t>(b) * 3 * P, sy0, sx0, H, W, P);
    {
        unsigned seen = 0u;
        for (int idx = tid; idx < CH * CW; idx += kBwdThreads) {
            const int ry = idx / CW, rx = idx - ry * CW;
            const int gy = cy0 + ry, gx = cx0 + rx;
            const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
            // min: the winning view (254 none, 255 auto-mask); mean: 253 = every view; mean with a clipped loss: bit v =
            // view v was not clipped at this pixel (mask_mode)
            const int raw = in && (use_min || mask_mode) ? sel_in[(static_cast<size_t>(ip) * B + b) * P + gy * W + gx] : 0;
            const int sv = mask_mode ? (in ? raw : 0) : (in ? (use_min ? raw : 253) : 254);
            selt[idx] = static_cast<uint8_t>(sv);
            seen |= mask_mode ? static_cast<unsigned>(sv) : (sv == 253 ? 0xffffffffu : (sv < 32 ? 1u << (sv & 7) : 0u));
        }
        seen = __reduce_or_sync(0xffffffffu, seen);
        if (lane == 0 && seen) atomicOr(&present, seen);
    }
    __syncthreads();

    // this thread: coefficient column lane (image column pgx), coefficient rows prow0 .. prow0+BRPT-1
    const int pgx = cx0 + lane;
    const int prow0 = grp * BRPT;
    const int cm = clampi(reflect_idx(pgx - 1, W) - sx0, 0, BSW - 1);
    const int c0 = lane + 1;
    const int cp = clampi(reflect_idx(pgx + 1, W) - sx0, 0, BSW - 1);
    int roff[BRPT + 2];
#pragma unroll
    for (int j = 0; j < BRPT + 2; ++j) roff[j] = clampi(reflect_idx(cy0 + prow0 - 1 + j, H) - sy0, 0, BSH - 1) * BSW;
    // interior pixels of this thread: column lane in [1, IW], rows r = prow0 + k in [1, IH]
    const bool col_ok = lane >= 1 && lane <= IW && pgx < W;
    // horizontal multiplicities of the reflected 3x3 window: how often pixel pgx occurs in the window of pgx+dx
    float wx[3];
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
        const int px = pgx + dx;
        wx[dx + 1] = (px < 0 || px >= W) ? 0.0f
                                         : 1.0f + ((px == 0 && pgx == 1) ? 1.0f : 0.0f) + ((px == W - 1 && pgx == W - 2) ? 1.0f : 0.0f);
    }
    const int lm = max(lane - 1, 0), lp = min(lane + 1, CW - 1);
    float gd[BRPT];
#pragma unroll
    for (int k = 0; k < BRPT; ++k) gd[k] = 0.0f;
    const float* invd = pp.inv_depth[ip] + static_cast<size_t>(b) * P;

    for (int v = 0; v < V; ++v) {
        const float* src = pp.context[v] + static_cast<size_t>(b) * 3 * P;
        const Cam cam = cam_s[SAVED ? 0 : v];
        const size_t vslot = (static_cast<size_t>(ip) * V + v) * B + b;
        float* gw = SAVED ? g_warped + vslot * 3 * P : nullptr;
        if (!((present >> v) & 1u)) {
            // no window of this block chose view v: its gradient is identically zero here
            if (SAVED) {
#pragma unroll
                for (int k = 0; k < BRPT; ++k) {
                    const int r = prow0 + k, qy = cy0 + r;
                    if (col_ok && r >= 1 && r <= IH && qy < H) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) gw[c * P + qy * W + pgx] = 0.0f;
                    }
                }
            }
            continue;
        }
        // phase A: warped source on tile + halo 2 -- reloaded from the forward's copy, or recomputed
        if (SAVED) fill_target_tile<BSH, BSW, kBwdThreads>(xs, warped_save + vslot * 3 * P, sy0, sx0, H, W, P);
        else fill_source_tile<BSH, BSW, kBwdThreads, true>(xs, src, invd, depth_kind, cam, sy0, sx0, H, W, P, wm1, hm1, opts.padding);
        __syncthreads();

#pragma unroll 1
        for (int c = 0; c < 3; ++c) {
            // phase B: coefficient maps of channel c on tile + halo 1 (zero where another map won the min)
            {
                const float* xc_ = xs + c * SPLANE;
                const float* yc_ = ys + c * SPLANE;
                Win ra, rb;
#pragma unroll
                for (int j = 0; j < BRPT + 2; ++j) {
                    float xc, yc;
                    const Win rc = row_sums(xc_ + roff[j], yc_ + roff[j], cm, c0, cp, xc, yc);
                    if (j >= 2) {
                        const int r = prow0 + j - 2;
                        float a = 0.0f, bb = 0.0f, cq = 0.0f;
                        const int sv = selt[r * CW + lane];
                        if (mask_mode ? ((sv >> v) & 1) != 0 : (sv == v || sv == 253)) {
                            const Ssim s = ssim_from(add3(ra, rb, rc), opts.C1, opts.C2);
                            const float l = (1.0f - s.s) * 0.5f;
                            if (l >= 0.0f && l <= 1.0f) {
                                const float q = __fdividef(kp, s.B1 * s.B2);
                                a = q * (s.mu_y * (s.A2 - s.A1) - s.s * s.mu_x * (s.B2 - s.B1));
                                bb = -q * s.s * s.B1;
                                cq = q * s.A1;
                            }
                        }
                        ca[r * CW + lane] = a;
                        cb[r * CW + lane] = bb;
                        cc[r * CW + lane] = cq;
                    }
                    ra = rb; rb = rc;
                }
            }
            __syncthreads();
            // phase C: 3x3 box sums (with the reflection multiplicities) of the three maps -> d loss / d warped pixel
            {
                float ha[BRPT + 2], hb[BRPT + 2], hc[BRPT + 2];
#pragma unroll
                for (int j = 0; j < BRPT + 2; ++j) {
                    const int row = clampi(prow0 - 1 + j, 0, CH - 1) * CW;
                    ha[j] = wx[0] * ca[row + lm] + wx[1] * ca[row + lane] + wx[2] * ca[row + lp];
                    hb[j] = wx[0] * cb[row + lm] + wx[1] * cb[row + lane] + wx[2] * cb[row + lp];
                    hc[j] = wx[0] * cc[row + lm] + wx[1] * cc[row + lane] + wx[2] * cc[row + lp];
                }
#pragma unroll
                for (int k = 0; k < BRPT; ++k) {
                    const int r = prow0 + k;
                    const int qy = cy0 + r;
                    if (col_ok && r >= 1 && r <= IH && qy < H) {
                        float ga = 0.0f, gb = 0.0f, gc = 0.0f;
#pragma unroll
                        for (int dy = -1; dy <= 1; ++dy) {
                            const int py = qy + dy;
                            const float wy = (py < 0 || py >= H) ? 0.0f
                                                                 : 1.0f + ((py == 0 && qy == 1) ? 1.0f : 0.0f) +
                                                                       ((py == H - 1 && qy == H - 2) ? 1.0f : 0.0f);
                            ga += wy * ha[k + 1 + dy];
                            gb += wy * hb[k + 1 + dy];
                            gc += wy * hc[k + 1 + dy];
                        }
                        const float xq = xs[c * SPLANE + (r + 1) * BSW + lane + 1], yq = ys[c * SPLANE + (r + 1) * BSW + lane + 1];
                        float gxv = ga + gb * xq + gc * yq;
                        const int sv = selt[r * CW + lane];
                        if (mask_mode ? ((sv >> v) & 1) != 0 : (sv == (l1_only ? v + 8 * c : v) || sv == 253)) {
                            const float df = xq - yq;
                            gxv += df > 0.0f ? kl1 : (df < 0.0f ? -kl1 : 0.0f);
                        }
                        if (SAVED) gw[c * P + qy * W + pgx] = gxv;
                        else gxs[(c * IH + r - 1) * CW + lane] = gxv;
                    }
                }
            }
            __syncthreads();
        }

        if (SAVED) continue;          // stages B and C end with a barrier; the adjoint runs in its own kernel

        // phase D: through the bilinear taps and the projection adjoint
        float gT[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) gT[i] = 0.0f;
#pragma unroll
        for (int k = 0; k < BRPT; ++k) {
            const int r = prow0 + k;
            const int qy = cy0 + r;
            if (col_ok && r >= 1 && r <= IH && qy < H) {
                const float d = to_depth(__ldg(invd + qy * W + pgx), depth_kind);
                Warp wp;
                warp_pixel<true>(cam, pgx, qy, d, wm1, hm1, true, wp);
                Taps t;
                make_taps(wp.p.u, wp.p.v, H, W, opts.padding, t);
                if (t.valid) {
                    const int x0 = max(t.x0, 0), y0 = max(t.y0, 0);
                    const int o00 = y0 * W + x0, dxo = min(t.x0 + 1, W - 1) - x0, dyo = (min(t.y0 + 1, H - 1) - y0) * W;
                    const float m0 = (t.valid & 1u) ? 1.f : 0.f, m1 = (t.valid & 2u) ? 1.f : 0.f;
                    const float m2 = (t.valid & 4u) ? 1.f : 0.f, m3 = (t.valid & 8u) ? 1.f : 0.f;
                    float gix = 0.0f, giy = 0.0f;
                    const float bx = 1.0f - t.ax, by = 1.0f - t.ay;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float* r0 = src + c * P + o00;
                        const float v0 = __ldg(r0) * m0, v1 = __ldg(r0 + dxo) * m1;
                        const float v2 = __ldg(r0 + dyo) * m2, v3 = __ldg(r0 + dyo + dxo) * m3;
                        const float g = gxs[(c * IH + r - 1) * CW + lane];
                        gix += g * ((v1 - v0) * by + (v3 - v2) * t.ay);
                        giy += g * ((v2 - v0) * bx + (v3 - v1) * t.ax);
                    }
                    gd[k] += warp_pixel_adjoint(cam, wp, d, wm1, hm1, true, gix * t.mx, giy * t.my, gT);
                }
            }
        }
        if (pg.g_pose[v * n_preds + ip] != nullptr)
            warp_accumulate12(gT, spread_acc(slot_at(ws, (v * n_preds + ip) * B + b)));
        __syncthreads();      // gxs / xs are rewritten by the next view
    }
    if (SAVED) return;

    float* gout = pg.g_inv_depth[ip];
    if (gout != nullptr) {
#pragma unroll
        for (int k = 0; k < BRPT; ++k) {
            const int r = prow0 + k;
            const int qy = cy0 + r;
            if (col_ok && r >= 1 && r <= IH && qy < H) {
                const int o = qy * W + pgx;
                float g = gd[k];
                if (depth_kind == DROSFM_INV_DEPTH) g = inv2depth_grad(__ldg(invd + o), g);
                gout[static_cast<size_t>(b) * P + o] = g;
            }
        }
    }
    // one ticket per (view, prediction, sample): the last block turns the fp64 sums into the caller's encoding
    __shared__ unsigned last_mask;
    finish_last_slots(
        V, static_cast<unsigned long long>(gridDim.x) * gridDim.y, &last_mask,
        [&](int v) { return pg.g_pose[v * n_preds + ip] != nullptr ? slot_at(ws, (v * n_preds + ip) * B + b) : nullptr; },
        [&](int v, Slot* slot) {
            const bool eul = cams.pose_kind == DROSFM_POSE_EULER6;
            finish_pose_grad_warp(slot, cams.pose_kind, eul ? pp.pose[v * n_preds + ip] + b * 6 : nullptr,
                                  pg.g_pose[v * n_preds + ip] + b * (eul ? 6 : 16));
        });
}

// ------------------------------------------------------------------------------------------
// staged path: stage 1 (warp every source view once, no halo) and stage 4 (its adjoint)
// ------------------------------------------------------------------------------------------
// With a warped_save buffer the loss runs as four lean kernels instead of two fused ones:
//   1 warp_sources_kernel          inv_depth, pose, source  -> warped [n,V,B,3,H,W]
//   2 photometric_fwd_kernel<0,1>  warped, target           -> loss, sel
//   3 photometric_bwd_kernel<1>    warped, target, sel      -> g_warped [n,V,B,3,H,W]
//   4 warp_sources_adjoint_kernel  g_warped, inv_depth, ... -> g_inv_depth, g_pose
// The SSIM stages then carry no camera state (fewer registers, more resident warps) and the geometry runs
// once per pixel instead of once per pixel of every tile + halo.
constexpr int kFlatThreads = 256, kFlatRows = 8;
constexpr int kFlatTileH = (kFlatThreads / 32) * kFlatRows;       // one tile: 32 columns x 64 rows
// Tuning (B200, 320x960, B=2, n=9, V=2; loss graph replay, scratch/loss_bench.py):
//   forward : 4 blocks/SM (64 registers) 130 us, 3 blocks 142 us, 2 blocks 152 us; tiles per block 1 / 2 / 5: 136 / 142 / 130 us
//   adjoint : 2 blocks/SM (128 registers, no spills) 202 us, 3 blocks (80 registers, 120 B spilled) 219 us;
//             tiles per block 1 / 2 / 3 / 5 (final build, one barrier set-up): 215 / 201 / 191 / 189 us;
//             row loop unrolled: 317 us (instruction cache)
#ifndef DROSFM_FWD_TILES
#define DROSFM_FWD_TILES 1
#endif
// (DROSFM_ADJ_TILES, when defined, pins the adjoint's tiles per block; by default launch_adjoint chooses per launch)
#ifndef DROSFM_FWD_MINBLOCKS
#define DROSFM_FWD_MINBLOCKS 4
#endif
#ifndef DROSFM_ADJ_MINBLOCKS
#define DROSFM_ADJ_MINBLOCKS 2
#endif
#ifndef DROSFM_ADJ_UNROLL
#define DROSFM_ADJ_UNROLL 1
#endif
#ifndef DROSFM_ADJ_ILP
#define DROSFM_ADJ_ILP 1      // rows interleaved per iteration: 1 -> 203 us, 2 -> 205 us, 4 -> 323 us (registers)
#endif
constexpr int kFwdTiles = DROSFM_FWD_TILES;   // tiles a block walks down (amortises its camera set-up)
constexpr int kAdjTilesMax = 8;               // the adjoint picks its tile count per launch (launch_adjoint)

// Source pictures as RGBx texels ([V][B][H][W][4] floats, x = 0): every bilinear tap of the flat warp and of its adjoint
// is ONE 128-bit gather instead of three 32-bit ones from three planes (a third of the load instructions and address
// arithmetic; the two horizontal taps of a pixel share a 32-byte sector).  Written once per step.
__global__ void __launch_bounds__(256)
pack_rgbx_kernel(const __grid_constant__ PhotoPtrs pp, int B, int P, float* __restrict__ rgbx) {
    const int v = blockIdx.z, b = blockIdx.y;
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= P) return;
    const float* __restrict__ src = pp.context[v] + static_cast<size_t>(b) * 3 * P;
    const float4 t = make_float4(__ldg(src + p), __ldg(src + P + p), __ldg(src + 2 * P + p), 0.0f);
    reinterpret_cast<float4*>(rgbx)[(static_cast<size_t>(v) * B + b) * P + p] = t;
}

struct FlatTap {
    int o00, dxo, dyo;                 // element offsets (pixels): north-west tap, +east, +south
    float w00, w01, w10, w11;
};

__device__ __forceinline__ FlatTap flat_taps(float u, float v, int H, int W, int padding) {
    float mx, my;
    const ATap tx = axis_taps(unnormalize(u, W, padding, mx), W);
    const ATap ty = axis_taps(unnormalize(v, H, padding, my), H);
    FlatTap tp;
    tp.o00 = ty.i0 * W + tx.i0;
    tp.dxo = tx.step;
    tp.dyo = ty.step * W;
    tp.w00 = tx.w0 * ty.w0; tp.w01 = tx.w1 * ty.w0; tp.w10 = tx.w0 * ty.w1; tp.w11 = tx.w1 * ty.w1;
    return tp;
}

// PACKED: the sources are read as RGBx texels (`rgbx`), otherwise as the caller's three planes.
template <bool PACKED>
__global__ void __launch_bounds__(kFlatThreads, DROSFM_FWD_MINBLOCKS)
warp_sources_kernel(const __grid_constant__ PhotoPtrs pp, int V, int depth_kind, int n_preds, drosfm_cams_t cams, int padding,
                    const float* __restrict__ rgbx, float* __restrict__ warped, int B, int H, int W) {
    __shared__ Cam cam_s;
    const int slot = blockIdx.z;                               // (ip * V + v) * B + b
    const int b = slot % B, v = (slot / B) % V, ip = slot / (B * V);
    setup_cam_split(cams, pp.pose[v * n_preds + ip], b, cam_s);
    const int P = H * W;
    const Norm nm = make_norm(W, H);
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const float* __restrict__ invd = pp.inv_depth[ip] + static_cast<size_t>(b) * P;
    const float* __restrict__ src = pp.context[v] + static_cast<size_t>(b) * 3 * P;
    const float4* __restrict__ tex = reinterpret_cast<const float4*>(rgbx) + (static_cast<size_t>(v) * B + b) * P;
    float* __restrict__ out = warped + static_cast<size_t>(slot) * 3 * P;
    constexpr int kStride = kFlatThreads / 32;
    // the depths of a tile are all requested up front -- those of the first tile before the barrier of the camera set-up,
    // whose latency they share
    float d[kFlatRows];
    auto load_depths = [&](int tile) {
        const int y0 = (blockIdx.y * kFwdTiles + tile) * kFlatTileH + (threadIdx.x >> 5);
#pragma unroll
        for (int k = 0; k < kFlatRows; ++k) {
            const int y = y0 + k * kStride;
            d[k] = (x < W && y < H) ? __ldg(invd + y * W + x) : 0.0f;
        }
    };
    load_depths(0);
    __syncthreads();
    const Cam cam = cam_s;
    if (x >= W) return;
    // a block walks down kFwdTiles tiles of 64 rows: the camera set-up (three threads, a barrier) is paid once for all of them
#pragma unroll 1
    for (int tile = 0; tile < kFwdTiles; ++tile) {
        const int y0 = (blockIdx.y * kFwdTiles + tile) * kFlatTileH + (threadIdx.x >> 5);
        if (y0 - static_cast<int>(threadIdx.x >> 5) >= H) break;
        // Software pipeline over the thread's rows: the gathers of row k are in flight while the coordinate chain of
        // row k+1 runs.
        if (tile > 0) load_depths(tile);
        auto taps_of = [&](int k) {
            // rows beyond the image run the chain on depth 0 (in-range addresses, results discarded): no divergence
            Warp wp;
            warp_pixel_fast(cam, x, y0 + k * kStride, to_depth_fast(d[k], depth_kind), nm, true, wp);
            return flat_taps(wp.p.u, wp.p.v, H, W, padding);
        };
        float val[12];
        float4 tex4[4];
        auto gather = [&](const FlatTap& tp) {
            if constexpr (PACKED) {
                const float4* r0 = tex + tp.o00;
                tex4[0] = __ldg(r0);
                tex4[1] = __ldg(r0 + tp.dxo);
                tex4[2] = __ldg(r0 + tp.dyo);
                tex4[3] = __ldg(r0 + tp.dyo + tp.dxo);
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float* r0 = src + (static_cast<unsigned>(c * P) + static_cast<unsigned>(tp.o00));
                    val[4 * c + 0] = __ldg(r0);
                    val[4 * c + 1] = __ldg(r0 + tp.dxo);
                    val[4 * c + 2] = __ldg(r0 + tp.dyo);
                    val[4 * c + 3] = __ldg(r0 + tp.dyo + tp.dxo);
                }
            }
        };
        FlatTap cur = taps_of(0);
        gather(cur);
#pragma unroll
        for (int k = 0; k < kFlatRows; ++k) {
            FlatTap nxt = cur;
            if (k + 1 < kFlatRows) nxt = taps_of(k + 1);
            const int y = y0 + k * kStride;
            if (y < H) {
                const unsigned o = static_cast<unsigned>(y * W + x);
                // one fixed evaluation order for both tap layouts (the compiler would otherwise contract them differently)
                auto blend = [&](float a0, float a1, float a2, float a3) {
                    return __fmaf_rn(a3, cur.w11, __fmaf_rn(a2, cur.w10, __fmaf_rn(a1, cur.w01, __fmul_rn(a0, cur.w00))));
                };
                if constexpr (PACKED) {
                    out[o] = blend(tex4[0].x, tex4[1].x, tex4[2].x, tex4[3].x);
                    out[o + P] = blend(tex4[0].y, tex4[1].y, tex4[2].y, tex4[3].y);
                    out[o + 2 * P] = blend(tex4[0].z, tex4[1].z, tex4[2].z, tex4[3].z);
                } else {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        out[o + static_cast<unsigned>(c * P)] = blend(val[4 * c], val[4 * c + 1], val[4 * c + 2], val[4 * c + 3]);
                }
            }
            if (k + 1 < kFlatRows) {
                cur = nxt;
                gather(cur);
            }
        }
    }
}

// red.global.add.f32: fire-and-forget accumulation of one view's depth gradient
__device__ __forceinline__ void red_add1(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }

// Adjoint of the flat warp.  Like the forward, one block owns a 32 x 64 tile of ONE (prediction, view, sample): a thread
// walks down 8 rows of its column, pushes the upstream gradient of the warped pixel through the bilinear taps and the
// projection chain, and keeps the 12 pose-gradient terms in registers over all its rows -- one warp-level
// reduce-scatter (13 shuffles) and 12 fp64 atomics per warp and block.  The depth gradients of the V views of a pixel
// meet in g_inv_depth through red.global.add (ADD) or are stored (V == 1, nothing to add to).
// A row is skipped when no lane of the warp has a non-zero upstream gradient (un-selected view / auto-masked region).
template <bool PACKED, bool ADD>
__global__ void __launch_bounds__(kFlatThreads, DROSFM_ADJ_MINBLOCKS)
warp_sources_adjoint_kernel(const __grid_constant__ PhotoPtrs pp, int V, int depth_kind, int n_preds, drosfm_cams_t cams,
                            int padding, const float* __restrict__ rgbx, const float* __restrict__ g_warped,
                            const float* __restrict__ g_scale, const __grid_constant__ PhotoGrads pg, Slot* ws, int tiles, int B,
                            int H, int W) {
    __shared__ Cam cam_s;
    __shared__ int flag;
    const int tid = threadIdx.x;
    const int slot = blockIdx.z;                               // (ip * V + v) * B + b
    const int b = slot % B, v = (slot / B) % V, ip = slot / (B * V);
    setup_cam_split(cams, pp.pose[v * n_preds + ip], b, cam_s);      // the barrier follows the first loads (below)
    const Cam& cam = cam_s;                                    // read from shared memory on use: the registers go to occupancy
    const int P = H * W;
    const Norm nm = make_norm(W, H);
    // g_warped may come unscaled from the training forward (ssim_train_stream2_kernel).  Everything below is linear in it, so
    // the loss's upstream gradient multiplies the OUTPUTS (depth gradient per row, pose sums once); it is re-read where it is
    // used instead of being carried in a register through the row loop (the kernel sits at its register cap)
    auto load_scale = [&]() {
        float v = 1.0f;
        if (g_scale != nullptr) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(g_scale));
        return v;
    };
    const int x = blockIdx.x * 32 + (tid & 31);
    const int row0 = blockIdx.y * tiles * kFlatTileH + (tid >> 5);           // first row of this thread; then every 8th
    const bool col_ok = x < W;
    constexpr int kStride = kFlatThreads / 32;
    const int kRows = tiles * kFlatRows;       // `tiles` tiles of 64 rows per block (host: as many as keep the grid a few waves deep)
    const float* __restrict__ invd = pp.inv_depth[ip] + static_cast<size_t>(b) * P;
    float* __restrict__ gout = pg.g_inv_depth[ip] != nullptr ? pg.g_inv_depth[ip] + static_cast<size_t>(b) * P : nullptr;
    const float* __restrict__ gw = g_warped + static_cast<size_t>(slot) * 3 * P;
    const float* __restrict__ src = pp.context[v] + static_cast<size_t>(b) * 3 * P;
    const float4* __restrict__ tex = reinterpret_cast<const float4*>(rgbx) + (static_cast<size_t>(v) * B + b) * P;
    float gT[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) gT[i] = 0.0f;
    // The upstream gradients and the depths of the next kIlp rows are requested while the current ones are processed; the
    // row loop is NOT unrolled beyond that (the body is ~400 instructions per row: eight copies thrash the instruction
    // cache).  kIlp rows are processed side by side in one basic block so that their chains and gathers overlap.
    constexpr int kIlp = DROSFM_ADJ_ILP;
    static_assert(kFlatRows % kIlp == 0, "rows per tile must be a multiple of the interleave");
    float gn[kIlp][3], dn[kIlp];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int r = 0; r < kIlp; ++r) {
            const int y = row0 + (k0 + r) * kStride;
            const bool in = col_ok && y < H;
            const unsigned o = in ? static_cast<unsigned>(y * W + x) : 0u;
#pragma unroll
            for (int c = 0; c < 3; ++c) gn[r][c] = in ? __ldg(gw + (o + static_cast<unsigned>(c * P))) : 0.0f;      // consumed one row later
            dn[r] = in ? __ldg(invd + o) : 0.0f;
        }
    };
    fetch(0);
    __syncthreads();                                           // camera set-up done (its latency shared with the loads above)
    constexpr int kAdjUnroll = DROSFM_ADJ_UNROLL;
#pragma unroll kAdjUnroll
    for (int k = 0; k < kRows; k += kIlp) {
        if (row0 + k * kStride - static_cast<int>(tid >> 5) >= H) break;        // block-uniform: rows beyond the image
        float g[kIlp][3], draw[kIlp];
        bool any = false;
#pragma unroll
        for (int r = 0; r < kIlp; ++r) {
            g[r][0] = gn[r][0]; g[r][1] = gn[r][1]; g[r][2] = gn[r][2];
            draw[r] = dn[r];
            any |= g[r][0] != 0.0f || g[r][1] != 0.0f || g[r][2] != 0.0f;
        }
        if (k + kIlp < kRows) fetch(k + kIlp);
        float gd[kIlp];
#pragma unroll
        for (int r = 0; r < kIlp; ++r) gd[r] = 0.0f;
        if (__any_sync(0xffffffffu, any)) {
            Warp wp[kIlp];
            float d[kIlp], mx[kIlp], my[kIlp], gix[kIlp], giy[kIlp];
#pragma unroll
            for (int r = 0; r < kIlp; ++r) {
                const int y = row0 + (k + r) * kStride;
                d[r] = to_depth_fast(draw[r], depth_kind);
                warp_pixel_fast(cam, x, y, d[r], nm, true, wp[r]);
                const ATap tx = axis_taps(unnormalize(wp[r].p.u, W, padding, mx[r]), W);
                const ATap ty = axis_taps(unnormalize(wp[r].p.v, H, padding, my[r]), H);
                const unsigned o00 = static_cast<unsigned>(ty.i0 * W + tx.i0);
                const unsigned dxo = static_cast<unsigned>(tx.step), dyo = static_cast<unsigned>(ty.step * W);
                // d(sample)/d(ix) = (ne - nw) wy0 + (se - sw) wy1 and d(sample)/d(iy) = (sw - nw) wx0 + (se - ne) wx1, a tap
                // outside the source counting as zero: four coefficients per direction, shared by the channels
                const float kx0 = -tx.f0 * ty.w0, kx1 = tx.f1 * ty.w0, kx2 = -tx.f0 * ty.w1, kx3 = tx.f1 * ty.w1;
                const float ky0 = -ty.f0 * tx.w0, ky1 = -ty.f0 * tx.w1, ky2 = ty.f1 * tx.w0, ky3 = ty.f1 * tx.w1;
                const float g0 = g[r][0], g1 = g[r][1], g2 = g[r][2];
                if constexpr (PACKED) {
                    const float4 t0 = __ldg(tex + o00), t1 = __ldg(tex + (o00 + dxo));
                    const float4 t2 = __ldg(tex + (o00 + dyo)), t3 = __ldg(tex + (o00 + dyo + dxo));
                    gix[r] = g0 * (t0.x * kx0 + t1.x * kx1 + t2.x * kx2 + t3.x * kx3) + g1 * (t0.y * kx0 + t1.y * kx1 + t2.y * kx2 + t3.y * kx3) +
                             g2 * (t0.z * kx0 + t1.z * kx1 + t2.z * kx2 + t3.z * kx3);
                    giy[r] = g0 * (t0.x * ky0 + t1.x * ky1 + t2.x * ky2 + t3.x * ky3) + g1 * (t0.y * ky0 + t1.y * ky1 + t2.y * ky2 + t3.y * ky3) +
                             g2 * (t0.z * ky0 + t1.z * ky1 + t2.z * ky2 + t3.z * ky3);
                } else {
                    gix[r] = giy[r] = 0.0f;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const unsigned oc = o00 + static_cast<unsigned>(c * P);
                        const float v0 = __ldg(src + oc), v1 = __ldg(src + (oc + dxo));
                        const float v2 = __ldg(src + (oc + dyo)), v3 = __ldg(src + (oc + dyo + dxo));
                        gix[r] += g[r][c] * (v0 * kx0 + v1 * kx1 + v2 * kx2 + v3 * kx3);
                        giy[r] += g[r][c] * (v0 * ky0 + v1 * ky1 + v2 * ky2 + v3 * ky3);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < kIlp; ++r) {
                const bool in = col_ok && row0 + (k + r) * kStride < H;
                // lanes without a gradient (or outside the image) carry gix = giy = 0: every term is then an exact zero
                if (in && (gix[r] != 0.0f || giy[r] != 0.0f))
                    gd[r] = warp_pixel_adjoint(cam, wp[r], d[r], nm.wm1, nm.hm1, true, gix[r] * mx[r], giy[r] * my[r], gT);
            }
        }
#pragma unroll
        for (int r = 0; r < kIlp; ++r) {
            const int y = row0 + (k + r) * kStride;
            if (gout != nullptr && col_ok && y < H) {
                const float gg = (depth_kind == DROSFM_INV_DEPTH ? inv2depth_grad(draw[r], gd[r]) : gd[r]) * load_scale();
                float* dst = gout + static_cast<unsigned>(y * W + x);
                if constexpr (ADD) { if (gg != 0.0f) red_add1(dst, gg); }
                else *dst = gg;
            }
        }
    }
    float* g_pose = pg.g_pose[v * n_preds + ip];
    if (g_pose == nullptr) return;
    Slot* slot_p = slot_at(ws, (v * n_preds + ip) * B + b);
    {
        const float gs = load_scale();
#pragma unroll
        for (int i = 0; i < 12; ++i) gT[i] *= gs;
    }
    warp_accumulate12(gT, spread_acc(slot_p));
    // one ticket per (view, prediction, sample): the last block turns the fp64 sums into the caller's encoding
    if (last_block(slot_p, gridDim.x * gridDim.y, &flag) && tid < 32) {
        const bool eul = cams.pose_kind == DROSFM_POSE_EULER6;
        finish_pose_grad_warp(slot_p, cams.pose_kind, eul ? pp.pose[v * n_preds + ip] + b * 6 : nullptr, g_pose + b * (eul ? 6 : 16));
    }
}

// zero-fill of the depth-gradient maps the adjoint accumulates into (several views, nothing written before)
__global__ void __launch_bounds__(256) zero_inv_grads_kernel(const __grid_constant__ PhotoGrads pg, size_t n) {
    float* dst = pg.g_inv_depth[blockIdx.y];
    if (dst == nullptr) return;
    for (size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * 256) dst[i] = 0.0f;
}

// host side of the adjoint launch: accumulate == 0 -> the depth gradients are (over)written, else added to
static int launch_adjoint(const PhotoPtrs& pp, const PhotoGrads& pg, int n_views, int depth_kind, int n_preds, const drosfm_cams_t* cams,
                          int padding, const float* rgbx, const float* g_warped, const float* g_scale, Slot* ws, int accumulate,
                          int B, int H, int W, cudaStream_t cs) {
    // tiles of 64 rows per block: the pose-gradient reduction, its atomics and the camera set-up are paid once per block, so
    // as many as the picture has (KITTI 320 rows: 5 -> 189 us, 3 -> 191, 2 -> 201, 1 -> 215) while the grid stays >= 3 waves
    const int slots = B * n_preds * n_views, cols = (W + 31) / 32;
    int tiles = (H + kFlatTileH - 1) / kFlatTileH;
    if (tiles > kAdjTilesMax) tiles = kAdjTilesMax;
#ifdef DROSFM_ADJ_TILES
    tiles = DROSFM_ADJ_TILES;
#else
    while (tiles > 1 && static_cast<long long>(cols) * ((H + tiles * kFlatTileH - 1) / (tiles * kFlatTileH)) * slots <
                            3ll * kNumSMs * DROSFM_ADJ_MINBLOCKS)
        --tiles;
#endif
    dim3 flat(cols, (H + tiles * kFlatTileH - 1) / (tiles * kFlatTileH), slots);
    const bool add = accumulate != 0 || n_views > 1;
    if (add && accumulate == 0) {
        const size_t n = static_cast<size_t>(B) * H * W;
        dim3 zgrid(static_cast<unsigned>(min(static_cast<size_t>(kNumSMs) * 4, (n + 255) / 256)), n_preds);
        zero_inv_grads_kernel<<<zgrid, 256, 0, cs>>>(pg, n);
        if (int e = launch_status("warp adjoint (zero fill)")) return e;
    }
#define ADJ(PK, AD) warp_sources_adjoint_kernel<PK, AD><<<flat, kFlatThreads, 0, cs>>>(pp, n_views, depth_kind, n_preds, *cams, padding, rgbx, \
                                                                                     g_warped, g_scale, pg, ws, tiles, B, H, W)
    if (rgbx != nullptr) { if (add) ADJ(true, true); else ADJ(true, false); }
    else { if (add) ADJ(false, true); else ADJ(false, false); }
#undef ADJ
    return DROSFM_OK;
}

// ------------------------------------------------------------------------------------------
// staged path, stages 2 and 3 as streaming kernels: no shared-memory tiles, no block barriers
// ------------------------------------------------------------------------------------------
// A warp owns a strip of 32 columns (outer 1 or 2 lanes are halo) and walks down a band of rows.  Horizontal
// neighbours come from warp shuffles, the vertical window from rotating register rows, so every value is loaded
// once per strip (coalesced 128-byte rows) and the 3x3 statistics cost 2 shuffles + 2 adds per sum.  Reflection
// padding: the row index is remapped before the load; at the image's first / last column the missing neighbour
// is the opposite one.
constexpr int kSsimThreads = 256, kSsimWarps = kSsimThreads / 32;
constexpr int kFwdStripW = 30, kFwdBandH = 16;       // forward : 1-lane halo, 18 rows loaded for 16 written
constexpr int kBwdStripW = 28, kBwdBandH = 32;       // backward: 2-lane halo, 36 rows loaded for 32 written

// Left / right neighbour lanes with the reflection at the image's first / last column folded into the source lane.
struct Lanes {
    int l, r;
};
__device__ __forceinline__ Lanes neighbour_lanes(int lane, bool left_edge, bool right_edge) {
    Lanes n;
    n.l = left_edge ? lane + 1 : lane - 1;
    n.r = right_edge ? lane - 1 : lane + 1;
    return n;
}
__device__ __forceinline__ void neighbours(float c, const Lanes& n, float& l, float& r) {
    l = __shfl_sync(0xffffffffu, c, n.l);
    r = __shfl_sync(0xffffffffu, c, n.r);
}
// Row gy of the reflection-padded image (one row of padding), clamped for rows beyond the padding.
__device__ __forceinline__ int padded_row(int gy, int H) { return max(min(abs(gy), 2 * H - 2 - gy), 0); }
// x / 3 correctly rounded (checked against IEEE division for every float in [0, 8]): the per-pixel channel mean.
__device__ __forceinline__ float third(float x) {
    const float r = 1.0f / 3.0f;
    const float q = __fmul_rn(x, r);
    return __fmaf_rn(__fmaf_rn(-3.0f, q, x), r, q);
}

template <int NV>
struct FwdSums {
    float sy[3], syy[3], sx[NV][3], sxx[NV][3], sxy[NV][3];
};
template <int NV>
struct FwdRow {
    FwdSums<NV> s;
    float yc[3], xc[NV][3];
};

template <int NV>
__global__ void __launch_bounds__(kSsimThreads, 2)
ssim_fwd_stream_kernel(const float* __restrict__ image, const float* __restrict__ warped, const __grid_constant__ PhotoPtrs pp,
                       int n_preds, const float* __restrict__ automask_in, drosfm_photo_opts_t opts, float l1_w,
                       uint8_t* __restrict__ sel_out, float* __restrict__ loss, Slot* ws, int B, int H, int W, int nstrips,
                       int nbands) {
    __shared__ double red[kSsimWarps];
    __shared__ int flag;
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    const int wg = blockIdx.x * kSsimWarps + wib;
    const bool active = wg < nstrips * nbands;
    const int strip = active ? wg % nstrips : 0, band = active ? wg / nstrips : 0;
    const int b = static_cast<int>(blockIdx.y) % B, ip = static_cast<int>(blockIdx.y) / B;
    const int P = H * W;
    const int gx = strip * kFwdStripW - 1 + lane, gy0 = band * kFwdBandH;
    const bool col_in = gx >= 0 && gx < W;
    const bool out_lane = active && lane >= 1 && lane <= kFwdStripW && gx < W;
    const Lanes nb = neighbour_lanes(lane, gx == 0, gx == W - 1);
    const int gxc = clampi(gx, 0, W - 1);
    const bool use_min = opts.reduce_op == DROSFM_REDUCE_MIN;
    // one base pointer per tensor (already at this lane's column); planes and rows are 32-bit offsets from it
    const float* __restrict__ ybase = image + static_cast<size_t>(b) * 3 * P + gxc;
    const float* __restrict__ xbase = warped + (static_cast<size_t>(ip) * NV * B + b) * 3 * P + gxc;
    const unsigned vstride = static_cast<unsigned>(B) * 3u * static_cast<unsigned>(P);            // < 2^31 (checked by the host)

    float ry[3], rx[NV][3];          // the next row's values, loaded one step ahead
    auto fetch = [&](int gy) {
        const unsigned off = static_cast<unsigned>(padded_row(gy, H) * W);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            ry[c] = col_in ? __ldg(ybase + (off + c * P)) : 0.0f;
#pragma unroll
            for (int v = 0; v < NV; ++v) rx[v][c] = col_in ? __ldg(xbase + (off + c * P + v * vstride)) : 0.0f;
        }
    };
    float local = 0.0f;
    FwdSums<NV> s12;
    FwdRow<NV> ra, rb;
    auto step = [&](FwdRow<NV>& prev, FwdRow<NV>& cur, int j) {
        const int gy = gy0 - 1 + j;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            cur.yc[c] = ry[c];
#pragma unroll
            for (int v = 0; v < NV; ++v) cur.xc[v][c] = rx[v][c];
        }
        fetch(gy + 1);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float yl, yr;
            neighbours(cur.yc[c], nb, yl, yr);
            const float y1 = cur.yc[c];
            cur.s.sy[c] = yl + y1 + yr;
            cur.s.syy[c] = yl * yl + y1 * y1 + yr * yr;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                float xl, xr;
                neighbours(cur.xc[v][c], nb, xl, xr);
                const float x1 = cur.xc[v][c];
                cur.s.sx[v][c] = xl + x1 + xr;
                cur.s.sxx[v][c] = xl * xl + x1 * x1 + xr * xr;
                cur.s.sxy[v][c] = xl * yl + x1 * y1 + xr * yr;
            }
        }
        if (j >= 2) {
            // window centred on row gy - 1 = rows (gy-2, gy-1) + gy
            float best = use_min ? __int_as_float(0x7f800000) : 0.0f;
            int sel = 254;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                float ssim_acc = 0.0f, l1_acc = 0.0f;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    Win w;
                    w.sx = s12.sx[v][c] + cur.s.sx[v][c];
                    w.sy = s12.sy[c] + cur.s.sy[c];
                    w.sxx = s12.sxx[v][c] + cur.s.sxx[v][c];
                    w.syy = s12.syy[c] + cur.s.syy[c];
                    w.sxy = s12.sxy[v][c] + cur.s.sxy[v][c];
                    const Ssim sm = ssim_from(w, opts.C1, opts.C2);
                    const float l = (1.0f - sm.s) * 0.5f;
                    ssim_acc += fminf(fmaxf(l, 0.0f), 1.0f);
                    l1_acc += fabsf(prev.xc[v][c] - prev.yc[c]);
                }
                const float pm = __fadd_rn(__fmul_rn(opts.ssim_w, third(ssim_acc)), __fmul_rn(l1_w, third(l1_acc)));
                if (use_min) {
                    if (pm < best) { best = pm; sel = v; }
                } else {
                    best += pm;
                }
            }
            const int gyo = gy - 1;
            if (out_lane && gyo < gy0 + kFwdBandH && gyo < H) {
                const size_t o = static_cast<size_t>(b) * P + gyo * W + gx;
                if (automask_in != nullptr) {
                    const float a = __ldg(automask_in + o);
                    if (a < best) { best = a; sel = 255; }
                }
                if (sel_out != nullptr) sel_out[static_cast<size_t>(ip) * B * P + o] = static_cast<uint8_t>(sel);
                local += best;
            }
        }
        // (prev + cur) is the first half of the next window
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            s12.sy[c] = prev.s.sy[c] + cur.s.sy[c];
            s12.syy[c] = prev.s.syy[c] + cur.s.syy[c];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                s12.sx[v][c] = prev.s.sx[v][c] + cur.s.sx[v][c];
                s12.sxx[v][c] = prev.s.sxx[v][c] + cur.s.sxx[v][c];
                s12.sxy[v][c] = prev.s.sxy[v][c] + cur.s.sxy[v][c];
            }
        }
    };
    static_assert((kFwdBandH + 2) % 2 == 0, "the row loop is unrolled by two");
    if (active) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            rb.s.sy[c] = rb.s.syy[c] = 0.0f;
#pragma unroll
            for (int v = 0; v < NV; ++v) rb.s.sx[v][c] = rb.s.sxx[v][c] = rb.s.sxy[v][c] = 0.0f;
        }
        fetch(gy0 - 1);
#pragma unroll 1
        for (int j = 0; j < kFwdBandH + 2; j += 2) {
            step(rb, ra, j);
            step(ra, rb, j + 1);
        }
    }

    // reduction of the loss, as in the tile kernel
    double part = warp_sum(static_cast<double>(local));
    if (lane == 0) red[wib] = part;
    __syncthreads();
    if (tid == 0) {
        double sacc = 0.0;
        for (int k = 0; k < kSsimWarps; ++k) sacc += red[k];
        atomicAdd(spread_acc(slot_at(ws, ip)), sacc);
    }
    Slot* ticket = slot_at(ws, n_preds);
    if (last_block(ticket, gridDim.x * gridDim.y, &flag) && tid < 32)
        finish_weighted_means(ws, n_preds, pp.weight, static_cast<double>(B) * P * (use_min ? 1.0 : static_cast<double>(NV)), ticket,
                              loss);
}

__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 shfl2(float2 v, int src) {
    return make_float2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}


// Forward, exactly two views: the same walk as ssim_fwd_stream_kernel<2> with the two views' sums, statistics and
// maps as float2 (half the issue slots for everything that is per view).
struct FwdSums2 {
    float sy[3], syy[3];
    float2 sx[3], sxx[3], sxy[3];
};
struct FwdRow2 {
    FwdSums2 s;
    float yc[3];
    float2 xc[3];
};

// blocks per SM (B200, benchmark workload): 1 -> 202 us, 2 -> 138, 3 -> 231, 4 -> 326 (spills)
#ifndef DROSFM_SSIMF_MINBLOCKS
#define DROSFM_SSIMF_MINBLOCKS 2
#endif
// MASK: the auto-mask pass -- the two "warped" pictures are the UN-warped sources pp.context[0..1], the per-pixel minimum of
// their photometric maps is written to mask_out [B,H,W]; no loss, no selection, no finisher.
template <bool MASK>
__global__ void __launch_bounds__(kSsimThreads, DROSFM_SSIMF_MINBLOCKS)
ssim_fwd_stream2_kernel(const float* __restrict__ image, const float* __restrict__ warped, const __grid_constant__ PhotoPtrs pp,
                        int n_preds, const float* __restrict__ automask_in, drosfm_photo_opts_t opts, float l1_w,
                        uint8_t* __restrict__ sel_out, float* __restrict__ loss, Slot* ws, float* __restrict__ mask_out, int B, int H,
                        int W, int nstrips, int nbands) {
    __shared__ double red[kSsimWarps];
    __shared__ int flag;
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    const int wg = blockIdx.x * kSsimWarps + wib;
    const bool active = wg < nstrips * nbands;
    const int strip = active ? wg % nstrips : 0, band = active ? wg / nstrips : 0;
    const int b = static_cast<int>(blockIdx.y) % B, ip = static_cast<int>(blockIdx.y) / B;
    const int P = H * W;
    const int gx = strip * kFwdStripW - 1 + lane, gy0 = band * kFwdBandH;
    const bool col_in = gx >= 0 && gx < W;
    const bool out_lane = active && lane >= 1 && lane <= kFwdStripW && gx < W;
    const Lanes nb = neighbour_lanes(lane, gx == 0, gx == W - 1);
    const int gxc = clampi(gx, 0, W - 1);
    const bool use_min = opts.reduce_op == DROSFM_REDUCE_MIN;
    const float* __restrict__ ybase = image + static_cast<size_t>(b) * 3 * P + gxc;
    const float* __restrict__ xbase = MASK ? pp.context[0] + static_cast<size_t>(b) * 3 * P + gxc
                                           : warped + (static_cast<size_t>(ip) * 2 * B + b) * 3 * P + gxc;
    const float* __restrict__ xbase1 = MASK ? pp.context[1] + static_cast<size_t>(b) * 3 * P + gxc
                                            : xbase + static_cast<size_t>(B) * 3u * static_cast<unsigned>(P);
    const float2 inv9 = bc2(1.0f / 9.0f), ninv9 = bc2(-1.0f / 9.0f), two = bc2(2.0f);
    const float2 C1 = bc2(opts.C1), C2 = bc2(opts.C2);

    float ry[3];
    float2 rx[3];
    auto fetch = [&](int gy) {
        const unsigned off = static_cast<unsigned>(padded_row(gy, H) * W);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            ry[c] = col_in ? __ldg(ybase + (off + c * P)) : 0.0f;
            rx[c].x = col_in ? __ldg(xbase + (off + c * P)) : 0.0f;
            rx[c].y = col_in ? __ldg(xbase1 + (off + c * P)) : 0.0f;
        }
    };
    float local = 0.0f;
    FwdSums2 s12;
    FwdRow2 ra, rb;
    auto step = [&](FwdRow2& prev, FwdRow2& cur, int j) {
        const int gy = gy0 - 1 + j;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            cur.yc[c] = ry[c];
            cur.xc[c] = rx[c];
        }
        fetch(gy + 1);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float yl, yr;
            neighbours(cur.yc[c], nb, yl, yr);
            const float y1 = cur.yc[c];
            cur.s.sy[c] = yl + y1 + yr;
            cur.s.syy[c] = yl * yl + y1 * y1 + yr * yr;
            const float2 xl = shfl2(cur.xc[c], nb.l), xr = shfl2(cur.xc[c], nb.r), x1 = cur.xc[c];
            cur.s.sx[c] = add2(add2(xl, x1), xr);
            cur.s.sxx[c] = fma2(xr, xr, fma2(x1, x1, mul2(xl, xl)));
            cur.s.sxy[c] = fma2(xr, bc2(yr), fma2(x1, bc2(y1), mul2(xl, bc2(yl))));
        }
        if (j >= 2) {
            float2 ssim_acc = bc2(0.0f), l1_acc = bc2(0.0f);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float2 wsx = add2(s12.sx[c], cur.s.sx[c]);
                const float2 wsxx = add2(s12.sxx[c], cur.s.sxx[c]);
                const float2 wsxy = add2(s12.sxy[c], cur.s.sxy[c]);
                const float wsy = s12.sy[c] + cur.s.sy[c], wsyy = s12.syy[c] + cur.s.syy[c];
                const float mu_y = wsy * (1.0f / 9.0f);
                const float mu_yy = mu_y * mu_y;
                const float sig_y = wsyy * (1.0f / 9.0f) - mu_yy;
                const float2 mu_y2 = bc2(mu_y);
                const float2 mu_x = mul2(wsx, inv9), nmu_x = mul2(wsx, ninv9);
                const float2 mu_xy = mul2(mu_x, mu_y2), mu_xx = mul2(mu_x, mu_x);
                const float2 sig_x = fma2(nmu_x, mu_x, mul2(wsxx, inv9));
                const float2 sig_xy = fma2(nmu_x, mu_y2, mul2(wsxy, inv9));
                const float2 A1 = fma2(two, mu_xy, C1), A2 = fma2(two, sig_xy, C2);
                const float2 B1 = add2(mu_xx, bc2(mu_yy + opts.C1)), B2 = add2(sig_x, bc2(sig_y + opts.C2));
                const float2 num = mul2(A1, A2), den = mul2(B1, B2);
                const float2 sm = make_float2(__fdividef(num.x, den.x), __fdividef(num.y, den.y));
                // clamp((1 - s) / 2, 0, 1)
                const float2 l = fma2(sm, bc2(-0.5f), bc2(0.5f));
                ssim_acc.x += fminf(fmaxf(l.x, 0.0f), 1.0f);
                ssim_acc.y += fminf(fmaxf(l.y, 0.0f), 1.0f);
                const float2 df = fma2(bc2(-1.0f), bc2(prev.yc[c]), prev.xc[c]);
                l1_acc.x += fabsf(df.x);
                l1_acc.y += fabsf(df.y);
            }
            const float pm0 = __fadd_rn(__fmul_rn(opts.ssim_w, third(ssim_acc.x)), __fmul_rn(l1_w, third(l1_acc.x)));
            const float pm1 = __fadd_rn(__fmul_rn(opts.ssim_w, third(ssim_acc.y)), __fmul_rn(l1_w, third(l1_acc.y)));
            float best;
            int sel;
            if (use_min) {
                best = __int_as_float(0x7f800000);
                sel = 254;
                if (pm0 < best) { best = pm0; sel = 0; }
                if (pm1 < best) { best = pm1; sel = 1; }
            } else {
                best = pm0 + pm1;
                sel = 254;
            }
            const int gyo = gy - 1;
            if (out_lane && gyo < gy0 + kFwdBandH && gyo < H) {
                const size_t o = static_cast<size_t>(b) * P + gyo * W + gx;
                if constexpr (MASK) {
                    mask_out[o] = best;
                } else {
                    if (automask_in != nullptr) {
                        const float a = __ldg(automask_in + o);
                        if (a < best) { best = a; sel = 255; }
                    }
                    if (sel_out != nullptr) sel_out[static_cast<size_t>(ip) * B * P + o] = static_cast<uint8_t>(sel);
                    local += best;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            s12.sy[c] = prev.s.sy[c] + cur.s.sy[c];
            s12.syy[c] = prev.s.syy[c] + cur.s.syy[c];
            s12.sx[c] = add2(prev.s.sx[c], cur.s.sx[c]);
            s12.sxx[c] = add2(prev.s.sxx[c], cur.s.sxx[c]);
            s12.sxy[c] = add2(prev.s.sxy[c], cur.s.sxy[c]);
        }
    };
    if (active) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            rb.s.sy[c] = rb.s.syy[c] = 0.0f;
            rb.s.sx[c] = rb.s.sxx[c] = rb.s.sxy[c] = bc2(0.0f);
        }
        fetch(gy0 - 1);
#pragma unroll 1
        for (int j = 0; j < kFwdBandH + 2; j += 2) {
            step(rb, ra, j);
            step(ra, rb, j + 1);
        }
    }

    if constexpr (MASK) return;
    double part = warp_sum(static_cast<double>(local));
    if (lane == 0) red[wib] = part;
    __syncthreads();
    if (tid == 0) {
        double sacc = 0.0;
        for (int k = 0; k < kSsimWarps; ++k) sacc += red[k];
        atomicAdd(spread_acc(slot_at(ws, ip)), sacc);
    }
    Slot* ticket = slot_at(ws, n_preds);
    if (last_block(ticket, gridDim.x * gridDim.y, &flag) && tid < 32)
        finish_weighted_means(ws, n_preds, pp.weight, static_cast<double>(B) * P * (use_min ? 1.0 : 2.0), ticket, loss);
}

struct BwdRow {
    float sy, syy, sx, sxx, sxy;     // horizontal 3-sums of this row
    float x, y;                      // its centre values
    float ha, hb, hc;                // horizontal (multiplicity-weighted) 3-sums of the coefficients of the windows centred on it
};

// One warp: one strip x band of ONE colour channel of one (prediction, view, sample).  Writes d loss / d warped.
__global__ void __launch_bounds__(kSsimThreads, 4)
ssim_bwd_stream_kernel(const float* __restrict__ g_loss, const float* __restrict__ image, const float* __restrict__ warped,
                       const __grid_constant__ PhotoPtrs pp, int V, const uint8_t* __restrict__ sel_in,
                       drosfm_photo_opts_t opts, float l1_w, float* __restrict__ g_warped, int B, int H, int W, int nstrips,
                       int nbands) {
    const int lane = threadIdx.x & 31;
    const int wg = blockIdx.x * kSsimWarps + (threadIdx.x >> 5);
    if (wg >= nstrips * nbands) return;
    const int strip = wg % nstrips, band = wg / nstrips;
    const int c = static_cast<int>(blockIdx.y) % 3, slot = static_cast<int>(blockIdx.y) / 3;      // slot = (ip * V + v) * B + b
    const int b = slot % B, v = (slot / B) % V, ip = slot / (B * V);
    const int P = H * W;
    const int gx = strip * kBwdStripW - 2 + lane, gy0 = band * kBwdBandH;
    const bool col_in = gx >= 0 && gx < W;
    const bool out_lane = lane >= 2 && lane <= kBwdStripW + 1 && gx < W;
    const Lanes nb = neighbour_lanes(lane, gx == 0, gx == W - 1);
    const int gxc = clampi(gx, 0, W - 1);
    const bool use_min = opts.reduce_op == DROSFM_REDUCE_MIN;
    const float G = __ldg(g_loss) * pp.weight[ip] /
                    (static_cast<float>(B) * static_cast<float>(P) * (use_min ? 1.0f : static_cast<float>(V)));
    const float kp = G * opts.ssim_w * (-1.0f / 6.0f) * (2.0f / 9.0f);   // d loss / d ssim  x  2/9 of the window derivative
    const float kl1 = G * l1_w * (1.0f / 3.0f);
    const float* __restrict__ ypl = image + (static_cast<size_t>(b) * 3 + c) * P + gxc;
    const float* __restrict__ xpl = warped + (static_cast<size_t>(slot) * 3 + c) * P + gxc;
    float* __restrict__ gpl = g_warped + (static_cast<size_t>(slot) * 3 + c) * P + gxc;
    const uint8_t* __restrict__ spl = sel_in + (static_cast<size_t>(ip) * B + b) * P + gxc;
    // how often column gx occurs in the (reflected) window centred on gx-1 / gx / gx+1
    const float wx0 = gx <= 0 ? 0.0f : (gx == 1 ? 2.0f : 1.0f);
    const float wx2 = gx >= W - 1 ? 0.0f : (gx == W - 2 ? 2.0f : 1.0f);

    float nx, ny;
    auto fetch = [&](int gy) {
        const unsigned off = static_cast<unsigned>(padded_row(gy, H) * W);
        nx = col_in ? __ldg(xpl + off) : 0.0f;
        ny = col_in ? __ldg(ypl + off) : 0.0f;
    };
    int sv_prev = 254;
    auto step = [&](BwdRow& p2, BwdRow& p1, BwdRow& cur, int j) {
        const int gy = gy0 - 2 + j;          // row loaded in this step; windows centred on gy-1; gradients of row gy-2
        const int gc = gy - 1;
        cur.x = nx;
        cur.y = ny;
        fetch(gy + 1);
        int sv = 254;
        if (j >= 2 && col_in && gc >= 0 && gc < H) sv = use_min ? static_cast<int>(__ldg(spl + static_cast<unsigned>(gc * W))) : 253;
        float xl, xr, yl, yr;
        neighbours(cur.x, nb, xl, xr);
        neighbours(cur.y, nb, yl, yr);
        cur.sx = xl + cur.x + xr;
        cur.sy = yl + cur.y + yr;
        cur.sxx = xl * xl + cur.x * cur.x + xr * xr;
        cur.syy = yl * yl + cur.y * cur.y + yr * yr;
        cur.sxy = xl * yl + cur.x * cur.y + xr * yr;
        if (j >= 2) {
            float a = 0.0f, bb = 0.0f, cq = 0.0f;
            if (sv == v || sv == 253) {
                Win w;
                w.sx = p2.sx + p1.sx + cur.sx;
                w.sy = p2.sy + p1.sy + cur.sy;
                w.sxx = p2.sxx + p1.sxx + cur.sxx;
                w.syy = p2.syy + p1.syy + cur.syy;
                w.sxy = p2.sxy + p1.sxy + cur.sxy;
                const Ssim sm = ssim_from(w, opts.C1, opts.C2);
                const float l = (1.0f - sm.s) * 0.5f;
                if (l >= 0.0f && l <= 1.0f) {
                    const float q = __fdividef(kp, sm.B1 * sm.B2);
                    a = q * (sm.mu_y * (sm.A2 - sm.A1) - sm.s * sm.mu_x * (sm.B2 - sm.B1));
                    bb = -q * sm.s * sm.B1;
                    cq = q * sm.A1;
                }
            }
            const float al = __shfl_up_sync(0xffffffffu, a, 1), ar = __shfl_down_sync(0xffffffffu, a, 1);
            const float bl = __shfl_up_sync(0xffffffffu, bb, 1), br = __shfl_down_sync(0xffffffffu, bb, 1);
            const float cl = __shfl_up_sync(0xffffffffu, cq, 1), cr = __shfl_down_sync(0xffffffffu, cq, 1);
            p1.ha = wx0 * al + a + wx2 * ar;
            p1.hb = wx0 * bl + bb + wx2 * br;
            p1.hc = wx0 * cl + cq + wx2 * cr;
        }
        if (j >= 4) {
            const int gq = gy - 2;
            if (out_lane && gq < gy0 + kBwdBandH && gq < H) {
                // vertical multiplicities of row gq in the windows centred on gq-1 / gq / gq+1
                const float wy0 = gq <= 0 ? 0.0f : (gq == 1 ? 2.0f : 1.0f);
                const float wy2 = gq >= H - 1 ? 0.0f : (gq == H - 2 ? 2.0f : 1.0f);
                const float ga = wy0 * cur.ha + p2.ha + wy2 * p1.ha;
                const float gb = wy0 * cur.hb + p2.hb + wy2 * p1.hb;
                const float gc_ = wy0 * cur.hc + p2.hc + wy2 * p1.hc;
                float gxv = ga + gb * p2.x + gc_ * p2.y;
                if (sv_prev == v || sv_prev == 253) {
                    const float df = p2.x - p2.y;
                    // sign(df) * kl1: kl1 with its sign flipped where df is negative
                    gxv += df == 0.0f ? 0.0f : __int_as_float(__float_as_int(kl1) ^ (__float_as_int(df) & 0x80000000));
                }
                gpl[static_cast<unsigned>(gq * W)] = gxv;
            }
        }
        sv_prev = sv;
    };
    static_assert((kBwdBandH + 4) % 3 == 0, "the row loop is unrolled by three");
    BwdRow r0, r1, r2;
    r0.ha = r0.hb = r0.hc = r1.ha = r1.hb = r1.hc = r2.ha = r2.hb = r2.hc = 0.0f;
    r0.sx = r0.sy = r0.sxx = r0.syy = r0.sxy = r1.sx = r1.sy = r1.sxx = r1.syy = r1.sxy = 0.0f;
    r0.x = r0.y = r1.x = r1.y = 0.0f;
    fetch(gy0 - 2);
#pragma unroll 1
    for (int j = 0; j < kBwdBandH + 4; j += 3) {
        step(r1, r2, r0, j);
        step(r2, r0, r1, j + 1);
        step(r0, r1, r2, j + 2);
    }
}

// ---- two views per warp, packed fp32x2 ---------------------------------------------------------
// These kernels are bound by instruction issue, not by the FMA pipe, and Blackwell's packed fp32x2 arithmetic
// (fma.rn.f32x2 & co: two IEEE fp32 operations per issued instruction) halves the issue slots of everything that is
// element-wise over a pair.  The backward stage pairs two source VIEWS in one warp: their statistics, coefficients
// and box sums travel as float2, and the target-image work (loads, shuffles, sums of y, the sel byte) is shared.
struct BwdRow2 {
    float sy, syy;                   // target row: horizontal 3-sums
    float2 sx, sxx, sxy;             // the two views' rows
    float2 x;
    float y;
    float2 ha, hb, hc;
};

// 128-thread blocks
constexpr int kBwd2Threads = 128, kBwd2Warps = kBwd2Threads / 32;

// blocks per SM (B200, benchmark workload): 3 -> 203.5 us, 4 -> 207, 5 -> 213, 6 -> 270 (spills)
#ifndef DROSFM_SSIMB_MINBLOCKS
#define DROSFM_SSIMB_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(kBwd2Threads, DROSFM_SSIMB_MINBLOCKS)
ssim_bwd_stream2_kernel(const float* __restrict__ g_loss, const float* __restrict__ image, const float* __restrict__ warped,
                        const __grid_constant__ PhotoPtrs pp, int V, const uint8_t* __restrict__ sel_in,
                        drosfm_photo_opts_t opts, float l1_w, float* __restrict__ g_warped, int B, int H, int W, int nstrips,
                        int nbands) {
    const int lane = threadIdx.x & 31;
    const int wg = blockIdx.x * kBwd2Warps + (threadIdx.x >> 5);
    if (wg >= nstrips * nbands) return;
    const int strip = wg % nstrips, band = wg / nstrips;
    const int c = static_cast<int>(blockIdx.y) % 3, pair = static_cast<int>(blockIdx.y) / 3;      // pair = (ip * V/2 + vp) * B + b
    const int VP = V >> 1;
    const int b = pair % B, vp = (pair / B) % VP, ip = pair / (B * VP);
    const int v0 = 2 * vp;
    const int P = H * W;
    const int gx = strip * kBwdStripW - 2 + lane, gy0 = band * kBwdBandH;
    const bool col_in = gx >= 0 && gx < W;
    const bool out_lane = lane >= 2 && lane <= kBwdStripW + 1 && gx < W;
    const Lanes nb = neighbour_lanes(lane, gx == 0, gx == W - 1);
    const int gxc = clampi(gx, 0, W - 1);
    const bool use_min = opts.reduce_op == DROSFM_REDUCE_MIN;
    const float G = __ldg(g_loss) * pp.weight[ip] /
                    (static_cast<float>(B) * static_cast<float>(P) * (use_min ? 1.0f : static_cast<float>(V)));
    const float kp = G * opts.ssim_w * (-1.0f / 6.0f) * (2.0f / 9.0f);
    const float kl1 = G * l1_w * (1.0f / 3.0f);
    const size_t slot0 = (static_cast<size_t>(ip) * V + v0) * B + b;
    const unsigned vstride = static_cast<unsigned>(B) * 3u * static_cast<unsigned>(P);     // view v0 + 1 relative to v0
    const float* __restrict__ ypl = image + (static_cast<size_t>(b) * 3 + c) * P + gxc;
    const float* __restrict__ xpl = warped + (slot0 * 3 + c) * P + gxc;
    float* __restrict__ gpl = g_warped + (slot0 * 3 + c) * P + gxc;
    const uint8_t* __restrict__ spl = sel_in + (static_cast<size_t>(ip) * B + b) * P + gxc;
    const float2 wx0 = bc2(gx <= 0 ? 0.0f : (gx == 1 ? 2.0f : 1.0f));
    const float2 wx2 = bc2(gx >= W - 1 ? 0.0f : (gx == W - 2 ? 2.0f : 1.0f));
    const float2 inv9 = bc2(1.0f / 9.0f), ninv9 = bc2(-1.0f / 9.0f), two = bc2(2.0f);
    const float2 C1 = bc2(opts.C1), C2 = bc2(opts.C2);

    // rows are requested two steps before they are used (three rotating buffers); the sel byte of a row travels with it
    // and is consumed one step later still, when that row is the window centre
    struct Pre {
        float2 x;
        float y;
        int sel;
    };
    Pre f0, f1, f2;
    auto fetch = [&](int gy, Pre& f) {
        const unsigned off = static_cast<unsigned>(padded_row(gy, H) * W);
        f.x.x = col_in ? __ldg(xpl + off) : 0.0f;
        f.x.y = col_in ? __ldg(xpl + (off + vstride)) : 0.0f;
        f.y = col_in ? __ldg(ypl + off) : 0.0f;
        f.sel = 254;
        if (col_in && gy >= 0 && gy < H) f.sel = use_min ? static_cast<int>(__ldg(spl + static_cast<unsigned>(gy * W))) : 253;
    };
    int sel_row = 254;      // sel of the row loaded in the previous step = the centre row of this step
    int sv_prev = 254;
    auto step = [&](BwdRow2& p2, BwdRow2& p1, BwdRow2& cur, Pre& mine, Pre& refill, int j) {
        const int gy = gy0 - 2 + j;
        cur.x = mine.x;
        cur.y = mine.y;
        const int sv = j >= 2 ? sel_row : 254;
        sel_row = mine.sel;
        fetch(gy + 2, refill);
        const float2 xl = shfl2(cur.x, nb.l), xr = shfl2(cur.x, nb.r);
        float yl, yr;
        neighbours(cur.y, nb, yl, yr);
        const float2 yl2 = bc2(yl), yc2 = bc2(cur.y), yr2 = bc2(yr);
        cur.sy = yl + cur.y + yr;
        cur.syy = yl * yl + cur.y * cur.y + yr * yr;
        cur.sx = add2(add2(xl, cur.x), xr);
        cur.sxx = fma2(xr, xr, fma2(cur.x, cur.x, mul2(xl, xl)));
        cur.sxy = fma2(xr, yr2, fma2(cur.x, yc2, mul2(xl, yl2)));
        if (j >= 2) {
            float2 a = bc2(0.0f), bb = a, cq = a;
            const bool on0 = sv == v0 || sv == 253, on1 = sv == v0 + 1 || sv == 253;
            if (on0 || on1) {
                const float2 wsx = add2(add2(p2.sx, p1.sx), cur.sx);
                const float2 wsxx = add2(add2(p2.sxx, p1.sxx), cur.sxx);
                const float2 wsxy = add2(add2(p2.sxy, p1.sxy), cur.sxy);
                const float wsy = p2.sy + p1.sy + cur.sy, wsyy = p2.syy + p1.syy + cur.syy;
                // statistics as in ssim_from(), two views at a time
                const float mu_y = wsy * (1.0f / 9.0f);
                const float mu_yy = mu_y * mu_y;
                const float sig_y = wsyy * (1.0f / 9.0f) - mu_yy;
                const float2 mu_y2 = bc2(mu_y);
                const float2 mu_x = mul2(wsx, inv9), nmu_x = mul2(wsx, ninv9);
                const float2 mu_xy = mul2(mu_x, mu_y2), mu_xx = mul2(mu_x, mu_x);
                const float2 sig_x = fma2(nmu_x, mu_x, mul2(wsxx, inv9));
                const float2 sig_xy = fma2(nmu_x, mu_y2, mul2(wsxy, inv9));
                const float2 A1 = fma2(two, mu_xy, C1), A2 = fma2(two, sig_xy, C2);
                const float2 B1 = add2(mu_xx, bc2(mu_yy + opts.C1)), B2 = add2(sig_x, bc2(sig_y + opts.C2));
                const float2 num = mul2(A1, A2), den = mul2(B1, B2);
                const float2 rden = make_float2(__fdividef(1.0f, den.x), __fdividef(1.0f, den.y));
                const float2 sm = mul2(num, rden);
                const float2 q = mul2(bc2(kp), rden);
                // a = q (mu_y (A2 - A1) - s mu_x (B2 - B1)),  b = -q s B1,  c = q A1
                const float2 neg1 = bc2(-1.0f);
                const float2 dA = fma2(neg1, A1, A2), dB = fma2(neg1, B1, B2);
                const float2 inner = fma2(mul2(sm, nmu_x), dB, mul2(mu_y2, dA));
                const float2 a_ = mul2(q, inner);
                const float2 b_ = mul2(mul2(mul2(bc2(-kp), rden), sm), B1);
                const float2 c_ = mul2(q, A1);
                const float l0 = (1.0f - sm.x) * 0.5f, l1 = (1.0f - sm.y) * 0.5f;
                const bool k0 = on0 && l0 >= 0.0f && l0 <= 1.0f, k1 = on1 && l1 >= 0.0f && l1 <= 1.0f;
                a = make_float2(k0 ? a_.x : 0.0f, k1 ? a_.y : 0.0f);
                bb = make_float2(k0 ? b_.x : 0.0f, k1 ? b_.y : 0.0f);
                cq = make_float2(k0 ? c_.x : 0.0f, k1 ? c_.y : 0.0f);
            }
            const int ll = (lane - 1) & 31, lr = (lane + 1) & 31;      // the outermost lanes' sums are never used
            p1.ha = fma2(wx2, shfl2(a, lr), fma2(wx0, shfl2(a, ll), a));
            p1.hb = fma2(wx2, shfl2(bb, lr), fma2(wx0, shfl2(bb, ll), bb));
            p1.hc = fma2(wx2, shfl2(cq, lr), fma2(wx0, shfl2(cq, ll), cq));
        }
        if (j >= 4) {
            const int gq = gy - 2;
            if (out_lane && gq < gy0 + kBwdBandH && gq < H) {
                const float2 wy0 = bc2(gq <= 0 ? 0.0f : (gq == 1 ? 2.0f : 1.0f));
                const float2 wy2 = bc2(gq >= H - 1 ? 0.0f : (gq == H - 2 ? 2.0f : 1.0f));
                const float2 ga = fma2(wy2, p1.ha, fma2(wy0, cur.ha, p2.ha));
                const float2 gb = fma2(wy2, p1.hb, fma2(wy0, cur.hb, p2.hb));
                const float2 gc_ = fma2(wy2, p1.hc, fma2(wy0, cur.hc, p2.hc));
                float2 gxv = fma2(gc_, bc2(p2.y), fma2(gb, p2.x, ga));
                const bool q0 = sv_prev == v0 || sv_prev == 253, q1 = sv_prev == v0 + 1 || sv_prev == 253;
                if (q0) {
                    const float df = p2.x.x - p2.y;
                    gxv.x += df == 0.0f ? 0.0f : __int_as_float(__float_as_int(kl1) ^ (__float_as_int(df) & 0x80000000));
                }
                if (q1) {
                    const float df = p2.x.y - p2.y;
                    gxv.y += df == 0.0f ? 0.0f : __int_as_float(__float_as_int(kl1) ^ (__float_as_int(df) & 0x80000000));
                }
                const unsigned o = static_cast<unsigned>(gq * W);
                gpl[o] = gxv.x;
                gpl[o + vstride] = gxv.y;
            }
        }
        sv_prev = sv;
    };
    BwdRow2 r0, r1, r2;
    r0.ha = r0.hb = r0.hc = r1.ha = r1.hb = r1.hc = r2.ha = r2.hb = r2.hc = bc2(0.0f);
    r0.sx = r0.sxx = r0.sxy = r1.sx = r1.sxx = r1.sxy = bc2(0.0f);
    r0.sy = r0.syy = r1.sy = r1.syy = 0.0f;
    r0.x = r1.x = bc2(0.0f);
    r0.y = r1.y = 0.0f;
    fetch(gy0 - 2, f0);
    fetch(gy0 - 1, f1);
#pragma unroll 1
    for (int j = 0; j < kBwdBandH + 4; j += 3) {
        step(r1, r2, r0, f0, f2, j);
        step(r2, r0, r1, f1, f0, j + 1);
        step(r0, r1, r2, f2, f1, j + 2);
    }
}

// ---- training forward, 2 / 4 / 6 / 8 views: SSIM forward AND the window gradients in one pass -------------------
// The backward SSIM stage recomputes every window statistic the forward stage had: with the upstream gradient being
// a scalar factor, d loss / d warped can be produced by the forward itself (unscaled: the warp adjoint multiplies by the
// upstream gradient of the loss), and the backward pass of the loss shrinks to the warp adjoint.
//
// Block = 3 warps: the three colour channels of one 28-column strip x 32-row band of one (prediction, sample), both views
// packed as float2 (the walk of ssim_bwd_stream2_kernel).  Per row step each channel warp computes its SSIM / L1 terms of
// the window just completed; the three warps exchange them through shared memory (three buffers at static offsets, ONE
// named barrier per step), every warp forms the photometric values of all views, the min / auto-mask selection and from it
// the coefficients of its own channel.  Channel 0 accumulates the loss and writes the selection.  More than two views: one
// such warp triple per view PAIR in the block (template PAIRS), all behind the same barrier.
constexpr int kTrainThreads = 96;      // per view PAIR: three channel warps
#ifndef DROSFM_SSIMT_INNER
#define DROSFM_SSIMT_INNER 1      // bands away from the top / bottom edge run a copy of the walk without the row tests
#endif
#ifndef DROSFM_SSIMT_PFD
#define DROSFM_SSIMT_PFD 3      // rows between a load and its use
#endif
#ifndef DROSFM_SSIMT_HSMEM
#define DROSFM_SSIMT_HSMEM 1    // the horizontal coefficient sums of the three rows in flight in shared memory (18 registers less)
#endif
#ifndef DROSFM_SSIMT_BAND
#define DROSFM_SSIMT_BAND 32
#endif
constexpr int kTrainBandH = DROSFM_SSIMT_BAND;      // rows written per block (4 more are loaded)

#ifndef DROSFM_SSIMT_MINBLOCKS
#define DROSFM_SSIMT_MINBLOCKS 5      // blocks per SM (15 warps, 124 registers, no spill); any spill costs > 10 %
#endif
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float2 third2(float2 x) {      // third() of both halves
    const float2 r = bc2(1.0f / 3.0f);
    const float2 q = mul2(x, r);
    return fma2(fma2(bc2(-3.0f), q, x), r, q);
}

// PAIRS = V / 2 view pairs per block (3 * PAIRS warps): every pair's warps exchange their terms, every warp forms the
// photometric values of ALL views and the selection, and keeps the coefficients of its own pair and channel.
template <int PAIRS>
__global__ void __launch_bounds__(kTrainThreads * PAIRS, PAIRS == 1 ? DROSFM_SSIMT_MINBLOCKS : (PAIRS == 2 ? 2 : 1))
ssim_train_stream2_kernel(const float* __restrict__ image, const float* __restrict__ warped, const __grid_constant__ PhotoPtrs pp,
                          int n_preds, const float* __restrict__ automask_in, drosfm_photo_opts_t opts, float l1_w,
                          uint8_t* __restrict__ sel_out, float* __restrict__ loss, Slot* ws, float* __restrict__ g_warped,
                          int B, int H, int W, int nstrips, int nbands) {
    __shared__ float4 xchg[3][PAIRS][3][32];      // [row step within the unrolled triple][pair][channel][lane]: static offsets
    __shared__ int flag;
#if DROSFM_SSIMT_HSMEM
    // the horizontal coefficient sums of the three rows in flight live in shared memory (private per thread: no barrier),
    // 18 registers less per thread
    __shared__ float2 hrow[3][3][kTrainThreads * PAIRS];
    hrow[0][0][threadIdx.x] = hrow[0][1][threadIdx.x] = hrow[0][2][threadIdx.x] = make_float2(0.0f, 0.0f);
    hrow[1][0][threadIdx.x] = hrow[1][1][threadIdx.x] = hrow[1][2][threadIdx.x] = make_float2(0.0f, 0.0f);
    hrow[2][0][threadIdx.x] = hrow[2][1][threadIdx.x] = hrow[2][2][threadIdx.x] = make_float2(0.0f, 0.0f);
#endif
    constexpr int V = 2 * PAIRS;
    const int pr = PAIRS == 1 ? 0 : static_cast<int>(threadIdx.x) / kTrainThreads;
    const int lane = threadIdx.x & 31, c = (static_cast<int>(threadIdx.x) - pr * kTrainThreads) >> 5;
    const int wg = blockIdx.x;
    const int strip = wg % nstrips, band = wg / nstrips;
    const int b = static_cast<int>(blockIdx.y) % B, ip = static_cast<int>(blockIdx.y) / B;
    const int P = H * W;
    const int gx = strip * kBwdStripW - 2 + lane, gy0 = band * kTrainBandH;
    const bool col_in = gx >= 0 && gx < W;
    const bool out_lane = lane >= 2 && lane <= kBwdStripW + 1 && gx < W;
    const Lanes nb = neighbour_lanes(lane, gx == 0, gx == W - 1);
    const int gxc = clampi(gx, 0, W - 1);
    const bool use_min = opts.reduce_op == DROSFM_REDUCE_MIN;
    // unscaled by the upstream gradient (see above)
    const float G = pp.weight[ip] / (static_cast<float>(B) * static_cast<float>(P) * (use_min ? 1.0f : static_cast<float>(V)));
    const float kp = G * opts.ssim_w * (-1.0f / 6.0f) * (2.0f / 9.0f);
    const float kl1 = G * l1_w * (1.0f / 3.0f);
    const size_t slot0 = (static_cast<size_t>(ip) * V + 2 * pr) * B + b;
    const unsigned vstride = static_cast<unsigned>(B) * 3u * static_cast<unsigned>(P);     // view 1 relative to view 0
    const float* __restrict__ ypl = image + (static_cast<size_t>(b) * 3 + c) * P + gxc;
    const float* __restrict__ xpl = warped + (slot0 * 3 + c) * P + gxc;
    float* __restrict__ gpl = g_warped + (slot0 * 3 + c) * P + gxc;
    const float* __restrict__ apl = automask_in != nullptr ? automask_in + static_cast<size_t>(b) * P + gxc : nullptr;
    uint8_t* __restrict__ spl = sel_out != nullptr ? sel_out + (static_cast<size_t>(ip) * B + b) * P + gxc : nullptr;
    const float2 wx0 = bc2(gx <= 0 ? 0.0f : (gx == 1 ? 2.0f : 1.0f));
    const float2 wx2 = bc2(gx >= W - 1 ? 0.0f : (gx == W - 2 ? 2.0f : 1.0f));
    const float2 inv9 = bc2(1.0f / 9.0f), ninv9 = bc2(-1.0f / 9.0f), two = bc2(2.0f);
    const float2 C1 = bc2(opts.C1), C2 = bc2(opts.C2);

    struct Pre {
        float2 x;
        float y, am;
    };
    Pre f0, f1, f2;
    // INNER (a block-uniform compile-time tag): every row the band touches lies inside the picture, at least two rows away
    // from its top and bottom edge -- no reflection, no row tests, unit vertical window weights
    // (the INNER walk reads and writes consecutive rows: running row pointers instead of row * W products)
    const float *xrow = nullptr, *yrow = nullptr, *arow = nullptr;
    float* grow = nullptr;
    uint8_t* srow = nullptr;
    auto fetch = [&](auto inner, int gy, Pre& f) {
        constexpr bool INNER = decltype(inner)::value;
        if constexpr (INNER) {
            f.x.x = col_in ? __ldg(xrow) : 0.0f;
            f.x.y = col_in ? __ldg(xrow + vstride) : 0.0f;
            f.y = col_in ? __ldg(yrow) : 0.0f;
            f.am = (apl != nullptr && col_in) ? __ldg(arow) : __int_as_float(0x7f800000);
            xrow += W;
            yrow += W;
            arow += W;
        } else {
            const unsigned off = static_cast<unsigned>(padded_row(gy, H) * W);
            f.x.x = col_in ? __ldg(xpl + off) : 0.0f;
            f.x.y = col_in ? __ldg(xpl + (off + vstride)) : 0.0f;
            f.y = col_in ? __ldg(ypl + off) : 0.0f;
            // the auto-mask value of a row travels with it (consumed one step later, when the row is a window centre)
            f.am = (apl != nullptr && col_in && gy >= 0 && gy < H) ? __ldg(apl + static_cast<unsigned>(gy * W)) : __int_as_float(0x7f800000);
        }
    };
    float am_row = __int_as_float(0x7f800000);      // auto-mask of the row loaded in the previous step = this step's window centre
    int sv_prev = 254;
    float local = 0.0f;
    auto step = [&](auto inner, auto kbuf, BwdRow2& p2, BwdRow2& p1, BwdRow2& cur, Pre& mine, Pre& refill, int j) {
        constexpr bool INNER = decltype(inner)::value;
        constexpr int KB = decltype(kbuf)::value;       // exchange buffer of this step (a buffer is reused three steps later)
        const int gy = gy0 - 2 + j;          // row loaded in this step; windows centred on gy-1; gradients of row gy-2
        const int gc = gy - 1;
        cur.x = mine.x;
        cur.y = mine.y;
        const float am = am_row;
        am_row = mine.am;
        fetch(inner, gy + DROSFM_SSIMT_PFD, refill);
        const float2 xl = shfl2(cur.x, nb.l), xr = shfl2(cur.x, nb.r);
        float yl, yr;
        neighbours(cur.y, nb, yl, yr);
        const float2 yl2 = bc2(yl), yc2 = bc2(cur.y), yr2 = bc2(yr);
        cur.sy = yl + cur.y + yr;
        cur.syy = yl * yl + cur.y * cur.y + yr * yr;
        cur.sx = add2(add2(xl, cur.x), xr);
        cur.sxx = fma2(xr, xr, fma2(cur.x, cur.x, mul2(xl, xl)));
        cur.sxy = fma2(xr, yr2, fma2(cur.x, yc2, mul2(xl, yl2)));
        int sv = 254;
#if DROSFM_SSIMT_HSMEM
        float2 h1a = bc2(0.0f), h1b = h1a, h1c = h1a;
#endif
        if (j >= 2) {
            // statistics of the window centred on (gc, gx), this channel, both views
            const float2 wsx = add2(add2(p2.sx, p1.sx), cur.sx);
            const float2 wsxx = add2(add2(p2.sxx, p1.sxx), cur.sxx);
            const float2 wsxy = add2(add2(p2.sxy, p1.sxy), cur.sxy);
            const float wsy = p2.sy + p1.sy + cur.sy, wsyy = p2.syy + p1.syy + cur.syy;
            const float mu_y = wsy * (1.0f / 9.0f);
            const float mu_yy = mu_y * mu_y;
            const float sig_y = wsyy * (1.0f / 9.0f) - mu_yy;
            const float2 mu_y2 = bc2(mu_y);
            const float2 mu_x = mul2(wsx, inv9), nmu_x = mul2(wsx, ninv9);
            const float2 mu_xy = mul2(mu_x, mu_y2), mu_xx = mul2(mu_x, mu_x);
            const float2 sig_x = fma2(nmu_x, mu_x, mul2(wsxx, inv9));
            const float2 sig_xy = fma2(nmu_x, mu_y2, mul2(wsxy, inv9));
            const float2 A1 = fma2(two, mu_xy, C1), A2 = fma2(two, sig_xy, C2);
            const float2 B1 = add2(mu_xx, bc2(mu_yy + opts.C1)), B2 = add2(sig_x, bc2(sig_y + opts.C2));
            const float2 num = mul2(A1, A2), den = mul2(B1, B2);
            // den >= C1 * C2 > 0, far inside the normal range: the bare reciprocal approximation (what __fdividef(1, x) uses,
            // without its range scaling)
            const float2 rden = make_float2(rcp_approx(den.x), rcp_approx(den.y));
            const float2 sm = mul2(num, rden);
            const float2 l01 = mul2(fma2(bc2(-1.0f), sm, bc2(1.0f)), bc2(0.5f));      // (1 - ssim) / 2, both views
            const float l0 = l01.x, l1 = l01.y;
            // this channel's terms of the photometric value of the window centre (row p1)
            const float2 dxy = fma2(bc2(-1.0f), bc2(p1.y), p1.x);                     // x - y, exact as a difference
            const float4 mineq = make_float4(fminf(fmaxf(l0, 0.0f), 1.0f), fminf(fmaxf(l1, 0.0f), 1.0f), fabsf(dxy.x), fabsf(dxy.y));
            xchg[KB][pr][c][lane] = mineq;
            asm volatile("bar.sync 1, %0;" ::"n"(kTrainThreads * PAIRS) : "memory");
            float best = use_min ? __int_as_float(0x7f800000) : 0.0f;
            if (!use_min) sv = 253;
#pragma unroll
            for (int q = 0; q < PAIRS; ++q) {
                const float4 q0 = xchg[KB][q][0][lane], q1 = xchg[KB][q][1][lane], q2 = xchg[KB][q][2][lane];
                // channel means and the weighted sum, both views at once (same operations, same order as the scalar form)
                const float2 ssum = add2(add2(make_float2(q0.x, q0.y), make_float2(q1.x, q1.y)), make_float2(q2.x, q2.y));
                const float2 lsum = add2(add2(make_float2(q0.z, q0.w), make_float2(q1.z, q1.w)), make_float2(q2.z, q2.w));
                const float2 pm = add2(mul2(bc2(opts.ssim_w), third2(ssum)), mul2(bc2(l1_w), third2(lsum)));
                if (use_min) {
                    if (pm.x < best) { best = pm.x; sv = 2 * q; }
                    if (pm.y < best) { best = pm.y; sv = 2 * q + 1; }
                } else {
                    best += pm.x;
                    best += pm.y;
                }
            }
            if (use_min && am < best) { best = am; sv = 255; }
            const bool centre_in = col_in && (INNER || (gc >= 0 && gc < H));
            if (!centre_in) sv = 254;
            if (c == 0 && pr == 0 && out_lane && j >= 3 && j < kTrainBandH + 3 && (INNER || gc < H)) {
                local += best;
                if (spl != nullptr) {
                    if constexpr (INNER) *srow = static_cast<uint8_t>(use_min ? sv : 254);
                    else spl[static_cast<unsigned>(gc * W)] = static_cast<uint8_t>(use_min ? sv : 254);
                }
            }
            // coefficients of the windows that carry a gradient
            // (branch-free: every quantity is finite -- den >= C1 * C2 -- so a window without a gradient is masked by a
            // multiplication with zero instead of a divergent branch around the block)
            const bool on0 = sv == 2 * pr || sv == 253, on1 = sv == 2 * pr + 1 || sv == 253;
            const bool k0 = on0 && l0 >= 0.0f && l0 <= 1.0f, k1 = on1 && l1 >= 0.0f && l1 <= 1.0f;
            const float2 q = mul2(mul2(bc2(kp), rden), make_float2(k0 ? 1.0f : 0.0f, k1 ? 1.0f : 0.0f));
            const float2 neg1 = bc2(-1.0f);
            const float2 dA = fma2(neg1, A1, A2), dB = fma2(neg1, B1, B2);
            const float2 inner = fma2(mul2(sm, nmu_x), dB, mul2(mu_y2, dA));
            const float2 a = mul2(q, inner);
            const float2 bb = mul2(mul2(mul2(q, neg1), sm), B1);
            const float2 cq = mul2(q, A1);
            const int ll = (lane - 1) & 31, lr = (lane + 1) & 31;      // the outermost lanes' sums are never used
#if DROSFM_SSIMT_HSMEM
            h1a = fma2(wx2, shfl2(a, lr), fma2(wx0, shfl2(a, ll), a));
            h1b = fma2(wx2, shfl2(bb, lr), fma2(wx0, shfl2(bb, ll), bb));
            h1c = fma2(wx2, shfl2(cq, lr), fma2(wx0, shfl2(cq, ll), cq));
            hrow[(KB + 2) % 3][0][threadIdx.x] = h1a;       // p1 of step KB is struct (KB + 2) % 3
            hrow[(KB + 2) % 3][1][threadIdx.x] = h1b;
            hrow[(KB + 2) % 3][2][threadIdx.x] = h1c;
#else
            p1.ha = fma2(wx2, shfl2(a, lr), fma2(wx0, shfl2(a, ll), a));
            p1.hb = fma2(wx2, shfl2(bb, lr), fma2(wx0, shfl2(bb, ll), bb));
            p1.hc = fma2(wx2, shfl2(cq, lr), fma2(wx0, shfl2(cq, ll), cq));
#endif
        }
        if (j >= 4) {
            const int gq = gy - 2;
            if (out_lane && j < kTrainBandH + 4 && (INNER || gq < H)) {
                float2 ga, gb, gc_;
#if DROSFM_SSIMT_HSMEM
                // cur of step KB is struct KB, p2 is struct (KB + 1) % 3
                const float2 cha = hrow[KB][0][threadIdx.x], chb = hrow[KB][1][threadIdx.x], chc = hrow[KB][2][threadIdx.x];
                const float2 p2a = hrow[(KB + 1) % 3][0][threadIdx.x], p2b = hrow[(KB + 1) % 3][1][threadIdx.x],
                             p2c = hrow[(KB + 1) % 3][2][threadIdx.x];
                if constexpr (INNER) {
                    ga = add2(h1a, add2(cha, p2a));
                    gb = add2(h1b, add2(chb, p2b));
                    gc_ = add2(h1c, add2(chc, p2c));
                } else {
                    const float2 wy0 = bc2(gq <= 0 ? 0.0f : (gq == 1 ? 2.0f : 1.0f));
                    const float2 wy2 = bc2(gq >= H - 1 ? 0.0f : (gq == H - 2 ? 2.0f : 1.0f));
                    ga = fma2(wy2, h1a, fma2(wy0, cha, p2a));
                    gb = fma2(wy2, h1b, fma2(wy0, chb, p2b));
                    gc_ = fma2(wy2, h1c, fma2(wy0, chc, p2c));
                }
                if constexpr (false) {
#else
                if constexpr (INNER) {      // unit weights: the same sums, bit for bit
                    ga = add2(p1.ha, add2(cur.ha, p2.ha));
                    gb = add2(p1.hb, add2(cur.hb, p2.hb));
                    gc_ = add2(p1.hc, add2(cur.hc, p2.hc));
                } else {
#endif
                    const float2 wy0 = bc2(gq <= 0 ? 0.0f : (gq == 1 ? 2.0f : 1.0f));
                    const float2 wy2 = bc2(gq >= H - 1 ? 0.0f : (gq == H - 2 ? 2.0f : 1.0f));
                    ga = fma2(wy2, p1.ha, fma2(wy0, cur.ha, p2.ha));
                    gb = fma2(wy2, p1.hb, fma2(wy0, cur.hb, p2.hb));
                    gc_ = fma2(wy2, p1.hc, fma2(wy0, cur.hc, p2.hc));
                }
                float2 gxv = fma2(gc_, bc2(p2.y), fma2(gb, p2.x, ga));
                const bool q0 = sv_prev == 2 * pr || sv_prev == 253, q1 = sv_prev == 2 * pr + 1 || sv_prev == 253;
                if (q0) {
                    const float df = p2.x.x - p2.y;
                    gxv.x += df == 0.0f ? 0.0f : __int_as_float(__float_as_int(kl1) ^ (__float_as_int(df) & 0x80000000));
                }
                if (q1) {
                    const float df = p2.x.y - p2.y;
                    gxv.y += df == 0.0f ? 0.0f : __int_as_float(__float_as_int(kl1) ^ (__float_as_int(df) & 0x80000000));
                }
                if constexpr (INNER) {
                    grow[0] = gxv.x;
                    grow[vstride] = gxv.y;
                } else {
                    const unsigned o = static_cast<unsigned>(gq * W);
                    gpl[o] = gxv.x;
                    gpl[o + vstride] = gxv.y;
                }
            }
        }
        if constexpr (INNER) {       // the output rows advance with the walk (written from step 3 / 4 on)
            srow += W;
            grow += W;
        }
        sv_prev = sv;
    };
    BwdRow2 r0, r1, r2;
    r0.ha = r0.hb = r0.hc = r1.ha = r1.hb = r1.hc = r2.ha = r2.hb = r2.hc = bc2(0.0f);
    r0.sx = r0.sxx = r0.sxy = r1.sx = r1.sxx = r1.sxy = bc2(0.0f);
    r0.sy = r0.syy = r1.sy = r1.syy = 0.0f;
    r0.x = r1.x = bc2(0.0f);
    r0.y = r1.y = 0.0f;
    auto walk = [&](auto inner) {
        if constexpr (decltype(inner)::value) {
            // step j loads row gy0 - 2 + j (+2 ahead), selects for row gy0 - 3 + j, writes gradients of row gy0 - 4 + j
            const unsigned first = static_cast<unsigned>((gy0 - 2) * W);
            xrow = xpl + first;
            yrow = ypl + first;
            arow = apl != nullptr ? apl + first : nullptr;
            grow = gpl + static_cast<unsigned>((gy0 - 4) * W);
            srow = spl != nullptr ? spl + static_cast<unsigned>((gy0 - 3) * W) : nullptr;
        }
        fetch(inner, gy0 - 2, f0);
        fetch(inner, gy0 - 1, f1);
#if DROSFM_SSIMT_PFD == 3
        fetch(inner, gy0, f2);
#endif
#pragma unroll 1
        for (int j = 0; j < kTrainBandH + 4; j += 3) {
#if DROSFM_SSIMT_PFD == 3
            // the row loaded for step j + 3 goes into the buffer step j has just emptied
            step(inner, std::integral_constant<int, 0>{}, r1, r2, r0, f0, f0, j);
            step(inner, std::integral_constant<int, 1>{}, r2, r0, r1, f1, f1, j + 1);
            step(inner, std::integral_constant<int, 2>{}, r0, r1, r2, f2, f2, j + 2);
#else
            step(inner, std::integral_constant<int, 0>{}, r1, r2, r0, f0, f2, j);
            step(inner, std::integral_constant<int, 1>{}, r2, r0, r1, f1, f0, j + 1);
            step(inner, std::integral_constant<int, 2>{}, r0, r1, r2, f2, f1, j + 2);
#endif
        }
    };
    constexpr int kSteps = (kTrainBandH + 4 + 2) / 3 * 3;      // the last row fetched is gy0 + kSteps - 1
#if DROSFM_SSIMT_INNER
    if (gy0 >= 2 && gy0 + kSteps - 1 + (DROSFM_SSIMT_PFD - 2) <= H - 1) walk(std::true_type{});
    else
#endif
        walk(std::false_type{});
    // loss: channel 0's lanes hold the per-pixel values of the band
    if (c == 0 && pr == 0) {
        const double part = warp_sum(static_cast<double>(local));
        if (lane == 0 && part != 0.0) atomicAdd(spread_acc(slot_at(ws, ip)), part);
    }
    Slot* ticket = slot_at(ws, n_preds);
    if (last_block(ticket, gridDim.x * gridDim.y, &flag) && threadIdx.x < 32)
        finish_weighted_means(ws, n_preds, pp.weight, static_cast<double>(B) * P * (use_min ? 1.0 : static_cast<double>(V)), ticket, loss);
}

static int check_photo(const float* image, const float* const* context, int n_views, const drosfm_photo_opts_t* opts,
                       int B, int H, int W) {
    DROSFM_REQUIRE(B >= 0 && H >= 0 && W >= 0, DROSFM_EINVAL, "photometric: negative dimension");
    DROSFM_REQUIRE(n_views >= 1 && n_views <= DROSFM_MAX_VIEWS, DROSFM_ERANGE, "photometric: n_views=%d outside [1,%d]",
                   n_views, DROSFM_MAX_VIEWS);
    DROSFM_REQUIRE(opts != nullptr, DROSFM_EINVAL, "photometric: NULL opts");
    DROSFM_REQUIRE(opts->padding == DROSFM_PAD_ZEROS || opts->padding == DROSFM_PAD_BORDER, DROSFM_EINVAL, "photometric: bad padding");
    DROSFM_REQUIRE(opts->reduce_op == DROSFM_REDUCE_MIN || opts->reduce_op == DROSFM_REDUCE_MEAN, DROSFM_EINVAL,
                   "photometric: bad reduce_op");
    DROSFM_REQUIRE(opts->ssim_w >= 0.0f, DROSFM_EINVAL, "photometric: negative ssim_loss_weight");
    DROSFM_REQUIRE(opts->ssim_w > 0.0f || !(opts->clip_loss > 0.0f && opts->reduce_op == DROSFM_REDUCE_MEAN), DROSFM_ENOTSUP,
                   "photometric: per-channel L1 maps (ssim_loss_weight == 0) with clip_loss > 0 need the 'min' reduce op");
    if (B == 0 || H * W == 0) return DROSFM_OK;
    DROSFM_REQUIRE(H >= 2 && W >= 2, DROSFM_ENOTSUP, "photometric: reflection padding needs H, W >= 2");
    DROSFM_REQUIRE(static_cast<long long>(H) * W < (1ll << 28) && B <= 4096, DROSFM_ERANGE, "photometric: dimension out of range");
    DROSFM_REQUIRE(image != nullptr && context != nullptr, DROSFM_EINVAL, "photometric: NULL image/context");
    for (int v = 0; v < n_views; ++v) DROSFM_REQUIRE(context[v] != nullptr, DROSFM_EINVAL, "photometric: context[%d] is NULL", v);
    return DROSFM_OK;
}

// Opt in to > 48 KB of dynamic shared memory (per device, done on first use; not a stream operation).
static int allow_big_smem() {
    static thread_local int done_for_device = -1;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess && dev != done_for_device) {
        e = cudaFuncSetAttribute(photometric_fwd_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(photometric_fwd_kernel<0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(photometric_fwd_kernel<1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(photometric_fwd_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(photometric_fwd_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmemBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(photometric_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSavedSmemBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(photometric_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmemBytes);
        if (e == cudaSuccess) done_for_device = dev;
    }
    if (e != cudaSuccess) {
        set_error("photometric: cannot configure shared memory: %s", cudaGetErrorString(e));
        return static_cast<int>(e);
    }
    return DROSFM_OK;
}

static float l1_weight(const drosfm_photo_opts_t* opts) {
    // (1 - ssim_loss_weight) is evaluated in double by the reference and rounded when it meets the fp32 tensor
    return static_cast<float>(1.0 - static_cast<double>(opts->ssim_w));
}

}  // namespace drosfm

using namespace drosfm;

extern "C" {

int drosfm_automask_fwd(const float* image, const float* const* context, int n_views, const drosfm_photo_opts_t* opts,
                        float* automask, int B, int H, int W, drosfm_stream_t stream) {
    if (int e = check_photo(image, context, n_views, opts, B, H, W)) return e;
    if (B == 0 || H * W == 0) return DROSFM_OK;
    DROSFM_REQUIRE(automask != nullptr, DROSFM_EINVAL, "automask_fwd: NULL output");
    PhotoPtrs pp{};
    for (int v = 0; v < n_views; ++v) pp.context[v] = context[v];
    drosfm_photo_opts_t o = *opts;
    o.reduce_op = DROSFM_REDUCE_MIN;
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    if (n_views == 2 && !(opts->clip_loss > 0.0f) && opts->ssim_w > 0.0f && static_cast<long long>(B) * 3 * H * W < (1ll << 31)) {
        // two views: the streaming SSIM walk of the loss forward on the un-warped sources (the tile kernel below executes
        // three times its instructions)
        const int nstrips = (W + kFwdStripW - 1) / kFwdStripW, nbands = (H + kFwdBandH - 1) / kFwdBandH;
        dim3 sgrid((nstrips * nbands + kSsimWarps - 1) / kSsimWarps, B);
        ssim_fwd_stream2_kernel<true><<<sgrid, kSsimThreads, 0, cs>>>(image, nullptr, pp, 1, nullptr, o, l1_weight(opts), nullptr, nullptr,
                                                                      nullptr, automask, B, H, W, nstrips, nbands);
        return launch_status("automask_fwd");
    }
    drosfm_cams_t none{};
    dim3 grid((W + FW - 1) / FW, (H + FH - 1) / FH, B);
    if (int e = allow_big_smem()) return e;
    ClipSlot* clip = nullptr;
    if (opts->clip_loss > 0.0f) {
        // statistics pass over the V un-warped maps, then their thresholds (slots 0..V-1 of the zero-filled scratch)
        DROSFM_REQUIRE(opts->clip_scratch != nullptr, DROSFM_EINVAL, "automask_fwd: clip_loss > 0 needs opts->clip_scratch");
        clip = reinterpret_cast<ClipSlot*>(opts->clip_scratch);
        photometric_fwd_kernel<1, false, true><<<grid, kFwdThreads, kFwdSmemBytes, cs>>>(
            image, pp, n_views, DROSFM_DEPTH, 1, none, nullptr, o, l1_weight(opts), nullptr, nullptr, nullptr, nullptr, nullptr, clip, B, H, W);
        if (int e = launch_status("automask_fwd (statistics)")) return e;
        clip_threshold_kernel<<<1, 64, 0, cs>>>(clip, n_views, static_cast<double>(B) * H * W * (opts->ssim_w > 0.0f ? 1.0 : 3.0), opts->clip_loss);
        if (int e = launch_status("automask_fwd (thresholds)")) return e;
    }
    photometric_fwd_kernel<1, false><<<grid, kFwdThreads, kFwdSmemBytes, cs>>>(
        image, pp, n_views, DROSFM_DEPTH, 1, none, nullptr, o, l1_weight(opts), nullptr, automask, nullptr, nullptr, nullptr, clip, B, H, W);
    return launch_status("automask_fwd");
}

static int fill_ptrs(PhotoPtrs& pp, const float* const* context, int n_views, const float* const* inv_depths, int n_preds,
                     const float* const* poses, float gamma) {
    DROSFM_REQUIRE(n_preds >= 1 && n_preds <= DROSFM_MAX_PREDS, DROSFM_ERANGE, "photometric: n_preds=%d outside [1,%d]",
                   n_preds, DROSFM_MAX_PREDS);
    DROSFM_REQUIRE(inv_depths != nullptr && poses != nullptr, DROSFM_EINVAL, "photometric: NULL depth/pose arrays");
    for (int v = 0; v < n_views; ++v) pp.context[v] = context[v];
    for (int i = 0; i < n_preds; ++i) {
        DROSFM_REQUIRE(inv_depths[i] != nullptr, DROSFM_EINVAL, "photometric: inv_depths[%d] is NULL", i);
        pp.inv_depth[i] = inv_depths[i];
        // gamma ** (n - i - 1) is a Python double that is rounded when it multiplies the fp32 loss
        double wgt = 1.0;
        for (int k = 0; k < n_preds - 1 - i; ++k) wgt *= static_cast<double>(gamma);
        pp.weight[i] = static_cast<float>(wgt);
    }
    for (int k = 0; k < n_views * n_preds; ++k) {
        DROSFM_REQUIRE(poses[k] != nullptr, DROSFM_EINVAL, "photometric: poses[%d] is NULL", k);
        pp.pose[k] = poses[k];
    }
    return DROSFM_OK;
}

int drosfm_photometric_fwd(const float* image, const float* const* context, int n_views, const float* const* inv_depths,
                           int depth_kind, int n_preds, const drosfm_cams_t* cams, const float* const* poses,
                           const float* automask, const drosfm_photo_opts_t* opts, uint8_t* sel, float* loss, void* ws,
                           float* warped_save, float* g_warped, int flags, int B, int H, int W, drosfm_stream_t stream) {
    if (int e = check_photo(image, context, n_views, opts, B, H, W)) return e;
    DROSFM_REQUIRE(B > 0 && H * W > 0, DROSFM_EINVAL, "photometric_fwd: empty batch (the mean over zero pixels is undefined)");
    DROSFM_REQUIRE(!(flags & DROSFM_PHOTO_WARPED_READY) || warped_save != nullptr, DROSFM_EINVAL,
                   "photometric_fwd: DROSFM_PHOTO_WARPED_READY without a warped buffer");
    DROSFM_REQUIRE(cams && cams->K && cams->Kref, DROSFM_EINVAL, "photometric_fwd: NULL cams");
    DROSFM_REQUIRE(cams->pose_kind == DROSFM_POSE_MAT4 || cams->pose_kind == DROSFM_POSE_EULER6, DROSFM_EINVAL,
                   "photometric_fwd: pose_kind must be MAT4 or EULER6");
    DROSFM_REQUIRE(loss != nullptr && ws != nullptr, DROSFM_EINVAL, "photometric_fwd: NULL loss/ws");
    DROSFM_REQUIRE(opts->reduce_op == DROSFM_REDUCE_MEAN || sel != nullptr, DROSFM_EINVAL, "photometric_fwd: min needs sel");
    DROSFM_REQUIRE(!(opts->clip_loss > 0.0f) || (warped_save == nullptr && opts->clip_scratch != nullptr && sel != nullptr), DROSFM_ENOTSUP,
                   "photometric_fwd: clip_loss > 0 runs on the fused path (warped_save == NULL) and needs opts->clip_scratch and sel");
    DROSFM_REQUIRE(opts->ssim_w > 0.0f || warped_save == nullptr, DROSFM_ENOTSUP,
                   "photometric_fwd: ssim_loss_weight == 0 runs on the fused path (warped_save == NULL)");
    DROSFM_REQUIRE(!(flags & DROSFM_PHOTO_FUSE_BWD) ||
                       (warped_save != nullptr && g_warped != nullptr && n_views % 2 == 0 && !(opts->clip_loss > 0.0f) &&
                        static_cast<long long>(B) * 3 * H * W < (1ll << 31) && static_cast<long long>(B) * n_preds <= 65535),
                   DROSFM_ENOTSUP, "photometric_fwd: DROSFM_PHOTO_FUSE_BWD needs the staged path (warped_save, g_warped), an even number of views, no clip");
    DROSFM_REQUIRE(!(opts->automask && opts->reduce_op != DROSFM_REDUCE_MIN), DROSFM_EINVAL,
                   "photometric_fwd: auto-masking needs the min reduce op");
    DROSFM_REQUIRE(!opts->automask || automask != nullptr, DROSFM_EINVAL, "photometric_fwd: automask map is NULL");
    PhotoPtrs pp{};
    if (int e = fill_ptrs(pp, context, n_views, inv_depths, n_preds, poses, opts->gamma)) return e;
    DROSFM_REQUIRE(static_cast<long long>(B) * n_preds * n_views <= 65535, DROSFM_ERANGE, "photometric_fwd: B * n_preds * n_views too large");
    dim3 grid((W + FW - 1) / FW, (H + FH - 1) / FH, B * n_preds);
    if (int e = allow_big_smem()) return e;
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    if (warped_save != nullptr) {
        if (!(flags & DROSFM_PHOTO_WARPED_READY)) {
            dim3 flat((W + 31) / 32, (H + kFwdTiles * kFlatTileH - 1) / (kFwdTiles * kFlatTileH), B * n_preds * n_views);
            warp_sources_kernel<false><<<flat, kFlatThreads, 0, cs>>>(pp, n_views, depth_kind, n_preds, *cams, opts->padding, nullptr,
                                                                      warped_save, B, H, W);
            if (int e = launch_status("photometric_fwd (warp_sources)")) return e;
        }
        if (flags & DROSFM_PHOTO_FUSE_BWD) {
            // training forward: loss, selection AND d loss / d warped (unscaled) in one pass over the warped copy
            const int nstrips = (W + kBwdStripW - 1) / kBwdStripW, nbands = (H + kTrainBandH - 1) / kTrainBandH;
            dim3 tgrid(nstrips * nbands, B * n_preds);
#define TRAIN(PAIRS_) ssim_train_stream2_kernel<PAIRS_><<<tgrid, kTrainThreads * PAIRS_, 0, cs>>>(                                     \
        image, warped_save, pp, n_preds, opts->automask ? automask : nullptr, *opts, l1_weight(opts), sel, loss, static_cast<Slot*>(ws), \
        g_warped, B, H, W, nstrips, nbands)
            if (n_views == 2) TRAIN(1);
            else if (n_views == 4) TRAIN(2);
            else if (n_views == 6) TRAIN(3);
            else TRAIN(4);
#undef TRAIN
        } else if (n_views <= 2 && static_cast<long long>(n_views) * B * 3 * H * W < (1ll << 31)) {
            const int nstrips = (W + kFwdStripW - 1) / kFwdStripW, nbands = (H + kFwdBandH - 1) / kFwdBandH;
            dim3 sgrid((nstrips * nbands + kSsimWarps - 1) / kSsimWarps, B * n_preds);
            if (n_views == 1)
                ssim_fwd_stream_kernel<1><<<sgrid, kSsimThreads, 0, cs>>>(image, warped_save, pp, n_preds, opts->automask ? automask : nullptr,
                                                                          *opts, l1_weight(opts), sel, loss, static_cast<Slot*>(ws), B, H,
                                                                          W, nstrips, nbands);
            else
                ssim_fwd_stream2_kernel<false><<<sgrid, kSsimThreads, 0, cs>>>(image, warped_save, pp, n_preds, opts->automask ? automask : nullptr,
                                                                               *opts, l1_weight(opts), sel, loss, static_cast<Slot*>(ws), nullptr,
                                                                               B, H, W, nstrips, nbands);
        } else {
            photometric_fwd_kernel<0, true><<<grid, kFwdThreads, kFwdSmemBytes, cs>>>(
                image, pp, n_views, depth_kind, n_preds, *cams, opts->automask ? automask : nullptr, *opts, l1_weight(opts), sel,
                nullptr, loss, static_cast<Slot*>(ws), warped_save, nullptr, B, H, W);
        }
    } else {
        ClipSlot* clip = nullptr;
        if (opts->clip_loss > 0.0f) {
            // statistics pass over the n_preds x V warped maps, then their thresholds (slots V.. of the scratch; the first V
            // belong to the un-warped maps of drosfm_automask_fwd)
            clip = reinterpret_cast<ClipSlot*>(opts->clip_scratch);
            photometric_fwd_kernel<0, false, true><<<grid, kFwdThreads, kFwdSmemBytes, cs>>>(
                image, pp, n_views, depth_kind, n_preds, *cams, nullptr, *opts, l1_weight(opts), nullptr, nullptr, nullptr, nullptr,
                nullptr, clip, B, H, W);
            if (int e = launch_status("photometric_fwd (statistics)")) return e;
            clip_threshold_kernel<<<(n_preds * n_views + 63) / 64, 64, 0, cs>>>(clip + n_views, n_preds * n_views,
                                                                                static_cast<double>(B) * H * W * (opts->ssim_w > 0.0f ? 1.0 : 3.0), opts->clip_loss);
            if (int e = launch_status("photometric_fwd (thresholds)")) return e;
        }
        photometric_fwd_kernel<0, false><<<grid, kFwdThreads, kFwdSmemBytes, cs>>>(
            image, pp, n_views, depth_kind, n_preds, *cams, opts->automask ? automask : nullptr, *opts, l1_weight(opts), sel,
            nullptr, loss, static_cast<Slot*>(ws), nullptr, clip, B, H, W);
    }
    return launch_status("photometric_fwd");
}

int drosfm_photometric_bwd(const float* g_loss, const float* image, const float* const* context, int n_views,
                           const float* const* inv_depths, int depth_kind, int n_preds, const drosfm_cams_t* cams,
                           const float* const* poses, const uint8_t* sel, const drosfm_photo_opts_t* opts,
                           float* const* g_inv_depths, float* const* g_poses, void* ws, const float* warped_save,
                           float* g_warped, int flags, int B, int H, int W, drosfm_stream_t stream) {
    if (int e = check_photo(image, context, n_views, opts, B, H, W)) return e;
    if (B == 0 || H * W == 0) return DROSFM_OK;
    DROSFM_REQUIRE(!(flags & DROSFM_PHOTO_NO_ADJOINT) || g_warped != nullptr, DROSFM_EINVAL,
                   "photometric_bwd: DROSFM_PHOTO_NO_ADJOINT needs the staged path (warped_save / g_warped)");
    DROSFM_REQUIRE(cams && cams->K && cams->Kref, DROSFM_EINVAL, "photometric_bwd: NULL cams");
    DROSFM_REQUIRE(cams->pose_kind == DROSFM_POSE_MAT4 || cams->pose_kind == DROSFM_POSE_EULER6, DROSFM_EINVAL,
                   "photometric_bwd: pose_kind must be MAT4 or EULER6");
    DROSFM_REQUIRE(g_loss != nullptr, DROSFM_EINVAL, "photometric_bwd: NULL g_loss");
    DROSFM_REQUIRE(opts->reduce_op == DROSFM_REDUCE_MEAN || sel != nullptr, DROSFM_EINVAL, "photometric_bwd: min needs sel");
    DROSFM_REQUIRE(!(opts->clip_loss > 0.0f) || (warped_save == nullptr && sel != nullptr), DROSFM_ENOTSUP,
                   "photometric_bwd: clip_loss > 0 runs on the fused path (warped_save == NULL) and needs the forward's sel");
    DROSFM_REQUIRE(opts->ssim_w > 0.0f || warped_save == nullptr, DROSFM_ENOTSUP,
                   "photometric_bwd: ssim_loss_weight == 0 runs on the fused path (warped_save == NULL)");
    DROSFM_REQUIRE((warped_save == nullptr) == (g_warped == nullptr), DROSFM_EINVAL,
                   "photometric_bwd: warped_save and g_warped go together (both NULL: fused path)");
    PhotoPtrs pp{};
    if (int e = fill_ptrs(pp, context, n_views, inv_depths, n_preds, poses, opts->gamma)) return e;
    PhotoGrads pg{};
    bool want_pose = false;
    for (int i = 0; i < n_preds; ++i) pg.g_inv_depth[i] = g_inv_depths ? g_inv_depths[i] : nullptr;
    for (int k = 0; k < n_views * n_preds; ++k) {
        pg.g_pose[k] = g_poses ? g_poses[k] : nullptr;
        want_pose |= pg.g_pose[k] != nullptr;
    }
    DROSFM_REQUIRE(!want_pose || ws != nullptr, DROSFM_EINVAL, "photometric_bwd: pose gradients need ws");
    DROSFM_REQUIRE(static_cast<long long>(B) * n_preds <= 65535, DROSFM_ERANGE, "photometric_bwd: B * n_preds too large");
    dim3 grid((W + IW - 1) / IW, (H + IH - 1) / IH, B * n_preds);
    if (int e = allow_big_smem()) return e;
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    if (warped_save != nullptr) {
        DROSFM_REQUIRE(static_cast<long long>(B) * n_preds * n_views * 3 <= 65535, DROSFM_ERANGE,
                       "photometric_bwd: B * n_preds * n_views too large");
        const int nstrips = (W + kBwdStripW - 1) / kBwdStripW, nbands = (H + kBwdBandH - 1) / kBwdBandH;
        dim3 sgrid((nstrips * nbands + kSsimWarps - 1) / kSsimWarps, B * n_preds * n_views * 3);
        if (n_views % 2 == 0 && static_cast<long long>(B) * 3 * H * W < (1ll << 31)) {
            dim3 pgrid((nstrips * nbands + kBwd2Warps - 1) / kBwd2Warps, B * n_preds * (n_views / 2) * 3);
            ssim_bwd_stream2_kernel<<<pgrid, kBwd2Threads, 0, cs>>>(g_loss, image, warped_save, pp, n_views, sel, *opts,
                                                                    l1_weight(opts), g_warped, B, H, W, nstrips, nbands);
        } else {
            ssim_bwd_stream_kernel<<<sgrid, kSsimThreads, 0, cs>>>(g_loss, image, warped_save, pp, n_views, sel, *opts,
                                                                   l1_weight(opts), g_warped, B, H, W, nstrips, nbands);
        }
        if (flags & DROSFM_PHOTO_NO_ADJOINT) return launch_status("photometric_bwd (window gradients)");
        DROSFM_REQUIRE(static_cast<long long>(n_views) * B * 3 * H * W < (1ll << 32), DROSFM_ERANGE,
                       "photometric_bwd: one prediction's warped views exceed 2^32 elements");
        if (int e = launch_status("photometric_bwd (window gradients)")) return e;
        DROSFM_REQUIRE(static_cast<long long>(B) * n_preds * n_views <= 65535, DROSFM_ERANGE, "photometric_bwd: B * n_preds * n_views too large");
        if (int e = launch_adjoint(pp, pg, n_views, depth_kind, n_preds, cams, opts->padding, nullptr, g_warped, nullptr,
                                   static_cast<Slot*>(ws), 0, B, H, W, cs)) return e;
    } else {
        photometric_bwd_kernel<false><<<grid, kBwdThreads, kBwdSmemBytes, cs>>>(
            g_loss, image, pp, n_views, depth_kind, n_preds, *cams, sel, *opts, l1_weight(opts), pg, static_cast<Slot*>(ws),
            nullptr, nullptr, B, H, W);
    }
    return launch_status("photometric_bwd");
}

int drosfm_warp_sources_fwd(const float* const* context, int n_views, const float* const* inv_depths, int depth_kind,
                            int n_preds, const drosfm_cams_t* cams, const float* const* poses, int padding, float* rgbx,
                            float* warped, int B, int H, int W, drosfm_stream_t stream) {
    DROSFM_REQUIRE(B >= 0 && H >= 0 && W >= 0, DROSFM_EINVAL, "warp_sources_fwd: negative dimension");
    DROSFM_REQUIRE(n_views >= 1 && n_views <= DROSFM_MAX_VIEWS, DROSFM_ERANGE, "warp_sources_fwd: n_views=%d outside [1,%d]",
                   n_views, DROSFM_MAX_VIEWS);
    DROSFM_REQUIRE(padding == DROSFM_PAD_ZEROS || padding == DROSFM_PAD_BORDER, DROSFM_EINVAL, "warp_sources_fwd: bad padding");
    if (B == 0 || H * W == 0) return DROSFM_OK;
    DROSFM_REQUIRE(context != nullptr && warped != nullptr, DROSFM_EINVAL, "warp_sources_fwd: NULL context/warped");
    for (int v = 0; v < n_views; ++v) DROSFM_REQUIRE(context[v] != nullptr, DROSFM_EINVAL, "warp_sources_fwd: context[%d] is NULL", v);
    DROSFM_REQUIRE(cams && cams->K && cams->Kref, DROSFM_EINVAL, "warp_sources_fwd: NULL cams");
    DROSFM_REQUIRE(cams->pose_kind == DROSFM_POSE_MAT4 || cams->pose_kind == DROSFM_POSE_EULER6, DROSFM_EINVAL,
                   "warp_sources_fwd: pose_kind must be MAT4 or EULER6");
    DROSFM_REQUIRE(static_cast<long long>(H) * W < (1ll << 28), DROSFM_ERANGE, "warp_sources_fwd: dimension out of range");
    PhotoPtrs pp{};
    if (int e = fill_ptrs(pp, context, n_views, inv_depths, n_preds, poses, 1.0f)) return e;
    DROSFM_REQUIRE(static_cast<long long>(B) * n_preds * n_views <= 65535, DROSFM_ERANGE, "warp_sources_fwd: B * n_preds * n_views too large");
    dim3 flat((W + 31) / 32, (H + kFwdTiles * kFlatTileH - 1) / (kFwdTiles * kFlatTileH), B * n_preds * n_views);
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    if (rgbx != nullptr) {
        DROSFM_REQUIRE(aligned16(rgbx), DROSFM_EALIGN, "warp_sources_fwd: rgbx must be 16-byte aligned");
        DROSFM_REQUIRE(B <= 65535, DROSFM_ERANGE, "warp_sources_fwd: B too large");
        dim3 pgrid((H * W + 255) / 256, B, n_views);
        pack_rgbx_kernel<<<pgrid, 256, 0, cs>>>(pp, B, H * W, rgbx);
        if (int e = launch_status("warp_sources_fwd (rgbx)")) return e;
        warp_sources_kernel<true><<<flat, kFlatThreads, 0, cs>>>(pp, n_views, depth_kind, n_preds, *cams, padding, rgbx, warped, B, H, W);
    } else {
        warp_sources_kernel<false><<<flat, kFlatThreads, 0, cs>>>(pp, n_views, depth_kind, n_preds, *cams, padding, nullptr, warped, B, H, W);
    }
    return launch_status("warp_sources_fwd");
}

int drosfm_warp_sources_bwd(const float* g_warped, const float* const* context, int n_views, const float* const* inv_depths,
                            int depth_kind, int n_preds, const drosfm_cams_t* cams, const float* const* poses, int padding,
                            const float* rgbx, const float* g_scale, float* const* g_inv_depths, float* const* g_poses, void* ws,
                            int accumulate, int B, int H, int W, drosfm_stream_t stream) {
    DROSFM_REQUIRE(B >= 0 && H >= 0 && W >= 0, DROSFM_EINVAL, "warp_sources_bwd: negative dimension");
    DROSFM_REQUIRE(n_views >= 1 && n_views <= DROSFM_MAX_VIEWS, DROSFM_ERANGE, "warp_sources_bwd: n_views=%d outside [1,%d]",
                   n_views, DROSFM_MAX_VIEWS);
    DROSFM_REQUIRE(padding == DROSFM_PAD_ZEROS || padding == DROSFM_PAD_BORDER, DROSFM_EINVAL, "warp_sources_bwd: bad padding");
    if (B == 0 || H * W == 0) return DROSFM_OK;
    DROSFM_REQUIRE(context != nullptr && g_warped != nullptr, DROSFM_EINVAL, "warp_sources_bwd: NULL context/g_warped");
    for (int v = 0; v < n_views; ++v) DROSFM_REQUIRE(context[v] != nullptr, DROSFM_EINVAL, "warp_sources_bwd: context[%d] is NULL", v);
    DROSFM_REQUIRE(cams && cams->K && cams->Kref, DROSFM_EINVAL, "warp_sources_bwd: NULL cams");
    DROSFM_REQUIRE(cams->pose_kind == DROSFM_POSE_MAT4 || cams->pose_kind == DROSFM_POSE_EULER6, DROSFM_EINVAL,
                   "warp_sources_bwd: pose_kind must be MAT4 or EULER6");
    DROSFM_REQUIRE(static_cast<long long>(H) * W < (1ll << 28), DROSFM_ERANGE, "warp_sources_bwd: dimension out of range");
    PhotoPtrs pp{};
    if (int e = fill_ptrs(pp, context, n_views, inv_depths, n_preds, poses, 1.0f)) return e;
    PhotoGrads pg{};
    bool want_pose = false;
    for (int i = 0; i < n_preds; ++i) pg.g_inv_depth[i] = g_inv_depths ? g_inv_depths[i] : nullptr;
    for (int k = 0; k < n_views * n_preds; ++k) {
        pg.g_pose[k] = g_poses ? g_poses[k] : nullptr;
        want_pose |= pg.g_pose[k] != nullptr;
    }
    DROSFM_REQUIRE(!want_pose || ws != nullptr, DROSFM_EINVAL, "warp_sources_bwd: pose gradients need ws");
    DROSFM_REQUIRE(static_cast<long long>(B) * n_preds <= 65535, DROSFM_ERANGE, "warp_sources_bwd: B * n_preds too large");
    DROSFM_REQUIRE(static_cast<long long>(n_views) * B * 3 * H * W < (1ll << 32), DROSFM_ERANGE,
                   "warp_sources_bwd: one prediction's warped views exceed 2^32 elements");
    DROSFM_REQUIRE(static_cast<long long>(B) * n_preds * n_views <= 65535, DROSFM_ERANGE, "warp_sources_bwd: B * n_preds * n_views too large");
    DROSFM_REQUIRE(rgbx == nullptr || aligned16(rgbx), DROSFM_EALIGN, "warp_sources_bwd: rgbx must be 16-byte aligned");
    if (int e = launch_adjoint(pp, pg, n_views, depth_kind, n_preds, cams, padding, rgbx, g_warped, g_scale, static_cast<Slot*>(ws), accumulate,
                               B, H, W, static_cast<cudaStream_t>(stream))) return e;
    return launch_status("warp_sources_bwd");
}

}  // extern "C"
