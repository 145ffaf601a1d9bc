// Scalar-loss kernels next to the photometric term.
//
//   drosfm_smoothness_*   calc_smoothness_loss (dro_sfm/losses/multiview_photometric_loss_mf.py:273-299),
//                         calc_smoothness / inv_depths_normalize (dro_sfm/utils/depth.py:147-199),
//                         gradient_x / gradient_y (dro_sfm/utils/image.py:134-162)
//   drosfm_reproj_loss_*  SupervisedDepthPoseLoss.get_ref_coords / calc_pose_loss
//                         (dro_sfm/losses/supervised_loss.py:279-325)
//
// Smoothness: the edge weights exp(-mean_c |dI|) do not depend on the prediction, so one pass over the
// image serves all n predictions (12 + 4n bytes per pixel instead of n * 16).  The loss is homogeneous
// of degree one in the mean-normalised inverse depth, which turns the gradient through the per-sample
// mean into  -L_{i,b} / (mean * P)  (Euler's theorem) and makes the backward a single streaming pass.
//
// Reprojection loss: the GT-pose coordinates are identical for every prediction; one pass over the GT
// depth evaluates all V + V*n projections per pixel (4 bytes per pixel of HBM traffic).
#include "common.cuh"

namespace drosfm {

constexpr int kLossThreads = 256;

struct DepthList {
    const float* d[DROSFM_MAX_PREDS];
};
struct DepthGrads {
    float* g[DROSFM_MAX_PREDS];
};

// ------------------------------------------------------------------------------------------
// smoothness
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float edge_weight(const float* __restrict__ img, int P, int p, int q) {
    const float a = fabsf(__ldg(img + p) - __ldg(img + q)) + fabsf(__ldg(img + P + p) - __ldg(img + P + q)) +
                    fabsf(__ldg(img + 2 * P + p) - __ldg(img + 2 * P + q));
    // a / 3, correctly rounded, without the division's ~15 instructions (q = RN(a/3) from one Newton-style correction)
    const float r = 1.0f / 3.0f, t = __fmul_rn(a, r);
    return expf(-__fmaf_rn(__fmaf_rn(-3.0f, t, a), r, t));
}

// sums of |dx|*wx and |dy|*wy per (prediction, sample); the last block folds them into the loss.
// NP: n_preds rounded up to a multiple of four (the per-prediction code is unrolled NP times; slots beyond n_preds are
// predicated off but still issue).
// ONE pass: the loss is homogeneous of degree one in the mean-normalised inverse depth, |d_p/m - d_q/m| = |d_p - d_q| / m, so
// the per-(prediction, sample) mean m is accumulated NEXT TO the un-normalised edge-weighted sums and divides them in the
// finisher -- no separate mean pass over the n maps (the products differ from the reference's by one rounding per term).
template <int NP>
__global__ void __launch_bounds__(kLossThreads)
smooth_fwd_kernel(const float* __restrict__ image, const __grid_constant__ DepthList dl, int n_preds, float weight,
                  float* __restrict__ stats, float* __restrict__ loss, Slot* ws, float* __restrict__ edge_w, int B, int H, int W) {
    __shared__ float red[3 * NP][kLossThreads / 32];
    __shared__ int flag;
    const int b = blockIdx.y, P = H * W;
    const float* img = image + static_cast<size_t>(b) * 3 * P;
    float sx[NP], sy[NP], sd[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) sx[i] = sy[i] = sd[i] = 0.0f;
    const int p0 = blockIdx.x * kLossThreads + threadIdx.x;
    int y = p0 / W, x = p0 - y * W;
    const int stride = gridDim.x * kLossThreads, sy_ = stride / W, sx_ = stride - sy_ * W;
    for (int p = p0; p < P; p += stride) {
        const bool hx = x + 1 < W, hy = y + 1 < H;
        const int px = hx ? p + 1 : p, py = hy ? p + W : p;
        // all 3 n depth loads of the pixel are requested before the first is consumed
        float dc[NP], dx[NP], dy[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            if (i < n_preds) {
                const float* d = dl.d[i] + static_cast<size_t>(b) * P;
                dc[i] = __ldg(d + p);
                dx[i] = __ldg(d + px);
                dy[i] = __ldg(d + py);
            }
        }
        const float wx = hx ? edge_weight(img, P, p, px) : 0.0f;
        const float wy = hy ? edge_weight(img, P, p, py) : 0.0f;
        if (edge_w != nullptr) {      // kept for the backward pass: the weights towards the right / lower neighbour
            edge_w[static_cast<size_t>(2 * b) * P + p] = wx;
            edge_w[static_cast<size_t>(2 * b + 1) * P + p] = wy;
        }
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            if (i < n_preds) {
                sd[i] += dc[i];
                sx[i] += fabsf((dc[i] - dx[i]) * wx);
                sy[i] += fabsf((dc[i] - dy[i]) * wy);
            }
        }
        // next pixel of this thread: (x, y) advance without a division
        x += sx_;
        y += sy_;
        if (x >= W) { x -= W; ++y; }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        if (i < n_preds) {
            const float a = warp_sum(sx[i]), c = warp_sum(sy[i]), m = warp_sum(sd[i]);
            if (lane == 0) { red[3 * i][wid] = a; red[3 * i + 1][wid] = c; red[3 * i + 2][wid] = m; }
        }
    }
    __syncthreads();
    if (threadIdx.x < 3 * n_preds) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < kLossThreads / 32; ++k) t += red[threadIdx.x][k];
        const int i = threadIdx.x / 3;
        if (t != 0.0) atomicAdd(spread_acc(slot_at(ws, i * B + b)) + (threadIdx.x - 3 * i), t);
    }
    Slot* ticket = slot_at(ws, n_preds * B);
    if (last_block(ticket, gridDim.x * gridDim.y, &flag)) {
        // every thread of the last block collects (prediction, sample) pairs, so the 3 n B accumulators are read with
        // one round trip; the per-prediction totals meet in shared memory (fp64), thread 0 folds them into the loss
        __shared__ double tot[2 * DROSFM_MAX_PREDS];
        if (threadIdx.x < 2 * DROSFM_MAX_PREDS) tot[threadIdx.x] = 0.0;
        __syncthreads();
        for (int k = threadIdx.x; k < n_preds * B; k += kLossThreads) {
            Slot* s = slot_at(ws, k);                          // k = i * B + bb
            const double axr = take_acc(s, 0), ayr = take_acc(s, 1);
            const float mean = static_cast<float>(take_acc(s, 2) / static_cast<double>(P));
            const double rm = static_cast<double>(1.0f / fmaxf(mean, 1e-6f));         // inv_depths_normalize (utils/depth.py:147-166)
            const double ax = axr * rm, ay = ayr * rm;
            stats[k * 4 + 0] = mean;
            stats[k * 4 + 1] = static_cast<float>(ax);
            stats[k * 4 + 2] = static_cast<float>(ay);
            atomicAdd(&tot[2 * (k / B)], ax);
            atomicAdd(&tot[2 * (k / B) + 1], ay);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const double nx = static_cast<double>(B) * H * (W - 1), ny = static_cast<double>(B) * (H - 1) * W;
            double total = 0.0, pw = 1.0;
            for (int i = 0; i < n_preds; ++i) {
                total += ((nx > 0 ? tot[2 * i] / nx : 0.0) + (ny > 0 ? tot[2 * i + 1] / ny : 0.0)) / pw;
                pw *= 2.0;
            }
            ticket->ticket = 0ull;
            *loss = static_cast<float>(static_cast<double>(weight) * (total / n_preds));
        }
    }
}

__device__ __forceinline__ float sgn(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }

__global__ void __launch_bounds__(kLossThreads)
smooth_bwd_kernel(const float* __restrict__ g_loss, const float* __restrict__ image, const __grid_constant__ DepthList dl,
                  int n_preds, float weight, const float* __restrict__ stats, const __grid_constant__ DepthGrads dg,
                  int accumulate, const float* __restrict__ edge_w, int B, int H, int W) {
    // per (prediction, sample) constants, computed once per BLOCK (they hold all the divisions of the formula):
    // 1/mean, kx/mean, ky/mean and the gradient through the mean
    __shared__ float4 cst[DROSFM_MAX_PREDS];
    const int b = blockIdx.y, P = H * W;
    if (threadIdx.x < n_preds) {
        const int i = threadIdx.x;
        const float* st = stats + (i * B + b) * 4;
        const float g0 = __ldg(g_loss) * weight / static_cast<float>(n_preds);
        const float nx = static_cast<float>(B) * H * (W - 1), ny = static_cast<float>(B) * (H - 1) * W;
        const float pw = static_cast<float>(1u << i);
        const float mean = st[0], rm = 1.0f / fmaxf(mean, 1e-6f);
        const float kx = nx > 0.0f ? g0 / (pw * nx) : 0.0f, ky = ny > 0.0f ? g0 / (pw * ny) : 0.0f;
        const float corr = mean >= 1e-6f ? (kx * st[1] + ky * st[2]) * rm * (1.0f / static_cast<float>(P)) : 0.0f;
        cst[i] = make_float4(rm, kx, ky, corr);
    }
    __syncthreads();
    const int p = blockIdx.x * kLossThreads + threadIdx.x;
    if (p >= P) return;
    const float* img = image + static_cast<size_t>(b) * 3 * P;
    const int y = p / W, x = p - y * W;
    const bool xr = x + 1 < W, xl = x > 0, yd = y + 1 < H, yu = y > 0;
    const int pr = xr ? p + 1 : p, pl = xl ? p - 1 : p, pd = yd ? p + W : p, pu = yu ? p - W : p;
    const size_t o = static_cast<size_t>(b) * P + p;
    // predictions in groups of four: the 5 (+1) loads of every prediction of a group are requested before any is consumed
    constexpr int kGroup = 4;
    float vc[kGroup], vr[kGroup], vl[kGroup], vd[kGroup], vu[kGroup], vo[kGroup];
    auto fetch = [&](int i0) {
#pragma unroll
        for (int k = 0; k < kGroup; ++k) {
            const int i = i0 + k;
            vo[k] = 0.0f;
            if (i < n_preds && dg.g[i] != nullptr) {
                const float* d = dl.d[i] + static_cast<size_t>(b) * P;
                vc[k] = __ldg(d + p); vr[k] = __ldg(d + pr); vl[k] = __ldg(d + pl); vd[k] = __ldg(d + pd); vu[k] = __ldg(d + pu);
                if (accumulate == 1) vo[k] = dg.g[i][o];
            }
        }
    };
    fetch(0);
    float w_r, w_l, w_d, w_u;
    if (edge_w != nullptr) {          // the forward pass left the weights behind: 4 loads instead of 24 loads + 4 exponentials
        const float* ex = edge_w + static_cast<size_t>(2 * b) * P;
        const float* ey = ex + P;
        w_r = xr ? __ldg(ex + p) : 0.0f;
        w_l = xl ? __ldg(ex + pl) : 0.0f;
        w_d = yd ? __ldg(ey + p) : 0.0f;
        w_u = yu ? __ldg(ey + pu) : 0.0f;
    } else {
        w_r = xr ? edge_weight(img, P, p, pr) : 0.0f;
        w_l = xl ? edge_weight(img, P, pl, p) : 0.0f;
        w_d = yd ? edge_weight(img, P, p, pd) : 0.0f;
        w_u = yu ? edge_weight(img, P, pu, p) : 0.0f;
    }
    for (int i0 = 0; i0 < n_preds; i0 += kGroup) {
#pragma unroll
        for (int k = 0; k < kGroup; ++k) {
            const int i = i0 + k;
            if (i < n_preds && dg.g[i] != nullptr) {
                const float4 c = cst[i];
                const float rm = c.x, kx = c.y, ky = c.z;
                const float dc = vc[k] * rm;
                // d L / d dn[p], dn = d * (1/mean) (the normalised values are formed exactly as in the forward pass)
                float h = 0.0f;
                h += kx * w_r * sgn(dc - vr[k] * rm);
                h -= kx * w_l * sgn(vl[k] * rm - dc);
                h += ky * w_d * sgn(dc - vd[k] * rm);
                h -= ky * w_u * sgn(vu[k] * rm - dc);
                const float g = h * rm - c.w;
                if (accumulate == 2) {      // another kernel may be adding into the same maps right now: fire-and-forget atomic
                    if (g != 0.0f) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dg.g[i] + o), "f"(g) : "memory");
                } else {
                    dg.g[i][o] = accumulate ? vo[k] + g : g;
                }
            }
        }
        if (i0 + kGroup < n_preds) fetch(i0 + kGroup);
    }
}

// ------------------------------------------------------------------------------------------
// reprojection pose loss
// ------------------------------------------------------------------------------------------
struct ReprojPtrs {
    const float* gt[DROSFM_MAX_VIEWS];
    const float* pred[DROSFM_MAX_VIEWS * DROSFM_MAX_PREDS];
    float* g_pred[DROSFM_MAX_VIEWS * DROSFM_MAX_PREDS];
    float weight[DROSFM_MAX_PREDS];
};

// source-camera part of the setup only (the target side is shared by all poses of a sample)
__device__ __forceinline__ void setup_src(const drosfm_cams_t& c, const float* pose, int b, Cam& base, Cam& out) {
    out = base;
    load_pose(pose, c.pose_kind, b, out.T, out.trig);
#pragma unroll
    for (int k = 0; k < 3; ++k)
        out.c[k] = out.T[4 * k] * out.Rt[3] + out.T[4 * k + 1] * out.Rt[7] + out.T[4 * k + 2] * out.Rt[11] + out.T[4 * k + 3];
}

// One block = a strip of pixels of one sample.  The predictions of a view are processed in groups of four: one pass over
// the block's pixels evaluates the GT projection once and the four predicted projections next to it (the GT chain, the
// camera set-up barrier and the block reductions are paid per group, not per prediction).
// MODE 0: forward sums per prediction.  MODE 1: backward, pose gradients per (view, prediction).
#ifndef DROSFM_REPROJ_BWD_GROUP
#define DROSFM_REPROJ_BWD_GROUP 2      // the backward keeps 12 pose-gradient sums per prediction of the group in registers
#endif
constexpr int kReprojGroupMax = 4;

template <int MODE>
__global__ void __launch_bounds__(kLossThreads)
reproj_kernel(const float* __restrict__ g_loss, const float* __restrict__ depth, int depth_kind, drosfm_cams_t cams,
              const __grid_constant__ ReprojPtrs rp, int V, int n_preds, float min_depth, float max_depth, float wsum,
              float* __restrict__ loss, Slot* ws, int B, int H, int W) {
    constexpr int kReprojGroup = MODE == 0 ? kReprojGroupMax : DROSFM_REPROJ_BWD_GROUP;
    __shared__ Cam base, cgt, cpr[kReprojGroup];
    __shared__ double red[12 * (kLossThreads / 32)];
    __shared__ int flag;
    const int b = blockIdx.y, P = H * W;
    const float wm1 = static_cast<float>(W - 1), hm1 = static_cast<float>(H - 1);
    if (threadIdx.x == 0) setup_cam(cams, nullptr, b, base);
    __syncthreads();
    const float dmax = max_depth / 4.0f;
    const int stride = gridDim.x * kLossThreads;
    const float gl = MODE == 1 ? __ldg(g_loss) : 0.0f;
    for (int v = 0; v < V; ++v) {
        for (int i0 = 0; i0 < n_preds; i0 += kReprojGroup) {
            const int ng = min(kReprojGroup, n_preds - i0);
            __syncthreads();
            // lane 0 of warps 0..ng sets up the GT camera / one predicted camera each
            if ((threadIdx.x & 31) == 0) {
                const int k = threadIdx.x >> 5;
                if (k == 0) setup_src(cams, rp.gt[v], b, base, cgt);
                else if (k <= ng) setup_src(cams, rp.pred[v * n_preds + i0 + k - 1], b, base, cpr[k - 1]);
            }
            __syncthreads();
            float acc[kReprojGroup];
            float gT[kReprojGroup][12];
            float kscale[kReprojGroup];
#pragma unroll
            for (int k = 0; k < kReprojGroup; ++k) {
                acc[k] = 0.0f;
#pragma unroll
                for (int q = 0; q < 12; ++q) gT[k][q] = 0.0f;
                // d loss / d |diff| for this (view, prediction): w_i / (V * wsum * numel)
                kscale[k] = (MODE == 1 && k < ng) ? gl * rp.weight[i0 + k] /
                                                        (static_cast<float>(V) * wsum * 2.0f * static_cast<float>(B) * static_cast<float>(P))
                                                  : 0.0f;
            }
            for (int p = blockIdx.x * kLossThreads + threadIdx.x; p < P; p += stride) {
                const int y = p / W, x = p - y * W;
                const float d = to_depth(__ldg(depth + static_cast<size_t>(b) * P + p), depth_kind);
                if (!(d > min_depth && d < dmax)) continue;
                Warp wg;
                warp_pixel<true>(cgt, x, y, d, wm1, hm1, true, wg);
                const bool gu_in = wg.p.u >= -1.0f && wg.p.u <= 1.0f, gv_in = wg.p.v >= -1.0f && wg.p.v <= 1.0f;
#pragma unroll
                for (int k = 0; k < kReprojGroup; ++k) {
                    if (k < ng) {
                        Warp wp;
                        warp_pixel<true>(cpr[k], x, y, d, wm1, hm1, true, wp);
                        const bool vu = gu_in && wp.p.u >= -1.0f && wp.p.u <= 1.0f;
                        const bool vv = gv_in && wp.p.v >= -1.0f && wp.p.v <= 1.0f;
                        const float du = wp.p.u - wg.p.u, dv = wp.p.v - wg.p.v;
                        if (MODE == 0) {
                            if (vu) acc[k] += fminf(fabsf(du), 1.0f);
                            if (vv) acc[k] += fminf(fabsf(dv), 1.0f);
                        } else {
                            const float gu = (vu && fabsf(du) <= 1.0f) ? kscale[k] * sgn(du) : 0.0f;
                            const float gv = (vv && fabsf(dv) <= 1.0f) ? kscale[k] * sgn(dv) : 0.0f;
                            if (gu != 0.0f || gv != 0.0f) warp_pixel_adjoint(cpr[k], wp, d, wm1, hm1, true, gu, gv, gT[k]);
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < kReprojGroup; ++k) {
                if (k >= ng) continue;                                  // block-uniform
                const int i = i0 + k;
                if (MODE == 0) {
                    block_accumulate<1>(&acc[k], red, spread_acc(slot_at(ws, i)));
                } else if (rp.g_pred[v * n_preds + i] != nullptr) {
                    Slot* slot = slot_at(ws, (v * n_preds + i) * B + b);
                    block_accumulate12(gT[k], red, spread_acc(slot));
                    if (last_block(slot, gridDim.x, &flag) && threadIdx.x < 32) {
                        const bool eul = cams.pose_kind == DROSFM_POSE_EULER6;
                        finish_pose_grad_warp(slot, cams.pose_kind, eul ? rp.pred[v * n_preds + i] + b * 6 : nullptr,
                                              rp.g_pred[v * n_preds + i] + b * (eul ? 6 : 16));
                    }
                }
            }
        }
    }
    if (MODE == 0) {
        Slot* ticket = slot_at(ws, n_preds);
        if (last_block(ticket, gridDim.x * gridDim.y, &flag) && threadIdx.x < 32) {
            // lane i collects prediction i (all loads in flight together), lane 0 adds them in order
            const int lane = threadIdx.x;
            const double numel = 2.0 * static_cast<double>(B) * P;
            double term = 0.0;
            if (lane < n_preds) term = static_cast<double>(rp.weight[lane]) * (take_acc(slot_at(ws, lane), 0) / numel / V);
            double total = 0.0;
            for (int i = 0; i < n_preds; ++i) total += __shfl_sync(0xffffffffu, term, i);
            if (lane == 0) {
                ticket->ticket = 0ull;
                *loss = static_cast<float>(total / static_cast<double>(wsum));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// supervised depth loss: gamma-weighted masked L1 on inverse depth (supervised_loss.py:244-277)
// ------------------------------------------------------------------------------------------
struct SupPtrs {
    const float* d[DROSFM_MAX_PREDS];
    float* g[DROSFM_MAX_PREDS];
    float weight[DROSFM_MAX_PREDS];
};

// MODE 0: loss = sum_i w_i * mean(valid * |gt - d_i|) / sum_i w_i;  MODE 1: gradients w.r.t. every d_i.
// NP: n_preds rounded up to a multiple of four (unrolled).  VEC: elements per thread and step (4: 128-bit accesses).  All
// loads of a step are requested before anything is stored (a store between them would serialise the memory round trips).
template <int MODE, int NP, int VEC>
__global__ void __launch_bounds__(kLossThreads)
sup_depth_kernel(const float* __restrict__ g_loss, const float* __restrict__ gt, const __grid_constant__ SupPtrs sp, int n_preds,
                 float lo, float hi, float wsum, float* __restrict__ loss, Slot* ws, long long N) {
    __shared__ double red[NP][kLossThreads / 32];
    __shared__ int flag;
    float acc[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) acc[i] = 0.0f;
    const float gscale = MODE == 1 ? __ldg(g_loss) / (wsum * static_cast<float>(N)) : 0.0f;
    const long long steps = N / VEC;                 // the host picks VEC = 4 only when N is a multiple of four
    for (long long q = static_cast<long long>(blockIdx.x) * kLossThreads + threadIdx.x; q < steps;
         q += static_cast<long long>(gridDim.x) * kLossThreads) {
        float t[VEC], d[NP][VEC];
        if constexpr (VEC == 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(gt) + q);
            t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
        } else {
            t[0] = __ldg(gt + q);
        }
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            if (i < n_preds) {
                if constexpr (VEC == 4) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(sp.d[i]) + q);
                    d[i][0] = v.x; d[i][1] = v.y; d[i][2] = v.z; d[i][3] = v.w;
                } else {
                    d[i][0] = __ldg(sp.d[i] + q);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            if (i < n_preds) {
                float g[VEC];
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    const bool valid = t[e] > lo && t[e] < hi;
                    const float df = t[e] - d[i][e];
                    if (MODE == 0) {
                        if (valid) acc[i] += fabsf(df);
                    } else {
                        // d|gt - d|/dd = -sign(gt - d)
                        g[e] = valid ? -gscale * sp.weight[i] * (df > 0.0f ? 1.0f : (df < 0.0f ? -1.0f : 0.0f)) : 0.0f;
                    }
                }
                if (MODE == 1 && sp.g[i] != nullptr) {
                    if constexpr (VEC == 4) reinterpret_cast<float4*>(sp.g[i])[q] = make_float4(g[0], g[1], g[2], g[3]);
                    else sp.g[i][q] = g[0];
                }
            }
        }
    }
    if (MODE == 1) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        if (i < n_preds) {
            const double s = warp_sum(static_cast<double>(acc[i]));
            if (lane == 0) red[i][wid] = s;
        }
    }
    __syncthreads();
    if (threadIdx.x < n_preds) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < kLossThreads / 32; ++k) t += red[threadIdx.x][k];
        if (t != 0.0) atomicAdd(spread_acc(slot_at(ws, threadIdx.x)), t);
    }
    Slot* ticket = slot_at(ws, n_preds);
    if (last_block(ticket, gridDim.x, &flag) && threadIdx.x < 32) {
        const int lane = threadIdx.x;
        double term = 0.0;
        if (lane < n_preds)
            term = static_cast<double>(sp.weight[lane]) *
                   static_cast<double>(static_cast<float>(take_acc(slot_at(ws, lane), 0) / static_cast<double>(N)));
        double total = 0.0;
        for (int i = 0; i < n_preds; ++i) total += __shfl_sync(0xffffffffu, term, i);
        if (lane == 0) {
            ticket->ticket = 0ull;
            *loss = static_cast<float>(total / static_cast<double>(wsum));
        }
    }
}

// launches sup_depth_kernel<MODE, NP, VEC> for the prediction count and the alignment of the maps
template <int MODE>
static void launch_sup_depth(unsigned blocks, cudaStream_t cs, bool vec4, const float* g_loss, const float* gt, const SupPtrs& sp,
                             int n_preds, float lo, float hi, float wsum, float* loss, Slot* ws, long long N) {
#define SUP(NP_, VEC_) sup_depth_kernel<MODE, NP_, VEC_><<<blocks, kLossThreads, 0, cs>>>(g_loss, gt, sp, n_preds, lo, hi, wsum, loss, ws, N)
#define SUP_NP(VEC_)                                  \
    do {                                              \
        if (n_preds <= 4) SUP(4, VEC_);               \
        else if (n_preds <= 8) SUP(8, VEC_);          \
        else if (n_preds <= 12) SUP(12, VEC_);        \
        else SUP(16, VEC_);                           \
    } while (0)
    if (vec4) SUP_NP(4);
    else SUP_NP(1);
#undef SUP_NP
#undef SUP
}

static bool sup_vec4(const float* gt, const SupPtrs& sp, int n_preds, long long N) {
    bool ok = (N % 4) == 0 && aligned16(gt);
    for (int i = 0; i < n_preds; ++i) ok = ok && aligned16(sp.d[i]) && (sp.g[i] == nullptr || aligned16(sp.g[i]));
    return ok;
}

static int fill_sup(SupPtrs& sp, float& wsum, const float* const* inv_depths, float* const* g, int n_preds, float gamma) {
    DROSFM_REQUIRE(n_preds >= 1 && n_preds <= DROSFM_MAX_PREDS, DROSFM_ERANGE, "sup_depth_loss: n_preds=%d outside [1,%d]", n_preds,
                   DROSFM_MAX_PREDS);
    DROSFM_REQUIRE(inv_depths != nullptr, DROSFM_EINVAL, "sup_depth_loss: NULL prediction array");
    double sum = 0.0;
    for (int i = 0; i < n_preds; ++i) {
        DROSFM_REQUIRE(inv_depths[i] != nullptr, DROSFM_EINVAL, "sup_depth_loss: inv_depths[%d] is NULL", i);
        sp.d[i] = inv_depths[i];
        sp.g[i] = g ? g[i] : nullptr;
        double wgt = 1.0;
        for (int k = 0; k < n_preds - 1 - i; ++k) wgt *= static_cast<double>(gamma);
        sp.weight[i] = static_cast<float>(wgt);
        sum += wgt;
    }
    wsum = static_cast<float>(sum);
    return DROSFM_OK;
}

static int strip_blocks(int P, int B) {
    int need = (P + kLossThreads - 1) / kLossThreads;
    int cap = (kNumSMs * 4 + B - 1) / (B > 0 ? B : 1);
    if (cap < 1) cap = 1;
    if (need < 1) need = 1;
    return need < cap ? need : cap;
}

static int fill_reproj(ReprojPtrs& rp, float& wsum, const float* const* gt_poses, const float* const* pred_poses,
                       int n_views, int n_preds, float gamma) {
    DROSFM_REQUIRE(n_views >= 1 && n_views <= DROSFM_MAX_VIEWS, DROSFM_ERANGE, "reproj_loss: n_views=%d outside [1,%d]", n_views,
                   DROSFM_MAX_VIEWS);
    DROSFM_REQUIRE(n_preds >= 1 && n_preds <= DROSFM_MAX_PREDS, DROSFM_ERANGE, "reproj_loss: n_preds=%d outside [1,%d]", n_preds,
                   DROSFM_MAX_PREDS);
    DROSFM_REQUIRE(gt_poses && pred_poses, DROSFM_EINVAL, "reproj_loss: NULL pose arrays");
    for (int v = 0; v < n_views; ++v) {
        DROSFM_REQUIRE(gt_poses[v] != nullptr, DROSFM_EINVAL, "reproj_loss: gt_poses[%d] is NULL", v);
        rp.gt[v] = gt_poses[v];
    }
    for (int k = 0; k < n_views * n_preds; ++k) {
        DROSFM_REQUIRE(pred_poses[k] != nullptr, DROSFM_EINVAL, "reproj_loss: pred_poses[%d] is NULL", k);
        rp.pred[k] = pred_poses[k];
    }
    double sum = 0.0;
    for (int i = 0; i < n_preds; ++i) {
        double wgt = 1.0;
        for (int k = 0; k < n_preds - 1 - i; ++k) wgt *= static_cast<double>(gamma);
        rp.weight[i] = static_cast<float>(wgt);
        sum += wgt;
    }
    wsum = static_cast<float>(sum);
    return DROSFM_OK;
}

}  // namespace drosfm

using namespace drosfm;

extern "C" {

int drosfm_smoothness_fwd(const float* image, const float* const* inv_depths, int n_preds, float weight, float* stats,
                          float* loss, void* ws, float* edge_w, int B, int H, int W, drosfm_stream_t stream) {
    DROSFM_REQUIRE(n_preds >= 1 && n_preds <= DROSFM_MAX_PREDS, DROSFM_ERANGE, "smoothness_fwd: n_preds=%d outside [1,%d]",
                   n_preds, DROSFM_MAX_PREDS);
    DROSFM_REQUIRE(B > 0 && H > 0 && W > 0, DROSFM_EINVAL, "smoothness_fwd: empty input");
    DROSFM_REQUIRE(B * n_preds <= 65535 && static_cast<long long>(H) * W < (1ll << 30), DROSFM_ERANGE, "smoothness_fwd: too large");
    DROSFM_REQUIRE(image && inv_depths && stats && loss && ws, DROSFM_EINVAL, "smoothness_fwd: NULL argument");
    DepthList dl{};
    for (int i = 0; i < n_preds; ++i) {
        DROSFM_REQUIRE(inv_depths[i] != nullptr, DROSFM_EINVAL, "smoothness_fwd: inv_depths[%d] is NULL", i);
        dl.d[i] = inv_depths[i];
    }
    const int P = H * W;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int fb = (P + kLossThreads * 4 - 1) / (kLossThreads * 4);     // four pixels per thread: the 2 n shuffle reductions + atomics of the epilogue are per thread
    if (fb < 1) fb = 1;
#define SMOOTH_FWD(NP_) smooth_fwd_kernel<NP_><<<dim3(fb, B), kLossThreads, 0, s>>>(image, dl, n_preds, weight, stats, loss, \
                                                                                  static_cast<Slot*>(ws), edge_w, B, H, W)
    if (n_preds <= 4) SMOOTH_FWD(4);
    else if (n_preds <= 8) SMOOTH_FWD(8);
    else if (n_preds <= 12) SMOOTH_FWD(12);
    else SMOOTH_FWD(16);
#undef SMOOTH_FWD
    return launch_status("smoothness_fwd");
}

int drosfm_smoothness_bwd(const float* g_loss, const float* image, const float* const* inv_depths, int n_preds, float weight,
                          const float* stats, float* const* g_inv_depths, int accumulate, const float* edge_w, int B, int H,
                          int W, drosfm_stream_t stream) {
    DROSFM_REQUIRE(n_preds >= 1 && n_preds <= DROSFM_MAX_PREDS, DROSFM_ERANGE, "smoothness_bwd: n_preds=%d outside [1,%d]",
                   n_preds, DROSFM_MAX_PREDS);
    DROSFM_REQUIRE(B > 0 && H > 0 && W > 0, DROSFM_EINVAL, "smoothness_bwd: empty input");
    DROSFM_REQUIRE(g_loss && image && inv_depths && stats && g_inv_depths, DROSFM_EINVAL, "smoothness_bwd: NULL argument");
    DepthList dl{};
    DepthGrads dg{};
    for (int i = 0; i < n_preds; ++i) {
        DROSFM_REQUIRE(inv_depths[i] != nullptr, DROSFM_EINVAL, "smoothness_bwd: inv_depths[%d] is NULL", i);
        dl.d[i] = inv_depths[i];
        dg.g[i] = g_inv_depths[i];
    }
    const int P = H * W;
    smooth_bwd_kernel<<<dim3((P + kLossThreads - 1) / kLossThreads, B), kLossThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        g_loss, image, dl, n_preds, weight, stats, dg, accumulate, edge_w, B, H, W);
    return launch_status("smoothness_bwd");
}

int drosfm_sup_depth_loss_fwd(const float* gt_inv_depth, const float* const* inv_depths, int n_preds, float min_depth,
                              float max_depth, float gamma, float* loss, void* ws, int B, int H, int W,
                              drosfm_stream_t stream) {
    DROSFM_REQUIRE(B > 0 && H > 0 && W > 0, DROSFM_EINVAL, "sup_depth_loss_fwd: empty input");
    DROSFM_REQUIRE(gt_inv_depth && loss && ws, DROSFM_EINVAL, "sup_depth_loss_fwd: NULL argument");
    SupPtrs sp{};
    float wsum = 1.0f;
    if (int e = fill_sup(sp, wsum, inv_depths, nullptr, n_preds, gamma)) return e;
    const long long N = static_cast<long long>(B) * H * W;
    const bool vec4 = sup_vec4(gt_inv_depth, sp, n_preds, N);
    const long long steps = vec4 ? N / 4 : N;
    long long blocks = (steps + kLossThreads - 1) / kLossThreads;       // one step per thread up to 16 blocks per SM
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    launch_sup_depth<0>(static_cast<unsigned>(blocks), static_cast<cudaStream_t>(stream), vec4, nullptr, gt_inv_depth, sp, n_preds,
                        1.0f / max_depth, 1.0f / min_depth, wsum, loss, static_cast<Slot*>(ws), N);
    return launch_status("sup_depth_loss_fwd");
}

int drosfm_sup_depth_loss_bwd(const float* g_loss, const float* gt_inv_depth, const float* const* inv_depths, int n_preds,
                              float min_depth, float max_depth, float gamma, float* const* g_inv_depths, int B, int H, int W,
                              drosfm_stream_t stream) {
    DROSFM_REQUIRE(B > 0 && H > 0 && W > 0, DROSFM_EINVAL, "sup_depth_loss_bwd: empty input");
    DROSFM_REQUIRE(g_loss && gt_inv_depth && g_inv_depths, DROSFM_EINVAL, "sup_depth_loss_bwd: NULL argument");
    SupPtrs sp{};
    float wsum = 1.0f;
    if (int e = fill_sup(sp, wsum, inv_depths, g_inv_depths, n_preds, gamma)) return e;
    const long long N = static_cast<long long>(B) * H * W;
    const bool vec4 = sup_vec4(gt_inv_depth, sp, n_preds, N);
    const long long steps = vec4 ? N / 4 : N;
    const long long blocks = (steps + kLossThreads - 1) / kLossThreads;
    DROSFM_REQUIRE(blocks < (1ll << 31), DROSFM_ERANGE, "sup_depth_loss_bwd: too large");
    launch_sup_depth<1>(static_cast<unsigned>(blocks), static_cast<cudaStream_t>(stream), vec4, g_loss, gt_inv_depth, sp, n_preds,
                        1.0f / max_depth, 1.0f / min_depth, wsum, nullptr, nullptr, N);
    return launch_status("sup_depth_loss_bwd");
}

int drosfm_reproj_loss_fwd(const float* depth, int depth_kind, const drosfm_cams_t* cams, const float* const* gt_poses,
                           const float* const* pred_poses, int n_views, int n_preds, float min_depth, float max_depth,
                           float gamma, float* loss, void* ws, int B, int H, int W, drosfm_stream_t stream) {
    DROSFM_REQUIRE(B > 0 && H > 0 && W > 0, DROSFM_EINVAL, "reproj_loss_fwd: empty input");
    DROSFM_REQUIRE(B <= 65535 && static_cast<long long>(H) * W < (1ll << 30), DROSFM_ERANGE, "reproj_loss_fwd: too large");
    DROSFM_REQUIRE(cams && cams->K && cams->Kref && depth && loss && ws, DROSFM_EINVAL, "reproj_loss_fwd: NULL argument");
    DROSFM_REQUIRE(cams->pose_kind == DROSFM_POSE_MAT4 || cams->pose_kind == DROSFM_POSE_EULER6, DROSFM_EINVAL,
                   "reproj_loss_fwd: pose_kind must be MAT4 or EULER6");
    ReprojPtrs rp{};
    float wsum = 1.0f;
    if (int e = fill_reproj(rp, wsum, gt_poses, pred_poses, n_views, n_preds, gamma)) return e;
    reproj_kernel<0><<<dim3(strip_blocks(H * W, B), B), kLossThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        nullptr, depth, depth_kind, *cams, rp, n_views, n_preds, min_depth, max_depth, wsum, loss, static_cast<Slot*>(ws), B, H, W);
    return launch_status("reproj_loss_fwd");
}

int drosfm_reproj_loss_bwd(const float* g_loss, const float* depth, int depth_kind, const drosfm_cams_t* cams,
                           const float* const* gt_poses, const float* const* pred_poses, int n_views, int n_preds,
                           float min_depth, float max_depth, float gamma, float* const* g_pred_poses, void* ws,
                           int B, int H, int W, drosfm_stream_t stream) {
    DROSFM_REQUIRE(B > 0 && H > 0 && W > 0, DROSFM_EINVAL, "reproj_loss_bwd: empty input");
    DROSFM_REQUIRE(B <= 65535 && static_cast<long long>(H) * W < (1ll << 30), DROSFM_ERANGE, "reproj_loss_bwd: too large");
    DROSFM_REQUIRE(cams && cams->K && cams->Kref && depth && g_loss && ws && g_pred_poses, DROSFM_EINVAL,
                   "reproj_loss_bwd: NULL argument");
    DROSFM_REQUIRE(cams->pose_kind == DROSFM_POSE_MAT4 || cams->pose_kind == DROSFM_POSE_EULER6, DROSFM_EINVAL,
                   "reproj_loss_bwd: pose_kind must be MAT4 or EULER6");
    ReprojPtrs rp{};
    float wsum = 1.0f;
    if (int e = fill_reproj(rp, wsum, gt_poses, pred_poses, n_views, n_preds, gamma)) return e;
    for (int k = 0; k < n_views * n_preds; ++k) rp.g_pred[k] = g_pred_poses[k];
    reproj_kernel<1><<<dim3(strip_blocks(H * W, B), B), kLossThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        g_loss, depth, depth_kind, *cams, rp, n_views, n_preds, min_depth, max_depth, wsum, nullptr, static_cast<Slot*>(ws), B, H, W);
    return launch_status("reproj_loss_bwd");
}

}  // extern "C"
