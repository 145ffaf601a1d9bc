// Evaluation path of the depth predictions.
//
//   post_process_inv_depth   dro_sfm/utils/depth.py:230-258   (fuse a prediction with the prediction of the flipped image)
//   compute_depth_metrics    dro_sfm/utils/depth.py:261-340   (abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3, SILog, iabs_diff)
//
// The reference evaluates the metrics sample by sample with boolean indexing, a sort-based torch.median and ~60 ATen
// launches per sample and mode.  Here: one launch for the post-processing; for the metrics one pass per radix digit of
// an EXACT selection of the median (the lower of the two middle values, as torch.median) over the valid pixels'
// gt / pred ratios -- three passes of 11 + 11 + 10 bits, histograms in the caller's workspace -- and one pass that
// accumulates every metric of every sample in fp64.  The prediction is interpolated to the ground-truth resolution
// (bilinear, align_corners=True) on the fly; nothing is materialised.
#include "common.cuh"

namespace drosfm {

#ifndef DROSFM_LINSPACE_FMA
#define DROSFM_LINSPACE_FMA 1
#endif

constexpr int kEvalThreads = 256;
constexpr int kBins = 2048;

// torch.linspace(0, 1, W, device='cuda')[x] in float32: the two-sided evaluation of ATen's kernel (RangeFactories.cu:
// ``ind < steps / 2 ? start + step * ind : end - step * (steps - ind - 1)``), whose multiply-subtract nvcc contracts
// into one fused operation -- reproduced explicitly so that the blend weights carry torch's bits
__device__ __forceinline__ float linspace01(int x, int W) {
    if (W == 1) return 0.0f;
    const float step = __fdiv_rn(1.0f, static_cast<float>(W - 1));
#if DROSFM_LINSPACE_FMA
    return x < W / 2 ? __fmul_rn(step, static_cast<float>(x)) : __fmaf_rn(-step, static_cast<float>(W - 1 - x), 1.0f);
#else
    return x < W / 2 ? __fmul_rn(step, static_cast<float>(x)) : __fsub_rn(1.0f, __fmul_rn(step, static_cast<float>(W - 1 - x)));
#endif
}

__global__ void __launch_bounds__(kEvalThreads)
post_process_kernel(const float* __restrict__ inv, const float* __restrict__ inv_flipped, float* __restrict__ out, int H, int W,
                    int method, size_t total) {
    const size_t i = static_cast<size_t>(blockIdx.x) * kEvalThreads + threadIdx.x;
    if (i >= total) return;
    const int x = static_cast<int>(i % W);
    const size_t row = i - x;
    const float a = __ldg(inv + i);
    const float hat = __ldg(inv_flipped + row + (W - 1 - x));            // flip_lr(inv_depth_flipped)
    float fused;
    if (method == 0) fused = __fmul_rn(0.5f, __fadd_rn(a, hat));
    else if (method == 1) fused = fmaxf(a, hat);
    else fused = fminf(a, hat);
    // mask = 1 - clamp(20 * (xs - 0.05), 0, 1), mask_hat = flip_lr(mask)
    auto mask_at = [&](int xx) {
        const float t = __fmul_rn(20.0f, __fsub_rn(linspace01(xx, W), 0.05f));
        return __fsub_rn(1.0f, fminf(fmaxf(t, 0.0f), 1.0f));
    };
    const float m = mask_at(x), mh = mask_at(W - 1 - x);
    // mask_hat * inv_depth + mask * inv_depth_hat + (1 - mask - mask_hat) * fused, in the reference's operation order
    out[i] = __fadd_rn(__fadd_rn(__fmul_rn(mh, a), __fmul_rn(m, hat)), __fmul_rn(__fsub_rn(__fsub_rn(1.0f, m), mh), fused));
}

// ---- metrics ---------------------------------------------------------------------------------
struct EvalCfg {
    int H, W, Hp, Wp;            // ground-truth and prediction resolution
    float min_depth, max_depth;
    int y1, y2, x1, x2;          // crop window (all pixels when crop is off)
    int use_gt_scale;
};

// workspace of one sample (all zero on entry; the last kernel leaves it zero again)
struct alignas(16) EvalSample {
    unsigned hist[kBins];
    unsigned long long count;    // valid pixels
    unsigned prefix;             // bits of the median ratio selected so far
    unsigned rank;               // rank still to be found inside the selected bucket
    double acc[12];              // metric sums
    unsigned long long ticket;
    unsigned long long pad_;
};

// F.interpolate(pred, size=(H, W), mode='bilinear', align_corners=True) at one ground-truth pixel, then clamp(min=1e-6)
__device__ __forceinline__ float pred_at(const float* __restrict__ pred, const EvalCfg& c, int y, int x) {
    float v;
    if (c.Hp == c.H && c.Wp == c.W) {
        v = __ldg(pred + y * c.W + x);
    } else {
        const float sy = c.H > 1 ? static_cast<float>(c.Hp - 1) / static_cast<float>(c.H - 1) : 0.0f;
        const float sx = c.W > 1 ? static_cast<float>(c.Wp - 1) / static_cast<float>(c.W - 1) : 0.0f;
        const float ry = sy * static_cast<float>(y), rx = sx * static_cast<float>(x);
        const int y0 = static_cast<int>(ry), x0 = static_cast<int>(rx);
        const int yp = y0 < c.Hp - 1 ? 1 : 0, xp = x0 < c.Wp - 1 ? 1 : 0;
        const float ly = ry - static_cast<float>(y0), lx = rx - static_cast<float>(x0);
        const float* p = pred + y0 * c.Wp + x0;
        const float top = (1.0f - lx) * __ldg(p) + lx * __ldg(p + xp);
        const float bot = (1.0f - lx) * __ldg(p + yp * c.Wp) + lx * __ldg(p + yp * c.Wp + xp);
        v = (1.0f - ly) * top + ly * bot;
    }
    return v < 1e-6f ? 1e-6f : v;
}

__device__ __forceinline__ bool valid_at(float gt, const EvalCfg& c, int y, int x) {
    return gt > c.min_depth && gt < c.max_depth && y >= c.y1 && y < c.y2 && x >= c.x1 && x < c.x2;
}

// PASS 0..2: histogram of one radix digit of the ratio gt / pred over the valid pixels whose higher digits equal the
// prefix selected so far (pass 0 also counts the valid pixels).
template <int PASS>
__global__ void __launch_bounds__(kEvalThreads)
ratio_histogram_kernel(const float* __restrict__ gt, const float* __restrict__ pred, EvalCfg c, EvalSample* ws) {
    __shared__ unsigned h[kBins];
    const int b = blockIdx.y;
    EvalSample& s = ws[b];
    for (int i = threadIdx.x; i < kBins; i += kEvalThreads) h[i] = 0u;
    __syncthreads();
    const unsigned prefix = PASS == 0 ? 0u : s.prefix;
    const int P = c.H * c.W;
    const float* g = gt + static_cast<size_t>(b) * P;
    const float* p = pred + static_cast<size_t>(b) * c.Hp * c.Wp;
    unsigned n = 0;
    for (int i = blockIdx.x * kEvalThreads + threadIdx.x; i < P; i += gridDim.x * kEvalThreads) {
        const int y = i / c.W, x = i - y * c.W;
        const float gv = __ldg(g + i);
        if (!valid_at(gv, c, y, x)) continue;
        const unsigned bits = __float_as_uint(__fdiv_rn(gv, pred_at(p, c, y, x)));       // positive: integer order == float order
        if (PASS == 0) { atomicAdd(&h[bits >> 21], 1u); ++n; }
        else if (PASS == 1) { if ((bits >> 21) == (prefix >> 21)) atomicAdd(&h[(bits >> 10) & 0x7ffu], 1u); }
        else { if ((bits >> 10) == (prefix >> 10)) atomicAdd(&h[bits & 0x3ffu], 1u); }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kBins; i += kEvalThreads)
        if (h[i] != 0u) atomicAdd(&s.hist[i], h[i]);
    if (PASS == 0) {
        n = __reduce_add_sync(0xffffffffu, n);
        if ((threadIdx.x & 31) == 0 && n != 0u) atomicAdd(&s.count, static_cast<unsigned long long>(n));
    }
}

// Between the passes: the bucket that contains the wanted rank (one block per sample); re-zeroes the histogram.
template <int PASS>
__global__ void __launch_bounds__(kEvalThreads)
select_bucket_kernel(EvalSample* ws) {
    __shared__ unsigned part[kEvalThreads];
    EvalSample& s = ws[blockIdx.x];
    // torch.median: the lower of the two middle values = sorted[(n - 1) / 2]
    const unsigned long long want64 = PASS == 0 ? (s.count == 0 ? 0 : (s.count - 1) / 2) : s.rank;
    const unsigned want = static_cast<unsigned>(want64);
    constexpr int PER = kBins / kEvalThreads;
    unsigned mine[PER], sum = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) { mine[k] = s.hist[threadIdx.x * PER + k]; sum += mine[k]; }
    part[threadIdx.x] = sum;
    __syncthreads();
    // exclusive prefix over the threads (256 values: a serial scan by every thread of its own prefix is cheap enough)
    unsigned before = 0;
    for (int t = 0; t < static_cast<int>(threadIdx.x); ++t) before += part[t];
    __syncthreads();
    if (want >= before && want < before + sum) {
        unsigned run = before;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            if (want >= run && want < run + mine[k]) {
                const unsigned bin = threadIdx.x * PER + k;
                s.prefix = PASS == 0 ? bin << 21 : (PASS == 1 ? s.prefix | (bin << 10) : s.prefix | bin);
                s.rank = want - run;
            }
            run += mine[k];
        }
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) s.hist[threadIdx.x * PER + k] = 0u;
}

// Final pass: every metric of compute_depth_metrics as fp64 sums over the valid pixels; the last block of the grid
// turns them into the nine averaged values.
__global__ void __launch_bounds__(kEvalThreads)
depth_metrics_kernel(const float* __restrict__ gt, const float* __restrict__ pred, EvalCfg c, EvalSample* ws, int B,
                     float* __restrict__ metrics) {
    __shared__ double red[12 * (kEvalThreads / 32)];
    __shared__ int flag;
    const int b = blockIdx.y;
    EvalSample& s = ws[b];
    const int P = c.H * c.W;
    const float* g = gt + static_cast<size_t>(b) * P;
    const float* p = pred + static_cast<size_t>(b) * c.Hp * c.Wp;
    const float scale = c.use_gt_scale ? __uint_as_float(s.prefix) : 1.0f;        // torch.median(gt_i / pred_i)
    float v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = 0.0f;
    double acc[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[k] = 0.0;
    int since = 0;
    for (int i = blockIdx.x * kEvalThreads + threadIdx.x; i < P; i += gridDim.x * kEvalThreads) {
        const int y = i / c.W, x = i - y * c.W;
        const float gv = __ldg(g + i);
        if (!valid_at(gv, c, y, x)) continue;
        float pv = pred_at(p, c, y, x);
        if (c.use_gt_scale) pv = fminf(fmaxf(__fmul_rn(pv, scale), c.min_depth), c.max_depth);
        pv = fminf(fmaxf(pv, c.min_depth), c.max_depth);
        const float thresh = fmaxf(__fdiv_rn(gv, pv), __fdiv_rn(pv, gv));
        const float diff = gv - pv, lg = logf(gv) - logf(pv);
        v[0] += fabsf(diff) / gv;                                  // abs_rel
        v[1] += diff * diff / gv;                                  // sq_rel
        v[2] += diff * diff;                                       // rmse (before the root)
        v[3] += lg * lg;                                           // rmse_log / SILog second moment
        v[4] += thresh < 1.25f ? 1.0f : 0.0f;                      // a1
        v[5] += thresh < 1.5625f ? 1.0f : 0.0f;                    // a2 (1.25 ** 2)
        v[6] += thresh < 1.953125f ? 1.0f : 0.0f;                  // a3 (1.25 ** 3)
        v[7] += lg;                                                // SILog first moment
        v[8] += fabsf(1.0f / pv - 1.0f / gv);                      // iabs_diff
        if (++since == 64) {                                       // fp32 partial sums are flushed into fp64 regularly
#pragma unroll
            for (int k = 0; k < 9; ++k) { acc[k] += v[k]; v[k] = 0.0f; }
            since = 0;
        }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] += v[k];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const double t = warp_sum(acc[k]);
        if (lane == 0) red[k * (kEvalThreads / 32) + wid] = t;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        double t = 0.0;
        for (int w = 0; w < kEvalThreads / 32; ++w) t += red[threadIdx.x * (kEvalThreads / 32) + w];
        if (t != 0.0) atomicAdd(&s.acc[threadIdx.x], t);
    }
    // last block of the whole grid: per-sample metrics, averaged over the batch (samples without a valid pixel add zero)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long t = atomicAdd(&ws[0].ticket, 1ull);
        flag = t == static_cast<unsigned long long>(gridDim.x) * gridDim.y - 1ull ? 1 : 0;
    }
    __syncthreads();
    if (!flag) return;
    __threadfence();
    if (threadIdx.x < 9) {
        double total = 0.0;
        for (int i = 0; i < B; ++i) {
            const double n = static_cast<double>(__ldcg(&ws[i].count));
            if (n == 0.0) continue;
            const double a = __ldcg(&ws[i].acc[threadIdx.x]);
            double m;
            // the reference takes float32 means, square roots and sums: rounded to float32 at the same points
            switch (threadIdx.x) {
                case 2: m = sqrtf(static_cast<float>(a / n)); break;                                   // rmse
                case 3: m = sqrtf(static_cast<float>(a / n)); break;                                   // rmse_log
                case 7: {                                                                              // SILog
                    const double second = __ldcg(&ws[i].acc[3]) / n, first = a;
                    const float inside = static_cast<float>(second) - static_cast<float>(first * first / (n * n));
                    m = sqrtf(inside);
                    break;
                }
                default: m = static_cast<float>(a / n);
            }
            total += m;
        }
        metrics[threadIdx.x] = static_cast<float>(total / static_cast<double>(B));
    }
    __syncthreads();
    // leave the workspace zeroed for the next call
    for (int i = threadIdx.x; i < B * static_cast<int>(sizeof(EvalSample) / 4); i += kEvalThreads) reinterpret_cast<unsigned*>(ws)[i] = 0u;
}

}  // namespace drosfm

using namespace drosfm;

extern "C" {

size_t drosfm_eval_ws_bytes(int B) { return static_cast<size_t>(B < 1 ? 1 : B) * sizeof(EvalSample); }

int drosfm_post_process_inv_depth(const float* inv_depth, const float* inv_depth_flipped, float* out, int B, int H, int W,
                                  int method, drosfm_stream_t stream) {
    DROSFM_REQUIRE(B >= 0 && H >= 0 && W >= 0, DROSFM_EINVAL, "post_process_inv_depth: negative dimension");
    DROSFM_REQUIRE(method >= 0 && method <= 2, DROSFM_EINVAL, "post_process_inv_depth: method must be 0 (mean), 1 (max) or 2 (min)");
    const size_t total = static_cast<size_t>(B) * H * W;
    if (total == 0) return DROSFM_OK;
    DROSFM_REQUIRE(inv_depth && inv_depth_flipped && out, DROSFM_EINVAL, "post_process_inv_depth: NULL argument");
    DROSFM_REQUIRE((total + kEvalThreads - 1) / kEvalThreads < (1ull << 31), DROSFM_ERANGE, "post_process_inv_depth: too many elements");
    post_process_kernel<<<static_cast<unsigned>((total + kEvalThreads - 1) / kEvalThreads), kEvalThreads, 0,
                          static_cast<cudaStream_t>(stream)>>>(inv_depth, inv_depth_flipped, out, H, W, method, total);
    return launch_status("post_process_inv_depth");
}

int drosfm_depth_metrics(const float* gt, const float* pred, int B, int H, int W, int Hp, int Wp, float min_depth, float max_depth,
                         int crop, int use_gt_scale, float* metrics, void* ws, drosfm_stream_t stream) {
    DROSFM_REQUIRE(B >= 1 && H >= 1 && W >= 1 && Hp >= 1 && Wp >= 1, DROSFM_EINVAL, "depth_metrics: empty input");
    DROSFM_REQUIRE(B <= 65535 && static_cast<long long>(H) * W < (1ll << 30), DROSFM_ERANGE, "depth_metrics: dimension out of range");
    DROSFM_REQUIRE(gt && pred && metrics && ws, DROSFM_EINVAL, "depth_metrics: NULL argument");
    DROSFM_REQUIRE(crop >= 0 && crop <= 2, DROSFM_EINVAL, "depth_metrics: crop must be 0 (none), 1 (garg) or 2 (eigen_nyu)");
    EvalCfg c{H, W, Hp, Wp, min_depth, max_depth, 0, H, 0, W, use_gt_scale ? 1 : 0};
    if (crop == 1) {           // depth.py:293-297
        c.y1 = static_cast<int>(0.40810811 * H); c.y2 = static_cast<int>(0.99189189 * H);
        c.x1 = static_cast<int>(0.03594771 * W); c.x2 = static_cast<int>(0.96405229 * W);
    } else if (crop == 2) {    // depth.py:298-303
        c.y1 = 20; c.y2 = 459; c.x1 = 24; c.x2 = 615;
    }
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    EvalSample* s = static_cast<EvalSample*>(ws);
    const int P = H * W;
    int blocks = (P + kEvalThreads * 8 - 1) / (kEvalThreads * 8);
    if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
    dim3 grid(blocks, B);
    if (use_gt_scale) {
        ratio_histogram_kernel<0><<<grid, kEvalThreads, 0, cs>>>(gt, pred, c, s);
        select_bucket_kernel<0><<<B, kEvalThreads, 0, cs>>>(s);
        ratio_histogram_kernel<1><<<grid, kEvalThreads, 0, cs>>>(gt, pred, c, s);
        select_bucket_kernel<1><<<B, kEvalThreads, 0, cs>>>(s);
        ratio_histogram_kernel<2><<<grid, kEvalThreads, 0, cs>>>(gt, pred, c, s);
        select_bucket_kernel<2><<<B, kEvalThreads, 0, cs>>>(s);
    } else {
        ratio_histogram_kernel<0><<<grid, kEvalThreads, 0, cs>>>(gt, pred, c, s);      // the valid-pixel counts
        select_bucket_kernel<0><<<B, kEvalThreads, 0, cs>>>(s);
    }
    if (int e = launch_status("depth_metrics (median)")) return e;
    depth_metrics_kernel<<<grid, kEvalThreads, 0, cs>>>(gt, pred, c, s, B, metrics);
    return launch_status("depth_metrics");
}

}  // extern "C"
