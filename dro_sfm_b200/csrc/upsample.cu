// Convex up-sampling of the low-resolution inverse depth (next-row f3 of SURVEY.md section 8).
//
//   DepthPoseNet.upsample_depth   dro_sfm/networks/depth_pose/DepthPoseNet.py:63-74
//
//   out[n,0,8y+i,8x+j] = sum_k softmax_k(mask[n, k*64 + i*8 + j, y, x]) * depth_pad[n, y+ky-1, x+kx-1],   k = 3*ky + kx
//
// The reference materialises the [N,1,9,8,8,H,W] soft-max, the unfolded depth and their product; here every
// mask value is read once and every output written once: 2304 + 256 + 4 bytes per low-resolution pixel forward,
// 4868 backward (mask, upstream gradient -> mask gradient, depth gradient).
// Mapping: block = 32 consecutive low-res pixels of one row x the 8x8 sub-positions; warp w owns sub-row i = w,
// lane = low-res pixel, so every mask load is a coalesced 128-byte row of one channel plane and every warp
// writes one contiguous 256-float output row (staged through shared memory).
#include "common.cuh"

namespace drosfm {

constexpr int kUpThreads = 256;   // 8 warps = 8 sub-rows

__device__ __forceinline__ void load_neighbours(const float* __restrict__ depth, int y, int x, int H, int W, bool on, float* d) {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int yy = y + ky - 1, xx = x + kx - 1;
            d[ky * 3 + kx] = (on && yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(depth + yy * W + xx) : 0.0f;
        }
}

__global__ void __launch_bounds__(kUpThreads)
upsample_fwd_kernel(const float* __restrict__ depth, const float* __restrict__ mask, float* __restrict__ out, int H, int W,
                    float disp_min, float disp_range) {
    __shared__ float rows[8][32 * 8 + 4];
    const int lane = threadIdx.x & 31, i = threadIdx.x >> 5;
    const int x0 = blockIdx.x * 32, y = blockIdx.y, n = blockIdx.z;
    const int x = x0 + lane;
    const bool on = x < W;
    const size_t P = static_cast<size_t>(H) * W;
    float d[9];
    load_neighbours(depth + n * P, y, x, H, W, on, d);
    const float* m = mask + (static_cast<size_t>(n) * 576 + i * 8) * P + static_cast<size_t>(y) * W + (on ? x : 0);
#pragma unroll 2
    for (int j = 0; j < 8; ++j) {
        float v[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) v[k] = __ldg(m + (static_cast<size_t>(k) * 64 + j) * P);
        float mx = v[0];
#pragma unroll
        for (int k = 1; k < 9; ++k) mx = fmaxf(mx, v[k]);
        float den = 0.0f, num = 0.0f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float e = expf(v[k] - mx);
            den += e;
            num += e * d[k];
        }
        // epilogue: disp_to_depth's scaling (layers.py:17; 0 + 1 * x for the plain operator)
        rows[i][lane * 8 + j] = __fadd_rn(disp_min, __fmul_rn(disp_range, num / den));
    }
    __syncwarp();
    const int Wo = W * 8;
    float* o = out + (static_cast<size_t>(n) * H * 8 + y * 8 + i) * Wo + x0 * 8;
    const int valid = min(32, W - x0) * 8;
    for (int c = lane; c < valid; c += 32) o[c] = rows[i][c];
}

__global__ void __launch_bounds__(kUpThreads)
upsample_bwd_kernel(const float* __restrict__ g_out, const float* __restrict__ depth, const float* __restrict__ mask,
                    float* __restrict__ g_depth, float* __restrict__ g_mask, int H, int W, float disp_range) {
    __shared__ float rows[8][32 * 8 + 4];
    __shared__ float part[8][9][32];
    const int lane = threadIdx.x & 31, i = threadIdx.x >> 5;
    const int x0 = blockIdx.x * 32, y = blockIdx.y, n = blockIdx.z;
    const int x = x0 + lane;
    const bool on = x < W;
    const size_t P = static_cast<size_t>(H) * W;
    const int Wo = W * 8;
    const float* g = g_out + (static_cast<size_t>(n) * H * 8 + y * 8 + i) * Wo + x0 * 8;
    const int valid = min(32, W - x0) * 8;
    for (int c = lane; c < 256; c += 32) rows[i][c] = c < valid ? __ldg(g + c) * disp_range : 0.0f;
    float d[9];
    load_neighbours(depth + n * P, y, x, H, W, on, d);
    __syncwarp();
    const size_t moff = (static_cast<size_t>(n) * 576 + i * 8) * P + static_cast<size_t>(y) * W + (on ? x : 0);
    float gd[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) gd[k] = 0.0f;
#pragma unroll 2
    for (int j = 0; j < 8; ++j) {
        float v[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) v[k] = __ldg(mask + moff + (static_cast<size_t>(k) * 64 + j) * P);
        float mx = v[0];
#pragma unroll
        for (int k = 1; k < 9; ++k) mx = fmaxf(mx, v[k]);
        float den = 0.0f, num = 0.0f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            v[k] = expf(v[k] - mx);
            den += v[k];
            num += v[k] * d[k];
        }
        const float inv = 1.0f / den, o = num * inv, go = rows[i][lane * 8 + j];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float s = v[k] * inv;
            gd[k] += s * go;
            if (on && g_mask != nullptr) g_mask[moff + (static_cast<size_t>(k) * 64 + j) * P] = s * (d[k] - o) * go;
        }
    }
    if (g_depth == nullptr) return;
#pragma unroll
    for (int k = 0; k < 9; ++k) part[i][k][lane] = gd[k];
    __syncthreads();
    // 9 x 32 sums over the 8 sub-rows, then one atomic per (pixel, neighbour)
    for (int idx = threadIdx.x; idx < 9 * 32; idx += kUpThreads) {
        const int k = idx >> 5, l = idx & 31;
        float s = 0.0f;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += part[r][k][l];
        const int xx = x0 + l + (k % 3) - 1, yy = y + (k / 3) - 1;
        if (x0 + l < W && xx >= 0 && xx < W && yy >= 0 && yy < H) atomicAdd(g_depth + n * P + yy * W + xx, s);
    }
}

static int check_up(int N, int H, int W, int ratio) {
    DROSFM_REQUIRE(ratio == 8, DROSFM_ENOTSUP, "upsample: only ratio 8 is supported (got %d)", ratio);
    DROSFM_REQUIRE(N >= 0 && H >= 0 && W >= 0, DROSFM_EINVAL, "upsample: negative dimension");
    DROSFM_REQUIRE(N <= 65535 && H <= 65535 && static_cast<long long>(H) * W < (1ll << 24), DROSFM_ERANGE, "upsample: dimension out of range");
    return DROSFM_OK;
}

}  // namespace drosfm

using namespace drosfm;

extern "C" {

int drosfm_upsample_depth_fwd(const float* depth, const float* mask, float* out, int N, int H, int W, int ratio,
                              float disp_min, float disp_range, drosfm_stream_t stream) {
    if (int e = check_up(N, H, W, ratio)) return e;
    if (N == 0 || H * W == 0) return DROSFM_OK;
    DROSFM_REQUIRE(depth && mask && out, DROSFM_EINVAL, "upsample_depth_fwd: NULL argument");
    dim3 grid((W + 31) / 32, H, N);
    upsample_fwd_kernel<<<grid, kUpThreads, 0, static_cast<cudaStream_t>(stream)>>>(depth, mask, out, H, W, disp_min, disp_range);
    return launch_status("upsample_depth_fwd");
}

int drosfm_upsample_depth_bwd(const float* g_out, const float* depth, const float* mask, float* g_depth, float* g_mask,
                              int N, int H, int W, int ratio, float disp_range, drosfm_stream_t stream) {
    if (int e = check_up(N, H, W, ratio)) return e;
    if (N == 0 || H * W == 0) return DROSFM_OK;
    DROSFM_REQUIRE(g_out && depth && mask, DROSFM_EINVAL, "upsample_depth_bwd: NULL argument");
    dim3 grid((W + 31) / 32, H, N);
    upsample_bwd_kernel<<<grid, kUpThreads, 0, static_cast<cudaStream_t>(stream)>>>(g_out, depth, mask, g_depth, g_mask, H, W, disp_range);
    return launch_status("upsample_depth_bwd");
}

}  // extern "C"
