// Storage-layout conversion of a feature map between NCHW and NHWC (torch channels_last).
//
// The reference keeps every feature map NCHW (DepthPoseNet.py:113-115); the channels-last cost kernels want every
// tap of the C-channel gather to be one contiguous segment.  networks/cost.py converts each distinct map once per
// forward and the summed cost gradient once per backward; this is that conversion as a tiled transpose of the
// per-sample [C, H*W] matrix: 32x32 tiles through padded shared memory, both sides coalesced 128-byte rows
// (the generic strided copy it replaces reaches a third of this).
#include "common.cuh"

namespace drosfm {

#ifndef DROSFM_TRANSPOSE_VEC4
#define DROSFM_TRANSPOSE_VEC4 1
#endif
constexpr int kTile = 32, kTileRows = 8;

// src: [rows][cols] row-major, dst: [cols][rows] row-major, one matrix per blockIdx.z
__global__ void __launch_bounds__(kTile * kTileRows)
transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
    __shared__ float tile[kTile][kTile + 1];
    const size_t base = static_cast<size_t>(blockIdx.z) * rows * cols;
    const int c0 = blockIdx.x * kTile, r0 = blockIdx.y * kTile;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kTile; k += kTileRows) {
        const int r = r0 + ty + k, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + k][tx] = __ldg(src + base + static_cast<size_t>(r) * cols + c);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTile; k += kTileRows) {
        const int c = c0 + ty + k, r = r0 + tx;
        if (r < rows && c < cols) dst[base + static_cast<size_t>(c) * rows + r] = tile[tx][ty + k];
    }
}

// The same transpose with 128-bit global accesses: 64x64 tiles, every thread moves four float4 in and four out (rows and
// cols multiples of 4 and 16-byte aligned buffers; the tile is stored transposed, its 65-float pitch keeps the scalar
// shared-memory accesses at two-way conflicts).  28 memory instructions per 16 elements instead of 64: the 32x32 kernel
// ran at 71 % issue utilisation and 20 % of DRAM peak on the stacked training maps (profiles/r2_summary.txt).
constexpr int kTile4 = 64;

__global__ void __launch_bounds__(256)
transpose4_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
    __shared__ float tile[kTile4][kTile4 + 1];               // tile[c][r]
    const size_t base = static_cast<size_t>(blockIdx.z) * rows * cols;
    const int c0 = blockIdx.x * kTile4, r0 = blockIdx.y * kTile4;
    const int q = threadIdx.x & 15, t = threadIdx.x >> 4;    // q: float4 index inside a 64-wide line, t: line 0..15 (+16k)
#pragma unroll
    for (int k = 0; k < kTile4; k += 16) {
        const int r = r0 + t + k, c = c0 + 4 * q;
        if (r < rows && c < cols) {                          // cols % 4 == 0: a float4 is inside or outside as a whole
            const float4 v = __ldg(reinterpret_cast<const float4*>(src + base + static_cast<size_t>(r) * cols + c));
            tile[4 * q + 0][t + k] = v.x;
            tile[4 * q + 1][t + k] = v.y;
            tile[4 * q + 2][t + k] = v.z;
            tile[4 * q + 3][t + k] = v.w;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTile4; k += 16) {
        const int c = c0 + t + k, r = r0 + 4 * q;
        if (c < cols && r < rows) {
            const float* line = tile[t + k] + 4 * q;
            *reinterpret_cast<float4*>(dst + base + static_cast<size_t>(c) * rows + r) = make_float4(line[0], line[1], line[2], line[3]);
        }
    }
}

// uint8 pictures -> float32 in [0,1]: x / 255 with the IEEE division ToTensor performs (datasets/augmentations.py:149-152).
// 16 pixels per thread: one 128-bit load, four 128-bit stores.
__global__ void __launch_bounds__(256) u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, size_t n) {
    const size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 16;
    if (i + 16 <= n) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(src + i));
        const unsigned w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float4 o;
            o.x = __fdiv_rn(static_cast<float>(w[k] & 0xffu), 255.0f);
            o.y = __fdiv_rn(static_cast<float>((w[k] >> 8) & 0xffu), 255.0f);
            o.z = __fdiv_rn(static_cast<float>((w[k] >> 16) & 0xffu), 255.0f);
            o.w = __fdiv_rn(static_cast<float>(w[k] >> 24), 255.0f);
            __stcs(reinterpret_cast<float4*>(dst + i + 4 * k), o);
        }
    } else {
        for (size_t j = i; j < n; ++j) dst[j] = __fdiv_rn(static_cast<float>(src[j]), 255.0f);
    }
}

}  // namespace drosfm

using namespace drosfm;

extern "C" {

int drosfm_relayout(const float* src, float* dst, int B, int C, int H, int W, int to_layout, drosfm_stream_t stream) {
    DROSFM_REQUIRE(B >= 0 && C >= 0 && H >= 0 && W >= 0, DROSFM_EINVAL, "relayout: negative dimension");
    DROSFM_REQUIRE(to_layout == DROSFM_NCHW || to_layout == DROSFM_NHWC, DROSFM_EINVAL, "relayout: bad layout %d", to_layout);
    if (B == 0 || C == 0 || H * W == 0) return DROSFM_OK;
    DROSFM_REQUIRE(src != nullptr && dst != nullptr && src != dst, DROSFM_EINVAL, "relayout: NULL or aliased buffers");
    DROSFM_REQUIRE(B <= 65535 && static_cast<long long>(H) * W < (1ll << 30) && C <= (1 << 20), DROSFM_ERANGE,
                   "relayout: dimension out of range");
    const int P = H * W;
    // to NHWC: the per-sample source is [C][P]; to NCHW: it is [P][C]
    const int rows = to_layout == DROSFM_NHWC ? C : P, cols = to_layout == DROSFM_NHWC ? P : C;
#if DROSFM_TRANSPOSE_VEC4
    if (rows % 4 == 0 && cols % 4 == 0 && aligned16(src) && aligned16(dst)) {
        dim3 grid4((cols + kTile4 - 1) / kTile4, (rows + kTile4 - 1) / kTile4, B);
        DROSFM_REQUIRE(grid4.y <= 65535, DROSFM_ERANGE, "relayout: too many rows");
        transpose4_kernel<<<grid4, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, rows, cols);
        return launch_status("relayout");
    }
#endif
    dim3 grid((cols + kTile - 1) / kTile, (rows + kTile - 1) / kTile, B);
    DROSFM_REQUIRE(grid.y <= 65535, DROSFM_ERANGE, "relayout: too many rows");
    transpose_kernel<<<grid, kTile * kTileRows, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, rows, cols);
    return launch_status("relayout");
}

int drosfm_images_u8_to_f32(const uint8_t* src, float* dst, size_t n, drosfm_stream_t stream) {
    if (n == 0) return DROSFM_OK;
    DROSFM_REQUIRE(src != nullptr && dst != nullptr, DROSFM_EINVAL, "images_u8_to_f32: NULL buffer");
    DROSFM_REQUIRE(aligned16(src) && aligned16(dst), DROSFM_EALIGN, "images_u8_to_f32: buffers must be 16-byte aligned");
    const size_t threads = (n + 15) / 16;
    DROSFM_REQUIRE(threads / 256 < (1ull << 31), DROSFM_ERANGE, "images_u8_to_f32: too many elements");
    u8_to_f32_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, n);
    return launch_status("images_u8_to_f32");
}

}  // extern "C"
