// Storage-layout conversion of a feature map between NCHW and NHWC (torch channels_last).
//
// The reference keeps every feature map NCHW (DepthPoseNet.py:113-115); the channels-last cost kernels want every
// tap of the C-channel gather to be one contiguous segment.  networks/cost.py converts each distinct map once per
// forward and the summed cost gradient once per backward; this is that conversion as a tiled transpose of the
// per-sample [C, H*W] matrix: 32x32 tiles through padded shared memory, both sides coalesced 128-byte rows
// (the generic strided copy it replaces reaches a third of this).
#include "common.cuh"

namespace drosfm {

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
    dim3 grid((cols + kTile - 1) / kTile, (rows + kTile - 1) / kTile, B);
    DROSFM_REQUIRE(grid.y <= 65535, DROSFM_ERANGE, "relayout: too many rows");
    transpose_kernel<<<grid, kTile * kTileRows, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, rows, cols);
    return launch_status("relayout");
}

}  // extern "C"
