// K1 -- Gaussian pyramid (cv::pyrDown chain of cv::buildOpticalFlowPyramid) for all streams.
//
// Reference: the pyramid OpenCV builds inside cv::calcOpticalFlowPyrLK, called at
// src/processing/lkorb_tracking.cpp:64-73 and src/processing/camera_frame.cpp:124-128.
// Arithmetic: separable [1 4 6 4 1], out = (sum + 128) >> 8, BORDER_REFLECT_101,
// output size ((w+1)/2, (h+1)/2).  All integer => bit-exact.
//
// HBM-bound byte work: each CTA stages a (2*TW+3) x (2*TH+3) u8 tile in shared memory with
// coalesced 4-byte loads, does the horizontal 5-tap pass into a u16 tile, then the vertical pass.
// Algorithmic bytes per image: read w*h + write (P - w*h) = P.
#include "ctx.h"

namespace {

constexpr int TW = 64;   // output tile width
constexpr int TH = 16;   // output tile height
constexpr int IN_W = 2 * TW + 3;
constexpr int IN_H = 2 * TH + 3;
constexpr int IN_PITCH = IN_W + 1;   // 132

__device__ __forceinline__ int reflect101(int i, int n) {
  i = i < 0 ? -i : i;
  return i >= n ? 2 * n - 2 - i : i;
}

__global__ void __launch_bounds__(256) pyr_down_kernel(const uint8_t* __restrict__ src_base,
                                                       uint8_t* __restrict__ dst_base,
                                                       int sw, int sh, int spitch, int dw, int dh,
                                                       int dpitch, size_t stream_stride) {
  __shared__ uint8_t tile[IN_H][IN_PITCH];
  __shared__ uint16_t rows[IN_H][TW];
  const uint8_t* src = src_base + (size_t)blockIdx.z * stream_stride;
  uint8_t* dst = dst_base + (size_t)blockIdx.z * stream_stride;
  const int ox0 = blockIdx.x * TW, oy0 = blockIdx.y * TH;
  const int ix0 = 2 * ox0 - 2, iy0 = 2 * oy0 - 2;
  const int tid = threadIdx.x;
  // stage input tile (reflect at the image border)
  for (int i = tid; i < IN_H * IN_W; i += 256) {
    int r = i / IN_W, c = i - r * IN_W;
    int y = reflect101(iy0 + r, sh), x = reflect101(ix0 + c, sw);
    tile[r][c] = src[(size_t)y * spitch + x];
  }
  __syncthreads();
  // horizontal pass: rows[r][ox] = sum_k k[k] * tile[r][2*ox + k]
  for (int i = tid; i < IN_H * TW; i += 256) {
    int r = i / TW, ox = i - r * TW;
    const uint8_t* t = &tile[r][2 * ox];
    rows[r][ox] = (uint16_t)(t[0] + 4 * t[1] + 6 * t[2] + 4 * t[3] + t[4]);
  }
  __syncthreads();
  // vertical pass
  for (int i = tid; i < TH * TW; i += 256) {
    int oy = i / TW, ox = i - oy * TW;
    int gx = ox0 + ox, gy = oy0 + oy;
    if (gx < dw && gy < dh) {
      int s = rows[2 * oy][ox] + 4 * rows[2 * oy + 1][ox] + 6 * rows[2 * oy + 2][ox] +
              4 * rows[2 * oy + 3][ox] + rows[2 * oy + 4][ox];
      dst[(size_t)gy * dpitch + gx] = (uint8_t)((s + 128) >> 8);
    }
  }
}

// level-0 ingest: tight or strided source images (device memory) -> the slot's pitched level-0 rows, all streams
// in one launch.  VEC = 16-byte path when every row start is 16-byte aligned and w % 16 == 0.
template <bool VEC>
__global__ void __launch_bounds__(256) unpack_images_kernel(const uint8_t* __restrict__ src, size_t row_stride,
                                                            size_t img_stride, uint8_t* __restrict__ dst_base,
                                                            size_t stream_stride, int pitch, int w, int h) {
  const int s = blockIdx.z;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (y >= h) return;
  const uint8_t* srow = src + (size_t)s * img_stride + (size_t)y * row_stride;
  uint8_t* drow = dst_base + (size_t)s * stream_stride + (size_t)y * pitch;
  if (VEC) {
    const int n16 = w >> 4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x)
      reinterpret_cast<uint4*>(drow)[i] = reinterpret_cast<const uint4*>(srow)[i];
  } else {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w; i += gridDim.x * blockDim.x) drow[i] = srow[i];
  }
}

}  // namespace

int flv_launch_unpack(flv_ctx* ctx, int slot, int n_streams, const uint8_t* d_src, size_t row_stride,
                      size_t img_stride) {
  const LevelGeom& L0 = ctx->geom.lv[0];
  const bool vec = (ctx->w % 16 == 0) && (row_stride % 16 == 0) && (img_stride % 16 == 0) &&
                   (reinterpret_cast<size_t>(d_src) % 16 == 0);
  dim3 block(64, 4);
  dim3 grid(1, (ctx->h + 3) / 4, n_streams);
  uint8_t* dst = ctx->pyr[slot] + L0.off;
  if (vec)
    unpack_images_kernel<true><<<grid, block, 0, ctx->stream>>>(d_src, row_stride, img_stride, dst,
                                                                 ctx->geom.stream_stride, L0.pitch, ctx->w, ctx->h);
  else
    unpack_images_kernel<false><<<grid, block, 0, ctx->stream>>>(d_src, row_stride, img_stride, dst,
                                                                  ctx->geom.stream_stride, L0.pitch, ctx->w, ctx->h);
  ctx->launches++;
  FLV_CUDA(ctx, cudaGetLastError());
  return FLV_OK;
}

int flv_launch_pyramid(flv_ctx* ctx, int slot, int n_streams) {
  const PyrGeom& g = ctx->geom;
  for (int l = 1; l < g.nlev; ++l) {
    const LevelGeom& a = g.lv[l - 1];
    const LevelGeom& b = g.lv[l];
    dim3 grid((b.w + TW - 1) / TW, (b.h + TH - 1) / TH, n_streams);
    pyr_down_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->pyr[slot] + a.off, ctx->pyr[slot] + b.off,
                                                   a.w, a.h, a.pitch, b.w, b.h, b.pitch,
                                                   g.stream_stride);
    ctx->launches++;
  }
  FLV_CUDA(ctx, cudaGetLastError());
  return FLV_OK;
}
