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


// Fused ingest + first pyrDown: one pass over the source image writes the slot's pitched level 0 AND level 1.
// A warp owns 64 output columns (lane = two adjacent outputs = one aligned input word + a byte from each neighbour,
// fetched by shuffle) and marches down ROWS output rows keeping the five horizontal sums of the vertical window in
// registers as packed u16 pairs -- no shared memory; (t0 + 4 t1 + 6 t2 + 4 t3 + t4) <= 4080 and 16 * 4080 + 128 < 65536,
// so both halves of a register are summed and rounded with ordinary 32-bit adds.  Requires w % 4 == 0 and 4-byte
// aligned source rows (otherwise the generic unpack + pyr_down path runs).
constexpr int FUSED_ROWS = 8;     // output rows per warp (16 owned input rows + 3 halo rows)
constexpr int FUSED_WARPS = 4;

__device__ __forceinline__ unsigned hsum_pair(unsigned A, unsigned B, unsigned C) {
  const unsigned lo = __dp4a(__byte_perm(A, B, 0x5432), 0x04060401u, (B >> 16) & 0xffu);   // taps X-2 .. X+2
  const unsigned hi = __dp4a(B, 0x04060401u, C & 0xffu);                                     // taps X   .. X+4
  return lo | (hi << 16);
}

// EQ: cv::equalizeHist is applied on the fly -- every source word goes through the stream's 256-entry LUT (lut_kernel, below)
// held in shared memory, so the equalised image is never written to / re-read from a landing area.
__device__ __forceinline__ unsigned lut4(const unsigned char* l, unsigned v) {
  return (unsigned)l[v & 0xffu] | ((unsigned)l[(v >> 8) & 0xffu] << 8) | ((unsigned)l[(v >> 16) & 0xffu] << 16) | ((unsigned)l[v >> 24] << 24);
}

template <bool EQ>
__global__ void __launch_bounds__(FUSED_WARPS * 32) ingest_l1_kernel(const uint8_t* __restrict__ src, size_t row_stride,
                                                                     size_t img_stride, uint8_t* __restrict__ slot_base,
                                                                     size_t stream_stride, size_t off0, int pitch0,
                                                                     size_t off1, int pitch1, int w, int h, int ow, int oh,
                                                                     const unsigned char* __restrict__ lut_g) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = blockIdx.z;
  __shared__ unsigned char s_lut[256];
  if (EQ) {
    reinterpret_cast<unsigned short*>(s_lut)[threadIdx.x] = reinterpret_cast<const unsigned short*>(lut_g + (size_t)s * 256)[threadIdx.x];
    __syncthreads();
  }
  const int oy0 = (blockIdx.y * FUSED_WARPS + warp) * FUSED_ROWS;
  if (oy0 >= oh) return;
  const int X = blockIdx.x * 128 + 4 * lane;          // first input column of this lane's word
  const bool act = X < w;                              // w % 4 == 0: a word is entirely inside or outside
  const uint8_t* img = src + (size_t)s * img_stride;
  uint8_t* L0 = slot_base + (size_t)s * stream_stride + off0;
  uint8_t* L1 = slot_base + (size_t)s * stream_stride + off1;
  const int y_lo = 2 * oy0, y_hi = min(2 * (oy0 + FUSED_ROWS), h);     // owned input rows (written to level 0)

  auto row_sums = [&](int y) -> unsigned {              // y may be outside [0,h): BORDER_REFLECT_101
    const int yr = y < 0 ? -y : (y >= h ? 2 * h - 2 - y : y);
    unsigned B = 0;
    if (act) B = *reinterpret_cast<const unsigned*>(img + (size_t)yr * row_stride + X);
    if (EQ) B = lut4(s_lut, B);
    unsigned A = __shfl_up_sync(0xffffffffu, B, 1);
    unsigned C = __shfl_down_sync(0xffffffffu, B, 1);
    if (act) {
      if (lane == 0) {
        if (X == 0) A = __byte_perm(B, 0, 0x1244);     // columns -2,-1 -> 2,1 into bytes 2,3
        else { A = *reinterpret_cast<const unsigned*>(img + (size_t)yr * row_stride + X - 4); if (EQ) A = lut4(s_lut, A); }
      }
      if (X + 4 >= w) C = (B >> 16) & 0xffu;            // column w -> w-2
      else if (lane == 31) { C = *reinterpret_cast<const unsigned*>(img + (size_t)yr * row_stride + X + 4); if (EQ) C = lut4(s_lut, C); }
      if (y == yr && y >= y_lo && y < y_hi) *reinterpret_cast<unsigned*>(L0 + (size_t)y * pitch0 + X) = B;
    }
    return hsum_pair(A, B, C);
  };

  unsigned h0 = row_sums(2 * oy0 - 2), h1 = row_sums(2 * oy0 - 1), h2 = row_sums(2 * oy0);
  const int ox = blockIdx.x * 64 + 2 * lane;
#pragma unroll 2
  for (int r = 0; r < FUSED_ROWS; ++r) {
    const int oy = oy0 + r;
    if (oy >= oh) break;
    const unsigned h3 = row_sums(2 * oy + 1), h4 = row_sums(2 * oy + 2);
    const unsigned v = h0 + 4u * h1 + 6u * h2 + 4u * h3 + h4 + 0x00800080u;
    if (act && ox < ow) {
      const unsigned short o2 = (unsigned short)(((v >> 8) & 0xffu) | ((v >> 16) & 0xff00u));
      if (ox + 1 < ow) *reinterpret_cast<unsigned short*>(L1 + (size_t)oy * pitch1 + ox) = o2;
      else L1[(size_t)oy * pitch1 + ox] = (uint8_t)(o2 & 0xff);
    }
    h0 = h2; h1 = h3; h2 = h4;
  }
}

// ---- cv::cvtColor(BGR/RGB/BGRA/RGBA -> GRAY) on ingest (src/frontend/f2f_tracking.cpp:78-110) --------------------------------
// OpenCV 4's 8-bit path (15-bit coefficients, probed against cv2 4.13): gray = (B * 3735 + G * 19235 + R * 9798 + (1 << 14)) >> 15.  One thread per pixel, interleaved
// source of `ch` (3 | 4) channels -> tight [S][h][w] landing area.
__global__ void __launch_bounds__(256) gray_kernel(const uint8_t* __restrict__ src, size_t row_stride, size_t img_stride, int ch,
                                                   int rgb, uint8_t* __restrict__ dst, int w, int h) {
  const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
  if (x >= w) return;
  const uint8_t* p = src + (size_t)blockIdx.z * img_stride + (size_t)y * row_stride + (size_t)x * ch;
  const int c0 = p[0], c1 = p[1], c2 = p[2];
  const int b = rgb ? c2 : c0, r = rgb ? c0 : c2;
  dst[(size_t)blockIdx.z * w * h + (size_t)y * w + x] = (uint8_t)((b * 3735 + c1 * 19235 + r * 9798 + (1 << 14)) >> 15);
}

// ---- cv::equalizeHist on ingest (need_equal_hist, src/frontend/f2f_tracking.cpp:125-145) -----------------------------
// OpenCV: hist over the whole image; i0 = first non-empty bin; if it holds every pixel the image becomes the constant i0;
// else scale = 255.f / (total - hist[i0]) and lut[i] = saturate_cast<uchar>(sum_{i0 < j <= i} hist[j] * scale) (float
// product, round half even).  Pinned to cv2.equalizeHist by tests/test_frontend_gpu.py.
constexpr int EQ_ROWS = 16;      // image rows per CTA

__global__ void __launch_bounds__(256) hist_kernel(const uint8_t* __restrict__ src, size_t row_stride, size_t img_stride, int w, int h,
                                                   int* __restrict__ hist) {
  __shared__ int sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const uint8_t* img = src + (size_t)blockIdx.y * img_stride;
  const int y0 = blockIdx.x * EQ_ROWS, y1 = min(y0 + EQ_ROWS, h);
  for (int y = y0; y < y1; ++y)
    for (int x = threadIdx.x; x < w; x += 256) atomicAdd(&sh[img[(size_t)y * row_stride + x]], 1);
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&hist[blockIdx.y * 256 + threadIdx.x], sh[threadIdx.x]);
}

// the LUT alone (one CTA per stream), for the fused ingest
__global__ void __launch_bounds__(256) lut_kernel(const int* __restrict__ hist, int total, unsigned char* __restrict__ lut_g) {
  __shared__ int sh[256];
  __shared__ int s_i0;
  const int t = threadIdx.x;
  sh[t] = hist[blockIdx.x * 256 + t];
  if (t == 0) s_i0 = 256;
  __syncthreads();
  if (sh[t]) atomicMin(&s_i0, t);
  __syncthreads();
  const int i0 = s_i0;
  unsigned char o;
  if (sh[i0] == total) o = (unsigned char)i0;
  else {
    const float scale = __fdiv_rn(255.f, (float)(total - sh[i0]));
    int sum = 0;
    for (int j = i0 + 1; j <= t; ++j) sum += sh[j];
    const int v = t <= i0 ? 0 : __float2int_rn(__fmul_rn((float)sum, scale));
    o = (unsigned char)(v > 255 ? 255 : v);
  }
  lut_g[blockIdx.x * 256 + t] = o;
}

__global__ void __launch_bounds__(256) equalize_apply_kernel(const uint8_t* __restrict__ src, size_t row_stride, size_t img_stride,
                                                             uint8_t* __restrict__ dst, int w, int h, const int* __restrict__ hist) {
  __shared__ int sh[256];
  __shared__ unsigned char lut[256];
  __shared__ int s_i0;
  const int t = threadIdx.x;
  sh[t] = hist[blockIdx.y * 256 + t];
  if (t == 0) s_i0 = 256;
  __syncthreads();
  if (sh[t]) atomicMin(&s_i0, t);
  __syncthreads();
  const int i0 = s_i0, total = w * h;
  if (sh[i0] == total) lut[t] = (unsigned char)i0;
  else {
    const float scale = __fdiv_rn(255.f, (float)(total - sh[i0]));
    int sum = 0;
    for (int j = i0 + 1; j <= t; ++j) sum += sh[j];
    int v = t <= i0 ? 0 : __float2int_rn(__fmul_rn((float)sum, scale));
    lut[t] = (unsigned char)(v > 255 ? 255 : v);
  }
  __syncthreads();
  const uint8_t* img = src + (size_t)blockIdx.y * img_stride;
  uint8_t* out = dst + (size_t)blockIdx.y * (size_t)w * h;        // tight [S][h][w]
  const int y0 = blockIdx.x * EQ_ROWS, y1 = min(y0 + EQ_ROWS, h);
  for (int y = y0; y < y1; ++y)
    for (int x = t; x < w; x += 256) out[(size_t)y * w + x] = lut[img[(size_t)y * row_stride + x]];
}

}  // namespace

int flv_launch_gray(flv_ctx* ctx, int n_streams, const uint8_t* d_src, size_t row_stride, size_t img_stride, int channels, int rgb,
                    uint8_t* d_dst_tight) {
  dim3 grid((ctx->w + 255) / 256, ctx->h, n_streams);
  gray_kernel<<<grid, 256, 0, ctx->stream>>>(d_src, row_stride, img_stride, channels, rgb, d_dst_tight, ctx->w, ctx->h);
  ctx->launches++;
  FLV_CUDA(ctx, cudaGetLastError());
  return FLV_OK;
}

int flv_launch_equalize(flv_ctx* ctx, int n_streams, const uint8_t* d_src, size_t row_stride, size_t img_stride, uint8_t* d_dst_tight) {
  if (!ctx->d_hist) FLV_CUDA(ctx, cudaMalloc(&ctx->d_hist, (size_t)ctx->S * 256 * sizeof(int)));
  FLV_CUDA(ctx, cudaMemsetAsync(ctx->d_hist, 0, (size_t)n_streams * 256 * sizeof(int), ctx->stream));
  dim3 grid((ctx->h + EQ_ROWS - 1) / EQ_ROWS, n_streams);
  hist_kernel<<<grid, 256, 0, ctx->stream>>>(d_src, row_stride, img_stride, ctx->w, ctx->h, ctx->d_hist);
  equalize_apply_kernel<<<grid, 256, 0, ctx->stream>>>(d_src, row_stride, img_stride, d_dst_tight, ctx->w, ctx->h, ctx->d_hist);
  ctx->launches += 2;
  FLV_CUDA(ctx, cudaGetLastError());
  return FLV_OK;
}

static bool fused_ingest_ok(const flv_ctx* ctx, const uint8_t* d_src, size_t row_stride, size_t img_stride) {
  return ctx->geom.nlev > 1 && ctx->w % 4 == 0 && row_stride % 4 == 0 && img_stride % 4 == 0 &&
         reinterpret_cast<size_t>(d_src) % 4 == 0 && !ctx->no_fused_ingest;
}

// cv::equalizeHist + ingest + first pyrDown without an equalised intermediate image: histogram -> LUT -> fused ingest that maps
// every source word through the LUT.  Returns FLV_ERR_UNSUPPORTED when the fused ingest cannot run on this source (the caller
// then equalises into a landing area and unpacks from there).
int flv_launch_unpack_equalized(flv_ctx* ctx, int slot, int n_streams, const uint8_t* d_src, size_t row_stride, size_t img_stride) {
  if (!fused_ingest_ok(ctx, d_src, row_stride, img_stride) || getenv("FLV_NO_FUSED_EQUALIZE")) return FLV_ERR_UNSUPPORTED;
  if (!ctx->d_hist) FLV_CUDA(ctx, cudaMalloc(&ctx->d_hist, (size_t)ctx->S * 256 * sizeof(int)));
  if (!ctx->d_lut) FLV_CUDA(ctx, cudaMalloc(&ctx->d_lut, (size_t)ctx->S * 256));
  FLV_CUDA(ctx, cudaMemsetAsync(ctx->d_hist, 0, (size_t)n_streams * 256 * sizeof(int), ctx->stream));
  dim3 hgrid((ctx->h + EQ_ROWS - 1) / EQ_ROWS, n_streams);
  hist_kernel<<<hgrid, 256, 0, ctx->stream>>>(d_src, row_stride, img_stride, ctx->w, ctx->h, ctx->d_hist);
  lut_kernel<<<n_streams, 256, 0, ctx->stream>>>(ctx->d_hist, ctx->w * ctx->h, (unsigned char*)ctx->d_lut);
  const LevelGeom& L0 = ctx->geom.lv[0];
  const LevelGeom& L1 = ctx->geom.lv[1];
  ctx->deriv_streams[slot] = 0;
  dim3 grid((ctx->w + 127) / 128, (L1.h + FUSED_ROWS * FUSED_WARPS - 1) / (FUSED_ROWS * FUSED_WARPS), n_streams);
  ingest_l1_kernel<true><<<grid, FUSED_WARPS * 32, 0, ctx->stream>>>(d_src, row_stride, img_stride, ctx->pyr[slot],
                                                                     ctx->geom.stream_stride, L0.off, L0.pitch, L1.off, L1.pitch,
                                                                     ctx->w, ctx->h, L1.w, L1.h, (const unsigned char*)ctx->d_lut);
  ctx->launches += 3;
  FLV_CUDA(ctx, cudaGetLastError());
  ctx->l1_valid[slot] = n_streams;        // streams whose level 1 came out of the fused ingest
  return FLV_OK;
}

int flv_launch_unpack(flv_ctx* ctx, int slot, int n_streams, const uint8_t* d_src, size_t row_stride,
                      size_t img_stride) {
  const LevelGeom& L0 = ctx->geom.lv[0];
  ctx->l1_valid[slot] = 0;
  ctx->deriv_streams[slot] = 0;
  if (fused_ingest_ok(ctx, d_src, row_stride, img_stride)) {
    const LevelGeom& L1 = ctx->geom.lv[1];
    dim3 grid((ctx->w + 127) / 128, (L1.h + FUSED_ROWS * FUSED_WARPS - 1) / (FUSED_ROWS * FUSED_WARPS), n_streams);
    ingest_l1_kernel<false><<<grid, FUSED_WARPS * 32, 0, ctx->stream>>>(d_src, row_stride, img_stride, ctx->pyr[slot],
                                                                 ctx->geom.stream_stride, L0.off, L0.pitch, L1.off, L1.pitch,
                                                                 ctx->w, ctx->h, L1.w, L1.h, nullptr);
    ctx->launches++;
    FLV_CUDA(ctx, cudaGetLastError());
    ctx->l1_valid[slot] = n_streams;        // streams whose level 1 came out of the fused ingest
    return FLV_OK;
  }
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
  ctx->deriv_streams[slot] = 0;
  const int first = ctx->l1_valid[slot] >= n_streams ? 2 : 1;      // level 1 already built by the fused ingest kernel (for all of these streams)
  ctx->l1_valid[slot] = 0;
  for (int l = first; l < g.nlev; ++l) {
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
