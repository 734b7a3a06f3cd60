// K4 -- Shi-Tomasi corner detection == cv::goodFeaturesToTrack(img, N, q, d) (blockSize 3, Sobel 3).
//
// Reference call sites: src/processing/feature_dem.cpp:160 (redetect) and :221 (detect).
// Bit-level contract: oracle/gftt_ref.py (pinned to cv2 4.13.0 by tests/golden/gftt_*.npz).
//
//   mineig_kernel   u8 image -> f32 min-eigenvalue map + per-stream max   (HBM: read w*h, write 4*w*h)
//   nms_kernel      threshold q*max, 3x3 non-max test, emit (value,y,x) keys
//   mindist_kernel  one CTA per stream: bin keys into d x d cells, decide OpenCV's sequential greedy
//                   "accept if no stronger accepted corner within d" as a parallel fixed point
//                   (a corner's fate depends only on stronger corners, so iterating "reject if a
//                   stronger neighbour is accepted / accept if all stronger neighbours are rejected"
//                   reproduces the sequential result exactly), then bitonic-sort the accepted keys
//                   and keep the strongest N.
//
// Float contract of the response map (found by probing cv2, see oracle/gftt_ref.py header): explicit
// __fmaf_rn where OpenCV's AVX2 body contracts, plain mul/add in its scalar tail (x >= w - w%32).
#include "ctx.h"

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int reflect101(int i, int n) {
  i = i < 0 ? -i : i;
  return i >= n ? 2 * n - 2 - i : i;
}

// ------------------------------------------------------------------------------------------------
constexpr int ETW = 64, ETH = 16;            // output tile
constexpr int ITW = ETW + 4, ITH = ETH + 4;  // image tile (halo 2)
constexpr int CTW = ETW + 2, CTH = ETH + 2;  // covariance tile (halo 1)

__global__ void __launch_bounds__(256)
mineig_kernel(const uint8_t* __restrict__ img_base, size_t stream_stride, int pitch, int w, int h,
              float* __restrict__ eig_base, int* __restrict__ eigmax) {
  __shared__ uint8_t It[ITH][ITW + 4];
  __shared__ float cxx[CTH][CTW + 1], cxy[CTH][CTW + 1], cyy[CTH][CTW + 1];
  __shared__ int smax[8];
  const int s = blockIdx.z;
  const uint8_t* img = img_base + (size_t)s * stream_stride;
  float* eig = eig_base + (size_t)s * w * h;
  const int x0 = blockIdx.x * ETW, y0 = blockIdx.y * ETH;
  const int tid = threadIdx.x;
  for (int i = tid; i < ITH * ITW; i += 256) {
    int r = i / ITW, c = i - r * ITW;
    int y = reflect101(y0 - 2 + r, h), x = reflect101(x0 - 2 + c, w);
    // tiles hanging far over the right/bottom edge can reflect twice; clamp keeps loads in range
    y = min(max(y, 0), h - 1); x = min(max(x, 0), w - 1);
    It[r][c] = img[(size_t)y * pitch + x];
  }
  __syncthreads();
  const float k1 = (float)(1.0 / (4 * 3 * 255.0));
  const float k0 = 2.f * k1;
  const int body = w - (w & 31);
  for (int i = tid; i < CTH * CTW; i += 256) {
    int r = i / CTW, c = i - r * CTW;
    int px = x0 - 1 + c, py = y0 - 1 + r;
    float vxx = 0.f, vxy = 0.f, vyy = 0.f;
    if (px <= w && py <= h) {
      // covariance at the REFLECTED coordinate (boxFilter border), computed there from scratch
      int qx = reflect101(px, w), qy = reflect101(py, h);
      int cxm = reflect101(qx - 1, w) - (x0 - 2), cx0 = qx - (x0 - 2), cxp = reflect101(qx + 1, w) - (x0 - 2);
      int rym = reflect101(qy - 1, h) - (y0 - 2), ry0 = qy - (y0 - 2), ryp = reflect101(qy + 1, h) - (y0 - 2);
      int a_m = It[rym][cxm], a_0 = It[rym][cx0], a_p = It[rym][cxp];
      int b_m = It[ry0][cxm], b_0 = It[ry0][cx0], b_p = It[ry0][cxp];
      int c_m = It[ryp][cxm], c_0 = It[ryp][cx0], c_p = It[ryp][cxp];
      // Dx: rows of [-1 0 1] (exact), columns k1*[1 2 1] as fma(k1, r(y-1)+r(y+1), k0*r(y))
      float dx = __fmaf_rn(k1, (float)((a_p - a_m) + (c_p - c_m)), __fmul_rn(k0, (float)(b_p - b_m)));
      // Dy: rows smoothed with k1*[1 2 1] (FMA body / plain tail), then row difference
      float sa, sc;
      if (qx < body) {
        sa = __fmaf_rn(k1, (float)a_p, __fmaf_rn(k0, (float)a_0, __fmul_rn(k1, (float)a_m)));
        sc = __fmaf_rn(k1, (float)c_p, __fmaf_rn(k0, (float)c_0, __fmul_rn(k1, (float)c_m)));
      } else {
        sa = __fadd_rn(__fadd_rn(__fmul_rn(k1, (float)a_m), __fmul_rn(k0, (float)a_0)), __fmul_rn(k1, (float)a_p));
        sc = __fadd_rn(__fadd_rn(__fmul_rn(k1, (float)c_m), __fmul_rn(k0, (float)c_0)), __fmul_rn(k1, (float)c_p));
      }
      float dy = __fsub_rn(sc, sa);
      vxx = __fmul_rn(dx, dx); vxy = __fmul_rn(dx, dy); vyy = __fmul_rn(dy, dy);
    }
    cxx[r][c] = vxx; cxy[r][c] = vxy; cyy[r][c] = vyy;
  }
  __syncthreads();
  float best = -1.f;
  for (int i = tid; i < ETH * ETW; i += 256) {
    int r = i / ETW, c = i - r * ETW;
    int x = x0 + c, y = y0 + r;
    if (x < w && y < h) {
      double sxx = 0, sxy = 0, syy = 0;
#pragma unroll
      for (int dr = 0; dr < 3; ++dr)
#pragma unroll
        for (int dc = 0; dc < 3; ++dc) {
          sxx += (double)cxx[r + dr][c + dc];
          sxy += (double)cxy[r + dr][c + dc];
          syy += (double)cyy[r + dr][c + dc];
        }
      float a = __fmul_rn((float)sxx, 0.5f), b = (float)sxy, cc = __fmul_rn((float)syy, 0.5f);
      float t = __fsub_rn(a, cc);
      float e = __fsub_rn(__fadd_rn(a, cc), __fsqrt_rn(__fadd_rn(__fmul_rn(t, t), __fmul_rn(b, b))));
      eig[(size_t)y * w + x] = e;
      best = fmaxf(best, e);
    }
  }
  // block max -> atomicMax on the int view (valid ordering for non-negative floats; negatives lose)
  int bi = __float_as_int(fmaxf(best, 0.f));
  bi = __reduce_max_sync(FULL, bi);
  if ((tid & 31) == 0) smax[tid >> 5] = bi;
  __syncthreads();
  if (tid < 8) {
    int v = smax[tid];
    v = max(v, __shfl_xor_sync(0xffu, v, 4));
    v = max(v, __shfl_xor_sync(0xffu, v, 2));
    v = max(v, __shfl_xor_sync(0xffu, v, 1));
    if (tid == 0) atomicMax(&eigmax[s], v);
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
nms_kernel(const float* __restrict__ eig_base, int w, int h, const int* __restrict__ eigmax,
           double quality, unsigned long long* __restrict__ cand_base, int* __restrict__ ncand,
           int cand_cap) {
  const int s = blockIdx.z;
  const float* eig = eig_base + (size_t)s * w * h;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const float thr = (float)((double)__int_as_float(eigmax[s]) * quality);
  bool is = false;
  float c = 0.f;
  if (x >= 1 && x < w - 1 && y >= 1 && y < h - 1) {
    c = eig[(size_t)y * w + x];
    if (c > thr) {
      const float* p = eig + (size_t)(y - 1) * w + x;
      float m = fmaxf(fmaxf(p[-1], p[0]), p[1]);
      p += w; m = fmaxf(m, fmaxf(p[-1], p[1]));
      p += w; m = fmaxf(m, fmaxf(fmaxf(p[-1], p[0]), p[1]));
      is = c >= m;
    }
  }
  unsigned mask = __ballot_sync(FULL, is);
  if (mask) {
    int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(&ncand[s], __popc(mask));
    base = __shfl_sync(FULL, base, 0);
    if (is) {
      int pos = base + __popc(mask & ((1u << lane) - 1));
      if (pos < cand_cap)
        cand_base[(size_t)s * cand_cap + pos] =
            ((unsigned long long)__float_as_uint(c) << 32) | ((unsigned)y << 16) | (unsigned)x;
    }
  }
}

// ------------------------------------------------------------------------------------------------
constexpr int MD_THREADS = 1024;
constexpr int ACC_CAP = 8192;   // accepted corners that can be sorted (packing bound of d >= 8 on 1241x376)

__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int* total) {
  // v: per-thread value; returns exclusive prefix over the block (MD_THREADS threads)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int ws = warp_sums[lane];
    int winc = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(FULL, winc, o);
      if (lane >= o) winc += t;
    }
    warp_sums[lane] = winc - ws;
    if (lane == 31) *total = winc;
  }
  __syncthreads();
  return warp_sums[wid] + inc - v;
}

__global__ void __launch_bounds__(MD_THREADS, 1)
mindist_kernel(const unsigned long long* __restrict__ cand_base, const int* __restrict__ ncand,
               int cand_cap, unsigned long long* __restrict__ sorted_base, int w, int h,
               double min_distance, int max_corners, float* __restrict__ corners_base,
               int* __restrict__ ncorners, int corner_stride, int* __restrict__ flags,
               int max_cells) {
  extern __shared__ unsigned char smem_raw[];
  // layout: acc keys [ACC_CAP] u64 | cell start [max_cells+1] int | cell count [max_cells] int | state [cand_cap] u8
  unsigned long long* acc = (unsigned long long*)smem_raw;
  int* cstart = (int*)(acc + ACC_CAP);
  int* ccount = cstart + (max_cells + 1);
  unsigned char* state = (unsigned char*)(ccount + max_cells);
  __shared__ int warp_sums[32];
  __shared__ int sh_total, sh_nacc;

  const int s = blockIdx.x;
  const int tid = threadIdx.x;
  const unsigned long long* cand = cand_base + (size_t)s * cand_cap;
  unsigned long long* skey = sorted_base + (size_t)s * cand_cap;
  int n = ncand[s];
  if (n > cand_cap) { n = cand_cap; if (tid == 0) atomicOr(&flags[s], 1); }
  float* corners = corners_base + (size_t)s * corner_stride * 2;

  if (tid == 0) sh_nacc = 0;
  const bool use_dist = min_distance >= 1.0;
  if (use_dist) {
    const int cell = (int)rint(min_distance);
    const int gw = (w + cell - 1) / cell, gh = (h + cell - 1) / cell;
    const int ncell = gw * gh;
    const double d2 = min_distance * min_distance;
    for (int i = tid; i < ncell; i += MD_THREADS) ccount[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += MD_THREADS) {
      unsigned lo = (unsigned)cand[i];
      int x = lo & 0xffff, y = lo >> 16;
      atomicAdd(&ccount[(y / cell) * gw + x / cell], 1);
    }
    __syncthreads();
    // exclusive scan of ccount -> cstart (each thread owns a contiguous chunk of cells)
    const int per = (ncell + MD_THREADS - 1) / MD_THREADS;
    int local = 0;
    for (int k = 0; k < per; ++k) { int c = tid * per + k; if (c < ncell) local += ccount[c]; }
    int ex = block_exclusive_scan(local, warp_sums, &sh_total);
    for (int k = 0; k < per; ++k) {
      int c = tid * per + k;
      if (c < ncell) { cstart[c] = ex; ex += ccount[c]; ccount[c] = 0; }
    }
    if (tid == 0) cstart[ncell] = n;
    __syncthreads();
    for (int i = tid; i < n; i += MD_THREADS) {
      unsigned long long k = cand[i];
      unsigned lo = (unsigned)k;
      int x = lo & 0xffff, y = lo >> 16;
      int c = (y / cell) * gw + x / cell;
      int pos = cstart[c] + atomicAdd(&ccount[c], 1);
      skey[pos] = k;
      state[pos] = 0;
    }
    __syncthreads();
    // fixed-point iteration of the greedy rule
    for (;;) {
      int pending = 0;
      for (int i = tid; i < n; i += MD_THREADS) {
        if (state[i] != 0) continue;
        const unsigned long long ki = skey[i];
        const unsigned lo = (unsigned)ki;
        const int x = lo & 0xffff, y = lo >> 16;
        const int xc = x / cell, yc = y / cell;
        const int x1 = max(0, xc - 1), x2 = min(gw - 1, xc + 1);
        const int y1 = max(0, yc - 1), y2 = min(gh - 1, yc + 1);
        int verdict = 1;   // 1 accept, 2 reject, 0 wait
        for (int yy = y1; yy <= y2 && verdict != 2; ++yy) {
          const int jb = cstart[yy * gw + x1], je = cstart[yy * gw + x2 + 1];
          for (int j = jb; j < je; ++j) {
            const unsigned long long kj = skey[j];
            if (kj <= ki) continue;
            const unsigned lj = (unsigned)kj;
            const int dx = x - (int)(lj & 0xffff), dy = y - (int)(lj >> 16);
            if ((double)(dx * dx + dy * dy) < d2) {
              const int sj = ((volatile unsigned char*)state)[j];
              if (sj == 1) { verdict = 2; break; }
              if (sj == 0) verdict = 0;
            }
          }
        }
        if (verdict == 0) pending = 1;
        else ((volatile unsigned char*)state)[i] = (unsigned char)verdict;
      }
      if (!__syncthreads_or(pending)) break;
    }
    // gather accepted keys
    for (int i = tid; i < n; i += MD_THREADS) {
      if (state[i] == 1) {
        int p = atomicAdd(&sh_nacc, 1);
        if (p < ACC_CAP) acc[p] = skey[i];
      }
    }
  } else {
    __syncthreads();
    for (int i = tid; i < n; i += MD_THREADS) {
      int p = atomicAdd(&sh_nacc, 1);
      if (p < ACC_CAP) acc[p] = cand[i];
    }
  }
  __syncthreads();
  int nacc = sh_nacc;
  if (nacc > ACC_CAP) { nacc = ACC_CAP; if (tid == 0) atomicOr(&flags[s], 2); }
  int np2 = 1;
  while (np2 < nacc) np2 <<= 1;
  for (int i = nacc + tid; i < np2; i += MD_THREADS) acc[i] = 0ull;
  __syncthreads();
  // bitonic sort, descending
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < np2; i += MD_THREADS) {
        int l = i ^ j;
        if (l > i) {
          unsigned long long a = acc[i], b = acc[l];
          bool desc = (i & k) == 0;
          if (desc ? (a < b) : (a > b)) { acc[i] = b; acc[l] = a; }
        }
      }
      __syncthreads();
    }
  }
  int nout = nacc;
  if (max_corners > 0 && nout > max_corners) nout = max_corners;
  if (nout > corner_stride) nout = corner_stride;
  for (int i = tid; i < nout; i += MD_THREADS) {
    unsigned lo = (unsigned)acc[i];
    corners[2 * i] = (float)(lo & 0xffff);
    corners[2 * i + 1] = (float)(lo >> 16);
  }
  if (tid == 0) ncorners[s] = nout;
}

__global__ void gftt_reset_kernel(int* eigmax, int* ncand, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { eigmax[i] = 0; ncand[i] = 0; }
}

}  // namespace

size_t flv_mindist_smem(int cand_cap, int max_cells) {
  return (size_t)ACC_CAP * 8 + (size_t)(2 * max_cells + 1) * 4 + (size_t)cand_cap;
}

int flv_launch_gftt(flv_ctx* ctx, int slot, int n_streams, int max_corners, double quality,
                    double min_distance) {
  const int w = ctx->w, h = ctx->h;
  if (min_distance >= 1.0) {
    int cell = (int)rint(min_distance);
    int ncell = ((w + cell - 1) / cell) * ((h + cell - 1) / cell);
    if (ncell > ctx->max_cells)
      FLV_FAIL(ctx, FLV_ERR_UNSUPPORTED, "min_distance %.2f needs %d cells > capacity %d", min_distance,
               ncell, ctx->max_cells);
  }
  if (max_corners > ctx->gftt_cap || max_corners <= 0)
    FLV_FAIL(ctx, FLV_ERR_INVALID, "max_corners %d outside (0, %d]", max_corners, ctx->gftt_cap);
  gftt_reset_kernel<<<(n_streams + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_eigmax, ctx->d_ncand, n_streams);
  dim3 g1((w + ETW - 1) / ETW, (h + ETH - 1) / ETH, n_streams);
  mineig_kernel<<<g1, 256, 0, ctx->stream>>>(ctx->pyr[slot] + ctx->geom.lv[0].off, ctx->geom.stream_stride,
                                             ctx->geom.lv[0].pitch, w, h, ctx->d_eig, ctx->d_eigmax);
  dim3 g2((w + 31) / 32, (h + 7) / 8, n_streams);
  nms_kernel<<<g2, 256, 0, ctx->stream>>>(ctx->d_eig, w, h, ctx->d_eigmax, quality, ctx->d_cand,
                                          ctx->d_ncand, ctx->cand_cap);
  size_t smem = flv_mindist_smem(ctx->cand_cap, ctx->max_cells);
  mindist_kernel<<<n_streams, MD_THREADS, smem, ctx->stream>>>(
      ctx->d_cand, ctx->d_ncand, ctx->cand_cap, ctx->d_sorted, w, h,
      min_distance, max_corners, ctx->d_corners, ctx->d_ncorners, ctx->gftt_cap, ctx->d_flags,
      ctx->max_cells);
  ctx->launches += 4;
  FLV_CUDA(ctx, cudaGetLastError());
  return FLV_OK;
}

int flv_gftt_init(flv_ctx* ctx) {
  size_t smem = flv_mindist_smem(ctx->cand_cap, ctx->max_cells);
  FLV_CUDA(ctx, cudaFuncSetAttribute(mindist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return FLV_OK;
}
