// A/B baseline (first version of the tracker kept for kernel comparisons; FLV_LK_VARIANT=1).
// K2/K3 -- pyramidal Lucas-Kanade tracker, one warp per (stream, point), all levels in one launch.
//
// Reference: cv::calcOpticalFlowPyrLK as called at src/processing/lkorb_tracking.cpp:64-73
// (frame->frame) and src/processing/camera_frame.cpp:124-128 (left->right); algorithm restated in
// SURVEY.md Appendix A.1 and oracle/lk_ref.py (the bit-level contract for this kernel).
//
// Arithmetic contract (must equal oracle/lk_ref.py bit for bit):
//   * template patch Ival/Ix/Iy: 14-bit fixed-point bilinear, int16-range, exactly OpenCV's;
//   * A11/A12/A22/b1/b2/err window sums are exact integers (per-lane int32 partials that cannot
//     overflow, combined with redux.sync on 16-bit halves), converted to f32 once;
//   * the f32 tail is compiled with -fmad=false: every op is an individually rounded IEEE op.
//
// Mapping: lane c owns window column c (0..30; lane 31 only feeds the right bilinear tap of lane
// 30).  The 31x31 template (Ival, Ix, Iy) lives in registers (fully unrolled rows), the J patch of
// every iteration is read as 32 row-bytes per lane through L1 (rows of 32 consecutive bytes ->
// 1-2 sectors per request).  Border handling is index reflection (BORDER_REFLECT_101 for
// intensities, zero derivative outside the image) instead of OpenCV's padded copies, so the
// kernel reads exactly the pyramid bytes (2P + 29N algorithmic bytes per call).
#include "ctx.h"

namespace {

constexpr int WIN = 31;
constexpr int WB = 14;
constexpr unsigned FULL = 0xffffffffu;

struct LKGeom {
  int nlev;
  int w[FLV_MAX_LEVELS], h[FLV_MAX_LEVELS], pitch[FLV_MAX_LEVELS];
  unsigned long long off[FLV_MAX_LEVELS];
  unsigned long long stream_stride;
};

__device__ __forceinline__ int reflect101(int i, int n) {
  i = i < 0 ? -i : i;
  return i >= n ? 2 * n - 2 - i : i;
}

// exact warp sum of int32 partials (|total| may exceed 2^31): split into 16-bit halves
__device__ __forceinline__ long long warp_sum_exact(int p) {
  unsigned lo = (unsigned)p & 0xffffu;
  int hi = p >> 16;
  unsigned slo = __reduce_add_sync(FULL, lo);
  int shi = __reduce_add_sync(FULL, hi);
  return ((long long)shi << 16) + (long long)slo;
}

__device__ __forceinline__ void bilinear_weights(float a, float b, int& w00, int& w01, int& w10,
                                                 int& w11) {
  const float s = (float)(1 << WB);
  float oa = 1.f - a, ob = 1.f - b;
  w00 = __float2int_rn((oa * ob) * s);
  w01 = __float2int_rn((a * ob) * s);
  w10 = __float2int_rn((oa * b) * s);
  w11 = (1 << WB) - w00 - w01 - w10;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
lk_track_kernel_v1(const uint8_t* __restrict__ pyrI, const uint8_t* __restrict__ pyrJ, LKGeom g,
                const int* __restrict__ npts, const float* __restrict__ prev_xy,
                const float* __restrict__ init_xy, float* __restrict__ next_xy,
                uint8_t* __restrict__ status, float* __restrict__ err, int max_pts, int nlev_used,
                int max_iter, double eps2, double min_eig_thr, float err_scale) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int s = blockIdx.y;
  const int pt = blockIdx.x * WARPS + warp;
  if (pt >= npts[s]) return;
  const size_t pidx = (size_t)s * max_pts + pt;
  const uint8_t* Ibase = pyrI + (size_t)s * g.stream_stride;
  const uint8_t* Jbase = pyrJ + (size_t)s * g.stream_stride;
  const float px0 = prev_xy[2 * pidx], py0 = prev_xy[2 * pidx + 1];
  float nx = init_xy[2 * pidx], ny = init_xy[2 * pidx + 1];
  int st = 1;
  float er = 0.f;
  const float half = (float)((WIN - 1) * 0.5);
  const float FLT_SCALE = 1.f / (float)(1 << 20);
  const bool live = lane < WIN;

  for (int level = nlev_used - 1; level >= 0; --level) {
    const int w = g.w[level], h = g.h[level], pitch = g.pitch[level];
    const uint8_t* I = Ibase + g.off[level];
    const uint8_t* J = Jbase + g.off[level];
    const float sc = 1.f / (float)(1 << level);
    float px = px0 * sc, py = py0 * sc;
    if (level == nlev_used - 1) { nx = nx * sc; ny = ny * sc; }
    else { nx = nx * 2.f; ny = ny * 2.f; }
    px = px - half; py = py - half;
    const int ipx = (int)floorf(px), ipy = (int)floorf(py);
    if (ipx < -WIN || ipx >= w || ipy < -WIN || ipy >= h) {
      if (level == 0) { st = 0; er = 0.f; }
      continue;
    }
    int w00, w01, w10, w11;
    bilinear_weights(px - (float)ipx, py - (float)ipy, w00, w01, w10, w11);

    // ---- template patch: Ival (5 guard bits), Ix, Iy (Scharr, int16 range) -------------------
    int tI[WIN], tX[WIN], tY[WIN];
    int a11 = 0, a12 = 0, a22 = 0;
    {
      const int x = ipx + lane;
      const bool in_x = (x >= 0) && (x < w);
      const int xi = reflect101(x, w);
      const int xm = reflect101(xi - 1, w), xp = reflect101(xi + 1, w);
      // rows y-1 (p), y (c), y+1 (n) of the 3 columns xm, xi, xp
      const uint8_t* rp = I + (size_t)reflect101(ipy - 1, h) * pitch;
      const uint8_t* rc = I + (size_t)reflect101(ipy, h) * pitch;
      int pm = rp[xm], p0 = rp[xi], pp = rp[xp];
      int cm = rc[xm], c0 = rc[xi], cp = rc[xp];
      int qI = 0, qX = 0, qY = 0, qIr = 0, qXr = 0, qYr = 0;   // previous row (own, right)
#pragma unroll
      for (int k = 0; k <= WIN; ++k) {
        const int y = ipy + k;
        const uint8_t* rn = I + (size_t)reflect101(y + 1, h) * pitch;
        const int nm = rn[xm], n0 = rn[xi], np = rn[xp];
        const bool in = in_x && (y >= 0) && (y < h);
        int vX = 3 * (pp - pm) + 10 * (cp - cm) + 3 * (np - nm);
        int vY = 3 * (nm - pm) + 10 * (n0 - p0) + 3 * (np - pp);
        vX = in ? vX : 0;
        vY = in ? vY : 0;
        const int vI = c0;
        const int vIr = __shfl_down_sync(FULL, vI, 1);
        const int vXr = __shfl_down_sync(FULL, vX, 1);
        const int vYr = __shfl_down_sync(FULL, vY, 1);
        if (k >= 1) {
          int iv = (qI * w00 + qIr * w01 + vI * w10 + vIr * w11 + (1 << (WB - 5 - 1))) >> (WB - 5);
          int ix = (qX * w00 + qXr * w01 + vX * w10 + vXr * w11 + (1 << (WB - 1))) >> WB;
          int iy = (qY * w00 + qYr * w01 + vY * w10 + vYr * w11 + (1 << (WB - 1))) >> WB;
          iv = live ? iv : 0; ix = live ? ix : 0; iy = live ? iy : 0;
          tI[k - 1] = iv; tX[k - 1] = ix; tY[k - 1] = iy;
          a11 += ix * ix; a12 += ix * iy; a22 += iy * iy;
        }
        qI = vI; qX = vX; qY = vY; qIr = vIr; qXr = vXr; qYr = vYr;
        pm = cm; p0 = c0; pp = cp;
        cm = nm; c0 = n0; cp = np;
      }
    }
    const float A11 = __ll2float_rn(warp_sum_exact(a11)) * FLT_SCALE;
    const float A12 = __ll2float_rn(warp_sum_exact(a12)) * FLT_SCALE;
    const float A22 = __ll2float_rn(warp_sum_exact(a22)) * FLT_SCALE;
    float D = A11 * A22 - A12 * A12;
    const float dA = A11 - A22;
    const float disc = dA * dA + (4.f * A12) * A12;
    const float min_eig = ((A22 + A11) - sqrtf(disc)) / (float)(2 * WIN * WIN);
    if ((double)min_eig < min_eig_thr || (double)D < 1.1920928955078125e-07) {
      if (level == 0) st = 0;
      continue;
    }
    D = 1.f / D;
    float cx = nx - half, cy = ny - half;
    float pdx = 0.f, pdy = 0.f;
    for (int j = 0; j < max_iter; ++j) {
      const int inx = (int)floorf(cx), iny = (int)floorf(cy);
      if (inx < -WIN || inx >= w || iny < -WIN || iny >= h) {
        if (level == 0) st = 0;
        break;
      }
      bilinear_weights(cx - (float)inx, cy - (float)iny, w00, w01, w10, w11);
      const int xj = reflect101(inx + lane, w);
      int b1 = 0, b2 = 0;
      int q = J[(size_t)reflect101(iny, h) * pitch + xj];
      int qr = __shfl_down_sync(FULL, q, 1);
#pragma unroll
      for (int r = 1; r <= WIN; ++r) {
        const int v = J[(size_t)reflect101(iny + r, h) * pitch + xj];
        const int vr = __shfl_down_sync(FULL, v, 1);
        const int jv = (q * w00 + qr * w01 + v * w10 + vr * w11 + (1 << (WB - 5 - 1))) >> (WB - 5);
        const int diff = jv - tI[r - 1];
        b1 += diff * tX[r - 1];
        b2 += diff * tY[r - 1];
        q = v; qr = vr;
      }
      const float fb1 = __ll2float_rn(warp_sum_exact(b1)) * FLT_SCALE;
      const float fb2 = __ll2float_rn(warp_sum_exact(b2)) * FLT_SCALE;
      const float dx = (A12 * fb2 - A22 * fb1) * D;
      const float dy = (A12 * fb1 - A11 * fb2) * D;
      cx = cx + dx; cy = cy + dy;
      nx = cx + half; ny = cy + half;
      if ((double)dx * (double)dx + (double)dy * (double)dy <= eps2) break;
      if (j > 0 && fabs((double)(dx + pdx)) < 0.01 && fabs((double)(dy + pdy)) < 0.01) {
        nx = nx - dx * 0.5f;
        ny = ny - dy * 0.5f;
        break;
      }
      pdx = dx; pdy = dy;
    }
    if (level == 0 && st) {
      const float qx = nx - half, qy = ny - half;
      const int inx = (int)floorf(qx), iny = (int)floorf(qy);
      if (inx < -WIN || inx >= w || iny < -WIN || iny >= h) {
        st = 0;
      } else {
        bilinear_weights(qx - (float)inx, qy - (float)iny, w00, w01, w10, w11);
        const int xj = reflect101(inx + lane, w);
        int e = 0;
        int q = J[(size_t)reflect101(iny, h) * pitch + xj];
        int qr = __shfl_down_sync(FULL, q, 1);
#pragma unroll
        for (int r = 1; r <= WIN; ++r) {
          const int v = J[(size_t)reflect101(iny + r, h) * pitch + xj];
          const int vr = __shfl_down_sync(FULL, v, 1);
          const int jv = (q * w00 + qr * w01 + v * w10 + vr * w11 + (1 << (WB - 5 - 1))) >> (WB - 5);
          const int diff = live ? (jv - tI[r - 1]) : 0;
          e += diff < 0 ? -diff : diff;
          q = v; qr = vr;
        }
        er = __ll2float_rn(warp_sum_exact(e)) * err_scale;
      }
    }
  }
  if (lane == 0) {
    next_xy[2 * pidx] = nx;
    next_xy[2 * pidx + 1] = ny;
    status[pidx] = (uint8_t)st;
    err[pidx] = er;
  }
}

}  // namespace

int flv_launch_lk_v1(flv_ctx* ctx, int src_slot, int dst_slot, int n_streams, const int* d_npts,
                  const float* d_prev, const float* d_init, float* d_next, uint8_t* d_status,
                  float* d_err, int nlev_used, int max_iter, double eps2, double min_eig_thr) {
  LKGeom g;
  g.nlev = ctx->geom.nlev;
  for (int l = 0; l < FLV_MAX_LEVELS; ++l) {
    g.w[l] = ctx->geom.lv[l].w; g.h[l] = ctx->geom.lv[l].h; g.pitch[l] = ctx->geom.lv[l].pitch;
    g.off[l] = ctx->geom.lv[l].off;
  }
  g.stream_stride = ctx->geom.stream_stride;
  constexpr int WARPS = 4;
  dim3 grid((ctx->max_pts + WARPS - 1) / WARPS, n_streams);
  const float err_scale = (float)(1.0 / (32 * WIN * WIN));
  lk_track_kernel_v1<WARPS><<<grid, WARPS * 32, 0, ctx->stream>>>(
      ctx->pyr[src_slot], ctx->pyr[dst_slot], g, d_npts, d_prev, d_init, d_next, d_status, d_err,
      ctx->max_pts, nlev_used, max_iter, eps2, min_eig_thr, err_scale);
  ctx->launches++;
  FLV_CUDA(ctx, cudaGetLastError());
  return FLV_OK;
}
