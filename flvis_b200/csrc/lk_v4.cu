// K2/K3 -- pyramidal Lucas-Kanade tracker, fourth version (FLV_LK_VARIANT=6).
//
// Same mapping and the same arithmetic contract as lk.cu / lk_v3.cu / oracle/lk_ref.py (one warp per (stream, point),
// lane = window column, exact integer window sums, every integer expression evaluated exactly), restructured around
// what the instruction-level profile of v3 showed (profiles/README.md): 33 % of the instructions built the template
// (Scharr derivatives recomputed per point and level) and the window pass spent 16 instructions per row, half of
// them on addressing and on the four bilinear multiplies.
//   * the Scharr derivative of the TEMPLATE image is computed once per frame and level (scharr_kernel, below) --
//     OpenCV does the same (calcSharrDeriv over the whole level) -- and reused by every point and by both LK calls
//     that use the frame as their first image (left->right now, frame->frame on the next frame);
//   * the 32x32 u8 patch of the second image is held in registers as packed columns (lane = column, 4 rows per
//     register); a window pass needs NO memory access for it, and the two vertical taps of the bilinear blend are
//     adjacent bytes, so one DP2A (16-bit weights x 8-bit pixels) does two multiplies: 2 DP2A per pixel instead of
//     4 IMAD + load + shuffle.  The patch is only reloaded when the integer window origin moves.
// Reference call sites: src/processing/lkorb_tracking.cpp:64-73, src/processing/camera_frame.cpp:124-128.
//
// FLV_LK_VARIANT=7 (the default since it measured faster; 6 = plain loads) is the TMA form of the patch staging: the 32x32 u8 patch of the second image comes through a
// cp.async.bulk.tensor.3d load (one tensor map per pyramid level: x, y, stream; box 32x32x1) into a per-warp shared-memory
// tile (48 x 32: tile loads start on 16-byte boundaries) and is packed into the same registers from there; everything else is
// identical.  Numbers: profiles/README.md.
#include <stdlib.h>
#include <cuda.h>
#include "ctx.h"

namespace {

constexpr int WIN = 31;
constexpr int WB = 14;
constexpr unsigned FULL = 0xffffffffu;
constexpr int V4_WARPS = 4;
constexpr int TMPL_WORDS = WIN * 32;          // int2 per (row, lane)

struct LKGeom {
  int nlev;
  int w[FLV_MAX_LEVELS], h[FLV_MAX_LEVELS], pitch[FLV_MAX_LEVELS];
  unsigned long long off[FLV_MAX_LEVELS];
  unsigned long long stream_stride;
  int blk0[FLV_MAX_LEVELS + 1];               // scharr_kernel: first block of each level
  int strips[FLV_MAX_LEVELS];
};

__device__ __forceinline__ int reflect101(int i, int n) {
  i = i < 0 ? -i : i;
  return i >= n ? 2 * n - 2 - i : i;
}

__device__ __forceinline__ long long warp_sum_exact(int p) {
  unsigned lo = (unsigned)p & 0xffffu;
  int hi = p >> 16;
  unsigned slo = __reduce_add_sync(FULL, lo);
  int shi = __reduce_add_sync(FULL, hi);
  return ((long long)shi << 16) + (long long)slo;
}

__device__ __forceinline__ void bilinear_weights(float a, float b, int& w00, int& w01, int& w10, int& w11) {
  const float s = (float)(1 << WB);
  float oa = 1.f - a, ob = 1.f - b;
  w00 = __float2int_rn((oa * ob) * s);
  w01 = __float2int_rn((a * ob) * s);
  w10 = __float2int_rn((oa * b) * s);
  w11 = (1 << WB) - w00 - w01 - w10;
}

// d = c + a.h0 * b.b0 + a.h1 * b.b1 (lo) / c + a.h0 * b.b2 + a.h1 * b.b3 (hi); a: signed 16-bit halves, b: unsigned bytes
__device__ __forceinline__ int dp2a_lo(int a, unsigned b, int c) {
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int dp2a_hi(int a, unsigned b, int c) {
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// ---- Scharr derivative of one pyramid (all levels), packed (Iy << 16) | (Ix & 0xffff) per pixel ------------------
// Ix = 3 (I[y-1][x+1] - I[y-1][x-1]) + 10 (I[y][x+1] - I[y][x-1]) + 3 (I[y+1][x+1] - I[y+1][x-1])
// Iy = 3 (I[y+1][x-1] - I[y-1][x-1]) + 10 (I[y+1][x] - I[y-1][x]) + 3 (I[y+1][x+1] - I[y-1][x+1]),  BORDER_REFLECT_101.
// A warp owns a 128-column strip (lane = 4 adjacent pixels = one aligned word per row, rows are 128-byte pitched) and
// marches down SD_ROWS rows; the two outer taps come from the neighbouring lanes' words.  HBM-bound: reads 1 B,
// writes 4 B per pixel.
constexpr int SD_ROWS = 16;

__global__ void __launch_bounds__(128) scharr_kernel(const uint8_t* __restrict__ pyr, unsigned* __restrict__ deriv, LKGeom g) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int level = 0;
#pragma unroll
  for (int l = 1; l < FLV_MAX_LEVELS; ++l) level += (l < g.nlev && (int)blockIdx.x >= g.blk0[l]) ? 1 : 0;
  int w, h, pitch, strips, b0;
  unsigned long long off;
  switch (level) {
    case 0: w = g.w[0]; h = g.h[0]; pitch = g.pitch[0]; off = g.off[0]; strips = g.strips[0]; b0 = g.blk0[0]; break;
    case 1: w = g.w[1]; h = g.h[1]; pitch = g.pitch[1]; off = g.off[1]; strips = g.strips[1]; b0 = g.blk0[1]; break;
    case 2: w = g.w[2]; h = g.h[2]; pitch = g.pitch[2]; off = g.off[2]; strips = g.strips[2]; b0 = g.blk0[2]; break;
    default: w = g.w[3]; h = g.h[3]; pitch = g.pitch[3]; off = g.off[3]; strips = g.strips[3]; b0 = g.blk0[3]; break;
  }
  const int task = ((int)blockIdx.x - b0) * 4 + warp;          // (row group, strip)
  const int rg = task / strips, strip = task - rg * strips;
  const int y0 = rg * SD_ROWS;
  if (y0 >= h) return;
  const int x0 = strip * 128 + 4 * lane;
  const bool act = x0 < w;                                      // the word may extend past w inside the row pitch
  const uint8_t* img = pyr + (size_t)blockIdx.y * g.stream_stride + off;
  unsigned* dv = deriv + (size_t)blockIdx.y * g.stream_stride + off;
  const int last = w - 1 - x0;                                   // index (0..3) of the image's last column in this word, if any

  auto row = [&](int y, int (&dx)[4], int (&sx)[4]) {
    const int yr = reflect101(y, h);
    const uint8_t* r = img + (size_t)yr * pitch;
    const unsigned wd = act ? *reinterpret_cast<const unsigned*>(r + x0) : 0u;
    const unsigned wl = __shfl_up_sync(FULL, wd, 1), wr = __shfl_down_sync(FULL, wd, 1);
    int a[6];
    a[1] = wd & 0xffu; a[2] = (wd >> 8) & 0xffu; a[3] = (wd >> 16) & 0xffu; a[4] = wd >> 24;
    a[0] = wl >> 24; a[5] = wr & 0xffu;
    if (act) {
      if (lane == 0) a[0] = x0 == 0 ? a[2] : r[x0 - 1];          // column -1 -> 1
      if (lane == 31 && x0 + 4 < w) a[5] = r[x0 + 4];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) if (i == last) a[i + 2] = a[i];  // column w -> w-2
#pragma unroll
    for (int i = 0; i < 4; ++i) { dx[i] = a[i + 2] - a[i]; sx[i] = 3 * a[i] + 10 * a[i + 1] + 3 * a[i + 2]; }
  };
  int dxp[4], sxp[4], dxc[4], sxc[4];
  row(y0 - 1, dxp, sxp);
  row(y0, dxc, sxc);
  const int y1 = min(y0 + SD_ROWS, h);
#pragma unroll 2
  for (int y = y0; y < y1; ++y) {
    int dxn[4], sxn[4];
    row(y + 1, dxn, sxn);
    unsigned o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int vX = 3 * dxp[i] + 10 * dxc[i] + 3 * dxn[i];
      const int vY = sxn[i] - sxp[i];
      o[i] = ((unsigned)vY << 16) | ((unsigned)vX & 0xffffu);
      dxp[i] = dxc[i]; sxp[i] = sxc[i]; dxc[i] = dxn[i]; sxc[i] = sxn[i];
    }
    if (act) {
      unsigned* q = dv + (size_t)y * pitch + x0;
      if (last >= 3) *reinterpret_cast<uint4*>(q) = make_uint4(o[0], o[1], o[2], o[3]);
      else {
#pragma unroll
        for (int i = 0; i < 4; ++i) if (i <= last) q[i] = o[i];
      }
    }
  }
}

// ---- packed 32x32 patch: P[k] = rows 4k..4k+3 of column (ox + lane) ----------------------------------------------
__device__ __forceinline__ unsigned pack4(unsigned b0, unsigned b1, unsigned b2, unsigned b3) {
  return __byte_perm(__byte_perm(b0, b1, 0x1140), __byte_perm(b2, b3, 0x1140), 0x5410);
}

__device__ __forceinline__ void load_patch(const uint8_t* __restrict__ img, int pitch, int w, int h, int ox, int oy,
                                           int lane, unsigned (&P)[8], unsigned (&Q)[8]) {
  const bool interior = ox >= 0 && ox + 32 <= w && oy >= 0 && oy + 32 <= h;
  if (interior) {
    const uint8_t* p = img + (size_t)oy * pitch + ox + lane;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const unsigned b0 = p[0], b1 = p[pitch], b2 = p[2 * pitch], b3 = p[3 * pitch];
      P[k] = pack4(b0, b1, b2, b3);
      p += 4 * pitch;
    }
  } else {
    const uint8_t* col = img + reflect101(ox + lane, w);
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
      unsigned b[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = col[(size_t)reflect101(oy + 4 * k + j, h) * pitch];
      const unsigned v = pack4(b[0], b[1], b[2], b[3]);
      // dynamic register index avoided: select
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) if (kk == k) P[kk] = v;
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) Q[k] = __shfl_down_sync(FULL, P[k], 1);
}

// TMA variant of the interior patch load: lane 0 issues one 3-D tile load, the warp waits on its mbarrier and packs its column
// from the row-major shared-memory tile.  A tile load must start on a 16-byte boundary of the innermost dimension (any other
// start coordinate faults with "illegal instruction": tools/tma_probe.cu), so the box is 48 x 32 x 1 bytes at column
// ox & ~15 and the lane reads column (ox & 15) + lane of it.
constexpr int TMA_BOX_W = 48;
constexpr int TMA_TILE_BYTES = TMA_BOX_W * 32;
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void load_patch_tma(const CUtensorMap* tm, uint8_t* tile, unsigned long long* bar, unsigned& parity,
                                               int ox, int oy, int s, int lane, unsigned (&P)[8], unsigned (&Q)[8]) {
  __syncwarp();                                   // every lane is done with the previous tile
  if (lane == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"((unsigned)TMA_TILE_BYTES) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_addr(tile)), "l"(tm), "r"(ox & ~15), "r"(oy), "r"(s), "r"(smem_addr(bar)) : "memory");
  }
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LKW_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LKD_%=;\n"
      "bra LKW_%=;\n"
      "LKD_%=:\n"
      "}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
  parity ^= 1u;
  const uint8_t* p = tile + (ox & 15) + lane;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const unsigned b0 = p[0], b1 = p[TMA_BOX_W], b2 = p[2 * TMA_BOX_W], b3 = p[3 * TMA_BOX_W];
    P[k] = pack4(b0, b1, b2, b3);
    p += 4 * TMA_BOX_W;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) Q[k] = __shfl_down_sync(FULL, P[k], 1);
}

// bilinear blend of rows (y, y+1) x columns (lane, lane+1), y = 4k + j:  c + w00 J[y][x] + w01 J[y][x+1] + w10 J[y+1][x] + w11 J[y+1][x+1]
// wv0 = w00 | w10 << 16 (own column), wv1 = w01 | w11 << 16 (right column)
template <int J>
__device__ __forceinline__ int blend_row(unsigned Pk, unsigned Pn, unsigned Qk, unsigned Qn, int wv0, int wv1, int c) {
  if (J == 0) return dp2a_lo(wv1, Qk, dp2a_lo(wv0, Pk, c));
  if (J == 1) return dp2a_lo(wv1, Qk >> 8, dp2a_lo(wv0, Pk >> 8, c));
  if (J == 2) return dp2a_hi(wv1, Qk, dp2a_hi(wv0, Pk, c));
  return dp2a_lo(wv1, __funnelshift_r(Qk, Qn, 24), dp2a_lo(wv0, __funnelshift_r(Pk, Pn, 24), c));
}

// One pass over the window: diff = (blend + tmpl.x) >> 9 with tmpl.x = (1<<8) - (Ival<<9), tmpl.y = Iy<<16 | Ix&0xffff.
//   ERR = false: o1 += diff*Ix, o2 += diff*Iy;   ERR = true: o1 += |diff| (live lanes only).
template <bool ERR>
__device__ __forceinline__ void window_pass(const unsigned (&P)[8], const unsigned (&Q)[8], int wv0, int wv1,
                                            const int2* __restrict__ tmpl, bool live, int& o1, int& o2) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const unsigned Pn = k < 7 ? P[k + 1] : 0u, Qn = k < 7 ? Q[k + 1] : 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = 4 * k + j;
      if (r >= WIN) continue;
      const int2 t = tmpl[r * 32];
      int s;
      if (j == 0) s = blend_row<0>(P[k], Pn, Q[k], Qn, wv0, wv1, t.x);
      else if (j == 1) s = blend_row<1>(P[k], Pn, Q[k], Qn, wv0, wv1, t.x);
      else if (j == 2) s = blend_row<2>(P[k], Pn, Q[k], Qn, wv0, wv1, t.x);
      else s = blend_row<3>(P[k], Pn, Q[k], Qn, wv0, wv1, t.x);
      const int diff = s >> (WB - 5);
      if (ERR) {
        const int d = live ? diff : 0;
        o1 += d < 0 ? -d : d;
      } else {
        o1 += diff * (int)(short)(t.y & 0xffff);
        o2 += diff * (t.y >> 16);
      }
    }
  }
}

// Template of one point and level -> shared memory: tmpl[r*32] = ((1<<8) - (Ival<<9), Iy<<16 | Ix&0xffff) for window
// rows r = 0..30 of this lane's column, plus the lane's share of the gradient matrix sums.  Ival / Ix / Iy are the
// bilinear blends of the u8 level (BORDER_REFLECT_101) and of the Scharr pyramid (zero outside the image).
// Rolled over groups of four window rows (small code: the instruction cache, not the ALUs, limited the unrolled form):
// a group loads its 4 new image rows + derivative rows up front, packs the u8 column, blends.
// INTERIOR: the whole 32x32 footprint is inside the image (no reflection, no range tests, pointer walk).
template <bool INTERIOR>
__device__ __forceinline__ void build_template(const uint8_t* __restrict__ I, const unsigned* __restrict__ Dv, int pitch,
                                               int w, int h, int ipx, int ipy, int lane, bool live, int w00, int w01,
                                               int w10, int w11, int2* __restrict__ tmpl, int& a11, int& a12, int& a22) {
  const int wv0 = (w00 & 0xffff) | (w10 << 16), wv1 = (w01 & 0xffff) | (w11 << 16);
  const int x = ipx + lane;
  const bool in_x = INTERIOR || (x >= 0 && x < w);
  const uint8_t* icol = INTERIOR ? I + (size_t)ipy * pitch + x : I + reflect101(x, w);
  const unsigned* dcol = INTERIOR ? Dv + (size_t)ipy * pitch + x : Dv + x;
  unsigned bI0, d0;
  if (INTERIOR) { bI0 = icol[0]; d0 = dcol[0]; }
  else {
    bI0 = icol[(size_t)reflect101(ipy, h) * pitch];
    d0 = (in_x && ipy >= 0 && ipy < h) ? dcol[(size_t)ipy * pitch] : 0u;
  }
  unsigned d0r = __shfl_down_sync(FULL, d0, 1);
  int qX = (int)(short)(d0 & 0xffffu), qY = (int)d0 >> 16, qXr = (int)(short)(d0r & 0xffffu), qYr = (int)d0r >> 16;
#pragma unroll 1
  for (int k = 0; k < 8; ++k) {
    unsigned bI[5], dd[4];
    bI[0] = bI0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (INTERIOR) {
        const bool used = 4 * k + j + 1 <= WIN;      // image row 32 of the footprint is never blended: do not touch it
        bI[j + 1] = used ? icol[(size_t)(j + 1) * pitch] : 0u;
        dd[j] = used ? dcol[(size_t)(j + 1) * pitch] : 0u;
      } else {
        const int y1 = ipy + 4 * k + j + 1;
        bI[j + 1] = icol[(size_t)reflect101(y1, h) * pitch];
        dd[j] = (in_x && y1 >= 0 && y1 < h) ? dcol[(size_t)y1 * pitch] : 0u;
      }
    }
    if (INTERIOR) { icol += 4 * (size_t)pitch; dcol += 4 * (size_t)pitch; }
    bI0 = bI[4];
    const unsigned Pk = pack4(bI[0], bI[1], bI[2], bI[3]), Pn = bI[4];
    const unsigned Qk = __shfl_down_sync(FULL, Pk, 1), Qn = __shfl_down_sync(FULL, Pn, 1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = 4 * k + j;                                  // window row r blends image rows r, r+1
      const unsigned d1 = dd[j];
      const unsigned d1r = __shfl_down_sync(FULL, d1, 1);
      const int vX = (int)(short)(d1 & 0xffffu), vY = (int)d1 >> 16, vXr = (int)(short)(d1r & 0xffffu), vYr = (int)d1r >> 16;
      int sI;
      if (j == 0) sI = blend_row<0>(Pk, Pn, Qk, Qn, wv0, wv1, 1 << (WB - 5 - 1));
      else if (j == 1) sI = blend_row<1>(Pk, Pn, Qk, Qn, wv0, wv1, 1 << (WB - 5 - 1));
      else if (j == 2) sI = blend_row<2>(Pk, Pn, Qk, Qn, wv0, wv1, 1 << (WB - 5 - 1));
      else sI = blend_row<3>(Pk, Pn, Qk, Qn, wv0, wv1, 1 << (WB - 5 - 1));
      const int iv = sI >> (WB - 5);                             // lane 31's value is never used (Ix = Iy = 0, ERR masks it)
      int ix = (qX * w00 + qXr * w01 + vX * w10 + vXr * w11 + (1 << (WB - 1))) >> WB;
      int iy = (qY * w00 + qYr * w01 + vY * w10 + vYr * w11 + (1 << (WB - 1))) >> WB;
      ix = live ? ix : 0; iy = live ? iy : 0;
      if (r < WIN) {
        tmpl[r * 32] = make_int2((1 << (WB - 5 - 1)) - (iv << (WB - 5)), (iy << 16) | (ix & 0xffff));
        a11 += ix * ix; a12 += ix * iy; a22 += iy * iy;
      }
      qX = vX; qY = vY; qXr = vXr; qYr = vYr;
    }
  }
}

struct LKTmaps { CUtensorMap lv[FLV_MAX_LEVELS]; };      // second image, one map per level (TMA variant only)

template <bool TMA>
__global__ void __launch_bounds__(V4_WARPS * 32, 5)
lk_track_kernel_v4(const uint8_t* __restrict__ pyrI, const unsigned* __restrict__ derivI, const uint8_t* __restrict__ pyrJ,
                   LKGeom g, const int* __restrict__ npts, const float* __restrict__ prev_xy,
                   const float* __restrict__ init_xy, float* __restrict__ next_xy, uint8_t* __restrict__ status,
                   float* __restrict__ err, int max_pts, int nlev_used, int max_iter, double eps2, double min_eig_thr,
                   float err_scale, const LKTmaps* __restrict__ tmaps_g) {
  extern __shared__ __align__(128) int2 tmpl_all[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int s = blockIdx.y;
  const int pt = blockIdx.x * V4_WARPS + warp;
  if (pt >= npts[s]) return;
  int2* tmpl = tmpl_all + warp * TMPL_WORDS + lane;      // this lane's column of the template: tmpl[row*32]
  // TMA variant: per-warp 1 KB tile + mbarrier behind the templates
  uint8_t* tile = reinterpret_cast<uint8_t*>(tmpl_all + V4_WARPS * TMPL_WORDS) + warp * TMA_TILE_BYTES;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(tmpl_all + V4_WARPS * TMPL_WORDS) + V4_WARPS * TMA_TILE_BYTES) + warp;
  unsigned parity = 0;
  if (TMA) {
    if (lane == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
  }
  const size_t pidx = (size_t)s * max_pts + pt;
  const uint8_t* Ibase = pyrI + (size_t)s * g.stream_stride;
  const unsigned* Dbase = derivI + (size_t)s * g.stream_stride;
  const uint8_t* Jbase = pyrJ + (size_t)s * g.stream_stride;
  const float px0 = prev_xy[2 * pidx], py0 = prev_xy[2 * pidx + 1];
  float nx = init_xy[2 * pidx], ny = init_xy[2 * pidx + 1];
  int st = 1;
  float er = 0.f;
  const float half = (float)((WIN - 1) * 0.5);
  const float FLT_SCALE = 1.f / (float)(1 << 20);
  const bool live = lane < WIN;

#pragma unroll 1
  for (int level = nlev_used - 1; level >= 0; --level) {
    int w, h, pitch;
    unsigned long long off;
    const CUtensorMap* tm;
    switch (level) {
      case 0: w = g.w[0]; h = g.h[0]; pitch = g.pitch[0]; off = g.off[0]; break;
      case 1: w = g.w[1]; h = g.h[1]; pitch = g.pitch[1]; off = g.off[1]; break;
      case 2: w = g.w[2]; h = g.h[2]; pitch = g.pitch[2]; off = g.off[2]; break;
      default: w = g.w[3]; h = g.h[3]; pitch = g.pitch[3]; off = g.off[3]; break;
    }
    tm = TMA ? &tmaps_g->lv[level] : nullptr;               // descriptors live in global memory (written once per context)
    const uint8_t* I = Ibase + off;
    const unsigned* Dv = Dbase + off;
    const uint8_t* J = Jbase + off;
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

    unsigned P[8], Q[8];
    // ---- template patch -> shared memory ------------------------------------------------------
    int a11 = 0, a12 = 0, a22 = 0;
    if (ipx >= 0 && ipx + 32 <= w && ipy >= 0 && ipy + 32 <= h)
      build_template<true>(I, Dv, pitch, w, h, ipx, ipy, lane, live, w00, w01, w10, w11, tmpl, a11, a12, a22);
    else
      build_template<false>(I, Dv, pitch, w, h, ipx, ipy, lane, live, w00, w01, w10, w11, tmpl, a11, a12, a22);
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
    int cox = 0x7fffffff, coy = 0;                 // origin of the cached J patch (none yet)
#pragma unroll 1
    for (int j = 0; j < max_iter; ++j) {
      const int inx = (int)floorf(cx), iny = (int)floorf(cy);
      if (inx < -WIN || inx >= w || iny < -WIN || iny >= h) {
        if (level == 0) st = 0;
        break;
      }
      bilinear_weights(cx - (float)inx, cy - (float)iny, w00, w01, w10, w11);
      if (inx != cox || iny != coy) {
        if (TMA && inx >= 0 && inx + 32 <= w && iny >= 0 && iny + 32 <= h) load_patch_tma(tm, tile, bar, parity, inx, iny, s, lane, P, Q);
        else load_patch(J, pitch, w, h, inx, iny, lane, P, Q);
        cox = inx; coy = iny;
      }
      int b1 = 0, b2 = 0;
      window_pass<false>(P, Q, (w00 & 0xffff) | (w10 << 16), (w01 & 0xffff) | (w11 << 16), tmpl, live, b1, b2);
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
    if (level == 0 && st && err != nullptr) {            // (the reference never reads `err`: callers may pass NULL and save the pass)
      const float qx = nx - half, qy = ny - half;
      const int inx = (int)floorf(qx), iny = (int)floorf(qy);
      if (inx < -WIN || inx >= w || iny < -WIN || iny >= h) {
        st = 0;
      } else {
        bilinear_weights(qx - (float)inx, qy - (float)iny, w00, w01, w10, w11);
        if (inx != cox || iny != coy) {
          if (TMA && inx >= 0 && inx + 32 <= w && iny >= 0 && iny + 32 <= h) load_patch_tma(tm, tile, bar, parity, inx, iny, s, lane, P, Q);
          else load_patch(J, pitch, w, h, inx, iny, lane, P, Q);
          cox = inx; coy = iny;
        }
        int e = 0, unused = 0;
        window_pass<true>(P, Q, (w00 & 0xffff) | (w10 << 16), (w01 & 0xffff) | (w11 << 16), tmpl, live, e, unused);
        er = __ll2float_rn(warp_sum_exact(e)) * err_scale;
      }
    }
  }
  if (lane == 0) {
    next_xy[2 * pidx] = nx;
    next_xy[2 * pidx + 1] = ny;
    status[pidx] = (uint8_t)st;
    if (err != nullptr) err[pidx] = er;
  }
}

}  // namespace

// tensor maps of the second image's pyramid levels: u8 [S][h][w] with the level's row pitch / the slot's stream stride
static int lk_make_tmaps(flv_ctx* ctx, int slot, LKTmaps& out) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    FLV_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    if (!fn || qr != cudaDriverEntryPointSuccess) FLV_FAIL(ctx, FLV_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    encode = (EncodeFn)fn;
  }
  memset(&out, 0, sizeof(out));
  for (int l = 0; l < ctx->geom.nlev; ++l) {
    const cuuint64_t dims[3] = {(cuuint64_t)ctx->geom.lv[l].w, (cuuint64_t)ctx->geom.lv[l].h, (cuuint64_t)ctx->S};
    const cuuint64_t strides[2] = {(cuuint64_t)ctx->geom.lv[l].pitch, (cuuint64_t)ctx->geom.stream_stride};
    const cuuint32_t box[3] = {TMA_BOX_W, 32, 1}, estr[3] = {1, 1, 1};
    const CUresult r = encode(&out.lv[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ctx->pyr[slot] + ctx->geom.lv[l].off, dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) FLV_FAIL(ctx, FLV_ERR_CUDA, "cuTensorMapEncodeTiled(level %d) failed: %d", l, (int)r);
  }
  return FLV_OK;
}

int flv_launch_lk_v4(flv_ctx* ctx, int src_slot, int dst_slot, int n_streams, const int* d_npts,
                     const float* d_prev, const float* d_init, float* d_next, uint8_t* d_status,
                     float* d_err, int nlev_used, int max_iter, double eps2, double min_eig_thr) {
  LKGeom g;
  g.nlev = ctx->geom.nlev;
  int nb = 0;
  for (int l = 0; l < FLV_MAX_LEVELS; ++l) {
    g.w[l] = ctx->geom.lv[l].w; g.h[l] = ctx->geom.lv[l].h; g.pitch[l] = ctx->geom.lv[l].pitch;
    g.off[l] = ctx->geom.lv[l].off;
    g.strips[l] = (g.w[l] + 127) / 128;
    g.blk0[l] = nb;
    if (l < g.nlev) nb += (g.strips[l] * ((g.h[l] + SD_ROWS - 1) / SD_ROWS) + 3) / 4;
  }
  g.blk0[FLV_MAX_LEVELS] = nb;
  g.stream_stride = ctx->geom.stream_stride;
  // derivative pyramid of the template slot: built on first use after the slot's images changed
  if (!ctx->deriv[src_slot]) {
    FLV_CUDA(ctx, cudaMalloc(&ctx->deriv[src_slot], (size_t)ctx->S * ctx->geom.stream_stride * sizeof(unsigned)));
  }
  if (ctx->deriv_streams[src_slot] < n_streams) {
    dim3 grid(nb, n_streams);
    scharr_kernel<<<grid, 128, 0, ctx->stream>>>(ctx->pyr[src_slot], ctx->deriv[src_slot], g);
    ctx->launches++;
    FLV_CUDA(ctx, cudaGetLastError());
    ctx->deriv_streams[src_slot] = n_streams;
  }
  const bool tma = !(getenv("FLV_LK_VARIANT") && atoi(getenv("FLV_LK_VARIANT")) == 6);      // default: TMA patch staging
  const size_t smem = (size_t)V4_WARPS * TMPL_WORDS * sizeof(int2) + (tma ? V4_WARPS * (TMA_TILE_BYTES + 8) : 0);
  if (!ctx->attr_lk4) {          // per context (= per device): function attributes do not carry across devices
    FLV_CUDA(ctx, cudaFuncSetAttribute(lk_track_kernel_v4<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FLV_CUDA(ctx, cudaFuncSetAttribute(lk_track_kernel_v4<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       getenv("FLV_LK_CARVEOUT") ? atoi(getenv("FLV_LK_CARVEOUT")) : 75));
    FLV_CUDA(ctx, cudaFuncSetAttribute(lk_track_kernel_v4<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)((size_t)V4_WARPS * TMPL_WORDS * sizeof(int2) + V4_WARPS * (TMA_TILE_BYTES + 8))));
    FLV_CUDA(ctx, cudaFuncSetAttribute(lk_track_kernel_v4<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       getenv("FLV_LK_CARVEOUT") ? atoi(getenv("FLV_LK_CARVEOUT")) : 75));
    ctx->attr_lk4 = 1;
  }
  dim3 grid((ctx->max_pts + V4_WARPS - 1) / V4_WARPS, n_streams);
  const float err_scale = (float)(1.0 / (32 * WIN * WIN));
  if (tma) {
    if (!ctx->d_lk_tmaps) {      // one descriptor set per image slot, encoded once (the pyramids never move)
      FLV_CUDA(ctx, cudaMalloc(&ctx->d_lk_tmaps, FLV_NUM_SLOTS * sizeof(LKTmaps)));
      for (int sl = 0; sl < FLV_NUM_SLOTS; ++sl) {
        LKTmaps maps;
        int rc = lk_make_tmaps(ctx, sl, maps);
        if (rc) return rc;
        FLV_CUDA(ctx, cudaMemcpy((LKTmaps*)ctx->d_lk_tmaps + sl, &maps, sizeof(maps), cudaMemcpyHostToDevice));
      }
    }
    lk_track_kernel_v4<true><<<grid, V4_WARPS * 32, smem, ctx->stream>>>(
        ctx->pyr[src_slot], ctx->deriv[src_slot], ctx->pyr[dst_slot], g, d_npts, d_prev, d_init, d_next, d_status, d_err,
        ctx->max_pts, nlev_used, max_iter, eps2, min_eig_thr, err_scale, (const LKTmaps*)ctx->d_lk_tmaps + dst_slot);
  } else
  lk_track_kernel_v4<false><<<grid, V4_WARPS * 32, smem, ctx->stream>>>(
      ctx->pyr[src_slot], ctx->deriv[src_slot], ctx->pyr[dst_slot], g, d_npts, d_prev, d_init, d_next, d_status, d_err,
      ctx->max_pts, nlev_used, max_iter, eps2, min_eig_thr, err_scale, nullptr);
  ctx->launches++;
  FLV_CUDA(ctx, cudaGetLastError());
  return FLV_OK;
}
