/* C twin of oracle/lk_ref.py -- TEST INFRASTRUCTURE (oracle), never linked into the product.
 *
 * Same restatement of OpenCV's calcOpticalFlowPyrLK as FLVIS calls it
 * (/root/reference/src/processing/lkorb_tracking.cpp:64-73, src/processing/camera_frame.cpp:124-128; algorithm:
 * SURVEY.md Appendix A.1), statement for statement: exact integer window sums, every f32 operation individually
 * rounded (build with -ffp-contract=off, no -ffast-math).  It exists so that long sequences (hundreds of frames) can be
 * checked LIVE on the GPU box in seconds; tests/test_oracle_cpu.py pins it bit-for-bit to the Python restatement, which
 * is pinned to cv2 4.13.0 by tests/golden/lk_*.npz.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define WB 14
#define MAXLEV 12

static inline int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

static void pyr_down(const uint8_t* src, int w, int h, uint8_t* dst, int ow, int oh) {
  static const int k[5] = {1, 4, 6, 4, 1};
  int* rows = (int*)malloc(sizeof(int) * (size_t)h * ow);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < ow; ++x) {
      int s = 0;
      for (int t = -2; t <= 2; ++t) s += k[t + 2] * src[(size_t)y * w + reflect101(2 * x + t, w)];
      rows[(size_t)y * ow + x] = s;
    }
  for (int y = 0; y < oh; ++y)
    for (int x = 0; x < ow; ++x) {
      int s = 0;
      for (int t = -2; t <= 2; ++t) s += k[t + 2] * rows[(size_t)reflect101(2 * y + t, h) * ow + x];
      dst[(size_t)y * ow + x] = (uint8_t)((s + 128) >> 8);
    }
  free(rows);
}

typedef struct {
  int w, h, B, pw, ph;    /* padded arrays are (h+2B) x (w+2B) */
  int32_t *I, *J, *DX, *DY;
} level_t;

static void make_level(level_t* L, const uint8_t* I, const uint8_t* J, int w, int h, int win) {
  const int B = win + 1;
  L->w = w; L->h = h; L->B = B; L->pw = w + 2 * B; L->ph = h + 2 * B;
  const size_t n = (size_t)L->pw * L->ph;
  L->I = (int32_t*)malloc(n * 4); L->J = (int32_t*)malloc(n * 4);
  L->DX = (int32_t*)calloc(n, 4); L->DY = (int32_t*)calloc(n, 4);
  for (int y = 0; y < L->ph; ++y) {
    const int sy = reflect101(y - B, h);
    for (int x = 0; x < L->pw; ++x) {
      const int sx = reflect101(x - B, w);
      L->I[(size_t)y * L->pw + x] = I[(size_t)sy * w + sx];
      L->J[(size_t)y * L->pw + x] = J[(size_t)sy * w + sx];
    }
  }
  /* calcSharrDeriv with REFLECT_101 at the image edge; zero outside the image (np.pad constant) */
  for (int y = 0; y < h; ++y) {
    const int ym = reflect101(y - 1, h), yp = reflect101(y + 1, h);
    for (int x = 0; x < w; ++x) {
      const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
#define PX(yy, xx) ((int)I[(size_t)(yy) * w + (xx)])
      const int t0p = (PX(ym, xp) + PX(yp, xp)) * 3 + PX(y, xp) * 10;
      const int t0m = (PX(ym, xm) + PX(yp, xm)) * 3 + PX(y, xm) * 10;
      const int t1p = PX(yp, xp) - PX(ym, xp), t1m = PX(yp, xm) - PX(ym, xm), t1c = PX(yp, x) - PX(ym, x);
#undef PX
      L->DX[(size_t)(y + B) * L->pw + x + B] = t0p - t0m;
      L->DY[(size_t)(y + B) * L->pw + x + B] = (t1p + t1m) * 3 + t1c * 10;
    }
  }
}
static void free_level(level_t* L) { free(L->I); free(L->J); free(L->DX); free(L->DY); }

static void weights(float a, float b, int* w4) {
  const float one = 1.f, s = (float)(1 << WB);
  float t;
  t = (one - a) * (one - b); w4[0] = (int)lrintf(t * s);
  t = a * (one - b);         w4[1] = (int)lrintf(t * s);
  t = (one - a) * b;         w4[2] = (int)lrintf(t * s);
  w4[3] = (1 << WB) - w4[0] - w4[1] - w4[2];
}

/* win x win bilinear blend of the (win+1)^2 window at (x0, y0) of padded array A, descaled by n bits */
static void interp(const int32_t* A, int pw, int x0, int y0, int win, const int* w4, int n, int32_t* out) {
  const int rnd = 1 << (n - 1);
  for (int y = 0; y < win; ++y) {
    const int32_t* r0 = A + (size_t)(y0 + y) * pw + x0;
    const int32_t* r1 = r0 + pw;
    for (int x = 0; x < win; ++x)
      out[y * win + x] = (r0[x] * w4[0] + r0[x + 1] * w4[1] + r1[x] * w4[2] + r1[x + 1] * w4[3] + rnd) >> n;
  }
}

/* prev/next: h x w u8 images (tightly packed); prev_pts/init_pts/next_pts: n x 2 f32; status n u8; err n f32.
 * Returns the number of pyramid levels used. */
int lk_ref_track(const uint8_t* prev, const uint8_t* next, int w, int h, int n, const float* prev_pts, const float* init_pts,
                 float* next_pts, uint8_t* status, float* errs, int win, int max_level, int max_iter, double eps, double min_eig_thr) {
  if (max_iter < 0) max_iter = 0;
  if (max_iter > 100) max_iter = 100;
  if (eps < 0) eps = 0;
  if (eps > 10) eps = 10;
  const double eps2 = eps * eps;
  /* pyramids: stop before a level with w <= win or h <= win */
  uint8_t *pI[MAXLEV], *pJ[MAXLEV];
  int lw[MAXLEV], lh[MAXLEV], nlev = 1;
  pI[0] = (uint8_t*)prev; pJ[0] = (uint8_t*)next; lw[0] = w; lh[0] = h;
  while (nlev - 1 < max_level && nlev < MAXLEV) {
    const int ow = (lw[nlev - 1] + 1) / 2, oh = (lh[nlev - 1] + 1) / 2;
    if (ow <= win || oh <= win) break;
    pI[nlev] = (uint8_t*)malloc((size_t)ow * oh); pJ[nlev] = (uint8_t*)malloc((size_t)ow * oh);
    pyr_down(pI[nlev - 1], lw[nlev - 1], lh[nlev - 1], pI[nlev], ow, oh);
    pyr_down(pJ[nlev - 1], lw[nlev - 1], lh[nlev - 1], pJ[nlev], ow, oh);
    lw[nlev] = ow; lh[nlev] = oh; nlev++;
  }
  level_t L[MAXLEV];
  for (int l = 0; l < nlev; ++l) make_level(&L[l], pI[l], pJ[l], lw[l], lh[l], win);
  for (int i = 0; i < n; ++i) { status[i] = 1; errs[i] = 0.f; next_pts[2 * i] = init_pts[2 * i]; next_pts[2 * i + 1] = init_pts[2 * i + 1]; }
  const float half = (float)((win - 1) * 0.5);
  const float FLT_SCALE = 1.f / (float)(1 << 20);
  const float FLT_EPS = 1.1920928955078125e-07f;
  const float err_scale = (float)(1.0 / (32 * win * win));
  const int ww = win * win;
  int32_t* Iw = (int32_t*)malloc(sizeof(int32_t) * ww * 4);
  int32_t *Ix = Iw + ww, *Iy = Ix + ww, *Jw = Iy + ww;
  for (int level = nlev - 1; level >= 0; --level) {
    const level_t* V = &L[level];
    const int lwid = V->w, lhei = V->h, B = V->B, pw = V->pw;
    const float sc = (float)(1.0 / (double)(1 << level));
    for (int i = 0; i < n; ++i) {
      float px = prev_pts[2 * i] * sc, py = prev_pts[2 * i + 1] * sc;
      float nx, ny;
      if (level == nlev - 1) { nx = next_pts[2 * i] * sc; ny = next_pts[2 * i + 1] * sc; }
      else { nx = next_pts[2 * i] * 2.f; ny = next_pts[2 * i + 1] * 2.f; }
      next_pts[2 * i] = nx; next_pts[2 * i + 1] = ny;
      px = px - half; py = py - half;
      const int ipx = (int)floorf(px), ipy = (int)floorf(py);
      if (ipx < -win || ipx >= lwid || ipy < -win || ipy >= lhei) {
        if (level == 0) { status[i] = 0; errs[i] = 0.f; }
        continue;
      }
      float a = px - (float)ipx, b = py - (float)ipy;
      int w4[4];
      weights(a, b, w4);
      interp(V->I, pw, ipx + B, ipy + B, win, w4, WB - 5, Iw);
      interp(V->DX, pw, ipx + B, ipy + B, win, w4, WB, Ix);
      interp(V->DY, pw, ipx + B, ipy + B, win, w4, WB, Iy);
      int64_t s11 = 0, s12 = 0, s22 = 0;
      for (int k = 0; k < ww; ++k) { s11 += (int64_t)Ix[k] * Ix[k]; s12 += (int64_t)Ix[k] * Iy[k]; s22 += (int64_t)Iy[k] * Iy[k]; }
      const float A11 = (float)(double)s11 * FLT_SCALE, A12 = (float)(double)s12 * FLT_SCALE, A22 = (float)(double)s22 * FLT_SCALE;
      float t1 = A11 * A22, t2 = A12 * A12;
      float D = t1 - t2;
      const float dA = A11 - A22;
      t1 = dA * dA; t2 = 4.f * A12; t2 = t2 * A12;
      const float disc = t1 + t2;
      t1 = A22 + A11; t2 = sqrtf(disc); t1 = t1 - t2;
      const float min_eig = t1 / (float)(2 * win * win);
      if ((double)min_eig < min_eig_thr || (double)D < (double)FLT_EPS) {
        if (level == 0) status[i] = 0;
        continue;
      }
      D = 1.f / D;
      nx = nx - half; ny = ny - half;
      float pdx = 0.f, pdy = 0.f;
      for (int j = 0; j < max_iter; ++j) {
        const int inx = (int)floorf(nx), iny = (int)floorf(ny);
        if (inx < -win || inx >= lwid || iny < -win || iny >= lhei) {
          if (level == 0) status[i] = 0;
          break;
        }
        a = nx - (float)inx; b = ny - (float)iny;
        weights(a, b, w4);
        interp(V->J, pw, inx + B, iny + B, win, w4, WB - 5, Jw);
        int64_t sb1 = 0, sb2 = 0;
        for (int k = 0; k < ww; ++k) { const int d = Jw[k] - Iw[k]; sb1 += (int64_t)d * Ix[k]; sb2 += (int64_t)d * Iy[k]; }
        const float b1 = (float)(double)sb1 * FLT_SCALE, b2 = (float)(double)sb2 * FLT_SCALE;
        float u1 = A12 * b2, u2 = A22 * b1;
        float dx = u1 - u2; dx = dx * D;
        u1 = A12 * b1; u2 = A11 * b2;
        float dy = u1 - u2; dy = dy * D;
        nx = nx + dx; ny = ny + dy;
        next_pts[2 * i] = nx + half; next_pts[2 * i + 1] = ny + half;
        if ((double)dx * (double)dx + (double)dy * (double)dy <= eps2) break;
        if (j > 0) {
          const float sx = dx + pdx, sy = dy + pdy;
          if (fabs((double)sx) < 0.01 && fabs((double)sy) < 0.01) {
            next_pts[2 * i] = next_pts[2 * i] - dx * 0.5f;
            next_pts[2 * i + 1] = next_pts[2 * i + 1] - dy * 0.5f;
            break;
          }
        }
        pdx = dx; pdy = dy;
      }
      if (level == 0 && status[i]) {
        const float qx = next_pts[2 * i] - half, qy = next_pts[2 * i + 1] - half;
        const int inx = (int)floorf(qx), iny = (int)floorf(qy);
        if (inx < -win || inx >= lwid || iny < -win || iny >= lhei) { status[i] = 0; continue; }
        a = qx - (float)inx; b = qy - (float)iny;
        weights(a, b, w4);
        interp(V->J, pw, inx + B, iny + B, win, w4, WB - 5, Jw);
        int64_t sa = 0;
        for (int k = 0; k < ww; ++k) { const int d = Jw[k] - Iw[k]; sa += d < 0 ? -d : d; }
        errs[i] = (float)(double)sa * err_scale;
      }
    }
  }
  free(Iw);
  for (int l = 0; l < nlev; ++l) free_level(&L[l]);
  for (int l = 1; l < nlev; ++l) { free(pI[l]); free(pJ[l]); }
  return nlev;
}
