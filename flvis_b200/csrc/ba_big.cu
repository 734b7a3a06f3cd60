// K7-K10 for LARGE windows (26 .. 100 poses; the reference accepts window sizes up to 100, src/backend/vo_localmap.cpp:441-447).
// ba.cu keeps a window's reduced camera system in one SM's shared memory (<= 24 free poses); here the system (n = 6 * free poses
// <= 594, 2.8 MB) and every other work array live in global memory / L2, and ONE THREAD-BLOCK CLUSTER (8 CTAs x 256 threads) per
// window runs the whole Levenberg-Marquardt loop as a sequence of data-parallel PHASES separated by cluster barriers.  Same
// algorithm and reference lines as ba.cu (see its header): robust chi2, analytic Jacobians, Huber, Schur complement on the
// landmarks, LDL^T of the reduced system, back-substitution, g2o's LM control with <= 10 trials, chi2 > 3 cull between the two
// optimize() calls.  Every sum has a fixed order (no floating-point atomics): results are run-to-run deterministic.
//
// A phase is a function phase(ctx, arguments, t, T) executed by "thread" t of T; the driver (big_run) strings the phases together
// and takes every control decision from values all threads read identically after a barrier.  The SAME driver and phases run
//   * on the GPU: t = cluster-wide thread index, barrier = cluster.sync();
//   * on the host, sequentially (for t in 0..T): flv_ba_big_emulate_host, a TEST aid that lets the CPU suite check the
//     arithmetic against the oracle without a GPU.  The product never calls it.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "ctx.h"
#include "ba_math.h"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

constexpr int BIG_MAX_POSES = 100;
constexpr int BIG_CTAS = 8, BIG_THREADS = 256, BIG_T = BIG_CTAS * BIG_THREADS;
constexpr int BIG_PARTS = 16;                   // partial sums per pose in the pose pass

struct BigCtx {
  // problem (global memory)
  int P, L, E, fixed, fix_landmarks;
  Cam cam;
  double delta, d2, cull_chi2;
  double* poses; double* lms;
  const int* ep; const int* el; const double* uv; uint8_t* act;
  // workspace: doubles
  double *pbk, *lbk, *Hll, *bl, *Dinv, *W, *part, *Hd, *x, *y, *S, *invd, *red, *red2;
  // workspace: ints
  int *lstart, *pstart, *cursor, *csr_e, *csr_p, *csr_l, *pcsr, *pidx, *pose_of, *ctl;   // ctl: [0] np, [1] nact, [2] fail
  int ME, ML;
};

#define BIG_HD __host__ __device__

// ---- layout ---------------------------------------------------------------------------------------------------------------
struct BigLayout { size_t pbk, lbk, Hll, bl, Dinv, W, part, Hd, x, y, S, invd, red, red2, n_doubles;
                   size_t lstart, pstart, cursor, csr_e, csr_p, csr_l, pcsr, pidx, pose_of, ctl, n_ints; };
BIG_HD inline BigLayout big_layout(int MP, int ML, int ME) {
  BigLayout o; size_t d = 0, i = 0;
  const size_t P = MP, L = ML, E = ME, n = 6 * (size_t)MP;
  o.pbk = d; d += 7 * P; o.lbk = d; d += 3 * L; o.Hll = d; d += 6 * L; o.bl = d; d += 3 * L; o.Dinv = d; d += 6 * L;
  o.W = d; d += 18 * E; o.part = d; d += 27 * P * BIG_PARTS; o.Hd = d; d += 27 * P; o.x = d; d += n; o.y = d; d += n;
  o.S = d; d += n * n; o.invd = d; d += n; o.red = d; d += 2 * BIG_T; o.red2 = d; d += 2 * (BIG_T / 32 + 1);
  d += d & 1; o.n_doubles = d;
  o.lstart = i; i += L + 2; o.pstart = i; i += P + 2; o.cursor = i; i += L + 2; o.csr_e = i; i += E; o.csr_p = i; i += E; o.csr_l = i; i += E;
  o.pcsr = i; i += E; o.pidx = i; i += P; o.pose_of = i; i += P; o.ctl = i; i += 8; o.n_ints = i;
  return o;
}
BIG_HD inline void big_bind(BigCtx& c, unsigned char* ws, int MP, int ML, int ME) {
  const BigLayout lo = big_layout(MP, ML, ME);
  double* d = (double*)ws; int* ib = (int*)(d + lo.n_doubles);
  c.pbk = d + lo.pbk; c.lbk = d + lo.lbk; c.Hll = d + lo.Hll; c.bl = d + lo.bl; c.Dinv = d + lo.Dinv; c.W = d + lo.W; c.part = d + lo.part;
  c.Hd = d + lo.Hd; c.x = d + lo.x; c.y = d + lo.y; c.S = d + lo.S; c.invd = d + lo.invd; c.red = d + lo.red; c.red2 = d + lo.red2;
  c.lstart = ib + lo.lstart; c.pstart = ib + lo.pstart; c.cursor = ib + lo.cursor; c.csr_e = ib + lo.csr_e; c.csr_p = ib + lo.csr_p;
  c.csr_l = ib + lo.csr_l; c.pcsr = ib + lo.pcsr; c.pidx = ib + lo.pidx; c.pose_of = ib + lo.pose_of; c.ctl = ib + lo.ctl;
  c.ME = ME; c.ML = ML;
}
size_t big_ws_bytes(int MP, int ML, int ME) {
  const BigLayout lo = big_layout(MP, ML, ME);
  return (lo.n_doubles * 8 + lo.n_ints * 4 + 255) & ~(size_t)255;
}

BIG_HD inline int big_atomic_inc(int* p) {
#ifdef __CUDA_ARCH__
  return atomicAdd(p, 1);
#else
  return (*p)++;
#endif
}
BIG_HD inline int sym21b(int i, int j) { return i * 6 - (i * (i - 1)) / 2 + (j - i); }

// ---- phases -----------------------------------------------------------------------------------------------------------------
enum Phase { PH_COUNT_ZERO, PH_COUNT, PH_SCAN, PH_CURSOR, PH_PLACE, PH_SORT, PH_POSE_CSR, PH_CHI2, PH_CHI2_EDGES, PH_RED2, PH_W, PH_LM, PH_POSE_PART,
             PH_POSE_SUM, PH_MAXDIAG, PH_S_ZERO, PH_S_INIT, PH_SCHUR, PH_CHOL, PH_BSUB, PH_UPD_LM, PH_UPD_POSE, PH_RESTORE, PH_CULL, PH_BACKUP };

BIG_HD inline double robust_rho(double c, double delta, double d2) { return (c <= d2) ? c : 2 * sqrt(c) * delta - d2; }

// thread t of T runs its share of phase ph; (ai, ad) are the phase's scalar arguments (column index, lambda)
BIG_HD void big_phase(BigCtx& c, int ph, int ai, double ad, int t, int T) {
  const int P = c.P, L = c.L, E = c.E;
  const int np = c.ctl[0], nact = c.ctl[1], n = 6 * np;
  switch (ph) {
    case PH_COUNT_ZERO:
      for (int i = t; i <= L + 1; i += T) c.lstart[i] = 0;
      for (int i = t; i <= P + 1; i += T) c.pstart[i] = 0;
      break;
    case PH_COUNT:                       // active edges per landmark / pose (integer atomics: exact, order-free)
      for (int e = t; e < E; e += T)
        if (c.act[e]) { big_atomic_inc(&c.lstart[c.el[e] + 1]); big_atomic_inc(&c.pstart[c.ep[e] + 1]); }
      break;
    case PH_SCAN:                        // prefixes + free-pose numbering (sparse_optimizer.cpp:168-272: a pose without active edges leaves)
      if (t == 0) {
        for (int l = 0; l < L; ++l) c.lstart[l + 1] += c.lstart[l];
        int k = 0;
        for (int p = 0; p < P; ++p) {
          const int cnt = c.pstart[p + 1];
          if (p != c.fixed && cnt > 0) { c.pidx[p] = k; c.pose_of[k] = p; ++k; } else c.pidx[p] = -1;
        }
        for (int p = 0; p < P; ++p) c.pstart[p + 1] += c.pstart[p];
        c.ctl[0] = k; c.ctl[1] = c.lstart[L]; c.ctl[2] = 0;
      }
      break;
    case PH_CURSOR:
      for (int l = t; l < L; l += T) c.cursor[l] = c.lstart[l];
      break;
    case PH_PLACE:                       // landmark-major edge list, arbitrary order inside a landmark ...
      for (int e = t; e < E; e += T)
        if (c.act[e]) c.csr_e[big_atomic_inc(&c.cursor[c.el[e]])] = e;
      break;
    case PH_SORT:                        // ... then ordered by pose inside every landmark (one edge per (pose, landmark)): fixed order
      for (int l = t; l < L; l += T) {
        const int j0 = c.lstart[l], j1 = c.lstart[l + 1];
        for (int a = j0 + 1; a < j1; ++a) {
          const int e = c.csr_e[a]; const int key = c.ep[e];
          int b = a - 1;
          while (b >= j0 && c.ep[c.csr_e[b]] > key) { c.csr_e[b + 1] = c.csr_e[b]; --b; }
          c.csr_e[b + 1] = e;
        }
        for (int a = j0; a < j1; ++a) { c.csr_p[a] = c.ep[c.csr_e[a]]; c.csr_l[a] = l; }
      }
      break;
    case PH_POSE_CSR:                    // pose-major list of CSR positions, in landmark order (stable pass per pose)
      for (int p = t; p < P; p += T) {
        int k = c.pstart[p];
        for (int j = 0; j < nact; ++j) if (c.csr_p[j] == p) c.pcsr[k++] = j;
      }
      break;
    case PH_CHI2: {                      // robust chi2 over the active CSR entries -> red[t]
      double acc = 0;
      for (int j = t; j < nact; j += T) {
        double r[2];
        edge_eval<false>(c.poses + 7 * c.csr_p[j], c.lms + 3 * (size_t)c.csr_l[j], c.uv + 2 * (size_t)c.csr_e[j], c.cam, r, nullptr, nullptr);
        acc += robust_rho(r[0] * r[0] + r[1] * r[1], c.delta, c.d2);
      }
      c.red[t] = acc; c.red[T + t] = 0;
    } break;
    case PH_CHI2_EDGES: {                // the same over the edge arrays (outside the LM loop)
      double acc = 0;
      for (int e = t; e < E; e += T) {
        if (!c.act[e]) continue;
        double r[2];
        edge_eval<false>(c.poses + 7 * c.ep[e], c.lms + 3 * (size_t)c.el[e], c.uv + 2 * (size_t)e, c.cam, r, nullptr, nullptr);
        acc += robust_rho(r[0] * r[0] + r[1] * r[1], c.delta, c.d2);
      }
      c.red[t] = acc; c.red[T + t] = 0;
    } break;
    case PH_RED2:                        // second level of the two-level reductions: 32 partials each (sum of channel 0 / 1, or max)
      for (int g = t; g < (T + 31) / 32; g += T) {
        double s0 = 0, s1 = 0;
        for (int i = 32 * g; i < 32 * g + 32 && i < T; ++i) {
          if (ai) { s0 = fmax(s0, c.red[i]); } else { s0 += c.red[i]; s1 += c.red[T + i]; }
        }
        c.red2[2 * g] = s0; c.red2[2 * g + 1] = s1;
      }
      break;
    case PH_W:                           // W = rho' B^T A of every active edge at the linearisation point
      for (int j = t; j < nact; j += T) {
        double* w = c.W + 18 * (size_t)j;
        if (c.pidx[c.csr_p[j]] < 0) { for (int k = 0; k < 18; ++k) w[k] = 0; continue; }
        edge_W(c.poses + 7 * c.csr_p[j], c.lms + 3 * (size_t)c.csr_l[j], c.uv + 2 * (size_t)c.csr_e[j], c.cam, c.delta, c.d2, w);
      }
      break;
    case PH_LM:                          // landmark blocks Hll, bl (pose order)
      for (int l = t; l < L; l += T) {
        const int j0 = c.lstart[l], j1 = c.lstart[l + 1];
        if (j0 == j1) continue;
        double H[6] = {0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
        for (int j = j0; j < j1; ++j) {
          double r[2], A[6], B[12];
          edge_eval<true>(c.poses + 7 * c.csr_p[j], c.lms + 3 * (size_t)l, c.uv + 2 * (size_t)c.csr_e[j], c.cam, r, A, B);
          const double cc = r[0] * r[0] + r[1] * r[1];
          const double rho1 = (cc <= c.d2) ? 1.0 : c.delta / sqrt(cc);
          const double o0 = -r[0] * rho1, o1 = -r[1] * rho1;
          for (int i = 0; i < 3; ++i) b[i] += A[i] * o0 + A[3 + i] * o1;
          H[0] += rho1 * (A[0] * A[0] + A[3] * A[3]); H[1] += rho1 * (A[0] * A[1] + A[3] * A[4]);
          H[2] += rho1 * (A[0] * A[2] + A[3] * A[5]); H[3] += rho1 * (A[1] * A[1] + A[4] * A[4]);
          H[4] += rho1 * (A[1] * A[2] + A[4] * A[5]); H[5] += rho1 * (A[2] * A[2] + A[5] * A[5]);
        }
        for (int i = 0; i < 6; ++i) c.Hll[6 * (size_t)l + i] = H[i];
        for (int i = 0; i < 3; ++i) c.bl[3 * (size_t)l + i] = b[i];
      }
      break;
    case PH_POSE_PART:                   // pose blocks: (free pose, part) sums a strided share of the pose's edges
      for (int task = t; task < np * BIG_PARTS; task += T) {
        const int pi = task / BIG_PARTS, part = task - pi * BIG_PARTS, p = c.pose_of[pi];
        double H[21], b[6];
        for (int i = 0; i < 21; ++i) H[i] = 0;
        for (int i = 0; i < 6; ++i) b[i] = 0;
        for (int q = c.pstart[p] + part; q < c.pstart[p + 1]; q += BIG_PARTS) {
          const int j = c.pcsr[q];
          double r[2], A[6], B[12];
          edge_eval<true>(c.poses + 7 * p, c.lms + 3 * (size_t)c.csr_l[j], c.uv + 2 * (size_t)c.csr_e[j], c.cam, r, A, B);
          const double cc = r[0] * r[0] + r[1] * r[1];
          const double sr = (cc <= c.d2) ? 1.0 : sqrt(c.delta / sqrt(cc));
          for (int i = 0; i < 12; ++i) B[i] *= sr;
          const double g0 = -sr * r[0], g1 = -sr * r[1];
          int k2 = 0;
          for (int i = 0; i < 6; ++i) {
            b[i] += B[i] * g0 + B[6 + i] * g1;
            for (int jj = i; jj < 6; ++jj) H[k2++] += B[i] * B[jj] + B[6 + i] * B[6 + jj];
          }
        }
        double* o = c.part + 27 * (size_t)task;
        for (int i = 0; i < 21; ++i) o[i] = H[i];
        for (int i = 0; i < 6; ++i) o[21 + i] = b[i];
      }
      break;
    case PH_POSE_SUM:
      for (int i = t; i < np * 27; i += T) {
        const int pi = i / 27, k = i - 27 * pi;
        double v = 0;
        for (int part = 0; part < BIG_PARTS; ++part) v += c.part[27 * (size_t)(pi * BIG_PARTS + part) + k];
        c.Hd[i] = v;
      }
      break;
    case PH_MAXDIAG: {                   // computeLambdaInit: largest diagonal entry of H
      double md = 0;
      for (int pi = t; pi < np; pi += T)
        for (int i = 0; i < 6; ++i) md = fmax(md, fabs(c.Hd[27 * pi + sym21b(i, i)]));
      if (!c.fix_landmarks)
        for (int l = t; l < L; l += T)
          if (c.lstart[l] != c.lstart[l + 1])
            md = fmax(md, fmax(fabs(c.Hll[6 * (size_t)l]), fmax(fabs(c.Hll[6 * (size_t)l + 3]), fabs(c.Hll[6 * (size_t)l + 5]))));
      c.red[t] = md;
    } break;
    case PH_S_ZERO:
      for (size_t i = t; i < (size_t)n * n; i += T) c.S[i] = 0;
      break;
    case PH_S_INIT:                      // S = diag blocks (+ lambda), y = bp, Dinv = (Hll + lambda)^-1
      for (int i = t; i < np * 36; i += T) {
        const int pi = i / 36, k = i - 36 * pi, r = k / 6, cc = k - 6 * r;
        double v = c.Hd[27 * pi + (r <= cc ? sym21b(r, cc) : sym21b(cc, r))];
        if (r == cc) v += ad;
        c.S[(size_t)(6 * pi + r) * n + 6 * pi + cc] = v;
      }
      for (int i = t; i < n; i += T) c.y[i] = c.Hd[27 * (i / 6) + 21 + i % 6];
      if (!c.fix_landmarks)
        for (int l = t; l < L; l += T) {
          if (c.lstart[l] == c.lstart[l + 1]) continue;
          const double* H = c.Hll + 6 * (size_t)l;
          const double a = H[0] + ad, b = H[1], cc = H[2], d = H[3] + ad, e = H[4], f = H[5] + ad;
          const double c00 = d * f - e * e, c01 = cc * e - b * f, c02 = b * e - cc * d;
          const double id = 1.0 / (a * c00 + b * c01 + cc * c02);
          double* D = c.Dinv + 6 * (size_t)l;
          D[0] = c00 * id; D[1] = c01 * id; D[2] = c02 * id; D[3] = (a * f - cc * cc) * id; D[4] = (b * cc - a * e) * id; D[5] = (a * d - b * b) * id;
        }
      break;
    case PH_SCHUR:                       // thread (free pose a, i, j'): element (i, j') of every block (a, b <= a) and, for j' = 0, rhs entry i
      if (!c.fix_landmarks)
        for (int task = t; task < np * 36; task += T) {
          const int a = task / 36, k = task - 36 * a, i = k / 6, jc = k - 6 * i, p = c.pose_of[a];
          double ysum = 0;
          for (int q = c.pstart[p]; q < c.pstart[p + 1]; ++q) {        // a's edges in landmark order
            const int ja = c.pcsr[q], l = c.csr_l[ja];
            const double* wa = c.W + 18 * (size_t)ja + 3 * i;          // row i of W_al
            const double* D = c.Dinv + 6 * (size_t)l;
            const double y0 = wa[0] * D[0] + wa[1] * D[1] + wa[2] * D[2], y1 = wa[0] * D[1] + wa[1] * D[3] + wa[2] * D[4],
                         y2 = wa[0] * D[2] + wa[1] * D[4] + wa[2] * D[5];  // row i of Y = W_al Dinv_l
            if (jc == 0) { const double* b = c.bl + 3 * (size_t)l; ysum += y0 * b[0] + y1 * b[1] + y2 * b[2]; }
            for (int jb = c.lstart[l]; jb < c.lstart[l + 1]; ++jb) {    // l's edges in pose order
              const int bb = c.pidx[c.csr_p[jb]];
              if (bb < 0 || bb > a) continue;                            // lower triangle of blocks (whole diagonal block)
              const double* wb = c.W + 18 * (size_t)jb + 3 * jc;        // row j' of W_bl
              c.S[(size_t)(6 * a + i) * n + 6 * bb + jc] -= y0 * wb[0] + y1 * wb[1] + y2 * wb[2];
            }
          }
          if (jc == 0) c.y[6 * a + i] -= ysum;
        }
      break;
    case PH_CHOL: {                      // one column step of the LDL^T factorisation of [S ; y^T] (see ba.cu); ai = column j
      const int j = ai;
      const double inv = 1.0 / c.S[(size_t)j * n + j];
      if (t == 0) c.invd[j] = inv;
      const int TR = 32;                                   // threads per row
      for (int row = j + 1 + t / TR; row <= n; row += T / TR) {
        double* Ai = row < n ? c.S + (size_t)row * n : c.y;
        const double f = Ai[j] * inv;
        const int kmax = row < n ? row : n - 1;
        for (int k = j + 1 + (t % TR); k <= kmax; k += TR) Ai[k] -= f * c.S[(size_t)k * n + j];
      }
    } break;
    case PH_BSUB: {                      // back substitution, row j = ai: x_j = w_j / d_j, then w_i -= A_ji x_j for i < j
      const int j = ai;
      const double xj = c.y[j] * c.invd[j];
      if (t == 0) c.x[j] = xj;
      for (int i = t; i < j; i += T) c.y[i] -= c.S[(size_t)j * n + i] * xj;
    } break;
    case PH_BACKUP:
      for (int i = t; i < 7 * P; i += T) c.pbk[i] = c.poses[i];
      break;
    case PH_UPD_LM: {                    // landmark back-substitution + update (block_solver.hpp:422-444); scale share -> red
      double sc = 0;
      if (!c.fix_landmarks)
        for (int l = t; l < L; l += T) {
          double* X = c.lms + 3 * (size_t)l;
          c.lbk[3 * (size_t)l] = X[0]; c.lbk[3 * (size_t)l + 1] = X[1]; c.lbk[3 * (size_t)l + 2] = X[2];
          const int j0 = c.lstart[l], j1 = c.lstart[l + 1];
          if (j0 == j1) continue;
          const double* b = c.bl + 3 * (size_t)l;
          double c0 = b[0], c1 = b[1], c2 = b[2];
          for (int j = j0; j < j1; ++j) {
            const int pi = c.pidx[c.csr_p[j]];
            if (pi < 0) continue;
            const double* w = c.W + 18 * (size_t)j; const double* xp = c.x + 6 * pi;
            double t0 = 0, t1 = 0, t2 = 0;
            for (int i = 0; i < 6; ++i) { t0 += w[3 * i] * xp[i]; t1 += w[3 * i + 1] * xp[i]; t2 += w[3 * i + 2] * xp[i]; }
            c0 -= t0; c1 -= t1; c2 -= t2;
          }
          const double* D = c.Dinv + 6 * (size_t)l;
          const double x0 = D[0] * c0 + D[1] * c1 + D[2] * c2, x1 = D[1] * c0 + D[3] * c1 + D[4] * c2, x2 = D[2] * c0 + D[4] * c1 + D[5] * c2;
          sc += x0 * (ad * x0 + b[0]) + x1 * (ad * x1 + b[1]) + x2 * (ad * x2 + b[2]);
          X[0] += x0; X[1] += x1; X[2] += x2;
        }
      for (int i = t; i < n; i += T) sc += c.x[i] * (ad * c.x[i] + c.Hd[27 * (i / 6) + 21 + i % 6]);
      c.red[T + t] = sc;                 // channel 1 (channel 0 receives the new chi2 in the next phase... see big_run)
    } break;
    case PH_UPD_POSE:
      for (int pi = t; pi < np; pi += T) pose_oplus(c.poses + 7 * c.pose_of[pi], c.x + 6 * pi);
      break;
    case PH_RESTORE:
      for (int i = t; i < 7 * P; i += T) c.poses[i] = c.pbk[i];
      if (!c.fix_landmarks) for (int i = t; i < 3 * L; i += T) c.lms[i] = c.lbk[i];
      break;
    case PH_CULL: {                      // un-robustified chi2 > threshold -> inactive; counts -> red (culled, remaining)
      double culled = 0, remaining = 0;
      for (int j = t; j < nact; j += T) {
        double r[2];
        edge_eval<false>(c.poses + 7 * c.csr_p[j], c.lms + 3 * (size_t)c.csr_l[j], c.uv + 2 * (size_t)c.csr_e[j], c.cam, r, nullptr, nullptr);
        if (r[0] * r[0] + r[1] * r[1] > c.cull_chi2) { c.act[c.csr_e[j]] = 0; culled += 1; } else remaining += 1;
      }
      c.red[t] = culled; c.red[T + t] = remaining;
    } break;
  }
}

// ---- driver: identical on the device (X = cluster executor) and on the host (X = sequential executor) -------------------------
template <class X>
BIG_HD void big_reduce(BigCtx& c, X& x, int T, bool is_max, double& v0, double& v1) {
  x.par(c, PH_RED2, is_max ? 1 : 0, 0.0);
  v0 = 0; v1 = 0;
  for (int g = 0; g < (T + 31) / 32; ++g) {
    if (is_max) v0 = fmax(v0, c.red2[2 * g]); else { v0 += c.red2[2 * g]; v1 += c.red2[2 * g + 1]; }
  }
  x.barrier();                           // everyone has read red2 before a later reduction rewrites it
}

template <class X>
BIG_HD void big_setup(BigCtx& c, X& x) {
  x.par(c, PH_COUNT_ZERO, 0, 0); x.par(c, PH_COUNT, 0, 0); x.par(c, PH_SCAN, 0, 0); x.par(c, PH_CURSOR, 0, 0);
  x.par(c, PH_PLACE, 0, 0); x.par(c, PH_SORT, 0, 0); x.par(c, PH_POSE_CSR, 0, 0);
}

template <class X>
BIG_HD void big_run(BigCtx& c, X& x, const flv_ba_params& prm, flv_ba_stats& st, double* trace) {
  const int T = x.threads();
  double u;
  st.iterations_run = 0; st.n_culled = 0; st.ok = 1; st.reserved = 0;
  st.chi2_initial = st.chi2_after1 = st.chi2_final = 0; st.lambda_final = 0;
  x.par(c, PH_CHI2_EDGES, 0, 0);
  big_reduce(c, x, T, false, st.chi2_initial, u);
  double lambda = 0;
  for (int phase = 0; phase < 2; ++phase) {
    const int iters = phase == 0 ? prm.iters1 : prm.iters2;
    big_setup(c, x);
    const int np = c.ctl[0], n = 6 * np;
    double ni = 2, currentChi = 0;
    for (int it = 0; it < iters; ++it) {
      if (it == 0) { x.par(c, PH_CHI2, 0, 0); big_reduce(c, x, T, false, currentChi, u); }
      x.par(c, PH_W, 0, 0); x.par(c, PH_LM, 0, 0); x.par(c, PH_POSE_PART, 0, 0); x.par(c, PH_POSE_SUM, 0, 0);
      if (it == 0) {
        double md;
        x.par(c, PH_MAXDIAG, 0, 0);
        big_reduce(c, x, T, true, md, u);
        lambda = 1e-5 * md; ni = 2;
      }
      double rho = 0;
      int qmax = 0;
      do {
        x.par(c, PH_S_ZERO, 0, 0); x.par(c, PH_S_INIT, 0, lambda); x.par(c, PH_SCHUR, 0, 0);
        bool ok2 = true;
        for (int j = 0; j < n; ++j) {
          if (!(c.S[(size_t)j * n + j] > 0)) { ok2 = false; break; }          // same value in every thread (read after a barrier)
          x.par(c, PH_CHOL, j, 0);
        }
        double scale = 0, tempChi = 1.7976931348623157e308;
        if (ok2) {
          for (int j = n - 1; j >= 0; --j) x.par(c, PH_BSUB, j, 0);
          x.par(c, PH_BACKUP, 0, 0);
          x.par(c, PH_UPD_LM, 0, lambda);                  // scale shares -> channel 1 of red
          double keep_sc;
          { double a0; big_reduce(c, x, T, false, a0, keep_sc); }
          x.par(c, PH_UPD_POSE, 0, 0);
          x.par(c, PH_CHI2, 0, 0);
          big_reduce(c, x, T, false, tempChi, u);
          scale = keep_sc;
        }
        rho = (currentChi - tempChi) / (scale + 1e-3);
        if (rho > 0 && isfinite(tempChi)) {
          double alpha = 1. - (2 * rho - 1) * (2 * rho - 1) * (2 * rho - 1);
          alpha = fmin(alpha, 2. / 3.);
          lambda *= fmax(1. / 3., alpha);
          ni = 2; currentChi = tempChi;
        } else {
          lambda *= ni; ni *= 2;
          if (ok2) x.par(c, PH_RESTORE, 0, 0);
          if (!isfinite(lambda)) break;
        }
        ++qmax;
      } while (rho < 0 && qmax < 10);
      if (trace && x.leader() && st.iterations_run < 32) {
        double* tr = trace + 4 * st.iterations_run;
        tr[0] = currentChi; tr[1] = lambda; tr[2] = rho; tr[3] = (double)qmax;
      }
      ++st.iterations_run;
      if (qmax == 10 || rho == 0 || !isfinite(lambda)) break;
    }
    if (phase == 0) {
      x.par(c, PH_CHI2, 0, 0);
      big_reduce(c, x, T, false, st.chi2_after1, u);
      x.par(c, PH_CULL, 0, 0);
      double culled, remaining;
      big_reduce(c, x, T, false, culled, remaining);
      st.n_culled = (int)(culled + 0.5);
      if ((int)(remaining + 0.5) < prm.min_edges_after_cull) { st.ok = 0; break; }
    }
  }
  x.par(c, PH_CHI2_EDGES, 0, 0);
  big_reduce(c, x, T, false, st.chi2_final, u);
  st.lambda_final = lambda;
}

struct HostExec {                        // sequential emulation: "thread" t of T one after the other; barriers are sequence points
  int T;
  BIG_HD int threads() const { return T; }
  BIG_HD bool leader() const { return true; }
  BIG_HD void barrier() {}
  BIG_HD void par(BigCtx& c, int ph, int ai, double ad) { for (int t = 0; t < T; ++t) big_phase(c, ph, ai, ad, t, T); }
};

struct DevExec {
  int t;
  __device__ int threads() const { return BIG_T; }
  __device__ bool leader() const { return t == 0; }
  __device__ void barrier() { cg::this_cluster().sync(); }
  __device__ void par(BigCtx& c, int ph, int ai, double ad) { big_phase(c, ph, ai, ad, t, BIG_T); cg::this_cluster().sync(); }
};

struct BigArgs {
  const flv_ba_problem* problems; flv_ba_params prm;
  double* poses; double* lms; const int* ep; const int* el; const double* uv; uint8_t* active; flv_ba_stats* stats;
  int max_poses, max_lms, max_edges;
  unsigned char* ws; size_t ws_stride;
  double* trace;
};

BIG_HD inline void big_make_ctx(BigCtx& c, const BigArgs& a, int s) {
  const flv_ba_problem& pb = a.problems[s];
  c.P = pb.n_poses; c.L = pb.n_landmarks; c.E = pb.n_edges; c.fixed = pb.fixed_pose; c.fix_landmarks = pb.fix_landmarks;
  c.cam.fx = pb.fx; c.cam.fy = pb.fy; c.cam.cx = pb.cx; c.cam.cy = pb.cy;
  c.delta = a.prm.huber_delta; c.d2 = c.delta * c.delta; c.cull_chi2 = a.prm.cull_chi2;
  c.poses = a.poses + (size_t)s * a.max_poses * 7; c.lms = a.lms + (size_t)s * a.max_lms * 3;
  c.ep = a.ep + (size_t)s * a.max_edges; c.el = a.el + (size_t)s * a.max_edges; c.uv = a.uv + (size_t)s * a.max_edges * 2;
  c.act = a.active + (size_t)s * a.max_edges;
  big_bind(c, a.ws + (size_t)s * a.ws_stride, a.max_poses, a.max_lms, a.max_edges);
}

__global__ void __launch_bounds__(BIG_THREADS, 1) ba_big_kernel(BigArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  const int s = blockIdx.x / BIG_CTAS;
  BigCtx c;
  big_make_ctx(c, a, s);
  DevExec x{(int)cluster.block_rank() * BIG_THREADS + (int)threadIdx.x};
  flv_ba_stats st;
  if (c.P < 1 || c.P > BIG_MAX_POSES || c.E < 0) {
    st.iterations_run = 0; st.n_culled = 0; st.ok = 0; st.reserved = 1; st.chi2_initial = st.chi2_after1 = st.chi2_final = st.lambda_final = 0;
    if (x.t == 0) a.stats[s] = st;
    return;
  }
  big_run(c, x, a.prm, st, a.trace ? a.trace + (size_t)s * 32 * 4 : nullptr);
  if (x.t == 0) a.stats[s] = st;
}

}  // namespace

// ---- host entry points (called from flv_ba_reserve / flv_ba_optimize in ba.cu) -----------------------------------------------
size_t flv_ba_big_ws_bytes(int max_poses, int max_lms, int max_edges) { return big_ws_bytes(max_poses, max_lms, max_edges); }
int flv_ba_big_max_poses() { return BIG_MAX_POSES; }

// device-resident arrays (strides max_*), workspace of n_streams * flv_ba_big_ws_bytes; trace may be null
cudaError_t flv_ba_big_launch(int n_streams, const flv_ba_problem* d_problems, const flv_ba_params* prm, double* d_poses, double* d_lms,
                              const int* d_ep, const int* d_el, const double* d_uv, uint8_t* d_act, flv_ba_stats* d_stats, int max_poses,
                              int max_lms, int max_edges, unsigned char* d_ws, double* d_trace, cudaStream_t stream) {
  BigArgs a;
  a.problems = d_problems; a.prm = *prm; a.poses = d_poses; a.lms = d_lms; a.ep = d_ep; a.el = d_el; a.uv = d_uv; a.active = d_act;
  a.stats = d_stats; a.max_poses = max_poses; a.max_lms = max_lms; a.max_edges = max_edges; a.ws = d_ws;
  a.ws_stride = big_ws_bytes(max_poses, max_lms, max_edges); a.trace = d_trace;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_streams * BIG_CTAS)); cfg.blockDim = dim3(BIG_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = BIG_CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, ba_big_kernel, a);
}

extern "C" {

/* TEST AID (no GPU involved): the large-window solver's phases executed sequentially on the host over host arrays -- the same
 * driver and phase code the kernel runs -- so that the CPU test suite can compare its arithmetic with the oracle.  One window. */
int flv_ba_big_emulate_host(const flv_ba_problem* problem, const flv_ba_params* prm, double* poses, double* landmarks, const int* edge_pose,
                            const int* edge_lm, const double* edge_uv, uint8_t* edge_active, flv_ba_stats* stats, int emulated_threads) {
  if (!problem || !prm || !poses || !landmarks || !edge_pose || !edge_lm || !edge_uv || !edge_active || !stats) return FLV_ERR_INVALID;
  if (problem->n_poses < 1 || problem->n_poses > BIG_MAX_POSES || emulated_threads < 32 || emulated_threads > BIG_T) return FLV_ERR_INVALID;
  const int MP = problem->n_poses, ML = problem->n_landmarks > 0 ? problem->n_landmarks : 1, ME = problem->n_edges > 0 ? problem->n_edges : 1;
  std::vector<unsigned char> ws(big_ws_bytes(MP, ML, ME) + 16);
  BigArgs a;
  a.problems = problem; a.prm = *prm; a.poses = poses; a.lms = landmarks; a.ep = edge_pose; a.el = edge_lm; a.uv = edge_uv;
  a.active = edge_active; a.stats = stats; a.max_poses = MP; a.max_lms = ML; a.max_edges = ME;
  a.ws = (unsigned char*)(((uintptr_t)ws.data() + 15) & ~(uintptr_t)15); a.ws_stride = 0; a.trace = nullptr;
  BigCtx c;
  big_make_ctx(c, a, 0);
  HostExec x{emulated_threads};
  big_run(c, x, *prm, *stats, nullptr);
  return FLV_OK;
}

}  // extern "C"
