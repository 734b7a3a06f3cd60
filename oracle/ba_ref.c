/* CPU restatement of the g2o call path FLVIS uses for bundle adjustment -- TEST INFRASTRUCTURE (oracle)
 * and the timed CPU baseline ("port": g2o itself cannot be built here, it needs Eigen3 + CHOLMOD).
 *
 * PARITY UNPINNED: the reference ships no golden vectors for this path and g2o is unbuildable in this
 * container; this file follows the vendored sources line by line and is validated by known-answer tests
 * (tests/test_ba_oracle_cpu.py: numeric Jacobians, zero-noise recovery, ba_demo-shaped problems).
 *
 * Follows (paths under /root/reference/3rdPartLib/g2o/g2o/ unless noted):
 *   residual      types/sba/types_six_dof_expmap.h:209-214, cam_project types_six_dof_expmap.cpp:427-433
 *   Jacobians     types/sba/types_six_dof_expmap.cpp:389-425
 *   SE3 algebra   types/slam3d/se3quat.h:99-116 (operator*), :212 (map), :218-260 (exp), :280 (normalize)
 *   vertex oplus  types/sba/types_six_dof_expmap.h:98-101 (pose <- exp(dx)*pose), types/sba/types_sba.h:149-153
 *   quadratic form + Huber   core/base_binary_edge.hpp:62-134, core/robust_kernel_impl.cpp:65-78,
 *                 core/base_edge.h:117-123 (robustInformation = rho' * Omega)
 *   system build / Schur / back-substitution   core/block_solver.hpp:463-521, :328-447, :525-565
 *   LM control    core/optimization_algorithm_levenberg.cpp:58-175
 *   outer loop    core/sparse_optimizer.cpp:366-430 (stop when solve() != OK), :102-116 (activeRobustChi2)
 *   active sets   core/sparse_optimizer.cpp:168-272 (vertices with >= 1 active edge; fixed => no index)
 * and the callers  /root/reference/src/backend/vo_localmap.cpp:292-319 (12 it, chi2>3 cull, 8 it) and
 *                  /root/reference/src/processing/optimize_in_frame.cpp:10-90 (pose-only 2+2).
 * Linear solver: dense Cholesky (the reference uses CHOLMOD / Eigen LDLT on the same SPD system).
 * Summation order = g2o's: edges in array (insertion) order, landmarks by index.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int n_poses, n_landmarks, n_edges;
  int fixed_pose;    /* -1 none */
  int fix_landmarks; /* pose-only BA */
  double fx, fy, cx, cy;
} oracle_ba_problem;

typedef struct {
  int iterations_run, n_culled, ok, reserved;
  double chi2_initial, chi2_after1, chi2_final, lambda_final;
} oracle_ba_stats;

/* ---- SE3 (quaternion xyzw + t) -------------------------------------------------------------- */
static void q_rotate(const double* q, const double* v, double* o) {
  /* Eigen QuaternionBase::_transformVector: uv = 2 * (q.vec x v); o = v + w*uv + q.vec x uv */
  double ux = q[1] * v[2] - q[2] * v[1], uy = q[2] * v[0] - q[0] * v[2], uz = q[0] * v[1] - q[1] * v[0];
  ux += ux; uy += uy; uz += uz;
  o[0] = v[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
  o[1] = v[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
  o[2] = v[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
}
static void q_mul(const double* a, const double* b, double* o) {
  double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
static void q_normalize_g2o(double* q) { /* se3quat.h:280 */
  if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
static void q_to_R(const double* q, double* R) { /* Eigen toRotationMatrix, row-major R[3*r+c] */
  double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
static void R_to_q(const double* m, double* q) { /* Eigen quaternionbase_assign_impl<Matrix3> */
  double t = m[0] + m[4] + m[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t; t = 0.5 / t;
    q[0] = (m[7] - m[5]) * t; q[1] = (m[2] - m[6]) * t; q[2] = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    q[i] = 0.5 * t; t = 0.5 / t;
    q[3] = (m[3 * k + j] - m[3 * j + k]) * t;
    q[j] = (m[3 * j + i] + m[3 * i + j]) * t;
    q[k] = (m[3 * k + i] + m[3 * i + k]) * t;
  }
}
static void se3_exp(const double* u, double* q, double* t) { /* se3quat.h:218-260; u = [omega, upsilon] */
  double wx = u[0], wy = u[1], wz = u[2];
  double theta = sqrt(wx * wx + wy * wy + wz * wz);
  double O[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
  double O2[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) O2[3 * r + c] = O[3 * r] * O[c] + O[3 * r + 1] * O[3 + c] + O[3 * r + 2] * O[6 + c];
  double a, b, c2, d;
  if (theta < 0.00001) { a = 1.0; b = 0.5; c2 = 0.5; d = 1.0 / 6.0; }
  else {
    a = sin(theta) / theta; b = (1 - cos(theta)) / (theta * theta);
    c2 = b; d = (theta - sin(theta)) / pow(theta, 3);
  }
  double R[9], V[9];
  for (int i = 0; i < 9; ++i) {
    double id = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
    R[i] = id + a * O[i] + b * O2[i];
    V[i] = id + c2 * O[i] + d * O2[i];
  }
  R_to_q(R, q);
  for (int r = 0; r < 3; ++r) t[r] = V[3 * r] * u[3] + V[3 * r + 1] * u[4] + V[3 * r + 2] * u[5];
}
static void pose_oplus(double* pose, const double* u) { /* estimate <- exp(u) * estimate, se3quat.h:99-105 */
  double qe[4], te[3], rt[3], qn[4];
  se3_exp(u, qe, te);
  q_rotate(qe, pose + 4, rt);
  q_mul(qe, pose, qn);
  pose[4] = te[0] + rt[0]; pose[5] = te[1] + rt[1]; pose[6] = te[2] + rt[2];
  q_normalize_g2o(qn);
  pose[0] = qn[0]; pose[1] = qn[1]; pose[2] = qn[2]; pose[3] = qn[3];
}

/* ---- one edge --------------------------------------------------------------------------------- */
void oracle_ba_edge(const double* pose, const double* X, const double* uv, double fx, double fy, double cx,
                    double cy, double* err2, double* Jp23, double* Jc26) {
  double Xc[3];
  q_rotate(pose, X, Xc);
  Xc[0] += pose[4]; Xc[1] += pose[5]; Xc[2] += pose[6];
  double x = Xc[0], y = Xc[1], z = Xc[2];
  err2[0] = uv[0] - (x / z * fx + cx);
  err2[1] = uv[1] - (y / z * fy + cy);
  if (!Jp23) return;
  double z2 = z * z;
  double R[9];
  q_to_R(pose, R);
  double tmp[6] = {fx, 0, -x / z * fx, 0, fy, -y / z * fy};
  for (int r = 0; r < 2; ++r)
    for (int c = 0; c < 3; ++c)
      Jp23[3 * r + c] = -1. / z * (tmp[3 * r] * R[c] + tmp[3 * r + 1] * R[3 + c] + tmp[3 * r + 2] * R[6 + c]);
  Jc26[0] = x * y / z2 * fx; Jc26[1] = -(1 + (x * x / z2)) * fx; Jc26[2] = y / z * fx;
  Jc26[3] = -1. / z * fx; Jc26[4] = 0; Jc26[5] = x / z2 * fx;
  Jc26[6] = (1 + y * y / z2) * fy; Jc26[7] = -x * y / z2 * fy; Jc26[8] = -x / z * fy;
  Jc26[9] = 0; Jc26[10] = -1. / z * fy; Jc26[11] = y / z2 * fy;
}

void oracle_se3_oplus(double* pose, const double* u) { pose_oplus(pose, u); }

/* ---- solver state ---------------------------------------------------------------------------------- */
typedef struct {
  const oracle_ba_problem* pb;
  double *poses, *lms;
  const int *ep, *el;
  const double* uv;
  const unsigned char* act;
  double delta;
  int P, L, E, np;           /* np = number of free active poses */
  int* pidx;                 /* pose -> index in reduced system or -1 */
  int* lact;                 /* landmark active flag */
  double *Hpp, *bp;          /* (6np)^2, 6np */
  double *Hll, *bl;          /* L*9, L*3 */
  double *Hpl;               /* E*18 (6x3 per edge; unused rows for fixed poses) */
  double *S, *bs, *x, *xl, *coef, *Dinv;
  double *pbk, *lbk;
} ba_ws;

static double robust_chi2(const ba_ws* w) {
  const oracle_ba_problem* pb = w->pb;
  double chi = 0, d2 = w->delta * w->delta;
  for (int e = 0; e < w->E; ++e) {
    if (!w->act[e]) continue;
    double r[2];
    oracle_ba_edge(w->poses + 7 * w->ep[e], w->lms + 3 * w->el[e], w->uv + 2 * e, pb->fx, pb->fy, pb->cx, pb->cy, r, 0, 0);
    double c = r[0] * r[0] + r[1] * r[1];
    chi += (c <= d2) ? c : 2 * sqrt(c) * w->delta - d2;
  }
  return chi;
}

static void build_system(ba_ws* w) {
  const oracle_ba_problem* pb = w->pb;
  int n = 6 * w->np;
  memset(w->Hpp, 0, sizeof(double) * n * n);
  memset(w->bp, 0, sizeof(double) * n);
  memset(w->Hll, 0, sizeof(double) * 9 * w->L);
  memset(w->bl, 0, sizeof(double) * 3 * w->L);
  double d2 = w->delta * w->delta;
  for (int e = 0; e < w->E; ++e) {
    if (!w->act[e]) continue;
    int p = w->ep[e], l = w->el[e];
    double r[2], A[6], B[12];
    oracle_ba_edge(w->poses + 7 * p, w->lms + 3 * l, w->uv + 2 * e, pb->fx, pb->fy, pb->cx, pb->cy, r, A, B);
    double c = r[0] * r[0] + r[1] * r[1];
    double rho1 = (c <= d2) ? 1.0 : w->delta / sqrt(c);
    double orr[2] = {-r[0] * rho1, -r[1] * rho1};          /* omega_r = -Omega*r * rho' */
    int pi = w->pidx[p];
    int lfree = !pb->fix_landmarks;
    if (lfree) {
      double* H = w->Hll + 9 * l; double* b = w->bl + 3 * l;
      for (int i = 0; i < 3; ++i) {
        b[i] += A[i] * orr[0] + A[3 + i] * orr[1];
        for (int j = 0; j < 3; ++j) H[3 * i + j] += rho1 * (A[i] * A[j] + A[3 + i] * A[3 + j]);
      }
    }
    if (pi >= 0) {
      for (int i = 0; i < 6; ++i) {
        w->bp[6 * pi + i] += B[i] * orr[0] + B[6 + i] * orr[1];
        for (int j = 0; j < 6; ++j) w->Hpp[(6 * pi + i) * n + 6 * pi + j] += rho1 * (B[i] * B[j] + B[6 + i] * B[6 + j]);
      }
      if (lfree) {
        double* W = w->Hpl + 18 * e;                      /* 6x3 = B^T * wOmega * A */
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 3; ++j) W[3 * i + j] = rho1 * (B[i] * A[j] + B[6 + i] * A[3 + j]);
      }
    }
  }
}

static int chol_solve(double* A, double* b, int n) { /* in place LL^T; returns 0 if not SPD */
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0)) return 0;
    d = sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / d;
    }
  }
  for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= A[i * n + k] * b[k]; b[i] = s / A[i * n + i]; }
  for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= A[k * n + i] * b[k]; b[i] = s / A[i * n + i]; }
  return 1;
}

static void inv3(const double* m, double* o) { /* Eigen 3x3 inverse (cofactors / det) */
  double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  double id = 1.0 / det;
  o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

/* solve (H + lambda I) x = b with the Schur complement; returns 0 on failure */
static int solve_system(ba_ws* w, double lambda) {
  const oracle_ba_problem* pb = w->pb;
  int n = 6 * w->np;
  memcpy(w->S, w->Hpp, sizeof(double) * n * n);
  for (int i = 0; i < n; ++i) w->S[i * n + i] += lambda;
  memcpy(w->bs, w->bp, sizeof(double) * n);
  if (!pb->fix_landmarks) {
    memset(w->coef, 0, sizeof(double) * n);
    /* per-landmark edge lists: build CSR on the fly (edges in array order) */
    int* start = (int*)calloc(w->L + 1, sizeof(int));
    for (int e = 0; e < w->E; ++e) if (w->act[e]) start[w->el[e] + 1]++;
    for (int l = 0; l < w->L; ++l) start[l + 1] += start[l];
    int* fill = (int*)malloc(sizeof(int) * (w->L + 1)); memcpy(fill, start, sizeof(int) * (w->L + 1));
    int* list = (int*)malloc(sizeof(int) * (start[w->L] + 1));
    for (int e = 0; e < w->E; ++e) if (w->act[e]) list[fill[w->el[e]]++] = e;
    for (int l = 0; l < w->L; ++l) {
      if (!w->lact[l]) continue;
      double D[9]; memcpy(D, w->Hll + 9 * l, sizeof(D));
      D[0] += lambda; D[4] += lambda; D[8] += lambda;
      double* Di = w->Dinv + 9 * l;
      inv3(D, Di);
      double db[3];
      for (int i = 0; i < 3; ++i) db[i] = Di[3 * i] * w->bl[3 * l] + Di[3 * i + 1] * w->bl[3 * l + 1] + Di[3 * i + 2] * w->bl[3 * l + 2];
      for (int a = start[l]; a < start[l + 1]; ++a) {
        int e1 = list[a], p1 = w->pidx[w->ep[e1]];
        if (p1 < 0) continue;
        const double* B1 = w->Hpl + 18 * e1;
        double BD[18];
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 3; ++j) BD[3 * i + j] = B1[3 * i] * Di[j] + B1[3 * i + 1] * Di[3 + j] + B1[3 * i + 2] * Di[6 + j];
        for (int i = 0; i < 6; ++i) w->coef[6 * p1 + i] += B1[3 * i] * db[0] + B1[3 * i + 1] * db[1] + B1[3 * i + 2] * db[2];
        for (int b = start[l]; b < start[l + 1]; ++b) {
          int e2 = list[b], p2 = w->pidx[w->ep[e2]];
          if (p2 < 0) continue;
          const double* B2 = w->Hpl + 18 * e2;
          for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j)
              w->S[(6 * p1 + i) * n + 6 * p2 + j] -= BD[3 * i] * B2[3 * j] + BD[3 * i + 1] * B2[3 * j + 1] + BD[3 * i + 2] * B2[3 * j + 2];
        }
      }
    }
    for (int i = 0; i < n; ++i) w->bs[i] -= w->coef[i];
    free(start); free(fill); free(list);
  }
  memcpy(w->x, w->bs, sizeof(double) * n);
  if (n > 0 && !chol_solve(w->S, w->x, n)) return 0;
  if (!pb->fix_landmarks) {
    /* xl = Dinv * (bl - Hpl^T xp) */
    for (int l = 0; l < w->L; ++l) { w->xl[3 * l] = w->bl[3 * l]; w->xl[3 * l + 1] = w->bl[3 * l + 1]; w->xl[3 * l + 2] = w->bl[3 * l + 2]; }
    for (int e = 0; e < w->E; ++e) {
      if (!w->act[e]) continue;
      int p = w->pidx[w->ep[e]];
      if (p < 0) continue;
      const double* B = w->Hpl + 18 * e; double* c = w->xl + 3 * w->el[e];
      for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 6; ++i) c[j] -= B[3 * i + j] * w->x[6 * p + i];
    }
    for (int l = 0; l < w->L; ++l) {
      if (!w->lact[l]) { w->xl[3 * l] = w->xl[3 * l + 1] = w->xl[3 * l + 2] = 0; continue; }
      const double* Di = w->Dinv + 9 * l; double c[3] = {w->xl[3 * l], w->xl[3 * l + 1], w->xl[3 * l + 2]};
      for (int i = 0; i < 3; ++i) w->xl[3 * l + i] = Di[3 * i] * c[0] + Di[3 * i + 1] * c[1] + Di[3 * i + 2] * c[2];
    }
  }
  return 1;
}

static void setup_active(ba_ws* w) {
  const oracle_ba_problem* pb = w->pb;
  int* pedges = (int*)calloc(w->P, sizeof(int));
  memset(w->lact, 0, sizeof(int) * w->L);
  for (int e = 0; e < w->E; ++e) if (w->act[e]) { pedges[w->ep[e]]++; w->lact[w->el[e]] = 1; }
  w->np = 0;
  for (int p = 0; p < w->P; ++p) w->pidx[p] = (p != pb->fixed_pose && pedges[p] > 0) ? w->np++ : -1;
  free(pedges);
}

/* optional per-iteration trace for the tests: rows of {chi2 at the end of the iteration, lambda, rho of the last trial,
 * trials}; set with oracle_ba_set_trace before a solve, rows accumulate over both optimize() phases */
static double* g_trace = 0; static int g_trace_cap = 0, g_trace_n = 0;
void oracle_ba_set_trace(double* buf, int cap_rows) { g_trace = buf; g_trace_cap = cap_rows; g_trace_n = 0; }
int oracle_ba_trace_rows(void) { return g_trace_n; }

/* g2o SparseOptimizer::optimize(iters) with OptimizationAlgorithmLevenberg; returns iterations run */
static int lm_optimize(ba_ws* w, int iters, double* lambda_out, double* chi_out) {
  const oracle_ba_problem* pb = w->pb;
  setup_active(w);
  int n = 6 * w->np, done = 0;
  double lambda = 0, ni = 2;
  double chi_last = robust_chi2(w);
  for (int it = 0; it < iters; ++it) {
    double currentChi = robust_chi2(w), tempChi = currentChi;
    build_system(w);
    if (it == 0) {
      double md = 0;
      for (int i = 0; i < n; ++i) md = fmax(fabs(w->Hpp[i * n + i]), md);
      if (!pb->fix_landmarks)
        for (int l = 0; l < w->L; ++l) if (w->lact[l]) for (int j = 0; j < 3; ++j) md = fmax(fabs(w->Hll[9 * l + 4 * j]), md);
      lambda = 1e-5 * md; ni = 2;
    }
    double rho = 0; int qmax = 0;
    do {
      memcpy(w->pbk, w->poses, sizeof(double) * 7 * w->P);
      memcpy(w->lbk, w->lms, sizeof(double) * 3 * w->L);
      int ok2 = solve_system(w, lambda);
      if (ok2) {
        for (int p = 0; p < w->P; ++p) if (w->pidx[p] >= 0) pose_oplus(w->poses + 7 * p, w->x + 6 * w->pidx[p]);
        if (!pb->fix_landmarks)
          for (int l = 0; l < w->L; ++l) if (w->lact[l]) for (int j = 0; j < 3; ++j) w->lms[3 * l + j] += w->xl[3 * l + j];
      }
      tempChi = robust_chi2(w);
      if (!ok2) tempChi = 1.7976931348623157e308;
      rho = currentChi - tempChi;
      double scale = 0;
      if (ok2) {
        for (int i = 0; i < n; ++i) scale += w->x[i] * (lambda * w->x[i] + w->bp[i]);
        if (!pb->fix_landmarks)
          for (int l = 0; l < w->L; ++l) if (w->lact[l]) for (int j = 0; j < 3; ++j) scale += w->xl[3 * l + j] * (lambda * w->xl[3 * l + j] + w->bl[3 * l + j]);
      }
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && isfinite(tempChi)) {
        double alpha = 1. - pow((2 * rho - 1), 3);
        alpha = fmin(alpha, 2. / 3.);
        double sf = fmax(1. / 3., alpha);
        lambda *= sf; ni = 2; currentChi = tempChi;
      } else {
        lambda *= ni; ni *= 2;
        memcpy(w->poses, w->pbk, sizeof(double) * 7 * w->P);
        memcpy(w->lms, w->lbk, sizeof(double) * 3 * w->L);
        if (!isfinite(lambda)) break;
      }
      qmax++;
    } while (rho < 0 && qmax < 10);
    if (g_trace && g_trace_n < g_trace_cap) {
      double* tr = g_trace + 4 * g_trace_n++;
      tr[0] = currentChi; tr[1] = lambda; tr[2] = rho; tr[3] = (double)qmax;
    }
    done++;
    chi_last = currentChi;
    if (qmax == 10 || rho == 0 || !isfinite(lambda)) break;   /* Terminate */
  }
  *lambda_out = lambda; *chi_out = chi_last;
  return done;
}

/* local BA: optimize(iters1); cull active edges with chi2 > cull_chi2; optimize(iters2) */
int oracle_ba_optimize(const oracle_ba_problem* pb, int iters1, int iters2, double huber_delta, double cull_chi2,
                       int min_edges_after_cull, double* poses, double* lms, const int* ep, const int* el,
                       const double* uv, unsigned char* active, oracle_ba_stats* st) {
  ba_ws w; memset(&w, 0, sizeof(w));
  int P = pb->n_poses, L = pb->n_landmarks, E = pb->n_edges;
  w.pb = pb; w.poses = poses; w.lms = lms; w.ep = ep; w.el = el; w.uv = uv; w.act = active; w.delta = huber_delta;
  w.P = P; w.L = L; w.E = E;
  size_t n = 6 * (size_t)P;
  w.pidx = (int*)malloc(sizeof(int) * (P + 1)); w.lact = (int*)malloc(sizeof(int) * (L + 1));
  w.Hpp = (double*)malloc(sizeof(double) * (n * n + 1)); w.bp = (double*)malloc(sizeof(double) * (n + 1));
  w.Hll = (double*)malloc(sizeof(double) * (9 * L + 1)); w.bl = (double*)malloc(sizeof(double) * (3 * L + 1));
  w.Hpl = (double*)malloc(sizeof(double) * (18 * (size_t)E + 1));
  w.S = (double*)malloc(sizeof(double) * (n * n + 1)); w.bs = (double*)malloc(sizeof(double) * (n + 1));
  w.x = (double*)malloc(sizeof(double) * (n + 1)); w.xl = (double*)malloc(sizeof(double) * (3 * L + 1));
  w.coef = (double*)malloc(sizeof(double) * (n + 1)); w.Dinv = (double*)malloc(sizeof(double) * (9 * L + 1));
  w.pbk = (double*)malloc(sizeof(double) * (7 * P + 1)); w.lbk = (double*)malloc(sizeof(double) * (3 * L + 1));
  memset(st, 0, sizeof(*st));
  st->ok = 1;
  st->chi2_initial = robust_chi2(&w);
  double lam = 0, chi = 0;
  st->iterations_run = lm_optimize(&w, iters1, &lam, &chi);
  st->chi2_after1 = robust_chi2(&w);
  int remaining = 0;
  for (int e = 0; e < E; ++e) {
    if (!active[e]) continue;
    double r[2];
    oracle_ba_edge(poses + 7 * ep[e], lms + 3 * el[e], uv + 2 * e, pb->fx, pb->fy, pb->cx, pb->cy, r, 0, 0);
    if (r[0] * r[0] + r[1] * r[1] > cull_chi2) { active[e] = 0; st->n_culled++; } else remaining++;
  }
  if (remaining < min_edges_after_cull) st->ok = 0;      /* optimize_in_frame.cpp:70-73 */
  else st->iterations_run += lm_optimize(&w, iters2, &lam, &chi);
  st->chi2_final = robust_chi2(&w);
  st->lambda_final = lam;
  free(w.pidx); free(w.lact); free(w.Hpp); free(w.bp); free(w.Hll); free(w.bl); free(w.Hpl); free(w.S); free(w.bs);
  free(w.x); free(w.xl); free(w.coef); free(w.Dinv); free(w.pbk); free(w.lbk);
  return 0;
}
