// K4 -- Shi-Tomasi corner detection == cv::goodFeaturesToTrack(img, N, q, d) (blockSize 3, Sobel 3).
//
// Reference call sites: src/processing/feature_dem.cpp:160 (redetect) and :221 (detect).
// Bit-level contract: oracle/gftt_ref.py (pinned to cv2 4.13.0 by tests/golden/gftt_*.npz).
//
//   corner_response_kernel  one warp per 26-column x 32-row strip marches down the image with the
//       I / covariance / box-sum / eigenvalue rows in registers (neighbours by warp shuffle, no shared
//       memory): min-eigenvalue response, 3x3 non-max test and candidate emission fused, the response
//       map is never written (only when the debug flag asks for it).  HBM traffic = the u8 image once.
//   mindist_fast_kernel     one CTA per stream: threshold q*max, take the strongest <= 8192 candidates
//       (a prefix of the priority order, cut at a value-histogram bin), sort them in shared memory,
//       and decide OpenCV's sequential greedy "accept if no stronger accepted corner within d" as a
//       parallel fixed point (a corner's fate depends only on stronger corners, so iterating "reject
//       if a stronger neighbour is accepted / accept once all stronger neighbours are rejected"
//       reproduces the sequential result exactly).  If the prefix yields fewer than N corners and
//       candidates were left out, the stream is flagged and
//   mindist_full_kernel     redoes it over all candidates out of global memory (rare slow path).
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
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ------------------------------------------------------------------------------------------------
constexpr int CS_ROWS = 32;   // output rows per strip
constexpr int CS_COLS = 26;   // output columns per warp (32 lanes - 3 halo lanes each side)
constexpr int CS_WARPS = 4;

struct Row3 { float l, c, r; };      // pixel values are small integers: exact in float, converted once per loaded row

// FAST: the strip and its halo lie strictly inside the image and left of OpenCV's scalar tail (x < w - w%32): no
// reflection selects, no range predicates, FMA body everywhere.  The generic instantiation handles the border strips.
template <bool FAST>
__device__ __forceinline__ void corner_strip(const uint8_t* __restrict__ img, int pitch, int w, int h, int x, int y0, int lane,
                                             float* __restrict__ eig_s, float pre_thr, unsigned long long* __restrict__ cand,
                                             int* __restrict__ ncand_s, int cand_cap, float& best) {
  const bool in_x = FAST || (x >= 0 && x < w);
  const int xl = FAST ? x : clampi(reflect101(x, w), 0, w - 1);      // column this lane loads
  const bool left_from_up = !FAST && (x - 1 < 0);                  // reflect101(-1) = 1  -> neighbour value sits in lane+1
  const bool right_from_down = !FAST && (x + 1 > w - 1);           // reflect101(w) = w-2 -> neighbour value sits in lane-1
  const float k1 = (float)(1.0 / (4 * 3 * 255.0));
  const float k0 = 2.f * k1;
  const bool body = FAST || (x < w - (w & 31));
  const bool out_lane = lane >= 3 && lane < 3 + CS_COLS && in_x;

  auto load_raw = [&](int r) -> float {
    const int yy = FAST ? r : clampi(reflect101(r, h), 0, h - 1);
    return (float)img[(size_t)yy * pitch + xl];
  };
  auto make_row = [&](float c) -> Row3 {
    const float up = __shfl_down_sync(FULL, c, 1), dn = __shfl_up_sync(FULL, c, 1);
    Row3 o;
    o.c = c;
    o.l = left_from_up ? up : dn;
    o.r = right_from_down ? dn : up;
    return o;
  };

  // software prefetch: the row consumed at step r was requested 4 steps earlier (ncu: 52 % long-scoreboard without it)
  Row3 Ia, Ib = make_row(load_raw(y0 - 3)), Ic = make_row(load_raw(y0 - 2));
  float pf0 = load_raw(y0 - 1), pf1 = load_raw(y0), pf2 = load_raw(y0 + 1), pf3 = load_raw(y0 + 2);
  double Rm[3] = {0, 0, 0}, R0[3] = {0, 0, 0}, Rp[3] = {0, 0, 0};
  float Eb = 0.f, m3a = 0.f, m3b = 0.f, nb = 0.f;   // eig row y-1 (b); nb = max(left,right) of row b
  for (int r = y0 - 2; r <= y0 + CS_ROWS + 1; ++r) {
    Ia = Ib; Ib = Ic; Ic = make_row(pf0);
    pf0 = pf1; pf1 = pf2; pf2 = pf3; pf3 = load_raw(r + 5 < h || !FAST ? r + 5 : h - 1);
    // covariance of row r at this lane's column (only inside the image)
    float vxx = 0.f, vxy = 0.f, vyy = 0.f;
    if (FAST || (r >= 0 && r < h && in_x)) {
      const float dx = __fmaf_rn(k1, __fadd_rn(__fsub_rn(Ia.r, Ia.l), __fsub_rn(Ic.r, Ic.l)), __fmul_rn(k0, __fsub_rn(Ib.r, Ib.l)));
      float sa, sc;
      if (body) {
        sa = __fmaf_rn(k1, Ia.r, __fmaf_rn(k0, Ia.c, __fmul_rn(k1, Ia.l)));
        sc = __fmaf_rn(k1, Ic.r, __fmaf_rn(k0, Ic.c, __fmul_rn(k1, Ic.l)));
      } else {
        sa = __fadd_rn(__fadd_rn(__fmul_rn(k1, Ia.l), __fmul_rn(k0, Ia.c)), __fmul_rn(k1, Ia.r));
        sc = __fadd_rn(__fadd_rn(__fmul_rn(k1, Ic.l), __fmul_rn(k0, Ic.c)), __fmul_rn(k1, Ic.r));
      }
      const float dy = __fsub_rn(sc, sa);
      vxx = __fmul_rn(dx, dx); vxy = __fmul_rn(dx, dy); vyy = __fmul_rn(dy, dy);
    }
    // horizontal 3-sum (column reflection = take the other neighbour twice); exact in double
    Rm[0] = R0[0]; Rm[1] = R0[1]; Rm[2] = R0[2];
    R0[0] = Rp[0]; R0[1] = Rp[1]; R0[2] = Rp[2];
    {
      const float v[3] = {vxx, vxy, vyy};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float up = __shfl_down_sync(FULL, v[k], 1), dn = __shfl_up_sync(FULL, v[k], 1);
        const float lf = left_from_up ? up : dn, rt = right_from_down ? dn : up;
        Rp[k] = (double)lf + (double)v[k] + (double)rt;
      }
    }
    // eigenvalue of row y = r-1 (rows y-1, y, y+1 of R are Rm, R0, Rp; row reflection at the image border)
    const int y = r - 1;
    float e = 0.f;
    if (FAST || (y >= 0 && y < h)) {
      const bool top = !FAST && (y == 0), bot = !FAST && (y == h - 1);
      double sxx = R0[0] + (top ? Rp[0] : Rm[0]) + (bot ? Rm[0] : Rp[0]);
      double sxy = R0[1] + (top ? Rp[1] : Rm[1]) + (bot ? Rm[1] : Rp[1]);
      double syy = R0[2] + (top ? Rp[2] : Rm[2]) + (bot ? Rm[2] : Rp[2]);
      const float a = __fmul_rn((float)sxx, 0.5f), b = (float)sxy, c = __fmul_rn((float)syy, 0.5f);
      const float t = __fsub_rn(a, c);
      e = __fsub_rn(__fadd_rn(a, c), __fsqrt_rn(__fadd_rn(__fmul_rn(t, t), __fmul_rn(b, b))));
      if (out_lane && y >= y0 && y < y0 + CS_ROWS) {
        best = fmaxf(best, e);
        if (eig_s) eig_s[(size_t)y * w + x] = e;
      }
    }
    const float el = __shfl_up_sync(FULL, e, 1), er = __shfl_down_sync(FULL, e, 1);
    const float nc = fmaxf(el, er);
    const float m3c = fmaxf(nc, e);
    // non-max test of row yy = y-1 (its neighbours are rows a, c and its own left/right)
    const int yy = y - 1;
    const bool is = out_lane && yy >= y0 && yy < y0 + CS_ROWS && (FAST || (yy >= 1 && yy < h - 1 && x >= 1 && x < w - 1)) &&
                    Eb > pre_thr && Eb > 0.f && Eb >= m3a && Eb >= m3c && Eb >= nb;
    const unsigned mask = __ballot_sync(FULL, is);
    if (mask) {
      int base = 0;
      if (lane == 0) base = atomicAdd(ncand_s, __popc(mask));
      base = __shfl_sync(FULL, base, 0);
      if (is) {
        const int pos = base + __popc(mask & ((1u << lane) - 1));
        if (pos < cand_cap)
          cand[pos] = ((unsigned long long)__float_as_uint(Eb) << 32) | ((unsigned)yy << 16) | (unsigned)x;
      }
    }
    m3a = m3b; Eb = e; m3b = m3c; nb = nc;
  }
}

__global__ void __launch_bounds__(CS_WARPS * 32)
corner_response_kernel(const uint8_t* __restrict__ img_base, size_t stream_stride, int pitch, int w, int h,
                       int nsx, int nsy, float* __restrict__ eig_out, int* __restrict__ eigmax, double quality,
                       unsigned long long* __restrict__ cand_base, int* __restrict__ ncand, int cand_cap) {
  const int lane = threadIdx.x & 31;
  const int strip = blockIdx.x * CS_WARPS + (threadIdx.x >> 5);
  if (strip >= nsx * nsy) return;
  const int s = blockIdx.y;
  const int sx = strip % nsx, sy = strip / nsx;
  const uint8_t* img = img_base + (size_t)s * stream_stride;
  const int x = sx * CS_COLS - 3 + lane;
  const int y0 = sy * CS_ROWS;
  const float pre_thr = (float)((double)__int_as_float(*(volatile int*)&eigmax[s]) * quality);
  float* eig_s = eig_out ? eig_out + (size_t)s * w * h : nullptr;
  unsigned long long* cand = cand_base + (size_t)s * cand_cap;
  const int x_lo = sx * CS_COLS - 3, x_hi = x_lo + 31;
  // rows touched: y0-3 .. y0+CS_ROWS+6 (incl. prefetch); all lanes' columns and their neighbours inside, FMA body
  const bool fast = x_lo >= 1 && x_hi <= w - 2 && x_hi < w - (w & 31) && y0 - 3 >= 1 && y0 + CS_ROWS + 2 <= h - 2;
  float best = 0.f;
  if (fast) corner_strip<true>(img, pitch, w, h, x, y0, lane, eig_s, pre_thr, cand, &ncand[s], cand_cap, best);
  else corner_strip<false>(img, pitch, w, h, x, y0, lane, eig_s, pre_thr, cand, &ncand[s], cand_cap, best);
  int bi = __float_as_int(best);
  bi = __reduce_max_sync(FULL, bi);
  if (lane == 0 && bi > 0) atomicMax(&eigmax[s], bi);
}

// ------------------------------------------------------------------------------------------------
constexpr int MD_THREADS = 1024;
constexpr int SEL_CAP = 16384;   // candidates handled by the shared-memory fast path
constexpr int HBINS = 4096;      // histogram over float bits [30:19]

__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
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

// Greedy min-distance selection over `n` keys sorted in DESCENDING priority order (index == rank).
// keys: shared or global; cstart/ccount: ints [ncell+1]/[ncell]; items: ranks grouped by cell; state: u8[n].
template <typename KeyT, typename ItemT>
__device__ void greedy_fixed_point(const KeyT* keys, int n, int cell, int gw, int gh, double d2,
                                   int* cstart, int* ccount, ItemT* items, unsigned char* state, int* warp_sums,
                                   int* sh_total) {
  const int tid = threadIdx.x;
  const int ncell = gw * gh;
  for (int i = tid; i < ncell; i += MD_THREADS) ccount[i] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += MD_THREADS) {
    const unsigned lo = (unsigned)keys[i];
    atomicAdd(&ccount[((lo >> 16) / cell) * gw + (lo & 0xffff) / cell], 1);
  }
  __syncthreads();
  const int per = (ncell + MD_THREADS - 1) / MD_THREADS;
  int local = 0;
  for (int k = 0; k < per; ++k) { const int c = tid * per + k; if (c < ncell) local += ccount[c]; }
  int ex = block_exclusive_scan(local, warp_sums, sh_total);
  for (int k = 0; k < per; ++k) {
    const int c = tid * per + k;
    if (c < ncell) { cstart[c] = ex; ex += ccount[c]; ccount[c] = 0; }
  }
  if (tid == 0) cstart[ncell] = n;
  __syncthreads();
  for (int i = tid; i < n; i += MD_THREADS) {
    const unsigned lo = (unsigned)keys[i];
    const int c = ((lo >> 16) / cell) * gw + (lo & 0xffff) / cell;
    items[cstart[c] + atomicAdd(&ccount[c], 1)] = (ItemT)i;
    state[i] = 0;
  }
  __syncthreads();
  volatile unsigned char* vstate = state;
  for (;;) {
    int pending = 0;
    for (int i = tid; i < n; i += MD_THREADS) {
      if (vstate[i] != 0) continue;
      const unsigned lo = (unsigned)keys[i];
      const int x = lo & 0xffff, y = lo >> 16;
      const int xc = x / cell, yc = y / cell;
      const int x1 = max(0, xc - 1), x2 = min(gw - 1, xc + 1);
      const int y1 = max(0, yc - 1), y2 = min(gh - 1, yc + 1);
      int verdict = 1;   // 1 accept, 2 reject, 0 wait
      for (int yy = y1; yy <= y2 && verdict != 2; ++yy) {
        const int jb = cstart[yy * gw + x1], je = cstart[yy * gw + x2 + 1];
        for (int j = jb; j < je; ++j) {
          const int rj = (int)items[j];
          if (rj >= i) continue;                    // only stronger corners (smaller rank) matter
          const unsigned lj = (unsigned)keys[rj];
          const int dx = x - (int)(lj & 0xffff), dy = y - (int)(lj >> 16);
          if ((double)(dx * dx + dy * dy) < d2) {
            const int sj = vstate[rj];
            if (sj == 1) { verdict = 2; break; }
            if (sj == 0) verdict = 0;
          }
        }
      }
      if (verdict == 0) pending = 1;
      else vstate[i] = (unsigned char)verdict;
    }
    if (!__syncthreads_or(pending)) break;
  }
}

// bitonic sort of np2 (power of two) u64 keys, descending
__device__ void bitonic_desc(unsigned long long* a, int np2) {
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < np2; i += MD_THREADS) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long x = a[i], y = a[l];
          const bool desc = (i & k) == 0;
          if (desc ? (x < y) : (x > y)) { a[i] = y; a[l] = x; }
        }
      }
      __syncthreads();
    }
  }
}

// write the first `max_corners` accepted keys (rank order) as corners; returns count via ncorners
template <typename KeyT>
__device__ void emit_accepted(const KeyT* keys, const unsigned char* state, int n, bool all,
                              int max_corners, float* corners, int corner_stride, int* ncorners_s, int* warp_sums,
                              int* sh_total) {
  const int tid = threadIdx.x;
  const int per = (n + MD_THREADS - 1) / MD_THREADS;
  int local = 0;
  for (int k = 0; k < per; ++k) { const int i = tid * per + k; if (i < n && (all || state[i] == 1)) ++local; }
  int pos = block_exclusive_scan(local, warp_sums, sh_total);
  int limit = max_corners < corner_stride ? max_corners : corner_stride;
  for (int k = 0; k < per; ++k) {
    const int i = tid * per + k;
    if (i < n && (all || state[i] == 1)) {
      if (pos < limit) {
        const unsigned lo = (unsigned)keys[i];
        corners[2 * pos] = (float)(lo & 0xffff);
        corners[2 * pos + 1] = (float)(lo >> 16);
      }
      ++pos;
    }
  }
  if (tid == 0) *ncorners_s = *sh_total < limit ? *sh_total : limit;
}

__global__ void __launch_bounds__(MD_THREADS, 1)
mindist_fast_kernel(const unsigned long long* __restrict__ cand_base, const int* __restrict__ ncand, int cand_cap,
                    const int* __restrict__ eigmax, double quality, int w, int h, double min_distance,
                    int max_corners, float* __restrict__ corners_base, int* __restrict__ ncorners,
                    int corner_stride, int* __restrict__ flags, int* __restrict__ need_full, int max_cells) {
  extern __shared__ unsigned char smem_raw[];
  // region A [0,128K): u64 keys[SEL_CAP] while selecting + sorting; afterwards pos u32[SEL_CAP] (low halves, rank order)
  //                    in its first half and cstart / ccount in its second half.
  // region B [128K, 176K): items u16[SEL_CAP] | state u8[SEL_CAP];  then hist int[HBINS].
  unsigned long long* keys = (unsigned long long*)smem_raw;
  unsigned* pos = (unsigned*)smem_raw;
  int* cstart = (int*)(smem_raw + (size_t)SEL_CAP * 4);
  int* ccount = cstart + (max_cells + 1);
  unsigned char* regB = smem_raw + (size_t)SEL_CAP * 8;
  int* hist = (int*)(regB + (size_t)SEL_CAP * 3);          // own 16 KB: survives the retries
  unsigned short* items = (unsigned short*)regB;
  unsigned char* state = regB + (size_t)SEL_CAP * 2;
  __shared__ int warp_sums[32];
  __shared__ int sh_total, sh_nsel, sh_cut, sh_npass;

  const int s = blockIdx.x, tid = threadIdx.x;
  const unsigned long long* cand = cand_base + (size_t)s * cand_cap;
  int n = ncand[s];
  if (n > cand_cap) { n = cand_cap; if (tid == 0) atomicOr(&flags[s], 1); }
  float* corners = corners_base + (size_t)s * corner_stride * 2;
  const float thr = (float)((double)__int_as_float(eigmax[s]) * quality);
  if (tid == 0) { need_full[s] = 0; sh_nsel = 0; sh_npass = 0; }
  // 1. histogram of the candidates that pass the quality threshold
  for (int i = tid; i < HBINS; i += MD_THREADS) hist[i] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += MD_THREADS) {
    const unsigned vb = (unsigned)(cand[i] >> 32);
    if (__uint_as_float(vb) > thr) atomicAdd(&hist[(vb >> 19) & (HBINS - 1)], 1);
  }
  __syncthreads();
  // Adaptive prefix: the greedy result only depends on stronger corners, so the strongest `cap` candidates decide the
  // first accepted corners exactly; start with 4*N and double while fewer than N corners come out.
  int cap = 4 * max_corners < 2048 ? 2048 : 4 * max_corners;
  if (cap > SEL_CAP) cap = SEL_CAP;
  bool decided = false;
  for (;;) {
    // 2. lowest bin `cut` such that everything in bins >= cut fits the fast path (suffix sums by one warp)
    if (tid < 32) {
      int run = 0, cut = HBINS, npass = 0;
      for (int b0 = HBINS - 32; b0 >= 0; b0 -= 32) {
        const int v = hist[b0 + tid];
        int suf = v;                                        // inclusive suffix sum within the 32-bin chunk
  #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_down_sync(FULL, suf, o);
          if (tid + o < 32) suf += t;
        }
        const bool fits = run + suf <= cap;
        const unsigned fm = __ballot_sync(FULL, fits);
        if (cut == b0 + 32) {                               // still contiguous from the top
          // lanes that fit form a suffix of the chunk (suffix sums are monotone)
          const int nfit = __popc(fm);
          if (nfit > 0) cut = b0 + 32 - nfit;
        }
        run += __shfl_sync(FULL, suf, 0);
        npass = run;
      }
      if (tid == 0) { sh_cut = cut; sh_npass = npass; }
    }
    __syncthreads();
    const int cut = sh_cut, npass = sh_npass;
    if (tid == 0) sh_nsel = 0;
    __syncthreads();
    // 3. gather the selected prefix into shared memory
    for (int i = tid; i < n; i += MD_THREADS) {
      const unsigned long long k = cand[i];
      const unsigned vb = (unsigned)(k >> 32);
      if (__uint_as_float(vb) > thr && (int)((vb >> 19) & (HBINS - 1)) >= cut) {
        const int p = atomicAdd(&sh_nsel, 1);
        if (p < SEL_CAP) keys[p] = k;
      }
    }
    __syncthreads();
    const int nsel = sh_nsel < SEL_CAP ? sh_nsel : SEL_CAP;
    int np2 = 1;
    while (np2 < nsel) np2 <<= 1;
    for (int i = nsel + tid; i < np2; i += MD_THREADS) keys[i] = 0ull;
    __syncthreads();
    bitonic_desc(keys, np2);
    const bool complete = nsel == npass;
    {   // compact the sorted keys to their 32-bit positions in place (registers in between: the arrays alias)
      constexpr int PER = SEL_CAP / MD_THREADS;
      unsigned lo[PER];
  #pragma unroll
      for (int k = 0; k < PER; ++k) { const int i = tid + k * MD_THREADS; lo[k] = i < nsel ? (unsigned)keys[i] : 0u; }
      __syncthreads();
  #pragma unroll
      for (int k = 0; k < PER; ++k) { const int i = tid + k * MD_THREADS; if (i < nsel) pos[i] = lo[k]; }
      __syncthreads();
    }
    if (min_distance >= 1.0) {
      const int cell = (int)rint(min_distance);
      const int gw = (w + cell - 1) / cell, gh = (h + cell - 1) / cell;
      greedy_fixed_point<unsigned, unsigned short>(pos, nsel, cell, gw, gh, min_distance * min_distance, cstart, ccount,
                                                   items, state, warp_sums, &sh_total);
      emit_accepted<unsigned>(pos, state, nsel, false, max_corners, corners, corner_stride, &ncorners[s], warp_sums, &sh_total);
    } else {
      emit_accepted<unsigned>(pos, state, nsel, true, max_corners, corners, corner_stride, &ncorners[s], warp_sums, &sh_total);
    }
    __syncthreads();
    decided = complete || sh_total >= max_corners || min_distance < 1.0;
    if (decided || cap >= SEL_CAP) break;
    cap = cap * 2 > SEL_CAP ? SEL_CAP : cap * 2;
    __syncthreads();
  }
  if (tid == 0 && !decided) need_full[s] = 1;   // prefix too short: slow path decides
}

// slow path: all passing candidates, keys sorted in global memory by a shared-memory-free odd route:
// (1) compact passing keys into `sorted`, (2) bitonic sort in global memory, (3) greedy with global state.
__global__ void __launch_bounds__(MD_THREADS, 1)
mindist_full_kernel(const unsigned long long* __restrict__ cand_base, const int* __restrict__ ncand, int cand_cap,
                    unsigned long long* __restrict__ sorted_base, int* __restrict__ items_base,
                    unsigned char* __restrict__ state_base, const int* __restrict__ eigmax, double quality, int w,
                    int h, double min_distance, int max_corners, float* __restrict__ corners_base,
                    int* __restrict__ ncorners, int corner_stride, const int* __restrict__ need_full, int max_cells) {
  extern __shared__ unsigned char smem_raw[];
  const int s = blockIdx.x, tid = threadIdx.x;
  if (!need_full[s]) return;
  int* cstart = (int*)smem_raw;
  int* ccount = cstart + (max_cells + 1);
  __shared__ int warp_sums[32];
  __shared__ int sh_total, sh_n;
  const unsigned long long* cand = cand_base + (size_t)s * cand_cap;
  unsigned long long* keys = sorted_base + (size_t)s * cand_cap;
  int* items = items_base + (size_t)s * cand_cap;
  unsigned char* state = state_base + (size_t)s * cand_cap;
  int n = ncand[s];
  if (n > cand_cap) n = cand_cap;
  const float thr = (float)((double)__int_as_float(eigmax[s]) * quality);
  if (tid == 0) sh_n = 0;
  __syncthreads();
  for (int i = tid; i < n; i += MD_THREADS) {
    const unsigned long long k = cand[i];
    if (__uint_as_float((unsigned)(k >> 32)) > thr) keys[atomicAdd(&sh_n, 1)] = k;
  }
  __syncthreads();
  const int m = sh_n;
  int np2 = 1;
  while (np2 < m) np2 <<= 1;                     // np2 <= cand_cap (power of two by construction)
  for (int i = m + tid; i < np2; i += MD_THREADS) keys[i] = 0ull;
  __syncthreads();
  bitonic_desc(keys, np2);
  float* corners = corners_base + (size_t)s * corner_stride * 2;
  if (min_distance >= 1.0) {
    const int cell = (int)rint(min_distance);
    const int gw = (w + cell - 1) / cell, gh = (h + cell - 1) / cell;
    greedy_fixed_point<unsigned long long, int>(keys, m, cell, gw, gh, min_distance * min_distance, cstart, ccount, items, state,
                            warp_sums, &sh_total);
    emit_accepted<unsigned long long>(keys, state, m, false, max_corners, corners, corner_stride, &ncorners[s], warp_sums, &sh_total);
  } else {
    emit_accepted<unsigned long long>(keys, state, m, true, max_corners, corners, corner_stride, &ncorners[s], warp_sums, &sh_total);
  }
}

__global__ void gftt_reset_kernel(int* eigmax, int* ncand, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { eigmax[i] = 0; ncand[i] = 0; }
}

size_t fast_smem(int max_cells) {
  (void)max_cells;                       // cstart/ccount live inside region A: needs 2*max_cells+1 ints <= SEL_CAP ints
  return (size_t)SEL_CAP * 8 + (size_t)SEL_CAP * 2 + SEL_CAP + (size_t)HBINS * 4;
}
size_t full_smem(int max_cells) { return (size_t)(2 * max_cells + 1) * 4; }

}  // namespace

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
  const int nsx = (w + CS_COLS - 1) / CS_COLS, nsy = (h + CS_ROWS - 1) / CS_ROWS;
  dim3 g1((nsx * nsy + CS_WARPS - 1) / CS_WARPS, n_streams);
  corner_response_kernel<<<g1, CS_WARPS * 32, 0, ctx->stream>>>(
      ctx->pyr[slot] + ctx->geom.lv[0].off, ctx->geom.stream_stride, ctx->geom.lv[0].pitch, w, h, nsx, nsy,
      ctx->keep_eig ? ctx->d_eig : nullptr, ctx->d_eigmax, quality, ctx->d_cand, ctx->d_ncand, ctx->cand_cap);
  mindist_fast_kernel<<<n_streams, MD_THREADS, fast_smem(ctx->max_cells), ctx->stream>>>(
      ctx->d_cand, ctx->d_ncand, ctx->cand_cap, ctx->d_eigmax, quality, w, h, min_distance, max_corners,
      ctx->d_corners, ctx->d_ncorners, ctx->gftt_cap, ctx->d_flags, ctx->d_need_full, ctx->max_cells);
  mindist_full_kernel<<<n_streams, MD_THREADS, full_smem(ctx->max_cells), ctx->stream>>>(
      ctx->d_cand, ctx->d_ncand, ctx->cand_cap, ctx->d_sorted, ctx->d_items, ctx->d_state, ctx->d_eigmax, quality,
      w, h, min_distance, max_corners, ctx->d_corners, ctx->d_ncorners, ctx->gftt_cap, ctx->d_need_full,
      ctx->max_cells);
  ctx->launches += 4;
  FLV_CUDA(ctx, cudaGetLastError());
  return FLV_OK;
}

int flv_gftt_init(flv_ctx* ctx) {
  FLV_CUDA(ctx, cudaFuncSetAttribute(mindist_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)fast_smem(ctx->max_cells)));
  FLV_CUDA(ctx, cudaFuncSetAttribute(mindist_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)full_smem(ctx->max_cells)));
  return FLV_OK;
}
