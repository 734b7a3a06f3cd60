// K5 -- FeatureDEM region-balanced corner selection, one CTA per stream, one warp per region.
//
// Reference: FeatureDEM::detect (src/processing/feature_dem.cpp:215-266), FeatureDEM::redetect
// (:124-213), calHarrisR (:59-88), fillIntoRegion (:92-121).  Reference quirks reproduced on
// purpose: patch[5] reads (x+1,y+1); IX,IY use integer division by 3;
// Y2 = IY*IX and XY = IX*IX; the spacing test is `dis_x <= bd || dis_y <= bd`.
//
// Ordering: the reference sorts each region with std::sort (libstdc++ introsort, unstable) on the
// coarse pseudo-Harris score, so ties are frequent and their order decides which corners survive.
// introsort_desc() below restates libstdc++'s algorithm (median-of-3 introsort loop with threshold
// 16 + final insertion sort) so the permutation -- ties included -- is the one the reference gets.
#include "ctx.h"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int RG_THREADS = FLV_NUM_REGIONS * 32;
constexpr int KCAP = 256;      // kept points per region (existing + new)

struct Item { float score; int idx; };   // idx = position in the GFTT list

__device__ __forceinline__ bool cmp_desc(const Item& a, const Item& b) { return a.score > b.score; }
__device__ __forceinline__ void iswap(Item& a, Item& b) { Item t = a; a = b; b = t; }

// libstdc++ std::__insertion_sort / __unguarded_linear_insert
__device__ void unguarded_linear_insert(Item* last) {
  Item val = *last;
  Item* next = last - 1;
  while (cmp_desc(val, *next)) { *last = *next; last = next; --next; }
  *last = val;
}
__device__ void insertion_sort(Item* first, Item* last) {
  if (first == last) return;
  for (Item* i = first + 1; i != last; ++i) {
    if (cmp_desc(*i, *first)) {
      Item val = *i;
      for (Item* p = i; p != first; --p) *p = *(p - 1);
      *first = val;
    } else {
      unguarded_linear_insert(i);
    }
  }
}
// std::__move_median_to_first + std::__unguarded_partition
__device__ Item* partition_pivot(Item* first, Item* last) {
  Item* mid = first + (last - first) / 2;
  Item *a = first + 1, *b = mid, *c = last - 1;
  if (cmp_desc(*a, *b)) {
    if (cmp_desc(*b, *c)) iswap(*first, *b);
    else if (cmp_desc(*a, *c)) iswap(*first, *c);
    else iswap(*first, *a);
  } else if (cmp_desc(*a, *c)) iswap(*first, *a);
  else if (cmp_desc(*b, *c)) iswap(*first, *c);
  else iswap(*first, *b);
  Item* lo = first + 1;
  Item* hi = last;
  for (;;) {
    while (cmp_desc(*lo, *first)) ++lo;
    --hi;
    while (cmp_desc(*first, *hi)) --hi;
    if (!(lo < hi)) return lo;
    iswap(*lo, *hi);
    ++lo;
  }
}
// std::sort(first, last, greater-by-score).  Returns false if the depth limit was hit (libstdc++
// would switch to heapsort there; never seen for these sizes -- reported as an overflow flag).
__device__ bool introsort_desc(Item* first, int n) {
  if (n <= 1) return true;
  int lg = 31 - __clz(n);
  struct Frame { Item* f; Item* l; int depth; };
  Frame stack[48];
  int sp = 0;
  stack[sp++] = Frame{first, first + n, 2 * lg};
  bool ok = true;
  while (sp > 0) {
    Frame fr = stack[--sp];
    Item* f = fr.f; Item* l = fr.l; int depth = fr.depth;
    // __introsort_loop: recurse on the right part, iterate on the left part.  The order in which
    // disjoint sub-ranges are processed does not change the result.
    while (l - f > 16) {
      if (depth == 0) { ok = false; break; }
      --depth;
      Item* cut = partition_pivot(f, l);
      if (sp < 48) stack[sp++] = Frame{cut, l, depth}; else ok = false;
      l = cut;
    }
  }
  // __final_insertion_sort
  if (n > 16) {
    insertion_sort(first, first + 16);
    for (Item* i = first + 16; i != first + n; ++i) unguarded_linear_insert(i);
  } else {
    insertion_sort(first, first + n);
  }
  return ok;
}

__device__ __forceinline__ float harris_score(const uint8_t* img, int pitch, int xx, int yy) {
  // feature_dem.cpp:63-87 verbatim semantics (including the patch[5] and Y2/XY slips)
  int p0 = img[(size_t)(yy - 1) * pitch + xx - 1], p1 = img[(size_t)(yy - 1) * pitch + xx],
      p2 = img[(size_t)(yy - 1) * pitch + xx + 1];
  int p3 = img[(size_t)yy * pitch + xx - 1];
  int p5 = img[(size_t)(yy + 1) * pitch + xx + 1];
  int p6 = img[(size_t)(yy + 1) * pitch + xx - 1], p7 = img[(size_t)(yy + 1) * pitch + xx],
      p8 = img[(size_t)(yy + 1) * pitch + xx + 1];
  float IX = (float)((p0 + p3 + p6 - (p2 + p5 + p8)) / 3);
  float IY = (float)((p0 + p1 + p2 - (p6 + p7 + p8)) / 3);
  float X2 = __fmul_rn(IX, IX), Y2 = __fmul_rn(IY, IX), XY = __fmul_rn(IX, IX);
  float t = __fadd_rn(X2, Y2);
  // R = (X2*Y2) - (XY*XY) - 0.05f*(X2+Y2)*(X2+Y2), evaluated left to right in float
  float r = __fsub_rn(__fmul_rn(X2, Y2), __fmul_rn(XY, XY));
  return __fsub_rn(r, __fmul_rn(__fmul_rn(0.05f, t), t));
}

__global__ void __launch_bounds__(RG_THREADS, 1)
region_select_kernel(const uint8_t* __restrict__ img_base, size_t stream_stride, int pitch, int w,
                     int h, const float* __restrict__ corners_base, const int* __restrict__ ncorners,
                     int corner_stride, const double* __restrict__ exist_base,
                     const int* __restrict__ nexist, int max_pts, int redetect, int max_region,
                     int boundary_dis, float* __restrict__ new_base, int* __restrict__ nnew,
                     int* __restrict__ flags) {
  extern __shared__ unsigned char smem_raw[];
  Item* items = (Item*)smem_raw;                         // [corner_stride] region-sorted pool
  unsigned char* reg_of = (unsigned char*)(items + corner_stride);   // [corner_stride] region id or 255
  __shared__ float2 kept[FLV_NUM_REGIONS][KCAP];
  __shared__ int kept_n[FLV_NUM_REGIONS], kept_first_new[FLV_NUM_REGIONS];
  __shared__ int reg_cnt[FLV_NUM_REGIONS], reg_off[FLV_NUM_REGIONS + 1], out_off[FLV_NUM_REGIONS + 1];

  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, reg = tid >> 5;
  const uint8_t* img = img_base + (size_t)s * stream_stride;
  const float* corners = corners_base + (size_t)s * corner_stride * 2;
  const int n = ncorners[s];
  const int rw = (int)floor(w / 4.0), rh = (int)floor(h / 4.0);
  float* out = new_base + (size_t)s * max_pts * 2;

  // region id of every GFTT corner (fillIntoRegion, existed_features=false)
  for (int i = tid; i < n; i += RG_THREADS) {
    float x = corners[2 * i], y = corners[2 * i + 1];
    int r = 255;
    if (x >= 3.f && x < (float)(w - 3) && y >= 3.f && y < (float)(h - 3)) {
      r = (int)(__fadd_rn(__fmul_rn(4.f, floorf(y / (float)rh)), x / (float)rw));
      if (r < 0 || r > 15) r = 255;
    }
    reg_of[i] = (unsigned char)r;
  }
  if (tid < FLV_NUM_REGIONS) { reg_cnt[tid] = 0; kept_n[tid] = 0; }
  __syncthreads();
  // count per region (warp `reg` counts its own), then offsets
  {
    int c = 0;
    for (int i = lane; i < n; i += 32) c += (reg_of[i] == reg);
    c = __reduce_add_sync(FULL, c);
    if (lane == 0) reg_cnt[reg] = c;
  }
  __syncthreads();
  if (tid == 0) {
    int a = 0;
    for (int r = 0; r < FLV_NUM_REGIONS; ++r) { reg_off[r] = a; a += reg_cnt[r]; }
    reg_off[FLV_NUM_REGIONS] = a;
  }
  __syncthreads();
  Item* mine = items + reg_off[reg];
  const int cnt = reg_cnt[reg];
  // ordered compaction (GFTT order is preserved, as push_back does) + score
  {
    int base = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
      int i = i0 + lane;
      bool in = i < n && reg_of[i] == reg;
      unsigned m = __ballot_sync(FULL, in);
      if (in) {
        int pos = base + __popc(m & ((1u << lane) - 1));
        int xx = (int)corners[2 * i], yy = (int)corners[2 * i + 1];
        mine[pos].score = harris_score(img, pitch, xx, yy);
        mine[pos].idx = i;
      }
      base += __popc(m);
    }
  }
  // existing features seed the region (redetect, fillIntoRegion existed_features=true)
  if (redetect) {
    const double* ex = exist_base + (size_t)s * max_pts * 2;
    const int ne = nexist[s];
    int base = 0;
    for (int i0 = 0; i0 < ne; i0 += 32) {
      int i = i0 + lane;
      bool in = false;
      float x = 0.f, y = 0.f;
      if (i < ne) {
        x = (float)ex[2 * i]; y = (float)ex[2 * i + 1];
        if (x >= 3.f && x < (float)(w - 3) && y >= 3.f && y < (float)(h - 3)) {
          int r = (int)(__fadd_rn(__fmul_rn(4.f, floorf(y / (float)rh)), x / (float)rw));
          in = (r == reg);
        }
      }
      unsigned m = __ballot_sync(FULL, in);
      if (in) {
        int pos = base + __popc(m & ((1u << lane) - 1));
        if (pos < KCAP) kept[reg][pos] = make_float2(x, y);
      }
      base += __popc(m);
    }
    if (base > KCAP) { base = KCAP; if (lane == 0) atomicOr(&flags[s], 4); }
    if (lane == 0) kept_n[reg] = base;
  }
  __syncwarp();
  if (lane == 0) {
    kept_first_new[reg] = kept_n[reg];
    if (!introsort_desc(mine, cnt)) atomicOr(&flags[s], 8);
  }
  __syncwarp();
  // greedy pick in sorted order; lanes test the candidate against the kept list
  {
    int kn = kept_n[reg];
    const float bd = (float)boundary_dis;
    int added = 0;
    bool full = false;
    for (int j = 0; j < cnt && !full; ++j) {
      const int ci = mine[j].idx;
      // detect keeps Point2f, redetect converts to cv::Point (int) first; GFTT coords are integral
      const float cx = corners[2 * ci], cy = corners[2 * ci + 1];
      bool conflict = false;
      for (int k = lane; k < kn; k += 32) {
        float dx = fabsf(__fsub_rn(cx, kept[reg][k].x)), dy = fabsf(__fsub_rn(cy, kept[reg][k].y));
        conflict |= (dx <= bd) || (dy <= bd);
      }
      conflict = __any_sync(FULL, conflict);
      if (!conflict) {
        if (kn < KCAP) { if (lane == 0) kept[reg][kn] = make_float2(cx, cy); }
        else if (lane == 0) atomicOr(&flags[s], 4);
        kn = min(kn + 1, KCAP);
        ++added;
        __syncwarp();
        if (redetect) full = kn >= max_region;      // regionKeyPts[i].size() >= max (feature_dem.cpp:188)
        else full = added >= max_region;            // count >= max (feature_dem.cpp:250)
      }
    }
    if (lane == 0) kept_n[reg] = kn;
  }
  __syncthreads();
  if (tid == 0) {
    int a = 0;
    for (int r = 0; r < FLV_NUM_REGIONS; ++r) { out_off[r] = a; a += kept_n[r] - kept_first_new[r]; }
    out_off[FLV_NUM_REGIONS] = a;
    nnew[s] = min(a, max_pts);
    if (a > max_pts) atomicOr(&flags[s], 16);
  }
  __syncthreads();
  {
    const int f = kept_first_new[reg], m = kept_n[reg] - f, o = out_off[reg];
    for (int k = lane; k < m; k += 32) {
      if (o + k < max_pts) { out[2 * (o + k)] = kept[reg][f + k].x; out[2 * (o + k) + 1] = kept[reg][f + k].y; }
    }
  }
}

}  // namespace

int flv_launch_region(flv_ctx* ctx, int slot, int n_streams, const flv_feature_params* prm,
                      int redetect) {
  size_t smem = (size_t)ctx->gftt_cap * (sizeof(Item) + 1);
  if (!ctx->attr_region) {          // per context (= per device): function attributes do not carry across devices
    FLV_CUDA(ctx, cudaFuncSetAttribute(region_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctx->attr_region = 1;
  }
  region_select_kernel<<<n_streams, RG_THREADS, smem, ctx->stream>>>(
      ctx->pyr[slot] + ctx->geom.lv[0].off, ctx->geom.stream_stride, ctx->geom.lv[0].pitch, ctx->w,
      ctx->h, ctx->d_corners, ctx->d_ncorners, ctx->gftt_cap, ctx->d_exist, ctx->d_nexist,
      ctx->max_pts, redetect, prm->max_region_feature_num, prm->boundary_dis, ctx->d_newxy,
      ctx->d_nnew, ctx->d_flags);
  ctx->launches++;
  FLV_CUDA(ctx, cudaGetLastError());
  return FLV_OK;
}
