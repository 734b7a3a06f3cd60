// K7-K10 -- sliding-window bundle adjustment, whole Levenberg-Marquardt loop on the device.
// One persistent CTA per stream (= one independent window / one g2o SparseOptimizer); all streams
// of a batch run concurrently on different SMs.  No atomics on floating-point data: every sum has a
// fixed order, so results are run-to-run deterministic.
//
// Reference (paths under 3rdPartLib/g2o/g2o/ unless noted; restated in oracle/ba_ref.c):
//   residual / Jacobians   types/sba/types_six_dof_expmap.h:209-214, types_six_dof_expmap.cpp:389-433
//   Huber + quadratic form core/robust_kernel_impl.cpp:65-78, core/base_binary_edge.hpp:62-134
//   Schur + back-subst     core/block_solver.hpp:328-447;  lambda handling :525-565
//   LM control             core/optimization_algorithm_levenberg.cpp:58-175
//   outer loop / chi2      core/sparse_optimizer.cpp:366-430, :102-116;  active sets :168-272
//   callers                src/backend/vo_localmap.cpp:292-319 (12, cull chi2>3, 8),
//                          src/processing/optimize_in_frame.cpp:64-80 (2, cull, <10 edges => fail, 2)
//
// Data layout (per stream, fp64; global workspace stays L2-resident, S/b/x + one landmark chunk in shared memory):
//   active edges get SLOTS ordered (landmark chunk, pose, landmark) and CSR positions ordered (landmark, pose).  Per
//   edge only the pixel measurement is kept (uvs by slot, luv by CSR position, with csr_p / csr_l); NO Jacobian product
//   is stored: every pass of this kernel waits on memory, not on the fp64 pipe, so W = rho' B^T A is recomputed from the
//   state wherever it is needed (the state equals the linearisation point in all those places).  Per-landmark
//   quantities are planes indexed by landmark (Hll[6][ML], bl[3][ML], Dinv[6][ML], Ld[6][ML], Dv[3][ML]).
//   tab[p][l] = slot of edge (p,l); lmask[l] = poses observing l (P <= 32); Hd[pi][21], bp[6 pi] in shared memory.
// Passes per LM iteration / trial:
//   build   thread per landmark over its CSR entries (point Jacobians -> Hll, bl); warp per (pose, part) (Hpp, bp)
//   Schur   per landmark chunk: stage Z = W chol((Hll + lambda)^-1) and Ld^T bl in shared memory, then warp per pose pair
//           (tasks drawn dynamically, largest first), ONE lane per member landmark with 36 accumulators, butterfly
//           reduce-scatter, one writer per block of S
//   solve   right-looking LDL^T of the reduced camera system in shared memory (n = 6*(free poses) <= 144), one barrier per
//           column; back substitution inside one warp
//   update  thread per CSR entry (W^T x_p into the idle chunk area), thread per landmark (back-substitution), thread per
//           pose (exp map), chi2 block reduction
// Algorithmic bytes: build 168E+392P+120L, Schur 144E+96L+288Pf^2+48Pf per trial.
#include "ctx.h"
#include "ba_math.h"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

constexpr int BA_THREADS = 384;          // 12 warps: 168 registers per thread keep the 36 Schur accumulators of a lane resident (512 threads x 128 regs measured slower: spills in the pair products)
constexpr int BA_WARPS = BA_THREADS / 32;
constexpr int BA_MAX_POSES = 32;
constexpr int BA_MAX_FREE = 24;          // reduced system n <= 144 -> S fits shared memory
constexpr unsigned FULL = 0xffffffffu;
constexpr int BA_TRACE_ITERS = 32;       // per-iteration debug trace rows kept per stream (flv_ba_trace)

struct BAArgs {
  const flv_ba_problem* problems;
  flv_ba_params prm;
  double* poses; double* lms;
  const int* ep; const int* el; const double* uv; uint8_t* active;
  flv_ba_stats* stats;
  int max_poses, max_lms, max_edges;
  unsigned char* ws; size_t ws_stride;
  long long* prof;   // [S][8] cycle counters or nullptr
  double* trace;     // [S][BA_TRACE_ITERS][4] per LM iteration: chi2 at the end, lambda, rho of the last trial, trials
  int dyn_doubles;   // dynamic shared memory of the launch, in doubles
  int member_buf;    // ints of shared memory for TMA-staged pose-pair member lists (0 = read them from L2; see BA_MEMBER_BUF)
  int ws_poses;      // pose capacity the workspace layout was computed for (min(max_poses, BA_MAX_POSES))
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
// deterministic block sum; result valid in all threads.  red: shared double[BA_WARPS]
__device__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0;
#pragma unroll
  for (int i = 0; i < BA_WARPS; ++i) t += red[i];
  return t;
}
__device__ double block_max(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0;
#pragma unroll
  for (int i = 0; i < BA_WARPS; ++i) t = fmax(t, red[i]);
  return t;
}


constexpr int BA_MAX_PAIRS = BA_MAX_FREE * (BA_MAX_FREE + 1) / 2;   // 300
constexpr int BA_MAX_CHUNKS = 128;
constexpr int BA_CHUNK_PLANES = 18 + 3;       // Z = W Ld (6x3) per slot, Ld^T bl per landmark
constexpr int BA_MAX_CAP = 1023;              // chunk-local positions are packed 10 bits each
// ints of shared memory for TMA-staged member lists (FLV_BA_TMA=1).  Off by default: measured on B200 (tools/ba_profile.py,
// profiles/README.md) the staged lists are 4 % slower at W=10 / 4 CTAs (the L2 reads were already prefetched one round ahead)
// and 13 % slower at W=20 / 1 CTA (the buffer shrinks the landmark chunks).
constexpr int BA_MEMBER_BUF = 6144;

struct Sh {   // fixed-size shared state (one copy per CTA of the window's cluster)
  double red[BA_WARPS];
  double Hd[BA_MAX_FREE][21];
  double part[2 * BA_MAX_FREE][27];     // pose-pass partial sums
  double Hdp[BA_MAX_FREE][27];          // this CTA's share of (Hpp, bp) over its own chunks; rank 0 sums the cluster's shares
  double bp[6 * BA_MAX_FREE];
  double x[6 * BA_MAX_FREE];
  double invd[6 * BA_MAX_FREE];
  double col1[6 * BA_MAX_FREE + 2];     // Cholesky: final entries of the odd column of a two-column step (written back a step later)
  double poses[7 * BA_MAX_POSES];       // the window's poses: every CTA of the cluster keeps (and updates) its own identical copy
  double pbk[7 * BA_MAX_POSES];         // backup of the poses for a rejected trial
  unsigned long long mbar;              // mbarrier of the bulk (TMA) copies that stage a chunk's pose-pair member lists
  double xchg[2][4];                    // cluster all-reduce exchange slots (double-buffered: one cluster barrier per reduction)
  int rank, C;                          // rank in / size of the window's cluster
  // per-CTA state of the member-list staging (NOT part of the range copied from rank 0 below)
  int mb_chunk;                          // what the member staging buffer holds: chunk index, -2 = all chunks of this CTA, -1 = nothing
  int m_base;                            // first (16-byte aligned) index of `pairs` held in the buffer
  unsigned m_issued;                     // bulk copies issued so far (a thread waits until it has seen as many completions)
  int c0, c1, l0, l1, s0, s1;           // this CTA's chunks and the landmark / slot (= CSR) ranges they cover
  // ---- written by setup_active on rank 0 and copied to the other CTAs (contiguous: pidx .. capq) ----
  int pidx[BA_MAX_POSES];
  int pose_of[BA_MAX_FREE];
  int pcount[BA_MAX_POSES];
  int chunk_lb[BA_MAX_CHUNKS + 1];      // landmark range of a chunk
  int chunk_sb[BA_MAX_CHUNKS + 1];      // slot range of a chunk
  int scan[BA_WARPS];
  unsigned char pair_a[BA_MAX_PAIRS], pair_b[BA_MAX_PAIRS];
  int task_bounds[BA_MAX_PAIRS + 1];    // member-list bounds of the pose pairs for the chunk being processed
  unsigned short task_order[BA_MAX_PAIRS];   // pose pairs by decreasing member count: the order warps pick them up in
  int next_task;
  int np, fail, nact, overflow, nch, mb, cap, capq;     // mb: ints of the member-list staging buffer behind the chunk planes (0 = none)
  long long prof[16], tlast;  // cycle counters: 0 chi2, 1 build (pose pass), 2 schur (pair products), 3 cholesky, 4 substitution,
                              // 5 update, 6 setup, 7 -, 8 schur init + Dinv, 9 schur chunk staging, 10 build edge pass, 11 build landmark pass
};

// Cluster-wide reductions of up to 4 block-uniform values: every CTA publishes its share in its own shared memory, one cluster
// barrier, then every CTA sums the shares in rank order through distributed shared memory -- identical result in all CTAs, so
// the LM control flow stays uniform over the cluster.  Slots alternate (xpar), so a CTA that runs ahead into the next
// reduction writes the other slot while a slower one still reads this one.
template <int K, bool MAX>
__device__ __forceinline__ void cluster_reduce(double (&v)[K], Sh& sh, int& xpar) {
  if (sh.C == 1) return;
  cg::cluster_group cl = cg::this_cluster();
  double* slot = sh.xchg[xpar];
  xpar ^= 1;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < K; ++i) slot[i] = v[i];
  }
  cl.sync();
#pragma unroll
  for (int i = 0; i < K; ++i) v[i] = 0;
  for (int r = 0; r < sh.C; ++r) {
    const double* rs = cl.map_shared_rank(slot, r);
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] = MAX ? fmax(v[i], rs[i]) : v[i] + rs[i];
  }
}

// ---- TMA bulk copy global -> shared memory, completion on an mbarrier (cp.async.bulk; SASS: UBLKCP + SYNCS) ----------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// one thread: announce `bytes`, then start the copy (src, dst 16-byte aligned, bytes a multiple of 16)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // earlier generic-proxy reads of dst are done (after a barrier)
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void mark(Sh& sh, int slot) {
  if (threadIdx.x == 0) { const long long now = clock64(); sh.prof[slot] += now - sh.tlast; sh.tlast = now; }
}

// per-stream global workspace views.  "slot" arrays are structure-of-arrays planes with stride ME (max edges),
// landmark arrays planes with stride ML: consecutive threads touch consecutive addresses in every pass.
// plane PAIRS: two consecutive planes are interleaved per slot so that every access is one 16-byte load / store
#define PL2(p) reinterpret_cast<double2*>(p)
#define CPL2(p) reinterpret_cast<const double2*>(p)

struct Ws {
  double *pbk, *lbk;
  double *luv, *uvs;                        // double2 [ME]: pixel measurement per slot (uvs) and per landmark-major CSR position (luv)
  double *Hll, *bl, *Dinv, *Dv, *Ld;        // [6|3|6|3|6][ML]
  int *tab;                                 // [P][L]: edge id during setup, then slot of edge (p,l) or -1
  unsigned* lmask;                          // [L]
  int *slot_e, *slot_pl, *slot_lp, *csr_p, *csr_l;   // [ME]: edge id, p | l << 8, CSR position of the slot; pose / landmark of a CSR entry
  int *lw, *lstart;                         // [L+1] exclusive prefixes (chunk weights, edge counts)
  int *cp_off;                              // [nch][P+1] slot offsets of (chunk, pose) runs
  int *poff;                                // [nch*nblk + 1] member-list offsets of (chunk, pose pair)
  int *pairs;                               // packed members: pos_a | pos_b << 10 | lpos << 20 (chunk-local)
  int pair_cap, ME, ML;
};

// One thread: stage the member lists of chunks [ch0, ch1) (a contiguous run of `pairs`, widened to 16-byte granules) unless
// the buffer already holds them (tag).  False if they do not fit.  Readers look at sh.mb_chunk / m_base / m_issued after the
// next barrier.
__device__ __noinline__ bool members_stage(int ch0, int ch1, int tag, int nblk, int* Mb, Ws& ws, Sh& sh) {
  const int m0 = ws.poff[ch0 * nblk] & ~3, m1 = (ws.poff[ch1 * nblk] + 3) & ~3;
  if (!(m1 > m0 && m1 - m0 <= sh.mb)) { if (sh.mb_chunk == tag) sh.mb_chunk = -1; return false; }
  if (sh.mb_chunk != tag) {
    bulk_g2s(Mb, ws.pairs + m0, (unsigned)(m1 - m0) * 4u, &sh.mbar);
    sh.m_issued++;
    sh.mb_chunk = tag; sh.m_base = m0;
  }
  return true;
}

__device__ __forceinline__ int sym21(int i, int j) {   // index into upper-triangular 6x6 (i<=j)
  return i * 6 - (i * (i - 1)) / 2 + (j - i);
}

// in-place exclusive prefix sum of a[0..n) (global ints), a[n] = total.  Deterministic; all threads call it.
__device__ void block_exclusive_scan(int* a, int n, Sh& sh) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int seg = (n + BA_THREADS - 1) / BA_THREADS;
  const int lo = min(tid * seg, n), hi = min(lo + seg, n);
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += a[i];
  int inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
  __syncthreads();
  if (lane == 31) sh.scan[warp] = inc;
  __syncthreads();
  int base = 0;
#pragma unroll
  for (int w = 0; w < BA_WARPS; ++w) if (w < warp) base += sh.scan[w];
  int run = base + inc - sum;
  for (int i = lo; i < hi; ++i) { const int v = a[i]; a[i] = run; run += v; }
  if (tid == BA_THREADS - 1) a[n] = base + inc;
  __syncthreads();
}

// chi2 over the edge arrays (used outside the LM loop, where no slot tables are required to be valid)
__device__ double robust_chi2_edges(const flv_ba_problem& pb, const Cam& cam, const double* poses, const double* lms,
                                    const int* ep, const int* el, const double* uv, const uint8_t* act, double delta,
                                    Sh& sh, int& xpar) {
  double acc = 0;
  const double d2 = delta * delta;
  for (int e = threadIdx.x + sh.rank * BA_THREADS; e < pb.n_edges; e += sh.C * BA_THREADS) {
    if (!act[e]) continue;
    double r[2];
    edge_eval<false>(poses + 7 * ep[e], lms + 3 * el[e], uv + 2 * e, cam, r, nullptr, nullptr);
    const double c = r[0] * r[0] + r[1] * r[1];
    acc += (c <= d2) ? c : 2 * sqrt(c) * delta - d2;
  }
  double v[1] = {block_sum(acc, sh.red)};
  cluster_reduce<1, false>(v, sh, xpar);
  return v[0];
}

// chi2 over this CTA's active slots (inside the LM loop): coalesced reads of the slot tables.  Returns the CTA's share;
// callers reduce over the cluster.
__device__ double robust_chi2_part(const Cam& cam, const double* poses, const double* lms, double delta, Ws& ws, Sh& sh) {
  double acc = 0;
  const double d2 = delta * delta;
  for (int s = sh.s0 + threadIdx.x; s < sh.s1; s += BA_THREADS) {
    const int pl = ws.slot_pl[s];
    const double2 q2 = CPL2(ws.uvs)[s];
    const double uv[2] = {q2.x, q2.y};
    double r[2];
    edge_eval<false>(poses + 7 * (pl & 255), lms + 3 * (size_t)(pl >> 8), uv, cam, r, nullptr, nullptr);
    const double c = r[0] * r[0] + r[1] * r[1];
    acc += (c <= d2) ? c : 2 * sqrt(c) * delta - d2;
  }
  return block_sum(acc, sh.red);
}
__device__ double robust_chi2(const Cam& cam, const double* poses, const double* lms, double delta, Ws& ws, Sh& sh, int& xpar) {
  double v[1] = {robust_chi2_part(cam, poses, lms, delta, ws, sh)};
  cluster_reduce<1, false>(v, sh, xpar);
  return v[0];
}

// Active sets + lookup tables (sparse_optimizer.cpp:168-272 semantics), built once per optimize() phase.
// Edges get "slots" ordered (chunk of landmarks, pose, landmark): a chunk is a run of consecutive landmarks whose
// edges fit the shared-memory staging area of the Schur pass; inside a chunk the edges of one pose are contiguous
// and sorted by landmark, so both the per-landmark and the per-pose passes read the planes nearly sequentially.
__device__ void setup_active(const flv_ba_problem& pb, const int* ep, const int* el, const double* uv,
                             const uint8_t* act, int chunk_area_doubles, int member_buf, Ws& ws, Sh& sh) {
  const int P = pb.n_poses, L = pb.n_landmarks, E = pb.n_edges, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < P * L; i += BA_THREADS) ws.tab[i] = -1;
  for (int i = tid; i < L; i += BA_THREADS) ws.lmask[i] = 0u;
  if (tid < BA_MAX_POSES) sh.pcount[tid] = 0;
  if (tid == 0) sh.overflow = 0;
  __syncthreads();
  for (int e = tid; e < E; e += BA_THREADS) {
    if (!act[e]) continue;
    const int p = ep[e], l = el[e];
    ws.tab[p * L + l] = e;
    atomicOr(&ws.lmask[l], 1u << p);
    atomicAdd(&sh.pcount[p], 1);
  }
  __syncthreads();
  if (tid == 0) {
    int np = 0, nact = 0;
    for (int p = 0; p < P; ++p) {
      nact += sh.pcount[p];
      if (p != pb.fixed_pose && sh.pcount[p] > 0) { sh.pidx[p] = np < BA_MAX_FREE ? np : -1; if (np < BA_MAX_FREE) sh.pose_of[np] = p; ++np; }
      else sh.pidx[p] = -1;
    }
    sh.np = np; sh.nact = nact;
    // staging capacity of the Schur pass: what the reduced system leaves of the dynamic shared memory
    const int npc = np < BA_MAX_FREE ? np : BA_MAX_FREE;
    const int n = 6 * npc;
    // behind the chunk planes: staging buffer of the chunk's pose-pair member lists (bulk-copied while Z is computed), unless
    // the reduced system leaves too little room
    int mb = member_buf;
    int cap = (chunk_area_doubles - (n * (n + 1) + n + 8) - mb / 2) / BA_CHUNK_PLANES;
    if (cap < 256 || pb.fix_landmarks) { mb = 0; cap = (chunk_area_doubles - (n * (n + 1) + n + 8)) / BA_CHUNK_PLANES; }
    cap = cap > BA_MAX_CAP ? BA_MAX_CAP : cap;
    cap &= ~1;                                    // keeps the member buffer behind the planes 16-byte aligned
    sh.mb = mb; sh.cap = cap; sh.capq = cap - P;
  }
  __syncthreads();
  if (sh.np > BA_MAX_FREE) return;
  if (sh.capq < 32 && !pb.fix_landmarks) { if (tid == 0) sh.overflow = 2; __syncthreads(); return; }
  // prefixes over landmarks: edge counts (slot bases) and chunk weights max(count, 1) (bounds landmarks per chunk too)
  for (int l = tid; l < L; l += BA_THREADS) {
    const int c = __popc(ws.lmask[l]);
    ws.lstart[l] = c; ws.lw[l] = c > 0 ? c : 1;
  }
  __syncthreads();
  block_exclusive_scan(ws.lstart, L, sh);
  block_exclusive_scan(ws.lw, L, sh);
  // a cluster of C CTAs splits the chunks evenly: shrink the chunks so that their number is (about) a multiple of C
  if (tid == 0 && sh.C > 1 && !pb.fix_landmarks) {
    const int Wt = ws.lw[L];
    int rounds = (Wt + sh.C * sh.capq - 1) / (sh.C * sh.capq);
    rounds = rounds < 1 ? 1 : rounds;
    const int cq = Wt / (sh.C * rounds) + 1;
    if (cq < sh.capq) sh.capq = cq;
  }
  __syncthreads();
  // chunk of landmark l = lw[l] / capq: a landmark has <= P edges, so a chunk holds < capq + P = cap weight
  const int capq = pb.fix_landmarks ? 0x3fffffff : sh.capq;
  if (tid == 0) { sh.nch = L > 0 ? ws.lw[L - 1] / capq + 1 : 0; }
  __syncthreads();
  const int nch = sh.nch;
  if (nch > BA_MAX_CHUNKS) { if (tid == 0) sh.overflow = 3; __syncthreads(); return; }
  for (int l = tid; l < L; l += BA_THREADS) {
    const int c = ws.lw[l] / capq;
    if (l == 0 || ws.lw[l - 1] / capq != c) { sh.chunk_lb[c] = l; sh.chunk_sb[c] = ws.lstart[l]; }
  }
  if (tid == 0) { sh.chunk_lb[nch] = L; sh.chunk_sb[nch] = sh.nact; }
  __syncthreads();
  // (chunk, pose) run lengths -> offsets
  const int P1 = P + 1;
  for (int task = warp; task < nch * P; task += BA_WARPS) {
    const int ch = task / P, p = task - ch * P;
    const int lb = sh.chunk_lb[ch], le = sh.chunk_lb[ch + 1];
    int cnt = 0;
    for (int l0 = lb; l0 < le; l0 += 32) {
      const int l = l0 + lane;
      cnt += __popc(__ballot_sync(FULL, l < le && ((ws.lmask[l] >> p) & 1u)));
    }
    if (lane == 0) ws.cp_off[ch * P1 + p] = cnt;
  }
  __syncthreads();
  for (int ch = tid; ch < nch; ch += BA_THREADS) {
    int run = sh.chunk_sb[ch];
    for (int p = 0; p < P; ++p) { const int c = ws.cp_off[ch * P1 + p]; ws.cp_off[ch * P1 + p] = run; run += c; }
    ws.cp_off[ch * P1 + P] = run;
  }
  __syncthreads();
  // slot assignment (deterministic: ballot compaction in landmark order)
  for (int task = warp; task < nch * P; task += BA_WARPS) {
    const int ch = task / P, p = task - ch * P;
    const int lb = sh.chunk_lb[ch], le = sh.chunk_lb[ch + 1];
    int base = ws.cp_off[ch * P1 + p];
    for (int l0 = lb; l0 < le; l0 += 32) {
      const int l = l0 + lane;
      const bool has = l < le && ((ws.lmask[l] >> p) & 1u);
      const unsigned bal = __ballot_sync(FULL, has);
      if (has) {
        const int s = base + __popc(bal & ((1u << lane) - 1));
        const int e = ws.tab[p * L + l];
        ws.slot_e[s] = e; ws.slot_pl[s] = p | (l << 8);
        const int lp = ws.lstart[l] + __popc(ws.lmask[l] & ((1u << p) - 1));
        ws.slot_lp[s] = lp;
        PL2(ws.luv)[lp] = make_double2(uv[2 * (size_t)e], uv[2 * (size_t)e + 1]);      // landmark-major copies for the landmark pass
        ws.csr_p[lp] = p; ws.csr_l[lp] = l;
        PL2(ws.uvs)[s] = make_double2(uv[2 * (size_t)e], uv[2 * (size_t)e + 1]);
        ws.tab[p * L + l] = s;
      }
      base += __popc(bal);
    }
  }
  __syncthreads();
  if (pb.fix_landmarks) return;
  // pose-pair member lists per chunk: count, prefix, fill
  const int np = sh.np, nblk = np * (np + 1) / 2;
  for (int blk = tid; blk < nblk; blk += BA_THREADS) {
    int a = 0, rem = blk;
    while (rem >= np - a) { rem -= np - a; ++a; }
    sh.pair_a[blk] = (unsigned char)a; sh.pair_b[blk] = (unsigned char)(a + rem);
  }
  __syncthreads();
  for (int task = warp; task < nch * nblk; task += BA_WARPS) {
    const int ch = task / nblk, blk = task - ch * nblk;
    const unsigned need = (1u << sh.pose_of[sh.pair_a[blk]]) | (1u << sh.pose_of[sh.pair_b[blk]]);
    const int lb = sh.chunk_lb[ch], le = sh.chunk_lb[ch + 1];
    int cnt = 0;
    for (int l0 = lb; l0 < le; l0 += 32) {
      const int l = l0 + lane;
      cnt += __popc(__ballot_sync(FULL, l < le && (ws.lmask[l] & need) == need));
    }
    if (lane == 0) ws.poff[task] = cnt;
  }
  __syncthreads();
  // pick-up order of the pair tasks: largest first (a warp that draws a diagonal pair late would make the others wait)
  for (int blk = tid; blk < nblk; blk += BA_THREADS) {
    int tot = 0;
    for (int ch = 0; ch < nch; ++ch) tot += ws.poff[ch * nblk + blk];
    sh.task_bounds[blk] = tot;
  }
  __syncthreads();
  for (int blk = tid; blk < nblk; blk += BA_THREADS) {
    const int mine = sh.task_bounds[blk];
    int rank = 0;
    for (int j = 0; j < nblk; ++j) { const int o = sh.task_bounds[j]; rank += (o > mine) || (o == mine && j < blk); }
    sh.task_order[rank] = (unsigned short)blk;
  }
  __syncthreads();
  block_exclusive_scan(ws.poff, nch * nblk, sh);
  if (ws.poff[nch * nblk] > ws.pair_cap) { if (tid == 0) sh.overflow = 1; __syncthreads(); return; }
  for (int task = warp; task < nch * nblk; task += BA_WARPS) {
    const int ch = task / nblk, blk = task - ch * nblk;
    const int pa = sh.pose_of[sh.pair_a[blk]], pb_ = sh.pose_of[sh.pair_b[blk]];
    const unsigned need = (1u << pa) | (1u << pb_);
    const int lb = sh.chunk_lb[ch], le = sh.chunk_lb[ch + 1], sb = sh.chunk_sb[ch];
    int base = ws.poff[task];
    for (int l0 = lb; l0 < le; l0 += 32) {
      const int l = l0 + lane;
      const bool mem = l < le && (ws.lmask[l] & need) == need;
      const unsigned bal = __ballot_sync(FULL, mem);
      if (mem)
        ws.pairs[base + __popc(bal & ((1u << lane) - 1))] =
            (ws.tab[pa * L + l] - sb) | ((ws.tab[pb_ * L + l] - sb) << 10) | ((l - lb) << 20);
      base += __popc(bal);
    }
  }
  __syncthreads();
}

// setup runs on rank 0 of the window's cluster (it writes the global tables); the other CTAs copy the shared-memory part of its
// result through distributed shared memory.  Then every CTA takes a contiguous run of chunks: all edges of a landmark lie in the
// landmark's chunk, so a CTA's landmarks, slots and CSR entries are private to it in every pass.
__device__ void setup_cluster(const flv_ba_problem& pb, const int* ep, const int* el, const double* uv,
                              const uint8_t* act, int chunk_area_doubles, int member_buf, Ws& ws, Sh& sh) {
  if (sh.rank == 0) setup_active(pb, ep, el, uv, act, chunk_area_doubles, member_buf, ws, sh);
  if (sh.C > 1) {
    cg::cluster_group cl = cg::this_cluster();
    cl.sync();
    if (sh.rank != 0) {
      const int* src = cl.map_shared_rank(sh.pidx, 0);
      int* dst = sh.pidx;
      const int nw = (int)((offsetof(Sh, capq) + sizeof(int) - offsetof(Sh, pidx)) / sizeof(int));
      for (int i = threadIdx.x; i < nw; i += BA_THREADS) dst[i] = src[i];
    }
    cl.sync();
  }
  if (threadIdx.x == 0) {
    sh.mb_chunk = -1;                              // new member lists
    const int nch = (sh.np > BA_MAX_FREE || sh.overflow) ? 0 : sh.nch;
    sh.c0 = (nch * sh.rank) / sh.C; sh.c1 = (nch * (sh.rank + 1)) / sh.C;
    sh.l0 = nch ? sh.chunk_lb[sh.c0] : 0; sh.l1 = nch ? sh.chunk_lb[sh.c1] : 0;
    sh.s0 = nch ? sh.chunk_sb[sh.c0] : 0; sh.s1 = nch ? sh.chunk_sb[sh.c1] : 0;
  }
  __syncthreads();
}

// Linearisation.  Pass A: thread per slot (edge): residual, Jacobians, W = rho' B^T A, Bw = sqrt(rho') B,
// g = -sqrt(rho') r and the edge's share of Hll / bl.  Pass B: thread per landmark sums the shares (pose order).
// Pass C: warp per (free pose, part of its slot runs): Hpp diagonal block and bp.
__device__ void build_system(const flv_ba_problem& pb, const Cam& cam, const double* poses, const double* lms,
                             double delta, Ws& ws, Sh& sh) {
  const int P = pb.n_poses, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ML = ws.ML;
  const double d2 = delta * delta;
  if (!pb.fix_landmarks) {
    mark(sh, 10);
    for (int l = sh.l0 + tid; l < sh.l1; l += BA_THREADS) {
      const int j0 = ws.lstart[l], j1 = ws.lstart[l + 1];
      if (j0 == j1) continue;
      double H[6] = {0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
      const double X[3] = {lms[3 * (size_t)l], lms[3 * (size_t)l + 1], lms[3 * (size_t)l + 2]};
      for (int j = j0; j < j1; ++j) {             // pose order; the landmark-major copies make the addresses data-independent
        const double2 q2 = CPL2(ws.luv)[j];
        const double uv[2] = {q2.x, q2.y};
        double r[2], A[6], Bu[12];
        edge_eval<true>(poses + 7 * ws.csr_p[j], X, uv, cam, r, A, Bu);          // the pose Jacobian is dead code here
        const double c = r[0] * r[0] + r[1] * r[1];
        const double rho1 = (c <= d2) ? 1.0 : delta / sqrt(c);
        const double o0 = -r[0] * rho1, o1 = -r[1] * rho1;
#pragma unroll
        for (int i = 0; i < 3; ++i) b[i] += A[i] * o0 + A[3 + i] * o1;
        H[0] += rho1 * (A[0] * A[0] + A[3] * A[3]); H[1] += rho1 * (A[0] * A[1] + A[3] * A[4]);
        H[2] += rho1 * (A[0] * A[2] + A[3] * A[5]); H[3] += rho1 * (A[1] * A[1] + A[4] * A[4]);
        H[4] += rho1 * (A[1] * A[2] + A[4] * A[5]); H[5] += rho1 * (A[2] * A[2] + A[5] * A[5]);
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) ws.Hll[i * ML + l] = H[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) ws.bl[i * ML + l] = b[i];
    }
    __syncthreads();
    mark(sh, 11);
  }
  // pose diagonal blocks: task = (free pose, split part); a part owns a contiguous range of chunks
  const int np = sh.np, nch = sh.c1 - sh.c0, P1 = P + 1;       // this CTA's chunks
  int split = BA_WARPS / (np > 0 ? np : 1);               // (4 parts per pose measured slower than 1: 0.80 M vs 0.68 M cycles)
  split = split < 1 ? 1 : (split > 4 ? 4 : split);
  if (split > nch) split = nch > 0 ? nch : 1;
  for (int task = warp; task < np * split; task += BA_WARPS) {
    const int pi = task / split, part = task - pi * split, p = sh.pose_of[pi];
    const int c0 = sh.c0 + (nch * part) / split, c1 = sh.c0 + (nch * (part + 1)) / split;
    double H[21], b[6];
#pragma unroll
    for (int i = 0; i < 21; ++i) H[i] = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) b[i] = 0;
    for (int ch = c0; ch < c1; ++ch) {
      const int s0 = ws.cp_off[ch * P1 + p], s1 = ws.cp_off[ch * P1 + p + 1];
      for (int s = s0 + lane; s < s1; s += 32) {
        double B[12], g0, g1;
        {                                             // sqrt(rho') B and -sqrt(rho') r recomputed: cheaper than 7 loads per edge
          const int pl = ws.slot_pl[s];
          const double2 q2 = CPL2(ws.uvs)[s];
          const double uv[2] = {q2.x, q2.y};
          double r[2], A[6];
          edge_eval<true>(poses + 7 * p, lms + 3 * (size_t)(pl >> 8), uv, cam, r, A, B);
          const double c = r[0] * r[0] + r[1] * r[1];
          const double sr = (c <= d2) ? 1.0 : sqrt(delta / sqrt(c));
#pragma unroll
          for (int i = 0; i < 12; ++i) B[i] *= sr;
          g0 = -sr * r[0]; g1 = -sr * r[1];
        }
        int k2 = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          b[i] += B[i] * g0 + B[6 + i] * g1;
#pragma unroll
          for (int j = i; j < 6; ++j) H[k2++] += B[i] * B[j] + B[6 + i] * B[6 + j];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 21; ++i) { const double v = warp_sum(H[i]); if (lane == 0) sh.part[task][i] = v; }
#pragma unroll
    for (int i = 0; i < 6; ++i) { const double v = warp_sum(b[i]); if (lane == 0) sh.part[task][21 + i] = v; }
  }
  __syncthreads();
  for (int i = tid; i < np * 27; i += BA_THREADS) {
    const int pi = i / 27, k = i - 27 * pi;
    double v = 0;
    for (int part = 0; part < split; ++part) v += sh.part[pi * split + part][k];
    sh.Hdp[pi][k] = v;
  }
  // the pose blocks live on rank 0 (it owns the reduced system): sum the cluster's shares in rank order
  if (sh.C > 1) cg::this_cluster().sync(); else __syncthreads();
  if (sh.rank == 0) {
    for (int i = tid; i < np * 27; i += BA_THREADS) {
      const int pi = i / 27, k = i - 27 * pi;
      double v = sh.Hdp[pi][k];
      if (sh.C > 1) {
        cg::cluster_group cl = cg::this_cluster();
        for (int r = 1; r < sh.C; ++r) v += cl.map_shared_rank(&sh.Hdp[pi][k], r)[0];
      }
      if (k < 21) sh.Hd[pi][k] = v; else sh.bp[6 * pi + k - 21] = v;
    }
  }
  __syncthreads();
}

// (H + lambda I) x = b through the Schur complement: S (n x ld, shared, lower triangle used), y, x in shared.
// The landmark side is streamed through shared memory chunk by chunk (W planes + Dinv + Dinv*bl of the chunk's
// landmarks); a warp owns a pose pair, two lanes share a member landmark (lane parity h owns columns 3h..3h+2 of the
// 6x6 block), member lists hold chunk-local positions, so every operand of the inner loop is a shared-memory read.
__device__ void solve_system(const flv_ba_problem& pb, const Cam& cam, const double* poses, const double* lms, double delta,
                             double lambda, double* S, double* y, double* chunk, int ld, Ws& ws, Sh& sh, unsigned& mpar) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int np = sh.np, n = 6 * np, ML = ws.ML;
  const int nblk = np * (np + 1) / 2;
  // reduced system starts as the pose blocks (+ lambda); the chunks subtract W Dinv W^T from it
  // pose-pair member lists -> shared memory by TMA bulk copy (members_stage): all chunks of this CTA at once if they fit (then
  // they stay resident over the LM trials of the phase), else chunk by chunk, issued early so the copy overlaps the Z pass
  int* Mb = reinterpret_cast<int*>(chunk + BA_CHUNK_PLANES * sh.cap);
  if (!pb.fix_landmarks && sh.c0 < sh.c1 && tid == 0) {
    if (!members_stage(sh.c0, sh.c1, -2, nblk, Mb, ws, sh)) members_stage(sh.c0, sh.c0 + 1, sh.c0, nblk, Mb, ws, sh);
  }
  // (in a cluster every CTA accumulates the products of its own chunks into its own S / y; rank 0 adds them up below)
  for (int i = tid; i < n * ld; i += BA_THREADS) S[i] = 0;
  __syncthreads();
  if (sh.rank == 0) {
    for (int i = tid; i < np * 36; i += BA_THREADS) {
      const int pi = i / 36, k = i - 36 * pi, r = k / 6, c = k - 6 * r;
      double v = sh.Hd[pi][r <= c ? sym21(r, c) : sym21(c, r)];
      if (r == c) v += lambda;
      S[(6 * pi + r) * ld + 6 * pi + c] = v;
    }
  }
  if (tid < n) y[tid] = sh.rank == 0 ? sh.bp[tid] : 0.0;
  if (!pb.fix_landmarks) {
    for (int l = sh.l0 + tid; l < sh.l1; l += BA_THREADS) {
      if (!ws.lmask[l]) continue;
      const double a = ws.Hll[l] + lambda, b = ws.Hll[ML + l], c = ws.Hll[2 * ML + l], d = ws.Hll[3 * ML + l] + lambda,
                   e = ws.Hll[4 * ML + l], f = ws.Hll[5 * ML + l] + lambda;
      const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
      const double id = 1.0 / (a * c00 + b * c01 + c * c02);
      const double D0 = c00 * id, D1 = c01 * id, D2 = c02 * id, D3 = (a * f - c * c) * id, D4 = (b * c - a * e) * id,
                   D5 = (a * d - b * b) * id;
      ws.Dinv[l] = D0; ws.Dinv[ML + l] = D1; ws.Dinv[2 * ML + l] = D2; ws.Dinv[3 * ML + l] = D3;
      ws.Dinv[4 * ML + l] = D4; ws.Dinv[5 * ML + l] = D5;
      // Dinv = Ld Ld^T (3x3 Cholesky): W Dinv W^T = (W Ld)(W Ld)^T, so the pair products need ONE factor Z = W Ld per edge
      const double l00 = sqrt(D0), il00 = 1.0 / l00, l10 = D1 * il00, l20 = D2 * il00;
      const double l11 = sqrt(D3 - l10 * l10), l21 = (D4 - l20 * l10) / l11, l22 = sqrt(D5 - l20 * l20 - l21 * l21);
      ws.Ld[l] = l00; ws.Ld[ML + l] = l10; ws.Ld[2 * ML + l] = l20; ws.Ld[3 * ML + l] = l11; ws.Ld[4 * ML + l] = l21;
      ws.Ld[5 * ML + l] = l22;
      const double b0 = ws.bl[l], b1 = ws.bl[ML + l], b2 = ws.bl[2 * ML + l];
      ws.Dv[l] = l00 * b0 + l10 * b1 + l20 * b2; ws.Dv[ML + l] = l11 * b1 + l21 * b2; ws.Dv[2 * ML + l] = l22 * b2;   // Ld^T bl
    }
    __syncthreads();
    mark(sh, 8);
    const int cap = sh.cap;
    double* Zc = chunk;                   // [18][cap]  Z = W Ld of the chunk's slots
    double* Vc = chunk + 18 * cap;        // [3][cap]   Ld^T bl of the chunk's landmarks
    for (int ch = sh.c0; ch < sh.c1; ++ch) {
      const int sb = sh.chunk_sb[ch], ns = sh.chunk_sb[ch + 1] - sb, lb = sh.chunk_lb[ch], nl = sh.chunk_lb[ch + 1] - lb;

      // member-list bounds of every pose pair for this chunk -> shared memory; tasks are drawn dynamically below
      for (int i = tid; i <= nblk; i += BA_THREADS) sh.task_bounds[i] = ws.poff[ch * nblk + i];
      if (tid == 0) sh.next_task = 0;
      for (int i = tid; i < ns; i += BA_THREADS) {
        const int sl = sb + i, pl = ws.slot_pl[sl];
        if (sh.pidx[pl & 255] < 0) continue;                            // fixed pose: Z is never read
        const int l = pl >> 8;
        const double2 q2 = CPL2(ws.uvs)[sl];
        const double uv[2] = {q2.x, q2.y};
        double w[18];
        edge_W(poses + 7 * (pl & 255), lms + 3 * (size_t)l, uv, cam, delta, delta * delta, w);   // state == linearisation point here
        const double l00 = ws.Ld[l], l10 = ws.Ld[ML + l], l20 = ws.Ld[2 * ML + l], l11 = ws.Ld[3 * ML + l],
                     l21 = ws.Ld[4 * ML + l], l22 = ws.Ld[5 * ML + l];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          Zc[(3 * r) * cap + i] = w[3 * r] * l00 + w[3 * r + 1] * l10 + w[3 * r + 2] * l20;
          Zc[(3 * r + 1) * cap + i] = w[3 * r + 1] * l11 + w[3 * r + 2] * l21;
          Zc[(3 * r + 2) * cap + i] = w[3 * r + 2] * l22;
        }
      }
      for (int i = tid; i < nl; i += BA_THREADS) {
        const double v0 = ws.Dv[lb + i], v1 = ws.Dv[ML + lb + i], v2 = ws.Dv[2 * ML + lb + i];
        Vc[i] = v0; Vc[cap + i] = v1; Vc[2 * cap + i] = v2;
      }
      __syncthreads();
      const bool staged = sh.mb_chunk == -2 || sh.mb_chunk == ch;
      if (staged && mpar != sh.m_issued) { mbar_wait(&sh.mbar, mpar & 1u); ++mpar; }     // mpar counts the copies this thread has waited for
      const int* members = staged ? Mb - sh.m_base : ws.pairs;
      mark(sh, 9);
      for (;;) {
        int tsk = 0;
        if (lane == 0) tsk = atomicAdd(&sh.next_task, 1);
        tsk = __shfl_sync(FULL, tsk, 0);
        if (tsk >= nblk) break;
        const int blk = sh.task_order[tsk];
        const int i0 = sh.task_bounds[blk], i1 = sh.task_bounds[blk + 1];
        if (i0 == i1) continue;
        const int a = sh.pair_a[blk], b = sh.pair_b[blk];
        // one lane per member: acc[6 i + jc] += sum_c Za[i][c] Zb[jc][c]  (block (a,b) of W Dinv W^T)
        double acc[36];
#pragma unroll
        for (int i = 0; i < 36; ++i) acc[i] = 0;
        int k = i0 + lane;
        int ent = k < i1 ? members[k] : 0;
        for (; k < i1; k += 32) {
          const int cur = ent;
          if (k + 32 < i1) ent = members[k + 32];           // prefetch the next round's member
          const int pa = cur & 1023, pbb = (cur >> 10) & 1023;
          double za[18];
#pragma unroll
          for (int i = 0; i < 18; ++i) za[i] = Zc[i * cap + pa];
#pragma unroll
          for (int jc = 0; jc < 6; ++jc) {
            const double z0 = Zc[(3 * jc) * cap + pbb], z1 = Zc[(3 * jc + 1) * cap + pbb], z2 = Zc[(3 * jc + 2) * cap + pbb];
#pragma unroll
            for (int i = 0; i < 6; ++i) acc[6 * i + jc] += za[3 * i] * z0 + za[3 * i + 1] * z1 + za[3 * i + 2] * z2;
          }
        }
        // butterfly reduce-scatter of the first 32 sums (lane e ends with the total of acc[e]); plain tree for the last 4
#pragma unroll
        for (int e = 32; e < 36; ++e) acc[e] = warp_sum(acc[e]);
#pragma unroll
        for (int w2 = 16; w2 >= 1; w2 >>= 1) {
          const bool up = (lane & w2) != 0;
#pragma unroll
          for (int i = 0; i < w2; ++i) {
            const double send = up ? acc[i] : acc[i + w2];
            const double keep = up ? acc[i + w2] : acc[i];
            acc[i] = keep + __shfl_xor_sync(FULL, send, w2);
          }
        }
        {
          const int e = lane, i = e / 6, jc = e - 6 * i;
          S[(6 * b + jc) * ld + 6 * a + i] -= acc[0];        // block (b,a), b >= a: lower triangle (both triangles if a == b)
          if (lane < 4) {
            const double v = lane == 0 ? acc[32] : lane == 1 ? acc[33] : lane == 2 ? acc[34] : acc[35];
            S[(6 * b + 2 + lane) * ld + 6 * a + 5] -= v;     // e = 32 + lane: i = 5, jc = 2 + lane
          }
        }
        if (a == b) {
          // right-hand side of pose a: y_a -= sum_l Z_al (Ld^T bl)_l
          double cf[6] = {0, 0, 0, 0, 0, 0};
          for (int k2 = i0 + lane; k2 < i1; k2 += 32) {
            const int cur = members[k2];
            const int pa = cur & 1023, lp = cur >> 20;
            const double v0 = Vc[lp], v1 = Vc[cap + lp], v2 = Vc[2 * cap + lp];
#pragma unroll
            for (int i = 0; i < 6; ++i)
              cf[i] += Zc[(3 * i) * cap + pa] * v0 + Zc[(3 * i + 1) * cap + pa] * v1 + Zc[(3 * i + 2) * cap + pa] * v2;
          }
#pragma unroll
          for (int i = 0; i < 6; ++i) { const double v = warp_sum(cf[i]); if (lane == 0) y[6 * a + i] -= v; }
        }
      }
      __syncthreads();
      mark(sh, 2);
      if (tid == 0 && sh.mb_chunk != -2 && ch + 1 < sh.c1) members_stage(ch + 1, ch + 2, ch + 1, nblk, Mb, ws, sh);   // (the buffer is free)
    }
  }
  if (tid == 0) sh.fail = 0;
  if (sh.C > 1) {
    // reduced system = sum of the cluster's shares (lower triangle + right-hand side), added in rank order on rank 0
    cg::cluster_group cl = cg::this_cluster();
    cl.sync();
    if (sh.rank == 0) {
      for (int i = tid; i < n * ld + n; i += BA_THREADS) {
        const int r = i / ld, c = i - r * ld;
        if (i < n * ld && c > r) continue;                    // upper triangle is never read (y follows S: rows n.. are y)
        double* dst = i < n * ld ? S + i : y + (i - n * ld);
        double v = *dst;
        for (int q = 1; q < sh.C; ++q) v += *cl.map_shared_rank(dst, q);
        *dst = v;
      }
    }
  }
  __syncthreads();
  mark(sh, 2);
  if (sh.rank == 0) {
  // Right-looking Cholesky (LDL^T form) of the augmented matrix [S ; y^T] (row n = right-hand side), no square roots: after
  // its step column j holds A_ij = L_ij sqrt(d_j), the diagonal d_j = L_jj^2, and x_j = (y_j - sum_{k>j} A_kj x_k) / d_j.
  // TWO columns per barrier (n = 6 np is even): a step is bound by its fixed latency (barrier + dependent shared-memory loads:
  // ~690 cycles per column measured with one column per barrier), not by the trailing update, so every thread redoes the tiny
  // 2x2 pivot block and the column-(j+1) entries it needs instead of waiting for another barrier.  Same operations in the same
  // order per element as the one-column form.  The final column-(j+1) entries are parked in col1 and written back one step
  // later (other threads still read the raw column during the step).
  // (re-spreading the threads over the live rows in every step measured slower: the integer division costs more than it saves)
  int T = BA_THREADS / (n + 1);                   // threads per row (no shuffles here: any count works)
  T = T < 1 ? 1 : (T > 32 ? 32 : T);
  const int row = tid / T, t = tid - row * T;
  bool bad = false;
  if (tid == 0) sh.invd[0] = 1.0 / S[0];
  __syncthreads();
  double* Ai = row < n ? S + row * ld : y;        // (dereferenced for row <= n only)
  const int kmax = row < n ? row : n - 1;
  for (int j = 0; j < n; j += 2) {
    if (j > 0 && t == 0 && row >= j && row <= n) Ai[j - 1] = sh.col1[row];     // column j-1, final since the previous step
    const double d0 = S[j * ld + j];
    const double inv0 = sh.invd[j];                          // 1 / d_j, computed by the owner of d_j during the previous step
    const double a10 = S[(j + 1) * ld + j];
    const double f10 = a10 * inv0;
    const double d1 = S[(j + 1) * ld + j + 1] - f10 * a10;
    if (!(d0 > 0) || !(d1 > 0)) { bad = true; break; }       // same values in every thread: uniform exit
    const double inv1 = 1.0 / d1;
    if (tid == 0) sh.invd[j + 1] = inv1;
    if (row >= j + 2 && row <= n) {
      const double fi0 = Ai[j] * inv0;
      const double ai1 = Ai[j + 1] - fi0 * a10;
      const double fi1 = ai1 * inv1;
#pragma unroll 2
      for (int k = j + 2 + t; k <= kmax; k += T) {
        const double ak0 = S[k * ld + j];
        const double ak1 = S[k * ld + j + 1] - (ak0 * inv0) * a10;
        double v = Ai[k] - fi0 * ak0;
        v = v - fi1 * ak1;
        Ai[k] = v;
        if (k == j + 2 && row == j + 2) sh.invd[j + 2] = 1.0 / v;   // next pivot: its reciprocal leaves the critical path
      }
      if (t == 0) sh.col1[row] = ai1;
    }
    __syncthreads();
  }
  if (!bad && t == 0 && row == n) y[n - 1] = sh.col1[n];     // last odd column: only the right-hand side row is left
  if (bad && tid == 0) sh.fail = 1;
  __syncthreads();
  mark(sh, 3);
  // back substitution inside one warp: lane holds w[lane + 32 q]; row j of the factor is read contiguously
  if (warp == 0 && !bad) {
    double wv[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) wv[q] = (lane + 32 * q) < n ? y[lane + 32 * q] : 0.0;
    for (int j = n - 1; j >= 0; --j) {
      const int jq = j >> 5, jl = j & 31;
      double own = wv[0];
#pragma unroll
      for (int q = 1; q < 5; ++q) own = jq == q ? wv[q] : own;
      const double xj = __shfl_sync(FULL, own, jl) * sh.invd[j];
      const double* Aj = S + j * ld;
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const int k = lane + 32 * q;
        if (k < j) wv[q] -= Aj[k] * xj;
      }
      if (lane == 0) sh.x[j] = xj;
    }
  }
  }   // rank 0
  if (sh.C > 1) {
    // the increment of the poses (and the failure flag) go back to every CTA
    cg::cluster_group cl = cg::this_cluster();
    cl.sync();
    if (sh.rank != 0) {
      if (tid < n) sh.x[tid] = cl.map_shared_rank(sh.x, 0)[tid];
      if (tid == 0) sh.fail = *cl.map_shared_rank(&sh.fail, 0);
    }
  }
  __syncthreads();
  mark(sh, 4);
}

// state update (sparse_optimizer.cpp:433-446) incl. landmark back-substitution (block_solver.hpp:422-444).
// Returns sum_j x_j (lambda x_j + b_j) (computeScale, optimization_algorithm_levenberg.cpp:168-175).
__device__ double apply_update(const flv_ba_problem& pb, const Cam& cam, double delta, double lambda, double* poses, double* lms,
                               double* scratch, int scratch_doubles, Ws& ws, Sh& sh) {
  const int P = pb.n_poses, tid = threadIdx.x, ML = ws.ML;
  double sc = 0;
  for (int i = tid; i < 7 * P; i += BA_THREADS) sh.pbk[i] = poses[i];
  if (!pb.fix_landmarks) {
    // thread per landmark: c = bl - sum_edges W^T x_p with W recomputed per edge (poses are still the linearisation point:
    // they move after the barrier below; a thread only writes its own landmark), dX = Dinv c
    const double d2 = delta * delta;
    const int jb = sh.s0, nown = sh.s1 - sh.s0;               // this CTA's CSR entries (= its slot range: same landmarks)
    const bool in_smem = 3 * nown <= scratch_doubles;         // t_j = W_j^T x_p of every own CSR entry fits the (idle) chunk area
    if (in_smem) {
      // thread per CSR entry: no divergence over the landmarks' edge counts, coalesced luv / csr_p reads
      for (int j = jb + tid; j < sh.s1; j += BA_THREADS) {
        const int p = ws.csr_p[j], pi = sh.pidx[p];
        double t0 = 0, t1 = 0, t2 = 0;
        if (pi >= 0) {
          const int l = ws.csr_l[j];
          const double2 q2 = CPL2(ws.luv)[j];
          const double uv[2] = {q2.x, q2.y};
          double w[18];
          edge_W(poses + 7 * p, lms + 3 * (size_t)l, uv, cam, delta, d2, w);
          const double* xp = sh.x + 6 * pi;
#pragma unroll
          for (int i = 0; i < 6; ++i) { t0 += w[3 * i] * xp[i]; t1 += w[3 * i + 1] * xp[i]; t2 += w[3 * i + 2] * xp[i]; }
        }
        scratch[j - jb] = t0; scratch[nown + j - jb] = t1; scratch[2 * nown + j - jb] = t2;
      }
      __syncthreads();
    }
    for (int l = sh.l0 + tid; l < sh.l1; l += BA_THREADS) {
      double* X = lms + 3 * (size_t)l;
      const double X0[3] = {X[0], X[1], X[2]};
      ws.lbk[3 * (size_t)l] = X0[0]; ws.lbk[3 * (size_t)l + 1] = X0[1]; ws.lbk[3 * (size_t)l + 2] = X0[2];
      const int j0 = ws.lstart[l], j1 = ws.lstart[l + 1];
      if (j0 == j1) continue;
      const double bl0 = ws.bl[l], bl1 = ws.bl[ML + l], bl2 = ws.bl[2 * ML + l];
      double c0 = bl0, c1 = bl1, c2 = bl2;
      if (in_smem) {
        for (int j = j0 - jb; j < j1 - jb; ++j) { c0 -= scratch[j]; c1 -= scratch[nown + j]; c2 -= scratch[2 * nown + j]; }
      } else
      for (int j = j0; j < j1; ++j) {
        const int p = ws.csr_p[j], pi = sh.pidx[p];
        if (pi < 0) continue;
        const double2 q2 = CPL2(ws.luv)[j];
        const double uv[2] = {q2.x, q2.y};
        double w[18];
        edge_W(poses + 7 * p, X0, uv, cam, delta, d2, w);
        const double* xp = sh.x + 6 * pi;
        double t0 = 0, t1 = 0, t2 = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) { t0 += w[3 * i] * xp[i]; t1 += w[3 * i + 1] * xp[i]; t2 += w[3 * i + 2] * xp[i]; }
        c0 -= t0; c1 -= t1; c2 -= t2;
      }
      const double D0 = ws.Dinv[l], D1 = ws.Dinv[ML + l], D2 = ws.Dinv[2 * ML + l], D3 = ws.Dinv[3 * ML + l],
                   D4 = ws.Dinv[4 * ML + l], D5 = ws.Dinv[5 * ML + l];
      const double x0 = D0 * c0 + D1 * c1 + D2 * c2, x1 = D1 * c0 + D3 * c1 + D4 * c2, x2 = D2 * c0 + D4 * c1 + D5 * c2;
      sc += x0 * (lambda * x0 + bl0) + x1 * (lambda * x1 + bl1) + x2 * (lambda * x2 + bl2);
      X[0] += x0; X[1] += x1; X[2] += x2;
    }
  }
  if (sh.rank == 0 && tid < 6 * sh.np) sc += sh.x[tid] * (lambda * sh.x[tid] + sh.bp[tid]);
  __syncthreads();     // backups of poses complete before anyone overwrites
  if (tid < sh.np) pose_oplus(poses + 7 * sh.pose_of[tid], sh.x + 6 * tid);     // every CTA moves its own copy of the poses
  return block_sum(sc, sh.red);      // this CTA's share
}

__device__ void restore_state(const flv_ba_problem& pb, double* poses, double* lms, Ws& ws, Sh& sh) {
  for (int i = threadIdx.x; i < 7 * pb.n_poses; i += BA_THREADS) poses[i] = sh.pbk[i];
  if (!pb.fix_landmarks)
    for (int i = 3 * sh.l0 + threadIdx.x; i < 3 * sh.l1; i += BA_THREADS) lms[i] = ws.lbk[i];
  __syncthreads();
}

__host__ __device__ inline size_t pair_capacity(int max_poses, int max_edges) {
  return (size_t)max_edges * ((max_poses + 2) / 2);
}
__host__ __device__ inline size_t poff_capacity() { return (size_t)BA_MAX_CHUNKS * BA_MAX_PAIRS + 1; }

// workspace carve-up (doubles first, then ints); shared by the kernel and ws_stride_bytes()
struct WsLayout {
  size_t pbk, lbk, luv, uvs, Hll, bl, Dinv, Dv, Ld, n_doubles;
  size_t tab, lmask, slot_e, slot_pl, slot_lp, csr_p, csr_l, lw, lstart, cp_off, poff, pairs, n_ints;
};
__host__ __device__ inline WsLayout ws_layout(int MP, int ML, int ME) {
  WsLayout o; size_t d = 0, i = 0;
  const size_t E = (size_t)ME, L = (size_t)ML, P = (size_t)MP;
  o.pbk = d; d += 7 * P + (P & 1);
  o.lbk = d; d += 3 * L + (L & 1);
  o.luv = d; d += 2 * E;
  o.uvs = d; d += 2 * E;
  o.Hll = d; d += 6 * L; o.bl = d; d += 3 * L; o.Dinv = d; d += 6 * L; o.Dv = d; d += 3 * L; o.Ld = d; d += 6 * L;
  d += d & 1;                                   // ints start 16-byte aligned (bulk copies of `pairs` need it)
  o.n_doubles = d;
  o.tab = i; i += P * L; o.lmask = i; i += L; o.slot_e = i; i += E; o.slot_pl = i; i += E; o.slot_lp = i; i += E; o.csr_p = i; i += E; o.csr_l = i; i += E;
  o.lw = i; i += L + 1; o.lstart = i; i += L + 1; o.cp_off = i; i += (size_t)BA_MAX_CHUNKS * (P + 1);
  o.poff = i; i += poff_capacity(); i = (i + 3) & ~(size_t)3; o.pairs = i; i += pair_capacity(MP, ME) + 8;
  o.n_ints = i;
  return o;
}

__global__ void __launch_bounds__(BA_THREADS, 1) ba_kernel(BAArgs a) {
  extern __shared__ __align__(16) double dyn[];
  __shared__ Sh sh;
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();     // C CTAs (1, 2 or 4) work on one window
  const int s = blockIdx.x / C, tid = threadIdx.x;
  int xpar = 0;
  const flv_ba_problem pb = a.problems[s];
  const Cam cam = {pb.fx, pb.fy, pb.cx, pb.cy};
  double* poses_g = a.poses + (size_t)s * a.max_poses * 7;
  double* poses = sh.poses;                         // the LM loop works on the shared-memory copy
  double* lms = a.lms + (size_t)s * a.max_lms * 3;
  const int* ep = a.ep + (size_t)s * a.max_edges;
  const int* el = a.el + (size_t)s * a.max_edges;
  const double* uv = a.uv + (size_t)s * a.max_edges * 2;
  uint8_t* act = a.active + (size_t)s * a.max_edges;
  const int P = pb.n_poses, E = pb.n_edges;
  Ws ws;
  {
    const WsLayout lo = ws_layout(a.ws_poses, a.max_lms, a.max_edges);
    double* d = (double*)(a.ws + (size_t)s * a.ws_stride);
    int* ib = (int*)(d + lo.n_doubles);
    ws.pbk = d + lo.pbk; ws.lbk = d + lo.lbk; ws.luv = d + lo.luv;
    ws.uvs = d + lo.uvs; ws.Hll = d + lo.Hll; ws.bl = d + lo.bl; ws.Dinv = d + lo.Dinv; ws.Dv = d + lo.Dv; ws.Ld = d + lo.Ld;
    ws.tab = ib + lo.tab; ws.lmask = (unsigned*)(ib + lo.lmask); ws.slot_e = ib + lo.slot_e; ws.slot_pl = ib + lo.slot_pl; ws.slot_lp = ib + lo.slot_lp; ws.csr_p = ib + lo.csr_p; ws.csr_l = ib + lo.csr_l;
    ws.lw = ib + lo.lw; ws.lstart = ib + lo.lstart; ws.cp_off = ib + lo.cp_off; ws.poff = ib + lo.poff;
    ws.pairs = ib + lo.pairs;
    ws.pair_cap = (int)pair_capacity(a.ws_poses, a.max_edges); ws.ME = a.max_edges; ws.ML = a.max_lms;
  }
  const double delta = a.prm.huber_delta;
  flv_ba_stats st;
  st.iterations_run = 0; st.n_culled = 0; st.ok = 1; st.reserved = 0;
  st.chi2_initial = st.chi2_after1 = st.chi2_final = 0; st.lambda_final = 0;
  if (P < 1 || P > BA_MAX_POSES || E < 0) {
    if (tid == 0 && rank == 0) { st.ok = 0; st.reserved = 1; a.stats[s] = st; }
    return;
  }
  if (tid == 0) { for (int i = 0; i < 16; ++i) sh.prof[i] = 0; sh.tlast = clock64(); sh.rank = rank; sh.C = C; mbar_init(&sh.mbar, 1); sh.m_issued = 0; sh.mb_chunk = -1; }
  unsigned mpar = 0;                                // phase parity of sh.mbar (one completed bulk copy per staged chunk)
  for (int i = tid; i < 7 * P; i += BA_THREADS) sh.poses[i] = poses_g[i];
  __syncthreads();
  st.chi2_initial = robust_chi2_edges(pb, cam, poses, lms, ep, el, uv, act, delta, sh, xpar);
  double lambda = 0;
  for (int phase = 0; phase < 2; ++phase) {
    const int iters = phase == 0 ? a.prm.iters1 : a.prm.iters2;
    setup_cluster(pb, ep, el, uv, act, a.dyn_doubles, a.member_buf, ws, sh);
    mark(sh, 6);
    if (sh.np > BA_MAX_FREE) { st.ok = 0; st.reserved = 2; break; }
    if (sh.overflow) { st.ok = 0; st.reserved = sh.overflow == 1 ? 3 : 4; break; }
    const int n = 6 * sh.np, ld = n + 1;
    double* S = dyn;
    double* y = dyn + (size_t)n * ld;
    double* chunk = y + ((n + 8) & ~1);
    double ni = 2;
    double currentChi = 0;
    for (int it = 0; it < iters; ++it) {
      // g2o recomputes activeRobustChi2 at the start of every iteration; the state only changes through accepted trials
      // (whose chi2 was just computed by the same code on the same state), so the value is carried over bit-identically
      if (it == 0) currentChi = robust_chi2(cam, poses, lms, delta, ws, sh, xpar);
      mark(sh, 0);
      build_system(pb, cam, poses, lms, delta, ws, sh);
      mark(sh, 1);
      if (it == 0) {
        double md = 0;
        if (rank == 0 && tid < sh.np) {
#pragma unroll
          for (int i = 0; i < 6; ++i) md = fmax(md, fabs(sh.Hd[tid][sym21(i, i)]));
        }
        if (!pb.fix_landmarks)
          for (int l = sh.l0 + tid; l < sh.l1; l += BA_THREADS)
            if (ws.lmask[l]) md = fmax(md, fmax(fabs(ws.Hll[l]), fmax(fabs(ws.Hll[3 * ws.ML + l]), fabs(ws.Hll[5 * ws.ML + l]))));
        double mv[1] = {block_max(md, sh.red)};
        cluster_reduce<1, true>(mv, sh, xpar);
        lambda = 1e-5 * mv[0];
        ni = 2;
      }
      double rho = 0;
      int qmax = 0;
      do {
        solve_system(pb, cam, poses, lms, delta, lambda, S, y, chunk, ld, ws, sh, mpar);
        const int ok2 = !sh.fail;
        double scale = 0, tempChi;
        if (ok2) {
          double sv[2];
          sv[0] = apply_update(pb, cam, delta, lambda, poses, lms, chunk, BA_CHUNK_PLANES * sh.cap, ws, sh);   // (planes only: the member lists stay resident)
          __syncthreads();
          mark(sh, 5);
          sv[1] = robust_chi2_part(cam, poses, lms, delta, ws, sh);
          cluster_reduce<2, false>(sv, sh, xpar);       // one cluster barrier for both sums
          scale = sv[0]; tempChi = sv[1];
          mark(sh, 0);
        } else {
          tempChi = 1.7976931348623157e308;
        }
        rho = (currentChi - tempChi) / (scale + 1e-3);
        if (rho > 0 && isfinite(tempChi)) {
          double alpha = 1. - (2 * rho - 1) * (2 * rho - 1) * (2 * rho - 1);
          alpha = fmin(alpha, 2. / 3.);
          lambda *= fmax(1. / 3., alpha);
          ni = 2;
          currentChi = tempChi;
        } else {
          lambda *= ni; ni *= 2;
          if (ok2) restore_state(pb, poses, lms, ws, sh);
          if (!isfinite(lambda)) break;
        }
        ++qmax;
      } while (rho < 0 && qmax < 10);
      if (a.trace && tid == 0 && rank == 0 && st.iterations_run < BA_TRACE_ITERS) {
        double* tr = a.trace + ((size_t)s * BA_TRACE_ITERS + st.iterations_run) * 4;
        tr[0] = currentChi; tr[1] = lambda; tr[2] = rho; tr[3] = (double)qmax;
      }
      ++st.iterations_run;
      if (qmax == 10 || rho == 0 || !isfinite(lambda)) break;
    }
    if (phase == 0) {
      st.chi2_after1 = robust_chi2(cam, poses, lms, delta, ws, sh, xpar);
      // cull: un-robustified chi2 > threshold (vo_localmap.cpp:303-316, optimize_in_frame.cpp:67-74)
      int culled = 0, remaining = 0;
      for (int sl = sh.s0 + tid; sl < sh.s1; sl += BA_THREADS) {
        const int pl = ws.slot_pl[sl];
        const double2 q2 = CPL2(ws.uvs)[sl];
        const double uvv[2] = {q2.x, q2.y};
        double r[2];
        edge_eval<false>(poses + 7 * (pl & 255), lms + 3 * (size_t)(pl >> 8), uvv, cam, r, nullptr, nullptr);
        if (r[0] * r[0] + r[1] * r[1] > a.prm.cull_chi2) { act[ws.slot_e[sl]] = 0; ++culled; } else ++remaining;
      }
      double cv[2];
      cv[0] = block_sum((double)culled, sh.red);
      cv[1] = block_sum((double)remaining, sh.red);
      cluster_reduce<2, false>(cv, sh, xpar);           // (its barrier also publishes the cleared `active` flags to rank 0's next setup)
      st.n_culled = (int)(cv[0] + 0.5);
      const int rem = (int)(cv[1] + 0.5);
      __syncthreads();
      if (rem < a.prm.min_edges_after_cull) { st.ok = 0; break; }
    }
  }
  if (C > 1) cluster.sync();                            // every CTA's landmarks are in global memory before the final chi2 reads them
  st.chi2_final = robust_chi2_edges(pb, cam, poses, lms, ep, el, uv, act, delta, sh, xpar);
  st.lambda_final = lambda;
  if (rank == 0) {
    for (int i = tid; i < 7 * P; i += BA_THREADS) poses_g[i] = sh.poses[i];
    if (tid == 0) {
      a.stats[s] = st;
      if (a.prof) for (int i = 0; i < 16; ++i) a.prof[16 * s + i] = sh.prof[i];
    }
  }
  if (C > 1) cluster.sync();                            // no CTA leaves while another may still read its shared memory
}


// ---- pose-only BA (OptimizeInFrame::optimize, src/processing/optimize_in_frame.cpp:10-90): one pose, all landmarks fixed --------
// The general kernel spends most of a 0.16 ms pose-only solve in its set-up passes (slot tables, CSR, chunking) that a problem
// with ONE 6x6 system does not need.  This kernel runs the same LM control flow (ba_kernel above / optimization_algorithm_
// levenberg.cpp:58-175) on one CTA per sequence with a thread per edge: residual + pose Jacobian -> block-reduced H (21) and
// b (6) -> LDL^T of the damped 6x6 -> exp-map update -> robust chi2; cull of chi2 > threshold between the two optimize() calls.
constexpr int PO_THREADS = 512;
constexpr int PO_WARPS = PO_THREADS / 32;

struct PoSh {
  double part[PO_WARPS][28];
  double H[21], b[6], x[6], pose[7], pbk[7];
  double red[PO_WARPS];
  int fail;
};

__device__ double po_block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0;
#pragma unroll
  for (int i = 0; i < PO_WARPS; ++i) t += red[i];
  return t;
}

__device__ double po_chi2(const flv_ba_problem& pb, const Cam& cam, const double* pose, const double* lms, const int* el, const double* uv,
                          const uint8_t* act, double delta, double* red) {
  double acc = 0;
  const double d2 = delta * delta;
  for (int e = threadIdx.x; e < pb.n_edges; e += PO_THREADS) {
    if (!act[e]) continue;
    double r[2];
    edge_eval<false>(pose, lms + 3 * (size_t)el[e], uv + 2 * (size_t)e, cam, r, nullptr, nullptr);
    const double c = r[0] * r[0] + r[1] * r[1];
    acc += (c <= d2) ? c : 2 * sqrt(c) * delta - d2;
  }
  return po_block_sum(acc, red);
}

__global__ void __launch_bounds__(PO_THREADS, 1) ba_pose_only_kernel(BAArgs a) {
  __shared__ PoSh sh;
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const flv_ba_problem pb = a.problems[s];
  const Cam cam = {pb.fx, pb.fy, pb.cx, pb.cy};
  double* pose_g = a.poses + (size_t)s * a.max_poses * 7;
  const double* lms = a.lms + (size_t)s * a.max_lms * 3;
  const int* el = a.el + (size_t)s * a.max_edges;
  const double* uv = a.uv + (size_t)s * a.max_edges * 2;
  uint8_t* act = a.active + (size_t)s * a.max_edges;
  const double delta = a.prm.huber_delta, d2 = delta * delta;
  flv_ba_stats st;
  st.iterations_run = 0; st.n_culled = 0; st.ok = 1; st.reserved = 0;
  st.chi2_initial = st.chi2_after1 = st.chi2_final = 0; st.lambda_final = 0;
  if (pb.n_poses != 1 || !pb.fix_landmarks || pb.n_edges < 0) {            // not a pose-only problem: this context cannot solve it
    if (tid == 0) { st.ok = 0; st.reserved = 5; a.stats[s] = st; }
    return;
  }
  if (tid < 7) sh.pose[tid] = pose_g[tid];
  __syncthreads();
  const double* pose = sh.pose;
  st.chi2_initial = po_chi2(pb, cam, pose, lms, el, uv, act, delta, sh.red);
  double lambda = 0;
  for (int phase = 0; phase < 2; ++phase) {
    const int iters = phase == 0 ? a.prm.iters1 : a.prm.iters2;
    // the pose is optimised iff it is not the fixed vertex and has at least one active edge (sparse_optimizer.cpp:168-272)
    int cnt = 0;
    for (int e = tid; e < pb.n_edges; e += PO_THREADS) cnt += act[e] ? 1 : 0;
    const int nact = (int)(po_block_sum((double)cnt, sh.red) + 0.5);
    const bool free_pose = pb.fixed_pose != 0 && nact > 0;
    double ni = 2, currentChi = 0;
    for (int it = 0; it < iters; ++it) {
      if (it == 0) currentChi = po_chi2(pb, cam, pose, lms, el, uv, act, delta, sh.red);
      // linearisation: H = sum rho' B^T B, b = -sum rho' B^T r (as sqrt(rho') B and -sqrt(rho') r, like ba_kernel's pose pass)
      double H[21], b[6];
#pragma unroll
      for (int i = 0; i < 21; ++i) H[i] = 0;
#pragma unroll
      for (int i = 0; i < 6; ++i) b[i] = 0;
      if (free_pose)
        for (int e = tid; e < pb.n_edges; e += PO_THREADS) {
          if (!act[e]) continue;
          double r[2], A[6], B[12];
          edge_eval<true>(pose, lms + 3 * (size_t)el[e], uv + 2 * (size_t)e, cam, r, A, B);
          const double c = r[0] * r[0] + r[1] * r[1];
          const double sr = (c <= d2) ? 1.0 : sqrt(delta / sqrt(c));
#pragma unroll
          for (int i = 0; i < 12; ++i) B[i] *= sr;
          const double g0 = -sr * r[0], g1 = -sr * r[1];
          int k2 = 0;
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            b[i] += B[i] * g0 + B[6 + i] * g1;
#pragma unroll
            for (int j = i; j < 6; ++j) H[k2++] += B[i] * B[j] + B[6 + i] * B[6 + j];
          }
        }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 21; ++i) { const double v = warp_sum(H[i]); if (lane == 0) sh.part[warp][i] = v; }
#pragma unroll
      for (int i = 0; i < 6; ++i) { const double v = warp_sum(b[i]); if (lane == 0) sh.part[warp][21 + i] = v; }
      __syncthreads();
      if (tid < 27) {
        double v = 0;
        for (int w = 0; w < PO_WARPS; ++w) v += sh.part[w][tid];
        if (tid < 21) sh.H[tid] = v; else sh.b[tid - 21] = v;
      }
      __syncthreads();
      if (it == 0) {
        double md = 0;
        if (free_pose) for (int i = 0; i < 6; ++i) md = fmax(md, fabs(sh.H[sym21(i, i)]));
        lambda = 1e-5 * md; ni = 2;
      }
      double rho = 0;
      int qmax = 0;
      do {
        // (H + lambda I) x = b by LDL^T without pivoting (every thread redundantly on registers: 6x6)
        bool ok2 = true;
        double scale = 0, tempChi = 1.7976931348623157e308;
        if (free_pose) {
          double M[6][6], y[6], d[6];
#pragma unroll
          for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int j = 0; j <= i; ++j) M[i][j] = sh.H[sym21(j, i)] + (i == j ? lambda : 0.0);
            y[i] = sh.b[i];
          }
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            double dj = M[j][j];
#pragma unroll
            for (int k = 0; k < j; ++k) dj -= M[j][k] * M[j][k] * d[k];
            if (!(dj > 0)) ok2 = false;
            d[j] = dj;
#pragma unroll
            for (int i = j + 1; i < 6; ++i) {
              double v = M[i][j];
#pragma unroll
              for (int k = 0; k < j; ++k) v -= M[i][k] * M[j][k] * d[k];
              M[i][j] = v / dj;
            }
          }
          if (ok2) {
#pragma unroll
            for (int i = 0; i < 6; ++i) {
#pragma unroll
              for (int k = 0; k < i; ++k) y[i] -= M[i][k] * y[k];
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) y[i] /= d[i];
#pragma unroll
            for (int i = 5; i >= 0; --i) {
#pragma unroll
              for (int k = i + 1; k < 6; ++k) y[i] -= M[k][i] * y[k];
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) scale += y[i] * (lambda * y[i] + sh.b[i]);
            __syncthreads();                                   // everyone has read H / b / pose
            if (tid < 7) sh.pbk[tid] = sh.pose[tid];
            if (tid < 6) sh.x[tid] = y[tid];
            __syncthreads();
            if (tid == 0) pose_oplus(sh.pose, sh.x);
            __syncthreads();
          }
        }
        if (ok2) tempChi = po_chi2(pb, cam, pose, lms, el, uv, act, delta, sh.red);
        rho = (currentChi - tempChi) / (scale + 1e-3);
        if (rho > 0 && isfinite(tempChi)) {
          double alpha = 1. - (2 * rho - 1) * (2 * rho - 1) * (2 * rho - 1);
          alpha = fmin(alpha, 2. / 3.);
          lambda *= fmax(1. / 3., alpha);
          ni = 2;
          currentChi = tempChi;
        } else {
          lambda *= ni; ni *= 2;
          if (ok2 && free_pose) { __syncthreads(); if (tid < 7) sh.pose[tid] = sh.pbk[tid]; __syncthreads(); }
          if (!isfinite(lambda)) break;
        }
        ++qmax;
      } while (rho < 0 && qmax < 10);
      if (a.trace && tid == 0 && st.iterations_run < BA_TRACE_ITERS) {
        double* tr = a.trace + ((size_t)s * BA_TRACE_ITERS + st.iterations_run) * 4;
        tr[0] = currentChi; tr[1] = lambda; tr[2] = rho; tr[3] = (double)qmax;
      }
      ++st.iterations_run;
      if (qmax == 10 || rho == 0 || !isfinite(lambda)) break;
    }
    if (phase == 0) {
      st.chi2_after1 = po_chi2(pb, cam, pose, lms, el, uv, act, delta, sh.red);
      int culled = 0, remaining = 0;
      for (int e = tid; e < pb.n_edges; e += PO_THREADS) {
        if (!act[e]) continue;
        double r[2];
        edge_eval<false>(pose, lms + 3 * (size_t)el[e], uv + 2 * (size_t)e, cam, r, nullptr, nullptr);
        if (r[0] * r[0] + r[1] * r[1] > a.prm.cull_chi2) { act[e] = 0; ++culled; } else ++remaining;
      }
      st.n_culled = (int)(po_block_sum((double)culled, sh.red) + 0.5);
      const int rem = (int)(po_block_sum((double)remaining, sh.red) + 0.5);
      __syncthreads();
      if (rem < a.prm.min_edges_after_cull) { st.ok = 0; break; }
    }
  }
  st.chi2_final = po_chi2(pb, cam, pose, lms, el, uv, act, delta, sh.red);
  st.lambda_final = lambda;
  if (tid < 7) pose_g[tid] = sh.pose[tid];
  if (tid == 0) a.stats[s] = st;
}

__global__ void ba_debug_edges_kernel(int n, const double* poses, const double* pts, const double* uv, Cam cam, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r[2], A[6], B[12];
  edge_eval<true>(poses + 7 * (size_t)i, pts + 3 * (size_t)i, uv + 2 * (size_t)i, cam, r, A, B);
  double* o = out + 20 * (size_t)i;
  o[0] = r[0]; o[1] = r[1];
  for (int k = 0; k < 6; ++k) o[2 + k] = A[k];
  for (int k = 0; k < 12; ++k) o[8 + k] = B[k];
}

size_t ws_stride_bytes(int max_poses, int max_lms, int max_edges) {
  const WsLayout lo = ws_layout(max_poses, max_lms, max_edges);
  const size_t b = lo.n_doubles * 8 + lo.n_ints * 4;
  return (b + 255) & ~(size_t)255;
}

// dynamic shared memory of the kernel: everything the SM has left after the static state
int ba_dyn_doubles() {
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, ba_kernel) != cudaSuccess) return 0;
  const long avail = 232448L - (long)fa.sharedSizeBytes - 1024;   // 227 KB per block on sm_100
  return (int)(avail / 8);
}

// one cluster of C CTAs per window (C = 1: plain launch semantics, same kernel)
cudaError_t launch_ba(const BAArgs& a, int n_streams, int C, size_t smem, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_streams * C)); cfg.blockDim = dim3(BA_THREADS);
  cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, ba_kernel, a);
}

// pose-only problems (one pose, fixed landmarks) run on ba_pose_only_kernel unless FLV_BA_POSE_ONLY_FAST=0 (A/B switch)
bool pose_only_fast() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FLV_BA_POSE_ONLY_FAST"); v = (e && atoi(e) == 0) ? 0 : 1; }
  return v == 1;
}

// CTAs per window: FLV_BA_CLUSTER / flv_set_ba_cluster (1, 2 or 4; default 4) for windows that optimise landmarks, 1 for pose-only
int ba_cluster_size(flv_ctx* ctx) {
  if (ctx->ba_cluster <= 0) {
    const char* e = getenv("FLV_BA_CLUSTER");
    const int v = e ? atoi(e) : 4;
    ctx->ba_cluster = (v == 1 || v == 2 || v == 4) ? v : 4;
  }
  return ctx->ba_cluster;
}

}  // namespace

// large windows (26 .. 100 poses): ba_big.cu
size_t flv_ba_big_ws_bytes(int max_poses, int max_lms, int max_edges);
int flv_ba_big_max_poses();
cudaError_t flv_ba_big_launch(int n_streams, const flv_ba_problem* d_problems, const flv_ba_params* prm, double* d_poses, double* d_lms,
                              const int* d_ep, const int* d_el, const double* d_uv, uint8_t* d_act, flv_ba_stats* d_stats, int max_poses,
                              int max_lms, int max_edges, unsigned char* d_ws, double* d_trace, cudaStream_t stream);

int flv_ba_free(flv_ctx* ctx) {
  if (ctx->ba_ws) cudaFree(ctx->ba_ws);
  ctx->ba_ws = nullptr; ctx->ba_ws_bytes = 0;
  return FLV_OK;
}

extern "C" {

int flv_ba_reserve(flv_ctx* ctx, int max_poses, int max_landmarks, int max_edges) {
  if (!ctx || max_poses < 1 || max_landmarks < 1 || max_edges < 1) return FLV_ERR_INVALID;
  if (max_poses > flv_ba_big_max_poses())
    FLV_FAIL(ctx, FLV_ERR_UNSUPPORTED, "window of %d poses: this build supports <= %d", max_poses, flv_ba_big_max_poses());
  // already large enough: nothing to do (the host tracker reserves per frame; a realloc would synchronise the device)
  if (ctx->ba_ws && max_poses <= ctx->ba_max_poses && max_landmarks <= ctx->ba_max_lms && max_edges <= ctx->ba_max_edges) return FLV_OK;
  if (ctx->ba_ws) {                        // grow: keep the union of the old and new capacities
    max_poses = max_poses > ctx->ba_max_poses ? max_poses : ctx->ba_max_poses;
    max_landmarks = max_landmarks > ctx->ba_max_lms ? max_landmarks : ctx->ba_max_lms;
    max_edges = max_edges > ctx->ba_max_edges ? max_edges : ctx->ba_max_edges;
    FLV_CUDA(ctx, cudaDeviceSynchronize());
  }
  flv_ba_free(ctx);
  // windows with more than BA_MAX_FREE + 1 poses run on the global-memory solver (ba_big.cu): its workspace also holds the
  // reduced camera system; the per-stream stride is the larger of the two layouts
  size_t stride = ws_stride_bytes(max_poses < BA_MAX_POSES ? max_poses : BA_MAX_POSES, max_landmarks, max_edges);
  if (max_poses > BA_MAX_FREE + 1) { const size_t b = flv_ba_big_ws_bytes(max_poses, max_landmarks, max_edges); stride = b > stride ? b : stride; }
  ctx->ba_ws_stride = stride;
  // tail: device copies of problems / stats / staging are carved after the per-stream blocks
  size_t total = stride * ctx->S + (size_t)ctx->S * (sizeof(flv_ba_problem) + sizeof(flv_ba_stats) + 128 + BA_TRACE_ITERS * 32) + 512;
  FLV_CUDA(ctx, cudaMalloc(&ctx->ba_ws, total));
  ctx->ba_ws_bytes = total;
  ctx->ba_max_poses = max_poses; ctx->ba_max_lms = max_landmarks; ctx->ba_max_edges = max_edges;
  const size_t smem = (size_t)ba_dyn_doubles() * 8;
  if (smem == 0) FLV_FAIL(ctx, FLV_ERR_CUDA, "cudaFuncGetAttributes(ba_kernel) failed");
  FLV_CUDA(ctx, cudaFuncSetAttribute(ba_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return FLV_OK;
}

int flv_ba_optimize(flv_ctx* ctx, int n_streams, const flv_ba_problem* problems, const flv_ba_params* prm,
                    double* poses, double* landmarks, const int* edge_pose, const int* edge_lm,
                    const double* edge_uv, uint8_t* edge_active, flv_ba_stats* stats, flv_memspace mem) {
  if (!ctx || !problems || !prm || !poses || !landmarks || !edge_pose || !edge_lm || !edge_uv || !edge_active ||
      !stats || n_streams < 1 || n_streams > ctx->S)
    return FLV_ERR_INVALID;
  if (!ctx->ba_ws) FLV_FAIL(ctx, FLV_ERR_INVALID, "flv_ba_reserve has not been called");
  const int MP = ctx->ba_max_poses, ML = ctx->ba_max_lms, ME = ctx->ba_max_edges;
  const size_t stride = ctx->ba_ws_stride;
  unsigned char* tail = (unsigned char*)ctx->ba_ws + stride * ctx->S;
  const int slot0 = prm->ws_slot0;
  if (slot0 < 0 || slot0 + n_streams > ctx->S) FLV_FAIL(ctx, FLV_ERR_INVALID, "ws_slot0 %d + %d streams exceeds %d slots", slot0, n_streams, ctx->S);
  flv_ba_problem* d_prob = (flv_ba_problem*)tail + slot0;
  flv_ba_stats* d_stats = (flv_ba_stats*)(tail + (size_t)ctx->S * sizeof(flv_ba_problem)) + slot0;
  BAArgs a;
  a.prof = (long long*)(tail + (size_t)ctx->S * (sizeof(flv_ba_problem) + sizeof(flv_ba_stats))) + 16 * slot0;
  a.trace = (double*)(tail + (size_t)ctx->S * (sizeof(flv_ba_problem) + sizeof(flv_ba_stats) + 128)) + (size_t)BA_TRACE_ITERS * 4 * slot0;
  a.prm = *prm; a.max_poses = MP; a.max_lms = ML; a.max_edges = ME;
  a.ws = (unsigned char*)ctx->ba_ws + stride * slot0; a.ws_stride = stride;
  // ba_kernel's own layout is computed for min(MP, BA_MAX_POSES) poses (reserve does the same)
  const int MPk = MP < BA_MAX_POSES ? MP : BA_MAX_POSES;
  a.ws_poses = MPk;
  if (!ctx->ba_dyn) ctx->ba_dyn = ba_dyn_doubles();
  a.dyn_doubles = ctx->ba_dyn;
  if (ctx->ba_member_buf < 0) { const char* e = getenv("FLV_BA_TMA"); ctx->ba_member_buf = (e && atoi(e) > 0) ? BA_MEMBER_BUF : 0; }
  a.member_buf = ctx->ba_member_buf;
  const size_t smem = (size_t)a.dyn_doubles * 8;
  const size_t S = n_streams;
  cudaStream_t stream = ctx->ba_stream_set ? ctx->ba_stream : ctx->stream;
  if (mem == FLV_MEM_DEVICE) {
    a.problems = problems; a.poses = poses; a.lms = landmarks; a.ep = edge_pose; a.el = edge_lm; a.uv = edge_uv;
    a.active = edge_active; a.stats = stats;
    if (MP == 1 && pose_only_fast()) {       // a context reserved for one pose (the trackers' OptimizeInFrame): pose-only kernel;
      ba_pose_only_kernel<<<n_streams, PO_THREADS, 0, stream>>>(a);     // a problem that is not pose-only comes back with reserved = 5
      ctx->launches++;
      FLV_CUDA(ctx, cudaGetLastError());
      return FLV_OK;
    }
    if (MP > BA_MAX_FREE + 1) {       // reserved for large windows: the global-memory solver takes any size
      FLV_CUDA(ctx, flv_ba_big_launch(n_streams, problems, prm, poses, landmarks, edge_pose, edge_lm, edge_uv, edge_active, stats, MP, ML, ME,
                                      a.ws, a.trace, stream));
      ctx->launches++;
      return FLV_OK;
    }
    // device-resident problems cannot be inspected here: one CTA per window unless the caller asked for clusters
    FLV_CUDA(ctx, launch_ba(a, n_streams, ctx->ba_cluster_device > 0 ? ctx->ba_cluster_device : 1, smem, stream));
    ctx->launches++;
    FLV_CUDA(ctx, cudaGetLastError());
    return FLV_OK;
  }
  int C = 1;
  bool big = false, pose_only = pose_only_fast();
  for (int s = 0; s < n_streams; ++s) {
    if (!problems[s].fix_landmarks && problems[s].n_poses >= 3) C = ba_cluster_size(ctx);
    if (problems[s].n_poses > BA_MAX_FREE + 1) big = true;       // (a fixed pose may not exist: be conservative)
    if (problems[s].n_poses != 1 || !problems[s].fix_landmarks) pose_only = false;
  }
  for (int s = 0; s < n_streams; ++s) {
    const flv_ba_problem& p = problems[s];
    if (p.n_poses < 1 || p.n_poses > MP || p.n_landmarks < 0 || p.n_landmarks > ML || p.n_edges < 0 || p.n_edges > ME)
      FLV_FAIL(ctx, FLV_ERR_INVALID, "stream %d: problem (P=%d L=%d E=%d) exceeds reserved (%d,%d,%d)", s, p.n_poses,
               p.n_landmarks, p.n_edges, MP, ML, ME);
  }
  // staging layout: poses | lms | uv | ep | el | active
  const size_t b_pose = S * MP * 7 * 8, b_lm = S * ML * 3 * 8, b_uv = S * ME * 2 * 8, b_i = S * ME * 4, b_a = S * ME;
  const size_t o_pose = 0, o_lm = o_pose + b_pose, o_uv = o_lm + b_lm, o_ep = o_uv + b_uv, o_el = o_ep + b_i,
               o_act = o_el + b_i, total = ((o_act + b_a + 255) & ~(size_t)255);
  int rc = flv_stage_reserve(ctx, total);
  if (rc) return rc;
  char* hs = (char*)ctx->h_stage; char* ds = (char*)ctx->d_stage;
  memcpy(hs + o_pose, poses, b_pose); memcpy(hs + o_lm, landmarks, b_lm); memcpy(hs + o_uv, edge_uv, b_uv);
  memcpy(hs + o_ep, edge_pose, b_i); memcpy(hs + o_el, edge_lm, b_i); memcpy(hs + o_act, edge_active, b_a);
  FLV_CUDA(ctx, cudaMemcpyAsync(ds, hs, total, cudaMemcpyHostToDevice, stream));
  FLV_CUDA(ctx, cudaMemcpyAsync(d_prob, problems, S * sizeof(flv_ba_problem), cudaMemcpyHostToDevice, stream));
  a.problems = d_prob; a.poses = (double*)(ds + o_pose); a.lms = (double*)(ds + o_lm); a.uv = (const double*)(ds + o_uv);
  a.ep = (const int*)(ds + o_ep); a.el = (const int*)(ds + o_el); a.active = (uint8_t*)(ds + o_act); a.stats = d_stats;
  if (pose_only)
    ba_pose_only_kernel<<<n_streams, PO_THREADS, 0, stream>>>(a);
  else if (big)
    FLV_CUDA(ctx, flv_ba_big_launch(n_streams, a.problems, prm, a.poses, a.lms, a.ep, a.el, a.uv, a.active, a.stats, MP, ML, ME, a.ws, a.trace, stream));
  else
    FLV_CUDA(ctx, launch_ba(a, n_streams, C, smem, stream));
  ctx->launches++;
  FLV_CUDA(ctx, cudaGetLastError());
  FLV_CUDA(ctx, cudaMemcpyAsync(hs + o_pose, ds + o_pose, b_pose + b_lm, cudaMemcpyDeviceToHost, stream));
  FLV_CUDA(ctx, cudaMemcpyAsync(hs + o_act, ds + o_act, b_a, cudaMemcpyDeviceToHost, stream));
  FLV_CUDA(ctx, cudaMemcpyAsync(stats, d_stats, S * sizeof(flv_ba_stats), cudaMemcpyDeviceToHost, stream));
  FLV_CUDA(ctx, cudaStreamSynchronize(stream));
  memcpy(poses, hs + o_pose, b_pose); memcpy(landmarks, hs + o_lm, b_lm); memcpy(edge_active, hs + o_act, b_a);
  for (int s = 0; s < n_streams; ++s)
    if (stats[s].reserved)
      FLV_FAIL(ctx, FLV_ERR_UNSUPPORTED, "stream %d: %s", s,
               stats[s].reserved == 1 ? "pose count outside [1,32]" : stats[s].reserved == 2 ? "more than 24 free poses (reduced system > 144)" : stats[s].reserved == 3 ? "pose-pair list capacity exceeded" : stats[s].reserved == 5 ? "not a pose-only problem" : "window too large for the shared-memory landmark chunks");
  return FLV_OK;
}

/* CTAs (thread-block cluster size) per window: 1, 2 or 4.  host_mode_cluster applies to FLV_MEM_HOST calls whose problems
 * optimise landmarks (default 4; pose-only problems always use 1); device_mode_cluster to FLV_MEM_DEVICE calls (default 1). */
int flv_set_ba_cluster(flv_ctx* ctx, int host_mode_cluster, int device_mode_cluster) {
  if (!ctx) return FLV_ERR_INVALID;
  auto ok = [](int v) { return v == 1 || v == 2 || v == 4; };
  if (!ok(host_mode_cluster) || !ok(device_mode_cluster)) FLV_FAIL(ctx, FLV_ERR_INVALID, "cluster size must be 1, 2 or 4");
  ctx->ba_cluster = host_mode_cluster; ctx->ba_cluster_device = device_mode_cluster;
  return FLV_OK;
}

int flv_set_ba_stream(flv_ctx* ctx, void* cuda_stream, int enable) {
  if (!ctx) return FLV_ERR_INVALID;
  ctx->ba_stream = (cudaStream_t)cuda_stream;
  ctx->ba_stream_set = enable ? 1 : 0;
  return FLV_OK;
}

/* debug: cycle counters of the last flv_ba_optimize for `stream` (8 values, see Sh::prof) */
int flv_ba_profile(flv_ctx* ctx, int stream, long long* out16) {
  if (!ctx || !ctx->ba_ws || !out16 || stream < 0 || stream >= ctx->S) return FLV_ERR_INVALID;
  const size_t stride = ctx->ba_ws_stride;
  unsigned char* tail = (unsigned char*)ctx->ba_ws + stride * ctx->S;
  long long* d = (long long*)(tail + (size_t)ctx->S * (sizeof(flv_ba_problem) + sizeof(flv_ba_stats)));
  FLV_CUDA(ctx, cudaDeviceSynchronize());
  FLV_CUDA(ctx, cudaMemcpy(out16, d + 16 * stream, 128, cudaMemcpyDeviceToHost));
  return FLV_OK;
}

/* debug: per-iteration LM trace of the last flv_ba_optimize for `stream`: out[it][4] = {robust chi2 after the iteration,
 * lambda after it, rho of its last trial, trials}; returns the rows written (<= cap_iters, <= 32). */
int flv_ba_trace(flv_ctx* ctx, int stream, double* out, int cap_iters, int iterations_run) {
  if (!ctx || !ctx->ba_ws || !out || stream < 0 || stream >= ctx->S || cap_iters < 0) return FLV_ERR_INVALID;
  const size_t stride = ctx->ba_ws_stride;
  unsigned char* tail = (unsigned char*)ctx->ba_ws + stride * ctx->S;
  const double* d = (const double*)(tail + (size_t)ctx->S * (sizeof(flv_ba_problem) + sizeof(flv_ba_stats) + 128)) + (size_t)BA_TRACE_ITERS * 4 * stream;
  int n = iterations_run < BA_TRACE_ITERS ? iterations_run : BA_TRACE_ITERS;
  n = n < cap_iters ? n : cap_iters;
  FLV_CUDA(ctx, cudaDeviceSynchronize());
  if (n > 0) FLV_CUDA(ctx, cudaMemcpy(out, d, (size_t)n * 32, cudaMemcpyDeviceToHost));
  return n;
}

/* debug: the kernel's own residual and analytic Jacobians (EdgeSE3ProjectXYZ::computeError / linearizeOplus as ba_kernel
 * evaluates them) for n independent (pose, point, measurement) triples: r[n][2], A[n][2][3] = dr/dpoint,
 * B[n][2][6] = dr/dpose (rotation first).  Host arrays; synchronises. */
int flv_ba_debug_edges(flv_ctx* ctx, int n, const double* poses7, const double* pts3, const double* uv2, const double* K4,
                       double* r, double* A, double* B) {
  if (!ctx || n < 1 || !poses7 || !pts3 || !uv2 || !K4 || !r || !A || !B) return FLV_ERR_INVALID;
  const size_t in_b = (size_t)n * (7 + 3 + 2) * 8, out_b = (size_t)n * (2 + 6 + 12) * 8;
  int rc = flv_stage_reserve(ctx, in_b + out_b + 64);
  if (rc) return rc;
  double* hs = (double*)ctx->h_stage; double* ds = (double*)ctx->d_stage;
  memcpy(hs, poses7, (size_t)n * 56); memcpy(hs + 7 * (size_t)n, pts3, (size_t)n * 24); memcpy(hs + 10 * (size_t)n, uv2, (size_t)n * 16);
  FLV_CUDA(ctx, cudaMemcpyAsync(ds, hs, in_b, cudaMemcpyHostToDevice, ctx->stream));
  const Cam cam = {K4[0], K4[1], K4[2], K4[3]};
  ba_debug_edges_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(n, ds, ds + 7 * (size_t)n, ds + 10 * (size_t)n, cam, ds + 12 * (size_t)n);
  ctx->launches++;
  FLV_CUDA(ctx, cudaGetLastError());
  FLV_CUDA(ctx, cudaMemcpyAsync(hs + 12 * (size_t)n, ds + 12 * (size_t)n, out_b, cudaMemcpyDeviceToHost, ctx->stream));
  FLV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const double* o = hs + 12 * (size_t)n;
  for (int i = 0; i < n; ++i) {
    memcpy(r + 2 * i, o + 20 * (size_t)i, 16); memcpy(A + 6 * i, o + 20 * (size_t)i + 2, 48); memcpy(B + 12 * i, o + 20 * (size_t)i + 8, 96);
  }
  return FLV_OK;
}

}  // extern "C"
