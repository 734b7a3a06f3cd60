// Internal context of libflvis_b200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/flvis_b200.h"

struct LevelGeom {
  int w, h, pitch;   // pitch in bytes (multiple of 128)
  size_t off;        // byte offset of the level inside one stream's pyramid block
};

struct PyrGeom {
  int nlev;
  LevelGeom lv[FLV_MAX_LEVELS];
  size_t stream_stride;  // bytes per stream per slot
};

struct flv_ctx {
  int device;
  int S;         // max streams
  int w, h;
  int max_pts;
  PyrGeom geom;
  uint8_t* pyr[FLV_NUM_SLOTS];   // each S * geom.stream_stride bytes
  cudaStream_t stream;
  bool own_stream;
  long long launches;
  int l1_valid[FLV_NUM_SLOTS];   // level 1 of the slot was produced by the fused ingest kernel (consumed by build_pyramid)
  unsigned* deriv[FLV_NUM_SLOTS];      // Scharr derivative pyramids (LK v4), allocated on first use
  int deriv_streams[FLV_NUM_SLOTS];    // streams for which deriv[slot] matches the slot's current images (0 = stale)
  // GFTT started early on an auxiliary stream by flv_feature_prepare, consumed by the next detect/redetect
  cudaStream_t aux_stream; cudaEvent_t ev_fork, ev_gftt;
  int prep_valid, prep_slot, prep_streams, prep_ncorn, prep_dis; double prep_ql;
  void* d_color_stage; size_t color_stage_bytes;   // landing area of interleaved colour uploads
  int equalize; int* d_hist;     // cv::equalizeHist on ingest (flv_set_equalize_hist)
  void* d_lut;                   // [S][256] equalisation LUTs of the fused ingest
  void* d_lk_tmaps;                        // LK TMA variant: tensor maps of the pyramid levels of every image slot
  int attr_lk3, attr_lk4, attr_region;     // cudaFuncSetAttribute done for this context's device
  int ba_dyn;                              // cached dynamic shared memory of ba_kernel (doubles)
  int ba_cluster, ba_cluster_device;       // CTAs per window for FLV_MEM_HOST / FLV_MEM_DEVICE solves (0 = default)
  int ba_member_buf;                       // ints of shared memory for TMA-staged member lists (-1 = read FLV_BA_TMA, 0 = off)
  int no_fused_ingest;           // FLV_NO_FUSED_INGEST=1: A/B switch for tests
  char err[512];

  // staging for FLV_MEM_HOST calls (pinned host + device mirrors)
  void* h_stage; size_t h_stage_bytes;
  void* d_stage; size_t d_stage_bytes;
  // host image uploads: ring of tight [S][h][w] landing areas filled by H2D copies on `copy_stream`, so the copy of the
  // next frame overlaps the kernels of the current one (the ROS image queue of the reference, tracking nodelet)
  static constexpr int IMG_RING = 4;
  void* d_img_stage[IMG_RING]; size_t img_stage_bytes;
  cudaEvent_t img_ready[IMG_RING], img_free[IMG_RING];
  cudaStream_t copy_stream;
  unsigned img_ring_pos;

  // LK
  int* d_npts;                    // [S]
  // GFTT
  int gftt_cap;                   // max corners out
  int cand_cap;                   // max NMS candidates per stream
  float* d_eig;                   // [S][h][w]
  int* d_eigmax;                  // [S] float bits (non-negative) via atomicMax
  unsigned long long* d_cand;     // [S][cand_cap] key = val_bits<<32 | addr
  int* d_ncand;                   // [S]
  unsigned long long* d_sorted;   // [S][cand_cap] slow path: keys sorted by priority
  int* d_items;                   // [S][cand_cap] slow path: ranks grouped by cell
  unsigned char* d_state;         // [S][cand_cap] slow path: greedy state
  int* d_need_full;               // [S] fast path could not decide -> slow path runs
  int keep_eig;                   // debug: also write the f32 response map
  float* d_corners;               // [S][gftt_cap][2]
  int* d_ncorners;                // [S]
  int* d_flags;                   // [S] overflow flags
  int max_cells;
  // FeatureDEM
  double* d_exist;                // [S][max_pts][2]
  int* d_nexist;                  // [S]
  float* d_newxy;                 // [S][max_pts][2]
  int* d_nnew;                    // [S]
  // BA
  int ba_max_poses, ba_max_lms, ba_max_edges;
  void* ba_ws;                    // device workspace
  cudaStream_t ba_stream;         // optional separate stream for flv_ba_optimize (local-map thread analogue)
  int ba_stream_set;
  size_t ba_ws_bytes;
  size_t ba_ws_stride;            // bytes of workspace per stream slot (max of the ba.cu and ba_big.cu layouts)
};

#define FLV_CUDA(ctx, call)                                                            \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess) {                                                          \
      snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d %s: %s", __FILE__, __LINE__, #call, \
               cudaGetErrorString(e__));                                               \
      return FLV_ERR_CUDA;                                                             \
    }                                                                                  \
  } while (0)

#define FLV_FAIL(ctx, code, ...)                          \
  do {                                                    \
    snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); \
    return (code);                                        \
  } while (0)

// staging helpers (capi.cu)
int flv_stage_reserve(flv_ctx* ctx, size_t bytes);

// kernel launchers (one per .cu)
int flv_launch_pyramid(flv_ctx* ctx, int slot, int n_streams);
int flv_launch_fmat_ransac(flv_ctx* ctx, int n_streams, const int* d_npts, const float* d_from, const float* d_to, double thr_px,
                           uint8_t* d_mask, double* d_F, int* d_ninl);
int flv_launch_pnp_ransac(flv_ctx* ctx, int n_streams, const int* d_npts, const float* d_p3d, const float* d_p2d, const double* d_K4,
                          const double* d_Tin, double thr_px, double* d_Tout, uint8_t* d_mask, int* d_ninl);
int flv_launch_gray(flv_ctx* ctx, int n_streams, const uint8_t* d_src, size_t row_stride, size_t img_stride, int channels, int rgb,
                    uint8_t* d_dst_tight);
int flv_launch_equalize(flv_ctx* ctx, int n_streams, const uint8_t* d_src, size_t row_stride, size_t img_stride, uint8_t* d_dst_tight);
int flv_launch_unpack(flv_ctx* ctx, int slot, int n_streams, const uint8_t* d_src, size_t row_stride, size_t img_stride);
int flv_launch_unpack_equalized(flv_ctx* ctx, int slot, int n_streams, const uint8_t* d_src, size_t row_stride, size_t img_stride);
int flv_launch_lk(flv_ctx* ctx, int src_slot, int dst_slot, int n_streams, const int* d_npts,
                  const float* d_prev, const float* d_init, float* d_next, uint8_t* d_status,
                  float* d_err, int nlev_used, int max_iter, double eps2, double min_eig_thr);
int flv_launch_gftt(flv_ctx* ctx, int slot, int n_streams, int max_corners, double quality,
                    double min_distance);
int flv_launch_region(flv_ctx* ctx, int slot, int n_streams, const flv_feature_params* prm,
                      int redetect);
int flv_launch_select(flv_ctx* ctx, int n_streams, const int* d_npts, const float* prev, const float* next,
                      const uint8_t* status, uint8_t* keep, float* out, double* out64);
int flv_gftt_init(flv_ctx* ctx);
size_t flv_mindist_smem(int cand_cap, int max_cells);
int flv_ba_free(flv_ctx* ctx);
