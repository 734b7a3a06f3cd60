// Internal types of the batched device-resident tracker (tracker.cu: stage kernels, batch_tracker.cu: host orchestration).
#pragma once
#include "ctx.h"
#include "../host/undistort.h"

constexpr int TRK_THREADS = 512;      // one thread per landmark; ctx->max_pts must be <= 512

// one frame's landmark list per stream, structure of arrays with stride max_pts (LandMarkInFrame, landmark.h:8-36)
struct TrkTable {
  long long* id;
  double *plane, *undist, *p3w, *p3c, *f2d, *fpose;
  unsigned char *has, *inl;
  int* n;        // [S]
  double* T;     // [S][7] T_c_w of the frame
};

// per-stream control block written by the host before every frame
struct TrkCtl {
  int mode;                 // 0 idle, 1 track (state Tracking), 2 initialise (init_frame)
  int use_guess;            // IMU pose guess available (viGetCorrFrameState)
  int rp_found;             // has_imu && viGetIMURollPitchAtTime found a state
  int commit_on_init_fail;  // UnInit: the frame object is kept even when init_frame fails (it is simply overwritten later)
  double guess[7], init_pose[7];
  double roll, pitch;
};

// per-stream summary read back after every frame
struct TrkOut {
  int ok, fail_stage;       // fail_stage: 1 optical flow < 10, 2 F inliers < 10, 3 PnP inliers < 10, 4 / 5 pose-only BA
  int of_cnt, f_cnt, pnp_cnt, n_final, valid_cnt, rand_used, n_new, committed, mode, reserved;
  double T[7];
  double reproj_err;
  long long id_index;
};

struct TrkCam {
  flv_camera c;             // cam0 fx fy cx cy, P0, P1, cam_type (0 depth / 1 stereo), depth_scale
  double cam1[4];           // K1 (rectified)
  double T_c1_c0[7], T_i_c[7], T_c_i[7];
  double vi_para2;
  int w, h, unrect;
  flv::LensModel lens0, lens1;
};

struct TrkBufs {
  int max_pts;
  TrkCtl* ctl; TrkOut* out;
  int* ok; long long* id_index; int* orig_size;
  // frame -> frame LK
  float *lk_prev, *lk_init, *lk_next, *lk_err; unsigned char* lk_st; int* n_lk;
  // fundamental-matrix test
  float *fa, *fb; int* n_f; unsigned char* maskF; double* Fm; int* f_ninl;
  // PnP
  float *pnp3, *pnp2; int* n_pnp; double *pnp_K4, *pnp_Tin, *pnp_Tout; unsigned char* maskP; int* pnp_ninl;
  // pose-only BA (strides = the context's reserved sizes)
  int ba_MP, ba_ML, ba_ME;
  flv_ba_problem* ba_prob; double *ba_poses, *ba_lms, *ba_uv; int *ba_ep, *ba_el; unsigned char* ba_act; flv_ba_stats* ba_stats;
  // reprojection test
  int* n_rep; double* rep_mean;
  int* n_exist;
  // left -> right LK / depth samples / depth innovation
  float *r_prev, *r_init, *r_next, *r_err; unsigned char* r_st; int* n_r;
  double* pt1; unsigned short* dat; float* rnd; int* n_rand_used;
};

struct TrkDev {
  TrkTable L, C;
  TrkBufs b;
  TrkCam cam;
  unsigned short* depth;    // [S][h][w] current depth images (DEPTH_D435)
};

int flv_trk_stage_prepare(flv_ctx* ctx, const TrkDev& d, int S);
int flv_trk_stage_keep(flv_ctx* ctx, const TrkDev& d, int S);
int flv_trk_stage_after_f(flv_ctx* ctx, const TrkDev& d, int S);
int flv_trk_stage_after_pnp(flv_ctx* ctx, const TrkDev& d, int S);
int flv_trk_stage_after_ba(flv_ctx* ctx, const TrkDev& d, int S);
int flv_trk_stage_erase_outliers(flv_ctx* ctx, const TrkDev& d, int S);
int flv_trk_stage_append(flv_ctx* ctx, const TrkDev& d, int S, int mode_sel);
int flv_trk_stage_pt1(flv_ctx* ctx, const TrkDev& d, int S);
int flv_trk_stage_finish(flv_ctx* ctx, const TrkDev& d, int S);
