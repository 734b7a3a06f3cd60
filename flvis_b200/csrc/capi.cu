// C ABI of libflvis_b200: context management, host<->device staging, argument checking.
// The entry points are declared (with the reference call each replaces) in include/flvis_b200.h.
#include <new>
#include <stdlib.h>
#include "ctx.h"

namespace {

constexpr int WIN = 31;

void build_geom(PyrGeom& g, int w, int h) {
  // cv::buildOpticalFlowPyramid: stop before a level whose width or height is <= winSize (31)
  g.nlev = 0;
  size_t off = 0;
  int lw = w, lh = h;
  for (int l = 0; l < FLV_MAX_LEVELS; ++l) {
    if (l > 0) {
      lw = (lw + 1) / 2; lh = (lh + 1) / 2;
      if (lw <= WIN || lh <= WIN) break;
    }
    LevelGeom& L = g.lv[l];
    L.w = lw; L.h = lh; L.pitch = (lw + 127) & ~127; L.off = off;
    off += (size_t)L.pitch * lh;
    g.nlev = l + 1;
  }
  for (int l = g.nlev; l < FLV_MAX_LEVELS; ++l) g.lv[l] = g.lv[g.nlev - 1];
  g.stream_stride = (off + 255) & ~(size_t)255;
}

template <class T>
int dev_alloc(flv_ctx* ctx, T** p, size_t count) {
  FLV_CUDA(ctx, cudaMalloc((void**)p, count * sizeof(T)));
  FLV_CUDA(ctx, cudaMemset(*p, 0, count * sizeof(T)));
  return FLV_OK;
}

}  // namespace

int flv_stage_reserve(flv_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->h_stage_bytes) return FLV_OK;
  bytes = (bytes + (1 << 20)) & ~(size_t)((1 << 20) - 1);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->d_stage) cudaFree(ctx->d_stage);
  ctx->h_stage = nullptr; ctx->d_stage = nullptr; ctx->h_stage_bytes = ctx->d_stage_bytes = 0;
  FLV_CUDA(ctx, cudaMallocHost(&ctx->h_stage, bytes));
  FLV_CUDA(ctx, cudaMalloc(&ctx->d_stage, bytes));
  ctx->h_stage_bytes = ctx->d_stage_bytes = bytes;
  return FLV_OK;
}

extern "C" {

const char* flv_version(void) { return "flvis_b200 0.1 (sm_100a)"; }

int flv_create(flv_ctx** out, int device, int max_streams, int img_w, int img_h, int max_pts) {
  if (!out || max_streams < 1 || img_w < 64 || img_h < 64 || img_w > 65535 || img_h > 65535 ||
      max_pts < 1)
    return FLV_ERR_INVALID;
  flv_ctx* ctx = new (std::nothrow) flv_ctx();
  if (!ctx) return FLV_ERR_NOMEM;
  memset(ctx, 0, sizeof(*ctx));
  *out = ctx;
  ctx->device = device; ctx->S = max_streams; ctx->w = img_w; ctx->h = img_h; ctx->max_pts = max_pts;
  ctx->ba_member_buf = -1;
  FLV_CUDA(ctx, cudaSetDevice(device));
  // FLV_BLOCKING_SYNC=1: host threads that wait for the GPU sleep instead of spinning (cudaDeviceScheduleBlockingSync).  Every
  // stream group, local-map shard and caller thread waits on the device most of the time; with several ranks per node that is
  // more spinning threads than cores (bench.py sets it when ranks x threads exceed the core count).  Best effort.
  if (const char* e = getenv("FLV_BLOCKING_SYNC")) {
    if (atoi(e) > 0 && cudaSetDeviceFlags(cudaDeviceScheduleBlockingSync) != cudaSuccess) cudaGetLastError();
  }
  build_geom(ctx->geom, img_w, img_h);
  {   // cv::buildOpticalFlowPyramid would build a further level for this size: LK would silently lose its coarsest level
    const LevelGeom& top = ctx->geom.lv[ctx->geom.nlev - 1];
    if (ctx->geom.nlev == FLV_MAX_LEVELS && (top.w + 1) / 2 > 31 && (top.h + 1) / 2 > 31)
      FLV_FAIL(ctx, FLV_ERR_UNSUPPORTED, "%dx%d images need more than %d pyramid levels (31x31 window); this build supports %d", img_w,
               img_h, FLV_MAX_LEVELS, FLV_MAX_LEVELS);
  }
  ctx->no_fused_ingest = getenv("FLV_NO_FUSED_INGEST") ? atoi(getenv("FLV_NO_FUSED_INGEST")) : 0;
  FLV_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  ctx->own_stream = true;
  const size_t S = max_streams;
  for (int i = 0; i < FLV_NUM_SLOTS; ++i) {
    int rc = dev_alloc(ctx, &ctx->pyr[i], S * ctx->geom.stream_stride);
    if (rc) return rc;
  }
  ctx->gftt_cap = 4096;
  ctx->cand_cap = 65536;
  ctx->max_cells = 8191;   // cstart[max_cells+1] + ccount[max_cells] must fit the 64 KB half of the fast path's region A
  int rc = 0;
  rc |= dev_alloc(ctx, &ctx->d_npts, S);
  rc |= dev_alloc(ctx, &ctx->d_eig, S * img_w * img_h);
  rc |= dev_alloc(ctx, &ctx->d_eigmax, S);
  rc |= dev_alloc(ctx, &ctx->d_cand, S * ctx->cand_cap);
  rc |= dev_alloc(ctx, &ctx->d_ncand, S);
  rc |= dev_alloc(ctx, &ctx->d_sorted, S * ctx->cand_cap);
  rc |= dev_alloc(ctx, &ctx->d_items, S * ctx->cand_cap);
  rc |= dev_alloc(ctx, &ctx->d_state, S * ctx->cand_cap);
  rc |= dev_alloc(ctx, &ctx->d_need_full, S);
  rc |= dev_alloc(ctx, &ctx->d_corners, S * ctx->gftt_cap * 2);
  rc |= dev_alloc(ctx, &ctx->d_ncorners, S);
  rc |= dev_alloc(ctx, &ctx->d_flags, S);
  rc |= dev_alloc(ctx, &ctx->d_exist, S * max_pts * 2);
  rc |= dev_alloc(ctx, &ctx->d_nexist, S);
  rc |= dev_alloc(ctx, &ctx->d_newxy, S * max_pts * 2);
  rc |= dev_alloc(ctx, &ctx->d_nnew, S);
  if (rc) return FLV_ERR_CUDA;
  rc = flv_gftt_init(ctx);
  if (rc) return rc;
  rc = flv_stage_reserve(ctx, (size_t)4 << 20);
  if (rc) return rc;
  return FLV_OK;
}

void flv_destroy(flv_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < FLV_NUM_SLOTS; ++i) { cudaFree(ctx->pyr[i]); if (ctx->deriv[i]) cudaFree(ctx->deriv[i]); }
  cudaFree(ctx->d_npts); cudaFree(ctx->d_eig); cudaFree(ctx->d_eigmax); cudaFree(ctx->d_cand);
  cudaFree(ctx->d_ncand); cudaFree(ctx->d_sorted); cudaFree(ctx->d_items); cudaFree(ctx->d_state); cudaFree(ctx->d_need_full); cudaFree(ctx->d_corners); cudaFree(ctx->d_ncorners);
  cudaFree(ctx->d_flags); cudaFree(ctx->d_exist); cudaFree(ctx->d_nexist); cudaFree(ctx->d_newxy);
  cudaFree(ctx->d_nnew);
  flv_ba_free(ctx);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  if (ctx->d_stage) cudaFree(ctx->d_stage);
  if (ctx->img_stage_bytes) for (int i = 0; i < flv_ctx::IMG_RING; ++i) cudaFree(ctx->d_img_stage[i]);
  if (ctx->d_hist) cudaFree(ctx->d_hist);
  if (ctx->d_lut) cudaFree(ctx->d_lut);
  if (ctx->d_lk_tmaps) cudaFree(ctx->d_lk_tmaps);
  if (ctx->d_color_stage) cudaFree(ctx->d_color_stage);
  if (ctx->aux_stream) { cudaEventDestroy(ctx->ev_fork); cudaEventDestroy(ctx->ev_gftt); cudaStreamDestroy(ctx->aux_stream); }
  if (ctx->copy_stream) {
    for (int i = 0; i < flv_ctx::IMG_RING; ++i) { cudaEventDestroy(ctx->img_ready[i]); cudaEventDestroy(ctx->img_free[i]); }
    cudaStreamDestroy(ctx->copy_stream);
  }
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int flv_set_stream(flv_ctx* ctx, void* cuda_stream) {
  if (!ctx) return FLV_ERR_INVALID;
  if (ctx->own_stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
  if (cuda_stream) { ctx->stream = (cudaStream_t)cuda_stream; ctx->own_stream = false; }
  else {
    FLV_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  return FLV_OK;
}

int flv_sync(flv_ctx* ctx) {
  if (!ctx) return FLV_ERR_INVALID;
  FLV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FLV_OK;
}

const char* flv_last_error(flv_ctx* ctx) { return ctx ? ctx->err : "null context"; }
long long flv_launch_count(flv_ctx* ctx) { return ctx ? ctx->launches : 0; }
int flv_num_levels(flv_ctx* ctx) { return ctx ? ctx->geom.nlev : 0; }
int flv_gftt_capacity(flv_ctx* ctx) { return ctx ? ctx->gftt_cap : 0; }
int flv_gftt_keep_response(flv_ctx* ctx, int enable) {
  if (!ctx) return FLV_ERR_INVALID;
  ctx->keep_eig = enable ? 1 : 0;
  return FLV_OK;
}

int flv_level_info(flv_ctx* ctx, int level, int* w, int* h, int* pitch, size_t* offset) {
  if (!ctx || level < 0 || level >= ctx->geom.nlev) return FLV_ERR_INVALID;
  const LevelGeom& L = ctx->geom.lv[level];
  if (w) *w = L.w; if (h) *h = L.h; if (pitch) *pitch = L.pitch; if (offset) *offset = L.off;
  return FLV_OK;
}

int flv_upload_images(flv_ctx* ctx, int slot, int n_streams, const uint8_t* imgs,
                      size_t row_stride_bytes, size_t img_stride_bytes, flv_memspace mem) {
  if (!ctx || !imgs || slot < 0 || slot >= FLV_NUM_SLOTS || n_streams < 1 || n_streams > ctx->S ||
      row_stride_bytes < (size_t)ctx->w)
    return FLV_ERR_INVALID;
  if (mem == FLV_MEM_DEVICE && !ctx->equalize) return flv_launch_unpack(ctx, slot, n_streams, imgs, row_stride_bytes, img_stride_bytes);
  if (mem == FLV_MEM_DEVICE) {       // equalizeHist fused into the ingest: no landing area at all
    const int rc = flv_launch_unpack_equalized(ctx, slot, n_streams, imgs, row_stride_bytes, img_stride_bytes);
    if (rc != FLV_ERR_UNSUPPORTED) return rc;
  }
  // host images: one (2D) H2D copy of all streams into a tight device landing area on the copy stream, then one unpack
  // launch on the compute stream.  The landing areas form a ring, so a caller that submits frame k+1 before it waits for
  // the results of frame k gets the copy of k+1 overlapped with the kernels of k.
  const size_t w = ctx->w, h = ctx->h, bytes = (size_t)n_streams * w * h;
  if (!ctx->copy_stream) {
    FLV_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < flv_ctx::IMG_RING; ++i) {
      FLV_CUDA(ctx, cudaEventCreateWithFlags(&ctx->img_ready[i], cudaEventDisableTiming));
      FLV_CUDA(ctx, cudaEventCreateWithFlags(&ctx->img_free[i], cudaEventDisableTiming));
    }
  }
  if (ctx->img_stage_bytes == 0) {
    for (int i = 0; i < flv_ctx::IMG_RING; ++i) FLV_CUDA(ctx, cudaMalloc(&ctx->d_img_stage[i], (size_t)ctx->S * w * h));
    ctx->img_stage_bytes = (size_t)ctx->S * w * h;
  }
  const int ring = (int)(ctx->img_ring_pos++ % flv_ctx::IMG_RING);
  uint8_t* st = (uint8_t*)ctx->d_img_stage[ring];
  cudaStream_t cs = ctx->copy_stream;
  if (mem == FLV_MEM_DEVICE) {       // equalizeHist of device-resident images: the landing area receives the equalized copy
    FLV_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->img_free[ring], 0));
    int rc = flv_launch_equalize(ctx, n_streams, imgs, row_stride_bytes, img_stride_bytes, st);
    if (rc) return rc;
    rc = flv_launch_unpack(ctx, slot, n_streams, st, w, w * h);
    if (rc) return rc;
    FLV_CUDA(ctx, cudaEventRecord(ctx->img_free[ring], ctx->stream));
    return FLV_OK;
  }
  FLV_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->img_free[ring], 0));        // last unpack that read this landing area
  if (row_stride_bytes == w && img_stride_bytes == w * h) {
    // in pieces: a copy engine serves one request at a time, and small copies of other streams (BA windows, results)
    // should not queue behind 11 MB of pixels
    const size_t piece = (size_t)2 << 20;
    for (size_t o = 0; o < bytes; o += piece)
      FLV_CUDA(ctx, cudaMemcpyAsync(st + o, imgs + o, bytes - o < piece ? bytes - o : piece, cudaMemcpyHostToDevice, cs));
  } else if (img_stride_bytes == row_stride_bytes * h) {
    FLV_CUDA(ctx, cudaMemcpy2DAsync(st, w, imgs, row_stride_bytes, w, h * (size_t)n_streams, cudaMemcpyHostToDevice, cs));
  } else {
    for (int s = 0; s < n_streams; ++s)
      FLV_CUDA(ctx, cudaMemcpy2DAsync(st + (size_t)s * w * h, w, imgs + (size_t)s * img_stride_bytes, row_stride_bytes,
                                      w, h, cudaMemcpyHostToDevice, cs));
  }
  FLV_CUDA(ctx, cudaEventRecord(ctx->img_ready[ring], cs));
  FLV_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->img_ready[ring], 0));
  int rc = ctx->equalize ? flv_launch_unpack_equalized(ctx, slot, n_streams, st, w, w * h) : FLV_ERR_UNSUPPORTED;
  if (rc == FLV_ERR_UNSUPPORTED) {
    if (ctx->equalize) {
      const int rc0 = flv_launch_equalize(ctx, n_streams, st, w, w * h, st);          // in place (element-wise LUT)
      if (rc0) return rc0;
    }
    rc = flv_launch_unpack(ctx, slot, n_streams, st, w, w * h);
  }
  if (rc) return rc;
  FLV_CUDA(ctx, cudaEventRecord(ctx->img_free[ring], ctx->stream));
  return FLV_OK;
}

int flv_upload_color_images(flv_ctx* ctx, int slot, int n_streams, const uint8_t* imgs, size_t row_stride_bytes,
                            size_t img_stride_bytes, int channels, int is_rgb, flv_memspace mem) {
  if (!ctx || !imgs || slot < 0 || slot >= FLV_NUM_SLOTS || n_streams < 1 || n_streams > ctx->S || (channels != 3 && channels != 4) ||
      row_stride_bytes < (size_t)ctx->w * channels || img_stride_bytes < row_stride_bytes * ctx->h)
    return FLV_ERR_INVALID;
  const size_t w = ctx->w, h = ctx->h;
  const uint8_t* d_src = imgs;
  if (mem == FLV_MEM_HOST) {                         // one H2D copy of the interleaved frames, converted on the device
    const size_t bytes = (size_t)n_streams * img_stride_bytes;
    if (bytes > ctx->color_stage_bytes) {
      if (ctx->d_color_stage) cudaFree(ctx->d_color_stage);
      ctx->d_color_stage = nullptr; ctx->color_stage_bytes = 0;
      FLV_CUDA(ctx, cudaMalloc(&ctx->d_color_stage, bytes));
      ctx->color_stage_bytes = bytes;
    }
    FLV_CUDA(ctx, cudaMemcpyAsync(ctx->d_color_stage, imgs, bytes, cudaMemcpyHostToDevice, ctx->stream));
    d_src = (const uint8_t*)ctx->d_color_stage;
  }
  if (ctx->img_stage_bytes == 0) {
    for (int i = 0; i < flv_ctx::IMG_RING; ++i) FLV_CUDA(ctx, cudaMalloc(&ctx->d_img_stage[i], (size_t)ctx->S * w * h));
    ctx->img_stage_bytes = (size_t)ctx->S * w * h;
  }
  if (!ctx->copy_stream) {
    FLV_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < flv_ctx::IMG_RING; ++i) {
      FLV_CUDA(ctx, cudaEventCreateWithFlags(&ctx->img_ready[i], cudaEventDisableTiming));
      FLV_CUDA(ctx, cudaEventCreateWithFlags(&ctx->img_free[i], cudaEventDisableTiming));
    }
  }
  const int ring = (int)(ctx->img_ring_pos++ % flv_ctx::IMG_RING);
  uint8_t* st = (uint8_t*)ctx->d_img_stage[ring];
  FLV_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->img_free[ring], 0));
  int rc = flv_launch_gray(ctx, n_streams, d_src, row_stride_bytes, img_stride_bytes, channels, is_rgb ? 1 : 0, st);
  if (rc) return rc;
  rc = ctx->equalize ? flv_launch_unpack_equalized(ctx, slot, n_streams, st, w, w * h) : FLV_ERR_UNSUPPORTED;
  if (rc == FLV_ERR_UNSUPPORTED) {
    if (ctx->equalize && (rc = flv_launch_equalize(ctx, n_streams, st, w, w * h, st))) return rc;
    rc = flv_launch_unpack(ctx, slot, n_streams, st, w, w * h);
  }
  if (rc) return rc;
  FLV_CUDA(ctx, cudaEventRecord(ctx->img_free[ring], ctx->stream));
  return FLV_OK;
}

int flv_set_equalize_hist(flv_ctx* ctx, int enable) {
  if (!ctx) return FLV_ERR_INVALID;
  ctx->equalize = enable ? 1 : 0;
  return FLV_OK;
}

int flv_build_pyramid(flv_ctx* ctx, int slot, int n_streams) {
  if (!ctx || slot < 0 || slot >= FLV_NUM_SLOTS || n_streams < 1 || n_streams > ctx->S)
    return FLV_ERR_INVALID;
  return flv_launch_pyramid(ctx, slot, n_streams);
}

int flv_download_level(flv_ctx* ctx, int slot, int stream, int level, uint8_t* out, flv_memspace mem) {
  if (!ctx || !out || slot < 0 || slot >= FLV_NUM_SLOTS || stream < 0 || stream >= ctx->S ||
      level < 0 || level >= ctx->geom.nlev)
    return FLV_ERR_INVALID;
  const LevelGeom& L = ctx->geom.lv[level];
  FLV_CUDA(ctx, cudaMemcpy2DAsync(out, L.w, ctx->pyr[slot] + (size_t)stream * ctx->geom.stream_stride + L.off,
                                  L.pitch, L.w, L.h,
                                  mem == FLV_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice,
                                  ctx->stream));
  if (mem == FLV_MEM_HOST) FLV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FLV_OK;
}

int flv_lk_track(flv_ctx* ctx, int src_slot, int dst_slot, int n_streams, const int* n_pts,
                 const float* prev_xy, const float* init_xy, float* next_xy, uint8_t* status,
                 float* err, const flv_lk_params* prm, flv_memspace mem) {
  if (!ctx || !prm || !n_pts || !prev_xy || !init_xy || !next_xy || !status || (!err && mem != FLV_MEM_DEVICE) ||
      src_slot < 0 || src_slot >= FLV_NUM_SLOTS || dst_slot < 0 || dst_slot >= FLV_NUM_SLOTS ||
      n_streams < 1 || n_streams > ctx->S)
    return FLV_ERR_INVALID;
  if (prm->win != WIN) FLV_FAIL(ctx, FLV_ERR_UNSUPPORTED, "LK window %d: only 31 is implemented", prm->win);
  if (prm->max_level < 0) return FLV_ERR_INVALID;
  int nlev_used = prm->max_level + 1 < ctx->geom.nlev ? prm->max_level + 1 : ctx->geom.nlev;
  int max_iter = prm->max_iter < 0 ? 0 : (prm->max_iter > 100 ? 100 : prm->max_iter);
  double eps = prm->eps < 0 ? 0 : (prm->eps > 10 ? 10 : prm->eps);
  const size_t np = (size_t)n_streams * ctx->max_pts;
  if (mem == FLV_MEM_DEVICE) {
    return flv_launch_lk(ctx, src_slot, dst_slot, n_streams, n_pts, prev_xy, init_xy, next_xy, status,
                         err, nlev_used, max_iter, eps * eps, prm->min_eig_threshold);
  }
  for (int s = 0; s < n_streams; ++s)
    if (n_pts[s] < 0 || n_pts[s] > ctx->max_pts) return FLV_ERR_INVALID;
  // layout of the staging block: npts | prev | init/next | err | status
  size_t o_np = 0, o_prev = 256 + (size_t)n_streams * 4, o_init = o_prev + np * 8, o_err = o_init + np * 8,
         o_st = o_err + np * 4, total = o_st + np;
  o_prev = (o_prev + 255) & ~(size_t)255;
  o_init = o_prev + np * 8; o_err = o_init + np * 8; o_st = o_err + np * 4; total = o_st + np;
  int rc = flv_stage_reserve(ctx, total);
  if (rc) return rc;
  char* hs = (char*)ctx->h_stage; char* ds = (char*)ctx->d_stage;
  memcpy(hs + o_np, n_pts, (size_t)n_streams * 4);
  memcpy(hs + o_prev, prev_xy, np * 8);
  memcpy(hs + o_init, init_xy, np * 8);
  FLV_CUDA(ctx, cudaMemcpyAsync(ds, hs, o_err, cudaMemcpyHostToDevice, ctx->stream));
  rc = flv_launch_lk(ctx, src_slot, dst_slot, n_streams, (const int*)(ds + o_np), (const float*)(ds + o_prev),
                     (const float*)(ds + o_init), (float*)(ds + o_init), (uint8_t*)(ds + o_st),
                     (float*)(ds + o_err), nlev_used, max_iter, eps * eps, prm->min_eig_threshold);
  if (rc) return rc;
  FLV_CUDA(ctx, cudaMemcpyAsync(hs + o_init, ds + o_init, total - o_init, cudaMemcpyDeviceToHost, ctx->stream));
  FLV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(next_xy, hs + o_init, np * 8);
  memcpy(err, hs + o_err, np * 4);
  memcpy(status, hs + o_st, np);
  return FLV_OK;
}

int flv_fundamental_ransac(flv_ctx* ctx, int n_streams, const int* n_pts, const float* from_xy, const float* to_xy,
                           const flv_ransac_params* prm, uint8_t* mask, double* F, int* n_inliers, flv_memspace mem) {
  if (!ctx || !n_pts || !from_xy || !to_xy || !prm || !mask || !F || !n_inliers || n_streams < 1 || n_streams > ctx->S ||
      !(prm->threshold_px > 0))
    return FLV_ERR_INVALID;
  if (mem == FLV_MEM_DEVICE) return flv_launch_fmat_ransac(ctx, n_streams, n_pts, from_xy, to_xy, prm->threshold_px, mask, F, n_inliers);
  for (int s = 0; s < n_streams; ++s)
    if (n_pts[s] < 0 || n_pts[s] > ctx->max_pts) return FLV_ERR_INVALID;
  const size_t np = (size_t)n_streams * ctx->max_pts, S = n_streams;
  // staging: npts | from | to || F | ninl | mask
  const size_t o_np = 0, o_from = 256 + ((S * 4 + 255) & ~(size_t)255), o_to = o_from + np * 8, o_F = o_to + np * 8,
               o_ni = o_F + S * 72, o_mask = o_ni + ((S * 4 + 255) & ~(size_t)255), total = o_mask + np;
  int rc = flv_stage_reserve(ctx, total);
  if (rc) return rc;
  char* hs = (char*)ctx->h_stage; char* ds = (char*)ctx->d_stage;
  memcpy(hs + o_np, n_pts, S * 4); memcpy(hs + o_from, from_xy, np * 8); memcpy(hs + o_to, to_xy, np * 8);
  FLV_CUDA(ctx, cudaMemcpyAsync(ds, hs, o_F, cudaMemcpyHostToDevice, ctx->stream));
  rc = flv_launch_fmat_ransac(ctx, n_streams, (const int*)(ds + o_np), (const float*)(ds + o_from), (const float*)(ds + o_to),
                              prm->threshold_px, (uint8_t*)(ds + o_mask), (double*)(ds + o_F), (int*)(ds + o_ni));
  if (rc) return rc;
  FLV_CUDA(ctx, cudaMemcpyAsync(hs + o_F, ds + o_F, total - o_F, cudaMemcpyDeviceToHost, ctx->stream));
  FLV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(F, hs + o_F, S * 72); memcpy(n_inliers, hs + o_ni, S * 4); memcpy(mask, hs + o_mask, np);
  return FLV_OK;
}

int flv_pnp_ransac(flv_ctx* ctx, int n_streams, const int* n_pts, const float* p3d, const float* p2d, const double* K4,
                   const double* T_c_w_in, const flv_ransac_params* prm, double* T_c_w_out, uint8_t* mask, int* n_inliers,
                   flv_memspace mem) {
  if (!ctx || !n_pts || !p3d || !p2d || !K4 || !T_c_w_in || !prm || !T_c_w_out || !mask || !n_inliers || n_streams < 1 ||
      n_streams > ctx->S || !(prm->threshold_px > 0))
    return FLV_ERR_INVALID;
  if (mem == FLV_MEM_DEVICE)
    return flv_launch_pnp_ransac(ctx, n_streams, n_pts, p3d, p2d, K4, T_c_w_in, prm->threshold_px, T_c_w_out, mask, n_inliers);
  for (int s = 0; s < n_streams; ++s)
    if (n_pts[s] < 0 || n_pts[s] > ctx->max_pts) return FLV_ERR_INVALID;
  const size_t np = (size_t)n_streams * ctx->max_pts, S = n_streams;
  // staging: npts | K | Tin | p3d | p2d || Tout | ninl | mask
  const size_t o_np = 0, o_K = 256 + ((S * 4 + 255) & ~(size_t)255), o_Ti = o_K + S * 32, o_p3 = (o_Ti + S * 56 + 255) & ~(size_t)255,
               o_p2 = o_p3 + np * 12, o_To = (o_p2 + np * 8 + 255) & ~(size_t)255, o_ni = o_To + S * 56,
               o_mask = (o_ni + S * 4 + 255) & ~(size_t)255, total = o_mask + np;
  int rc = flv_stage_reserve(ctx, total);
  if (rc) return rc;
  char* hs = (char*)ctx->h_stage; char* ds = (char*)ctx->d_stage;
  memcpy(hs + o_np, n_pts, S * 4); memcpy(hs + o_K, K4, S * 32); memcpy(hs + o_Ti, T_c_w_in, S * 56);
  memcpy(hs + o_p3, p3d, np * 12); memcpy(hs + o_p2, p2d, np * 8);
  FLV_CUDA(ctx, cudaMemcpyAsync(ds, hs, o_To, cudaMemcpyHostToDevice, ctx->stream));
  rc = flv_launch_pnp_ransac(ctx, n_streams, (const int*)(ds + o_np), (const float*)(ds + o_p3), (const float*)(ds + o_p2),
                             (const double*)(ds + o_K), (const double*)(ds + o_Ti), prm->threshold_px, (double*)(ds + o_To),
                             (uint8_t*)(ds + o_mask), (int*)(ds + o_ni));
  if (rc) return rc;
  FLV_CUDA(ctx, cudaMemcpyAsync(hs + o_To, ds + o_To, total - o_To, cudaMemcpyDeviceToHost, ctx->stream));
  FLV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(T_c_w_out, hs + o_To, S * 56); memcpy(n_inliers, hs + o_ni, S * 4); memcpy(mask, hs + o_mask, np);
  return FLV_OK;
}

int flv_select_tracked(flv_ctx* ctx, int n_streams, const int* n_pts, const float* prev_xy,
                       const float* next_xy, const uint8_t* status, uint8_t* keep, float* out_xy,
                       double* out_xy_f64) {
  if (!ctx || !n_pts || !prev_xy || !next_xy || !status || n_streams < 1 || n_streams > ctx->S)
    return FLV_ERR_INVALID;
  return flv_launch_select(ctx, n_streams, n_pts, prev_xy, next_xy, status, keep, out_xy, out_xy_f64);
}

static int check_flags(flv_ctx* ctx, int n_streams) {
  // host-memory calls synchronise anyway: surface device-side capacity overflows as an error
  int* hf = (int*)ctx->h_stage;
  FLV_CUDA(ctx, cudaMemcpyAsync(hf, ctx->d_flags, (size_t)n_streams * 4, cudaMemcpyDeviceToHost, ctx->stream));
  FLV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int s = 0; s < n_streams; ++s)
    if (hf[s]) {
      int f = hf[s];
      cudaMemsetAsync(ctx->d_flags, 0, (size_t)n_streams * 4, ctx->stream);
      FLV_FAIL(ctx, FLV_ERR_OVERFLOW, "stream %d: device capacity exceeded (flags 0x%x: 1=candidates 2=accepted "
               "4=region kept 8=sort depth 16=max_pts)", s, f);
    }
  return FLV_OK;
}

/* FLV_MEM_DEVICE callers never synchronise inside the library, so the per-stream overflow flags of the Shi-Tomasi / FeatureDEM
 * kernels stay on the device until somebody asks: copies them to flags_out[n_streams] (host), clears them, synchronises the
 * context's stream.  Returns FLV_ERR_OVERFLOW if any flag was set. */
int flv_get_flags(flv_ctx* ctx, int n_streams, int* flags_out) {
  if (!ctx || n_streams < 1 || n_streams > ctx->S) return FLV_ERR_INVALID;
  int rc = flv_stage_reserve(ctx, (size_t)n_streams * 4);
  if (rc) return rc;
  int* hf = (int*)ctx->h_stage;
  FLV_CUDA(ctx, cudaMemcpyAsync(hf, ctx->d_flags, (size_t)n_streams * 4, cudaMemcpyDeviceToHost, ctx->stream));
  FLV_CUDA(ctx, cudaMemsetAsync(ctx->d_flags, 0, (size_t)n_streams * 4, ctx->stream));
  FLV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int any = 0;
  for (int s = 0; s < n_streams; ++s) { if (flags_out) flags_out[s] = hf[s]; any |= hf[s]; }
  return any ? FLV_ERR_OVERFLOW : FLV_OK;
}

int flv_gftt(flv_ctx* ctx, int slot, int n_streams, int max_corners, double quality,
             double min_distance, float* xy_out, int* n_out, int out_stride_pts, flv_memspace mem) {
  if (!ctx || !xy_out || !n_out || slot < 0 || slot >= FLV_NUM_SLOTS || n_streams < 1 ||
      n_streams > ctx->S || out_stride_pts < max_corners)
    return FLV_ERR_INVALID;
  if (ctx->prep_valid) { FLV_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_gftt, 0)); ctx->prep_valid = 0; }
  int rc = flv_launch_gftt(ctx, slot, n_streams, max_corners, quality, min_distance);
  if (rc) return rc;
  cudaMemcpyKind kind = mem == FLV_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  FLV_CUDA(ctx, cudaMemcpy2DAsync(xy_out, (size_t)out_stride_pts * 8, ctx->d_corners, (size_t)ctx->gftt_cap * 8,
                                  (size_t)max_corners * 8, n_streams, kind, ctx->stream));
  FLV_CUDA(ctx, cudaMemcpyAsync(n_out, ctx->d_ncorners, (size_t)n_streams * 4, kind, ctx->stream));
  if (mem == FLV_MEM_HOST) return check_flags(ctx, n_streams);
  return FLV_OK;
}

int flv_download_eig(flv_ctx* ctx, int stream, float* out, flv_memspace mem) {
  if (!ctx || !out || stream < 0 || stream >= ctx->S) return FLV_ERR_INVALID;
  if (!ctx->keep_eig) FLV_FAIL(ctx, FLV_ERR_INVALID, "response map not kept: call flv_gftt_keep_response(ctx, 1) first");
  FLV_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_eig + (size_t)stream * ctx->w * ctx->h, (size_t)ctx->w * ctx->h * 4,
                                mem == FLV_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice,
                                ctx->stream));
  if (mem == FLV_MEM_HOST) FLV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FLV_OK;
}

static int feature_common(flv_ctx* ctx, int slot, int n_streams, const flv_feature_params* prm,
                          const double* existing_xy, const int* n_existing, float* new_xy, int* n_new,
                          flv_memspace mem, int redetect) {
  if (!ctx || !prm || !new_xy || !n_new || slot < 0 || slot >= FLV_NUM_SLOTS || n_streams < 1 ||
      n_streams > ctx->S || (redetect && (!existing_xy || !n_existing)))
    return FLV_ERR_INVALID;
  const int ncorn = redetect ? prm->gftt_num : 2 * prm->gftt_num;
  if (prm->max_region_feature_num < 1 || prm->boundary_dis < 0) return FLV_ERR_INVALID;
  const size_t np = (size_t)n_streams * ctx->max_pts;
  cudaMemcpyKind up = mem == FLV_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  cudaMemcpyKind down = mem == FLV_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  if (redetect) {
    if (mem == FLV_MEM_HOST)
      for (int s = 0; s < n_streams; ++s)
        if (n_existing[s] < 0 || n_existing[s] > ctx->max_pts) return FLV_ERR_INVALID;
    FLV_CUDA(ctx, cudaMemcpyAsync(ctx->d_exist, existing_xy, np * 16, up, ctx->stream));
    FLV_CUDA(ctx, cudaMemcpyAsync(ctx->d_nexist, n_existing, (size_t)n_streams * 4, up, ctx->stream));
  }
  int rc = FLV_OK;
  if (ctx->prep_valid && ctx->prep_slot == slot && ctx->prep_streams == n_streams && ctx->prep_ncorn == ncorn &&
      ctx->prep_ql == prm->gftt_ql && ctx->prep_dis == prm->gftt_dis) {
    FLV_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_gftt, 0));      // corners were computed ahead of time
  } else {
    if (ctx->prep_valid) FLV_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_gftt, 0));   // scratch buffers in use
    rc = flv_launch_gftt(ctx, slot, n_streams, ncorn, prm->gftt_ql, (double)prm->gftt_dis);
  }
  ctx->prep_valid = 0;
  if (rc) return rc;
  rc = flv_launch_region(ctx, slot, n_streams, prm, redetect);
  if (rc) return rc;
  FLV_CUDA(ctx, cudaMemcpyAsync(new_xy, ctx->d_newxy, np * 8, down, ctx->stream));
  FLV_CUDA(ctx, cudaMemcpyAsync(n_new, ctx->d_nnew, (size_t)n_streams * 4, down, ctx->stream));
  if (mem == FLV_MEM_HOST) return check_flags(ctx, n_streams);
  return FLV_OK;
}

int flv_feature_prepare(flv_ctx* ctx, int slot, int n_streams, const flv_feature_params* prm, int redetect) {
  if (!ctx || !prm || slot < 0 || slot >= FLV_NUM_SLOTS || n_streams < 1 || n_streams > ctx->S) return FLV_ERR_INVALID;
  if (!ctx->aux_stream) {
    FLV_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
    FLV_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    FLV_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_gftt, cudaEventDisableTiming));
  }
  const int ncorn = redetect ? prm->gftt_num : 2 * prm->gftt_num;
  // fork: the auxiliary stream sees everything enqueued so far (the slot's pyramid, the previous consumer of the
  // GFTT scratch buffers), then runs corner response + min-distance selection concurrently with the caller's stream
  FLV_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
  FLV_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
  cudaStream_t main_stream = ctx->stream;
  ctx->stream = ctx->aux_stream;
  const int rc = flv_launch_gftt(ctx, slot, n_streams, ncorn, prm->gftt_ql, (double)prm->gftt_dis);
  ctx->stream = main_stream;
  if (rc) return rc;
  FLV_CUDA(ctx, cudaEventRecord(ctx->ev_gftt, ctx->aux_stream));
  ctx->prep_valid = 1; ctx->prep_slot = slot; ctx->prep_streams = n_streams; ctx->prep_ncorn = ncorn;
  ctx->prep_ql = prm->gftt_ql; ctx->prep_dis = prm->gftt_dis;
  return FLV_OK;
}

int flv_feature_detect(flv_ctx* ctx, int slot, int n_streams, const flv_feature_params* prm,
                       float* new_xy, int* n_new, flv_memspace mem) {
  return feature_common(ctx, slot, n_streams, prm, nullptr, nullptr, new_xy, n_new, mem, 0);
}

int flv_feature_redetect(flv_ctx* ctx, int slot, int n_streams, const flv_feature_params* prm,
                         const double* existing_xy, const int* n_existing, float* new_xy,
                         int* n_new, flv_memspace mem) {
  return feature_common(ctx, slot, n_streams, prm, existing_xy, n_existing, new_xy, n_new, mem, 1);
}

}  // extern "C"
