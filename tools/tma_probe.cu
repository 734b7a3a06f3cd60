// TMA tile-load probe (debug aid): loads one box of a u8 image with cp.async.bulk.tensor.{2d,3d} and checks it.
// usage: tma_probe <rank 2|3> <box_w> <box_h> <x> <y>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__device__ __forceinline__ unsigned sa(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void probe(const CUtensorMap* tm, int x, int y, int z, int bytes, unsigned char* out) {
  extern __shared__ __align__(128) unsigned char tile[];
  __shared__ unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sa(&bar)), "r"(bytes) : "memory");
    if (RANK == 2)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(sa(tile)), "l"(tm), "r"(x), "r"(y), "r"(sa(&bar)) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(sa(tile)), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(sa(&bar)) : "memory");
  }
  asm volatile(
      "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(sa(&bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = tile[i];
}

int main(int argc, char** argv) {
  const int rank = atoi(argv[1]), bw = atoi(argv[2]), bh = atoi(argv[3]), x = atoi(argv[4]), y = atoi(argv[5]);
  const int w = 752, h = 480, pitch = 768, S = 2;
  std::vector<unsigned char> img((size_t)S * pitch * h);
  for (size_t i = 0; i < img.size(); ++i) img[i] = (unsigned char)((i * 2654435761u) >> 24);
  unsigned char *d_img, *d_out; CUtensorMap* d_tm;
  cudaMalloc(&d_img, img.size()); cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice);
  cudaMalloc(&d_out, bw * bh); cudaMalloc(&d_tm, sizeof(CUtensorMap));
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
  CUtensorMap tm; memset(&tm, 0, sizeof(tm));
  const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)S};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * h};
  const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
  CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d_img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("rank %d box %dx%d at (%d,%d): encode=%d ", rank, bw, bh, x, y, (int)r);
  cudaMemcpy(d_tm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
  if (rank == 2) probe<2><<<1, 128, bw * bh>>>(d_tm, x, y, 1, bw * bh, d_out);
  else probe<3><<<1, 128, bw * bh>>>(d_tm, x, y, 1, bw * bh, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("run=%s ", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<unsigned char> out(bw * bh); cudaMemcpy(out.data(), d_out, bw * bh, cudaMemcpyDeviceToHost);
    int bad = 0; const int z = rank == 3 ? 1 : 0;
    for (int j = 0; j < bh; ++j) for (int i = 0; i < bw; ++i) bad += out[j * bw + i] != img[(size_t)z * pitch * h + (size_t)(y + j) * pitch + x + i];
    printf("mismatches=%d", bad);
  }
  printf("\n");
  return 0;
}
