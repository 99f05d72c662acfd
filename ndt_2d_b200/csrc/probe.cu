// probe.cu -- roofline denominators measured on the box the bench runs on.
//
// ndt2d_probe_gather: random 32-byte record reads from a table of a given size
// ("achievable gather roofline"): every thread issues a
// chain-free stream of 32-B loads at pseudo-random record indices and folds
// them into a checksum.  A table of the search model's size (tens of KB) stays
// in L1; larger ones measure L2, then HBM.  ndt2d_probe_copy: plain device copy.
#include <cuda_runtime.h>
#include <stdint.h>

#include "ndt2d_internal.h"

namespace
{

__global__ void __launch_bounds__(256) gather_probe_kernel(
  const uint4 * __restrict__ table, uint32_t n_records, uint32_t reads_per_thread,
  uint32_t * __restrict__ sink)
{
  uint32_t state = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  uint32_t acc = 0;
#pragma unroll 8
  for (uint32_t k = 0; k < reads_per_thread; ++k) {
    state = state * 1664525u + 1013904223u;
    const uint32_t r = __umulhi(state, n_records);   // uniform in [0, n_records)
    const uint4 a = table[2 * r], b = table[2 * r + 1];
    acc ^= a.x ^ a.w ^ b.y ^ b.z;
  }
  if (acc == 0x9e3779b9u) {sink[0] = acc;}  // keeps the loads alive
}

__global__ void __launch_bounds__(256) copy_probe_kernel(
  const uint4 * __restrict__ src, uint4 * __restrict__ dst, size_t n)
{
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
    i += static_cast<size_t>(gridDim.x) * blockDim.x)
  {
    dst[i] = src[i];
  }
}

}  // namespace

extern "C" {

NDT2D_API int ndt2d_probe_gather(int device, size_t table_bytes, double * out_gbps)
{
  if (!out_gbps || table_bytes < 64) {return NDT2D_ERR_INVALID;}
  if (ndt2d_device_count() <= 0) {return NDT2D_ERR_NO_DEVICE;}
  if (device >= 0) {NDT2D_CUDA_TRY(cudaSetDevice(device));}
  const uint32_t n_records = static_cast<uint32_t>(table_bytes / 32);
  void * table = nullptr;
  uint32_t * sink = nullptr;
  NDT2D_CUDA_TRY(cudaMalloc(&table, static_cast<size_t>(n_records) * 32));
  NDT2D_CUDA_TRY(cudaMalloc(&sink, 64));
  NDT2D_CUDA_TRY(cudaMemset(table, 1, static_cast<size_t>(n_records) * 32));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device >= 0 ? device : 0);
  const uint32_t grid = sms * 8, reads = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    gather_probe_kernel<<<grid, 256>>>(static_cast<const uint4 *>(table), n_records, reads, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double gbps = static_cast<double>(grid) * 256 * reads * 32 / (ms * 1e-3) / 1e9;
    if (rep > 0 && gbps > best) {best = gbps;}
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(table);
  cudaFree(sink);
  NDT2D_CUDA_TRY(cudaGetLastError());
  *out_gbps = best;
  return NDT2D_OK;
}

NDT2D_API int ndt2d_probe_copy(int device, size_t bytes, double * out_gbps)
{
  if (!out_gbps || bytes < 1024) {return NDT2D_ERR_INVALID;}
  if (ndt2d_device_count() <= 0) {return NDT2D_ERR_NO_DEVICE;}
  if (device >= 0) {NDT2D_CUDA_TRY(cudaSetDevice(device));}
  const size_t n = bytes / 16;
  void * a = nullptr, * b = nullptr;
  NDT2D_CUDA_TRY(cudaMalloc(&a, n * 16));
  NDT2D_CUDA_TRY(cudaMalloc(&b, n * 16));
  NDT2D_CUDA_TRY(cudaMemset(a, 1, n * 16));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device >= 0 ? device : 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    copy_probe_kernel<<<sms * 16, 256>>>(static_cast<const uint4 *>(a), static_cast<uint4 *>(b), n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double gbps = 2.0 * n * 16 / (ms * 1e-3) / 1e9;
    if (rep > 0 && gbps > best) {best = gbps;}
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(a);
  cudaFree(b);
  NDT2D_CUDA_TRY(cudaGetLastError());
  *out_gbps = best;
  return NDT2D_OK;
}

}  // extern "C"
