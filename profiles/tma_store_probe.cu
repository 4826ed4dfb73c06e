// Probe: can a few warps per SM stream row-sized blocks (28 KB) from shared memory to HBM with bulk-copy stores
// (cp.async.bulk.global.shared::cta) at the fill rate?  W writer warps per CTA, 3 CTAs per SM.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#ifndef ROW
#define ROW 28432   // 13536 + 14896
#endif
__global__ void __launch_bounds__(256, 3) probe(uint8_t* out, int rows, int writers, int compose) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= writers) return;
  unsigned char* img = sm + (size_t)warp * ROW;
  for (int i = lane; i < ROW / 16; i += 32) reinterpret_cast<uint4*>(img)[i] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  const int wid = blockIdx.x * writers + warp, nw = gridDim.x * writers;
  for (int r = wid; r < rows; r += nw) {
    if (compose) {  // ~100 sparse 16-byte updates, as a row composer would do
      for (int k = lane; k < 128; k += 32) reinterpret_cast<uint4*>(img)[(k * 37 + r) % (ROW / 16)] = make_uint4(r, k, 1, 0);
    }
    __syncwarp();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (lane == 0) {
      const uint32_t src = (uint32_t)__cvta_generic_to_shared(img);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(out + (size_t)r * ROW), "r"(src), "n"(ROW) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
int main() {
  const int rows = 65536;
  uint8_t* out;
  cudaMalloc(&out, (size_t)rows * ROW);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int writers = 1; writers <= 2; writers++)
    for (int compose = 0; compose <= 1; compose++) {
      const size_t dyn = (size_t)writers * ROW;
      cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int i = 0; i < 3; i++) probe<<<sms * 3, 256, dyn>>>(out, rows, writers, compose);
      cudaEventRecord(e0);
      for (int i = 0; i < 10; i++) probe<<<sms * 3, 256, dyn>>>(out, rows, writers, compose);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
      printf("writers/CTA %d compose %d: %.4f ms per %d rows of %d B -> %.0f GB/s (%s)\n", writers, compose, ms, rows, ROW,
             (double)rows * ROW / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
