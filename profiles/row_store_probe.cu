// Probe (round 2): what bounds the ROW WRITES of kz_step on B200, with no move generation at all?
// One warp per game writes that game's 13,536-byte mask row and 14,904-byte observation row (the real layouts:
// mask rows 16-byte aligned with stride 13,536, observation rows contiguous = 8-byte aligned) in several styles:
//   bit 0  mask row: zero fill with 256-bit stores          bit 1  obs row: zero fill (float2 head/tail + 256-bit body)
//   bit 2  mask row: ~25 sparse 16-byte overwrites          bit 3  obs row: ~40 single floats + 4 constant planes
//   bit 4  obs row written ONCE from a shared-memory image of its non-zero floats (no overwrite of zeroed sectors)
//   bit 5  flat fill of the same bytes (every warp writes 1 KB pieces round-robin): the "pure fill" reference
//   bit 6  mask row written ONCE: every 32-byte piece composed from a bit image (no overwrite)
//   bit 7  CTA-cooperative: the 8 warps build their games' images, then all 256 threads stream the tile's 8 mask rows
//          (108 KB contiguous) and 8 observation rows (119 KB contiguous) once, 32 bytes per thread per store
//   bit 8  with bit 7: zeros only (no images), the ceiling of that access pattern
// swept over resident CTAs per SM.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o row_store_probe row_store_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define FULL 0xffffffffu
#define MASK_STRIDE 13536
#define OBS_FLOATS 3726

__device__ __forceinline__ void st_zero256(void* p) {
  const uint32_t z = 0;
  asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(z) : "memory");
}
__device__ __forceinline__ void st256(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ uint32_t mix(uint32_t h) {
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}

__global__ void __launch_bounds__(256) probe(uint8_t* mask, float* obs, int n, int mode, int* counter) {
  __shared__ uint32_t s_img[8][128];  // per warp: bit image of the non-zero floats (obs) / legal bits (mask: 423 words > 128,
  __shared__ uint32_t s_bm[8][448];   // so the mask image has its own array)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (mode & 32) {
    uint8_t* base = mask;  // the two buffers are one allocation
    if (mode & 512) {  // 128-bit stores: 512 B per warp instruction
      const size_t total = (size_t)n * (MASK_STRIDE + OBS_FLOATS * 4) / 512;
      for (size_t i = (size_t)blockIdx.x * 8 + warp; i < total; i += (size_t)gridDim.x * 8)
        reinterpret_cast<uint4*>(base + i * 512)[lane] = make_uint4(0, 0, 0, 0);
      return;
    }
    const size_t total = (size_t)n * (MASK_STRIDE + OBS_FLOATS * 4) / 1024;
    for (size_t i = (size_t)blockIdx.x * 8 + warp; i < total; i += (size_t)gridDim.x * 8) st_zero256(base + i * 1024 + lane * 32);
    return;
  }
  if (mode & 128) {
    __shared__ float s_pv[8][20];
    for (int tile = blockIdx.x; tile < n / 8; tile += gridDim.x) {
      const int g = tile * 8 + warp;
      if (!(mode & 256)) {
        for (int i = lane; i < 448; i += 32) s_bm[warp][i] = 0;
        for (int i = lane; i < 128; i += 32) s_img[warp][i] = 0;
        __syncwarp();
        if (lane < 25) { const uint32_t b = mix(g * 131 + lane) % 13527u; atomicOr(&s_bm[warp][b >> 5], 1u << (b & 31)); }
        for (int j = 0; j < 2; j++) {
          const int i = lane + 32 * j;
          if (i < 40) { const uint32_t f = (mix(g * 977 + i) % 28u) * 81 + (mix(g * 31 + i) % 81u); atomicOr(&s_img[warp][f >> 5], 1u << (f & 31)); }
        }
        if (lane < 18) s_pv[warp][lane] = (lane & 3) == 0 && lane < 16 ? 0.5f : 0.f;
      }
      __syncthreads();
      uint8_t* mbase = mask + (size_t)tile * 8 * MASK_STRIDE;
#pragma unroll 1
      for (int p = threadIdx.x; p < 8 * 423; p += 256) {
        if (mode & 256) { st_zero256(mbase + 32 * p); continue; }
        const int gg = p / 423, q = p - gg * 423;
        const uint32_t w = s_bm[gg][q];
        if (w == 0) st_zero256(mbase + 32 * p);
        else {
          uint32_t v[8];
#pragma unroll
          for (int k = 0; k < 8; k++) v[k] = (((w >> (4 * k)) & 0xF) * 0x00204081u) & 0x01010101u;
          st256(mbase + 32 * p, v);
        }
      }
      char* obase = reinterpret_cast<char*>(obs + (size_t)tile * 8 * OBS_FLOATS);
#pragma unroll 1
      for (int p = threadIdx.x; p < OBS_FLOATS; p += 256) {  // 8 rows x 3726 floats = 3726 pieces of 8 floats
        if (mode & 256) { st_zero256(obase + 32 * p); continue; }
        const int f0 = 8 * p;
        const int gg = f0 / OBS_FLOATS, r0 = f0 - gg * OBS_FLOATS;
        if (r0 + 8 <= 28 * 81) {
          const uint32_t lo = s_img[gg][r0 >> 5], hi = s_img[gg][(r0 >> 5) + 1];
          const uint32_t bits = (uint32_t)((((unsigned long long)hi << 32) | lo) >> (r0 & 31)) & 0xFF;
          if (bits == 0) { st_zero256(obase + 32 * p); continue; }
          uint32_t v[8];
#pragma unroll
          for (int k = 0; k < 8; k++) v[k] = ((bits >> k) & 1) ? 0x3F800000u : 0u;
          st256(obase + 32 * p, v);
        } else {
          uint32_t v[8];
#pragma unroll
          for (int k = 0; k < 8; k++) {
            int r = r0 + k, g2 = gg;
            if (r >= OBS_FLOATS) { r -= OBS_FLOATS; g2++; }
            float x;
            if (r < 28 * 81) x = ((s_img[g2 & 7][r >> 5] >> (r & 31)) & 1) ? 1.0f : 0.f;
            else x = s_pv[g2 & 7][(r - 28 * 81) / 81];
            v[k] = __float_as_uint(x);
          }
          st256(obase + 32 * p, v);
        }
      }
      __syncthreads();
    }
    return;
  }
  for (int g = blockIdx.x * 8 + warp; g < n; g += gridDim.x * 8) {
    uint8_t* mrow = mask + (size_t)g * MASK_STRIDE;
    float* orow = obs + (size_t)g * OBS_FLOATS;
    if (mode & 64) {
      // compose the bitmap image (25 random bits set), then write every 32-byte piece once
      for (int i = lane; i < 448; i += 32) s_bm[warp][i] = 0;
      __syncwarp();
      if (lane < 25) { const uint32_t b = mix(g * 131 + lane) % 13527u; atomicOr(&s_bm[warp][b >> 5], 1u << (b & 31)); }
      __syncwarp();
#pragma unroll 1
      for (int q = lane; q < 423; q += 32) {  // piece q = bits [32q, 32q + 32) = word q
        const uint32_t w = s_bm[warp][q];
        if (w == 0) st_zero256(mrow + 32 * q);
        else {
          uint32_t v[8];
#pragma unroll
          for (int k = 0; k < 8; k++) v[k] = (((w >> (4 * k)) & 0xF) * 0x00204081u) & 0x01010101u;
          st256(mrow + 32 * q, v);
        }
      }
    }
    if (mode & 1) {
#pragma unroll 1
      for (int q = lane; q < 423; q += 32) st_zero256(mrow + 32 * q);
    }
    if (mode & 4) {
      __syncwarp();
      if (lane < 25) {
        const uint32_t q = mix(g * 131 + lane) % 846u;
        reinterpret_cast<uint4*>(mrow)[q] = make_uint4(1, 0x0100, 0, 0x01000000);
      }
    }
    if (mode & 2) {
      float2* o2 = reinterpret_cast<float2*>(orow);
      const int head = (int)(((32u - (unsigned)((uintptr_t)orow & 31)) & 31u) >> 3);
      const int nb = (OBS_FLOATS * 4 - head * 8) >> 5;
      char* body = reinterpret_cast<char*>(orow) + head * 8;
      if (lane < head) o2[lane] = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int q = lane; q < nb; q += 32) st_zero256(body + 32 * q);
      const int tail0 = head + 4 * nb;
      if (lane < OBS_FLOATS / 2 - tail0) o2[tail0 + lane] = make_float2(0.f, 0.f);
    }
    if (mode & 8) {
      __syncwarp();
      for (int j = 0; j < 2; j++) {
        const int i = lane + 32 * j;
        if (i < 40) orow[(mix(g * 977 + i) % 28u) * 81 + (mix(g * 31 + i) % 81u)] = 1.0f;
      }
      for (int p = 0; p < 4; p++) {
        float* pl = orow + (28 + 4 * p) * 81;
        pl[lane] = 0.5f; pl[lane + 32] = 0.5f;
        if (lane < 17) pl[lane + 64] = 0.5f;
      }
    }
    if (mode & 16) {
      // image: 40 piece bits; the 4 constant planes are ranges, handled arithmetically
      for (int i = lane; i < 128; i += 32) s_img[warp][i] = 0;
      __syncwarp();
      for (int j = 0; j < 2; j++) {
        const int i = lane + 32 * j;
        if (i < 40) { const uint32_t f = (mix(g * 977 + i) % 28u) * 81 + (mix(g * 31 + i) % 81u); atomicOr(&s_img[warp][f >> 5], 1u << (f & 31)); }
      }
      __syncwarp();
      // row = 1863 float2; head float2s to reach a 32-byte line, then 32-byte pieces of 8 floats, then the tail
      float2* o2 = reinterpret_cast<float2*>(orow);
      const int head = (int)(((32u - (unsigned)((uintptr_t)orow & 31)) & 31u) >> 3);
      const int nb = (OBS_FLOATS * 4 - head * 8) >> 5;
      auto val = [&](int f) -> float {  // final value of float f of the row
        if (f >= 28 * 81) { const int p = f / 81 - 28; return (p & 3) == 0 && p < 16 ? 0.5f : 0.f; }
        return ((s_img[warp][f >> 5] >> (f & 31)) & 1) ? 1.0f : 0.f;
      };
      if (lane < head) o2[lane] = make_float2(val(2 * lane), val(2 * lane + 1));
#pragma unroll 1
      for (int q = lane; q < nb; q += 32) {
        const int f0 = head * 2 + 8 * q;
        // 8 consecutive floats: a 8-bit field of the image (may straddle two words) or a constant-plane range
        uint32_t bits;
        if (f0 + 8 <= 28 * 81) {
          const uint32_t lo = s_img[warp][f0 >> 5], hi = s_img[warp][(f0 >> 5) + 1];
          bits = (uint32_t)((((unsigned long long)hi << 32) | lo) >> (f0 & 31)) & 0xFF;
          if (bits == 0) { st_zero256(reinterpret_cast<char*>(orow) + head * 8 + 32 * q); continue; }
          uint32_t v[8];
#pragma unroll
          for (int k = 0; k < 8; k++) v[k] = ((bits >> k) & 1) ? 0x3F800000u : 0u;
          st256(reinterpret_cast<char*>(orow) + head * 8 + 32 * q, v);
        } else {
          uint32_t v[8];
#pragma unroll
          for (int k = 0; k < 8; k++) v[k] = __float_as_uint(val(f0 + k));
          st256(reinterpret_cast<char*>(orow) + head * 8 + 32 * q, v);
        }
      }
      const int tail0 = head + 4 * nb;
      if (lane < OBS_FLOATS / 2 - tail0) o2[tail0 + lane] = make_float2(val(2 * (tail0 + lane)), val(2 * (tail0 + lane) + 1));
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(128) fill_np(uint4* p, size_t n16) {  // non-persistent: 4 x 16 B per thread, like an elementwise fill
  const size_t i0 = ((size_t)blockIdx.x * 128 + threadIdx.x);
  const size_t stride = (size_t)gridDim.x * 128;
#pragma unroll
  for (int k = 0; k < 4; k++) { const size_t i = i0 + k * stride; if (i < n16) p[i] = make_uint4(0, 0, 0, 0); }
}


// ---- round-2 follow-up: is it the persistent grid or the access pattern that costs the 15 % between 0.30 and 0.255 ms?
// np_flat256: one CTA per 8 KB, 256-bit stores (the persistent flat fill's pattern, non-persistent)
__global__ void __launch_bounds__(256) np_flat256(uint8_t* base, size_t total_kb) {
  const size_t i = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i < total_kb) st_zero256(base + i * 1024 + (threadIdx.x & 31) * 32);
}
// np_rows: one CTA per tile of 8 games, one warp per game: its mask row, then its observation row (kz_step's layout), zeros
__global__ void __launch_bounds__(256) np_rows(uint8_t* mask, float* obs, int n) {
  const int lane = threadIdx.x & 31, g = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (g >= n) return;
  uint8_t* mrow = mask + (size_t)g * MASK_STRIDE;
  for (int q = lane; q < MASK_STRIDE / 32; q += 32) st_zero256(mrow + 32 * q);
  char* orow = reinterpret_cast<char*>(obs) + (size_t)g * OBS_FLOATS * 4;
  const int head = (int)(((32u - (unsigned)((uintptr_t)orow & 31)) & 31u) >> 3);
  const int nb = (OBS_FLOATS * 4 - head * 8) >> 5;
  if (lane < head) reinterpret_cast<float2*>(orow)[lane] = make_float2(0.f, 0.f);
  for (int q = lane; q < nb; q += 32) st_zero256(orow + head * 8 + 32 * q);
  const int tail0 = head + 4 * nb;
  if (lane < OBS_FLOATS / 2 - tail0) reinterpret_cast<float2*>(orow)[tail0 + lane] = make_float2(0.f, 0.f);
}
// p_streams: persistent grid, every thread keeps 4 far-apart streams of 16-byte stores (fill_np's pattern, persistent)
__global__ void __launch_bounds__(128) p_streams(uint4* p, size_t n16) {
  const size_t q = n16 / 4;
  for (size_t i = (size_t)blockIdx.x * 128 + threadIdx.x; i < q; i += (size_t)gridDim.x * 128) {
#pragma unroll
    for (int k = 0; k < 4; k++) p[i + k * q] = make_uint4(0, 0, 0, 0);
  }
}
// np_contig: non-persistent, 64 contiguous bytes per thread as 4 x 16 B, CTA of 128 threads covers 8 KB
__global__ void __launch_bounds__(128) np_contig(uint4* p, size_t n16) {
  const size_t i0 = (size_t)blockIdx.x * 512 + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 4; k++) { const size_t i = i0 + k * 128; if (i < n16) p[i] = make_uint4(0, 0, 0, 0); }
}

// ---- second follow-up: which property of the fast non-persistent fills matters?
// p_chunk: persistent, CTA b owns one contiguous chunk and walks it 8 KB per iteration (page-local, unlike the grid stride)
__global__ void __launch_bounds__(256) p_chunk(uint8_t* base, size_t total_kb) {
  const size_t per = (total_kb / 8 + gridDim.x - 1) / gridDim.x;  // 8 KB pieces per CTA
  const size_t lo = (size_t)blockIdx.x * per, hi = lo + per < total_kb / 8 ? lo + per : total_kb / 8;
  for (size_t i = lo; i < hi; i++) st_zero256(base + (i * 8 + (threadIdx.x >> 5)) * 1024 + (threadIdx.x & 31) * 32);
}
// np_run: non-persistent, CTA writes R consecutive 8 KB pieces
__global__ void __launch_bounds__(256) np_run(uint8_t* base, size_t total_kb, int R) {
  for (int r = 0; r < R; r++) {
    const size_t i = ((size_t)blockIdx.x * R + r) * 8 + (threadIdx.x >> 5);
    if (i < total_kb) st_zero256(base + i * 1024 + (threadIdx.x & 31) * 32);
  }
}
// np_tile_flat: non-persistent, CTA = tile of 8 games, but the CTA's 8 warps write the tile's 8 mask rows as ONE flat block
// (consecutive warps -> consecutive KB), then the 8 observation rows the same way
__global__ void __launch_bounds__(256) np_tile_flat(uint8_t* mask, float* obs, int n) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* mb = mask + (size_t)blockIdx.x * 8 * MASK_STRIDE;        // 108,288 B, 32-byte aligned
  for (int q = w; q < 8 * MASK_STRIDE / 1024; q += 8) st_zero256(mb + q * 1024 + lane * 32);
  if (w == 0 && lane < (8 * MASK_STRIDE % 1024) / 32) st_zero256(mb + (8 * MASK_STRIDE / 1024) * 1024 + lane * 32);
  uint8_t* ob = reinterpret_cast<uint8_t*>(obs) + (size_t)blockIdx.x * 8 * OBS_FLOATS * 4;  // 119,232 B, 32-byte aligned
  for (int q = w; q < 8 * OBS_FLOATS * 4 / 1024; q += 8) st_zero256(ob + q * 1024 + lane * 32);
  if (w == 0 && lane < (8 * OBS_FLOATS * 4 % 1024) / 32) st_zero256(ob + (8 * OBS_FLOATS * 4 / 1024) * 1024 + lane * 32);
}
// p_rows_dyn: persistent per-warp rows (zeros) with dynamically claimed tiles (no static-stride tail)
__global__ void __launch_bounds__(256) p_rows_dyn(uint8_t* mask, float* obs, int n, int* counter) {
  __shared__ int s_t;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int tile = blockIdx.x;
  while (tile < n / 8) {
    const int g = tile * 8 + w;
    uint8_t* mrow = mask + (size_t)g * MASK_STRIDE;
    for (int q = lane; q < MASK_STRIDE / 32; q += 32) st_zero256(mrow + 32 * q);
    char* orow = reinterpret_cast<char*>(obs) + (size_t)g * OBS_FLOATS * 4;
    const int head = (int)(((32u - (unsigned)((uintptr_t)orow & 31)) & 31u) >> 3);
    const int nb = (OBS_FLOATS * 4 - head * 8) >> 5;
    if (lane < head) reinterpret_cast<float2*>(orow)[lane] = make_float2(0.f, 0.f);
    for (int q = lane; q < nb; q += 32) st_zero256(orow + head * 8 + 32 * q);
    const int tail0 = head + 4 * nb;
    if (lane < OBS_FLOATS / 2 - tail0) reinterpret_cast<float2*>(orow)[tail0 + lane] = make_float2(0.f, 0.f);
    __syncthreads();
    if (threadIdx.x == 0) s_t = gridDim.x + atomicAdd(counter, 1);
    __syncthreads();
    tile = s_t;
  }
}

int main() {
  const int n = 65536;
  uint8_t* buf;
  const size_t mask_bytes = (size_t)n * MASK_STRIDE, obs_bytes = (size_t)n * OBS_FLOATS * 4;
  cudaMalloc(&buf, mask_bytes + obs_bytes + 1024);
  uint8_t* mask = buf;
  float* obs = reinterpret_cast<float*>(buf + mask_bytes);
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int modes[] = {32, 32 | 512, 1, 2, 3, 3 | 4, 3 | 8, 3 | 4 | 8, 1 | 16, 1 | 4 | 16, 64 | 16, 64, 16, 128, 128 | 256};
  const char* names[] = {"flat fill (reference)", "flat fill, 128-bit stores", "mask fill", "obs fill", "mask+obs fill", "fill + mask sparse", "fill + obs sparse",
                         "fill + both sparse (= kz_step's writers)", "mask fill + obs ONCE", "mask fill+sparse + obs ONCE",
                         "mask ONCE + obs ONCE", "mask ONCE only", "obs ONCE only", "CTA-cooperative tile, composed ONCE",
                         "CTA-cooperative tile, zeros"};
  printf("%-44s", "style \\ CTAs(8 warps)/SM");
  const int cps[] = {1, 2, 3, 4, 6, 8};
  for (int c : cps) printf("%9d", c);
  printf("   [ms per %d games; GB/s of the best]\n", n);
  for (size_t mi = 0; mi < sizeof(modes) / sizeof(int); mi++) {
    const int mode = modes[mi];
    double bytes = 0;
    if (mode & (1 | 64)) bytes += mask_bytes;
    if (mode & (2 | 16)) bytes += obs_bytes;
    if (mode & (32 | 128)) bytes = mask_bytes + obs_bytes;
    printf("%-44s", names[mi]);
    float best = 1e9;
    for (int c : cps) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int i = 0; i < 3; i++) probe<<<sms * c, 256>>>(mask, obs, n, mode, nullptr);
      cudaEventRecord(e0);
      for (int i = 0; i < 10; i++) probe<<<sms * c, 256>>>(mask, obs, n, mode, nullptr);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      ms /= 10;
      if (ms < best) best = ms;
      printf("%9.4f", ms);
    }
    printf("   %.0f GB/s  (%s)\n", bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  {  // non-persistent elementwise-style fill of the same bytes
    const size_t n16 = (mask_bytes + obs_bytes) / 16;
    const unsigned grid = (unsigned)((n16 + 511) / 512);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) fill_np<<<grid, 128>>>(reinterpret_cast<uint4*>(buf), n16);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; i++) fill_np<<<grid, 128>>>(reinterpret_cast<uint4*>(buf), n16);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 10;
    printf("%-44s%9.4f   %.0f GB/s  (non-persistent grid of %u CTAs x 128 threads x 64 B)\n", "flat fill, elementwise-style", ms,
           (double)(mask_bytes + obs_bytes) / ms / 1e6, grid);
  }
  {  // round-2 follow-up: non-persistent / multi-stream variants of the same bytes
    const size_t total = mask_bytes + obs_bytes, n16 = total / 16;
    auto timeit = [&](const char* name, auto launch) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int i = 0; i < 3; i++) launch();
      cudaEventRecord(e0);
      for (int i = 0; i < 10; i++) launch();
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      ms /= 10;
      printf("%-60s%9.4f   %.0f GB/s  (%s)\n", name, ms, (double)total / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    timeit("non-persistent, CTA = 8 KB flat, 256-bit stores", [&] { np_flat256<<<(unsigned)((total / 1024 + 7) / 8), 256>>>(buf, total / 1024); });
    timeit("non-persistent, CTA = tile of 8 games, warp = one game's rows", [&] { np_rows<<<n / 8, 256>>>(mask, obs, n); });
    timeit("non-persistent, 64 contiguous bytes per thread", [&] { np_contig<<<(unsigned)((n16 + 511) / 512), 128>>>(reinterpret_cast<uint4*>(buf), n16); });
    for (int c : {4, 8, 16})
      timeit(c == 4 ? "persistent x4 CTAs(128)/SM, 4 far streams per thread" : c == 8 ? "persistent x8" : "persistent x16",
             [&] { p_streams<<<sms * c, 128>>>(reinterpret_cast<uint4*>(buf), n16); });
  }
  {  // second follow-up
    const size_t total = mask_bytes + obs_bytes;
    int* counter;
    cudaMalloc(&counter, 4);
    auto timeit = [&](const char* name, auto launch) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int i = 0; i < 3; i++) launch();
      cudaEventRecord(e0);
      for (int i = 0; i < 10; i++) launch();
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      ms /= 10;
      printf("%-72s%9.4f   %.0f GB/s  (%s)\n", name, ms, (double)total / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    char nm[128];
    for (int c : {1, 3, 8}) { snprintf(nm, sizeof nm, "persistent flat, CTA-contiguous chunks, %d CTAs/SM", c); timeit(nm, [&] { p_chunk<<<sms * c, 256>>>(buf, total / 1024); }); }
    for (int R : {4, 16, 64, 256}) { snprintf(nm, sizeof nm, "non-persistent flat, %d x 8 KB per CTA", R); timeit(nm, [&] { np_run<<<(unsigned)((total / 8192 + R - 1) / R), 256>>>(buf, total / 1024, R); }); }
    timeit("non-persistent, CTA = tile, rows written as flat blocks by the CTA", [&] { np_tile_flat<<<n / 8, 256>>>(mask, obs, n); });
    for (int kb : {0, 24, 48, 72, 100, 200}) {  // dynamic shared memory caps the resident CTAs per SM: 8, 8, 4, 3, 2, 1
      cudaFuncSetAttribute(np_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      snprintf(nm, sizeof nm, "non-persistent per-warp rows, %d KB dynamic smem per CTA (caps residency)", kb);
      timeit(nm, [&] { np_rows<<<n / 8, 256, kb * 1024>>>(mask, obs, n); });
    }
    for (int c : {1, 2, 3, 4}) {
      snprintf(nm, sizeof nm, "persistent per-warp rows, dynamic tile claims, %d CTAs/SM", c);
      timeit(nm, [&] { cudaMemsetAsync(counter, 0, 4); p_rows_dyn<<<sms * c, 256>>>(mask, obs, n, counter); });
    }
  }
  return 0;
}
