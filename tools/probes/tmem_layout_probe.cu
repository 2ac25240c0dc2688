// Probe: register <-> (lane, column) mapping of tcgen05.ld.16x256b.x2 on sm_100a (used to size the gate-warp I/O in gru_scan.cu).
// nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_layout_probe tmem_layout_probe.cu && ./tmem_layout_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void probe(float* out) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_ptr)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_ptr + ((uint32_t)(warp * 32) << 16);
  const int row = warp * 32 + lane;
  uint32_t v[16];
  for (int c = 0; c < 16; ++c) v[c] = __float_as_uint((float)(row * 100 + c));
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(base),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
      "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  for (int half = 0; half < 2; ++half) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(base + ((uint32_t)(half * 16) << 16))
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out[((warp * 2 + half) * 32 + lane) * 8 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_ptr), "r"(32u) : "memory");
}

int main() {
  float* d;
  cudaMalloc(&d, 4 * 2 * 32 * 8 * sizeof(float));
  probe<<<1, 128>>>(d);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
  static float h[4 * 2 * 32 * 8];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int warp = 0; warp < 2; ++warp)
    for (int half = 0; half < 2; ++half)
      for (int lane = 0; lane < 32; ++lane) {
        printf("w%d h%d lane %2d:", warp, half, lane);
        for (int i = 0; i < 8; ++i) printf(" r%d=(%3d,%2d)", i, (int)h[((warp * 2 + half) * 32 + lane) * 8 + i] / 100, (int)h[((warp * 2 + half) * 32 + lane) * 8 + i] % 100);
        printf("\n");
      }
  return 0;
}
