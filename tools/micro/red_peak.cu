// Microbenchmark: peak rate of scattered vector reductions into an L2-resident histogram.
// This is the "atomic roofline" of the chaos-game kernel: nothing but index generation and
// red.global.add into random buckets. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t xs(uint32_t &s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

template <int MODE>
__global__ void __launch_bounds__(256) k(float4 *h4, float *h1, uint32_t nb, int iters) {
   uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
   for (int i = 0; i < iters; i++) {
      uint32_t r = xs(s);
      uint32_t b = (uint32_t)(((uint64_t)r * nb) >> 32);
      if (MODE == 0) { // one v4 red
         asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(h4 + b), "f"(1.f), "f"(2.f), "f"(3.f), "f"(255.f) : "memory");
      } else if (MODE == 1) { // v4 + scalar count in a second plane
         asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(h4 + b), "f"(1.f), "f"(2.f), "f"(3.f), "f"(255.f) : "memory");
         asm volatile("red.global.add.f32 [%0], %1;" ::"l"(h1 + b), "f"(255.f) : "memory");
      } else if (MODE == 2) { // scalar only
         asm volatile("red.global.add.f32 [%0], %1;" ::"l"(h1 + b), "f"(255.f) : "memory");
      } else if (MODE == 3) { // 4 scalar reds (what an integer histogram would need)
         float *p = (float *)(h4 + b);
         asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(1.f) : "memory");
         asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + 1), "f"(1.f) : "memory");
         asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + 2), "f"(1.f) : "memory");
         asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + 3), "f"(1.f) : "memory");
      } else if (MODE == 4) { // v2
         asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"((float2 *)(h4 + b)), "f"(1.f), "f"(2.f) : "memory");
      } else if (MODE == 5) { // u32 scalar x1
         asm volatile("red.global.add.u32 [%0], %1;" ::"l"((unsigned *)(h1 + b)), "r"(255u) : "memory");
      }
   }
}

template <int MODE> double run(float4 *h4, float *h1, uint32_t nb, int blocks, int iters) {
   cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
   k<MODE><<<blocks, 256>>>(h4, h1, nb, iters / 8);
   cudaDeviceSynchronize();
   cudaEventRecord(a);
   k<MODE><<<blocks, 256>>>(h4, h1, nb, iters);
   cudaEventRecord(b); cudaEventSynchronize(b);
   float ms; cudaEventElapsedTime(&ms, a, b);
   return (double)blocks * 256 * iters / (ms * 1e-3);
}

int main() {
   const char *names[] = {"v4.f32", "v4.f32+f32", "f32", "4x f32", "v2.f32", "u32"};
   size_t sizes[] = {1u << 19, 2127924, 1u << 23, 33635044};
   for (size_t nb : sizes) {
      float4 *h4; float *h1;
      cudaMalloc(&h4, nb * 16); cudaMalloc(&h1, nb * 4);
      cudaMemset(h4, 0, nb * 16); cudaMemset(h1, 0, nb * 4);
      for (int occ : {4, 8}) {
         int blocks = 148 * occ, iters = 4000;
         double r[6];
         r[0] = run<0>(h4, h1, (uint32_t)nb, blocks, iters);
         r[1] = run<1>(h4, h1, (uint32_t)nb, blocks, iters);
         r[2] = run<2>(h4, h1, (uint32_t)nb, blocks, iters);
         r[3] = run<3>(h4, h1, (uint32_t)nb, blocks, iters);
         r[4] = run<4>(h4, h1, (uint32_t)nb, blocks, iters);
         r[5] = run<5>(h4, h1, (uint32_t)nb, blocks, iters);
         printf("buckets=%zu (%.1f MB v4 plane) blocks/SM=%d:", nb, nb * 16 / 1e6, occ);
         for (int i = 0; i < 6; i++) printf("  %s %.3e hits/s", names[i], r[i]);
         printf("\n");
      }
      cudaFree(h4); cudaFree(h1);
   }
   return 0;
}
