// Microbenchmark: latency of the hand-offs the pipelined tcgen05 kernels are built from.
//   (a) mbarrier.arrive by one warp -> try_wait / test_wait wake-up in another warp          (pure barrier hop)
//   (b) tcgen05.mma (M=128, N, K=16 x ksteps, operands in shared memory) + tcgen05.commit -> waiter wake-up
// One CTA; warp 0 issues, warp 1 waits; both read clock64 of the same SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../include -I../guided_attention_b200/csrc hop_latency.cu -o hop_latency
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"

namespace ga { int fail(int code, const char*, ...) { return code; } }
using namespace ga::tc;

__device__ __forceinline__ bool test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

// mode 0: plain arrive; mode 1: MMA + commit.  spin 0: try_wait (may suspend), 1: test_wait (busy poll)
__global__ void __launch_bounds__(64) hop_kernel(int mode, int spin, int N, int ksteps, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  __shared__ long long t_issue;
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 64) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = uniform_u32(tmem_slot);
  const uint32_t go = smem_u32(&bars[0]), back = smem_u32(&bars[1]);
  const uint32_t idesc = make_idesc(0, 0, N, 128);
  const uint64_t da = smem_desc_sw128(base, 16, 1024), db = smem_desc_sw128(base + 16384, 16, 1024);
  long long total = 0, tmax = 0;
  for (int it = 0; it < iters; ++it) {
    const uint32_t par = (uint32_t)it & 1u;
    if (warp == 0) {
      if (it > 0) mbar_wait(back, par ^ 1u);
      __syncwarp();
      if (elect_one()) {
        t_issue = clock64();
        if (mode == 0) {
          mbar_arrive(go);
        } else {
          for (int ks = 0; ks < ksteps; ++ks)
            mma_ss(tmem, desc_advance(da, (ks & 3) * 32), desc_advance(db, (ks & 3) * 32), idesc, ks > 0);
          tc_commit(go);
        }
      }
      __syncwarp();
    } else {
      if (spin) { while (!test_wait(go, par)) {} } else { mbar_wait(go, par); }
      const long long t1 = clock64();
      if (lane == 0) {
        const long long d = t1 - *(volatile long long*)&t_issue;
        if (it >= 10) { total += d; if (d > tmax) tmax = d; }
        mbar_arrive(back);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (warp == 1 && lane == 0) { out[0] = total / (iters - 10); out[1] = tmax; }
  if (warp == 0) tmem_dealloc(tmem, 512u);
}

int main() {
  long long* out;
  cudaMalloc(&out, 16);
  cudaFuncSetAttribute(hop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  struct { int mode, spin, N, ks; const char* what; } cases[] = {
      {0, 0, 0, 0, "mbarrier.arrive -> try_wait"},      {0, 1, 0, 0, "mbarrier.arrive -> test_wait spin"},
      {1, 0, 48, 1, "1 MMA N=48 + commit -> try_wait"}, {1, 1, 48, 1, "1 MMA N=48 + commit -> test_wait spin"},
      {1, 1, 80, 3, "3 MMA N=80 + commit -> spin"},     {1, 1, 48, 5, "5 MMA N=48 + commit -> spin"},
      {1, 1, 128, 3, "3 MMA N=128 + commit -> spin"},   {1, 1, 256, 4, "4 MMA N=256 + commit -> spin"},
      {1, 0, 80, 3, "3 MMA N=80 + commit -> try_wait"}, {1, 0, 256, 4, "4 MMA N=256 + commit -> try_wait"}};
  for (auto& c : cases) {
    hop_kernel<<<1, 64, 64 * 1024>>>(c.mode, c.spin, c.N ? c.N : 16, c.ks, 2010, out);
    long long h[2] = {0, 0};
    cudaError_t e = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%-44s avg %6lld cycles  max %6lld  (%s)\n", c.what, h[0], h[1], cudaGetErrorString(e));
  }
  return 0;
}
