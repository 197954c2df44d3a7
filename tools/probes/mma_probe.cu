// Micro-probe: cycles per tcgen05.mma (M=128, N, K=16, bf16) issued back-to-back by one elected thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I. -o /tmp/mma_probe tools/probes/mma_probe.cu
#include <cstdio>
#include "../../video_diffusion_nnx_b200/csrc/vdn_common.cuh"
namespace vdn { void set_last_error(const char*, ...) {} int check_launch(const char*) { return 0; } }
using namespace vdn;

template <int N, int SW, int NACC, int ROWSTEP, int SPIN>
__global__ void probe(long long* out, int reps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t spin_bar;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* smem = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&spin_bar, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base_smem, 512); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
      const uint32_t sa0 = smem_u32(smem), sb0 = sa0 + 48 * 1024;
      constexpr uint32_t kLayout = umma_layout_type(SW);
      for (int round = 0; round < 3; ++round) {
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
          const uint32_t sa = sa0 + (uint32_t)((r % 4) * ROWSTEP);
          const uint64_t da = umma_smem_desc(sa, 16, 8 * SW, kLayout);
          const uint64_t db = umma_smem_desc(sb0, 16, 8 * SW, kLayout);
          umma_bf16(tmem_base + (uint32_t)((r % NACC) * N), da + (uint64_t)((r & 1) * 2), db, idesc, r >= NACC);
        }
        long long t1 = clock64();
        tc_commit(&bar);
        mbar_wait(&bar, round & 1);
        long long t2 = clock64();
        out[round * 2] = t1 - t0;
        out[round * 2 + 1] = t2 - t0;
      }
      if (SPIN) mbar_arrive(&spin_bar);
    }
    __syncwarp();
  } else if (SPIN) {
    mbar_wait(&spin_bar, 0);  // other warps poll an mbarrier like epilogue warps waiting for the accumulator
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int N, int SW, int NACC, int ROWSTEP, int SPIN = 0>
void run(const char* name, int reps) {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(probe<N, SW, NACC, ROWSTEP, SPIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  probe<N, SW, NACC, ROWSTEP, SPIN><<<1, 192, 80 * 1024>>>(d, reps);
  long long h[6];
  cudaError_t e = cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
  printf("%-34s reps %3d: issue %6lld cyc, done %6lld cyc  -> %.1f cyc/MMA (%s)\n", name, reps, h[4], h[5],
         (double)h[5] / reps, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<32, 64, 1, 0>("N=32 SW64 1acc", 18);
  run<32, 64, 1, 0>("N=32 SW64 1acc", 72);
  run<32, 64, 2, 0>("N=32 SW64 2acc", 72);
  run<32, 64, 4, 0>("N=32 SW64 4acc", 72);
  run<32, 128, 1, 0>("N=32 SW128 1acc", 72);
  run<64, 128, 1, 0>("N=64 SW128 1acc", 72);
  run<128, 128, 1, 0>("N=128 SW128 1acc", 72);
  run<256, 128, 1, 0>("N=256 SW128 1acc", 72);
  run<256, 128, 2, 0>("N=256 SW128 2acc", 72);
  run<32, 64, 1, 4096>("N=32 SW64 1acc rowstep4096", 72);
  run<32, 64, 1, 64>("N=32 SW64 1acc rowstep64(misaligned)", 72);
  run<32, 64, 1, 4096, 1>("N=32 SW64 1acc + 5 spinning warps", 72);
  return 0;
}
