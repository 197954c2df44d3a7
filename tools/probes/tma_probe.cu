// Micro-probe: bytes per clock one SM can pull from L2 into shared memory through TMA (2-D tiled boxes of `rows` x 128 B,
// 128B swizzle, a ring of S stages, nobody consumes the data), as a function of how many SMs pull at the same time.
// Answers: is the ~51 B/clk per SM seen in the conv kernels a per-SM port limit or the chip-wide L2 cap divided by 148?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probes/tma_probe tools/probes/tma_probe.cu -lcuda
#include <cstdio>
#include <vector>
#include "../../video_diffusion_nnx_b200/csrc/vdn_common.cuh"
namespace vdn { void set_last_error(const char*, ...) {} int check_launch(const char*) { return 0; } }
using namespace vdn;

__device__ __forceinline__ void mbar_wait_poll(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

// K boxes per stage barrier; poll = 1 uses mbarrier.test_wait (pure polling) instead of try_wait
__global__ void __launch_bounds__(64) probe(const __grid_constant__ CUtensorMap map, int rows_per_box, int n_boxes, int S,
                                            int boxes_in_buffer, long long* cycles, int K, int poll) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[16];
  uint8_t* smem = smem_raw + (((smem_u32(smem_raw) + 1023u) & ~1023u) - smem_u32(smem_raw));
  const int box_bytes = rows_per_box * 128;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    // keep S boxes in flight: issue box i into slot i % S once box i - S has landed
    const int n_groups = n_boxes / K;
    int box = (int)((blockIdx.x * 131u) % (unsigned)boxes_in_buffer);  // no division in the issue loop
    int st = 0;
    uint32_t ph = 0;  // parity of the phase the NEXT wait on slot st has to see completed
    uint32_t slot_addr = smem_u32(smem);
    const uint32_t slot_bytes = (uint32_t)(box_bytes * K);
    bool primed = false;
    for (int i = 0; i < n_groups + S; ++i) {
      if (primed) mbar_wait(&full[st], ph);
      if (i < n_groups) {
        mbar_expect_tx(&full[st], slot_bytes);
        for (int k = 0; k < K; ++k) {
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                  slot_addr + (uint32_t)(k * box_bytes)),
              "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&full[st])), "r"(0), "r"(box * rows_per_box)
              : "memory");
          if (++box == boxes_in_buffer) box = 0;
        }
      }
      slot_addr += slot_bytes;
      if (++st == S) {
        st = 0;
        slot_addr = smem_u32(smem);
        if (primed) ph ^= 1u;
        primed = true;
      }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  const size_t buf_bytes = 48u << 20;  // L2-resident working set
  uint8_t* buf;
  cudaMalloc(&buf, buf_bytes);
  cudaMemset(buf, 1, buf_bytes);
  long long* cyc;
  cudaMalloc(&cyc, 148 * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("rows_per_box K poll stages  ctas  B/clk/SM(avg)  chip B/clk\n");
  for (int rows : {32, 64, 128, 256}) {
    const cuuint64_t dims[2] = {64, buf_bytes / 128};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {64, (cuuint32_t)rows};
    const cuuint32_t es[2] = {1, 1};
    CUtensorMap map;
    CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    const int box_bytes = rows * 128;
    const int boxes_in_buffer = (int)(buf_bytes / box_bytes);
    for (int K : {1, 2})
    for (int poll : {0})
    for (int S : {2, 4, 8}) {
      if ((size_t)S * K * box_bytes > 190 * 1024) continue;
      for (int ctas : {1, 40, 80, 148}) {
        const int n_boxes = (int)((8u << 20) / box_bytes);  // 8 MB per CTA
        for (int rep = 0; rep < 2; ++rep)
          probe<<<ctas, 64, (size_t)S * K * box_bytes + 1024>>>(map, rows, n_boxes, S, boxes_in_buffer, cyc, K, poll);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> h(ctas);
        cudaMemcpy(h.data(), cyc, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
        double sum = 0, mx = 0;
        for (long long v : h) { sum += (double)v; mx = v > mx ? (double)v : mx; }
        const double bytes = (double)n_boxes * box_bytes;
        printf("%12d %d %4d %6d %5d  %12.1f  %10.0f\n", rows, K, poll, S, ctas, bytes / (sum / ctas), bytes * ctas / mx);
      }
    }
  }
  return 0;
}
