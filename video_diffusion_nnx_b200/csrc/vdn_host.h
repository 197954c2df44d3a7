// Internal host-side declarations shared by the .cu translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vdn.h"

namespace vdn {

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

// conv3x3_rows.cu: persistent row-ring (1,3,3) conv; applicable() decides, launch() has vdn_tapgemm's contract
bool rowconv_applicable(const vdn_tapgemm_desc* d, const void* residual, const float* gn_sums);
int rowconv_launch(const vdn_tapgemm_desc* d, const void* src0, const void* src1, const void* wp, const float* bias,
                   const void* residual, const void* residual2, void* out, void* out2, float* gn_sums,
                   cudaStream_t st);

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int num_sms() { return 148; }

}  // namespace vdn
