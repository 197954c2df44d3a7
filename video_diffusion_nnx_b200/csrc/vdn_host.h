// Internal host-side declarations shared by the .cu translation units.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace vdn {

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int num_sms() { return 148; }

}  // namespace vdn
