// Shared host/device declarations of the aclgan_b200 extension.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/aclgan_b200.h"
#include "sm100_ptx.cuh"

namespace aclgan {

// Encodes a bf16 / 128B-swizzle / zero-fill tiled tensor map from the plain-C spec (driver entry point is
// resolved lazily through the runtime, so the library has no link-time dependency on libcuda).
int encode_tmap(const aclgan_tmap_spec* spec, CUtensorMap* out);
int num_sms();

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

}  // namespace aclgan
