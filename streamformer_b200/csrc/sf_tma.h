// sf_tma.h — host-side access to cuTensorMapEncodeTiled through the runtime's driver entry point
// (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace sf {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// nullptr (with the error message set) if the entry point cannot be resolved
EncodeTiledFn get_encode_fn();
int num_sms();

}  // namespace sf
