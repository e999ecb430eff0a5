// Host-side helpers shared by the translation units of librsb200.so.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rsb {

// record a thread-local error message and return `code`
int set_error(int code, const char* fmt, ...);
int set_cuda_error(cudaError_t e, const char* what);

// cuTensorMapEncodeTiled through cudaGetDriverEntryPoint (no link-time dependency on libcuda, so the
// library still loads on a machine without a driver). fp16 elements, 128-byte swizzle, zero OOB fill.
int encode_tiled_f16(CUtensorMap* map, int rank, const void* base, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle_bytes = 128);

int num_sms();

}  // namespace rsb
