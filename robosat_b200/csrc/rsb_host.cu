// Error reporting, device probing and TMA descriptor encoding for librsb200.so.

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/rsb200.h"
#include "rsb_host.h"

namespace rsb {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int set_cuda_error(cudaError_t e, const char* what) {
    return set_error(RSB_E_CUDA, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess) {
            fn = reinterpret_cast<PFN_encodeTiled>(p);
        }
    }
    return fn;
}

int encode_tiled_f16(CUtensorMap* map, int rank, const void* base, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle_bytes) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return set_error(RSB_E_NODEVICE, "cuTensorMapEncodeTiled driver entry point unavailable (no CUDA driver?)");
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bdim[5];
    cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    const CUtensorMapSwizzle swz = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                   : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE);
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[256];
        int o = 0;
        for (int i = 0; i < rank; ++i) o += snprintf(buf + o, sizeof(buf) - o, "%llu/%u ", (unsigned long long)dims[i], box[i]);
        o += snprintf(buf + o, sizeof(buf) - o, "| strides ");
        for (int i = 0; i + 1 < rank; ++i) o += snprintf(buf + o, sizeof(buf) - o, "%llu ", (unsigned long long)strides_bytes[i]);
        return set_error(RSB_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims/box %s)", (int)r, rank, buf);
    }
    return RSB_OK;
}

int num_sms() {
    // per device: one process may drive several GPUs (e.g. nn.DataParallel with two device_ids)
    static int sms[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!sms[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        sms[dev] = n;
    }
    return sms[dev];
}

}  // namespace rsb

using namespace rsb;

extern "C" int rsb_version(void) { return 200; }

extern "C" void rsb_abi_layout(int32_t* out4) {
    out4[0] = static_cast<int32_t>(sizeof(rsb_conv_src));
    out4[1] = static_cast<int32_t>(sizeof(rsb_conv_seg));
    out4[2] = static_cast<int32_t>(sizeof(rsb_conv_desc));
    out4[3] = static_cast<int32_t>(sizeof(rsb_rowconv_desc));
}

extern "C" const char* rsb_last_error(void) { return g_err; }

extern "C" int rsb_device_ok(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return set_error(RSB_E_NODEVICE, "no CUDA device: %s", cudaGetErrorString(e));
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaDeviceGetAttribute");
    if (major != 10) return set_error(RSB_E_NODEVICE, "device compute capability %d.x is not sm_100 (B200)", major);
    if (!get_encode()) return set_error(RSB_E_NODEVICE, "cuTensorMapEncodeTiled driver entry point unavailable");
    return RSB_OK;
}
