// Hardware-behaviour probes (tests / bring-up only, never on the product path).
//
// rsb_debug_umma answers two questions about tcgen05 shared-memory descriptors on sm_100a that the public
// guides leave open and that decide the design of the next kernels:
//   (1) K-major SWIZZLE_128B operand whose start address is NOT 1024-byte aligned (start = tile + r*128 B):
//       does a row-shifted window of a larger TMA-written box work, and which `base_offset` does it need?
//       (line-buffer convolution: one halo row in shared memory serves all three horizontal taps)
//   (2) MN-major SWIZZLE_128B A operand (rows = K index, 64 M-elements contiguous per row), as produced by a TMA
//       load of an NHWC activation tile: which LBO / SBO encode it? (wgrad contracts over pixels)

#include <cuda_fp16.h>

#include "../../include/rsb200.h"
#include "../../include/rsb200_debug.h"
#include "rsb_host.h"
#include "rsb_ptx.cuh"

namespace rsb {

struct DebugParams {
    CUtensorMap tmA;  // 2D [rows][64] fp16, box {64, a_rows}
    CUtensorMap tmB;  // 2D [64 n][64 k] fp16, box {64, 64}
    float* out;       // [128][64] fp32
    int a_rows;       // rows loaded for A (<= 256)
    int a_blocks;     // number of A boxes loaded back to back (MN-major: 2 blocks of 64 M columns)
    int mode;         // 0: K-major shifted window; 1: MN-major A
    int row_offset;   // mode 0: window start row
    int base_offset;  // value for descriptor bits [49,52)
    int lbo, sbo;     // mode 1: descriptor LBO / SBO in bytes
    int k_step_bytes; // mode 1: start-address advance per K=16 MMA
};

__global__ void __launch_bounds__(128, 1) debug_umma_kernel(const __grid_constant__ DebugParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sa = smem;                  // up to 2 x 256 rows x 128 B = 64 KB
    uint8_t* sb = smem + 65536;          // 64 x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536 + 8192);
    uint64_t* mma_bar = bar + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(mma_bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc<64>(tmem_ptr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, p.a_blocks * p.a_rows * 128 + 64 * 128);
        for (int b = 0; b < p.a_blocks; ++b) tma_load_2d(sa + b * p.a_rows * 128, &p.tmA, bar, b * 64, 0);
        tma_load_2d(sb, &p.tmB, bar, 0, 0);
        mbar_wait(bar, 0);
        tc_fence_after();
        const uint64_t db = make_sw128_kmajor_desc(smem_u32(sb));
        for (int k = 0; k < 4; ++k) {
            uint64_t da;
            uint32_t idesc = make_idesc_f16(128, 64);
            if (p.mode == 0) {
                da = make_sw128_kmajor_desc(smem_u32(sa) + p.row_offset * 128) + 2 * k;
                da |= static_cast<uint64_t>(p.base_offset & 7) << 49;
            } else {
                const uint32_t addr = smem_u32(sa) + k * p.k_step_bytes;
                da = static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
                da |= static_cast<uint64_t>((p.lbo >> 4) & 0x3FFF) << 16;
                da |= static_cast<uint64_t>((p.sbo >> 4) & 0x3FFF) << 32;
                da |= static_cast<uint64_t>(1) << 46;
                da |= static_cast<uint64_t>(2) << 61;
                idesc |= (1u << 15);  // A is MN-major
            }
            umma_f16(tmem, da, db + 2 * k, idesc, k > 0 ? 1u : 0u);
        }
        umma_commit(mma_bar);
    }
    __syncthreads();
    mbar_wait(mma_bar, 0);
    tc_fence_after();
    for (int c = 0; c < 64; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, r);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) p.out[(warp * 32 + lane) * 64 + c + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tmem);
}

// tcgen05.mma issue-rate probe: `iters` K=64 blocks (4 MMAs each) on zeroed shared-memory operands, one CTA or one
// CTA pair per tile, optionally a commit per block as the conv kernel does. out[blockIdx.x] = cycles per MMA.
template <bool TWO>
__global__ void __launch_bounds__(128, 1) debug_mma_rate_kernel(float* out, int block_n, int iters, int commit_each) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sa = smem;            // 4 stages x 16 KB
    uint8_t* sb = smem + 65536;    // 4 stages x 32 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536 + 131072);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 8);
    const int warp = threadIdx.x >> 5;
    const int rank = TWO ? static_cast<int>(cluster_ctarank()) : 0;
    for (int i = threadIdx.x; i < (65536 + 131072) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1);
        mbar_fence_init();
    }
    fence_proxy_async_smem();
    if (warp == 0) {
        if (TWO) tmem_alloc_pair<256>(tmem_ptr);
        else tmem_alloc<256>(tmem_ptr);
    }
    tc_fence_before();
    if (TWO) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    if (threadIdx.x == 0 && rank == 0) {
        const bool mn_major = (commit_each & 2) != 0;  // both operands MN-major (the wgrad kernel's layout)
        commit_each &= 1;
        uint32_t idesc = make_idesc_f16(TWO ? 256 : 128, block_n);
        uint64_t da0 = make_sw128_kmajor_desc(smem_u32(sa));
        uint64_t db0 = make_sw128_kmajor_desc(smem_u32(sb));
        if (mn_major) {
            idesc |= (1u << 15) | (1u << 16);
            // rows = K (pixels), 64 M/N elements per 128-byte row; LBO = 16 KB between 64-wide column blocks, SBO = 1024 B
            da0 = (da0 & ~(0x3FFFull << 16)) | (static_cast<uint64_t>(16384 >> 4) << 16);
            db0 = (db0 & ~(0x3FFFull << 16)) | (static_cast<uint64_t>(16384 >> 4) << 16);
        }
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const int st = it & 3;
            const uint64_t da = mn_major ? da0 + static_cast<uint64_t>((st & 1) * 512) : da0 + static_cast<uint64_t>(st * (16384 >> 4));
            const uint64_t db = mn_major ? db0 + static_cast<uint64_t>((st & 1) * 512) : db0 + static_cast<uint64_t>(st * (32768 >> 4));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t ka = mn_major ? 128 * k : 2 * k;  // K=16 step: 2048 B (16 pixel rows) vs 32 B inside the swizzle row
                if (TWO) umma_f16_pair(tmem, da + ka, db + ka, idesc, 1u);
                else umma_f16(tmem, da + ka, db + ka, idesc, 1u);
            }
            if (commit_each) {
                if (TWO) umma_commit_pair(&bar[1 + (it & 3)]);
                else umma_commit(&bar[1 + (it & 3)]);
            }
        }
        if (TWO) umma_commit_pair(&bar[0]);
        else umma_commit(&bar[0]);
        mbar_wait(&bar[0], 0);
        const long long t1 = clock64();
        out[blockIdx.x] = static_cast<float>(t1 - t0) / (4.0f * iters);
    }
    tc_fence_before();
    if (TWO) cluster_sync_all();
    else __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        if (TWO) tmem_dealloc_pair<256>(tmem);
        else tmem_dealloc<256>(tmem);
    }
}

}  // namespace rsb

using namespace rsb;

extern "C" int rsb_debug_umma(const void* a, int32_t a_total_rows, int32_t a_cols, const void* b, float* out, int32_t mode,
                              int32_t a_rows, int32_t a_blocks, int32_t row_offset, int32_t base_offset, int32_t lbo, int32_t sbo,
                              int32_t k_step_bytes, void* stream) {
    if (!a || !b || !out || a_rows < 1 || a_rows > 256 || a_blocks < 1 || a_blocks > 2) return set_error(RSB_E_INVALID, "debug_umma: bad arguments");
    int rc = rsb_device_ok();
    if (rc) return rc;
    DebugParams p = {};
    {
        const uint64_t dims[2] = {(uint64_t)a_cols, (uint64_t)a_total_rows};
        const uint64_t strides[1] = {(uint64_t)a_cols * 2};
        const uint32_t box[2] = {64, (uint32_t)a_rows};
        rc = encode_tiled_f16(&p.tmA, 2, a, dims, strides, box);
        if (rc) return rc;
    }
    {
        const uint64_t dims[2] = {64, 64};
        const uint64_t strides[1] = {128};
        const uint32_t box[2] = {64, 64};
        rc = encode_tiled_f16(&p.tmB, 2, b, dims, strides, box);
        if (rc) return rc;
    }
    p.out = out;
    p.a_rows = a_rows;
    p.a_blocks = a_blocks;
    p.mode = mode;
    p.row_offset = row_offset;
    p.base_offset = base_offset;
    p.lbo = lbo;
    p.sbo = sbo;
    p.k_step_bytes = k_step_bytes;
    const int smem = 65536 + 8192 + 64 + 1024;
    cudaError_t e = cudaFuncSetAttribute(debug_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_cuda_error(e, "debug_umma attr");
    debug_umma_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(p);
    e = cudaGetLastError();
    return e == cudaSuccess ? RSB_OK : set_cuda_error(e, "debug_umma launch");
}

extern "C" int rsb_debug_mma_rate(float* out, int32_t grid, int32_t pair, int32_t block_n, int32_t iters, int32_t commit_each, void* stream) {
    if (!out || grid < 1 || iters < 1 || (pair && (grid & 1))) return set_error(RSB_E_INVALID, "debug_mma_rate: bad arguments");
    int rc = rsb_device_ok();
    if (rc) return rc;
    const int smem = 65536 + 131072 + 128 + 1024;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pair ? 1 : 0;
    cudaError_t e;
    if (pair) {
        e = cudaFuncSetAttribute(debug_mma_rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, debug_mma_rate_kernel<true>, out, block_n, iters, commit_each);
    } else {
        e = cudaFuncSetAttribute(debug_mma_rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, debug_mma_rate_kernel<false>, out, block_n, iters, commit_each);
    }
    return e == cudaSuccess ? RSB_OK : set_cuda_error(e, "debug_mma_rate launch");
}
