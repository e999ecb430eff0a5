// Implicit-GEMM convolution for sm_100a: TMA box loads -> 128B-swizzled shared memory ->
// tcgen05.mma (fp16 x fp16 -> fp32 accumulators in tensor memory) -> fused epilogue.
//
// Persistent, warp-specialised CTA (192 threads, one CTA per SM):
//   warp 0      TMA producer   one elected lane issues cp.async.bulk.tensor for the A box and the weight tile
//   warp 1      MMA issuer     one elected lane issues 4 x tcgen05.mma (K=16) per 64-wide K block; owns TMEM alloc
//   warps 2..5  epilogue       tcgen05.ld 32 rows x 32 columns per warp, bias / residual / ReLU, fp16 NHWC stores
// Pipelines: smem ring (full/empty mbarriers, TMA <-> MMA) and a 2-deep TMEM accumulator ring
// (tmem_full / tmem_empty, MMA <-> epilogue) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// A-operand addressing (what makes this a convolution rather than a GEMM) is entirely in the TMA
// coordinates: a tile is a TW x TH x TN box of pixels of one source *view*; every filter tap is the same
// box displaced by (dh, dw), zero-filled by TMA outside the view (= the conv's zero padding). Strided
// convs read parity views, the nearest x2 upsample runs as 4 output phases with 2x2 taps on the low-res
// input, channel concat is two views. See include/rsb200.h for the reference call sites this replaces.

#include <stdio.h>
#include <string.h>

#include <new>

#include "../../include/rsb200.h"
#include "rsb_host.h"
#include "rsb_ptx.cuh"

namespace rsb {

static constexpr int kBlockM = 128;
static constexpr int kBlockK = 64;  // fp16 elements = 128 bytes = one swizzle row
static constexpr int kABytes = kBlockM * kBlockK * 2;

struct alignas(64) ConvKParams {
    CUtensorMap tmA[RSB_MAX_SRCS];
    CUtensorMap tmB;
    CUtensorMap tmC[4];  // output view per phase (TMA store)
    CUtensorMap tmR;     // residual view (TMA load), phases == 1 only
    int32_t nseg;
    int32_t seg_src[RSB_MAX_SEGS];
    int32_t seg_dh[RSB_MAX_SEGS];
    int32_t seg_dw[RSB_MAX_SEGS];
    int32_t seg_cb[RSB_MAX_SEGS];
    int32_t kblocks;
    int32_t tiles_w, tiles_h, tiles_n, n_blocks, phases, total_tiles;
    int32_t pair_tiles;  // CTA-pair schedule: ceil(spatial tiles / 2) * phases * n_blocks
    int32_t TW, TH, TN;
    int32_t Wt, Ht, Nt;
    int32_t Cout;
    int32_t out_sy, out_sx;
    int64_t out_pitch_w, out_pitch_h, out_pitch_n;
    __half* out;
    const __half* residual;
    const float* bias;
    int32_t relu;
    int32_t head_classes;
    const float* head_w;
    const float* head_b;
    float* head_out;
    float acc_scale;
    // STATS instances (training forward of a conv feeding BatchNorm): per 32-row quarter of every tile, the column sums and
    // sums of squares of the fp16 outputs as stored: float [spatial tiles * 4][2][Cout]; every entry is written exactly once
    float* stats;
    // K chunking (strict precision, long K loops): the tensor core truncates its fp32 accumulator after every MMA, a bias
    // that grows with the number of accumulated MMAs. A tile's K loop is therefore cut into chunks of `kchunk` K blocks; each
    // chunk accumulates in its own tensor-memory stage and the epilogue warps add the chunk results with round-to-nearest fp32
    // adds, parked in a per-CTA scratch tile between chunks (thread-private addresses, L2 resident). 0 = one chunk.
    int32_t kchunk;
    float* scratch;  // [grid][128][BLOCK_N] fp32
};

// TWO: the tile is computed by a CTA pair (cluster of 2, tcgen05 cta_group::2): rank r owns 128 of the pair's 256 tile
// rows and stages half of the weight tile (BLOCK_N/2 rows); one MMA of the leader reads both halves, so each SM moves
// 16 KB + BLOCK_N*64 B of shared memory per K block instead of 16 KB + BLOCK_N*128 B.
// SPLIT: strict precision. Every operand is a (hi, lo) pair of fp16 planes; a pipeline stage holds both planes of the A box
// and of the weight tile (one rank-5 / rank-3 TMA box each) and every K step issues three MMAs: hi*lo, lo*hi, hi*hi.
// SPLIT == 2 (narrow tiles, BLOCK_N <= 64): the weight planes sit back to back in shared memory, so ONE MMA of width
// 2*BLOCK_N computes A_hi x [W_hi | W_lo] into [main | cross] accumulator columns and a second one of width BLOCK_N adds
// A_lo x W_hi onto the cross columns: two MMAs per K step instead of three (an M128 MMA costs the same ~66 cycles for any
// N <= 128: it is bound by the shared-memory read of the A operand), and the small cross terms no longer share -- and
// truncate -- the main accumulator. The epilogue adds main + cross in fp32.
template <int BLOCK_N, int MODE, bool HAS_RES, bool TWO = false, int SPLIT = 0>
struct ConvCfg {
    static_assert(SPLIT != 2 || (BLOCK_N <= 64 && !TWO), "the N-concatenated split schedule is for narrow single-CTA tiles");
    static constexpr int kAccCols = SPLIT == 2 ? 2 * BLOCK_N : BLOCK_N;  // tensor-memory columns of one accumulator stage
    static constexpr int kPlanes = SPLIT ? 2 : 1;
    static constexpr int kBRows = TWO ? BLOCK_N / 2 : BLOCK_N;
    static constexpr int kBBytes = kBRows * kBlockK * 2;                 // one plane
    static constexpr int kStageA = kPlanes * kABytes;
    static constexpr int kStageB = kPlanes * kBBytes;
    static constexpr int kStageBytes = kStageA + kStageB;
    // epilogue staging: the C tile leaves through shared memory in chunks of kChunk columns (TMA store), and the
    // residual tile arrives the same way (TMA load)
    static constexpr int kChunk = BLOCK_N >= 64 ? 64 : 32;
    static constexpr int kNumChunks = BLOCK_N / kChunk;
    // every epilogue warp owns its 32 rows of the tile end to end (own staging slices, own TMA stores / residual
    // loads, no cross-warp barrier): kStoreBufs store slices and kResBufs residual slices of kSliceBytes each
    // A single warp per SM sub-partition is latency-bound (ncu: ~5 cycles per issued instruction), so wide tiles get
    // two epilogue warps per TMEM lane quarter; the pair splits the tile's column chunks (even / odd). In split mode the K
    // loop of a tile is three times as long, which hides a 4-warp epilogue, and shared memory is needed for the stages.
    static constexpr int kEpiWarps = (MODE == 0 && BLOCK_N >= 128 && !SPLIT) ? 8 : 4;
    static constexpr int kChunkStride = kEpiWarps / 4;
    static constexpr int kThreads = 64 + 32 * kEpiWarps;
    static constexpr int kWarpChunkBytes = 32 * kChunk * 2;               // one plane of one warp's slice
    static constexpr int kSliceBytes = kPlanes * kWarpChunkBytes;
    static constexpr int kStoreBufs = (kEpiWarps == 8 || SPLIT) ? 1 : 2;                       // per warp
    static constexpr int kResBufs = HAS_RES ? ((kEpiWarps == 8 || SPLIT) ? 2 : 4) : 0;        // per warp
    static constexpr int kEpiBytes = MODE == 0 ? (kStoreBufs + kResBufs) * kEpiWarps * kSliceBytes : 0;
    static constexpr int kBarBytes = 512;
    static constexpr int kMaxSmem = 232448;  // 227 KB opt-in limit per CTA
    static constexpr int kAvail = kMaxSmem - 1024 - kBarBytes - kEpiBytes;
    static constexpr int kStages = (kAvail / kStageBytes) > 8 ? 8 : (kAvail / kStageBytes);
    static_assert(kStages >= 2, "pipeline too shallow");
    static constexpr int kTmemCols = (2 * kAccCols <= 32) ? 32 : (2 * kAccCols <= 64 ? 64 : (2 * kAccCols <= 128 ? 128 : (2 * kAccCols <= 256 ? 256 : 512)));
    static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + kBarBytes + 1024;
};

struct TileCoord {
    int n_blk, pa, pb, phase, w0, h0, n0;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvKParams& p, int id) {
    TileCoord t;
    t.n_blk = id % p.n_blocks;
    id /= p.n_blocks;
    t.phase = id % p.phases;
    id /= p.phases;
    t.pa = t.phase >> 1;
    t.pb = t.phase & 1;
    t.w0 = (id % p.tiles_w) * p.TW;
    id /= p.tiles_w;
    t.h0 = (id % p.tiles_h) * p.TH;
    id /= p.tiles_h;
    t.n0 = id * p.TN;
    return t;
}

// Work item `it` of this CTA -> flat tile id. One CTA per tile: identity. CTA pair: item = (spatial pair, phase, n block),
// rank r takes spatial tile 2*pair + r (a tile past the end loads zeros and stores nothing).
template <bool TWO>
__device__ __forceinline__ int item_tile(const ConvKParams& p, int it, int rank) {
    if (!TWO) return it;
    const int per = p.n_blocks * p.phases;
    return ((it / per) * 2 + rank) * per + it % per;
}

template <int BLOCK_N, int MODE, bool HAS_RES, bool TWO, int SPLIT, bool STATS = false>
__global__ void __launch_bounds__((ConvCfg<BLOCK_N, MODE, HAS_RES, TWO, SPLIT>::kThreads), 1) conv_tc_kernel(const __grid_constant__ ConvKParams p) {
    using Cfg = ConvCfg<BLOCK_N, MODE, HAS_RES, TWO, SPLIT>;
    const int rank = TWO ? static_cast<int>(cluster_ctarank()) : 0;
    const int item0 = TWO ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int item_step = TWO ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    const int num_items = TWO ? p.pair_tiles : p.total_tiles;
    extern __shared__ uint8_t smem_raw[];
    // 128B swizzle atoms repeat every 1024 bytes: tile bases must be 1024-byte aligned
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + Cfg::kStages * Cfg::kStageA;
    uint8_t* smem_c = smem + Cfg::kStages * Cfg::kStageBytes;          // store slices (1024-aligned)
    uint8_t* smem_r = smem_c + Cfg::kStoreBufs * Cfg::kEpiWarps * Cfg::kSliceBytes;  // residual slices (HAS_RES)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kEpiBytes);
    uint64_t* empty_bar = full_bar + Cfg::kStages;
    uint64_t* tmem_full_bar = empty_bar + Cfg::kStages;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint64_t* res_full_bar = tmem_empty_bar + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_full_bar + 32);  // res_full_bar[epilogue warp][buffer]

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp_idx == 0 && lane == 0) {
        for (int i = 0; i < RSB_MAX_SRCS; ++i) tma_prefetch_desc(&p.tmA[i]);
        tma_prefetch_desc(&p.tmB);
        for (int i = 0; i < Cfg::kStages; ++i) {
            mbar_init(&full_bar[i], 1);  // pair: only the leader's is used; it expects the bytes of both CTAs' boxes
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], (TWO ? 2 : 1) * Cfg::kEpiWarps);  // one arrive per epilogue warp (of both CTAs)
        }
        for (int i = 0; i < 32; ++i) mbar_init(&res_full_bar[i], 1);
        if (MODE == 0) {
            for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.tmC[i]);
            if (HAS_RES) tma_prefetch_desc(&p.tmR);
        }
        mbar_fence_init();
    }
    if (warp_idx == 1) {
        if (TWO) tmem_alloc_pair<Cfg::kTmemCols>(tmem_ptr);
        else tmem_alloc<Cfg::kTmemCols>(tmem_ptr);
    }
    tc_fence_before();
    if (TWO) cluster_sync_all();  // the peer's barriers must be initialised before anything arrives on them remotely
    // (pair: the cluster barrier already orders the allocator's shared-memory write of the TMEM address; the CTA barrier is
    // there for tools that only track CTA-scope barriers -- compute-sanitizer racecheck reported the read below otherwise)
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) may overlap
    // the tail of the previous kernel in the stream; from here on we touch memory it produced.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp_idx == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t lead_full0 = TWO ? mapa_u32(smem_u32(&full_bar[0]), 0) : 0;
            for (int it = item0; it < num_items; it += item_step) {
                const TileCoord t = decode_tile(p, item_tile<TWO>(p, it, rank));
                const int b_row = t.phase * p.Cout + t.n_blk * BLOCK_N + rank * Cfg::kBRows;
                int kb = 0;
                for (int s = 0; s < p.nseg; ++s) {
                    const CUtensorMap* tm = &p.tmA[p.seg_src[s]];
                    const int cw = t.w0 + p.seg_dw[s] + t.pb;
                    const int ch = t.h0 + p.seg_dh[s] + t.pa;
                    for (int cb = 0; cb < p.seg_cb[s]; ++cb, ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        if constexpr (TWO) {
                            // both CTAs' boxes complete on the leader's barrier
                            const uint32_t lead_full = lead_full0 + stage * 8;
                            // (the peer's bytes may land before this expect_tx: the barrier cannot complete until the
                            // leader's arrive, and the peer cannot run a ring lap ahead of the MMA's commits)
                            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
                            if constexpr (SPLIT) {
                                tma_load_5d_pair(smem_a + stage * Cfg::kStageA, tm, lead_full, cb * kBlockK, cw, ch, t.n0, 0);
                                tma_load_3d_pair(smem_b + stage * Cfg::kStageB, &p.tmB, lead_full, kb * kBlockK, b_row, 0);
                            } else {
                                tma_load_4d_pair(smem_a + stage * Cfg::kStageA, tm, lead_full, cb * kBlockK, cw, ch, t.n0);
                                tma_load_2d_pair(smem_b + stage * Cfg::kStageB, &p.tmB, lead_full, kb * kBlockK, b_row);
                            }
                        } else {
                            mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                            if constexpr (SPLIT) {
                                // one box carries both planes: [plane][TN][TH][TW][64 ch] / [plane][rows][64]
                                tma_load_5d(smem_a + stage * Cfg::kStageA, tm, &full_bar[stage], cb * kBlockK, cw, ch, t.n0, 0);
                                tma_load_3d(smem_b + stage * Cfg::kStageB, &p.tmB, &full_bar[stage], kb * kBlockK, b_row, 0);
                            } else {
                                tma_load_4d(smem_a + stage * Cfg::kStageA, tm, &full_bar[stage], cb * kBlockK, cw, ch, t.n0);
                                tma_load_2d(smem_b + stage * Cfg::kStageB, &p.tmB, &full_bar[stage], kb * kBlockK, b_row);
                            }
                        }
                        if (++stage == Cfg::kStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (rank == 0 && elect_one()) {
            constexpr uint32_t idesc = make_idesc_f16(TWO ? 2 * kBlockM : kBlockM, BLOCK_N);
            constexpr uint32_t idesc2 = make_idesc_f16(kBlockM, 2 * BLOCK_N);  // SPLIT == 2: [W_hi | W_lo] in one MMA
            // operand descriptors of stage 0; stage s adds its byte offset >> 4 to the 14-bit start-address field
            const uint64_t da0 = make_sw128_kmajor_desc(smem_u32(smem_a));
            const uint64_t db0 = make_sw128_kmajor_desc(smem_u32(smem_b));
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            const int kchunk = p.kchunk > 0 ? p.kchunk : p.kblocks;
            for (int it = item0; it < num_items; it += item_step) {
              for (int kb0 = 0; kb0 < p.kblocks; kb0 += kchunk) {
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * Cfg::kAccCols;
                const int kb1 = kb0 + kchunk < p.kblocks ? kb0 + kchunk : p.kblocks;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t da = da0 + static_cast<uint64_t>(stage * (Cfg::kStageA >> 4));
                    const uint64_t db = db0 + static_cast<uint64_t>(stage * (Cfg::kStageB >> 4));
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        // advance 16 fp16 = 32 bytes along K inside the swizzle row: +2 in 16-byte units
                        const uint32_t first = ((kb - kb0) | k) != 0 ? 1u : 0u;
                        if constexpr (SPLIT == 2) {
                            // columns [0, N) = A_hi W_hi, columns [N, 2N) = A_hi W_lo + A_lo W_hi
                            umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc2, first);
                            umma_f16(d_tmem + BLOCK_N, da + (kABytes >> 4) + 2 * k, db + 2 * k, idesc, 1u);
                        } else if constexpr (SPLIT == 1) {
                            // lo planes sit kABytes / kBBytes behind the hi planes; cross terms first (smallest magnitude)
                            const uint64_t dal = da + (kABytes >> 4), dbl = db + (Cfg::kBBytes >> 4);
                            if (TWO) {
                                umma_f16_pair(d_tmem, da + 2 * k, dbl + 2 * k, idesc, first);
                                umma_f16_pair(d_tmem, dal + 2 * k, db + 2 * k, idesc, 1u);
                                umma_f16_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, 1u);
                            } else {
                                umma_f16(d_tmem, da + 2 * k, dbl + 2 * k, idesc, first);
                                umma_f16(d_tmem, dal + 2 * k, db + 2 * k, idesc, 1u);
                                umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, 1u);
                            }
                        } else {
                            if (TWO) umma_f16_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, first);
                            else umma_f16(d_tmem, da + 2 * k, db + 2 * k, idesc, first);
                        }
                    }
                    // smem slot is free once these MMAs retire (pair: in both CTAs)
                    if (TWO) umma_commit_pair(&empty_bar[stage]);
                    else umma_commit(&empty_bar[stage]);
                    if (++stage == Cfg::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                // accumulator (of this K chunk) complete -> epilogue
                if (TWO) umma_commit_pair(&tmem_full_bar[acc]);
                else umma_commit(&tmem_full_bar[acc]);
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
              }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        const int q = warp_idx & 3;  // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;
        const int tw = row % p.TW;
        const int th = (row / p.TW) % p.TH;
        const int tn = row / (p.TW * p.TH);
        int acc = 0;
        uint32_t acc_phase = 0;
        // this warp's quarter of the tile box: rows [32q, 32q+32) -> box offset (qw, qh, qn) inside the TW x TH x TN tile
        const int qrow0 = q * 32;
        const int qw = qrow0 % p.TW, qh = (qrow0 / p.TW) % p.TH, qn = qrow0 / (p.TW * p.TH);
        const int ew = warp_idx - 2;                 // epilogue warp index
        const int chunk0 = ew >> 2;                  // first column chunk of this warp; it takes every kChunkStride-th
        uint8_t* my_c = smem_c + ew * Cfg::kStoreBufs * Cfg::kSliceBytes;
        uint8_t* my_r = smem_r + ew * (Cfg::kResBufs > 0 ? Cfg::kResBufs : 1) * Cfg::kSliceBytes;
        uint64_t* my_res_bar = res_full_bar + ew * 4;
        uint32_t wchunk = 0;          // running chunk counter of this warp (selects staging buffers)
        int res_item = item0;         // residual prefetch cursor (lane 0)
        const uint32_t lead_tmem_empty0 = TWO ? mapa_u32(smem_u32(&tmem_empty_bar[0]), 0) : 0;
        auto release_acc = [&](int a) {
            if (TWO) mbar_arrive_remote(lead_tmem_empty0 + a * 8);
            else mbar_arrive(&tmem_empty_bar[a]);
        };
        int res_chunk = chunk0;
        auto issue_residual = [&](int rb) {
            if (res_item < num_items) {
                const TileCoord rt = decode_tile(p, item_tile<TWO>(p, res_item, rank));
                mbar_expect_tx(&my_res_bar[rb], Cfg::kSliceBytes);
                if constexpr (SPLIT)
                    tma_load_5d(my_r + rb * Cfg::kSliceBytes, &p.tmR, &my_res_bar[rb], rt.n_blk * BLOCK_N + res_chunk * Cfg::kChunk,
                                rt.w0 + qw, rt.h0 + qh, rt.n0 + qn, 0);
                else
                    tma_load_4d(my_r + rb * Cfg::kSliceBytes, &p.tmR, &my_res_bar[rb], rt.n_blk * BLOCK_N + res_chunk * Cfg::kChunk,
                                rt.w0 + qw, rt.h0 + qh, rt.n0 + qn);
                res_chunk += Cfg::kChunkStride;
                if (res_chunk >= Cfg::kNumChunks) {
                    res_chunk = chunk0;
                    res_item += item_step;
                }
            }
        };
        if (MODE == 0 && HAS_RES && lane == 0) {
#pragma unroll 1
            for (int i = 0; i < Cfg::kResBufs; ++i) issue_residual(i);
        }
        for (int it = item0; it < num_items; it += item_step) {
            const TileCoord t = decode_tile(p, item_tile<TWO>(p, it, rank));
            const int w = t.w0 + tw, h = t.h0 + th, n = t.n0 + tn;
            const bool valid = (w < p.Wt) && (h < p.Ht) && (n < p.Nt);
            const unsigned vmask = STATS ? __ballot_sync(0xffffffffu, valid) : 0u;
            // K chunking: every chunk but the last is added (fp32, round to nearest) into this thread's row of the CTA's scratch tile
            const int nchunks = (MODE == 0 && p.kchunk > 0) ? (p.kblocks + p.kchunk - 1) / p.kchunk : 1;
            float* srow = nullptr;
            if (MODE == 0 && nchunks > 1) {
                srow = p.scratch + (static_cast<size_t>(blockIdx.x) * kBlockM + row) * BLOCK_N;
                for (int ch = 0; ch + 1 < nchunks; ++ch) {
                    mbar_wait(&tmem_full_bar[acc], acc_phase);
                    tc_fence_after();
                    const uint32_t ta = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * Cfg::kAccCols;
                    // same column ownership as the final pass below: a thread only ever re-reads what it wrote itself
#pragma unroll 1
                    for (int cc = chunk0 * (Cfg::kChunk / 32); cc < BLOCK_N / 32; cc += ((cc + 1) % (Cfg::kChunk / 32) == 0) ? (Cfg::kChunkStride - 1) * (Cfg::kChunk / 32) + 1 : 1) {
                        const int c = cc * 32;
                        uint32_t r[32];
                        tmem_ld_32x32(ta + c, r);
                        if constexpr (SPLIT == 2) {
                            uint32_t x[32];
                            tmem_ld_32x32(ta + BLOCK_N + c, x);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(x[j]));
                        } else {
                            tmem_ld_wait();
                        }
                        float4* sp = reinterpret_cast<float4*>(srow + c);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float4 v4 = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                                    __uint_as_float(r[4 * j + 3]));
                            if (ch > 0) {
                                const float4 o = sp[j];
                                v4.x += o.x;
                                v4.y += o.y;
                                v4.z += o.z;
                                v4.w += o.w;
                            }
                            sp[j] = v4;
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) release_acc(acc);
                    if (++acc == 2) {
                        acc = 0;
                        acc_phase ^= 1;
                    }
                }
            }
            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * Cfg::kAccCols;

            if constexpr (MODE == 0) {
                const float* bptr = p.bias ? p.bias + t.n_blk * BLOCK_N : nullptr;
#pragma unroll 1
                for (int ck = chunk0; ck < Cfg::kNumChunks; ck += Cfg::kChunkStride, ++wchunk) {
                    constexpr int kRB = Cfg::kResBufs > 0 ? Cfg::kResBufs : 1;
                    const int rb = wchunk % kRB;
                    uint8_t* cbuf = my_c + (wchunk % Cfg::kStoreBufs) * Cfg::kSliceBytes;
                    const uint8_t* rbuf = my_r + rb * Cfg::kSliceBytes;
                    if (HAS_RES) mbar_wait(&my_res_bar[rb], (wchunk / kRB) & 1);
                    // the TMA store this warp issued kStoreBufs chunks ago (same slice) must have finished reading shared memory
                    if (lane == 0) tma_store_wait_read<Cfg::kStoreBufs - 1>();
                    __syncwarp();
#pragma unroll
                    for (int half = 0; half < Cfg::kChunk / 32; ++half) {
                        const int c = ck * Cfg::kChunk + half * 32;
                        uint32_t r[32];
                        tmem_ld_32x32(taddr + c, r);
                        if constexpr (SPLIT == 2) {
                            uint32_t x[32];
                            tmem_ld_32x32(taddr + BLOCK_N + c, x);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(x[j]));
                        } else {
                            tmem_ld_wait();
                        }
                        if (srow != nullptr) {
                            const float4* sp = reinterpret_cast<const float4*>(srow + c);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 o = sp[j];
                                r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) + o.x);
                                r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + o.y);
                                r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + o.z);
                                r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + o.w);
                            }
                        }
                        if (ck + Cfg::kChunkStride >= Cfg::kNumChunks && half == Cfg::kChunk / 32 - 1) {
                            // accumulator fully drained into registers: hand the TMEM stage back to the MMA warp
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) release_acc(acc);
                        }
                        float v[32];
                        const float asc = p.acc_scale;  // power of two (1 in fast mode): exact
                        if (bptr) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bptr + c + j));
                                v[j] = fmaf(__uint_as_float(r[j]), asc, b4.x);
                                v[j + 1] = fmaf(__uint_as_float(r[j + 1]), asc, b4.y);
                                v[j + 2] = fmaf(__uint_as_float(r[j + 2]), asc, b4.z);
                                v[j + 3] = fmaf(__uint_as_float(r[j + 3]), asc, b4.w);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * asc;
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            // 16-byte unit `u` of this row inside the chunk, at its swizzled position
                            const int u = half * 4 + j;
                            const int su = Cfg::kChunk == 64 ? (u ^ (lane & 7)) : (u ^ ((lane >> 1) & 3));
                            const int soff = lane * (Cfg::kChunk * 2) + su * 16;
                            if (HAS_RES) {
                                const uint4 r4 = *reinterpret_cast<const uint4*>(rbuf + soff);
                                const __half2* h2 = reinterpret_cast<const __half2*>(&r4);
                                if constexpr (SPLIT) {
                                    const uint4 l4 = *reinterpret_cast<const uint4*>(rbuf + Cfg::kWarpChunkBytes + soff);
                                    const __half2* l2 = reinterpret_cast<const __half2*>(&l4);
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const float2 f = __half22float2(h2[e]);
                                        const float2 g = __half22float2(l2[e]);
                                        v[j * 8 + 2 * e] += f.x + g.x;  // hi + lo is exact in fp32
                                        v[j * 8 + 2 * e + 1] += f.y + g.y;
                                    }
                                } else {
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const float2 f = __half22float2(h2[e]);
                                        v[j * 8 + 2 * e] += f.x;
                                        v[j * 8 + 2 * e + 1] += f.y;
                                    }
                                }
                            }
                            uint4 o4, q4;
                            __half2* o2 = reinterpret_cast<__half2*>(&o4);
                            __half2* q2 = reinterpret_cast<__half2*>(&q4);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float a = v[j * 8 + 2 * e], b = v[j * 8 + 2 * e + 1];
                                if (p.relu) {
                                    a = fmaxf(a, 0.0f);
                                    b = fmaxf(b, 0.0f);
                                }
                                o2[e] = __floats2half2_rn(a, b);
                                if constexpr (SPLIT) {
                                    const float2 hf = __half22float2(o2[e]);
                                    q2[e] = __floats2half2_rn(a - hf.x, b - hf.y);  // the residue is exact in fp32
                                }
                            }
                            *reinterpret_cast<uint4*>(cbuf + soff) = o4;
                            if constexpr (SPLIT) *reinterpret_cast<uint4*>(cbuf + Cfg::kWarpChunkBytes + soff) = q4;
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if constexpr (STATS) {
                        // BatchNorm batch statistics of this warp's 32 rows x kChunk columns, read back column-wise from the staged
                        // fp16 slice (the values the consumer will read): lane -> one 4-byte word (2 channels) of a row, conflict free
                        constexpr int kWords = Cfg::kChunk / 2;   // 32 or 16 words per staged row
                        constexpr int kGroups = 32 / kWords;      // row groups read in parallel: 1 or 2
                        const int wd = lane % kWords, g = lane / kWords;
                        float s0x = 0.f, s0y = 0.f, s1x = 0.f, s1y = 0.f;
#pragma unroll 8
                        for (int r = g; r < 32; r += kGroups) {
                            const int u = wd >> 2;
                            const int su = Cfg::kChunk == 64 ? (u ^ (r & 7)) : (u ^ ((r >> 1) & 3));
                            const __half2 hv = *reinterpret_cast<const __half2*>(cbuf + r * (Cfg::kChunk * 2) + su * 16 + (wd & 3) * 4);
                            if ((vmask >> r) & 1u) {
                                const float2 f = __half22float2(hv);
                                s0x += f.x;
                                s0y += f.y;
                                s1x = fmaf(f.x, f.x, s1x);
                                s1y = fmaf(f.y, f.y, s1y);
                            }
                        }
                        if (kGroups == 2) {
                            s0x += __shfl_xor_sync(0xffffffffu, s0x, 16);
                            s0y += __shfl_xor_sync(0xffffffffu, s0y, 16);
                            s1x += __shfl_xor_sync(0xffffffffu, s1x, 16);
                            s1y += __shfl_xor_sync(0xffffffffu, s1y, 16);
                        }
                        const int spatial = item_tile<TWO>(p, it, rank) / p.n_blocks;  // phases == 1
                        if (g == 0 && spatial < p.tiles_w * p.tiles_h * p.tiles_n) {
                            float* dst = p.stats + (static_cast<size_t>(spatial) * 4 + q) * 2 * p.Cout + t.n_blk * BLOCK_N + ck * Cfg::kChunk + 2 * wd;
                            *reinterpret_cast<float2*>(dst) = make_float2(s0x, s0y);
                            *reinterpret_cast<float2*>(dst + p.Cout) = make_float2(s1x, s1y);
                        }
                    }
                    if (lane == 0) {
                        if constexpr (SPLIT)
                            tma_store_5d(&p.tmC[t.phase], cbuf, t.n_blk * BLOCK_N + ck * Cfg::kChunk, t.w0 + qw, t.h0 + qh, t.n0 + qn, 0);
                        else
                            tma_store_4d(&p.tmC[t.phase], cbuf, t.n_blk * BLOCK_N + ck * Cfg::kChunk, t.w0 + qw, t.h0 + qh, t.n0 + qn);
                        tma_store_commit();
                        // residual slice `rb` has been consumed by the whole warp: refill it kResBufs chunks ahead
                        if (HAS_RES) issue_residual(rb);
                    }
                }
            } else {
                // head: ReLU(acc) [32 ch] -> fp32 1x1 conv to `classes` logits, NCHW fp32 store
                static_assert(MODE == 0 || BLOCK_N == 32, "head mode needs BLOCK_N == 32");
                uint32_t r[32];
                tmem_ld_32x32(taddr, r);
                if constexpr (SPLIT == 2) {
                    uint32_t x[32];
                    tmem_ld_32x32(taddr + BLOCK_N, x);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(x[j]));
                } else {
                    tmem_ld_wait();
                }
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float a = __uint_as_float(r[j]) * p.acc_scale;
                    if (p.bias) a += __ldg(p.bias + j);
                    v[j] = p.relu ? fmaxf(a, 0.0f) : a;
                }
                for (int k = 0; k < p.head_classes; ++k) {
                    float s = __ldg(p.head_b + k);
#pragma unroll
                    for (int j = 0; j < 32; ++j) s = fmaf(__ldg(p.head_w + k * 32 + j), v[j], s);
                    if (valid) {
                        p.head_out[((static_cast<int64_t>(n) * p.head_classes + k) * p.Ht + h) * p.Wt + w] = s;
                    }
                }
            }

            if constexpr (MODE == 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) release_acc(acc);
            }
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
        if (MODE == 0 && lane == 0) tma_store_wait<0>();  // all output slices written before the CTA retires
    }

    tc_fence_before();
    if (TWO) cluster_sync_all();  // neither CTA may retire (or free tensor memory) while the pair's MMAs can still touch it
    else __syncthreads();
    if (warp_idx == 1) {
        tc_fence_after();
        if (TWO) tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
        else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------
// SIMT checker: the same contraction read straight from global memory (tests only).
__global__ void conv_simt_check_kernel(rsb_conv_desc d, int K) {
    const int64_t total_px = static_cast<int64_t>(d.Nt) * d.Ht * d.Wt * d.phases;
    const int co_per_thread = d.mode == 1 ? d.Cout : 1;
    const int co_threads = d.Cout / co_per_thread;
    const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (gid >= total_px * co_threads) return;
    const int co0 = static_cast<int>(gid % co_threads) * co_per_thread;
    int64_t px = gid / co_threads;
    const int phase = static_cast<int>(px % d.phases);
    px /= d.phases;
    const int w = static_cast<int>(px % d.Wt);
    px /= d.Wt;
    const int h = static_cast<int>(px % d.Ht);
    const int n = static_cast<int>(px / d.Ht);
    const int pa = phase >> 1, pb = phase & 1;
    const __half* wts = static_cast<const __half*>(d.weights);
    const int64_t w_plane = static_cast<int64_t>(d.phases) * d.Cout * K;

    float head_acc[8];
    for (int k = 0; k < 8; ++k) head_acc[k] = (d.mode == 1 && k < d.head_classes) ? d.head_b[k] : 0.f;

    for (int co = co0; co < co0 + co_per_thread; ++co) {
        const __half* wrow = wts + (static_cast<int64_t>(phase) * d.Cout + co) * K;
        float acc = 0.f;
        int kidx = 0;
        for (int s = 0; s < d.nseg; ++s) {
            const rsb_conv_src& src = d.srcs[d.segs[s].src];
            const int hh = h + d.segs[s].dh + pa, ww = w + d.segs[s].dw + pb;
            const bool inb = hh >= 0 && hh < src.H && ww >= 0 && ww < src.W && n < src.N;
            const __half* a = static_cast<const __half*>(src.ptr) + n * src.pitch_n + hh * src.pitch_h + ww * src.pitch_w;
            for (int c = 0; c < d.segs[s].cblocks * 64; ++c, ++kidx) {
                if (inb && c < src.C) {
                    float av = __half2float(a[c]), wv = __half2float(wrow[kidx]);
                    if (d.split) {
                        av += __half2float(a[c + src.plane]);
                        wv += __half2float(wrow[kidx + w_plane]);
                    }
                    acc = fmaf(av, wv, acc);
                }
            }
        }
        acc *= d.acc_scale != 0.f ? d.acc_scale : 1.f;
        if (d.bias) acc += d.bias[co];
        const int64_t off = n * d.out_pitch_n + static_cast<int64_t>(h * d.out_sy + pa) * d.out_pitch_h +
                            static_cast<int64_t>(w * d.out_sx + pb) * d.out_pitch_w + co;
        if (d.mode == 0) {
            if (d.residual) {
                float rv = __half2float(static_cast<const __half*>(d.residual)[off]);
                if (d.split) rv += __half2float(static_cast<const __half*>(d.residual)[off + d.res_plane]);
                acc += rv;
            }
            if (d.relu) acc = fmaxf(acc, 0.f);
            const __half hi = __float2half_rn(acc);
            static_cast<__half*>(d.out)[off] = hi;
            if (d.split) static_cast<__half*>(d.out)[off + d.out_plane] = __float2half_rn(acc - __half2float(hi));
        } else {
            if (d.relu) acc = fmaxf(acc, 0.f);
            for (int k = 0; k < d.head_classes; ++k) head_acc[k] = fmaf(d.head_w[k * 32 + co], acc, head_acc[k]);
        }
    }
    if (d.mode == 1) {
        for (int k = 0; k < d.head_classes; ++k)
            d.head_out[((static_cast<int64_t>(n) * d.head_classes + k) * d.Ht + h) * d.Wt + w] = head_acc[k];
    }
}

}  // namespace rsb

// ================================================================================================
// host side
// ================================================================================================
using namespace rsb;

struct rsb_conv_plan {
    ConvKParams kp;
    int block_n;
    int mode;
    int grid;
    int smem;
    bool has_res;
    bool pair;   // CTA-pair (cta_group::2) schedule
    bool split;  // strict precision (hi + lo planes)
    bool stats;  // the epilogue also writes per-quarter-tile BatchNorm partial sums
};

static int validate_desc(const rsb_conv_desc* d, int* K_out) {
    if (!d) return set_error(RSB_E_INVALID, "conv: null desc");
    if (d->nsrc < 1 || d->nsrc > RSB_MAX_SRCS) return set_error(RSB_E_INVALID, "conv: nsrc out of range");
    if (d->nseg < 1 || d->nseg > RSB_MAX_SEGS) return set_error(RSB_E_INVALID, "conv: nseg out of range");
    if (!(d->block_n == 32 || d->block_n == 64 || d->block_n == 128 || d->block_n == 256))
        return set_error(RSB_E_INVALID, "conv: block_n must be 32/64/128/256");
    if (d->Cout <= 0 || d->Cout % d->block_n) return set_error(RSB_E_INVALID, "conv: Cout not a multiple of block_n");
    if (!(d->phases == 1 || d->phases == 4)) return set_error(RSB_E_INVALID, "conv: phases must be 1 or 4");
    if (d->TW * d->TH * d->TN != kBlockM) return set_error(RSB_E_INVALID, "conv: TW*TH*TN must be 128");
    if (d->Wt <= 0 || d->Ht <= 0 || d->Nt <= 0) return set_error(RSB_E_INVALID, "conv: empty tile space");
    if (!d->weights) return set_error(RSB_E_INVALID, "conv: null weights");
    if (d->split != 0 && d->split != 1) return set_error(RSB_E_INVALID, "conv: split must be 0 or 1");
    int kblocks = 0;
    for (int s = 0; s < d->nseg; ++s) {
        const rsb_conv_seg& g = d->segs[s];
        if (g.src < 0 || g.src >= d->nsrc || g.cblocks < 1) return set_error(RSB_E_INVALID, "conv: bad segment");
        if (g.cblocks * kBlockK > ((d->srcs[g.src].C + kBlockK - 1) / kBlockK) * kBlockK)
            return set_error(RSB_E_INVALID, "conv: segment reads past the source's inner extent");
        kblocks += g.cblocks;
    }
    for (int i = 0; i < d->nsrc; ++i) {
        const rsb_conv_src& s = d->srcs[i];
        if (!s.ptr || (reinterpret_cast<uintptr_t>(s.ptr) & 15)) return set_error(RSB_E_INVALID, "conv: source pointer null or not 16B aligned");
        if ((s.pitch_w * 2) % 16 || (s.pitch_h * 2) % 16 || (s.pitch_n * 2) % 16)
            return set_error(RSB_E_INVALID, "conv: source pitches must be multiples of 16 bytes");
        if (s.C < 1 || s.W < 1 || s.H < 1 || s.N < 1) return set_error(RSB_E_INVALID, "conv: empty source view");
        if (d->split && (s.plane <= 0 || (s.plane * 2) % 16)) return set_error(RSB_E_INVALID, "conv: split needs a positive 16B-multiple plane stride per source");
    }
    if (d->mode == 0) {
        if (!d->out) return set_error(RSB_E_INVALID, "conv: null out");
        if ((reinterpret_cast<uintptr_t>(d->out) & 15) || (d->out_pitch_w * 2) % 16 || (d->out_pitch_h * 2) % 16 ||
            (d->out_pitch_n * 2) % 16)
            return set_error(RSB_E_INVALID, "conv: out must be 16B aligned with 16B-multiple pitches");
        if (d->residual && (reinterpret_cast<uintptr_t>(d->residual) & 15))
            return set_error(RSB_E_INVALID, "conv: residual must be 16B aligned");
        if (d->split && (d->out_plane <= 0 || (d->out_plane * 2) % 16)) return set_error(RSB_E_INVALID, "conv: split needs out_plane");
        if (d->split && d->residual && (d->res_plane <= 0 || (d->res_plane * 2) % 16)) return set_error(RSB_E_INVALID, "conv: split needs res_plane");
        if (d->split && d->residual && d->block_n == 256 && !d->cta_pair)
            return set_error(RSB_E_INVALID, "conv: split + residual supports block_n <= 128 (or a CTA pair)");
    } else if (d->mode == 1) {
        if (d->block_n != 32 || d->Cout != 32 || d->phases != 1) return set_error(RSB_E_INVALID, "conv: head mode needs Cout == block_n == 32, phases == 1");
        if (d->head_classes < 1 || d->head_classes > 8 || !d->head_w || !d->head_b || !d->head_out)
            return set_error(RSB_E_INVALID, "conv: bad head arguments");
    } else {
        return set_error(RSB_E_INVALID, "conv: unknown mode");
    }
    if (d->cta_pair && (d->mode != 0 || d->block_n < 128)) return set_error(RSB_E_INVALID, "conv: cta_pair needs mode 0 and block_n >= 128");
    if (d->stats) {
        if (d->mode != 0 || d->split || d->residual || d->phases != 1)
            return set_error(RSB_E_INVALID, "conv: stats needs mode 0, fast precision, no residual, phases == 1");
        const int64_t tiles = static_cast<int64_t>((d->Wt + d->TW - 1) / d->TW) * ((d->Ht + d->TH - 1) / d->TH) * ((d->Nt + d->TN - 1) / d->TN);
        const int64_t need = tiles * 4 * 2 * d->Cout * 4;
        if ((reinterpret_cast<uintptr_t>(d->stats) & 7) || d->stats_bytes < need)
            return set_error(RSB_E_INVALID, "conv: stats needs an 8B-aligned buffer of >= %lld bytes (tiles x 4 x 2 x Cout fp32)", (long long)need);
    }
    *K_out = kblocks * kBlockK;
    return RSB_OK;
}

static constexpr int kMaxDevices = 64;
static int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
    return dev;
}

template <int BLOCK_N, int MODE, bool HAS_RES, bool TWO, int SPLIT, bool STATS = false>
struct ConvInst {
    using Cfg = ConvCfg<BLOCK_N, MODE, HAS_RES, TWO, SPLIT>;
    static constexpr int kSmem = Cfg::kSmemBytes;

    // the opt-in shared-memory limit is a per-device function attribute: one process may drive several GPUs
    static int prepare() {
        static bool attr_set[kMaxDevices] = {};
        const int dev = current_device();
        if (!attr_set[dev]) {
            cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, MODE, HAS_RES, TWO, SPLIT, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
            if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(conv)");
            attr_set[dev] = true;
        }
        return RSB_OK;
    }

    static int launch(const rsb_conv_plan* plan, cudaStream_t stream) {
        int rc = prepare();
        if (rc) return rc;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(plan->grid);
        cfg.blockDim = dim3(Cfg::kThreads);
        cfg.dynamicSmemBytes = kSmem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (TWO) {
            attr[1].id = cudaLaunchAttributeClusterDimension;
            attr[1].val.clusterDim.x = 2;
            attr[1].val.clusterDim.y = 1;
            attr[1].val.clusterDim.z = 1;
            cfg.numAttrs = 2;
        }
        cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<BLOCK_N, MODE, HAS_RES, TWO, SPLIT, STATS>, plan->kp);
        if (e != cudaSuccess) return set_cuda_error(e, "conv_tc_kernel launch");
        return RSB_OK;
    }

    // how many CTA pairs the device can hold at once (pairs need two SMs of one TPC; persistent grid = that many clusters)
    static int pair_clusters() {
        static int cached[kMaxDevices];
        static bool have[kMaxDevices] = {};
        const int dev = current_device();
        if (have[dev]) return cached[dev];
        if (prepare() != RSB_OK) return 0;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * num_sms());
        cfg.blockDim = dim3(Cfg::kThreads);
        cfg.dynamicSmemBytes = kSmem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, conv_tc_kernel<BLOCK_N, MODE, HAS_RES, TWO, SPLIT, STATS>, &cfg) != cudaSuccess) {
            cudaGetLastError();
            n = 0;
        }
        const int cap = num_sms() / 2;
        cached[dev] = n < cap ? n : cap;
        have[dev] = true;
        return cached[dev];
    }
};

// what to do with the instantiation a plan selects
enum ConvAction { kLaunch, kSmemOf, kPairClusters };

template <int BLOCK_N, int MODE, bool HAS_RES, bool TWO, int SPLIT, bool STATS = false>
static int conv_act(ConvAction a, const rsb_conv_plan* plan, cudaStream_t stream) {
    using I = ConvInst<BLOCK_N, MODE, HAS_RES, TWO, SPLIT, STATS>;
    switch (a) {
        case kLaunch: return I::launch(plan, stream);
        case kSmemOf: return I::kSmem;
        default: return I::pair_clusters();
    }
}

// SPLIT: 0 = fast; 1 = strict. Narrow strict tiles (block_n <= 64, never paired) take the N-concatenated schedule (2).
template <int SPLIT>
static int conv_dispatch(ConvAction a, const rsb_conv_plan* plan, cudaStream_t stream) {
    constexpr int NARROW = SPLIT ? 2 : 0;
    if (plan->mode == 1) return conv_act<32, 1, false, false, NARROW>(a, plan, stream);
    if (plan->pair) {
        if (plan->block_n == 128) return plan->has_res ? conv_act<128, 0, true, true, SPLIT>(a, plan, stream) : conv_act<128, 0, false, true, SPLIT>(a, plan, stream);
        return plan->has_res ? conv_act<256, 0, true, true, SPLIT>(a, plan, stream) : conv_act<256, 0, false, true, SPLIT>(a, plan, stream);
    }
    if (plan->has_res) {
        switch (plan->block_n) {
            case 32: return conv_act<32, 0, true, false, NARROW>(a, plan, stream);
            case 64: return conv_act<64, 0, true, false, NARROW>(a, plan, stream);
            case 128: return conv_act<128, 0, true, false, SPLIT>(a, plan, stream);
            case 256:
                if constexpr (!SPLIT) return conv_act<256, 0, true, false, 0>(a, plan, stream);
                break;  // split + residual + 256 columns does not fit one SM (rejected by validate_desc)
        }
    } else {
        switch (plan->block_n) {
            case 32: return conv_act<32, 0, false, false, NARROW>(a, plan, stream);
            case 64: return conv_act<64, 0, false, false, NARROW>(a, plan, stream);
            case 128: return conv_act<128, 0, false, false, SPLIT>(a, plan, stream);
            case 256: return conv_act<256, 0, false, false, SPLIT>(a, plan, stream);
        }
    }
    return a == kLaunch ? set_error(RSB_E_INVALID, "conv: unsupported block_n / residual / split combination") : 0;
}

// training forward of a conv that feeds BatchNorm: fast precision, no residual, mode 0 (validate_desc)
static int conv_dispatch_stats(ConvAction a, const rsb_conv_plan* plan, cudaStream_t stream) {
    if (plan->pair) {
        if (plan->block_n == 128) return conv_act<128, 0, false, true, 0, true>(a, plan, stream);
        return conv_act<256, 0, false, true, 0, true>(a, plan, stream);
    }
    switch (plan->block_n) {
        case 32: return conv_act<32, 0, false, false, 0, true>(a, plan, stream);
        case 64: return conv_act<64, 0, false, false, 0, true>(a, plan, stream);
        case 128: return conv_act<128, 0, false, false, 0, true>(a, plan, stream);
        default: return conv_act<256, 0, false, false, 0, true>(a, plan, stream);
    }
}

static int conv_select(ConvAction a, const rsb_conv_plan* plan, cudaStream_t stream) {
    if (plan->stats) return conv_dispatch_stats(a, plan, stream);
    return plan->split ? conv_dispatch<1>(a, plan, stream) : conv_dispatch<0>(a, plan, stream);
}

extern "C" int rsb_conv_plan_create(const rsb_conv_desc* d, rsb_conv_plan** out_plan) {
    if (!out_plan) return set_error(RSB_E_INVALID, "conv: null out_plan");
    *out_plan = nullptr;
    int K = 0;
    int rc = validate_desc(d, &K);
    if (rc) return rc;
    rc = rsb_device_ok();
    if (rc) return rc;

    rsb_conv_plan* plan = new (std::nothrow) rsb_conv_plan();
    if (!plan) return set_error(RSB_E_INVALID, "conv: out of host memory");
    memset(&plan->kp, 0, sizeof(plan->kp));
    ConvKParams& kp = plan->kp;
    const bool split = d->split != 0;
    // split precision: every view gains an outermost "plane" dimension of extent 2 and every box takes both planes
    const int xr = split ? 1 : 0;

    for (int i = 0; i < d->nsrc; ++i) {
        const rsb_conv_src& s = d->srcs[i];
        const uint64_t dims[5] = {(uint64_t)s.C, (uint64_t)s.W, (uint64_t)s.H, (uint64_t)s.N, 2};
        const uint64_t strides[4] = {(uint64_t)s.pitch_w * 2, (uint64_t)s.pitch_h * 2, (uint64_t)s.pitch_n * 2, (uint64_t)s.plane * 2};
        const uint32_t box[5] = {(uint32_t)kBlockK, (uint32_t)d->TW, (uint32_t)d->TH, (uint32_t)d->TN, 2};
        rc = encode_tiled_f16(&kp.tmA[i], 4 + xr, s.ptr, dims, strides, box);
        if (rc) {
            delete plan;
            return rc;
        }
    }
    for (int i = d->nsrc; i < RSB_MAX_SRCS; ++i) kp.tmA[i] = kp.tmA[0];
    {
        const uint64_t rows = (uint64_t)d->phases * d->Cout;
        const uint64_t dims[3] = {(uint64_t)K, rows, 2};
        const uint64_t strides[2] = {(uint64_t)K * 2, rows * (uint64_t)K * 2};
        // a CTA of a pair stages half of the weight tile
        const uint32_t box[3] = {(uint32_t)kBlockK, (uint32_t)(d->cta_pair ? d->block_n / 2 : d->block_n), 2};
        rc = encode_tiled_f16(&kp.tmB, 2 + xr, d->weights, dims, strides, box);
        if (rc) {
            delete plan;
            return rc;
        }
    }
    if (d->mode == 0) {
        // output (and residual) views in tile space: pixel (w, h, n) of phase (a, b) lives at
        // out + n*pitch_n + (h*sy + a)*pitch_h + (w*sx + b)*pitch_w; stored / loaded in 64- (or 32-) channel chunks
        const int sy = d->out_sy > 0 ? d->out_sy : 1, sx = d->out_sx > 0 ? d->out_sx : 1;
        const int chunk = d->block_n >= 64 ? 64 : 32;
        const uint64_t dims[5] = {(uint64_t)d->Cout, (uint64_t)d->Wt, (uint64_t)d->Ht, (uint64_t)d->Nt, 2};
        uint64_t strides[4] = {(uint64_t)sx * d->out_pitch_w * 2, (uint64_t)sy * d->out_pitch_h * 2, (uint64_t)d->out_pitch_n * 2,
                               (uint64_t)d->out_plane * 2};
        // one box per epilogue warp = 32 consecutive tile rows
        const int bw = d->TW < 32 ? d->TW : 32;
        const int bh = (32 / bw) < d->TH ? (32 / bw) : d->TH;
        const int bn = 32 / (bw * bh);
        if (32 % d->TW && d->TW % 32) {
            delete plan;
            return set_error(RSB_E_INVALID, "conv: TW must divide or be a multiple of 32");
        }
        const uint32_t box[5] = {(uint32_t)chunk, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn, 2};
        for (int ph = 0; ph < 4; ++ph) {
            const int a = ph >> 1, b = ph & 1;
            const __half* base = static_cast<const __half*>(d->out) + (ph < d->phases ? a * d->out_pitch_h + b * d->out_pitch_w : 0);
            rc = encode_tiled_f16(&kp.tmC[ph], 4 + xr, base, dims, strides, box, chunk * 2);
            if (rc) {
                delete plan;
                return rc;
            }
        }
        if (d->residual) {
            if (d->phases != 1) {
                delete plan;
                return set_error(RSB_E_INVALID, "conv: residual needs phases == 1");
            }
            strides[3] = (uint64_t)d->res_plane * 2;
            rc = encode_tiled_f16(&kp.tmR, 4 + xr, d->residual, dims, strides, box, chunk * 2);
            if (rc) {
                delete plan;
                return rc;
            }
        } else {
            kp.tmR = kp.tmC[0];
        }
    } else {
        for (int ph = 0; ph < 4; ++ph) kp.tmC[ph] = kp.tmA[0];
        kp.tmR = kp.tmA[0];
    }
    kp.nseg = d->nseg;
    for (int s = 0; s < d->nseg; ++s) {
        kp.seg_src[s] = d->segs[s].src;
        kp.seg_dh[s] = d->segs[s].dh;
        kp.seg_dw[s] = d->segs[s].dw;
        kp.seg_cb[s] = d->segs[s].cblocks;
    }
    kp.kblocks = K / kBlockK;
    kp.TW = d->TW;
    kp.TH = d->TH;
    kp.TN = d->TN;
    kp.Wt = d->Wt;
    kp.Ht = d->Ht;
    kp.Nt = d->Nt;
    kp.tiles_w = (d->Wt + d->TW - 1) / d->TW;
    kp.tiles_h = (d->Ht + d->TH - 1) / d->TH;
    kp.tiles_n = (d->Nt + d->TN - 1) / d->TN;
    kp.n_blocks = d->Cout / d->block_n;
    kp.phases = d->phases;
    kp.total_tiles = kp.tiles_w * kp.tiles_h * kp.tiles_n * kp.n_blocks * kp.phases;
    kp.pair_tiles = ((kp.tiles_w * kp.tiles_h * kp.tiles_n + 1) / 2) * kp.n_blocks * kp.phases;
    kp.Cout = d->Cout;
    kp.out_sy = d->out_sy > 0 ? d->out_sy : 1;
    kp.out_sx = d->out_sx > 0 ? d->out_sx : 1;
    kp.out_pitch_w = d->out_pitch_w;
    kp.out_pitch_h = d->out_pitch_h;
    kp.out_pitch_n = d->out_pitch_n;
    kp.out = static_cast<__half*>(d->out);
    kp.residual = static_cast<const __half*>(d->residual);
    kp.bias = d->bias;
    kp.relu = d->relu;
    kp.head_classes = d->head_classes;
    kp.head_w = d->head_w;
    kp.head_b = d->head_b;
    kp.head_out = d->head_out;
    kp.acc_scale = d->acc_scale != 0.f ? d->acc_scale : 1.f;
    kp.kchunk = (d->kchunk > 0 && d->kchunk < kp.kblocks) ? d->kchunk : 0;
    kp.scratch = d->scratch;
    kp.stats = d->stats;

    plan->block_n = d->block_n;
    plan->mode = d->mode;
    plan->split = split;
    plan->stats = d->stats != nullptr;
    const int sms = num_sms();
    plan->grid = kp.total_tiles < sms ? kp.total_tiles : sms;
    plan->has_res = d->mode == 0 && d->residual != nullptr;
    plan->pair = d->cta_pair != 0;
    if (plan->pair) {
        const int clusters = conv_select(kPairClusters, plan, nullptr);
        if (clusters < 1) {
            delete plan;
            return set_error(RSB_E_CUDA, "conv: no CTA pair can be resident");
        }
        plan->grid = 2 * (kp.pair_tiles < clusters ? kp.pair_tiles : clusters);
    }
    plan->smem = conv_select(kSmemOf, plan, nullptr);
    if (kp.kchunk > 0) {
        const int64_t need = static_cast<int64_t>(plan->grid) * kBlockM * d->block_n * 4;
        if (d->mode != 0 || !d->scratch || d->scratch_bytes < need || (reinterpret_cast<uintptr_t>(d->scratch) & 15)) {
            delete plan;
            return set_error(RSB_E_INVALID, "conv: kchunk needs mode 0 and a 16B-aligned scratch of >= %lld bytes (grid %d x 128 x block_n fp32)", (long long)need,
                             plan->grid);
        }
    }
    *out_plan = plan;
    return RSB_OK;
}

extern "C" int64_t rsb_conv_scratch_bytes(int32_t block_n) { return static_cast<int64_t>(num_sms()) * kBlockM * block_n * 4; }

extern "C" void rsb_conv_plan_destroy(rsb_conv_plan* plan) { delete plan; }

extern "C" int rsb_conv_plan_info(const rsb_conv_plan* plan, int32_t* grid, int32_t* tiles, int32_t* kblocks, int32_t* smem_bytes) {
    if (!plan) return set_error(RSB_E_INVALID, "conv: null plan");
    if (grid) *grid = plan->grid;
    if (tiles) *tiles = plan->kp.total_tiles;
    if (kblocks) *kblocks = plan->kp.kblocks;
    if (smem_bytes) *smem_bytes = plan->smem;
    return RSB_OK;
}

extern "C" int rsb_conv_run(const rsb_conv_plan* plan, void* stream_) {
    if (!plan) return set_error(RSB_E_INVALID, "conv: null plan");
    return conv_select(kLaunch, plan, static_cast<cudaStream_t>(stream_));
}

extern "C" int rsb_conv_run_simt_check(const rsb_conv_desc* d, void* stream_) {
    int K = 0;
    int rc = validate_desc(d, &K);
    if (rc) return rc;
    const int64_t px = static_cast<int64_t>(d->Nt) * d->Ht * d->Wt * d->phases;
    const int64_t threads = px * (d->mode == 1 ? 1 : d->Cout);
    const int block = 128;
    const int64_t grid = (threads + block - 1) / block;
    if (grid > 0x7fffffff) return set_error(RSB_E_INVALID, "conv check: problem too large");
    conv_simt_check_kernel<<<static_cast<unsigned>(grid), block, 0, static_cast<cudaStream_t>(stream_)>>>(*d, K);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "conv_simt_check_kernel launch");
    return RSB_OK;
}
