// K3 (tensor-core variant) -- patch-moment projection on tcgen05 / TMEM, fed by TMA.
// Replaces ZPs._transform_dot_product (mtflearn/features/_zps.py:146-157) and, through its
// epilogue, zmoments.to_complex / np.abs / np.angle (mtflearn/features/_zmoments.py:300-316).
//
//   D[patch, mode] = sum_k X[patch, k] * B[mode, k]          B = V/area, K-major, zero padded
//
// Roles per CTA (persistent, one CTA per SM, static round-robin over tiles of 128*S patches):
//   warp 0      TMA producer: per k-block of 32 floats, one box of X (128*S rows x 128 B) and one
//               box of B (n_pad rows x 128 B), both landing in the 128B-swizzled K-major layout
//   warp 1      MMA issuer: tcgen05.mma.kind::tf32, M=128 (patches) x N=n_pad (modes) x K=8,
//               fp32 accumulators in TMEM, S accumulators per tile, double-buffered when they fit
//   warp 2      TMEM allocator
//   warps 4-7   epilogue: tcgen05.ld the accumulator rows (thread == patch), fuse complex packing,
//               modulus, phase or the n-fold scores, store
// The kernel is HBM-bound at the metric shape: every patch byte is read once, the basis
// (n_pad x 16 KiB) is re-streamed per tile from L2.
//
// tf32x3 (fp32-grade) is a second kernel, project_tc3_kernel, with two more ideas:
//   * operand split  X.B ~= Xhi.Bhi + (Xlo.Bhi + Xhi.Blo): the first term is a kind::tf32 MMA whose
//     raw fp32 X operand the tensor core truncates to Xhi itself; the two correction terms need only
//     ~8 bits each, so they are ONE kind::f16 (bf16) MMA with K=16 = 8 x (Xlo,Bhi) + 8 x (Xhi,Blo).
//     A splitter warpgroup computes Xlo = x - trunc_tf32(x), packs bf16 [Xlo | Xhi] and writes it
//     straight into TMEM (tcgen05.st), from where the MMA reads its A operand -- no second copy of X
//     in shared memory; the matching bf16 [Bhi | Blo] operand is precomputed in HBM (op_cb);
//   * K-chunked accumulation: the tensor core accumulates in fp32 with round-toward-zero, which
//     biases a 512-step sum by ~1.5e-5 relative (measured); accumulators are therefore drained
//     every 8 k-blocks into fp32 registers of the epilogue warps (round-to-nearest adds), with two
//     accumulator buffers in TMEM so draining overlaps the next chunk's MMAs;
//   * separate X (HBM, TMA or gathered) and B (L2, multicast) rings with their own producers; with
//     p.gather the splitter warps fill the X ring themselves from a frame (K2 fused into K3).
#include "zb200_common.cuh"
#include "zb200_tc_ptx.cuh"
#include "zb200_project_shared.cuh"

#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <stdlib.h>

namespace zb200 {

namespace tc {

// experiment support (debug-hook builds only): cycles spent blocked per wait site, summed over CTAs (ZB200_TC_DEBUG & 16)
#if ZB200_DEBUG_HOOKS
__device__ unsigned long long g_wait_cycles[16];
#endif
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, unsigned long long& acc, bool on) {
    if (!on) { mbar_wait(bar, parity); return; }
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += (unsigned long long)(clock64() - t0);
}




struct Params {
    long long n_patches;
    int n_tiles;          // ceil(n_patches / (128*subtiles))
    int subtiles;         // 1 or 2 accumulators (128 patches each) per tile
    int k_blocks;         // 32-tap k-blocks that hold taps inside the unit disk (window rows entirely outside are skipped)
    const unsigned char* kmask;   // per k-block (absolute index): bit j = taps [8j, 8j+8) of the block touch the unit disk
                          //   (the tf32x3 issuer skips the MMAs of all-zero K steps: 16 % of them at k = 64)
    int kb0;              // first of them: k-block kb of the loops is taps [(kb0 + kb) * 32, +32)
    int n_pad;            // UMMA N (operand rows, multiple of 16)
    int n_cols;           // meaningful accumulator columns
    int n_stages;
    int b_stages;         // basis ring slots (tf32x3 kernel)
    int acc_bufs;         // 1 or 2
    int out_kind;         // ZB200_OUT_* or 100 = scores
    int row_len;          // output row length in floats (REAL: M, COMPLEX: 2Mc, ABS: Mc)
    float* out;
    float* out2;
    const float* w;       // [n_folds][n_pad] score weights
    const unsigned char* sel;
    int n_folds;
    int norm_kind;
    int chunk_kb;         // k-blocks per accumulation chunk (tf32x3 kernel)
    int lo_bufs;          // operand staging buffers in TMEM: 1, 2 or 4       (tf32x3 kernel)
    int epi_solo;         // 1: epilogue warpgroup 1 owns all columns, warpgroup 2 idles
    int cluster;          // CTAs per cluster sharing the B operand through TMA multicast (1, 2 or 4)
    int pair;             // 1 (tf32x3, cluster 2): cta_group::2 -- the leader CTA issues M=256 MMAs for both SMs, each CTA
                          //   stages only its half of the B rows (shared-memory traffic per k-block 168 -> 132 KB)
    int gather;           // 1: X tiles are gathered from a frame at peak windows instead of TMA-loaded (tf32x3)
    const float* g_img;   // frame [g_H][g_W]
    const float* g_planes;  // optional: 4 shifted, zero-padded copies [4][g_H][g_Wp], plane r [y][u] = img0[y][u - g_L + r]
    int g_Wp, g_L;          //   (16-byte gathers: a window row starting at column X is 16-B aligned in plane (X + g_L) % 4)
    int g_H, g_W, g_k;
    const int2* g_xy;     // top-left corner (x0, y0) of every patch window
    int dbg;              // ZB200_TC_DEBUG bitmask: experiments only (results are wrong when set)
    // K5 fused into K3: copies of the output rows go to the same rows of up to 7 peer GPUs' result arrays over
    // NVLink (P2P stores from an otherwise idle warp, tile by tile while the next tiles are being computed)
    int n_peers;
    uint32_t push_off;    // byte offset of the pusher's staging region (mbarrier + kPushChunk bytes) in dynamic shared memory
    float* peer_out[kMaxPeers];
    // f16x3 mode: inputs are multiplied by x_scale (a power of two that brings |x| <= 2^14) before the fp16 split,
    // the basis operand holds V (not V/area): results are multiplied by out_scale = 1 / (area * x_scale)
    float x_scale, out_scale;
    // conditional launch (fallback of the auto-ranged folded kernel): every CTA returns at once unless *run_if != 0
    const unsigned* run_if;
};

template <int kOut>
__device__ __forceinline__ void epilogue_chunk(const Params& p, long long row, int c0, const uint32_t (&v)[16],
                                               ScoreAcc& sc) {
    if constexpr (kOut == kOutPlain) {
        float* dst = p.out + row * (long long)p.row_len + c0;
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (c0 + i < p.row_len) dst[i] = __uint_as_float(v[i]);
    } else if constexpr (kOut == kOutAbs || kOut == kOutAbsPhase) {
        const int m0 = c0 >> 1;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (m0 + i < p.row_len) {
                const float re = __uint_as_float(v[2 * i]), im = __uint_as_float(v[2 * i + 1]);
                p.out[row * (long long)p.row_len + m0 + i] = fast_abs2(re, im);
                if constexpr (kOut == kOutAbsPhase) p.out2[row * (long long)p.row_len + m0 + i] = compact_atan2(im, re);
            }
        }
    } else {   // fused n-fold scores over real-order columns
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int c = c0 + i;
            if (c < p.n_cols && p.sel[c]) {
                const float z = __uint_as_float(v[i]), z2 = z * z, a = fabsf(z);
                sc.s1 += a;
                sc.s2 += z2;
                sc.sm = fmaxf(sc.sm, a);
#pragma unroll
                for (int f = 0; f < kFusedFolds; ++f)
                    if (f < p.n_folds) sc.num[f] = fmaf(__ldg(p.w + f * p.n_pad + c), z2, sc.num[f]);
            }
        }
    }
}

__device__ __forceinline__ void finish_scores(const Params& p, long long row, const ScoreAcc& sc) {
    float den = 1.f;
    if (p.norm_kind == ZB200_NORM_L1) den = sc.s1 * sc.s1;
    else if (p.norm_kind == ZB200_NORM_L2) den = sc.s2;
    else if (p.norm_kind == ZB200_NORM_INF) den = sc.sm * sc.sm;
#pragma unroll
    for (int f = 0; f < kFusedFolds; ++f)
        if (f < p.n_folds) p.out[row * p.n_folds + f] = sc.num[f] / den;
}

// ================================================================================================
// 1 x TF32: raw fp32 operands, the tensor core truncates them to tf32 (stated bound 1e-3 * max|Z|)
// ================================================================================================
template <int kOut>
__global__ void __launch_bounds__(256, 1)
project_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_b, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t x_bytes = (uint32_t)p.subtiles * kTileRows * 128;
    const uint32_t b_bytes = (uint32_t)p.n_pad * 128;
    const uint32_t stage_bytes = x_bytes + b_bytes;
    auto stage_x = [&](int s) { return smem + (size_t)s * stage_bytes; };
    auto stage_b = [&](int s) { return smem + (size_t)s * stage_bytes + x_bytes; };
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.n_stages * stage_bytes);
    uint64_t* full = bars;                        // TMA landed
    uint64_t* empty = bars + p.n_stages;          // MMAs reading the stage retired
    uint64_t* acc_full = bars + 2 * p.n_stages;   // accumulator complete      [2]
    uint64_t* acc_empty = acc_full + 2;           // accumulator drained       [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    unsigned* out_done = reinterpret_cast<unsigned*>(tmem_slot + 1);     // epilogue warps that finished storing a tile

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_x);
        prefetch_tmap(&map_b);
        *out_done = 0;
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.n_stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], p.cluster);      // every CTA of the cluster releases the shared B slot
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 4);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    if (p.cluster > 1) cluster_sync_all();        // peers' barriers are initialised before any multicast
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // all CTAs of a cluster run the same number of tiles (they consume B in lock-step); tiles past
    // the end are all-zero (TMA out-of-bounds fill) and their rows are never stored
    const uint32_t crank = p.cluster > 1 ? cluster_rank() : 0u;
    const uint16_t cmask = (uint16_t)((1u << p.cluster) - 1u);
    const int my_tiles = (p.n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int b_rows = p.n_pad / p.cluster;       // B rows this CTA fetches and multicasts

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int s = 0;
            uint32_t ph = 0;
            for (int t = 0; t < my_tiles; ++t) {
                const int tile = blockIdx.x + t * gridDim.x;
                const int row0 = tile * p.subtiles * kTileRows;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full[s], x_bytes + b_bytes);
                    tma_load_2d(stage_x(s), &map_x, &full[s], (p.kb0 + kb) * kBlockK, row0, kEvictFirst);
                    if (p.cluster == 1)
                        tma_load_2d(stage_b(s), &map_b, &full[s], (p.kb0 + kb) * kBlockK, 0, kEvictLast);
                    else
                        tma_load_2d_mc(stage_b(s) + (size_t)crank * b_rows * 128, &map_b, &full[s], (p.kb0 + kb) * kBlockK,
                                       (int)crank * b_rows, cmask, kEvictLast);
                    if (++s == p.n_stages) { s = 0; ph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t idesc = make_idesc_tf32(p.n_pad);
            const uint32_t x_lo0 = desc_lo_sw128(smem_u32(smem));
            const uint32_t stage_step = stage_bytes >> 4;
            int s = 0;
            uint32_t ph = 0;
            for (int t = 0; t < my_tiles; ++t) {
                const int buf = t % p.acc_bufs;
                const uint32_t acc_ph = (uint32_t)(t / p.acc_bufs) & 1u;
                mbar_wait(&acc_empty[buf], acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem0 = tmem_base + (uint32_t)(buf * p.subtiles * p.n_pad);
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t xl = x_lo0 + (uint32_t)s * stage_step;
                    const uint32_t bl = xl + (x_bytes >> 4);
                    const uint32_t acc0 = kb > 0 ? 1u : 0u;
#pragma unroll
                    for (int k4 = 0; k4 < kBlockK / kUmmaK; ++k4) {      // 32-byte steps inside the swizzle atom
                        umma_tf32(d_tmem0, desc_from_lo(xl + 2 * k4), desc_from_lo(bl + 2 * k4), idesc, k4 ? 1u : acc0);
                    }
                    if (p.subtiles == 2) {
#pragma unroll
                        for (int k4 = 0; k4 < kBlockK / kUmmaK; ++k4)
                            umma_tf32(d_tmem0 + p.n_pad, desc_from_lo(xl + 1024 + 2 * k4), desc_from_lo(bl + 2 * k4), idesc,
                                      k4 ? 1u : acc0);
                    }
                    if (p.cluster == 1) umma_commit(&empty[s]);   // stage reusable once these MMAs retire
                    else umma_commit_mc(&empty[s], cmask);
                    if (kb == p.k_blocks - 1) umma_commit(&acc_full[buf]);
                    if (++s == p.n_stages) { s = 0; ph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp - 4;                                   // TMEM lane quarter of this warp
        for (int t = 0; t < my_tiles; ++t) {
            const int tile = blockIdx.x + t * gridDim.x;
            const int buf = t % p.acc_bufs;
            const uint32_t acc_ph = (uint32_t)(t / p.acc_bufs) & 1u;
            mbar_wait(&acc_full[buf], acc_ph);
            tc_fence_after();
            for (int sub = 0; sub < p.subtiles; ++sub) {
                const long long row = (long long)tile * p.subtiles * kTileRows + sub * kTileRows + q * 32 + lane;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * p.subtiles + sub) * p.n_pad);
                ScoreAcc sc;
                if constexpr (kOut == kOutScores) sc.clear();
                for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(taddr + c0, v);
                    tmem_ld_wait();
                    if (row < p.n_patches) epilogue_chunk<kOut>(p, row, c0, v, sc);
                }
                if constexpr (kOut == kOutScores)
                    if (row < p.n_patches) finish_scores(p, row, sc);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
            if (p.n_peers) {
                __threadfence();                              // this tile's rows are in L2 before the pusher is told
                asm volatile("fence.proxy.async.global;" ::: "memory");
                __syncwarp();
                if (lane == 0) atomicAdd(out_done, 1u);
            }
        }
    } else if (warp == 3 && p.n_peers) {
        pusher_loop(p, out_done, 4, my_tiles, lane, smem + p.push_off, p.subtiles * kTileRows);
    }

    tc_fence_before();
    __syncthreads();
    if (p.cluster > 1) cluster_sync_all();        // no CTA leaves while a peer may still write into it
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// ================================================================================================
// 3 x TF32 (fp32-grade): Xlo in TMEM, K-chunked accumulation drained into registers
//   warpgroup 0  warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator      (few registers)
//   warpgroup 1,2  epilogue: running fp32 sums of up to 128 accumulator columns per thread
//   warpgroup 3  splitter: Xlo = x - trunc_tf32(x) for its 128 rows, tcgen05.st into TMEM
// ================================================================================================
#ifndef ZB200_TC3_PROF
#define ZB200_TC3_PROF ZB200_DEBUG_HOOKS
#endif
constexpr int kRegsCtl = 48, kRegsEpi = 176, kRegsSplit = 112;     // (48 + 2*176 + 112) * 128 = 64 Ki
constexpr int kRegsSplit2 = 96, kRegsEpi2 = 248;                   // kSplit2: (48 + 2*96 + 248) * 128 = 61 Ki
constexpr int kCC2 = 6;                                            // kSplit2: 6 x 16 accumulator columns per sub-tile
constexpr int kMaxColChunks = 8;     // 8 x 16 = 128 running sums per epilogue thread

// kF16 = the fp16-split arithmetic (ZB200_PREC_F16X3): x = x1 + x2, V = b1 + b2 in fp16, three kind::f16 MMAs
// (x1.b1, x2.b1, x1.b2) per 16 taps, all with A in TMEM -- 6 MMAs per 32-tap k-block and sub-tile instead of 8, no
// MMA reads X from shared memory (the splitter releases the X stage as soon as it holds the row in registers), and
// the basis stream is ONE 128-byte row per k-block ([32 x b1 | 32 x b2]) instead of two ([Bhi] + [Bcb]).
// kSplit2 = the layout for the metric shape (two 128-patch sub-tiles, <= 96 operand rows, TMA-fed X): TWO splitter
// warpgroups (warpgroup g splits sub-tile g: one row per thread and k-block instead of two, so the serial chain
// LDS -> split -> tcgen05.st -> arrive that paces the kernel under the power cap is half as long) and ONE epilogue
// warpgroup that drains both accumulators (2 x 96 running sums per thread, 248 registers).
template <int kOut, bool kPair, bool kF16, bool kSplit2>
__global__ void __launch_bounds__(512, 1)
project_tc3_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_bhi,
                   const __grid_constant__ CUtensorMap map_blo, const Params p) {
    if (p.run_if && *reinterpret_cast<const volatile unsigned*>(p.run_if) == 0u) return;     // uniform over the grid
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t x_bytes = (uint32_t)p.subtiles * kTileRows * 128;
    const uint32_t b_bytes = (uint32_t)p.n_pad * 128;
    // two rings: X (HBM stream, deep: bytes in flight hide the DRAM latency) and B (L2 stream, 2 slots)
    // [Bhi | Bcb] of one k-block; a CTA of a pair stages only its half of the rows of each
    const uint32_t bop_bytes = kPair ? b_bytes / 2 : b_bytes;
    const uint32_t bst_bytes = kF16 ? bop_bytes : 2 * bop_bytes;
    uint8_t* b_ring = smem + (size_t)p.n_stages * x_bytes;
    auto stage_x = [&](int s) { return smem + (size_t)s * x_bytes; };
    auto stage_bhi = [&](int s) { return b_ring + (size_t)s * bst_bytes; };
    auto stage_blo = [&](int s) { return b_ring + (size_t)s * bst_bytes + bop_bytes; };
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + (size_t)p.b_stages * bst_bytes);
    uint64_t* full = bars;                        // TMA landed                               [stages]
    uint64_t* empty = bars + p.n_stages;          // MMAs reading the stage retired           [stages]
    uint64_t* bfull = bars + 2 * p.n_stages;      // B k-block landed                          [4]
    uint64_t* bempty = bfull + 4;                 // MMAs reading it retired (whole cluster)    [4]
    uint64_t* lo_full = bempty + 4;               // staged operands of a k-block are in TMEM [4]
    uint64_t* lo_empty = lo_full + 4;             // MMAs reading them retired                [4]
    uint64_t* acc_full = lo_empty + 4;            // accumulation chunk complete              [2]
    uint64_t* acc_empty = acc_full + 2;           // chunk drained by the epilogue            [2]
    uint64_t* bpeer = acc_empty + 2;              // pair mode: the peer CTA's half of a B k-block landed [4]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bpeer + 4);
    unsigned* out_done = reinterpret_cast<unsigned*>(tmem_slot + 1);     // epilogue warps that finished storing a tile

    // Warp-role layout.  The SM's issue arbiter prefers the highest warp id of a sub-partition, and
    // every sub-partition hosts one busy splitter warp, so the latency-critical single-thread roles
    // (TMA producer, MMA issuer) take the LAST warpgroup and the splitter the first:
    //   warps 0-3 splitter | warps 4-11 epilogue (two warpgroups) | warp 12 TMA, 13 MMA, 14 TMEM alloc
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wg = warp >> 2;
    constexpr int kWarpTma = 12, kWarpMma = 13, kWarpAlloc = 14, kWarpBasis = 15;
    // experiment hooks (blocked-cycle counters, ablation bits) are compiled out unless -DZB200_TC3_PROF=1:
    // every extra branch or clock read in the single-thread issue loop costs issue slots (profiles/r01_maph_issue_study.md)
    const bool prof = ZB200_TC3_PROF && (p.dbg & 16) != 0;
    unsigned long long w0 = 0, w1 = 0, w2 = 0, w3 = 0;      // per-thread blocked-cycle counters (experiments)

    if (warp == kWarpTma && lane == 0) {
        prefetch_tmap(&map_x);
        prefetch_tmap(&map_bhi);
        prefetch_tmap(&map_blo);
        *out_done = 0;
    }
    if (warp == kWarpMma && lane == 0) {
        for (int s = 0; s < p.n_stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kF16 ? (kSplit2 ? 8 : 4) : 1);   // f16x3: the splitter warps release the X stage
        }
        // pair mode: the leader's MMA thread is the only committer (multicast to both CTAs) and collects the
        // splitter / epilogue arrivals of both CTAs
        for (int b = 0; b < 4; ++b) {
            mbar_init(&bfull[b], 1);
            mbar_init(&bempty[b], kPair ? 1 : p.cluster);
            mbar_init(&lo_full[b], (kPair ? 8 : 4) * (kSplit2 ? 2 : 1));
            mbar_init(&lo_empty[b], 1);
            mbar_init(&bpeer[b], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], ((p.epi_solo || kSplit2) ? 4 : 8) * (kPair ? 2 : 1));
        }
        fence_barrier_init();
    }
    if (warp == kWarpAlloc) {
        if constexpr (kPair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"(kTmemCols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"(kTmemCols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.cluster > 1) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // TMEM columns: [2 accumulator sets][lo_bufs x subtiles x (Xraw? 32 | bf16 pack 32)]
    const uint32_t lo_base = tmem_base + (uint32_t)(p.acc_bufs * p.subtiles * p.n_pad);
    const uint32_t lo_mask = (uint32_t)p.lo_bufs - 1u;               // lo_bufs is a power of two
    const uint32_t lo_shift = p.lo_bufs == 4 ? 2u : (p.lo_bufs == 2 ? 1u : 0u);

    const uint32_t crank = p.cluster > 1 ? cluster_rank() : 0u;
    const uint16_t cmask = (uint16_t)((1u << p.cluster) - 1u);
    const int my_tiles = (p.n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;     // lock-step within the cluster
    const int b_rows = p.n_pad / p.cluster;
    const int n_chunks = (p.k_blocks + p.chunk_kb - 1) / p.chunk_kb;

    if (wg == 3) {
        reg_dec<kRegsCtl>();
        if (warp == kWarpTma) {
            // ===================== TMA producer =====================
            if (elect_one()) {
                int s = 0;
                uint32_t ph = 0;
                for (int t = 0; t < my_tiles; ++t) {
                    const int tile = blockIdx.x + t * gridDim.x;
                    const int row0 = tile * p.subtiles * kTileRows;
                    for (int kb = 0; kb < p.k_blocks; ++kb) {
                        if (p.gather) break;                       // X is gathered by the splitter warps
                        mbar_wait_t(&empty[s], ph ^ 1, w0, prof);
                        mbar_arrive_expect_tx(&full[s], x_bytes);
                        tma_load_2d(stage_x(s), &map_x, &full[s], (p.kb0 + kb) * kBlockK, row0, kEvictFirst);
                        if (++s == p.n_stages) { s = 0; ph ^= 1; }
                    }
                }
            }
            __syncwarp();
        } else if (warp == kWarpBasis) {
            // ===================== basis producer (L2 -> smem, multicast across the cluster) =====================
            if (elect_one()) {
                int sb = 0;
                uint32_t phb = 0;
                for (int t = 0; t < my_tiles; ++t) {
                    for (int kb = 0; kb < p.k_blocks; ++kb) {
                        mbar_wait(&bempty[sb], phb ^ 1);
                        if (kPair) {
                            // own half of the rows only, at the start of the stage (the M=256 MMA takes N/2 rows of B
                            // from each CTA of the pair)
                            mbar_arrive_expect_tx(&bfull[sb], bst_bytes);
                            tma_load_2d(stage_bhi(sb), &map_bhi, &bfull[sb], (p.kb0 + kb) * kBlockK, (int)crank * b_rows, kEvictLast);
                            if (!kF16)
                            tma_load_2d(stage_blo(sb), &map_blo, &bfull[sb], (p.kb0 + kb) * kBlockK, (int)crank * b_rows, kEvictLast);
                            if (++sb == p.b_stages) { sb = 0; phb ^= 1; }
                            continue;
                        }
                        mbar_arrive_expect_tx(&bfull[sb], bst_bytes);
                        if (p.cluster == 1) {
                            tma_load_2d(stage_bhi(sb), &map_bhi, &bfull[sb], (p.kb0 + kb) * kBlockK, 0, kEvictLast);
                            if (!kF16) tma_load_2d(stage_blo(sb), &map_blo, &bfull[sb], (p.kb0 + kb) * kBlockK, 0, kEvictLast);
                        } else {
                            const size_t off = (size_t)crank * b_rows * 128;
                            tma_load_2d_mc(stage_bhi(sb) + off, &map_bhi, &bfull[sb], (p.kb0 + kb) * kBlockK, (int)crank * b_rows, cmask,
                                           kEvictLast);
                            if (!kF16)
                            tma_load_2d_mc(stage_blo(sb) + off, &map_blo, &bfull[sb], (p.kb0 + kb) * kBlockK, (int)crank * b_rows, cmask,
                                           kEvictLast);
                        }
                        if (++sb == p.b_stages) { sb = 0; phb ^= 1; }
                    }
                }
            }
            __syncwarp();
        } else if (warp == kWarpMma) {
            // ===================== MMA issuer =====================
            if (kPair && crank != 0) {
                // peer CTA of a pair: no MMAs to issue; relay "my half of the B k-block landed" to the leader
                if (elect_one()) {
                    int sb = 0;
                    uint32_t phb = 0;
                    for (int t = 0; t < my_tiles; ++t)
                        for (int kb = 0; kb < p.k_blocks; ++kb) {
                            mbar_wait(&bfull[sb], phb);
                            mbar_arrive_cluster(mapa_cluster(smem_u32(&bpeer[sb]), 0));
                            if (++sb == p.b_stages) { sb = 0; phb ^= 1; }
                        }
                }
                __syncwarp();
            } else
            if (elect_one()) {
                const uint32_t idesc = make_idesc_tf32(p.n_pad, kPair ? 256 : 128);
                const uint32_t idesc_c = make_idesc_bf16(p.n_pad, kPair ? 256 : 128);
                const uint32_t x_lo0 = desc_lo_sw128(smem_u32(smem));
                const uint32_t stage_step = x_bytes >> 4;
                const uint32_t b_lo0 = desc_lo_sw128(smem_u32(b_ring));
                const uint32_t b_step = bst_bytes >> 4;
                int sb = 0;
                uint32_t phb = 0;
                int s = 0;
                uint32_t ph = 0;
                uint32_t it = 0;            // running k-block counter  -> Xlo buffer / phase
                uint32_t ck = 0;            // running chunk counter    -> accumulator buffer / phase
                for (int t = 0; t < my_tiles; ++t) {
                    for (int c = 0; c < n_chunks; ++c, ++ck) {
                        // two accumulator sets alternate; with one set (wide operands: TMEM goes to the staging
                        // buffers instead) the chunk waits for its own drain
                        const int buf = p.acc_bufs == 2 ? (int)(ck & 1) : 0;
                        const uint32_t acc_par = p.acc_bufs == 2 ? ((ck >> 1) & 1u) : (ck & 1u);
                        if (kPair) mbar_wait_cluster(&acc_empty[buf], acc_par ^ 1u);
                        else mbar_wait_t(&acc_empty[buf], acc_par ^ 1u, w0, prof);
                        tc_fence_after();
                        const int kb_end = min(p.k_blocks, (c + 1) * p.chunk_kb);
                        const uint32_t d0 = tmem_base + (uint32_t)(buf * p.subtiles * p.n_pad);
                        const int kb_begin = c * p.chunk_kb;
                        uint32_t acc_run = 0u;                       // 0 until the chunk's first MMA has been issued
                        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
                            const uint32_t km = __ldg(p.kmask + p.kb0 + kb);      // issued before the waits
                            const int lb = (int)(it & lo_mask);
                            const uint32_t lo_ph = (it >> lo_shift) & 1u;
                            if (kPair) {
                                mbar_wait_cluster(&lo_full[lb], lo_ph);       // both CTAs' splitters: X landed and staged
                                mbar_wait_cluster(&bpeer[sb], phb);           // the peer's half of B
                            } else {
                                mbar_wait_t(&lo_full[lb], lo_ph, w1, prof);   // implies X landed (TMA or gathered)
                            }
                            mbar_wait(&bfull[sb], phb);
                            const long long t_fence = prof ? clock64() : 0;
                            tc_fence_after();
                            const long long t_issue = prof ? clock64() : 0;
                            if (prof) w0 += (unsigned long long)(t_issue - t_fence);   // (reported as mma.acc_empty+fence)
                            const uint32_t xl = x_lo0 + (uint32_t)s * stage_step;
                            const uint32_t bhl = b_lo0 + (uint32_t)sb * b_step;
                            const uint32_t bcl = bhl + (bop_bytes >> 4);
                            const uint32_t a0 = lo_base + (uint32_t)(lb * p.subtiles) * kBlockK;
                            if constexpr (kF16) {
                                // fp16 split: per 16 taps x1.b1, x2.b1, x1.b2 -- A from TMEM (x1 in columns [0,16) of the
                                // staging buffer, x2 in [16,32)), B = [b1 | b2] halves of the 128-byte operand row
                                const uint32_t idesc_h = make_idesc_f16(p.n_pad, kPair ? 256 : 128);
#pragma unroll
                                for (int j = 0; j < 2; ++j) {
                                    if (!((km >> (2 * j)) & 3u)) continue;       // 16 taps outside the unit disk
                                    const uint64_t b1d = desc_from_lo(bhl + 2 * j), b2d = desc_from_lo(bhl + 4 + 2 * j);
                                    const uint32_t a1 = a0 + 8 * j, a2 = a0 + 16 + 8 * j;
                                    if constexpr (kPair) {
                                        umma_bf16_ts_2(d0, a1, b1d, idesc_h, acc_run);
                                        umma_bf16_ts_2(d0, a2, b1d, idesc_h, 1u);
                                        umma_bf16_ts_2(d0, a1, b2d, idesc_h, 1u);
                                        if (p.subtiles == 2) {
                                            umma_bf16_ts_2(d0 + p.n_pad, a1 + kBlockK, b1d, idesc_h, acc_run);
                                            umma_bf16_ts_2(d0 + p.n_pad, a2 + kBlockK, b1d, idesc_h, 1u);
                                            umma_bf16_ts_2(d0 + p.n_pad, a1 + kBlockK, b2d, idesc_h, 1u);
                                        }
                                    } else {
                                        umma_bf16_ts(d0, a1, b1d, idesc_h, acc_run);
                                        umma_bf16_ts(d0, a2, b1d, idesc_h, 1u);
                                        umma_bf16_ts(d0, a1, b2d, idesc_h, 1u);
                                        if (p.subtiles == 2) {
                                            umma_bf16_ts(d0 + p.n_pad, a1 + kBlockK, b1d, idesc_h, acc_run);
                                            umma_bf16_ts(d0 + p.n_pad, a2 + kBlockK, b1d, idesc_h, 1u);
                                            umma_bf16_ts(d0 + p.n_pad, a1 + kBlockK, b2d, idesc_h, 1u);
                                        }
                                    }
                                    acc_run = 1u;
                                }
                                if constexpr (kPair) {
                                    umma_commit_2mc(&bempty[sb], 3);
                                    umma_commit_2mc(&lo_empty[lb], 3);
                                    if (kb == kb_end - 1) umma_commit_2mc(&acc_full[buf], 3);
                                } else {
                                    if (p.cluster == 1) umma_commit(&bempty[sb]);
                                    else umma_commit_mc(&bempty[sb], cmask);
                                    umma_commit(&lo_empty[lb]);
                                    if (kb == kb_end - 1) umma_commit(&acc_full[buf]);
                                }
                                if (++sb == p.b_stages) { sb = 0; phb ^= 1; }
                                if (++s == p.n_stages) { s = 0; ph ^= 1; }
                                continue;
                            }
                            // Xhi.Bhi (tf32; the tensor core truncates the raw X itself) and the bf16 correction
                            // Xlo.Bhi + Xhi.Blo, for the 4 K-steps of the k-block and each sub-tile
                            if constexpr (kPair) {
#pragma unroll
                                for (int k4 = 0; k4 < kBlockK / kUmmaK; ++k4) {
                                    if (!((km >> k4) & 1u)) continue;           // 8 taps outside the unit disk: basis all zero
                                    umma_tf32_2(d0, desc_from_lo(xl + 2 * k4), desc_from_lo(bhl + 2 * k4), idesc, acc_run);
                                    umma_bf16_ts_2(d0, a0 + k4 * kUmmaK, desc_from_lo(bcl + 2 * k4), idesc_c, 1u);
                                    if (p.subtiles == 2) {
                                        umma_tf32_2(d0 + p.n_pad, desc_from_lo(xl + 1024 + 2 * k4), desc_from_lo(bhl + 2 * k4), idesc, acc_run);
                                        umma_bf16_ts_2(d0 + p.n_pad, a0 + kBlockK + k4 * kUmmaK, desc_from_lo(bcl + 2 * k4), idesc_c, 1u);
                                    }
                                    acc_run = 1u;
                                }
                                umma_commit_2mc(&empty[s], 3);
                                umma_commit_2mc(&bempty[sb], 3);
                                umma_commit_2mc(&lo_empty[lb], 3);
                                if (++sb == p.b_stages) { sb = 0; phb ^= 1; }
                                if (kb == kb_end - 1) umma_commit_2mc(&acc_full[buf], 3);
                                if (++s == p.n_stages) { s = 0; ph ^= 1; }
                                continue;
                            }
#pragma unroll
                            for (int k4 = 0; k4 < kBlockK / kUmmaK; ++k4) {
                                if (!((km >> k4) & 1u)) continue;               // 8 taps outside the unit disk: basis all zero
                                umma_tf32(d0, desc_from_lo(xl + 2 * k4), desc_from_lo(bhl + 2 * k4), idesc, acc_run);
                                if (!(ZB200_TC3_PROF && (p.dbg & 1))) umma_bf16_ts(d0, a0 + k4 * kUmmaK, desc_from_lo(bcl + 2 * k4), idesc_c, 1u);
                                if (p.subtiles == 2) {
                                    umma_tf32(d0 + p.n_pad, desc_from_lo(xl + 1024 + 2 * k4), desc_from_lo(bhl + 2 * k4), idesc, acc_run);
                                    if (!(ZB200_TC3_PROF && (p.dbg & 1)))
                                        umma_bf16_ts(d0 + p.n_pad, a0 + kBlockK + k4 * kUmmaK, desc_from_lo(bcl + 2 * k4), idesc_c, 1u);
                                }
                                acc_run = 1u;
                            }
                            if (prof) w2 += (unsigned long long)(clock64() - t_issue);
                            const long long t_commit = prof ? clock64() : 0;
                            umma_commit(&empty[s]);
                            if (p.cluster == 1) umma_commit(&bempty[sb]);
                            else umma_commit_mc(&bempty[sb], cmask);
                            umma_commit(&lo_empty[lb]);
                            if (++sb == p.b_stages) { sb = 0; phb ^= 1; }
                            if (kb == kb_end - 1) umma_commit(&acc_full[buf]);
                            if (prof) w3 += (unsigned long long)(clock64() - t_commit);
                            if (++s == p.n_stages) { s = 0; ph ^= 1; }
                        }
                    }
                }
            }
            __syncwarp();
        } else if (warp == kWarpAlloc && p.n_peers) {
            // ===================== K5: forward finished tiles to the peer GPUs =====================
            pusher_loop(p, out_done, (p.epi_solo || kSplit2) ? 4 : 8, my_tiles, lane, smem + p.push_off, p.subtiles * kTileRows);
        }
    } else if (kSplit2 && wg <= 1) {
        // ===================== splitter, one sub-tile per warpgroup =====================
        reg_dec<kRegsSplit2>();
        const int q = warp & 3;
        const int r = q * 32 + lane;                              // row inside this warpgroup's 128-row sub-tile
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const uint32_t x_ring_u32 = smem_u32(smem);
        const uint32_t swz = (uint32_t)(r & 7) << 4;
        const float xsc = p.x_scale;
        int s = 0;
        uint32_t ph = 0, it = 0;
        for (int t = 0; t < my_tiles; ++t) {
            for (int kb = 0; kb < p.k_blocks; ++kb, ++it) {
                const int lb = (int)(it & lo_mask);
                const uint32_t lo_ph = (it >> lo_shift) & 1u;
                mbar_wait(&full[s], ph);
                const uint32_t rowp = x_ring_u32 + (uint32_t)s * x_bytes + (uint32_t)(wg * kTileRows + r) * 128u;
                float4 x[8];
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(x[c].x), "=f"(x[c].y), "=f"(x[c].z), "=f"(x[c].w)
                                 : "r"(rowp + (((uint32_t)c << 4) ^ swz)));
                mbar_wait(&lo_empty[lb], lo_ph ^ 1u);
                tc_fence_after();
                uint32_t lo[32];
                auto pack_h = [](float hi_half, float lo_half) {
                    uint32_t v;
                    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(v) : "f"(hi_half), "f"(lo_half));
                    return v;
                };
                auto pack_b = [](float hi_half, float lo_half) {
                    uint32_t v;
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(v) : "f"(hi_half), "f"(lo_half));
                    return v;
                };
                // (a - trunc(a), b - trunc(b)) as one packed add; trunc = the top 11 significand bits
                auto residual2 = [](float a, float b, uint32_t ta, uint32_t tb, float& ra, float& rb) {
                    uint64_t va, vb, vd;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(va) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b)));
                    asm("mov.b64 %0, {%1, %2};" : "=l"(vb) : "r"(ta ^ 0x80000000u), "r"(tb ^ 0x80000000u));
                    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(vd) : "l"(va), "l"(vb));
                    uint32_t l0, l1;
                    asm("mov.b64 {%0, %1}, %2;" : "=r"(l0), "=r"(l1) : "l"(vd));
                    ra = __uint_as_float(l0);
                    rb = __uint_as_float(l1);
                };
                if constexpr (kF16) {
                    // columns [0,16): x1 pairs, [16,32): x2 pairs (see split_row_h of the one-warpgroup splitter)
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const float a = (h ? x[c].z : x[c].x) * xsc, b = (h ? x[c].w : x[c].y) * xsc;
                            const uint32_t ta = __float_as_uint(a) & 0xFFFFE000u, tb = __float_as_uint(b) & 0xFFFFE000u;
                            float ra, rb;
                            residual2(a, b, ta, tb, ra, rb);
                            lo[2 * c + h] = pack_h(__uint_as_float(tb), __uint_as_float(ta));
                            lo[16 + 2 * c + h] = pack_h(rb, ra);
                        }
                    }
                } else {
                    // per 8 taps: columns 0..3 Xlo pairs, 4..7 Xhi pairs (bf16), like split8
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4& v = x[2 * c8 + (j >> 1)];
                            const float a = (j & 1) ? v.z : v.x, b = (j & 1) ? v.w : v.y;
                            const uint32_t ta = __float_as_uint(a) & 0xFFFFE000u, tb = __float_as_uint(b) & 0xFFFFE000u;
                            float ra, rb;
                            residual2(a, b, ta, tb, ra, rb);
                            lo[8 * c8 + j] = pack_b(rb, ra);
                            lo[8 * c8 + 4 + j] = pack_b(b, a);
                        }
                    }
                }
                tmem_st32(lo_base + lane_addr + (uint32_t)(lb * 2 + wg) * kBlockK, lo);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (kPair && crank != 0) mbar_arrive_cluster(mapa_cluster(smem_u32(&lo_full[lb]), 0));
                    else mbar_arrive(&lo_full[lb]);
                    if constexpr (kF16) mbar_arrive(&empty[s]);
                }
                if (++s == p.n_stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (kSplit2) {
        // ===================== epilogue, one warpgroup for both sub-tiles =====================
        reg_inc<kRegsEpi2>();
        const int q = warp & 3;
        const int n_cc = p.n_pad / 16;                            // <= kCC2
        uint32_t ck = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const int tile = blockIdx.x + t * gridDim.x;
            float sum[2][kCC2][16];
#pragma unroll
            for (int sb2 = 0; sb2 < 2; ++sb2)
#pragma unroll
                for (int cc = 0; cc < kCC2; ++cc)
#pragma unroll
                    for (int i = 0; i < 16; ++i) sum[sb2][cc][i] = 0.f;
            for (int c = 0; c < n_chunks; ++c, ++ck) {
                const int buf = p.acc_bufs == 2 ? (int)(ck & 1) : 0;
                mbar_wait(&acc_full[buf], p.acc_bufs == 2 ? ((ck >> 1) & 1u) : (ck & 1u));
                tc_fence_after();
#pragma unroll
                for (int sb2 = 0; sb2 < 2; ++sb2) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * 2 + sb2) * p.n_pad);
#pragma unroll
                    for (int cc = 0; cc < kCC2; ++cc) {
                        if (cc < n_cc) {
                            uint32_t v[16];
                            tmem_ld16(taddr + cc * 16, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) sum[sb2][cc][i] += __uint_as_float(v[i]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (kPair && crank != 0) mbar_arrive_cluster(mapa_cluster(smem_u32(&acc_empty[buf]), 0));
                    else mbar_arrive(&acc_empty[buf]);
                }
            }
#pragma unroll
            for (int sb2 = 0; sb2 < 2; ++sb2) {
                const long long row = (long long)tile * 2 * kTileRows + sb2 * kTileRows + q * 32 + lane;
                if (row < p.n_patches) {
                    ScoreAcc sc;
                    if constexpr (kOut == kOutScores) sc.clear();
#pragma unroll
                    for (int cc = 0; cc < kCC2; ++cc) {
                        if (cc < n_cc) {
                            uint32_t v[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                v[i] = __float_as_uint(kF16 ? sum[sb2][cc][i] * p.out_scale : sum[sb2][cc][i]);
                            epilogue_chunk<kOut>(p, row, cc * 16, v, sc);
                        }
                    }
                    if constexpr (kOut == kOutScores) finish_scores(p, row, sc);
                }
            }
            if (p.n_peers) {
                __threadfence();
                asm volatile("fence.proxy.async.global;" ::: "memory");
                __syncwarp();
                if (lane == 0) atomicAdd(out_done, 1u);
            }
        }
    } else if (wg == 0) {
        // ===================== splitter =====================
        // Xlo = x - trunc_tf32(x): exact in fp32, and exactly what the tensor core drops when it
        // truncates the raw X operand.  Both sub-tiles' rows are loaded before the buffer wait so
        // the shared-memory latency overlaps the previous k-block's MMAs.
        reg_dec<kRegsSplit>();
        const int q = warp & 3;
        const int r = q * 32 + lane;                              // row inside a 128-row sub-tile
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        int s = 0;
        uint32_t ph = 0;
        uint32_t it = 0;
        // Eight k values (two 16-B chunks) -> 8 TMEM columns of bf16 pairs: columns 0..3 hold
        // Xlo[0..7], columns 4..7 hold Xhi[0..7] (K slot 2c in the low half-word, 2c+1 in the high).
        // The four splitter warps are this kernel's critical path (ncu source page: ~100 % of their samples sit in
        // the per-k-block loop, 365 SASS instructions), so the split is written for instruction count: the pair
        // (x0 - trunc(x0), x1 - trunc(x1)) is ONE packed add.f32x2 on (x, -(x & mask)), each negated truncation one
        // LOP3 ((x & mask) ^ sign), and the tile rows are read with ld.shared.v4 from 32-bit shared addresses.
        auto pack2 = [](float even, float odd) {
            uint32_t v;
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(v) : "f"(odd), "f"(even));   // .x (low half-word) = even
            return v;
        };
        auto lo_pair = [&](float x0, float x1) {
            const uint32_t n0 = (__float_as_uint(x0) & 0xFFFFE000u) ^ 0x80000000u;   // -trunc_tf32(x0)
            const uint32_t n1 = (__float_as_uint(x1) & 0xFFFFE000u) ^ 0x80000000u;
            uint64_t a, b, d;
            asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(__float_as_uint(x0)), "r"(__float_as_uint(x1)));
            asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "r"(n0), "r"(n1));
            asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
            uint32_t l0, l1;
            asm("mov.b64 {%0, %1}, %2;" : "=r"(l0), "=r"(l1) : "l"(d));
            return pack2(__uint_as_float(l0), __uint_as_float(l1));
        };
        auto split8 = [&](const float4& a, const float4& b, uint32_t* w) {
            w[0] = lo_pair(a.x, a.y);
            w[1] = lo_pair(a.z, a.w);
            w[2] = lo_pair(b.x, b.y);
            w[3] = lo_pair(b.z, b.w);
            w[4] = pack2(a.x, a.y);
            w[5] = pack2(a.z, a.w);
            w[6] = pack2(b.x, b.y);
            w[7] = pack2(b.z, b.w);
        };
        // fp16 split of a pair: x1 = the top 11 significand bits of x' = x * scale (exact in fp16), x2 = RN_f16(x' - x1)
        const float xsc = p.x_scale;
        auto split2h = [&](float a, float b, uint32_t& w1, uint32_t& w2) {
            a *= xsc;
            b *= xsc;
            const uint32_t ta = __float_as_uint(a) & 0xFFFFE000u, tb = __float_as_uint(b) & 0xFFFFE000u;
            asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w1) : "f"(__uint_as_float(tb)), "f"(__uint_as_float(ta)));   // low half = a
            uint64_t va, vb, vd;
            asm("mov.b64 %0, {%1, %2};" : "=l"(va) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b)));
            asm("mov.b64 %0, {%1, %2};" : "=l"(vb) : "r"(ta ^ 0x80000000u), "r"(tb ^ 0x80000000u));
            asm("add.rn.f32x2 %0, %1, %2;" : "=l"(vd) : "l"(va), "l"(vb));
            uint32_t l0, l1;
            asm("mov.b64 {%0, %1}, %2;" : "=r"(l0), "=r"(l1) : "l"(vd));
            asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w2) : "f"(__uint_as_float(l1)), "f"(__uint_as_float(l0)));
        };
        // one row of a k-block (8 chunks of 4 taps) -> 32 TMEM columns: x1 pairs in [0,16), x2 pairs in [16,32)
        auto split_row_h = [&](const float4 (&x)[8], uint32_t* w) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                split2h(x[c].x, x[c].y, w[2 * c], w[16 + 2 * c]);
                split2h(x[c].z, x[c].w, w[2 * c + 1], w[16 + 2 * c + 1]);
            }
        };
        auto lds128 = [](uint32_t addr) {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
            return v;
        };
        const uint32_t x_ring_u32 = smem_u32(smem);
        const uint32_t swz = (uint32_t)(r & 7) << 4;                  // logical 16-B chunk c sits at (c << 4) ^ swz
        const int g_rows = 32 * p.subtiles;                       // tile rows this warp gathers: [q*g_rows, +g_rows)
        for (int t = 0; t < my_tiles; ++t) {
            int2 ca = make_int2(-(1 << 28), -(1 << 28)), cb = ca;  // windows of the rows q*g_rows+lane (and +32)
            if (p.gather) {
                const long long base = ((long long)blockIdx.x + (long long)t * gridDim.x) * p.subtiles * kTileRows + q * g_rows;
                if (base + lane < p.n_patches) ca = __ldg(p.g_xy + base + lane);
                if (p.subtiles == 2 && base + 32 + lane < p.n_patches) cb = __ldg(p.g_xy + base + 32 + lane);
            }
            auto gather_issue = [&](int kbi, int sg, uint32_t phg) {
                mbar_wait(&empty[sg], phg ^ 1u);                   // the stage's previous MMAs have retired
                const int e0 = (p.kb0 + kbi) * kBlockK;
                const uint32_t xs = smem_u32(stage_x(sg)) + (uint32_t)(q * g_rows) * 128u;
                if (p.g_planes) {
                    // 16-byte copies from the shifted planes: 8 lanes move one 32-float segment, a warp instruction
                    // moves the segments of 4 tile rows (4x fewer LSU operations than the 4-byte path)
                    const int c = lane & 7;                          // 16-B chunk of the segment = 4 taps
                    const int e = e0 + 4 * c;
                    const int wr = e / p.g_k, wc = e - wr * p.g_k;   // g_k % 4 == 0: a chunk never straddles window rows
                    for (int j0 = 0; j0 < g_rows; j0 += 4) {
                        const int j = j0 + (lane >> 3);
                        const int2 src = (j0 & 32) ? cb : ca;
                        const int x0 = __shfl_sync(0xffffffffu, src.x, j & 31);
                        const int y0 = __shfl_sync(0xffffffffu, src.y, j & 31);
                        const int yy = y0 + wr;
                        const int xl = x0 + wc + p.g_L;              // column in the padded frame
                        const int rr = xl & 3, u = xl - rr;
                        const bool inb = yy >= 0 && yy < p.g_H && xl >= 0 && u + 4 <= p.g_Wp;
                        const float* gp = p.g_planes + (inb ? ((size_t)rr * p.g_H + yy) * p.g_Wp + u : 0);
                        const int row = q * g_rows + j;              // tile row (swizzle phase = row % 8)
                        const uint32_t dst = xs + (uint32_t)j * 128u + (((uint32_t)c ^ (uint32_t)(row & 7)) << 4);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gp), "r"(inb ? 16 : 0)
                                     : "memory");                    // src-size 0 -> zero fill outside the frame
                    }
                } else {
                    int wr = e0 / p.g_k, wc = e0 - wr * p.g_k + lane;   // window row / column of this lane
                    if (wc >= p.g_k) { wc -= p.g_k; ++wr; }
                    for (int j = 0; j < g_rows; ++j) {
                        const int2 src = (j & 32) ? cb : ca;
                        const int x0 = __shfl_sync(0xffffffffu, src.x, j & 31);
                        const int y0 = __shfl_sync(0xffffffffu, src.y, j & 31);
                        const int yy = y0 + wr, xx = x0 + wc;
                        const bool inb = yy >= 0 && yy < p.g_H && xx >= 0 && xx < p.g_W;
                        const float* gp = p.g_img + (inb ? (size_t)yy * p.g_W + xx : 0);
                        const int row = q * g_rows + j;                 // tile row (swizzle phase = row % 8)
                        const uint32_t dst = xs + (uint32_t)j * 128u + ((((uint32_t)lane >> 2) ^ (uint32_t)(row & 7)) << 4) +
                                             ((uint32_t)lane & 3u) * 4u;
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(gp), "r"(inb ? 4 : 0)
                                     : "memory");                       // src-size 0 -> zero fill outside the frame
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
            for (int kb = 0; kb < p.k_blocks; ++kb, ++it) {
                const int lb = (int)(it & lo_mask);
                const uint32_t lo_ph = (it >> lo_shift) & 1u;
                if (!p.gather) {
                    mbar_wait_t(&full[s], ph, w0, prof);
                } else {
                    // K2 fused: the four warps copy 32-float window segments (one per tile row, coalesced
                    // 128-B reads of the L2-resident frame) into the 128-B-swizzled K-major X stage with
                    // cp.async (16-byte copies from the shifted planes, else 4-byte; no register staging, every
                    // copy of a k-block in flight at once), one k-block ahead of the split so the L2 latency
                    // hides behind it (two ahead measured no faster: the splitter warps' instruction stream,
                    // not the latency, bounds this path).
                    if (kb == 0) gather_issue(0, s, ph);
                    if (kb + 1 < p.k_blocks) {
                        const int sn = (s + 1 == p.n_stages) ? 0 : s + 1;
                        gather_issue(kb + 1, sn, sn == 0 ? ph ^ 1u : ph);
                        asm volatile("cp.async.wait_group 1;" ::: "memory");
                    } else {
                        asm volatile("cp.async.wait_group 0;" ::: "memory");
                    }
                    fence_proxy_async();                           // the MMA (async proxy) reads these rows
                    asm volatile("bar.sync 1, 128;" ::: "memory");  // all four gather warps done with this stage
                }
                const uint32_t rowp = x_ring_u32 + (uint32_t)s * x_bytes + (uint32_t)r * 128u;
                float4 x0[8], x1[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) x0[c] = lds128(rowp + (((uint32_t)c << 4) ^ swz));
                if (p.subtiles == 2) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) x1[c] = lds128(rowp + kTileRows * 128u + (((uint32_t)c << 4) ^ swz));
                }
                mbar_wait_t(&lo_empty[lb], lo_ph ^ 1u, w1, prof);
                tc_fence_after();
                if (!(ZB200_TC3_PROF && (p.dbg & 2))) {
                    uint32_t lo[32];
                    if constexpr (kF16) split_row_h(x0, lo);
                    else {
#pragma unroll
                        for (int c = 0; c < 4; ++c) split8(x0[2 * c], x0[2 * c + 1], lo + 8 * c);
                    }
                    tmem_st32(lo_base + lane_addr + (uint32_t)(lb * p.subtiles + 0) * kBlockK, lo);
                }
                if (p.subtiles == 2 && !(ZB200_TC3_PROF && (p.dbg & 2))) {
                    uint32_t lo[32];
                    if constexpr (kF16) split_row_h(x1, lo);
                    else {
#pragma unroll
                        for (int c = 0; c < 4; ++c) split8(x1[2 * c], x1[2 * c + 1], lo + 8 * c);
                    }
                    tmem_st32(lo_base + lane_addr + (uint32_t)(lb * p.subtiles + 1) * kBlockK, lo);
                }
                const long long t_st = prof ? clock64() : 0;
                tmem_st_wait();
                if (prof) w2 += (unsigned long long)(clock64() - t_st);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (kPair && crank != 0) mbar_arrive_cluster(mapa_cluster(smem_u32(&lo_full[lb]), 0));
                    else mbar_arrive(&lo_full[lb]);
                    if constexpr (kF16) mbar_arrive(&empty[s]);      // the row lives in TMEM now: no MMA reads this X stage
                }
                if (++s == p.n_stages) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warpgroups 1 and 2) =====================
        reg_inc<kRegsEpi>();
        const int g = wg - 1;
        const int q = warp & 3;
        // two accumulators: group g owns sub-tile g, all columns; one accumulator: group g owns a column half
        const int sub = (p.subtiles == 2) ? g : 0;
        const int half = ((p.n_pad / 16 + 1) / 2) * 16;
        const int c_beg = (p.subtiles == 2 || p.epi_solo) ? 0 : g * half;
        const int c_end = (p.subtiles == 2 || p.epi_solo) ? p.n_pad : min(p.n_pad, (g + 1) * half);
        const bool idle = p.epi_solo && g == 1;                   // second warpgroup has nothing to own
        const int n_cc = (c_end - c_beg + 15) / 16;               // <= kMaxColChunks
        uint32_t ck = 0;
        for (int t = 0; t < (idle ? 0 : my_tiles); ++t) {
            const int tile = blockIdx.x + t * gridDim.x;
            const long long row = (long long)tile * p.subtiles * kTileRows + sub * kTileRows + q * 32 + lane;
            float sum[kMaxColChunks][16];
#pragma unroll
            for (int cc = 0; cc < kMaxColChunks; ++cc)
#pragma unroll
                for (int i = 0; i < 16; ++i) sum[cc][i] = 0.f;
            for (int c = 0; c < n_chunks; ++c, ++ck) {
                const int buf = p.acc_bufs == 2 ? (int)(ck & 1) : 0;
                mbar_wait_t(&acc_full[buf], p.acc_bufs == 2 ? ((ck >> 1) & 1u) : (ck & 1u), w0, prof);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) +
                                       (uint32_t)((buf * p.subtiles + sub) * p.n_pad + c_beg);
#pragma unroll
                for (int cc = 0; cc < kMaxColChunks; ++cc) {
                    if (cc < n_cc) {
                        uint32_t v[16];
                        tmem_ld16(taddr + cc * 16, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) sum[cc][i] += __uint_as_float(v[i]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (kPair && crank != 0) mbar_arrive_cluster(mapa_cluster(smem_u32(&acc_empty[buf]), 0));
                    else mbar_arrive(&acc_empty[buf]);
                }
            }
            if (row < p.n_patches) {
                ScoreAcc sc;
                if constexpr (kOut == kOutScores) sc.clear();
#pragma unroll
                for (int cc = 0; cc < kMaxColChunks; ++cc) {
                    if (cc < n_cc) {
                        uint32_t v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(kF16 ? sum[cc][i] * p.out_scale : sum[cc][i]);
                        epilogue_chunk<kOut>(p, row, c_beg + cc * 16, v, sc);
                    }
                }
                if constexpr (kOut == kOutScores) finish_scores(p, row, sc);   // host guarantees subtiles == 2
            }
            if (p.n_peers) {
                __threadfence();                              // this tile's rows are in L2 before the pusher is told
                asm volatile("fence.proxy.async.global;" ::: "memory");
                __syncwarp();
                if (lane == 0) atomicAdd(out_done, 1u);
            }
        }
    }

#if ZB200_DEBUG_HOOKS
    if (prof && lane == 0 && (warp == kWarpTma || warp == kWarpMma || warp == 4 || warp == 0)) {
        const int base = warp == kWarpTma ? 0 : (warp == kWarpMma ? 1 : (warp == 4 ? 5 : 8));
        atomicAdd(&g_wait_cycles[base], w0);        // 0 producer.empty | 1 mma.acc_empty | 5 epi.acc_full | 8 split.full
        atomicAdd(&g_wait_cycles[base + 1], w1);    // 2 mma.lo_full | 9 split.lo_empty
        atomicAdd(&g_wait_cycles[base + 2], w2);    // 3 mma.issue   | 10 split.st_wait
        if (warp == kWarpMma) atomicAdd(&g_wait_cycles[4], w3);   // 4 mma.commit
    }
#endif
    tc_fence_before();
    __syncthreads();
    if (p.cluster > 1) cluster_sync_all();        // no CTA leaves while a peer may still write into it
    if (warp == kWarpAlloc) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// ---- host side -----------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    }
    return fn;
}

// 2-D fp32 tensor [rows][cols] (row pitch pitch_bytes), box = 32 floats x box_rows, 128B swizzle
static int encode_2d(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t pitch_bytes,
                     uint32_t box_rows) {
    auto enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return ZB200_ECUDA;
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch_bytes};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (cols=%llu rows=%llu pitch=%llu box_rows=%u)", (int)r,
                  (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)pitch_bytes, box_rows);
        return ZB200_ECUDA;
    }
    return ZB200_OK;
}

}  // namespace tc

// Both kernels need one UMMA N (<= 256 operand rows).  tf32x3 keeps two accumulator sets plus the operand staging
// in the 512 TMEM columns while 2*n_pad + 32 <= 512; wider operands (n_max = 20: 231 real rows -> 240, 242
// complex-interleaved rows -> 256) run with ONE set and four staging buffers (256 + 4*32 columns).
static bool operand_ok(const Operand& op, int precision) {
    (void)precision;
    return op.has_tmap && op.rows_pad <= 256;
}

bool tc_supported(const zb200_plan* p, int precision, bool complex_order) {
    // sm_100 family and 16-byte aligned patch rows (k*k % 4 == 0) for TMA
    if (p->cc_major != 10 || p->kk % 4 != 0) return false;
    return operand_ok(complex_order ? p->cplx : p->real, precision);
}

int init_tensor_maps(zb200_plan* p) {
    if (p->cc_major != 10 || p->kk % 4 != 0) return ZB200_OK;       // SIMT only; not an error
    for (Operand* op : {&p->real, &p->cplx}) {
        if (op->rows_pad > 256) continue;
        op->max_cluster = 1;
        for (int lg = 0; lg < 3; ++lg) {
            const int c = 1 << lg;
            if (op->rows_pad % (8 * c) != 0) break;                  // B slices must start on a swizzle atom
            int rc = tc::encode_2d(&op->tmap_hi[lg], op->hi, (uint64_t)p->k_pad, (uint64_t)op->rows_pad,
                                   (uint64_t)p->k_pad * 4, (uint32_t)(op->rows_pad / c));
            if (rc) return rc;
            rc = tc::encode_2d(&op->tmap_cb[lg], op->cb, (uint64_t)p->k_pad, (uint64_t)op->rows_pad,
                               (uint64_t)p->k_pad * 4, (uint32_t)(op->rows_pad / c));
            if (rc) return rc;
            rc = tc::encode_2d(&op->tmap_lo_f32[lg], op->lo, (uint64_t)p->k_pad, (uint64_t)op->rows_pad,
                               (uint64_t)p->k_pad * 4, (uint32_t)(op->rows_pad / c));
            if (rc) return rc;
            rc = tc::encode_2d(&op->tmap_hb[lg], op->hb, (uint64_t)p->k_pad, (uint64_t)op->rows_pad,
                               (uint64_t)p->k_pad * 4, (uint32_t)(op->rows_pad / c));
            if (rc) return rc;
            op->max_cluster = c;
        }
        op->has_tmap = true;
    }
    return ZB200_OK;
}

int project_tc(const zb200_plan* p, const float* d_patches, int64_t n, int precision, int out_kind, void* d_out,
               void* d_out2, const float* d_w, const uint8_t* d_sel, int n_folds, int norm_kind, cudaStream_t s,
               const GatherSource* gsrc, const PeerTargets* peers, double value_max, const unsigned* run_if) {
    using namespace tc;
    if (n == 0) return ZB200_OK;
    ZB_CHECK_ARG(gsrc || (reinterpret_cast<uintptr_t>(d_patches) & 15) == 0,
                 "project: patch pointer must be 16-byte aligned");
    const bool h3 = precision == ZB200_PREC_F16X3;
    const bool folded = h3 && !gsrc && !d_w && fold_supported(p) && knobs().tc_fold != 0;
    if (h3 && !folded && !(value_max > 0.0)) {
        set_error("the f16x3 projection needs an upper bound of |patch values| (value_max > 0) to scale them into fp16 range");
        return ZB200_EINVAL;
    }
    if (gsrc && (precision != ZB200_PREC_TF32X3 || p->size < 32)) {
        set_error("fused gather+projection needs precision tf32x3 and a window of at least 32 pixels");
        return ZB200_EUNSUP;
    }
    const bool x3 = precision == ZB200_PREC_TF32X3 || h3;       // the fp32-grade kernel family
    const bool scores = d_w != nullptr;
    // windows of 64 / 128 pixels: the mirror-folded kernel (a quarter of the multiply-adds, zb200_project_fold.cu)
    if (folded) {
        if (value_max > 0.0) return project_fold(p, d_patches, n, out_kind, d_out, d_out2, s, peers, value_max);
        // no bound from the caller: scale by the largest |x| of a sample of the stack; should an unsampled value be
        // more than 8x larger (fp16 overflow, flagged by the kernel), the range-free tf32x3 kernel behind it runs --
        // otherwise its CTAs return at once
        uint32_t* aux = nullptr;
        ZB_CUDA(scratch_alloc(&aux, 2 * sizeof(uint32_t), s));
        ZB_CUDA(cudaMemsetAsync(aux, 0, 2 * sizeof(uint32_t), s));
        int rc = project_fold(p, d_patches, n, out_kind, d_out, d_out2, s, peers, 0.0, aux);
        if (!rc && tc_supported(p, ZB200_PREC_TF32X3, out_kind != ZB200_OUT_REAL))
            rc = project_tc(p, d_patches, n, ZB200_PREC_TF32X3, out_kind, d_out, d_out2, nullptr, nullptr, 0, 0, s, nullptr, peers,
                            0.0, aux + 1);
        cudaFreeAsync(aux, s);
        return rc;
    }
    const bool cplx = !scores && out_kind != ZB200_OUT_REAL;
    const Operand& op = cplx ? p->cplx : p->real;
    if (!tc_supported(p, precision, cplx)) {
        set_error("tcgen05 projection (precision %d, %s order) unsupported for n_max=%d size=%d: needs sm_100, even "
                  "size and <= %d operand rows (have %d)", precision, cplx ? "complex" : "real", p->n_max, p->size,
                  256, op.rows_pad);
        return ZB200_EUNSUP;
    }

    Params prm{};
    prm.n_patches = n;
    prm.n_pad = op.rows_pad;
    prm.n_cols = op.rows;
    prm.kmask = p->d_kmask;
    const Knobs& kn = knobs();
    if (kn.tc_kskip == 0) prm.kmask = p->d_kmask + p->k_pad / 32;      // experiment: every K step issued
    prm.kb0 = p->kb_first;                                  // leading / trailing k-blocks with an all-zero basis
    prm.k_blocks = p->kb_last - p->kb_first;                //   (window rows outside the unit disk) are never loaded
    prm.out = static_cast<float*>(d_out);
    prm.out2 = static_cast<float*>(d_out2);
    prm.w = d_w;
    prm.sel = d_sel;
    prm.n_folds = n_folds;
    prm.norm_kind = norm_kind;
    prm.out_kind = scores ? 100 : out_kind;
    prm.row_len = scores ? n_folds
                         : (out_kind == ZB200_OUT_REAL ? p->n_modes
                                                       : (out_kind == ZB200_OUT_COMPLEX ? 2 * p->n_complex : p->n_complex));
    prm.chunk_kb = 8;
    if (kn.tc_chunk) prm.chunk_kb = kn.tc_chunk;
    prm.x_scale = prm.out_scale = 1.f;
    if (h3) {
        int e = 0;
        frexp(value_max, &e);                               // value_max = f * 2^e, f in [0.5, 1): |x| * 2^(14-e) <= 2^14
        const int sh = 14 - e < -100 ? -100 : (14 - e > 100 ? 100 : 14 - e);
        prm.x_scale = (float)ldexp(1.0, sh);
        prm.out_scale = (float)(p->inv_area * ldexp(1.0, -sh));
    }
    if (gsrc) {
        prm.gather = 1;
        prm.g_img = gsrc->img;
        prm.g_planes = gsrc->planes;
        prm.g_Wp = gsrc->Wp;
        prm.g_L = gsrc->L;
        prm.g_H = gsrc->H;
        prm.g_W = gsrc->W;
        prm.g_k = p->size;
        prm.g_xy = gsrc->xy0;
    }
    if (peers && peers->n > 0) {
        if (peers->n > kMaxPeers || out_kind == ZB200_OUT_ABS_PHASE) {
            set_error("project: peer push supports at most %d peers and one output array", kMaxPeers);
            return ZB200_EUNSUP;
        }
        prm.n_peers = peers->n;
        for (int g = 0; g < peers->n; ++g) prm.peer_out[g] = peers->out[g];
    }
    prm.run_if = run_if;
    prm.dbg = kn.tc_debug;                                  // 0 unless a debug-hook build runs with ZB200_EXPERIMENT=1
    if (prm.dbg & 4) prm.chunk_kb = prm.k_blocks;

    // tile shape: two 128-patch accumulators per tile when TMEM/smem allow and there is enough work
    // to keep every SM busy with 256-patch tiles
    const bool pushing = peers && peers->n > 0;
    // barriers (+ alignment slack); with peers also the pusher's staging region, 128-byte aligned behind them
    const int bar_core = 1024 + 8 * (2 * 8 + 24) + 16;
    const int bar_bytes = pushing ? round_up(bar_core, 128) + 128 + (int)kPushRegion : bar_core;
    auto stage_bytes = [&](int sub) { return sub * kTileRows * 128 + (x3 ? 2 : 1) * prm.n_pad * 128; };
    int sub = 2;
    if (scores && n_folds > kFusedFolds) {
        set_error("fused n-fold scores support at most %d folds (got %d)", kFusedFolds, n_folds);
        return ZB200_EUNSUP;
    }
    if (x3) {
        sub = prm.n_pad <= 96 ? 2 : 1;                    // 2 accumulator sets x sub x n_pad + staging <= 512 columns
        if (ceil_div(n, 256) < p->sm_count || (prm.dbg & 8)) sub = 1;
        prm.epi_solo = (sub == 1 && prm.n_pad <= 16 * kMaxColChunks) ? 1 : 0;
        if (scores && sub == 1 && !prm.epi_solo) {
            set_error("fused n-fold scores in tf32x3 need <= 128 operand rows (have %d); use the unfused path", prm.n_pad);
            return ZB200_EUNSUP;
        }
    } else {
        if (sub * prm.n_pad > (int)kTmemCols) sub = 1;
        if (sub == 2 && ceil_div(n, 256) < p->sm_count) sub = 1;
        if (prm.dbg & 8) sub = 1;
        if ((kSmemLimit - bar_bytes) / stage_bytes(sub) < 2 && sub == 2) sub = 1;
    }
    prm.subtiles = sub;
    prm.n_tiles = (int)ceil_div(n, (int64_t)sub * kTileRows);
    // cluster of C CTAs shares each B k-block through TMA multicast (L2 -> SM traffic of B / C)
    int cluster = 2;
    if (kn.tc_cluster) cluster = kn.tc_cluster;
    while (cluster > 1 && (cluster > op.max_cluster || prm.n_tiles < 2 * cluster)) cluster >>= 1;
    prm.cluster = cluster;
    const int lg = cluster == 4 ? 2 : (cluster == 2 ? 1 : 0);
    // CTA pairs (cta_group::2) for the fp32-grade kernel: needs the 2-CTA cluster and the TMA-fed X ring
    prm.pair = (x3 && cluster == 2 && !gsrc && op.rows_pad % 16 == 0) ? 1 : 0;
    if (kn.tc_pair == 0) prm.pair = 0;

    prm.b_stages = 3;
    if (kn.tc_bstages) prm.b_stages = kn.tc_bstages;
    if (x3) {
        const int bst = (prm.pair ? 1 : 2) * prm.n_pad * 128 / (h3 ? 2 : 1), xb = sub * kTileRows * 128;
        prm.n_stages = (kSmemLimit - bar_bytes - prm.b_stages * bst) / xb;
    } else
    prm.n_stages = (kSmemLimit - bar_bytes) / stage_bytes(sub);
    if (prm.n_stages > 8) prm.n_stages = 8;
    if (kn.tc_stages && kn.tc_stages < prm.n_stages) prm.n_stages = kn.tc_stages;
    if (prm.n_stages < 1) {
        set_error("project_tc: operand of %d rows does not fit shared memory", prm.n_pad);
        return ZB200_EUNSUP;
    }
    prm.acc_bufs = (2 * sub * prm.n_pad <= (int)kTmemCols) ? 2 : 1;
    {
        const int per_buf = sub * kBlockK;
        int room = ((int)kTmemCols - prm.acc_bufs * sub * prm.n_pad) / per_buf;
        if (x3 && prm.acc_bufs == 2 && room < 2) {
            // wide operands (n_max = 20: 240 rows): two accumulator sets would leave ONE staging buffer and serialise
            // the splitter with the MMAs on every k-block; one set + four staging buffers only stalls at the drains
            prm.acc_bufs = 1;
            room = ((int)kTmemCols - sub * prm.n_pad) / per_buf;
        }
        if (kn.tc_accbufs) {
            const int v = kn.tc_accbufs;
            if (x3 && (v == 1 || (v == 2 && 2 * sub * prm.n_pad + per_buf <= (int)kTmemCols))) {
                prm.acc_bufs = v;
                room = ((int)kTmemCols - v * sub * prm.n_pad) / per_buf;
            }
        }
        prm.lo_bufs = room >= 4 ? 4 : (room >= 2 ? 2 : 1);
    }
    prm.n_tiles = (int)ceil_div(n, (int64_t)sub * kTileRows);

    // two splitter warpgroups + one epilogue warpgroup: two 128-patch sub-tiles, <= 96 operand rows, TMA-fed X, CTA pairs
    bool split2 = x3 && prm.pair && sub == 2 && prm.n_pad <= 16 * kCC2 && !gsrc;
    if (kn.tc_split2 == 0) split2 = false;
    CUtensorMap map_x;
    int rc = gsrc ? encode_2d(&map_x, op.hi, (uint64_t)p->k_pad, (uint64_t)op.rows_pad, (uint64_t)p->k_pad * 4, 8)   // unused
                  : encode_2d(&map_x, d_patches, (uint64_t)p->kk, (uint64_t)n, (uint64_t)p->kk * 4, (uint32_t)(sub * kTileRows));
    if (rc) return rc;

    const size_t ring_bytes = x3 ? (size_t)prm.n_stages * sub * kTileRows * 128 + (size_t)prm.b_stages * (prm.pair ? 1 : 2) * prm.n_pad * 128 / (h3 ? 2 : 1)
                                 : (size_t)prm.n_stages * stage_bytes(sub);
    const size_t smem = ring_bytes + bar_bytes;
    // rings start at the 1024-byte aligned base (<= 1023 bytes of slack inside bar_core), barriers follow the rings
    prm.push_off = (uint32_t)round_up((int)ring_bytes + (bar_core - 1024), 128);
    int grid = prm.n_tiles < p->sm_count ? prm.n_tiles : p->sm_count;
    grid = (grid / cluster) * cluster;                               // whole clusters only (148 = 2*74 = 4*37)
    if (grid < cluster) grid = cluster;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int kout = scores ? kOutScores
                            : (out_kind == ZB200_OUT_ABS ? kOutAbs : (out_kind == ZB200_OUT_ABS_PHASE ? kOutAbsPhase : kOutPlain));
#define ZB_TC_LAUNCH(KOUT)                                                                                            \
    if (kout == KOUT) {                                                                                               \
        if (split2 && h3) {                                                                                           \
            ZB_CUDA(cudaFuncSetAttribute(project_tc3_kernel<KOUT, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem));                                                                 \
            cfg.blockDim = dim3(512);                                                                                 \
            ZB_CUDA(cudaLaunchKernelEx(&cfg, project_tc3_kernel<KOUT, true, true, true>, map_x, op.tmap_hb[lg], op.tmap_hb[lg], prm)); \
        } else if (split2) {                                                                                          \
            ZB_CUDA(cudaFuncSetAttribute(project_tc3_kernel<KOUT, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem));                                                                 \
            cfg.blockDim = dim3(512);                                                                                 \
            ZB_CUDA(cudaLaunchKernelEx(&cfg, project_tc3_kernel<KOUT, true, false, true>, map_x, op.tmap_hi[lg], op.tmap_cb[lg], prm)); \
        } else if (h3 && prm.pair) {                                                                                  \
            ZB_CUDA(cudaFuncSetAttribute(project_tc3_kernel<KOUT, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem));                                                                 \
            cfg.blockDim = dim3(512);                                                                                 \
            ZB_CUDA(cudaLaunchKernelEx(&cfg, project_tc3_kernel<KOUT, true, true, false>, map_x, op.tmap_hb[lg], op.tmap_hb[lg], prm)); \
        } else if (h3) {                                                                                              \
            ZB_CUDA(cudaFuncSetAttribute(project_tc3_kernel<KOUT, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem));                                                                 \
            cfg.blockDim = dim3(512);                                                                                 \
            ZB_CUDA(cudaLaunchKernelEx(&cfg, project_tc3_kernel<KOUT, false, true, false>, map_x, op.tmap_hb[lg], op.tmap_hb[lg], prm)); \
        } else if (x3 && prm.pair) {                                                                                  \
            ZB_CUDA(cudaFuncSetAttribute(project_tc3_kernel<KOUT, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem));                                                                 \
            cfg.blockDim = dim3(512);                                                                                 \
            ZB_CUDA(cudaLaunchKernelEx(&cfg, project_tc3_kernel<KOUT, true, false, false>, map_x, op.tmap_hi[lg], op.tmap_cb[lg], prm)); \
        } else if (x3) {                                                                                              \
            ZB_CUDA(cudaFuncSetAttribute(project_tc3_kernel<KOUT, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem));                                                                 \
            cfg.blockDim = dim3(512);                                                                                 \
            ZB_CUDA(cudaLaunchKernelEx(&cfg, project_tc3_kernel<KOUT, false, false, false>, map_x, op.tmap_hi[lg], op.tmap_cb[lg], prm)); \
        } else {                                                                                                      \
            ZB_CUDA(cudaFuncSetAttribute(project_tc_kernel<KOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                         (int)smem));                                                                 \
            cfg.blockDim = dim3(256);                                                                                 \
            ZB_CUDA(cudaLaunchKernelEx(&cfg, project_tc_kernel<KOUT>, map_x, op.tmap_hi[lg], prm));                   \
        }                                                                                                             \
    }
    ZB_TC_LAUNCH(kOutPlain)
    ZB_TC_LAUNCH(kOutAbs)
    ZB_TC_LAUNCH(kOutAbsPhase)
    ZB_TC_LAUNCH(kOutScores)
#undef ZB_TC_LAUNCH
    ZB_LAUNCHED();
#if ZB200_DEBUG_HOOKS
    if (prm.dbg & 16) {
        unsigned long long h[16];
        cudaDeviceSynchronize();
        cudaMemcpyFromSymbol(h, g_wait_cycles, sizeof(h));
        fprintf(stderr, "[zb200 wait cycles / CTA] producer.empty=%llu mma.acc_empty=%llu mma.lo_full=%llu mma.issue=%llu "
                "mma.commit=%llu split.full=%llu split.lo_empty=%llu split.st_wait=%llu epi.acc_full=%llu (grid %d, tiles %d, k_blocks %d, stages %d)\n",
                h[0] / grid, h[1] / grid, h[2] / grid, h[3] / grid, h[4] / grid, h[8] / grid, h[9] / grid, h[10] / grid, h[5] / grid, grid, prm.n_tiles, prm.k_blocks,
                prm.n_stages);
        unsigned long long z[16] = {0};
        cudaMemcpyToSymbol(g_wait_cycles, z, sizeof(z));
    }
#endif
    return ZB200_OK;
}

}  // namespace zb200
