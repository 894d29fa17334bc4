// K3, mirror-folded -- the fp32-grade patch projection for windows whose side is a multiple of 64 pixels
// (the metric shape and BASELINE configs 3 and 5).  Same contract as project_tc3_kernel<.., kF16>
// (ZPs._transform_dot_product, mtflearn/features/_zps.py:146-157, + to_complex / abs / angle,
// _zmoments.py:300-316), a quarter of the multiply-adds.
//
// Every Zernike plane has a definite parity under the two mirror operations of the pixel grid (x = column,
// y = row, theta = arctan2(y, x), _zps.py:68-75):
//     cos(m theta): even in y, (-1)^m     in x          sin(m theta): odd in y, (-1)^(m+1) in x
// so the modes fall in four classes and, with i' = k-1-i, j' = k-1-j,
//     sum_{i,j} x[i][j] V[i][j]  =  sum_{i<k/2, j<k/2} F_c[i][j] V[i][j]
//     F_c = (x[i][j] + sj x[i][j']) + si (x[i'][j] + sj x[i'][j'])          (sj, si) = the class's parities
// One butterfly (8 adds per 4 pixels) turns the four mirror images of a quadrant pixel into the four class inputs;
// each class is then a GEMM with K = k^2/4 taps and N = its own modes only: 4 x (k^2/4) x (M/4) multiply-adds per
// patch instead of k^2 x M.  The patch bytes still cross HBM once -- the kernel stays HBM-bound, but with a third
// of the tensor-pipe work (padding of the class widths to 16 costs the rest) and a quarter of the basis stream
// it no longer runs into the board's power limit, and wide operands (n_max = 20) stop being bound by the
// L2 -> SM traffic of the basis.
//
// Arithmetic: F in fp32 (two rounded adds per value), scaled by a power of two so that |F| <= 2^14 (no scale at all for
// data bounded in [2^-2, 2^13]), then the fp16 split of project_tc3_kernel: F = x1 + x2, V = b1 + b2, three kind::f16
// MMAs x1.b1 + x2.b1 + x1.b2 per 16 taps.  The bound of |x| comes from the caller (value_max), from a sample of the
// stack with the range-free tf32x3 kernel enqueued behind as a conditional fallback (auto-range), or -- gathering
// form -- from the frame itself.
//
// Layout per CTA pair (cta_group::2, M = 256 = 128 patches per CTA, each CTA stages half of every class's rows):
//   X ring     per stage the four 32-tap boxes (row i | its mirror half | row i' | its mirror half) of 128 patches
//   B ring     per super-block (32 folded taps) the CTA's half of the class rows, [32 x b1 | 32 x b2] per row
//   TMEM       accumulators [class A_re | A_im | B_re | B_im] (two sets when they fit) + staging units of
//              4 classes x (8 columns x1 | 8 columns x2) = 16 folded taps each
//   n_max <= 13: warps 0-7 two splitter warpgroups (butterfly + split + tcgen05.st, active units in turn), warps 8-11
//                one epilogue warpgroup (running sums of the K chunks for all four classes, fused store path)
//   n_max <= 20: warps 0-3 splitter, warps 4-11 two epilogue warpgroups (class pair A / pair B)
//   warp 12    X TMA   13 MMA issuer   14 TMEM allocation + K5 pusher   15 basis TMA
//   gathering form (K2 fused, n_max <= 13): warps 16-19 copy the windows from four shifted frame planes into the X ring
#include "zb200_common.cuh"
#include "zb200_tc_ptx.cuh"
#include "zb200_project_shared.cuh"
#include "zb200_basis_math.h"

#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <exception>
#include <type_traits>
#include <vector>

namespace zb200 {
namespace fold {
using namespace tc;

// class widths (operand rows = accumulator columns, multiples of 16) of the two compiled configurations:
//   0: n_max <= 13 (every class <= 32 modes)      1: n_max <= 20 (A_re <= 80, the others <= 64)
// class order = accumulator column order: A_re (m even, cos; m = 0 last), A_im (m even, sin), B_re (m odd, cos),
// B_im (m odd, sin): slot s of X_re pairs with slot s of X_im (the complex moment Z[n,+m] + i Z[n,-m]).
__host__ __device__ constexpr int cls_w(int cfg, int c) { return cfg == 0 ? 32 : (c == 0 ? 80 : 64); }
__host__ __device__ constexpr int cls_off(int cfg, int c) { return c == 0 ? 0 : cls_off(cfg, c - 1) + cls_w(cfg, c - 1); }
__host__ __device__ constexpr int cfg_cols(int cfg) { return cls_off(cfg, 3) + cls_w(cfg, 3); }
constexpr int kSlotMax = 80;
constexpr int kUnitCols = 64;                 // TMEM columns of one staging unit
constexpr int kXStage = 4 * kTileRows * 128;  // four boxes of 128 rows x 128 bytes
// waits of the roles that are idle most of the time (epilogue between chunks, the producers) carry a suspend-time
// hint: the hardware parks the warp until the phase completes instead of re-issuing try_wait -- the spin loops were
// 56 % of the kernel's executed instructions (ncu source page), issue slots and power the HBM-bound kernel needs
constexpr uint32_t kParkNs = 20000;
__device__ __forceinline__ void wait_idle(uint64_t* bar, uint32_t parity, uint32_t park_ns) {
    if (park_ns) mbar_wait_parked(bar, parity, park_ns);
    else mbar_wait(bar, parity);
}

template <int kCfg, bool kGather> struct Regs;   // setmaxnreg budgets: (ctl + split + epi_a + epi_b) * 128 <= 64 Ki
template <> struct Regs<0, false> { static constexpr int ctl = 48, split = 144, epi_a = 176, epi_b = 176, gather = 0; };   // [split | split | epilogue | control]
template <> struct Regs<1, false> { static constexpr int ctl = 48, split = 96, epi_a = 200, epi_b = 168, gather = 0; };   // 144 / 128 running sums
// K2 fused (a fifth warpgroup gathers the windows): 640 threads start with 96 registers, the pool is 480 x 128
// and the warpgroups are [splitter | splitter | epilogue (both class pairs, 128 running sums) | control | gather]
template <> struct Regs<0, true> { static constexpr int ctl = 40, split = 96, epi_a = 176, epi_b = 176, gather = 72; };

struct Params {
    long long n_patches;
    int n_tiles;              // 128-patch tiles
    // balanced partition (gathering form; 0 = tiles dealt round-robin): CTA b owns rows [b R, (b+1) R) and walks them in
    // 128-row tiles -- a frame's ~21 k windows are 164 tiles on 148 CTAs, i.e. two full tile times round-robin, but 1.1
    // when every CTA takes 141 rows (rows past its range are zero-filled copies that touch no memory)
    long long rows_per_cta;
    int sb_count;             // super-blocks (32 folded taps) with taps inside the unit disk
    const unsigned short* sb_list;   // their indices (window row i = sb / nj, 32-tap column block jb = sb % nj), ascending
    const unsigned char* umask;   // per super-block (absolute index): bit h = its 16-tap unit h holds taps of the unit disk
    int k, nj;                // window side, 32-tap boxes per half row (k / 64)
    int n_stages, b_stages, n_units, acc_bufs, chunk_sb;
    int out_kind;             // ZB200_OUT_*
    int row_len;              // output row length in floats
    float* out;
    float* out2;
    const short* col_real;    // [cols] real mode of an accumulator column, -1 = padding
    const short* slot_cplx;   // [2][kSlotMax] complex mode of slot s of the class pair A / B, -1 = none
    int n_peers;
    uint32_t push_off;
    float* peer_out[kMaxPeers];
    float x_scale, out_scale;
    // auto-range (no value_max from the caller): absmax = bits of the largest sampled |x| (range_sample_kernel); the
    // kernel derives the power-of-two scale from it; flag is raised when a result is not finite (an unsampled value
    // beyond 8x the sampled maximum overflowed fp16) -- the caller then recomputes with the range-free kernel
    // K2 fused into K3: windows are gathered from four shifted, zero-padded copies of the frame (16-byte aligned
    // window rows: plane r [y][u] = frame[y][u - g_L + r], pitch g_Wp) at the corners g_xy
    const float* g_planes;
    const int2* g_xy;
    int g_H, g_Wp, g_L;
    uint32_t park_ns;         // suspend-time hint of the idle roles' waits (0 = plain try_wait loops)
    const uint32_t* absmax;
    unsigned* flag;
    float inv_area;
    // per class, for the issuing thread (read through the constant bank in a rolled loop: with the class constants
    // unrolled into the instruction stream ptxas ran out of UNIFORM registers as soon as the widths differed)
    uint32_t cls_idesc[4];    // instruction descriptor: kind::f16, M = 256, N = class width
    uint32_t cls_brow[4];     // (byte offset of the class's rows inside a B stage) >> 4
    uint32_t cls_col[4];      // first accumulator column
};

__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(__float_as_uint(lo)), "r"(__float_as_uint(hi)));
    return v;
}
__device__ __forceinline__ void unpk2(uint64_t v, uint32_t& lo, uint32_t& hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// fp16 split of a pair of folded values: x1 = their top 11 significand bits (exact in fp16), x2 = RN_f16(F - x1)
__device__ __forceinline__ void split_pair(uint64_t f, uint32_t& w1, uint32_t& w2) {
    uint32_t lo, hi;
    unpk2(f, lo, hi);
    const uint32_t tl = lo & 0xFFFFE000u, th = hi & 0xFFFFE000u;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w1) : "f"(__uint_as_float(th)), "f"(__uint_as_float(tl)));   // low half = first tap
    const uint64_t r = sub2(f, pk2(__uint_as_float(tl), __uint_as_float(th)));
    uint32_t rl, rh;
    unpk2(r, rl, rh);
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w2) : "f"(__uint_as_float(rh)), "f"(__uint_as_float(rl)));
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// power-of-two input scale for |x| <= bound (bits of a non-negative float): |x| 2^sh <= 2^11, so the four-term fold
// stays below 2^13 and an unsampled value up to 8x the bound still converts to a finite fp16
__device__ __forceinline__ int auto_shift(uint32_t bound_bits) {
    const int e = (int)((bound_bits >> 23) & 255u) - 126;          // bound < 2^e  (frexp exponent)
    const int sh = (bound_bits == 0u || (e >= -1 && e <= 13)) ? 0 : 11 - e;     // 2^-2 <= bound < 2^13: fp16 range as is
    return sh < -100 ? -100 : (sh > 100 ? 100 : sh);
}
__device__ __forceinline__ float pow2i(int sh) { return __uint_as_float((uint32_t)(127 + sh) << 23); }

// largest |x| of `n_sample` patches spread evenly over the stack (one block per sampled patch)
__global__ void range_sample_kernel(const float* __restrict__ x, long long n, int kk, int n_sample, uint32_t* __restrict__ absmax) {
    const long long patch = (long long)blockIdx.x * n / n_sample;
    const float4* src = reinterpret_cast<const float4*>(x + patch * (long long)kk);
    uint32_t m = 0;
    for (int i = threadIdx.x; i < kk / 4; i += blockDim.x) {
        const float4 v = __ldg(src + i);
        m = max(max(m, __float_as_uint(v.x) & 0x7FFFFFFFu), max(__float_as_uint(v.y) & 0x7FFFFFFFu,
                max(__float_as_uint(v.z) & 0x7FFFFFFFu, __float_as_uint(v.w) & 0x7FFFFFFFu)));
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(absmax, m);       // non-negative floats (and NaNs above them) order like their bits
}

// ---- epilogue of one class pair (X_re, X_im): kRe / kIm = their widths in 16-column chunks ----------------------
// The thread owns one patch: running fp32 sums of the K chunks (round-to-nearest adds; the tensor core's own fp32
// accumulation truncates), then the store path of kOut.  Columns are class-ordered, so real moments go out through
// the column -> mode table and complex ones pair chunk cc of X_re with chunk cc of X_im.
// kPairs = 2 (the gathering form, n_max <= 13): ONE epilogue warpgroup owns both class pairs (A then B, equal widths),
// because its second warpgroup's threads went to a second splitter warpgroup.
template <int kOut, int kRe, int kIm, int kPairs = 1>
__device__ __forceinline__ void epilogue_pair(const Params& p, uint32_t tmem_base, int n_cols, int col0, int g, int q, int lane,
                                              bool remote, uint64_t* acc_full, uint64_t* acc_empty, unsigned* out_done,
                                              int my_tiles, int n_chunks, float out_scale) {
    constexpr int kCP = kRe + kIm;            // chunks of one pair
    constexpr int kCC = kPairs * kCP;
    uint32_t ck = 0;
    for (int t = 0; t < my_tiles; ++t) {
        const long long row = (p.rows_per_cta ? (long long)blockIdx.x * p.rows_per_cta + (long long)t * kTileRows
                                              : ((long long)blockIdx.x + (long long)t * gridDim.x) * kTileRows) + q * 32 + lane;
        const long long row_end = p.rows_per_cta ? min(p.n_patches, ((long long)blockIdx.x + 1) * p.rows_per_cta) : p.n_patches;
        float sum[kCC][16];
#pragma unroll
        for (int cc = 0; cc < kCC; ++cc)
#pragma unroll
            for (int i = 0; i < 16; ++i) sum[cc][i] = 0.f;
        for (int c = 0; c < n_chunks; ++c, ++ck) {
            const int buf = p.acc_bufs == 2 ? (int)(ck & 1) : 0;
            wait_idle(&acc_full[buf], p.acc_bufs == 2 ? ((ck >> 1) & 1u) : (ck & 1u), p.park_ns);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * n_cols + col0);
#pragma unroll
            for (int cc = 0; cc < kCC; ++cc) {
                uint32_t v[16];
                tmem_ld16(taddr + cc * 16, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) sum[cc][i] += __uint_as_float(v[i]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (remote) mbar_arrive_cluster(mapa_cluster(smem_u32(&acc_empty[buf]), 0));
                else mbar_arrive(&acc_empty[buf]);
            }
        }
        if (row < row_end) {
            const float sc = out_scale;
            if (p.flag) {
                uint32_t worst = 0;
#pragma unroll
                for (int cc = 0; cc < kCC; ++cc)
#pragma unroll
                    for (int i = 0; i < 16; ++i) worst = max(worst, __float_as_uint(sum[cc][i]) & 0x7FFFFFFFu);
                if (worst >= 0x7F800000u) atomicOr(p.flag, 1u);
            }
            if (kOut == kOutPlain && p.out_kind == ZB200_OUT_REAL) {
                float* dst = p.out + row * (long long)p.row_len;
#pragma unroll
                for (int cc = 0; cc < kCC; ++cc)
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int j = __ldg(p.col_real + col0 + cc * 16 + i);
                        if (j >= 0) dst[j] = sum[cc][i] * sc;
                    }
            } else {
#pragma unroll
                for (int pr = 0; pr < kPairs; ++pr)
#pragma unroll
                for (int cc = 0; cc < kRe; ++cc)
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int c = __ldg(p.slot_cplx + (g + pr) * kSlotMax + cc * 16 + i);
                        if (c < 0) continue;
                        const float re = sum[pr * kCP + cc][i] * sc;
                        const float im = cc < kIm ? sum[pr * kCP + (cc < kIm ? kRe + cc : 0)][i] * sc : 0.f;
                        if constexpr (kOut == kOutPlain) {
                            *reinterpret_cast<float2*>(p.out + row * (long long)p.row_len + 2 * c) = make_float2(re, im);
                        } else {
                            p.out[row * (long long)p.row_len + c] = fast_abs2(re, im);
                            if constexpr (kOut == kOutAbsPhase) p.out2[row * (long long)p.row_len + c] = compact_atan2(im, re);
                        }
                    }
            }
        }
        if (p.n_peers) {
            __threadfence();                              // this tile's rows are in L2 before the pusher is told
            asm volatile("fence.proxy.async.global;" ::: "memory");
            __syncwarp();
            if (lane == 0) atomicAdd(out_done, 1u);
        }
    }
}

template <int kOut, int kCfg, bool kGather>
__global__ void __launch_bounds__(kGather ? 640 : 512, 1)
project_fold_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_b, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // n_max <= 13 (128 accumulator columns): warpgroups [splitter | splitter | epilogue for both class pairs | control
    // (| gather)]; n_max <= 20 (272 columns, two epilogue warpgroups of running sums): [splitter | epilogue A | epilogue B |
    // control].  A unit is ~0.8 us of ONE warp's instruction chain per sub-partition -- 45 us per tile against 49 us of HBM
    // time at full clock, but the chain stretches with the SM clock under the power cap while HBM does not.
    constexpr bool kTwoSplit = kCfg == 0;
    static_assert(!kGather || kTwoSplit, "the gathering form exists for n_max <= 13 only");
    constexpr int kCols = cfg_cols(kCfg);
    constexpr uint32_t kBStage = (uint32_t)(kCols / 2) * 128u;    // this CTA's half of the class rows, one super-block
    uint8_t* b_ring = smem + (size_t)p.n_stages * kXStage;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + (size_t)p.b_stages * kBStage);
    uint64_t* xfull = bars;                        // TMA landed                                  [n_stages]
    uint64_t* xempty = bars + p.n_stages;          // the splitter warps hold the stage's rows    [n_stages]
    uint64_t* bfull = bars + 2 * p.n_stages;       // own half of a B super-block landed          [4]
    uint64_t* bempty = bfull + 4;                  // MMAs reading it retired (both CTAs)         [4]
    uint64_t* bpeer = bempty + 4;                  // leader: the peer's half landed              [4]
    uint64_t* lo_full = bpeer + 4;                 // leader: a staging unit is in TMEM, both CTAs [4]
    uint64_t* lo_empty = lo_full + 4;              // MMAs reading the unit retired               [4]
    uint64_t* acc_full = lo_empty + 4;             // accumulation chunk complete                 [2]
    uint64_t* acc_empty = acc_full + 2;            // leader: chunk drained by both CTAs' epilogues [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    unsigned* out_done = reinterpret_cast<unsigned*>(tmem_slot + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wg = warp >> 2;
    constexpr int kWarpTma = 12, kWarpMma = 13, kWarpAlloc = 14, kWarpBasis = 15;

    if (warp == kWarpTma && lane == 0) {
        if (smem_u32(smem) & 1023u) asm volatile("trap;");        // the swizzled stages rely on the 1024-byte base
        prefetch_tmap(&map_x);
        prefetch_tmap(&map_b);
        *out_done = 0;
    }
    if (warp == kWarpMma && lane == 0) {
        for (int s = 0; s < p.n_stages; ++s) {
            mbar_init(&xfull[s], kGather ? 128 : 1);     // gathered: every gathering thread's copies have landed
            mbar_init(&xempty[s], kTwoSplit ? 8 : 4);    // the splitter warps of one or two warpgroups read a stage
        }
        for (int b = 0; b < 4; ++b) {
            mbar_init(&bfull[b], 1);
            mbar_init(&bempty[b], 1);
            mbar_init(&bpeer[b], 1);
            mbar_init(&lo_full[b], 8);
            mbar_init(&lo_empty[b], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], kTwoSplit ? 8 : 16);  // epilogue warps of both CTAs (one or two warpgroups each)
        }
        fence_barrier_init();
    }
    if (warp == kWarpAlloc) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lo_base = tmem_base + (uint32_t)(p.acc_bufs * kCols);
    const uint32_t crank = cluster_rank();
    const int my_tiles = p.rows_per_cta ? (int)((p.rows_per_cta + kTileRows - 1) / kTileRows)
                                        : (p.n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;     // lock-step within the pair
    const int n_chunks = (p.sb_count + p.chunk_sb - 1) / p.chunk_sb;
    float x_scale = p.x_scale, out_scale = p.out_scale;
    if (p.absmax) {
        const int sh = auto_shift(__ldg(p.absmax));
        x_scale = pow2i(sh);
        out_scale = p.inv_area * pow2i(-sh);
    }

    if (wg == 3) {
        reg_dec<Regs<kCfg, kGather>::ctl>();
        if (warp == kWarpTma) {
            // ===================== X producer: the four mirror boxes of a super-block =====================
            if (!kGather && elect_one()) {
                int s = 0;
                uint32_t ph = 0;
                for (int t = 0; t < my_tiles; ++t) {
                    const int row0 = (blockIdx.x + t * gridDim.x) * kTileRows;
                    for (int sbi = 0; sbi < p.sb_count; ++sbi) {
                        const int sb = __ldg(p.sb_list + sbi);
                        const int i = sb / p.nj, jb = sb - i * p.nj;
                        const int c_top = i * p.k, c_bot = (p.k - 1 - i) * p.k;
                        const int c_l = jb * 32, c_r = p.k - 32 - jb * 32;
                        wait_idle(&xempty[s], ph ^ 1, p.park_ns);
                        mbar_arrive_expect_tx(&xfull[s], kXStage);
                        uint8_t* st = smem + (size_t)s * kXStage;
                        tma_load_2d(st, &map_x, &xfull[s], c_top + c_l, row0, kEvictFirst);
                        tma_load_2d(st + 16384, &map_x, &xfull[s], c_top + c_r, row0, kEvictFirst);
                        tma_load_2d(st + 32768, &map_x, &xfull[s], c_bot + c_l, row0, kEvictFirst);
                        tma_load_2d(st + 49152, &map_x, &xfull[s], c_bot + c_r, row0, kEvictFirst);
                        if (++s == p.n_stages) { s = 0; ph ^= 1; }
                    }
                }
            }
            __syncwarp();
        } else if (warp == kWarpBasis) {
            // ===================== basis producer: this CTA's half of the class rows =====================
            if (elect_one()) {
                int sb = 0;
                uint32_t phb = 0;
                for (int t = 0; t < my_tiles; ++t)
                    for (int sbi = 0; sbi < p.sb_count; ++sbi) {
                        wait_idle(&bempty[sb], phb ^ 1, p.park_ns);
                        mbar_arrive_expect_tx(&bfull[sb], kBStage);
                        tma_load_2d(b_ring + (size_t)sb * kBStage, &map_b, &bfull[sb], (int)__ldg(p.sb_list + sbi) * 32,
                                    (int)crank * (kCols / 2), kEvictLast);
                        if (++sb == p.b_stages) { sb = 0; phb ^= 1; }
                    }
            }
            __syncwarp();
        } else if (warp == kWarpMma) {
            if (crank != 0) {
                // peer CTA: no MMAs to issue; relay "my half of the B super-block landed" to the leader
                if (elect_one()) {
                    int sb = 0;
                    uint32_t phb = 0;
                    for (int t = 0; t < my_tiles; ++t)
                        for (int sbi = 0; sbi < p.sb_count; ++sbi) {
                            wait_idle(&bfull[sb], phb, p.park_ns);
                            mbar_arrive_cluster(mapa_cluster(smem_u32(&bpeer[sb]), 0));
                            if (++sb == p.b_stages) { sb = 0; phb ^= 1; }
                        }
                }
                __syncwarp();
            } else if (elect_one()) {
                // ===================== MMA issuer (leader): 4 classes x 3 MMAs per 16-tap unit =====================
                const uint32_t b_lo0 = desc_lo_sw128(smem_u32(b_ring));
                int sb = 0;
                uint32_t phb = 0;
                int u = 0;
                uint32_t uph = 0;
                uint32_t ck = 0;
                for (int t = 0; t < my_tiles; ++t) {
                    for (int c = 0; c < n_chunks; ++c, ++ck) {
                        const int buf = p.acc_bufs == 2 ? (int)(ck & 1) : 0;
                        const uint32_t acc_par = p.acc_bufs == 2 ? ((ck >> 1) & 1u) : (ck & 1u);
                        mbar_wait_cluster(&acc_empty[buf], acc_par ^ 1u);
                        tc_fence_after();
                        const uint32_t d0 = tmem_base + (uint32_t)(buf * kCols);
                        const int sb_end = min(p.sb_count, (c + 1) * p.chunk_sb);
                        uint32_t acc_run = 0u;
                        for (int sbi = c * p.chunk_sb; sbi < sb_end; ++sbi) {
                            const uint32_t um = __ldg(p.umask + __ldg(p.sb_list + sbi));
                            mbar_wait_cluster(&bpeer[sb], phb);
                            mbar_wait(&bfull[sb], phb);
                            const uint32_t bl = b_lo0 + (uint32_t)sb * (kBStage >> 4);
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                if (!((um >> h) & 1u)) continue;
                                mbar_wait_cluster(&lo_full[u], uph);
                                tc_fence_after();
                                const uint32_t a0 = lo_base + (uint32_t)u * kUnitCols;
#pragma unroll 1
                                for (int cl = 0; cl < 4; ++cl) {
                                    const uint32_t idesc = p.cls_idesc[cl];
                                    const uint32_t brow = bl + p.cls_brow[cl];
                                    const uint64_t b1d = desc_from_lo(brow + 2 * h), b2d = desc_from_lo(brow + 4 + 2 * h);
                                    const uint32_t d = d0 + p.cls_col[cl], a1 = a0 + 16 * cl, a2 = a1 + 8;
                                    umma_bf16_ts_2(d, a1, b1d, idesc, acc_run);
                                    umma_bf16_ts_2(d, a2, b1d, idesc, 1u);
                                    umma_bf16_ts_2(d, a1, b2d, idesc, 1u);
                                }
                                acc_run = 1u;
                                umma_commit_2mc(&lo_empty[u], 3);
                                if (++u == p.n_units) { u = 0; uph ^= 1u; }
                            }
                            umma_commit_2mc(&bempty[sb], 3);
                            if (sbi == sb_end - 1) umma_commit_2mc(&acc_full[buf], 3);
                            if (++sb == p.b_stages) { sb = 0; phb ^= 1; }
                        }
                    }
                }
            }
            __syncwarp();
        } else if (warp == kWarpAlloc && p.n_peers) {
            pusher_loop(p, out_done, kTwoSplit ? 4 : 8, my_tiles, lane, smem + p.push_off, kTileRows);
        }
    } else if (wg == 0 || (kTwoSplit && wg == 1)) {
        // ===================== butterfly + fp16 split, one patch row per thread =====================
        // kTwoSplit: two splitter warpgroups take the active units in turn
        if constexpr (Regs<kCfg, kGather>::split > (kGather ? 96 : 128)) reg_inc<Regs<kCfg, kGather>::split>();
        else reg_dec<Regs<kCfg, kGather>::split>();
        const int r = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        uint32_t unit_no = 0;             // running count of active units: warpgroup wg owns those with unit_no % 2 == wg
        const uint32_t swz = (uint32_t)(r & 7) << 4;
        const uint32_t row_u32 = smem_u32(smem) + (uint32_t)r * 128u;
        const uint64_t sc2 = pk2(x_scale, x_scale);
        const bool unscaled = x_scale == 1.f;
        int s = 0;
        uint32_t ph = 0;
        int u = 0;
        uint32_t uph = 0;
        for (int t = 0; t < my_tiles; ++t) {
            for (int sbi = 0; sbi < p.sb_count; ++sbi) {
                const uint32_t um = __ldg(p.umask + __ldg(p.sb_list + sbi));
                mbar_wait(&xfull[s], ph);
                const uint32_t rowp = row_u32 + (uint32_t)s * kXStage;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (!((um >> h) & 1u)) continue;
                    if (kTwoSplit && (int)(unit_no++ & 1u) != wg) {        // the other splitter warpgroup's unit
                        if (++u == p.n_units) { u = 0; uph ^= 1u; }
                        continue;
                    }
                    // folded taps e = 16h .. 16h+15 of the super-block: a = row i, b = its mirror (box 1, tap 31-e),
                    // c = row i', d = its mirror (box 3)
                    float4 A[4], B[4], C[4], D[4];
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const uint32_t fw = ((uint32_t)(4 * h + c4) << 4) ^ swz;           // forward chunk
                        const uint32_t bw = ((uint32_t)(7 - 4 * h - c4) << 4) ^ swz;       // chunk of the mirrored taps
                        A[c4] = lds128(rowp + fw);
                        B[c4] = lds128(rowp + 16384u + bw);
                        C[c4] = lds128(rowp + 32768u + fw);
                        D[c4] = lds128(rowp + 49152u + bw);
                    }
                    mbar_wait(&lo_empty[u], uph ^ 1u);
                    tc_fence_after();
                    uint32_t w[4][16];       // per class: 8 columns x1 (tap pairs), 8 columns x2
                    // kScaled = false: data whose bound lies in [2^-2, 2^13] needs no power-of-two scale to sit in fp16
                    // range (|F| <= 2^15; the residual's subnormal rounding is 2^-25 absolute): four packed multiplies
                    // per tap pair less, a tenth of the splitter's instructions
                    auto butterfly_split = [&](auto scaled_tag) {
                        constexpr bool kScaled = decltype(scaled_tag)::value;
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) {
#pragma unroll
                            for (int hp = 0; hp < 2; ++hp) {
                                // taps (e, e+1), e = 4 c4 + 2 hp; mirrored chunk holds taps 31-e-3 .. 31-e: reversed order
                                const uint64_t a = hp ? pk2(A[c4].z, A[c4].w) : pk2(A[c4].x, A[c4].y);
                                const uint64_t c = hp ? pk2(C[c4].z, C[c4].w) : pk2(C[c4].x, C[c4].y);
                                const uint64_t b = hp ? pk2(B[c4].y, B[c4].x) : pk2(B[c4].w, B[c4].z);
                                const uint64_t d = hp ? pk2(D[c4].y, D[c4].x) : pk2(D[c4].w, D[c4].z);
                                uint64_t pp = add2(a, b), qq = sub2(a, b), rr = add2(c, d), ss = sub2(c, d);
                                if constexpr (kScaled) {
                                    pp = mul2(pp, sc2);
                                    qq = mul2(qq, sc2);
                                    rr = mul2(rr, sc2);
                                    ss = mul2(ss, sc2);
                                }
                                const int col = 2 * c4 + hp;
                                split_pair(add2(pp, rr), w[0][col], w[0][8 + col]);     // A_re: even in x, even in y
                                split_pair(sub2(qq, ss), w[1][col], w[1][8 + col]);     // A_im: odd, odd
                                split_pair(add2(qq, ss), w[2][col], w[2][8 + col]);     // B_re: odd in x, even in y
                                split_pair(sub2(pp, rr), w[3][col], w[3][8 + col]);     // B_im: even in x, odd in y
                            }
                        }
                    };
                    if (unscaled) butterfly_split(std::false_type{});
                    else butterfly_split(std::true_type{});
                    const uint32_t ta = lo_base + lane_addr + (uint32_t)u * kUnitCols;
#pragma unroll
                    for (int cl = 0; cl < 4; ++cl) tmem_st16(ta + 16 * cl, w[cl]);
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (crank != 0) mbar_arrive_cluster(mapa_cluster(smem_u32(&lo_full[u]), 0));
                        else mbar_arrive(&lo_full[u]);
                    }
                    if (++u == p.n_units) { u = 0; uph ^= 1u; }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&xempty[s]);          // every row of the stage is in registers / TMEM
                if (++s == p.n_stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (kGather && wg == 4) {
        // ===================== K2: gather the windows of 128 patches into the X ring =====================
        // one warp per 32 tile rows; a warp instruction copies the 32-float segments of 4 rows (8 lanes x 16 bytes each)
        reg_dec<Regs<kCfg, kGather>::gather>();
        const int gw = warp - 16;
        const int ch = lane & 7;
        int s = 0;
        uint32_t ph = 0;
        // per lane, the 8 tile rows it serves (row j = 4 it + lane / 8 of this warp's 32): everything that does not
        // depend on the super-block is computed once per tile -- the first version spent ~45 instructions per copy
        uint32_t dst_row[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int j = 4 * it + (lane >> 3);
            dst_row[it] = (uint32_t)j * 128u + (((uint32_t)ch ^ (uint32_t)(j & 7)) << 4);
        }
        const unsigned x_hi = (unsigned)(p.g_Wp - 4);
        for (int t = 0; t < my_tiles; ++t) {
            const long long base = (p.rows_per_cta ? (long long)blockIdx.x * p.rows_per_cta + (long long)t * kTileRows
                                                   : ((long long)blockIdx.x + (long long)t * gridDim.x) * kTileRows) + gw * 32;
            const long long row_end = p.rows_per_cta ? min(p.n_patches, ((long long)blockIdx.x + 1) * p.rows_per_cta) : p.n_patches;
            int2 cn = make_int2(-(1 << 28), -(1 << 28));               // rows past the end: every copy zero-fills
            if (base + lane < row_end) cn = __ldg(p.g_xy + base + lane);
            // all 32 windows of this warp valid and fully inside the frame?  (padded planes: x0 in [0, W - k], y0 in [0, H - k])
            const bool interior = __all_sync(0xffffffffu, base + lane < row_end && cn.x >= 0 && cn.y >= 0 &&
                                                              cn.x + p.k <= p.g_Wp - 2 * p.g_L && cn.y + p.k <= p.g_H);
            const float* row_ptr[8];      // plane (x0 & 3), frame row y0, padded column x0 - (x0 & 3) + g_L + 4 ch
            int row_y[8], row_x[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int j = 4 * it + (lane >> 3);
                const int x0 = __shfl_sync(0xffffffffu, cn.x, j), y0 = __shfl_sync(0xffffffffu, cn.y, j);
                const int xl = x0 + p.g_L + 4 * ch;                   // segment starts and g_L are multiples of 4
                const int rr = xl & 3;
                row_y[it] = y0;
                row_x[it] = xl - rr;
                const bool sane = y0 > -(1 << 20) && y0 < (1 << 20) && x0 > -(1 << 20) && x0 < (1 << 20);
                row_ptr[it] = p.g_planes + (sane ? ((long long)rr * p.g_H + y0) * p.g_Wp + (xl - rr) : 0);
                if (!sane) row_y[it] = -(1 << 28);
            }
            for (int sbi = 0; sbi < p.sb_count; ++sbi) {
                const int sb = __ldg(p.sb_list + sbi);
                const int i = sb / p.nj, jb = sb - i * p.nj;
                wait_idle(&xempty[s], ph ^ 1, p.park_ns);
                const uint32_t xs = smem_u32(smem) + (uint32_t)s * kXStage + (uint32_t)(gw * 32) * 128u;
#pragma unroll
                for (int box = 0; box < 4; ++box) {
                    const int wr = (box < 2) ? i : p.k - 1 - i;
                    const int wc = (box & 1) ? p.k - 32 - jb * 32 : jb * 32;
                    const long long off = (long long)wr * p.g_Wp + wc;
                    const uint32_t xb = xs + (uint32_t)box * 16384u;
                    if (interior) {
                        // every window of this warp lies inside the frame (what clear_border leaves): no range checks --
                        // the gather warps' instruction stream, at ~15 instructions per copy, paced the whole tile
#pragma unroll
                        for (int it = 0; it < 8; ++it)
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(xb + dst_row[it]), "l"(row_ptr[it] + off) : "memory");
                    } else {
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const bool inb = (unsigned)(row_y[it] + wr) < (unsigned)p.g_H && (unsigned)(row_x[it] + wc) <= x_hi;
                            const float* gp = inb ? row_ptr[it] + off : p.g_planes;
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(xb + dst_row[it]), "l"(gp), "r"(inb ? 16 : 0) : "memory");
                        }
                    }
                }
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&xfull[s])) : "memory");
                if (++s == p.n_stages) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // ===================== epilogue: warpgroup 1 owns the class pair A, warpgroup 2 the pair B =====================
        const int g = wg - 1, q = warp & 3;
        if (g == 0 || kTwoSplit) reg_inc<Regs<kCfg, kGather>::epi_a>();
        else reg_inc<Regs<kCfg, kGather>::epi_b>();
        if constexpr (kTwoSplit) {
            // warpgroup 2 alone: both class pairs (cfg 0: four classes of equal width)
            epilogue_pair<kOut, cls_w(kCfg, 0) / 16, cls_w(kCfg, 0) / 16, 2>(p, tmem_base, kCols, 0, 0, q, lane, crank != 0,
                                                                            acc_full, acc_empty, out_done, my_tiles, n_chunks, out_scale);
        } else if (g == 0)
            epilogue_pair<kOut, cls_w(kCfg, 0) / 16, cls_w(kCfg, 1) / 16>(p, tmem_base, kCols, cls_off(kCfg, 0), 0, q, lane, crank != 0,
                                                                          acc_full, acc_empty, out_done, my_tiles, n_chunks, out_scale);
        else
            epilogue_pair<kOut, cls_w(kCfg, 2) / 16, cls_w(kCfg, 3) / 16>(p, tmem_base, kCols, cls_off(kCfg, 2), 1, q, lane, crank != 0,
                                                                          acc_full, acc_empty, out_done, my_tiles, n_chunks, out_scale);
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                           // no CTA leaves while its peer may still signal it
    if (warp == kWarpAlloc) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// ---- plan time: the folded fp16-split operand ---------------------------------------------------------------------
// row_mode[r] = real mode of operand row r (-1 = zero row), row_sj / row_si = its parities under the column / row
// mirror.  Row r, super-block sb (window row i = sb / nj, columns jb*32 .. +31), tap t:
//   Vq = (V[i][j] + sj V[i][j'] + si V[i'][j] + sj si V[i'][j']) / 4     (= V[i][j] up to the basis' own rounding)
//   [32 x b1 | 32 x b2] halves, b1 = RN_f16(Vq), b2 = RN_f16(Vq - b1)
__global__ void fold_pack_kernel(const double* __restrict__ basis, const short* __restrict__ row_mode,
                                 const signed char* __restrict__ row_sj, const signed char* __restrict__ row_si, int k, int nj,
                                 int n_sb, __half* __restrict__ fb, double* __restrict__ asym) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;          // folded tap index: sb * 32 + t
    const int r = blockIdx.y;
    if (e >= n_sb * 32) return;
    const int sb = e >> 5, t = e & 31;
    const int i = sb / nj, j = (sb - i * nj) * 32 + t;
    const int src = row_mode[r];
    double vq = 0.0;
    if (src >= 0) {
        const double* v = basis + (size_t)src * k * k;
        const double sj = row_sj[r], si = row_si[r];
        const double v00 = v[i * k + j], v01 = v[i * k + (k - 1 - j)], v10 = v[(k - 1 - i) * k + j], v11 = v[(k - 1 - i) * k + (k - 1 - j)];
        vq = 0.25 * (v00 + sj * v01 + si * v10 + sj * si * v11);
        const double dev = fmax(fmax(fabs(v00 - sj * v01), fabs(v00 - si * v10)), fabs(v00 - sj * si * v11));
        if (dev > 0.0) atomicMax(reinterpret_cast<unsigned long long*>(asym), (unsigned long long)__double_as_longlong(dev));
    }
    const __half b1 = __double2half(vq);
    const size_t o = (size_t)r * n_sb * 64 + (size_t)sb * 64 + t;
    fb[o] = b1;
    fb[o + 32] = __double2half(vq - (double)__half2float(b1));
}

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
static int encode_2d(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t pitch_bytes, uint32_t box_rows) {
    auto enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return ZB200_ECUDA;
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch_bytes};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (fold: cols=%llu rows=%llu box_rows=%u)", (int)r,
                  (unsigned long long)cols, (unsigned long long)rows, box_rows);
        return ZB200_ECUDA;
    }
    return ZB200_OK;
}

}  // namespace fold

void free_fold_operand(zb200_plan* p) {
    FoldOperand& f = p->fold;
    cudaFree(f.fb);
    cudaFree(f.d_umask);
    cudaFree(f.d_sb_list);
    cudaFree(f.d_col_real);
    cudaFree(f.d_slot_cplx);
    f = FoldOperand();
}

// Builds the folded operand when the shape qualifies (sm_100, window side a multiple of 64, n_max <= 20); any other
// plan simply has fold.ready == false and keeps the unfolded kernels.
static int init_fold_operand_impl(zb200_plan* p);
int init_fold_operand(zb200_plan* p) {
    try {                                   // host tables live in std::vector: nothing may throw across the C ABI
        return init_fold_operand_impl(p);
    } catch (const std::exception& e) {
        set_error("fold operand: %s", e.what());
        free_fold_operand(p);
        return ZB200_ENOMEM;
    }
}
static int init_fold_operand_impl(zb200_plan* p) {
    using namespace fold;
    FoldOperand& f = p->fold;
    const int k = p->size;
    if (p->cc_major != 10 || k % 64 != 0 || k > 128 || p->n_max > 20) return ZB200_OK;
    // class sizes: m even cos (incl. m = 0) | m even sin | m odd cos | m odd sin
    int count[4] = {0, 0, 0, 0};
    for (int j = 0; j < p->n_modes; ++j) {
        const int m = p->h_m[j];
        count[(m & 1) * 2 + (m < 0 ? 1 : 0)]++;
    }
    int cfg = -1;
    for (int c = 0; c < 2 && cfg < 0; ++c) {
        bool ok = true;
        for (int cl = 0; cl < 4; ++cl) ok = ok && count[cl] <= cls_w(c, cl);
        if (ok) cfg = c;
    }
    if (cfg < 0) return ZB200_OK;
    const int cols = cfg_cols(cfg);
    f.cfg = cfg;
    f.cols = cols;
    f.nj = k / 64;
    f.n_sb = (k / 2) * f.nj;

    // slots: (n, m) ascending inside a class, m = 0 modes behind the m > 0 ones of A_re, so that slot s of X_re and
    // slot s of X_im are the two real planes of one complex mode
    std::vector<short> col_real((size_t)cols, -1), slot_cplx((size_t)2 * kSlotMax, -1), row_mode((size_t)cols, -1);
    std::vector<signed char> row_sj((size_t)cols, 1), row_si((size_t)cols, 1);
    int fill[4] = {0, 0, 0, 0};
    auto place = [&](int cl, int mode) {
        const int slot = fill[cl]++;
        col_real[(size_t)cls_off(cfg, cl) + slot] = (short)mode;
        // HBM row: [rank][class][slot within the rank's half]
        const int half = cls_w(cfg, cl) / 2, rank = slot / half;
        const int row = rank * (cols / 2) + cls_off(cfg, cl) / 2 + (slot - rank * half);
        row_mode[(size_t)row] = (short)mode;
        const bool odd = cl >= 2, sine = cl & 1;
        row_sj[(size_t)row] = (signed char)((odd != sine) ? -1 : 1);     // cos: (-1)^m, sin: (-1)^(m+1) under the column mirror
        row_si[(size_t)row] = (signed char)(sine ? -1 : 1);              // cos even, sin odd under the row mirror
        return slot;
    };
    auto mode_of = [&](int n, int m) { return mode_index(n, m); };                   // ZPs order: n ascending, m = -n, -n+2, ..
    int c_idx = 0;
    std::vector<int> cplx_of((size_t)p->n_modes, -1);                    // complex index of the (n, |m|) pair, nm2j_complex order
    for (int n = 0; n <= p->n_max; ++n)
        for (int m = n & 1; m <= n; m += 2, ++c_idx) cplx_of[(size_t)mode_of(n, m)] = c_idx;
    for (int pass = 0; pass < 2; ++pass)                                  // pass 0: m > 0, pass 1: m = 0
        for (int n = 0; n <= p->n_max; ++n)
            for (int m = n & 1; m <= n; m += 2) {
                if ((pass == 0) != (m > 0)) continue;
                const int pair = m & 1;                                   // 0 = A (m even), 1 = B (m odd)
                const int s_re = place(2 * pair, mode_of(n, m));
                slot_cplx[(size_t)pair * kSlotMax + s_re] = (short)cplx_of[(size_t)mode_of(n, m)];
                if (m > 0) {
                    const int s_im = place(2 * pair + 1, mode_of(n, -m));
                    if (s_im != s_re) {
                        set_error("fold operand: class slots of (n=%d, m=+-%d) do not line up", n, m);
                        return ZB200_ECUDA;
                    }
                }
            }

    // active units / super-block range (same disk rule as the unfolded k-block mask)
    std::vector<unsigned char> um((size_t)f.n_sb, 0);
    for (int sb = 0; sb < f.n_sb; ++sb) {
        const int i = sb / f.nj, jb = sb % f.nj;
        for (int t = 0; t < 32; ++t) {
            const double y = -1.0 + 2.0 * i / (k - 1), x = -1.0 + 2.0 * (jb * 32 + t) / (k - 1);
            if (x * x + y * y <= 1.0 + 1e-9) um[(size_t)sb] |= (unsigned char)(1u << (t / 16));
        }
    }
    std::vector<unsigned short> sb_list;
    for (int sb = 0; sb < f.n_sb; ++sb)
        if (um[(size_t)sb]) sb_list.push_back((unsigned short)sb);
    if (sb_list.empty()) return ZB200_OK;
    f.sb_count = (int)sb_list.size();

    short *d_row_mode = nullptr;
    signed char *d_sj = nullptr, *d_si = nullptr;
    double* d_asym = nullptr;
    const size_t fb_bytes = (size_t)cols * f.n_sb * 128;
    ZB_CUDA(cudaMalloc(&f.fb, fb_bytes));
    ZB_CUDA(cudaMalloc(&f.d_umask, (size_t)f.n_sb));
    ZB_CUDA(cudaMalloc(&f.d_sb_list, sizeof(unsigned short) * sb_list.size()));
    ZB_CUDA(cudaMemcpy(f.d_sb_list, sb_list.data(), sizeof(unsigned short) * sb_list.size(), cudaMemcpyHostToDevice));
    ZB_CUDA(cudaMalloc(&f.d_col_real, sizeof(short) * col_real.size()));
    ZB_CUDA(cudaMalloc(&f.d_slot_cplx, sizeof(short) * slot_cplx.size()));
    ZB_CUDA(cudaMalloc(&d_row_mode, sizeof(short) * row_mode.size()));
    ZB_CUDA(cudaMalloc(&d_sj, row_sj.size()));
    ZB_CUDA(cudaMalloc(&d_si, row_si.size()));
    ZB_CUDA(cudaMalloc(&d_asym, sizeof(double)));
    ZB_CUDA(cudaMemset(d_asym, 0, sizeof(double)));
    ZB_CUDA(cudaMemcpy(f.d_umask, um.data(), um.size(), cudaMemcpyHostToDevice));
    ZB_CUDA(cudaMemcpy(f.d_col_real, col_real.data(), sizeof(short) * col_real.size(), cudaMemcpyHostToDevice));
    ZB_CUDA(cudaMemcpy(f.d_slot_cplx, slot_cplx.data(), sizeof(short) * slot_cplx.size(), cudaMemcpyHostToDevice));
    ZB_CUDA(cudaMemcpy(d_row_mode, row_mode.data(), sizeof(short) * row_mode.size(), cudaMemcpyHostToDevice));
    ZB_CUDA(cudaMemcpy(d_sj, row_sj.data(), row_sj.size(), cudaMemcpyHostToDevice));
    ZB_CUDA(cudaMemcpy(d_si, row_si.data(), row_si.size(), cudaMemcpyHostToDevice));
    dim3 grid((unsigned)ceil_div((int64_t)f.n_sb * 32, 128), (unsigned)cols);
    fold_pack_kernel<<<grid, 128>>>(p->basis64, d_row_mode, d_sj, d_si, k, f.nj, f.n_sb, reinterpret_cast<__half*>(f.fb), d_asym);
    ZB_LAUNCHED();
    double asym = 0.0;
    ZB_CUDA(cudaMemcpy(&asym, d_asym, sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(d_row_mode);
    cudaFree(d_sj);
    cudaFree(d_si);
    cudaFree(d_asym);
    // the generated planes are mirror-symmetric to rounding (1e-13 at n_max = 20); anything else means the class
    // table above does not describe this basis, and the folded kernel must not be used
    if (!(asym < 1e-9)) {
        free_fold_operand(p);
        return ZB200_OK;
    }
    int rc = encode_2d(&f.tmap, f.fb, (uint64_t)f.n_sb * 32, (uint64_t)cols, (uint64_t)f.n_sb * 128, (uint32_t)(cols / 2));
    if (rc) return rc;
    f.ready = true;
    return ZB200_OK;
}

bool fold_supported(const zb200_plan* p) { return p->fold.ready; }
bool fold_gather_supported(const zb200_plan* p) { return p->fold.ready && p->fold.cfg == 0; }

// value_max > 0: the caller's bound of |x|.  value_max <= 0: auto-range -- d_aux (two zeroed 32-bit words owned by the
// caller: [0] sampled absmax bits, [1] overflow flag) receives the largest |x| of up to 1024 evenly spread patches, the
// kernel scales by it and raises d_aux[1] if any result is not finite (then every row must be recomputed).
// gsrc (optional, n_max <= 13): K2 fused -- the windows are gathered from the shifted planes of a frame by a fifth
// warpgroup instead of being loaded from a patch stack; d_aux[0] must then hold the bits of max |frame| (or value_max > 0).
int project_fold(const zb200_plan* p, const float* d_patches, int64_t n, int out_kind, void* d_out, void* d_out2,
                 cudaStream_t s, const PeerTargets* peers, double value_max, uint32_t* d_aux, const GatherSource* gsrc) {
    using namespace fold;
    const FoldOperand& f = p->fold;
    if (!f.ready) {
        set_error("folded projection: no folded operand for n_max=%d size=%d", p->n_max, p->size);
        return ZB200_EUNSUP;
    }
    if (n == 0) return ZB200_OK;
    ZB_CHECK_ARG(gsrc || (reinterpret_cast<uintptr_t>(d_patches) & 15) == 0, "project: patch pointer must be 16-byte aligned");
    if (gsrc && (f.cfg != 0 || !gsrc->planes)) {
        set_error("folded projection: the fused gather needs n_max <= 13 and the shifted frame planes");
        return ZB200_EUNSUP;
    }
    ZB_CHECK_ARG(value_max > 0.0 || d_aux, "the folded projection needs an upper bound of |patch values| or auto-range scratch");
    ZB_CHECK_ARG(out_kind != ZB200_OUT_COMPLEX || (reinterpret_cast<uintptr_t>(d_out) & 7) == 0,
                 "project: complex output must be 8-byte aligned");
    Params prm{};
    prm.n_patches = n;
    prm.n_tiles = (int)ceil_div(n, kTileRows);
    prm.sb_count = f.sb_count;
    prm.sb_list = f.d_sb_list;
    prm.umask = f.d_umask;
    prm.k = p->size;
    prm.nj = f.nj;
    prm.out_kind = out_kind;
    prm.row_len = out_kind == ZB200_OUT_REAL ? p->n_modes : (out_kind == ZB200_OUT_COMPLEX ? 2 * p->n_complex : p->n_complex);
    prm.out = static_cast<float*>(d_out);
    prm.out2 = static_cast<float*>(d_out2);
    prm.col_real = f.d_col_real;
    prm.slot_cplx = f.d_slot_cplx;
    for (int cl = 0; cl < 4; ++cl) {
        const int w = cls_w(f.cfg, cl);
        prm.cls_idesc[cl] = (1u << 4) | ((uint32_t)(w >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // make_idesc_f16(w, 256)
        prm.cls_brow[cl] = (uint32_t)((cls_off(f.cfg, cl) / 2) * 128) >> 4;
        prm.cls_col[cl] = (uint32_t)cls_off(f.cfg, cl);
    }
    prm.inv_area = (float)p->inv_area;
    if (gsrc) {
        prm.g_planes = gsrc->planes;
        prm.g_xy = gsrc->xy0;
        prm.g_H = gsrc->H;
        prm.g_Wp = gsrc->Wp;
        prm.g_L = gsrc->L;
    }
    if (!(value_max > 0.0)) {
        if (!gsrc) {
            const int n_sample = (int)(n < 1024 ? n : 1024);
            range_sample_kernel<<<n_sample, 256, 0, s>>>(d_patches, (long long)n, p->kk, n_sample, d_aux);
            ZB_LAUNCHED();
            prm.flag = d_aux + 1;
        }                                                   // gathered: d_aux[0] bounds the whole frame, nothing can overflow
        prm.absmax = d_aux;
        prm.x_scale = prm.out_scale = 1.f;
    } else {
        int e = 0;
        frexp(value_max, &e);                               // |x| < 2^e: |x| 2^(12-e) <= 2^12, the four-term fold <= 2^14
        int sh = 12 - e < -100 ? -100 : (12 - e > 100 ? 100 : 12 - e);
        if (e >= -1 && e <= 13) sh = 0;                     // 2^-2 <= value_max < 2^13: no scale (see butterfly_split)
        prm.x_scale = (float)ldexp(1.0, sh);
        prm.out_scale = (float)(p->inv_area * ldexp(1.0, -sh));
    }
    const bool pushing = peers && peers->n > 0;
    if (pushing) {
        if (peers->n > kMaxPeers || out_kind == ZB200_OUT_ABS_PHASE) {
            set_error("project: peer push supports at most %d peers and one output array", kMaxPeers);
            return ZB200_EUNSUP;
        }
        prm.n_peers = peers->n;
        for (int g = 0; g < peers->n; ++g) prm.peer_out[g] = peers->out[g];
    }
    const Knobs& kn = knobs();
    const int cols = f.cols;
    prm.acc_bufs = 2 * cols + 2 * kUnitCols <= (int)kTmemCols ? 2 : 1;
    prm.n_units = ((int)kTmemCols - prm.acc_bufs * cols) / kUnitCols;
    if (prm.n_units > 4) prm.n_units = 4;
    // accumulation chunks: the tensor core's fp32 accumulation truncates, so the accumulators are drained into
    // round-to-nearest register sums every few super-blocks.  Measured (scripts/fold_chunk_probe.py, worst element
    // against the gate / max error / time at 262 144 patches): 31 (no chunking) 1.27-1.72 -- FAILS -- / 5.4e-6;
    // 8: 0.29 / 1.5e-6 / 0.699 ms; 4: 0.21 / 8.6e-7 / 0.699; 2: 0.13 / 5.6e-7 / 0.702 (n_max = 20, one accumulator
    // set: 0.903 / 0.902 / 0.915 ms).
    prm.chunk_sb = f.cfg == 0 ? 2 : 4;
    if (kn.tc_chunk) prm.chunk_sb = kn.tc_chunk;
    prm.park_ns = kn.tc_park == 0 ? 0u : kParkNs;
    const int bst = (cols / 2) * 128;
    const int bar_core = 8 * (2 * 8 + 24) + 16;
    const int tail = pushing ? round_up(bar_core, 128) + 128 + (int)kPushRegion : bar_core;
    prm.b_stages = 3;
    if (kn.tc_bstages) prm.b_stages = kn.tc_bstages;
    prm.n_stages = 3;
    while (prm.n_stages * kXStage + prm.b_stages * bst + tail > kSmemLimit && prm.b_stages > 2) --prm.b_stages;
    while (prm.n_stages * kXStage + prm.b_stages * bst + tail > kSmemLimit && prm.n_stages > 1) --prm.n_stages;
    if (kn.tc_stages && kn.tc_stages < prm.n_stages) prm.n_stages = kn.tc_stages;
    const size_t ring_bytes = (size_t)prm.n_stages * kXStage + (size_t)prm.b_stages * bst;
    const size_t smem = ring_bytes + tail;
    if (smem > (size_t)kSmemLimit) {
        set_error("folded projection: %zu bytes of shared memory needed", smem);
        return ZB200_EUNSUP;
    }
    prm.push_off = (uint32_t)round_up((int)ring_bytes + bar_core, 128);

    CUtensorMap map_x;
    int rc = gsrc ? encode_2d(&map_x, f.fb, (uint64_t)f.n_sb * 32, (uint64_t)cols, (uint64_t)f.n_sb * 128, 8)       // unused
                  : encode_2d(&map_x, d_patches, (uint64_t)p->kk, (uint64_t)n, (uint64_t)p->kk * 4, (uint32_t)kTileRows);
    if (rc) return rc;
    int grid = prm.n_tiles < p->sm_count ? prm.n_tiles : p->sm_count;
    grid = (grid / 2) * 2;
    if (grid < 2) grid = 2;
    if (gsrc) prm.rows_per_cta = ceil_div(n, grid);          // the pusher (round-robin tiles) is never combined with the gather
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int kout = out_kind == ZB200_OUT_ABS ? kOutAbs : (out_kind == ZB200_OUT_ABS_PHASE ? kOutAbsPhase : kOutPlain);
#define ZB_FOLD_LAUNCH(KOUT, CFG, GATHER)                                                                                      \
    if (kout == KOUT && f.cfg == CFG && (gsrc != nullptr) == GATHER) {                                                         \
        ZB_CUDA(cudaFuncSetAttribute(project_fold_kernel<KOUT, CFG, GATHER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        cfg.blockDim = dim3(GATHER ? 640 : 512);                                                                               \
        ZB_CUDA(cudaLaunchKernelEx(&cfg, project_fold_kernel<KOUT, CFG, GATHER>, map_x, f.tmap, prm));                         \
    }
    ZB_FOLD_LAUNCH(kOutPlain, 0, false)
    ZB_FOLD_LAUNCH(kOutAbs, 0, false)
    ZB_FOLD_LAUNCH(kOutAbsPhase, 0, false)
    ZB_FOLD_LAUNCH(kOutPlain, 1, false)
    ZB_FOLD_LAUNCH(kOutAbs, 1, false)
    ZB_FOLD_LAUNCH(kOutAbsPhase, 1, false)
    ZB_FOLD_LAUNCH(kOutPlain, 0, true)
    ZB_FOLD_LAUNCH(kOutAbs, 0, true)
    ZB_FOLD_LAUNCH(kOutAbsPhase, 0, true)
#undef ZB_FOLD_LAUNCH
    ZB_LAUNCHED();
    return ZB200_OK;
}

}  // namespace zb200
