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
//   * operand split  X.B ~= Xhi.Bhi + Xlo.Bhi + Xhi.Blo  where the tensor core itself truncates the
//     raw fp32 X to Xhi and a splitter warpgroup writes Xlo = x - trunc_tf32(x) straight into TMEM
//     (tcgen05.st), from where the MMA reads it as its A operand (no second copy of X in smem);
//   * K-chunked accumulation: the tensor core accumulates in fp32 with round-toward-zero, which
//     biases a 512-step sum by ~1.5e-5 relative (measured); accumulators are therefore drained
//     every 8 k-blocks into fp32 registers of the epilogue warps (round-to-nearest adds), with two
//     accumulator buffers in TMEM so draining overlaps the next chunk's MMAs.
#include "zb200_common.cuh"

#include <cudaTypedefs.h>

namespace zb200 {

namespace tc {

constexpr int kBlockK = 32;                 // floats per k-block = one 128-byte swizzle atom
constexpr int kUmmaK = 8;                   // tf32 MMA K
constexpr int kTileRows = 128;              // UMMA M
constexpr uint32_t kTmemCols = 512;
constexpr int kSmemLimit = 227 * 1024;

constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

struct Params {
    long long n_patches;
    int n_tiles;          // ceil(n_patches / (128*subtiles))
    int subtiles;         // 1 or 2 accumulators (128 patches each) per tile
    int k_blocks;         // k_pad / 32
    int n_pad;            // UMMA N (operand rows, multiple of 16)
    int n_cols;           // meaningful accumulator columns
    int n_stages;
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
    int lo_bufs;          // Xlo staging buffers in TMEM     (tf32x3 kernel)
};

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)),
        "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// D[tmem] (+)= A[smem] . B[smem]^T, both K-major, tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
        "[%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile whose rows are 128 B apart: 8-row groups are 1024 B apart
// (SBO); LBO is unused for swizzled K-major layouts; descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);            // [0,14)  start address >> 4
    d |= (uint64_t)1 << 16;                                 // [16,30) leading byte offset >> 4 (ignored)
    d |= (uint64_t)(1024 >> 4) << 32;                       // [32,46) stride byte offset >> 4
    d |= (uint64_t)1 << 46;                                 // [46,48) descriptor version
    d |= (uint64_t)2 << 61;                                 // [61,64) SWIZZLE_128B
    return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M=128, N=n
__device__ __forceinline__ uint32_t make_idesc_tf32(int n) {
    uint32_t d = 0;
    d |= 1u << 4;                   // c_format = F32
    d |= 2u << 7;                   // a_format = TF32
    d |= 2u << 10;                  // b_format = TF32
    d |= (uint32_t)(n >> 3) << 17;  // N / 8
    d |= (uint32_t)(128 >> 4) << 24;  // M / 16
    return d;
}

__device__ __forceinline__ float tf32_rn(float f) {
    uint32_t u = __float_as_uint(f);
    u += 0x0FFFu + ((u >> 13) & 1u);
    u &= 0xFFFFE000u;
    return __uint_as_float(u);
}

// ---- epilogue: 16 accumulator columns of one patch -----------------------------------------------
// kOut is compile-time so every kernel instance carries exactly one store path.
constexpr int kOutPlain = 0;      // row store of the accumulator columns (REAL and COMPLEX orders)
constexpr int kOutAbs = 1;        // |Zc|
constexpr int kOutAbsPhase = 2;   // |Zc| and angle(Zc)
constexpr int kOutScores = 3;     // fused n-fold scores
constexpr int kFusedFolds = 8;    // fused scores keep at most this many folds in registers

struct ScoreAcc {
    float s1, s2, sm;
    float num[kFusedFolds];
    __device__ __forceinline__ void clear() {
        s1 = s2 = sm = 0.f;
#pragma unroll
        for (int f = 0; f < kFusedFolds; ++f) num[f] = 0.f;
    }
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
                p.out[row * (long long)p.row_len + m0 + i] = sqrtf(re * re + im * im);
                if constexpr (kOut == kOutAbsPhase) p.out2[row * (long long)p.row_len + m0 + i] = atan2f(im, re);
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

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_x);
        prefetch_tmap(&map_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.n_stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
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
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int t = 0; t < my_tiles; ++t) {
                const int tile = blockIdx.x + t * gridDim.x;
                const int row0 = tile * p.subtiles * kTileRows;
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full[s], x_bytes + b_bytes);
                    tma_load_2d(stage_x(s), &map_x, &full[s], kb * kBlockK, row0, kEvictFirst);
                    tma_load_2d(stage_b(s), &map_b, &full[s], kb * kBlockK, 0, kEvictLast);
                    if (++s == p.n_stages) { s = 0; ph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(p.n_pad);
            int s = 0;
            uint32_t ph = 0;
            for (int t = 0; t < my_tiles; ++t) {
                const int buf = t % p.acc_bufs;
                const uint32_t acc_ph = (uint32_t)(t / p.acc_bufs) & 1u;
                mbar_wait(&acc_empty[buf], acc_ph ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t xa = smem_u32(stage_x(s));
                    const uint64_t db = make_desc_sw128(smem_u32(stage_b(s)));
                    for (int sub = 0; sub < p.subtiles; ++sub) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)((buf * p.subtiles + sub) * p.n_pad);
                        const uint64_t dx = make_desc_sw128(xa + sub * kTileRows * 128);
#pragma unroll
                        for (int k4 = 0; k4 < kBlockK / kUmmaK; ++k4) {
                            const uint64_t koff = (uint64_t)((k4 * kUmmaK * 4) >> 4);   // 32 B steps inside the atom
                            umma_tf32(d_tmem, dx + koff, db + koff, idesc, (kb > 0 || k4 > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty[s]);                       // stage reusable once these MMAs retire
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
        }
    }

    tc_fence_before();
    __syncthreads();
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
constexpr int kRegsCtl = 40, kRegsEpi = 192, kRegsSplit = 88;      // (40 + 2*192 + 88) * 128 = 64 Ki

template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// D[tmem] (+)= A[tmem] . B[smem]^T
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int kMaxColChunks = 8;     // 8 x 16 = 128 running sums per epilogue thread

template <int kOut>
__global__ void __launch_bounds__(512, 1)
project_tc3_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_bhi,
                   const __grid_constant__ CUtensorMap map_blo, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t x_bytes = (uint32_t)p.subtiles * kTileRows * 128;
    const uint32_t b_bytes = (uint32_t)p.n_pad * 128;
    const uint32_t stage_bytes = x_bytes + 2 * b_bytes;
    auto stage_x = [&](int s) { return smem + (size_t)s * stage_bytes; };
    auto stage_bhi = [&](int s) { return smem + (size_t)s * stage_bytes + x_bytes; };
    auto stage_blo = [&](int s) { return smem + (size_t)s * stage_bytes + x_bytes + b_bytes; };
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.n_stages * stage_bytes);
    uint64_t* full = bars;                        // TMA landed                               [stages]
    uint64_t* empty = bars + p.n_stages;          // MMAs reading the stage retired           [stages]
    uint64_t* lo_full = bars + 2 * p.n_stages;    // Xlo of a k-block is in TMEM              [2]
    uint64_t* lo_empty = lo_full + 2;             // MMAs reading that Xlo retired            [2]
    uint64_t* acc_full = lo_empty + 2;            // accumulation chunk complete              [2]
    uint64_t* acc_empty = acc_full + 2;           // chunk drained by the epilogue            [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wg = warp >> 2;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_x);
        prefetch_tmap(&map_bhi);
        prefetch_tmap(&map_blo);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.n_stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&lo_full[b], 4);
            mbar_init(&lo_empty[b], 1);
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 8);
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
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lo_base = tmem_base + (uint32_t)(2 * p.subtiles * p.n_pad);      // after the two accumulator sets

    const int my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int n_chunks = (p.k_blocks + p.chunk_kb - 1) / p.chunk_kb;

    if (wg == 0) {
        reg_dec<kRegsCtl>();
        if (warp == 0) {
            // ===================== TMA producer =====================
            if (lane == 0) {
                int s = 0;
                uint32_t ph = 0;
                for (int t = 0; t < my_tiles; ++t) {
                    const int tile = blockIdx.x + t * gridDim.x;
                    const int row0 = tile * p.subtiles * kTileRows;
                    for (int kb = 0; kb < p.k_blocks; ++kb) {
                        mbar_wait(&empty[s], ph ^ 1);
                        mbar_arrive_expect_tx(&full[s], x_bytes + 2 * b_bytes);
                        tma_load_2d(stage_x(s), &map_x, &full[s], kb * kBlockK, row0, kEvictFirst);
                        tma_load_2d(stage_bhi(s), &map_bhi, &full[s], kb * kBlockK, 0, kEvictLast);
                        tma_load_2d(stage_blo(s), &map_blo, &full[s], kb * kBlockK, 0, kEvictLast);
                        if (++s == p.n_stages) { s = 0; ph ^= 1; }
                    }
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            // ===================== MMA issuer =====================
            if (lane == 0) {
                const uint32_t idesc = make_idesc_tf32(p.n_pad);
                int s = 0;
                uint32_t ph = 0;
                uint32_t it = 0;            // running k-block counter  -> Xlo buffer / phase
                uint32_t ck = 0;            // running chunk counter    -> accumulator buffer / phase
                for (int t = 0; t < my_tiles; ++t) {
                    for (int c = 0; c < n_chunks; ++c, ++ck) {
                        const int buf = ck & 1;
                        mbar_wait(&acc_empty[buf], ((ck >> 1) & 1u) ^ 1u);
                        tc_fence_after();
                        const int kb_end = min(p.k_blocks, (c + 1) * p.chunk_kb);
                        for (int kb = c * p.chunk_kb; kb < kb_end; ++kb, ++it) {
                            const int lb = (p.lo_bufs == 2) ? (int)(it & 1u) : 0;
                            const uint32_t lo_ph = (p.lo_bufs == 2) ? ((it >> 1) & 1u) : (it & 1u);
                            mbar_wait(&lo_full[lb], lo_ph);                  // implies full[s]
                            tc_fence_after();
                            const uint32_t xa = smem_u32(stage_x(s));
                            const uint64_t dbh = make_desc_sw128(smem_u32(stage_bhi(s)));
                            const uint64_t dbl = make_desc_sw128(smem_u32(stage_blo(s)));
                            for (int sub = 0; sub < p.subtiles; ++sub) {
                                const uint32_t d_tmem = tmem_base + (uint32_t)((buf * p.subtiles + sub) * p.n_pad);
                                const uint32_t a_lo = lo_base + (uint32_t)((lb * p.subtiles + sub) * kBlockK);
                                const uint64_t dx = make_desc_sw128(xa + sub * kTileRows * 128);
#pragma unroll
                                for (int k4 = 0; k4 < kBlockK / kUmmaK; ++k4) {
                                    const uint64_t koff = (uint64_t)((k4 * kUmmaK * 4) >> 4);
                                    const uint32_t first = (kb == c * p.chunk_kb && k4 == 0) ? 0u : 1u;
                                    umma_tf32(d_tmem, dx + koff, dbh + koff, idesc, first);          // Xhi . Bhi
                                    umma_tf32_ts(d_tmem, a_lo + k4 * kUmmaK, dbh + koff, idesc, 1u);  // Xlo . Bhi
                                    umma_tf32(d_tmem, dx + koff, dbl + koff, idesc, 1u);             // Xhi . Blo
                                }
                            }
                            umma_commit(&empty[s]);
                            umma_commit(&lo_empty[lb]);
                            if (kb == kb_end - 1) umma_commit(&acc_full[buf]);
                            if (++s == p.n_stages) { s = 0; ph ^= 1; }
                        }
                    }
                }
            }
            __syncwarp();
        }
    } else if (wg == 3) {
        // ===================== splitter =====================
        reg_dec<kRegsSplit>();
        const int q = warp & 3;
        const int r = q * 32 + lane;                              // row inside a 128-row sub-tile
        int s = 0;
        uint32_t ph = 0;
        uint32_t it = 0;
        for (int t = 0; t < my_tiles; ++t) {
            for (int kb = 0; kb < p.k_blocks; ++kb, ++it) {
                const int lb = (p.lo_bufs == 2) ? (int)(it & 1u) : 0;
                const uint32_t lo_ph = (p.lo_bufs == 2) ? ((it >> 1) & 1u) : (it & 1u);
                mbar_wait(&full[s], ph);
                mbar_wait(&lo_empty[lb], lo_ph ^ 1u);
                tc_fence_after();
                for (int sub = 0; sub < p.subtiles; ++sub) {
                    const uint8_t* rowp = stage_x(s) + (size_t)sub * kTileRows * 128 + (size_t)r * 128;
                    uint32_t lo[32];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {                 // logical 16-B chunk c sits at c ^ (r % 8)
                        const float4 x = *reinterpret_cast<const float4*>(rowp + ((c ^ (r & 7)) << 4));
                        lo[4 * c + 0] = __float_as_uint(x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u));
                        lo[4 * c + 1] = __float_as_uint(x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u));
                        lo[4 * c + 2] = __float_as_uint(x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u));
                        lo[4 * c + 3] = __float_as_uint(x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u));
                    }
                    tmem_st32(lo_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((lb * p.subtiles + sub) * kBlockK), lo);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&lo_full[lb]);
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
        const int c_beg = (p.subtiles == 2) ? 0 : g * half;
        const int c_end = (p.subtiles == 2) ? p.n_pad : min(p.n_pad, (g + 1) * half);
        const int n_cc = (c_end - c_beg + 15) / 16;               // <= kMaxColChunks
        uint32_t ck = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const int tile = blockIdx.x + t * gridDim.x;
            const long long row = (long long)tile * p.subtiles * kTileRows + sub * kTileRows + q * 32 + lane;
            float sum[kMaxColChunks][16];
#pragma unroll
            for (int cc = 0; cc < kMaxColChunks; ++cc)
#pragma unroll
                for (int i = 0; i < 16; ++i) sum[cc][i] = 0.f;
            for (int c = 0; c < n_chunks; ++c, ++ck) {
                const int buf = ck & 1;
                mbar_wait(&acc_full[buf], (ck >> 1) & 1u);
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
                if (lane == 0) mbar_arrive(&acc_empty[buf]);
            }
            if (row < p.n_patches) {
                ScoreAcc sc;
                if constexpr (kOut == kOutScores) sc.clear();
#pragma unroll
                for (int cc = 0; cc < kMaxColChunks; ++cc) {
                    if (cc < n_cc) {
                        uint32_t v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(sum[cc][i]);
                        epilogue_chunk<kOut>(p, row, c_beg + cc * 16, v, sc);
                    }
                }
                if constexpr (kOut == kOutScores) finish_scores(p, row, sc);   // host guarantees subtiles == 2
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
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

// 1xTF32 needs one UMMA N (<= 256 operand rows); tf32x3 additionally keeps two accumulator sets
// plus the Xlo staging in the 512 TMEM columns: 2*n_pad + 32 <= 512  ->  n_pad <= 240.
static bool operand_ok(const Operand& op, int precision) {
    if (!op.has_tmap) return false;
    return precision == ZB200_PREC_TF32X3 ? op.rows_pad <= 240 : op.rows_pad <= 256;
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
        int rc = tc::encode_2d(&op->tmap_hi, op->hi, (uint64_t)p->k_pad, (uint64_t)op->rows_pad,
                               (uint64_t)p->k_pad * 4, (uint32_t)op->rows_pad);
        if (rc) return rc;
        rc = tc::encode_2d(&op->tmap_lo, op->lo, (uint64_t)p->k_pad, (uint64_t)op->rows_pad, (uint64_t)p->k_pad * 4,
                           (uint32_t)op->rows_pad);
        if (rc) return rc;
        op->has_tmap = true;
    }
    return ZB200_OK;
}

int project_tc(const zb200_plan* p, const float* d_patches, int64_t n, int precision, int out_kind, void* d_out,
               void* d_out2, const float* d_w, const uint8_t* d_sel, int n_folds, int norm_kind, cudaStream_t s) {
    using namespace tc;
    if (n == 0) return ZB200_OK;
    ZB_CHECK_ARG((reinterpret_cast<uintptr_t>(d_patches) & 15) == 0, "project: patch pointer must be 16-byte aligned");
    const bool x3 = precision == ZB200_PREC_TF32X3;
    const bool scores = d_w != nullptr;
    const bool cplx = !scores && out_kind != ZB200_OUT_REAL;
    const Operand& op = cplx ? p->cplx : p->real;
    if (!tc_supported(p, precision, cplx)) {
        set_error("tcgen05 projection (precision %d, %s order) unsupported for n_max=%d size=%d: needs sm_100, even "
                  "size and <= %d operand rows (have %d)", precision, cplx ? "complex" : "real", p->n_max, p->size,
                  x3 ? 240 : 256, op.rows_pad);
        return ZB200_EUNSUP;
    }

    Params prm{};
    prm.n_patches = n;
    prm.n_pad = op.rows_pad;
    prm.n_cols = op.rows;
    prm.k_blocks = p->k_pad / kBlockK;
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

    // tile shape: two 128-patch accumulators per tile when TMEM/smem allow and there is enough work
    // to keep every SM busy with 256-patch tiles
    const int bar_bytes = 1024 + 8 * (2 * 8 + 8) + 16;
    auto stage_bytes = [&](int sub) { return sub * kTileRows * 128 + (x3 ? 2 : 1) * prm.n_pad * 128; };
    int sub = 2;
    if (x3) {
        if (prm.n_pad > 96) sub = 1;                                  // 2 acc sets * 2 * n_pad + 2*64 Xlo <= 512
        if (scores && n_folds > kFusedFolds) {
            set_error("fused n-fold scores support at most %d folds (got %d)", kFusedFolds, n_folds);
            return ZB200_EUNSUP;
        }
        if (scores && sub == 1) {
            set_error("fused n-fold scores in tf32x3 need <= 96 operand rows (have %d); use the unfused path", prm.n_pad);
            return ZB200_EUNSUP;
        }
    } else if (sub * prm.n_pad > (int)kTmemCols) {
        sub = 1;
    }
    if (sub == 2 && !(x3 && scores) && ceil_div(n, 256) < p->sm_count) sub = 1;
    if ((kSmemLimit - bar_bytes) / stage_bytes(sub) < 2 && sub == 2 && !(x3 && scores)) sub = 1;
    prm.subtiles = sub;
    prm.n_stages = (kSmemLimit - bar_bytes) / stage_bytes(sub);
    if (prm.n_stages > 8) prm.n_stages = 8;
    if (prm.n_stages < 1) {
        set_error("project_tc: operand of %d rows does not fit shared memory", prm.n_pad);
        return ZB200_EUNSUP;
    }
    prm.acc_bufs = (2 * sub * prm.n_pad <= (int)kTmemCols) ? 2 : 1;
    prm.lo_bufs = (2 * sub * prm.n_pad + 2 * sub * kBlockK <= (int)kTmemCols) ? 2 : 1;
    prm.n_tiles = (int)ceil_div(n, (int64_t)sub * kTileRows);

    CUtensorMap map_x;
    int rc = encode_2d(&map_x, d_patches, (uint64_t)p->kk, (uint64_t)n, (uint64_t)p->kk * 4, (uint32_t)(sub * kTileRows));
    if (rc) return rc;

    const size_t smem = (size_t)prm.n_stages * stage_bytes(sub) + bar_bytes;
    const int grid = prm.n_tiles < p->sm_count ? prm.n_tiles : p->sm_count;
    const int kout = scores ? kOutScores
                            : (out_kind == ZB200_OUT_ABS ? kOutAbs : (out_kind == ZB200_OUT_ABS_PHASE ? kOutAbsPhase : kOutPlain));
#define ZB_TC_LAUNCH(KOUT)                                                                                            \
    if (kout == KOUT) {                                                                                               \
        if (x3) {                                                                                                     \
            ZB_CUDA(cudaFuncSetAttribute(project_tc3_kernel<KOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                         (int)smem));                                                                 \
            project_tc3_kernel<KOUT><<<grid, 512, smem, s>>>(map_x, op.tmap_hi, op.tmap_lo, prm);                      \
        } else {                                                                                                      \
            ZB_CUDA(cudaFuncSetAttribute(project_tc_kernel<KOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                         (int)smem));                                                                 \
            project_tc_kernel<KOUT><<<grid, 256, smem, s>>>(map_x, op.tmap_hi, prm);                                   \
        }                                                                                                             \
    }
    ZB_TC_LAUNCH(kOutPlain)
    ZB_TC_LAUNCH(kOutAbs)
    ZB_TC_LAUNCH(kOutAbsPhase)
    ZB_TC_LAUNCH(kOutScores)
#undef ZB_TC_LAUNCH
    ZB_LAUNCHED();
    return ZB200_OK;
}

}  // namespace zb200
