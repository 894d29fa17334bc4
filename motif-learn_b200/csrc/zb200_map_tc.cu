// K4 (tensor-core variant) -- dense sliding-window Zernike correlation as an implicit GEMM on
// tcgen05 / TMEM, with the fused n-fold symmetry-score epilogue.
// Replaces ZPs._transform_fft_convolve (mtflearn/features/_zps.py:159-193) and zmoments.rot_maps
// (mtflearn/features/_zmoments.py:420-462):
//   Z[j,y,x] = 1/area * sum_{a,b<k} img0[y-k/2+a, x-k/2+b] * V[j,a,b]        (img0 zero-extended)
//
// GEMM view: for one output row y, M = output pixels, N = modes, K = the k*k taps.  The A operand
// (pixels x taps) is Toeplitz -- A[p, (a,b)] = img0[y-k/2+a, p-k/2+b] -- and is never materialised:
// the window row a of the frame is staged once in shared memory and the UMMA descriptor itself
// walks it.  In the un-swizzled K-major canonical layout a row of the operand is 16 B and
// consecutive rows are 16 B apart, so with the MMA row i standing for pixel x0 + 4i + r the
// element (i, t) sits at  row_copy_r + 16 i + 4 t  bytes: a plain contiguous image row, read with
// leading-byte-offset 16 B (next K chunk) and stride-byte-offset 128 B (next 8 rows).  The four
// pixel phases r = 0..3 need four copies of the row shifted by r floats (16-B aligned start
// addresses).  TMA cannot shift by single floats (inner coordinates must be 16-B aligned), so a
// pre-pass writes the four shifted copies of the frame to HBM and TMA loads rows of those (zero
// fill outside the frame = the reference's zero extension).  B (modes x taps) is the same packed basis
// operand the projection kernel uses (128-B swizzled k-blocks of 32 taps).
//
// Tile = (output row, 512-pixel span, phase pair): two accumulators of 128 pixels x n_pad modes.
// fp32-grade mode: frame split on the fly in a pre-pass into hi = RN_tf32(img), lo = img - hi and
// three MMAs per step (hi.Bhi + lo.Bhi + hi.Blo); accumulators are drained every ~32 K-steps into
// fp32 registers (the tensor core accumulates with round-toward-zero), two accumulator sets in TMEM.
// Warp roles (384 threads): warp 0 frame-row TMA producer, warp 1 basis TMA producer (multicast
// across the cluster), warp 2 MMA issuer, warp 3 TMEM allocator, warpgroups 1-2 epilogue (one
// pixel phase each; thread == pixel, so the score is thread-local).
#include "zb200_common.cuh"
#include "zb200_tc_ptx.cuh"

#include <cudaTypedefs.h>
#include <stdlib.h>

namespace zb200 {
namespace tcmap {

using namespace tc;

constexpr int kSpan = 512;              // pixels per tile span (4 phases x 128 MMA rows)
constexpr int kMaxColChunks = 8;        // running sums per epilogue thread: 8 x 16 columns
constexpr int kRegsCtl = 64, kRegsEpi = 208;     // (64 + 2*208) * 128 < 64 Ki registers

struct MapParams {
    int H, W;
    int row0, rows;
    int k, half;
    int n_pad, n_modes;
    int n_planes;        // 1: frame rounded to tf32;  2: hi/lo split
    int n_terms;         // 1 or 3
    int n_span;
    long long n_tiles;   // rows * n_span * 2
    int rowlen_pad, n_box, bw;
    int img_slots, b_stages;
    int chunk_rows;
    int cluster;
    float* out_moments;
    float* out_scores;
    const float* w;
    const unsigned char* sel;
    int n_folds;
    int norm_kind;
    int dbg;             // ZB200_MAP_DEBUG experiment bits (results wrong when set)
};

// un-swizzled K-major operand whose rows are 16 B apart: LBO (next 16-B K chunk) = 16 B,
// SBO (next group of 8 rows) = 128 B, descriptor version 1
constexpr uint32_t kDescHiToeplitz = (uint32_t)(128 >> 4) | (1u << 14);
__device__ __forceinline__ uint32_t desc_lo_toeplitz(uint32_t smem_addr) { return ((smem_addr & 0x3FFFF) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t desc_toeplitz(uint32_t lo) { return ((uint64_t)kDescHiToeplitz << 32) | lo; }

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)),
        "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(hint)
        : "memory");
}

__device__ __forceinline__ float tf32_rn(float f) {
    uint32_t u = __float_as_uint(f);
    u += 0x0FFFu + ((u >> 13) & 1u);
    u &= 0xFFFFE000u;
    return __uint_as_float(u);
}

// frame -> operand planes [n_planes][4 shifts][H][Wp] (Wp = W rounded up to 4):
//   plane(pl, r)[y][u] = v_pl(img[y][u - 4 + r]),  v_0 = RN_tf32, v_1 = RN_tf32(img - v_0), zero outside the row
// (four columns of left padding keep the r pixels that the shift moves across the left frame edge).
// TMA needs 16-byte aligned inner coordinates, so the four pixel-phase shifts cannot be expressed as
// element offsets of one plane; they are materialised here (a few hundred MB/s of extra HBM writes
// against a multi-TFLOP contraction).
__global__ void map_prepare_kernel(const float* __restrict__ img, int H, int W, int Wp, int n_planes,
                                   float* __restrict__ planes) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long plane_elems = (long long)H * Wp;
    if (i >= plane_elems) return;
    const int y = (int)(i / Wp), u = (int)(i - (long long)y * Wp);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int x = u - 4 + r;
        const float v = (x >= 0 && x < W) ? __ldg(img + (long long)y * W + x) : 0.f;
        const float hi = tf32_rn(v);
        planes[(long long)r * plane_elems + i] = hi;
        if (n_planes == 2) planes[(long long)(4 + r) * plane_elems + i] = tf32_rn(v - hi);
    }
}

template <bool kScores>
__global__ void __launch_bounds__(384, 1)
map_tc_kernel(const __grid_constant__ CUtensorMap map_img, const __grid_constant__ CUtensorMap map_bhi,
              const __grid_constant__ CUtensorMap map_blo, const MapParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int n_bops = p.n_terms == 3 ? 2 : 1;
    const uint32_t b_bytes = (uint32_t)p.n_pad * 128;                 // one basis operand k-block
    const uint32_t b_stage = (uint32_t)n_bops * b_bytes;
    const uint32_t copy_bytes = (uint32_t)p.rowlen_pad * 4;           // one shifted copy of a frame row
    const uint32_t slot_bytes = 2u * p.n_planes * copy_bytes;         // [phase g][plane t]
    uint8_t* b_ring = smem;
    uint8_t* img_ring = smem + (size_t)p.b_stages * b_stage;
    uint64_t* bars = reinterpret_cast<uint64_t*>(img_ring + (((size_t)p.img_slots * slot_bytes + 15) & ~(size_t)15));
    uint64_t* img_full = bars;
    uint64_t* img_empty = img_full + p.img_slots;
    uint64_t* b_full = img_empty + p.img_slots;
    uint64_t* b_empty = b_full + p.b_stages;
    uint64_t* acc_full = b_empty + p.b_stages;      // [2]
    uint64_t* acc_empty = acc_full + 2;             // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wg = warp >> 2;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_img);
        prefetch_tmap(&map_bhi);
        prefetch_tmap(&map_blo);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.img_slots; ++s) {
            mbar_init(&img_full[s], 1);
            mbar_init(&img_empty[s], 1);
        }
        for (int s = 0; s < p.b_stages; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], p.cluster);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 8);
        }
        fence_barrier_init();
    }
    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    if (p.cluster > 1) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t crank = p.cluster > 1 ? cluster_rank() : 0u;
    const uint16_t cmask = (uint16_t)((1u << p.cluster) - 1u);
    const long long my_tiles = (p.n_tiles + gridDim.x - 1) / gridDim.x;      // lock-step within the cluster
    const int k = p.k;
    const int steps_per_row = k >> 3;                                          // 8-tap K steps per window row
    const int k_blocks = (k * k) >> 5;                                         // 32-tap basis k-blocks per tile
    const int n_chunks = (k + p.chunk_rows - 1) / p.chunk_rows;

    auto decode = [&](long long tile, int& yl, int& span, int& pp) {
        pp = (int)(tile & 1);
        const long long rest = tile >> 1;
        span = (int)(rest % p.n_span);
        yl = (int)(rest / p.n_span);
    };

    if (wg == 0) {
        reg_dec<kRegsCtl>();
        if (warp == 0) {
            // ===================== frame-row producer =====================
            if (elect_one()) {
                int s = 0;
                uint32_t ph = 0;
                for (long long t = 0; t < my_tiles; ++t) {
                    const long long tile = blockIdx.x + t * gridDim.x;
                    int yl, span, pp;
                    decode(tile, yl, span, pp);
                    const bool live = tile < p.n_tiles;
                    const int y = p.row0 + yl;
                    const int xs = span * kSpan - p.half + 4;                 // plane column of copy element 0 (multiple of 4)
                    for (int a = 0; a < k; ++a) {
                        mbar_wait(&img_empty[s], ph ^ 1);
                        if (ZB200_DEBUG_HOOKS && (p.dbg & 1)) {
                            mbar_arrive(&img_full[s]);
                            if (++s == p.img_slots) { s = 0; ph ^= 1; }
                            continue;
                        }
                        mbar_arrive_expect_tx(&img_full[s], slot_bytes);
                        uint8_t* slot = img_ring + (size_t)s * slot_bytes;
                        // dead tiles (past the end, cluster padding) read far outside the frame: all zeros
                        const int yy = live ? y - p.half + a : -4 * k;
                        for (int g = 0; g < 2; ++g)
                            for (int pl = 0; pl < p.n_planes; ++pl)
                                for (int bx = 0; bx < p.n_box; ++bx)
                                    tma_load_3d(slot + (size_t)(g * p.n_planes + pl) * copy_bytes + (size_t)bx * p.bw * 4, &map_img,
                                                &img_full[s], xs + bx * p.bw, yy, pl * 4 + 2 * pp + g, kEvictLast);
                        if (++s == p.img_slots) { s = 0; ph ^= 1; }
                    }
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            // ===================== basis producer =====================
            if (elect_one()) {
                const int b_rows = p.n_pad / p.cluster;
                int s = 0;
                uint32_t ph = 0;
                for (long long t = 0; t < my_tiles; ++t) {
                    for (int kb = 0; kb < k_blocks; ++kb) {
                        mbar_wait(&b_empty[s], ph ^ 1);
                        mbar_arrive_expect_tx(&b_full[s], b_stage);
                        uint8_t* st = b_ring + (size_t)s * b_stage;
                        if (p.cluster == 1) {
                            tma_load_2d(st, &map_bhi, &b_full[s], kb * kBlockK, 0, kEvictLast);
                            if (n_bops == 2) tma_load_2d(st + b_bytes, &map_blo, &b_full[s], kb * kBlockK, 0, kEvictLast);
                        } else {
                            const size_t off = (size_t)crank * b_rows * 128;
                            tma_load_2d_mc(st + off, &map_bhi, &b_full[s], kb * kBlockK, (int)crank * b_rows, cmask, kEvictLast);
                            if (n_bops == 2)
                                tma_load_2d_mc(st + b_bytes + off, &map_blo, &b_full[s], kb * kBlockK, (int)crank * b_rows, cmask,
                                               kEvictLast);
                        }
                        if (++s == p.b_stages) { s = 0; ph ^= 1; }
                    }
                }
            }
            __syncwarp();
        } else if (warp == 2) {
            // ===================== MMA issuer =====================
            if (elect_one()) {
                const uint32_t idesc = make_idesc_tf32(p.n_pad);
                const uint32_t img_lo0 = desc_lo_toeplitz(smem_u32(img_ring));
                const uint32_t b_lo0 = desc_lo_sw128(smem_u32(b_ring));
                const uint32_t slot_step = slot_bytes >> 4, copy_step = copy_bytes >> 4;
                const uint32_t b_step = b_stage >> 4, bop_step = b_bytes >> 4;
                const bool x3 = p.n_terms == 3;
                int si = 0, sb = 0;
                uint32_t phi = 0, phb = 0;
                uint32_t ck = 0;
                for (long long t = 0; t < my_tiles; ++t) {
                    uint32_t step = 0;                           // 8-tap K step inside the tile: 0 .. k*k/8
                    uint32_t bhl = 0, bll = 0;
                    for (int c = 0; c < n_chunks; ++c, ++ck) {
                        const int buf = ck & 1;
                        mbar_wait(&acc_empty[buf], ((ck >> 1) & 1u) ^ 1u);
                        tc_fence_after();
                        const uint32_t d0 = tmem_base + (uint32_t)(buf * 2 * p.n_pad);
                        const uint32_t d1 = d0 + (uint32_t)p.n_pad;
                        const int a_begin = c * p.chunk_rows, a_end = min(k, a_begin + p.chunk_rows);
                        for (int a = a_begin; a < a_end; ++a) {
                            mbar_wait(&img_full[si], phi);
                            tc_fence_after();
                            // copies of this window row: [phase g][plane]; plane 1 (lo) directly after plane 0
                            const uint32_t a00 = img_lo0 + (uint32_t)si * slot_step;          // g=0 hi
                            const uint32_t a10 = a00 + (uint32_t)p.n_planes * copy_step;       // g=1 hi
                            for (int b8 = 0; b8 < steps_per_row; ++b8, ++step) {
                                const uint32_t k4 = step & 3u;
                                if (k4 == 0) {
                                    mbar_wait(&b_full[sb], phb);
                                    tc_fence_after();
                                    bhl = b_lo0 + (uint32_t)sb * b_step;
                                    bll = bhl + bop_step;
                                }
                                const uint32_t ao = 2u * (uint32_t)b8;                      // 8 taps = 32 B
                                const uint32_t bo = 2u * k4;
                                const uint32_t acc = (a > a_begin || b8 > 0) ? 1u : 0u;
                                if (!(ZB200_DEBUG_HOOKS && (p.dbg & 2))) {
                                umma_tf32(d0, desc_toeplitz(a00 + ao), desc_from_lo(bhl + bo), idesc, acc);
                                if (x3) {
                                    umma_tf32(d0, desc_toeplitz(a00 + copy_step + ao), desc_from_lo(bhl + bo), idesc, 1u);
                                    umma_tf32(d0, desc_toeplitz(a00 + ao), desc_from_lo(bll + bo), idesc, 1u);
                                }
                                umma_tf32(d1, desc_toeplitz(a10 + ao), desc_from_lo(bhl + bo), idesc, acc);
                                if (x3) {
                                    umma_tf32(d1, desc_toeplitz(a10 + copy_step + ao), desc_from_lo(bhl + bo), idesc, 1u);
                                    umma_tf32(d1, desc_toeplitz(a10 + ao), desc_from_lo(bll + bo), idesc, 1u);
                                }
                                }
                                if (k4 == 3) {
                                    if (p.cluster == 1) umma_commit(&b_empty[sb]);
                                    else umma_commit_mc(&b_empty[sb], cmask);
                                    if (++sb == p.b_stages) { sb = 0; phb ^= 1; }
                                }
                            }
                            umma_commit(&img_empty[si]);
                            if (++si == p.img_slots) { si = 0; phi ^= 1; }
                        }
                        umma_commit(&acc_full[buf]);
                    }
                }
            }
            __syncwarp();
        }
    } else {
        // ===================== epilogue: warpgroup 1 -> phase 2pp, warpgroup 2 -> phase 2pp+1 =====================
        reg_inc<kRegsEpi>();
        const int g = wg - 1;
        const int q = warp & 3;
        const int n_cc = p.n_pad >> 4;
        uint32_t ck = 0;
        for (long long t = 0; t < my_tiles; ++t) {
            const long long tile = blockIdx.x + t * gridDim.x;
            int yl, span, pp;
            decode(tile, yl, span, pp);
            float sum[kMaxColChunks][16];
#pragma unroll
            for (int cc = 0; cc < kMaxColChunks; ++cc)
#pragma unroll
                for (int i = 0; i < 16; ++i) sum[cc][i] = 0.f;
            for (int c = 0; c < n_chunks; ++c, ++ck) {
                const int buf = ck & 1;
                mbar_wait(&acc_full[buf], (ck >> 1) & 1u);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * 2 + g) * p.n_pad);
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
            const int x = span * kSpan + 4 * (q * 32 + lane) + 2 * pp + g;
            if (tile < p.n_tiles && x < p.W) {
                if (kScores) {
                    // pass 1: norms over the selected modes; the sums are squared in place
                    float s1 = 0.f, s2 = 0.f, sm = 0.f;
#pragma unroll
                    for (int cc = 0; cc < kMaxColChunks; ++cc) {
                        if (cc < n_cc) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const int col = cc * 16 + i;
                                const bool on = col < p.n_modes && p.sel[col];
                                const float z = on ? sum[cc][i] : 0.f;
                                const float az = fabsf(z);
                                s1 += az;
                                s2 = fmaf(z, z, s2);
                                sm = fmaxf(sm, az);
                                sum[cc][i] = z * z;
                            }
                        }
                    }
                    float den = 1.f;
                    if (p.norm_kind == ZB200_NORM_L1) den = s1 * s1;
                    else if (p.norm_kind == ZB200_NORM_L2) den = s2;
                    else if (p.norm_kind == ZB200_NORM_INF) den = sm * sm;
                    // pass 2: one weighted sum per fold (weights are warp-uniform loads)
#pragma unroll 1
                    for (int f = 0; f < p.n_folds; ++f) {
                        const float* wf = p.w + (size_t)f * p.n_pad;
                        float num = 0.f;
#pragma unroll
                        for (int cc = 0; cc < kMaxColChunks; ++cc) {
                            if (cc < n_cc) {
#pragma unroll
                                for (int i = 0; i < 16; ++i) num = fmaf(__ldg(wf + cc * 16 + i), sum[cc][i], num);
                            }
                        }
                        p.out_scores[((size_t)f * p.rows + yl) * p.W + x] = num / den;
                    }
                } else {
#pragma unroll
                    for (int cc = 0; cc < kMaxColChunks; ++cc) {
                        if (cc < n_cc) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const int col = cc * 16 + i;
                                if (col < p.n_modes) p.out_moments[((size_t)col * p.rows + yl) * p.W + x] = sum[cc][i];
                            }
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (p.cluster > 1) cluster_sync_all();
    if (warp == 3) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
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

}  // namespace tcmap

bool map_tc_supported(const zb200_plan* p, int precision) {
    if (precision != ZB200_PREC_TF32 && precision != ZB200_PREC_TF32X3) return false;
    // sm_100, window a multiple of 8 taps (an MMA K step never straddles a window row), basis operand
    // with TMA descriptors, two accumulator sets x two phases in 512 TMEM columns
    return p->cc_major == 10 && p->size % 8 == 0 && p->real.has_tmap && p->real.rows_pad <= 128;
}

int map_tc(const zb200_plan* p, const float* d_img, int H, int W, int row0, int rows, int precision,
           float* d_moments, float* d_scores, const float* d_w, const uint8_t* d_sel, int n_folds, int norm_kind,
           cudaStream_t s) {
    using namespace tcmap;
    if (rows == 0) return ZB200_OK;
    if (!map_tc_supported(p, precision)) {
        set_error("tcgen05 dense map unsupported for n_max=%d size=%d (needs sm_100, size %% 8 == 0, <= 128 operand rows)",
                  p->n_max, p->size);
        return ZB200_EUNSUP;
    }
    const Operand& op = p->real;
    MapParams prm{};
    prm.H = H; prm.W = W; prm.row0 = row0; prm.rows = rows;
    prm.k = p->size; prm.half = p->size / 2;
    prm.n_pad = op.rows_pad; prm.n_modes = p->n_modes;
    prm.n_planes = precision == ZB200_PREC_TF32X3 ? 2 : 1;
    prm.n_terms = precision == ZB200_PREC_TF32X3 ? 3 : 1;
    prm.n_span = (int)ceil_div(W, kSpan);
    prm.n_tiles = (long long)rows * prm.n_span * 2;
    const int rowlen = 4 * 127 + prm.k + 4;                     // last element touched: 4*127 + (k-1)
    prm.n_box = (int)ceil_div(rowlen, 256);
    prm.bw = round_up((int)ceil_div(rowlen, prm.n_box), 32);      // TMA destinations must be 128-B aligned
    prm.rowlen_pad = prm.n_box * prm.bw;
    prm.chunk_rows = 32 / (prm.k / 8) > 0 ? 32 / (prm.k / 8) : 1;     // drain after ~32 accumulation steps
    prm.out_moments = d_moments; prm.out_scores = d_scores; prm.w = d_w; prm.sel = d_sel;
    prm.n_folds = n_folds; prm.norm_kind = norm_kind;
    prm.dbg = knobs().map_debug;

    // operand planes of the frame: (hi | lo) x 4 pixel-phase shifts, pitch padded to 16 B
    const int Wp = round_up(W + 4, 4);
    const int n_maps = 4 * prm.n_planes;
    float* planes = nullptr;
    ZB_CUDA(scratch_alloc(&planes, sizeof(float) * (size_t)n_maps * H * Wp, s));
    {
        const long long n = (long long)H * Wp;
        map_prepare_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(d_img, H, W, Wp, prm.n_planes, planes);
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    // 3-D TMA descriptor [plane][row][col], box = bw floats x 1 row x 1 plane, no swizzle, zero fill outside
    CUtensorMap map_img;
    {
        auto enc = get_encode();
        if (!enc) { cudaFreeAsync(planes, s); set_error("cuTensorMapEncodeTiled is not available"); return ZB200_ECUDA; }
        cuuint64_t dims[3] = {(cuuint64_t)Wp, (cuuint64_t)H, (cuuint64_t)n_maps};
        cuuint64_t strides[2] = {(cuuint64_t)Wp * 4, (cuuint64_t)Wp * 4 * H};
        cuuint32_t box[3] = {(cuuint32_t)prm.bw, 1, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&map_img, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, planes, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            cudaFreeAsync(planes, s);
            set_error("cuTensorMapEncodeTiled(frame) failed with CUresult %d", (int)r);
            return ZB200_ECUDA;
        }
    }

    int cluster = 2;
    if (knobs().tc_cluster) cluster = knobs().tc_cluster;
    while (cluster > 1 && (cluster > op.max_cluster || prm.n_tiles < 2 * cluster)) cluster >>= 1;
    prm.cluster = cluster;
    const int lg = cluster == 4 ? 2 : (cluster == 2 ? 1 : 0);

    const int b_stage = (prm.n_terms == 3 ? 2 : 1) * prm.n_pad * 128;
    const int slot_bytes = 2 * prm.n_planes * prm.rowlen_pad * 4;
    prm.b_stages = 4;
    prm.img_slots = 8;
    const size_t smem = 1024 + (size_t)prm.b_stages * b_stage + (((size_t)prm.img_slots * slot_bytes + 15) & ~(size_t)15) +
                        8 * (2 * prm.img_slots + 2 * prm.b_stages + 4) + 16;
    if (smem > (size_t)kSmemLimit) {
        cudaFreeAsync(planes, s);
        set_error("tcgen05 dense map: window %d needs %zu B of shared memory", prm.k, smem);
        return ZB200_EUNSUP;
    }
    long long grid = prm.n_tiles < p->sm_count ? prm.n_tiles : p->sm_count;
    grid = (grid / cluster) * cluster;
    if (grid < cluster) grid = cluster;

    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e;
    if (d_scores) {
        e = cudaFuncSetAttribute(map_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, map_tc_kernel<true>, map_img, op.tmap_hi[lg], op.tmap_lo_f32[lg], prm);
    } else {
        e = cudaFuncSetAttribute(map_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, map_tc_kernel<false>, map_img, op.tmap_hi[lg], op.tmap_lo_f32[lg], prm);
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaFreeAsync(planes, s);
    if (e != cudaSuccess) {
        set_error("map_tc_kernel launch failed: %s", cudaGetErrorString(e));
        return ZB200_ECUDA;
    }
    ZB_CUDA(cudaGetLastError());
    return ZB200_OK;
}

}  // namespace zb200
