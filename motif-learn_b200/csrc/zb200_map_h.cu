// K4 (fp16-split tensor-core variant) -- dense sliding-window Zernike correlation as an implicit GEMM
// on tcgen05 / TMEM with the fused n-fold symmetry-score epilogue.
// Replaces ZPs._transform_fft_convolve (mtflearn/features/_zps.py:159-193) and zmoments.rot_maps
// (mtflearn/features/_zmoments.py:420-462):
//   Z[j,y,x] = 1/area * sum_{a,b<k} img0[y-k/2+a, x-k/2+b] * V[j,a,b]        (img0 zero-extended)
//
// Same GEMM view as zb200_map_tc.cu (M = output pixels of one row, N = modes, K = window taps, the
// Toeplitz A operand never materialised), but the arithmetic is kind::f16 instead of kind::tf32:
//   frame  x' = x * 2^-e  (e from max|x|, so |x'| < 2^14),   x1 = RN_f16(x'), x2 = RN_f16(x' - x1)
//   basis  b = V,                                             b1 = RN_f16(b),  b2 = RN_f16(b - b1)
//   Z = 2^e/area * sum (x1.b1 + x2.b1 + x1.b2)                        -> 22-bit operands, fp32-grade
// A K=16 f16 MMA costs what a K=8 tf32 MMA costs, so the three terms cost 1.5 tf32 passes instead of 3.
//
// Operand geometry.  kind::f16 core matrices are 8 rows x 16 B = 8 taps, and in the un-swizzled K-major
// canonical layout rows of one core matrix are 16 B apart, so with MMA row i standing for pixel
// x0 + 4i + r the pre-pass writes, per pixel phase r and part (x1|x2), the "overlapped" row
//   E_r[q][t] = part(img[y][4q + t + r - pad]),  t = 0..7          (16 B per q; 2x duplication)
// and element (i, tap t) of the 16-tap group g of window row a sits at  E_r + 16 (q0 + i + 4g) + 2t,
// taps 8..15 at +32 B (leading byte offset 32 B, stride byte offset 128 B).  TMA loads rows of E
// (zero fill outside the frame = the reference's zero extension).
// Window rows that lie entirely outside the unit disk are never issued; the packed B operand (b1 | b2, 128-B
// swizzled k-blocks of 4 groups, ceil(G/4) k-blocks per window row) holds the remaining rows in issue order.
//
// Two output rows per MMA.  A staged frame row f is window row a = f - y0 + k/2 of output row y0 and window row
// a - 1 of output row y0 + 1, so ONE MMA with N = 2 n_pad whose B operand is the basis of window rows (a-1, a) --
// adjacent slots of the basis ring, slot S mirroring slot 0 for the wrap-around -- feeds both rows: accumulator
// columns [0, n_pad) belong to row y0 + 1, [n_pad, 2 n_pad) to row y0.  An N = 96 SS MMA is shared-memory-bound
// (7 KB per 48 tensor clk), the N = 192 one is tensor-bound.  Rows are paired by ABSOLUTE parity so that row bands
// computed separately (image-tile sharding) reproduce the single-call result bit for bit.
//
// Tile = (row pair, 512-pixel span, pixel phase r): 128 pixels x 2 rows, two accumulator sets of 2 n_pad TMEM
// columns, drained every ~12 groups per row into fp32 registers (the tensor core accumulates with
// round-toward-zero).  Warp roles (384 threads): warps 0-7 epilogue (thread == pixel, so the score is
// thread-local; warpgroup 0 = row y0+1, warpgroup 1 = row y0), warp 8 frame-row TMA producer, warp 9 basis TMA
// producer (multicast across the cluster), warp 10 MMA issuer (loop specialised on G and unrolled per step: the
// first version was issue-bound, profiles/r01_maph_issue_study.md), warp 11 TMEM allocator.
#include "zb200_common.cuh"
#include "zb200_tc_ptx.cuh"

#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <vector>

namespace zb200 {
namespace hmap {

using namespace tc;

constexpr int kSpan = 512;              // pixels per tile span (4 phases x 128 MMA rows)
constexpr int kMaxColChunks = 8;        // running sums per epilogue thread: 8 x 16 columns
constexpr int kRegsCtl = 64, kRegsEpi = 208;     // (64 + 2*208) * 128 < 64 Ki registers
constexpr int kMaxWindow = kMapHalfMaxWindow;
constexpr int kFoldsPerLaunch = 8;

struct MapParams {
    int H, W;
    int row0, rows;
    int plane_row0;      // first frame row held by the operand planes (they cover only the rows this band reads)
    int k, half, pad;
    int n_pad, n_modes;
    int n_terms;         // 1: x1.b1 only (11-bit operands);  3: fp32-grade split
    int n_span;
    long long n_tiles;   // n_pairs * n_span * 4
    int pair0, n_pairs;  // absolute output-row pairs (2j, 2j+1) that overlap [row0, row0+rows)
    int n_groups;        // 16-tap groups per window row (G)
    int kb_per_row;      // basis k-blocks (4 groups) per window row = ceil(G/4)
    int a_first, n_rows; // window rows with at least one tap inside the disk: [a_first, a_first+n_rows)
    int chunk_steps;     // window-row steps per accumulator chunk
    int n_chunks;
    int copy_q, n_box, bw_q;   // one staged row copy: n_box TMA boxes of bw_q 16-byte units
    int img_slots, b_slots;
    int cluster;
    float* out_moments;
    float* out_scores;
    const float* scale;  // [1]: 2^e / area, written by the pre-pass
    int n_folds;
    int norm_kind;
    int dbg;             // ZB200_MAP_DEBUG experiment bits (results wrong when set)
    // score tables live in the kernel parameters: with the column loops unrolled every weight is a
    // constant-bank operand of its FFMA (no loads in the epilogue)
    unsigned char step_mask[kMaxWindow + 1];   // per window-row step: bit g = 16-tap group g has taps inside the disk in
                                      //   window row a_first+s or a_first+s-1 (all-zero basis otherwise: MMAs skipped)
    float selw[128];                  // 1 for modes that enter the norm (unselect()), else 0
    float wts[kFoldsPerLaunch][128];  // construct_rot_maps_matrix rows, zero on unselected / padding modes
};

// un-swizzled K-major operand whose rows are 16 B apart: SBO (next group of 8 rows) = 128 B, LBO (next
// 8-tap core matrix along K) = 32 B, descriptor version 1
constexpr uint32_t kDescHiToeplitz = (uint32_t)(128 >> 4) | (1u << 14);
__device__ __forceinline__ uint32_t desc_lo_toeplitz(uint32_t smem_addr) { return ((smem_addr & 0x3FFFF) >> 4) | (2u << 16); }
__device__ __forceinline__ uint64_t desc_toeplitz(uint32_t lo) { return ((uint64_t)kDescHiToeplitz << 32) | lo; }

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)),
        "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(hint)
        : "memory");
}

// kind::f16 with fp16 inputs, fp32 accumulate, A and B K-major, M=128, N=n
__device__ __forceinline__ uint32_t make_idesc_f16(int n) {
    uint32_t d = 0;
    d |= 1u << 4;                     // c_format = F32;  a_format = b_format = 0 (F16)
    d |= (uint32_t)(n >> 3) << 17;    // N / 8
    d |= (uint32_t)(128 >> 4) << 24;  // M / 16
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ---- pre-pass 1: max |img| (bit pattern; non-negative floats order like their bits) ----------------
__global__ void map_absmax_kernel(const float* __restrict__ img, long long n, unsigned int* __restrict__ out_bits) {
    unsigned int m = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = max(m, __float_as_uint(__ldg(img + i)) & 0x7FFFFFFFu);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out_bits, m);
}

// biased exponent of the frame maximum, clamped so that both 2^(140-E) and 2^(E-140) are normal floats
__device__ __forceinline__ int frame_exponent(unsigned int absmax_bits) {
    int E = (int)((absmax_bits >> 23) & 0xFFu);
    if (absmax_bits == 0) E = 127;
    return min(max(E, 16), 254);
}

// ---- pre-pass 2: frame -> overlapped fp16 operand planes [part][phase r][H][Wq][8] --------------------
//   plane(part, r)[y][q][t] = part(img[y][4q + t + r - pad] * 2^(140-E)),   zero outside the row
// also writes scale[0] = 2^(E-140) / area for the epilogue.
__global__ void map_prepare_kernel(const float* __restrict__ img, int y_first, int n_rows, int W, int Wq, int pad, int n_parts,
                                   const unsigned int* __restrict__ absmax_bits, double inv_area,
                                   uint4* __restrict__ planes, float* __restrict__ scale) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long plane_units = (long long)n_rows * Wq;
    const int E = frame_exponent(*absmax_bits);
    if (i == 0) *scale = (float)ldexp(inv_area, E - 140);
    if (i >= plane_units) return;
    const float mult = __uint_as_float((uint32_t)(267 - E) << 23);       // 2^(140-E)
    const int y = (int)(i / Wq), q = (int)(i - (long long)y * Wq);
    const float* row = img + (long long)(y_first + y) * W;
    const int xb = 4 * q - pad;
    __half h1[11], h2[11];
#pragma unroll
    for (int j = 0; j < 11; ++j) {
        const int x = xb + j;
        const float v = (x >= 0 && x < W) ? __ldg(row + x) * mult : 0.f;
        h1[j] = __float2half_rn(v);
        h2[j] = __float2half_rn(v - __half2float(h1[j]));
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        uint32_t w1[4], w2[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            w1[t] = (uint32_t)__half_as_ushort(h1[r + 2 * t]) | ((uint32_t)__half_as_ushort(h1[r + 2 * t + 1]) << 16);
            w2[t] = (uint32_t)__half_as_ushort(h2[r + 2 * t]) | ((uint32_t)__half_as_ushort(h2[r + 2 * t + 1]) << 16);
        }
        planes[(long long)r * plane_units + i] = make_uint4(w1[0], w1[1], w1[2], w1[3]);
        if (n_parts == 2) planes[(long long)(4 + r) * plane_units + i] = make_uint4(w2[0], w2[1], w2[2], w2[3]);
    }
}

// ---- plan-time: packed fp16 basis operands, active 16-tap groups only ---------------------------------
//   b1[r][gi*16 + t] = RN_f16(V[r][a][16 g + t]),  b2 = RN_f16(V - b1);  (a, g) = groups[gi];  zero for taps
//   beyond the window row, padding rows and padding groups.  Row pitch = n_kb * 64 halves (128 B per k-block).
__global__ void map_pack_basis_kernel(const double* __restrict__ basis, int mode0, int n_modes, int rows_pad, int k,
                                      const unsigned short* __restrict__ groups, int n_act, int n_kb,
                                      __half* __restrict__ b1, __half* __restrict__ b2) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;       // gi * 16 + t
    const int r = blockIdx.y;
    if (col >= n_kb * 64) return;
    const int gi = col >> 4, t = col & 15;
    double v = 0.0;
    if (r < n_modes && gi < n_act && groups[gi] != 0xFFFFu) {
        const int a = groups[gi] >> 8, g = groups[gi] & 0xFF;
        const int b = 16 * g + t;
        if (b < k) v = basis[((size_t)(mode0 + r) * k + a) * k + b];
    }
    const __half h = __double2half(v);
    const size_t o = (size_t)r * n_kb * 64 + col;
    b1[o] = h;
    b2[o] = __double2half(v - (double)__half2float(h));
}

struct IssueCtx {
    uint32_t idesc1, idesc2;          // N = n_pad (one output row) / N = 2 n_pad (both rows of the pair)
    uint32_t slot_step, copy_step, b_slot_step, b_ring_step, img_lo0, b_lo0, tmem_base;
    uint16_t cmask;
    long long my_tiles;
    uint64_t *img_full, *img_empty, *b_full, *b_empty, *acc_full, *acc_empty;
};

// MMAs of one window-row step.  kMode: 0 = both output rows (N = 2 n_pad, B = the contiguous slot pair),
// 1 = upper row only (first step: accumulator columns [n_pad, 2 n_pad), N = n_pad), 2 = lower row only (last step).
// a0: descriptor word of the staged frame row (part x1; part x2 follows at +copy_step); bb: basis slot of the
// first window row the step reads (mode 0: row a-1, row a follows at +b_slot_step); operand rings (k-block,
// b1|b2) are b_ring_step apart.
template <int G, bool kX3, int kMode>
__device__ __forceinline__ void issue_step(const IssueCtx& c, uint32_t d, int n_pad, uint32_t a0, uint32_t bb,
                                           uint32_t acc_first, bool split_first, uint32_t mask) {
    constexpr int n_bops = kX3 ? 2 : 1;
    const uint32_t dd = kMode == 1 ? d + (uint32_t)n_pad : d;
    const uint32_t idesc = kMode == 0 ? c.idesc2 : c.idesc1;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        if (!((mask >> g) & 1u)) continue;                                          // group outside the disk in both rows
        const uint32_t ag = a0 + 4u * g;                                            // 16 taps = 4 units of 16 B
        const uint32_t bg = bb + (uint32_t)((g / 4) * n_bops) * c.b_ring_step + 2u * (uint32_t)(g % 4);   // 32 B per group
        if (g == 0) {
            if (kMode == 0 && split_first) {
                // first step that touches the lower row while the upper one already holds a partial sum
                umma_f16(d, desc_toeplitz(ag), desc_from_lo(bg), c.idesc1, 0u);
                umma_f16(d + (uint32_t)n_pad, desc_toeplitz(ag), desc_from_lo(bg + c.b_slot_step), c.idesc1, 1u);
            } else {
                umma_f16(dd, desc_toeplitz(ag), desc_from_lo(bg), idesc, acc_first);
            }
        } else {
            umma_f16(dd, desc_toeplitz(ag), desc_from_lo(bg), idesc, 1u);
        }
        if (kX3) {
            umma_f16(dd, desc_toeplitz(ag + c.copy_step), desc_from_lo(bg), idesc, 1u);
            umma_f16(dd, desc_toeplitz(ag), desc_from_lo(bg + c.b_ring_step), idesc, 1u);
        }
    }
}

// The MMA schedule of the issuer thread, specialised on G = 16-tap groups per window row.  One tile = one pair
// of output rows (y0, y0+1), 128 pixels of one phase; step s stages frame row y0 - k/2 + a_first + s, which is
// window row a = a_first + s of output row y0 and window row a - 1 of output row y0 + 1: ONE MMA with
// N = 2 n_pad whose B operand is the basis of window rows (a-1, a) -- adjacent slots of the basis ring -- feeds
// both rows.  Against two N = n_pad MMAs this halves the A-side shared-memory reads (the N = 96 SS MMA is
// shared-memory-bound: 7 KB per 48 tensor clk) and the number of instructions the issuing thread must retire.
template <int G, bool kX3, bool kCluster2>
__device__ __forceinline__ void issue_tiles(const MapParams& p, const IssueCtx& c) {
    const int S = p.b_slots;
    int si = 0;
    uint32_t phi = 0;
    int sj = 0;                 // basis slot of the window row this step waits for
    uint32_t pj = 0;
    int sprev = 0;              // slot of the previous window row (released after this step)
    uint32_t ck = 0;
    auto release_prev = [&]() {
        if (kCluster2) umma_commit_mc(&c.b_empty[sprev], c.cmask);
        else umma_commit(&c.b_empty[sprev]);
    };
    for (long long t = 0; t < c.my_tiles; ++t) {
        int s = 0;
        for (int ch = 0; ch < p.n_chunks; ++ch, ++ck) {
            const int buf = ck & 1;
            mbar_wait(&c.acc_empty[buf], ((ck >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d = c.tmem_base + (uint32_t)(buf * 2 * p.n_pad);
            const int s_stop = ch == p.n_chunks - 1 ? p.n_rows + 1 : s + p.chunk_steps;
            uint32_t acc = 0u;
            for (; s < s_stop; ++s) {
                mbar_wait(&c.img_full[si], phi);
                const uint32_t a0 = c.img_lo0 + (uint32_t)si * c.slot_step;
                if (s < p.n_rows) mbar_wait(&c.b_full[sj], pj);
                tc_fence_after();
                if (s == 0) {
                    // only output row y0 (upper accumulator half) has a window row on this frame row
                    issue_step<G, kX3, 1>(c, d, p.n_pad, a0, c.b_lo0 + (uint32_t)sj * c.b_slot_step, 0u, false, 0xFFu);
                } else if (s == p.n_rows) {
                    issue_step<G, kX3, 2>(c, d, p.n_pad, a0, c.b_lo0 + (uint32_t)sprev * c.b_slot_step, s == 1 ? 0u : acc, false,
                                          s == 1 ? 0xFFu : (uint32_t)p.step_mask[s]);
                } else {
                    // slots (sprev, sprev + 1) hold window rows (a-1, a); slot S mirrors slot 0 for the wrap-around
                    // the group-0 MMA carries the accumulate flag of a chunk's first step: skip nothing there
                    issue_step<G, kX3, 0>(c, d, p.n_pad, a0, c.b_lo0 + (uint32_t)sprev * c.b_slot_step, acc, s == 1,
                                          (acc && s != 1) ? (uint32_t)p.step_mask[s] : 0xFFu);
                }
                acc = 1u;
                umma_commit(&c.img_empty[si]);
                if (++si == p.img_slots) { si = 0; phi ^= 1; }
                if (s >= 1) release_prev();
                if (s < p.n_rows) {
                    sprev = sj;
                    if (++sj == S) { sj = 0; pj ^= 1; }
                }
            }
            umma_commit(&c.acc_full[buf]);
        }
    }
}

template <bool kScores, bool kX3, bool kCluster2>
__global__ void __launch_bounds__(384, 1)
map_h_kernel(const __grid_constant__ CUtensorMap map_img, const __grid_constant__ CUtensorMap map_b1,
             const __grid_constant__ CUtensorMap map_b2, const __grid_constant__ MapParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int n_bops = kX3 ? 2 : 1;
    const uint32_t b_bytes = (uint32_t)p.n_pad * 128;                 // one basis operand k-block (4 groups)
    const uint32_t ring_bytes = (uint32_t)(p.b_slots + 1) * b_bytes;  // one (k-block, operand) ring incl. the mirror slot
    const uint32_t n_rings = (uint32_t)(p.kb_per_row * n_bops);
    const uint32_t copy_bytes = (uint32_t)p.copy_q * 16;              // one staged row of one part
    const uint32_t slot_bytes = (uint32_t)n_bops * copy_bytes;        // [part]
    uint8_t* b_ring = smem;
    uint8_t* img_ring = smem + (size_t)n_rings * ring_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(img_ring + (((size_t)p.img_slots * slot_bytes + 15) & ~(size_t)15));
    uint64_t* img_full = bars;
    uint64_t* img_empty = img_full + p.img_slots;
    uint64_t* b_full = img_empty + p.img_slots;
    uint64_t* b_empty = b_full + p.b_slots;
    uint64_t* acc_full = b_empty + p.b_slots;       // [2]
    uint64_t* acc_empty = acc_full + 2;             // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    // Warp-role layout.  The SM's issue arbiter prefers the highest warp id of a sub-partition, so the
    // latency-critical single-thread roles take the LAST warpgroup: warps 0-7 epilogue (two warpgroups),
    // warp 8 frame-row TMA, warp 9 basis TMA, warp 10 MMA issuer, warp 11 TMEM allocation.
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wg = warp >> 2;
    constexpr int kWarpFrame = 8, kWarpBasis = 9, kWarpMma = 10, kWarpAlloc = 11;

    if (warp == kWarpFrame && lane == 0) {
        prefetch_tmap(&map_img);
        prefetch_tmap(&map_b1);
        prefetch_tmap(&map_b2);
    }
    if (warp == kWarpBasis && lane == 0) {
        for (int s = 0; s < p.img_slots; ++s) {
            mbar_init(&img_full[s], 1);
            mbar_init(&img_empty[s], 1);
        }
        for (int s = 0; s < p.b_slots; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], p.cluster);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], 8);
        }
        fence_barrier_init();
    }
    if (warp == kWarpAlloc) {
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

    // tile -> (output row pair, 512-pixel span, pixel phase r)
    auto decode = [&](long long tile, int& y0, int& span, int& r) {
        r = (int)(tile & 3);
        const long long rest = tile >> 2;
        span = (int)(rest % p.n_span);
        y0 = 2 * (p.pair0 + (int)(rest / p.n_span));
    };

    if (wg == 2) {
        reg_dec<kRegsCtl>();
        if (warp == kWarpFrame) {
            // ===================== frame-row producer =====================
            if (elect_one()) {
                int s = 0;
                uint32_t ph = 0;
                for (long long t = 0; t < my_tiles; ++t) {
                    const long long tile = blockIdx.x + t * gridDim.x;
                    int y0, span, r;
                    decode(tile, y0, span, r);
                    const bool live = tile < p.n_tiles;
                    const int c0 = span * kSpan - p.half + p.pad;            // = 4 q0 (4-byte words of a plane row)
                    for (int st = 0; st <= p.n_rows; ++st) {
                        mbar_wait(&img_empty[s], ph ^ 1);
                        if (ZB200_DEBUG_HOOKS && (p.dbg & 1)) {
                            mbar_arrive(&img_full[s]);
                            if (++s == p.img_slots) { s = 0; ph ^= 1; }
                            continue;
                        }
                        mbar_arrive_expect_tx(&img_full[s], slot_bytes);
                        uint8_t* slot = img_ring + (size_t)s * slot_bytes;
                        // dead tiles (past the end, cluster padding) read far outside the frame: all zeros
                        const int yy = live ? y0 - p.half + p.a_first + st - p.plane_row0 : -4 * p.k - 8;
                        for (int pt = 0; pt < n_bops; ++pt)
                            for (int bx = 0; bx < p.n_box; ++bx)
                                tma_load_3d(slot + (size_t)pt * copy_bytes + (size_t)bx * p.bw_q * 16, &map_img, &img_full[s],
                                            c0 + bx * p.bw_q * 4, yy, pt * 4 + r, kEvictLast);
                        if (++s == p.img_slots) { s = 0; ph ^= 1; }
                    }
                }
            }
            __syncwarp();
        } else if (warp == kWarpBasis) {
            // ===================== basis producer =====================
            // window row i of the tile -> slot (running row counter mod S); a row that lands in slot 0 is also
            // written to the mirror slot S, so that the slot pair (S-1, S) is contiguous like every other pair
            if (elect_one()) {
                const int b_rows = p.n_pad / p.cluster;
                const size_t off = (size_t)crank * b_rows * 128;
                int s = 0;
                uint32_t ph = 0;
                for (long long t = 0; t < my_tiles; ++t) {
                    for (int i = 0; i < p.n_rows; ++i) {
                        mbar_wait(&b_empty[s], ph ^ 1);
                        if (ZB200_DEBUG_HOOKS && (p.dbg & 16)) {
                            mbar_arrive(&b_full[s]);
                            if (++s == p.b_slots) { s = 0; ph ^= 1; }
                            continue;
                        }
                        const int copies = s == 0 ? 2 : 1;
                        mbar_arrive_expect_tx(&b_full[s], (uint32_t)copies * n_rings * b_bytes);
                        for (int cp = 0; cp < copies; ++cp) {
                            const int slot = cp == 0 ? s : p.b_slots;
                            for (int kb = 0; kb < p.kb_per_row; ++kb) {
                                const int kcol = (i * p.kb_per_row + kb) * kBlockK;
                                for (int op = 0; op < n_bops; ++op) {
                                    uint8_t* dst = b_ring + (size_t)(kb * n_bops + op) * ring_bytes + (size_t)slot * b_bytes;
                                    const CUtensorMap* mp = op == 0 ? &map_b1 : &map_b2;
                                    if (p.cluster == 1) tma_load_2d(dst, mp, &b_full[s], kcol, 0, kEvictLast);
                                    else tma_load_2d_mc(dst + off, mp, &b_full[s], kcol, (int)crank * b_rows, cmask, kEvictLast);
                                }
                            }
                        }
                        if (++s == p.b_slots) { s = 0; ph ^= 1; }
                    }
                }
            }
            __syncwarp();
        } else if (warp == kWarpMma) {
            // ===================== MMA issuer =====================
            // One elected thread; the loop body is specialised on G and fully unrolled per window-row step: at
            // N = 96..192 an MMA lasts 50-100 clk, and a loop with per-group branches, bit scans and
            // register->uniform moves (~60 SASS instructions per group, ncu) is issue-bound.
            if (elect_one()) {
                IssueCtx ctx;
                ctx.idesc1 = make_idesc_f16(p.n_pad);
                ctx.idesc2 = make_idesc_f16(2 * p.n_pad);
                ctx.slot_step = slot_bytes >> 4; ctx.copy_step = copy_bytes >> 4;
                ctx.b_slot_step = b_bytes >> 4; ctx.b_ring_step = ring_bytes >> 4;
                ctx.img_lo0 = desc_lo_toeplitz(smem_u32(img_ring));
                ctx.b_lo0 = desc_lo_sw128(smem_u32(b_ring));
                ctx.tmem_base = tmem_base; ctx.cmask = cmask; ctx.my_tiles = my_tiles;
                ctx.img_full = img_full; ctx.img_empty = img_empty; ctx.b_full = b_full; ctx.b_empty = b_empty;
                ctx.acc_full = acc_full; ctx.acc_empty = acc_empty;
                switch (p.n_groups) {
                    case 1: issue_tiles<1, kX3, kCluster2>(p, ctx); break;
                    case 2: issue_tiles<2, kX3, kCluster2>(p, ctx); break;
                    case 3: issue_tiles<3, kX3, kCluster2>(p, ctx); break;
                    case 4: issue_tiles<4, kX3, kCluster2>(p, ctx); break;
                    case 5: issue_tiles<5, kX3, kCluster2>(p, ctx); break;
                    case 6: issue_tiles<6, kX3, kCluster2>(p, ctx); break;
                    case 7: issue_tiles<7, kX3, kCluster2>(p, ctx); break;
                    default: issue_tiles<8, kX3, kCluster2>(p, ctx); break;
                }
            }
            __syncwarp();
        }
    } else {
        // ===== epilogue: warpgroup 0 -> accumulator columns [0, n_pad) = output row y0+1, warpgroup 1 -> [n_pad, 2 n_pad) = row y0
        reg_inc<kRegsEpi>();
        const int g = wg;
        const int q = warp & 3;
        const int n_cc = p.n_pad >> 4;
        const float scale = __ldg(p.scale);
        uint32_t ck = 0;
        for (long long t = 0; t < my_tiles; ++t) {
            const long long tile = blockIdx.x + t * gridDim.x;
            int y0, span, r;
            decode(tile, y0, span, r);
            float sum[kMaxColChunks][16];
#pragma unroll
            for (int cc = 0; cc < kMaxColChunks; ++cc)
#pragma unroll
                for (int i = 0; i < 16; ++i) sum[cc][i] = 0.f;
            for (int c = 0; c < p.n_chunks; ++c, ++ck) {
                const int buf = ck & 1;
                mbar_wait(&acc_full[buf], (ck >> 1) & 1u);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * 2 + g) * p.n_pad);
#pragma unroll
                for (int cc = 0; cc < kMaxColChunks; ++cc) {
                    if (cc < n_cc && !(ZB200_DEBUG_HOOKS && (p.dbg & 8))) {
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
            const int x = span * kSpan + 4 * (q * 32 + lane) + r;
            const int yl = y0 + (g == 0 ? 1 : 0) - p.row0;                 // local output row of this warpgroup's half
            const bool live = tile < p.n_tiles && x < p.W && yl >= 0 && yl < p.rows;
            if (ZB200_DEBUG_HOOKS && (p.dbg & 4)) {
                if (live) p.out_scores[(size_t)yl * p.W + x] = sum[0][0] + sum[5][15];
            } else if (live) {
                if (kScores) {
                    // pass 1: norms over the selected modes; the sums are squared in place
                    float s1 = 0.f, s2 = 0.f, sm = 0.f;
#pragma unroll
                    for (int cc = 0; cc < kMaxColChunks; ++cc) {
                        if (cc < n_cc) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const float z = sum[cc][i] * scale * p.selw[cc * 16 + i];
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
                    // pass 2: four independent weighted sums at a time
#pragma unroll
                    for (int fb = 0; fb < kFoldsPerLaunch; fb += 4) {
                        if (fb < p.n_folds) {
                            float num[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                            for (int cc = 0; cc < kMaxColChunks; ++cc) {
                                if (cc < n_cc) {
#pragma unroll
                                    for (int i = 0; i < 16; ++i)
#pragma unroll
                                        for (int j = 0; j < 4; ++j) num[j] = fmaf(p.wts[fb + j][cc * 16 + i], sum[cc][i], num[j]);
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (fb + j < p.n_folds) p.out_scores[((size_t)(fb + j) * p.rows + yl) * p.W + x] = num[j] / den;
                        }
                    }
                } else {
#pragma unroll
                    for (int cc = 0; cc < kMaxColChunks; ++cc) {
                        if (cc < n_cc) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const int col = cc * 16 + i;
                                if (col < p.n_modes) p.out_moments[((size_t)col * p.rows + yl) * p.W + x] = sum[cc][i] * scale;
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
    if (warp == kWarpAlloc) {
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

// 2-D view [rows][words] of 4-byte words (row pitch pitch_bytes), box = 32 words (one 128-B swizzle atom) x box_rows
static int encode_basis(CUtensorMap* map, const void* base, uint64_t words, uint64_t rows, uint64_t pitch_bytes,
                        uint32_t box_rows) {
    auto enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return ZB200_ECUDA;
    }
    cuuint64_t dims[2] = {words, rows};
    cuuint64_t strides[1] = {pitch_bytes};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(map basis) failed with CUresult %d", (int)r);
        return ZB200_ECUDA;
    }
    return ZB200_OK;
}

}  // namespace hmap

// Plan-time: which 16-tap groups of each window row touch the unit disk, and the packed fp16 basis.
// The grid is the reference's (x_i = -1 + 2 i/(k-1), rho <= 1, _zps.py:68-75); a group is kept when any of its
// taps has rho^2 <= 1 + 1e-9 (conservative: taps outside the disk are exact zeros of the basis anyway).
static int init_one_map_operand(zb200_plan* p, MapHalf& mh, int mode0, int n_modes_part) {
    using namespace hmap;
    const int k = p->size;
    const int G = (k + 15) / 16;
    mh.n_groups = G;
    mh.mode0 = mode0;
    mh.n_modes = n_modes_part;
    mh.rows_pad = round_up(n_modes_part, 16);
    std::vector<unsigned short> groups;
    mh.a_first = -1;
    mh.a_end = 0;
    for (int a = 0; a < k; ++a) {
        const double y = k > 1 ? -1.0 + 2.0 * a / (k - 1) : 0.0;
        unsigned short m = 0;
        for (int b = 0; b < k; ++b) {
            const double x = k > 1 ? -1.0 + 2.0 * b / (k - 1) : 0.0;
            if (x * x + y * y <= 1.0 + 1e-9) m |= (unsigned short)(1u << (b >> 4));
        }
        mh.act[a] = m;
        if (m) {
            if (mh.a_first < 0) mh.a_first = a;
            mh.a_end = a + 1;
        }
    }
    if (mh.a_first < 0) { mh.a_first = 0; mh.a_end = 1; mh.act[0] = 1; }
    // rows between the first and the last active one are active (the disk is convex); every group of those
    // rows is issued (skipping single groups saves 3-7 % of the MMAs and costs a branchy issue loop).
    // Operand order = issue order: window row i -> KB = ceil(G/4) k-blocks of 4 groups (see issue_tiles).
    const int KB = (G + 3) / 4;
    const int n_rows = mh.a_end - mh.a_first;
    mh.n_blocks = n_rows;
    mh.n_kb = n_rows * KB;
    groups.assign((size_t)mh.n_kb * 4, (unsigned short)0xFFFFu);
    for (int i = 0; i < n_rows; ++i)
        for (int g = 0; g < G; ++g) groups[(size_t)(i * KB + g / 4) * 4 + g % 4] = (unsigned short)(((mh.a_first + i) << 8) | g);
    mh.n_act = (int)groups.size();
    const int rows_pad = mh.rows_pad;
    const size_t halves = (size_t)rows_pad * mh.n_kb * 64;
    ZB_CUDA(cudaMalloc(&mh.b1, halves * 2));
    ZB_CUDA(cudaMalloc(&mh.b2, halves * 2));
    unsigned short* d_groups = nullptr;
    ZB_CUDA(cudaMalloc(&d_groups, groups.size() * sizeof(unsigned short)));
    ZB_CUDA(cudaMemcpy(d_groups, groups.data(), groups.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
    dim3 grid((unsigned)ceil_div(mh.n_kb * 64, 128), (unsigned)rows_pad);
    map_pack_basis_kernel<<<grid, 128>>>(p->basis64, mode0, n_modes_part, rows_pad, k, d_groups, mh.n_act, mh.n_kb,
                                         static_cast<__half*>(mh.b1), static_cast<__half*>(mh.b2));
    ZB_LAUNCHED();
    ZB_CUDA(cudaDeviceSynchronize());
    ZB_CUDA(cudaFree(d_groups));
    mh.max_cluster = 1;
    for (int lg = 0; lg < 2; ++lg) {
        const int c = 1 << lg;
        if (rows_pad % (8 * c) != 0) break;
        int rc = encode_basis(&mh.tmap_b1[lg], mh.b1, (uint64_t)mh.n_kb * 32, (uint64_t)rows_pad, (uint64_t)mh.n_kb * 128,
                              (uint32_t)(rows_pad / c));
        if (rc) return rc;
        rc = encode_basis(&mh.tmap_b2[lg], mh.b2, (uint64_t)mh.n_kb * 32, (uint64_t)rows_pad, (uint64_t)mh.n_kb * 128,
                          (uint32_t)(rows_pad / c));
        if (rc) return rc;
        mh.max_cluster = c;
    }
    mh.ready = true;
    return ZB200_OK;
}

// One operand holds at most 128 padded modes (two accumulator sets of 2 x n_pad TMEM columns); plans with more
// modes (n_max >= 15) are split into equal parts, each a separate pass of the kernel over the frame.
int init_map_half_operand(zb200_plan* p) {
    using namespace hmap;
    p->n_map_parts = 0;
    if (p->cc_major != 10 || p->size > kMaxWindow) return ZB200_OK;      // the SIMT map serves these
    const int parts = (int)ceil_div(p->n_modes, 128);
    if (parts > 4) return ZB200_OK;
    const int per = round_up((int)ceil_div(p->n_modes, parts), 16);
    int done = 0;
    for (int i = 0; i < parts && done < p->n_modes; ++i) {
        const int cnt = p->n_modes - done < per ? p->n_modes - done : per;
        int rc = init_one_map_operand(p, p->map_half[i], done, cnt);
        if (rc) return rc;
        done += cnt;
        p->n_map_parts = i + 1;
    }
    return ZB200_OK;
}

void free_map_half_operand(zb200_plan* p) {
    for (MapHalf& mh : p->map_half) {
        cudaFree(mh.b1);
        cudaFree(mh.b2);
        mh = MapHalf();
    }
    p->n_map_parts = 0;
}

bool map_h_supported(const zb200_plan* p, int precision) {
    if (precision != ZB200_PREC_F16 && precision != ZB200_PREC_F16X3) return false;
    return p->n_map_parts > 0;
}

// one pass of the kernel: the modes of one operand (all of them when the plan has a single part)
static int map_h_part(const zb200_plan* p, const MapHalf& mh, const float* d_img, int H, int W, int row0, int rows,
                      int precision, float* d_moments, float* d_scores, const float* h_w, const uint8_t* h_sel,
                      int n_folds, int norm_kind, cudaStream_t s) {
    using namespace hmap;
    const bool x3 = precision == ZB200_PREC_F16X3;
    MapParams prm{};
    prm.H = H; prm.W = W; prm.row0 = row0; prm.rows = rows;
    prm.k = p->size; prm.half = p->size / 2;
    prm.pad = 8 + (prm.half & 3);                                  // (pad - half) % 4 == 0, pad >= 8
    prm.n_pad = mh.rows_pad; prm.n_modes = mh.n_modes;
    prm.n_terms = x3 ? 3 : 1;
    prm.n_span = (int)ceil_div(W, kSpan);
    // output rows are paired by ABSOLUTE parity (2j, 2j+1), so a row's arithmetic does not depend on the band
    prm.pair0 = row0 >> 1;
    prm.n_pairs = ((row0 + rows - 1) >> 1) - prm.pair0 + 1;
    prm.n_tiles = (long long)prm.n_pairs * prm.n_span * 4;
    prm.n_groups = mh.n_groups;
    prm.kb_per_row = (mh.n_groups + 3) / 4;
    prm.a_first = mh.a_first; prm.n_rows = mh.a_end - mh.a_first;
    {
        const bool skip = knobs().map_gskip != 0;
        for (int st = 0; st <= prm.n_rows; ++st) {
            const unsigned up = st < prm.n_rows ? mh.act[mh.a_first + st] : 0u;
            const unsigned lo = st >= 1 ? mh.act[mh.a_first + st - 1] : 0u;
            prm.step_mask[st] = skip ? (unsigned char)((up | lo) & 0xFFu) : (unsigned char)0xFF;
        }
    }
    const int nq = 128 + 4 * mh.n_groups - 2;                     // 16-byte units one staged row spans
    prm.n_box = (int)ceil_div(nq, 64);
    prm.bw_q = round_up((int)ceil_div(nq, prm.n_box), 8);          // TMA destinations stay 128-B aligned
    prm.copy_q = prm.n_box * prm.bw_q;
    // drain after ~12 groups per output row (36 accumulating MMAs): the tensor core accumulates with
    // round-toward-zero.  At least 2 steps (step 1 completes the first chunk's lower row), and the last chunk
    // absorbs the remainder so that the final single-row step never stands alone.
    prm.chunk_steps = 12 / mh.n_groups > 2 ? 12 / mh.n_groups : 2;
    prm.n_chunks = prm.n_rows / prm.chunk_steps > 1 ? prm.n_rows / prm.chunk_steps : 1;
    prm.out_moments = d_moments; prm.out_scores = d_scores;
    prm.n_folds = n_folds; prm.norm_kind = norm_kind;
    if (d_scores) {
        for (int c = 0; c < p->n_modes; ++c) prm.selw[c] = h_sel[c] ? 1.f : 0.f;
        for (int f = 0; f < n_folds; ++f)
            for (int c = 0; c < p->n_modes; ++c) prm.wts[f][c] = h_sel[c] ? h_w[(size_t)f * p->n_modes + c] : 0.f;
    }
    prm.dbg = knobs().map_debug;

    // operand planes of the frame: (x1 | x2) x 4 pixel phases, overlapped 16-byte units
    const int n_parts = x3 ? 2 : 1;
    const int Wq = (W - 1 + prm.pad) / 4 + 1;
    const int n_maps = 4 * n_parts;
    // frame rows the band's windows read: output row y uses rows [y - half + a_first, y - half + a_first + n_rows]
    // (the +1: rows are computed in pairs), zero outside the frame.  Only those rows are expanded -- a row band of a
    // large frame (image-tile sharding, ZPs.mirror_map) no longer pays for the whole frame (0.1 ms at 4096^2).
    const int first_row = (row0 & ~1) - prm.half + prm.a_first;
    const int last_row = ((row0 + rows - 1) | 1) - prm.half + prm.a_first + prm.n_rows;
    prm.plane_row0 = first_row < 0 ? 0 : first_row;
    const int plane_end = last_row + 1 > H ? H : last_row + 1;
    const int plane_rows = plane_end > prm.plane_row0 ? plane_end - prm.plane_row0 : 1;
    const size_t plane_bytes = (size_t)plane_rows * Wq * 16;
    uint8_t* scratch = nullptr;
    ZB_CUDA(scratch_alloc(&scratch, 256 + (size_t)n_maps * plane_bytes, s));
    unsigned int* d_bits = reinterpret_cast<unsigned int*>(scratch);
    float* d_scale = reinterpret_cast<float*>(scratch + 16);
    uint4* planes = reinterpret_cast<uint4*>(scratch + 256);
    prm.scale = d_scale;
    {
        cudaError_t e = cudaMemsetAsync(d_bits, 0, 16, s);
        if (e != cudaSuccess) { cudaFreeAsync(scratch, s); set_error("cudaMemsetAsync failed: %s", cudaGetErrorString(e)); return ZB200_ECUDA; }
        const long long n = (long long)H * W;
        const int blocks = (int)(ceil_div(n, 256 * 8) < 4 * p->sm_count ? ceil_div(n, 256 * 8) : 4 * p->sm_count);
        map_absmax_kernel<<<blocks, 256, 0, s>>>(d_img, n, d_bits);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        const long long units = (long long)plane_rows * Wq;
        map_prepare_kernel<<<(unsigned)ceil_div(units, 128), 128, 0, s>>>(d_img, prm.plane_row0, plane_rows, W, Wq, prm.pad, n_parts,
                                                                           d_bits, p->inv_area, planes, d_scale);
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    // 3-D TMA descriptor [plane][row][4-byte words], box = 4*bw_q words x 1 row x 1 plane, no swizzle, zero fill outside
    CUtensorMap map_img;
    {
        auto enc = get_encode();
        if (!enc) { cudaFreeAsync(scratch, s); set_error("cuTensorMapEncodeTiled is not available"); return ZB200_ECUDA; }
        cuuint64_t dims[3] = {(cuuint64_t)Wq * 4, (cuuint64_t)plane_rows, (cuuint64_t)n_maps};
        cuuint64_t strides[2] = {(cuuint64_t)Wq * 16, (cuuint64_t)plane_bytes};
        cuuint32_t box[3] = {(cuuint32_t)prm.bw_q * 4, 1, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&map_img, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, planes, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            cudaFreeAsync(scratch, s);
            set_error("cuTensorMapEncodeTiled(frame planes) failed with CUresult %d", (int)r);
            return ZB200_ECUDA;
        }
    }

    int cluster = 2;
    if (knobs().tc_cluster == 1) cluster = 1;
    while (cluster > 1 && (cluster > mh.max_cluster || prm.n_tiles < 2 * cluster)) cluster >>= 1;
    prm.cluster = cluster;
    const int lg = cluster == 2 ? 1 : 0;

    const int n_rings = prm.kb_per_row * (x3 ? 2 : 1);
    const int slot_bytes = n_parts * prm.copy_q * 16;
    prm.img_slots = 8;
    prm.b_slots = prm.kb_per_row == 1 ? 4 : 2;
    if (knobs().map_bstages) prm.b_slots = knobs().map_bstages;
    if (knobs().map_slots) prm.img_slots = knobs().map_slots;
    auto smem_need = [&]() {
        return 1024 + (size_t)n_rings * (prm.b_slots + 1) * prm.n_pad * 128 + (((size_t)prm.img_slots * slot_bytes + 15) & ~(size_t)15) +
               8 * (2 * prm.img_slots + 2 * prm.b_slots + 4) + 16;
    };
    while (smem_need() > (size_t)kSmemLimit && prm.b_slots > 2) --prm.b_slots;
    const size_t smem = smem_need();
    if (smem > (size_t)kSmemLimit) {
        cudaFreeAsync(scratch, s);
        set_error("fp16-split dense map: window %d needs %zu B of shared memory", prm.k, smem);
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
    auto launch = [&](auto kernel) {
        cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err == cudaSuccess) err = cudaLaunchKernelEx(&cfg, kernel, map_img, mh.tmap_b1[lg], mh.tmap_b2[lg], prm);
        return err;
    };
    const int variant = (d_scores ? 4 : 0) | (x3 ? 2 : 0) | (cluster == 2 ? 1 : 0);
    switch (variant) {
        case 0: e = launch(map_h_kernel<false, false, false>); break;
        case 1: e = launch(map_h_kernel<false, false, true>); break;
        case 2: e = launch(map_h_kernel<false, true, false>); break;
        case 3: e = launch(map_h_kernel<false, true, true>); break;
        case 4: e = launch(map_h_kernel<true, false, false>); break;
        case 5: e = launch(map_h_kernel<true, false, true>); break;
        case 6: e = launch(map_h_kernel<true, true, false>); break;
        default: e = launch(map_h_kernel<true, true, true>); break;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaFreeAsync(scratch, s);
    if (e != cudaSuccess) {
        set_error("map_h_kernel launch failed: %s", cudaGetErrorString(e));
        return ZB200_ECUDA;
    }
    ZB_CUDA(cudaGetLastError());
    return ZB200_OK;
}

int map_h(const zb200_plan* p, const float* d_img, int H, int W, int row0, int rows, int precision,
          float* d_moments, float* d_scores, const float* h_w, const uint8_t* h_sel, int n_folds, int norm_kind,
          cudaStream_t s) {
    using namespace hmap;
    if (rows == 0) return ZB200_OK;
    if (!map_h_supported(p, precision)) {
        set_error("fp16-split tcgen05 dense map unsupported for n_max=%d size=%d (needs sm_100, size <= %d, <= 512 modes)",
                  p->n_max, p->size, kMaxWindow);
        return ZB200_EUNSUP;
    }
    if (d_scores) {
        ZB_CHECK_ARG(n_folds >= 1 && n_folds <= kMaxFolds, "n_folds=%d out of range [1,%d]", n_folds, kMaxFolds);
        ZB_CHECK_ARG(h_w && h_sel, "weights/select must not be null");
    }
    if (p->n_map_parts == 1) {
        if (d_scores && n_folds > kFoldsPerLaunch) {
            // the score tables of one launch hold kFoldsPerLaunch folds: split the fold list
            for (int f0 = 0; f0 < n_folds; f0 += kFoldsPerLaunch) {
                const int nf = n_folds - f0 < kFoldsPerLaunch ? n_folds - f0 : kFoldsPerLaunch;
                int rc = map_h_part(p, p->map_half[0], d_img, H, W, row0, rows, precision, nullptr,
                                    d_scores + (size_t)f0 * rows * W, h_w + (size_t)f0 * p->n_modes, h_sel, nf, norm_kind, s);
                if (rc) return rc;
            }
            return ZB200_OK;
        }
        return map_h_part(p, p->map_half[0], d_img, H, W, row0, rows, precision, d_moments, d_scores, h_w, h_sel, n_folds,
                          norm_kind, s);
    }
    // more than 128 padded modes (n_max >= 15): one pass per operand.  The scores need every mode of a pixel, so
    // they are computed from the materialised moment maps by the rot-score kernel (zb200_algebra.cu).
    float* moments = d_moments;
    const size_t plane = (size_t)rows * W;
    if (d_scores) ZB_CUDA(scratch_alloc(&moments, sizeof(float) * plane * p->n_modes, s));
    int rc = ZB200_OK;
    for (int i = 0; i < p->n_map_parts && rc == ZB200_OK; ++i) {
        const MapHalf& mh = p->map_half[i];
        rc = map_h_part(p, mh, d_img, H, W, row0, rows, precision, moments + plane * mh.mode0, nullptr, nullptr, nullptr, 0, 0, s);
    }
    if (rc == ZB200_OK && d_scores) {
        std::vector<double> wd((size_t)n_folds * p->n_modes);
        for (size_t i = 0; i < wd.size(); ++i) wd[i] = h_w[i];
        rc = zb200_rot_scores(ZB200_F32, moments, (int64_t)plane, p->n_modes, 1, (int64_t)plane, wd.data(), h_sel, n_folds,
                              norm_kind, d_scores, 1, (int64_t)plane, s);
    }
    if (d_scores) cudaFreeAsync(moments, s);
    return rc;
}

}  // namespace zb200
