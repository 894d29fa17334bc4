// Pieces shared by the projection kernels (zb200_project_tc.cu, zb200_project_fold.cu): epilogue arithmetic and the
// K5 pusher warp.  P is the kernel's parameter struct: it provides out, row_len, n_peers, peer_out[], n_patches.
#pragma once

#include "zb200_common.cuh"
#include "zb200_tc_ptx.cuh"

namespace zb200 {
namespace tc {

constexpr int kMaxPeers = 7;

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

// |z| and angle(z) for the fused epilogues, written for register count: the libm versions carry slow paths
// (denormal fix-ups behind a call) that made ptxas spill up to 228 bytes around them while 128 running sums are
// live.  sqrt.approx.ftz is within 1 ulp; the arctangent is the classic two-step range reduction
// (tan(pi/8), tan(3pi/8)) with a degree-4 polynomial in z = t^2: max error 1.2e-7 rad over the plane.
__device__ __forceinline__ float fast_abs2(float re, float im) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(re, re, im * im)));
    return r;
}
__device__ __forceinline__ float compact_atan2(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float t = mx > 0.f ? __fdividef(mn, mx) : 0.f;           // in [0, 1]
    float base = 0.f;
    if (t > 0.4142135623730950f) {                            // tan(pi/8)
        t = __fdividef(t - 1.f, t + 1.f);
        base = 0.7853981633974483f;
    }
    const float z = t * t;
    float r = fmaf(fmaf(fmaf(fmaf(8.05374449538e-2f, z, -1.38776856032e-1f), z, 1.99777106478e-1f), z, -3.33329491539e-1f) * z, t, t);
    r += base;
    if (ay > ax) r = 1.5707963267948966f - r;
    if (x < 0.f) r = 3.141592653589793f - r;
    return copysignf(r, y);
}


// ---- K5 inside K3: one warp forwards finished output tiles to the peer GPUs ----------------------------------
// The epilogue warps store a tile's rows to the local result array, fence, and bump `done` (shared memory, one
// count per storing warp; a monotonic counter, not an mbarrier: the epilogue never waits for the pusher, so phases
// could wrap).  The pusher warp then moves the tile's rows (contiguous in the row-major result) from the local array
// to the same offsets of every peer's array in 512-byte warp stores -- large NVLink packets, where the epilogue's own
// 4-byte stores at a 364-byte stride would not be.
constexpr uint32_t kPushChunk = 8192;               // bytes per staged chunk; two staging buffers
constexpr uint32_t kPushRegion = 2 * kPushChunk + 128;  // + their mbarriers

// Body of a tile (16-byte aligned, `body` bytes at `sb` locally, at `off_bytes` inside every peer's array):
// chunks are bulk-loaded from the local result array (L2 hits) into two shared-memory buffers -- the load of chunk
// c+1 is in flight while chunk c is being stored -- and written to the peers with 16-byte st.global from the whole
// warp (512 contiguous bytes per instruction and peer).  Remote stores are posted, so the warp streams at the rate
// NVLink accepts them; the first version issued one bulk STORE per peer instead and ran at ~3 GB/s per SM (the copy
// engine's window of outstanding remote writes), a register loop with ld.cg at ~0.7 GB/s per warp (L2 load latency).
template <class P>
__device__ __forceinline__ void push_body(const P& p, const uint8_t* sb, size_t body, size_t off_bytes, int lane,
                                          uint8_t* stage, uint64_t* bars, uint32_t (&phase)[2]) {
    const int n_chunks = (int)((body + kPushChunk - 1) / kPushChunk);
    auto issue = [&](int c) {
        const size_t o = (size_t)c * kPushChunk;
        const uint32_t n = (uint32_t)(body - o < kPushChunk ? body - o : kPushChunk);
        const int b = c & 1;
        mbar_arrive_expect_tx(&bars[b], n);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(stage + (size_t)b * kPushChunk)),
                     "l"(sb + o), "r"(n), "r"(smem_u32(&bars[b]))
                     : "memory");
    };
    if (n_chunks > 0 && lane == 0) issue(0);
    for (int c = 0; c < n_chunks; ++c) {
        const int b = c & 1;
        if (c + 1 < n_chunks && lane == 0) issue(c + 1);          // buffer (c+1)&1 was drained before the last __syncwarp
        mbar_wait(&bars[b], phase[b]);
        phase[b] ^= 1u;
        const size_t o = (size_t)c * kPushChunk;
        const uint32_t nv = (uint32_t)((body - o < kPushChunk ? body - o : kPushChunk) >> 4);
        const uint32_t src = smem_u32(stage + (size_t)b * kPushChunk);
        for (uint32_t i = lane; i < nv; i += 32) {
            uint4 v;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src + (i << 4)));
            for (int g = 0; g < p.n_peers; ++g)
                *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.peer_out[g]) + off_bytes + o + ((size_t)i << 4)) = v;
        }
        __syncwarp();
    }
}

template <class P>
__device__ __forceinline__ void push_tile(const P& p, long long r0, int n_rows, int lane, uint8_t* stage,
                                          uint64_t* bars, uint32_t (&phase)[2]) {
    const size_t nf = (size_t)n_rows * p.row_len;
    const float* src = p.out + (size_t)r0 * p.row_len;
    size_t head = ((16u - (unsigned)(reinterpret_cast<uintptr_t>(src) & 15u)) & 15u) >> 2;
    if (head > nf) head = nf;
    const size_t nv = (nf - head) >> 2, tail0 = head + (nv << 2);
    const size_t off = (size_t)r0 * p.row_len;
    // at most 3 leading and 3 trailing floats around the 16-byte aligned body
    if ((size_t)lane < head) {
        const float v = __ldcg(src + lane);
        for (int g = 0; g < p.n_peers; ++g) p.peer_out[g][off + lane] = v;
    }
    if (tail0 + lane < nf) {
        const float v = __ldcg(src + tail0 + lane);
        for (int g = 0; g < p.n_peers; ++g) p.peer_out[g][off + tail0 + lane] = v;
    }
    push_body(p, reinterpret_cast<const uint8_t*>(src + head), nv << 4, (off + head) * sizeof(float), lane, stage, bars, phase);
}

template <class P>
__device__ __forceinline__ void pusher_loop(const P& p, volatile unsigned* done, int warps_per_tile, int my_tiles, int lane,
                                            uint8_t* region, int tile_rows) {
    uint64_t* bars = reinterpret_cast<uint64_t*>(region);
    uint8_t* stage = region + 128;
    uint32_t phase[2] = {0u, 0u};
    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
    }
    __syncwarp();
    for (int t = 0; t < my_tiles; ++t) {
        const long long r0 = ((long long)blockIdx.x + (long long)t * gridDim.x) * tile_rows;
        if (r0 >= p.n_patches) break;
        const unsigned want = (unsigned)(t + 1) * (unsigned)warps_per_tile;
        while (*done < want) __nanosleep(256);
        __syncwarp();
        asm volatile("fence.proxy.async.global;" ::: "memory");      // the rows were written through the generic proxy
        const long long left = p.n_patches - r0;
        push_tile(p, r0, (int)(left < tile_rows ? left : tile_rows), lane, stage, bars, phase);
    }
    __threadfence_system();
}

}  // namespace tc
}  // namespace zb200
