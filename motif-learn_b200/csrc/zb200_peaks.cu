// Peak detection ("next" row f2 of SURVEY section 8): the step right before the Zernike hot path.
// Replaces mtflearn.features.local_max (mtflearn/features/_local_max_v2.py:6-66):
//   peaks = skimage.feature.peak_local_max(image, min_distance=1, threshold_abs=threshold)
//   keep  = filter_peaks_by_distance(image, peaks, min_distance)      # intensity-ordered radius NMS
// skimage's routine at min_distance=1 reduces to (published algorithm, skimage/feature/peak.py):
//   candidate <=> image == maximum_filter(image, 3x3, mode='nearest')  and  image > threshold,
//   the 1-pixel frame border excluded, no candidate at all for a constant image; candidates ordered by
//   intensity (descending, stable = raster order among equals); ensure_spacing(spacing=1) removes nothing
//   (it rejects points strictly closer than 1 pixel).
// The reference's greedy suppression (visit peaks by descending intensity; a visited, still-kept peak
// suppresses every other peak within Euclidean distance <= min_distance) is sequential as written.  Here it
// runs as a fixed point over a rank map of the frame: peak i is SUPPRESSED as soon as one higher-ranked
// neighbour is KEPT, and KEPT as soon as all higher-ranked neighbours are SUPPRESSED -- the same set, a
// handful of sweeps.  Ordering among exactly equal intensities is raster order (the reference leaves it to
// numpy's unstable argsort).
#include "zb200_common.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

namespace zb200 {

// order-preserving map float -> uint32 (ascending); -0 is folded onto +0 first (numpy compares them equal)
__device__ __forceinline__ uint32_t float_order(float f) {
    if (f == 0.f) f = 0.f;
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// [0] = min, [1] = max of the frame, as ordered uint32 (init: [0] = 0xFFFFFFFF, [1] = 0)
__global__ void peaks_minmax_kernel(const float* __restrict__ img, long long n, uint32_t* __restrict__ mm) {
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    const long long stride = (long long)gridDim.x * blockDim.x, tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(img) & 15) == 0) {
        const long long n4 = n >> 2;
        const float4* img4 = reinterpret_cast<const float4*>(img);
        for (long long i = tid; i < n4; i += stride) {
            const float4 v = __ldg(img4 + i);
            const uint32_t a = float_order(v.x), b = float_order(v.y), c = float_order(v.z), d = float_order(v.w);
            lo = min(min(lo, a), min(b, min(c, d)));
            hi = max(max(hi, a), max(b, max(c, d)));
        }
        for (long long i = (n4 << 2) + tid; i < n; i += stride) {
            const uint32_t o = float_order(__ldg(img + i));
            lo = min(lo, o);
            hi = max(hi, o);
        }
    } else {
        for (long long i = tid; i < n; i += stride) {
            const uint32_t o = float_order(__ldg(img + i));
            lo = min(lo, o);
            hi = max(hi, o);
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, s));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, s));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(mm, lo);
        atomicMax(mm + 1, hi);
    }
}

// candidates -> keys (descending-intensity order in the high word, raster index in the low word)
// One warp walks a strip of 30 columns x kStripRows rows top-down: a lane loads ONE pixel per row, its neighbours come
// from the lanes beside it (lanes 0 and 31 only carry the halo columns), the 3-row window lives in registers -- one
// load, two shuffles and a handful of max per pixel instead of nine loads.
constexpr int kStripRows = 32;
__global__ void __launch_bounds__(128)
peaks_candidates_kernel(const float* __restrict__ img, int H, int W, int has_thr, double thr,
                        const uint32_t* __restrict__ mm, unsigned long long* __restrict__ keys,
                        unsigned long long capacity, unsigned long long* __restrict__ count) {
    if (mm[0] == mm[1]) return;                                     // constant frame: no peaks (skimage)
    const int lane = threadIdx.x & 31;
    const int strip = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int x = strip * 30 + lane;                                // lanes 1..30 own columns; 0 and 31 are halo
    const int y_begin = blockIdx.y * kStripRows;
    if (strip * 30 + 1 >= W - 1) return;                            // warp-uniform: no interior column in this strip
    const bool col_ok = x < W;
    const int xc = col_ok ? x : W - 1;                              // clamp loads; such lanes never report a hit
    const uint32_t floor_order = mm[0];
    auto row_val = [&](int y) { return (y >= 0 && y < H) ? __ldg(img + (size_t)y * W + xc) : -INFINITY; };
    auto hmax3 = [&](float v) {                                     // max of the lane's pixel and its two neighbours
        const float l = __shfl_up_sync(0xffffffffu, v, 1), r = __shfl_down_sync(0xffffffffu, v, 1);
        return fmaxf(v, fmaxf(l, r));
    };
    float v_mid = row_val(y_begin), h_top = hmax3(row_val(y_begin - 1)), h_mid = hmax3(v_mid);
    const int y_end = min(y_begin + kStripRows, H);
    for (int y = y_begin; y < y_end; ++y) {
        const float v_bot = row_val(y + 1);
        const float h_bot = hmax3(v_bot);
        const float v = v_mid, m = fmaxf(h_mid, fmaxf(h_top, h_bot));
        bool hit = false;
        if (lane >= 1 && lane <= 30 && x >= 1 && x < W - 1 && y >= 1 && y < H - 1) {
            // threshold None -> image.min(): compare in the ordered domain so the default needs no host round trip
            const bool above = has_thr ? ((double)v > thr) : (float_order(v) > floor_order);
            hit = (v == m) && above;
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, hit);
        if (ballot) {
            unsigned long long base = 0;
            if (lane == (__ffs(ballot) - 1)) base = atomicAdd(count, (unsigned long long)__popc(ballot));
            base = __shfl_sync(0xffffffffu, base, __ffs(ballot) - 1);
            if (hit) {
                const unsigned long long slot = base + __popc(ballot & ((1u << lane) - 1u));
                if (slot < capacity)
                    keys[slot] = ((unsigned long long)(~float_order(v)) << 32) | (unsigned long long)((unsigned)y * (unsigned)W + (unsigned)x);
            }
        }
        h_top = h_mid;
        h_mid = h_bot;
        v_mid = v_bot;
    }
}

__global__ void peaks_rank_kernel(const unsigned long long* __restrict__ keys, long long n, int* __restrict__ rank_map) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rank_map[(uint32_t)keys[i]] = (int)i;
}

// one sweep of the suppression fixed point; state: 0 undecided, 1 kept, 2 suppressed
__global__ void peaks_nms_kernel(const unsigned long long* __restrict__ keys, long long n, const int* __restrict__ rank_map,
                                 int H, int W, int R, double r2, unsigned char* __restrict__ state,
                                 unsigned int* __restrict__ undecided, const unsigned int* __restrict__ undecided_before) {
    // undecided_before: the previous sweep's count (nullptr for the first sweep of a batch): 0 = already settled
    if (undecided_before && *undecided_before == 0u) return;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || state[i] != 0) return;
    const uint32_t pix = (uint32_t)keys[i];
    const int y = (int)(pix / (uint32_t)W), x = (int)(pix - (uint32_t)y * (uint32_t)W);
    bool pending = false, dead = false;
    for (int dy = -R; dy <= R && !dead; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
        for (int dx = -R; dx <= R; ++dx) {
            const int xx = x + dx;
            if (xx < 0 || xx >= W || (double)(dx * dx + dy * dy) > r2) continue;
            const int j = __ldg(rank_map + (size_t)yy * W + xx);
            if (j < 0 || j >= i) continue;                 // not a peak / lower priority / itself
            const unsigned char sj = state[j];
            if (sj == 1) { dead = true; break; }
            if (sj == 0) pending = true;
        }
    }
    if (dead) state[i] = 2;
    else if (!pending) state[i] = 1;
    else atomicAdd(undecided, 1u);
}

__global__ void peaks_flags_kernel(const unsigned char* __restrict__ state, long long n, unsigned char* __restrict__ flags) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = state[i] == 1 ? 1 : 0;
}

__global__ void peaks_emit_kernel(const unsigned long long* __restrict__ keys, long long n, int W, int32_t* __restrict__ xy) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t pix = (uint32_t)keys[i];
    const uint32_t y = pix / (uint32_t)W;
    xy[2 * i] = (int32_t)(pix - y * (uint32_t)W);
    xy[2 * i + 1] = (int32_t)y;
}

}  // namespace zb200

using namespace zb200;

#define ZB_PEAKS_CUDA(call)                                                                   \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__)); \
            rc = ZB200_ECUDA;                                                                 \
            goto done;                                                                        \
        }                                                                                     \
    } while (0)

extern "C" int zb200_local_max_f32(const float* d_img, int H, int W, double min_distance, int has_threshold,
                                   double threshold, int32_t* d_xy_out, int64_t capacity, int64_t* h_count,
                                   int64_t* h_candidates, void* stream) {
    ZB_CHECK_ARG(d_img && h_count, "local_max: null argument");
    ZB_CHECK_ARG(H >= 1 && W >= 1 && (long long)H * W < (1ll << 31), "local_max: bad frame size %dx%d", H, W);
    ZB_CHECK_ARG(min_distance >= 0.0 && min_distance <= 1024.0, "local_max: min_distance %g out of range", min_distance);
    ZB_CHECK_ARG(capacity >= 0 && (d_xy_out || capacity == 0), "local_max: null output with non-zero capacity");
    cudaStream_t s = as_stream(stream);
    *h_count = 0;
    if (h_candidates) *h_candidates = 0;
    const long long n_pix = (long long)H * W;
    // candidates of a 3x3 maximum filter: at most every pixel (plateaus); size the key buffers for that
    const size_t key_bytes = sizeof(unsigned long long) * (size_t)n_pix;
    int rc = ZB200_OK;
    uint8_t* pool = nullptr;
    void* cub_tmp = nullptr;
    unsigned long long n_cand = 0;
    long long n = 0;
    // layout of the scratch block
    const size_t off_keys0 = 256, off_keys1 = off_keys0 + key_bytes, off_rank = off_keys1 + key_bytes;
    const size_t off_state = off_rank + sizeof(int) * (size_t)n_pix, off_flags = off_state + (size_t)n_pix;
    const size_t total = off_flags + (size_t)n_pix + 256;
    ZB_CUDA(scratch_alloc(&pool, total, s));
    {
        uint32_t* mm = reinterpret_cast<uint32_t*>(pool);                       // [0] min, [1] max
        unsigned long long* count = reinterpret_cast<unsigned long long*>(pool + 16);
        unsigned int* undecided = reinterpret_cast<unsigned int*>(pool + 32);     // [4]: one per sweep of a batch
        unsigned long long* n_sel = reinterpret_cast<unsigned long long*>(pool + 64);
        unsigned long long* keys0 = reinterpret_cast<unsigned long long*>(pool + off_keys0);
        unsigned long long* keys1 = reinterpret_cast<unsigned long long*>(pool + off_keys1);
        int* rank_map = reinterpret_cast<int*>(pool + off_rank);
        unsigned char* state = pool + off_state;
        unsigned char* flags = pool + off_flags;
        const uint32_t init[2] = {0xFFFFFFFFu, 0u};
        ZB_PEAKS_CUDA(cudaMemsetAsync(pool, 0, 256, s));
        ZB_PEAKS_CUDA(cudaMemcpyAsync(mm, init, sizeof(init), cudaMemcpyHostToDevice, s));
        {
            const int blocks = (int)(ceil_div(n_pix, 256 * 8) < 1184 ? ceil_div(n_pix, 256 * 8) : 1184);
            peaks_minmax_kernel<<<blocks, 256, 0, s>>>(d_img, n_pix, mm);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            dim3 blk(128), grd((unsigned)ceil_div(ceil_div(W, 30), 4), (unsigned)ceil_div(H, kStripRows));
            peaks_candidates_kernel<<<grd, blk, 0, s>>>(d_img, H, W, has_threshold, threshold, mm, keys0,
                                                        (unsigned long long)n_pix, count);
            g_launches.fetch_add(1, std::memory_order_relaxed);
        }
        ZB_PEAKS_CUDA(cudaMemcpyAsync(&n_cand, count, sizeof(n_cand), cudaMemcpyDeviceToHost, s));
        ZB_PEAKS_CUDA(cudaStreamSynchronize(s));
        n = (long long)n_cand;
        if (h_candidates) *h_candidates = n;
        if (n == 0) goto done;
        {
            size_t tmp_bytes = 0;
            ZB_PEAKS_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys0, keys1, (int)n, 0, 64, s));
            ZB_PEAKS_CUDA(scratch_alloc(&cub_tmp, tmp_bytes, s));
            ZB_PEAKS_CUDA(cub::DeviceRadixSort::SortKeys(cub_tmp, tmp_bytes, keys0, keys1, (int)n, 0, 64, s));
            g_launches.fetch_add(1, std::memory_order_relaxed);
            ZB_PEAKS_CUDA(cudaFreeAsync(cub_tmp, s));
            cub_tmp = nullptr;
        }
        ZB_PEAKS_CUDA(cudaMemsetAsync(rank_map, 0xFF, sizeof(int) * (size_t)n_pix, s));
        ZB_PEAKS_CUDA(cudaMemsetAsync(state, 0, (size_t)n, s));
        peaks_rank_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(keys1, n, rank_map);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        {
            const int R = (int)floor(min_distance);
            const double r2 = min_distance * min_distance;
            // Sweeps run in batches of 4 between host checks (a lattice frame settles in 2-4 sweeps; a sweep over
            // an already settled list is a few microseconds, a host round trip costs more).  counters[j] = peaks
            // still undecided after sweep j of the batch.
            unsigned int left = 1;
            constexpr int kBatch = 4, kMaxSweeps = 8192;
            for (int sweep = 0; sweep < kMaxSweeps && left != 0; sweep += kBatch) {
                ZB_PEAKS_CUDA(cudaMemsetAsync(undecided, 0, sizeof(unsigned int) * kBatch, s));
                for (int j = 0; j < kBatch; ++j) {
                    peaks_nms_kernel<<<(unsigned)ceil_div(n, 128), 128, 0, s>>>(keys1, n, rank_map, H, W, R, r2, state, undecided + j,
                                                                                j ? undecided + j - 1 : nullptr);
                    g_launches.fetch_add(1, std::memory_order_relaxed);
                }
                ZB_PEAKS_CUDA(cudaMemcpyAsync(&left, undecided + (kBatch - 1), sizeof(left), cudaMemcpyDeviceToHost, s));
                ZB_PEAKS_CUDA(cudaStreamSynchronize(s));
            }
            if (left != 0) {
                // a dependency chain longer than the sweep cap (e.g. a huge plateau of equal maxima): report it
                // rather than silently dropping the undecided peaks
                set_error("local_max: suppression did not settle within %d sweeps (%u peaks undecided)", kMaxSweeps, left);
                rc = ZB200_EUNSUP;
                goto done;
            }
        }
        peaks_flags_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(state, n, flags);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        {
            size_t tmp_bytes = 0;
            ZB_PEAKS_CUDA(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, keys1, flags, keys0, n_sel, (int)n, s));
            ZB_PEAKS_CUDA(scratch_alloc(&cub_tmp, tmp_bytes, s));
            ZB_PEAKS_CUDA(cub::DeviceSelect::Flagged(cub_tmp, tmp_bytes, keys1, flags, keys0, n_sel, (int)n, s));
            g_launches.fetch_add(1, std::memory_order_relaxed);
            ZB_PEAKS_CUDA(cudaFreeAsync(cub_tmp, s));
            cub_tmp = nullptr;
        }
        unsigned long long kept = 0;
        ZB_PEAKS_CUDA(cudaMemcpyAsync(&kept, n_sel, sizeof(kept), cudaMemcpyDeviceToHost, s));
        ZB_PEAKS_CUDA(cudaStreamSynchronize(s));
        *h_count = (int64_t)kept;
        if ((int64_t)kept > capacity) {
            if (capacity > 0 || d_xy_out) {
                set_error("local_max: %llu peaks do not fit the output capacity %lld", kept, (long long)capacity);
                rc = ZB200_EINVAL;
            }
            goto done;                                     // capacity 0 + null output = count-only query
        }
        if (kept) {
            peaks_emit_kernel<<<(unsigned)ceil_div((long long)kept, 256), 256, 0, s>>>(keys0, (long long)kept, W, d_xy_out);
            g_launches.fetch_add(1, std::memory_order_relaxed);   // stream-ordered: no host wait for the coordinates
        }
    }
done:
    if (cub_tmp) cudaFreeAsync(cub_tmp, s);
    cudaFreeAsync(pool, s);
    if (rc == ZB200_OK) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { set_error("local_max: %s", cudaGetErrorString(e)); rc = ZB200_ECUDA; }
    }
    return rc;
}
