// Synthetic STEM frame renderer ("next" row f1 of SURVEY section 8): sum of tapered Gaussians.
// Replaces mtflearn/datasets/_tapered_gaussian.py:3-97 (add_tapered_gaussian), the per-atom Python
// loop behind HoneyCombLattice.to_image (_honeycomb_lattice.py:169-226, 3.5 s per 2048^2 frame):
//   img[y,x] += A * exp(-r^2 / (2 sigma^2)) * (1 - 3 t^2 + 2 t^3),  t = r/R, R = r_factor*sigma, r <= R
// evaluated in float64 and accumulated atom by atom (in input order) into a float32 frame, exactly
// like the reference does, so frames agree to the last few float32 ulps.
#include "zb200_common.cuh"

namespace zb200 {

constexpr int RT_W = 32, RT_H = 8, RT_THREADS = RT_W * RT_H, RT_CAP = 256;

// One CTA = one 32x8 pixel tile.  Atoms whose support touches the tile are compacted, in input
// order, into shared memory (ballot-based ordered compaction over chunks of 256 atoms); every
// thread then walks the list for its pixel.
__global__ void __launch_bounds__(RT_THREADS)
render_kernel(const double* __restrict__ pts, const double* __restrict__ amps, long long n_atoms, double amp_scalar,
              double sigma, double cut, int H, int W, float* __restrict__ img, int accumulate) {
    __shared__ double sx[RT_CAP], sy[RT_CAP], sa[RT_CAP];
    __shared__ int warp_count[RT_THREADS / 32];
    __shared__ int n_list;
    const int tx = threadIdx.x % RT_W, ty = threadIdx.x / RT_W;
    const int x = blockIdx.x * RT_W + tx, y = blockIdx.y * RT_H + ty;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double x_lo = blockIdx.x * RT_W - cut - 1.0, x_hi = blockIdx.x * RT_W + RT_W + cut;
    const double y_lo = blockIdx.y * RT_H - cut - 1.0, y_hi = blockIdx.y * RT_H + RT_H + cut;
    const bool live = x < W && y < H;
    float acc = (live && accumulate) ? img[(size_t)y * W + x] : 0.f;
    const double inv2s2 = 0.5 / (sigma * sigma);
    if (threadIdx.x == 0) n_list = 0;
    __syncthreads();

    auto flush = [&]() {
        const int n = n_list;
        if (live) {
            for (int i = 0; i < n; ++i) {
                const double dx = (double)x - sx[i], dy = (double)y - sy[i];
                const double r = sqrt(dx * dx + dy * dy);
                if (r <= cut) {
                    const double t = r / cut;
                    const double v = sa[i] * exp(-(r * r) * inv2s2) * (1.0 - 3.0 * t * t + 2.0 * t * t * t);
                    acc = (float)((double)acc + v);          // float32 frame += float64 value, like numpy
                }
            }
        }
    };

    for (long long base = 0; base < n_atoms; base += RT_THREADS) {
        const long long i = base + threadIdx.x;
        double px = 0.0, py = 0.0;
        bool hit = false;
        if (i < n_atoms) {
            px = pts[2 * i];
            py = pts[2 * i + 1];
            hit = px >= x_lo && px <= x_hi && py >= y_lo && py <= y_hi;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) warp_count[warp] = __popc(mask);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < RT_THREADS / 32; ++w) {
            if (w < warp) before += warp_count[w];
            total += warp_count[w];
        }
        if (n_list + total > RT_CAP) {       // list full: consume what is there, then start over
            flush();
            __syncthreads();
            if (threadIdx.x == 0) n_list = 0;
            __syncthreads();
        }
        if (hit) {
            const int slot = n_list + before + __popc(mask & ((1u << lane) - 1u));
            sx[slot] = px;
            sy[slot] = py;
            sa[slot] = amps ? amps[i] : amp_scalar;
        }
        __syncthreads();
        if (threadIdx.x == 0) n_list += total;
        __syncthreads();
    }
    flush();
    if (live) img[(size_t)y * W + x] = acc;
}

}  // namespace zb200

extern "C" int zb200_render_atoms_f32(const double* d_pts_xy, const double* d_amps, double amp_scalar, int64_t n_atoms,
                                      double sigma, double r_factor, int H, int W, float* d_img, int accumulate,
                                      void* stream) {
    using namespace zb200;
    ZB_CHECK_ARG(H > 0 && W > 0 && n_atoms >= 0, "render: bad shape");
    ZB_CHECK_ARG(sigma > 0 && r_factor > 0, "sigma must be positive");
    ZB_CHECK_ARG(d_img && (d_pts_xy || n_atoms == 0), "render: null pointer");
    dim3 grid((unsigned)ceil_div(W, RT_W), (unsigned)ceil_div(H, RT_H));
    render_kernel<<<grid, RT_THREADS, 0, as_stream(stream)>>>(d_pts_xy, d_amps, (long long)n_atoms, amp_scalar, sigma,
                                                              r_factor * sigma, H, W, d_img, accumulate);
    ZB_LAUNCHED();
    return ZB200_OK;
}
