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


// Honeycomb lattice sites (replaces the Python double loop of HoneyCombLattice._generate_coordinates,
// mtflearn/datasets/_honeycomb_lattice.py:103-140): for (n1, n2) in [-N, N]^2, in that order,
//   R = n1*a1 + n2*a2;  A = (R + dA) + offset;  B = (R + dB) + offset;  rotate by Rmat;  + box centre
// in float64 with the reference's operation order (the rotation is a (P,2)x(2,2) product there: x*c + y*(-s),
// x*s + y*c).  Optional per-site jitter (host-drawn, the reference's own RNG stream) is added last.
__global__ void lattice_coords_kernel(int N, double a1x, double a1y, double a2x, double a2y, double dAx, double dAy,
                                      double dBx, double dBy, double offx, double offy, double c, double s,
                                      double centre, const double* __restrict__ jitter_a, const double* __restrict__ jitter_b,
                                      double* __restrict__ out_a, double* __restrict__ out_b) {
    const long long side = 2ll * N + 1, total = side * side;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const double n1 = (double)(i / side - N), n2 = (double)(i % side - N);
    const double rx = __dadd_rn(__dmul_rn(n1, a1x), __dmul_rn(n2, a2x));
    const double ry = __dadd_rn(__dmul_rn(n1, a1y), __dmul_rn(n2, a2y));
    const double ax = __dadd_rn(__dadd_rn(rx, dAx), offx), ay = __dadd_rn(__dadd_rn(ry, dAy), offy);
    const double bx = __dadd_rn(__dadd_rn(rx, dBx), offx), by = __dadd_rn(__dadd_rn(ry, dBy), offy);
    auto rot = [&](double x, double y, const double* jit, double* out) {
        double xr = __dadd_rn(__dmul_rn(x, c), __dmul_rn(y, -s)) + centre;
        double yr = __dadd_rn(__dmul_rn(x, s), __dmul_rn(y, c)) + centre;
        if (jit) { xr += jit[2 * i]; yr += jit[2 * i + 1]; }
        out[2 * i] = xr;
        out[2 * i + 1] = yr;
    };
    rot(ax, ay, jitter_a, out_a);
    rot(bx, by, jitter_b, out_b);
}

// Delta placement + Gaussian blur of TMDImageSimulator.simulate (mtflearn/datasets/_tmd_simulator.py:151-186):
// atoms are rounded to pixels (np.round: half to even), atoms outside the frame are dropped, each kept atom stamps
// scale * A * exp(-(dx^2 + dy^2) / (2 sigma^2)) on the (2*half+1)^2 pixels around it -- what the reference's
// fftconvolve(delta, kernel, 'same') evaluates up to its FFT round-off.  One species per call:
//   img = float32(img + blurred)    (the reference's `total_img += blurred`), blurred summed in float64.
__global__ void __launch_bounds__(RT_THREADS)
stamp_kernel(const double* __restrict__ pts, const float* __restrict__ scales, long long n_atoms, double amp, double sigma,
             int half, int H, int W, float* __restrict__ img, int accumulate) {
    __shared__ int sx[RT_CAP], sy[RT_CAP];
    __shared__ float sv[RT_CAP];
    __shared__ int warp_count[RT_THREADS / 32];
    __shared__ int n_list;
    const int tx = threadIdx.x % RT_W, ty = threadIdx.x / RT_W;
    const int x = blockIdx.x * RT_W + tx, y = blockIdx.y * RT_H + ty;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x_lo = blockIdx.x * RT_W - half, x_hi = blockIdx.x * RT_W + RT_W - 1 + half;
    const int y_lo = blockIdx.y * RT_H - half, y_hi = blockIdx.y * RT_H + RT_H - 1 + half;
    const bool live = x < W && y < H;
    double acc = 0.0;
    const double inv2s2 = 1.0 / (2.0 * sigma * sigma);
    if (threadIdx.x == 0) n_list = 0;
    __syncthreads();
    auto flush = [&]() {
        const int n = n_list;
        if (live) {
            for (int i = 0; i < n; ++i) {
                const int dx = x - sx[i], dy = y - sy[i];
                if (abs(dx) <= half && abs(dy) <= half)
                    acc += (double)sv[i] * (amp * exp(-(double)(dx * dx + dy * dy) * inv2s2));
            }
        }
    };
    for (long long base = 0; base < n_atoms; base += RT_THREADS) {
        const long long i = base + threadIdx.x;
        int xi = 0, yi = 0;
        float sc = 0.f;
        bool hit = false;
        if (i < n_atoms) {
            xi = (int)rint(pts[2 * i]);
            yi = (int)rint(pts[2 * i + 1]);
            sc = scales[i];
            hit = xi >= 0 && xi < W && yi >= 0 && yi < H && xi >= x_lo && xi <= x_hi && yi >= y_lo && yi <= y_hi;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) warp_count[warp] = __popc(mask);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < RT_THREADS / 32; ++w) {
            if (w < warp) before += warp_count[w];
            total += warp_count[w];
        }
        if (n_list + total > RT_CAP) {
            flush();
            __syncthreads();
            if (threadIdx.x == 0) n_list = 0;
            __syncthreads();
        }
        if (hit) {
            const int slot = n_list + before + __popc(mask & ((1u << lane) - 1u));
            sx[slot] = xi;
            sy[slot] = yi;
            sv[slot] = sc;
        }
        __syncthreads();
        if (threadIdx.x == 0) n_list += total;
        __syncthreads();
    }
    flush();
    if (live) {
        const size_t o = (size_t)y * W + x;
        img[o] = accumulate ? (float)((double)img[o] + acc) : (float)acc;
    }
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

extern "C" int zb200_lattice_coords_f64(int n_index, const double* h_a1, const double* h_a2, const double* h_dA,
                                        const double* h_dB, const double* h_offset, double angle_rad, double centre,
                                        const double* d_jitter_a, const double* d_jitter_b, double* d_coords_a,
                                        double* d_coords_b, void* stream) {
    using namespace zb200;
    ZB_CHECK_ARG(n_index >= 0 && n_index <= 16384, "lattice_coords: index range %d out of bounds", n_index);
    ZB_CHECK_ARG(h_a1 && h_a2 && h_dA && h_dB && h_offset && d_coords_a && d_coords_b, "lattice_coords: null pointer");
    const long long side = 2ll * n_index + 1, total = side * side;
    lattice_coords_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(
        n_index, h_a1[0], h_a1[1], h_a2[0], h_a2[1], h_dA[0], h_dA[1], h_dB[0], h_dB[1], h_offset[0], h_offset[1],
        cos(angle_rad), sin(angle_rad), centre, d_jitter_a, d_jitter_b, d_coords_a, d_coords_b);
    ZB_LAUNCHED();
    return ZB200_OK;
}

extern "C" int zb200_render_stamps_f32(const double* d_pts_xy, const float* d_scales, int64_t n_atoms, double amplitude,
                                       double sigma, int kernel_size, int H, int W, float* d_img, int accumulate,
                                       void* stream) {
    using namespace zb200;
    ZB_CHECK_ARG(H > 0 && W > 0 && n_atoms >= 0, "render_stamps: bad shape");
    ZB_CHECK_ARG(sigma > 0 && kernel_size >= 1 && (kernel_size & 1) && kernel_size <= 511, "render_stamps: bad kernel");
    ZB_CHECK_ARG(d_img && ((d_pts_xy && d_scales) || n_atoms == 0), "render_stamps: null pointer");
    dim3 grid((unsigned)ceil_div(W, RT_W), (unsigned)ceil_div(H, RT_H));
    stamp_kernel<<<grid, RT_THREADS, 0, as_stream(stream)>>>(d_pts_xy, d_scales, (long long)n_atoms, amplitude, sigma,
                                                             kernel_size / 2, H, W, d_img, accumulate);
    ZB_LAUNCHED();
    return ZB200_OK;
}
