// K2 -- patch gather at peak coordinates.
// Replaces KeyPoints.extract_patches (mtflearn/features/_keypoint.py:60-78):
//   (x, y) = rint(pts[i])  (half-to-even, np.rint),  s1 = k//2,
//   patch[i] = img[y-s1 : y-s1+k, x-s1 : x-s1+k]
// HBM-bound: every output byte is written once (N*k*k*4 B); the frame (<= 67 MB) stays
// L2-resident and each pixel is re-read ~k^2/spacing^2 times from L2, not from HBM.
#include "zb200_common.cuh"

namespace zb200 {

// One CTA streams whole patches.  Windows start at arbitrary columns, so reads are 4-byte loads
// (consecutive lanes -> consecutive pixels of a window row, served by L2: the frame is <= 67 MB),
// while every thread assembles four consecutive output floats and writes them with one 128-bit
// store (fully coalesced 512-B warp stores; the output stream is the HBM traffic of this kernel).
// Grid is a multiple of the SM count; CTAs stride over patches.
template <int kThreads>
__global__ void __launch_bounds__(kThreads)
gather_kernel(const float* __restrict__ img, int H, int W, const double* __restrict__ pts,
              long long n_pts, int k, float* __restrict__ out) {
    const int kk = k * k;
    const int half = k / 2;
    const bool vec = (k % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const int k4 = k >> 2;
    for (long long p = blockIdx.x; p < n_pts; p += gridDim.x) {
        // np.rint == round-half-to-even == rint() in the default rounding mode
        const int cx = (int)rint(pts[2 * p]);
        const int cy = (int)rint(pts[2 * p + 1]);
        const int x0 = cx - half, y0 = cy - half;
        float* dst = out + p * (long long)kk;
        const bool interior = x0 >= 0 && y0 >= 0 && x0 + k <= W && y0 + k <= H;
        if (interior && vec) {
            const float* src = img + (long long)y0 * W + x0;
            for (int v = threadIdx.x; v < (kk >> 2); v += kThreads) {
                const int r = v / k4, c = (v - r * k4) << 2;
                const float* s4 = src + (long long)r * W + c;
                const float4 val = make_float4(__ldg(s4), __ldg(s4 + 1), __ldg(s4 + 2), __ldg(s4 + 3));
                __stcs(reinterpret_cast<float4*>(dst) + v, val);          // streaming: written once
            }
        } else {
            for (int e = threadIdx.x; e < kk; e += kThreads) {
                const int r = e / k, c = e - r * k;
                const int yy = y0 + r, xx = x0 + c;
                dst[e] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(img + (long long)yy * W + xx) : 0.f;
            }
        }
    }
}

// window corners for the fused path: (x0, y0) = rint(pts) - k/2
__global__ void corners_kernel(const double* __restrict__ pts, long long n, int k, int2* __restrict__ xy0) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    xy0[i] = make_int2((int)rint(pts[2 * i]) - k / 2, (int)rint(pts[2 * i + 1]) - k / 2);
}

// shifted, zero-padded copies of the frame for the 16-byte gathers of the fused path:
//   plane r [y][u] = img0[y][u - L + r]   (img0 = frame, zero outside), u in [0, Wp), Wp % 4 == 0
// absmax (optional): receives the bits of max |frame| (plane 0 reads every pixel once) -- the exact range bound of the
// folded fp16-split projection that gathers from these planes.
// One thread per row and group of four padded columns: the seven frame pixels u-L .. u-L+6 give the float4 of every
// plane; stores are 16 bytes per plane and thread (512 contiguous bytes per warp), 72 MB for a 2048^2 frame.
__global__ void shift_planes_kernel(const float* __restrict__ img, int H, int W, int Wp, int L, float* __restrict__ planes,
                                    uint32_t* __restrict__ absmax) {
    const int u = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    uint32_t m = 0;
    if (u < Wp) {
        const float* row = img + (size_t)y * W;
        float v[7];
#pragma unroll
        for (int e = 0; e < 7; ++e) {
            const int x = u - L + e;
            v[e] = (x >= 0 && x < W) ? __ldg(row + x) : 0.f;
        }
        const size_t n = (size_t)H * Wp, o = (size_t)y * Wp + u;
#pragma unroll
        for (int r = 0; r < 4; ++r)
            *reinterpret_cast<float4*>(planes + (size_t)r * n + o) = make_float4(v[r], v[r + 1], v[r + 2], v[r + 3]);
#pragma unroll
        for (int e = 0; e < 4; ++e) m = max(m, __float_as_uint(v[e]) & 0x7FFFFFFFu);
    }
    if (absmax) {
        m = __reduce_max_sync(0xffffffffu, m);
        if ((threadIdx.x & 31) == 0 && m) atomicMax(absmax, m);
    }
}

}  // namespace zb200

extern "C" int zb200_project_peaks_f32(const zb200_plan* plan, const float* d_img, int H, int W, const double* d_pts_xy,
                                       int64_t n_pts, int precision, int out_kind, void* d_out, void* d_out2,
                                       void* stream) {
    using namespace zb200;
    ZB_CHECK_ARG(plan, "project_peaks: plan is null");
    ZB_CHECK_ARG(H > 0 && W > 0 && n_pts >= 0, "project_peaks: bad shape");
    ZB_CHECK_ARG(out_kind >= ZB200_OUT_REAL && out_kind <= ZB200_OUT_ABS_PHASE, "project_peaks: bad out_kind %d", out_kind);
    if (n_pts == 0) return ZB200_OK;
    ZB_CHECK_ARG(d_img && d_pts_xy && d_out, "project_peaks: null pointer");
    cudaStream_t s = as_stream(stream);
    // 64-pixel windows, n_max <= 13: the mirror-folded kernel gathers the windows itself (fp32-grade like tf32x3; the
    // range bound is the frame maximum, measured by the plane pre-pass)
    const bool folded = precision == ZB200_PREC_F16X3 && fold_gather_supported(plan) && knobs().tc_fold != 0 && H <= 65535;
    if (precision == ZB200_PREC_F16X3 && !folded) {
        set_error("project_peaks: the fp16-split fused gather needs a 64-pixel window and n_max <= 13");
        return ZB200_EUNSUP;
    }
    int2* xy0 = nullptr;
    ZB_CUDA(scratch_alloc(&xy0, sizeof(int2) * (size_t)n_pts + 16, s));
    corners_kernel<<<(unsigned)ceil_div(n_pts, 256), 256, 0, s>>>(d_pts_xy, (long long)n_pts, plan->size, xy0);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    GatherSource src{d_img, H, W, xy0};
    float* planes = nullptr;
    if (folded) {
        const int L = plan->size, Wp = round_up(W + 2 * L, 4);
        uint32_t* aux = reinterpret_cast<uint32_t*>(xy0 + n_pts);          // [0] absmax bits, [1] unused
        int rc = ZB200_OK;
        if (scratch_alloc(&planes, sizeof(float) * 4 * (size_t)H * Wp, s) == cudaSuccess) {
            ZB_CUDA(cudaMemsetAsync(aux, 0, 16, s));
            shift_planes_kernel<<<dim3((unsigned)ceil_div(Wp / 4, 128), (unsigned)H), 128, 0, s>>>(d_img, H, W, Wp, L, planes, aux);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            src.planes = planes;
            src.Wp = Wp;
            src.L = L;
            rc = project_fold(plan, nullptr, n_pts, out_kind, d_out, d_out2, s, nullptr, 0.0, aux, &src);
            cudaFreeAsync(planes, s);
            cudaFreeAsync(xy0, s);
            return rc;
        }
        cudaGetLastError();                                 // no memory for the planes: the unfolded tf32x3 route below
        planes = nullptr;
        precision = ZB200_PREC_TF32X3;
    }
    if (plan->size % 4 == 0 && !knobs().gather_4b && H <= 65535) {
        // windows start at arbitrary columns; four copies shifted by 0..3 pixels make every window row 16-byte aligned
        const int L = plan->size, Wp = round_up(W + 2 * L, 4);
        if (scratch_alloc(&planes, sizeof(float) * 4 * (size_t)H * Wp, s) == cudaSuccess) {
            shift_planes_kernel<<<dim3((unsigned)ceil_div(Wp / 4, 128), (unsigned)H), 128, 0, s>>>(d_img, H, W, Wp, L, planes, nullptr);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            src.planes = planes;
            src.Wp = Wp;
            src.L = L;
        } else {
            cudaGetLastError();
            planes = nullptr;                               // no memory for the planes: 4-byte gathers
        }
    }
    int rc = project_tc(plan, nullptr, n_pts, precision, out_kind, d_out, d_out2, nullptr, nullptr, 0, 0, s, &src);
    if (planes) cudaFreeAsync(planes, s);
    cudaFreeAsync(xy0, s);
    return rc;
}

extern "C" int zb200_gather_patches_f32(const float* d_img, int H, int W, const double* d_pts_xy,
                                        int64_t n_pts, int k, float* d_out, void* stream) {
    using namespace zb200;
    ZB_CHECK_ARG(H > 0 && W > 0 && k > 0 && n_pts >= 0, "gather: bad shape H=%d W=%d k=%d n=%lld", H, W, k,
                 (long long)n_pts);
    if (n_pts == 0) return ZB200_OK;
    ZB_CHECK_ARG(d_img && d_out && d_pts_xy, "gather: null pointer");
    int dev = 0, sms = 148;
    ZB_CUDA(cudaGetDevice(&dev));
    ZB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long want = (long long)sms * 8;          // 8 resident CTAs of 256 threads per SM
    const unsigned grid = (unsigned)(n_pts < want ? n_pts : want);
    gather_kernel<256><<<grid, 256, 0, as_stream(stream)>>>(d_img, H, W, d_pts_xy, (long long)n_pts, k, d_out);
    ZB_LAUNCHED();
    return ZB200_OK;
}

// ---- "next" row f4: overlap-add of the patch-SVD denoiser -----------------------------------------------------------
// Replaces reconstruct_patches (mtflearn/denoise/_denoise_svd.py:51-70): patches laid back at their start indices
// (ys x xs grid, row-major), summed, divided by the overlap count.  One thread per pixel walks the patches that
// cover it in the reference's accumulation order (increasing start row, then start column): float64, deterministic.
namespace zb200 {
__global__ void overlap_add_kernel(const double* __restrict__ patches, const int* __restrict__ ys, int ny,
                                   const int* __restrict__ xs, int nx, int kh, int kw, int H, int W, double* __restrict__ img) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W || y >= H) return;
    double acc = 0.0, cnt = 0.0;
    for (int a = 0; a < ny; ++a) {
        const int dy = y - ys[a];
        if (dy < 0 || dy >= kh) continue;
        for (int b = 0; b < nx; ++b) {
            const int dx = x - xs[b];
            if (dx < 0 || dx >= kw) continue;
            acc += patches[((size_t)(a * nx + b) * kh + dy) * kw + dx];
            cnt += 1.0;
        }
    }
    img[(size_t)y * W + x] = acc / cnt;                       // the start grids cover every pixel: cnt >= 1
}
}  // namespace zb200

extern "C" int zb200_overlap_add_f64(const double* d_patches, const int32_t* d_ys, int ny, const int32_t* d_xs, int nx,
                                     int kh, int kw, int H, int W, double* d_img, void* stream) {
    using namespace zb200;
    ZB_CHECK_ARG(d_patches && d_ys && d_xs && d_img, "overlap_add: null pointer");
    ZB_CHECK_ARG(ny >= 1 && nx >= 1 && kh >= 1 && kw >= 1 && H >= kh && W >= kw, "overlap_add: bad shape");
    dim3 grid((unsigned)ceil_div(W, 128), (unsigned)H);
    overlap_add_kernel<<<grid, 128, 0, as_stream(stream)>>>(d_patches, d_ys, ny, d_xs, nx, kh, kw, H, W, d_img);
    ZB_LAUNCHED();
    return ZB200_OK;
}
