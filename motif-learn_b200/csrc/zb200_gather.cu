// K2 -- patch gather at peak coordinates.
// Replaces KeyPoints.extract_patches (mtflearn/features/_keypoint.py:60-78):
//   (x, y) = rint(pts[i])  (half-to-even, np.rint),  s1 = k//2,
//   patch[i] = img[y-s1 : y-s1+k, x-s1 : x-s1+k]
// HBM-bound: every output byte is written once (N*k*k*4 B); the frame (<= 67 MB) stays
// L2-resident and each pixel is re-read ~k^2/spacing^2 times from L2, not from HBM.
#include "zb200_common.cuh"

namespace zb200 {

// One CTA streams whole patches: consecutive lanes read consecutive pixels of a window
// row (coalesced, arbitrary 4-B alignment) and write consecutive output floats (fully
// coalesced 128-B lines).  Grid is a multiple of the SM count; CTAs stride over patches.
template <int kThreads>
__global__ void __launch_bounds__(kThreads)
gather_kernel(const float* __restrict__ img, int H, int W, const double* __restrict__ pts,
              long long n_pts, int k, float* __restrict__ out) {
    const int kk = k * k;
    const int half = k / 2;
    for (long long p = blockIdx.x; p < n_pts; p += gridDim.x) {
        // np.rint == round-half-to-even == rint() in the default rounding mode
        const int cx = (int)rint(pts[2 * p]);
        const int cy = (int)rint(pts[2 * p + 1]);
        const int x0 = cx - half, y0 = cy - half;
        float* dst = out + p * (long long)kk;
        const bool interior = x0 >= 0 && y0 >= 0 && x0 + k <= W && y0 + k <= H;
        if (interior) {
            const float* src = img + (long long)y0 * W + x0;
            for (int e = threadIdx.x; e < kk; e += kThreads) {
                const int r = e / k, c = e - r * k;
                dst[e] = __ldg(src + (long long)r * W + c);
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

}  // namespace zb200

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
