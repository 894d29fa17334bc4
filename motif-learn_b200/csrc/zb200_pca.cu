// PCA of the feature matrix ("next" row f4 of SURVEY section 8): the first consumer of ZPs features.
// Replaces mtflearn.features.pca (mtflearn/features/_dimension_reduction.py:3-6 = sklearn PCA(n_components)
// .fit_transform; for n_samples >= 10 n_features scikit-learn takes the covariance route: C = (X^T X - n mu mu^T)
// / (n - 1), eigh(C), components sign-fixed by svd_flip(u_based_decision=False), scores (X - mu) V^T).
// Device side = the two passes over the (N, M) matrix, both HBM-bound:
//   gram_kernel     G = X^T X and column sums, accumulated in float64 (deterministic: per-CTA partials in a
//                   fixed order, reduced by gram_reduce_kernel)
//   scores_kernel   out[i, c] = sum_k (x[i, k] - mu[k]) V[c, k]  in float64
// The M x M symmetric eigenproblem (M <= a few hundred) is host-side glue, as it is in scikit-learn.
#include "zb200_common.cuh"

namespace zb200 {

constexpr int PG_T = 16;           // 16 x 16 threads
constexpr int PG_B = 6;            // each thread owns a 6 x 6 block of a 96 x 96 output tile
constexpr int PG_TILE = PG_T * PG_B;
constexpr int PG_ROWS = 32;        // rows of X staged per step

// grid: (slabs, tiles_a, tiles_b); every CTA walks rows slab, slab + gridDim.x, ... in chunks of PG_ROWS
__global__ void __launch_bounds__(PG_T * PG_T)
gram_kernel(const float* __restrict__ x, long long n, int m, long long rows_per_cta, double* __restrict__ partial,
            double* __restrict__ partial_sum) {
    __shared__ float sa[PG_ROWS][PG_TILE + 1], sb[PG_ROWS][PG_TILE + 1];
    const int ta = blockIdx.y * PG_TILE, tb = blockIdx.z * PG_TILE;
    const int ti = threadIdx.x / PG_T, tj = threadIdx.x % PG_T;
    double acc[PG_B][PG_B];
    double csum[PG_B];
#pragma unroll
    for (int a = 0; a < PG_B; ++a) {
        csum[a] = 0.0;
#pragma unroll
        for (int b = 0; b < PG_B; ++b) acc[a][b] = 0.0;
    }
    const long long r_begin = (long long)blockIdx.x * rows_per_cta;
    const long long r_end = r_begin + rows_per_cta < n ? r_begin + rows_per_cta : n;
    for (long long r0 = r_begin; r0 < r_end; r0 += PG_ROWS) {
        const int nr = (int)(r_end - r0 < PG_ROWS ? r_end - r0 : PG_ROWS);
        for (int e = threadIdx.x; e < PG_ROWS * PG_TILE; e += PG_T * PG_T) {
            const int r = e / PG_TILE, c = e - r * PG_TILE;
            const bool live = r < nr;
            sa[r][c] = (live && ta + c < m) ? __ldg(x + (r0 + r) * m + ta + c) : 0.f;
            sb[r][c] = (live && tb + c < m) ? __ldg(x + (r0 + r) * m + tb + c) : 0.f;
        }
        __syncthreads();
        for (int r = 0; r < nr; ++r) {
            double va[PG_B], vb[PG_B];
#pragma unroll
            for (int a = 0; a < PG_B; ++a) va[a] = (double)sa[r][ti * PG_B + a];
#pragma unroll
            for (int b = 0; b < PG_B; ++b) vb[b] = (double)sb[r][tj * PG_B + b];
#pragma unroll
            for (int a = 0; a < PG_B; ++a)
#pragma unroll
                for (int b = 0; b < PG_B; ++b) acc[a][b] = fma(va[a], vb[b], acc[a][b]);
            if (blockIdx.z == 0 && tj == 0) {
#pragma unroll
                for (int a = 0; a < PG_B; ++a) csum[a] += va[a];
            }
        }
        __syncthreads();
    }
    double* dst = partial + (size_t)blockIdx.x * m * m;
#pragma unroll
    for (int a = 0; a < PG_B; ++a)
#pragma unroll
        for (int b = 0; b < PG_B; ++b) {
            const int ia = ta + ti * PG_B + a, ib = tb + tj * PG_B + b;
            if (ia < m && ib < m) dst[(size_t)ia * m + ib] = acc[a][b];
        }
    if (blockIdx.z == 0 && tj == 0) {
#pragma unroll
        for (int a = 0; a < PG_B; ++a) {
            const int ia = ta + ti * PG_B + a;
            if (ia < m) partial_sum[(size_t)blockIdx.x * m + ia] = csum[a];
        }
    }
}

// out[e] = sum over slabs, in slab order (deterministic)
__global__ void gram_reduce_kernel(const double* __restrict__ partial, int n_slabs, long long elems, double* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= elems) return;
    double s = 0.0;
    for (int k = 0; k < n_slabs; ++k) s += partial[(size_t)k * elems + e];
    out[e] = s;
}

// tab: [mu (m) | V (c x m)] doubles on the device; one thread per (row, component)
__global__ void scores_kernel(const float* __restrict__ x, long long n, int m, const double* __restrict__ tab, int n_comp,
                              double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * n_comp) return;
    const long long row = i / n_comp;
    const int c = (int)(i - row * n_comp);
    const double* mu = tab;
    const double* v = tab + m + (size_t)c * m;
    const float* xr = x + row * m;
    double s = 0.0;
    for (int k = 0; k < m; ++k) s = fma((double)__ldg(xr + k) - mu[k], v[k], s);
    out[i] = s;
}

}  // namespace zb200

using namespace zb200;

extern "C" int zb200_gram_f32(const float* d_x, int64_t n, int m, double* d_gram, double* d_colsum, void* stream) {
    ZB_CHECK_ARG(d_x && d_gram && d_colsum, "gram: null pointer");
    ZB_CHECK_ARG(n >= 1 && m >= 1 && m <= 1024, "gram: bad shape n=%lld m=%d", (long long)n, m);
    cudaStream_t s = as_stream(stream);
    int dev = 0, sms = 148;
    ZB_CUDA(cudaGetDevice(&dev));
    ZB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int tiles = (int)ceil_div(m, PG_TILE);
    long long slabs = (long long)sms * 4 / (tiles * tiles);
    if (slabs < 1) slabs = 1;
    if (slabs > ceil_div(n, PG_ROWS)) slabs = ceil_div(n, PG_ROWS);
    const long long rows_per_cta = round_up((int)ceil_div(n, slabs), PG_ROWS);
    slabs = ceil_div(n, rows_per_cta);
    double* partial = nullptr;
    const size_t mm = (size_t)m * m;
    ZB_CUDA(scratch_alloc(&partial, sizeof(double) * (size_t)slabs * (mm + m), s));
    double* partial_sum = partial + (size_t)slabs * mm;
    gram_kernel<<<dim3((unsigned)slabs, (unsigned)tiles, (unsigned)tiles), PG_T * PG_T, 0, s>>>(d_x, (long long)n, m, rows_per_cta,
                                                                                             partial, partial_sum);
    gram_reduce_kernel<<<(unsigned)ceil_div((long long)mm, 256), 256, 0, s>>>(partial, (int)slabs, (long long)mm, d_gram);
    gram_reduce_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, s>>>(partial_sum, (int)slabs, (long long)m, d_colsum);
    g_launches.fetch_add(3, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(partial, s);
    if (e != cudaSuccess) { set_error("gram kernels failed: %s", cudaGetErrorString(e)); return ZB200_ECUDA; }
    return ZB200_OK;
}

extern "C" int zb200_pca_scores_f32(const float* d_x, int64_t n, int m, const double* h_mean, const double* h_components,
                                    int n_comp, double* d_out, void* stream) {
    ZB_CHECK_ARG(d_x && h_mean && h_components && d_out, "pca_scores: null pointer");
    ZB_CHECK_ARG(n >= 1 && m >= 1 && n_comp >= 1 && n_comp <= m, "pca_scores: bad shape");
    cudaStream_t s = as_stream(stream);
    double* tab = nullptr;
    const size_t words = (size_t)m * (n_comp + 1);
    ZB_CUDA(scratch_alloc(&tab, sizeof(double) * words, s));
    ZB_CUDA(cudaMemcpyAsync(tab, h_mean, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    ZB_CUDA(cudaMemcpyAsync(tab + m, h_components, sizeof(double) * (size_t)m * n_comp, cudaMemcpyHostToDevice, s));
    scores_kernel<<<(unsigned)ceil_div((long long)n * n_comp, 256), 256, 0, s>>>(d_x, (long long)n, m, tab, n_comp, d_out);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(tab, s);
    if (e != cudaSuccess) { set_error("pca scores kernel failed: %s", cudaGetErrorString(e)); return ZB200_ECUDA; }
    return ZB200_OK;
}
