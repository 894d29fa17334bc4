// Clustering of the feature matrix ("next" row f4 of SURVEY section 8): the consumers of ZPs features.
// Replaces the passes over the (N, d) matrix inside mtflearn.clustering.kmeans_lbs / gmm_lbs
// (mtflearn/clustering/_clustering_functions.py:8-34 = scikit-learn KMeans(n, random_state=0) and
// GaussianMixture(n, 'full', random_state=0); scikit-learn 1.9 is the third-party dependency whose published
// algorithm is restated: k-means++ seeding (sklearn/cluster/_kmeans.py:_kmeans_plusplus), Lloyd iterations
// (_kmeans_single_lloyd) and EM with full covariances (sklearn/mixture/_gaussian_mixture.py)).
// Division of labour: everything that touches all N samples runs here, in float64 on the float32 feature matrix
// (mean-centred on the fly, like scikit-learn centres X); the k x d / k x d x d parameter updates, the random
// draws (the reference's own numpy RandomState stream) and the d x d Cholesky factors are host glue.
// All reductions are deterministic: per-slab partials in a fixed order, summed in slab order.
#include "zb200_common.cuh"

#include <vector>

namespace zb200 {

constexpr int KC_SLAB = 2048;      // samples per partial-sum slab

// squared distances of every (centred) sample to t candidate centres, min-ed with the running closest distance:
//   out[j][i] = min(closest[i], |x_i - mean - cand_j|^2)      closest == nullptr: no min (first centre)
__global__ void kmeans_mindist_kernel(const float* __restrict__ x, long long n, int d, const double* __restrict__ mean,
                                      const double* __restrict__ cand, int t, const double* __restrict__ closest,
                                      double* __restrict__ out) {
    extern __shared__ double sh[];                 // [d] mean, [t][d] candidates
    for (int e = threadIdx.x; e < d; e += blockDim.x) sh[e] = mean[e];
    for (int e = threadIdx.x; e < t * d; e += blockDim.x) sh[d + e] = cand[e];
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* xr = x + i * d;
    for (int j = 0; j < t; ++j) {
        const double* c = sh + d + (size_t)j * d;
        double s = 0.0;
        for (int f = 0; f < d; ++f) {
            const double v = ((double)__ldg(xr + f) - sh[f]) - c[f];
            s = fma(v, v, s);
        }
        if (closest) s = fmin(s, closest[i]);
        out[(size_t)j * n + i] = s;
    }
}

// per-slab sums of t rows of length n (potentials of the k-means++ candidates), slab order is fixed
__global__ void rows_slab_sum_kernel(const double* __restrict__ v, long long n, int t, double* __restrict__ partial) {
    __shared__ double red[256];
    const int j = blockIdx.y;
    const long long b = (long long)blockIdx.x * KC_SLAB;
    const long long e = b + KC_SLAB < n ? b + KC_SLAB : n;
    double s = 0.0;
    for (long long i = b + threadIdx.x; i < e; i += blockDim.x) s += v[(size_t)j * n + i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = blockDim.x / 2; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(size_t)j * gridDim.x + blockIdx.x] = red[0];
}

// E-step of Lloyd: label = argmin_c |x - mean - centre_c|^2 (first minimum), count of labels that changed
__global__ void kmeans_assign_kernel(const float* __restrict__ x, long long n, int d, const double* __restrict__ mean,
                                     const double* __restrict__ centres, int k, int32_t* __restrict__ labels,
                                     unsigned long long* __restrict__ changed) {
    extern __shared__ double sh[];                 // [d] mean, [k][d] centres
    for (int e = threadIdx.x; e < d; e += blockDim.x) sh[e] = mean[e];
    for (int e = threadIdx.x; e < k * d; e += blockDim.x) sh[d + e] = centres[e];
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool diff = false;
    if (i < n) {
        const float* xr = x + i * d;
        double best = 0.0;
        int arg = 0;
        for (int c = 0; c < k; ++c) {
            const double* ctr = sh + d + (size_t)c * d;
            double s = 0.0;
            for (int f = 0; f < d; ++f) {
                const double v = ((double)__ldg(xr + f) - sh[f]) - ctr[f];
                s = fma(v, v, s);
            }
            if (c == 0 || s < best) { best = s; arg = c; }
        }
        diff = labels[i] != arg;
        labels[i] = arg;
    }
    const unsigned m = __ballot_sync(0xffffffffu, diff);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(changed, (unsigned long long)__popc(m));
}

// M-step of Lloyd: per-slab, per-cluster sums of the centred samples and counts.  Thread f owns column f of the
// slab's [k][d] accumulator in shared memory (no atomics: the order of the additions is the sample order).
__global__ void kmeans_sums_kernel(const float* __restrict__ x, long long n, int d, const double* __restrict__ mean,
                                   const int32_t* __restrict__ labels, int k, double* __restrict__ partial,
                                   double* __restrict__ partial_count) {
    extern __shared__ double acc[];                // [k][d] + [k]
    double* cnt = acc + (size_t)k * d;
    for (int e = threadIdx.x; e < k * d + k; e += blockDim.x) acc[e] = 0.0;
    __syncthreads();
    const long long b = (long long)blockIdx.x * KC_SLAB;
    const long long e = b + KC_SLAB < n ? b + KC_SLAB : n;
    for (int f = threadIdx.x; f < d; f += blockDim.x) {
        const double mu = mean[f];
        for (long long i = b; i < e; ++i) {
            const int c = labels[i];
            acc[(size_t)c * d + f] += (double)__ldg(x + i * d + f) - mu;
        }
    }
    if (threadIdx.x == 0)
        for (long long i = b; i < e; ++i) cnt[labels[i]] += 1.0;
    __syncthreads();
    for (int q = threadIdx.x; q < k * d; q += blockDim.x) partial[(size_t)blockIdx.x * k * d + q] = acc[q];
    for (int q = threadIdx.x; q < k; q += blockDim.x) partial_count[(size_t)blockIdx.x * k + q] = cnt[q];
}

// ---- Gaussian mixture, full covariances -------------------------------------------------------------------------
// E-step: y = (x - mu_c) . P_c (P_c = precision Cholesky factor, d x d, upper triangular in the sklearn convention:
// prec = P P^T), log N = -(d log 2pi + |y|^2)/2 + log_det_c; weighted by log w_c; log-sum-exp over components.
// Writes log_resp [n][k] (or labels only) and the per-sample log-probability norm.
__global__ void gmm_estep_kernel(const float* __restrict__ x, long long n, int d, const double* __restrict__ tab, int k,
                                 double* __restrict__ log_resp, double* __restrict__ log_norm, int32_t* __restrict__ labels) {
    // tab: [k] log weights, [k] log dets, [k][d] means, [k][d][d] precision factors (row-major P[c][r][col])
    extern __shared__ double sh[];
    const size_t tab_len = (size_t)2 * k + (size_t)k * d + (size_t)k * d * d;
    for (size_t e = threadIdx.x; e < tab_len; e += blockDim.x) sh[e] = tab[e];
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* logw = sh;
    const double* logdet = sh + k;
    const double* mu = sh + 2 * k;
    const double* prec = mu + (size_t)k * d;
    const float* xr = x + i * d;
    const double c0 = -0.5 * d * 1.8378770664093453;       // log(2 pi)
    double best = 0.0, run_max = 0.0, run_sum = 0.0;
    int arg = 0;
    for (int c = 0; c < k; ++c) {
        const double* m = mu + (size_t)c * d;
        const double* p = prec + (size_t)c * d * d;
        double q = 0.0;
        for (int col = 0; col < d; ++col) {
            double y = 0.0;
            for (int r = 0; r <= col; ++r) y = fma((double)__ldg(xr + r) - m[r], p[(size_t)r * d + col], y);   // upper triangular
            q = fma(y, y, q);
        }
        const double w = c0 - 0.5 * q + logdet[c] + logw[c];
        if (log_resp) log_resp[(size_t)i * k + c] = w;
        if (c == 0 || w > best) { best = w; arg = c; }
        if (c == 0) { run_max = w; run_sum = 1.0; }
        else if (w > run_max) { run_sum = run_sum * exp(run_max - w) + 1.0; run_max = w; }
        else run_sum += exp(w - run_max);
    }
    const double norm = run_max + log(run_sum);
    if (log_norm) log_norm[i] = norm;
    if (labels) labels[i] = arg;
    if (log_resp)
        for (int c = 0; c < k; ++c) log_resp[(size_t)i * k + c] -= norm;
}

// M-step accumulators per slab: nk[c] = sum r, sx[c][f] = sum r x_f, sxx[c][a][b] = sum r x_a x_b (x centred by the
// global mean, r = exp(log_resp) or the one-hot of a label).  Thread (a, b) owns its accumulators (fixed order).
__global__ void gmm_mstep_kernel(const float* __restrict__ x, long long n, int d, const double* __restrict__ mean,
                                 const double* __restrict__ log_resp, const int32_t* __restrict__ labels, int k, int c,
                                 double* __restrict__ partial) {
    // one launch per component c; partial: [slab][1 + d + d*d]
    extern __shared__ double sh[];                 // [rows][d] centred samples, [rows] weights
    constexpr int kRows = 32;
    double* sx = sh;
    double* sw = sh + (size_t)kRows * d;
    const long long b = (long long)blockIdx.x * KC_SLAB;
    const long long e = b + KC_SLAB < n ? b + KC_SLAB : n;
    const int per = (d * d + blockDim.x - 1) / blockDim.x;          // (a, b) pairs per thread, <= 8
    double acc[8], acc1 = 0.0, acc0 = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.0;
    for (long long r0 = b; r0 < e; r0 += kRows) {
        const int nr = (int)(e - r0 < kRows ? e - r0 : kRows);
        for (int q = threadIdx.x; q < nr * d; q += blockDim.x) {
            const int r = q / d, f = q - r * d;
            sx[(size_t)r * d + f] = (double)__ldg(x + (r0 + r) * d + f) - mean[f];
        }
        for (int r = threadIdx.x; r < nr; r += blockDim.x)
            sw[r] = log_resp ? exp(log_resp[(size_t)(r0 + r) * k + c]) : (labels[r0 + r] == c ? 1.0 : 0.0);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int pair = threadIdx.x + q * blockDim.x;
            if (q < per && pair < d * d) {
                const int a = pair / d, bb = pair - a * d;
                double s = acc[q];
                for (int r = 0; r < nr; ++r) s = fma(sw[r] * sx[(size_t)r * d + a], sx[(size_t)r * d + bb], s);
                acc[q] = s;
            }
        }
        if ((int)threadIdx.x < d) {
            double s = acc1;
            for (int r = 0; r < nr; ++r) s = fma(sw[r], sx[(size_t)r * d + threadIdx.x], s);
            acc1 = s;
        }
        if (threadIdx.x == blockDim.x - 1) {
            double s = acc0;
            for (int r = 0; r < nr; ++r) s += sw[r];
            acc0 = s;
        }
        __syncthreads();
    }
    double* dst = partial + (size_t)blockIdx.x * (1 + d + (size_t)d * d);
    if (threadIdx.x == blockDim.x - 1) dst[0] = acc0;
    if ((int)threadIdx.x < d) dst[1 + threadIdx.x] = acc1;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int pair = threadIdx.x + q * blockDim.x;
        if (q < per && pair < d * d) dst[1 + d + pair] = acc[q];
    }
}

// out[e] = sum over slabs in slab order
__global__ void slab_reduce_kernel(const double* __restrict__ partial, int n_slabs, long long elems, double* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= elems) return;
    double s = 0.0;
    for (int q = 0; q < n_slabs; ++q) s += partial[(size_t)q * elems + e];
    out[e] = s;
}

struct DevBuf {
    void* p = nullptr;
    cudaStream_t s;
    explicit DevBuf(cudaStream_t st) : s(st) {}
    ~DevBuf() { if (p) cudaFreeAsync(p, s); }
    int alloc(size_t bytes) {
        cudaError_t e = scratch_alloc(&p, bytes ? bytes : 16, s);
        if (e != cudaSuccess) { p = nullptr; set_error("cluster: scratch_alloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return ZB200_ENOMEM; }
        return ZB200_OK;
    }
    int upload(const void* h, size_t bytes) {
        int rc = alloc(bytes);
        if (rc) return rc;
        ZB_CUDA(cudaMemcpyAsync(p, h, bytes, cudaMemcpyHostToDevice, s));
        return ZB200_OK;
    }
};

static int n_slabs_of(int64_t n) { return (int)ceil_div(n, KC_SLAB); }

}  // namespace zb200

using namespace zb200;

extern "C" int zb200_kmeans_mindist_f32(const float* d_x, int64_t n, int d, const double* h_mean, const double* h_cand, int t,
                                        const double* d_closest, double* d_out, double* h_pot, void* stream) {
    ZB_CHECK_ARG(d_x && h_mean && h_cand && d_out && h_pot, "kmeans_mindist: null pointer");
    ZB_CHECK_ARG(n >= 1 && d >= 1 && d <= 1024 && t >= 1 && t <= 64, "kmeans_mindist: bad shape n=%lld d=%d t=%d", (long long)n, d, t);
    cudaStream_t s = as_stream(stream);
    std::vector<double> tab((size_t)d + (size_t)t * d);
    for (int f = 0; f < d; ++f) tab[f] = h_mean[f];
    for (size_t e = 0; e < (size_t)t * d; ++e) tab[d + e] = h_cand[e];
    DevBuf dtab(s), part(s);
    int rc = dtab.upload(tab.data(), tab.size() * sizeof(double));
    if (rc) return rc;
    const size_t smem = tab.size() * sizeof(double);
    ZB_CHECK_ARG(smem <= 200 * 1024, "kmeans_mindist: %d candidates of %d features do not fit shared memory", t, d);
    ZB_CUDA(cudaFuncSetAttribute(kmeans_mindist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const double* dt = static_cast<const double*>(dtab.p);
    kmeans_mindist_kernel<<<(unsigned)ceil_div(n, 128), 128, smem, s>>>(d_x, (long long)n, d, dt, dt + d, t, d_closest, d_out);
    ZB_LAUNCHED();
    const int slabs = n_slabs_of(n);
    rc = part.alloc(sizeof(double) * (size_t)t * slabs);
    if (rc) return rc;
    rows_slab_sum_kernel<<<dim3((unsigned)slabs, (unsigned)t), 256, 0, s>>>(d_out, (long long)n, t, static_cast<double*>(part.p));
    ZB_LAUNCHED();
    std::vector<double> hp((size_t)t * slabs);
    ZB_CUDA(cudaMemcpyAsync(hp.data(), part.p, hp.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    ZB_CUDA(cudaStreamSynchronize(s));
    for (int j = 0; j < t; ++j) {
        double acc = 0.0;
        for (int q = 0; q < slabs; ++q) acc += hp[(size_t)j * slabs + q];
        h_pot[j] = acc;
    }
    return ZB200_OK;
}

extern "C" int zb200_kmeans_step_f32(const float* d_x, int64_t n, int d, const double* h_mean, const double* h_centres, int k,
                                     int32_t* d_labels, int update, double* h_sums, double* h_counts, int64_t* h_changed,
                                     void* stream) {
    ZB_CHECK_ARG(d_x && h_mean && h_centres && d_labels && h_changed, "kmeans_step: null pointer");
    ZB_CHECK_ARG(n >= 1 && d >= 1 && d <= 1024 && k >= 1 && k <= 256, "kmeans_step: bad shape n=%lld d=%d k=%d", (long long)n, d, k);
    ZB_CHECK_ARG(!update || (h_sums && h_counts), "kmeans_step: update needs h_sums / h_counts");
    cudaStream_t s = as_stream(stream);
    std::vector<double> tab((size_t)d + (size_t)k * d);
    for (int f = 0; f < d; ++f) tab[f] = h_mean[f];
    for (size_t e = 0; e < (size_t)k * d; ++e) tab[d + e] = h_centres[e];
    DevBuf dtab(s), dchg(s), part(s), red(s);
    int rc = dtab.upload(tab.data(), tab.size() * sizeof(double));
    if (rc) return rc;
    rc = dchg.alloc(sizeof(unsigned long long));
    if (rc) return rc;
    ZB_CUDA(cudaMemsetAsync(dchg.p, 0, sizeof(unsigned long long), s));
    const size_t smem = tab.size() * sizeof(double);
    ZB_CHECK_ARG(smem + (size_t)k * 8 <= 200 * 1024, "kmeans_step: %d centres of %d features do not fit shared memory", k, d);
    ZB_CUDA(cudaFuncSetAttribute(kmeans_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const double* dt = static_cast<const double*>(dtab.p);
    kmeans_assign_kernel<<<(unsigned)ceil_div(n, 128), 128, smem, s>>>(d_x, (long long)n, d, dt, dt + d, k, d_labels,
                                                                        static_cast<unsigned long long*>(dchg.p));
    ZB_LAUNCHED();
    unsigned long long chg = 0;
    ZB_CUDA(cudaMemcpyAsync(&chg, dchg.p, sizeof(chg), cudaMemcpyDeviceToHost, s));
    if (update) {
        const int slabs = n_slabs_of(n);
        const size_t per = (size_t)k * d + k;
        rc = part.alloc(sizeof(double) * per * slabs);
        if (rc) return rc;
        rc = red.alloc(sizeof(double) * per);
        if (rc) return rc;
        double* pp = static_cast<double*>(part.p);
        const size_t smem2 = per * sizeof(double);
        ZB_CUDA(cudaFuncSetAttribute(kmeans_sums_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        kmeans_sums_kernel<<<(unsigned)slabs, 128, smem2, s>>>(d_x, (long long)n, d, dt, d_labels, k, pp, pp + (size_t)slabs * k * d);
        ZB_LAUNCHED();
        double* rr = static_cast<double*>(red.p);
        slab_reduce_kernel<<<(unsigned)ceil_div((long long)k * d, 128), 128, 0, s>>>(pp, slabs, (long long)k * d, rr);
        ZB_LAUNCHED();
        slab_reduce_kernel<<<1, 256, 0, s>>>(pp + (size_t)slabs * k * d, slabs, (long long)k, rr + (size_t)k * d);
        ZB_LAUNCHED();
        std::vector<double> host(per);
        ZB_CUDA(cudaMemcpyAsync(host.data(), rr, per * sizeof(double), cudaMemcpyDeviceToHost, s));
        ZB_CUDA(cudaStreamSynchronize(s));
        for (size_t e = 0; e < (size_t)k * d; ++e) h_sums[e] = host[e];
        for (int c = 0; c < k; ++c) h_counts[c] = host[(size_t)k * d + c];
    } else {
        ZB_CUDA(cudaStreamSynchronize(s));
    }
    *h_changed = (int64_t)chg;
    return ZB200_OK;
}

extern "C" int zb200_gmm_estep_f32(const float* d_x, int64_t n, int d, const double* h_log_weights, const double* h_log_dets,
                                   const double* h_means, const double* h_prec_chol, int k, double* d_log_resp,
                                   int32_t* d_labels, double* h_mean_log_norm, void* stream) {
    ZB_CHECK_ARG(d_x && h_log_weights && h_log_dets && h_means && h_prec_chol, "gmm_estep: null pointer");
    ZB_CHECK_ARG(n >= 1 && d >= 1 && k >= 1 && k <= 64, "gmm_estep: bad shape n=%lld d=%d k=%d", (long long)n, d, k);
    cudaStream_t s = as_stream(stream);
    const size_t tab_len = (size_t)2 * k + (size_t)k * d + (size_t)k * d * d;
    ZB_CHECK_ARG(tab_len * sizeof(double) <= 200 * 1024, "gmm_estep: %d components of %d features do not fit shared memory", k, d);
    std::vector<double> tab(tab_len);
    for (int c = 0; c < k; ++c) { tab[c] = h_log_weights[c]; tab[k + c] = h_log_dets[c]; }
    for (size_t e = 0; e < (size_t)k * d; ++e) tab[2 * k + e] = h_means[e];
    for (size_t e = 0; e < (size_t)k * d * d; ++e) tab[2 * k + (size_t)k * d + e] = h_prec_chol[e];
    DevBuf dtab(s), norm(s), part(s);
    int rc = dtab.upload(tab.data(), tab_len * sizeof(double));
    if (rc) return rc;
    if (h_mean_log_norm) { rc = norm.alloc(sizeof(double) * (size_t)n); if (rc) return rc; }
    ZB_CUDA(cudaFuncSetAttribute(gmm_estep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(tab_len * sizeof(double))));
    gmm_estep_kernel<<<(unsigned)ceil_div(n, 128), 128, tab_len * sizeof(double), s>>>(
        d_x, (long long)n, d, static_cast<const double*>(dtab.p), k, d_log_resp, static_cast<double*>(norm.p), d_labels);
    ZB_LAUNCHED();
    if (h_mean_log_norm) {
        const int slabs = n_slabs_of(n);
        rc = part.alloc(sizeof(double) * slabs);
        if (rc) return rc;
        rows_slab_sum_kernel<<<dim3((unsigned)slabs, 1), 256, 0, s>>>(static_cast<const double*>(norm.p), (long long)n, 1,
                                                                     static_cast<double*>(part.p));
        ZB_LAUNCHED();
        std::vector<double> hp(slabs);
        ZB_CUDA(cudaMemcpyAsync(hp.data(), part.p, sizeof(double) * slabs, cudaMemcpyDeviceToHost, s));
        ZB_CUDA(cudaStreamSynchronize(s));
        double acc = 0.0;
        for (int q = 0; q < slabs; ++q) acc += hp[q];
        *h_mean_log_norm = acc / (double)n;
    }
    return ZB200_OK;
}

extern "C" int zb200_gmm_mstep_f32(const float* d_x, int64_t n, int d, const double* h_mean, const double* d_log_resp,
                                   const int32_t* d_labels, int k, double* h_nk, double* h_sx, double* h_sxx, void* stream) {
    ZB_CHECK_ARG(d_x && h_mean && (d_log_resp || d_labels) && h_nk && h_sx && h_sxx, "gmm_mstep: null pointer");
    ZB_CHECK_ARG(n >= 1 && d >= 1 && d <= 90 && k >= 1 && k <= 64, "gmm_mstep: bad shape n=%lld d=%d k=%d (d <= 90)", (long long)n, d, k);
    cudaStream_t s = as_stream(stream);
    const int slabs = n_slabs_of(n);
    const size_t per = 1 + (size_t)d + (size_t)d * d;
    DevBuf dmean(s), part(s), red(s);
    int rc = dmean.upload(h_mean, sizeof(double) * d);
    if (rc) return rc;
    rc = part.alloc(sizeof(double) * per * slabs);
    if (rc) return rc;
    rc = red.alloc(sizeof(double) * per * k);
    if (rc) return rc;
    int threads = ((d * d + 7) / 8 + 31) / 32 * 32;                  // <= 8 (a, b) pairs per thread
    if (threads < 64) threads = 64;
    if (threads < d + 1) threads = (d + 1 + 31) / 32 * 32;
    const size_t smem = sizeof(double) * (32 * (size_t)d + 32);
    ZB_CUDA(cudaFuncSetAttribute(gmm_mstep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int c = 0; c < k; ++c) {
        gmm_mstep_kernel<<<(unsigned)slabs, threads, smem, s>>>(d_x, (long long)n, d, static_cast<const double*>(dmean.p),
                                                                d_log_resp, d_labels, k, c, static_cast<double*>(part.p));
        ZB_LAUNCHED();
        slab_reduce_kernel<<<(unsigned)ceil_div((long long)per, 128), 128, 0, s>>>(static_cast<const double*>(part.p), slabs,
                                                                                  (long long)per, static_cast<double*>(red.p) + (size_t)c * per);
        ZB_LAUNCHED();
    }
    std::vector<double> host(per * k);
    ZB_CUDA(cudaMemcpyAsync(host.data(), red.p, sizeof(double) * per * k, cudaMemcpyDeviceToHost, s));
    ZB_CUDA(cudaStreamSynchronize(s));
    for (int c = 0; c < k; ++c) {
        const double* src = host.data() + (size_t)c * per;
        h_nk[c] = src[0];
        for (int f = 0; f < d; ++f) h_sx[(size_t)c * d + f] = src[1 + f];
        for (size_t e = 0; e < (size_t)d * d; ++e) h_sxx[(size_t)c * d * d + e] = src[1 + d + e];
    }
    return ZB200_OK;
}
