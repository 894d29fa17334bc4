// zmoments algebra on device-resident moment arrays: real<->complex packing, normalisation,
// mode selection, rotation, n-fold and mirror scores.  Replaces the numpy passes of
// mtflearn/features/_zmoments.py:300-493.  All HBM-bound element-wise / small-reduction work:
// one thread per item (patch or pixel), items of a warp are consecutive so the planar (M,H,W)
// layout is fully coalesced; the (N,M) layout walks a row per thread through L1.
#include "zb200_common.cuh"

#include <string.h>
#include <vector>

namespace zb200 {

// stream-ordered scratch holding small host tables on the device for one call
struct Scratch {
    void* ptr = nullptr;
    cudaStream_t s;
    explicit Scratch(cudaStream_t st) : s(st) {}
    ~Scratch() { if (ptr) cudaFreeAsync(ptr, s); }
    int upload(const void* host, size_t bytes) {
        ZB_CUDA(scratch_alloc(&ptr, bytes ? bytes : 1, s));
        if (bytes) ZB_CUDA(cudaMemcpyAsync(ptr, host, bytes, cudaMemcpyHostToDevice, s));
        return ZB200_OK;
    }
};

template <typename T> struct Cx { T re, im; };

static inline unsigned grid_for(int64_t n, int threads) { return (unsigned)ceil_div(n, threads); }

// ---- real -> complex: Zc[c] = Z[pos[c]] + i Z[neg[c]]  (_zmoments.py:111-132, 300-316) -------
// Work mapping of the three re-indexing kernels.  kRow = false: one thread per item, looping over the modes --
// coalesced for the planar (M,H,W) layout, where items of a warp are consecutive pixels.  kRow = true: one thread
// per (item, output mode), modes fastest -- coalesced for the row-major (N,M) layout of patch moments (a thread
// per item would walk its own row there: 0.6 TB/s instead of HBM speed).
template <bool kRow>
__device__ __forceinline__ bool map_work(long long n_items, int n_out, long long& it, int& q0, int& q1) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (kRow) {
        it = i / n_out;
        q0 = (int)(i - it * n_out);
        q1 = q0 + 1;
    } else {
        it = i;
        q0 = 0;
        q1 = n_out;
    }
    return it < n_items;
}

template <typename T, bool kRow>
__global__ void to_complex_kernel(const T* __restrict__ in, long long n_items, long long iis, long long ims,
                                  const int* __restrict__ pos, const int* __restrict__ neg, int n_c,
                                  Cx<T>* __restrict__ out, long long ois, long long oms) {
    long long it;
    int c0, c1;
    if (!map_work<kRow>(n_items, n_c, it, c0, c1)) return;
    for (int c = c0; c < c1; ++c) {
        Cx<T> v;
        v.re = pos[c] >= 0 ? in[it * iis + pos[c] * ims] : (T)0;
        v.im = neg[c] >= 0 ? in[it * iis + neg[c] * ims] : (T)0;
        out[it * ois + c * oms] = v;
    }
}

// ---- complex -> real: Z[j] = Re or Im of Zc[src[j]]  (_zmoments.py:134-196, 318-341) ---------
template <typename T, bool kRow>
__global__ void to_real_kernel(const Cx<T>* __restrict__ in, long long n_items, long long iis, long long ims,
                               const int* __restrict__ src, const unsigned char* __restrict__ take_im, int n_r,
                               T* __restrict__ out, long long ois, long long oms) {
    long long it;
    int j0, j1;
    if (!map_work<kRow>(n_items, n_r, it, j0, j1)) return;
    for (int j = j0; j < j1; ++j) {
        const Cx<T> v = in[it * iis + src[j] * ims];
        out[it * ois + j * oms] = take_im[j] ? v.im : v.re;
    }
}

// ---- select: out[q] = in[index[q]]  (_zmoments.py:359-374) -----------------------------------
template <typename E, bool kRow>
__global__ void select_kernel(const E* __restrict__ in, long long n_items, long long iis, long long ims,
                              const int* __restrict__ index, int n_out, E* __restrict__ out, long long ois,
                              long long oms) {
    long long it;
    int q0, q1;
    if (!map_work<kRow>(n_items, n_out, it, q0, q1)) return;
    for (int q = q0; q < q1; ++q) out[it * ois + q * oms] = in[it * iis + index[q] * ims];
}

// ---- normalize: x / ||x||_p over modes  (_zmoments.py:344-356, np.linalg.norm semantics) ----
template <typename T> __device__ __forceinline__ T mag(T v) { return fabs(v); }
template <typename T> __device__ __forceinline__ T mag(Cx<T> v) { return hypot(v.re, v.im); }
template <typename T> __device__ __forceinline__ T scale(T v, T s) { return v / s; }
template <typename T> __device__ __forceinline__ Cx<T> scale(Cx<T> v, T s) { return Cx<T>{v.re / s, v.im / s}; }

template <typename T>
__device__ __forceinline__ void norm_accum(T a, int kind, double p, T& acc) {
    switch (kind) {
        case ZB200_NORM_L1: acc += a; break;
        case ZB200_NORM_L2: acc += a * a; break;
        case ZB200_NORM_INF: acc = a > acc ? a : acc; break;
        default: acc += (T)pow((double)a, p); break;
    }
}
template <typename T>
__device__ __forceinline__ T norm_finish(T acc, int kind, double p) {
    switch (kind) {
        case ZB200_NORM_L1: return acc;
        case ZB200_NORM_L2: return sqrt(acc);
        case ZB200_NORM_INF: return acc;
        default: return (T)pow((double)acc, 1.0 / p);
    }
}

template <typename T, typename E>
__global__ void normalize_kernel(const E* __restrict__ in, long long n_items, int n_modes, long long is,
                                 long long ms, int kind, double p, E* __restrict__ out) {
    const long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= n_items) return;
    T acc = 0;
    for (int j = 0; j < n_modes; ++j) norm_accum<T>(mag(in[it * is + j * ms]), kind, p, acc);
    const T nrm = norm_finish<T>(acc, kind, p);
    for (int j = 0; j < n_modes; ++j) out[it * is + j * ms] = scale(in[it * is + j * ms], nrm);
}

// ---- rotate: Zc * exp(-i m theta)  (_zmoments.py:377-418) ------------------------------------
template <typename T>
__global__ void rotate_kernel(const Cx<T>* __restrict__ in, long long n_items, int n_modes, long long is,
                              long long ms, const double* __restrict__ fac, Cx<T>* __restrict__ out) {
    const long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= n_items) return;
    for (int j = 0; j < n_modes; ++j) {
        const Cx<T> v = in[it * is + j * ms];
        const T c = (T)fac[2 * j], s = (T)fac[2 * j + 1];       // exp(-i m theta) = c + i s
        out[it * is + j * ms] = Cx<T>{v.re * c - v.im * s, v.re * s + v.im * c};
    }
}

// ---- n-fold scores on real moments  (_zmoments.py:420-462) -----------------------------------
template <typename T>
__global__ void rot_scores_kernel(const T* __restrict__ in, long long n_items, int n_modes, long long is,
                                  long long ms, const double* __restrict__ w, const unsigned char* __restrict__ sel,
                                  int n_folds, int kind, T* __restrict__ out, long long ois, long long ofs) {
    const long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= n_items) return;
    T s1 = 0, s2 = 0, sm = 0;
    T num[kMaxFolds];
    for (int f = 0; f < n_folds; ++f) num[f] = 0;
    for (int j = 0; j < n_modes; ++j) {
        if (!sel[j]) continue;
        const T z = in[it * is + j * ms], z2 = z * z, a = fabs(z);
        s1 += a;
        s2 += z2;
        sm = a > sm ? a : sm;
        for (int f = 0; f < n_folds; ++f) num[f] += (T)w[f * n_modes + j] * z2;
    }
    T den = 1;
    if (kind == ZB200_NORM_L1) den = s1 * s1;
    else if (kind == ZB200_NORM_L2) den = s2;
    else if (kind == ZB200_NORM_INF) den = sm * sm;
    for (int f = 0; f < n_folds; ++f) out[it * ois + f * ofs] = num[f] / den;
}

// ---- mirror score: max_theta sum_c Re(Zc^2 e^{-i m_c theta})  (_zmoments.py:464-493) ---------
// tab[t][c] = (cos(m_c theta_t), sin(m_c theta_t)) in the precision of the data (float64 moments keep float64
// trigonometry); eight angles share each pass over the modes.
template <typename T>
__global__ void mirror_kernel(const Cx<T>* __restrict__ in, long long n_items, int n_c, long long is,
                              long long ms, const Cx<T>* __restrict__ tab, int n_theta, T* __restrict__ out) {
    const long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= n_items) return;
    T best = -INFINITY;
    for (int t0 = 0; t0 < n_theta; t0 += 8) {
        T acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = 0;
        for (int c = 0; c < n_c; ++c) {
            const Cx<T> v = in[it * is + c * ms];
            const T p = v.re * v.re - v.im * v.im, q = (T)2 * v.re * v.im;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (t0 + u < n_theta) {
                    const Cx<T> cs = tab[(size_t)(t0 + u) * n_c + c];
                    acc[u] += p * cs.re + q * cs.im;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (t0 + u < n_theta) best = acc[u] > best ? acc[u] : best;
    }
    out[it] = best;
}

// Same score with the modes grouped by |m| first: sum_c Re(Zc^2 e^{-i m_c theta}) = sum_m (P_m cos m theta +
// Q_m sin m theta) with P_m = sum_{c: m_c = m} Re(Zc^2), Q_m = sum Im(Zc^2).  A basis up to n_max has at most
// n_max + 1 distinct m against (n_max/2 + 1)^2 complex modes, so the angle loop shrinks 3-4x (n_max = 12:
// 11 slots instead of 40 modes).  slot[c] = group of mode c; tab[t][s] = (cos, sin)(m_s theta_t); kSlots >= groups.
template <typename T, int kSlots>
__global__ void mirror_grouped_kernel(const Cx<T>* __restrict__ in, long long n_items, int n_c, long long is,
                                      long long ms, const int* __restrict__ slot, const Cx<T>* __restrict__ tab,
                                      int n_theta, T* __restrict__ out) {
    const long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= n_items) return;
    T P[kSlots], Q[kSlots];
#pragma unroll
    for (int s = 0; s < kSlots; ++s) P[s] = Q[s] = 0;
    for (int c = 0; c < n_c; ++c) {
        const Cx<T> v = in[it * is + c * ms];
        const T p = v.re * v.re - v.im * v.im, q = (T)2 * v.re * v.im;
        const int sc = __ldg(slot + c);                       // warp-uniform
#pragma unroll
        for (int s = 0; s < kSlots; ++s)
            if (s == sc) { P[s] += p; Q[s] += q; }
    }
    T best = -INFINITY;
    for (int t = 0; t < n_theta; ++t) {
        const Cx<T>* row = tab + (size_t)t * kSlots;
        T a0 = 0, a1 = 0;
#pragma unroll
        for (int s = 0; s < kSlots; s += 2) {
            const Cx<T> c0 = row[s], c1 = row[s + 1];
            a0 += P[s] * c0.re + Q[s] * c0.im;
            a1 += P[s + 1] * c1.re + Q[s + 1] * c1.im;
        }
        const T acc = a0 + a1;
        best = acc > best ? acc : best;
    }
    out[it] = best;
}

template <typename T>
__global__ void abs_phase_kernel(const Cx<T>* __restrict__ in, long long n, T* __restrict__ mag_out, T* __restrict__ ph_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Cx<T> v = in[i];
    mag_out[i] = hypot(v.re, v.im);
    if (ph_out) ph_out[i] = atan2(v.im, v.re);
}

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ in, D* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (D)in[i];
}

// Score tables for the fused kernels: weights [n_folds][cols_pad] (zero padded) followed by the
// select mask [cols_pad], in one stream-ordered allocation (no host/device synchronisation: the
// copy from pageable memory is staged by the runtime before it returns).  Free with cudaFreeAsync
// on the same stream after the consuming kernel has been enqueued.
int upload_weights(const float* h_weights, const uint8_t* h_select, int n_folds, int n_cols, int cols_pad,
                   cudaStream_t s, float** d_w, uint8_t** d_sel) {
    ZB_CHECK_ARG(n_folds >= 1 && n_folds <= kMaxFolds, "n_folds=%d out of range [1,%d]", n_folds, kMaxFolds);
    ZB_CHECK_ARG(h_weights && h_select, "weights/select must not be null");
    const size_t wbytes = sizeof(float) * (size_t)n_folds * cols_pad;
    std::vector<unsigned char> host(wbytes + cols_pad, 0);
    float* hw = reinterpret_cast<float*>(host.data());
    for (int f = 0; f < n_folds; ++f)
        for (int c = 0; c < n_cols; ++c) hw[(size_t)f * cols_pad + c] = h_weights[(size_t)f * n_cols + c];
    for (int c = 0; c < n_cols; ++c) host[wbytes + c] = h_select[c] ? 1 : 0;
    void* dev = nullptr;
    ZB_CUDA(scratch_alloc(&dev, host.size(), s));
    ZB_CUDA(cudaMemcpyAsync(dev, host.data(), host.size(), cudaMemcpyHostToDevice, s));
    *d_w = static_cast<float*>(dev);
    *d_sel = static_cast<uint8_t*>(dev) + wbytes;
    return ZB200_OK;
}

}  // namespace zb200

using namespace zb200;

#define ZB_DISPATCH_DTYPE(dtype, ...)                                         \
    if ((dtype) == ZB200_F32) { using T = float; __VA_ARGS__ }                \
    else if ((dtype) == ZB200_F64) { using T = double; __VA_ARGS__ }          \
    else { set_error("unknown dtype %d", (int)(dtype)); return ZB200_EINVAL; }

extern "C" int zb200_to_complex(int dtype, const void* d_in, int64_t n_items, int64_t iis, int64_t ims,
                                const int32_t* h_pos, const int32_t* h_neg, int n_c, void* d_out, int64_t ois,
                                int64_t oms, void* stream) {
    ZB_CHECK_ARG(d_in && d_out && h_pos && h_neg && n_c > 0 && n_items >= 0, "to_complex: bad arguments");
    if (n_items == 0) return ZB200_OK;
    cudaStream_t s = as_stream(stream);
    Scratch sc(s);
    std::vector<int32_t> tab(2 * (size_t)n_c);
    for (int c = 0; c < n_c; ++c) { tab[c] = h_pos[c]; tab[n_c + c] = h_neg[c]; }
    int rc = sc.upload(tab.data(), tab.size() * sizeof(int32_t));
    if (rc) return rc;
    const int* pos = static_cast<const int*>(sc.ptr);
    ZB_DISPATCH_DTYPE(dtype, {
        if (oms == 1 && ims == 1)
            to_complex_kernel<T, true><<<grid_for(n_items * n_c, 256), 256, 0, s>>>(
                static_cast<const T*>(d_in), n_items, iis, ims, pos, pos + n_c, n_c, static_cast<Cx<T>*>(d_out), ois, oms);
        else
            to_complex_kernel<T, false><<<grid_for(n_items, 256), 256, 0, s>>>(
                static_cast<const T*>(d_in), n_items, iis, ims, pos, pos + n_c, n_c, static_cast<Cx<T>*>(d_out), ois, oms);
    })
    ZB_LAUNCHED();
    return ZB200_OK;
}

extern "C" int zb200_to_real(int dtype, const void* d_in, int64_t n_items, int64_t iis, int64_t ims,
                             const int32_t* h_src, const uint8_t* h_take_imag, int n_r, void* d_out, int64_t ois,
                             int64_t oms, void* stream) {
    ZB_CHECK_ARG(d_in && d_out && h_src && h_take_imag && n_r > 0 && n_items >= 0, "to_real: bad arguments");
    if (n_items == 0) return ZB200_OK;
    cudaStream_t s = as_stream(stream);
    Scratch sc(s);
    std::vector<unsigned char> tab((size_t)n_r * 5);
    memcpy(tab.data(), h_src, (size_t)n_r * 4);
    memcpy(tab.data() + (size_t)n_r * 4, h_take_imag, n_r);
    int rc = sc.upload(tab.data(), tab.size());
    if (rc) return rc;
    const int* src = static_cast<const int*>(sc.ptr);
    const unsigned char* tk = static_cast<const unsigned char*>(sc.ptr) + (size_t)n_r * 4;
    ZB_DISPATCH_DTYPE(dtype, {
        if (oms == 1 && ims == 1)
            to_real_kernel<T, true><<<grid_for(n_items * n_r, 256), 256, 0, s>>>(
                static_cast<const Cx<T>*>(d_in), n_items, iis, ims, src, tk, n_r, static_cast<T*>(d_out), ois, oms);
        else
            to_real_kernel<T, false><<<grid_for(n_items, 256), 256, 0, s>>>(
                static_cast<const Cx<T>*>(d_in), n_items, iis, ims, src, tk, n_r, static_cast<T*>(d_out), ois, oms);
    })
    ZB_LAUNCHED();
    return ZB200_OK;
}

extern "C" int zb200_select_modes(int dtype, int is_complex, const void* d_in, int64_t n_items, int64_t iis,
                                  int64_t ims, const int32_t* h_index, int n_out, void* d_out, int64_t ois,
                                  int64_t oms, void* stream) {
    ZB_CHECK_ARG(d_in && d_out && h_index && n_out > 0 && n_items >= 0, "select_modes: bad arguments");
    if (n_items == 0) return ZB200_OK;
    cudaStream_t s = as_stream(stream);
    Scratch sc(s);
    int rc = sc.upload(h_index, sizeof(int32_t) * n_out);
    if (rc) return rc;
    const int* idx = static_cast<const int*>(sc.ptr);
    ZB_DISPATCH_DTYPE(dtype, {
        const bool row = oms == 1 && ims == 1;
        const unsigned grid = grid_for(row ? n_items * n_out : n_items, 256);
        if (is_complex && row)
            select_kernel<Cx<T>, true><<<grid, 256, 0, s>>>(static_cast<const Cx<T>*>(d_in), n_items, iis, ims, idx, n_out,
                                                            static_cast<Cx<T>*>(d_out), ois, oms);
        else if (is_complex)
            select_kernel<Cx<T>, false><<<grid, 256, 0, s>>>(static_cast<const Cx<T>*>(d_in), n_items, iis, ims, idx, n_out,
                                                             static_cast<Cx<T>*>(d_out), ois, oms);
        else if (row)
            select_kernel<T, true><<<grid, 256, 0, s>>>(static_cast<const T*>(d_in), n_items, iis, ims, idx, n_out,
                                                        static_cast<T*>(d_out), ois, oms);
        else
            select_kernel<T, false><<<grid, 256, 0, s>>>(static_cast<const T*>(d_in), n_items, iis, ims, idx, n_out,
                                                         static_cast<T*>(d_out), ois, oms);
    })
    ZB_LAUNCHED();
    return ZB200_OK;
}

extern "C" int zb200_normalize(int dtype, int is_complex, const void* d_in, int64_t n_items, int n_modes,
                               int64_t is, int64_t ms, int norm_kind, double order_p, void* d_out, void* stream) {
    ZB_CHECK_ARG(d_in && d_out && n_modes > 0 && n_items >= 0, "normalize: bad arguments");
    ZB_CHECK_ARG(norm_kind == ZB200_NORM_L1 || norm_kind == ZB200_NORM_L2 || norm_kind == ZB200_NORM_INF ||
                     (norm_kind < 0 && order_p > 0),
                 "normalize: unsupported norm kind %d (p=%g)", norm_kind, order_p);
    if (n_items == 0) return ZB200_OK;
    cudaStream_t s = as_stream(stream);
    ZB_DISPATCH_DTYPE(dtype, {
        if (is_complex)
            normalize_kernel<T, Cx<T>><<<grid_for(n_items, 256), 256, 0, s>>>(static_cast<const Cx<T>*>(d_in), n_items,
                                                                            n_modes, is, ms, norm_kind, order_p,
                                                                            static_cast<Cx<T>*>(d_out));
        else
            normalize_kernel<T, T><<<grid_for(n_items, 256), 256, 0, s>>>(static_cast<const T*>(d_in), n_items, n_modes,
                                                                        is, ms, norm_kind, order_p, static_cast<T*>(d_out));
    })
    ZB_LAUNCHED();
    return ZB200_OK;
}

extern "C" int zb200_rotate(int dtype, const void* d_in, int64_t n_items, int n_modes, int64_t is, int64_t ms,
                            const int32_t* h_m, double theta_rad, void* d_out, void* stream) {
    ZB_CHECK_ARG(d_in && d_out && h_m && n_modes > 0 && n_items >= 0, "rotate: bad arguments");
    if (n_items == 0) return ZB200_OK;
    cudaStream_t s = as_stream(stream);
    std::vector<double> fac(2 * (size_t)n_modes);
    for (int j = 0; j < n_modes; ++j) {   // exp(-1j * theta * m), _zmoments.py:401
        fac[2 * j] = cos(-theta_rad * h_m[j]);
        fac[2 * j + 1] = sin(-theta_rad * h_m[j]);
    }
    Scratch sc(s);
    int rc = sc.upload(fac.data(), fac.size() * sizeof(double));
    if (rc) return rc;
    ZB_DISPATCH_DTYPE(dtype, {
        rotate_kernel<T><<<grid_for(n_items, 256), 256, 0, s>>>(static_cast<const Cx<T>*>(d_in), n_items, n_modes, is, ms,
                                                              static_cast<const double*>(sc.ptr),
                                                              static_cast<Cx<T>*>(d_out));
    })
    ZB_LAUNCHED();
    return ZB200_OK;
}

extern "C" int zb200_rot_scores(int dtype, const void* d_in, int64_t n_items, int n_modes, int64_t is, int64_t ms,
                                const double* h_weights, const uint8_t* h_select, int n_folds, int norm_kind,
                                void* d_out, int64_t ois, int64_t ofs, void* stream) {
    ZB_CHECK_ARG(d_in && d_out && h_weights && h_select && n_modes > 0 && n_items >= 0, "rot_scores: bad arguments");
    ZB_CHECK_ARG(n_folds >= 1 && n_folds <= kMaxFolds, "rot_scores: n_folds=%d out of range [1,%d]", n_folds, kMaxFolds);
    ZB_CHECK_ARG(norm_kind >= ZB200_NORM_NONE && norm_kind <= ZB200_NORM_INF, "rot_scores: bad norm kind %d", norm_kind);
    if (n_items == 0) return ZB200_OK;
    cudaStream_t s = as_stream(stream);
    const size_t wbytes = sizeof(double) * (size_t)n_folds * n_modes;
    std::vector<unsigned char> tab(wbytes + n_modes);
    memcpy(tab.data(), h_weights, wbytes);
    memcpy(tab.data() + wbytes, h_select, n_modes);
    Scratch sc(s);
    int rc = sc.upload(tab.data(), tab.size());
    if (rc) return rc;
    const double* w = static_cast<const double*>(sc.ptr);
    const unsigned char* sel = static_cast<const unsigned char*>(sc.ptr) + wbytes;
    ZB_DISPATCH_DTYPE(dtype, {
        rot_scores_kernel<T><<<grid_for(n_items, 128), 128, 0, s>>>(static_cast<const T*>(d_in), n_items, n_modes, is, ms,
                                                                  w, sel, n_folds, norm_kind, static_cast<T*>(d_out), ois,
                                                                  ofs);
    })
    ZB_LAUNCHED();
    return ZB200_OK;
}

// host side of zb200_mirror_scores for one dtype: trig tables in that dtype, grouped kernel for <= 32 distinct m
template <typename T>
static int mirror_impl(const void* d_in, int64_t n_items, int n_c, int64_t is, int64_t ms, const int32_t* h_m,
                       const double* h_theta, int n_theta, void* d_out, cudaStream_t s, const std::vector<int>& distinct,
                       const std::vector<int>& slots, int n_groups) {
    Scratch sc(s);
    const unsigned grid = grid_for(n_items, 128);
    const Cx<T>* src = static_cast<const Cx<T>*>(d_in);
    T* dst = static_cast<T*>(d_out);
    if (n_groups <= 32) {
        const int k_slots = n_groups <= 8 ? 8 : (n_groups <= 16 ? 16 : 32);
        const size_t tab_off = (sizeof(int) * (size_t)n_c + 15) & ~(size_t)15;
        std::vector<unsigned char> blob(tab_off + sizeof(Cx<T>) * (size_t)n_theta * k_slots, 0);
        memcpy(blob.data(), slots.data(), sizeof(int) * (size_t)n_c);
        Cx<T>* tab = reinterpret_cast<Cx<T>*>(blob.data() + tab_off);
        for (int t = 0; t < n_theta; ++t)
            for (int g = 0; g < n_groups; ++g) {
                tab[(size_t)t * k_slots + g].re = (T)cos(distinct[g] * h_theta[t]);
                tab[(size_t)t * k_slots + g].im = (T)sin(distinct[g] * h_theta[t]);
            }
        int rc = sc.upload(blob.data(), blob.size());
        if (rc) return rc;
        const int* d_slot = static_cast<const int*>(sc.ptr);
        const Cx<T>* d_tab = reinterpret_cast<const Cx<T>*>(static_cast<const unsigned char*>(sc.ptr) + tab_off);
        if (k_slots == 8) mirror_grouped_kernel<T, 8><<<grid, 128, 0, s>>>(src, n_items, n_c, is, ms, d_slot, d_tab, n_theta, dst);
        else if (k_slots == 16) mirror_grouped_kernel<T, 16><<<grid, 128, 0, s>>>(src, n_items, n_c, is, ms, d_slot, d_tab, n_theta, dst);
        else mirror_grouped_kernel<T, 32><<<grid, 128, 0, s>>>(src, n_items, n_c, is, ms, d_slot, d_tab, n_theta, dst);
        ZB_LAUNCHED();
        return ZB200_OK;
    }
    std::vector<Cx<T>> tab((size_t)n_theta * n_c);
    for (int t = 0; t < n_theta; ++t)
        for (int c = 0; c < n_c; ++c) {
            tab[(size_t)t * n_c + c].re = (T)cos(h_m[c] * h_theta[t]);
            tab[(size_t)t * n_c + c].im = (T)sin(h_m[c] * h_theta[t]);
        }
    int rc = sc.upload(tab.data(), tab.size() * sizeof(Cx<T>));
    if (rc) return rc;
    mirror_kernel<T><<<grid, 128, 0, s>>>(src, n_items, n_c, is, ms, static_cast<const Cx<T>*>(sc.ptr), n_theta, dst);
    ZB_LAUNCHED();
    return ZB200_OK;
}

extern "C" int zb200_mirror_scores(int dtype, const void* d_in, int64_t n_items, int n_c, int64_t is, int64_t ms,
                                   const int32_t* h_m, const double* h_theta, int n_theta, void* d_out, void* stream) {
    ZB_CHECK_ARG(d_in && d_out && h_m && h_theta && n_c > 0 && n_theta > 0 && n_items >= 0, "mirror_scores: bad arguments");
    if (n_items == 0) return ZB200_OK;
    cudaStream_t s = as_stream(stream);
    // group the modes by m: slots[c] = index of m_c among the distinct m values
    std::vector<int> distinct, slots(n_c);
    for (int c = 0; c < n_c; ++c) {
        int sidx = -1;
        for (size_t i = 0; i < distinct.size(); ++i)
            if (distinct[i] == h_m[c]) sidx = (int)i;
        if (sidx < 0) { sidx = (int)distinct.size(); distinct.push_back(h_m[c]); }
        slots[c] = sidx;
    }
    const int n_groups = (int)distinct.size();
    int rc = ZB200_OK;
    ZB_DISPATCH_DTYPE(dtype, { rc = mirror_impl<T>(d_in, n_items, n_c, is, ms, h_m, h_theta, n_theta, d_out, s, distinct, slots, n_groups); })
    return rc;
}

extern "C" int zb200_complex_abs_phase(int dtype, const void* d_in, int64_t n, void* d_abs, void* d_phase, void* stream) {
    ZB_CHECK_ARG(n >= 0, "complex_abs_phase: negative count");
    if (n == 0) return ZB200_OK;
    ZB_CHECK_ARG(d_in && d_abs, "complex_abs_phase: null pointer");
    cudaStream_t s = as_stream(stream);
    ZB_DISPATCH_DTYPE(dtype, {
        abs_phase_kernel<T><<<grid_for(n, 256), 256, 0, s>>>(static_cast<const Cx<T>*>(d_in), n, static_cast<T*>(d_abs),
                                                           static_cast<T*>(d_phase));
    })
    ZB_LAUNCHED();
    return ZB200_OK;
}

extern "C" int zb200_cast(int src_dtype, const void* d_in, int dst_dtype, void* d_out, int64_t n, void* stream) {
    ZB_CHECK_ARG(d_in && d_out && n >= 0, "cast: bad arguments");
    if (n == 0) return ZB200_OK;
    cudaStream_t s = as_stream(stream);
    const unsigned g = grid_for(n, 256);
    if (src_dtype == ZB200_F32 && dst_dtype == ZB200_F64)
        cast_kernel<float, double><<<g, 256, 0, s>>>(static_cast<const float*>(d_in), static_cast<double*>(d_out), n);
    else if (src_dtype == ZB200_F64 && dst_dtype == ZB200_F32)
        cast_kernel<double, float><<<g, 256, 0, s>>>(static_cast<const double*>(d_in), static_cast<float*>(d_out), n);
    else {
        set_error("cast: unsupported dtype pair %d -> %d", src_dtype, dst_dtype);
        return ZB200_EINVAL;
    }
    ZB_LAUNCHED();
    return ZB200_OK;
}
