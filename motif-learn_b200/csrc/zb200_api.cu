// C-ABI entry points: plan lifetime, projection / map dispatch, host-buffer pipeline.
// See include/zernike_b200.h for the contract of every function.
#include "zb200_common.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

namespace zb200 {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static int knob(const char* name, int lo, int hi, int fallback) {
    const char* e = getenv(name);
    if (!e || !*e) return fallback;
    char* end = nullptr;
    const long v = strtol(e, &end, 10);
    if (end == e || *end != '\0' || v < lo || v > hi) {
        fprintf(stderr, "[zb200] ignoring %s=%s (allowed range %d..%d)\n", name, e, lo, hi);
        return fallback;
    }
    return (int)v;
}

const Knobs& knobs() {
    static const Knobs k = [] {
        Knobs v;
        const char* on = getenv("ZB200_EXPERIMENT");
        if (!on || strcmp(on, "1") != 0) return v;
        v.tc_kskip = knob("ZB200_TC_KSKIP", 0, 1, -1);
        v.tc_chunk = knob("ZB200_TC_CHUNK", 1, 64, 0);
        v.tc_cluster = knob("ZB200_TC_CLUSTER", 1, 4, 0);
        if (v.tc_cluster == 3) v.tc_cluster = 0;
        v.tc_pair = knob("ZB200_TC_PAIR", 0, 1, -1);
        v.tc_bstages = knob("ZB200_TC_BSTAGES", 1, 4, 0);
        v.tc_stages = knob("ZB200_TC_STAGES", 1, 8, 0);
        v.tc_accbufs = knob("ZB200_TC_ACCBUFS", 1, 2, 0);
        v.tc_split2 = knob("ZB200_TC_SPLIT2", 0, 1, -1);
        v.tc_fold = knob("ZB200_TC_FOLD", 0, 1, -1);
        v.tc_park = knob("ZB200_TC_PARK", 0, 1, -1);
        v.map_gskip = knob("ZB200_MAP_GSKIP", 0, 1, -1);
        v.map_bstages = knob("ZB200_MAP_BSTAGES", 2, 8, 0);
        v.map_slots = knob("ZB200_MAP_SLOTS", 2, 16, 0);
        v.gather_4b = knob("ZB200_GATHER_4B", 0, 1, 0);
#if ZB200_DEBUG_HOOKS
        v.tc_debug = knob("ZB200_TC_DEBUG", 0, 255, 0);
        v.map_debug = knob("ZB200_MAP_DEBUG", 0, 255, 0);
#endif
        return v;
    }();
    return k;
}

// Stream-ordered scratch (score tables, operand planes of the dense map, peak-detection keys, shifted frame planes:
// 0.1-0.6 GB per call at 2048^2..4096^2, 6 GB for materialised 4096^2 moment maps) comes from a PRIVATE memory pool per
// device that keeps its blocks cached across synchronisation points: a frame series allocates the same blocks again
// every frame, on several streams, and a pool that trims at every synchronisation (a bounded release threshold, once
// the cache has grown past it) turns each of them into a driver allocation -- measured: config 5 fell from 2700 to
// 1700 frames/s after a 4096^2 map had run in the same process.  The application's default pool is never touched
// (ADVICE r1); zb200_trim_scratch() hands the cache back.
static std::mutex g_pool_mu;
static cudaMemPool_t g_pools[64] = {};
cudaError_t scratch_alloc(void** ptr, size_t bytes, cudaStream_t s) {
    std::mutex& mu = g_pool_mu;
    cudaMemPool_t* pools = g_pools;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaMallocAsync(ptr, bytes, s);
    cudaMemPool_t pool;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!pools[dev]) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            cudaMemPool_t created = nullptr;
            if (cudaMemPoolCreate(&created, &props) == cudaSuccess) {
                uint64_t keep = UINT64_MAX;
                cudaMemPoolSetAttribute(created, cudaMemPoolAttrReleaseThreshold, &keep);
                pools[dev] = created;
            } else {
                cudaGetLastError();
            }
        }
        pool = pools[dev];
    }
    return pool ? cudaMallocFromPoolAsync(ptr, bytes, pool, s) : cudaMallocAsync(ptr, bytes, s);
}

static int check_device(int* sms, int* major, int* minor) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device available (%s); this library has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return ZB200_ENODEV;
    }
    int dev = 0;
    ZB_CUDA(cudaGetDevice(&dev));
    ZB_CUDA(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
    ZB_CUDA(cudaDeviceGetAttribute(major, cudaDevAttrComputeCapabilityMajor, dev));
    ZB_CUDA(cudaDeviceGetAttribute(minor, cudaDevAttrComputeCapabilityMinor, dev));
    return ZB200_OK;
}

static int alloc_operand(Operand& op, int rows, int k_pad) {
    op.rows = rows;
    op.rows_pad = round_up(rows, 16);
    const size_t bytes = sizeof(float) * (size_t)op.rows_pad * k_pad;
    ZB_CUDA(cudaMalloc(&op.full, bytes));
    ZB_CUDA(cudaMalloc(&op.hi, bytes));
    ZB_CUDA(cudaMalloc(&op.lo, bytes));
    ZB_CUDA(cudaMalloc(&op.cb, bytes));
    ZB_CUDA(cudaMalloc(&op.t, bytes));
    ZB_CUDA(cudaMalloc(&op.hb, bytes));
    return ZB200_OK;
}
static void free_operand(Operand& op) {
    cudaFree(op.full); cudaFree(op.hi); cudaFree(op.lo); cudaFree(op.cb); cudaFree(op.t); cudaFree(op.hb);
    op = Operand();
}

}  // namespace zb200

using namespace zb200;

extern "C" int zb200_abi_version(void) { return ZB200_ABI_VERSION; }

extern "C" int zb200_trim_scratch(void) {
    int dev = 0;
    ZB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_pool_mu);
    if (dev >= 0 && dev < 64 && g_pools[dev]) {
        ZB_CUDA(cudaDeviceSynchronize());
        ZB_CUDA(cudaMemPoolTrimTo(g_pools[dev], 0));
    }
    return ZB200_OK;
}
extern "C" const char* zb200_last_error(void) { return g_err; }
extern "C" int64_t zb200_launch_count(void) { return g_launches.load(); }
extern "C" void zb200_reset_launch_count(void) { g_launches.store(0); }

extern "C" int zb200_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int s = 0, a = 0, b = 0;
    int rc = check_device(&s, &a, &b);
    if (rc) return rc;
    if (sm_count) *sm_count = s;
    if (cc_major) *cc_major = a;
    if (cc_minor) *cc_minor = b;
    return ZB200_OK;
}

extern "C" int zb200_num_modes(int n_max) {
    if (n_max < 0) return ZB200_EINVAL;
    return (n_max + 1) * (n_max + 2) / 2;
}
extern "C" int zb200_num_complex_modes(int n_max) {
    if (n_max < 0) return ZB200_EINVAL;
    int c = 0;
    for (int n = 0; n <= n_max; ++n) c += n / 2 + 1;
    return c;
}
extern "C" int zb200_mode_table(int n_max, int32_t* h_n, int32_t* h_m) {
    ZB_CHECK_ARG(n_max >= 0 && h_n && h_m, "mode_table: bad arguments");
    int j = 0;
    for (int n = 0; n <= n_max; ++n)
        for (int m = -n; m <= n; m += 2, ++j) { h_n[j] = n; h_m[j] = m; }
    return j;
}

extern "C" void zb200_plan_destroy(zb200_plan* p) {
    if (!p) return;
    cudaFree(p->basis64);
    cudaFree(p->d_kmask);
    cudaFree(p->d_n);
    cudaFree(p->d_m);
    free_operand(p->real);
    free_operand(p->cplx);
    free_map_half_operand(p);
    free_fold_operand(p);
    free_host_pipe(p);
    delete p;
}

extern "C" int zb200_plan_create(int n_max, int size, zb200_plan** out_plan) {
    ZB_CHECK_ARG(out_plan, "plan_create: out_plan is null");
    *out_plan = nullptr;
    ZB_CHECK_ARG(n_max >= 0, "n_max must be non-negative.");
    ZB_CHECK_ARG(size > 0, "size must be positive.");
    ZB_CHECK_ARG(zb200_num_modes(n_max) <= kMaxModes, "n_max=%d gives more than %d modes", n_max, kMaxModes);
    ZB_CHECK_ARG(size <= 4096, "size=%d too large", size);
    int sms = 0, major = 0, minor = 0;
    int rc = check_device(&sms, &major, &minor);
    if (rc) return rc;
    zb200_plan* p = new (std::nothrow) zb200_plan();
    if (!p) { set_error("out of host memory"); return ZB200_ENOMEM; }
    p->n_max = n_max;
    p->size = size;
    p->n_modes = zb200_num_modes(n_max);
    p->n_complex = zb200_num_complex_modes(n_max);
    p->kk = size * size;
    p->k_pad = round_up(p->kk, 32);
    p->sm_count = sms;
    p->cc_major = major;
    cudaGetDevice(&p->device);
    p->inv_area = 1.0 / (M_PI * (double)size * (double)size / 4.0);    // area = pi k^2/4, _zps.py:154
    zb200_mode_table(n_max, p->h_n, p->h_m);
    {
        // k-blocks of 32 taps that contain a tap of the unit disk (grid x_i = -1 + 2 i/(k-1), rho <= 1, _zps.py:68-75;
        // the 1e-9 margin only ever keeps a block): the tensor-core projection streams and multiplies only those
        const int k = size;
        int first = -1, last = 0;
        for (int e = 0; e < p->kk; ++e) {
            const double y = k > 1 ? -1.0 + 2.0 * (e / k) / (k - 1) : 0.0, x = k > 1 ? -1.0 + 2.0 * (e % k) / (k - 1) : 0.0;
            if (x * x + y * y <= 1.0 + 1e-9) {
                if (first < 0) first = e / 32;
                last = e / 32 + 1;
            }
        }
        p->kb_first = first < 0 ? 0 : first;
        p->kb_last = first < 0 ? 1 : last;
    }

#define ZB_PLAN_TRY(expr)                      \
    do {                                       \
        int rc__ = (expr);                     \
        if (rc__) { zb200_plan_destroy(p); return rc__; } \
    } while (0)
#define ZB_PLAN_CUDA(call)                                                             \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            set_error("%s failed: %s", #call, cudaGetErrorString(e__));                \
            zb200_plan_destroy(p);                                                     \
            return e__ == cudaErrorMemoryAllocation ? ZB200_ENOMEM : ZB200_ECUDA;      \
        }                                                                              \
    } while (0)

    {
        // 8-tap K steps that touch the unit disk, per 32-tap k-block (same rule as above)
        const int k = size, nkb = p->k_pad / 32;
        std::vector<unsigned char> km((size_t)nkb, 0);
        for (int e = 0; e < p->kk; ++e) {
            const double y = k > 1 ? -1.0 + 2.0 * (e / k) / (k - 1) : 0.0, x = k > 1 ? -1.0 + 2.0 * (e % k) / (k - 1) : 0.0;
            if (x * x + y * y <= 1.0 + 1e-9) km[e / 32] |= (unsigned char)(1u << ((e % 32) / 8));
        }
        // second half of the allocation: all steps on (ZB200_TC_KSKIP=0, A/B measurements)
        ZB_PLAN_CUDA(cudaMalloc(&p->d_kmask, 2 * (size_t)nkb));
        ZB_PLAN_CUDA(cudaMemcpy(p->d_kmask, km.data(), (size_t)nkb, cudaMemcpyHostToDevice));
        ZB_PLAN_CUDA(cudaMemset(p->d_kmask + nkb, 0x0F, (size_t)nkb));
    }
    ZB_PLAN_CUDA(cudaMalloc(&p->basis64, sizeof(double) * (size_t)p->n_modes * p->kk));
    ZB_PLAN_CUDA(cudaMalloc(&p->d_n, sizeof(int32_t) * p->n_modes));
    ZB_PLAN_CUDA(cudaMalloc(&p->d_m, sizeof(int32_t) * p->n_modes));
    ZB_PLAN_CUDA(cudaMemcpy(p->d_n, p->h_n, sizeof(int32_t) * p->n_modes, cudaMemcpyHostToDevice));
    ZB_PLAN_CUDA(cudaMemcpy(p->d_m, p->h_m, sizeof(int32_t) * p->n_modes, cudaMemcpyHostToDevice));
    ZB_PLAN_TRY(alloc_operand(p->real, p->n_modes, p->k_pad));
    ZB_PLAN_TRY(alloc_operand(p->cplx, 2 * p->n_complex, p->k_pad));
    ZB_PLAN_TRY(launch_basis(p, nullptr));
    ZB_PLAN_TRY(launch_pack(p, nullptr));
    ZB_PLAN_CUDA(cudaDeviceSynchronize());
    ZB_PLAN_TRY(init_tensor_maps(p));
    ZB_PLAN_TRY(init_map_half_operand(p));
    ZB_PLAN_TRY(init_fold_operand(p));
    *out_plan = p;
    return ZB200_OK;
}

extern "C" int zb200_plan_n_max(const zb200_plan* p) { return p ? p->n_max : ZB200_EINVAL; }
extern "C" int zb200_plan_size(const zb200_plan* p) { return p ? p->size : ZB200_EINVAL; }
extern "C" const double* zb200_plan_basis_device(const zb200_plan* p) { return p ? p->basis64 : nullptr; }
extern "C" int zb200_plan_basis_to_host(const zb200_plan* p, double* h_out) {
    ZB_CHECK_ARG(p && h_out, "plan_basis_to_host: null argument");
    ZB_CUDA(cudaMemcpy(h_out, p->basis64, sizeof(double) * (size_t)p->n_modes * p->kk, cudaMemcpyDeviceToHost));
    return ZB200_OK;
}
extern "C" int zb200_plan_supports(const zb200_plan* p, int precision, int out_kind) {
    if (!p || out_kind < ZB200_OUT_REAL || out_kind > ZB200_OUT_ABS_PHASE) return 0;
    if (precision == ZB200_PREC_FP32) return out_kind == ZB200_OUT_REAL ? 1 : 0;
    if (precision == ZB200_PREC_TF32 || precision == ZB200_PREC_TF32X3 || precision == ZB200_PREC_F16X3)
        return tc_supported(p, precision, out_kind != ZB200_OUT_REAL) ? 1 : 0;
    return 0;
}

extern "C" int zb200_plan_supports_autorange(const zb200_plan* p) { return p && fold_supported(p) && knobs().tc_fold != 0 ? 1 : 0; }
extern "C" int zb200_plan_supports_folded_gather(const zb200_plan* p) { return p && fold_gather_supported(p) && knobs().tc_fold != 0 ? 1 : 0; }

extern "C" int zb200_plan_supports_map(const zb200_plan* p, int precision) {
    if (!p) return 0;
    if (precision == ZB200_PREC_FP32) return 1;
    if (precision == ZB200_PREC_F16 || precision == ZB200_PREC_F16X3) return map_h_supported(p, precision) ? 1 : 0;
    return map_tc_supported(p, precision) ? 1 : 0;
}

// ---- K3 ---------------------------------------------------------------------------------------
namespace zb200 {
int project_any(const zb200_plan* p, const float* d_patches, int64_t n, int precision, int out_kind,
                void* d_out, void* d_out2, const float* d_w, const uint8_t* d_sel, int n_folds,
                int norm_kind, cudaStream_t s, double value_max) {
    if (precision == ZB200_PREC_FP32) {
        if (out_kind != ZB200_OUT_REAL || d_w) {
            set_error("the fp32 SIMT projection only produces real moments (out_kind REAL)");
            return ZB200_EUNSUP;
        }
        return project_simt(p, d_patches, n, static_cast<float*>(d_out), s);
    }
    if (precision == ZB200_PREC_TF32 || precision == ZB200_PREC_TF32X3 || precision == ZB200_PREC_F16X3) {
        return project_tc(p, d_patches, n, precision, out_kind, d_out, d_out2, d_w, d_sel, n_folds, norm_kind, s, nullptr,
                          nullptr, value_max);
    }
    set_error("unknown precision %d", precision);
    return ZB200_EINVAL;
}
}  // namespace zb200

extern "C" int zb200_project_patches_f32(const zb200_plan* p, const float* d_patches, int64_t n, int precision,
                                         int out_kind, void* d_out, void* d_out2, void* stream) {
    ZB_CHECK_ARG(p, "project: plan is null");
    ZB_CHECK_ARG(n >= 0, "project: negative patch count");
    ZB_CHECK_ARG(out_kind >= ZB200_OUT_REAL && out_kind <= ZB200_OUT_ABS_PHASE, "project: bad out_kind %d", out_kind);
    if (n == 0) return ZB200_OK;
    ZB_CHECK_ARG(d_patches && d_out, "project: null device pointer");
    ZB_CHECK_ARG(out_kind != ZB200_OUT_ABS_PHASE || d_out2, "project: ABS_PHASE needs d_out2");
    return project_any(p, d_patches, n, precision, out_kind, d_out, d_out2, nullptr, nullptr, 0, 0, as_stream(stream));
}

extern "C" int zb200_project_patches_ranged_f32(const zb200_plan* p, const float* d_patches, int64_t n, double value_max,
                                                int out_kind, void* d_out, void* d_out2, void* stream) {
    ZB_CHECK_ARG(p, "project_ranged: plan is null");
    ZB_CHECK_ARG(n >= 0, "project_ranged: negative patch count");
    ZB_CHECK_ARG(out_kind >= ZB200_OUT_REAL && out_kind <= ZB200_OUT_ABS_PHASE, "project_ranged: bad out_kind %d", out_kind);
    ZB_CHECK_ARG(value_max >= 0.0 && value_max < 1e30, "project_ranged: value_max must be a positive finite bound of |x| (0 = auto-range)");
    if (value_max == 0.0 && !fold_supported(p)) {
        set_error("project_ranged: auto-range (value_max = 0) needs the mirror-folded kernel (window side a multiple of 64, n_max <= 20)");
        return ZB200_EUNSUP;
    }
    if (n == 0) return ZB200_OK;
    ZB_CHECK_ARG(d_patches && d_out, "project_ranged: null device pointer");
    ZB_CHECK_ARG(out_kind != ZB200_OUT_ABS_PHASE || d_out2, "project_ranged: ABS_PHASE needs d_out2");
    return project_any(p, d_patches, n, ZB200_PREC_F16X3, out_kind, d_out, d_out2, nullptr, nullptr, 0, 0, as_stream(stream),
                       value_max);
}

extern "C" int zb200_project_patches_scores_f32(const zb200_plan* p, const float* d_patches, int64_t n,
                                                int precision, const float* h_weights, const uint8_t* h_select,
                                                int n_folds, int norm_kind, float* d_scores, void* stream) {
    ZB_CHECK_ARG(p, "project_scores: plan is null");
    ZB_CHECK_ARG(n >= 0, "project_scores: negative patch count");
    ZB_CHECK_ARG(norm_kind >= ZB200_NORM_NONE && norm_kind <= ZB200_NORM_INF, "project_scores: bad norm kind");
    if (n == 0) return ZB200_OK;
    ZB_CHECK_ARG(d_patches && d_scores, "project_scores: null device pointer");
    ZB_CHECK_ARG(n_folds >= 1 && n_folds <= kMaxFolds, "n_folds=%d out of range [1,%d]", n_folds, kMaxFolds);
    ZB_CHECK_ARG(h_weights && h_select, "weights/select must not be null");
    cudaStream_t s = as_stream(stream);
    if (precision == ZB200_PREC_FP32) {
        // SIMT contraction into a temporary, then the score kernel (two launches, no fusion)
        float* tmp = nullptr;
        ZB_CUDA(scratch_alloc(&tmp, sizeof(float) * (size_t)n * p->n_modes, s));
        int rc = project_simt(p, d_patches, n, tmp, s);
        if (!rc) {
            double wd[kMaxFolds * kMaxModes];          // bounded by the checks above: nothing throws across the C ABI
            for (size_t i = 0; i < (size_t)n_folds * p->n_modes; ++i) wd[i] = h_weights[i];
            rc = zb200_rot_scores(ZB200_F32, tmp, n, p->n_modes, p->n_modes, 1, wd, h_select, n_folds,
                                  norm_kind, d_scores, n_folds, 1, stream);
        }
        cudaFreeAsync(tmp, s);
        return rc;
    }
    float* d_w = nullptr;
    uint8_t* d_sel = nullptr;
    int rc = upload_weights(h_weights, h_select, n_folds, p->n_modes, p->real.rows_pad, s, &d_w, &d_sel);
    if (rc) return rc;
    rc = project_any(p, d_patches, n, precision, ZB200_OUT_REAL, d_scores, nullptr, d_w, d_sel, n_folds, norm_kind, s);
    cudaFreeAsync(d_w, s);
    return rc;
}

// ---- K4 ---------------------------------------------------------------------------------------
static int check_map_args(const zb200_plan* p, const float* d_img, int H, int W, int row0, int rows) {
    ZB_CHECK_ARG(p && d_img, "map: null argument");
    ZB_CHECK_ARG(H >= p->size && W >= p->size,
                 "For FFT convolution, image size (%dx%d) must be at least as large as polynomial size (%dx%d)", H, W,
                 p->size, p->size);
    ZB_CHECK_ARG(row0 >= 0 && rows >= 0 && row0 + rows <= H, "map: row band [%d,%d) outside image of %d rows", row0,
                 row0 + rows, H);
    return ZB200_OK;
}

extern "C" int zb200_moment_map_f32(const zb200_plan* p, const float* d_img, int H, int W, int row0, int rows,
                                    int precision, float* d_out, void* stream) {
    int rc = check_map_args(p, d_img, H, W, row0, rows);
    if (rc) return rc;
    ZB_CHECK_ARG(d_out || rows == 0, "map: null output");
    if (precision == ZB200_PREC_F16 || precision == ZB200_PREC_F16X3)
        return map_h(p, d_img, H, W, row0, rows, precision, d_out, nullptr, nullptr, nullptr, 0, 0, as_stream(stream));
    if (precision != ZB200_PREC_FP32)
        return map_tc(p, d_img, H, W, row0, rows, precision, d_out, nullptr, nullptr, nullptr, 0, 0, as_stream(stream));
    return map_simt(p, d_img, H, W, row0, rows, d_out, nullptr, nullptr, nullptr, 0, 0, as_stream(stream));
}

extern "C" int zb200_symmetry_map_f32(const zb200_plan* p, const float* d_img, int H, int W, int row0, int rows,
                                      int precision, const float* h_weights, const uint8_t* h_select, int n_folds,
                                      int norm_kind, float* d_scores, void* stream) {
    int rc = check_map_args(p, d_img, H, W, row0, rows);
    if (rc) return rc;
    ZB_CHECK_ARG(d_scores || rows == 0, "symmetry map: null output");
    ZB_CHECK_ARG(norm_kind >= ZB200_NORM_NONE && norm_kind <= ZB200_NORM_INF, "symmetry map: bad norm kind");
    cudaStream_t s = as_stream(stream);
    if (precision == ZB200_PREC_F16 || precision == ZB200_PREC_F16X3)
        return map_h(p, d_img, H, W, row0, rows, precision, nullptr, d_scores, h_weights, h_select, n_folds, norm_kind, s);
    float* d_w = nullptr;
    uint8_t* d_sel = nullptr;
    rc = upload_weights(h_weights, h_select, n_folds, p->n_modes, p->real.rows_pad, s, &d_w, &d_sel);
    if (rc) return rc;
    if (precision != ZB200_PREC_FP32)
        rc = map_tc(p, d_img, H, W, row0, rows, precision, nullptr, d_scores, d_w, d_sel, n_folds, norm_kind, s);
    else
        rc = map_simt(p, d_img, H, W, row0, rows, nullptr, d_scores, d_w, d_sel, n_folds, norm_kind, s);
    cudaFreeAsync(d_w, s);
    return rc;
}

