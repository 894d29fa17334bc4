// Host-buffer entry points: the calls a numpy user of the reference makes (host arrays in, float64 host
// arrays out).  Everything here is plumbing around the kernels: a persistent host thread pool (staging copies
// and float32 -> float64 widening are memory-bound loops that one core cannot feed at PCIe speed), two-slot
// pipelines  host -> pinned -> HBM -> kernels -> pinned -> host  whose result side runs on a second host
// thread, and the per-plan staging buffers, guarded by the plan's mutex (calls on one plan serialise).
#include "zb200_common.cuh"

#include <stdlib.h>
#include <string.h>
#if defined(__linux__)
#include <sys/mman.h>
#include <unistd.h>
#endif
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace zb200 {

// ---- persistent host pool -----------------------------------------------------------------------------
namespace {
class HostPool {
public:
    static HostPool& get() {
        static HostPool* pool = new HostPool();      // leaked on purpose: no destructor order problems at exit
        return *pool;
    }
    int width() const { return (int)workers_.size() + 1; }
    // fn(begin, end) over [0, n) in contiguous slices; the caller works on the first slice itself
    void parallel_for(int64_t n, int64_t grain, const std::function<void(int64_t, int64_t)>& fn) {
        int parts = (int)((n + grain - 1) / grain);
        if (parts > width()) parts = width();
        if (parts <= 1) { fn(0, n); return; }
        struct Latch { std::mutex m; std::condition_variable cv; int left; } latch;
        latch.left = parts - 1;
        const int64_t per = (n + parts - 1) / parts;
        {
            std::lock_guard<std::mutex> lk(m_);
            for (int t = 1; t < parts; ++t) {
                const int64_t a = t * per, b = a + per < n ? a + per : n;
                q_.emplace_back([a, b, &fn, &latch]() {
                    if (a < b) fn(a, b);
                    std::lock_guard<std::mutex> lk2(latch.m);
                    if (--latch.left == 0) latch.cv.notify_one();
                });
            }
        }
        cv_.notify_all();
        fn(0, per < n ? per : n);
        std::unique_lock<std::mutex> lk(latch.m);
        latch.cv.wait(lk, [&] { return latch.left == 0; });
    }

private:
    HostPool() {
        unsigned hw = std::thread::hardware_concurrency();
        int n = hw >= 8 ? (int)hw / 2 : (hw >= 3 ? (int)hw - 2 : 0);
        if (n > 16) n = 16;
        // the only environment variable a release build honours besides ZB200_EXPERIMENT: the pool width (1..64),
        // read once when the pool is created
        if (const char* e = getenv("ZB200_HOST_THREADS")) { const int v = atoi(e); if (v >= 1 && v <= 64) n = v - 1; }
        for (int i = 0; i < n; ++i) workers_.emplace_back([this] { run(); });
        for (auto& w : workers_) w.detach();
    }
    void run() {
        for (;;) {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return !q_.empty(); });
                job = std::move(q_.front());
                q_.pop_front();
            }
            job();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> q_;
};

// float32 -> float64 of a range.  Plain cached stores on purpose: non-temporal stores (_mm_stream_pd) halve the
// widening itself (3.4 -> 1.1 ms per 8 frames) but the frame route as a whole got SLOWER with them -- the |Zc| route
// from 4.0 to 7.0 ms per 8 frames, the GPU-side completions arriving late while the write-combining traffic and the
// D2H DMA writes share the memory controller (measured with ZB200_HOST_TRACE, profiles/r02_e2e_host_route_probe.log).
static void widen_range(const float* src, double* dst, int64_t a, int64_t b) {
    for (int64_t i = a; i < b; ++i) dst[i] = (double)src[i];
}
void widen_parallel(const float* src, double* dst, int64_t n) {
    HostPool::get().parallel_for(n, 1 << 16, [=](int64_t a, int64_t b) { widen_range(src, dst, a, b); });
}
// A fresh result array (np.empty) is first touched by the widening threads: 4 KB page faults were the largest part of
// a call that returns 100+ MB (frame route 22 M patches/s into a fresh array against 40 M/s into a reused one).  Where
// transparent huge pages are in "madvise" mode, asking for them on the 2 MB-aligned interior of the array makes the
// first touch 512x rarer; elsewhere the call is a no-op.
void advise_huge(void* ptr, size_t bytes) {
#if defined(__linux__) && defined(MADV_HUGEPAGE)
    if (bytes < (8u << 20)) return;
    const uintptr_t two_mb = (uintptr_t)2 << 20;
    const uintptr_t lo = (reinterpret_cast<uintptr_t>(ptr) + two_mb - 1) & ~(two_mb - 1);
    const uintptr_t hi = (reinterpret_cast<uintptr_t>(ptr) + bytes) & ~(two_mb - 1);
    if (hi > lo) madvise(reinterpret_cast<void*>(lo), hi - lo, MADV_HUGEPAGE);
#else
    (void)ptr; (void)bytes;
#endif
}
// First touch of a result array while the calling thread would only wait for a pipeline slot.  A fresh np.empty of
// 134 MB costs 18 ms of page faults on one thread; spread over the pool and hidden behind the first items they no longer
// sit in the widening of every item.  The touch is an atomic OR of 0: it never changes a value, so it is safe next to
// the drainer thread that may already be writing results into the same array, and for a reused array; arrays whose
// sampled pages are already resident are skipped.
void first_touch_parallel(void* ptr, size_t bytes) {
#if defined(__linux__)
    if (bytes < (8u << 20)) return;
    const uintptr_t page = 4096;
    const uintptr_t lo = (reinterpret_cast<uintptr_t>(ptr) + page - 1) & ~(page - 1);
    const uintptr_t hi = (reinterpret_cast<uintptr_t>(ptr) + bytes) & ~(page - 1);
    if (hi <= lo) return;
    const int64_t n_pages = (int64_t)((hi - lo) / page);
    int resident = 0;
    for (int i = 0; i < 8; ++i) {                            // eight pages spread over the array
        unsigned char vec = 0;
        void* q = reinterpret_cast<void*>(lo + (uintptr_t)((n_pages - 1) * i / 7) * page);
        if (mincore(q, page, &vec) == 0 && (vec & 1)) ++resident;
    }
    if (resident == 8) return;
    HostPool::get().parallel_for(n_pages, 256, [=](int64_t a, int64_t e) {
        for (int64_t pg = a; pg < e; ++pg)
            __atomic_fetch_or(reinterpret_cast<int*>(lo + (uintptr_t)pg * page), 0, __ATOMIC_RELAXED);
    });
#else
    (void)ptr; (void)bytes;
#endif
}
// float32 [F][rows][W] band -> rows [row0, row0 + rows) of a float64 [F][H][W] array, ONE parallel loop over the band
void widen_band_parallel(const float* src, double* dst, int n_folds, int rows, int W, int H, int row0) {
    const int64_t per = (int64_t)rows * W;
    HostPool::get().parallel_for((int64_t)n_folds * per, 1 << 16, [=](int64_t a, int64_t b) {
        while (a < b) {
            const int64_t f = a / per, o = a - f * per;
            const int64_t stop = (f + 1) * per < b ? (f + 1) * per : b;
            widen_range(src + f * per, dst + ((size_t)f * H + row0) * W, o, o + (stop - a));
            a = stop;
        }
    });
}
void copy_parallel(void* dst, const void* src, size_t bytes) {
    HostPool::get().parallel_for((int64_t)bytes, 1 << 20, [=](int64_t a, int64_t b) {
        memcpy(static_cast<char*>(dst) + a, static_cast<const char*>(src) + a, (size_t)(b - a));
    });
}
bool is_pinned(const void* h) {
    cudaPointerAttributes attr;
    bool pinned = false;
    if (cudaPointerGetAttributes(&attr, h) == cudaSuccess) pinned = attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    return pinned;
}

// ---- kSlots-deep pipeline: enqueue(c, slot) on the calling thread, drain(c, slot) on a second thread once the
// event of the slot has fired.  Slot c % kSlots is reused by item c only after item c - kSlots has been drained.
// Four slots: the chain upload -> gather -> projection -> download -> widening of one 2048^2 frame is ~1.1 ms long
// while its longest stage (the upload) is 0.32 ms; two slots ran at 0.56 ms per frame, three at 0.50 ms.
constexpr int kSlots = 4;
struct PipeSync {
    std::mutex m;
    std::condition_variable cv;
    int64_t enqueued = 0, drained = 0;
    int rc = 0;
    std::string msg;
};

static bool host_trace() {
    static const bool on = getenv("ZB200_HOST_TRACE") != nullptr;      // read once; diagnostics on stderr
    return on;
}
static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <class Enqueue, class Drain>
int run_pipeline(int device, int64_t n_items, cudaEvent_t* ev, Enqueue enq, Drain drain) {
    if (n_items <= 0) return ZB200_OK;
    PipeSync ps;
    const bool trace = host_trace();
    double t_enq = 0, t_slot = 0, t_evsync = 0, t_drain = 0;
    const double t_begin = now_ms();
    std::vector<double> tl(trace ? (size_t)n_items * 4 : 0, 0.0);     // per item: enqueue begin / end, event ready, drained
    std::thread drainer([&] {
        cudaSetDevice(device);
        for (int64_t c = 0; c < n_items; ++c) {
            {
                std::unique_lock<std::mutex> lk(ps.m);
                ps.cv.wait(lk, [&] { return ps.enqueued > c || ps.rc; });
                if (ps.rc) return;
            }
            int r = ZB200_OK;
            const double t0 = trace ? now_ms() : 0;
            cudaError_t e = cudaEventSynchronize(ev[c % kSlots]);
            const double t1 = trace ? now_ms() : 0;
            if (e != cudaSuccess) {
                set_error("host pipeline: %s", cudaGetErrorString(e));
                r = ZB200_ECUDA;
            } else {
                r = drain(c, (int)(c % kSlots));
            }
            if (trace) { t_evsync += t1 - t0; t_drain += now_ms() - t1; tl[(size_t)c * 4 + 2] = t1 - t_begin; tl[(size_t)c * 4 + 3] = now_ms() - t_begin; }
            std::lock_guard<std::mutex> lk(ps.m);
            if (r) { ps.rc = r; ps.msg = zb200_last_error(); }
            ps.drained = c + 1;
            ps.cv.notify_all();
            if (r) return;
        }
    });
    for (int64_t c = 0; c < n_items; ++c) {
        const double t0 = trace ? now_ms() : 0;
        {
            std::unique_lock<std::mutex> lk(ps.m);
            ps.cv.wait(lk, [&] { return ps.drained >= c - (kSlots - 1) || ps.rc; });
            if (ps.rc) break;
        }
        const double t1 = trace ? now_ms() : 0;
        const int r = enq(c, (int)(c % kSlots));
        if (trace) { t_slot += t1 - t0; t_enq += now_ms() - t1; tl[(size_t)c * 4] = t1 - t_begin; tl[(size_t)c * 4 + 1] = now_ms() - t_begin; }
        std::lock_guard<std::mutex> lk(ps.m);
        if (r) { ps.rc = r; ps.msg = zb200_last_error(); }
        else ps.enqueued = c + 1;
        ps.cv.notify_all();
        if (r) break;
    }
    drainer.join();
    if (trace)
        fprintf(stderr, "[zb200 host pipeline] items %lld total %.2f ms | main: enqueue %.2f, waiting for a slot %.2f | drainer: "
                "event wait %.2f, drain %.2f\n", (long long)n_items, now_ms() - t_begin, t_enq, t_slot, t_evsync, t_drain);
    if (trace && getenv("ZB200_HOST_TRACE")[0] == '2')
        for (int64_t c = 0; c < n_items; ++c)
            fprintf(stderr, "    item %2lld: enqueue %.2f..%.2f  event ready %.2f  drained %.2f\n", (long long)c, tl[(size_t)c * 4],
                    tl[(size_t)c * 4 + 1], tl[(size_t)c * 4 + 2], tl[(size_t)c * 4 + 3]);
    if (ps.rc) {
        set_error("%s", ps.msg.c_str());
        cudaDeviceSynchronize();                     // nothing of this call is left in flight on the staging buffers
        cudaGetLastError();
    }
    return ps.rc;
}

// growable buffers of the per-plan staging set
enum BufKind { kPinned, kDevice };
int ensure_buf(void** ptr, size_t* cap, size_t need, BufKind kind) {
    if (*cap >= need && *ptr) return ZB200_OK;
    if (*ptr) {
        if (kind == kPinned) cudaFreeHost(*ptr); else cudaFree(*ptr);
        *ptr = nullptr;
        *cap = 0;
    }
    const size_t want = need + need / 4 + 4096;      // headroom: frames of a series differ a little in peak count
    cudaError_t e = kind == kPinned ? cudaMallocHost(ptr, want) : cudaMalloc(ptr, want);
    if (e != cudaSuccess) {
        *ptr = nullptr;
        cudaGetLastError();
        set_error("host pipeline: cannot allocate %zu bytes of %s memory: %s", want, kind == kPinned ? "pinned" : "device",
                  cudaGetErrorString(e));
        return ZB200_ENOMEM;
    }
    *cap = want;
    return ZB200_OK;
}
}  // namespace

struct HostPipe {
    cudaStream_t st[kSlots] = {};
    cudaEvent_t ev[kSlots] = {};
    // ALL uploads go through one stream, in item order: with one upload per slot stream the copy engine interleaves
    // the queued 16.8 MB frame copies, all of them finish together and late (ZB200_HOST_TRACE=2: the second frame of a
    // call was ready 0.95 ms after the first instead of 0.31 ms); up_ev[slot] hands the item over to its slot stream
    cudaStream_t up = nullptr;
    cudaEvent_t up_ev[kSlots] = {};
    // per slot: input staging (pinned), result staging (pinned), device input, device patches, device points, device result
    void* pin_in[kSlots] = {};   size_t pin_in_cap[kSlots] = {};
    void* pin_out[kSlots] = {};  size_t pin_out_cap[kSlots] = {};
    void* dev_in[kSlots] = {};   size_t dev_in_cap[kSlots] = {};
    void* dev_pat[kSlots] = {};  size_t dev_pat_cap[kSlots] = {};
    void* dev_pts[kSlots] = {};  size_t dev_pts_cap[kSlots] = {};
    void* pin_pts[kSlots] = {};  size_t pin_pts_cap[kSlots] = {};   // peak lists are staged through pinned memory: an async
                                                                    // copy from the caller's pageable list blocks the enqueuing
                                                                    // thread until the frame copy in front of it has finished
    void* dev_out[kSlots] = {};  size_t dev_out_cap[kSlots] = {};
};

void free_host_pipe(zb200_plan* p) {
    HostPipe* h = p->host;
    if (!h) return;
    for (int i = 0; i < kSlots; ++i) {
        if (h->pin_in[i]) cudaFreeHost(h->pin_in[i]);
        if (h->pin_out[i]) cudaFreeHost(h->pin_out[i]);
        if (h->pin_pts[i]) cudaFreeHost(h->pin_pts[i]);
        cudaFree(h->dev_in[i]);
        cudaFree(h->dev_pat[i]);
        cudaFree(h->dev_pts[i]);
        cudaFree(h->dev_out[i]);
        if (h->ev[i]) cudaEventDestroy(h->ev[i]);
        if (h->up_ev[i]) cudaEventDestroy(h->up_ev[i]);
        if (h->st[i]) cudaStreamDestroy(h->st[i]);
    }
    if (h->up) cudaStreamDestroy(h->up);
    delete h;
    p->host = nullptr;
}

static int host_pipe(zb200_plan* p, HostPipe** out) {
    if (!p->host) {
        HostPipe* h = new (std::nothrow) HostPipe();
        if (!h) { set_error("out of host memory"); return ZB200_ENOMEM; }
        p->host = h;
        bool ok = cudaStreamCreateWithFlags(&h->up, cudaStreamNonBlocking) == cudaSuccess;
        for (int i = 0; i < kSlots; ++i) {
            if (!ok || cudaStreamCreateWithFlags(&h->st[i], cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&h->up_ev[i], cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&h->ev[i], cudaEventDisableTiming) != cudaSuccess) {
                set_error("host pipeline: cannot create stream/event: %s", cudaGetErrorString(cudaGetLastError()));
                free_host_pipe(p);                   // nothing half-built is kept
                return ZB200_ECUDA;
            }
        }
    }
    *out = p->host;
    return ZB200_OK;
}

static int row_len_of(const zb200_plan* p, int out_kind) {
    return out_kind == ZB200_OUT_REAL ? p->n_modes : (out_kind == ZB200_OUT_COMPLEX ? 2 * p->n_complex : p->n_complex);
}

}  // namespace zb200

using namespace zb200;

// ---- patch stack from host memory (ZPs.transform(numpy), _zps.py:146-157) ------------------------------------
extern "C" int zb200_project_patches_host(const zb200_plan* plan, const float* h_patches, int64_t n, int precision,
                                          double* h_out) {
    ZB_CHECK_ARG(plan, "project_host: plan is null");
    ZB_CHECK_ARG(n >= 0, "project_host: negative patch count");
    if (n == 0) return ZB200_OK;
    ZB_CHECK_ARG(h_patches && h_out, "project_host: null host pointer");
    zb200_plan* p = const_cast<zb200_plan*>(plan);
    std::lock_guard<std::mutex> lock(p->host_mu);     // one host pipeline per plan: concurrent callers take turns
    HostPipe* h = nullptr;
    int rc = host_pipe(p, &h);
    if (rc) return rc;
    int64_t chunk = (64ll << 20) / ((int64_t)p->kk * (int64_t)sizeof(float));     // 64 MiB of patches per chunk
    if (chunk < 256) chunk = 256;
    if (chunk > n) chunk = n;
    const int M = p->n_modes;
    advise_huge(h_out, sizeof(double) * (size_t)n * M);
    const bool pinned = is_pinned(h_patches);
    for (int b = 0; b < kSlots && b < ceil_div(n, chunk); ++b) {
        if (!pinned && (rc = ensure_buf(&h->pin_in[b], &h->pin_in_cap[b], sizeof(float) * chunk * p->kk, kPinned))) return rc;
        if ((rc = ensure_buf(&h->pin_out[b], &h->pin_out_cap[b], sizeof(float) * chunk * M, kPinned))) return rc;
        if ((rc = ensure_buf(&h->dev_in[b], &h->dev_in_cap[b], sizeof(float) * chunk * p->kk, kDevice))) return rc;
        if ((rc = ensure_buf(&h->dev_out[b], &h->dev_out_cap[b], sizeof(float) * chunk * M, kDevice))) return rc;
    }
    const int64_t n_chunks = ceil_div(n, chunk);
    auto count_of = [&](int64_t c) { return n - c * chunk < chunk ? n - c * chunk : chunk; };
    auto enq = [&](int64_t c, int b) -> int {
        const int64_t cnt = count_of(c);
        const size_t in_bytes = sizeof(float) * (size_t)cnt * p->kk;
        const float* src = h_patches + c * chunk * p->kk;
        if (!pinned) {
            copy_parallel(h->pin_in[b], src, in_bytes);
            src = static_cast<const float*>(h->pin_in[b]);
        }
        ZB_CUDA(cudaMemcpyAsync(h->dev_in[b], src, in_bytes, cudaMemcpyHostToDevice, h->up));
        ZB_CUDA(cudaEventRecord(h->up_ev[b], h->up));
        ZB_CUDA(cudaStreamWaitEvent(h->st[b], h->up_ev[b], 0));
        int r = project_any(p, static_cast<const float*>(h->dev_in[b]), cnt, precision, ZB200_OUT_REAL, h->dev_out[b],
                            nullptr, nullptr, nullptr, 0, 0, h->st[b]);
        if (r) return r;
        ZB_CUDA(cudaMemcpyAsync(h->pin_out[b], h->dev_out[b], sizeof(float) * (size_t)cnt * M, cudaMemcpyDeviceToHost, h->st[b]));
        ZB_CUDA(cudaEventRecord(h->ev[b], h->st[b]));
        return ZB200_OK;
    };
    auto drain = [&](int64_t c, int b) -> int {
        widen_parallel(static_cast<const float*>(h->pin_out[b]), h_out + c * chunk * M, count_of(c) * M);
        return ZB200_OK;
    };
    return run_pipeline(p->device, n_chunks, h->ev, enq, drain);
}

// ---- the reference pipeline from host memory: frames + peak coordinates in, features out ---------------------
// KeyPoints(pts, img, size).extract_patches() -> ZPs.transform -> (to_complex / np.abs), _keypoint.py:60-78,
// _zps.py:146-157, _zmoments.py:300-316; 16.8 MB per 2048^2 frame cross the bus instead of 16 KB per patch.
extern "C" int zb200_project_peaks_host(const zb200_plan* plan, const float* const* h_frames, int n_frames, int H, int W,
                                        const double* h_pts_xy, const int64_t* h_counts, int precision, int out_kind,
                                        int out_dtype, void* h_out) {
    ZB_CHECK_ARG(plan, "project_peaks_host: plan is null");
    ZB_CHECK_ARG(n_frames >= 0 && H > 0 && W > 0, "project_peaks_host: bad shape");
    ZB_CHECK_ARG(out_kind >= ZB200_OUT_REAL && out_kind <= ZB200_OUT_ABS, "project_peaks_host: out_kind %d not supported here",
                 out_kind);
    ZB_CHECK_ARG(out_dtype == ZB200_F32 || out_dtype == ZB200_F64, "project_peaks_host: bad out_dtype %d", out_dtype);
    if (n_frames == 0) return ZB200_OK;
    ZB_CHECK_ARG(h_frames && h_counts, "project_peaks_host: null host pointer");
    int64_t max_count = 0, total = 0;
    for (int f = 0; f < n_frames; ++f) {
        ZB_CHECK_ARG(h_counts[f] >= 0 && h_frames[f], "project_peaks_host: bad frame %d", f);
        if (h_counts[f] > max_count) max_count = h_counts[f];
        total += h_counts[f];
    }
    if (total == 0) return ZB200_OK;
    ZB_CHECK_ARG(h_pts_xy && h_out, "project_peaks_host: null host pointer");
    if (precision != ZB200_PREC_FP32 && !zb200_plan_supports(plan, precision, out_kind)) {
        set_error("project_peaks_host: precision %d with out_kind %d is not available for this plan", precision, out_kind);
        return ZB200_EUNSUP;
    }
    zb200_plan* p = const_cast<zb200_plan*>(plan);
    std::lock_guard<std::mutex> lock(p->host_mu);
    HostPipe* h = nullptr;
    int rc = host_pipe(p, &h);
    if (rc) return rc;
    const int L = row_len_of(p, out_kind);
    const size_t frame_bytes = sizeof(float) * (size_t)H * W;
    advise_huge(h_out, (out_dtype == ZB200_F64 ? sizeof(double) : sizeof(float)) * (size_t)total * L);
    // 64-pixel windows, n_max <= 13: one kernel gathers and projects (no patch stack in HBM)
    const bool fused_gather = precision == ZB200_PREC_F16X3 && fold_gather_supported(p) && knobs().tc_fold != 0;
    bool all_pinned = true;
    for (int f = 0; f < n_frames && all_pinned; ++f) all_pinned = is_pinned(h_frames[f]);
    for (int b = 0; b < kSlots && b < n_frames; ++b) {
        if (!all_pinned && (rc = ensure_buf(&h->pin_in[b], &h->pin_in_cap[b], frame_bytes, kPinned))) return rc;
        if ((rc = ensure_buf(&h->pin_out[b], &h->pin_out_cap[b], sizeof(float) * max_count * L, kPinned))) return rc;
        if ((rc = ensure_buf(&h->dev_in[b], &h->dev_in_cap[b], frame_bytes, kDevice))) return rc;
        if ((rc = ensure_buf(&h->dev_pts[b], &h->dev_pts_cap[b], sizeof(double) * 2 * max_count, kDevice))) return rc;
        if ((rc = ensure_buf(&h->pin_pts[b], &h->pin_pts_cap[b], sizeof(double) * 2 * max_count, kPinned))) return rc;
        if (!fused_gather && (rc = ensure_buf(&h->dev_pat[b], &h->dev_pat_cap[b], sizeof(float) * max_count * p->kk, kDevice))) return rc;
        if ((rc = ensure_buf(&h->dev_out[b], &h->dev_out_cap[b], sizeof(float) * max_count * L, kDevice))) return rc;
    }
    std::vector<int64_t> first((size_t)n_frames + 1, 0);
    for (int f = 0; f < n_frames; ++f) first[f + 1] = first[f] + h_counts[f];
    auto enq = [&](int64_t f, int b) -> int {
        const int64_t cnt = h_counts[f];
        const float* src = h_frames[f];
        if (!all_pinned) {
            copy_parallel(h->pin_in[b], src, frame_bytes);
            src = static_cast<const float*>(h->pin_in[b]);
        }
        if (cnt > 0) {
            memcpy(h->pin_pts[b], h_pts_xy + 2 * first[f], sizeof(double) * 2 * (size_t)cnt);
            ZB_CUDA(cudaMemcpyAsync(h->dev_pts[b], h->pin_pts[b], sizeof(double) * 2 * (size_t)cnt,
                                    cudaMemcpyHostToDevice, h->up));
        }
        ZB_CUDA(cudaMemcpyAsync(h->dev_in[b], src, frame_bytes, cudaMemcpyHostToDevice, h->up));
        ZB_CUDA(cudaEventRecord(h->up_ev[b], h->up));
        ZB_CUDA(cudaStreamWaitEvent(h->st[b], h->up_ev[b], 0));
        if (cnt > 0) {
            int r;
            if (fused_gather) {
                r = zb200_project_peaks_f32(p, static_cast<const float*>(h->dev_in[b]), H, W, static_cast<const double*>(h->dev_pts[b]),
                                            cnt, precision, out_kind, h->dev_out[b], nullptr, h->st[b]);
            } else {
                r = zb200_gather_patches_f32(static_cast<const float*>(h->dev_in[b]), H, W,
                                             static_cast<const double*>(h->dev_pts[b]), cnt, p->size,
                                             static_cast<float*>(h->dev_pat[b]), h->st[b]);
                if (r) return r;
                r = project_any(p, static_cast<const float*>(h->dev_pat[b]), cnt, precision, out_kind, h->dev_out[b], nullptr,
                                nullptr, nullptr, 0, 0, h->st[b]);
            }
            if (r) return r;
            ZB_CUDA(cudaMemcpyAsync(h->pin_out[b], h->dev_out[b], sizeof(float) * (size_t)cnt * L, cudaMemcpyDeviceToHost,
                                    h->st[b]));
        }
        ZB_CUDA(cudaEventRecord(h->ev[b], h->st[b]));
        if (f == (n_frames < kSlots ? n_frames : kSlots) - 1)      // every slot is busy: this thread would only wait
            first_touch_parallel(h_out, (out_dtype == ZB200_F64 ? sizeof(double) : sizeof(float)) * (size_t)total * L);
        return ZB200_OK;
    };
    auto drain = [&](int64_t f, int b) -> int {
        const int64_t cnt = h_counts[f] * L;
        if (out_dtype == ZB200_F64)
            widen_parallel(static_cast<const float*>(h->pin_out[b]), static_cast<double*>(h_out) + first[f] * L, cnt);
        else
            copy_parallel(static_cast<float*>(h_out) + first[f] * L, h->pin_out[b], sizeof(float) * (size_t)cnt);
        return ZB200_OK;
    };
    return run_pipeline(p->device, n_frames, h->ev, enq, drain);
}

// ---- the dense symmetry map from host memory: one frame in, float64 n-fold score maps out -----------------------------
// ZPs.transform(img).rot_maps(n_folds) of the reference (_zps.py:159-193 + _zmoments.py:420-462) for a numpy frame.
// The frame goes up once; the map runs in row bands (bit-identical to a single call: rows pair by absolute parity and
// the input scale comes from the whole frame), and band b's scores come down and are widened to float64 while band
// b+1 is being computed -- download (F x 16.8 MB at 2048^2) and widening used to run after the whole map.
extern "C" int zb200_symmetry_map_host(const zb200_plan* plan, const float* h_img, int H, int W, int precision,
                                       const float* h_weights, const uint8_t* h_select, int n_folds, int norm_kind,
                                       double* h_out) {
    ZB_CHECK_ARG(plan && h_img && h_out && h_weights && h_select, "symmetry_map_host: null argument");
    ZB_CHECK_ARG(H >= plan->size && W >= plan->size,
                 "For FFT convolution, image size (%dx%d) must be at least as large as polynomial size (%dx%d)", H, W,
                 plan->size, plan->size);
    ZB_CHECK_ARG(n_folds >= 1 && n_folds <= kMaxFolds, "n_folds=%d out of range [1,%d]", n_folds, kMaxFolds);
    zb200_plan* p = const_cast<zb200_plan*>(plan);
    std::lock_guard<std::mutex> lock(p->host_mu);
    HostPipe* h = nullptr;
    int rc = host_pipe(p, &h);
    if (rc) return rc;
    const size_t frame_bytes = sizeof(float) * (size_t)H * W;
    advise_huge(h_out, sizeof(double) * (size_t)n_folds * H * W);
    int band = ((H + 5) / 6 + 1) & ~1;                       // ~6 bands, an even number of rows each
    if (band < 64) band = 64;
    if (band > H) band = H;
    const int64_t n_bands = (H + band - 1) / band;
    const size_t band_bytes = sizeof(float) * (size_t)n_folds * band * W;
    const bool pinned = is_pinned(h_img);
    if (!pinned && (rc = ensure_buf(&h->pin_in[0], &h->pin_in_cap[0], frame_bytes, kPinned))) return rc;
    if ((rc = ensure_buf(&h->dev_in[0], &h->dev_in_cap[0], frame_bytes, kDevice))) return rc;
    for (int b = 0; b < kSlots && b < n_bands; ++b) {
        if ((rc = ensure_buf(&h->pin_out[b], &h->pin_out_cap[b], band_bytes, kPinned))) return rc;
        if ((rc = ensure_buf(&h->dev_out[b], &h->dev_out_cap[b], band_bytes, kDevice))) return rc;
    }
    const float* src = h_img;
    if (!pinned) {
        copy_parallel(h->pin_in[0], h_img, frame_bytes);
        src = static_cast<const float*>(h->pin_in[0]);
    }
    ZB_CUDA(cudaMemcpyAsync(h->dev_in[0], src, frame_bytes, cudaMemcpyHostToDevice, h->up));
    auto rows_of = [&](int64_t c) { return (int)(H - c * band < band ? H - c * band : band); };
    // All bands are computed on ONE stream (the upload stream), in order: on four streams the bands' kernels shared
    // the SMs, all of them finished together and the first scores arrived after 3.7 ms instead of 1.2
    // (ZB200_HOST_TRACE=2); only the download of band c runs on its slot stream, behind an event.
    const size_t out_bytes = sizeof(double) * (size_t)n_folds * H * W;
    auto enq = [&](int64_t c, int b) -> int {
        const int rows = rows_of(c);
        int r = zb200_symmetry_map_f32(p, static_cast<const float*>(h->dev_in[0]), H, W, (int)(c * band), rows, precision,
                                       h_weights, h_select, n_folds, norm_kind, static_cast<float*>(h->dev_out[b]), h->up);
        if (r) return r;
        ZB_CUDA(cudaEventRecord(h->up_ev[b], h->up));
        ZB_CUDA(cudaStreamWaitEvent(h->st[b], h->up_ev[b], 0));
        ZB_CUDA(cudaMemcpyAsync(h->pin_out[b], h->dev_out[b], sizeof(float) * (size_t)n_folds * rows * W,
                                cudaMemcpyDeviceToHost, h->st[b]));
        ZB_CUDA(cudaEventRecord(h->ev[b], h->st[b]));
        if (c == (n_bands < kSlots ? n_bands : kSlots) - 1) first_touch_parallel(h_out, out_bytes);   // this thread would only wait
        return ZB200_OK;
    };
    auto drain = [&](int64_t c, int b) -> int {
        const int rows = rows_of(c);
        // band scores are [F][rows][W]; the result is [F][H][W]
        widen_band_parallel(static_cast<const float*>(h->pin_out[b]), h_out, n_folds, rows, W, H, (int)(c * band));
        return ZB200_OK;
    };
    return run_pipeline(p->device, n_bands, h->ev, enq, drain);
}

// ---- result download: float32 in HBM -> float64 host array (what the reference returns) -----------------
// Chunked D2H into two pinned staging buffers on a private stream, widened to float64 by the host pool
// while the next chunk is in flight (the destination's first-touch page faults are spread over the threads).
namespace {
struct Downloader {
    std::mutex mu;
    float* pin[2] = {nullptr, nullptr};
    cudaStream_t stream = nullptr;
    cudaEvent_t ready = nullptr, done[2] = {nullptr, nullptr};
    int device = -1;
    static constexpr int64_t kChunk = 8ll << 20;      // floats per chunk (32 MiB)
};
constexpr int kMaxDevices = 64;
Downloader g_downloaders[kMaxDevices];
}  // namespace

extern "C" int zb200_download_as_f64(const float* d_src, int64_t n, double* h_dst, void* stream) {
    ZB_CHECK_ARG(n >= 0, "download: negative count");
    if (n == 0) return ZB200_OK;
    ZB_CHECK_ARG(d_src && h_dst, "download: null pointer");
    advise_huge(h_dst, sizeof(double) * (size_t)n);
    int dev = 0;
    ZB_CUDA(cudaGetDevice(&dev));
    ZB_CHECK_ARG(dev >= 0 && dev < kMaxDevices, "download: device ordinal %d out of range", dev);
    Downloader& g_dl = g_downloaders[dev];           // one staging set per device (each lives in that device's context)
    std::lock_guard<std::mutex> lock(g_dl.mu);
    if (g_dl.device != dev) {
        for (int i = 0; i < 2; ++i) {
            ZB_CUDA(cudaMallocHost(&g_dl.pin[i], sizeof(float) * Downloader::kChunk));
            ZB_CUDA(cudaEventCreateWithFlags(&g_dl.done[i], cudaEventDisableTiming));
        }
        ZB_CUDA(cudaEventCreateWithFlags(&g_dl.ready, cudaEventDisableTiming));
        ZB_CUDA(cudaStreamCreateWithFlags(&g_dl.stream, cudaStreamNonBlocking));
        g_dl.device = dev;
    }
    // the producing kernels run on the caller's stream
    ZB_CUDA(cudaEventRecord(g_dl.ready, as_stream(stream)));
    ZB_CUDA(cudaStreamWaitEvent(g_dl.stream, g_dl.ready, 0));
    const int64_t n_chunks = ceil_div(n, Downloader::kChunk);
    auto count_of = [&](int64_t c) { const int64_t off = c * Downloader::kChunk; return n - off < Downloader::kChunk ? n - off : Downloader::kChunk; };
    for (int64_t c = 0; c <= n_chunks; ++c) {
        if (c < n_chunks) {
            ZB_CUDA(cudaMemcpyAsync(g_dl.pin[c & 1], d_src + c * Downloader::kChunk, sizeof(float) * count_of(c),
                                    cudaMemcpyDeviceToHost, g_dl.stream));
            ZB_CUDA(cudaEventRecord(g_dl.done[c & 1], g_dl.stream));
        }
        if (c >= 1) {
            ZB_CUDA(cudaEventSynchronize(g_dl.done[(c - 1) & 1]));
            widen_parallel(g_dl.pin[(c - 1) & 1], h_dst + (c - 1) * Downloader::kChunk, count_of(c - 1));
        }
    }
    return ZB200_OK;
}
