// Shared internals of libzernike_b200.so (sm_100a).  Not part of the C ABI.
#pragma once

#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include <mutex>

#include "zernike_b200.h"

namespace zb200 {

// ---- error plumbing ---------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

#define ZB_CHECK_ARG(cond, ...)                                   \
    do {                                                          \
        if (!(cond)) {                                            \
            ::zb200::set_error(__VA_ARGS__);                      \
            return ZB200_EINVAL;                                  \
        }                                                         \
    } while (0)

#define ZB_CUDA(call)                                                                  \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            ::zb200::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,    \
                               cudaGetErrorString(e__));                               \
            return (e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver)    \
                       ? ZB200_ENODEV : ZB200_ECUDA;                                   \
        }                                                                              \
    } while (0)

// call after every kernel launch: counts it and surfaces launch-config errors
#define ZB_LAUNCHED()                                                                  \
    do {                                                                               \
        ::zb200::g_launches.fetch_add(1, std::memory_order_relaxed);                   \
        ZB_CUDA(cudaGetLastError());                                                   \
    } while (0)

// Experiment knobs (A/B measurements of DESIGN.md).  They exist ONLY when the process is started with
// ZB200_EXPERIMENT=1; then each ZB200_* variable below is read ONCE (first use), range-checked, and out-of-range
// values are ignored.  Without the switch the library never looks at the environment on a launch path.
struct Knobs {
    int tc_kskip = -1;     // ZB200_TC_KSKIP   0|1   issue every K step / skip the all-zero ones
    int tc_chunk = 0;      // ZB200_TC_CHUNK   1..64 k-blocks per accumulation chunk
    int tc_cluster = 0;    // ZB200_TC_CLUSTER 1|2|4 CTAs sharing the basis stream
    int tc_pair = -1;      // ZB200_TC_PAIR    0|1   cta_group::2 pairs
    int tc_bstages = 0;    // ZB200_TC_BSTAGES 1..4
    int tc_stages = 0;     // ZB200_TC_STAGES  1..8
    int tc_accbufs = 0;    // ZB200_TC_ACCBUFS 1|2
    int tc_fold = -1;      // ZB200_TC_FOLD    0|1   mirror-folded projection when the plan has a folded operand
    int tc_park = -1;      // ZB200_TC_PARK    0|1   parked (suspend-time hint) waits of the idle roles in the folded kernel
    int tc_split2 = -1;    // ZB200_TC_SPLIT2  0|1   two splitter warpgroups + one epilogue warpgroup (metric shape)
    int tc_debug = 0;      // ZB200_TC_DEBUG   bit mask, needs a -DZB200_DEBUG_HOOKS=1 build (results are wrong when set)
    int map_gskip = -1;    // ZB200_MAP_GSKIP  0|1
    int map_bstages = 0;   // ZB200_MAP_BSTAGES 2..8
    int map_slots = 0;     // ZB200_MAP_SLOTS  2..16 (checked against shared memory by the launcher)
    int map_debug = 0;     // ZB200_MAP_DEBUG  bit mask, debug-hook builds only
    int gather_4b = 0;     // ZB200_GATHER_4B  1: 4-byte gathers in the fused gather->projection path
};
const Knobs& knobs();
#ifndef ZB200_DEBUG_HOOKS
#define ZB200_DEBUG_HOOKS 0      // ablation bits / blocked-cycle counters inside the hot kernels: compiled out of releases
#endif

// stream-ordered scratch from the library's private memory pool (zb200_api.cu); release with cudaFreeAsync
cudaError_t scratch_alloc(void** ptr, size_t bytes, cudaStream_t s);
template <class T> static inline cudaError_t scratch_alloc(T** ptr, size_t bytes, cudaStream_t s) {
    return scratch_alloc(reinterpret_cast<void**>(ptr), bytes, s);
}

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// ---- the plan: everything cached in HBM for one (n_max, size) -----------------
// HBM layout (all allocations 256-B aligned by cudaMalloc):
//   basis64   double [M][k][k]                the reference's ZPs.polynomials
//   op_full   float  [rows_pad][k_pad]        K-major GEMM operand, value = V/area (fp32 RN)
//   op_hi     float  [rows_pad][k_pad]        tf32-exact high part  (RN of V/area to 11 bits)
//   op_lo     float  [rows_pad][k_pad]        tf32-exact low part   (RN of V/area - hi)
//   op_cb     bf16   [rows_pad][2*k_pad]      correction operand of the fp32-grade tensor path: per
//                                             group of 8 k: 8 x bf16(V/area) then 8 x bf16(V/area - hi)
//   op_t      float  [k_pad][rows_pad]        transpose of op_full for the SIMT kernels
//   two row orders of each: REAL (row r = mode j) and CPLX (row 2c = (n,+m), 2c+1 = (n,-m))
struct Operand {
    int rows = 0;        // meaningful rows
    int rows_pad = 0;    // padded to a multiple of 16 (UMMA N granularity at M=128)
    float* full = nullptr;
    float* hi = nullptr;
    float* lo = nullptr;
    uint32_t* cb = nullptr;  // bf16 pairs, same byte geometry as one fp32 row per operand row
    float* t = nullptr;
    // fp16-split operand of the f16x3 projection: per 32-tap k-block and row 128 bytes = 32 x half b1 (RN_f16(V))
    // then 32 x half b2 (RN_f16(V - b1)); V, not V/area (fp16 range) -- same byte geometry as one fp32 row
    uint32_t* hb = nullptr;
    // TMA descriptors over [rows_pad][k_pad] 4-byte words, SW128, box 32 x rows_pad/C for cluster
    // sizes C = 1, 2, 4 (index log2 C): each CTA of a cluster fetches 1/C of the rows and multicasts
    CUtensorMap tmap_hi[3];
    CUtensorMap tmap_cb[3];
    CUtensorMap tmap_lo_f32[3];   // over the fp32 `lo` operand (dense-map 3xTF32 path)
    CUtensorMap tmap_hb[3];       // over `hb`
    int max_cluster = 1;     // largest C with rows_pad % (8 C) == 0
    bool has_tmap = false;
};

// fp16-split operand of the dense map (zb200_map_h.cu): b1 = RN_f16(V), b2 = RN_f16(V - b1), only the
// 16-tap groups of each window row that touch the unit disk, packed in issue order.
constexpr int kMapHalfMaxWindow = 128;
struct MapHalf {
    bool ready = false;
    int n_groups = 0;            // 16-tap groups per window row
    int a_first = 0, a_end = 0;  // window rows with taps inside the disk
    int n_act = 0;               // group positions of the packed operand (incl. zero padding)
    int n_kb = 0;                // 128-byte k-blocks (4 groups) per operand row
    int n_blocks = 0;            // window rows per tile
    int mode0 = 0, n_modes = 0;  // the modes [mode0, mode0 + n_modes) this operand holds (a plan with more than 128
    int rows_pad = 0;            //   padded modes is served by several operands = several passes over the frame)
    int max_cluster = 1;
    unsigned short act[kMapHalfMaxWindow] = {};
    void* b1 = nullptr;          // half [rows_pad][n_kb*64]
    void* b2 = nullptr;
    CUtensorMap tmap_b1[2];
    CUtensorMap tmap_b2[2];
};

// mirror-folded fp16-split operand of the patch projection (zb200_project_fold.cu): accumulator columns grouped in the
// four mirror-parity classes, K = the window's upper-left quadrant
struct FoldOperand {
    bool ready = false;
    int cfg = -1;                // compiled class-width configuration
    int cols = 0;                // accumulator columns = operand rows (all four classes, padded)
    int nj = 0;                  // 32-tap boxes per half window row
    int n_sb = 0;                // super-blocks (32 folded taps) in the quadrant
    int sb_count = 0;            // those with taps inside the unit disk
    unsigned short* d_sb_list = nullptr;   // their indices, ascending
    uint32_t* fb = nullptr;      // [cols][n_sb * 32] words: per super-block 32 x half b1 | 32 x half b2; rows [rank][class][slot]
    unsigned char* d_umask = nullptr;   // per super-block: bit h = 16-tap unit h is active
    short* d_col_real = nullptr;        // [cols] real mode of an accumulator column (-1 = padding)
    short* d_slot_cplx = nullptr;       // [2][80] complex mode of a slot of the class pairs A (m even) / B (m odd)
    CUtensorMap tmap;
};

struct HostPipe;
}  // namespace zb200

struct zb200_plan {
    int n_max = 0;
    int size = 0;
    int n_modes = 0;      // M
    int n_complex = 0;    // Mc
    int kk = 0;           // k*k
    int k_pad = 0;        // kk rounded up to 32
    int kb_first = 0, kb_last = 0;   // 32-tap k-blocks [kb_first, kb_last) contain every tap inside the unit disk
    unsigned char* d_kmask = nullptr;   // [k_pad/32] bit j = taps [8j, 8j+8) of the k-block touch the unit disk
    int device = 0;
    int sm_count = 0;
    int cc_major = 0;
    double inv_area = 0.0;
    double* basis64 = nullptr;
    int32_t* d_n = nullptr;       // device copies of the mode table
    int32_t* d_m = nullptr;
    int32_t h_n[1024];
    int32_t h_m[1024];
    zb200::Operand real;          // real row order
    zb200::Operand cplx;          // complex-interleaved row order
    zb200::MapHalf map_half[4];   // fp16-split dense-map operands (real row order), <= 128 padded modes each
    int n_map_parts = 0;
    zb200::FoldOperand fold;      // mirror-folded projection operand (window side % 64 == 0, n_max <= 20)
    // staging set of the host-buffer entry points (zb200_host.cu); calls on one plan take turns under host_mu
    std::mutex host_mu;
    zb200::HostPipe* host = nullptr;
};

namespace zb200 {
constexpr int kMaxFolds = 16;
constexpr int kMaxModes = 1024;   // n_max <= 43

// kernels / launchers implemented across the .cu files
int launch_basis(zb200_plan* plan, cudaStream_t s);
int launch_pack(zb200_plan* plan, cudaStream_t s);
int init_tensor_maps(zb200_plan* plan);

void free_host_pipe(zb200_plan* plan);
// precision / epilogue dispatch of the patch projection (zb200_api.cu)
int project_any(const zb200_plan* plan, const float* d_patches, int64_t n, int precision, int out_kind, void* d_out,
                void* d_out2, const float* d_w, const uint8_t* d_sel, int n_folds, int norm_kind, cudaStream_t s,
                double value_max = 0.0);

int project_simt(const zb200_plan* plan, const float* d_patches, int64_t n, float* d_out_real, cudaStream_t s);
// frame + window corners for the fused gather -> projection path (K2 inside K3)
struct GatherSource {
    const float* img;
    int H, W;
    const int2* xy0;      // top-left corner (x0, y0) of each window
    // optional 16-byte gather source: 4 shifted zero-padded copies, plane r [y][u] = img0[y][u - L + r], pitch Wp
    const float* planes = nullptr;
    int Wp = 0, L = 0;
};
// K5 fused into K3: the same output rows inside up to 7 peer GPUs' result arrays (device pointers valid on this
// device: CUDA IPC mappings with peer access, zb200_peer.cu)
struct PeerTargets {
    int n = 0;
    float* out[7] = {};
};
// value_max: an upper bound of |patch values| for the f16x3 mode (sets the power-of-two input scale); ignored otherwise
int project_tc(const zb200_plan* plan, const float* d_patches, int64_t n, int precision, int out_kind,
               void* d_out, void* d_out2, const float* d_w, const uint8_t* d_sel, int n_folds, int norm_kind,
               cudaStream_t s, const GatherSource* gather = nullptr, const PeerTargets* peers = nullptr,
               double value_max = 0.0, const unsigned* run_if = nullptr);
bool tc_supported(const zb200_plan* plan, int precision, bool complex_order);
// mirror-folded fp16-split projection (value_max required; REAL / COMPLEX / ABS / ABS_PHASE outputs, peer push)
int init_fold_operand(zb200_plan* plan);
void free_fold_operand(zb200_plan* plan);
bool fold_supported(const zb200_plan* plan);
int project_fold(const zb200_plan* plan, const float* d_patches, int64_t n, int out_kind, void* d_out, void* d_out2,
                 cudaStream_t s, const PeerTargets* peers, double value_max, uint32_t* d_aux = nullptr,
                 const GatherSource* gather = nullptr);
bool fold_gather_supported(const zb200_plan* plan);

int map_simt(const zb200_plan* plan, const float* d_img, int H, int W, int row0, int rows,
             float* d_moments, float* d_scores, const float* d_w, const uint8_t* d_sel, int n_folds,
             int norm_kind, cudaStream_t s);

bool map_tc_supported(const zb200_plan* plan, int precision);
int map_tc(const zb200_plan* plan, const float* d_img, int H, int W, int row0, int rows, int precision,
           float* d_moments, float* d_scores, const float* d_w, const uint8_t* d_sel, int n_folds, int norm_kind,
           cudaStream_t s);

int init_map_half_operand(zb200_plan* plan);
void free_map_half_operand(zb200_plan* plan);
bool map_h_supported(const zb200_plan* plan, int precision);
// h_w [n_folds][n_modes] / h_sel [n_modes] are HOST tables (they travel in the kernel parameters)
int map_h(const zb200_plan* plan, const float* d_img, int H, int W, int row0, int rows, int precision,
          float* d_moments, float* d_scores, const float* h_w, const uint8_t* h_sel, int n_folds, int norm_kind,
          cudaStream_t s);

int upload_weights(const float* h_weights, const uint8_t* h_select, int n_folds, int n_cols, int cols_pad,
                   cudaStream_t s, float** d_w, uint8_t** d_sel);
}  // namespace zb200
