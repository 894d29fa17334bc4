/*
 * zernike_b200.h -- C ABI of the B200-native Zernike hot path (libzernike_b200.so).
 *
 * Drop-in boundary for ONE path of jiadongdan/motif-learn (mtflearn 0.1.2): the
 * ZPs patch-moment transform and the zmoments dense symmetry map.  The reference
 * has no FFI for this path (it is pure Python); the entry points
 * below are what a binding for that path would bind, one per reference step, and
 * each cites the reference interface it replaces (file:line in /root/reference).
 *
 * Conventions
 *   - every function returns 0 on success, a negative ZB200_E* code on failure;
 *     zb200_last_error() gives the message of the calling thread's last failure.
 *   - d_* pointers are DEVICE pointers on the current CUDA device, h_* are HOST
 *     pointers.  `stream` is a cudaStream_t passed as void* (NULL = default
 *     stream).  No call synchronises the device unless documented.
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary.
 *   - layouts are C-order, exactly the reference's: patches (N,k,k); real moments
 *     (N,M) with modes ordered n ascending, m=-n,-n+2..n (j=((n+2)n+m)/2); complex
 *     moments (N,Mc) ordered by nm2j_complex; maps (M,H,W) / (F,H,W).
 *   - threads: device-pointer entry points may be called concurrently from several host threads (they only
 *     enqueue on the given stream).  The HOST-buffer entry points (zb200_project_patches_host,
 *     zb200_project_peaks_host, zb200_download_as_f64) own per-plan / per-device staging buffers and take a
 *     mutex for the whole call: concurrent callers on one plan are served one after the other.
 *   - sm_100a only.  There is no CPU fallback: without a CUDA device every compute
 *     entry point fails with ZB200_ENODEV.
 */
#ifndef ZERNIKE_B200_H_
#define ZERNIKE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* exported with default visibility; everything else in the library is hidden */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define ZB200_ABI_VERSION 1

/* error codes */
#define ZB200_OK        0
#define ZB200_EINVAL   -1   /* bad argument (message says which)              */
#define ZB200_ECUDA    -2   /* a CUDA runtime / driver call failed            */
#define ZB200_ENODEV   -3   /* no usable sm_100 device                        */
#define ZB200_ENOMEM   -4
#define ZB200_EUNSUP   -5   /* combination not supported by the requested path*/

/* arithmetic of the projection / map contraction */
#define ZB200_PREC_FP32    0   /* fp32 FFMA (SIMT) -- always available            */
#define ZB200_PREC_TF32    1   /* tcgen05 kind::tf32, one pass (fast, ~1e-3 rel)  */
#define ZB200_PREC_TF32X3  2   /* tcgen05 kind::tf32, 3-pass split (fp32-grade)   */
#define ZB200_PREC_F16     3   /* dense map only: tcgen05 kind::f16, one pass on a range-scaled frame (~tf32-grade) */
#define ZB200_PREC_F16X3   4   /* fp16 operand split x1+x2, b1+b2, three passes (fp32-grade): dense map; patch stacks through zb200_project_patches_ranged_f32 */

/* what the projection epilogue writes */
#define ZB200_OUT_REAL       0 /* float  [N, M]       real moments  (_zps.py:146-157)             */
#define ZB200_OUT_COMPLEX    1 /* float2 [N, Mc]      Z[n,+m]+i Z[n,-m] (_zmoments.py:300-316)     */
#define ZB200_OUT_ABS        2 /* float  [N, Mc]      |Zc|  (notebook "2 How to use ZPs" cell 13)  */
#define ZB200_OUT_ABS_PHASE  3 /* float  [N, Mc] x2   |Zc| -> d_out, angle(Zc) -> d_out2           */

/* norm used by normalize / rot scores (np.linalg.norm ord, _zmoments.py:344-356) */
#define ZB200_NORM_NONE 0
#define ZB200_NORM_L1   1
#define ZB200_NORM_L2   2
#define ZB200_NORM_INF  3

/* element types of the zmoments algebra kernels */
#define ZB200_F32 0
#define ZB200_F64 1

typedef struct zb200_plan zb200_plan;

/* ---- meta ---------------------------------------------------------------- */
int         zb200_abi_version(void);
const char* zb200_last_error(void);
/* SM count and compute capability of the current device. */
int         zb200_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Stream-ordered scratch of the library (frame planes, peak-detection keys, score tables, map operand planes) lives in
 * a library-private cudaMemPool per device that keeps freed blocks cached; this synchronises the current device and
 * returns the cached blocks to the driver.  The application's default memory pool is never reconfigured. */
int         zb200_trim_scratch(void);
/* Kernels launched by this library since the last reset (bench.py gpu_launches). */
int64_t     zb200_launch_count(void);
void        zb200_reset_launch_count(void);
/* 1 if the projection can run at this precision with this ZB200_OUT_* epilogue for the
 * plan's shape (the tcgen05 kernels need sm_100, an even window and a bounded mode count). */
int         zb200_plan_supports(const zb200_plan* plan, int precision, int out_kind);
/* Same question for the dense map (K4): the tf32 tcgen05 map needs a window that is a multiple of 8, the
 * fp16-split map (ZB200_PREC_F16 / F16X3) any window up to 128 px; both need <= 128 padded modes. */
int         zb200_plan_supports_map(const zb200_plan* plan, int precision);
/* 1 if ZB200_PREC_F16X3 projections of this plan need no value_max: windows whose side is a multiple of 64 pixels
 * (n_max <= 20) run the mirror-folded kernel, which can take its input scale from a sample of the stack and is backed
 * by the range-free TF32X3 kernel when an unsampled value overflows (see zb200_project_patches_ranged_f32). */
int         zb200_plan_supports_autorange(const zb200_plan* plan);
/* 1 if zb200_project_peaks_f32 accepts ZB200_PREC_F16X3 for this plan (64-pixel windows, n_max <= 13): the
 * mirror-folded kernel then gathers the windows from the frame itself -- KeyPoints.extract_patches + ZPs.transform
 * (_keypoint.py:60-78 + _zps.py:146-157) in one launch, the frame maximum as the exact range bound. */
int         zb200_plan_supports_folded_gather(const zb200_plan* plan);

/* ---- K1: basis (replaces ZPs.__init__/_generate_polynomials, _zps.py:23-90) - */
/* (n_max+1)(n_max+2)/2, or negative on bad n_max. */
int zb200_num_modes(int n_max);
/* Complex (m>=0) mode count. */
int zb200_num_complex_modes(int n_max);
/* Host mode table in ZPs order (_zps.py:77-80); arrays of zb200_num_modes(). */
int zb200_mode_table(int n_max, int32_t* h_n, int32_t* h_m);
/* Build the plan on the current device: fp64 basis V[M,k,k] generated by a stable
 * radial recurrence and cached in HBM with the GEMM-ready fp32 operands.  */
int  zb200_plan_create(int n_max, int size, zb200_plan** out_plan);
void zb200_plan_destroy(zb200_plan* plan);
int  zb200_plan_n_max(const zb200_plan* plan);
int  zb200_plan_size(const zb200_plan* plan);
/* Device pointer of the cached fp64 basis [M,k,k] (ZPs.polynomials). */
const double* zb200_plan_basis_device(const zb200_plan* plan);
/* Copy the fp64 basis to a host buffer of M*k*k doubles (synchronises). */
int  zb200_plan_basis_to_host(const zb200_plan* plan, double* h_out);

/* ---- K2: patch gather (replaces KeyPoints.extract_patches, _keypoint.py:60-78) */
/* d_out[i] = img[y-k/2 : y-k/2+k, x-k/2 : x-k/2+k], (x,y)=rint(pts[i]) half-to-even;
 * pixels outside the image read as 0 (the reference requires clear_border first). */
int zb200_gather_patches_f32(const float* d_img, int H, int W,
                             const double* d_pts_xy, int64_t n_pts, int k,
                             float* d_out, void* stream);

/* ---- K3: patch projection (replaces ZPs._transform_dot_product, _zps.py:146-157) */
/* Z = X[N,k*k] . V[M,k*k]^T / (pi k^2/4) with the epilogue selected by out_kind. */
int zb200_project_patches_f32(const zb200_plan* plan, const float* d_patches, int64_t n_patches,
                              int precision, int out_kind, void* d_out, void* d_out2, void* stream);
/* The same projection in the fp16-split arithmetic (ZB200_PREC_F16X3: x = x1 + x2, V = b1 + b2 in fp16, three
 * tcgen05 kind::f16 passes -- fp32-grade like TF32X3 with a quarter fewer tensor-core instructions and half the basis
 * traffic).  fp16 has a narrow exponent range: value_max must bound |patch values|; inputs are scaled by the power
 * of two that brings value_max to <= 2^14.  Values beyond 4 x value_max overflow to inf/NaN in the output.
 * Windows whose side is a multiple of 64 pixels (n_max <= 20) run the MIRROR-FOLDED kernel: the four mirror images of
 * a quadrant pixel are combined first (every Zernike plane is even or odd under the two mirrors of the grid), so each
 * of the four parity classes is a GEMM over a quarter of the taps and its own modes only.  For those plans
 * (zb200_plan_supports_autorange) value_max = 0 asks for AUTO-RANGE: the scale comes from the largest |x| of up to
 * 1024 evenly spread patches, and if an unsampled value overflows (a result is not finite) the range-free TF32X3
 * kernel enqueued behind recomputes the stack -- same stream, no host synchronisation.  zb200_project_patches_f32
 * with ZB200_PREC_F16X3 is the same auto-range call. */
int zb200_project_patches_ranged_f32(const zb200_plan* plan, const float* d_patches, int64_t n_patches, double value_max,
                                     int out_kind, void* d_out, void* d_out2, void* stream);
/* Fused n-fold scores of patches (rot_maps on real moments, _zmoments.py:420-462):
 * h_weights[F,M] is construct_rot_maps_matrix on the FULL mode list with zeros on
 * unselected modes, h_select[M] is 1 for modes kept by unselect(). d_scores [N,F]. */
int zb200_project_patches_scores_f32(const zb200_plan* plan, const float* d_patches, int64_t n_patches,
                                     int precision, const float* h_weights, const uint8_t* h_select,
                                     int n_folds, int norm_kind, float* d_scores, void* stream);
/* K2 fused into K3: moments of the windows centred at the peaks, gathered straight from the frame
 * inside the projection kernel -- the patch stack never exists in HBM (replaces
 * KeyPoints.extract_patches + ZPs.transform, _keypoint.py:60-78 + _zps.py:146-157).  TF32X3 (windows >= 32 px) or,
 * where zb200_plan_supports_folded_gather says so, F16X3 (the mirror-folded kernel with a gathering warpgroup: the
 * faster route, no patch stack in HBM); pixels outside the frame read as 0.  Same epilogues as zb200_project_patches_f32. */
int zb200_project_peaks_f32(const zb200_plan* plan, const float* d_img, int H, int W,
                            const double* d_pts_xy, int64_t n_pts, int precision, int out_kind,
                            void* d_out, void* d_out2, void* stream);
/* Same contraction from HOST buffers (the call a numpy user makes): chunked
 * H2D -> kernel -> D2H pipeline; h_out is double [N,M] like the reference. Blocks. */
int zb200_project_patches_host(const zb200_plan* plan, const float* h_patches, int64_t n_patches,
                               int precision, double* h_out);
/* The reference's patch route end to end from HOST buffers, for a series of frames of one shape:
 * KeyPoints(pts, frame, size).extract_patches() -> ZPs.transform(patches) [-> to_complex() / np.abs()]
 * (_keypoint.py:60-78, _zps.py:146-157, _zmoments.py:300-316).  h_frames[f] is frame f (float [H,W]),
 * h_counts[f] its number of peaks, h_pts_xy the peaks of all frames back to back as (x, y) doubles (already
 * filtered by clear_border).  Only the frames and the coordinates cross the bus (16.8 MB per 2048^2 frame
 * instead of 16 KB per patch); frames are pipelined over two streams (H2D of frame f+1 and the result
 * download / widening of frame f-1 overlap the kernels of frame f).  out_kind REAL | COMPLEX | ABS;
 * h_out is [sum(counts), M | 2Mc | Mc] of out_dtype (ZB200_F64: what the reference returns).  Blocks. */
int zb200_project_peaks_host(const zb200_plan* plan, const float* const* h_frames, int n_frames, int H, int W,
                             const double* h_pts_xy, const int64_t* h_counts, int precision, int out_kind,
                             int out_dtype, void* h_out);

/* ---- K5: final feature gather over NVLink peer memory (the reference is single-process) ------ */
/* One process per GPU.  Every rank allocates its copy of the gathered array with zb200_peer_buffer_alloc, hands the
 * 64-byte handle to the other ranks (any transport), which map it with zb200_peer_buffer_open (CUDA IPC, peer
 * access enabled lazily).  zb200_project_patches_push_f32 is zb200_project_patches_f32 whose kernel ALSO writes
 * every finished tile of output rows to the same rows of the peers' copies (P2P stores over NVLink from a warp of
 * the projection kernel, overlapped with the computation of the next tiles): d_out = this rank's rows inside its
 * own copy, d_out_peers[g] = the same rows inside peer g's copy (equal modulo 16 bytes).  The rows are complete on
 * every rank once all ranks' kernels have finished (order that with a barrier / collective on the stream).
 * Tensor-core precisions only; out_kind REAL | COMPLEX | ABS. */
#define ZB200_IPC_HANDLE_BYTES 64
int zb200_peer_buffer_alloc(size_t bytes, void** d_ptr, unsigned char* handle /* [64] */);
int zb200_peer_buffer_free(void* d_ptr);
int zb200_peer_buffer_open(const unsigned char* handle /* [64] */, void** d_ptr);
int zb200_peer_buffer_close(void* d_ptr);
int zb200_project_patches_push_f32(const zb200_plan* plan, const float* d_patches, int64_t n_patches, int precision,
                                   int out_kind, double value_max /* ZB200_PREC_F16X3 only */, void* d_out,
                                   void* const* d_out_peers, int n_peers, void* stream);
/* Copy-engine forwarding of a 2-D block (e.g. the F row bands of a score map) into a peer's array. */
int zb200_peer_copy_2d(void* d_dst, size_t dst_pitch, const void* d_src, size_t src_pitch, size_t width_bytes,
                       size_t height, void* stream);

/* ---- K4: dense map (replaces ZPs._transform_fft_convolve, _zps.py:159-193) --- */
/* Moments at every pixel of rows [row0,row0+rows): d_out float [M,rows,W].
 * Z[j,y,x] = 1/area sum_{a,b} img0[y-k/2+a, x-k/2+b] V[j,a,b], img0 zero-extended. */
int zb200_moment_map_f32(const zb200_plan* plan, const float* d_img, int H, int W,
                         int row0, int rows, int precision, float* d_out, void* stream);
/* Fused map -> n-fold scores, never materialising the moments: d_scores [F,rows,W]. */
int zb200_symmetry_map_f32(const zb200_plan* plan, const float* d_img, int H, int W,
                           int row0, int rows, int precision,
                           const float* h_weights, const uint8_t* h_select, int n_folds,
                           int norm_kind, float* d_scores, void* stream);

/* The same from HOST memory: one float32 frame in (pinned or pageable), float64 score maps [F,H,W] out -- what
 * ZPs.transform(img).rot_maps(n_folds) returns in the reference for a numpy frame (_zps.py:159-193 + _zmoments.py:420-462).
 * The map runs in row bands (bit-identical to one call); the scores of a band are downloaded and widened to float64
 * while the next band is computed.  Owns per-plan staging like the other *_host entry points. */
int zb200_symmetry_map_host(const zb200_plan* plan, const float* h_img, int H, int W, int precision,
                            const float* h_weights, const uint8_t* h_select, int n_folds, int norm_kind, double* h_out);

/* ---- zmoments algebra on device arrays (replaces _zmoments.py:300-462) ------- */
/* A moment array is addressed as elem(item,mode) = base[item*item_stride + mode*mode_stride]
 * (2-D (N,M): item_stride=M, mode_stride=1; 3-D (M,H,W): item_stride=1, mode_stride=H*W).
 * dtype is ZB200_F32 / ZB200_F64; complex arrays are interleaved (re,im) of that type. */
int zb200_to_complex(int dtype, const void* d_in, int64_t n_items, int64_t in_item_stride, int64_t in_mode_stride,
                     const int32_t* h_pos, const int32_t* h_neg, int n_complex,
                     void* d_out, int64_t out_item_stride, int64_t out_mode_stride, void* stream);
int zb200_to_real(int dtype, const void* d_in, int64_t n_items, int64_t in_item_stride, int64_t in_mode_stride,
                  const int32_t* h_src, const uint8_t* h_take_imag, int n_real,
                  void* d_out, int64_t out_item_stride, int64_t out_mode_stride, void* stream);
int zb200_select_modes(int dtype, int is_complex, const void* d_in, int64_t n_items,
                       int64_t in_item_stride, int64_t in_mode_stride,
                       const int32_t* h_index, int n_out,
                       void* d_out, int64_t out_item_stride, int64_t out_mode_stride, void* stream);
/* norm_kind as above; order_p used when norm_kind<0 (generic p-norm). In place allowed. */
int zb200_normalize(int dtype, int is_complex, const void* d_in, int64_t n_items, int n_modes,
                    int64_t item_stride, int64_t mode_stride, int norm_kind, double order_p,
                    void* d_out, void* stream);
/* data * exp(-i m theta) on complex input (_zmoments.py:377-418); h_m[n_modes]. */
int zb200_rotate(int dtype, const void* d_in, int64_t n_items, int n_modes,
                 int64_t item_stride, int64_t mode_stride, const int32_t* h_m, double theta_rad,
                 void* d_out, void* stream);
/* scores[item,f] = sum_j w[f,j] * (x_j/norm)^2 over the n_modes given (already
 * unselected by the caller or masked through zero weights + h_select). */
int zb200_rot_scores(int dtype, const void* d_in, int64_t n_items, int n_modes,
                     int64_t item_stride, int64_t mode_stride,
                     const double* h_weights, const uint8_t* h_select, int n_folds, int norm_kind,
                     void* d_out, int64_t out_item_stride, int64_t out_fold_stride, void* stream);
/* max over n_theta angles of sum_c Re(Zc^2 e^{-i m theta}) (_zmoments.py:464-493);
 * input: complex moments already unselected+normalised; h_m[n_complex]. */
int zb200_mirror_scores(int dtype, const void* d_in, int64_t n_items, int n_complex,
                        int64_t item_stride, int64_t mode_stride, const int32_t* h_m,
                        const double* h_theta, int n_theta, void* d_out, void* stream);
/* |z| (and angle(z) when d_phase != NULL) of n interleaved complex numbers: the rotation-invariant
 * features np.abs(m.to_complex().data) / np.angle(...) of notebook "2 How to use ZPs" cell 13. */
int zb200_complex_abs_phase(int dtype, const void* d_in, int64_t n, void* d_abs, void* d_phase, void* stream);
/* ---- "next" row f1: synthetic frame renderer (replaces add_tapered_gaussian,
 * mtflearn/datasets/_tapered_gaussian.py:3-97, used by HoneyCombLattice.to_image) ------------- */
/* img (+)= sum over atoms of A*exp(-r^2/(2 sigma^2))*(1-3t^2+2t^3), t=r/(r_factor*sigma), r<=r_factor*sigma.
 * d_pts_xy double [n,2] (x,y); d_amps double [n] or NULL (then amp_scalar for all); accumulate=0 starts
 * from a zero frame, 1 adds to the existing one (second sub-lattice). */
int zb200_render_atoms_f32(const double* d_pts_xy, const double* d_amps, double amp_scalar, int64_t n_atoms,
                           double sigma, double r_factor, int H, int W, float* d_img, int accumulate, void* stream);
/* Honeycomb lattice sites on the device (replaces HoneyCombLattice._generate_coordinates,
 * mtflearn/datasets/_honeycomb_lattice.py:103-140): sites (n1, n2) in [-n_index, n_index]^2, n1-major, sub-lattices
 * A and B -> d_coords_a / d_coords_b, double [(2 n_index + 1)^2, 2] as (x, y).  h_* are host 2-vectors; the optional
 * d_jitter_* (same shape as the outputs, host-drawn with the reference's RNG stream) are added last. */
int zb200_lattice_coords_f64(int n_index, const double* h_a1, const double* h_a2, const double* h_dA, const double* h_dB,
                             const double* h_offset, double angle_rad, double centre, const double* d_jitter_a,
                             const double* d_jitter_b, double* d_coords_a, double* d_coords_b, void* stream);
/* Delta placement + Gaussian blur of one species (replaces TMDImageSimulator.place_atoms_delta + the fftconvolve
 * with gaussian_kernel in simulate, mtflearn/datasets/_tmd_simulator.py:151-186): atoms rounded to pixels (half to
 * even), dropped outside the frame, each stamps scale * amplitude * exp(-(dx^2+dy^2)/(2 sigma^2)) on the
 * kernel_size^2 pixels around it (kernel_size odd).  accumulate=1: img = float32(img + blurred). */
int zb200_render_stamps_f32(const double* d_pts_xy, const float* d_scales, int64_t n_atoms, double amplitude, double sigma,
                            int kernel_size, int H, int W, float* d_img, int accumulate, void* stream);
/* ---- "next" row f2: peak detection (replaces local_max, mtflearn/features/_local_max_v2.py:6-66 =
 * skimage peak_local_max(min_distance=1, threshold_abs) + intensity-ordered radius suppression) ---------- */
/* d_xy_out int32 [capacity,2] receives the kept peaks as (x, y), brightest first (equal intensities in
 * raster order).  has_threshold=0 means threshold=None (= image.min()).  *h_count = number of kept peaks,
 * *h_candidates (may be NULL) = local maxima before suppression.  capacity=0 with d_xy_out=NULL only counts;
 * a too small capacity fails with ZB200_EINVAL after setting *h_count.  Blocks until *h_count is known; the coordinates
 * in d_xy_out are ordered on `stream` like any kernel output. */
int zb200_local_max_f32(const float* d_img, int H, int W, double min_distance, int has_threshold, double threshold,
                        int32_t* d_xy_out, int64_t capacity, int64_t* h_count, int64_t* h_candidates, void* stream);
/* ---- "next" row f4: PCA of the feature matrix (replaces pca, mtflearn/features/_dimension_reduction.py:3-6 =
 * sklearn PCA(n_components).fit_transform on the covariance route) ------------------------------------------ */
/* d_gram[m*m] = X^T X and d_colsum[m] = column sums of the row-major float matrix X (n, m), accumulated in float64
 * in a fixed order (deterministic).  The m x m eigenproblem is the caller's (host) job. */
int zb200_gram_f32(const float* d_x, int64_t n, int m, double* d_gram, double* d_colsum, void* stream);
/* d_out double [n, n_comp]: (x_i - mean) . components[c]  with host tables mean[m], components[n_comp, m]. */
int zb200_pca_scores_f32(const float* d_x, int64_t n, int m, const double* h_mean, const double* h_components,
                         int n_comp, double* d_out, void* stream);
/* ---- "next" row f4: cluster labels of the feature matrix (replaces the passes over X inside kmeans_lbs / gmm_lbs,
 * mtflearn/clustering/_clustering_functions.py:8-34 = scikit-learn KMeans / GaussianMixture('full')).  All arithmetic
 * in float64 on the float32 row-major matrix d_x (n, d), centred by h_mean on the fly; deterministic reductions. ---- */
/* k-means++ seeding step: d_out[j][i] = min(d_closest[i], |x_i - mean - cand_j|^2) (no min when d_closest is NULL),
 * h_pot[j] = sum_i d_out[j][i]; h_cand (t, d) in centred coordinates.  Blocks. */
int zb200_kmeans_mindist_f32(const float* d_x, int64_t n, int d, const double* h_mean, const double* h_cand, int t,
                             const double* d_closest, double* d_out, double* h_pot, void* stream);
/* One Lloyd iteration: labels = argmin_c |x - mean - centre_c|^2 (first minimum), *h_changed = labels that changed;
 * update=1 also returns the per-cluster sums of the centred samples (k, d) and the counts (k).  Blocks. */
int zb200_kmeans_step_f32(const float* d_x, int64_t n, int d, const double* h_mean, const double* h_centres, int k,
                          int32_t* d_labels, int update, double* h_sums, double* h_counts, int64_t* h_changed, void* stream);
/* E-step of a full-covariance Gaussian mixture: h_prec_chol (k, d, d) upper-triangular factors (precision = P P^T),
 * h_means (k, d) absolute.  d_log_resp (n, k) and / or d_labels (argmax) may be NULL; *h_mean_log_norm (may be NULL) =
 * mean over the samples of the log-sum-exp (the lower bound scikit-learn monitors).  Blocks when it is requested. */
int zb200_gmm_estep_f32(const float* d_x, int64_t n, int d, const double* h_log_weights, const double* h_log_dets,
                        const double* h_means, const double* h_prec_chol, int k, double* d_log_resp, int32_t* d_labels,
                        double* h_mean_log_norm, void* stream);
/* M-step accumulators with r = exp(d_log_resp) (or the one-hot of d_labels when d_log_resp is NULL), x centred by
 * h_mean: h_nk (k), h_sx (k, d) = sum r x, h_sxx (k, d, d) = sum r x x^T.  d <= 90.  Blocks. */
int zb200_gmm_mstep_f32(const float* d_x, int64_t n, int d, const double* h_mean, const double* d_log_resp,
                        const int32_t* d_labels, int k, double* h_nk, double* h_sx, double* h_sxx, void* stream);
/* Overlap-add of the patch-SVD denoiser (replaces reconstruct_patches, mtflearn/denoise/_denoise_svd.py:51-70):
 * d_patches double [ny*nx, kh, kw] laid back at start rows d_ys[ny] x start columns d_xs[nx] (row-major grid), summed
 * in that order and divided by the overlap count -> d_img double [H, W]. */
int zb200_overlap_add_f64(const double* d_patches, const int32_t* d_ys, int ny, const int32_t* d_xs, int nx, int kh, int kw,
                          int H, int W, double* d_img, void* stream);
/* Result download: n floats in HBM -> float64 host array (the dtype the reference returns), chunked D2H through
 * pinned staging overlapped with multi-threaded widening.  Ordered after the work queued on `stream`; blocks. */
int zb200_download_as_f64(const float* d_src, int64_t n, double* h_dst, void* stream);
/* element-wise casts between the two supported float types. */
int zb200_cast(int src_dtype, const void* d_in, int dst_dtype, void* d_out, int64_t n, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* ZERNIKE_B200_H_ */
