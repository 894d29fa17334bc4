// K4 (fp32 SIMT variant) -- dense sliding-window Zernike correlation with halo tiles in
// shared memory and a fused n-fold symmetry-score epilogue.
// Replaces ZPs._transform_fft_convolve (mtflearn/features/_zps.py:159-193) and, when fused,
// zmoments.rot_maps (mtflearn/features/_zmoments.py:420-462):
//   Z[j,y,x] = 1/area * sum_{a,b<k} img0[y-k/2+a, x-k/2+b] * V[j,a,b]     (img0 zero-extended)
//   S[f,y,x] = sum_j w[f,j] Z_j^2 / ||Z_sel||_p^2
// The reference computes the same numbers with M FFT convolutions and flips the sign of odd
// n afterwards; the direct correlation needs neither.
#include "zb200_common.cuh"

namespace zb200 {

constexpr int MS_PX = 128;        // output pixels (one row segment) per CTA
constexpr int MS_MODES = 96;      // operand rows per mode tile
constexpr int MS_PG = 16;         // pixel groups  (8 px each)
constexpr int MS_MG = 12;         // mode groups   (8 modes each)
constexpr int MS_THREADS = MS_PG * MS_MG;   // 192

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// kScores=false: writes moments of mode tile blockIdx.z.  kScores=true: loops all mode tiles
// and writes only the F score planes.
template <bool kScores>
__global__ void __launch_bounds__(MS_THREADS)
map_simt_kernel(const float* __restrict__ img, int H, int W, int row0, int rows,
                const float* __restrict__ Bt, int rows_pad, int n_modes, int k, int k8,
                float* __restrict__ out_moments, float* __restrict__ out_scores,
                const float* __restrict__ w, const unsigned char* __restrict__ sel, int n_folds,
                int norm_kind) {
    extern __shared__ __align__(16) float smem[];
    const int iw = MS_PX + k8 + 8;                 // halo-tile row pitch (floats), multiple of 4
    float* img_s = smem;                           // [k][iw]
    float* bs = smem + (size_t)k * iw;             // [2][k8][MS_MODES]
    float* red = bs + (size_t)2 * k8 * MS_MODES;   // [kMaxFolds+3][MS_PX]   (scores only)

    const int tid = threadIdx.x;
    const int pg = tid & (MS_PG - 1), mg = tid / MS_PG;
    const int x0 = blockIdx.x * MS_PX;
    const int yl = blockIdx.y;                     // local output row
    const int y = row0 + yl;
    const int half = k / 2;

    // ---- halo tile: img_s[a][c] = img0[y-half+a][x0-half+c] -------------------------------
    for (int e = tid; e < k * iw; e += MS_THREADS) {
        const int a = e / iw, c = e - a * iw;
        const int yy = y - half + a, xx = x0 - half + c;
        float v = 0.f;
        if (c < MS_PX + k - 1 && yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(img + (long long)yy * W + xx);
        img_s[e] = v;
    }
    if (kScores)
        for (int e = tid; e < (kMaxFolds + 3) * MS_PX; e += MS_THREADS) red[e] = 0.f;
    // zero the tap padding rows b in [k, k8) of both slab buffers once
    for (int e = tid; e < 2 * (k8 - k) * MS_MODES; e += MS_THREADS) {
        const int buf = e / ((k8 - k) * MS_MODES);
        const int rem = e - buf * (k8 - k) * MS_MODES;
        bs[((size_t)buf * k8 + k) * MS_MODES + rem] = 0.f;
    }

    const int n_tiles = kScores ? (rows_pad + MS_MODES - 1) / MS_MODES : 1;
    for (int tile = 0; tile < n_tiles; ++tile) {
        const int r0 = (kScores ? tile : (int)blockIdx.z) * MS_MODES;

        auto load_slab = [&](int a, int buf) {
            // slab rows b=0..k-1 of window row a: Bt[(a*k+b)*rows_pad + r0 + c], c<96
            float* dst = bs + (size_t)buf * k8 * MS_MODES;
            for (int e = tid; e < k * (MS_MODES / 4); e += MS_THREADS) {
                const int b = e / (MS_MODES / 4), c4 = (e - b * (MS_MODES / 4)) * 4;
                float* d = dst + b * MS_MODES + c4;
                if (r0 + c4 < rows_pad)
                    cp_async16(d, Bt + ((size_t)a * k + b) * rows_pad + r0 + c4);
                else
                    *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            cp_async_commit();
        };

        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

        __syncthreads();          // previous tile's readers are done with bs / img_s is complete
        load_slab(0, 0);
        for (int a = 0; a < k; ++a) {
            const int buf = a & 1;
            cp_async_wait<0>();
            __syncthreads();      // slab a visible; everyone finished slab a-1 (buffer buf^1 free)
            if (a + 1 < k) load_slab(a + 1, buf ^ 1);

            const float* rowp = img_s + (size_t)a * iw + pg * 8;
            const float* bp = bs + (size_t)buf * k8 * MS_MODES + mg * 8;
            float win[16];
            {
                const float4 v0 = *reinterpret_cast<const float4*>(rowp);
                const float4 v1 = *reinterpret_cast<const float4*>(rowp + 4);
                win[0] = v0.x; win[1] = v0.y; win[2] = v0.z; win[3] = v0.w;
                win[4] = v1.x; win[5] = v1.y; win[6] = v1.z; win[7] = v1.w;
            }
            for (int b0 = 0; b0 < k8; b0 += 8) {
                {
                    const float4 v0 = *reinterpret_cast<const float4*>(rowp + b0 + 8);
                    const float4 v1 = *reinterpret_cast<const float4*>(rowp + b0 + 12);
                    win[8] = v0.x; win[9] = v0.y; win[10] = v0.z; win[11] = v0.w;
                    win[12] = v1.x; win[13] = v1.y; win[14] = v1.z; win[15] = v1.w;
                }
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const float4 w0 = *reinterpret_cast<const float4*>(bp + (size_t)(b0 + t) * MS_MODES);
                    const float4 w1 = *reinterpret_cast<const float4*>(bp + (size_t)(b0 + t) * MS_MODES + 4);
                    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(win[t + i], wv[j], acc[i][j]);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) win[i] = win[i + 8];
            }
        }

        if (!kScores) {
            // moments: out[(r0+8mg+j)][yl][x0+8pg+i]
            const int xb = x0 + pg * 8;
            const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(out_moments) & 15) == 0) && xb + 7 < W;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int mode = r0 + mg * 8 + j;
                if (mode >= n_modes) continue;
                float* dst = out_moments + ((size_t)mode * rows + yl) * W + xb;
                if (vec) {
                    *reinterpret_cast<float4*>(dst) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
                    *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4][j], acc[5][j], acc[6][j], acc[7][j]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (xb + i < W) dst[i] = acc[i][j];
                }
            }
        } else {
            // fused scores: deterministic accumulation over the 12 mode groups into red[q][px]
            for (int g = 0; g < MS_MG; ++g) {
                if (mg == g) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int px = pg * 8 + i;
                        float s1 = 0.f, s2 = 0.f, sm = 0.f;
                        float num[kMaxFolds];
                        for (int f = 0; f < n_folds; ++f) num[f] = 0.f;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int mode = r0 + mg * 8 + j;
                            if (mode < n_modes && sel[mode]) {
                                const float z = acc[i][j], z2 = z * z;
                                s1 += fabsf(z);
                                s2 += z2;
                                sm = fmaxf(sm, fabsf(z));
                                for (int f = 0; f < n_folds; ++f) num[f] = fmaf(w[f * rows_pad + mode], z2, num[f]);
                            }
                        }
                        red[0 * MS_PX + px] += s1;
                        red[1 * MS_PX + px] += s2;
                        red[2 * MS_PX + px] = fmaxf(red[2 * MS_PX + px], sm);
                        for (int f = 0; f < n_folds; ++f) red[(3 + f) * MS_PX + px] += num[f];
                    }
                }
                __syncthreads();
            }
        }
    }

    if (kScores && tid < MS_PX) {
        const int x = x0 + tid;
        if (x < W) {
            float den = 1.f;
            if (norm_kind == ZB200_NORM_L1) den = red[0 * MS_PX + tid] * red[0 * MS_PX + tid];
            else if (norm_kind == ZB200_NORM_L2) den = red[1 * MS_PX + tid];
            else if (norm_kind == ZB200_NORM_INF) den = red[2 * MS_PX + tid] * red[2 * MS_PX + tid];
            for (int f = 0; f < n_folds; ++f)
                out_scores[((size_t)f * rows + yl) * W + x] = red[(3 + f) * MS_PX + tid] / den;
        }
    }
}

int map_simt(const zb200_plan* p, const float* d_img, int H, int W, int row0, int rows,
             float* d_moments, float* d_scores, const float* d_w, const uint8_t* d_sel, int n_folds,
             int norm_kind, cudaStream_t s) {
    if (rows == 0) return ZB200_OK;
    const int k = p->size, k8 = round_up(k, 8);
    const int iw = MS_PX + k8 + 8;
    const bool scores = d_scores != nullptr;
    size_t smem = ((size_t)k * iw + (size_t)2 * k8 * MS_MODES) * sizeof(float);
    if (scores) smem += (size_t)(kMaxFolds + 3) * MS_PX * sizeof(float);
    if (smem > 227 * 1024) {
        set_error("dense map: window %d needs %zu B of shared memory (> 227 KB)", k, smem);
        return ZB200_EUNSUP;
    }
    dim3 grid((unsigned)ceil_div(W, MS_PX), (unsigned)rows,
              scores ? 1u : (unsigned)ceil_div(p->real.rows_pad, MS_MODES));
    if (scores) {
        ZB_CUDA(cudaFuncSetAttribute(map_simt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        map_simt_kernel<true><<<grid, MS_THREADS, smem, s>>>(d_img, H, W, row0, rows, p->real.t, p->real.rows_pad,
                                                             p->n_modes, k, k8, nullptr, d_scores, d_w, d_sel,
                                                             n_folds, norm_kind);
    } else {
        ZB_CUDA(cudaFuncSetAttribute(map_simt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        map_simt_kernel<false><<<grid, MS_THREADS, smem, s>>>(d_img, H, W, row0, rows, p->real.t, p->real.rows_pad,
                                                              p->n_modes, k, k8, d_moments, nullptr, nullptr,
                                                              nullptr, 0, 0);
    }
    ZB_LAUNCHED();
    return ZB200_OK;
}

}  // namespace zb200
