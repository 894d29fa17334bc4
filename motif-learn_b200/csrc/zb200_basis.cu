// K1 -- unit-disk Zernike basis generator (fp64) and GEMM-operand packer.
// Replaces ZPs._generate_polynomials / _radial_polynomial (mtflearn/features/_zps.py:52-90).
// One-off per (n_max, size): the result is cached in HBM inside the plan.
#include "zb200_common.cuh"
#include "zb200_basis_math.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace zb200 {

// One thread per (pixel, |m|).  Each thread walks n = |m|, |m|+2, ... with the Jacobi
// recurrence and writes the cos (m>=0) and sin (m<0) planes.  Planes are zero outside
// the disk (the host memset the buffer before launch, threads outside return early).
__global__ void basis_kernel(double* __restrict__ basis, int n_max, int k) {
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    const int am = blockIdx.y;
    if (pix >= k * k) return;
    const int row = pix / k, col = pix - row * k;
    const double x = grid_coord(col, k), y = grid_coord(row, k);
    const double rho = grid_rho(x, y);
    if (!(rho <= 1.0)) return;
    const double theta = atan2(y, x);            // arctan2(yv, xv), _zps.py:72
    double sn, cs;
    sincos((double)am * theta, &sn, &cs);
    RadialIter it(rho, am);
    const size_t plane = (size_t)k * k;
    for (int n = am; n <= n_max; n += 2) {
        const double r = it.value() * mode_norm(n, am);
        basis[(size_t)mode_index(n, am) * plane + pix] = r * cs;
        if (am > 0) basis[(size_t)mode_index(n, -am) * plane + pix] = r * sn;
        it.next();
    }
}

// round-to-nearest-even to the 11-bit tf32 significand, result kept in an fp32 container
__device__ __forceinline__ float to_tf32_rn(float f) {
    uint32_t u = __float_as_uint(f);
    u += 0x0FFFu + ((u >> 13) & 1u);
    u &= 0xFFFFE000u;
    return __uint_as_float(u);
}

// row_map[r] = source mode of operand row r, or -1 for a zero (padding / m==0 imaginary) row.
__global__ void pack_kernel(const double* __restrict__ basis, const int* __restrict__ row_map,
                            int rows_pad, int kk, int k_pad, double inv_area,
                            float* __restrict__ full, float* __restrict__ hi, float* __restrict__ lo,
                            __nv_bfloat16* __restrict__ cb, float* __restrict__ tr, __half* __restrict__ hb) {
    const int kidx = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (kidx >= k_pad) return;
    const int src = row_map[r];
    double v = 0.0;
    if (src >= 0 && kidx < kk) v = basis[(size_t)src * kk + kidx] * inv_area;
    const float f = (float)v;
    const float h = to_tf32_rn(f);
    const float l = to_tf32_rn((float)(v - (double)h));
    const size_t o = (size_t)r * k_pad + kidx;
    full[o] = f;
    hi[o] = h;
    lo[o] = l;
    // correction operand: 16 bf16 per group of 8 k -- slots 0..7 pair with Xlo, slots 8..15 with Xhi
    const size_t cbo = (size_t)r * 2 * k_pad + (size_t)(kidx >> 3) * 16 + (kidx & 7);
    cb[cbo] = __float2bfloat16_rn(f);
    cb[cbo + 8] = __float2bfloat16_rn((float)(v - (double)h));
    tr[(size_t)kidx * rows_pad + r] = f;
    // fp16-split operand (V itself, not V/area): k-block kb of the row = [32 x b1 | 32 x b2]
    const double raw = (src >= 0 && kidx < kk) ? basis[(size_t)src * kk + kidx] : 0.0;
    const __half b1 = __double2half(raw);
    const size_t hbo = (size_t)r * 2 * k_pad + (size_t)(kidx >> 5) * 64 + (kidx & 31);
    hb[hbo] = b1;
    hb[hbo + 32] = __double2half(raw - (double)__half2float(b1));
}

int launch_basis(zb200_plan* p, cudaStream_t s) {
    const size_t bytes = (size_t)p->n_modes * p->kk * sizeof(double);
    ZB_CUDA(cudaMemsetAsync(p->basis64, 0, bytes, s));
    dim3 grid((unsigned)ceil_div(p->kk, 128), (unsigned)(p->n_max + 1));
    basis_kernel<<<grid, 128, 0, s>>>(p->basis64, p->n_max, p->size);
    ZB_LAUNCHED();
    return ZB200_OK;
}

static int pack_one(zb200_plan* p, Operand& op, const int* h_map, cudaStream_t s) {
    int* d_map = nullptr;
    ZB_CUDA(cudaMalloc(&d_map, sizeof(int) * op.rows_pad));
    ZB_CUDA(cudaMemcpyAsync(d_map, h_map, sizeof(int) * op.rows_pad, cudaMemcpyHostToDevice, s));
    dim3 grid((unsigned)ceil_div(p->k_pad, 128), (unsigned)op.rows_pad);
    pack_kernel<<<grid, 128, 0, s>>>(p->basis64, d_map, op.rows_pad, p->kk, p->k_pad, p->inv_area,
                                     op.full, op.hi, op.lo, reinterpret_cast<__nv_bfloat16*>(op.cb), op.t,
                                     reinterpret_cast<__half*>(op.hb));
    ZB_LAUNCHED();
    ZB_CUDA(cudaStreamSynchronize(s));
    ZB_CUDA(cudaFree(d_map));
    return ZB200_OK;
}

int launch_pack(zb200_plan* p, cudaStream_t s) {
    int map[kMaxModes * 2 + 32];
    // REAL order: operand row r = mode r
    for (int r = 0; r < p->real.rows_pad; ++r) map[r] = r < p->n_modes ? r : -1;
    int rc = pack_one(p, p->real, map, s);
    if (rc) return rc;
    // CPLX order: rows (2c, 2c+1) = ((n,+m), (n,-m)) for the c-th complex mode in
    // nm2j_complex order (n ascending, m = n%2, n%2+2, ..., n), _zmoments.py:71-91,111-132.
    for (int r = 0; r < p->cplx.rows_pad; ++r) map[r] = -1;
    int c = 0;
    for (int n = 0; n <= p->n_max; ++n)
        for (int m = n & 1; m <= n; m += 2, ++c) {
            map[2 * c] = mode_index(n, m);
            map[2 * c + 1] = m > 0 ? mode_index(n, -m) : -1;
        }
    return pack_one(p, p->cplx, map, s);
}

}  // namespace zb200
