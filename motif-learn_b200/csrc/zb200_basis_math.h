// Unit-disk Zernike basis arithmetic shared by the CUDA generator (zb200_basis.cu)
// and a host-only harness used by the CPU tests (tests/host_tools/basis_host.cpp).
//
// What it computes is fixed by the reference, ZPs._generate_polynomials
// (mtflearn/features/_zps.py:66-90):
//   x_i = linspace(-1,1,k)[i]  (column -> x, row -> y),  rho = sqrt(x^2+y^2),
//   theta = arctan2(y, x),
//   V[n,m] = [rho<=1] * R_n^{|m|}(rho) * sqrt(2(n+1)/(1+[m==0])) * (m<0 ? sin(|m| theta) : cos(m theta))
// HOW it is computed is ours: instead of the reference's float-factorial power sum
// (_zps.py:52-64, which loses ~7e-10 at n_max=20 and breaks down above n_max~30) the
// radial part uses the Jacobi three-term recurrence
//   R_n^m(rho) = rho^m P_s^{(0,m)}(2 rho^2 - 1),  s=(n-m)/2,
// which is stable on [-1,1] (<= ~2.5e-13 from the exact polynomial at n_max=20).
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define ZB_HD __host__ __device__ __forceinline__
#else
#define ZB_HD inline
#endif

namespace zb200 {

// numpy.linspace(-1, 1, k)[i]: arange(i)*step + start with separately rounded
// multiply and add, last element forced to the stop value (numpy/_core/function_base.py).
ZB_HD double grid_coord(int i, int k) {
    if (k == 1) return -1.0;
    if (i == k - 1) return 1.0;
    const double step = 2.0 / (double)(k - 1);
#if defined(__CUDA_ARCH__)
    return __dadd_rn(__dmul_rn((double)i, step), -1.0);   // no FMA contraction: bit-equal to numpy
#else
    volatile double prod = (double)i * step;
    return prod + (-1.0);
#endif
}

// rho = sqrt(x*x + y*y) with every operation rounded separately (numpy semantics), so
// the disk mask rho<=1 ties exactly like the reference (SURVEY 7 "fp64 basis parity").
ZB_HD double grid_rho(double x, double y) {
#if defined(__CUDA_ARCH__)
    return __dsqrt_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
#else
    volatile double xx = x * x;
    volatile double yy = y * y;
    volatile double s = xx + yy;
    return sqrt(s);
#endif
}

// sqrt(2(n+1)/(1+[m==0]))  (_zps.py:83)
ZB_HD double mode_norm(int n, int m) {
    return sqrt((double)(2 * (n + 1)) / (m == 0 ? 2.0 : 1.0));
}

// Iterates R_{am+2s}^{am}(rho), s = 0,1,...  Call next() after reading value().
struct RadialIter {
    double x;        // 2 rho^2 - 1
    double rpow;     // rho^am
    double p_prev;   // P_{s-1}
    double p_cur;    // P_s
    int am;
    int s;
    ZB_HD RadialIter(double rho, int am_) : am(am_), s(0) {
        x = 2.0 * rho * rho - 1.0;
        rpow = 1.0;
        for (int i = 0; i < am_; ++i) rpow *= rho;
        p_prev = 0.0;
        p_cur = 1.0;
    }
    ZB_HD double value() const { return rpow * p_cur; }
    ZB_HD void next() {
        const int k = s + 1;
        double p_next;
        if (k == 1) {
            p_next = 0.5 * ((double)(am + 2) * x - (double)am);
        } else {
            const double b = (double)am;
            const double c = (double)(2 * k + am);
            const double lhs = 2.0 * k * (k + b) * (c - 2.0);
            const double t1 = (c - 1.0) * (c * (c - 2.0) * x - b * b);
            const double t2 = 2.0 * (k - 1.0) * (k + b - 1.0) * c;
            p_next = (t1 * p_cur - t2 * p_prev) / lhs;
        }
        p_prev = p_cur;
        p_cur = p_next;
        s = k;
    }
};

// index of mode (n, m) in ZPs order: j = ((n+2) n + m) / 2  (_zmoments.py:61)
ZB_HD int mode_index(int n, int m) { return ((n + 2) * n + m) / 2; }

}  // namespace zb200
