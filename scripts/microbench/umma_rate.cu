// Micro-benchmark: issue cost of tcgen05.mma at small N on B200 (one CTA per SM, 148 CTAs).
// Measures cycles per MMA for different operand sources (SS: A from smem, TS: A from TMEM), kinds,
// N, and accumulator switching patterns.  Operands are garbage; only timing matters.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t a) {
    uint64_t d = 0;
    d |= (uint64_t)((a & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t idesc(int n, int fmt) {   // fmt 2 = tf32, 1 = bf16
    uint32_t d = 0;
    d |= 1u << 4;
    d |= (uint32_t)fmt << 7;
    d |= (uint32_t)fmt << 10;
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(128 >> 4) << 24;
    return d;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, int tf32) {
    if (tf32)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(id) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id, int tf32) {
    if (tf32)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(id) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(id) : "memory");
}

// mode bits: 0 = A from TMEM; 1 = bf16 kind (else tf32); 2 = alternate kinds per MMA;
// n_acc = number of accumulators cycled; group = consecutive MMAs on the same accumulator
__global__ void __launch_bounds__(128, 1) bench(int n, int mode, int n_acc, int group, int iters, unsigned long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 255);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (warp == 1 && lane == 0) {
        const uint64_t da = make_desc_sw128(smem_u32(smem));
        const uint64_t db = make_desc_sw128(smem_u32(smem + 16384));
        const uint32_t a_t = tmem + 448;                     // A staging columns (garbage)
        const long long t0 = clock64();
        int cnt = 0;
        for (int it = 0; it < iters; ++it) {
            for (int acc = 0; acc < n_acc; ++acc) {
                for (int g = 0; g < group; ++g, ++cnt) {
                    const int k4 = g & 3;
                    int tf32 = (mode & 2) ? 0 : 1;
                    if (mode & 4) tf32 = cnt & 1;
                    const uint32_t id = idesc(n, tf32 ? 2 : 1);
                    const uint32_t d = tmem + acc * n;
                    if (mode & 1) mma_ts(d, a_t + k4 * 8, db + 2 * k4, id, tf32);
                    else mma_ss(d, da + 2 * k4, db + 2 * k4, id, tf32);
                }
            }
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        const long long t2 = clock64();
        if (blockIdx.x == 0) {
            out[0] = (unsigned long long)(t1 - t0);
            out[1] = (unsigned long long)(t2 - t0);
            out[2] = (unsigned long long)cnt;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// Lean issue loop: descriptors pre-built, 8 MMAs fully unrolled per iteration, one predicate.
template <int kMode>   // 0 SS same D, 1 TS same D, 2 SS two D alternating, 3 SS: 4 on D0 then 4 on D1
__global__ void __launch_bounds__(128, 1) bench_lean(int n, int iters, unsigned long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 255);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (warp == 1 && lane == 0) {
        uint64_t da = make_desc_sw128(smem_u32(smem));
        if (kMode == 4) {   // un-swizzled K-major, rows 16 B apart (overlapping): LBO 16 B, SBO 128 B
            da = (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)8 << 32) | ((uint64_t)1 << 46);
        }
        if (kMode == 5) {   // un-swizzled K-major canonical: LBO 128 B, SBO 256 B
            da = (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4) | ((uint64_t)8 << 16) | ((uint64_t)16 << 32) | ((uint64_t)1 << 46);
        }
        const uint64_t db = make_desc_sw128(smem_u32(smem + 16384));
        const uint32_t a_t = tmem + 448;
        const uint32_t id = idesc(n, 2);
        const uint32_t d0 = tmem, d1 = tmem + n;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const int k4 = g & 3;
                const uint32_t d = (kMode == 2) ? ((g & 1) ? d1 : d0) : ((kMode == 3) ? ((g & 4) ? d1 : d0) : d0);
                if (kMode == 1) mma_ts(d, a_t + k4 * 8, db + 2 * k4, id, 1);
                else mma_ss(d, da + 2 * k4, db + 2 * k4, id, 1);
            }
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        const long long t2 = clock64();
        if (blockIdx.x == 0) {
            out[0] = (unsigned long long)(t1 - t0);
            out[1] = (unsigned long long)(t2 - t0);
            out[2] = (unsigned long long)iters * 8;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

template <int kMode>
void run_lean(const char* name, int n, unsigned long long* d) {
    cudaFuncSetAttribute(bench_lean<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int rep = 0; rep < 2; ++rep) {
        bench_lean<kMode><<<148, 128, 64 * 1024>>>(n, 400, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
    }
    unsigned long long h[3];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("LEAN %-34s N=%3d  issue %.1f clk/MMA   complete %.1f clk/MMA\n", name, n, (double)h[0] / h[2], (double)h[1] / h[2]);
}

int main() {
    unsigned long long* d;
    cudaMalloc(&d, 64);
    for (int n : {32, 64, 96, 128, 192, 240, 256}) {
        run_lean<0>("SS same D", n, d);
        run_lean<1>("TS same D", n, d);
    }
    run_lean<5>("SS A un-swizzled canonical", 96, d);
    run_lean<4>("SS A Toeplitz (rows 16 B apart)", 96, d);
    for (int n : {96, 192}) {
        run_lean<2>("SS 2 D alternating", n, d);
        run_lean<3>("SS 4 on D0 then 4 on D1", n, d);
    }
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    struct Cfg { int n, mode, n_acc, group; const char* name; };
    Cfg cfgs[] = {
        {96, 0, 1, 8, "SS tf32 N=96 same D"},      {96, 0, 2, 4, "SS tf32 N=96 2 D, groups of 4"},
        {96, 0, 2, 1, "SS tf32 N=96 2 D alternating"}, {96, 0, 4, 1, "SS tf32 N=96 4 D alternating"},
        {96, 1, 1, 8, "TS tf32 N=96 same D"},      {96, 1, 2, 1, "TS tf32 N=96 2 D alternating"},
        {96, 2, 1, 8, "SS bf16 N=96 same D"},      {96, 3, 1, 8, "TS bf16 N=96 same D"},
        {96, 4, 1, 8, "SS tf32/bf16 alternating kinds N=96 same D"}, {96, 5, 1, 8, "TS tf32/bf16 alternating kinds N=96"},
        {192, 0, 1, 8, "SS tf32 N=192 same D"},    {192, 0, 2, 4, "SS tf32 N=192 2 D groups of 4"},
        {256, 0, 1, 8, "SS tf32 N=256 same D"},    {256, 1, 1, 8, "TS tf32 N=256 same D"},
        {128, 0, 1, 8, "SS tf32 N=128 same D"},    {64, 0, 1, 8, "SS tf32 N=64 same D"},
        {32, 0, 1, 8, "SS tf32 N=32 same D"},      {256, 2, 1, 8, "SS bf16 N=256 same D"},
        {112, 0, 1, 8, "SS tf32 N=112 same D"},    {240, 0, 1, 8, "SS tf32 N=240 same D"},
    };
    for (auto& c : cfgs) {
        for (int rep = 0; rep < 2; ++rep) {
            bench<<<148, 128, 64 * 1024>>>(c.n, c.mode, c.n_acc, c.group, 200, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
        }
        unsigned long long h[3];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%-48s issue %.1f clk/MMA   complete %.1f clk/MMA   (%llu MMAs)\n", c.name, (double)h[0] / h[2], (double)h[1] / h[2], h[2]);
    }
    return 0;
}
