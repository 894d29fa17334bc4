// K5 -- the final feature gather over NVLink peer memory (the reference is single-process).
// One process per GPU: every rank allocates its copy of the gathered array here (cudaMalloc), exports it as a
// CUDA IPC handle, and maps the other ranks' copies (cudaIpcOpenMemHandle enables peer access lazily).  The
// projection kernel then writes each finished tile of rows into every copy (zb200_project_patches_push_f32:
// P2P stores from a warp of the kernel itself, overlapped with the computation of the next tiles), and row bands
// of the dense map are forwarded by the copy engines (zb200_peer_copy_2d).  The handles travel between the
// processes through whatever the host application uses (torch.distributed in motif_learn_b200.parallel).
#include "zb200_common.cuh"

#include <string.h>

using namespace zb200;

static_assert(sizeof(cudaIpcMemHandle_t) == ZB200_IPC_HANDLE_BYTES, "IPC handle size");

extern "C" int zb200_peer_buffer_alloc(size_t bytes, void** d_ptr, unsigned char* handle) {
    ZB_CHECK_ARG(d_ptr && handle && bytes > 0, "peer_buffer_alloc: bad arguments");
    *d_ptr = nullptr;
    void* ptr = nullptr;
    cudaError_t e = cudaMalloc(&ptr, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("peer_buffer_alloc: cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? ZB200_ENOMEM : ZB200_ECUDA;
    }
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) {
        cudaFree(ptr);
        cudaGetLastError();
        set_error("peer_buffer_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        return ZB200_ECUDA;
    }
    memcpy(handle, &h, sizeof(h));
    *d_ptr = ptr;
    return ZB200_OK;
}

extern "C" int zb200_peer_buffer_free(void* d_ptr) {
    if (d_ptr) ZB_CUDA(cudaFree(d_ptr));
    return ZB200_OK;
}

extern "C" int zb200_peer_buffer_open(const unsigned char* handle, void** d_ptr) {
    ZB_CHECK_ARG(handle && d_ptr, "peer_buffer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("peer_buffer_open: cudaIpcOpenMemHandle failed: %s (the owning process must be alive, on the same "
                  "node, and the GPUs must be peer-accessible)", cudaGetErrorString(e));
        return ZB200_ECUDA;
    }
    *d_ptr = ptr;
    return ZB200_OK;
}

extern "C" int zb200_peer_buffer_close(void* d_ptr) {
    if (d_ptr) ZB_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return ZB200_OK;
}

extern "C" int zb200_peer_copy_2d(void* d_dst, size_t dst_pitch, const void* d_src, size_t src_pitch, size_t width_bytes,
                                  size_t height, void* stream) {
    ZB_CHECK_ARG(d_dst && d_src, "peer_copy_2d: null pointer");
    if (width_bytes == 0 || height == 0) return ZB200_OK;
    ZB_CUDA(cudaMemcpy2DAsync(d_dst, dst_pitch, d_src, src_pitch, width_bytes, height, cudaMemcpyDeviceToDevice,
                              as_stream(stream)));
    return ZB200_OK;
}

extern "C" int zb200_project_patches_push_f32(const zb200_plan* p, const float* d_patches, int64_t n, int precision,
                                              int out_kind, double value_max, void* d_out, void* const* d_out_peers,
                                              int n_peers, void* stream) {
    ZB_CHECK_ARG(p, "project_push: plan is null");
    ZB_CHECK_ARG(n >= 0, "project_push: negative patch count");
    ZB_CHECK_ARG(out_kind >= ZB200_OUT_REAL && out_kind <= ZB200_OUT_ABS, "project_push: out_kind %d not supported", out_kind);
    ZB_CHECK_ARG(n_peers >= 0 && n_peers <= 7 && (n_peers == 0 || d_out_peers), "project_push: 0..7 peers");
    if (n == 0) return ZB200_OK;
    ZB_CHECK_ARG(d_patches && d_out, "project_push: null device pointer");
    if (precision != ZB200_PREC_TF32 && precision != ZB200_PREC_TF32X3 && precision != ZB200_PREC_F16X3) {
        set_error("project_push: the peer push lives in the tensor-core projection kernels (precision tf32 / tf32x3 / f16x3)");
        return ZB200_EUNSUP;
    }
    PeerTargets peers;
    peers.n = n_peers;
    for (int g = 0; g < n_peers; ++g) {
        ZB_CHECK_ARG(d_out_peers[g], "project_push: peer pointer %d is null", g);
        ZB_CHECK_ARG(((reinterpret_cast<uintptr_t>(d_out_peers[g]) ^ reinterpret_cast<uintptr_t>(d_out)) & 15) == 0,
                     "project_push: peer pointer %d and the local pointer must agree modulo 16 bytes", g);
        peers.out[g] = static_cast<float*>(d_out_peers[g]);
    }
    return project_tc(p, d_patches, n, precision, out_kind, d_out, nullptr, nullptr, nullptr, 0, 0, as_stream(stream), nullptr,
                      &peers, value_max);
}
