// qt_common.cuh — context object and launch helpers shared by the translation units of libqtrack.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <mutex>

#include "../../include/qtrack.h"

struct qt_context {
    int device;
    cudaStream_t stream;      // stream work is enqueued on
    cudaStream_t own_stream;  // created by qt_create, destroyed by qt_destroy
    cudaStream_t aux_stream;  // second lane of the host-buffer entry points' copy/compute pipeline
    cudaEvent_t ev_inputs, ev_aux_done;  // ordering between the two lanes (timing disabled)
    int sm_count;
    int dare_path;            // QT_DARE_PATH_* (qt_set_dare_path)
    int64_t launches;
    // scratch owned by the library (host-buffer entry points only; never handed to the caller)
    void *scratch_dev;
    size_t scratch_bytes;
    void *sweep_scratch_dev;  // qt_rollout_sweep: metric rows nobody asked for, when the fused reduction cannot be used
    size_t sweep_scratch_bytes;
    void *partials_dev;       // qt_reduce_summary block partials (partials_cap x 16 doubles)
    int partials_cap;
    char err[512];
    // The *_host entry points swap `stream` between two lanes and share scratch_dev: they serialise on this mutex.
    // The *_dev entry points are not locked (they only enqueue on `stream`): use one context per host thread.
    std::mutex host_call_mutex;
};

namespace qt {

inline int fail(qt_context *ctx, int code, const char *what, const char *detail) {
    if (ctx) snprintf(ctx->err, sizeof(ctx->err), "%s: %s", what, detail ? detail : "");
    return code;
}

inline int check_launch(qt_context *ctx, const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, QT_ERR_CUDA, what, cudaGetErrorString(e));
    ctx->launches += 1;
    return QT_OK;
}

#define QT_CUDA_TRY(ctx, expr)                                                    \
    do {                                                                          \
        cudaError_t qt_e_ = (expr);                                               \
        if (qt_e_ != cudaSuccess) return qt::fail((ctx), QT_ERR_CUDA, #expr, cudaGetErrorString(qt_e_)); \
    } while (0)

#define QT_REQUIRE(ctx, cond, msg)                                     \
    do {                                                               \
        if (!(cond)) return qt::fail((ctx), QT_ERR_ARG, "invalid argument", msg); \
    } while (0)

// Grow-only device scratch of the host-buffer entry points (caller holds host_call_mutex).
inline int ensure_scratch(qt_context *ctx, size_t total) {
    if (ctx->scratch_bytes >= total) return QT_OK;
    if (ctx->scratch_dev) {
        cudaError_t e = cudaFree(ctx->scratch_dev);
        if (e != cudaSuccess) return fail(ctx, QT_ERR_CUDA, "cudaFree(scratch)", cudaGetErrorString(e));
    }
    ctx->scratch_dev = nullptr;
    ctx->scratch_bytes = 0;
    cudaError_t e = cudaMalloc(&ctx->scratch_dev, total);
    if (e != cudaSuccess) return fail(ctx, QT_ERR_CUDA, "cudaMalloc(scratch)", cudaGetErrorString(e));
    ctx->scratch_bytes = total;
    return QT_OK;
}

struct DeviceGuard {
    int prev;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace qt
