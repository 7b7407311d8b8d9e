// ke_common.cuh — context, error plumbing and small device helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <map>
#include <utility>
#include <vector>

#include "kobato_b200.h"

void ke_set_error(const char* fmt, ...);

#define KE_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            ke_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return KE_E_CUDA;                                                                  \
        }                                                                                      \
    } while (0)

#define KE_REQUIRE(cond, ...)        \
    do {                             \
        if (!(cond)) {               \
            ke_set_error(__VA_ARGS__); \
            return KE_E_INVALID;     \
        }                            \
    } while (0)

// Device-resident Pillow coefficient tables, cached per image geometry (ke_phash.cu).
struct KeTableCache;
void ke_tables_free(KeTableCache* cache);
// Device-resident tap tables of the generic gray resize (ke_refine.cu), cached per (in, out, filter).
struct KeResizeCache;
void ke_resize_tables_free(KeResizeCache* cache);
// Tables and entry of the streaming tensor-pipe resize (ke_resize_mma.cu); *taken = 0 when the shape is not its.
struct ke_ctx;
struct KeResizeMmaCache;
void ke_resize_mma_tables_free(KeResizeMmaCache* cache);
int ke_gray_resize_mma(ke_ctx* ctx, const uint8_t* d_img, int64_t n, int h, int w, int c, int64_t img_stride,
                       int64_t row_stride, int out_w, int out_h, int filter, uint8_t* d_out, cudaStream_t s, int* taken);

struct ke_ctx {
    int device = 0;
    int sm_count = 0;
    int64_t launches = 0;
    cudaStream_t copy_stream[2] = {nullptr, nullptr};
    cudaStream_t aux_stream = nullptr;  // second lane for kernels that run concurrently (hybrid join)
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // growable device / pinned scratch owned by the context
    void* d_scratch[12] = {};
    size_t d_scratch_bytes[12] = {};
    void* h_pinned[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t h_pinned_bytes[4] = {0, 0, 0, 0};
    KeTableCache* tables = nullptr;
    KeResizeCache* resize_tables = nullptr;
    KeResizeMmaCache* resize_mma = nullptr;
    int force_generic_phash = 0;
    int join_mode = 0;  // KE_OPT_JOIN_MODE: 0 auto, 1 POPC kernel only, 2 hybrid (POPC + bit-sliced), 3 bit-sliced only  // KE_OPT_PHASH_GENERIC: route every geometry through the generic K1 kernel
};

int ke_ctx_scratch(ke_ctx* ctx, int slot, size_t bytes, void** out);
int ke_ctx_pinned(ke_ctx* ctx, int slot, size_t bytes, void** out);

struct KeDeviceGuard {
    int prev = -1;
    explicit KeDeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~KeDeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
