// ke_common.cuh — context, error plumbing and small device helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <functional>
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

constexpr int KE_MAX_DEVICES = 16;

struct ke_ctx {
    // A context drives one device.  A multi-device context (ke_ctx_create_multi) is the context of its first device
    // plus one child context per further device: dev_ctx[0] == this, dev_ctx[k] owns device k of the list.  The `d_`
    // entry points act on the context they are handed; the `_host` entry points fan over dev_ctx[0..n_dev).
    int n_dev = 1;
    ke_ctx* dev_ctx[KE_MAX_DEVICES] = {};
    int device = 0;
    int sm_count = 0;
    int64_t launches = 0;
    cudaStream_t copy_stream[2] = {nullptr, nullptr};
    cudaStream_t aux_stream = nullptr;  // second lane for kernels that run concurrently (hybrid join)
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // growable device / pinned scratch owned by the context
    void* d_scratch[16] = {};
    size_t d_scratch_bytes[16] = {};
    void* h_pinned[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t h_pinned_bytes[4] = {0, 0, 0, 0};
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};  // "pinned staging buffer b has been read by the copy engine"
    KeTableCache* tables = nullptr;
    KeResizeCache* resize_tables = nullptr;
    KeResizeMmaCache* resize_mma = nullptr;
    int force_generic_phash = 0;   // KE_OPT_PHASH_GENERIC
    int phash_cfg = 0;             // KE_OPT_PHASH_CFG
    int force_ssim_v1 = 0;         // KE_OPT_SSIM_V1
    int force_generic_resize = 0;  // KE_OPT_RESIZE_GENERIC
    int join_mode = 0;  // KE_OPT_JOIN_MODE: 0 auto, 1 POPC kernel only, 2 hybrid (POPC + bit-sliced), 3 bit-sliced only  // KE_OPT_PHASH_GENERIC: route every geometry through the generic K1 kernel
};

int ke_ctx_scratch(ke_ctx* ctx, int slot, size_t bytes, void** out);
int ke_ctx_pinned(ke_ctx* ctx, int slot, size_t bytes, void** out);
const char* ke_last_error_cstr();

// Run fn(k, dev_ctx[k]) for every device of the context, one host thread per device (the calling thread takes
// device 0; n_use > 0 limits the fan to the first n_use devices).  Returns the first non-zero status; its message becomes the caller's ke_last_error().
int ke_fan_out(ke_ctx* ctx, const std::function<int(int, ke_ctx*)>& fn, int n_use = 0);

// Host -> device copy of `bytes` on `stream`.  Pinned sources (cudaHostAlloc / cudaHostRegister, e.g. torch pinned
// tensors) are DMA'd in place; pageable sources go through the context's two pinned staging buffers in `piece`-byte
// pieces, filled by `threads` host threads, so the copy engine always reads pinned memory.
int ke_h2d_staged(ke_ctx* ctx, void* d_dst, const void* h_src, size_t bytes, cudaStream_t stream);
// the same for `rows` rows of `row_bytes` packed on the host and `dst_pitch` apart on the device
int ke_h2d_staged_2d(ke_ctx* ctx, void* d_dst, size_t dst_pitch, const void* h_src, size_t row_bytes, size_t rows,
                     cudaStream_t stream);
// 1 when the range starts in page-locked host memory
int ke_host_is_pinned(const void* p);

// single-device bodies of the `_host` entry points (the exported functions fan these over the devices)
int ke_hamming_join_host_one(ke_ctx* ctx, const uint64_t* h_hashes, int64_t n, int threshold, uint32_t flags, int band_bits,
                             int band_count, const uint64_t* h_band_allow, int part_index, int part_count,
                             uint32_t** d_i, uint32_t** d_j, uint8_t** d_d, int64_t capacity, int64_t* out_count);
int ke_phash_batch_host_one(ke_ctx* ctx, const uint8_t* h_img, int64_t n, int h, int w, int c, uint64_t* h_phash,
                            uint64_t* h_dhash, float* h_min_margin);
int ke_ssim_pairs_host_one(ke_ctx* ctx, const uint8_t* h_a, const uint8_t* h_b, int64_t n_pairs, int h, int w, int c,
                           int flags, double* h_ssim);

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
