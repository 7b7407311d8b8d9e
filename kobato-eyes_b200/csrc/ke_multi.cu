// ke_multi.cu — single-process multi-device plumbing behind the `_host` entry points.
//
// the reference drives the whole path from ONE Qt worker thread of ONE process
// (src/ui/dup_tab.py:118), so the drop-in seams (DuplicateScanner behind DupViewModel(scanner_factory=...),
// core.fastsig.compute_signatures_mp, dup.refine.refine_pairs_batch) can only use several GPUs if the library
// fans the work out itself.  A context created with ke_ctx_create_multi owns one child context per device;
// the `_host` entry points split their units over the devices with one host thread per device:
//     K1  images   -> contiguous image ranges, no exchange
//     K2  triangle -> tiles t % (part_count * n_dev), the (small) table uploaded to every device, candidate
//                     lists concatenated on the host
//     K3  pairs    -> contiguous pair ranges, no exchange
// There is no device-to-device traffic on this path, hence no collective.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>

#include "ke_common.cuh"

int ke_fan_out(ke_ctx* ctx, const std::function<int(int, ke_ctx*)>& fn, int n_use) {
    const int n = n_use > 0 ? std::min(n_use, ctx->n_dev) : ctx->n_dev;
    if (n <= 1) return fn(0, ctx);
    std::vector<int> rc((size_t)n, KE_OK);
    std::vector<std::string> msg((size_t)n);
    std::vector<std::thread> pool;
    pool.reserve((size_t)n - 1);
    auto body = [&](int k) {
        rc[(size_t)k] = fn(k, ctx->dev_ctx[k]);
        if (rc[(size_t)k] != KE_OK) msg[(size_t)k] = ke_last_error_cstr();  // thread-local: hand it to the caller below
    };
    for (int k = 1; k < n; ++k) pool.emplace_back(body, k);
    body(0);
    for (auto& t : pool) t.join();
    for (int k = 0; k < n; ++k)
        if (rc[(size_t)k] != KE_OK) {
            ke_set_error("device %d: %s", ctx->dev_ctx[k]->device, msg[(size_t)k].c_str());
            return rc[(size_t)k];
        }
    return KE_OK;
}

int ke_host_is_pinned(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();  // unregistered host memory reports an error on some drivers: clear it
        return 0;
    }
    return attr.type == cudaMemoryTypeHost ? 1 : 0;
}

namespace {

constexpr size_t kStagePiece = 32u << 20;

int stage_threads() {
    static int n = [] {
        const char* env = getenv("KE_STAGE_THREADS");
        const int hw = (int)std::thread::hardware_concurrency();
        // measured (tools/probe_host_path.py, 512x512 RGB from pageable memory): 2 threads 19.6, 4 threads 29.5, 8 threads
        // 35.3 GB/s against 55.5 GB/s from page-locked memory: the host-side copy into the staging buffers is the limit
        int v = env ? atoi(env) : std::max(2, std::min(8, hw / 2));
        if (hw > 0 && v > hw) v = hw;
        return v < 1 ? 1 : v;
    }();
    return n;
}

void parallel_memcpy(void* dst, const void* src, size_t bytes, int threads) {
    if (threads <= 1 || bytes < (4u << 20)) {
        memcpy(dst, src, bytes);
        return;
    }
    const size_t slice = (bytes / (size_t)threads + 4095) & ~(size_t)4095;
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) {
        const size_t lo = std::min(bytes, slice * (size_t)t), hi = std::min(bytes, lo + slice);
        if (hi > lo) pool.emplace_back([=] { memcpy((char*)dst + lo, (const char*)src + lo, hi - lo); });
    }
    memcpy(dst, src, std::min(bytes, slice));
    for (auto& t : pool) t.join();
}

}  // namespace

int ke_h2d_staged(ke_ctx* ctx, void* d_dst, const void* h_src, size_t bytes, cudaStream_t stream) {
    if (bytes == 0) return KE_OK;
    if (ke_host_is_pinned(h_src)) {
        KE_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, stream));
        return KE_OK;
    }
    void* stage[2];
    int rc;
    for (int b = 0; b < 2; ++b)
        if ((rc = ke_ctx_pinned(ctx, b, kStagePiece, &stage[b]))) return rc;
    const int threads = stage_threads();
    int k = 0;
    for (size_t off = 0; off < bytes; off += kStagePiece, ++k) {
        const int b = k & 1;
        const size_t len = std::min(kStagePiece, bytes - off);
        KE_CUDA(cudaEventSynchronize(ctx->stage_ev[b]));  // the copy engine is done with this buffer (no-op when never recorded)
        parallel_memcpy(stage[b], (const char*)h_src + off, len, threads);
        KE_CUDA(cudaMemcpyAsync((char*)d_dst + off, stage[b], len, cudaMemcpyHostToDevice, stream));
        KE_CUDA(cudaEventRecord(ctx->stage_ev[b], stream));
    }
    return KE_OK;
}

int ke_h2d_staged_2d(ke_ctx* ctx, void* d_dst, size_t dst_pitch, const void* h_src, size_t row_bytes, size_t rows,
                     cudaStream_t stream) {
    if (rows == 0 || row_bytes == 0) return KE_OK;
    if (dst_pitch == row_bytes) return ke_h2d_staged(ctx, d_dst, h_src, row_bytes * rows, stream);
    if (ke_host_is_pinned(h_src)) {
        KE_CUDA(cudaMemcpy2DAsync(d_dst, dst_pitch, h_src, row_bytes, row_bytes, rows, cudaMemcpyHostToDevice, stream));
        return KE_OK;
    }
    int rc;
    if (row_bytes > kStagePiece) {  // rows longer than a staging piece: one flat staged copy per row
        for (size_t r = 0; r < rows; ++r)
            if ((rc = ke_h2d_staged(ctx, (char*)d_dst + r * dst_pitch, (const char*)h_src + r * row_bytes, row_bytes, stream)))
                return rc;
        return KE_OK;
    }
    void* stage[2];
    for (int b = 0; b < 2; ++b)
        if ((rc = ke_ctx_pinned(ctx, b, kStagePiece, &stage[b]))) return rc;
    const size_t per = kStagePiece / row_bytes;
    const int threads = stage_threads();
    int k = 0;
    for (size_t r0 = 0; r0 < rows; r0 += per, ++k) {
        const int b = k & 1;
        const size_t cnt = std::min(per, rows - r0);
        KE_CUDA(cudaEventSynchronize(ctx->stage_ev[b]));
        parallel_memcpy(stage[b], (const char*)h_src + r0 * row_bytes, cnt * row_bytes, threads);
        KE_CUDA(cudaMemcpy2DAsync((char*)d_dst + r0 * dst_pitch, dst_pitch, stage[b], row_bytes, row_bytes, cnt,
                                  cudaMemcpyHostToDevice, stream));
        KE_CUDA(cudaEventRecord(ctx->stage_ev[b], stream));
    }
    return KE_OK;
}

// ------------------------------------------------------------------------------------------
// K2

extern "C" int ke_hamming_join_host(ke_ctx* ctx, const uint64_t* h_hashes, int64_t n, int threshold, uint32_t flags,
                                    int band_bits, int band_count, const uint64_t* h_band_allow, int part_index,
                                    int part_count, uint32_t* h_out_i, uint32_t* h_out_j, uint8_t* h_out_dist,
                                    int64_t capacity, int64_t* out_count) {
    KE_REQUIRE(ctx != nullptr && out_count != nullptr, "ke_hamming_join_host: NULL argument");
    KE_REQUIRE(n >= 0 && capacity >= 0, "ke_hamming_join_host: negative size");
    KE_REQUIRE(n == 0 || h_hashes != nullptr, "ke_hamming_join_host: h_hashes is NULL");
    KE_REQUIRE(part_count >= 1 && part_index >= 0 && part_index < part_count, "ke_hamming_join_host: bad partition %d/%d",
               part_index, part_count);
    *out_count = 0;
    // one more device per ~5e9 pairs of the caller's share (about 2 ms of one GPU): below that the per-device table upload
    // and launch cost more than the split saves
    int nd = ctx->n_dev;
    {
        const double pairs = 0.5 * (double)n * (double)(n > 0 ? n - 1 : 0) / (double)part_count;
        while (nd > 1 && pairs / nd < 5e9) --nd;
    }
    uint32_t *d_i[KE_MAX_DEVICES] = {}, *d_j[KE_MAX_DEVICES] = {};
    uint8_t* d_d[KE_MAX_DEVICES] = {};
    int64_t cnt[KE_MAX_DEVICES] = {};
    // device k takes the tiles t with t % (part_count * nd) == part_index + part_count * k: the caller's share, dealt on
    int rc = ke_fan_out(ctx, [&](int k, ke_ctx* c) {
        return ke_hamming_join_host_one(c, h_hashes, n, threshold, flags, band_bits, band_count, h_band_allow,
                                        part_index + part_count * k, part_count * nd, &d_i[k], &d_j[k], &d_d[k], capacity,
                                        &cnt[k]);
    }, nd);
    if (rc) return rc;
    int64_t total = 0, off[KE_MAX_DEVICES + 1] = {};
    for (int k = 0; k < nd; ++k) {
        off[k] = total;
        total += cnt[k];
    }
    *out_count = total;
    rc = ke_fan_out(ctx, [&](int k, ke_ctx* c) -> int {
        const int64_t room = capacity - std::min(capacity, off[k]);
        const size_t take = (size_t)std::min(cnt[k], room);
        if (!take) return (int)KE_OK;
        KeDeviceGuard guard(c->device);
        cudaStream_t s = c->copy_stream[0];
        KE_CUDA(cudaMemcpyAsync(h_out_i + off[k], d_i[k], take * 4, cudaMemcpyDeviceToHost, s));
        KE_CUDA(cudaMemcpyAsync(h_out_j + off[k], d_j[k], take * 4, cudaMemcpyDeviceToHost, s));
        KE_CUDA(cudaMemcpyAsync(h_out_dist + off[k], d_d[k], take, cudaMemcpyDeviceToHost, s));
        KE_CUDA(cudaStreamSynchronize(s));
        return (int)KE_OK;
    }, nd);
    if (rc) return rc;
    if (total > capacity) {
        ke_set_error("ke_hamming_join_host: %lld pairs qualify but capacity is %lld", (long long)total, (long long)capacity);
        return KE_E_CAPACITY;
    }
    return KE_OK;
}

// ------------------------------------------------------------------------------------------
// K1

extern "C" int ke_phash_batch_host(ke_ctx* ctx, const uint8_t* h_img, int64_t n, int h, int w, int c,
                                   uint64_t* h_phash, uint64_t* h_dhash, float* h_min_margin) {
    KE_REQUIRE(ctx != nullptr, "ke_phash_batch_host: ctx is NULL");
    KE_REQUIRE(n >= 0, "ke_phash_batch_host: n < 0");
    if (n == 0) return KE_OK;
    KE_REQUIRE(h_img && h_phash && h_dhash, "ke_phash_batch_host: NULL buffer");
    KE_REQUIRE(h > 0 && w > 0 && (c == 1 || c == 3 || c == 4), "ke_phash_batch_host: bad geometry %dx%dx%d", w, h, c);
    const int64_t img_bytes = (int64_t)h * w * c;
    // more devices than a few hundred MB of images each buys nothing: the per-device launch + table set-up would dominate
    int nd = ctx->n_dev;
    while (nd > 1 && n * img_bytes / nd < (64ll << 20)) --nd;
    if (nd == 1) return ke_phash_batch_host_one(ctx, h_img, n, h, w, c, h_phash, h_dhash, h_min_margin);
    return ke_fan_out(ctx, [&](int k, ke_ctx* cdev) {
        const int64_t lo = n * k / nd, hi = n * (k + 1) / nd;
        return ke_phash_batch_host_one(cdev, h_img + lo * img_bytes, hi - lo, h, w, c, h_phash + lo, h_dhash + lo,
                                       h_min_margin ? h_min_margin + lo : nullptr);
    }, nd);  // the first nd devices only
}

// ------------------------------------------------------------------------------------------
// K3

extern "C" int ke_ssim_pairs_host(ke_ctx* ctx, const uint8_t* h_a, const uint8_t* h_b, int64_t n_pairs, int h, int w, int c,
                                  int gaussian, double* h_ssim) {
    KE_REQUIRE(ctx != nullptr, "ke_ssim_pairs_host: ctx is NULL");
    KE_REQUIRE(n_pairs >= 0, "ke_ssim_pairs_host: n_pairs < 0");
    if (n_pairs == 0) return KE_OK;
    KE_REQUIRE(h_a && h_b && h_ssim, "ke_ssim_pairs_host: NULL buffer");
    const int64_t plane = (int64_t)h * w * c;
    int nd = ctx->n_dev;
    while (nd > 1 && 2 * n_pairs * plane / nd < (32ll << 20)) --nd;
    if (nd == 1) return ke_ssim_pairs_host_one(ctx, h_a, h_b, n_pairs, h, w, c, gaussian, h_ssim);
    return ke_fan_out(ctx, [&](int k, ke_ctx* cdev) {
        const int64_t lo = n_pairs * k / nd, hi = n_pairs * (k + 1) / nd;
        return ke_ssim_pairs_host_one(cdev, h_a + lo * plane, h_b + lo * plane, hi - lo, h, w, c, gaussian, h_ssim + lo);
    }, nd);
}
