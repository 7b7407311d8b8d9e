// ke_refine.cu — N1 ("next" row of SURVEY §8f): the refinement the shipped UI runs after a scan.
//
// Replaces the arithmetic of ui.dup_refine_parallel (reference src/ui/dup_refine_parallel.py):
//   tile_ahash_bits   :59-83   convert("L") -> resize((side,side), BILINEAR) -> per-tile mean
//                              threshold -> (grid*tile)^2 bits in (gy, gx, ty, tx) order, little endian
//   tile_hamming      :86-88   popcount of the XOR of two such bit strings
//   _load_small_gray  :203-207 convert("L") -> resize((size,size), BILINEAR)
//   _mae01            :210-212 mean |a-b| / 255 over the two planes
//
// Kernels (all exact integer arithmetic, byte-identical to Pillow):
//   ke_resize_h_kernel   luma (Pillow rgb2l, fused) + horizontal 8bpc taps  [n,h,w,c] -> u8 [n,h,ow]
//   ke_resize_v_kernel   vertical 8bpc taps                                 [n,h,ow]  -> u8 [n,oh,ow]
//   ke_tile_bits_kernel  per-tile sum, pixel*tile^2 > sum  (== pixel > mean), ballot-packed bits
//   ke_bits_hamming_kernel / ke_plane_sad_kernel   pair (ia, ib) -> popcount(xor) / sum |a-b|
// The intermediate [n,h,ow] plane costs ow/(w*c) of the input traffic (8 % at 512x512x3 -> 128).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>
#include <tuple>
#include <vector>

#include "ke_common.cuh"

namespace {

constexpr int kPrec = 22;

struct RTable {  // one Pillow 8bpc tap table on the device
    int32_t* d_kk = nullptr;
    int32_t* d_bounds = nullptr;
    int ksize = 0;
};

__device__ __forceinline__ uint8_t clip8(int32_t v) {
    v >>= kPrec;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

template <int C>
__device__ __forceinline__ uint32_t luma_at(const uint8_t* p) {
    if (C == 1) return p[0];
    return (p[0] * 19595u + p[1] * 38470u + p[2] * 7471u + 0x8000u) >> 16;
}

// one thread per output (image, row y, out column x); consecutive threads = consecutive x
template <int C>
__global__ void __launch_bounds__(256) ke_resize_h_kernel(const uint8_t* __restrict__ img, long long n, int h, int w,
                                                          long long img_stride, long long row_stride,
                                                          const int32_t* __restrict__ kk,
                                                          const int32_t* __restrict__ bounds, int ksize, int ow,
                                                          int identity, uint8_t* __restrict__ out) {
    const long long total = n * h * ow;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % ow);
        const long long t = idx / ow;
        const int y = (int)(t % h);
        const long long im = t / h;
        const uint8_t* row = img + im * img_stride + (long long)y * row_stride;
        if (identity) {  // Pillow skips the pass when the width already matches
            out[idx] = (uint8_t)luma_at<C>(row + (long long)x * C);
            continue;
        }
        const int first = __ldg(bounds + 2 * x), count = __ldg(bounds + 2 * x + 1);
        const int32_t* k = kk + (long long)x * ksize;
        int32_t acc = 1 << (kPrec - 1);
        const uint8_t* p = row + (long long)first * C;
        for (int tp = 0; tp < count; ++tp) acc += (int32_t)luma_at<C>(p + (long long)tp * C) * __ldg(k + tp);
        out[idx] = clip8(acc);
    }
}

__global__ void __launch_bounds__(256) ke_resize_v_kernel(const uint8_t* __restrict__ mid, long long n, int h, int ow,
                                                          const int32_t* __restrict__ kk,
                                                          const int32_t* __restrict__ bounds, int ksize, int oh,
                                                          int identity, uint8_t* __restrict__ out) {
    const long long total = n * oh * ow;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % ow);
        const long long t = idx / ow;
        const int yy = (int)(t % oh);
        const long long im = t / oh;
        const uint8_t* base = mid + im * (long long)h * ow + x;
        if (identity) {
            out[idx] = base[(long long)yy * ow];
            continue;
        }
        const int first = __ldg(bounds + 2 * yy), count = __ldg(bounds + 2 * yy + 1);
        const int32_t* k = kk + (long long)yy * ksize;
        int32_t acc = 1 << (kPrec - 1);
        for (int tp = 0; tp < count; ++tp) acc += (int32_t)base[(long long)(first + tp) * ow] * __ldg(k + tp);
        out[idx] = clip8(acc);
    }
}

// One warp per (image, tile): lanes stride over the tile's pixels; bit index of pixel (gy,gx,ty,tx) is
// ((gy*grid + gx)*tile + ty)*tile + tx, little endian inside 32-bit words.
__global__ void __launch_bounds__(256) ke_tile_bits_kernel(const uint8_t* __restrict__ planes, long long n, int grid,
                                                           int tile, uint32_t* __restrict__ bits) {
    const int side = grid * tile, tiles = grid * grid, tt = tile * tile;
    const int words_per_img = (side * side + 31) / 32;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long job = warp; job < n * tiles; job += n_warps) {
        const long long im = job / tiles;
        const int tl = (int)(job - im * tiles), gy = tl / grid, gx = tl - gy * grid;
        const uint8_t* p = planes + im * (long long)side * side + (long long)gy * tile * side + gx * tile;
        uint32_t sum = 0;
        for (int e = lane; e < tt; e += 32) sum += p[(e / tile) * side + (e % tile)];
#pragma unroll
        for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        // a > mean  <=>  a * tile^2 > sum  (exact; the reference compares uint8 with a float64 mean)
        for (int e0 = 0; e0 < tt; e0 += 32) {
            const int e = e0 + lane;
            const bool on = e < tt && (uint32_t)p[(e / tile) * side + (e % tile)] * (uint32_t)tt > sum;
            const uint32_t word = __ballot_sync(0xffffffffu, on);
            // this ballot covers bit positions [tl*tt + e0, +32) of the image's bit string
            const long long bit0 = (long long)tl * tt + e0;
            const int nvalid = min(32, tt - e0);
            if (lane == 0) {
                uint32_t* dst = bits + im * words_per_img;
                const int wi = (int)(bit0 >> 5), sh = (int)(bit0 & 31);
                const uint32_t mask = nvalid == 32 ? 0xFFFFFFFFu : ((1u << nvalid) - 1u);
                atomicOr(dst + wi, (word & mask) << sh);
                if (sh && sh + nvalid > 32) atomicOr(dst + wi + 1, (word & mask) >> (32 - sh));
            }
        }
    }
}

__global__ void __launch_bounds__(256) ke_bits_hamming_kernel(const uint32_t* __restrict__ bits, int words,
                                                              const long long* __restrict__ ia,
                                                              const long long* __restrict__ ib, long long n_pairs,
                                                              int* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p = warp; p < n_pairs; p += n_warps) {
        const uint32_t* a = bits + ia[p] * words;
        const uint32_t* b = bits + ib[p] * words;
        int d = 0;
        for (int k = lane; k < words; k += 32) d += __popc(a[k] ^ b[k]);
#pragma unroll
        for (int off = 16; off; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
        if (lane == 0) out[p] = d;
    }
}

__global__ void __launch_bounds__(256) ke_plane_sad_kernel(const uint8_t* __restrict__ planes, long long plane_bytes,
                                                           const long long* __restrict__ ia,
                                                           const long long* __restrict__ ib, long long n_pairs,
                                                           unsigned long long* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p = warp; p < n_pairs; p += n_warps) {
        const uint8_t* a = planes + ia[p] * plane_bytes;
        const uint8_t* b = planes + ib[p] * plane_bytes;
        unsigned long long s = 0;
        if ((plane_bytes & 3) == 0) {
            const uint32_t* a4 = reinterpret_cast<const uint32_t*>(a);
            const uint32_t* b4 = reinterpret_cast<const uint32_t*>(b);
            uint32_t part = 0;
            for (long long k = lane; k < plane_bytes / 4; k += 32) part = __vsadu4(a4[k], b4[k]) + part;
            s = part;
        } else {
            for (long long k = lane; k < plane_bytes; k += 32) s += (unsigned)abs((int)a[k] - (int)b[k]);
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) out[p] = s;
    }
}

// ------------------------------------------------------------------ host side

// Pillow precompute_coeffs / normalize_coeffs_8bpc for BILINEAR (filter 2) and LANCZOS (filter 1).
int build_table(int in_size, int out_size, int filter, std::vector<int32_t>& kk, std::vector<int32_t>& bd, int& ksize) {
    const double support0 = filter == 2 ? 1.0 : 3.0;
    const double scale = (double)((float)in_size - 0.0f) / (double)out_size;
    const double fs = scale < 1.0 ? 1.0 : scale;
    const double support = support0 * fs;
    ksize = (int)std::ceil(support) * 2 + 1;
    kk.assign((size_t)ksize * out_size, 0);
    bd.assign(2 * (size_t)out_size, 0);
    std::vector<double> w((size_t)ksize);
    const double inv = 1.0 / fs;
    for (int o = 0; o < out_size; ++o) {
        const double center = (o + 0.5) * scale;
        int first = (int)(center - support + 0.5);
        if (first < 0) first = 0;
        int last = (int)(center + support + 0.5);
        if (last > in_size) last = in_size;
        const int count = last - first;
        double total = 0.0;
        for (int t = 0; t < count; ++t) {
            double x = (t + first - center + 0.5) * inv;
            double v;
            if (filter == 2) {
                if (x < 0.0) x = -x;
                v = x < 1.0 ? 1.0 - x : 0.0;
            } else {
                if (x >= -3.0 && x < 3.0) {
                    auto sinc = [](double z) { return z == 0.0 ? 1.0 : std::sin(z * M_PI) / (z * M_PI); };
                    v = sinc(x) * sinc(x / 3.0);
                } else {
                    v = 0.0;
                }
            }
            w[t] = v;
            total += v;
        }
        for (int t = 0; t < count; ++t) {
            const double v = (total != 0.0 ? w[t] / total : w[t]) * (double)(1 << kPrec);
            kk[(size_t)o * ksize + t] = v < 0 ? (int32_t)(v - 0.5) : (int32_t)(v + 0.5);
        }
        bd[2 * o] = first;
        bd[2 * o + 1] = count;
    }
    return KE_OK;
}

}  // namespace

struct KeResizeCache {
    std::map<std::tuple<int, int, int>, RTable> tables;  // (in, out, filter)
};

void ke_resize_tables_free(KeResizeCache* cache) {
    if (!cache) return;
    for (auto& kv : cache->tables) {
        cudaFree(kv.second.d_kk);
        cudaFree(kv.second.d_bounds);
    }
    delete cache;
}

namespace {

int get_table(ke_ctx* ctx, int in_size, int out_size, int filter, const RTable** out) {
    if (!ctx->resize_tables) ctx->resize_tables = new KeResizeCache();
    auto& tables = ctx->resize_tables->tables;
    auto key = std::make_tuple(in_size, out_size, filter);
    auto it = tables.find(key);
    if (it == tables.end()) {
        std::vector<int32_t> kk, bd;
        RTable t;
        build_table(in_size, out_size, filter, kk, bd, t.ksize);
        KE_CUDA(cudaMalloc((void**)&t.d_kk, kk.size() * 4));
        KE_CUDA(cudaMalloc((void**)&t.d_bounds, bd.size() * 4));
        KE_CUDA(cudaMemcpy(t.d_kk, kk.data(), kk.size() * 4, cudaMemcpyHostToDevice));
        KE_CUDA(cudaMemcpy(t.d_bounds, bd.data(), bd.size() * 4, cudaMemcpyHostToDevice));
        it = tables.emplace(key, t).first;
    }
    *out = &it->second;
    return KE_OK;
}

}  // namespace

int ke_pillow_table(int in_size, int out_size, int filter, std::vector<int32_t>& kk, std::vector<int32_t>& bd, int& ksize) {
    return build_table(in_size, out_size, filter, kk, bd, ksize);
}

namespace {

unsigned grid_for(long long work_items, int per_block, const ke_ctx* ctx) {
    long long blocks = (work_items + per_block - 1) / per_block;
    const long long cap = (long long)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace

extern "C" int ke_gray_resize_batch(ke_ctx* ctx, const uint8_t* d_img, int64_t n, int h, int w, int c,
                                    int64_t img_stride, int64_t row_stride, int out_w, int out_h, int filter,
                                    uint8_t* d_mid, uint8_t* d_out, void* stream) {
    KE_REQUIRE(ctx != nullptr, "ke_gray_resize_batch: ctx is NULL");
    KE_REQUIRE(n >= 0, "ke_gray_resize_batch: n < 0");
    if (n == 0) return KE_OK;
    KE_REQUIRE(d_img && d_mid && d_out, "ke_gray_resize_batch: NULL buffer");
    KE_REQUIRE(h > 0 && w > 0 && out_w > 0 && out_h > 0, "ke_gray_resize_batch: empty geometry");
    KE_REQUIRE(c == 1 || c == 3 || c == 4, "ke_gray_resize_batch: channels must be 1, 3 or 4 (got %d)", c);
    KE_REQUIRE(filter == 1 || filter == 2, "ke_gray_resize_batch: filter must be 1 (LANCZOS) or 2 (BILINEAR)");
    KE_REQUIRE(row_stride >= (int64_t)w * c && img_stride >= (int64_t)(h - 1) * row_stride + (int64_t)w * c,
               "ke_gray_resize_batch: strides smaller than the image");
    KeDeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if (!ctx->force_generic_resize) {  // streaming tensor-pipe kernel for the shapes it takes (ke_resize_mma.cu)
        int taken = 0;
        if ((rc = ke_gray_resize_mma(ctx, d_img, n, h, w, c, img_stride, row_stride, out_w, out_h, filter, d_out, s, &taken)))
            return rc;
        if (taken) return KE_OK;
    }
    const RTable *th, *tv;
    if ((rc = get_table(ctx, w, out_w, filter, &th))) return rc;
    if ((rc = get_table(ctx, h, out_h, filter, &tv))) return rc;
    const unsigned gh = grid_for(n * h * out_w, 256, ctx), gv = grid_for(n * out_h * out_w, 256, ctx);
    const int id_h = w == out_w, id_v = h == out_h;
    switch (c) {
        case 1:
            ke_resize_h_kernel<1><<<gh, 256, 0, s>>>(d_img, n, h, w, img_stride, row_stride, th->d_kk, th->d_bounds,
                                                     th->ksize, out_w, id_h, d_mid);
            break;
        case 3:
            ke_resize_h_kernel<3><<<gh, 256, 0, s>>>(d_img, n, h, w, img_stride, row_stride, th->d_kk, th->d_bounds,
                                                     th->ksize, out_w, id_h, d_mid);
            break;
        default:
            ke_resize_h_kernel<4><<<gh, 256, 0, s>>>(d_img, n, h, w, img_stride, row_stride, th->d_kk, th->d_bounds,
                                                     th->ksize, out_w, id_h, d_mid);
    }
    KE_CUDA(cudaGetLastError());
    ke_resize_v_kernel<<<gv, 256, 0, s>>>(d_mid, n, h, out_w, tv->d_kk, tv->d_bounds, tv->ksize, out_h, id_v, d_out);
    KE_CUDA(cudaGetLastError());
    ctx->launches += 2;
    return KE_OK;
}

extern "C" int ke_tile_ahash_bits(ke_ctx* ctx, const uint8_t* d_planes, int64_t n, int grid, int tile,
                                  uint32_t* d_bits, void* stream) {
    KE_REQUIRE(ctx && (n == 0 || (d_planes && d_bits)), "ke_tile_ahash_bits: NULL argument");
    KE_REQUIRE(n >= 0 && grid > 0 && tile > 0 && (long long)grid * tile <= 4096, "ke_tile_ahash_bits: bad geometry");
    if (n == 0) return KE_OK;
    KeDeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    const long long side = (long long)grid * tile, words = (side * side + 31) / 32;
    KE_CUDA(cudaMemsetAsync(d_bits, 0, (size_t)(n * words * 4), s));
    ke_tile_bits_kernel<<<grid_for(n * grid * grid * 32, 256, ctx), 256, 0, s>>>(d_planes, n, grid, tile, d_bits);
    KE_CUDA(cudaGetLastError());
    ctx->launches++;
    return KE_OK;
}

extern "C" int ke_bits_hamming_pairs(ke_ctx* ctx, const uint32_t* d_bits, int words, const int64_t* d_ia,
                                     const int64_t* d_ib, int64_t n_pairs, int32_t* d_out, void* stream) {
    KE_REQUIRE(ctx && words > 0 && n_pairs >= 0, "ke_bits_hamming_pairs: bad arguments");
    if (n_pairs == 0) return KE_OK;
    KE_REQUIRE(d_bits && d_ia && d_ib && d_out, "ke_bits_hamming_pairs: NULL buffer");
    KeDeviceGuard guard(ctx->device);
    ke_bits_hamming_kernel<<<grid_for(n_pairs * 32, 256, ctx), 256, 0, (cudaStream_t)stream>>>(
        d_bits, words, (const long long*)d_ia, (const long long*)d_ib, n_pairs, d_out);
    KE_CUDA(cudaGetLastError());
    ctx->launches++;
    return KE_OK;
}

extern "C" int ke_plane_sad_pairs(ke_ctx* ctx, const uint8_t* d_planes, int64_t plane_bytes, const int64_t* d_ia,
                                  const int64_t* d_ib, int64_t n_pairs, uint64_t* d_out, void* stream) {
    KE_REQUIRE(ctx && plane_bytes > 0 && n_pairs >= 0, "ke_plane_sad_pairs: bad arguments");
    if (n_pairs == 0) return KE_OK;
    KE_REQUIRE(d_planes && d_ia && d_ib && d_out, "ke_plane_sad_pairs: NULL buffer");
    KeDeviceGuard guard(ctx->device);
    ke_plane_sad_kernel<<<grid_for(n_pairs * 32, 256, ctx), 256, 0, (cudaStream_t)stream>>>(
        d_planes, plane_bytes, (const long long*)d_ia, (const long long*)d_ib, n_pairs, (unsigned long long*)d_out);
    KE_CUDA(cudaGetLastError());
    ctx->launches++;
    return KE_OK;
}

// ------------------------------------------------------------------------------------------
// convert("L") of selected bank images into packed planes (Pillow rgb2l, src/dup/refine.py:48-49): what the
// multi-GPU scan ships between ranks for its cross-shard SSIM pairs (a third of the RGB bytes).

namespace {
// grid = (blocks over one plane, images): 32-bit index arithmetic inside a plane (the first version divided 64-bit element
// indices per thread and ran at a fifth of the HBM rate)
template <int C>
__global__ void __launch_bounds__(256) ke_luma_planes_kernel(const uint8_t* __restrict__ bank, int h, int w, long long img_stride,
                                                             long long row_stride, const long long* __restrict__ idx,
                                                             uint8_t* __restrict__ out) {
    const uint8_t* img = bank + idx[blockIdx.y] * img_stride;
    uint8_t* dst = out + (long long)blockIdx.y * h * w;
    const int per = h * w;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < per; e += gridDim.x * blockDim.x) {
        const int y = e / w, x = e - y * w;
        dst[e] = (uint8_t)luma_at<C>(img + (long long)y * row_stride + (long long)x * C);
    }
}

// RGB rows on 16-byte boundaries, w % 16 == 0: a thread turns 48 bytes (three coalesced 16-byte loads) into 16 luma bytes
__global__ void __launch_bounds__(256) ke_luma_planes_rgb16_kernel(const uint8_t* __restrict__ bank, int h, int w,
                                                                   long long img_stride, long long row_stride,
                                                                   const long long* __restrict__ idx, uint4* __restrict__ out) {
    const uint8_t* img = bank + idx[blockIdx.y] * img_stride;
    const int wg = w >> 4, per = h * wg;
    uint4* dst = out + (long long)blockIdx.y * per;
    auto lum = [](uint32_t r_, uint32_t g_, uint32_t b_) { return (r_ * 19595u + g_ * 38470u + b_ * 7471u + 0x8000u) >> 16; };
    auto four = [&](uint32_t w0, uint32_t w1, uint32_t w2) {
        return lum(w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u) | (lum(w0 >> 24, w1 & 255u, (w1 >> 8) & 255u) << 8) |
               (lum((w1 >> 16) & 255u, w1 >> 24, w2 & 255u) << 16) | (lum((w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24) << 24);
    };
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < per; e += gridDim.x * blockDim.x) {
        const int y = e / wg, g = e - y * wg;
        const uint4* src = reinterpret_cast<const uint4*>(img + (long long)y * row_stride) + 3 * g;
        const uint4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
        dst[e] = make_uint4(four(a.x, a.y, a.z), four(a.w, b.x, b.y), four(b.z, b.w, c.x), four(c.y, c.z, c.w));
    }
}
}  // namespace

extern "C" int ke_luma_planes(ke_ctx* ctx, const uint8_t* d_bank, int h, int w, int c, int64_t img_stride, int64_t row_stride,
                              const int64_t* d_idx, int64_t n, uint8_t* d_out, void* stream) {
    KE_REQUIRE(ctx && n >= 0 && h > 0 && w > 0, "ke_luma_planes: bad arguments");
    KE_REQUIRE(c == 1 || c == 3 || c == 4, "ke_luma_planes: channels must be 1, 3 or 4 (got %d)", c);
    if (n == 0) return KE_OK;
    KE_REQUIRE(d_bank && d_idx && d_out, "ke_luma_planes: NULL buffer");
    KeDeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    KE_REQUIRE(n <= 65535 * 1024ll, "ke_luma_planes: too many planes in one call");
    const long long per = (long long)h * w;
    KE_REQUIRE(per < (1ll << 31), "ke_luma_planes: plane too large");
    for (int64_t k0 = 0; k0 < n; k0 += 65535) {  // gridDim.y limit
        const unsigned ny = (unsigned)std::min<int64_t>(65535, n - k0);
        const long long* idx = (const long long*)d_idx + k0;
        uint8_t* out = d_out + k0 * per;
        if (c == 3 && (w & 15) == 0 && (row_stride & 15) == 0 && (img_stride & 15) == 0 &&
            (reinterpret_cast<uintptr_t>(d_bank) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0) {
            const unsigned nx = (unsigned)std::min<long long>((per / 16 + 255) / 256, 64);
            ke_luma_planes_rgb16_kernel<<<dim3(nx, ny), 256, 0, s>>>(d_bank, h, w, img_stride, row_stride, idx, (uint4*)out);
        } else {
            const unsigned nx = (unsigned)std::min<long long>((per + 255) / 256, 256);
            switch (c) {
                case 1: ke_luma_planes_kernel<1><<<dim3(nx, ny), 256, 0, s>>>(d_bank, h, w, img_stride, row_stride, idx, out); break;
                case 3: ke_luma_planes_kernel<3><<<dim3(nx, ny), 256, 0, s>>>(d_bank, h, w, img_stride, row_stride, idx, out); break;
                default: ke_luma_planes_kernel<4><<<dim3(nx, ny), 256, 0, s>>>(d_bank, h, w, img_stride, row_stride, idx, out);
            }
        }
        ctx->launches++;
    }
    KE_CUDA(cudaGetLastError());
    return KE_OK;
}
