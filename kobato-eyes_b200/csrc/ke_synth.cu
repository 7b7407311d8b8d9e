// ke_synth.cu — synthetic image generator for bench.py / tests (measurement helper).
// Integer-only twin of kobato_b200/synth.py::synth_image: identical bytes on CPU and GPU.
#include "ke_common.cuh"

namespace {

constexpr unsigned long long K_ITEM = 0x9E3779B97F4A7C15ull;
constexpr unsigned long long K_ELEM = 0xD1B54A32D192ED03ull;
constexpr unsigned long long K_CHAN = 0x8CB92BA72F3D8DD7ull;

__host__ __device__ inline unsigned long long splitmix64(unsigned long long x) {
    unsigned long long z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ inline unsigned long long mix(unsigned long long seed, unsigned long long item,
                                                  unsigned long long elem) {
    return splitmix64(splitmix64(seed ^ (item * K_ITEM)) ^ (elem * K_ELEM));
}

constexpr int G1 = 9, G2 = 33, kSlabRows = 32, kThreads = 256;

__device__ __forceinline__ int lerp_at(const unsigned char* grid, int g, int y, int x, int h, int w) {
    const long long fy = ((long long)y * (g - 1) * 65536) / h, fx = ((long long)x * (g - 1) * 65536) / w;
    const int y0 = (int)(fy >> 16), ty = (int)((fy >> 8) & 255), x0 = (int)(fx >> 16), tx = (int)((fx >> 8) & 255);
    const int g00 = grid[y0 * g + x0], g01 = grid[y0 * g + x0 + 1], g10 = grid[(y0 + 1) * g + x0],
              g11 = grid[(y0 + 1) * g + x0 + 1];
    const int top = g00 * (256 - tx) + g01 * tx, bot = g10 * (256 - tx) + g11 * tx;
    return (top * (256 - ty) + bot * ty) >> 16;
}

__global__ void __launch_bounds__(kThreads) ke_synth_kernel(unsigned char* out, long long start, long long item_stride, int h,
                                                            int w, int c, long long n_set, unsigned long long seed,
                                                            int planted_permille, int slabs) {
    __shared__ unsigned char g1[4][G1 * G1], g2[4][G2 * G2];
    const long long k = blockIdx.x / slabs;
    const int slab = blockIdx.x % slabs;
    const long long i = start + k * item_stride;
    const long long n_base = n_set - (n_set * planted_permille) / 1000;
    long long src = i;
    int variant = 0;
    if (i >= n_base && n_base >= 1) {
        src = (long long)(mix(seed ^ 0x5EEDull, (unsigned long long)i, 1ull) % (unsigned long long)n_base);
        variant = 1 + (int)(mix(seed ^ 0x5EEDull, (unsigned long long)i, 2ull) % 3ull);
    }
    const int shift = variant == 3 ? 1 : 0, amp = variant == 2 ? 3 : 8;
    const unsigned long long noise_item = variant == 2 ? (unsigned long long)i : (unsigned long long)src;
    for (int ch = 0; ch < c; ++ch) {
        for (int e = threadIdx.x; e < G1 * G1; e += kThreads)
            g1[ch][e] = (unsigned char)(mix(seed, (unsigned long long)src,
                                            (unsigned long long)e + 1ull * (1ull << 20) + (unsigned long long)ch * K_CHAN) >> 56);
        for (int e = threadIdx.x; e < G2 * G2; e += kThreads)
            g2[ch][e] = (unsigned char)(mix(seed, (unsigned long long)src,
                                            (unsigned long long)e + 2ull * (1ull << 20) + (unsigned long long)ch * K_CHAN) >> 56);
    }
    __syncthreads();
    const int y_begin = slab * kSlabRows, y_end = min(h, y_begin + kSlabRows);
    unsigned char* img = out + k * ((long long)h * w * c);
    for (int idx = threadIdx.x; idx < (y_end - y_begin) * w; idx += kThreads) {
        const int y = y_begin + idx / w, x = idx % w;
        const int xs = min(x + shift, w - 1);
        const unsigned long long pix = (unsigned long long)y * (unsigned long long)w + (unsigned long long)xs;
        for (int ch = 0; ch < c; ++ch) {
            const int coarse = lerp_at(g1[ch], G1, y, xs, h, w), fine = lerp_at(g2[ch], G2, y, xs, h, w);
            const int base = (coarse * 3 + fine) >> 2;
            const unsigned long long r = mix(seed ^ 0xA5A5ull, noise_item, pix * 4ull + (unsigned long long)ch);
            const int noise = (int)((r >> 40) % (unsigned long long)(2 * amp + 1)) - amp;
            int v = base + noise;
            if (variant == 1) v = (v * 261 + 128) >> 8;
            img[((long long)y * w + x) * c + ch] = (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
}

}  // namespace

extern "C" int ke_synth_images(ke_ctx* ctx, uint8_t* d_out, int64_t start, int64_t item_stride, int64_t count, int h, int w,
                               int c, int64_t n_set, uint64_t seed, int planted_permille, void* stream) {
    KE_REQUIRE(ctx && d_out, "ke_synth_images: NULL argument");
    KE_REQUIRE(count >= 0 && h > 0 && w > 0 && c >= 1 && c <= 4, "ke_synth_images: bad geometry");
    if (count == 0) return KE_OK;
    KeDeviceGuard guard(ctx->device);
    const int slabs = (h + kSlabRows - 1) / kSlabRows;
    KE_REQUIRE(count * slabs < (1ll << 31), "ke_synth_images: too many images in one call");
    ke_synth_kernel<<<(unsigned)(count * slabs), kThreads, 0, (cudaStream_t)stream>>>(
        d_out, start, item_stride, h, w, c, n_set, seed, planted_permille, slabs);
    ctx->launches++;
    KE_CUDA(cudaGetLastError());
    return KE_OK;
}
