// ke_ssim.cu — K3: batched SSIM on candidate pairs, sm_100a.
//
// Replaces skimage.metrics.structural_similarity(a, b, data_range=1.0) as called by
// dup.refine._compute_ssim (reference src/dup/refine.py:44-52): 7x7 UNIFORM window, sample
// covariance (x49/48), K1=.01, K2=.03, 3-pixel border cropped, mean of the SSIM map.
//
// Arithmetic: the five window sums (u, v, u^2, v^2, uv over 49 pixels) are EXACT integers
// (they fit int32), the variance/covariance numerators 49*Suu - Su^2 ... are exact integers too,
// so there is no cancellation error; only the final ratio is FP32 and the mean FP64.
// With N=49 and pixel scale 255 everything is kept multiplied by 49^2*255^2 (means) or
// 48*49*255^2 (covariances):
//     S = (2*Su*Sv + c1) (2*(49*Suv - Su*Sv) + c2) / ((Su^2 + Sv^2 + c1) (49*(Suu+Svv) - Su^2 - Sv^2 + c2))
//     c1 = 1e-4 * 49^2 * 255^2,  c2 = 9e-4 * 48*49 * 255^2.
//
// Schedule: a work unit is (pair, block of 256 output columns).  Thread = output column.
// Rows stream through shared memory in strips (one 1-D TMA bulk copy per image per strip when
// the plane is contiguous, double buffered).  Per row a thread forms the horizontal 7-sums of its
// window straight from the packed bytes with dp4a (4+3 bytes, no sliding dependency), then keeps
// the vertical 7-row running sums in registers with a 7-deep register ring: no inter-thread
// exchange at all.
//
// Algorithmic HBM bytes per pair: 2*h*w*c read + 8 written.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "ke_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kStripRows = 28;        // multiple of 7 (ring unroll)
constexpr int kBlockCols = kThreads;  // output columns per unit
constexpr int kWin = 7;

struct SsimArgs {
    const uint8_t* bank;
    int h, w, c;
    long long img_stride, row_stride;
    const long long* ia;
    const long long* ib;
    long long n_pairs;
    int n_cblocks;
    int pitch;     // shared-memory row pitch in bytes (multiple of 4)
    int use_bulk;  // contiguous 'L' planes, one column block: bulk copies straight into the strip
    int rgb_words; // RGB bank whose rows and images start on 4-byte boundaries: vectorised luma staging
    double inv_count;
    double* out;
    double* partial;  // [n_pairs][n_cblocks] column-block sums when a plane spans several column blocks
};

__device__ __forceinline__ uint32_t dp4a_uu(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct HSum {
    uint32_t s;   // Su | Sv << 16 (each <= 7*255)
    uint32_t t;   // Suu + Svv (only their sum enters the formula)
    uint32_t uv;  // Suv
};

// Horizontal 7-sums of the window starting at byte `k` (0..3) of word w0.
__device__ __forceinline__ HSum hsum7(const uint32_t* __restrict__ ru, const uint32_t* __restrict__ rv, uint32_t sel) {
    const uint32_t u0 = ru[0], u1 = ru[1], u2 = ru[2];
    const uint32_t v0 = rv[0], v1 = rv[1], v2 = rv[2];
    const uint32_t ua = __byte_perm(u0, u1, sel), ub = __byte_perm(u1, u2, sel);
    const uint32_t va = __byte_perm(v0, v1, sel), vb = __byte_perm(v1, v2, sel);
    const uint32_t ubm = ub & 0x00FFFFFFu, vbm = vb & 0x00FFFFFFu;
    HSum r;
    const uint32_t su = dp4a_uu(ua, 0x01010101u, dp4a_uu(ub, 0x00010101u, 0u));
    const uint32_t sv = dp4a_uu(va, 0x01010101u, dp4a_uu(vb, 0x00010101u, 0u));
    r.s = su + (sv << 16);
    r.t = dp4a_uu(ua, ua, dp4a_uu(ub, ubm, dp4a_uu(va, va, dp4a_uu(vb, vbm, 0u))));
    r.uv = dp4a_uu(ua, va, dp4a_uu(ub, vbm, 0u));
    return r;
}

__device__ __forceinline__ float ssim_point(uint32_t s, uint32_t st, uint32_t suv) {
    constexpr float C1 = 1e-4f * 49.0f * 49.0f * 255.0f * 255.0f;
    constexpr float C2 = 9e-4f * 48.0f * 49.0f * 255.0f * 255.0f;
    const int a = (int)(s & 0xFFFFu), b = (int)(s >> 16);
    const int p = a * b;
    const int q = a * a + b * b;
    const int vxy = 49 * (int)suv - p;
    const int vs = 49 * (int)st - q;
    const float num = fmaf(2.0f, (float)p, C1) * fmaf(2.0f, (float)vxy, C2);
    const float den = ((float)q + C1) * ((float)vs + C2);
    return __fdividef(num, den);
}

template <int C>
__device__ __forceinline__ uint8_t luma_of(const uint8_t* p) {
    if (C == 1) return p[0];
    return (uint8_t)((p[0] * 19595u + p[1] * 38470u + p[2] * 7471u + 0x8000u) >> 16);
}

template <int C>
__global__ void __launch_bounds__(kThreads) ke_ssim_kernel(const SsimArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    // [buf 0: u strip | v strip][buf 1: u strip | v strip][2 mbarriers]
    const int strip_bytes = (kStripRows * a.pitch + 16 + 127) / 128 * 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * strip_bytes);
    __shared__ double s_red[kThreads / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity[2] = {0u, 0u};
    const int n_strips = (a.h + kStripRows - 1) / kStripRows;
    const long long n_units = a.n_pairs * a.n_cblocks;

    for (long long unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const long long pair = unit / a.n_cblocks;
        const int cb = (int)(unit - pair * a.n_cblocks);
        const int col0 = cb * kBlockCols;                       // first output column of this unit
        const int out_cols = min(kBlockCols, (a.w - 6) - col0);  // valid output columns
        const int in_cols = out_cols + 6;
        const uint8_t* img_u = a.bank + a.ia[pair] * a.img_stride;
        const uint8_t* img_v = a.bank + a.ib[pair] * a.img_stride;

        auto load_strip = [&](int s) {
            const int buf = s & 1;
            const int r0 = s * kStripRows;
            const int rows = min(kStripRows, a.h - r0);
            uint8_t* du = smem + (2 * buf) * strip_bytes;
            uint8_t* dv = smem + (2 * buf + 1) * strip_bytes;
            if (a.use_bulk) {
                if (tid == 0) {
                    const uint32_t bytes = (uint32_t)(rows * a.w);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(&bars[buf], 2u * bytes);
                    bulk_g2s(du, img_u + (long long)r0 * a.row_stride, bytes, &bars[buf]);
                    bulk_g2s(dv, img_v + (long long)r0 * a.row_stride, bytes, &bars[buf]);
                }
            } else {
                int x_done = 0;
                if (C == 3 && a.rgb_words) {
                    // RGB rows whose start is word aligned: Pillow luma of 4 pixels from 3 coalesced words
                    constexpr uint32_t LO = 0x002F468Bu, HI = 0x001D964Cu;  // 19595, 38470, 7471 split in bytes
                    const int groups = in_cols >> 2;
                    x_done = groups << 2;
                    for (int idx = tid; idx < rows * groups; idx += kThreads) {
                        const int r = idx / groups, q = idx - r * groups;
                        const long long g = (long long)(r0 + r) * a.row_stride + (long long)col0 * 3 + (long long)q * 12;
#pragma unroll
                        for (int im = 0; im < 2; ++im) {
                            const uint32_t* src = reinterpret_cast<const uint32_t*>((im ? img_v : img_u) + g);
                            const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
                            const uint32_t l0 = dp4a_uu(w0, LO, 0x8000u), h0 = dp4a_uu(w0, HI, 0u);
                            uint32_t l1 = dp4a_uu(w0, LO << 24, 0x8000u), h1 = dp4a_uu(w0, HI << 24, 0u);
                            l1 = dp4a_uu(w1, LO >> 8, l1), h1 = dp4a_uu(w1, HI >> 8, h1);
                            uint32_t l2 = dp4a_uu(w1, LO << 16, 0x8000u), h2 = dp4a_uu(w1, HI << 16, 0u);
                            l2 = dp4a_uu(w2, LO >> 16, l2), h2 = dp4a_uu(w2, HI >> 16, h2);
                            const uint32_t l3 = dp4a_uu(w2, LO << 8, 0x8000u), h3 = dp4a_uu(w2, HI << 8, 0u);
                            const uint32_t s0 = l0 + (h0 << 8), s1 = l1 + (h1 << 8), s2 = l2 + (h2 << 8), s3 = l3 + (h3 << 8);
                            const uint32_t packed = __byte_perm(__byte_perm(s0, s1, 0x0062), __byte_perm(s2, s3, 0x0062), 0x5410);
                            *reinterpret_cast<uint32_t*>((im ? dv : du) + r * a.pitch + 4 * q) = packed;
                        }
                    }
                }
                const int tail = in_cols - x_done;
                for (int idx = tid; idx < rows * tail; idx += kThreads) {
                    const int r = idx / tail, x = x_done + (idx - r * tail);
                    const long long g = (long long)(r0 + r) * a.row_stride + (long long)(col0 + x) * C;
                    du[r * a.pitch + x] = luma_of<C>(img_u + g);
                    dv[r * a.pitch + x] = luma_of<C>(img_v + g);
                }
                __syncthreads();
                if (tid == 0) mbar_arrive(&bars[buf]);
            }
        };

        HSum ring[kWin];
#pragma unroll
        for (int i = 0; i < kWin; ++i) ring[i] = HSum{0u, 0u, 0u};
        uint32_t acc_s = 0, acc_t = 0, acc_uv = 0;
        double total = 0.0;
        const bool active = tid < out_cols;
        const uint32_t sel = 0x3210u + 0x1111u * (uint32_t)(tid & 3);
        const int word = tid >> 2;

        __syncthreads();  // every thread is done with the previous unit's buffers
        load_strip(0);
        for (int s = 0; s < n_strips; ++s) {
            const int buf = s & 1;
            if (s + 1 < n_strips) load_strip(s + 1);  // buffer (s+1)&1 was released by the barrier below
            mbar_wait(&bars[buf], parity[buf]);
            parity[buf] ^= 1u;
            const int r0 = s * kStripRows;
            const int rows = min(kStripRows, a.h - r0);
            const uint32_t* su = reinterpret_cast<const uint32_t*>(smem + (2 * buf) * strip_bytes) + word;
            const uint32_t* sv = reinterpret_cast<const uint32_t*>(smem + (2 * buf + 1) * strip_bytes) + word;
            const int pw = a.pitch >> 2;
            float part = 0.f;
            if (active) {
                // One 7-row group; STEADY = all 7 rows exist and the window is already full, so the
                // loop body carries no predicates (the common case: every strip but the first/last).
                auto group = [&](int rb, auto steady) {
                    constexpr bool STEADY = decltype(steady)::value;
#pragma unroll
                    for (int k = 0; k < kWin; ++k) {
                        const int r = rb + k;  // (r0 + r) % 7 == k because strips are multiples of 7
                        if (STEADY || r < rows) {
                            // the slot of row (r0+r-7) is dead: build the new row's sums in place
                            ring[k] = hsum7(su + r * pw, sv + r * pw, sel);
                            acc_s += ring[k].s;
                            acc_t += ring[k].t;
                            acc_uv += ring[k].uv;
                            if (STEADY || r0 + r >= kWin - 1) part += ssim_point(acc_s, acc_t, acc_uv);
                            // slot (k+1)%7 holds row (r0+r-6): it leaves the window
                            acc_s -= ring[(k + 1) % kWin].s;
                            acc_t -= ring[(k + 1) % kWin].t;
                            acc_uv -= ring[(k + 1) % kWin].uv;
                        }
                    }
                };
                for (int rb = 0; rb < rows; rb += kWin) {
                    if (r0 + rb >= kWin && rb + kWin <= rows) group(rb, std::true_type{});
                    else group(rb, std::false_type{});
                }
            }
            total += (double)part;
            __syncthreads();  // strip buffer `buf` may be refilled
        }

        // block reduction of the per-column sums
#pragma unroll
        for (int off = 16; off; off >>= 1) total += __shfl_xor_sync(0xffffffffu, total, off);
        if (lane == 0) s_red[warp] = total;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
#pragma unroll
            for (int i = 0; i < kThreads / 32; ++i) t += s_red[i];
            if (a.n_cblocks == 1) a.out[pair] = t * a.inv_count;
            else a.partial[unit] = t;  // summed in column-block order by ke_ssim_finish_kernel: bit-reproducible
        }
    }
}


// ------------------------------------------------------------------------------------------
// v2: four output columns per thread.
//
// The v1 kernel (thread = one output column) is bound by instruction issue: per pixel it loads six
// words, realigns them with four PRMT and evaluates the point formula on scalars.  Here a thread owns the
// outputs x0..x0+3 (x0 = 4 tid): the three words per image and row are loaded ONCE for the four windows
// (bytes x0..x0+9), the window of output 0 is word aligned (no PRMT), and the point formula runs on
// output PAIRS with the packed FP32 instructions of sm_100 (FFMA2 / FADD2 / FMUL2).  A unit is still
// (pair, 256 output columns), now walked by 64 threads; the 7-row ring lives in registers (84 of them).

constexpr int kT4 = 64;
#ifndef KE_SSIM4_CTAS
#define KE_SSIM4_CTAS 6
#endif
#ifndef KE_SSIM4S_CTAS
#define KE_SSIM4S_CTAS 6
#endif

__device__ __forceinline__ HSum hsum7_words(uint32_t ua, uint32_t ub, uint32_t va, uint32_t vb) {
    const uint32_t ubm = ub & 0x00FFFFFFu, vbm = vb & 0x00FFFFFFu;
    HSum r;
    const uint32_t su = dp4a_uu(ua, 0x01010101u, dp4a_uu(ub, 0x00010101u, 0u));
    const uint32_t sv = dp4a_uu(va, 0x01010101u, dp4a_uu(vb, 0x00010101u, 0u));
    r.s = su + (sv << 16);
    r.t = dp4a_uu(ua, ua, dp4a_uu(ub, ubm, dp4a_uu(va, va, dp4a_uu(vb, vbm, 0u))));
    r.uv = dp4a_uu(ua, va, dp4a_uu(ub, vbm, 0u));
    return r;
}

// SSIM of two neighbouring windows from their exact integer sums (same arithmetic as ssim_point)
__device__ __forceinline__ float2 ssim_point2(uint32_t s0, uint32_t t0, uint32_t uv0, uint32_t s1, uint32_t t1, uint32_t uv1) {
    constexpr float C1 = 1e-4f * 49.0f * 49.0f * 255.0f * 255.0f;
    constexpr float C2 = 9e-4f * 48.0f * 49.0f * 255.0f * 255.0f;
    const int a0 = (int)(s0 & 0xFFFFu), b0 = (int)(s0 >> 16), a1 = (int)(s1 & 0xFFFFu), b1 = (int)(s1 >> 16);
    const int p0 = a0 * b0, p1 = a1 * b1;
    const int q0 = a0 * a0 + b0 * b0, q1 = a1 * a1 + b1 * b1;
    const int x0 = 49 * (int)uv0 - p0, x1 = 49 * (int)uv1 - p1;
    const int v0 = 49 * (int)t0 - q0, v1 = 49 * (int)t1 - q1;
    const float2 two = make_float2(2.0f, 2.0f), c1 = make_float2(C1, C1), c2 = make_float2(C2, C2);
    const float2 A1 = __ffma2_rn(two, make_float2((float)p0, (float)p1), c1);
    const float2 A2 = __ffma2_rn(two, make_float2((float)x0, (float)x1), c2);
    const float2 B1 = __fadd2_rn(make_float2((float)q0, (float)q1), c1);
    const float2 B2 = __fadd2_rn(make_float2((float)v0, (float)v1), c2);
    const float2 num = __fmul2_rn(A1, A2), den = __fmul2_rn(B1, B2);  // den >= c1*c2 > 0, far from the rcp range limits
    float rx, ry;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rx) : "f"(den.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ry) : "f"(den.y));
    return __fmul2_rn(num, make_float2(rx, ry));
}

template <int C>
__global__ void __launch_bounds__(kT4, KE_SSIM4_CTAS) ke_ssim4_kernel(const SsimArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int strip_bytes = (kStripRows * a.pitch + 32 + 127) / 128 * 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * strip_bytes);
    __shared__ double s_red[kT4 / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity[2] = {0u, 0u};
    const int n_strips = (a.h + kStripRows - 1) / kStripRows;
    const long long n_units = a.n_pairs * a.n_cblocks;

    for (long long unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const long long pair = unit / a.n_cblocks;
        const int cb = (int)(unit - pair * a.n_cblocks);
        const int col0 = cb * kBlockCols;
        const int out_cols = min(kBlockCols, (a.w - 6) - col0);
        const int in_cols = out_cols + 6;
        const uint8_t* img_u = a.bank + a.ia[pair] * a.img_stride;
        const uint8_t* img_v = a.bank + a.ib[pair] * a.img_stride;

        auto load_strip = [&](int s) {
            const int buf = s & 1;
            const int r0 = s * kStripRows;
            const int rows = min(kStripRows, a.h - r0);
            uint8_t* du = smem + (2 * buf) * strip_bytes;
            uint8_t* dv = smem + (2 * buf + 1) * strip_bytes;
            if (a.use_bulk) {
                if (tid == 0) {
                    const uint32_t bytes = (uint32_t)(rows * a.w);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(&bars[buf], 2u * bytes);
                    bulk_g2s(du, img_u + (long long)r0 * a.row_stride, bytes, &bars[buf]);
                    bulk_g2s(dv, img_v + (long long)r0 * a.row_stride, bytes, &bars[buf]);
                }
            } else {
                int x_done = 0;
                if (C == 3 && a.rgb_words) {
                    constexpr uint32_t LO = 0x002F468Bu, HI = 0x001D964Cu;
                    const int groups = in_cols >> 2;
                    x_done = groups << 2;
                    for (int idx = tid; idx < rows * groups; idx += kT4) {
                        const int r = idx / groups, q = idx - r * groups;
                        const long long g = (long long)(r0 + r) * a.row_stride + (long long)col0 * 3 + (long long)q * 12;
#pragma unroll
                        for (int im = 0; im < 2; ++im) {
                            const uint32_t* src = reinterpret_cast<const uint32_t*>((im ? img_v : img_u) + g);
                            const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
                            const uint32_t l0 = dp4a_uu(w0, LO, 0x8000u), h0 = dp4a_uu(w0, HI, 0u);
                            uint32_t l1 = dp4a_uu(w0, LO << 24, 0x8000u), h1 = dp4a_uu(w0, HI << 24, 0u);
                            l1 = dp4a_uu(w1, LO >> 8, l1), h1 = dp4a_uu(w1, HI >> 8, h1);
                            uint32_t l2 = dp4a_uu(w1, LO << 16, 0x8000u), h2 = dp4a_uu(w1, HI << 16, 0u);
                            l2 = dp4a_uu(w2, LO >> 16, l2), h2 = dp4a_uu(w2, HI >> 16, h2);
                            const uint32_t l3 = dp4a_uu(w2, LO << 8, 0x8000u), h3 = dp4a_uu(w2, HI << 8, 0u);
                            const uint32_t s0 = l0 + (h0 << 8), s1 = l1 + (h1 << 8), s2 = l2 + (h2 << 8), s3 = l3 + (h3 << 8);
                            const uint32_t packed = __byte_perm(__byte_perm(s0, s1, 0x0062), __byte_perm(s2, s3, 0x0062), 0x5410);
                            *reinterpret_cast<uint32_t*>((im ? dv : du) + r * a.pitch + 4 * q) = packed;
                        }
                    }
                }
                const int tail = in_cols - x_done;
                for (int idx = tid; idx < rows * tail; idx += kT4) {
                    const int r = idx / tail, x = x_done + (idx - r * tail);
                    const long long g = (long long)(r0 + r) * a.row_stride + (long long)(col0 + x) * C;
                    du[r * a.pitch + x] = luma_of<C>(img_u + g);
                    dv[r * a.pitch + x] = luma_of<C>(img_v + g);
                }
                __syncthreads();
                if (tid == 0) mbar_arrive(&bars[buf]);
            }
        };

        HSum ring[4][kWin];
        uint32_t acc_s[4], acc_t[4], acc_uv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc_s[j] = acc_t[j] = acc_uv[j] = 0u;
#pragma unroll
            for (int i = 0; i < kWin; ++i) ring[j][i] = HSum{0u, 0u, 0u};
        }
        double total = 0.0;
        const int x0 = 4 * tid;
        const bool active = x0 < out_cols;
        // outputs past the last valid column read padding / the next row's bytes: computed, then masked out
        const float m0 = x0 < out_cols ? 1.f : 0.f, m1 = x0 + 1 < out_cols ? 1.f : 0.f, m2 = x0 + 2 < out_cols ? 1.f : 0.f,
                    m3 = x0 + 3 < out_cols ? 1.f : 0.f;

        __syncthreads();  // every thread is done with the previous unit's buffers
        load_strip(0);
        for (int s = 0; s < n_strips; ++s) {
            const int buf = s & 1;
            if (s + 1 < n_strips) load_strip(s + 1);
            mbar_wait(&bars[buf], parity[buf]);
            parity[buf] ^= 1u;
            const int r0 = s * kStripRows;
            const int rows = min(kStripRows, a.h - r0);
            const uint32_t* su = reinterpret_cast<const uint32_t*>(smem + (2 * buf) * strip_bytes) + tid;
            const uint32_t* sv = reinterpret_cast<const uint32_t*>(smem + (2 * buf + 1) * strip_bytes) + tid;
            const int pw = a.pitch >> 2;
            float part = 0.f;
            if (active) {
                auto group = [&](int rb, auto steady) {
                    constexpr bool STEADY = decltype(steady)::value;
#pragma unroll
                    for (int k = 0; k < kWin; ++k) {
                        const int r = rb + k;  // (r0 + r) % 7 == k because strips are multiples of 7
                        if (STEADY || r < rows) {
                            const uint32_t* pu = su + r * pw;
                            const uint32_t* pv = sv + r * pw;
                            const uint32_t u0 = pu[0], u1 = pu[1], u2 = pu[2];
                            const uint32_t v0 = pv[0], v1 = pv[1], v2 = pv[2];
                            ring[0][k] = hsum7_words(u0, u1, v0, v1);
                            ring[1][k] = hsum7_words(__byte_perm(u0, u1, 0x4321), __byte_perm(u1, u2, 0x4321),
                                                     __byte_perm(v0, v1, 0x4321), __byte_perm(v1, v2, 0x4321));
                            ring[2][k] = hsum7_words(__byte_perm(u0, u1, 0x5432), __byte_perm(u1, u2, 0x5432),
                                                     __byte_perm(v0, v1, 0x5432), __byte_perm(v1, v2, 0x5432));
                            ring[3][k] = hsum7_words(__byte_perm(u0, u1, 0x6543), __byte_perm(u1, u2, 0x6543),
                                                     __byte_perm(v0, v1, 0x6543), __byte_perm(v1, v2, 0x6543));
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                acc_s[j] += ring[j][k].s;
                                acc_t[j] += ring[j][k].t;
                                acc_uv[j] += ring[j][k].uv;
                            }
                            if (STEADY || r0 + r >= kWin - 1) {
                                const float2 e01 = ssim_point2(acc_s[0], acc_t[0], acc_uv[0], acc_s[1], acc_t[1], acc_uv[1]);
                                const float2 e23 = ssim_point2(acc_s[2], acc_t[2], acc_uv[2], acc_s[3], acc_t[3], acc_uv[3]);
                                const float2 sm = __ffma2_rn(e23, make_float2(m2, m3), __fmul2_rn(e01, make_float2(m0, m1)));
                                part += sm.x + sm.y;
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                acc_s[j] -= ring[j][(k + 1) % kWin].s;
                                acc_t[j] -= ring[j][(k + 1) % kWin].t;
                                acc_uv[j] -= ring[j][(k + 1) % kWin].uv;
                            }
                        }
                    }
                };
                for (int rb = 0; rb < rows; rb += kWin) {
                    if (r0 + rb >= kWin && rb + kWin <= rows) group(rb, std::true_type{});
                    else group(rb, std::false_type{});
                }
            }
            total += (double)part;
            __syncthreads();  // strip buffer `buf` may be refilled
        }

#pragma unroll
        for (int off = 16; off; off >>= 1) total += __shfl_xor_sync(0xffffffffu, total, off);
        if (lane == 0) s_red[warp] = total;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
#pragma unroll
            for (int i = 0; i < kT4 / 32; ++i) t += s_red[i];
            if (a.n_cblocks == 1) a.out[pair] = t * a.inv_count;
            else a.partial[unit] = t;  // summed in column-block order by ke_ssim_finish_kernel: bit-reproducible
        }
    }
}

template <int C>
int launch_ssim4(ke_ctx* ctx, SsimArgs& a, cudaStream_t s) {
    const int strip_bytes = (kStripRows * a.pitch + 32 + 127) / 128 * 128;
    const int smem = 4 * strip_bytes + 16;
    KE_CUDA(cudaFuncSetAttribute(ke_ssim4_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int per_sm = 0;
    KE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ke_ssim4_kernel<C>, kT4, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)ctx->sm_count * per_sm;
    const long long units = a.n_pairs * a.n_cblocks;
    if (grid > units) grid = units;
    ke_ssim4_kernel<C><<<(unsigned)grid, kT4, smem, s>>>(a);
    ctx->launches++;
    KE_CUDA(cudaGetLastError());
    return KE_OK;
}


// ------------------------------------------------------------------------------------------
// v2, staged feed (RGB / RGBA banks, 'L' planes wider than one column block): the strips cannot come as one
// contiguous bulk copy, so every thread streams 16-byte pieces of the raw rows into shared memory with cp.async
// (no registers held, the copy of strip s+1 runs under the arithmetic of strip s), and the raw strip is turned into
// the luma strip shared memory to shared memory (dp2a, 16 pixels per step).  Same 4-columns-per-thread arithmetic.

constexpr int kSR2 = 14;  // rows per staged strip (multiple of 7)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ uint32_t dp2a_lo(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t dp2a_hi(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// 4 RGB pixels (12 bytes) -> 4 luma bytes, two dp2a per pixel (see ke_phash.cu luma4_rgb)
__device__ __forceinline__ uint32_t ssim_luma4_rgb(uint32_t w0, uint32_t w1, uint32_t w2) {
    constexpr uint32_t cR = 19595u, cG = 38470u, cB = 7471u;
    constexpr uint32_t RG = cR | (cG << 16), B_ = cB, _R = cR << 16, GB = cG | (cB << 16);
    const uint32_t s0 = dp2a_hi(B_, w0, dp2a_lo(RG, w0, 0x8000u));
    const uint32_t s1 = dp2a_lo(GB, w1, dp2a_hi(_R, w0, 0x8000u));
    const uint32_t s2 = dp2a_lo(B_, w2, dp2a_hi(RG, w1, 0x8000u));
    const uint32_t s3 = dp2a_hi(GB, w2, dp2a_lo(_R, w2, 0x8000u));
    return __byte_perm(__byte_perm(s0, s1, 0x0062), __byte_perm(s2, s3, 0x0062), 0x5410);
}
__device__ __forceinline__ uint32_t ssim_luma4_rgba(uint4 px) {
    constexpr uint32_t RG = 19595u | (38470u << 16), B_ = 7471u;
    const uint32_t s0 = dp2a_hi(B_, px.x, dp2a_lo(RG, px.x, 0x8000u)), s1 = dp2a_hi(B_, px.y, dp2a_lo(RG, px.y, 0x8000u));
    const uint32_t s2 = dp2a_hi(B_, px.z, dp2a_lo(RG, px.z, 0x8000u)), s3 = dp2a_hi(B_, px.w, dp2a_lo(RG, px.w, 0x8000u));
    return __byte_perm(__byte_perm(s0, s1, 0x0062), __byte_perm(s2, s3, 0x0062), 0x5410);
}

struct Ssim4State {  // the register state a thread carries down the rows of one unit
    HSum ring[4][kWin];
    uint32_t acc_s[4], acc_t[4], acc_uv[4];
};

// rows [0, rows) of a luma strip pair (word pointers already offset by the thread's first word) -> partial sum
__device__ __forceinline__ float ssim4_strip(Ssim4State& st, const uint32_t* su, const uint32_t* sv, int pw, int rows, int r0,
                                             float m0, float m1, float m2, float m3) {
    float part = 0.f;
    auto group = [&](int rb, auto steady) {
        constexpr bool STEADY = decltype(steady)::value;
#pragma unroll
        for (int k = 0; k < kWin; ++k) {
            const int r = rb + k;  // (r0 + r) % 7 == k because strips are multiples of 7
            if (STEADY || r < rows) {
                const uint32_t* pu = su + r * pw;
                const uint32_t* pv = sv + r * pw;
                const uint32_t u0 = pu[0], u1 = pu[1], u2 = pu[2];
                const uint32_t v0 = pv[0], v1 = pv[1], v2 = pv[2];
                st.ring[0][k] = hsum7_words(u0, u1, v0, v1);
                st.ring[1][k] = hsum7_words(__byte_perm(u0, u1, 0x4321), __byte_perm(u1, u2, 0x4321),
                                            __byte_perm(v0, v1, 0x4321), __byte_perm(v1, v2, 0x4321));
                st.ring[2][k] = hsum7_words(__byte_perm(u0, u1, 0x5432), __byte_perm(u1, u2, 0x5432),
                                            __byte_perm(v0, v1, 0x5432), __byte_perm(v1, v2, 0x5432));
                st.ring[3][k] = hsum7_words(__byte_perm(u0, u1, 0x6543), __byte_perm(u1, u2, 0x6543),
                                            __byte_perm(v0, v1, 0x6543), __byte_perm(v1, v2, 0x6543));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    st.acc_s[j] += st.ring[j][k].s;
                    st.acc_t[j] += st.ring[j][k].t;
                    st.acc_uv[j] += st.ring[j][k].uv;
                }
                if (STEADY || r0 + r >= kWin - 1) {
                    const float2 e01 = ssim_point2(st.acc_s[0], st.acc_t[0], st.acc_uv[0], st.acc_s[1], st.acc_t[1], st.acc_uv[1]);
                    const float2 e23 = ssim_point2(st.acc_s[2], st.acc_t[2], st.acc_uv[2], st.acc_s[3], st.acc_t[3], st.acc_uv[3]);
                    const float2 sm = __ffma2_rn(e23, make_float2(m2, m3), __fmul2_rn(e01, make_float2(m0, m1)));
                    part += sm.x + sm.y;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    st.acc_s[j] -= st.ring[j][(k + 1) % kWin].s;
                    st.acc_t[j] -= st.ring[j][(k + 1) % kWin].t;
                    st.acc_uv[j] -= st.ring[j][(k + 1) % kWin].uv;
                }
            }
        }
    };
    for (int rb = 0; rb < rows; rb += kWin) {
        if (r0 + rb >= kWin && rb + kWin <= rows) group(rb, std::true_type{});
        else group(rb, std::false_type{});
    }
    return part;
}

template <int C>
__global__ void __launch_bounds__(kT4, KE_SSIM4S_CTAS) ke_ssim4s_kernel(const SsimArgs a, const int raw_pitch) {
    extern __shared__ __align__(128) uint8_t smem[];
    // C > 1: [raw u | raw v] (one strip, refilled under the arithmetic) + [luma u | luma v];  C == 1: 2 x [luma u | luma v]
    const int luma_bytes = (kSR2 * a.pitch + 32 + 127) / 128 * 128;
    const int raw_bytes = C == 1 ? 0 : (kSR2 * raw_pitch + 64 + 127) / 128 * 128;
    uint8_t* s_rawbuf = smem;
    uint8_t* s_lum = smem + 2 * raw_bytes;
    __shared__ double s_red[kT4 / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_strips = (a.h + kSR2 - 1) / kSR2;
    const long long n_units = a.n_pairs * a.n_cblocks;
    const int row_bytes_img = a.w * C;

    for (long long unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const long long pair = unit / a.n_cblocks;
        const int cb = (int)(unit - pair * a.n_cblocks);
        const int col0 = cb * kBlockCols;
        const int out_cols = min(kBlockCols, (a.w - 6) - col0);
        const int in_cols = out_cols + 6;
        const uint8_t* img_u = a.bank + a.ia[pair] * a.img_stride + (long long)col0 * C;
        const uint8_t* img_v = a.bank + a.ib[pair] * a.img_stride + (long long)col0 * C;
        // 16-byte pieces per row: everything the block needs, never past the end of the image row
        const int piece_bytes = min((in_cols * C + 15) / 16 * 16, row_bytes_img - col0 * C);
        const int pieces = piece_bytes >> 4;

        auto fetch = [&](int s, int buf) {  // raw rows of strip s -> shared memory, asynchronously
            const int r0 = s * kSR2;
            const int rows = min(kSR2, a.h - r0);
            uint8_t* du = C == 1 ? s_lum + (2 * buf) * luma_bytes : s_rawbuf;
            uint8_t* dv = C == 1 ? s_lum + (2 * buf + 1) * luma_bytes : s_rawbuf + raw_bytes;
            const int pitch = C == 1 ? a.pitch : raw_pitch;
            for (int idx = tid; idx < rows * pieces; idx += kT4) {
                const int r = idx / pieces, q = idx - r * pieces;
                const long long g = (long long)(r0 + r) * a.row_stride + 16 * q;
                cp_async16(du + r * pitch + 16 * q, img_u + g);
                cp_async16(dv + r * pitch + 16 * q, img_v + g);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto convert = [&](int s) {  // raw strip -> luma strip (C > 1)
            const int rows = min(kSR2, a.h - s * kSR2);
            if (C == 3) {
                const int groups = (in_cols + 15) >> 4;  // 16 pixels = 48 raw bytes per step
                for (int idx = tid; idx < 2 * rows * groups; idx += kT4) {
                    const int im = idx / (rows * groups), rem = idx - im * rows * groups;
                    const int r = rem / groups, q = rem - r * groups;
                    const uint4* src = reinterpret_cast<const uint4*>(s_rawbuf + im * raw_bytes + r * raw_pitch + 48 * q);
                    const uint4 x = src[0], y = src[1], z = src[2];
                    *reinterpret_cast<uint4*>(s_lum + im * luma_bytes + r * a.pitch + 16 * q) =
                        make_uint4(ssim_luma4_rgb(x.x, x.y, x.z), ssim_luma4_rgb(x.w, y.x, y.y), ssim_luma4_rgb(y.z, y.w, z.x),
                                   ssim_luma4_rgb(z.y, z.z, z.w));
                }
            } else if (C == 4) {
                const int groups = (in_cols + 3) >> 2;
                for (int idx = tid; idx < 2 * rows * groups; idx += kT4) {
                    const int im = idx / (rows * groups), rem = idx - im * rows * groups;
                    const int r = rem / groups, q = rem - r * groups;
                    const uint4 px = *reinterpret_cast<const uint4*>(s_rawbuf + im * raw_bytes + r * raw_pitch + 16 * q);
                    *reinterpret_cast<uint32_t*>(s_lum + im * luma_bytes + r * a.pitch + 4 * q) = ssim_luma4_rgba(px);
                }
            }
        };

        Ssim4State st;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            st.acc_s[j] = st.acc_t[j] = st.acc_uv[j] = 0u;
#pragma unroll
            for (int i = 0; i < kWin; ++i) st.ring[j][i] = HSum{0u, 0u, 0u};
        }
        double total = 0.0;
        const int x0 = 4 * tid;
        const bool active = x0 < out_cols;
        const float m0 = x0 < out_cols ? 1.f : 0.f, m1 = x0 + 1 < out_cols ? 1.f : 0.f, m2 = x0 + 2 < out_cols ? 1.f : 0.f,
                    m3 = x0 + 3 < out_cols ? 1.f : 0.f;
        const int pw = a.pitch >> 2;

        __syncthreads();  // every thread is done with the previous unit's buffers
        fetch(0, 0);
        for (int s = 0; s < n_strips; ++s) {
            const int buf = C == 1 ? (s & 1) : 0;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();  // strip s has landed for every thread
            if (C > 1) {
                convert(s);
                __syncthreads();  // luma strip complete, raw strip free
            }
            if (s + 1 < n_strips) fetch(s + 1, (s + 1) & 1);  // runs under the arithmetic below
            const int r0 = s * kSR2;
            const int rows = min(kSR2, a.h - r0);
            const uint32_t* su = reinterpret_cast<const uint32_t*>(s_lum + (2 * buf) * luma_bytes) + tid;
            const uint32_t* sv = reinterpret_cast<const uint32_t*>(s_lum + (C == 1 ? (2 * buf + 1) : 1) * luma_bytes) + tid;
            if (active) total += (double)ssim4_strip(st, su, sv, pw, rows, r0, m0, m1, m2, m3);
            if (C > 1) __syncthreads();  // the luma strip is rewritten by the next convert
        }

#pragma unroll
        for (int off = 16; off; off >>= 1) total += __shfl_xor_sync(0xffffffffu, total, off);
        if (lane == 0) s_red[warp] = total;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
#pragma unroll
            for (int i = 0; i < kT4 / 32; ++i) t += s_red[i];
            if (a.n_cblocks == 1) a.out[pair] = t * a.inv_count;
            else a.partial[unit] = t;  // summed in column-block order by ke_ssim_finish_kernel: bit-reproducible
        }
    }
}

template <int C>
int launch_ssim4s(ke_ctx* ctx, SsimArgs& a, cudaStream_t s) {
    const int in_max = std::min(a.w, kBlockCols + 6);
    const int raw_pitch = (in_max * C + 15) / 16 * 16 + 16;
    const int luma_bytes = (kSR2 * a.pitch + 32 + 127) / 128 * 128;
    const int raw_bytes = C == 1 ? 0 : (kSR2 * raw_pitch + 64 + 127) / 128 * 128;
    const int smem = 2 * raw_bytes + (C == 1 ? 4 : 2) * luma_bytes;
    KE_CUDA(cudaFuncSetAttribute(ke_ssim4s_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int per_sm = 0;
    KE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ke_ssim4s_kernel<C>, kT4, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)ctx->sm_count * per_sm;
    const long long units = a.n_pairs * a.n_cblocks;
    if (grid > units) grid = units;
    ke_ssim4s_kernel<C><<<(unsigned)grid, kT4, smem, s>>>(a, raw_pitch);
    ctx->launches++;
    KE_CUDA(cudaGetLastError());
    return KE_OK;
}

template <int C>
int launch_ssim(ke_ctx* ctx, SsimArgs& a, cudaStream_t s) {
    const int strip_bytes = (kStripRows * a.pitch + 16 + 127) / 128 * 128;
    const int smem = 4 * strip_bytes + 16;
    KE_CUDA(cudaFuncSetAttribute(ke_ssim_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int per_sm = 0;
    KE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ke_ssim_kernel<C>, kThreads, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)ctx->sm_count * per_sm;
    const long long units = a.n_pairs * a.n_cblocks;
    if (grid > units) grid = units;
    ke_ssim_kernel<C><<<(unsigned)grid, kThreads, smem, s>>>(a);
    ctx->launches++;
    KE_CUDA(cudaGetLastError());
    return KE_OK;
}

// ------------------------------------------------------------------------------------------
// Gaussian-window variant (`gaussian != 0`): skimage's structural_similarity(..., gaussian_weights=True) — sigma 1.5,
// truncate 3.5 -> 11 taps, win_size 11 (crop 5, cov_norm 121/120).  NOT the reference's path (src/dup/refine.py:52 uses
// the uniform window); offered because BASELINE.json's north_star words kernel 3 as "a separable Gaussian window".
// It mirrors scipy.ndimage.gaussian_filter on float32 images operation by operation: axis 0 then axis 1, every 1-D pass
// accumulated in FP64 in NI_Correlate1D's symmetric order (centre tap, then (x[-j] + x[+j]) * w[j] from the outermost
// pair inwards, no FMA contraction) and rounded to float32 in between; the point formula in float32, the mean in FP64.
// So constant and identical images give exactly the values the float32 reference path gives.

constexpr int kGaussWin = 11, kGaussR = 5;
constexpr int kGTW = 32, kGTH = 16, kGThreads = 128;
constexpr int kGIW = kGTW + 2 * kGaussR, kGIH = kGTH + 2 * kGaussR;  // 42 x 26 input pixels per tile

__constant__ double c_gauss[kGaussR + 1];  // w[0] = outermost tap ... w[5] = centre

struct GaussArgs {
    const uint8_t* bank;
    int h, w, c;
    long long img_stride, row_stride;
    const long long* ia;
    const long long* ib;
    long long n_pairs;
    int tiles_x, tiles_y;
    double* partial;  // [n_pairs][tiles_y * tiles_x]
};

__device__ __forceinline__ float gauss_line(const float* x, int stride) {  // 11 taps centred on x[0]
    double t = __dmul_rn((double)x[0], c_gauss[kGaussR]);
#pragma unroll
    for (int j = kGaussR; j >= 1; --j)
        t = __dadd_rn(t, __dmul_rn(__dadd_rn((double)x[-j * stride], (double)x[j * stride]), c_gauss[kGaussR - j]));
    return (float)t;
}

template <int C>
__global__ void __launch_bounds__(kGThreads) ke_ssim_gauss_kernel(const GaussArgs a) {
    __shared__ float s_in[2][kGIH][kGIW];        // u, v as float32 pixel / 255
    __shared__ float s_mid[5][kGTH][kGIW + 1];   // axis-0 pass of u, v, uu, vv, uv (float32, like scipy's intermediate)
    __shared__ double s_red[kGThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles = a.tiles_x * a.tiles_y;
    const long long n_units = a.n_pairs * tiles;
    const float cov_norm = (float)(121.0 / 120.0), C1 = (float)(0.01 * 0.01), C2 = (float)(0.03 * 0.03);
    for (long long unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const long long pair = unit / tiles;
        const int tl = (int)(unit - pair * tiles), ty = tl / a.tiles_x, tx = tl - ty * a.tiles_x;
        const int oy0 = kGaussR + ty * kGTH, ox0 = kGaussR + tx * kGTW;  // first output pixel of the tile (cropped map)
        const int out_rows = min(kGTH, (a.h - kGaussR) - oy0), out_cols = min(kGTW, (a.w - kGaussR) - ox0);
        const uint8_t* img[2] = {a.bank + a.ia[pair] * a.img_stride, a.bank + a.ib[pair] * a.img_stride};
        __syncthreads();  // previous unit's readers are done
        for (int idx = tid; idx < 2 * kGIH * kGIW; idx += kGThreads) {
            const int im = idx / (kGIH * kGIW), rem = idx - im * (kGIH * kGIW), r = rem / kGIW, x = rem - r * kGIW;
            const int gy = oy0 - kGaussR + r, gx = ox0 - kGaussR + x;
            float v = 0.f;
            if (gy < a.h && gx < a.w) v = __fdiv_rn((float)luma_of<C>(img[im] + (long long)gy * a.row_stride + (long long)gx * C), 255.0f);
            s_in[im][r][x] = v;
        }
        __syncthreads();
        for (int idx = tid; idx < kGTH * kGIW; idx += kGThreads) {  // axis 0 (rows) first, as scipy does
            const int r = idx / kGIW, x = idx - r * kGIW;
            float col[5][kGaussWin];
#pragma unroll
            for (int j = 0; j < kGaussWin; ++j) {
                const float u = s_in[0][r + j][x], v = s_in[1][r + j][x];
                col[0][j] = u, col[1][j] = v, col[2][j] = __fmul_rn(u, u), col[3][j] = __fmul_rn(v, v), col[4][j] = __fmul_rn(u, v);
            }
#pragma unroll
            for (int q = 0; q < 5; ++q) s_mid[q][r][x] = gauss_line(&col[q][kGaussR], 1);
        }
        __syncthreads();
        double total = 0.0;
        for (int idx = tid; idx < kGTH * kGTW; idx += kGThreads) {
            const int r = idx / kGTW, x = idx - r * kGTW;
            if (r >= out_rows || x >= out_cols) continue;
            const float ux = gauss_line(&s_mid[0][r][x + kGaussR], 1), uy = gauss_line(&s_mid[1][r][x + kGaussR], 1);
            const float uxx = gauss_line(&s_mid[2][r][x + kGaussR], 1), uyy = gauss_line(&s_mid[3][r][x + kGaussR], 1);
            const float uxy = gauss_line(&s_mid[4][r][x + kGaussR], 1);
            const float vx = __fmul_rn(cov_norm, __fsub_rn(uxx, __fmul_rn(ux, ux)));
            const float vy = __fmul_rn(cov_norm, __fsub_rn(uyy, __fmul_rn(uy, uy)));
            const float vxy = __fmul_rn(cov_norm, __fsub_rn(uxy, __fmul_rn(ux, uy)));
            const float A1 = __fadd_rn(__fmul_rn(__fmul_rn(2.0f, ux), uy), C1), A2 = __fadd_rn(__fmul_rn(2.0f, vxy), C2);
            const float B1 = __fadd_rn(__fadd_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy)), C1), B2 = __fadd_rn(__fadd_rn(vx, vy), C2);
            total += (double)__fdiv_rn(__fmul_rn(A1, A2), __fmul_rn(B1, B2));
        }
#pragma unroll
        for (int off = 16; off; off >>= 1) total += __shfl_xor_sync(0xffffffffu, total, off);
        if (lane == 0) s_red[warp] = total;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
#pragma unroll
            for (int i = 0; i < kGThreads / 32; ++i) t += s_red[i];
            a.partial[unit] = t;
        }
    }
}

__global__ void __launch_bounds__(256) ke_ssim_finish_kernel(const double* __restrict__ partial, long long n_pairs, int n_cblocks,
                                                             double inv_count, double* __restrict__ out) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    double t = 0.0;
    for (int cb = 0; cb < n_cblocks; ++cb) t += partial[p * n_cblocks + cb];
    out[p] = t * inv_count;
}

int g_gauss_uploaded_device = -1;

int launch_ssim_gauss(ke_ctx* ctx, const uint8_t* d_bank, int h, int w, int c, int64_t img_stride, int64_t row_stride,
                      const int64_t* d_ia, const int64_t* d_ib, int64_t n_pairs, double* d_ssim, cudaStream_t s) {
    if (g_gauss_uploaded_device != ctx->device) {
        // scipy.ndimage._filters._gaussian_kernel1d(sigma=1.5, order=0, radius=5): exp(-0.5 / sigma^2 * x^2) / sum, in double
        double phi[kGaussWin], sum = 0.0;
        for (int x = -kGaussR; x <= kGaussR; ++x) sum += (phi[x + kGaussR] = std::exp(-0.5 / (1.5 * 1.5) * (double)(x * x)));
        double wts[kGaussR + 1];
        for (int j = 0; j <= kGaussR; ++j) wts[j] = phi[j] / sum;
        KE_CUDA(cudaMemcpyToSymbol(c_gauss, wts, sizeof(wts)));
        g_gauss_uploaded_device = ctx->device;
    }
    GaussArgs a;
    a.bank = d_bank, a.h = h, a.w = w, a.c = c, a.img_stride = img_stride, a.row_stride = row_stride;
    a.ia = (const long long*)d_ia, a.ib = (const long long*)d_ib, a.n_pairs = n_pairs;
    a.tiles_x = (w - 2 * kGaussR + kGTW - 1) / kGTW, a.tiles_y = (h - 2 * kGaussR + kGTH - 1) / kGTH;
    const int tiles = a.tiles_x * a.tiles_y;
    KE_CUDA(cudaMallocAsync((void**)&a.partial, (size_t)n_pairs * tiles * sizeof(double), s));
    const long long units = n_pairs * tiles;
    const unsigned grid = (unsigned)std::min<long long>(units, (long long)ctx->sm_count * 8);
    switch (c) {
        case 1: ke_ssim_gauss_kernel<1><<<grid, kGThreads, 0, s>>>(a); break;
        case 3: ke_ssim_gauss_kernel<3><<<grid, kGThreads, 0, s>>>(a); break;
        default: ke_ssim_gauss_kernel<4><<<grid, kGThreads, 0, s>>>(a); break;
    }
    ke_ssim_finish_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, s>>>(
        a.partial, n_pairs, tiles, 1.0 / ((double)(h - 2 * kGaussR) * (double)(w - 2 * kGaussR)), d_ssim);
    ctx->launches += 2;
    cudaFreeAsync(a.partial, s);
    KE_CUDA(cudaGetLastError());
    return KE_OK;
}

}  // namespace

extern "C" int ke_ssim_batch(ke_ctx* ctx, const uint8_t* d_bank, int h, int w, int c, int64_t img_stride,
                             int64_t row_stride, const int64_t* d_ia, const int64_t* d_ib, int64_t n_pairs, int gaussian,
                             double* d_ssim, void* stream) {
    KE_REQUIRE(ctx != nullptr, "ke_ssim_batch: ctx is NULL");
    KE_REQUIRE(n_pairs >= 0, "ke_ssim_batch: n_pairs < 0");
    if (n_pairs == 0) return KE_OK;
    KE_REQUIRE(d_bank && d_ia && d_ib && d_ssim, "ke_ssim_batch: NULL buffer");
    KE_REQUIRE(c == 1 || c == 3 || c == 4, "ke_ssim_batch: channels must be 1, 3 or 4 (got %d)", c);
    {
        const int win = gaussian ? kGaussWin : kWin;
        if (h < win || w < win) {
            ke_set_error("win_size exceeds image extent (%dx%d < %d)", w, h, win);
            return KE_E_UNSUPPORTED;
        }
    }
    KE_REQUIRE(row_stride >= (int64_t)w * c && img_stride >= (int64_t)(h - 1) * row_stride + (int64_t)w * c,
               "ke_ssim_batch: strides smaller than the image");
    KeDeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (gaussian) return launch_ssim_gauss(ctx, d_bank, h, w, c, img_stride, row_stride, d_ia, d_ib, n_pairs, d_ssim, s);
    SsimArgs a;
    a.bank = d_bank;
    a.h = h;
    a.w = w;
    a.c = c;
    a.img_stride = img_stride;
    a.row_stride = row_stride;
    a.ia = (const long long*)d_ia;
    a.ib = (const long long*)d_ib;
    a.n_pairs = n_pairs;
    a.n_cblocks = ((w - 6) + kBlockCols - 1) / kBlockCols;
    a.use_bulk = (c == 1 && a.n_cblocks == 1 && (w % 16) == 0 && row_stride == w && (img_stride % 16) == 0 &&
                  (reinterpret_cast<uintptr_t>(d_bank) & 15) == 0);
    a.rgb_words = (c == 3 && (row_stride % 4) == 0 && (img_stride % 4) == 0 && (reinterpret_cast<uintptr_t>(d_bank) & 3) == 0);
    a.pitch = a.use_bulk ? w : ((std::min(w, kBlockCols + 6) + 3) / 4 * 4 + 4);
    a.inv_count = 1.0 / ((double)(h - 6) * (double)(w - 6));
    a.out = d_ssim;
    a.partial = nullptr;
    if (a.n_cblocks > 1)  // stream-ordered: belongs to this call, freed after the finish kernel
        KE_CUDA(cudaMallocAsync((void**)&a.partial, (size_t)n_pairs * a.n_cblocks * sizeof(double), s));
    // v2 (four output columns per thread): strips by one bulk copy per image when the plane is a contiguous 'L' block,
    // else by cp.async pieces (needs 16-byte aligned rows); v1 (one column per thread, plain loads) takes the rest.
    const bool can_stage = (row_stride % 16) == 0 && (img_stride % 16) == 0 && ((int64_t)w * c) % 16 == 0 &&
                           (reinterpret_cast<uintptr_t>(d_bank) & 15) == 0;
    bool use_v1 = !(a.use_bulk || can_stage);
#ifdef KE_TUNING_PROBES
    if (const char* which = getenv("KE_SSIM_KERNEL")) use_v1 = use_v1 || !strcmp(which, "v1");  // "v1" | "v2"
#endif
    if (ctx->force_ssim_v1) use_v1 = true;
    int rc;
    if (use_v1) {
        rc = c == 1 ? launch_ssim<1>(ctx, a, s) : c == 3 ? launch_ssim<3>(ctx, a, s) : launch_ssim<4>(ctx, a, s);
    } else if (a.use_bulk) {
        rc = launch_ssim4<1>(ctx, a, s);
    } else {
        // staged v2 reads up to 11 bytes past a thread's first column and converts 16 pixels per step
        a.pitch = (std::min(w, kBlockCols + 6) + 15) / 16 * 16 + 16;
        rc = c == 1 ? launch_ssim4s<1>(ctx, a, s) : c == 3 ? launch_ssim4s<3>(ctx, a, s) : launch_ssim4s<4>(ctx, a, s);
    }
    if (a.partial) {
        if (rc == KE_OK) {
            ke_ssim_finish_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, s>>>(a.partial, n_pairs, a.n_cblocks, a.inv_count,
                                                                                   d_ssim);
            ctx->launches++;
        }
        cudaFreeAsync(a.partial, s);
        if (rc == KE_OK) KE_CUDA(cudaGetLastError());
    }
    return rc;
}

// Single-device body of ke_ssim_pairs_host (ke_multi.cu fans it over the devices of a context).
int ke_ssim_pairs_host_one(ke_ctx* ctx, const uint8_t* h_a, const uint8_t* h_b, int64_t n_pairs, int h, int w, int c,
                           int gaussian, double* h_ssim) {
    if (n_pairs == 0) return KE_OK;
    KE_REQUIRE(c == 1 || c == 3 || c == 4, "ke_ssim_pairs_host: channels must be 1, 3 or 4 (got %d)", c);
    const int win = gaussian ? kGaussWin : kWin;
    if (h < win || w < win) {
        ke_set_error("win_size exceeds image extent (%dx%d < %d)", w, h, win);
        return KE_E_UNSUPPORTED;
    }
    KeDeviceGuard guard(ctx->device);
    const int64_t plane = (int64_t)h * w * c;
    const int64_t stride = (plane + 15) / 16 * 16;
    int64_t per_chunk = (128ll << 20) / (2 * stride);
    if (per_chunk < 1) per_chunk = 1;
    if (per_chunk > n_pairs) per_chunk = n_pairs;
    // device bank per chunk: [a planes | b planes], pair p -> (p, per_chunk + p)
    void *d_bank[2], *d_idx = nullptr, *d_out = nullptr;
    int rc;
    for (int b = 0; b < 2; ++b)
        if ((rc = ke_ctx_scratch(ctx, b, (size_t)(2 * per_chunk * stride) + 64, &d_bank[b]))) return rc;
    if ((rc = ke_ctx_scratch(ctx, 2, (size_t)per_chunk * 16, &d_idx))) return rc;
    if ((rc = ke_ctx_scratch(ctx, 3, (size_t)n_pairs * 8, &d_out))) return rc;
    std::vector<long long> idx((size_t)per_chunk * 2);
    for (int64_t p = 0; p < per_chunk; ++p) {
        idx[(size_t)p] = p;
        idx[(size_t)(per_chunk + p)] = per_chunk + p;
    }
    KE_CUDA(cudaMemcpy(d_idx, idx.data(), idx.size() * 8, cudaMemcpyHostToDevice));
    int k = 0;
    for (int64_t p0 = 0; p0 < n_pairs; p0 += per_chunk, ++k) {
        const int b = k & 1;
        const int64_t cnt = std::min<int64_t>(per_chunk, n_pairs - p0);
        cudaStream_t s = ctx->copy_stream[b];
        uint8_t* da = (uint8_t*)d_bank[b];
        uint8_t* db = da + per_chunk * stride;
        if ((rc = ke_h2d_staged_2d(ctx, da, (size_t)stride, h_a + p0 * plane, (size_t)plane, (size_t)cnt, s))) return rc;
        if ((rc = ke_h2d_staged_2d(ctx, db, (size_t)stride, h_b + p0 * plane, (size_t)plane, (size_t)cnt, s))) return rc;
        rc = ke_ssim_batch(ctx, da, h, w, c, stride, (int64_t)w * c, (const int64_t*)d_idx, (const int64_t*)d_idx + per_chunk, cnt,
                           gaussian, (double*)d_out + p0, s);
        if (rc) return rc;
    }
    for (auto s : ctx->copy_stream) KE_CUDA(cudaStreamSynchronize(s));
    KE_CUDA(cudaMemcpy(h_ssim, d_out, (size_t)n_pairs * 8, cudaMemcpyDeviceToHost));
    return KE_OK;
}
