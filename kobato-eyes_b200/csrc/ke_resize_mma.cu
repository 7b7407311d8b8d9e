// ke_resize_mma.cu — N1 fast path: convert("L").resize((out_w, out_h), BILINEAR | LANCZOS) of a batch of decoded
// images as a streaming kernel, byte-identical to Pillow (reference src/ui/dup_refine_parallel.py:66-69, :203-207).
//
// Same machinery as K1 v5 (ke_phash.cu), generalised in the output geometry:
//   luma warps (4)  raw rows -> 1-D TMA bulk copies into two 16-row slots -> Pillow luma (2 dp2a per pixel, 16 pixels
//                   per lane) -> ring of two 32-row luma chunks
//   tap warps (8)   warp q owns output columns 8q..8q+7 END TO END: the horizontal resample as an exact integer matrix
//                   product on the tensor pipe (mma.m16n8k32 u8 x s8 over balanced base-256 tap digits, B fragments
//                   from shared memory), clip to bytes, the eight columns go transposed into a private scratch and
//                   straight back as the B fragment of the vertical resample (mma.m16n8k32 s8 x u8, A = vertical tap
//                   digits).  Output rows come in units of 16; a unit's support spans a few 32-row chunks and at most
//                   two units (of different parity) are live at a time, so the vertical accumulators are two register
//                   slots that are flushed to the output plane when their unit's last chunk has passed.
// One launch covers up to 128 output columns (tap warp q owns the 8-column groups q and q + 8) and 128 output rows.  Shapes it does not take (w % 16, w > 512, out sizes not multiples of 8 / 16, nearly 1:1 scales) stay
// on the generic kernels of ke_refine.cu.
//
// Algorithmic HBM bytes per image: h*w*c read + out_w*out_h written.
#include <algorithm>
#include <cmath>
#include <map>
#include <tuple>
#include <vector>

#include "ke_common.cuh"

int ke_pillow_table(int in_size, int out_size, int filter, std::vector<int32_t>& kk, std::vector<int32_t>& bd, int& ksize);

namespace {

constexpr int kTap = 8, kLuma = 4, kThreads = (kTap + kLuma) * 32;
constexpr int kTapRegs = 88, kLumaRegs = 64;  // setmaxnreg per 4-warp group; the CTA launches with 80 per thread
static_assert(2 * kTapRegs + kLumaRegs == 3 * 80, "register pool of the 3 warp groups");
constexpr int kPrec = 22, kCR = 32, kHP = 48, kSlots = 2, kMaxUnits = 8, kMaxGroups = 16;  // groups per launch: 8 tap warps x <= 2

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
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, uint32_t ns) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(ns);
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
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
// Pillow rgb2l, two dp2a per pixel (see ke_phash.cu: every pixel of a 12-byte group splits at byte-pair boundaries)
__device__ __forceinline__ uint32_t luma4_rgb(uint32_t w0, uint32_t w1, uint32_t w2) {
    constexpr uint32_t cR = 19595u, cG = 38470u, cB = 7471u;
    constexpr uint32_t RG = cR | (cG << 16), B_ = cB, _R = cR << 16, GB = cG | (cB << 16);
    const uint32_t s0 = dp2a_hi(B_, w0, dp2a_lo(RG, w0, 0x8000u));
    const uint32_t s1 = dp2a_lo(GB, w1, dp2a_hi(_R, w0, 0x8000u));
    const uint32_t s2 = dp2a_lo(B_, w2, dp2a_hi(RG, w1, 0x8000u));
    const uint32_t s3 = dp2a_hi(GB, w2, dp2a_lo(_R, w2, 0x8000u));
    return __byte_perm(__byte_perm(s0, s1, 0x0062), __byte_perm(s2, s3, 0x0062), 0x5410);
}
__device__ __forceinline__ uint4 luma16_rgb(const uint4 a, const uint4 b, const uint4 c) {
    return make_uint4(luma4_rgb(a.x, a.y, a.z), luma4_rgb(a.w, b.x, b.y), luma4_rgb(b.z, b.w, c.x), luma4_rgb(c.y, c.z, c.w));
}
__device__ __forceinline__ uint32_t luma4_rgba(const uint4 px) {
    constexpr uint32_t RG = 19595u | (38470u << 16), B_ = 7471u;
    const uint32_t s0 = dp2a_hi(B_, px.x, dp2a_lo(RG, px.x, 0x8000u)), s1 = dp2a_hi(B_, px.y, dp2a_lo(RG, px.y, 0x8000u));
    const uint32_t s2 = dp2a_hi(B_, px.z, dp2a_lo(RG, px.z, 0x8000u)), s3 = dp2a_hi(B_, px.w, dp2a_lo(RG, px.w, 0x8000u));
    return __byte_perm(__byte_perm(s0, s1, 0x0062), __byte_perm(s2, s3, 0x0062), 0x5410);
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void mma_u8s8(int32_t (&c)[4], const uint32_t (&a)[4], const uint2 b) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
__device__ __forceinline__ void mma_s8u8(int32_t (&c)[4], const uint4 a, const uint32_t b0, const uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_sat_u8(int32_t hi, int32_t lo) {  // sat_u8(hi) << 8 | sat_u8(lo)
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(hi), "r"(lo), "r"(0));
    return d;
}

struct ResizeArgs {
    const uint8_t* img;
    long long n, img_stride;
    int h, w;
    int sub_rows, pitch_bytes;
    int n_groups;                 // 8-column output groups in this launch (<= 16), warp q owns groups q and q + 8
    int col_begin;                // first output column of this launch
    int out_w, out_h, n_units;    // full output plane geometry; n_units = out_h / 16
    const uint2* hb;              // horizontal B fragments [group][k][digit][lane]
    int hb_words;
    int h_k0[kMaxGroups], h_nk[kMaxGroups], h_off[kMaxGroups];
    const uint4* va;              // vertical A fragments [chunk][unit][digit][lane]
    int v_lo[kMaxUnits], v_hi[kMaxUnits];
    uint8_t* out;                 // [n][out_h][out_w]
};

struct Layout {
    int raw, luma, bfrag, scratch, bar, luma_bytes, total;
};

__host__ __device__ inline Layout make_layout(int sub_bytes, int pitch_bytes, int hb_words) {
    Layout L;
    int off = 0;
    auto take = [&](int bytes, int align) {
        off = (off + align - 1) / align * align;
        int at = off;
        off += bytes;
        return at;
    };
    L.luma_bytes = (kCR * pitch_bytes + 127) / 128 * 128;
    L.raw = take(kSlots * sub_bytes, 128);
    L.luma = take(2 * L.luma_bytes, 128);
    L.bfrag = take(hb_words * 8, 16);
    L.scratch = take(kTap * 8 * kHP, 16);
    L.bar = take((kSlots + 4) * 8 + kSlots * 4, 8);
    L.total = off;
    return L;
}

template <int C, int GP>
__global__ void __launch_bounds__(kThreads, 2) ke_resize_mma_kernel(const ResizeArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int row_bytes = a.w * C;
    const int sub_rows = a.sub_rows, sub_bytes = sub_rows * row_bytes, pitch_bytes = a.pitch_bytes;
    const Layout L = make_layout(sub_bytes, pitch_bytes, a.hb_words);
    uint8_t* s_raw = smem + L.raw;
    uint8_t* s_luma = smem + L.luma;
    uint2* s_b = reinterpret_cast<uint2*>(smem + L.bfrag);
    uint8_t* s_scr = smem + L.scratch;
    uint64_t* s_full = reinterpret_cast<uint64_t*>(smem + L.bar);  // raw slots
    uint64_t* l_full = s_full + kSlots;                            // luma chunk ring [2]
    uint64_t* l_empty = l_full + 2;
    uint32_t* s_cnt = reinterpret_cast<uint32_t*>(l_empty + 2);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_sub = (a.h + sub_rows - 1) / sub_rows;
    const int subs_per_chunk = kCR / sub_rows;

    if (tid == 0) {
        for (int b = 0; b < kSlots; ++b) {
            mbar_init(&s_full[b], 1);
            s_cnt[b] = 0u;
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&l_full[b], kLuma);
            mbar_init(&l_empty[b], kTap);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < a.hb_words; i += kThreads) s_b[i] = __ldg(a.hb + i);
    for (int i = tid; i < 2 * L.luma_bytes / 4; i += kThreads) reinterpret_cast<uint32_t*>(s_luma)[i] = 0u;
    for (int i = tid; i < kTap * 8 * kHP / 4; i += kThreads) reinterpret_cast<uint32_t*>(s_scr)[i] = 0u;
    __syncthreads();

    if (warp >= kTap) {
        // ===== luma warps (see ke_phash.cu v5): the last reader of a raw slot refills it =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kLumaRegs));
        const int lw = warp - kTap;
        const long long my_images = blockIdx.x < a.n ? (a.n - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const uint32_t total_seq = (uint32_t)(my_images * n_sub);
        auto issue = [&](uint32_t q) {
            const uint32_t k = q / (uint32_t)n_sub, sq = q - k * (uint32_t)n_sub;
            const int b = (int)(q % kSlots);
            const int rows = min(sub_rows, a.h - (int)sq * sub_rows);
            mbar_expect_tx(&s_full[b], (uint32_t)(rows * row_bytes));
            bulk_g2s(s_raw + b * sub_bytes, a.img + (blockIdx.x + (long long)k * gridDim.x) * a.img_stride + (long long)sq * sub_bytes,
                     (uint32_t)(rows * row_bytes), &s_full[b]);
        };
        auto release = [&](uint32_t seq_, int b) {
            uint32_t old;
            asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(&s_cnt[b])) : "memory");
            if ((old & (kLuma - 1)) == kLuma - 1 && seq_ + kSlots < total_seq) issue(seq_ + kSlots);
        };
        if (lw == 0 && lane == 0)
            for (uint32_t q = 0; q < kSlots && q < total_seq; ++q) issue(q);
        const bool act = lane < (a.w >> 4);
        uint32_t seq = 0, chunk = 0;
        for (long long im = blockIdx.x; im < a.n; im += gridDim.x) {
            for (int r0 = 0; r0 < a.h; r0 += kCR, ++chunk) {
                const int lb = chunk & 1;
                mbar_wait(&l_empty[lb], ((chunk >> 1) & 1u) ^ 1u);
                uint8_t* dst8 = s_luma + lb * L.luma_bytes;
                for (int s = 0; s < subs_per_chunk && r0 + s * sub_rows < a.h; ++s, ++seq) {
                    const int b = (int)(seq % kSlots);
                    const int srows = min(sub_rows, a.h - (r0 + s * sub_rows));
                    mbar_wait(&s_full[b], (seq / kSlots) & 1u);
                    for (int rr = lw; rr < srows; rr += 2 * kLuma) {
                        const bool r_two = rr + kLuma < srows;
                        const uint8_t* s0 = s_raw + b * sub_bytes + rr * row_bytes;
                        const uint8_t* s1 = s0 + kLuma * row_bytes;
                        uint8_t* d0 = dst8 + (s * sub_rows + rr) * pitch_bytes;
                        uint8_t* d1 = d0 + kLuma * pitch_bytes;
                        if (C == 3) {
                            const bool one = act, two = act && r_two;
                            uint4 x0[3], x1[3];
#pragma unroll
                            for (int i = 0; i < 3; ++i) {
                                x0[i] = one ? reinterpret_cast<const uint4*>(s0 + 48 * lane)[i] : make_uint4(0, 0, 0, 0);
                                x1[i] = two ? reinterpret_cast<const uint4*>(s1 + 48 * lane)[i] : make_uint4(0, 0, 0, 0);
                            }
                            if (one) *reinterpret_cast<uint4*>(d0 + 16 * lane) = luma16_rgb(x0[0], x0[1], x0[2]);
                            if (two) *reinterpret_cast<uint4*>(d1 + 16 * lane) = luma16_rgb(x1[0], x1[1], x1[2]);
                        } else if (C == 4) {
                            const int pieces = a.w >> 2;
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const bool in = 32 * i + lane < pieces;
                                if (in) reinterpret_cast<uint32_t*>(d0)[32 * i + lane] = luma4_rgba(reinterpret_cast<const uint4*>(s0)[32 * i + lane]);
                                if (in && r_two)
                                    reinterpret_cast<uint32_t*>(d1)[32 * i + lane] = luma4_rgba(reinterpret_cast<const uint4*>(s1)[32 * i + lane]);
                            }
                        } else {
                            if (act) reinterpret_cast<uint4*>(d0)[lane] = reinterpret_cast<const uint4*>(s0)[lane];
                            if (act && r_two) reinterpret_cast<uint4*>(d1)[lane] = reinterpret_cast<const uint4*>(s1)[lane];
                        }
                    }
                    __syncwarp();
                    if (lane == 0) release(seq, b);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&l_full[lb]);
            }
        }
        return;
    }

    // ===== tap warps: warp q owns output column groups q (and q + 8 when GP == 2) end to end =====
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kTapRegs));
    const int g = lane >> 2, t = lane & 3;
    const uint32_t row_off = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * pitch_bytes + (lane >> 4) * 16);
    uint8_t* scr = s_scr + warp * (8 * kHP);
    const uint32_t* col = reinterpret_cast<const uint32_t*>(scr + g * kHP);
    int32_t vc[GP][2][3][4];  // per group: two live units (slot = unit & 1), three digits
    uint32_t chunk = 0;
    for (long long im = blockIdx.x; im < a.n; im += gridDim.x) {
#pragma unroll
        for (int gp = 0; gp < GP; ++gp)
#pragma unroll
            for (int sl = 0; sl < 2; ++sl)
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int i = 0; i < 4; ++i) vc[gp][sl][d][i] = d == 0 ? (1 << (kPrec - 1)) : 0;
        uint8_t* out_img = a.out + im * (long long)a.out_h * a.out_w;
        int ci = 0;
        for (int r0 = 0; r0 < a.h; r0 += kCR, ++chunk, ++ci) {
            const int lb = chunk & 1;
            mbar_wait_sleep(&l_full[lb], (chunk >> 1) & 1u, 200);
            const uint32_t luma_addr = smem_u32(s_luma + lb * L.luma_bytes) + row_off;
#pragma unroll
            for (int gp = 0; gp < GP; ++gp) {
                const int grp = warp + kTap * gp;
                if (grp >= a.n_groups) continue;  // idle tap warps still take part in the luma ring hand-shake below
                const int nk = a.h_nk[grp];
                const uint2* bw = s_b + a.h_off[grp] + lane;
                const uint32_t a_addr = luma_addr + a.h_k0[grp] * 32;
                if (GP == 1) {
                    // horizontal: two 16-row blocks x three digit tiles over the group's band, one B load per two MMAs
                    int32_t c[2][3][4];
#pragma unroll
                    for (int rb = 0; rb < 2; ++rb)
#pragma unroll
                        for (int tl = 0; tl < 3; ++tl)
#pragma unroll
                            for (int i = 0; i < 4; ++i) c[rb][tl][i] = tl == 0 ? (1 << (kPrec - 1)) : 0;
#pragma unroll 2
                    for (int k = 0; k < nk; ++k) {
                        uint32_t a0[4], a1[4];
                        ldmatrix_x4(a0, a_addr + k * 32);
                        ldmatrix_x4(a1, a_addr + k * 32 + 16 * pitch_bytes);
#pragma unroll
                        for (int tl = 0; tl < 3; ++tl) {
                            const uint2 b = bw[(k * 3 + tl) * 32];
                            mma_u8s8(c[0][tl], a0, b);
                            mma_u8s8(c[1][tl], a1, b);
                        }
                    }
#pragma unroll
                    for (int rb = 0; rb < 2; ++rb)
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            const int32_t v0 = c[rb][0][2 * hf] + (c[rb][1][2 * hf] << 8) + (c[rb][2][2 * hf] << 16);
                            const int32_t v1 = c[rb][0][2 * hf + 1] + (c[rb][1][2 * hf + 1] << 8) + (c[rb][2][2 * hf + 1] << 16);
                            const int row = rb * 16 + hf * 8 + g;
                            scr[(2 * t) * kHP + row] = (uint8_t)pack_sat_u8(0, v0 >> kPrec);
                            scr[(2 * t + 1) * kHP + row] = (uint8_t)pack_sat_u8(0, v1 >> kPrec);
                        }
                } else {
                    // two groups per warp: one 16-row block at a time keeps the horizontal accumulators at 12 registers
#pragma unroll
                    for (int rb = 0; rb < 2; ++rb) {
                        int32_t c[3][4];
#pragma unroll
                        for (int tl = 0; tl < 3; ++tl)
#pragma unroll
                            for (int i = 0; i < 4; ++i) c[tl][i] = tl == 0 ? (1 << (kPrec - 1)) : 0;
                        for (int k = 0; k < nk; ++k) {
                            uint32_t a0[4];
                            ldmatrix_x4(a0, a_addr + k * 32 + rb * 16 * pitch_bytes);
#pragma unroll
                            for (int tl = 0; tl < 3; ++tl) mma_u8s8(c[tl], a0, bw[(k * 3 + tl) * 32]);
                        }
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            const int32_t v0 = c[0][2 * hf] + (c[1][2 * hf] << 8) + (c[2][2 * hf] << 16);
                            const int32_t v1 = c[0][2 * hf + 1] + (c[1][2 * hf + 1] << 8) + (c[2][2 * hf + 1] << 16);
                            const int row = rb * 16 + hf * 8 + g;
                            scr[(2 * t) * kHP + row] = (uint8_t)pack_sat_u8(0, v0 >> kPrec);
                            scr[(2 * t + 1) * kHP + row] = (uint8_t)pack_sat_u8(0, v1 >> kPrec);
                        }
                    }
                }
                __syncwarp();  // the eight columns of this chunk are in the scratch
                // vertical: this chunk is one k-step for every live unit; a unit is flushed after its last chunk
                const uint32_t b0 = col[t], b1 = col[4 + t];
                const int out_col = a.col_begin + 8 * grp + 2 * t;
#pragma unroll
                for (int u = 0; u < kMaxUnits; ++u) {
                    if (u >= a.n_units || ci < a.v_lo[u] || ci > a.v_hi[u]) continue;
                    const uint4* af = a.va + ((size_t)(ci * a.n_units + u) * 3) * 32 + lane;
#pragma unroll
                    for (int d = 0; d < 3; ++d) mma_s8u8(vc[gp][u & 1][d], __ldg(af + d * 32), b0, b1);
                    if (ci == a.v_hi[u]) {
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            const int32_t v0 = vc[gp][u & 1][0][2 * hf] + (vc[gp][u & 1][1][2 * hf] << 8) + (vc[gp][u & 1][2][2 * hf] << 16);
                            const int32_t v1 = vc[gp][u & 1][0][2 * hf + 1] + (vc[gp][u & 1][1][2 * hf + 1] << 8) +
                                               (vc[gp][u & 1][2][2 * hf + 1] << 16);
                            *reinterpret_cast<uint16_t*>(out_img + (long long)(16 * u + hf * 8 + g) * a.out_w + out_col) =
                                (uint16_t)pack_sat_u8(v1 >> kPrec, v0 >> kPrec);
                        }
#pragma unroll
                        for (int d = 0; d < 3; ++d)
#pragma unroll
                            for (int i = 0; i < 4; ++i) vc[gp][u & 1][d][i] = d == 0 ? (1 << (kPrec - 1)) : 0;
                    }
                }
                __syncwarp();  // every lane has read the scratch columns before they are overwritten
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&l_empty[lb]);  // this warp no longer reads the luma buffer
        }
    }
}

// ------------------------------------------------------------------ host side

struct MmaTables {
    uint2* d_hb = nullptr;
    uint4* d_va = nullptr;
    int hb_words = 0;
    int k0[16] = {}, nk[16] = {}, off[16] = {};  // per 8-column output group
    int n_groups = 0, n_units = 0;
    int v_lo[kMaxUnits] = {}, v_hi[kMaxUnits] = {};
    bool ok = false;
};

int digit_of(int32_t k, int d, bool& fits) {
    int dd[3];
    for (int i = 0; i < 3; ++i) {
        dd[i] = ((k & 255) ^ 128) - 128;
        k = (k - dd[i]) >> 8;
    }
    if (k != 0) fits = false;
    return dd[d];
}

}  // namespace

struct KeResizeMmaCache {
    std::map<std::tuple<int, int, int, int, int>, MmaTables> tables;  // (w, h, out_w, out_h, filter)
};

void ke_resize_mma_tables_free(KeResizeMmaCache* cache) {
    if (!cache) return;
    for (auto& kv : cache->tables) {
        cudaFree(kv.second.d_hb);
        cudaFree(kv.second.d_va);
    }
    delete cache;
}

namespace {

int build_tables(int w, int h, int out_w, int out_h, int filter, MmaTables& T) {
    std::vector<int32_t> kh, bh, kv, bv;
    int ksh = 0, ksv = 0;
    ke_pillow_table(w, out_w, filter, kh, bh, ksh);
    ke_pillow_table(h, out_h, filter, kv, bv, ksv);
    bool fits = true;
    auto tap_h = [&](int o, int x) -> int32_t {
        const int tp = x - bh[2 * o];
        return (o < out_w && tp >= 0 && tp < bh[2 * o + 1]) ? kh[(size_t)o * ksh + tp] : 0;
    };
    auto tap_v = [&](int yy, int y) -> int32_t {
        const int tp = y - bv[2 * yy];
        return (yy < out_h && y < h && tp >= 0 && tp < bv[2 * yy + 1]) ? kv[(size_t)yy * ksv + tp] : 0;
    };
    T.n_groups = out_w / 8;
    T.n_units = out_h / 16;
    std::vector<uint2> hb;
    for (int gq = 0; gq < T.n_groups; ++gq) {
        int first = w, last = 0;
        for (int o = 8 * gq; o < 8 * gq + 8; ++o) {
            first = std::min(first, bh[2 * o]);
            last = std::max(last, bh[2 * o] + bh[2 * o + 1] - 1);
        }
        T.k0[gq] = first / 32;
        T.nk[gq] = last / 32 - T.k0[gq] + 1;
        T.off[gq] = (int)hb.size();
        for (int k = 0; k < T.nk[gq]; ++k)
            for (int d = 0; d < 3; ++d)
                for (int lane = 0; lane < 32; ++lane) {
                    const int n = lane >> 2, t4 = lane & 3;
                    uint32_t wd[2] = {0, 0};
                    for (int half = 0; half < 2; ++half)
                        for (int i = 0; i < 4; ++i) {
                            const int x = (T.k0[gq] + k) * 32 + half * 16 + 4 * t4 + i;
                            wd[half] |= ((uint32_t)digit_of(tap_h(8 * gq + n, x), d, fits) & 0xFFu) << (8 * i);
                        }
                    hb.push_back(make_uint2(wd[0], wd[1]));
                }
    }
    const int nch = (h + kCR - 1) / kCR;
    std::vector<uint4> va((size_t)nch * T.n_units * 3 * 32, make_uint4(0, 0, 0, 0));
    for (int u = 0; u < T.n_units; ++u) T.v_lo[u] = nch, T.v_hi[u] = -1;
    for (int c = 0; c < nch; ++c)
        for (int u = 0; u < T.n_units; ++u)
            for (int lane = 0; lane < 32; ++lane) {
                const int g = lane >> 2, t4 = lane & 3;
                uint32_t reg[3][4] = {};
                for (int ri = 0; ri < 4; ++ri) {
                    const int r = g + 8 * (ri & 1), kb = 16 * (ri >> 1) + 4 * t4;
                    for (int i = 0; i < 4; ++i) {
                        const int32_t k = tap_v(16 * u + r, kCR * c + kb + i);
                        if (k) T.v_lo[u] = std::min(T.v_lo[u], c), T.v_hi[u] = std::max(T.v_hi[u], c);
                        for (int d = 0; d < 3; ++d) reg[d][ri] |= ((uint32_t)digit_of(k, d, fits) & 0xFFu) << (8 * i);
                    }
                }
                for (int d = 0; d < 3; ++d)
                    va[(((size_t)c * T.n_units + u) * 3 + d) * 32 + lane] = make_uint4(reg[d][0], reg[d][1], reg[d][2], reg[d][3]);
            }
    // at most two live units, of different parity: unit u+2 must start after unit u has ended
    bool two_live = true;
    for (int u = 0; u + 2 < T.n_units; ++u)
        if (T.v_lo[u + 2] <= T.v_hi[u]) two_live = false;
    for (int u = 0; u < T.n_units; ++u)
        if (T.v_hi[u] < 0) two_live = false;
    T.ok = fits && two_live;
    if (!T.ok) return KE_OK;
    T.hb_words = (int)hb.size();
    KE_CUDA(cudaMalloc((void**)&T.d_hb, hb.size() * sizeof(uint2)));
    KE_CUDA(cudaMalloc((void**)&T.d_va, va.size() * sizeof(uint4)));
    KE_CUDA(cudaMemcpy(T.d_hb, hb.data(), hb.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    KE_CUDA(cudaMemcpy(T.d_va, va.data(), va.size() * sizeof(uint4), cudaMemcpyHostToDevice));
    return KE_OK;
}

template <int C, int GP>
int launch(ke_ctx* ctx, const ResizeArgs& a, int smem, cudaStream_t s) {
    KE_CUDA(cudaFuncSetAttribute(ke_resize_mma_kernel<C, GP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int per_sm = 0;
    KE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ke_resize_mma_kernel<C, GP>, kThreads, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)ctx->sm_count * per_sm;
    if (grid > a.n) grid = a.n;
    ke_resize_mma_kernel<C, GP><<<(unsigned)grid, kThreads, smem, s>>>(a);
    ctx->launches++;
    KE_CUDA(cudaGetLastError());
    return KE_OK;
}

}  // namespace

// Returns KE_OK and sets *taken = 1 when the streaming kernel served the call, *taken = 0 when the shape is not its.
int ke_gray_resize_mma(ke_ctx* ctx, const uint8_t* d_img, int64_t n, int h, int w, int c, int64_t img_stride,
                       int64_t row_stride, int out_w, int out_h, int filter, uint8_t* d_out, cudaStream_t s, int* taken) {
    *taken = 0;
    const long long row_bytes = (long long)w * c;
    if (row_stride != row_bytes || (row_bytes & 15) || (img_stride & 15) || (w & 15) || w > 512 || w == out_w || h == out_h ||
        (reinterpret_cast<uintptr_t>(d_img) & 15) || (out_w & 7) || (out_h & 15) || out_h > 16 * kMaxUnits || out_w > 128 ||
        (reinterpret_cast<uintptr_t>(d_out) & 1))
        return KE_OK;
    if (!ctx->resize_mma) ctx->resize_mma = new KeResizeMmaCache();
    auto key = std::make_tuple(w, h, out_w, out_h, filter);
    auto it = ctx->resize_mma->tables.find(key);
    if (it == ctx->resize_mma->tables.end()) {
        MmaTables T;
        int rc = build_tables(w, h, out_w, out_h, filter, T);
        if (rc) return rc;
        it = ctx->resize_mma->tables.emplace(key, T).first;
    }
    const MmaTables& T = it->second;
    if (!T.ok) return KE_OK;
    const int pitch_bytes = (w + 31) / 32 * 32 + 16;
    int sub_rows = 0;
    {
        int max_words = 0;
        for (int g0 = 0; g0 < T.n_groups; g0 += kMaxGroups) {
            const int ng = std::min(kMaxGroups, T.n_groups - g0);
            max_words = std::max(max_words, (g0 + ng < T.n_groups ? T.off[g0 + ng] : T.hb_words) - T.off[g0]);
        }
        for (int sub : {16, 8, 4, 2, 1})
            if (make_layout((int)(sub * row_bytes), pitch_bytes, max_words).total <= 113 * 1024 &&
                n * ((h + sub - 1) / sub) < (1ll << 31)) {
                sub_rows = sub;
                break;
            }
    }
    if (!sub_rows) return KE_OK;
    for (int g0 = 0; g0 < T.n_groups; g0 += kMaxGroups) {  // up to 128 output columns per launch
        ResizeArgs a;
        a.img = d_img;
        a.n = n;
        a.img_stride = img_stride;
        a.h = h;
        a.w = w;
        a.pitch_bytes = pitch_bytes;
        a.n_groups = std::min(kMaxGroups, T.n_groups - g0);
        a.col_begin = 8 * g0;
        a.out_w = out_w;
        a.out_h = out_h;
        a.n_units = T.n_units;
        a.hb = T.d_hb + T.off[g0];
        a.hb_words = (g0 + a.n_groups < T.n_groups ? T.off[g0 + a.n_groups] : T.hb_words) - T.off[g0];
        for (int q = 0; q < kMaxGroups; ++q) {
            const bool in = q < a.n_groups;
            a.h_k0[q] = in ? T.k0[g0 + q] : 0;
            a.h_nk[q] = in ? T.nk[g0 + q] : 0;
            a.h_off[q] = in ? T.off[g0 + q] - T.off[g0] : 0;
        }
        a.va = T.d_va;
        for (int u = 0; u < kMaxUnits; ++u) a.v_lo[u] = T.v_lo[u], a.v_hi[u] = T.v_hi[u];
        a.out = d_out;
        a.sub_rows = sub_rows;
        const int smem = make_layout((int)(sub_rows * row_bytes), pitch_bytes, a.hb_words).total;
        int rc;
        const bool two = a.n_groups > kTap;
        switch (c) {
            case 1: rc = two ? launch<1, 2>(ctx, a, smem, s) : launch<1, 1>(ctx, a, smem, s); break;
            case 3: rc = two ? launch<3, 2>(ctx, a, smem, s) : launch<3, 1>(ctx, a, smem, s); break;
            default: rc = two ? launch<4, 2>(ctx, a, smem, s) : launch<4, 1>(ctx, a, smem, s); break;
        }
        if (rc) return rc;
    }
    *taken = 1;
    return KE_OK;
}
