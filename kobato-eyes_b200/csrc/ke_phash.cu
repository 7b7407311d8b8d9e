// ke_phash.cu — K1: batched pHash + dHash of decoded uint8 images, sm_100a.
//
// Replaces sig.phash.phash / dhash (reference src/sig/phash.py:21-57): convert("L") ->
// resize(LANCZOS) to 32x32 and 9x8 (Pillow's fixed-point arithmetic, byte-identical) -> DCT-II of
// the 32x32 plane (FP64 on CUDA cores; tensor cores deliberately unused: a bit must not flip at
// the threshold) -> 8x8 low block compared with the mean of its 63 AC terms -> 64-bit hash;
// dHash = left<right compares on the 8x9 plane.
//
// Two kernels:
//   * ke_phash_v5_kernel (default; "tensor-core kernel (v5)" below): persistent CTAs, raw rows by 1-D TMA bulk copies into
//     a ring of sub-chunks, luma warps (dp2a), and eight tap warps that run BOTH Lanczos resamples exactly on the
//     tensor pipe (mma.sync u8 x s8 over balanced base-256 tap digits), then the FP64 DCT.  It takes every batch of
//     contiguous rows: any width (the resample bands' B fragments sit in registers up to ~2200 pixels — two CTAs per SM up
//     to ~512 pixels, one CTA per SM with setmaxnreg-enlarged tap warps beyond — and behind pointers into shared memory /
//     L2 for longer rows), any byte alignment (the copies move the 16-byte aligned superset of a sub-chunk and the luma
//     warps funnel-shift), 'L' / RGB / RGBA.
//   * ke_phash_kernel (generic): one CTA of 256 threads per image stream, dp4a taps on CUDA cores, plain loads when the
//     rows are strided.  It takes everything else (row_stride != w * c) and is the in-library reference the parity
//     tests compare v5 with (KE_OPT_PHASH_GENERIC).  Its steps:
//   1. raw rows HBM -> shared memory with one 1-D TMA bulk copy (cp.async.bulk + mbarrier;
//      16-byte aligned superset of the chunk), issued one chunk ahead of the compute;
//   2. luma: L = (R*19595 + G*38470 + B*7471 + 0x8000) >> 16 with dp4a on the packed bytes
//      -> uint8 luma rows in shared memory (odd word pitch: lane=row reads are conflict free);
//   3. horizontal Lanczos taps for both targets (32 and 9 outputs per row).  Taps are 22-bit
//      fixed point split into three byte planes so that 4 pixels x 1 plane = one dp4a:
//      sum = D0 + 256*D1 + 65536*D2 (mod 2^32, exact because the true sum fits in int32).
//      Lanes map to rows, so the tap words are warp-uniform (shared-memory broadcast);
//   4. (2^21 + sum) >> 22, clip to uint8 -> the chunk's rows of the [H,32] and [H,9] planes;
//   5. vertical taps are accumulated on the fly into per-thread registers (each thread owns 4
//      of the 32x32 outputs; 72 threads own the 8x9 outputs), so no [H,32] plane is ever stored.
// After the last chunk: vertical rounding, FP64 DCT (8x32 * 32x32 * 32x8), warp-shuffle mean
// of the AC terms, ballot -> hash bits (first element = MSB).
//
// The aligned-superset copies of v5 read up to 15 bytes in front of the first and behind the last image of an
// unaligned batch: inside the caller's allocation (device allocations are at least 256-byte granular), never written.
//
// Algorithmic HBM bytes per image: h*w*c read + 16 written.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>

#include "ke_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kOutW = 32, kOutH = 32, kDW = 9, kDH = 8;
constexpr int kOuts = kOutW + kDW;  // 41 horizontal outputs per row
constexpr int kPrec = 22;

// ------------------------------------------------------------------ host-side table cache

struct HTable {           // horizontal pass, one per input width
    uint4* d_coef = nullptr;   // packed byte-plane tap words, outputs back to back
    int4* d_items = nullptr;   // work items {out, word_begin, word_count, 0}
    int* d_meta = nullptr;     // [kOuts] first pixel word, [kOuts] offset into d_coef, [kWarps+1] list starts
    int n_items = 0;
    int coef_words = 0;
    // tensor-core kernel (v5): per tap warp a run of mma.m16n8k32 B fragments [k-step][tile][lane]
    uint2* d_mma_b = nullptr;
    int mma_words = 0;           // uint2 words in d_mma_b (0: this width is not served by v5)
    int mma_k0[8] = {}, mma_nk[8] = {}, mma_boff[8] = {};
    // narrow target split by k range: warp 4 + j owns k-steps [mmaq_k0[j], + mmaq_nk[j]) of ALL nine outputs (four tiles)
    int mmaq_k0[4] = {}, mmaq_nk[4] = {}, mmaq_boff[4] = {}, mmaq_kq = 0;
    // narrow target in two output groups x two halves of the group's band (the two-CTA kernel): warp 4 + 2 grp + half
    int mmah_k0[4] = {}, mmah_nk[4] = {}, mmah_boff[4] = {}, mmah_end = 0;
};
struct VTable {           // vertical pass, one per input height
    int* d_kk32 = nullptr;
    int* d_b32 = nullptr;
    int* d_kk8 = nullptr;
    int* d_b8 = nullptr;
    int ks32 = 0, ks8 = 0;
    // tensor-core vertical pass (v5): A fragments [chunk of 32 rows][unit][digit][lane]; unit 0/1 = output rows
    // 0..15 / 16..31 of the 32x32 plane, unit 2 = the 8 rows of the 8x9 plane
    uint4* d_vmma = nullptr;
    int vmma_ok = 0;
    int v_lo[3] = {}, v_hi[3] = {};  // chunks [lo, hi] in which a unit has non-zero taps
};

}  // namespace

struct KeTableCache {
    std::map<int, HTable> h;
    std::map<int, VTable> v;
};

void ke_tables_free(KeTableCache* cache) {
    if (!cache) return;
    for (auto& kv : cache->h) {
        cudaFree(kv.second.d_coef);
        cudaFree(kv.second.d_items);
        cudaFree(kv.second.d_meta);
        cudaFree(kv.second.d_mma_b);
    }
    for (auto& kv : cache->v) {
        cudaFree(kv.second.d_kk32);
        cudaFree(kv.second.d_b32);
        cudaFree(kv.second.d_kk8);
        cudaFree(kv.second.d_b8);
        cudaFree(kv.second.d_vmma);
    }
    delete cache;
}

namespace {

template <typename T>
int upload(const std::vector<T>& v, T** d) {
    KE_CUDA(cudaMalloc((void**)d, std::max<size_t>(v.size(), 1) * sizeof(T)));
    if (!v.empty()) KE_CUDA(cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return KE_OK;
}

int get_htable(ke_ctx* ctx, int w, const HTable** out) {
    if (!ctx->tables) ctx->tables = new KeTableCache();
    auto it = ctx->tables->h.find(w);
    if (it != ctx->tables->h.end()) {
        *out = &it->second;
        return KE_OK;
    }
    std::vector<uint4> coef;
    std::vector<int> meta(2 * kOuts);
    std::vector<int> nwords(kOuts);
    const int outs[2] = {kOutW, kDW};
    int o_base = 0;
    for (int tbl = 0; tbl < 2; ++tbl) {
        const int ow = outs[tbl];
        const int ks = ke_resample_ksize(w, ow);
        std::vector<int32_t> kk((size_t)ks * ow), bd(2 * (size_t)ow);
        int rc = ke_resample_table(w, ow, kk.data(), bd.data(), ks);
        if (rc) return rc;
        for (int o = 0; o < ow; ++o) {
            const int first = bd[2 * o], count = bd[2 * o + 1];
            const int w0 = first / 4, lead = first % 4;
            const int nw = (lead + count + 3) / 4;
            meta[o_base + o] = w0;
            meta[kOuts + o_base + o] = (int)coef.size();
            nwords[o_base + o] = nw;
            for (int t = 0; t < nw; ++t) {
                uint32_t plane[3] = {0, 0, 0};
                for (int b = 0; b < 4; ++b) {
                    const int tap = 4 * t + b - lead;
                    const int32_t k = (tap >= 0 && tap < count) ? kk[(size_t)o * ks + tap] : 0;
                    // k = b0 + 256*b1 + 65536*b2 with b0,b1 unsigned bytes and b2 a signed byte
                    const uint32_t u = (uint32_t)k;
                    const int32_t top = k >> 16;  // arithmetic: floor(k / 65536)
                    if (top < -128 || top > 127) {
                        ke_set_error("resample tap %d does not fit 24 bits (w=%d -> %d)", k, w, ow);
                        return KE_E_UNSUPPORTED;
                    }
                    plane[0] |= (u & 0xFFu) << (8 * b);
                    plane[1] |= ((u >> 8) & 0xFFu) << (8 * b);
                    plane[2] |= ((uint32_t)top & 0xFFu) << (8 * b);
                }
                coef.push_back(make_uint4(plane[0], plane[1], plane[2], 0u));
            }
        }
        o_base += ow;
    }
    // Work items: split every output's word range into pieces no longer than the longest
    // 32-wide output so the 8 warps get pieces of similar cost.
    int piece = 1;
    for (int o = 0; o < kOutW; ++o) piece = std::max(piece, nwords[o]);
    std::vector<int4> items;
    for (int o = 0; o < kOuts; ++o) {
        const int parts = (nwords[o] + piece - 1) / piece;
        for (int p = 0; p < parts; ++p) {
            const int b = (int)((long long)nwords[o] * p / parts), e = (int)((long long)nwords[o] * (p + 1) / parts);
            if (e > b) items.push_back(make_int4(o, b, e - b, 0));
        }
    }
    std::stable_sort(items.begin(), items.end(), [](const int4& a, const int4& b) { return a.z > b.z; });
    // Tensor-core tables (v5).  The horizontal pass is the banded product  luma[rows, w] x taps[w, 41 x 3 digits]:
    // taps are written in balanced base-256 digits k = d0 + 256 d1 + 65536 d2, every d in [-128, 127], so one
    // mma.m16n8k32.s32.u8.s8 per (16 rows, 8 columns, 32 pixels) accumulates a digit plane exactly.
    // Tap warp q < 4: outputs 8q..8q+7 of the 32-wide target, three tiles (= digits) over one k range.
    // Tap warps 4..7: the 9-wide target, one tile per output pair (a, b) with columns
    // {a.d0, a.d1, b.d0, b.d1, a.d2, b.d2, 0, 0}; warp 7 carries the pairs (6,7) and (8,-).
    std::vector<uint2> mma_b;
    int mma_k0[8], mma_nk[8], mma_boff[8];
    int mmaq_k0[4] = {}, mmaq_nk[4] = {}, mmaq_boff[4] = {}, mmaq_kq = 0;
    int mmah_k0[4] = {}, mmah_nk[4] = {}, mmah_boff[4] = {}, mmah_end = 0;
    int mma_words_ptr = 0;  // words of the per-output-band tables above (what the pointer-fed kernel may copy to shared memory)
    bool mma_ok = true;  // any width: taps are zero outside an output's support, so the padding of the last k-step adds nothing
    if (mma_ok) {
        std::vector<int32_t> kkP, bdP, kkD, bdD;
        const int ksP = ke_resample_ksize(w, kOutW), ksD = ke_resample_ksize(w, kDW);
        kkP.resize((size_t)ksP * kOutW), bdP.resize(2 * kOutW), kkD.resize((size_t)ksD * kDW), bdD.resize(2 * kDW);
        int rc = ke_resample_table(w, kOutW, kkP.data(), bdP.data(), ksP);
        if (rc) return rc;
        if ((rc = ke_resample_table(w, kDW, kkD.data(), bdD.data(), ksD))) return rc;
        auto tap = [&](int out, int x) -> int32_t {  // out: 0..31 wide target, 32..40 narrow target, else none
            if (out < 0 || out >= kOuts) return 0;
            const bool P = out < kOutW;
            const int o = P ? out : out - kOutW;
            const int first = P ? bdP[2 * o] : bdD[2 * o], count = P ? bdP[2 * o + 1] : bdD[2 * o + 1];
            const int t = x - first;
            if (t < 0 || t >= count) return 0;
            return P ? kkP[(size_t)o * ksP + t] : kkD[(size_t)o * ksD + t];
        };
        auto digit = [&](int32_t k, int d) -> int {
            int dd[3];
            for (int i = 0; i < 3; ++i) {
                dd[i] = ((k & 255) ^ 128) - 128;
                k = (k - dd[i]) >> 8;
            }
            if (k != 0) mma_ok = false;  // does not fit three balanced digits
            return dd[d];
        };
        auto range = [&](int o_lo, int o_hi, int& k0, int& nk) {  // k-steps of 32 pixels touched by outputs [o_lo, o_hi]
            int first = w, last = 0;
            for (int o = o_lo; o <= o_hi && o < kOuts; ++o) {
                const bool P = o < kOutW;
                const int f = P ? bdP[2 * o] : bdD[2 * (o - kOutW)], c = P ? bdP[2 * o + 1] : bdD[2 * (o - kOutW) + 1];
                first = std::min(first, f), last = std::max(last, f + c - 1);
            }
            k0 = first / 32, nk = last / 32 - k0 + 1;
        };
        for (int wv = 0; wv < 8 && mma_ok; ++wv) {
            int nt, o_lo, o_hi;
            if (wv < 4) nt = 3, o_lo = 8 * wv, o_hi = 8 * wv + 7;
            else if (wv < 7) nt = 1, o_lo = kOutW + 2 * (wv - 4), o_hi = o_lo + 1;
            else nt = 2, o_lo = kOutW + 6, o_hi = kOutW + 8;
            range(o_lo, o_hi, mma_k0[wv], mma_nk[wv]);
            mma_boff[wv] = (int)mma_b.size();
            for (int k = 0; k < mma_nk[wv]; ++k)
                for (int tl = 0; tl < nt; ++tl)
                    for (int lane = 0; lane < 32; ++lane) {
                        const int n = lane >> 2, t4 = lane & 3;
                        int out, dg;
                        if (wv < 4) out = o_lo + n, dg = tl;
                        else {
                            const int a_out = o_lo + 2 * tl, b_out = a_out + 1;
                            static const int sel_out[8] = {0, 0, 1, 1, 0, 1, -1, -1}, sel_d[8] = {0, 1, 0, 1, 2, 2, 0, 0};
                            out = sel_out[n] < 0 ? -1 : (sel_out[n] ? b_out : a_out);
                            dg = sel_d[n];
                        }
                        uint32_t wd[2] = {0, 0};
                        for (int half = 0; half < 2; ++half)
                            for (int i = 0; i < 4; ++i) {
                                const int x = (mma_k0[wv] + k) * 32 + half * 16 + 4 * t4 + i;
                                const int dv = digit(tap(out, x), dg);
                                wd[half] |= ((uint32_t)dv & 0xFFu) << (8 * i);
                            }
                        mma_b.push_back(make_uint2(wd[0], wd[1]));
                    }
        }
    // The same nine outputs split by K instead: narrow warp j takes a quarter of the row's k-steps for every output, so
    // each luma byte is read by ONE narrow warp (not by all four, whose output bands nearly cover the row) and the
    // quarter-band fits registers.  Four tiles per k-step: tile T = {a.d0, a.d1, b.d0, b.d1, a.d2, b.d2, e0, e1} with
    // (a, b) = outputs (2T, 2T + 1); the ninth output rides in the spare columns: e = (8.d0, 8.d1) in tile 0, (8.d2, -) in
    // tile 1.  The warps exchange their partial sums through shared memory (see the kernel).
    if (mma_ok) {
        mma_words_ptr = (int)mma_b.size();
        const int KT = (w + 31) / 32;
        mmaq_kq = (KT + 3) / 4;
        for (int j = 0; j < 4; ++j) {
            mmaq_k0[j] = j * mmaq_kq;
            mmaq_nk[j] = std::max(0, std::min(mmaq_kq, KT - j * mmaq_kq));
            mmaq_boff[j] = (int)mma_b.size();
            for (int k = 0; k < mmaq_nk[j]; ++k)
                for (int T = 0; T < 4; ++T)
                    for (int lane = 0; lane < 32; ++lane) {
                        const int n = lane >> 2, t4 = lane & 3;
                        static const int sel_out[6] = {0, 0, 1, 1, 0, 1}, sel_d[6] = {0, 1, 0, 1, 2, 2};
                        int out = -1, dg = 0;
                        if (n < 6) out = kOutW + 2 * T + sel_out[n], dg = sel_d[n];
                        else if (T == 0) out = kOutW + 8, dg = n - 6;
                        else if (T == 1 && n == 6) out = kOutW + 8, dg = 2;
                        uint32_t wd[2] = {0, 0};
                        for (int half = 0; half < 2; ++half)
                            for (int i = 0; i < 4; ++i) {
                                const int x = (mmaq_k0[j] + k) * 32 + half * 16 + 4 * t4 + i;
                                const int dv = out < 0 ? 0 : digit(tap(out, x), dg);
                                wd[half] |= ((uint32_t)dv & 0xFFu) << (8 * i);
                            }
                        mma_b.push_back(make_uint2(wd[0], wd[1]));
                    }
            }
            // Two-CTA kernel (72 registers in the narrow warps: no room for the K-split): outputs in two groups,
            // (0..3) and (4..8), each group's band cut in two halves -> warp 4 + 2 grp + half multiplies half of the band
            // into the group's two tiles {a.d0, a.d1, b.d0, b.d1, a.d2, b.d2, e0, e1}, (a, b) = (g0 + 2 tile, + 1); the
            // ninth output rides in the spare columns of group 1: e = (8.d0, 8.d1) in tile 0, (8.d2, -) in tile 1.  Each
            // luma byte is read by at most two narrow warps instead of four; the upper half hands its sums to the lower
            // one through shared memory.
            for (int jw = 0; jw < 4; ++jw) {
                const int grp = jw >> 1, half = jw & 1, g0 = kOutW + 4 * grp;
                int k0g, nkg;
                range(g0, grp ? kOutW + 8 : g0 + 3, k0g, nkg);
                const int n_lo = (nkg + 1) / 2;
                mmah_k0[jw] = half ? k0g + n_lo : k0g;
                mmah_nk[jw] = half ? nkg - n_lo : n_lo;
                mmah_boff[jw] = (int)mma_b.size();
                for (int k = 0; k < mmah_nk[jw]; ++k)
                    for (int T = 0; T < 2; ++T)
                        for (int lane = 0; lane < 32; ++lane) {
                            const int n = lane >> 2, t4 = lane & 3;
                            static const int sel_out[6] = {0, 0, 1, 1, 0, 1}, sel_d[6] = {0, 1, 0, 1, 2, 2};
                            int out = -1, dg = 0;
                            if (n < 6) out = g0 + 2 * T + sel_out[n], dg = sel_d[n];
                            else if (grp == 1 && T == 0) out = kOutW + 8, dg = n - 6;
                            else if (grp == 1 && T == 1 && n == 6) out = kOutW + 8, dg = 2;
                            uint32_t wd[2] = {0, 0};
                            for (int hf = 0; hf < 2; ++hf)
                                for (int i = 0; i < 4; ++i) {
                                    const int x = (mmah_k0[jw] + k) * 32 + hf * 16 + 4 * t4 + i;
                                    const int dv = out < 0 ? 0 : digit(tap(out, x), dg);
                                    wd[hf] |= ((uint32_t)dv & 0xFFu) << (8 * i);
                                }
                            mma_b.push_back(make_uint2(wd[0], wd[1]));
                        }
            }
            mmah_end = (int)mma_b.size();
        }
    }
    HTable t;
    t.n_items = (int)items.size();
    t.coef_words = (int)coef.size();
    int rc;
    if ((rc = upload(coef, &t.d_coef))) return rc;
    if ((rc = upload(items, &t.d_items))) return rc;
    if ((rc = upload(meta, &t.d_meta))) return rc;
    if (mma_ok) {
        if ((rc = upload(mma_b, &t.d_mma_b))) return rc;
        t.mma_words = mma_words_ptr;
        for (int i = 0; i < 8; ++i) t.mma_k0[i] = mma_k0[i], t.mma_nk[i] = mma_nk[i], t.mma_boff[i] = mma_boff[i];
        for (int i = 0; i < 4; ++i) t.mmaq_k0[i] = mmaq_k0[i], t.mmaq_nk[i] = mmaq_nk[i], t.mmaq_boff[i] = mmaq_boff[i];
        t.mmaq_kq = mmaq_kq;
        for (int i = 0; i < 4; ++i) t.mmah_k0[i] = mmah_k0[i], t.mmah_nk[i] = mmah_nk[i], t.mmah_boff[i] = mmah_boff[i];
        t.mmah_end = mmah_end;
    }
    auto ins = ctx->tables->h.emplace(w, t);
    *out = &ins.first->second;
    return KE_OK;
}

int get_vtable(ke_ctx* ctx, int h, const VTable** out) {
    if (!ctx->tables) ctx->tables = new KeTableCache();
    auto it = ctx->tables->v.find(h);
    if (it != ctx->tables->v.end()) {
        *out = &it->second;
        return KE_OK;
    }
    VTable t;
    t.ks32 = ke_resample_ksize(h, kOutH);
    t.ks8 = ke_resample_ksize(h, kDH);
    std::vector<int32_t> kk32((size_t)t.ks32 * kOutH), b32(2 * kOutH), kk8((size_t)t.ks8 * kDH), b8(2 * kDH);
    int rc;
    if ((rc = ke_resample_table(h, kOutH, kk32.data(), b32.data(), t.ks32))) return rc;
    if ((rc = ke_resample_table(h, kDH, kk8.data(), b8.data(), t.ks8))) return rc;
    if ((rc = upload(kk32, &t.d_kk32))) return rc;
    if ((rc = upload(b32, &t.d_b32))) return rc;
    if ((rc = upload(kk8, &t.d_kk8))) return rc;
    if ((rc = upload(b8, &t.d_b8))) return rc;
    {
        const int nch = (h + 31) / 32;
        std::vector<uint4> vm((size_t)nch * 3 * 3 * 32, make_uint4(0, 0, 0, 0));
        bool ok = true;
        auto tap = [&](int unit, int r, int y) -> int32_t {  // tap of output row r of the unit at input row y
            if (y >= h) return 0;
            if (unit < 2) {
                const int yy = 16 * unit + r, tpos = y - b32[2 * yy];
                return (tpos >= 0 && tpos < b32[2 * yy + 1]) ? kk32[(size_t)yy * t.ks32 + tpos] : 0;
            }
            if (r >= kDH) return 0;
            const int tpos = y - b8[2 * r];
            return (tpos >= 0 && tpos < b8[2 * r + 1]) ? kk8[(size_t)r * t.ks8 + tpos] : 0;
        };
        for (int u = 0; u < 3; ++u) t.v_lo[u] = nch, t.v_hi[u] = -1;
        for (int c = 0; c < nch; ++c)
            for (int u = 0; u < 3; ++u)
                for (int lane = 0; lane < 32; ++lane) {
                    const int g = lane >> 2, t4 = lane & 3;
                    uint32_t reg[3][4] = {};
                    for (int ri = 0; ri < 4; ++ri) {  // a0: (g, k lo) a1: (g+8, k lo) a2: (g, k hi) a3: (g+8, k hi)
                        const int r = g + 8 * (ri & 1), kb = 16 * (ri >> 1) + 4 * t4;
                        for (int i = 0; i < 4; ++i) {
                            int32_t k = tap(u, r, 32 * c + kb + i);
                            if (k) t.v_lo[u] = std::min(t.v_lo[u], c), t.v_hi[u] = std::max(t.v_hi[u], c);
                            for (int d = 0; d < 3; ++d) {
                                const int dd = ((k & 255) ^ 128) - 128;
                                k = (k - dd) >> 8;
                                reg[d][ri] |= ((uint32_t)dd & 0xFFu) << (8 * i);
                            }
                            if (k != 0) ok = false;
                        }
                    }
                    for (int d = 0; d < 3; ++d)
                        vm[(((size_t)c * 3 + u) * 3 + d) * 32 + lane] = make_uint4(reg[d][0], reg[d][1], reg[d][2], reg[d][3]);
                }
        if (ok) {
            if ((rc = upload(vm, &t.d_vmma))) return rc;
            t.vmma_ok = 1;
        }
    }
    auto ins = ctx->tables->v.emplace(h, t);
    *out = &ins.first->second;
    return KE_OK;
}

// ------------------------------------------------------------------ device helpers

__constant__ double c_dct[8 * 32];  // orthonormal DCT-II rows 0..7 (what cv2.dct applies)

__device__ __forceinline__ uint32_t dp4a_uu(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int32_t dp4a_us(uint32_t a, uint32_t b, int32_t c) {
    int32_t d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
    // consumer side: poll politely so the spinning warps do not take issue slots from the producers
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
__device__ __forceinline__ uint8_t clip8(int32_t v) {
    v >>= kPrec;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

struct PhashArgs {
    const uint8_t* img;
    long long n;
    int h, w, c;
    long long img_stride, row_stride;
    long long total_bytes;  // bytes addressable from img (bounds the aligned superset copies)
    int rows_per_chunk;     // power of two <= 32
    int pitch_words;        // odd
    int raw_bytes;          // shared-memory bytes reserved for a raw chunk
    int use_bulk;           // rows contiguous: chunk = one contiguous byte range
    // tables
    const uint4* coef;
    const int4* items;
    const int* meta;
    int n_items, coef_words;
    const uint2* mma_b;
    int mma_words;
    int mma_k0[8], mma_nk[8], mma_boff[8];
    const uint4* vmma;
    int v_lo[3], v_hi[3];
    const int* kk32;
    const int* b32;
    const int* kk8;
    const int* b8;
    int ks32, ks8;
    // outputs
    uint64_t* phash;
    uint64_t* dhash;
    float* min_margin;
    uint8_t* plane32;
    uint8_t* plane98;
    // K-split narrow-target tables (one-CTA kernels); kept behind the fields the two-CTA kernel reads
    int mmaq_k0[4], mmaq_nk[4], mmaq_boff[4], mmaq_kq;
    int mmah_k0[4], mmah_nk[4], mmah_boff[4], mmah_end;
};

struct SmemLayout {
    int coef, raw, luma, acc, hrow, x32, x98, tmat, ymat, meta, bar, total;
};

__host__ __device__ inline SmemLayout smem_layout(int coef_words, int raw_bytes, int rows, int pitch_words) {
    SmemLayout L;
    int off = 0;
    auto take = [&](int bytes, int align) {
        off = (off + align - 1) / align * align;
        int at = off;
        off += bytes;
        return at;
    };
    L.raw = take(raw_bytes, 128);
    L.coef = take(coef_words * 16, 16);
    L.luma = take(rows * pitch_words * 4, 16);
    L.acc = take(rows * kOuts * 4, 16);
    L.hrow = take(rows * kOuts, 16);
    L.x32 = take(1024, 16);
    L.x98 = take(80, 16);
    L.tmat = take(8 * 32 * 8, 16);
    L.ymat = take(64 * 8, 16);
    L.meta = take(2 * kOuts * 4, 16);
    L.bar = take(8, 8);
    L.total = off;
    return L;
}

// ------------------------------------------------------------------ the kernel

template <int C>
__device__ __forceinline__ void luma_chunk(const uint8_t* __restrict__ raw, uint32_t* __restrict__ luma, int rows, int w,
                                           int pitch_words, bool aligned) {
    // Pillow rgb2l, split into low/high coefficient bytes for dp4a:
    // 19595 = 0x4C8B, 38470 = 0x9646, 7471 = 0x1D2F.
    constexpr uint32_t LO = 0x002F468Bu, HI = 0x001D964Cu;  // bytes (R,G,B,0) little endian
    const int wq = w >> 2;  // full 4-pixel groups per row
    if (C == 1) {
        uint8_t* lb = reinterpret_cast<uint8_t*>(luma);
        for (int idx = threadIdx.x; idx < rows * w; idx += kThreads) {
            const int r = idx / w, x = idx - r * w;
            lb[r * pitch_words * 4 + x] = raw[(size_t)r * w + x];
        }
        return;
    }
    if (aligned && C == 3) {
        for (int idx = threadIdx.x; idx < rows * wq; idx += kThreads) {
            const int r = idx / wq, q = idx - r * wq;
            const uint32_t* src = reinterpret_cast<const uint32_t*>(raw + (size_t)r * w * 3) + q * 3;
            const uint32_t w0 = src[0], w1 = src[1], w2 = src[2];
            // pixel byte positions inside the 12-byte group: p0=0..2, p1=3..5, p2=6..8, p3=9..11
            uint32_t l0 = dp4a_uu(w0, LO, 0x8000u), h0 = dp4a_uu(w0, HI, 0u);
            uint32_t l1 = dp4a_uu(w0, LO << 24, 0x8000u), h1 = dp4a_uu(w0, HI << 24, 0u);
            l1 = dp4a_uu(w1, LO >> 8, l1), h1 = dp4a_uu(w1, HI >> 8, h1);
            uint32_t l2 = dp4a_uu(w1, LO << 16, 0x8000u), h2 = dp4a_uu(w1, HI << 16, 0u);
            l2 = dp4a_uu(w2, LO >> 16, l2), h2 = dp4a_uu(w2, HI >> 16, h2);
            uint32_t l3 = dp4a_uu(w2, LO << 8, 0x8000u), h3 = dp4a_uu(w2, HI << 8, 0u);
            const uint32_t s0 = l0 + (h0 << 8), s1 = l1 + (h1 << 8), s2 = l2 + (h2 << 8), s3 = l3 + (h3 << 8);
            // L = byte 2 of each sum (sums < 2^24)
            const uint32_t lo2 = __byte_perm(s0, s1, 0x0062), hi2 = __byte_perm(s2, s3, 0x0062);
            luma[r * pitch_words + q] = __byte_perm(lo2, hi2, 0x5410);
        }
    } else if (aligned && C == 4) {
        for (int idx = threadIdx.x; idx < rows * wq; idx += kThreads) {
            const int r = idx / wq, q = idx - r * wq;
            const uint4 px = *reinterpret_cast<const uint4*>(raw + (size_t)r * w * 4 + (size_t)q * 16);
            const uint32_t s0 = dp4a_uu(px.x, LO, 0x8000u) + (dp4a_uu(px.x, HI, 0u) << 8);
            const uint32_t s1 = dp4a_uu(px.y, LO, 0x8000u) + (dp4a_uu(px.y, HI, 0u) << 8);
            const uint32_t s2 = dp4a_uu(px.z, LO, 0x8000u) + (dp4a_uu(px.z, HI, 0u) << 8);
            const uint32_t s3 = dp4a_uu(px.w, LO, 0x8000u) + (dp4a_uu(px.w, HI, 0u) << 8);
            const uint32_t lo2 = __byte_perm(s0, s1, 0x0062), hi2 = __byte_perm(s2, s3, 0x0062);
            luma[r * pitch_words + q] = __byte_perm(lo2, hi2, 0x5410);
        }
    }
    // generic bytes: everything when unaligned, else only the w % 4 tail pixels of each row
    const int x_begin = aligned ? (wq << 2) : 0;
    const int tail = w - x_begin;
    if (tail > 0) {
        uint8_t* lb = reinterpret_cast<uint8_t*>(luma);
        for (int idx = threadIdx.x; idx < rows * tail; idx += kThreads) {
            const int r = idx / tail, x = x_begin + (idx - r * tail);
            const uint8_t* p = raw + ((size_t)r * w + x) * C;
            lb[r * pitch_words * 4 + x] = (uint8_t)((p[0] * 19595u + p[1] * 38470u + p[2] * 7471u + 0x8000u) >> 16);
        }
    }
}

template <int C>
__global__ void __launch_bounds__(kThreads) ke_phash_kernel(const PhashArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const SmemLayout L = smem_layout(a.coef_words, a.raw_bytes, a.rows_per_chunk, a.pitch_words);
    uint8_t* s_raw = smem + L.raw;
    uint4* s_coef = reinterpret_cast<uint4*>(smem + L.coef);
    uint32_t* s_luma = reinterpret_cast<uint32_t*>(smem + L.luma);
    uint32_t* s_acc = reinterpret_cast<uint32_t*>(smem + L.acc);
    uint8_t* s_hrow = smem + L.hrow;
    uint8_t* s_x32 = smem + L.x32;
    uint8_t* s_x98 = smem + L.x98;
    double* s_t = reinterpret_cast<double*>(smem + L.tmat);
    double* s_y = reinterpret_cast<double*>(smem + L.ymat);
    int* s_meta = reinterpret_cast<int*>(smem + L.meta);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L.bar);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int RC = a.rows_per_chunk;
    const int nparts = 32 / RC;
    const int row_l = lane & (RC - 1), part = lane / RC;
    const int n_chunks = (a.h + RC - 1) / RC;
    const long long row_bytes = (long long)a.w * C;

    for (int i = tid; i < a.coef_words; i += kThreads) s_coef[i] = a.coef[i];
    for (int i = tid; i < 2 * kOuts; i += kThreads) s_meta[i] = a.meta[i];
    for (int i = tid; i < RC * kOuts; i += kThreads) s_acc[i] = 1u << (kPrec - 1);
    if (tid == 0) {
        mbar_init(s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity = 0;

    // Issue (or perform) the load of chunk `ck` of image `im`; returns the byte offset of the
    // chunk's first row inside s_raw.
    auto chunk_geometry = [&](long long im, int ck, long long& g_begin, int& bytes, int& lead) {
        const long long start = im * a.img_stride + (long long)ck * RC * a.row_stride;
        const int rows = min(RC, a.h - ck * RC);
        const long long end = start + (long long)rows * row_bytes;
        // 16-byte aligned superset relative to the (16-byte aligned) base pointer
        const long long a0 = start & ~15ll;
        long long a1 = (end + 15ll) & ~15ll;
        if (a1 > a.total_bytes) a1 = a.total_bytes;  // never read past the caller's buffer
        g_begin = a0;
        bytes = (int)(a1 - a0);
        lead = (int)(start - a0);
    };
    auto issue_load = [&](long long im, int ck) {
        long long g0;
        int bytes, lead;
        chunk_geometry(im, ck, g0, bytes, lead);
        if (a.use_bulk && (bytes & 15) == 0) {
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(s_bar, (uint32_t)bytes);
                bulk_g2s(s_raw, a.img + g0, (uint32_t)bytes, s_bar);
            }
        } else {
            // generic path (strided rows or an unaligned tail): plain loads, then arrive
            const int rows = min(RC, a.h - ck * RC);
            const long long start = im * a.img_stride + (long long)ck * RC * a.row_stride;
            for (long long idx = tid; idx < rows * row_bytes; idx += kThreads) {
                const long long r = idx / row_bytes, x = idx - r * row_bytes;
                s_raw[lead + idx] = a.img[start + r * a.row_stride + x];
            }
            __syncthreads();
            if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(s_bar)) : "memory");
        }
        return lead;
    };

    long long im = blockIdx.x;
    int lead = 0;
    if (im < a.n) lead = issue_load(im, 0);

    for (; im < a.n; im += gridDim.x) {
        int32_t vacc[4], dacc = 1 << (kPrec - 1);
#pragma unroll
        for (int q = 0; q < 4; ++q) vacc[q] = 1 << (kPrec - 1);
        const int vx = tid & 31, vg = tid >> 5;

        for (int ck = 0; ck < n_chunks; ++ck) {
            const int r0 = ck * RC;
            const int rows = min(RC, a.h - r0);
            mbar_wait(s_bar, parity);
            parity ^= 1u;
            // vector fast path needs word (RGB) / 16-byte (RGBA) aligned rows in shared memory
            const bool aligned = C == 3 ? (((lead & 3) == 0) && ((row_bytes & 3) == 0))
                                        : (C == 4 ? (((lead & 15) == 0) && ((a.w & 3) == 0)) : false);
            luma_chunk<C>(s_raw + lead, s_luma, rows, a.w, a.pitch_words, aligned);
            __syncthreads();  // luma complete, s_raw free

            // prefetch the next chunk (of this or the next image) behind the horizontal pass
            {
                long long nim = im;
                int nck = ck + 1;
                if (nck == n_chunks) {
                    nim = im + gridDim.x;
                    nck = 0;
                }
                if (nim < a.n) lead = issue_load(nim, nck);
            }

            // ---- horizontal taps: lane -> (row, part); items dealt round-robin to warps
            for (int it = warp; it < a.n_items; it += kWarps) {
                const int4 item = a.items[it];
                const int o = item.x;
                const uint4* cf = s_coef + s_meta[kOuts + o] + item.y;
                const uint32_t* px = s_luma + row_l * a.pitch_words + s_meta[o] + item.y;
                uint32_t d0 = 0, d1 = 0;
                int32_t d2 = 0;
#pragma unroll 4
                for (int t = part; t < item.z; t += nparts) {
                    const uint4 cw = cf[t];
                    const uint32_t p = px[t];
                    d0 = dp4a_uu(p, cw.x, d0);
                    d1 = dp4a_uu(p, cw.y, d1);
                    d2 = dp4a_us(p, cw.z, d2);
                }
                if (row_l < rows) atomicAdd(&s_acc[row_l * kOuts + o], d0 + (d1 << 8) + ((uint32_t)d2 << 16));
            }
            __syncthreads();

            // ---- round, clip, reset accumulators
            for (int i = tid; i < rows * kOuts; i += kThreads) {
                s_hrow[i] = clip8((int32_t)s_acc[i]);
                s_acc[i] = 1u << (kPrec - 1);
            }
            __syncthreads();

            // ---- vertical taps, streamed: each thread owns outputs (yy = vg*4+q, x = vx)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int yy = vg * 4 + q;
                const int ymin = __ldg(a.b32 + 2 * yy), ylen = __ldg(a.b32 + 2 * yy + 1);
                const int lo = max(ymin, r0), hi = min(ymin + ylen, r0 + rows);
                const int* kk = a.kk32 + yy * a.ks32 - ymin;
                for (int y = lo; y < hi; ++y) vacc[q] += (int32_t)s_hrow[(y - r0) * kOuts + vx] * __ldg(kk + y);
            }
            if (tid < kDW * kDH) {
                const int yy = tid / kDW, x = tid - yy * kDW;
                const int ymin = __ldg(a.b8 + 2 * yy), ylen = __ldg(a.b8 + 2 * yy + 1);
                const int lo = max(ymin, r0), hi = min(ymin + ylen, r0 + rows);
                const int* kk = a.kk8 + yy * a.ks8 - ymin;
                for (int y = lo; y < hi; ++y) dacc += (int32_t)s_hrow[(y - r0) * kOuts + kOutW + x] * __ldg(kk + y);
            }
            // s_hrow is rewritten only after the next chunk's two barriers
        }

        // ---- planes
#pragma unroll
        for (int q = 0; q < 4; ++q) s_x32[(vg * 4 + q) * 32 + vx] = clip8(vacc[q]);
        if (tid < kDW * kDH) s_x98[tid] = clip8(dacc);
        __syncthreads();
        if (a.plane32)
            for (int i = tid; i < 1024; i += kThreads) a.plane32[im * 1024 + i] = s_x32[i];
        if (a.plane98 && tid < kDW * kDH) a.plane98[im * 72 + tid] = s_x98[tid];

        // ---- DCT low block: T = C(8x32) * X(32x32);  Y = T * C^T (8x8)
        {
            const int k = tid >> 5, x = tid & 31;
            double s = 0.0;
#pragma unroll 8
            for (int nn = 0; nn < 32; ++nn) s = fma(c_dct[k * 32 + nn], (double)s_x32[nn * 32 + x], s);
            s_t[k * 32 + x] = s;
        }
        __syncthreads();
        if (tid < 64) {
            const int k = tid >> 3, l = tid & 7;
            double s = 0.0;
#pragma unroll 8
            for (int nn = 0; nn < 32; ++nn) s = fma(s_t[k * 32 + nn], c_dct[l * 32 + nn], s);
            s_y[tid] = s;
        }
        __syncthreads();
        if (warp == 0) {
            const double y0 = s_y[lane], y1 = s_y[lane + 32];
            double sum = (lane == 0 ? 0.0 : y0) + y1;  // DC is compared but not averaged
#pragma unroll
            for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
            const double mean = sum / 63.0;
            const uint32_t bhi = __ballot_sync(0xffffffffu, y0 > mean), blo = __ballot_sync(0xffffffffu, y1 > mean);
            double mg = fmin(fabs(y0 - mean), fabs(y1 - mean));
#pragma unroll
            for (int off = 16; off; off >>= 1) mg = fmin(mg, __shfl_xor_sync(0xffffffffu, mg, off));
            // dHash: bit (r*8+c) = plane[r][c+1] > plane[r][c], first = MSB
            const int r_a = lane >> 3, c_a = lane & 7;
            const uint32_t dhi = __ballot_sync(0xffffffffu, s_x98[r_a * 9 + c_a + 1] > s_x98[r_a * 9 + c_a]);
            const uint32_t dlo =
                __ballot_sync(0xffffffffu, s_x98[(r_a + 4) * 9 + c_a + 1] > s_x98[(r_a + 4) * 9 + c_a]);
            if (lane == 0) {
                a.phash[im] = ((uint64_t)__brev(bhi) << 32) | (uint64_t)__brev(blo);
                a.dhash[im] = ((uint64_t)__brev(dhi) << 32) | (uint64_t)__brev(dlo);
                if (a.min_margin) a.min_margin[im] = (float)mg;
            }
        }
        __syncthreads();
    }
}


template <int NW = kWarps>
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory"); }

// DCT low block + hash bits from s_x32 / s_x98 (all NW compute warps call this)
template <int NW>
__device__ __forceinline__ void dct_and_bits(const PhashArgs& a, long long im, const uint8_t* s_x32, const uint8_t* s_x98,
                                             double* s_t, double* s_y, int tid, int lane, int warp) {
    if (a.plane32)
        for (int i = tid; i < 1024; i += NW * 32) a.plane32[im * 1024 + i] = s_x32[i];
    if (a.plane98 && tid < kDW * kDH) a.plane98[im * 72 + tid] = s_x98[tid];
    for (int idx = tid; idx < 256; idx += NW * 32) {
        const int k = idx >> 5, x = idx & 31;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;  // four independent chains: the DFMA latency, not its rate, bounds this
#pragma unroll
        for (int nn = 0; nn < 32; nn += 4) {
            s0 = fma(c_dct[k * 32 + nn], (double)s_x32[nn * 32 + x], s0);
            s1 = fma(c_dct[k * 32 + nn + 1], (double)s_x32[(nn + 1) * 32 + x], s1);
            s2 = fma(c_dct[k * 32 + nn + 2], (double)s_x32[(nn + 2) * 32 + x], s2);
            s3 = fma(c_dct[k * 32 + nn + 3], (double)s_x32[(nn + 3) * 32 + x], s3);
        }
        s_t[k * 32 + x] = (s0 + s1) + (s2 + s3);
    }
    compute_sync<NW>();
    if (tid < 64) {
        const int k = tid >> 3, l = tid & 7;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int nn = 0; nn < 32; nn += 4) {
            s0 = fma(s_t[k * 32 + nn], c_dct[l * 32 + nn], s0);
            s1 = fma(s_t[k * 32 + nn + 1], c_dct[l * 32 + nn + 1], s1);
            s2 = fma(s_t[k * 32 + nn + 2], c_dct[l * 32 + nn + 2], s2);
            s3 = fma(s_t[k * 32 + nn + 3], c_dct[l * 32 + nn + 3], s3);
        }
        s_y[tid] = (s0 + s1) + (s2 + s3);
    }
    compute_sync<NW>();
    if (warp == 0) {
        const double y0 = s_y[lane], y1 = s_y[lane + 32];
        double sum = (lane == 0 ? 0.0 : y0) + y1;
#pragma unroll
        for (int off = 16; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        const double mean = sum / 63.0;
        const uint32_t bhi = __ballot_sync(0xffffffffu, y0 > mean), blo = __ballot_sync(0xffffffffu, y1 > mean);
        double mg = fmin(fabs(y0 - mean), fabs(y1 - mean));
#pragma unroll
        for (int off = 16; off; off >>= 1) mg = fmin(mg, __shfl_xor_sync(0xffffffffu, mg, off));
        const int r_a = lane >> 3, c_a = lane & 7;
        const uint32_t dhi = __ballot_sync(0xffffffffu, s_x98[r_a * 9 + c_a + 1] > s_x98[r_a * 9 + c_a]);
        const uint32_t dlo = __ballot_sync(0xffffffffu, s_x98[(r_a + 4) * 9 + c_a + 1] > s_x98[(r_a + 4) * 9 + c_a]);
        if (lane == 0) {
            a.phash[im] = ((uint64_t)__brev(bhi) << 32) | (uint64_t)__brev(blo);
            a.dhash[im] = ((uint64_t)__brev(dhi) << 32) | (uint64_t)__brev(dlo);
            if (a.min_margin) a.min_margin[im] = (float)mg;
        }
    }
}

constexpr int kMaxSlots = 8;

// luma of `rows` raw rows (row r at raw + r*row_bytes) -> luma rows (word pitch `pitch_words`).
// warp -> rows, lanes -> 4-pixel groups: conflict-free LDS.32 x3 / STS.32 x1, no divisions.
template <int C, int NW = kWarps>
__device__ __forceinline__ void luma_rows_fast(const uint8_t* __restrict__ raw, uint32_t* __restrict__ luma, int rows,
                                               int w, int pitch_words, int warp, int lane) {
    constexpr uint32_t LO = 0x002F468Bu, HI = 0x001D964Cu;
    const int wq = w >> 2;
    for (int r = warp; r < rows; r += NW) {
        uint32_t* dst = luma + r * pitch_words;
        if (C == 1) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(raw + (size_t)r * w);
            for (int q = lane; q < wq; q += 32) dst[q] = src[q];
        } else if (C == 3) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(raw + (size_t)r * w * 3);
#pragma unroll 4
            for (int q = lane; q < wq; q += 32) {
                const uint32_t w0 = src[3 * q], w1 = src[3 * q + 1], w2 = src[3 * q + 2];
                const uint32_t l0 = dp4a_uu(w0, LO, 0x8000u), h0 = dp4a_uu(w0, HI, 0u);
                uint32_t l1 = dp4a_uu(w0, LO << 24, 0x8000u), h1 = dp4a_uu(w0, HI << 24, 0u);
                l1 = dp4a_uu(w1, LO >> 8, l1), h1 = dp4a_uu(w1, HI >> 8, h1);
                uint32_t l2 = dp4a_uu(w1, LO << 16, 0x8000u), h2 = dp4a_uu(w1, HI << 16, 0u);
                l2 = dp4a_uu(w2, LO >> 16, l2), h2 = dp4a_uu(w2, HI >> 16, h2);
                const uint32_t l3 = dp4a_uu(w2, LO << 8, 0x8000u), h3 = dp4a_uu(w2, HI << 8, 0u);
                const uint32_t s0 = l0 + (h0 << 8), s1 = l1 + (h1 << 8), s2 = l2 + (h2 << 8), s3 = l3 + (h3 << 8);
                dst[q] = __byte_perm(__byte_perm(s0, s1, 0x0062), __byte_perm(s2, s3, 0x0062), 0x5410);
            }
        } else {
            const uint4* src = reinterpret_cast<const uint4*>(raw + (size_t)r * w * 4);
            for (int q = lane; q < wq; q += 32) {
                const uint4 px = src[q];
                const uint32_t s0 = dp4a_uu(px.x, LO, 0x8000u) + (dp4a_uu(px.x, HI, 0u) << 8);
                const uint32_t s1 = dp4a_uu(px.y, LO, 0x8000u) + (dp4a_uu(px.y, HI, 0u) << 8);
                const uint32_t s2 = dp4a_uu(px.z, LO, 0x8000u) + (dp4a_uu(px.z, HI, 0u) << 8);
                const uint32_t s3 = dp4a_uu(px.w, LO, 0x8000u) + (dp4a_uu(px.w, HI, 0u) << 8);
                dst[q] = __byte_perm(__byte_perm(s0, s1, 0x0062), __byte_perm(s2, s3, 0x0062), 0x5410);
            }
        }
    }
}


// RGB rows -> luma rows, 16 pixels per lane: 3 LDS.128 of raw bytes (lane stride 48 B: conflict free), 48 dp4a,
// one STS.128.  Two rows are in flight per warp so that the loads of one overlap the arithmetic of the other.
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
// 4 RGB pixels (12 bytes in w0..w2) -> 4 luma bytes.  dp2a multiplies two 16-bit coefficients by a byte PAIR of
// the pixel word, and every pixel of the 12-byte group splits into pairs at (0,1)|(2,3) boundaries:
//   p0 = w0.b0..b2   p1 = w0.b3, w1.b0, w1.b1   p2 = w1.b2, w1.b3, w2.b0   p3 = w2.b1..b3
// so each pixel costs exactly two dp2a (Pillow: L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16).
__device__ __forceinline__ uint32_t luma4_rgb(uint32_t w0, uint32_t w1, uint32_t w2) {
    constexpr uint32_t cR = 19595u, cG = 38470u, cB = 7471u;
    constexpr uint32_t RG = cR | (cG << 16), B_ = cB, _R = cR << 16, GB = cG | (cB << 16);
    const uint32_t s0 = dp2a_hi(B_, w0, dp2a_lo(RG, w0, 0x8000u));
    const uint32_t s1 = dp2a_lo(GB, w1, dp2a_hi(_R, w0, 0x8000u));
    const uint32_t s2 = dp2a_lo(B_, w2, dp2a_hi(RG, w1, 0x8000u));
    const uint32_t s3 = dp2a_hi(GB, w2, dp2a_lo(_R, w2, 0x8000u));
    return __byte_perm(__byte_perm(s0, s1, 0x0062), __byte_perm(s2, s3, 0x0062), 0x5410);  // byte 2 of each sum
}
__device__ __forceinline__ uint32_t luma4_rgba(const uint4 px) {  // 4 RGBA pixels -> 4 luma bytes (alpha ignored, as Pillow does)
    constexpr uint32_t RG = 19595u | (38470u << 16), B_ = 7471u;
    const uint32_t s0 = dp2a_hi(B_, px.x, dp2a_lo(RG, px.x, 0x8000u)), s1 = dp2a_hi(B_, px.y, dp2a_lo(RG, px.y, 0x8000u));
    const uint32_t s2 = dp2a_hi(B_, px.z, dp2a_lo(RG, px.z, 0x8000u)), s3 = dp2a_hi(B_, px.w, dp2a_lo(RG, px.w, 0x8000u));
    return __byte_perm(__byte_perm(s0, s1, 0x0062), __byte_perm(s2, s3, 0x0062), 0x5410);
}
__device__ __forceinline__ uint4 luma16_rgb(const uint4 a, const uint4 b, const uint4 c) {
    return make_uint4(luma4_rgb(a.x, a.y, a.z), luma4_rgb(a.w, b.x, b.y), luma4_rgb(b.z, b.w, c.x), luma4_rgb(c.y, c.z, c.w));
}
template <int NLW>
__device__ __forceinline__ void luma_rows_rgb16(const uint8_t* __restrict__ raw, uint8_t* __restrict__ luma, int rows,
                                                int w, int row_bytes, int pitch_bytes, int lw, int lane) {
    const int ng = w >> 4;  // 16-pixel groups per row
    for (int r = lw; r < rows; r += 2 * NLW) {
        const bool two = r + NLW < rows;
        const uint8_t* s0 = raw + r * row_bytes;
        const uint8_t* s1 = s0 + (two ? NLW * row_bytes : 0);
        uint8_t* d0 = luma + r * pitch_bytes;
        uint8_t* d1 = d0 + NLW * pitch_bytes;
        for (int g = lane; g < ng; g += 32) {
            const uint4* p0 = reinterpret_cast<const uint4*>(s0 + 48 * g);
            const uint4* p1 = reinterpret_cast<const uint4*>(s1 + 48 * g);
            const uint4 a0 = p0[0], b0 = p0[1], c0 = p0[2];
            const uint4 a1 = p1[0], b1 = p1[1], c1 = p1[2];
            *reinterpret_cast<uint4*>(d0 + 16 * g) = luma16_rgb(a0, b0, c0);
            if (two) *reinterpret_cast<uint4*>(d1 + 16 * g) = luma16_rgb(a1, b1, c1);
        }
    }
}

__device__ __forceinline__ void mbar_arrive1(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------ tensor-core kernel (v5)
//
// Both resamples ARE matrix products (rows x banded tap matrix), and with the 22-bit taps in balanced base-256
// digits they are exact u8 x s8 -> s32 ones: v5 keeps v4's rings (1-D TMA raw slots -> luma warps -> 2 x 32-row
// luma ring) and runs them on the tensor pipe, which otherwise idles (mma.sync.m16n8k32, IMMA.16832).
//     warps 0..3   "wide target": warp q owns outputs 8q..8q+7 of the 32-wide plane END TO END.  Per 32-row chunk:
//                  A fragments by ldmatrix.x4 from the luma ring (pitch = w rounded to 32, + 16: conflict free), B
//                  fragments (3 digits x <= 8 k-steps, constant per width) in registers, epilogue
//                  (2^21 + d0 + 256 d1 + 65536 d2) >> 22 -> cvt.pack.sat -> its eight columns TRANSPOSED into a
//                  private scratch -> read back as the B fragment of the vertical pass (A = vertical tap digits of
//                  this chunk, accumulators for both 16-row halves of the 32x32 plane in registers all image long).
//                  No barrier with any other warp before the image is finished.
//     warps 4..7   "narrow target": output pairs of the 9-wide plane (B fragments in shared memory), the same way end to
//                  end: the pair's columns through a private scratch into the vertical pass of the 8x9 plane (an
//                  8-column tile with 2-3 live columns: three cheap MMAs per chunk buy the absence of any barrier).
//     warps 8..11  luma warps: wait for a 16-row raw slot, 16 pixels per lane through registers (2 dp2a per pixel),
//                  the LAST reader of a slot (shared-memory counter) issues the next TMA copy into it.
// Image end: all eight tap warps meet for the FP64 DCT and the bits.  setmaxnreg gives the three 4-warp groups
// 104 / 72 / 64 registers out of the 80 per thread the CTA launches with.

constexpr int kV5Tap = 8, kV5Luma = 4;
// Registers per role (setmaxnreg per 4-warp group; the CTA launches with 80 per thread): the wide-target warps hold their
// B fragments and the vertical accumulators in registers and take what the other two groups hand back.
constexpr int kV5WideRegs = 104, kV5NarrowRegs = 72, kV5LumaRegs = 64;
// one CTA per SM (168 per thread at launch): wide warps holding 16 / 32 k-steps of fragments (96 / 192 registers)
// and narrow warps holding their quarter of the row: 9 / 18 k-steps x 4 tiles (72 / 144 registers)
constexpr int kV5WideRegs16 = 200, kV5WideRegs32 = 232, kV5NarrowRegs16 = 144, kV5NarrowRegs32 = 200;
static_assert(kV5WideRegs32 + kV5NarrowRegs32 + kV5LumaRegs <= 3 * 168 && kV5WideRegs16 + kV5NarrowRegs16 + kV5LumaRegs <= 3 * 168,
              "register pool of the 3 warp groups, one CTA per SM");
static_assert(kV5WideRegs + kV5NarrowRegs + kV5LumaRegs == 3 * 80, "register pool of the 3 warp groups");
constexpr int kV5Threads = (kV5Tap + kV5Luma) * 32;
constexpr int kHP = 48;   // horizontal-pass output plane, stored TRANSPOSED: [column 0..47][row 0..31], column pitch 48 B
constexpr int kHCols = 48;

constexpr int kMaxLumaBufs = 4;
// narrow-target partial sums: [2 buffers][4 source warps][9 outputs][32 rows (stride 40: outputs 8 banks apart)] int32
constexpr int kPartOut = 40, kPartWords = 2 * 4 * kDW * kPartOut;
// two-CTA kernel: the upper half-band warp of a group hands 12 accumulator slots per lane to the lower one:
// [2 groups][2 buffers][12 slots][32 lanes] int32
constexpr int kHalfSlots = 12, kHalfWords = 2 * 2 * kHalfSlots * 32;
__host__ __device__ constexpr int v5_part_bytes(int nkw) { return nkw > 8 ? kPartWords * 4 : nkw == 8 ? kHalfWords * 4 : 0; }

struct V5Layout {
    int raw, luma, bfrag, hrow, part, x32, x98, tmat, ymat, bar, luma_bytes, total;
};

__host__ __device__ inline V5Layout v5_layout(int slot_bytes, int pitch_bytes, int n_slots, int mma_words /* in shared memory */,
                                                  int nlb /* luma chunk buffers */, int chunk_rows = 32,
                                                  int part_bytes = 0 /* narrow-target partial sums */) {
    V5Layout L;
    int off = 0;
    auto take = [&](int bytes, int align) {
        off = (off + align - 1) / align * align;
        int at = off;
        off += bytes;
        return at;
    };
    L.luma_bytes = (chunk_rows * pitch_bytes + 127) / 128 * 128;
    L.raw = take(n_slots * slot_bytes, 128);
    L.luma = take(nlb * L.luma_bytes, 128);
    L.bfrag = take(mma_words * 8, 16);
    L.hrow = take(2 * kHCols * kHP, 16);
    L.x32 = take(1024, 16);
    L.x98 = take(80, 16);
    L.tmat = take(8 * 32 * 8, 16);
    L.ymat = take(64 * 8, 16);
    L.bar = take((2 * kMaxSlots + 2 * kMaxLumaBufs) * 8 + kMaxSlots * 4, 8);
    // last: with this region in the middle of the layout the two-CTA kernel lost 2.5 % (9.25 -> 9.46 ms per 70 000 images,
    // A/B on one box) although it never touches it
    L.part = take(part_bytes, 16);
    L.total = off;
    return L;
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

// One 32-row luma chunk -> this warp's columns of the row plane.  NT tiles share the k range.
// WIDE: tiles are the three digits of outputs out0..out0+7; else tile tl is the output pair (out0 + 2 tl, +1).
template <int NT, bool WIDE, int NRB = 2>
__device__ __forceinline__ void v5_taps(uint32_t a_addr, const uint2* __restrict__ bw, int nk, int pitch_bytes,
                                        uint8_t* __restrict__ hrow, int out0, int lane, int row_off = 0) {
    int32_t c[NRB][NT][4];
#pragma unroll
    for (int rb = 0; rb < NRB; ++rb)
#pragma unroll
        for (int tl = 0; tl < NT; ++tl)
#pragma unroll
            for (int i = 0; i < 4; ++i)  // the rounding term 2^21 rides in the d0 accumulators
                c[rb][tl][i] = WIDE ? (tl == 0 ? (1 << (kPrec - 1)) : 0) : ((i & 1) == 0 && (lane & 3) < 2 ? (1 << (kPrec - 1)) : 0);
#pragma unroll 2
    for (int k = 0; k < nk; ++k) {
        uint32_t a0[4], a1[4];
        ldmatrix_x4(a0, a_addr + k * 32);
        if (NRB == 2) ldmatrix_x4(a1, a_addr + k * 32 + 16 * pitch_bytes);
#pragma unroll
        for (int tl = 0; tl < NT; ++tl) {
            const uint2 b = bw[(k * NT + tl) * 32];
            mma_u8s8(c[0][tl], a0, b);
            if (NRB == 2) mma_u8s8(c[NRB - 1][tl], a1, b);
        }
    }
    const int g = lane >> 2, t = lane & 3;
    {
        const int src = (lane & ~3) | 2;  // the quad's lane holding the third digits {a.d2, b.d2}
#pragma unroll
        for (int rb = 0; rb < NRB; ++rb)
#pragma unroll
            for (int tl = 0; tl < NT; ++tl)
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const int32_t xa = __shfl_sync(0xffffffffu, c[rb][tl][2 * hf], src);
                    const int32_t xb = __shfl_sync(0xffffffffu, c[rb][tl][2 * hf + 1], src);
                    const int32_t d2 = t == 0 ? xa : xb;
                    const int32_t v = c[rb][tl][2 * hf] + (c[rb][tl][2 * hf + 1] << 8) + (d2 << 16);
                    if (t < 2) hrow[(out0 + 2 * tl + t) * kHP + row_off + rb * 16 + hf * 8 + g] = (uint8_t)pack_sat_u8(0, v >> kPrec);
                }
    }
}


// Wide-target warps keep their B fragments (3 digits x <= kNKP k-steps) in REGISTERS for the whole kernel: no
// shared-memory table, no LDS per MMA.  One 16-row block at a time keeps the accumulators at 12 registers.
constexpr int kNKP = 8;
template <int NKW, int NRB>
__device__ __forceinline__ void v5_taps_wide_reg(uint32_t a_addr, const uint2 (&breg)[NKW][3], int nk, int pitch_bytes,
                                                 uint8_t* __restrict__ hrow, int out0, int lane, int row_off) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int rb = 0; rb < NRB; ++rb) {
        int32_t c[3][4];
#pragma unroll
        for (int tl = 0; tl < 3; ++tl)
#pragma unroll
            for (int i = 0; i < 4; ++i) c[tl][i] = tl == 0 ? (1 << (kPrec - 1)) : 0;
#pragma unroll
        for (int k = 0; k < NKW; ++k) {
            if (k < nk) {
                uint32_t a0[4];
                ldmatrix_x4(a0, a_addr + k * 32 + rb * 16 * pitch_bytes);
#pragma unroll
                for (int tl = 0; tl < 3; ++tl) mma_u8s8(c[tl], a0, breg[k][tl]);
            }
        }
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const int32_t v0 = c[0][2 * hf] + (c[1][2 * hf] << 8) + (c[2][2 * hf] << 16);
            const int32_t v1 = c[0][2 * hf + 1] + (c[1][2 * hf + 1] << 8) + (c[2][2 * hf + 1] << 16);
            const int row = row_off + rb * 16 + hf * 8 + g;
            hrow[(out0 + 2 * t) * kHP + row] = (uint8_t)pack_sat_u8(0, v0 >> kPrec);
            hrow[(out0 + 2 * t + 1) * kHP + row] = (uint8_t)pack_sat_u8(0, v1 >> kPrec);
        }
    }
}

// Where the horizontal B fragments of a warp group live: registers (wide target only, <= kNKP k-steps), shared memory,
// or global memory (L2-resident, for bands too long for shared memory next to the luma ring).
enum { kBReg = 0, kBSmem = 1, kBGmem = 2 };

struct V5Config {
    int sub_rows, slot_shift, pitch_bytes, nlb;
    int cr;          // rows per luma ring buffer: 32, or 16 (pointer-fed tap loops only)
    int slot_bytes;  // stride of a raw slot: sub_rows * row_bytes, + slack for the aligned superset when !aligned
    int aligned;     // base, images and rows on 16-byte boundaries and w % 16 == 0: exact copies, vector luma path
    int wide_b, narrow_b;  // kBReg / kBSmem / kBGmem
    int nkw;         // k-steps of wide-target fragments a warp keeps in registers: 8 (two CTAs per SM), 16 / 32 (one), 0: none
    int smem_words;  // uint2 words of B fragments kept in shared memory
    int dbg;
    V5Layout L;
};

// Wide-target taps with the B fragments behind a pointer (shared or global memory): bands of more than kNKP k-steps.
template <int NRB>
__device__ __forceinline__ void v5_taps_wide_mem(uint32_t a_addr, const uint2* __restrict__ bw, int nk, int pitch_bytes,
                                                 uint8_t* __restrict__ hrow, int out0, int lane, int row_off) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int rb = 0; rb < NRB; ++rb) {
        int32_t c[3][4];
#pragma unroll
        for (int tl = 0; tl < 3; ++tl)
#pragma unroll
            for (int i = 0; i < 4; ++i) c[tl][i] = tl == 0 ? (1 << (kPrec - 1)) : 0;
#pragma unroll 4
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
            const int row = row_off + rb * 16 + hf * 8 + g;
            hrow[(out0 + 2 * t) * kHP + row] = (uint8_t)pack_sat_u8(0, v0 >> kPrec);
            hrow[(out0 + 2 * t + 1) * kHP + row] = (uint8_t)pack_sat_u8(0, v1 >> kPrec);
        }
    }
}

// Luma of the rows lw, lw + NLW, ... of a raw sub-chunk that starts `off` bytes into a 16-byte aligned slot, for ANY
// byte alignment of the rows (widths that are not a multiple of 16, odd image strides, sliced batches): aligned 32-bit
// loads + funnel shifts.  RGB / L: a lane takes 4 pixels per step (word stride 3 / 1 between lanes: no bank conflicts);
// RGBA: one pixel per lane.  Pixels past the row end are converted too (they land in the ring's padding, where every tap
// is zero).
template <int C, int NLW>
__device__ __noinline__ void luma_rows_any(const uint8_t* __restrict__ slot, int off, uint8_t* __restrict__ luma, int rows,
                                              int w, int row_bytes, int pitch_bytes, int lw, int lane) {
    for (int r = lw; r < rows; r += NLW) {
        const int o = off + r * row_bytes;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(slot + (o & ~3));
        const uint32_t sh = (uint32_t)(o & 3) * 8u;
        uint32_t* dst = reinterpret_cast<uint32_t*>(luma + r * pitch_bytes);
        if (C == 3) {
            const int nq = (w + 3) >> 2;
            for (int q = lane; q < nq; q += 32) {
                const uint32_t* p = src + 3 * q;
                const uint32_t x0 = p[0], x1 = p[1], x2 = p[2], x3 = p[3];
                dst[q] = luma4_rgb(__funnelshift_r(x0, x1, sh), __funnelshift_r(x1, x2, sh), __funnelshift_r(x2, x3, sh));
            }
        } else if (C == 4) {
            constexpr uint32_t RG = 19595u | (38470u << 16), B_ = 7471u;
            for (int x = lane; x < w; x += 32) {
                const uint32_t px = __funnelshift_r(src[x], src[x + 1], sh);
                luma[r * pitch_bytes + x] = (uint8_t)(dp2a_hi(B_, px, dp2a_lo(RG, px, 0x8000u)) >> 16);
            }
        } else {
            const int nq = (w + 3) >> 2;
            for (int q = lane; q < nq; q += 32) dst[q] = __funnelshift_r(src[q], src[q + 1], sh);
        }
    }
}

// NKW = k-steps of wide-target resample fragments each wide warp keeps in REGISTERS:
//   8   bands of <= 8 k-steps (widths up to ~512: the benchmark geometry); 80 registers per thread, two CTAs per SM;
//       narrow-target fragments in shared memory, the narrow warps work in two output groups x two band halves.
//   16 / 32   longer rows (up to ~1100 / ~2200 pixels): ONE CTA per SM with 168 registers per thread at launch, which
//       setmaxnreg redistributes — the wide warps take 200 / 232 and hold their whole band, the narrow warps take
//       144 / 200 and hold a quarter of the row's k-steps for all nine outputs (K-split).  Shared memory was the bound
//       there (ncu: the LSU data pipe at 60 % + the bulk-copy writes at 0.70 of HBM, a third of it fragment reads, a
//       quarter the four narrow warps each reading nearly the whole row); nothing is left in shared memory but the rings.
//   0   bands beyond that: every fragment behind a pointer into shared or global memory (L2), narrow warps own output
//       pairs end to end.
// ALIGNED = true: base, images and rows on 16-byte boundaries and w % 16 == 0 (exact copies, vector luma path).
// The staging geometry comes as SCALAR kernel parameters and the shared-memory layout is recomputed in the kernel: with
// the same values read from a parameter struct the compiler kept them off the uniform datapath and the 512x512 RGB case
// lost 15 % (measured: 10.7 ms against 9.2 ms per 70 000 images).
// CR = rows per luma ring buffer: 32 (one vertical k-step per buffer), or 16 for long rows — half the ring and a finer
// hand-off between luma and tap warps; the tap warps then run the vertical pass after every second buffer.
template <int C, int NKW, bool ALIGNED, int CR = 32>
__global__ void __launch_bounds__(kV5Threads, NKW > kNKP ? 1 : 2)
ke_phash_v5_kernel(const PhashArgs a, const int sub_rows, const int slot_shift, const int pitch_bytes, const int nlb,
                   const int dbg, const V5Config cfg) {
    constexpr int NW = kV5Tap, NRB = CR / 16;
    static_assert(NKW == 0 || NKW == kNKP || NKW == 16 || NKW == 32, "register-resident bands of 8, 16 or 32 k-steps");
    static_assert(CR == 32 || (CR == 16 && NKW != kNKP), "the two-CTA kernel runs 32-row luma buffers only");
    extern __shared__ __align__(128) uint8_t smem[];
    const int row_bytes = a.w * C;
    const int sub_bytes = sub_rows * row_bytes;
    const int slot_bytes = ALIGNED ? sub_bytes : cfg.slot_bytes;  // aligned rows: slots are exactly one sub-chunk apart
    const int n_slots = 1 << slot_shift;
    const uint32_t slot_mask = (uint32_t)n_slots - 1u;
    // B fragments in shared memory: [wide-target warps 0..3 when cfg.wide_b == kBSmem][narrow-target warps 4..7 when
    // cfg.narrow_b == kBSmem], in table order
    const int b_first = NKW == kNKP ? a.mmah_boff[0] : (NKW == 0 && cfg.wide_b == kBSmem) ? 0 : a.mma_boff[4];
    // (NKW == 8 must not read cfg here: a value selected through the parameter struct leaves the uniform datapath and
    // every shared-memory address derived from the layout with it — measured 9.20 -> 9.44 ms per 70 000 images)
    const int b_last = NKW > kNKP ? b_first : NKW == kNKP ? a.mmah_end : (cfg.narrow_b != kBSmem ? a.mma_boff[4] : a.mma_words);
    const V5Layout L = v5_layout(slot_bytes, pitch_bytes, n_slots, b_last - b_first, nlb, CR, v5_part_bytes(NKW));
    uint8_t* s_raw = smem + L.raw;
    uint8_t* s_luma = smem + L.luma;
    uint2* s_b = reinterpret_cast<uint2*>(smem + L.bfrag);
    uint8_t* s_hrow = smem + L.hrow;
    int32_t* s_part = reinterpret_cast<int32_t*>(smem + L.part);
    uint8_t* s_x32 = smem + L.x32;
    uint8_t* s_x98 = smem + L.x98;
    double* s_t = reinterpret_cast<double*>(smem + L.tmat);
    double* s_y = reinterpret_cast<double*>(smem + L.ymat);
    uint64_t* s_full = reinterpret_cast<uint64_t*>(smem + L.bar);  // raw ring
    uint64_t* s_empty = s_full + kMaxSlots;
    uint64_t* l_full = s_empty + kMaxSlots;  // luma chunk ring [nlb]
    uint64_t* l_empty = l_full + kMaxLumaBufs;
    uint32_t* s_cnt = reinterpret_cast<uint32_t*>(l_empty + kMaxLumaBufs);  // readers done with a raw slot (mod kV5Luma)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_sub = (a.h + sub_rows - 1) / sub_rows;
    const int subs_per_chunk = CR / sub_rows;
    const int pitch_words = pitch_bytes >> 2;

    if (tid == 0) {
        for (int b = 0; b < n_slots; ++b) {
            mbar_init(&s_full[b], 1);
            s_cnt[b] = 0u;
        }
        for (int b = 0; b < nlb; ++b) {
            mbar_init(&l_full[b], kV5Luma);
            mbar_init(&l_empty[b], kV5Tap);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < b_last - b_first; i += kV5Threads) s_b[i] = __ldg(a.mma_b + b_first + i);
    for (int i = tid; i < nlb * L.luma_bytes / 4; i += kV5Threads) reinterpret_cast<uint32_t*>(s_luma)[i] = 0u;
    for (int i = tid; i < 2 * kHCols * kHP / 4; i += kV5Threads) reinterpret_cast<uint32_t*>(s_hrow)[i] = 0u;
    __syncthreads();

    if (warp >= kV5Tap) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kV5LumaRegs));  // (from 80, or 168 with one CTA per SM)
        // ===== luma warps: raw rows -> (TMA) raw ring -> luma chunk ring =====
        // A raw slot is refilled by whichever luma warp finishes reading it LAST (a shared-memory counter
        // per slot tells): no issuer warp, no "slot empty" barrier to wait on, the copy of sub-chunk
        // seq + n_slots leaves the moment slot seq % n_slots is free.
        const int lw = warp - kV5Tap;
        const long long my_images = blockIdx.x < a.n ? (a.n - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const uint32_t total_seq = (uint32_t)(my_images * n_sub);  // per-CTA sub-chunk count fits 32 bits (launch checks)
        auto issue = [&](uint32_t q) {  // one lane: copy sub-chunk q of this CTA's stream into slot q % n_slots
            const uint32_t k = q / (uint32_t)n_sub, sq = q - k * (uint32_t)n_sub;
            const int b = (int)(q & slot_mask);
            const int rows = min(sub_rows, a.h - (int)sq * sub_rows);
            const uint8_t* src = a.img + (blockIdx.x + (long long)k * gridDim.x) * a.img_stride + (long long)sq * sub_bytes;
            uint32_t bytes = (uint32_t)(rows * row_bytes);
            if constexpr (!ALIGNED) {  // the 16-byte aligned superset of the sub-chunk (bulk copies move whole 16-byte units)
                const uint32_t lead = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
                src -= lead;
                bytes = (lead + bytes + 15u) & ~15u;
            }
            mbar_expect_tx(&s_full[b], bytes);
            bulk_g2s(s_raw + b * slot_bytes, src, bytes, &s_full[b]);
        };
        auto release = [&](uint32_t seq_, int b) {  // lane 0, after the warp's reads of slot b (ordered by __syncwarp)
            uint32_t old;  // acq_rel: the warp's reads are ordered before the count, the refill after the last count
            asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(&s_cnt[b])) : "memory");
            if ((old & (kV5Luma - 1)) == kV5Luma - 1 && seq_ + (uint32_t)n_slots < total_seq) issue(seq_ + (uint32_t)n_slots);
        };
        if (lw == 0 && lane == 0)
            for (uint32_t q = 0; q < (uint32_t)n_slots && q < total_seq; ++q) issue(q);
        // early release: when a warp's share of a sub-chunk (<= 2 rows of <= 512 pixels) fits its registers the
        // raw slot is handed back right after the loads, before the arithmetic
        const bool early = ALIGNED && a.w <= 512 && !(dbg & 16);  // 16 pixels (RGB, L) or 4 x 4 pixels (RGBA) per lane cover a row
        const int ng = a.w >> 4;
        const bool act = lane < ng;
        uint32_t seq = 0;
        uint32_t chunk = 0, lph = 0;  // luma ring: buffer lb, phase lph
        int lb = 0;
        for (long long im = blockIdx.x; im < a.n; im += gridDim.x) {
            for (int r0 = 0; r0 < a.h; r0 += CR, ++chunk, lph ^= (lb + 1 == nlb), lb = lb + 1 == nlb ? 0 : lb + 1) {
                mbar_wait(&l_empty[lb], lph ^ 1u);  // tap warps are done with this buffer
                uint8_t* dst8 = s_luma + lb * L.luma_bytes;
                for (int s = 0; s < subs_per_chunk && r0 + s * sub_rows < a.h; ++s, ++seq) {
                    const int b = (int)(seq & slot_mask);
                    const int srows = min(sub_rows, a.h - (r0 + s * sub_rows));
                    mbar_wait(&s_full[b], (seq >> slot_shift) & 1u);
                    if (dbg & 8) {  // tuning probe: the TMA feed alone
                        __syncwarp();
                        if (lane == 0) release(seq, b);
                    } else if (early) {
                        // rows lw, lw+NL, lw+2NL, ... of the sub-chunk, two at a time through registers.  The slot is handed back
                        // after the warp's last row (measured: with 16-row sub-chunks handing it back before the arithmetic
                        // of the last pair costs more in the luma warps than the earlier refill gains)
                        for (int rr = lw; rr < srows; rr += 2 * kV5Luma) {
                            const bool r_two = rr + kV5Luma < srows;
                            const uint8_t* s0 = s_raw + b * slot_bytes + rr * row_bytes;
                            const uint8_t* s1 = s0 + kV5Luma * row_bytes;
                            uint8_t* d0 = dst8 + (s * sub_rows + rr) * pitch_bytes;
                            uint8_t* d1 = d0 + kV5Luma * pitch_bytes;
                            if (C == 3) {  // 16 pixels = 48 bytes per lane (lane stride 48 B: conflict-free LDS.128)
                                const bool one = act, two = act && r_two;
                                uint4 x0[3], x1[3];
#pragma unroll
                                for (int i = 0; i < 3; ++i) {
                                    x0[i] = one ? reinterpret_cast<const uint4*>(s0 + 48 * lane)[i] : make_uint4(0, 0, 0, 0);
                                    x1[i] = two ? reinterpret_cast<const uint4*>(s1 + 48 * lane)[i] : make_uint4(0, 0, 0, 0);
                                }
                                if (one) *reinterpret_cast<uint4*>(d0 + 16 * lane) = luma16_rgb(x0[0], x0[1], x0[2]);
                                if (two) *reinterpret_cast<uint4*>(d1 + 16 * lane) = luma16_rgb(x1[0], x1[1], x1[2]);
                            } else if (C == 4) {  // 4 pixels per 16-byte piece, piece 32 i + lane: consecutive lanes, no conflicts
                                const int pieces = a.w >> 2;
                                uint4 x0[4], x1[4];
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const bool in = 32 * i + lane < pieces;
                                    x0[i] = in ? reinterpret_cast<const uint4*>(s0)[32 * i + lane] : make_uint4(0, 0, 0, 0);
                                    x1[i] = in && r_two ? reinterpret_cast<const uint4*>(s1)[32 * i + lane] : make_uint4(0, 0, 0, 0);
                                }
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const bool in = 32 * i + lane < pieces;
                                    if (in) reinterpret_cast<uint32_t*>(d0)[32 * i + lane] = luma4_rgba(x0[i]);
                                    if (in && r_two) reinterpret_cast<uint32_t*>(d1)[32 * i + lane] = luma4_rgba(x1[i]);
                                }
                            } else {  // 'L': the rows are the luma rows already
                                const bool one = act, two = act && r_two;
                                const uint4 x0 = one ? reinterpret_cast<const uint4*>(s0)[lane] : make_uint4(0, 0, 0, 0);
                                const uint4 x1 = two ? reinterpret_cast<const uint4*>(s1)[lane] : make_uint4(0, 0, 0, 0);
                                if (one) reinterpret_cast<uint4*>(d0)[lane] = x0;
                                if (two) reinterpret_cast<uint4*>(d1)[lane] = x1;
                            }
                        }
                        __syncwarp();  // (a warp without rows in a short last sub-chunk still counts as a reader)
                        if (lane == 0) release(seq, b);
                    } else if constexpr (ALIGNED) {
                        if (C == 3)
                            luma_rows_rgb16<kV5Luma>(s_raw + b * slot_bytes, dst8 + s * sub_rows * pitch_bytes, srows, a.w,
                                                     row_bytes, pitch_bytes, lw, lane);
                        else
                            luma_rows_fast<C, kV5Luma>(s_raw + b * slot_bytes,
                                                       reinterpret_cast<uint32_t*>(dst8) + s * sub_rows * pitch_words, srows,
                                                       a.w, pitch_words, lw, lane);
                        __syncwarp();
                        if (lane == 0) release(seq, b);
                    } else {
                        // rows at any byte alignment: where the sub-chunk starts inside its slot follows from its address
                        const long long g0 = im * a.img_stride + (long long)(r0 + s * sub_rows) * row_bytes;
                        const int lead = (int)(reinterpret_cast<uintptr_t>(a.img + g0) & 15u);
                        luma_rows_any<C, kV5Luma>(s_raw + b * slot_bytes, lead, dst8 + s * sub_rows * pitch_bytes, srows, a.w,
                                                  row_bytes, pitch_bytes, lw, lane);
                        __syncwarp();
                        if (lane == 0) release(seq, b);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive1(&l_full[lb]);  // release: the chunk's luma rows are written
            }
        }
        return;
    }

    // ===== tap warps =====
#ifdef KE_TUNING_PROBES
    const int nk = ((dbg & 2) || ((dbg & 32) && warp < 4) || ((dbg & 64) && warp >= 4)) ? 0 : a.mma_nk[warp];
#else
    const int nk = (dbg & 2) ? 0 : a.mma_nk[warp];
#endif
    const uint32_t a_off = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * pitch_bytes + (lane >> 4) * 16 + a.mma_k0[warp] * 32);
    const uint32_t poll_ns = (uint32_t)dbg >> 8;
    const int g = lane >> 2, t = lane & 3;
    // Vertical pass, also on the tensor pipe: out[yy, x] = sum_y V[yy, y] * hrow[y, x]; a 32-row chunk is one k-step of
    // mma.m16n8k32.s8.u8 (A = tap digits from the per-height table, B = the transposed row plane).  Accumulators live
    // in registers for the whole image.
    if (warp < 4) {
        // ---- wide target: warp q owns outputs 8q..8q+7 END TO END — horizontal taps (B fragments in registers), then
        // the vertical pass of exactly those eight columns (both 16-row halves of the 32x32 plane).  It reads back only
        // what it wrote itself, so it never meets another tap warp before the image is finished.
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(NKW == 32 ? kV5WideRegs32 : NKW == 16 ? kV5WideRegs16 : kV5WideRegs));
        uint2 breg[NKW ? NKW : 1][3];
        if constexpr (NKW > 0) {
#pragma unroll
            for (int k = 0; k < NKW; ++k)
#pragma unroll
                for (int tl = 0; tl < 3; ++tl)
                    breg[k][tl] = k < a.mma_nk[warp] ? __ldg(a.mma_b + a.mma_boff[warp] + (k * 3 + tl) * 32 + lane) : make_uint2(0u, 0u);
        }
        // NKW == 0: bands of more than 32 k-steps — fragments from shared memory, or from global memory (L2) when they do
        // not fit next to the luma ring
        const uint2* bmem = (cfg.wide_b == kBSmem ? s_b + (a.mma_boff[warp] - b_first) : a.mma_b + a.mma_boff[warp]) + lane;
        uint8_t* hrow = s_hrow;  // columns 8q..8q+7 of plane 0 are this warp's private scratch
        const uint32_t* col = reinterpret_cast<const uint32_t*>(hrow + (8 * warp + g) * kHP);
        int32_t vc[2][3][4];
        uint32_t lph = 0;
        int lb = 0;
        for (long long im = blockIdx.x; im < a.n; im += gridDim.x) {
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int d = 0; d < 3; ++d)
#pragma unroll
                    for (int i = 0; i < 4; ++i) vc[u][d][i] = d == 0 ? (1 << (kPrec - 1)) : 0;
            int hc = 0;  // ring buffers consumed of this image; vertical k-step ci = rows / 32
            for (int r0 = 0; r0 < a.h; r0 += CR, ++hc, lph ^= (lb + 1 == nlb), lb = lb + 1 == nlb ? 0 : lb + 1) {
                mbar_wait_sleep(&l_full[lb], lph, poll_ns);
                const int row_off = CR == 32 ? 0 : (hc & 1) * 16;
                {
                    const uint32_t a_addr = smem_u32(s_luma + lb * L.luma_bytes) + a_off;
                    if constexpr (NKW > 0) v5_taps_wide_reg<NKW, NRB>(a_addr, breg, nk, pitch_bytes, hrow, 8 * warp, lane, row_off);
                    else v5_taps_wide_mem<NRB>(a_addr, bmem, nk, pitch_bytes, hrow, 8 * warp, lane, row_off);
                }
                __syncwarp();  // the eight columns are written
                if (lane == 0) mbar_arrive1(&l_empty[lb]);  // this warp no longer reads the luma buffer
                // 16-row buffers: the vertical k-step needs both halves of its 32 rows (or the image's last rows; whatever
                // an earlier chunk left in the other half meets zero taps there)
                const bool vstep = CR == 32 || (hc & 1) || r0 + CR >= a.h;
                const int ci = CR == 32 ? hc : hc >> 1;
                if (vstep && !(dbg & 4)) {
                    const uint32_t b0 = col[t], b1 = col[4 + t];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if (ci < a.v_lo[u] || ci > a.v_hi[u]) continue;
                        const uint4* af = a.vmma + ((size_t)(ci * 3 + u) * 3) * 32 + lane;
#pragma unroll
                        for (int d = 0; d < 3; ++d) mma_s8u8(vc[u][d], __ldg(af + d * 32), b0, b1);
                    }
                }
                __syncwarp();  // every lane has read the columns before the next chunk overwrites them
            }
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const int32_t v0 = vc[u][0][2 * hf] + (vc[u][1][2 * hf] << 8) + (vc[u][2][2 * hf] << 16);
                    const int32_t v1 = vc[u][0][2 * hf + 1] + (vc[u][1][2 * hf + 1] << 8) + (vc[u][2][2 * hf + 1] << 16);
                    *reinterpret_cast<uint16_t*>(s_x32 + (u * 16 + hf * 8 + g) * 32 + 8 * warp + 2 * t) =
                        (uint16_t)pack_sat_u8(v1 >> kPrec, v0 >> kPrec);
                }
            compute_sync<NW>();  // both planes are complete
            dct_and_bits<NW>(a, im, s_x32, s_x98, s_t, s_y, tid, lane, warp);
        }
        return;
    }

    // ---- narrow target (9 x 8 plane), warps 4..7.
    if constexpr (NKW == 32) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kV5NarrowRegs32));  // above the 168 of the launch
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(NKW == 16 ? kV5NarrowRegs16 : kV5NarrowRegs));
    uint8_t* scr = s_hrow + kHCols * kHP;                 // plane 1
    const int scr_col = 8 * (warp - 4);                   // this warp's private columns scr_col .. scr_col + 7
    const uint32_t* col = reinterpret_cast<const uint32_t*>(scr + (scr_col + g) * kHP);
    int32_t vc[3][4];
    uint32_t lph = 0;
    int lb = 0;
    if constexpr (NKW > kNKP) {
        // Horizontal pass split by K (one CTA per SM, where the registers for it exist): warp 4 + j multiplies ITS quarter of the row's k-steps (fragments in registers) into
        // all nine outputs — every luma byte is read by one narrow warp instead of four, and no fragment is re-read per
        // chunk (measured before: the narrow warps cost 0.09 of HBM at 512 pixels, 0.10 at 1024).  The quarter sums meet in
        // shared memory (double-buffered: one 128-thread barrier per chunk), warp 4 + j finishes the output pair
        // (2j, 2j + 1) (warp 7: 6, 7 and 8) — sum of four, round, clip — into its private columns and runs the
        // vertical pass of exactly those columns, as before.
        constexpr int NKN = NKW == 32 ? 18 : 9;
        const int j = warp - 4;
        const int nkq = ((dbg & 2) || (dbg & 64)) ? 0 : a.mmaq_nk[j];
        uint2 nreg[NKN][4];
#pragma unroll
        for (int k = 0; k < NKN; ++k)
#pragma unroll
            for (int T = 0; T < 4; ++T)
                nreg[k][T] = k < a.mmaq_nk[j] ? __ldg(a.mma_b + a.mmaq_boff[j] + (k * 4 + T) * 32 + lane) : make_uint2(0u, 0u);
        const uint32_t aq_off = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * pitch_bytes + (lane >> 4) * 16 + a.mmaq_k0[j] * 32);
        const int src = (lane & ~3) | 2;  // the quad's lane holding the third digits {a.d2, b.d2}
        const int n_own = j == 3 ? 3 : 2;
        int pb = 0;
        for (long long im = blockIdx.x; im < a.n; im += gridDim.x) {
#pragma unroll
            for (int d = 0; d < 3; ++d)
#pragma unroll
                for (int i = 0; i < 4; ++i) vc[d][i] = d == 0 ? (1 << (kPrec - 1)) : 0;
            int hc = 0;
            for (int r0 = 0; r0 < a.h; r0 += CR, ++hc, lph ^= (lb + 1 == nlb), lb = lb + 1 == nlb ? 0 : lb + 1, pb ^= 1) {
                mbar_wait_sleep(&l_full[lb], lph, poll_ns);
                const uint32_t a_addr = smem_u32(s_luma + lb * L.luma_bytes) + aq_off;
                const int row_off = CR == 32 ? 0 : (hc & 1) * 16;
                int32_t* mine = s_part + (pb * 4 + j) * (kDW * kPartOut);
#pragma unroll
                for (int rb = 0; rb < NRB; ++rb) {
                    int32_t c[4][4];
#pragma unroll
                    for (int T = 0; T < 4; ++T)
#pragma unroll
                        for (int i = 0; i < 4; ++i) c[T][i] = 0;
#pragma unroll
                    for (int k = 0; k < NKN; ++k) {
                        if (k < nkq) {
                            uint32_t a0[4];
                            ldmatrix_x4(a0, a_addr + k * 32 + rb * 16 * pitch_bytes);
#pragma unroll
                            for (int T = 0; T < 4; ++T) mma_u8s8(c[T], a0, nreg[k][T]);
                        }
                    }
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        const int row = rb * 16 + hf * 8 + g;
#pragma unroll
                        for (int T = 0; T < 4; ++T) {
                            const int32_t xa = __shfl_sync(0xffffffffu, c[T][2 * hf], src);
                            const int32_t xb = __shfl_sync(0xffffffffu, c[T][2 * hf + 1], src);
                            const int32_t v = c[T][2 * hf] + (c[T][2 * hf + 1] << 8) + ((t == 0 ? xa : xb) << 16);
                            if (t < 2) mine[(2 * T + t) * kPartOut + row] = v;
                        }
                        if (t == 3) mine[8 * kPartOut + row] = c[0][2 * hf] + (c[0][2 * hf + 1] << 8) + (c[1][2 * hf] << 16);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive1(&l_empty[lb]);  // this warp no longer reads the luma buffer
                asm volatile("bar.sync 2, 128;" ::: "memory");  // the four quarter sums of this chunk are written
                {
                    const int32_t* all = s_part + pb * 4 * (kDW * kPartOut);
                    // CR == 32: lane = row, one output after the other; CR == 16: lanes 0..15 / 16..31 = first / second output
                    const int row = CR == 32 ? lane : (lane & 15);
#pragma unroll
                    for (int pass = 0; pass < (CR == 32 ? 3 : 2); ++pass) {
                        const int oi = CR == 32 ? pass : 2 * pass + (lane >> 4);
                        if (oi < n_own) {
                            const int32_t* p = all + (2 * j + oi) * kPartOut + row;
                            const int32_t v = p[0] + p[kDW * kPartOut] + p[2 * kDW * kPartOut] + p[3 * kDW * kPartOut] + (1 << (kPrec - 1));
                            scr[(scr_col + oi) * kHP + row_off + row] = (uint8_t)pack_sat_u8(0, v >> kPrec);
                        }
                    }
                }
                __syncwarp();  // the pair's columns are written
                const bool vstep = CR == 32 || (hc & 1) || r0 + CR >= a.h;
                const int ci = CR == 32 ? hc : hc >> 1;
                if (vstep && !(dbg & 4) && ci >= a.v_lo[2] && ci <= a.v_hi[2]) {
                    const uint32_t b0 = col[t], b1 = col[4 + t];
                    const uint4* af = a.vmma + ((size_t)(ci * 3 + 2) * 3) * 32 + lane;
#pragma unroll
                    for (int d = 0; d < 3; ++d) mma_s8u8(vc[d], __ldg(af + d * 32), b0, b1);
                }
                __syncwarp();  // every lane has read the columns before the next chunk overwrites them
            }
            {   // 8x9 plane: rows g; tile columns 0, 1 (warp 7: 0, 1, 2) are outputs 2 (warp - 4) + 0, 1 (, 2)
                const int32_t v0 = vc[0][0] + (vc[1][0] << 8) + (vc[2][0] << 16);
                const int32_t v1 = vc[0][1] + (vc[1][1] << 8) + (vc[2][1] << 16);
                const uint32_t pk = pack_sat_u8(v1 >> kPrec, v0 >> kPrec);
                const int x = 2 * (warp - 4) + 2 * t;
                if (t == 0) {
                    s_x98[g * kDW + x] = (uint8_t)(pk & 0xFFu);
                    s_x98[g * kDW + x + 1] = (uint8_t)(pk >> 8);
                } else if (t == 1 && warp == 7) {
                    s_x98[g * kDW + x] = (uint8_t)(pk & 0xFFu);
                }
            }
            compute_sync<NW>();
            dct_and_bits<NW>(a, im, s_x32, s_x98, s_t, s_y, tid, lane, warp);
        }
        return;
    }
    if constexpr (NKW == kNKP) {
        // Two CTAs per SM (72 registers here): outputs in two groups, (0..3) -> warps 4, 5 and (4..8) -> warps 6, 7; the two
        // warps of a group split the group's band in halves (fragments in shared memory, two tiles per k-step), so each
        // luma byte is read by at most two narrow warps instead of four.  The upper half hands its twelve accumulator
        // words per lane to the lower one through shared memory (same lane mapping on both sides: conflict free, double
        // buffered, one 64-thread named barrier per chunk); the lower warp rounds, writes the group's columns and runs
        // their vertical pass.
        const int j = warp - 4, grp = j >> 1;
        const bool upper = j & 1;
#ifdef KE_TUNING_PROBES
        const int nkh = ((dbg & 2) || (dbg & 64)) ? 0 : a.mmah_nk[j];
#else
        const int nkh = (dbg & 2) ? 0 : a.mmah_nk[j];
#endif
        const uint2* bh = s_b + (a.mmah_boff[j] - b_first) + lane;
        const uint32_t ah_off = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * pitch_bytes + (lane >> 4) * 16 + a.mmah_k0[j] * 32);
        const int src = (lane & ~3) | 2;  // the quad's lane holding the third digits {a.d2, b.d2}
        int32_t* xch = s_part + grp * (2 * kHalfSlots * 32) + lane;
        const int gcol = 8 * j;  // the lower warp's private columns of plane 1 (j = 0 or 2)
        const uint32_t* colh = reinterpret_cast<const uint32_t*>(scr + (gcol + g) * kHP);
        int pb = 0;
        for (long long im = blockIdx.x; im < a.n; im += gridDim.x) {
#pragma unroll
            for (int d = 0; d < 3; ++d)
#pragma unroll
                for (int i = 0; i < 4; ++i) vc[d][i] = d == 0 ? (1 << (kPrec - 1)) : 0;
            int hc = 0;
            for (int r0 = 0; r0 < a.h; r0 += CR, ++hc, lph ^= (lb + 1 == nlb), lb = lb + 1 == nlb ? 0 : lb + 1, pb ^= 1) {
                mbar_wait_sleep(&l_full[lb], lph, poll_ns);
                const uint32_t a_addr = smem_u32(s_luma + lb * L.luma_bytes) + ah_off;
                int32_t c[NRB][2][4];
#pragma unroll
                for (int rb = 0; rb < NRB; ++rb)
#pragma unroll
                    for (int T = 0; T < 2; ++T)
#pragma unroll
                        for (int i = 0; i < 4; ++i) c[rb][T][i] = 0;
#pragma unroll 2
                for (int k = 0; k < nkh; ++k) {
                    uint32_t a0[4], a1[4];
                    ldmatrix_x4(a0, a_addr + k * 32);
                    if (NRB == 2) ldmatrix_x4(a1, a_addr + k * 32 + 16 * pitch_bytes);
#pragma unroll
                    for (int T = 0; T < 2; ++T) {
                        const uint2 b = bh[(k * 2 + T) * 32];
                        mma_u8s8(c[0][T], a0, b);
                        if (NRB == 2) mma_u8s8(c[NRB - 1][T], a1, b);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive1(&l_empty[lb]);  // this warp no longer reads the luma buffer
                // digits -> values: lanes t < 2 hold output 2 T + t of the group, lanes t == 3 the ninth output (group 1)
                int32_t v[NRB][3][2];
#pragma unroll
                for (int rb = 0; rb < NRB; ++rb)
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
                        for (int T = 0; T < 2; ++T) {
                            const int32_t xa = __shfl_sync(0xffffffffu, c[rb][T][2 * hf], src);
                            const int32_t xb = __shfl_sync(0xffffffffu, c[rb][T][2 * hf + 1], src);
                            v[rb][T][hf] = c[rb][T][2 * hf] + (c[rb][T][2 * hf + 1] << 8) + ((t == 0 ? xa : xb) << 16);
                        }
                        v[rb][2][hf] = c[rb][0][2 * hf] + (c[rb][0][2 * hf + 1] << 8) + (c[rb][1][2 * hf] << 16);
                    }
                int32_t* x = xch + pb * (kHalfSlots * 32);
                if (upper) {
#pragma unroll
                    for (int rb = 0; rb < NRB; ++rb)
#pragma unroll
                        for (int T = 0; T < 3; ++T)
#pragma unroll
                            for (int hf = 0; hf < 2; ++hf) x[((rb * 3 + T) * 2 + hf) * 32] = v[rb][T][hf];
                }
                if (grp == 0) asm volatile("bar.sync 2, 64;" ::: "memory");
                else asm volatile("bar.sync 3, 64;" ::: "memory");
                if (!upper) {
                    const int row_off = CR == 32 ? 0 : (hc & 1) * 16;
#pragma unroll
                    for (int rb = 0; rb < NRB; ++rb)
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            const int row = row_off + rb * 16 + hf * 8 + g;
#pragma unroll
                            for (int T = 0; T < 2; ++T) {
                                const int32_t sum = v[rb][T][hf] + x[((rb * 3 + T) * 2 + hf) * 32] + (1 << (kPrec - 1));
                                if (t < 2) scr[(gcol + 2 * T + t) * kHP + row] = (uint8_t)pack_sat_u8(0, sum >> kPrec);
                            }
                            const int32_t s8 = v[rb][2][hf] + x[((rb * 3 + 2) * 2 + hf) * 32] + (1 << (kPrec - 1));
                            if (t == 3 && grp == 1) scr[(gcol + 4) * kHP + row] = (uint8_t)pack_sat_u8(0, s8 >> kPrec);
                        }
                    __syncwarp();  // the group's columns are written
                    const bool vstep = CR == 32 || (hc & 1) || r0 + CR >= a.h;
                    const int ci = CR == 32 ? hc : hc >> 1;
                    if (vstep && !(dbg & 4) && ci >= a.v_lo[2] && ci <= a.v_hi[2]) {
                        const uint32_t b0 = colh[t], b1 = colh[4 + t];
                        const uint4* af = a.vmma + ((size_t)(ci * 3 + 2) * 3) * 32 + lane;
#pragma unroll
                        for (int d = 0; d < 3; ++d) mma_s8u8(vc[d], __ldg(af + d * 32), b0, b1);
                    }
                    __syncwarp();  // every lane has read the columns before the next chunk overwrites them
                }
            }
            if (!upper) {  // 8x9 plane: rows g; tile column 2 t + e is output 4 grp + 2 t + e (group 1: five outputs)
                const int32_t v0 = vc[0][0] + (vc[1][0] << 8) + (vc[2][0] << 16);
                const int32_t v1 = vc[0][1] + (vc[1][1] << 8) + (vc[2][1] << 16);
                const uint32_t pk = pack_sat_u8(v1 >> kPrec, v0 >> kPrec);
                const int xo = 4 * grp + 2 * t;
                if (t < 2) {
                    s_x98[g * kDW + xo] = (uint8_t)(pk & 0xFFu);
                    s_x98[g * kDW + xo + 1] = (uint8_t)(pk >> 8);
                } else if (t == 2 && grp == 1) {
                    s_x98[g * kDW + 8] = (uint8_t)(pk & 0xFFu);
                }
            }
            compute_sync<NW>();
            dct_and_bits<NW>(a, im, s_x32, s_x98, s_t, s_y, tid, lane, warp);
        }
        return;
    }
    // NKW == 0 (rows beyond ~2200 pixels): warps 4..7 own output pairs of the 9-wide target (warp 7: outputs 6, 7 and 8) END
    // TO END, fragments behind a pointer: the pair's columns go into a private 8-column scratch (plane 1 of the row plane)
    // and come straight back as the B fragment of the 8x9 plane's vertical pass — six of the tile's eight columns are
    // padding, three MMAs per chunk are cheap, and no tap warp waits for another before the image is finished.
    const uint2* bw = s_b + (a.mma_boff[warp] - b_first) + lane;
    if (NKW == 0 && cfg.narrow_b == kBGmem) bw = a.mma_b + a.mma_boff[warp] + lane;  // (NKW == 8: always shared memory, LDS)
    for (long long im = blockIdx.x; im < a.n; im += gridDim.x) {
#pragma unroll
        for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int i = 0; i < 4; ++i) vc[d][i] = d == 0 ? (1 << (kPrec - 1)) : 0;
        int hc = 0;
        for (int r0 = 0; r0 < a.h; r0 += CR, ++hc, lph ^= (lb + 1 == nlb), lb = lb + 1 == nlb ? 0 : lb + 1) {
            mbar_wait_sleep(&l_full[lb], lph, poll_ns);
            const uint32_t a_addr = smem_u32(s_luma + lb * L.luma_bytes) + a_off;
            const int row_off = CR == 32 ? 0 : (hc & 1) * 16;
            if (warp < 7) v5_taps<1, false, NRB>(a_addr, bw, nk, pitch_bytes, scr, scr_col, lane, row_off);
            else v5_taps<2, false, NRB>(a_addr, bw, nk, pitch_bytes, scr, scr_col, lane, row_off);
            __syncwarp();  // the pair's columns are written
            if (lane == 0) mbar_arrive1(&l_empty[lb]);  // this warp no longer reads the luma buffer
            const bool vstep = CR == 32 || (hc & 1) || r0 + CR >= a.h;
            const int ci = CR == 32 ? hc : hc >> 1;
            if (vstep && !(dbg & 4) && ci >= a.v_lo[2] && ci <= a.v_hi[2]) {
                const uint32_t b0 = col[t], b1 = col[4 + t];
                const uint4* af = a.vmma + ((size_t)(ci * 3 + 2) * 3) * 32 + lane;
#pragma unroll
                for (int d = 0; d < 3; ++d) mma_s8u8(vc[d], __ldg(af + d * 32), b0, b1);
            }
            __syncwarp();  // every lane has read the columns before the next chunk overwrites them
        }
        {   // 8x9 plane: rows g; tile columns 0, 1 (warp 7: 0, 1, 2) are outputs 2 (warp - 4) + 0, 1 (, 2)
            const int32_t v0 = vc[0][0] + (vc[1][0] << 8) + (vc[2][0] << 16);
            const int32_t v1 = vc[0][1] + (vc[1][1] << 8) + (vc[2][1] << 16);
            const uint32_t pk = pack_sat_u8(v1 >> kPrec, v0 >> kPrec);
            const int x = 2 * (warp - 4) + 2 * t;
            if (t == 0) {
                s_x98[g * kDW + x] = (uint8_t)(pk & 0xFFu);
                s_x98[g * kDW + x + 1] = (uint8_t)(pk >> 8);
            } else if (t == 1 && warp == 7) {
                s_x98[g * kDW + x] = (uint8_t)(pk & 0xFFu);
            }
        }
        compute_sync<NW>();
        dct_and_bits<NW>(a, im, s_x32, s_x98, s_t, s_y, tid, lane, warp);
    }
}

template <int C>
bool v5_config(const PhashArgs& a, V5Config& cfg, int forced) {
    const long long row_bytes = (long long)a.w * C;
    if (a.row_stride != row_bytes || a.mma_words < 1 || !a.vmma) return false;  // strided rows: generic kernel
    cfg.aligned = (row_bytes & 15) == 0 && (a.img_stride & 15) == 0 && (a.w & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(a.img) & 15) == 0;
    int nk_wide = 0, nk_narrow = 0;
    for (int q = 0; q < 4; ++q) nk_wide = std::max(nk_wide, a.mma_nk[q]), nk_narrow = std::max(nk_narrow, a.mma_nk[4 + q]);
    cfg.pitch_bytes = (a.w + 31) / 32 * 32 + 16;  // rows of 16-byte units, odd count: ldmatrix reads are conflict free
    cfg.dbg = 0;
    const int wide_words = a.mma_boff[4], narrow_words = a.mma_words - a.mma_boff[4];
    // wide-target bands of <= 8 k-steps stay in registers at two CTAs per SM, of <= 32 at one CTA per SM
    // (and the narrow warps a quarter of the row each: <= 4 / 9 / 18 k-steps)
    const int kq = a.mmaq_kq;
    const bool force_one = forced && ((forced >> 21) & 1);  // tuning: the one-CTA kernel on short rows too
    const int nkw_reg = nk_wide <= kNKP && !force_one ? kNKP : nk_wide <= 16 && kq <= 9 ? 16 : nk_wide <= 32 && kq <= 18 ? 32 : 0;
    auto fits = [&](int budget, int wide, int narrow, int bufs, int sub, int shift, int cr) -> bool {
        if (sub < 1 || sub > cr || (cr % sub) || sub * row_bytes > (1 << 20) || shift < 1 || shift > 3 || bufs < 1 ||
            bufs > kMaxLumaBufs)
            return false;
        if (wide == kBReg && nkw_reg == 0) return false;
        const int nkw = wide == kBReg ? nkw_reg : 0;
        if (nkw == kNKP && (cr == 16 || narrow != kBSmem)) return false;  // the two-CTA kernel: 32-row buffers, fragments on chip
        if (nkw > kNKP && budget < 227 * 1024) return false;                // 168 registers per thread: one CTA per SM
        if (a.n * ((a.h + sub - 1) / sub) >= (1ll << 31)) return false;
        // unaligned rows: room for the aligned superset (<= 15 B in front, <= 15 B behind) and for the luma loads that
        // run a few words past the last pixel
        const int slot = cfg.aligned ? (int)(sub * row_bytes) : (int)((sub * row_bytes + 64 + 127) / 128 * 128);
        if (nkw > kNKP) narrow = kBReg;  // one CTA per SM: the K-split narrow warps hold their fragments too, nothing in shared memory
        const int words = nkw == kNKP ? a.mmah_end - a.mmah_boff[0] : (wide == kBSmem ? wide_words : 0) + (narrow == kBSmem ? narrow_words : 0);
        const V5Layout L = v5_layout(slot, cfg.pitch_bytes, 1 << shift, words, bufs, cr, v5_part_bytes(nkw));
        if (L.total > budget) return false;
        cfg.sub_rows = sub, cfg.slot_shift = shift, cfg.nlb = bufs, cfg.slot_bytes = slot, cfg.cr = cr;
        cfg.wide_b = wide, cfg.narrow_b = narrow, cfg.smem_words = words, cfg.L = L, cfg.nkw = nkw;
        return true;
    };
    if (forced) {  // KE_OPT_PHASH_CFG (tests / tuning): sub_rows | slot_shift << 8 | luma buffers << 12 | placement << 16 | cr16 << 20
        // placement 0: wide fragments in registers where the band allows (else shared memory), narrow in shared memory;
        // 1: wide in L2 (registers where the band fits the two-CTA kernel), narrow in shared memory; 2: wide as 1, narrow in
        // L2; 3: both in shared memory (the pointer-fed loops)
        const int place = (forced >> 16) & 3;
        const int cr = ((forced >> 20) & 1) ? 16 : 32;
        const int chip = nkw_reg && !(nkw_reg == kNKP && cr == 16) ? kBReg : kBSmem;
        const int l2 = nkw_reg == kNKP && cr == 32 ? kBReg : kBGmem;
        const int wide = place == 0 ? chip : place == 3 ? kBSmem : l2;
        return fits(227 * 1024, wide, place == 2 ? kBGmem : kBSmem, (forced >> 12) & 15, forced & 255, (forced >> 8) & 15, cr);
    }
    // What matters (measured, tools/probe_phash_large.py and probe_phash_roles.py: 512 / 1024 / 2048 / 4000-pixel rows):
    // a double-buffered luma ring, raw sub-chunks of >= 12 KB (small bulk copies are latency bound: 2-row copies cost
    // half the bandwidth at 2048 pixels), two CTAs per SM where everything still fits — and, on longer rows, as little
    // shared-memory traffic per pixel as possible: the wide-target fragments in registers (one CTA per SM).
    struct Place { int budget, wide, narrow; };
    // two CTAs per SM with everything on chip, as long as a sub-chunk of >= 12 KB (or 16 rows) still fits beside them
    const long long min_slot = std::min<long long>(12 * 1024, 16 * row_bytes);
    if (nkw_reg == kNKP)
        for (int sub : {16, 8, 4, 2, 1})
            if (sub * row_bytes >= min_slot && fits(113 * 1024, kBReg, kBSmem, 2, sub, 1, 32)) return true;
    // Long rows: one CTA per SM.  With pointer-fed fragments 16-row luma buffers — a finer hand-off between luma and tap
    // warps — were worth 0.61 -> 0.70 of HBM at 1024 pixels and 0.59 -> 0.71 at 2048; the largest sub-chunk that fits, the
    // narrow-target fragments moved out to L2 where that is what it takes.
    const int wide_chip = nkw_reg ? kBReg : kBSmem, wide_l2 = nkw_reg ? kBReg : kBGmem;
    const Place places[] = {{227 * 1024, wide_chip, kBSmem}, {227 * 1024, wide_l2, kBSmem}, {227 * 1024, wide_l2, kBGmem}};
    // Between the regimes (bands of 9+ k-steps on rows still short enough for TWO CTAs of the pointer-fed kernel, every
    // fragment in L2): two CTAs hide each other's per-chunk chain, which a single CTA on short rows cannot — measured
    // 0.71 against 0.58 at 640 pixels, 0.71 / 0.72 against 0.68 / 0.69 at 768 / 800; at 1024 the register kernel wins (0.84+)
    if (nkw_reg > kNKP && a.w < 900)
        for (int sub : {16, 8, 4})
            for (int narrow : {kBSmem, kBGmem})
                if (sub * row_bytes >= min_slot && fits(113 * 1024, kBGmem, narrow, 2, sub, 1, 32)) return true;
    // register-resident bands: 32-row buffers while two of them fit beside sub-chunks of >= 8 rows (1024 pixels: 0.85 of
    // HBM against 0.82-0.84 with 16-row buffers), else 16-row buffers (2048 pixels: 0.887 with 8-row sub-chunks; 4-row
    // sub-chunks, whatever the buffers, stay at 0.63)
    if (nkw_reg > kNKP)
        for (int sub : {16, 8})
            if (sub * row_bytes >= min_slot && fits(227 * 1024, kBReg, kBSmem, 2, sub, 1, 32)) return true;
    if (nk_wide > kNKP)
        for (int sub : {16, 8, 4, 2, 1})
            for (const Place& pl : places)
                if (fits(pl.budget, pl.wide, pl.narrow, 2, sub, 1, 16)) return true;
    for (int bufs : {2, 1})
        for (int sub : {16, 8, 4, 2, 1})
            for (const Place& pl : places)
                if (fits(pl.budget, pl.wide, pl.narrow, bufs, sub, 1, 32)) return true;
    if (nk_wide > kNKP)
        for (int sub : {16, 8, 4, 2, 1})
            for (const Place& pl : places)
                if (fits(pl.budget, pl.wide, pl.narrow, 1, sub, 1, 16)) return true;
    return false;
}

template <int C>
int launch_v5(ke_ctx* ctx, const PhashArgs& a, V5Config& cfg, cudaStream_t s) {
    using Kernel = void (*)(const PhashArgs, const int, const int, const int, const int, const int, const V5Config);
    Kernel kernel = nullptr;
    const bool al = cfg.aligned != 0, c16 = cfg.cr == 16;
    switch (cfg.nkw) {
        case kNKP: kernel = al ? ke_phash_v5_kernel<C, kNKP, true, 32> : ke_phash_v5_kernel<C, kNKP, false, 32>; break;
        case 16:
            kernel = c16 ? (al ? ke_phash_v5_kernel<C, 16, true, 16> : ke_phash_v5_kernel<C, 16, false, 16>)
                         : (al ? ke_phash_v5_kernel<C, 16, true, 32> : ke_phash_v5_kernel<C, 16, false, 32>);
            break;
        case 32:
            kernel = c16 ? (al ? ke_phash_v5_kernel<C, 32, true, 16> : ke_phash_v5_kernel<C, 32, false, 16>)
                         : (al ? ke_phash_v5_kernel<C, 32, true, 32> : ke_phash_v5_kernel<C, 32, false, 32>);
            break;
        default:
            kernel = c16 ? (al ? ke_phash_v5_kernel<C, 0, true, 16> : ke_phash_v5_kernel<C, 0, false, 16>)
                         : (al ? ke_phash_v5_kernel<C, 0, true, 32> : ke_phash_v5_kernel<C, 0, false, 32>);
    }
    KE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg.L.total));
    int per_sm = 0;
    KE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kV5Threads, cfg.L.total));
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)ctx->sm_count * per_sm;
    if (grid > a.n) grid = a.n;
    int dbg = 0, ns = 200;  // consumer poll interval in ns
#ifdef KE_TUNING_PROBES
    if (const char* dbg_env = getenv("KE_PHASH_DBG")) dbg = atoi(dbg_env) & 0xFF;  // bit 2: no horizontal MMA, 4: no vertical pass, ...
    if (const char* ns_env = getenv("KE_PHASH_SLEEP")) ns = atoi(ns_env);
#endif
    cfg.dbg = dbg | (ns << 8);
    kernel<<<(unsigned)grid, kV5Threads, cfg.L.total, s>>>(a, cfg.sub_rows, cfg.slot_shift, cfg.pitch_bytes, cfg.nlb, cfg.dbg, cfg);
    ctx->launches++;
    KE_CUDA(cudaGetLastError());
    return KE_OK;
}

int g_dct_uploaded_device = -1;

int ensure_dct(ke_ctx* ctx) {
    if (g_dct_uploaded_device == ctx->device) return KE_OK;
    double c[8 * 32];
    for (int k = 0; k < 8; ++k)
        for (int n = 0; n < 32; ++n)
            c[k * 32 + n] = (k == 0 ? std::sqrt(1.0 / 32.0) : std::sqrt(2.0 / 32.0)) * std::cos(M_PI * (2 * n + 1) * k / 64.0);
    KE_CUDA(cudaMemcpyToSymbol(c_dct, c, sizeof(c)));
    g_dct_uploaded_device = ctx->device;
    return KE_OK;
}

template <int C>
int launch_phash(ke_ctx* ctx, PhashArgs& a, cudaStream_t s) {
    // rows per chunk: as many (power of two, <= 32) as keep the CTA's shared memory <= ~110 KB
    // (two CTAs per SM); shrink further until it fits the 227 KB hardware limit.
    const long long row_bytes = (long long)a.w * C;
    a.pitch_words = ((a.w + 3) / 4) | 1;
    // Two kernels: the streaming tensor-pipe kernel (v5) takes every batch of contiguous rows whose resample band fits
    // its fragment budget — any width, any byte alignment; the generic kernel takes the rest (strided rows, very wide
    // images) and is the in-library reference the parity tests compare v5 with (KE_OPT_PHASH_GENERIC).
    if (!ctx->force_generic_phash) {
        V5Config cfg;
        if (v5_config<C>(a, cfg, ctx->phash_cfg)) return launch_v5<C>(ctx, a, cfg, s);
        if (ctx->phash_cfg) {
            ke_set_error("ke_phash_batch: the pinned staging configuration 0x%x does not fit this geometry", ctx->phash_cfg);
            return KE_E_UNSUPPORTED;
        }
    }
    int rc_rows = 32;
    SmemLayout L;
    for (;;) {
        a.rows_per_chunk = rc_rows;
        a.raw_bytes = (int)(((long long)rc_rows * row_bytes + 15 + 16 + 15) / 16 * 16);
        L = smem_layout(a.coef_words, a.raw_bytes, rc_rows, a.pitch_words);
        const int budget = rc_rows > 1 ? 112 * 1024 : 227 * 1024;
        if (L.total <= budget || rc_rows == 1) break;
        rc_rows >>= 1;
    }
    if (L.total > 227 * 1024) {
        // retry preferring one CTA per SM with more rows
        ke_set_error("ke_phash_batch: image rows of %lld bytes do not fit shared memory", row_bytes);
        return KE_E_UNSUPPORTED;
    }
    KE_CUDA(cudaFuncSetAttribute(ke_phash_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    int per_sm = 0;
    KE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ke_phash_kernel<C>, kThreads, L.total));
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)ctx->sm_count * per_sm;
    if (grid > a.n) grid = a.n;
    ke_phash_kernel<C><<<(unsigned)grid, kThreads, L.total, s>>>(a);
    ctx->launches++;
    KE_CUDA(cudaGetLastError());
    return KE_OK;
}

}  // namespace

extern "C" int ke_phash_batch(ke_ctx* ctx, const uint8_t* d_img, int64_t n, int h, int w, int c, int64_t img_stride,
                              int64_t row_stride, uint64_t* d_phash, uint64_t* d_dhash, float* d_min_margin,
                              uint8_t* d_plane32, uint8_t* d_plane9x8, void* stream) {
    KE_REQUIRE(ctx != nullptr, "ke_phash_batch: ctx is NULL");
    KE_REQUIRE(n >= 0, "ke_phash_batch: n < 0");
    if (n == 0) return KE_OK;
    KE_REQUIRE(d_img && d_phash && d_dhash, "ke_phash_batch: NULL buffer");
    KE_REQUIRE(h > 0 && w > 0, "ke_phash_batch: empty image %dx%d", w, h);
    KE_REQUIRE(c == 1 || c == 3 || c == 4, "ke_phash_batch: channels must be 1, 3 or 4 (got %d)", c);
    KE_REQUIRE(row_stride >= (int64_t)w * c && img_stride >= (int64_t)(h - 1) * row_stride + (int64_t)w * c,
               "ke_phash_batch: strides smaller than the image");
    KeDeviceGuard guard(ctx->device);
    int rc;
    if ((rc = ensure_dct(ctx))) return rc;
    const HTable* ht;
    const VTable* vt;
    if ((rc = get_htable(ctx, w, &ht))) return rc;
    if ((rc = get_vtable(ctx, h, &vt))) return rc;
    PhashArgs a;
    a.img = d_img;
    a.n = n;
    a.h = h;
    a.w = w;
    a.c = c;
    a.img_stride = img_stride;
    a.row_stride = row_stride;
    a.total_bytes = (n - 1) * img_stride + (int64_t)(h - 1) * row_stride + (int64_t)w * c;
    a.use_bulk = (row_stride == (int64_t)w * c) && ((reinterpret_cast<uintptr_t>(d_img) & 15) == 0);
    a.coef = ht->d_coef;
    a.items = ht->d_items;
    a.mma_b = ht->d_mma_b;
    a.mma_words = ht->mma_words;
    for (int i = 0; i < 8; ++i) a.mma_k0[i] = ht->mma_k0[i], a.mma_nk[i] = ht->mma_nk[i], a.mma_boff[i] = ht->mma_boff[i];
    for (int i = 0; i < 4; ++i) a.mmaq_k0[i] = ht->mmaq_k0[i], a.mmaq_nk[i] = ht->mmaq_nk[i], a.mmaq_boff[i] = ht->mmaq_boff[i];
    a.mmaq_kq = ht->mmaq_kq;
    for (int i = 0; i < 4; ++i) a.mmah_k0[i] = ht->mmah_k0[i], a.mmah_nk[i] = ht->mmah_nk[i], a.mmah_boff[i] = ht->mmah_boff[i];
    a.mmah_end = ht->mmah_end;
    a.vmma = vt->vmma_ok ? vt->d_vmma : nullptr;
    for (int i = 0; i < 3; ++i) a.v_lo[i] = vt->v_lo[i], a.v_hi[i] = vt->v_hi[i];
    a.meta = ht->d_meta;
    a.n_items = ht->n_items;
    a.coef_words = ht->coef_words;
    a.kk32 = vt->d_kk32;
    a.b32 = vt->d_b32;
    a.kk8 = vt->d_kk8;
    a.b8 = vt->d_b8;
    a.ks32 = vt->ks32;
    a.ks8 = vt->ks8;
    a.phash = d_phash;
    a.dhash = d_dhash;
    a.min_margin = d_min_margin;
    a.plane32 = d_plane32;
    a.plane98 = d_plane9x8;
    cudaStream_t s = (cudaStream_t)stream;
    switch (c) {
        case 1: return launch_phash<1>(ctx, a, s);
        case 3: return launch_phash<3>(ctx, a, s);
        default: return launch_phash<4>(ctx, a, s);
    }
}

// Single-device body of ke_phash_batch_host (ke_multi.cu fans it over the devices of a context): chunks of <= 256 MB go
// host -> device on two alternating streams (pinned sources by DMA in place, pageable ones through the context's pinned
// staging buffers), each chunk's kernel follows its copy on the same stream, so copy k+1 overlaps kernel k.
int ke_phash_batch_host_one(ke_ctx* ctx, const uint8_t* h_img, int64_t n, int h, int w, int c, uint64_t* h_phash,
                            uint64_t* h_dhash, float* h_min_margin) {
    if (n == 0) return KE_OK;
    KeDeviceGuard guard(ctx->device);
    const int64_t img_bytes = (int64_t)h * w * c;
    const int64_t img_stride = (img_bytes + 15) / 16 * 16;  // keep every image 16-byte aligned on the device
    int64_t per_chunk = (256ll << 20) / img_stride;
    if (per_chunk < 1) per_chunk = 1;
    if (per_chunk > n) per_chunk = n;
    void *d_buf[2], *d_ph = nullptr, *d_dh = nullptr, *d_mm = nullptr;
    int rc;
    for (int b = 0; b < 2; ++b)
        if ((rc = ke_ctx_scratch(ctx, b, (size_t)(per_chunk * img_stride) + 64, &d_buf[b]))) return rc;
    if ((rc = ke_ctx_scratch(ctx, 2, (size_t)n * 8, &d_ph))) return rc;
    if ((rc = ke_ctx_scratch(ctx, 3, (size_t)n * 8, &d_dh))) return rc;
    if ((rc = ke_ctx_scratch(ctx, 4, (size_t)n * 4, &d_mm))) return rc;
    int k = 0;
    for (int64_t i0 = 0; i0 < n; i0 += per_chunk, ++k) {
        const int b = k & 1;
        const int64_t cnt = std::min<int64_t>(per_chunk, n - i0);
        cudaStream_t s = ctx->copy_stream[b];
        if ((rc = ke_h2d_staged_2d(ctx, d_buf[b], (size_t)img_stride, h_img + i0 * img_bytes, (size_t)img_bytes, (size_t)cnt, s)))
            return rc;
        rc = ke_phash_batch(ctx, (const uint8_t*)d_buf[b], cnt, h, w, c, img_stride, (int64_t)w * c,
                            (uint64_t*)d_ph + i0, (uint64_t*)d_dh + i0, (float*)d_mm + i0, nullptr, nullptr, s);
        if (rc) return rc;
    }
    for (auto s : ctx->copy_stream) KE_CUDA(cudaStreamSynchronize(s));
    KE_CUDA(cudaMemcpy(h_phash, d_ph, (size_t)n * 8, cudaMemcpyDeviceToHost));
    KE_CUDA(cudaMemcpy(h_dhash, d_dh, (size_t)n * 8, cudaMemcpyDeviceToHost));
    if (h_min_margin) KE_CUDA(cudaMemcpy(h_min_margin, d_mm, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return KE_OK;
}
