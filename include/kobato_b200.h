/*
 * kobato_b200.h — C ABI of libkobato_b200.so: the B200 (sm_100a) implementation of
 * kobato-eyes' duplicate-detection hot path.
 *
 * The reference (srndpty/kobato-eyes v0.6.0) is pure Python and has no FFI layer; each entry
 * point below replaces the arithmetic behind a Python seam of the reference (paths relative
 * to the reference checkout).  INTEGRATION.md shows the ctypes stub a maintainer adds.
 *
 * Conventions: plain C, caller owns every buffer, the library owns only the opaque context.
 * Nothing throws: every call returns KE_OK (0) or a negative ke_status and records a
 * thread-local message readable through ke_last_error().  `d_` pointers are device memory on
 * the context's device, `h_` pointers are host memory.  `stream` is a cudaStream_t passed as
 * void* (NULL = default stream); `d_` entry points only enqueue work on it.  There is no CPU
 * fallback anywhere: without a usable CUDA device ke_ctx_create fails.
 */
#ifndef KOBATO_B200_H
#define KOBATO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KE_ABI_VERSION 1

typedef enum ke_status {
    KE_OK = 0,
    KE_E_INVALID = -1,     /* bad argument (the reference raises ValueError there) */
    KE_E_CUDA = -2,        /* CUDA runtime error, see ke_last_error() */
    KE_E_CAPACITY = -3,    /* output buffer too small; required count was written */
    KE_E_NOMEM = -4,
    KE_E_UNSUPPORTED = -5  /* e.g. SSIM on an image smaller than the 7x7 window */
} ke_status;

typedef struct ke_ctx ke_ctx;

int ke_abi_version(void);
const char* ke_last_error(void);

/* One context per (process, device).  Calls on one context must be serialised by the caller
 * (the reference drives this path from a single Qt worker thread: src/ui/dup_tab.py:118). */
int ke_ctx_create(int device, ke_ctx** out);
void ke_ctx_destroy(ke_ctx* ctx);
int ke_ctx_device(const ke_ctx* ctx);
/* Tuning / test knobs.  KE_OPT_PHASH_GENERIC=1 routes every image geometry through K1's generic
 * kernel (the one used for unaligned or very wide rows) instead of the fast one. */
#define KE_OPT_PHASH_GENERIC 1
/* KE_OPT_JOIN_MODE: 0 auto (hybrid for large tables with threshold <= 15), 1 POPC kernel only,
 * 2 hybrid (POPC kernel + bit-sliced LOP3 kernel running concurrently), 3 bit-sliced kernel only. */
#define KE_OPT_JOIN_MODE 2
int ke_ctx_set_option(ke_ctx* ctx, int option, int value);
int ke_ctx_sm_count(const ke_ctx* ctx);

/* ---------------------------------------------------------------------------------------
 * Pillow-compatible 8bpc LANCZOS coefficient tables (host, double precision + libm sin).
 * Replaces the table construction inside `image.resize(size, LANCZOS)` called at
 * src/sig/phash.py:25 (Pillow libImaging/Resample.c precompute_coeffs + normalize_coeffs_8bpc).
 * kk: out_size*ksize int32 taps (22-bit fixed point), bounds: out_size*2 int32 (first, count). */
int ke_resample_ksize(int in_size, int out_size);
int ke_resample_table(int in_size, int out_size, int32_t* kk, int32_t* bounds, int ksize);

/* ---------------------------------------------------------------------------------------
 * K1 — batched pHash + dHash.  Replaces sig.phash.phash / sig.phash.dhash
 * (src/sig/phash.py:33-57, incl. _to_grayscale :21-26) for a batch of decoded images, i.e. the
 * body of core.fastsig._compute_worker (src/core/fastsig.py:24-37) after Image.open and of
 * core.signature.compute_signatures_from_image (src/core/signature.py:24-28).
 *
 * d_img: n images of h x w x c uint8 (c = 1 'L', 3 'RGB', 4 'RGBA'/'RGBX', alpha ignored exactly
 * like Pillow's RGBA->L), image k at d_img + k*img_stride, row y at + y*row_stride.
 * Outputs (device): phash/dhash as UNSIGNED 64-bit, first compared element = MSB
 * (src/sig/phash.py:43-46); the caller wraps to signed for SQLite like _to_signed (:29-30).
 * d_min_margin (nullable): min |coef - mean| of the 64 pHash comparisons (a near-tie flag).
 * d_plane32 / d_plane9x8 (nullable): the 32x32 and 8x9 uint8 planes, byte-identical to
 * `convert("L").resize(...)`, for verification. */
int ke_phash_batch(ke_ctx* ctx, const uint8_t* d_img, int64_t n, int h, int w, int c, int64_t img_stride,
                   int64_t row_stride, uint64_t* d_phash, uint64_t* d_dhash, float* d_min_margin,
                   uint8_t* d_plane32, uint8_t* d_plane9x8, void* stream);

/* Same, host buffers in and out (densely packed n*h*w*c): chunked H2D through pinned staging
 * overlapped with the kernel, results copied back.  This is the call behind the drop-in
 * core.fastsig.compute_signatures_mp (src/core/fastsig.py:65-99). */
int ke_phash_batch_host(ke_ctx* ctx, const uint8_t* h_img, int64_t n, int h, int w, int c, uint64_t* h_phash,
                        uint64_t* h_dhash, float* h_min_margin);

/* ---------------------------------------------------------------------------------------
 * K2 — all-pairs 64-bit Hamming threshold join.  Replaces the candidate search of
 * dup.scanner.DuplicateScanner.build_clusters (src/dup/scanner.py:227-290; distance =
 * sig.phash.hamming64, src/sig/phash.py:60-63): emits every i<j with popcount(h[i]^h[j]) <=
 * threshold.  With KE_JOIN_REQUIRE_BAND a pair must also agree on at least one of the
 * `band_count` bands of `band_bits` bits ((h >> band*band_bits) & mask), which makes the output
 * exactly the reference's LSH edge candidates; d_band_allow (nullable, one uint64 per hash, bit
 * b = "my bucket in band b is not skipped by KE_DUP_BUCKET_PAIR_CAP", src/dup/scanner.py:239-266)
 * additionally masks bands.  The N x N upper triangle is tiled; this call evaluates the tiles
 * t with t % part_count == part_index (multi-GPU split: same table on every GPU, no exchange).
 * Results are unordered.  *d_count receives the number of hits even beyond `capacity`. */
#define KE_JOIN_REQUIRE_BAND 1u

int ke_hamming_join(ke_ctx* ctx, const uint64_t* d_hashes, int64_t n, int threshold, uint32_t flags, int band_bits,
                    int band_count, const uint64_t* d_band_allow, int part_index, int part_count, uint32_t* d_out_i,
                    uint32_t* d_out_j, uint8_t* d_out_dist, int64_t capacity, unsigned long long* d_count,
                    void* stream);

/* Host buffers in and out.  Returns KE_E_CAPACITY (and the required count in *out_count) when
 * more than `capacity` pairs qualify: nothing is silently truncated. */
int ke_hamming_join_host(ke_ctx* ctx, const uint64_t* h_hashes, int64_t n, int threshold, uint32_t flags,
                         int band_bits, int band_count, const uint64_t* h_band_allow, int part_index, int part_count,
                         uint32_t* h_out_i, uint32_t* h_out_j, uint8_t* h_out_dist, int64_t capacity,
                         int64_t* out_count);

/* Number of pairs the (part_index, part_count) share of the triangle covers (for pairs/s). */
int64_t ke_hamming_join_pairs(int64_t n, int part_index, int part_count);

/* ---------------------------------------------------------------------------------------
 * K3 — batched SSIM.  Replaces `structural_similarity(a, b, data_range=1.0)` as called by
 * dup.refine._compute_ssim (src/dup/refine.py:44-52): 7x7 uniform window, sample covariance,
 * K1=0.01, K2=0.03, 3-pixel border cropped, mean over the interior.  Window sums are exact
 * integers, the per-pixel formula FP32, the mean FP64 (|delta| vs the reference's float32 path
 * < 1e-5).  Pair p compares images ia[p] and ib[p] of a bank of h x w x c uint8 images
 * (c = 1: 'L' planes as _compute_ssim prepares them; c = 3/4: RGB(A), converted with Pillow's
 * fixed-point luma on the fly).  h, w >= 7 or KE_E_UNSUPPORTED (the reference raises). */
int ke_ssim_batch(ke_ctx* ctx, const uint8_t* d_bank, int h, int w, int c, int64_t img_stride, int64_t row_stride,
                  const int64_t* d_ia, const int64_t* d_ib, int64_t n_pairs, double* d_ssim, void* stream);

/* Host buffers: pair p compares h_a + p*h*w*c with h_b + p*h*w*c (c = 1: 'L' planes), copied in
 * chunks overlapped with the kernel.  Behind the drop-in dup.refine._compute_ssim /
 * refine_pairs_batch. */
int ke_ssim_pairs_host(ke_ctx* ctx, const uint8_t* h_a, const uint8_t* h_b, int64_t n_pairs, int h, int w, int c,
                       double* h_ssim);

/* ---------------------------------------------------------------------------------------
 * N1 — the refinement the shipped UI runs after a scan (SURVEY §8f "next" row): tile aHash and
 * small-gray pixel MAE of ui.dup_refine_parallel (src/ui/dup_refine_parallel.py).
 *
 * ke_gray_resize_batch: `convert("L").resize((out_w, out_h), filter)` for a batch of decoded images
 * (filter 1 = LANCZOS, 2 = BILINEAR; Pillow's 8bpc fixed-point arithmetic, byte-identical) —
 * replaces the resize in tile_ahash_bits (:66-69) and _load_small_gray (:203-207).
 * d_mid is caller scratch of n*h*out_w bytes (untouched when the streaming kernel serves the shape), d_out receives
 * n*out_h*out_w bytes.
 * ke_tile_ahash_bits: planes n x (grid*tile)^2 -> bit strings of ceil((grid*tile)^2/32) uint32 words per
 * image, bit order (gy, gx, ty, tx), little endian (:71-83), bit = pixel > mean of its tile.
 * ke_bits_hamming_pairs: popcount(bits[ia] ^ bits[ib]) per pair (tile_hamming :86-88).
 * ke_plane_sad_pairs: sum |a-b| over two planes per pair; _mae01 (:210-212) = sad / plane_bytes / 255. */
int ke_gray_resize_batch(ke_ctx* ctx, const uint8_t* d_img, int64_t n, int h, int w, int c, int64_t img_stride,
                         int64_t row_stride, int out_w, int out_h, int filter, uint8_t* d_mid, uint8_t* d_out,
                         void* stream);
int ke_tile_ahash_bits(ke_ctx* ctx, const uint8_t* d_planes, int64_t n, int grid, int tile, uint32_t* d_bits,
                       void* stream);
int ke_bits_hamming_pairs(ke_ctx* ctx, const uint32_t* d_bits, int words, const int64_t* d_ia, const int64_t* d_ib,
                          int64_t n_pairs, int32_t* d_out, void* stream);
int ke_plane_sad_pairs(ke_ctx* ctx, const uint8_t* d_planes, int64_t plane_bytes, const int64_t* d_ia,
                       const int64_t* d_ib, int64_t n_pairs, uint64_t* d_out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Cluster assembly (host).  Union-find over accepted pairs with the reference's "smaller root wins" rule
 * (ClusterBuilder.build, src/dup/cluster.py:22-70; DisjointSet of dup.scanner, src/dup/scanner.py:176-200):
 * h_nodes[0..*n_nodes) receives the distinct ids that occur in the pairs GROUPED BY COMPONENT (components by ascending
 * representative, members ascending inside each) and h_node_rep[i] the representative (= smallest id) of the component
 * of h_nodes[i]; both buffers hold 2*n_pairs entries.
 * Ids are arbitrary int64 values (file ids or table indices). */
int ke_cluster_pairs_host(const int64_t* h_a, const int64_t* h_b, int64_t n_pairs, int64_t* h_nodes,
                          int64_t* h_node_rep, int64_t* n_nodes);

/* ---------------------------------------------------------------------------------------
 * Measurement helpers (bench.py / tests only). */

/* Synthetic image generator, the CUDA twin of kobato_b200.synth.synth_image (identical bytes). */
int ke_synth_images(ke_ctx* ctx, uint8_t* d_out, int64_t start, int64_t count, int h, int w, int c, int64_t n_set,
                    uint64_t seed, int planted_permille, void* stream);

/* POPC issue-rate microbenchmark: the integer roofline denominator of K2.
 * Returns POPC thread-instructions per SM clock per SM, and the SM clock it derived. */
int ke_microbench_popc(ke_ctx* ctx, int iters, double* popc_per_clk_per_sm, double* sm_clock_mhz);

/* Kernels launched by this library on this context since creation (bench.py `gpu_launches`). */
int64_t ke_ctx_launch_count(const ke_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* KOBATO_B200_H */
