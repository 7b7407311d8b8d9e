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

#define KE_ABI_VERSION 2

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

/* A context drives the devices it was created over (`ke_ctx_create(const int* devices, int n, ...)`).
 * The reference runs this whole path from ONE Qt worker thread of ONE process (src/ui/dup_tab.py:118,
 * src/core/jobs.py:299), so using several GPUs behind its seams means the library fans the work out itself:
 *   - `d_` entry points act on the context they are handed: the first device of a multi-device context, or the
 *     per-device child returned by ke_ctx_child(ctx, k);
 *   - `_host` entry points split their units (images, triangle tiles, pairs) over ALL devices of the context with
 *     one host thread per device; there is no device-to-device traffic on this path, hence no collective.
 * Calls on one context must be serialised by the caller.  ke_ctx_create(device, ..) == ke_ctx_create_multi(&device, 1, ..). */
int ke_ctx_create(int device, ke_ctx** out);
int ke_ctx_create_multi(const int* devices, int n, ke_ctx** out);
void ke_ctx_destroy(ke_ctx* ctx);
int ke_ctx_device(const ke_ctx* ctx);
int ke_ctx_device_count(const ke_ctx* ctx);
ke_ctx* ke_ctx_child(ke_ctx* ctx, int k); /* k in [0, device_count): child 0 is ctx itself; owned by ctx */
/* Test knobs (the parity tests compare the kernel variants with each other through these).
 * KE_OPT_PHASH_GENERIC=1 routes every image geometry through K1's generic kernel (the one that takes strided /
 * unaligned / very wide rows) instead of the streaming tensor-pipe kernel. */
#define KE_OPT_PHASH_GENERIC 1
/* KE_OPT_JOIN_MODE: 0 auto (hybrid for large tables with threshold <= 15), 1 POPC kernel only,
 * 2 hybrid (POPC role + bit-sliced LOP3 role in one kernel), 3 bit-sliced kernel only. */
#define KE_OPT_JOIN_MODE 2
/* KE_OPT_PHASH_CFG: pin the streaming K1 kernel's staging (0 = automatic; tests and tuning): sub_rows | slot_shift << 8 |
 * luma_buffers << 12 | placement << 16 (0 resample fragments in registers where the band allows, else shared memory;
 * 1 wide-target fragments in L2; 2 both in L2; 3 both in shared memory) | 16-row ring buffers << 20 | the one-CTA-per-SM
 * kernel also on short rows << 21.  A configuration that does not fit the geometry makes the call return KE_E_UNSUPPORTED. */
#define KE_OPT_PHASH_CFG 5
/* KE_OPT_SSIM_V1=1: K3 on the one-column-per-thread kernel (the one that takes unaligned banks). */
#define KE_OPT_SSIM_V1 3
/* KE_OPT_RESIZE_GENERIC=1: N1 gray resize on the generic two-pass kernels instead of the streaming one. */
#define KE_OPT_RESIZE_GENERIC 4
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

/* Same, host buffers in and out (densely packed n*h*w*c).  Images are split into contiguous ranges over the context's
 * devices; per device, chunks of <= 256 MB go host -> device on two alternating streams with each chunk's kernel behind
 * its copy.  Page-locked sources (cudaHostAlloc / cudaHostRegister, torch pinned tensors) are DMA'd in place; pageable
 * ones pass through the context's pinned staging buffers (2 x 32 MB, filled by KE_STAGE_THREADS host threads, default
 * min(8, cores / 2): 35 GB/s against 55 GB/s from page-locked memory on the test box).
 * This is the call behind the drop-in core.fastsig.compute_signatures_mp (src/core/fastsig.py:65-99). */
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
 * more than `capacity` pairs qualify: nothing is silently truncated.  On a multi-device context the caller's share of
 * the tiles is dealt on over the devices (device k: t % (part_count * n_dev) == part_index + part_count * k; the table
 * is uploaded to each, the lists are concatenated on the host), one more device per ~5e9 pairs. */
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
 * fixed-point luma on the fly).  h, w >= 7 or KE_E_UNSUPPORTED (the reference raises).
 *
 * gaussian = 0 is the reference's path.  gaussian = 1 is skimage's `gaussian_weights=True` variant (sigma 1.5, 11 taps,
 * crop 5, cov_norm 121/120; h, w >= 11), which is how BASELINE.json words kernel 3 in its north_star; it mirrors
 * scipy.ndimage.gaussian_filter on float32 images (FP64 accumulation per 1-D pass, float32 in between). */
int ke_ssim_batch(ke_ctx* ctx, const uint8_t* d_bank, int h, int w, int c, int64_t img_stride, int64_t row_stride,
                  const int64_t* d_ia, const int64_t* d_ib, int64_t n_pairs, int gaussian, double* d_ssim, void* stream);

/* Host buffers: pair p compares h_a + p*h*w*c with h_b + p*h*w*c (c = 1: 'L' planes), copied in
 * chunks overlapped with the kernel (staged like ke_phash_batch_host), pairs split over the context's devices.
 * Behind the drop-in dup.refine._compute_ssim / refine_pairs_batch. */
int ke_ssim_pairs_host(ke_ctx* ctx, const uint8_t* h_a, const uint8_t* h_b, int64_t n_pairs, int h, int w, int c,
                       int gaussian, double* h_ssim);

/* convert("L") (Pillow rgb2l) of the bank images d_idx[0..n) into packed h*w planes: what a multi-process scan ships
 * between ranks for its cross-shard SSIM pairs instead of the RGB images (src/dup/refine.py:48-49). */
int ke_luma_planes(ke_ctx* ctx, const uint8_t* d_bank, int h, int w, int c, int64_t img_stride, int64_t row_stride,
                   const int64_t* d_idx, int64_t n, uint8_t* d_out, void* stream);

/* ---------------------------------------------------------------------------------------
 * N2 — the matching half of dup.refine._compute_orb_ratio (src/dup/refine.py:55-68):
 *     matches = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(da, db);  ratio = len(matches) / min(len(kpa), len(kpb))
 * batched over pairs.  d_desc holds 256-bit ORB descriptors (32 bytes each, 16-byte aligned); pair p matches the
 * d_cnt_a[p] query descriptors starting at row d_off_a[p] against the d_cnt_b[p] train descriptors at row d_off_b[p]
 * (max_a / max_b = the largest such counts).  A match is a MUTUAL nearest neighbour in Hamming distance, the first index
 * winning ties on either side — OpenCV's cross-check.  d_n_matches[p] = len(matches); d_match_train / d_match_dist
 * (nullable, [n_pairs][max_a]) = trainIdx / distance of query i or -1.  The FAST/Harris detector and the rBRIEF
 * descriptor stay with OpenCV on the host. */
int ke_orb_match_pairs(ke_ctx* ctx, const uint8_t* d_desc, const int64_t* d_off_a, const int32_t* d_cnt_a,
                       const int64_t* d_off_b, const int32_t* d_cnt_b, int64_t n_pairs, int max_a, int max_b,
                       int32_t* d_n_matches, int32_t* d_match_train, int32_t* d_match_dist, void* stream);

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
/* The same on the device for pairs of table indices < n_nodes (lock-free union-find, larger root hooked under the
 * smaller): d_label[v] = smallest index of v's component for every v that occurs in a pair, 0xFFFFFFFF otherwise. */
int ke_cluster_pairs(ke_ctx* ctx, const uint32_t* d_a, const uint32_t* d_b, int64_t n_pairs, int64_t n_nodes,
                     uint32_t* d_label, void* stream);

/* ---------------------------------------------------------------------------------------
 * N3 — table-level duplicate scan: DuplicateScanner.build_clusters (src/dup/scanner.py:211-356) on the
 * COLUMNS of `iter_files_for_dup` (src/db/repository.py:416-455) instead of one DuplicateFile object per row:
 * h_phash = the signed 64-bit `signatures.phash_u64` column (src/db/schema.py:65-72; same bits as the unsigned hash),
 * h_file_id (nullable: rows are then distinct files), h_size (nullable: no size gate).  Steps, all on the device(s):
 * band-bucket statistics and the KE_DUP_BUCKET_PAIR_CAP mask (:227-253; pair_cap <= 0: off), the all-pairs join with
 * the band predicate over every device of the context (:262-290), the same-id and size-ratio gates (:271-279,
 * :358-370; size_ratio <= 0: off), union-find + best_hamming (:304-318).
 * Out: the members (rows with at least one edge) grouped by component — h_member_index (table row), h_member_label
 * (= smallest row of the component), h_member_best (min edge distance), components by ascending label, rows ascending
 * inside; and (edge_capacity > 0) the surviving edges sorted by (i, j).  KE_E_CAPACITY when a buffer is too small
 * (stats->members / stats->edges then hold the required sizes).
 * Requires distinct file ids (the reference de-duplicates edges by id pair, which only matters when they repeat);
 * the cosine gate (:372-400) needs embeddings and stays with the caller. */
typedef struct ke_scan_stats {
    int64_t n_buckets, buckets_ge2, max_bucket; /* what the reference logs at :240-247 */
    int64_t candidates;                          /* pairs within the threshold that share an allowed band */
    int64_t after_same_id, edges;                /* after the same-id gate, after the size gate */
    int64_t members, clusters;
} ke_scan_stats;
int ke_scan_table_host(ke_ctx* ctx, const int64_t* h_phash, const int64_t* h_file_id, const int64_t* h_size, int64_t n,
                       int threshold, int band_bits, int band_count, double size_ratio, int64_t pair_cap,
                       int64_t* h_member_index, int64_t* h_member_label, int32_t* h_member_best, int64_t member_capacity,
                       uint32_t* h_edge_i, uint32_t* h_edge_j, uint8_t* h_edge_dist, int64_t edge_capacity,
                       ke_scan_stats* stats);

/* ---------------------------------------------------------------------------------------
 * Measurement helpers (bench.py / tests only). */

/* Synthetic image generator, the CUDA twin of kobato_b200.synth.synth_image (identical bytes): images
 * start, start + item_stride, ... of a set of n_set. */
int ke_synth_images(ke_ctx* ctx, uint8_t* d_out, int64_t start, int64_t item_stride, int64_t count, int h, int w, int c,
                    int64_t n_set, uint64_t seed, int planted_permille, void* stream);

/* POPC issue-rate microbenchmark: the integer roofline denominator of K2.
 * Returns POPC thread-instructions per SM clock per SM, and the SM clock it derived. */
int ke_microbench_popc(ke_ctx* ctx, int iters, double* popc_per_clk_per_sm, double* sm_clock_mhz);

/* Kernels launched by this library on this context since creation (bench.py `gpu_launches`). */
int64_t ke_ctx_launch_count(const ke_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* KOBATO_B200_H */
