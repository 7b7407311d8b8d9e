// ke_join.cu — K2: all-pairs 64-bit Hamming threshold join (XOR + POPC), sm_100a.
//
// Replaces the candidate search of DuplicateScanner.build_clusters (reference
// src/dup/scanner.py:227-290, distance = src/sig/phash.py:60-63).
//
// Layout / schedule
//   * hashes: uint64[n] in HBM; 8 MB (1 M) .. 80 MB (10 M) -> L2 resident (126 MB), so the kernel
//     is integer-issue bound, not memory bound.
//   * The N x N upper triangle is cut into square tiles of TILE = 256*RPT hashes.  Tiles are
//     enumerated linearly (row-major over c >= r) and dealt round-robin:
//         tile t  ->  GPU  (t % part_count),  then CTA ((t / part_count) % gridDim.x)
//     A persistent grid of sm_count * CTAS_PER_SM CTAs walks its tiles.
//   * Per tile: each thread keeps RPT row hashes in registers (2*RPT b32), the CTA stages the
//     TILE column hashes in shared memory once, and every thread streams them as warp-broadcast
//     LDS.64 (all lanes read the same address: one wavefront).
//   * Per pair (fast path): 2 LOP3 (xor) + 2 POPC + 1 IADD3 (pl + ph - (T+1)) + 1/2 LOP3 (OR of
//     sign bits).  A check of the accumulated sign every CB columns x RPT rows enters the rare
//     slow path, which re-evaluates those pairs with bounds, i<j, the optional band predicate,
//     and appends (i, j, d) through one global atomic per hit.
//   * Bounding pipe: POPC (2 per pair).  Algorithmic work per pair = 6 integer instructions.
#include <algorithm>
#include <cstdlib>

#include "ke_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kColBatch = 4;  // columns per fast-path check

struct JoinArgs {
    const uint64_t* hashes;
    const uint64_t* band_allow;  // nullable
    long long n;
    int threshold;
    unsigned flags;
    int band_bits;
    int band_count;
    int part_index;
    int part_count;
    long long tiles_per_dim;
    long long tile_total;  // tiles_per_dim*(tiles_per_dim+1)/2
    uint32_t* out_i;
    uint32_t* out_j;
    uint8_t* out_d;
    long long capacity;
    unsigned long long* count;
    unsigned long long* queue;   // nullable: dynamic tile queue shared by concurrently running kernels
    const uint32_t* sliced;      // hybrid: bit-sliced copy of the table, [ceil(n/32)][64] words
};

__device__ __forceinline__ void tile_coords(long long t, long long nt, long long& r, long long& c) {
    // row r owns tiles [r*nt - r(r-1)/2, ...) of length nt - r
    double disc = (2.0 * nt + 1.0) * (2.0 * nt + 1.0) - 8.0 * (double)t;
    long long rr = (long long)(((2.0 * nt + 1.0) - sqrt(disc)) * 0.5);
    if (rr < 0) rr = 0;
    if (rr >= nt) rr = nt - 1;
    while (rr > 0 && rr * nt - rr * (rr - 1) / 2 > t) --rr;
    while ((rr + 1) * nt - (rr + 1) * rr / 2 <= t) ++rr;
    r = rr;
    c = rr + (t - (rr * nt - rr * (rr - 1) / 2));
}

__device__ __forceinline__ bool band_match(uint64_t x, const JoinArgs& a, uint64_t allow) {
    const uint64_t mask = a.band_bits >= 64 ? ~0ull : ((1ull << a.band_bits) - 1ull);
    for (int b = 0; b < a.band_count; ++b) {
        if ((((x >> (b * a.band_bits)) & mask) == 0ull) && ((allow >> b) & 1ull)) return true;
    }
    return false;
}

// Slow path of the POPC role, run by the WHOLE warp.  `flagged` = lanes whose 4-column batch holds a pair within the
// threshold in one of their rows (`rowmask`).  For every flagged lane the warp re-evaluates that lane's RPT x 4 pairs in
// one step — lane L takes (row L / 4, column L % 4) — with every filter (bounds, i < j, exact distance, band predicate,
// allow mask) and appends (i, j, d) through one global atomic per hit.  A flagged lane used to walk its own pairs
// serially while the other 31 waited; on tables of real pHashes (about one near pair per thousand, most of them
// rejected by the band predicate) that serial walk was a third of the kernel's time.
template <int RPT>
__device__ __noinline__ void emit_hits_warp(const JoinArgs& a, long long row0, const uint64_t* cols, long long col0,
                                            int jbase, uint32_t rowmask, uint32_t flagged) {
    const int lane = threadIdx.x & 31;
    const int k = lane >> 2, jj = lane & 3;
    const uint64_t colv = cols[jbase + jj];
    const long long j = col0 + jbase + jj;
    for (uint32_t f = flagged; f; f &= f - 1u) {
        const int src = __ffs(f) - 1;
        const uint32_t rm = __shfl_sync(0xffffffffu, rowmask, src);
        if (k >= RPT || !((rm >> k) & 1u)) continue;
        const long long i = row0 + (threadIdx.x & ~31) + src + (long long)k * kThreads;
        if (i >= a.n || j >= a.n || j <= i) continue;
        const uint64_t x = __ldg(a.hashes + i) ^ colv;
        const int d = __popcll(x);
        if (d > a.threshold) continue;
        if (a.flags & KE_JOIN_REQUIRE_BAND) {
            const uint64_t allow = a.band_allow ? (a.band_allow[i] & a.band_allow[j]) : ~0ull;
            if (!band_match(x, a, allow)) continue;
        }
        const unsigned long long slot = atomicAdd(a.count, 1ull);
        if ((long long)slot < a.capacity) {
            a.out_i[slot] = (uint32_t)i;
            a.out_j[slot] = (uint32_t)j;
            a.out_d[slot] = (uint8_t)d;
        }
    }
}

// bar.sync on a named barrier: lets two roles inside one CTA synchronise independently
template <int ID, int N>
__device__ __forceinline__ void role_sync() {
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory");
}

template <int RPT, int BAR>
__device__ __forceinline__ void popc_role(const JoinArgs& a, uint64_t* cols, long long* s_next_p) {
    constexpr int TILE = kThreads * RPT;
    const int neg_t1 = -(a.threshold + 1);
    long long& s_next = *s_next_p;

    for (long long local = blockIdx.x;; local += gridDim.x) {
        role_sync<BAR, kThreads>();  // previous tile's readers are done with cols[] (and with s_next)
        if (a.queue) {
            if (threadIdx.x == 0) s_next = (long long)atomicAdd(a.queue, 1ull);
            role_sync<BAR, kThreads>();
            local = s_next;
        }
        const long long t = local * a.part_count + a.part_index;
        if (t >= a.tile_total) break;
        if (a.queue && threadIdx.x == 0) atomicAdd(a.queue + 1, 1ull);
        long long tr, tc;
        tile_coords(t, a.tiles_per_dim, tr, tc);
        const long long row0 = tr * TILE, col0 = tc * TILE;

#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const long long j = col0 + threadIdx.x + (long long)k * kThreads;
            cols[threadIdx.x + k * kThreads] = j < a.n ? __ldg(a.hashes + j) : ~0ull;  // pad: distance 64 from row pad 0
        }
        uint32_t al[RPT], ah[RPT];
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const long long i = row0 + threadIdx.x + (long long)k * kThreads;
            const uint64_t row = i < a.n ? __ldg(a.hashes + i) : 0ull;
            al[k] = (uint32_t)row;
            ah[k] = (uint32_t)(row >> 32);
        }
        role_sync<BAR, kThreads>();

        const long long col_valid = a.n - col0 < TILE ? a.n - col0 : TILE;
        const int jend = (int)((col_valid + kColBatch - 1) / kColBatch) * kColBatch;
#pragma unroll 1
        for (int j = 0; j < jend; j += kColBatch) {
            // one sign accumulator per ROW (same LOP3 count as a single one: each 3-input OR folds two pairs), so that
            // the slow path re-checks only the flagged rows' 4 pairs instead of all RPT x 4 — on tables with many near
            // pairs (real pHashes) the slow path otherwise dominates
            int acc[RPT];
#pragma unroll
            for (int k = 0; k < RPT; ++k) acc[k] = 0;
#pragma unroll
            for (int jj = 0; jj < kColBatch; ++jj) {
                const uint2 b = *reinterpret_cast<const uint2*>(&cols[j + jj]);
#pragma unroll
                for (int k = 0; k < RPT; ++k) {
                    // sign bit of (popc_lo + popc_hi - (T+1)) is set iff the pair is within T
                    acc[k] |= __popc(al[k] ^ b.x) + __popc(ah[k] ^ b.y) + neg_t1;
                }
            }
            int any = 0;
#pragma unroll
            for (int k = 0; k < RPT; ++k) any |= acc[k];
            const uint32_t flagged = __ballot_sync(0xffffffffu, any < 0);  // the j loop is warp-uniform: all lanes are here
            if (flagged) {
                uint32_t rowmask = 0u;
#pragma unroll
                for (int k = 0; k < RPT; ++k) rowmask |= ((uint32_t)acc[k] >> 31) << k;
                emit_hits_warp<RPT>(a, row0, cols, col0, j, rowmask, flagged);
            }
        }
    }
}


template <int RPT>
__global__ void __launch_bounds__(kThreads) ke_join_kernel(const JoinArgs a) {
    __shared__ __align__(16) uint64_t cols[kThreads * RPT];
    __shared__ long long s_next;
    popc_role<RPT, 0>(a, cols, &s_next);
}

// ------------------------------------------------------------------------------------------
// Bit-sliced partner kernel (hybrid mode).  The POPC kernel saturates the XU pipe (16 POPC
// lanes/clk/SM) and leaves most of the ALU pipe idle; this kernel computes the same distances
// with LOP3 only, so the two run CONCURRENTLY on every SM and pull tiles from one queue.
//
// The table is also kept bit-sliced: for every block of 32 hashes, word k holds bit k of the 32
// hashes.  A thread owns one row hash a and its 64 masks M_k = -(bit k of a); for a block of 32
// columns, x_k = W_k ^ M_k has bit j set iff column j differs from a in bit k, so the per-column
// distance is the bitwise population count over the 64 words x_k.  A Harley-Seal carry-save tree
// (60 CSAs = 120 LOP3) reduces them to ones/twos/fours/eights planes plus four "sixteen" planes;
// for a threshold <= 15 a column hits iff no sixteen plane is set and the 4-bit value <= T.
// ~200 LOP3 per 32 pairs and lane = 6.3 ALU instructions per pair, no POPC at all.

constexpr int kBThreads = 128;
constexpr int kBTile = 2048;  // the fused kernel's tile for large tables (= the POPC role's with RPT = 8); 1024 for mid sizes

__global__ void __launch_bounds__(256) ke_bitslice_kernel(const uint64_t* __restrict__ hashes, long long n,
                                                          uint32_t* __restrict__ sliced) {
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long nblk = (n + 31) / 32;
    if (warp_global >= nblk) return;
    const long long i = warp_global * 32 + lane;
    const uint64_t h = i < n ? hashes[i] : 0ull;
    uint32_t mine_lo = 0, mine_hi = 0;  // lane k keeps words k and k+32
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        const uint32_t wlo = __ballot_sync(0xffffffffu, (h >> k) & 1ull);
        const uint32_t whi = __ballot_sync(0xffffffffu, (h >> (k + 32)) & 1ull);
        if (lane == k) mine_lo = wlo, mine_hi = whi;
    }
    sliced[warp_global * 64 + lane] = mine_lo;
    sliced[warp_global * 64 + 32 + lane] = mine_hi;
}

__device__ __forceinline__ void csa(uint32_t& h, uint32_t& l, uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t u = a ^ b;
    h = (a & b) | (u & c);
    l = u ^ c;
}

__device__ __noinline__ void emit_sliced_hits(const JoinArgs& a, uint64_t row, long long i, long long colbase, uint32_t hits) {
    while (hits) {
        const int j = __ffs(hits) - 1;
        hits &= hits - 1;
        const long long col = colbase + j;
        if (i >= a.n || col >= a.n || col <= i) continue;
        const uint64_t x = row ^ __ldg(a.hashes + col);
        const int d = __popcll(x);
        if (d > a.threshold) continue;
        if (a.flags & KE_JOIN_REQUIRE_BAND) {
            const uint64_t allow = a.band_allow ? (a.band_allow[i] & a.band_allow[col]) : ~0ull;
            if (!band_match(x, a, allow)) continue;
        }
        const unsigned long long slot = atomicAdd(a.count, 1ull);
        if ((long long)slot < a.capacity) {
            a.out_i[slot] = (uint32_t)i;
            a.out_j[slot] = (uint32_t)col;
            a.out_d[slot] = (uint8_t)d;
        }
    }
}

template <int BAR, int TILE>
__device__ __forceinline__ void sliced_role(const JoinArgs& a, uint32_t* cols, long long* s_next_p, const int tid) {
    long long& s_next = *s_next_p;
    const long long nblk_total = (a.n + 31) / 32;
    // 4-bit compare planes for the uniform threshold
    uint32_t tb[4], ntb[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        tb[b] = ((a.threshold >> b) & 1) ? 0xFFFFFFFFu : 0u;
        ntb[b] = ~tb[b];
    }

    for (long long local = blockIdx.x;; local += gridDim.x) {
        role_sync<BAR, kBThreads>();
        if (a.queue) {
            if (tid == 0) s_next = (long long)atomicAdd(a.queue, 1ull);
            role_sync<BAR, kBThreads>();
            local = s_next;
        }
        const long long t = local * a.part_count + a.part_index;
        if (t >= a.tile_total) break;
        if (a.queue && tid == 0) atomicAdd(a.queue + 2, 1ull);
        long long tr, tc;
        tile_coords(t, a.tiles_per_dim, tr, tc);
        const long long row0 = tr * TILE, col0 = tc * TILE;
        const long long blk0 = col0 / 32;
        const int nblk = (int)min((long long)(TILE / 32), nblk_total - blk0);
        {
            const uint4* src = reinterpret_cast<const uint4*>(a.sliced + blk0 * 64);
            uint4* dst = reinterpret_cast<uint4*>(cols);
            for (int q = tid; q < nblk * 16; q += kBThreads) dst[q] = __ldg(src + q);
        }
        role_sync<BAR, kBThreads>();

        for (int pass = 0; pass < TILE / kBThreads; ++pass) {
            const long long i = row0 + (long long)pass * kBThreads + tid;
            if (row0 + (long long)pass * kBThreads >= a.n) break;
            const uint64_t row = i < a.n ? __ldg(a.hashes + i) : 0ull;
            const uint32_t rlo = (uint32_t)row, rhi = (uint32_t)(row >> 32);
            uint32_t M[64];
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                M[k] = (uint32_t)((int32_t)(rlo << (31 - k)) >> 31);
                M[k + 32] = (uint32_t)((int32_t)(rhi << (31 - k)) >> 31);
            }
#pragma unroll 1
            for (int blk = 0; blk < nblk; ++blk) {
                const uint4* w4 = reinterpret_cast<const uint4*>(cols + blk * 64);
                uint32_t ones = 0, twos = 0, fours = 0, eights = 0, big = 0;
#pragma unroll
                for (int g = 0; g < 4; ++g) {  // 16 inputs per round
                    const uint4 wa = w4[g * 4], wb = w4[g * 4 + 1], wc = w4[g * 4 + 2], wd = w4[g * 4 + 3];
                    const int k = g * 16;
                    uint32_t twosA, twosB, foursA, foursB, eightsA, eightsB, sixteen;
                    csa(twosA, ones, ones, wa.x ^ M[k + 0], wa.y ^ M[k + 1]);
                    csa(twosB, ones, ones, wa.z ^ M[k + 2], wa.w ^ M[k + 3]);
                    csa(foursA, twos, twos, twosA, twosB);
                    csa(twosA, ones, ones, wb.x ^ M[k + 4], wb.y ^ M[k + 5]);
                    csa(twosB, ones, ones, wb.z ^ M[k + 6], wb.w ^ M[k + 7]);
                    csa(foursB, twos, twos, twosA, twosB);
                    csa(eightsA, fours, fours, foursA, foursB);
                    csa(twosA, ones, ones, wc.x ^ M[k + 8], wc.y ^ M[k + 9]);
                    csa(twosB, ones, ones, wc.z ^ M[k + 10], wc.w ^ M[k + 11]);
                    csa(foursA, twos, twos, twosA, twosB);
                    csa(twosA, ones, ones, wd.x ^ M[k + 12], wd.y ^ M[k + 13]);
                    csa(twosB, ones, ones, wd.z ^ M[k + 14], wd.w ^ M[k + 15]);
                    csa(foursB, twos, twos, twosA, twosB);
                    csa(eightsB, fours, fours, foursA, foursB);
                    csa(sixteen, eights, eights, eightsA, eightsB);
                    big |= sixteen;  // any weight-16 carry means distance >= 16 > T
                }
                // value = 8*eights + 4*fours + 2*twos + ones; hit iff !big && value <= T
                const uint32_t v[4] = {ones, twos, fours, eights};
                uint32_t gt = 0u, eq = 0xFFFFFFFFu;
#pragma unroll
                for (int b = 3; b >= 0; --b) {
                    gt |= eq & v[b] & ntb[b];
                    eq &= ~(v[b] ^ tb[b]);
                }
                const uint32_t hits = ~(gt | big);
                if (hits) emit_sliced_hits(a, row, i, col0 + (long long)blk * 32, hits);
            }
        }
    }
}

__global__ void __launch_bounds__(kBThreads) ke_join_sliced_kernel(const JoinArgs a) {
    __shared__ __align__(16) uint32_t cols[(kBTile / 32) * 64];  // 16 KB: 64 blocks x 64 words
    __shared__ long long s_next;
    sliced_role<0, kBTile>(a, cols, &s_next, threadIdx.x);
}

// Fused hybrid: warps 0..7 run the POPC role, warps 8..11 the bit-sliced role, each with its own
// column tile and named barrier, all pulling tiles from one queue.  Co-residency of the two
// instruction mixes on every SM is then guaranteed by construction (two separate kernels on two
// streams only co-run when the block scheduler happens to interleave them).
// RPT = 8: tiles of 2048 hashes; RPT = 4: tiles of 1024 for mid-size tables, where 2048-hash tiles leave a CTA only a
// handful of tiles and the roles' different tile times show as a tail (a rank's share of the 560 k table of the 8-GPU
// step: 16 tiles per CTA; the 70 k table of the 1-GPU step: 2).
template <int RPT>
__global__ void __launch_bounds__(kThreads + kBThreads, 2) ke_join_fused_kernel(const JoinArgs a) {
    constexpr int TILE = kThreads * RPT;
    __shared__ __align__(16) uint64_t cols_p[TILE];
    __shared__ __align__(16) uint32_t cols_b[(TILE / 32) * 64];
    __shared__ long long s_next[2];
    if (threadIdx.x < kThreads) popc_role<RPT, 1>(a, cols_p, &s_next[0]);
    else sliced_role<2, TILE>(a, cols_b, &s_next[1], threadIdx.x - kThreads);
}

template <int RPT>
int launch_join(ke_ctx* ctx, JoinArgs& a, cudaStream_t stream) {
    constexpr int TILE = kThreads * RPT;
    a.tiles_per_dim = (a.n + TILE - 1) / TILE;
    a.tile_total = a.tiles_per_dim * (a.tiles_per_dim + 1) / 2;
    int per_sm = 0;
    KE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ke_join_kernel<RPT>, kThreads, 0));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    long long mine = (a.tile_total - a.part_index + a.part_count - 1) / a.part_count;
    long long grid = (long long)ctx->sm_count * per_sm;
    if (grid > mine) grid = mine;
    if (grid < 1) return KE_OK;
    ke_join_kernel<RPT><<<(unsigned)grid, kThreads, 0, stream>>>(a);
    ctx->launches++;
    KE_CUDA(cudaGetLastError());
    return KE_OK;
}


// Hybrid launch: bit-slice the table, then the fused kernel (or the bit-sliced kernel alone) with a
// dynamic tile queue.
int launch_hybrid(ke_ctx* ctx, JoinArgs& a, cudaStream_t stream, int mode, int tile) {
    const int TILE = mode == 3 ? kBTile : tile;
    a.tiles_per_dim = (a.n + TILE - 1) / TILE;
    a.tile_total = a.tiles_per_dim * (a.tiles_per_dim + 1) / 2;
    const long long nblk = (a.n + 31) / 32;
    // The bit-sliced table and the tile queue belong to THIS call: stream-ordered allocations, so that two joins in
    // flight on one context (different streams / threads) never share them.
    void *d_sliced = nullptr, *d_queue = nullptr;
    KE_CUDA(cudaMallocAsync(&d_sliced, (size_t)nblk * 256 + 256, stream));
    if (cudaMallocAsync(&d_queue, 64, stream) != cudaSuccess) {
        cudaFreeAsync(d_sliced, stream);
        ke_set_error("ke_hamming_join: cudaMallocAsync failed");
        return KE_E_NOMEM;
    }
    a.sliced = (const uint32_t*)d_sliced;
    a.queue = (unsigned long long*)d_queue;
    KE_CUDA(cudaMemsetAsync(d_queue, 0, 32, stream));
    ke_bitslice_kernel<<<(unsigned)((nblk * 32 + 255) / 256), 256, 0, stream>>>(a.hashes, a.n, (uint32_t*)d_sliced);
    ctx->launches++;
    KE_CUDA(cudaGetLastError());
    const long long mine = (a.tile_total - a.part_index + a.part_count - 1) / a.part_count;
    if (mode == 3) {
        long long grid = std::min<long long>((long long)ctx->sm_count * 4, mine);
        ke_join_sliced_kernel<<<(unsigned)grid, kBThreads, 0, stream>>>(a);
    } else {
        long long grid = std::min<long long>((long long)ctx->sm_count * 2, mine);
        if (TILE == kBTile) ke_join_fused_kernel<8><<<(unsigned)grid, kThreads + kBThreads, 0, stream>>>(a);
        else ke_join_fused_kernel<4><<<(unsigned)grid, kThreads + kBThreads, 0, stream>>>(a);
    }
    ctx->launches++;
    KE_CUDA(cudaGetLastError());
#ifdef KE_TUNING_PROBES
    if (getenv("KE_JOIN_DEBUG")) {  // tuning probe: how many tiles each role took
        unsigned long long q[4] = {0, 0, 0, 0};
        KE_CUDA(cudaStreamSynchronize(stream));
        KE_CUDA(cudaMemcpy(q, d_queue, 32, cudaMemcpyDeviceToHost));
        fprintf(stderr, "[ke_join] tiles=%lld popc=%llu sliced=%llu\n", a.tile_total, q[1], q[2]);
    }
#endif
    KE_CUDA(cudaFreeAsync(d_sliced, stream));
    KE_CUDA(cudaFreeAsync(d_queue, stream));
    return KE_OK;
}

int pick_rpt(const ke_ctx* ctx, long long n, int part_count) {
    // Largest tile that still leaves every CTA >= 8 tiles (tail balance), smallest otherwise.
    const int opts[3] = {8, 4, 2};
    for (int rpt : opts) {
        long long tile = (long long)kThreads * rpt;
        long long nt = (n + tile - 1) / tile;
        long long tiles = nt * (nt + 1) / 2 / part_count;
        if (tiles >= (long long)ctx->sm_count * 4 * 8) return rpt;
    }
    return 2;
}

}  // namespace

extern "C" int64_t ke_hamming_join_pairs(int64_t n, int part_index, int part_count) {
    // Pairs are attributed to parts by tile; for reporting, an even split of the triangle.
    if (n < 2 || part_count <= 0 || part_index < 0 || part_index >= part_count) return 0;
    const int64_t total = n * (n - 1) / 2;
    return total / part_count + (part_index < total % part_count ? 1 : 0);
}

extern "C" int ke_hamming_join(ke_ctx* ctx, const uint64_t* d_hashes, int64_t n, int threshold, uint32_t flags,
                               int band_bits, int band_count, const uint64_t* d_band_allow, int part_index,
                               int part_count, uint32_t* d_out_i, uint32_t* d_out_j, uint8_t* d_out_dist,
                               int64_t capacity, unsigned long long* d_count, void* stream) {
    KE_REQUIRE(ctx != nullptr, "ke_hamming_join: ctx is NULL");
    KE_REQUIRE(n >= 0 && n <= 0xFFFFFFFFll, "ke_hamming_join: n=%lld out of range", (long long)n);
    KE_REQUIRE(threshold >= 0 && threshold <= 64, "hamming_threshold must be in [0, 64]");
    KE_REQUIRE(part_count >= 1 && part_index >= 0 && part_index < part_count, "ke_hamming_join: bad partition %d/%d",
               part_index, part_count);
    KE_REQUIRE(capacity >= 0 && d_count != nullptr, "ke_hamming_join: bad output arguments");
    KE_REQUIRE(capacity == 0 || (d_out_i && d_out_j && d_out_dist), "ke_hamming_join: NULL output buffers");
    if (flags & KE_JOIN_REQUIRE_BAND) {
        KE_REQUIRE(band_bits > 0, "band_bits must be positive");
        KE_REQUIRE(band_count > 0, "band_count must be positive");
        KE_REQUIRE((long long)band_bits * band_count <= 64, "band config too large");
    }
    KeDeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    KE_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), s));
    if (n < 2) return KE_OK;
    KE_REQUIRE(d_hashes != nullptr, "ke_hamming_join: d_hashes is NULL");
    JoinArgs a;
    a.hashes = d_hashes;
    a.band_allow = (flags & KE_JOIN_REQUIRE_BAND) ? d_band_allow : nullptr;
    a.n = n;
    a.threshold = threshold;
    a.flags = flags;
    a.band_bits = band_bits;
    a.band_count = band_count;
    a.part_index = part_index;
    a.part_count = part_count;
    a.out_i = d_out_i;
    a.out_j = d_out_j;
    a.out_d = d_out_dist;
    a.capacity = capacity;
    a.count = d_count;
    a.queue = nullptr;
    a.sliced = nullptr;
    const int rpt = pick_rpt(ctx, n, part_count);
    const int mode = ctx->join_mode;
    // hybrid (POPC role + bit-sliced LOP3 role) as soon as every CTA of the fused kernel gets a 2048-hash tile or more:
    // measured on B200 it is at least as fast as the POPC kernel from 70 k hashes on (1.14 vs 1.23 ms) and 28 % faster at
    // 140 k.  (Before the slow path re-checked flagged ROWS only, the pHashes of a real scan — many near pairs that the band
    // predicate rejects — made this switch 8x slower at 70 k; `tools/probe_join70k.py` is the regression probe.)
    const long long bt = (n + kBTile - 1) / kBTile, btiles = bt * (bt + 1) / 2 / part_count;
    const bool hybrid_auto = rpt == 8 || btiles >= 2ll * ctx->sm_count;
    // fewer than 12 tiles of 2048 per CTA of the fused kernel: tiles of 1024 (four times as many) even out the tail
    // (measured, tools/probe_join_sizes.py: 70 k hashes 1.16 -> 0.98 ms, a rank's share of 140 k / 280 k 2.05 -> 1.83 /
    // 3.67 -> 3.53 ms; from 16 tiles per CTA on — a share of 560 k — the larger tile wins again)
    const int tile = btiles >= 12ll * 2 * ctx->sm_count ? kBTile : kBTile / 2;
    if (threshold <= 15 && mode != 1 && (mode >= 2 || hybrid_auto))
        return launch_hybrid(ctx, a, s, mode == 0 ? 2 : mode, tile);
    switch (rpt) {
        case 8: return launch_join<8>(ctx, a, s);
        case 4: return launch_join<4>(ctx, a, s);
        default: return launch_join<2>(ctx, a, s);
    }
}

// Single-device body of ke_hamming_join_host (ke_multi.cu fans it over the devices of a context): table up, join,
// count back; the candidate lists stay in the context's scratch (*d_i, *d_j, *d_d) for the caller to copy out.
int ke_hamming_join_host_one(ke_ctx* ctx, const uint64_t* h_hashes, int64_t n, int threshold, uint32_t flags, int band_bits,
                             int band_count, const uint64_t* h_band_allow, int part_index, int part_count,
                             uint32_t** d_i_out, uint32_t** d_j_out, uint8_t** d_d_out, int64_t capacity,
                             int64_t* out_count) {
    *out_count = 0;
    KeDeviceGuard guard(ctx->device);
    cudaStream_t s = ctx->copy_stream[0];
    void *d_h = nullptr, *d_allow = nullptr, *d_i = nullptr, *d_j = nullptr, *d_d = nullptr, *d_cnt = nullptr;
    int rc;
    if ((rc = ke_ctx_scratch(ctx, 0, (size_t)n * 8 + 8, &d_h))) return rc;
    if ((rc = ke_ctx_scratch(ctx, 1, (size_t)capacity * 4 + 8, &d_i))) return rc;
    if ((rc = ke_ctx_scratch(ctx, 2, (size_t)capacity * 4 + 8, &d_j))) return rc;
    if ((rc = ke_ctx_scratch(ctx, 3, (size_t)capacity + 8, &d_d))) return rc;
    if ((rc = ke_ctx_scratch(ctx, 4, 64, &d_cnt))) return rc;
    if (n && (rc = ke_h2d_staged(ctx, d_h, h_hashes, (size_t)n * 8, s))) return rc;
    if (h_band_allow && (flags & KE_JOIN_REQUIRE_BAND) && n) {
        if ((rc = ke_ctx_scratch(ctx, 5, (size_t)n * 8 + 8, &d_allow))) return rc;
        if ((rc = ke_h2d_staged(ctx, d_allow, h_band_allow, (size_t)n * 8, s))) return rc;
    }
    rc = ke_hamming_join(ctx, (const uint64_t*)d_h, n, threshold, flags, band_bits, band_count,
                         (const uint64_t*)d_allow, part_index, part_count, (uint32_t*)d_i, (uint32_t*)d_j,
                         (uint8_t*)d_d, capacity, (unsigned long long*)d_cnt, s);
    if (rc) return rc;
    unsigned long long cnt = 0;
    KE_CUDA(cudaMemcpyAsync(&cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, s));
    KE_CUDA(cudaStreamSynchronize(s));
    *out_count = (int64_t)cnt;
    *d_i_out = (uint32_t*)d_i;
    *d_j_out = (uint32_t*)d_j;
    *d_d_out = (uint8_t*)d_d;
    return KE_OK;
}

// ------------------------------------------------------------------------------------------
// POPC issue-rate microbenchmark: 8 independent popc chains per thread, enough warps to
// saturate the pipe.  Reports thread-level POPC per SM clock per SM.

__global__ void __launch_bounds__(256) ke_popc_bench_kernel(int iters, uint32_t seed, uint32_t* sink,
                                                            long long* clocks) {
    uint32_t x0 = seed + threadIdx.x, x1 = x0 * 3u, x2 = x0 * 5u, x3 = x0 * 7u, x4 = x0 * 11u, x5 = x0 * 13u,
             x6 = x0 * 17u, x7 = x0 * 19u;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = __popc(x0) + seed;
            x1 = __popc(x1) + seed;
            x2 = __popc(x2) + seed;
            x3 = __popc(x3) + seed;
            x4 = __popc(x4) + seed;
            x5 = __popc(x5) + seed;
            x6 = __popc(x6) + seed;
            x7 = __popc(x7) + seed;
        }
    }
    const long long t1 = clock64();
    sink[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
    if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}

extern "C" int ke_microbench_popc(ke_ctx* ctx, int iters, double* popc_per_clk_per_sm, double* sm_clock_mhz) {
    KE_REQUIRE(ctx && iters > 0 && popc_per_clk_per_sm, "ke_microbench_popc: bad arguments");
    KeDeviceGuard guard(ctx->device);
    const int ctas_per_sm = 4, threads = 256;
    const int grid = ctx->sm_count * ctas_per_sm;
    void *d_sink = nullptr, *d_clk = nullptr;
    int rc;
    if ((rc = ke_ctx_scratch(ctx, 6, (size_t)grid * threads * 4, &d_sink))) return rc;
    if ((rc = ke_ctx_scratch(ctx, 7, (size_t)grid * 8, &d_clk))) return rc;
    cudaEvent_t e0, e1;
    KE_CUDA(cudaEventCreate(&e0));
    KE_CUDA(cudaEventCreate(&e1));
    ke_popc_bench_kernel<<<grid, threads>>>(iters / 8 + 1, 12345u, (uint32_t*)d_sink, (long long*)d_clk);  // warm
    KE_CUDA(cudaEventRecord(e0));
    ke_popc_bench_kernel<<<grid, threads>>>(iters, 12345u, (uint32_t*)d_sink, (long long*)d_clk);
    KE_CUDA(cudaEventRecord(e1));
    KE_CUDA(cudaEventSynchronize(e1));
    ctx->launches += 2;
    float ms = 0.f;
    KE_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> clk((size_t)grid);
    KE_CUDA(cudaMemcpy(clk.data(), d_clk, (size_t)grid * 8, cudaMemcpyDeviceToHost));
    long long worst = 0;
    for (long long c : clk) worst = c > worst ? c : worst;
    const double popc_per_sm = (double)iters * 64.0 * threads * ctas_per_sm;
    *popc_per_clk_per_sm = popc_per_sm / (double)worst;
    if (sm_clock_mhz) *sm_clock_mhz = (double)worst / (ms * 1e-3) / 1e6;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return KE_OK;
}
