// ke_scan.cu — N3: table-level duplicate scan ("DB wire format <-> device table"), sm_100a.
//
// Replaces DuplicateScanner.build_clusters (reference src/dup/scanner.py:211-356) for callers that hold the
// `files LEFT JOIN signatures` rows (src/db/repository.py:416-455; phash_u64 is a signed 64-bit INTEGER column,
// src/db/schema.py:65-72) as COLUMN ARRAYS instead of 10 M DuplicateFile objects:
//     signed-64 phash[n], file_id[n], size[n]
//       -> band bucket statistics + KE_DUP_BUCKET_PAIR_CAP mask           (:227-253; device histograms)
//       -> all-pairs Hamming join with the band predicate, all devices      (:262-290; K2, ke_join.cu)
//       -> same-id / size-ratio gates on the candidate list                 (:271-279, :358-370; device)
//       -> connected components + best_hamming per member                   (:304-318; device union-find)
//       -> members grouped by component, back on the host.
// Python then builds DuplicateCluster objects for the members only (keeper choice and the sorts, :320-356).
//
// Union-find on the device: parent[] over table indices, roots hooked larger-under-smaller with atomicCAS (parents only
// ever decrease, so concurrent finds with path halving stay correct), then every member is flattened to its root.  The
// label of a component is therefore its smallest table index, independent of scheduling.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

#include "ke_common.cuh"

namespace {

constexpr int kT = 256;
constexpr uint32_t kNoBest = 0xFFFFFFFFu;

__global__ void __launch_bounds__(kT) ke_band_hist_kernel(const uint64_t* __restrict__ h, long long n, int band_bits,
                                                          int band_count, uint32_t* __restrict__ hist) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    if (i >= n) return;
    const uint64_t v = h[i], mask = band_bits >= 64 ? ~0ull : ((1ull << band_bits) - 1ull);
    for (int b = 0; b < band_count; ++b) atomicAdd(&hist[((size_t)b << band_bits) + (size_t)((v >> (b * band_bits)) & mask)], 1u);
}

// stats[0] = buckets, stats[1] = buckets with >= 2 members, stats[2] = largest bucket
__global__ void __launch_bounds__(kT) ke_band_stats_kernel(const uint32_t* __restrict__ hist, long long entries,
                                                           unsigned long long* __restrict__ stats) {
    unsigned long long nb = 0, ge2 = 0, mx = 0;
    for (long long e = (long long)blockIdx.x * kT + threadIdx.x; e < entries; e += (long long)gridDim.x * kT) {
        const uint32_t c = hist[e];
        nb += c > 0, ge2 += c >= 2, mx = c > mx ? c : mx;
    }
    for (int off = 16; off; off >>= 1) {
        nb += __shfl_xor_sync(0xffffffffu, nb, off);
        ge2 += __shfl_xor_sync(0xffffffffu, ge2, off);
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, mx, off);
        mx = o > mx ? o : mx;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&stats[0], nb);
        atomicAdd(&stats[1], ge2);
        atomicMax(&stats[2], mx);
    }
}

// bit b of allow[i] = "my bucket in band b has at most pair_cap pairs" (reference :239-266 skips the others)
__global__ void __launch_bounds__(kT) ke_band_allow_kernel(const uint64_t* __restrict__ h, long long n, int band_bits,
                                                           int band_count, const uint32_t* __restrict__ hist,
                                                           long long pair_cap, uint64_t* __restrict__ allow) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    if (i >= n) return;
    const uint64_t v = h[i], mask = band_bits >= 64 ? ~0ull : ((1ull << band_bits) - 1ull);
    uint64_t a = 0;
    for (int b = 0; b < band_count; ++b) {
        const unsigned long long s = hist[((size_t)b << band_bits) + (size_t)((v >> (b * band_bits)) & mask)];
        if (s * (s - 1ull) / 2ull <= (unsigned long long)pair_cap) a |= 1ull << b;
    }
    allow[i] = a;
}

__device__ __forceinline__ uint32_t uf_find(uint32_t* parent, uint32_t x) {
    uint32_t cur = __ldcg(parent + x);
    if (cur != x) {
        uint32_t prev = x, next;
        while (cur > (next = __ldcg(parent + cur))) {
            parent[prev] = next;  // path halving: any ancestor is a valid parent, parents only decrease
            prev = cur;
            cur = next;
        }
    }
    return cur;
}

__global__ void __launch_bounds__(kT) ke_uf_init_kernel(uint32_t* parent, uint32_t* best, long long n) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    if (i < n) parent[i] = (uint32_t)i, best[i] = kNoBest;
}

// One thread per candidate: gates of the reference (:271-279), then union + best_hamming.  counters: [0] pairs with
// different ids, [1] after the size gate (= edges).  keep[e] = 1 for the edges.
__global__ void __launch_bounds__(kT) ke_scan_edges_kernel(const uint32_t* __restrict__ ci, const uint32_t* __restrict__ cj,
                                                           const uint8_t* __restrict__ cd, long long m,
                                                           const long long* __restrict__ file_id,
                                                           const long long* __restrict__ size, double size_ratio,
                                                           uint32_t* parent, uint32_t* best, uint8_t* __restrict__ keep,
                                                           unsigned long long* counters) {
    const long long e = (long long)blockIdx.x * kT + threadIdx.x;
    if (e >= m) return;
    const uint32_t i = ci[e], j = cj[e];
    uint8_t k = 0;
    if (!file_id || file_id[i] != file_id[j]) {
        bool pass = true;
        if (size && size_ratio > 0.0) {
            const long long ls = size[i], rs = size[j];
            if (ls > 0 && rs > 0) {  // `(min / max) >= ratio` on Python ints is a correctly rounded double division
                const double lo = (double)(ls < rs ? ls : rs), hi = (double)(ls < rs ? rs : ls);
                pass = __ddiv_rn(lo, hi) >= size_ratio;
            }
        }
        atomicAdd(&counters[0], 1ull);
        if (pass) {
            atomicAdd(&counters[1], 1ull);
            k = 1;
            const uint32_t d = cd[e];
            atomicMin(&best[i], d);
            atomicMin(&best[j], d);
            uint32_t ra = uf_find(parent, i), rb = uf_find(parent, j);
            while (ra != rb) {
                if (ra < rb) {
                    const uint32_t t = ra;
                    ra = rb, rb = t;
                }
                const uint32_t old = atomicCAS(parent + ra, ra, rb);  // hook the larger root under the smaller
                if (old == ra) break;
                ra = uf_find(parent, old);
                rb = uf_find(parent, rb);
            }
        }
    }
    keep[e] = k;
}

// members = table rows with at least one edge; out rows {index, label, best} appended in any order (sorted on the host)
__global__ void __launch_bounds__(kT) ke_scan_members_kernel(uint32_t* parent, const uint32_t* __restrict__ best, long long n,
                                                             uint32_t* __restrict__ out_idx, uint32_t* __restrict__ out_label,
                                                             uint32_t* __restrict__ out_best, long long capacity,
                                                             unsigned long long* count) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    if (i >= n) return;
    const uint32_t b = best[i];
    if (b == kNoBest) return;
    const uint32_t root = uf_find(parent, (uint32_t)i);
    const unsigned long long slot = atomicAdd(count, 1ull);
    if ((long long)slot < capacity) out_idx[slot] = (uint32_t)i, out_label[slot] = root, out_best[slot] = b;
}

__global__ void __launch_bounds__(kT) ke_uf_label_kernel(uint32_t* parent, const uint32_t* __restrict__ best, long long n,
                                                         uint32_t* __restrict__ label) {
    const long long i = (long long)blockIdx.x * kT + threadIdx.x;
    if (i < n) label[i] = best[i] == kNoBest ? 0xFFFFFFFFu : uf_find(parent, (uint32_t)i);
}

struct DevBuf {  // cudaMalloc'd scratch freed on scope exit
    void* p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    int alloc(size_t bytes) {
        KE_CUDA(cudaMalloc(&p, bytes ? bytes : 1));
        return KE_OK;
    }
    template <typename T>
    T* as() const {
        return (T*)p;
    }
};

// bucket statistics and the pair-cap mask on the host for band widths a device histogram cannot hold (> 24 bits)
void host_band_stats(const uint64_t* h, int64_t n, int band_bits, int band_count, int64_t pair_cap, int64_t stats[3],
                     uint64_t* allow) {
    std::vector<uint32_t> order((size_t)n);
    const uint64_t mask = band_bits >= 64 ? ~0ull : ((1ull << band_bits) - 1ull);
    stats[0] = stats[1] = stats[2] = 0;
    if (allow) std::fill(allow, allow + n, 0ull);
    for (int b = 0; b < band_count; ++b) {
        std::iota(order.begin(), order.end(), 0u);
        auto key = [&](uint32_t i) { return (h[i] >> (b * band_bits)) & mask; };
        std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return key(x) < key(y); });
        for (int64_t s = 0; s < n;) {
            int64_t e = s + 1;
            while (e < n && key(order[(size_t)e]) == key(order[(size_t)s])) ++e;
            const int64_t c = e - s;
            stats[0] += 1, stats[1] += c >= 2, stats[2] = std::max(stats[2], c);
            if (allow && c * (c - 1) / 2 <= pair_cap)
                for (int64_t q = s; q < e; ++q) allow[order[(size_t)q]] |= 1ull << b;
            s = e;
        }
    }
}

}  // namespace

extern "C" int ke_scan_table_host(ke_ctx* ctx, const int64_t* h_phash, const int64_t* h_file_id, const int64_t* h_size,
                                  int64_t n, int threshold, int band_bits, int band_count, double size_ratio,
                                  int64_t pair_cap, int64_t* h_member_index, int64_t* h_member_label,
                                  int32_t* h_member_best, int64_t member_capacity, uint32_t* h_edge_i, uint32_t* h_edge_j,
                                  uint8_t* h_edge_dist, int64_t edge_capacity, ke_scan_stats* stats) {
    KE_REQUIRE(ctx != nullptr && stats != nullptr, "ke_scan_table_host: NULL argument");
    memset(stats, 0, sizeof(*stats));
    KE_REQUIRE(n >= 0 && n <= 0xFFFFFFFFll, "ke_scan_table_host: n=%lld out of range", (long long)n);
    KE_REQUIRE(threshold >= 0 && threshold <= 64, "hamming_threshold must be in [0, 64]");
    KE_REQUIRE(band_bits > 0, "band_bits must be positive");
    KE_REQUIRE(band_count > 0, "band_count must be positive");
    KE_REQUIRE((long long)band_bits * band_count <= 64, "band config too large");
    KE_REQUIRE(member_capacity >= 0 && edge_capacity >= 0, "ke_scan_table_host: negative capacity");
    KE_REQUIRE(member_capacity == 0 || (h_member_index && h_member_label && h_member_best), "ke_scan_table_host: NULL member buffers");
    KE_REQUIRE(edge_capacity == 0 || (h_edge_i && h_edge_j && h_edge_dist), "ke_scan_table_host: NULL edge buffers");
    if (n < 2) return KE_OK;
    KE_REQUIRE(h_phash != nullptr, "ke_scan_table_host: h_phash is NULL");
    const uint64_t* h_hashes = reinterpret_cast<const uint64_t*>(h_phash);  // signed wrap = same bits (src/sig/phash.py:29-30)
    KeDeviceGuard guard(ctx->device);
    cudaStream_t s = ctx->copy_stream[0];
    int rc;

    // ---- bucket statistics (the reference logs them and stops early when no bucket has two members) + pair-cap mask
    std::vector<uint64_t> allow_host;
    const bool want_allow = pair_cap > 0;
    if (band_bits <= 24) {
        DevBuf d_h, d_hist, d_stats, d_allow;
        const long long entries = (long long)band_count << band_bits;
        if ((rc = d_h.alloc((size_t)n * 8)) || (rc = d_hist.alloc((size_t)entries * 4)) || (rc = d_stats.alloc(32))) return rc;
        if ((rc = ke_h2d_staged(ctx, d_h.p, h_hashes, (size_t)n * 8, s))) return rc;
        KE_CUDA(cudaMemsetAsync(d_hist.p, 0, (size_t)entries * 4, s));
        KE_CUDA(cudaMemsetAsync(d_stats.p, 0, 32, s));
        const unsigned grid_n = (unsigned)((n + kT - 1) / kT);
        ke_band_hist_kernel<<<grid_n, kT, 0, s>>>(d_h.as<uint64_t>(), n, band_bits, band_count, d_hist.as<uint32_t>());
        ke_band_stats_kernel<<<(unsigned)std::min<long long>((entries + kT - 1) / kT, 4096), kT, 0, s>>>(
            d_hist.as<uint32_t>(), entries, d_stats.as<unsigned long long>());
        ctx->launches += 2;
        if (want_allow) {
            if ((rc = d_allow.alloc((size_t)n * 8))) return rc;
            ke_band_allow_kernel<<<grid_n, kT, 0, s>>>(d_h.as<uint64_t>(), n, band_bits, band_count, d_hist.as<uint32_t>(),
                                                       pair_cap, d_allow.as<uint64_t>());
            ctx->launches++;
            allow_host.resize((size_t)n);
            KE_CUDA(cudaMemcpyAsync(allow_host.data(), d_allow.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
        }
        unsigned long long st[4] = {};
        KE_CUDA(cudaMemcpyAsync(st, d_stats.p, 24, cudaMemcpyDeviceToHost, s));
        KE_CUDA(cudaStreamSynchronize(s));
        KE_CUDA(cudaGetLastError());
        stats->n_buckets = (int64_t)st[0], stats->buckets_ge2 = (int64_t)st[1], stats->max_bucket = (int64_t)st[2];
    } else {
        int64_t st[3];
        if (want_allow) allow_host.resize((size_t)n);
        host_band_stats(h_hashes, n, band_bits, band_count, pair_cap, st, want_allow ? allow_host.data() : nullptr);
        stats->n_buckets = st[0], stats->buckets_ge2 = st[1], stats->max_bucket = st[2];
    }
    if (stats->buckets_ge2 == 0) return KE_OK;  // "no bucket has 2+ items -> edges=0" (reference :255-257)

    // ---- candidate search on every device of the context (K2); the lists grow until nothing is truncated
    std::vector<uint32_t> ci, cj;
    std::vector<uint8_t> cd;
    int64_t cap = std::max<int64_t>(1 << 20, 2 * n), found = 0;
    for (;;) {
        ci.resize((size_t)cap), cj.resize((size_t)cap), cd.resize((size_t)cap);
        rc = ke_hamming_join_host(ctx, h_hashes, n, threshold, KE_JOIN_REQUIRE_BAND, band_bits, band_count,
                                  want_allow ? allow_host.data() : nullptr, 0, 1, ci.data(), cj.data(), cd.data(), cap, &found);
        if (rc == KE_E_CAPACITY) {
            cap = found;
            continue;
        }
        if (rc) return rc;
        break;
    }
    stats->candidates = found;
    if (found == 0) return KE_OK;

    // ---- gates + union-find + members on the context's first device
    DevBuf d_ci, d_cj, d_cd, d_keep, d_id, d_size, d_parent, d_best, d_cnt, d_mi, d_ml, d_mb;
    if ((rc = d_ci.alloc((size_t)found * 4)) || (rc = d_cj.alloc((size_t)found * 4)) || (rc = d_cd.alloc((size_t)found)) ||
        (rc = d_keep.alloc((size_t)found)) || (rc = d_parent.alloc((size_t)n * 4)) || (rc = d_best.alloc((size_t)n * 4)) ||
        (rc = d_cnt.alloc(64)))
        return rc;
    KE_CUDA(cudaMemcpyAsync(d_ci.p, ci.data(), (size_t)found * 4, cudaMemcpyHostToDevice, s));
    KE_CUDA(cudaMemcpyAsync(d_cj.p, cj.data(), (size_t)found * 4, cudaMemcpyHostToDevice, s));
    KE_CUDA(cudaMemcpyAsync(d_cd.p, cd.data(), (size_t)found, cudaMemcpyHostToDevice, s));
    if (h_file_id) {
        if ((rc = d_id.alloc((size_t)n * 8))) return rc;
        if ((rc = ke_h2d_staged(ctx, d_id.p, h_file_id, (size_t)n * 8, s))) return rc;
    }
    const bool gate = h_size != nullptr && size_ratio > 0.0;
    if (gate) {
        if ((rc = d_size.alloc((size_t)n * 8))) return rc;
        if ((rc = ke_h2d_staged(ctx, d_size.p, h_size, (size_t)n * 8, s))) return rc;
    }
    KE_CUDA(cudaMemsetAsync(d_cnt.p, 0, 64, s));
    const unsigned grid_n = (unsigned)((n + kT - 1) / kT), grid_m = (unsigned)((found + kT - 1) / kT);
    ke_uf_init_kernel<<<grid_n, kT, 0, s>>>(d_parent.as<uint32_t>(), d_best.as<uint32_t>(), n);
    ke_scan_edges_kernel<<<grid_m, kT, 0, s>>>(d_ci.as<uint32_t>(), d_cj.as<uint32_t>(), d_cd.as<uint8_t>(), found,
                                               d_id.as<long long>(), gate ? d_size.as<long long>() : nullptr, size_ratio,
                                               d_parent.as<uint32_t>(), d_best.as<uint32_t>(), d_keep.as<uint8_t>(),
                                               d_cnt.as<unsigned long long>());
    const int64_t mcap = std::min<int64_t>(2 * found, n);
    if ((rc = d_mi.alloc((size_t)mcap * 4)) || (rc = d_ml.alloc((size_t)mcap * 4)) || (rc = d_mb.alloc((size_t)mcap * 4))) return rc;
    ke_scan_members_kernel<<<grid_n, kT, 0, s>>>(d_parent.as<uint32_t>(), d_best.as<uint32_t>(), n, d_mi.as<uint32_t>(),
                                                 d_ml.as<uint32_t>(), d_mb.as<uint32_t>(), mcap,
                                                 d_cnt.as<unsigned long long>() + 2);
    ctx->launches += 3;
    unsigned long long cnt[4] = {};
    KE_CUDA(cudaMemcpyAsync(cnt, d_cnt.p, 32, cudaMemcpyDeviceToHost, s));
    std::vector<uint8_t> keep((size_t)found);
    KE_CUDA(cudaMemcpyAsync(keep.data(), d_keep.p, (size_t)found, cudaMemcpyDeviceToHost, s));
    KE_CUDA(cudaStreamSynchronize(s));
    KE_CUDA(cudaGetLastError());
    stats->after_same_id = (int64_t)cnt[0];
    stats->edges = (int64_t)cnt[1];
    stats->members = (int64_t)cnt[2];

    // ---- edges (i < j table indices, sorted) for callers that verify them further (SSIM)
    if (edge_capacity > 0) {
        std::vector<uint64_t> keys;
        keys.reserve((size_t)stats->edges);
        for (int64_t e = 0; e < found; ++e)
            if (keep[(size_t)e]) keys.push_back(((uint64_t)ci[(size_t)e] << 40) | ((uint64_t)cj[(size_t)e] << 8) | cd[(size_t)e]);
        // i, j < 2^32 do not fit 40 + 32 + 8 bits together: sort (i, j) pairs through an index instead when n is large
        if (n <= (1ll << 24)) {
            std::sort(keys.begin(), keys.end());
            const int64_t take = std::min<int64_t>((int64_t)keys.size(), edge_capacity);
            for (int64_t e = 0; e < take; ++e) {
                h_edge_i[e] = (uint32_t)(keys[(size_t)e] >> 40);
                h_edge_j[e] = (uint32_t)((keys[(size_t)e] >> 8) & 0xFFFFFFFFull);
                h_edge_dist[e] = (uint8_t)(keys[(size_t)e] & 0xFF);
            }
        } else {
            std::vector<int64_t> order;
            order.reserve((size_t)stats->edges);
            for (int64_t e = 0; e < found; ++e)
                if (keep[(size_t)e]) order.push_back(e);
            std::sort(order.begin(), order.end(), [&](int64_t x, int64_t y) {
                return ci[(size_t)x] != ci[(size_t)y] ? ci[(size_t)x] < ci[(size_t)y] : cj[(size_t)x] < cj[(size_t)y];
            });
            const int64_t take = std::min<int64_t>((int64_t)order.size(), edge_capacity);
            for (int64_t e = 0; e < take; ++e) {
                h_edge_i[e] = ci[(size_t)order[(size_t)e]];
                h_edge_j[e] = cj[(size_t)order[(size_t)e]];
                h_edge_dist[e] = cd[(size_t)order[(size_t)e]];
            }
        }
    }

    // ---- members grouped by component (label = smallest table index), ascending inside each
    const int64_t m = stats->members;
    if (m > mcap) {
        ke_set_error("ke_scan_table_host: internal member buffer too small (%lld > %lld)", (long long)m, (long long)mcap);
        return KE_E_CAPACITY;
    }
    std::vector<uint32_t> mi((size_t)m), ml((size_t)m), mb((size_t)m);
    if (m) {
        KE_CUDA(cudaMemcpy(mi.data(), d_mi.p, (size_t)m * 4, cudaMemcpyDeviceToHost));
        KE_CUDA(cudaMemcpy(ml.data(), d_ml.p, (size_t)m * 4, cudaMemcpyDeviceToHost));
        KE_CUDA(cudaMemcpy(mb.data(), d_mb.p, (size_t)m * 4, cudaMemcpyDeviceToHost));
    }
    std::vector<uint64_t> key((size_t)m);
    for (int64_t q = 0; q < m; ++q) key[(size_t)q] = ((uint64_t)ml[(size_t)q] << 32) | mi[(size_t)q];
    std::vector<uint32_t> order((size_t)m);
    std::iota(order.begin(), order.end(), 0u);
    std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return key[x] < key[y]; });
    int64_t clusters = 0;
    for (int64_t q = 0; q < m; ++q) {
        const uint32_t o = order[(size_t)q];
        if (q == 0 || ml[o] != ml[order[(size_t)q - 1]]) ++clusters;
        if (q < member_capacity) {
            h_member_index[q] = mi[o];
            h_member_label[q] = ml[o];
            h_member_best[q] = (int32_t)mb[o];
        }
    }
    stats->clusters = clusters;
    if (m > member_capacity || stats->edges > (edge_capacity > 0 ? edge_capacity : stats->edges)) {
        ke_set_error("ke_scan_table_host: %lld members / %lld edges exceed the output capacity (%lld / %lld)", (long long)m,
                     (long long)stats->edges, (long long)member_capacity, (long long)edge_capacity);
        return KE_E_CAPACITY;
    }
    return KE_OK;
}

// Device union-find for callers that already hold accepted pairs (table indices < n_nodes): the device form of
// ke_cluster_pairs_host.  d_label[v] = smallest index of v's component for every v that occurs in a pair, 0xFFFFFFFF else.
extern "C" int ke_cluster_pairs(ke_ctx* ctx, const uint32_t* d_a, const uint32_t* d_b, int64_t n_pairs, int64_t n_nodes,
                                uint32_t* d_label, void* stream) {
    KE_REQUIRE(ctx != nullptr && n_pairs >= 0 && n_nodes >= 0 && n_nodes <= 0xFFFFFFFFll, "ke_cluster_pairs: bad arguments");
    if (n_nodes == 0) return KE_OK;
    KE_REQUIRE(d_label != nullptr && (n_pairs == 0 || (d_a && d_b)), "ke_cluster_pairs: NULL buffer");
    KeDeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t *parent = nullptr, *best = nullptr;
    uint8_t *zero_d = nullptr, *keep = nullptr;
    unsigned long long* cnt = nullptr;
    const size_t pbytes = (size_t)std::max<int64_t>(n_pairs, 1);
    KE_CUDA(cudaMallocAsync((void**)&parent, (size_t)n_nodes * 4, s));
    KE_CUDA(cudaMallocAsync((void**)&best, (size_t)n_nodes * 4, s));
    KE_CUDA(cudaMallocAsync((void**)&zero_d, pbytes, s));
    KE_CUDA(cudaMallocAsync((void**)&keep, pbytes, s));
    KE_CUDA(cudaMallocAsync((void**)&cnt, 64, s));
    KE_CUDA(cudaMemsetAsync(cnt, 0, 64, s));
    KE_CUDA(cudaMemsetAsync(zero_d, 0, pbytes, s));
    const unsigned grid_n = (unsigned)((n_nodes + kT - 1) / kT);
    ke_uf_init_kernel<<<grid_n, kT, 0, s>>>(parent, best, n_nodes);
    if (n_pairs)
        ke_scan_edges_kernel<<<(unsigned)((n_pairs + kT - 1) / kT), kT, 0, s>>>(d_a, d_b, zero_d, n_pairs, nullptr, nullptr, 0.0,
                                                                             parent, best, keep, cnt);
    ke_uf_label_kernel<<<grid_n, kT, 0, s>>>(parent, best, n_nodes, d_label);
    ctx->launches += n_pairs ? 3 : 2;
    KE_CUDA(cudaFreeAsync(parent, s));
    KE_CUDA(cudaFreeAsync(best, s));
    KE_CUDA(cudaFreeAsync(zero_d, s));
    KE_CUDA(cudaFreeAsync(keep, s));
    KE_CUDA(cudaFreeAsync(cnt, s));
    KE_CUDA(cudaGetLastError());
    return KE_OK;
}
