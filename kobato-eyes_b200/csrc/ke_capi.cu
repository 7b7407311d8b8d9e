// ke_capi.cu — context lifetime, error reporting, host-side Pillow coefficient tables.
#include <cmath>
#include <cstring>

#include <algorithm>
#include <vector>

#include "ke_common.cuh"

static thread_local char g_err[512] = "";

void ke_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int ke_abi_version(void) { return KE_ABI_VERSION; }
extern "C" const char* ke_last_error(void) { return g_err; }

const char* ke_last_error_cstr() { return g_err; }

namespace {

int ctx_create_one(int device, ke_ctx** out) {
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        ke_set_error("ke_ctx_create: no CUDA device (%s); this library has no CPU fallback",
                     e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return KE_E_CUDA;
    }
    KE_REQUIRE(device >= 0 && device < count, "ke_ctx_create: device %d out of range [0,%d)", device, count);
    KeDeviceGuard guard(device);
    cudaDeviceProp prop;
    KE_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        ke_set_error("ke_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                     prop.minor);
        return KE_E_UNSUPPORTED;
    }
    ke_ctx* ctx = new (std::nothrow) ke_ctx();
    if (!ctx) return KE_E_NOMEM;
    ctx->device = device;
    ctx->dev_ctx[0] = ctx;
    ctx->sm_count = prop.multiProcessorCount;
    for (auto& s : ctx->copy_stream) KE_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    KE_CUDA(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
    for (auto& ev : ctx->ev) KE_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto& ev : ctx->stage_ev) KE_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    {
        // per-call scratch (join queue / bit-sliced table, SSIM partials, union-find arrays) comes from the device's default
        // stream-ordered pool: keep what it has handed out instead of returning it to the driver at every synchronisation
        // (the default release threshold of 0 turns every cudaMallocAsync after a sync into a fresh physical allocation)
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    *out = ctx;
    return KE_OK;
}

void ctx_destroy_one(ke_ctx* ctx) {
    KeDeviceGuard guard(ctx->device);
    cudaDeviceSynchronize();
    ke_tables_free(ctx->tables);
    ke_resize_tables_free(ctx->resize_tables);
    ke_resize_mma_tables_free(ctx->resize_mma);
    for (auto p : ctx->d_scratch) cudaFree(p);
    for (auto p : ctx->h_pinned) cudaFreeHost(p);
    for (auto s : ctx->copy_stream)
        if (s) cudaStreamDestroy(s);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    for (auto ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto ev : ctx->stage_ev)
        if (ev) cudaEventDestroy(ev);
    delete ctx;
}

}  // namespace

extern "C" int ke_ctx_create_multi(const int* devices, int n, ke_ctx** out) {
    KE_REQUIRE(out != nullptr, "ke_ctx_create_multi: out is NULL");
    *out = nullptr;
    KE_REQUIRE(devices != nullptr && n >= 1 && n <= KE_MAX_DEVICES, "ke_ctx_create_multi: need 1..%d devices (got %d)",
               KE_MAX_DEVICES, n);
    for (int a = 0; a < n; ++a)
        for (int b = a + 1; b < n; ++b) KE_REQUIRE(devices[a] != devices[b], "ke_ctx_create_multi: device %d listed twice", devices[a]);
    ke_ctx* root = nullptr;
    int rc = ctx_create_one(devices[0], &root);
    if (rc) return rc;
    for (int k = 1; k < n; ++k) {
        ke_ctx* child = nullptr;
        if ((rc = ctx_create_one(devices[k], &child))) {
            for (int q = 1; q < k; ++q) ctx_destroy_one(root->dev_ctx[q]);
            ctx_destroy_one(root);
            return rc;
        }
        root->dev_ctx[k] = child;
    }
    root->n_dev = n;
    *out = root;
    return KE_OK;
}

extern "C" int ke_ctx_create(int device, ke_ctx** out) { return ke_ctx_create_multi(&device, 1, out); }

extern "C" void ke_ctx_destroy(ke_ctx* ctx) {
    if (!ctx) return;
    for (int k = 1; k < ctx->n_dev; ++k)
        if (ctx->dev_ctx[k]) ctx_destroy_one(ctx->dev_ctx[k]);
    ctx_destroy_one(ctx);
}

extern "C" int ke_ctx_set_option(ke_ctx* ctx, int option, int value) {
    KE_REQUIRE(ctx != nullptr, "ke_ctx_set_option: ctx is NULL");
    switch (option) {
        case KE_OPT_PHASH_GENERIC:
        case KE_OPT_PHASH_CFG:
        case KE_OPT_SSIM_V1:
        case KE_OPT_RESIZE_GENERIC: break;
        case KE_OPT_JOIN_MODE: KE_REQUIRE(value >= 0 && value <= 3, "ke_ctx_set_option: join mode must be 0..3"); break;
        default: ke_set_error("ke_ctx_set_option: unknown option %d", option); return KE_E_INVALID;
    }
    for (int k = 0; k < ctx->n_dev; ++k) {
        ke_ctx* c = ctx->dev_ctx[k];
        switch (option) {
            case KE_OPT_PHASH_GENERIC: c->force_generic_phash = value ? 1 : 0; break;
            case KE_OPT_PHASH_CFG: c->phash_cfg = value; break;
            case KE_OPT_SSIM_V1: c->force_ssim_v1 = value ? 1 : 0; break;
            case KE_OPT_RESIZE_GENERIC: c->force_generic_resize = value ? 1 : 0; break;
            default: c->join_mode = value;
        }
    }
    return KE_OK;
}

extern "C" int ke_ctx_device(const ke_ctx* ctx) { return ctx ? ctx->device : -1; }
extern "C" int ke_ctx_device_count(const ke_ctx* ctx) { return ctx ? ctx->n_dev : -1; }
extern "C" ke_ctx* ke_ctx_child(ke_ctx* ctx, int k) { return (ctx && k >= 0 && k < ctx->n_dev) ? ctx->dev_ctx[k] : nullptr; }
extern "C" int ke_ctx_sm_count(const ke_ctx* ctx) { return ctx ? ctx->sm_count : -1; }
extern "C" int64_t ke_ctx_launch_count(const ke_ctx* ctx) {
    if (!ctx) return -1;
    int64_t total = 0;
    for (int k = 0; k < ctx->n_dev; ++k) total += ctx->dev_ctx[k]->launches;
    return total;
}

int ke_ctx_scratch(ke_ctx* ctx, int slot, size_t bytes, void** out) {
    if (ctx->d_scratch_bytes[slot] < bytes) {
        if (ctx->d_scratch[slot]) KE_CUDA(cudaFree(ctx->d_scratch[slot]));
        ctx->d_scratch[slot] = nullptr;
        ctx->d_scratch_bytes[slot] = 0;
        size_t want = bytes + bytes / 4 + 256;
        KE_CUDA(cudaMalloc(&ctx->d_scratch[slot], want));
        ctx->d_scratch_bytes[slot] = want;
    }
    *out = ctx->d_scratch[slot];
    return KE_OK;
}

int ke_ctx_pinned(ke_ctx* ctx, int slot, size_t bytes, void** out) {
    if (ctx->h_pinned_bytes[slot] < bytes) {
        if (ctx->h_pinned[slot]) KE_CUDA(cudaFreeHost(ctx->h_pinned[slot]));
        ctx->h_pinned[slot] = nullptr;
        ctx->h_pinned_bytes[slot] = 0;
        KE_CUDA(cudaMallocHost(&ctx->h_pinned[slot], bytes));
        ctx->h_pinned_bytes[slot] = bytes;
    }
    *out = ctx->h_pinned[slot];
    return KE_OK;
}

// ------------------------------------------------------------------------------------------
// Pillow 8bpc LANCZOS tables.  Double precision with libm sin on the HOST: the device's sin may
// differ by an ulp and flip a quantised tap, and tables depend on (in,out) only.

namespace {
constexpr int kPrecisionBits = 32 - 8 - 2;
constexpr double kLanczosSupport = 3.0;

inline double sinc_pi(double x) {
    if (x == 0.0) return 1.0;
    const double px = x * M_PI;
    return std::sin(px) / px;
}
inline double lanczos3(double x) { return (x >= -3.0 && x < 3.0) ? sinc_pi(x) * sinc_pi(x / 3.0) : 0.0; }

struct Geometry {
    double scale, filterscale, support;
    int ksize;
};
inline Geometry geometry(int in_size, int out_size) {
    Geometry g;
    // Pillow keeps the box edges as C floats: (double)(in1 - in0) / outSize
    g.scale = (double)((float)in_size - 0.0f) / (double)out_size;
    g.filterscale = g.scale < 1.0 ? 1.0 : g.scale;
    g.support = kLanczosSupport * g.filterscale;
    g.ksize = (int)std::ceil(g.support) * 2 + 1;
    return g;
}
}  // namespace

extern "C" int ke_resample_ksize(int in_size, int out_size) {
    if (in_size <= 0 || out_size <= 0) return KE_E_INVALID;
    return geometry(in_size, out_size).ksize;
}

extern "C" int ke_resample_table(int in_size, int out_size, int32_t* kk, int32_t* bounds, int ksize) {
    KE_REQUIRE(in_size > 0 && out_size > 0 && kk && bounds, "ke_resample_table: bad arguments");
    const Geometry g = geometry(in_size, out_size);
    KE_REQUIRE(ksize == g.ksize, "ke_resample_table: ksize %d != %d", ksize, g.ksize);
    std::vector<double> w((size_t)g.ksize);
    const double inv_fs = 1.0 / g.filterscale;
    for (int o = 0; o < out_size; ++o) {
        const double center = (o + 0.5) * g.scale;
        int first = (int)(center - g.support + 0.5);
        if (first < 0) first = 0;
        int last = (int)(center + g.support + 0.5);
        if (last > in_size) last = in_size;
        const int count = last - first;
        double total = 0.0;
        for (int t = 0; t < count; ++t) {
            w[t] = lanczos3((t + first - center + 0.5) * inv_fs);
            total += w[t];
        }
        int32_t* row = kk + (size_t)o * g.ksize;
        for (int t = 0; t < g.ksize; ++t) {
            double v = 0.0;
            if (t < count) v = (total != 0.0 ? w[t] / total : w[t]) * (double)(1 << kPrecisionBits);
            row[t] = v < 0 ? (int32_t)(v - 0.5) : (int32_t)(v + 0.5);
        }
        bounds[2 * o] = first;
        bounds[2 * o + 1] = count;
    }
    return KE_OK;
}


// ------------------------------------------------------------------------------------------
// Host union-find over accepted pairs: representative = smallest id of the component.

extern "C" int ke_cluster_pairs_host(const int64_t* h_a, const int64_t* h_b, int64_t n_pairs, int64_t* h_nodes,
                                     int64_t* h_node_rep, int64_t* n_nodes) {
    KE_REQUIRE(n_pairs >= 0 && n_nodes != nullptr, "ke_cluster_pairs_host: bad arguments");
    *n_nodes = 0;
    if (n_pairs == 0) return KE_OK;
    KE_REQUIRE(h_a && h_b && h_nodes && h_node_rep, "ke_cluster_pairs_host: NULL buffer");
    int64_t lo = h_a[0], hi = h_a[0];
    for (int64_t k = 0; k < n_pairs; ++k) {
        lo = std::min(lo, std::min(h_a[k], h_b[k]));
        hi = std::max(hi, std::max(h_a[k], h_b[k]));
    }
    // slot of an id: direct index when the id range is compact (table indices), else rank among the sorted ids
    const bool direct = (hi - lo) >= 0 && (hi - lo) < 8 * n_pairs + 4096;  // else O(range) passes would dominate
    std::vector<int64_t> ids, hkey;
    std::vector<int32_t> parent, hval;
    size_t hmask = 0;
    auto hash_of = [](int64_t v) -> size_t {
        uint64_t x = (uint64_t)v * 0x9E3779B97F4A7C15ull;
        return (size_t)(x ^ (x >> 29));
    };
    if (direct) {
        parent.assign((size_t)(hi - lo + 1), -1);
    } else {
        ids.reserve((size_t)n_pairs * 2);
        ids.insert(ids.end(), h_a, h_a + n_pairs);
        ids.insert(ids.end(), h_b, h_b + n_pairs);
        std::sort(ids.begin(), ids.end());
        ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
        parent.resize(ids.size());
        for (size_t i = 0; i < parent.size(); ++i) parent[i] = (int32_t)i;
        // id -> rank among the sorted ids through an open-addressing table (a binary search per lookup was the hot spot)
        size_t cap = 16;
        while (cap < 4 * ids.size()) cap <<= 1;
        hkey.assign(cap, 0);
        hval.assign(cap, -1);
        hmask = cap - 1;
        for (size_t i = 0; i < ids.size(); ++i) {
            size_t at = hash_of(ids[i]) & hmask;
            while (hval[at] >= 0) at = (at + 1) & hmask;
            hkey[at] = ids[i];
            hval[at] = (int32_t)i;
        }
    }
    auto slot = [&](int64_t v) -> int32_t {
        if (direct) {
            const int32_t s = (int32_t)(v - lo);
            if (parent[(size_t)s] < 0) parent[(size_t)s] = s;  // first sight of this id
            return s;
        }
        size_t at = hash_of(v) & hmask;
        while (hkey[at] != v || hval[at] < 0) at = (at + 1) & hmask;  // every id of the pairs is in the table
        return hval[at];
    };
    auto find = [&](int32_t x) {
        int32_t r = x;
        while (parent[(size_t)r] != r) r = parent[(size_t)r];
        while (parent[(size_t)x] != r) {
            const int32_t nx = parent[(size_t)x];
            parent[(size_t)x] = r;
            x = nx;
        }
        return r;
    };
    for (int64_t k = 0; k < n_pairs; ++k) {
        const int32_t rx = find(slot(h_a[k])), ry = find(slot(h_b[k]));
        if (rx != ry) {  // slots are ordered like the ids: the smaller one stays the root
            if (rx < ry) parent[(size_t)ry] = rx;
            else parent[(size_t)rx] = ry;
        }
    }
    // output grouped by component: components by ascending representative, members ascending inside each
    std::vector<int32_t> place(parent.size() + 1, 0);
    int64_t n = 0;
    for (size_t sidx = 0; sidx < parent.size(); ++sidx) {
        if (parent[sidx] < 0) continue;
        ++place[(size_t)find((int32_t)sidx) + 1];  // size of the component, counted at its root
        ++n;
    }
    for (size_t i = 1; i < place.size(); ++i) place[i] += place[i - 1];  // first output position of root i-1 at place[i-1]
    for (size_t sidx = 0; sidx < parent.size(); ++sidx) {
        if (parent[sidx] < 0) continue;
        const int32_t r = parent[sidx];  // fully compressed by the find above
        const int64_t at = place[(size_t)r]++;
        h_nodes[at] = direct ? lo + (int64_t)sidx : ids[sidx];
        h_node_rep[at] = direct ? lo + (int64_t)r : ids[(size_t)r];
    }
    *n_nodes = n;
    return KE_OK;
}
