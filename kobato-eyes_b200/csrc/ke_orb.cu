// ke_orb.cu — N2: the matching half of dup.refine._compute_orb_ratio (reference src/dup/refine.py:55-68),
// batched over candidate pairs, sm_100a.
//
//     matches = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(da, db)
//     ratio   = len(matches) / min(len(kpa), len(kpb))
//
// What OpenCV's cross-check keeps (checked against cv2.BFMatcher of opencv-python-headless 4.13 on 300 random
// descriptor sets with many ties; 4.12 is pinned by the reference): the MUTUAL nearest neighbours — query i matches train
// t(i) = argmin_j d(i, j) iff q(t(i)) = argmin_i' d(i', t(i)) is i again, the first index winning a tie on either side.
// One CTA per pair: both descriptor sets sit in shared memory (<= 500 x 32 B each by ORB's default nfeatures); a thread
// walks the other set for its descriptor (8 XOR + 8 POPC per candidate, the walked words are a shared-memory broadcast),
// first the queries, then the trains, then the mutual check.
// The FAST/Harris detector and the rBRIEF descriptor stay with OpenCV on host threads (not built: §8f ranks a
// bit-exact GPU ORB as "a larger, fuzzier parity problem").
#include <algorithm>

#include "ke_common.cuh"

namespace {

constexpr int kT = 256;
constexpr int kWords = 8;  // 256-bit descriptors

struct OrbArgs {
    const uint32_t* desc;    // [total][8]
    const long long* off_a;  // first descriptor row of the query set of pair p
    const int* cnt_a;
    const long long* off_b;  // ... of the train set
    const int* cnt_b;
    long long n_pairs;
    int max_a, max_b;        // largest query / train set (shared-memory sizing)
    int* n_matches;          // [n_pairs]
    int* match_train;        // nullable: [n_pairs][max_a] train index matched to query i, or -1
    int* match_dist;         // nullable: [n_pairs][max_a]
};

__device__ __forceinline__ int nearest(const uint32_t* mine, const uint32_t* others, int n_others, int& best_d) {
    uint32_t t[kWords];
#pragma unroll
    for (int w = 0; w < kWords; ++w) t[w] = mine[w];
    int best = -1;
    best_d = 1 << 30;
    for (int i = 0; i < n_others; ++i) {  // every lane reads the same words: shared-memory broadcast
        const uint4 q0 = *reinterpret_cast<const uint4*>(others + i * kWords);
        const uint4 q1 = *reinterpret_cast<const uint4*>(others + i * kWords + 4);
        const int d = __popc(q0.x ^ t[0]) + __popc(q0.y ^ t[1]) + __popc(q0.z ^ t[2]) + __popc(q0.w ^ t[3]) +
                      __popc(q1.x ^ t[4]) + __popc(q1.y ^ t[5]) + __popc(q1.z ^ t[6]) + __popc(q1.w ^ t[7]);
        if (d < best_d) best_d = d, best = i;  // strict: the first index wins a tie
    }
    return best;
}

__global__ void __launch_bounds__(kT) ke_orb_match_kernel(const OrbArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t* s_q = smem;                                   // [max_a][8] query descriptors
    uint32_t* s_t = s_q + (size_t)a.max_a * kWords;         // [max_b][8] train descriptors
    int* s_tq = reinterpret_cast<int*>(s_t + (size_t)a.max_b * kWords);  // [max_b] nearest query of train j
    __shared__ int s_count;
    for (long long p = blockIdx.x; p < a.n_pairs; p += gridDim.x) {
        const int na = a.cnt_a[p], nb = a.cnt_b[p];
        const uint32_t* qa = a.desc + a.off_a[p] * kWords;
        const uint32_t* tb = a.desc + a.off_b[p] * kWords;
        __syncthreads();  // previous pair's readers are done
        for (int i = threadIdx.x; i < na * kWords; i += kT) s_q[i] = qa[i];
        for (int i = threadIdx.x; i < nb * kWords; i += kT) s_t[i] = tb[i];
        if (threadIdx.x == 0) s_count = 0;
        __syncthreads();
        for (int j = threadIdx.x; j < nb; j += kT) {
            int d;
            s_tq[j] = nearest(s_t + j * kWords, s_q, na, d);
        }
        __syncthreads();
        int mine = 0;
        for (int i = threadIdx.x; i < na; i += kT) {
            int d;
            const int t = nearest(s_q + i * kWords, s_t, nb, d);
            const bool hit = t >= 0 && s_tq[t] == i;
            mine += hit;
            if (a.match_train) a.match_train[p * a.max_a + i] = hit ? t : -1;
            if (a.match_dist) a.match_dist[p * a.max_a + i] = hit ? d : -1;
        }
        for (int off = 16; off; off >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, off);
        if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_count, mine);
        __syncthreads();
        if (threadIdx.x == 0) a.n_matches[p] = s_count;
    }
}

}  // namespace

extern "C" int ke_orb_match_pairs(ke_ctx* ctx, const uint8_t* d_desc, const int64_t* d_off_a, const int32_t* d_cnt_a,
                                  const int64_t* d_off_b, const int32_t* d_cnt_b, int64_t n_pairs, int max_a, int max_b,
                                  int32_t* d_n_matches, int32_t* d_match_train, int32_t* d_match_dist, void* stream) {
    KE_REQUIRE(ctx != nullptr && n_pairs >= 0, "ke_orb_match_pairs: bad arguments");
    if (n_pairs == 0) return KE_OK;
    KE_REQUIRE(d_desc && d_off_a && d_cnt_a && d_off_b && d_cnt_b && d_n_matches, "ke_orb_match_pairs: NULL buffer");
    KE_REQUIRE(max_a >= 1 && max_b >= 1 && (long long)(max_a + max_b) * (kWords * 4) + max_b * 4 <= 200 * 1024,
               "ke_orb_match_pairs: %d + %d descriptors do not fit shared memory", max_a, max_b);
    KE_REQUIRE((reinterpret_cast<uintptr_t>(d_desc) & 15) == 0, "ke_orb_match_pairs: descriptors must be 16-byte aligned");
    KeDeviceGuard guard(ctx->device);
    OrbArgs a;
    a.desc = reinterpret_cast<const uint32_t*>(d_desc);
    a.off_a = (const long long*)d_off_a, a.cnt_a = d_cnt_a, a.off_b = (const long long*)d_off_b, a.cnt_b = d_cnt_b;
    a.n_pairs = n_pairs, a.max_a = max_a, a.max_b = max_b, a.n_matches = d_n_matches, a.match_train = d_match_train, a.match_dist = d_match_dist;
    const int smem = (max_a + max_b) * kWords * 4 + max_b * 4;
    KE_CUDA(cudaFuncSetAttribute(ke_orb_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const unsigned grid = (unsigned)std::min<long long>(n_pairs, (long long)ctx->sm_count * 8);
    ke_orb_match_kernel<<<grid, kT, smem, (cudaStream_t)stream>>>(a);
    ctx->launches++;
    KE_CUDA(cudaGetLastError());
    return KE_OK;
}
