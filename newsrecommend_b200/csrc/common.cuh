// Shared helpers for libnrb200: error plumbing, ordered keys, and the warp-cooperative
// candidate-buffer prune used by every selection epilogue.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/nrb200.h"

namespace nrb {

// ---------------------------------------------------------------- host-side error plumbing
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define NRB_CUDA_CHECK(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            nrb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                  \
                           cudaGetErrorString(_e));                                       \
            return NRB_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

#define NRB_LAUNCH_CHECK()                                                                \
    do {                                                                                  \
        nrb::count_launch();                                                              \
        NRB_CUDA_CHECK(cudaGetLastError());                                               \
    } while (0)

#define NRB_REQUIRE(cond, ...)                                                            \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            nrb::set_error(__VA_ARGS__);                                                  \
            return NRB_ERR_INVALID;                                                       \
        }                                                                                 \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------- selection primitives
// All selection works on "keys" where LARGER IS BETTER: key = score for inner product and
// key = -distance for L2. A candidate is the pair (key, idx); idx is a 32-bit row index
// (>= 0). Candidates are ordered by (key descending, idx ascending), so exact ties resolve to
// the lowest id the way faiss's k=1 strict-compare heap does.
constexpr int CAND_CAP = 256;  // per-row candidate buffer capacity (entries)
constexpr float NEG_INF = -__builtin_huge_valf();

__device__ __forceinline__ uint64_t pack_cand(float key, int idx) {
    uint32_t u = __float_as_uint(key);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((uint64_t)u << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)idx);
}
__device__ __forceinline__ float cand_key(uint64_t c) {
    uint32_t u = (uint32_t)(c >> 32);
    u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
    return __uint_as_float(u);
}
__device__ __forceinline__ int cand_idx(uint64_t c) {
    return (int)(0xFFFFFFFFu - (uint32_t)(c & 0xFFFFFFFFu));
}
// Monotone float -> uint map (0 is below every float, including -inf): shared running bounds are
// raised with atomicMax on this encoding.
__device__ __forceinline__ uint32_t ordered_u32(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_u32(uint32_t u) {
    if (u == 0) return NEG_INF;
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}
// The empty candidate: (-inf, idx -1) packs below every real candidate.
__device__ __forceinline__ uint64_t empty_cand() { return pack_cand(NEG_INF, -1); }

__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
    uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, m);
    uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), m);
    return ((uint64_t)hi << 32) | lo;
}

// Bitonic sort (descending) of 32*R packed candidates held striped across a warp:
// element e = i*32 + lane lives in c[i].
template <int R>
__device__ __forceinline__ void warp_bitonic_desc(uint64_t (&c)[R], int lane) {
#pragma unroll
    for (int k2 = 2; k2 <= 32 * R; k2 <<= 1) {
#pragma unroll
        for (int j = k2 >> 1; j >= 1; j >>= 1) {
            if (j >= 32) {
                const int jj = j >> 5;
#pragma unroll
                for (int i = 0; i < R; i++) {
                    if ((i & jj) == 0) {
                        const bool desc = (((i * 32) & k2) == 0);  // k2 >= 64 here
                        uint64_t a = c[i], b = c[i | jj];
                        uint64_t mx = a > b ? a : b, mn = a > b ? b : a;
                        c[i] = desc ? mx : mn;
                        c[i | jj] = desc ? mn : mx;
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < R; i++) {
                    const int e = i * 32 + lane;
                    const bool desc = ((e & k2) == 0);
                    const bool lower = ((lane & j) == 0);
                    uint64_t a = c[i];
                    uint64_t b = shfl_xor_u64(a, j);
                    const bool take_max = (lower == desc);
                    c[i] = take_max ? (a > b ? a : b) : (a > b ? b : a);
                }
            }
        }
    }
}

// Candidate storage views. The SIMT kernel keeps keys and ids in two arrays (SoA); the tcgen05
// kernels keep one array of 8-byte (key, idx) entries (AoS): an append is ONE predicated 64-bit store
// (half the store wavefronts of the first tiles of a cold unit, where 32 lanes write 32 different
// rows, and 6 instead of 8 instructions per expanded element), a prune loads one 64-bit word per entry.
struct SoAView {
    float* k;
    int* i;
    __device__ __forceinline__ void load(int e, float& key, int& idx) const {
        key = k[e];
        idx = i[e];
    }
    __device__ __forceinline__ void store(int e, float key, int idx) const {
        k[e] = key;
        i[e] = idx;
    }
    __device__ __forceinline__ const void* base() const { return k; }
};
struct AoSView {
    uint2* p;
    __device__ __forceinline__ void load(int e, float& key, int& idx) const {
        const uint2 v = p[e];
        key = __uint_as_float(v.x);
        idx = (int)v.y;
    }
    __device__ __forceinline__ void store(int e, float key, int idx) const {
        p[e] = make_uint2(__float_as_uint(key), (uint32_t)idx);
    }
    __device__ __forceinline__ const void* base() const { return p; }
};
// Output of a prune: back into the AoS buffer (p != null) or into a pair of partial-row arrays.
struct EitherView {
    uint2* p;
    float* k;
    int* i;
    __device__ __forceinline__ void store(int e, float key, int idx) const {
        if (p) {
            p[e] = make_uint2(__float_as_uint(key), (uint32_t)idx);
        } else {
            k[e] = key;
            i[e] = idx;
        }
    }
    __device__ __forceinline__ const void* base() const { return p ? (const void*)p : (const void*)k; }
};

// Warp-cooperative prune of one row's candidate buffer ((bk, bi) must have 32*R readable entries:
// the candidate buffers always have CAND_CAP). Sorts the first `n` entries of (bk, bi)
// best-first, finds kth = key of the k-th best (NEG_INF while fewer than k exist) and keeps the
// best k PLUS every entry whose key is within `margin` of kth (key > kth - margin), at most
// `keep_max`; writes `width` entries to (ok, oi) (kept ones, then empty candidates; ok/oi may
// alias bk/bi). Returns the new append threshold kth - margin; *kept = entries kept; *overflow
// is set when more than keep_max entries were within the margin (the row is then incomplete
// and must be recomputed by the exact path). margin = 0, keep_max = width = k is the plain
// top-k prune. All 32 lanes must call with identical arguments.
// R = registers per lane: the sort covers 32*R entries, so n must be <= 32*R (width may be larger:
// the rest of the output row is padding). When ok aliases bk, width must be <= 32*R.
template <int R, class In, class Out>
__device__ __forceinline__ float warp_prune_row_v(const In in, int n, int k, float margin, int keep_max, int width,
                                                  const Out out, int lane, int* kept, bool* overflow,
                                                  float floor_thr = NEG_INF, float* kth_out = nullptr) {
    // unconditional loads (32*R <= CAND_CAP entries are always allocated; the tail is masked): the
    // compiler batches them, whereas predicated loads were issued a few at a time
    float rk[R];
    int ri[R];
#pragma unroll
    for (int i = 0; i < R; i++) in.load(i * 32 + lane, rk[i], ri[i]);
    uint64_t c[R];
#pragma unroll
    for (int i = 0; i < R; i++) c[i] = (i * 32 + lane < n) ? pack_cand(rk[i], ri[i]) : empty_cand();
    __syncwarp();
    warp_bitonic_desc<R>(c, lane);
    float kth = NEG_INF;
    const int kl = (k - 1) & 31, ki = (k - 1) >> 5;
#pragma unroll
    for (int i = 0; i < R; i++)
        if (i == ki) kth = cand_key(c[i]);
    kth = __shfl_sync(0xffffffffu, kth, kl);
    if (kth_out) *kth_out = kth;
    // Keep (a prefix of the sorted entries): the best k plus everything within `margin` of the
    // k-th, but nothing below floor_thr -- a bound already established elsewhere (the query's
    // shared bound minus margin, or this row's previous threshold). The floor is NON-strict:
    // the row's own k-th entry has key == its threshold and must survive.
    const float thr = kth - margin;  // NEG_INF stays NEG_INF
    int c_thr = 0, c_floor = 0;
#pragma unroll
    for (int i = 0; i < R; i++) {
        const int e = i * 32 + lane;
        const float key = cand_key(c[i]);
        c_thr += (e < n && key > thr) ? 1 : 0;
        c_floor += (e < n && key >= floor_thr) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        c_thr += __shfl_xor_sync(0xffffffffu, c_thr, o);
        c_floor += __shfl_xor_sync(0xffffffffu, c_floor, o);
    }
    const int base = n < k ? n : k;
    int cnt = c_thr > base ? c_thr : base;
    cnt = cnt < c_floor ? cnt : c_floor;
    *overflow = cnt > keep_max;
    cnt = cnt > keep_max ? keep_max : cnt;
    *kept = cnt;
#pragma unroll
    for (int i = 0; i < R; i++) {
        const int e = i * 32 + lane;
        if (e < width) {
            const bool keep = e < cnt;
            out.store(e, keep ? cand_key(c[i]) : NEG_INF, keep ? cand_idx(c[i]) : -1);
        }
    }
    for (int e = 32 * R + lane; e < width; e += 32) out.store(e, NEG_INF, -1);  // output wider than the sort: padding
    __syncwarp();
    return fmaxf(thr, floor_thr);  // threshold for further appends
}

// The two-array form (SIMT kernel, tests).
template <int R = CAND_CAP / 32>
__device__ __forceinline__ float warp_prune_row_m(const float* bk, const int* bi, int n, int k, float margin,
                                                  int keep_max, int width, float* ok, int* oi, int lane,
                                                  int* kept, bool* overflow, float floor_thr = NEG_INF,
                                                  float* kth_out = nullptr) {
    return warp_prune_row_v<R>(SoAView{const_cast<float*>(bk), const_cast<int*>(bi)}, n, k, margin, keep_max, width,
                               SoAView{ok, oi}, lane, kept, overflow, floor_thr, kth_out);
}

// Cheap mid-unit prune: instead of sorting, find a LOWER BOUND lb of the row's k-th best key by
// bisection on the ordered-uint keys (count(key >= lb) is between k and k + slack, or exactly the
// k-th after 32 halvings), then keep every entry with key >= max(lb - margin, floor_thr) by
// stream compaction (order not preserved; ties are all kept, so no (key, idx) tie rule is
// needed here -- the exact sort at the end of the unit applies it). Dropped entries are below
// a valid bound of the k-th minus the margin, hence outside the final top-k / margin set.
// Returns the new append threshold; *kept = entries left in (bk, bi); *lb_u = ordered-uint lower
// bound of the k-th (0 while fewer than k entries exist). n <= 32*R. All lanes call together.
template <int R, class Buf>
__device__ __forceinline__ float warp_tighten_row(const Buf buf, int n, int k, float margin, float floor_thr,
                                                  int lane, int* kept, uint32_t* lb_u) {
    constexpr int SLACK = 3;
    float rk[R];
    int id[R];
#pragma unroll
    for (int i = 0; i < R; i++) buf.load(i * 32 + lane, rk[i], id[i]);  // unconditional, batched loads (see warp_prune_row_v)
    uint32_t u[R];
#pragma unroll
    for (int i = 0; i < R; i++) u[i] = (i * 32 + lane < n) ? ordered_u32(rk[i]) : 0u;  // 0 sorts below every float
    __syncwarp();
    uint32_t lo = 0;
    if (n >= k) {
        uint32_t mn = 0xffffffffu, mx = 0u;
#pragma unroll
        for (int i = 0; i < R; i++) {
            mn = (u[i] != 0u && u[i] < mn) ? u[i] : mn;
            mx = u[i] > mx ? u[i] : mx;
        }
        mn = __reduce_min_sync(0xffffffffu, mn);
        mx = __reduce_max_sync(0xffffffffu, mx);
        // invariant: count(u >= lo) >= k > count(u >= hi)
        lo = mn;
        uint32_t hi = mx + 1u;
        while (hi - lo > 1u) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            int c = 0;
#pragma unroll
            for (int i = 0; i < R; i++) c += (u[i] >= mid) ? 1 : 0;
            c = __reduce_add_sync(0xffffffffu, c);
            if (c >= k) {
                lo = mid;
                if (c <= k + SLACK) break;
            } else {
                hi = mid;
            }
        }
    }
    *lb_u = lo;
    const float thr = fmaxf(from_ordered_u32(lo) - margin, floor_thr);  // lo == 0 -> NEG_INF
    const uint32_t lt = (1u << lane) - 1u;
    int base = 0;
#pragma unroll
    for (int i = 0; i < R; i++) {
        const float key = from_ordered_u32(u[i]);
        const bool keep = (u[i] != 0u) && key >= thr;
        const uint32_t b = __ballot_sync(0xffffffffu, keep);
        if (keep) buf.store(base + __popc(b & lt), key, id[i]);
        base += __popc(b);
    }
    __syncwarp();
    *kept = base;
    return thr;
}

// Plain top-k prune (margin 0): keeps min(n, k) entries, returns kth.
__device__ __forceinline__ float warp_prune_row(const float* bk, const int* bi, int n, int k,
                                                float* ok, int* oi, int lane) {
    int kept;
    bool ovf;
    return warp_prune_row_m(bk, bi, n, k, 0.f, k, k, ok, oi, lane, &kept, &ovf);
}

}  // namespace nrb
