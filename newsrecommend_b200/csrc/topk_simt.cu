// fp32 CUDA-core distance + selection over a Unit list (the NRB_PATH_SIMT path).
//
// A classic register-tiled SGEMM (128 query rows x 64 item rows per step, K chunks of 16) whose
// epilogue never writes scores: every score above the row's running threshold is appended to
// the row's candidate buffer, which a warp prunes back to the best k when it fills up. Used for
// small batches (faiss itself switches away from BLAS below 20 queries, distances.cpp), as the
// GPU-side cross-check of the tcgen05 path, and for matrices the TMA path cannot address.
#include "common.cuh"
#include "internal.h"

namespace nrb {

namespace {
constexpr int SBM = 128, SBN = 64, SBK = 16, STHREADS = 256;

struct SimtSmem {
    float As[SBK][SBM + 4];
    float Bs[SBK][SBN + 4];
    float an[SBM];
    float bn[SBN];
    float thr[SBM];
    int cnt[SBM];
};

template <bool L2>
__global__ void __launch_bounds__(STHREADS)
topk_simt_kernel(const float* __restrict__ A, const float* __restrict__ An,
                 const float* __restrict__ B, const float* __restrict__ Bn, int kp, int64_t a_total,
                 int64_t b_total, const Unit* __restrict__ units, const int* __restrict__ n_units_p,
                 int k, float* __restrict__ part_key, int* __restrict__ part_idx,
                 float* __restrict__ cand_key_buf, int* __restrict__ cand_idx_buf) {
    __shared__ SimtSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;  // 16 x 16 threads -> 8 rows x 4 cols each
    const int n_units = *n_units_p;
    float* ck = cand_key_buf + (int64_t)blockIdx.x * SBM * CAND_CAP;
    int* ci = cand_idx_buf + (int64_t)blockIdx.x * SBM * CAND_CAP;

    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        Unit un = units[u];
        if (un.a_rows == 0) un.b_rows = 0;  // phantom unit (pair padding): only writes empty rows
        for (int r = tid; r < SBM; r += STHREADS) {
            sm.cnt[r] = 0;
            sm.thr[r] = (r < un.a_rows) ? NEG_INF : __builtin_huge_valf();
            const int64_t ar = (int64_t)un.a_row0 + r;
            sm.an[r] = (L2 && r < un.a_rows && ar < a_total) ? An[ar] : 0.f;
        }
        __syncthreads();
        for (int n0 = 0; n0 < un.b_rows; n0 += SBN) {
            float acc[8][4];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
            if (L2 && tid < SBN) {
                const int64_t br = (int64_t)un.b_row0 + n0 + tid;
                sm.bn[tid] = (n0 + tid < un.b_rows && br < b_total) ? Bn[br] : 0.f;
            }
            for (int k0 = 0; k0 < kp; k0 += SBK) {
                // A chunk: 128 rows x 16 floats = 512 float4, two per thread
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int f = tid + h * STHREADS;
                    const int r = f >> 2, c4 = f & 3;
                    const int64_t ar = (int64_t)un.a_row0 + r;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < un.a_rows && ar < a_total)
                        v = __ldg(reinterpret_cast<const float4*>(A + ar * kp + k0) + c4);
                    sm.As[c4 * 4 + 0][r] = v.x;
                    sm.As[c4 * 4 + 1][r] = v.y;
                    sm.As[c4 * 4 + 2][r] = v.z;
                    sm.As[c4 * 4 + 3][r] = v.w;
                }
                {
                    const int r = tid >> 2, c4 = tid & 3;
                    const int64_t br = (int64_t)un.b_row0 + n0 + r;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (n0 + r < un.b_rows && br < b_total)
                        v = __ldg(reinterpret_cast<const float4*>(B + br * kp + k0) + c4);
                    sm.Bs[c4 * 4 + 0][r] = v.x;
                    sm.Bs[c4 * 4 + 1][r] = v.y;
                    sm.Bs[c4 * 4 + 2][r] = v.z;
                    sm.Bs[c4 * 4 + 3][r] = v.w;
                }
                __syncthreads();
#pragma unroll
                for (int kk = 0; kk < SBK; kk++) {
                    const float4 a0 = *reinterpret_cast<const float4*>(&sm.As[kk][ty * 8]);
                    const float4 a1 = *reinterpret_cast<const float4*>(&sm.As[kk][ty * 8 + 4]);
                    const float4 b0 = *reinterpret_cast<const float4*>(&sm.Bs[kk][tx * 4]);
                    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                    const float b[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
                    for (int i = 0; i < 8; i++)
#pragma unroll
                        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
                }
                __syncthreads();
            }
            // selection epilogue: append candidates above the row threshold
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int r = ty * 8 + i;
                const float th = sm.thr[r];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int c = n0 + tx * 4 + j;
                    float key = acc[i][j];
                    if (L2) {
                        float dis = sm.an[r] + sm.bn[tx * 4 + j] - 2.f * key;
                        key = -fmaxf(dis, 0.f);
                    }
                    if (c < un.b_rows && key > th) {
                        const int p = atomicAdd(&sm.cnt[r], 1);
                        ck[r * CAND_CAP + p] = key;
                        ci[r * CAND_CAP + p] = un.b_row0 + c;
                    }
                }
            }
            __syncthreads();
            // prune rows that could overflow on the next step (<= SBN appends per step)
            for (int r = warp; r < SBM; r += STHREADS / 32) {
                const int n = sm.cnt[r];
                if (n > CAND_CAP - SBN) {
                    float t = warp_prune_row(ck + r * CAND_CAP, ci + r * CAND_CAP, n, k,
                                             ck + r * CAND_CAP, ci + r * CAND_CAP, lane);
                    if (lane == 0) {
                        sm.cnt[r] = n < k ? n : k;
                        sm.thr[r] = t;
                    }
                }
            }
            __syncthreads();
        }
        // final prune of every row straight into the unit's partial rows
        for (int r = warp; r < SBM; r += STHREADS / 32) {
            const int n = sm.cnt[r];
            const int64_t o = ((int64_t)u * UNIT_ROWS + r) * k;
            warp_prune_row(ck + r * CAND_CAP, ci + r * CAND_CAP, n, k, part_key + o, part_idx + o, lane);
        }
        __syncthreads();
    }
}
}  // namespace

int simt_grid(int n_units) {
    int g = sm_count() * 2;
    if (n_units > 0 && n_units < g) g = n_units;
    return g < 1 ? 1 : g;
}

size_t simt_scratch_bytes(int grid) {
    return (size_t)grid * SBM * CAND_CAP * (sizeof(float) + sizeof(int)) + 256;
}

int launch_topk_simt_dev(const nrb_matrix* a, const nrb_matrix* b, const Unit* units,
                         const int* n_units_dev, int grid, int metric, int k, float* part_key,
                         int* part_idx, void* scratch, size_t scratch_bytes, cudaStream_t st) {
    NRB_REQUIRE(a->raw && b->raw, "simt: raw planes required");
    NRB_REQUIRE(a->kp == b->kp && a->kp % SBK == 0, "simt: kp mismatch");
    NRB_REQUIRE(k >= 1 && k <= NRB_MAX_K, "simt: k=%d out of range [1,%d]", k, NRB_MAX_K);
    NRB_REQUIRE(metric != NRB_METRIC_L2 || (a->norms && b->norms), "simt: norms required for L2");
    if (scratch_bytes < simt_scratch_bytes(grid)) {
        set_error("simt: scratch too small");
        return NRB_ERR_WORKSPACE;
    }
    float* ck = (float*)scratch;
    int* ci = (int*)((char*)scratch + (size_t)grid * SBM * CAND_CAP * sizeof(float));
    if (metric == NRB_METRIC_L2)
        topk_simt_kernel<true><<<grid, STHREADS, 0, st>>>(a->raw, a->norms, b->raw, b->norms, a->kp,
                                                         a->n, b->n, units, n_units_dev, k,
                                                         part_key, part_idx, ck, ci);
    else
        topk_simt_kernel<false><<<grid, STHREADS, 0, st>>>(a->raw, a->norms, b->raw, b->norms, a->kp,
                                                          a->n, b->n, units, n_units_dev, k,
                                                          part_key, part_idx, ck, ci);
    NRB_LAUNCH_CHECK();
    return NRB_OK;
}

}  // namespace nrb
