// kernels_pool.cuh — "next" row N3 (TARL half): radius-mean pooling of per-scan-point features onto the
// major (0.35 m) voxel points, the step of ncuts_chunk right before the affinity build.
//
// Reference: tarl_features_per_patch, pipeline/utils/point_cloud/chunk_generation.py:205-258 (called at
// pipeline/ncuts/ncuts_utils.py:135-141): the scan points of the ±10 neighbouring scans are cropped to the
// chunk cube with STRICT comparisons (:233-236), concatenated, put into an Open3D KD-tree, and every major
// point takes `np.mean(features[idx], axis=0)` over `search_radius_vector_3d(point, MAJOR_VOXEL_SIZE / 2)`
// (:249-252) in a Python loop; points without a neighbour keep a zero row, which the affinity build then
// neutralises (ncuts_utils.py:143-146).  Open3D 0.17 is not vendored: its radius search is nanoflann's
// RadiusResultSet, which keeps a point when the squared distance is strictly below radius^2.
//
// Here: scan points inside the cube get a cell key on a grid of pitch >= radius, one stable radix sort
// groups them by cell (CUB), and one warp per major point walks the 3 x 3 runs of three x-adjacent cells
// (contiguous key ranges found by binary search), tests the float64 squared distance, and accumulates the
// feature rows of the hits in float64 in a fixed order (run, key, original index): results do not depend
// on scheduling.  HBM-bound: the feature rows of the hit points are read once (coalesced 128-byte rows).
#pragma once
#include "common.cuh"

namespace ancuts {

struct PoolGrid {
    double lo[3], hi[3];      // open box: a scan point takes part iff lo < p < hi on every axis
    double inv_h;             // 1 / cell pitch (pitch >= radius)
    int n[3];                 // cells per axis
};

__device__ __forceinline__ int pool_cell(const PoolGrid& g, double x, int a) {
    int c = (int)floor((x - g.lo[a]) * g.inv_h);
    return max(0, min(g.n[a] - 1, c));
}

// key per scan point (UINT_MAX outside the cube, so the sort moves those behind every cell), identity payload
__global__ void __launch_bounds__(256)
k_pool_keys(int m, const double* __restrict__ pts, PoolGrid g, unsigned* __restrict__ keys, int* __restrict__ idx,
            int* __restrict__ inside_count) {
    int j = blockIdx.x * 256 + threadIdx.x;
    bool in = false;
    if (j < m) {
        const double x = pts[(size_t)j * 3], y = pts[(size_t)j * 3 + 1], z = pts[(size_t)j * 3 + 2];
        in = x > g.lo[0] && x < g.hi[0] && y > g.lo[1] && y < g.hi[1] && z > g.lo[2] && z < g.hi[2];
        unsigned key = 0xFFFFFFFFu;
        if (in) key = ((unsigned)pool_cell(g, z, 2) * (unsigned)g.n[1] + (unsigned)pool_cell(g, y, 1)) * (unsigned)g.n[0]
                      + (unsigned)pool_cell(g, x, 0);
        keys[j] = key;
        idx[j] = j;
    }
    unsigned b = __ballot_sync(0xffffffffu, in);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(inside_count, __popc(b));
}

__device__ __forceinline__ int pool_lower_bound(const unsigned* __restrict__ keys, int n, unsigned v) {
    int lo = 0, hi = n;                     // first position with keys[pos] >= v
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// one warp per major point; FPL = feature values per lane (feat_dim <= 32 * FPL)
template <int FPL>
__global__ void __launch_bounds__(256)
k_pool_gather(int nmajor, const double* __restrict__ major, int m, const double* __restrict__ pts,
              const float* __restrict__ feat, int fdim, PoolGrid g, double r2, int normalise,
              const unsigned* __restrict__ keys, const int* __restrict__ idx, const int* __restrict__ inside_count,
              double* __restrict__ out, int* __restrict__ out_count) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= nmajor) return;
    const int nin = min(*inside_count, m);
    const double qx = major[(size_t)i * 3], qy = major[(size_t)i * 3 + 1], qz = major[(size_t)i * 3 + 2];
    // runs: lane r < 9 owns the three x-adjacent cells at (cy + r % 3 - 1, cz + r / 3 - 1)
    int run_lo = 0, run_hi = 0;
    {
        // unclamped cell of the query (it may lie outside the cube; hits can then only sit in border cells)
        const int cx = (int)floor((qx - g.lo[0]) * g.inv_h);
        const int cy = (int)floor((qy - g.lo[1]) * g.inv_h) + (lane % 3) - 1;
        const int cz = (int)floor((qz - g.lo[2]) * g.inv_h) + (lane / 3) - 1;
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.n[0] - 1);
        if (lane < 9 && cy >= 0 && cy < g.n[1] && cz >= 0 && cz < g.n[2] && x0 <= x1) {
            const unsigned base = ((unsigned)cz * (unsigned)g.n[1] + (unsigned)cy) * (unsigned)g.n[0];
            run_lo = pool_lower_bound(keys, nin, base + (unsigned)x0);
            run_hi = pool_lower_bound(keys, nin, base + (unsigned)x1 + 1u);
        }
    }
    double acc[FPL];
#pragma unroll
    for (int t = 0; t < FPL; ++t) acc[t] = 0.0;
    int cnt = 0;
    for (int r = 0; r < 9; ++r) {
        const int lo = __shfl_sync(0xffffffffu, run_lo, r), hi = __shfl_sync(0xffffffffu, run_hi, r);
        for (int b0 = lo; b0 < hi; b0 += 32) {
            const int pos = b0 + lane;
            int j = -1;
            bool hit = false;
            if (pos < hi) {
                j = __ldg(idx + pos);
                const double dx = pts[(size_t)j * 3] - qx, dy = pts[(size_t)j * 3 + 1] - qy, dz = pts[(size_t)j * 3 + 2] - qz;
                hit = (dx * dx + dy * dy + dz * dz) < r2;          // nanoflann RadiusResultSet: strictly inside
            }
            unsigned mask = __ballot_sync(0xffffffffu, hit);
            cnt += __popc(mask);
            while (mask) {
                const int src = __ffs(mask) - 1;
                mask &= mask - 1;
                const int jj = __shfl_sync(0xffffffffu, j, src);
                const float* row = feat + (size_t)jj * fdim;
#pragma unroll
                for (int t = 0; t < FPL; ++t) {
                    const int c = lane + 32 * t;
                    if (c < fdim) acc[t] += (double)__ldg(row + c);
                }
            }
        }
    }
    if (cnt > 0) {
        double nn = 0.0;
#pragma unroll
        for (int t = 0; t < FPL; ++t) { acc[t] = acc[t] / (double)cnt; nn += acc[t] * acc[t]; }      // np.mean: sum / count
        if (normalise) {                                       // TARL_NORM (chunk_generation.py:253-254; False in config.py:64)
            nn = sqrt(warp_sum(nn));
#pragma unroll
            for (int t = 0; t < FPL; ++t) acc[t] = acc[t] / nn;
        }
    }
#pragma unroll
    for (int t = 0; t < FPL; ++t) {
        const int c = lane + 32 * t;
        if (c < fdim) out[(size_t)i * fdim + c] = acc[t];
    }
    if (out_count && lane == 0) out_count[i] = cnt;
}

}  // namespace ancuts
