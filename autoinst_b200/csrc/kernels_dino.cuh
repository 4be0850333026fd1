// kernels_dino.cuh — "next" row N3, DINOv2 half: per-view feature look-up of the major points and the mean over views.
//   image_based_features_per_patch, pipeline/utils/image/image_utils.py:264-346 (per view, behind the visibility
//   bookkeeping), point_to_pixel, pipeline/utils/image/point_to_pixels.py:21-30, dinov2_mean, image_utils.py:363-371.
// The reference queries an Open3D KD-tree per major point and view in a Python loop, fills an N x views x 384 float64
// array (0.7 GB at N = 8 k, 29 views) and averages it per point; here one kernel per view finds the feature-map pixel of
// every major point and one kernel forms the means straight from the float32 feature maps.
#pragma once
#include "common.cuh"

namespace ancuts {

// One thread per major point (camera frame): nearest visible chunk point by brute force over shared-memory tiles, kept if
// the distance is strictly below max_dist (:271-276), then K p, division by the depth, np.round (half to even), image
// bounds and depth > 0 (point_to_pixels.py:21-30), feature-map pixel = int(factor * pixel) (:255-256, 341-342).
// out[i] = row * map_w + col of the feature map, or -1.
__global__ void __launch_bounds__(256)
k_dino_view_pixels(int n, const double* __restrict__ major, int m, const double* __restrict__ vis, double max_dist,
                   double k00, double k01, double k02, double k10, double k11, double k12, double k20, double k21, double k22,
                   int img_h, int img_w, int map_h, int map_w, int* __restrict__ out) {
    __shared__ double sx[1024], sy[1024], sz[1024];
    const int i = blockIdx.x * 256 + threadIdx.x;
    double px = 0.0, py = 0.0, pz = 0.0;
    if (i < n) { px = major[(size_t)i * 3]; py = major[(size_t)i * 3 + 1]; pz = major[(size_t)i * 3 + 2]; }
    double best = 1e300;
    for (int t0 = 0; t0 < m; t0 += 1024) {
        const int tn = min(1024, m - t0);
        __syncthreads();
        for (int j = threadIdx.x; j < tn; j += 256) {
            sx[j] = vis[(size_t)(t0 + j) * 3];
            sy[j] = vis[(size_t)(t0 + j) * 3 + 1];
            sz[j] = vis[(size_t)(t0 + j) * 3 + 2];
        }
        __syncthreads();
        if (i < n) {
#pragma unroll 4
            for (int j = 0; j < tn; ++j) {
                const double dx = px - sx[j], dy = py - sy[j], dz = pz - sz[j];
                const double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                best = fmin(best, d);
            }
        }
    }
    if (i >= n) return;
    int res = -1;
    if (m > 0 && __dsqrt_rn(best) < max_dist) {
        // K @ p the way a BLAS kernel accumulates it (k = 0, 1, 2 with fused multiply-adds); only a value within one
        // rounding of x.5 could land on another pixel than with any other summation order
        double u = fma(k02, pz, fma(k01, py, k00 * px));
        double v = fma(k12, pz, fma(k11, py, k10 * px));
        const double w = fma(k22, pz, fma(k21, py, k20 * px));
        u = rint(u / w);
        v = rint(v / w);
        if (u < (double)img_w && u >= 0.0 && v < (double)img_h && v >= 0.0 && w > 0.0) {
            const double f0 = (double)map_h / (double)img_h, f1 = (double)map_w / (double)img_w;
            const int p0 = (int)(f0 * (double)(long long)v), p1 = (int)(f1 * (double)(long long)u);
            res = p0 * map_w + p1;
        }
    }
    out[i] = res;
}

// One warp per major point: views in order, a view counts if its looked-up feature vector has any non-zero entry (:366);
// float64 sum in view order, divided by the number of such views (np.mean along the view axis, :369-370); zero row if none.
template <int FPL>      // features per lane = ceil(F / 32)
__global__ void __launch_bounds__(256)
k_dino_mean(int n, int num_views, const int* __restrict__ view_pixel, const float* const* __restrict__ maps, int fdim,
            double* __restrict__ out, int* __restrict__ out_count) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    double acc[FPL];
#pragma unroll
    for (int q = 0; q < FPL; ++q) acc[q] = 0.0;
    int cnt = 0;
    for (int v = 0; v < num_views; ++v) {
        const int pix = view_pixel[(size_t)v * n + i];
        if (pix < 0) continue;                                     // warp-uniform
        const float* f = maps[v] + (size_t)pix * fdim;
        float x[FPL];
        bool any = false;
#pragma unroll
        for (int q = 0; q < FPL; ++q) {
            const int k = lane + 32 * q;
            x[q] = (k < fdim) ? f[k] : 0.0f;
            any |= (x[q] != 0.0f);
        }
        if (!__any_sync(0xffffffffu, any)) continue;
#pragma unroll
        for (int q = 0; q < FPL; ++q) acc[q] += (double)x[q];
        ++cnt;
    }
#pragma unroll
    for (int q = 0; q < FPL; ++q) {
        const int k = lane + 32 * q;
        if (k < fdim) out[(size_t)i * fdim + k] = cnt ? acc[q] / (double)cnt : 0.0;
    }
    if (lane == 0 && out_count) out_count[i] = cnt;
}

}  // namespace ancuts
