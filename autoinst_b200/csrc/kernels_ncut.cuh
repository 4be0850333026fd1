// kernels_ncut.cuh — stage 4a: the ten threshold cuts of every node from ONE pass over its block.
//   Replaces get_min_ncut / ncut_cost / cut_cost (normalized_cut.py:4-34), which densify the
//   degree matrix twice per threshold.
// Every point gets bucket b_i = #{k : ev_i > t_k}; the reference's mask for threshold k is
// {b_i > k}.  An edge (i,j) with b_j < b_i is cut exactly for k in [b_j, b_i - 1], so a difference
// array over the buckets gives all ten cut weights.  Cut weights are accumulated in 2^-40 fixed
// point with integer atomics: the sums are independent of scheduling, and two thresholds that
// give the same mask get bit-identical costs (the reference keeps the first, :30).
#pragma once
#include "common.cuh"
#include "kernels_graph.cuh"

namespace ancuts {

// Fixed-point exponent of a node's cut sums from its volume (sum of degrees >= twice the total weight): 2^-40 steps
// unless the running sums could leave 62 bits (caller-provided w with weights far above 1, or a huge
// PROXIMITY_THRESHOLD); then fewer fraction bits.  Every CTA of k_scan and k_decide derive the same value.
__device__ __forceinline__ int fix_shift(double vol) {
    const int e = ilogb(fmax(vol, 1.0));                 // vol < 2^(e+1)
    return min(FIX_SHIFT, 61 - e);
}

// one thread per active node: sign (sum(ev) >= 0, the oracle's canonical sign), min/max,
// np.allclose(mn, mx) (normalized_cut.py:22-23), thresholds exactly as np.linspace(endpoint=False)
// computes them (k*step + mn with separate roundings), and accumulator reset.
__global__ void k_ev_final(Eng e, int num_active) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= num_active) return;
    if (e.sel) a = e.sel[a];
    if (e.a_fused[a]) return;                  // decided inside the cluster kernel (cl_fused_cut)
    int s0 = e.a_slot0[a], nch = e.a_nch[a];
    double sum = 0.0, mn = 1e300, mx = -1e300, q = 0.0;
    for (int c = 0; c < nch; ++c) {
        const double* o = e.p_stat + (size_t)(s0 + c) * 4;
        sum += o[0];
        mn = fmin(mn, o[1]);
        mx = fmax(mx, o[2]);
        q += o[3];
    }
    double scale = 1.0 / sqrt(q);            // unit norm, as eigsh returns
    double sg = (sum < 0.0) ? -scale : scale;
    if (sum < 0.0) { double t = mn; mn = -mx; mx = -t; }
    mn *= scale;
    mx *= scale;
    e.a_sign[a] = sg;
    // np.allclose(mn, mx): |mn - mx| <= atol + rtol*|mx|, rtol 1e-5, atol 1e-8
    e.a_nocut[a] = (fabs(mn - mx) <= 1e-8 + 1e-5 * fabs(mx)) ? 1 : 0;
    double step = (mx - mn) / (double)NCUT;
    for (int k = 0; k < NCUT; ++k) e.a_thr[a * NCUT + k] = __dadd_rn(__dmul_rn((double)k, step), mn);
    for (int b = 0; b <= NB; ++b) e.a_diff[a * (NB + 1) + b] = 0ull;
    for (int b = 0; b < NB; ++b) e.a_cnt[a * NB + b] = 0;
}

// bucket per point, signed unit-norm ev written back, per-chunk bucket volumes (sum of degrees,
// deterministic block reduction) and bucket counts.  grid: (chunks, active)
__global__ void __launch_bounds__(256)
k_bucket(Eng e) {
    __shared__ double red[8];
    __shared__ int scnt[NB];
    __shared__ double thr[NCUT];
    int a = e.sel ? e.sel[blockIdx.y] : blockIdx.y;
    if (e.a_fused[a]) return;
    int nch = e.a_nch[a];
    int ch = blockIdx.x;
    if (ch >= nch) return;
    int r = e.a_rid[a];
    int start = e.r_start[r], n = e.r_n[r];
    if (threadIdx.x < NB) scnt[threadIdx.x] = 0;
    if (threadIdx.x < NCUT) thr[threadIdx.x] = e.a_thr[a * NCUT + threadIdx.x];
    __syncthreads();
    double sg = e.a_sign[a];
    int cc[2] = {ch * CH + (int)threadIdx.x, ch * CH + (int)threadIdx.x + 256};
    int bk[2] = {-1, -1};
    double dg[2] = {0.0, 0.0};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        if (cc[u] < n) {
            int g = start + cc[u];
            double x = e.ev[g] * sg;
            e.ev[g] = x;
            int b = 0;
#pragma unroll
            for (int k = 0; k < NCUT; ++k) b += (x > thr[k]) ? 1 : 0;
            e.bucket[g] = (uint8_t)b;
            bk[u] = b;
            dg[u] = e.deg[g];
            atomicAdd(&scnt[b], 1);
        }
    }
    double* pv = e.p_vol + (size_t)(e.a_slot0[a] + ch) * NB;
    for (int b = 0; b < NB; ++b) {
        double v = (bk[0] == b ? dg[0] : 0.0) + (bk[1] == b ? dg[1] : 0.0);
        v = block_sum_256(v, red);
        if (threadIdx.x == 0) pv[b] = v;
    }
    __syncthreads();
    if (threadIdx.x < NB && scnt[threadIdx.x]) atomicAdd(&e.a_cnt[a * NB + threadIdx.x], scnt[threadIdx.x]);
}

// one pass over the upper triangle of the block: difference array of cut weights.
// grid: (row blocks of 8 rows, active)
__global__ void __launch_bounds__(256)
k_scan(Eng e, int cur) {
    __shared__ unsigned long long sdiff[NB + 1];
    __shared__ double svol[NB];
    __shared__ double sscale;
    int a = e.sel ? e.sel[blockIdx.y] : blockIdx.y;
    if (e.a_fused[a] || e.a_nocut[a]) return;
    NodeView v = node_view(e, e.a_rid[a], cur);
    int row0 = blockIdx.x * 8;
    if (row0 >= v.n) return;
    if (threadIdx.x <= NB) sdiff[threadIdx.x] = 0ull;
    if (threadIdx.x < NB) {                        // volume per bucket, chunk slots in order (as k_decide)
        const int s0 = e.a_slot0[a], nch = e.a_nch[a];
        double t = 0.0;
        for (int ch = 0; ch < nch; ++ch) t += e.p_vol[(size_t)(s0 + ch) * NB + threadIdx.x];
        svol[threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int b = 0; b < NB; ++b) t += svol[b];
        sscale = ldexp(1.0, fix_shift(t));
    }
    __syncthreads();
    const double fscale = sscale;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int row = row0 + warp;
    if (row < v.n) {
        int bi = e.bucket[v.start + row];
        const float* rowp = v.W + (size_t)(v.ro + row) * v.ld;
        const uint8_t* bkt = e.bucket + v.start - v.ro;       // indexed by chunk-local column
        int c_lo = v.ro + row + 1, c_hi = v.ro + v.n;
        int a0 = c_lo & ~3;
        for (int c = a0 + lane * 4; c < c_hi; c += 128) {
            float4 w = ld_stream4(rowp + c);
            float in[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int cc = c + k;
                if (cc >= c_lo && cc < c_hi && in[k] != 0.0f) {
                    int bj = bkt[cc];
                    if (bj != bi) {
                        int lo = min(bi, bj), hi = max(bi, bj);
                        long long q = __double2ll_rn((double)in[k] * fscale);
                        atomicAdd(&sdiff[lo], (unsigned long long)q);
                        atomicAdd(&sdiff[hi], (unsigned long long)(-q));
                    }
                }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x <= NB && sdiff[threadIdx.x] != 0ull)
        atomicAdd(&e.a_diff[a * (NB + 1) + threadIdx.x], sdiff[threadIdx.x]);
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(&e.acct[SG_SCAN], 2ull * v.n * v.n);
}

// one thread per active node: N-cut value of the ten cuts (normalized_cut.py:7-11), first strictly
// smallest (:27-32), decision mcut < T (:56), side sizes and their stop rule (:39-40), log record.
__global__ void k_decide(Eng e, int num_active) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= num_active) return;
    if (e.sel) a = e.sel[a];
    if (e.a_fused[a]) return;
    int r = e.a_rid[a];
    int n = e.r_n[r], c = e.r_chunk[r];
    int best = -1;
    double bestc = INFINITY;
    int cnt[NB];
    for (int b = 0; b < NB; ++b) cnt[b] = e.a_cnt[a * NB + b];
    if (!e.a_nocut[a]) {
        double vol[NB];
        int s0 = e.a_slot0[a], nch = e.a_nch[a];
        for (int b = 0; b < NB; ++b) vol[b] = 0.0;
        for (int ch = 0; ch < nch; ++ch)
            for (int b = 0; b < NB; ++b) vol[b] += e.p_vol[(size_t)(s0 + ch) * NB + b];
        double vtot = 0.0;
        for (int b = 0; b < NB; ++b) vtot += vol[b];
        const double unfix = ldexp(1.0, -fix_shift(vtot));
        long long run = 0;
        for (int k = 0; k < NCUT; ++k) {
            run += (long long)e.a_diff[a * (NB + 1) + k];
            double cut = (double)run * unfix;
            double assoc_b = 0.0, assoc_a = 0.0;
            for (int b = 0; b <= k; ++b) assoc_b += vol[b];          // mask false: ev <= t_k
            for (int b = k + 1; b < NB; ++b) assoc_a += vol[b];       // mask true
            double cost = (cut / assoc_a) + (cut / assoc_b);
            e.a_costs[a * NCUT + k] = cost;
            if (cost < bestc) { bestc = cost; best = k; }
        }
    } else {
        for (int k = 0; k < NCUT; ++k) e.a_costs[a * NCUT + k] = INFINITY;
    }
    e.a_bestk[a] = best;
    e.a_mcut[a] = bestc;
    bool split = (best >= 0) && (bestc < e.T);
    int n_a = 0;
    if (best >= 0) for (int b = best + 1; b < NB; ++b) n_a += cnt[b];
    if (split) {
        int n_b = n - n_a;
        double no = (double)e.c_norig[c] + 1e-8;
        e.r_pass[2 * r + 0] = (n_a > 2 && (double)n_a / no > CHILD_SPLIT_LIM) ? 1 : 0;   // side 0 = mask side
        e.r_pass[2 * r + 1] = (n_b > 2 && (double)n_b / no > CHILD_SPLIT_LIM) ? 1 : 0;
        e.r_status[r] = ST_SPLIT;
        int s = atomicAdd(&e.ctr[5], 1);
        e.split_ids[s] = r;
        atomicMax(&e.ctr[7], n);
    } else {
        e.r_status[r] = ST_LEAF;
    }
    if (!e.a_conv[a]) atomicAdd(&e.ctr[16], 1);      // stopped at lanczos_max_steps: reported by every segment call
    if (e.stats != nullptr) {
        int s = atomicAdd(&e.ctr[6], 1);
        if (s < e.stats_cap) {
            ancuts_node_stat st;
            st.chunk = c; st.n = n; st.steps = e.a_k[a]; st.converged = e.a_conv[a];
            st.best_k = best; st.split = split ? 1 : 0; st.level = e.r_level[r]; st.n_side = n_a;
            st.lambda2 = 1.0 - e.a_theta[2 * a]; st.mcut = bestc;
            e.stats[s] = st;
        }
    }
}

// side flag per point of the nodes that split: 0 = mask side (ev > t), 1 = the rest.
// grid: (chunks, active)
__global__ void __launch_bounds__(256)
k_sides(Eng e) {
    int a = e.sel ? e.sel[blockIdx.y] : blockIdx.y;
    int r = e.a_rid[a];
    if (e.a_fused[a] || e.r_status[r] != ST_SPLIT) return;
    int start = e.r_start[r], n = e.r_n[r];
    int bk = e.a_bestk[a];
    for (int c = blockIdx.x * CH + threadIdx.x; c < min(n, (int)(blockIdx.x + 1) * CH); c += 256) {
        int g = start + c;
        e.side[g] = (e.bucket[g] > bk) ? 0 : 1;
    }
}

}  // namespace ancuts
