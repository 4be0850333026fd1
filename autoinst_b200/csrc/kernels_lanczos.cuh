// kernels_lanczos.cuh — stage 3: batched Lanczos for the Fiedler vector of every active node.
//   Replaces eigsh(A, 2, sigma=1e-10, which='LM') + argsort (normalized_cut.py:49-53).
// Operator: M = S (w + I) S with S = D^-1/2; the Fiedler vector of L = I - M is the eigenvector
// of the second largest eigenvalue of M.  The largest (1, vector D^1/2 1) is deflated by keeping
// it as row 0 of the basis V and re-orthogonalising against it like any Lanczos vector.
// W is read as float32; every vector and every sum is float64.
#pragma once
#include "common.cuh"
#include "kernels_graph.cuh"

namespace ancuts {

__device__ __forceinline__ double start_value(int i) {
    // same integer hash as oracle/device_model.py::start_vector
    unsigned long long x = ((unsigned long long)i + 1ull) * 0x9E3779B97F4A7C15ull;
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    return (double)(x >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

// Start vector of a node.  With coordinates at hand (segment calls) it is the SMOOTH vector D^1/2 (q - q_0), q = the
// points projected on the fixed direction (1, 0.7, 0.4), plus 1 % of the hash: the Fiedler vector of a proximity graph
// varies slowly along the object, so this start has a far larger component along it than a random vector
// (numpy model, 33 root nodes: 1745 -> 1613 steps; the principal axis of the node gives 1600).  The deflation
// against D^1/2 1 removes the choice of origin.  Without coordinates (stage entry, caller-provided W): the hash.
struct StartVec {
    const double* base;      // coordinates of the node's chunk, or NULL
    const int* perm;         // global position -> chunk-local input index
    const double* deg;
    double o0, o1, o2;
    __device__ __forceinline__ StartVec(const Eng& e, int chunk, int start) {
        base = e.pts ? e.pts + (size_t)e.c_base[chunk] * 3 : nullptr;
        perm = e.perm;
        deg = e.deg;
        o0 = o1 = o2 = 0.0;
        if (base) { const double* q = base + (size_t)perm[start] * 3; o0 = q[0]; o1 = q[1]; o2 = q[2]; }
    }
    __device__ __forceinline__ double at(int pos, int i) const {      // pos = global position, i = index inside the node
        const double h = start_value(i);
        if (!base) return h;
        const double* q = base + (size_t)perm[pos] * 3;
        return sqrt(deg[pos]) * ((q[0] - o0) + 0.7 * (q[1] - o1) + 0.4 * (q[2] - o2)) + 0.01 * h;
    }
};

// One CTA per active node: u1 = sqrt(d)/||sqrt(d)|| -> V row 0; w0 = start - u1 (u1.start) -> wbuf,
// ||w0|| -> a_bprev.  Also resets the per-node Lanczos state.
__global__ void __launch_bounds__(256)
k_lanczos_init(Eng e) {
    __shared__ double red[8];
    int a = blockIdx.x;
    if (e.a_done[a] != DONE_NO) return;          // already finished by the cluster kernel
    int r = e.a_rid[a];
    int start = e.r_start[r], n = e.r_n[r];
    const StartVec sv0(e, e.r_chunk[r], start);
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += e.deg[start + i];
    double vol = block_sum_256(s, red);
    double inv = 1.0 / sqrt(vol);
    double dot = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        double u = sqrt(e.deg[start + i]) * inv;
        e.V[start + i] = u;                                   // row 0
        dot += u * sv0.at(start + i, i);
    }
    dot = block_sum_256(dot, red);
    double nn = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        double x = sv0.at(start + i, i) - dot * e.V[start + i];
        e.wbuf[start + i] = x;
        e.zbuf[start + i] = e.sinv[start + i] * x;
        nn += x * x;
    }
    nn = block_sum_256(nn, red);
    if (threadIdx.x == 0) {
        e.a_bprev[a] = sqrt(nn);
        e.a_k[a] = 0;
        int kcap = min(e.kmax, n - 1);
        e.a_kcap[a] = kcap;
        e.a_done[a] = DONE_NO;
        e.a_conv[a] = 0;
    }
}

// y = S (w+I) S v  with v = wbuf / bprev (the normalisation of the previous step is folded in
// here; every row also stores its entry of v into V row k+1).  zbuf = S wbuf is written by the
// producer of wbuf, so  y_i = s_i/bprev * (sum_j w_ij zbuf_j + zbuf_i).
// grid: (row blocks of 8*R rows, active).  One warp owns R rows and walks the 16-byte window of the
// block's columns 256 columns at a time: all 2R 128-bit loads of W are issued before the first use.
template <int R>
__global__ void __launch_bounds__(256)
k_matvec(Eng e, int cur) {
    int a = blockIdx.y;
    if (e.a_done[a] != DONE_NO) return;
    NodeView v = node_view(e, e.a_rid[a], cur);
    int row0 = blockIdx.x * (8 * R);
    if (row0 >= v.n) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int myrow0 = row0 + warp * R;
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(&e.acct[SG_MATVEC], 4ull * v.n * v.n + 8ull * v.n);
    if (myrow0 >= v.n) return;
    const double invb = 1.0 / e.a_bprev[a];
    const int k = e.a_k[a];
    const double* __restrict__ z = e.zbuf + (v.start - v.ro);      // indexed by chunk-local column
    const float* rp[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr)
        rp[rr] = v.W + (size_t)(v.ro + min(myrow0 + rr, v.n - 1)) * v.ld;
    double acc[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) acc[rr] = 0.0;
    const int c_lo = v.ro, c_hi = v.ro + v.n;
    const int a0 = c_lo & ~3;
    for (int c = a0 + lane * 4; c < c_hi; c += 256) {
        const int cb = c + 128;
        const bool hasb = cb < c_hi;
        float4 wa[R], wb[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) wa[rr] = ld_stream4(rp[rr] + c);
#pragma unroll
        for (int rr = 0; rr < R; ++rr) wb[rr] = hasb ? ld_stream4(rp[rr] + cb) : make_float4(0.f, 0.f, 0.f, 0.f);
        double za[4], zb[4];
        bool va[4], vb[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            va[q] = (c + q >= c_lo) & (c + q < c_hi);
            vb[q] = hasb & (cb + q < c_hi);
            za[q] = va[q] ? __ldg(z + c + q) : 0.0;
            zb[q] = vb[q] ? __ldg(z + cb + q) : 0.0;
        }
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            // entries outside the block belong to other nodes or are uninitialised: select, do not multiply
            double s0 = (va[0] ? (double)wa[rr].x : 0.0) * za[0] + (va[1] ? (double)wa[rr].y : 0.0) * za[1];
            double s1 = (va[2] ? (double)wa[rr].z : 0.0) * za[2] + (va[3] ? (double)wa[rr].w : 0.0) * za[3];
            double s2 = (vb[0] ? (double)wb[rr].x : 0.0) * zb[0] + (vb[1] ? (double)wb[rr].y : 0.0) * zb[1];
            double s3 = (vb[2] ? (double)wb[rr].z : 0.0) * zb[2] + (vb[3] ? (double)wb[rr].w : 0.0) * zb[3];
            acc[rr] += (s0 + s1) + (s2 + s3);
        }
    }
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
        double s = warp_sum(acc[rr]);
        int row = myrow0 + rr;
        if (lane == 0 && row < v.n) {
            int g = v.start + row;
            double si = e.sinv[g];
            e.ybuf[g] = si * invb * (s + e.zbuf[g]);              // (w + I) z
            e.V[(size_t)(k + 1) * e.P + g] = e.wbuf[g] * invb;
        }
    }
}

// partial dots of ybuf with basis rows 0..k+1 over one CH-column chunk.
// grid: (chunks, row tiles of 32 basis rows, active)
__global__ void __launch_bounds__(256)
k_dots(Eng e) {
    int a = blockIdx.z;
    if (e.a_done[a] != DONE_NO) return;
    int nch = e.a_nch[a];
    int ch = blockIdx.x;
    if (ch >= nch) return;
    int rows = e.a_k[a] + 2;
    int j0 = blockIdx.y * 32;
    if (j0 >= rows) return;
    int r = e.a_rid[a];
    int start = e.r_start[r], n = e.r_n[r];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int cbase = ch * CH;
    double yr[CH / 32];
#pragma unroll
    for (int m = 0; m < CH / 32; ++m) {
        int c = cbase + lane + 32 * m;
        yr[m] = c < n ? e.ybuf[start + c] : 0.0;
    }
    double* out = e.p_dot + (size_t)(e.a_slot0[a] + ch) * e.KS;
    for (int j = j0 + warp; j < min(j0 + 32, rows); j += 8) {
        const double* vr = e.V + (size_t)j * e.P + start;
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < CH / 32; ++m) {
            int c = cbase + lane + 32 * m;
            if (c < n) s += vr[c] * yr[m];
        }
        s = warp_sum(s);
        if (lane == 0) out[j] = s;
    }
    if (blockIdx.y == 0 && warp == 0) {          // |y|^2 over the chunk, stored behind the last basis row
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < CH / 32; ++m) s += yr[m] * yr[m];
        s = warp_sum(s);
        if (lane == 0) out[rows] = s;
    }
}

// y -= V^T h over one chunk (classical Gram-Schmidt against u1 and every Lanczos vector).
// Pass 1 decides, from |y|^2 and |h|^2, whether a second pass is needed (DGKS test as in ARPACK:
// |y - V h| < 0.717 |y|); if not, it finishes the step (wbuf, zbuf, partial squared norm); if so,
// it leaves the partial dots of the updated y for pass 2, which exits at once for nodes that
// do not need it.   grid: (chunks, active)
template <int PASS>
__global__ void __launch_bounds__(256)
k_update(Eng e) {
    extern __shared__ double sm[];          // hs[KS] | ych[CH] | red[8]
    double* hs = sm;
    double* ych = sm + e.KS;
    double* red = ych + CH;
    int a = blockIdx.y;
    if (e.a_done[a] != DONE_NO) return;
    if (PASS == 2 && !e.a_need2[a]) return;
    int nch = e.a_nch[a];
    int ch = blockIdx.x;
    if (ch >= nch) return;
    int rows = e.a_k[a] + 2;
    int r = e.a_rid[a];
    int start = e.r_start[r], n = e.r_n[r];
    const double* pd = (PASS == 1 ? e.p_dot : e.p_dot2) + (size_t)e.a_slot0[a] * e.KS;
    double hh = 0.0;
    for (int j = threadIdx.x; j < rows; j += 256) {
        double h = 0.0;
        for (int c = 0; c < nch; ++c) h += pd[(size_t)c * e.KS + j];     // fixed order
        hs[j] = h;
        hh += h * h;
    }
    bool second = false;
    if (PASS == 1) {
        hh = block_sum_256(hh, red);          // includes a __syncthreads: hs[] is complete
        double yy = 0.0;
        for (int c = 0; c < nch; ++c) yy += pd[(size_t)c * e.KS + rows];
        second = (yy - hh) < 0.5 * yy;        // DGKS test on the whole projection (eta^2 = 1/2): in Lanczos the
                                              // three-term part alone usually trips it, i.e. two passes almost always
    }
    __syncthreads();
    if (ch == 0 && threadIdx.x == 0) {
        if (PASS == 1) { e.a_h1[a] = hs[rows - 1]; e.a_h2[a] = 0.0; e.a_need2[a] = second ? 1 : 0; }
        else e.a_h2[a] = hs[rows - 1];
    }
    int c0 = ch * CH + threadIdx.x, c1 = c0 + 256;
    double y0 = c0 < n ? e.ybuf[start + c0] : 0.0;
    double y1 = c1 < n ? e.ybuf[start + c1] : 0.0;
    const double* vb = e.V + start;
    const int cc0 = min(c0, n - 1), cc1 = min(c1, n - 1);                 // clamped: loads are always legal
    int j = 0;
    for (; j + 8 <= rows; j += 8) {
        double t0[8], t1[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double* vr = vb + (size_t)(j + u) * e.P;
            t0[u] = vr[cc0];
            t1[u] = vr[cc1];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { y0 -= hs[j + u] * t0[u]; y1 -= hs[j + u] * t1[u]; }
    }
    for (; j < rows; ++j) {
        const double* vr = vb + (size_t)j * e.P;
        y0 -= hs[j] * vr[cc0];
        y1 -= hs[j] * vr[cc1];
    }
    if (c0 >= n) y0 = 0.0;
    if (c1 >= n) y1 = 0.0;
    if (PASS == 1 && second) {
        if (c0 < n) e.ybuf[start + c0] = y0;
        if (c1 < n) e.ybuf[start + c1] = y1;
        ych[threadIdx.x] = y0;
        ych[threadIdx.x + 256] = y1;
        __syncthreads();
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        double yr[CH / 32];
#pragma unroll
        for (int m = 0; m < CH / 32; ++m) yr[m] = ych[lane + 32 * m];
        // other CTAs of this node may still be reading the pass-1 partials: pass 2 has its own buffer
        double* out = e.p_dot2 + (size_t)(e.a_slot0[a] + ch) * e.KS;
        for (int jj = warp; jj < rows; jj += 8) {
            const double* vr = vb + (size_t)jj * e.P;
            double s = 0.0;
#pragma unroll
            for (int m = 0; m < CH / 32; ++m) {
                int c = ch * CH + lane + 32 * m;
                if (c < n) s += vr[c] * yr[m];
            }
            s = warp_sum(s);
            if (lane == 0) out[jj] = s;
        }
    } else {
        if (c0 < n) { e.wbuf[start + c0] = y0; e.zbuf[start + c0] = e.sinv[start + c0] * y0; }
        if (c1 < n) { e.wbuf[start + c1] = y1; e.zbuf[start + c1] = e.sinv[start + c1] * y1; }
        double nn = block_sum_256(y0 * y0 + y1 * y1, red);
        if (threadIdx.x == 0) e.p_norm[e.a_slot0[a] + ch] = nn;
    }
}

// alpha_k, beta_k, step counter.  One thread per active node.
__global__ void k_lanczos_finalize(Eng e, int num_active) {
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= num_active) return;
    if (e.a_done[a] != DONE_NO) return;
    int k = e.a_k[a];
    double nn = 0.0;
    int s0 = e.a_slot0[a], nch = e.a_nch[a];
    for (int c = 0; c < nch; ++c) nn += e.p_norm[s0 + c];
    double beta = sqrt(nn);
    e.a_alpha[(size_t)a * e.KS + k] = e.a_h1[a] + e.a_h2[a];
    e.a_beta[(size_t)a * e.KS + k] = beta;
    e.a_bprev[a] = beta;
    k += 1;
    e.a_k[a] = k;
    // breakdown (invariant subspace) or Krylov space exhausted: wait for the check kernel
    if (beta < 1e-13 || k >= e.a_kcap[a]) e.a_done[a] = DONE_HOLD;
}

// ---------------------------------------------------------------------------------------------
// Convergence check: two largest eigenvalues of the k x k tridiagonal by Sturm-count
// multisection (128 shifts per round), eigenvector of the largest by inverse iteration with a
// pivoted tridiagonal LU, residual = beta_{k-1} |y_{k-1}|.
// One CTA of 128 threads per active node.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int sturm_count(const double* al, const double* be2, int k, double x, double pivmin) {
    int cnt = 0;
    double q = al[0] - x;
    if (fabs(q) < pivmin) q = -pivmin;
    cnt += (q < 0.0);
    for (int i = 1; i < k; ++i) {
        q = al[i] - x - be2[i - 1] / q;
        if (fabs(q) < pivmin) q = -pivmin;
        cnt += (q < 0.0);
    }
    return cnt;     // number of eigenvalues < x
}

// whole CTA (128 threads): the two largest eigenvalues of the k x k tridiagonal at once, 64 shifts per
// eigenvalue and round (65^10 > 2^60: full float64 resolution of the Gershgorin bracket in 10 rounds).
__device__ void tridiag_top2_bisect(const double* al, const double* be2, int k, double glo, double ghi,
                                    double pivmin, int* cnts, double* bounds, double* th1, double* th2) {
    const int tid = threadIdx.x;
    const int which = tid >> 6, t64 = tid & 63;      // 0: largest, 1: second largest
    const int m = k - 1 - which;
    double lo = glo, hi = ghi;
    for (int round = 0; round < 10; ++round) {
        double x = lo + (hi - lo) * ((double)(t64 + 1) / 65.0);
        cnts[tid] = (m >= 0) ? sturm_count(al, be2, k, x, pivmin) : 0;
        __syncthreads();
        if (t64 == 0) {
            // eigenvalue m lies in (x_t, x_{t+1}] with count(x_t) <= m < count(x_{t+1})
            int t = -1;
            for (int i = 0; i < 64; ++i) if (cnts[which * 64 + i] <= m) t = i;
            bounds[2 * which] = (t >= 0) ? lo + (hi - lo) * ((double)(t + 1) / 65.0) : lo;
            bounds[2 * which + 1] = (t < 63) ? lo + (hi - lo) * ((double)(t + 2) / 65.0) : hi;
        }
        __syncthreads();
        lo = bounds[2 * which];
        hi = bounds[2 * which + 1];
        __syncthreads();
    }
    *th1 = 0.5 * (bounds[0] + bounds[1]);
    *th2 = (k > 1) ? 0.5 * (bounds[2] + bounds[3]) : -1e300;
}

__global__ void __launch_bounds__(128)
k_lanczos_check(Eng e, int force) {
    extern __shared__ double sm[];
    const int KS = e.KS;
    double* al = sm;             // alpha
    double* be = al + KS;        // beta (off-diagonals)
    double* be2 = be + KS;       // beta^2
    double* dd = be2 + KS;       // LU diagonal
    double* du = dd + KS;        // LU first super-diagonal
    double* du2 = du + KS;       // LU second super-diagonal
    double* dl = du2 + KS;       // LU multipliers
    double* yv = dl + KS;        // inverse-iteration vector
    int* swp = reinterpret_cast<int*>(yv + KS);   // row interchange flags
    __shared__ int cnts[128];
    __shared__ double bounds[4];
    __shared__ double gb[3];

    int a = blockIdx.x;
    int done = e.a_done[a];
    if (done == DONE_YES) return;
    int k = e.a_k[a];
    bool due = (done == DONE_HOLD) || force || (k > 0 && (k % e.check_every) == 0);
    if (!due || k == 0) {
        if (threadIdx.x == 0 && done == DONE_NO) atomicAdd(&e.ctr[4], 1);
        return;
    }
    const int tid = threadIdx.x;
    for (int i = tid; i < k; i += 128) {
        al[i] = e.a_alpha[(size_t)a * KS + i];
        double b = e.a_beta[(size_t)a * KS + i];
        be[i] = b;
        be2[i] = b * b;
    }
    __syncthreads();
    if (tid == 0) {                     // Gershgorin interval and pivmin
        double lo = 1e300, hi = -1e300, bmax = 0.0;
        for (int i = 0; i < k; ++i) {
            double rad = (i > 0 ? fabs(be[i - 1]) : 0.0) + (i < k - 1 ? fabs(be[i]) : 0.0);
            lo = fmin(lo, al[i] - rad);
            hi = fmax(hi, al[i] + rad);
            if (i < k - 1) bmax = fmax(bmax, be2[i]);
        }
        double w = fmax(fmax(fabs(lo), fabs(hi)), 1e-300);
        gb[0] = lo - 1e-10 * w;
        gb[1] = hi + 1e-10 * w;
        gb[2] = fmax(bmax, 1.0) * 1.0020841800044864e-292;      // safmin / eps
    }
    __syncthreads();
    const double glo = gb[0], ghi = gb[1], pivmin = gb[2];
    double th1, th2;
    tridiag_top2_bisect(al, be2, k, glo, ghi, pivmin, cnts, bounds, &th1, &th2);
    if (tid == 0) {
        // eigenvector of th1 from the factorisation of T - th1 I twisted at the first index (bottom-up
        // pivots, all safely negative below the largest eigenvalue; see kernels_cluster.cuh)
        double d = al[k - 1] - th1;
        if (fabs(d) < pivmin) d = -pivmin;
        dd[k - 1] = d;
        for (int i = k - 2; i >= 0; --i) {
            d = al[i] - th1 - be2[i] / d;
            if (fabs(d) < pivmin) d = -pivmin;
            dd[i] = d;
        }
        double z = 1.0, ss = 1.0;
        yv[0] = 1.0;
        for (int i = 0; i < k - 1; ++i) {
            z = -be[i] * z / dd[i + 1];
            if (fabs(z) > 1e150) {
                for (int j = 0; j <= i; ++j) yv[j] *= 1e-150;
                z *= 1e-150;
                ss *= 1e-300;
            }
            yv[i + 1] = z;
            ss += z * z;
        }
        double inv = 1.0 / sqrt(ss);
        for (int i = 0; i < k; ++i) yv[i] *= inv;
        double res = fabs(be[k - 1] * yv[k - 1]);
        double gap = fmax(th1 - th2, 1e-300);
        int n = e.r_n[e.a_rid[a]];
        bool exhausted = (k >= n - 1);
        bool breakdown = be[k - 1] < 1e-13;
        bool conv = exhausted || breakdown || (res <= e.tol * gap);
        bool stop = conv || (k >= e.a_kcap[a]) || force;
        e.a_theta[2 * a] = th1;
        e.a_theta[2 * a + 1] = th2;
        if (stop) {
            for (int i = 0; i < k; ++i) e.a_y[(size_t)a * KS + i] = yv[i];
            e.a_conv[a] = conv ? 1 : 0;
            e.a_done[a] = DONE_YES;
        } else {
            atomicAdd(&e.ctr[4], 1);
        }
    }
}

// ev = sum_j y_j V[j+1] over one chunk, plus per-chunk (sum, min, max).  grid: (chunks, active)
__global__ void __launch_bounds__(256)
k_ritz(Eng e) {
    extern __shared__ double sm[];      // y[KS] | red[3*8]
    double* ys = sm;
    double* red = sm + e.KS;
    int a = e.sel ? e.sel[blockIdx.y] : blockIdx.y;
    if (e.a_path[a] == 0) return;                // the cluster kernel already wrote ev and its statistics
    int nch = e.a_nch[a];
    int ch = blockIdx.x;
    if (ch >= nch) return;
    int k = e.a_k[a];
    int r = e.a_rid[a];
    int start = e.r_start[r], n = e.r_n[r];
    for (int j = threadIdx.x; j < k; j += 256) ys[j] = e.a_y[(size_t)a * e.KS + j];
    __syncthreads();
    int c0 = ch * CH + threadIdx.x, c1 = c0 + 256;
    double x0 = 0.0, x1 = 0.0;
    const double* vb = e.V + start + (size_t)e.P;     // row 1
#pragma unroll 4
    for (int j = 0; j < k; ++j) {
        const double* vr = vb + (size_t)j * e.P;
        double y = ys[j];
        if (c0 < n) x0 += y * vr[c0];
        if (c1 < n) x1 += y * vr[c1];
    }
    if (c0 < n) e.ev[start + c0] = x0;
    if (c1 < n) e.ev[start + c1] = x1;
    double s = (c0 < n ? x0 : 0.0) + (c1 < n ? x1 : 0.0);
    double mn = fmin(c0 < n ? x0 : 1e300, c1 < n ? x1 : 1e300);
    double mx = fmax(c0 < n ? x0 : -1e300, c1 < n ? x1 : -1e300);
    double q = (c0 < n ? x0 * x0 : 0.0) + (c1 < n ? x1 * x1 : 0.0);
    s = warp_sum(s); mn = warp_min(mn); mx = warp_max(mx); q = warp_sum(q);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { red[w] = s; red[8 + w] = mn; red[16 + w] = mx; red[24 + w] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0.0, tmn = 1e300, tmx = -1e300, tq = 0.0;
        for (int i = 0; i < 8; ++i) { ts += red[i]; tmn = fmin(tmn, red[8 + i]); tmx = fmax(tmx, red[16 + i]); tq += red[24 + i]; }
        double* o = e.p_stat + (size_t)(e.a_slot0[a] + ch) * 4;   // sum, min, max, sum of squares
        o[0] = ts; o[1] = tmn; o[2] = tmx; o[3] = tq;
    }
}

}  // namespace ancuts
