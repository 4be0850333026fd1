// kernels_cluster.cuh — stage 3, persistent variant: ONE thread-block cluster runs the whole Lanczos
// iteration of ONE node (init, matvec, re-orthogonalisation, convergence checks, Ritz vector) without
// returning to the host.  Replaces eigsh(A, 2, sigma=1e-10) + argsort (normalized_cut.py:49-53) for
// nodes of up to CL_NMAX points; larger nodes keep the grid-wide multi-launch path (kernels_lanczos.cuh).
//
// The C CTAs of a cluster own C row slices of the node's block.  Per step:
//   matvec of the slice (W streamed from L2/HBM, z = S v for the whole node in shared memory),
//   the three-term recurrence (alpha comes out of the matvec epilogue), then ONE classical Gram-Schmidt pass
//   against u1 = D^1/2 1 and every Lanczos vector: each CTA forms the partial dots over its slice, the
//   partials are exchanged through distributed shared memory (cluster barrier, then every CTA sums the C
//   partials in rank order, so all CTAs hold bit-identical alpha/beta and take identical decisions); the
//   squared norm is accumulated by the update pass.  Three cluster barriers and four block barriers per step.
// Every sum is float64 and evaluated in a fixed order: results do not depend on scheduling.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"
#include "kernels_graph.cuh"
#include "kernels_lanczos.cuh"

namespace ancuts {
namespace cg = cooperative_groups;

constexpr int CL_KMAX = 256;             // steps the cluster kernel can take; nodes needing more fall back
constexpr int CL_KS = CL_KMAX + 4;
constexpr int CL_RPMAX = 512;            // rows per CTA (two per thread)
constexpr int CL_NMAX = 4096;            // largest node handled here (8 CTAs x 512 rows)
constexpr int CL_THREADS = 512;          // 1024 threads (64 registers, 2 KB ring stages) measured 9 % slower
constexpr int CL_WARPS = CL_THREADS / 32;
constexpr int CL_HALF = CL_THREADS / 2;     // shifts per eigenvalue and bisection round
constexpr int CL_CLASSES = 6;            // node size bins (kernels_graph.cuh::cluster_class); cluster sizes 1, 2, 4, 8
constexpr int CL_DYN_SMEM = 195 * 1024;  // z for the whole node + as many basis rows of the slice as fit
                                         // (one CTA per SM; 256 threads x 2 CTAs per SM measured 25 % slower)

struct ClusterShared {
    double hpart[2][CL_KS];      // this CTA's partial dots (pass 1 / pass 2), read by the peers
    double npart[2];             // partial squared norm of the slice
    double spart[4];             // at the end: partial (sum, min, max, sum of squares) of the Ritz vector
    double hs[CL_KS];            // reduced projection coefficients
    double alpha[CL_KS], beta[CL_KS];
    double be2[CL_KS], dd[CL_KS], yv[CL_KS];
    double ysl[2][CL_RPMAX];     // this CTA's slice of the current vector (ping) and of the matvec result (pong)
    double wred[CL_WARPS];       // per-warp partials of the fused reductions (alpha in the matvec, norm in the update)
    double sv[CL_RPMAX];         // D^-1/2 of the slice
    double red[32];
    double bounds[4];
    double gb[3];
    int cnts[CL_THREADS];
    long long tmark;             // debug phase clock: end of the multisection rounds inside cluster_tridiag
    double prev_res;             // adaptive check schedule: residual estimate and step of the last check
    int prev_k;
    int next_check;
};

// Adaptive convergence checks.  The fixed schedule (every 16 steps) runs 8 steps past convergence on average
// (14 % of the steps of the benchmark workload; no node converges before step 38).  Here the residual estimates
// of the last two checks give a geometric rate, the next check is placed at CL_CHECK_SAFETY x the predicted
// number of steps (Lanczos converges faster than geometrically at the end), clamped to [LO, HI].  numpy model
// over 87 nodes: 1.009 x the minimum number of steps (fixed 16: 1.14) with the same number of checks.
constexpr int CL_CHECK_FIRST = 28;
constexpr int CL_CHECK_LO = 2;
constexpr int CL_CHECK_HI = 12;
constexpr double CL_CHECK_SAFETY = 0.7;

__device__ __forceinline__ double block_sum_512(double v, double* red) {
    v = warp_sum(v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < CL_WARPS; ++i) t += red[i];
    return t;
}
__device__ __forceinline__ double block_min_512(double v, double* red) {
    v = warp_min(v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double t = red[0];
#pragma unroll
    for (int i = 1; i < CL_WARPS; ++i) t = fmin(t, red[i]);
    return t;
}

// Sturm count without divisions: the characteristic polynomials of the leading blocks,
// p_i = (a_i - x) p_{i-1} - b_{i-1}^2 p_{i-2}, change sign between consecutive i exactly where the pivots
// q_i = p_i / p_{i-1} of sturm_count() are negative.  One dependent FMA per element instead of one division
// (a float64 division is a ~12-instruction dependent sequence; the convergence checks were 8-13 % of the CTA time).
// |a_i - x| <= 2.3 and b^2 <= 1 bound the growth to 3.3x per step, so the pair is rescaled every 4 elements;
// an exact zero takes the sign opposite to its predecessor (the q = -pivmin rule).
// Signs are compared on the high words (integer pipe): an exact +-0 then takes the sign of its zero, and the total
// over the two steps around it is the same as with the q = -pivmin rule (p_{i+1} = -b^2 p_{i-1} there).
// al and be2 must be readable up to index k + 3 (CL_KS = CL_KMAX + 4 entries): the loads of the next group of four
// are unconditional so that they sit in front of the dependent FMA chain of the current one.
__device__ __forceinline__ int sturm_count_poly(const double* al, const double* be2, int k, double x) {
    double pm = 1.0, p = al[0] - x;
    int cnt = (int)((unsigned)__double2hiint(p) >> 31);       // p_{-1} = 1 > 0
#define STURM_STEP(c, e) do {                                                             \
        const double pn_ = fma((c), p, -((e) * pm));                                      \
        cnt += (int)((unsigned)(__double2hiint(pn_) ^ __double2hiint(p)) >> 31);          \
        pm = p; p = pn_;                                                                  \
    } while (0)
    int i = 1;
    double a0 = al[1], a1 = al[2], a2 = al[3], a3 = al[4], b0 = be2[0], b1 = be2[1], b2 = be2[2], b3 = be2[3];
    while (i + 3 < k) {
        const double c0 = a0 - x, c1 = a1 - x, c2 = a2 - x, c3 = a3 - x;
        const double e0 = b0, e1 = b1, e2 = b2, e3 = b3;
        i += 4;
        a0 = al[i]; a1 = al[i + 1]; a2 = al[i + 2]; a3 = al[i + 3];
        b0 = be2[i - 1]; b1 = be2[i]; b2 = be2[i + 1]; b3 = be2[i + 2];
        STURM_STEP(c0, e0); STURM_STEP(c1, e1); STURM_STEP(c2, e2); STURM_STEP(c3, e3);
        const int ex = (__double2hiint(p) >> 20) & 0x7ff;          // biased exponent
        if (ex > 1023 + 256) { p *= 0x1p-256; pm *= 0x1p-256; }
        else if (ex < 1023 - 256 && ex != 0) { p *= 0x1p256; pm *= 0x1p256; }
    }
    if (i < k) STURM_STEP(a0 - x, b0);                             // at most three elements left: no rescaling needed
    if (i + 1 < k) STURM_STEP(a1 - x, b1);
    if (i + 2 < k) STURM_STEP(a2 - x, b2);
#undef STURM_STEP
    return cnt;     // number of eigenvalues < x
}

// tridiagonal analysis by the whole CTA (512 threads): top two eigenvalues by multisection with 256
// shifts each (257^8 > 2^64), eigenvector of the largest by a twisted factorisation (thread 0).
// Returns the residual estimate beta_{k-1} |y_{k-1}|; th[0], th[1] = eigenvalues; S.yv = eigenvector.
__device__ double cluster_tridiag(ClusterShared& S, int k, double* th, bool poly, int sh) {
    const int tid = threadIdx.x;
    // Gershgorin bracket of the spectrum and the pivot floor, by the whole CTA
    double lo_t = 1e300, hi_t = -1e300, bm_t = 0.0;
    for (int i = tid; i < k; i += CL_THREADS) {
        const double b = S.beta[i];
        S.be2[i] = b * b;
        const double rad = (i > 0 ? fabs(S.beta[i - 1]) : 0.0) + (i < k - 1 ? fabs(b) : 0.0);
        lo_t = fmin(lo_t, S.alpha[i] - rad);
        hi_t = fmax(hi_t, S.alpha[i] + rad);
        if (i < k - 1) bm_t = fmax(bm_t, b * b);
    }
    lo_t = block_min_512(lo_t, S.red);
    hi_t = -block_min_512(-hi_t, S.red);
    bm_t = -block_min_512(-bm_t, S.red);
    if (tid == 0) {
        double w = fmax(fmax(fabs(lo_t), fabs(hi_t)), 1e-300);
        S.gb[0] = lo_t - 1e-10 * w;
        S.gb[1] = hi_t + 1e-10 * w;
        S.gb[2] = fmax(bm_t, 1.0) * 1.0020841800044864e-292;
    }
    __syncthreads();
    const double glo = S.gb[0], ghi = S.gb[1], pivmin = S.gb[2];
    {
        // sh shifts per eigenvalue and round.  256 (all 512 threads) resolves 8 bits per round, 128 resolves 7:
        // 129^8 > 2^56 still brackets a float64 in 8 rounds, with half as many warps competing for the FP64 pipe.
        const int which = tid / sh, t256 = tid % sh;
        const bool act = which < 2;
        const int m = k - 1 - which;
        double lo = glo, hi = ghi;
        const int rounds = 8;                       // full float64 resolution (the residual estimate below is only
                                                    // as good as the eigenvalue)
        const double den = 1.0 / (double)(sh + 1);
        for (int round = 0; round < rounds; ++round) {
            if (act) {
                double x = lo + (hi - lo) * ((double)(t256 + 1) * den);
                S.cnts[tid] = (m < 0) ? 0 : poly ? sturm_count_poly(S.alpha, S.be2, k, x) : sturm_count(S.alpha, S.be2, k, x, pivmin);
            }
            __syncthreads();
            if (act && t256 < 32) {          // one warp per eigenvalue finds the last shift with count <= m
                int best = -1;
                for (int i = t256; i < sh; i += 32) if (S.cnts[which * sh + i] <= m) best = max(best, i);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
                if (t256 == 0) {
                    S.bounds[2 * which] = (best >= 0) ? lo + (hi - lo) * ((double)(best + 1) * den) : lo;
                    S.bounds[2 * which + 1] = (best < sh - 1) ? lo + (hi - lo) * ((double)(best + 2) * den) : hi;
                }
            }
            __syncthreads();
            if (act) { lo = S.bounds[2 * which]; hi = S.bounds[2 * which + 1]; }
            __syncthreads();
        }
    }
    const double th1 = 0.5 * (S.bounds[0] + S.bounds[1]);
    const double th2 = (k > 1) ? 0.5 * (S.bounds[2] + S.bounds[3]) : -1e300;
    if (tid == 0) {
        S.tmark = clock64();
        // Eigenvector of th1 from the factorisation of T - th1 I twisted at the FIRST index: bottom-up
        // pivots p_k = a_k - th, p_i = a_i - th - b_i^2 / p_{i+1}.  The trailing blocks of T do not contain
        // the converged part of the Krylov space, their eigenvalues stay below th1 (interlacing), so every
        // pivot is safely negative and the recurrence z_1 = 1, z_{i+1} = -b_i z_i / p_{i+1} is stable; it
        // replaces a pivoted LU plus inverse iteration (2k divisions, agreement with LAPACK to 1e-14).
        // S.dd holds the RECIPROCAL pivots, so the vector recurrence below has no division of its own.
        double d = S.alpha[k - 1] - th1;
        if (fabs(d) < pivmin) d = -pivmin;
        double r = 1.0 / d;
        S.dd[k - 1] = r;
        for (int i = k - 2; i >= 0; --i) {
            d = fma(-S.be2[i], r, S.alpha[i] - th1);
            if (fabs(d) < pivmin) d = -pivmin;
            r = 1.0 / d;
            S.dd[i] = r;
        }
        double z = 1.0, ss = 1.0;
        S.yv[0] = 1.0;
        for (int i = 0; i < k - 1; ++i) {
            z = -(S.beta[i] * S.dd[i + 1]) * z;
            if (fabs(z) > 1e150) {                       // start vector almost orthogonal to the Ritz vector
                for (int j = 0; j <= i; ++j) S.yv[j] *= 1e-150;
                z *= 1e-150;
                ss *= 1e-300;
            }
            S.yv[i + 1] = z;
            ss = fma(z, z, ss);
        }
        S.gb[1] = 1.0 / sqrt(ss);
    }
    __syncthreads();
    {
        const double inv = S.gb[1];
        for (int i = tid; i < k; i += CL_THREADS) S.yv[i] *= inv;
    }
    __syncthreads();
    if (tid == 0) S.gb[0] = fabs(S.beta[k - 1] * S.yv[k - 1]);
    __syncthreads();
    th[0] = th1;
    th[1] = th2;
    return S.gb[0];
}

template <int C>
__device__ __forceinline__ void cl_sync(cg::cluster_group& cl) {
    if (C > 1) cl.sync(); else __syncthreads();
}

// Basis rows restricted to this CTA's slice: the first rows_s rows live in shared memory, later rows
// (long runs) spill to the global basis V.
struct SliceBasis {
    double* smem;          // rows_s rows of stride nrp
    double* glob;          // e.V + first global position of the slice; row stride P
    size_t P;
    int rows_s;
    int nrp;
    __device__ __forceinline__ double* row(int j) const {
        return (j < rows_s) ? smem + (size_t)j * nrp : glob + (size_t)j * P;
    }
};

// partial dots of the slice vector y with basis rows [0, rows): S.hpart[buf][j].  One warp per row,
// all loads of a row in flight at once (nr <= CL_RPMAX = 16 * 32).
__device__ __forceinline__ void cl_partial_dots(ClusterShared& S, const double* __restrict__ y, const SliceBasis& B,
                                                int rows, int nr, int buf) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int M = (CL_RPMAX + 31) / 32;
    double yr[M];
#pragma unroll
    for (int m = 0; m < M; ++m) { int i = lane + 32 * m; yr[m] = (i < nr) ? y[i] : 0.0; }
    for (int j = warp; j < rows; j += CL_WARPS) {
        const double* vr = B.row(j);
        double t[M];
#pragma unroll
        for (int m = 0; m < M; ++m) { int i = lane + 32 * m; t[m] = (i < nr) ? vr[i] : 0.0; }
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < M; ++m) s += t[m] * yr[m];
        s = warp_sum(s);
        if (lane == 0) S.hpart[buf][j] = s;
    }
}

template <int C>
__device__ __forceinline__ void cl_reduce_h(cg::cluster_group& cl, ClusterShared& S, int rows, int buf) {
    for (int j = threadIdx.x; j < rows; j += CL_THREADS) {
        double h = 0.0;
        if (C > 1) {
#pragma unroll
            for (int r = 0; r < C; ++r) h += cl.map_shared_rank(&S.hpart[buf][0], r)[j];     // rank order
        } else {
            h = S.hpart[buf][j];
        }
        S.hs[j] = h;
    }
    __syncthreads();
}

// y_i -= sum_j hs[j] V[j][i] for the slice; returns this thread's share of |y|^2 (the norm rides on the update pass)
__device__ __forceinline__ double cl_update_norm(ClusterShared& S, double* __restrict__ y, const SliceBasis& B, int rows, int nr) {
    double q = 0.0;
    for (int i = threadIdx.x; i < nr; i += CL_THREADS) {
        double v = y[i];
        int j = 0;
        for (; j + 8 <= rows; j += 8) {
            double t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = B.row(j + u)[i];
#pragma unroll
            for (int u = 0; u < 8; ++u) v -= S.hs[j + u] * t[u];
        }
        for (; j < rows; ++j) v -= S.hs[j] * B.row(j)[i];
        y[i] = v;
        q = fma(v, v, q);
    }
    return q;
}

// sum over the CTA of one value per thread: warp partials in S.wred, ONE block barrier, fixed order
__device__ __forceinline__ double cl_block_sum1(ClusterShared& S, double v) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) S.wred[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < CL_WARPS; ++i) t += S.wred[i];
    return t;
}

// float -> double widening on the integer pipe, for weights in {0} U [2^-126, 2): the bits move 29 places and the
// exponent is re-biased; +0 becomes 2^-127 (5.9e-39), which every later float64 sum absorbs exactly.
// cvt.f64.f32 runs on the XU pipe at 4 lanes per clock and scheduler: ncu shows it half busy during the matvec and
// the warps waiting on it, so in the MIX variant every second element is widened with three integer instructions.
__device__ __forceinline__ double widen_int(float f) {
    const unsigned b = __float_as_uint(f);
    return __hiloint2double((int)((b >> 3) + 0x38000000u), (int)(b << 29));
}
template <bool MIX>
__device__ __forceinline__ double widen_alt(float f) { return MIX ? widen_int(f) : (double)f; }

// ---- guarded matvec (MODE 0, stage entry points on the caller's W): register-staged loads, entries outside the block
// are selected away (they may be uninitialised or belong to other nodes).
// yout_i = s_i * invb * (sum_j w_ij z_j + z_i) for the rows [r0, r0+nr); returns (lane 0) the warp's share of
// alpha = v . (M v), v_i = yin_i * invb.
__device__ __forceinline__ double cl_matvec_guard(ClusterShared& S, const double* zs, const double* __restrict__ yin,
                                                  double* __restrict__ yout, const NodeView& v, int r0, int nr, int pad,
                                                  double invb) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c_lo = v.ro, c_hi = v.ro + v.n, a0 = c_lo & ~3;
    double pa = 0.0;
    for (int rb = warp * 2; rb < nr; rb += CL_WARPS * 2) {
        const float* rpt[2];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) rpt[rr] = v.W + (size_t)(v.ro + r0 + min(rb + rr, nr - 1)) * v.ld;
        double acc[2] = {0.0, 0.0};
        for (int c = a0 + lane * 4; c < c_hi; c += 512) {
            float4 w[2][4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int cg_ = c + 128 * g;
                const bool has = cg_ < c_hi;
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) w[rr][g] = has ? ld_stream4(rpt[rr] + cg_) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int cg_ = c + 128 * g;
                if (cg_ < c_hi) {
                    const double2 z0 = *reinterpret_cast<const double2*>(&zs[cg_ - a0]);
                    const double2 z1 = *reinterpret_cast<const double2*>(&zs[cg_ - a0 + 2]);
                    const bool v0 = (cg_ >= c_lo), v1 = (cg_ + 1 >= c_lo) & (cg_ + 1 < c_hi);
                    const bool v2 = (cg_ + 2 >= c_lo) & (cg_ + 2 < c_hi), v3 = (cg_ + 3 >= c_lo) & (cg_ + 3 < c_hi);
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        double q0 = (v0 ? (double)w[rr][g].x : 0.0) * z0.x + (v1 ? (double)w[rr][g].y : 0.0) * z0.y;
                        double q1 = (v2 ? (double)w[rr][g].z : 0.0) * z1.x + (v3 ? (double)w[rr][g].w : 0.0) * z1.y;
                        acc[rr] += q0 + q1;
                    }
                }
            }
        }
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const double t = warp_sum(acc[rr]);
            const int i = rb + rr;
            if (lane == 0 && i < nr) {
                const double yn = S.sv[i] * invb * (t + zs[r0 + i + pad]);      // (w + I) z
                yout[i] = yn;
                pa = fma(yin[i] * invb, yn, pa);
            }
        }
    }
    return pa;
}

// ---- TMA-fed matvec (MODE 4 / 6): every warp owns a ring of RING_ST stages in shared memory; a stage holds up to
// RING_COLS columns of the warp's two current rows (4 KB).  Lane 0 issues one bulk copy per row and stage
// (cp.async.bulk, completion counted on the stage's mbarrier); the warp waits for the stage, multiplies it with z
// from shared memory and hands the slot back.  Registers limit a register-staged form to 4 KB per warp in flight and
// only in bursts (issue 8 loads, wait, multiply: ncu put 32 % of that kernel's stall samples on the first use of the
// loaded registers); the ring keeps 4-8 KB per warp in flight all the time.  It takes the shared memory that held
// basis rows: Gram-Schmidt then reads the basis from L2, which costs less than the matvec gains.
// W does not change between steps, so the first stages of the NEXT matvec are fetched during Gram-Schmidt.
// The blocks come from k_gather_blocks_cur / k_zero_blocks, which zero the <= 3-column fringe of every block, and
// zs is zero there: no selects.  (Measured and dropped, DESIGN.md 5a: 1 KB stages x 4, register-staged loads with L2
// bulk prefetch, per-lane prefetch, L2 prefetch of the next matvec's stages, all loads of a stage hoisted.)
constexpr int RING_ST = 2;
constexpr int RING_COLS = 512;                               // floats per row and stage
constexpr int RING_STAGE_FLOATS = 2 * RING_COLS;
constexpr int RING_WARP_FLOATS = RING_ST * RING_STAGE_FLOATS;
constexpr int RING_BYTES = CL_WARPS * RING_WARP_FLOATS * 4;      // 128 KB

struct RingGeom { int a0, width, nseg, npass, T; };
__device__ __forceinline__ RingGeom ring_geom(const NodeView& v, int nr) {
    RingGeom q;
    const int warp = threadIdx.x >> 5;
    q.a0 = v.ro & ~3;
    q.width = (v.ro + v.n - q.a0 + 3) & ~3;            // columns fetched per row (multiple of 4)
    q.nseg = (q.width + RING_COLS - 1) / RING_COLS;
    q.npass = (warp * 2 < nr) ? (nr - warp * 2 + CL_WARPS * 2 - 1) / (CL_WARPS * 2) : 0;
    q.T = q.npass * q.nseg;
    return q;
}
// lane 0: fetch stage (p, sg) of the matvec into the slot of global stage index gi
__device__ __forceinline__ void ring_issue(const NodeView& v, const RingGeom& q, int r0, int nr, int p, int sg,
                                           float* ring_w, uint64_t* bars_w, uint32_t gi) {
    const int warp = threadIdx.x >> 5;
    const int rb = warp * 2 + p * (CL_WARPS * 2);
    const int c = sg * RING_COLS;
    const uint32_t bytes = (uint32_t)min(RING_COLS, q.width - c) * 4u;
    const uint32_t slot = gi % RING_ST;
    float* dst = ring_w + slot * RING_STAGE_FLOATS;
    uint64_t* bar = bars_w + slot;
    mbarrier_expect_tx(bar, 2u * bytes);
    bulk_copy_g2s(dst, v.W + (size_t)(v.ro + r0 + min(rb, nr - 1)) * v.ld + q.a0 + c, bytes, bar);
    bulk_copy_g2s(dst + RING_COLS, v.W + (size_t)(v.ro + r0 + min(rb + 1, nr - 1)) * v.ld + q.a0 + c, bytes, bar);
}
// the first min(RING_ST, T) stages of a matvec (before the first step and after every matvec)
__device__ __forceinline__ void ring_prologue(const NodeView& v, const RingGeom& q, int r0, int nr, float* ring_w,
                                              uint64_t* bars_w, uint32_t g) {
    if ((threadIdx.x & 31) == 0) {
        int p = 0, sg = 0;
        for (int t = 0; t < min(RING_ST, q.T); ++t) {
            ring_issue(v, q, r0, nr, p, sg, ring_w, bars_w, g + t);
            if (++sg == q.nseg) { sg = 0; ++p; }
        }
    }
}
__device__ __forceinline__ void ring_drain(const RingGeom& q, uint64_t* bars_w, uint32_t g) {
    for (int t = 0; t < min(RING_ST, q.T); ++t) mbarrier_wait(bars_w + ((g + t) % RING_ST), ((g + t) / RING_ST) & 1u);
}

// returns (lane 0) the warp's share of alpha = v . (M v), v_i = yin_i * invb
template <bool MIX>
__device__ __forceinline__ double cl_matvec_ring(ClusterShared& S, const double* zs, const double* __restrict__ yin,
                                                 double* __restrict__ yout, const NodeView& v, const RingGeom& q,
                                                 int r0, int nr, int pad, double invb, float* ring_w, uint64_t* bars_w,
                                                 uint32_t& g) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int ip = 0, isg = 0;                               // next stage to issue = stage t + RING_ST
    for (int t = 0; t < min(RING_ST, q.T); ++t) if (++isg == q.nseg) { isg = 0; ++ip; }
    int t = 0;
    double pa = 0.0;
    for (int p = 0; p < q.npass; ++p) {
        double acc0 = 0.0, acb0 = 0.0, acc1 = 0.0, acb1 = 0.0;
        for (int sg = 0; sg < q.nseg; ++sg, ++t) {
            const uint32_t slot = g % RING_ST;
            mbarrier_wait(bars_w + slot, (g / RING_ST) & 1u);
            const float* buf = ring_w + slot * RING_STAGE_FLOATS;
            auto fma8 = [&](const float4& wa, const float4& wb, const double2& z0, const double2& z1) {
                acc0 = fma((double)wa.x, z0.x, acc0); acb0 = fma(widen_alt<MIX>(wa.y), z0.y, acb0);
                acc0 = fma((double)wa.z, z1.x, acc0); acb0 = fma(widen_alt<MIX>(wa.w), z1.y, acb0);
                acc1 = fma((double)wb.x, z0.x, acc1); acb1 = fma(widen_alt<MIX>(wb.y), z0.y, acb1);
                acc1 = fma((double)wb.z, z1.x, acc1); acb1 = fma(widen_alt<MIX>(wb.w), z1.y, acb1);
            };
#pragma unroll
            for (int gq = 0; gq < RING_COLS / 128; ++gq) {
                const int cs = 128 * gq + 4 * lane;                // column inside the stage
                const int cofs = sg * RING_COLS + cs;              // column - a0
                if (cofs < q.width) {
                    const float4 wa = *reinterpret_cast<const float4*>(buf + cs);
                    const float4 wb = *reinterpret_cast<const float4*>(buf + RING_COLS + cs);
                    const double2 z0 = *reinterpret_cast<const double2*>(&zs[cofs]);
                    const double2 z1 = *reinterpret_cast<const double2*>(&zs[cofs + 2]);
                    fma8(wa, wb, z0, z1);
                }
            }
            __syncwarp();
            ++g;
            if (lane == 0 && t + RING_ST < q.T) {
                ring_issue(v, q, r0, nr, ip, isg, ring_w, bars_w, g + RING_ST - 1);
                if (++isg == q.nseg) { isg = 0; ++ip; }
            }
        }
        const int rb = warp * 2 + p * (CL_WARPS * 2);
        const double t0 = warp_sum(acc0 + acb0), t1 = warp_sum(acc1 + acb1);
        if (lane == 0) {
            const double y0 = S.sv[rb] * invb * (t0 + zs[r0 + rb + pad]);               // (w + I) z
            yout[rb] = y0;
            pa = fma(yin[rb] * invb, y0, pa);
            if (rb + 1 < nr) {
                const double y1 = S.sv[rb + 1] * invb * (t1 + zs[r0 + rb + 1 + pad]);
                yout[rb + 1] = y1;
                pa = fma(yin[rb + 1] * invb, y1, pa);
            }
        }
    }
    ring_prologue(v, q, r0, nr, ring_w, bars_w, g);
    return pa;
}

// ---- shared-memory sparse matvec (MODE 7, ANCUTS_OPT_MATVEC = 1): W is 99.5 % zeros (27-66 stored entries per row
// against n <= 4096 columns), yet the dense form streams every block from HBM once per Lanczos step.  Here the CTA reads
// its row slice of the dense block TWICE at the start of the node (count, then fill) and keeps it as CSR in shared memory
// (float32 value + uint16 column, rows in slice order, columns ascending) for all the steps: 8 n^2 bytes per node instead
// of k (4 n^2 + 8 n).  No ring, so the basis rows get the shared memory the ring took.  Same float64 products; the sum of a
// row is taken by four lanes over its entries round-robin and combined (l0 + l1) + (l2 + l3): a fixed order, but not
// the dense kernels' order, so alpha/beta differ from the dense form in the last bits (parity is against the oracle).
// A slice that does not fit (dense little nodes far above 100 entries per row) sends the node to the grid-wide path.
struct SparseSlice {
    const float* val; const unsigned short* col; const int* ptr; int nnz;
};
constexpr int SP_LANES = 4;

// entries of row `rowp` inside the block's aligned window, counted (FILL = false) or written (FILL = true) by one warp
template <bool FILL>
__device__ __forceinline__ int sp_scan_row(const float* __restrict__ rowp, int a0, int c_lo, int c_hi, int lane,
                                           float* val, unsigned short* col, int base) {
    int total = 0;
    for (int c = a0 + lane * 4; c - lane * 4 < c_hi; c += 128) {          // warp-uniform trip count
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < c_hi) w = ld_stream4(rowp + c);
        const float in[4] = {w.x, w.y, w.z, w.w};
        unsigned m = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) m |= (in[q] != 0.0f && c + q >= c_lo && c + q < c_hi) ? (1u << q) : 0u;
        const int cnt = __popc(m);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (FILL) {
            int pos = base + total + incl - cnt;
#pragma unroll
            for (int q = 0; q < 4; ++q) if (m & (1u << q)) { val[pos] = in[q]; col[pos] = (unsigned short)(c + q - c_lo); ++pos; }
        }
        total += __shfl_sync(0xffffffffu, incl, 31);
    }
    return total;
}

// returns (every lane) the thread's share of alpha = v . (M v), v_i = yin_i * invb; the caller reduces it over the warp
__device__ __forceinline__ double cl_matvec_sparse(ClusterShared& S, const double* zs, const double* __restrict__ yin,
                                                   double* __restrict__ yout, const SparseSlice& sp, int r0, int nr, int pad,
                                                   double invb) {
    const int tid = threadIdx.x, sub = tid & (SP_LANES - 1);
    double pa = 0.0;
    for (int base = 0; base < nr; base += CL_THREADS / SP_LANES) {         // warp-uniform loop: the shuffles need all lanes
        const int i = base + tid / SP_LANES;
        double acc = 0.0;
        if (i < nr) {
            const int b = sp.ptr[i], e = sp.ptr[i + 1];
            for (int q = b + sub; q < e; q += SP_LANES) acc = fma((double)sp.val[q], zs[sp.col[q] + pad], acc);
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (sub == 0 && i < nr) {
            const double y = S.sv[i] * invb * (acc + zs[r0 + i + pad]);      // (w + I) z
            yout[i] = y;
            pa = fma(yin[i] * invb, y, pa);
        }
    }
    return pa;
}

// grid: count * C CTAs, cluster (C,1,1); ids[cluster index] = active slot.
// MODE 0: guarded register-staged matvec (caller's W read in place); 4: TMA ring; 6: TMA ring + integer widening of
// every second element (weights in {0} U [2^-126, 2): the library's own affinities); 7: row slice as CSR in shared memory.
template <int C, int MODE>
__global__ void __launch_bounds__(CL_THREADS, 1)
k_lanczos_cluster(Eng e, int cur, const int* __restrict__ ids, int dyn_doubles) {
    extern __shared__ __align__(16) double zs[];   // z = S v for the whole node on the 16-byte window of W's columns;
                                                   // behind it: the warps' W rings, then the first rows of the basis,
                                                   // restricted to this CTA's slice
    __shared__ ClusterShared S;
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (C > 1) ? (int)cl.block_rank() : 0;
    const int a = ids[blockIdx.x / C];
    const NodeView v = node_view(e, e.a_rid[a], cur);
    const int n = v.n;
    const int rp = (n + C - 1) / C;
    const int r0 = min(n, rank * rp), r1 = min(n, r0 + rp);
    const int nr = r1 - r0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t P = (size_t)e.P;
    const int g0 = v.start + r0;                       // global position of the slice
    const int pad = v.ro & 3;                          // zs[j + pad] <-> column v.ro + j
    constexpr bool RING = (MODE == 4 || MODE == 6);
    constexpr bool SP = (MODE == 7);
    const int nz = (n + 8 + 3) & ~3;                   // doubles used by zs
    SparseSlice sp;
    sp.val = nullptr; sp.col = nullptr; sp.ptr = nullptr; sp.nnz = 0;
    int sp_doubles = 0;
    if (SP) {
        // ---- build the CSR slice: count the entries per row, scan, fill (two passes over the slice) ----
        const int c_lo = v.ro, c_hi = v.ro + n, a0 = c_lo & ~3;
        int* rowcnt = S.cnts;                          // free until the first convergence check
        __shared__ int sp_wtot[CL_WARPS + 1];
        for (int i = warp; i < nr; i += CL_WARPS) {
            const int t = sp_scan_row<false>(v.W + (size_t)(v.ro + r0 + i) * v.ld, a0, c_lo, c_hi, lane, nullptr, nullptr, 0);
            if (lane == 0) rowcnt[i] = t;
        }
        __syncthreads();
        const int mine = (tid < nr) ? rowcnt[tid] : 0;
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) sp_wtot[warp] = incl;
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            for (int w = 0; w < CL_WARPS; ++w) { int t = sp_wtot[w]; sp_wtot[w] = run; run += t; }
            sp_wtot[CL_WARPS] = run;
        }
        __syncthreads();
        const int total = sp_wtot[CL_WARPS];
        int* ptr = reinterpret_cast<int*>(zs + nz);
        const int ptr_d = (nr + 1 + 1) / 2;                                  // doubles
        const int val_d = (total + 1) / 2, col_d = (total + 3) / 4;
        sp_doubles = ptr_d + val_d + col_d;
        double misfit = (nz + sp_doubles > dyn_doubles) ? 1.0 : 0.0;
        if (C > 1) {                                                         // the whole cluster takes the same decision
            if (tid == 0) S.npart[1] = misfit;
            cl.sync();
            misfit = 0.0;
#pragma unroll
            for (int r = 0; r < C; ++r) misfit += cl.map_shared_rank(&S.npart[0], r)[1];
            cl.sync();                                                       // peers have read the flag: the slot is free again
        }
        if (misfit != 0.0) {
            if (rank == 0 && tid == 0) { e.a_done[a] = DONE_NO; e.a_path[a] = 1; atomicAdd(&e.ctr[4], 1); }
            return;                                                          // grid-wide path (kernels_lanczos.cuh)
        }
        float* val = reinterpret_cast<float*>(zs + nz + ptr_d);
        unsigned short* col = reinterpret_cast<unsigned short*>(zs + nz + ptr_d + val_d);
        if (tid < nr) ptr[tid] = sp_wtot[warp] + incl - mine;
        if (tid == 0) ptr[nr] = total;
        __syncthreads();
        for (int i = warp; i < nr; i += CL_WARPS)
            sp_scan_row<true>(v.W + (size_t)(v.ro + r0 + i) * v.ld, a0, c_lo, c_hi, lane, val, col, ptr[i]);
        sp.val = val; sp.col = col; sp.ptr = ptr; sp.nnz = total;
        __syncthreads();
    }
    SliceBasis B;
    {
        B.nrp = (max(nr, 1) + 3) & ~3;
        const int nfront = RING ? RING_BYTES / 8 : sp_doubles;              // the warps' W rings or the CSR slice sit in front
        B.smem = zs + nz + nfront;
        B.rows_s = max(0, (dyn_doubles - nz - nfront) / B.nrp);
        B.glob = e.V + g0;
        B.P = P;
    }
    const int kcap = min(min(CL_KMAX, e.kmax), n - 1);
    __shared__ uint64_t ring_bar[CL_WARPS * RING_ST];
    float* ring_w = reinterpret_cast<float*>(zs + nz) + (size_t)warp * RING_WARP_FLOATS;
    uint64_t* bars_w = ring_bar + warp * RING_ST;
    const RingGeom rq = ring_geom(v, nr);
    uint32_t ring_g = 0;                               // stages consumed by this warp so far
    if (RING) {
        if (lane == 0) {
            for (int i = 0; i < RING_ST; ++i) mbarrier_init(bars_w + i, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        ring_prologue(v, rq, r0, nr, ring_w, bars_w, ring_g);     // W is ready: start streaming during the set-up
    }
    // ---- init (every CTA redundantly: bit-identical scalars, no exchange needed) ----
    double s = 0.0;
    for (int i = tid; i < n; i += CL_THREADS) s += e.deg[v.start + i];
    const double vol = block_sum_512(s, S.red);
    const double ivol = 1.0 / sqrt(vol);
    s = 0.0;
    const StartVec sv0(e, v.chunk, v.start);
    for (int i = tid; i < n; i += CL_THREADS) s += sqrt(e.deg[v.start + i]) * ivol * sv0.at(v.start + i, i);
    const double dot = block_sum_512(s, S.red);
    s = 0.0;
    for (int i = tid; i < n + 8; i += CL_THREADS) {
        int j = i - pad;
        double z = 0.0;
        if (j >= 0 && j < n) {
            double u = sqrt(e.deg[v.start + j]) * ivol;
            double x = sv0.at(v.start + j, j) - dot * u;
            z = e.sinv[v.start + j] * x;
            s += x * x;
        }
        zs[i] = z;
    }
    double bprev = sqrt(block_sum_512(s, S.red));
    double* yc = S.ysl[0];                               // current (unnormalised) Lanczos vector of the slice
    double* yn = S.ysl[1];                               // matvec result, orthogonalised in place
    for (int i = tid; i < nr; i += CL_THREADS) {
        int j = r0 + i;
        double u = sqrt(e.deg[v.start + j]) * ivol;
        B.row(0)[i] = u;                                 // basis row 0 = u1
        yc[i] = sv0.at(v.start + j, j) - dot * u;
        S.sv[i] = e.sinv[v.start + j];
    }
    __syncthreads();

    int k = 0;
    int conv = 0;
    double th[2] = {0.0, 0.0};
    const bool adapt = e.check_adapt != 0;
    if (tid == 0) { S.prev_res = 1.0; S.prev_k = 0; S.next_check = adapt ? CL_CHECK_FIRST : e.check_every; }
    // optional phase clock (debug): cycles of thread 0 of rank 0 between the block-wide syncs that end each phase
    // (a barrier's wait is charged to the phase AFTER it: BAR.SYNC does not block at issue)
    const bool prof = (e.dbg != nullptr) && tid == 0 && rank == 0;
    long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = prof ? clock64() : 0;
#define CL_PHASE(i) do { if (prof) { long long t_ = clock64(); tph[i] += t_ - tlast; tlast = t_; } } while (0)
    while (true) {
        const double invb = 1.0 / bprev;
        // basis row k+1 = current vector (slice).  The matvec below only reads yc and writes yn: no barrier in between.
        double* vk = B.row(k + 1);
        for (int i = tid; i < nr; i += CL_THREADS) vk[i] = yc[i] * invb;
        CL_PHASE(0);
        // ---- matvec of the slice; alpha = v_k . (M v_k) rides on its epilogue ----
        double pa;
        if (RING) pa = cl_matvec_ring<MODE == 6>(S, zs, yc, yn, v, rq, r0, nr, pad, invb, ring_w, bars_w, ring_g);
        else if (SP) pa = warp_sum(cl_matvec_sparse(S, zs, yc, yn, sp, r0, nr, pad, invb));
        else pa = cl_matvec_guard(S, zs, yc, yn, v, r0, nr, pad, invb);
        if (lane == 0) S.wred[warp] = pa;
        __syncthreads();
        CL_PHASE(1);
        double a1 = 0.0;
#pragma unroll
        for (int i = 0; i < CL_WARPS; ++i) a1 += S.wred[i];
        if (C > 1) {
            if (tid == 0) S.npart[1] = a1;
            cl.sync();
            a1 = 0.0;
#pragma unroll
            for (int r = 0; r < C; ++r) a1 += cl.map_shared_rank(&S.npart[0], r)[1];      // rank order
        }
        const int rows = k + 2;
        // ---- three-term recurrence first, then ONE classical Gram-Schmidt pass against u1 and every Lanczos
        //      vector.  In exact arithmetic w - alpha v_k - beta v_{k-1} is already orthogonal to the basis, so the
        //      full pass only removes rounding-level components and plays the role of the SECOND pass of "twice is
        //      enough": orthogonality stays at 2e-15 like CGS2 (numpy model, 151 nodes, identical steps and cuts),
        //      with two sweeps over the basis instead of four.  A single CGS pass WITHOUT the three-term part
        //      loses orthogonality within dozens of steps (measured in round 1 and again in the model). ----
        {
            const double bk = (k > 0) ? S.beta[k - 1] : 0.0;
            const double* vp = B.row(k);                                                // v_{k-1} (u1 when k = 0: bk = 0)
            for (int i = tid; i < nr; i += CL_THREADS) yn[i] = (yn[i] - a1 * vk[i]) - bk * vp[i];   // own elements of vk, vp
        }
        __syncthreads();
        CL_PHASE(2);
        cl_partial_dots(S, yn, B, rows, nr, 1);
        cl_sync<C>(cl);
        cl_reduce_h<C>(cl, S, rows, 1);
        const double a2 = S.hs[rows - 1];
        CL_PHASE(4);
        // ---- update, squared norm, z = S y for the next matvec ----
        double q = cl_update_norm(S, yn, B, rows, nr);
        if (C > 1) { for (int i = tid; i < nr; i += CL_THREADS) e.zbuf[g0 + i] = S.sv[i] * yn[i]; }
        else { for (int i = tid; i < nr; i += CL_THREADS) zs[i + pad] = S.sv[i] * yn[i]; }
        q = cl_block_sum1(S, q);
        CL_PHASE(5);
        double nn = q;
        if (C > 1) {
            if (tid == 0) S.npart[0] = q;
            cl.sync();
            nn = 0.0;
#pragma unroll
            for (int r = 0; r < C; ++r) nn += cl.map_shared_rank(&S.npart[0], r)[0];
            // z of the whole node.  (Through global memory: a CTA reads its peers' slices from L2 at ~64 B/clk, distributed
            // shared memory moves 17-21 B/clk per SM, B300_MICROARCH.md; the cluster barrier above orders the two sides.)
            for (int i = tid; i < n; i += CL_THREADS) zs[i + pad] = __ldcg(e.zbuf + v.start + i);
        }
        const double beta = sqrt(nn);
        if (tid == 0) { S.alpha[k] = a1 + a2; S.beta[k] = beta; }
        bprev = beta;
        k += 1;
        { double* t_ = yc; yc = yn; yn = t_; }
        __syncthreads();
        CL_PHASE(6);
        const bool breakdown = beta < 1e-13;
        if (breakdown || k >= kcap || k == S.next_check) {
            double res = cluster_tridiag(S, k, th, true, CL_HALF / 2);
            if (prof) { tph[3] += S.tmark - tlast; tlast = S.tmark; }     // slot 3: multisection part of the check
            double gap = fmax(th[0] - th[1], 1e-300);
            bool c1 = (k >= n - 1) || breakdown || (res <= e.tol * gap);
            if (c1) { conv = 1; break; }
            if (k >= kcap) break;
            if (tid == 0) {                  // every CTA of the cluster holds the same alpha/beta: identical schedules
                int adv = e.check_every;
                if (adapt) {
                    adv = CL_CHECK_HI;
                    const double pr = S.prev_res;
                    if (res < pr && res > 0.0) {
                        const double rate = log(res / pr) / (double)(k - S.prev_k);          // < 0
                        const double need = CL_CHECK_SAFETY * log(e.tol * gap / res) / rate;   // > 0: not converged yet
                        adv = (int)fmin((double)CL_CHECK_HI, fmax((double)CL_CHECK_LO, ceil(need)));
                    }
                    S.prev_res = res;
                    S.prev_k = k;
                }
                S.next_check = k + adv;
            }
            __syncthreads();
            CL_PHASE(7);
        }
    }
    CL_PHASE(7);
    if (RING) ring_drain(rq, bars_w, ring_g);     // stages fetched ahead for a step that will not run
    if (prof) {
        const int ci = (C == 1) ? 0 : (C == 2) ? 1 : (C == 4) ? 2 : 3;
        for (int i = 0; i < 8; ++i) atomicAdd(&e.dbg[ci * 8 + i], (unsigned long long)tph[i]);
    }
#undef CL_PHASE
    // ---- Ritz vector of the slice, statistics for the cut kernels ----
    if (conv) {
        double xs = 0.0, xq = 0.0, xmn = 1e300, xmx = -1e300;
        for (int i = tid; i < nr; i += CL_THREADS) {
            double x = 0.0;
            for (int j = 0; j < k; ++j) x += S.yv[j] * B.row(j + 1)[i];
            e.ev[g0 + i] = x;
            xs += x; xq += x * x; xmn = fmin(xmn, x); xmx = fmax(xmx, x);
        }
        double sm_ = block_sum_512(xs, S.red);
        double sq = block_sum_512(xq, S.red);
        double mn = block_min_512(xmn, S.red);
        double mx = -block_min_512(-xmx, S.red);
        if (tid == 0) { S.spart[0] = sm_; S.spart[1] = mn; S.spart[2] = mx; S.spart[3] = sq; }
        cl_sync<C>(cl);
        if (rank == 0 && tid == 0) {
            double ts = 0.0, tmn = 1e300, tmx = -1e300, tq = 0.0;
            for (int r = 0; r < C; ++r) {
                const double* np_ = (C > 1) ? cl.map_shared_rank(&S.spart[0], r) : &S.spart[0];
                ts += np_[0]; tmn = fmin(tmn, np_[1]); tmx = fmax(tmx, np_[2]); tq += np_[3];
            }
            const int s0 = e.a_slot0[a], nch = e.a_nch[a];
            double* o = e.p_stat + (size_t)s0 * 4;
            o[0] = ts; o[1] = tmn; o[2] = tmx; o[3] = tq;
            for (int c = 1; c < nch; ++c) { double* oc = e.p_stat + (size_t)(s0 + c) * 4; oc[0] = 0.0; oc[1] = 1e300; oc[2] = -1e300; oc[3] = 0.0; }
            e.a_k[a] = k;
            e.a_kcap[a] = kcap;
            e.a_conv[a] = 1;
            e.a_theta[2 * a] = th[0];
            e.a_theta[2 * a + 1] = th[1];
            e.a_done[a] = DONE_YES;
            e.a_path[a] = 0;                    // Ritz vector and statistics are already in place
            if (!SP) atomicAdd(&e.acct[SG_MATVEC], (unsigned long long)k * (4ull * n * n + 8ull * n));
        }
        if (SP && tid == 0) {      // what this form really moves: the slice twice from HBM, then k sweeps over its entries
            atomicAdd(&e.acct[SG_MATVEC], 8ull * (unsigned long long)nr * n);
            atomicAdd(&e.acct[SG_SPARSE_STEPS], (unsigned long long)k * (unsigned long long)sp.nnz);
            atomicAdd(&e.acct[SG_SPARSE_NNZ], (unsigned long long)sp.nnz);
        }
    } else if (rank == 0 && tid == 0) {
        e.a_done[a] = DONE_NO;                  // left for the multi-launch path (more steps)
        e.a_path[a] = 1;
        atomicAdd(&e.ctr[4], 1);
    }
    cl_sync<C>(cl);                             // peers may still be reading this CTA's shared memory
}

}  // namespace ancuts
