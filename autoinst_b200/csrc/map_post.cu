// map_post.cu — "next" rows N2 and N4 of SURVEY.md §8f on the device:
//   N2  map-level merge / instance association across overlapping chunks
//         merge_chunks_unite_instances2   pipeline/utils/point_cloud/point_cloud_utils.py:387-491
//         remove_semantics                :253-287          (glue: pipeline/run_pipeline.py:197-223)
//   N4  instance metrics
//         Metrics.update_stats / filter_labels / get_tp_fp / average_precision
//                                         pipeline/metrics/metrics_class.py:61-117,137-179,181-235,296-340
//         evaluator.add_batch / get_eval  pipeline/metrics/modified_LSTQ.py:23-80
// The reference walks Python dictionaries of per-instance point arrays, chunk after chunk; here every map point keeps
// ONE slot in the concatenated arrays (chunks in file-name order), labels are dense ranks of the sorted label values
// (np.unique order, which the reference's greedy rules depend on), and each merge step is a handful of kernels over
// the 40 m crop: per-instance boxes and sizes by atomics, points-in-box counts into a dense instance x instance table,
// the reference's odd "union" (distinct SCALAR coordinate values, :457) from one sort of (value, instance) keys plus a
// binary-search intersection per candidate pair, the greedy resolve (:465-477) one thread per new instance, and
// first-occurrence-wins duplicate removal (:488-489) through a hash table that keeps the smallest point index.
// Integer / index work is exact; the few float64 expressions repeat the reference's operation order.
#include <cub/cub.cuh>
#include <algorithm>

#include "handle.cuh"

using namespace ancuts;

namespace ancuts {

static inline size_t pa_align(size_t x) { return (x + 255) / 256 * 256; }

struct PArena {
    char* base; size_t off = 0;
    template <typename T> T* take(size_t count) {
        off = pa_align(off);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

static int post_ws(ancuts_handle* h, size_t bytes) {
    if (!h->h_post) ANCUTS_CUDA(cudaMallocHost((void**)&h->h_post, 16 * sizeof(long long)));
    if (bytes <= h->post_ws_bytes) return ANCUTS_OK;
    if (h->post_ws) cudaFree(h->post_ws);
    h->post_ws = nullptr; h->post_ws_bytes = 0;
    cudaError_t err = cudaMalloc((void**)&h->post_ws, bytes);
    if (err != cudaSuccess) {
        set_error("map post-processing workspace cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(err));
        cudaGetLastError();
        return ANCUTS_ENOMEM;
    }
    h->post_ws_bytes = bytes;
    return ANCUTS_OK;
}

#define PLAUNCH(...) do { h->launches_total++; __VA_ARGS__; } while (0)

// ---- helpers ------------------------------------------------------------------------------------------------------
// total order on float64 bit patterns that agrees with numeric '<' and '==' for the values np.unique compares
// (-0.0 == 0.0: canonicalised; no NaNs in coordinates)
__device__ __forceinline__ unsigned long long f64_key(double x) {
    if (x == 0.0) x = 0.0;
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double atomic_min_f64(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        if (__longlong_as_double((long long)assumed) <= v) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
    return __longlong_as_double((long long)old);
}
__device__ __forceinline__ double atomic_max_f64(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        if (__longlong_as_double((long long)assumed) >= v) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
    return __longlong_as_double((long long)old);
}
// position of `v` in the ascending array a[0..n) (must be present)
__device__ __forceinline__ int find_sorted(const int* __restrict__ a, int n, int v) {
    int lo = 0, hi = n - 1;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}

// ---- dense ranks of label values (np.unique order) -----------------------------------------------------------------
__global__ void k_flag_new(const int* __restrict__ sorted, long long n, int* __restrict__ flag) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i == 0 || sorted[i] != sorted[i - 1]) ? 1 : 0;
}
__global__ void k_scatter_unique(const int* __restrict__ sorted, const int* __restrict__ flag, const int* __restrict__ incl,
                                 long long n, int* __restrict__ uniq, long long* __restrict__ count_out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flag[i]) uniq[incl[i] - 1] = sorted[i];
    if (i == n - 1) *count_out = incl[i];
}
__global__ void k_rank_of(const int* __restrict__ lab, long long n, const int* __restrict__ uniq, const long long* __restrict__ nu,
                          int zero_is_background, int* __restrict__ rank) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int v = lab[i];
    rank[i] = (zero_is_background && v == 0) ? -1 : find_sorted(uniq, (int)*nu, v);
}

// sorted unique values of lab[0..n) -> uniq, count -> d_count (device); tmp arrays are the caller's
static int unique_labels(ancuts_handle* h, const int* lab, long long n, int* sorted, int* flag, int* incl, int* uniq,
                         long long* d_count, void* cub_tmp, size_t cub_bytes, cudaStream_t st) {
    const int tb = 256; const unsigned g = (unsigned)((n + tb - 1) / tb);
    size_t tmp = cub_bytes;
    h->launches_total += 2;
    ANCUTS_CUDA(cub::DeviceRadixSort::SortKeys(cub_tmp, tmp, lab, sorted, (int)n, 0, 32, st));
    PLAUNCH(k_flag_new<<<g, tb, 0, st>>>(sorted, n, flag));
    tmp = cub_bytes;
    ANCUTS_CUDA(cub::DeviceScan::InclusiveSum(cub_tmp, tmp, flag, incl, (int)n, st));
    PLAUNCH(k_scatter_unique<<<g, tb, 0, st>>>(sorted, flag, incl, n, uniq, d_count));
    return ANCUTS_OK;
}
// NB radix sort orders int32 keys as signed values (CUB handles the sign bit): ascending like np.unique.

// =====================================================================================================================
// N2: merge
// =====================================================================================================================
struct MergeState {
    long long P;                  // all points of all chunks
    int G;                        // distinct non-background labels in the map
    int n2max;                    // largest number of instances in one chunk
    const double* pts;            // [P][3]
    int* rank;                    // [P] current instance rank (-1 = background); rewritten by the association
    int* local;                   // [P] index of the point's ORIGINAL label in its chunk's sorted instance list, or -1
    unsigned char* alive;         // [P]
    // per step, side 1 = cropped merge, side 2 = new chunk
    int* cnt1;                    // [G] cropped points per instance
    double* box1;                 // [G][6] min xyz, max xyz
    int* cnt2;                    // [n2max]
    int* inter;                   // [G][n2max] points of instance 2 inside the box of instance 1
    double* iou;                  // [G][n2max]
    int* match;                   // [n2max] rank of the instance the new instance is united with, or -1
    // distinct scalar values per instance (the reference's np.union1d / intersect1d of FLATTENED coordinates, :457): one
    // open-addressing set per instance that takes part in a candidate pair, all sets of a step in one table
    unsigned long long* sv;       // set slots (f64_key of a coordinate, SET_EMPTY = free)
    long long* ustart;            // [G + n2max] first slot of the instance's set (side 2: G + local)
    int* ucap;                    // [G + n2max] slots of the set (a power of two, >= 4/3 of the values it can receive), 0 = none
    int* usz;                     // [G + n2max] distinct values in the set
    int* need;                    // [G + n2max] the instance is in a pair with a non-empty box intersection
    int* pairs;                   // [G][n2max] candidate pairs (r * n2 + j) of the step, in no particular order
    long long* ctr;               // [4]: [1] slots in use in this step, [2] candidate pairs
    int* htab; long long hmask;   // open-addressing table of point indices (first occurrence wins)
    const int* chunk_rank0;       // [B+1] offsets into chunk_ranks
    const int* chunk_ranks;       // sorted original instance ranks of every chunk
};

__global__ void k_merge_local(MergeState m, long long a, long long b, int c) {
    long long p = a + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= b) return;
    int r = m.rank[p];
    int o = m.chunk_rank0[c], cnt = m.chunk_rank0[c + 1] - o;
    m.local[p] = (r < 0) ? -1 : find_sorted(m.chunk_ranks + o, cnt, r);
}

__global__ void k_merge_reset(MergeState m, int n2) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < m.G) {
        m.cnt1[i] = 0;
        for (int k = 0; k < 3; ++k) { m.box1[i * 6 + k] = 1e300; m.box1[i * 6 + 3 + k] = -1e300; }
    }
    if (i < (long long)m.G + n2) { m.ustart[i] = 0; m.ucap[i] = 0; m.usz[i] = 0; m.need[i] = 0; }
    if (i < n2) { m.cnt2[i] = 0; m.match[i] = -1; }
    if (i < (long long)m.G * n2) { m.inter[i] = 0; m.iou[i] = 0.0; }
    if (i == 0) { m.ctr[0] = 0; m.ctr[1] = 0; m.ctr[2] = 0; }
}

constexpr unsigned long long SET_EMPTY = 0xFFFFFFFFFFFFFFFFull;      // f64_key never yields it (a NaN pattern)
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
    return x;
}
__device__ __forceinline__ bool in_crop(const MergeState& m, long long p, double lx, double ly, double lz, double hx, double hy,
                                        double hz) {
    const double x = m.pts[p * 3], y = m.pts[p * 3 + 1], z = m.pts[p * 3 + 2];
    return x >= lx && x <= hx && y >= ly && y <= hy && z >= lz && z <= hz;
}
// side 1: alive points of the earlier chunks inside the crop box (inclusive, :405-417) -> instance sizes and boxes (:447-448)
__global__ void k_merge_crop(MergeState m, long long p1, double lx, double ly, double lz, double hx, double hy, double hz) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= p1 || !m.alive[p]) return;
    int r = m.rank[p];
    if (r < 0) return;                                          // black = ground / background (:428)
    if (!in_crop(m, p, lx, ly, lz, hx, hy, hz)) return;
    const double x = m.pts[p * 3], y = m.pts[p * 3 + 1], z = m.pts[p * 3 + 2];
    atomicAdd(&m.cnt1[r], 1);
    atomic_min_f64(&m.box1[r * 6 + 0], x); atomic_min_f64(&m.box1[r * 6 + 1], y); atomic_min_f64(&m.box1[r * 6 + 2], z);
    atomic_max_f64(&m.box1[r * 6 + 3], x); atomic_max_f64(&m.box1[r * 6 + 4], y); atomic_max_f64(&m.box1[r * 6 + 5], z);
}
// side 2: the new chunk's instances (all of their points, :434-442)
__global__ void k_merge_cnt2(MergeState m, long long a, long long b) {
    long long p = a + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= b) return;
    int j = m.local[p];
    if (j >= 0) atomicAdd(&m.cnt2[j], 1);
}
// one block: slots of the sets of the instances that are in a candidate pair (3 scalar values per point, load <= 3/4)
__global__ void __launch_bounds__(1024)
k_merge_regions(MergeState m, int n2) {
    __shared__ long long wsum[32];
    __shared__ long long carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = m.G + n2;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < total; base += 1024) {
        const int id = base + tid;
        int cap = 0;
        if (id < total && m.need[id]) {
            const int cnt = (id < m.G) ? m.cnt1[id] : m.cnt2[id - m.G];
            cap = 4;
            while (cap < 4 * cnt) cap <<= 1;
        }
        long long v = cap;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { long long t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
        if (lane == 31) wsum[warp] = v;
        __syncthreads();
        if (warp == 0) {
            long long w = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { long long t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
            wsum[lane] = w;
        }
        __syncthreads();
        const long long before = carry + (warp ? wsum[warp - 1] : 0) + v - cap;
        if (id < total) { m.ustart[id] = before; m.ucap[id] = cap; }
        __syncthreads();
        if (tid == 1023) carry = before + cap;
        __syncthreads();
    }
    if (tid == 0) m.ctr[1] = carry;
}
__global__ void k_merge_set_clear(MergeState m, long long bound) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < bound && i < m.ctr[1]) m.sv[i] = SET_EMPTY;
}
__device__ __forceinline__ void set_insert(const MergeState& m, int id, unsigned long long v) {
    const unsigned mask = (unsigned)m.ucap[id] - 1u;
    unsigned long long* T = m.sv + m.ustart[id];
    unsigned s = (unsigned)mix64(v) & mask;
    while (true) {
        const unsigned long long old = atomicCAS(&T[s], SET_EMPTY, v);
        if (old == SET_EMPTY) { atomicAdd(&m.usz[id], 1); return; }
        if (old == v) return;
        s = (s + 1u) & mask;
    }
}
__device__ __forceinline__ bool set_has(const MergeState& m, int id, unsigned long long v) {
    const unsigned mask = (unsigned)m.ucap[id] - 1u;
    const unsigned long long* T = m.sv + m.ustart[id];
    unsigned s = (unsigned)mix64(v) & mask;
    while (true) {
        const unsigned long long cur = T[s];
        if (cur == v) return true;
        if (cur == SET_EMPTY) return false;
        s = (s + 1u) & mask;
    }
}
// the three coordinates of every point of a needed instance go into the instance's set; side 1 = [0, p1) cropped, side 2
// = the chunk [a, b)
__global__ void k_merge_set_fill(MergeState m, long long p1, long long b, double lx, double ly, double lz, double hx, double hy,
                                 double hz) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= b) return;
    int id;
    if (p < p1) {
        if (!m.alive[p]) return;
        id = m.rank[p];
        if (id < 0 || !m.need[id] || !in_crop(m, p, lx, ly, lz, hx, hy, hz)) return;
    } else {
        const int j = m.local[p];
        if (j < 0 || !m.need[m.G + j]) return;
        id = m.G + j;
    }
    set_insert(m, id, f64_key(m.pts[p * 3]));
    set_insert(m, id, f64_key(m.pts[p * 3 + 1]));
    set_insert(m, id, f64_key(m.pts[p * 3 + 2]));
}
// points of every new instance inside the box of every cropped instance (:452-455)
__global__ void __launch_bounds__(256)
k_merge_inter(MergeState m, long long a, long long b, int n2) {
    // few of the map's G instances have points in the crop: the block compacts the boxes of those into shared memory
    // (256 candidates per round) and every point tests only them
    __shared__ double sbox[256][6];
    __shared__ int sr[256];
    __shared__ int sn;
    const long long p = a + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int j = (p < b) ? m.local[p] : -1;
    double x = 0.0, y = 0.0, z = 0.0;
    if (j >= 0) { x = m.pts[p * 3]; y = m.pts[p * 3 + 1]; z = m.pts[p * 3 + 2]; }
    for (int base = 0; base < m.G; base += 256) {
        if (threadIdx.x == 0) sn = 0;
        __syncthreads();
        const int rc = base + threadIdx.x;
        if (rc < m.G && m.cnt1[rc] > 0) {
            const int k = atomicAdd(&sn, 1);
            sr[k] = rc;
#pragma unroll
            for (int q = 0; q < 6; ++q) sbox[k][q] = m.box1[(size_t)rc * 6 + q];
        }
        __syncthreads();
        if (j >= 0) {
            const int cnt = sn;
            for (int k = 0; k < cnt; ++k) {
                const double* bx = sbox[k];
                if (x >= bx[0] && y >= bx[1] && z >= bx[2] && x <= bx[3] && y <= bx[4] && z <= bx[5]) {
                    const int r = sr[k];
                    if (atomicAdd(&m.inter[(size_t)r * n2 + j], 1) == 0) {      // first point of the pair: it becomes a candidate
                        m.pairs[atomicAdd((unsigned long long*)&m.ctr[2], 1ull)] = r * n2 + j;
                        m.need[r] = 1; m.need[m.G + j] = 1;
                    }
                }
            }
        }
        __syncthreads();
    }
}
// one block per candidate pair (cropped instance r, new instance j): |U1 n U2| by looking the values of the smaller set up
// in the larger one, four look-ups in flight per thread; union = |U1| + |U2| - common; iou = float(inter) / float(union)
// (:457-458)
constexpr int IOU_BLOCKS = 296;
__global__ void __launch_bounds__(256)
k_merge_iou(MergeState m, int n2, double min_iou) {
    __shared__ int red[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int npairs = (int)m.ctr[2];
    for (int pi = blockIdx.x; pi < npairs; pi += gridDim.x) {
        const int w = m.pairs[pi];
        const int r = w / n2, j = w % n2;
        int ia = r, ib = m.G + j;
        if (m.usz[ia] > m.usz[ib]) { int t = ia; ia = ib; ib = t; }
        const unsigned long long* A = m.sv + m.ustart[ia];
        const int capa = m.ucap[ia];
        int common = 0;
        for (int i = tid; i < capa; i += 4 * 256) {
            unsigned long long v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = (i + 256 * u < capa) ? A[i + 256 * u] : SET_EMPTY;
#pragma unroll
            for (int u = 0; u < 4; ++u) if (v[u] != SET_EMPTY && set_has(m, ib, v[u])) ++common;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) common += __shfl_xor_sync(0xffffffffu, common, o);
        if (lane == 0) red[warp] = common;
        __syncthreads();
        if (tid == 0) {
            int c = 0;
            for (int k = 0; k < 8; ++k) c += red[k];
            const int uni = m.usz[ia] + m.usz[ib] - c;
            const double v = (double)m.inter[w] / (double)uni;
            m.iou[w] = (v > min_iou) ? v : 0.0;                          // :459
        }
        __syncthreads();
    }
}
// every new instance keeps the cropped instance with the largest iou; pairs come in (id1, id2) ascending order and a
// later pair replaces an earlier one only if strictly larger (:465-477)
__global__ void k_merge_resolve(MergeState m, int n2) {
    // one warp per new instance: largest iou, the smallest rank among equals (what "strictly larger replaces" leaves)
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (j >= n2) return;
    double best = 0.0; int br = -1;
    for (int r = lane; r < m.G; r += 32) {
        const double v = m.iou[(size_t)r * n2 + j];
        if (v > 0.0 && (br < 0 || v > best)) { best = v; br = r; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int orr = __shfl_xor_sync(0xffffffffu, br, o);
        if (orr >= 0 && (br < 0 || ob > best || (ob == best && orr < br))) { best = ob; br = orr; }
    }
    if (lane == 0) m.match[j] = br;
}
__global__ void k_merge_apply(MergeState m, long long a, long long b) {
    long long p = a + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= b) return;
    int j = m.local[p];
    if (j >= 0 && m.match[j] >= 0) m.rank[p] = m.match[j];              // recolour (:479-481)
}

// ---- duplicate removal: exact-coordinate duplicates dropped, first occurrence (smallest index) kept (:488-489) ----
__device__ __forceinline__ unsigned long long hash3(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long h = a * 0x9E3779B97F4A7C15ull;
    h ^= (h >> 29); h += b * 0xBF58476D1CE4E5B9ull; h ^= (h >> 32);
    h += c * 0x94D049BB133111EBull; h ^= (h >> 29); h *= 0xD6E8FEB86659FD93ull; h ^= (h >> 32);
    return h;
}
__device__ __forceinline__ bool same_point(const double* pts, long long p, long long q) {
    return pts[p * 3] == pts[q * 3] && pts[p * 3 + 1] == pts[q * 3 + 1] && pts[p * 3 + 2] == pts[q * 3 + 2];
}
__global__ void k_merge_hash_insert(MergeState m, long long a, long long b) {
    long long p = a + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= b) return;
    unsigned long long s = hash3(f64_key(m.pts[p * 3]), f64_key(m.pts[p * 3 + 1]), f64_key(m.pts[p * 3 + 2])) & (unsigned long long)m.hmask;
    while (true) {
        int cur = m.htab[s];
        if (cur < 0) {
            int old = atomicCAS(&m.htab[s], -1, (int)p);
            if (old < 0) return;
            cur = old;
        }
        if (same_point(m.pts, p, cur)) { atomicMin(&m.htab[s], (int)p); return; }   // owner changes only among equal points
        s = (s + 1) & (unsigned long long)m.hmask;
    }
}
__global__ void k_merge_hash_alive(MergeState m, long long a, long long b) {
    long long p = a + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= b) return;
    unsigned long long s = hash3(f64_key(m.pts[p * 3]), f64_key(m.pts[p * 3 + 1]), f64_key(m.pts[p * 3 + 2])) & (unsigned long long)m.hmask;
    while (true) {
        int cur = m.htab[s];
        if (cur < 0) { m.alive[p] = 1; return; }                        // cannot happen after the insert pass
        if (same_point(m.pts, p, cur)) { m.alive[p] = (cur == (int)p) ? 1 : 0; return; }
        s = (s + 1) & (unsigned long long)m.hmask;
    }
}
__global__ void k_fill_i32(int* p, long long n, int v) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_fill_u8(unsigned char* p, long long n, unsigned char v) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_chunk_keys(const int* __restrict__ rank, const int* __restrict__ chunk_of, long long n, int G,
                             unsigned long long* __restrict__ key) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) key[i] = (unsigned long long)chunk_of[i] * (unsigned long long)(G + 1) + (unsigned long long)(rank[i] + 1);
}
__global__ void k_chunk_of(const long long* __restrict__ off, int B, long long n, int* __restrict__ chunk_of) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = 0, hi = B - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (off[mid] <= i) lo = mid; else hi = mid - 1; }
    chunk_of[i] = lo;
}
// distinct (chunk, rank) keys, sorted: per-chunk instance lists
__global__ void k_chunk_lists(const unsigned long long* __restrict__ skey, const int* __restrict__ flag, const int* __restrict__ incl,
                              long long n, int G, int B, int* __restrict__ chunk_ranks, int* __restrict__ chunk_cnt) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    const unsigned long long k = skey[i];
    const int c = (int)(k / (unsigned long long)(G + 1));
    const int r = (int)(k % (unsigned long long)(G + 1)) - 1;
    // background entries (r = -1) are dropped: position = distinct keys before this one minus background keys of chunks <= c
    if (r < 0) return;
    chunk_ranks[incl[i] - 1] = r;                     // provisional position (compacted by k_chunk_lists_fix)
    atomicAdd(&chunk_cnt[c], 1);
    (void)B;
}
__global__ void k_flag_new_u64(const unsigned long long* __restrict__ s, long long n, int* __restrict__ flag) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i == 0 || s[i] != s[i - 1]) ? 1 : 0;
}
__global__ void k_flag_fg_u64(const unsigned long long* __restrict__ s, long long n, int G, int* __restrict__ flag) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) flag[i] = ((i == 0 || s[i] != s[i - 1]) && (s[i] % (unsigned long long)(G + 1)) != 0ull) ? 1 : 0;
}
__global__ void k_rank_to_label(const int* __restrict__ rank, const int* __restrict__ uniq, long long n, int* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = rank[i] < 0 ? 0 : uniq[rank[i]];
}
struct IsAlive {
    const unsigned char* alive;
    __device__ __forceinline__ bool operator()(const long long& i) const { return alive[i] != 0; }
};

}  // namespace ancuts

extern "C" {

// Merge of the chunk labelings of one map (merge_chunks_unite_instances2, point_cloud_utils.py:387-491).
int ancuts_merge_chunks(ancuts_handle* h, int num_chunks, const int64_t* h_chunk_off, const double* d_points,
                        const int32_t* d_labels, const double* h_centers, double crop_half_side, double min_iou,
                        int32_t* d_out_labels, int64_t* d_out_index, int64_t* h_num_kept, void* stream) {
    if (!h || num_chunks <= 0 || !h_chunk_off || !d_points || !d_labels || !h_centers || !d_out_labels || h_chunk_off[0] != 0) {
        set_error("bad argument to ancuts_merge_chunks");
        return ANCUTS_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    ANCUTS_CUDA(cudaSetDevice(h->device));
    const int B = num_chunks;
    const long long P = h_chunk_off[B];
    if (P <= 0 || P > 0x7ffffff0LL / 3) { set_error("ancuts_merge_chunks: %lld points (supported: 1 .. 7e8)", P); return ANCUTS_EINVAL; }
    long long maxc = 0;
    for (int c = 0; c < B; ++c) {
        if (h_chunk_off[c + 1] <= h_chunk_off[c]) { set_error("chunk %d is empty", c); return ANCUTS_EINVAL; }
        maxc = std::max<long long>(maxc, h_chunk_off[c + 1] - h_chunk_off[c]);
    }
    const int tb = 256;
    auto grid = [&](long long n) { return (unsigned)std::max<long long>(1, (n + tb - 1) / tb); };

    // ---- phase A: dense ranks of the label values, per-chunk instance lists (sizes come back to the host once) ----
    size_t cubA = 0, cubB = 0, cubC = 0, cubD = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, cubA, (int*)nullptr, (int*)nullptr, (int)P, 0, 32);
    cub::DeviceScan::InclusiveSum(nullptr, cubB, (int*)nullptr, (int*)nullptr, (int)P);
    cub::DeviceRadixSort::SortKeys(nullptr, cubC, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)P, 0, 64);
    size_t cubE = 0;
    cub::DeviceSelect::If(nullptr, cubE, cub::CountingInputIterator<long long>(0), (long long*)nullptr, (long long*)nullptr,
                          (long long)P, IsAlive{nullptr});
    const size_t cub_bytes = std::max({cubA, cubB, cubC, cubD, cubE}) + 256;
    long long hsize = 1;
    while (hsize < 2 * P) hsize <<= 1;

    PArena ar{nullptr};
    auto layoutA = [&](PArena& a, int*& sorted, int*& flag, int*& incl, int*& uniq, long long*& dctr, int*& rank, int*& local,
                       unsigned char*& alive, int*& chunk_of, unsigned long long*& ckey, unsigned long long*& ckey2,
                       int*& chunk_ranks, int*& chunk_cnt, long long*& d_off, int*& htab, void*& cub_tmp) {
        sorted = a.take<int>(P); flag = a.take<int>(P); incl = a.take<int>(P); uniq = a.take<int>(P);
        dctr = a.take<long long>(8); rank = a.take<int>(P); local = a.take<int>(P); alive = a.take<unsigned char>(P);
        chunk_of = a.take<int>(P); ckey = a.take<unsigned long long>(P); ckey2 = a.take<unsigned long long>(P);
        chunk_ranks = a.take<int>(P); chunk_cnt = a.take<int>(B + 1); d_off = a.take<long long>(B + 1);
        htab = a.take<int>(hsize); cub_tmp = a.take<char>(cub_bytes);
    };
    int *sorted, *flag, *incl, *uniq, *rank, *local, *chunk_of, *chunk_ranks, *chunk_cnt, *htab;
    long long *dctr, *d_off; unsigned char* alive; unsigned long long *ckey, *ckey2; void* cub_tmp;
    layoutA(ar, sorted, flag, incl, uniq, dctr, rank, local, alive, chunk_of, ckey, ckey2, chunk_ranks, chunk_cnt, d_off, htab, cub_tmp);
    // phase B buffers sized after G / n2max are known: reserve the set table now (3 values per point, sets at most 8/3 of
    // their values: 8 slots per point) and the instance tables later
    const size_t fixed_bytes = ar.off;
    const long long set_slots = 8 * P + 64;
    size_t keys_bytes = pa_align((size_t)set_slots * 8) + 4096;
    int rc = post_ws(h, fixed_bytes + keys_bytes + (64u << 20));
    if (rc) return rc;
    ar = PArena{h->post_ws};
    layoutA(ar, sorted, flag, incl, uniq, dctr, rank, local, alive, chunk_of, ckey, ckey2, chunk_ranks, chunk_cnt, d_off, htab, cub_tmp);

    ANCUTS_CUDA(cudaMemcpyAsync(d_off, h_chunk_off, (B + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    rc = unique_labels(h, d_labels, P, sorted, flag, incl, uniq, dctr, cub_tmp, cub_bytes, st);
    if (rc) return rc;
    // uniq[] may start with the background 0: ranks are taken among the NON-background values, so shift when it is there
    ANCUTS_CUDA(cudaMemcpyAsync(h->h_post, dctr, sizeof(long long), cudaMemcpyDeviceToHost, st));
    int first_label = 0;
    ANCUTS_CUDA(cudaMemcpyAsync(&first_label, uniq, sizeof(int), cudaMemcpyDeviceToHost, st));
    ANCUTS_CUDA(cudaStreamSynchronize(st));
    long long nu = h->h_post[0];
    // negative labels would sort before 0; the reference has none (colours): reject
    if (first_label < 0) { set_error("ancuts_merge_chunks: negative labels are not supported (0 = background)"); return ANCUTS_EINVAL; }
    const int has_bg = (first_label == 0) ? 1 : 0;
    const int G = (int)(nu - has_bg);
    const int* uniq_fg = uniq + has_bg;
    if (G == 0) {                                       // nothing but background: only the duplicate removal is left
        // fall through with G = 0 tables of size 1
    }
    {
        // rank among the non-background labels
        long long* d_nufg = dctr + 1;
        long long nufg = G;
        ANCUTS_CUDA(cudaMemcpyAsync(d_nufg, &nufg, sizeof(long long), cudaMemcpyHostToDevice, st));
        ANCUTS_CUDA(cudaStreamSynchronize(st));          // nufg is a stack variable
        PLAUNCH(k_rank_of<<<grid(P), tb, 0, st>>>(d_labels, P, uniq_fg, d_nufg, 1, rank));
    }
    PLAUNCH(k_chunk_of<<<grid(P), tb, 0, st>>>(d_off, B, P, chunk_of));
    PLAUNCH(k_chunk_keys<<<grid(P), tb, 0, st>>>(rank, chunk_of, P, G, ckey));
    {
        size_t tmp = cub_bytes;
        h->launches_total += 2;
        ANCUTS_CUDA(cub::DeviceRadixSort::SortKeys(cub_tmp, tmp, ckey, ckey2, (int)P, 0, 64, st));
        PLAUNCH(k_flag_fg_u64<<<grid(P), tb, 0, st>>>(ckey2, P, G, flag));          // distinct (chunk, instance) keys, background dropped
        tmp = cub_bytes;
        ANCUTS_CUDA(cub::DeviceScan::InclusiveSum(cub_tmp, tmp, flag, incl, (int)P, st));
        ANCUTS_CUDA(cudaMemsetAsync(chunk_cnt, 0, (B + 1) * sizeof(int), st));
        PLAUNCH(k_chunk_lists<<<grid(P), tb, 0, st>>>(ckey2, flag, incl, P, G, B, chunk_ranks, chunk_cnt));
    }
    std::vector<int> h_cnt(B + 1, 0), h_rank0(B + 1, 0);
    ANCUTS_CUDA(cudaMemcpyAsync(h_cnt.data(), chunk_cnt, B * sizeof(int), cudaMemcpyDeviceToHost, st));
    ANCUTS_CUDA(cudaStreamSynchronize(st));
    int n2max = 1;
    for (int c = 0; c < B; ++c) { h_rank0[c + 1] = h_rank0[c] + h_cnt[c]; n2max = std::max(n2max, h_cnt[c]); }
    // a label that shows up in two chunks would break the per-chunk instance lists' meaning: the lists stay correct
    // (they are per chunk), nothing to check.
    int* d_rank0 = chunk_cnt;                           // reuse: offsets
    ANCUTS_CUDA(cudaMemcpyAsync(d_rank0, h_rank0.data(), (B + 1) * sizeof(int), cudaMemcpyHostToDevice, st));

    // ---- phase B buffers ----
    const int Gs = std::max(G, 1);
    const size_t table = (size_t)Gs * n2max;
    if (table > (1ull << 30)) { set_error("ancuts_merge_chunks: instance table %d x %d too large", G, n2max); return ANCUTS_ENOMEM; }
    PArena br{nullptr};
    br.off = fixed_bytes;
    auto layoutB = [&](PArena& a, MergeState& m) {
        m.cnt1 = a.take<int>(Gs); m.box1 = a.take<double>((size_t)Gs * 6); m.cnt2 = a.take<int>(n2max);
        m.inter = a.take<int>(table); m.iou = a.take<double>(table); m.match = a.take<int>(n2max);
        m.pairs = a.take<int>(table);
        m.sv = a.take<unsigned long long>((size_t)set_slots);
        m.ustart = a.take<long long>((size_t)Gs + n2max); m.ucap = a.take<int>((size_t)Gs + n2max);
        m.usz = a.take<int>((size_t)Gs + n2max); m.need = a.take<int>((size_t)Gs + n2max);
    };
    MergeState m;
    memset(&m, 0, sizeof(m));
    layoutB(br, m);
    const size_t need = br.off + 256;
    if (need > h->post_ws_bytes) {
        // grow: phase A results must survive -> allocate a new block and copy the fixed part over
        char* old = h->post_ws; size_t oldb = h->post_ws_bytes;
        h->post_ws = nullptr; h->post_ws_bytes = 0;
        rc = post_ws(h, need);
        if (rc) { h->post_ws = old; h->post_ws_bytes = oldb; return rc; }
        ANCUTS_CUDA(cudaMemcpyAsync(h->post_ws, old, fixed_bytes, cudaMemcpyDeviceToDevice, st));
        ANCUTS_CUDA(cudaStreamSynchronize(st));
        cudaFree(old);
        ar = PArena{h->post_ws};
        layoutA(ar, sorted, flag, incl, uniq, dctr, rank, local, alive, chunk_of, ckey, ckey2, chunk_ranks, chunk_cnt, d_off, htab, cub_tmp);
        d_rank0 = chunk_cnt;
        uniq_fg = uniq + has_bg;
    }
    br = PArena{h->post_ws};
    br.off = fixed_bytes;
    layoutB(br, m);
    m.P = P; m.G = G; m.n2max = n2max; m.pts = d_points; m.rank = rank; m.local = local; m.alive = alive;
    m.ctr = dctr + 2; m.htab = htab; m.hmask = hsize - 1;
    m.chunk_rank0 = d_rank0; m.chunk_ranks = chunk_ranks;

    PLAUNCH(k_fill_i32<<<grid(hsize), tb, 0, st>>>(htab, hsize, -1));
    PLAUNCH(k_fill_u8<<<grid(P), tb, 0, st>>>(alive, P, 1));
    for (int c = 0; c < B; ++c) {
        const long long a = h_chunk_off[c], b = h_chunk_off[c + 1];
        PLAUNCH(k_merge_local<<<grid(b - a), tb, 0, st>>>(m, a, b, c));
    }
    // ---- phase C: chunk after chunk (:393-489) ----
    for (int c = 1; c < B; ++c) {
        const long long a = h_chunk_off[c], b = h_chunk_off[c + 1];
        const int n2 = h_cnt[c];
        if (c == 1) {                                    // the first chunk enters the merge as it is (:390-391)
            PLAUNCH(k_merge_hash_insert<<<grid(h_chunk_off[1]), tb, 0, st>>>(m, 0, h_chunk_off[1]));
        }
        if (n2 > 0 && G > 0) {
            const double* ctr = h_centers + 3 * (size_t)c;
            const double lx = ctr[0] - crop_half_side, ly = ctr[1] - crop_half_side, lz = ctr[2] - crop_half_side;
            const double hx = ctr[0] + crop_half_side, hy = ctr[1] + crop_half_side, hz = ctr[2] + crop_half_side;
            const long long tbl = std::max<long long>((long long)G * n2, (long long)G + n2);
            PLAUNCH(k_merge_reset<<<grid(tbl), tb, 0, st>>>(m, n2));
            // no read-back inside the loop: sizes of the step stay on the device
            PLAUNCH(k_merge_crop<<<grid(a), tb, 0, st>>>(m, a, lx, ly, lz, hx, hy, hz));
            PLAUNCH(k_merge_cnt2<<<grid(b - a), tb, 0, st>>>(m, a, b));
            PLAUNCH(k_merge_inter<<<grid(b - a), tb, 0, st>>>(m, a, b, n2));
            PLAUNCH(k_merge_regions<<<1, 1024, 0, st>>>(m, n2));
            const long long bound = std::min<long long>(set_slots, 8 * b + 64);
            PLAUNCH(k_merge_set_clear<<<grid(bound), tb, 0, st>>>(m, bound));
            PLAUNCH(k_merge_set_fill<<<grid(b), tb, 0, st>>>(m, a, b, lx, ly, lz, hx, hy, hz));
            PLAUNCH(k_merge_iou<<<IOU_BLOCKS, 256, 0, st>>>(m, n2, min_iou));
            PLAUNCH(k_merge_resolve<<<grid((long long)n2 * 32), tb, 0, st>>>(m, n2));
            PLAUNCH(k_merge_apply<<<grid(b - a), tb, 0, st>>>(m, a, b));
        }
        PLAUNCH(k_merge_hash_insert<<<grid(b - a), tb, 0, st>>>(m, a, b));
        // remove_duplicated_points (:489): a new point never displaces an earlier owner (larger index), so only the new chunk
        // (and, at the first step, chunk 0 with its own duplicates) needs a look-up
        const long long a0 = (c == 1) ? 0 : a;
        PLAUNCH(k_merge_hash_alive<<<grid(b - a0), tb, 0, st>>>(m, a0, b));
    }
    ANCUTS_CUDA(cudaGetLastError());
    // ---- output: labels of every slot, indices of the surviving points in order ----
    PLAUNCH(k_rank_to_label<<<grid(P), tb, 0, st>>>(rank, uniq_fg, P, d_out_labels));
    long long kept = P;
    if (d_out_index) {
        size_t tmp = cub_bytes;
        h->launches_total++;
        ANCUTS_CUDA(cub::DeviceSelect::If(cub_tmp, tmp, cub::CountingInputIterator<long long>(0), (long long*)d_out_index,
                                          dctr + 4, P, IsAlive{alive}, st));
        ANCUTS_CUDA(cudaMemcpyAsync(h->h_post, dctr + 4, sizeof(long long), cudaMemcpyDeviceToHost, st));
        ANCUTS_CUDA(cudaStreamSynchronize(st));
        kept = h->h_post[0];
    } else {
        ANCUTS_CUDA(cudaStreamSynchronize(st));
    }
    if (h_num_kept) *h_num_kept = kept;
    return ANCUTS_OK;
}

}  // extern "C"

// =====================================================================================================================
// glue of run_pipeline.py:216-218 for integer labels: per-chunk segment ids -> ids unique across the map
// =====================================================================================================================
namespace ancuts {
__global__ void k_first_min(const int* __restrict__ seg, const int* __restrict__ chunk_of, const long long* __restrict__ off,
                            long long n, int* __restrict__ first) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= n) return;
    const long long o = off[chunk_of[p]];
    atomicMin(&first[o + seg[p]], (int)(p - o));
}
__global__ void k_mark_first(const int* __restrict__ seg, const int* __restrict__ chunk_of, const long long* __restrict__ off,
                             long long n, const int* __restrict__ first, int* __restrict__ flag) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= n) return;
    const long long o = off[chunk_of[p]];
    flag[p] = (first[o + seg[p]] == (int)(p - o)) ? 1 : 0;
}
__global__ void k_emit_global(const int* __restrict__ seg, const int* __restrict__ chunk_of, const long long* __restrict__ off,
                              long long n, const int* __restrict__ first, const int* __restrict__ incl, int shift,
                              int* __restrict__ out) {
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int c = chunk_of[p];
    const long long o = off[c];
    const int base = (o > 0) ? incl[o - 1] : 0;
    const int rank = incl[o + first[o + seg[p]]] - 1 - base;            // segments numbered by first occurrence in point order
    out[p] = ((c + 1) << shift) + rank + 1;
}
}  // namespace ancuts

extern "C" {
// Per-chunk segment ids (0 .. n_c - 1 within chunk c, as ancuts_segment_chunks writes them) -> labels unique across the
// map: ((c + 1) << id_shift) + r + 1, r = rank of the segment by FIRST OCCURRENCE in the chunk's point order.  The
// reference gives every segment a random colour (ncuts_utils.py:177-183) and turns colours into integers with np.unique
// (run_pipeline.py:216-218); its greedy merge / matching rules depend on the ORDER of those values, so a deterministic,
// labeling-independent numbering is needed to compare two labelings.  Asynchronous.
int ancuts_map_labels(ancuts_handle* h, int num_chunks, const int64_t* h_chunk_off, const int32_t* d_seg_labels,
                      int id_shift, int32_t* d_out_labels, void* stream) {
    if (!h || num_chunks <= 0 || !h_chunk_off || !d_seg_labels || !d_out_labels || id_shift < 1 || id_shift > 24 ||
        (long long)(num_chunks + 1) >= (1ll << (31 - id_shift))) {
        set_error("bad argument to ancuts_map_labels");
        return ANCUTS_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    ANCUTS_CUDA(cudaSetDevice(h->device));
    const long long P = h_chunk_off[num_chunks];
    if (P <= 0 || P > 0x7ffffff0LL) { set_error("ancuts_map_labels: bad point count"); return ANCUTS_EINVAL; }
    for (int c = 0; c < num_chunks; ++c)
        if (h_chunk_off[c + 1] - h_chunk_off[c] >= (1ll << id_shift)) { set_error("chunk %d too large for id_shift %d", c, id_shift); return ANCUTS_EINVAL; }
    size_t cubb = 0;
    cub::DeviceScan::InclusiveSum(nullptr, cubb, (int*)nullptr, (int*)nullptr, (int)P);
    PArena ar{nullptr};
    auto lay = [&](PArena& a, int*& first, int*& flag, int*& incl, int*& chunk_of, long long*& d_off, void*& tmp) {
        first = a.take<int>(P); flag = a.take<int>(P); incl = a.take<int>(P); chunk_of = a.take<int>(P);
        d_off = a.take<long long>(num_chunks + 1); tmp = a.take<char>(cubb + 256);
    };
    int *first, *flag, *incl, *chunk_of; long long* d_off; void* tmp;
    lay(ar, first, flag, incl, chunk_of, d_off, tmp);
    int rc = post_ws(h, ar.off + 256);
    if (rc) return rc;
    ar = PArena{h->post_ws};
    lay(ar, first, flag, incl, chunk_of, d_off, tmp);
    const int tb = 256; const unsigned g = (unsigned)((P + tb - 1) / tb);
    ANCUTS_CUDA(cudaMemcpyAsync(d_off, h_chunk_off, (num_chunks + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    PLAUNCH(k_chunk_of<<<g, tb, 0, st>>>(d_off, num_chunks, P, chunk_of));
    PLAUNCH(k_fill_i32<<<g, tb, 0, st>>>(first, P, 0x7fffffff));
    PLAUNCH(k_first_min<<<g, tb, 0, st>>>(d_seg_labels, chunk_of, d_off, P, first));
    PLAUNCH(k_mark_first<<<g, tb, 0, st>>>(d_seg_labels, chunk_of, d_off, P, first, flag));
    size_t tb2 = cubb + 256;
    h->launches_total++;
    ANCUTS_CUDA(cub::DeviceScan::InclusiveSum(tmp, tb2, flag, incl, (int)P, st));
    PLAUNCH(k_emit_global<<<g, tb, 0, st>>>(d_seg_labels, chunk_of, d_off, P, first, incl, id_shift, d_out_labels));
    ANCUTS_CUDA(cudaGetLastError());
    ANCUTS_CUDA(cudaStreamSynchronize(st));               // h_chunk_off was copied asynchronously from the caller's memory
    return ANCUTS_OK;
}
}  // extern "C"

// =====================================================================================================================
// remove_semantics + N4: instance metrics
// =====================================================================================================================
namespace ancuts {

__global__ void k_hist_rank(const int* __restrict__ rank, long long n, int* __restrict__ cnt) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n && rank[i] >= 0) atomicAdd(&cnt[rank[i]], 1);
}
// remove_semantics (:253-287): points of a predicted label on GT background
__global__ void k_sem_count(const int* __restrict__ prank, const int* __restrict__ gt, long long n, int* __restrict__ cnt,
                            int* __restrict__ bg) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = prank[i];
    atomicAdd(&cnt[r], 1);
    if (gt[i] == 0) atomicAdd(&bg[r], 1);
}
__global__ void k_sem_apply(const int* __restrict__ pred, const int* __restrict__ prank, long long n, const int* __restrict__ cnt,
                            const int* __restrict__ bg, double thr, int* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = prank[i];
    out[i] = ((double)bg[r] > thr * (double)cnt[r]) ? 0 : pred[i];       // cur_intersect > threshold * len(pred_idcs) (:257)
}
// filter_labels (metrics_class.py:302-309): labels with fewer than min_points points become 0
__global__ void k_filter_small(const int* __restrict__ lab, const int* __restrict__ rk, long long n, const int* __restrict__ cnt,
                               int min_points, int* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (cnt[rk[i]] < min_points) ? 0 : lab[i];
}
// (pred rank, gt rank) key of every point where both are non-background
__global__ void k_pair_keys(const int* __restrict__ pr, const int* __restrict__ gr, long long n, long long stride,
                            unsigned long long* __restrict__ key, int major_gt) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = pr[i], g = gr[i];
    if (p < 0 || g < 0) key[i] = 0xffffffffffffffffull;                  // sorts behind every real pair
    else key[i] = major_gt ? (unsigned long long)g * (unsigned long long)stride + (unsigned long long)p
                           : (unsigned long long)p * (unsigned long long)stride + (unsigned long long)g;
}
__global__ void k_run_heads(const unsigned long long* __restrict__ s, long long n, int* __restrict__ flag) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (s[i] != 0xffffffffffffffffull && (i == 0 || s[i] != s[i - 1])) ? 1 : 0;
}
// run-length encode the sorted keys: pair list (key, count)
__global__ void k_run_emit(const unsigned long long* __restrict__ s, const int* __restrict__ flag, const int* __restrict__ incl,
                           long long n, unsigned long long* __restrict__ pkey, int* __restrict__ pstart, long long* __restrict__ npairs,
                           long long* __restrict__ nvalid) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flag[i]) { pkey[incl[i] - 1] = s[i]; pstart[incl[i] - 1] = (int)i; }
    if (i == n - 1) *npairs = incl[i];
    if (s[i] != 0xffffffffffffffffull && (i == n - 1 || s[i + 1] == 0xffffffffffffffffull)) *nvalid = i + 1;
}

struct MetricsState {
    int Up, Ug, Ua;                 // distinct non-background labels: filtered predictions, GT, filtered all-labels
    const int* psz; const int* gsz; const int* asz;      // points per rank
    int has_zero_pred, has_zero_gt; // 0 present among the values (:321-323)
    int up_total;                   // np.unique(pred).shape[0]
    int ug_total;
    const unsigned long long* pkey; const int* pstart; long long np_; long long nvalid;   // (pred, gt) pairs, pred-major
    const unsigned long long* akey; const int* astart; long long na; long long navalid;   // (gt, all) pairs, gt-major
    double* iou;                    // per (pred, gt) pair
    int* pair_first;                // [Up + 1] first pair of every prediction
    unsigned char* used;            // [12][Ug]
    double* out;                    // [16]
    int min_points;
};
__global__ void k_pair_iou(MetricsState m) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= m.np_) return;
    const unsigned long long k = m.pkey[i];
    const int p = (int)(k / (unsigned long long)(m.Ug + 1)), g = (int)(k % (unsigned long long)(m.Ug + 1));
    const long long end = (i + 1 < m.np_) ? m.pstart[i + 1] : m.nvalid;
    const int c = (int)(end - m.pstart[i]);
    m.iou[i] = (double)c / (double)(m.psz[p] + m.gsz[g] - c);           // Metrics.iou (:296-300)
    if (i == 0 || (int)(m.pkey[i - 1] / (unsigned long long)(m.Ug + 1)) != p) m.pair_first[p] = (int)i;
}
// thread t < 11: average_precision at OVERLAPS[t] (:181-235); thread 11: P / R / F1 at 0.5 (:61-117,315-340).
// Predictions and GT instances are visited in np.unique order; a prediction takes the first unused GT with
// iou >= thr.  Pairs absent from the table have iou 0 < thr.
__global__ void k_greedy(MetricsState m) {
    const int t = threadIdx.x;
    if (t >= 12) return;
    const double overlaps[11] = {0.25, 0.5, 0.55, 0.6, 0.65, 0.7, 0.75, 0.8, 0.85, 0.9, 0.95};
    const double thr = (t < 11) ? overlaps[t] : 0.5;
    unsigned char* used = m.used + (size_t)t * (m.Ug + 1);
    for (int g = 0; g < m.Ug; ++g) used[g] = 0;
    int tp = 0, fp = 0, fn = m.Ug;
    double area = 0.0, pprev = 1.0, rprev = 0.0;         // np.trapz(precision, recall) with the sentinels (1, 0)
    for (int p = 0; p < m.Up; ++p) {
        bool hit = false;
        const int b = m.pair_first[p], e = m.pair_first[p + 1];
        for (int i = b; i < e; ++i) {                    // this prediction's GT candidates, ascending
            const int g = (int)(m.pkey[i] % (unsigned long long)(m.Ug + 1));
            if (m.iou[i] >= thr && !used[g]) { used[g] = 1; hit = true; break; }
        }
        if (hit) { tp += 1; fn -= 1; } else fp += 1;
        const double prec = (double)tp / (double)(tp + fp), rec = (double)tp / (double)(tp + fn);
        area += (rec - rprev) * (prec + pprev) / 2.0;
        pprev = prec; rprev = rec;
    }
    if (t < 11) m.out[t] = area;
    else {
        const int n_gt = m.has_zero_gt ? (m.ug_total - 1) : 0;           // :321-322
        const int n_pred = m.up_total - 1;                                 // :323
        m.out[11] = (double)tp; m.out[12] = (double)n_pred; m.out[13] = (double)n_gt;
    }
}
__global__ void k_pair_first_fill(MetricsState m) {
    // predictions without any pair: empty ranges.  pair_first was set at the first pair of every prediction that has one.
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int next = (int)m.np_;
    m.pair_first[m.Up] = next;
    for (int p = m.Up - 1; p >= 0; --p) { if (m.pair_first[p] < 0) m.pair_first[p] = next; else next = m.pair_first[p]; }
}
// S_assoc (modified_LSTQ.py:57-80): one thread per GT instance with more than min_points points, pairs gt-major with
// the prediction ranks ascending inside, sums in that order
__global__ void k_assoc(MetricsState m, double* __restrict__ per_gt) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= m.na) return;
    const unsigned long long k = m.akey[i];
    const int g = (int)(k / (unsigned long long)(m.Ua + 1));
    if (i > 0 && (int)(m.akey[i - 1] / (unsigned long long)(m.Ua + 1)) == g) return;     // first pair of this GT only
    const double garea = (double)m.gsz[g];
    double inner = 0.0;
    for (long long q = i; q < m.na; ++q) {
        const unsigned long long kq = m.akey[q];
        if ((int)(kq / (unsigned long long)(m.Ua + 1)) != g) break;
        const int a = (int)(kq % (unsigned long long)(m.Ua + 1));
        const long long end = (q + 1 < m.na) ? m.astart[q + 1] : m.navalid;
        const double tpa = (double)(end - m.astart[q]);
        inner += tpa * (tpa / (garea + (double)m.asz[a] - tpa));
    }
    per_gt[g] = inner / garea;
}
__global__ void k_assoc_sum(MetricsState m, const double* __restrict__ per_gt) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double outer = 0.0; int cnt = 0;
    for (int g = 0; g < m.Ug; ++g) {
        if (m.gsz[g] > m.min_points) { outer += per_gt[g]; cnt += 1; }   // strict '>' (:31-32); GT without overlap adds 0
    }
    m.out[14] = outer; m.out[15] = (double)cnt;
}

}  // namespace ancuts

extern "C" {

// remove_semantics(labels, preds, threshold) (point_cloud_utils.py:253-287): every predicted label with more than
// `threshold` of its points on GT background (gt == 0) becomes 0.
int ancuts_remove_semantics(ancuts_handle* h, int64_t n, const int32_t* d_gt_labels, const int32_t* d_pred_labels,
                            double threshold, int32_t* d_out_labels, void* stream) {
    if (!h || n <= 0 || n > 0x7ffffff0LL || !d_gt_labels || !d_pred_labels || !d_out_labels) { set_error("bad argument to ancuts_remove_semantics"); return ANCUTS_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    ANCUTS_CUDA(cudaSetDevice(h->device));
    size_t c1 = 0, c2 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, c1, (int*)nullptr, (int*)nullptr, (int)n, 0, 32);
    cub::DeviceScan::InclusiveSum(nullptr, c2, (int*)nullptr, (int*)nullptr, (int)n);
    const size_t cub_bytes = std::max(c1, c2) + 256;
    PArena ar{nullptr};
    auto lay = [&](PArena& a, int*& sorted, int*& flag, int*& incl, int*& uniq, int*& rk, int*& cnt, int*& bg, long long*& dc, void*& tmp) {
        sorted = a.take<int>(n); flag = a.take<int>(n); incl = a.take<int>(n); uniq = a.take<int>(n); rk = a.take<int>(n);
        cnt = a.take<int>(n); bg = a.take<int>(n); dc = a.take<long long>(4); tmp = a.take<char>(cub_bytes);
    };
    int *sorted, *flag, *incl, *uniq, *rk, *cnt, *bg; long long* dc; void* tmp;
    lay(ar, sorted, flag, incl, uniq, rk, cnt, bg, dc, tmp);
    int rc = post_ws(h, ar.off + 256);
    if (rc) return rc;
    ar = PArena{h->post_ws};
    lay(ar, sorted, flag, incl, uniq, rk, cnt, bg, dc, tmp);
    const int tb = 256; const unsigned g = (unsigned)((n + tb - 1) / tb);
    rc = unique_labels(h, d_pred_labels, n, sorted, flag, incl, uniq, dc, tmp, cub_bytes, st);
    if (rc) return rc;
    PLAUNCH(k_rank_of<<<g, tb, 0, st>>>(d_pred_labels, n, uniq, dc, 0, rk));
    ANCUTS_CUDA(cudaMemsetAsync(cnt, 0, n * sizeof(int), st));
    ANCUTS_CUDA(cudaMemsetAsync(bg, 0, n * sizeof(int), st));
    PLAUNCH(k_sem_count<<<g, tb, 0, st>>>(rk, d_gt_labels, n, cnt, bg));
    PLAUNCH(k_sem_apply<<<g, tb, 0, st>>>(d_pred_labels, rk, n, cnt, bg, threshold, d_out_labels));
    ANCUTS_CUDA(cudaGetLastError());
    return ANCUTS_OK;
}

// Metrics(...).update_stats(all_labels, pred_labels, gt_labels) for a fresh Metrics object, one map
// (metrics_class.py:137-179): h_out[0..6] = p, r, f1, ap, ap0.25, ap0.5, S_assoc; h_out[7..9] = tp, n_pred, n_gt
// at IoU 0.5; h_out[10..20] = AP at the 11 overlaps of metrics_class.py:40.  Divisions by zero give NaN (the
// reference raises ZeroDivisionError for precision / recall, returns 0 for f1).  Blocks until done.
int ancuts_instance_metrics(ancuts_handle* h, int64_t n, const int32_t* d_all_labels, const int32_t* d_pred_labels,
                            const int32_t* d_gt_labels, int min_points, double* h_out, void* stream) {
    if (!h || n <= 0 || n > 0x7ffffff0LL || !d_all_labels || !d_pred_labels || !d_gt_labels || !h_out) { set_error("bad argument to ancuts_instance_metrics"); return ANCUTS_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    ANCUTS_CUDA(cudaSetDevice(h->device));
    size_t c1 = 0, c2 = 0, c3 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, c1, (int*)nullptr, (int*)nullptr, (int)n, 0, 32);
    cub::DeviceScan::InclusiveSum(nullptr, c2, (int*)nullptr, (int*)nullptr, (int)n);
    cub::DeviceRadixSort::SortKeys(nullptr, c3, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)n, 0, 64);
    const size_t cub_bytes = std::max({c1, c2, c3}) + 256;
    struct Bufs {
        int *sorted, *flag, *incl; void* tmp; long long* dc;
        int *uq[3], *rk[3], *cnt[3], *filt[2];          // 0 = pred (filtered), 1 = gt, 2 = all (filtered)
        unsigned long long *key, *skey, *pkey, *akey; int *pstart, *astart; double *iou, *per_gt, *out; int* pair_first;
        unsigned char* used;
    } b;
    auto lay = [&](PArena& a) {
        b.sorted = a.take<int>(n); b.flag = a.take<int>(n); b.incl = a.take<int>(n); b.tmp = a.take<char>(cub_bytes);
        b.dc = a.take<long long>(16);
        for (int i = 0; i < 3; ++i) { b.uq[i] = a.take<int>(n); b.rk[i] = a.take<int>(n); b.cnt[i] = a.take<int>(n + 1); }
        b.filt[0] = a.take<int>(n); b.filt[1] = a.take<int>(n);
        b.key = a.take<unsigned long long>(n); b.skey = a.take<unsigned long long>(n);
        b.pkey = a.take<unsigned long long>(n); b.akey = a.take<unsigned long long>(n);
        b.pstart = a.take<int>(n); b.astart = a.take<int>(n); b.iou = a.take<double>(n); b.per_gt = a.take<double>(n);
        b.out = a.take<double>(16); b.pair_first = a.take<int>(n + 2); b.used = a.take<unsigned char>(12 * (size_t)(n + 1));
    };
    PArena ar{nullptr};
    lay(ar);
    int rc = post_ws(h, ar.off + 256);
    if (rc) return rc;
    ar = PArena{h->post_ws};
    lay(ar);
    const int tb = 256; const unsigned g = (unsigned)((n + tb - 1) / tb);
    long long hu[3]; int first[3];
    // ---- filter_labels on pred and all (:147-148): count per value, small ones -> 0 ----
    const int32_t* src[2] = {d_pred_labels, d_all_labels};
    for (int s = 0; s < 2; ++s) {
        rc = unique_labels(h, src[s], n, b.sorted, b.flag, b.incl, b.uq[0], b.dc, b.tmp, cub_bytes, st);
        if (rc) return rc;
        PLAUNCH(k_rank_of<<<g, tb, 0, st>>>(src[s], n, b.uq[0], b.dc, 0, b.rk[0]));
        ANCUTS_CUDA(cudaMemsetAsync(b.cnt[0], 0, (n + 1) * sizeof(int), st));
        PLAUNCH(k_hist_rank<<<g, tb, 0, st>>>(b.rk[0], n, b.cnt[0]));
        PLAUNCH(k_filter_small<<<g, tb, 0, st>>>(src[s], b.rk[0], n, b.cnt[0], min_points, b.filt[s]));
    }
    // ---- ranks among the non-zero values of filtered pred (0), gt (1), filtered all (2) ----
    const int32_t* lab[3] = {b.filt[0], d_gt_labels, b.filt[1]};
    int U[3], has0[3], total[3];
    for (int s = 0; s < 3; ++s) {
        rc = unique_labels(h, lab[s], n, b.sorted, b.flag, b.incl, b.uq[s], b.dc + s, b.tmp, cub_bytes, st);
        if (rc) return rc;
    }
    ANCUTS_CUDA(cudaMemcpyAsync(h->h_post, b.dc, 3 * sizeof(long long), cudaMemcpyDeviceToHost, st));
    for (int s = 0; s < 3; ++s) ANCUTS_CUDA(cudaMemcpyAsync(&first[s], b.uq[s], sizeof(int), cudaMemcpyDeviceToHost, st));
    ANCUTS_CUDA(cudaStreamSynchronize(st));
    for (int s = 0; s < 3; ++s) {
        hu[s] = h->h_post[s];
        if (first[s] < 0) { set_error("ancuts_instance_metrics: negative labels are not supported"); return ANCUTS_EUNSUPPORTED; }
        has0[s] = (first[s] == 0) ? 1 : 0;
        total[s] = (int)hu[s];
        U[s] = total[s] - has0[s];
    }
    for (int s = 0; s < 3; ++s) {
        long long nf = U[s];
        ANCUTS_CUDA(cudaMemcpyAsync(b.dc + 4 + s, &nf, sizeof(long long), cudaMemcpyHostToDevice, st));
        ANCUTS_CUDA(cudaStreamSynchronize(st));
        PLAUNCH(k_rank_of<<<g, tb, 0, st>>>(lab[s], n, b.uq[s] + has0[s], b.dc + 4 + s, 1, b.rk[s]));
        ANCUTS_CUDA(cudaMemsetAsync(b.cnt[s], 0, (n + 1) * sizeof(int), st));
        PLAUNCH(k_hist_rank<<<g, tb, 0, st>>>(b.rk[s], n, b.cnt[s]));
    }
    MetricsState m;
    memset(&m, 0, sizeof(m));
    m.Up = U[0]; m.Ug = U[1]; m.Ua = U[2];
    m.psz = b.cnt[0]; m.gsz = b.cnt[1]; m.asz = b.cnt[2];
    m.has_zero_pred = has0[0]; m.has_zero_gt = has0[1]; m.up_total = total[0]; m.ug_total = total[1];
    m.iou = b.iou; m.pair_first = b.pair_first; m.used = b.used; m.out = b.out; m.min_points = min_points;
    ANCUTS_CUDA(cudaMemsetAsync(b.out, 0, 16 * sizeof(double), st));
    // ---- (pred, gt) pairs, pred-major: IoU table, greedy matches, AP ----
    auto pair_table = [&](const int* r0, const int* r1, long long stride, int major_gt, unsigned long long* pk, int* ps,
                          long long& npairs, long long& nvalid) -> int {
        PLAUNCH(k_pair_keys<<<g, tb, 0, st>>>(r0, r1, n, stride, b.key, major_gt));
        size_t tmp = cub_bytes;
        h->launches_total += 2;
        ANCUTS_CUDA(cub::DeviceRadixSort::SortKeys(b.tmp, tmp, b.key, b.skey, (int)n, 0, 64, st));
        PLAUNCH(k_run_heads<<<g, tb, 0, st>>>(b.skey, n, b.flag));
        tmp = cub_bytes;
        ANCUTS_CUDA(cub::DeviceScan::InclusiveSum(b.tmp, tmp, b.flag, b.incl, (int)n, st));
        ANCUTS_CUDA(cudaMemsetAsync(b.dc + 8, 0, 2 * sizeof(long long), st));
        PLAUNCH(k_run_emit<<<g, tb, 0, st>>>(b.skey, b.flag, b.incl, n, pk, ps, b.dc + 8, b.dc + 9));
        ANCUTS_CUDA(cudaMemcpyAsync(h->h_post, b.dc + 8, 2 * sizeof(long long), cudaMemcpyDeviceToHost, st));
        ANCUTS_CUDA(cudaStreamSynchronize(st));
        npairs = h->h_post[0]; nvalid = h->h_post[1];
        return ANCUTS_OK;
    };
    long long np_ = 0, nvalid = 0;
    rc = pair_table(b.rk[0], b.rk[1], (long long)m.Ug + 1, 0, b.pkey, b.pstart, np_, nvalid);
    if (rc) return rc;
    m.pkey = b.pkey; m.pstart = b.pstart; m.np_ = np_; m.nvalid = nvalid;
    PLAUNCH(k_fill_i32<<<(unsigned)((m.Up + 2 + tb - 1) / tb), tb, 0, st>>>(b.pair_first, m.Up + 2, -1));
    if (np_ > 0) PLAUNCH(k_pair_iou<<<(unsigned)((np_ + tb - 1) / tb), tb, 0, st>>>(m));
    PLAUNCH(k_pair_first_fill<<<1, 32, 0, st>>>(m));
    PLAUNCH(k_greedy<<<1, 32, 0, st>>>(m));
    // ---- (gt, all) pairs, gt-major: S_assoc ----
    long long na = 0, navalid = 0;
    rc = pair_table(b.rk[2], b.rk[1], (long long)m.Ua + 1, 1, b.akey, b.astart, na, navalid);
    if (rc) return rc;
    m.akey = b.akey; m.astart = b.astart; m.na = na; m.navalid = navalid;
    ANCUTS_CUDA(cudaMemsetAsync(b.per_gt, 0, (size_t)std::max(m.Ug, 1) * sizeof(double), st));
    if (na > 0) PLAUNCH(k_assoc<<<(unsigned)((na + tb - 1) / tb), tb, 0, st>>>(m, b.per_gt));
    PLAUNCH(k_assoc_sum<<<1, 32, 0, st>>>(m, b.per_gt));
    ANCUTS_CUDA(cudaGetLastError());
    double o[16];
    ANCUTS_CUDA(cudaMemcpyAsync(o, b.out, sizeof(o), cudaMemcpyDeviceToHost, st));
    ANCUTS_CUDA(cudaStreamSynchronize(st));
    const double tp = o[11], n_pred = o[12], n_gt = o[13];
    const double prec = tp / n_pred, rec = tp / n_gt;                      // :325-326
    double f1 = (prec + rec) != 0.0 ? 2.0 * (prec * rec) / (prec + rec) : 0.0;    // :327-330
    double ap = 0.0;
    for (int t = 1; t < 11; ++t) ap += o[t];                               // AP_OVERLAPS = OVERLAPS[1:], in order
    ap /= 10.0;
    h_out[0] = prec; h_out[1] = rec; h_out[2] = f1; h_out[3] = ap; h_out[4] = o[0]; h_out[5] = o[1];
    h_out[6] = (o[15] > 0.0) ? o[14] / o[15] : nan("");
    h_out[7] = tp; h_out[8] = n_pred; h_out[9] = n_gt;
    for (int t = 0; t < 11; ++t) h_out[10 + t] = o[t];
    return ANCUTS_OK;
}

}  // extern "C"
