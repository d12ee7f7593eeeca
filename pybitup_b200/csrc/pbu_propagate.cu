// K3: Monte-Carlo uncertainty propagation (Propagation.propagate, emulator == "None",
// solve_problem.py:467-647): evaluate the forward model at S parameter vectors with the same device model
// functions the samplers use, keep the evaluations in HBM point-major, and reduce per design point
//   mean, min, max                      (:596-627, written to <id>_interval.csv :629-637)
//   two percentiles (2.5 / 97.5)        (np.percentile over all evaluations, :639-645, <id>_CI.csv)
// The percentiles are EXACT order statistics found by an 8-pass most-significant-digit radix select over the
// HBM-resident evaluations: each pass builds, per design point and per wanted rank, a 256-bin histogram of the next
// byte of the order-preserving 64-bit key among the values that match the prefix found so far.  The histograms
// (and the moment partials) are the only thing that has to be summed over GPUs when samples are sharded
// (one small all-reduce per pass, done by the host layer), replacing the reference's pickled MPI send of every
// evaluation to rank 0 (:578-592).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "pbu_engine_internal.h"

using namespace pbu;

struct pbu_propagation {
    pbu_engine *e = nullptr;
    int64_t S = 0;
    int N = 0;
    double *d_X = nullptr;       // [d][S] parameter vectors (structure of arrays)
    double *d_f = nullptr;       // [N][S] model evaluations, point-major
    double *d_part = nullptr;    // [N][NSPLIT][3] moment partials
    // radix select state
    int T = 0;                   // wanted ranks per design point
    unsigned long long *d_prefix = nullptr;   // [N][T] key prefix found so far
    long long *d_rank = nullptr;              // [N][T] remaining rank inside the prefix class
    unsigned long long *d_hist = nullptr;     // [N][T][256]
    bool evaluated = false;
    // bracketed select (large S): subsample, bracket values, candidates between the brackets
    double *d_sub = nullptr;                  // [N][sub_m] every stride-th 256-chunk of each row
    int64_t sub_m = 0;
    double *d_brk = nullptr;                  // [N][nq][2] bracket (a, b) per design point and percentile
    unsigned long long *d_cnt = nullptr;      // [2][N][nq]: values below a | values inside [a, b]
    double *d_cand = nullptr;                 // [N][nq][cap] the values inside [a, b], unordered
    int64_t cap = 0;
    int nq = 0;
    bool cand_mode = false;                   // select_pass reads the candidates instead of the evaluations
};

static constexpr int NSPLIT = 16;
static constexpr int MAX_T = 4;

// ------------------------------------------------------------------------------------------------ kernels
// rows of N(mean, diag(std^2)) by Philox: X[i][s] = mean_i + std_i z(seed, sample_offset + s, i)
__global__ void gaussian_samples_kernel(double *X, int64_t S, int d, const double *mean, const double *stdv, uint64_t seed, uint64_t sample_offset) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const uint64_t g = sample_offset + (uint64_t)s;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    for (int b = 0; b * 4 < d; ++b) {
        const uint4 r = Philox::round10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), 0x50524f50u /* 'PROP' */, (uint32_t)b), key);
        double z[4];
        box_muller(r.x, r.y, z[0], z[1]);
        box_muller(r.z, r.w, z[2], z[3]);
        for (int j = 0; j < 4 && b * 4 + j < d; ++j) X[(int64_t)(b * 4 + j) * S + s] = mean[b * 4 + j] + stdv[b * 4 + j] * z[j];
    }
}

// Slice `split` of `n_split` of row `row` of f[N][S]: [lo, hi).  Interior boundaries fall on 256-byte lines of the
// ARRAY (not of the row), so that a warp's 256-byte loads never straddle lines: with streaming (evict-first) loads a
// straddled line is fetched from DRAM twice (ncu: 60 GB read for a 50 GB pass before this).
__device__ __forceinline__ void row_slice(int row, int split, int n_split, int64_t S, int64_t &lo, int64_t &hi) {
    const int64_t per = (S + n_split - 1) / n_split, off = (int64_t)row * S;
    auto cut = [&](int64_t k) -> int64_t {          // boundary k of the row, moved up to the next line of the array
        if (k <= 0) return 0;
        if (k >= S) return S;
        const int64_t a = ((off + k + 31) & ~(int64_t)31) - off;
        return a < S ? a : S;
    };
    lo = cut((int64_t)split * per);
    hi = cut((int64_t)(split + 1) * per);
}

// per design point (row of f) and slice: sum, min, max
__global__ void __launch_bounds__(256) row_moments_kernel(const double *f, int64_t S, double *part) {
    const int row = blockIdx.x, split = blockIdx.y;
    int64_t lo, hi;
    row_slice(row, split, gridDim.y, S, lo, hi);
    const double *r = f + (int64_t)row * S;
    double sum = 0.0, mn = INFINITY, mx = -INFINITY;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const double v = r[i];
        sum += v;
        mn = fmin(mn, v);
        mx = fmax(mx, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    __shared__ double sh[3][8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        sh[0][w] = sum;
        sh[1][w] = mn;
        sh[2][w] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) {
            sum += sh[0][i];
            mn = fmin(mn, sh[1][i]);
            mx = fmax(mx, sh[2][i]);
        }
        double *o = part + ((int64_t)row * gridDim.y + split) * 3;
        o[0] = sum;
        o[1] = mn;
        o[2] = mx;
    }
}

// order-preserving map double -> uint64 (negative values reversed, sign bit flipped)
__device__ __forceinline__ unsigned long long order_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
static inline unsigned long long order_key_host(double v) {
    unsigned long long b;
    memcpy(&b, &v, sizeof b);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ inline double key_to_double(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    double v;
#ifdef __CUDA_ARCH__
    v = __longlong_as_double((long long)b);
#else
    memcpy(&v, &b, sizeof v);
#endif
    return v;
}

// one pass of the MSD radix select: byte `pass` (0 = most significant) of the keys that match prefix[row][t].
//   * eight loads in flight per thread (the pass is a pure HBM stream of the evaluations);
//   * ranks whose prefixes coincide (the two order statistics around one percentile nearly always do) are counted once;
//   * each thread run-length encodes its matches: the samples of a design point cluster, so in the early passes nearly
//     every key falls in the same bin -- one shared-memory atomic per run instead of a 32-way conflict per key.
// run-length counting of one rank's matches (all state in scalars: arrays indexed by the rank end up in local memory)
__device__ __forceinline__ void select_feed(bool on, unsigned long long km, unsigned long long pf, int digit, int &cur, unsigned int &cnt,
                                            unsigned int *bins) {
    if (on && km == pf) {
        if (digit == cur) {
            cnt += 1u;
        } else {
            if (cnt) atomicAdd(&bins[cur], cnt);
            cur = digit;
            cnt = 1u;
        }
    }
}

__global__ void __launch_bounds__(256) select_hist_kernel(const double *f, int64_t S, int T, int pass, const unsigned long long *prefix,
                                                          unsigned long long *hist) {
    static_assert(MAX_T == 4, "the kernel is written out for four ranks");
    __shared__ unsigned int sh[MAX_T][256];
    const int row = blockIdx.x, split = blockIdx.y;
    for (int i = threadIdx.x; i < MAX_T * 256; i += blockDim.x) (&sh[0][0])[i] = 0u;
    __syncthreads();
    const int shift = 56 - 8 * pass;
    const unsigned long long mask = (pass == 0) ? 0ull : (~0ull << (shift + 8));
    const unsigned long long *pr = prefix + (int64_t)row * T;
    const unsigned long long pf0 = pr[0] & mask, pf1 = (T > 1) ? (pr[1] & mask) : 0ull, pf2 = (T > 2) ? (pr[2] & mask) : 0ull,
                             pf3 = (T > 3) ? (pr[3] & mask) : 0ull;
    // rep_t: first rank with the same (masked) prefix -- the one that is counted
    const int rep1 = (pf1 == pf0) ? 0 : 1;
    const int rep2 = (pf2 == pf0) ? 0 : ((pf2 == pf1) ? 1 : 2);
    const int rep3 = (pf3 == pf0) ? 0 : ((pf3 == pf1) ? 1 : ((pf3 == pf2) ? 2 : 3));
    const bool on1 = T > 1 && rep1 == 1, on2 = T > 2 && rep2 == 2, on3 = T > 3 && rep3 == 3;
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    unsigned int n0 = 0u, n1 = 0u, n2 = 0u, n3 = 0u;
    int64_t lo, hi;
    row_slice(row, split, gridDim.y, S, lo, hi);
    const double *r = f + (int64_t)row * S;
    int64_t i = lo + threadIdx.x;
    const int64_t st = blockDim.x;
#define PBU_SELECT_FEED(v)                                    \
    do {                                                      \
        const unsigned long long k_ = order_key(v);           \
        const int dg_ = (int)((k_ >> shift) & 255ull);        \
        const unsigned long long km_ = k_ & mask;             \
        select_feed(true, km_, pf0, dg_, c0, n0, sh[0]);      \
        select_feed(on1, km_, pf1, dg_, c1, n1, sh[1]);       \
        select_feed(on2, km_, pf2, dg_, c2, n2, sh[2]);       \
        select_feed(on3, km_, pf3, dg_, c3, n3, sh[3]);       \
    } while (0)
    for (; i + 7 * st < hi; i += 8 * st) {          // eight loads in flight per thread: ~64 KB per SM, what 6 TB/s needs
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcs(r + i + u * st);
#pragma unroll
        for (int u = 0; u < 8; ++u) PBU_SELECT_FEED(v[u]);
    }
    for (; i < hi; i += st) {
        const double v = r[i];
        PBU_SELECT_FEED(v);
    }
#undef PBU_SELECT_FEED
    if (n0) atomicAdd(&sh[0][c0], n0);
    if (n1) atomicAdd(&sh[1][c1], n1);
    if (n2) atomicAdd(&sh[2][c2], n2);
    if (n3) atomicAdd(&sh[3][c3], n3);
    __syncthreads();
    for (int j = threadIdx.x; j < T * 256; j += blockDim.x) {
        const int t = j >> 8;
        const int src = (t == 0) ? 0 : ((t == 1) ? rep1 : ((t == 2) ? rep2 : rep3));
        const unsigned int v = sh[src][j & 255];
        if (v) atomicAdd(&hist[(int64_t)row * T * 256 + j], (unsigned long long)v);
    }
}

// after the (all-reduced) histogram of a pass: pick the digit that holds the wanted rank.  One warp per (point, rank):
// each lane owns eight consecutive bins, a shuffle scan finds the lane, the lane finds the bin.
__global__ void select_update_kernel(int n_rows, int T, int pass, unsigned long long *prefix, long long *rank, const unsigned long long *hist) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, l = threadIdx.x & 31;
    if (i >= n_rows * T) return;
    const unsigned long long *h = hist + (int64_t)i * 256 + 8 * l;
    long long c[8], tot = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        c[b] = (long long)h[b];
        tot += c[b];
    }
    long long incl = tot;
    for (int o = 1; o < 32; o <<= 1) {
        const long long up = __shfl_up_sync(0xffffffffu, incl, o);
        if (l >= o) incl += up;
    }
    const long long k = rank[i];
    // first lane whose inclusive count exceeds k holds the bin (none: the last bin, as a guard against a bad rank)
    const unsigned hit = __ballot_sync(0xffffffffu, k < incl);
    const int owner = hit ? (__ffs(hit) - 1) : 31;
    if (l == owner) {
        long long r = k - (incl - tot);
        int digit = 8 * l + 7;
        for (int b = 0; b < 8; ++b) {
            if (r < c[b]) {
                digit = 8 * l + b;
                break;
            }
            r -= c[b];
        }
        if (!hit) digit = 255;
        rank[i] = r;
        prefix[i] |= (unsigned long long)digit << (56 - 8 * pass);
    }
}

// ------------------------------------------------------------------------------------------------ bracketed select
// For large S the eight full passes of the radix select are replaced by ONE pass over the evaluations:
//   1. a subsample (every stride-th chunk of 256 values of each row) is radix-selected at ranks a safe distance either
//      side of each percentile: a bracket [a, b] that holds the wanted order statistics unless the subsample is freakishly
//      unrepresentative (the counts of step 2 tell; the caller then falls back to the full radix select);
//   2. one pass over the evaluations accumulates sum / min / max, counts the values below a and copies the values
//      inside [a, b] (a fraction of a percent) to a candidate buffer;
//   3. the radix select runs on the candidates alone.
// Sharded over GPUs the brackets are min / max-reduced, the counts summed, and step 3 sums histograms as before.
__global__ void subsample_gather_kernel(const double *f, int64_t S, int stride, int64_t m, double *sub) {
    const int row = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int64_t src = (i >> 8) * 256 * stride + (i & 255);
    sub[(int64_t)row * m + i] = f[(int64_t)row * S + src];
}

// the warp hands its n parked values over: one global atomic reserves the slots; `store` false (a == b, a tie across the
// bracket: a constant design point) counts without copying
__device__ __noinline__ void bracket_flush(const double *buf, int n, unsigned long long *counter, double *dst, int64_t cap, int lane, bool store) {
    __syncwarp();
    unsigned long long off = 0ull;
    if (lane == 0) off = atomicAdd(counter, (unsigned long long)n);
    off = __shfl_sync(0xffffffffu, off, 0);
    if (store)
        for (int j = lane; j < n; j += 32)
            if ((int64_t)(off + j) < cap) dst[off + j] = buf[j];
    __syncwarp();
}

template <int NQ>
__global__ void __launch_bounds__(256) bracket_pass_kernel(const double *f, int64_t S, const double *brk, double *part, unsigned long long *below,
                                                           unsigned long long *inside, double *cand, int64_t cap) {
    __shared__ double wbuf[8][NQ][64];        // per warp and percentile: values waiting to be flushed (at most 63)
    __shared__ double sh[3][8];
    const int row = blockIdx.x, split = blockIdx.y, w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const unsigned lanes_below = (1u << l) - 1u;
    double a[NQ], b[NQ];
    unsigned int nb[NQ];
    int wc[NQ];                               // warp-uniform fill of wbuf[w][q]
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        a[q] = brk[((int64_t)row * NQ + q) * 2];
        b[q] = brk[((int64_t)row * NQ + q) * 2 + 1];
        nb[q] = 0u;
        wc[q] = 0;
    }
    int64_t lo, hi;
    row_slice(row, split, gridDim.y, S, lo, hi);
    const double *r = f + (int64_t)row * S;
    double sum = 0.0, mn = INFINITY, mx = -INFINITY;
    // one warp-wide flush of wbuf[w][q] (bracket_flush: out of line, the pass is sixteen copies of feed() and the
    // instruction cache was missing as often as the loads were stalling)
    auto flush = [&](int q) {
        bracket_flush(&wbuf[w][q][0], wc[q], &inside[(int64_t)row * NQ + q], cand + ((int64_t)row * NQ + q) * cap, cap, l, a[q] < b[q]);
        wc[q] = 0;
    };
    // per value: four compares, one add, two compare-and-keep (fmin / fmax cost seven instructions each for their NaN
    // rules; NaN is ignored here as there) -- the pass is bound by instruction issue, not by HBM, beyond ~44 issue
    // cycles per warp and value
    auto feed = [&](bool valid, double v) {
        bool in[NQ], any = false;
        if (valid) {
            sum += v;
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const bool ge = v >= a[q];          // one compare serves the count and the range test (a NaN counts as below)
            nb[q] += (valid && !ge) ? 1u : 0u;
            in[q] = valid && ge && v <= b[q];
            any = any || in[q];
        }
        if (__any_sync(0xffffffffu, any)) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const unsigned msk = __ballot_sync(0xffffffffu, in[q]);
                if (msk) {
                    if (in[q]) wbuf[w][q][wc[q] + __popc(msk & lanes_below)] = v;
                    wc[q] += __popc(msk);
                    if (wc[q] >= 32) flush(q);
                }
            }
        }
    };
    const int64_t st = blockDim.x;
    int64_t base = lo;
    for (; base + 8 * st <= hi; base += 8 * st) {          // block-uniform trip count: the warp votes need every lane
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcs(r + base + threadIdx.x + u * st);
#pragma unroll
        for (int u = 0; u < 8; ++u) feed(true, v[u]);
    }
    for (; base < hi; base += st) {
        const int64_t i = base + threadIdx.x;
        const bool valid = i < hi;
        feed(valid, valid ? r[i] : 0.0);
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        if (wc[q] > 0) flush(q);
        unsigned int c = nb[q];
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (l == 0 && c) atomicAdd(&below[(int64_t)row * NQ + q], (unsigned long long)c);
    }
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (l == 0) {
        sh[0][w] = sum;
        sh[1][w] = mn;
        sh[2][w] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) {
            sum += sh[0][i];
            mn = fmin(mn, sh[1][i]);
            mx = fmax(mx, sh[2][i]);
        }
        double *o = part + ((int64_t)row * gridDim.y + split) * 3;
        o[0] = sum;
        o[1] = mn;
        o[2] = mx;
    }
}

// one radix pass over the candidates of (row, percentile): the two ranks 2q, 2q+1 of that percentile; one block each,
// so the histogram rows are written, not accumulated
__global__ void __launch_bounds__(256) cand_hist_kernel(const double *cand, const unsigned long long *inside, int64_t cap, int nq, int pass,
                                                        const unsigned long long *prefix, unsigned long long *hist) {
    __shared__ unsigned int sh[2][256];
    const int rq = blockIdx.x, row = rq / nq, q = rq - row * nq, T = 2 * nq;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) (&sh[0][0])[i] = 0u;
    __syncthreads();
    const int shift = 56 - 8 * pass;
    const unsigned long long mask = (pass == 0) ? 0ull : (~0ull << (shift + 8));
    const unsigned long long pf0 = prefix[(int64_t)row * T + 2 * q] & mask, pf1 = prefix[(int64_t)row * T + 2 * q + 1] & mask;
    const bool two = pf1 != pf0;
    unsigned long long n = inside[rq];
    if (n > (unsigned long long)cap) n = (unsigned long long)cap;
    const double *c = cand + (int64_t)rq * cap;
    for (unsigned long long i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long k = order_key(c[i]);
        const unsigned long long km = k & mask;
        const int dg = (int)((k >> shift) & 255ull);
        if (km == pf0) atomicAdd(&sh[0][dg], 1u);
        if (two && km == pf1) atomicAdd(&sh[1][dg], 1u);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < 512; j += blockDim.x) {
        const int t = j >> 8;
        hist[((int64_t)row * T + 2 * q + t) * 256 + (j & 255)] = sh[(t == 1 && !two) ? 0 : t][j & 255];
    }
}

// ------------------------------------------------------------------------------------------------ C ABI
// bracketed select: subsample length for a stride, and the candidate slots that hold four times the expected number of
// values in a bracket six binomial standard deviations either side of fraction f of an m-value subsample
static int64_t bracket_sub_len(int64_t S, int stride) {
    const int64_t group = (int64_t)256 * stride, full = S / group, rem = S - full * group;
    return full * 256 + std::min<int64_t>(rem, 256);
}
static int64_t bracket_cap(int64_t cap_hint) { return ((4 * cap_hint + 4096 + 255) / 256) * 256; }
static constexpr int BRACKET_STRIDE_DEFAULT = 64;
static constexpr int64_t BRACKET_MIN_SAMPLES = 1 << 18;

PBU_API int pbu_propagation_create(pbu_engine *e, int64_t n_samples, pbu_propagation **out) {
    if (!e || !out || n_samples < 1) return fail(PBU_ERR_INVALID, "bad argument");
    CUDA_TRY(cudaSetDevice(e->device));
    pbu_propagation *p = new pbu_propagation();
    p->e = e;
    p->S = n_samples;
    p->N = e->prob.n_points;
    const size_t fbytes = (size_t)p->N * (size_t)n_samples * sizeof(double);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    // large S: the buffers of the bracketed select are taken now (any one or two percentiles fit: sized for the median),
    // so that statistics never allocates -- 1.6 % + ~10 % of the evaluations
    int64_t sub_m = 0, cap = 0;
    if (n_samples >= BRACKET_MIN_SAMPLES) {
        sub_m = bracket_sub_len(n_samples, BRACKET_STRIDE_DEFAULT);
        const double dl = 6.0 * std::sqrt(0.25 * (double)sub_m) + 8.0;
        cap = bracket_cap((int64_t)((double)n_samples * (2.0 * dl + 3.0) / (double)sub_m));
    }
    const size_t bbytes = ((size_t)sub_m + 2 * (size_t)cap) * p->N * sizeof(double);
    if (fbytes + bbytes + (size_t)e->d * n_samples * 8 + (64u << 20) > free_b) {
        delete p;
        return fail(PBU_ERR_INVALID, "propagation of %lld samples x %d points needs %.1f GB of device memory, %.1f GB free: shard the samples",
                    (long long)n_samples, e->prob.n_points, fbytes / 1e9, free_b / 1e9);
    }
    cudaError_t ce = cudaMalloc(&p->d_X, (size_t)e->d * n_samples * sizeof(double));
    if (ce == cudaSuccess) ce = cudaMalloc(&p->d_f, fbytes);
    if (ce == cudaSuccess) ce = cudaMalloc(&p->d_part, (size_t)p->N * NSPLIT * 3 * sizeof(double));
    if (ce == cudaSuccess) ce = cudaMalloc(&p->d_prefix, (size_t)p->N * MAX_T * sizeof(unsigned long long));
    if (ce == cudaSuccess) ce = cudaMalloc(&p->d_rank, (size_t)p->N * MAX_T * sizeof(long long));
    if (ce == cudaSuccess) ce = cudaMalloc(&p->d_hist, (size_t)p->N * MAX_T * 256 * sizeof(unsigned long long));
    if (ce == cudaSuccess) ce = cudaMalloc(&p->d_brk, (size_t)p->N * 2 * 2 * sizeof(double));
    if (ce == cudaSuccess) ce = cudaMalloc(&p->d_cnt, (size_t)2 * p->N * 2 * sizeof(unsigned long long));
    if (ce == cudaSuccess && sub_m > 0) {
        ce = cudaMalloc(&p->d_sub, (size_t)p->N * sub_m * sizeof(double));
        if (ce == cudaSuccess) ce = cudaMalloc(&p->d_cand, (size_t)p->N * 2 * cap * sizeof(double));
        if (ce == cudaSuccess) p->sub_m = sub_m, p->cap = cap;
    }
    if (ce != cudaSuccess) {
        cudaFree(p->d_sub), cudaFree(p->d_cand), cudaFree(p->d_brk), cudaFree(p->d_cnt);
        cudaFree(p->d_X), cudaFree(p->d_f), cudaFree(p->d_part), cudaFree(p->d_prefix), cudaFree(p->d_rank), cudaFree(p->d_hist);
        delete p;
        return fail(PBU_ERR_CUDA, "propagation buffers: %s", cudaGetErrorString(ce));
    }
    *out = p;
    return PBU_OK;
}

PBU_API int pbu_propagation_destroy(pbu_propagation *p) {
    if (!p) return PBU_OK;
    cudaSetDevice(p->e->device);
    cudaStreamSynchronize(p->e->stream);
    cudaFree(p->d_X), cudaFree(p->d_f), cudaFree(p->d_part), cudaFree(p->d_prefix), cudaFree(p->d_rank), cudaFree(p->d_hist);
    cudaFree(p->d_sub), cudaFree(p->d_brk), cudaFree(p->d_cnt), cudaFree(p->d_cand);
    delete p;
    return PBU_OK;
}

static int run_model_eval(pbu_propagation *p) {
    pbu_engine *e = p->e;
    // Models with lookup tables (Arrhenius): conflict-free replicated tables in 1024-thread persistent blocks, one per SM,
    // once there are samples for two rounds of them (PBU_REP_TABLES=0 / 1 overrides: diagnostics)
    if (e->ev->eval_model_rep && e->prob.n_blocks <= 1 && e->tile_points_raw >= e->prob.n_points) {
        int n_sm = 0, max_smem = 0;
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, e->device);
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, e->device);
        const size_t smem = e->smem_raw_bytes + 128 + PBU_REP_TABLE_BYTES;
        bool use = p->S >= (int64_t)2 * 1024 * n_sm;
        if (const char *f = getenv("PBU_REP_TABLES")) use = atoi(f) != 0;
        if (use && smem <= (size_t)max_smem) {
            CUDA_TRY(cudaFuncSetAttribute(e->ev->eval_model_rep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int64_t need = (p->S + 1023) / 1024;
            const unsigned grid = (unsigned)std::min<int64_t>(need, n_sm);
            void *args[] = {(void *)&e->prob, (void *)&p->d_X, (void *)&p->S, (void *)&p->d_f, (void *)&e->tile_points_raw};
            CUDA_TRY(cudaLaunchKernel(e->ev->eval_model_rep, dim3(grid), dim3(1024), args, smem, e->stream));
            p->evaluated = true;
            return PBU_OK;
        }
    }
    CUDA_TRY(cudaFuncSetAttribute(e->ev->eval_model, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_raw_bytes));
    const int block = 128;
    const unsigned grid = (unsigned)((p->S + block - 1) / block);
    void *args[] = {(void *)&e->prob, (void *)&p->d_X, (void *)&p->S, (void *)&p->d_f, (void *)&e->tile_points_raw};
    CUDA_TRY(cudaLaunchKernel(e->ev->eval_model, dim3(grid), dim3(block), args, e->smem_raw_bytes, e->stream));
    p->evaluated = true;
    return PBU_OK;
}

// the loop body of solve_problem.py:553-576 for all S samples: X host [S][d] -> f = run_model(X[s]) on the device
PBU_API int pbu_propagation_eval_host(pbu_propagation *p, const double *X_host) {
    if (!p || !X_host) return fail(PBU_ERR_INVALID, "null argument");
    pbu_engine *e = p->e;
    CUDA_TRY(cudaSetDevice(e->device));
    std::vector<double> soa((size_t)e->d * p->S);
    for (int64_t s = 0; s < p->S; ++s)
        for (int i = 0; i < e->d; ++i) soa[(size_t)i * p->S + s] = X_host[(size_t)s * e->d + i];
    CUDA_TRY(cudaMemcpyAsync(p->d_X, soa.data(), soa.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return run_model_eval(p);
}

// same with the samples drawn on the device: row s ~ N(mean, diag(std^2)), Philox keyed by (seed, sample_offset + s)
PBU_API int pbu_propagation_eval_gaussian(pbu_propagation *p, const double *mean_host, const double *std_host, uint64_t seed, uint64_t sample_offset) {
    if (!p || !mean_host || !std_host) return fail(PBU_ERR_INVALID, "null argument");
    pbu_engine *e = p->e;
    CUDA_TRY(cudaSetDevice(e->device));
    double *d_ms = nullptr;
    CUDA_TRY(cudaMalloc(&d_ms, 2 * e->d * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(d_ms, mean_host, e->d * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaMemcpyAsync(d_ms + e->d, std_host, e->d * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    gaussian_samples_kernel<<<(unsigned)((p->S + 255) / 256), 256, 0, e->stream>>>(p->d_X, p->S, e->d, d_ms, d_ms + e->d, seed, sample_offset);
    CUDA_TRY(cudaGetLastError());
    int rc = run_model_eval(p);
    cudaStreamSynchronize(e->stream);
    cudaFree(d_ms);
    return rc;
}

// device pointers of the sample matrix [d][S] and the evaluations [N][S] (for callers that hand over torch tensors)
PBU_API int pbu_propagation_buffers(pbu_propagation *p, double **X_dev, double **f_dev) {
    if (!p) return fail(PBU_ERR_INVALID, "null argument");
    if (X_dev) *X_dev = p->d_X;
    if (f_dev) *f_dev = p->d_f;
    return PBU_OK;
}

// per design point: sum, min, max over this engine's samples (host [N] each); sums are added over GPUs by the caller
PBU_API int pbu_propagation_moments(pbu_propagation *p, double *sum_host, double *min_host, double *max_host) {
    if (!p || !p->evaluated) return fail(PBU_ERR_INVALID, "propagation has not been evaluated");
    pbu_engine *e = p->e;
    CUDA_TRY(cudaSetDevice(e->device));
    row_moments_kernel<<<dim3(p->N, NSPLIT), 256, 0, e->stream>>>(p->d_f, p->S, p->d_part);
    CUDA_TRY(cudaGetLastError());
    std::vector<double> part((size_t)p->N * NSPLIT * 3);
    CUDA_TRY(cudaMemcpyAsync(part.data(), p->d_part, part.size() * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    for (int k = 0; k < p->N; ++k) {
        double s = 0.0, mn = INFINITY, mx = -INFINITY;
        for (int j = 0; j < NSPLIT; ++j) {
            const double *o = &part[((size_t)k * NSPLIT + j) * 3];
            s += o[0];
            mn = std::min(mn, o[1]);
            mx = std::max(mx, o[2]);
        }
        if (sum_host) sum_host[k] = s;
        if (min_host) min_host[k] = mn;
        if (max_host) max_host[k] = mx;
    }
    return PBU_OK;
}

// radix select of `n_ranks` order statistics per design point.  ranks host [n_ranks]: 0-based ranks in the sorted
// GLOBAL sample (all GPUs).  Protocol per pass first_pass..7: select_pass fills hist_host [N][n_ranks][256] with this
// engine's counts; the caller sums them over GPUs and hands the total to select_update.
PBU_API int pbu_propagation_select_begin(pbu_propagation *p, const int64_t *ranks, int n_ranks, const double *row_min, const double *row_max,
                                         int *first_pass) {
    if (!p || !ranks || n_ranks < 1 || n_ranks > MAX_T) return fail(PBU_ERR_INVALID, "need 1..%d ranks", MAX_T);
    if (!p->evaluated) return fail(PBU_ERR_INVALID, "propagation has not been evaluated");
    if ((row_min == nullptr) != (row_max == nullptr)) return fail(PBU_ERR_INVALID, "row_min and row_max go together");
    pbu_engine *e = p->e;
    CUDA_TRY(cudaSetDevice(e->device));
    p->T = n_ranks;
    p->cand_mode = false;
    std::vector<long long> r((size_t)p->N * n_ranks);
    std::vector<unsigned long long> pre((size_t)p->N * n_ranks, 0ull);
    for (int k = 0; k < p->N; ++k)
        for (int t = 0; t < n_ranks; ++t) r[(size_t)k * n_ranks + t] = ranks[t];
    // With the GLOBAL minimum and maximum of every design point (the moments, reduced over all GPUs) the leading bytes
    // that all its keys share are known: those passes are skipped, for every point alike (the smallest shared length).
    int skip = 0;
    if (row_min) {
        skip = 8;
        for (int k = 0; k < p->N; ++k) {
            if (!(row_min[k] <= row_max[k])) {          // NaN bounds: no information
                skip = 0;
                break;
            }
            const unsigned long long x = order_key_host(row_min[k]) ^ order_key_host(row_max[k]);
            int same = 8;
            for (int b = 0; b < 8; ++b)
                if ((x >> (56 - 8 * b)) & 255ull) {
                    same = b;
                    break;
                }
            skip = std::min(skip, same);
        }
        const unsigned long long keep = (skip == 0) ? 0ull : ((skip >= 8) ? ~0ull : (~0ull << (64 - 8 * skip)));
        for (int k = 0; k < p->N; ++k)
            for (int t = 0; t < n_ranks; ++t) pre[(size_t)k * n_ranks + t] = order_key_host(row_min[k]) & keep;
    }
    if (first_pass) *first_pass = skip;
    CUDA_TRY(cudaMemcpyAsync(p->d_rank, r.data(), r.size() * sizeof(long long), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaMemcpyAsync(p->d_prefix, pre.data(), pre.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return PBU_OK;
}

// one pass into `hist` (device, [N][T][256]): the evaluations, or the candidates of the bracketed select
static int select_pass_into(pbu_propagation *p, int pass, unsigned long long *hist) {
    pbu_engine *e = p->e;
    const size_t n = (size_t)p->N * p->T * 256;
    CUDA_TRY(cudaMemsetAsync(hist, 0, n * sizeof(unsigned long long), e->stream));
    if (p->cand_mode)
        cand_hist_kernel<<<p->N * p->nq, 256, 0, e->stream>>>(p->d_cand, p->d_cnt + (size_t)p->N * p->nq, p->cap, p->nq, pass, p->d_prefix, hist);
    else
        select_hist_kernel<<<dim3(p->N, NSPLIT), 256, 0, e->stream>>>(p->d_f, p->S, p->T, pass, p->d_prefix, hist);
    CUDA_TRY(cudaGetLastError());
    return PBU_OK;
}

PBU_API int pbu_propagation_select_pass(pbu_propagation *p, int pass, uint64_t *hist_host) {
    if (!p || p->T < 1 || pass < 0 || pass > 7) return fail(PBU_ERR_INVALID, "bad argument");
    pbu_engine *e = p->e;
    CUDA_TRY(cudaSetDevice(e->device));
    const size_t n = (size_t)p->N * p->T * 256;
    if (int rc = select_pass_into(p, pass, p->d_hist)) return rc;
    if (hist_host) {          // NULL: single GPU, the counts stay on the device for select_update(..., NULL)
        CUDA_TRY(cudaMemcpyAsync(hist_host, p->d_hist, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
        CUDA_TRY(cudaStreamSynchronize(e->stream));
    }
    return PBU_OK;
}

PBU_API int pbu_propagation_select_update(pbu_propagation *p, int pass, const uint64_t *hist_total_host) {
    if (!p || p->T < 1 || pass < 0 || pass > 7) return fail(PBU_ERR_INVALID, "bad argument");
    pbu_engine *e = p->e;
    CUDA_TRY(cudaSetDevice(e->device));
    const size_t n = (size_t)p->N * p->T * 256;
    if (hist_total_host)      // NULL: use the counts select_pass left on the device (single GPU, nothing to sum)
        CUDA_TRY(cudaMemcpyAsync(p->d_hist, hist_total_host, n * sizeof(unsigned long long), cudaMemcpyHostToDevice, e->stream));
    const int items = p->N * p->T;
    select_update_kernel<<<(items * 32 + 127) / 128, 128, 0, e->stream>>>(p->N, p->T, pass, p->d_prefix, p->d_rank, p->d_hist);
    CUDA_TRY(cudaGetLastError());
    if (hist_total_host) CUDA_TRY(cudaStreamSynchronize(e->stream));
    return PBU_OK;
}

// The same two steps with the histogram in a caller-owned DEVICE buffer ([N][n_ranks][256] uint64, e.g. a torch tensor):
// select_pass_dev fills it on the engine's stream, the caller all-reduces it in place on that stream (NCCL), and
// select_update_dev consumes it -- no host copy and no synchronisation per pass.
PBU_API int pbu_propagation_select_pass_dev(pbu_propagation *p, int pass, uint64_t *hist_dev) {
    if (!p || p->T < 1 || pass < 0 || pass > 7 || !hist_dev) return fail(PBU_ERR_INVALID, "bad argument");
    CUDA_TRY(cudaSetDevice(p->e->device));
    return select_pass_into(p, pass, (unsigned long long *)hist_dev);
}

PBU_API int pbu_propagation_select_update_dev(pbu_propagation *p, int pass, const uint64_t *hist_total_dev) {
    if (!p || p->T < 1 || pass < 0 || pass > 7 || !hist_total_dev) return fail(PBU_ERR_INVALID, "bad argument");
    pbu_engine *e = p->e;
    CUDA_TRY(cudaSetDevice(e->device));
    const int items = p->N * p->T;
    select_update_kernel<<<(items * 32 + 127) / 128, 128, 0, e->stream>>>(p->N, p->T, pass, p->d_prefix, p->d_rank,
                                                                          (const unsigned long long *)hist_total_dev);
    CUDA_TRY(cudaGetLastError());
    return PBU_OK;
}

// the order statistics after pass 7: values host [N][n_ranks]
PBU_API int pbu_propagation_select_result(pbu_propagation *p, double *values_host) {
    if (!p || p->T < 1 || !values_host) return fail(PBU_ERR_INVALID, "bad argument");
    pbu_engine *e = p->e;
    CUDA_TRY(cudaSetDevice(e->device));
    std::vector<unsigned long long> k((size_t)p->N * p->T);
    CUDA_TRY(cudaMemcpyAsync(k.data(), p->d_prefix, k.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    for (size_t i = 0; i < k.size(); ++i) values_host[i] = key_to_double(k[i]);
    return PBU_OK;
}

// ---- bracketed select (see the kernels above).  n_q = 1 or 2 percentiles, fractions[q] in [0, 1].
// Step 1: brackets from this engine's subsample.  brackets_host [N][n_q][2] = (a, b); -inf / +inf where the safety margin
// reaches the ends of the subsample.  The caller min-reduces a and max-reduces b over GPUs.
static int shared_key_bytes(double lo, double hi) {
    const unsigned long long x = order_key_host(lo) ^ order_key_host(hi);
    for (int b = 0; b < 8; ++b)
        if ((x >> (56 - 8 * b)) & 255ull) return b;
    return 8;
}

PBU_API int pbu_propagation_bracket_sample(pbu_propagation *p, const double *fractions, int n_q, int stride, double *brackets_host) {
    if (!p || !fractions || n_q < 1 || 2 * n_q > MAX_T || stride < 1 || !brackets_host) return fail(PBU_ERR_INVALID, "bad argument");
    if (!p->evaluated) return fail(PBU_ERR_INVALID, "propagation has not been evaluated");
    pbu_engine *e = p->e;
    CUDA_TRY(cudaSetDevice(e->device));
    const int64_t m = bracket_sub_len(p->S, stride);
    if (m < 1024) return fail(PBU_ERR_INVALID, "subsample of %lld values is too small for a bracket: use the full radix select", (long long)m);
    if (p->sub_m != m) {
        cudaFree(p->d_sub);
        p->d_sub = nullptr;
        p->sub_m = 0;
        CUDA_TRY(cudaMalloc(&p->d_sub, (size_t)p->N * m * sizeof(double)));
        p->sub_m = m;
    }
    subsample_gather_kernel<<<dim3((unsigned)((m + 255) / 256), p->N), 256, 0, e->stream>>>(p->d_f, p->S, stride, m, p->d_sub);
    CUDA_TRY(cudaGetLastError());
    // ranks in the subsample: the percentile's position -/+ 6 standard deviations of a binomial count (+ slack)
    const int T = 2 * n_q;
    std::vector<long long> rk((size_t)p->N * T);
    std::vector<int> open_lo(n_q), open_hi(n_q);
    for (int q = 0; q < n_q; ++q) {
        const double f = std::min(1.0, std::max(0.0, fractions[q]));
        const double k = f * (double)(m - 1), dl = 6.0 * std::sqrt((double)m * f * (1.0 - f)) + 8.0;
        long long lo = (long long)std::floor(k - dl), hi = (long long)std::ceil(k + dl) + 1;
        open_lo[q] = lo < 0;
        open_hi[q] = hi > m - 1;
        lo = std::max<long long>(lo, 0);
        hi = std::min<long long>(hi, m - 1);
        for (int r = 0; r < p->N; ++r) {
            rk[(size_t)r * T + 2 * q] = lo;
            rk[(size_t)r * T + 2 * q + 1] = hi;
        }
    }
    p->T = T;
    p->cand_mode = false;
    CUDA_TRY(cudaMemcpyAsync(p->d_rank, rk.data(), rk.size() * sizeof(long long), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaMemsetAsync(p->d_prefix, 0, (size_t)p->N * T * sizeof(unsigned long long), e->stream));
    // all eight key bytes: design points whose values agree to many digits (an ODE output before the reaction starts:
    // 1 - 1e-11) need the low bits, a bracket from the leading bytes alone would hold every value
    for (int pass = 0; pass < 8; ++pass) {
        CUDA_TRY(cudaMemsetAsync(p->d_hist, 0, (size_t)p->N * T * 256 * sizeof(unsigned long long), e->stream));
        select_hist_kernel<<<dim3(p->N, NSPLIT), 256, 0, e->stream>>>(p->d_sub, m, T, pass, p->d_prefix, p->d_hist);
        select_update_kernel<<<(p->N * T * 32 + 127) / 128, 128, 0, e->stream>>>(p->N, T, pass, p->d_prefix, p->d_rank, p->d_hist);
    }
    CUDA_TRY(cudaGetLastError());
    std::vector<unsigned long long> k((size_t)p->N * T);
    CUDA_TRY(cudaMemcpyAsync(k.data(), p->d_prefix, k.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    for (int r = 0; r < p->N; ++r)
        for (int q = 0; q < n_q; ++q) {
            brackets_host[((size_t)r * n_q + q) * 2] = open_lo[q] ? -INFINITY : key_to_double(k[(size_t)r * T + 2 * q]);
            brackets_host[((size_t)r * n_q + q) * 2 + 1] = open_hi[q] ? INFINITY : key_to_double(k[(size_t)r * T + 2 * q + 1]);
        }
    p->T = 0;
    return PBU_OK;
}

// Step 2: the one pass over the evaluations with the (global) brackets.  Per design point: sum, min, max as
// pbu_propagation_moments; below / inside host [N][n_q]: this engine's counts of values < a and in [a, b].  *overflow != 0:
// more candidates than the buffer holds (at least 4 cap_hint + 4096 per point and percentile; cap_hint = the expected
// number) -- the caller falls back to the full radix select.
PBU_API int pbu_propagation_bracket_pass(pbu_propagation *p, const double *brackets_host, int n_q, int64_t cap_hint, double *sum_host, double *min_host,
                                         double *max_host, int64_t *below_host, int64_t *inside_host, int *overflow) {
    if (!p || !brackets_host || n_q < 1 || 2 * n_q > MAX_T || cap_hint < 0 || !below_host || !inside_host || !overflow)
        return fail(PBU_ERR_INVALID, "bad argument");
    if (!p->evaluated) return fail(PBU_ERR_INVALID, "propagation has not been evaluated");
    pbu_engine *e = p->e;
    CUDA_TRY(cudaSetDevice(e->device));
    const size_t nrq = (size_t)p->N * n_q;
    if (bracket_cap(cap_hint) > p->cap) {          // buffers are for two percentiles whatever n_q (pbu_propagation_create took them for large S)
        const int64_t want = bracket_cap(cap_hint);
        CUDA_TRY(cudaStreamSynchronize(e->stream));
        cudaFree(p->d_cand);
        p->d_cand = nullptr;
        p->cap = 0;
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        if ((size_t)p->N * 2 * (size_t)want * 8 + (64u << 20) > free_b)
            return fail(PBU_ERR_INVALID, "no room for %.2f GB of bracket candidates", (double)p->N * 2 * (double)want * 8 / 1e9);
        CUDA_TRY(cudaMalloc(&p->d_cand, (size_t)p->N * 2 * (size_t)want * sizeof(double)));
        p->cap = want;
    }
    const int64_t cap = p->cap;
    p->nq = n_q;
    CUDA_TRY(cudaMemcpyAsync(p->d_brk, brackets_host, nrq * 2 * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaMemsetAsync(p->d_cnt, 0, 2 * nrq * sizeof(unsigned long long), e->stream));
    if (n_q == 1)
        bracket_pass_kernel<1><<<dim3(p->N, NSPLIT), 256, 0, e->stream>>>(p->d_f, p->S, p->d_brk, p->d_part, p->d_cnt, p->d_cnt + nrq, p->d_cand, cap);
    else
        bracket_pass_kernel<2><<<dim3(p->N, NSPLIT), 256, 0, e->stream>>>(p->d_f, p->S, p->d_brk, p->d_part, p->d_cnt, p->d_cnt + nrq, p->d_cand, cap);
    CUDA_TRY(cudaGetLastError());
    std::vector<double> part((size_t)p->N * NSPLIT * 3);
    std::vector<unsigned long long> cnt(2 * nrq);
    CUDA_TRY(cudaMemcpyAsync(part.data(), p->d_part, part.size() * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaMemcpyAsync(cnt.data(), p->d_cnt, cnt.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    for (int k = 0; k < p->N; ++k) {
        double s = 0.0, mn = INFINITY, mx = -INFINITY;
        for (int j = 0; j < NSPLIT; ++j) {
            const double *o = &part[((size_t)k * NSPLIT + j) * 3];
            s += o[0];
            mn = std::min(mn, o[1]);
            mx = std::max(mx, o[2]);
        }
        if (sum_host) sum_host[k] = s;
        if (min_host) min_host[k] = mn;
        if (max_host) max_host[k] = mx;
    }
    *overflow = 0;
    for (size_t i = 0; i < nrq; ++i) {
        below_host[i] = (int64_t)cnt[i];
        inside_host[i] = (int64_t)cnt[nrq + i];
        if (cnt[nrq + i] > (unsigned long long)cap && brackets_host[2 * i] < brackets_host[2 * i + 1]) *overflow = 1;
    }
    return PBU_OK;
}

// Step 3: radix select among the candidates.  ranks_in_bracket host [N][2 n_q]: 0-based ranks among ALL GPUs'
// candidates of that design point and percentile (global rank - global count below a).  brackets_host: the global
// brackets of step 2 (their shared leading key bytes are skipped).  Then pbu_propagation_select_pass / _update /
// _result as for the full select, passes *first_pass .. 7, histograms [N][2 n_q][256].
PBU_API int pbu_propagation_bracket_select_begin(pbu_propagation *p, const int64_t *ranks_in_bracket, const double *brackets_host, int *first_pass) {
    if (!p || !ranks_in_bracket || !brackets_host || p->nq < 1 || !p->d_cand) return fail(PBU_ERR_INVALID, "no bracket pass has been run");
    pbu_engine *e = p->e;
    CUDA_TRY(cudaSetDevice(e->device));
    const int T = 2 * p->nq;
    int skip = 8;
    for (size_t i = 0; i < (size_t)p->N * p->nq; ++i) {
        const double a = brackets_host[2 * i], b = brackets_host[2 * i + 1];
        // a zero bound: -0.0 and +0.0 compare equal but their keys differ in the first byte
        skip = std::min(skip, (!(a <= b) || a == 0.0 || b == 0.0) ? 0 : shared_key_bytes(a, b));
    }
    if (skip > 7) skip = 7;
    const unsigned long long keep = (skip == 0) ? 0ull : (~0ull << (64 - 8 * skip));
    std::vector<unsigned long long> pre((size_t)p->N * T);
    std::vector<long long> r((size_t)p->N * T);
    for (int k = 0; k < p->N; ++k)
        for (int t = 0; t < T; ++t) {
            pre[(size_t)k * T + t] = order_key_host(brackets_host[((size_t)k * p->nq + t / 2) * 2]) & keep;
            r[(size_t)k * T + t] = ranks_in_bracket[(size_t)k * T + t];
        }
    p->T = T;
    p->cand_mode = true;
    if (first_pass) *first_pass = skip;
    CUDA_TRY(cudaMemcpyAsync(p->d_rank, r.data(), r.size() * sizeof(long long), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaMemcpyAsync(p->d_prefix, pre.data(), pre.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return PBU_OK;
}

// copy the evaluations of samples [s0, s0 + n) to the host, sample-major [n][N] (fun_eval, solve_problem.py:571)
PBU_API int pbu_propagation_get_evals(pbu_propagation *p, int64_t s0, int64_t n, double *f_host) {
    if (!p || !p->evaluated || s0 < 0 || n < 1 || s0 + n > p->S || !f_host) return fail(PBU_ERR_INVALID, "bad argument");
    pbu_engine *e = p->e;
    CUDA_TRY(cudaSetDevice(e->device));
    std::vector<double> pm((size_t)p->N * n);
    CUDA_TRY(cudaMemcpy2DAsync(pm.data(), (size_t)n * sizeof(double), p->d_f + s0, (size_t)p->S * sizeof(double), (size_t)n * sizeof(double), p->N,
                               cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    for (int64_t s = 0; s < n; ++s)
        for (int k = 0; k < p->N; ++k) f_host[(size_t)s * p->N + k] = pm[(size_t)k * n + s];
    return PBU_OK;
}

// the samples themselves (the rows handed to eval_host, or drawn on the device by eval_gaussian), [n][d] on the host
PBU_API int pbu_propagation_get_samples(pbu_propagation *p, int64_t s0, int64_t n, double *X_host) {
    if (!p || !p->evaluated || s0 < 0 || n < 1 || s0 + n > p->S || !X_host) return fail(PBU_ERR_INVALID, "bad argument");
    pbu_engine *e = p->e;
    CUDA_TRY(cudaSetDevice(e->device));
    const int d = e->d;
    std::vector<double> pm((size_t)d * n);
    CUDA_TRY(cudaMemcpy2DAsync(pm.data(), (size_t)n * sizeof(double), p->d_X + s0, (size_t)p->S * sizeof(double), (size_t)n * sizeof(double), d,
                               cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    for (int64_t s = 0; s < n; ++s)
        for (int i = 0; i < d; ++i) X_host[(size_t)s * d + i] = pm[(size_t)i * n + s];
    return PBU_OK;
}
