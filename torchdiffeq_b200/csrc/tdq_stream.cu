// tdq_stream.cu -- the HBM-bound kernels of the explicit RK hot path.
//
//   k_combine      y_i = y0 + sum_j k_j * c_ij                         rk_common.py:79, :85
//   k_reduce       sum ((err or x) / tol)^2 per norm segment            rk_common.py:89, misc.py:80-82, :22-23, :55-58, :69
//   k_fit_commit   y_mid, quartic coefficients, y0 <- y1, k0 <- k_S     rk_common.py:363-369, interp.py:1-22, rk_common.py:338-352
//   k_interp_eval  solution[j] = p((t_j - t0)/(t1 - t0))                interp.py:25-48
//
// All of them stream each operand exactly once with 128-bit transactions, keep stage slots as
// separate contiguous arrays (structure of arrays; the reference interleaves the stage index
// innermost, rk_common.py:69) and read their scalars from the device control block, so the same
// launch sequence is valid for every attempt and can be replayed from a CUDA graph.
// Arithmetic is contraction free and follows the reference's order of roundings.
#include "tdq_common.cuh"

namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------------
// Stage combine.  NK = number of non-zero tableau entries in the row (compile time => the NK+1 loads
// of a thread are all in flight before the first use).
// ------------------------------------------------------------------------------------------------
// Launch shape, from the r1 tuning sweep on B200 (profiles/r1_combine_variants.txt): a persistent
// grid-stride kernel beats one-tile-per-block by 1-15 % at 33.5 MB per operand because the last wave no
// longer drains alone.  Rows with few operands need more bytes in flight per thread:
//   NK <= 2 : 512 threads, 4 vectors per operand per thread, 2 blocks per SM
//   NK >= 3 : 256 threads, 2 vectors per operand per thread, 8 blocks per SM
// The bulk-async (TMA, cp.async.bulk + mbarrier ring through shared memory) variant measured 5-10 %
// SLOWER than plain 128-bit loads for this pure streaming pattern, so it is not used.
template <typename T, int NK, int THREADS, int U, bool VECTOR>
__global__ void __launch_bounds__(THREADS)
k_combine(const TdqCtrl *__restrict__ c, int row, T *__restrict__ out, const T *__restrict__ y0, KPtrs kp, size_t n) {
    if (c->halt) return;
    using A = Ar<T>;
    T cf[NK];
    const T *k[NK];
#pragma unroll
    for (int m = 0; m < NK; ++m) {
        cf[m] = (T)c->coef[row][m];
        k[m] = reinterpret_cast<const T *>(kp.p[m]);
    }
    if (VECTOR) {
        using V = Vec<T>;
        const size_t nvec = n / V::N;
        const size_t stride = (size_t)gridDim.x * (THREADS * U);
        for (size_t base = (size_t)blockIdx.x * (THREADS * U) + threadIdx.x; base < nvec; base += stride) {
            V a[U], kv[U][NK];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const size_t v = base + (size_t)u * THREADS;
                if (v < nvec) {
                    a[u] = ld_stream<T>(y0 + v * V::N);
#pragma unroll
                    for (int m = 0; m < NK; ++m) kv[u][m] = ld_stream<T>(k[m] + v * V::N);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const size_t v = base + (size_t)u * THREADS;
                if (v < nvec) {
                    V r;
#pragma unroll
                    for (int e = 0; e < V::N; ++e) {
                        T acc = A::mul(kv[u][0].v[e], cf[0]);
#pragma unroll
                        for (int m = 1; m < NK; ++m) acc = A::add(acc, A::mul(kv[u][m].v[e], cf[m]));
                        r.v[e] = A::add(a[u].v[e], acc);
                    }
                    st_vec<T>(out + v * V::N, r);
                }
            }
        }
        // scalar tail (n not a multiple of the vector width): first threads of block 0
        if (blockIdx.x == 0) {
            const size_t i = nvec * V::N + threadIdx.x;
            if (i < n) {
                T acc = A::mul(k[0][i], cf[0]);
#pragma unroll
                for (int m = 1; m < NK; ++m) acc = A::add(acc, A::mul(k[m][i], cf[m]));
                out[i] = A::add(y0[i], acc);
            }
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * THREADS) {
            T acc = A::mul(k[0][i], cf[0]);
#pragma unroll
            for (int m = 1; m < NK; ++m) acc = A::add(acc, A::mul(k[m][i], cf[m]));
            out[i] = A::add(y0[i], acc);
        }
    }
}

static int sm_count() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
    }
    return sms;
}

template <typename T, int NK>
int launch_combine(const TdqCtrl *c, int row, void *out, const void *y0, const KPtrs &kp, size_t n, bool vec,
                   cudaStream_t st) {
    constexpr bool kFew = NK <= 2;
    constexpr int THREADS = kFew ? 512 : 256;
    constexpr int U = kFew ? 4 : 2;
    constexpr int PER_SM = kFew ? 2 : 8;
    if (vec) {
        using V = Vec<T>;
        const size_t nvec = n / V::N;
        size_t blocks = (nvec + (size_t)THREADS * U - 1) / ((size_t)THREADS * U);
        const size_t cap = (size_t)sm_count() * PER_SM;       // one resident wave; the loop covers the rest
        if (blocks > cap) blocks = cap;
        if (blocks == 0) blocks = 1;
        k_combine<T, NK, THREADS, U, true><<<(unsigned)blocks, THREADS, 0, st>>>(c, row, (T *)out, (const T *)y0, kp, n);
    } else {
        size_t blocks = (n + THREADS - 1) / THREADS;
        const size_t cap = (size_t)sm_count() * PER_SM * 2;
        if (blocks > cap) blocks = cap;
        if (blocks == 0) blocks = 1;
        k_combine<T, NK, THREADS, U, false><<<(unsigned)blocks, THREADS, 0, st>>>(c, row, (T *)out, (const T *)y0, kp, n);
    }
    return 0;
}

template <typename T>
int dispatch_combine(int nk, const TdqCtrl *c, int row, void *out, const void *y0, const KPtrs &kp, size_t n,
                     bool vec, cudaStream_t st) {
    switch (nk) {
#define TDQ_CASE(N) case N: return launch_combine<T, N>(c, row, out, y0, kp, n, vec, st);
        TDQ_CASE(1) TDQ_CASE(2) TDQ_CASE(3) TDQ_CASE(4) TDQ_CASE(5) TDQ_CASE(6) TDQ_CASE(7) TDQ_CASE(8)
        TDQ_CASE(9) TDQ_CASE(10) TDQ_CASE(11) TDQ_CASE(12) TDQ_CASE(13) TDQ_CASE(14) TDQ_CASE(15)
        TDQ_CASE(16) TDQ_CASE(17)
#undef TDQ_CASE
    }
    return -1;
}

// ------------------------------------------------------------------------------------------------
// Segmented scaled sum of squares with a deterministic two-level reduction.
//   MODE 0: q = err / tol, err = sum_m k_m*ecoef_m, tol = atol + rtol*max(|y0|,|y1|)   (error ratio)
//   MODE 1: q = x / scale or (x - x2) / scale, scale = atol + |y0|*rtol               (initial step)
// grid = (tiles of the longest segment, n_seg).  Block (tx, s) reduces one tile of segment s into
// partials[s*tiles + tx]; the last block to finish (ticket) sums every segment's partials in index
// order, so the result does not depend on the schedule or the SM count.
// ------------------------------------------------------------------------------------------------
struct SegDesc {
    int64_t off[TDQ_MAX_SEGS];
    int64_t len[TDQ_MAX_SEGS];
};

template <typename T, bool VTOL> struct TolT { using type = T; };
template <typename T> struct TolT<T, true> { using type = double; };

template <typename T, int NK, int MODE, bool VECTOR, bool VTOL, bool WRITEQ>
__global__ void __launch_bounds__(kThreads)
k_reduce(const TdqCtrl *__restrict__ c, const T *__restrict__ y0, const T *__restrict__ y1, KPtrs kp,
         const double *__restrict__ rtol_v, const double *__restrict__ atol_v, SegDesc seg, int n_seg,
         int tiles, double *__restrict__ partials, unsigned int *ticket, double *__restrict__ out,
         void *__restrict__ q_out) {
    if (c->halt) return;
    using A = Ar<T>;
    using Q = typename TolT<T, VTOL>::type;     // dtype of tol and of err/tol (float64 with vector tolerances)
    __shared__ double red[kThreads / 32];
    __shared__ bool is_last;

    const int s = blockIdx.y;
    const int64_t off = seg.off[s], len = seg.len[s];
    constexpr int U = 4;
    constexpr int VN = VECTOR ? Vec<T>::N : 1;
    constexpr int64_t TILE = (int64_t)kThreads * U * VN;
    const int64_t tile_lo = (int64_t)blockIdx.x * TILE;

    double acc = 0.0, bad = 0.0;
    if (tile_lo < len) {
        T cf[NK > 0 ? NK : 1];
        const T *k[NK > 0 ? NK : 1];
#pragma unroll
        for (int m = 0; m < NK; ++m) {
            cf[m] = (MODE == 0) ? (T)c->ecoef[m] : (T)0;
            k[m] = reinterpret_cast<const T *>(kp.p[m]);
        }
        const T rtolT = (T)c->rtol, atolT = (T)c->atol;   // 0-dim float64 tensors act as scalars of T (misc.py:81)

        auto element = [&](int64_t i, T v0, T v1, const T *kv) {
            // i: global element index; v0 = y0[i]; v1 = y1[i] (MODE 0) ; kv = operands
            T num;
            if (MODE == 0) {
                num = A::mul(kv[0], cf[0]);
#pragma unroll
                for (int m = 1; m < NK; ++m) num = A::add(num, A::mul(kv[m], cf[m]));
            } else {
                num = (NK == 2) ? A::sub(kv[0], kv[1]) : kv[0];
            }
            Q q;
            if (VTOL) {
                const double rt = rtol_v[i], at = atol_v[i];
                double tol;
                if (MODE == 0) tol = at + rt * (double)A::max_nan(A::abs(v0), A::abs(v1));
                else tol = at + (double)A::abs(v0) * rt;
                q = (Q)((double)num / tol);
            } else {
                T tol;
                if (MODE == 0) tol = A::add(atolT, A::mul(rtolT, A::max_nan(A::abs(v0), A::abs(v1))));
                else tol = A::add(atolT, A::mul(A::abs(v0), rtolT));
                q = (Q)A::div(num, tol);
            }
            if (WRITEQ) reinterpret_cast<Q *>(q_out)[i] = q;
            if (MODE == 0 && !A::finite(v1)) bad += 1.0;
            const Q q2 = Ar<Q>::mul(q, q);                 // .abs().pow(2)
            acc += (double)q2;
        };

        if (VECTOR) {
            using V = Vec<T>;
            const int64_t nvec = len / V::N;               // full vectors in this segment
            const int64_t vbase = tile_lo / V::N + threadIdx.x;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t v = vbase + (int64_t)u * kThreads;
                if (v < nvec) {
                    const int64_t i0 = off + v * V::N;
                    V a0 = ld_stream<T>(y0 + i0), a1;
                    if (MODE == 0) a1 = ld_stream<T>(y1 + i0);
                    V kv[NK > 0 ? NK : 1];
#pragma unroll
                    for (int m = 0; m < NK; ++m) kv[m] = ld_stream<T>(k[m] + i0);
#pragma unroll
                    for (int e = 0; e < V::N; ++e) {
                        T ke[NK > 0 ? NK : 1];
#pragma unroll
                        for (int m = 0; m < NK; ++m) ke[m] = kv[m].v[e];
                        element(i0 + e, a0.v[e], (MODE == 0) ? a1.v[e] : (T)0, ke);
                    }
                }
            }
            // tail elements of the segment (len % VN), handled by the tile that would contain them
            const int64_t tail0 = nvec * V::N;
            if (tail0 < len && tail0 >= tile_lo && tail0 < tile_lo + TILE) {
                const int64_t i = tail0 + threadIdx.x;
                if (i < len) {
                    T ke[NK > 0 ? NK : 1];
#pragma unroll
                    for (int m = 0; m < NK; ++m) ke[m] = k[m][off + i];
                    element(off + i, y0[off + i], (MODE == 0) ? y1[off + i] : (T)0, ke);
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = tile_lo + (int64_t)u * kThreads + threadIdx.x;
                if (i < len) {
                    T ke[NK > 0 ? NK : 1];
#pragma unroll
                    for (int m = 0; m < NK; ++m) ke[m] = k[m][off + i];
                    element(off + i, y0[off + i], (MODE == 0) ? y1[off + i] : (T)0, ke);
                }
            }
        }
    }

    const double tsum = block_sum<kThreads>(acc, red);
    const double tbad = block_sum<kThreads>(bad, red);
    double *p_sum = partials + 2;                              // [n_seg][tiles]; partials[0..1] hold the ticket
    double *p_bad = p_sum + (size_t)n_seg * tiles;             // [n_seg][tiles]
    if (threadIdx.x == 0) {
        p_sum[(size_t)s * tiles + blockIdx.x] = tsum;
        p_bad[(size_t)s * tiles + blockIdx.x] = tbad;
        __threadfence();
        const unsigned int total = gridDim.x * gridDim.y;
        const unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == total - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // Final pass, fixed order: thread t adds partials t, t+256, ... then a fixed tree.
    double nb = 0.0;
    for (int sg = 0; sg < n_seg; ++sg) {
        double a = 0.0, b = 0.0;
        for (int i = threadIdx.x; i < tiles; i += kThreads) {
            a += __ldcg(&p_sum[(size_t)sg * tiles + i]);
            b += __ldcg(&p_bad[(size_t)sg * tiles + i]);
        }
        const double sa = block_sum<kThreads>(a, red);
        const double sb = block_sum<kThreads>(b, red);
        if (threadIdx.x == 0) out[sg] = sa;
        nb += sb;                                           // only thread 0's value is meaningful
    }
    if (threadIdx.x == 0) {
        out[n_seg] = nb;
        *ticket = 0;                                        // self-reset for the next launch
    }
}

template <typename T, int NK, int MODE>
int launch_reduce(const TdqCtrl *c, const void *y0, const void *y1, const KPtrs &kp, const double *rtol_v,
                  const double *atol_v, const SegDesc &seg, int n_seg, int64_t max_len, bool vec,
                  double *partials, double *out, void *q_out, cudaStream_t st, int *tiles_out) {
    const bool vtol = rtol_v != nullptr;
    const bool wq = q_out != nullptr;
    const int vn = vec ? Vec<T>::N : 1;
    const int64_t TILE = (int64_t)kThreads * 4 * vn;
    int64_t tiles = (max_len + TILE - 1) / TILE;
    if (tiles < 1) tiles = 1;
    if (tiles_out) *tiles_out = (int)tiles;
    // the ticket word sits in front of the two partial planes (zeroed once by the caller, self-resetting)
    unsigned int *ticket = reinterpret_cast<unsigned int *>(partials);
    dim3 grid((unsigned)tiles, (unsigned)n_seg);
#define TDQ_LAUNCH(V_, VT_, WQ_)                                                                         \
    k_reduce<T, NK, MODE, V_, VT_, WQ_><<<grid, kThreads, 0, st>>>(                                      \
        c, (const T *)y0, (const T *)y1, kp, rtol_v, atol_v, seg, n_seg, (int)tiles, partials, ticket, out, q_out)
    if (vec) {
        if (vtol) { if (wq) TDQ_LAUNCH(true, true, true); else TDQ_LAUNCH(true, true, false); }
        else      { if (wq) TDQ_LAUNCH(true, false, true); else TDQ_LAUNCH(true, false, false); }
    } else {
        if (vtol) { if (wq) TDQ_LAUNCH(false, true, true); else TDQ_LAUNCH(false, true, false); }
        else      { if (wq) TDQ_LAUNCH(false, false, true); else TDQ_LAUNCH(false, false, false); }
    }
#undef TDQ_LAUNCH
    return 0;
}

template <typename T>
int dispatch_reduce_err(int nk, const TdqCtrl *c, const void *y0, const void *y1, const KPtrs &kp,
                        const double *rtol_v, const double *atol_v, const SegDesc &seg, int n_seg,
                        int64_t max_len, bool vec, double *partials, double *out, void *q_out, cudaStream_t st) {
    switch (nk) {
#define TDQ_CASE(N) case N: return launch_reduce<T, N, 0>(c, y0, y1, kp, rtol_v, atol_v, seg, n_seg, max_len, vec, partials, out, q_out, st, nullptr);
        TDQ_CASE(1) TDQ_CASE(2) TDQ_CASE(3) TDQ_CASE(4) TDQ_CASE(5) TDQ_CASE(6) TDQ_CASE(7) TDQ_CASE(8)
        TDQ_CASE(9) TDQ_CASE(10) TDQ_CASE(11) TDQ_CASE(12) TDQ_CASE(13) TDQ_CASE(14) TDQ_CASE(15)
        TDQ_CASE(16) TDQ_CASE(17)
#undef TDQ_CASE
    }
    return -1;
}

// ------------------------------------------------------------------------------------------------
// Interpolant fit + state commit, accepted attempts only.
//   y_mid = y0 + sum_m k_m * fl(dt*mid_m)                     rk_common.py:366
//   a,b,c,d,e per interp.py:17-22 with f0 = s*k_0, f1 = s*k_S (s = t_sign, folded into sdt)
//   y0 <- y1 ; k_0 <- k_S                                     rk_common.py:341, :352 (FSAL carry)
// ------------------------------------------------------------------------------------------------
template <typename T, int NK, bool VECTOR>
__global__ void __launch_bounds__(kThreads)
k_fit_commit(const TdqCtrl *__restrict__ c, T *y0p, const T *__restrict__ y1p, KPtrsMut kfirst_last, KPtrs kmid,
             T *__restrict__ ce, T *__restrict__ cd, T *__restrict__ cc, T *__restrict__ cb, T *__restrict__ ca,
             size_t n) {
    if (c->halt && !c->done) return;       // a failed attempt never commits
    if (!c->accept) return;
    if (c->seq == 0) return;
    using A = Ar<T>;
    T mf[NK];
    const T *km[NK];
#pragma unroll
    for (int m = 0; m < NK; ++m) {
        mf[m] = (T)c->fit_mcoef[m];
        km[m] = reinterpret_cast<const T *>(kmid.p[m]);
    }
    T *k0 = reinterpret_cast<T *>(kfirst_last.p[0]);
    const T *kS = reinterpret_cast<const T *>(kfirst_last.p[1]);
    const T sdt = (T)c->fit_sdt;
    const T two_sdt = A::mul((T)2, sdt);                    // 2 * dt (exact)

    auto fit = [&](T y0v, T y1v, T f0, T f1, const T *kv, T &e, T &d, T &cq, T &b, T &a) {
        T acc = A::mul(kv[0], mf[0]);
#pragma unroll
        for (int m = 1; m < NK; ++m) acc = A::add(acc, A::mul(kv[m], mf[m]));
        const T ymid = A::add(y0v, acc);
        // a = 2*dt*(f1 - f0) - 8*(y1 + y0) + 16*y_mid
        a = A::add(A::sub(A::mul(two_sdt, A::sub(f1, f0)), A::mul((T)8, A::add(y1v, y0v))), A::mul((T)16, ymid));
        // b = dt*(5*f0 - 3*f1) + 18*y0 + 14*y1 - 32*y_mid
        b = A::sub(A::add(A::add(A::mul(sdt, A::sub(A::mul((T)5, f0), A::mul((T)3, f1))), A::mul((T)18, y0v)),
                          A::mul((T)14, y1v)),
                   A::mul((T)32, ymid));
        // c = dt*(f1 - 4*f0) - 11*y0 - 5*y1 + 16*y_mid
        cq = A::add(A::sub(A::sub(A::mul(sdt, A::sub(f1, A::mul((T)4, f0))), A::mul((T)11, y0v)), A::mul((T)5, y1v)),
                    A::mul((T)16, ymid));
        d = A::mul(sdt, f0);
        e = y0v;
    };

    if (VECTOR) {
        using V = Vec<T>;
        const size_t nvec = n / V::N;
        for (size_t v = (size_t)blockIdx.x * kThreads + threadIdx.x; v < nvec; v += (size_t)gridDim.x * kThreads) {
            const size_t i0 = v * V::N;
            V a0 = ld_stream<T>(y0p + i0), a1 = ld_stream<T>(y1p + i0);
            V f0 = ld_stream<T>(k0 + i0), f1 = ld_stream<T>(kS + i0);
            V kv[NK];
#pragma unroll
            for (int m = 0; m < NK; ++m) kv[m] = ld_stream<T>(km[m] + i0);
            V re, rd, rc, rb, ra;
#pragma unroll
            for (int e = 0; e < V::N; ++e) {
                T ke[NK];
#pragma unroll
                for (int m = 0; m < NK; ++m) ke[m] = kv[m].v[e];
                fit(a0.v[e], a1.v[e], f0.v[e], f1.v[e], ke, re.v[e], rd.v[e], rc.v[e], rb.v[e], ra.v[e]);
            }
            st_vec<T>(ce + i0, re); st_vec<T>(cd + i0, rd); st_vec<T>(cc + i0, rc);
            st_vec<T>(cb + i0, rb); st_vec<T>(ca + i0, ra);
            st_vec<T>(y0p + i0, a1);
            st_vec<T>(k0 + i0, f1);
        }
        if (blockIdx.x == gridDim.x - 1) {
            const size_t i = nvec * V::N + threadIdx.x;
            if (i < n) {
                T ke[NK];
#pragma unroll
                for (int m = 0; m < NK; ++m) ke[m] = km[m][i];
                const T f0 = k0[i], f1 = kS[i], a0 = y0p[i], a1 = y1p[i];
                T e, d, cq, b, a;
                fit(a0, a1, f0, f1, ke, e, d, cq, b, a);
                ce[i] = e; cd[i] = d; cc[i] = cq; cb[i] = b; ca[i] = a;
                y0p[i] = a1; k0[i] = f1;
            }
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads) {
            T ke[NK];
#pragma unroll
            for (int m = 0; m < NK; ++m) ke[m] = km[m][i];
            const T f0 = k0[i], f1 = kS[i], a0 = y0p[i], a1 = y1p[i];
            T e, d, cq, b, a;
            fit(a0, a1, f0, f1, ke, e, d, cq, b, a);
            ce[i] = e; cd[i] = d; cc[i] = cq; cb[i] = b; ca[i] = a;
            y0p[i] = a1; k0[i] = f1;
        }
    }
}

template <typename T, int NK>
int launch_fit(const TdqCtrl *c, void *y0, const void *y1, const KPtrsMut &fl, const KPtrs &kmid, void *const *coeff,
               size_t n, bool vec, cudaStream_t st) {
    if (vec) {
        const size_t nvec = n / Vec<T>::N;
        size_t blocks = (nvec + kThreads - 1) / kThreads;
        const size_t cap = (size_t)sm_count() * 8;           // persistent: a rejected attempt's no-op launch stays cheap
        if (blocks > cap) blocks = cap;
        if (blocks == 0) blocks = 1;
        k_fit_commit<T, NK, true><<<(unsigned)blocks, kThreads, 0, st>>>(
            c, (T *)y0, (const T *)y1, fl, kmid, (T *)coeff[0], (T *)coeff[1], (T *)coeff[2], (T *)coeff[3],
            (T *)coeff[4], n);
    } else {
        size_t blocks = (n + kThreads - 1) / kThreads;
        if (blocks == 0) blocks = 1;
        if (blocks > 148 * 16) blocks = 148 * 16;
        k_fit_commit<T, NK, false><<<(unsigned)blocks, kThreads, 0, st>>>(
            c, (T *)y0, (const T *)y1, fl, kmid, (T *)coeff[0], (T *)coeff[1], (T *)coeff[2], (T *)coeff[3],
            (T *)coeff[4], n);
    }
    return 0;
}

template <typename T>
int dispatch_fit(int nk, const TdqCtrl *c, void *y0, const void *y1, const KPtrsMut &fl, const KPtrs &kmid,
                 void *const *coeff, size_t n, bool vec, cudaStream_t st) {
    switch (nk) {
#define TDQ_CASE(N) case N: return launch_fit<T, N>(c, y0, y1, fl, kmid, coeff, n, vec, st);
        TDQ_CASE(1) TDQ_CASE(2) TDQ_CASE(3) TDQ_CASE(4) TDQ_CASE(5) TDQ_CASE(6) TDQ_CASE(7) TDQ_CASE(8)
        TDQ_CASE(9) TDQ_CASE(10) TDQ_CASE(11) TDQ_CASE(12) TDQ_CASE(13) TDQ_CASE(14) TDQ_CASE(15)
        TDQ_CASE(16) TDQ_CASE(17)
#undef TDQ_CASE
    }
    return -1;
}

// ------------------------------------------------------------------------------------------------
// Dense output.  x = T((t - t0)/(t1 - t0)) in float64 then cast (interp.py:39-40); running powers,
// not Horner (interp.py:42-46).  One pass reads the five coefficients and writes every pending
// output row.
// ------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T eval_poly(T e, T d, T cq, T b, T a, T x) {
    using A = Ar<T>;
    T total = A::add(e, A::mul(x, d));
    T xp = A::mul(x, x);
    total = A::add(total, A::mul(xp, cq));
    xp = A::mul(xp, x);
    total = A::add(total, A::mul(xp, b));
    xp = A::mul(xp, x);
    total = A::add(total, A::mul(xp, a));
    return total;
}

// p(x) for a caller-supplied abscissa x (float64, cast to T like interp.py:39-40); used by dense-output
// closures that keep their own (t0, t1, coefficients) per accepted step (odeint.py:111-157).
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_poly_eval(const T *__restrict__ ce, const T *__restrict__ cd, const T *__restrict__ cc, const T *__restrict__ cb,
            const T *__restrict__ ca, T *__restrict__ out, double x64, size_t n) {
    const T x = (T)x64;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads)
        out[i] = eval_poly<T>(ce[i], cd[i], cc[i], cb[i], ca[i], x);
}

template <typename T, bool VECTOR, bool AT>
__global__ void __launch_bounds__(kThreads)
k_interp_eval(const TdqCtrl *__restrict__ c, const T *__restrict__ ce, const T *__restrict__ cd,
              const T *__restrict__ cc, const T *__restrict__ cb, const T *__restrict__ ca,
              T *__restrict__ solution, const double *__restrict__ t_at, size_t n) {
    int lo, hi;
    if (AT) { lo = 0; hi = 1; }
    else {
        if (c->halt && !c->done) return;
        if (!c->accept || c->seq == 0) return;
        lo = c->emit_lo; hi = c->emit_hi;
        if (lo >= hi) return;
    }
    const double t0 = c->t0, t1 = c->t1;
    auto xof = [&](int j) -> T {
        const double t = AT ? *t_at : c->t_out[j];
        return (T)((t - t0) / (t1 - t0));
    };
    if (VECTOR) {
        using V = Vec<T>;
        const size_t nvec = n / V::N;
        for (size_t v = (size_t)blockIdx.x * kThreads + threadIdx.x; v < nvec; v += (size_t)gridDim.x * kThreads) {
            const size_t i0 = v * V::N;
            const V e = ld_stream<T>(ce + i0), d = ld_stream<T>(cd + i0), q = ld_stream<T>(cc + i0),
                    b = ld_stream<T>(cb + i0), a = ld_stream<T>(ca + i0);
            for (int j = lo; j < hi; ++j) {
                const T x = xof(j);
                V r;
#pragma unroll
                for (int l = 0; l < V::N; ++l) r.v[l] = eval_poly<T>(e.v[l], d.v[l], q.v[l], b.v[l], a.v[l], x);
                st_vec<T>(solution + (AT ? 0 : (size_t)j * n) + i0, r);
            }
        }
        if (blockIdx.x == gridDim.x - 1) {
            const size_t i = nvec * V::N + threadIdx.x;
            if (i < n)
                for (int j = lo; j < hi; ++j)
                    solution[(AT ? 0 : (size_t)j * n) + i] = eval_poly<T>(ce[i], cd[i], cc[i], cb[i], ca[i], xof(j));
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads)
            for (int j = lo; j < hi; ++j)
                solution[(AT ? 0 : (size_t)j * n) + i] = eval_poly<T>(ce[i], cd[i], cc[i], cb[i], ca[i], xof(j));
    }
}

template <typename T, bool AT>
int launch_eval(const TdqCtrl *c, const void *const *coeff, void *solution, const double *t_at, size_t n, bool vec,
                cudaStream_t st) {
    if (vec) {
        const size_t nvec = n / Vec<T>::N;
        size_t blocks = (nvec + kThreads - 1) / kThreads;
        const size_t cap = (size_t)sm_count() * 8;           // most launches are no-ops (no output time in the step)
        if (blocks > cap) blocks = cap;
        if (blocks == 0) blocks = 1;
        k_interp_eval<T, true, AT><<<(unsigned)blocks, kThreads, 0, st>>>(
            c, (const T *)coeff[0], (const T *)coeff[1], (const T *)coeff[2], (const T *)coeff[3],
            (const T *)coeff[4], (T *)solution, t_at, n);
    } else {
        size_t blocks = (n + kThreads - 1) / kThreads;
        if (blocks == 0) blocks = 1;
        if (blocks > 148 * 16) blocks = 148 * 16;
        k_interp_eval<T, false, AT><<<(unsigned)blocks, kThreads, 0, st>>>(
            c, (const T *)coeff[0], (const T *)coeff[1], (const T *)coeff[2], (const T *)coeff[3],
            (const T *)coeff[4], (T *)solution, t_at, n);
    }
    return 0;
}

// y_probe = y0 + h0*f0 with f0 = s*k0 (misc.py:66)
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_probe(const TdqCtrl *__restrict__ c, T *__restrict__ out, const T *__restrict__ y0, const T *__restrict__ f0, size_t n) {
    using A = Ar<T>;
    const T h = A::mul((T)c->t_sign, (T)c->h0);
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads)
        out[i] = A::add(y0[i], A::mul(h, f0[i]));
}

// Host mirror of the tableau sparsity (so launchers can pick NK without reading device memory).
struct HostShape {
    int valid;
    int n_stages, fsal;
    int row_nnz[TDQ_ROWS];
    int row_idx[TDQ_ROWS][TDQ_MAX_K];
    int err_nnz, mid_nnz;
    int err_idx[TDQ_MAX_K], mid_idx[TDQ_MAX_K];
};

}  // namespace

// Tableau sparsity is a function of the tableau only; launchers recompute it from the tdq_tableau the
// caller passes (cheap) instead of caching per control block.
static void shape_from_tableau(const tdq_tableau *tab, HostShape *h) {
    memset(h, 0, sizeof(*h));
    const int S = tab->n_stages;
    h->n_stages = S;
    h->fsal = tab->fsal;
    for (int i = 0; i < S; ++i) {
        int m = 0;
        for (int j = 0; j <= i; ++j)
            if (tab->beta[i][j] != 0.0) h->row_idx[i][m++] = j;
        h->row_nnz[i] = m;
    }
    int m = 0;
    for (int j = 0; j <= S; ++j)
        if (tab->c_sol[j] != 0.0) h->row_idx[S][m++] = j;
    h->row_nnz[S] = m;
    m = 0;
    for (int j = 0; j <= S; ++j)
        if (tab->c_err[j] != 0.0) h->err_idx[m++] = j;
    h->err_nnz = m;
    m = 0;
    for (int j = 0; j <= S; ++j)
        if (tab->c_mid[j] != 0.0) h->mid_idx[m++] = j;
    h->mid_nnz = m;
    h->valid = 1;
}

#define TDQ_DISPATCH_T(dtype, ...)                                         \
    do {                                                                   \
        if ((dtype) == TDQ_F32) { using T = float; __VA_ARGS__; }          \
        else if ((dtype) == TDQ_F64) { using T = double; __VA_ARGS__; }    \
        else { tdq_set_error("unsupported dtype %d", (int)(dtype)); return TDQ_ERR_INVALID; } \
    } while (0)

extern "C" {

size_t tdq_norm_partials_len(size_t n_max_seg_len, int32_t n_seg) {
    // tiles are sized for the scalar path (smallest tile) so the buffer fits either path
    const size_t tile = (size_t)kThreads * 4;
    size_t tiles = (n_max_seg_len + tile - 1) / tile;
    if (tiles < 1) tiles = 1;
    return 2 * (size_t)n_seg * tiles + 2;   // two planes + ticket word (8 bytes)
}

int tdq_stage_combine(void *ctrl_dev, const tdq_tableau *tab, int32_t dtype, int32_t row, void *y_out,
                      const void *y0, const void *const *k, size_t n, void *stream) {
    TDQ_REQUIRE(ctrl_dev && tab && y_out && y0 && k, "null argument");
    HostShape hs;
    shape_from_tableau(tab, &hs);
    TDQ_REQUIRE(row >= 0 && row <= hs.n_stages, "row out of range");
    const int nk = hs.row_nnz[row];
    TDQ_REQUIRE(nk >= 1, "empty tableau row");
    KPtrs kp;
    memset(&kp, 0, sizeof(kp));
    bool vec = tdq_aligned16(y_out) && tdq_aligned16(y0);
    for (int m = 0; m < nk; ++m) {
        kp.p[m] = k[hs.row_idx[row][m]];
        TDQ_REQUIRE(kp.p[m] != nullptr, "missing stage slot for a non-zero tableau entry");
        vec = vec && tdq_aligned16(kp.p[m]);
    }
    if (n == 0) return TDQ_OK;
    int rc = -1;
    TDQ_DISPATCH_T(dtype, rc = dispatch_combine<T>(nk, (const TdqCtrl *)ctrl_dev, row, y_out, y0, kp, n, vec,
                                                   (cudaStream_t)stream));
    TDQ_REQUIRE(rc == 0, "unsupported number of stage terms");
    TDQ_CHECK_CUDA(cudaGetLastError());
    return TDQ_OK;
}

static int fill_segments(const int64_t *seg_offsets, const int64_t *seg_lens, int32_t n_seg, size_t n,
                         SegDesc *sd, int64_t *max_len, bool *aligned, size_t elem) {
    *max_len = 0;
    *aligned = true;
    for (int s = 0; s < n_seg; ++s) {
        sd->off[s] = seg_offsets ? seg_offsets[s] : 0;
        sd->len[s] = seg_lens ? seg_lens[s] : (int64_t)n;
        if (sd->off[s] < 0 || sd->len[s] < 0 || (size_t)(sd->off[s] + sd->len[s]) > n) return -1;
        if (sd->len[s] > *max_len) *max_len = sd->len[s];
        if ((sd->off[s] * (int64_t)elem) % 16 != 0) *aligned = false;
    }
    return 0;
}

int tdq_error_norm(void *ctrl_dev, const tdq_tableau *tab, int32_t dtype, const void *y0, const void *y1,
                   const void *const *k, const double *rtol_vec, const double *atol_vec,
                   const int64_t *seg_offsets, const int64_t *seg_lens, int32_t n_seg, size_t n,
                   double *partials, double *out, void *err_over_tol_out, void *stream) {
    TDQ_REQUIRE(ctrl_dev && tab && y0 && y1 && k && partials && out, "null argument");
    TDQ_REQUIRE(n_seg >= 1 && n_seg <= TDQ_MAX_SEGS, "n_seg out of range");
    TDQ_REQUIRE((rtol_vec == nullptr) == (atol_vec == nullptr), "rtol_vec and atol_vec go together");
    HostShape hs;
    shape_from_tableau(tab, &hs);
    const int nk = hs.err_nnz;
    TDQ_REQUIRE(nk >= 1, "tableau has no error weights");
    KPtrs kp;
    memset(&kp, 0, sizeof(kp));
    bool vec = tdq_aligned16(y0) && tdq_aligned16(y1);
    for (int m = 0; m < nk; ++m) {
        kp.p[m] = k[hs.err_idx[m]];
        TDQ_REQUIRE(kp.p[m] != nullptr, "missing stage slot for a non-zero error weight");
        vec = vec && tdq_aligned16(kp.p[m]);
    }
    SegDesc sd;
    int64_t max_len;
    bool seg_aligned;
    const size_t elem = dtype == TDQ_F32 ? 4 : 8;
    TDQ_REQUIRE(fill_segments(seg_offsets, seg_lens, n_seg, n, &sd, &max_len, &seg_aligned, elem) == 0,
                "segment out of bounds");
    vec = vec && seg_aligned;
    if (err_over_tol_out) vec = false;   // q may be float64 while the state is float32: keep it simple
    int rc = -1;
    TDQ_DISPATCH_T(dtype, rc = dispatch_reduce_err<T>(nk, (const TdqCtrl *)ctrl_dev, y0, y1, kp, rtol_vec, atol_vec,
                                                      sd, n_seg, max_len, vec, partials, out, err_over_tol_out,
                                                      (cudaStream_t)stream));
    TDQ_REQUIRE(rc == 0, "unsupported number of error terms");
    TDQ_CHECK_CUDA(cudaGetLastError());
    return TDQ_OK;
}

int tdq_scaled_sumsq(void *ctrl_dev, int32_t dtype, const void *x, const void *x2, const void *y0,
                     const double *rtol_vec, const double *atol_vec, const int64_t *seg_offsets,
                     const int64_t *seg_lens, int32_t n_seg, size_t n, double *partials, double *out,
                     void *stream) {
    TDQ_REQUIRE(ctrl_dev && x && y0 && partials && out, "null argument");
    TDQ_REQUIRE(n_seg >= 1 && n_seg <= TDQ_MAX_SEGS, "n_seg out of range");
    TDQ_REQUIRE((rtol_vec == nullptr) == (atol_vec == nullptr), "rtol_vec and atol_vec go together");
    KPtrs kp;
    memset(&kp, 0, sizeof(kp));
    kp.p[0] = x;
    kp.p[1] = x2;
    bool vec = tdq_aligned16(x) && tdq_aligned16(y0) && (x2 == nullptr || tdq_aligned16(x2));
    SegDesc sd;
    int64_t max_len;
    bool seg_aligned;
    const size_t elem = dtype == TDQ_F32 ? 4 : 8;
    TDQ_REQUIRE(fill_segments(seg_offsets, seg_lens, n_seg, n, &sd, &max_len, &seg_aligned, elem) == 0,
                "segment out of bounds");
    vec = vec && seg_aligned;
    const TdqCtrl *c = (const TdqCtrl *)ctrl_dev;
    cudaStream_t st = (cudaStream_t)stream;
    if (x2) {
        TDQ_DISPATCH_T(dtype, (launch_reduce<T, 2, 1>(c, y0, nullptr, kp, rtol_vec, atol_vec, sd, n_seg, max_len, vec,
                                                      partials, out, nullptr, st, nullptr)));
    } else {
        TDQ_DISPATCH_T(dtype, (launch_reduce<T, 1, 1>(c, y0, nullptr, kp, rtol_vec, atol_vec, sd, n_seg, max_len, vec,
                                                      partials, out, nullptr, st, nullptr)));
    }
    TDQ_CHECK_CUDA(cudaGetLastError());
    return TDQ_OK;
}

int tdq_initial_step_probe(void *ctrl_dev, int32_t dtype, void *y_probe, const void *y0, const void *f0, size_t n,
                           void *stream) {
    TDQ_REQUIRE(ctrl_dev && y_probe && y0 && f0, "null argument");
    if (n == 0) return TDQ_OK;
    size_t blocks = (n + kThreads - 1) / kThreads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    TDQ_DISPATCH_T(dtype, (k_probe<T><<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(
                               (const TdqCtrl *)ctrl_dev, (T *)y_probe, (const T *)y0, (const T *)f0, n)));
    TDQ_CHECK_CUDA(cudaGetLastError());
    return TDQ_OK;
}

int tdq_interp_fit_commit(void *ctrl_dev, const tdq_tableau *tab, int32_t dtype, void *y0, const void *y1,
                          void *const *k, void *const *coeff, size_t n, void *stream) {
    TDQ_REQUIRE(ctrl_dev && tab && y0 && y1 && k && coeff, "null argument");
    HostShape hs;
    shape_from_tableau(tab, &hs);
    const int S = hs.n_stages;
    const int nk = hs.mid_nnz;
    TDQ_REQUIRE(nk >= 1, "tableau has no mid-point weights");
    KPtrs kmid;
    memset(&kmid, 0, sizeof(kmid));
    KPtrsMut fl;
    memset(&fl, 0, sizeof(fl));
    fl.p[0] = k[0];
    fl.p[1] = k[S];
    TDQ_REQUIRE(fl.p[0] && fl.p[1], "k_0 and k_S are required");
    bool vec = tdq_aligned16(y0) && tdq_aligned16(y1) && tdq_aligned16(fl.p[0]) && tdq_aligned16(fl.p[1]);
    for (int m = 0; m < nk; ++m) {
        kmid.p[m] = k[hs.mid_idx[m]];
        TDQ_REQUIRE(kmid.p[m] != nullptr, "missing stage slot for a non-zero mid-point weight");
        vec = vec && tdq_aligned16(kmid.p[m]);
    }
    for (int i = 0; i < 5; ++i) {
        TDQ_REQUIRE(coeff[i] != nullptr, "five coefficient buffers are required");
        vec = vec && tdq_aligned16(coeff[i]);
    }
    if (n == 0) return TDQ_OK;
    int rc = -1;
    TDQ_DISPATCH_T(dtype, rc = dispatch_fit<T>(nk, (const TdqCtrl *)ctrl_dev, y0, y1, fl, kmid, coeff, n, vec,
                                               (cudaStream_t)stream));
    TDQ_REQUIRE(rc == 0, "unsupported number of mid-point terms");
    TDQ_CHECK_CUDA(cudaGetLastError());
    return TDQ_OK;
}

int tdq_interp_eval(void *ctrl_dev, int32_t dtype, const void *const *coeff, void *solution, size_t n, void *stream) {
    TDQ_REQUIRE(ctrl_dev && coeff && solution, "null argument");
    bool vec = tdq_aligned16(solution) && ((n * (dtype == TDQ_F32 ? 4 : 8)) % 16 == 0);
    for (int i = 0; i < 5; ++i) {
        TDQ_REQUIRE(coeff[i] != nullptr, "five coefficient buffers are required");
        vec = vec && tdq_aligned16(coeff[i]);
    }
    if (n == 0) return TDQ_OK;
    TDQ_DISPATCH_T(dtype, (launch_eval<T, false>((const TdqCtrl *)ctrl_dev, coeff, solution, nullptr, n, vec,
                                                 (cudaStream_t)stream)));
    TDQ_CHECK_CUDA(cudaGetLastError());
    return TDQ_OK;
}

int tdq_poly_eval(int32_t dtype, const void *const *coeff, double x, void *out, size_t n, void *stream) {
    TDQ_REQUIRE(coeff && out, "null argument");
    for (int i = 0; i < 5; ++i) TDQ_REQUIRE(coeff[i] != nullptr, "five coefficient buffers are required");
    if (n == 0) return TDQ_OK;
    size_t blocks = (n + kThreads - 1) / kThreads;
    if (blocks > (size_t)sm_count() * 8) blocks = (size_t)sm_count() * 8;
    TDQ_DISPATCH_T(dtype, (k_poly_eval<T><<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(
                               (const T *)coeff[0], (const T *)coeff[1], (const T *)coeff[2], (const T *)coeff[3],
                               (const T *)coeff[4], (T *)out, x, n)));
    TDQ_CHECK_CUDA(cudaGetLastError());
    return TDQ_OK;
}

int tdq_interp_eval_at(void *ctrl_dev, int32_t dtype, const void *const *coeff, const double *t_dev, void *out,
                       size_t n, void *stream) {
    TDQ_REQUIRE(ctrl_dev && coeff && out && t_dev, "null argument");
    bool vec = tdq_aligned16(out);
    for (int i = 0; i < 5; ++i) {
        TDQ_REQUIRE(coeff[i] != nullptr, "five coefficient buffers are required");
        vec = vec && tdq_aligned16(coeff[i]);
    }
    if (n == 0) return TDQ_OK;
    TDQ_DISPATCH_T(dtype, (launch_eval<T, true>((const TdqCtrl *)ctrl_dev, coeff, out, t_dev, n, vec,
                                                (cudaStream_t)stream)));
    TDQ_CHECK_CUDA(cudaGetLastError());
    return TDQ_OK;
}

}  // extern "C"
