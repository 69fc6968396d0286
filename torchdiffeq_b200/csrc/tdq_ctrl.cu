// tdq_ctrl.cu -- scalar side of the adaptive loop, kept on the device:
//   start-of-attempt bookkeeping   rk_common.py:266-308, :61-78
//   accept / reject + I controller rk_common.py:323-361, misc.py:85-95
//   initial step selection         misc.py:36-77
// One thread does the arithmetic (a few hundred flops); what matters is that nothing here needs the
// host, so a whole attempt can sit inside a CUDA graph.
#include "tdq_common.cuh"
#include "tdq_shape.cuh"

namespace {

template <typename T> __device__ __forceinline__ T prev_repr(T t);   // misc.py:358-364, Perturb.PREV
template <> __device__ __forceinline__ float prev_repr<float>(float t) { return nextafterf(t, __fsub_rn(t, 1.0f)); }
template <> __device__ __forceinline__ double prev_repr<double>(double t) { return nextafter(t, __dsub_rn(t, 1.0)); }

template <typename T> __device__ __forceinline__ T next_repr(T t);   // Perturb.NEXT
template <> __device__ __forceinline__ float next_repr<float>(float t) { return nextafterf(t, __fadd_rn(t, 1.0f)); }
template <> __device__ __forceinline__ double next_repr<double>(double t) { return nextafter(t, __dadd_rn(t, 1.0)); }

template <typename T> __device__ __forceinline__ void store_T(unsigned char *raw, int i, T v) {
    reinterpret_cast<T *>(raw)[i] = v;
}

// rk_common.py:266-308 (+ :246-247) for the attempt that starts at rk_state.t1 with rk_state.dt,
// then the casts and products of _runge_kutta_step (:61-79, :89) and _interp_fit's dt (:365-366).
// Split in two: the scalar decisions (one thread) and the per-attempt tables -- stage times and the
// coefficients fl_T(beta_ij * T(dt)) -- which are independent entries and are filled by the whole block.
template <typename T> __device__ void prepare_scalar(TdqCtrl &c) {
    if (c.halt) return;
    if (c.n_steps_interval >= c.max_num_steps) {                     // rk_common.py:247
        c.status = TDQ_RUN_MAX_STEPS;
        c.halt = 1;
        return;
    }
    double dt = c.dt;
    if (!isfinite(dt)) dt = c.min_step;                               // :269-270
    dt = fmin(fmax(dt, c.min_step), c.max_step);                      // :271
    const double t0 = c.t1;
    double t1 = t0 + dt;                                              // :273
    c.att_t0 = t0;
    c.att_dt = dt;
    if (!(t0 + dt > t0)) {                                            // :286
        c.status = TDQ_RUN_DT_UNDERFLOW;
        c.halt = 1;
        return;
    }
    if (c.y0_bad) {                                                   // :287 on the FIRST attempt (later ones: controller)
        c.status = TDQ_RUN_NONFINITE;
        c.halt = 1;
        return;
    }
    c.on_step_t = 0;
    if (c.n_step_t > 0) {                                             // :293-300
        const double nxt = c.step_t[c.next_step_index];
        if (t0 < nxt && nxt < t0 + dt) {
            c.on_step_t = 1;
            t1 = nxt;
            dt = t1 - t0;
        }
    }
    c.on_jump_t = 0;
    if (c.n_jump_t > 0) {                                             // :302-308 (after the step_t handling)
        const double nxt = c.jump_t[c.next_jump_index];
        if (t0 < nxt && nxt < t0 + dt) {
            c.on_jump_t = 1;
            c.on_step_t = 0;
            t1 = nxt;
            dt = t1 - t0;
        }
    }
    c.att_dt = dt;
    c.att_t1 = t1;
    c.att_dtT = (double)(T)dt;                                        // :61-65
}

template <typename T> __device__ void prepare_tables(TdqCtrl &c, int tid, int nthreads) {
    if (c.halt) return;
    using A = Ar<T>;
    const T t0T = (T)c.att_t0, dtT = (T)c.att_dt, t1T = (T)c.att_t1;  // :61-65
    const T sgn = (T)c.t_sign;
    const int S = c.n_stages;
    for (int i = tid; i < S; i += nthreads) {                         // :72-78
        const T a = (T)c.alpha[i];
        T ti;
        if (a == (T)1) ti = prev_repr<T>(t1T);
        else ti = A::add(t0T, A::mul(a, dtT));
        store_T<T>(c.tstage, i, A::mul(sgn, ti));
    }
    const int rows = c.fsal ? S : S + 1;
    for (int e = tid; e < rows * TDQ_MAX_K; e += nthreads) {          // :79 (beta_i * dt), :85 (dt * c_sol)
        const int r = e / TDQ_MAX_K, m = e % TDQ_MAX_K;
        if (m < c.row_nnz[r]) c.coef[r][m] = (double)A::mul(sgn, A::mul((T)c.beta[r][m], dtT));
    }
    for (int m = tid; m < c.err_nnz; m += nthreads)                   // :89
        c.ecoef[m] = (double)A::mul(sgn, A::mul(dtT, (T)c.c_err[m]));
}

template <typename T> __device__ void prepare_attempt(TdqCtrl &c) {
    prepare_scalar<T>(c);
    prepare_tables<T>(c, 0, 1);
}

// Value of the norm from per-segment sums: max over segments of sqrt(mean), each rounded to the
// ratio dtype (misc.py:22-23 _rms_norm, misc.py:30-33 _mixed_norm, adjoint.py:247-250).
// Computed by a whole block (any number of segments): thread t takes segments t, t+B, ...; max is order
// independent, so the result equals a serial loop's.  Every thread returns the value.
template <typename T, int THREADS>
__device__ double block_norm_from_sums(const TdqCtrl &c, const double *sums, const int64_t *counts, int n_seg,
                                       double *smem /* THREADS/32 + 1 */) {
    double best = 0.0;
    int nan = 0;
    for (int s = threadIdx.x; s < n_seg; s += THREADS) {
        const double cnt = counts ? (double)counts[s] : (double)c.n_global;
        if (cnt <= 0.0) continue;
        double r = sqrt(sums[s] / cnt);
        if (!c.ratio_f64) r = (double)(T)r;
        if (r != r) nan = 1;
        if (r > best) best = r;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_down_sync(0xffffffffu, best, o);
        const int on = __shfl_down_sync(0xffffffffu, nan, o);
        if (ob > best) best = ob;
        nan |= on;
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[w] = nan ? CUDART_NAN : best;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = 0.0;
        bool bn = false;
        for (int i = 0; i < THREADS / 32; ++i) {
            const double v = smem[i];
            if (v != v) bn = true;
            else if (v > b) b = v;
        }
        smem[THREADS / 32] = bn ? CUDART_NAN : b;
    }
    __syncthreads();
    return smem[THREADS / 32];
}

__device__ void write_mailbox(TdqCtrl &c, double fin_t0, double fin_dt, int jumped = 0) {
    c.seq += 1;
    tdq_mailbox *m = c.mbox;
    if (!m) return;
    // inside the device-side loop nobody polls between attempts: only the attempt that ends the solve reports
    // (saves the system-scope fence and the stores over PCIe on every other attempt)
    if (c.loop_handle != 0ull && !c.halt) return;
    m->status = c.status;
    m->accept = c.accept;
    m->done = c.done;
    m->out_cursor = c.out_cursor;
    m->n_accept = c.n_accept;
    m->n_reject = c.n_reject;
    m->t0 = c.t0;
    m->t1 = c.t1;
    m->dt = c.dt;
    m->ratio = c.ratio;
    m->att_t0 = fin_t0;
    m->att_dt = fin_dt;
    m->next_t0 = c.att_t0;
    m->next_dt = c.att_dt;
    m->on_jump_t = jumped;
    m->par = c.par;
    __threadfence_system();
    m->seq = c.seq;                  // kernel completion flushes this last store; no second fence needed
}

// rk_common.py:323-361 + misc.py:85-95, then the next attempt's constants.
template <typename T>
__device__ void controller(TdqCtrl &c, const double *norm_in, int n_seg, const void *ratio_dev, double ratio_pre) {
    if (c.halt) {
        // Attempts issued after the end are no-ops; the mailbox still ticks so a host that runs
        // ahead can account for every attempt it queued.  Clearing `accept` keeps fit/eval of such an
        // attempt from touching the finished solution.
        c.accept = 0;
        c.fit_now = 0;
        c.emit_lo = c.emit_hi;
        write_mailbox(c, c.att_t0, c.att_dt);
        return;
    }
    using A = Ar<T>;
    double ratio;
    if (ratio_dev) {
        ratio = c.ratio_f64 ? *reinterpret_cast<const double *>(ratio_dev)
                            : (double)*reinterpret_cast<const T *>(ratio_dev);
        ratio = fabs(ratio);                                          // misc.py:82 .abs()
    } else {
        ratio = ratio_pre;                                            // block_norm_from_sums
    }
    const bool y1_nonfinite = norm_in && norm_in[n_seg] > 0.0;
    if (y1_nonfinite && !ratio_dev) ratio = CUDART_NAN;               // a non-finite y1 poisons err/tol
    c.ratio = ratio;

    const double dt = c.att_dt;
    bool accept = ratio <= 1.0;                                       // :324
    if (dt > c.max_step) accept = false;                              // :327-328
    if (dt <= c.min_step) accept = true;                              // :329-330
    c.accept = accept ? 1 : 0;

    if (accept) {                                                     // :338-352
        c.t0 = c.att_t0;
        c.t1 = c.att_t1;
        c.n_accept += 1;
        // y_next = y1, f_next = f1 (:341, :352): the error-norm kernel has already written both into the other
        // pair of the pointer table; accepting is a flip.  The old pair stays valid for the interpolant fit.
        c.y0_prev = c.y0_cur;
        c.k0_prev = c.k0_cur;
        c.par ^= 1;
        c.y0_cur = c.ybuf[c.par];
        c.k0_cur = c.kbuf[c.par];
        if (c.on_step_t && c.next_step_index != c.n_step_t - 1) c.next_step_index += 1;
        if (c.on_jump_t) {                                            // :346-351
            if (c.next_jump_index != c.n_jump_t - 1) c.next_jump_index += 1;
            store_T<T>(c.taux, 2, A::mul((T)c.t_sign, next_repr<T>((T)c.att_t1)));
        }
        // constants _interp_fit needs from THIS attempt (rk_common.py:363-369)
        const T dtT = (T)c.att_dtT, sgn = (T)c.t_sign;
        c.fit_sdt = (double)A::mul(sgn, dtT);
        for (int m = 0; m < c.mid_nnz; ++m)
            c.fit_mcoef[m] = (double)A::mul(sgn, A::mul(dtT, (T)c.c_mid[m]));
        if (y1_nonfinite) {                                           // the next attempt would trip :287
            c.status = TDQ_RUN_NONFINITE;
            c.halt = 1;
        }
    } else {                                                          // :353-357
        c.t0 = c.att_t0;
        c.t1 = c.att_t0;
        c.n_reject += 1;
    }

    // misc.py:85-95 _optimal_step_size (float64), then the clamp of rk_common.py:359
    double dt_next;
    if (ratio == 0.0) {
        dt_next = dt * c.ifactor;
    } else {
        const double dfac = (ratio < 1.0) ? 1.0 : c.dfactor;
        const double expo = 1.0 / (double)c.order;
        const double cand = c.safety / pow(ratio, expo);
        double inner = (cand != cand || dfac != dfac) ? CUDART_NAN : fmax(cand, dfac);   // torch.max
        double factor = (inner != inner) ? CUDART_NAN : fmin(c.ifactor, inner);          // torch.min
        dt_next = dt * factor;
    }
    if (dt_next == dt_next) dt_next = fmin(fmax(dt_next, c.min_step), c.max_step);
    c.dt = dt_next;

    // Output cursor: solvers.py:33-34 asks for t[i] one at a time; every t[i] <= t1 is now covered
    // by this accepted interval (rk_common.py:246 loop condition `next_t > t1` is false for them).
    c.emit_lo = c.out_cursor;
    c.n_steps_interval += 1;
    if (accept) {
        int cur = c.out_cursor;
        while (cur < c.n_out && !(c.t_out[cur] > c.t1)) ++cur;
        if (cur != c.out_cursor) c.n_steps_interval = 0;
        c.out_cursor = cur;
    }
    c.emit_hi = c.out_cursor;
    // the interpolant is needed only when an output time fell into this step, or when the caller keeps it
    c.fit_now = (accept && (c.always_fit || c.emit_hi > c.emit_lo)) ? 1 : 0;
    if (c.out_cursor >= c.n_out) {
        c.done = 1;
        c.halt = 1;
    }
    const double fin_t0 = c.att_t0, fin_dt = c.att_dt;
    const int jumped = (accept && c.on_jump_t) ? 1 : 0;
    prepare_scalar<T>(c);                 // the tables of the next attempt are filled by the whole block (k_controller)
    write_mailbox(c, fin_t0, fin_dt, jumped);
}

// misc.py:55-63: h0 from d0, d1.
template <typename T>
__device__ void initial_h0(TdqCtrl &c, double d0d, double d1d) {
    using A = Ar<T>;
    double h0;
    if (c.ratio_f64) {
        h0 = (d0d < 1e-5 || d1d < 1e-5) ? (double)(T)1e-6 : fabs(0.01 * d0d / d1d);
    } else {
        const T d0 = (T)d0d, d1 = (T)d1d;
        T h;
        if (d0 < (T)1e-5 || d1 < (T)1e-5) h = (T)1e-6;                 // :60-61 (compare after promoting 1e-5)
        else h = A::div(A::mul((T)0.01, d0), d1);                      // :63
        h0 = (double)A::abs(h);
    }
    c.h0 = h0;
    c.ratio = d1d;                                                     // stash d1 for the finish kernel
    // probe time: t0 (f64) + h0 -> f64, cast to T by _PerturbFunc (misc.py:66-67, :187)
    const T tp = (T)(c.t1 + h0);
    store_T<T>(c.taux, 1, A::mul((T)c.t_sign, tp));
}

// misc.py:69-77
template <typename T>
__device__ void initial_finish(TdqCtrl &c, double nd) {
    using A = Ar<T>;
    double dt;
    const double order_p1 = (double)c.order;                           // called with order-1 (rk_common.py:217)
    if (c.ratio_f64) {
        const double d1 = c.ratio, h0 = c.h0;
        const double d2 = fabs(nd / h0);
        double h1;
        if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax((double)(T)1e-6, h0 * 1e-3);
        else h1 = pow(0.01 / ((d2 > d1) ? d2 : d1), 1.0 / order_p1);
        h1 = fabs(h1);
        dt = fmin(100.0 * h0, h1);
    } else {
        const T d1 = (T)c.ratio, h0 = (T)c.h0;
        const T d2 = A::abs(A::div((T)nd, h0));
        T h1;
        if (d1 <= (T)1e-15 && d2 <= (T)1e-15) {
            const T a = (T)1e-6, b = A::mul(h0, (T)1e-3);
            h1 = (a != a || b != b) ? (T)CUDART_NAN : (a > b ? a : b);
        } else {
            const T m = (d2 > d1) ? d2 : d1;                           // Python max(d1, d2)
            const T base = A::div((T)0.01, m);
            const T ex = (T)(1.0 / order_p1);
            h1 = (T)pow((double)base, (double)ex);
        }
        h1 = A::abs(h1);
        const T a = A::mul((T)100, h0);
        const T r = (a != a || h1 != h1) ? (T)CUDART_NAN : (a < h1 ? a : h1);   // torch.min
        dt = (double)r;
    }
    c.dt = dt;
}

template <typename T> __global__ void k_prepare(TdqCtrl *c, const double *y0_bad) {
    if (y0_bad && *y0_bad > 0.0) c->y0_bad = 1;
    prepare_attempt<T>(*c);
    c->y0_bad = 0;                       // later attempts start from states the controller has checked
    if (c->mbox) {                       // first attempt of a solve: let the host see its (t0, dt) and status
        c->mbox->status = c->status;
        c->mbox->next_t0 = c->att_t0;
        c->mbox->next_dt = c->att_dt;
        __threadfence_system();
    }
}
// The control block is ~10 KB of scalars that one thread reads and writes hundreds of times; from global
// memory every access is an L2 round trip (the r1 launch list showed 18 us per launch for ~300 flops).
// The block is staged through shared memory instead: 256 threads copy it in, thread 0 works on the shared
// copy, everybody copies it back.
constexpr int kCtrlThreads = 256;
static_assert(sizeof(TdqCtrl) % 8 == 0, "control block must be a whole number of 8-byte words");
static_assert(sizeof(TdqCtrl) <= 40 * 1024, "control block must fit static shared memory");

template <typename T>
__global__ void __launch_bounds__(kCtrlThreads)
k_controller(TdqCtrl *c, const double *norm_in, const int64_t *cnt, int n_seg, const void *ratio_dev) {
    __shared__ __align__(16) unsigned char raw[sizeof(TdqCtrl)];
    constexpr int kWords = (int)(sizeof(TdqCtrl) / 8);
    unsigned long long *sw = reinterpret_cast<unsigned long long *>(raw);
    const unsigned long long *gw = reinterpret_cast<const unsigned long long *>(c);
    for (int i = threadIdx.x; i < kWords; i += kCtrlThreads) sw[i] = gw[i];
    __syncthreads();
    TdqCtrl &sc = *reinterpret_cast<TdqCtrl *>(raw);
    __shared__ double xsum[TDQ_MAX_SEGS + 2];
    __shared__ int xfail;
    if (sc.xworld > 1 && !sc.halt && norm_in != nullptr && ratio_dev == nullptr && n_seg <= TDQ_MAX_SEGS) {
        // Fused all-reduce over NVLink peer memory: thread t talks to rank t.
        const int R = sc.xworld, me = sc.xrank, nv = n_seg + 1;
        const int par = (int)(((sc.xepoch & 1ull) << 1) | (sc.seq & 1ull));
        const unsigned long long want = (sc.xepoch << 32) | (sc.seq + 1ull);
        if (threadIdx.x == 0) xfail = 0;
        __syncthreads();
        if ((int)threadIdx.x < R) {
            const int t = threadIdx.x;
            TdqXBuf *peer = reinterpret_cast<TdqXBuf *>(sc.xpeer[t]);
            for (int i = 0; i < nv; ++i) peer->vals[par][me][i] = norm_in[i];          // P2P store
            __threadfence_system();
            asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(&peer->flags[par][me]), "l"(want) : "memory");
            TdqXBuf *mine = reinterpret_cast<TdqXBuf *>(sc.xpeer[me]);
            unsigned long long seen = 0, t0 = 0, now = 0;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            do {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(&mine->flags[par][t]) : "memory");
                if (seen == want) break;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            } while (now - t0 < 10000000000ull);                                       // 10 s: a peer died
            if (seen != want) atomicExch(&xfail, 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (xfail) {
                sc.status = TDQ_RUN_EXCHANGE_TIMEOUT;
                sc.halt = 1;
            } else {
                const TdqXBuf *mine = reinterpret_cast<const TdqXBuf *>(sc.xpeer[me]);
                for (int i = 0; i < nv; ++i) {
                    double a = 0.0;
                    for (int r = 0; r < R; ++r) a += __ldcg(&mine->vals[par][r][i]);       // rank order: same sum everywhere
                    xsum[i] = a;
                }
            }
        }
        __syncthreads();
        norm_in = xsum;
    }
    __shared__ double nsm[kCtrlThreads / 32 + 1];
    double ratio_pre = 0.0;
    if (!sc.halt && ratio_dev == nullptr)
        ratio_pre = block_norm_from_sums<T, kCtrlThreads>(sc, norm_in, cnt, n_seg, nsm);
    __shared__ int was_halted;
    if (threadIdx.x == 0) {
        was_halted = sc.halt;
        controller<T>(sc, norm_in, n_seg, ratio_dev, ratio_pre);
    }
    __syncthreads();
    if (!was_halted) prepare_tables<T>(sc, threadIdx.x, kCtrlThreads);   // a no-op once the solve has halted
    __syncthreads();
    unsigned long long *go = reinterpret_cast<unsigned long long *>(c);
    for (int i = threadIdx.x; i < kWords; i += kCtrlThreads) go[i] = sw[i];
    // Device-side while loop (tdq_loop_create): this attempt's graph is the body of a conditional WHILE node;
    // another iteration runs only while the solve has neither finished nor failed.
    if (threadIdx.x == 0 && sc.loop_handle != 0ull)
        cudaGraphSetConditional((cudaGraphConditionalHandle)sc.loop_handle, sc.halt ? 0u : 1u);
}
constexpr int kInitThreads = 128;
template <typename T>
__global__ void __launch_bounds__(kInitThreads)
k_initial_h0(TdqCtrl *c, const double *s0, const double *s1, const int64_t *cnt, int n_seg) {
    __shared__ double nsm[kInitThreads / 32 + 1];
    const double d0 = block_norm_from_sums<T, kInitThreads>(*c, s0, cnt, n_seg, nsm);
    const double d1 = block_norm_from_sums<T, kInitThreads>(*c, s1, cnt, n_seg, nsm);
    if (threadIdx.x == 0) initial_h0<T>(*c, d0, d1);
}
template <typename T>
__global__ void __launch_bounds__(kInitThreads)
k_initial_finish(TdqCtrl *c, const double *s2, const int64_t *cnt, int n_seg) {
    __shared__ double nsm[kInitThreads / 32 + 1];
    const double nd = block_norm_from_sums<T, kInitThreads>(*c, s2, cnt, n_seg, nsm);
    if (threadIdx.x == 0) initial_finish<T>(*c, nd);
}
__global__ void k_set_loop(TdqCtrl *c, unsigned long long h) { c->loop_handle = h; }
__global__ void k_set_first_step(TdqCtrl *c, double dt) { c->dt = dt; }
struct XPtrs { const void *p[TDQ_MAX_RANKS]; };
__global__ void k_set_exchange(TdqCtrl *c, XPtrs xp, int rank, int world, unsigned long long epoch) {
    for (int r = 0; r < TDQ_MAX_RANKS; ++r) c->xpeer[r] = const_cast<void *>(xp.p[r]);
    c->xrank = rank;
    c->xworld = world;
    c->xepoch = epoch;
}
__global__ void k_reset_interval(TdqCtrl *c) { c->n_steps_interval = 0; }
__global__ void k_set_jump_t(TdqCtrl *c, const double *jt, int n) {
    c->jump_t = jt;
    c->n_jump_t = n;
    int idx = 0;                                                      // rk_common.py:241
    while (idx < n && !(jt[idx] > c->t1)) ++idx;
    c->next_jump_index = (idx < n - 1) ? idx : (n - 1);
    if (n <= 0) c->next_jump_index = 0;
}
__global__ void k_set_step_t(TdqCtrl *c, const double *st, int n) {
    c->step_t = st;
    c->n_step_t = n;
    // rk_common.py:240: min(bisect(step_t, t0), len-1)
    int idx = 0;
    while (idx < n && !(st[idx] > c->t1)) ++idx;
    c->next_step_index = (idx < n - 1) ? idx : (n - 1);
    if (n <= 0) c->next_step_index = 0;
}

}  // namespace

// The dtype is stored in the block, but launchers must not read device memory: the host passes it.
extern "C" {

size_t tdq_ctrl_size(void) { return (sizeof(TdqCtrl) + 255) & ~size_t(255); }
size_t tdq_ctrl_tstage_offset(void) { return offsetof(TdqCtrl, tstage); }
size_t tdq_ctrl_taux_offset(void) { return offsetof(TdqCtrl, taux); }

static float round_f32(double x) { return (float)x; }

int tdq_ctrl_init(void *ctrl_dev, const tdq_tableau *tab, const tdq_options *opt, const double *t_out_dev,
                  double t_start, int32_t n_out, void *mailbox_dev, void *stream) {
    TDQ_REQUIRE(ctrl_dev && tab && opt, "null argument");
    TDQ_REQUIRE(tab->n_stages >= 1 && tab->n_stages <= TDQ_MAX_STAGES, "n_stages out of range");
    TDQ_REQUIRE(opt->dtype == TDQ_F32 || opt->dtype == TDQ_F64, "unsupported dtype");
    TDQ_REQUIRE(n_out >= 1, "need at least one output time");
    static thread_local TdqCtrl h;   // ~10 KB: keep off the stack
    memset(&h, 0, sizeof(h));
    const bool f32 = opt->dtype == TDQ_F32;
    auto rT = [&](double x) { return f32 ? (double)round_f32(x) : x; };   // rk_common.py:201-205 cast
    const int S = tab->n_stages;
    h.dtype = opt->dtype;
    h.n_stages = S;
    h.order = tab->order;
    h.fsal = tab->fsal ? 1 : 0;
    h.ratio_f64 = (opt->ratio_f64 || !f32) ? 1 : 0;
    h.n_out = n_out;
    for (int i = 0; i < S; ++i) {
        h.alpha[i] = rT(tab->alpha[i]);
        int m = 0;
        for (int j = 0; j <= i; ++j)
            if (tab->beta[i][j] != 0.0) { h.row_idx[i][m] = j; h.beta[i][m] = rT(tab->beta[i][j]); ++m; }
        h.row_nnz[i] = m;
    }
    {
        int m = 0;
        for (int j = 0; j <= S; ++j)
            if (tab->c_sol[j] != 0.0) { h.row_idx[S][m] = j; h.beta[S][m] = rT(tab->c_sol[j]); ++m; }
        h.row_nnz[S] = m;
        m = 0;
        for (int j = 0; j <= S; ++j)
            if (tab->c_err[j] != 0.0) { h.err_idx[m] = j; h.c_err[m] = rT(tab->c_err[j]); ++m; }
        h.err_nnz = m;
        m = 0;
        for (int j = 0; j <= S; ++j)
            if (tab->c_mid[j] != 0.0) { h.mid_idx[m] = j; h.c_mid[m] = rT(tab->c_mid[j]); ++m; }
        h.mid_nnz = m;
    }
    h.rtol = opt->rtol;
    h.atol = opt->atol;
    h.min_step = opt->min_step;
    h.max_step = opt->max_step;
    h.safety = opt->safety;
    h.ifactor = opt->ifactor;
    h.dfactor = opt->dfactor;
    h.t_sign = (opt->t_sign < 0) ? -1.0 : 1.0;
    h.max_num_steps = opt->max_num_steps;
    h.n_global = opt->n_global;
    h.t_out = t_out_dev;
    h.mbox = reinterpret_cast<tdq_mailbox *>(mailbox_dev);
    h.t0 = h.t1 = t_start;                                            // rk_common.py:221
    h.dt = 0.0;
    TDQ_REQUIRE((opt->ybuf[0] == nullptr) == (opt->ybuf[1] == nullptr) &&
                    (opt->ybuf[0] == nullptr) == (opt->kbuf[0] == nullptr) &&
                    (opt->ybuf[0] == nullptr) == (opt->kbuf[1] == nullptr),
                "ybuf/kbuf: give all four state buffers or none");
    for (int i = 0; i < 2; ++i) {
        TDQ_REQUIRE(tdq_aligned16(opt->ybuf[i]) && tdq_aligned16(opt->kbuf[i]), "state buffers must be 16-byte aligned");
        h.ybuf[i] = opt->ybuf[i];
        h.kbuf[i] = opt->kbuf[i];
    }
    h.par = 0;
    h.y0_cur = h.y0_prev = h.ybuf[0];
    h.k0_cur = h.k0_prev = h.kbuf[0];
    h.always_fit = opt->always_fit ? 1 : 0;
    h.loop_handle = opt->loop_handle;
    h.out_cursor = 1;                                                 // solution[0] = y0 (solvers.py:30)
    h.emit_lo = h.emit_hi = 1;
    if (n_out <= 1) { h.done = 1; h.halt = 1; }
    if (f32) {
        float v = (float)h.t_sign * (float)t_start;
        memcpy(h.taux, &v, sizeof(v));
    } else {
        double v = h.t_sign * t_start;
        memcpy(h.taux, &v, sizeof(v));
    }
    // Pageable source: the runtime stages the bytes before returning, so `h` may be reused.
    TDQ_CHECK_CUDA(cudaMemcpyAsync(ctrl_dev, &h, sizeof(h), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return TDQ_OK;
}

int tdq_ctrl_set_step_t(void *ctrl_dev, const double *step_t_dev, int32_t n, void *stream) {
    TDQ_REQUIRE(ctrl_dev, "null ctrl");
    k_set_step_t<<<1, 1, 0, (cudaStream_t)stream>>>((TdqCtrl *)ctrl_dev, step_t_dev, n);
    TDQ_CHECK_CUDA(cudaGetLastError());
    return TDQ_OK;
}

int tdq_ctrl_set_jump_t(void *ctrl_dev, const double *jump_t_dev, int32_t n, void *stream) {
    TDQ_REQUIRE(ctrl_dev, "null ctrl");
    k_set_jump_t<<<1, 1, 0, (cudaStream_t)stream>>>((TdqCtrl *)ctrl_dev, jump_t_dev, n);
    TDQ_CHECK_CUDA(cudaGetLastError());
    return TDQ_OK;
}

int tdq_prepare_attempt(void *ctrl_dev, int32_t dtype, const double *y0_nonfinite_count_dev, void *stream) {
    TDQ_REQUIRE(ctrl_dev, "null ctrl");
    TDQ_DISPATCH_T(dtype, (k_prepare<T><<<1, 1, 0, (cudaStream_t)stream>>>((TdqCtrl *)ctrl_dev, y0_nonfinite_count_dev)));
    TDQ_CHECK_CUDA(cudaGetLastError());
    return TDQ_OK;
}

int tdq_controller(void *ctrl_dev, int32_t dtype, const double *norm_in, const int64_t *seg_counts_dev,
                   int32_t n_seg, const void *ratio_dev, void *stream) {
    TDQ_REQUIRE(ctrl_dev, "null ctrl");
    TDQ_REQUIRE(norm_in || ratio_dev, "need norm sums or an explicit ratio");
    TDQ_REQUIRE(n_seg >= 1, "n_seg out of range");      // any number of segments (the peer exchange: <= TDQ_MAX_SEGS)
    TDQ_DISPATCH_T(dtype, (k_controller<T><<<1, kCtrlThreads, 0, (cudaStream_t)stream>>>(
                               (TdqCtrl *)ctrl_dev, norm_in, seg_counts_dev, n_seg, ratio_dev)));
    TDQ_CHECK_CUDA(cudaGetLastError());
    return TDQ_OK;
}

int tdq_initial_step_h0(void *ctrl_dev, int32_t dtype, const double *d0_sumsq, const double *d1_sumsq,
                        const int64_t *seg_counts_dev, int32_t n_seg, void *stream) {
    TDQ_REQUIRE(ctrl_dev && d0_sumsq && d1_sumsq, "null argument");
    TDQ_DISPATCH_T(dtype, (k_initial_h0<T><<<1, kInitThreads, 0, (cudaStream_t)stream>>>(
                               (TdqCtrl *)ctrl_dev, d0_sumsq, d1_sumsq, seg_counts_dev, n_seg)));
    TDQ_CHECK_CUDA(cudaGetLastError());
    return TDQ_OK;
}

int tdq_initial_step_finish(void *ctrl_dev, int32_t dtype, const double *d2_sumsq,
                            const int64_t *seg_counts_dev, int32_t n_seg, void *stream) {
    TDQ_REQUIRE(ctrl_dev && d2_sumsq, "null argument");
    TDQ_DISPATCH_T(dtype, (k_initial_finish<T><<<1, kInitThreads, 0, (cudaStream_t)stream>>>(
                               (TdqCtrl *)ctrl_dev, d2_sumsq, seg_counts_dev, n_seg)));
    TDQ_CHECK_CUDA(cudaGetLastError());
    return TDQ_OK;
}

int tdq_set_first_step(void *ctrl_dev, double first_step, void *stream) {
    TDQ_REQUIRE(ctrl_dev, "null ctrl");
    k_set_first_step<<<1, 1, 0, (cudaStream_t)stream>>>((TdqCtrl *)ctrl_dev, first_step);
    TDQ_CHECK_CUDA(cudaGetLastError());
    return TDQ_OK;
}

int tdq_ctrl_set_exchange(void *ctrl_dev, const void *const *peer_ptrs, int32_t rank, int32_t world, uint64_t epoch,
                          void *stream) {
    TDQ_REQUIRE(ctrl_dev && peer_ptrs, "null argument");
    TDQ_REQUIRE(world >= 1 && world <= TDQ_MAX_RANKS && rank >= 0 && rank < world, "rank/world out of range");
    XPtrs xp;
    memset(&xp, 0, sizeof(xp));
    for (int r = 0; r < world; ++r) {
        TDQ_REQUIRE(peer_ptrs[r] != nullptr, "missing peer buffer");
        xp.p[r] = peer_ptrs[r];
    }
    k_set_exchange<<<1, 1, 0, (cudaStream_t)stream>>>((TdqCtrl *)ctrl_dev, xp, rank, world, (unsigned long long)epoch);
    TDQ_CHECK_CUDA(cudaGetLastError());
    return TDQ_OK;
}

int tdq_xchg_create(void **dev_ptr, tdq_ipc_handle *handle_out) {
    TDQ_REQUIRE(dev_ptr && handle_out, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(tdq_ipc_handle), "IPC handle does not fit");
    void *p = nullptr;
    TDQ_CHECK_CUDA(cudaMalloc(&p, sizeof(TdqXBuf)));
    TDQ_CHECK_CUDA(cudaMemset(p, 0, sizeof(TdqXBuf)));
    TDQ_CHECK_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        tdq_set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        return TDQ_ERR_CUDA;
    }
    memset(handle_out, 0, sizeof(*handle_out));
    memcpy(handle_out->bytes, &h, sizeof(h));
    *dev_ptr = p;
    return TDQ_OK;
}

int tdq_xchg_open(const tdq_ipc_handle *handle, void **peer_ptr) {
    TDQ_REQUIRE(handle && peer_ptr, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle->bytes, sizeof(h));
    TDQ_CHECK_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return TDQ_OK;
}

int tdq_xchg_close(void *peer_ptr) {
    if (peer_ptr) TDQ_CHECK_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return TDQ_OK;
}

int tdq_xchg_destroy(void *dev_ptr) {
    if (dev_ptr) TDQ_CHECK_CUDA(cudaFree(dev_ptr));
    return TDQ_OK;
}

int tdq_ctrl_set_loop(void *ctrl_dev, uint64_t loop_handle, void *stream) {
    TDQ_REQUIRE(ctrl_dev, "null ctrl");
    k_set_loop<<<1, 1, 0, (cudaStream_t)stream>>>((TdqCtrl *)ctrl_dev, (unsigned long long)loop_handle);
    TDQ_CHECK_CUDA(cudaGetLastError());
    return TDQ_OK;
}

int tdq_ctrl_reset_interval(void *ctrl_dev, void *stream) {
    TDQ_REQUIRE(ctrl_dev, "null ctrl");
    k_reset_interval<<<1, 1, 0, (cudaStream_t)stream>>>((TdqCtrl *)ctrl_dev);
    TDQ_CHECK_CUDA(cudaGetLastError());
    return TDQ_OK;
}

}  // extern "C"
