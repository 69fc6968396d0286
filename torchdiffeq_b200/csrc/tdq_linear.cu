// tdq_linear.cu -- a Runge-Kutta stage fused with a LINEAR vector field f(t, y) = y W^T on the 5th-generation tensor
// cores (float32 states of width 128, the field of BASELINE.json configs[1]).
//
//   y_i = y0 + sum_j coef_ij k_j        the stage combination of tdq_stream.cu, same products, same order, no fma
//   k_i = y_i W^T                       what the reference obtains by calling func(t_i, y_i)   (rk_common.py:79-81)
//
// y_i never goes to HBM: the producer warps form it in registers, split every float32 into three bfloat16 planes
// (hi + mid + lo = the 24-bit significand exactly) and store the planes as the B operand of tcgen05.mma in shared memory
// (K-major, 128-byte swizzle, two stages).  The weights are the stationary A operand, split the same way and kept in
// TENSOR MEMORY for the life of the CTA.  Six bf16 x bf16 products accumulate in float32 tensor-memory accumulators
// (the five cross terms >= 2^-16 in ascending magnitude into one, hi x hi into the other; the epilogue adds the two), which
// reproduces a float32 GEMM to float32 rounding (rel. rms error 1.0e-7 against float64; cuBLAS' own SIMT SGEMM: 2e-7) --
// the split of cuBLAS 12.9's CUBLAS_COMPUTE_32F_EMULATED_16BFX9 without its three terms below float32 resolution, here
// fused with the operand's producer.
// The accumulator is D^T (lane = output feature, column = state row), so a warp's store of one column is 128 contiguous
// bytes of k_i: no staging.
//
// Roles (384 threads, one CTA per SM, persistent over a contiguous range of 64-row units):
//   warps 0-7   producers: 128-bit streaming loads of y0 and the k_j, the combination, the split, st.shared of the planes;
//               for the row that yields y1 (FSAL) also y1 and the error-sum prefix, as k_combine_final does
//   warps 8-11  weights -> tensor memory once; per tile one thread issues the 48 MMAs, then all four drain the accumulators
// Measured on B200 (scripts/exp_fused_linear.cu, profiles/README.md), cold: 30 / 33 / 37 / 41 / 46 us for rows with 1..5
// terms at 65536 x 128 (k_combine + cuBLAS SGEMM: 65 .. 89 us), 60 us for the last row with y1 and the error prefix.
#include "tdq_shape.cuh"
#include "tdq_tc.cuh"

#include <cstdint>

namespace {

constexpr int LD = 128;                       // state width = number of output features = GEMM K and M
constexpr int ATOM_BYTES = 128 * 128;         // 128 rows x 128 bytes: one swizzle-atom column (64 bf16 of K)
constexpr int PLANE_BYTES = 2 * ATOM_BYTES;   // K = 128
constexpr int STAGE_BYTES = 3 * PLANE_BYTES;  // hi, mid, lo
constexpr int L_THREADS = 384;
constexpr int L_SMEM = 2 * STAGE_BYTES + 1024 + 128;
constexpr int TMEM_COLS = 512;
constexpr int COL_BIG = 0, COL_SMALL = 128, COL_W = 256;
constexpr int MAX_TERMS = 8;                  // stage terms per fused row

struct LinMap {                               // union of the row's and the error weights' slots (tdq_stream.cu FinalMap)
    signed char rpos[MAX_TERMS], epos[MAX_TERMS];
};
struct LinK {
    const float *p[MAX_TERMS];
};

using namespace tdq_tc;

constexpr uint32_t IDESC = tc_idesc(128);

// elements [4q, 4q+4) of row r inside one plane: K-major, 128-byte swizzle (atoms of 8 rows x 128 B, 16-byte chunk
// index XOR row mod 8), rows 128 B apart, the second 64 elements of K one ATOM further
__device__ __forceinline__ void store_split(uint8_t *stage, int r, int q, float4 v) {
    uint32_t h0, m0, l0, h1, m1, l1;
    split2(v.x, v.y, h0, m0, l0);
    split2(v.z, v.w, h1, m1, l1);
    const uint32_t off = (q >> 4) * ATOM_BYTES + r * 128 + ((((q & 15) >> 1) ^ (r & 7)) << 4) + (q & 1) * 8;
    *reinterpret_cast<uint2 *>(stage + off) = make_uint2(h0, h1);
    *reinterpret_cast<uint2 *>(stage + PLANE_BYTES + off) = make_uint2(m0, m1);
    *reinterpret_cast<uint2 *>(stage + 2 * PLANE_BYTES + off) = make_uint2(l0, l1);
}

// W [128 features][128] float32 -> wt[plane][feature][64 x u32]: what the thread that owns tensor-memory lane `feature` stores
__global__ void k_split_weights(const float *__restrict__ W, uint32_t *__restrict__ wt) {
    const int n = blockIdx.x, c = threadIdx.x;
    uint32_t h, m, l;
    split2(W[(size_t)n * LD + 2 * c], W[(size_t)n * LD + 2 * c + 1], h, m, l);
    wt[(0 * LD + n) * 64 + c] = h;
    wt[(1 * LD + n) * 64 + c] = m;
    wt[(2 * LD + n) * 64 + c] = l;
}

// NU: stage terms read.  MODE 0: no control block, NU = 0 (k = y W^T).  MODE 1: a middle row (coef[row][m], m < NU).
// MODE 2: the row that yields y1 of an FSAL tableau, with the error-sum prefix (LinMap; tdq_stream.cu k_combine_final).
// U: rows per producer warp in flight (registers: U * (NU + 1) 128-bit loads).
template <int NU, int MODE, int U>
__global__ void __launch_bounds__(L_THREADS, 1)
k_linear_stage(const TdqCtrl *__restrict__ c, int row, const float *y0, LinK kp, LinMap fm, const uint32_t *__restrict__ wt,
               float *__restrict__ kout, float *__restrict__ yout, float *__restrict__ eout, int n_rows) {
    if (MODE != 0 && c->halt) return;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 2 * STAGE_BYTES);     // full[2], empty[2], accf
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 2 * STAGE_BYTES + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t b_full[2] = {bar0, bar0 + 8}, b_empty[2] = {bar0 + 16, bar0 + 24};
    const uint32_t b_accf = bar0 + 32;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(b_full[0], 256);
        mbar_init(b_full[1], 256);
        mbar_init(b_empty[0], 1);
        mbar_init(b_empty[1], 1);
        mbar_init(b_accf, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;

    // contiguous range of 64-row units per CTA; a tile is two units (one at the end of an odd range)
    const int units = (n_rows + 63) / 64;
    const int u_begin = (int)((long long)blockIdx.x * units / gridDim.x);
    const int u_end = (int)((long long)(blockIdx.x + 1) * units / gridDim.x);

    if (warp < 8) {
        // ================= producers =================
        constexpr int NKK = NU > 0 ? NU : 1;
        float cr[NKK], ce[NKK];
        unsigned mask_r = 0, mask_e = 0;
        const float *k[NKK];
        if (MODE != 0 && y0 == nullptr) y0 = reinterpret_cast<const float *>(c->y0_cur);
#pragma unroll
        for (int m = 0; m < NU; ++m) {
            if (MODE == 1) {
                cr[m] = (float)c->coef[row][m];
                ce[m] = 0.f;
                mask_r |= 1u << m;
            } else if (MODE == 2) {
                const bool ur = fm.rpos[m] >= 0, ue = fm.epos[m] >= 0;
                if (ur) mask_r |= 1u << m;
                if (ue) mask_e |= 1u << m;
                cr[m] = ur ? (float)c->coef[row][fm.rpos[m]] : 0.f;
                ce[m] = ue ? (float)c->ecoef[fm.epos[m]] : 0.f;
            }
            k[m] = kp.p[m] ? kp.p[m] : reinterpret_cast<const float *>(c->k0_cur);
        }
        // The pointers above may come straight out of a global load (the control block's pointer table).  Used as they
        // are, every address computation in the loop below inherits that load's scoreboard -- the one the loop's own
        // loads are counted on -- and waits for the PREVIOUS rows' data before issuing the next rows' loads (measured:
        // 55 us instead of 39 us for a 3-term row).  One integer add with a value the compiler cannot fold (0 at run
        // time) makes them ALU results.
        {
            const size_t zero = (size_t)((unsigned)n_rows >> 31);
            y0 = reinterpret_cast<const float *>(reinterpret_cast<const char *>(y0) + zero);
#pragma unroll
            for (int m = 0; m < NU; ++m) k[m] = reinterpret_cast<const float *>(reinterpret_cast<const char *>(k[m]) + zero);
        }
        uint32_t it = 0;
        for (int u = u_begin; u < u_end; ++it) {
            const int nu = u_end - u < 2 ? u_end - u : 2;
            const int row0 = u * 64, rows_here = min(nu * 64, n_rows - row0);
            u += nu;
            const int s = it & 1;
            uint8_t *stage = smem + s * STAGE_BYTES;
            mbar_wait(b_empty[s], ((it >> 1) & 1) ^ 1);
#pragma unroll 1
            for (int i = 0; i < 16 / U; ++i) {
                if (U * i * 8 >= rows_here) break;
                float4 a[U], kv[U][NKK];
#pragma unroll
                for (int uu = 0; uu < U; ++uu) {
                    const int r = warp + 8 * (U * i + uu);
                    if (r < rows_here) {
                        const size_t off = (size_t)(row0 + r) * LD + lane * 4;
                        a[uu] = __ldcs(reinterpret_cast<const float4 *>(y0 + off));
#pragma unroll
                        for (int m = 0; m < NU; ++m) kv[uu][m] = __ldcs(reinterpret_cast<const float4 *>(k[m] + off));
                    }
                }
#pragma unroll
                for (int uu = 0; uu < U; ++uu) {
                    const int r = warp + 8 * (U * i + uu);
                    if (r < rows_here) {
                        float4 y = a[uu];
                        if (NU > 0) {
                            // each sum keeps the reference's ascending-j order over its own non-zero terms
                            // (rk_common.py:79, :89); products and sums rounded separately (--fmad=false)
                            float4 ar = make_float4(0.f, 0.f, 0.f, 0.f), ae = ar;
                            bool fr = true, fe = true;
#pragma unroll
                            for (int m = 0; m < NU; ++m) {
                                if ((mask_r >> m) & 1u) {
                                    const float4 p = make_float4(kv[uu][m].x * cr[m], kv[uu][m].y * cr[m], kv[uu][m].z * cr[m], kv[uu][m].w * cr[m]);
                                    ar = fr ? p : make_float4(ar.x + p.x, ar.y + p.y, ar.z + p.z, ar.w + p.w);
                                    fr = false;
                                }
                                if (MODE == 2 && ((mask_e >> m) & 1u)) {
                                    const float4 p = make_float4(kv[uu][m].x * ce[m], kv[uu][m].y * ce[m], kv[uu][m].z * ce[m], kv[uu][m].w * ce[m]);
                                    ae = fe ? p : make_float4(ae.x + p.x, ae.y + p.y, ae.z + p.z, ae.w + p.w);
                                    fe = false;
                                }
                            }
                            y = make_float4(y.x + ar.x, y.y + ar.y, y.z + ar.z, y.w + ar.w);
                            if (MODE == 2) {
                                const size_t off = (size_t)(row0 + r) * LD + lane * 4;
                                *reinterpret_cast<float4 *>(yout + off) = y;
                                *reinterpret_cast<float4 *>(eout + off) = ae;
                            }
                        }
                        store_split(stage, r, lane, y);
                    }
                }
            }
            fence_async_smem();
            mbar_arrive(b_full[s]);
        }
    } else {
        // ================= weights into tensor memory once; per tile: MMA issue (one thread), accumulators -> k =================
        const int e = warp & 3, f = e * 32 + lane;                 // this thread's tensor-memory lane = output feature
        const uint32_t lane_base = tmem + ((uint32_t)(e * 32) << 16);
#pragma unroll 1
        for (int pl = 0; pl < 3; ++pl) {
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 32) {
                uint32_t r[32];
                const uint4 *src = reinterpret_cast<const uint4 *>(wt + ((size_t)pl * LD + f) * 64 + c0);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint4 v = src[j];
                    r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
                }
                tmem_st32(lane_base + COL_W + pl * 64 + c0, r);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        uint32_t it = 0;
        for (int u = u_begin; u < u_end; ++it) {
            const int nu = u_end - u < 2 ? u_end - u : 2;
            const int row0 = u * 64, rows_here = min(nu * 64, n_rows - row0);
            u += nu;
            // the previous tile's accumulators are drained (the first time: the weights are in tensor memory)
            fence_before();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (warp == 8 && lane == 0) {
                const int s = it & 1;
                const uint32_t y_base = smem_u32(smem + s * STAGE_BYTES);
                mbar_wait(b_full[s], (it >> 1) & 1);
                fence_after();
                // weights plane PW (tensor memory) x stage-value plane PY (shared memory): mid.mid, lo.hi, hi.lo, mid.hi, hi.mid
                // into the small accumulator, hi.hi into the big one.  The three remaining cross terms (lo.lo, lo.mid, mid.lo)
                // are below 2^-24 of a product -- the rounding of a float32 product itself -- and measured irrelevant
                // (rel. rms error 1.03e-7 with six products, 1.02e-7 with nine; scripts/exp_fused_linear.cu), so they are not
                // computed: 48 instead of 72 MMAs per tile.
                constexpr int NPROD = 6;
                constexpr int PW[NPROD] = {1, 2, 0, 1, 0, 0}, PY[NPROD] = {1, 0, 2, 0, 1, 0};
#pragma unroll
                for (int p = 0; p < NPROD; ++p) {
                    const uint32_t dcol = tmem + (p == NPROD - 1 ? COL_BIG : COL_SMALL);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint32_t koff = (ks >> 2) * ATOM_BYTES + (ks & 3) * 32;
                        mma_ts(dcol, tmem + COL_W + PW[p] * 64 + ks * 8, make_desc(y_base + PY[p] * PLANE_BYTES + koff), IDESC,
                               (p == 0 || p == NPROD - 1) && ks == 0 ? 0u : 1u);
                    }
                }
                mma_commit(b_empty[s]);      // the stage may be refilled
                mma_commit(b_accf);          // the accumulators are complete
            }
            __syncwarp();
            mbar_wait(b_accf, it & 1);
            fence_after();
#pragma unroll 1
            for (int cc = 0; cc < 8; ++cc) {
                if (cc * 16 >= rows_here) break;
                uint32_t big[16], small[16];
                tmem_ld16(lane_base + COL_SMALL + cc * 16, small);
                tmem_ld16(lane_base + COL_BIG + cc * 16, big);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float *dst = kout + (size_t)(row0 + cc * 16) * LD + f;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (cc * 16 + j < rows_here) dst[(size_t)j * LD] = __uint_as_float(small[j]) + __uint_as_float(big[j]);
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(TMEM_COLS) : "memory");
    }
}

template <int NU, int MODE>
int launch_linear(const TdqCtrl *c, int row, const float *y0, const LinK &kp, const LinMap &fm, const uint32_t *wt, float *kout,
                  float *yout, float *eout, size_t n_rows, cudaStream_t st) {
    constexpr int U = NU <= 5 ? 4 : 2;
    auto kern = k_linear_stage<NU, MODE, U>;
    // per function AND per device; idempotent and cheap, so set on every launch (legal during stream capture)
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L_SMEM) != cudaSuccess) return -2;
    const size_t units = (n_rows + 63) / 64;
    size_t grid = (units + 1) / 2;
    const size_t cap = (size_t)tdq_sm_count();
    if (grid > cap) grid = cap;
    if (grid == 0) grid = 1;
    kern<<<(unsigned)grid, L_THREADS, L_SMEM, st>>>(c, row, y0, kp, fm, wt, kout, yout, eout, (int)n_rows);
    return 0;
}

template <int MODE>
int dispatch_linear(int nu, const TdqCtrl *c, int row, const float *y0, const LinK &kp, const LinMap &fm, const uint32_t *wt,
                    float *kout, float *yout, float *eout, size_t n_rows, cudaStream_t st) {
    switch (nu) {
#define TDQ_CASE(N) case N: return launch_linear<N, MODE>(c, row, y0, kp, fm, wt, kout, yout, eout, n_rows, st);
        TDQ_CASE(1) TDQ_CASE(2) TDQ_CASE(3) TDQ_CASE(4) TDQ_CASE(5) TDQ_CASE(6) TDQ_CASE(7) TDQ_CASE(8)
#undef TDQ_CASE
    }
    return -1;
}

bool linear_shape_ok(int32_t dtype, int32_t width) { return dtype == TDQ_F32 && width == LD; }

}  // namespace

extern "C" {

int tdq_linear_supported(int32_t dtype, int32_t width) { return linear_shape_ok(dtype, width) ? 1 : 0; }

size_t tdq_linear_weights_bytes(int32_t width) { return (size_t)3 * width * (width / 2) * 4; }

int tdq_linear_prepare(int32_t dtype, const void *weight, int32_t width, void *planes, void *stream) {
    TDQ_REQUIRE(weight && planes, "null argument");
    TDQ_REQUIRE(linear_shape_ok(dtype, width), "the fused linear field is float32, width 128");
    k_split_weights<<<LD, 64, 0, (cudaStream_t)stream>>>((const float *)weight, (uint32_t *)planes);
    TDQ_CHECK_CUDA(cudaGetLastError());
    return TDQ_OK;
}

int tdq_linear_apply(int32_t dtype, const void *y, const void *planes, int32_t width, size_t n_rows, void *k_out,
                     void *stream) {
    TDQ_REQUIRE(y && planes && k_out, "null argument");
    TDQ_REQUIRE(linear_shape_ok(dtype, width), "the fused linear field is float32, width 128");
    TDQ_REQUIRE(tdq_aligned16(y) && tdq_aligned16(k_out) && tdq_aligned16(planes), "operands must be 16-byte aligned");
    TDQ_REQUIRE(n_rows < ((size_t)1 << 31) - 64, "too many rows");
    if (n_rows == 0) return TDQ_OK;
    LinK kp;
    LinMap fm;
    memset(&kp, 0, sizeof(kp));
    memset(&fm, 0xff, sizeof(fm));
    const int rc = launch_linear<0, 0>(nullptr, 0, (const float *)y, kp, fm, (const uint32_t *)planes, (float *)k_out, nullptr,
                                       nullptr, n_rows, (cudaStream_t)stream);
    TDQ_REQUIRE(rc == 0, "launch configuration failed");
    TDQ_CHECK_CUDA(cudaGetLastError());
    return TDQ_OK;
}

int tdq_linear_stage(void *ctrl_dev, const tdq_tableau *tab, int32_t dtype, int32_t row, void *k_out, void *y1_out,
                     void *err_out, const void *y0, const void *const *k, const void *planes, int32_t width,
                     size_t n, void *stream) {
    TDQ_REQUIRE(ctrl_dev && tab && k_out && k && planes, "null argument");
    TDQ_REQUIRE(linear_shape_ok(dtype, width), "the fused linear field is float32, width 128");
    TDQ_REQUIRE(n % (size_t)width == 0, "state size is not a multiple of the field width");
    const size_t n_rows = n / (size_t)width;
    TDQ_REQUIRE(n_rows < ((size_t)1 << 31) - 64, "too many rows");
    TdqHostShape hs;
    tdq_shape_from_tableau(tab, &hs);
    const int S = hs.n_stages;
    TDQ_REQUIRE(row >= 0 && row < S, "row out of range");
    const bool final_row = hs.fsal && row == S - 1;
    TDQ_REQUIRE(final_row == (y1_out != nullptr) && final_row == (err_out != nullptr),
                "y1_out / err_out are given for, and only for, the row that yields y1 of an FSAL tableau");
    LinK kp;
    LinMap fm;
    memset(&kp, 0, sizeof(kp));
    memset(&fm, 0xff, sizeof(fm));
    bool aligned = tdq_aligned16(k_out) && tdq_aligned16(y0) && tdq_aligned16(planes) && tdq_aligned16(y1_out) &&
                   tdq_aligned16(err_out);
    int nu = 0;
    if (!final_row) {
        nu = hs.row_nnz[row];
        TDQ_REQUIRE(nu >= 1 && nu <= MAX_TERMS, "unsupported number of stage terms for the fused row");
        for (int m = 0; m < nu; ++m) {
            const int j = hs.row_idx[row][m];
            kp.p[m] = (const float *)k[j];
            TDQ_REQUIRE(kp.p[m] != nullptr || j == 0, "missing stage slot for a non-zero tableau entry");
            aligned = aligned && tdq_aligned16(kp.p[m]);
        }
    } else {
        // union of the row's and the error weights' slots, ascending (tdq_stage_combine_final)
        int used_r[TDQ_MAX_K], used_e[TDQ_MAX_K];
        for (int j = 0; j < TDQ_MAX_K; ++j) used_r[j] = used_e[j] = -1;
        for (int m = 0; m < hs.row_nnz[row]; ++m) used_r[hs.row_idx[row][m]] = m;
        for (int m = 0; m < hs.err_nnz; ++m)
            if (hs.err_idx[m] <= S - 1) used_e[hs.err_idx[m]] = m;
        for (int j = 0; j <= S - 1; ++j) {
            if (used_r[j] < 0 && used_e[j] < 0) continue;
            TDQ_REQUIRE(nu < MAX_TERMS, "unsupported number of stage terms for the fused row");
            kp.p[nu] = (const float *)k[j];
            TDQ_REQUIRE(kp.p[nu] != nullptr || j == 0, "missing stage slot for a non-zero tableau entry");
            aligned = aligned && tdq_aligned16(kp.p[nu]);
            fm.rpos[nu] = (signed char)used_r[j];
            fm.epos[nu] = (signed char)used_e[j];
            ++nu;
        }
        TDQ_REQUIRE(nu >= 1, "empty tableau row");
    }
    TDQ_REQUIRE(aligned, "operands must be 16-byte aligned");
    if (n_rows == 0) return TDQ_OK;
    const int rc = final_row ? dispatch_linear<2>(nu, (const TdqCtrl *)ctrl_dev, row, (const float *)y0, kp, fm,
                                                  (const uint32_t *)planes, (float *)k_out, (float *)y1_out, (float *)err_out,
                                                  n_rows, (cudaStream_t)stream)
                             : dispatch_linear<1>(nu, (const TdqCtrl *)ctrl_dev, row, (const float *)y0, kp, fm,
                                                  (const uint32_t *)planes, (float *)k_out, nullptr, nullptr, n_rows,
                                                  (cudaStream_t)stream);
    TDQ_REQUIRE(rc == 0, "launch configuration failed");
    TDQ_CHECK_CUDA(cudaGetLastError());
    return TDQ_OK;
}

}  // extern "C"
