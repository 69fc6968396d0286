// tdq_attempt.cu -- a WHOLE Runge-Kutta attempt of a LINEAR vector field f(t, y) = y W^T in one launch: every stage
// combination, every evaluation of the field, the error estimate, its squared norm and the candidate commit.
//
// Why this is possible: for a linear field an attempt is ROW-LOCAL.  k_i = (y0 + sum_j coef_ij k_j) W^T needs nothing from
// other state rows, so a tile of rows can be taken through all S stages without leaving the SM: y0 and the k_j stay in
// registers / shared memory, the stage value goes to shared memory as the B operand of tcgen05.mma (three bfloat16 planes,
// as in tdq_linear.cu), the product comes back from tensor memory.  HBM sees 2 reads (y0, k_0) and 2 writes (the candidate pair
// y1, k_S) per element and attempt instead of the 34 + 6 of six fused stage launches plus the norm launch (tdq_linear.cu,
// tdq_norm.cu), and one launch instead of seven; the attempt becomes tensor bound (S x 6 bf16 products per element).
// What rk_common.py:43-90 (_runge_kutta_step), misc.py:80-82 (_compute_error_ratio up to the mean) and the assignment
// y_next = y1, f_next = f1 of rk_common.py:341/:352 do, for the FSAL tableaus dopri5 and bosh3.
//
// Arithmetic is the stage kernels' arithmetic, operation for operation: products and sums of the combination rounded
// separately in ascending j (rk_common.py:79), the same float32 -> hi + mid + lo split, the same six bf16 products in the
// same accumulation order -- k_i, y1 and the error-sum prefix are BITWISE what tdq_linear_stage writes, (err/tol)^2 per
// element bitwise what k_norm computes; only the order of the float64 sum over elements differs (tests/test_gpu_linear.py).
//
// Layout: 512 threads = 4 independent tile pipelines ("groups") of 4 warps, one CTA per SM.  A tile is 16 state rows x 128
// features.  The accumulator is D^T (lane = output feature, column = state row; M = 128, N = 16), so thread (warp e, lane l)
// of a group owns feature f = 32 e + l of all 16 rows: a warp's access to one row is 128 contiguous bytes of global memory, and
// its 8 consecutive rows of one feature are one 16-byte vector of an MN-major core matrix of the B operand (6 st.shared.v4 per
// stage).  State per thread: k_0..k_3 in registers (64), y0 in shared memory; once k_0..k_3 are known the remaining rows and the
// error estimate are running sums that each later k_j is folded into.  Per stage a group does: newest term + split + st.shared,
// fence.proxy.async, bar.sync (its 128 threads), one elected thread issues 48 MMAs (weights stationary in tensor memory, 192
// columns, shared by the groups) and a commit; while they run, the prefix of the next row's sum; then everybody waits on the
// group's mbarrier and drains 2 x 16 columns.  The chain of a tile is serial by nature (stage i+1 needs k_i); the four groups
// interleave, one hiding the other's latency.  The last block to finish adds the per-block partials of the squared error norm
// and runs the controller step (tdq_ctrl_dev.cuh).  Measurements and the variants tried: DESIGN.md section 3c.
//
// Stage derivatives, y1 and the error prefix are written to HBM only for attempts that can contain an output time (the
// lazy interpolant fit needs them, tdq_interp.cu) or when the caller keeps every step (dense output, events).
#include "tdq_shape.cuh"
#include "tdq_tc.cuh"
#include "tdq_ctrl_dev.cuh"

#include <type_traits>

namespace {

using namespace tdq_tc;

constexpr int LD = 128;                        // state width = output features = GEMM K and M
constexpr int AT_ROWS = 16;                    // state rows per tile = MMA N
constexpr int AT_PLANE = AT_ROWS * LD * 2;      // one bfloat16 plane of a tile: 16 rows x 128 features
constexpr int AT_STAGE = 3 * AT_PLANE;         // hi, mid, lo
constexpr int AT_AUX = 2048;                   // barriers, tensor-memory slot, coefficient tables, reduction scratch
constexpr int AT_Y0 = AT_ROWS * LD * 4;         // a tile's y0 (float32) stays in shared memory: read once per stage
// G: tile pipelines (groups of 4 warps) per CTA
constexpr int at_smem(int G) { return G * (AT_STAGE + AT_Y0) + AT_AUX + 1024; }
constexpr int AT_TMEM_COLS = 512;
constexpr int AT_COL_W = 256;                  // weights: 3 planes x 64 columns; accumulators of group g: 32 ACCS columns from 32 ACCS g
constexpr int AT_MAX_S = 7;
constexpr uint32_t IDESC16 = tc_idesc(AT_ROWS) | (1u << 16);   // B (the stage planes) is MN-major: bit 16

// Stage planes: MN-major, no swizzle.  Core matrix = 8 rows (MN, contiguous: 16 bytes) x 8 features (K, 16 bytes apart) =
// 128 contiguous bytes; core matrices 128 B apart along K (leading byte offset), 2048 B apart along the rows (stride byte
// offset).  Element (row n, feature k) of a plane sits at (n >> 3) * 2048 + (k >> 3) * 128 + (k & 7) * 16 + (n & 7) * 2, so the
// thread that owns feature k stores rows 0-7 and rows 8-15 as two 16-byte vectors, and a warp's store is 512 contiguous bytes.
constexpr uint32_t AT_LBO = 128, AT_SBO = 2048;
// descriptor: start >> 4 in [0,14), leading byte offset >> 4 in [16,30), stride byte offset >> 4 in [32,46), version 1 in
// [46,48), layout type 0 (no swizzle) in [61,64)  (verified on a B200: LBO is the K direction, SBO the MN direction)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(AT_LBO >> 4) << 16) | ((uint64_t)(AT_SBO >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// A value that is zero at run time but depends on every element of `a`: added to the address of the mbarrier the group is
// about to wait on, it forces the arithmetic of the MMA window to be issued BEFORE the wait (ptxas otherwise sinks part of it
// below the wait, onto the critical path).  `zero` is a run-time zero the compiler cannot fold.  8 LOP3 per array.
template <int N>
__device__ __forceinline__ uint32_t dep16(const float (&a)[N], uint32_t zero) {
    uint32_t x = 0;
#pragma unroll
    for (int r = 0; r < N; ++r) x ^= __float_as_uint(a[r]);
    return x & zero;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
template <int N> __device__ __forceinline__ void tmem_ldn(uint32_t taddr, uint32_t (&r)[N]);
template <> __device__ __forceinline__ void tmem_ldn<16>(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld16(taddr, r); }
template <> __device__ __forceinline__ void tmem_ldn<8>(uint32_t taddr, uint32_t (&r)[8]) { tmem_ld8(taddr, r); }
// one lane of a converged warp (the compiler keeps the operands of what follows in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\t@p mov.s32 %0, 1;\n\t}\n" : "+r"(pred));
    return pred != 0;
}
// explicit shared-space accesses with 32-bit addresses (a pointer derived from the aligned dynamic shared memory base is
// generic to the compiler: 64-bit address registers and generic ST/LD otherwise)
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

struct AttOut {
    float *k[AT_MAX_S + 1];                    // k[i], i = 1..S: where k_i goes when the attempt's stages are kept
    float *y1, *err;
};

__host__ __device__ constexpr int popc_below(unsigned mask, int j) {
    int n = 0;
    for (int b = 0; b < j; ++b) n += (mask >> b) & 1u;
    return n;
}

// S: stages of an FSAL tableau (rows 0..S-1, the last one is c_sol and yields y1).  RM: 8 bits per row, bit j set <=> slot j has
// a non-zero coefficient in that row.  EM: the same for the error weights of slots 0..S-1 (k_S always carries the last one).
// CTRL: the last block to finish also runs the controller step (tdq_ctrl_dev.cuh: accept / reject, next step size, the next
// attempt's tables, the device-side loop's condition) -- one launch per attempt instead of two.  That instantiation carries a
// device-runtime call (cudaGraphSetConditional), which kernel-level profilers refuse; the CTRL = false one is what ncu sees.
// RT: rows of a tile per thread.  16: one warpgroup (4 warps) per tile; 8: two warpgroups per tile, each thread half the rows
// (warps w and w + 4 own the same tensor-memory lanes and drain columns [0, 8) / [8, 16)): half the registers and half the
// dependent work per thread, twice the warps to hide latencies with.
// ACCS: accumulators per partial product group (1: one small + one big; 2: the K steps alternate between two of each --
// consecutive MMAs then never accumulate into the same tensor-memory tile).
template <int S, unsigned long long RM, unsigned EM, int G, int RT, bool CTRL, int ACCS>
__global__ void __launch_bounds__(G * 128 * (16 / RT), 1)
k_linear_attempt(TdqCtrl *c, const float *y0, const float *k0, AttOut out, const uint32_t *__restrict__ wt,
                 double *partials, double *norm_out, const int64_t *seg_counts, int store_always, size_t n_rows_sz) {
    if (c->halt) {
        // An attempt issued after the end of the solve is a no-op -- but its controller step still has to tick the mailbox:
        // a host that runs ahead (eager run_ahead, graph replay) accounts for every attempt it queued (tdq_ctrl_dev.cuh).
        if (CTRL && blockIdx.x == 0) tdq_ctrl_dev::controller_block<float, G * 128 * (16 / RT)>(c, norm_out, seg_counts, 1, nullptr);
        return;
    }
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int AT_GROUPS = G, HN = AT_ROWS / RT, WPG = 4 * HN, AT_THREADS = G * 128 * HN;
    uint8_t *aux = smem + G * (AT_STAGE + AT_Y0);
    uint64_t *bars = reinterpret_cast<uint64_t *>(aux);                    // one "accumulators complete" barrier per group
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(aux + 32);
    int *s_flag = reinterpret_cast<int *>(aux + 40);
    float *s_cr = reinterpret_cast<float *>(aux + 64);                    // [S][8]: coef[i][m] as float32
    float *s_ce = s_cr + AT_MAX_S * 8;                                    // [8]:    ecoef[m]
    double *s_red = reinterpret_cast<double *>(aux + 512);                // [2][32]
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);                  // warp-uniform for the compiler as well
    const int n_rows = (int)n_rows_sz;
    const uint32_t bar0 = smem_u32(bars);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(AT_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
#pragma unroll
        for (int g = 0; g < AT_GROUPS; ++g) mbar_init(bar0 + 8 * g, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid >= 64 && tid < 64 + AT_MAX_S * 8) {                           // this attempt's coefficients (prepare_tables)
        const int i = (tid - 64) >> 3, m = (tid - 64) & 7;
        s_cr[i * 8 + m] = (i < S) ? (float)c->coef[i][m] : 0.f;
        if (i == 0) s_ce[m] = (float)c->ecoef[m];
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    // group g = one tile pipeline; e = tensor-memory lane quarter (= warp index mod 4); h = which RT rows of the tile
    const int g = warp / WPG, wl = warp % WPG, e = wl & 3, h = wl >> 2, f = e * 32 + lane;   // f: this thread's lane = feature
    const uint32_t lane_base = tmem + ((uint32_t)(e * 32) << 16);
    {   // weights -> tensor memory, once: six chunks (plane, half of K) shared out over the warpgroups
#pragma unroll 1
        for (int ch = warp >> 2; ch < 6; ch += AT_THREADS / 128) {
            const int pl = ch >> 1, c0 = (ch & 1) * 32;
            uint32_t r[32];
            const uint4 *src = reinterpret_cast<const uint4 *>(wt + ((size_t)pl * LD + f) * 64 + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint4 v = src[j];
                r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
            }
            tmem_st32(lane_base + AT_COL_W + pl * 64 + c0, r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();

    // ---- per-attempt scalars ---------------------------------------------------------------------------------
    if (y0 == nullptr) y0 = reinterpret_cast<const float *>(c->y0_cur);
    if (k0 == nullptr) k0 = reinterpret_cast<const float *>(c->k0_cur);
    y0 = tdq_detach(y0, n_rows_sz);
    k0 = tdq_detach(k0, n_rows_sz);
    const bool fold = partials != nullptr;                                // squared error norm + candidate commit in here
    float *ycand = nullptr, *kcand = nullptr;
    if (fold && c->ybuf[0] != nullptr) {
        ycand = tdq_detach(reinterpret_cast<float *>(c->ybuf[c->par ^ 1]), n_rows_sz);
        kcand = tdq_detach(reinterpret_cast<float *>(c->kbuf[c->par ^ 1]), n_rows_sz);
    }
    // the stages of this attempt are needed afterwards only if an output time can fall into it (the controller's
    // test `!(t_out[cursor] > t1)` for t1 = att_t1, rk_common.py:246) or the caller keeps every step
    bool store = store_always != 0 || c->always_fit != 0;
    if (!store) {
        const int cur = c->out_cursor;
        store = cur < c->n_out && !(c->t_out[cur] > c->att_t1);
    }
    const float rtolT = (float)c->rtol, atolT = (float)c->atol;
    constexpr int EK = popc_below(EM, S);                                 // index of k_S's error weight in ecoef
    const float ecS = s_ce[EK];

    uint8_t *stage = smem + g * AT_STAGE;
    const uint32_t sy0 = smem_u32(smem + AT_GROUPS * AT_STAGE + g * AT_Y0) + (uint32_t)(h * RT * LD + f) * 4;   // [row][feature]: this thread's rows, its column
    const uint32_t stage_u32 = smem_u32(stage);
    const uint32_t bar = bar0 + 8 * g;
    constexpr int ACOLS = 32 * ACCS;                                       // accumulator columns per group: ACCS x (big, small)
    const uint32_t acc0 = tmem + g * ACOLS;                                // [big_0, small_0, (big_1, small_1)] x 16 columns
    uint32_t phase = 0;
    const uint32_t st_f = stage_u32 + (uint32_t)f * 16 + (uint32_t)(h * (RT / 8)) * AT_SBO;   // this thread's 16-byte slot of its first core-matrix row group

    const int tiles = (n_rows + AT_ROWS - 1) / AT_ROWS;
    const int workers = (int)gridDim.x * AT_GROUPS;
    double acc = 0.0;
    int nbad = 0;
    const uint32_t zero32 = (uint32_t)(n_rows_sz >> 63);                  // 0, unknown to the compiler

    // One tile through all S stages.  FULL: all 16 rows exist (no per-row predicates); the one partial tile of a launch
    // takes the predicated copy of the same code.
    //
    // Schedule of stage i (row i needs k_0 .. k_i, the newest one, k_i, has just come out of tensor memory):
    //   critical path   y_i = y0 + (prefix_i + k_i c_ii)  ->  split  ->  planes  ->  fence, bar.sync, 48 MMAs + commit
    //   MMA window      everything that does not depend on the product in flight: the prefix of the NEXT row's sum
    //                   (sum over j <= i of k_j c_{i+1,j}: ascending j, so the newest term is always added last and the
    //                   value is bitwise the one a single ascending loop produces), the running sums, y1's bookkeeping
    //   then            wait for the accumulators, drain them: k_{i+1}
    // Registers: k_0 .. k_{KEEP-1} are kept; once they are all known the remaining rows (and the error estimate) become
    // running sums that each later k_j is folded into as it arrives, so at most four state-sized arrays are live.
    constexpr int KEEP = S < 4 ? S : 4;
    constexpr int NACC = S - KEEP;
    auto row_mask = [](int i) -> unsigned { return (unsigned)((RM >> (8 * i)) & 0xffull); };
    auto do_tile = [&](auto full_tag, const int t) {
        constexpr bool FULL = decltype(full_tag)::value;
        const int row0 = t * AT_ROWS;
        const int rows_here = FULL ? RT : n_rows - row0 - h * RT;      // of this thread's RT rows
        const size_t base = (size_t)(row0 + h * RT) * LD + f;
        float K[KEEP][RT];
        float A[NACC > 0 ? NACC : 1][RT];                             // A[q - KEEP]: running sum of row q
        float AE[RT], PRE[RT], KN[RT], Y1[RT];
        {
            float Y0[RT];
#pragma unroll
            for (int r = 0; r < RT; ++r) {
                Y0[r] = 0.f;
                KN[r] = 0.f;
                if (FULL || r < rows_here) {
                    Y0[r] = __ldcs(y0 + base + (size_t)r * LD);
                    KN[r] = __ldcs(k0 + base + (size_t)r * LD);
                }
            }
#pragma unroll
            for (int r = 0; r < RT; ++r) sts_f32(sy0 + r * LD * 4, Y0[r]);   // only this thread reads it back: no barrier
        }
#pragma unroll
        for (int i = 0; i < S; ++i) {
            const unsigned mask = row_mask(i);
            const bool last = i == S - 1;
            const bool has_prefix = (mask & ((1u << i) - 1u)) != 0u, has_new = ((mask >> i) & 1u) != 0u;
            const float c_new = has_new ? s_cr[i * 8 + popc_below(mask, i)] : 0.f;
            if (i < KEEP) {
#pragma unroll
                for (int r = 0; r < RT; ++r) K[i < KEEP ? i : 0][r] = KN[r];
            }
            // ---- critical path: y_i, split, planes ----
#pragma unroll
            for (int r8 = 0; r8 < RT; r8 += 8) {
                uint32_t H[4], M[4], L[4];
#pragma unroll
                for (int r = r8; r < r8 + 8; r += 2) {
                    float yv[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const float pre = i < KEEP ? PRE[r + q] : A[i >= KEEP ? i - KEEP : 0][r + q];
                        float sum;
                        if (has_prefix && has_new) sum = pre + KN[r + q] * c_new;
                        else if (has_prefix) sum = pre;
                        else sum = KN[r + q] * c_new;
                        yv[q] = lds_f32(sy0 + (r + q) * LD * 4) + sum;
                        if (last) Y1[r + q] = yv[q];
                    }
                    split2(yv[0], yv[1], H[(r - r8) >> 1], M[(r - r8) >> 1], L[(r - r8) >> 1]);
                }
                sts_v4(st_f + (r8 >> 3) * AT_SBO, H[0], H[1], H[2], H[3]);
                sts_v4(st_f + AT_PLANE + (r8 >> 3) * AT_SBO, M[0], M[1], M[2], M[3]);
                sts_v4(st_f + 2 * AT_PLANE + (r8 >> 3) * AT_SBO, L[0], L[1], L[2], L[3]);
            }
            // ---- k_{i+1} = y_i W^T ----
            fence_async_smem();
            fence_before();
            asm volatile("bar.sync %0, %1;" :: "r"(g + 1), "n"(WPG * 32) : "memory");
            if (wl == g % 4 && elect_one()) {                         // group g issues from sub-partition g mod 4
                fence_after();
                // weights plane PW (tensor memory) x stage plane PY (shared memory): the five cross terms >= 2^-16 in
                // ascending magnitude into the small accumulator, hi.hi into the big one (tdq_linear.cu)
                constexpr int NPROD = 6;
                constexpr int PW[NPROD] = {1, 2, 0, 1, 0, 0}, PY[NPROD] = {1, 0, 2, 0, 1, 0};
#pragma unroll
                for (int p = 0; p < NPROD; ++p) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint32_t dcol = acc0 + (ks % ACCS) * 32 + (p == NPROD - 1 ? 0 : 16);
                        mma_ts(dcol, tmem + AT_COL_W + PW[p] * 64 + ks * 8, make_desc_mn(stage_u32 + PY[p] * AT_PLANE + ks * 2 * AT_LBO), IDESC16,
                               (p == 0 || p == NPROD - 1) && ks < ACCS ? 0u : 1u);
                    }
                }
                mma_commit(bar);
            }
            __syncwarp();
            // ---- MMA window ----
            uint32_t dep = 0;
            if (i + 1 < KEEP && i + 1 < S) {
                // prefix of the next row's sum over the slots known so far
                const unsigned mn = row_mask(i + 1);
                bool first = true;
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    if ((mn >> j) & 1u) {
                        const float cj = s_cr[(i + 1) * 8 + popc_below(mn, j)];
#pragma unroll
                        for (int r = 0; r < RT; ++r) {
                            const float p = K[j < KEEP ? j : 0][r] * cj;
                            PRE[r] = first ? p : PRE[r] + p;
                        }
                        first = false;
                    }
                }
                if (!first) dep |= dep16(PRE, zero32);
            }
            if (i == KEEP - 1) {
                // every kept slot is known: the remaining rows and the error estimate become running sums (row by row of
                // the tile, so that the kept slots die as the sums are born)
#pragma unroll
                for (int r = 0; r < RT; ++r) {
#pragma unroll
                    for (int qrow = KEEP; qrow < S; ++qrow) {
                        const unsigned mq = row_mask(qrow);
                        float a_ = 0.f;
                        bool first = true;
#pragma unroll
                        for (int j = 0; j <= i; ++j) {
                            if ((mq >> j) & 1u) {
                                const float p = K[j < KEEP ? j : 0][r] * s_cr[qrow * 8 + popc_below(mq, j)];
                                a_ = first ? p : a_ + p;
                                first = false;
                            }
                        }
                        A[qrow - KEEP][r] = a_;
                    }
                    float e_ = 0.f;
                    bool first = true;
#pragma unroll
                    for (int j = 0; j <= i; ++j) {
                        if ((EM >> j) & 1u) {
                            const float p = K[j < KEEP ? j : 0][r] * s_ce[popc_below(EM, j)];
                            e_ = first ? p : e_ + p;
                            first = false;
                        }
                    }
                    AE[r] = e_;
                }
#pragma unroll
                for (int qrow = KEEP; qrow < S; ++qrow) dep |= dep16(A[qrow - KEEP], zero32);
                dep |= dep16(AE, zero32);
            }
            if (i >= KEEP) {
                // fold k_i into the running sums of the later rows and of the error estimate
#pragma unroll
                for (int qrow = i + 1; qrow < S; ++qrow) {
                    const unsigned mq = row_mask(qrow);
                    if ((mq >> i) & 1u) {
                        const float cj = s_cr[qrow * 8 + popc_below(mq, i)];
                        const bool started = (mq & ((1u << i) - 1u)) != 0u;
#pragma unroll
                        for (int r = 0; r < RT; ++r) {
                            const float p = KN[r] * cj;
                            A[qrow - KEEP][r] = started ? A[qrow - KEEP][r] + p : p;
                        }
                    }
                }
                if ((EM >> i) & 1u) {
                    const float cj = s_ce[popc_below(EM, i)];
                    const bool started = (EM & ((1u << i) - 1u)) != 0u;
#pragma unroll
                    for (int r = 0; r < RT; ++r) {
                        const float p = KN[r] * cj;
                        AE[r] = started ? AE[r] + p : p;
                    }
                }
#pragma unroll
                for (int qrow = i + 1; qrow < S; ++qrow) dep |= dep16(A[qrow - KEEP], zero32);
                dep |= dep16(AE, zero32);
            }
            if (last) {
                // y1: non-finite count, candidate commit, (optional) y1 and the error prefix; tol = atol + rtol * max(|y0|, |y1|)
                // (misc.py:81) replaces y1 in its registers
#pragma unroll
                for (int r = 0; r < RT; ++r) {
                    const float y1v = Y1[r];
                    if (FULL || r < rows_here) {
                        if (!isfinite(y1v)) nbad += 1;
                        const size_t o = base + (size_t)r * LD;
                        if (ycand) ycand[o] = y1v;
                        if (store) {
                            out.y1[o] = y1v;
                            out.err[o] = AE[r];
                        }
                    }
                    Y1[r] = Ar<float>::add(atolT, Ar<float>::mul(rtolT, Ar<float>::max_nan(fabsf(lds_f32(sy0 + r * LD * 4)), fabsf(y1v))));
                }
                dep |= dep16(Y1, zero32);
            }
            mbar_wait(bar + dep, phase);
            phase ^= 1u;
            fence_after();
            // drain: the small accumulator first, then the big one on top of it (16 registers of staging, not 32)
            {
                uint32_t tq[RT];
#pragma unroll
                for (int a_ = 0; a_ < 2 * ACCS; ++a_) {
                    // small_0, (small_1,) big_0 (, big_1): ascending magnitude
                    const int col = (a_ < ACCS ? a_ * 32 + 16 : (a_ - ACCS) * 32) + h * RT;
                    tmem_ldn<RT>(lane_base + (uint32_t)(g * ACOLS + col), tq);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int r = 0; r < RT; ++r) KN[r] = a_ == 0 ? __uint_as_float(tq[r]) : KN[r] + __uint_as_float(tq[r]);
                }
            }
            if (store) {
#pragma unroll
                for (int r = 0; r < RT; ++r)
                    if (FULL || r < rows_here) out.k[i + 1][base + (size_t)r * LD] = KN[r];
            }
            if (last) {
                // ---- k_S: candidate commit, error ratio (misc.py:80-82 up to the mean) ----
#pragma unroll
                for (int r = 0; r < RT; ++r) {
                    if (FULL || r < rows_here) {
                        if (kcand) kcand[base + (size_t)r * LD] = KN[r];
                        if (fold) {
                            const float num = Ar<float>::add(AE[r], Ar<float>::mul(KN[r], ecS));
                            const float q = Ar<float>::div(num, Y1[r]);
                            acc += (double)Ar<float>::mul(q, q);
                        }
                    }
                }
            }
        }
    };
#pragma unroll 1
    for (int t = g * (int)gridDim.x + (int)blockIdx.x; t < tiles; t += workers) {
        {   // the next tile's y0 / k_0 rows of this warp (128 bytes each) towards L2: lanes 0..RT-1 -> y0, 16..16+RT-1 -> k_0
            const int tn = t + workers, pr = lane & 15;
            const long long prow = (long long)tn * AT_ROWS + h * RT + pr;
            if (tn < tiles && pr < RT && prow < (long long)n_rows) {
                const float *pp = (lane < 16 ? y0 : k0) + (size_t)prow * LD + e * 32;
                asm volatile("prefetch.global.L2 [%0];" :: "l"(pp));
            }
        }
        if (n_rows - t * AT_ROWS >= AT_ROWS) do_tile(std::true_type{}, t);          // (a partial tile: rows_here may be <= 0 for h = 1)
        else do_tile(std::false_type{}, t);
    }
    const double bad = (double)nbad;

    // ---- per-CTA partial of the squared norm and of the non-finite count; the last CTA adds them in index order ----
    if (fold) {
        const double wa = warp_sum(acc), wb = warp_sum(bad);
        if (lane == 0) {
            s_red[warp] = wa;
            s_red[32 + warp] = wb;
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(AT_TMEM_COLS) : "memory");
    }
    if (!fold) return;
    const int P = (int)gridDim.x;
    double *p_sum = partials + 2, *p_bad = p_sum + P;
    unsigned int *ticket = reinterpret_cast<unsigned int *>(partials);
    if (tid == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < AT_THREADS / 32; ++w) {
            a += s_red[w];
            b += s_red[32 + w];
        }
        p_sum[blockIdx.x] = a;
        p_bad[blockIdx.x] = b;
        __threadfence();
        const unsigned int tk = atomicAdd(ticket, 1u);
        *s_flag = (tk == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!*s_flag) return;
    __threadfence();
    if (warp == 0) {
        double a = 0.0, b = 0.0;
        for (int i = lane; i < P; i += 32) {
            a += __ldcg(&p_sum[i]);
            b += __ldcg(&p_bad[i]);
        }
        a = warp_sum(a);
        b = warp_sum(b);
        if (lane == 0) {
            norm_out[0] = a;
            norm_out[1] = b;
            *ticket = 0;                                                  // self-reset for the next launch
        }
    }
    if (CTRL) {
        __threadfence();
        __syncthreads();
        tdq_ctrl_dev::controller_block<float, AT_THREADS>(c, norm_out, seg_counts, 1, nullptr);
    }
}

// the sparsity of the supported FSAL tableaus (row masks 8 bits per row, error-prefix mask); tsit5 as the reference
// tabulates it is not FSAL (its c_sol has a weight on k_S), fehlberg2 / adaptive_heun neither, dopri8 has 13 stages
constexpr unsigned long long RM_DOPRI5 = 0x01ull | (0x03ull << 8) | (0x07ull << 16) | (0x0full << 24) | (0x1full << 32) | (0x3dull << 40);
constexpr unsigned long long RM_BOSH3 = 0x01ull | (0x02ull << 8) | (0x07ull << 16);

template <int S, unsigned long long RM, unsigned EM, int G, int RT, bool CTRL, int ACCS>
int launch_attempt_g(TdqCtrl *c, const float *y0, const float *k0, const AttOut &out, const uint32_t *wt, double *partials,
                     double *norm_out, const int64_t *seg_counts, int store_always, size_t n_rows, cudaStream_t st) {
    auto kern = k_linear_attempt<S, RM, EM, G, RT, CTRL, ACCS>;
    constexpr int AT_SMEM = at_smem(G);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM) != cudaSuccess) return -2;
    const size_t tiles = (n_rows + AT_ROWS - 1) / AT_ROWS;
    size_t grid = (tiles + G - 1) / G;
    const size_t cap = (size_t)tdq_sm_count();
    if (grid > cap) grid = cap;
    if (grid == 0) grid = 1;
    kern<<<(unsigned)grid, G * 128 * (AT_ROWS / RT), AT_SMEM, st>>>(c, y0, k0, out, wt, partials, norm_out, seg_counts, store_always, n_rows);
    return 0;
}

template <int S, unsigned long long RM, unsigned EM>
int launch_attempt(TdqCtrl *c, const float *y0, const float *k0, const AttOut &out, const uint32_t *wt, double *partials,
                   double *norm_out, const int64_t *seg_counts, int store_always, size_t n_rows, cudaStream_t st) {
    // Four tile pipelines of one warpgroup each, 16 rows per thread (128 registers).  Measured alternatives at 65,536 x 128,
    // dopri5 (profiles/README.md): 3 pipelines 106.6 us, 4: 96.1, 5 (96 registers, spills): 103.4; two warpgroups per tile
    // (8 rows per thread) 3 x 2: 102.2, 4 x 2: 102.0.  Two accumulator pairs per tile (ACCS = 2: consecutive MMAs never
    // accumulate into the same tensor-memory tile): 97.4 against 92.3 -- the MMA chain is not what paces a stage.
#define TDQ_GO(G_, RT_, A_) (seg_counts ? launch_attempt_g<S, RM, EM, G_, RT_, true, A_>(c, y0, k0, out, wt, partials, norm_out, seg_counts, store_always, n_rows, st) \
                                        : launch_attempt_g<S, RM, EM, G_, RT_, false, A_>(c, y0, k0, out, wt, partials, norm_out, seg_counts, store_always, n_rows, st))
    return TDQ_GO(4, 16, 1);
#undef TDQ_GO
}

// row / error masks of a tableau, or false if it is not FSAL with 2..AT_MAX_S stages and an error weight on k_S
bool attempt_masks(const TdqHostShape &hs, unsigned long long *rm, unsigned *em) {
    const int S = hs.n_stages;
    if (!hs.fsal || S < 2 || S > AT_MAX_S) return false;
    *rm = 0;
    *em = 0;
    for (int i = 0; i < S; ++i) {
        unsigned m = 0;
        for (int q = 0; q < hs.row_nnz[i]; ++q) {
            if (hs.row_idx[i][q] > i) return false;
            m |= 1u << hs.row_idx[i][q];
        }
        if (m == 0) return false;
        *rm |= (unsigned long long)m << (8 * i);
    }
    bool has_S = false;
    for (int q = 0; q < hs.err_nnz; ++q) {
        if (hs.err_idx[q] == S) has_S = true;
        else *em |= 1u << hs.err_idx[q];
    }
    return has_S && hs.err_idx[hs.err_nnz - 1] == S;
}

}  // namespace

extern "C" {

int tdq_linear_attempt_supported(const tdq_tableau *tab, int32_t dtype, int32_t width) {
    if (!tab || dtype != TDQ_F32 || width != LD) return 0;
    TdqHostShape hs;
    tdq_shape_from_tableau(tab, &hs);
    unsigned long long rm;
    unsigned em;
    if (!attempt_masks(hs, &rm, &em)) return 0;
    const int S = hs.n_stages;
    return ((S == 6 && rm == RM_DOPRI5 && em == 0x3du) || (S == 3 && rm == RM_BOSH3 && em == 0x07u)) ? 1 : 0;
}

int tdq_linear_attempt(void *ctrl_dev, const tdq_tableau *tab, int32_t dtype, void *const *k_out, void *y1_out, void *err_out,
                       const void *y0, const void *k0, const void *planes, int32_t width, size_t n, double *partials,
                       double *norm_out, const int64_t *seg_counts_dev, int32_t store_always, void *stream) {
    TDQ_REQUIRE(ctrl_dev && tab && k_out && y1_out && err_out && planes, "null argument");
    TDQ_REQUIRE(dtype == TDQ_F32 && width == LD, "the fused linear field is float32, width 128");
    TDQ_REQUIRE(n % (size_t)width == 0, "state size is not a multiple of the field width");
    TDQ_REQUIRE((partials == nullptr) == (norm_out == nullptr), "partials and norm_out go together");
    TDQ_REQUIRE(seg_counts_dev == nullptr || partials != nullptr, "the controller step needs the folded error norm");
    const size_t n_rows = n / (size_t)width;
    TDQ_REQUIRE(n_rows < ((size_t)1 << 31) - 64, "too many rows");
    TdqHostShape hs;
    tdq_shape_from_tableau(tab, &hs);
    unsigned long long rm = 0;
    unsigned em = 0;
    TDQ_REQUIRE(attempt_masks(hs, &rm, &em), "the whole-attempt kernel takes FSAL tableaus of at most 7 stages");
    const int S = hs.n_stages;
    AttOut out;
    memset(&out, 0, sizeof(out));
    for (int i = 1; i <= S; ++i) {
        TDQ_REQUIRE(k_out[i] != nullptr, "missing stage slot");
        out.k[i] = (float *)k_out[i];
    }
    out.y1 = (float *)y1_out;
    out.err = (float *)err_out;
    if (n_rows == 0) return TDQ_OK;
    TdqCtrl *c = (TdqCtrl *)ctrl_dev;
    const uint32_t *wt = (const uint32_t *)planes;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = -1;
    if (S == 6 && rm == RM_DOPRI5 && em == 0x3du)
        rc = launch_attempt<6, RM_DOPRI5, 0x3du>(c, (const float *)y0, (const float *)k0, out, wt, partials, norm_out, seg_counts_dev, store_always, n_rows, st);
    else if (S == 3 && rm == RM_BOSH3 && em == 0x07u)
        rc = launch_attempt<3, RM_BOSH3, 0x07u>(c, (const float *)y0, (const float *)k0, out, wt, partials, norm_out, seg_counts_dev, store_always, n_rows, st);
    TDQ_REQUIRE(rc != -1, "no whole-attempt kernel for this tableau (tdq_linear_attempt_supported)");
    TDQ_REQUIRE(rc == 0, "launch configuration failed");
    TDQ_CHECK_CUDA(cudaGetLastError());
    return TDQ_OK;
}

}  // extern "C"
