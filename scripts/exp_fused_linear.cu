// Experiment (not part of libtdq): a Runge-Kutta stage fused with a LINEAR vector field on the 5th-generation tensor cores.
//
//   y_i = y0 + sum_j cf_j * k_j          (the stage combination of tdq_stream.cu, same expression order)
//   k_i = y_i @ A^T                      (the field of BASELINE.json configs[1]: 65536 x 128 states, A 128 x 128)
//
// The float32 product is computed as a BF16x9 emulation (each float32 operand = hi + mid + lo bfloat16 planes, exact to 24 bits;
// nine bf16 products accumulated in float32 in tensor memory), the scheme cuBLAS 12.9 offers as
// CUBLAS_COMPUTE_32F_EMULATED_16BFX9.  y_i never goes to HBM: it is split in registers and stored as three bf16 operand planes
// in shared memory (128-byte swizzle, K-major), tcgen05.mma reads them, tcgen05.ld brings the accumulator tile back.
//
// build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -lineinfo -o /tmp/exp_fused_linear scripts/exp_fused_linear.cu
// run:    /tmp/exp_fused_linear [rows]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int D = 128;                 // state width = GEMM N = GEMM K
constexpr int TM = 128;                // rows per tile = GEMM M
constexpr int ATOM_BYTES = 128 * 128;  // one swizzle atom column: 128 rows x 128 bytes (64 bf16 of K)
constexpr int PLANE_BYTES = 2 * ATOM_BYTES;   // K = 128 bf16 = two atoms
constexpr int MAXK = 7;

struct KP { const float *p[MAXK]; };
struct CF { float c[MAXK]; };

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- float32 -> three bf16 planes, two elements at a time (packed: element 0 in the low half) --------------------------------
__device__ __forceinline__ void split2(float a, float b, uint32_t &h, uint32_t &m, uint32_t &l) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));
    const float ra = a - __uint_as_float(h << 16), rb = b - __uint_as_float(h & 0xffff0000u);       // exact
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(m) : "f"(rb), "f"(ra));
    const float sa = ra - __uint_as_float(m << 16), sb = rb - __uint_as_float(m & 0xffff0000u);     // exact
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(sb), "f"(sa));
}

// byte offset, inside one plane, of the 8-byte group holding elements [4*q, 4*q+4) of row r (q = 0..31):
// K-major, 128-byte swizzle: atoms of 8 rows x 128 bytes, 16-byte chunk index XOR (row mod 8); rows 128 B apart,
// the second 64 elements of K one ATOM (16 KB) further
__device__ __forceinline__ uint32_t plane_offset(int r, int q) {
    const int katom = q >> 4, chunk = (q & 15) >> 1, half = q & 1;
    return katom * ATOM_BYTES + r * 128 + ((chunk ^ (r & 7)) << 4) + half * 8;
}

template <typename P>
__device__ __forceinline__ void store_split(P *base, int r, int q, float4 v) {
    uint32_t h0, m0, l0, h1, m1, l1;
    split2(v.x, v.y, h0, m0, l0);
    split2(v.z, v.w, h1, m1, l1);
    const uint32_t off = plane_offset(r, q);
    *reinterpret_cast<uint2 *>(base + off) = make_uint2(h0, h1);
    *reinterpret_cast<uint2 *>(base + PLANE_BYTES + off) = make_uint2(m0, m1);
    *reinterpret_cast<uint2 *>(base + 2 * PLANE_BYTES + off) = make_uint2(l0, l1);
}

// A (128 x 128 float32, row n = output column n, K contiguous) -> three pre-swizzled bf16 planes in global memory
__global__ void k_split_weights(const float *__restrict__ A, uint8_t *__restrict__ planes) {
    const int r = blockIdx.x, q = threadIdx.x;          // 128 blocks x 32 threads
    const float4 v = *reinterpret_cast<const float4 *>(A + (size_t)r * D + q * 4);
    store_split(planes, r, q, v);
}

// ---- tcgen05 / mbarrier wrappers ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    // UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start address >> 4 in [0,14), leading byte
    // offset >> 4 in [16,30) (unused for swizzled K-major), stride byte offset >> 4 in [32,46) = 1024 B between 8-row groups,
    // version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64)
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (InstrDescriptor): D = F32 (1 at [4,6)), A = B = BF16 (1 at [7,10), [10,13)), both K-major,
// N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- the fused stage ----------------------------------------------------------------------------------------------------------
constexpr int THREADS = 256;
constexpr int TMEM_COLS = 128;
constexpr int SMEM_BYTES = 6 * PLANE_BYTES + 1024 + 64;

template <int NK>
__global__ void __launch_bounds__(THREADS, 1)
k_linear_stage(const float *__restrict__ y0, KP kp, CF cf, const uint8_t *__restrict__ wplanes, float *__restrict__ kout,
               int ntiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sB = smem, *sA = smem + 3 * PLANE_BYTES;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 6 * PLANE_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 6 * PLANE_BYTES + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(smem_u32(bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 3 * PLANE_BYTES / 16; i += THREADS)
        reinterpret_cast<uint4 *>(sB)[i] = reinterpret_cast<const uint4 *>(wplanes)[i];
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_d = *tmem_slot;
    const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB), bar_a = smem_u32(bar);

    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const size_t row0 = (size_t)tile * TM;
        // ---- stage combination, split, operand planes ----
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
            float4 a[4], kv[4][NK > 0 ? NK : 1];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = warp + 8 * (4 * i + u);
                const size_t off = (row0 + r) * D + lane * 4;
                a[u] = __ldcs(reinterpret_cast<const float4 *>(y0 + off));
#pragma unroll
                for (int m = 0; m < NK; ++m) kv[u][m] = __ldcs(reinterpret_cast<const float4 *>(kp.p[m] + off));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = warp + 8 * (4 * i + u);
                float4 y = a[u];
                if (NK > 0) {
                    float4 acc = make_float4(kv[u][0].x * cf.c[0], kv[u][0].y * cf.c[0], kv[u][0].z * cf.c[0], kv[u][0].w * cf.c[0]);
#pragma unroll
                    for (int m = 1; m < NK; ++m) {
                        acc.x = acc.x + kv[u][m].x * cf.c[m];
                        acc.y = acc.y + kv[u][m].y * cf.c[m];
                        acc.z = acc.z + kv[u][m].z * cf.c[m];
                        acc.w = acc.w + kv[u][m].w * cf.c[m];
                    }
                    y = make_float4(y.x + acc.x, y.y + acc.y, y.z + acc.z, y.w + acc.w);
                }
                store_split(sA, r, lane, y);
            }
        }
        fence_async_smem();
        fence_before();
        __syncthreads();
        // ---- nine bf16 products, smallest first, accumulated in tensor memory ----
        if (tid == 0) {
            fence_after();
            constexpr int PA[9] = {2, 2, 1, 2, 0, 1, 1, 0, 0};
            constexpr int PB[9] = {2, 1, 2, 0, 2, 1, 0, 1, 0};
            uint32_t acc = 0;
#pragma unroll
            for (int p = 0; p < 9; ++p) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint32_t koff = (ks >> 2) * ATOM_BYTES + (ks & 3) * 32;
                    mma_bf16(tmem_d, make_desc(a_base + PA[p] * PLANE_BYTES + koff), make_desc(b_base + PB[p] * PLANE_BYTES + koff), acc);
                    acc = 1;
                }
            }
            mma_commit(bar_a);
        }
        mbar_wait(bar_a, it & 1);
        fence_after();
        // ---- accumulator tile -> k_i ----
        {
            const int q = warp & 3, h = warp >> 2;
            float *dst = kout + (row0 + q * 32 + lane) * D + h * 64;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                uint32_t r[32];
                tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + h * 64 + cc * 32, r);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4 *>(dst + cc * 32 + j * 4) =
                        make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
            }
        }
        fence_before();
        __syncthreads();
    }
    if (warp == 0) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "n"(TMEM_COLS) : "memory");
    }
}

// ---- version 2: balanced 64-row units, K-atom split MMA issue, register double buffering with prefetch across tiles,
//      accumulator tile transposed through shared memory for coalesced stores, optional last-row outputs (y1, error prefix) ----
constexpr int SCRATCH_ROW = 272;                       // 64 floats + 16 bytes: conflict-free 128-bit writes down a column
constexpr int SCRATCH_WARP = 32 * SCRATCH_ROW;

template <int NK, int NPROD, bool FINAL>
__global__ void __launch_bounds__(THREADS, 1)
k_linear_stage2(const float *__restrict__ y0, KP kp, CF cr, CF ce, const uint8_t *__restrict__ wplanes,
                float *__restrict__ kout, float *__restrict__ yout, float *__restrict__ eout, int n_rows) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sB = smem, *sA = smem + 3 * PLANE_BYTES;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 6 * PLANE_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 6 * PLANE_BYTES + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, half = lane >> 4, q = lane & 15;
    constexpr int NKK = NK > 0 ? NK : 1;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(smem_u32(bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 3 * PLANE_BYTES / 16; i += THREADS)
        reinterpret_cast<uint4 *>(sB)[i] = reinterpret_cast<const uint4 *>(wplanes)[i];
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_d = *tmem_slot;
    const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB), bar_a = smem_u32(bar);

    struct Regs { float4 a[2]; float4 k[2][NKK]; };
    Regs R[2];
    // batch b (0..3) of K-atom `atom`: rows (2b+u)*16 + 2*warp + half, u = 0,1; 16 lanes cover the 64 columns of the atom
    auto load = [&](Regs &r, int row0, int rows_here, int atom, int b) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int rr = (b * 2 + u) * 16 + warp * 2 + half;
            if (rr < rows_here) {
                const size_t off = (size_t)(row0 + rr) * D + atom * 64 + q * 4;
                r.a[u] = __ldcs(reinterpret_cast<const float4 *>(y0 + off));
#pragma unroll
                for (int m = 0; m < NK; ++m) r.k[u][m] = __ldcs(reinterpret_cast<const float4 *>(kp.p[m] + off));
            }
        }
    };
    auto process = [&](Regs &r, int row0, int rows_here, int atom, int b) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int rr = (b * 2 + u) * 16 + warp * 2 + half;
            if (rr < rows_here) {
                float4 y = r.a[u];
                if (NK > 0) {
                    float4 acc = make_float4(r.k[u][0].x * cr.c[0], r.k[u][0].y * cr.c[0], r.k[u][0].z * cr.c[0], r.k[u][0].w * cr.c[0]);
#pragma unroll
                    for (int m = 1; m < NK; ++m) {
                        acc.x = acc.x + r.k[u][m].x * cr.c[m];
                        acc.y = acc.y + r.k[u][m].y * cr.c[m];
                        acc.z = acc.z + r.k[u][m].z * cr.c[m];
                        acc.w = acc.w + r.k[u][m].w * cr.c[m];
                    }
                    y = make_float4(y.x + acc.x, y.y + acc.y, y.z + acc.z, y.w + acc.w);
                    if (FINAL) {
                        float4 e = make_float4(r.k[u][0].x * ce.c[0], r.k[u][0].y * ce.c[0], r.k[u][0].z * ce.c[0], r.k[u][0].w * ce.c[0]);
#pragma unroll
                        for (int m = 1; m < NK; ++m) {
                            e.x = e.x + r.k[u][m].x * ce.c[m];
                            e.y = e.y + r.k[u][m].y * ce.c[m];
                            e.z = e.z + r.k[u][m].z * ce.c[m];
                            e.w = e.w + r.k[u][m].w * ce.c[m];
                        }
                        const size_t off = (size_t)(row0 + rr) * D + atom * 64 + q * 4;
                        *reinterpret_cast<float4 *>(yout + off) = y;
                        *reinterpret_cast<float4 *>(eout + off) = e;
                    }
                }
                store_split(sA, rr, atom * 16 + q, y);
            }
        }
    };
    auto issue = [&](int atom, bool first) {
        constexpr int PA9[9] = {2, 2, 1, 2, 0, 1, 1, 0, 0}, PB9[9] = {2, 1, 2, 0, 2, 1, 0, 1, 0};
        constexpr int PA6[6] = {2, 0, 1, 1, 0, 0}, PB6[6] = {0, 2, 1, 0, 1, 0};
        uint32_t acc = first ? 0u : 1u;
#pragma unroll
        for (int p = 0; p < NPROD; ++p) {
            const int pa = NPROD == 9 ? PA9[p] : PA6[p], pb = NPROD == 9 ? PB9[p] : PB6[p];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint32_t koff = atom * ATOM_BYTES + ks * 32;
                mma_bf16(tmem_d, make_desc(a_base + pa * PLANE_BYTES + koff), make_desc(b_base + pb * PLANE_BYTES + koff), acc);
                acc = 1;
            }
        }
    };

    const int units = (n_rows + 63) / 64;
    int u = (int)((long long)blockIdx.x * units / gridDim.x);
    const int u_end = (int)((long long)(blockIdx.x + 1) * units / gridDim.x);
    uint32_t it = 0;
    if (u < u_end) {
        const int nu = u_end - u < 2 ? u_end - u : 2;
        const int rows_here = min(nu * 64, n_rows - u * 64);
        load(R[0], u * 64, rows_here, 0, 0);
    }
    while (u < u_end) {
        const int nu = u_end - u < 2 ? u_end - u : 2;
        const int row0 = u * 64, rows_here = min(nu * 64, n_rows - row0);
        const int un = u + nu;
        const bool has_next = un < u_end;
        const int nnu = u_end - un < 2 ? u_end - un : 2;
        const int nrow0 = un * 64, nrows_here = has_next ? min(nnu * 64, n_rows - nrow0) : 0;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            if (s < 7) load(R[(s + 1) & 1], row0, rows_here, (s + 1) >> 2, (s + 1) & 3);
            else if (has_next) load(R[0], nrow0, nrows_here, 0, 0);
            process(R[s & 1], row0, rows_here, s >> 2, s & 3);
            if (s == 3 || s == 7) {
                fence_async_smem();
                fence_before();
                __syncthreads();
                if (tid == 0) {
                    fence_after();
                    issue(s >> 2, s == 3);
                    if (s == 7) mma_commit(bar_a);
                }
            }
        }
        mbar_wait(bar_a, it & 1);
        fence_after();
        {
            const int qw = warp & 3, h = warp >> 2;
            if (qw * 32 < rows_here) {
                uint8_t *scratch = sA + warp * SCRATCH_WARP;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    uint32_t r[32];
                    tmem_ld32(tmem_d + ((uint32_t)(qw * 32) << 16) + h * 64 + cc * 32, r);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<uint4 *>(scratch + lane * SCRATCH_ROW + (cc * 32 + j * 4) * 4) =
                            make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int rr = 2 * i + half, grow = qw * 32 + rr;
                    const uint4 v = *reinterpret_cast<const uint4 *>(scratch + rr * SCRATCH_ROW + q * 16);
                    if (grow < rows_here) *reinterpret_cast<uint4 *>(kout + (size_t)(row0 + grow) * D + h * 64 + q * 4) = v;
                }
            }
        }
        fence_before();
        __syncthreads();
        u = un;
        ++it;
    }
    if (warp == 0) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "n"(TMEM_COLS) : "memory");
    }
}

// ---- version 3: warp specialised; the WEIGHTS are the stationary A operand in tensor memory (M = 128 output features),
//      the stage values are the B operand (N = 128 rows) in a two-deep shared-memory ring; the accumulator comes out
//      transposed (lane = feature, column = row), so the epilogue stores are 128-byte coalesced without staging. ----
constexpr int V3_THREADS = 384;
constexpr int STAGE_BYTES = 3 * PLANE_BYTES;
constexpr int V3_SMEM = 2 * STAGE_BYTES + 1024 + 128;
constexpr int V3_TMEM_COLS = 512;
constexpr int COL_BIG = 0, COL_SMALL = 128, COL_W = 256;      // tensor-memory columns

// weights -> [plane][feature n][64 x u32] (bf16 pairs along K): what thread n stores into its tensor-memory lane
__global__ void k_split_weights_t(const float *__restrict__ A, uint32_t *__restrict__ wt) {
    const int n = blockIdx.x, c = threadIdx.x;            // 128 blocks x 64 threads
    uint32_t h, m, l;
    split2(A[(size_t)n * D + 2 * c], A[(size_t)n * D + 2 * c + 1], h, m, l);
    wt[(0 * D + n) * 64 + c] = h;
    wt[(1 * D + n) * 64 + c] = m;
    wt[(2 * D + n) * 64 + c] = l;
}

__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "r"(tmem_a), "l"(db), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                    "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
                    "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
                    "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}

template <int NK, bool FINAL, int PF = 0, int SCHED = 0, int PIPE = 0, int NPROD = 9>
__global__ void __launch_bounds__(V3_THREADS, 1)
k_linear_stage3(const float *__restrict__ y0, KP kp, CF cr, CF ce, const uint32_t *__restrict__ wt,
                float *__restrict__ kout, float *__restrict__ yout, float *__restrict__ eout, int n_rows) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 2 * STAGE_BYTES);   // full[2], empty[2], accf, acce, wready
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 2 * STAGE_BYTES + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wg = warp >> 2;
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t b_full[2] = {bar0, bar0 + 8}, b_empty[2] = {bar0 + 16, bar0 + 24};
    const uint32_t b_accf = bar0 + 32, b_acce = bar0 + 40, b_wready = bar0 + 48;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(V3_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(b_full[0], 256); mbar_init(b_full[1], 256);
        mbar_init(b_empty[0], 1); mbar_init(b_empty[1], 1);
        mbar_init(b_accf, 1); mbar_init(b_acce, 128); mbar_init(b_wready, 128); mbar_init(bar0 + 56, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;
    const int dbg = (int)ce.c[6];          // ablation only: 1 skip epilogue stores, 2 skip MMAs, 4 skip y0 loads, 8 skip split+STS

    const int units = (n_rows + 63) / 64;
    const int u_begin = (int)((long long)blockIdx.x * units / gridDim.x);
    const int u_end = (int)((long long)(blockIdx.x + 1) * units / gridDim.x);
    // SCHED 1: round robin -- W full waves of 128-row tiles (tile it * G + b), then the remaining 64-row units one per CTA and
    // wave: at any time the CTAs read ONE contiguous window of every array (like a grid-stride streaming kernel)
    const int G = gridDim.x, W_full = (units / 2) / G, u_tail = 2 * W_full * G;
    const int n_tail = units - u_tail > (int)blockIdx.x ? (units - u_tail - (int)blockIdx.x + G - 1) / G : 0;
    const int n_tiles = SCHED == 0 ? (u_end - u_begin + 1) / 2 : W_full + n_tail;
    auto tile_of = [&](int it, int &row0, int &rows_here) {
        if (SCHED == 0) {
            const int u = u_begin + 2 * it;
            row0 = u * 64;
            rows_here = min(min(128, (u_end - u) * 64), n_rows - row0);
        } else if (it < W_full) {
            row0 = (it * G + (int)blockIdx.x) * 128;
            rows_here = min(128, n_rows - row0);
        } else {
            row0 = (u_tail + (it - W_full) * G + (int)blockIdx.x) * 64;
            rows_here = min(64, n_rows - row0);
        }
    };

    if (wg < 2) {
        // ================= producers: stage combination, split, operand planes =================
        constexpr int NKK = NK > 0 ? NK : 1;
        constexpr int BPT = 4;                                   // batches of 4 rows per warp and tile
        const int nb = n_tiles * BPT;
        // L2 prefetch of a later batch: lane l < 4 * (NK + 1) fetches the 512-byte row uu = l / (NK + 1) of array l % (NK + 1)
        auto prefetch = [&](int b) {
            int row0, rows_here;
            tile_of(b / BPT, row0, rows_here);
            const int uu = lane / (NK + 1), m = lane % (NK + 1);
            const int r = warp + 8 * (4 * (b % BPT) + uu);
            if (uu < 4 && r < rows_here) {
                const float *base = y0;
#pragma unroll
                for (int mm = 0; mm < NK; ++mm) if (m == mm + 1) base = kp.p[mm];
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], 512;" :: "l"(base + (size_t)(row0 + r) * D) : "memory");
            }
        };
        if (PF > 0)
            for (int b = 0; b < PF && b < nb; ++b) prefetch(b);
#pragma unroll 1
        for (int b = 0; b < nb; ++b) {
            const int it = b / BPT, i = b % BPT, s = it & 1;
            int row0, rows_here;
            tile_of(it, row0, rows_here);
            uint8_t *sY = smem + s * STAGE_BYTES;
            if (PF > 0 && b + PF < nb) prefetch(b + PF);
            if (i == 0) mbar_wait(b_empty[s], ((it >> 1) & 1) ^ 1);
            if (4 * i * 8 < rows_here) {
                float4 a[4], kv[4][NKK];
#pragma unroll
                for (int uu = 0; uu < 4; ++uu) {
                    const int r = warp + 8 * (4 * i + uu);
                    if (r < rows_here) {
                        const size_t off = (size_t)(row0 + r) * D + lane * 4;
                        if (dbg & 4) { a[uu] = make_float4(1.f, 2.f, 3.f, (float)off); }
                        else a[uu] = __ldcs(reinterpret_cast<const float4 *>(y0 + off));
#pragma unroll
                        for (int m = 0; m < NK; ++m) kv[uu][m] = __ldcs(reinterpret_cast<const float4 *>(kp.p[m] + off));
                    }
                }
#pragma unroll
                for (int uu = 0; uu < 4; ++uu) {
                    const int r = warp + 8 * (4 * i + uu);
                    if (r < rows_here) {
                        float4 y = a[uu];
                        if (NK > 0) {
                            float4 acc = make_float4(kv[uu][0].x * cr.c[0], kv[uu][0].y * cr.c[0], kv[uu][0].z * cr.c[0], kv[uu][0].w * cr.c[0]);
#pragma unroll
                            for (int m = 1; m < NK; ++m) {
                                acc.x = acc.x + kv[uu][m].x * cr.c[m];
                                acc.y = acc.y + kv[uu][m].y * cr.c[m];
                                acc.z = acc.z + kv[uu][m].z * cr.c[m];
                                acc.w = acc.w + kv[uu][m].w * cr.c[m];
                            }
                            y = make_float4(y.x + acc.x, y.y + acc.y, y.z + acc.z, y.w + acc.w);
                            if (FINAL) {
                                float4 e = make_float4(kv[uu][0].x * ce.c[0], kv[uu][0].y * ce.c[0], kv[uu][0].z * ce.c[0], kv[uu][0].w * ce.c[0]);
#pragma unroll
                                for (int m = 1; m < NK; ++m) {
                                    e.x = e.x + kv[uu][m].x * ce.c[m];
                                    e.y = e.y + kv[uu][m].y * ce.c[m];
                                    e.z = e.z + kv[uu][m].z * ce.c[m];
                                    e.w = e.w + kv[uu][m].w * ce.c[m];
                                }
                                const size_t off = (size_t)(row0 + r) * D + lane * 4;
                                *reinterpret_cast<float4 *>(yout + off) = y;
                                *reinterpret_cast<float4 *>(eout + off) = e;
                            }
                        }
                        if (!(dbg & 8)) store_split(sY, r, lane, y);
                    }
                }
            }
            if (i == BPT - 1) {
                fence_async_smem();
                mbar_arrive(b_full[s]);
            }
        }
    } else if (wg == 2) {
        // ================= epilogue warps: weights into tensor memory once, then accumulator tiles -> k =================
        const int e = warp & 3, f = e * 32 + lane;                // this thread's tensor-memory lane = output feature
        const uint32_t lane_base = tmem + ((uint32_t)(e * 32) << 16);
#pragma unroll 1
        for (int pl = 0; pl < 3; ++pl) {
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 32) {
                uint32_t r[32];
                const uint4 *src = reinterpret_cast<const uint4 *>(wt + ((size_t)pl * D + f) * 64 + c0);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint4 v = src[j];
                    r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
                }
                tmem_st32(lane_base + COL_W + pl * 64 + c0, r);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if (PIPE) {
            // ---- version 4: one accumulator per tile, two of them: the MMAs of tile t+1 run while tile t is drained ----
            const uint32_t b_acc2[2] = {b_accf, bar0 + 56};
            const bool leader = warp == 8 && lane == 0;
            auto issue = [&](int t) {
                const int s = t & 1;
                const uint32_t y_base = smem_u32(smem + s * STAGE_BYTES);
                mbar_wait(b_full[s], (t >> 1) & 1);
                fence_after();
                constexpr int PW[9] = {2, 2, 1, 1, 2, 0, 1, 0, 0}, PY[9] = {2, 1, 2, 1, 0, 2, 0, 1, 0};
#pragma unroll
                for (int p = 0; p < 9; ++p) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint32_t koff = (ks >> 2) * ATOM_BYTES + (ks & 3) * 32;
                        mma_bf16_ts(tmem + s * 128, tmem + COL_W + PW[p] * 64 + ks * 8, make_desc(y_base + PY[p] * PLANE_BYTES + koff),
                                    p == 0 && ks == 0 ? 0u : 1u);
                    }
                }
                mma_commit(b_empty[s]);
                mma_commit(b_acc2[s]);
            };
            fence_before();
            asm volatile("bar.sync 1, 128;" ::: "memory");          // the weights are in tensor memory
            if (leader && n_tiles > 0) issue(0);
            for (int t = 0; t < n_tiles; ++t) {
                int row0, rows_here;
                tile_of(t, row0, rows_here);
                if (t > 0) {
                    fence_before();
                    asm volatile("bar.sync 1, 128;" ::: "memory");  // tile t-1 is drained: its accumulator is free for t+1
                }
                if (leader && t + 1 < n_tiles) issue(t + 1);
                __syncwarp();
                mbar_wait(b_acc2[t & 1], (t >> 1) & 1);
                fence_after();
                const uint32_t acc = lane_base + (t & 1) * 128;
                const int nch = (rows_here + 15) >> 4;
                uint32_t buf[2][16];
                tmem_ld16_nowait(acc, buf[0]);
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    if (cc < nch) {
                        tmem_ld_wait();
                        if (cc + 1 < nch) tmem_ld16_nowait(acc + (cc + 1) * 16, buf[(cc + 1) & 1]);
                        float *dst = kout + (size_t)(row0 + cc * 16) * D + f;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (cc * 16 + j < rows_here) dst[(size_t)j * D] = __uint_as_float(buf[cc & 1][j]);
                    }
                }
            }
        } else
        for (uint32_t it = 0; it < (uint32_t)n_tiles; ++it) {
            int row0, rows_here;
            tile_of((int)it, row0, rows_here);
            // the previous tile's accumulators are drained (and, the first time, the weights are in tensor memory)
            fence_before();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (warp == 8 && lane == 0) {
                // ---- MMA issue: one thread ----
                const int s = it & 1;
                const uint32_t y_base = smem_u32(smem + s * STAGE_BYTES);
                mbar_wait(b_full[s], (it >> 1) & 1);
                fence_after();
                // weights plane pw (A, tensor memory) x stage-value plane py (B, shared memory); eight small products
                // in ascending magnitude into one accumulator, hi x hi into the other
                constexpr int PW9[9] = {2, 2, 1, 1, 2, 0, 1, 0, 0}, PY9[9] = {2, 1, 2, 1, 0, 2, 0, 1, 0};
                constexpr int PW6[9] = {1, 2, 0, 1, 0, 0, 0, 0, 0}, PY6[9] = {1, 0, 2, 0, 1, 0, 0, 0, 0};
                constexpr int PW3[9] = {1, 0, 0, 0, 0, 0, 0, 0, 0}, PY3[9] = {0, 1, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
                for (int p = 0; p < NPROD; ++p) {
                    const int *PW = NPROD == 9 ? PW9 : (NPROD == 6 ? PW6 : PW3), *PY = NPROD == 9 ? PY9 : (NPROD == 6 ? PY6 : PY3);
                    const uint32_t dcol = tmem + (p == NPROD - 1 ? COL_BIG : COL_SMALL);
                    if (dbg & 2) continue;
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint32_t koff = (ks >> 2) * ATOM_BYTES + (ks & 3) * 32;
                        mma_bf16_ts(dcol, tmem + COL_W + PW[p] * 64 + ks * 8, make_desc(y_base + PY[p] * PLANE_BYTES + koff),
                                    (p == 0 || p == NPROD - 1) && ks == 0 ? 0u : 1u);
                    }
                }
                mma_commit(b_empty[s]);
                mma_commit(b_accf);
            }
            __syncwarp();
            mbar_wait(b_accf, it & 1);
            fence_after();
#pragma unroll 1
            for (int cc = 0; cc < 8; ++cc) {
                if (cc * 16 >= rows_here) break;
                uint32_t big[16], small[16];
                tmem_ld16_nowait(lane_base + COL_SMALL + cc * 16, small);
                tmem_ld16_nowait(lane_base + COL_BIG + cc * 16, big);
                tmem_ld_wait();
                float *dst = kout + (size_t)(row0 + cc * 16) * D + f;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (cc * 16 + j < rows_here && !((dbg & 1) && j > 0)) dst[(size_t)j * D] = __uint_as_float(small[j]) + __uint_as_float(big[j]);
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(V3_TMEM_COLS) : "memory");
    }
}

// ---- host ---------------------------------------------------------------------------------------------------------------------
static float frand(uint64_t &s) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    return (float)((s >> 40) & 0xFFFFFF) / 16777216.0f * 2.0f - 1.0f;
}

template <int NK>
static void run(int rows, const float *d_y0, float *const *d_k, const float *cfh, const uint8_t *d_planes, float *d_out,
                const std::vector<float> &h_y0, const std::vector<std::vector<float>> &h_k, const std::vector<float> &h_A,
                int sms) {
    KP kp{};
    CF cf{};
    for (int m = 0; m < NK; ++m) { kp.p[m] = d_k[m]; cf.c[m] = cfh[m]; }
    const int ntiles = rows / TM;
    CK(cudaFuncSetAttribute(k_linear_stage<NK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    const int grid = ntiles < sms ? ntiles : sms;
    CK(cudaMemset(d_out, 0xff, (size_t)rows * D * 4));
    k_linear_stage<NK><<<grid, THREADS, SMEM_BYTES>>>(d_y0, kp, cf, d_planes, d_out, ntiles);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> out((size_t)rows * D);
    CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
    // reference in double from the float32 stage value (same expression order, no fma)
    double max_err = 0, sum_sq = 0, ref_sq = 0;
    size_t bad = 0;
    const int check_rows[] = {0, 1, 7, 8, 31, 32, 63, 64, 127, 128, 129, 255, rows / 2 + 5, rows - 129, rows - 128, rows - 1};
    std::vector<int> rr(check_rows, check_rows + 16);
    for (int r = 300; r < rows; r += 997) rr.push_back(r);
    for (int r : rr) {
        float y[D];
        for (int c = 0; c < D; ++c) {
            float v = h_y0[(size_t)r * D + c];
            if (NK > 0) {
                volatile float acc = h_k[0][(size_t)r * D + c] * cfh[0];
                for (int m = 1; m < NK; ++m) { volatile float p = h_k[m][(size_t)r * D + c] * cfh[m]; acc = acc + p; }
                v = v + acc;
            }
            y[c] = v;
        }
        for (int n = 0; n < D; ++n) {
            double s = 0;
            for (int c = 0; c < D; ++c) s += (double)y[c] * (double)h_A[(size_t)n * D + c];
            const double e = fabs((double)out[(size_t)r * D + n] - s);
            if (!(e < 1e-3)) ++bad;
            if (e > max_err) max_err = e;
            sum_sq += e * e;
            ref_sq += s * s;
        }
    }
    // timing
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) k_linear_stage<NK><<<grid, THREADS, SMEM_BYTES>>>(d_y0, kp, cf, d_planes, d_out, ntiles);
    CK(cudaEventRecord(e0));
    const int reps = 20;
    for (int i = 0; i < reps; ++i) k_linear_stage<NK><<<grid, THREADS, SMEM_BYTES>>>(d_y0, kp, cf, d_planes, d_out, ntiles);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = ms * 1e3 / reps, bytes = (double)rows * D * 4 * (NK + 2);
    printf("NK=%d rows=%d grid=%d: %.2f us  (%.0f GB/s on %d arrays)  rel rms err %.3e  max abs err %.3e  bad %zu of %zu\n", NK, rows,
           grid, us, bytes / us * 1e-3, NK + 2, sqrt(sum_sq / ref_sq), max_err, bad, rr.size() * D);
}

static const uint32_t *g_wt = nullptr;
static int g_dbg = 0;
constexpr int NSETS = 3;                       // distinct copies of the inputs, rotated per launch: a cold L2 for every launch
static float *g_set_y0[NSETS];
static float *g_set_k[NSETS][MAXK];
static float *g_set_out[NSETS][3];
template <int NK, int NPROD, bool FINAL, int VER = 2, int PF = 0, int SCHED = 0, int PIPE = 0>
static void run2(int rows, const float *d_y0, float *const *d_k, const float *cfh, const float *ceh, const uint8_t *d_planes,
                 float *d_out, float *d_yout, float *d_eout, const std::vector<float> &h_y0,
                 const std::vector<std::vector<float>> &h_k, const std::vector<float> &h_A, int sms) {
    KP kp{};
    CF cr{}, ce{};
    for (int m = 0; m < NK; ++m) { kp.p[m] = d_k[m]; cr.c[m] = cfh[m]; ce.c[m] = ceh[m]; }
    ce.c[6] = (float)g_dbg;
    const int units = (rows + 63) / 64;
    const int grid = units / 2 < sms ? (units + 1) / 2 : sms;
    if (VER == 2) CK(cudaFuncSetAttribute(k_linear_stage2<NK, NPROD, FINAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    else CK(cudaFuncSetAttribute(k_linear_stage3<NK, FINAL, PF, SCHED, PIPE, NPROD>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3_SMEM));
    const uint32_t *wt = g_wt;
    int set = -1;
    auto launch = [&]() {
        KP kq = kp;
        const float *yy = d_y0;
        if (set >= 0) {
            yy = g_set_y0[set % NSETS];
            for (int m = 0; m < NK; ++m) kq.p[m] = g_set_k[set % NSETS][m];
            ++set;
        }
        float *o0 = d_out, *o1 = d_yout, *o2 = d_eout;
        if (set > 0 && g_set_out[0][0]) { o0 = g_set_out[set % NSETS][0]; o1 = g_set_out[set % NSETS][1]; o2 = g_set_out[set % NSETS][2]; }
        if (VER == 2) k_linear_stage2<NK, NPROD, FINAL><<<grid, THREADS, SMEM_BYTES>>>(yy, kq, cr, ce, d_planes, o0, o1, o2, rows);
        else k_linear_stage3<NK, FINAL, PF, SCHED, PIPE, NPROD><<<grid, V3_THREADS, V3_SMEM>>>(yy, kq, cr, ce, wt, o0, o1, o2, rows);
    };
    CK(cudaMemset(d_out, 0xff, (size_t)rows * D * 4));
    CK(cudaMemset(d_yout, 0xff, (size_t)rows * D * 4));
    CK(cudaMemset(d_eout, 0xff, (size_t)rows * D * 4));
    launch();
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> out((size_t)rows * D), yo((size_t)rows * D), eo((size_t)rows * D);
    CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(yo.data(), d_yout, out.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(eo.data(), d_eout, out.size() * 4, cudaMemcpyDeviceToHost));
    double max_err = 0, sum_sq = 0, ref_sq = 0;
    size_t bad = 0, ybad = 0;
    std::vector<int> rr = {0, 1, 7, 8, 31, 32, 63, 64, 65, 127, 128, 129, 255, rows / 2 + 5, rows - 129, rows - 128, rows - 65, rows - 64, rows - 1};
    for (int r = 300; r < rows; r += 499) rr.push_back(r);
    for (int r : rr) {
        if (r < 0 || r >= rows) continue;
        float y[D];
        for (int c = 0; c < D; ++c) {
            float v = h_y0[(size_t)r * D + c];
            if (NK > 0) {
                volatile float acc = h_k[0][(size_t)r * D + c] * cfh[0];
                volatile float ee = h_k[0][(size_t)r * D + c] * ceh[0];
                for (int m = 1; m < NK; ++m) {
                    volatile float p = h_k[m][(size_t)r * D + c] * cfh[m]; acc = acc + p;
                    volatile float pe = h_k[m][(size_t)r * D + c] * ceh[m]; ee = ee + pe;
                }
                v = v + acc;
                if (FINAL && (yo[(size_t)r * D + c] != v || eo[(size_t)r * D + c] != ee)) ++ybad;
            }
            y[c] = v;
        }
        for (int n = 0; n < D; ++n) {
            double s = 0;
            for (int c = 0; c < D; ++c) s += (double)y[c] * (double)h_A[(size_t)n * D + c];
            const double e = fabs((double)out[(size_t)r * D + n] - s);
            if (!(e < 1e-3)) ++bad;
            if (e > max_err) max_err = e;
            sum_sq += e * e;
            ref_sq += s * s;
        }
    }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    set = g_set_y0[0] ? 0 : -1;
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaEventRecord(e0));
    const int reps = 21;
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const int arrays = NK + 2 + (FINAL ? 2 : 0);
    const double us = ms * 1e3 / reps, bytes = (double)rows * D * 4 * arrays;
    printf("v%d PIPE=%d SCHED=%d PF=%d NK=%d NPROD=%d FINAL=%d rows=%d grid=%d: %.2f us  (%.0f GB/s on %d arrays)  rel rms err %.3e  max abs err %.3e  bad %zu  y/err mismatches %zu\n",
           VER, PIPE, SCHED, PF, NK, NPROD, (int)FINAL, rows, grid, us, bytes / us * 1e-3, arrays, sqrt(sum_sq / ref_sq), max_err, bad, ybad);
}

int main(int argc, char **argv) {
    const int rows = argc > 1 ? atoi(argv[1]) : 65536;
    int sms;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    uint64_t seed = 12345;
    std::vector<float> h_A((size_t)D * D), h_y0((size_t)rows * D);
    for (auto &v : h_A) v = frand(seed) * 0.09f;
    const bool gauss = getenv("EXP_GAUSS") != nullptr;
    auto gen = [&]() { return gauss ? (frand(seed) + frand(seed) + frand(seed) + frand(seed)) * 0.866f : frand(seed); };
    for (auto &v : h_y0) v = gen();
    std::vector<std::vector<float>> h_k(MAXK, std::vector<float>((size_t)rows * D));
    for (auto &k : h_k) for (auto &v : k) v = gen();
    const float cfh[MAXK] = {0.0123f, -0.031f, 0.027f, 0.0451f, -0.0083f, 0.019f, 0.0071f};
    float *d_A, *d_y0, *d_out, *d_k[MAXK];
    uint8_t *d_planes;
    CK(cudaMalloc(&d_A, h_A.size() * 4)); CK(cudaMalloc(&d_y0, h_y0.size() * 4)); CK(cudaMalloc(&d_out, h_y0.size() * 4));
    CK(cudaMalloc(&d_planes, 3 * PLANE_BYTES));
    CK(cudaMemcpy(d_A, h_A.data(), h_A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_y0, h_y0.data(), h_y0.size() * 4, cudaMemcpyHostToDevice));
    for (int m = 0; m < MAXK; ++m) {
        CK(cudaMalloc(&d_k[m], h_y0.size() * 4));
        CK(cudaMemcpy(d_k[m], h_k[m].data(), h_y0.size() * 4, cudaMemcpyHostToDevice));
    }
    k_split_weights<<<D, 32>>>(d_A, d_planes);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    run<0>(rows, d_y0, d_k, cfh, d_planes, d_out, h_y0, h_k, h_A, sms);
    run<1>(rows, d_y0, d_k, cfh, d_planes, d_out, h_y0, h_k, h_A, sms);
    run<3>(rows, d_y0, d_k, cfh, d_planes, d_out, h_y0, h_k, h_A, sms);
    run<5>(rows, d_y0, d_k, cfh, d_planes, d_out, h_y0, h_k, h_A, sms);
    float *d_yout, *d_eout;
    CK(cudaMalloc(&d_yout, h_y0.size() * 4)); CK(cudaMalloc(&d_eout, h_y0.size() * 4));
    const float ceh[MAXK] = {0.0012f, 0.0005f, -0.0034f, 0.0021f, -0.0018f, 0.0041f, -0.0009f};
    uint32_t *d_wt;
    CK(cudaMalloc(&d_wt, 3 * D * 64 * 4));
    k_split_weights_t<<<D, 64>>>(d_A, d_wt);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    g_wt = d_wt;
#define R3(NK, F) run2<NK, 9, F, 3>(rows, d_y0, d_k, cfh, ceh, d_planes, d_out, d_yout, d_eout, h_y0, h_k, h_A, sms)
#define R4(NK, F, PF) run2<NK, 9, F, 3, PF>(rows, d_y0, d_k, cfh, ceh, d_planes, d_out, d_yout, d_eout, h_y0, h_k, h_A, sms)
    if (argc > 2) {
        // one slab, arrays carved at a fixed stride: argv[3] = padding in bytes added to the array size (0: the arrays are
        // exactly 2^25 bytes apart at B = 65536, which is how a caching allocator lays out equal-sized buffers)
        const size_t pad = argc > 3 ? (size_t)atol(argv[3]) : 0, stride = h_y0.size() * 4 + pad;
        char *slab;
        CK(cudaMalloc(&slab, stride * NSETS * (MAXK + 1) + (1 << 25)));
        slab = (char *)(((uintptr_t)slab + (1 << 25) - 1) & ~(uintptr_t)((1 << 25) - 1));
        for (int st = 0; st < NSETS; ++st) {
            g_set_y0[st] = (float *)(slab + stride * (st * (MAXK + 1)));
            CK(cudaMemcpy(g_set_y0[st], d_y0, h_y0.size() * 4, cudaMemcpyDeviceToDevice));
            for (int m = 0; m < MAXK; ++m) {
                g_set_k[st][m] = (float *)(slab + stride * (st * (MAXK + 1) + 1 + m));
                CK(cudaMemcpy(g_set_k[st][m], d_k[m], h_y0.size() * 4, cudaMemcpyDeviceToDevice));
            }
        }
        for (int st = 0; st < NSETS; ++st)
            for (int j = 0; j < 3; ++j) CK(cudaMalloc(&g_set_out[st][j], h_y0.size() * 4));
        printf("cold inputs and outputs (3 rotating copies), arrays %zu bytes apart:\n", stride);
        R3(0, false); R3(1, false); R3(2, false); R3(3, false); R3(4, false); R3(5, false); R3(5, true);
#define R5(NK, F) run2<NK, 9, F, 3, 0, 0, 1>(rows, d_y0, d_k, cfh, ceh, d_planes, d_out, d_yout, d_eout, h_y0, h_k, h_A, sms)
        R5(0, false); R5(1, false); R5(2, false); R5(3, false); R5(4, false); R5(5, false); R5(5, true);
        run2<5, 9, true, 3, 0, 0, 1>(rows - 40, d_y0, d_k, cfh, ceh, d_planes, d_out, d_yout, d_eout, h_y0, h_k, h_A, sms);
        run2<3, 9, false, 3, 0, 0, 1>(1000, d_y0, d_k, cfh, ceh, d_planes, d_out, d_yout, d_eout, h_y0, h_k, h_A, sms);
        run2<3, 9, false, 3, 0, 0, 1>(64, d_y0, d_k, cfh, ceh, d_planes, d_out, d_yout, d_eout, h_y0, h_k, h_A, sms);
#define R6(NK, F, NP) run2<NK, NP, F, 3, 0, 0, 0>(rows, d_y0, d_k, cfh, ceh, d_planes, d_out, d_yout, d_eout, h_y0, h_k, h_A, sms)
        printf("six and three bf16 products instead of nine:\n");
        R6(0, false, 6); R6(1, false, 6); R6(2, false, 6); R6(3, false, 6); R6(4, false, 6); R6(5, false, 6); R6(5, true, 6);
        R6(0, false, 3); R6(3, false, 3); R6(5, false, 3);
        for (int dbg : {1, 2, 3, 4, 8, 12, 15}) {
            g_dbg = dbg;
            printf("ablation %d (1 no epilogue stores, 2 no MMAs, 4 no y0 loads, 8 no split/STS):\n", dbg);
            R3(0, false); R3(3, false);
        }
        g_dbg = 0;
        run2<3, 9, false, 3, 0, 0, 1>(148 * 128 * 2 + 64 * 100 + 13, d_y0, d_k, cfh, ceh, d_planes, d_out, d_yout, d_eout, h_y0, h_k, h_A, sms);
        run2<5, 9, true, 3>(rows - 40, d_y0, d_k, cfh, ceh, d_planes, d_out, d_yout, d_eout, h_y0, h_k, h_A, sms);
        run2<3, 9, false, 3>(1000, d_y0, d_k, cfh, ceh, d_planes, d_out, d_yout, d_eout, h_y0, h_k, h_A, sms);
        printf("done v3\n");
        return 0;
    }
#define R2(NK, NP, F) run2<NK, NP, F>(rows, d_y0, d_k, cfh, ceh, d_planes, d_out, d_yout, d_eout, h_y0, h_k, h_A, sms)
    R2(0, 9, false); R2(1, 9, false); R2(2, 9, false); R2(3, 9, false); R2(4, 9, false); R2(5, 9, false); R2(5, 9, true);
    R2(0, 6, false); R2(3, 6, false); R2(5, 6, false); R2(5, 6, true);
    run2<5, 9, true>(rows - 40, d_y0, d_k, cfh, ceh, d_planes, d_out, d_yout, d_eout, h_y0, h_k, h_A, sms);
    run2<3, 9, false>(1000, d_y0, d_k, cfh, ceh, d_planes, d_out, d_yout, d_eout, h_y0, h_k, h_A, sms);
    printf("done\n");
    return 0;
}
