// exp_combine.cu -- tuning experiment for the stage-combine kernel (not part of libtdq).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o gpurun_out/exp_combine scripts/exp_combine.cu
// Variants of out = y0 + sum_m k_m*c_m over N = 65536*128 float32 elements, NK in {1, 3, 5}:
//   tile      one tile per block, U 16-byte vectors per thread per operand (the libtdq r1 kernel is U=2, 256 thr)
//   persist   grid = 148*R blocks, grid-stride loop
//   tma       cp.async.bulk global->shared ring (mbarrier), compute from shared, st.global
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

struct KP { const float *p[8]; float c[8]; };

__device__ __forceinline__ float4 ldv(const float *p) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ldv_plain(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float4 ldv_nc(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
template <int LD> __device__ __forceinline__ float4 ldx(const float *p) {
    if (LD == 0) return ldv(p);
    if (LD == 1) return ldv_plain(p);
    return ldv_nc(p);
}
__device__ __forceinline__ void stv(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ void stv_cs(float *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int NK> __device__ __forceinline__ float4 comb(float4 a, const float4 *k, const float *c) {
    float4 acc;
    acc.x = __fmul_rn(k[0].x, c[0]); acc.y = __fmul_rn(k[0].y, c[0]); acc.z = __fmul_rn(k[0].z, c[0]); acc.w = __fmul_rn(k[0].w, c[0]);
#pragma unroll
    for (int m = 1; m < NK; ++m) {
        acc.x = __fadd_rn(acc.x, __fmul_rn(k[m].x, c[m])); acc.y = __fadd_rn(acc.y, __fmul_rn(k[m].y, c[m]));
        acc.z = __fadd_rn(acc.z, __fmul_rn(k[m].z, c[m])); acc.w = __fadd_rn(acc.w, __fmul_rn(k[m].w, c[m]));
    }
    float4 r;
    r.x = __fadd_rn(a.x, acc.x); r.y = __fadd_rn(a.y, acc.y); r.z = __fadd_rn(a.z, acc.z); r.w = __fadd_rn(a.w, acc.w);
    return r;
}

template <int NK, int U, int THREADS, int LD, int ST>
__global__ void __launch_bounds__(THREADS) k_tile(float *out, const float *y0, KP kp, size_t nvec) {
    const size_t base = (size_t)blockIdx.x * (THREADS * U) + threadIdx.x;
    float4 a[U], kv[U][NK];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const size_t v = base + (size_t)u * THREADS;
        if (v < nvec) {
            a[u] = ldx<LD>(y0 + v * 4);
#pragma unroll
            for (int m = 0; m < NK; ++m) kv[u][m] = ldx<LD>(kp.p[m] + v * 4);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const size_t v = base + (size_t)u * THREADS;
        if (v < nvec) {
            float4 r = comb<NK>(a[u], kv[u], kp.c);
            if (ST == 0) stv(out + v * 4, r); else stv_cs(out + v * 4, r);
        }
    }
}

template <int NK, int U, int THREADS, int LD, int ST>
__global__ void __launch_bounds__(THREADS) k_persist(float *out, const float *y0, KP kp, size_t nvec) {
    const size_t stride = (size_t)gridDim.x * THREADS * U;
    for (size_t base = (size_t)blockIdx.x * (THREADS * U) + threadIdx.x; base < nvec; base += stride) {
        float4 a[U], kv[U][NK];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t v = base + (size_t)u * THREADS;
            if (v < nvec) {
                a[u] = ldx<LD>(y0 + v * 4);
#pragma unroll
                for (int m = 0; m < NK; ++m) kv[u][m] = ldx<LD>(kp.p[m] + v * 4);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t v = base + (size_t)u * THREADS;
            if (v < nvec) {
                float4 r = comb<NK>(a[u], kv[u], kp.c);
                if (ST == 0) stv(out + v * 4, r); else stv_cs(out + v * 4, r);
            }
        }
    }
}

// ---- 256-bit loads (sm_100: ld.global.v8.b32, the only form that takes L2 eviction priorities) ------
struct F8 { float v[8]; };
template <int EV> __device__ __forceinline__ F8 ld256(const float *p) {
    F8 r;
    if (EV)
        asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
    else
        asm volatile("ld.global.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
    return r;
}
__device__ __forceinline__ void st256(float *p, const F8 &r) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]) : "memory");
}
template <int NK, int U, int THREADS, int EV>
__global__ void __launch_bounds__(THREADS) k_tile256(float *out, const float *y0, KP kp, size_t nv8) {
    const size_t base = (size_t)blockIdx.x * (THREADS * U) + threadIdx.x;
    F8 a[U], kv[U][NK];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const size_t v = base + (size_t)u * THREADS;
        if (v < nv8) {
            a[u] = ld256<EV>(y0 + v * 8);
#pragma unroll
            for (int m = 0; m < NK; ++m) kv[u][m] = ld256<EV>(kp.p[m] + v * 8);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const size_t v = base + (size_t)u * THREADS;
        if (v < nv8) {
            F8 r;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float acc = __fmul_rn(kv[u][0].v[e], kp.c[0]);
#pragma unroll
                for (int m = 1; m < NK; ++m) acc = __fadd_rn(acc, __fmul_rn(kv[u][m].v[e], kp.c[m]));
                r.v[e] = __fadd_rn(a[u].v[e], acc);
            }
            st256(out + v * 8, r);
        }
    }
}

// ---- TMA (bulk async copy) variant --------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_LOOP:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE;\n bra WAIT_LOOP;\n DONE:\n}\n"
        :: "r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// TILE_V float4 vectors per operand per stage; STAGES-deep ring; one producer thread.
template <int NK, int THREADS, int TILE_V, int STAGES>
__global__ void __launch_bounds__(THREADS) k_tma(float *out, const float *y0, KP kp, size_t nvec) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *buf = reinterpret_cast<float4 *>(smem_raw);                 // [STAGES][NK+1][TILE_V]
    __shared__ uint64_t full[STAGES];
    const size_t ntiles = (nvec + TILE_V - 1) / TILE_V;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](size_t tile, int s) {
        const size_t v0 = tile * TILE_V;
        const uint32_t nv = (uint32_t)((nvec - v0 < (size_t)TILE_V) ? (nvec - v0) : TILE_V);
        const uint32_t bytes = nv * 16;
        mbar_expect_tx(&full[s], bytes * (NK + 1));
        float4 *st = buf + (size_t)s * (NK + 1) * TILE_V;
        bulk_g2s(st, y0 + v0 * 4, bytes, &full[s]);
#pragma unroll
        for (int m = 0; m < NK; ++m) bulk_g2s(st + (size_t)(m + 1) * TILE_V, kp.p[m] + v0 * 4, bytes, &full[s]);
    };
    // prologue
    size_t t_issue = blockIdx.x;
    if (threadIdx.x == 0)
        for (int s = 0; s < STAGES && t_issue < ntiles; ++s, t_issue += gridDim.x) issue(t_issue, s);
    int s = 0;
    uint32_t phase = 0;
    size_t n_done = 0;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++n_done) {
        mbar_wait(&full[s], phase);
        const size_t v0 = tile * TILE_V;
        float4 *st = buf + (size_t)s * (NK + 1) * TILE_V;
#pragma unroll
        for (int i = threadIdx.x; i < TILE_V; i += THREADS) {
            if (v0 + i < nvec) {
                float4 kv[NK];
#pragma unroll
                for (int m = 0; m < NK; ++m) kv[m] = st[(size_t)(m + 1) * TILE_V + i];
                stv(out + (v0 + i) * 4, comb<NK>(st[i], kv, kp.c));
            }
        }
        __syncthreads();                                  // stage consumed
        if (threadIdx.x == 0) {
            const size_t nxt = tile + (size_t)STAGES * gridDim.x;
            if (nxt < ntiles) issue(nxt, s);
        }
        if (++s == STAGES) { s = 0; phase ^= 1; }
    }
}

// ---- round 2: warp-specialised bulk-async pipeline ---------------------------------------------------------
// the r1 sweep kept too few bytes in flight for some shapes and synchronised the whole block
// once per stage.  This variant has ONE producer warp (lane 0 issues cp.async.bulk for all NK+1 operands of a stage
// as soon as the consumers have released it through an `empty` mbarrier) and CONSUMERS-1 consumer warps that never
// meet the producer at a __syncthreads: full[s] (tx-count) -> compute from shared -> arrive on empty[s].
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
template <int NK, int THREADS, int TILE_V, int STAGES>
__global__ void __launch_bounds__(THREADS) k_tma2(float *out, const float *y0, KP kp, size_t nvec) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4 *buf = reinterpret_cast<float4 *>(smem_raw);                 // [STAGES][NK+1][TILE_V]
    __shared__ uint64_t full[STAGES], empty[STAGES];
    constexpr int CONS = THREADS - 32;                                  // consumer threads
    const size_t ntiles = (nvec + TILE_V - 1) / TILE_V;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CONS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x < 32) {                                             // producer warp
        if (threadIdx.x == 0) {
            int s = 0;
            uint32_t phase = 0;
            size_t it = 0;
            for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                if (it >= (size_t)STAGES) mbar_wait(&empty[s], phase ^ 1);      // consumers released the slot
                const size_t v0 = tile * TILE_V;
                const uint32_t nv = (uint32_t)((nvec - v0 < (size_t)TILE_V) ? (nvec - v0) : TILE_V);
                const uint32_t bytes = nv * 16;
                mbar_expect_tx(&full[s], bytes * (NK + 1));
                float4 *st = buf + (size_t)s * (NK + 1) * TILE_V;
                bulk_g2s(st, y0 + v0 * 4, bytes, &full[s]);
#pragma unroll
                for (int m = 0; m < NK; ++m) bulk_g2s(st + (size_t)(m + 1) * TILE_V, kp.p[m] + v0 * 4, bytes, &full[s]);
                if (++s == STAGES) { s = 0; phase ^= 1; }
            }
        }
        return;
    }
    const int ct = threadIdx.x - 32;
    int s = 0;
    uint32_t phase = 0;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        mbar_wait(&full[s], phase);
        const size_t v0 = tile * TILE_V;
        float4 *st = buf + (size_t)s * (NK + 1) * TILE_V;
        constexpr int PER = (TILE_V + CONS - 1) / CONS;
        float4 r[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int i = ct + j * CONS;
            if (i < TILE_V) {
                float4 kv[NK];
#pragma unroll
                for (int m = 0; m < NK; ++m) kv[m] = st[(size_t)(m + 1) * TILE_V + i];
                r[j] = comb<NK>(st[i], kv, kp.c);
            }
        }
        mbar_arrive(&empty[s]);                                          // operands are in registers: release the slot
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int i = ct + j * CONS;
            if (i < TILE_V && v0 + i < nvec) stv(out + (v0 + i) * 4, r[j]);
        }
        if (++s == STAGES) { s = 0; phase ^= 1; }
    }
}

struct Bufs { float *y0, *out, *k[8]; size_t n; };

template <typename F> float time_ms(F f, int reps = 30) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) f();
    CK(cudaDeviceSynchronize());
    float tot = 0;
    for (int i = 0; i < reps; ++i) {
        CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); tot += ms;
    }
    CK(cudaGetLastError());
    return tot / reps;
}

template <int NK> void run_all(const Bufs &B) {
    KP kp;
    for (int m = 0; m < 8; ++m) { kp.p[m] = B.k[m]; kp.c[m] = 0.01f * (m + 1); }
    const size_t nvec = B.n / 4;
    const double bytes = (double)(NK + 2) * B.n * 4;
    auto report = [&](const char *name, float ms) { printf("NK=%d %-34s %8.4f ms  %7.1f GB/s\n", NK, name, ms, bytes / ms / 1e6); fflush(stdout); };
#define TILE(U, T, LD, ST, label) report(label, time_ms([&] { k_tile<NK, U, T, LD, ST><<<(unsigned)((nvec + (size_t)T * U - 1) / ((size_t)T * U)), T>>>(B.out, B.y0, kp, nvec); }))
    TILE(2, 256, 0, 0, "tile U=2 T=256 (r1 baseline)");
    TILE(1, 256, 0, 0, "tile U=1 T=256");
    TILE(4, 256, 0, 0, "tile U=4 T=256");
    TILE(8, 256, 0, 0, "tile U=8 T=256");
    TILE(4, 128, 0, 0, "tile U=4 T=128");
    TILE(4, 512, 0, 0, "tile U=4 T=512");
    TILE(2, 512, 0, 0, "tile U=2 T=512");
    TILE(4, 256, 1, 0, "tile U=4 T=256 ld.plain");
    TILE(4, 256, 2, 0, "tile U=4 T=256 ld.nc");
    TILE(4, 256, 0, 1, "tile U=4 T=256 st.cs");
    TILE(4, 256, 2, 1, "tile U=4 T=256 nc + st.cs");
#define T256(U, T, EV, label) report(label, time_ms([&] { k_tile256<NK, U, T, EV><<<(unsigned)((nvec / 2 + (size_t)T * U - 1) / ((size_t)T * U)), T>>>(B.out, B.y0, kp, nvec / 2); }))
    T256(1, 256, 0, "tile256 U=1 T=256");
    T256(2, 256, 0, "tile256 U=2 T=256");
    T256(2, 256, 1, "tile256 U=2 T=256 evict_first");
    T256(4, 256, 0, "tile256 U=4 T=256");
    T256(2, 128, 0, "tile256 U=2 T=128");
    T256(1, 512, 0, "tile256 U=1 T=512");
#define PERS(U, T, R, LD, ST, label) report(label, time_ms([&] { k_persist<NK, U, T, LD, ST><<<148 * R, T>>>(B.out, B.y0, kp, nvec); }))
    PERS(2, 256, 8, 0, 0, "persist U=2 T=256 R=8");
    PERS(4, 256, 4, 0, 0, "persist U=4 T=256 R=4");
    PERS(4, 256, 8, 0, 0, "persist U=4 T=256 R=8");
    PERS(2, 512, 4, 0, 0, "persist U=2 T=512 R=4");
    PERS(4, 512, 2, 0, 0, "persist U=4 T=512 R=2");
    PERS(4, 512, 4, 0, 0, "persist U=4 T=512 R=4");
    PERS(8, 256, 4, 0, 0, "persist U=8 T=256 R=4");
    PERS(4, 256, 8, 2, 1, "persist U=4 T=256 R=8 nc+cs");
#define TMA(T, TV, S, R, label) do { \
        size_t sm = (size_t)S * (NK + 1) * TV * 16; \
        CK(cudaFuncSetAttribute(k_tma<NK, T, TV, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
        report(label, time_ms([&] { k_tma<NK, T, TV, S><<<148 * R, T, sm>>>(B.out, B.y0, kp, nvec); })); } while (0)
    TMA(256, 256, 3, 2, "tma T=256 tile=4KB S=3 R=2");
    TMA(256, 256, 4, 2, "tma T=256 tile=4KB S=4 R=2");
    TMA(256, 512, 3, 1, "tma T=256 tile=8KB S=3 R=1");
    TMA(512, 512, 3, 1, "tma T=512 tile=8KB S=3 R=1");
    TMA(256, 128, 4, 4, "tma T=256 tile=2KB S=4 R=4");
    // round 2: >= 96 KB in flight per SM wherever shared memory allows it (227 KB per block), warp-specialised
    // pipeline (k_tma2: producer warp + empty/full mbarriers, no block-wide sync per stage)
#define TMA2(T, TV, S, R, label) do { \
        size_t sm = (size_t)S * (NK + 1) * TV * 16; \
        if (sm * R <= 220 * 1024 && sm <= 220 * 1024) { \
            char nm[96]; snprintf(nm, sizeof(nm), "%s [%zu KB/SM]", label, sm * R / 1024); \
            CK(cudaFuncSetAttribute(k_tma2<NK, T, TV, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
            report(nm, time_ms([&] { k_tma2<NK, T, TV, S><<<148 * R, T, sm>>>(B.out, B.y0, kp, nvec); })); } } while (0)
    TMA2(288, 256, 4, 2, "tma2 T=256+32 tile=4KB S=4 R=2");
    TMA2(288, 256, 6, 2, "tma2 T=256+32 tile=4KB S=6 R=2");
    TMA2(288, 256, 8, 2, "tma2 T=256+32 tile=4KB S=8 R=2");
    TMA2(288, 512, 4, 2, "tma2 T=256+32 tile=8KB S=4 R=2");
    TMA2(288, 512, 6, 1, "tma2 T=256+32 tile=8KB S=6 R=1");
    TMA2(288, 512, 8, 1, "tma2 T=256+32 tile=8KB S=8 R=1");
    TMA2(544, 512, 4, 1, "tma2 T=512+32 tile=8KB S=4 R=1");
    TMA2(544, 1024, 4, 1, "tma2 T=512+32 tile=16KB S=4 R=1");
    TMA2(544, 1024, 6, 1, "tma2 T=512+32 tile=16KB S=6 R=1");
    TMA2(288, 1024, 3, 1, "tma2 T=256+32 tile=16KB S=3 R=1");
    TMA2(288, 128, 8, 4, "tma2 T=256+32 tile=2KB S=8 R=4");
}

__global__ void k_copy(float4 *o, const float4 *i, size_t nvec) {
    size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < nvec) o[v] = i[v];
}

int main() {
    Bufs B;
    B.n = (size_t)65536 * 128;
    CK(cudaMalloc(&B.y0, B.n * 4)); CK(cudaMalloc(&B.out, B.n * 4));
    for (int m = 0; m < 8; ++m) { CK(cudaMalloc(&B.k[m], B.n * 4)); CK(cudaMemset(B.k[m], 0, B.n * 4)); }
    CK(cudaMemset(B.y0, 0, B.n * 4));
    // reference points: cudaMemcpy D2D and a plain float4 copy kernel at this size, and at 1 GiB
    {
        float ms = time_ms([&] { CK(cudaMemcpyAsync(B.out, B.y0, B.n * 4, cudaMemcpyDeviceToDevice)); });
        printf("cudaMemcpy D2D 33.5MB             %8.4f ms  %7.1f GB/s\n", ms, 2.0 * B.n * 4 / ms / 1e6);
        ms = time_ms([&] { k_copy<<<(unsigned)((B.n / 4 + 255) / 256), 256>>>((float4 *)B.out, (const float4 *)B.y0, B.n / 4); });
        printf("float4 copy kernel 33.5MB         %8.4f ms  %7.1f GB/s\n", ms, 2.0 * B.n * 4 / ms / 1e6);
        float *a, *b; size_t big = (size_t)1 << 28;   // 1 GiB each
        CK(cudaMalloc(&a, big * 4)); CK(cudaMalloc(&b, big * 4));
        ms = time_ms([&] { CK(cudaMemcpyAsync(b, a, big * 4, cudaMemcpyDeviceToDevice)); }, 10);
        printf("cudaMemcpy D2D 1GiB               %8.4f ms  %7.1f GB/s\n", ms, 2.0 * big * 4 / ms / 1e6);
        ms = time_ms([&] { k_copy<<<(unsigned)((big / 4 + 255) / 256), 256>>>((float4 *)b, (const float4 *)a, big / 4); }, 10);
        printf("float4 copy kernel 1GiB           %8.4f ms  %7.1f GB/s\n", ms, 2.0 * big * 4 / ms / 1e6);
        CK(cudaFree(a)); CK(cudaFree(b));
    }
    run_all<1>(B);
    run_all<3>(B);
    run_all<5>(B);
    return 0;
}
