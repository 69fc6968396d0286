// exp_fused_norm_fit.cu -- round-2 experiment (not part of libtdq): is it worth fusing the error norm with a
// SPECULATIVE interpolant fit?  
//
// dopri5, N = 65536*128 float32.  Per attempt today (r1):
//     k_reduce      reads y0,y1,k0,k2..k6 (8 arrays)                          =  8 N*s
//     k_fit_commit  reads the same 8 arrays, writes e,d,c,b,a,y0,k0 (7 arrays) = 15 N*s      -> 23 N*s
// Variant measured here:
//     k_norm_fit    reads the 8 arrays once, accumulates the norm, writes c,b,a speculatively  = 11 N*s
//     k_commit      on accept: reads y0,k0,y1,k6, writes e,d (old y0, dt*old k0), y0<-y1, k0<-k6 =  8 N*s  -> 19 N*s
// (with a device pointer table for e/y0 and f0/k0 the commit drops to 4 N*s -> 15 N*s.)
// The speculative write is legal because a rejected attempt never needs the previous interval's coefficients
// again: all outputs inside it were emitted right after its accept.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o scripts/exp_fused_norm_fit.bin scripts/exp_fused_norm_fit.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

struct P8 { const float *y0, *y1, *k[6]; };          // k[0]=k0, k[1..5]=k2..k6
struct C7 { float ec[6], mc[6]; float rtol, atol, dt; };

__device__ __forceinline__ float4 ldv(const float *p) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stv(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ float lane(const float4 &v, int e) { return e == 0 ? v.x : e == 1 ? v.y : e == 2 ? v.z : v.w; }
__device__ __forceinline__ void setlane(float4 &v, int e, float x) { if (e == 0) v.x = x; else if (e == 1) v.y = x; else if (e == 2) v.z = x; else v.w = x; }

__device__ __forceinline__ double block_sum(double v, double *sm) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0;
    if (threadIdx.x < 32) {
        r = threadIdx.x < blockDim.x / 32 ? sm[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    return r;
}

// ---- today's pair ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_reduce(P8 p, C7 c, double *partials, size_t nvec) {
    __shared__ double sm[8];
    double acc = 0;
    for (size_t v = (size_t)blockIdx.x * 256 + threadIdx.x; v < nvec; v += (size_t)gridDim.x * 256) {
        float4 a0 = ldv(p.y0 + v * 4), a1 = ldv(p.y1 + v * 4), kv[6];
#pragma unroll
        for (int m = 0; m < 6; ++m) kv[m] = ldv(p.k[m] + v * 4);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float err = __fmul_rn(lane(kv[0], e), c.ec[0]);
#pragma unroll
            for (int m = 1; m < 6; ++m) err = __fadd_rn(err, __fmul_rn(lane(kv[m], e), c.ec[m]));
            const float tol = __fadd_rn(c.atol, __fmul_rn(c.rtol, fmaxf(fabsf(lane(a0, e)), fabsf(lane(a1, e)))));
            const float q = __fdiv_rn(err, tol);
            acc += (double)__fmul_rn(q, q);
        }
    }
    const double s = block_sum(acc, sm);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

__device__ __forceinline__ void fit1(float y0, float y1, float f0, float f1, float ymid, float dt, float &cq, float &b, float &a) {
    const float two_dt = __fmul_rn(2.f, dt);
    a = __fadd_rn(__fsub_rn(__fmul_rn(two_dt, __fsub_rn(f1, f0)), __fmul_rn(8.f, __fadd_rn(y1, y0))), __fmul_rn(16.f, ymid));
    b = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(dt, __fsub_rn(__fmul_rn(5.f, f0), __fmul_rn(3.f, f1))), __fmul_rn(18.f, y0)), __fmul_rn(14.f, y1)), __fmul_rn(32.f, ymid));
    cq = __fadd_rn(__fsub_rn(__fsub_rn(__fmul_rn(dt, __fsub_rn(f1, __fmul_rn(4.f, f0))), __fmul_rn(11.f, y0)), __fmul_rn(5.f, y1)), __fmul_rn(16.f, ymid));
}

__global__ void __launch_bounds__(256) k_fit_commit(P8 p, C7 c, float *y0w, float *k0w, float *ce, float *cd, float *cc, float *cb, float *ca, size_t nvec) {
    for (size_t v = (size_t)blockIdx.x * 256 + threadIdx.x; v < nvec; v += (size_t)gridDim.x * 256) {
        float4 a0 = ldv(p.y0 + v * 4), a1 = ldv(p.y1 + v * 4), kv[6], rc, rb, ra, rd;
#pragma unroll
        for (int m = 0; m < 6; ++m) kv[m] = ldv(p.k[m] + v * 4);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float ym = __fmul_rn(lane(kv[0], e), c.mc[0]);
#pragma unroll
            for (int m = 1; m < 6; ++m) ym = __fadd_rn(ym, __fmul_rn(lane(kv[m], e), c.mc[m]));
            ym = __fadd_rn(lane(a0, e), ym);
            float cq, b, a;
            fit1(lane(a0, e), lane(a1, e), lane(kv[0], e), lane(kv[5], e), ym, c.dt, cq, b, a);
            setlane(rc, e, cq); setlane(rb, e, b); setlane(ra, e, a); setlane(rd, e, __fmul_rn(c.dt, lane(kv[0], e)));
        }
        stv(ce + v * 4, a0); stv(cd + v * 4, rd); stv(cc + v * 4, rc); stv(cb + v * 4, rb); stv(ca + v * 4, ra);
        stv(y0w + v * 4, a1); stv(k0w + v * 4, kv[5]);
    }
}

// ---- fused variant ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_norm_fit(P8 p, C7 c, double *partials, float *cc, float *cb, float *ca, size_t nvec) {
    __shared__ double sm[8];
    double acc = 0;
    for (size_t v = (size_t)blockIdx.x * 256 + threadIdx.x; v < nvec; v += (size_t)gridDim.x * 256) {
        float4 a0 = ldv(p.y0 + v * 4), a1 = ldv(p.y1 + v * 4), kv[6], rc, rb, ra;
#pragma unroll
        for (int m = 0; m < 6; ++m) kv[m] = ldv(p.k[m] + v * 4);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float err = __fmul_rn(lane(kv[0], e), c.ec[0]), ym = __fmul_rn(lane(kv[0], e), c.mc[0]);
#pragma unroll
            for (int m = 1; m < 6; ++m) {
                err = __fadd_rn(err, __fmul_rn(lane(kv[m], e), c.ec[m]));
                ym = __fadd_rn(ym, __fmul_rn(lane(kv[m], e), c.mc[m]));
            }
            ym = __fadd_rn(lane(a0, e), ym);
            const float tol = __fadd_rn(c.atol, __fmul_rn(c.rtol, fmaxf(fabsf(lane(a0, e)), fabsf(lane(a1, e)))));
            const float q = __fdiv_rn(err, tol);
            acc += (double)__fmul_rn(q, q);
            float cq, b, a;
            fit1(lane(a0, e), lane(a1, e), lane(kv[0], e), lane(kv[5], e), ym, c.dt, cq, b, a);
            setlane(rc, e, cq); setlane(rb, e, b); setlane(ra, e, a);
        }
        stv(cc + v * 4, rc); stv(cb + v * 4, rb); stv(ca + v * 4, ra);
    }
    const double s = block_sum(acc, sm);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_commit(float *y0w, float *k0w, const float *y1, const float *k6, float *ce, float *cd, float dt, size_t nvec) {
    for (size_t v = (size_t)blockIdx.x * 256 + threadIdx.x; v < nvec; v += (size_t)gridDim.x * 256) {
        float4 a0 = ldv(y0w + v * 4), f0 = ldv(k0w + v * 4), a1 = ldv(y1 + v * 4), f1 = ldv(k6 + v * 4), d;
        d.x = __fmul_rn(dt, f0.x); d.y = __fmul_rn(dt, f0.y); d.z = __fmul_rn(dt, f0.z); d.w = __fmul_rn(dt, f0.w);
        stv(ce + v * 4, a0); stv(cd + v * 4, d); stv(y0w + v * 4, a1); stv(k0w + v * 4, f1);
    }
}
// pointer-table flavour: e and d are the OLD y0 / k0 buffers themselves (swapped on the device), so only the carries are copied
__global__ void __launch_bounds__(256) k_commit_swap(float *y0_new, float *k0_new, const float *y1, const float *k6, size_t nvec) {
    for (size_t v = (size_t)blockIdx.x * 256 + threadIdx.x; v < nvec; v += (size_t)gridDim.x * 256) {
        stv(y0_new + v * 4, ldv(y1 + v * 4));
        stv(k0_new + v * 4, ldv(k6 + v * 4));
    }
}

template <typename F> float time_ms(F f, int reps = 20) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    CK(cudaGetLastError());
    return ms / reps;
}

int main() {
    const size_t n = (size_t)65536 * 128, nvec = n / 4;
    float *buf[18];
    for (int i = 0; i < 18; ++i) { CK(cudaMalloc(&buf[i], n * 4)); CK(cudaMemset(buf[i], 0, n * 4)); }
    double *partials; CK(cudaMalloc(&partials, 8192 * 8));
    P8 p = {buf[0], buf[1], {buf[2], buf[3], buf[4], buf[5], buf[6], buf[7]}};
    C7 c; for (int m = 0; m < 6; ++m) { c.ec[m] = 0.01f * (m + 1); c.mc[m] = 0.02f * (m + 1); } c.rtol = 1e-5f; c.atol = 1e-7f; c.dt = 0.1f;
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * 8;
    const double NS = (double)n * 4 / 1e6;     // MB per array
    float t_red = time_ms([&] { k_reduce<<<grid, 256>>>(p, c, partials, nvec); });
    float t_fit = time_ms([&] { k_fit_commit<<<grid, 256>>>(p, c, buf[8], buf[9], buf[10], buf[11], buf[12], buf[13], buf[14], nvec); });
    float t_nf = time_ms([&] { k_norm_fit<<<grid, 256>>>(p, c, partials, buf[12], buf[13], buf[14], nvec); });
    float t_cm = time_ms([&] { k_commit<<<grid, 256>>>(buf[8], buf[9], buf[1], buf[7], buf[10], buf[11], c.dt, nvec); });
    float t_sw = time_ms([&] { k_commit_swap<<<grid, 256>>>(buf[15], buf[16], buf[1], buf[7], nvec); });
    printf("r1   k_reduce      %7.4f ms (%6.0f GB/s)   k_fit_commit %7.4f ms (%6.0f GB/s)   sum %7.4f ms\n", t_red, 8 * NS / t_red, t_fit, 15 * NS / t_fit, t_red + t_fit);
    printf("new  k_norm_fit    %7.4f ms (%6.0f GB/s)   k_commit     %7.4f ms (%6.0f GB/s)   sum %7.4f ms\n", t_nf, 11 * NS / t_nf, t_cm, 8 * NS / t_cm, t_nf + t_cm);
    printf("new  k_norm_fit    %7.4f ms                 k_commit_swap %7.4f ms (%6.0f GB/s)   sum %7.4f ms\n", t_nf, t_sw, 4 * NS / t_sw, t_nf + t_sw);
    return 0;
}
