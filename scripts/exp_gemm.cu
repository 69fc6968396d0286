// exp_gemm.cu -- experiment (not part of libtdq): what would the user's func of BASELINE configs[1] cost on the tensor
// cores WITHOUT giving up float32 accuracy?  k = y @ A^T with y [65536, 128], A [128, 128], float32, through cuBLAS 12.9:
//   CUBLAS_COMPUTE_32F                   the SIMT path torch takes (cutlass_80_simt_sgemm, 54.7 us in profiles/)
//   CUBLAS_COMPUTE_32F_FAST_TF32         1 x TF32 (10-bit mantissa: NOT acceptable for this solve)
//   CUBLAS_COMPUTE_32F_EMULATED_16BFX9   float32 emulated with 9 BF16 tensor-core products (Blackwell)
// Reports time per GEMM and the max / rms error against a float64 GEMM of the same inputs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/exp_gemm.bin scripts/exp_gemm.cu -lcublas
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
#define CB(x) do { cublasStatus_t s = (x); if (s != CUBLAS_STATUS_SUCCESS) { printf("cuBLAS error %d at %d\n", (int)s, __LINE__); return 1; } } while (0)

__global__ void to_double(const float *a, double *b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (double)a[i];
}
__global__ void err_kernel(const float *c, const double *ref, size_t n, double *out /* max, sumsq, refsumsq */) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double d = fabs((double)c[i] - ref[i]);
    atomicMax((unsigned long long *)&out[0], __double_as_longlong(d));      // non-negative doubles order like integers
    atomicAdd(&out[1], d * d);
    atomicAdd(&out[2], ref[i] * ref[i]);
}

int main() {
    const int M = 65536, N = 128, K = 128;
    std::vector<float> hy((size_t)M * K), ha((size_t)N * K);
    srand(1);
    for (auto &v : hy) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto &v : ha) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.3f;
    float *y, *a, *c;
    double *yd, *ad, *cd, *errs;
    CK(cudaMalloc(&y, sizeof(float) * M * K)); CK(cudaMalloc(&a, sizeof(float) * N * K)); CK(cudaMalloc(&c, sizeof(float) * M * N));
    CK(cudaMalloc(&yd, sizeof(double) * M * K)); CK(cudaMalloc(&ad, sizeof(double) * N * K)); CK(cudaMalloc(&cd, sizeof(double) * M * N));
    CK(cudaMalloc(&errs, 3 * sizeof(double)));
    CK(cudaMemcpy(y, hy.data(), sizeof(float) * M * K, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(a, ha.data(), sizeof(float) * N * K, cudaMemcpyHostToDevice));
    to_double<<<(M * K + 255) / 256, 256>>>(y, yd, (size_t)M * K);
    to_double<<<(N * K + 255) / 256, 256>>>(a, ad, (size_t)N * K);
    cublasHandle_t h;
    CB(cublasCreate(&h));
    // row-major k[M,N] = y[M,K] @ A^T[K,N]  ==  column-major C^T[N,M] = A[N,K](as op T of col-major KxN) * y^T
    const double one_d = 1.0, zero_d = 0.0;
    CB(cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, N, M, K, &one_d, ad, K, yd, K, &zero_d, cd, N));
    const float one = 1.f, zero = 0.f;
    struct { const char *name; cublasComputeType_t ct; } modes[] = {
        {"CUBLAS_COMPUTE_32F (SIMT)", CUBLAS_COMPUTE_32F},
        {"CUBLAS_COMPUTE_32F_FAST_TF32", CUBLAS_COMPUTE_32F_FAST_TF32},
        {"CUBLAS_COMPUTE_32F_EMULATED_16BFX9", CUBLAS_COMPUTE_32F_EMULATED_16BFX9},
    };
    for (auto &m : modes) {
        cublasStatus_t s = cublasGemmEx(h, CUBLAS_OP_T, CUBLAS_OP_N, N, M, K, &one, a, CUDA_R_32F, K, y, CUDA_R_32F, K, &zero, c,
                                        CUDA_R_32F, N, m.ct, CUBLAS_GEMM_DEFAULT);
        if (s != CUBLAS_STATUS_SUCCESS) { printf("%-40s not supported (status %d)\n", m.name, (int)s); continue; }
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        const int reps = 50;
        CK(cudaEventRecord(e0));
        for (int r = 0; r < reps; ++r)
            cublasGemmEx(h, CUBLAS_OP_T, CUBLAS_OP_N, N, M, K, &one, a, CUDA_R_32F, K, y, CUDA_R_32F, K, &zero, c, CUDA_R_32F, N,
                         m.ct, CUBLAS_GEMM_DEFAULT);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaMemset(errs, 0, 3 * sizeof(double)));
        err_kernel<<<(M * N + 255) / 256, 256>>>(c, cd, (size_t)M * N, errs);
        double he[3];
        CK(cudaMemcpy(he, errs, sizeof(he), cudaMemcpyDeviceToHost));
        printf("%-40s %8.2f us per GEMM   max|err| %.3e   rms err / rms ref %.3e\n", m.name, ms / reps * 1e3, he[0],
               sqrt(he[1] / he[2]));
    }
    return 0;
}
