// exp_graph_while.cu -- round-2 experiment (not part of libtdq): a whole adaptive solve as ONE graph launch.
//
// CUDA 12.4+ conditional graph nodes: a WHILE node whose body graph is re-executed while a device-side
// condition value is non-zero.  The plan for libtdq:
//   outer graph = [ WHILE(handle) { child graph = the step body torch captured (stage combines + func + norm +
//                   controller + fit + eval) } ]
//   k_controller calls cudaGraphSetConditional(handle, !halt) at the end of every attempt.
// torch exposes the captured body as a raw cudaGraph_t (torch.cuda.CUDAGraph(keep_graph=True).raw_cuda_graph()),
// which cudaGraphAddChildGraphNode clones into the body.
//
// This standalone program checks the two mechanisms the plan relies on, without torch:
//   (1) a kernel inside a CHILD graph of the while-body may set the condition handle of the outer graph;
//   (2) the loop terminates from the device and the host sees one launch.
// Expected output: "iterations=37 expected=37 OK".
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/exp_graph_while.bin scripts/exp_graph_while.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

struct State { int iterations; int limit; double t, dt; };

// stand-in for the step body's stream kernels
__global__ void k_work(State *s, float *buf, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) buf[i] = buf[i] * 1.0001f + 1.0f;
}

// stand-in for k_controller: advances the state and decides whether another attempt runs
__global__ void k_controller(State *s, cudaGraphConditionalHandle h) {
    s->iterations += 1;
    s->t += s->dt;
    const bool more = s->iterations < s->limit;
    cudaGraphSetConditional(h, more ? 1u : 0u);
}

int main() {
    State *s;
    float *buf;
    const int n = 1 << 20;
    CK(cudaMalloc(&s, sizeof(State)));
    CK(cudaMalloc(&buf, n * sizeof(float)));
    CK(cudaMemset(buf, 0, n * sizeof(float)));
    State h0 = {0, 37, 0.0, 0.25};
    CK(cudaMemcpy(s, &h0, sizeof(State), cudaMemcpyHostToDevice));

    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));

    // outer graph with a WHILE node
    cudaGraph_t outer;
    CK(cudaGraphCreate(&outer, 0));
    cudaGraphConditionalHandle handle;
    CK(cudaGraphConditionalHandleCreate(&handle, outer, 1 /* default: run the first iteration */,
                                        cudaGraphCondAssignDefault));
    cudaGraphNodeParams wp = {};
    wp.type = cudaGraphNodeTypeConditional;
    wp.conditional.handle = handle;
    wp.conditional.type = cudaGraphCondTypeWhile;
    wp.conditional.size = 1;
    cudaGraphNode_t while_node;
    CK(cudaGraphAddNode(&while_node, outer, nullptr, 0, &wp));
    cudaGraph_t body = wp.conditional.phGraph_out[0];

    // the step body, captured from a stream into its own graph (as torch would hand it to us) ...
    cudaGraph_t step;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
    k_work<<<(n + 255) / 256, 256, 0, st>>>(s, buf, n);
    k_controller<<<1, 1, 0, st>>>(s, handle);
    CK(cudaStreamEndCapture(st, &step));
    // ... and embedded as a CHILD graph node of the while body (mechanism 1)
    cudaGraphNode_t child;
    CK(cudaGraphAddChildGraphNode(&child, body, nullptr, 0, step));

    cudaGraphExec_t exec;
    CK(cudaGraphInstantiate(&exec, outer, 0));
    CK(cudaGraphLaunch(exec, st));                       // ONE launch for the whole loop (mechanism 2)
    CK(cudaStreamSynchronize(st));

    State h1;
    CK(cudaMemcpy(&h1, s, sizeof(State), cudaMemcpyDeviceToHost));
    printf("iterations=%d expected=%d %s  t=%g\n", h1.iterations, h0.limit, h1.iterations == h0.limit ? "OK" : "MISMATCH", h1.t);

    // a second launch must run again from the default condition value
    State h2 = {0, 5, 0.0, 0.5};
    CK(cudaMemcpy(s, &h2, sizeof(State), cudaMemcpyHostToDevice));
    CK(cudaGraphLaunch(exec, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaMemcpy(&h1, s, sizeof(State), cudaMemcpyDeviceToHost));
    printf("relaunch: iterations=%d expected=5 %s\n", h1.iterations, h1.iterations == 5 ? "OK" : "MISMATCH");
    return 0;
}
