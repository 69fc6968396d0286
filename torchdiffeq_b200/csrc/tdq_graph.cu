// tdq_graph.cu -- the adaptive loop as a device-side WHILE (CUDA conditional graph node).
//
// Reference: AdaptiveStepsizeODESolver.integrate / _advance (solvers.py:28-35, rk_common.py:243-250) loop on
// the host and decide `next_t > t1` with a device-to-host sync per iteration.  Here the body of the loop --
// one step attempt, captured by the caller as a CUDA graph -- becomes the body of a conditional WHILE node
// whose condition k_controller sets from the device (tdq_ctrl.cu), so a solve is one graph launch.
#include "tdq_common.cuh"

namespace {

struct TdqLoop {
    cudaGraph_t outer;
    cudaGraphExec_t exec;
    cudaGraphConditionalHandle handle;
};

}  // namespace

extern "C" {

int tdq_loop_create(void *body_graph, void **loop_out, uint64_t *handle_out) {
    TDQ_REQUIRE(body_graph && loop_out && handle_out, "null argument");
    TdqLoop *lp = new TdqLoop();
    memset(lp, 0, sizeof(*lp));
    cudaError_t e = cudaGraphCreate(&lp->outer, 0);
    if (e == cudaSuccess)
        // default value 1, re-assigned at every launch: the first iteration always runs (the caller does not
        // launch a loop for a solve that has nothing to integrate)
        e = cudaGraphConditionalHandleCreate(&lp->handle, lp->outer, 1, cudaGraphCondAssignDefault);
    cudaGraphNode_t while_node = nullptr, child = nullptr;
    if (e == cudaSuccess) {
        cudaGraphNodeParams wp = {};
        wp.type = cudaGraphNodeTypeConditional;
        wp.conditional.handle = lp->handle;
        wp.conditional.type = cudaGraphCondTypeWhile;
        wp.conditional.size = 1;
        e = cudaGraphAddNode(&while_node, lp->outer, nullptr, 0, &wp);
        if (e == cudaSuccess) {
            cudaGraph_t body = wp.conditional.phGraph_out[0];
            e = cudaGraphAddChildGraphNode(&child, body, nullptr, 0, (cudaGraph_t)body_graph);
        }
    }
    if (e == cudaSuccess) e = cudaGraphInstantiate(&lp->exec, lp->outer, 0);
    if (e != cudaSuccess) {
        tdq_set_error("tdq_loop_create: %s", cudaGetErrorString(e));
        if (lp->exec) cudaGraphExecDestroy(lp->exec);
        if (lp->outer) cudaGraphDestroy(lp->outer);
        delete lp;
        cudaGetLastError();
        return TDQ_ERR_CUDA;
    }
    *loop_out = lp;
    *handle_out = (uint64_t)lp->handle;
    return TDQ_OK;
}

int tdq_loop_launch(void *loop, void *stream) {
    TDQ_REQUIRE(loop, "null loop");
    TDQ_CHECK_CUDA(cudaGraphLaunch(((TdqLoop *)loop)->exec, (cudaStream_t)stream));
    return TDQ_OK;
}

int tdq_loop_destroy(void *loop) {
    if (!loop) return TDQ_OK;
    TdqLoop *lp = (TdqLoop *)loop;
    if (lp->exec) cudaGraphExecDestroy(lp->exec);
    if (lp->outer) cudaGraphDestroy(lp->outer);
    delete lp;
    return TDQ_OK;
}

}  // extern "C"
