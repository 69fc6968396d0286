// tdq_shape.cuh -- host-side helpers shared by the launchers of libtdq's translation units.
#pragma once

#include "tdq_common.cuh"

#define TDQ_ROWS_H (TDQ_MAX_STAGES + 1)

// Host mirror of the tableau sparsity (so launchers can pick template arguments without reading device
// memory).  It is a function of the tableau only; launchers recompute it from the tdq_tableau the
// caller passes (cheap) instead of caching it per control block.
struct TdqHostShape {
    int valid;
    int n_stages, fsal;
    int row_nnz[TDQ_ROWS_H];
    int row_idx[TDQ_ROWS_H][TDQ_MAX_K];
    int err_nnz, mid_nnz;
    int err_idx[TDQ_MAX_K], mid_idx[TDQ_MAX_K];
};

void tdq_shape_from_tableau(const tdq_tableau *tab, TdqHostShape *h);   // tdq_stream.cu
int tdq_sm_count();                                                     // tdq_stream.cu

#define TDQ_DISPATCH_T(dtype, ...)                                         \
    do {                                                                   \
        if ((dtype) == TDQ_F32) { using T = float; __VA_ARGS__; }          \
        else if ((dtype) == TDQ_F64) { using T = double; __VA_ARGS__; }    \
        else { tdq_set_error("unsupported dtype %d", (int)(dtype)); return TDQ_ERR_INVALID; } \
    } while (0)
