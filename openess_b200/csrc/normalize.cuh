// normalize.cuh -- nonzero mean/std standardisation (data_util.py:38-48, inference_utils.py:77-85,
// representations.py:45-53).
#pragma once
#include "common.cuh"

// x: [n_groups, group_numel] in place; stats: device float64 [n_groups, 3] = {sum, sumsq, nnz}.
// phase 0 = stats + apply, 1 = stats only, 2 = apply only.
int launch_nonzero_standardize(float* x, int64_t group_numel, int n_groups, double* stats, int phase,
                               int unbiased, cudaStream_t st);
