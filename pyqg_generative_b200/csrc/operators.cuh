// operators.cuh -- coarse-graining operators (placeholder until the kernels land)
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "../../include/qgb200.h"
namespace qgb {
inline int op_coarsegrain(int, int, int, int, int, const double*, double*, int, cudaStream_t, long long*, std::string* e) {
  *e = "coarse-graining kernels not built"; return QGB_EUNSUPPORTED; }
inline int op_subgrid_forcing(const qgb_config*, int, int, int, const double*, double*, double*, double*, double*, double*,
                              int, cudaStream_t, long long*, std::string* e) {
  *e = "coarse-graining kernels not built"; return QGB_EUNSUPPORTED; }
}  // namespace qgb
