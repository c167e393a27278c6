// cnn_tc.cuh -- tcgen05 implicit-GEMM convolution path (placeholder until the kernels land)
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "../../include/qgb200.h"
namespace qgb {
struct TcNet { bool ready = false; };
struct TcWorkspace {};
inline int tc_pack_net(TcNet& n, int, const qgb_cnn_layer*, std::string*) { n.ready = false; return 0; }
inline void tc_free_net(TcNet& n) { n.ready = false; }
inline void tc_free_workspace(TcWorkspace&) {}
inline int tc_launches_per_forward(const TcNet&) { return 0; }
inline int tc_forward(const TcNet&, TcWorkspace&, const float*, long long, float*, long long, int, int, int, int, int, int,
                      cudaStream_t, std::string* e) { *e = "tcgen05 path not built"; return QGB_EUNSUPPORTED; }
}  // namespace qgb
