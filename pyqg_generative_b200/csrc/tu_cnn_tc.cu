// tu_cnn_tc.cu -- translation unit of the tcgen05 convolution kernels and their host-side launch code (cnn_tc.cuh)
#include "cnn_tc.cuh"
