// acq.cu -- FFT parallel-code-phase acquisition (placeholder until the kernels land)
#include "common.cuh"
void acq_free_workspace(gnssb200_handle *h) { (void)h; }
