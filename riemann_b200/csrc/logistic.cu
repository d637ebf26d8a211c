#include "common.cuh"
SamplerImpl* make_logistic_sampler(rmn_sampler* s) { rmn_set_error("logistic sampler not built yet"); return nullptr; }
int logistic_pointwise(rmn_model* m, int which, int64_t n, const double* d_theta, double* d_out,
                       double* d_grad, double* d_metric, cudaStream_t st) { rmn_set_error("logistic not built yet"); return RMN_ERR_UNSUPPORTED; }
