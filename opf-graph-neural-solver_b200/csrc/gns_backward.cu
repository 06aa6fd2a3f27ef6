#include "gns_host.h"
namespace gns {
int backward_extra_floats(int, int, int, int, int, int) { return 0; }
int backward_ctas(const gns_plan* plan, const ModelDims&, const Geometry& g) { return std::min(g.nbatch, plan->num_sms); }
int run_backward(gns_plan*, const ModelDims&, const float*, const float*, const float*, const float*, long long, float,
                 const float*, const float*, const float*, const float*, float*, void*, long long, cudaStream_t) {
  set_error("backward not built yet");
  return -1;
}
}  // namespace gns
