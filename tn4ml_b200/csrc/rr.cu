// Register-resident small-bond family (float32, fp16x3 on mma.sync): host-side launch code.  Kernels: smpo_rr.cuh.
#include "host.h"

namespace tnb {

cudaError_t rr_set_attrs() { return cudaSuccess; }

}  // namespace tnb
