// placeholder until the tcgen05 kernel lands
#include "common.cuh"
namespace bsdfdiff {
int launch_tc(const FlowParams&, cudaStream_t) { return -2; }
}
