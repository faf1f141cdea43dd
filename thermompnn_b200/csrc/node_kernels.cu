// Which form of the per-residue kernels a batch runs (the kernels themselves: node_mma_kernels.cu -- warp-level tensor
// cores inside 4-CTA clusters, small batches -- and tc_node_kernels.cu / tc_head_kernels.cu -- tcgen05, large batches).
#include "kernels.h"

namespace tmpnn {

// residues: up to here the clusters (one 16-residue tile each, 4 CTAs) fit in ONE wave on 148 SMs; above, the tcgen05
// forms win (their cost per 128-residue tile is flat)
constexpr int NODE_MMA_MAX = 593;

bool node_update_uses_tcgen05(int N) { return N >= NODE_MMA_MAX; }
bool head_uses_tcgen05(int N) { return N >= NODE_MMA_MAX; }

cudaError_t node_kernels_init() { return node_mma_kernels_init(); }

void launch_node_update(const NodeUpdArgs& a, int sms, cudaStream_t st) {
    // Inside a form a residue's result does not depend on the rest of the batch; the two forms agree to fp32 rounding.
    if (a.N < NODE_MMA_MAX) launch_node_update_mma(a, st);
    else launch_tc_node_update(a, sms, st);
}

}  // namespace tmpnn
