// gmm_inst.cu — instantiates the fused EM kernel for ONE latent dimension (-DSCC_DIM=<d>).
#include "gmm_kernels.cuh"

#ifndef SCC_DIM
#error "compile with -DSCC_DIM=<latent dimension>"
#endif

namespace scc {

#define SCC_CAT_(a, b) a##b
#define SCC_CAT(a, b) SCC_CAT_(a, b)

int SCC_CAT(gmm_em_dim, SCC_DIM)(const GmmArgs& a, cudaStream_t st) {
    const int kp = a.K <= 4 ? 4 : (a.K <= 8 ? 8 : 16);
    if (kp == 4) return launch_gmm<SCC_DIM, 4>(a, st);
    if (kp == 8) return launch_gmm<SCC_DIM, 8>(a, st);
    return launch_gmm<SCC_DIM, 16>(a, st);
}

}  // namespace scc
