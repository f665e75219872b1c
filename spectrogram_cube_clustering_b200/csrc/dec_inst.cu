// dec_inst.cu — instantiates the DEC kernels for ONE latent dimension (-DSCC_DIM=<d>), so the
// nine dimensions build in parallel.  Exports dec_assign_dim<d> / dec_grad_dim<d>.
#include "dec_kernels.cuh"

#ifndef SCC_DIM
#error "compile with -DSCC_DIM=<latent dimension>"
#endif

namespace scc {

#define SCC_CAT_(a, b) a##b
#define SCC_CAT(a, b) SCC_CAT_(a, b)

int SCC_CAT(dec_assign_dim, SCC_DIM)(const DecArgs& a, cudaStream_t st) {
    const int kp = a.K <= 4 ? 4 : (a.K <= 8 ? 8 : 16);
    if (kp == 4) return DecOps<SCC_DIM, 4>::assign(a, st);
    if (kp == 8) return DecOps<SCC_DIM, 8>::assign(a, st);
    return DecOps<SCC_DIM, 16>::assign(a, st);
}

int SCC_CAT(dec_grad_dim, SCC_DIM)(const DecArgs& a, int mode, cudaStream_t st) {
    const int kp = a.K <= 4 ? 4 : (a.K <= 8 ? 8 : 16);
    if (mode == MODE_KL) {
        if (kp == 4) return DecOps<SCC_DIM, 4>::template grad<MODE_KL>(a, st);
        if (kp == 8) return DecOps<SCC_DIM, 8>::template grad<MODE_KL>(a, st);
        return DecOps<SCC_DIM, 16>::template grad<MODE_KL>(a, st);
    }
    if (mode == MODE_KLF) {
        if (kp == 4) return DecOps<SCC_DIM, 4>::template grad<MODE_KLF>(a, st);
        if (kp == 8) return DecOps<SCC_DIM, 8>::template grad<MODE_KLF>(a, st);
        return DecOps<SCC_DIM, 16>::template grad<MODE_KLF>(a, st);
    }
    if (mode == MODE_KLU) {
        if (kp == 4) return DecOps<SCC_DIM, 4>::template grad<MODE_KLU>(a, st);
        if (kp == 8) return DecOps<SCC_DIM, 8>::template grad<MODE_KLU>(a, st);
        return DecOps<SCC_DIM, 16>::template grad<MODE_KLU>(a, st);
    }
    if (mode == MODE_STEP) {
        if (kp == 4) return DecOps<SCC_DIM, 4>::template grad<MODE_STEP>(a, st);
        if (kp == 8) return DecOps<SCC_DIM, 8>::template grad<MODE_STEP>(a, st);
        return DecOps<SCC_DIM, 16>::template grad<MODE_STEP>(a, st);
    }
    if (mode == MODE_KMEANS) {
        if (kp == 4) return DecOps<SCC_DIM, 4>::template grad<MODE_KMEANS>(a, st);
        if (kp == 8) return DecOps<SCC_DIM, 8>::template grad<MODE_KMEANS>(a, st);
        return DecOps<SCC_DIM, 16>::template grad<MODE_KMEANS>(a, st);
    }
    if (kp == 4) return DecOps<SCC_DIM, 4>::template grad<MODE_GENERIC>(a, st);
    if (kp == 8) return DecOps<SCC_DIM, 8>::template grad<MODE_GENERIC>(a, st);
    return DecOps<SCC_DIM, 16>::template grad<MODE_GENERIC>(a, st);
}

}  // namespace scc
