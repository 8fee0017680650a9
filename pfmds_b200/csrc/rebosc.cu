// pfmds_b200 — kernels and launchers of `rebosc`; the arithmetic (one pair term, the per-thread energy sum and the per-thread
// central difference) lives in rebosc_core.cuh, which tests/rebosc_host.cpp also runs on the host, thread by thread.
#include <string>

#include "ctx.hpp"
#include "rebosc_core.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw std::string("CUDA: ") + cudaGetErrorString(e_) + " at " #x; } while (0)
#define RT 128

__global__ void __launch_bounds__(RT) k_rebosc_energy(int N, const double4* __restrict__ pos, ListView lv, REBp P, BoxD box, double* __restrict__ part) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double e = i < N ? reb_energy_thread(i, pos, lv, P, box) : 0.;
    e = block_sum(e);
    if (threadIdx.x == 0) part[blockIdx.x] = e;
}

// thread t -> atom m = t/3, axis k = t%3; the three threads of an atom write different components of frc[m]
__global__ void __launch_bounds__(RT) k_rebosc_numforce(int N, const double4* __restrict__ pos, double4* __restrict__ frc, ListView lv, REBp P, BoxD box,
                                                        const int* __restrict__ orig, int* err) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * N) return;
    const int m = t / 3;
    reb_numforce_thread(m, t - 3 * m, pos, frc, lv, P, box, orig, err);
}
void rebosc_forces(pfmds_ctx* c, Inter& it) {
    const int N = c->N;
    KTimer kt(c, KS_REBOSC_FORCE);
    LAUNCH((k_rebosc_numforce), (3 * N + RT - 1) / RT, RT, c->fst ? c->fst : c->st, N, c->pos, c->fout ? c->fout : c->frc, it.nl[0].view(c->stride), it.reb, c->box, c->orig, c->err);
    c->launches += 1;
    CK(cudaGetLastError());
}
// block partials into c->part; returns their number
int rebosc_energy_partials(pfmds_ctx* c, Inter& it) {
    const int N = c->N, nb = (N + RT - 1) / RT;
    KTimer kt(c, KS_REBOSC_ENERGY);
    LAUNCH((k_rebosc_energy), nb, RT, c->st, N, c->pos, it.nl[0].view(c->stride), it.reb, c->box, c->part);
    c->launches += 1;
    CK(cudaGetLastError());
    return nb;
}
