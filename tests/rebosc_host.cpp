// Host build of pfmds_b200/csrc/rebosc_core.cuh for tests/test_rebosc.py: the per-thread bodies of the rebosc kernels run in a
// serial loop over the thread index (CPU emulation of the launch), on the same ELL list layout the device uses.
#include <cstddef>
#include <cstring>

#include "../pfmds_b200/csrc/rebosc_core.cuh"

extern "C" {
// pos4/frc4: N records {x,y,z,w}; nlist: ELL [maxn][stride]; params: the 18 numbers of the parameter file; box: L[3]
int reb_host_run(int N, const double* pos4, double* frc4, size_t stride, const int* nlist, const int* nnum, const double* p, const double* L,
                 double* energy, int* err) {
    REBp P{p[0], p[1], p[2], {p[3], p[4], p[5]}, {p[6], p[7], p[8]}, p[9], {p[10], p[11], p[12], p[13], p[14], p[15]}, p[16], p[17]};
    BoxD box;
    for (int k = 0; k < 3; ++k) { box.L[k] = L[k]; box.h[k] = 0.5 * L[k]; }
    ListView lv{nlist, nnum, stride};
    const double4* pos = reinterpret_cast<const double4*>(pos4);
    double4* frc = reinterpret_cast<double4*>(frc4);
    static int orig_dummy[1 << 20];
    double e = 0.;
    for (int i = 0; i < N; ++i) e += reb_energy_thread(i, pos, lv, P, box);
    *energy = e;
    for (int t = 0; t < 3 * N; ++t) reb_numforce_thread(t / 3, t % 3, pos, frc, lv, P, box, orig_dummy, err);
    return 0;
}
}
