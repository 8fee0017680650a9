// Host build of pfmds_b200/csrc/forces.cu for tests/test_kernels_host.py: the force / energy kernels (thread-per-atom variants) run
// as plain functions, one call per (block, thread) (tests/emu/host_emu.hpp), on the device's data layout — double4 records,
// ELL lists with slot indices — with the launch sequences of forces_interaction / energy_interaction (forces.cu, bottom).
// This checks kernel arithmetic and indexing against the CPU oracle without a GPU; it is not a CPU path of the product.
#include <algorithm>
#include <cstddef>
#include <vector>

#include "../pfmds_b200/csrc/forces.cu"

namespace {
BoxD make_box(const double* L) {
    BoxD b;
    for (int k = 0; k < 3; ++k) { b.L[k] = L[k]; b.h[k] = 0.5 * L[k]; }
    return b;
}
double sum_parts(std::vector<double>& part, int n, double scale) {
    double out = 0;
    emu_launch(k_sum_partials, 1, 1, 1024, n, (const double*)part.data(), scale, &out);
    return out;
}
}  // namespace

extern "C" {

// lj: lists 0 (g1 -> g2) and 1 (converse); energy from list 0 (LJ_energy on nl(1))
int fh_lj(int N, const double* pos4, double* frc4, size_t stride, const int* nl0, const int* nn0, const int* nl1, const int* nn1, const double* p,
          const double* L, double* energy) {
    const double4* pos = (const double4*)pos4;
    double4* frc = (double4*)frc4;
    BoxD box = make_box(L);
    LJp P{p[0], p[1], p[2], p[3]};
    const int nb = (N + FT - 1) / FT;
    std::vector<double> part((size_t)nb + 1, 0.);
    emu_launch(k_lj<true, true, 1>, nb, 1, FT, N, pos, frc, ListView{nl0, nn0, stride}, P, box, part.data());
    *energy = sum_parts(part, nb, 1.0);
    emu_launch(k_lj<true, false, 1>, nb, 1, FT, N, pos, frc, ListView{nl1, nn1, stride}, P, box, (double*)nullptr);
    return 0;
}

int fh_lj1g(int N, const double* pos4, double* frc4, size_t stride, const int* nl0, const int* nn0, const double* p, const double* L, double* energy, int pipelined) {
    const double4* pos = (const double4*)pos4;
    double4* frc = (double4*)frc4;
    BoxD box = make_box(L);
    double s2 = p[1] * p[1], s6 = s2 * s2 * s2, s12 = s6 * s6;  // as pfmds_add_interaction (capi.cu), LennardJones_1g.f90:21-24
    LJ1Gp P{p[2], p[3], 4. * p[0] * s6, 4. * p[0] * s12, 6. * 4. * p[0] * s6, 12. * 4. * p[0] * s12};
    const int nb = (N + FT - 1) / FT;
    std::vector<double> part((size_t)nb + 1, 0.);
    if (pipelined) emu_launch(k_lj1g_pipe<true>, nb, 1, FT, N, pos, frc, ListView{nl0, nn0, stride}, P, box, wrap_consts(box), part.data(), 0);
    else emu_launch(k_lj1g<true, true, 1>, nb, 1, FT, N, pos, frc, ListView{nl0, nn0, stride}, P, box, part.data());
    *energy = sum_parts(part, nb, 0.5);
    return 0;
}

// rjl: density pass (1/Eb into pos.w, energy partials) then force pass; `overwrite` = the store-instead-of-accumulate variant.
// gen = 2: second-generation pair routines when the parameters allow them (as forces_interaction decides), 1: first generation.
// Returns the generation that ran.  The list must carry two spare rows of valid slot numbers (capi.cu allocates them).
int fh_rjl(int N, double* pos4, double* frc4, size_t stride, const int* nl0, const int* nn0, const double* p, const double* L, int overwrite, double* energy,
           int gen) {
    double4* pos = (double4*)pos4;
    double4* frc = (double4*)frc4;
    BoxD box = make_box(L);
    RJLp R{p[0], p[1], p[2], p[3], p[4], p[5], p[6]};
    const WrapC W = wrap_consts(box);
    const int nb = (N + FT - 1) / FT;
    std::vector<double> part((size_t)nb + 1, 0.);
    const ListView lv{nl0, nn0, stride};
    if (gen == 3 && rjl_gen3_ok(R, box)) {   // third generation: node-table exponentials (the table lives in host memory here)
        const RjlTabSpec t = rjl_tab_spec(R);
        std::vector<double2> tab((size_t)(t.J_hi - t.J_lo + 1));
        rjl_tab_fill(R, t, tab.data());
        const RjlG G = rjl_g3_consts(R, t, tab.data());
        emu_launch(k_rjl_density<true, RjlG>, nb, 1, FT, N, pos, lv, G, box, W, part.data(), SlabDev{});
        *energy = sum_parts(part, nb, 1.0);
        if (overwrite >= 2) {   // the force pass that also yields the energy (k_rjl_force_e): energy from it, forces as usual
            std::fill(part.begin(), part.end(), 0.);
            emu_launch(k_rjl_force_e<RjlG>, nb, 1, FT, N, (const double4*)pos, frc, lv, G, box, W, SlabDev{}, overwrite & 1, R.r0 / (2. * R.p), R.xi, part.data());
            *energy = sum_parts(part, nb, 1.0);
        } else
        emu_launch(k_rjl_force<RjlG>, nb, 1, FT, N, (const double4*)pos, frc, lv, G, box, W, SlabDev{}, overwrite);
        return 3;
    }
    if (gen != 1 && rjl_gen2_ok(R, box)) {
        const RjlD CD = rjl_dens_consts(R);
        const RjlF CF = rjl_force_consts(R);
        emu_launch(k_rjl_density<true, RjlD>, nb, 1, FT, N, pos, lv, CD, box, W, part.data(), SlabDev{});
        *energy = sum_parts(part, nb, 1.0);
        emu_launch(k_rjl_force<RjlF>, nb, 1, FT, N, (const double4*)pos, frc, lv, CF, box, W, SlabDev{}, overwrite);
        return 2;
    }
    const RjlC C = rjl_consts(R);
    emu_launch(k_rjl_density<true, RjlC>, nb, 1, FT, N, pos, lv, C, box, W, part.data(), SlabDev{});
    *energy = sum_parts(part, nb, 1.0);
    emu_launch(k_rjl_force<RjlC>, nb, 1, FT, N, (const double4*)pos, frc, lv, C, box, W, SlabDev{}, overwrite);
    return 1;
}

// tb: bond orders, per-(slot, atom) force parts, reduction; energy from a second sweep (energy_interaction)
int fh_tb(int N, const double* pos4, double* frc4, size_t stride, int maxn, const int* nl0, const int* nn0, const double* p, const double* L, double* energy) {
    const double4* pos = (const double4*)pos4;
    double4* frc = (double4*)frc4;
    BoxD box = make_box(L);
    TBp T{p[0], p[1], p[2], p[3], p[4], p[5], p[6] * p[6], p[7] * p[7], p[8], p[9]};  // as pfmds_add_interaction, TersoffBrenner.f90:19-20
    const int nb = (N + FT - 1) / FT;
    std::vector<double> aux((size_t)maxn * stride, 0.), aux2((size_t)maxn * stride, 0.), part((size_t)nb * maxn + 16, 0.);
    std::vector<double4> fpart((size_t)maxn * stride, double4{0, 0, 0, 0});
    ListView lv{nl0, nn0, stride};
    emu_launch(k_tb_bond, nb, maxn, FT, N, pos, lv, T, box, aux.data(), aux2.data());
    emu_launch(k_tb_force<true, false>, nb, maxn, FT, N, pos, fpart.data(), lv, T, box, (const double*)aux.data(), (const double*)aux2.data(), (double*)nullptr);
    emu_launch(k_tb_reduce, nb, 1, FT, N, (const double4*)fpart.data(), frc, lv);
    emu_launch(k_tb_force<false, true>, nb, maxn, FT, N, pos, fpart.data(), lv, T, box, (const double*)aux.data(), (const double*)aux2.data(), part.data());
    *energy = sum_parts(part, nb * maxn, 1.0);
    return 0;
}

// ljc / morsec: normals, graphene side (+ T vectors), normal-derivative term, metal side; energy from the graphene list
int fh_cos(int morse, int N, const double* pos4, double* frc4, size_t stride, const int* nl0, const int* nn0, const int* nl1, const int* nn1, const int* nl2,
           const int* nn2, const double* p, const double* L, double* gnorm4, double* energy) {
    const double4* pos = (const double4*)pos4;
    double4* frc = (double4*)frc4;
    double4* gnorm = (double4*)gnorm4;
    BoxD box = make_box(L);
    CosP P{};
    int simp;
    if (!morse) { P.pe = 4. * p[0]; P.sig = p[1]; P.delt = p[2]; P.R1 = p[3]; P.R2 = p[4]; simp = p[5] != 0.; }       // cosp_of(), forces.cu
    else { P.pe = p[0]; P.r0 = p[1]; P.a = p[2]; P.delt = p[3]; P.R1 = p[4]; P.R2 = p[5]; simp = p[6] != 0.; }
    const int nb = (N + FT - 1) / FT;
    std::vector<double4> tvec(stride, double4{0, 0, 0, 0});
    std::vector<double> part((size_t)nb + 1, 0.);
    ListView l0{nl0, nn0, stride}, l1{nl1, nn1, stride}, l2{nl2, nn2, stride};
    emu_launch(k_normals, nb, 1, FT, N, pos, l2, box, simp, gnorm);
    if (!morse) {
        emu_launch(k_cos_direct<false, true, true, false, 1>, nb, 1, FT, N, pos, frc, l0, P, box, gnorm, tvec.data(), (double*)nullptr);
        if (!simp) emu_launch(k_cos_indirect, nb, 1, FT, N, pos, frc, l2, P.pe * P.delt, box, gnorm, tvec.data());
        emu_launch(k_cos_direct<false, false, true, false, 1>, nb, 1, FT, N, pos, frc, l1, P, box, gnorm, tvec.data(), (double*)nullptr);
        emu_launch(k_cos_direct<false, true, false, true, 1>, nb, 1, FT, N, pos, frc, l0, P, box, gnorm, tvec.data(), part.data());
    } else {
        emu_launch(k_cos_direct<true, true, true, false, 1>, nb, 1, FT, N, pos, frc, l0, P, box, gnorm, tvec.data(), (double*)nullptr);
        if (!simp) emu_launch(k_cos_indirect, nb, 1, FT, N, pos, frc, l2, 2. * P.pe * P.delt, box, gnorm, tvec.data());
        emu_launch(k_cos_direct<true, false, true, false, 1>, nb, 1, FT, N, pos, frc, l1, P, box, gnorm, tvec.data(), (double*)nullptr);
        emu_launch(k_cos_direct<true, true, false, true, 1>, nb, 1, FT, N, pos, frc, l0, P, box, gnorm, tvec.data(), part.data());
    }
    *energy = sum_parts(part, nb, 1.0);
    return 0;
}
}
