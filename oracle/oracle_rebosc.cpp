// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).  PARITY UNPINNED.
// Restates INTERACTION_POTENTIALS/REBOsolidcarbon.f90 (REBOsc_energy: the reference has no analytic rebosc force)
// and the generic numerical-force engine of MOLECULAR_DYNAMICS/md_interactions.f90:273-399 (truncated neighbour
// lists around one atom, +-dx shifts of the cached dr, central differences of the potential energy).
#include <omp.h>

#include <cmath>

#include "oracle.hpp"

namespace oracle {

static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// REBOsolidcarbon.f90:27-91.  bsp: sigma-pi bond order of bond p of atom i from the angles to i's other bonds; bdh: dihedral
// term from i's other bonds q and j's other bonds l; the pair energy is counted once, from the lower list-local index.
void REBOsc_energy(double& energy, const NeighbourList& nl, const REBOscParams& P) {
    const int m = nl.neighb_num_max, N = nl.N;
    std::vector<double> bsp((size_t)m * N), bdh((size_t)m * N);  // :31 automatic arrays
    energy = 0.;
#pragma omp parallel if (!omp_in_parallel())
    {
        double energy_priv = 0.;
#pragma omp for
        for (int i = 0; i < N; ++i) {
            for (int p = 0; p < m; ++p) { bsp[(size_t)i * m + p] = 0.; bdh[(size_t)i * m + p] = 0.; }  // :38-39 whole row, every call
            for (int p = 0; p < nl.nnum[i]; ++p) {
                size_t sp = (size_t)i * m + p;
                if (nl.moddr[sp] < P.R2) {
                    int j = nl.nlist[sp];
                    const double* dp = &nl.dr[3 * sp];
                    for (int q = 0; q < nl.nnum[i]; ++q) {
                        size_t sq = (size_t)i * m + q;
                        if (p != q && nl.moddr[sq] < P.R2) {
                            const double* dq = &nl.dr[3 * sq];
                            double cosine = dot3(dp, dq) / (nl.moddr[sp] * nl.moddr[sq]);
                            double c2 = cosine * cosine, c3 = c2 * cosine, c4 = c2 * c2, c5 = c4 * cosine;  // x**n by repeated products
                            bsp[sp] = bsp[sp] + f_cut(nl.moddr[sq], P.R1, P.R2) *
                                                    (P.g[0] + P.g[1] * cosine + P.g[2] * c2 + P.g[3] * c3 + P.g[4] * c4 + P.g[5] * c5);
                            for (int l = 0; l < nl.nnum[j]; ++l) {
                                size_t sl = (size_t)j * m + l;
                                if (nl.nlist[sl] != i && nl.moddr[sl] < P.R2) {
                                    const double* dl = &nl.dr[3 * sl];
                                    double aa = nl.moddr[sp] * nl.moddr[sp];
                                    double ab = dot3(dp, dq), ac = dot3(dp, dl), bc = dot3(dq, dl);
                                    double num = aa * bc - ab * ac;
                                    bdh[sp] = bdh[sp] + f_cut(nl.moddr[sq], P.R1, P.R2) * f_cut(nl.moddr[sl], P.R1, P.R2) *
                                                            (1. - num * num / (aa * (nl.moddr[sq] * nl.moddr[sq]) - ab * ab) /
                                                                      (aa * (nl.moddr[sl] * nl.moddr[sl]) - ac * ac));
                                }
                            }
                        }
                    }
                    bsp[sp] = std::pow(1. + bsp[sp], -0.5);
                    bdh[sp] = P.T * bdh[sp];
                }
            }
        }
#pragma omp for
        for (int i = 0; i < N; ++i)
            for (int p = 0; p < nl.nnum[i]; ++p) {
                size_t sp = (size_t)i * m + p;
                if (nl.moddr[sp] < P.R2) {
                    int j = nl.nlist[sp];
                    if (j > i) {
                        int q = 0;
                        for (; q < nl.nnum[j]; ++q)
                            if (nl.nlist[(size_t)j * m + q] == i) break;
                        // not found (only in truncated lists): the Fortran loop leaves q = nnum(j)+1, an entry zeroed at :38
                        double bq = (q < nl.nnum[j]) ? bsp[(size_t)j * m + q] : 0.;
                        double r = nl.moddr[sp];
                        energy_priv = energy_priv +
                                      f_cut(r, P.R1, P.R2) * ((1 + P.Q / r) * P.A * std::exp(-P.alpha * r) -
                                                              ((bsp[sp] + bq) / 2 + bdh[sp]) * (P.B[0] * std::exp(-P.beta[0] * r) + P.B[1] * std::exp(-P.beta[1] * r) +
                                                                                                  P.B[2] * std::exp(-P.beta[2] * r)));
                    }
                }
            }
#pragma omp atomic
        energy = energy + energy_priv;
    }
}

// ---- md_interactions.f90:313-399 ---------------------------------------------------------------
// :313-320
static void create_truncated_nl(NeighbourList& tnl, const NeighbourList& nl) {
    tnl.N = nl.N;
    tnl.neighb_num_max = nl.neighb_num_max;
    create_neighbour_list(tnl);
}
static void copy_row(NeighbourList& tnl, const NeighbourList& nl, int r) {
    const int m = nl.neighb_num_max;
    for (int p = 0; p < nl.nnum[r]; ++p) {
        size_t s = (size_t)r * m + p;
        tnl.nlist[s] = nl.nlist[s];
        tnl.moddr[s] = nl.moddr[s];
        for (int k = 0; k < 3; ++k) tnl.dr[3 * s + k] = nl.dr[3 * s + k];
    }
}
// :335-381  rows of atom i, of its neighbours up to order n-1 (whole rows), and for the outermost shell one back-pointer
// entry per atom (the `nnum(q)==0` test fails after the first).  `present` stands for particle_index /= 0 (indices are 0-based here).
static void calculate_truncated_nl(NeighbourList& tnl, std::vector<char>& present, const NeighbourList& nl, int i, int n) {
    const int m = nl.neighb_num_max;
    for (int j = 0; j < tnl.N; ++j) { tnl.particle_index[j] = 0; tnl.nnum[j] = 0; present[j] = 0; }
    tnl.N = nl.N;
    tnl.neighb_num_max = nl.neighb_num_max;
    copy_row(tnl, nl, i);
    tnl.nnum[i] = nl.nnum[i];
    tnl.particle_index[i] = nl.particle_index[i];
    present[i] = 1;
    for (int ni = 1; ni <= n; ++ni) {
        if (ni < n) {
            for (int j = 0; j < tnl.N; ++j)
                for (int p = 0; p < tnl.nnum[j]; ++p) {
                    int q = tnl.nlist[(size_t)j * m + p];
                    if (tnl.nnum[q] == 0) {
                        copy_row(tnl, nl, q);
                        tnl.particle_index[q] = nl.particle_index[q];
                        present[q] = 1;
                    }
                }
            for (int j = 0; j < tnl.N; ++j)
                if (present[j]) tnl.nnum[j] = nl.nnum[j];
        } else {
            for (int j = 0; j < tnl.N; ++j)
                for (int p = 0; p < tnl.nnum[j]; ++p) {
                    size_t s = (size_t)j * m + p;
                    int q = tnl.nlist[s];
                    if (tnl.nnum[q] == 0) {
                        tnl.nnum[q] = tnl.nnum[q] + 1;
                        size_t t = (size_t)q * m + (tnl.nnum[q] - 1);
                        tnl.nlist[t] = j;
                        for (int k = 0; k < 3; ++k) tnl.dr[3 * t + k] = -nl.dr[3 * s + k];
                        tnl.moddr[t] = nl.moddr[s];
                        tnl.particle_index[q] = nl.particle_index[q];
                        present[q] = 1;
                    }
                }
        }
    }
}
// :383-403  atom inl moves by +dx along k: its own dr shrink by dx, the dr of every row that points at it grow by dx
static void shift_drs(NeighbourList& tnl, int inl, int k, int nl_n, double dx) {
    const int m = tnl.neighb_num_max;
    for (int p = 0; p < tnl.nnum[inl]; ++p) {
        size_t s = (size_t)inl * m + p;
        tnl.dr[3 * s + k] = tnl.dr[3 * s + k] - dx;
        tnl.moddr[s] = std::sqrt(tnl.dr[3 * s] * tnl.dr[3 * s] + tnl.dr[3 * s + 1] * tnl.dr[3 * s + 1] + tnl.dr[3 * s + 2] * tnl.dr[3 * s + 2]);
    }
    if (nl_n == 1)
        for (int j = 0; j < tnl.N; ++j)
            for (int p = 0; p < tnl.nnum[j]; ++p) {
                size_t s = (size_t)j * m + p;
                if (tnl.nlist[s] == inl) {
                    tnl.dr[3 * s + k] = tnl.dr[3 * s + k] + dx;
                    tnl.moddr[s] = std::sqrt(tnl.dr[3 * s] * tnl.dr[3 * s] + tnl.dr[3 * s + 1] * tnl.dr[3 * s + 1] + tnl.dr[3 * s + 2] * tnl.dr[3 * s + 2]);
                }
            }
}

// :244-258 dispatch used by the numerical engine
static void energy_of(const Interaction& it, double& e, const NeighbourList& nl) {
    const std::string& nm = it.interaction_name;
    if (nm == "lj") LJ_energy(e, nl, it.lj);
    else if (nm == "lj1g") LJ1g_energy(e, nl, it.lj1g);
    else if (nm == "tb") TB_energy(e, nl, it.tb);
    else if (nm == "rebosc") REBOsc_energy(e, nl, it.rebosc);
    else if (nm == "rjl") RJL_energy(e, nl, it.rjl);
}

// :273-311  F_k(i) += (E(x_i - dx e_k) - E(x_i + dx e_k)) / 2 / dx on the truncated list around atom i
void calculate_forces_numerically(Particles& a, std::vector<Interaction>& its) {
    const double dx = std::pow(10., -6);
    for (auto& it : its) {
        if (!it.numerical_force) continue;
        const std::string& nm = it.interaction_name;
        if (nm == "ljc" || nm == "morsec") continue;  // :283 empty case in the reference
        for (int j = 0; j < it.nl_n; ++j) {
            const NeighbourList& nl = it.nl[(size_t)j];
#pragma omp parallel
            {
                NeighbourList tnl;
                create_truncated_nl(tnl, nl);
                std::vector<char> present((size_t)(nl.N > 0 ? nl.N : 0), 0);
#pragma omp for
                for (int inl = 0; inl < nl.N; ++inl) {
                    calculate_truncated_nl(tnl, present, nl, inl, it.neib_order);
                    for (int k = 0; k < 3; ++k) {
                        double e1, e2;
                        shift_drs(tnl, inl, k, it.nl_n, -dx);
                        energy_of(it, e1, tnl);
                        shift_drs(tnl, inl, k, it.nl_n, 2 * dx);
                        energy_of(it, e2, tnl);
                        if (k != 2) shift_drs(tnl, inl, k, it.nl_n, -dx);
                        a.forces[3 * nl.particle_index[inl] + k] = a.forces[3 * nl.particle_index[inl] + k] + (e1 - e2) / 2 / dx;
                    }
                }
            }
        }
    }
}

}  // namespace oracle
