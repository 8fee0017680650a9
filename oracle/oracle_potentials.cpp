// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).  PARITY UNPINNED.
// Restates INTERACTION_POTENTIALS/{cut_off_function,cut_off_poly,LennardJones,LennardJones_1g,
// graphenenorm,LennardJonesCosine,MorseCosine,RosatoGuillopeLegrand,TersoffBrenner}.f90 and the
// dispatch in MOLECULAR_DYNAMICS/md_interactions.f90:138-271.  All routines read only the cached
// nl.dr / nl.moddr, never positions, exactly like the reference.
#include <omp.h>

#include <cmath>

#include "oracle.hpp"

namespace oracle {

static inline double pw2(double x) { return x * x; }
static inline double pw6(double x) { double x2 = x * x; return x2 * x2 * x2; }  // x**6 with an integer exponent
static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// ---- cut_off_function.f90:6-28 (pi is the reference's 15-digit literal) ---------------------
double f_cut(double r, double R1, double R2) {
    const double pi = 3.14159265358979;
    if (r < R1) return 1.;
    else if (r < R2) return (1. + std::cos(pi * (r - R1) / (R2 - R1))) / 2;
    return 0.;
}
double df_cut(double r, double R1, double R2) {  // f'(r)/r
    const double pi = 3.14159265358979;
    if (r < R1) return 0.;
    else if (r < R2) return -std::sin(pi * (r - R1) / (R2 - R1)) * pi / (R2 - R1) / r / 2;
    return 0.;
}
// ---- cut_off_poly.f90:6-17, 30-43 -----------------------------------------------------------
double f_cut_poly(double r, double R1, double R2) {
    if (r > R1 && r < R2) {
        double a = r - R1, w = R2 - R1;
        return 1. + (-10. * (a * a * a) * (w * w) + 15. * (a * a * a * a) * w - 6. * (a * a * a * a * a)) / (w * w * w * w * w);
    } else if (r <= R1) return 1.;
    return 0.;
}
// dfr lacks the 1/(R2-R1) factor in the reference (:41); reproduced literally.
void f_dfr_cut(double& f, double& dfr, double r, double R1, double R2) {
    double tempr = r;
    tempr = std::fmax(tempr, R1);
    tempr = std::fmin(tempr, R2);
    double x = (tempr - R1) / (R2 - R1);
    double x2 = x * x;
    f = 1. + x2 * x * (-10. + 15. * x - 6. * x2);
    dfr = tempr * x2 * (-30. + 60. * x - 30. * x2);
}

// ---- LennardJones.f90 ------------------------------------------------------------------------
// :23-46
void LJ_energy(double& energy, const NeighbourList& nl, const LJParams& P) {
    energy = 0.;
    const int m = nl.neighb_num_max;
#pragma omp parallel
    {
        double priv = 0.;
#pragma omp for
        for (int i = 0; i < nl.N; ++i)
            for (int p = 0; p < nl.nnum[i]; ++p) {
                double r = nl.moddr[(size_t)i * m + p];
                if (r < P.R2) {
                    double V = pw6(P.sig / r);
                    priv = priv + 4 * P.eps * V * (V - 1.) * f_cut(r, P.R1, P.R2);
                }
            }
#pragma omp atomic
        energy = energy + priv;
    }
}
// :48-69
void LJ_forces(Particles& a, const NeighbourList& nl, const LJParams& P) {
    const int m = nl.neighb_num_max;
#pragma omp parallel for
    for (int i = 0; i < nl.N; ++i)
        for (int p = 0; p < nl.nnum[i]; ++p) {
            size_t s = (size_t)i * m + p;
            double r = nl.moddr[s];
            if (r < P.R2) {
                double V = pw6(P.sig / r);
                double c = 4. * P.eps * (V * (12. * V - 6.) / (r * r) * f_cut(r, P.R1, P.R2) - V * (V - 1.) * df_cut(r, P.R1, P.R2));
                double* f = &a.forces[3 * nl.particle_index[i]];
                for (int k = 0; k < 3; ++k) f[k] = f[k] - c * nl.dr[3 * s + k];
            }
        }
}

// ---- LennardJones_1g.f90 ---------------------------------------------------------------------
// :21-24
void LJ1g_finish_parameters(LJ1gParams& P) {
    double s6 = pw6(P.sig), s12 = s6 * s6;
    P.c6 = 4. * P.eps * s6;
    P.c12 = 4. * P.eps * s12;
    P.c6t6 = 6. * 4. * P.eps * s6;
    P.c12t12 = 12. * 4. * P.eps * s12;
}
// :28-52  half list: p <= lessnnum(i)
void LJ1g_energy(double& energy, const NeighbourList& nl, const LJ1gParams& P) {
    energy = 0.;
    const int m = nl.neighb_num_max;
#pragma omp parallel
    {
        double priv = 0.;
#pragma omp for schedule(dynamic, 64)
        for (int i = 0; i < nl.N; ++i)
            for (int p = 0; p < nl.lessnnum[i]; ++p) {
                double r = nl.moddr[(size_t)i * m + p];
                double U = 1. / (r * r * r * r * r * r);
                priv = priv + U * (P.c12 * U - P.c6) * f_cut_poly(r, P.R1, P.R2);
            }
#pragma omp atomic
        energy = energy + priv;
    }
}
// :105-117
static inline double scalar_lj_force(double r, double R1, double R2, double c12, double c6, double c12t12, double c6t6) {
    double invr2 = 1. / (r * r);
    double U = invr2 * invr2 * invr2;
    double fcut, dfrcut;
    f_dfr_cut(fcut, dfrcut, r, R1, R2);
    return U * invr2 * ((c12t12 * U - c6t6) * fcut - (c12 * U - c6) * dfrcut);
}
// :54-103  Newton-3 scatter into per-thread force buffers, then atomic merge
void LJ1g_forces(Particles& a, const NeighbourList& nl, const LJ1gParams& P) {
    const int m = nl.neighb_num_max;
#pragma omp parallel
    {
        std::vector<double> priv_force((size_t)3 * a.N, 0.), F((size_t)m), fp((size_t)3 * m);
#pragma omp for schedule(dynamic, 64)
        for (int i = 0; i < nl.N; ++i) {
            const int h = nl.lessnnum[i];
            for (int p = 0; p < h; ++p) F[p] = scalar_lj_force(nl.moddr[(size_t)i * m + p], P.R1, P.R2, P.c12, P.c6, P.c12t12, P.c6t6);
            for (int p = 0; p < h; ++p)
                for (int k = 0; k < 3; ++k) fp[3 * p + k] = F[p] * nl.dr[3 * ((size_t)i * m + p) + k];
            for (int p = 0; p < h; ++p) {
                int ind = nl.particle_index[i];
                int jnd = nl.particle_index[nl.nlist[(size_t)i * m + p]];
                for (int k = 0; k < 3; ++k) {
                    priv_force[3 * ind + k] = priv_force[3 * ind + k] - fp[3 * p + k];
                    priv_force[3 * jnd + k] = priv_force[3 * jnd + k] + fp[3 * p + k];
                }
            }
        }
        for (int i = 0; i < a.N; ++i)
            for (int k = 0; k < 3; ++k) {
#pragma omp atomic
                a.forces[3 * i + k] = a.forces[3 * i + k] + priv_force[3 * i + k];
            }
    }
}

// ---- graphenenorm.f90 ------------------------------------------------------------------------
// :8-36  first three entries of the carbon list closer than nl_nn.r_cut; exactly three required
void find_gr_nearest_neighbors(NeighbourList& nn, const NeighbourList& nl) {
    const int nnum_nn = 3;
    if (nn.neighb_num_max != nnum_nn) throw StopError("error: nl_nn%neighb_num_max/=nnum_nn");
    const int m = nl.neighb_num_max;
    std::string err;
#pragma omp parallel for
    for (int i = 0; i < nl.N; ++i) {
        int k = 0;
        for (int p = 0; p < nl.nnum[i]; ++p) {
            size_t s = (size_t)i * m + p;
            if (nl.moddr[s] < nn.r_cut) {
                k = k + 1;
                if (k <= nnum_nn) {
                    size_t d = (size_t)i * nnum_nn + (k - 1);
                    nn.nlist[d] = nl.nlist[s];
                    nn.moddr[d] = nl.moddr[s];
                    for (int c = 0; c < 3; ++c) nn.dr[3 * d + c] = nl.dr[3 * s + c];
                } else {
#pragma omp critical
                    if (err.empty()) err = "error: too many gr nearest neibs " + std::to_string(i + 1);
                }
            }
        }
        if (k != nnum_nn) {
#pragma omp critical
            if (err.empty()) err = "error: not enough gr nearest neibs " + std::to_string(i + 1);
        }
        nn.nnum[i] = nnum_nn;
    }
    if (!err.empty()) throw StopError(err);
}
// :38-56  n = (d2-d1) x (d1-d3), flipped to n_z >= 0, normalised
void find_norm_in_graphene(std::vector<double>& gr_norm, const std::vector<double>& dr_nn, int n_rows, int maxn) {
#pragma omp parallel for
    for (int i = 0; i < n_rows; ++i) {
        const double* d1 = &dr_nn[3 * ((size_t)i * maxn + 0)];
        const double* d2 = &dr_nn[3 * ((size_t)i * maxn + 1)];
        const double* d3 = &dr_nn[3 * ((size_t)i * maxn + 2)];
        double drj12[3], drj31[3];
        for (int k = 0; k < 3; ++k) { drj12[k] = d2[k] - d1[k]; drj31[k] = d1[k] - d3[k]; }
        double* n = &gr_norm[3 * (size_t)i];
        for (int k = 0; k < 3; ++k) {  // Fortran k=1..3: (k mod 3)+1, ((k+1) mod 3)+1
            int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
            n[k] = drj12[k1] * drj31[k2] - drj12[k2] * drj31[k1];
        }
        if (n[2] < 0.) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
        double len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        for (int k = 0; k < 3; ++k) n[k] = n[k] / len;
    }
}

// ---- LennardJonesCosine.f90 / MorseCosine.f90 -------------------------------------------------
// The two modules share their structure; V2 and the prefactors differ.
// LJC :26-51
void LJC_energy(double& energy, const NeighbourList& nl, const LJCParams& P) {
    energy = 0.;
    const int m = nl.neighb_num_max;
#pragma omp parallel
    {
        double priv = 0.;
#pragma omp for
        for (int i = 0; i < nl.N; ++i)
            for (int p = 0; p < nl.nnum[i]; ++p) {
                size_t s = (size_t)i * m + p;
                double r = nl.moddr[s];
                if (r < P.R2) {
                    double V2 = pw6(P.sig / r);
                    double V1 = V2 * V2;
                    double V3 = std::pow(std::fabs(dot3(&P.gr_norm[3 * (size_t)i], &nl.dr[3 * s])) / r, P.delt);
                    priv = priv + 4 * P.eps * (V1 - V2 * V3) * f_cut(r, P.R1, P.R2);
                }
            }
#pragma omp atomic
        energy = energy + priv;
    }
}
// MorseC :26-51
void MorseC_energy(double& energy, const NeighbourList& nl, const MorseCParams& P) {
    energy = 0.;
    const int m = nl.neighb_num_max;
#pragma omp parallel
    {
        double priv = 0.;
#pragma omp for
        for (int i = 0; i < nl.N; ++i)
            for (int p = 0; p < nl.nnum[i]; ++p) {
                size_t s = (size_t)i * m + p;
                double r = nl.moddr[s];
                if (r < P.R2) {
                    double V2 = std::exp(-P.a * (r - P.r));
                    double V1 = V2 * V2;
                    double V3 = std::pow(std::fabs(dot3(&P.gr_norm[3 * (size_t)i], &nl.dr[3 * s])) / r, P.delt);
                    priv = priv + P.d * (V1 - 2. * V2 * V3) * f_cut(r, P.R1, P.R2);
                }
            }
#pragma omp atomic
        energy = energy + priv;
    }
}

// second loop of *_forces_for_graphene (LJC :81-106, MorseC :81-106): derivative of the normal of
// neighbour j with respect to the position of atom i.  `pref(V2)` = 4*eps*delt*V2 (ljc) or 2*d*delt*V2 (morsec).
template <class V2F>
static void normal_derivative_forces(Particles& a, const NeighbourList& nl, const NeighbourList& nn, const std::vector<double>& gr_norm,
                                     int nnum_nn, double R1, double R2, double delt, double pref, V2F v2f) {
    const int m = nl.neighb_num_max;
    std::string err;
#pragma omp parallel for
    for (int i = 0; i < nl.N; ++i) {
        for (int q = 0; q < nnum_nn; ++q) {
            int j = nn.nlist[(size_t)i * 3 + q];
            int l1 = 0;
            for (; l1 < nnum_nn; ++l1)
                if (nn.nlist[(size_t)j * 3 + l1] == i) break;
            if (l1 > 2) {
#pragma omp critical
                err = "l1>3";
                continue;
            }
            int l2 = (l1 + 1) % 3, l3 = (l1 + 2) % 3;
            const double* e1 = &nn.dr[3 * ((size_t)j * 3 + l1)];
            const double* e2 = &nn.dr[3 * ((size_t)j * 3 + l2)];
            const double* e3 = &nn.dr[3 * ((size_t)j * 3 + l3)];
            double drj12[3], drj31[3], drj23[3];
            for (int k = 0; k < 3; ++k) { drj12[k] = e2[k] - e1[k]; drj31[k] = e1[k] - e3[k]; drj23[k] = e3[k] - e2[k]; }
            const double* nj = &gr_norm[3 * (size_t)j];
            for (int p = 0; p < nl.nnum[j]; ++p) {
                size_t s = (size_t)j * m + p;
                double r = nl.moddr[s];
                if (r < R2) {
                    const double* d = &nl.dr[3 * s];
                    double V2 = v2f(r);
                    double nd = dot3(nj, d);
                    double V3 = std::pow(std::fabs(nd) / r, delt);
                    double f_c = f_cut(r, R1, R2);
                    double s2331 = dot3(drj23, drj31), s2312 = dot3(drj23, drj12);
                    double w = 0.;
                    for (int k = 0; k < 3; ++k) w += d[k] * (drj12[k] * s2331 - drj31[k] * s2312);
                    double den = dot3(drj12, drj12) * dot3(drj31, drj31) - pw2(dot3(drj12, drj31));
                    double c = pref * delt * V2 * V3 * f_c / den / nd;
                    double* f = &a.forces[3 * nl.particle_index[i]];
                    for (int k = 0; k < 3; ++k) f[k] = f[k] - c * nj[k] * w;
                }
            }
        }
    }
    if (!err.empty()) throw StopError(err);
}

// LJC :53-110
void LJC_forces_for_graphene(Particles& a, const NeighbourList& nl, const NeighbourList& nn, LJCParams& P) {
    if (nl.N != nn.N && nl.N != 0) throw StopError("error: nl%N/=nl_nn%N");
    int nnum_nn = 3;
    if (P.simplified) {
        nnum_nn = 0;
        for (size_t i = 0; i < P.gr_norm.size() / 3; ++i) { P.gr_norm[3 * i] = 0.; P.gr_norm[3 * i + 1] = 0.; P.gr_norm[3 * i + 2] = 1.; }
    }
    const int m = nl.neighb_num_max;
#pragma omp parallel for
    for (int i = 0; i < nl.N; ++i)
        for (int p = 0; p < nl.nnum[i]; ++p) {
            size_t s = (size_t)i * m + p;
            double r = nl.moddr[s];
            if (r < P.R2) {
                const double* n = &P.gr_norm[3 * (size_t)i];
                const double* d = &nl.dr[3 * s];
                double V2 = pw6(P.sig / r), V1 = V2 * V2;
                double nd = dot3(n, d);
                double V3 = std::pow(std::fabs(nd) / r, P.delt);
                double f_c = f_cut(r, P.R1, P.R2), df_c = df_cut(r, P.R1, P.R2);
                double cr = (12. * V1 - (6. + P.delt) * V2 * V3) / (r * r) * f_c - (V1 - V2 * V3) * df_c;
                double cn = P.delt * V2 * V3 / nd * f_c;
                double* f = &a.forces[3 * nl.particle_index[i]];
                for (int k = 0; k < 3; ++k) f[k] = f[k] - 4. * P.eps * (cr * d[k] + cn * n[k]);
            }
        }
    const double sig = P.sig;
    normal_derivative_forces(a, nl, nn, P.gr_norm, nnum_nn, P.R1, P.R2, P.delt, 4. * P.eps, [sig](double r) { return pw6(sig / r); });
}
// LJC :112-139
void LJC_forces_for_other_atoms(Particles& a, const NeighbourList& nl, const LJCParams& P) {
    const int m = nl.neighb_num_max;
#pragma omp parallel for
    for (int i = 0; i < nl.N; ++i)
        for (int p = 0; p < nl.nnum[i]; ++p) {
            size_t s = (size_t)i * m + p;
            double r = nl.moddr[s];
            if (r < P.R2) {
                const double* n = &P.gr_norm[3 * (size_t)nl.nlist[s]];
                const double* d = &nl.dr[3 * s];
                double V2 = pw6(P.sig / r), V1 = V2 * V2;
                double nd = dot3(n, d);
                double V3 = std::pow(std::fabs(nd) / r, P.delt);
                double f_c = f_cut(r, P.R1, P.R2), df_c = df_cut(r, P.R1, P.R2);
                double cr = (12. * V1 - (6. + P.delt) * V2 * V3) / (r * r) * f_c - (V1 - V2 * V3) * df_c;
                double cn = P.delt * V2 * V3 / nd * f_c;
                double* f = &a.forces[3 * nl.particle_index[i]];
                for (int k = 0; k < 3; ++k) f[k] = f[k] - 4. * P.eps * (cr * d[k] + cn * n[k]);
            }
        }
}
// MorseC :53-110
void MorseC_forces_for_graphene(Particles& a, const NeighbourList& nl, const NeighbourList& nn, MorseCParams& P) {
    if (nl.N != nn.N && nl.N != 0) throw StopError("error: nl%N/=nl_nn%N");
    int nnum_nn = 3;
    if (P.simplified) {
        nnum_nn = 0;
        for (size_t i = 0; i < P.gr_norm.size() / 3; ++i) { P.gr_norm[3 * i] = 0.; P.gr_norm[3 * i + 1] = 0.; P.gr_norm[3 * i + 2] = 1.; }
    }
    const int m = nl.neighb_num_max;
#pragma omp parallel for
    for (int i = 0; i < nl.N; ++i)
        for (int p = 0; p < nl.nnum[i]; ++p) {
            size_t s = (size_t)i * m + p;
            double r = nl.moddr[s];
            if (r < P.R2) {
                const double* n = &P.gr_norm[3 * (size_t)i];
                const double* d = &nl.dr[3 * s];
                double V2 = std::exp(-P.a * (r - P.r)), V1 = V2 * V2;
                double nd = dot3(n, d);
                double V3 = std::pow(std::fabs(nd) / r, P.delt);
                double f_c = f_cut(r, P.R1, P.R2), df_c = df_cut(r, P.R1, P.R2);
                double cr = 2. * (P.a * V1 - (P.a + P.delt / r) * V2 * V3) / r * f_c - (V1 - 2. * V2 * V3) * df_c;
                double cn = 2. * P.delt * V2 * V3 / nd * f_c;
                double* f = &a.forces[3 * nl.particle_index[i]];
                for (int k = 0; k < 3; ++k) f[k] = f[k] - P.d * (cr * d[k] + cn * n[k]);
            }
        }
    const double aa = P.a, rr = P.r;
    normal_derivative_forces(a, nl, nn, P.gr_norm, nnum_nn, P.R1, P.R2, P.delt, 2. * P.d, [aa, rr](double r) { return std::exp(-aa * (r - rr)); });
}
// MorseC :112-139  (V3 additionally divides by |n|, :127)
void MorseC_forces_for_other_atoms(Particles& a, const NeighbourList& nl, const MorseCParams& P) {
    const int m = nl.neighb_num_max;
#pragma omp parallel for
    for (int i = 0; i < nl.N; ++i)
        for (int p = 0; p < nl.nnum[i]; ++p) {
            size_t s = (size_t)i * m + p;
            double r = nl.moddr[s];
            if (r < P.R2) {
                const double* n = &P.gr_norm[3 * (size_t)nl.nlist[s]];
                const double* d = &nl.dr[3 * s];
                double V2 = std::exp(-P.a * (r - P.r)), V1 = V2 * V2;
                double nd = dot3(n, d);
                double V3 = std::pow(std::fabs(nd) / (std::sqrt(dot3(n, n)) * r), P.delt);
                double f_c = f_cut(r, P.R1, P.R2), df_c = df_cut(r, P.R1, P.R2);
                double cr = 2. * (P.a * V1 - (P.a + P.delt / r) * V2 * V3) / r * f_c - (V1 - 2. * V2 * V3) * df_c;
                double cn = 2. * P.delt * V2 * V3 / nd * f_c;
                double* f = &a.forces[3 * nl.particle_index[i]];
                for (int k = 0; k < 3; ++k) f[k] = f[k] - P.d * (cr * d[k] + cn * n[k]);
            }
        }
}

// ---- RosatoGuillopeLegrand.f90 ----------------------------------------------------------------
// :23-49
void RJL_energy(double& energy, const NeighbourList& nl, const RJLParams& P) {
    energy = 0.;
    const int m = nl.neighb_num_max;
#pragma omp parallel
    {
        double priv = 0.;
#pragma omp for
        for (int i = 0; i < nl.N; ++i) {
            double Eb2 = 0., Er = 0.;
            for (int p = 0; p < nl.nnum[i]; ++p) {
                double r = nl.moddr[(size_t)i * m + p];
                if (r <= P.R2) {
                    Eb2 = Eb2 + std::exp(-2. * P.q * (r / P.r0 - 1.)) * f_cut(r, P.R1, P.R2);
                    Er = Er + std::exp(-P.p * (r / P.r0 - 1.)) * f_cut(r, P.R1, P.R2);
                }
            }
            priv = priv + P.A0 * Er - P.xi * std::sqrt(Eb2);
        }
#pragma omp atomic
        energy = energy + priv;
    }
}
// :51-94  three sweeps: exp cache, band term Eb(i), gather
void RJL_forces(Particles& a, const NeighbourList& nl, const RJLParams& P) {
    const int m = nl.neighb_num_max;
    std::vector<double> Eb((size_t)nl.N), expp((size_t)m * nl.N), expq((size_t)m * nl.N);
#pragma omp parallel
    {
#pragma omp for
        for (int i = 0; i < nl.N; ++i)
            for (int p = 0; p < nl.nnum[i]; ++p) {
                size_t s = (size_t)i * m + p;
                double r = nl.moddr[s];
                if (r <= P.R2) {
                    expq[s] = std::exp(-2. * P.q * (r / P.r0 - 1.));
                    expp[s] = std::exp(-P.p * (r / P.r0 - 1.));
                }
            }
#pragma omp for
        for (int i = 0; i < nl.N; ++i) {
            Eb[i] = 0.;
            for (int p = 0; p < nl.nnum[i]; ++p) {
                size_t s = (size_t)i * m + p;
                double r = nl.moddr[s];
                if (r <= P.R2) Eb[i] = Eb[i] + expq[s] * f_cut(r, P.R1, P.R2);
            }
            Eb[i] = std::sqrt(Eb[i]);
        }
#pragma omp for
        for (int i = 0; i < nl.N; ++i)
            for (int p = 0; p < nl.nnum[i]; ++p) {
                size_t s = (size_t)i * m + p;
                double r = nl.moddr[s];
                if (r <= P.R2) {
                    double fc = f_cut(r, P.R1, P.R2), dfc = df_cut(r, P.R1, P.R2);
                    double c = (2. * P.A0 * (P.p / P.r0 * fc - dfc * r) * expp[s] -
                                P.xi * (P.q / P.r0 * fc - dfc / 2. * r) * (1. / Eb[i] + 1. / Eb[nl.nlist[s]]) * expq[s]) / r;
                    double* f = &a.forces[3 * nl.particle_index[i]];
                    for (int k = 0; k < 3; ++k) f[k] = f[k] - c * nl.dr[3 * s + k];
                }
            }
    }
}

// ---- TersoffBrenner.f90 -----------------------------------------------------------------------
void TB_finish_parameters(TBParams& P) { P.c02 = P.c0 * P.c0; P.d02 = P.d0 * P.d0; }  // :19-20

// bond order B(p,i), :36-50 and :83-96
static void tb_bond_orders(std::vector<double>& B, const NeighbourList& nl, const TBParams& T) {
    const int m = nl.neighb_num_max;
#pragma omp parallel for
    for (int i = 0; i < nl.N; ++i)
        for (int p = 0; p < nl.nnum[i]; ++p) {
            size_t sp = (size_t)i * m + p;
            B[sp] = 0.;
            if (nl.moddr[sp] < T.R2) {
                for (int q = 0; q < nl.nnum[i]; ++q) {
                    size_t sq = (size_t)i * m + q;
                    if (p != q && nl.moddr[sq] < T.R2) {
                        double c = dot3(&nl.dr[3 * sp], &nl.dr[3 * sq]) / (nl.moddr[sp] * nl.moddr[sq]);
                        B[sp] = B[sp] + f_cut(nl.moddr[sq], T.R1, T.R2) * (1. + T.c02 / T.d02 - T.c02 / (T.d02 + pw2(1. + c)));
                    }
                }
                B[sp] = std::pow(1. + T.a0 * B[sp], -T.delt);
            }
        }
}
// :26-72  pairs with j>i in group-local numbering, average of B_ij and B_ji
void TB_energy(double& energy, const NeighbourList& nl, const TBParams& T) {
    energy = 0.;
    const int m = nl.neighb_num_max;
    std::vector<double> B((size_t)m * nl.N, 0.);
    tb_bond_orders(B, nl, T);
#pragma omp parallel
    {
        double priv = 0.;
#pragma omp for
        for (int i = 0; i < nl.N; ++i)
            for (int p = 0; p < nl.nnum[i]; ++p) {
                size_t sp = (size_t)i * m + p;
                double r = nl.moddr[sp];
                if (r < T.R2) {
                    int j = nl.nlist[sp];
                    if (j > i) {
                        int q = 0;
                        for (; q < nl.nnum[j]; ++q)
                            if (nl.nlist[(size_t)j * m + q] == i) break;
                        double Bji = (q < nl.nnum[j]) ? B[(size_t)j * m + q] : 0.;
                        double aa = -std::sqrt(2. * T.s) * T.b * (r - T.r0);
                        priv = priv + f_cut(r, T.R1, T.R2) * T.d / (T.s - 1.) * (std::exp(aa) - (B[sp] + Bji) / 2 * T.s * std::exp(aa / T.s));
                    }
                }
            }
#pragma omp atomic
        energy = energy + priv;
    }
}
// :74-150
void TB_forces(Particles& a, const NeighbourList& nl, const TBParams& T) {
    const int m = nl.neighb_num_max;
    std::vector<double> B((size_t)m * nl.N, 0.);
    tb_bond_orders(B, nl, T);
    const double ex = 1. / T.delt + 1.;
    std::string err;
#pragma omp parallel for
    for (int i = 0; i < nl.N; ++i) {
        double* f = &a.forces[3 * nl.particle_index[i]];
        for (int p = 0; p < nl.nnum[i]; ++p) {
            size_t sp = (size_t)i * m + p;
            const double* dp = &nl.dr[3 * sp];
            double rp = nl.moddr[sp];
            double dB[3] = {0., 0., 0.};
            for (int q = 0; q < nl.nnum[i]; ++q) {  // :112-120
                if (p != q) {
                    size_t sq = (size_t)i * m + q;
                    const double* dq = &nl.dr[3 * sq];
                    double rq = nl.moddr[sq];
                    double rr = 1. / rp / rq;
                    double cosi = dot3(dp, dq) * rr;
                    double g1 = f_cut(rq, T.R1, T.R2) * 2. * T.a0 * T.c02 * (1. + cosi) / pw2(T.d02 + pw2(1. + cosi));
                    double g2 = df_cut(rq, T.R1, T.R2) * T.a0 * (1. + T.c02 / T.d02 - T.c02 / (T.d02 + pw2(1. + cosi)));
                    for (int k = 0; k < 3; ++k)
                        dB[k] = dB[k] + g1 * ((dp[k] + dq[k]) * rr - cosi * (dp[k] / (rp * rp) + dq[k] / (rq * rq))) + g2 * dq[k];
                }
            }
            double bp = std::pow(B[sp], ex);
            for (int k = 0; k < 3; ++k) dB[k] = dB[k] * bp;
            int j = nl.nlist[sp];
            int l = 0;
            for (; l < nl.nnum[j]; ++l)
                if (nl.nlist[(size_t)j * m + l] == i) break;
            if (l >= nl.nnum[j]) {
#pragma omp critical
                err = "tb: neighbour list is not symmetric";
                continue;
            }
            size_t sl = (size_t)j * m + l;
            const double* dl = &nl.dr[3 * sl];
            double rl = nl.moddr[sl];
            double bl = std::pow(B[sl], ex);
            for (int q = 0; q < nl.nnum[j]; ++q) {  // :126-133
                if (q != l) {
                    size_t sq = (size_t)j * m + q;
                    const double* dq = &nl.dr[3 * sq];
                    double rq = nl.moddr[sq];
                    double rr = 1. / rl / rq;
                    double cosi = dot3(dl, dq) * rr;
                    double g = bl * f_cut(rq, T.R1, T.R2) * 2. * T.a0 * T.c02 * (1. + cosi) / pw2(T.d02 + pw2(1. + cosi));
                    for (int k = 0; k < 3; ++k) dB[k] = dB[k] + g * (-dq[k] * rr + cosi * dl[k] / (rl * rl));
                }
            }
            for (int k = 0; k < 3; ++k) dB[k] = dB[k] * (-T.delt) / 2.;
            if (rp < T.R2) {  // :135-141
                double f_c = f_cut(rp, T.R1, T.R2);
                double dff = df_cut(rp, T.R1, T.R2) / f_c;
                double aa = -std::sqrt(2. * T.s) * T.b * (rp - T.r0);
                double ea = std::exp(aa), eas = std::exp(aa / T.s);
                for (int k = 0; k < 3; ++k) {
                    double t1 = (dp[k] * dff - std::sqrt(2. * T.s) * T.b / rp * dp[k]) * ea;
                    double t2 = (dB[k] + (B[sp] + B[sl]) / 2 * (dp[k] * dff - std::sqrt(2. / T.s) * T.b / rp * dp[k])) * T.s * eas;
                    f[k] = f[k] + f_c * T.d / (T.s - 1.) * (t1 - t2);
                }
            }
            for (int q = 0; q < nl.nnum[j]; ++q) {  // :142-152
                if (q != l) {
                    size_t sq = (size_t)j * m + q;
                    const double* dq = &nl.dr[3 * sq];
                    double rq = nl.moddr[sq];
                    double f_c = f_cut(rl, T.R1, T.R2), df_c = df_cut(rl, T.R1, T.R2);
                    double rr = 1. / rl / rq;
                    double cosi = dot3(dl, dq) * rr;
                    double pre = T.delt / 2 * std::pow(B[sq], ex);
                    double g1 = f_c * 2. * T.a0 * T.c02 * (1. + cosi) / pw2(T.d02 + pw2(1. + cosi));
                    double g2 = df_c * T.a0 * (1. + T.c02 / T.d02 - T.c02 / (T.d02 + pw2(1. + cosi)));
                    double tail = f_cut(rq, T.R1, T.R2) * T.d / (T.s - 1.) * T.s * std::exp(-std::sqrt(2. * T.s) * T.b * (rq - T.r0) / T.s);
                    for (int k = 0; k < 3; ++k)
                        f[k] = f[k] + pre * (g1 * (-dq[k] * rr + cosi * dl[k] / (rl * rl)) + dp[k] * g2) * tail;
                }
            }
        }
    }
    if (!err.empty()) throw StopError(err);
}

// ---- md_interactions.f90 ----------------------------------------------------------------------
// :65-118
int nl_n_for(const std::string& name) {
    if (name == "lj") return 2;
    if (name == "lj1g") return 1;
    if (name == "ljc") return 3;
    if (name == "morsec") return 3;
    if (name == "tb") return 1;
    if (name == "rjl") return 1;
    if (name == "rebosc") return 1;
    return -1;
}
// :123-132  (the caller has filled group_nums and nl[j].{neighb_num_max,r_cut,update_period})
void setup_interaction_lists(Interaction& it, const std::vector<ParticleGroup>& groups) {
    for (int j = 0; j < it.nl_n; ++j) {
        it.nl[j].N = groups[it.group_nums[2 * j] - 1].N;
        create_neighbour_list(it.nl[j]);
    }
}
// :180-193
void allocate_graphene_norm(std::vector<Interaction>& its) {
    for (auto& it : its) {
        if (it.interaction_name == "ljc") it.ljc.gr_norm.assign((size_t)3 * it.nl[2].N, 0.);
        if (it.interaction_name == "morsec") it.morsec.gr_norm.assign((size_t)3 * it.nl[2].N, 0.);
    }
}
// graphenenorm.f90:58-71
static void update_nearest_neighbours_in_graphene(int md_step, NeighbourList& nn, const NeighbourList& nl, const Particles& a,
                                                  const ParticleGroup& g, const SimulationCell& box) {
    if (md_step % nl.update_period == 0) find_gr_nearest_neighbors(nn, nl);
    else find_neighbour_distances(nn, a, g, g, box);
}
// :138-178 and :195-208
void update_interactions_neighbour_lists(int md_step, std::vector<Interaction>& its, const Particles& a,
                                         const std::vector<ParticleGroup>& groups, const SimulationCell& cell, double& ts, double& td) {
    for (size_t i = 0; i < its.size(); ++i) {
        Interaction& it = its[i];
        update_neighbour_list(md_step, it.nl[0], a, groups[it.group_nums[0] - 1], groups[it.group_nums[1] - 1], cell, ts, td);
        const std::string& nm = it.interaction_name;
        if (nm == "lj" || nm == "ljc" || nm == "morsec") converce_neighbour_list(it.nl[1], groups[it.group_nums[1] - 1], it.nl[0]);
        if (nm == "ljc" || nm == "morsec") {
            size_t j = 0;
            for (; j < its.size(); ++j)
                if (its[j].interaction_name == "tb" || its[j].interaction_name == "rebosc") break;
            if (j < its.size())
                update_nearest_neighbours_in_graphene(md_step, it.nl[2], its[j].nl[0], a, groups[its[j].group_nums[0] - 1], cell);
            else
                update_neighbour_list(md_step, it.nl[2], a, groups[it.group_nums[4] - 1], groups[it.group_nums[5] - 1], cell, ts, td);
        }
    }
    for (auto& it : its) {
        if (it.interaction_name == "ljc") find_norm_in_graphene(it.ljc.gr_norm, it.nl[2].dr, it.nl[2].N, it.nl[2].neighb_num_max);
        if (it.interaction_name == "morsec") find_norm_in_graphene(it.morsec.gr_norm, it.nl[2].dr, it.nl[2].N, it.nl[2].neighb_num_max);
    }
}
// :210-242  forces accumulate over the interactions in file order
void calculate_forces(Particles& a, std::vector<Interaction>& its) {
    for (auto& it : its) {
        if (it.numerical_force) continue;
        const std::string& nm = it.interaction_name;
        if (nm == "lj") { LJ_forces(a, it.nl[0], it.lj); LJ_forces(a, it.nl[1], it.lj); }
        else if (nm == "lj1g") LJ1g_forces(a, it.nl[0], it.lj1g);
        else if (nm == "ljc") { LJC_forces_for_graphene(a, it.nl[0], it.nl[2], it.ljc); LJC_forces_for_other_atoms(a, it.nl[1], it.ljc); }
        else if (nm == "morsec") { MorseC_forces_for_graphene(a, it.nl[0], it.nl[2], it.morsec); MorseC_forces_for_other_atoms(a, it.nl[1], it.morsec); }
        else if (nm == "tb") TB_forces(a, it.nl[0], it.tb);
        else if (nm == "rjl") RJL_forces(a, it.nl[0], it.rjl);
    }
}
// :244-271  always on nl(1)
void calculate_potential_energies(std::vector<Interaction>& its) {
    for (auto& it : its) {
        const std::string& nm = it.interaction_name;
        if (nm == "lj") LJ_energy(it.energy, it.nl[0], it.lj);
        else if (nm == "lj1g") LJ1g_energy(it.energy, it.nl[0], it.lj1g);
        else if (nm == "ljc") LJC_energy(it.energy, it.nl[0], it.ljc);
        else if (nm == "morsec") MorseC_energy(it.energy, it.nl[0], it.morsec);
        else if (nm == "tb") TB_energy(it.energy, it.nl[0], it.tb);
        else if (nm == "rjl") RJL_energy(it.energy, it.nl[0], it.rjl);
        else if (nm == "rebosc") REBOsc_energy(it.energy, it.nl[0], it.rebosc);
    }
}

}  // namespace oracle
