// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).  PARITY UNPINNED.
// Restates MOLECULAR_DYNAMICS/{md_general,md_neighbours,md_integrators}.f90 and the step body of
// md_simulation.f90.  Build with -ffp-contract=off so that dr2 = dx*dx+dy*dy+dz*dz is not fused.
#include "oracle.hpp"

#include <omp.h>

#include <cmath>
#include <cstdio>
#include <sstream>

namespace oracle {

// literal constants of the reference (FP64 because of -fdefault-real-8)
static const double mass_coef = 1.6605389217 / 1.6021765654 * 100.0;     // md_general.f90:165, md_integrators.f90:62
static const double kt_a_degree = 1.3806488 / 1.6021765654 * 1.0e-4;     // md_general.f90:304, md_integrators.f90:211

// ---- md_general.f90 -------------------------------------------------------------------------
// :45-55
void init_time_steps(TimeSteps& dt, double delta_t) {
    int pw = 1;
    for (int i = 0; i < 4; ++i) { dt.ts[i] = delta_t / pw; pw *= 2; }
}

// :57-80  indexes ordered by type name (column order of the group table) first, file order second
void create_particle_group(ParticleGroup& g, const std::vector<std::string>& type_names, const Particles& atoms) {
    g.N = 0;
    g.indexes.clear();
    for (const auto& nm : type_names)
        for (int i = 0; i < atoms.N; ++i)
            if (atoms.atom_types[i] == nm) g.indexes.push_back(i);
    g.N = (int)g.indexes.size();
}

// :82-94  deposition: the target group exposes the first N of its indexes; N follows the source group's N up to
// change_ts1, then grows by one atom at change_ts1 and every change_frec steps until change_ts2
void change_particle_group_N(ParticleGroup& group, int md_step, int change_ts1, int change_ts2, int change_frec, const ParticleGroup& init_group) {
    if (md_step <= change_ts1) {
        group.N = init_group.N;
        if (md_step == change_ts1) group.N = group.N + 1;
    } else {
        if (md_step < change_ts2) {
            if (change_frec == 0) throw StopError("error: change_frec is zero (integer division by zero in mod)");
            if ((md_step - change_ts1) % change_frec == 0) group.N = group.N + 1;
        }
    }
    if (group.N > (int)group.indexes.size()) group.N = (int)group.indexes.size();
}

// :96-112
void scale_velocities(Particles& a, const ParticleGroup& g, double s) {
#pragma omp parallel for
    for (int ind = 0; ind < g.N; ++ind) {
        int i = g.indexes[ind];
        for (int k = 0; k < 3; ++k) a.velocities[3 * i + k] = a.velocities[3 * i + k] * s;
    }
}

// :161-182  thread partial sums merged by an atomic add
void calculate_kinetic_energy(double& ke, const Particles& a, const ParticleGroup& g) {
    ke = 0.;
#pragma omp parallel
    {
        double ke_priv = 0.;
#pragma omp for
        for (int ind = 0; ind < g.N; ++ind) {
            int i = g.indexes[ind];
            const double* v = &a.velocities[3 * i];
            ke_priv = ke_priv + a.masses[i] * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) / 2 * mass_coef;
        }
#pragma omp atomic
        ke = ke + ke_priv;
    }
}

// :255-275
void calculate_masses_sum(double& totm, const Particles& a, const ParticleGroup& g) {
    totm = 0.;
#pragma omp parallel
    {
        double priv = 0.;
#pragma omp for
        for (int ind = 0; ind < g.N; ++ind) priv = priv + a.masses[g.indexes[ind]];
#pragma omp atomic
        totm = totm + priv;
    }
}

static void mass_weighted_sum(double out[3], const std::vector<double>& arr, const Particles& a, const ParticleGroup& g) {
    out[0] = out[1] = out[2] = 0.;
#pragma omp parallel
    {
        double priv[3] = {0., 0., 0.};
#pragma omp for
        for (int ind = 0; ind < g.N; ++ind) {
            int i = g.indexes[ind];
            for (int k = 0; k < 3; ++k) priv[k] = priv[k] + a.masses[i] * arr[3 * i + k];
        }
        for (int k = 0; k < 3; ++k) {
#pragma omp atomic
            out[k] = out[k] + priv[k];
        }
    }
    double totm;
    calculate_masses_sum(totm, a, g);
    for (int k = 0; k < 3; ++k) out[k] = out[k] / totm;
}
// :184-208
void calculate_mass_center(double mc[3], const Particles& a, const ParticleGroup& g) { mass_weighted_sum(mc, a.positions, a, g); }
// :210-234
void calculate_mass_center_velocity(double mcv[3], const Particles& a, const ParticleGroup& g) { mass_weighted_sum(mcv, a.velocities, a, g); }

// :236-253
void zero_momentum(Particles& a, const ParticleGroup& g) {
    double mcv[3];
    calculate_mass_center_velocity(mcv, a, g);
#pragma omp parallel for
    for (int ind = 0; ind < g.N; ++ind) {
        int i = g.indexes[ind];
        for (int k = 0; k < 3; ++k) a.velocities[3 * i + k] = a.velocities[3 * i + k] - mcv[k];
    }
}

// :277-299
void calculate_force_sum(double fs[3], const Particles& a, const ParticleGroup& g) {
    fs[0] = fs[1] = fs[2] = 0.;
#pragma omp parallel
    {
        double priv[3] = {0., 0., 0.};
#pragma omp for
        for (int ind = 0; ind < g.N; ++ind) {
            int i = g.indexes[ind];
            for (int k = 0; k < 3; ++k) priv[k] = priv[k] + a.forces[3 * i + k];
        }
        for (int k = 0; k < 3; ++k) {
#pragma omp atomic
            fs[k] = fs[k] + priv[k];
        }
    }
}

// :301-311   3N degrees of freedom, no constraint correction
void calculate_temperature(double& temp, double& ke, const Particles& a, const ParticleGroup& g) {
    calculate_kinetic_energy(ke, a, g);
    temp = 2 * ke / kt_a_degree / (3 * (g.N));
}

// :342-364
void check_positions(const Particles& a, const SimulationCell& box) {
    const double tolerance = 0.0000001;
    int bad = 0;
    std::string msg;
#pragma omp parallel for reduction(+ : bad)
    for (int i = 0; i < a.N; ++i)
        for (int k = 0; k < 3; ++k) {
            double x = a.positions[3 * i + k];
            if (!(x > (0. - tolerance) && x < (box.box_size[k] + tolerance))) {
#pragma omp critical
                {
                    std::ostringstream os;
                    os << " " << (i + 1) << "  particle out of cell  " << a.positions[3 * i] << " " << a.positions[3 * i + 1] << " "
                       << a.positions[3 * i + 2] << "\n";
                    msg += os.str();
                }
                bad = bad + 1;
            }
        }
    if (bad > 0) throw StopError(msg);
}

// :366-380
void find_max_velocity(double& v, const Particles& a) {
    double maxvel2 = -1.;
#pragma omp parallel for reduction(max : maxvel2)
    for (int i = 0; i < a.N; ++i) {
        const double* w = &a.velocities[3 * i];
        double v2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
        if (maxvel2 < v2) maxvel2 = v2;
    }
    v = std::sqrt(maxvel2);
}

// :382-398  elastic wall inside the slab [z_low, z_high]
void invert_z_velocities(Particles& a, double zl, double zh) {
#pragma omp parallel for
    for (int i = 0; i < a.N; ++i) {
        double z = a.positions[3 * i + 2], vz = a.velocities[3 * i + 2];
        if ((z > zl && z < (zl + zh) / 2 && vz > 0.) || (z < zh && z > (zl + zh) / 2 && vz < 0.)) a.velocities[3 * i + 2] = -vz;
    }
}

// :423-441  minimum image through Fortran sign(1.,x) (+1 for x>=+0, -1 otherwise)
static inline double fsign1(double x) { return std::copysign(1.0, x); }
void find_distance(double dr[3], double& dr2, const double* vec1, const double* vec2, const SimulationCell& box) {
    dr[0] = vec2[0] - vec1[0];
    dr[1] = vec2[1] - vec1[1];
    dr[2] = vec2[2] - vec1[2];
    for (int k = 0; k < 3; ++k) {
        double h = box.half_box_size[k];
        dr[k] = dr[k] - h * (fsign1(dr[k] - h) + fsign1(dr[k] + h));
    }
    dr2 = dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2];
}

// ---- md_neighbours.f90 ----------------------------------------------------------------------
// :9-28
void create_neighbour_list(NeighbourList& nl) {
    if (nl.N > 0) {
        size_t n = (size_t)nl.N, m = (size_t)nl.neighb_num_max;
        nl.particle_index.assign(n, 0);
        nl.nnum.assign(n, 0);
        nl.lessnnum.assign(n, 0);
        nl.nlist.assign(m * n, 0);
        nl.dr.assign(3 * m * n, 0.);
        nl.moddr.assign(m * n, 0.);
    }
}

// :32-52
void update_neighbour_list(int md_step, NeighbourList& nl, const Particles& a, const ParticleGroup& g1, const ParticleGroup& g2,
                           const SimulationCell& box, double& t_search, double& t_distance) {
    if (g1.N > 0 && g2.N > 0 && nl.N > 0) {
        double t = omp_get_wtime();
        if (md_step % nl.update_period == 0) {
            find_neighbours(nl, a, g1, g2, box);
            t_search += omp_get_wtime() - t;
        } else {
            find_neighbour_distances(nl, a, g1, g2, box);
            t_distance += omp_get_wtime() - t;
        }
    }
}

// :56-100  brute force O(N1*N2); entries in ascending group-2 local index; lessnnum = number of
// entries before the first one whose GLOBAL index exceeds the owner's
void find_neighbours(NeighbourList& nl, const Particles& a, const ParticleGroup& g1, const ParticleGroup& g2, const SimulationCell& box) {
    if (g1.N > nl.N) throw StopError("error: group1%N>nl%N");
    const int maxn = nl.neighb_num_max;
    std::string err;
#pragma omp parallel for
    for (int ind = 0; ind < g1.N; ++ind) {
        int i = g1.indexes[ind];
        nl.particle_index[ind] = g1.indexes[ind];
        int nnumind = 0, lessnnumind = -1;
        double dr[3], dr2;
        for (int jnd = 0; jnd < g2.N; ++jnd) {
            int j = g2.indexes[jnd];
            if (i != j) {
                find_distance(dr, dr2, &a.positions[3 * i], &a.positions[3 * j], box);
                if (dr2 < nl.r_cut * nl.r_cut) {
                    if (lessnnumind == -1 && i < j) lessnnumind = nnumind;
                    nnumind = nnumind + 1;
                    if (nnumind > maxn) {
#pragma omp critical
                        if (err.empty()) err = "error: too many neighbours " + std::to_string(ind + 1) + " " + std::to_string(nnumind);
                        nnumind = maxn;  // keep memory safe until the stop below
                        break;
                    }
                    size_t s = (size_t)ind * maxn + (nnumind - 1);
                    nl.nlist[s] = jnd;
                    nl.dr[3 * s] = dr[0];
                    nl.dr[3 * s + 1] = dr[1];
                    nl.dr[3 * s + 2] = dr[2];
                    nl.moddr[s] = std::sqrt(dr2);
                }
            }
        }
        nl.lessnnum[ind] = (lessnnumind == -1) ? nnumind : lessnnumind;
        nl.nnum[ind] = nnumind;
    }
    if (!err.empty()) throw StopError(err);
}

// :104-124
void find_neighbour_distances(NeighbourList& nl, const Particles& a, const ParticleGroup& g1, const ParticleGroup& g2, const SimulationCell& box) {
    const int maxn = nl.neighb_num_max;
#pragma omp parallel for
    for (int i = 0; i < nl.N; ++i) {
        double dr2;
        for (int p = 0; p < nl.nnum[i]; ++p) {
            size_t s = (size_t)i * maxn + p;
            find_distance(&nl.dr[3 * s], dr2, &a.positions[3 * g1.indexes[i]], &a.positions[3 * g2.indexes[nl.nlist[s]]], box);
            nl.moddr[s] = std::sqrt(dr2);
        }
    }
}

// :128-160  serial transpose; rows of the converse list come out in ascending i
void converce_neighbour_list(NeighbourList& cnl, const ParticleGroup& g2, const NeighbourList& nl) {
    if (g2.N > 0 && nl.N > 0 && cnl.N > 0) {
        if (g2.N > cnl.N) throw StopError("error: group2%N>cnl%N");
        for (int i = 0; i < g2.N; ++i) cnl.particle_index[i] = g2.indexes[i];
        for (int i = 0; i < cnl.N; ++i) cnl.nnum[i] = 0;
        const int maxn = nl.neighb_num_max, cmax = cnl.neighb_num_max;
        for (int i = 0; i < nl.N; ++i) {
            for (int p = 0; p < nl.nnum[i]; ++p) {
                size_t s = (size_t)i * maxn + p;
                int j = nl.nlist[s];
                cnl.nnum[j] = cnl.nnum[j] + 1;
                if (cnl.nnum[j] > cmax)
                    throw StopError("error: too many neighbours " + std::to_string(j + 1) + " " + std::to_string(cnl.nnum[j]));
                size_t c = (size_t)j * cmax + (cnl.nnum[j] - 1);
                cnl.nlist[c] = i;
                for (int k = 0; k < 3; ++k) cnl.dr[3 * c + k] = -nl.dr[3 * s + k];
                cnl.moddr[c] = nl.moddr[s];
            }
        }
    }
}

// ---- md_integrators.f90 ---------------------------------------------------------------------
// :7-31
void integrate_verlet_xyz_positions(Particles& a, const ParticleGroup& g, const TimeSteps& s, const SimulationCell& box) {
#pragma omp parallel for
    for (int ind = 0; ind < g.N; ++ind) {
        int i = g.indexes[ind];
        for (int k = 0; k < 3; ++k) {
            double& x = a.positions[3 * i + k];
            x = x + a.velocities[3 * i + k] * s.ts[0];
            if (x > box.box_size[k]) x = x - box.box_size[k];
            else if (x < 0.) x = x + box.box_size[k];
        }
    }
}
// :33-56
void integrate_verlet_z_positions(Particles& a, const ParticleGroup& g, const TimeSteps& s, const SimulationCell& box) {
#pragma omp parallel for
    for (int ind = 0; ind < g.N; ++ind) {
        int i = g.indexes[ind];
        const int k = 2;
        double& x = a.positions[3 * i + k];
        x = x + a.velocities[3 * i + k] * s.ts[0];
        if (x > box.box_size[k]) x = x - box.box_size[k];
        else if (x < 0.) x = x + box.box_size[k];
    }
}
// :58-77
void integrate_verlet_xyz_velocities(Particles& a, const ParticleGroup& g, const TimeSteps& s) {
#pragma omp parallel for
    for (int ind = 0; ind < g.N; ++ind) {
        int i = g.indexes[ind];
        for (int k = 0; k < 3; ++k)
            a.velocities[3 * i + k] = a.velocities[3 * i + k] + a.forces[3 * i + k] / a.masses[i] / mass_coef * s.ts[1];
    }
}
// :79-97
void integrate_verlet_z_velocities(Particles& a, const ParticleGroup& g, const TimeSteps& s) {
#pragma omp parallel for
    for (int ind = 0; ind < g.N; ++ind) {
        int i = g.indexes[ind];
        const int k = 2;
        a.velocities[3 * i + k] = a.velocities[3 * i + k] + a.forces[3 * i + k] / a.masses[i] / mass_coef * s.ts[1];
    }
}
// :99-123
void molecular_static_xyz_velocities(Particles& a, const ParticleGroup& g) {
#pragma omp parallel for
    for (int ind = 0; ind < g.N; ++ind) {
        int i = g.indexes[ind];
        const double* f = &a.forces[3 * i];
        double* v = &a.velocities[3 * i];
        double fv = f[0] * v[0] + f[1] * v[1] + f[2] * v[2];
        double ff = f[0] * f[0] + f[1] * f[1] + f[2] * f[2];
        if (fv > 0. && ff > 1.0e-12) {
            for (int k = 0; k < 3; ++k) v[k] = fv / ff * f[k];
        } else {
            v[0] = v[1] = v[2] = 0.;
        }
    }
}
// :125-145
void molecular_static_1D_velocities(Particles& a, const ParticleGroup& g) {
#pragma omp parallel for
    for (int ind = 0; ind < g.N; ++ind) {
        int i = g.indexes[ind];
        const double* f = &a.forces[3 * i];
        double* v = &a.velocities[3 * i];
        double fv = f[0] * v[0] + f[1] * v[1] + f[2] * v[2];
        if (fv > 0.) {
        } else {
            v[0] = v[1] = v[2] = 0.;
        }
    }
}
// :147-163
void zero_forces(Particles& a, const ParticleGroup& g) {
#pragma omp parallel for
    for (int ind = 0; ind < g.N; ++ind) {
        int i = g.indexes[ind];
        a.forces[3 * i] = a.forces[3 * i + 1] = a.forces[3 * i + 2] = 0.;
    }
}
// :165-180
void create_nose_hoover_chain(NoseHooverChain& nhc, int M) {
    nhc.x.assign((size_t)M, 0.);
    nhc.v.assign((size_t)M, 0.);
    nhc.q.assign((size_t)M, 0.);
    nhc.M = M;
    nhc.e = 0.;
    nhc.s = 1.;
}
// :182-198
void set_nose_hoover_chain(NoseHooverChain& nhc, double temp, double q1, int gn, int l) {
    if (l < 1 || q1 < 0. || temp < 0.) throw StopError("error: wrong nhc parameters");
    nhc.group_num = gn;
    nhc.L = l;
    nhc.temperature = temp;
    nhc.q[0] = q1;
    for (int i = 1; i < nhc.M; ++i) nhc.q[i] = nhc.q[0] / (3. * nhc.L);
}
// :200-245   (Fortran v(1..M) -> v[0..M-1]; ts(2)=dt/2, ts(3)=dt/4, ts(4)=dt/8)
void integrate_nose_hoover_chain(NoseHooverChain& n, Particles& a, const ParticleGroup& g, const TimeSteps& dt) {
    double ke, b = 0.;
    calculate_kinetic_energy(ke, a, g);
    const int M = n.M;
    double kt = 1.3806488 / 1.6021765654 * 1.0e-4 * n.temperature;
    double kedif = 2. * ke - 3. * n.L * kt;
    auto& v = n.v;
    auto& q = n.q;
    if (M == 1) {
        v[0] = v[0] + kedif / q[0] * dt.ts[2];
    } else {
        v[M - 1] = v[M - 1] + (q[M - 2] * v[M - 2] * v[M - 2] - kt) / q[M - 1] * dt.ts[2];
        for (int i = M - 2; i >= 1; --i) {  // Fortran i = M-1 .. 2
            b = std::exp(-v[i + 1] * dt.ts[3]);
            v[i] = v[i] * (b * b) + (q[i - 1] * v[i - 1] * v[i - 1] - kt) / q[i] * dt.ts[2] * b;
        }
        b = std::exp(-v[1] * dt.ts[3]);
        v[0] = v[0] * (b * b) + kedif / q[0] * dt.ts[2] * b;
    }
    n.s = std::exp(-v[0] * dt.ts[1]);
    scale_velocities(a, g, n.s);
    kedif = 2. * ke * (n.s * n.s) - 3. * n.L * kt;
    for (int i = 0; i < M; ++i) n.x[i] = n.x[i] + v[i] * dt.ts[1];
    if (M == 1) {
        v[0] = v[0] + kedif / q[0] * dt.ts[2];
    } else {
        v[0] = v[0] * (b * b) + kedif / q[0] * dt.ts[2] * b;  // same b as the last one above (:236)
        for (int i = 1; i <= M - 2; ++i) {                     // Fortran i = 2 .. M-1
            b = std::exp(-v[i + 1] * dt.ts[3]);
            v[i] = v[i] * (b * b) + (q[i - 1] * v[i - 1] * v[i - 1] - kt) / q[i] * dt.ts[2] * b;
        }
        v[M - 1] = v[M - 1] + (q[M - 2] * v[M - 2] * v[M - 2] - kt) / q[M - 1] * dt.ts[2];
    }
}
// :247-260
void calculate_nose_hoover_chain_energy(NoseHooverChain& n) {
    double kt = 1.3806488 / 1.6021765654 * 1.0e-4 * n.temperature;
    n.e = n.q[0] / 2 * (n.v[0] * n.v[0]) + 3. * n.L * kt * n.x[0];
    for (int i = 1; i < n.M; ++i) n.e = n.e + n.q[i] / 2 * (n.v[i] * n.v[i]) + kt * n.x[i];
}

// ---- md_simulation.f90:138-186, one md_step without the I/O ------------------------------------
void System::step(int md_step, const std::string& integrator_name) {
    for (const auto& ch : changes)  // :116-119
        change_particle_group_N(groups.at((size_t)ch.to - 1), md_step, ch.ts1, ch.ts2, ch.frec, groups.at((size_t)ch.from - 1));
    double t = omp_get_wtime();
    check_positions(atoms, cell);
    if (invert_z_vel) invert_z_velocities(atoms, 0.8 * cell.box_size[2], 0.9 * cell.box_size[2]);
    if (md_step != 0) {
        if (integrator_name == "nvt")
            for (auto& th : nhc) integrate_nose_hoover_chain(th, atoms, groups[th.group_num - 1], dt);
        integrate_verlet_xyz_velocities(atoms, groups[xyz_moving - 1], dt);
        integrate_verlet_z_velocities(atoms, groups[z_moving - 1], dt);
        integrate_verlet_xyz_positions(atoms, groups[xyz_moving - 1], dt, cell);
        integrate_verlet_z_positions(atoms, groups[z_moving - 1], dt, cell);
    }
    t_pos_vel += omp_get_wtime() - t;

    t = omp_get_wtime();
    update_interactions_neighbour_lists(md_step, interactions, atoms, groups, cell, t_nlsearch, t_nldistance);
    t_nlists += omp_get_wtime() - t;

    t = omp_get_wtime();
    if (md_step % zero_momentum_period == 0) zero_momentum(atoms, groups[all_atoms - 1]);
    zero_forces(atoms, groups[all_atoms - 1]);
    calculate_forces(atoms, interactions);
    calculate_forces_numerically(atoms, interactions);  // :165
    t_forces += omp_get_wtime() - t;

    t = omp_get_wtime();
    if (md_step != 0) {
        integrate_verlet_xyz_velocities(atoms, groups[xyz_moving - 1], dt);
        integrate_verlet_z_velocities(atoms, groups[z_moving - 1], dt);
        if (integrator_name == "nvt")
            for (auto& th : nhc) integrate_nose_hoover_chain(th, atoms, groups[th.group_num - 1], dt);
        if (integrator_name == "nvms") {
            molecular_static_xyz_velocities(atoms, groups[xyz_moving - 1]);
            molecular_static_1D_velocities(atoms, groups[z_moving - 1]);
        }
        dt.simulation_time = dt.simulation_time + dt.ts[0];
    }
    t_pos_vel += omp_get_wtime() - t;
}

}  // namespace oracle
