// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement (C++17 + OpenMP, FP64) of the PFMDS MD inner
// loop.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
// build, link, import or execute anything under oracle/.  The product (pfmds_b200/) never does.
//
// PARITY UNPINNED: the reference (Fortran 90) cannot be compiled in this image (no Fortran compiler,
// the shipped binaries are Win32) and it holds no tests, golden vectors or sample outputs.  This
// restatement follows the reference line by line (citations below are relative to
// /root/reference/code_source/) and is validated by physics invariants (force = -grad E by finite
// differences, sum F = 0, conserved energy) in tests/test_oracle.py, and anchored from outside the repository by
// tests/test_zz_anchors.py: closed forms of rjl / tb / lj1g on perfect lattices written in numpy from the formulas (1e-12) and the
// figures the parameter sets were fitted to in their source papers (Cleri-Rosato Cu 3.544 eV/atom at a = 3.615 A; Brenner set I
// graphite 7.3756 eV/atom at 1.42 A).
//
// Conventions: all reals are double (the reference builds with -fdefault-real-8); indices are
// 0-based here and 1-based in the reference; arrays dimensioned (3,N) in Fortran are [3*i+k] here;
// nlist(max,N) / moddr(max,N) are [i*max+p]; dr(3,max,N) is [(i*max+p)*3+k].
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

namespace oracle {

struct StopError : std::runtime_error { using std::runtime_error::runtime_error; };  // Fortran `stop`

// MOLECULAR_DYNAMICS/md_general.f90:5-41
struct TimeSteps { double ts[4]{0, 0, 0, 0}; double simulation_time = 0; };
struct SimulationCell { double box_size[3]{0, 0, 0}, half_box_size[3]{0, 0, 0}; };
struct Particles {
    int N = 0;
    std::vector<double> positions, velocities, forces, masses;
    std::vector<std::string> atom_types;
};
struct ParticleGroup { int N = 0; std::vector<int> indexes; };
struct NoseHooverChain { int M = 0, L = 0, group_num = 0; std::vector<double> x, v, q; double temperature = 0, s = 1, e = 0; };
struct NeighbourList {
    int N = 0, neighb_num_max = 0, update_period = 1;
    double r_cut = 0;
    std::vector<int> nlist, nnum, lessnnum, particle_index;
    std::vector<double> dr, moddr;
};
struct IntegratorParams { int l = 0, period_snapshot = 1, period_log = 1; double dt = 0; std::string int_name = "none"; };

// INTERACTION_POTENTIALS/*.f90 parameter types
struct LJParams { double eps, sig, R1, R2; };
struct LJ1gParams { double eps, sig, R1, R2, c6, c12, c6t6, c12t12; };
struct LJCParams { double eps, sig, delt, R1, R2; bool simplified; std::vector<double> gr_norm; };
struct MorseCParams { double d, r, a, delt, R1, R2; bool simplified; std::vector<double> gr_norm; };
struct RJLParams { double A0, xi, p, q, r0, R1, R2; };
struct TBParams { double d, s, b, r0, delt, a0, c0, d0, R1, R2, c02, d02; };
struct REBOscParams { double A, Q, alpha, B[3], beta[3], T, g[6], R1, R2; };  // REBOsolidcarbon.f90:6-8

// MOLECULAR_DYNAMICS/md_interactions.f90:15-34
struct Interaction {
    int nl_n = 0, neib_order = 0;
    std::vector<int> group_nums;  // 1-based group numbers as in the settings file
    std::vector<NeighbourList> nl;
    double energy = 0;
    std::string interaction_name, parameters_file;
    LJParams lj{}; LJ1gParams lj1g{}; LJCParams ljc{}; MorseCParams morsec{}; RJLParams rjl{}; TBParams tb{}; REBOscParams rebosc{};
    bool numerical_force = false;
};

// md_general.f90
void init_time_steps(TimeSteps& dt, double delta_t);
void create_particle_group(ParticleGroup& g, const std::vector<std::string>& type_names, const Particles& atoms);
void change_particle_group_N(ParticleGroup& group, int md_step, int change_ts1, int change_ts2, int change_frec, const ParticleGroup& init_group);
void scale_velocities(Particles& a, const ParticleGroup& g, double s);
void calculate_kinetic_energy(double& ke, const Particles& a, const ParticleGroup& g);
void calculate_mass_center(double mc[3], const Particles& a, const ParticleGroup& g);
void calculate_mass_center_velocity(double mcv[3], const Particles& a, const ParticleGroup& g);
void zero_momentum(Particles& a, const ParticleGroup& g);
void calculate_masses_sum(double& totm, const Particles& a, const ParticleGroup& g);
void calculate_force_sum(double fs[3], const Particles& a, const ParticleGroup& g);
void calculate_temperature(double& temp, double& ke, const Particles& a, const ParticleGroup& g);
void check_positions(const Particles& a, const SimulationCell& box);
void find_max_velocity(double& v, const Particles& a);
void invert_z_velocities(Particles& a, double z_low_border, double z_high_border);
void find_distance(double dr[3], double& dr2, const double* vec1, const double* vec2, const SimulationCell& box);

// md_neighbours.f90
void create_neighbour_list(NeighbourList& nl);
void update_neighbour_list(int md_step, NeighbourList& nl, const Particles& a, const ParticleGroup& g1, const ParticleGroup& g2,
                           const SimulationCell& box, double& t_search, double& t_distance);
void find_neighbours(NeighbourList& nl, const Particles& a, const ParticleGroup& g1, const ParticleGroup& g2, const SimulationCell& box);
void find_neighbour_distances(NeighbourList& nl, const Particles& a, const ParticleGroup& g1, const ParticleGroup& g2, const SimulationCell& box);
void converce_neighbour_list(NeighbourList& cnl, const ParticleGroup& g2, const NeighbourList& nl);

// md_integrators.f90
void integrate_verlet_xyz_positions(Particles& a, const ParticleGroup& g, const TimeSteps& s, const SimulationCell& box);
void integrate_verlet_z_positions(Particles& a, const ParticleGroup& g, const TimeSteps& s, const SimulationCell& box);
void integrate_verlet_xyz_velocities(Particles& a, const ParticleGroup& g, const TimeSteps& s);
void integrate_verlet_z_velocities(Particles& a, const ParticleGroup& g, const TimeSteps& s);
void molecular_static_xyz_velocities(Particles& a, const ParticleGroup& g);
void molecular_static_1D_velocities(Particles& a, const ParticleGroup& g);
void zero_forces(Particles& a, const ParticleGroup& g);
void create_nose_hoover_chain(NoseHooverChain& nhc, int M);
void set_nose_hoover_chain(NoseHooverChain& nhc, double temp, double q1, int gn, int l);
void integrate_nose_hoover_chain(NoseHooverChain& nhc, Particles& a, const ParticleGroup& g, const TimeSteps& dt);
void calculate_nose_hoover_chain_energy(NoseHooverChain& nhc);

// INTERACTION_POTENTIALS
double f_cut(double r, double R1, double R2);
double df_cut(double r, double R1, double R2);
double f_cut_poly(double r, double R1, double R2);
void f_dfr_cut(double& f, double& dfr, double r, double R1, double R2);
void LJ_energy(double& e, const NeighbourList& nl, const LJParams& p);
void LJ_forces(Particles& a, const NeighbourList& nl, const LJParams& p);
void LJ1g_finish_parameters(LJ1gParams& p);
void LJ1g_energy(double& e, const NeighbourList& nl, const LJ1gParams& p);
void LJ1g_forces(Particles& a, const NeighbourList& nl, const LJ1gParams& p);
void find_gr_nearest_neighbors(NeighbourList& nl_nn, const NeighbourList& nl);
void find_norm_in_graphene(std::vector<double>& gr_norm, const std::vector<double>& dr_nn, int n_rows, int maxn);
void LJC_energy(double& e, const NeighbourList& nl, const LJCParams& p);
void LJC_forces_for_graphene(Particles& a, const NeighbourList& nl, const NeighbourList& nl_nn, LJCParams& p);
void LJC_forces_for_other_atoms(Particles& a, const NeighbourList& nl, const LJCParams& p);
void MorseC_energy(double& e, const NeighbourList& nl, const MorseCParams& p);
void MorseC_forces_for_graphene(Particles& a, const NeighbourList& nl, const NeighbourList& nl_nn, MorseCParams& p);
void MorseC_forces_for_other_atoms(Particles& a, const NeighbourList& nl, const MorseCParams& p);
void RJL_energy(double& e, const NeighbourList& nl, const RJLParams& p);
void RJL_forces(Particles& a, const NeighbourList& nl, const RJLParams& p);
void TB_finish_parameters(TBParams& p);
void TB_energy(double& e, const NeighbourList& nl, const TBParams& p);
void TB_forces(Particles& a, const NeighbourList& nl, const TBParams& p);
void REBOsc_energy(double& e, const NeighbourList& nl, const REBOscParams& p);

// md_interactions.f90
int nl_n_for(const std::string& name);  // lj 2, lj1g 1, ljc 3, morsec 3, tb 1, rebosc 1, rjl 1; -1 unknown
void setup_interaction_lists(Interaction& it, const std::vector<ParticleGroup>& groups);
void allocate_graphene_norm(std::vector<Interaction>& its);
void update_interactions_neighbour_lists(int md_step, std::vector<Interaction>& its, const Particles& a,
                                         const std::vector<ParticleGroup>& groups, const SimulationCell& cell,
                                         double& t_search, double& t_distance);
void calculate_forces(Particles& a, std::vector<Interaction>& its);
void calculate_forces_numerically(Particles& a, std::vector<Interaction>& its);
void calculate_potential_energies(std::vector<Interaction>& its);

// The whole simulation state + one md step (body of the loop in md_simulation.f90:114-243 without I/O).
struct System {
    SimulationCell cell;
    TimeSteps dt;
    Particles atoms;
    std::vector<ParticleGroup> groups;  // groups[g-1] is group number g
    std::vector<Interaction> interactions;
    std::vector<NoseHooverChain> nhc;
    int all_moving = 1, xyz_moving = 1, z_moving = 1, all_atoms = 1;
    int zero_momentum_period = 1;
    bool invert_z_vel = false;
    // md_simulation.f90:63-71: `change_group_num` entries, applied at the top of every step (:116-119)
    struct GroupChange { int from, to, ts1, ts2, frec; };
    std::vector<GroupChange> changes;
    double t_pos_vel = 0, t_nlists = 0, t_nlsearch = 0, t_nldistance = 0, t_forces = 0, t_energy = 0;
    // md_simulation.f90:138-186 for one md_step; integrator_name in {"nve","nvt","nvms"}.
    void step(int md_step, const std::string& integrator_name);
};

// md_simulation.f90:19-274 with all file and stdout I/O (oracle_md.cpp).
int md(std::FILE* out, std::FILE* all_out, const std::string& input_path, const std::string& settings_filename,
       const std::string& output_prefix, int out_period, int num_of_omp_threads, int rand_seed);

}  // namespace oracle
