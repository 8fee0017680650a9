// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).  PARITY UNPINNED.
// Adapter that lets the shared md() driver (pfmds_b200/host/md_driver.hpp) and the ctypes test API
// drive the CPU restatement.  The driver only supplies text I/O and phase control; every number
// comes from oracle::System.
#pragma once
#include <cmath>
#include <string>
#include <vector>

#include "../pfmds_b200/host/md_inputs.hpp"
#include "oracle.hpp"

namespace oracle {

struct OracleEngine {
    System sys;
    std::vector<std::string> kind_names{"nve", "nvt", "nvms"};

    void create(int n, const double* pos, const double* vel, const double* mass, const double* box) {
        sys.atoms.N = n;
        sys.atoms.positions.assign(pos, pos + 3 * (size_t)n);
        sys.atoms.velocities.assign(vel, vel + 3 * (size_t)n);
        sys.atoms.forces.assign(3 * (size_t)n, 0.);
        sys.atoms.masses.assign(mass, mass + n);
        sys.atoms.atom_types.assign((size_t)n, "");
        for (int k = 0; k < 3; ++k) { sys.cell.box_size[k] = box[k]; sys.cell.half_box_size[k] = 0.5 * box[k]; }  // md_read_write.f90:32-35
    }
    void set_group(int g, const std::vector<int>& idx1) {
        if ((int)sys.groups.size() < g) sys.groups.resize((size_t)g);
        ParticleGroup& G = sys.groups[(size_t)g - 1];
        G.indexes.clear();
        for (int i : idx1) G.indexes.push_back(i - 1);
        G.N = (int)idx1.size();
    }
    void set_roles(int am, int xyz, int z, int all) { sys.all_moving = am; sys.xyz_moving = xyz; sys.z_moving = z; sys.all_atoms = all; }
    void add_nhc(const pfmds_host::NhcSpec& n) {  // md_simulation.f90:84-89
        NoseHooverChain c;
        create_nose_hoover_chain(c, n.M);
        set_nose_hoover_chain(c, n.temperature, n.q1, n.group, sys.groups.at((size_t)n.group - 1).N);
        sys.nhc.push_back(c);
    }
    void add_group_change(int from, int to, int ts1, int ts2, int frec) {  // md_simulation.f90:63-71
        if (from < 1 || to < 1 || from > (int)sys.groups.size() || to > (int)sys.groups.size()) throw StopError("error: group number out of range in a group change");
        sys.changes.push_back(System::GroupChange{from, to, ts1, ts2, frec});
    }
    int group_size(int g) const { return sys.groups.at((size_t)g - 1).N; }
    void set_misc(int zmp, bool inv) { sys.zero_momentum_period = zmp; sys.invert_z_vel = inv; }
    void add_interaction(const pfmds_host::InteractionSpec& sp) {  // md_interactions.f90:59-136
        Interaction it;
        it.interaction_name = sp.name;
        it.parameters_file = sp.parameters_file;
        it.nl_n = nl_n_for(sp.name);
        if (it.nl_n < 0 || it.nl_n != (int)sp.lists.size()) throw StopError("error: unknown interaction name");
        const auto& p = sp.params;
        if (sp.name == "lj") it.lj = LJParams{p.at(0), p.at(1), p.at(2), p.at(3)};
        else if (sp.name == "lj1g") { it.lj1g = LJ1gParams{p.at(0), p.at(1), p.at(2), p.at(3), 0, 0, 0, 0}; LJ1g_finish_parameters(it.lj1g); }
        else if (sp.name == "ljc") it.ljc = LJCParams{p.at(0), p.at(1), p.at(2), p.at(3), p.at(4), p.at(5) != 0., {}};
        else if (sp.name == "morsec") it.morsec = MorseCParams{p.at(0), p.at(1), p.at(2), p.at(3), p.at(4), p.at(5), p.at(6) != 0., {}};
        else if (sp.name == "tb") { it.tb = TBParams{p.at(0), p.at(1), p.at(2), p.at(3), p.at(4), p.at(5), p.at(6), p.at(7), p.at(8), p.at(9), 0, 0}; TB_finish_parameters(it.tb); }
        else if (sp.name == "rjl") it.rjl = RJLParams{p.at(0), p.at(1), p.at(2), p.at(3), p.at(4), p.at(5), p.at(6)};
        else if (sp.name == "rebosc") {  // REBOsolidcarbon.f90:12-25: A Q alpha / B(3) / beta(3) / T / g(6) / R1 R2
            it.rebosc = REBOscParams{p.at(0), p.at(1), p.at(2), {p.at(3), p.at(4), p.at(5)}, {p.at(6), p.at(7), p.at(8)}, p.at(9),
                                     {p.at(10), p.at(11), p.at(12), p.at(13), p.at(14), p.at(15)}, p.at(16), p.at(17)};
            it.numerical_force = true;  // md_interactions.f90:103
        }
        it.neib_order = (sp.name == "tb" || sp.name == "rjl") ? 2 : (sp.name == "rebosc" ? 3 : 0);
        it.nl.resize((size_t)it.nl_n);
        for (int j = 0; j < it.nl_n; ++j) {
            it.group_nums.push_back(sp.lists[j].g1);
            it.group_nums.push_back(sp.lists[j].g2);
            it.nl[j].neighb_num_max = sp.lists[j].neighb_num_max;
            it.nl[j].r_cut = sp.lists[j].r_cut;
            it.nl[j].update_period = sp.lists[j].update_period;
        }
        setup_interaction_lists(it, sys.groups);
        sys.interactions.push_back(std::move(it));
        allocate_graphene_norm(sys.interactions);
    }
    void advance(int kind, double dt, int first, int n, bool = false) {
        init_time_steps(sys.dt, dt);
        for (int t = first; t < first + n; ++t) sys.step(t, kind_names.at((size_t)kind));
    }
    // the reference's own sequence: one step, one energy evaluation (md_simulation.f90:138-199)
    void advance_logged(int kind, double dt, int first, int n, int log_period, std::vector<pfmds_host::EnergyRow>& rows) {
        rows.clear();
        for (int t = first; t < first + n; ++t) {
            advance(kind, dt, t, 1);
            if (t % log_period != 0) continue;
            pfmds_host::EnergyRow r;
            energies(r.e_inter, r.ke, r.temp, r.e_nhc);
            rows.push_back(r);
        }
    }
    void energies(std::vector<double>& e_inter, double& ke, double& temp, std::vector<double>& e_nhc) {  // md_simulation.f90:190-199
        double t0 = omp_wtime();
        calculate_potential_energies(sys.interactions);
        e_inter.resize(sys.interactions.size());
        for (size_t i = 0; i < sys.interactions.size(); ++i) e_inter[i] = sys.interactions[i].energy;
        calculate_temperature(temp, ke, sys.atoms, sys.groups[(size_t)sys.all_moving - 1]);
        e_nhc.resize(sys.nhc.size());
        for (size_t i = 0; i < sys.nhc.size(); ++i) { calculate_nose_hoover_chain_energy(sys.nhc[i]); e_nhc[i] = sys.nhc[i].e; }
        sys.t_energy += omp_wtime() - t0;
    }
    void diagnostics(double fs[3], double mc[3], double mcv[3], double& vmax, std::vector<int>& nl_load) {  // :212-223
        const ParticleGroup& all = sys.groups[(size_t)sys.all_atoms - 1];
        calculate_force_sum(fs, sys.atoms, all);
        calculate_mass_center(mc, sys.atoms, all);
        calculate_mass_center_velocity(mcv, sys.atoms, all);
        find_max_velocity(vmax, sys.atoms);
        nl_load.clear();
        for (auto& it : sys.interactions)
            for (auto& l : it.nl) {
                int m = 0;
                if (l.N > 0) for (int v : l.nnum) m = v > m ? v : m;
                nl_load.push_back(m);
            }
    }
    void download(double* pos, double* vel, double* frc) {
        size_t n = 3 * (size_t)sys.atoms.N;
        if (pos) std::copy(sys.atoms.positions.begin(), sys.atoms.positions.begin() + n, pos);
        if (vel) std::copy(sys.atoms.velocities.begin(), sys.atoms.velocities.begin() + n, vel);
        if (frc) std::copy(sys.atoms.forces.begin(), sys.atoms.forces.begin() + n, frc);
    }
    // exact restart: what positions and velocities do not carry (thermostat chains, group%N); same role as pfmds_save_state
    void save_state(std::vector<double>& blob) {
        blob.clear();
        blob.push_back((double)sys.nhc.size());
        for (auto& n : sys.nhc) {
            blob.push_back((double)n.M);
            blob.insert(blob.end(), n.x.begin(), n.x.end());
            blob.insert(blob.end(), n.v.begin(), n.v.end());
        }
        blob.push_back((double)sys.groups.size());
        for (auto& g : sys.groups) blob.push_back((double)g.N);
    }
    void restore_state(const double* pos, const double* vel, const std::vector<double>& blob) {
        size_t n3 = 3 * (size_t)sys.atoms.N, k = 0;
        std::copy(pos, pos + n3, sys.atoms.positions.begin());
        std::copy(vel, vel + n3, sys.atoms.velocities.begin());
        if ((size_t)blob.at(k++) != sys.nhc.size()) throw StopError("error: the checkpoint does not belong to this settings file");
        for (auto& n : sys.nhc) {
            if ((int)blob.at(k++) != n.M) throw StopError("error: the checkpoint does not belong to this settings file");
            for (int i = 0; i < n.M; ++i) n.x[(size_t)i] = blob.at(k++);
            for (int i = 0; i < n.M; ++i) n.v[(size_t)i] = blob.at(k++);
        }
        if ((size_t)blob.at(k++) != sys.groups.size()) throw StopError("error: the checkpoint does not belong to this settings file");
        for (auto& g : sys.groups) g.N = (int)blob.at(k++);
        // lists and forces of the checkpointed step (md_simulation.f90:160-165 without the momentum removal)
        double ts = 0, td = 0;
        update_interactions_neighbour_lists(0, sys.interactions, sys.atoms, sys.groups, sys.cell, ts, td);
        zero_forces(sys.atoms, sys.groups[(size_t)sys.all_atoms - 1]);
        calculate_forces(sys.atoms, sys.interactions);
        calculate_forces_numerically(sys.atoms, sys.interactions);
    }
    void timers(double t[6]) {
        t[0] = sys.t_pos_vel; t[1] = sys.t_nlists; t[2] = sys.t_nlsearch; t[3] = sys.t_nldistance; t[4] = sys.t_forces; t[5] = sys.t_energy;
    }
    static double omp_wtime();
};

}  // namespace oracle
