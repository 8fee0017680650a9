// Input side of the run_md_simulation drop-in: the fixed-order settings file, the extended-xyz
// file and the per-potential parameter files, read with the reference's list-directed grammar.
// Mirrors code_source/MOLECULAR_DYNAMICS/md_simulation.f90:48-93, md_interactions.f90:38-136,
// md_read_write.f90:10-61 and the read_*_parameters routines in INTERACTION_POTENTIALS/*.f90.
#pragma once
#include <string>
#include <vector>

#include "fortran_io.hpp"

namespace pfmds_host {

struct IntegratorParams { std::string int_name = "none"; double dt = 0; int l = 0, period_snapshot = 1, period_log = 1; };
struct NhcSpec { int group = 1; double temperature = 0; int M = 1; double q1 = 1; };
struct EnergyRow { std::vector<double> e_inter; double ke = 0, temp = 0; std::vector<double> e_nhc; };  // what energies() returns, per logged step
struct ListSpec { int g1 = 0, g2 = 0, neighb_num_max = 0; double r_cut = 0; int update_period = 1; };
struct InteractionSpec {
    std::string name, parameters_file;
    int nl_n = 0;
    std::vector<ListSpec> lists;
    std::vector<double> params;  // parameter-file order, logical `simplified` as 0/1 at the end
};
struct GroupSpec { std::string aux; std::vector<std::string> type_names; };

struct Settings {
    // labels are read and echoed, never interpreted (only the order matters)
    std::string l_step_limit, l_log, l_xyz, l_newvel, l_zmp, l_types, l_groups, l_allmoving, l_xyzmoving, l_zmoving, l_allatoms, l_traj,
        l_ptraj, l_change, l_invert, l_integrators, l_msde, l_nhc, l_temp, l_inter;
    int md_step_limit = 0;
    std::string logfilename, init_xyz_filename;
    bool new_velocities = false;
    int zero_momentum_period = 1;
    int particle_types_num = 0, groups_num = 0;
    std::vector<GroupSpec> groups;
    int all_moving = 1, xyz_moving = 1, z_moving = 1, all_atoms = 1, traj_group = 1, period_traj = 1, change_group_num = 0;
    struct Change { std::string l1, l2; int from, to, ts1, ts2, frec; };
    std::vector<Change> changes;
    bool invert_z_vel = false;
    int integrators_num = 0;
    std::string integrators_header;
    std::vector<IntegratorParams> integrators;  // [0] is the 'none' sentinel of the reference
    double ms_de = 0;
    int nhc_num = 0;
    std::vector<NhcSpec> nhc;
    double initial_temperature = 0;
    int interactions_num = 0;
    std::vector<InteractionSpec> interactions;
};

struct XyzFile {
    double box[3]{0, 0, 0};
    int N = 0;
    std::vector<double> positions, velocities, masses;
    std::vector<std::string> atom_types;
};

// md_read_write.f90:22-61.  Line 2 is read as one character token + 9 reals, so a blank must follow
// `Lattice="`; only elements 1, 5, 9 of the matrix are used (rectangular cells).
inline XyzFile read_xyz(const std::string& path) {
    XyzFile x;
    fio::ListReader r(path);
    x.N = (int)fio::to_int(r.record(1)[0]);
    auto t = r.record(10);
    double m[9];
    for (int k = 0; k < 9; ++k) m[k] = fio::to_real(t[1 + k]);
    x.box[0] = m[0]; x.box[1] = m[4]; x.box[2] = m[8];
    x.positions.resize((size_t)3 * x.N); x.velocities.resize((size_t)3 * x.N); x.masses.resize((size_t)x.N); x.atom_types.resize((size_t)x.N);
    for (int i = 0; i < x.N; ++i) {
        auto a = r.record(8);
        for (int k = 0; k < 3; ++k) { x.positions[3 * i + k] = fio::to_real(a[k]); x.velocities[3 * i + k] = fio::to_real(a[3 + k]); }
        x.masses[i] = fio::to_real(a[6]);
        x.atom_types[i] = a[7].substr(0, 32);
    }
    return x;
}

// number of neighbour-list lines per interaction, md_interactions.f90:65-118
inline int nl_n_for(const std::string& name) {
    if (name == "lj") return 2;
    if (name == "lj1g") return 1;
    if (name == "ljc") return 3;
    if (name == "morsec") return 3;
    if (name == "tb") return 1;
    if (name == "rjl") return 1;
    if (name == "rebosc") return 1;
    return -1;
}

// read_*_parameters: LennardJones.f90:12-21, LennardJones_1g.f90:13-26, LennardJonesCosine.f90:14-24,
// MorseCosine.f90:14-24, TersoffBrenner.f90:13-22, RosatoGuillopeLegrand.f90:12-21, REBOsolidcarbon.f90:12-25
inline std::vector<double> read_parameters(const std::string& name, const std::string& path) {
    fio::ListReader r(path);
    std::vector<double> p;
    auto reals = [&](size_t n) { auto t = r.record(n); for (auto& s : t) p.push_back(fio::to_real(s)); };
    if (name == "lj" || name == "lj1g") { reals(2); reals(2); }
    else if (name == "ljc") { reals(3); reals(2); p.push_back(fio::to_logical(r.record(1)[0]) ? 1. : 0.); }
    else if (name == "morsec") { reals(4); reals(2); p.push_back(fio::to_logical(r.record(1)[0]) ? 1. : 0.); }
    else if (name == "tb") { reals(8); reals(2); }
    else if (name == "rjl") { reals(5); reals(2); }
    else if (name == "rebosc") { reals(3); reals(3); reals(3); reals(1); reals(6); reals(2); }  // REBOsolidcarbon.f90:12-25
    else throw std::runtime_error("error: unknown interaction name " + name);
    return p;
}

// md_simulation.f90:48-93 (order of the reads is the grammar; README's example is stale)
inline Settings read_settings(const std::string& input_path, const std::string& settings_filename) {
    Settings s;
    fio::ListReader r(input_path + settings_filename);
    auto li = [&](std::string& label, int& v) { auto t = r.record(2); label = t[0]; v = (int)fio::to_int(t[1]); };
    auto lb = [&](std::string& label, bool& v) { auto t = r.record(2); label = t[0]; v = fio::to_logical(t[1]); };
    auto ld = [&](std::string& label, double& v) { auto t = r.record(2); label = t[0]; v = fio::to_real(t[1]); };
    auto ls = [&](std::string& label, std::string& v) { auto t = r.record(2); label = t[0]; v = t[1]; };
    li(s.l_step_limit, s.md_step_limit);
    ls(s.l_log, s.logfilename);
    ls(s.l_xyz, s.init_xyz_filename);
    lb(s.l_newvel, s.new_velocities);
    li(s.l_zmp, s.zero_momentum_period);
    li(s.l_types, s.particle_types_num);
    li(s.l_groups, s.groups_num);
    for (int i = 0; i < s.groups_num; ++i) {
        auto t = r.record((size_t)1 + s.particle_types_num);
        GroupSpec g;
        g.aux = t[0];
        for (int k = 0; k < s.particle_types_num; ++k) g.type_names.push_back(t[1 + k].substr(0, 32));
        s.groups.push_back(g);
    }
    li(s.l_allmoving, s.all_moving);
    li(s.l_xyzmoving, s.xyz_moving);
    li(s.l_zmoving, s.z_moving);
    li(s.l_allatoms, s.all_atoms);
    li(s.l_traj, s.traj_group);
    li(s.l_ptraj, s.period_traj);
    li(s.l_change, s.change_group_num);
    for (int i = 0; i < s.change_group_num; ++i) {
        Settings::Change c;
        auto t = r.record(3); c.l1 = t[0]; c.from = (int)fio::to_int(t[1]); c.to = (int)fio::to_int(t[2]);
        auto u = r.record(4); c.l2 = u[0]; c.ts1 = (int)fio::to_int(u[1]); c.ts2 = (int)fio::to_int(u[2]); c.frec = (int)fio::to_int(u[3]);
        s.changes.push_back(c);
    }
    lb(s.l_invert, s.invert_z_vel);
    li(s.l_integrators, s.integrators_num);
    s.integrators_header = r.line();
    s.integrators.emplace_back();  // integrators(0): name 'none', dt 0, l 0
    for (int i = 0; i < s.integrators_num; ++i) {
        auto t = r.record(5);
        IntegratorParams p;
        p.int_name = t[0].substr(0, 32); p.dt = fio::to_real(t[1]); p.l = (int)fio::to_int(t[2]);
        p.period_snapshot = (int)fio::to_int(t[3]); p.period_log = (int)fio::to_int(t[4]);
        s.integrators.push_back(p);
    }
    ld(s.l_msde, s.ms_de);
    li(s.l_nhc, s.nhc_num);
    for (int i = 0; i < s.nhc_num; ++i) {
        auto t = r.record(4);
        NhcSpec n; n.group = (int)fio::to_int(t[0]); n.temperature = fio::to_real(t[1]); n.M = (int)fio::to_int(t[2]); n.q1 = fio::to_real(t[3]);
        s.nhc.push_back(n);
    }
    ld(s.l_temp, s.initial_temperature);
    li(s.l_inter, s.interactions_num);
    for (int i = 0; i < s.interactions_num; ++i) {
        auto t = r.record(2);
        InteractionSpec it;
        it.name = t[0].substr(0, 32); it.parameters_file = t[1].substr(0, 32);
        it.nl_n = nl_n_for(it.name);
        if (it.nl_n < 0) throw std::runtime_error("error: unknown interaction name " + it.name);
        it.params = read_parameters(it.name, input_path + it.parameters_file);
        for (int j = 0; j < it.nl_n; ++j) {
            auto u = r.record(5);
            ListSpec l; l.g1 = (int)fio::to_int(u[0]); l.g2 = (int)fio::to_int(u[1]); l.neighb_num_max = (int)fio::to_int(u[2]);
            l.r_cut = fio::to_real(u[3]); l.update_period = (int)fio::to_int(u[4]);
            it.lists.push_back(l);
        }
        s.interactions.push_back(it);
    }
    return s;
}

}  // namespace pfmds_host
