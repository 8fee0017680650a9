// The md() driver of the run_md_simulation drop-in: settings echo, integrator phases, log / stdout /
// snapshot / trajectory cadence and the end-of-run report, written against an abstract engine so the
// step loop itself (forces, neighbour lists, integration) runs wherever the engine lives — on the
// B200 through the C ABI of include/pfmds_b200.h for the product.
// Mirrors code_source/MOLECULAR_DYNAMICS/md_simulation.f90:19-274 (control flow and every format),
// md_read_write.f90:65-107 (writers), md_interactions.f90:38-57,122-128,427-444 (echo, list load).
//
// Engine concept (all methods throw std::runtime_error on failure, with the reference's message):
//   create(n,pos,vel,mass,box) set_group(g,idx1) set_roles(am,xyz,z,all) add_interaction(spec)
//   add_nhc(spec) set_misc(zero_momentum_period,invert_z) add_group_change(from,to,ts1,ts2,frec) group_size(g) advance(kind,dt,first_md_step,n_steps,energy_after_last)
//   energies(e_inter,ke,temp,e_nhc) diagnostics(fs,mc,mcv,vmax,nl_load) download(pos,vel,frc) save_state(blob) restore_state(pos,vel,blob)
//   timers(t[6])   -> seconds: pos_vel, nlists, nlsearch, nldistance, forces, energy
#pragma once
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "md_inputs.hpp"

namespace pfmds_host {

enum IntegratorKind { KIND_NVE = 0, KIND_NVT = 1, KIND_NVMS = 2 };
inline int kind_of(const std::string& name) { return name == "nvt" ? KIND_NVT : (name == "nvms" ? KIND_NVMS : KIND_NVE); }

inline double wall() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// md_general.f90:57-80: atoms of a group are ordered by the type-name column first, file order second.
inline std::vector<int> group_indexes_1based(const GroupSpec& g, const std::vector<std::string>& atom_types) {
    std::vector<int> idx;
    for (const auto& nm : g.type_names)
        for (size_t i = 0; i < atom_types.size(); ++i)
            if (atom_types[i] == nm) idx.push_back((int)i + 1);
    return idx;
}

// md_general.f90:114-159,328-340 (set_new_temperature).  The reference draws from libgfortran's rand()
// inside an OpenMP region (thread-order dependent, source not in the tree): parity is unpinned there.
// This generator is COUNTER-BASED: the uniform behind draw d of component k of atom i is a pure function
// (SplitMix64 finaliser) of (rand_seed, i, k, d), so the velocities do not depend on the order in which atoms are
// processed, on thread counts or on which rank / device generates them.  Same Marsaglia polar method, same
// sqrt(coef/m) scaling, momentum removal and rescale to the requested temperature as the reference.
inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline double counter_uniform(int rand_seed, uint64_t atom, int k, uint64_t draw) {  // [0, 1), 53 bits
    uint64_t key = mix64((uint64_t)(uint32_t)rand_seed);
    uint64_t ctr = (atom * 3 + (uint64_t)k) * 0x100000000ull + draw;
    return (double)(mix64(key ^ mix64(ctr)) >> 11) * (1.0 / 9007199254740992.0);
}
inline void set_new_temperature(XyzFile& x, const std::vector<int>& idx1, double temp, int rand_seed) {
    const double coef = 1.3806488 / 1.6605389217 * 1.0e-6;
    const double mass_coef = 1.6605389217 / 1.6021765654 * 100.0;
    const double kt_a_degree = 1.3806488 / 1.6021765654 * 1.0e-4;
    for (int i1 : idx1) {
        int i = i1 - 1;
        for (int k = 0; k < 3; ++k) {
            double b = 2., a1 = 0., a2 = 0.;
            for (uint64_t d = 0; b >= 1. || b == 0.; d += 2) {
                a1 = 2. * counter_uniform(rand_seed, (uint64_t)i, k, d) - 1.;
                a2 = 2. * counter_uniform(rand_seed, (uint64_t)i, k, d + 1) - 1.;
                b = a1 * a1 + a2 * a2;
            }
            x.velocities[3 * i + k] = a1 * std::sqrt(-2. * std::log(b) / b) * std::sqrt(coef / x.masses[i]);
        }
    }
    double mcv[3] = {0, 0, 0}, totm = 0;
    for (int i1 : idx1) { int i = i1 - 1; totm += x.masses[i]; for (int k = 0; k < 3; ++k) mcv[k] += x.masses[i] * x.velocities[3 * i + k]; }
    double ke = 0;
    for (int i1 : idx1) {
        int i = i1 - 1;
        double v2 = 0;
        for (int k = 0; k < 3; ++k) { x.velocities[3 * i + k] -= mcv[k] / totm; v2 += x.velocities[3 * i + k] * x.velocities[3 * i + k]; }
        ke += x.masses[i] * v2 / 2 * mass_coef;
    }
    double temperature = 2 * ke / kt_a_degree / (3 * (double)idx1.size());
    double s = std::sqrt(temp / temperature);
    for (int i1 : idx1) for (int k = 0; k < 3; ++k) x.velocities[3 * (i1 - 1) + k] *= s;
}

// md_read_write.f90:65-83
inline void write_particle_group(const std::string& filename, const std::vector<int>& idx1, const double* pos, const double* vel,
                                 const XyzFile& x) {
    std::FILE* f = std::fopen(filename.c_str(), "w");
    if (!f) throw std::runtime_error("cannot open " + filename);
    std::fprintf(f, "%s\n", fio::LI((long)idx1.size()).c_str());
    std::string l = "Lattice=\"";
    const double m[9] = {x.box[0], 0, 0, 0, x.box[1], 0, 0, 0, x.box[2]};
    for (double v : m) l += fio::F(v, 16, 6);
    l += " \" Properties=pos:R:3:vel:R:3:mass:R:1:species:S:1";
    std::fprintf(f, "%s\n", l.c_str());
    for (int i1 : idx1) {
        int i = i1 - 1;
        std::string r;
        for (int k = 0; k < 3; ++k) r += fio::F(pos[3 * i + k], 27, 16);
        for (int k = 0; k < 3; ++k) r += fio::F(vel[3 * i + k], 27, 16);
        r += fio::F(x.masses[i], 27, 16);
        r += "    " + fio::Apad(x.atom_types[i], 32);
        std::fprintf(f, "%s\n", r.c_str());
    }
    std::fclose(f);
}
// md_read_write.f90:86-107
inline void write_particle_group_append(const std::string& filename, const std::vector<int>& idx1, const double* pos, const double* vel,
                                        const XyzFile& x, int md_step) {
    if (idx1.empty()) return;
    std::FILE* f = std::fopen(filename.c_str(), "a");
    if (!f) throw std::runtime_error("cannot open " + filename);
    std::fprintf(f, "%s\n", fio::LI((long)idx1.size()).c_str());
    std::string l = "time_step: " + fio::I(md_step, 9) + "    Lattice=\"";
    const double m[9] = {x.box[0], 0, 0, 0, x.box[1], 0, 0, 0, x.box[2]};
    for (double v : m) l += fio::F(v, 10, 4);
    l += " \" Properties=pos:R:3:vel:R:3:mass:R:1:species:S:1";
    std::fprintf(f, "%s\n", l.c_str());
    for (int i1 : idx1) {
        int i = i1 - 1;
        std::string r;
        for (int k = 0; k < 3; ++k) r += fio::F(pos[3 * i + k], 10, 4);
        for (int k = 0; k < 3; ++k) r += fio::F(vel[3 * i + k], 10, 4);
        r += fio::F(x.masses[i], 10, 4);
        r += "    " + fio::trim(x.atom_types[i]);
        std::fprintf(f, "%s\n", r.c_str());
    }
    std::fclose(f);
}

// Exact restart (not in the reference, whose only restart is a snapshot xyz: md_simulation.f90:233-236).
// A checkpoint holds the full-precision state, the thermostat chains and the driver's own counters, so that
// `-restart <file>` continues the interrupted run: same log rows, same final xyz.
struct MdExtras {
    int checkpoint_period = 0;   // > 0: write <prefix>checkpoint_NNNNNN.chk every that many md steps (rebuild steps of every list only)
    std::string restart_file;    // non-empty: continue from this checkpoint instead of starting at md step 0
};
struct Checkpoint {
    long long n = 0, md_step = 0, integrator_index = 0;
    double simulation_time = 0, potential_energy = 0, prev_potential_energy = 0, kinetic_energy = 0, temperature = 0, total_energy = 0,
           conserved_energy = 0, nose_hoover_energy = 0;
    std::vector<double> e_inter, e_nhc, blob, pos, vel;
};
inline void write_checkpoint(const std::string& path, const Checkpoint& c) {
    std::FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot open " + path);
    const char magic[8] = {'P', 'F', 'M', 'D', 'S', 'C', 'K', '1'};
    std::fwrite(magic, 1, 8, f);
    long long h[3] = {c.n, c.md_step, c.integrator_index};
    std::fwrite(h, sizeof(long long), 3, f);
    double d[8] = {c.simulation_time, c.potential_energy, c.prev_potential_energy, c.kinetic_energy, c.temperature, c.total_energy, c.conserved_energy,
                   c.nose_hoover_energy};
    std::fwrite(d, sizeof(double), 8, f);
    for (const std::vector<double>* v : {&c.e_inter, &c.e_nhc, &c.blob, &c.pos, &c.vel}) {
        long long m = (long long)v->size();
        std::fwrite(&m, sizeof m, 1, f);
        if (m) std::fwrite(v->data(), sizeof(double), (size_t)m, f);
    }
    if (std::fclose(f) != 0) throw std::runtime_error("cannot write " + path);
}
inline Checkpoint read_checkpoint(const std::string& path) {
    std::FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    Checkpoint c;
    char magic[8];
    long long h[3];
    double d[8];
    bool ok = std::fread(magic, 1, 8, f) == 8 && std::string(magic, 8) == "PFMDSCK1" && std::fread(h, sizeof(long long), 3, f) == 3 &&
              std::fread(d, sizeof(double), 8, f) == 8;
    if (ok) {
        c.n = h[0]; c.md_step = h[1]; c.integrator_index = h[2];
        c.simulation_time = d[0]; c.potential_energy = d[1]; c.prev_potential_energy = d[2]; c.kinetic_energy = d[3]; c.temperature = d[4];
        c.total_energy = d[5]; c.conserved_energy = d[6]; c.nose_hoover_energy = d[7];
        for (std::vector<double>* v : {&c.e_inter, &c.e_nhc, &c.blob, &c.pos, &c.vel}) {
            long long m = -1;
            ok = ok && std::fread(&m, sizeof m, 1, f) == 1 && m >= 0 && m < (1ll << 40);
            if (!ok) break;
            v->resize((size_t)m);
            if (m) ok = std::fread(v->data(), sizeof(double), (size_t)m, f) == (size_t)m;
        }
    }
    std::fclose(f);
    if (!ok) throw std::runtime_error("error: " + path + " is not a pfmds checkpoint");
    return c;
}

struct MdResult { int last_step = -1; double simulation_time = 0, total = 0, potential = 0, kinetic = 0, temperature = 0, md_seconds = 0; long atoms = 0; };

template <class Engine>
MdResult md(Engine& eng, std::FILE* out, std::FILE* all_out, const std::string& input_path, const std::string& settings_filename,
            const std::string& output_prefix, int out_period, int num_of_omp_treads, int rand_seed, const MdExtras& extras = MdExtras()) {
    using namespace fio;
    const double exe_time_start = wall();
    auto P = [&](const std::string& s) { std::fprintf(out, "%s\n", s.c_str()); };

    // ---- settings + echo, md_simulation.f90:48-93 ----
    Settings s = read_settings(input_path, settings_filename);
    XyzFile xyz = read_xyz(input_path + s.init_xyz_filename);
    P("settings_filename: " + trim(settings_filename));
    P(A(s.l_step_limit, 32, 128) + I(s.md_step_limit, 12));
    P(A(s.l_log, 32, 128) + "\t" + trim(s.logfilename));
    P(A(s.l_xyz, 32, 128) + "\t" + trim(s.init_xyz_filename));
    P("box_size: " + F(xyz.box[0], 16, 6) + F(xyz.box[1], 16, 6) + F(xyz.box[2], 16, 6));
    P("particles_num: " + I(xyz.N, 12));
    P(A(s.l_newvel, 32, 128) + L(s.new_velocities, 8));
    P(A(s.l_zmp, 32, 128) + I(s.zero_momentum_period, 12));
    P(A(s.l_types, 32, 128) + I(s.particle_types_num, 12));
    P(A(s.l_groups, 32, 128) + I(s.groups_num, 12));
    std::vector<std::vector<int>> groups;
    for (int g = 0; g < s.groups_num; ++g) {
        groups.push_back(group_indexes_1based(s.groups[g], xyz.atom_types));
        std::string l = I(g + 1, 6) + " ";
        for (auto& nm : s.groups[g].type_names) l += A(nm, 12, 32);
        P(l + I((long)groups[g].size(), 9));
    }
    P(A(s.l_allmoving, 32, 128) + I(s.all_moving, 12));
    P(A(s.l_xyzmoving, 32, 128) + I(s.xyz_moving, 12));
    P(A(s.l_zmoving, 32, 128) + I(s.z_moving, 12));
    P(A(s.l_allatoms, 32, 128) + I(s.all_atoms, 12));
    P(A(s.l_traj, 32, 128) + I(s.traj_group, 12));
    P(A(s.l_ptraj, 32, 128) + I(s.period_traj, 12));
    P(A(s.l_change, 32, 128) + I(s.change_group_num, 12));
    for (auto& c : s.changes) {
        P(A(c.l1, 32, 128) + I(c.from, 12) + I(c.to, 12));
        P(A(c.l2, 32, 128) + I(c.ts1, 12) + I(c.ts2, 12) + I(c.frec, 12));
    }
    P(A(s.l_invert, 32, 128) + L(s.invert_z_vel, 8));
    P(A(s.l_integrators, 32, 128) + I(s.integrators_num, 12));
    P(Apad(s.integrators_header, 128));
    for (int i = 1; i <= s.integrators_num; ++i) {
        auto& p = s.integrators[i];
        P("  " + A(p.int_name, 6, 32) + F(p.dt, 10, 5) + I(p.l, 9) + I(p.period_snapshot, 9) + I(p.period_log, 9));
    }
    P(A(s.l_msde, 32, 128) + ES(s.ms_de, 16, 6));
    P(A(s.l_nhc, 32, 128) + I(s.nhc_num, 12));
    for (auto& n : s.nhc) P(I(n.group, 6) + F(n.temperature, 16, 6) + I(n.M, 6) + F(n.q1, 16, 6));
    P(A(s.l_temp, 32, 128) + F(s.initial_temperature, 16, 6));
    if (s.initial_temperature < 0.) s.initial_temperature = 0.;
    auto check_group = [&](int g, const char* what) {
        if (g < 1 || g > s.groups_num) throw std::runtime_error(std::string("error: group number out of range: ") + what);
    };
    check_group(s.all_moving, "all_moving"); check_group(s.xyz_moving, "xyz_moving"); check_group(s.z_moving, "z_moving");
    check_group(s.all_atoms, "all_atoms"); check_group(s.traj_group, "traj_group");
    if (s.new_velocities) set_new_temperature(xyz, groups[s.all_moving - 1], s.initial_temperature, rand_seed);
    P(A(s.l_inter, 32, 128) + I(s.interactions_num, 12));
    for (auto& it : s.interactions) {
        P(A(it.name, 32, 32) + A(it.parameters_file, 32, 32) + I(it.nl_n, 6));
        for (auto& l : it.lists) P(I(l.g1, 6) + I(l.g2, 6) + I(l.neighb_num_max, 6) + F(l.r_cut, 16, 6) + I(l.update_period, 9));
    }

    // ---- hand the system to the engine ----
    eng.create(xyz.N, xyz.positions.data(), xyz.velocities.data(), xyz.masses.data(), xyz.box);
    for (int g = 0; g < s.groups_num; ++g) eng.set_group(g + 1, groups[g]);
    eng.set_roles(s.all_moving, s.xyz_moving, s.z_moving, s.all_atoms);
    for (auto& n : s.nhc) { check_group(n.group, "nhc"); eng.add_nhc(n); }
    eng.set_misc(s.zero_momentum_period, s.invert_z_vel);
    for (auto& c : s.changes) { check_group(c.from, "group_change_from"); check_group(c.to, "group_change_to"); eng.add_group_change(c.from, c.to, c.ts1, c.ts2, c.frec); }
    for (auto& it : s.interactions) eng.add_interaction(it);
    // group%N of a group right now (deposition: the first N entries of its index list, md_general.f90:82-94)
    auto current = [&](int g) {
        std::vector<int> v = groups[(size_t)g - 1];
        if (!s.changes.empty()) v.resize((size_t)eng.group_size(g));
        return v;
    };

    const size_t n_inter = s.interactions.size(), n_nhc = s.nhc.size();
    std::FILE* logf = std::fopen((trim(output_prefix) + trim(s.logfilename)).c_str(), extras.restart_file.empty() ? "w" : "a");
    if (!logf) throw std::runtime_error("cannot open log file " + trim(output_prefix) + trim(s.logfilename));
    P(A("PREPARATIONS TIME: ", 24, 19) + F(wall() - exe_time_start, 10, 2) + " S ");
    P(A("RUNNING ON ", 24, 11) + I(num_of_omp_treads, 6) + " OPENMP THREADS");
    P("");

    const double exe_time_md0 = wall();
    int integrator_index = 0;
    double simulation_time = 0., ts1 = 0.;
    std::string integrator_name;
    double potential_energy = 0., prev_potential_energy = 0., kinetic_energy = 0., temperature = 0., total_energy = 0., conserved_energy = 0.,
           nose_hoover_energy = 0.;
    std::vector<double> e_inter(n_inter, 0.), e_nhc(n_nhc, 0.), e_nhc_dev(n_nhc, 0.);
    std::vector<double> pos((size_t)3 * xyz.N), vel((size_t)3 * xyz.N);
    std::vector<int> nl_load;
    const int limit = s.md_step_limit;
    auto cum_len = [&](int upto) { long c = 0; for (int i = 0; i <= upto && i <= s.integrators_num; ++i) c += s.integrators[i].l; return c; };
    auto needs_energy = [&](int t, int idx, int kind) { return t % s.integrators[idx].period_log == 0 || t % out_period == 0 || kind == KIND_NVMS; };
    auto checkpoint_step = [&](int t) {  // only steps on which every list is rebuilt reproduce the interrupted run exactly
        if (extras.checkpoint_period <= 0 || t == 0 || t % extras.checkpoint_period != 0) return false;
        for (auto& it : s.interactions)
            for (auto& l : it.lists)
                if (l.update_period > 0 && t % l.update_period != 0) return false;
        return true;
    };
    auto event_after = [&](int t, int idx, int kind) {
        return needs_energy(t, idx, kind) || (t % s.integrators[idx].period_snapshot == 0 && t != 0) || t % s.period_traj == 0 || checkpoint_step(t);
    };

    // PFMDS_HOST_STEPWISE_LOG=1: one engine call and one energy read per logged step, as the reference's loop does (A/B and tests)
    const char* sw_env = std::getenv("PFMDS_HOST_STEPWISE_LOG");
    const bool stepwise_log = sw_env && sw_env[0] == '1';
    int md_step = 0;
    bool exited = false;
    if (!extras.restart_file.empty()) {
        Checkpoint ck = read_checkpoint(extras.restart_file);
        if (ck.n != xyz.N || ck.pos.size() != (size_t)3 * xyz.N || ck.vel.size() != (size_t)3 * xyz.N || ck.e_inter.size() != n_inter || ck.e_nhc.size() != n_nhc ||
            ck.integrator_index < 1 || ck.integrator_index > s.integrators_num)
            throw std::runtime_error("error: the checkpoint does not belong to this settings file");
        eng.restore_state(ck.pos.data(), ck.vel.data(), ck.blob);
        {   // The interrupted run may have gone past the checkpoint: keep the header and the rows up to the checkpointed step, drop
            // later rows and its final summary line, so that the continued log holds every row exactly once.
            std::fclose(logf);
            const std::string path = trim(output_prefix) + trim(s.logfilename);
            std::vector<std::string> keep;
            if (std::FILE* in = std::fopen(path.c_str(), "r")) {
                std::string line;
                bool rows_started = false;
                auto flush = [&]() {
                    if (line.size() >= 15) {
                        std::string nm = trim(line.substr(0, 6));
                        if (nm == "nve" || nm == "nvt" || nm == "nvms") {
                            char* end = nullptr;
                            const std::string num = line.substr(6, 9);
                            long st = std::strtol(num.c_str(), &end, 10);
                            if (end && *end == 0) { rows_started = true; if (st <= ck.md_step) keep.push_back(line); line.clear(); return; }
                        }
                    }
                    if (!rows_started) keep.push_back(line);
                    line.clear();
                };
                for (int ch; (ch = std::fgetc(in)) != EOF;) { if (ch == '\n') flush(); else line.push_back((char)ch); }
                if (!line.empty()) flush();
                std::fclose(in);
            }
            logf = std::fopen(path.c_str(), "w");
            if (!logf) throw std::runtime_error("cannot open log file " + path);
            for (auto& l : keep) std::fprintf(logf, "%s\n", l.c_str());
        }
        md_step = (int)ck.md_step + 1;
        integrator_index = (int)ck.integrator_index;
        integrator_name = s.integrators[integrator_index].int_name;
        ts1 = s.integrators[integrator_index].dt;
        simulation_time = ck.simulation_time; potential_energy = ck.potential_energy; prev_potential_energy = ck.prev_potential_energy;
        kinetic_energy = ck.kinetic_energy; temperature = ck.temperature; total_energy = ck.total_energy; conserved_energy = ck.conserved_energy;
        nose_hoover_energy = ck.nose_hoover_energy; e_inter = ck.e_inter; e_nhc = ck.e_nhc;
        P(" restarted from " + trim(extras.restart_file) + " after step" + LI(ck.md_step));
    }
    while (md_step <= limit) {
        // integrator phase switch, md_simulation.f90:121-136
        if (md_step - 1 == cum_len(integrator_index) || integrator_index == 0) {
            integrator_index = integrator_index + 1;
            int i = integrator_index;
            for (; i <= s.integrators_num; ++i)
                if (s.integrators[i].l > 0) { integrator_name = s.integrators[i].int_name; ts1 = s.integrators[i].dt; break; }
            if (i > s.integrators_num || (i == s.integrators_num && s.integrators[s.integrators_num].l < 1)) {
                P(" no more integrators");
                exited = true;
                break;
            }
            integrator_index = i;
        }
        const int kind = kind_of(integrator_name);
        // nvms convergence test, :142-147 (uses the energies of the last evaluated step)
        if (md_step != 0 && kind == KIND_NVMS && std::fabs(potential_energy - prev_potential_energy) <= s.ms_de) {
            P(" potential energy diffrence is small enough");
            exited = true;
            break;
        }
        // How many steps can be queued on the device before the host has to look at anything.  A step whose only event is its
        // log row (`soft`) does not end the queue: its energies are logged on the device (advance_logged) and read with the rest.
        const int period_log = s.integrators[integrator_index].period_log;
        auto soft = [&](int t) {
            return !stepwise_log && kind != KIND_NVMS && t % period_log == 0 && t % out_period != 0 &&
                   !(t % s.integrators[integrator_index].period_snapshot == 0 && t != 0) && t % s.period_traj != 0 && !checkpoint_step(t);
        };
        int n = 1;
        bool soft_inside = false;
        while (md_step + n <= limit && (!event_after(md_step + n - 1, integrator_index, kind) || soft(md_step + n - 1)) &&
               !(md_step + n - 1 == cum_len(integrator_index))) {
            soft_inside |= soft(md_step + n - 1);
            ++n;
        }
        // the energies of a step, folded into the running totals and the log (md_simulation.f90:190-210)
        auto absorb = [&](int t) {
            potential_energy = 0.;
            for (double e : e_inter) potential_energy += e;
            if (kind == KIND_NVT) e_nhc = e_nhc_dev;  // nhc%e is refreshed only while nvt runs (:194-198)
            nose_hoover_energy = 0.;
            for (double e : e_nhc) nose_hoover_energy += e;
            total_energy = potential_energy + kinetic_energy;
            conserved_energy = total_energy + nose_hoover_energy;
            if (t % period_log == 0) {
                std::string l = A(trim(integrator_name), 6, (int)trim(integrator_name).size()) + I(t, 9) + F(simulation_time, 24, 6) +
                                F(conserved_energy, 24, 6) + F(nose_hoover_energy, 24, 6) + F(total_energy, 24, 6) + F(potential_energy, 24, 6) +
                                F(kinetic_energy, 24, 6) + F(temperature, 24, 6);
                for (double e : e_inter) l += F(e, 20, 9);
                for (double e : e_nhc) l += F(e, 20, 9);
                std::fprintf(logf, "%s\n", l.c_str());
            }
        };
        std::vector<EnergyRow> rows;
        if (soft_inside) eng.advance_logged(kind, ts1, md_step, n, period_log, rows);
        else eng.advance(kind, ts1, md_step, n, needs_energy(md_step + n - 1, integrator_index, kind));
        size_t row = 0;
        for (int t = md_step; t < md_step + n; ++t) {
            prev_potential_energy = potential_energy;  // :158
            if (t != 0) simulation_time = simulation_time + ts1;  // :184
            if (soft_inside && t % period_log == 0) {
                if (row >= rows.size()) throw std::runtime_error("error: the engine logged fewer energy rows than steps asked for");
                if (t != md_step + n - 1) {  // the last step of the queue takes the ordinary path below
                    e_inter = rows[row].e_inter; kinetic_energy = rows[row].ke; temperature = rows[row].temp; e_nhc_dev = rows[row].e_nhc;
                    absorb(t);
                }
                ++row;
            }
        }
        md_step += n - 1;  // md_step is now the last executed step

        if (needs_energy(md_step, integrator_index, kind)) {  // :188-231
            eng.energies(e_inter, kinetic_energy, temperature, e_nhc_dev);
            absorb(md_step);
            if (md_step % out_period == 0) {
                double fs[3], mc[3], mcv[3], mav_vel;
                eng.diagnostics(fs, mc, mcv, mav_vel, nl_load);
                double fsn = std::sqrt(fs[0] * fs[0] + fs[1] * fs[1] + fs[2] * fs[2]);
                double mcvn = std::sqrt(mcv[0] * mcv[0] + mcv[1] * mcv[1] + mcv[2] * mcv[2]);
                P("exe time (s) = " + F(wall() - exe_time_start, 10, 2));
                P("step = " + I(md_step, 12) + " integrator = " + A(trim(integrator_name), 6, (int)trim(integrator_name).size()));
                P("conserved energy (eV) = " + ES(conserved_energy, 21, 9));
                P("forces sum (eV/A) = " + ES(fsn, 16, 6));
                P("c.o.m. velocity (A/fs) = " + ES(mcvn, 16, 6));
                P("c.o.m. position (A) = " + F(mc[0], 12, 6) + F(mc[1], 12, 6) + F(mc[2], 12, 6));
                P("maximum velocity (A/fs) = " + F(mav_vel, 12, 6));
                if (n_inter > 0) P("neib lists load:");  // md_interactions.f90:427-444
                size_t k = 0;
                for (auto& it : s.interactions)
                    for (int j = 0; j < it.nl_n; ++j, ++k)
                        P(A(trim(it.name), 8, (int)trim(it.name).size()) + I(j + 1, 3) + I(nl_load[k], 6) + " /" + I(it.lists[j].neighb_num_max, 6));
                if (kind != KIND_NVMS && fsn > 1.0e-14) P(" WARNING: forces sum is too big");
                if (kind == KIND_NVT && s.xyz_moving == s.all_atoms && mcvn > 1.0e-14) P(" WARNING: center of mass velocity is too big");
                P("");
                std::fflush(out);
            }
        }
        const bool snap = md_step % s.integrators[integrator_index].period_snapshot == 0 && md_step != 0;
        const bool traj = md_step % s.period_traj == 0;
        if (snap || traj) eng.download(pos.data(), vel.data(), nullptr);
        if (snap)  // :233-236
            write_particle_group(trim(output_prefix) + "snapshot_" + I(md_step, 0, 6) + ".xyz", current(s.all_atoms), pos.data(), vel.data(), xyz);
        if (traj)  // :238-241
            write_particle_group_append(trim(output_prefix) + "traj_" + I(s.traj_group, 0, 2) + ".xyz", current(s.traj_group), pos.data(),
                                        vel.data(), xyz, md_step);
        if (checkpoint_step(md_step)) {
            Checkpoint ck;
            ck.n = xyz.N; ck.md_step = md_step; ck.integrator_index = integrator_index;
            ck.simulation_time = simulation_time; ck.potential_energy = potential_energy; ck.prev_potential_energy = prev_potential_energy;
            ck.kinetic_energy = kinetic_energy; ck.temperature = temperature; ck.total_energy = total_energy; ck.conserved_energy = conserved_energy;
            ck.nose_hoover_energy = nose_hoover_energy; ck.e_inter = e_inter; ck.e_nhc = e_nhc;
            ck.pos.resize((size_t)3 * xyz.N); ck.vel.resize((size_t)3 * xyz.N);
            eng.save_state(ck.blob);
            eng.download(ck.pos.data(), ck.vel.data(), nullptr);
            write_checkpoint(trim(output_prefix) + "checkpoint_" + I(md_step, 0, 6) + ".chk", ck);
        }
        ++md_step;
    }
    (void)exited;  // after a normal end md_step == limit+1, after an exit it is the step that was not run

    // ---- end-of-run report, :245-272 ----
    const double exe_time_md = wall() - exe_time_md0;
    double t[6] = {0, 0, 0, 0, 0, 0};
    eng.timers(t);
    if (kind_of(integrator_name) == KIND_NVMS) P(" potential energy difference: " + LR(potential_energy - prev_potential_energy));
    P(" steps number:" + LI(md_step - 1));
    P("");
    P(" PERFOMANCE:");
    auto row = [&](const char* name, double v) { P(A(name, 24, (int)std::string(name).size()) + F(v, 10, 2) + " S " + F(v / exe_time_md * 100, 10, 2) + "%"); };
    row("MD:", exe_time_md);
    row("POSITION AND VELOCITY:", t[0]);
    row("NEIGHBOURS:", t[1]);
    row("NEIGHBOURS SEARCH:", t[2]);
    row("NEIGHBOURS DISTANCE:", t[3]);
    row("FORCES:", t[4]);
    row("ENERGY:", t[5]);
    row("REST:", exe_time_md - t[0] - t[1] - t[4] - t[5]);
    P(A("TIME STEPS PER HOUR:", 24, 20) + F(md_step / exe_time_md * 3600, 16, 2));
    P("");

    std::string fin = A(trim(output_prefix), 32, (int)trim(output_prefix).size()) + I(md_step - 1, 9) + F(simulation_time, 20, 9) + F(total_energy, 20, 9) +
                      F(potential_energy, 20, 9) + F(kinetic_energy, 20, 9) + F(temperature, 20, 9);
    for (double e : e_inter) fin += F(e, 20, 9);
    std::fprintf(all_out, "%s", fin.c_str());  // advance='no'
    std::fprintf(logf, "%s", fin.c_str());
    if (md_step != 0) {
        eng.download(pos.data(), vel.data(), nullptr);
        write_particle_group(trim(output_prefix) + "final_" + trim(s.init_xyz_filename), current(s.all_atoms), pos.data(), vel.data(), xyz);
    }
    std::fprintf(logf, "\n");
    std::fclose(logf);
    std::fflush(out);
    std::fflush(all_out);

    MdResult r;
    r.last_step = md_step - 1; r.simulation_time = simulation_time; r.total = total_energy; r.potential = potential_energy;
    r.kinetic = kinetic_energy; r.temperature = temperature; r.md_seconds = exe_time_md; r.atoms = (long)groups[s.all_atoms - 1].size();
    return r;
}

}  // namespace pfmds_host
