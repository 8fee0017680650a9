// The consumers of md() around the hot path: the z-height analysis of graphene on a metal surface and the
// golden-section fitting loop over the ljc / morsec parameters that calls md() twice per evaluation.
// Mirrors code_source/graphene_on_surface_analysis/graphene_on_surface_analysis.f90:8-30, md_general.f90:400-421
// (position_analysis), code_source/runners/run_gr_analysis.f90, code_source/ljc_and_morsec_moire_graphene_fitting/
// fit_gr_moire.f90:16-182 and code_source/runners/run_gr_moire_fitting.f90 — same command lines, same fitting-parameters
// file, same file renames, same fit_out.txt rows.  Written against the engine concept of md_driver.hpp, so every md() call of
// the fit runs its step loop on the B200 through the C ABI (or on the CPU oracle in the tests).
// Extension: `-pair` runs the two cells of one evaluation concurrently (two host threads, two contexts / CUDA streams).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "md_driver.hpp"

namespace pfmds_host {

// md_general.f90:400-421: average, minimum and maximum of one coordinate over the atoms of a group inside (minimum, maximum)
inline void position_analysis(double& av, double& mi, double& ma, const XyzFile& x, const std::vector<int>& idx1, int direction, double minimum,
                              double maximum) {
    int k = 0;
    av = 0.;
    mi = maximum;
    ma = minimum;
    for (int i1 : idx1) {
        double v = x.positions[3 * (size_t)(i1 - 1) + (size_t)(direction - 1)];
        if (v < maximum && v > minimum) {
            av = av + v;
            k = k + 1;
            if (v > ma) ma = v;
            if (v < mi) mi = v;
        }
    }
    av = av / k;  // k == 0 gives NaN, like the reference
}

// graphene_on_surface_analysis.f90:8-30: arr1 = (average, min, max) z of the carbon atoms, arr2 = of the metal atoms
inline void gr_on_cu_analysis(double arr1[3], double arr2[3], const std::string& filename, double z) {
    XyzFile x = read_xyz(filename);
    GroupSpec gc, gcu;
    gc.type_names = {"C", "C_a", "C_b"};
    gcu.type_names = {"CU", "CU_fixed", "#"};
    position_analysis(arr1[0], arr1[1], arr1[2], x, group_indexes_1based(gc, x.atom_types), 3, z, 1000.);
    position_analysis(arr2[0], arr2[1], arr2[2], x, group_indexes_1based(gcu, x.atom_types), 3, z, 1000.);
}

// runners/run_gr_analysis.f90
inline int run_gr_analysis_cli(int argc, char** argv) {
    using namespace fio;
    std::string path, filelist = "filelist.txt", outfilename = "outfilename.txt";
    double z = 0.;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> std::string { return (i + 1 < argc) ? std::string(argv[++i]) : std::string(); };
        if (a == "-path") path = next();
        else if (a == "-fl") filelist = next();
        else if (a == "-o") outfilename = next();
        else if (a == "-z") z = to_real(next());
    }
    try {
        std::printf(" %s\n", trim(path + filelist).c_str());
        std::printf(" %s\n", trim(path + outfilename).c_str());
        ListReader in(path + filelist);
        std::FILE* out = std::fopen((path + outfilename).c_str(), "w");
        if (!out) throw std::runtime_error("cannot open " + path + outfilename);
        int n = (int)to_int(in.record(1)[0]);
        std::printf("%s\n", LI(n).c_str());
        for (int i = 1; i <= n; ++i) {
            std::string filename = in.record(1)[0];
            double arr1[3], arr2[3];
            gr_on_cu_analysis(arr1, arr2, path + filename, z);
            std::printf("%s", I(i, 10).c_str());
            std::fprintf(out, " %s    %s%s%s\n", trim(filename).c_str(), LR(arr1[0] - arr2[0]).c_str(), LR(arr1[1] - arr2[0]).c_str(), LR(arr1[2] - arr2[0]).c_str());
        }
        std::fclose(out);
        std::printf("\n");
    } catch (const std::exception& e) {
        std::printf(" %s\n", e.what());
        return 1;
    }
    return 0;
}

// ---- fit_gr_moire.f90 ----------------------------------------------------------------------------
struct FitGrMoire {
    int sim_num = 0, out_period = 1, num_of_omp_treads = 1;
    int ar_c_num[2]{0, 0};
    std::string interaction_name, ar_settings_filename[2], output_prefix, input_path, out_path, ar_final_file[2], param_file, ar_start_xyz_file[2],
        ar_xyz_file[2];
    double z = 0, ar_zero_energy_level[2]{0, 0}, be0 = 0, ar_grd0[2]{0, 0}, Rcut[2]{0, 0};
    bool simplified = false, pair_concurrent = false;
    std::string line = std::string(80, '_');
    std::FILE* out = stdout;   // out_id
    std::FILE* oid = nullptr;  // fit_out.txt

    // fit_gr_moire.f90:107-182
    void set_fitting_parameters(const std::string& fitting_parameters_file_name, double init_min_params[4], double init_max_params[4]) {
        using namespace fio;
        ListReader r(fitting_parameters_file_name);
        auto echo_s = [&](const std::string& l, const std::string& v) { std::fprintf(out, " %s  %s\n", trim(l).c_str(), trim(v).c_str()); };
        auto a16 = [&](std::string& v) {  // read(9,'(A16,A)') str,path
            std::string l = r.line();
            std::string s = l.substr(0, std::min<size_t>(16, l.size()));
            v = l.size() > 16 ? trim(l.substr(16)) : std::string();
            echo_s(s, v);
        };
        auto ls = [&](std::string& v) { auto t = r.record(2); v = t[1]; echo_s(t[0], v); };
        auto li = [&](int& v) { auto t = r.record(2); v = (int)to_int(t[1]); std::fprintf(out, " %s  %s\n", trim(t[0]).c_str(), LI(v).c_str()); };
        auto ld = [&](double& v) { auto t = r.record(2); v = to_real(t[1]); std::fprintf(out, " %s  %s\n", trim(t[0]).c_str(), LR(v).c_str()); };
        std::string min_param_file, max_param_file;
        a16(input_path);
        a16(out_path);
        ls(interaction_name);
        ls(output_prefix);
        for (int i = 0; i < 2; ++i) {
            ls(ar_settings_filename[i]);
            ls(ar_start_xyz_file[i]);
            li(ar_c_num[i]);
            ld(ar_zero_energy_level[i]);
            ls(ar_final_file[i]);
        }
        ls(min_param_file);
        ls(max_param_file);
        ld(z);
        ld(be0);
        ld(ar_grd0[0]);
        ld(ar_grd0[1]);

        simplified = false;
        if (interaction_name == "ljc") {
            ListReader a(input_path + min_param_file);
            auto t = a.record(3);
            init_min_params[2] = to_real(t[0]); init_min_params[0] = to_real(t[1]); init_min_params[1] = to_real(t[2]);
            auto c = a.record(2);
            Rcut[0] = to_real(c[0]); Rcut[1] = to_real(c[1]);
            ListReader b(input_path + max_param_file);
            auto u = b.record(3);
            init_max_params[2] = to_real(u[0]); init_max_params[0] = to_real(u[1]); init_max_params[1] = to_real(u[2]);
            init_min_params[3] = 0.;
            init_max_params[3] = 0.;
        }
        if (interaction_name == "morsec") {
            ListReader a(input_path + min_param_file);
            auto t = a.record(4);
            init_min_params[2] = to_real(t[0]); init_min_params[0] = to_real(t[1]); init_min_params[3] = to_real(t[2]); init_min_params[1] = to_real(t[3]);
            auto c = a.record(2);
            Rcut[0] = to_real(c[0]); Rcut[1] = to_real(c[1]);
            ListReader b(input_path + max_param_file);
            auto u = b.record(4);
            init_max_params[2] = to_real(u[0]); init_max_params[0] = to_real(u[1]); init_max_params[3] = to_real(u[2]); init_max_params[1] = to_real(u[3]);
        }
        sim_num = 0;
        {   // the xyz file name is on the third line of a settings file; the parameter file follows the interaction name at column 1
            ListReader s1(input_path + ar_settings_filename[0]);
            s1.skip(); s1.skip();
            ar_xyz_file[0] = s1.record(2)[1];
            const size_t nl = interaction_name.size();
            while (true) {
                std::string str = Apad(s1.line(), 128);
                auto parse = [&](size_t from) {
                    std::string rest = str.substr(from);
                    size_t p = rest.find_first_not_of(" \t");
                    if (p == std::string::npos) throw std::runtime_error("fit: no parameter file after the interaction name");
                    size_t q = rest.find_first_of(" \t", p);
                    param_file = rest.substr(p, q == std::string::npos ? std::string::npos : q - p);
                };
                // Fortran compares str(1:3) / str(1:6) with the name blank-padded to the same length
                if (nl <= 3 && trim(str.substr(0, 3)) == interaction_name) { parse(3); break; }
                if (nl <= 6 && trim(str.substr(0, 6)) == interaction_name) { parse(6); break; }
            }
            ListReader s2(input_path + ar_settings_filename[1]);
            s2.skip(); s2.skip();
            ar_xyz_file[1] = s2.record(2)[1];
        }
    }

    // fit_gr_moire.f90:16-105: write the parameter file, relax both cells with md(), measure binding energy / distance / corrugation
    template <class Factory>
    void calc_error(double& error, bool from_init_xyz, const double params[4], Factory& make_engine) {
        using namespace fio;
        error = 0;
        sim_num = sim_num + 1;
        std::fprintf(out, "\n");
        std::fprintf(out, " %s\n", (trim(input_path) + trim(param_file)).c_str());
        if (interaction_name == "ljc") {
            std::fprintf(out, "%s%s%s%s\n", LI(sim_num).c_str(), LR(params[2]).c_str(), LR(params[0]).c_str(), LR(params[1]).c_str());
            std::fprintf(oid, "%s%s%s%s", I(sim_num, 6).c_str(), F(params[2], 21, 6).c_str(), F(params[0], 21, 6).c_str(), F(params[1], 21, 6).c_str());
            std::FILE* f = std::fopen((trim(input_path) + trim(param_file)).c_str(), "w");
            if (!f) throw std::runtime_error("cannot write " + input_path + param_file);
            std::fprintf(f, "%s%s%s\n", LR(params[2]).c_str(), LR(params[0]).c_str(), LR(params[1]).c_str());
            std::fprintf(f, "%s%s\n", LR(Rcut[0]).c_str(), LR(Rcut[1]).c_str());
            std::fprintf(f, " %s\n", simplified ? "T" : "F");
            std::fclose(f);
        }
        if (interaction_name == "morsec") {
            std::fprintf(out, "%s%s%s%s%s\n", LI(sim_num).c_str(), LR(params[2]).c_str(), LR(params[0]).c_str(), LR(params[3]).c_str(), LR(params[1]).c_str());
            std::fprintf(oid, "%s%s%s%s%s", I(sim_num, 6).c_str(), F(params[2], 21, 6).c_str(), F(params[0], 21, 6).c_str(), F(params[3], 21, 6).c_str(),
                         F(params[1], 21, 6).c_str());
            std::FILE* f = std::fopen((trim(input_path) + trim(param_file)).c_str(), "w");
            if (!f) throw std::runtime_error("cannot write " + input_path + param_file);
            std::fprintf(f, "%s%s%s%s\n", LR(params[2]).c_str(), LR(params[0]).c_str(), LR(params[3]).c_str(), LR(params[1]).c_str());
            std::fprintf(f, "%s%s\n", LR(Rcut[0]).c_str(), LR(Rcut[1]).c_str());
            std::fprintf(f, " %s\n", simplified ? "T" : "F");
            std::fclose(f);
        }
        const std::string str = I(sim_num, 0, 6), prev = I(sim_num - 1, 0, 6);
        const std::string op = trim(out_path) + trim(output_prefix) + str + "_", prevop = trim(out_path) + trim(output_prefix) + prev + "_";

        struct Cell { std::string xyz_in, final_path, md_out, err; double arr1[3], arr2[3], be = 0; };
        Cell cell[2];
        auto stage_in = [&](int i) {
            cell[i].xyz_in = trim(input_path) + trim(ar_xyz_file[i]);
            if (from_init_xyz) std::rename((trim(input_path) + trim(ar_start_xyz_file[i])).c_str(), cell[i].xyz_in.c_str());
            else std::rename((prevop + "final_" + trim(ar_xyz_file[i])).c_str(), cell[i].xyz_in.c_str());
            cell[i].final_path = op + trim(ar_final_file[i]);
        };
        auto relax = [&](int i) {  // md() + analysis; stdout of md() goes to a buffer so that concurrent cells do not interleave
            char* buf = nullptr;
            size_t len = 0;
            std::FILE* mo = open_memstream(&buf, &len);
            std::FILE* fo = std::fopen(cell[i].final_path.c_str(), "w");
            try {
                if (!fo) throw std::runtime_error("cannot open " + cell[i].final_path);
                auto eng = make_engine(num_of_omp_treads);
                md(eng, mo, fo, input_path, ar_settings_filename[i], op, out_period, num_of_omp_treads, 1);
                gr_on_cu_analysis(cell[i].arr1, cell[i].arr2, op + "final_" + trim(ar_xyz_file[i]), z);
                // the all_out row was written without a line end: the analysis lands on the same record (fit_gr_moire.f90:70)
                std::fprintf(fo, "%s%s%s\n", F(cell[i].arr1[0] - cell[i].arr2[0], 16, 6).c_str(), F(cell[i].arr1[1] - cell[i].arr2[0], 16, 6).c_str(),
                             F(cell[i].arr1[2] - cell[i].arr2[0], 16, 6).c_str());
            } catch (const std::exception& e) { cell[i].err = e.what(); }
            if (fo) std::fclose(fo);
            std::fclose(mo);
            cell[i].md_out.assign(buf ? buf : "", len);
            std::free(buf);
        };
        auto stage_out = [&](int i) {
            if (from_init_xyz) std::rename(cell[i].xyz_in.c_str(), (trim(input_path) + trim(ar_start_xyz_file[i])).c_str());
            else std::rename(cell[i].xyz_in.c_str(), (prevop + "final_" + trim(ar_xyz_file[i])).c_str());
        };
        const bool together = pair_concurrent && ar_xyz_file[0] != ar_xyz_file[1] && ar_final_file[0] != ar_final_file[1] &&
                              ar_settings_filename[0] != ar_settings_filename[1];
        if (together) {
            stage_in(0); stage_in(1);
            std::thread t0(relax, 0), t1(relax, 1);
            t0.join(); t1.join();
        }
        for (int i = 0; i < 2; ++i) {
            if (!together) { stage_in(i); relax(i); }
            std::fprintf(out, "%s\n", line.c_str());
            std::fwrite(cell[i].md_out.data(), 1, cell[i].md_out.size(), out);
            std::fprintf(out, "%s\n", line.c_str());
            stage_out(i);
            if (!cell[i].err.empty()) throw std::runtime_error(cell[i].err);
            const double* arr1 = cell[i].arr1;
            const double* arr2 = cell[i].arr2;
            std::fprintf(out, " gr_on_cu_analysis:%s%s%s%s%s%s\n", LR(arr1[0]).c_str(), LR(arr1[1]).c_str(), LR(arr1[2]).c_str(), LR(arr2[0]).c_str(),
                         LR(arr2[1]).c_str(), LR(arr2[2]).c_str());
            {   // read(final_out_id,'(A61,f20.9,A)'): the total energy of the all_out row (A32, i9, then the second f20.9)
                std::ifstream fin(cell[i].final_path);
                std::string l;
                std::getline(fin, l);
                if (l.size() < 81) throw std::runtime_error("fit: short all_out row in " + cell[i].final_path);
                std::string f20 = trim(l.substr(61, 20));
                size_t p0 = f20.find_first_not_of(' ');
                cell[i].be = to_real(p0 == std::string::npos ? std::string() : f20.substr(p0));
            }
            double bd = arr1[0] - arr2[0];
            double grd = arr1[2] - arr1[1];
            double be = (cell[i].be - ar_zero_energy_level[i]) / ar_c_num[i];
            std::fprintf(oid, "%s%s%s", F(be, 21, 6).c_str(), F(bd, 21, 6).c_str(), F(grd, 21, 6).c_str());
            if (i == 0) {
                double e1 = (be / be0 - 1.) * (be / be0 - 1.), e2 = (grd / ar_grd0[i] - 1.) * (grd / ar_grd0[i] - 1.);
                error = e1 + e2;
                std::fprintf(oid, "%s%s", F(e1, 21, 6).c_str(), F(e2, 21, 6).c_str());
            }
            if (i == 1) {
                double e2 = (grd / ar_grd0[i] - 1.) * (grd / ar_grd0[i] - 1.);
                error = error + e2;
                std::fprintf(oid, "%s", F(e2, 21, 6).c_str());
            }
        }
        std::fprintf(oid, "%s\n", F(error, 21, 6).c_str());
        std::fflush(oid);
        std::fflush(out);
        std::rename((trim(input_path) + trim(param_file)).c_str(), (trim(input_path) + str + trim(param_file)).c_str());
    }
};

// runners/run_gr_moire_fitting.f90: golden-section search over the first three parameters, one at a time, until the fit error
// stops changing.  Evaluations inside a bracket restart from the previous relaxed configuration (from_init_xyz = .false.).
template <class Factory>
int run_gr_moire_fitting_cli(int argc, char** argv, int default_threads, Factory make_engine) {
    using namespace fio;
    const double gold = (std::sqrt(5.) - 1.) / (std::sqrt(5.) + 1.);
    FitGrMoire fit;
    fit.num_of_omp_treads = default_threads;
    double delta_error_gold = 0., delta_error_fit = 0.;
    std::string fitting_parameters_file_name;
    std::FILE* out = stdout;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> std::string { return (i + 1 < argc) ? std::string(argv[++i]) : std::string(); };
        if (a == "-op" || a == "--out_period") { std::string v = next(); fit.out_period = (int)to_int(v); std::fprintf(out, "out_period: %s\n", v.c_str()); }
        else if (a == "-omp_n" || a == "--openmp_threads_num") { std::string v = next(); fit.num_of_omp_treads = (int)to_int(v); std::fprintf(out, "openmp_threads_num: %s\n", v.c_str()); }
        else if (a == "-fpfn") { fitting_parameters_file_name = next(); std::fprintf(out, "fitting_parameters_file_name: %s\n", fitting_parameters_file_name.c_str()); }
        else if (a == "-delta_error_gold") { std::string v = next(); delta_error_gold = to_real(v); std::fprintf(out, "delta_error_gold: %s\n", v.c_str()); }
        else if (a == "-delta_error_fit") { std::string v = next(); delta_error_fit = to_real(v); std::fprintf(out, "delta_error_fit: %s\n", v.c_str()); }
        else if (a == "-pair") fit.pair_concurrent = true;
    }
    try {
        double init_min_params[4] = {0, 0, 0, 0}, init_max_params[4] = {0, 0, 0, 0}, params[4], min_params[4], max_params[4];
        fit.set_fitting_parameters(fitting_parameters_file_name, init_min_params, init_max_params);
        fit.oid = std::fopen((trim(fit.out_path) + trim(fit.output_prefix) + "fit_out.txt").c_str(), "w");
        if (!fit.oid) throw std::runtime_error("cannot open " + fit.out_path + fit.output_prefix + "fit_out.txt");
        std::fprintf(out, "%s\n", fit.line.c_str());
        for (int k = 0; k < 4; ++k) { min_params[k] = init_min_params[k]; max_params[k] = init_max_params[k]; params[k] = (init_min_params[k] + init_max_params[k]) / 2; }
        double error = 0., prev_error_fit = 0., error_array[4] = {0, 0, 0, 0};
        auto delta_error = [&]() {
            double mx = error_array[0], mn = error_array[0];
            for (double e : error_array) { if (e > mx) mx = e; if (e < mn) mn = e; }
            return mx - mn;
        };
        while (std::fabs(prev_error_fit - error) > delta_error_fit || fit.sim_num == 0) {
            for (int k = 0; k < 3; ++k) {
                prev_error_fit = error;
                if (fit.sim_num == 0) prev_error_fit = 1000000.;
                min_params[k] = init_min_params[k];
                max_params[k] = init_max_params[k];
                params[k] = min_params[k];
                fit.calc_error(error_array[0], true, params, make_engine);
                params[k] = max_params[k];
                fit.calc_error(error_array[3], true, params, make_engine);
                params[k] = min_params[k] + (max_params[k] - min_params[k]) * gold;
                fit.calc_error(error_array[1], true, params, make_engine);
                params[k] = max_params[k] - (max_params[k] - min_params[k]) * gold;
                fit.calc_error(error_array[2], true, params, make_engine);
                double de = delta_error();
                std::fprintf(out, "  delta_error: %s\n", LR(de).c_str());
                int gold_i = 0;
                while (de > delta_error_gold && gold_i <= 20) {
                    gold_i = gold_i + 1;
                    if (error_array[1] < error_array[2]) {
                        error_array[3] = error_array[2];
                        error_array[2] = error_array[1];
                        max_params[k] = max_params[k] - (max_params[k] - min_params[k]) * gold;
                        params[k] = min_params[k] + (max_params[k] - min_params[k]) * gold;
                        fit.calc_error(error_array[1], false, params, make_engine);
                    } else {
                        error_array[0] = error_array[1];
                        error_array[1] = error_array[2];
                        min_params[k] = min_params[k] + (max_params[k] - min_params[k]) * gold;
                        params[k] = max_params[k] - (max_params[k] - min_params[k]) * gold;
                        fit.calc_error(error_array[2], false, params, make_engine);
                    }
                    de = delta_error();
                    std::fprintf(out, "  delta_error: %s\n", LR(de).c_str());
                }
                int loc = 0;
                for (int q = 1; q < 4; ++q) if (error_array[q] < error_array[loc]) loc = q;  // minloc: first minimum
                error = error_array[loc];
                switch (loc) {
                case 1: params[k] = min_params[k] + (max_params[k] - min_params[k]) * gold; break;
                case 2: params[k] = max_params[k] - (max_params[k] - min_params[k]) * gold; break;
                case 0: params[k] = min_params[k]; break;
                case 3: params[k] = max_params[k]; break;
                }
                std::fprintf(out, "  parameters: %s%s%s%s\n", LR(params[0]).c_str(), LR(params[1]).c_str(), LR(params[2]).c_str(), LR(params[3]).c_str());
            }
        }
        std::fclose(fit.oid);
        std::fprintf(out, "%s\n", fit.line.c_str());
    } catch (const std::exception& e) {
        std::fflush(out);
        std::fprintf(out, " %s\n", e.what());
        if (fit.oid) std::fclose(fit.oid);
        return 1;
    }
    return 0;
}

}  // namespace pfmds_host
