// Command-line front end shared by the run_md_simulation hosts.
// Mirrors code_source/runners/run_md_simulation.f90:11-77 (flags, single run, sequential list) and
// run_md_simulation_mpi.f90:12-105 (ensemble: rank r of n runs list entries i with
// mod(i-1,n)==r-1, per-rank `NNNN-` prefixes, rand_seed = rank).  There is no MPI in this image, so
// the ensemble rank comes from `-node r -nodes n`, or from RANK / WORLD_SIZE (torchrun) when
// `-mpi` is given; ranks never talk to each other in the reference either.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "md_driver.hpp"

namespace pfmds_host {

struct CliOptions {
    int out_period = 1, threads = 1, node_id = 0, nodes = 0;  // node_id 1-based; nodes==0: plain runner
    MdExtras extras;  // -checkpoint_period n, -restart file (single-run mode)
    int streams = 1;  // ensemble mode: runs of this rank executed concurrently (one host thread + one context/stream each)
    std::string settings_filename = "md_run_settings.txt", settings_files_list, all_out_file = "all_out.txt", output_prefix, input_path, out_path;
};

template <class Factory>
int run_cli(int argc, char** argv, int default_threads, Factory make_engine) {
    using namespace fio;
    CliOptions o;
    o.threads = default_threads;
    const std::string line(80, '_');
    std::vector<std::string> echo;
    bool mpi = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> std::string { return (i + 1 < argc) ? std::string(argv[++i]) : std::string(); };
        if (a == "-i" || a == "--input") o.settings_filename = next();
        else if (a == "-p" || a == "--prefix") { o.output_prefix = next(); echo.push_back("output_prefix:  " + o.output_prefix); }
        else if (a == "-op" || a == "--out_period") { std::string v = next(); echo.push_back("out_period:         " + v); o.out_period = (int)to_int(v); }
        else if (a == "-omp_n" || a == "--openmp_threads_num") { std::string v = next(); echo.push_back("openmp_threads_num: " + v); o.threads = (int)to_int(v); }
        else if (a == "-ipath" || a == "--input_path") { o.input_path = next(); echo.push_back("input_path:         " + o.input_path); }
        else if (a == "-ilist" || a == "--input_list") { o.settings_files_list = next(); echo.push_back("input_list:         " + o.settings_files_list); }
        else if (a == "-opath" || a == "--out_path") { o.out_path = next(); echo.push_back("out_path:           " + o.out_path); }
        else if (a == "-ofile" || a == "--all_out_file") { o.all_out_file = next(); echo.push_back("all_out_file:   " + o.all_out_file); }
        else if (a == "-mpi") mpi = true;
        else if (a == "-node") { o.node_id = (int)to_int(next()); mpi = true; }
        else if (a == "-nodes") { o.nodes = (int)to_int(next()); mpi = true; }
        else if (a == "-streams") o.streams = (int)to_int(next());
        else if (a == "-checkpoint_period") o.extras.checkpoint_period = (int)to_int(next());
        else if (a == "-restart") o.extras.restart_file = next();
        // unknown flags are silently ignored, like the reference's select case
    }
    if (mpi) {
        if (o.nodes == 0) {
            const char* ws = std::getenv("WORLD_SIZE"); const char* rk = std::getenv("RANK");
            o.nodes = ws ? std::atoi(ws) : 1;
            o.node_id = (rk ? std::atoi(rk) : 0) + 1;
        }
        if (o.node_id < 1) o.node_id = 1;
    }
    int rc = 0;
    try {
        if (!mpi) {  // run_md_simulation.f90
            std::FILE* out = stdout;
            std::fprintf(out, "%s\n", line.c_str());
            for (auto& e : echo) std::fprintf(out, "%s\n", e.c_str());
            if (!o.settings_files_list.empty()) {
                ListReader lst(o.input_path + o.settings_files_list);
                std::FILE* all_out = std::fopen((o.out_path + o.all_out_file).c_str(), "w");
                if (!all_out) throw std::runtime_error("cannot open " + o.out_path + o.all_out_file);
                int set_num = (int)to_int(lst.record(1)[0]);
                for (int i = 1; i <= set_num; ++i) {
                    auto t = lst.record(2);
                    std::string prefix = o.out_path + t[1];
                    std::fprintf(out, "%s\n", line.c_str());
                    std::fprintf(out, "%s\t%s%s\n", I(i, 6).c_str(), A(t[0], 32, 128).c_str(), A(prefix, 32, 128).c_str());
                    auto eng = make_engine(o.threads);
                    md(eng, out, all_out, o.input_path, t[0], prefix, o.out_period, o.threads, 1);
                    std::fprintf(all_out, "\n");
                    std::fprintf(out, "%s\n", line.c_str());
                }
                std::fclose(all_out);
            } else {
                std::fprintf(out, "%s\n", line.c_str());
                auto eng = make_engine(o.threads);
                md(eng, out, out, o.input_path, o.settings_filename, o.output_prefix, o.out_period, o.threads, 1, o.extras);
                std::fprintf(out, "\n%s\n", line.c_str());
            }
            std::fprintf(out, "%s\n", line.c_str());
        } else {  // run_md_simulation_mpi.f90
            const std::string node = I(o.node_id, 0, 4) + "-";
            std::FILE* out = std::fopen((o.output_prefix + node + "out.txt").c_str(), "w");
            if (!out) throw std::runtime_error("cannot open " + o.output_prefix + node + "out.txt");
            std::fprintf(out, "output_prefix:      %s\n", o.output_prefix.c_str());
            std::fprintf(out, "out_period:         %s\n", I(o.out_period, 9).c_str());
            std::fprintf(out, "openmp_threads_num: %s\n", I(o.threads, 9).c_str());
            std::fprintf(out, "input_path:         %s\n", o.input_path.c_str());
            std::fprintf(out, "input_list:         %s\n", o.settings_files_list.c_str());
            std::fprintf(out, "out_path:           %s\n", o.out_path.c_str());
            std::fprintf(out, "all_out_file:       %s\n", o.all_out_file.c_str());
            std::fprintf(out, "%s\n", line.c_str());
            const int rand_seed = o.node_id;
            if (!o.settings_files_list.empty()) {
                std::FILE* all_out = std::fopen((o.out_path + o.output_prefix + node + o.all_out_file).c_str(), "w");
                if (!all_out) throw std::runtime_error("cannot open the all_out file");
                ListReader lst(o.input_path + o.settings_files_list);
                int set_num = (int)to_int(lst.record(1)[0]);
                if (o.nodes <= set_num) {
                    // this rank's entries; with -streams k > 1 they run k at a time, each on its own host thread with its own
                    // engine (context + stream), and their outputs are appended in list order afterwards
                    struct Job { int i; std::string settings, prefix; char* obuf = nullptr; size_t olen = 0; char* abuf = nullptr; size_t alen = 0; std::string err; };
                    std::vector<Job> jobs;
                    for (int i = 1; i <= set_num; ++i) {
                        auto t = lst.record(2);
                        if ((i - 1) % o.nodes == o.node_id - 1) { Job j; j.i = i; j.settings = t[0]; j.prefix = o.out_path + o.output_prefix + node + t[1]; jobs.push_back(j); }
                    }
                    auto run_job = [&](Job& j) {
                        std::FILE* jo = open_memstream(&j.obuf, &j.olen);
                        std::FILE* ja = open_memstream(&j.abuf, &j.alen);
                        std::fprintf(jo, "%s\n", line.c_str());
                        std::fprintf(jo, "RUNNING ON NODE %s OUT OF%s NODES\n", I(o.node_id, 6).c_str(), I(o.nodes, 6).c_str());
                        std::fprintf(jo, "%s\t%s\t%s\n", I(j.i, 6).c_str(), j.settings.c_str(), j.prefix.c_str());
                        try {
                            auto eng = make_engine(o.threads);
                            md(eng, jo, ja, o.input_path, j.settings, j.prefix, o.out_period, o.threads, rand_seed);
                        } catch (const std::exception& e) { j.err = e.what(); }
                        std::fprintf(ja, "\n");
                        std::fprintf(jo, "%s\n", line.c_str());
                        std::fclose(jo); std::fclose(ja);
                    };
                    const size_t k = (size_t)(o.streams < 1 ? 1 : o.streams);
                    for (size_t b = 0; b < jobs.size(); b += k) {
                        std::vector<std::thread> th;
                        for (size_t q = b; q < jobs.size() && q < b + k; ++q) th.emplace_back(run_job, std::ref(jobs[q]));
                        for (auto& t : th) t.join();
                        for (size_t q = b; q < jobs.size() && q < b + k; ++q) {
                            std::fwrite(jobs[q].obuf, 1, jobs[q].olen, out);
                            std::fwrite(jobs[q].abuf, 1, jobs[q].alen, all_out);
                            std::free(jobs[q].obuf); std::free(jobs[q].abuf);
                            if (!jobs[q].err.empty()) { std::fclose(all_out); std::fclose(out); throw std::runtime_error(jobs[q].err); }
                        }
                    }
                } else {
                    std::fprintf(out, "error: too many mpi nodes (%s) for this list (%s)%s\n", I(o.nodes, 6).c_str(), I(set_num, 6).c_str(),
                                 o.settings_files_list.c_str());
                }
                std::fclose(all_out);
            } else {
                std::fprintf(out, "%s\n", line.c_str());
                std::fprintf(out, "RUNNING ON NODE %s OUT OF%s NODES. EACH NODE RUNS THE SAME SIMULATION.\n", I(o.node_id, 6).c_str(), I(o.nodes, 6).c_str());
                std::string str = o.out_path + o.output_prefix + node;
                std::fprintf(out, "%s\t%s\t%s\n", I(1, 6).c_str(), o.settings_filename.c_str(), str.c_str());
                auto eng = make_engine(o.threads);
                md(eng, out, out, o.input_path, o.settings_filename, str, o.out_period, o.threads, rand_seed);
                std::fprintf(out, "\n%s\n", line.c_str());
            }
            std::fprintf(out, "%s\n", line.c_str());
            std::fclose(out);
        }
    } catch (const std::exception& e) {
        std::fflush(stdout);
        std::fprintf(stdout, " %s\n", e.what());  // the reference prints its message and `stop`s
        std::fflush(stdout);
        rc = 1;
    }
    return rc;
}

}  // namespace pfmds_host
