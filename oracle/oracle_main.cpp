// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).  PARITY UNPINNED.
// run_md_simulation on the CPU restatement: same flags as code_source/runners/run_md_simulation.f90:26-76.
// This binary is the CPU baseline that bench.py times (`cpu_baseline`, `--impl reference`).
#include <omp.h>

#include <cstdio>
#include <cstring>
#include <string>

#include "../pfmds_b200/host/md_driver.hpp"
#include "../pfmds_b200/host/runner_cli.hpp"
#include "oracle_engine.hpp"

int main(int argc, char** argv) {
    return pfmds_host::run_cli(argc, argv, omp_get_max_threads(), [](int threads) {
        omp_set_num_threads(threads);
        return oracle::OracleEngine();
    });
}
