// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).  PARITY UNPINNED.
// run_gr_moire_fitting on the CPU restatement (code_source/runners/run_gr_moire_fitting.f90): the host logic of the fit
// (pfmds_b200/host/fit_gr_moire.hpp) is shared, only the engine behind md() differs.
#include <omp.h>

#include "../pfmds_b200/host/fit_gr_moire.hpp"
#include "oracle_engine.hpp"

int main(int argc, char** argv) {
    return pfmds_host::run_gr_moire_fitting_cli(argc, argv, omp_get_max_threads(), [](int threads) {
        omp_set_num_threads(threads);
        return oracle::OracleEngine();
    });
}
