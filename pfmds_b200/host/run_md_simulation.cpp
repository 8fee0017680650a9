// run_md_simulation — drop-in host for the reference's runners (code_source/runners/
// run_md_simulation.f90 and run_md_simulation_mpi.f90): same flags, same settings-file grammar,
// same log / xyz outputs; the MD step loop runs on a B200 through libpfmds_b200.so.
// Extra flags: `-gpu d` picks the CUDA device; `-mpi` / `-node r -nodes n` select the ensemble
// (task-farm) mode, where by default rank r uses device (r-1) mod <number of GPUs>.
#include <cstdlib>
#include <cstring>
#include <string>

#include "cabi_engine.hpp"
#include "md_driver.hpp"
#include "runner_cli.hpp"

int main(int argc, char** argv) {
    int device = -1, node = 0;
    for (int i = 1; i + 1 < argc; ++i) {
        if (!std::strcmp(argv[i], "-gpu")) device = std::atoi(argv[i + 1]);
        if (!std::strcmp(argv[i], "-node")) node = std::atoi(argv[i + 1]);
    }
    if (device < 0) {
        const char* lr = std::getenv("LOCAL_RANK");
        device = lr ? std::atoi(lr) : (node > 0 ? node - 1 : 0);
        const char* nd = std::getenv("PFMDS_NUM_GPUS");
        if (nd && std::atoi(nd) > 0) device %= std::atoi(nd);
    }
    return pfmds_host::run_cli(argc, argv, 1, [device](int) { return pfmds_host::CabiEngine(device); });
}
