// run_gr_moire_fitting — drop-in for code_source/runners/run_gr_moire_fitting.f90: golden-section fit of the ljc / morsec
// graphene-metal parameters; every error evaluation relaxes two cells with md(), whose step loop runs on a B200 through
// libpfmds_b200.so.  Extra flags: `-gpu d` (CUDA device), `-pair` (the two cells of an evaluation run concurrently).
#include <cstdlib>
#include <cstring>

#include "cabi_engine.hpp"
#include "fit_gr_moire.hpp"

int main(int argc, char** argv) {
    int device = 0;
    for (int i = 1; i + 1 < argc; ++i)
        if (!std::strcmp(argv[i], "-gpu")) device = std::atoi(argv[i + 1]);
    return pfmds_host::run_gr_moire_fitting_cli(argc, argv, 1, [device](int) { return pfmds_host::CabiEngine(device); });
}
