// run_gr_analysis — drop-in for code_source/runners/run_gr_analysis.f90: z-height statistics of graphene over a metal surface
// for a list of xyz files (pure host post-processing, no device work).
#include "fit_gr_moire.hpp"

int main(int argc, char** argv) { return pfmds_host::run_gr_analysis_cli(argc, argv); }
