// TEST SUPPORT ONLY: the SERIAL flavour of the host replay cannot run the slab decomposition (pfmds_b200/csrc/slab.cu uses warp
// scans and spin waits) and links these stand-ins; c->slab is never set there, so none of them runs.  The LOCK-STEP flavour
// compiles slab.cu itself (tests/test_slab_lockstep.py).
#define PFMDS_EMU_LIB 1
#include <string>

#include "../../pfmds_b200/csrc/ctx.hpp"

static void no_slab() { throw std::string("slab decomposition is not available in the host replay"); }
int slab_unique_id(char*) { return 20; }
void slab_init(pfmds_ctx*, int, int, const char*, long long, int, int) { no_slab(); }
void slab_destroy(pfmds_ctx*) {}
void slab_redistribute(pfmds_ctx*) { no_slab(); }
void slab_after_reorder(pfmds_ctx*) { no_slab(); }
void slab_exchange(pfmds_ctx*, int) { no_slab(); }
void slab_step_done(pfmds_ctx*) { no_slab(); }
bool slab_uses_p2p(pfmds_ctx*) { return false; }
bool slab_fused(pfmds_ctx*) { return false; }
SlabDev slab_dev(pfmds_ctx*, int) { return SlabDev{}; }
bool slab_pos_pushed_by_kick(pfmds_ctx*, bool) { return false; }
void slab_allreduce_sum(pfmds_ctx*, double*, int) { no_slab(); }
void slab_allreduce_max(pfmds_ctx*, double*, int) { no_slab(); }
void slab_allreduce_max_int(pfmds_ctx*, int*, int) { no_slab(); }
void slab_allreduce_sum_ll(pfmds_ctx*, unsigned long long*, int) { no_slab(); }
int slab_rank(pfmds_ctx*) { return 0; }
int slab_nranks(pfmds_ctx*) { return 1; }
int slab_n_local(pfmds_ctx*) { return 0; }
long long slab_n_global(pfmds_ctx*) { return 0; }
bool slab_ke_close(pfmds_ctx*, const NhcPack&, int, const double*, double, double, double) { return false; }
void slab_download(pfmds_ctx*, int*, int*, double*, double*, double*) { no_slab(); }
void slab_upload(pfmds_ctx*, int, const double*, const double*) { no_slab(); }
