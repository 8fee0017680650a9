// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).  PARITY UNPINNED.
// C entry points with the same shapes as include/pfmds_b200.h (prefix oracle_ instead of pfmds_), so the
// parity tests drive the CPU restatement and the CUDA library through one ctypes wrapper.
#include <omp.h>

#include <cstring>
#include <string>

#include "oracle_engine.hpp"

using namespace oracle;

double OracleEngine::omp_wtime() { return omp_get_wtime(); }

struct oracle_ctx {
    OracleEngine eng;
    std::string err;
};

#define GUARD(...)                                                           \
    try { __VA_ARGS__; return 0; }                                                  \
    catch (const StopError& e) { c->err = e.what(); return 10; }             \
    catch (const std::exception& e) { c->err = e.what(); return 1; }

extern "C" {

int oracle_create(oracle_ctx** out, int /*device*/, int n, const double* pos, const double* vel, const double* mass, const double* box) {
    oracle_ctx* c = new oracle_ctx;
    *out = c;
    GUARD(c->eng.create(n, pos, vel, mass, box));
}
int oracle_set_group(oracle_ctx* c, int g, int n, const int* idx) { GUARD(c->eng.set_group(g, std::vector<int>(idx, idx + n))); }
int oracle_set_roles(oracle_ctx* c, int am, int xyz, int z, int all) { GUARD(c->eng.set_roles(am, xyz, z, all)); }
int oracle_add_nhc(oracle_ctx* c, int g, double T, int M, double q1) {
    GUARD(pfmds_host::NhcSpec s; s.group = g; s.temperature = T; s.M = M; s.q1 = q1; c->eng.add_nhc(s));
}
int oracle_add_group_change(oracle_ctx* c, int from, int to, int ts1, int ts2, int frec) { GUARD(c->eng.add_group_change(from, to, ts1, ts2, frec)); }
int oracle_group_size(oracle_ctx* c, int g, int* n) { GUARD(*n = c->eng.group_size(g)); }
int oracle_set_misc(oracle_ctx* c, int zmp, int inv) { GUARD(c->eng.set_misc(zmp, inv != 0)); }
int oracle_add_interaction(oracle_ctx* c, const char* name, int np, const double* params, int nl_n, const int* gn, const int* maxn,
                           const double* rcut, const int* period) {
    GUARD(
        pfmds_host::InteractionSpec s; s.name = name; s.nl_n = nl_n; s.params.assign(params, params + np);
        for (int j = 0; j < nl_n; ++j) {
            pfmds_host::ListSpec l; l.g1 = gn[2 * j]; l.g2 = gn[2 * j + 1]; l.neighb_num_max = maxn[j]; l.r_cut = rcut[j]; l.update_period = period[j];
            s.lists.push_back(l);
        }
        c->eng.add_interaction(s));
}
int oracle_advance(oracle_ctx* c, int kind, double dt, int first, int n) { GUARD(c->eng.advance(kind, dt, first, n)); }
// checker twin of pfmds_advance_logged: the reference's sequence, one step and one energy evaluation at a time (md_simulation.f90:138-199)
int oracle_advance_logged(oracle_ctx* c, int kind, double dt, int first, int n, int log_period, double* rows, int row_len, int* n_rows) {
    GUARD(
        int r = 0;
        for (int s = first; s < first + n; ++s) {
            c->eng.advance(kind, dt, s, 1);
            if (log_period < 1 || s % log_period != 0) continue;
            std::vector<double> ei, en;
            double ke = 0, temp = 0;
            c->eng.energies(ei, ke, temp, en);
            if ((int)(ei.size() + 2 + en.size()) > row_len) throw std::runtime_error("error: log rows too short");
            double* out = rows + (size_t)r * row_len;
            std::copy(ei.begin(), ei.end(), out);
            out[ei.size()] = ke; out[ei.size() + 1] = temp;
            std::copy(en.begin(), en.end(), out + ei.size() + 2);
            ++r;
        }
        if (n_rows) *n_rows = r);
}
int oracle_energies(oracle_ctx* c, double* e_inter, double* ke, double* temp, double* e_nhc) {
    GUARD(
        std::vector<double> ei, en; c->eng.energies(ei, *ke, *temp, en);
        if (e_inter) std::copy(ei.begin(), ei.end(), e_inter);
        if (e_nhc) std::copy(en.begin(), en.end(), e_nhc));
}
int oracle_diagnostics(oracle_ctx* c, double* fs, double* mc, double* mcv, double* vmax, int* nl_load) {
    GUARD(std::vector<int> l; c->eng.diagnostics(fs, mc, mcv, *vmax, l); if (nl_load) std::copy(l.begin(), l.end(), nl_load));
}
int oracle_download(oracle_ctx* c, double* pos, double* vel, double* frc) { GUARD(c->eng.download(pos, vel, frc)); }
int oracle_neighbours(oracle_ctx* c, int inter, int list, int* nlist, int* nnum, int* lessnnum) {
    GUARD(
        const NeighbourList& nl = c->eng.sys.interactions.at((size_t)inter).nl.at((size_t)list);
        for (int i = 0; i < nl.N; ++i) {
            if (nnum) nnum[i] = nl.nnum[i];
            if (lessnnum) lessnnum[i] = nl.lessnnum[i];
            if (nlist) for (int p = 0; p < nl.neighb_num_max; ++p)
                nlist[(size_t)i * nl.neighb_num_max + p] = p < nl.nnum[i] ? nl.nlist[(size_t)i * nl.neighb_num_max + p] + 1 : 0;
        });
}
int oracle_normals(oracle_ctx* c, int inter, double* out) {
    GUARD(
        const Interaction& it = c->eng.sys.interactions.at((size_t)inter);
        const std::vector<double>& g = it.interaction_name == "ljc" ? it.ljc.gr_norm : it.morsec.gr_norm;
        std::copy(g.begin(), g.end(), out));
}
int oracle_get_nhc(oracle_ctx* c, int k, double* x, double* v) {
    GUARD(const NoseHooverChain& n = c->eng.sys.nhc.at((size_t)k); std::copy(n.x.begin(), n.x.end(), x); std::copy(n.v.begin(), n.v.end(), v));
}
int oracle_set_nhc(oracle_ctx* c, int k, const double* x, const double* v) {
    GUARD(NoseHooverChain& n = c->eng.sys.nhc.at((size_t)k); std::copy(x, x + n.M, n.x.begin()); std::copy(v, v + n.M, n.v.begin()));
}
int oracle_timers(oracle_ctx* c, double* t) { GUARD(c->eng.timers(t)); }
int oracle_launch_count(oracle_ctx*, long long* n) { *n = 0; return 0; }
int oracle_synchronize(oracle_ctx*) { return 0; }
const char* oracle_last_error(oracle_ctx* c) { return c ? c->err.c_str() : ""; }
int oracle_destroy(oracle_ctx* c) { delete c; return 0; }
const char* oracle_version(void) { return "pfmds oracle (C++ restatement, parity unpinned)"; }
int oracle_set_threads(int n) { omp_set_num_threads(n); return omp_get_max_threads(); }

}  // extern "C"
