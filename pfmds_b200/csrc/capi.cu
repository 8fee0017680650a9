// pfmds_b200 — the C ABI of include/pfmds_b200.h and the per-step orchestration of the device
// kernels (the body of `do md_step` in code_source/MOLECULAR_DYNAMICS/md_simulation.f90:138-186).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pfmds_b200.h"
#include "ctx.hpp"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw std::string("CUDA: ") + cudaGetErrorString(e_) + " at " #x; } while (0)

static bool rjl_short_forms_ok(int dev);  // defined next to the device self-tests below

// Live contexts of this process per device.  A context that shares its GPU with others (ensemble mode: many small runs, one
// stream each) leaves the parallel branches to them: measured on a B200, 64 graphene-on-Cu replicas reach 2.8e8 atom-steps/s with
// one stream per context and 1.2e8 when every context forks its interactions onto streams of its own.
#include <atomic>
static std::atomic<int> g_live_contexts[64];

namespace {

struct Fail { int code; std::string msg; };
[[noreturn]] void fail(int code, const std::string& m) { throw Fail{code, m}; }

template <class Fn>
int guarded(pfmds_ctx* c, Fn fn) {
    try { fn(); return PFMDS_OK; }
    catch (const Fail& f) { if (c) c->err_msg = f.msg; return f.code; }
    catch (const std::string& s) { if (c) c->err_msg = s; return PFMDS_ERR_CUDA; }
    catch (const std::exception& e) { if (c) c->err_msg = e.what(); return PFMDS_ERR_INVALID; }
}

int kind_of(const std::string& n) {
    if (n == "lj") return K_LJ;
    if (n == "lj1g") return K_LJ1G;
    if (n == "ljc") return K_LJC;
    if (n == "morsec") return K_MORSEC;
    if (n == "tb") return K_TB;
    if (n == "rjl") return K_RJL;
    if (n == "rebosc") return K_REBOSC;
    return -1;
}
int nl_n_of(int kind) { return kind == K_LJ ? 2 : (kind == K_LJC || kind == K_MORSEC) ? 3 : 1; }

void finalize_slab(pfmds_ctx* c);

// Kernel-variant and path switches from the environment, read once per context (pfmds_create / pfmds_create_slab).
int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    if (!v || !*v) return dflt;
    char* end = nullptr;
    long x = std::strtol(v, &end, 10);
    return (end && *end == 0 && x >= 0 && x <= 2000000000L) ? (int)x : dflt;
}
void read_env(pfmds_ctx* c, int n_atoms, bool slab) {
    c->timers_on = env_int("PFMDS_TIMERS", 0) == 1;
    c->use_graphs = !slab && env_int("PFMDS_GRAPHS", n_atoms < 200000 ? 1 : 0) == 1;
    c->lj1g_pipe = env_int("PFMDS_LJ1G_PIPE", 1) != 0;
    c->nl_mask = env_int("PFMDS_NL_MASK", 1) != 0;
    c->nl_cell = env_int("PFMDS_NL_CELL", 1) != 0;
    c->pre_open_enabled = env_int("PFMDS_PRE_OPEN", 1) != 0;
    c->graph_rebuilds = env_int("PFMDS_GRAPH_REBUILDS", 1) != 0;
    c->graph_steps = env_int("PFMDS_GRAPH_STEPS", 4);
    if (c->graph_steps < 1 || c->graph_steps > 64) c->graph_steps = 1;
    // 2 is the default: the third generation (node-table exponentials) removes 13 of 43 FP64 instructions per pair but its table
    // look-ups double the L1 data-pipe wavefronts, and that pipe is what bounds these kernels (ncu, profiles/r2b_*): measured
    // 0.343 against 0.277 ms (density) and 0.362 against 0.360 ms (force) per launch at 10^6 atoms
    c->rjl_gen = env_int("PFMDS_RJL_GEN", 2);
    if (c->rjl_gen < 1 || c->rjl_gen > 3) c->rjl_gen = 2;
#ifdef PFMDS_COOP
    c->small_n = env_int("PFMDS_SMALL_N", 100000);
    c->nl_warp_n = env_int("PFMDS_NL_WARP_N", 200000);
#else  // serial host replay of the test suite: no lane exchange, thread-per-atom kernels only
    c->small_n = 0;
    c->nl_warp_n = 0;
#endif
}

const std::vector<int>& group_of(pfmds_ctx* c, int g) {
    if (g < 1 || g > (int)c->groups.size()) fail(PFMDS_ERR_INVALID, "error: group number " + std::to_string(g) + " is not defined");
    return c->groups[(size_t)g - 1];
}

long long group_size(pfmds_ctx* c, int g) {
    if (c->slab) {
        if (g < 1 || g > (int)c->group_count.size()) fail(PFMDS_ERR_INVALID, "error: group number " + std::to_string(g) + " is not defined");
        return c->group_count[(size_t)g - 1];
    }
    group_of(c, g);
    if (c->cur_n.size() == c->groups.size()) return c->cur_n[(size_t)g - 1];  // group%N (deposition changes it)
    return (long long)group_of(c, g).size();
}

// poll the device error word; turns the first recorded condition into the reference's message
void check_device_error(pfmds_ctx* c) {
    int h[PFMDS_ERRW];
    if (c->slab) slab_allreduce_max_int(c, c->err, 1);  // every rank stops together
    CK(cudaMemcpyAsync(h, c->err, sizeof h, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    if (h[0] == 0) return;
    if (h[0] == 31) fail(PFMDS_ERR_CUDA, "slab decomposition: a neighbour rank did not signal within the flag time-out (peer-memory halo; PFMDS_SLAB_TIMEOUT_S, default 120 s): it failed, or its host thread fell that far behind");
    if (h[0] == 30) fail(PFMDS_ERR_UNSUPPORTED, "slab decomposition: an atom moved farther than one slab between two list rebuilds");
    if (h[0] == E_OUT_OF_CELL) fail(PFMDS_ERR_OUT_OF_CELL, " " + std::to_string(h[1] + 1) + "  particle out of cell");
    if (h[0] == E_TOO_MANY) fail(PFMDS_ERR_TOO_MANY_NEIGHBOURS, "error: too many neighbours (atom " + std::to_string(h[1] + 1) + ", " + std::to_string(h[2]) + " found)");
    if (h[0] == E_GR_NEIB)
        fail(PFMDS_ERR_GR_NEIGHBOURS, std::string(h[2] > 3 ? "error: too many gr nearest neibs" : "error: not enough gr nearest neibs") + " (atom " +
                                          std::to_string(h[1] + 1) + ", " + std::to_string(h[2]) + " found)");
    fail(PFMDS_ERR_CUDA, "device error " + std::to_string(h[0]));
}


// The second-generation rjl kernels lean on properties of this device's MUFU.RSQ64H seed (relative error, zero low word) and on
// bit-level exponent arithmetic (mathx.cuh short forms).  The first rjl context of a process has the device check them itself;
// a device that misses the bounds gets the first generation, and says so.
void check_rjl_generation(pfmds_ctx* c) {
    if (c->rjl_gen == 1) return;
    bool has_rjl = false;
    for (auto& it : c->inter) has_rjl |= it.kind == K_RJL;
    if (!has_rjl || rjl_short_forms_ok(c->dev)) return;
    std::fprintf(stderr, "pfmds_b200: the short elementary functions of the second-generation rjl kernels missed their error bounds on device %d; "
                         "using the first generation (PFMDS_RJL_GEN=1)\n", c->dev);
    c->rjl_gen = 1;
}

// device-resident energy log of pfmds_advance_logged: allocated once with the description (no first-use allocation inside a call)
#define LOG_ROWS_PREALLOC 4096
void alloc_log(pfmds_ctx* c) {
    size_t D = c->inter.size() + 1;
    for (auto& t : c->nhc) D += (size_t)3 * t.M;
    if (c->logbuf) return;
    CK(cudaMalloc(&c->logbuf, sizeof(double) * D * LOG_ROWS_PREALLOC));
    c->log_cap = D * LOG_ROWS_PREALLOC;
}

// Slab mode: the masks came with pfmds_create_slab; validation works on group numbers and global sizes.
void finalize_slab(pfmds_ctx* c) {
    if (!c->changes.empty()) fail(PFMDS_ERR_UNSUPPORTED, "unsupported: group changes (deposition) in slab decomposition mode");
    int period = -1;
    for (auto& it : c->inter) {
        if (it.kind != K_LJ && it.kind != K_LJ1G && it.kind != K_RJL)
            fail(PFMDS_ERR_UNSUPPORTED, "unsupported: slab decomposition handles lj, lj1g and rjl (" + it.name + " needs ghost bond orders / normals / second-shell rows)");
        NList& a = it.nl[0];
        for (int j = 0; j < it.nl_n; ++j) { group_size(c, it.nl[j].g1); group_size(c, it.nl[j].g2); }
        if ((it.kind == K_LJ1G || it.kind == K_RJL) && a.g1 != a.g2) fail(PFMDS_ERR_UNSUPPORTED, "unsupported: " + it.name + " needs group1 == group2");
        if (it.kind == K_LJ) {
            NList& b = it.nl[1];
            if (group_size(c, a.g2) > group_size(c, b.g1)) fail(PFMDS_ERR_LIST_SIZE, "error: group2%N>cnl%N");
            b.g1 = a.g2; b.g2 = a.g1; b.rcut = a.rcut; b.period = a.period;
        }
        double R1 = it.kind == K_LJ ? it.lj.R1 : it.kind == K_LJ1G ? it.lj1g.R1 : it.rjl.R1;
        double R2 = it.kind == K_LJ ? it.lj.R2 : it.kind == K_LJ1G ? it.lj1g.R2 : it.rjl.R2;
        for (int j = 0; j < it.nl_n; ++j) {
            NList& l = it.nl[j];
            if (l.maxn < 1 || l.period < 1 || !(l.rcut > 0)) fail(PFMDS_ERR_INVALID, "error: bad neighbour list parameters");
            if (period < 0) period = l.period;
            if (l.period != period) fail(PFMDS_ERR_UNSUPPORTED, "unsupported: slab decomposition needs one update_period for all lists");
            l.partition = true; l.part_r1sq = R1 * R1; l.part_r2sq = R2 * R2;
            CK(cudaMalloc(&l.nlist_alt, sizeof(int) * (size_t)l.maxn * c->stride));
            // two spare rows, and every entry a valid slot number from the start (0, later stale ones): the pipelined row walks of
            // forces.cu prefetch up to slot n+1 without bounds tests and gather the partner record of whatever they find there
            CK(cudaMalloc(&l.nlist, sizeof(int) * (size_t)(l.maxn + 2) * c->stride));
            CK(cudaMemsetAsync(l.nlist, 0, sizeof(int) * (size_t)(l.maxn + 2) * c->stride, c->st));
            CK(cudaMalloc(&l.nnum, sizeof(int) * c->stride));
            CK(cudaMemsetAsync(l.nnum, 0, sizeof(int) * c->stride, c->st));
        }
    }
    group_size(c, c->all_moving); group_size(c, c->xyz_moving); group_size(c, c->z_moving); group_size(c, c->all_atoms);
    for (auto& t : c->nhc) {
        t.L = (int)group_size(c, t.group);
        if (t.L < 1) fail(PFMDS_ERR_NHC_PARAMS, "error: wrong nhc parameters");
        std::vector<double> st((size_t)3 * t.M + 4, 0.);
        double q1;
        CK(cudaMemcpy(&q1, t.state + 2 * t.M, sizeof(double), cudaMemcpyDeviceToHost));
        st[2 * (size_t)t.M] = q1;
        for (int i = 1; i < t.M; ++i) st[2 * (size_t)t.M + i] = q1 / (3. * t.L);
        st[3 * (size_t)t.M] = 1.;
        st[3 * (size_t)t.M + 2] = 1.;
        CK(cudaMemcpy(t.state, st.data(), sizeof(double) * st.size(), cudaMemcpyHostToDevice));
    }
    c->nhc_fusable = !c->nhc.empty() && c->nhc.size() <= NHC_MAXF;  // distinct groups are assumed disjoint: checked below
    for (size_t a = 0; a < c->nhc.size(); ++a)
        for (size_t b = a + 1; b < c->nhc.size(); ++b)
            if (c->nhc[a].group == c->nhc[b].group) c->nhc_fusable = false;
    if (c->nhc.size() > 1) c->nhc_fusable = false;  // several thermostats: masks are not on the host in slab mode, keep the plain path
    if (!c->inter.empty()) CK(cudaMalloc(&c->energy, sizeof(double) * c->inter.size()));
    alloc_log(c);
    check_rjl_generation(c);
    for (auto& it : c->inter) rjl_prepare(c, it);
    c->first_overwrites = !c->inter.empty() && c->inter[0].kind == K_RJL && group_size(c, c->inter[0].nl[0].g1) == slab_n_global(c);
    nl_setup_grid(c);
    CK(cudaStreamSynchronize(c->st));
    c->finalized = true;
}

// Validate the description, build the masks, allocate the lists.  Runs once, at the first advance.
void finalize(pfmds_ctx* c) {
    if (c->finalized) return;
    if (c->slab) { finalize_slab(c); return; }
    const int N = c->N;
    if (c->groups.size() > PFMDS_MAX_GROUPS) fail(PFMDS_ERR_UNSUPPORTED, "unsupported: more than 32 atom groups");
    c->h_gmask.assign((size_t)N, 0u);
    for (size_t g = 0; g < c->groups.size(); ++g)
        for (int i1 : c->groups[g]) {
            if (i1 < 1 || i1 > N) fail(PFMDS_ERR_INVALID, "error: atom index out of range in group " + std::to_string(g + 1));
            uint32_t b = 1u << g;
            if (c->h_gmask[(size_t)i1 - 1] & b)
                fail(PFMDS_ERR_UNSUPPORTED, "unsupported: atom " + std::to_string(i1) + " listed twice in group " + std::to_string(g + 1));
            c->h_gmask[(size_t)i1 - 1] |= b;
        }
    group_of(c, c->all_moving); group_of(c, c->xyz_moving); group_of(c, c->z_moving); group_of(c, c->all_atoms);
    auto monotone = [&](int g) { const auto& v = group_of(c, g); return std::is_sorted(v.begin(), v.end()); };
    int first_tb = -1;
    for (size_t k = 0; k < c->inter.size(); ++k)
        if (c->inter[k].kind == K_TB || c->inter[k].kind == K_REBOSC) { first_tb = (int)k; break; }  // md_interactions.f90:157-162
    for (size_t k = 0; k < c->inter.size(); ++k) {
        Inter& it = c->inter[k];
        for (int j = 0; j < it.nl_n; ++j) { group_of(c, it.nl[j].g1); group_of(c, it.nl[j].g2); }
        NList& a = it.nl[0];
        if (it.kind == K_LJ1G || it.kind == K_TB || it.kind == K_RJL || it.kind == K_REBOSC) {
            // the reference indexes group-1 rows with group-2 local numbers in these potentials
            // (LennardJones_1g.f90:83, TersoffBrenner.f90:56-59, RosatoGuillopeLegrand.f90:88)
            if (group_of(c, a.g1) != group_of(c, a.g2)) fail(PFMDS_ERR_UNSUPPORTED, "unsupported: " + it.name + " needs group1 == group2");
        }
        if (it.kind == K_LJ1G && !monotone(a.g1))
            fail(PFMDS_ERR_UNSUPPORTED, "unsupported: lj1g group indexes are not ascending; the reference's half-list rule (lessnnum, "
                                        "md_neighbours.f90:78) silently drops pairs for such groups");
        if (it.kind == K_LJ || it.kind == K_LJC || it.kind == K_MORSEC) {
            // nl(2) is the converse of nl(1): rows = group2 of line 1, r_cut and period of line 1, capacity of line 2
            NList& b = it.nl[1];
            if (group_of(c, a.g2).size() > group_of(c, b.g1).size()) fail(PFMDS_ERR_LIST_SIZE, "error: group2%N>cnl%N");
            b.g1 = a.g2; b.g2 = a.g1; b.rcut = a.rcut; b.period = a.period;
        }
        if (it.kind == K_LJC || it.kind == K_MORSEC) {
            NList& n3 = it.nl[2];
            if (first_tb >= 0) {
                if (first_tb > (int)k)
                    fail(PFMDS_ERR_UNSUPPORTED, "unsupported: 'tb' must be listed before '" + it.name + "' (the nearest-neighbour list is taken from it, md_interactions.f90:157-167)");
                if (n3.maxn != 3) fail(PFMDS_ERR_LIST_SIZE, "error: nl_nn%neighb_num_max/=nnum_nn");
                if (group_of(c, c->inter[first_tb].nl[0].g1) != group_of(c, a.g1)) fail(PFMDS_ERR_LIST_SIZE, "error: nl%N/=nl_nn%N");
                n3.from_tb = true; n3.src_inter = first_tb;
                n3.g1 = n3.g2 = a.g1;
                n3.period = c->inter[first_tb].nl[0].period;
            } else {
                if (n3.maxn != 3) fail(PFMDS_ERR_UNSUPPORTED, "unsupported: the nearest-neighbour list of " + it.name + " must have capacity 3");
                if (group_of(c, n3.g1) != group_of(c, a.g1)) fail(PFMDS_ERR_LIST_SIZE, "error: nl%N/=nl_nn%N");
            }
        }
        {   // switch radii of the owning potential drive the row partition (nl.cu k_partition)
            double R1 = 0, R2 = 0;
            switch (it.kind) {
            case K_LJ: R1 = it.lj.R1; R2 = it.lj.R2; break;
            case K_LJ1G: R1 = it.lj1g.R1; R2 = it.lj1g.R2; break;
            case K_LJC: R1 = it.ljc.R1; R2 = it.ljc.R2; break;
            case K_MORSEC: R1 = it.mor.R1; R2 = it.mor.R2; break;
            case K_RJL: R1 = it.rjl.R1; R2 = it.rjl.R2; break;
            default: break;
            }
            int np = (it.kind == K_TB || it.kind == K_REBOSC) ? 0 : (it.kind == K_LJ1G || it.kind == K_RJL ? 1 : 2);
            for (int j = 0; j < np; ++j) { it.nl[j].partition = true; it.nl[j].part_r1sq = R1 * R1; it.nl[j].part_r2sq = R2 * R2; }
        }
        for (int j = 0; j < it.nl_n; ++j) {
            NList& l = it.nl[j];
            if (l.maxn < 1 || l.period < 1 || !(l.rcut > 0)) fail(PFMDS_ERR_INVALID, "error: bad neighbour list parameters");
            if (l.partition) CK(cudaMalloc(&l.nlist_alt, sizeof(int) * (size_t)l.maxn * c->stride));
            // two spare rows, and every entry a valid slot number from the start (0, later stale ones): the pipelined row walks of
            // forces.cu prefetch up to slot n+1 without bounds tests and gather the partner record of whatever they find there
            CK(cudaMalloc(&l.nlist, sizeof(int) * (size_t)(l.maxn + 2) * c->stride));
            CK(cudaMemsetAsync(l.nlist, 0, sizeof(int) * (size_t)(l.maxn + 2) * c->stride, c->st));
            CK(cudaMalloc(&l.nnum, sizeof(int) * c->stride));
            CK(cudaMemsetAsync(l.nnum, 0, sizeof(int) * c->stride, c->st));
        }
        if (it.kind == K_TB) {
            size_t need = ((size_t)N + 127) / 128 * (size_t)a.maxn + 16;  // energy partials: one per (block, slot)
            if (need > c->part_cap) { cudaFree(c->part); CK(cudaMalloc(&c->part, sizeof(double) * need)); c->part_cap = need; }
            CK(cudaMalloc(&it.aux, sizeof(double) * (size_t)a.maxn * c->stride));
            CK(cudaMalloc(&it.aux2, sizeof(double) * (size_t)a.maxn * c->stride));
            CK(cudaMalloc(&it.fpart, sizeof(double4) * (size_t)a.maxn * c->stride));
        }
        if (it.kind == K_LJC || it.kind == K_MORSEC) {
            CK(cudaMalloc(&it.gnorm, sizeof(double4) * c->stride));
            CK(cudaMalloc(&it.tvec, sizeof(double4) * c->stride));
            CK(cudaMemsetAsync(it.gnorm, 0, sizeof(double4) * c->stride, c->st));
            CK(cudaMemsetAsync(it.tvec, 0, sizeof(double4) * c->stride, c->st));
        }
    }
    for (auto& t : c->nhc) {
        t.L = (int)group_of(c, t.group).size();
        if (t.L < 1) fail(PFMDS_ERR_NHC_PARAMS, "error: wrong nhc parameters");
        // q(i) = q(1)/(3L), i >= 2  (md_integrators.f90:192-195)
        std::vector<double> st((size_t)3 * t.M + 4, 0.);
        double q1;
        CK(cudaMemcpy(&q1, t.state + 2 * t.M, sizeof(double), cudaMemcpyDeviceToHost));
        st[2 * (size_t)t.M] = q1;
        for (int i = 1; i < t.M; ++i) st[2 * (size_t)t.M + i] = q1 / (3. * t.L);
        st[3 * (size_t)t.M] = 1.;
        st[3 * (size_t)t.M + 2] = 1.;
        CK(cudaMemcpy(t.state, st.data(), sizeof(double) * st.size(), cudaMemcpyHostToDevice));
    }
    {   // fused NVT path needs pairwise disjoint thermostat groups
        uint32_t seen = 0;
        bool ok = !c->nhc.empty() && c->nhc.size() <= NHC_MAXF;
        std::vector<uint32_t> cover((size_t)N, 0u);
        for (size_t k = 0; ok && k < c->nhc.size(); ++k) {
            uint32_t b = 1u << (c->nhc[k].group - 1);
            if (seen & b) ok = false;
            seen |= b;
        }
        if (ok)
            for (size_t i = 0; i < (size_t)N && ok; ++i) {
                int cnt = 0;
                for (auto& t : c->nhc) cnt += (c->h_gmask[i] >> (t.group - 1)) & 1u;
                if (cnt > 1) ok = false;
            }
        c->nhc_fusable = ok;
    }
    CK(cudaMemcpyAsync(c->gmask, c->h_gmask.data(), sizeof(uint32_t) * (size_t)N, cudaMemcpyHostToDevice, c->st));
    if (!c->inter.empty()) CK(cudaMalloc(&c->energy, sizeof(double) * c->inter.size()));
    // group%N starts at the full size (create_particle_group); change entries take over from the first step on
    c->cur_n.resize(c->groups.size());
    for (size_t g = 0; g < c->groups.size(); ++g) c->cur_n[g] = (int)c->groups[g].size();
    c->d_grank.assign(c->groups.size(), nullptr);
    c->zero_all = (int)group_of(c, c->all_atoms).size() == N;
    for (auto& ch : c->changes) {
        group_of(c, ch.from);
        const auto& G = group_of(c, ch.to);
        if (ch.to == c->all_atoms) c->zero_all = false;
        if (c->d_grank[(size_t)ch.to - 1]) continue;
        std::vector<int> rank((size_t)N, 0x7fffffff);
        for (size_t r = 0; r < G.size(); ++r) rank[(size_t)G[r] - 1] = (int)r;
        CK(cudaMalloc(&c->d_grank[(size_t)ch.to - 1], sizeof(int) * (size_t)N));
        CK(cudaMemcpyAsync(c->d_grank[(size_t)ch.to - 1], rank.data(), sizeof(int) * (size_t)N, cudaMemcpyHostToDevice, c->st));
    }
    alloc_log(c);
    check_rjl_generation(c);
    for (auto& it : c->inter) rjl_prepare(c, it);
    {   // per-interaction force buffers of small systems (compute_forces); lj gets two: its two lists run as separate branches
        size_t nbuf = 0;
        for (auto& it : c->inter) nbuf += (it.kind == K_LJ || it.kind == K_LJC || it.kind == K_MORSEC) ? 2 : 1;
        // (decided once: a context finalized while it shares its device with other contexts of the process -- ensemble mode -- keeps the
        //  one-stream, one-force-array path for good: the sum kernel would only cost it time, 1.94e9 against 2.17e9 atom-steps/s for 64 runs on 8 GPUs)
        if (c->N < c->small_n && c->inter.size() >= 2 && nbuf <= 12 && env_int("PFMDS_SMALL_FORK", 1) != 0 && g_live_contexts[c->dev & 63].load() <= 1) {
            for (size_t k = 0; k < nbuf; ++k) {
                double4* b = nullptr;
                CK(cudaMalloc(&b, sizeof(double4) * c->stride));
                CK(cudaMemset(b, 0, sizeof(double4) * c->stride));
                c->fbuf.push_back(b);
                cudaStream_t st = nullptr;
                CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
                c->aux_st.push_back(st);
            }
            for (size_t k = 0; k < nbuf + 1; ++k) {
                cudaEvent_t e = nullptr;
                CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                c->aux_ev.push_back(e);
            }
            CK(cudaEventCreateWithFlags(&c->aux_ev_mid, cudaEventDisableTiming));
            CK(cudaMalloc(&c->ticket, sizeof(unsigned int)));
            CK(cudaMemset(c->ticket, 0, sizeof(unsigned int)));
            c->fbuf_on = true;
        }
    }
    c->first_overwrites = c->zero_all && c->changes.empty() && !c->inter.empty() && (c->inter[0].kind == K_RJL || (c->inter[0].kind == K_LJ1G && c->lj1g_pipe)) &&
                          (int)group_of(c, c->inter[0].nl[0].g1).size() == N;
    nl_setup_grid(c);
    CK(cudaStreamSynchronize(c->st));
    c->finalized = true;
}

struct PhaseTimer {
    pfmds_ctx* c; int slot;
    PhaseTimer(pfmds_ctx* c_, int s) : c(c_), slot(s) { if (c->timers_on) cudaEventRecord(c->ev0, c->st); }
    ~PhaseTimer() {
        if (!c->timers_on) return;
        cudaEventRecord(c->ev1, c->st);
        cudaEventSynchronize(c->ev1);
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        c->t_phase[slot] += ms * 1e-3;
    }
};

// change_particle_group_N, md_general.f90:82-94: membership bit of group `bit` = (position in the group's index list < n)
__global__ void k_group_resize(int N, const int* __restrict__ orig, const int* __restrict__ rank, uint32_t bit, int n, uint32_t* __restrict__ gmask) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    uint32_t g = gmask[i];
    gmask[i] = rank[orig[i]] < n ? (g | bit) : (g & ~bit);
}
// The `do i=1,change_group_num` loop at the top of every md step (md_simulation.f90:116-119).  Returns true when a group changed.
bool apply_group_changes(pfmds_ctx* c, int step) {
    if (c->changes.empty()) return false;
    std::vector<int> before = c->cur_n;
    for (const auto& ch : c->changes) {
        int& n = c->cur_n[(size_t)ch.to - 1];
        if (step <= ch.ts1) {
            n = c->cur_n[(size_t)ch.from - 1];
            if (step == ch.ts1) n = n + 1;
        } else if (step < ch.ts2) {
            if (ch.frec == 0) fail(PFMDS_ERR_INVALID, "error: change_frec is zero (integer division by zero in mod)");
            if ((step - ch.ts1) % ch.frec == 0) n = n + 1;
        }
        const int cap = (int)c->groups[(size_t)ch.to - 1].size();
        if (n > cap) n = cap;
    }
    bool changed = false;
    for (size_t g = 0; g < c->cur_n.size(); ++g) {
        if (c->cur_n[g] == before[g]) continue;
        if (!changed) {
            // velocity scalings still pending from the last thermostat step belong to the OLD membership
            integ_flush_pending(c);
            c->nhc_ke_valid = false;
            c->energy_valid = false;
        }
        changed = true;
        LAUNCH((k_group_resize), (c->N + 255) / 256, 256, c->st, c->N, c->orig, c->d_grank[g], 1u << g, c->cur_n[g], c->gmask);
        c->launches += 1;
    }
    return changed;
}

// update_interactions_neighbour_lists, md_interactions.f90:138-178
void update_lists(pfmds_ctx* c, int step) {
    bool any = false, all = true;
    for (auto& it : c->inter) {
        for (int j = 0; j < it.nl_n; ++j) {
            NList& l = it.nl[j];
            bool rb = (step % l.period == 0) || !l.built;
            any |= rb;
            all &= rb;
        }
    }
    if (c->slab && !any) slab_exchange(c, 0);  // ghost positions follow their owners every step
    if (any) {
        PhaseTimer t(c, 2);
        if (c->slab) slab_redistribute(c);   // finalize_slab guarantees any == all
        nl_bin_atoms(c, all);
        if (c->slab) slab_after_reorder(c);
        for (auto& it : c->inter) {
            for (int j = 0; j < it.nl_n; ++j) {
                NList& l = it.nl[j];
                if (!((step % l.period == 0) || !l.built)) continue;
                if (l.from_tb) nl_nearest3_from(c, l, c->inter[(size_t)l.src_inter].nl[0]);
                else nl_build(c, l);
            }
        }
    }
    if (!c->fbuf_on)   // (small systems: the normals are the first kernel of the ljc / morsec branch of compute_forces)
        for (size_t k = 0; k < c->inter.size(); ++k) normals_interaction(c, (int)k);  // update_norm_in_graphene, every step
}

// zero_forces + calculate_forces + calculate_forces_numerically, md_simulation.f90:163-165
// `defer_sum`: the caller follows with integ_sum_forces(mode 1 / 2), which adds the closing kick to the sum of the buffers
void compute_forces(pfmds_ctx* c, bool with_energy, bool defer_sum = false) {
    // Small systems, steps that do not report energies: one branch per interaction (own stream, own force buffer): in the step's CUDA
    // graph the interactions become parallel branches, so its depth is the longest kernel chain of ONE interaction, not their sum.
    // The buffers are added in file order, analytic interactions first and calculate_forces_numerically (rebosc) after them, as below.
    // Steps that report energies (or are being profiled) walk the same buffers one interaction after the other on the context's
    // stream: the per-atom additions are the same in both cases, so a run's bits do not depend on its logging cadence.
    if (c->fbuf_on) {
        const bool fork = !with_energy && !c->prof_on && !c->timers_on && g_live_contexts[c->dev & 63].load() <= 1;
        std::vector<size_t> order;
        for (size_t k = 0; k < c->inter.size(); ++k) if (c->inter[k].kind != K_REBOSC) order.push_back(k);
        for (size_t k = 0; k < c->inter.size(); ++k) if (c->inter[k].kind == K_REBOSC) order.push_back(k);
        c->fbuf_active = true;
        if (fork) CK(cudaEventRecord(c->aux_ev[0], c->st));
        size_t b = 0;   // next buffer, in summation order
        for (size_t q = 0; q < order.size(); ++q) {
            const int kq = c->inter[order[q]].kind;
            const bool two = kq == K_LJ || kq == K_LJC || kq == K_MORSEC;   // second list of lj, metal side of ljc / morsec: own branch and buffer
            cudaStream_t s1 = fork ? c->aux_st[b] : nullptr, s2 = (fork && two) ? c->aux_st[b + 1] : nullptr;
            if (fork) { CK(cudaStreamWaitEvent(s1, c->aux_ev[0], 0)); if (two) CK(cudaStreamWaitEvent(s2, c->aux_ev[0], 0)); }
            c->fst = s1; c->fout = c->fbuf[b];
            c->fst2 = two ? s2 : nullptr; c->fout2 = two ? c->fbuf[b + 1] : nullptr;
            try { forces_interaction(c, (int)order[q], with_energy); } catch (...) { c->fst = c->fst2 = nullptr; c->fout = c->fout2 = nullptr; throw; }
            if (fork) { CK(cudaEventRecord(c->aux_ev[1 + b], s1)); if (two) CK(cudaEventRecord(c->aux_ev[2 + b], s2)); }
            b += two ? 2 : 1;
        }
        c->fst = c->fst2 = nullptr; c->fout = c->fout2 = nullptr;
        if (fork)
            for (size_t q = 0; q < b; ++q) CK(cudaStreamWaitEvent(c->st, c->aux_ev[1 + q], 0));
        if (!defer_sum) integ_sum_forces(c, 0, 0.);
        c->energy_valid = with_energy;
        return;
    }
    forces_zero(c);
    for (size_t k = 0; k < c->inter.size(); ++k)   // calculate_forces: the analytic interactions in file order
        if (c->inter[k].kind != K_REBOSC) forces_interaction(c, (int)k, with_energy);
    for (size_t k = 0; k < c->inter.size(); ++k)   // calculate_forces_numerically comes after all of them
        if (c->inter[k].kind == K_REBOSC) forces_interaction(c, (int)k, with_energy);
    if (c->slab) slab_step_done(c);
    c->energy_valid = with_energy;
}

void do_step(pfmds_ctx* c, int step, int kind, double dt, bool first_of_call, bool with_energy = false) {
    c->energy_valid = false;
    {
        PhaseTimer t(c, 0);
        if (first_of_call) integ_check_positions(c);
        if (step == 0) {  // a re-run from md step 0 after NVT steps: the closing scale of the last thermostat step is applied before anything reads velocities
            integ_flush_pending(c);
            c->nhc_ke_valid = false;
        }
        if (c->invert_z) integ_invert_z(c);
        if (step != 0) {
            if (kind == PFMDS_NVT && c->nhc_fusable && c->nhc_ke_valid) {
                bool rebuild = false;
                for (auto& it : c->inter)
                    for (int j = 0; j < it.nl_n; ++j) rebuild |= (step % it.nl[j].period == 0) || !it.nl[j].built;
                integ_nvt_open_kick_drift(c, dt, rebuild);  // consumes the pending closing scale of the previous step
                c->nhc_pending = false;
            } else {
                integ_flush_pending(c);
                if (kind == PFMDS_NVT)
                    for (auto& th : c->nhc) integ_nhc_half(c, th, dt);
                integ_kick_drift(c, dt);
            }
        }
    }
    {
        PhaseTimer t(c, 1);
        update_lists(c, step);
    }
    {
        PhaseTimer t(c, 4);
        if (step % c->zero_momentum_period == 0) integ_zero_momentum(c);
        compute_forces(c, with_energy, step != 0);
    }
    if (step != 0) {
        PhaseTimer t(c, 0);
        const bool summed = c->fbuf_active;   // the forces are still in the per-interaction buffers: sum and kick in one kernel
        if (kind == PFMDS_NVT && c->nhc_fusable) {
            if (summed) integ_sum_forces(c, 2, dt);
            else integ_nvt_kick_close(c, dt);
            c->nhc_pending = true;
            c->nhc_ke_valid = true;
        } else {
            if (summed) integ_sum_forces(c, 1, dt);
            else integ_kick(c, dt);
            if (kind == PFMDS_NVT)
                for (auto& th : c->nhc) integ_nhc_half(c, th, dt);
            if (kind == PFMDS_NVMS) integ_quench(c);
            c->nhc_ke_valid = false;
        }
    }
}

__global__ void k_max_int(int n, const int* __restrict__ a, int* out) {
    int m = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, a[i]);
#ifdef PFMDS_COOP
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
#else  // host replay: no lane exchange, every thread contributes
    atomicMax(out, m);
#endif
}

}  // namespace

// ---- per-kernel profiling -------------------------------------------------------------------------
void prof_flush(pfmds_ctx* c) {
    if (c->prof_used == 0) return;
    cudaStreamSynchronize(c->st);
    for (size_t k = 0; k < c->prof_used; ++k) {
        float ms = 0;
        cudaEventElapsedTime(&ms, c->prof_ev[2 * k], c->prof_ev[2 * k + 1]);
        c->prof_ms[c->prof_slot[k]] += ms;
        c->prof_cnt[c->prof_slot[k]] += 1;
    }
    c->prof_used = 0;
}
KTimer::KTimer(pfmds_ctx* c_, int slot_) : c(c_), slot(slot_) {
    if (!c->prof_on) return;
    if (c->prof_used * 2 + 2 > c->prof_ev.size()) {
        if (c->prof_ev.size() >= 16384) prof_flush(c);
        else for (int k = 0; k < 1024; ++k) { cudaEvent_t e; cudaEventCreate(&e); c->prof_ev.push_back(e); }
    }
    if (c->prof_slot.size() < c->prof_ev.size() / 2) c->prof_slot.resize(c->prof_ev.size() / 2);
    cudaEventRecord(c->prof_ev[2 * c->prof_used], c->st);
}
KTimer::~KTimer() {
    if (!c->prof_on) return;
    cudaEventRecord(c->prof_ev[2 * c->prof_used + 1], c->st);
    c->prof_slot[c->prof_used] = slot;
    c->prof_used += 1;
}

#ifdef __CUDACC__
// ---- peak micro-benchmarks (the roofline denominators MEASURED_PEAKS.json does not carry) ----------
// 8 independent DFMA chains per thread; 2 flop per DFMA.
__global__ void __launch_bounds__(256) k_dfma_peak(int iters, double seed, double* out) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, b = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
        a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) out[0] = s;
}
// device self-test of mathx.cuh against the CUDA math library: max errors over a sweep
__global__ void k_math_selftest(int n, double* out) {
    double e_exp = 0, e_sw = 0, e_rs = 0, e_seed = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double u = (i + 0.5) / n;
        double x = -600. + 1200. * u;
        double a = exp(x), b = mx::exp_nc(x);
        e_exp = fmax(e_exp, fabs(a - b) / a);
        b = mx::exp_fast(x);
        e_exp = fmax(e_exp, fabs(a - b) / a);
        double ang = 3.14159265358979 * u, f, s, sr, cr;
        mx::cos_switch(ang, f, s);
        sincos(ang, &sr, &cr);
        e_sw = fmax(e_sw, fmax(fabs(f - (1. + cr) / 2), fabs(s - sr)));
        mx::sincos_0pi(ang, s, f);
        e_sw = fmax(e_sw, fmax(fabs(f - cr), fabs(s - sr)));
        double v = exp(-20. + 40. * u);
        e_rs = fmax(e_rs, fabs(mx::rsqrt_fast(v) * sqrt(v) - 1.));
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v));
        e_seed = fmax(e_seed, fabs(y * sqrt(v) - 1.));
    }
    // order-independent max through the bit pattern (all values are non-negative)
    atomicMax((unsigned long long*)&out[0], (unsigned long long)__double_as_longlong(e_exp));
    atomicMax((unsigned long long*)&out[1], (unsigned long long)__double_as_longlong(e_sw));
    atomicMax((unsigned long long*)&out[2], (unsigned long long)__double_as_longlong(e_rs));
    atomicMax((unsigned long long*)&out[3], (unsigned long long)__double_as_longlong(e_seed));
}
// running maximum that a NaN cannot hide behind (fmax returns its other argument)
__device__ __forceinline__ double worst(double e, double v) { return v == v ? fmax(e, v) : 1.0; }
// the short forms of the second-generation rjl kernels (exp_m / exp_m2, rsqrt_q, cos_switch_m, half_switch)
__global__ void k_math_selftest2(int n, double* out) {
    double e_exp = 0, e_sw = 0, e_rs = 0, e_wide = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double u = (i + 0.5) / n;
        double x = -40. + 80. * u, xb = 3. - 19. * u;
        double a = exp(x), b = mx::exp_m(x), ea, eb;
        e_exp = worst(e_exp, fabs(a - b) / a);
        mx::exp_m2(x, xb, ea, eb, 1.4426950408889634, -6.93147180559945309417e-01);
        e_exp = worst(worst(e_exp, fabs(a - ea) / a), fabs(exp(xb) - eb) / exp(xb));
        x = -600. + 1200. * u;
        a = exp(x);
        e_wide = worst(e_wide, fabs(a - mx::exp_m(x)) / a);
        double ang = 3.14159265358979 * u, f, s, sr, cr;
        mx::cos_switch_m(fma(ang, 0.5, -0.78539816339744830962), f, s);
        sincos(ang, &sr, &cr);
        e_sw = worst(worst(e_sw, fabs(f - (1. + cr) / 2)), fabs(s - sr));
        e_sw = worst(e_sw, fabs(mx::half_switch(ang - 1.57079632679489661923) - (1. + cr) / 2));
        double v = exp(-20. + 40. * u);
        e_rs = worst(e_rs, fabs(mx::rsqrt_q(v) * sqrt(v) - 1.));
    }
    atomicMax((unsigned long long*)&out[0], (unsigned long long)__double_as_longlong(e_exp));
    atomicMax((unsigned long long*)&out[1], (unsigned long long)__double_as_longlong(e_sw));
    atomicMax((unsigned long long*)&out[2], (unsigned long long)__double_as_longlong(e_rs));
    atomicMax((unsigned long long*)&out[3], (unsigned long long)__double_as_longlong(e_wide));
}
#endif  // __CUDACC__
static bool rjl_short_forms_ok(int dev) {
#ifndef __CUDACC__
    (void)dev;
    return true;  // host replay: mathx.cuh is checked by tests/test_mathx.py
#else
    static std::mutex mu;
    static int verdict[64] = {0};  // per device: 0 unknown, 1 ok, 2 out of bounds
    std::lock_guard<std::mutex> lock(mu);
    int& v = verdict[dev & 63];
    if (v == 0) {
        double* d = nullptr;
        double err[4] = {1., 1., 1., 1.};
        CK(cudaMalloc(&d, sizeof err));
        CK(cudaMemset(d, 0, sizeof err));
        LAUNCH((k_math_selftest2), 64, 256, 0, 1 << 16, d);
        CK(cudaMemcpy(err, d, sizeof err, cudaMemcpyDeviceToHost));
        CK(cudaFree(d));
        // exp_m / exp_m2 relative (6e-15 by construction), switches absolute (5e-15), rsqrt_q relative (1.5 x seed error^2 = 1.2e-12)
        v = (err[0] < 2e-14 && err[1] < 4e-14 && err[2] < 1e-11) ? 1 : 2;
    }
    return v == 1;
#endif
}
__global__ void k_copy(size_t n, const double4* __restrict__ a, double4* __restrict__ b) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void k_sum_int(int n, const int* __restrict__ a, unsigned long long* out) {
    unsigned long long s = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += (unsigned long long)a[i];
#ifdef PFMDS_COOP
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
#else  // host replay: no lane exchange, every thread contributes
    atomicAdd(out, s);
#endif
}
__global__ void k_count_within(int N, const double4* __restrict__ pos, ListView lv, BoxD box, double r2max, unsigned long long* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long s = 0;
    if (i < N) {
        const int n = lv.nnum[i];
        const double4 pi = pos[i];
        for (int p = 0; p < n; ++p) {
            const double4 pj = pos[lv.nlist[(size_t)p * lv.stride + i]];
            double dx = min_image(pj.x - pi.x, box.h[0], box.L[0]), dy = min_image(pj.y - pi.y, box.h[1], box.L[1]), dz = min_image(pj.z - pi.z, box.h[2], box.L[2]);
            s += (dx * dx + dy * dy + dz * dz < r2max) ? 1ull : 0ull;
        }
    }
    if (s) atomicAdd(out, s);
}
__global__ void k_upload_scatter(int N, const int* __restrict__ orig, const double* __restrict__ hp, const double* __restrict__ hv,
                                 double4* __restrict__ pos, double4* __restrict__ vel) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    int f = orig[s];
    if (hp) { double4 p = pos[s]; p.x = hp[3 * f]; p.y = hp[3 * f + 1]; p.z = hp[3 * f + 2]; pos[s] = p; }
    if (hv) { double4 v = vel[s]; v.x = hv[3 * f]; v.y = hv[3 * f + 1]; v.z = hv[3 * f + 2]; vel[s] = v; }
}

__global__ void k_upload_forces(int N, const int* __restrict__ orig, const double* __restrict__ hf, double4* __restrict__ frc) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    const double* f = hf + 3 * (size_t)orig[s];
    frc[s] = make_double4(f[0], f[1], f[2], 0.);
}

// file-order copy of one state array: out[3 f + k] = component k of the atom whose file index is f = orig[slot]
__global__ void k_download_gather(int N, const int* __restrict__ orig, const double4* __restrict__ a, double* __restrict__ out) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= N) return;
    const double4 v = a[s];
    double* o = out + 3 * (size_t)orig[s];
    o[0] = v.x; o[1] = v.y; o[2] = v.z;
}

extern "C" {

int pfmds_create(pfmds_ctx** out, int device, int n_atoms, const double* pos, const double* vel, const double* mass, const double box[3]) {
    if (!out) return PFMDS_ERR_INVALID;
    pfmds_ctx* c = new pfmds_ctx;
    *out = c;
    return guarded(c, [&] {
        if (n_atoms < 1 || !pos || !vel || !mass || !box) fail(PFMDS_ERR_INVALID, "error: bad arguments to pfmds_create");
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
            fail(PFMDS_ERR_CUDA, "no CUDA device: pfmds_b200 has no CPU fallback");
        if (device < 0 || device >= ndev) fail(PFMDS_ERR_INVALID, "error: CUDA device " + std::to_string(device) + " does not exist");
        c->dev = device;
        CK(cudaSetDevice(device));
        CK(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
        g_live_contexts[device & 63].fetch_add(1);
        c->counted = true;
        c->N = n_atoms;
        c->stride = ((size_t)n_atoms + 31) / 32 * 32;
        for (int k = 0; k < 3; ++k) { c->box.L[k] = box[k]; c->box.h[k] = 0.5 * box[k]; }  // md_read_write.f90:32-35
        const size_t S = c->stride;
        CK(cudaMalloc(&c->pos, sizeof(double4) * S)); CK(cudaMalloc(&c->pos2, sizeof(double4) * S));
        CK(cudaMalloc(&c->vel, sizeof(double4) * S)); CK(cudaMalloc(&c->vel2, sizeof(double4) * S));
        CK(cudaMalloc(&c->frc, sizeof(double4) * S));
        CK(cudaMalloc(&c->gmask, sizeof(uint32_t) * S)); CK(cudaMalloc(&c->gmask2, sizeof(uint32_t) * S));
        CK(cudaMalloc(&c->orig, sizeof(int) * S)); CK(cudaMalloc(&c->orig2, sizeof(int) * S));
        CK(cudaMalloc(&c->cell_atoms, sizeof(int) * S)); CK(cudaMalloc(&c->cid, sizeof(int) * S));
        CK(cudaMalloc(&c->posf, sizeof(float4) * S));
        size_t nparts = (S + 127) / 128 + RED_BLOCKS;
        CK(cudaMalloc(&c->part, sizeof(double) * 16 * nparts));
        c->part_cap = 16 * nparts;
        CK(cudaMalloc(&c->red, sizeof(double) * 64));
        CK(cudaMalloc(&c->err, sizeof(int) * PFMDS_ERRW));
        CK(cudaMemset(c->err, 0, sizeof(int) * PFMDS_ERRW));
        CK(cudaMemset(c->frc, 0, sizeof(double4) * S));
        std::vector<double4> hp(S, make_double4(0, 0, 0, 0)), hv(S, make_double4(0, 0, 0, 1));
        std::vector<int> ho(S, 0);
        for (int i = 0; i < n_atoms; ++i) {
            hp[(size_t)i] = make_double4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], 0.);
            hv[(size_t)i] = make_double4(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2], mass[i]);
            ho[(size_t)i] = i;
        }
        CK(cudaMemcpy(c->pos, hp.data(), sizeof(double4) * S, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->vel, hv.data(), sizeof(double4) * S, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->orig, ho.data(), sizeof(int) * S, cudaMemcpyHostToDevice));
        CK(cudaMemset(c->gmask, 0, sizeof(uint32_t) * S));
        read_env(c, n_atoms, false);
        CK(cudaEventCreate(&c->ev0)); CK(cudaEventCreate(&c->ev1));
        // staging block of pfmds_upload / pfmds_download (file-order x y z of up to three arrays): allocated here, once, so that
        // no call on the data path allocates (a first-use cudaMalloc of 72 MB inside a timed region costs milliseconds)
        c->io_cap = 9 * (size_t)n_atoms;
        CK(cudaMalloc(&c->io_stage, sizeof(double) * c->io_cap));
    });
}

int pfmds_set_group(pfmds_ctx* c, int g, int n, const int* idx) {
    if (!c) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        if (c->finalized) fail(PFMDS_ERR_INVALID, "error: groups must be defined before the first pfmds_advance");
        if (g < 1 || n < 0 || (n > 0 && !idx)) fail(PFMDS_ERR_INVALID, "error: bad arguments to pfmds_set_group");
        if ((int)c->groups.size() < g) c->groups.resize((size_t)g);
        c->groups[(size_t)g - 1].assign(idx, idx + n);
    });
}

int pfmds_set_roles(pfmds_ctx* c, int am, int xyz, int z, int all) {
    if (!c) return PFMDS_ERR_INVALID;
    return guarded(c, [&] { c->all_moving = am; c->xyz_moving = xyz; c->z_moving = z; c->all_atoms = all; });
}

int pfmds_add_nhc(pfmds_ctx* c, int g, double T, int M, double q1) {
    if (!c) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        if (c->finalized) fail(PFMDS_ERR_INVALID, "error: thermostats must be defined before the first pfmds_advance");
        if (M < 1 || q1 < 0. || T < 0.) fail(PFMDS_ERR_NHC_PARAMS, "error: wrong nhc parameters");  // md_integrators.f90:187
        Nhc t; t.group = g; t.M = M; t.temperature = T;
        CK(cudaMalloc(&t.state, sizeof(double) * ((size_t)3 * M + 4)));
        std::vector<double> st((size_t)3 * M + 4, 0.);
        st[2 * (size_t)M] = q1;
        CK(cudaMemcpy(t.state, st.data(), sizeof(double) * st.size(), cudaMemcpyHostToDevice));
        c->nhc.push_back(t);
    });
}

int pfmds_set_misc(pfmds_ctx* c, int zmp, int inv) {
    if (!c) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        if (zmp < 1) fail(PFMDS_ERR_INVALID, "error: zero_momentum_period must be positive");
        c->zero_momentum_period = zmp; c->invert_z = inv != 0;
    });
}

int pfmds_add_group_change(pfmds_ctx* c, int from, int to, int ts1, int ts2, int frec) {
    if (!c) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        if (c->finalized) fail(PFMDS_ERR_INVALID, "error: group changes must be defined before the first pfmds_advance");
        if (from < 1 || to < 1) fail(PFMDS_ERR_INVALID, "error: group number out of range in a group change");
        c->changes.push_back(pfmds_ctx::GroupChange{from, to, ts1, ts2, frec});
    });
}

int pfmds_group_size(pfmds_ctx* c, int g, int* n) {
    if (!c || !n) return PFMDS_ERR_INVALID;
    return guarded(c, [&] { *n = (int)group_size(c, g); });
}

int pfmds_add_interaction(pfmds_ctx* c, const char* name, int np, const double* p, int nl_n, const int* gn, const int* maxn, const double* rcut,
                          const int* period) {
    if (!c) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        if (c->finalized) fail(PFMDS_ERR_INVALID, "error: interactions must be defined before the first pfmds_advance");
        Inter it;
        it.name = name ? name : "";
        it.kind = kind_of(it.name);
        if (it.kind < 0) fail(PFMDS_ERR_UNKNOWN_INTERACTION, "error: unknown interaction name");  // md_interactions.f90:119-120
        it.nl_n = nl_n_of(it.kind);
        if (nl_n != it.nl_n) fail(PFMDS_ERR_INVALID, "error: " + it.name + " needs " + std::to_string(it.nl_n) + " neighbour list lines");
        auto need = [&](int n) { if (np != n || !p) fail(PFMDS_ERR_INVALID, "error: " + it.name + " needs " + std::to_string(n) + " parameters"); };
        switch (it.kind) {
        case K_LJ: need(4); it.lj = LJp{p[0], p[1], p[2], p[3]}; break;
        case K_LJ1G: {  // LennardJones_1g.f90:21-24
            need(4);
            double s2 = p[1] * p[1], s6 = s2 * s2 * s2, s12 = s6 * s6;
            it.lj1g = LJ1Gp{p[2], p[3], 4. * p[0] * s6, 4. * p[0] * s12, 6. * 4. * p[0] * s6, 12. * 4. * p[0] * s12};
            break;
        }
        case K_LJC: need(6); it.ljc = LJCp{p[0], p[1], p[2], p[3], p[4], p[5] != 0.}; break;
        case K_MORSEC: need(7); it.mor = MORp{p[0], p[1], p[2], p[3], p[4], p[5], p[6] != 0.}; break;
        case K_TB: need(10); it.tb = TBp{p[0], p[1], p[2], p[3], p[4], p[5], p[6] * p[6], p[7] * p[7], p[8], p[9]}; break;  // TersoffBrenner.f90:19-20
        case K_RJL: need(7); it.rjl = RJLp{p[0], p[1], p[2], p[3], p[4], p[5], p[6]}; break;
        case K_REBOSC:  // REBOsolidcarbon.f90:12-25
            need(18);
            it.reb = REBp{p[0], p[1], p[2], {p[3], p[4], p[5]}, {p[6], p[7], p[8]}, p[9], {p[10], p[11], p[12], p[13], p[14], p[15]}, p[16], p[17]};
            break;
        }
        for (int j = 0; j < nl_n; ++j) {
            it.nl[j].g1 = gn[2 * j]; it.nl[j].g2 = gn[2 * j + 1]; it.nl[j].maxn = maxn[j]; it.nl[j].rcut = rcut[j]; it.nl[j].period = period[j];
        }
        c->inter.push_back(it);
    });
}

// One md step of a call that started at step `first`.  Steady-state steps (no list rebuild, no momentum removal, no energy
// request, not the first of the call) of small systems are replayed from a CUDA graph captured from this very code path: same
// kernels, same order, fewer launch gaps.
static void run_step(pfmds_ctx* c, int s, int first, int kind, double dt, bool with_energy, bool next_follows = false, int nrep = 1) {
    // the thermostat's opening half step of step s+1 can ride in the closing kernel of step s when s+1 follows inside this call,
    // nothing reads or regroups the chains in between (no log row, no deposition) and the fused NVT path is in use
    c->pre_open = next_follows && kind == PFMDS_NVT && c->nhc_fusable && !c->slab && c->changes.empty() && !with_energy && c->pre_open_enabled;
    bool rebuild = false, rebuild_all = true, built_all = true;
    for (auto& it : c->inter)
        for (int j = 0; j < it.nl_n; ++j) {
            const bool rb = (s % it.nl[j].period == 0) || !it.nl[j].built;
            rebuild |= rb; rebuild_all &= rb; built_all &= it.nl[j].built;
        }
    const bool regrouped = apply_group_changes(c, s);
    // a step that rebuilds every list (binning, cell re-sort into the other half of the double-buffered state, all builds) is a fixed
    // sequence of launches too: its graph is keyed by the buffer it starts from, the replay redoes the host's pointer swap
#ifdef PFMDS_COOP
    const bool rebuild_ok = !rebuild || (rebuild_all && built_all && c->graph_rebuilds && nrep == 1);
#else   // serial host replay of the test suite: nl_bin_atoms runs its prefix sums as host loops there, which a capture cannot record
    const bool rebuild_ok = !rebuild;
    (void)rebuild_all; (void)built_all;
#endif
    const bool graphable = !regrouped && c->use_graphs && !c->slab && !c->prof_on && !c->timers_on && s != 0 && s != first && rebuild_ok &&
                           (s % c->zero_momentum_period != 0) && !with_energy;
    if (!graphable) {
        for (int r = 1; r < nrep; ++r) {   // (a run the caller judged replayable and this test did not: still every step, one by one)
            do_step(c, s + r - 1, kind, dt, s + r - 1 == first, false);
            apply_group_changes(c, s + r);
        }
        if (nrep > 1) s += nrep - 1;
        do_step(c, s, kind, dt, s == first, with_energy);
        if (c->slab && std::getenv("PFMDS_SLAB_DEBUG")) {
            std::fprintf(stderr, "[slab %d] step %d queued\n", slab_rank(c), s); std::fflush(stderr);
            cudaError_t e = cudaStreamSynchronize(c->st);
            std::fprintf(stderr, "[slab %d] step %d done (%s)\n", slab_rank(c), s, cudaGetErrorString(e)); std::fflush(stderr);
        }
        return;
    }
    pfmds_ctx::StepGraph* g = nullptr;
    for (auto& e : c->graphs)
        if (e.kind == kind && e.dt == dt && e.pos == (const void*)c->pos && e.pending == c->nhc_pending && e.ke_valid == c->nhc_ke_valid &&
            e.opened == c->nhc_opened && e.pre_open == c->pre_open && e.alone == (g_live_contexts[c->dev & 63].load() <= 1) && e.nsteps == nrep && e.rebuild == rebuild) g = &e;
    if (!g) {
        pfmds_ctx::StepGraph e{kind, dt, (const void*)c->pos, c->nhc_pending, c->nhc_ke_valid, c->nhc_opened, c->pre_open, g_live_contexts[c->dev & 63].load() <= 1, nullptr, 0, nrep, rebuild};
        const long long l0 = c->launches;
        cudaGraph_t graph = nullptr;
        CK(cudaStreamBeginCapture(c->st, cudaStreamCaptureModeThreadLocal));
        try { for (int r = 0; r < nrep; ++r) do_step(c, s + r, kind, dt, false); } catch (...) { cudaStreamEndCapture(c->st, &graph); if (graph) cudaGraphDestroy(graph); throw; }
        CK(cudaStreamEndCapture(c->st, &graph));
        CK(cudaGraphInstantiate(&e.exec, graph, 0));
        CK(cudaGraphDestroy(graph));
        e.launches = c->launches - l0;
        c->launches = l0;
        // capturing ran the host-side bookkeeping of one step: the flags now describe the state AFTER a step; a step
        // is only graphable again from the same entry state, which holds in steady state (checked by the key)
        if (c->graphs.size() >= 16) { cudaGraphExecDestroy(c->graphs.front().exec); c->graphs.erase(c->graphs.begin()); }
        c->graphs.push_back(e);
        g = &c->graphs.back();
        CK(cudaGraphLaunch(g->exec, c->st));
        c->launches += g->launches;
        return;
    }
    CK(cudaGraphLaunch(g->exec, c->st));
    c->launches += g->launches;
    c->energy_valid = false;
    if (rebuild) {   // nl_bin_atoms(reorder): the re-sorted state is in the other buffers
        std::swap(c->pos, c->pos2); std::swap(c->vel, c->vel2); std::swap(c->gmask, c->gmask2); std::swap(c->orig, c->orig2);
        c->identity_order = true;
    }
    // host-side bookkeeping of do_step for this integrator
    if (kind == PFMDS_NVT && c->nhc_fusable) { c->nhc_pending = true; c->nhc_ke_valid = true; c->nhc_opened = c->pre_open; }
    else { c->nhc_pending = false; c->nhc_ke_valid = false; c->nhc_opened = false; }
}

}  // extern "C"
// Steady-state steps of small systems are replayed from CUDA graphs; where a run of them allows it one graph launch carries
// c->graph_steps steps (the gap between two graph launches is 4.4 us, between two nodes of one graph 2.3 us).  From step s: can the
// next `g` steps of the call [first, end) go as one graph?  All of them must be plain (run_step's `graphable`) and followed by
// another step of the call, so that the thermostat flags are those of the steady state on both sides of every step.
template <class EnergyAt>
static bool graph_run_ok(pfmds_ctx* c, int s, int first, int end, int g, EnergyAt energy_at) {
    if (g < 2 || !c->use_graphs || c->slab || c->prof_on || c->timers_on || !c->changes.empty() || s + g >= end) return false;
    for (int t = s; t < s + g; ++t) {
        if (t == 0 || t == first || (t % c->zero_momentum_period == 0) || energy_at(t)) return false;
        for (auto& it : c->inter)
            for (int j = 0; j < it.nl_n; ++j)
                if ((t % it.nl[j].period == 0) || !it.nl[j].built) return false;
    }
    return true;
}
extern "C" {
static int advance_impl(pfmds_ctx* c, int kind, double dt, int first, int n, bool energy_last);
int pfmds_advance(pfmds_ctx* c, int kind, double dt, int first, int n) { return advance_impl(c, kind, dt, first, n, false); }
int pfmds_advance_with_energy(pfmds_ctx* c, int kind, double dt, int first, int n) { return advance_impl(c, kind, dt, first, n, true); }
static int advance_impl(pfmds_ctx* c, int kind, double dt, int first, int n, bool energy_last) {
    if (!c) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        if (n < 0 || first < 0) fail(PFMDS_ERR_INVALID, "error: bad step range");
        CK(cudaSetDevice(c->dev));
        finalize(c);
        const int end = first + n, g = c->graph_steps;
        for (int s = first; s < end; ++s) {
            if (graph_run_ok(c, s, first, end, g, [&](int t) { return energy_last && t == end - 1; })) { run_step(c, s, first, kind, dt, false, true, g); s += g - 1; continue; }
            run_step(c, s, first, kind, dt, energy_last && s == end - 1, s + 1 < end);
        }
        CK(cudaGetLastError());
    });
}

// one row of the device-resident energy log: [e_inter(n_inter), KE(all_moving), x v q of every chain]
// (head: energies and KE, with the first pack of thermostats; further packs of up to NHC_MAXF chains append at `o`)
__global__ void k_log_row(int n_inter, const double* __restrict__ energy, const double* __restrict__ ke, NhcPack P, double* __restrict__ row, int o) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (n_inter >= 0) {
        for (int k = 0; k < n_inter; ++k) row[k] = energy[k];
        row[n_inter] = ke[0];
    }
    for (int t = 0; t < P.n; ++t)
        for (int i = 0; i < 3 * P.M[t]; ++i) row[o++] = P.state[t][i];
}

// md() with period_log = 1 asks for the energies after every step (md_simulation.f90:188-199); through pfmds_advance +
// pfmds_energies that is one host round trip per step.  Here the steps of the call that satisfy mod(step, log_period) == 0
// evaluate the energies in their force pass and append what pfmds_energies would return to a log kept on the device; the
// host gets all rows with one copy when the call ends.  Row = e_inter[n_inter], KE, temperature, e_nhc[n_nhc]; the numbers
// are those of the one-step-at-a-time sequence, bit for bit (same kernels in the same order).
int pfmds_advance_logged(pfmds_ctx* c, int kind, double dt, int first, int n, int log_period, double* rows, int row_len, int* n_rows) {
    if (!c) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        if (n < 0 || first < 0 || log_period < 1) fail(PFMDS_ERR_INVALID, "error: bad step range");
        CK(cudaSetDevice(c->dev));
        finalize(c);
        const int nI = (int)c->inter.size(), nT = (int)c->nhc.size();
        if (!rows || row_len < nI + 2 + nT) fail(PFMDS_ERR_INVALID, "error: pfmds_advance_logged needs rows of n_interactions + 2 + n_nhc doubles");
        int D = nI + 1;
        for (auto& t : c->nhc) D += 3 * t.M;
        int want = 0;
        for (int s = first; s < first + n; ++s) want += (s % log_period == 0);
        if ((size_t)want * D > c->log_cap) {   // beyond the block finalize() allocated (LOG_ROWS_PREALLOC rows): grow, outside any step
            if (c->logbuf) CK(cudaFree(c->logbuf));
            c->logbuf = nullptr; c->log_cap = 0;
            CK(cudaMalloc(&c->logbuf, sizeof(double) * (size_t)want * D));
            c->log_cap = (size_t)want * D;
        }
        // thermostat states go out in packs of NHC_MAXF chains (kernel parameter block), any number of thermostats
        std::vector<NhcPack> packs;
        std::vector<int> pack_at;
        {
            int o = nI + 1;
            for (int k0 = 0; k0 < nT || packs.empty(); k0 += NHC_MAXF) {
                NhcPack P{};
                pack_at.push_back(o);
                for (int k = k0; k < nT && k < k0 + NHC_MAXF; ++k) { P.state[P.n] = c->nhc[k].state; P.M[P.n] = c->nhc[k].M; o += 3 * c->nhc[k].M; ++P.n; }
                packs.push_back(P);
            }
        }
        std::vector<int> gsize;
        int r = 0;
        for (int s = first; s < first + n; ++s) {
            const bool logged = s % log_period == 0;
            if (!logged && graph_run_ok(c, s, first, first + n, c->graph_steps, [&](int t) { return t % log_period == 0; })) {
                run_step(c, s, first, kind, dt, false, true, c->graph_steps);
                s += c->graph_steps - 1;
                continue;
            }
            run_step(c, s, first, kind, dt, logged, s + 1 < first + n);   // unlogged steady-state steps of small systems replay their CUDA graph, as in pfmds_advance
            if (!logged) continue;
            integ_flush_pending(c);                              // as pfmds_energies: KE of the velocities the host would download
            integ_kinetic_energy(c, c->all_moving, c->red);
            for (size_t q = 0; q < packs.size(); ++q) {
                LAUNCH((k_log_row), 1, 32, c->st, q == 0 ? nI : -1, c->energy, c->red, packs[q], c->logbuf + (size_t)r * D, pack_at[q]);
                c->launches += 1;
            }
            gsize.push_back(group_size(c, c->all_moving));
            ++r;
        }
        std::vector<double> h((size_t)r * D + 1, 0.);
        if (r) CK(cudaMemcpyAsync(h.data(), c->logbuf, sizeof(double) * (size_t)r * D, cudaMemcpyDeviceToHost, c->st));
        check_device_error(c);  // synchronises
        for (int i = 0; i < r; ++i) {
            const double* in = h.data() + (size_t)i * D;
            double* out = rows + (size_t)i * row_len;
            for (int k = 0; k < nI; ++k) out[k] = in[k];
            const double ke = in[nI];
            out[nI] = ke;
            out[nI + 1] = 2 * ke / PFMDS_KB / (3 * (double)gsize[(size_t)i]);  // calculate_temperature, md_general.f90:301-311
            const double* x = in + nI + 1;
            for (int k = 0; k < nT; ++k) {  // calculate_nose_hoover_chain_energy, md_integrators.f90:247-260
                const Nhc& t = c->nhc[(size_t)k];
                const double *v = x + t.M, *q = x + 2 * t.M;
                const double kt = PFMDS_KB * t.temperature;
                double e = q[0] / 2 * (v[0] * v[0]) + 3. * t.L * kt * x[0];
                for (int j = 1; j < t.M; ++j) e = e + q[j] / 2 * (v[j] * v[j]) + kt * x[j];
                out[nI + 2 + k] = e;
                x += 3 * t.M;
            }
        }
        if (n_rows) *n_rows = r;
    });
}

int pfmds_synchronize(pfmds_ctx* c) {
    if (!c) return PFMDS_ERR_INVALID;
    return guarded(c, [&] { CK(cudaSetDevice(c->dev)); CK(cudaStreamSynchronize(c->st)); check_device_error(c); });
}

int pfmds_energies(pfmds_ctx* c, double* e_inter, double* ke, double* temp, double* e_nhc) {
    if (!c) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        CK(cudaSetDevice(c->dev));
        finalize(c);
        integ_flush_pending(c);
        {
            PhaseTimer t(c, 5);
            if (!c->energy_valid)  // else: computed by the force pass of the last step (pfmds_advance_with_energy)
                for (size_t k = 0; k < c->inter.size(); ++k) energy_interaction(c, (int)k);
            integ_kinetic_energy(c, c->all_moving, c->red);
        }
        std::vector<double> he(c->inter.size() + 1, 0.);
        if (!c->inter.empty()) CK(cudaMemcpyAsync(he.data(), c->energy, sizeof(double) * c->inter.size(), cudaMemcpyDeviceToHost, c->st));
        double hke = 0;
        CK(cudaMemcpyAsync(&hke, c->red, sizeof(double), cudaMemcpyDeviceToHost, c->st));
        std::vector<std::vector<double>> hs(c->nhc.size());
        for (size_t k = 0; k < c->nhc.size(); ++k) {
            hs[k].resize((size_t)3 * c->nhc[k].M + 2);
            CK(cudaMemcpyAsync(hs[k].data(), c->nhc[k].state, sizeof(double) * hs[k].size(), cudaMemcpyDeviceToHost, c->st));
        }
        check_device_error(c);  // synchronises
        if (e_inter) for (size_t k = 0; k < c->inter.size(); ++k) e_inter[k] = he[k];
        if (ke) *ke = hke;
        // calculate_temperature, md_general.f90:301-311
        if (temp) *temp = 2 * hke / PFMDS_KB / (3 * (double)group_size(c, c->all_moving));
        if (e_nhc)
            for (size_t k = 0; k < c->nhc.size(); ++k) {  // calculate_nose_hoover_chain_energy, md_integrators.f90:247-260
                const Nhc& t = c->nhc[k];
                const double *x = hs[k].data(), *v = x + t.M, *q = x + 2 * t.M;
                double kt = PFMDS_KB * t.temperature;
                double e = q[0] / 2 * (v[0] * v[0]) + 3. * t.L * kt * x[0];
                for (int i = 1; i < t.M; ++i) e = e + q[i] / 2 * (v[i] * v[i]) + kt * x[i];
                e_nhc[k] = e;
            }
    });
}

int pfmds_diagnostics(pfmds_ctx* c, double fs[3], double mc[3], double mcv[3], double* vmax, int* nl_load) {
    if (!c) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        CK(cudaSetDevice(c->dev));
        finalize(c);
        integ_flush_pending(c);
        integ_diagnostics(c, c->red + 32);
        size_t nl_total = 0;
        for (auto& it : c->inter) nl_total += (size_t)it.nl_n;
        int* d_max = nullptr;
        std::vector<int> hmax(nl_total + 1, 0);
        if (nl_total) {
            CK(cudaMalloc(&d_max, sizeof(int) * nl_total));
            CK(cudaMemsetAsync(d_max, 0, sizeof(int) * nl_total, c->st));
            size_t k = 0;
            for (auto& it : c->inter)
                for (int j = 0; j < it.nl_n; ++j, ++k) { LAUNCH((k_max_int), 64, 256, c->st, c->N, it.nl[j].nnum, d_max + k); c->launches += 1; }
            if (c->slab) slab_allreduce_max_int(c, d_max, (int)nl_total);
            CK(cudaMemcpyAsync(hmax.data(), d_max, sizeof(int) * nl_total, cudaMemcpyDeviceToHost, c->st));
        }
        double h[11];
        CK(cudaMemcpyAsync(h, c->red + 32, sizeof h, cudaMemcpyDeviceToHost, c->st));
        check_device_error(c);
        if (d_max) cudaFree(d_max);
        for (int k = 0; k < 3; ++k) {
            if (fs) fs[k] = h[k];
            if (mc) mc[k] = h[3 + k] / h[9];
            if (mcv) mcv[k] = h[6 + k] / h[9];
        }
        if (vmax) *vmax = std::sqrt(h[10]);
        if (nl_load) for (size_t k = 0; k < nl_total; ++k) nl_load[k] = hmax[k];
    });
}

int pfmds_download(pfmds_ctx* c, double* pos, double* vel, double* frc) {
    if (!c) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        CK(cudaSetDevice(c->dev));
        if (c->slab) fail(PFMDS_ERR_INVALID, "error: use pfmds_slab_download on a slab context");
        const size_t n3 = 3 * (size_t)c->N;
        integ_flush_pending(c);
        // The device undoes its own cell ordering: each requested array is gathered into file order (x y z per atom, the
        // layout of the caller's buffer) in a staging block, then copied out with one transfer per array straight into the
        // caller's memory -- no host-side permutation, no intermediate host buffer.
        double* out[3] = {pos, vel, frc};
        const double4* src[3] = {c->pos, c->vel, c->frc};
        int want = 0;
        for (double* o : out) want += o != nullptr;
        if (n3 * (size_t)want > c->io_cap) {  // staging block of the context, grown on demand and kept (no allocation per call)
            if (c->io_stage) CK(cudaFree(c->io_stage));
            c->io_stage = nullptr; c->io_cap = 0;
            CK(cudaMalloc(&c->io_stage, sizeof(double) * n3 * (size_t)want));
            c->io_cap = n3 * (size_t)want;
        }
        double* stage = c->io_stage;
        int k = 0;
        for (int a = 0; a < 3; ++a) {
            if (!out[a]) continue;
            LAUNCH((k_download_gather), (c->N + 255) / 256, 256, c->st, c->N, c->orig, src[a], stage + n3 * (size_t)k);
            c->launches += 1;
            ++k;
        }
        k = 0;
        for (int a = 0; a < 3; ++a) {
            if (!out[a]) continue;
            CK(cudaMemcpyAsync(out[a], stage + n3 * (size_t)k, sizeof(double) * n3, cudaMemcpyDeviceToHost, c->st));
            ++k;
        }
        check_device_error(c);  // synchronises
    });
}

int pfmds_neighbours(pfmds_ctx* c, int inter, int list, int* nlist, int* nnum, int* lessnnum) {
    if (!c) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        CK(cudaSetDevice(c->dev));
        if (!c->finalized) fail(PFMDS_ERR_INVALID, "error: no neighbour lists before the first pfmds_advance");
        if (c->slab) fail(PFMDS_ERR_INVALID, "error: pfmds_neighbours is not available on a slab context");
        if (inter < 0 || inter >= (int)c->inter.size() || list < 0 || list >= c->inter[(size_t)inter].nl_n) fail(PFMDS_ERR_INVALID, "error: no such neighbour list");
        const NList& l = c->inter[(size_t)inter].nl[list];
        const size_t N = (size_t)c->N, S = c->stride;
        std::vector<int> ho(N), hn(N), hl((size_t)l.maxn * S);
        CK(cudaMemcpyAsync(ho.data(), c->orig, sizeof(int) * N, cudaMemcpyDeviceToHost, c->st));
        CK(cudaMemcpyAsync(hn.data(), l.nnum, sizeof(int) * N, cudaMemcpyDeviceToHost, c->st));
        CK(cudaMemcpyAsync(hl.data(), l.nlist, sizeof(int) * hl.size(), cudaMemcpyDeviceToHost, c->st));
        check_device_error(c);
        const auto& G1 = group_of(c, l.g1);
        const auto& G2 = group_of(c, l.g2);
        std::vector<int> slot_of(N), local2(N, -1);
        for (size_t s = 0; s < N; ++s) slot_of[(size_t)ho[s]] = (int)s;
        for (size_t k = 0; k < G2.size(); ++k) local2[(size_t)G2[k] - 1] = (int)k;
        std::vector<std::pair<int, int>> row;  // (group-2 local, file index)
        for (size_t r = 0; r < G1.size(); ++r) {
            int fi = G1[r] - 1, s = slot_of[(size_t)fi];
            row.clear();
            for (int p = 0; p < hn[(size_t)s]; ++p) {
                int fj = ho[(size_t)hl[(size_t)p * S + (size_t)s]];
                row.emplace_back(local2[(size_t)fj], fj);
            }
            std::sort(row.begin(), row.end());
            int less = -1;
            for (size_t p = 0; p < row.size(); ++p)
                if (less == -1 && fi < row[p].second) less = (int)p;  // md_neighbours.f90:78
            if (nnum) nnum[r] = (int)row.size();
            if (lessnnum) lessnnum[r] = less == -1 ? (int)row.size() : less;
            if (nlist)
                for (int p = 0; p < l.maxn; ++p) nlist[r * (size_t)l.maxn + (size_t)p] = p < (int)row.size() ? row[(size_t)p].first + 1 : 0;
        }
    });
}

int pfmds_normals(pfmds_ctx* c, int inter, double* out) {
    if (!c) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        CK(cudaSetDevice(c->dev));
        if (inter < 0 || inter >= (int)c->inter.size() || !c->inter[(size_t)inter].gnorm) fail(PFMDS_ERR_INVALID, "error: not an ljc/morsec interaction");
        const Inter& it = c->inter[(size_t)inter];
        const size_t N = (size_t)c->N;
        std::vector<int> ho(N);
        std::vector<double4> g(N);
        CK(cudaMemcpyAsync(ho.data(), c->orig, sizeof(int) * N, cudaMemcpyDeviceToHost, c->st));
        CK(cudaMemcpyAsync(g.data(), it.gnorm, sizeof(double4) * N, cudaMemcpyDeviceToHost, c->st));
        check_device_error(c);
        std::vector<int> slot_of(N);
        for (size_t s = 0; s < N; ++s) slot_of[(size_t)ho[s]] = (int)s;
        const auto& G1 = group_of(c, it.nl[0].g1);
        for (size_t r = 0; r < G1.size(); ++r) {
            const double4& v = g[(size_t)slot_of[(size_t)G1[r] - 1]];
            out[3 * r] = v.x; out[3 * r + 1] = v.y; out[3 * r + 2] = v.z;
        }
    });
}

int pfmds_get_nhc(pfmds_ctx* c, int k, double* x, double* v) {
    if (!c) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        if (k < 0 || k >= (int)c->nhc.size()) fail(PFMDS_ERR_INVALID, "error: no such thermostat");
        CK(cudaSetDevice(c->dev));
        CK(cudaStreamSynchronize(c->st));
        const Nhc& t = c->nhc[(size_t)k];
        if (x) CK(cudaMemcpy(x, t.state, sizeof(double) * t.M, cudaMemcpyDeviceToHost));
        if (v) CK(cudaMemcpy(v, t.state + t.M, sizeof(double) * t.M, cudaMemcpyDeviceToHost));
    });
}
int pfmds_set_nhc(pfmds_ctx* c, int k, const double* x, const double* v) {
    if (!c) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        if (k < 0 || k >= (int)c->nhc.size()) fail(PFMDS_ERR_INVALID, "error: no such thermostat");
        CK(cudaSetDevice(c->dev));
        CK(cudaStreamSynchronize(c->st));
        const Nhc& t = c->nhc[(size_t)k];
        if (x) CK(cudaMemcpy(t.state, x, sizeof(double) * t.M, cudaMemcpyHostToDevice));
        if (v) CK(cudaMemcpy(t.state + t.M, v, sizeof(double) * t.M, cudaMemcpyHostToDevice));
    });
}


int pfmds_upload(pfmds_ctx* c, const double* pos, const double* vel) {
    if (!c) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        CK(cudaSetDevice(c->dev));
        if (c->slab) fail(PFMDS_ERR_INVALID, "error: pfmds_upload is not available on a slab context");
        const size_t n3 = 3 * (size_t)c->N;
        integ_flush_pending(c);
        c->nhc_ke_valid = false;
        c->energy_valid = false;
        // file-order input lands in the context's staging block (allocated in pfmds_create; stream order keeps it safe to reuse)
        double *dp = nullptr, *dv = nullptr;
        if (pos) { dp = c->io_stage; CK(cudaMemcpyAsync(dp, pos, sizeof(double) * n3, cudaMemcpyHostToDevice, c->st)); }
        if (vel) { dv = c->io_stage + n3; CK(cudaMemcpyAsync(dv, vel, sizeof(double) * n3, cudaMemcpyHostToDevice, c->st)); }
        LAUNCH((k_upload_scatter), (c->N + 255) / 256, 256, c->st, c->N, c->orig, dp, dv, c->pos, c->vel);
        c->launches += 1;
        // membership of every list was decided on the old positions: rebuild at the next step
        for (auto& it : c->inter) for (int j = 0; j < it.nl_n; ++j) it.nl[j].built = false;
    });
}

// ---- exact restart -------------------------------------------------------------
// blob = [magic, n_nhc, n_groups, ke_valid] + per thermostat [M, x(M) v(M) q(M) s ke_cached s_pending spare] + group%N per group
static const double STATE_MAGIC = 20240731.0;
static size_t state_doubles(pfmds_ctx* c) {
    size_t n = 4 + c->groups.size();
    for (auto& t : c->nhc) n += 1 + (size_t)3 * t.M + 4;
    if (c->finalized && !c->zero_all) n += 3 * (size_t)c->N;  // accumulated forces, see pfmds_save_state
    return n;
}
int pfmds_state_size(pfmds_ctx* c, long long* n) {
    if (!c || !n) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        if (c->slab) fail(PFMDS_ERR_UNSUPPORTED, "unsupported: checkpoints of a slab context");
        CK(cudaSetDevice(c->dev));
        finalize(c);
        *n = (long long)state_doubles(c);
    });
}
int pfmds_save_state(pfmds_ctx* c, double* blob) {
    if (!c || !blob) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        CK(cudaSetDevice(c->dev));
        if (c->slab) fail(PFMDS_ERR_UNSUPPORTED, "unsupported: checkpoints of a slab context");
        finalize(c);
        integ_flush_pending(c);  // the velocities a following pfmds_download returns carry every thermostat scaling
        CK(cudaStreamSynchronize(c->st));
        size_t k = 0;
        blob[k++] = STATE_MAGIC; blob[k++] = (double)c->nhc.size(); blob[k++] = (double)c->groups.size(); blob[k++] = c->nhc_ke_valid ? 1. : 0.;
        for (auto& t : c->nhc) {
            blob[k++] = (double)t.M;
            CK(cudaMemcpy(blob + k, t.state, sizeof(double) * ((size_t)3 * t.M + 4), cudaMemcpyDeviceToHost));
            k += (size_t)3 * t.M + 4;
        }
        for (size_t g = 0; g < c->groups.size(); ++g) blob[k++] = (double)c->cur_n[g];
        if (!c->zero_all) {
            // zero_forces only touches the all_atoms group (md_integrators.f90:147-163): forces of atoms outside it accumulate over
            // the run and cannot be recomputed from the positions, so they travel with the checkpoint (file order)
            const size_t n3 = 3 * (size_t)c->N;
            LAUNCH((k_download_gather), (c->N + 255) / 256, 256, c->st, c->N, c->orig, c->frc, c->io_stage);
            c->launches += 1;
            CK(cudaMemcpyAsync(blob + k, c->io_stage, sizeof(double) * n3, cudaMemcpyDeviceToHost, c->st));
        }
        check_device_error(c);
    });
}
int pfmds_restore_state(pfmds_ctx* c, const double* pos, const double* vel, const double* blob) {
    if (!c || !pos || !vel || !blob) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        CK(cudaSetDevice(c->dev));
        if (c->slab) fail(PFMDS_ERR_UNSUPPORTED, "unsupported: checkpoints of a slab context");
        finalize(c);
        size_t k = 0;
        if (blob[k++] != STATE_MAGIC || (size_t)blob[k] != c->nhc.size() || (size_t)blob[k + 1] != c->groups.size())
            fail(PFMDS_ERR_INVALID, "error: the checkpoint does not belong to this settings file (thermostats / groups differ)");
        k += 2;
        const bool ke_valid = blob[k++] != 0.;
        integ_flush_pending(c);
        {   // state in file order -> slots
            const size_t n3 = 3 * (size_t)c->N;
            double *dp = c->io_stage, *dv = c->io_stage + n3;
            CK(cudaMemcpyAsync(dp, pos, sizeof(double) * n3, cudaMemcpyHostToDevice, c->st));
            CK(cudaMemcpyAsync(dv, vel, sizeof(double) * n3, cudaMemcpyHostToDevice, c->st));
            LAUNCH((k_upload_scatter), (c->N + 255) / 256, 256, c->st, c->N, c->orig, dp, dv, c->pos, c->vel);
            c->launches += 1;
            CK(cudaStreamSynchronize(c->st));
        }
        for (auto& t : c->nhc) {
            if ((int)blob[k] != t.M) fail(PFMDS_ERR_INVALID, "error: the checkpoint does not belong to this settings file (chain length differs)");
            ++k;
            CK(cudaMemcpy(t.state, blob + k, sizeof(double) * ((size_t)3 * t.M + 4), cudaMemcpyHostToDevice));
            k += (size_t)3 * t.M + 4;
        }
        c->nhc_pending = false;
        c->nhc_ke_valid = ke_valid && c->nhc_fusable;
        for (size_t g = 0; g < c->groups.size(); ++g) {
            int n = (int)blob[k++];
            if (n < 0 || n > (int)c->groups[g].size()) fail(PFMDS_ERR_INVALID, "error: bad group size in the checkpoint");
            if (n != c->cur_n[g]) {
                if (!c->d_grank[g]) fail(PFMDS_ERR_INVALID, "error: the checkpoint changes a group that has no change entry");
                c->cur_n[g] = n;
                LAUNCH((k_group_resize), (c->N + 255) / 256, 256, c->st, c->N, c->orig, c->d_grank[g], 1u << g, n, c->gmask);
                c->launches += 1;
            }
        }
        // the lists and forces of the checkpointed step: rebuilt from the checkpointed positions, which is what the interrupted
        // run held when every update_period divides that step (the host only writes checkpoints on such steps)
        for (auto& it : c->inter) for (int j = 0; j < it.nl_n; ++j) it.nl[j].built = false;
        update_lists(c, 0);
        compute_forces(c, false);
        if (!c->zero_all) {  // the forces of the checkpointed step, including what had accumulated outside the all_atoms group
            const size_t n3 = 3 * (size_t)c->N;
            CK(cudaMemcpyAsync(c->io_stage, blob + k, sizeof(double) * n3, cudaMemcpyHostToDevice, c->st));
            LAUNCH((k_upload_forces), (c->N + 255) / 256, 256, c->st, c->N, c->orig, c->io_stage, c->frc);
            c->launches += 1;
        }
        CK(cudaGetLastError());
        check_device_error(c);
    });
}

int pfmds_pair_count(pfmds_ctx* c, int inter, int list, long long* pairs) {
    if (!c || !pairs) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        CK(cudaSetDevice(c->dev));
        if (!c->finalized || inter < 0 || inter >= (int)c->inter.size() || list < 0 || list >= c->inter[(size_t)inter].nl_n) fail(PFMDS_ERR_INVALID, "error: no such neighbour list");
        unsigned long long* d = nullptr;
        CK(cudaMalloc(&d, sizeof(unsigned long long)));
        CK(cudaMemsetAsync(d, 0, sizeof(unsigned long long), c->st));
        LAUNCH((k_sum_int), 256, 256, c->st, c->N, c->inter[(size_t)inter].nl[list].nnum, d);
        if (c->slab) slab_allreduce_sum_ll(c, d, 1);
        unsigned long long h = 0;
        CK(cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        cudaFree(d);
        *pairs = (long long)h;
    });
}

int pfmds_pair_count_within(pfmds_ctx* c, int inter, int list, double r, long long* pairs) {
    if (!c || !pairs) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        CK(cudaSetDevice(c->dev));
        if (!c->finalized || inter < 0 || inter >= (int)c->inter.size() || list < 0 || list >= c->inter[(size_t)inter].nl_n) fail(PFMDS_ERR_INVALID, "error: no such neighbour list");
        unsigned long long* d = nullptr;
        CK(cudaMalloc(&d, sizeof(unsigned long long)));
        CK(cudaMemsetAsync(d, 0, sizeof(unsigned long long), c->st));
        LAUNCH((k_count_within), (c->N + 255) / 256, 256, c->st, c->N, c->pos, c->inter[(size_t)inter].nl[list].view(c->stride), c->box, r * r, d);
        c->launches += 1;
        if (c->slab) slab_allreduce_sum_ll(c, d, 1);
        unsigned long long h = 0;
        CK(cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, c->st));
        CK(cudaStreamSynchronize(c->st));
        cudaFree(d);
        *pairs = (long long)h;
    });
}

int pfmds_set_profiling(pfmds_ctx* c, int on) {
    if (!c) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        CK(cudaSetDevice(c->dev));
        prof_flush(c);
        c->prof_on = on != 0;
        if (on) { for (int k = 0; k < 32; ++k) { c->prof_ms[k] = 0; c->prof_cnt[k] = 0; } }
    });
}
int pfmds_kernel_times(pfmds_ctx* c, int n, double* ms, long long* count) {
    if (!c) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        CK(cudaSetDevice(c->dev));
        prof_flush(c);
        for (int k = 0; k < n && k < KS_COUNT; ++k) { if (ms) ms[k] = c->prof_ms[k]; if (count) count[k] = c->prof_cnt[k]; }
    });
}
const char* pfmds_kernel_name(int k) {
    static const char* names[KS_COUNT] = {"nl_bin", "nl_build", "lj", "lj1g", "rjl_density", "rjl_force", "tb_bond", "tb_force", "cos_graphene",
                                          "cos_indirect", "cos_metal", "normals", "kick_drift", "kick", "nhc", "zero_forces", "other", "rebosc_force", "rebosc_energy"};
    return (k >= 0 && k < KS_COUNT) ? names[k] : "";
}

int pfmds_timer_start(pfmds_ctx* c) {
    if (!c) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        CK(cudaSetDevice(c->dev));
        if (!c->tm0) { CK(cudaEventCreate(&c->tm0)); CK(cudaEventCreate(&c->tm1)); }
        CK(cudaStreamSynchronize(c->st));
        CK(cudaEventRecord(c->tm0, c->st));
    });
}
int pfmds_timer_stop(pfmds_ctx* c, double* ms) {
    if (!c || !ms) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        CK(cudaSetDevice(c->dev));
        if (!c->tm0) fail(PFMDS_ERR_INVALID, "error: pfmds_timer_stop without pfmds_timer_start");
        CK(cudaEventRecord(c->tm1, c->st));
        CK(cudaEventSynchronize(c->tm1));
        float f = 0;
        CK(cudaEventElapsedTime(&f, c->tm0, c->tm1));
        *ms = f;
        check_device_error(c);
    });
}

#if defined(__CUDACC__) && defined(PFMDS_STAMPS)
// debug build only (tools/stamps_probe.py): bind / read the per-step time stamps of common.cuh
}  // extern "C"
void forces_stamps_bind(unsigned long long* p);
void integ_stamps_bind(unsigned long long* p);
extern "C" {
static unsigned long long* g_stamps_dev = nullptr;
int pfmds_debug_stamps_begin(void) {
    const size_t n = 1 + (size_t)STAMP_STEPS * 2 * STAMP_SLOTS;
    if (!g_stamps_dev && cudaMalloc(&g_stamps_dev, n * 8) != cudaSuccess) return PFMDS_ERR_CUDA;
    std::vector<unsigned long long> h(n, 0ull);
    for (size_t s = 0; s < STAMP_STEPS; ++s)
        for (int k = 0; k < STAMP_SLOTS; ++k) h[1 + s * 2 * STAMP_SLOTS + k] = ~0ull;
    cudaDeviceSynchronize();
    cudaMemcpy(g_stamps_dev, h.data(), n * 8, cudaMemcpyHostToDevice);
    forces_stamps_bind(g_stamps_dev);
    integ_stamps_bind(g_stamps_dev);
    cudaDeviceSynchronize();
    return PFMDS_OK;
}
int pfmds_debug_stamps_read(unsigned long long* out) {  // 1 + STAMP_STEPS * 2 * STAMP_SLOTS words
    cudaDeviceSynchronize();
    cudaMemcpy(out, g_stamps_dev, (1 + (size_t)STAMP_STEPS * 2 * STAMP_SLOTS) * 8, cudaMemcpyDeviceToHost);
    forces_stamps_bind(nullptr);
    integ_stamps_bind(nullptr);
    return PFMDS_OK;
}
#endif
// Max errors of the device elementary functions: [0] exp relative, [1] switch/sincos absolute,
// [2] rsqrt relative, [3] raw MUFU.RSQ64H seed relative.
int pfmds_selftest_math(int device, double err[4]) {
#ifndef __CUDACC__
    (void)device; (void)err;
    return PFMDS_ERR_UNSUPPORTED;  // host replay of the test suite: there is no device to test
#else
    try {
        CK(cudaSetDevice(device));
        double* d = nullptr;
        CK(cudaMalloc(&d, 4 * sizeof(double)));
        CK(cudaMemset(d, 0, 4 * sizeof(double)));
        LAUNCH((k_math_selftest), 256, 256, 0, 1 << 22, d);
        CK(cudaMemcpy(err, d, 4 * sizeof(double), cudaMemcpyDeviceToHost));
        cudaFree(d);
        return PFMDS_OK;
    } catch (...) { return PFMDS_ERR_CUDA; }
#endif
}

// The short forms used by the second-generation rjl kernels: [0] exp_m / exp_m2 relative on [-40, 40], [1] cos_switch_m /
// half_switch absolute, [2] rsqrt_q relative, [3] exp_m relative on [-600, 600].
int pfmds_selftest_math2(int device, double err[4]) {
#ifndef __CUDACC__
    (void)device; (void)err;
    return PFMDS_ERR_UNSUPPORTED;
#else
    try {
        CK(cudaSetDevice(device));
        double* d = nullptr;
        CK(cudaMalloc(&d, 4 * sizeof(double)));
        CK(cudaMemset(d, 0, 4 * sizeof(double)));
        LAUNCH((k_math_selftest2), 256, 256, 0, 1 << 22, d);
        CK(cudaMemcpy(err, d, 4 * sizeof(double), cudaMemcpyDeviceToHost));
        cudaFree(d);
        return PFMDS_OK;
    } catch (...) { return PFMDS_ERR_CUDA; }
#endif
}

// FP64 FMA peak (TFLOP/s) and device-to-device copy bandwidth (GB/s, read+write) of `device`.
int pfmds_measure_peaks(int device, double* dfma_tflops, double* copy_gbs) {
#ifndef __CUDACC__
    (void)device; (void)dfma_tflops; (void)copy_gbs;
    return PFMDS_ERR_UNSUPPORTED;
#else
    try {
        CK(cudaSetDevice(device));
        cudaDeviceProp pr;
        CK(cudaGetDeviceProperties(&pr, device));
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        double* d = nullptr;
        CK(cudaMalloc(&d, 64));
        const int iters = 1 << 16, blocks = pr.multiProcessorCount * 8, threads = 256;
        double best = 0;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaEventRecord(e0));
            LAUNCH((k_dfma_peak), blocks, threads, 0, iters, 1.0 + rep, d);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) * 1e-12;
            if (rep > 0 && tf > best) best = tf;
        }
        if (dfma_tflops) *dfma_tflops = best;
        cudaFree(d);
        if (copy_gbs) {
            const size_t n = (size_t)1 << 25;  // 2 x 1 GiB buffers of double4
            double4 *a = nullptr, *b = nullptr;
            CK(cudaMalloc(&a, n * sizeof(double4))); CK(cudaMalloc(&b, n * sizeof(double4)));
            CK(cudaMemset(a, 0, n * sizeof(double4)));
            double bw = 0;
            for (int rep = 0; rep < 5; ++rep) {
                CK(cudaEventRecord(e0));
                LAUNCH((k_copy), pr.multiProcessorCount * 16, 512, 0, n, a, b);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms = 0;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                double g = 2.0 * n * sizeof(double4) / (ms * 1e-3) * 1e-9;
                if (rep > 0 && g > bw) bw = g;
            }
            *copy_gbs = bw;
            cudaFree(a); cudaFree(b);
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        return PFMDS_OK;
    } catch (...) { return PFMDS_ERR_CUDA; }
#endif
}


// ---- slab decomposition (BASELINE.json configs[3]) ----------------------------------------------------
int pfmds_slab_unique_id(char id[128]) {
    try { return slab_unique_id(id); } catch (...) { return PFMDS_ERR_CUDA; }
}

int pfmds_create_slab(pfmds_ctx** out, int device, int rank, int nranks, const char id[128], long long n_global, int n_local, const int* global_index,
                      const double* pos, const double* vel, const double* mass, const unsigned int* group_mask, int n_groups,
                      const long long* group_sizes, const double box[3], int capacity) {
    if (!out) return PFMDS_ERR_INVALID;
    pfmds_ctx* c = new pfmds_ctx;
    *out = c;
    return guarded(c, [&] {
        if (n_local < 0 || capacity < n_local || nranks < 2 || rank < 0 || rank >= nranks || !box || n_groups < 1 || n_groups > 31)
            fail(PFMDS_ERR_INVALID, "error: bad arguments to pfmds_create_slab");
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) fail(PFMDS_ERR_CUDA, "no CUDA device: pfmds_b200 has no CPU fallback");
        if (device < 0 || device >= ndev) fail(PFMDS_ERR_INVALID, "error: CUDA device " + std::to_string(device) + " does not exist");
        c->dev = device;
        CK(cudaSetDevice(device));
        CK(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
        g_live_contexts[device & 63].fetch_add(1);
        c->counted = true;
        c->N = n_local;
        c->stride = ((size_t)capacity + 31) / 32 * 32;
        for (int k = 0; k < 3; ++k) { c->box.L[k] = box[k]; c->box.h[k] = 0.5 * box[k]; }
        const size_t S = c->stride;
        CK(cudaMalloc(&c->pos, sizeof(double4) * S)); CK(cudaMalloc(&c->pos2, sizeof(double4) * S));
        CK(cudaMalloc(&c->vel, sizeof(double4) * S)); CK(cudaMalloc(&c->vel2, sizeof(double4) * S));
        CK(cudaMalloc(&c->frc, sizeof(double4) * S));
        CK(cudaMalloc(&c->gmask, sizeof(uint32_t) * S)); CK(cudaMalloc(&c->gmask2, sizeof(uint32_t) * S));
        CK(cudaMalloc(&c->orig, sizeof(int) * S)); CK(cudaMalloc(&c->orig2, sizeof(int) * S));
        CK(cudaMalloc(&c->cell_atoms, sizeof(int) * S)); CK(cudaMalloc(&c->cid, sizeof(int) * S));
        CK(cudaMalloc(&c->posf, sizeof(float4) * S));
        size_t nparts = (S + 127) / 128 + RED_BLOCKS;
        CK(cudaMalloc(&c->part, sizeof(double) * 16 * nparts));
        CK(cudaMalloc(&c->red, sizeof(double) * 64));
        CK(cudaMalloc(&c->err, sizeof(int) * PFMDS_ERRW));
        CK(cudaMemset(c->err, 0, sizeof(int) * PFMDS_ERRW));
        CK(cudaMemset(c->frc, 0, sizeof(double4) * S));
        std::vector<double4> hp(S, make_double4(0, 0, 0, 0)), hv(S, make_double4(0, 0, 0, 1));
        std::vector<int> ho(S, 0);
        std::vector<uint32_t> hm(S, 0u);
        for (int i = 0; i < n_local; ++i) {
            hp[(size_t)i] = make_double4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], 0.);
            hv[(size_t)i] = make_double4(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2], mass[i]);
            ho[(size_t)i] = global_index[i] - 1;
            hm[(size_t)i] = group_mask[i] & 0x7fffffffu;
        }
        CK(cudaMemcpy(c->pos, hp.data(), sizeof(double4) * S, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->vel, hv.data(), sizeof(double4) * S, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->orig, ho.data(), sizeof(int) * S, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->gmask, hm.data(), sizeof(uint32_t) * S, cudaMemcpyHostToDevice));
        c->group_count.assign(group_sizes, group_sizes + n_groups);
        read_env(c, n_local, true);
        CK(cudaEventCreate(&c->ev0)); CK(cudaEventCreate(&c->ev1));
        slab_init(c, rank, nranks, id, n_global, n_local, capacity);
    });
}

int pfmds_slab_counts(pfmds_ctx* c, int* n_local, int* n_ghost) {
    if (!c || !c->slab) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        CK(cudaSetDevice(c->dev));
        CK(cudaStreamSynchronize(c->st));
        if (n_local) *n_local = slab_n_local(c);
        if (n_ghost) *n_ghost = c->N - slab_n_local(c);
    });
}

int pfmds_slab_download(pfmds_ctx* c, int* n_local, int* global_index, double* pos, double* vel, double* frc) {
    if (!c || !c->slab) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        CK(cudaSetDevice(c->dev));
        integ_flush_pending(c);
        slab_download(c, n_local, global_index, pos, vel, frc);
        check_device_error(c);
    });
}

int pfmds_slab_upload(pfmds_ctx* c, int n_local, const double* pos, const double* vel) {
    if (!c || !c->slab) return PFMDS_ERR_INVALID;
    return guarded(c, [&] {
        CK(cudaSetDevice(c->dev));
        integ_flush_pending(c);
        c->nhc_ke_valid = false;
        c->energy_valid = false;
        slab_upload(c, n_local, pos, vel);
        for (auto& it : c->inter) for (int j = 0; j < it.nl_n; ++j) it.nl[j].built = false;
    });
}

int pfmds_timers(pfmds_ctx* c, double s[6]) {
    if (!c || !s) return PFMDS_ERR_INVALID;
    // slots: 0 pos_vel, 1 nlists, 2 nlsearch, 3 nldistance (no such pass on the device), 4 forces, 5 energy
    for (int k = 0; k < 6; ++k) s[k] = c->t_phase[k];
    return PFMDS_OK;
}
int pfmds_live_contexts(int device) { return g_live_contexts[device & 63].load(); }
int pfmds_launch_count(pfmds_ctx* c, long long* n) {
    if (!c || !n) return PFMDS_ERR_INVALID;
    *n = c->launches;
    return PFMDS_OK;
}
const char* pfmds_last_error(pfmds_ctx* c) { return c ? c->err_msg.c_str() : "null context"; }

int pfmds_destroy(pfmds_ctx* c) {
    if (!c) return PFMDS_OK;
    cudaSetDevice(c->dev);
    if (c->counted) g_live_contexts[c->dev & 63].fetch_sub(1);
    if (c->st) cudaStreamSynchronize(c->st);
    slab_destroy(c);
    for (auto& it : c->inter) {
        for (int j = 0; j < 3; ++j) { cudaFree(it.nl[j].nlist); cudaFree(it.nl[j].nlist_alt); cudaFree(it.nl[j].nnum); }
        cudaFree(it.aux); cudaFree(it.aux2); cudaFree(it.fpart); cudaFree(it.gnorm); cudaFree(it.tvec);
    }
    for (auto& t : c->nhc) cudaFree(t.state);
    for (int* r : c->d_grank) cudaFree(r);
    for (auto& g : c->graphs) cudaGraphExecDestroy(g.exec);
    for (auto b : c->fbuf) cudaFree(b);
    cudaFree(c->ticket);
    if (c->aux_ev_mid) cudaEventDestroy(c->aux_ev_mid);
    for (auto s : c->aux_st) cudaStreamDestroy(s);
    for (auto e : c->aux_ev) cudaEventDestroy(e);
    void* ptrs[] = {c->pos, c->pos2, c->vel, c->vel2, c->frc, c->gmask, c->gmask2, c->orig, c->orig2, c->cell_cnt, c->cell_start, c->cell_atoms,
                    c->cid, c->posf, c->scan_tmp, c->part, c->red, c->energy, c->err, c->logbuf, c->io_stage};
    for (void* p : ptrs) cudaFree(p);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->tm0) { cudaEventDestroy(c->tm0); cudaEventDestroy(c->tm1); }
    for (auto e : c->prof_ev) cudaEventDestroy(e);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
    return PFMDS_OK;
}
#ifdef __CUDACC__
const char* pfmds_version(void) { return "pfmds_b200 0.1 (sm_100a)"; }
#else
const char* pfmds_version(void) { return "pfmds_b200 0.1 HOST REPLAY of the device code (test suite only, not a product path)"; }
#endif

}  // extern "C"
