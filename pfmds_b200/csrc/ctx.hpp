// pfmds_b200 — host-side context behind the opaque pfmds_ctx of include/pfmds_b200.h, and the
// launch wrappers the C ABI calls (defined in nl.cu, forces.cu, integrate.cu).
#pragma once
#ifdef __CUDACC__
#include <cuda_runtime.h>
#endif

#include <string>
#include <vector>

#include "common.cuh"

struct NList {
    int g1 = 0, g2 = 0, maxn = 0, period = 1;
    double rcut = 0;
    int* nlist = nullptr;  // ELL [maxn][stride]
    int* nnum = nullptr;   // [stride]
    bool from_tb = false;  // ljc/morsec nl(3) derived from the first tb list (graphenenorm.f90:58-71)
    int src_inter = -1;
    bool built = false;
    // rows are re-ordered after every build into [r < R1 | R1 <= r < R2 | r >= R2] (distances at build time)
    bool partition = false;
    double part_r1sq = 0, part_r2sq = 0;
    int* nlist_alt = nullptr;
    ListView view(size_t stride) const { return ListView{nlist, nnum, stride}; }
};

struct Inter {
    int kind = -1;
    std::string name;
    int nl_n = 0;
    NList nl[3];
    LJp lj{}; LJ1Gp lj1g{}; LJCp ljc{}; MORp mor{}; TBp tb{}; RJLp rjl{}; REBp reb{};
    double* aux = nullptr;    // tb: bond orders B, ELL [maxn][stride];  rjl: node table of the third-generation routines (double2 entries; 1/Eb lives in pos[].w)
    double* aux2 = nullptr;   // tb: B^(1/delt+1)
    double4* fpart = nullptr; // tb: per-(slot, atom) force contributions, summed per atom in slot order
    double4* gnorm = nullptr; // ljc/morsec: unit normal per carbon atom {nx,ny,nz,-}
    double4* tvec = nullptr;  // ljc/morsec: T_i = sum_p V2 V3 f_c/(n_i.dr) dr  (normal-derivative term)
};

struct Nhc {
    int group = 0, M = 0, L = 0;
    double temperature = 0;
    double* state = nullptr;  // device: x[M], v[M], q[M], then s, e
};

struct Slab;

struct pfmds_ctx {
    int dev = 0;
    cudaStream_t st = nullptr;
    int N = 0;
    size_t stride = 0;  // N rounded up to 32
    BoxD box{};
    // state, double-buffered for the cell re-sort
    double4 *pos = nullptr, *vel = nullptr, *frc = nullptr, *pos2 = nullptr, *vel2 = nullptr;
    uint32_t *gmask = nullptr, *gmask2 = nullptr;
    float4* posf = nullptr;  // float copy of the positions + group mask, for the list-build prefilter
    int *orig = nullptr, *orig2 = nullptr;
    // cell grid
    int ncell[3]{1, 1, 1};
    int ncells = 0;
    double cell_rc = 0;
    int *cell_cnt = nullptr, *cell_start = nullptr, *cell_atoms = nullptr, *cid = nullptr, *scan_tmp = nullptr;
    bool identity_order = false;  // cell_atoms[k]==k (atoms physically in cell order)
    // reductions
    double* part = nullptr;   // block partials of the two-stage reductions
    size_t part_cap = 0;
    double* red = nullptr;    // [64] reduced values
    double* energy = nullptr; // [n_inter] device energies
    double* logbuf = nullptr; // device-resident energy log of pfmds_advance_logged
    size_t log_cap = 0;
    double* io_stage = nullptr; // file-order staging block of pfmds_download / pfmds_upload
    size_t io_cap = 0;
    int* err = nullptr;       // [PFMDS_ERRW]
    // host description
    std::vector<std::vector<int>> groups;  // 1-based group -> 1-based file indexes
    std::vector<uint32_t> h_gmask;         // by file index
    bool all_in_group(int g) const { return g >= 1 && g <= (int)groups.size() && cur_n.empty() ? groups[(size_t)g - 1].size() == (size_t)N : (g >= 1 && g <= (int)groups.size() && groups[(size_t)g - 1].size() == (size_t)N && changes.empty()); }
    int all_moving = 1, xyz_moving = 1, z_moving = 1, all_atoms = 1;
    int zero_momentum_period = 1;
    bool invert_z = false;
    // deposition (change_particle_group_N, md_general.f90:82-94): group `to` exposes the first cur_n[to] atoms of its index list
    struct GroupChange { int from, to, ts1, ts2, frec; };
    std::vector<GroupChange> changes;
    std::vector<int> cur_n;       // group%N of every group; differs from groups[g].size() only for the targets of change entries
    std::vector<int*> d_grank;    // per group, targets only: device int[N] by FILE index = position in the group's index list (INT_MAX: not listed)
    bool zero_all = true;         // the all_atoms group is every atom at every step: zero_forces is one memset
    std::vector<Inter> inter;
    std::vector<Nhc> nhc;
    // slab decomposition (slab.cu); null in single-GPU and ensemble runs
    Slab* slab = nullptr;
    int* newslot = nullptr;  // old slot -> new slot of the last cell re-sort
    std::vector<long long> group_count;  // slab mode: global size of every group
    // CUDA graphs of the steady-state step (small systems are launch-latency bound), keyed by what is baked in
    struct StepGraph { int kind; double dt; const void* pos; bool pending, ke_valid, opened, pre_open, alone; cudaGraphExec_t exec; long long launches; int nsteps; bool rebuild; };
    std::vector<StepGraph> graphs;
    bool use_graphs = false;
    bool graph_rebuilds = true;     // steps that rebuild EVERY list (cell re-sort included) are replayed from graphs too, one per direction of the state's double buffer (PFMDS_GRAPH_REBUILDS=0: launched one by one)
    int graph_steps = 4;            // steady-state steps per graph launch where a run of them allows it (PFMDS_GRAPH_STEPS; graph-to-graph gap 4.4 us, node-to-node 2.3 us: tools/stamps_probe.py)
    int rjl_gen = 2;                // rjl pair routines: 2 = analytic short forms (default), 3 = node-table exponentials (measured slower: L1-bound), 1 = first generation (PFMDS_RJL_GEN)
    bool nl_mask = true;            // thread-per-atom list build with the FP32 prefilter and the exact test in separate loops (measured 8 % faster, BENCH_r01); PFMDS_NL_MASK=0: k_build
    bool lj1g_pipe = true;          // pipelined lj1g force kernel for systems of small_n atoms and more (measured 0.174 -> 0.102 ms, BENCH_r01); PFMDS_LJ1G_PIPE=0: k_lj1g
    // Path switches by system size.  Runtime fields (PFMDS_SMALL_N, PFMDS_NL_WARP_N) so that the parity tests can drive
    // the kernels of BOTH sides of each switch against the oracle on systems the O(N^2) oracle can handle.
    bool nl_cell = true;            // cell-tiled list build (one warp per cell, candidates staged in shared memory by bulk copies) for large systems; PFMDS_NL_CELL=0: k_build_mask
    int small_n = 100000;           // below: 8 lanes per atom in the pair kernels (latency bound); from it on: thread per atom, pipelined
    int nl_warp_n = 200000;         // below: warp-per-atom list build; from it on: thread per atom
    // Small systems (latency bound): the interactions of a step run as parallel branches of its CUDA graph, each accumulating into
    // its own force buffer (fbuf[k]); integrate.cu k_sum_kick_ke adds the buffers per atom in file order (the bits of the sequential
    // accumulation), applies the closing kick and clears them.  fst / fout redirect the launches of forces_interaction.
    std::vector<double4*> fbuf;
    std::vector<cudaStream_t> aux_st;
    std::vector<cudaEvent_t> aux_ev;     // [0] fork, [1 + k] join of branch k
    cudaEvent_t aux_ev_mid = nullptr;    // inside an interaction: normals done -> the metal-side branch of ljc / morsec may start
    cudaStream_t fst = nullptr, fst2 = nullptr;   // (fst2 / fout2: the converse-list launch of lj)
    double4 *fout = nullptr, *fout2 = nullptr;
    unsigned int* ticket = nullptr;      // block counter of k_sum_kick_ke (its last block closes the thermostat step)
    bool fbuf_on = false;                // buffers exist (finalize): small system, at most 8 interactions, PFMDS_SMALL_FORK != 0
    bool fbuf_active = false;            // this step's forces went into the buffers: the sum kernel must run
    bool first_overwrites = false;  // interaction 0 is rjl (or lj1g with the pipelined kernel) and owns every atom: its force kernel stores, no zero pass
    bool energy_valid = false;  // c->energy[] holds the potential energies of the current positions (computed inside the last step)
    bool finalized = false;
    bool counted = false;       // this context is in the process-wide count of live contexts of its device (capi.cu)
    // fused NVT path (integrate.cu): usable when the thermostat groups are pairwise disjoint
    bool nhc_fusable = false;
    bool nhc_ke_valid = false;  // state[3M+1] holds the current kinetic energy of each thermostat group
    bool nhc_pending = false;   // state[3M+2] holds a velocity scale that has not been applied yet
    bool pre_open_enabled = true;  // PFMDS_PRE_OPEN=0: k_nhc_open stays a launch of its own
    bool pre_open = false;      // set per step by the caller: the next NVT step follows in the same call and nothing reads the chains before it
    bool nhc_opened = false;    // the pending scale already contains the opening half step of the next step (k_nhc_close also_open)
    long long launches = 0;
    std::string err_msg;
    // phase timers (PFMDS_TIMERS=1)
    bool timers_on = false;
    double t_phase[6]{0, 0, 0, 0, 0, 0};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t tm0 = nullptr, tm1 = nullptr;  // pfmds_timer_start / pfmds_timer_stop
    // per-kernel profiling
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;  // pairs
    std::vector<int> prof_slot;
    size_t prof_used = 0;
    double prof_ms[32]{};
    long long prof_cnt[32]{};
};

// per-kernel device timing (pfmds_set_profiling): CUDA events recorded on the context's stream around
// every launch of a kernel class, accumulated at synchronisation points
enum { KS_NL_BIN = 0, KS_NL_BUILD, KS_LJ, KS_LJ1G, KS_RJL_DENSITY, KS_RJL_FORCE, KS_TB_BOND, KS_TB_FORCE, KS_COS_GRAPHENE, KS_COS_INDIRECT,
       KS_COS_METAL, KS_NORMALS, KS_KICK_DRIFT, KS_KICK, KS_NHC, KS_ZERO_FORCES, KS_OTHER, KS_REBOSC_FORCE, KS_REBOSC_ENERGY, KS_COUNT };

void prof_flush(pfmds_ctx* c);
struct KTimer {
    pfmds_ctx* c;
    int slot;
    KTimer(pfmds_ctx* c_, int slot_);
    ~KTimer();
};

#define RED_BLOCKS 592  // 148 SMs x 4

// ---- nl.cu ----
void nl_setup_grid(pfmds_ctx* c);
void nl_bin_atoms(pfmds_ctx* c, bool reorder);
void nl_build(pfmds_ctx* c, NList& l);
void nl_nearest3_from(pfmds_ctx* c, NList& nn, const NList& src);

// ---- slab.cu ----
int slab_unique_id(char* id128);
void slab_init(pfmds_ctx* c, int rank, int nranks, const char* id128, long long n_global, int n_local, int capacity);
void slab_destroy(pfmds_ctx* c);
void slab_redistribute(pfmds_ctx* c);
void slab_after_reorder(pfmds_ctx* c);
void slab_exchange(pfmds_ctx* c, int field);
void slab_step_done(pfmds_ctx* c);
bool slab_uses_p2p(pfmds_ctx* c);
bool slab_fused(pfmds_ctx* c);
// stage 0: kick+drift pushes positions; 1: rjl density waits for positions, pushes 1/Eb; 2: rjl force waits for 1/Eb
SlabDev slab_dev(pfmds_ctx* c, int stage);
bool slab_pos_pushed_by_kick(pfmds_ctx* c, bool rebuild_step);
void slab_allreduce_sum(pfmds_ctx* c, double* d, int n);
bool slab_ke_close(pfmds_ctx* c, const NhcPack& P, int nparts, const double* part, double ts2, double ts3, double ts4);  // thermostat KE over the ranks by peer-memory mailboxes + chain update
void slab_allreduce_max(pfmds_ctx* c, double* d, int n);
void slab_allreduce_max_int(pfmds_ctx* c, int* d, int n);
void slab_allreduce_sum_ll(pfmds_ctx* c, unsigned long long* d, int n);
void slab_download(pfmds_ctx* c, int* n_local, int* gid, double* pos, double* vel, double* frc);  // owned atoms, compacted on the device
void slab_upload(pfmds_ctx* c, int n_local, const double* pos, const double* vel);
int slab_rank(pfmds_ctx* c);
int slab_nranks(pfmds_ctx* c);
int slab_n_local(pfmds_ctx* c);
long long slab_n_global(pfmds_ctx* c);

// ---- forces.cu ----
void forces_zero(pfmds_ctx* c);
void rjl_prepare(pfmds_ctx* c, Inter& it);  // third-generation rjl: node table of the exponentials, built once
void forces_interaction(pfmds_ctx* c, int k, bool with_energy);
void normals_interaction(pfmds_ctx* c, int k, cudaStream_t st = nullptr);
void energy_interaction(pfmds_ctx* c, int k);  // result in c->energy[k]

// ---- rebosc.cu ----
void rebosc_forces(pfmds_ctx* c, Inter& it);          // numerical forces, md_interactions.f90:273-311
int rebosc_energy_partials(pfmds_ctx* c, Inter& it);  // REBOsc_energy block partials into c->part; returns their number

// ---- integrate.cu ----
void integ_check_positions(pfmds_ctx* c);
void integ_invert_z(pfmds_ctx* c);
void integ_nhc_half(pfmds_ctx* c, Nhc& t, double dt);
void integ_kick_drift(pfmds_ctx* c, double dt);
void integ_kick(pfmds_ctx* c, double dt);
void integ_quench(pfmds_ctx* c);
void integ_zero_momentum(pfmds_ctx* c);
void integ_nvt_open_kick_drift(pfmds_ctx* c, double dt, bool rebuild_step);
void integ_nvt_kick_close(pfmds_ctx* c, double dt);
void integ_sum_forces(pfmds_ctx* c, int mode, double dt);   // fbuf mode: per-atom sum of the interaction buffers; mode 0 sum, 1 + closing kick, 2 + thermostat KE partials and chain update
void integ_flush_pending(pfmds_ctx* c);
// out[0]=KE(group) ; group sums for diagnostics: out[0..2]=sum F, [3..5]=sum m x, [6..8]=sum m v, [9]=sum m, [10]=max v^2
void integ_kinetic_energy(pfmds_ctx* c, int group, double* d_out);
void integ_diagnostics(pfmds_ctx* c, double* d_out);
