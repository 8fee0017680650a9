// pfmds_b200 — cell-binned Verlet neighbour-list build (replaces the reference's O(N1*N2) search,
// code_source/MOLECULAR_DYNAMICS/md_neighbours.f90:56-100, its per-step distance refresh :104-124 and
// the serial converse list :128-160).
//
// Membership rule reproduced bit-for-bit: global i != j and dx*dx+dy*dy+dz*dz < r_cut*r_cut with the
// products and sums rounded separately (no FMA contraction) on the min-image dr of find_distance
// (md_general.f90:423-441).  Lists are rebuilt only when mod(md_step,update_period)==0, exactly like
// update_neighbour_list (:32-52); between rebuilds membership is frozen and distances are recomputed
// in registers by the force kernels, so no "distance" or "converse" pass exists here.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#if defined(__CUDACC__) || defined(PFMDS_EMU_LIB)
#include "ctx.hpp"
#else
#include "common.cuh"  // host emulation of the kernels (tests/nl_host.cpp): no context, no launchers
#endif

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw std::string("CUDA: ") + cudaGetErrorString(e_) + " at " #x; } while (0)

// ------------------------------------------------------------------------------------------------
// Cell grid for a box and the largest r_cut of the run (pure arithmetic: shared by nl_setup_grid and the host emulation tests).
// Positions may sit up to 1e-7 outside [0,L] (check_positions tolerance, md_general.f90:346) and are
// clamped into the edge cells, so the cells are made a little wider than r_cut.
inline long long nl_grid_dims(const BoxD& box, double rc, int ncell[3], double& cell_rc) {
    if (rc <= 0) rc = 1.0;
    cell_rc = rc * (1.0 + 1e-9) + 1e-5;
    long long total = 1;
    for (int k = 0; k < 3; ++k) {
        int n = (int)std::floor(box.L[k] / cell_rc);
        if (n < 1) n = 1;
        if (n > 1024) n = 1024;
        ncell[k] = n;
        total *= n;
    }
    // keep the cell table bounded for sparse, huge boxes
    while (total > 64ll * 1024 * 1024) {
        int k = (ncell[0] >= ncell[1] && ncell[0] >= ncell[2]) ? 0 : (ncell[1] >= ncell[2] ? 1 : 2);
        total /= ncell[k];
        ncell[k] = (ncell[k] + 1) / 2;
        total *= ncell[k];
    }
    return total;
}
#ifdef PFMDS_HAVE_CTX
void nl_setup_grid(pfmds_ctx* c) {
    double rc = 0;
    for (auto& it : c->inter)
        for (int j = 0; j < it.nl_n; ++j) rc = std::max(rc, it.nl[j].rcut);
    c->ncells = (int)nl_grid_dims(c->box, rc, c->ncell, c->cell_rc);
    if (c->cell_cnt) { cudaFree(c->cell_cnt); cudaFree(c->cell_start); cudaFree(c->scan_tmp); }
    CK(cudaMalloc(&c->cell_cnt, sizeof(int) * (size_t)(c->ncells + 1)));
    CK(cudaMalloc(&c->cell_start, sizeof(int) * (size_t)(c->ncells + 1)));
    CK(cudaMalloc(&c->scan_tmp, sizeof(int) * (size_t)(c->ncells / 2048 + 2)));
}
#endif

struct GridD { int n[3]; double inv[3]; };
__global__ void k_make_posf(int N, const double4* __restrict__ pos, const uint32_t* __restrict__ gmask, float4* __restrict__ posf);

__device__ __forceinline__ int cell_coord(double x, double inv, int n) {
    int c = (int)floor(x * inv);
    return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

__global__ void k_cell_count(int N, const double4* __restrict__ pos, GridD g, int* __restrict__ cid, int* __restrict__ cnt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double4 p = pos[i];
    int cx = cell_coord(p.x, g.inv[0], g.n[0]), cy = cell_coord(p.y, g.inv[1], g.n[1]), cz = cell_coord(p.z, g.inv[2], g.n[2]);
    int c = (cz * g.n[1] + cy) * g.n[0] + cx;
    cid[i] = c;
    atomicAdd(&cnt[c], 1);
}

#ifdef PFMDS_COOP  // the scans exchange data between lanes: device, or the lock-step host replay (the serial replay uses a plain prefix sum)
// exclusive scan, 2048 items per block (1024 threads x 2), block totals to `sums`
__global__ void k_scan_block(int n, const int* __restrict__ in, int* __restrict__ out, int* __restrict__ sums) {
    __shared__ int sh[32];
    int base = blockIdx.x * 2048 + threadIdx.x * 2;
    int a = base < n ? in[base] : 0, b = base + 1 < n ? in[base + 1] : 0;
    int v = a + b, incl = v;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) sh[w] = incl;
    __syncthreads();
    if (w == 0) {
        int s = sh[lane], si = s;
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, si, o); if (lane >= o) si += t; }
        sh[lane] = si - s;
        if (lane == 31) sums[blockIdx.x] = si;
    }
    __syncthreads();
    int excl = incl - v + sh[w];
    if (base < n) out[base] = excl;
    if (base + 1 < n) out[base + 1] = excl + a;
}
__global__ void k_scan_sums(int nb, int* sums) {  // single block, serial over chunks of 1024
    __shared__ int carry;
    __shared__ int sh[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < nb ? sums[i] : 0, incl = v;
        int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) sh[w] = incl;
        __syncthreads();
        if (w == 0) {
            int s = sh[lane], si = s;
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, si, o); if (lane >= o) si += t; }
            sh[lane] = si - s;
        }
        __syncthreads();
        int excl = incl - v + sh[w] + carry;
        if (i < nb) sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
}
__global__ void k_scan_add(int n, int* out, const int* __restrict__ sums, int total_slot, int total) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += sums[i / 2048];
    if (i == 0) out[total_slot] = total;
}
#endif

__global__ void k_cell_scatter(int N, const int* __restrict__ cid, const int* __restrict__ start, int* __restrict__ cursor, int* __restrict__ atoms) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int c = cid[i];
    atoms[start[c] + atomicAdd(&cursor[c], 1)] = i;
}
// deterministic order inside each cell: ascending FILE index (insertion sort, cells hold tens of atoms).  The key is the
// atom's identity, not the slot it happens to sit in, so the cell order — and with it every row order and summation order —
// is a function of the positions alone: a run restarted from a checkpoint reproduces the interrupted run bit for bit.
__global__ void k_cell_sort(int ncells, const int* __restrict__ start, int* __restrict__ atoms, const int* __restrict__ orig) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    int b = start[c], e = start[c + 1];
    for (int i = b + 1; i < e; ++i) {
        int v = atoms[i], kv = orig[v], k = i - 1;
        while (k >= b && orig[atoms[k]] > kv) { atoms[k + 1] = atoms[k]; --k; }
        atoms[k + 1] = v;
    }
}
__global__ void k_permute(int N, const int* __restrict__ order, const double4* __restrict__ pos, const double4* __restrict__ vel,
                          const uint32_t* __restrict__ gm, const int* __restrict__ orig, double4* __restrict__ pos2, double4* __restrict__ vel2,
                          uint32_t* __restrict__ gm2, int* __restrict__ orig2, int* __restrict__ newslot) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    int s = order[k];
    if (newslot) newslot[s] = k;
    pos2[k] = pos[s];
    vel2[k] = vel[s];
    gm2[k] = gm[s];
    orig2[k] = orig[s];
}
__global__ void k_iota(int N, int* a) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < N) a[k] = k;
}

#ifdef PFMDS_HAVE_CTX
// Bin every atom into the cell grid; with `reorder` the state arrays are physically permuted into
// cell order (only legal when every neighbour list is rebuilt in the same step, since lists hold
// slot indices).
void nl_bin_atoms(pfmds_ctx* c, bool reorder) {
    const int N = c->N, T = 256, nb = (N + T - 1) / T;
    KTimer kt(c, KS_NL_BIN);
    GridD g;
    for (int k = 0; k < 3; ++k) { g.n[k] = c->ncell[k]; g.inv[k] = c->ncell[k] / c->box.L[k]; }
    CK(cudaMemsetAsync(c->cell_cnt, 0, sizeof(int) * (size_t)(c->ncells + 1), c->st));
    LAUNCH((k_cell_count), nb, T, c->st, N, c->pos, g, c->cid, c->cell_cnt);
#ifdef PFMDS_COOP
    int sb = (c->ncells + 2047) / 2048;
    LAUNCH((k_scan_block), sb, 1024, c->st, c->ncells, c->cell_cnt, c->cell_start, c->scan_tmp);
    LAUNCH((k_scan_sums), 1, 1024, c->st, sb, c->scan_tmp);
    LAUNCH((k_scan_add), (c->ncells + T - 1) / T, T, c->st, c->ncells, c->cell_start, c->scan_tmp, c->ncells, N);
#else  // host replay: the scans exchange data between lanes; same exclusive prefix sum, serially
    c->cell_start[0] = 0;
    for (int k = 0; k < c->ncells; ++k) c->cell_start[k + 1] = c->cell_start[k] + c->cell_cnt[k];
#endif
    CK(cudaMemsetAsync(c->cell_cnt, 0, sizeof(int) * (size_t)(c->ncells + 1), c->st));
    LAUNCH((k_cell_scatter), nb, T, c->st, N, c->cid, c->cell_start, c->cell_cnt, c->cell_atoms);
    LAUNCH((k_cell_sort), (c->ncells + T - 1) / T, T, c->st, c->ncells, c->cell_start, c->cell_atoms, c->orig);
    c->launches += 6;
    if (reorder) {
        LAUNCH((k_permute), nb, T, c->st, N, c->cell_atoms, c->pos, c->vel, c->gmask, c->orig, c->pos2, c->vel2, c->gmask2, c->orig2, c->newslot);
        std::swap(c->pos, c->pos2);
        std::swap(c->vel, c->vel2);
        std::swap(c->gmask, c->gmask2);
        std::swap(c->orig, c->orig2);
        LAUNCH((k_iota), nb, T, c->st, N, c->cell_atoms);
        c->identity_order = true;
        c->launches += 2;
    } else {
        c->identity_order = false;
    }
    LAUNCH((k_make_posf), nb, T, c->st, N, c->pos, c->gmask, c->posf);
    c->launches += 1;
    CK(cudaGetLastError());
}
#endif

// ------------------------------------------------------------------------------------------------
// One thread per list-owner atom walks the 27 (or fewer, for boxes under three cells wide) cells
// around it.  Candidates first pass an FP32 prefilter on a float4 copy of the positions that also carries
// the group mask (one 16-byte load per candidate, FP32 pipe): r2f < r_cut^2 + margin with the margin
// covering every float rounding, so no true neighbour is ever rejected.  Survivors (~15 %) take the exact
// FP64 test: dr2 uses __dmul_rn/__dadd_rn so the compiler cannot contract it into FMAs and the set
// {j : dr2 < rc2} is bit-identical to the reference's brute-force scan (md_neighbours.f90:73-87).
// With PART the accepted entries are written class by class (r < R1 | switch zone | r >= R2, distances at
// build time) so that the lanes of a warp later agree on the branch the potential takes: class 0 grows
// from the front of the row, classes 1 and 2 are staged in a scratch row and appended.
struct PrefD { float L[3], h[3], lim; int on; };
__global__ void k_make_posf(int N, const double4* __restrict__ pos, const uint32_t* __restrict__ gmask, float4* __restrict__ posf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double4 p = pos[i];
    posf[i] = make_float4((float)p.x, (float)p.y, (float)p.z, __uint_as_float(gmask[i]));
}

// Thread-per-atom variant (large systems: the grid already fills the GPU and a serial scan issues fewer
// warp instructions than 32 lanes sharing ~23-atom cells).
template <bool IDENT, bool PART>
__device__ __forceinline__ void build_row(int i, const double4* __restrict__ pos, const float4* __restrict__ posf, const int* __restrict__ orig,
                                          const int* __restrict__ cstart, const int* __restrict__ catoms, const GridD& g, const BoxD& box, const PrefD& pf,
                                          uint32_t bit1, uint32_t bit2, double rc2, double r1sq, double r2sq, int maxn, size_t stride,
                                          int* __restrict__ nlist, int* __restrict__ alt, int* __restrict__ nnum, int* err) {
    const float4 pif = posf[i];
    if ((__float_as_uint(pif.w) & (bit1 | PFMDS_GHOST)) != bit1) { nnum[i] = 0; return; }  // owners: in group 1 and not a ghost copy
    const double4 pi = pos[i];
    // same binning expression as k_cell_count
    const int cx = cell_coord(pi.x, g.inv[0], g.n[0]), cy = cell_coord(pi.y, g.inv[1], g.n[1]), cz = cell_coord(pi.z, g.inv[2], g.n[2]);
    const int lo0 = g.n[0] >= 3 ? -1 : 0, hi0 = g.n[0] >= 2 ? 1 : 0;
    const int lo1 = g.n[1] >= 3 ? -1 : 0, hi1 = g.n[1] >= 2 ? 1 : 0;
    const int lo2 = g.n[2] >= 3 ? -1 : 0, hi2 = g.n[2] >= 2 ? 1 : 0;
    int cnt = 0, c0 = 0, c1 = 0, c2 = 0;
    for (int oz = lo2; oz <= hi2; ++oz) {
        int z = cz + oz; z = z < 0 ? z + g.n[2] : (z >= g.n[2] ? z - g.n[2] : z);
        for (int oy = lo1; oy <= hi1; ++oy) {
            int y = cy + oy; y = y < 0 ? y + g.n[1] : (y >= g.n[1] ? y - g.n[1] : y);
            for (int ox = lo0; ox <= hi0; ++ox) {
                int x = cx + ox; x = x < 0 ? x + g.n[0] : (x >= g.n[0] ? x - g.n[0] : x);
                int cc = (z * g.n[1] + y) * g.n[0] + x;
                int b = cstart[cc], e = cstart[cc + 1];
                for (int s = b; s < e; ++s) {
                    int j = IDENT ? s : catoms[s];
                    const float4 q = posf[j];
                    if (j == i) continue;
                    if (!(__float_as_uint(q.w) & bit2)) continue;
                    if (pf.on) {
                        float fx = q.x - pif.x, fy = q.y - pif.y, fz = q.z - pif.z;
                        fx = fx >= pf.h[0] ? fx - pf.L[0] : (fx < -pf.h[0] ? fx + pf.L[0] : fx);
                        fy = fy >= pf.h[1] ? fy - pf.L[1] : (fy < -pf.h[1] ? fy + pf.L[1] : fy);
                        fz = fz >= pf.h[2] ? fz - pf.L[2] : (fz < -pf.h[2] ? fz + pf.L[2] : fz);
                        if (!(fmaf(fz, fz, fmaf(fy, fy, fx * fx)) < pf.lim)) continue;
                    }
                    double4 pj = pos[j];
                    double dx = min_image(pj.x - pi.x, box.h[0], box.L[0]);
                    double dy = min_image(pj.y - pi.y, box.h[1], box.L[1]);
                    double dz = min_image(pj.z - pi.z, box.h[2], box.L[2]);
                    double dr2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                    if (dr2 < rc2) {
                        if (cnt < maxn) {
                            if (!PART) nlist[(size_t)cnt * stride + i] = j;
                            else if (dr2 < r1sq) nlist[(size_t)(c0++) * stride + i] = j;
                            else if (dr2 < r2sq) alt[(size_t)(c1++) * stride + i] = j;
                            else alt[(size_t)(maxn - 1 - (c2++)) * stride + i] = j;
                        }
                        ++cnt;
                    }
                }
            }
        }
    }
    if (cnt > maxn) { raise_error(err, E_TOO_MANY, orig[i], cnt); cnt = maxn; }  // md_neighbours.f90:80
    if (PART) {
        for (int k = 0; k < c1; ++k) nlist[(size_t)(c0 + k) * stride + i] = alt[(size_t)k * stride + i];
        for (int k = 0; k < c2; ++k) nlist[(size_t)(c0 + c1 + k) * stride + i] = alt[(size_t)(maxn - 1 - k) * stride + i];
    }
    nnum[i] = cnt;
}
template <bool IDENT, bool PART>
__global__ void __launch_bounds__(128) k_build(int N, const double4* __restrict__ pos, const float4* __restrict__ posf, const int* __restrict__ orig,
                                               const int* __restrict__ cstart, const int* __restrict__ catoms, GridD g, BoxD box, PrefD pf,
                                               uint32_t bit1, uint32_t bit2, double rc2, double r1sq, double r2sq, int maxn, size_t stride,
                                               int* __restrict__ nlist, int* __restrict__ alt, int* __restrict__ nnum, int* err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    build_row<IDENT, PART>(i, pos, posf, orig, cstart, catoms, g, box, pf, bit1, bit2, rc2, r1sq, r2sq, maxn, stride, nlist, alt, nnum, err);
}

// Thread-per-atom variant with the two tests separated (the default; PFMDS_NL_MASK=0 selects k_build; 8 % faster per rebuild at 10^6 atoms, BENCH_r01).  In k_build the exact
// FP64 test sits inside the candidate loop, and because nearly every candidate survives the FP32 prefilter for SOME lane of the
// warp, the warp executes that block for all ~630 candidates of an atom with ~15 % of its lanes active.  Here the prefilter runs
// over one cell range (32 candidates at a time) and records its survivors in a register bit mask; the exact tests then walk the
// set bits, the lanes of the warp side by side: max-over-lanes popc(mask), about 6 trips per cell instead of ~23.  Candidate order
// is unchanged (ascending bit = ascending cell slot, ranges in order), so every row is k_build's row, entry for entry.
#ifdef __CUDA_ARCH__
#define PFMDS_FFS(m) __ffs((int)(m))
#else
#define PFMDS_FFS(m) __builtin_ffs((int)(m))
#endif
template <bool IDENT, bool PART>
__global__ void __launch_bounds__(128) k_build_mask(int N, const double4* __restrict__ pos, const float4* __restrict__ posf, const int* __restrict__ orig,
                                                    const int* __restrict__ cstart, const int* __restrict__ catoms, GridD g, BoxD box, PrefD pf,
                                                    uint32_t bit1, uint32_t bit2, double rc2, double r1sq, double r2sq, int maxn, size_t stride,
                                                    int* __restrict__ nlist, int* __restrict__ alt, int* __restrict__ nnum, int* err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float4 pif = posf[i];
    if ((__float_as_uint(pif.w) & (bit1 | PFMDS_GHOST)) != bit1) { nnum[i] = 0; return; }
    const double4 pi = pos[i];
    const int cx = cell_coord(pi.x, g.inv[0], g.n[0]), cy = cell_coord(pi.y, g.inv[1], g.n[1]), cz = cell_coord(pi.z, g.inv[2], g.n[2]);
    const int lo0 = g.n[0] >= 3 ? -1 : 0, hi0 = g.n[0] >= 2 ? 1 : 0;
    const int lo1 = g.n[1] >= 3 ? -1 : 0, hi1 = g.n[1] >= 2 ? 1 : 0;
    const int lo2 = g.n[2] >= 3 ? -1 : 0, hi2 = g.n[2] >= 2 ? 1 : 0;
    int cnt = 0, c0 = 0, c1 = 0, c2 = 0;
    for (int oz = lo2; oz <= hi2; ++oz) {
        int z = cz + oz; z = z < 0 ? z + g.n[2] : (z >= g.n[2] ? z - g.n[2] : z);
        for (int oy = lo1; oy <= hi1; ++oy) {
            int y = cy + oy; y = y < 0 ? y + g.n[1] : (y >= g.n[1] ? y - g.n[1] : y);
            for (int ox = lo0; ox <= hi0; ++ox) {
                int x = cx + ox; x = x < 0 ? x + g.n[0] : (x >= g.n[0] ? x - g.n[0] : x);
                int cc = (z * g.n[1] + y) * g.n[0] + x;
                const int b = cstart[cc], e = cstart[cc + 1];
                for (int b0 = b; b0 < e; b0 += 32) {
                    const int m = e - b0 < 32 ? e - b0 : 32;
                    uint32_t mask = 0;
                    for (int t = 0; t < m; ++t) {  // FP32 prefilter: no branch on the outcome, one bit per candidate
                        const int j = IDENT ? b0 + t : catoms[b0 + t];
                        const float4 q = posf[j];
                        bool keep = j != i && (__float_as_uint(q.w) & bit2) != 0u;
                        if (pf.on) {
                            float fx = q.x - pif.x, fy = q.y - pif.y, fz = q.z - pif.z;
                            fx = fx >= pf.h[0] ? fx - pf.L[0] : (fx < -pf.h[0] ? fx + pf.L[0] : fx);
                            fy = fy >= pf.h[1] ? fy - pf.L[1] : (fy < -pf.h[1] ? fy + pf.L[1] : fy);
                            fz = fz >= pf.h[2] ? fz - pf.L[2] : (fz < -pf.h[2] ? fz + pf.L[2] : fz);
                            keep = keep && (fmaf(fz, fz, fmaf(fy, fy, fx * fx)) < pf.lim);
                        }
                        mask |= keep ? (1u << t) : 0u;
                    }
                    while (mask) {                  // exact test of the survivors, in candidate order
                        const int t = PFMDS_FFS(mask) - 1;
                        mask &= mask - 1u;
                        const int j = IDENT ? b0 + t : catoms[b0 + t];
                        double4 pj = pos[j];
                        double dx = min_image(pj.x - pi.x, box.h[0], box.L[0]);
                        double dy = min_image(pj.y - pi.y, box.h[1], box.L[1]);
                        double dz = min_image(pj.z - pi.z, box.h[2], box.L[2]);
                        double dr2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                        if (dr2 < rc2) {
                            if (cnt < maxn) {
                                if (!PART) nlist[(size_t)cnt * stride + i] = j;
                                else if (dr2 < r1sq) nlist[(size_t)(c0++) * stride + i] = j;
                                else if (dr2 < r2sq) alt[(size_t)(c1++) * stride + i] = j;
                                else alt[(size_t)(maxn - 1 - (c2++)) * stride + i] = j;
                            }
                            ++cnt;
                        }
                    }
                }
            }
        }
    }
    if (cnt > maxn) { raise_error(err, E_TOO_MANY, orig[i], cnt); cnt = maxn; }  // md_neighbours.f90:80
    if (PART) {
        for (int k = 0; k < c1; ++k) nlist[(size_t)(c0 + k) * stride + i] = alt[(size_t)k * stride + i];
        for (int k = 0; k < c2; ++k) nlist[(size_t)(c0 + c1 + k) * stride + i] = alt[(size_t)(maxn - 1 - k) * stride + i];
    }
    nnum[i] = cnt;
}

#ifdef PFMDS_COOP
// ------------------------------------------------------------------------------------------------
// Cell-tiled build (large systems, atoms physically in cell order, every axis at least three cells wide).
// ncu on k_build_mask (profiles/r2a_k_build_mask.txt): 84 warp instructions per candidate, issue slots 83 % busy, 19 of 32 lanes
// active -- every thread walks the 27 cells on its own, fetching each candidate's 16-byte record with its own load and folding it
// into the box with three compare/select pairs.  Here ONE WARP owns ONE CELL: its lanes are the cell's atoms (list owners), and
// the candidates of the 27 surrounding cells are staged in shared memory, one x-row of three cells (one or two contiguous ranges
// of the float4 copy) at a time, double buffered, by 1-D bulk copies (cp.async.bulk, completion on an mbarrier) -- the one
// tile-shaped movement of this code.  Every lane then reads the SAME staged record (shared-memory broadcast, one wavefront per
// warp and candidate) and the periodic image is a property of the range, not of the pair: the owner's coordinates are shifted
// once per range, so the FP32 prefilter is three subtractions, three multiply-adds and a compare per candidate.  Survivors are
// recorded in a per-lane bit mask (32 candidates at a time) and take the exact FP64 test in candidate order, as in k_build_mask.
//   Bit-identical rows: the candidates are visited in k_build's order (oz, oy, ox; slot order inside a cell), and the exact test
// computes d = (x_j - x_i) + s with s = 0 or -+L of the range, which is min_image's own arithmetic whenever both agree on the
// image -- they do for every survivor: a survivor has |d| < r_cut + margin < L/2 in each component under the range's image, and
// with three or more cells per axis each neighbouring cell is visited under exactly one image.
#define CB_WARPS 4     // cells (warps) per block; each warp works alone: no block-wide barrier after the prologue
#define CB_CAP 64      // records per stage buffer (1 KB); a three-cell row goes through in one or two pieces
struct CbRange { int start, count; float sx, sy, sz; int shifted; };
#ifdef __CUDA_ARCH__
__device__ __forceinline__ unsigned cb_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
#endif
// lane 0: arm the barrier with the byte count and start the bulk copy of `n` records; other lanes: nothing
__device__ __forceinline__ void cb_issue(float4* dst, const float4* src, int n, unsigned long long* bar, int lane) {
#ifdef __CUDA_ARCH__
    if (lane == 0) {
        const unsigned bytes = (unsigned)n * 16u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cb_smem(bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(cb_smem(dst)), "l"(src), "r"(bytes),
                     "r"(cb_smem(bar))
                     : "memory");
    }
#else  // host replay: a plain copy by the lanes of the warp
    (void)bar;
    for (int k = lane; k < n; k += 32) dst[k] = src[k];
#endif
}
__device__ __forceinline__ void cb_wait(unsigned long long* bar, unsigned parity) {
#ifdef __CUDA_ARCH__
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(cb_smem(bar)),
        "r"(parity)
        : "memory");
#else
    (void)bar; (void)parity;
    __syncwarp();
#endif
}
// Phase 1 (lanes = the cell's atoms): every staged candidate is read by all lanes at once (broadcast) and the FP32 prefilter's
// survivors (bit mask per 32 candidates) are appended, in candidate order, to a per-lane list in shared memory (16-bit entries: range
// number, offset in range).  Phase 2: every lane takes the exact FP64 test over its own list and writes its own row, as k_build_mask
// does per 32 candidates -- but once per ~60 listed survivors, when the lanes' trip counts are nearly equal (the first two versions
// of this kernel, profiles/r2e_* r2f_* r2h_*: exact test inside the candidate loop, 13 of 32 lanes active; one atom's survivors spread
// over the lanes with ballot / popc compaction, 100 instructions per trip and six shuffles per atom: both slower than k_build_mask).
#define CB_OFF_BITS 11                      // offset of a candidate inside its range (three cells): ranges longer than 2048 atoms use k_build_mask
template <bool PART, bool CHECK2>
__global__ void __launch_bounds__(32 * CB_WARPS, 8) k_build_cell(int ncells, const double4* __restrict__ pos, const float4* __restrict__ posf,
                                                              const int* __restrict__ orig, const int* __restrict__ cstart, GridD g, BoxD box, PrefD pf,
                                                              uint32_t bit1, uint32_t bit2, double rc2, double r1sq, double r2sq, int maxn, size_t stride,
                                                              int* __restrict__ nlist, int* __restrict__ alt, int* __restrict__ nnum, int* err, int lcap) {
    __shared__ __align__(128) float4 buf[CB_WARPS][2][CB_CAP];
    __shared__ __align__(8) unsigned long long bars[CB_WARPS][2];
    __shared__ CbRange rng[CB_WARPS][18];
    __shared__ double dshift[CB_WARPS][18][3];   // the ranges' periodic shifts in FP64 (0 or -+L)
#ifdef __CUDACC__
    extern __shared__ __align__(16) unsigned short surv_all[];  // [CB_WARPS][32][lcap]: prefilter survivors of each lane's atom
#else
    unsigned short* const surv_all = reinterpret_cast<unsigned short*>(emu_dyn_buf().data());  // host replay: the block's dynamic shared memory
#endif
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * CB_WARPS + w;
    if (c >= ncells) return;
    const int ob = cstart[c], oe = cstart[c + 1];
    if (ob >= oe) return;  // empty cell
    unsigned short* const surv = surv_all + (size_t)w * 32 * lcap;
    // ---- the 18 candidate ranges of this cell: 9 (oz, oy) rows x [cells x-1..x+1, split where the row wraps] ----
    const int cx = c % g.n[0], cy = (c / g.n[0]) % g.n[1], cz = c / (g.n[0] * g.n[1]);
    if (lane < 18) {
        const int r = lane >> 1, part = lane & 1;
        const int oz = r / 3 - 1, oy = r % 3 - 1;
        int z = cz + oz, y = cy + oy;
        float sz = 0.f, sy = 0.f, sx = 0.f;
        if (z < 0) { z += g.n[2]; sz = -pf.L[2]; } else if (z >= g.n[2]) { z -= g.n[2]; sz = pf.L[2]; }
        if (y < 0) { y += g.n[1]; sy = -pf.L[1]; } else if (y >= g.n[1]) { y -= g.n[1]; sy = pf.L[1]; }
        const int rowc = (z * g.n[1] + y) * g.n[0];
        int first, last;  // cells of this part, in visiting order x-1, x, x+1
        if (cx == 0) { if (part == 0) { first = last = g.n[0] - 1; sx = -pf.L[0]; } else { first = 0; last = 1; } }
        else if (cx == g.n[0] - 1) { if (part == 0) { first = cx - 1; last = cx; } else { first = last = 0; sx = pf.L[0]; } }
        else { if (part == 0) { first = cx - 1; last = cx + 1; } else { first = 0; last = -1; } }
        CbRange R;
        R.start = cstart[rowc + first];
        R.count = last >= first ? cstart[rowc + last + 1] - R.start : 0;
        R.sx = sx; R.sy = sy; R.sz = sz;
        R.shifted = (sx != 0.f) || (sy != 0.f) || (sz != 0.f);
        rng[w][lane] = R;
        dshift[w][lane][0] = sx == 0.f ? 0. : (sx < 0.f ? -box.L[0] : box.L[0]);
        dshift[w][lane][1] = sy == 0.f ? 0. : (sy < 0.f ? -box.L[1] : box.L[1]);
        dshift[w][lane][2] = sz == 0.f ? 0. : (sz < 0.f ? -box.L[2] : box.L[2]);
    }
#ifdef __CUDA_ARCH__
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cb_smem(&bars[w][0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cb_smem(&bars[w][1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
#endif
    __syncwarp();
    unsigned use0 = 0u, use1 = 0u;  // completed uses of each stage buffer: parity of the next wait
    for (int o0 = ob; o0 < oe; o0 += 32) {  // the cell's atoms, 32 at a time (one pass for a crystal)
        const int i = o0 + lane;
        bool owner = false;
        float4 pif = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < oe) {
            pif = posf[i];
            owner = (__float_as_uint(pif.w) & (bit1 | PFMDS_GHOST)) == bit1;
            if (!owner) nnum[i] = 0;
        }
        // Row state of this lane's atom (phase 2 reads and updates it through shuffles): entries so far, per class
        int cnt = 0, c0 = 0, c1 = 0, c2 = 0;
        bool ovf = false;                             // a survivor list overflowed: the atom is redone by the serial routine at the end
        int ns = 0;                                   // survivors of this lane's atom waiting in its list
        unsigned short* const mine = surv + (size_t)lane * lcap;
        // ---- phase 2: exact FP64 test of the listed survivors; called whenever a list could overflow, and at the end ----
        auto flush = [&]() {
            // every lane walks its OWN list (the lists of a cell's atoms have nearly the same length: in the crystal 60 +- 5 when the first
            // one fills up), so the lanes stay together without shuffles, ballots or a second mapping of lanes to work
            const double4 pi = owner ? pos[i] : make_double4(0., 0., 0., 0.);
            for (int sidx = 0; sidx < ns; ++sidx) {
                const unsigned e = mine[sidx];
                const int r = (int)(e >> CB_OFF_BITS);
                const int j = rng[w][r].start + (int)(e & ((1u << CB_OFF_BITS) - 1u));
                const double4 pj = pos[j];
                double dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
                if (rng[w][r].shifted) {  // min_image's own d - L / d + L (see the header of this kernel)
                    dx += dshift[w][r][0]; dy += dshift[w][r][1]; dz += dshift[w][r][2];
                }
                const double dr2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                // one store on one path for the three classes (a branch per class left 9 of 32 lanes active here, ncu r2i)
                const bool in = (dr2 < rc2) && (j != i);
                const bool k0 = !PART || dr2 < r1sq, k1 = PART && !k0 && dr2 < r2sq;
                int* const dst = k0 ? nlist : alt;
                const int row = !PART ? cnt : (k0 ? c0 : (k1 ? c1 : maxn - 1 - c2));
                if (in && cnt < maxn) dst[(size_t)row * stride + i] = j;
                if (PART && in && cnt < maxn) { c0 += k0; c1 += k1; c2 += !k0 && !k1; }
                cnt += in;
            }
            ns = 0;
        };
        // ---- phase 1: FP32 prefilter, lanes = atoms of the cell ----
        int qr = 0, qo = 0, k_issue = 0, k_done = 0;  // piece being fetched (q*) / processed (p*)
#define CB_SKIP_EMPTY(r, o) while ((r) < 18 && (o) >= rng[w][(r)].count) { ++(r); (o) = 0; }
#define CB_ISSUE_NEXT() do { if (qr < 18) { const int s_ = rng[w][qr].start, c_ = rng[w][qr].count; const int n_ = c_ - qo < CB_CAP ? c_ - qo : CB_CAP; \
            cb_issue(buf[w][k_issue & 1], posf + s_ + qo, n_, &bars[w][k_issue & 1], lane); ++k_issue; qo += n_; CB_SKIP_EMPTY(qr, qo) } } while (0)
        CB_SKIP_EMPTY(qr, qo)
        int pr = 0, po = 0;
        CB_SKIP_EMPTY(pr, po)
        CB_ISSUE_NEXT();
        CB_ISSUE_NEXT();
        while (pr < 18) {
            const CbRange R = rng[w][pr];
            const int n = R.count - po < CB_CAP ? R.count - po : CB_CAP;
            const int b = k_done & 1;
            cb_wait(&bars[w][b], (b ? use1 : use0) & 1u);
            if (b) use1 += 1u; else use0 += 1u;
            const float4* cand = buf[w][b];
            const unsigned tag = ((unsigned)pr << CB_OFF_BITS) + (unsigned)po;
            const float ox = pif.x - R.sx, oy = pif.y - R.sy, oz = pif.z - R.sz;  // candidate + s - owner = candidate - (owner - s)
            if (R.count > (1 << CB_OFF_BITS)) ovf = owner;  // offsets would not fit the 16-bit entries: these atoms take the serial path
            const bool act = owner && !ovf;
            for (int g0 = 0; g0 < n; g0 += 32) {
                const int m = n - g0 < 32 ? n - g0 : 32;
                // a group adds at most 32 survivors per lane: make room first (phase 2 empties every list), so no list ever overflows
                if (__ballot_sync(0xffffffffu, act && ns + m > lcap) != 0u) flush();
                if (act) {
                    // every candidate's tag is stored at the end of the list and the list grows only when the candidate passes: no mask,
                    // no loop over set bits (the atom itself passes too and is dropped by phase 2)
                    const float4* cg = cand + g0;
                    const unsigned short tg = (unsigned short)(tag + (unsigned)g0);
                    if (m == 32) {
#pragma unroll
                        for (int t = 0; t < 32; ++t) {
                            const float4 q = cg[t];
                            const float fx = q.x - ox, fy = q.y - oy, fz = q.z - oz;
                            bool keep = fmaf(fz, fz, fmaf(fy, fy, fx * fx)) < pf.lim;
                            if (CHECK2) keep = keep && (__float_as_uint(q.w) & bit2) != 0u;
                            mine[ns] = (unsigned short)(tg + t);
                            ns += keep;
                        }
                    } else {
                        for (int t = 0; t < m; ++t) {
                            const float4 q = cg[t];
                            const float fx = q.x - ox, fy = q.y - oy, fz = q.z - oz;
                            bool keep = fmaf(fz, fz, fmaf(fy, fy, fx * fx)) < pf.lim;
                            if (CHECK2) keep = keep && (__float_as_uint(q.w) & bit2) != 0u;
                            mine[ns] = (unsigned short)(tg + t);
                            ns += keep;
                        }
                    }
                }
            }
            __syncwarp();  // every lane is done with this stage buffer: it may be refilled
            ++k_done;
            po += n;
            CB_SKIP_EMPTY(pr, po)
            CB_ISSUE_NEXT();
        }
#undef CB_SKIP_EMPTY
#undef CB_ISSUE_NEXT
        flush();
        // ---- row epilogue (as k_build) ----
        if (owner) {
            if (ovf) build_row<true, PART>(i, pos, posf, orig, cstart, nullptr, g, box, pf, bit1, bit2, rc2, r1sq, r2sq, maxn, stride, nlist, alt, nnum, err);
            else {
                if (cnt > maxn) { raise_error(err, E_TOO_MANY, orig[i], cnt); cnt = maxn; }  // md_neighbours.f90:80
                if (PART) {
                    for (int k = 0; k < c1; ++k) nlist[(size_t)(c0 + k) * stride + i] = alt[(size_t)k * stride + i];
                    for (int k = 0; k < c2; ++k) nlist[(size_t)(c0 + c1 + k) * stride + i] = alt[(size_t)(maxn - 1 - k) * stride + i];
                }
                nnum[i] = cnt;
            }
        }
        __syncwarp();
    }
}
#endif  // PFMDS_COOP

#ifdef PFMDS_COOP  // ballot / popc compaction across the lanes of a warp: device, or the lock-step host replay
// One WARP per list-owner atom: the lanes test 32 candidates of a cell range at a time (coalesced 16-byte
// loads of the float4 copy), the survivors of the exact FP64 test are compacted with ballot + popc into the
// row in candidate order (deterministic), one counter per class when PART.  The three x-neighbour cells of a
// (y,z) pair are contiguous in the cell order, so a row is assembled from at most 9 ranges.
template <bool IDENT, bool PART>
__global__ void __launch_bounds__(256) k_build_warp(int N, const double4* __restrict__ pos, const float4* __restrict__ posf, const int* __restrict__ orig,
                                               const int* __restrict__ cstart, const int* __restrict__ catoms, GridD g, BoxD box, PrefD pf,
                                               uint32_t bit1, uint32_t bit2, double rc2, double r1sq, double r2sq, int maxn, size_t stride,
                                               int* __restrict__ nlist, int* __restrict__ alt, int* __restrict__ nnum, int* err) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // warp-uniform
    const int lane = threadIdx.x & 31;
    if (i >= N) return;
    const float4 pif = posf[i];
    if ((__float_as_uint(pif.w) & (bit1 | PFMDS_GHOST)) != bit1) { if (lane == 0) nnum[i] = 0; return; }  // owners: in group 1 and not a ghost copy
    const double4 pi = pos[i];
    // same binning expression as k_cell_count
    const int cx = cell_coord(pi.x, g.inv[0], g.n[0]), cy = cell_coord(pi.y, g.inv[1], g.n[1]), cz = cell_coord(pi.z, g.inv[2], g.n[2]);
    const int lo1 = g.n[1] >= 3 ? -1 : 0, hi1 = g.n[1] >= 2 ? 1 : 0;
    const int lo2 = g.n[2] >= 3 ? -1 : 0, hi2 = g.n[2] >= 2 ? 1 : 0;
    // x ranges: [x0,x1] without wrap plus an optional wrapped single cell on either side
    int xa = cx - (g.n[0] >= 3 ? 1 : 0), xb = cx + (g.n[0] >= 2 ? 1 : 0);
    int wrap_lo = -1, wrap_hi = -1;
    if (xa < 0) { wrap_lo = g.n[0] - 1; xa = 0; }
    if (xb >= g.n[0]) { wrap_hi = 0; xb = g.n[0] - 1; }
    if (g.n[0] == 2) { xa = 0; xb = 1; wrap_lo = wrap_hi = -1; }
    const unsigned lt = (1u << lane) - 1u;
    int cnt = 0, c0 = 0, c1 = 0, c2 = 0;
    for (int oz = lo2; oz <= hi2; ++oz) {
        int z = cz + oz; z = z < 0 ? z + g.n[2] : (z >= g.n[2] ? z - g.n[2] : z);
        for (int oy = lo1; oy <= hi1; ++oy) {
            int y = cy + oy; y = y < 0 ? y + g.n[1] : (y >= g.n[1] ? y - g.n[1] : y);
            const int rowc = (z * g.n[1] + y) * g.n[0];
            for (int seg = 0; seg < 3; ++seg) {
                int b, e;
                if (seg == 0) { b = cstart[rowc + xa]; e = cstart[rowc + xb + 1]; }
                else if (seg == 1) { if (wrap_lo < 0) continue; b = cstart[rowc + wrap_lo]; e = cstart[rowc + wrap_lo + 1]; }
                else { if (wrap_hi < 0) continue; b = cstart[rowc + wrap_hi]; e = cstart[rowc + wrap_hi + 1]; }
                for (int s0 = b; s0 < e; s0 += 32) {
                    const int s = s0 + lane;
                    bool ok = s < e;
                    int j = 0;
                    double dr2 = 0.;
                    if (ok) {
                        j = IDENT ? s : catoms[s];
                        const float4 q = posf[j];
                        ok = (j != i) && (__float_as_uint(q.w) & bit2);
                        if (ok && pf.on) {
                            float fx = q.x - pif.x, fy = q.y - pif.y, fz = q.z - pif.z;
                            fx = fx >= pf.h[0] ? fx - pf.L[0] : (fx < -pf.h[0] ? fx + pf.L[0] : fx);
                            fy = fy >= pf.h[1] ? fy - pf.L[1] : (fy < -pf.h[1] ? fy + pf.L[1] : fy);
                            fz = fz >= pf.h[2] ? fz - pf.L[2] : (fz < -pf.h[2] ? fz + pf.L[2] : fz);
                            ok = fmaf(fz, fz, fmaf(fy, fy, fx * fx)) < pf.lim;
                        }
                        if (ok) {
                            const double4 pj = pos[j];
                            double dx = min_image(pj.x - pi.x, box.h[0], box.L[0]);
                            double dy = min_image(pj.y - pi.y, box.h[1], box.L[1]);
                            double dz = min_image(pj.z - pi.z, box.h[2], box.L[2]);
                            dr2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                            ok = dr2 < rc2;
                        }
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, ok);
                    if (m == 0) continue;
                    if (!PART) {
                        int slot = cnt + __popc(m & lt);
                        if (ok && slot < maxn) nlist[(size_t)slot * stride + i] = j;
                    } else {
                        const int cls = dr2 < r1sq ? 0 : (dr2 < r2sq ? 1 : 2);
                        const unsigned m0 = __ballot_sync(0xffffffffu, ok && cls == 0), m1 = __ballot_sync(0xffffffffu, ok && cls == 1), m2 = m & ~(m0 | m1);
                        if (ok && cnt + __popc(m & lt) < maxn) {  // entries beyond the capacity are only counted
                            if (cls == 0) nlist[(size_t)(c0 + __popc(m0 & lt)) * stride + i] = j;
                            else if (cls == 1) alt[(size_t)(c1 + __popc(m1 & lt)) * stride + i] = j;
                            else alt[(size_t)(maxn - 1 - (c2 + __popc(m2 & lt))) * stride + i] = j;
                        }
                        c0 += __popc(m0); c1 += __popc(m1); c2 += __popc(m2);
                    }
                    cnt += __popc(m);
                }
            }
        }
    }
    if (cnt > maxn) {  // md_neighbours.f90:80
        if (lane == 0) { raise_error(err, E_TOO_MANY, orig[i], cnt); nnum[i] = maxn; }
        return;        // the run stops at the next synchronisation; the row content is irrelevant
    }
    if (PART) {
        __syncwarp();
        for (int k = lane; k < c1; k += 32) nlist[(size_t)(c0 + k) * stride + i] = alt[(size_t)k * stride + i];
        for (int k = lane; k < c2; k += 32) nlist[(size_t)(c0 + c1 + k) * stride + i] = alt[(size_t)(maxn - 1 - k) * stride + i];
    }
    if (lane == 0) nnum[i] = cnt;
}

#endif

// FP32 prefilter limits for a list with cut-off rcut in `box` (pure arithmetic: shared by nl_build and the host emulation tests)
inline PrefD nl_prefilter(const BoxD& box, double rcut) {
    PrefD pf;
    const double rc2 = rcut * rcut;
    double Lmax = std::max(box.L[0], std::max(box.L[1], box.L[2])), hmin = std::min(box.h[0], std::min(box.h[1], box.h[2]));
    for (int k = 0; k < 3; ++k) { pf.L[k] = (float)box.L[k]; pf.h[k] = (float)box.h[k]; }
    // |delta d| <= 3 ulp-halves of L per component, r2 error <= 2 sqrt(3) r |delta d| + float rounding of r2; doubled
    double dd = 2e-7 * Lmax, margin = 2.0 * (3.5 * rcut * dd + 4e-7 * rc2 + 3 * dd * dd);
    pf.lim = (float)(rc2 + margin) * (1.0f + 2e-7f);
    pf.on = hmin > rcut * 1.01 + 10 * dd;  // tiny boxes: the float wrap could pick another image, use the exact test only
    return pf;
}
inline GridD nl_grid(const int ncell[3], const BoxD& box) {
    GridD g;
    for (int k = 0; k < 3; ++k) { g.n[k] = ncell[k]; g.inv[k] = ncell[k] / box.L[k]; }
    return g;
}

#ifdef PFMDS_HAVE_CTX
void nl_build(pfmds_ctx* c, NList& l) {
    const int N = c->N;
#ifdef PFMDS_COOP
    const bool warp_per_atom = N < c->nl_warp_n;  // measured: at 1e6 atoms the thread-per-atom scan is 2x faster, at 1e4 atoms 5x slower
#else
    const bool warp_per_atom = false;       // host replay: the warp-per-atom kernel compacts with ballots
#endif
    const int T = warp_per_atom ? 256 : 128, nb = warp_per_atom ? (int)(((size_t)N * 32 + T - 1) / T) : (N + T - 1) / T;
    const GridD g = nl_grid(c->ncell, c->box);
    uint32_t b1 = 1u << (l.g1 - 1), b2 = 1u << (l.g2 - 1);
    double rc2 = l.rcut * l.rcut;
    const PrefD pf = nl_prefilter(c->box, l.rcut);
    KTimer kt(c, KS_NL_BUILD);
#ifdef PFMDS_COOP
    // cell-tiled build: large, dense systems in cell order, three or more cells per axis, FP32 prefilter usable
    // (sparse cells leave most lanes of a cell's warp without an atom: below 12 atoms per cell the thread-per-atom build is faster, measured on the LJ fluid)
    // (slab mode: the grid spans the whole box, this rank's atoms sit in its 1 / nranks share of the cells -- with the global count the
    //  8-GPU runs of round 2 fell back to k_build_mask: 0.092 against 0.065 ms/step at 2 GPUs)
    const double cells_here = c->slab ? (double)c->ncells / slab_nranks(c) : (double)c->ncells;
    if (c->nl_cell && !warp_per_atom && c->identity_order && pf.on && g.n[0] >= 3 && g.n[1] >= 3 && g.n[2] >= 3 && (double)N >= 12. * cells_here) {
        const int nbc = (c->ncells + CB_WARPS - 1) / CB_WARPS;
        // the partner-group test can be dropped when every atom is in group 2 (slab mode: from the global group sizes)
        const bool chk = c->slab ? !(l.g2 >= 1 && l.g2 <= (int)c->group_count.size() && c->group_count[(size_t)l.g2 - 1] == slab_n_global(c))
                                 : !(c->h_gmask.size() == (size_t)N && c->all_in_group(l.g2));
        // prefilter survivors per atom kept in shared memory between two runs of phase 2 (16-bit entries)
        int lcap = 122;   // 61 words per lane: the lanes' list tails fall into different shared-memory banks
        if (const char* lc = std::getenv("PFMDS_NL_LCAP")) { int v = std::atoi(lc); if (v >= 40 && v <= 512) lcap = (v & ~3) | 2; }
        const size_t dyn = (size_t)CB_WARPS * 32 * lcap * sizeof(unsigned short);
#define CELL_ARGS c->ncells, c->pos, c->posf, c->orig, c->cell_start, g, c->box, pf, b1, b2, rc2, l.part_r1sq, l.part_r2sq, l.maxn, c->stride, l.nlist, l.nlist_alt, l.nnum, c->err, lcap
#ifdef __CUDACC__
#define CELL_LAUNCH(PT, CH) do { CK(cudaFuncSetAttribute(k_build_cell<PT, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)); \
        k_build_cell<PT, CH><<<nbc, 32 * CB_WARPS, dyn, c->st>>>(CELL_ARGS); } while (0)
#else
#define CELL_LAUNCH(PT, CH) do { emu_dynamic_smem(dyn); LAUNCH((k_build_cell<PT, CH>), nbc, 32 * CB_WARPS, c->st, CELL_ARGS); } while (0)
#endif
        if (l.partition) { if (chk) CELL_LAUNCH(true, true); else CELL_LAUNCH(true, false); }
        else { if (chk) CELL_LAUNCH(false, true); else CELL_LAUNCH(false, false); }
#undef CELL_LAUNCH
#undef CELL_ARGS
        c->launches += 1;
        l.built = true;
        CK(cudaGetLastError());
        return;
    }
#endif
#define BUILD_ARGS N, c->pos, c->posf, c->orig, c->cell_start, c->cell_atoms, g, c->box, pf, b1, b2, rc2, l.part_r1sq, l.part_r2sq, l.maxn, c->stride, l.nlist, \
        l.nlist_alt, l.nnum, c->err
#ifdef PFMDS_COOP
#define LAUNCH_BUILD(ID, PT) do { if (warp_per_atom) LAUNCH((k_build_warp<ID, PT>), nb, T, c->st, BUILD_ARGS); \
        else if (c->nl_mask) LAUNCH((k_build_mask<ID, PT>), nb, T, c->st, BUILD_ARGS); \
        else LAUNCH((k_build<ID, PT>), nb, T, c->st, BUILD_ARGS); } while (0)
#else
#define LAUNCH_BUILD(ID, PT) do { if (c->nl_mask) LAUNCH((k_build_mask<ID, PT>), nb, T, c->st, BUILD_ARGS); \
        else LAUNCH((k_build<ID, PT>), nb, T, c->st, BUILD_ARGS); } while (0)
#endif
    if (c->identity_order) { if (l.partition) LAUNCH_BUILD(true, true); else LAUNCH_BUILD(true, false); }
    else { if (l.partition) LAUNCH_BUILD(false, true); else LAUNCH_BUILD(false, false); }
#undef LAUNCH_BUILD
#undef BUILD_ARGS
    c->launches += 1;
    l.built = true;
    CK(cudaGetLastError());
}
#endif  // PFMDS_HAVE_CTX

// graphenenorm.f90:8-36: the entries of the carbon (tb) list closer than r_cut_nn; exactly three.
__global__ void k_nearest3(int N, const double4* __restrict__ pos, const int* __restrict__ orig, ListView src, BoxD box, double rc_nn,
                           size_t stride, int* __restrict__ nn, int* __restrict__ nnnum, int* err) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int n = src.nnum[i];
    if (n == 0) { nnnum[i] = 0; return; }
    const double4 pi = pos[i];
    int k = 0;
    for (int p = 0; p < n; ++p) {
        int j = src.nlist[(size_t)p * src.stride + i];
        double4 pj = pos[j];
        double dx = min_image(pj.x - pi.x, box.h[0], box.L[0]);
        double dy = min_image(pj.y - pi.y, box.h[1], box.L[1]);
        double dz = min_image(pj.z - pi.z, box.h[2], box.L[2]);
        double dr2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (sqrt(dr2) < rc_nn) {
            if (k < 3) nn[(size_t)k * stride + i] = j;
            ++k;
        }
    }
    if (k != 3) raise_error(err, E_GR_NEIB, orig[i], k);
    nnnum[i] = k < 3 ? k : 3;
}

#ifdef PFMDS_HAVE_CTX
void nl_nearest3_from(pfmds_ctx* c, NList& nn, const NList& src) {
    const int N = c->N, T = 128, nb = (N + T - 1) / T;
    KTimer kt(c, KS_NL_BUILD);
    LAUNCH((k_nearest3), nb, T, c->st, N, c->pos, c->orig, src.view(c->stride), c->box, nn.rcut, c->stride, nn.nlist, nn.nnum, c->err);
    c->launches += 1;
    nn.built = true;
    CK(cudaGetLastError());
}
#endif  // PFMDS_HAVE_CTX
