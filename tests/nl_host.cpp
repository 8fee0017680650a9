// Host build of pfmds_b200/csrc/nl.cu for tests/test_kernels_host.py: binning, cell sort, re-sort and the thread-per-atom list
// build run as plain functions (tests/emu/host_emu.hpp) in the order of nl_bin_atoms / nl_build; the two device scans are
// replaced by a plain prefix sum (they exchange data between lanes).  Not a CPU path of the product.
#include <cstddef>
#include <vector>

#include "../pfmds_b200/csrc/nl.cu"

extern "C" {

// in:  N atoms in slots (pos4, gmask, orig), box L, largest r_cut of the run (cell size), one list (g1, g2, maxn, rcut, class radii)
// out: slot_orig[N] file index per slot after the optional re-sort, nlist[maxn*stride] / nnum[stride] with slot numbers,
//      err[4] device error word, ncell[3]
int nh_build(int N, const double* pos4_in, const unsigned* gmask_in, const int* orig_in, const double* L, double rc_max, int g1, int g2, int maxn,
             double rcut, int partition, double r1, double r2, int reorder, int* slot_orig, int* nlist, int* nnum, int* err, int* ncell_out) {
    const size_t stride = ((size_t)N + 31) / 32 * 32;
    BoxD box;
    for (int k = 0; k < 3; ++k) { box.L[k] = L[k]; box.h[k] = 0.5 * L[k]; }
    int ncell[3];
    double cell_rc;
    const int ncells = (int)nl_grid_dims(box, rc_max, ncell, cell_rc);
    for (int k = 0; k < 3; ++k) ncell_out[k] = ncell[k];
    const GridD g = nl_grid(ncell, box);
    std::vector<double4> pos((const double4*)pos4_in, (const double4*)pos4_in + N), vel((size_t)N, double4{0, 0, 0, 1}), pos2((size_t)N), vel2((size_t)N);
    std::vector<uint32_t> gm(gmask_in, gmask_in + N), gm2((size_t)N);
    std::vector<int> orig(orig_in, orig_in + N), orig2((size_t)N), cid((size_t)N), cnt((size_t)ncells + 1, 0), start((size_t)ncells + 1, 0), atoms((size_t)N);
    std::vector<float4> posf((size_t)N);
    const int T = 256, nb = (N + T - 1) / T;
    // nl_bin_atoms
    emu_launch(k_cell_count, nb, 1, T, N, (const double4*)pos.data(), g, cid.data(), cnt.data());
    for (int c = 0; c < ncells; ++c) start[(size_t)c + 1] = start[(size_t)c] + cnt[(size_t)c];  // k_scan_block / k_scan_sums / k_scan_add
    std::fill(cnt.begin(), cnt.end(), 0);
    emu_launch(k_cell_scatter, nb, 1, T, N, (const int*)cid.data(), (const int*)start.data(), cnt.data(), atoms.data());
    emu_launch(k_cell_sort, (ncells + T - 1) / T, 1, T, ncells, (const int*)start.data(), atoms.data(), (const int*)orig.data());
    bool ident = false;
    if (reorder) {
        emu_launch(k_permute, nb, 1, T, N, (const int*)atoms.data(), (const double4*)pos.data(), (const double4*)vel.data(), (const uint32_t*)gm.data(),
                   (const int*)orig.data(), pos2.data(), vel2.data(), gm2.data(), orig2.data(), (int*)nullptr);
        pos.swap(pos2); vel.swap(vel2); gm.swap(gm2); orig.swap(orig2);
        emu_launch(k_iota, nb, 1, T, N, atoms.data());
        ident = true;
    }
    emu_launch(k_make_posf, nb, 1, T, N, (const double4*)pos.data(), (const uint32_t*)gm.data(), posf.data());
    for (int s = 0; s < N; ++s) slot_orig[s] = orig[(size_t)s];
    // nl_build, thread-per-atom variant
    const uint32_t b1 = 1u << (g1 - 1), b2 = 1u << (g2 - 1);
    const double rc2 = rcut * rcut;
    const PrefD pf = nl_prefilter(box, rcut);
    std::vector<int> alt((size_t)maxn * stride, 0);
    const int TB = 128, nbb = (N + TB - 1) / TB;
#define EMU_BUILD(ID, PT) emu_launch(k_build<ID, PT>, nbb, 1, TB, N, (const double4*)pos.data(), (const float4*)posf.data(), (const int*)orig.data(), \
        (const int*)start.data(), (const int*)atoms.data(), g, box, pf, b1, b2, rc2, r1 * r1, r2 * r2, maxn, stride, nlist, alt.data(), nnum, err)
    if (ident) { if (partition) EMU_BUILD(true, true); else EMU_BUILD(true, false); }
    else { if (partition) EMU_BUILD(false, true); else EMU_BUILD(false, false); }
#undef EMU_BUILD
    return pf.on;
}
}
