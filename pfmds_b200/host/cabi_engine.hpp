// Engine adapter for md_driver.hpp over the C ABI of include/pfmds_b200.h: this is all the host
// program knows about the GPU.  An ISO_C_BINDING Fortran md() would make the same calls
// (pfmds_b200/fortran/pfmds_b200_iso_c.f90).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pfmds_b200.h"
#include "md_inputs.hpp"

namespace pfmds_host {

class CabiEngine {
public:
    explicit CabiEngine(int device = 0) : device_(device) {}
    CabiEngine(CabiEngine&& o) noexcept : ctx_(o.ctx_), device_(o.device_), n_inter_(o.n_inter_), n_nhc_(o.n_nhc_), n_lists_(o.n_lists_) { o.ctx_ = nullptr; }
    CabiEngine(const CabiEngine&) = delete;
    ~CabiEngine() { if (ctx_) pfmds_destroy(ctx_); }

    void create(int n, const double* pos, const double* vel, const double* mass, const double* box) { ck(pfmds_create(&ctx_, device_, n, pos, vel, mass, box)); }
    void set_group(int g, const std::vector<int>& idx1) { ck(pfmds_set_group(ctx_, g, (int)idx1.size(), idx1.data())); }
    void set_roles(int am, int xyz, int z, int all) { ck(pfmds_set_roles(ctx_, am, xyz, z, all)); }
    void add_nhc(const NhcSpec& n) { ck(pfmds_add_nhc(ctx_, n.group, n.temperature, n.M, n.q1)); ++n_nhc_; }
    void add_group_change(int from, int to, int ts1, int ts2, int frec) { ck(pfmds_add_group_change(ctx_, from, to, ts1, ts2, frec)); }
    int group_size(int g) { int n = 0; ck(pfmds_group_size(ctx_, g, &n)); return n; }
    void set_misc(int zmp, bool inv) { ck(pfmds_set_misc(ctx_, zmp, inv ? 1 : 0)); }
    void add_interaction(const InteractionSpec& s) {
        std::vector<int> gn, mx, pe;
        std::vector<double> rc;
        for (auto& l : s.lists) { gn.push_back(l.g1); gn.push_back(l.g2); mx.push_back(l.neighb_num_max); rc.push_back(l.r_cut); pe.push_back(l.update_period); }
        ck(pfmds_add_interaction(ctx_, s.name.c_str(), (int)s.params.size(), s.params.data(), s.nl_n, gn.data(), mx.data(), rc.data(), pe.data()));
        ++n_inter_;
        n_lists_ += s.nl_n;
    }
    void advance(int kind, double dt, int first, int n, bool energy_after_last) {
        ck(energy_after_last ? pfmds_advance_with_energy(ctx_, kind, dt, first, n) : pfmds_advance(ctx_, kind, dt, first, n));
    }
    // n steps; the energies of every step with step % log_period == 0, kept on the device and fetched with one copy
    void advance_logged(int kind, double dt, int first, int n, int log_period, std::vector<EnergyRow>& rows) {
        const int w = n_inter_ + 2 + n_nhc_;
        std::vector<double> buf((size_t)w * (size_t)(n > 0 ? n : 1), 0.);
        int nr = 0;
        ck(pfmds_advance_logged(ctx_, kind, dt, first, n, log_period, buf.data(), w, &nr));
        rows.assign((size_t)nr, EnergyRow());
        for (int r = 0; r < nr; ++r) {
            const double* p = buf.data() + (size_t)r * w;
            rows[(size_t)r].e_inter.assign(p, p + n_inter_);
            rows[(size_t)r].ke = p[n_inter_];
            rows[(size_t)r].temp = p[n_inter_ + 1];
            rows[(size_t)r].e_nhc.assign(p + n_inter_ + 2, p + n_inter_ + 2 + n_nhc_);
        }
    }
    void energies(std::vector<double>& e_inter, double& ke, double& temp, std::vector<double>& e_nhc) {
        e_inter.assign((size_t)n_inter_ + 1, 0.);
        e_nhc.assign((size_t)n_nhc_ + 1, 0.);
        ck(pfmds_energies(ctx_, e_inter.data(), &ke, &temp, e_nhc.data()));
        e_inter.resize((size_t)n_inter_);
        e_nhc.resize((size_t)n_nhc_);
    }
    void diagnostics(double fs[3], double mc[3], double mcv[3], double& vmax, std::vector<int>& nl_load) {
        nl_load.assign((size_t)n_lists_ + 1, 0);
        ck(pfmds_diagnostics(ctx_, fs, mc, mcv, &vmax, nl_load.data()));
        nl_load.resize((size_t)n_lists_);
    }
    void download(double* pos, double* vel, double* frc) { ck(pfmds_download(ctx_, pos, vel, frc)); }
    void save_state(std::vector<double>& blob) {
        long long n = 0;
        ck(pfmds_state_size(ctx_, &n));
        blob.assign((size_t)n, 0.);
        ck(pfmds_save_state(ctx_, blob.data()));
    }
    void restore_state(const double* pos, const double* vel, const std::vector<double>& blob) {
        long long n = 0;
        ck(pfmds_state_size(ctx_, &n));
        if ((long long)blob.size() != n) throw std::runtime_error("error: the checkpoint does not belong to this settings file (state size differs)");
        ck(pfmds_restore_state(ctx_, pos, vel, blob.data()));
    }
    void timers(double t[6]) { pfmds_synchronize(ctx_); pfmds_timers(ctx_, t); }

private:
    void ck(int rc) { if (rc != PFMDS_OK) throw std::runtime_error(pfmds_last_error(ctx_)); }
    pfmds_ctx* ctx_ = nullptr;
    int device_ = 0, n_inter_ = 0, n_nhc_ = 0, n_lists_ = 0;
};

}  // namespace pfmds_host
