// host build of pfmds_b200/host/fortran_io.hpp + md_inputs.hpp for the format / grammar tests (tests/test_fortran_io.py)
#include <cstring>
#include <string>
#include "../pfmds_b200/host/md_inputs.hpp"
static void put(const std::string& s, char* out, int cap) { std::strncpy(out, s.c_str(), (size_t)cap - 1); out[cap - 1] = 0; }
extern "C" {
void fio_F(double v, int w, int d, char* out, int cap) { put(fio::F(v, w, d), out, cap); }
void fio_ES(double v, int w, int d, char* out, int cap) { put(fio::ES(v, w, d), out, cap); }
void fio_I(long v, int w, int m, char* out, int cap) { put(fio::I(v, w, m), out, cap); }
void fio_A(const char* v, int w, int len, char* out, int cap) { put(fio::A(v, w, len), out, cap); }
void fio_LR(double v, char* out, int cap) { put(fio::LR(v), out, cap); }
// read `nrec` records of `n` items each from a file; tokens separated by \x1f, records by \n
int fio_records(const char* path, int nrec, int n, char* out, int cap) {
    try {
        fio::ListReader r(path);
        std::string s;
        for (int k = 0; k < nrec; ++k) {
            auto t = r.record((size_t)n);
            for (size_t i = 0; i < t.size(); ++i) s += (i ? "\x1f" : "") + t[i];
            s += "\n";
        }
        put(s, out, cap);
        return 0;
    } catch (const std::exception& e) { put(e.what(), out, cap); return 1; }
}
double fio_real(const char* t) { return fio::to_real(t); }
int fio_logical(const char* t) { try { return fio::to_logical(t) ? 1 : 0; } catch (...) { return -1; } }
// parse a settings file; returns a summary string
int fio_settings(const char* dir, const char* file, char* out, int cap) {
    try {
        auto s = pfmds_host::read_settings(dir, file);
        std::string r = std::to_string(s.md_step_limit) + "|" + s.logfilename + "|" + s.init_xyz_filename + "|" + (s.new_velocities ? "T" : "F") + "|" +
                        std::to_string(s.groups_num) + "|" + std::to_string(s.integrators_num) + "|" + std::to_string(s.nhc_num) + "|" +
                        std::to_string(s.interactions_num);
        for (auto& it : s.interactions) { r += "|" + it.name + ":" + std::to_string(it.params.size()) + ":" + std::to_string(it.lists.size()); }
        put(r, out, cap);
        return 0;
    } catch (const std::exception& e) { put(e.what(), out, cap); return 1; }
}
}
