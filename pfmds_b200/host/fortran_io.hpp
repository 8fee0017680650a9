// Host-side helpers that reproduce the Fortran list-directed READ grammar and the fixed
// edit descriptors the reference uses for its settings file, parameter files, .xyz files and logs
// (reference: code_source/MOLECULAR_DYNAMICS/md_simulation.f90:48-93,
//  md_read_write.f90:10-107, INTERACTION_POTENTIALS/*.f90 read_* routines).
// Plain C++17, no CUDA, no dependencies: shared by the run_md_simulation host and by the CPU
// oracle's driver (the oracle's arithmetic never comes through here, only text I/O does).
#pragma once
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace fio {

// One Fortran unit opened for list-directed reads. Every read(u,*) consumes whole records:
// it keeps pulling lines until it has the number of items it asked for and then drops the
// rest of the last line it touched.
class ListReader {
public:
    explicit ListReader(const std::string& path) : in_(path), path_(path) {
        if (!in_) throw std::runtime_error("cannot open file: " + path);
    }
    // read(u,*) with n items; returns the raw tokens.
    std::vector<std::string> record(size_t n) {
        std::vector<std::string> out;
        bool slash = false;
        while (out.size() < n && !slash) {
            std::string line;
            if (!std::getline(in_, line))
                throw std::runtime_error("end of file while reading " + path_);
            tokenize(line, out, n, slash);
        }
        while (out.size() < n) out.emplace_back();  // '/' leaves the remaining items untouched
        return out;
    }
    // read(u,*) with an empty item list: skip one record.
    void skip() {
        std::string line;
        if (!std::getline(in_, line)) throw std::runtime_error("end of file while reading " + path_);
    }
    // read(u,'(A)') str
    std::string line() {
        std::string l;
        if (!std::getline(in_, l)) throw std::runtime_error("end of file while reading " + path_);
        if (!l.empty() && l.back() == '\r') l.pop_back();
        return l;
    }

private:
    static void tokenize(const std::string& line, std::vector<std::string>& out, size_t n, bool& slash) {
        size_t i = 0, L = line.size();
        while (i < L && out.size() < n) {
            char c = line[i];
            if (c == ' ' || c == '\t' || c == '\r' || c == ',') { ++i; continue; }
            if (c == '/') { slash = true; return; }
            std::string tok;
            if (c == '"' || c == '\'') {  // delimited character constant, doubled delimiter = literal
                char q = c; ++i;
                while (i < L) {
                    if (line[i] == q) { if (i + 1 < L && line[i + 1] == q) { tok += q; i += 2; continue; } ++i; break; }
                    tok += line[i++];
                }
            } else {
                while (i < L && line[i] != ' ' && line[i] != '\t' && line[i] != '\r' && line[i] != ',' && line[i] != '/')
                    tok += line[i++];
            }
            out.push_back(tok);
        }
    }
    std::ifstream in_;
    std::string path_;
};

inline long to_int(const std::string& t) {
    if (t.empty()) throw std::runtime_error("list-directed read: missing integer item");
    char* e = nullptr;
    long v = std::strtol(t.c_str(), &e, 10);
    if (*e != 0) throw std::runtime_error("list-directed read: bad integer '" + t + "'");
    return v;
}
inline double to_real(const std::string& t) {
    if (t.empty()) throw std::runtime_error("list-directed read: missing real item");
    std::string s = t;
    for (auto& c : s) if (c == 'd' || c == 'D') c = 'e';
    char* e = nullptr;
    double v = std::strtod(s.c_str(), &e);
    if (*e != 0) throw std::runtime_error("list-directed read: bad real '" + t + "'");
    return v;
}
inline bool to_logical(const std::string& t) {
    size_t k = 0;
    if (k < t.size() && t[k] == '.') ++k;
    if (k < t.size()) {
        char c = (char)std::toupper((unsigned char)t[k]);
        if (c == 'T') return true;
        if (c == 'F') return false;
    }
    throw std::runtime_error("list-directed read: bad logical '" + t + "'");
}

// ---- formatted output -----------------------------------------------------------------------
inline std::string fmt(const char* f, double v) { char b[128]; std::snprintf(b, sizeof b, f, v); return b; }
// fW.D : right-justified fixed; a field that does not fit becomes W asterisks.
inline std::string F(double v, int w, int d) {
    char b[512];
    int n = std::snprintf(b, sizeof b, "%*.*f", w, d, v);
    if (n > w) {  // Fortran may drop the optional leading zero before giving up
        std::string s(b);
        size_t p = s.find("0.");
        if (p != std::string::npos && (p == 0 || s[p - 1] == '-' ) && (int)s.size() - 1 == w) { s.erase(p, 1); return s; }
        return std::string((size_t)w, '*');
    }
    return b;
}
// esW.D : d.dddE+xx
inline std::string ES(double v, int w, int d) {
    char b[128];
    std::snprintf(b, sizeof b, "%*.*E", w, d, v);
    std::string s(b);
    // C prints at least two exponent digits, like Fortran; three-digit exponents drop the 'E' in Fortran.
    size_t e = s.find('E');
    if (e != std::string::npos && s.size() - e - 2 == 3) { s.erase(e, 1); if ((int)s.size() < w) s.insert(0, (size_t)w - s.size(), ' '); }
    if ((int)s.size() > w) return std::string((size_t)w, '*');
    return s;
}
// iW and iW.M
inline std::string I(long v, int w, int m = 0) {
    char b[64];
    if (m > 0) std::snprintf(b, sizeof b, "%0*ld", m, v); else std::snprintf(b, sizeof b, "%ld", v);
    std::string s(b);
    if (w == 0) return s;
    if ((int)s.size() > w) return std::string((size_t)w, '*');
    return std::string((size_t)w - s.size(), ' ') + s;
}
// AW : right-justified when the string is shorter than W, leftmost W characters otherwise.
// The argument is a Fortran character variable of length `len` (blank padded).
inline std::string A(const std::string& v, int w, int len) {
    std::string s = v;
    if ((int)s.size() < len) s.append((size_t)len - s.size(), ' '); else s.resize((size_t)len);
    if (w >= len) return std::string((size_t)(w - len), ' ') + s;
    return s.substr(0, (size_t)w);
}
// plain A with a character(len) variable
inline std::string Apad(const std::string& v, int len) {
    std::string s = v;
    if ((int)s.size() < len) s.append((size_t)len - s.size(), ' '); else s.resize((size_t)len);
    return s;
}
inline std::string L(bool v, int w) { return std::string((size_t)w - 1, ' ') + (v ? "T" : "F"); }
inline std::string trim(const std::string& s) {
    size_t e = s.find_last_not_of(' ');
    return e == std::string::npos ? std::string() : s.substr(0, e + 1);
}
// list-directed output of a default integer: one blank + right-justified in 11 (gfortran).
inline std::string LI(long v) { return I(v, 12); }
// list-directed output of a real(8) (gfortran: G-style, 17 significant digits, E+ddd exponent).
inline std::string LR(double v) {
    char b[64];
    double a = std::fabs(v);
    if (v == 0.0 || (a >= 0.1 && a < 1e16)) {
        int lead = (a < 1.0) ? 0 : (int)std::floor(std::log10(a)) + 1;
        int dec = 17 - (lead > 0 ? lead : 1) + (lead == 0 ? 1 : 0);
        if (dec < 0) dec = 0;
        std::snprintf(b, sizeof b, "%.*f", dec, v);
        std::string s(b);
        if ((int)s.size() < 21) s.insert(0, 21 - s.size(), ' ');
        return "  " + s + "     ";
    }
    std::snprintf(b, sizeof b, "%.16E", v);
    std::string s(b);
    size_t e = s.find('E');
    std::string mant = s.substr(0, e), ex = s.substr(e + 2);
    char sign = s[e + 1];
    while (ex.size() < 3) ex.insert(0, "0");
    std::string r = mant + "E" + sign + ex;
    if ((int)r.size() < 25) r.insert(0, 25 - r.size(), ' ');
    return " " + r;
}

}  // namespace fio
