// host build of pfmds_b200/csrc/mathx.cuh for accuracy tests (tests/test_mathx.py)
#include <cmath>
#include <cstdio>
#include "../pfmds_b200/csrc/mathx.cuh"
extern "C" {
double mx_exp(double x) { return mx::exp_fast(x); }
void mx_sincos(double a, double* s, double* c) { mx::sincos_0pi(a, *s, *c); }
double mx_exp_nc(double x) { return mx::exp_nc(x); }
void mx_cos_switch(double a, double* f, double* s) { mx::cos_switch(a, *f, *s); }
double mx_log(double x) { return mx::log_fast(x); }
double mx_pow(double x, double y) { return mx::pow_pos(x, y); }
double mx_rcp(double x) { return mx::rcp_fast(x); }
double mx_rsqrt(double x) { return mx::rsqrt_fast(x); }
}
extern "C" {
double mx_exp_m(double x) { return mx::exp_m(x); }
double mx_rsqrt_q(double x) { return mx::rsqrt_q(x); }
void mx_cos_switch_m(double a, double* f, double* s) { mx::cos_switch_m(fma(a, 0.5, -7.85398163397448279e-01), *f, *s); }
double mx_half_switch(double a) { return mx::half_switch(a - 1.57079632679489655800e+00); }
}
