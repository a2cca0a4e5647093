// Host build of csrc/yh_libm.cuh: lets a CPU-only test compare the libm
// restatement with the glibc the oracle links (tests/test_libm_host.py).
#include <math.h>
#include <string.h>

#include "yh_libm.cuh"

extern "C" {
// double sinh of x = 1 / v as mp() takes it (v float): which 0 = restatement (yh::sinh_d), 1 = glibc; raw double bits out
void yh_host_sinh_batch(int which, const float* v, long n, unsigned long long* out) {
  for (long i = 0; i < n; i++) {
    double x = (double)(1.0f / v[i]);
    double r = which == 0 ? yh::sinh_d(x) : sinh(x);
    memcpy(&out[i], &r, 8);
  }
}
// fn: 0 expf 1 logf 2 sinf 3 cosf 4 atan2f(a,b) 5 acosf; which: 0 = restatement, 1 = glibc
void yh_host_libm_batch(int fn, int which, const float* a, const float* b, long n, float* out) {
  for (long i = 0; i < n; i++) {
    float r = 0;
    if (which == 0) {
      switch (fn) {
        case 0: r = yh::expf_(a[i]); break;
        case 1: r = yh::logf_(a[i]); break;
        case 2: r = yh::sinf_(a[i]); break;
        case 3: r = yh::cosf_(a[i]); break;
        case 4: r = yh::atan2f_(a[i], b[i]); break;
        case 5: r = yh::acosf_(a[i]); break;
      }
    } else {
      switch (fn) {
        case 0: r = expf(a[i]); break;
        case 1: r = logf(a[i]); break;
        case 2: r = sinf(a[i]); break;
        case 3: r = cosf(a[i]); break;
        case 4: r = atan2f(a[i], b[i]); break;
        case 5: r = acosf(a[i]); break;
      }
    }
    out[i] = r;
  }
}
}
