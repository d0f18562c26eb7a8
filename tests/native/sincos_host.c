/* Host build of csrc/sincos_cr.h and csrc/sincos_libm.h for the CPU unit test (tests/test_sincos.py) -- TEST
 * INFRASTRUCTURE ONLY.  The product calls these routines only on the device (csrc/kepler.cu); this file lets
 * `-m "not gpu"` hold the very same operation sequences to mpmath (correct rounding) and to the host libm.
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see the test). */
#include "../../orbital-physics_b200/csrc/sincos_cr.h"
#include "../../orbital-physics_b200/csrc/sincos_libm.h"

int sc_host_sincos(const double* x, double* s, double* c, long n) {
    int ok = 1;
    for (long i = 0; i < n; ++i) ok &= sc_sincos(x[i], &s[i], &c[i]);
    return ok;
}

int sl_host_sincos(const double* x, double* s, double* c, long n) {
    int ok = 1;
    for (long i = 0; i < n; ++i) ok &= sl_sincos(x[i], &s[i], &c[i]);
    return ok;
}
