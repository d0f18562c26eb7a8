// kepler.cu -- batched Keplerian elements -> Cartesian state on the device (sm_100a).
//
// The step before the hot path: the reference turns orbital elements into the
// initial condition one body at a time on the host --
//   solve_kepler            core/physics.py:43-71   Newton on M = E - e sin E, start E=M (e<0.8) else pi,
//                                                   |dE| < tol (1e-12) after the update, <= 50 iterations
//   Body.get_state          core/body.py:184-249    perifocal r, v and R = Rz(Omega) Rx(i) Rz(omega)
// Here one thread handles one body and follows the reference's operation order with explicitly rounded
// multiplies/adds (no FMA contraction); what is left is the trigonometry, which the reference takes from the host
// libm through Python's math.sin / math.cos.  Three modes (orb_set_trig_mode / ORBITAL_B200_TRIG):
//   ORB_TRIG_LIBM (default)  sincos_libm.h: glibc 2.39's sin / cos restated operation by operation -> the device
//                            states are BIT-IDENTICAL to the reference run on an x86-64 glibc host with FMA (this image)
//   ORB_TRIG_CR              sincos_cr.h: correctly rounded sin / cos -- host independent; equals any libm wherever
//                            that libm rounds correctly (99.3 % of the golden states vs glibc, which is not)
//   ORB_TRIG_FAST            CUDA's sincos (<= 2 ulp): tolerance path, what round 1 shipped
// Arguments outside a routine's domain (|x| >= 1e8 / 2^20, inf, nan) fall through to CUDA's sincos.
//
// ens_elements_kernel fills an ensemble in place: body 0 of every system is the central mass at rest at
// the origin, bodies 1.. get parent-relative states with mean motion n = sqrt(G m_0 / a^3) and
// b = a sqrt(1 - e^2)  (core/body.py:159-169 mean_motion, :120-124 get_b).
#include "../../include/orbital_b200.h"
#include "ensemble.h"
#include "kernels.h"
#include "sincos_cr.h"
#include "sincos_libm.h"

namespace orb {

struct KeplerOut { double rx, ry, rz, vx, vy, vz, E; };

__device__ __forceinline__ double kmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double kadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double ksub(double a, double b) { return __dsub_rn(a, b); }

// math.sin / math.cos of the reference in the selected mode (see the file header)
__device__ __noinline__ void trig_dev(double x, double* s, double* c, int trig) {
    if (trig == ORB_TRIG_LIBM && sl_sincos(x, s, c)) return;
    if (trig == ORB_TRIG_CR && sc_sincos(x, s, c)) return;
    sincos(x, s, c);
}

// core/physics.py:43-71
__device__ double solve_kepler_dev(double M, double e, double tol, int max_iter, int trig) {
    double E = e < 0.8 ? M : 3.141592653589793;
    for (int it = 0; it < max_iter; ++it) {
        double s, c;
        trig_dev(E, &s, &c, trig);
        const double f = ksub(ksub(E, kmul(e, s)), M);       // E - e*sin(E) - M
        const double fp = ksub(1.0, kmul(e, c));             // 1.0 - e*cos(E)
        const double dE = __ddiv_rn(-f, fp);
        E = kadd(E, dE);
        if (fabs(dE) < tol) break;
    }
    return E;
}

// core/body.py:184-249 (parent-relative; a, b in metres, n in rad/s, angles in radians)
// e2 = e ** 2 as the host evaluates it (api.cu host_pow_plane)
__device__ KeplerOut kepler_state_dev(double M, double e, double e2, double a, double b, double n, double inc,
                                      double Omega, double omega, double tol, int max_iter, int trig) {
    KeplerOut o;
    const double E = solve_kepler_dev(M, e, tol, max_iter, trig);
    double sE, cE, sw, cw, si, ci, sO, cO;
    trig_dev(E, &sE, &cE, trig);
    trig_dev(omega, &sw, &cw, trig);
    trig_dev(inc, &si, &ci, trig);
    trig_dev(Omega, &sO, &cO, trig);
    const double den = ksub(1.0, kmul(e, cE));
    const double x_op = kmul(a, ksub(cE, e));
    const double y_op = kmul(b, sE);
    const double vx_op = __ddiv_rn(kmul(kmul(-a, n), sE), den);                        // -a*n*sin_E/(1-e*cos_E)
    const double root = __dsqrt_rn(ksub(1.0, e2));                                     // sqrt(1 - e**2)
    const double vy_op = __ddiv_rn(kmul(kmul(kmul(a, n), root), cE), den);
    const double R11 = ksub(kmul(cO, cw), kmul(kmul(sO, sw), ci));
    const double R12 = ksub(kmul(-cO, sw), kmul(kmul(sO, cw), ci));
    const double R13 = kmul(sO, si);
    const double R21 = kadd(kmul(sO, cw), kmul(kmul(cO, sw), ci));
    const double R22 = kadd(kmul(-sO, sw), kmul(kmul(cO, cw), ci));
    const double R23 = kmul(-cO, si);
    const double R31 = kmul(sw, si);
    const double R32 = kmul(cw, si);
    const double R33 = ci;
    // R . [p, q, 0]: the reference keeps the third product (R_k3 * 0.0) in the sum
    o.rx = kadd(kadd(kmul(R11, x_op), kmul(R12, y_op)), kmul(R13, 0.0));
    o.ry = kadd(kadd(kmul(R21, x_op), kmul(R22, y_op)), kmul(R23, 0.0));
    o.rz = kadd(kadd(kmul(R31, x_op), kmul(R32, y_op)), kmul(R33, 0.0));
    o.vx = kadd(kadd(kmul(R11, vx_op), kmul(R12, vy_op)), kmul(R13, 0.0));
    o.vy = kadd(kadd(kmul(R21, vx_op), kmul(R22, vy_op)), kmul(R23, 0.0));
    o.vz = kadd(kadd(kmul(R31, vx_op), kmul(R32, vy_op)), kmul(R33, 0.0));
    o.E = E;
    return o;
}

// el: 9 planes of `count` doubles (M e a b n inc Omega omega e**2); out: 7 planes (rx ry rz vx vy vz E)
__global__ void __launch_bounds__(256) kepler_states_kernel(const double* __restrict__ el, double* __restrict__ out,
                                                            long long count, double tol, int max_iter, int trig) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= count) return;
    const KeplerOut o = kepler_state_dev(el[i], el[count + i], el[8 * count + i], el[2 * count + i], el[3 * count + i],
                                         el[4 * count + i], el[5 * count + i], el[6 * count + i], el[7 * count + i],
                                         tol, max_iter, trig);
    out[i] = o.rx; out[count + i] = o.ry; out[2 * count + i] = o.rz;
    out[3 * count + i] = o.vx; out[4 * count + i] = o.vy; out[5 * count + i] = o.vz;
    out[6 * count + i] = o.E;
}

// el: 8 planes [nsys][nb-1] (M e a inc Omega omega e**2 a**3); masses are already resident in g.m
__global__ void __launch_bounds__(256) ens_elements_kernel(const EnsArgs g, const double* __restrict__ el, double tol,
                                                           int max_iter, int trig) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long tot = g.nsys * g.nb;
    if (t >= tot) return;
    const long long sys = t / g.nb;
    const int body = (int)(t - sys * g.nb);
    KeplerOut o = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (body > 0) {
        const long long k = sys * (g.nb - 1) + (body - 1);
        const long long pl = g.nsys * (long long)(g.nb - 1);
        const double M = el[k], e = el[pl + k], a = el[2 * pl + k];
        const double mu = kmul(g.G, g.m[sys * g.nb]);
        const double e2 = el[6 * pl + k], a3 = el[7 * pl + k];
        const double n = __dsqrt_rn(__ddiv_rn(mu, a3));                            // sqrt(mu / a**3)
        const double b = kmul(a, __dsqrt_rn(ksub(1.0, e2)));                       // a * sqrt(1 - e**2)
        o = kepler_state_dev(M, e, e2, a, b, n, el[3 * pl + k], el[4 * pl + k], el[5 * pl + k], tol, max_iter, trig);
    }
    g.x[t] = o.rx; g.y[t] = o.ry; g.z[t] = o.rz;
    g.vx[t] = g.vel_f32 ? (double)__double2float_rn(o.vx) : o.vx;     // Object() stores float32 velocities
    g.vy[t] = g.vel_f32 ? (double)__double2float_rn(o.vy) : o.vy;
    g.vz[t] = g.vel_f32 ? (double)__double2float_rn(o.vz) : o.vz;
}

cudaError_t launch_kepler_states(const double* d_el, double* d_out, long long count, double tol, int max_iter,
                                 int trig, cudaStream_t st) {
    if (count <= 0) return cudaSuccess;
    kepler_states_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(d_el, d_out, count, tol, max_iter, trig);
    return cudaGetLastError();
}

cudaError_t launch_ens_elements(const EnsArgs& a, const double* d_el, double tol, int max_iter, int trig,
                                cudaStream_t st) {
    const long long tot = a.nsys * a.nb;
    ens_elements_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(a, d_el, tol, max_iter, trig);
    return cudaGetLastError();
}

}  // namespace orb
