/*
 * oracle/nbody_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU restatement of the reference's per-timestep hot path, used
 * only as the parity checker (tests/, __graft_entry__.smoke()) and as the
 * timed CPU baseline (bench.py cpu_baseline / --impl reference).  The product
 * path (orbital-physics_b200/) never links, loads or calls anything here.
 *
 * Reference (Python, /root/reference):
 *   core/physics.py:125-159   pairwise_accelerations   -> orc_pairwise_half / orc_pairwise_rows
 *   core/engine.py:65-97      SimulationEngine.step    -> orc_step
 *   core/physics.py:510-535   handle_collisions        -> orc_collisions
 *   core/physics.py:391-422   collide_spheres          -> collide_pair
 *   core/engine.py:104-121    total_energy / angular_momentum -> orc_energy / orc_angmom
 *
 * Parity status: PINNED.  Every function is checked bit-for-bit against
 * outputs of the unmodified reference (tests/golden/ .npz files, produced by
 * tests/golden/make_golden.py) in tests/test_oracle.py.
 *
 * Rounding contract (re-verified by golden ddot3.npz):
 *   NumPy's 3-element `rij @ rij` is OpenBLAS ddot whose scalar tail contracts
 *   to fma(dz,dz, fma(dy,dy, dx*dx)); everything else is one IEEE operation
 *   per Python/NumPy operator.  Build with -ffp-contract=off so the compiler
 *   adds no contractions of its own; fma() below is explicit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* physics.py:145-146  rij = rj - ri ; float(rij @ rij) */
static inline double dot3_numpy(double dx, double dy, double dz) {
    return fma(dz, dz, fma(dy, dy, dx * dx));
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/*
 * Literal loop structure of physics.py:136-158: half matrix i<j, both bodies
 * updated per pair, U accumulated in lexicographic pair order.
 */
void orc_pairwise_half(int64_t n, const double* x, const double* y, const double* z, const double* m,
                       double eps, double G, double* ax, double* ay, double* az, double* U_out) {
    double U = 0.0;
    const double eps2 = eps * eps;                       /* :134 */
    for (int64_t i = 0; i < n; ++i) ax[i] = ay[i] = az[i] = 0.0;   /* :132 */
    for (int64_t i = 0; i < n; ++i) {
        const double mi = m[i];
        for (int64_t j = i + 1; j < n; ++j) {
            const double mj = m[j];
            const double dx = x[j] - x[i], dy = y[j] - y[i], dz = z[j] - z[i];  /* :145 */
            const double r2 = dot3_numpy(dx, dy, dz) + eps2;                    /* :146 */
            const double inv_r = 1.0 / sqrt(r2);                                /* :147 */
            const double inv_r3 = inv_r / r2;                                   /* :148 */
            const double si = (G * mj) * inv_r3;                                /* :151 */
            const double sj = ((-G) * mi) * inv_r3;                             /* :152 */
            ax[i] += si * dx; ay[i] += si * dy; az[i] += si * dz;               /* :154 */
            ax[j] += sj * dx; ay[j] += sj * dy; az[j] += sj * dz;               /* :155 */
            U += (((-G) * mi) * mj) * inv_r;                                    /* :158 */
        }
    }
    if (U_out) *U_out = U;
}

/* One target row, ascending j != i.  Bit-identical to the half-matrix form
 * because IEEE negation is exact. */
static inline void row_accel(int64_t i, int64_t n, const double* x, const double* y, const double* z,
                             const double* m, double eps2, double G, double* a3) {
    double axi = 0.0, ayi = 0.0, azi = 0.0;
    const double xi = x[i], yi = y[i], zi = z[i];
    for (int64_t j = 0; j < n; ++j) {
        if (j == i) continue;
        double dx = x[j] - xi, dy = y[j] - yi, dz = z[j] - zi;
        const double r2 = dot3_numpy(dx, dy, dz) + eps2;
        const double inv_r = 1.0 / sqrt(r2);
        const double inv_r3 = inv_r / r2;
        const double s = (G * m[j]) * inv_r3;
        axi += s * dx; ayi += s * dy; azi += s * dz;
    }
    a3[0] = axi; a3[1] = ayi; a3[2] = azi;
}

/* Target-centric all rows, OpenMP over targets (nthreads <= 0: all cores). */
void orc_pairwise_rows(int64_t n, const double* x, const double* y, const double* z, const double* m,
                       double eps, double G, double* ax, double* ay, double* az, int nthreads) {
    const double eps2 = eps * eps;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads)
#endif
    for (int64_t i = 0; i < n; ++i) {
        double a3[3];
        row_accel(i, n, x, y, z, m, eps2, G, a3);
        ax[i] = a3[0]; ay[i] = a3[1]; az[i] = a3[2];
    }
}

/* Sampled target rows for N where the full pass is too slow. */
void orc_pairwise_sample(int64_t n, const double* x, const double* y, const double* z, const double* m,
                         double eps, double G, const int64_t* rows, int64_t nrows,
                         double* ax, double* ay, double* az, int nthreads) {
    const double eps2 = eps * eps;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
#endif
    for (int64_t k = 0; k < nrows; ++k) {
        double a3[3];
        row_accel(rows[k], n, x, y, z, m, eps2, G, a3);
        ax[k] = a3[0]; ay[k] = a3[1]; az[k] = a3[2];
    }
}

/* Same rows in 80-bit long double with exact-ish arithmetic order: the
 * accuracy yardstick (not the reference's rounding). Also returns
 * sum_j |a_ij| per row (conditioning denominator). */
void orc_pairwise_sample_ld(int64_t n, const double* x, const double* y, const double* z, const double* m,
                            double eps, double G, const int64_t* rows, int64_t nrows,
                            double* ax, double* ay, double* az, double* sum_abs, int nthreads) {
    const long double eps2 = (long double)eps * (long double)eps;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
#endif
    for (int64_t k = 0; k < nrows; ++k) {
        const int64_t i = rows[k];
        long double sx = 0, sy = 0, sz = 0, sa = 0;
        for (int64_t j = 0; j < n; ++j) {
            if (j == i) continue;
            long double dx = (long double)x[j] - x[i], dy = (long double)y[j] - y[i], dz = (long double)z[j] - z[i];
            long double r2 = dx * dx + dy * dy + dz * dz + eps2;
            long double inv_r = 1.0L / sqrtl(r2);
            long double s = (long double)G * m[j] * inv_r * inv_r * inv_r;
            sx += s * dx; sy += s * dy; sz += s * dz;
            sa += s * sqrtl(dx * dx + dy * dy + dz * dz);
        }
        ax[k] = (double)sx; ay[k] = (double)sy; az[k] = (double)sz;
        if (sum_abs) sum_abs[k] = (double)sa;
    }
}

/* Potential only, lexicographic pair order (physics.py:158). */
double orc_potential(int64_t n, const double* x, const double* y, const double* z, const double* m,
                     double eps, double G) {
    double U = 0.0;
    const double eps2 = eps * eps;
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = i + 1; j < n; ++j) {
            const double dx = x[j] - x[i], dy = y[j] - y[i], dz = z[j] - z[i];
            const double r2 = dot3_numpy(dx, dy, dz) + eps2;
            const double inv_r = 1.0 / sqrt(r2);
            U += (((-G) * m[i]) * m[j]) * inv_r;
        }
    return U;
}

/* engine.py:69-70 / :81-82  `obj.velocity += 0.5 * dt * acc`
 * f64 velocity: v = v + (h*a).  f32 velocity array: the in-place += runs the
 * f64 loop and casts on store: v32 = (float)((double)v32 + h*a)  (SURVEY A.2).
 * Velocities are carried in doubles; f32 bodies hold exactly-representable values. */
static inline double kick1(double v, double h, double a, int is_f32) {
    double r = v + h * a;
    return is_f32 ? (double)(float)r : r;
}

/* engine.py:74  `obj.position() + obj.velocity * dt`
 * f32 velocity: NEP-50 weak scalar -> product in float32 with float32(dt). */
static inline double drift1(double r, double v, double dt, int is_f32) {
    if (is_f32) {
        float p = (float)v * (float)dt;
        return r + (double)p;
    }
    return r + v * dt;
}

/* physics.py:391-422 */
static void collide_pair(int64_t a, int64_t b, double* x, double* y, double* z,
                         double* vx, double* vy, double* vz, const double* m, const double* radius,
                         const uint8_t* vf32, double restitution) {
    double nx = x[a] - x[b], ny = y[a] - y[b], nz = z[a] - z[b];     /* :394 */
    const double dist = sqrt(dot3_numpy(nx, ny, nz));                /* :395 np.linalg.norm */
    if (dist == 0.0) return;                                         /* :396 */
    nx /= dist; ny /= dist; nz /= dist;                              /* :398 */
    const double m1 = m[a], m2 = m[b];
    /* :401 np.dot(obj1.velocity - obj2.velocity, n): the difference keeps the
     * common dtype (f32 - f32 -> f32), the dot with f64 n upcasts to ddot. */
    double wx, wy, wz;
    if (vf32[a] && vf32[b]) {
        wx = (double)((float)vx[a] - (float)vx[b]);
        wy = (double)((float)vy[a] - (float)vy[b]);
        wz = (double)((float)vz[a] - (float)vz[b]);
    } else {
        wx = vx[a] - vx[b]; wy = vy[a] - vy[b]; wz = vz[a] - vz[b];
    }
    const double v_rel = fma(wz, nz, fma(wy, ny, wx * nx));
    if (v_rel >= 0.0) return;                                        /* :402 */
    const double m1_inv = 1.0 / m1, m2_inv = 1.0 / m2;               /* :408-409 */
    double e = restitution; if (e < 0.0) e = 0.0; if (e > 1.0) e = 1.0;   /* :410 */
    const double j = (-(1.0 + e)) * v_rel / (m1_inv + m2_inv);       /* :412 */
    const double ix = j * nx, iy = j * ny, iz = j * nz;              /* :413 */
    /* :414-415 in-place += / -= on the velocity arrays (dtype preserved) */
    double t;
    t = vx[a] + ix / m1; vx[a] = vf32[a] ? (double)(float)t : t;
    t = vy[a] + iy / m1; vy[a] = vf32[a] ? (double)(float)t : t;
    t = vz[a] + iz / m1; vz[a] = vf32[a] ? (double)(float)t : t;
    t = vx[b] - ix / m2; vx[b] = vf32[b] ? (double)(float)t : t;
    t = vy[b] - iy / m2; vy[b] = vf32[b] ? (double)(float)t : t;
    t = vz[b] - iz / m2; vz[b] = vf32[b] ? (double)(float)t : t;
    const double overlap = radius[a] + radius[b] - dist;             /* :418 */
    if (overlap > 0.0) {
        const double corr = overlap / (m1_inv + m2_inv);             /* :420 */
        const double c1 = corr / m1, c2 = corr / m2;
        const double ax_ = x[a], ay_ = y[a], az_ = z[a], bx_ = x[b], by_ = y[b], bz_ = z[b];
        x[a] = ax_ + nx * c1; y[a] = ay_ + ny * c1; z[a] = az_ + nz * c1;   /* :421 */
        x[b] = bx_ - nx * c2; y[b] = by_ - ny * c2; z[b] = bz_ - nz * c2;   /* :422 */
    }
}

/* physics.py:510-535 with merge_on_capture=False (the only branch the engine
 * reaches, engine.py:85).  Sequential, in place.  Returns the number of
 * overlapping pairs seen. */
int64_t orc_collisions(int64_t n, double* x, double* y, double* z, double* vx, double* vy, double* vz,
                       const double* m, const double* radius, const uint8_t* vf32, double restitution) {
    int64_t hits = 0;
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = i + 1; j < n; ++j) {
            const double dx = x[i] - x[j], dy = y[i] - y[j], dz = z[i] - z[j];   /* :517 */
            const double r = sqrt(dot3_numpy(dx, dy, dz));
            if (r <= radius[i] + radius[j]) {                                     /* :518 */
                ++hits;
                collide_pair(i, j, x, y, z, vx, vy, vz, m, radius, vf32, restitution);
            }
        }
    return hits;
}

/* engine.py:65-97, `nsteps` times.  ax/ay/az carry self.acc in and out.
 * do_collisions=0 skips step 5 (for timing the force path alone).
 * nthreads: 1 = the literal single-threaded half-matrix loop; otherwise the
 * (bit-identical) row form on that many threads. Returns total collision hits. */
int64_t orc_step(int64_t n, double* x, double* y, double* z, double* vx, double* vy, double* vz,
                 const double* m, const double* radius, const uint8_t* vf32,
                 double* ax, double* ay, double* az, double dt, double eps, double G,
                 double restitution, int do_collisions, int64_t nsteps, int nthreads, double* U_out) {
    const double h = 0.5 * dt;                                        /* (0.5*dt)*acc */
    int64_t hits = 0;
    double U = 0.0;
    for (int64_t s = 0; s < nsteps; ++s) {
        for (int64_t i = 0; i < n; ++i) {                             /* :69-70 */
            vx[i] = kick1(vx[i], h, ax[i], vf32[i]);
            vy[i] = kick1(vy[i], h, ay[i], vf32[i]);
            vz[i] = kick1(vz[i], h, az[i], vf32[i]);
        }
        for (int64_t i = 0; i < n; ++i) {                             /* :73-75 */
            x[i] = drift1(x[i], vx[i], dt, vf32[i]);
            y[i] = drift1(y[i], vy[i], dt, vf32[i]);
            z[i] = drift1(z[i], vz[i], dt, vf32[i]);
        }
        if (nthreads == 1) {
            orc_pairwise_half(n, x, y, z, m, eps, G, ax, ay, az, &U); /* :78 */
        } else {
            orc_pairwise_rows(n, x, y, z, m, eps, G, ax, ay, az, nthreads);
        }
        for (int64_t i = 0; i < n; ++i) {                             /* :81-82 */
            vx[i] = kick1(vx[i], h, ax[i], vf32[i]);
            vy[i] = kick1(vy[i], h, ay[i], vf32[i]);
            vz[i] = kick1(vz[i], h, az[i], vf32[i]);
        }
        if (do_collisions)                                            /* :85 */
            hits += orc_collisions(n, x, y, z, vx, vy, vz, m, radius, vf32, restitution);
    }
    if (U_out) *U_out = U;
    return hits;
}

/* Batched ensemble: nsys independent systems of nb bodies, arrays [nsys][nb]. */
void orc_ensemble_step(int64_t nsys, int64_t nb, double* x, double* y, double* z,
                       double* vx, double* vy, double* vz, const double* m,
                       double dt, double eps, double G, int64_t nsteps, int f32_velocity, int nthreads) {
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
    for (int64_t s = 0; s < nsys; ++s) {
        const int64_t o = s * nb;
        double* ax = (double*)malloc(sizeof(double) * 3 * nb);
        double* ay = ax + nb; double* az = ay + nb;
        uint8_t* f = (uint8_t*)malloc(nb);
        double* rad = (double*)calloc(nb, sizeof(double));
        memset(f, f32_velocity ? 1 : 0, nb);
        orc_pairwise_half(nb, x + o, y + o, z + o, m + o, eps, G, ax, ay, az, 0);   /* engine.py:41 */
        orc_step(nb, x + o, y + o, z + o, vx + o, vy + o, vz + o, m + o, rad, f,
                 ax, ay, az, dt, eps, G, 1.0, 0, nsteps, 1, 0);
        free(ax); free(f); free(rad);
    }
}

/* engine.py:104-112: K = sum 0.5*m*(v@v).  With an f32 velocity array `v @ v`
 * is an f32 dot (sdot); its rounding is host-BLAS dependent, so the f32 form
 * here (fmaf chain) is only indicative -- tests compare energies with a
 * tolerance, never bitwise. */
double orc_kinetic(int64_t n, const double* vx, const double* vy, const double* vz, const double* m,
                   const uint8_t* vf32) {
    double K = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        double v2;
        if (vf32[i]) {
            float a = (float)vx[i], b = (float)vy[i], c = (float)vz[i];
            v2 = (double)fmaf(c, c, fmaf(b, b, a * a));
        } else {
            v2 = dot3_numpy(vx[i], vy[i], vz[i]);
        }
        K += (0.5 * m[i]) * v2;
    }
    return K;
}

/* engine.py:114-121: L = sum r x (m v), fp64 velocities form. */
void orc_angmom(int64_t n, const double* x, const double* y, const double* z,
                const double* vx, const double* vy, const double* vz, const double* m, double* L) {
    double lx = 0, ly = 0, lz = 0;
    for (int64_t i = 0; i < n; ++i) {
        const double px = m[i] * vx[i], py = m[i] * vy[i], pz = m[i] * vz[i];
        lx += y[i] * pz - z[i] * py;
        ly += z[i] * px - x[i] * pz;
        lz += x[i] * py - y[i] * px;
    }
    L[0] = lx; L[1] = ly; L[2] = lz;
}
