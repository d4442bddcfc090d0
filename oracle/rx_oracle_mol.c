/* rx_oracle_mol.c -- CPU restatement (f64) of the small-molecule path: the force field AmberPrmtopFile.createSystem builds for
 * testsystems.AlanineDipeptideVacuum (/root/reference/openmmtools/testsystems.py:3352-3388) and the constrained Langevin
 * splitting steps of /root/reference/openmmtools/integrators.py:1404-1460.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README): used by tests/, __graft_entry__.smoke() and bench.py's CPU arms, never by the
 * product.  Parity status: UNPINNED against OpenMM (not installable here) -- the functional forms below are OpenMM's documented
 * ones (HarmonicBondForce 1/2 k (r-r0)^2, HarmonicAngleForce 1/2 k (t-t0)^2, PeriodicTorsionForce k (1 + cos(n phi - phase)),
 * NonbondedForce NoCutoff: 138.935456 q_i q_j / r + 4 eps ((s/r)^12 - (s/r)^6) with Lorentz-Berthelot rules, exceptions
 * replacing the 1-4 pairs, exclusions for 1-2/1-3); forces are checked against finite differences of the energy
 * (tests/test_oracle_molecule.py).  The integrator algebra follows integrators.py line by line:
 *   R (:1404-1422)  x1 = x + (dt/nR) v;  x = constrain(x1);  v += (x - x1) / (dt/nR);  constrain velocities
 *   V (:1424-1447)  v += (dt/nV) f / m;  constrain velocities
 *   O (:1449-1460)  v = a v + b sqrt(kT/m) xi;  constrain velocities
 * Position constraints: SHAKE iterations to a relative tolerance on the squared length; velocity constraints: RATTLE
 * iterations (both Gauss-Seidel over the constraint list; OpenMM uses CCMA for positions -- another solver of the same
 * equations, so results agree to the tolerance, not bitwise).  CMMotionRemover: the centre-of-mass velocity is removed at the
 * beginning of every step (the integrator's addUpdateContextState, integrators.py:1346). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ONE_4PI_EPS0 138.935456
#define ORC_API __attribute__((visibility("default")))

typedef struct {
    int n_atoms, n_bonds, n_angles, n_torsions, n_excl, n_exc, n_cons, remove_cm;
    const double *mass, *charge, *sigma, *eps;
    const double *bonds;      /* [n_bonds][4]   i, j, K, r0 */
    const double *angles;     /* [n_angles][5]  i, j, k, K, t0 */
    const double *torsions;   /* [n_torsions][7] i, j, k, l, n, phase, k */
    const int64_t *excl;      /* [n_excl][2] */
    const double *exc;        /* [n_exc][5]  i, j, qq, sigma, eps */
    const double *cons;       /* [n_cons][3] i, j, d */
} orc_mol;

static void cross(const double *a, const double *b, double *c) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
static double dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* Potential energy (kJ/mol); f (may be NULL) receives the forces. */
ORC_API double orc_mol_energy(const orc_mol *m, const double *x, double *f) {
    const int n = m->n_atoms;
    double U = 0.0;
    if (f) memset(f, 0, sizeof(double) * 3 * (size_t)n);
    for (int b = 0; b < m->n_bonds; b++) {
        const double *p = m->bonds + 4 * b;
        const int i = (int)p[0], j = (int)p[1];
        double d[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
        const double r = sqrt(dot(d, d)), dr = r - p[3];
        U += 0.5 * p[2] * dr * dr;
        if (f) { const double c = -p[2] * dr / r; for (int q = 0; q < 3; q++) { f[3 * i + q] += c * d[q]; f[3 * j + q] -= c * d[q]; } }
    }
    for (int a = 0; a < m->n_angles; a++) {
        const double *p = m->angles + 5 * a;
        const int i = (int)p[0], j = (int)p[1], k = (int)p[2];
        double u[3], v[3];
        for (int q = 0; q < 3; q++) { u[q] = x[3 * i + q] - x[3 * j + q]; v[q] = x[3 * k + q] - x[3 * j + q]; }
        const double ru = sqrt(dot(u, u)), rv = sqrt(dot(v, v));
        double c = dot(u, v) / (ru * rv);
        if (c > 1.0) c = 1.0;
        if (c < -1.0) c = -1.0;
        const double t = acos(c), dt = t - p[4];
        U += 0.5 * p[3] * dt * dt;
        if (f) {
            const double s = sqrt(1.0 - c * c);
            const double g = (s > 1e-12) ? p[3] * dt / s : 0.0;   /* -dU/dt * dt/dcos = K dt / sin t */
            for (int q = 0; q < 3; q++) {
                const double fi = g * (v[q] / (ru * rv) - c * u[q] / (ru * ru));
                const double fk = g * (u[q] / (ru * rv) - c * v[q] / (rv * rv));
                f[3 * i + q] += fi; f[3 * k + q] += fk; f[3 * j + q] -= fi + fk;
            }
        }
    }
    for (int t = 0; t < m->n_torsions; t++) {
        const double *p = m->torsions + 7 * t;
        const int i = (int)p[0], j = (int)p[1], k = (int)p[2], l = (int)p[3];
        const double per = p[4], phase = p[5], kk = p[6];
        /* IUPAC dihedral: r_ij = x_i - x_j, r_kj = x_k - x_j, r_kl = x_k - x_l, m = r_ij x r_kj, n = r_kj x r_kl,
         * phi = atan2(|r_kj| r_ij.n, m.n); gradient as in Bekker et al. (the form GROMACS documents):
         * F_i = -U' |r_kj| / |m|^2 m,  F_l = +U' |r_kj| / |n|^2 n,  F_j = -F_i + p F_i - q F_l,  F_k = -F_l - p F_i + q F_l
         * with p = r_ij.r_kj / |r_kj|^2, q = r_kl.r_kj / |r_kj|^2. */
        double rij[3], rkj[3], rkl[3], mm[3], nn[3];
        for (int q = 0; q < 3; q++) { rij[q] = x[3 * i + q] - x[3 * j + q]; rkj[q] = x[3 * k + q] - x[3 * j + q]; rkl[q] = x[3 * k + q] - x[3 * l + q]; }
        cross(rij, rkj, mm); cross(rkj, rkl, nn);
        const double nrkj = sqrt(dot(rkj, rkj));
        const double phi = atan2(nrkj * dot(rij, nn), dot(mm, nn));
        U += kk * (1.0 + cos(per * phi - phase));
        if (f) {
            const double dU = -kk * per * sin(per * phi - phase);
            const double m2 = dot(mm, mm), n2 = dot(nn, nn);
            const double pp = dot(rij, rkj) / (nrkj * nrkj), qq = dot(rkl, rkj) / (nrkj * nrkj);
            for (int q = 0; q < 3; q++) {
                const double fi = -dU * nrkj / m2 * mm[q], fl = dU * nrkj / n2 * nn[q];
                const double sv = pp * fi - qq * fl;
                f[3 * i + q] += fi; f[3 * l + q] += fl; f[3 * j + q] += sv - fi; f[3 * k + q] += -sv - fl;
            }
        }
    }
    /* nonbonded: all pairs that are neither excluded nor exceptions */
    unsigned char *skip = (unsigned char *)calloc((size_t)n * n, 1);
    for (int e = 0; e < m->n_excl; e++) { const int i = (int)m->excl[2 * e], j = (int)m->excl[2 * e + 1]; skip[i * n + j] = skip[j * n + i] = 1; }
    for (int e = 0; e < m->n_exc; e++) { const int i = (int)m->exc[5 * e], j = (int)m->exc[5 * e + 1]; skip[i * n + j] = skip[j * n + i] = 1; }
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) {
            if (skip[i * n + j]) continue;
            double d[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
            const double r2 = dot(d, d), r = sqrt(r2);
            const double qq = ONE_4PI_EPS0 * m->charge[i] * m->charge[j];
            const double s = 0.5 * (m->sigma[i] + m->sigma[j]), e = sqrt(m->eps[i] * m->eps[j]);
            const double s6 = pow(s * s / r2, 3.0);
            U += qq / r + 4.0 * e * (s6 * s6 - s6);
            if (f) {
                const double c = (qq / r + 24.0 * e * (2.0 * s6 * s6 - s6)) / r2;
                for (int q = 0; q < 3; q++) { f[3 * i + q] += c * d[q]; f[3 * j + q] -= c * d[q]; }
            }
        }
    free(skip);
    for (int e = 0; e < m->n_exc; e++) {
        const double *p = m->exc + 5 * e;
        const int i = (int)p[0], j = (int)p[1];
        double d[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
        const double r2 = dot(d, d), r = sqrt(r2);
        const double qq = ONE_4PI_EPS0 * p[2], s = p[3], ee = p[4];
        const double s6 = pow(s * s / r2, 3.0);
        U += qq / r + 4.0 * ee * (s6 * s6 - s6);
        if (f) {
            const double c = (qq / r + 24.0 * ee * (2.0 * s6 * s6 - s6)) / r2;
            for (int q = 0; q < 3; q++) { f[3 * i + q] += c * d[q]; f[3 * j + q] -= c * d[q]; }
        }
    }
    return U;
}

/* SHAKE: move x so that every constraint has its length, displacements along the pre-move bond vectors of x0. */
static void shake(const orc_mol *m, const double *x0, double *x, double tol) {
    for (int it = 0; it < 500; it++) {
        int done = 1;
        for (int c = 0; c < m->n_cons; c++) {
            const double *p = m->cons + 3 * c;
            const int i = (int)p[0], j = (int)p[1];
            const double d2 = p[2] * p[2];
            double r[3], r0[3];
            for (int q = 0; q < 3; q++) { r[q] = x[3 * i + q] - x[3 * j + q]; r0[q] = x0[3 * i + q] - x0[3 * j + q]; }
            const double diff = d2 - dot(r, r);
            if (fabs(diff) > tol * d2) {
                done = 0;
                const double wi = 1.0 / m->mass[i], wj = 1.0 / m->mass[j];
                const double g = diff / (2.0 * (wi + wj) * dot(r, r0));
                for (int q = 0; q < 3; q++) { x[3 * i + q] += g * wi * r0[q]; x[3 * j + q] -= g * wj * r0[q]; }
            }
        }
        if (done) break;
    }
}

/* RATTLE: remove the velocity components along the constraints. */
static void rattle(const orc_mol *m, const double *x, double *v, double tol) {
    for (int it = 0; it < 500; it++) {
        int done = 1;
        for (int c = 0; c < m->n_cons; c++) {
            const double *p = m->cons + 3 * c;
            const int i = (int)p[0], j = (int)p[1];
            double r[3], dv[3];
            for (int q = 0; q < 3; q++) { r[q] = x[3 * i + q] - x[3 * j + q]; dv[q] = v[3 * i + q] - v[3 * j + q]; }
            const double rv = dot(r, dv), r2 = dot(r, r);
            if (fabs(rv) > tol * r2) {   /* (units of 1/ps: relative rate of change of the squared length / 2) */
                done = 0;
                const double wi = 1.0 / m->mass[i], wj = 1.0 / m->mass[j];
                const double g = rv / ((wi + wj) * r2);
                for (int q = 0; q < 3; q++) { v[3 * i + q] -= g * wi * r[q]; v[3 * j + q] += g * wj * r[q]; }
            }
        }
        if (done) break;
    }
}

static void remove_cm(const orc_mol *m, double *v) {
    double p[3] = {0, 0, 0}, mt = 0;
    for (int i = 0; i < m->n_atoms; i++) { mt += m->mass[i]; for (int q = 0; q < 3; q++) p[q] += m->mass[i] * v[3 * i + q]; }
    for (int i = 0; i < m->n_atoms; i++) for (int q = 0; q < 3; q++) v[3 * i + q] -= p[q] / mt;
}

/* n_steps of the splitting `program` (characters V, R, O) with injected standard normals noise[n_steps*nO][n][3]. */
ORC_API double orc_mol_langevin(const orc_mol *m, double *x, double *v, const double *noise, double kT, double dt, double gamma,
                        int n_steps, const char *program, double tol) {
    const int n = m->n_atoms;
    int nV = 0, nR = 0, nO = 0;
    for (const char *q = program; *q; q++) { if (*q == 'V') nV++; else if (*q == 'R') nR++; else if (*q == 'O') nO++; }
    const double hO = dt / (nO > 0 ? nO : 1);
    const double a = exp(-gamma * hO), b = sqrt(1.0 - exp(-2.0 * gamma * hO));
    double *f = (double *)malloc(sizeof(double) * 3 * n), *x1 = (double *)malloc(sizeof(double) * 3 * n);
    int f_valid = 0;
    long oc = 0;
    for (int s = 0; s < n_steps; s++) {
        if (m->remove_cm) remove_cm(m, v);
        for (const char *q = program; *q; q++) {
            if (*q == 'V') {
                if (!f_valid) { orc_mol_energy(m, x, f); f_valid = 1; }
                const double h = dt / nV;
                for (int i = 0; i < n; i++) for (int c = 0; c < 3; c++) v[3 * i + c] += h * f[3 * i + c] / m->mass[i];
                if (m->n_cons) rattle(m, x, v, tol);
            } else if (*q == 'R') {
                const double h = dt / nR;
                memcpy(x1, x, sizeof(double) * 3 * n);           /* bond vectors before the move (SHAKE directions) */
                for (int i = 0; i < 3 * n; i++) x[i] += h * v[i];
                if (m->n_cons) {
                    double *xu = (double *)malloc(sizeof(double) * 3 * n);
                    memcpy(xu, x, sizeof(double) * 3 * n);       /* "x1" of integrators.py:1415: the unconstrained positions */
                    shake(m, x1, x, tol);
                    for (int i = 0; i < 3 * n; i++) v[i] += (x[i] - xu[i]) / h;
                    free(xu);
                    rattle(m, x, v, tol);
                }
                f_valid = 0;
            } else if (*q == 'O') {
                for (int i = 0; i < n; i++) {
                    const double sg = sqrt(kT / m->mass[i]);
                    for (int c = 0; c < 3; c++) v[3 * i + c] = a * v[3 * i + c] + b * sg * noise[(oc * n + i) * 3 + c];
                }
                oc++;
                if (m->n_cons) rattle(m, x, v, tol);
            }
        }
    }
    const double U = orc_mol_energy(m, x, NULL);
    free(f); free(x1);
    return U;
}

ORC_API void orc_mol_rattle(const orc_mol *m, const double *x, double *v, double tol) { rattle(m, x, v, tol); }

ORC_API double orc_mol_kinetic(const orc_mol *m, const double *v) {
    double ke = 0;
    for (int i = 0; i < m->n_atoms; i++) ke += 0.5 * m->mass[i] * (v[3 * i] * v[3 * i] + v[3 * i + 1] * v[3 * i + 1] + v[3 * i + 2] * v[3 * i + 2]);
    return ke;
}
