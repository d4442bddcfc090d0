// rx_molecule.cuh -- a small molecule in vacuum per replica (included at the end of rx_dynamics.cu).
//
// The system testsystems.AlanineDipeptideVacuum builds (/root/reference/openmmtools/testsystems.py:3352-3388:
// AmberPrmtopFile.createSystem(implicitSolvent=None, constraints=HBonds, nonbondedCutoff=None)) propagated by the constrained
// Langevin splitting of /root/reference/openmmtools/integrators.py:1404-1460:
//   R  x1 = x + (dt/nR) v;  x = constrain(x1);  v += (x - x1) / (dt/nR);  constrain velocities      (:1404-1422)
//   V  v += (dt/nV) f / m;  constrain velocities                                                    (:1424-1447)
//   O  v = a v + b sqrt(kT/m) xi;  constrain velocities                                             (:1449-1460)
// with the centre-of-mass velocity removed at the start of every step (CMMotionRemover through addUpdateContextState).
//
// One warp per replica, one lane per atom (N <= 32), everything in f64 (22 atoms: the step is a latency chain, not
// arithmetic).  No atomics: a lane computes the force on ITS atom from the terms that contain it (per-atom term lists built
// on the host; an angle is evaluated by its three lanes, a torsion by four), so results are bit-reproducible.  Constraints
// are solved cluster by cluster (connected components of the constraint graph, e.g. a CH3 group), one lane per cluster,
// Gauss-Seidel in list order -- the same arithmetic, in the same order, as oracle/rx_oracle_mol.c.
#pragma once

#define MOL_MAX_ATOMS 32
#define MOL_ONE_4PI_EPS0 138.935456

struct MolBond { int j, pad; double K, r0; };
struct MolAngle { int i, j, k, role; double K, t0; };
struct MolTorsion { int i, j, k, l, n, role; double phase, kk; };
struct MolExc { int j, pad; double qq, sig, eps; };
struct MolCons { int i, j; double d; };

struct MolDev {
    int n, n_clusters, remove_cm, pad;
    double tol;
    const double *mass, *charge, *sigma, *eps;
    const int *b_off, *a_off, *t_off, *x_off, *c_off;
    const MolBond *bonds;
    const MolAngle *angles;
    const MolTorsion *torsions;
    const MolExc *exc;
    const MolCons *cons;
    const unsigned *nb_mask;     // bit b of nb_mask[a]: the pair (a, b) has the plain Coulomb + LJ interaction
};

__device__ __forceinline__ void mol_cross(const double *a, const double *b, double *c) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double mol_dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// Force on atom a (ENERGY: instead the energy of the terms this atom owns: bonds/exceptions/pairs with the larger partner
// index, angles and torsions in which it has role 0).  X: positions of all atoms in shared memory.
template <bool ENERGY>
__device__ double mol_atom(const MolDev &m, const double (*X)[3], int a, double *f) {
    double e = 0.0, fx = 0.0, fy = 0.0, fz = 0.0;
    const double xa[3] = {X[a][0], X[a][1], X[a][2]};
    for (int q = m.b_off[a]; q < m.b_off[a + 1]; q++) {
        const MolBond b = m.bonds[q];
        const double d[3] = {xa[0] - X[b.j][0], xa[1] - X[b.j][1], xa[2] - X[b.j][2]};
        const double r = sqrt(mol_dot(d, d)), dr = r - b.r0;
        if (ENERGY) { if (b.j > a) e += 0.5 * b.K * dr * dr; }
        else { const double c = -b.K * dr / r; fx += c * d[0]; fy += c * d[1]; fz += c * d[2]; }
    }
    for (int q = m.a_off[a]; q < m.a_off[a + 1]; q++) {
        const MolAngle g = m.angles[q];
        double u[3], v[3];
        for (int c = 0; c < 3; c++) { u[c] = X[g.i][c] - X[g.j][c]; v[c] = X[g.k][c] - X[g.j][c]; }
        const double ru = sqrt(mol_dot(u, u)), rv = sqrt(mol_dot(v, v));
        double cs = mol_dot(u, v) / (ru * rv);
        cs = fmin(1.0, fmax(-1.0, cs));
        const double dt = acos(cs) - g.t0;
        if (ENERGY) { if (g.role == 0) e += 0.5 * g.K * dt * dt; }
        else {
            const double sn = sqrt(1.0 - cs * cs);
            const double gg = (sn > 1e-12) ? g.K * dt / sn : 0.0;
            double o[3];
            for (int c = 0; c < 3; c++) {
                const double fi = gg * (v[c] / (ru * rv) - cs * u[c] / (ru * ru));
                const double fk = gg * (u[c] / (ru * rv) - cs * v[c] / (rv * rv));
                o[c] = g.role == 0 ? fi : (g.role == 2 ? fk : -(fi + fk));
            }
            fx += o[0]; fy += o[1]; fz += o[2];
        }
    }
    for (int q = m.t_off[a]; q < m.t_off[a + 1]; q++) {
        const MolTorsion t = m.torsions[q];
        double rij[3], rkj[3], rkl[3], mm[3], nn[3];
        for (int c = 0; c < 3; c++) { rij[c] = X[t.i][c] - X[t.j][c]; rkj[c] = X[t.k][c] - X[t.j][c]; rkl[c] = X[t.k][c] - X[t.l][c]; }
        mol_cross(rij, rkj, mm); mol_cross(rkj, rkl, nn);
        const double nrkj = sqrt(mol_dot(rkj, rkj));
        const double phi = atan2(nrkj * mol_dot(rij, nn), mol_dot(mm, nn));
        if (ENERGY) { if (t.role == 0) e += t.kk * (1.0 + cos((double)t.n * phi - t.phase)); }
        else {
            const double dU = -t.kk * (double)t.n * sin((double)t.n * phi - t.phase);
            const double m2 = mol_dot(mm, mm), n2 = mol_dot(nn, nn);
            const double pp = mol_dot(rij, rkj) / (nrkj * nrkj), qq = mol_dot(rkl, rkj) / (nrkj * nrkj);
            double o[3];
            for (int c = 0; c < 3; c++) {
                const double fi = -dU * nrkj / m2 * mm[c], fl = dU * nrkj / n2 * nn[c];
                const double sv = pp * fi - qq * fl;
                o[c] = t.role == 0 ? fi : (t.role == 1 ? sv - fi : (t.role == 2 ? -sv - fl : fl));
            }
            fx += o[0]; fy += o[1]; fz += o[2];
        }
    }
    const unsigned mask = m.nb_mask[a];
    const double qa = MOL_ONE_4PI_EPS0 * m.charge[a], sa = m.sigma[a], ea = m.eps[a];
    for (int b = 0; b < m.n; b++) {
        if (!((mask >> b) & 1u)) continue;
        if (ENERGY && b < a) continue;
        const double d[3] = {xa[0] - X[b][0], xa[1] - X[b][1], xa[2] - X[b][2]};
        const double r2 = mol_dot(d, d), r = sqrt(r2);
        const double qq = qa * m.charge[b], s = 0.5 * (sa + m.sigma[b]), ee = sqrt(ea * m.eps[b]);
        const double s2 = s * s / r2, s6 = s2 * s2 * s2;
        if (ENERGY) e += qq / r + 4.0 * ee * (s6 * s6 - s6);
        else { const double c = (qq / r + 24.0 * ee * (2.0 * s6 * s6 - s6)) / r2; fx += c * d[0]; fy += c * d[1]; fz += c * d[2]; }
    }
    for (int q = m.x_off[a]; q < m.x_off[a + 1]; q++) {
        const MolExc x = m.exc[q];
        if (ENERGY && x.j < a) continue;
        const double d[3] = {xa[0] - X[x.j][0], xa[1] - X[x.j][1], xa[2] - X[x.j][2]};
        const double r2 = mol_dot(d, d), r = sqrt(r2);
        const double qq = MOL_ONE_4PI_EPS0 * x.qq;
        const double s2 = x.sig * x.sig / r2, s6 = s2 * s2 * s2;
        if (ENERGY) e += qq / r + 4.0 * x.eps * (s6 * s6 - s6);
        else { const double c = (qq / r + 24.0 * x.eps * (2.0 * s6 * s6 - s6)) / r2; fx += c * d[0]; fy += c * d[1]; fz += c * d[2]; }
    }
    if (!ENERGY) { f[0] = fx; f[1] = fy; f[2] = fz; }
    return e;
}

// SHAKE on one cluster: positions X moved so that its constraints have their lengths, along the bond vectors of XO.
__device__ void mol_shake_cluster(const MolDev &m, int c, const double (*XO)[3], double (*X)[3]) {
    for (int it = 0; it < 500; it++) {
        bool done = true;
        for (int q = m.c_off[c]; q < m.c_off[c + 1]; q++) {
            const MolCons k = m.cons[q];
            const double d2 = k.d * k.d;
            double r[3], r0[3];
            for (int a = 0; a < 3; a++) { r[a] = X[k.i][a] - X[k.j][a]; r0[a] = XO[k.i][a] - XO[k.j][a]; }
            const double diff = d2 - mol_dot(r, r);
            if (fabs(diff) > m.tol * d2) {
                done = false;
                const double wi = 1.0 / m.mass[k.i], wj = 1.0 / m.mass[k.j];
                const double g = diff / (2.0 * (wi + wj) * mol_dot(r, r0));
                for (int a = 0; a < 3; a++) { X[k.i][a] += g * wi * r0[a]; X[k.j][a] -= g * wj * r0[a]; }
            }
        }
        if (done) break;
    }
}

// RATTLE on one cluster: the velocity components along its constraints are removed.
__device__ void mol_rattle_cluster(const MolDev &m, int c, const double (*X)[3], double (*V)[3]) {
    for (int it = 0; it < 500; it++) {
        bool done = true;
        for (int q = m.c_off[c]; q < m.c_off[c + 1]; q++) {
            const MolCons k = m.cons[q];
            double r[3], dv[3];
            for (int a = 0; a < 3; a++) { r[a] = X[k.i][a] - X[k.j][a]; dv[a] = V[k.i][a] - V[k.j][a]; }
            const double rv = mol_dot(r, dv), r2 = mol_dot(r, r);
            if (fabs(rv) > m.tol * r2) {
                done = false;
                const double wi = 1.0 / m.mass[k.i], wj = 1.0 / m.mass[k.j];
                const double g = rv / ((wi + wj) * r2);
                for (int a = 0; a < 3; a++) { V[k.i][a] -= g * wi * r[a]; V[k.j][a] += g * wj * r[a]; }
            }
        }
        if (done) break;
    }
}

__device__ __forceinline__ double mol_warp_sum(double v) {   // fixed tree: bit-reproducible
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One warp per owned replica.  pos / vel: double[kloc][n][3].
__global__ void __launch_bounds__(32) k_propagate_mol(MolDev m, DynParams p, const StateDev *__restrict__ states,
                                                      const int *__restrict__ perm, double *__restrict__ pos,
                                                      double *__restrict__ vel, int k0, uint2 key, uint32_t iteration,
                                                      int reassign, double *__restrict__ pot, double *__restrict__ kin,
                                                      int *__restrict__ nan_flag, const int *__restrict__ only) {
    __shared__ double X[MOL_MAX_ATOMS][3], XO[MOL_MAX_ATOMS][3], V[MOL_MAX_ATOMS][3];
    const int r = blockIdx.x, k = k0 + r, a = threadIdx.x, n = m.n;
    if (only && !only[k]) return;
    const bool active = a < n;
    const StateDev st = states[perm[k]];
    const double mass = active ? m.mass[a] : 1.0, sg = sqrt(st.kT / mass);
    double mt = mol_warp_sum(active ? mass : 0.0);
    if (active) {
        for (int c = 0; c < 3; c++) { X[a][c] = pos[((size_t)r * n + a) * 3 + c]; V[a][c] = vel[((size_t)r * n + a) * 3 + c]; }
        if (reassign) {   // context.setVelocitiesToTemperature, mcmc.py:711
            const float3 g = philox_normal3(philox4x32_10(make_uint4(a, 0x80000000u, k, iteration), key));
            V[a][0] = sg * g.x; V[a][1] = sg * g.y; V[a][2] = sg * g.z;
        }
    }
    __syncwarp();
    if (a < m.n_clusters) mol_rattle_cluster(m, a, X, V);   // incoming velocities obey the constraints
    __syncwarp();
    int nV = p.nV, nR = p.nR;
    double f[3] = {0, 0, 0};
    bool f_valid = false;
    uint32_t ocount = 0;
    for (int s = 0; s < p.n_steps; s++) {
        if (m.remove_cm) {
            double px = mol_warp_sum(active ? mass * V[a][0] : 0.0), py = mol_warp_sum(active ? mass * V[a][1] : 0.0),
                   pz = mol_warp_sum(active ? mass * V[a][2] : 0.0);
            if (active) { V[a][0] -= px / mt; V[a][1] -= py / mt; V[a][2] -= pz / mt; }
            __syncwarp();
        }
        for (int q = 0; q < p.n_prog; q++) {
            const char op = p.prog[q];
            if (op == 'V') {
                if (!f_valid) { if (active) mol_atom<false>(m, X, a, f); f_valid = true; }
                const double h = (double)p.dt_d / nV;
                if (active) for (int c = 0; c < 3; c++) V[a][c] += h * f[c] / mass;
                __syncwarp();
                if (a < m.n_clusters) mol_rattle_cluster(m, a, X, V);
                __syncwarp();
            } else if (op == 'R') {
                const double h = (double)p.dt_d / nR;
                double xu[3] = {0, 0, 0};
                if (active) for (int c = 0; c < 3; c++) { XO[a][c] = X[a][c]; X[a][c] += h * V[a][c]; xu[c] = X[a][c]; }
                __syncwarp();
                if (m.n_clusters) {
                    if (a < m.n_clusters) mol_shake_cluster(m, a, XO, X);
                    __syncwarp();
                    if (active) for (int c = 0; c < 3; c++) V[a][c] += (X[a][c] - xu[c]) / h;
                    __syncwarp();
                    if (a < m.n_clusters) mol_rattle_cluster(m, a, X, V);
                    __syncwarp();
                }
                f_valid = false;
            } else {   // 'O'
                if (active) {
                    const float3 g = philox_normal3(philox4x32_10(make_uint4(a, ocount, k, iteration), key));
                    V[a][0] = p.a_d * V[a][0] + p.b_d * sg * (double)g.x;
                    V[a][1] = p.a_d * V[a][1] + p.b_d * sg * (double)g.y;
                    V[a][2] = p.a_d * V[a][2] + p.b_d * sg * (double)g.z;
                }
                ocount++;
                __syncwarp();
                if (a < m.n_clusters) mol_rattle_cluster(m, a, X, V);
                __syncwarp();
            }
        }
    }
    const double U = mol_warp_sum(active ? mol_atom<true>(m, X, a, nullptr) : 0.0);
    const double KE = mol_warp_sum(active ? 0.5 * mass * (V[a][0] * V[a][0] + V[a][1] * V[a][1] + V[a][2] * V[a][2]) : 0.0);
    const bool bad = active && !(isfinite(X[a][0]) && isfinite(X[a][1]) && isfinite(X[a][2]) && isfinite(V[a][0]) &&
                                 isfinite(V[a][1]) && isfinite(V[a][2]));
    const unsigned any_bad = __ballot_sync(0xffffffffu, bad);
    if (a == 0) { pot[k] = U + st.offset; kin[k] = KE; nan_flag[k] = (any_bad || !isfinite(U)) ? 1 : 0; }
    if (active)
        for (int c = 0; c < 3; c++) { pos[((size_t)r * n + a) * 3 + c] = X[a][c]; vel[((size_t)r * n + a) * 3 + c] = V[a][c]; }
}

// u[k][l] = beta_l (U(x_k) + offset_l): the states of a molecule differ in temperature only (parallel tempering,
// paralleltempering.py:175-237: one potential evaluation per replica, scaled by beta_l).
__global__ void __launch_bounds__(32) k_energy_mol(MolDev m, const StateDev *__restrict__ states, int n_states,
                                                   const double *__restrict__ pos, int k0, double *__restrict__ u_out) {
    __shared__ double X[MOL_MAX_ATOMS][3];
    const int r = blockIdx.x, k = k0 + r, a = threadIdx.x, n = m.n;
    const bool active = a < n;
    if (active) for (int c = 0; c < 3; c++) X[a][c] = pos[((size_t)r * n + a) * 3 + c];
    __syncwarp();
    const double U = mol_warp_sum(active ? mol_atom<true>(m, X, a, nullptr) : 0.0);
    for (int l = a; l < n_states; l += 32) u_out[(size_t)k * n_states + l] = states[l].beta * (U + states[l].offset);
}

__global__ void k_randomize_velocities_mol(MolDev m, const StateDev *__restrict__ states, const int *__restrict__ perm,
                                           double *__restrict__ vel, int k0, uint2 key, uint32_t stream_id) {
    const int r = blockIdx.x, k = k0 + r, a = threadIdx.x;
    if (a >= m.n) return;
    const StateDev st = states[perm[k]];
    const double sv = sqrt(st.kT / m.mass[a]);
    const float3 g = philox_normal3(philox4x32_10(make_uint4(a, 0xC0000000u, k, stream_id), key));
    vel[((size_t)r * m.n + a) * 3 + 0] = sv * g.x;
    vel[((size_t)r * m.n + a) * 3 + 1] = sv * g.y;
    vel[((size_t)r * m.n + a) * 3 + 2] = sv * g.z;
}

__global__ void k_restore_failed_mol(const int *__restrict__ nan_flag, int *__restrict__ retry, int k0, int N,
                                     const double *__restrict__ pos_snap, const double *__restrict__ vel_snap,
                                     double *__restrict__ pos, double *__restrict__ vel) {
    const int r = blockIdx.x, k = k0 + r;
    const int failed = nan_flag[k];
    if (threadIdx.x == 0) retry[k] = failed;
    if (!failed) return;
    for (int q = threadIdx.x; q < 3 * N; q += blockDim.x) {
        pos[(size_t)r * N * 3 + q] = pos_snap[(size_t)r * N * 3 + q];
        vel[(size_t)r * N * 3 + q] = vel_snap[(size_t)r * N * 3 + q];
    }
}

