"""Shared builders for the tests: the alchemical LJ fluid of SURVEY.md section 8(d) in md units, the numpy
restatement of the device's Philox noise, and a guarded GPU-engine import."""
import os
import numpy as np

KB = 8.31446261815324e-3   # kJ/mol/K

# testsystems.LennardJonesFluid defaults (/root/reference/openmmtools/testsystems.py:1898-1909)
SIGMA = 0.34
EPSILON = 0.238 * 4.184
MASS = 39.9


def lj_setup(N=512, n_alch=10, reduced_density=0.05, seed=1, jitter=0.05):
    """Box, parameters and a jittered sub-random configuration (float32-rounded, as the reference stores it)."""
    L = (N * SIGMA ** 3 / reduced_density) ** (1.0 / 3.0)
    rc = 3.0 * SIGMA
    rs = rc - SIGMA
    rng = np.random.default_rng(seed)
    x = rng.random((N, 3)) * L
    # push apart overlapping atoms crudely: reject points closer than 0.9 sigma
    pts = []
    while len(pts) < N:
        c = rng.random(3) * L
        ok = True
        for q in pts:
            d = c - q
            d -= L * np.round(d / L)
            if d @ d < (0.9 * SIGMA) ** 2:
                ok = False
                break
        if ok:
            pts.append(c)
    x = np.array(pts).astype(np.float32).astype(np.float64)
    alch = np.zeros(N, np.uint8)
    alch[:n_alch] = 1
    return dict(N=N, L=L, rc=rc, rs=rs, x=x, alch=alch, sigma=np.full(N, SIGMA), eps=np.full(N, EPSILON),
                mass=np.full(N, MASS))


def oracle_system(s, annihilate=False, alpha=0.5, a=1.0, b=1.0, c=6.0, use_switch=True):
    from oracle import oracle
    return oracle.LJSystem(s['sigma'], s['eps'], s['mass'], s['alch'], (s['L'],) * 3, s['rc'], s['rs'],
                           use_switch=use_switch, alpha=alpha, a=a, b=b, c=c, annihilate_sterics=annihilate)


def gpu_engine(*args, **kw):
    from openmmtools_b200._engine import Engine
    return Engine(*args, **kw)


# ---- Philox4x32-10 + Box-Muller exactly as openmmtools_b200/csrc/rx_dynamics.cu -------------------------
def philox4x32_10(ctr, key):
    """ctr: uint32 array [..., 4]; key: (k0, k1). Vectorised."""
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        n0 = ((p1 >> np.uint64(32)) ^ c[1] ^ k0) & mask
        n1 = p1 & mask
        n2 = ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & mask
        n3 = p0 & mask
        c = [n0, n1, n2, n3]
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return np.stack(c, axis=-1).astype(np.uint32)


def philox_normal3(r):
    u = ((r >> 8).astype(np.float64) + 0.5) / 16777216.0
    ra = np.sqrt(-2.0 * np.log(u[..., 0]))
    rb = np.sqrt(-2.0 * np.log(u[..., 2]))
    return np.stack([ra * np.cos(2 * np.pi * u[..., 1]), ra * np.sin(2 * np.pi * u[..., 1]),
                     rb * np.cos(2 * np.pi * u[..., 3])], axis=-1)


def device_noise(seed, iteration, replica, N, n_osteps):
    """Standard normals the device uses for replica `replica`: [n_osteps, N, 3]."""
    key = (seed & 0xFFFFFFFF, ((seed >> 32) ^ (iteration >> 32)) & 0xFFFFFFFF)
    ctr = np.zeros((n_osteps, N, 4), np.uint32)
    ctr[..., 0] = np.arange(N)[None, :]
    ctr[..., 1] = np.arange(n_osteps)[:, None]
    ctr[..., 2] = replica
    ctr[..., 3] = iteration & 0xFFFFFFFF
    return philox_normal3(philox4x32_10(ctr, key))


def device_reassign_noise(seed, iteration, replica, N):
    key = (seed & 0xFFFFFFFF, ((seed >> 32) ^ (iteration >> 32)) & 0xFFFFFFFF)
    ctr = np.zeros((N, 4), np.uint32)
    ctr[:, 0] = np.arange(N); ctr[:, 1] = 0x80000000; ctr[:, 2] = replica; ctr[:, 3] = iteration & 0xFFFFFFFF
    return philox_normal3(philox4x32_10(ctr, key))
