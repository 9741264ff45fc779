"""Synthetic haloes for parity tests and benchmarks.

Units follow the reference (pyHALMA.py:523-526): positions in comoving Mpc, masses in
Msun, velocities in km/s; G = 4.3e-9 (km/s)^2 Mpc/Msun (halo_gas.py:459-465).
Arrays are float64, like the reference's particle arrays; the float32 cast happens at
the boundary (halo_gas.py:172-178).  A centre offset and a bulk velocity are added so
float32 cancellation in (x_j - x_i) is realistic.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

G = (4.3 * 1e-3) * 1e-6
KPC = 1e-3
CENTRE = (3.0, -7.0, 11.0)
BULK_V = (150.0, -80.0, 40.0)
# level-9 AMR cell of a 40 Mpc box with a 128^3 base grid (pyHALMA.dat:14-15, pyHALMA.py:1478)
CELL = 40.0 / 128 / 2 ** 9
BASE_SEED = 20240215


def rng_for(cfg: int, extra: int = 0) -> np.random.Generator:
    return np.random.default_rng(BASE_SEED + cfg + 1000003 * extra)


def _isotropic(n, rng):
    mu = rng.uniform(-1.0, 1.0, n)
    ph = rng.uniform(0.0, 2 * np.pi, n)
    s = np.sqrt(1.0 - mu * mu)
    return s * np.cos(ph), s * np.sin(ph), mu


@dataclass
class Particles:
    x: np.ndarray
    y: np.ndarray
    z: np.ndarray
    mass: np.ndarray
    vx: Optional[np.ndarray] = None
    vy: Optional[np.ndarray] = None
    vz: Optional[np.ndarray] = None
    temp: Optional[np.ndarray] = None

    def __len__(self):
        return len(self.x)

    def pos_mass(self):
        """(mass, x, y, z): the argument order of the reference kernel."""
        return self.mass, self.x, self.y, self.z


def plummer_stars(n: int, a: float, m_part: float, rng, centre=CENTRE, bulk_v=BULK_V,
                  interloper_frac: float = 0.1, interloper_boost: float = 8.0) -> Particles:
    """Equal-mass Plummer sphere; speeds by Aarseth-Henon-Wielen rejection on
    q^2 (1-q^2)^{7/2} times the escape speed; a fraction gets |v| boosted so the unbinding
    takes more than one pass."""
    M = n * m_part
    u = rng.uniform(1e-10, 1.0, n)
    r = a / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
    r = np.minimum(r, 50 * a)
    ux, uy, uz = _isotropic(n, rng)
    q = np.empty(n)
    todo = np.arange(n)
    while len(todo):
        qq = rng.uniform(0, 1, len(todo))
        yy = rng.uniform(0, 0.1, len(todo))
        ok = yy < qq * qq * (1 - qq * qq) ** 3.5
        q[todo[ok]] = qq[ok]
        todo = todo[~ok]
    vesc = np.sqrt(2 * G * M / a) * (1 + (r / a) ** 2) ** (-0.25)
    v = q * vesc
    if interloper_frac > 0:
        k = rng.uniform(0, 1, n) < interloper_frac
        v = np.where(k, v * interloper_boost, v)
    wx, wy, wz = _isotropic(n, rng)
    return Particles(centre[0] + r * ux, centre[1] + r * uy, centre[2] + r * uz,
                     np.full(n, float(m_part)), bulk_v[0] + v * wx, bulk_v[1] + v * wy,
                     bulk_v[2] + v * wz)


def _nfw_tables(rs, c, M):
    xg = np.geomspace(1e-4, c, 4096)
    mu = np.log1p(xg) - xg / (1 + xg)
    mu_c = np.log1p(c) - c / (1 + c)
    Mr = M * mu / mu_c
    rho = 1.0 / (xg * (1 + xg) ** 2)
    # isotropic Jeans: sigma^2(r) = 1/rho * int_r^rvir rho G M(<s)/s^2 ds
    integrand = rho * G * Mr / (xg * rs) ** 2
    dr = np.diff(xg * rs)
    seg = 0.5 * (integrand[1:] + integrand[:-1]) * dr
    tail = np.concatenate((np.cumsum(seg[::-1])[::-1], [0.0]))
    sig2 = tail / rho
    return xg, mu / mu_c, np.sqrt(np.maximum(sig2, 0))


def nfw_stars(n: int, rs: float, c: float, m_part: float, rng, centre=CENTRE, bulk_v=BULK_V,
              interloper_frac: float = 0.05, interloper_boost: float = 7.0) -> Particles:
    """NFW profile truncated at r_vir = c rs; Gaussian isotropic velocities with the Jeans
    dispersion of the profile's own mass."""
    M = n * m_part
    xg, cdf, sig = _nfw_tables(rs, c, M)
    u = rng.uniform(cdf[0], 1.0, n)
    xr = np.interp(u, cdf, xg)
    r = xr * rs
    ux, uy, uz = _isotropic(n, rng)
    s = np.interp(xr, xg, sig)
    vx, vy, vz = (rng.normal(0, 1, n) * s for _ in range(3))
    if interloper_frac > 0:
        k = rng.uniform(0, 1, n) < interloper_frac
        f = np.where(k, interloper_boost, 1.0)
        vx, vy, vz = vx * f, vy * f, vz * f
    return Particles(centre[0] + r * ux, centre[1] + r * uy, centre[2] + r * uz,
                     np.full(n, float(m_part)), bulk_v[0] + vx, bulk_v[1] + vy, bulk_v[2] + vz)


def lattice_gas(n_target: int, spacing: float, rng, centre=CENTRE, bulk_v=BULK_V,
                core: Optional[float] = None, m_total: float = 1e9, two_levels: bool = False,
                sigma_v: float = 100.0, wind: float = 300.0) -> Particles:
    """Gas pseudo-particles at AMR cell centres (halo_gas.py:27-38): a cubic lattice clipped to
    a sphere, so every pair in a common x-, y- or z-plane trips the reference's exclusion
    predicate (particle_subroutines.f90:499-501).  Masses follow a beta profile, T is
    log-uniform 1e3..1e7 K, velocities are bulk + Gaussian + a wind on half the cells.
    two_levels: the inner half (by radius) sits on `spacing`, the rest on 2*spacing."""
    if n_target <= 0:
        e = np.zeros(0)
        return Particles(e, e.copy(), e.copy(), e.copy(), e.copy(), e.copy(), e.copy(), e.copy())

    def ball(n_cells, h, rmin=0.0, rmax=None):
        if rmax is None:
            vol = n_cells * h ** 3 + 4.0 / 3 * np.pi * rmin ** 3
            rmax = (3 * vol / (4 * np.pi)) ** (1.0 / 3) * 1.06 + h
        k = int(np.ceil(rmax / h)) + 1
        g = (np.arange(-k, k) + 0.5) * h
        X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
        rr = np.sqrt(X * X + Y * Y + Z * Z)
        keep = (rr < rmax) & (rr >= rmin)
        return X[keep], Y[keep], Z[keep], rr[keep], rmax

    if two_levels:
        n_in = n_target // 2
        xi, yi, zi, ri, rcut = ball(n_in, spacing)
        rcut = (rcut - spacing) / 1.06
        # the coarse shell starts on the fine/coarse boundary box to avoid overlapping cells
        rcut = np.floor(rcut / (2 * spacing)) * 2 * spacing
        sel = np.maximum.reduce([np.abs(xi), np.abs(yi), np.abs(zi)]) < rcut
        xi, yi, zi, ri = xi[sel], yi[sel], zi[sel], ri[sel]
        n_out = max(n_target - len(xi), 1)
        k = int(np.ceil(((n_out * (2 * spacing) ** 3 + (2 * rcut) ** 3) * 3 / (4 * np.pi)) ** (1 / 3)
                        / (2 * spacing))) + 1
        g = (np.arange(-k, k) + 0.5) * 2 * spacing
        X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
        outside = np.maximum.reduce([np.abs(X), np.abs(Y), np.abs(Z)]) >= rcut
        rr = np.sqrt(X * X + Y * Y + Z * Z)
        order = np.argsort(rr[outside], kind="stable")[:n_out]
        xo, yo, zo, ro = X[outside][order], Y[outside][order], Z[outside][order], rr[outside][order]
        x = np.concatenate((xi, xo)); y = np.concatenate((yi, yo)); z = np.concatenate((zi, zo))
        r = np.concatenate((ri, ro))
        vol = np.concatenate((np.full(len(xi), spacing ** 3), np.full(len(xo), (2 * spacing) ** 3)))
    else:
        x, y, z, r, _ = ball(n_target, spacing)
        order = np.argsort(r, kind="stable")[:n_target]
        x, y, z, r = x[order], y[order], z[order], r[order]
        vol = np.full(len(x), spacing ** 3)
    n = len(x)
    rc = core if core is not None else 0.2 * r.max()
    dens = (1 + (r / rc) ** 2) ** (-1.5 * 0.7)
    mass = dens * vol
    mass *= m_total / mass.sum()
    temp = 10 ** rng.uniform(3, 7, n)
    vx = bulk_v[0] + rng.normal(0, sigma_v, n)
    vy = bulk_v[1] + rng.normal(0, sigma_v, n)
    vz = bulk_v[2] + rng.normal(0, sigma_v, n)
    blown = rng.uniform(0, 1, n) < 0.5
    vx = vx + np.where(blown, wind, 0.0)
    # snap the centre to the lattice so cell centres stay on exact planes after the shift
    cx, cy, cz = (np.round(c / spacing) * spacing for c in centre)
    return Particles(cx + x, cy + y, cz + z, mass, vx, vy, vz, temp)


def dm_cloud(n: int, scale: float, m_part: float, rng, centre=CENTRE, two_species: bool = False) -> Particles:
    """Dark-matter particles around the halo (Hernquist-like radial distribution)."""
    u = rng.uniform(0, 0.98, n)
    r = scale * np.sqrt(u) / (1 - np.sqrt(u))
    ux, uy, uz = _isotropic(n, rng)
    mass = np.full(n, float(m_part))
    if two_species:
        light = rng.uniform(0, 1, n) < 0.7
        mass = np.where(light, m_part / 64.0, m_part)
    return Particles(centre[0] + r * ux, centre[1] + r * uy, centre[2] + r * uz, mass)


def add_coincident_pairs(p: Particles, n_pairs: int, rng) -> None:
    """Make a few stars share one float32 coordinate (stars closer than
    ~2 pc in a coordinate collapse to equal float32 and get excluded)."""
    n = len(p)
    if n < 2 * n_pairs + 2:
        return
    pick = rng.choice(n, 2 * n_pairs, replace=False)
    a, b = pick[:n_pairs], pick[n_pairs:]
    which = rng.integers(0, 3, n_pairs)
    for arr, k in ((p.x, 0), (p.y, 1), (p.z, 2)):
        sel = which == k
        arr[b[sel]] = arr[a[sel]]


# --------------------------------------------------------------------------------------
# The BASELINE.json configurations
# --------------------------------------------------------------------------------------
@dataclass
class HaloCase:
    """One halo: stars, gas cells, DM, plus the stellar bulk velocity the gas pass uses."""
    stars: Particles
    gas: Particles
    dm: Particles
    mass_dm_part: float = 8e7
    factor_v: float = 3.0          # pyHALMA.dat:38-39
    name: str = ""


def config1(n_star: int = 10_000, n_gas: int = 10_000, seed_extra: int = 0, n_dm: int = 2_000) -> HaloCase:
    """cfg1: Plummer stellar halo (a = 2 kpc, 1e6 Msun stars) + level-9 lattice gas."""
    rng = rng_for(1, seed_extra)
    stars = plummer_stars(n_star, 2 * KPC, 1e6, rng)
    add_coincident_pairs(stars, max(1, n_star // 2000), rng)
    gas = lattice_gas(n_gas, CELL, rng, m_total=0.2 * n_star * 1e6)
    dm = dm_cloud(n_dm, 6 * KPC, 8e7 / 64, rng)
    return HaloCase(stars, gas, dm, mass_dm_part=8e7, name="cfg1")


def config2(n_star: int = 200_000, n_gas: int = 500_000, seed_extra: int = 0, n_dm: int = 0) -> HaloCase:
    """cfg2: NFW galaxy halo (r_s = 5 kpc, c = 10) + two-level lattice gas."""
    rng = rng_for(2, seed_extra)
    stars = nfw_stars(n_star, 5 * KPC, 10.0, 1e6, rng)
    add_coincident_pairs(stars, max(1, n_star // 20000), rng)
    gas = lattice_gas(n_gas, CELL, rng, m_total=0.15 * n_star * 1e6, two_levels=True)
    dm = dm_cloud(n_dm, 30 * KPC, 8e7 / 8, rng) if n_dm else Particles(*(np.zeros(0) for _ in range(4)))
    return HaloCase(stars, gas, dm, mass_dm_part=8e7, name="cfg2")


def powerlaw_sizes(n_halo: int, nmin: int, nmax: int, alpha: float, rng) -> np.ndarray:
    """dN/dN_h ~ N_h^-alpha on [nmin, nmax] by inverse CDF."""
    u = rng.uniform(0, 1, n_halo)
    e = 1.0 - alpha
    n = (nmin ** e + u * (nmax ** e - nmin ** e)) ** (1.0 / e)
    return np.clip(np.round(n).astype(np.int64), nmin, nmax)


@dataclass
class Catalogue:
    """Concatenated member arrays + CSR offsets, the input of unbind_catalogue."""
    offsets: np.ndarray
    x: np.ndarray
    y: np.ndarray
    z: np.ndarray
    vx: np.ndarray
    vy: np.ndarray
    vz: np.ndarray
    mass: np.ndarray
    meta: Dict = field(default_factory=dict)

    @property
    def n_halo(self) -> int:
        return len(self.offsets) - 1

    def sizes(self) -> np.ndarray:
        return np.diff(self.offsets)

    def halo(self, h: int):
        a, b = self.offsets[h], self.offsets[h + 1]
        return tuple(arr[a:b] for arr in (self.x, self.y, self.z, self.vx, self.vy, self.vz, self.mass))

    def subset(self, halo_ids) -> "Catalogue":
        halo_ids = np.asarray(halo_ids, dtype=np.int64)
        sizes = self.sizes()[halo_ids]
        off = np.concatenate(([0], np.cumsum(sizes)))
        if len(halo_ids):
            take = np.concatenate([np.arange(self.offsets[h], self.offsets[h + 1]) for h in halo_ids])
        else:
            take = np.zeros(0, np.int64)
        return Catalogue(off, *(arr[take] for arr in (self.x, self.y, self.z, self.vx, self.vy,
                                                      self.vz, self.mass)),
                         meta=dict(self.meta, parent_ids=halo_ids))


def config3(n_halo: int = 10_000, nmin: int = 100, nmax: int = 100_000, alpha: float = 1.9,
            seed_extra: int = 0, box: float = 40.0) -> Catalogue:
    """cfg3: catalogue of Plummer haloes, power-law sizes, a ~ N^(1/3), random centres."""
    rng = rng_for(3, seed_extra)
    sizes = powerlaw_sizes(n_halo, nmin, nmax, alpha, rng)
    off = np.concatenate(([0], np.cumsum(sizes)))
    tot = int(off[-1])
    cols = [np.empty(tot) for _ in range(7)]
    for h, n in enumerate(sizes):
        centre = rng.uniform(0.05 * box, 0.95 * box, 3) - box / 2
        bulk = rng.normal(0, 200.0, 3)
        a = 2 * KPC * (n / 1e4) ** (1.0 / 3)
        p = plummer_stars(int(n), a, 1e6, rng, centre=centre, bulk_v=bulk)
        s = slice(off[h], off[h + 1])
        for c, arr in zip(cols, (p.x, p.y, p.z, p.vx, p.vy, p.vz, p.mass)):
            c[s] = arr
    return Catalogue(off, *cols, meta={"cfg": 3, "alpha": alpha, "sum_n": tot,
                                       "sum_n2": float(np.sum(sizes.astype(np.float64) ** 2))})


def config4(n_star: int = 2_000_000, seed_extra: int = 0) -> Particles:
    """cfg4: cluster-scale stellar halo, half Plummer (a = 30 kpc) half NFW, no externals."""
    rng = rng_for(4, seed_extra)
    n1 = n_star // 2
    p = plummer_stars(n1, 30 * KPC, 1e6, rng)
    q = nfw_stars(n_star - n1, 60 * KPC, 6.0, 1e6, rng)
    return Particles(*(np.concatenate((getattr(p, k), getattr(q, k)))
                       for k in ("x", "y", "z", "mass", "vx", "vy", "vz")))


def config5(n_gas: int = 10_000_000, n_star: int = 2_000_000, n_dm: int = 500_000,
            seed_extra: int = 0) -> HaloCase:
    """cfg5: cfg4 stars as sources + lattice gas (targets and self-sources) + DM."""
    rng = rng_for(5, seed_extra)
    stars = config4(n_star, seed_extra)
    gas = lattice_gas(n_gas, 4 * CELL, rng, m_total=0.1 * n_star * 1e6, two_levels=True)
    dm = dm_cloud(n_dm, 200 * KPC, 8e7, rng)
    return HaloCase(stars, gas, dm, mass_dm_part=8e7, name="cfg5")


# --------------------------------------------------------------------------------------
# A synthetic MASCLET-like AMR snapshot: the input format of the gather step before the
# hot path (python_scripts/halo_gas.py:56-141, 223-277)
# --------------------------------------------------------------------------------------
@dataclass
class AmrSnapshot:
    """The objects pyHALMA.py hands to halo_gas.st_gas_dm_particles_inside (:1027-1037), with
    only the entries that function reads filled in:
      grid_data[2] nl, [5] npatch per level (npatch[0] = 0), [6..8] patchnx/ny/nz, [12..14]
      patchrx/ry/rz (centre of the patch's first PARENT cell), index 0 = the base grid;
      gas_data[0] delta, [1..3] velocity in units of c, [4] temperature, [5] cr0amr (True =
      not refined), [6] solapst (True = not overlapped): per-patch (nx, ny, nz) arrays,
      float32 / bool, Fortran-ordered like masclet_framework's reader returns them;
      masclet_dm_data[0..3] x, y, z, mass; masclet_st_data[0..2] position, [6] mass, [9] id."""
    L: float
    ncoarse: int
    grid_data: list
    gas_data: list
    masclet_dm_data: list
    masclet_st_data: list
    rho_B: float
    rete: float
    centre: tuple

    @property
    def n_cells(self) -> int:
        return int(sum(int(a.size) for a in self.gas_data[0][1:]))


def amr_snapshot(n_levels: int = 5, patches_per_level: int = 3, n_dm: int = 20_000, n_st: int = 30_000,
                 seed_extra: int = 0, L: float = 40.0, ncoarse: int = 128, centre=CENTRE,
                 max_cells: int = 28, base_cells: int = 4, background_frac: float = 0.0) -> AmrSnapshot:
    """Nested patches around `centre`: level l has `patches_per_level` patches of up to
    max_cells^3 cells of size (L/ncoarse)/2^l, aligned to the parent grid, partly overlapping;
    cr0amr marks cells covered by a finer patch, solapst cells already covered by an earlier
    patch of the same level; a few flags are flipped at random.  background_frac of the DM and
    star particles is spread uniformly over the whole box (a cosmological volume around the halo)."""
    rng = rng_for(6, seed_extra)
    c = np.asarray(centre, dtype=np.float64)
    npatch = [0]
    nx, ny, nz = [ncoarse], [ncoarse], [ncoarse]
    rx, ry, rz = [-L / 2 + L / ncoarse / 2] * 1, [-L / 2 + L / ncoarse / 2], [-L / 2 + L / ncoarse / 2]
    level_of = [0]
    lo_edges = [np.full(3, -L / 2)]
    for lev in range(1, n_levels + 1):
        res = (L / ncoarse) / 2 ** lev
        npatch.append(patches_per_level)
        for _ in range(patches_per_level):
            n = 2 * rng.integers(4, max_cells // 2 + 1, 3)              # even: whole parent cells
            jitter = rng.integers(-3, 4, 3) * 2 * res
            left = c - n * res / 2 + jitter
            left = -L / 2 + np.round((left + L / 2) / (2 * res)) * (2 * res)     # parent-cell edges
            nx.append(int(n[0])); ny.append(int(n[1])); nz.append(int(n[2]))
            rx.append(left[0] + res); ry.append(left[1] + res); rz.append(left[2] + res)
            level_of.append(lev)
            lo_edges.append(left)
    npt = len(nx)
    nxa, nya, nza = (np.array(a, dtype=np.int64) for a in (nx, ny, nz))
    hi_edges = [lo_edges[p] + np.array([nx[p], ny[p], nz[p]]) * (L / ncoarse) / 2 ** level_of[p] for p in range(npt)]
    delta, vx, vy, vz, temp, cr0, sol = ([None] * npt for _ in range(7))
    bc = base_cells                                     # the base grid is never read (l = 0 is skipped)
    for p in range(npt):
        shape = (bc, bc, bc) if p == 0 else (nx[p], ny[p], nz[p])
        res = (L / ncoarse) / 2 ** level_of[p]
        ax = [lo_edges[p][k] + (np.arange(shape[k]) + 0.5) * res for k in range(3)]
        X, Y, Z = np.meshgrid(*ax, indexing="ij")
        r = np.sqrt((X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2)
        rc = 20 * KPC
        d = 2e4 / (1 + (r / rc) ** 2) ** 1.5 * rng.lognormal(0, 0.3, shape) - 0.5
        delta[p] = np.asfortranarray(d.astype(np.float32))
        for arr, v0, sig in ((vx, BULK_V[0], 120.0), (vy, BULK_V[1], 120.0), (vz, BULK_V[2], 120.0)):
            arr[p] = np.asfortranarray(((v0 + rng.normal(0, sig, shape)) / 3e5).astype(np.float32))
        temp[p] = np.asfortranarray((10 ** rng.uniform(3, 7, shape)).astype(np.float32))
        refined = np.zeros(shape, dtype=bool)
        overl = np.zeros(shape, dtype=bool)
        for q in range(1, npt):
            if q == p or p == 0 or level_of[q] not in (level_of[p], level_of[p] + 1):
                continue
            inside = ((X > lo_edges[q][0]) & (X < hi_edges[q][0]) & (Y > lo_edges[q][1]) & (Y < hi_edges[q][1])
                      & (Z > lo_edges[q][2]) & (Z < hi_edges[q][2]))
            if level_of[q] == level_of[p] + 1:
                refined |= inside
            elif level_of[q] == level_of[p] and q < p:
                overl |= inside
        flip = rng.uniform(0, 1, shape) < 0.01
        cr0[p] = np.asfortranarray(~refined ^ flip)
        sol[p] = np.asfortranarray(~overl ^ (rng.uniform(0, 1, shape) < 0.01))
    grid_data = [None] * 15
    grid_data[2] = n_levels
    grid_data[5] = np.array(npatch, dtype=np.int64)
    grid_data[6], grid_data[7], grid_data[8] = nxa, nya, nza
    grid_data[12], grid_data[13], grid_data[14] = (np.array(a, dtype=np.float64) for a in (rx, ry, rz))
    gas_data = [delta, vx, vy, vz, temp, cr0, sol]
    dm = dm_cloud(n_dm, 60 * KPC, 8e7 / 8, rng, centre=centre)
    st = plummer_stars(n_st, 8 * KPC, 1e6, rng, centre=centre)
    if background_frac > 0:
        for p_, n_ in ((dm, n_dm), (st, n_st)):
            k = int(background_frac * n_)
            which = rng.choice(n_, k, replace=False)
            p_.x[which], p_.y[which], p_.z[which] = (rng.uniform(-L / 2, L / 2, k) for _ in range(3))
    dm_data = [dm.x, dm.y, dm.z, dm.mass]
    st_data = [st.x, st.y, st.z, st.vx, st.vy, st.vz, st.mass, None, None, np.arange(n_st, dtype=np.int64) + 7_000_000]
    return AmrSnapshot(L, ncoarse, grid_data, gas_data, dm_data, st_data, rho_B=3.9e10, rete=0.8, centre=tuple(c))
