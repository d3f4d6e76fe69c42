"""Host-side mirror of the set-up slice of SPH_Poiseuille.m (S2/S3): derived parameters, the lattice
initial condition and a seeded synthetic "developed" state for benchmarks.

Reference: SPH_Poiseuille.m:62-80 (geometry snapping, g, h, wall thickness, p0, inv_sigma0, nu),
:95-123 (fluid lattice, masses, B = I) and build_shell_wall_particles.m:14-38 (wall blocks).  Pure
host bookkeeping (numpy); every particle-loop operation lives behind the C ABI.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


def _round_half_away(v: float) -> float:   # MATLAB round()
    return math.copysign(math.floor(abs(v) + 0.5), v)


@dataclass
class ChannelConfig:
    """[domain]/[physics]/[simulation] keys of config.ini plus the script's hard-coded constants."""
    DL: float = 3.0
    DH: float = 1.0
    dp: float = 0.05
    rho0: float = 1.0
    mu: float = 0.1
    U_bulk: float = 0.666667
    c_f: float = 15.0
    end_time: float = 20.0
    output_interval: float = 1.0
    sort_interval: int = 100
    transport_coeff: float = 0.30      # m:77
    h_over_dp: float = 1.3             # m:73
    wall_layers_min: float = 4.0       # m:74

    def derive(self) -> "ChannelConfig":
        self.DL = _round_half_away(self.DL / self.dp) * self.dp          # m:64
        self.DH = _round_half_away(self.DH / self.dp) * self.dp          # m:65
        self.gravity_g = 12.0 * self.mu * self.U_bulk / (self.rho0 * self.DH ** 2)   # m:71
        self.U_max = 1.5 * self.U_bulk                                    # m:72
        self.h = self.h_over_dp * self.dp                                 # m:73
        self.wall_thickness = max(self.wall_layers_min * self.dp, math.ceil(2.0 * self.h / self.dp) * self.dp)
        self.p0 = self.rho0 * self.c_f ** 2                               # m:78
        self.inv_sigma0 = self.dp ** 2                                    # m:79
        self.nu = self.mu / self.rho0                                     # m:80
        return self


@dataclass
class Particles:
    pos: np.ndarray       # [n_total, 2] column-major
    vel: np.ndarray
    rho: np.ndarray
    drho_dt: np.ndarray
    mass: np.ndarray
    wall_vel: np.ndarray
    n_fluid: int
    n_total: int


def lattice(cfg: ChannelConfig) -> Particles:
    """Fluid rows first (x-major, y fastest: meshgrid(x, y)(:), m:97-98), then the bottom wall block,
    then the top wall block (build_shell_wall_particles.m:22-31)."""
    dp = cfg.dp
    nx = int(_round_half_away(cfg.DL / dp))
    ny = int(_round_half_away(cfg.DH / dp))
    nl = int(_round_half_away(cfg.wall_thickness / dp))
    xs = dp / 2 + dp * np.arange(nx, dtype=np.float64)
    cols = []
    for y0, n in ((dp / 2, ny), (-cfg.wall_thickness + dp / 2, nl), (cfg.DH + dp / 2, nl)):
        ys = y0 + dp * np.arange(n, dtype=np.float64)
        cols.append(np.stack([np.repeat(xs, n), np.tile(ys, nx)], axis=1))
    pos = np.asfortranarray(np.concatenate(cols, axis=0))
    n_fluid, n_total = nx * ny, pos.shape[0]
    zeros2 = lambda: np.zeros((n_total, 2), order="F")   # noqa: E731
    return Particles(pos=pos, vel=zeros2(), rho=np.full(n_total, cfg.rho0), drho_dt=np.zeros(n_total),
                     mass=np.full(n_total, cfg.rho0 * dp * dp), wall_vel=zeros2(), n_fluid=n_fluid, n_total=n_total)


def developed(cfg: ChannelConfig, seed: int = 12345, jitter: float = 0.1, noise: float = 0.01) -> Particles:
    """Synthetic developed flow: the lattice with a seeded uniform +-jitter*dp displacement of the
    fluid rows, the analytic Poiseuille profile plus noise*U_max velocity noise, small random
    drho_dt (numpy PCG64, deterministic per seed)."""
    pt = lattice(cfg)
    rng = np.random.Generator(np.random.PCG64(seed))
    nf = pt.n_fluid
    pt.pos[:nf, :] += (rng.random((nf, 2)) * 2.0 - 1.0) * jitter * cfg.dp
    y = pt.pos[:nf, 1]
    u = cfg.gravity_g / (2.0 * cfg.nu) * y * (cfg.DH - y)
    pt.vel[:nf, 0] = u + noise * cfg.U_max * (rng.random(nf) * 2.0 - 1.0)
    pt.vel[:nf, 1] = noise * cfg.U_max * (rng.random(nf) * 2.0 - 1.0)
    pt.drho_dt[:nf] = 0.05 * cfg.rho0 * (rng.random(nf) * 2.0 - 1.0)
    pt.pos[:nf, 0] = np.mod(pt.pos[:nf, 0], cfg.DL)
    return pt


def cell_sorted(cfg: ChannelConfig, pt: Particles) -> Particles:
    """The fluid rows in the order the reference's periodic cell sort gives them (sort_particles_by_cell,
    SPH_Poiseuille.m:529-568: key (cy, cx), cell = 2h, cy from min(fluid y), stable; walls stay last) -- the
    order a running simulation is in after its first sort point (m:272)."""
    nf = pt.n_fluid
    cell = 2.0 * cfg.h
    x = np.mod(pt.pos[:nf, 0], cfg.DL)
    cx = np.floor(x / cell).astype(np.int64)
    cy = np.floor((pt.pos[:nf, 1] - pt.pos[:nf, 1].min()) / cell).astype(np.int64)
    perm = np.concatenate([np.lexsort((cx, cy)), np.arange(nf, pt.n_total)])     # stable, like sortrows
    take2 = lambda a: np.asfortranarray(a[perm, :])   # noqa: E731
    return Particles(pos=take2(pt.pos), vel=take2(pt.vel), rho=pt.rho[perm].copy(), drho_dt=pt.drho_dt[perm].copy(),
                     mass=pt.mass[perm].copy(), wall_vel=take2(pt.wall_vel), n_fluid=nf, n_total=pt.n_total)


def analytic_profile(cfg: ChannelConfig, y):
    """u(y) = g / (2 nu) y (DH - y)  (SPH_Poiseuille_postprocess.m:36-38)."""
    y = np.asarray(y, dtype=np.float64)
    return cfg.gravity_g / (2.0 * cfg.nu) * y * (cfg.DH - y)
