"""GPU parity of the STATELESS MEX path (sph_neighbor_search_mex / sph_physics_shell_mex served by
libsph_b200 through the C ABI) against the CPU oracle and the committed golden vectors.

Bars (BASELINE.json north_star, SURVEY.md 7 gates G1-G3):
  * neighbour lists: BIT-EXACT, all seven columns, in the reference's emission order;
  * every physics mode: max-norm relative <= 1e-11 per output field (TOL below).
"""
import numpy as np
import pytest

from parity_util import PAIR_NAMES, TOL, assert_close, pairs_as_sets

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be(sphb):
    if not sphb.device_available():
        pytest.fail("no CUDA device: the product path has no CPU fallback")
    return sphb.MexBackend()


# ------------------------------------------------------------------------------------------------
# neighbour search
# ------------------------------------------------------------------------------------------------
def test_neighbor_golden_edge_cases_bit_exact(be, golden):
    g = golden("nbr_edge_cases.npz")
    for tag in "abc":
        nf, nt, h, DL = g[f"{tag}_meta"]
        got = be.neighbor_search(g[f"{tag}_pos"], int(nf), int(nt), float(h), float(DL))
        for k, name in enumerate(PAIR_NAMES):
            assert np.array_equal(got[k], g[f"{tag}_{name}"]), f"case {tag}: {name} differs from the reference"


def test_neighbor_golden_ops_state_bit_exact(be, golden):
    g = golden("ops_dp01_dev.npz")
    got = be.neighbor_search(g["in_pos"], int(g["n_fluid"]), int(g["n_total"]), 1.3 * float(g["dp"]), 3.0)
    for k, name in enumerate(PAIR_NAMES):
        assert np.array_equal(got[k], g[name]), name


@pytest.mark.parametrize("dp,DL,seed", [(0.05, 3.0, 3), (0.02, 3.0, 4), (0.01, 3.0, 5), (0.04, 1.0, 6)])
def test_neighbor_matches_oracle_on_developed_states(be, H, oracle_be, dp, DL, seed):
    prm = H.derive_params(dp=dp, DL=DL)
    st = H.developed_state(prm, seed=seed, jitter=0.3)
    want = oracle_be.neighbor_search(st.pos, st.n_fluid, st.n_total, prm.h, prm.DL)
    got = be.neighbor_search(st.pos, st.n_fluid, st.n_total, prm.h, prm.DL)
    assert got[0].size == want[0].size
    assert pairs_as_sets(got, st.n_fluid) == pairs_as_sets(want, st.n_fluid)
    for k, name in enumerate(PAIR_NAMES):
        assert np.array_equal(got[k], want[k]), name


def test_neighbor_lattice_counts(be, H):
    """SURVEY.md appendix B sanity values (lattice initial condition)."""
    for dp, total, fw in ((0.05, 12660, 1320), (0.01, 303300, 6600)):
        prm = H.derive_params(dp=dp)
        st = H.init_particles(prm)
        p = be.neighbor_search(st.pos, st.n_fluid, st.n_total, prm.h, prm.DL)
        assert p[0].size == total
        assert int((p[1] > st.n_fluid).sum()) == fw


def test_neighbor_degenerate_inputs(be, sphb):
    # a single fluid particle and one wall particle far apart: no pairs, empty outputs
    pos = np.asfortranarray(np.array([[0.5, 0.5], [0.5, 5.0]]))
    out = be.neighbor_search(pos, 1, 2, 0.1, 1.0)
    assert all(o.size == 0 for o in out)
    # everything in ONE cell column (DL < 2h): every particle is its own periodic image's neighbour cell
    rng = np.random.default_rng(0)
    pos = np.asfortranarray(rng.random((40, 2)) * np.array([0.15, 1.0]))
    out = be.neighbor_search(pos, 30, 40, 0.1, 0.15)
    assert out[0].size > 0


def test_neighbor_single_column_matches_oracle(be, oracle_be):
    rng = np.random.default_rng(11)
    for DL, h in ((0.15, 0.1), (0.25, 0.1), (0.41, 0.1)):   # n_cell_x = 1, 2, 3
        pos = np.asfortranarray(rng.random((60, 2)) * np.array([DL, 0.8]))
        want = oracle_be.neighbor_search(pos, 45, 60, h, DL)
        got = be.neighbor_search(pos, 45, 60, h, DL)
        for k, name in enumerate(PAIR_NAMES):
            assert np.array_equal(got[k], want[k]), f"DL={DL}: {name}"


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_neighbor_clustered_clouds_match_oracle(be, oracle_be, seed):
    """Dense clusters (hundreds of particles per cell, rows of several hundred neighbours), empty regions,
    particles outside [0, DL) and on both sides of the periodic seam: the tile width is narrowed and the
    list grown on the fly; the result is still the reference's list, bit for bit."""
    rng = np.random.default_rng(seed)
    h, DL = 0.05, 2.0
    parts = [rng.random((300, 2)) * np.array([DL, 1.0])]                                # background
    for cx, cy, n, s in ((0.02, 0.5, 400, 0.03), (1.98, 0.52, 300, 0.04), (1.0, 0.2, 500, 0.02), (0.6, 0.9, 250, 0.05)):
        parts.append(np.array([cx, cy]) + rng.normal(size=(n, 2)) * s)                  # clusters (two at the seam)
    pos = np.concatenate(parts)
    pos[::7, 0] += DL * rng.integers(-2, 3, size=pos[::7, 0].shape)                     # unwrapped x
    rng.shuffle(pos)
    pos = np.asfortranarray(pos)
    nt = pos.shape[0]
    nf = int(0.8 * nt)
    want = oracle_be.neighbor_search(pos, nf, nt, h, DL)
    got = be.neighbor_search(pos, nf, nt, h, DL)
    assert got[0].size == want[0].size and want[0].size > 50000
    for k, name in enumerate(PAIR_NAMES):
        assert np.array_equal(got[k], want[k]), name


@pytest.mark.parametrize("ncols", [4, 5, 6, 7, 9, 12])
def test_neighbor_short_periodic_boxes_match_oracle(be, oracle_be, ncols):
    """Boxes only a few cell columns long: every walker is near the seam (ghost entries, both entries of a
    particle inside one stencil, minimum-image shifts), the general walk does all the work."""
    rng = np.random.default_rng(100 + ncols)
    h = 0.05
    DL = (ncols - 0.37) * 2.0 * h            # a sliver last column
    n = 90 * ncols
    pos = np.asfortranarray(rng.random((n, 2)) * np.array([DL, 0.9]))
    nf = int(0.85 * n)
    want = oracle_be.neighbor_search(pos, nf, n, h, DL)
    got = be.neighbor_search(pos, nf, n, h, DL)
    for k, name in enumerate(PAIR_NAMES):
        assert np.array_equal(got[k], want[k]), f"{ncols} columns: {name}"


def test_neighbor_error_identifiers(be, sphb):
    pos = np.zeros((4, 2), order="F")
    for args, ident in (((pos, 2, 5, 0.1, 1.0), "SPH:Neighbor:count"), ((pos, 5, 4, 0.1, 1.0), "SPH:Neighbor:count"),
                        ((pos, 2, 4, 0.0, 1.0), "SPH:Neighbor:param"), ((pos, 2, 4, 0.1, -1.0), "SPH:Neighbor:param")):
        with pytest.raises(sphb.SphError) as e:
            be.neighbor_search(*args)
        assert e.value.identifier == ident


# ------------------------------------------------------------------------------------------------
# physics modes
# ------------------------------------------------------------------------------------------------
def test_all_modes_against_golden(be, H, golden):
    """Every mode once on the committed dp=0.1 developed state (reference outputs, 1 thread)."""
    from oracle.gen_golden import ops_case
    g = golden("ops_dp01_dev.npz")
    prm = H.derive_params(dp=float(g["dp"]), c_f=float(g["c_f"]))
    st = H.init_particles(prm)
    st.pos = np.asfortranarray(g["in_pos"])
    st.vel = np.asfortranarray(g["in_vel"])
    st.drho_dt = g["in_drho_dt"].copy()
    got = ops_case(be, prm, st)
    for name in PAIR_NAMES:
        assert np.array_equal(got[name], g[name]), name
    for k, v in got.items():
        if k in PAIR_NAMES:
            continue
        assert_close(np.asarray(v), g[k], k)


@pytest.mark.parametrize("dp,seed", [(0.05, 21), (0.02, 22)])
def test_all_modes_against_oracle(be, H, oracle_be, dp, seed):
    from oracle.gen_golden import ops_case
    prm = H.derive_params(dp=dp, c_f=10.0)
    st = H.developed_state(prm, seed=seed)
    want = ops_case(oracle_be, prm, st.copy())
    got = ops_case(be, prm, st.copy())
    for k in want:
        if k in PAIR_NAMES:
            assert np.array_equal(got[k], want[k]), k
        else:
            assert_close(np.asarray(got[k]), np.asarray(want[k]), k)


def test_modes_skip_bad_pairs_like_reference(be, H, oracle_be):
    """Out-of-range indices and r <= 1e-12 pairs are silently skipped."""
    prm = H.derive_params(dp=0.1, c_f=10.0)
    st = H.developed_state(prm, seed=5)
    nf, nt = st.n_fluid, st.n_total
    pairs = [p.copy() for p in oracle_be.neighbor_search(st.pos, nf, nt, prm.h, prm.DL)]
    pairs[0][3] = 0.0            # i out of range (0 -> -1)
    pairs[1][7] = nt + 5.0       # j out of range
    pairs[0][9] = nf + 1.0       # i is a wall row: not a fluid index
    pairs[4][11] = 0.0           # r <= 1e-12
    pi, pj, dx, dy, r, W, dW = pairs
    args = (pi, pj, dx, dy, r, W, dW, st.mass, nf, nt, prm.rho0, prm.h, prm.inv_sigma0)
    want = oracle_be.physics("density_correction", *args, nargout=3)
    got = be.physics("density_correction", *args, nargout=3)
    for a, b, name in zip(got, want, ("rho", "Vol", "B")):
        assert_close(a, b, name)
    rho, Vol, B = want
    args = (pi, pj, dx, dy, r, dW, st.vel, Vol, B, prm.mu, prm.h, nf, nt, st.mass, st.wall_vel)
    assert_close(be.physics("viscous_force", *args, nargout=1)[0],
                 oracle_be.physics("viscous_force", *args, nargout=1)[0], "force")


def test_physics_error_identifiers(be, sphb, H, oracle_be):
    prm = H.derive_params(dp=0.1)
    st = H.init_particles(prm)
    nf, nt = st.n_fluid, st.n_total
    pi, pj, dx, dy, r, W, dW = oracle_be.neighbor_search(st.pos, nf, nt, prm.h, prm.DL)
    cases = [
        ("density_correction", (pi, pj[:-1], dx, dy, r, W, dW, st.mass, nf, nt, prm.rho0, prm.h, prm.inv_sigma0), 3,
         "SPH:Physics:density:pairs"),
        ("density_correction", (pi, pj, dx, dy, r, W, dW, st.mass[:-1], nf, nt, prm.rho0, prm.h, prm.inv_sigma0), 3,
         "SPH:Physics:density:mass"),
        ("density_correction", (pi, pj, dx, dy, r, W, dW, st.mass, nf, nt, -1.0, prm.h, prm.inv_sigma0), 3,
         "SPH:Physics:density:param"),
        ("viscous_force", (pi, pj, dx, dy, r, dW, st.vel[:-1], st.Vol, st.B, prm.mu, prm.h, nf, nt, st.mass,
                           st.wall_vel), 1, "SPH:Physics:viscous:vel"),
        ("viscous_force", (pi, pj, dx, dy, r, dW, st.vel, st.Vol, st.B, prm.mu, prm.h, nf, nt, st.mass,
                           st.wall_vel[:-1]), 1, "SPH:Physics:viscous:wallvel"),
        ("transport_correction", (pi, pj, dx, dy, r, dW, st.Vol, st.B, st.pos, prm.h, nf, nt, -0.5), 1,
         "SPH:Physics:transport:coeff"),
        ("integration_2nd", (pi, pj, dx, dy, r, dW, st.Vol, st.rho[:-1], st.pos, st.vel, 1e-3, nf, nt, st.wall_vel), 3,
         "SPH:Physics:int2:rho"),
    ]
    for mode, args, nout, ident in cases:
        with pytest.raises(sphb.SphError) as e:
            be.physics(mode, *args, nargout=nout)
        assert e.value.identifier == ident, (mode, e.value.identifier)
        # and the reference raises the same identifier
        from oracle.mexhost import MexError
        if oracle_be.kind == "reference":
            with pytest.raises(MexError) as e2:
                oracle_be.physics(mode, *args, nargout=nout)
            assert e2.value.identifier == ident


# ------------------------------------------------------------------------------------------------
# the MATLAB main loop driven through the GPU MEX path, against the golden trajectory
# ------------------------------------------------------------------------------------------------
def test_reference_loop_through_gpu_mex(be, H, golden):
    """SPH_Poiseuille.m's loop body with the GPU serving both MEX files: 120 steps from the dp=0.05
    lattice (cell sort at step 100).  dt and pair counts equal the reference's exactly; state fields
    within 1e-9 at step 10 and 1e-7 at step 120 (EOS amplification of ulp-level density
    differences, then chaotic growth: a report-style gate)."""
    g = golden("traj_dp005.npz")
    prm = H.derive_params(dp=0.05, c_f=15.0, t_end=20.0, output_interval=1.0)
    st = H.init_particles(prm)
    H.initialise(be, st, prm)
    assert st.pairs[0].size == int(g["n_pairs0"])
    for k in range(1, 121):
        dt = H.reference_step(be, st, prm)
        assert abs(dt - g["dt"][k - 1]) <= 1e-13 * g["dt"][k - 1]
        assert st.pairs[0].size == g["n_pairs"][k - 1], f"pair count at step {k}"
        if k in (10, 120):
            tol = 1e-9 if k == 10 else 1e-7   # EOS amplification + chaotic growth (gate G4)
            for f in ("pos", "vel", "rho", "p", "drho_dt", "force", "force_prior", "Vol", "B"):
                assert_close(getattr(st, f), g[f"s{k}_{f}"], f"step {k} {f}", tol=tol)
    assert_close(np.array([st.tau_bottom, st.tau_top]), g["tau"][119], "tau", tol=1e-7)


# ------------------------------------------------------------------------------------------------
# the compiled MEX shims, driven by the fake MATLAB host exactly like the reference's MEX files
# ------------------------------------------------------------------------------------------------
class _ShimBackend:
    kind = "b200-mex-shim"

    def __init__(self):
        import os
        from conftest import ROOT
        from oracle.mexhost import MexLibrary
        # (SPHB_CPUSIM_MEXHOST: test infrastructure, the shims linked against the simulator build of the library)
        self.mex = MexLibrary(os.environ.get("SPHB_CPUSIM_MEXHOST") or
                              os.path.join(ROOT, "sph-poiseuille-flow_b200", "libsphb_mex_stubhost.so"))

    def neighbor_search(self, pos, n_fluid, n_total, h, DL):
        out = self.mex.call("sphb_neighbor_mexFunction", 7, pos, n_fluid, n_total, h, DL)
        return tuple(o.ravel(order="F") for o in out)

    def physics(self, mode, *args, nargout):
        out = self.mex.call("sphb_physics_mexFunction", nargout, mode, *args)
        return [o if o.shape[1] > 1 else o.ravel(order="F") for o in out]


def test_mex_shims_all_modes_against_golden(be, H, golden):
    """feval(mex, ...) as SPH_Poiseuille.m issues it -> MEX shim (C) -> C ABI -> sm_100a kernels."""
    from oracle.gen_golden import ops_case
    shim = _ShimBackend()
    g = golden("ops_dp01_dev.npz")
    prm = H.derive_params(dp=float(g["dp"]), c_f=float(g["c_f"]))
    st = H.init_particles(prm)
    st.pos = np.asfortranarray(g["in_pos"])
    st.vel = np.asfortranarray(g["in_vel"])
    st.drho_dt = g["in_drho_dt"].copy()
    got = ops_case(shim, prm, st)
    for k, v in got.items():
        if k in PAIR_NAMES:
            assert np.array_equal(v, g[k]), k
        else:
            assert_close(np.asarray(v), g[k], k)


def test_mex_shims_error_identifiers_match_reference(be, H, oracle_be):
    from oracle.mexhost import MexError
    if oracle_be.kind != "reference":
        pytest.skip("needs oracle/_ref")
    shim = _ShimBackend()
    prm = H.derive_params(dp=0.1)
    st = H.init_particles(prm)
    nf, nt = st.n_fluid, st.n_total
    pi, pj, dx, dy, r, W, dW = oracle_be.neighbor_search(st.pos, nf, nt, prm.h, prm.DL)
    cases = [
        ("viscous_force", (pi, pj, dx, dy, r, dW, st.vel, st.Vol[:-1], st.B, prm.mu, prm.h, nf, nt, st.mass, st.wall_vel), 1),
        ("viscous_force", (pi, pj, dx, dy, r, dW, st.vel, st.Vol, st.B[:, :3], prm.mu, prm.h, nf, nt, st.mass, st.wall_vel), 1),
        ("transport_correction", (pi, pj, dx, dy, r, dW, st.Vol, st.B, st.pos[:-1], prm.h, nf, nt), 1),
        ("integration_verlet", (pi, pj, dx, dy, r, dW, st.Vol, st.B, st.rho, st.mass, st.pos, st.vel, st.drho_dt[:-1],
                                st.force_prior, 1e-3, nf, nt, prm.rho0, prm.p0, prm.c_f, st.wall_vel), 6),
        ("integration_1st", (pi, pj, dx, dy, r, dW, st.Vol, st.B, st.rho, st.mass, st.pos, st.vel, st.drho_dt,
                             st.force_prior[:-1], 1e-3, nf, nt, prm.rho0, prm.p0, prm.c_f, st.wall_vel), 5),
        ("wall_shear_monitor", (pi, pj, dx, dy, r, dW, st.pos, st.vel[:-1], st.wall_vel, st.Vol, st.B, nf, prm.DL, prm.DH,
                                prm.mu, prm.h), 2),
        ("advance_shell_step", (pi, pj, dx, dy, r, W, dW, st.mass, st.pos, st.vel, st.wall_vel, st.rho[:-1], st.drho_dt,
                                1e-3, nf, nt, prm.rho0, prm.p0, prm.c_f, prm.mu, prm.h, prm.inv_sigma0, prm.gravity_g), 9),
        ("density_correction", (pi, pj, dx, dy, r, W, dW, st.mass, 0, nt, prm.rho0, prm.h, prm.inv_sigma0), 3),
    ]
    for mode, args, nout in cases:
        with pytest.raises(MexError) as want:
            oracle_be.physics(mode, *args, nargout=nout)
        with pytest.raises(MexError) as got:
            shim.physics(mode, *args, nargout=nout)
        assert got.value.identifier == want.value.identifier, (mode, got.value.identifier, want.value.identifier)
