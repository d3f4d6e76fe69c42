"""Pins the CPU oracle.

The reference ships no golden vectors, so the pin is the reference itself: its MEX sources compiled
unmodified (oracle/_ref) produced tests/golden/*.npz (oracle/gen_golden.py, OMP_NUM_THREADS=1).  The
plain-C port (oracle/sph_oracle.c) must reproduce them BIT FOR BIT, and so must the reference build
when it is present.  All CPU-only.
"""
import numpy as np
import pytest

PAIR_NAMES = ("pair_i", "pair_j", "dx", "dy", "r", "W", "dW")


def _prm_state(H, g):
    prm = H.derive_params(dp=float(g["dp"]), c_f=float(g["c_f"]))
    st = H.init_particles(prm)
    st.pos = np.asfortranarray(g["in_pos"])
    st.vel = np.asfortranarray(g["in_vel"])
    st.drho_dt = g["in_drho_dt"].copy()
    return prm, st


def _backends(H, port, request):
    out = [port]
    import os
    if os.path.exists(H.REF_SO):
        out.append(H.RefBackend())
    return out


def test_lattice_pair_counts(H, port):
    """SURVEY.md appendix B sanity values."""
    for dp, total, fw in ((0.05, 12660, 1320), (0.04, 19575, 1650)):
        prm = H.derive_params(dp=dp)
        st = H.init_particles(prm)
        p = port.neighbor_search(st.pos, st.n_fluid, st.n_total, prm.h, prm.DL)
        assert p[0].size == total
        assert int((p[1] > st.n_fluid).sum()) == fw
        # fluid-fluid pairs are stored once with i < j; fluid-wall with i = fluid
        ff = p[1] <= st.n_fluid
        assert np.all(p[0][ff] < p[1][ff]) and np.all(p[0] <= st.n_fluid)


def test_ops_golden_bit_exact(H, port, golden, request):
    from oracle.gen_golden import ops_case
    g = golden("ops_dp01_dev.npz")
    for be in _backends(H, port, request):
        prm, st = _prm_state(H, g)
        got = ops_case(be, prm, st)
        for k, v in got.items():
            assert np.array_equal(np.asarray(v), g[k]), f"{be.kind}: {k} differs from the golden vector"


def test_neighbor_edge_cases_bit_exact(H, port, golden, request):
    g = golden("nbr_edge_cases.npz")
    for be in _backends(H, port, request):
        for tag in "abc":
            nf, nt, h, DL = g[f"{tag}_meta"]
            got = be.neighbor_search(g[f"{tag}_pos"], int(nf), int(nt), float(h), float(DL))
            for k, name in enumerate(PAIR_NAMES):
                assert np.array_equal(got[k], g[f"{tag}_{name}"]), f"{be.kind}: case {tag} {name}"


def test_trajectory_golden_bit_exact(H, port, golden, request):
    """120 reference steps from the dp=0.05 lattice, crossing the cell sort at step 100."""
    g = golden("traj_dp005.npz")
    for be in _backends(H, port, request):
        prm = H.derive_params(dp=0.05, c_f=15.0, t_end=20.0, output_interval=1.0)
        st = H.init_particles(prm)
        H.initialise(be, st, prm)
        assert st.pairs[0].size == int(g["n_pairs0"])
        for k in range(1, 121):
            dt = H.reference_step(be, st, prm)
            assert dt == g["dt"][k - 1]
            assert st.pairs[0].size == g["n_pairs"][k - 1]
            assert (st.tau_bottom, st.tau_top) == tuple(g["tau"][k - 1])
            if k in (10, 120):
                for f in ("pos", "vel", "rho", "p", "drho_dt", "force", "force_prior", "Vol", "B"):
                    assert np.array_equal(getattr(st, f), g[f"s{k}_{f}"]), f"{be.kind}: step {k} {f}"


def test_port_matches_reference_on_random_states(H, port, ref):
    """Port == reference build, bitwise, on seeded developed states at two resolutions."""
    from oracle.gen_golden import ops_case
    for dp, seed in ((0.05, 1), (0.04, 2)):
        prm = H.derive_params(dp=dp, c_f=10.0)
        st = H.developed_state(prm, seed=seed)
        a = ops_case(port, prm, st.copy())
        b = ops_case(ref, prm, st.copy())
        for k in a:
            assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k


def test_error_identifiers_of_reference(ref):
    """The boundary's machine-checkable contract, as raised by the reference."""
    from oracle.mexhost import MexError
    pos = np.zeros((4, 2), order="F")
    with pytest.raises(MexError) as e:
        ref.mex.call("ref_neighbor_mexFunction", 7, pos, 2, 4, 0.1)
    assert e.value.identifier == "SPH:Neighbor:nrhs"
    with pytest.raises(MexError) as e:
        ref.mex.call("ref_neighbor_mexFunction", 7, pos, 2, 5, 0.1, 1.0)
    assert e.value.identifier == "SPH:Neighbor:count"
    with pytest.raises(MexError) as e:
        ref.mex.call("ref_neighbor_mexFunction", 7, pos, 2, 4, -0.1, 1.0)
    assert e.value.identifier == "SPH:Neighbor:param"
    with pytest.raises(MexError) as e:
        ref.mex.call("ref_physics_mexFunction", 1, "no_such_mode")
    assert e.value.identifier == "SPH:Physics:mode"


def test_profile_helpers(H):
    prm = H.derive_params(dp=0.05)
    st = H.init_particles(prm)
    nf = st.n_fluid
    y = st.pos[:nf, 1]
    u = prm.gravity_g / (2 * prm.nu) * y * (prm.DH - y)
    assert H.l2_error(prm, st.pos[:nf], u) < 1e-12   # lattice rows sit on bin centres
    ym, um = H.mid_channel_profile(st.pos[:nf], u, prm.DL, prm.DH, 0.5 * prm.DL, max(prm.dp, prm.h), 20)
    assert ym.shape == (20,) and not np.isnan(um).any()


def test_port_matches_reference_on_fuzz_clouds(port, ref):
    """The clouds of tests/test_gpu_zx_neighbor_fuzz.py (few-column boxes, points on cell boundaries and on the seam,
    coincident points, pairs at exactly the cut-off, unwrapped x): the plain-C port reproduces the reference build's
    seven pair columns bit for bit, so either may serve as the checker of that GPU test."""
    import test_gpu_zx_neighbor_fuzz as F
    for block in range(6):
        for label, pos, nf, n, h, DL in F.cases(block):
            a, b = ref.neighbor_search(pos, nf, n, h, DL), port.neighbor_search(pos, nf, n, h, DL)
            assert all(np.array_equal(x, y) for x, y in zip(a, b)), label
