"""sphb_ctx_step_uploaded (one reference step for a caller whose state lives on the host: list for the uploaded
positions + shear monitor + physics, no trailing list rebuild) against the oracle and against sphb_ctx_step.

Written after the GPU budget of round 1 was spent (the entry point composes calls the other GPU tests cover);
the file sorts late so that a surprise here cannot mask the rest of the `-m gpu` run.
"""
import numpy as np
import pytest

from parity_util import assert_close

pytestmark = pytest.mark.gpu

FIELDS = ("pos", "vel", "rho", "p", "drho_dt", "force", "force_prior", "Vol", "B")


def _channel(sphb, prm, st):
    ch = sphb.channel_from_oracle_params(prm, st.n_fluid, st.n_total, sort_interval=0)
    ch.upload(pos=st.pos, vel=st.vel, rho=st.rho, drho_dt=st.drho_dt, mass=st.mass, wall_vel=st.wall_vel)
    return ch


def test_step_uploaded_is_the_physics_of_a_step(sphb, H, oracle_be):
    """Three passes of upload -> step_uploaded -> download (state through host arrays every time) give the fields
    of three sphb_ctx_step calls on a resident context to the bit; the first pass is also held against the oracle."""
    prm = H.derive_params(dp=0.05, c_f=15.0)
    st = H.developed_state(prm, seed=77)
    ref = st.copy()
    H.initialise(oracle_be, ref, prm)

    a = _channel(sphb, prm, st)          # resident: reference steps on the device
    a.initialise()
    b = _channel(sphb, prm, st)          # host-state chain
    b.initialise()
    for k in range(3):
        a.step()
        d = b.download(("pos", "vel", "rho", "drho_dt"))
        b.upload(pos=d["pos"], vel=d["vel"], rho=d["rho"], drho_dt=d["drho_dt"])
        b.step_uploaded()
        if k == 0:                       # one step against the oracle, at the one-step tolerance
            H.reference_step(oracle_be, ref, prm)
            d1 = b.download()
            for f in FIELDS:
                assert_close(d1[f], getattr(ref, f), f)
    da, db = a.download(), b.download()
    for f in FIELDS:
        assert np.array_equal(da[f], db[f]), f
    sa, sb = a.status(), b.status()
    assert sa["t"] == sb["t"] and sa["dt"] == sb["dt"] and sb["error_flags"] == 0
    # the monitor of the chain runs at the start of the next pass, on the list of the uploaded positions and with
    # the volumes and B of the last step: the shear the resident run reported at the end of its step
    d = b.download(("pos", "vel", "rho", "drho_dt"))
    b.upload(pos=d["pos"], vel=d["vel"], rho=d["rho"], drho_dt=d["drho_dt"])
    b.step_uploaded()
    sb = b.status()
    assert sb["tau_bottom"] == sa["tau_bottom"] and sb["tau_top"] == sa["tau_top"]
    a.close()
    b.close()


def test_step_uploaded_keeps_a_valid_list(sphb, H):
    """With a list already built for the resident positions the call does not build another one (same launch
    count as shear + physics), and it refuses to run before any upload."""
    prm = H.derive_params(dp=0.05, c_f=15.0)
    st = H.developed_state(prm, seed=78)
    ch = _channel(sphb, prm, st)
    ch.initialise()
    l0 = ch.timing()["launches"]
    ch.step_uploaded()
    ch.sync()
    with_list = ch.timing()["launches"] - l0
    d = ch.download(("pos",))
    ch.upload(pos=d["pos"])               # invalidates the list
    l1 = ch.timing()["launches"]
    ch.step_uploaded()
    ch.sync()
    assert ch.timing()["launches"] - l1 > with_list
    assert ch.status()["error_flags"] == 0
    ch.close()


def test_step_host_equals_the_three_calls_bitwise(sphb, H):
    """sphb_ctx_step_host (upload -> list -> shear -> physics -> download in ONE call, the list build overlapping the
    rest of the upload on a second stream) gives the arrays of upload + step_uploaded + download to the bit, over
    several passes with the state going through host arrays every time, and with a remaining-time cut."""
    prm = H.derive_params(dp=0.05, c_f=15.0)
    st = H.developed_state(prm, seed=79)
    a = _channel(sphb, prm, st)
    a.initialise()
    b = _channel(sphb, prm, st)
    b.initialise()
    nt = st.n_total
    sa = a.download(("pos", "vel", "rho", "drho_dt"))
    sb = {k: v.copy(order="F") for k, v in sa.items()}
    names = ("pos", "vel", "rho", "p", "drho_dt", "force")
    for k in range(4):
        remain = float("inf") if k != 2 else 0.5 * a.status()["dt"]
        a.upload(pos=sa["pos"], vel=sa["vel"], rho=sa["rho"], drho_dt=sa["drho_dt"])
        a.step_uploaded(remain)
        da = a.download(names)
        out = {f: (np.zeros((nt, 2), order="F") if f in ("pos", "vel", "force") else np.zeros(nt)) for f in names}
        b.step_host(sb, out, remain)
        for f in names:
            assert np.array_equal(da[f], out[f]), (k, f)
        ta, tb = a.status(), b.status()
        # (in the pass after the wall moved, the three-call route recovers from the outgrown list the way a resident
        # run does -- list and monitor again AFTER the physics -- so its monitor and pair count are those of the new
        # positions; the one-call route repeats the step and reports the monitor of the uploaded ones)
        keys = ("t", "dt", "vmax", "step", "error_flags") + (() if k == 2 else ("tau_bottom", "tau_top", "n_half_pairs"))
        for key in keys:
            assert ta[key] == tb[key], (k, key)
        sa = {f: da[f] for f in ("pos", "vel", "rho", "drho_dt")}
        sb = {f: out[f] for f in ("pos", "vel", "rho", "drho_dt")}
        if k == 1:      # rows leave the y range the lists were planned for (the top wall moves three channel heights
            # up: more cell rows than the plan's slack): sphb_ctx_upload plans again from the host array, the one-call
            # route learns it from the device's capacity flag and repeats the step -- same arrays either way
            for d in (sa, sb):
                d["pos"] = d["pos"].copy(order="F")
                top = d["pos"][:, 1] > prm.DH
                top[:st.n_fluid] = False
                d["pos"][top, 1] += 3.0
    a.close()
    b.close()
