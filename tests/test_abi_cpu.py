"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/sph_b200.h declares, keeps the reference's argument-count / error-identifier contract on the
host side, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_product


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sph_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"SPHB_API\s+[^;{]*?\b(sphb_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    sphb = load_product()
    L = sphb.lib()
    names = _declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(L, n), f"libsph_b200.so does not export {n}"
    assert L.sphb_version() == 100


def test_no_cpu_fallback_without_device():
    sphb = load_product()
    if sphb.device_available():
        pytest.skip("a CUDA device is present")
    pos = np.zeros((4, 2), order="F")
    with pytest.raises(sphb.SphError) as e:
        sphb.sph_neighbor_search_mex(pos, 2, 4, 0.1, 1.0)
    assert e.value.identifier == "SPHB:cuda"
    with pytest.raises(sphb.SphError) as e:
        sphb.Channel(sphb.ChannelParams(n_fluid=2, n_total=4, DL=1.0, DH=1.0, dp=0.1, h=0.13, rho0=1.0, p0=100.0,
                                        c_f=10.0, mu=0.1, gravity_g=0.8, inv_sigma0=0.01))
    assert e.value.identifier == "SPHB:cuda"


def test_gateway_argument_count_identifiers():
    """nrhs/nlhs identifiers of the two gateways (nb:216-221, ph:1749-1770 and each mode's header)."""
    sphb = load_product()
    with pytest.raises(sphb.SphError) as e:
        sphb.sph_neighbor_search_mex(np.zeros((4, 2)), 2, 4, 0.1)
    assert e.value.identifier == "SPH:Neighbor:nrhs"
    with pytest.raises(sphb.SphError) as e:
        sphb.sph_neighbor_search_mex(np.zeros((4, 2)), 2, 4, 0.1, 1.0, nargout=3)
    assert e.value.identifier == "SPH:Neighbor:nlhs"
    with pytest.raises(sphb.SphError) as e:
        sphb.sph_physics_shell_mex("no_such_mode", nargout=1)
    assert e.value.identifier == "SPH:Physics:mode"
    with pytest.raises(sphb.SphError) as e:
        sphb.sph_physics_shell_mex(3.0, nargout=1)
    assert e.value.identifier == "SPH:Physics:mode"
    for mode, tag, nout in (("density_correction", "density", 3), ("viscous_force", "viscous", 1),
                            ("transport_correction", "transport", 1), ("integration_1st", "int1", 5),
                            ("integration_2nd", "int2", 3), ("integration_verlet", "verlet", 6),
                            ("advance_shell_step", "advance", 9), ("wall_shear_monitor", "wallshear", 2)):
        with pytest.raises(sphb.SphError) as e:
            sphb.sph_physics_shell_mex(mode, 1.0, 2.0, nargout=nout)
        assert e.value.identifier == f"SPH:Physics:{tag}:nrhs"


def test_stateless_validation_identifiers_precede_device_use():
    """Size checks that the reference performs before touching data are reproduced inside the library
    and fire before any CUDA call (so they are testable without a GPU)."""
    sphb = load_product()
    pos = np.zeros((4, 2), order="F")
    with pytest.raises(sphb.SphError) as e:
        sphb.sph_neighbor_search_mex(pos, 2, 5, 0.1, 1.0)          # n_total != rows(pos)
    assert e.value.identifier == "SPH:Neighbor:count"
    with pytest.raises(sphb.SphError) as e:
        sphb.sph_neighbor_search_mex(pos, 2, 4, -0.1, 1.0)
    assert e.value.identifier == "SPH:Neighbor:param"
    with pytest.raises(sphb.SphError) as e:
        sphb.sph_neighbor_search_mex(np.zeros((4, 3), order="F"), 2, 4, 0.1, 1.0)
    assert e.value.identifier == "SPH:Neighbor:pos"


def test_struct_layouts_match_header():
    """ctypes mirrors of sphb_params / sphb_status / sphb_timing have the C sizes (x86-64 SysV)."""
    sphb = load_product()
    assert ctypes.sizeof(sphb._Params) == 2 * 4 + 11 * 8 + 3 * 4 + 4   # ints, doubles, ints + tail pad
    assert ctypes.sizeof(sphb._Status) == 5 * 8 + 3 * 8 + 8
    assert ctypes.sizeof(sphb._Mat) == 24


def test_mex_shims_load_and_keep_gateway_contract():
    """The two MEX gateways (sph-poiseuille-flow_b200/mex/*.c) built against the stub MATLAB host:
    symbols present, argument-count errors raised with the reference's identifiers (no GPU needed)."""
    from oracle.mexhost import MexError, MexLibrary
    path = os.path.join(ROOT, "sph-poiseuille-flow_b200", "libsphb_mex_stubhost.so")
    if not os.path.exists(path):
        import __graft_entry__ as ge
        ge.build()
    mex = MexLibrary(path)
    assert mex.has("sphb_neighbor_mexFunction") and mex.has("sphb_physics_mexFunction")
    pos = np.zeros((4, 2), order="F")
    with pytest.raises(MexError) as e:
        mex.call("sphb_neighbor_mexFunction", 7, pos, 2, 4, 0.1)
    assert e.value.identifier == "SPH:Neighbor:nrhs"
    with pytest.raises(MexError) as e:
        mex.call("sphb_neighbor_mexFunction", 6, pos, 2, 4, 0.1, 1.0)
    assert e.value.identifier == "SPH:Neighbor:nlhs"
    with pytest.raises(MexError) as e:
        mex.call("sphb_neighbor_mexFunction", 7, np.zeros((4, 3), order="F"), 2, 4, 0.1, 1.0)
    assert e.value.identifier == "SPH:Neighbor:pos"
    with pytest.raises(MexError) as e:
        mex.call("sphb_neighbor_mexFunction", 7, pos, 2, 5, 0.1, 1.0)
    assert e.value.identifier == "SPH:Neighbor:count"
    with pytest.raises(MexError) as e:
        mex.call("sphb_physics_mexFunction", 1, "no_such_mode")
    assert e.value.identifier == "SPH:Physics:mode"
    with pytest.raises(MexError) as e:
        mex.call("sphb_physics_mexFunction", 3, "density_correction", 1.0)
    assert e.value.identifier == "SPH:Physics:density:nrhs"
    with pytest.raises(MexError) as e:
        mex.call("sphb_physics_mexFunction", 2, "integration_verlet", *([1.0] * 21))
    assert e.value.identifier == "SPH:Physics:verlet:nlhs"


def test_batch_entry_points_validate_arguments_without_device():
    """sphb_batch_* (batched ensemble launches): argument checks fire before any CUDA call, NULL handles are
    tolerated by the accessors and by destroy."""
    sphb = load_product()
    L = sphb.lib()
    out = ctypes.c_void_p()
    assert L.sphb_batch_create(None, ctypes.c_int(0), ctypes.byref(out)) != 0
    assert L.sphb_last_error_id() == b"SPHB:arg"
    arr = (ctypes.c_void_p * 1)(None)
    assert L.sphb_batch_create(arr, ctypes.c_int(1), ctypes.byref(out)) != 0      # null member
    assert L.sphb_last_error_id() == b"SPHB:arg" and not out.value
    assert L.sphb_batch_create(arr, ctypes.c_int(70000), ctypes.byref(out)) != 0   # more than grid.y allows
    assert L.sphb_last_error_id() == b"SPHB:batch"
    for fn in ("sphb_batch_refresh", "sphb_batch_sync"):
        f = getattr(L, fn)
        f.argtypes = [ctypes.c_void_p]
        assert f(None) != 0 and L.sphb_last_error_id() == b"SPHB:arg"
    L.sphb_batch_run_dual.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    assert L.sphb_batch_run_dual(None, 1, 1) != 0 and L.sphb_last_error_id() == b"SPHB:arg"
    assert L.sphb_batch_launches(None) == 0
    L.sphb_batch_destroy(None)


def test_slab_group_entry_points_validate_arguments_without_device():
    """sphb_slabs_* (the C host of the sharded channel): a channel too short for the slab count is refused before any
    CUDA call; NULL handles are tolerated by the accessors and by destroy."""
    import importlib
    sphb = load_product()
    L = sphb.lib()
    cs = importlib.import_module("sph_b200.channel_setup")
    cfg = cs.ChannelConfig(DL=1.0, DH=1.0, dp=0.05, sort_interval=0).derive()    # 8 cell columns: no room for two slabs
    pt = cs.lattice(cfg)
    cp = sphb._Params(pt.n_fluid, pt.n_total, cfg.DL, cfg.DH, cfg.dp, cfg.h, cfg.rho0, cfg.p0, cfg.c_f, cfg.mu,
                      cfg.gravity_g, cfg.inv_sigma0, cfg.transport_coeff, 0, -1, 0)
    f = np.asfortranarray
    out = ctypes.c_void_p()
    p = sphb._p
    rc = L.sphb_slabs_create(ctypes.byref(cp), ctypes.c_int(2), None, ctypes.c_int(0), p(f(pt.pos)), p(f(pt.vel)),
                             p(pt.rho), p(pt.drho_dt), p(pt.mass), p(f(pt.wall_vel)), ctypes.byref(out))
    assert rc != 0 and not out.value
    assert L.sphb_last_error_id().decode() == "SPHB:arg" and b"do not fit" in L.sphb_last_error_msg()
    rc = L.sphb_slabs_create(None, ctypes.c_int(2), None, ctypes.c_int(0), None, None, None, None, None, None,
                             ctypes.byref(out))
    assert rc != 0 and L.sphb_last_error_id().decode() == "SPHB:arg"
    assert L.sphb_slabs_count(None) == 0
    assert L.sphb_slabs_step(None, ctypes.c_double(1.0)) != 0
    L.sphb_slabs_destroy.restype = None
    L.sphb_slabs_destroy(None)
