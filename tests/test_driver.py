"""The standalone C driver (sph-poiseuille-flow_b200/driver/sph_poiseuille_main.c): config.ini
grammar and derived parameters on the CPU; a short run against the golden trajectory on the GPU."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

DRIVER = os.path.join(ROOT, "sph-poiseuille-flow_b200", "sph_poiseuille_b200")

INI = """# SPH Poiseuille Flow Configuration File
; a semicolon comment

[physical]
DL = 3.0          # length, inline comment
DH = 1.0 ; inline semicolon comment
dp = {dp}
rho0 = 1.0
mu = 0.1
U_bulk = 0.666667  # U_max = 1.5*U_bulk
c_f = {c_f}
not a key value line
[simulation]
end_time = {t_end}
output_interval = {oi}
sort_interval = 100
restart_from_file = {restart}
"""


def _driver():
    sim = os.environ.get("SPHB_CPUSIM_DRIVER")   # test infrastructure: the driver linked against tests/cpusim
    if sim:
        return sim
    if not os.path.exists(DRIVER):
        import __graft_entry__ as ge
        ge.build()
    return DRIVER


def _write(tmp_path, **kw):
    args = dict(dp=0.05, c_f=15.0, t_end=20.0, oi=1.0, restart=0)
    args.update(kw)
    path = tmp_path / "config.ini"
    path.write_text(INI.format(**args))
    return str(path)


def test_dry_run_parses_reference_ini_grammar(tmp_path, H):
    cfgp = _write(tmp_path)
    out = subprocess.run([_driver(), cfgp, "--dry-run"], capture_output=True, text=True, check=True).stdout
    prm = H.derive_params(dp=0.05, c_f=15.0)
    st = H.init_particles(prm)
    m = re.search(r"n_fluid=(\d+) n_wall=(\d+) n_total=(\d+), n_bins=(\d+), tau_target=([\d.]+)", out)
    assert (int(m[1]), int(m[3])) == (st.n_fluid, st.n_total) and int(m[4]) == 20
    assert abs(float(m[5]) - prm.gravity_g * prm.rho0 * prm.DH / 2) < 1e-6
    vals = dict(re.findall(r"(\w+)=([-+0-9.e]+)", out.splitlines()[1]))
    assert float(vals["h"]) == pytest.approx(prm.h, rel=1e-12)
    assert float(vals["g"]) == pytest.approx(prm.gravity_g, rel=1e-11)
    assert float(vals["p0"]) == pytest.approx(prm.p0, rel=1e-12)
    assert float(vals["wall"]) == pytest.approx(prm.wall_thickness, rel=1e-12)
    # the restart signature string of the reference (m:514-517)
    assert "signature: DL=3|DH=1|dp=0.05|rho0=1|mu=0.1|Ub=0.666667|cf=15|t=20|oi=1|si=100|wall=thick-wall-noslip-dual-dt" in out


def test_env_override_and_errors(tmp_path):
    cfgp = _write(tmp_path, dp=0.04, c_f=10.0)
    env = dict(os.environ, SPH_CONFIG_OVERRIDE=cfgp)
    out = subprocess.run([_driver(), "--dry-run"], capture_output=True, text=True, env=env, check=True).stdout
    assert "n_fluid=1875 n_wall=600 n_total=2475" in out      # README defaults
    bad = tmp_path / "bad.ini"
    bad.write_text("[physical]\nDL = 3.0\n")
    r = subprocess.run([_driver(), str(bad), "--dry-run"], capture_output=True, text=True)
    assert r.returncode != 0 and "missing key" in r.stderr
    r = subprocess.run([_driver(), str(tmp_path / "nope.ini"), "--dry-run"], capture_output=True, text=True)
    assert r.returncode != 0 and "not found" in r.stderr


def test_driver_fails_loudly_without_gpu(tmp_path, sphb):
    if sphb.device_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([_driver(), _write(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 2 and "SPHB:cuda" in r.stderr


@pytest.mark.gpu
def test_driver_run_matches_golden_trajectory(tmp_path, golden):
    """120 main-loop steps from the dp=0.05 lattice: the log lines carry the reference's pair counts
    and wall shear (golden trajectory, reference MEX run); restart round trip."""
    g = golden("traj_dp005.npz")
    cfgp = _write(tmp_path)
    env = dict(os.environ, SPH_RESTART_PATH_OVERRIDE=str(tmp_path / "restart.sphb"),
               SPH_PROFILE_CSV_OVERRIDE=str(tmp_path / "profile.csv"))
    r = subprocess.run([_driver(), cfgp, "--max-steps", "120"], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    logs = re.findall(r"step=(\d+), t=([\d.]+)/[\d.]+, dt=([\d.e+-]+), pairs=(\d+), vmax=([\d.]+)\n"
                      r"  \[thick-wall-noslip\] tau_bot=([-\d.]+), tau_top=([-\d.]+), tau_target=([\d.]+)", r.stdout)
    assert [int(x[0]) for x in logs] == [20, 40, 60, 80, 100, 120]
    for x in logs:
        k = int(x[0])
        assert int(x[3]) == g["n_pairs"][k - 1]
        assert abs(float(x[2]) - g["dt"][k - 1]) <= 1e-4 * g["dt"][k - 1]
        assert abs(float(x[5]) - g["tau"][k - 1][0]) < 1e-4 and abs(float(x[6]) - g["tau"][k - 1][1]) < 1e-4
    assert abs(float(logs[-1][1]) - float(g["s120_t"])) < 1e-6
    assert "L2 relative error" in r.stdout
    prof = np.genfromtxt(tmp_path / "profile.csv", delimiter=",", names=True)
    assert prof.shape[0] == 20 and np.all(np.isfinite(prof["u_mean"]))
    # restart: written at the output point, accepted by a run with restart_from_file = 1
    assert (tmp_path / "restart.sphb").exists()
    cfg2 = _write(tmp_path, restart=1)
    r2 = subprocess.run([_driver(), cfg2, "--max-steps", "20"], capture_output=True, text=True, env=env)
    assert r2.returncode == 0, r2.stderr
    assert "restart loaded: t=" in r2.stdout and "step=140," in r2.stdout


@pytest.mark.gpu
def test_driver_sharded_run_matches_the_single_domain_run(tmp_path):
    """`--slabs 2`: the C host of the sharded channel (sphb_slabs_*, no Python in the loop) against the same driver on
    one context: same step count and clock, dt / max|v| / wall shear within the print precision, same L2, and a restart
    written by the sharded run is accepted by a single-context run."""
    cfgp = _write(tmp_path, dp=0.04, c_f=10.0)
    out = {}
    for tag, extra in (("one", []), ("slabs", ["--slabs", "2"])):
        env = dict(os.environ, SPH_RESTART_PATH_OVERRIDE=str(tmp_path / f"restart_{tag}.sphb"),
                   SPH_PROFILE_CSV_OVERRIDE=str(tmp_path / f"profile_{tag}.csv"))
        r = subprocess.run([_driver(), cfgp, "--max-steps", "60"] + extra, capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr + r.stdout[-400:]
        out[tag] = r.stdout
    assert "sharded: 2 x-slabs: columns [0, " in out["slabs"]
    pat = (r"step=(\d+), t=([\d.]+)/[\d.]+, dt=([\d.e+-]+), pairs=(\d+), vmax=([\d.]+)\n"
           r"  \[thick-wall-noslip\] tau_bot=([-\d.]+), tau_top=([-\d.]+), tau_target=([\d.]+)")
    a, b = re.findall(pat, out["one"]), re.findall(pat, out["slabs"])
    assert [int(x[0]) for x in a] == [20, 40, 60] == [int(x[0]) for x in b]
    for x, y in zip(a, b):
        assert x[1] == y[1]                                              # t to the printed 1e-6
        assert abs(float(x[2]) - float(y[2])) <= 2e-4 * float(x[2])      # dt (4 digits printed)
        assert abs(float(x[4]) - float(y[4])) <= 2e-4                    # max|v|
        assert abs(float(x[5]) - float(y[5])) <= 2e-4 and abs(float(x[6]) - float(y[6])) <= 2e-4
    l2 = [float(re.search(r"L2 relative error = ([\d.]+)%", out[k]).group(1)) for k in ("one", "slabs")]
    assert abs(l2[0] - l2[1]) <= 1e-3
    pa = np.genfromtxt(tmp_path / "profile_one.csv", delimiter=",", names=True)
    pb = np.genfromtxt(tmp_path / "profile_slabs.csv", delimiter=",", names=True)
    assert np.allclose(pa["u_mean"], pb["u_mean"], rtol=0, atol=1e-9)
    cfg2 = _write(tmp_path, dp=0.04, c_f=10.0, restart=1)
    env = dict(os.environ, SPH_RESTART_PATH_OVERRIDE=str(tmp_path / "restart_slabs.sphb"),
               SPH_PROFILE_CSV_OVERRIDE=str(tmp_path / "profile_2.csv"))
    r2 = subprocess.run([_driver(), cfg2, "--max-steps", "20"], capture_output=True, text=True, env=env)
    assert r2.returncode == 0, r2.stderr
    assert "restart loaded: t=" in r2.stdout and "step=80," in r2.stdout
