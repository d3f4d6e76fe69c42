"""CPU-only checks of the built device code (cuobjdump on libsph_b200.so; nothing is executed): the library holds
sm_100a code only, every kernel of the step exists in its single-channel and its batched form, and the register /
stack budgets the occupancy reasoning of DESIGN.md 3 relies on still hold (a regression here is a silent
slow-down on the GPU, which this container cannot measure)."""
import os
import re
import shutil
import subprocess

import pytest

from conftest import ROOT

LIB = os.path.join(ROOT, "sph-poiseuille-flow_b200", "libsph_b200.so")


def _res_usage():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(LIB):
        import __graft_entry__ as ge
        ge.build()
    out = subprocess.run([exe, "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"arch = (\S+)", out))
    kernels = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", out):
        name = m.group(1)
        short = re.search(r"\d+((?:k_|scan_)\w+?)(?:I[A-Za-z0-9_]*?E{1,3}v|E)(?=N|P|i|d|m|v)", name)
        key = short.group(1) if short else name
        rec = dict(reg=int(m.group(2)), stack=int(m.group(3)), shared=int(m.group(4)), local=int(m.group(5)))
        kernels.setdefault(key, []).append(rec)
    return archs, kernels


def test_library_is_sm_100a_only_and_holds_every_step_kernel():
    archs, k = _res_usage()
    assert archs == {"sm_100a"}, archs
    # (canonical cell order and tile descriptors are made inside the walk's CTA since round 2: no kernels of their own)
    step = ["k_yrange_grid", "k_bin_count", "scan_tiles", "scan_finish", "k_scatter",
            "k_walk_tile", "k_density", "k_kgc", "k_force_fused", "k_continuity", "k_time_step",
            "k_pressure", "k_wall_shear"]
    for name in step:
        assert name in k, f"{name} missing from libsph_b200.so"
        assert name + "_batch" in k, f"{name}_batch missing from libsph_b200.so"
    assert "k_zero_cells_batch" in k and "k_refresh_mirrors" in k
    # MEX-path (pair list) operators and the slab kernels
    for name in ("k_dens_sigma", "k_visc_pairs", "k_transp_pairs", "k_int1_pairs", "k_int2_pairs", "k_shear_pairs",
                 "k_export_pairs", "k_slab_scatter", "k_slab_unpack"):
        assert name in k, name


def test_gather_kernels_keep_their_register_budget_and_do_not_spill():
    """128 threads per CTA, two staging buffers per CTA (ops_csr.cu run_tiles): the fused force kernel is held to
    three CTAs per SM by its shared memory (<= 170 registers), continuity to five (<= 102), KGC / density to six
    (<= 85); a few spilled words at most (measured: every further CTA bought with spills costs time)."""
    _, k = _res_usage()
    budget = {"k_force_fused": 170, "k_continuity": 102, "k_kgc": 85, "k_density": 85, "k_wall_shear": 96,
              "k_pressure": 128}
    for name, regs in budget.items():
        for rec in k[name]:
            assert rec["reg"] <= regs, (name, rec)
            assert rec["stack"] <= 16 and rec["local"] == 0, (name, rec)
    # the batched forms read their arguments from a device record instead of the constant bank: same bounds, a few
    # more spilled words
    for name, regs in {"k_force_fused_batch": 170, "k_continuity_batch": 102, "k_kgc_batch": 85, "k_density_batch": 85,
                       "k_pressure_batch": 96, "k_wall_shear_batch": 96}.items():
        for rec in k[name]:
            assert rec["reg"] <= regs and rec["stack"] <= 64 and rec["local"] == 0, (name, rec)
    # the walk is capped at 64 registers for 8 resident CTAs (__launch_bounds__(128, 8): measured faster than 6 CTAs
    # of 80 registers, r2 call 36, although it spills ~30 words); more than that is a regression
    for rec in k["k_walk_tile"] + k["k_walk_tile_batch"]:
        assert rec["reg"] <= 64 and rec["stack"] <= 144, rec
