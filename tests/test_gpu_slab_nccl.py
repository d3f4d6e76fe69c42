"""Two-rank runs of the x-slab sharded channel against the single-domain run (need >= 2 GPUs; the one-GPU form of the
same kernels is covered by test_gpu_slab.py): the NCCL exchange, and the peer-to-peer mailboxes across processes
(CUDA IPC, kernels of one GPU storing into the memory of the other, no collective inside the step)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import importlib, json, os, sys
import numpy as np
import torch, torch.distributed as dist
root = sys.argv[1]
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
from conftest import load_product
sphb = load_product()
cs = importlib.import_module("sph_b200.channel_setup"); slab = importlib.import_module("sph_b200.slab")
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
cfg = cs.ChannelConfig(DL=3.0, DH=1.0, dp=0.01, c_f=10.0, sort_interval=0).derive()
pt = cs.developed(cfg, seed=21, jitter=0.2)
pt.vel[:pt.n_fluid, 0] += 2.0
# one CUDA graph per step (library kernels + NCCL all-to-all), compaction + re-capture every 16 steps
sc = slab.SlabChannel(cfg, pt, rank, world, torch.device("cuda", lr), use_graph=True, compact_every=16,
                      p2p=(sys.argv[2] == "p2p"))
n_steps = 40
for _ in range(n_steps):
    sc.step()
rows = sc.slab.owned(("pos", "vel", "rho", "drho_dt"))
st = sc.slab.ch.status()
gathered = [None] * world
dist.all_gather_object(gathered, {k: v for k, v in rows.items()})
res = {"ok": True}
if rank == 0:
    ids = np.concatenate([g["ids"] for g in gathered]); order = np.argsort(ids)
    ch = sphb.channel_from_config(cfg, pt, sort_interval=0, device=lr)
    ch.initialise()
    for _ in range(n_steps):
        ch.step()
    want = ch.download(("pos", "vel", "rho", "drho_dt")); s1 = ch.status()
    res["ids_complete"] = bool(np.array_equal(ids[order], np.arange(pt.n_fluid)))
    worst = 0.0
    for f in ("pos", "vel", "rho", "drho_dt"):
        got = np.concatenate([g[f] for g in gathered], axis=0)[order]
        w = want[f][:pt.n_fluid]
        worst = max(worst, float(np.max(np.abs(got - w)) / np.max(np.abs(w))))
    res.update(graph=bool(sc._graph_ok and sc._graph is not None), worst=worst, dt_equal=bool(abs(st["dt"] - s1["dt"]) <= 1e-15 * s1["dt"]), t_equal=bool(abs(st["t"] - s1["t"]) <= 1e-12 * s1["t"]),
               flags=int(st["error_flags"]), owned=[int(g["ids"].size) for g in gathered])
    print("RESULT " + json.dumps(res), flush=True)
sc.close()              # drops the captured graph (it holds NCCL work) before the process group goes away
torch.cuda.synchronize()
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.parametrize("mode", ["nccl", "p2p"])
def test_two_rank_slabs_match_single_domain(tmp_path, mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533" if mode == "nccl" else "29534", str(script),
                        ROOT, mode],
                       capture_output=True, text=True, timeout=180)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    res = json.loads(line[len("RESULT "):])
    assert res["ids_complete"] and res["flags"] == 0 and res["dt_equal"] and res["t_equal"] and res["graph"], res
    assert res["worst"] <= 1e-10, res       # 40 steps of independent rounding (summation order differs per slab)
