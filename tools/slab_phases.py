import importlib, os, sys, time
import numpy as np, torch, torch.distributed as dist
root = "/root/repo"; sys.path.insert(0, root); sys.path.insert(0, root + "/tests")
from conftest import load_product
sphb = load_product()
cs = importlib.import_module("sph_b200.channel_setup"); slab = importlib.import_module("sph_b200.slab")
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
DLg = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
cfg = cs.ChannelConfig(DL=DLg * world, dp=0.0025, sort_interval=0).derive()
pt = cs.cell_sorted(cfg, cs.developed(cfg, seed=12345))
sc = slab.SlabChannel(cfg, pt, rank, world, torch.device("cuda", lr))
for _ in range(5): sc.step()
K = 40
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(K)]
host = np.zeros((K, 5))
torch.cuda.synchronize(); dist.barrier()
s = sc.slab
for k in range(K):
    with torch.cuda.stream(sc.stream):
        t0 = time.perf_counter(); ev[k][0].record(sc.stream)
        s.step_begin(); ev[k][1].record(sc.stream); t1 = time.perf_counter()
        s.pack(); ev[k][2].record(sc.stream); t2 = time.perf_counter()
        ops = [dist.P2POp(dist.isend, s.send_left, sc.left), dist.P2POp(dist.isend, s.send_right, sc.right),
               dist.P2POp(dist.irecv, s.recv_right, sc.right), dist.P2POp(dist.irecv, s.recv_left, sc.left)]
        works = dist.batch_isend_irecv(ops)
        dist.all_gather_into_tensor(sc._gath, s.scal[0:3])
        for w in works: w.wait()
        ev[k][3].record(sc.stream); t3 = time.perf_counter()
        slab._check(slab.lib().sphb_ctx_slab_scalars_reduce(s.ch._ctx, slab.C.c_void_p(sc._gath.data_ptr()), slab.C.c_int(world), slab.C.c_void_p(sc._red.data_ptr())))
        s.unpack(import_vmax=False); ev[k][4].record(sc.stream); t4 = time.perf_counter()
        s.step_end(); ev[k][5].record(sc.stream); t5 = time.perf_counter()
    host[k] = [t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4]
torch.cuda.synchronize()
g = np.array([[ev[k][i].elapsed_time(ev[k][i + 1]) for i in range(5)] for k in range(K)])
if rank == 0:
    print("GPU ms  [physics, pack, nccl, unpack, list+shear]:", np.round(g[5:].mean(0), 4), "sum", round(g[5:].sum(1).mean(), 4))
    print("host ms [physics, pack, nccl, unpack, list+shear]:", np.round(host[5:].mean(0), 4), "sum", round(host[5:].sum(1).mean(), 4))
dist.barrier(); dist.destroy_process_group()
