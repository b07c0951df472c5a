"""One rank of the NCCL block-cyclic solver (launch with torchrun, one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/dist_nccl.py n nb [reps] [key=value ...]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
m = importlib.import_module("mixed-precision_lu_factorization_b200")
pos = [a for a in sys.argv[1:] if "=" not in a]
kv = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
n, nb = int(pos[0]), int(pos[1])
reps = int(pos[2]) if len(pos) > 2 else 2
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("gloo")  # host-side plumbing only (unique id, timing reduction)
P, Q = m.grid_shape(world)
uid = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    uid = torch.tensor(list(m.DistSolver.unique_id()), dtype=torch.uint8)
dist.broadcast(uid, 0)
ds = m.DistSolver(local, P, Q, rank=rank, unique_id=bytes(uid.tolist()))
As, bs = ds.generate(n, nb, seed=1)
opts = m.default_options(**{k: int(v) for k, v in kv.items()})
best = None
for rep in range(reps):
    dist.barrier()
    xs, st = ds.gesv(n, nb, As, bs, opts, allow_noconv=True)
    t = torch.tensor([st.total_ms, st.factor_ms, st.solve_ms], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if best is None or t[0].item() < best[0]:
        best = t.tolist()
err = torch.tensor([(xs[0] - 1).abs().max().item()], dtype=torch.float64)
dist.all_reduce(err, op=dist.ReduceOp.MAX)
if rank == 0:
    tf = 2 / 3 * n ** 3 / (best[0] * 1e-3) / 1e12
    print(f"DIST n={n} nb={nb} grid {P}x{Q} nccl {kv}: total {best[0]:.2f} ms = {tf:.1f} TFLOP/s | factor {best[1]:.2f} solve {best[2]:.2f} "
          f"iters {st.iters} conv {st.converged} be {st.backward_error:.2e} status {st.status_bits} err {err.item():.1e}", flush=True)
ds.close()
dist.destroy_process_group()
