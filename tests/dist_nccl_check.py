"""One rank of the NCCL block-cyclic solver with result checks (test infrastructure; launched by
tests/test_gpu_configs.py through torch.distributed.run, one process per GPU):

    ... tests/dist_nccl_check.py n nb [check=1] [key=value options]

Prints  DIST n=.. iters=.. conv=.. be=.. status=.. err=..  (rank 0), and with check=1 gathers the local L\\U factors on
rank 0, re-solves the same system on rank 0's GPU with the single-GPU path and with host LAPACK, and prints
CHECK ok / CHECK FAIL <why>.
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

m = importlib.import_module("mixed-precision_lu_factorization_b200")
pos = [a for a in sys.argv[1:] if "=" not in a]
kv = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
n, nb = int(pos[0]), int(pos[1])
check = int(kv.pop("check", 0))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("gloo")  # host-side plumbing only (unique id, gathers of the check)
P, Q = m.grid_shape(world)
uid = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    uid = torch.tensor(list(m.DistSolver.unique_id()), dtype=torch.uint8)
dist.broadcast(uid, 0)
ds = m.DistSolver(local, P, Q, rank=rank, unique_id=bytes(uid.tolist()))
As, bs = ds.generate(n, nb, seed=1)
opts = m.default_options(**{k: int(v) for k, v in kv.items()})
dist.barrier()
xs, st = ds.gesv(n, nb, As, bs, opts, allow_noconv=True)
err = torch.tensor([(xs[0] - 1).abs().max().item()], dtype=torch.float64)
dist.all_reduce(err, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"DIST n={n} nb={nb} grid={P}x{Q} iters={st.iters} conv={st.converged} be={st.backward_error:.3e} "
          f"status={st.status_bits} err={err.item():.3e} ms={st.total_ms:.2f}", flush=True)

if check:
    p, q, mloc, nloc = ds.local_shape(0, n, nb)
    mine = ds.local_factors(0, n, nb)[:mloc, :nloc].cpu().contiguous()
    shapes = [None] * world
    dist.all_gather_object(shapes, (p, q, mloc, nloc))
    parts = [torch.empty(s[2], s[3], dtype=torch.float64) for s in shapes] if rank == 0 else None
    dist.gather(mine, parts, dst=0)
    if rank == 0:
        import mplu_oracle as orc
        order = [None] * (P * Q)
        for s, part in zip(shapes, parts):
            order[s[0] * Q + s[1]] = part.numpy()
        LU = m.gather_block_cyclic(order, n, nb, P, Q)
        A = orc.counter_matrix(n, seed=1)
        b = A.sum(axis=1)
        x_ref, lu_ref, piv = orc.lapack_gesv(A, b)
        s1 = m.Solver(local)
        dA = torch.tensor(A, dtype=torch.float64, device="cuda").t().contiguous().t()
        x1, st1 = s1.gesv(dA, torch.tensor(b, device="cuda"), m.default_options(nb=nb))
        LU1 = s1.factors(n).cpu().numpy()
        s1.close()
        u16 = 2.0 ** -11
        why = []
        if not np.array_equal(piv, np.arange(n)):
            why.append("lapack pivots not identity")
        if np.abs(np.triu(LU - lu_ref)).max() > 0.02 * u16 * np.abs(lu_ref).max():
            why.append("U vs lapack")
        if np.abs(np.tril(LU - lu_ref, -1)).max() > 2 * u16 * np.abs(np.tril(lu_ref, -1)).max():
            why.append("L vs lapack")
        if np.abs(np.triu(LU - LU1)).max() > 0.02 * u16 * np.abs(lu_ref).max():
            why.append("U vs single GPU")
        if np.abs(np.tril(LU - LU1, -1)).max() > 2 * u16 * np.abs(np.tril(lu_ref, -1)).max():
            why.append("L vs single GPU")
        if not (st.converged == 1 and st.iters <= st1.iters + 1 and st.backward_error <= 2 * n * 1.1e-16):
            why.append(f"refinement {st.iters} vs {st1.iters} be {st.backward_error:.2e}")
        if np.abs(xs[0].cpu().numpy() - x_ref).max() > 1e-11:
            why.append("x vs lapack")
        print("CHECK ok" if not why else "CHECK FAIL " + "; ".join(why), flush=True)
ds.close()
dist.destroy_process_group()
