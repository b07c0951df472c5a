"""CPU tests of the host-side logic of the multi-GPU path: block-cyclic index maps, process-grid choice, and the
world_size-2 rendezvous/plumbing bench.py uses (gloo backend; the GPU kernels are not involved)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_grid_shape(mplu):
    assert mplu.grid_shape(1) == (1, 1)
    assert mplu.grid_shape(2) == (1, 2)
    assert mplu.grid_shape(4) == (2, 2)
    assert mplu.grid_shape(8) == (2, 4)
    assert mplu.grid_shape(6) == (2, 3)


@pytest.mark.parametrize("T,P", [(1, 1), (4, 2), (5, 2), (7, 4), (3, 4), (64, 2), (64, 4)])
def test_tiles_local_partition(mplu, T, P):
    counts = [mplu.tiles_local(T, P, p) for p in range(P)]
    assert sum(counts) == T
    assert max(counts) - min(counts) <= 1
    assert counts == [len(range(p, T, P)) for p in range(P)]


@pytest.mark.parametrize("nb,P", [(128, 1), (128, 2), (256, 4), (1024, 2)])
def test_index_maps_are_inverse(mplu, nb, P):
    n = nb * 9
    seen = set()
    for g in range(0, n, 37):
        p, l = mplu.global_to_local(g, nb, P)
        assert 0 <= p < P
        assert mplu.local_to_global(l, nb, P, p) == g
        seen.add((p, l))
    assert len(seen) == len(range(0, n, 37))


@pytest.mark.parametrize("P,Q", [(1, 1), (1, 2), (2, 2), (2, 4), (3, 2)])
def test_scatter_gather_roundtrip(mplu, P, Q):
    nb, T = 4, 7
    n = nb * T
    A = np.arange(n * n, dtype=np.float64).reshape(n, n)
    parts = mplu.scatter_block_cyclic(A, nb, P, Q)
    assert len(parts) == P * Q
    for p in range(P):
        for q in range(Q):
            part = parts[p * Q + q]
            assert part.shape == (mplu.tiles_local(T, P, p) * nb, mplu.tiles_local(T, Q, q) * nb)
            for li in range(0, part.shape[0], 3):
                for lj in range(0, part.shape[1], 5):
                    assert part[li, lj] == A[mplu.local_to_global(li, nb, P, p), mplu.local_to_global(lj, nb, Q, q)]
    assert np.array_equal(mplu.gather_block_cyclic(parts, n, nb, P, Q), A)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


WORKER = r"""
import importlib, os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["MPLU_ROOT"]); sys.path.insert(0, os.path.join(os.environ["MPLU_ROOT"], "oracle"))
m = importlib.import_module("mixed-precision_lu_factorization_b200")
import mplu_oracle as orc
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
P, Q = m.grid_shape(world)
p, q = rank // Q, rank % Q
# 1. the unique-id hand-off bench.py does (rank 0 creates 128 bytes, everybody receives them)
uid = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    uid = torch.tensor(list(bytes(range(128))), dtype=torch.uint8)
dist.broadcast(uid, 0)
assert bytes(uid.tolist()) == bytes(range(128))
# 2. a block-cyclic no-pivot LU + solve in numpy over the SAME layout and message pattern as csrc/dist.cu
#    (diagonal tile to everyone, L panel along the process row, U panel along the process column), checked against
#    the un-distributed oracle: validates ownership / local index logic under a real multi-process exchange
n, nb = 24, 4
T = n // nb
A = orc.counter_matrix(n, seed=3)
mine = m.scatter_block_cyclic(A, nb, P, Q)[rank].copy()
def lt(k, PP, pp): return (k - pp) // PP + 1 if k >= pp else 0     # local tiles with global index <= k
def ltx(k, PP, pp): return (k - 1 - pp) // PP + 1 if k > pp else 0  # ... < k
Dtiles = []
for k in range(T):
    pk, qk = k % P, k % Q
    D = torch.zeros(nb, nb, dtype=torch.float64)
    if (p, q) == (pk, qk):
        i0, j0 = (k // P) * nb, (k // Q) * nb
        t = mine[i0:i0 + nb, j0:j0 + nb]
        for j in range(nb):
            t[j + 1:, j] /= t[j, j]
            t[j + 1:, j + 1:] -= np.outer(t[j + 1:, j], t[j, j + 1:])
        D = torch.tensor(t)
    dist.broadcast(D, pk * Q + qk)
    D = D.numpy()
    Dtiles.append(D.copy())
    Lt, Ut = np.tril(D, -1) + np.eye(nb), np.triu(D)
    ilo, jlo = lt(k, P, p) * nb, lt(k, Q, q) * nb
    Lp = torch.zeros(mine.shape[0] - ilo, nb, dtype=torch.float64)
    Up = torch.zeros(nb, mine.shape[1] - jlo, dtype=torch.float64)
    if q == qk and Lp.numel():
        j0 = (k // Q) * nb
        mine[ilo:, j0:j0 + nb] = mine[ilo:, j0:j0 + nb] @ np.linalg.inv(Ut)
        Lp = torch.tensor(mine[ilo:, j0:j0 + nb].copy())
    if p == pk and Up.numel():
        i0 = (k // P) * nb
        mine[i0:i0 + nb, jlo:] = np.linalg.inv(Lt) @ mine[i0:i0 + nb, jlo:]
        Up = torch.tensor(mine[i0:i0 + nb, jlo:].copy())
    # row / column broadcasts emulated on the world group: every (row, root) pair in turn
    rows = [torch.zeros((m.tiles_local(T, P, pp) - lt(k, P, pp)) * nb, nb, dtype=torch.float64) for pp in range(P)]
    cols = [torch.zeros(nb, (m.tiles_local(T, Q, qq) - lt(k, Q, qq)) * nb, dtype=torch.float64) for qq in range(Q)]
    for pp in range(P):
        if p == pp and q == qk: rows[pp] = Lp
        if rows[pp].numel(): dist.broadcast(rows[pp], pp * Q + qk)
    for qq in range(Q):
        if q == qq and p == pk: cols[qq] = Up
        if cols[qq].numel(): dist.broadcast(cols[qq], pk * Q + qq)
    if mine.shape[0] > ilo and mine.shape[1] > jlo:
        mine[ilo:, jlo:] -= rows[p].numpy() @ cols[q].numpy()
parts = [torch.zeros(m.tiles_local(T, P, r // Q) * nb, m.tiles_local(T, Q, r % Q) * nb, dtype=torch.float64) for r in range(world)]
for r in range(world):
    if r == rank: parts[r] = torch.tensor(mine)
    dist.broadcast(parts[r], r)
LU = m.gather_block_cyclic([t.numpy() for t in parts], n, nb, P, Q)
ref = orc.lu_nopivot(A) if hasattr(orc, "lu_nopivot") else None
if ref is None:
    ref = A.copy()
    for j in range(n):
        ref[j + 1:, j] /= ref[j, j]
        ref[j + 1:, j + 1:] -= np.outer(ref[j + 1:, j], ref[j, j + 1:])
assert np.abs(LU - ref).max() <= 1e-9 * np.abs(ref).max(), np.abs(LU - ref).max()
# 2b. the block-cyclic triangular solves as csrc/dist.cu (solve_body) runs them: right-looking with look-ahead.  Once the
#     solution block of tile column k is known, the ranks of process column k mod Q add W_loc(:, tile column k) * block into
#     full-length partial sums -- the next D tile rows ("near", on the chain) and every tile row beyond ("far", off the chain) --
#     and a step is one all-reduce of the tile row's slice (ranks outside its process row hold zeros) + the replicated
#     diagonal tile's solve.  Same ownership / range arithmetic (cnt_le, cnt_lt), checked against a dense solve.
rhs = A @ np.arange(1.0, n + 1.0)
mt = m.tiles_local(T, P, p)
for D_ in (1, 2, 3):
    y = np.zeros(n); x = np.zeros(n)
    for sweep in (0, 1):
        part = np.zeros(n)
        sol = y if sweep == 0 else x
        for kk in range(T):
            k = kk if sweep == 0 else T - 1 - kk
            sl = slice(k * nb, (k + 1) * nb)
            red = torch.tensor(part[sl].copy())
            if kk > 0: dist.all_reduce(red)
            Dk = Dtiles[k]
            if sweep == 0: sol[sl] = np.linalg.solve(np.tril(Dk, -1) + np.eye(nb), rhs[sl] - red.numpy())
            else: sol[sl] = np.linalg.solve(np.triu(Dk), y[sl] - red.numpy())
            if kk == T - 1 or q != k % Q or mine.size == 0: continue
            if sweep == 0:
                near0, near1 = lt(k, P, p), lt(min(k + D_, T - 1), P, p); far0, far1 = near1, mt
            else:
                near0, near1 = ltx(max(k - D_, 0), P, p), ltx(k, P, p); far0, far1 = 0, near0
            c0 = (k // Q) * nb
            for t0, t1 in ((near0, near1), (far0, far1)):
                for tl in range(t0, t1):
                    g = tl * P + p  # global tile row of local tile row tl
                    assert (g > k if sweep == 0 else g < k)
                    part[g * nb:(g + 1) * nb] += mine[tl * nb:(tl + 1) * nb, c0:c0 + nb] @ sol[sl]
    assert np.abs(x - np.arange(1.0, n + 1.0)).max() < 1e-8, (D_, np.abs(x - np.arange(1.0, n + 1.0)).max())
# 3. max-over-ranks timing reduction used by bench.py
t = torch.tensor([float(rank + 1)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.item() == world
dist.barrier()
if rank == 0:
    print("WORKER_OK")
dist.destroy_process_group()
"""


@pytest.mark.parametrize("world", [2, 4])
def test_block_cyclic_exchange_pattern_gloo(tmp_path, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MPLU_ROOT=ROOT, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "WORKER_OK" in r.stdout
