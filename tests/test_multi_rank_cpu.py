"""CPU, world_size 2 (gloo): the host-side logic of the N > 1 path.

The GPU path shards the permuted system by control-point rows (no data-path collective in the assembly) and the
Krylov solvers all-gather one padded vector per matvec (machline_b200/shard.py is the host statement of that layout,
csrc/gpu/solve_kernels.cu Sys::matvec the device one).  Here two gloo ranks play the two GPUs with the ORACLE as the
per-rank compute (test infrastructure standing in for the CUDA kernels, which cannot run on this box):
  * each rank assembles only its own rows; stacked, they are the full matrix, bit for bit;
  * the sharded matvec + padded all-gather + compaction reproduces the full matvec, and a GMRES driven by it
    reaches the oracle's solution;
  * bench.py's reference arm under torchrun: rank 0 alone prints one JSON line, the other rank exits 0."""
import json
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_row_shards_cover_every_row_once():
    from machline_b200 import shard
    for n in [1, 63, 64, 65, 1202, 10513, 100000]:
        for world in [1, 2, 3, 4, 8]:
            sh = shard.all_shards(n, world)
            assert sh[0][0] == 0 and sum(nr for _, nr in sh) == n
            for (a0, an), (b0, _) in zip(sh, sh[1:]):
                assert a0 + an == b0
            assert all(r0 % 64 == 0 or nr == 0 for r0, nr in sh)          # aligned shard starts
            assert max(nr for _, nr in sh) - min(nr for _, nr in sh if nr) <= 64 * world or n < 64 * world
            pad = shard.shard_pad(sh)
            assert pad % 64 == 0 and pad >= max(nr for _, nr in sh)
    with pytest.raises(ValueError):
        shard.row_shard(10, 2, 2)


def test_cyclic_rows_cover_every_row_once():
    from machline_b200 import shard
    for n in [1, 63, 128, 129, 1202, 10513]:
        for world in [1, 2, 3, 8]:
            for block in [64, 128]:
                rows = [shard.cyclic_rows(n, r, world, block) for r in range(world)]
                allr = np.sort(np.concatenate(rows))
                assert (allr == np.arange(n)).all()
                assert all((np.diff(r) > 0).all() for r in rows)
                assert max(len(r) for r in rows) - min(len(r) for r in rows) <= block
                for r, rr in enumerate(rows):
                    assert ((rr // block) % world == r).all()


def _worker(rank: int, world: int, port: int, out_dir: str):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import fixtures
    import oracle_binding as ob
    from machline_b200 import shard

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        case, _, _ = fixtures.make_case("test_08")      # sphere, 1202 unknowns
        N = case.n_cp
        shards = shard.all_shards(N, world)
        row0, nrows = shards[rank]
        pad = shard.shard_pad(shards)
        A_loc, I_loc = ob.assemble(case, row0=row0, nrows=nrows)           # this rank's rows only

        def gather_vec(v_loc):
            buf = torch.zeros(pad, dtype=torch.float64)
            buf[:nrows] = torch.from_numpy(np.ascontiguousarray(v_loc))
            allb = [torch.zeros(pad, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(allb, buf)
            return shard.compact_gathered(torch.cat(allb).numpy(), shards, pad)

        I_full = gather_vec(I_loc)
        b = np.asarray(case.BC) - I_full

        def matvec(q):                                                     # Sys::matvec: local rows, all-gather, compact
            return gather_vec(A_loc @ q)

        # GMRES (x0 = 0, MGS, Givens; linalg.f90:1235-1334) on the unscaled system, replicated small problem per rank
        tol, kmax = 1e-12, 200
        beta = np.linalg.norm(b)
        Q = [b / beta]
        H = np.zeros((kmax + 1, kmax))
        cs, sn, e = np.zeros(kmax), np.zeros(kmax), np.zeros(kmax + 1)
        e[0] = beta
        k_done = 0
        for k in range(kmax):
            w = matvec(Q[k])
            for i in range(k + 1):
                H[i, k] = w @ Q[i]
                w = w - H[i, k] * Q[i]
            H[k + 1, k] = np.linalg.norm(w)
            Q.append(w / H[k + 1, k])
            for i in range(k):
                t = cs[i] * H[i, k] + sn[i] * H[i + 1, k]
                H[i + 1, k] = -sn[i] * H[i, k] + cs[i] * H[i + 1, k]
                H[i, k] = t
            d = np.hypot(H[k, k], H[k + 1, k])
            cs[k], sn[k] = H[k, k] / d, H[k + 1, k] / d
            H[k, k], H[k + 1, k] = d, 0.0
            e[k + 1] = -sn[k] * e[k]
            e[k] = cs[k] * e[k]
            k_done = k + 1
            if abs(e[k + 1]) < tol:
                break
        y = np.linalg.solve(np.triu(H[:k_done, :k_done]), e[:k_done])
        x = sum(yi * qi for yi, qi in zip(y, Q))
        res = np.linalg.norm(matvec(x) - b)

        # every rank must hold the same replicated vectors
        chk = torch.tensor([float(x.sum()), float(res)], dtype=torch.float64)
        allc = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(allc, chk)
        same = all(torch.equal(allc[0], c) for c in allc)
        np.savez(Path(out_dir) / f"rank{rank}.npz", A_loc=A_loc, row0=row0, nrows=nrows, x=x, res=res, iters=k_done, same=same,
                 I_full=I_full)
        case.close()
    finally:
        dist.destroy_process_group()


def test_two_rank_row_sharded_assembly_and_gmres(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sys.path.insert(0, str(ROOT / "tests"))
    import fixtures
    import oracle_binding as ob
    case, _, _ = fixtures.make_case("test_08")
    A_ref, I_ref = ob.assemble(case)
    opts = case.solver_opts()
    opts.preconditioner = 0
    x_ref, info_ref = ob.solve_system(A_ref, I_ref, case.BC, opts)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(2)]
    assert int(r[0]["row0"]) == 0 and int(r[0]["nrows"]) + int(r[1]["nrows"]) == case.n_cp
    assert int(r[1]["row0"]) == int(r[0]["nrows"])
    assert np.array_equal(np.vstack([r[0]["A_loc"], r[1]["A_loc"]]), A_ref)     # rows do not depend on who builds them
    for k in range(2):
        assert bool(r[k]["same"])
        assert np.array_equal(r[k]["I_full"], I_ref)
        assert float(r[k]["res"]) < 1e-10
        assert abs(int(r[k]["iters"]) - info_ref.iterations) <= 1
        assert np.abs(r[k]["x"] - x_ref).max() <= 1e-9 * np.abs(x_ref).max()
    case.close()


def test_bench_reference_arm_under_torchrun_two_ranks():
    """Driver contract for --impl reference at N > 1: launched through torch.distributed.run, rank 0 alone runs and
    prints ONE JSON line; the other rank exits 0 without work.  (torchrun exports OMP_NUM_THREADS=1; the arm resets it.)"""
    port = _free_port()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
           "--dims", "16x8"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, res.stdout
    doc = json.loads(lines[0])
    assert doc["impl"] == "reference" and doc["n_gpus"] == 2 and doc["value"] > 0
    assert doc["metric"] == "aic_pair_influences_per_s" and doc["higher_is_better"] is True
    assert doc["cpu_baseline"]["kind"] == "port" and doc["cpu_baseline"]["cores"] >= 1
    assert doc["e2e"]["h2d_bytes_per_step"] == 0 and doc["e2e"]["value"] == doc["value"]


def test_bench_b200_arm_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--dims", "16x8"], capture_output=True, text=True,
                         timeout=600, cwd=str(ROOT))
    assert res.returncode != 0
    assert not any(ln.startswith("{") for ln in res.stdout.splitlines())      # no number without the CUDA path


# ---- row-sharded LU: the algorithm of csrc/gpu/lu_sharded.cu stated with numpy + gloo collectives --------------------
def _sharded_lu_model(A_loc, my_rows, b, N, rank, world, dist, torch, nb=64):
    """Every rank keeps its rows in place.  perm[position] = (owner, local row); per panel: all-gather the panel columns,
    factor them redundantly (implicit scaling, LAST maximal row on ties, linalg.f90:242), sum the pivot rows' trailing
    entries over the ranks (one non-zero contributor each), triangular solve, local rank-nb update of the rows still
    below the panel; the right-hand side rides along as column N.  Returns (x, pivot positions)."""
    S = max(int(v) for v in _all_gather_obj(dist, len(my_rows)))
    n_loc = len(my_rows)
    M = np.zeros((n_loc, N + 1))
    M[:, :N] = A_loc
    M[:, N] = b[my_rows]
    rows_all = _all_gather_obj(dist, list(map(int, my_rows)))
    slot_of_g = {}
    for r, rows in enumerate(rows_all):
        for i, g in enumerate(rows):
            slot_of_g[g] = (r, i)
    perm = [slot_of_g[g] for g in range(N)]                       # position -> (rank, local row)

    def allgather_rows(x_loc, width):                              # [n_loc, width] -> dict (rank, i) -> row
        buf = torch.zeros(S, width, dtype=torch.float64)
        buf[:n_loc] = torch.from_numpy(np.ascontiguousarray(x_loc))
        out = [torch.zeros(S, width, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(out, buf)
        return [o.numpy() for o in out]

    amax = allgather_rows(np.abs(M[:, :N]).max(axis=1, keepdims=True), 1)
    vv = np.array([1.0 / amax[r][i, 0] for r, i in perm])        # by position
    piv_pos, D = [], {}
    for k0 in range(0, N, nb):
        k1 = min(k0 + nb, N)
        w = k1 - k0
        G = allgather_rows(M[:, k0:k1], w)
        P = np.array([G[r][i] for r, i in perm])                   # panel in position order (rows < k0 unused)
        for j in range(k0, k1):
            cand = vv[j:] * np.abs(P[j:, j - k0])
            p = j + int(np.flatnonzero(cand == cand.max())[-1])    # last maximal row
            piv_pos.append(p)
            if p != j:
                P[[j, p]] = P[[p, j]]
                perm[j], perm[p] = perm[p], perm[j]
                vv[p] = vv[j]
            P[j + 1:, j - k0] *= 1.0 / P[j, j - k0]
            P[j + 1:, j - k0 + 1:] -= np.outer(P[j + 1:, j - k0], P[j, j - k0 + 1:])
        D[k0] = P[k0:k1, :].copy()
        # U rows: owners contribute, everyone sums
        Wd = N + 1 - k1
        U = torch.zeros(w, Wd, dtype=torch.float64)
        for jj in range(w):
            r, i = perm[k0 + jj]
            if r == rank:
                U[jj] = torch.from_numpy(M[i, k1:])
        dist.all_reduce(U)
        U = U.numpy()
        Lk = np.tril(D[k0], -1) + np.eye(w)
        U = np.linalg.solve(Lk, U) if w > 1 else U
        for jj in range(w):
            r, i = perm[k0 + jj]
            if r == rank:
                M[i, k1:] = U[jj]
        pos_of = {perm[pos][1]: pos for pos in range(k1, N) if perm[pos][0] == rank}
        for i, pos in pos_of.items():                               # local rows still below the panel
            M[i, k1:] -= P[pos, :] @ U
    x = np.zeros(N)
    for k0 in reversed(range(0, N, nb)):
        k1 = min(k0 + nb, N)
        y = torch.zeros(k1 - k0, dtype=torch.float64)
        for jj in range(k1 - k0):
            r, i = perm[k0 + jj]
            if r == rank:
                y[jj] = M[i, N]
        dist.all_reduce(y)
        xk = np.linalg.solve(np.triu(D[k0]), y.numpy())
        x[k0:k1] = xk
        M[:, N] -= M[:, k0:k1] @ xk
    return x, piv_pos


def _all_gather_obj(dist, obj):
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def _lu_worker(rank: int, world: int, port: int, out_dir: str, cyclic: int):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    from machline_b200 import shard
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        N = 300
        rng = np.random.default_rng(7)
        A = rng.standard_normal((N, N))
        A[::5] *= 1e3
        A[10] = A[11]                      # exact ties of vv * |a| in every column between two rows ...
        A[11, 200] += 1.0                  # ... that are not copies of each other
        b = rng.standard_normal(N)
        if cyclic:
            rows = shard.cyclic_rows(N, rank, world, cyclic)
        else:
            r0, nr = shard.row_shard(N, rank, world)
            rows = np.arange(r0, r0 + nr)
        x, piv = _sharded_lu_model(A[rows], rows, b, N, rank, world, dist, torch)
        np.savez(Path(out_dir) / f"lu_rank{rank}.npz", x=x, piv=np.array(piv), A=A, b=b)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("cyclic", [0, 64])
def test_two_rank_sharded_lu_algorithm(tmp_path, cyclic):
    """The sharded LU's bookkeeping (positions vs slots, replicated panel, exact U-row sums, right-hand side as a column,
    distributed back substitution) on two gloo ranks: same pivot sequence as the oracle's lu_decomp and the same solution."""
    import ctypes as C
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_lu_worker, args=(2, port, str(tmp_path), cyclic), nprocs=2, join=True)
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_binding as ob
    from machline_b200 import _abi
    r = [np.load(tmp_path / f"lu_rank{k}.npz") for k in range(2)]
    A, b = r[0]["A"], r[0]["b"]
    assert np.array_equal(r[0]["x"], r[1]["x"]) and np.array_equal(r[0]["piv"], r[1]["piv"])
    x_or, _ = ob.solve_system(np.asfortranarray(A), np.zeros(len(b)), b, _abi.solver_opts("LU"))
    assert np.abs(r[0]["x"] - x_or).max() <= 1e-9 * np.abs(x_or).max()
    # pivot positions = the oracle's indx (0-based)
    N = len(b)
    Af = np.asfortranarray(A.copy())
    indx = np.zeros(N, dtype=np.int32)
    L = ob.lib()
    if hasattr(L, "orc_lu_decomp"):
        L.orc_lu_decomp.argtypes = [C.c_int, _abi.c_double_p, _abi.c_int_p]
        assert L.orc_lu_decomp(N, Af.ctypes.data_as(_abi.c_double_p), indx.ctypes.data_as(_abi.c_int_p)) == 0
        assert np.array_equal(indx, r[0]["piv"])


def _bjac_worker(rank: int, world: int, port: int, out_dir: str, cyclic: int):
    """The algorithm of block_jacobi_sharded (csrc/gpu/lu_kernels.cu) on gloo ranks: every rank keeps its rows; the diagonal
    blocks are assembled with one all-reduce each (one non-zero contributor per entry) and solved on every rank; the block
    right-hand sides and the residual of the local rows travel through an all-gather of the padded local parts."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    from machline_b200 import shard
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        N, bs, rel, tol = 203, 41, 0.9, 1e-12
        rng = np.random.default_rng(3)
        A = rng.standard_normal((N, N)) * 0.05 + np.diag(2.0 + rng.random(N))      # diagonally dominant: block Jacobi converges
        b = rng.standard_normal(N)
        if cyclic:
            rows = shard.cyclic_rows(N, rank, world, cyclic)
        else:
            r0, nr = shard.row_shard(N, rank, world)
            rows = np.arange(r0, r0 + nr)
        A_loc = A[rows]
        n_blocks = (N + bs - 1) // bs
        blocks = []
        for i in range(n_blocks):
            s, e = i * bs, min(N, (i + 1) * bs)
            Bi = torch.zeros((e - s, e - s), dtype=torch.float64)
            mine = (rows >= s) & (rows < e)
            Bi[torch.from_numpy(rows[mine] - s)] = torch.from_numpy(A_loc[mine][:, s:e])
            dist.all_reduce(Bi)                                   # exact: one contributor per entry
            blocks.append(np.linalg.inv(Bi.numpy()))

        def exchange(v_loc):                                      # local parts -> the replicated vector
            pad = int(max(_all_gather_obj(dist, len(rows))))
            buf = torch.zeros(pad, dtype=torch.float64)
            buf[:len(rows)] = torch.from_numpy(np.ascontiguousarray(v_loc))
            out = [torch.zeros(pad, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(out, buf)
            all_rows = _all_gather_obj(dist, rows)
            full = np.zeros(N)
            for r in range(world):
                full[all_rows[r]] = out[r].numpy()[:len(all_rows[r])]
            return full

        blk = rows // bs
        x = exchange(b[rows] / A_loc[np.arange(len(rows)), rows])  # linalg.f90:645-647
        it, err = 0, 1.0
        while err >= tol and it < 500:
            it += 1
            rhs_loc = np.empty(len(rows))
            for k, g in enumerate(rows):                          # b_i - sum over the columns outside the row's block
                s, e = blk[k] * bs, min(N, (blk[k] + 1) * bs)
                rhs_loc[k] = b[g] - (A_loc[k] @ x - A_loc[k, s:e] @ x[s:e])
            rhs = exchange(rhs_loc)
            x_new = np.concatenate([blocks[i] @ rhs[i * bs:min(N, (i + 1) * bs)] for i in range(n_blocks)])
            x_new = (1. - rel) * x + rel * x_new
            err = float(np.linalg.norm(exchange(b[rows] - A_loc @ x_new)))
            x = x_new
        np.savez(Path(out_dir) / f"bjac_rank{rank}.npz", x=x, it=it, A=A, b=b, bs=bs, rel=rel)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("cyclic", [0, 64])
def test_two_rank_sharded_block_jacobi_algorithm(tmp_path, cyclic):
    """block_jacobi_solve (common/linalg.f90:601-728) on a row-sharded system, two gloo ranks: both ranks hold the same x bit for
    bit, the same iteration count as the oracle's block Jacobi on the whole matrix, the same solution."""
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_bjac_worker, args=(2, port, str(tmp_path), cyclic), nprocs=2, join=True)
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_binding as ob
    from machline_b200 import _abi
    r = [np.load(tmp_path / f"bjac_rank{k}.npz") for k in range(2)]
    assert np.array_equal(r[0]["x"], r[1]["x"]) and int(r[0]["it"]) == int(r[1]["it"])
    A, b = r[0]["A"], r[0]["b"]
    opts = _abi.solver_opts("BJAC", preconditioner="none", rel=float(r[0]["rel"]), block_size=int(r[0]["bs"]))
    x_or, info = ob.solve_system(np.asfortranarray(A), np.zeros(len(b)), b, opts)
    assert abs(int(r[0]["it"]) - info.iterations) <= 1
    assert np.abs(r[0]["x"] - x_or).max() <= 1e-10 * np.abs(x_or).max()
