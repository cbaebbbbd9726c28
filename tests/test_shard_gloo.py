"""The N>1 path on CPU: world_size-2 gloo processes shard a batch of streams by contiguous ranges, each rank
inflates its range (the CPU oracle stands in for the GPU here — test infrastructure), results are gathered and
checked for completeness and order.  Also: bench.py's reference arm under a 2-rank launch (rank 0 alone prints)."""
import json
import os
import subprocess
import sys
import zlib

import numpy as np
import pytest

import workloads
from decompress_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_properties():
    rng = np.random.default_rng(0)
    for n in (0, 1, 2, 7, 100, 4096):
        w = rng.integers(1, 100000, n)
        for world in (1, 2, 3, 8):
            parts = shard.partition(w, world)
            assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
            if n >= 16 * world:
                loads = [w[a:b].sum() for a, b in parts]
                assert max(loads) <= 1.2 * (w.sum() / world) + w.max()
            # the tensor form (what the C4 driver runs on the device) makes the same cuts
            import torch
            cuts = shard.partition_cuts(torch.from_numpy(np.asarray(w, dtype=np.int64)), world).tolist()
            assert cuts == [parts[0][0]] + [b for _, b in parts], (n, world)


def _worker(rank, world, port, n, q):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from tests import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comps, plains = workloads.c2_streams(n, size=8192, workers=1)
    lo, hi = shard.my_range([len(c) + 8192 for c in comps], rank, world)
    rows = []
    for k in range(lo, hi):
        rc, cons, prod, out = oracle.zl_inflate_ns(comps[k], 8192)
        rows.append([k, rc, cons, prod, zlib.adler32(out)])
    local = torch.tensor(rows, dtype=torch.int64).reshape(-1, 5)
    allrows = shard.gather_results(local)
    dist.barrier()
    if rank == 0:
        q.put(allrows.numpy().tolist())
    dist.destroy_process_group()


def test_two_ranks_gloo_shard_and_gather():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n, world, port = 24, 2, 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    rows = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    comps, plains = workloads.c2_streams(n, size=8192, workers=1)
    assert [r[0] for r in rows] == list(range(n))            # every unit exactly once, in order
    for k, rc, cons, prod, ad in rows:
        assert (rc, cons, prod) == (0, len(comps[k]), 8192) and ad == zlib.adler32(plains[k])


def test_bench_reference_arm_two_ranks():
    """`bench.py --impl reference` under torchrun-style env: rank 0 prints the line, rank 1 exits 0 silently."""
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999")
    outs = []
    for rank in (0, 1):
        e = dict(env, RANK=str(rank), LOCAL_RANK=str(rank))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                            "--warmup", "1", "--streams", "64"], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-400:]
        outs.append(r.stdout.strip())
    line = json.loads(outs[0].splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and outs[1] == ""


def _worker_c4(rank, world, port, q):
    """A batch that lives on rank 0 (BASELINE C4): index broadcast, scatter of compressed slabs, decode, gather of the
    output slabs in pieces — shard.py's collectives over gloo; the CPU oracle stands in for the GPU (test infrastructure)."""
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from tests import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    index = blob = out_full = None
    if rank == 0:
        members = workloads.c4_members(400000, seed=3, lo=8000, hi=60000)
        lens = np.array([len(m) for m, _ in members], dtype=np.int64)
        sizes = np.array([len(p) for _, p in members], dtype=np.int64)
        offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
        index = np.stack([offs, lens, sizes], axis=1)
        blob = torch.from_numpy(np.frombuffer(b"".join(m for m, _ in members), dtype=np.uint8).copy())
        out_full = torch.zeros(int(sizes.sum()), dtype=torch.uint8)
    idx = shard.broadcast_index(index, 0)
    parts = shard.partition(idx[:, 1] + idx[:, 2], world)
    ec, eo = np.concatenate([[0], np.cumsum(idx[:, 1])]), np.concatenate([[0], np.cumsum(idx[:, 2])])
    rc_ = [(int(ec[a]), int(ec[b])) for a, b in parts]
    ro = [(int(eo[a]), int(eo[b])) for a, b in parts]
    local, works = shard.scatter_slab(blob, rc_, 0)
    for w in works:
        w.wait()
    a, b = parts[rank]
    mine = torch.zeros(ro[rank][1] - ro[rank][0], dtype=torch.uint8)
    lb = local.numpy().tobytes()
    for k in range(a, b):
        body = lb[idx[k, 0] - ec[a] + 10: idx[k, 0] - ec[a] + idx[k, 1] - 8]          # (no optional header fields)
        st, cons, prod, out = oracle.inflate_ns(body, int(idx[k, 2]))
        assert st == 0 and prod == idx[k, 2]
        mine[eo[k] - eo[a]: eo[k + 1] - eo[a]] = torch.from_numpy(np.frombuffer(out, dtype=np.uint8).copy())
    half = (ro[1][1] - ro[1][0]) // 2
    pieces = [[(0, ro[0][1] - ro[0][0])], [(0, half), (half, ro[1][1] - ro[1][0])]]
    for w in shard.gather_slab(mine, out_full, ro, 0, pieces=pieces):
        w.wait()
    dist.barrier()
    if rank == 0:
        q.put((out_full.numpy().tobytes() == b"".join(p for _, p in members), [list(p) for p in parts]))
    dist.destroy_process_group()


def test_two_ranks_gloo_scatter_decode_gather():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_c4, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    same, parts = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same and parts[0][1] == parts[1][0] and parts[0][0] == 0 and parts[0][1] > 0
