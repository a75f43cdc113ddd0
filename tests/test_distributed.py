"""The N>1 path on CPU: two `gloo` ranks each trace a contiguous block of the source
(with the oracle standing in for the GPU tracer -- this tests the sharding / renumbering /
gather logic of raypier_optics_b200.distributed, not a kernel) and the stitched result
must equal the single-process trace of the whole source."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, kw, rl, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import raypier_optics_b200.core as core
    from oracle import oracle as O
    from raypier_optics_b200 import distributed as rd, scene as SC
    from util import build_case
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = build_case(core, name, kw, rl)
    sc = SC.Scene(cfg['face_lists'], cfg['wavelengths'])

    def trace_fn(block):
        return O.trace_rays(sc, block, cfg['recursion_limit'], cfg['max_length'])

    out = rd.trace_sharded(trace_fn, cfg['rays'], world, rank, gather_to=0)
    if rank == 0:
        q.put(([g.tobytes() for g in out['gathered']], out['counts_all'].tolist(), out['face_counts'].tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,kw,rl", [("config2", dict(n=3001, reflection_threshold=1e-3,
                                                         transmission_threshold=1e-3), 5),
                                        ("config5", dict(n=501, gausslets=True), None)])
def test_two_rank_trace_equals_single_process(name, kw, rl):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import raypier_optics_b200.core as core
    from oracle import oracle as O
    from raypier_optics_b200 import scene as SC
    from util import build_case
    cfg = build_case(core, name, kw, rl)
    sc = SC.Scene(cfg['face_lists'], cfg['wavelengths'])
    want, want_counts = O.trace_rays(sc, cfg['rays'], cfg['recursion_limit'], cfg['max_length'])

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, kw, rl, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, counts_all, face_counts = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert len(gathered) == len(want)
    assert np.sum(counts_all, axis=0).tolist() == [len(g) for g in want]
    for g, (got, w) in enumerate(zip(gathered, want)):
        assert got == w.tobytes(), "generation %d differs from the single-process trace" % g
    assert face_counts == want_counts.tolist()


def _field_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from oracle import oracle as O
    from raypier_optics_b200 import _abi as A, distributed as rd
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = np.load(os.path.join(ROOT, "tests", "golden", "fields_michelson.npz"))
    g = z['gausslets'].view(A.gausslet_dtype).reshape(-1)
    mine = rd.shard_rays(g, world, rank)
    part = O.eval_Efield_from_gausslets(mine, z['points'], z['wavelengths'], float(z['blending']), float(z['time_ps']))
    total = rd.allreduce_field(part)
    if rank == 0:
        q.put(total.tobytes())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_field_sum_equals_single_process():
    """Detector field of ray shards: every rank sums ITS rays at all points, one all-reduce of
    the grid (raypier_optics_b200.distributed.allreduce_field) gives the full field."""
    import torch.multiprocessing as mp
    z = np.load(os.path.join(ROOT, "tests", "golden", "fields_michelson.npz"))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_field_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = np.frombuffer(q.get(timeout=300), dtype=np.complex128).reshape(-1, 3)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = z['E']
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-13  # only the summation order differs


def _records_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from raypier_optics_b200 import _abi as A, distributed as rd
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank r holds 3 + 5 r "terminal rays" with a recognisable ident; rank 1 of the ragged case holds none
    for ragged in (False, True):
        n = 0 if (ragged and rank == 1) else 3 + 5 * rank
        rec = np.zeros(n, dtype=A.ray_dtype)
        rec['ray_ident'] = 1000 * rank + np.arange(n)
        rec['length'] = rank + 0.5
        local = torch.from_numpy(rec.view(np.uint8).reshape(-1).copy()) if n else torch.zeros(4, dtype=torch.uint8)
        everywhere, counts = rd.gather_records(local, n, A.ray_dtype.itemsize, dst=None)
        on0, counts0 = rd.gather_records(local, n, A.ray_dtype.itemsize, dst=0)
        assert counts == counts0
        assert (on0 is None) == (rank != 0)
        if rank == 0:
            assert torch.equal(on0, everywhere)
            q.put((ragged, counts, everywhere.numpy().tobytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_of_terminal_records():
    """distributed.gather_records (the collective behind gather_terminal): variable-length runs of
    packed ray records concatenated in rank order, an empty rank included."""
    import torch.multiprocessing as mp
    from raypier_optics_b200 import _abi as A
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_records_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for ragged, counts, raw in results:
        rec = np.frombuffer(raw, dtype=A.ray_dtype)
        assert counts == ([3, 0] if ragged else [3, 8])
        want_ident = list(range(3)) + ([] if ragged else [1000 + i for i in range(8)])
        assert rec['ray_ident'].tolist() == want_ident
        assert rec['length'].tolist() == [0.5] * 3 + ([] if ragged else [1.5] * 8)
