"""Multi-GPU self-check, run on a GPU box (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py

Every rank traces its contiguous shard of the Michelson gausslets on its own GPU, captures the
output port on the device, and the detector field of the sharded gausslets is reduced with one
NCCL all-reduce (distributed.field_sharded).  Rank 0 repeats everything on one GPU and compares.
Then the round-2 paths: the streamed trace with device-side consumers (rpx_trace_consume) feeding a
detector that is all-reduced IN PLACE in the library's device buffer (distributed.allreduce_detector),
and the device-resident NCCL gather of the terminal rays (distributed.gather_terminal) with
its measured rate.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import raypier_optics_b200.core as core  # noqa: E402
from raypier_optics_b200 import configs, distributed as rd, scene  # noqa: E402
from raypier_optics_b200.engine import Engine  # noqa: E402


def capture_plane():
    face = core.cfaces.RectangularFace(length=12.0, width=12.0, offset=0.0, z_plane=0.0)
    fl = core.ctracer.FaceList(owner=configs.Pose(centre=(0.0, -12.0, 0.0), direction=(0.0, 1.0, 0.0)))
    fl.faces = [face]
    fl.sync_transforms()
    return scene.Scene([fl], np.asarray([1.0]))


def traced_output(eng, cfg, rays):
    res = eng.trace(np.ascontiguousarray(rays), cfg["max_length"], cfg["recursion_limit"])
    g, _, counts = res.capture(cfg["wavelengths"])
    segs, fc = res.segments, res.face_counts.copy()
    res.free()
    return g, segs, fc


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = configs.build(core, "config5", n=40000, gausslets=True, seed=9)
    eng = Engine(local)
    eng.set_scene(scene.Scene(cfg["face_lists"], cfg["wavelengths"]))
    eng.set_capture_scene(capture_plane())
    mine = rd.shard_rays(cfg["rays"], world, rank)
    g, segs, fc = traced_output(eng, cfg, mine)
    xs = np.linspace(-3.0, 3.0, 96)
    gx, gz = np.meshgrid(xs, xs)
    pts = np.stack([gx.ravel(), np.full(gx.size, -14.0), gz.ravel()], axis=1)
    E = rd.field_sharded(eng, g, cfg["wavelengths"], pts)
    t = torch.tensor([segs, len(g)], dtype=torch.int64, device="cuda")
    dist.all_reduce(t)
    fct = torch.from_numpy(fc.astype(np.int64)).cuda()
    dist.all_reduce(fct)
    ok = True
    if rank == 0:
        g1, segs1, fc1 = traced_output(eng, cfg, cfg["rays"])
        fm = eng.field_prepare(g1, cfg["wavelengths"])
        E1 = fm.evaluate(pts)
        fm.free()
        err = float(np.abs(E - E1).max() / np.abs(E1).max())
        ok = (int(t[0]) == segs1 and int(t[1]) == len(g1) and fct.cpu().numpy().tolist() == fc1.tolist()
              and err < 1e-12)
        print("multi_gpu_check: world=%d segments %d/%d captured %d/%d face counts %s field rel err %.2e -> %s"
              % (world, int(t[0]), segs1, int(t[1]), len(g1), fct.cpu().numpy().tolist() == fc1.tolist(), err,
                 "OK" if ok else "FAIL"))
    # ---- streamed trace + consumers, detector all-reduced in place (no host bounce)
    det = eng.detector(pts[:1024], cfg["wavelengths"])
    out = eng.trace_consume(np.ascontiguousarray(mine), cfg["max_length"], cfg["recursion_limit"], chunk_rays=4096,
                            terminal=True, terminal_capacity=8 * len(mine) + 64, capture=True, detector=det)
    rd.allreduce_detector(eng, det)
    E2 = det.read()
    tt = torch.tensor([out.segments, out.n_captured, out.n_terminal], dtype=torch.int64, device="cuda")
    dist.all_reduce(tt)
    # ---- device-resident gather of the terminal rays (all ranks get all of them); the second call is timed
    import time
    warm, _ = rd.gather_terminal(eng, out.terminal, dst=None)
    warm.free()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    allterm, counts = rd.gather_terminal(eng, out.terminal, dst=None)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n_term_all = len(allterm)
    idents = eng.download(allterm)['base_ray']['ray_ident']
    allterm.free()
    out.free()
    det.free()
    if rank == 0:
        err2 = float(np.abs(E2 - E1[:1024]).max() / np.abs(E1).max())
        res1 = eng.trace(np.ascontiguousarray(cfg["rays"]), cfg["max_length"], cfg["recursion_limit"])
        term1, tcounts = eng.select_terminal(res1.device_generations(), True)
        want_idents = np.sort(eng.download(term1)['base_ray']['ray_ident'])
        term1.free()
        res1.free()
        ok2 = (int(tt[0]) == segs1 and int(tt[1]) == len(g1) and int(tt[2]) == n_term_all == len(want_idents)
               and sum(counts) == n_term_all and np.array_equal(np.sort(idents), want_idents) and err2 < 1e-12)
        ok = ok and ok2
        print("multi_gpu_check: consume path segments %d captured %d terminal %d (per rank %s), detector all-reduced "
              "in place rel err %.2e, terminal gather %.1f MB in %.2f ms = %.1f GB/s into every rank -> %s"
              % (int(tt[0]), int(tt[1]), int(tt[2]), counts, err2, n_term_all * 668 / 1e6, dt * 1e3,
                 n_term_all * 668 / dt / 1e9, "OK" if ok2 else "FAIL"))
    # ---- the same gather at a size where the transfer dominates: 5e5 gausslets per rank -> 2e6 terminal rays each
    # (four per source gausslet: both output ports of both arms)
    big = configs.build(core, "config5", n=500000, gausslets=True, seed=11 + rank)
    outb = eng.trace_consume(np.ascontiguousarray(big["rays"]), big["max_length"], big["recursion_limit"], terminal=True,
                             terminal_capacity=4 * len(big["rays"]) + 64)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    allb, countsb = rd.gather_terminal(eng, outb.terminal, dst=None)
    torch.cuda.synchronize()
    dtb = time.perf_counter() - t0
    nb = len(allb)
    allb.free()
    outb.free()
    if rank == 0:
        okb = nb == sum(countsb) == world * 4 * len(big["rays"])
        ok = ok and okb
        print("multi_gpu_check: terminal gather of %d gausslets (%.0f MB into every rank, %d ranks) in %.1f ms = %.1f GB/s "
              "received per rank (export + NCCL all-gather + import) -> %s"
              % (nb, nb * 668 / 1e6, world, dtb * 1e3, nb * 668 / dtb / 1e9, "OK" if okb else "FAIL"))
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
