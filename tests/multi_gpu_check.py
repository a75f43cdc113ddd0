"""Multi-GPU self-check, run on a GPU box (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py

Every rank traces its contiguous shard of the Michelson gausslets on its own GPU, captures the
output port on the device, and the detector field of the sharded gausslets is reduced with one
NCCL all-reduce (distributed.field_sharded).  Rank 0 repeats everything on one GPU and compares.
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
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
