"""Multi-GPU sharding of a trace.

A ray never interacts with another ray, so the source rays are split into contiguous
blocks in index order, one block per rank (one process per GPU), scene tables replicated,
and every rank runs the whole generation loop on its block with NO collective inside the
loop.  Because children are emitted in parent order, the concatenation of the ranks'
generation g in rank order *is* the single-GPU generation g, up to the parent index:

    global_parent_idx = local_parent_idx + sum_{r' < rank} count_{g-1}(r')

The only collectives (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests)
run after the trace: an all-gather of the per-generation counts, an all-reduce of the
per-face hit counts, and an optional gather of generations to rank 0.
"""
import numpy as np


def shard_bounds(n, world_size, rank):
    """Contiguous block [lo, hi) of rank ``rank`` when ``n`` rays are split over ranks."""
    return (rank * n) // world_size, ((rank + 1) * n) // world_size


def shard_rays(rays, world_size, rank):
    lo, hi = shard_bounds(rays.shape[0], world_size, rank)
    return np.ascontiguousarray(rays[lo:hi])


def _dist():
    import torch.distributed as dist
    return dist


def exchange_counts(counts, face_counts, device=None, group=None):
    """All-gather the per-generation ray counts and all-reduce the per-face hit counts.

    Returns (counts_all, face_counts_total): counts_all[r][g] (int64, zero padded to the
    longest trace) and the summed Face.count array.
    """
    import torch
    dist = _dist()
    world = dist.get_world_size(group)
    n_gen = torch.tensor([len(counts)], dtype=torch.int64, device=device)
    dist.all_reduce(n_gen, op=dist.ReduceOp.MAX, group=group)
    g_max = int(n_gen.item())
    local = torch.zeros(max(g_max, 1), dtype=torch.int64, device=device)
    if counts:
        local[:len(counts)] = torch.tensor(list(counts), dtype=torch.int64, device=device)
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local, group=group)
    fc = torch.as_tensor(np.asarray(face_counts, dtype=np.int64), device=device).clone()
    dist.all_reduce(fc, op=dist.ReduceOp.SUM, group=group)
    counts_all = np.stack([g.cpu().numpy() for g in gathered])[:, :g_max]
    return counts_all, fc.cpu().numpy()


def parent_offsets(counts_all, rank):
    """offset[g] = number of rays of generation g held by lower ranks."""
    return counts_all[:rank].sum(axis=0) if rank > 0 else np.zeros(counts_all.shape[1], dtype=np.int64)


def globalize_generation(arr, g, offsets):
    """Rewrite ``parent_idx`` of local generation ``g`` (g >= 1) into the global numbering."""
    if g == 0 or arr.shape[0] == 0:
        return arr
    base = arr['base_ray'] if 'base_ray' in (arr.dtype.names or ()) else arr
    base['parent_idx'] = (base['parent_idx'].astype(np.int64) + int(offsets[g - 1])).astype(np.uint32)
    return arr


def gather_records(local, n_local, itemsize, dst=None, group=None):
    """Gather variable-length runs of packed records that already live in torch tensors ON THE
    COLLECTIVE'S DEVICE (CUDA tensors under NCCL: the transfer is GPU to GPU over NVLink / NVSwitch with
    no host bounce; CPU tensors under gloo in the tests).  ``local``: uint8 tensor holding this rank's
    ``n_local`` records of ``itemsize`` bytes.  Returns ``(records, counts)``: the ranks' records
    concatenated in rank order as one uint8 tensor on the same device -- on every rank when ``dst`` is
    None (all-gather), else on rank ``dst`` only (None elsewhere) -- and the per-rank record counts."""
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = local.device
    cnt = torch.tensor([int(n_local)], dtype=torch.int64, device=dev)
    all_cnt = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(all_cnt, cnt, group=group)
    counts = [int(c.item()) for c in all_cnt]
    nmax = max(counts) if counts else 0
    pad = torch.zeros(max(nmax * itemsize, 1), dtype=torch.uint8, device=dev)
    pad[:int(n_local) * itemsize] = local[:int(n_local) * itemsize]
    if dst is None:
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
    elif rank == dst:
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.gather(pad, parts, dst=dst, group=group)
    else:
        dist.gather(pad, None, dst=dst, group=group)
        return None, counts
    out = torch.cat([p[:c * itemsize] for p, c in zip(parts, counts)]) if sum(counts) else pad[:0]
    return out, counts


def gather_terminal(engine, dev_rays, dst=None, group=None):
    """The end-of-trace gather of SURVEY 8e, device resident: this rank's terminal rays (a DeviceRays, e.g.
    from ``Engine.select_terminal`` or ``trace_consume(..., terminal=True)``) are written as packed
    ray_t / gausslet_t records into a CUDA tensor (``rpx_rays_export_device``), gathered with NCCL, and
    the concatenation is handed back to the library as a device collection (``rpx_rays_import_device``).
    Returns ``(DeviceRays or None, counts per rank)``.  ``parent_idx`` keeps each rank's numbering."""
    import torch
    from . import _abi as A
    dev = torch.device("cuda", engine.device)
    itemsize = A.gausslet_dtype.itemsize if dev_rays.is_gausslet else A.ray_dtype.itemsize
    n = len(dev_rays)
    send = torch.empty(max(n * itemsize, 4), dtype=torch.uint8, device=dev)
    engine.export_device(dev_rays, send.data_ptr(), n)  # returns after the copy finished on the library's stream
    out, counts = gather_records(send, n, itemsize, dst=dst, group=group)
    if out is None:
        return None, counts
    torch.cuda.current_stream(dev).synchronize()  # the library imports on its own (non-blocking) stream
    total = sum(counts)
    return engine.import_device(out.data_ptr() if total else send.data_ptr(), total, dev_rays.is_gausslet), counts


def gather_generation(arr, counts_g, dst=0, device=None, group=None):
    """Gather one (globalized) generation to ``dst`` in rank order, HOST arrays in and out (the
    reference-facing convenience for small traces; large traces gather terminal rays on the device with
    ``gather_terminal``).  ``counts_g[r]`` is the size of rank r's part.  Returns the concatenated array
    on ``dst`` and None elsewhere."""
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    itemsize = arr.dtype.itemsize
    nmax = int(max(counts_g)) if len(counts_g) else 0
    buf = torch.zeros(max(nmax * itemsize, 1), dtype=torch.uint8, device=device)
    raw = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1).copy())
    buf[:raw.numel()] = raw.to(buf.device)
    if rank == dst:
        parts = [torch.zeros_like(buf) for _ in range(world)]
        dist.gather(buf, parts, dst=dst, group=group)
        out = [p.cpu().numpy()[:int(c) * itemsize].view(arr.dtype) for p, c in zip(parts, counts_g)]
        return np.concatenate(out) if out else arr[:0]
    dist.gather(buf, None, dst=dst, group=group)
    return None


def trace_sharded(trace_fn, rays, world_size, rank, device=None, group=None, gather_to=None):
    """Trace this rank's contiguous block of ``rays`` with ``trace_fn(block) ->
    (list_of_generation_arrays, face_counts)`` and stitch the global numbering.

    Returns dict(local=list of globalized generations, counts_all, face_counts, gathered)
    where ``gathered`` is the full trace on rank ``gather_to`` (if requested)."""
    block = shard_rays(rays, world_size, rank)
    gens, face_counts = trace_fn(block)
    counts_all, fc = exchange_counts([len(g) for g in gens], face_counts, device=device, group=group)
    offs = parent_offsets(counts_all, rank)
    n_gen = counts_all.shape[1]
    dtype = rays.dtype
    local = []
    for g in range(n_gen):
        arr = gens[g] if g < len(gens) else np.zeros(0, dtype=dtype)
        local.append(globalize_generation(arr, g, offs))
    gathered = None
    if gather_to is not None:
        gathered = []
        for g in range(n_gen):
            full = gather_generation(local[g], counts_all[:, g], dst=gather_to, device=device, group=group)
            gathered.append(full)
        if rank != gather_to:
            gathered = None
    return dict(local=local, counts_all=counts_all, face_counts=fc, gathered=gathered)


# ---- detector field: the one reduction of the pipeline -------------------------
def allreduce_field(E, device=None, group=None):
    """Sum the partial E-fields of the ranks.  Every rank evaluates the modes of ITS rays at ALL
    detector points (the sum over rays is associative), so the only exchange is one all-reduce
    of the npt x 3 complex grid.  ``E`` may be a numpy complex128 array (copied through a tensor
    on ``device``; gloo in the CPU tests) or a float64 torch tensor of shape (npt, 6) already on
    the GPU (reduced in place by NCCL over NVLink -- no host round trip)."""
    import torch
    dist = _dist()
    if isinstance(E, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(E).view(np.float64).copy())
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t.cpu().numpy().view(np.complex128).reshape(E.shape)
    dist.all_reduce(E, op=dist.ReduceOp.SUM, group=group)
    return E


def field_sharded(engine, gausslets_shard, wavelengths, points, blending=1.0, time_ps=0.0, group=None):
    """E-field of gausslets that are sharded over the ranks: mode fit + summation of the local
    shard on this rank's GPU straight into a torch tensor, then one NCCL all-reduce of the grid.
    Returns the (npt, 3) complex128 field, identical on every rank."""
    import torch
    dev = torch.device("cuda", engine.device)
    pts = torch.from_numpy(np.ascontiguousarray(points, dtype=np.double).reshape(-1, 3)).to(dev)
    out = torch.zeros((pts.shape[0], 6), dtype=torch.float64, device=dev)
    # the zero fill and the points upload run on torch's stream; the library accumulates on its own
    # non-blocking stream (rpx_stream): order the two before handing the pointers over
    torch.cuda.current_stream(dev).synchronize()
    if len(gausslets_shard):
        fm = engine.field_prepare(gausslets_shard, wavelengths, blending=blending)
        try:
            fm.evaluate_device(pts.data_ptr(), pts.shape[0], out.data_ptr(), time_ps)
        finally:
            fm.free()
    allreduce_field(out, group=group)
    return out.cpu().numpy().view(np.complex128).reshape(-1, 3)


class _DeviceBuffer(object):
    """Zero-copy view of library-owned device memory for torch (``__cuda_array_interface__`` v3)."""

    def __init__(self, ptr, n_doubles):
        self.__cuda_array_interface__ = {"shape": (int(n_doubles),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def allreduce_detector(engine, detector, group=None):
    """Sum the partial detector fields of the ranks IN PLACE in the library's own device buffer: one NCCL
    all-reduce of npt x 6 doubles over NVLink, no copy and no host round trip ("probe/detector
    results at the end of a trace").  The detector's summation kernels run on the library's stream; they
    are waited for first (``Detector.ms`` synchronises on their events)."""
    import torch
    dist = _dist()
    _ = detector.ms  # waits for every accumulate enqueued so far
    dev = torch.device("cuda", engine.device)
    t = torch.as_tensor(_DeviceBuffer(detector.field_device_ptr, max(detector.npt, 1) * 6), device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    torch.cuda.current_stream(dev).synchronize()  # the library reads the buffer on its own stream next
    return t


def bind_to_gpu_numa(device_index):
    """Pin this process (and, by first-touch, the pinned staging memory it allocates next) to the CPU cores
    of the NUMA node the GPU hangs off: with one rank per GPU streaming its source through pinned host
    memory, a rank on the wrong socket moves every byte across the inter-socket link.  Reads
    /sys/bus/pci/devices/<bus id>/local_cpulist; returns the core list, or None when the topology is not
    visible (containers often hide it) or has a single node -- never an error."""
    import os
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device_index), "pci_device_id", 0)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (dom, bus, dev)
        spec = open(path).read().strip()
        cores = []
        for part in spec.split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cores.extend(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if not allowed or len(allowed) == len(os.sched_getaffinity(0)):
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None
