#!/usr/bin/env python3
"""h2d_probe.py -- the host-to-device copy ceiling of this box for the bench's end-to-end path.

Plain cudaMemcpyAsync from pinned host memory to device memory (torch's non_blocking copy_ of a pinned tensor is exactly that),
same size as one step's input of bench.py (256 frames x 640 x 480 bytes = 78.6 MB per GPU), on N GPUs at once:

    python tools/h2d_probe.py                      # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_probe.py

Every rank times K back-to-back copies with CUDA events between two barriers; the line printed by rank 0 carries the per-rank
rates and the aggregate = total bytes / slowest rank's time. bench.py calls measure() itself after its end-to-end region and
reports e2e.frac_of_copy_ceiling = (e2e frames/s x input bytes per frame) / that aggregate.
"""
import json
import os
import sys


def measure(torch, dev, nbytes, reps=20, dist=None, also_d2h_bytes=0):
    """Returns (seconds for `reps` copies on this rank, gathered list over ranks or None). Call on every rank at the same time."""
    src = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    src.fill_(7)
    dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda:%d" % dev)
    back_d = back_h = None
    if also_d2h_bytes:
        back_d = torch.empty(also_d2h_bytes, dtype=torch.uint8, device="cuda:%d" % dev)
        back_h = torch.empty(also_d2h_bytes, dtype=torch.uint8, pin_memory=True)
    up, down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    for _ in range(3):
        with torch.cuda.stream(up):
            dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(dev)
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(up)
    for _ in range(reps):
        with torch.cuda.stream(up):
            dst.copy_(src, non_blocking=True)
        if back_d is not None:   # the result copies of the real path run the other way at the same time
            with torch.cuda.stream(down):
                back_h.copy_(back_d, non_blocking=True)
    e1.record(up)
    torch.cuda.synchronize(dev)
    secs = e0.elapsed_time(e1) * 1e-3
    allsecs = None
    if dist is not None:
        t = torch.tensor([secs], dtype=torch.float64)
        out = [torch.zeros(1, dtype=torch.float64) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t)
        allsecs = [float(x.item()) for x in out]
        dist.barrier()
    return secs, allsecs


def main():
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import torch
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    dev = local % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    nbytes, reps = 256 * 640 * 480, 20
    out = {"bytes_per_copy": nbytes, "reps": reps, "n_gpus": world}
    for label, d2h in (("h2d_only", 0), ("h2d_with_d2h_16MB", 16 * 1024 * 1024)):
        secs, allsecs = measure(torch, dev, nbytes, reps, dist, also_d2h_bytes=d2h)
        per = [nbytes * reps / s / 1e9 for s in (allsecs or [secs])]
        out[label] = {"per_gpu_GBps": per, "aggregate_GBps": world * nbytes * reps / max(allsecs or [secs]) / 1e9}
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
