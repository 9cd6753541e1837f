#!/usr/bin/env python3
"""bench.py -- frames/s of ORB extract + dynamic-keypoint filter on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[1]): 256 synthetic 640x480 frames per GPU, nFeatures=1000, 8 levels x1.2,
FAST 20/7 (30/10 on the area_flag path), YOLO person boxes + moving points injected per frame; frames are
sharded across ranks with no collective (weak scaling). A "step" is one pass of the hot path over the
rank's 256-frame shard.

  python bench.py [--gpus N] [--steps K] [--warmup W]           # CUDA arm (default N=1)
  python bench.py --impl reference ...                           # the reference's own CPU extractor (oracle/_ref), all host threads

Prints ONE JSON line on rank 0. torch is used only for device buffers, the timing events on the launch
stream and the (gloo) barrier / max-over-ranks; the hot path is libcoeb_frontend.so.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

# one rank per GPU and two lane threads per rank: library thread pools (OpenMP / MKL behind numpy and torch) would oversubscribe the
# host cores of a many-GPU box, so they are held to one thread before anything imports them
for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
    os.environ.setdefault(_v, "1")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(ROOT, "coeb-slam_b200", "python"), os.path.join(ROOT, "tools")]

import numpy as np  # noqa: E402

METRIC = "frames/s ORB extract+dyn-filter (640x480,1k kps)"
W, H, NFEAT, NLEVELS = 640, 480, 1000, 8
FRAMES_PER_GPU = 256
LEVELS = [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231), (257, 193), (214, 161), (179, 134)]
SIGMA_P = sum(w * h for w, h in LEVELS)


def stage_bytes(n_kp, n_cand):
    """ALGORITHMIC bytes per frame of each stage."""
    p0, p7 = LEVELS[0][0] * LEVELS[0][1], LEVELS[-1][0] * LEVELS[-1][1]
    return {
        "classify": 0.0,
        "pyramid": float((SIGMA_P - p7) + (SIGMA_P - p0)),      # read resize sources + write L1..L7
        "blur": float(2 * SIGMA_P),                               # read + write every level
        "fast": float(SIGMA_P + 4 * n_cand),                      # read every level + candidate words out
        "select": float(4 * n_cand + 16 * n_kp),                  # candidates in + level keys out
        "describe": float(n_kp * (749 + 512 + 60 + 16)),          # IC patches + pattern samples + 28 B kp + 32 B desc (+ keys in)
    }


STAGE_KERNELS = {"classify": ["classify_kernel"], "pyramid": ["resize_kernel"], "blur": ["blur_kernel"],
                 "fast": ["fast_kernel", "fast_fallback_kernel"], "select": ["select_kernel"], "describe": ["describe_tma_kernel"]}


def newest_traffic_capture():
    """Path of the newest committed `ncu --set full` traffic summary (profiles/rNNx_traffic.json: round number, then letter)."""
    import glob
    import re
    best = None
    for f in glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")):
        m = re.match(r"r(\d+)([a-z]*)_traffic\.json$", os.path.basename(f))
        if m and (best is None or (int(m.group(1)), m.group(2)) > best[0]):
            best = ((int(m.group(1)), m.group(2)), f)
    return best[1] if best else None


def profiled_traffic(stage):
    """dram__bytes_read.sum + dram__bytes_write.sum per step of the stage's kernels, from the newest committed `ncu --set full`
    capture of this same command; (None, None, None) if there is none. Kernels a later build no longer launches are skipped."""
    try:
        path = newest_traffic_capture()
        k = json.load(open(path))["kernels"]
        names = [n for n in STAGE_KERNELS[stage] if n in k]
        if not names:
            return None, None, None
        return (float(sum(k[n]["dram_bytes"] for n in names)), {n: k[n]["alu_pipe_pct"] for n in names},
                "profiles/" + os.path.basename(path) + " (ncu --set full, per 256-frame step)")
    except Exception:
        return None, None, None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled while the timed region runs."""

    def __init__(self, gpu_index):
        self.idx, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile(prefix="clocks", suffix=".csv", delete=False)
            self.path = f.name
            q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap,timestamp")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=f, stderr=subprocess.DEVNULL)
            # nvidia-smi needs a moment to come up (longer on an 8-GPU box): wait for its first line so that it is
            # already looping when the short timed region starts
            t = time.time()
            while time.time() - t < 5.0 and os.path.getsize(self.path) == 0:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def stop(self, t_start=None, t_end=None):
        """Samples whose timestamp lies inside [t_start, t_end] (the timed region); if the region was shorter than the
        sampling period and caught none, the samples taken since the warm-up (same load) are used and `window` says so."""
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [[c.strip() for c in r.split(",")] for r in open(self.path).read().strip().splitlines() if r.strip()]
            rows = [r for r in rows if len(r) >= 8]
            out["window"] = "timed region"
            if t_start is not None:
                import datetime
                def ts(r):
                    try:
                        return datetime.datetime.strptime(r[7], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    except Exception:
                        return None
                inside = [r for r in rows if ts(r) is not None and t_start <= ts(r) <= t_end]
                if inside:
                    rows = inside
                else:
                    under_load = [r for r in rows if ts(r) is not None and ts(r) >= t_start - 0.5]
                    rows = under_load or rows[-2:]
                    out["window"] = "no sample fell inside the %.0f ms timed region: nearest samples under the same load (warm-up / stage pass)" % (1e3 * (t_end - t_start))
            sm = [float(r[0]) for r in rows]
            out["sm_mhz"] = float(np.median(sm)) if sm else None
            out["sm_max_mhz"] = float(rows[0][1]) if rows else None
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            out["reasons"] = [n for i, n in enumerate(names) if any("Active" in r[3 + i] and "Not" not in r[3 + i] for r in rows)]
            out["samples"] = len(rows)
        except Exception:
            pass
        finally:
            try:
                os.unlink(self.path)
            except Exception:
                pass
        return out


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist_mod.init_process_group(backend="gloo", rank=rank, world_size=world)
        dist = dist_mod
    return rank, world, local, dist


def reduce_max(dist, value):
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(dist, value):
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def base_config(n_gpus):
    return {"workload": "configs[1]: batched ORB extract + dynamic filter, %d synthetic 640x480 frames per GPU, "
                        "nFeatures=1000, 8 levels x1.2, FAST 20/7, YOLO boxes + T_M per frame" % FRAMES_PER_GPU,
            "frames_per_gpu": FRAMES_PER_GPU, "width": W, "height": H, "nfeatures": NFEAT, "nlevels": NLEVELS,
            "sharding": "independent frame shards, no collective", "n_gpus": n_gpus,
            "l2": "per step the inputs + pyramid arenas (~0.6 GB per shard) exceed the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's OWN CPU implementation (oracle/_ref: src/ORBextractor.cc compiled unchanged against
# the test-only OpenCV shim), one extractor per host thread, all host threads. Falls back to the
# oracle port (kind "port") only if the prebuilt oracle/_ref library did not travel with the snapshot.
# ------------------------------------------------------------------------------------------------
def cpu_arm():
    """(module, kind, description) of the CPU implementation timed beside the GPU: oracle/_ref if present, else the port."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    try:
        import ref
        if ref.available():
            ref.lib("ref")
            return ref, orc, "reference", "oracle/_ref: reference src/ORBextractor.cc unchanged (-O3 -march=x86-64-v3), OpenCV primitives from the cv2-pinned shim"
    except Exception as e:   # noqa: BLE001 -- a missing prebuilt library must not kill the bench line
        sys.stderr.write("bench.py: oracle/_ref unavailable (%s); timing the oracle port instead\n" % e)
    return orc, orc, "port", "oracle port (oracle/libcoeb_oracle.so)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    impl, orc, kind, what = cpu_arm()
    from coeb_b200 import synth
    threads = orc.hardware_threads()
    sample = int(args.ref_frames)
    # the first `sample` frames of the rank-0 shard of the CUDA arm (same seeds, same boxes / T_M / blur flags)
    batch = synth.make_batch(FRAMES_PER_GPU, base_seed=0, w=W, h=H, unique=args.unique)
    sub = {k: np.ascontiguousarray(batch[k][:sample]) for k in batch}
    params = orc.OrbParams(NFEAT, 1.2, NLEVELS, 20, 7)

    def step():
        secs, counts, _, _ = impl.extract_batch_mt(params, sub["gray"], sub["boxes"], sub["nbox"], sub["tm"], sub["ntm"], sub["blur"], threads)
        return secs, counts
    for _ in range(max(args.warmup, 1)):
        step()
    t = 0.0
    for _ in range(args.steps):
        s, counts = step()
        t += s
    fps = sample * args.steps / t
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": base_config(args.gpus),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                             "sample": "each step = the first %d frames of the rank-0 shard (of %d), one extractor per thread, %d threads; %s"
                                       % (sample, FRAMES_PER_GPU, threads, what)},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "mean_keypoints": float(np.mean(counts))}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(torch, dev):
    """Multi-GPU runs: keep this rank's threads (and so the first touch of its pinned staging buffers) on the NUMA node the GPU
    hangs off, so that 8 ranks do not pull their host->device traffic across the socket link. Returns the node or None."""
    try:
        p = torch.cuda.get_device_properties(dev)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def run_b200(args):
    import torch
    import coeb_b200 as cb
    from coeb_b200 import synth

    rank, world, local, dist = dist_setup(args.gpus)
    if not torch.cuda.is_available() or cb.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the CUDA arm has no CPU fallback (use --impl reference for the CPU arm)")
    dev = local % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    numa_node = bind_to_gpu_numa_node(torch, dev) if world > 1 else None
    torch.set_num_threads(1)
    # each rank keeps to its own slice of the host cores (after the NUMA restriction above), and inside the slice the two lane
    # threads of the end-to-end region get a core each: eight ranks on a 32-vCPU box otherwise migrate over each other
    my_cores = sorted(os.sched_getaffinity(0))
    if world > 1 and len(my_cores) >= 2 * world:
        ranks_here = [r for r in range(world)]   # one node: LOCAL_RANK == RANK
        k = len(my_cores) // len(ranks_here)
        my_cores = my_cores[local * k:(local + 1) * k]
        os.sched_setaffinity(0, set(my_cores))
    B = FRAMES_PER_GPU
    # seeds 0..255 are the frame ids within a shard, the rank is added x1000
    batch = synth.make_batch(B, base_seed=rank * 1000, w=W, h=H, unique=args.unique)
    ex = cb.Extractor(NFEAT, 1.2, NLEVELS, 20, 7, device=dev)
    cap = ex.default_cap()
    ex.reserve(W, H, B)
    stream = torch.cuda.Stream(device=dev)  # a real (non-default) stream: the library treats stream 0 as "use my own"
    torch.cuda.set_stream(stream)
    ex.set_stream(stream.cuda_stream)

    def to_dev(a):
        return torch.from_numpy(a).to("cuda:%d" % dev)
    d_gray, d_boxes, d_nbox = to_dev(batch["gray"]), to_dev(batch["boxes"]), to_dev(batch["nbox"])
    d_tm, d_ntm, d_blur = to_dev(batch["tm"]), to_dev(batch["ntm"]), to_dev(batch["blur"])
    d_kps = torch.empty((B, cap, 28), dtype=torch.uint8, device=d_gray.device)
    d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=d_gray.device)
    d_counts = torch.zeros(B, dtype=torch.int32, device=d_gray.device)
    d_status = torch.zeros(B, dtype=torch.int32, device=d_gray.device)

    def step_device():
        ex.extract_batch_device(B, d_gray.data_ptr(), W, H, W, W * H, d_boxes.data_ptr(), d_nbox.data_ptr(), synth.MAX_BOX,
                                d_tm.data_ptr(), d_ntm.data_ptr(), synth.MAX_TM, d_blur.data_ptr(), d_kps.data_ptr(),
                                d_desc.data_ptr(), d_counts.data_ptr(), d_status.data_ptr(), cap)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value"); per-stage times for the roofline come from a second pass ------------
    ex.set_profiling(bool(args.stage_sync))
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_device()
    barrier()
    t_clk0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_acc = {k: 0.0 for k in ex.STAGES}
    stage_events = []
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
        if args.stage_sync:  # read the stage events of this step (blocks on the step's end; off by default)
            for k, v in ex.stage_ms().items():
                stage_acc[k] += v
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    t_clk1 = time.time()
    launches_per_step = ex.launches_per_call()   # of the device-resident call just timed
    if rank == 0 and t_clk1 - t_clk0 < 0.06:   # a region shorter than a few sampling periods: keep the GPU under the same load a little longer
        for _ in range(args.steps):
            step_device()
        torch.cuda.synchronize()
    clocks = sampler.stop(t_clk0, t_clk1) if rank == 0 else None
    counts = d_counts.cpu().numpy()
    status = d_status.cpu().numpy()
    assert (status == 0).all(), "device reported per-frame failures: %s" % status[status != 0][:8]
    # stage breakdown: a second, identical pass with per-stage events (serialises blur with FAST) and a sync per step;
    # not part of `value`
    if not args.stage_sync:
        ex.set_profiling(True)
        for _ in range(args.steps):
            step_device()
            for k, v in ex.stage_ms().items():
                stage_acc[k] += v
    stage_ms = {k: v / args.steps for k, v in stage_acc.items()}
    ex.set_profiling(False)
    ms_total = reduce_max(dist, ms_total)
    frames_total = reduce_sum(dist, float(B * args.steps))
    value = frames_total / (ms_total * 1e-3)

    if args.match_only:   # profiling runs of the matcher kernels: one untimed host-path extraction for the inputs, then the matching section
        if rank == 0:
            kps_h, desc_h, cnt_h, st_h = ex.extract_batch_host(batch["gray"], batch["boxes"], batch["nbox"], batch["tm"], batch["ntm"], batch["blur"], cap=cap)
            print(json.dumps({"match": bench_matching(cb, dev, batch, kps_h, desc_h, cnt_h, ex, args), "note": "--match-only profiling run, not a bench line"}))
        return 0
    if args.skip_e2e:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": "frames/s",
                              "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
                              "stage_ms": stage_ms, "note": "--skip-e2e profiling run, not a bench line"}))
        return 0
    # ---- end to end through the host-buffer C ABI call (pinned host memory in, host results out) ----
    pin = {}
    owners = []
    for k in ("gray", "boxes", "nbox", "tm", "ntm", "blur"):
        arr, own = cb.pinned_array(batch[k].shape, batch[k].dtype)
        arr[...] = batch[k]
        pin[k] = arr
        owners.append(own)
    o_kps, p1 = cb.pinned_array((B, cap), cb.KP_DTYPE)
    o_desc, p2 = cb.pinned_array((B, cap, 32), np.uint8)
    o_cnt, p3 = cb.pinned_array((B,), np.int32)
    o_st, p4 = cb.pinned_array((B,), np.int32)
    owners += [p1, p2, p3, p4]

    # Several batches in flight (--e2e-lanes, default 3): one handle (own streams, arenas and staging) per host thread, so the H2D
    # copy of one batch overlaps the kernels and the D2H copy of the others (each call is blocking; ctypes drops the GIL). Two lanes
    # already run at the copy ceiling on a quiet host (172.0 k frames/s, three: 172.1 k, four: 169.7 k; 8 GPUs: 476.0 k / 475.1 k);
    # the third keeps the copy engine fed when a host thread is late (one run on a noisy box fell to 132 k with two).
    import threading
    lanes = [(ex, pin, (o_kps, o_desc, o_cnt, o_st))]
    n_lanes = max(2, int(args.e2e_lanes))
    for _ in range(n_lanes - 1):
        ex_b = cb.Extractor(NFEAT, 1.2, NLEVELS, 20, 7, device=dev)
        pin_b = {}
        for k in pin:
            arr, own = cb.pinned_array(pin[k].shape, pin[k].dtype)
            arr[...] = pin[k]
            pin_b[k] = arr
            owners.append(own)
        outs_b = []
        for shape, dt in (((B, cap), cb.KP_DTYPE), ((B, cap, 32), np.uint8), ((B,), np.int32), ((B,), np.int32)):
            arr, own = cb.pinned_array(shape, dt)
            outs_b.append(arr)
            owners.append(own)
        lanes.append((ex_b, pin_b, tuple(outs_b)))

    def step_host(lane):
        e, pn, out = lanes[lane]
        e.extract_batch_host(pn["gray"], pn["boxes"], pn["nbox"], pn["tm"], pn["ntm"], pn["blur"], cap=cap, out=out)

    def run_lane(lane, n):
        torch.cuda.set_device(dev)
        if len(my_cores) >= 3:
            try:
                os.sched_setaffinity(0, {my_cores[lane % len(my_cores)]})   # pid 0 = the calling thread
            except OSError:
                pass
        for _ in range(n):
            step_host(lane)

    for _ in range(max(1, min(args.warmup, 3))):
        for ln in range(n_lanes):
            step_host(ln)
    barrier()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(2, min(args.steps, args.e2e_steps))
    e2e_steps -= e2e_steps % n_lanes
    e2e_steps = max(e2e_steps, n_lanes)
    t0 = time.perf_counter()
    h0.record(stream)
    workers = [threading.Thread(target=run_lane, args=(i, e2e_steps // n_lanes)) for i in range(n_lanes)]
    for w in workers:
        w.start()
    for w in workers:
        w.join()
    h1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    e2e_ms = reduce_max(dist, h0.elapsed_time(h1))  # device time between the bracketing events, max over ranks
    e2e_wall_ms = reduce_max(dist, wall * 1e3)
    e2e_frames = reduce_sum(dist, float(B * e2e_steps))
    for _, _, ob in lanes[1:]:
        assert np.array_equal(o_cnt, counts) and np.array_equal(ob[2], counts), "host-path and device-path keypoint counts differ"
        assert o_kps.tobytes() == ob[0].tobytes() and o_desc.tobytes() == ob[1].tobytes(), "the in-flight lanes disagree"
    h2d = sum(int(pin[k].nbytes) for k in pin)
    d2h = int(o_kps.nbytes + o_desc.nbytes + o_cnt.nbytes + o_st.nbytes)
    # the copy ceiling of this box at this N: plain pinned cudaMemcpyAsync of one step's frames on every rank at once, with the
    # result copies running the other way (tools/h2d_probe.py)
    import h2d_probe
    barrier()
    probe_secs, probe_all = h2d_probe.measure(torch, dev, int(pin["gray"].nbytes), 10, dist, also_d2h_bytes=d2h)
    probe_gbs = world * int(pin["gray"].nbytes) * 10 / max(probe_all or [probe_secs]) / 1e9

    # ---- single-frame latency (the tracking thread consumes one frame at a time) -------------------
    lat = None
    if rank == 0:
        ex1 = cb.Extractor(NFEAT, 1.2, NLEVELS, 20, 7, device=dev)
        f = 0
        nb, nt = int(batch["nbox"][f]), int(batch["ntm"][f])
        args1 = (batch["gray"][f], batch["boxes"][f, :nb], batch["tm"][f, :nt], batch["blur"][f, :nb])
        for _ in range(5):
            ex1.extract(*args1)
        import ctypes as C
        L = cb.lib()
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        kbuf, dbuf, nout = np.empty(cap, cb.KP_DTYPE), np.empty((cap, 32), np.uint8), C.c_int()
        ts = []
        # the argument marshalling (numpy -> ctypes pointers, ~10 us of interpreter time) is done up front: the clock brackets the
        # blocking C call alone, which is what a C++ caller of ORBextractor::operator() pays
        pk, pd, pn, hx = P(kbuf), P(dbuf), C.byref(nout), ex1.h
        calls = []
        for f in range(16):
            nb, nt = int(batch["nbox"][f]), int(batch["ntm"][f])
            g_, b_, t_, f_ = batch["gray"][f], batch["boxes"][f], batch["tm"][f], batch["blur"][f]
            calls.append((hx, P(g_), W, H, W, P(b_), nb, P(t_), nt, P(f_), nb, pk, pd, cap, pn))
        fn = L.coeb_extract
        for i in range(200):
            a = calls[i % 16]
            t = time.perf_counter()   # the blocking C-ABI call itself (ORBextractor::operator() equivalent)
            fn(*a)
            ts.append(time.perf_counter() - t)
        ex1.set_profiling(True)
        L.coeb_extract(ex1.h, P(g_), W, H, W, P(b_), nb, P(t_), nt, P(f_), nb, P(kbuf), P(dbuf), cap, C.byref(nout))
        st1 = {k: 1e3 * v for k, v in ex1.stage_ms().items()}
        lat = {"extract_filter_us_median": 1e6 * float(np.median(ts)), "extract_filter_us_p90": 1e6 * float(np.percentile(ts, 90)),
               "stage_us_device": st1,
               "note": "host buffers in, host results out, one 640x480 frame per call (wall clock around the blocking C call)"}
        ex1.close()

    # ---- matching: us per call through the C ABI (host arrays in, results out), CPU oracle beside it ----------------
    match = None
    if rank == 0 and not args.no_match:
        match = bench_matching(cb, dev, batch, o_kps, o_desc, o_cnt, ex, args)

    # ---- CPU baseline beside it (rank 0, N=1 only) --------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        impl, orc, cpu_kind, cpu_what = cpu_arm()
        threads = orc.hardware_threads()
        sample = int(args.cpu_frames)
        params = orc.OrbParams(NFEAT, 1.2, NLEVELS, 20, 7)
        sub = {k: np.ascontiguousarray(batch[k][:sample]) for k in batch}
        impl.extract_batch_mt(params, sub["gray"][:threads], sub["boxes"][:threads], sub["nbox"][:threads], sub["tm"][:threads],
                              sub["ntm"][:threads], sub["blur"][:threads], threads)   # warm-up: first touch of every thread's buffers
        secs, _, _, _ = impl.extract_batch_mt(params, sub["gray"], sub["boxes"], sub["nbox"], sub["tm"], sub["ntm"], sub["blur"], threads)

        # the CPU sample doubles as a parity check of the timed GPU run: against the oracle port and, when it travelled with the
        # snapshot, against the monotonic-heap build of the reference itself (the plain build's octree tie order depends on glibc's
        # allocation history, tests/test_ref_parity_cpu.py)
        def same_as_gpu(mod, **kw):
            _, cc, ck, cd = mod.extract_batch_mt(params, sub["gray"], sub["boxes"], sub["nbox"], sub["tm"], sub["ntm"], sub["blur"], threads,
                                                 cap=cap, want_outputs=True, **kw)
            return bool(all(cc[i] == o_cnt[i] and ck[i, :cc[i]].tobytes() == o_kps[i, :o_cnt[i]].tobytes()
                            and cd[i, :cc[i]].tobytes() == o_desc[i, :o_cnt[i]].tobytes() for i in range(sample)))
        same = same_as_gpu(orc)
        same_ref = None
        if cpu_kind == "reference":
            try:
                same_ref = same_as_gpu(impl, variant="mono")
            except Exception:
                same_ref = None
        # single-thread per-frame latency with the per-stage breakdown: 4 warm-up frames, then 24 frames
        oex = orc.Extractor(NFEAT, 1.2, NLEVELS, 20, 7)
        st_lat = []
        for i in range(28):
            nb, nt = int(batch["nbox"][i]), int(batch["ntm"][i])
            t = time.perf_counter()
            oex.extract(batch["gray"][i], batch["boxes"][i, :nb], batch["tm"][i, :nt], batch["blur"][i, :nb])
            if i >= 4:
                st_lat.append(time.perf_counter() - t)
        stage_t, nfr = oex.stage_times()
        # secondary, OpenCV-backed number: the cv2 calls the reference makes per frame (resize chain, one cv::FAST per 30 px cell with
        # the minTh retry, one GaussianBlur per level) driven like src/ORBextractor.cc drives them, single thread; its octree,
        # orientation and descriptor loops are the reference's own C++ and are not included, so this is a lower bound of its CPU path
        cv_prims_ms = None
        try:
            import cv2
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from oracle_cv2 import ExtractorA
            cv2.setNumThreads(1)
            xa = ExtractorA(NFEAT, 1.2, NLEVELS)
            det_ini, det_min = cv2.FastFeatureDetector_create(20, True), cv2.FastFeatureDetector_create(7, True)
            tt = []
            for i in range(10):
                t = time.perf_counter()
                for im in xa.pyramid(batch["gray"][i]):
                    hh, ww = im.shape
                    nC, nR = int((ww - 32 + 6) / 30), int((hh - 32 + 6) / 30)
                    wC, hC = -(-(ww - 32 + 6) // nC), -(-(hh - 32 + 6) // nR)
                    for ci in range(nR):
                        y0 = 16 + ci * hC
                        if y0 >= hh - 16 - 3:
                            continue
                        for cj in range(nC):
                            x0 = 16 + cj * wC
                            if x0 >= ww - 16 - 6:
                                continue
                            roi = im[y0:min(y0 + hC + 6, hh - 16), x0:min(x0 + wC + 6, ww - 16)]
                            if not det_ini.detect(roi):
                                det_min.detect(roi)
                    cv2.GaussianBlur(im, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
                if i >= 2:
                    tt.append(time.perf_counter() - t)
            cv_prims_ms = 1e3 * float(np.median(tt))
        except Exception:
            pass
        cpu = {"value": sample / secs, "unit": "frames/s", "cores": threads, "kind": cpu_kind,
               "sample": "first %d frames of the rank-0 shard, one extractor per thread, %d threads; %s" % (sample, threads, cpu_what),
               "bit_exact_vs_gpu": bool(same) if same_ref is None else bool(same and same_ref),
               "gpu_equals_oracle_port": bool(same), "gpu_equals_reference_monotonic_heap": same_ref,
               "single_thread_ms_per_frame_median": 1e3 * float(np.median(st_lat)),
               "single_thread_stage_ms_per_frame": {k: 1e3 * v / max(nfr, 1) for k, v in stage_t.items()},
               "opencv_primitives_only_ms_per_frame": cv_prims_ms}

    if rank != 0:
        return 0
    n_kp = float(np.mean(counts))
    # FAST candidates per frame, measured: the lists the timed run handed to the octree (frames 0..7 of the shard, all levels)
    n_cand = float(np.mean([sum(len(ex.level_candidates(l, frame=f)) for l in range(NLEVELS)) for f in range(8)]))
    sb = stage_bytes(n_kp, n_cand)
    dom = max(stage_ms, key=lambda k: stage_ms[k])
    peak, peak_kind = measured_peak_gbs()
    achieved = sb[dom] * B / (stage_ms[dom] * 1e-3) / 1e9
    traffic, alu_pct, traffic_src = profiled_traffic(dom)
    pipeline_bytes = 6049674.0
    per_stage = {}
    for st_name, ms in stage_ms.items():   # every stage against the same roofline: algorithmic GB/s and the DRAM GB/s ncu saw for its kernels
        t_bytes, t_alu, _ = profiled_traffic(st_name)
        per_stage[st_name] = {"algorithmic_gbs": sb.get(st_name, 0.0) * B / (ms * 1e-3) / 1e9 if ms > 0 else None,
                              "frac": sb.get(st_name, 0.0) * B / (ms * 1e-3) / 1e9 / peak if ms > 0 else None,
                              "dram_gbs_ncu": (t_bytes / (ms * 1e-3) / 1e9) if (t_bytes and ms > 0) else None, "alu_pipe_pct_ncu": t_alu}
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": base_config(world), "mean_keypoints": n_kp,
        "e2e": {"value": e2e_frames / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "wall_frames_per_s": e2e_frames / (e2e_wall_ms * 1e-3),
                "in_flight": "%d batches: one handle per host thread, blocking coeb_extract_batch_host calls, alternating steps" % n_lanes,
                "copy_ceiling_GBps": probe_gbs, "frac_of_copy_ceiling": (e2e_frames / (e2e_ms * 1e-3)) * (h2d / B) / 1e9 / probe_gbs,
                "copy_ceiling_note": "tools/h2d_probe.py run in-line: pinned cudaMemcpyAsync of one step's frames on all %d GPUs at once, "
                                     "D2H of the result size the other way" % world,
                "host_cores_of_rank0": len(my_cores),
                "numa_node_of_rank0": numa_node},
        "gpu_launches": args.steps * launches_per_step,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "candidates_per_frame_measured": n_cand,
                     "alu_pipe_pct_ncu": alu_pct,
                     "note": "the stage is integer-ALU bound (packed 16-bit min/max), not HBM bound: see DESIGN.md section 5",
                     "algorithmic_bytes_per_launch": sb[dom] * B,
                     "stage_ms": stage_ms, "stages": per_stage,
                     "pipeline": {"bytes_per_frame": pipeline_bytes, "achieved_gbs": pipeline_bytes * value / world / 1e9,
                                  "frac": pipeline_bytes * value / world / 1e9 / peak}},
        "cpu_baseline": cpu, "clocks": clocks, "latency": lat, "match": match,
    }
    print(json.dumps(line))
    return 0


def _median_us(fn, reps):
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return 1e6 * float(np.median(ts))


def bench_matching(cb, dev, batch, o_kps, o_desc, o_cnt, ex, args):
    """BASELINE.json configs[2] (1000 keypoints vs a 5k-point local map; frame-to-frame; initialisation) and configs[4]
    (4000 x 100k brute-force k=2). Wall time around the blocking C calls (host arrays in, host results out); the kNN is
    also timed device-resident with CUDA events. The CPU oracle is timed on the same inputs when available."""
    import torch
    from coeb_b200 import synth
    out = {}
    # configs[2] is quoted on ~1000 keypoints: take the frames of the shard that carry no person box (nothing culled)
    nobox = [i for i in range(len(o_cnt)) if int(batch["nbox"][i]) == 0]
    fa, fb = nobox[0], nobox[1]
    kps = np.ascontiguousarray(o_kps[fa, :o_cnt[fa]])
    desc = np.ascontiguousarray(o_desc[fa, :o_cnt[fa]])
    scale = ex.tables()["scale"]
    cam_args = (535.4, 539.2, 320.1, 247.6, 40.0, 40.0 / 535.4, 0.0, 640.0, 0.0, 480.0)
    m = cb.Matcher(device=dev)
    mp, uright = synth.make_map_points(kps, desc, scale, seed=1)
    last, Tc, Tl = synth.make_last_frame(kps, desc, seed=2)
    state = np.full(len(kps), -1, np.int32)
    f = m.frame(kps, desc, cb.Camera(*cam_args), scale, uright)
    out["n_keypoints"] = int(len(kps))
    out["frame_upload_grid_us"] = _median_us(lambda: m.frame(kps, desc, cb.Camera(*cam_args), scale, uright).close(), 20)
    # pre-marshalled ctypes calls: the timed region is the blocking C-ABI call itself, not numpy conversions
    import ctypes as C
    L = cb.lib()
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    nm = C.c_int()
    a2 = [u8(mp["track_in_view"]), u8(mp["bad"]), u8(mp["has_obs"]), f32(mp["proj_x"]), f32(mp["proj_y"]), f32(mp["proj_xr"]),
          i32(mp["level"]), f32(mp["view_cos"]), u8(mp["desc"])]
    km = state.copy()

    # (the numpy -> ctypes pointer objects are built once: ~1.2 us each of interpreter time that a C++ caller does not pay)
    pa2, pkm, pnm = [P(a) for a in a2], P(km), C.byref(nm)
    c30, c08, c15, c05, c09 = C.c_float(3.0), C.c_float(0.8), C.c_float(15.0), C.c_float(0.5), C.c_float(0.9)

    def call_m2():
        km[:] = state
        L.coeb_match_projection(m.h, f.h, len(a2[3]), *pa2, c30, c08, pkm, pnm)
    a3 = [u8(last["valid"]), u8(last["has_obs"]), f32(last["xyz"]), i32(last["octave"]), f32(last["angle"]), u8(last["desc"])]
    tc, tl = f32(Tc).reshape(12), f32(Tl).reshape(12)

    pa3, ptc, ptl = [P(a) for a in a3], P(tc), P(tl)

    def call_m3():
        km[:] = state
        L.coeb_match_lastframe(m.h, f.h, len(a3[0]), *pa3, ptc, ptl, c15, 0, 1, pkm, pnm)
    out["search_by_projection_map5k_us"] = _median_us(call_m2, 50)
    out["search_by_projection_lastframe_us"] = _median_us(call_m3, 50)
    k2 = np.ascontiguousarray(o_kps[fb, :o_cnt[fb]])
    d2 = np.ascontiguousarray(o_desc[fb, :o_cnt[fb]])
    f2 = m.frame(k2, d2, cb.Camera(*cam_args), scale, None)
    prev = np.stack([kps["x"], kps["y"]], axis=1).astype(np.float32)
    pv, m12 = prev.copy(), np.empty(len(kps), np.int32)

    ppv, pm12 = P(pv), P(m12)

    def call_m4():
        pv[:] = prev
        L.coeb_match_init(m.h, f.h, f2.h, ppv, pm12, 100, c09, 1, pnm)
    out["search_for_initialization_us"] = _median_us(call_m4, 50)
    # ---- tracking-thread chain: extract -> Frame tail on the device -> SearchLocalPoints on a resident 5k map ---------
    gray0 = np.ascontiguousarray(batch["gray"][fa])
    nb0, nt0 = int(batch["nbox"][fa]), int(batch["ntm"][fa])
    b0, t0, fl0 = batch["boxes"][fa], batch["tm"][fa], batch["blur"][fa]
    ex1 = cb.Extractor(NFEAT, 1.2, NLEVELS, 20, 7, device=dev)
    cap1 = ex1.default_cap()
    kb1, db1, n1 = np.empty(cap1, cb.KP_DTYPE), np.empty((cap1, 32), np.uint8), C.c_int()
    depth_pin, depth_owner = cb.pinned_array((H, W), np.uint16)
    depth_pin[:] = synth.make_depth(0, W, H)
    dimg = cb.DepthImage(depth_pin.ctypes.data, 2, depth_pin.strides[0], float(np.float32(1.0) / np.float32(5000.0)), 0, W, H)
    cam = cb.Camera(*cam_args)
    Tcw, Ow = synth.make_pose(0)
    lm, skip, has_obs = synth.make_local_map(kps, desc, scale, Tcw, seed=0)
    dev_map = m.local_map(lm)
    tcw, ow = f32(Tcw).reshape(12), f32(Ow).reshape(3)
    kun, ur1, dp1 = np.empty(cap1, cb.KP_DTYPE), np.empty(cap1, np.float32), np.empty(cap1, np.float32)
    inview = np.empty(dev_map.n, np.uint8)
    fh = C.c_void_p()
    nfr = C.c_int()

    b0, t0, fl0 = np.ascontiguousarray(b0), np.ascontiguousarray(t0), np.ascontiguousarray(fl0)
    xargs = (ex1.h, P(gray0), W, H, W, P(b0), nb0, P(t0), nt0, P(fl0), nb0, P(kb1), P(db1), cap1, C.byref(n1))
    pcam, pdimg, pkun, pur1, pdp1, pnfr, pfh = C.byref(cam), C.byref(dimg), P(kun), P(ur1), P(dp1), C.byref(nfr), C.byref(fh)
    pskip, pobs, ptcw, pow_, pinview = P(skip), P(has_obs), P(tcw), P(ow), P(inview)

    def call_extract():
        L.coeb_extract(*xargs)

    def call_tail():
        if fh.value:
            L.coeb_frame_destroy(fh)
        L.coeb_frame_from_extractor(m.h, ex1.h, 0, n1.value, pcam, None, pdimg, pkun, pur1, pdp1, pnfr, pfh)

    def call_local():
        km[:] = state
        L.coeb_search_local_points(m.h, fh, dev_map.h, pskip, pobs, ptcw, pow_, c05, c30, c08, pkm, pinview, None, pnm)

    def call_chain():
        call_extract(); call_tail(); call_local()

    def call_chain_head():
        call_extract(); call_tail()
    for _ in range(5):
        call_chain()
    assert n1.value == len(kps) and nfr.value == len(kps)
    out["frame_tail_from_extractor_us"] = _median_us(call_tail, 50)
    out["search_local_points_map5k_us"] = _median_us(call_local, 50)
    out["track_frame_chain_us"] = _median_us(call_chain, 50)
    out["track_frame_chain_note"] = ("blocking C calls, host buffers: coeb_extract (640x480 + boxes/T_M in, keypoints/descriptors out) -> "
                                     "coeb_frame_from_extractor (raw uint16 depth in, mvKeysUn/mvuRight/mvDepth out) -> coeb_search_local_points "
                                     "(5000-point resident local map; flags in, matches + visibility out)")
    n_local_gpu, km_local_gpu, inview_gpu = int(nm.value), km.copy(), inview.copy()
    # ---- BASELINE.json configs[3] (KITTI-shaped stereo pair, nFeatures 2000) and the extraction of configs[4] (1080p, 4000) ----
    sw, sh = 1241, 376
    left = synth.make_frame(300, sw, sh)
    right = synth.make_stereo_right(left, seed=300)
    exL, exR = cb.Extractor(2000, 1.2, NLEVELS, 20, 7, device=dev), cb.Extractor(2000, 1.2, NLEVELS, 20, 7, device=dev)
    for _ in range(3):
        kl, dl = exL.extract(left)
        kr, dr = exR.extract(right)
    bf_s, b_s = 386.1448, 386.1448 / 718.856

    def stereo_frame():
        a = exL.extract(left)
        b = exR.extract(right)
        return m.stereo_match(exL, exR, a[0], a[1], b[0], b[1], bf_s, b_s)
    n_st, ur_st, dp_st = stereo_frame()
    out["stereo_1241x376_extract_pair_us"] = _median_us(lambda: (exL.extract(left), exR.extract(right)), 20)
    # the same pair as ONE two-frame call of one extractor (left / right extractors with equal parameters, as ORB-SLAM constructs them);
    # coeb_stereo_match_frames then takes frames 0 and 1 of it
    ex2 = cb.Extractor(2000, 1.2, NLEVELS, 20, 7, device=dev)
    pair = np.ascontiguousarray(np.stack([left, right]))
    cap2 = ex2.default_cap(sw, sh)
    out2 = (np.empty((2, cap2), cb.KP_DTYPE), np.empty((2, cap2, 32), np.uint8), np.empty(2, np.int32), np.empty(2, np.int32))
    for _ in range(3):
        ex2.extract_batch_host(pair, out=out2)
    assert out2[2][0] == len(kl) and out2[2][1] == len(kr) and np.array_equal(out2[1][0, :len(kl)], dl)
    out["stereo_1241x376_extract_pair_one_call_us"] = _median_us(lambda: ex2.extract_batch_host(pair, out=out2), 20)
    n_st2, ur_st2, _ = m.stereo_match_frames(ex2, 0, ex2, 1, kl, dl, kr, dr, bf_s, b_s)
    out["stereo_one_call_equals_two_calls"] = bool(n_st2 == n_st and np.array_equal(ur_st2, ur_st))
    ex2.close()
    out["stereo_1241x376_compute_stereo_matches_us"] = _median_us(lambda: m.stereo_match(exL, exR, kl, dl, kr, dr, bf_s, b_s), 20)
    out["stereo_1241x376_points_with_depth"] = int(n_st)
    big = synth.make_frame(500, 1920, 1080)
    exB = cb.Extractor(4000, 1.2, NLEVELS, 20, 7, device=dev)
    for _ in range(3):
        kb, db = exB.extract(big)
    out["extract_1920x1080_nf4000_us"] = _median_us(lambda: exB.extract(big), 20)
    out["extract_1920x1080_keypoints"] = int(len(kb))
    # kNN 4000 x 100k, device resident
    q, t = synth.make_knn_sets(4000, 100000, seed=3)
    dq, dt = torch.from_numpy(q).cuda(dev), torch.from_numpy(t).cuda(dev)
    di, d1, d2_ = (torch.empty(4000, dtype=torch.int32, device=dq.device) for _ in range(3))
    stream = torch.cuda.current_stream(dev)
    m.set_stream(stream.cuda_stream)
    for _ in range(3):
        m.knn2_device(dq.data_ptr(), 4000, dt.data_ptr(), 100000, 0.7, di.data_ptr(), d1.data_ptr(), d2_.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    reps = 10
    for _ in range(reps):
        m.knn2_device(dq.data_ptr(), 4000, dt.data_ptr(), 100000, 0.7, di.data_ptr(), d1.data_ptr(), d2_.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    knn_ms = e0.elapsed_time(e1) / reps
    popc = 4000 * 100000 * 8
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    out["knn2_4000x100k_ms"] = knn_ms
    out["knn2_popc_per_s"] = popc / (knn_ms * 1e-3)
    # POPC issues at 16 lanes/clk/SM on CC 10.0 (XU pipe). The figure above counts the ALGORITHMIC 8 POPC per pair; the kernel folds
    # seven of the eight difference words with carry-save adders first and executes 5 POPC per pair (ncu: XU pipe 91.6 % before the
    # folding, profiles/r02a_match_pipes.json), so the algorithmic rate may exceed the pipe's peak
    out["knn2_popc_executed_per_pair"] = 5
    out["knn2_algorithmic_popc_rate_over_xu_peak"] = popc / (knn_ms * 1e-3) / (n_sm * 16 * 1.965e9)
    out["knn2_xu_pipe_frac_at_max_clock"] = (popc * 5 / 8) / (knn_ms * 1e-3) / (n_sm * 16 * 1.965e9)
    m.set_stream(0)
    # ---- Frame::ProcessMovingObject (src/Frame.cc:311-393): the producer of T_M, one frame pair per blocking call ----
    try:
        from coeb_b200 import motion as cmotion
        mo = cmotion.Motion(device=dev)
        mp_prev, mp_cur, _ = synth.make_motion_pair(0)
        for _ in range(3):
            tm_g, tr_g = mo.process(mp_prev, mp_cur)
        pmo_call = mo.prepared_process(mp_prev, mp_cur)   # ctypes arguments marshalled outside the clock, as for the other calls
        assert pmo_call() == len(tm_g)
        out["process_moving_object_us"] = _median_us(pmo_call, 20)
        # the same as the reference runs it: imGrayPre is the previous call's frame, resident on the device (coeb_process_moving_object_next)
        # (frames fed alternately; the clock brackets the calls that pair (mp_prev -> mp_cur), the pair the two-frame figure above is for)
        seq_call = mo.prepared_process_next([mp_prev, mp_cur])
        for _ in range(4):
            seq_call()
        ts = []
        for _ in range(20):
            seq_call()                      # mp_prev arrives (pairs mp_cur -> mp_prev: not timed)
            t_seq = time.perf_counter()
            n_seq = seq_call()              # mp_cur arrives
            ts.append(time.perf_counter() - t_seq)
        assert n_seq == len(tm_g)
        out["process_moving_object_next_us"] = 1e6 * float(np.median(ts))
        # one frame of the COEB front end as the tracking thread runs it, five blocking C calls back to back: ProcessMovingObject (sequence
        # form) -> extract + dynamic filter -> Frame tail -> SearchByProjection(cur, last) -> SearchLocalPoints
        ts = []
        for _ in range(20):
            seq_call()
            t_seq = time.perf_counter()
            seq_call(); call_chain_head(); call_m3(); call_local()
            ts.append(time.perf_counter() - t_seq)
        out["coeb_frame_front_end_us"] = 1e6 * float(np.median(ts))
        out["coeb_frame_front_end_note"] = ("coeb_process_moving_object_next -> coeb_extract -> coeb_frame_from_extractor -> coeb_match_lastframe -> "
                                            "coeb_search_local_points, host buffers in and out, one 640x480 frame")
        out["process_moving_object_points"] = int(tr_g["n_points"])
        out["process_moving_object_tm"] = int(len(tm_g))
        if not args.no_cpu:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import pmo
            import cv2
            cv2.setNumThreads(1)
            out["cpu_process_moving_object_us_cv2_1thread"] = _median_us(lambda: pmo.process_moving_object(mp_prev, mp_cur), 5)
            cv2.setNumThreads(0)
            out["cpu_process_moving_object_us_cv2_all_threads"] = _median_us(lambda: pmo.process_moving_object(mp_prev, mp_cur), 5)
        mo.close()
    except Exception as e:   # noqa: BLE001
        out["process_moving_object_error"] = str(e)[:200]
    if not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import orc
        fc = orc.Frame(kps, desc, orc.Camera(*cam_args), scale, uright)
        f2c = orc.Frame(k2, d2, orc.Camera(*cam_args), scale, None)
        out["cpu_search_by_projection_map5k_us"] = _median_us(lambda: orc.match_projection(fc, mp, 3.0, 0.8, state), 10)
        out["cpu_search_by_projection_lastframe_us"] = _median_us(lambda: orc.match_lastframe(fc, last, Tc, Tl, 15.0, False, True, state), 10)
        out["cpu_search_for_initialization_us"] = _median_us(lambda: orc.match_init(fc, f2c, prev, 100, 0.9, True), 10)
        th = orc.hardware_threads()
        _, _, _, secs = orc.knn2(q[:1000], t, 0.7, nthreads=th)
        out["cpu_knn2_4000x100k_ms_extrapolated"] = 4e3 * secs
        out["cpu_threads"] = th
        kun_c = orc.undistort_keypoints(kps, orc.Camera(*cam_args), None)
        ur_c, dp_c = orc.stereo_from_rgbd(kps, kun_c, np.asarray(depth_pin), cam_args[4], np.float32(1.0) / np.float32(5000.0))
        fl = orc.Frame(kun_c, desc, orc.Camera(*cam_args), scale, ur_c)
        out["cpu_frame_tail_us"] = _median_us(lambda: (orc.stereo_from_rgbd(kps, kun_c, np.asarray(depth_pin), cam_args[4], 2e-4),
                                                       orc.Frame(kun_c, desc, orc.Camera(*cam_args), scale, ur_c)), 10)
        out["cpu_search_local_points_map5k_us"] = _median_us(lambda: orc.search_local_points(fl, lm, skip, has_obs, Tcw, Ow, 3.0, 0.8, state), 10)
        n_lc, km_lc, iv_lc, _ = orc.search_local_points(fl, lm, skip, has_obs, Tcw, Ow, 3.0, 0.8, state)
        out["local_points_bit_exact_vs_cpu"] = bool(n_lc == n_local_gpu and np.array_equal(km_lc, km_local_gpu) and np.array_equal(iv_lc, inview_gpu)
                                                    and ur_c.tobytes() == ur1[:len(kps)].tobytes() and dp_c.tobytes() == dp1[:len(kps)].tobytes())
        out["local_points_in_view"] = int(iv_lc.sum())
        out["local_points_matches"] = int(n_lc)
        oL, oR = orc.Extractor(2000, 1.2, NLEVELS, 20, 7), orc.Extractor(2000, 1.2, NLEVELS, 20, 7)
        okl, odl = oL.extract(left)
        okr, odr = oR.extract(right)
        out["cpu_stereo_1241x376_compute_stereo_matches_us"] = _median_us(lambda: orc.stereo_match(oL, oR, okl, odl, okr, odr, bf_s, b_s), 5)
        n_sc, ur_sc, dp_sc = orc.stereo_match(oL, oR, okl, odl, okr, odr, bf_s, b_s)
        out["stereo_bit_exact_vs_cpu"] = bool(n_sc == n_st and ur_sc.tobytes() == ur_st.tobytes() and dp_sc.tobytes() == dp_st.tobytes())
        n_g, km_g = m.match_projection(f, mp, 3.0, 0.8, state)
        n_c, km_c = orc.match_projection(fc, mp, 3.0, 0.8, state)
        out["bit_exact_vs_cpu"] = bool(n_g == n_c and np.array_equal(km_g, km_c))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--unique", type=int, default=64, help="distinct generated images per shard (the rest are shifted copies)")
    ap.add_argument("--e2e-steps", type=int, default=60)
    ap.add_argument("--e2e-lanes", type=int, default=3, help="batches in flight in the end-to-end region (one handle and host thread each)")
    ap.add_argument("--cpu-frames", type=int, default=128, help="frames of the shard timed on the host cores (CPU baseline)")
    ap.add_argument("--ref-frames", type=int, default=FRAMES_PER_GPU, help="--impl reference: frames per step (default: the whole 256-frame shard)")
    ap.add_argument("--stage-sync", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-match", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs: device-resident part only")
    ap.add_argument("--match-only", action="store_true", help="profiling runs: the matching section only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
