"""CPU: synthetic generators are deterministic, the bench's sharding / aggregation works across 2 ranks (gloo),
and the bench's reference arm runs."""
import json
import os
import subprocess
import sys
import zlib

import numpy as np

from coeb_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synth_is_seed_deterministic_and_shaped():
    a, b = synth.make_frame(5), synth.make_frame(5)
    assert a.dtype == np.uint8 and a.shape == (480, 640) and a.flags.c_contiguous and np.array_equal(a, b)
    assert not np.array_equal(a, synth.make_frame(6))
    assert synth.make_frame(1, 1241, 376).shape == (376, 1241)
    batch = synth.make_batch(16, base_seed=0, unique=8)
    assert batch["gray"].shape == (16, 480, 640) and batch["boxes"].shape == (16, synth.MAX_BOX, 4)
    assert np.array_equal(batch["gray"][3], synth.make_frame(3))
    assert not np.array_equal(batch["gray"][8], batch["gray"][0])      # shifted copy, not a duplicate
    # every 8th frame is built to take the area_flag path: two boxes whose summed area exceeds 200000 px^2
    b3 = batch["boxes"][3, :batch["nbox"][3]]
    assert ((b3[:, 2] - b3[:, 0]) * (b3[:, 3] - b3[:, 1])).sum() > 200000
    for i in range(16):
        bx = batch["boxes"][i, :batch["nbox"][i]]
        assert (bx[:, 0] >= 0).all() and (bx[:, 2] <= 640).all() and (bx[:, 3] <= 480).all()
        assert batch["ntm"][i] <= synth.MAX_TM


def test_rank_shards_are_disjoint():
    a = synth.make_batch(4, base_seed=0, unique=4)
    b = synth.make_batch(4, base_seed=1000, unique=4)
    assert not any(np.array_equal(a["gray"][i], b["gray"][j]) for i in range(4) for j in range(4))


WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import bench
rank, world, local, dist = bench.dist_setup(2)
assert world == 2 and dist is not None
ms = bench.reduce_max(dist, 10.0 + 5.0 * rank)        # slowest rank defines the step time
frames = bench.reduce_sum(dist, 256.0 * 3)            # whole-job frames over all ranks
dist.barrier()
if rank == 0:
    print("RESULT", ms, frames)
dist.destroy_process_group()
"""


def test_two_rank_aggregation_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29617")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", "29617", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][0].split()
    assert float(line[1]) == 15.0 and float(line[2]) == 2 * 256.0 * 3


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--ref-frames", "8"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libcoeb_ref.so")) or os.path.isdir("/root/reference/src")
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")   # oracle/_ref = the reference's own extractor
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["config"]["workload"].startswith("configs[1]")
    import bench
    assert d["config"] == bench.base_config(1)   # both arms print the same config object


def test_stage_byte_model_matches_survey_total():
    import bench
    sb = bench.stage_bytes(1000, 0)
    # 307,200 + 643,332 + 926,546 + 950,532 + 1,901,064 + 1,321,000 = 6,049,674 B/frame
    total = sb["pyramid"] + sb["blur"] + sb["fast"] + 1000 * (749 + 512 + 60) + 0
    assert bench.SIGMA_P == 950532
    assert int(sb["pyramid"]) == 643332 + 926546 and int(sb["blur"]) == 1901064 and int(sb["fast"]) == 950532
    assert int(total) + 0 == 6049674 - 307200 + 0 or int(total) + 307200 == 6049674


def test_documents_only_cite_profile_files_that_exist():
    """DESIGN.md / README.md / profiles/README.md / bench.py name captures under profiles/: every named file must be committed."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    have = set(os.listdir(os.path.join(root, "profiles")))
    missing = []
    for doc in ("DESIGN.md", "README.md", os.path.join("profiles", "README.md"), "bench.py"):
        text = open(os.path.join(root, doc)).read()
        for name in re.findall(r"\b(r\d\d[a-z]_[A-Za-z0-9_]+\.(?:json|csv|txt))\b", text):
            if name not in have:
                missing.append((doc, name))
    assert not missing, missing
