"""N>1 host logic on CPU: two gloo ranks shard a frame range, receive the plan bytes from rank 0 and merge results."""
import os
import subprocess
import sys
import textwrap

from video_subtitle_extractor_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_ranges_partition_the_stream():
    for n in (0, 1, 7, 64, 50000):
        for world in (1, 2, 3, 8):
            spans = [shard.frame_range(r, world, n) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


WORKER = textwrap.dedent("""
    import os, sys, hashlib
    sys.path.insert(0, %r)
    import torch.distributed as dist
    from video_subtitle_extractor_b200 import shard, weights
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    blobs = [weights.load_plan_blob("V4/ch_det_fast")[:200000], b"rec-plan-bytes"] if rank == 0 else None
    got = shard.broadcast_blobs(blobs, 0)
    lo, hi = shard.frame_range(rank, world, 11)
    local = [(f, "rank%%d:frame%%d" %% (rank, f)) for f in range(lo, hi)]
    merged = shard.gather_by_frame(local)
    mx = shard.max_over_ranks(float(rank + 1))
    line = "RESULT %%d %%s %%s %%s %%s" %% (rank, hashlib.sha1(got[0]).hexdigest(), got[1].decode(), [m[0] for m in merged], mx)
    open(os.path.join(os.environ["VSE_TEST_OUT"], "rank%%d.txt" %% rank), "w").write(line)   # one file per rank: stdout of two ranks interleaves
    dist.destroy_process_group()
""") % ROOT


def test_two_gloo_ranks_broadcast_and_merge(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1", VSE_TEST_OUT=str(tmp_path))
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = sorted(p.read_text() for p in tmp_path.glob("rank*.txt"))
    assert len(lines) == 2
    a, b = (l.split(" ", 3) for l in lines)
    assert a[2] == b[2]                                   # same plan bytes on both ranks
    assert "rec-plan-bytes" in lines[0] and "rec-plan-bytes" in lines[1]
    assert "[0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10]" in lines[0] and "[0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10]" in lines[1]
    assert lines[0].rstrip().endswith("2.0") and lines[1].rstrip().endswith("2.0")


def test_numa_binding_helpers_are_safe_without_a_gpu():
    """bench.py binds every rank of a multi-GPU run to the CPUs next to its GPU before allocating page-locked buffers; where the
    topology cannot be read (this container: no GPU) nothing changes."""
    import os
    assert shard._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    before = os.sched_getaffinity(0)
    assert shard.bind_to_gpu_numa_node(0) is None
    assert os.sched_getaffinity(0) == before
