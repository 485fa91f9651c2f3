"""N > 1 path on CPU: world_size-2 gloo run of the plumbing bench.py uses (views shard
with no data-path collective; results must not depend on the number of ranks)."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_views_for_rank_partitions_views():
    from mvsim_b200 import views_for_rank
    for n, w in [(6, 1), (6, 2), (6, 4), (6, 8), (8, 8), (7, 3), (0, 2)]:
        parts = [views_for_rank(n, r, w) for r in range(w)]
        assert sorted(v for p in parts for v in p) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        views_for_rank(6, 2, 2)


@pytest.mark.timeout(300)
def test_two_rank_gloo_run_matches_single_rank(tmp_path):
    sys.path.insert(0, HERE)
    from dist_worker import fake_view
    out = tmp_path / "r.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(HERE, "dist_worker.py"), "6", str(out)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=280)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(out.read_text())
    assert d["world"] == 2 and d["total"] == 6 and d["slowest"] == 11.0
    assert d["records"] == {str(v): fake_view(v) for v in range(6)}
