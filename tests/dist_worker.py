"""Worker of tests/test_sharding_gloo.py: launched twice by torch.distributed.run with the gloo backend.
Exercises the multi-process plumbing bench.py uses at N > 1 (view sharding, barrier, max/sum over ranks,
gather of per-view records) with a CPU stand-in for the per-view work."""
import json
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from mvsim_b200.distributed import Group  # noqa: E402


def fake_view(view_id):
    rng = np.random.default_rng(1000 + view_id)      # keyed by view id only, like the Philox stream
    return zlib.crc32(rng.integers(0, 255, 4096, dtype=np.uint8).tobytes())


def main():
    n_views = int(sys.argv[1])
    out = sys.argv[2]
    g = Group("gloo")
    mine = g.my_views(n_views)
    records = {v: fake_view(v) for v in mine}
    g.barrier()
    slowest = g.max(10.0 + g.rank)
    total = g.sum(len(mine))
    merged = g.gather_records(records)
    if g.rank == 0:
        with open(out, "w") as f:
            json.dump({"world": g.world, "slowest": slowest, "total": total, "records": {str(k): v for k, v in merged.items()}}, f)
    g.close()


if __name__ == "__main__":
    main()
