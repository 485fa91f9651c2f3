"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for the control plane only.

The path shards by view (sharding.py) with NO data-path collective: every rank simulates its own
views from its own copy of the ground truth; what crosses ranks is a barrier, the max of a timing and
(optionally) a gather of small per-view records.  Backend "nccl" on the GPU box, "gloo" in the CPU tests.
"""
import os

from .sharding import views_for_rank


class Group:
    def __init__(self, backend=None, device=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.device = device
        if self.world > 1:
            import torch.distributed as dist
            if not dist.is_initialized():
                kw = {}
                if backend == "nccl" and device is not None:
                    kw["device_id"] = device
                dist.init_process_group(backend or "gloo", **kw)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def _tensor(self, values, dtype):
        import torch
        dev = self.device if (self.dist is not None and self.dist.get_backend() == "nccl") else "cpu"
        return torch.tensor(values, dtype=dtype, device=dev)

    def max(self, value):
        """max over ranks of a float (device timings are reported as the slowest rank)."""
        if self.dist is None:
            return float(value)
        import torch
        t = self._tensor([float(value)], torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def min(self, value):
        return -self.max(-float(value))

    def sum(self, value):
        if self.dist is None:
            return int(value)
        import torch
        t = self._tensor([int(value)], torch.int64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return int(t.item())

    def broadcast_ground_truth(self, ctx, shape_zyx, host=None, src=0, tensor=None):
        """One dataset, views sharded over the ranks: rank `src` uploads the ground truth ONCE from (pinned) host memory and
        NCCL broadcasts it over NVLink; every rank gets a DeviceVolume over the received buffer.  This replaces world_size
        uploads through the shared host link by one upload plus an on-fabric copy.  `tensor` (optional)
        is a reusable float32 CUDA tensor of that shape.  Returns (volume, tensor)."""
        import torch
        from .api import DeviceVolume
        if tensor is None:
            tensor = torch.empty(tuple(shape_zyx), dtype=torch.float32, device=self.device)
        if self.rank == src:
            h = torch.from_numpy(host)
            tensor.copy_(h, non_blocking=True)
        if self.dist is not None:
            self.dist.broadcast(tensor, src=src)
        if tensor.is_cuda:
            # the copy and the broadcast are ordered on torch's current stream; the context may own another (non-blocking)
            # stream, so the ground truth must have ARRIVED before the volume is handed to it
            torch.cuda.current_stream(tensor.device).synchronize()
        return DeviceVolume.wrap(ctx, shape_zyx, tensor.data_ptr(), keepalive=tensor), tensor

    def my_views(self, n_views):
        return views_for_rank(n_views, self.rank, self.world)

    def gather_records(self, records):
        """dict view_id -> small picklable record; returns the merged dict on rank 0, None elsewhere."""
        if self.dist is None:
            return dict(records)
        out = [None] * self.world if self.rank == 0 else None
        self.dist.gather_object(dict(records), out, dst=0)
        if self.rank != 0:
            return None
        merged = {}
        for d in out:
            overlap = set(merged) & set(d)
            if overlap:
                raise RuntimeError(f"views simulated twice: {sorted(overlap)}")
            merged.update(d)
        return merged

    def close(self):
        if self.dist is not None and self.dist.is_initialized():
            self.dist.destroy_process_group()
