"""Slab-decomposed convolution of one large volume over several GPUs (BASELINE config 5).

The reference cannot even allocate this case (ArrayImg holds < 2^31 voxels, S/SimulateMultiViewDataset.java:109).
Here the volume is distributed by z slabs, one process per GPU; libmvsim.so runs the FFT passes and
torch.distributed (NCCL over NVLink) runs the two all-to-all transposes directly on the library's exchange
buffers, whose layout IS the all_to_all_single send/receive layout (no pack / unpack kernels).
"""
import ctypes as C
import os

from ._lib import check, dims3


class SlabConvolution:
    """One rank's share of `convolve` for a global (Z, Y, X) volume.  Buffers are torch CUDA tensors."""

    def __init__(self, ctx, shape_zyx, kshape_zyx, rank=0, world=1, dist=None, p2p=True):
        """p2p (world > 1): exchanges fused into the kernels as NVLink peer stores (CUDA IPC buffers, one stream-ordered
        barrier after the y pass and after the z pass); p2p=False: two NCCL all_to_all_single per y block."""
        import torch
        self.ctx, self.rank, self.world, self.dist = ctx, rank, world, dist
        self.h = C.c_void_p()
        check(ctx._lib.mvsim_slabconv_create(ctx.h, dims3(shape_zyx), dims3(kshape_zyx), rank, world, C.byref(self.h)), ctx.h)
        info = (C.c_int64 * 8)()
        check(ctx._lib.mvsim_slabconv_info(self.h, info))
        self.z_local, self.z0, self.y_blocks, self.exchange_elems = int(info[0]), int(info[1]), int(info[2]), int(info[3])
        self.nfft = (int(info[4]), int(info[5]), int(info[6]))
        dev = torch.device("cuda", ctx.device)
        self.shape = tuple(shape_zyx)
        self._side, self._side_ctx, self._flags = None, None, None
        self.overlap_blocks = os.environ.get("MVSIM_SLAB_OVERLAP", "1") != "0"      # A/B knob: 0 = y blocks one after the other
        self.p2p = bool(p2p) and world > 1
        # host_plane: the process group cannot move CUDA tensors (gloo; ranks that SHARE one GPU, where NCCL refuses to run).
        # Handles travel as CPU tensors, the cross-rank barrier is stream-synchronise + host barrier, and the all-to-all of the
        # non-p2p mode is staged through host memory.  Same kernels, same buffers, same layouts -- used by the 1-GPU tests.
        self.host_plane = world > 1 and dist.get_backend() != "nccl"
        if self.p2p:
            import numpy as np
            self.nbuf = min(2, self.y_blocks)
            mine = np.zeros(self.nbuf * 2 * 64, dtype=np.uint8)
            check(ctx._lib.mvsim_slabconv_p2p_alloc(ctx.h, self.h, self.nbuf, C.c_void_p(mine.ctypes.data)), ctx.h)
            t = torch.from_numpy(mine) if self.host_plane else torch.from_numpy(mine).to(dev)
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
            allh = torch.cat(parts).cpu().numpy()
            check(ctx._lib.mvsim_slabconv_p2p_open(ctx.h, self.h, C.c_void_p(allh.ctypes.data)), ctx.h)
            check(ctx._lib.mvsim_slabconv_p2p_select(self.h, 0), ctx.h)
            self._flag = torch.zeros(1, dtype=torch.float32, device="cpu" if self.host_plane else dev)
            dist.barrier()
            return
        # up to three exchange buffer sets: the all-to-all of one y block overlaps the kernels of its neighbours
        self.nbuf = 1 if world == 1 else min(3, self.y_blocks)
        self.send = [torch.empty(self.exchange_elems, dtype=torch.complex64, device=dev) for _ in range(self.nbuf)]
        self.recv = [torch.empty(self.exchange_elems, dtype=torch.complex64, device=dev) if world > 1 else None for _ in range(self.nbuf)]
        self._bind(0)
        self.shape = tuple(shape_zyx)

    def _bind(self, i):
        r = self.recv[i]
        check(self.ctx._lib.mvsim_slabconv_bind(self.h, C.c_void_p(self.send[i].data_ptr()),
                                                C.c_void_p(r.data_ptr()) if r is not None else None), self.ctx.h)

    def _barrier(self, flag=None):
        """All ranks' kernels enqueued so far ON THE CURRENT STREAM have finished before any rank's later kernels on it start."""
        if self.host_plane:
            import torch
            torch.cuda.current_stream().synchronize()
            self.dist.barrier()
        else:
            self.dist.all_reduce(self._flag if flag is None else flag)      # stream-ordered (NCCL joins the current stream's chain)

    def _all_to_all(self, dst, src, async_op):
        if not self.host_plane:
            return self.dist.all_to_all_single(dst, src, async_op=async_op)
        import torch
        torch.cuda.current_stream().synchronize()
        s = torch.view_as_real(src).cpu()
        d = torch.empty_like(s)
        self.dist.all_to_all_single(d, s)
        torch.view_as_real(dst).copy_(d)

        class _Done:
            def wait(self):
                return True
        return _Done()

    def exchange_bytes_per_rank(self):
        """bytes this rank sends to OTHER ranks per convolution (two transposes per y block; in p2p mode the same bytes
        leave as peer stores of the y and z kernels)."""
        return 2 * self.y_blocks * self.exchange_elems * 8 * (self.world - 1) // self.world

    def convolve(self, img_slab, psf, out_slab):
        """img_slab/out_slab: float32 CUDA tensors (z_local, Y, X); psf: normalised float32 CUDA tensor (KZ, KY, KX).
        Must be called on the stream the context was created on (torch's current stream)."""
        lib, ctx = self.ctx._lib, self.ctx
        assert img_slab.is_contiguous() and out_slab.is_contiguous() and psf.is_contiguous()
        assert tuple(img_slab.shape) == (self.z_local,) + self.shape[1:] == tuple(out_slab.shape)
        check(lib.mvsim_slabconv_prepare(ctx.h, self.h, C.c_void_p(psf.data_ptr()), C.c_void_p(img_slab.data_ptr())), ctx.h)
        nb = self.y_blocks
        if self.p2p and nb >= 2 and self.nbuf >= 2 and self.overlap_blocks:
            self._convolve_p2p_two_streams(nb)
        elif self.p2p:
            for b in range(nb):
                check(lib.mvsim_slabconv_p2p_select(self.h, b % self.nbuf), ctx.h)
                check(lib.mvsim_slabconv_forward_y(ctx.h, self.h, b), ctx.h)       # stores into the owners' z-pass buffers
                self._barrier()                                                     # stream-ordered cross-rank barrier
                check(lib.mvsim_slabconv_middle_z(ctx.h, self.h), ctx.h)            # stores into the owners' inverse buffers
                self._barrier()
                check(lib.mvsim_slabconv_inverse_y(ctx.h, self.h, b), ctx.h)
        elif self.world == 1:
            for b in range(nb):
                check(lib.mvsim_slabconv_forward_y(ctx.h, self.h, b), ctx.h)
                check(lib.mvsim_slabconv_middle_z(ctx.h, self.h), ctx.h)
                check(lib.mvsim_slabconv_inverse_y(ctx.h, self.h, b), ctx.h)
        else:
            # software pipeline over the y blocks: forward(t) | middle(t-1) | inverse(t-2); each exchange is
            # asynchronous (NCCL's stream) and is waited for only where its data is consumed
            works = {}
            for t in range(nb + 2):
                if t < nb:
                    i = t % self.nbuf
                    self._bind(i)
                    check(lib.mvsim_slabconv_forward_y(ctx.h, self.h, t), ctx.h)
                    works[t] = self._all_to_all(self.recv[i], self.send[i], True)
                if 0 <= t - 1 < nb:
                    b, i = t - 1, (t - 1) % self.nbuf
                    works.pop(b).wait()
                    self._bind(i)
                    check(lib.mvsim_slabconv_middle_z(ctx.h, self.h), ctx.h)
                    works[b] = self._all_to_all(self.send[i], self.recv[i], True)
                if 0 <= t - 2 < nb:
                    b, i = t - 2, (t - 2) % self.nbuf
                    works.pop(b).wait()
                    self._bind(i)
                    check(lib.mvsim_slabconv_inverse_y(ctx.h, self.h, b), ctx.h)
        check(lib.mvsim_slabconv_finish(ctx.h, self.h, C.c_void_p(out_slab.data_ptr())), ctx.h)
        return out_slab

    def _convolve_p2p_two_streams(self, nb):
        """Overlap-save y blocks on two streams: the y forward pass of block b + 1 (its peer stores are the first transpose: NVLink
        bound, 1.9 ms per block at config 5 on 8 GPUs) runs under the fused z pass of block b (compute bound, 2.0 ms), and the inverse y
        pass of block b under the z pass of block b + 1.  Blocks alternate between the context's stream and a side stream with its own
        context (same plan, same buffers: the two buffer sets make consecutive blocks independent); forward y passes are chained so
        that two of them never share the link; every stream ends in the main one.  The cross-rank barriers stay stream ordered."""
        import torch
        from .api import Context
        lib = self.ctx._lib
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
            self._side_ctx = Context(self.ctx.device, cuda_stream=self._side.cuda_stream)
            self._flags = [self._flag, self._flag.clone()]
        prepared = torch.cuda.Event()
        prepared.record(main)
        fwd_done, blk_done = None, []
        for b in range(nb):
            stream, ctx = (main, self.ctx) if b % 2 == 0 else (self._side, self._side_ctx)
            with torch.cuda.stream(stream):
                stream.wait_event(prepared)
                if fwd_done is not None:
                    stream.wait_event(fwd_done)                     # one forward y pass on the link at a time
                if b >= self.nbuf:
                    stream.wait_event(blk_done[b - self.nbuf])      # the buffer set is free again
                check(lib.mvsim_slabconv_p2p_select(self.h, b % self.nbuf), ctx.h)
                check(lib.mvsim_slabconv_forward_y(ctx.h, self.h, b), ctx.h)
                fwd_done = torch.cuda.Event()
                fwd_done.record(stream)
                self._barrier(self._flags[b % 2])
                check(lib.mvsim_slabconv_middle_z(ctx.h, self.h), ctx.h)
                self._barrier(self._flags[b % 2])
                check(lib.mvsim_slabconv_inverse_y(ctx.h, self.h, b), ctx.h)
                e = torch.cuda.Event()
                e.record(stream)
                blk_done.append(e)
        for e in blk_done:
            main.wait_event(e)

    def close(self):
        if self.h:
            if self._side_ctx is not None:
                self._side.synchronize()
                self._side_ctx.close()
                self._side_ctx = None
            self.ctx._lib.mvsim_slabconv_destroy(self.ctx.h, self.h)
            self.h = None


class SlabView:
    """ONE view (loop body S/SimulateMultiViewDataset.java:570-585) of a volume distributed by z slabs: rank r produces the
    planes [z0, z0 + z_local) of every stage.  rotate + attenuate reads the WHOLE ground truth (a rotation about x reaches planes
    far outside the slab; every rank holds the volume, e.g. from broadcast_ground_truth), the convolution is SlabConvolution,
    adjustImage's mean (S/Tools.java:143-147) is an all-gather of one double per rank summed in rank order, and extractSlices keeps
    the global planes z % inc == 0 with Philox counters of GLOBAL voxel indices -- so the result equals the undecomposed view."""

    def __init__(self, ctx, shape_zyx, kshape_zyx, rank=0, world=1, dist=None, p2p=True):
        self.ctx, self.rank, self.world, self.dist = ctx, rank, world, dist
        self.shape = tuple(int(s) for s in shape_zyx)
        self.conv = SlabConvolution(ctx, shape_zyx, kshape_zyx, rank, world, dist, p2p=p2p)
        self.z0, self.z_local = self.conv.z0, self.conv.z_local

    def kept_planes(self, inc):
        """(first kept global plane index, number of kept planes) of this rank's slab."""
        k0 = (self.z0 + inc - 1) // inc
        k1 = (self.z0 + self.z_local - 1) // inc
        return k0, max(0, k1 - k0 + 1) if k0 * inc < self.z0 + self.z_local else 0

    def simulate(self, gt, psf, degrees, axis=0, delta=0.01, min_value=0.0001, target_avg=1.0, inc=3, snr=25.0, seed=0, stream=0,
                 strict_reference=True):
        """gt: float32 CUDA tensor (Z, Y, X), the whole ground truth; psf: raw float32 CUDA tensor, normalised IN PLACE like :255.
        Returns (kept slices of this slab as a CUDA tensor (n_kept, Y, X), convolved + adjusted slab (z_local, Y, X))."""
        import torch
        lib, ctx = self.ctx._lib, self.ctx
        z, y, x = self.shape
        assert tuple(gt.shape) == self.shape and gt.is_contiguous() and psf.is_contiguous()
        dev = gt.device
        att = torch.empty((self.z_local, y, x), dtype=torch.float32, device=dev)
        check(lib.mvsim_slab_rotate_attenuate(ctx.h, C.c_void_p(gt.data_ptr()), dims3(self.shape), axis, int(degrees), float(delta),
                                              int(strict_reference), self.z0, self.z_local, C.c_void_p(att.data_ptr())), ctx.h)
        from .api import DeviceVolume
        k = DeviceVolume.wrap(ctx, tuple(psf.shape), psf.data_ptr(), keepalive=psf)
        check(lib.mvsim_dev_psf_normalize(ctx.h, k.h, None), ctx.h)              # every rank normalises its copy identically
        k.free()
        con = torch.empty_like(att)
        self.conv.convolve(att, psf, con)
        del att
        # adjustImage over the WHOLE volume: one double per rank, gathered and added in rank order
        mine = torch.zeros(1, dtype=torch.float64, device=dev)
        check(lib.mvsim_slab_sum(ctx.h, C.c_void_p(con.data_ptr()), con.numel(), C.c_void_p(mine.data_ptr())), ctx.h)
        if self.world > 1:
            if self.conv.host_plane:
                torch.cuda.current_stream().synchronize()
                parts = [torch.zeros(1, dtype=torch.float64) for _ in range(self.world)]
                self.dist.all_gather(parts, mine.cpu())
                sums = torch.cat(parts).to(dev)
            else:
                sums = torch.empty(self.world, dtype=torch.float64, device=dev)
                self.dist.all_gather_into_tensor(sums, mine)
        else:
            sums = mine
        check(lib.mvsim_slab_adjust(ctx.h, C.c_void_p(con.data_ptr()), con.numel(), C.c_void_p(sums.data_ptr()), self.world,
                                    float(z) * y * x, min_value, target_avg), ctx.h)
        _, nk = self.kept_planes(inc)
        out = torch.empty((max(nk, 1), y, x), dtype=torch.float32, device=dev)
        n = C.c_int64(0)
        check(lib.mvsim_slab_extract(ctx.h, C.c_void_p(con.data_ptr()), dims3(self.shape), self.z0, self.z_local, inc, float(snr),
                                     seed & ((1 << 64) - 1), stream, C.c_void_p(out.data_ptr()), C.byref(n)), ctx.h)
        assert n.value == nk
        return out[:nk], con

    def close(self):
        self.conv.close()
