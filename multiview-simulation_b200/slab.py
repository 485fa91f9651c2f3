"""Slab-decomposed convolution of one large volume over several GPUs (BASELINE config 5).

The reference cannot even allocate this case (ArrayImg holds < 2^31 voxels, S/SimulateMultiViewDataset.java:109).
Here the volume is distributed by z slabs, one process per GPU; libmvsim.so runs the FFT passes and
torch.distributed (NCCL over NVLink) runs the two all-to-all transposes directly on the library's exchange
buffers, whose layout IS the all_to_all_single send/receive layout (no pack / unpack kernels).
"""
import ctypes as C

from ._lib import check, dims3


class SlabConvolution:
    """One rank's share of `convolve` for a global (Z, Y, X) volume.  Buffers are torch CUDA tensors."""

    def __init__(self, ctx, shape_zyx, kshape_zyx, rank=0, world=1, dist=None):
        import torch
        self.ctx, self.rank, self.world, self.dist = ctx, rank, world, dist
        self.h = C.c_void_p()
        check(ctx._lib.mvsim_slabconv_create(ctx.h, dims3(shape_zyx), dims3(kshape_zyx), rank, world, C.byref(self.h)), ctx.h)
        info = (C.c_int64 * 8)()
        check(ctx._lib.mvsim_slabconv_info(self.h, info))
        self.z_local, self.z0, self.y_blocks, self.exchange_elems = int(info[0]), int(info[1]), int(info[2]), int(info[3])
        self.nfft = (int(info[4]), int(info[5]), int(info[6]))
        dev = torch.device("cuda", ctx.device)
        self.send = torch.empty(self.exchange_elems, dtype=torch.complex64, device=dev)
        self.recv = torch.empty(self.exchange_elems, dtype=torch.complex64, device=dev) if world > 1 else None
        check(ctx._lib.mvsim_slabconv_bind(self.h, C.c_void_p(self.send.data_ptr()),
                                           C.c_void_p(self.recv.data_ptr()) if self.recv is not None else None), ctx.h)
        self.shape = tuple(shape_zyx)

    def exchange_bytes_per_rank(self):
        """bytes this rank sends to OTHER ranks per convolution (two transposes per y block)."""
        return 2 * self.y_blocks * self.exchange_elems * 8 * (self.world - 1) // self.world

    def convolve(self, img_slab, psf, out_slab):
        """img_slab/out_slab: float32 CUDA tensors (z_local, Y, X); psf: normalised float32 CUDA tensor (KZ, KY, KX).
        Must be called on the stream the context was created on (torch's current stream)."""
        lib, ctx = self.ctx._lib, self.ctx
        assert img_slab.is_contiguous() and out_slab.is_contiguous() and psf.is_contiguous()
        assert tuple(img_slab.shape) == (self.z_local,) + self.shape[1:] == tuple(out_slab.shape)
        check(lib.mvsim_slabconv_prepare(ctx.h, self.h, C.c_void_p(psf.data_ptr()), C.c_void_p(img_slab.data_ptr())), ctx.h)
        for b in range(self.y_blocks):
            check(lib.mvsim_slabconv_forward_y(ctx.h, self.h, b), ctx.h)
            if self.world > 1:
                self.dist.all_to_all_single(self.recv, self.send)
            check(lib.mvsim_slabconv_middle_z(ctx.h, self.h), ctx.h)
            if self.world > 1:
                self.dist.all_to_all_single(self.send, self.recv)
            check(lib.mvsim_slabconv_inverse_y(ctx.h, self.h, b), ctx.h)
        check(lib.mvsim_slabconv_finish(ctx.h, self.h, C.c_void_p(out_slab.data_ptr())), ctx.h)
        return out_slab

    def close(self):
        if self.h:
            self.ctx._lib.mvsim_slabconv_destroy(self.ctx.h, self.h)
            self.h = None
