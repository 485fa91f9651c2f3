"""The drop-in boundary without a GPU: libmvsim.so loads, exports every symbol include/mvsim.h declares,
the ctypes mirror of the parameter struct matches the C layout, the host-only entry points agree with the
oracle, and every compute entry point fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mvsim.h")


@pytest.fixture(scope="module")
def lib():
    from mvsim_b200 import _lib
    return _lib.load()


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mvsim_[a-z0-9_]+)\s*\(", src)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_every_declared_symbol_is_exported_and_bound(lib):
    from mvsim_b200 import _lib
    declared = _declared()
    assert len(declared) >= 35
    bound = {name for name, _, _ in _lib.SYMBOLS}
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in mvsim.h but not exported"
        assert name in bound, f"{name} not bound in _lib.SYMBOLS"
    assert bound <= set(declared)


def test_header_is_plain_c_and_struct_layout_matches(tmp_path):
    from mvsim_b200 import ViewParams
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "mvsim.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(mvsim_view_params), offsetof(mvsim_view_params,kdims), offsetof(mvsim_view_params,degrees),"
                   "offsetof(mvsim_view_params,delta), offsetof(mvsim_view_params,snr), offsetof(mvsim_view_params,seed),"
                   "offsetof(mvsim_view_params,strict_reference));return MVSIM_NSTAGES;}\n")
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    from mvsim_b200 import _lib
    assert r.returncode == _lib.NSTAGES
    got = [int(v) for v in r.stdout.split()]
    exp = [C.sizeof(ViewParams)] + [getattr(ViewParams, f).offset for f in ("kdims", "degrees", "delta", "snr", "seed", "strict_reference")]
    assert got == exp


def test_host_only_entry_points_match_oracle(lib, oracle):
    from mvsim_b200 import SimulateMultiViewDataset as S
    for dims in [(289, 289, 289), (1024, 1024, 512), (7, 12, 9)]:
        for axis in range(3):
            for deg in (0, 15, 52, 90, 180, -45, 330):
                fwd, inv = S.axisRotation((dims[2], dims[1], dims[0]), axis, deg)
                ref = oracle.axis_rotation(dims, axis, deg)
                assert np.array_equal(fwd, ref)
                assert np.array_equal(inv, oracle.affine_invert(ref))
    nfft = (C.c_int64 * 3)()
    assert lib.mvsim_conv_padded_dims((C.c_int64 * 3)(1024, 1024, 512), (C.c_int64 * 3)(128, 128, 128), nfft) == 0
    assert list(nfft) == [1152, 1152, 640]
    assert lib.mvsim_conv_padded_dims((C.c_int64 * 3)(289, 289, 289), (C.c_int64 * 3)(51, 51, 51), nfft) == 0
    assert all(n >= 339 for n in nfft)
    assert lib.mvsim_conv_padded_dims((C.c_int64 * 3)(4000, 8, 8), (C.c_int64 * 3)(3, 3, 3), nfft) == 5
    assert lib.mvsim_conv_padded_dims((C.c_int64 * 3)(0, 8, 8), (C.c_int64 * 3)(3, 3, 3), nfft) == 1
    assert lib.mvsim_version() >= 100


def test_java_random_mirror_matches_jdk_sequence(oracle):
    from mvsim_b200 import JavaRandom
    a, b = JavaRandom(464232194), oracle.JavaRandom(464232194)
    for _ in range(50):
        assert a.nextInt() == b.next_int()
        assert a.nextLong() == b.next_long()
        assert a.nextDouble() == b.next_double()


def test_tiff_round_trip(tmp_path):
    from mvsim_b200 import tiff
    v = np.random.default_rng(0).random((5, 7, 9), dtype=np.float32)
    p = tmp_path / "v.tif"
    tiff.write_float_stack(str(p), v)
    assert np.array_equal(tiff.read_float_stack(str(p)), v)


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a CUDA device")
def test_no_cpu_fallback(lib):
    import mvsim_b200 as mv
    with pytest.raises(mv.MvsimError) as e:
        mv.Context(0)
    assert e.value.status == 3 and "no CPU fallback" in str(e.value)
    with pytest.raises(mv.MvsimError):
        mv.SimulateMultiViewDataset.rotateAroundAxis(np.ones((2, 2, 2), dtype=np.float32), 0, 10)
    # null context: compute entry points refuse instead of computing anything on the host
    f = np.ones(8, dtype=np.float32)
    fp = f.ctypes.data_as(C.POINTER(C.c_float))
    assert lib.mvsim_rotate_axis(None, fp, fp, (C.c_int64 * 3)(2, 2, 2), 0, 10) == 1
    assert lib.mvsim_poisson(None, fp, 8, 1.0, 0, 0) == 1
