#!/usr/bin/env python3
"""Generates the committed golden vectors of tests/golden/.

    python tests/golden/make_golden.py          (in the build container, where /root/reference is mounted)

Two kinds of vectors:
  * psf_angle0.npz   -- the reference's OWN fixture src/main/resources/Angle0.tif (51^3 float32 measured-style PSF, the
                        only kind of fixture the reference ships: inputs, no outputs), stored losslessly as a sparse
                        (index, value) list so that the GPU box, which has no /root/reference, can run the path on it.
  * stages_small.npz -- inputs from fixed seeds and the outputs of every stage of the path computed by the CPU ORACLE
                        (oracle/mvsim_oracle.c).  The reference is Java on un-vendored jars and cannot run here, so these
                        are NOT outputs of the reference: they pin the oracle against accidental change and give the GPU
                        tests a committed target that does not depend on the oracle being rebuilt on the box.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import sphere_phantom  # noqa: E402
from oracle import oracle as orc  # noqa: E402

REF_PSF = "/root/reference/src/main/resources/Angle0.tif"


def read_tiff(path):
    import importlib.util
    spec = importlib.util.spec_from_file_location("mvsim_tiff", os.path.join(ROOT, "multiview-simulation_b200", "tiff.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.read_float_stack(path)


def main():
    if os.path.exists(REF_PSF):
        psf = read_tiff(REF_PSF)
        idx = np.flatnonzero(psf).astype(np.int32)
        np.savez_compressed(os.path.join(HERE, "psf_angle0.npz"), shape=np.array(psf.shape, dtype=np.int32), index=idx,
                            value=psf.reshape(-1)[idx], source=np.array("PreibischLab/multiview-simulation src/main/resources/Angle0.tif (GPL-2)"))
        print("psf_angle0.npz:", psf.shape, len(idx), "non-zero voxels, sum", float(psf.astype(np.float64).sum()))
    else:
        print("reference not mounted: psf_angle0.npz left as committed")
    psf = load_psf()

    shape = (40, 44, 44)                       # X <= Y for the strict attenuation loop bound
    gt = sphere_phantom(shape, seed=20260101, n_spheres=120)
    out = {"gt": gt}
    out["rot_75"] = orc.rotate(gt, 0, 75)
    out["rot_axis1_33"] = orc.rotate(gt, 1, 33)
    out["att"] = orc.attenuate(out["rot_75"], 0.01)
    k = psf[13:38, 18:33, 18:33].copy()        # 25 x 15 x 15 centre crop of the fixture (keeps the test volume small)
    out["psf_crop"] = k.copy()
    kn = k.copy()
    out["con"] = orc.convolve(out["att"], kn, "direct")
    out["psf_norm"] = kn
    adj = out["con"].copy()
    out["corr"] = np.array(orc.adjust(adj, 0.0001, 1.0))
    out["adj"] = adj
    out["acq_nonoise"] = orc.extract_slices(adj, 3, -1.0)
    out["iso"] = orc.make_isotropic(out["acq_nonoise"], 3)
    out["weight"] = orc.weight_image((6, 100, 5))
    # the exact sampler of the reference (java.util.Random replay) on a small lambda ramp: pins the oracle's Poisson stage
    ramp = np.linspace(0.0, 8.0, 4096, dtype=np.float32)
    noisy = ramp.copy()
    orc.poisson(noisy, 25.0, orc.JavaRandom(464232194))
    out["poisson_ramp_in"] = ramp
    out["poisson_ramp_out"] = noisy
    pts = orc.random_points(40, (64, 48, 40), 535)
    out["bead_points"] = pts
    out["beads"] = orc.render_beads(orc.transform_points(pts, (64, 48, 40), 0, 45), (1.0, 1.0, 3.0), (0, 0, 0), (63, 47, 39))
    big, lst = orc.draw_spheres((242, 242, 242), scale=2, seed=464232194)
    out["sphere_list"] = lst
    out["phantom_120"] = orc.downsample2x(big)
    sq = np.random.default_rng(5).random((5, 9, 6), dtype=np.float32)
    out["square_in"] = sq
    out["square_out"] = orc.make_square(sq)
    np.savez_compressed(os.path.join(HERE, "stages_small.npz"), **out)
    print("stages_small.npz:", {k: v.shape for k, v in out.items()})


def load_psf():
    d = np.load(os.path.join(HERE, "psf_angle0.npz"))
    psf = np.zeros(int(np.prod(d["shape"])), dtype=np.float32)
    psf[d["index"]] = d["value"]
    return psf.reshape(tuple(d["shape"]))


if __name__ == "__main__":
    main()
