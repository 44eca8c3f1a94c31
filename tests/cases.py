"""Seeded small problems shared by the CPU and GPU tests (sizes the oracle finishes in < 1 s)."""

from fullwave25_b200 import synthetic

CASES = {
    # name: (kwargs for synthetic.make_problem)
    "het3d": dict(shape=(48, 52, 56), nT=60, modT=2, seed=1),
    "het3d_ragged": dict(shape=(45, 47, 53), nT=40, modT=3, seed=5, n_sensors=37, n_air=5),
    "het3d_long": dict(shape=(44, 44, 44), nT=300, modT=7, seed=7, n_pml=5, n_trans=3),
    "hom3d": dict(shape=(44, 46, 48), nT=50, modT=1, seed=3, homogeneous=True, n_air=0),
    "het2d": dict(shape=(80, 90), nT=200, modT=3, seed=2),
    "het2d_ragged": dict(shape=(83, 77), nT=150, modT=4, seed=9, n_sensors=50),
    "het2d_long": dict(shape=(120, 100), nT=1200, modT=5, seed=11),
    "hom2d": dict(shape=(64, 70), nT=100, modT=1, seed=4, homogeneous=True, n_air=0),
}

# Cases whose sources and air voxels all lie more than 8 planes away from the 2-GPU x-slab interface: the only kind
# of input for which the reference's own 2-GPU run equals its 1-GPU run (it injects sources and zeroes air voxels
# only in a GPU's OWNED planes, so a neighbour's ghost copy of such a cell goes stale; measured on a B200 pair:
# 0.2-0.4 rel-L2 between the reference's 1- and 2-GPU traces on the cases above, bit-identical on these).
CASES_2GPU = {
    "far3d": dict(shape=(72, 44, 46), nT=80, modT=3, seed=13, n_pml=5, n_trans=3, n_air=0, n_sensors=48),
    "far2d": dict(shape=(96, 70), nT=160, modT=4, seed=17, n_pml=6, n_trans=4, n_air=0, n_sensors=40),
}
CASES.update(CASES_2GPU)

# The anisotropic-relaxation engine family (per-axis kappa / a / b maps, every array different): goldens from the
# reference's fullwave2_{2d,3d}_2_relax_multi_gpu_sm_100_cuda129 binaries.
CASES_ANISO = {
    "aniso2d": dict(shape=(70, 76), nT=150, modT=3, seed=21, aniso=True, n_air=0),
    "aniso3d": dict(shape=(40, 44, 46), nT=60, modT=2, seed=22, aniso=True, n_air=0),
}
CASES.update(CASES_ANISO)


# BASELINE.json-sized grids (configs[3]: 280^3 = examples/wave_3d; configs[2]: 1457 x 2178 = examples/convex_transducer)
# with the boundary layer the reference's Solver uses (36 + 36 + 8 cells): a few hundred steps, 400 sensors in the part
# of the domain the pulse reaches.  Goldens: tests/golden/ref_big{3d,2d}.npz (reference sm_100 binaries on a B200).
CASES_BIG = {
    "big3d": dict(shape=(280, 280, 280), nT=320, modT=4, seed=51, n_pml=36, n_trans=36, block=20, n_sensors=6000, n_air=300),
    "big2d": dict(shape=(1457, 2178), nT=480, modT=4, seed=52, n_pml=36, n_trans=36, block=30, n_sensors=12000, n_air=60),
}


def make(name):
    if name in CASES_BIG:
        import numpy as np
        kw = dict(CASES_BIG[name])
        pb = synthetic.make_problem(**kw)
        nb = 8 + kw["n_pml"] + kw["n_trans"]
        reach = nb + int(kw["nT"] * 0.2) + 12                 # cfl 0.2: the pulse front after nT steps
        near = pb.outc[pb.outc[:, 0] < reach]
        pb.outc = np.ascontiguousarray(near[:: max(1, len(near) // 400)][:400])
        return pb.normalise()
    return synthetic.make_problem(**CASES[name])
