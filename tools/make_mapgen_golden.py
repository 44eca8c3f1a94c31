"""Golden vectors for the GPU map builder (fw25_mapgen), produced by the UNMODIFIED reference package.

Runs in the build container (imports /root/reference, numpy only, no GPU):

    python tools/make_mapgen_golden.py

For every seeded case of tests/mapgen_cases.py it drives the reference exactly as `Solver.__init__` / `Solver.run` do
(solver.py:527-536, :694, :734-754): `PMLBuilder(...)`, `.run(use_pml=...)`, `InputFileWriter(...).run(...)`, then reads
the .dat files the writer produced.  Saved per case (tests/golden/mapgen_<case>.npz):
  * small cases: the 13 float32 maps + dcmap byte for byte, and the float64 d / alpha maps after the PML ramps;
  * every case: sha256 of each .dat file; the larger cases keep these plus every 4th point per axis of the a / b maps.
"""

from __future__ import annotations

import hashlib
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from tests import mapgen_cases  # noqa: E402
from tools.ref_import import import_fullwave  # noqa: E402
from tools.ref_objects import ref_bin  # noqa: E402

SUB_STRIDE = 4     # larger cases: a / b maps kept at every 4th point per axis (+ sha256 of every file)
DAT = ("rho", "K", "beta", "kappax", "kappau", "apmlx1", "bpmlx1", "apmlx2", "bpmlx2",
       "apmlu1", "bpmlu1", "apmlu2", "bpmlu2")


def run_reference(case: dict) -> dict:
    fw = import_fullwave()
    from fullwave.solver.input_file_writer import InputFileWriter
    from fullwave.solver.pml_builder import PMLBuilder
    m = mapgen_cases.medium_arrays(case)
    shape = m["sound_speed"].shape
    grid = mapgen_cases.make_grid(fw, case)
    assert tuple(int(getattr(grid, a)) for a in ("nx", "ny", "nz")[: len(shape)]) == shape
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        if case.get("lut"):
            from scipy.io import savemat
            lut = mapgen_cases.synthetic_lut(case["lut"])
            savemat(td / "lut.mat", {"database": lut["database"], "alpha_0_list": lut["alpha_list"][None, :],
                                     "power_list": lut["power_list"][None, :], "invalid_matrix": lut["invalid_matrix"]})
            medium = fw.Medium(grid, m["sound_speed"], m["density"], m["alpha_coeff"], m["alpha_power"], m["beta"],
                               path_relaxation_parameters_database=td / "lut.mat")
        else:
            medium = fw.MediumRelaxationMaps(grid, m["sound_speed"], m["density"], m["beta"], m["relax"])
        mask = np.zeros(shape, dtype=bool)
        mask[0] = True
        source = fw.Source(np.zeros((int(mask.sum()), int(grid.nt))), mask)
        sensor = fw.Sensor(mask=mask, sampling_modulus_time=1)
        use_pml = case.get("use_pml", True)
        pml = PMLBuilder(grid=grid, medium=medium, source=source, sensor=sensor, m_spatial_order=8,
                         n_pml_layer=case["n_pml"] if use_pml else 0,
                         n_transition_layer=case["n_trans"] if use_pml else 0, use_isotropic_relaxation=True)
        ext = pml.run(use_pml=use_pml)
        w = InputFileWriter(work_dir=td, grid=pml.extended_grid, medium=ext, source=pml.extended_source,
                            sensor=pml.extended_sensor, path_fullwave_simulation_bin=ref_bin(len(shape)),
                            use_exponential_attenuation=False, use_isotropic_relaxation=True)
        sim = w.run("sim")
        eshape = tuple(int(getattr(pml.extended_grid, a)) for a in ("nx", "ny", "nz")[: len(shape)])
        out = {"ext_shape": np.array(eshape), "dt": np.float64(pml.extended_grid.dt),
               "dx": np.float64(pml.extended_grid.dx)}
        sha = {}
        for stem in DAT + ("dcmap",):
            raw = (sim / f"{stem}.dat").read_bytes()
            sha[stem] = hashlib.sha256(raw).hexdigest()
            full = np.frombuffer(raw, np.int32 if stem == "dcmap" else np.float32).reshape(eshape)
            if case.get("store", True):
                out[stem] = full.copy()
            elif stem[:4] in ("apml", "bpml"):     # exp() may differ in its last bit between CPUs: keep a sample
                out[f"sub_{stem}"] = full[(slice(None, None, SUB_STRIDE),) * len(eshape)].copy()
        out["ndmap"] = np.fromfile(sim / "ndmap.dat", np.int32)[0]
        out["dmap"] = np.fromfile(sim / "dmap.dat", np.float32)
        out["sha_keys"] = np.array(list(sha))
        out["sha_vals"] = np.array([sha[k] for k in sha])
        if case.get("store_f64", False) and use_pml:
            fw2 = ext.relaxation_param_dict_for_fw2
            for letter in ("x", "u"):
                for nu in (1, 2):
                    out[f"d_{letter}_nu{nu}"] = np.asarray(fw2[f"d_{letter}_nu{nu}"], np.float64)
                    out[f"alpha_{letter}_nu{nu}"] = np.asarray(fw2[f"alpha_{letter}_nu{nu}"], np.float64)
    return out


def main():
    gold = ROOT / "tests" / "golden"
    for name, case in mapgen_cases.CASES.items():
        out = run_reference(case)
        np.savez_compressed(gold / f"mapgen_{name}.npz", **out)
        print(name, tuple(out["ext_shape"]), f"{(gold / f'mapgen_{name}.npz').stat().st_size / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
