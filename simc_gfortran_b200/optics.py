"""COSY optics tables as arrays (the layout simc_b200_set_optics takes).

The reference reads them from ``hms/forward_cosy.dat`` etc. (shared/transp.f:294-474,
hms/mc_hms_recon.f:70-102).  ``simc_gfortran_b200/data/optics_*.npz`` holds the parsed tables of the
reference's five file pairs (data the package ships, made by tools/make_fixtures.py) so that the bench and the GPU
tests do not need the reference tree; ``write_cosy_files`` turns
tables back into the reference's fixed-column text format so the file reader of the library
can be exercised anywhere.
"""
from __future__ import annotations

import dataclasses
import os

import numpy as np

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


@dataclasses.dataclass
class OpticsTables:
    arm: int
    class_start: np.ndarray   # int32 [n_classes+1]
    fwd_coeff: np.ndarray     # float64 [n_terms,5]
    fwd_expon: np.ndarray     # int8 [n_terms,5]  (x, theta, y, phi, delta)
    length_cm: np.ndarray     # float64 [n_classes]
    adrift: np.ndarray        # int8 [n_classes]
    driftdist: np.ndarray     # float64 [n_classes]
    rec_coeff: np.ndarray     # float64 [n_rec,4]
    rec_expon: np.ndarray     # int8 [n_rec,5]

    @property
    def n_classes(self) -> int:
        return len(self.class_start) - 1

    def save(self, path: str) -> None:
        np.savez_compressed(path, **{f.name: getattr(self, f.name) for f in dataclasses.fields(self)})

    @staticmethod
    def load(path: str) -> "OpticsTables":
        z = np.load(path)
        return OpticsTables(arm=int(z["arm"]), **{k: z[k] for k in z.files if k != "arm"})


_ARM_NAME = {1: "hms", 2: "sos", 3: "hrsr", 4: "hrsl", 5: "shms"}


def load_optics_fixture(arm: int) -> OpticsTables:
    return OpticsTables.load(os.path.join(DATA_DIR, f"optics_{_ARM_NAME[arm]}.npz"))


def _g(v: float, width: int, digits: int) -> str:
    """A real that reads back to exactly the same double (17 significant digits)."""
    s = f"{v:.{digits}E}"
    return s.rjust(width)


def write_cosy_files(t: OpticsTables, fwd_path: str, rec_path: str) -> None:
    """Writes the tables in the reference's text formats, (1x,5g..,1x,6i1) / (1x,4g..,1x,5i1).

    Field widths are kept at 14 / 16 columns; values are written with as many digits as fit,
    so files written here are for exercising the reader, not for bit-exact storage.
    """
    with open(fwd_path, "w") as f:
        f.write("! forward map written by simc_gfortran_b200.optics.write_cosy_files\n")
        for k in range(t.n_classes):
            if t.length_cm[k] > 0:
                f.write(f"!LENGTH:  {t.length_cm[k] / 100.0!r} (canonical length in meters)\n")
            for i in range(t.class_start[k], t.class_start[k + 1]):
                c = "".join(_g(float(x), 14, 6) for x in t.fwd_coeff[i])
                e = t.fwd_expon[i]
                f.write(f" {c} {e[0]}{e[1]}{e[2]}{e[3]}0{e[4]}\n")
            f.write(" " + "-" * 78 + "\n")
    with open(rec_path, "w") as f:
        f.write("! reconstruction map written by simc_gfortran_b200.optics.write_cosy_files\n")
        for i in range(len(t.rec_coeff)):
            c = "".join(_g(float(x), 16, 8) for x in t.rec_coeff[i])
            e = t.rec_expon[i]
            f.write(f" {c} {e[0]}{e[1]}{e[2]}{e[3]}{e[4]}\n")
        f.write(" " + "-" * 78 + "\n")
