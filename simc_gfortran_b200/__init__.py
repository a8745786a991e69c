"""simc_gfortran_b200 -- B200-native implementation of SIMC's per-event Monte Carlo loop.

The product is ``libsimc_b200.so`` (hand-written sm_100a CUDA kernels behind the C ABI of
``include/simc_b200.h``); this package is the thin Python host side used by the tests and by
``bench.py``.  There is no CPU fallback: importing works anywhere, computing needs a GPU.
"""
from .lib import (config_from_deck, Simc, SimcError, lib_path, load_library, RunConfig, Accum, ARM_HMS, ARM_SOS, ARM_HRSR,
                  ARM_HRSL, ARM_SHMS, TRANSPORT_NIN, TRANSPORT_NOUT, precompile_optics, report_info_from_deck, central_event,
                  write_reports, ReportInfo, Central)
from .optics import load_optics_fixture, OpticsTables

__all__ = ["config_from_deck", "Simc", "SimcError", "lib_path", "load_library", "RunConfig", "Accum", "ARM_HMS", "ARM_SOS",
           "ARM_HRSR", "ARM_HRSL", "ARM_SHMS", "TRANSPORT_NIN", "TRANSPORT_NOUT", "load_optics_fixture",
           "OpticsTables", "precompile_optics", "report_info_from_deck", "central_event", "write_reports", "ReportInfo",
           "Central"]
