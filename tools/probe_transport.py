"""Scratch probe: single-arm transport throughput on the device (HBM-resident rows)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from simc_gfortran_b200 import Simc, load_optics_fixture
from tests.oracle_lib import transport_inputs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
for mode in ("strict", "fast"):
    sim = Simc(mode=mode)
    ext = torch.cuda.ExternalStream(sim.stream)
    for arm in (1, 5):
        sim.set_optics(load_optics_fixture(arm))
        inp = torch.from_numpy(transport_inputs(arm, n, seed=1)).cuda()
        out = torch.empty((12, n), dtype=torch.float64, device="cuda")
        flags = torch.empty(n, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        for ms in (True, False):
            with torch.cuda.stream(ext):
                for _ in range(2):
                    sim.transport_batch_device(arm, n, inp.data_ptr(), 7, out.data_ptr(), flags.data_ptr(), ms_flag=ms, wcs_flag=ms)
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(ext)
                for _ in range(reps):
                    sim.transport_batch_device(arm, n, inp.data_ptr(), 7, out.data_ptr(), flags.data_ptr(), ms_flag=ms, wcs_flag=ms)
                e1.record(ext)
            e1.synchronize()
            ms_t = e0.elapsed_time(e1) / reps
            acc = float((flags == 0).float().mean())
            print(f"{mode:6s} arm {arm} ms/wcs={ms}: {ms_t:8.3f} ms for {n} rows -> {n/ms_t*1e3/1e6:8.2f} M rows/s, accepted {acc:.3f}")
    sim.close()
