"""GPU test of the multi-GPU fold (simc_b200_reduce_gathered) on ONE device: the accumulator blocks three ranks would
all-gather are produced one after the other by the same handle, laid out rank after rank, folded by the library's
kernel and compared, bit for bit, with a single run over the union of the three try ranges."""
import os

import pytest

from simc_gfortran_b200 import Simc, config_from_deck, load_optics_fixture

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reduce_gathered_equals_single_run():
    import torch
    from simc_gfortran_b200.multi import _DeviceWords
    cfg = config_from_deck(os.path.join(ROOT, "decks", "c1_eep_hydrogen_hms_shms.inp"))[0]
    sim = Simc(cfg, mode="strict")
    try:
        for arm in (1, 5):
            sim.set_optics(load_optics_fixture(arm))
        n, seed, world = 30000, 5, 3
        whole = sim.run(0, world * n, seed, sim.accum_clear())
        ptr, words = sim.device_accum()
        local = torch.as_tensor(_DeviceWords(ptr, words), device="cuda:0")
        gathered = torch.empty(world * words, dtype=torch.int64, device="cuda:0")
        ext = torch.cuda.ExternalStream(sim.stream)
        for r in range(world):
            sim.run_async(r * n, n, seed)
            with torch.cuda.stream(ext):
                gathered[r * words:(r + 1) * words].copy_(local)
            sim.fetch(sim.accum_clear())                    # clears the device block for the next "rank"
        sim.reduce_gathered(gathered.data_ptr(), world)
        total = sim.fetch(sim.accum_clear())
        assert total.ntried == world * n and total.nsuccess > 1000
        assert bytes(total) == bytes(whole)
    finally:
        sim.close()
