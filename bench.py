#!/usr/bin/env python
"""bench.py -- generated / accepted events per second of SIMC's event loop on B200.

Workload (BASELINE.json configs[0], the configuration the metric and the 1e9 ev/s target are
quoted on): C1 = H(e,e'p) elastic, HMS electron + SHMS proton, decks/c1_eep_hydrogen_hms_shms.inp.
A step = one pass of the loop (simc.f:169-351) over --tries tries on every GPU; every rank
works on its own range of the try index (weak scaling), and the integer accumulators are
all-reduced over NCCL once at the end.

    python bench.py --gpus 1 --steps K --warmup W            # this framework
    python bench.py --impl reference --steps K --warmup W    # CPU loop on the host cores (oracle port)

Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# BASELINE.json configs; C1 is the one the metric is quoted on and the default.
CONFIGS = {
    "c1": ("c1_eep_hydrogen_hms_shms.inp", "C1 H(e,e'p) elastic, HMS e + SHMS p, radiative corrections on"),
    "c2": ("c2_eep_carbon_hms_sos.inp", "C2 C-12 A(e,e'p), Benhar spectral function, HMS e + SOS p, radiative corrections on"),
    "c3": ("c3_eepi_hydrogen_hms_shms.inp", "C3 H(e,e'pi+)n with pion decay in flight, HMS e + SHMS pi"),
    "c4": ("c4_semi_deuterium_hms_shms.inp", "C4 D(e,e'pi-)X semi-inclusive, CTEQ5M + Christy 2021 fit, pion decay in flight, HMS e + SHMS pi"),
    "c5": ("c5_eek_hydrogen_hrsl_hrsr.inp", "C5 H(e,e'K+)Lambda, HRS-L e + HRS-R K"),
}
METRIC = "generated events/s (ntried per second), H(e,e'p) HMS+SHMS"


def deck_of(args):
    name, what = CONFIGS[args.config]
    return os.path.join(ROOT, "decks", name), f"{what} (decks/{name})"


def sf_table():
    """Spectral function of C2 (simc_gfortran_b200/data/benharsf_12.npz = the reference's benharsf_12.dat)."""
    import numpy as np
    z = np.load(os.path.join(ROOT, "simc_gfortran_b200", "data", "benharsf_12.npz"))
    return z["pm"], z["em"], z["sf_proton"]


# ---- algorithmic work per stage ---------------------------------------------------------------
# One FP64 add, mul, div or sqrt = 1 flop; library calls (log, log10, pow, exp, sin, cos, acos, atan) are counted
# separately as "transcendentals" and enter `achieved` at FLOPS_PER_TRANS flops each (the table-driven log of
# csrc/fastlog.cuh is 22 FP64 operations; the CUDA library's pow / sincos are more).
#   COSY maps   : forward class with T terms 25 + 14 T, reconstruction 25 + 12 T, times the measured call counts
#   hut         : ~4 kflop + ~330 sqrt and ~280 transcendentals per track that reaches the hut (202 gauss1 + 77 musc)
#   generation  : ~2.5 kflop + ~250 transcendentals per try (bremos, enerloss_new, trip_thru_target; C1 figures)
#   finish      : per event that passed both arms: recon kinematics + 2 x sigep ~0.4 kflop + 12, and the deferred
#                 peaked_rad_weight (2 x bremos) ~0.6 kflop + 130 when radiation is on
FLOPS_PER_TRANS = 25.0
HUT_FLOPS, HUT_TRANS = 4330.0, 280.0
GEN_FLOPS, GEN_TRANS = 2500.0, 250.0
FIN_FLOPS, FIN_TRANS = 400.0, 12.0
RADW_FLOPS, RADW_TRANS = 600.0, 130.0


def map_flops(acc, optics):
    per_arm = []
    for which, arm in ((0, optics["e"]), (1, optics["p"])):
        calls = [int(x) for x in acc.transp_calls[which]]
        n_terms = [int(arm.class_start[k + 1] - arm.class_start[k]) for k in range(arm.n_classes)]
        f = sum(calls[k] * (25 + 14 * n_terms[k]) for k in range(arm.n_classes))
        f += calls[47] * (25 + 12 * len(arm.rec_coeff))
        per_arm.append(float(f))
    return per_arm


def stage_model(acc, optics, using_rad):
    """Algorithmic (flops, transcendentals) of the four stages for the tries in `acc`."""
    maps_e, maps_p = map_flops(acc, optics)
    hut_e, hut_p = float(acc.stop[0][2]), float(acc.stop[1][2])          # tracks that reached the hut
    n_fin = float(acc.stop[0][1])                                         # passed both arms (E arm runs last)
    fin_f = FIN_FLOPS + (RADW_FLOPS if using_rad else 0.0)
    fin_t = FIN_TRANS + (RADW_TRANS if using_rad else 0.0)
    return {"k_generate": (acc.ntried * GEN_FLOPS, acc.ntried * GEN_TRANS),
            "k_arm<hadron>": (maps_p + hut_p * HUT_FLOPS, hut_p * HUT_TRANS),
            "k_arm<electron>": (maps_e + hut_e * HUT_FLOPS, hut_e * HUT_TRANS),
            "k_finish": (n_fin * fin_f, n_fin * fin_t)}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def cpu_loop(cfg, n, seed, threads, first=0, ranlux=True):
    """The oracle's loop (CPU restatement of simc.f:169-351): test infrastructure used here only as
    the CPU baseline / reference arm, never on the product path."""
    from tests.oracle_lib import Oracle
    from simc_gfortran_b200 import load_optics_fixture
    orc = Oracle()
    for arm in (cfg.electron_arm, cfg.hadron_arm):
        orc.set_optics(load_optics_fixture(arm))
    if cfg.doing_heavy:
        orc.set_sf_table(*sf_table())
    if cfg.doing_semi:
        from tests.oracle_lib import load_cteq5_fixture, load_pfermi_fixture
        orc.set_cteq5_table(load_cteq5_fixture())
        orc.set_pfermi_table(*load_pfermi_fixture())
    t0 = time.perf_counter()
    acc = orc.run(cfg, first, n, seed, threads=threads, ranlux=ranlux)
    dt = time.perf_counter() - t0
    return acc, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from simc_gfortran_b200 import config_from_deck
    deck, workload = deck_of(args)
    cfg = config_from_deck(deck)[0]
    cores = os.cpu_count() or 1
    n = args.cpu_tries if args.cpu_tries else 25000 * cores
    for w in range(args.warmup):
        cpu_loop(cfg, max(n // 8, 1000), 900 + w, cores)
    tot_t, tot_n, tot_acc = 0.0, 0, 0
    for k in range(args.steps):
        acc, dt = cpu_loop(cfg, n, 1000 + k, cores, first=k * n)
        tot_t += dt
        tot_n += acc.ntried
        tot_acc += acc.nsuccess
    v = tot_n / tot_t
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "events/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "accepted_per_s": tot_acc / tot_t,
            "config": {"workload": workload, "tries_per_step": n,
                       "note": "the Fortran reference cannot be built here (no gfortran, CTP needs SunRPC); this is its "
                               "C++ restatement (oracle/), -O2 -ffp-contract=off, one thread per host core, each with its own "
                               "RANLUX luxury-3 stream like independent simc processes"},
            "cpu_baseline": {"value": v, "unit": "events/s", "cores": cores, "kind": "port",
                             "sample": f"{n} tries per step x {args.steps} steps, all host cores"},
            "e2e": {"value": v, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tries", type=int, default=1 << 23, help="tries per step per GPU")
    ap.add_argument("--batch", type=int, default=1 << 22, help="tries per pass of the stage pipeline")
    ap.add_argument("--mode", default=os.environ.get("SIMC_B200_MODE", "strict"), choices=["strict", "fast"])
    ap.add_argument("--cpu-tries", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true",
                    help="long sweeps (BASELINE configs[4]: 1e11 tries): the device-timed steps only, `e2e` is null")
    ap.add_argument("--config", default="c1", choices=sorted(CONFIGS), help="BASELINE.json configuration (default: the headline C1)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from simc_gfortran_b200 import Accum, Simc, config_from_deck, load_optics_fixture
    from simc_gfortran_b200.multi import allreduce_device

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libsimc_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    deck, workload = deck_of(args)
    cfg, _, charge = config_from_deck(deck)
    sim = Simc(cfg, device=local, mode=args.mode)
    if cfg.doing_heavy:
        sim.set_sf_table(*sf_table())
    if cfg.doing_semi:          # tables of C4: the reference's cteq5/cteq5m.tbl and deut.dat as fixtures
        import numpy as np
        z = np.load(os.path.join(ROOT, "simc_gfortran_b200", "data", "cteq5m.npz"))
        sim.set_cteq5_table({k: (z[k] if z[k].ndim else z[k].item()) for k in z.files})
        z = np.load(os.path.join(ROOT, "simc_gfortran_b200", "data", "pfermi_deut.npz"))
        sim.set_pfermi_table(z["pval"], z["mprob"])
    optics = {"e": load_optics_fixture(cfg.electron_arm), "p": load_optics_fixture(cfg.hadron_arm)}
    sim.set_optics(optics["e"])
    sim.set_optics(optics["p"])
    sim.set_batch(args.batch)
    ext = torch.cuda.ExternalStream(sim.stream)
    n = args.tries
    seed = 20240611
    peak_fma, peak_muladd = sim.fp64_peak()

    def first_try(step):
        return (step * world + rank) * n          # disjoint ranges: results do not depend on the GPU count

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    acc = sim.accum_clear()
    for w in range(args.warmup):
        sim.run(first_try(10_000 + w), n, seed, acc)
    # ---- device-timed region: K steps back to back, accumulators stay on the device
    sim.stage_times(enable=True)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = sim.launch_count
    e0.record(ext)
    for k in range(args.steps):
        sim.run_async(first_try(k), n, seed)
    e1.record(ext)
    e1.synchronize()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = sim.launch_count - launches0
    stage_ms, stage_launches = sim.stage_times(enable=False)
    acc = sim.accum_clear()
    sim.fetch(acc)
    # ---- end-to-end: the public call with host accumulators (launch parameters in, accumulators out)
    barrier()
    t0 = time.perf_counter()
    # every step ends like a run ends: ONE collective over NVLink (all-gather of the device accumulator blocks on the
    # handle's stream + the library's fold kernel, multi.py), then the total comes back to the host of every rank
    acc_e2e = sim.accum_clear()
    ar0, ar1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    allreduce_ms = 0.0
    for k in range(0 if args.skip_e2e else args.steps):
        if world == 1:
            sim.run(first_try(100 + k), n, seed, acc_e2e)
        else:
            sim.run_async(first_try(100 + k), n, seed)
            ar0.record(ext)
            allreduce_device(sim)
            ar1.record(ext)
            sim.fetch(acc_e2e)            # every rank adds the all-rank total of this step
            allreduce_ms += ar0.elapsed_time(ar1)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag.set()
    sampler.join()
    # ---- the device-timed steps end the same way (outside their timed region: `value` is the loop alone)
    times = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        # the accumulators of the device-timed steps were fetched per rank above: fold them on the host
        from simc_gfortran_b200.multi import allreduce_accum
        acc = allreduce_accum(acc)
        assert args.skip_e2e or acc_e2e.ntried == n * args.steps * world, (acc_e2e.ntried, n, args.steps, world)
    dev_ms, e2e_ms = float(times[0]), float(times[1])

    if rank == 0:
        tries_total = acc.ntried
        assert tries_total == n * args.steps * world, (tries_total, n, args.steps, world)
        gen_per_s = tries_total / (dev_ms * 1e-3)
        acc_per_s = acc.nsuccess / (dev_ms * 1e-3)
        names = ["k_generate", "k_arm<hadron>", "k_arm<electron>", "k_finish"]
        model = stage_model(acc, optics, bool(cfg.using_rad))
        total_f = sum(v[0] for v in model.values())
        total_t = sum(v[1] for v in model.values())
        flops = total_f + FLOPS_PER_TRANS * total_t
        flops_per_try = flops / tries_total
        # dominant kernel = the stage with the largest device time (per-rank numbers of rank 0; its share of the
        # all-rank counts is 1 / world: weak scaling)
        dom = max(range(4), key=lambda k: stage_ms[k])
        stages = {}
        for k, nm in enumerate(names):
            f, t = model[nm]
            w = (f + FLOPS_PER_TRANS * t) / world
            stages[nm] = {"ms": stage_ms[k], "launches": stage_launches[k], "flops": f / world, "transcendentals": t / world,
                          "tflops": w / (stage_ms[k] * 1e-3) / 1e12 if stage_ms[k] > 0 else 0.0}
            stages[nm]["frac"] = stages[nm]["tflops"] / peak_muladd if peak_muladd else None
        achieved = stages[names[dom]]["tflops"]
        # DRAM bytes of that stage per launch (= per batch of --batch tries), from one `ncu --set full` capture
        traffic, traffic_src = None, None
        tfile = os.path.join(ROOT, "profiles", "r2_traffic.json")
        if os.path.exists(tfile) and args.config == "c1":
            tj = json.load(open(tfile))
            traffic = tj["bytes_per_4M_tries"][names[dom]] * args.batch / 4194304
            traffic_src = tj["source"]
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak = json.load(open(peaks_file)).get("hbm_gbs") if os.path.exists(peaks_file) else 6650.0
        line = {
            "metric": METRIC if args.config == "c1" else "generated events/s (ntried per second), " + workload.split(",")[0],
            "value": gen_per_s, "unit": "events/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "accepted_per_s": acc_per_s, "accepted_fraction": acc.nsuccess / tries_total,
            "config": {"workload": workload, "tries_per_step_per_gpu": n, "stage_batch": args.batch, "mode": args.mode,
                       "rng": "Philox4x32-10 keyed (seed, try index)",
                       "l2": "inputs are generated on chip; the per-stage state buffers (%.0f MB per batch) exceed the 126 MB L2"
                             % (110 * 8 * args.batch / 1e6)},
            "e2e": None if args.skip_e2e else
                   {"value": tries_total / (e2e_ms * 1e-3), "unit": "events/s", "h2d_bytes_per_step": 32,
                    "d2h_bytes_per_step": C.sizeof(Accum),
                    "note": "simc_b200_run() through the C ABI with host accumulators; a Monte Carlo step's only input "
                            "is (first_try, n_tries, seed)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "kernel": names[dom], "achieved": achieved, "peak": peak_muladd, "unit": "TFLOP/s",
                         "frac": achieved / peak_muladd if peak_muladd else None, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_kind": "measured in this run: DMUL+DADD microbenchmark (no FMA, like the strict arithmetic); "
                                      "DFMA peak %.1f TFLOP/s" % peak_fma,
                         "flops_per_generated_event": flops_per_try,
                         "whole_loop_tflops": flops / (dev_ms * 1e-3) / 1e12,
                         "whole_loop_frac": flops / (dev_ms * 1e-3) / 1e12 / peak_muladd if peak_muladd else None,
                         "model": "SURVEY 8(d): maps 25+14T / 25+12T per call (measured calls), hut %.0f flop + %.0f transc. "
                                  "per track in a hut, generation %.0f + %.0f per try, finish %.0f + %.0f (+ %.0f + %.0f "
                                  "peaked_rad_weight) per event through both arms; 1 transcendental = %.0f flops"
                                  % (HUT_FLOPS, HUT_TRANS, GEN_FLOPS, GEN_TRANS, FIN_FLOPS, FIN_TRANS, RADW_FLOPS, RADW_TRANS,
                                     FLOPS_PER_TRANS),
                         "stages": stages,
                         "stage_ms": dict(zip(names, stage_ms)), "stage_launches": dict(zip(names, stage_launches)),
                         "hbm_peak_gbs": hbm_peak},
            "clocks": sampler.summary(),
            "allreduce_ms_per_step": allreduce_ms / args.steps,
            "collective": "inside the e2e region, once per step: all_gather_into_tensor (NCCL) of the device accumulator "
                          "blocks + k_reduce_gathered, on the handle's stream" if world > 1 else "none (one GPU)",
            "yield_per_mC": acc.wtcontribute.value() / tries_total *
                            (1.0 / (cfg.targ.mass_amu / 3.75914e6 / (cfg.targ.abundancy / 100.) * abs(np.cos(cfg.targ.angle)) / (cfg.targ.thick * 1000.))) *
                            (cfg.gen.e.yptar.max - cfg.gen.e.yptar.min) * (cfg.gen.e.xptar.max - cfg.gen.e.xptar.min) * charge,
        }
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            n_cpu = args.cpu_tries if args.cpu_tries else 25000 * cores
            acc_c, dt = cpu_loop(cfg, n_cpu, 7, cores)
            line["cpu_baseline"] = {"value": acc_c.ntried / dt, "unit": "events/s", "cores": cores, "kind": "port",
                                    "accepted_per_s": acc_c.nsuccess / dt,
                                    "sample": f"{n_cpu} tries of the same deck on {cores} host threads (oracle restatement, not gfortran)"}
        print(json.dumps(line))
    sim.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
