#!/usr/bin/env python
"""bench.py — IF Msamples/s through 60-channel B1C (wide-band, data + QMBOC pilot) closed-loop
tracking on N B200s of one node (BASELINE.json metric, config 4), plus %-of-HBM roofline,
an end-to-end number through the public API with host buffers, and the CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* = one pass of the hot path over one batch: all channels tracked closed-loop over the
whole synthetic IF record (state reset between steps).  Channels are sharded round-robin over
the ranks (strong scaling: 60 channels in total), the IF record is replicated in every GPU's
HBM, and rank 0 gathers the packed correlator-output blocks with one NCCL gather per step.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

FS = 99.375e6
N_CHANNELS = 60
METRIC = "IF Msamples/s through 60-ch B1C tracking"
UNIT = "Msamples/s"

# tracking workloads: signal, reference function mirrored, samples per epoch, closed-loop Doppler range of the synthetic
# scenario (B2a gets no code-Doppler aiding: SURVEY 8(d)), self spectral separation coefficient (dB/Hz) for the expected
# C/N0 with the record's other satellites as noise, dominant kernel, CPU-baseline epochs per step
TRACK = {
    "track": dict(sig="B1C", mode="WB", spc=993750, max_doppler=4500.0, kappa_db=-64.8, kernel="trk_fw_kernel", cpu_epochs=2,
                  cn0_total=45.0,   # synth: data 11/44 + pilot 33/44 of the 45 dB-Hz carrier
                  what="WB tracking (data + QMBOC pilot)"),
    "track_nb": dict(sig="B1C", mode="NB", spc=993750, max_doppler=4500.0, kappa_db=-64.8, kernel="trk_fw_kernel", cpu_epochs=2,
                     cn0_total=44.59,  # data 11/44 + BOC(1,1) pilot 29/44 of the 45 dB-Hz carrier
                     what="NB tracking (data + BOC(1,1) pilot)"),
    "track_b2a": dict(sig="B2a", mode="B2a", spc=99375, max_doppler=100.0, kappa_db=-71.9, kernel="trk_b2a_unit_kernel",
                      cpu_epochs=20, what="tracking (data + pilot)",
                      cn0_total=48.01),   # synth: data and pilot at 45 dB-Hz EACH (equal amplitudes), B2a_CNo is their sum
}


def metric_name(args):
    w = TRACK[args.workload]
    return METRIC if args.workload == "track" else f"IF Msamples/s through {args.channels}-ch {w['sig']} tracking"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seconds", type=float, default=float(os.environ.get("BDS_BENCH_SECONDS", 30.0)),
                    help="length of the synthetic IF record (BASELINE config 4: 30 s)")
    ap.add_argument("--channels", type=int, default=N_CHANNELS)
    ap.add_argument("--fs", type=float, default=FS,
                    help="--workload track / acq_b1c: sampling rate of the synthetic record [Hz]; 99.375e6 = BASELINE config 4, "
                         "53e6 = the reference's shipped B1C setting (B1C/initSettings.m:57)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "general", "fast"])
    ap.add_argument("--workload", default="track", choices=["track", "track_nb", "track_b2a", "dual", "pipeline", "acq_b2a", "acq_b1c"],
                    help="track = the headline metric (BASELINE config 4; --channels 12 = config 3); track_nb = the same through "
                         "NB_tracking (data + BOC(1,1) pilot, pilotTRKflag 1); track_b2a = 60-channel "
                         "B2a tracking; dual = BASELINE config 5 (--channels B1C + --channels B2a channels co-scheduled on "
                         "the same GPUs); pipeline = config 5 as a joint run: 63-PRN acquisition of both bands (PRN-sharded), "
                         "preRun, then tracking of the channels it found; acq_b2a = BASELINE config 2 (B2a 63-PRN x +-5 kHz "
                         "acquisition grid); acq_b1c")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: --channels in total, sharded over the GPUs (BASELINE config 4); weak: --channels per GPU")
    ap.add_argument("--b2a-cluster", type=int, default=0, choices=[0, 1, 2, 4, 8],
                    help="B2a kernel: CTAs per channel (0 = library default / for --workload dual the largest cluster that leaves "
                         "half of the SMs to the B1C grid)")
    ap.add_argument("--gather-every", type=int, default=0,
                    help="epochs per launch + NCCL gather of the correlator outputs to rank 0 (0 = one gather per step, the "
                         "default; 1 = the per-epoch gather the north star words, measured for SURVEY 7.7)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-file", action="store_true",
                    help="skip the file-backed end-to-end leg (N = 1): the record read from a file like the reference's fid, "
                         "pageable result arrays - the path the MEX drop-in takes")
    return ap.parse_args()


def settings_b1c(n_ch, seconds, nb=False):
    import bds3_b200 as B
    return B.b1c.initSettings(samplingFreq=FS, numberOfChannels=n_ch, pilotTRKflag=1 if nb else 2,
                              msToProcess=int(round(seconds * 1000)))


def settings_for(workload, n_ch, seconds):
    import bds3_b200 as B
    if TRACK[workload]["sig"] == "B1C":
        return settings_b1c(n_ch, seconds, nb=TRACK[workload]["mode"] == "NB")
    return B.b2a.initSettings(numberOfChannels=n_ch, msToProcess=int(round(seconds * 1000)))


def measured_traffic(kernel, channels, seconds, world, key="dram_bytes_per_launch"):
    """DRAM bytes per launch of the dominant kernel (or, key="issue", its issue-slot figures) from the committed
    `ncu --set full` capture of this exact configuration (profiles/traffic.json), else None."""
    if FS != 99.375e6:
        return None                                     # the captures are of the BASELINE sampling rate
    try:
        for e in json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))):
            if (e["kernel"], e["channels"], e["seconds"], e["n_gpus"]) == (kernel, channels, seconds, world):
                return e.get(key)
    except Exception:
        pass
    return None


def acq_traffic(b1c, cells):
    """DRAM bytes of the inverse passes of one call: the per-cell figure of the committed capture (profiles/traffic.json)
    x the cells of the grid; else None"""
    try:
        for e in json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))):
            if e["kernel"] == ("acq_inv_b1c" if b1c else "acq_inv_b2a"):
                return int(e["dram_bytes_per_cell"] * cells)
    except Exception:
        pass
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """index of the next sample (call at the start / end of the timed region)"""
        return len(self.rows)

    def stop(self, i0=0, i1=None):
        """summary of the samples [i0, i1) = the ones taken during the timed region (the sampler is started before
        the warm-up so that nvidia-smi is already streaming when the timed region begins)"""
        if self.proc:
            self.proc.terminate()
        rows = self.rows[i0:i1] if i1 is not None else self.rows[i0:]
        if not rows:            # timed region shorter than one sampling period: the samples right before it (same load)
            rows = self.rows[max(0, i0 - 3):i0 + 1]
        self.rows = rows
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = []
        for i, name in ((4, "hw_slowdown"), (5, "hw_thermal_slowdown"), (6, "sw_thermal_slowdown"), (7, "sw_power_cap")):
            if any(len(r) >= 8 and r[i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement (C correlator + python loop closure), all host threads
# ----------------------------------------------------------------------------------------------
def cpu_track_sample(x, st, ch, n_epochs, threads, mode="WB"):
    """Tracks every channel of ``ch`` for n_epochs with the oracle on ``threads`` host threads.
    Returns wall seconds."""
    from concurrent.futures import ThreadPoolExecutor
    import bds_oracle as O
    import c_oracle
    c_oracle.lib()
    so = O.Settings(dict(st))
    so.numberOfChannels = 1
    if mode == "WB":
        O.CalcWeighingFactor(so)  # warm scipy import outside the timed region

    def one(c):
        O.tracking(mode, x, [O.Settings(dict(c))], so, n_epochs=n_epochs, correlator=c_oracle.correlate_epoch)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, ch))
    return time.perf_counter() - t0


def workload_string(channels, seconds, workload="track"):
    """config.workload shared by both arms (the reference arm times a bounded sample of it, see config.sample)"""
    w = TRACK[workload]
    return (f"{w['sig']} {channels}-channel {w['what']}, fs={FS / 1e6:g} MHz int8 IF, {seconds:g} s record")


def host_record_numpy(st, sats, n, sig="B1C"):
    """Host record for the reference arm, rendered with numpy and the ORACLE's code generators: nothing of the product
    (libbdsgpu.so) is loaded in a `--impl reference` process.  Cached under /tmp (rendering 60 satellites takes a while)."""
    import hashlib
    import numpy as np
    import bds_oracle as O
    from bds3_b200 import synth          # pure-python scenario / signal model; does not load the CUDA library
    key = hashlib.sha1(repr((sig, n, [(s_.PRN, s_.doppler, s_.codeDelay, s_.carrPhase, s_.amplitude) for s_ in sats])).encode()).hexdigest()[:16]
    path = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"bds_ref_record_{key}.npy")
    if os.path.exists(path):
        try:
            x = np.load(path)
            if x.size == n:
                return x
        except Exception:
            pass
    from concurrent.futures import ThreadPoolExecutor
    if sig == "B1C":
        codes = {s_.PRN: (O.b1c_data_primary(s_.PRN), O.b1c_pilot_primary(s_.PRN)) for s_ in sats}
    else:
        codes = {s_.PRN: (O.generateB2aDataCode(s_.PRN), O.generateB2aPilotCode(s_.PRN)) for s_ in sats}
    # numpy releases the GIL in its kernels: render blocks of the record on all host threads
    nthr = os.cpu_count() or 1
    blk = max(1 << 16, (n + nthr - 1) // nthr)
    parts = [(o, min(blk, n - o)) for o in range(0, n, blk)]
    with ThreadPoolExecutor(max_workers=nthr) as ex:
        xs = list(ex.map(lambda a: synth.synth_numpy(sig, st, sats, a[1], first_sample=a[0], chunk=1 << 18,
                                                     noise_seed=7919 + a[0], primary_codes=lambda prn: codes[prn]), parts))
    x = np.concatenate(xs)
    try:
        np.save(path, x)
    except Exception:
        pass
    return x


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  MATLAB cannot run here,
    so this is the oracle port (kind "port"), all host threads, a bounded sample per step.
    Nothing of the product is executed: the input is rendered with numpy + the oracle's code generators."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from bds3_b200 import synth
    from bds3_b200.settings import Settings
    import bds_oracle as O
    threads = os.cpu_count() or 1
    if args.workload == "dual":
        return run_reference_dual(args)
    if args.workload not in TRACK:
        print(json.dumps({"impl": "reference", "unavailable": f"no reference arm for workload {args.workload}"}))
        return
    w = TRACK[args.workload]
    n_ep, spc, mode = w["cpu_epochs"], w["spc"], w["mode"]
    ms = int(round(args.seconds * 1000))
    if w["sig"] == "B1C":
        st = Settings(dict(O.initSettings_B1C(samplingFreq=FS, numberOfChannels=args.channels, pilotTRKflag=1 if mode == "NB" else 2,
                                              msToProcess=ms)))
    else:
        st = Settings(dict(O.initSettings_B2a(numberOfChannels=args.channels, msToProcess=ms)))
    sats = synth.make_sats(args.channels, st, w["sig"], max_doppler=w["max_doppler"])
    n = int((n_ep + 1.2) * spc) + spc
    x = host_record_numpy(st, sats, n, w["sig"])
    ch = synth.channels_from_sats(sats, st, w["sig"], freq_error=2.0)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_track_sample(x, st, ch[:threads], 1, threads, mode)
    times = []
    for _ in range(args.steps):
        times.append(cpu_track_sample(x, st, ch, n_ep, threads, mode))
    t = sum(times) / len(times)
    val = n_ep * spc / t / 1e6
    sample = (f"{args.channels} channels x {n_ep} epochs ({spc / FS * 1e3:g} ms each) per step; Msamples/s = IF samples the "
              "channels advanced through / wall time, the same normalisation as the GPU arm (which runs every epoch of the record)")
    line = {"impl": "reference", "metric": metric_name(args), "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(args.channels, args.seconds, args.workload), "sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{args.channels} ch x {n_ep} epochs per step; float64 oracle restatement "
                                       f"of {'WB_tracking.m' if mode == 'WB' else 'B2a tracking.m'} (C correlator + python loop "
                                       "closure); MATLAB unavailable"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "native_so_loaded": repo_native_libraries()}
    print(json.dumps(line))


def run_reference_dual(args):
    """--impl reference --workload dual: the two CPU ports one after the other on 20 ms of each band's record."""
    from bds3_b200 import synth
    from bds3_b200.settings import Settings
    import bds_oracle as O
    threads = os.cpu_count() or 1
    ms = int(round(args.seconds * 1000))
    parts = []
    for wl in ("track", "track_b2a"):
        w = TRACK[wl]
        if w["sig"] == "B1C":
            st = Settings(dict(O.initSettings_B1C(samplingFreq=FS, numberOfChannels=args.channels, pilotTRKflag=2, msToProcess=ms)))
        else:
            st = Settings(dict(O.initSettings_B2a(numberOfChannels=args.channels, msToProcess=ms)))
        sats = synth.make_sats(args.channels, st, w["sig"], max_doppler=w["max_doppler"])
        n = int((w["cpu_epochs"] + 1.2) * w["spc"]) + w["spc"]
        parts.append((w, st, host_record_numpy(st, sats, n, w["sig"]), synth.channels_from_sats(sats, st, w["sig"], freq_error=2.0)))
    times = []
    for i in range(max(0, min(args.warmup, 1)) + args.steps):
        t = sum(cpu_track_sample(x, st, ch, w["cpu_epochs"], threads, w["mode"]) for w, st, x, ch in parts)
        if i >= max(0, min(args.warmup, 1)):
            times.append(t)
    t = sum(times) / len(times)
    val = sum(w["cpu_epochs"] * w["spc"] for w, _, _, _ in parts) / t / 1e6
    sample = (f"{args.channels} B1C channels x 2 epochs + {args.channels} B2a channels x 20 epochs (20 ms of each band) per step; "
              "Msamples/s = IF samples both bands advanced through / wall time, the GPU arm's normalisation")
    print(json.dumps({"impl": "reference", "metric": f"IF Msamples/s through {2 * args.channels}-ch dual-band ({args.channels} B1C + {args.channels} B2a) tracking",
                      "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                      "dtype": "f64", "data": "synthetic",
                      "config": {"workload": f"BASELINE config 5: {args.channels} B1C WB + {args.channels} B2a channels co-scheduled, two int8 IF "
                                             f"records at 99.375 MHz, {args.seconds:g} s", "sample": sample},
                      "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
                      "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "native_so_loaded": repo_native_libraries()}))


def repo_native_libraries():
    """shared objects of THIS repo mapped into the process (the reference arm must list the oracle only)"""
    out = set()
    try:
        for ln in open("/proc/self/maps"):
            path = ln.split()[-1]
            if path.startswith(ROOT) and ".so" in os.path.basename(path):
                out.add(os.path.relpath(path, ROOT))
    except OSError:
        pass
    return sorted(out)


# ----------------------------------------------------------------------------------------------
def self_check(sess, st, chans, n_epochs, x_dev, n_samples, mode="WB", injected_cn0=45.0, n_sampled_channels=3,
               n_sampled_epochs=20, kappa_db=-64.8, seconds_tail=10.0):
    """Untimed checks on the result of the timed run (VERDICT r1 'make the headline run prove its own correctness'):
      * every channel completed every epoch;
      * the loop bookkeeping of EVERY epoch of EVERY channel follows from the device's own discriminators through the
        reference's loop filters / block arithmetic in float64 (tests/util.replay_loop_chain);
      * lock over the last 10 s: DataPLD and PilotPLD > 0.9 (Calc_CNo_PLD.m:70-73), total C/N0 (WB_tracking.m:467-477)
        in the band expected for the injected C/N0 with the record's other satellites acting as noise;
      * sampled one-step parity at the END of the record: the 18 sums of 3 channels x 20 epochs against the float64
        oracle correlator evaluated at the device's NCO state, <= 1e-4 of max(|I_P|,|Q_P|) (tests/util);
      * share of chips that left the chip-synchronous body for the exact per-sample path < 1e-4."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    import bds_oracle as O
    so = O.Settings(dict(st))
    N = n_epochs
    planes = sess.fetch(N, raw=True)
    done = planes["epochsDone"]
    assert int(done.min()) == N, f"tracking stopped early: epochsDone min {int(done.min())} of {N}"
    act = [c for c in chans if c.PRN != 0]
    assert len(act) == len(chans)
    mem = util.replay_loop_chain(mode, so, act, planes, N)                      # [N, 4, nch]
    lock = util.lock_report(mode, so, planes, N, seconds_tail=seconds_tail)
    # effective C/N0 with K-1 equal-power interferers: C / (N0 + (K-1) C kappa), kappa = the signal's self spectral
    # separation coefficient (BOC(1,1): -64.8 dB/Hz, BPSK(10): -71.9 dB/Hz); the estimator itself scatters by about
    # +-1 dB over an interval
    K = int(st.numberOfChannels_total) if "numberOfChannels_total" in st else len(act)
    cn = 10 ** (injected_cn0 / 10)
    eff = 10 * math.log10(1.0 / (1.0 / cn + max(0, K - 1) * 10 ** (kappa_db / 10)))
    lo, hi = eff - 1.5, injected_cn0 + 1.0
    locked = (lock["data_pld_min"] > 0.9) & (lock["pilot_pld_min"] > 0.9) & (lock["cno_median"] > lo) & (lock["cno_median"] < hi)
    assert bool(locked.all()), ("channels out of lock", np.nonzero(~locked)[0].tolist(), lock["cno_median"].round(2).tolist(),
                                lock["data_pld_min"].round(3).tolist())

    def get_block(pos, n):
        return x_dev[pos: pos + n].cpu().numpy()

    pick = sorted(set(int(round(i)) for i in np.linspace(0, len(act) - 1, min(n_sampled_channels, len(act)))))
    epochs = list(range(max(0, N - n_sampled_epochs), N))
    worst = 0.0
    for c in pick:
        g = {k: planes[k][c] for k in planes if k not in ("raw", "epochsDone") and planes[k].ndim == 2}
        worst = max(worst, util.one_step_parity_sampled(mode, so, get_block, act[c], g, planes["raw"][c], mem[:, :, c], epochs))
    fast, exact, general, _ = sess.counters()
    assert general == 0 or fast + exact == 0, "the general kernel ran in a chip-synchronous session"   # (--kernel general: all general)
    frac = exact / max(1, fast + exact)
    assert frac < 1e-4, f"exact-path share {frac:.2e}"
    return {"channels": len(act), "locked_channels": int(locked.sum()), "pld_min": float(min(lock["data_pld_min"].min(), lock["pilot_pld_min"].min())),
            "cno_db_min": float(lock["cno_median"].min()), "cno_db_max": float(lock["cno_median"].max()),
            "cno_db_expected": [round(lo, 2), round(hi, 2)], "cno_points": int(lock["points"]),
            "loop_chain_replayed_epochs": int(N * len(act)), "parity_max_rel": worst, "parity_epochs": len(pick) * len(epochs),
            "parity_channels": [int(act[c].PRN) for c in pick], "parity_tolerance": 1e-4,
            "fast_chips": int(fast), "exact_chips": int(exact), "exact_chip_frac": frac}


# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import bds3_b200 as B
    from bds3_b200 import _lib as L, _shard, _track, synth

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L.init(local)
    kern = {"auto": L.KERNEL_AUTO, "general": L.KERNEL_GENERAL, "fast": L.KERNEL_FAST}[args.kernel]

    w = TRACK[args.workload]
    sig, mode, spc = w["sig"], w["mode"], w["spc"]
    # strong scaling (BASELINE config 4): --channels in total over the ranks; weak: --channels on every rank
    total_channels = args.channels * (world if args.scaling == "weak" else 1)
    if total_channels > 63:
        raise SystemExit("the synthetic scenario has one satellite per PRN: at most 63 channels in total")
    st = settings_for(args.workload, total_channels, args.seconds)
    sats = synth.make_sats(total_channels, st, sig, max_doppler=w["max_doppler"])
    chans = synth.channels_from_sats(sats, st, sig, freq_error=2.0)
    n_samples = int(round(args.seconds * FS))
    n_epochs = max(1, int(math.floor((n_samples - spc) / (spc * (1 + 1e-5)))) - 1)
    # ---- synthetic IF, generated straight into HBM (identical on every rank: replicated record)
    x_dev = torch.empty(n_samples + 64, dtype=torch.int8, device="cuda")
    synth.synth_device(sig, st, sats, n_samples, out_ptr=x_dev.data_ptr())
    torch.cuda.synchronize()
    # ---- channel shard of this rank (round robin)
    mine = _shard.shard_list(chans, rank, world)
    st_local = st.copy()
    st_local.numberOfChannels = len(mine)
    st_local.numberOfChannels_total = total_channels   # satellites in the record (self_check: multiple-access noise)
    tuning = {"b2aClusterSize": args.b2a_cluster} if (sig == "B2a" and args.b2a_cluster) else None
    if os.environ.get("BDS_BENCH_TRK_TUNING"):        # developer A/B of bds_trk_cfg tuning fields: "fwPassesPerTask=3,fwPrefetch=2"
        tuning = dict(tuning or {}, **{k: int(v) for k, v in (kv.split("=") for kv in os.environ["BDS_BENCH_TRK_TUNING"].split(","))})
    sess = _track.TrackSession(mode, st_local, mine, kernel=kern, device_ptr=x_dev.data_ptr(), n_samples=n_samples, tuning=tuning)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def _as_tensor(ptr, n):
        class _Arr:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_Arr(), device="cuda")

    def gather_block(s_=None):
        """one NCCL gather of the packed correlator outputs per step (rank 0 receives)"""
        if dist is None:
            return None
        p, nbytes, nf, cap = (s_ or sess).device_block()
        t = _as_tensor(p, nbytes // 8)   # the library's device block, wrapped without a copy
        return _shard.gather_blocks(t, max_block_elems, dist, dst=0)

    # size of the largest rank block (ranks differ by at most one channel)
    sess.run_async(n_epochs)
    sess.sync()
    _, nbytes0, _, _ = sess.device_block()
    max_block_elems = _shard.max_block_elems(nbytes0 // 8, dist)

    gather_buf = {}

    def gather_epochs(e0, k):
        """NCCL gather of epochs [e0, e0 + k) of every output plane (21 trackResults planes + 18 raw sums) to rank 0"""
        if dist is None:
            return
        p, nbytes, nf, cap = sess.device_block()
        nch = nbytes // 8 // (nf * cap)
        nmax = max_block_elems // (nf * cap)
        if k not in gather_buf:
            gather_buf[k] = (torch.zeros(nmax * nf * k, dtype=torch.float64, device="cuda"),
                             [torch.empty(nmax * nf * k, dtype=torch.float64, device="cuda") for _ in range(world)] if rank == 0 else None)
        buf, outs = gather_buf[k]
        buf[: nch * nf * k].view(nch, nf, k).copy_(_as_tensor(p, nbytes // 8).view(nch, nf, cap)[:, :, e0: e0 + k])
        dist.gather(buf, outs, dst=0)

    def step():
        sess.reset()
        if args.gather_every <= 0:
            sess.run_async(n_epochs)
            sess.sync()
            gather_block()
            return sess.stats()
        e, ms_total = 0, 0.0
        while e < n_epochs:     # one launch + one gather per --gather-every epochs
            k = min(args.gather_every, n_epochs - e)
            sess.run_async(k)
            cs, ep, ms = sess.stats()
            ms_total += ms
            gather_epochs(e, k)
            e += k
        return cs, ep, ms_total

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    launches0 = B.launch_count()
    barrier()
    s0 = sampler.mark()
    t0 = time.perf_counter()
    kernel_ms, ch_samples = [], 0
    for _ in range(args.steps):
        cs, ep, ms = step()
        kernel_ms.append(ms)
        ch_samples = cs
    barrier()
    wall = time.perf_counter() - t0
    s1 = sampler.mark()
    launches = B.launch_count() - launches0
    clocks = sampler.stop(s0, s1) if rank == 0 else None
    # ---- untimed: the run that was just timed proves its own correctness (every rank, its own channels)
    check = self_check(sess, st_local, mine, n_epochs, x_dev, n_samples, mode=mode, kappa_db=w["kappa_db"],
                       injected_cn0=w["cn0_total"])
    if dist is not None:
        agg = torch.tensor([check["locked_channels"], check["channels"], check["fast_chips"], check["exact_chips"],
                            check["parity_epochs"]], device="cuda", dtype=torch.float64)
        mx = torch.tensor([check["parity_max_rel"], -check["cno_db_min"], check["cno_db_max"], -check["pld_min"]],
                          device="cuda", dtype=torch.float64)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        check.update(locked_channels=int(agg[0]), channels=int(agg[1]), fast_chips=int(agg[2]), exact_chips=int(agg[3]),
                     parity_epochs=int(agg[4]), parity_max_rel=float(mx[0]), cno_db_min=-float(mx[1]),
                     cno_db_max=float(mx[2]), pld_min=-float(mx[3]))
        check["exact_chip_frac"] = check["exact_chips"] / max(1, check["fast_chips"] + check["exact_chips"])
    # device time: max over ranks of the CUDA-event time of the persistent kernel, per step
    dev_ms = sum(kernel_ms) / len(kernel_ms)
    tmax = torch.tensor([dev_ms, wall * 1e3 / args.steps], device="cuda", dtype=torch.float64)
    tot = torch.tensor([float(ch_samples)], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms_max, step_ms_max = float(tmax[0]), float(tmax[1])
    total_ch_samples = float(tot[0])
    if_samples = n_epochs * float(spc)  # IF samples the channels advanced through (nominal epoch length)
    value = if_samples / (step_ms_max * 1e-3) / 1e6

    # ---- e2e: the public host-buffer call: pinned host IF -> (chunked H2D overlapped with tracking) -> D2H of the
    #      trackResults planes, every step, all inside the timed region
    e2e = None
    if not args.no_e2e:
        x_host = torch.empty(n_samples, dtype=torch.int8).pin_memory()
        x_host.copy_(x_dev[:n_samples])
        torch.cuda.synchronize()
        tuning2 = dict(tuning or {})
        if world > 1 and sig == "B1C" and args.kernel != "general":
            # leave a few SMs to the NCCL all-gather kernels that bring the next chunk in (a persistent grid on every SM
            # serialises them between its launches; measured on 8 x B200: 41.2 -> 36.0 ms per step, profiles/r02/scaling.md)
            n_sms_ = torch.cuda.get_device_properties(local).multi_processor_count
            tuning2["fwMaxCtas"] = int(os.environ.get("BDS_BENCH_E2E_FW_CTAS", n_sms_ - 12))
        sess2 = _track.TrackSession(mode, st_local, mine, kernel=kern, tuning=tuning2 or None)     # no resident record: fed from the host
        # caller-owned result planes, pinned like the input (the MEX gateway would hand mxArrays here)
        res = {name: torch.empty((len(mine), n_epochs), dtype=torch.float64).pin_memory().numpy() for name in L.TRK_PLANES}
        times, parts = [], []
        d2h = 0
        if world > 1:
            # N GPUs: every rank uploads 1/N of each chunk from its pinned copy of the record and the ranks all-gather the
            # chunk over NVLink (NCCL) on a side stream; the tracking kernel of chunk i runs under the upload + gather of
            # chunk i+1.  PCIe carries n_samples / N bytes per rank instead of the whole replicated record.
            chunk = (128 << 20) // (16 * world) * (16 * world)
            n_pad = (n_samples + chunk - 1) // chunk * chunk
            xh = torch.zeros(n_pad, dtype=torch.int8).pin_memory()
            xh[:n_samples].copy_(x_host)
            x_full = torch.empty(n_pad + 64, dtype=torch.int8, device="cuda")
            stage = [torch.empty(chunk // world, dtype=torch.int8, device="cuda") for _ in range(2)]
            side = torch.cuda.Stream()
            n_chunks = n_pad // chunk
            # Exchange of the uploaded slices between the GPUs.  Preferred: copy-engine pushes into the peers' records
            # (CUDA IPC mappings, device-to-device copies over NVLink) - they need no SM, so they run under the persistent
            # tracking grid, which occupies every SM and serialises an NCCL all-gather kernel between its launches.
            # A host-side (gloo) barrier per chunk tells every rank that all pushes of that chunk have landed.
            peers, cpu_group, e2e_exchange = None, None, "NCCL all-gather"
            try:
                # measured on 2 x B200 (profiles/r02/scaling_n2.md): the pushes into IPC-mapped peer memory ran at a fraction of
                # the NVLink rate (117 ms per step against 46 ms with the all-gather), so NCCL stays the default
                if os.environ.get("BDS_BENCH_E2E_EXCHANGE", "nccl") == "peer":
                    import torch.multiprocessing.reductions as red
                    fn, a = red.reduce_tensor(x_full)
                    handles = [None] * world
                    dist.all_gather_object(handles, (fn, a))
                    peers = [x_full if r == rank else h_[0](*h_[1]) for r, h_ in enumerate(handles)]
                    cpu_group = dist.new_group(backend="gloo")
                    probe = torch.full((16,), rank, dtype=torch.int8, device="cuda")
                    for r in range(world):          # one small push to every peer: fails here rather than in the timed loop
                        if r != rank:
                            peers[r][n_pad + 16 * 0: n_pad + 16].copy_(probe, non_blocking=True)
                    torch.cuda.synchronize()
                    dist.barrier(group=cpu_group)
                    e2e_exchange = "copy-engine peer pushes (CUDA IPC) + gloo barrier per chunk"
            except Exception as ex:   # noqa: BLE001
                print(f"[bench] peer-copy exchange unavailable ({type(ex).__name__}: {ex}); using the NCCL all-gather", file=sys.stderr)
                peers, cpu_group = None, None

            def e2e_step():
                evs = []
                with torch.cuda.stream(side):
                    for k in range(n_chunks):
                        o, sl = k * chunk, chunk // world
                        if peers is not None:
                            mine_ = x_full[o + rank * sl: o + (rank + 1) * sl]
                            mine_.copy_(xh[o + rank * sl: o + (rank + 1) * sl], non_blocking=True)
                            for r in range(1, world):      # staggered so that not everybody pushes to the same GPU at once
                                q = (rank + r) % world
                                peers[q][o + rank * sl: o + (rank + 1) * sl].copy_(mine_, non_blocking=True)
                        else:
                            st_ = stage[k & 1]
                            st_.copy_(xh[o + rank * sl: o + (rank + 1) * sl], non_blocking=True)
                            dist.all_gather_into_tensor(x_full[o: o + chunk], st_)
                        ev = torch.cuda.Event()
                        ev.record(side)
                        evs.append(ev)
                for k, ev in enumerate(evs):
                    ev.synchronize()
                    if peers is not None:
                        dist.barrier(group=cpu_group)   # every rank's pushes of chunk k have completed
                    sess2.run_window(x_full.data_ptr(), min(n_samples, (k + 1) * chunk), n_epochs)
        else:
            def e2e_step():
                sess2.run_streamed(x_host.data_ptr(), n_samples, n_epochs)
        for i in range(2 + args.steps):
            barrier()
            t1 = time.perf_counter()
            sess2.reset()
            t2 = time.perf_counter()
            e2e_step()
            t3 = time.perf_counter()
            sess2.sync()
            t4 = time.perf_counter()
            planes = sess2.fetch(n_epochs, into=res)
            gather_block(sess2)
            barrier()
            t5 = time.perf_counter()
            if i >= 2:
                times.append(t5 - t1)
                parts.append((t2 - t1, t3 - t2, t4 - t3, t5 - t4))
            d2h = sum(v.nbytes for v in planes.values())
        # plain pinned H2D rate of this box, for context (the streamed path cannot beat it)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        x_dev[:n_samples].copy_(x_host, non_blocking=True)
        torch.cuda.synchronize()
        h2d_gbs = n_samples / (time.perf_counter() - t1) / 1e9
        assert int(planes["epochsDone"].min()) == n_epochs, "e2e run did not complete every epoch"
        sess2.close()
        # ---- the MEX drop-in's path (matlab/bds_mex.c do_track): the record comes from a FILE (bds_track_open_file: mmap,
        #      pageable, only the range the epochs touch is streamed under the kernel) and the result planes go into
        #      ordinary pageable arrays.  Per step: open, run, fetch, close.  N = 1 only (a shared file needs no sharding).
        e2e_file = None
        if world == 1 and not args.no_e2e_file:
            import tempfile
            d = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
            path = os.path.join(d, f"bds_bench_if_{os.getpid()}.bin")
            try:
                x_host.numpy().tofile(path)
                tf = []
                for i in range(1 + min(3, args.steps)):
                    t1 = time.perf_counter()
                    with _track.TrackSession(mode, st_local, mine, source=path, kernel=kern, tuning=tuning) as sf:
                        sf.run_async(n_epochs)
                        pf = sf.fetch(n_epochs)
                    if i >= 1:
                        tf.append(time.perf_counter() - t1)
                    assert int(pf["epochsDone"].min()) == n_epochs, "file-backed e2e run did not complete every epoch"
                tfm = sum(tf) / len(tf)
                e2e_file = {"value": if_samples / tfm / 1e6, "unit": UNIT, "ms_per_step": tfm * 1e3,
                            "path": "bds_track_open_file (page-cache resident file read by parallel pread()s through pinned staging "
                                    "buffers) + bds_track_run_async + bds_track_fetch into pageable arrays + bds_track_close, every step"}
            except (OSError, AssertionError, L.BdsError) as ex:     # e.g. no room for the 3 GB file: the leg is optional
                e2e_file = {"unavailable": f"{type(ex).__name__}: {ex}"[:200]}
            finally:
                try:
                    os.unlink(path)
                except OSError:
                    pass
        te = sum(times) / len(times)
        tt = torch.tensor([te], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": if_samples / float(tt[0]) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(n_samples),   # summed over ranks
               "d2h_bytes_per_step": int(d2h), "ms_per_step": float(tt[0]) * 1e3,
               "ms_breakdown": {k: round(1e3 * sum(p_[j] for p_ in parts) / len(parts), 2)
                                for j, k in enumerate(("reset", "enqueue", "h2d+kernels", "fetch+gather"))},
               "fw_ctas": tuning2.get("fwMaxCtas", 0), "file_backed": e2e_file,
               "pinned_h2d_GBps_this_box": round(h2d_gbs, 1),
               "path": ("bds_track_run_streamed (128 MiB chunks on a copy stream) + bds_track_fetch" if world == 1 else
                        f"per 128 MiB chunk: H2D of 1/{world} per rank + {e2e_exchange} over NVLink on a side stream, "
                        "bds_track_run_window per chunk, bds_track_fetch")}

    if rank == 0:
        peak, peak_src = peaks()
        # algorithmic bytes: 1 byte per channel-sample; dominant kernel = trk_fw_kernel
        per_launch_bytes = total_ch_samples / world  # per-GPU launch
        achieved = per_launch_bytes / (dev_ms_max * 1e-3) / 1e9
        kname = w["kernel"] if args.kernel != "general" else "trk_persistent_kernel"
        line = {"metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": step_ms_max, "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f32 accumulate / f64 loop closure (int8 IF)", "data": "synthetic",
                "config": {"workload": workload_string(total_channels, args.seconds, args.workload),
                           "epochs_per_channel": n_epochs,
                           "parallelism": f"{total_channels} channels round-robin over {world} GPU(s), IF replicated",
                           "l2": f"input {n_samples / 1e6:.0f} MB > 126 MB L2 (no flush needed)",
                           "kernel": args.kernel, "x_realtime": value / (FS / 1e6),
                           "gather": ("one NCCL gather of the packed output block per step" if args.gather_every <= 0 else
                                      f"one launch + one NCCL gather of the outputs every {args.gather_every} epoch(s)")},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "frac_of_nominal_8000": achieved / 8000.0,
                             "traffic": measured_traffic(kname, args.channels, args.seconds, world),
                             "peak_source": peak_src,
                             "kernel": kname,
                             "kernel_ms_per_launch": dev_ms_max,
                             "algorithmic_bytes_per_launch": per_launch_bytes,
                             # what ncu shows as the limiter: instruction issue (L2 serves 98 % of the bytes, DRAM is < 2 %
                             # busy); figures of the committed capture of this configuration, None for other configurations
                             "measured_limiter": "instruction issue",
                             "issue": measured_traffic(kname, args.channels, args.seconds, world, key="issue")},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                # correctness of the timed run itself (untimed checks after the timed region, see self_check)
                "parity_max_rel": check["parity_max_rel"], "locked_channels": check["locked_channels"],
                "exact_chip_frac": check["exact_chip_frac"], "self_check": check}
        if not args.no_cpu_baseline and world >= 1:
            threads = os.cpu_count() or 1
            n_ep = w["cpu_epochs"]
            nh = int((n_ep + 1.2) * spc) + spc
            xh = x_dev[:nh].cpu().numpy()
            t = cpu_track_sample(xh, st, chans, n_ep, threads, mode)
            line["cpu_baseline"] = {"value": n_ep * spc / t / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{total_channels} ch x {n_ep} epochs, float64 oracle restatement "
                                              "(C correlator), all host threads; MATLAB unavailable"}
        print(json.dumps(line))
    sess.close()
    if dist is not None:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
# BASELINE config 5: dual band.  The reference has no joint B1C/B2a processing - they are two programs sharing Common/
# (B1C/init.m:42-44, B2a/init.m:38-40; SURVEY 8 mismatch 3) - so this is two independent pipelines co-scheduled on the
# same GPUs: per rank one B1C wide-band session (chip-synchronous persistent grid, limited to the SMs the B2a channels
# leave free) and one B2a session (one CTA per channel), each on its own stream over its own IF record.
# ----------------------------------------------------------------------------------------------
def run_dual(args):
    import numpy as np
    import torch
    import bds3_b200 as B
    from bds3_b200 import _lib as L, _shard, _track, synth

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L.init(local)
    n_sms = torch.cuda.get_device_properties(local).multi_processor_count
    n_samples = int(round(args.seconds * FS))
    total = args.channels * (world if args.scaling == "weak" else 1)      # per band
    if total > 63:
        raise SystemExit("the synthetic scenario has one satellite per PRN: at most 63 channels per band")
    bands = []
    for wl in ("track_b2a", "track"):            # B2a first: its CTA count bounds the B1C grid
        w = TRACK[wl]
        st = settings_for(wl, total, args.seconds)
        sats = synth.make_sats(total, st, w["sig"], max_doppler=w["max_doppler"])
        chans = synth.channels_from_sats(sats, st, w["sig"], freq_error=2.0)
        x_dev = torch.empty(n_samples + 64, dtype=torch.int8, device="cuda")
        synth.synth_device(w["sig"], st, sats, n_samples, out_ptr=x_dev.data_ptr())
        mine = _shard.shard_list(chans, rank, world)
        st_local = st.copy()
        st_local.numberOfChannels = len(mine)
        st_local.numberOfChannels_total = total
        n_epochs = max(1, int(math.floor((n_samples - w["spc"]) / (w["spc"] * (1 + 1e-5)))) - 1)
        if wl == "track_b2a":   # B2a: a cluster of CTAs per channel, on at most half of the SMs
            cs = args.b2a_cluster
            if cs == 0:
                cs = 8
                while cs > 1 and len(mine) * cs > n_sms // 2:
                    cs //= 2
            tuning = {"b2aClusterSize": cs}
            b2a_ctas = len(mine) * cs
        else:                   # the persistent B1C grid takes the SMs the B2a clusters leave
            tuning = {"fwMaxCtas": max(16, n_sms - b2a_ctas)}
        bands.append(dict(wl=wl, w=w, st=st, st_local=st_local, chans=chans, mine=mine, x_dev=x_dev, n_epochs=n_epochs,
                          tuning=tuning,
                          sess=_track.TrackSession(w["mode"], st_local, mine, device_ptr=x_dev.data_ptr(), n_samples=n_samples,
                                                   tuning=tuning)))
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def _as_tensor(ptr, n):
        class _Arr:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_Arr(), device="cuda")

    for b in bands:   # sizing run (allocates the output blocks)
        b["sess"].run_async(b["n_epochs"])
    for b in bands:
        b["sess"].sync()
        b["max_block"] = _shard.max_block_elems(b["sess"].device_block()[1] // 8, dist)

    def gather(b, s_=None):
        if dist is None:
            return
        p, nbytes, _, _ = (s_ or b["sess"]).device_block()
        _shard.gather_blocks(_as_tensor(p, nbytes // 8), b["max_block"], dist, dst=0)

    def step():
        for b in bands:
            b["sess"].reset()
        for b in bands:
            b["sess"].run_async(b["n_epochs"])       # both launches are asynchronous: the two kernels run concurrently
        for b in bands:
            b["sess"].sync()
        for b in bands:
            gather(b)
        return [b["sess"].stats() for b in bands]

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    launches0 = B.launch_count()
    barrier()
    s0 = sampler.mark()
    t0 = time.perf_counter()
    stats = None
    kms = [[], []]
    for _ in range(args.steps):
        stats = step()
        for i in range(2):
            kms[i].append(stats[i][2])
    barrier()
    wall = time.perf_counter() - t0
    s1 = sampler.mark()
    launches = B.launch_count() - launches0
    clocks = sampler.stop(s0, s1) if rank == 0 else None
    checks = [self_check(b["sess"], b["st_local"], b["mine"], b["n_epochs"], b["x_dev"], n_samples, mode=b["w"]["mode"],
                         kappa_db=b["w"]["kappa_db"], injected_cn0=b["w"]["cn0_total"]) for b in bands]
    agg = torch.tensor([sum(c["locked_channels"] for c in checks), sum(c["channels"] for c in checks),
                        float(stats[0][0] + stats[1][0])], device="cuda", dtype=torch.float64)
    mx = torch.tensor([max(c["parity_max_rel"] for c in checks), max(c["exact_chip_frac"] for c in checks),
                       wall * 1e3 / args.steps, sum(kms[0]) / len(kms[0]), sum(kms[1]) / len(kms[1])],
                      device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    step_ms = float(mx[2])
    if_samples = sum(b["n_epochs"] * float(b["w"]["spc"]) for b in bands)    # IF samples advanced through, both bands
    value = if_samples / (step_ms * 1e-3) / 1e6

    e2e = None
    if not args.no_e2e and world == 1:
        # both records from pinned host memory, streamed under the two kernels; result planes of both bands back to the host
        hosts, sess2, res = [], [], []
        for b in bands:
            xh = torch.empty(n_samples, dtype=torch.int8).pin_memory()
            xh.copy_(b["x_dev"][:n_samples])
            hosts.append(xh)
            sess2.append(_track.TrackSession(b["w"]["mode"], b["st_local"], b["mine"], tuning=b["tuning"]))
            res.append({name: torch.empty((len(b["mine"]), b["n_epochs"]), dtype=torch.float64).pin_memory().numpy()
                        for name in L.TRK_PLANES})
        torch.cuda.synchronize()
        times, d2h = [], 0
        for i in range(2 + args.steps):
            barrier()
            t1 = time.perf_counter()
            for s_ in sess2:
                s_.reset()
            for s_, xh, b in zip(sess2, hosts, bands):
                s_.run_streamed(xh.data_ptr(), n_samples, b["n_epochs"])
            d2h = 0
            for s_, b, r_ in zip(sess2, bands, res):
                planes = s_.fetch(b["n_epochs"], into=r_)
                assert int(planes["epochsDone"].min()) == b["n_epochs"], "e2e run did not complete every epoch"
                d2h += sum(v.nbytes for v in planes.values())
            barrier()
            if i >= 2:
                times.append(time.perf_counter() - t1)
        for s_ in sess2:
            s_.close()
        te = sum(times) / len(times)
        e2e = {"value": if_samples / te / 1e6, "unit": UNIT, "h2d_bytes_per_step": 2 * int(n_samples), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": te * 1e3, "path": "bds_track_run_streamed on both sessions (two pinned host records, chunked H2D "
                                                "on two copy streams under the two kernels) + bds_track_fetch of both"}
    if rank == 0:
        peak, peak_src = peaks()
        alg = float(agg[2]) / world           # channel-samples of both bands per GPU and step
        line = {"metric": f"IF Msamples/s through {2 * total}-ch dual-band ({total} B1C + {total} B2a) tracking",
                "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f32 accumulate / f64 loop closure (int8 IF)", "data": "synthetic",
                "config": {"workload": f"BASELINE config 5: {total} B1C WB + {total} B2a channels co-scheduled, two int8 IF records "
                                       f"at 99.375 MHz, {args.seconds:g} s",
                           "epochs_per_channel": {b["w"]["sig"]: b["n_epochs"] for b in bands},
                           "parallelism": f"channels of both bands round-robin over {world} GPU(s), both IF records replicated; per GPU "
                                          "two sessions on two streams, the B1C persistent grid limited to the SMs the B2a clusters leave",
                           "sm_split": {"b2a_ctas": b2a_ctas, "b2a_cluster": bands[0]["tuning"]["b2aClusterSize"],
                                        "b1c_ctas": bands[1]["tuning"]["fwMaxCtas"]},
                           "l2": f"inputs 2 x {n_samples / 1e6:.0f} MB > 126 MB L2 (no flush needed)",
                           "x_realtime": step_ms and args.seconds * 1e3 / step_ms},
                "roofline": {"bound": "hbm", "achieved": alg / (step_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": alg / (step_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                             "kernel": "trk_b2a_unit_kernel || trk_fw_kernel (concurrent; the step lasts as long as the slower one)",
                             "kernel_ms_per_launch": {"trk_b2a_unit_kernel": float(mx[3]), "trk_fw_kernel": float(mx[4])},
                             "algorithmic_bytes_per_launch": alg, "measured_limiter": "loop-closure latency of the B2a channels"},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "parity_max_rel": float(mx[0]), "locked_channels": int(agg[0]), "exact_chip_frac": float(mx[1]),
                "self_check": {b["w"]["sig"]: c for b, c in zip(bands, checks)}}
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            tt, smp = 0.0, 0.0
            for b in bands:
                w = b["w"]
                nh = int((w["cpu_epochs"] + 1.2) * w["spc"]) + w["spc"]
                tt += cpu_track_sample(b["x_dev"][:nh].cpu().numpy(), b["st"], b["chans"], w["cpu_epochs"], threads, w["mode"])
                smp += w["cpu_epochs"] * w["spc"]
            line["cpu_baseline"] = {"value": smp / tt / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{total} B1C ch x 2 epochs + {total} B2a ch x 20 epochs (20 ms of each band), float64 "
                                              "oracle restatement (C correlator), all host threads; MATLAB unavailable"}
        print(json.dumps(line))
    for b in bands:
        b["sess"].close()
    if dist is not None:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
# BASELINE config 5 as one joint run: acquisition -> preRun -> tracking of both bands.  Per step and band: the 63-PRN
# acquisition grid (PRN ranges sharded over the ranks, one all-reduce of the 3 x 63 result doubles), the reference's
# preRun on every rank (identical results), then the channels it produced are tracked (sharded round robin, sessions
# opened for them in the step) over the whole record, both bands concurrently.
# ----------------------------------------------------------------------------------------------
def run_pipeline(args):
    import numpy as np
    import torch
    import bds3_b200 as B
    from bds3_b200 import _acq, _lib as L, _shard, _track, synth
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L.init(local)
    n_sms = torch.cuda.get_device_properties(local).multi_processor_count
    n_samples = int(round(args.seconds * FS))
    total = min(args.channels, 60)
    bands = []
    for wl in ("track_b2a", "track"):
        w = TRACK[wl]
        st = settings_for(wl, total, args.seconds)
        st.acqSatelliteList = list(range(1, 64))
        sats = synth.make_sats(total, st, w["sig"], max_doppler=w["max_doppler"])
        x_dev = torch.empty(n_samples + 64, dtype=torch.int8, device="cuda")
        synth.synth_device(w["sig"], st, sats, n_samples, out_ptr=x_dev.data_ptr())
        spc = w["spc"]
        # what postProcessing reads for the acquisition: 20 code periods (B1C, postProcessing.m:94) / fineNoncoh + 2 ms (B2a)
        n_acq = 20 * spc if w["sig"] == "B1C" else spc * (int(st.fineNoncoh) + 2)
        n_epochs = max(1, int(math.floor((n_samples - 2 * spc) / (spc * (1 + 1e-5)))) - 1)
        bands.append(dict(w=w, st=st, sats=sats, x_dev=x_dev, n_acq=min(n_acq, n_samples), n_epochs=n_epochs,
                          sig=L.SIG_B1C if w["sig"] == "B1C" else L.SIG_B2A))
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def acquire(b):
        lo, hi = _shard.prn_range(63, rank, world)
        acq = _acq.acquire(b["sig"], None, b["st"], prn_range=(lo, hi), device_ptr=b["x_dev"].data_ptr(), n_samples=b["n_acq"])
        part = torch.tensor(np.stack([acq.carrFreq, acq.codePhase, acq.peakMetric]), device="cuda")
        if dist is not None:
            dist.all_reduce(part, op=dist.ReduceOp.SUM)      # disjoint PRN shards, zero = not found (acquisition.m:161-165)
        r = part.cpu().numpy()
        return B.Settings(carrFreq=r[0], codePhase=r[1], peakMetric=r[2])

    def step(keep=False):
        t = [time.perf_counter()]
        chans = []
        for b in bands:
            acq = acquire(b)
            chans.append(_acq.preRun(acq, b["st"], b1c=b["w"]["sig"] == "B1C"))       # preRun.m:44-76, every rank alike
        t.append(time.perf_counter())
        sess, b2a_ctas = [], 0
        for b, ch in zip(bands, chans):
            ch = [c for c in ch if c.PRN != 0]
            mine = _shard.shard_list(ch, rank, world)
            st_local = b["st"].copy()
            st_local.numberOfChannels = len(mine)
            st_local.numberOfChannels_total = total
            if b["w"]["sig"] == "B2a":
                cs = 8
                while cs > 1 and len(mine) * cs > n_sms // 2:
                    cs //= 2
                tuning = {"b2aClusterSize": cs}
                b2a_ctas = len(mine) * cs
            else:
                tuning = {"fwMaxCtas": max(16, n_sms - b2a_ctas)}
            s_ = _track.TrackSession(b["w"]["mode"], st_local, mine, device_ptr=b["x_dev"].data_ptr(), n_samples=n_samples,
                                     tuning=tuning) if mine else None
            sess.append((s_, mine, st_local))
        for (s_, _, _), b in zip(sess, bands):
            if s_ is not None:
                s_.run_async(b["n_epochs"])
        for s_, _, _ in sess:
            if s_ is not None:
                s_.sync()
        t.append(time.perf_counter())
        if keep:
            return t, chans, sess
        for s_, _, _ in sess:
            if s_ is not None:
                s_.close()
        return t, chans, None

    for _ in range(max(1, args.warmup)):
        step()
    launches0 = B.launch_count()
    barrier()
    t0 = time.perf_counter()
    acq_ms, trk_ms = [], []
    for _ in range(args.steps):
        t, _, _ = step()
        acq_ms.append((t[1] - t[0]) * 1e3)
        trk_ms.append((t[2] - t[1]) * 1e3)
    barrier()
    wall = time.perf_counter() - t0
    launches = B.launch_count() - launches0
    # ---- untimed: what the pipeline found and how the channels it started are doing at the end of the record
    _, chans, sess = step(keep=True)
    report = {}
    for b, ch, (s_, mine, st_local) in zip(bands, chans, sess):
        want = {st_.PRN: st_ for st_ in b["sats"]}
        got = [c for c in ch if c.PRN != 0]
        spc = b["w"]["spc"]
        ok_code = ok_freq = 0
        for c in got:
            if c.PRN in want:
                d = (c.codePhase - 1 - want[c.PRN].codeDelay) % spc
                ok_code += min(d, spc - d) <= 1.5
                ok_freq += abs(c.acquiredFreq - (b["st"].IF + want[c.PRN].doppler)) <= 25.0
        locked = 0
        if s_ is not None:
            planes = s_.fetch(b["n_epochs"])
            lock = util.lock_report(b["w"]["mode"], B.Settings(dict(st_local)), planes, b["n_epochs"], seconds_tail=10.0)
            locked = int(((lock["data_pld_min"] > 0.9) & (lock["pilot_pld_min"] > 0.9)).sum())
            assert int(planes["epochsDone"].min()) == b["n_epochs"]
            s_.close()
        agg = torch.tensor([float(locked)], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        report[b["w"]["sig"]] = {"satellites_in_record": len(want), "acquired": len(got), "false_alarms": sum(c.PRN not in want for c in got),
                                 "code_phase_within_1.5_samples": int(ok_code), "carrier_within_25_Hz": int(ok_freq),
                                 "channels_locked_over_last_10_s": int(agg[0])}
    tt = torch.tensor([wall * 1e3 / args.steps, sum(acq_ms) / len(acq_ms), sum(trk_ms) / len(trk_ms)], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        step_ms = float(tt[0])
        if_samples = sum(b["n_epochs"] * float(b["w"]["spc"]) for b in bands)
        print(json.dumps({
            "metric": f"IF Msamples/s through the joint {2 * total}-ch B1C + B2a acquisition -> tracking pipeline",
            "value": if_samples / (step_ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 (FFT, accumulate) / f64 loop closure (int8 IF)", "data": "synthetic",
            "config": {"workload": f"BASELINE config 5, joint: 63-PRN x +-5 kHz acquisition of both bands -> preRun -> tracking of the "
                                   f"{total} + {total} channels found, two int8 IF records at 99.375 MHz, {args.seconds:g} s",
                       "ms_acquisition_both_bands": float(tt[1]), "ms_sessions_and_tracking": float(tt[2]),
                       "parallelism": f"PRN ranges (acquisition) and channels (tracking) sharded over {world} GPU(s), records replicated",
                       "x_realtime": args.seconds * 1e3 / step_ms},
            "gpu_launches": int(launches), "pipeline": report}))
    if dist is not None:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
# secondary workload: BASELINE config 2 — B2a full 63-PRN x +-5 kHz acquisition grid (one GPU; PRNs shard across
# ranks through prn_lo/prn_hi when launched under torchrun)
# ----------------------------------------------------------------------------------------------
def run_acq_b2a(args, b1c=False):
    import numpy as np
    import torch
    import bds3_b200 as B
    from bds3_b200 import _acq, _lib as L, _shard, synth
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L.init(local)
    sig_name, SIG = ("B1C", L.SIG_B1C) if b1c else ("B2a", L.SIG_B2A)
    n_prn = int(os.environ.get("BDS_BENCH_ACQ_PRNS", 8 if b1c else 63))   # B1C: 201 bins x 2^22-point FFTs per PRN
    if b1c:
        st = B.b1c.initSettings(samplingFreq=FS, acqSatelliteList=list(range(1, n_prn + 1)))
        inj = [2, 5]
        n = 4 * int(round(FS * 0.01))                         # >= len10PlusXms + samplesPerCode (B1C/acquisition.m:135,239-241)
    else:
        st = B.b2a.initSettings(acqSatelliteList=list(range(1, n_prn + 1)))
        inj = [2, 9, 17, 23, 31, 40, 52, 61]
        n = 17 * 99375                                        # (fineNoncoh + 2) ms, B2a/postProcessing.m:89-90
    if os.environ.get("BDS_BENCH_ACQ_TUNE"):                   # developer A/B of the inverse passes (bdsgpu.h: cfg.tune)
        st["_tune"] = int(os.environ["BDS_BENCH_ACQ_TUNE"])
    sats = synth.make_sats(len(inj), st, sig_name, seed=3, prns=inj, cn0=47.0)
    x_dev = torch.empty(n + 64, dtype=torch.int8, device="cuda")
    synth.synth_device(sig_name, st, sats, n, out_ptr=x_dev.data_ptr())
    x_host = x_dev[:n].cpu().pin_memory().numpy()
    lo, hi = _shard.prn_range(n_prn, rank, world)
    nbins = int(round(st.acqSearchBand * 2 / st.acqStep)) + 1

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def timed(fn):
        for _ in range(max(args.warmup, 5)):   # the first calls pay one-off cudaMalloc growth of ~0.7-10 GB of work buffers
            fn()
        barrier()
        t0 = time.perf_counter()
        per = []
        for _ in range(args.steps):
            t1 = time.perf_counter()
            r = fn()
            per.append(round((time.perf_counter() - t1) * 1e3, 1))
        barrier()
        print("acq step ms:", per, file=sys.stderr)
        return sorted(per)[len(per) // 2] * 1e-3, r   # median step (secondary workload; allocation outliers happen)

    l0 = B.launch_count()
    t_dev, acq = timed(lambda: _acq.acquire(SIG, None, st, prn_range=(lo, hi), device_ptr=x_dev.data_ptr(), n_samples=n))
    launches = (B.launch_count() - l0) // (max(args.warmup, 5) + args.steps)   # calls made by timed()
    t_e2e, acq2 = timed(lambda: _acq.acquire(SIG, x_host, st, prn_range=(lo, hi)))
    tt = torch.tensor([t_dev, t_e2e], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    found = sorted(int(p) + 1 for p in np.nonzero(acq.carrFreq)[0])
    if rank == 0:
        cells = n_prn * nbins
        lgP = (22 if FS == 99.375e6 else max(8, int(math.ceil(math.log2(3 * round(FS * 0.01) - 1))))) if b1c else 19
        P = 1 << lgP                                          # 2^22 at 99.375 MHz, 2^21 at the shipped 53 MHz (N + M - 1 = 3 code periods - 1)
        # algorithmic bytes: per (PRN, bin) cell 2 inverse P-point complex-fp32 FFTs x 2 passes x (read+write)
        # x 8 B; + per bin one forward FFT (shared by all PRNs) and per PRN two code FFTs, 2 passes each
        alg = (cells * 2 + nbins + n_prn * 2) * 2 * 2 * 8 * P / world
        peak, peak_src = peaks()
        line = {"metric": f"{sig_name} acquisition grid cells/s ({n_prn} PRN x {nbins} Doppler bins, 2^{lgP}-point FFT, data+pilot)",
                "value": cells / float(tt[0]),
                "unit": "cells/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(tt[0]) * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (complex FFT), int8 IF",
                "data": "synthetic", "config": {"workload": (f"B1C {n_prn}-PRN x +-5 kHz acquisition (50 Hz bins, 10 ms coherent), int8 IF at {FS / 1e6:g} MHz" if b1c else
                                                               "BASELINE config 2: B2a 63-PRN x +-5 kHz acquisition, 17 ms int8 IF at 99.375 MHz"),
                                                  "prns_found": found, "injected": [s_.PRN for s_ in sats]},
                "roofline": {"bound": "hbm", "achieved": alg / float(tt[0]) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": alg / float(tt[0]) / 1e9 / peak, "traffic": acq_traffic(b1c, cells) if FS == 99.375e6 else None,
                             "peak_source": peak_src,
                             "kernel": "acq_inv_row_ct_kernel + acq_inv_col_ct_kernel (whole bds_acquire call: host code generation, allocation and the three phase synchronisations included)",
                             "algorithmic_bytes_per_launch": alg},
                "e2e": {"value": cells / float(tt[1]), "unit": "cells/s", "h2d_bytes_per_step": int(n), "d2h_bytes_per_step": 3 * n_prn * 8},
                "gpu_launches": int(launches)}
        if not args.no_cpu_baseline and not b1c:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import bds_oracle as O
            so = O.initSettings_B2a(acqSatelliteList=[2, 3])
            t0 = time.perf_counter()
            O.acquisition_B2a(x_host, so)
            tc = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": 2 * nbins / tc, "unit": "cells/s", "cores": 1, "kind": "port",
                                    "sample": "2 PRNs x 26 bins, numpy float64 oracle restatement (forward FFTs shared across PRNs)"}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    global FS
    args = parse()
    if args.fs != FS:
        if args.workload not in ("track", "track_nb", "acq_b1c"):
            raise SystemExit("--fs applies to --workload track, track_nb and acq_b1c")
        FS = float(args.fs)
        TRACK["track"]["spc"] = TRACK["track_nb"]["spc"] = int(round(FS * 0.01))     # samples per 10 ms B1C code period
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in ("acq_b2a", "acq_b1c"):
        run_acq_b2a(args, b1c=args.workload == "acq_b1c")
    elif args.workload == "dual":
        run_dual(args)
    elif args.workload == "pipeline":
        run_pipeline(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
