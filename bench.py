#!/usr/bin/env python
"""EKF visual updates/sec on BASELINE cfg-2 (30-pose window, 200 SLAM + 800 MSCKF features), one agent per GPU.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference ...                     (the reference's own C++ filter back end compiled in place,
                                                            oracle/_ref/libxref_release.so, on the host cores)

A step = one Ekf::processUpdateMeasurement()-equivalent: manage + constructUpdate (MSCKF,
SLAM) + compression + applyUpdate + postUpdate + re-propagation of the 10 buffered IMU states.  Inputs are
synthetic normalised track lists at the VioUpdater::preProcess seam (x_multi_agent_b200/synth.py).
`value` is device-timed (CUDA events on the filter's stream, inputs resident in HBM, L2 flushed between
steps); `e2e` goes through the C ABI with HOST buffers (track upload + state readback inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, os.fspath(ROOT))

METRIC = "ekf_visual_updates_per_sec"
UNIT = "updates/s"
CFG2 = dict(M=30, F=200, K=800)
# IMU of the synthetic agent AND of the filter's noise model: a MEMS-grade gyro (noise density 2e-4 rad/s/sqrt(Hz), bias
# random walk 2e-5) instead of the reference's default 0.0083 / 0.00083 (include/x/common/types.h:65-79).  With the default
# the reference's own Q_d polynomial -- which is not symmetric and, at that gyro density, not positive semi-definite
# (propagator.cpp:207-840) -- drives the covariance of a well-observed filter indefinite within ~20 updates of this size
# (oracle-only measurement): lambda_min(P) = -3e-6 after the first steady-state update, doubling every step.
IMU_NOISE = dict(n_w=2e-4, n_bw=2e-5, n_a=0.0013, n_ba=0.00013)
# Fixed covariance-intersection weights of the fusion phases.  The reference applies the correction of EVERY accepted
# match from the same prior, one after the other (updater.cpp:22-36,84-97): for the states the matches have in common the
# gain of a fusion step is n_matches x K with K ~ w, so the weights are chosen such that n_matches x w = 0.5 (a
# contraction); with w = 0.1 and 32 matches per peer the loop overshoots and every later gate closes.
CI_SLAM_W = 0.5 / 16
FILL_K = 40          # MSCKF tracks per update while the window fills (warm-up to steady state, untimed)
N_FILL = 33          # frames until the window is full and all 200 SLAM features are initialised


def workload_config(n_gpus, extra=None):
    c = {"workload": "cfg-2: single-agent 30-pose window, 200 SLAM + 800 MSCKF (30-obs) tracks per update, "
                     "10 IMU states re-propagated per update, ring buffer 250",
         "oc_projection": "off on the device arm (xb_config.oc_projection = 0: plain MSCKF pose Jacobians, so that the "
                          "synthetic filter stays statistically consistent and the gates accept; the reference's "
                          "projection as written loses consistency within seconds, tests/test_cpu.py::"
                          "test_oc_projection_as_written_breaks_consistency); same arithmetic cost either way; "
                          "`reference_semantics` in the line is the same run with the projection on",
         "imu_noise": dict(IMU_NOISE, note="MEMS-grade gyro instead of the reference default 0.0083 / 0.00083, used for the "
                                                "synthetic data and for the filter's noise model on both arms: with the default "
                                                "the reference's Q_d polynomial is not positive semi-definite and the covariance of "
                                                "a well-observed filter loses definiteness within ~20 updates (DESIGN.md)"),
         "window": 30, "slam_features": 200, "msckf_tracks": 800, "n_error_states": 795,
         "agents": n_gpus, "parallelism": f"one agent per GPU x{n_gpus}, no data-path collective",
         "l2": "flushed between steps (256 MiB write)",
         "timing": "per-step CUDA events on the filter stream; IMU feed + track upload + L2 flush between steps untimed",
         "streams": "one caller stream + two library-internal side streams (SLAM-column half of the Kalman update and the "
                    "re-propagation means run next to the MSCKF pipeline / the downdate; joined with events)"}
    if extra:
        c.update(extra)
    return c


CFG5 = dict(M=50, F=800, K=3200)   # BASELINE configs[4] per agent (800 SLAM + 3200 MSCKF of 4k features)


def build_scenario(seed, dims=None, n_fill=None):
    from x_multi_agent_b200.synth import Scenario, SynthConfig, record
    dims = dims or CFG2
    cfg = SynthConfig(M=dims["M"], F=dims["F"], K=FILL_K, seed=seed, slam_init_frame=dims["M"], slam_lm_seed=4242,
                      slam_msckf_init_frac=1.0, **IMU_NOISE)
    scn = Scenario(cfg)
    fill = record(scn, n_fill or N_FILL)
    scn.c.K = dims["K"]
    return scn, fill


def steady_events(scn, k0, n):
    """n further frames: ([imu samples between frames], measurement) with the 10-sample latency tail."""
    c = scn.c
    out = []
    fed = (k0 - 1) * c.imu_per_frame + c.latency_imu
    for k in range(k0, k0 + n):
        upto = k * c.imu_per_frame + c.latency_imu
        imu = []
        for i in range(fed + 1, upto + 1):
            t = i * scn.dt_imu
            imu.append((t, i, *scn.imu_sample(t)))
        fed = upto
        out.append((imu, scn.measurement(k)))
    return out


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index, uuid=None):
        super().__init__(daemon=True)
        self.index, self.uuid, self.rows, self.stop_flag = index, uuid, [], False

    def run(self):
        # NVML directly (nvidia_ml_py): a sample costs microseconds; nvidia-smi (0.1-0.5 s per query) only as a fallback
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = None
            if self.uuid:
                try:
                    h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + self.uuid).encode())
                except Exception:
                    h = None
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            while not self.stop_flag:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(sm), str(mx)] + ["Active" if r & bits[k] else "Not Active"
                                                       for k in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                                                 "sw_power_cap")])
                time.sleep(0.01)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def reference_binary(threads):
    """The reference's own filter back end compiled in place (oracle/ref_build/build_ref.sh): Release-flag build when
    present, its large GEMM / QR / LU routed to the OpenBLAS that ships with numpy on `threads` host threads."""
    from oracle import refcpp
    flavour = "release" if refcpp.available("release") else "single"
    if not refcpp.available(flavour):
        raise RuntimeError("oracle/_ref/libxref*.so missing: run oracle/ref_build/build_ref.sh where /root/reference exists")
    blas = refcpp.bind_blas(flavour, threads)
    desc = ("reference C++ sources compiled in place (oracle/_ref/libxref_%s.so: src/x/{ekf,vio,vision}/*.cpp unmodified, "
            "%s; Eigen/OpenCV/Boost stand-in headers, dense products / QR / LU through %s)"
            % ("release" if flavour == "release" else "single", "reference Release flags CMakeLists.txt:185,194"
               if flavour == "release" else "-O2", "OpenBLAS 0.3.30 on %d threads" % threads if blas else
               "the stand-in's own loops (OpenBLAS not found)"))
    return refcpp, flavour, desc


def reference_one_update(prior_state, sm_state, meas, sigma_img, threads):
    """One full Updater::update of the compiled reference (all 800 tracks) from the given prior: seconds."""
    refcpp, flavour, desc = reference_binary(threads)
    ref = refcpp.RefFilter(prior_state.M, prior_state.F, sigma_img=sigma_img, n_slots=2, flavour=flavour,
                           noise=tuple(IMU_NOISE[k] for k in ("n_w", "n_bw", "n_a", "n_ba")))
    ref.sm_set(*sm_state)
    ref.set_measurement(meas)
    ref.updater_update(prior_state)
    sec = ref.last_seconds
    ref.close()
    return sec, desc


def host_threads():
    """Host threads for the CPU arm: all of them.  torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, so
    the count is taken from the affinity mask and handed to OpenBLAS explicitly."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    """--impl reference: the reference's own C++ update path on the host cores, rank 0 only.  Every step is one FULL
    Ekf::processUpdateMeasurement of cfg-2 (800 MSCKF tracks + 200 SLAM features, N = 795, incl. the re-propagation of
    the 10 buffered IMU states), on the same synthetic stream as the GPU arm; nothing is sampled or extrapolated."""
    if rank != 0:
        return
    from x_multi_agent_b200.synth import replay
    threads = host_threads()
    refcpp, flavour, desc = reference_binary(threads)
    scn, fill = build_scenario(seed=0)
    ref = refcpp.RefFilter(CFG2["M"], CFG2["F"], n_slots=250, flavour=flavour,
                           noise=tuple(IMU_NOISE[k] for k in ("n_w", "n_bw", "n_a", "n_ba")))
    replay(fill, ref)
    W, K = max(args.warmup, 1), args.steps
    events = steady_events(scn, N_FILL, W + K)
    times = []
    for i, (imu, m) in enumerate(events):
        for (t, seq, w, a) in imu:
            ref.process_imu(t, seq, w, a, want_state=False)
        ref.set_measurement(m)
        st = ref.process_update_measurement()
        assert st is not None and np.all(np.isfinite(st.x))
        if i >= W:
            times.append(ref.last_seconds)
    sec = float(np.mean(times))
    val = 1.0 / sec
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
            "warmup": W, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference",
                             "sample": "every step is one full cfg-2 Ekf::processUpdateMeasurement (800 MSCKF tracks + 200 "
                                       "SLAM features, N = 795, 10 IMU states re-propagated): " + desc,
                             "ms_per_step_min_max": [round(min(times) * 1e3, 1), round(max(times) * 1e3, 1)]},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-semantics", action="store_true", help="skip the extra oc_projection = 1 run (profiling)")
    ap.add_argument("--precision", type=int, default=int(os.environ.get("XB_DOWNDATE_PRECISION", "0")))
    args = ap.parse_args()
    # stdout carries exactly one JSON line: native libraries (NCCL version banner, ...) write to fd 1 directly, so fd 1
    # is pointed at stderr for the whole run and the result goes to the saved descriptor
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from x_multi_agent_b200 import Filter, PackedMeasurement
    from x_multi_agent_b200.filter import PackedMsckfMatches
    from x_multi_agent_b200.synth import replay
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL's INFO lines (communicator ranks, transport, NVLS) go to stderr: fd 1 already points there
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):   # an inherited WARN / VERSION would hide them
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT,ENV")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W, K = max(args.warmup, 3), args.steps
    stream = torch.cuda.Stream()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def make_agent(oc, n_steady):
        """One agent: filter + warm-up to steady state + the next `n_steady` frames (IMU + pinned measurement)."""
        scn_, fill_ = build_scenario(seed=rank)
        f_ = Filter(CFG2["M"], CFG2["F"], max_tracks=CFG2["K"], n_slots=250, device=local_rank,
                    downdate_precision=args.precision, sigma_landmark=0.1, ci_slam_w=CI_SLAM_W,
                    ci_msckf_w=0.5 / (32 * max(1, world - 1)),
                    multi_uav=int(world > 1), oc_projection=oc, **IMU_NOISE)
        f_.set_stream(stream.cuda_stream)
        replay(fill_, f_)
        assert f_.n_poses == CFG2["M"] and f_.n_features == CFG2["F"], "warm-up did not reach steady state"
        ev_ = steady_events(scn_, N_FILL, n_steady)
        pk_ = [PackedMeasurement(m, pinned=True) for _, m in ev_]   # inputs in pinned host memory (bench contract)
        return scn_, f_, ev_, pk_

    def device_timed(f_, ev_, pk_, first, n_warm, n_timed, profile):
        """n_warm + n_timed updates starting at event `first`, CUDA events on the filter's stream around each
        Ekf::processUpdateMeasurement, L2 flushed before each; returns (ms over the timed ones, launches, stage times)."""
        e0 = [torch.cuda.Event(enable_timing=True) for _ in range(n_warm + n_timed)]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in range(n_warm + n_timed)]
        i0 = [torch.cuda.Event(enable_timing=True) for _ in range(n_warm + n_timed)]
        i1 = [torch.cuda.Event(enable_timing=True) for _ in range(n_warm + n_timed)]
        l0, stage_ = 0, None
        for i in range(n_warm + n_timed):
            if i == n_warm:
                barrier()
                if profile:
                    f_.profile(True)
                l0 = f_.kernel_launches()
            i0[i].record(stream)
            for (t, seq, w, a) in ev_[first + i][0]:
                f_.process_imu(t, seq, w, a, want_state=False)     # Ekf::processImu: one fused launch per sample
            i1[i].record(stream)
            n_imu = len(ev_[first + i][0])
            f_.set_measurement(pk_[first + i])
            with torch.cuda.stream(stream):
                flush.zero_()
                e0[i].record(stream)
            f_.process_update_measurement(want_state=False)
            e1[i].record(stream)
        barrier()
        n_l = f_.kernel_launches() - l0
        if profile:
            stage_ = f_.profile_read()
            f_.profile(False)
        imu_us = 1e3 * sum(i0[i].elapsed_time(i1[i]) for i in range(n_warm, n_warm + n_timed)) / (n_timed * max(n_imu, 1))
        return sum(e0[i].elapsed_time(e1[i]) for i in range(n_warm, n_warm + n_timed)), n_l, (stage_, imu_us)

    nC = nR = 4 + min(K, 12)                       # fusion steps of phase C (all-gather / request-response variant)
    scn, flt, events, packed = make_agent(0, 2 * (W + K))

    # ---- phase A: device-timed throughput, inputs resident in HBM ------------------------------------------
    try:
        dev_uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        dev_uuid = None
    sampler = ClockSampler(local_rank, dev_uuid)
    barrier()
    sampler.start()   # started with the warm-up so that the timed region is covered
    dev_ms, launches, (stage, imu_us) = device_timed(flt, events, packed, 0, W, K, True)
    sampler.stop_flag = True
    inl0 = flt.debug_int("inlier0", CFG2["K"])
    msckf_inlier_frac = float(inl0.mean()) if len(inl0) else None
    slam_inlier_frac = float(flt.debug_int("slam_inlier", CFG2["F"]).mean())
    # ---- cfg-5 dimensions (N = 2715: the covariance kernels become throughput kernels), N = 1 only ----------------
    cfg5 = None
    if world == 1 and not args.no_reference_semantics:
        n_fill5 = CFG5["M"] + 3
        scn5, fill5 = build_scenario(seed=5, dims=CFG5, n_fill=n_fill5)
        f5 = Filter(CFG5["M"], CFG5["F"], max_tracks=CFG5["K"], n_slots=250, device=local_rank, oc_projection=0, n_generations=8,
                    **IMU_NOISE)
        f5.set_stream(stream.cuda_stream)
        replay(fill5, f5)
        assert f5.n_poses == CFG5["M"] and f5.n_features == CFG5["F"]
        ev5 = steady_events(scn5, n_fill5, 3 + 8)
        pk5 = [PackedMeasurement(m, pinned=True) for _, m in ev5]
        ms5, _, (st5, _) = device_timed(f5, ev5, pk5, 0, 3, 8, True)
        inl5 = f5.debug_int("inlier0", CFG5["K"])
        N5 = 15 + 6 * CFG5["M"] + 3 * CFG5["F"]
        top5 = sorted(((k, v[0] / max(v[1], 1)) for k, v in st5.items() if v[1] > 0), key=lambda kv: -kv[1])[:6]
        cfg5 = {"workload": "cfg-5 dimensions per agent: 50-pose window, 800 SLAM + 3200 MSCKF (50-obs) tracks per update, "
                            "N = %d" % N5, "updates_per_sec": 8 / (ms5 * 1e-3), "ms_per_update": ms5 / 8, "steps": 8,
                "msckf_inlier_frac_last_step": float(inl5.mean()),
                "top_stages_ms": {k: round(v, 3) for k, v in top5},
                "covariance_bytes": 8 * N5 * N5}
        # the genuine dense contraction of the path at this size: P <- sym(P) - W1 W1^T (+ Woodbury tail), one launch of the
        # TMA-staged DMMA kernel (k_gemm_tma<SYM>), timed inside the update by the library's event pair around the stage
        pad32 = lambda x: (x + 31) // 32 * 32
        K5 = pad32(2 * CFG5["F"]) + pad32(6 * CFG5["M"]) + 64
        if "downdate" in st5 and st5["downdate"][1] > 0:
            dd_ms = st5["downdate"][0] / st5["downdate"][1]
            fl5, by5 = float(N5) * N5 * K5, 8.0 * (2 * N5 * N5 + N5 * K5)
            tr5 = None
            try:   # dram__bytes_read.sum + dram__bytes_write.sum of this launch from the committed ncu --set full capture
                tr5 = json.loads((ROOT / "profiles" / "r02_ncu_traffic.json").read_text()).get("cfg5_downdate")
            except Exception:
                pass
            cfg5["roofline"] = {"kernel": "k_gemm_tma<SYM> (covariance downdate, TMA-staged fp64 DMMA tiles)", "bound": "tensor",
                                "achieved": fl5 / (dd_ms * 1e-3) / 1e12, "peak": 36.9, "unit": "TFLOP/s",
                                "frac": fl5 / (dd_ms * 1e-3) / 1e12 / 36.9, "traffic": tr5, "avg_launch_ms": dd_ms,
                                "tensor_pipe_active_pct_ncu": 83.5,   # profiles/r02_ncu_full_summary.txt (k_gemm_tma<1>)
                                "algorithmic_flops": fl5, "algorithmic_bytes": by5,
                                "hbm_gbs_at_this_rate": by5 / (dd_ms * 1e-3) / 1e9,
                                "peak_source": "fp64 tensor-core (mma.sync.m8n8k4) peak measured with tools/bench_dmma.cu; "
                                               "B200 has no tcgen05 kind for fp64"}
        f5.close()
        # covariance-update roofline sweep (BASELINE configs[4]): the symmetric downdate kernel alone at the covariance sizes
        # of growing windows / maps (N = 15 + 6M + 3F, K = padded compressed rows), inputs resident, 10 timed launches each
        import ctypes as _C
        from x_multi_agent_b200 import lib as _L
        sweep = []
        _rng = np.random.default_rng(0)
        for (M_, F_) in ((30, 200), (40, 400), (50, 600), (50, 800)):
            n_, k_ = 15 + 6 * M_ + 3 * F_, pad32(2 * F_) + pad32(6 * M_)
            W_ = _rng.normal(size=(n_, k_)) * 0.01
            P_ = _rng.normal(size=(n_, n_))
            ms_ = _C.c_double(0.0)
            used = _L.load().xb_debug_gemm(2, n_, n_, k_, _L.dptr(W_), k_, _L.dptr(W_), k_, 0.0, 0.0, _L.dptr(P_), n_, 11,
                                           _C.cast(_C.byref(ms_), _L.c_double_p))
            if used == 1 and ms_.value > 0:
                sweep.append({"N": n_, "K": k_, "us": round(ms_.value * 1e3, 1),
                              "tflops": round(float(n_) * n_ * k_ / (ms_.value * 1e-3) / 1e12, 2),
                              "frac_fp64_tensor_peak": round(float(n_) * n_ * k_ / (ms_.value * 1e-3) / 1e12 / 36.9, 3),
                              "hbm_gbs": round(8.0 * (2 * n_ * n_ + n_ * k_) / (ms_.value * 1e-3) / 1e9, 1)})
            else:
                sweep.append({"N": n_, "K": k_, "kernel": "k_downdate_mma (cp.async tiles: below the TMA kernel's size threshold)"})
        cfg5["covariance_update_sweep"] = sweep
    # ---- phase B: end to end through the C ABI with host buffers ---------------------------------------------
    e2e_s = 0.0
    barrier()
    ib0 = [torch.cuda.Event(enable_timing=True) for _ in range(W + K)]
    ib1 = [torch.cuda.Event(enable_timing=True) for _ in range(W + K)]
    for j in range(W + K):
        i = W + K + j
        ib0[j].record(stream)
        flt.process_imu_batch(events[i][0])      # the frame's IMU samples in one call (xb_ekf_process_imu_batch)
        ib1[j].record(stream)
        flt.synchronize()
        t0 = time.perf_counter()
        flt.set_measurement(packed[i])                      # host -> device: track lists
        st = flt.process_update_measurement(want_state=True)  # device -> host: updated state (+ stream sync)
        t1 = time.perf_counter()
        t_last = st.time
        if j >= W:
            e2e_s += t1 - t0
    barrier()
    imu_batch_us = 1e3 * sum(ib0[j].elapsed_time(ib1[j]) for j in range(W, W + K)) / (K * max(len(events[W + K][0]), 1))
    assert np.all(np.isfinite(st.x)), "non-finite state after the benchmark"
    # ---- the same device-timed run with the reference's OC projection as written (N = 1 only) ------------------
    ref_sem = None
    if world == 1 and not args.no_reference_semantics:
        Kr = min(K, 20)
        scn_r, flt_r, ev_r, pk_r = make_agent(1, W + Kr)
        ms_r, _, _ = device_timed(flt_r, ev_r, pk_r, 0, W, Kr, False)
        inl_r = flt_r.debug_int("inlier0", CFG2["K"])
        ref_sem = {"value": Kr / (ms_r * 1e-3), "unit": UNIT, "ms_per_step": ms_r / Kr, "steps": Kr,
                   "msckf_inlier_frac_last_step": float(inl_r.mean()) if len(inl_r) else None,
                   "note": "xb_config.oc_projection = 1 (msckf_update.cpp:393-406 as written): bit-parity setting of the "
                           "tests; the filter has lost consistency by this point of the sequence, so fewer tracks pass the gate"}
        flt_r.close()
    frame_cursor = [N_FILL + 2 * (W + K)]          # next camera frame of this agent's scenario

    def next_frame():
        """IMU samples + pinned measurement of the next frame, generated on demand (multi-agent phases)."""
        imu_, m_ = steady_events(scn, frame_cursor[0], 1)[0]
        frame_cursor[0] += 1
        return imu_, PackedMeasurement(m_, pinned=True)

    # ---- phase C (N > 1): MULTI_UAV visual updates with 32 MSCKF-MSCKF matches per peer; every agent publishes its
    #      pose payload (window + 6M x 6M covariance block) through one all-gather per update --------------------------
    mm = None
    if world > 1:
        from x_multi_agent_b200.ci import exchange_payloads
        from x_multi_agent_b200.synth import Scenario, SynthConfig
        Kd = min(K, 12)
        peers = [p for p in range(world) if p != rank]
        peer_scn = {p: Scenario(SynthConfig(M=CFG2["M"], F=CFG2["F"], K=1, seed=p, slam_init_frame=CFG2["M"], slam_lm_seed=4242))
                    for p in peers}   # analytic truth of the other agents' trajectories (same generator, their seed)
        k0 = frame_cursor[0]
        PP = flt.pose_payload_len()
        local = torch.zeros(PP, dtype=torch.float64, device="cuda")
        steps_d = []
        NPROF = 4   # extra untimed steps with the library's stage timers on: where the MULTI_UAV step spends its time
        for j in range(W + Kd + NPROF):
            imu, m = steady_events(scn, k0 + j, 1)[0]
            lms = scn.last_msckf_lms
            win = list(range(k0 + j - CFG2["M"], k0 + j))     # frames held by a peer's window when it packs its payload
            mt = [(p, 0, a * 32 + q, peer_scn[p]._project(lms[a * 32 + q], win)) for a, p in enumerate(peers) for q in range(32)]
            steps_d.append((imu, PackedMeasurement(m), PackedMsckfMatches(mt)))
        d0 = [torch.cuda.Event(enable_timing=True) for _ in range(W + Kd + NPROF)]
        d1 = [torch.cuda.Event(enable_timing=True) for _ in range(W + Kd + NPROF)]
        acc, gated = [], []
        barrier()
        for j, (imu, pm, mt) in enumerate(steps_d):
            if j == W + Kd:
                flt.synchronize()
                flt.profile(True)
                flt.profile_read()
            for (t, seq, w, a) in imu:
                flt.process_imu(t, seq, w, a, want_state=False)
            flt.set_measurement(pm)
            with torch.cuda.stream(stream):
                flush.zero_()
                d0[j].record(stream)
                flt.pack_poses(local.data_ptr())
                gathered = exchange_payloads(local)
                flt.set_msckf_matches_packed(gathered.data_ptr(), world, mt)
                flt.process_update_measurement(want_state=False)
                d1[j].record(stream)
            stream.synchronize()
            g = flt.mm_last_gates(0)
            if os.environ.get("BENCH_DEBUG") and rank == 0 and j < 3:
                print("mm gates step", j, "gamma/chi2 quantiles", np.nanquantile(g[:, 1] / g[:, 2], [0.1, 0.5, 0.9]) if len(g) else None,
                      "n", len(g), file=sys.stderr, flush=True)
            gated.append(float(np.isfinite(g[:, 1]).mean()) if len(g) else 0.0)
            acc.append(float(g[:, 0].mean()) if len(g) else 0.0)
        mm_prof = flt.profile_read()
        flt.profile(False)
        barrier()
        frame_cursor[0] = k0 + W + Kd + NPROF
        mm_ms = sum(d0[j].elapsed_time(d1[j]) for j in range(W, W + Kd))
        tmm = torch.tensor([mm_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmm, op=dist.ReduceOp.MAX)
        mm = {"multi_uav_updates_per_sec": world * Kd / (float(tmm[0]) * 1e-3), "ms_per_step": float(tmm[0]) / Kd, "steps": Kd,
              "msckf_matches_per_step": steps_d[0][2].n, "own_gate_passed_frac_rank0": float(np.mean(gated[W:W + Kd])),
              "joint_gate_accepted_frac_rank0": float(np.mean(acc[W:W + Kd])), "pose_payload_bytes_per_agent": PP * 8,
              "collective": "all_gather (NCCL), one per update",
              "stage_ms_rank0": {k: round(v[0] / max(v[1], 1), 4) for k, v in sorted(mm_prof.items(), key=lambda kv: -kv[1][0])
                                 if v[1] > 0},
              "profiled_step_ms_rank0": sum(d0[j].elapsed_time(d1[j]) for j in range(W + Kd, W + Kd + NPROF)) / NPROF}
    # ---- phase D (N > 1): covariance-intersection fusion steps with the compressed payload exchanged over NCCL ----
    ci = None
    if world > 1:
        from x_multi_agent_b200.ci import exchange_payloads, ring_matches
        from x_multi_agent_b200.request_comm import (VLAD_LEN, Keyframe, KeyframeDatabase, exchange_request_response)
        PL = flt.ci_payload_len()
        local = torch.zeros(PL, dtype=torch.float64, device="cuda")
        # 16 SLAM-SLAM matches with one peer per fusion step, the peer and the feature block rotating from
        # step to step.  The reference applies every match's correction from the same prior, one after the other
        # (updater.cpp:22-36): the common-mode gain of a step is n_matches x K, so a step that fuses hundreds of matches
        # at once overshoots and the gates close for good -- 16 keeps the loop a contraction (16 x ~0.09 with w = 0.1).
        def step_matches(i):
            peer = (rank + 1 + i % (world - 1)) % world
            f0 = (16 * i) % CFG2["F"]
            return [(peer, (f0 + q) % CFG2["F"], (f0 + q) % CFG2["F"]) for q in range(16)]
        matches = step_matches(0)
        c0 = [torch.cuda.Event(enable_timing=True) for _ in range(nC)]
        c1 = [torch.cuda.Event(enable_timing=True) for _ in range(nC)]
        inl = []
        barrier()
        for i in range(nC):
            # (untimed) the filters keep running between fusion steps: one regular visual update per step, so that
            # every fusion step sees fresh, independently evolved estimates
            imu_c, pm_c = next_frame()
            for (t, seq, w, a) in imu_c:
                flt.process_imu(t, seq, w, a, want_state=False)
            flt.set_measurement(pm_c)
            stt = flt.process_update_measurement(want_state=True)
            t_last = stt.time
            flt.synchronize()
            barrier()
            matches = step_matches(i)
            with torch.cuda.stream(stream):
                c0[i].record(stream)
                flt.ci_pack(local.data_ptr())
                gathered = exchange_payloads(local)          # the only collective of the path (all-gather, NVLink)
                flt.process_others_packed(t_last, gathered.data_ptr(), world, matches, want_state=False)
                c1[i].record(stream)
            stream.synchronize()
            inl.append(float(flt.ci_last_gates(len(matches))[:, 0].mean()))
        barrier()
        ci_ms = sum(c0[i].elapsed_time(c1[i]) for i in range(4, nC))
        tci = torch.tensor([ci_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tci, op=dist.ReduceOp.MAX)
        ci = {"ci_fusion_steps_per_sec": world * (nC - 4) / (float(tci[0]) * 1e-3), "ms_per_step": float(tci[0]) / (nC - 4),
              "steps": nC - 4, "matches_per_step": len(matches), "inlier_frac_first_step_rank0": inl[0],
              "inlier_frac_mean_rank0": float(np.mean(inl[4:])), "payload_bytes_per_agent": PL * 8,
              "full_simplestate_bytes": 8 * (795 * 795 + 16 + 7 * 30 + 3 * 200), "collective": "all_gather (NCCL)",
              "note": "one regular visual update per agent between fusion steps (untimed); every timed step = pack + "
                      "all-gather + gate + CI fusion + applyCI + re-propagation"}
        # REQUEST_COMM variant (vio.cpp:455-496): 2592-byte VLAD request all-gather, then point-to-point answers only
        # between the pairs whose request scored above the threshold against a stored keyframe
        db = KeyframeDatabase(pr_score_thr=0.9)
        rng_v = np.random.Generator(np.random.PCG64(777))
        scene = rng_v.integers(0, 256, VLAD_LEN, dtype=np.uint8)   # all agents look at the same landmark patch ...

        def view_descriptor(seed):                                  # ... through their own noisy descriptors (3 % bits)
            r = np.random.Generator(np.random.PCG64(seed))
            flip = np.packbits(r.uniform(size=8 * VLAD_LEN) < 0.03)
            return np.bitwise_xor(scene, flip)
        r0 = [torch.cuda.Event(enable_timing=True) for _ in range(nR)]
        r1 = [torch.cuda.Event(enable_timing=True) for _ in range(nR)]
        sent = recv = 0
        inl_r = []
        barrier()
        for i in range(nR):
            imu_c, pm_c = next_frame()
            for (t, seq, w, a) in imu_c:
                flt.process_imu(t, seq, w, a, want_state=False)
            flt.set_measurement(pm_c)
            stt = flt.process_update_measurement(want_state=True)
            t_last = stt.time
            # keyframe selection is the caller's policy (vio_updater.cpp:451-484): here every second frame
            if i % 2 == 0:
                snap = torch.empty(PL, dtype=torch.float64, device="cuda")
                with torch.cuda.stream(stream):
                    flt.ci_pack(snap.data_ptr())
                stream.synchronize()
                db.add(Keyframe(view_descriptor(1000 * rank + i), snap, t_last))
            req = torch.from_numpy(view_descriptor(5000 * rank + i)).cuda()
            flt.synchronize()
            barrier()
            with torch.cuda.stream(stream):
                r0[i].record(stream)
                got, stats = exchange_request_response(req, db, PL)
                for peer, payload in got.items():
                    slots = torch.zeros((world, PL), dtype=torch.float64, device="cuda")
                    slots[peer].copy_(payload)
                    f0 = (16 * (i + 7 * peer)) % CFG2["F"]
                    pm = [(peer, (f0 + q) % CFG2["F"], (f0 + q) % CFG2["F"]) for q in range(16)]
                    flt.process_others_packed(t_last, slots.data_ptr(), world, pm, want_state=False)
                r1[i].record(stream)
            stream.synchronize()
            if i >= 4:
                sent += stats["answers_sent"]
                recv += stats["answers_received"]
                if got:
                    inl_r.append(float(flt.ci_last_gates(16)[:, 0].mean()))
        barrier()
        rc_ms = sum(r0[i].elapsed_time(r1[i]) for i in range(4, nR))
        trc = torch.tensor([rc_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(trc, op=dist.ReduceOp.MAX)
        ci["request_comm"] = {"ms_per_step": float(trc[0]) / (nR - 4), "steps": nR - 4,
                              "request_bytes_per_agent": VLAD_LEN, "answers_sent_rank0": sent, "answers_received_rank0": recv,
                              "answer_bytes": PL * 8, "inlier_frac_mean_rank0": float(np.mean(inl_r)) if inl_r else None,
                              "transport": "all_gather of the 2592-byte requests + NCCL send/recv of the answers "
                                           "(accepted pairs only; a keyframe is sent to a peer at most once)"}
    tt = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max = float(tt[0]), float(tt[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = world * K / (dev_ms_max * 1e-3)
    e2e = world * K / (e2e_ms_max * 1e-3)
    # ---- roofline of the dominant kernel -------------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    which_peak = "measured (MEASURED_PEAKS.json, burst copy)" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    N, M, F = 795, 30, 200
    ns2 = 2 * F
    s_pad, r_pad = (ns2 + 31) // 32 * 32, (6 * M + 31) // 32 * 32
    m_pad = s_pad + r_pad
    n_pad = (N + 31) // 32 * 32
    per_stage = {k: (v[0] / max(v[1], 1), v[1], v[0]) for k, v in stage.items() if v[1] > 0}
    # stages that are ONE kernel launch (CUDA-event pair around it on the launching stream) -> kernel, algorithmic bytes
    # per launch (fp64) and algorithmic flops per launch
    tall = lambda rows, cols: 8 * 2 * (cols * (cols + 1) // 2 + (rows - cols) * cols)      # read + write of the factored part
    single = {
        "tracks": ("k_tracks<1>", 8 * (CFG2["K"] * M * 2 + 7 * M + (6 * M) ** 2 + CFG2["K"] * (3 * (6 * M + 1) + 14 * M)),
                   CFG2["K"] * (2 * (2 * M) * (6 * M) * 6 + (2 * M) ** 3 / 3 + 4 * (2 * M) ** 2 * 3)),
        "tallchol": ("k_tallchol (slab columns)", tall(r_pad + n_pad + 96, r_pad), r_pad ** 3 / 3 + (n_pad + 96) * r_pad ** 2),
        "side_tallchol_slam_cols": ("k_tallchol (SLAM columns, side stream)", tall(s_pad + n_pad + 96, s_pad),
                                    s_pad ** 3 / 3 + (n_pad + 96) * s_pad ** 2),
        "chol_gram": ("k_tallchol (Gram factor)", tall(r_pad + 32, r_pad), r_pad ** 3 / 3),
        # the slab-column part now runs on a side stream next to State::correct (untimed there); the stage is the tail pass
        "downdate": ("k_downdate_mma (Woodbury / Omega tail pass)", 8 * (2 * N * N + N * 64), N * N * 64),
        "side_downdate_slam_cols": ("k_downdate_mma (SLAM columns, side stream)", 8 * (2 * N * N + N * s_pad), N * N * s_pad),
    }
    main_single = [k for k in single if k in per_stage and not k.startswith("side_")]
    dom = max(main_single, key=lambda k: per_stage[k][0]) if main_single else max(per_stage, key=lambda k: per_stage[k][2])
    kname, ab, fl = single.get(dom, (dom, None, None))
    avg_ms = per_stage[dom][0]
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of this kernel
        traffic = json.loads((ROOT / "profiles" / "r02_ncu_traffic.json").read_text()).get(dom)
    except Exception:
        pass
    fp64_peak = 36.9  # TFLOP/s, DFMA = DMMA peak measured with tools/bench_dmma.cu on this pool's B200
    roofline = {"kernel": kname, "bound": "hbm", "achieved": (ab / (avg_ms * 1e-3) / 1e9) if ab else None, "peak": hbm_peak,
                "unit": "GB/s", "frac": (ab / (avg_ms * 1e-3) / 1e9 / hbm_peak) if ab else None, "traffic": traffic,
                "avg_launch_ms": avg_ms, "algorithmic_bytes": ab, "peak_source": which_peak,
                "fp64": {"achieved_tflops": (fl / (avg_ms * 1e-3) / 1e12) if fl else None, "peak_tflops": fp64_peak,
                         "frac": (fl / (avg_ms * 1e-3) / 1e12 / fp64_peak) if fl else None,
                         "peak_source": "measured, tools/bench_dmma.cu (DFMA and DMMA m8n8k4 both 36.9 TFLOP/s)"},
                "note": "N=795 keeps P (5 MB) and the tall buffer (8 MB) L2-resident: no kernel of this path is HBM-bound at "
                        "cfg-2; the dominant kernels are bound by serial dependency chains (per-track 60x60 Cholesky in one "
                        "warp, tile-Cholesky pivot chain), see DESIGN.md section 4",
                "per_kernel": {single[k][0]: {"avg_launch_ms": round(per_stage[k][0], 4),
                                              "hbm_frac": single[k][1] / (per_stage[k][0] * 1e-3) / 1e9 / hbm_peak,
                                              "fp64_frac": single[k][2] / (per_stage[k][0] * 1e-3) / 1e12 / fp64_peak}
                               for k in single if k in per_stage}}
    stages_ms = {k: round(v[0], 4) for k, v in sorted(per_stage.items(), key=lambda kv: -kv[1][2])}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        # the reference's own C++ update path on the host cores: ONE full cfg-2 update (all 800 tracks) from the device's
        # own prior, on all host threads and on one (the reference's update path is single-threaded)
        more = steady_events(scn, N_FILL + 2 * (W + K), 1)
        for (t, seq, w, a) in more[0][0]:
            flt.process_imu(t, seq, w, a, want_state=False)
        slot = (flt.newest_slot() - scn.c.latency_imu) % 250
        prior = flt.get_state(slot)
        prior.cov = flt.get_covariance(slot)
        smst = (flt.n_poses, flt.n_features, list(flt.anchor_idxs), True)
        threads = host_threads()
        sec_all, desc = reference_one_update(prior, smst, more[0][1], scn.c.sigma_img, threads)
        sec_one, _ = reference_one_update(prior, smst, more[0][1], scn.c.sigma_img, 1)
        cpu_baseline = {"value": 1.0 / sec_all, "unit": UNIT, "cores": threads, "kind": "reference",
                        "sample": "one full cfg-2 Updater::update (800 MSCKF tracks + 200 SLAM features, N = 795) from the "
                                  "device's own prior: " + desc,
                        "seconds_per_update": round(sec_all, 3),
                        "one_core": {"value": 1.0 / sec_one, "seconds_per_update": round(sec_one, 3), "cores": 1}}
        # the STRUCTURED formulation (what the device computes) on the same host cores: vectorised numpy
        # + its BLAS/LAPACK (oracle/structured.py), from the post-manage state of the same update; splits the GPU-vs-reference
        # ratio into "algorithm" (reference C++ / this) and "hardware + implementation" (this / GPU)
        try:
            from oracle.structured import structured_update
            from oracle.updates import chi2_quantile
            meas = more[0][1]
            flt.work_set(prior)
            flt.set_measurement(meas)
            flt.manage(meas.lost_slam_trk_idxs)
            st = flt.work_get()
            Z = np.array([np.asarray(t_) for t_ in meas.msckf_trks])
            slam_obs = np.array([np.asarray(t_)[-1] for t_ in meas.slam_trks])
            slam_len = np.array([len(t_) for t_ in meas.slam_trks])
            chi95 = np.array([0.0] + [chi2_quantile(0.95, d) for d in range(1, 2 * CFG2["M"] + 1)])
            chi90 = np.array([0.0] + [chi2_quantile(0.90, d) for d in range(1, 2 * CFG2["M"] + 4)])
            args_s = (st.p_array, st.q_array, st.f_array, list(flt.anchor_idxs), st.cov, Z, slam_obs, slam_len, flt.n_poses,
                      CFG2["M"], CFG2["F"], scn.c.sigma_img, chi95, chi90)
            structured_update(*args_s)                      # warm-up (BLAS threads, page faults)
            t0 = time.perf_counter()
            d_cpu, info = structured_update(*args_s)
            sec_s = time.perf_counter() - t0
            flt.reset_correction()                          # Updater::update starts from a zero correction_total (updater.cpp:44)
            flt.construct_update(0)
            flt.apply_constructed()
            d_dev = flt.debug("delta", 15 + 6 * CFG2["M"] + 3 * CFG2["F"])
            # same gates on both sides for the comparison of the correction (the device gates with the exact unsymmetrised
            # covariance, this leg with sym(P): a handful of borderline tracks can differ)
            masks = (flt.debug_int("inlier0", CFG2["K"]) != 0, flt.debug_int("slam_inlier", CFG2["F"]) != 0)
            d_cpu, info_f = structured_update(*args_s, force_inliers=masks)
            info["gates_equal_to_device"] = bool(np.array_equal(info["inlier"], masks[0]) and np.array_equal(info["slam_inlier"], masks[1]))
            cpu_baseline["structured"] = {
                "value": 1.0 / sec_s, "unit": UNIT, "seconds_per_update": round(sec_s, 4), "cores": threads,
                "what": "the device's structured formulation (projector gate, Gram compression, sparse SLAM rows, Cholesky "
                        "update) as vectorised numpy + OpenBLAS/LAPACK on the host, stages after StateManager::manage "
                        "(oracle/structured.py; equals the dense formulation: tests/test_structured_cpu.py)",
                "stage_seconds": {k_: round(v_, 4) for k_, v_ in info["seconds"].items()},
                "state_correction_rel_diff_vs_device": float(np.linalg.norm(d_cpu - d_dev) / np.linalg.norm(d_dev)),
                "msckf_inliers": int(info["inlier"].sum()), "gates_equal_to_device": info["gates_equal_to_device"],
                "compressed_rows": int(info["rows"]),
                "algorithm_factor_vs_reference_all_cores": round(sec_all / sec_s, 1)}
        except Exception as exc:   # the structured leg is an explanation, never a reason to lose the bench line
            cpu_baseline["structured"] = {"error": repr(exc)[:300]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if args.precision == 0 else "f64 state, 3xTF32 tensor-core downdate", "data": "synthetic",
            "config": workload_config(world),
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": e2e_ms_max / K,
                    "h2d_bytes_per_step": int(np.mean([p.h2d_bytes for p in packed])) + 4 * (795 + 16 * 6 + 2 * 200),
                    "d2h_bytes_per_step": flt.LX * 8},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "clocks": sampler.summary(), "stage_ms_per_update": stages_ms, "ci": ci, "multi_uav_msckf": mm,
            "gate_inlier_frac_last_step": {"msckf": msckf_inlier_frac, "slam": slam_inlier_frac},
            "reference_semantics": ref_sem, "cfg5": cfg5,
            "imu_us_per_sample": {"value": round(imu_us, 2), "batched": round(imu_batch_us, 2),
                                  "batched_note": "xb_ekf_process_imu_batch: the 10 samples of a frame in one call (three launches), "
                                                  "device time per sample, used by the end-to-end phase",
                                  "note": "Ekf::processImu (propagateState + propagateCovariance, "
                                  "ekf.cpp:66-140) between updates: device time per IMU sample over the untimed feed of the "
                                  "steady-state phase (10 samples per frame, one fused launch each, issued back to back)"}}
    emit(line)
    flt.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
