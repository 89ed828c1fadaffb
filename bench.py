#!/usr/bin/env python
"""bench.py — headline benchmark: ICP scan-pairs/s on the KITTI-shaped synthetic odometry sequence
(BASELINE.json configs[1]: 1000 frame pairs, 64-beam ~120k-point scans, voxel 0.5 m, normals k=20, ICP 50 it / 1e-6,
Scan Context of every frame).

A step = one pass of the whole front end over the sequence: voxel grid -> Scan Context -> index build -> normals ->
batched point-to-plane ICP of the F consecutive pairs -> results on the host.
  value : pairs/s with the raw scans already resident in HBM (sb_register_batch_dev), CUDA-event timed.
  e2e   : the same through sb_register_batch with HOST (pinned) scans: H2D of every scan + D2H of results inside the
          timed region.
  N > 1 : one process per GPU, every rank registers its own sequence (weak scaling), results all-gathered over NCCL.
--impl reference times the CPU oracle (the reference's algorithm restated: it cannot be compiled here) on all host
cores over a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "lidar-slam-from-scratch_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

SENSOR = dict(beams=64, azimuth_steps=1875, elev_top_deg=2.0, elev_bot_deg=-24.8, max_range=120.0, noise_sigma=0.02,
              sensor_height=1.73)
VOXEL = 0.5
LOOP_LEN_M = 1200.0
RADIUS = LOOP_LEN_M / (2.0 * np.pi)


def make_world(synth):
    """Box city around a 1.2 km circular loop."""
    half = RADIUS + 100.0
    n_boxes = int(400 * (2 * half) ** 2 / 180.0 ** 2)
    return synth.scene(11, n_boxes=n_boxes, half_extent=half, path_kind=1, radius=RADIUS, corridor_half=6.0)


def make_poses(synth, n_scans, step_m=1.0):
    return np.stack([synth.pose(1, RADIUS, i * step_m) for i in range(n_scans)])


class ClockSampler:
    def __init__(self, device):
        self.device = device
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        # NVML in-process (a sample costs microseconds, so a 0.2 s timed region still gets ~10 of them); nvidia-smi
        # as a subprocess only if the binding is missing
        try:
            import pynvml
            pynvml.nvmlInit()
            try:  # the CUDA ordinal need not be the NVML index (CUDA_VISIBLE_DEVICES): go by PCI address
                import torch
                pr = torch.cuda.get_device_properties(self.device)
                h = pynvml.nvmlDeviceGetHandleByPciBusId(
                    ("%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)).encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {"hw_slowdown": pynvml.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": pynvml.nvmlClocksThrottleReasonSwPowerCap}
            while not self._stop.is_set():
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for n, b in bits.items():
                    if r & b:
                        self.reasons.add(n)
                self._stop.wait(0.02)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.check_output(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                               str(self.device)], timeout=5).decode().strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=10)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_pair(orc, raw_src, raw_tgt):
    """The reference's per-frame front end on the CPU oracle: voxel_downsample of the new frame + icp_point_to_plane
    (slam_node.cpp:122,132-138) + Scan Context (loop_closure.hpp:55), with the reference's doubled NN search."""
    ds, _ = orc.voxel_downsample(raw_src, VOXEL)
    dt, _ = orc.voxel_downsample(raw_tgt, VOXEL)  # (the reference keeps the previous frame's downsampled cloud; see below)
    orc.sc_compute(ds)
    return orc.icp_point_to_plane(ds, dt, faithful_cost=1)


def run_reference(args):
    """--impl reference: CPU oracle on all host cores over a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib
    from concurrent.futures import ThreadPoolExecutor
    orc, syn = oracle_lib.Oracle(), oracle_lib.Synth()
    cores = os.cpu_count() or 1
    world = make_world(syn)
    n_pairs = max(cores, min(2 * cores, 64))
    poses = make_poses(syn, n_pairs + 1)
    scans = [syn.scan(SENSOR, world, poses[i], 1000 + i, threads=cores) for i in range(n_pairs + 1)]
    ds = [None] * (n_pairs + 1)

    def one(i):  # ctypes releases the GIL: real parallelism over independent pairs
        d, _ = orc.voxel_downsample(scans[i + 1], VOXEL)
        orc.sc_compute(d)
        t, _ = orc.voxel_downsample(scans[i], VOXEL)
        return orc.icp_point_to_plane(d, t, faithful_cost=1)["num_iterations"]

    def step():
        with ThreadPoolExecutor(cores) as ex:
            return list(ex.map(one, range(n_pairs)))

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = n_pairs / dt
    sample = f"{n_pairs} consecutive pairs of the sequence per step, {cores} threads, one pair per task"
    print(json.dumps({
        "impl": "reference", "metric": "ICP scan-pairs/s", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n_pairs, note="bounded sample"),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(frames, note=None):
    c = {"workload": f"C2: {frames}-pair frame-to-frame odometry on a KITTI-shaped synthetic sequence "
                     "(64 beams x 1875 az, ~119k pts/scan, 1 m/frame on a 1.2 km loop)",
         "voxel_m": VOXEL, "normals_k": 20, "icp_max_iterations": 50, "icp_tolerance": 1e-6,
         "stages": "voxel_downsample + ScanContext + index + normals + point-to-plane ICP",
         "l2": "inputs (GBs of raw scans) are larger than the 126 MB L2"}
    if note:
        c["note"] = note
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=1000, help="pairs per step (1000 = the full C2 sequence)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import oracle_lib
    import slam_b200

    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    F = args.frames
    with torch.cuda.stream(stream):
        eng = slam_b200.Engine(local_rank, stream=stream.cuda_stream)
        syn = oracle_lib.Synth()
        world = make_world(syn)
        poses = make_poses(syn, F + 1)
        rays = SENSOR["beams"] * SENSOR["azimuth_steps"]
        d_raw = torch.empty((F + 1) * rays * 3, dtype=torch.float64, device="cuda")
        off = eng.synth_scans_dev(SENSOR, world, poses, 1000 + 7919 * rank, d_raw.data_ptr())  # input generation
        n_raw = int(off[-1])
        pair_src = np.arange(1, F + 1, dtype=np.int32)   # source = current frame (slam_node.cpp:132-138)
        pair_tgt = np.arange(0, F, dtype=np.int32)       # target = previous frame
        cfg = eng.icp_config()

        def step_dev():
            res, sc = eng.register_batch(None, off, pair_src, pair_tgt, voxel=VOXEL, cfg=cfg, want_sc=True,
                                         device_ptr=d_raw.data_ptr())
            return res, sc

        gathered = None
        # exchange buffers of the sharded path: one 160-byte record per pair and rank, gathered on every rank
        if world_size > 1:
            g_host = torch.empty((F, 20), dtype=torch.float64, pin_memory=True)
            g_dev = torch.empty((F, 20), dtype=torch.float64, device="cuda")
            g_all = torch.empty((world_size * F, 20), dtype=torch.float64, device="cuda")
        g_work = [None]

        def gather(res):
            """NCCL all-gather of the result records: the only exchange of the sharded path.  It is
            queued behind the step on the engine's stream and overlaps the next step; the closing barrier +
            synchronize waits for the last one, so every step's records have arrived inside the timed region."""
            if world_size == 1:
                return
            if g_work[0] is not None:
                g_work[0].wait()   # stream-level wait: the previous gather must have read g_dev before it is rewritten
            g_host.numpy()[...] = res.records20()
            g_dev.copy_(g_host, non_blocking=True)
            g_work[0] = dist.all_gather_into_tensor(g_all, g_dev, async_op=True)
            return g_all

        def barrier():
            if world_size > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # ---- warm-up
        for _ in range(max(args.warmup, 3)):
            res, sc = step_dev()
            gather(res)
        # ---- timed: device-resident inputs
        eng.set_profiling(True)
        sampler = ClockSampler(local_rank)
        barrier()
        if rank == 0:  # one sampler per node: nvidia-smi from every rank perturbs the run it is meant to watch
            sampler.start()
        l0 = eng.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stage_acc, host_acc = {}, {}
        e0.record(stream)
        for _ in range(args.steps):
            res, sc = step_dev()
            gathered = gather(res)
            for k, v in eng.stage_ms().items():
                stage_acc[k] = stage_acc.get(k, 0.0) + v
            for k, v in eng.stage_host_ms().items():
                host_acc[k] = host_acc.get(k, 0.0) + v
        e1.record(stream)
        barrier()
        clocks = sampler.stop()
        launches = eng.launch_count - l0
        if gathered is not None:  # the exchange really delivered this rank's records
            assert np.array_equal(gathered[rank * F:(rank + 1) * F].cpu().numpy(), res.records20())
        ms = e0.elapsed_time(e1) / args.steps
        counts = eng.last_counts()
        eng.set_profiling(False)
        t_ms = torch.tensor([ms], device="cuda", dtype=torch.float64)
        if world_size > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        ms_max = float(t_ms.item())
        value = world_size * F / (ms_max * 1e-3)
        # per-rank view: own elapsed time and device-busy time (sum of the stage events) -- tells a slow GPU
        # (busy grows) from host-side gaps or waiting in the gather (busy flat, elapsed grows)
        per_rank = None
        if world_size > 1:
            mine = torch.tensor([ms, sum(stage_acc.values()) / args.steps], device="cuda", dtype=torch.float64)
            allr = [torch.empty_like(mine) for _ in range(world_size)]
            dist.all_gather(allr, mine)
            per_rank = {"step_ms": [round(float(a[0]), 2) for a in allr],
                        "device_busy_ms": [round(float(a[1]), 2) for a in allr]}

        # ---- e2e: host (pinned) scans through the public entry points, H2D + D2H inside the timed region.
        # "e2e" takes the scans the way the reference's loader gets them from disk: float32 x, y, z records
        # (file_utils.cpp:91-97), widened on the device (sb_register_batch_f32).  "e2e_f64" takes the widened
        # PointCloud::Matrix rows (sb_register_batch), twice the bytes.
        e2e = None
        e2e_f64 = None
        if not args.no_e2e:
            d_view = d_raw[:n_raw * 3]
            h64 = torch.empty(n_raw * 3, dtype=torch.float64, pin_memory=True)
            h64.copy_(d_view)
            h32 = torch.empty(n_raw * 3, dtype=torch.float32, pin_memory=True)
            h32.copy_(d_view.to(torch.float32))
            torch.cuda.synchronize()
            assert torch.equal(h32.to(torch.float64), h64)  # the scans are float32-born: nothing is lost
            d2h = F * 1184 + (F + 1) * 9600

            def timed_host(h_np, bytes_per_row):
                def step_host():
                    return eng.register_batch(h_np, off, pair_src, pair_tgt, voxel=VOXEL, cfg=cfg, want_sc=True)
                for _ in range(2):
                    r2, s2 = step_host()
                    gather(r2)
                barrier()
                t0 = time.perf_counter()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record(stream)
                for _ in range(args.steps):
                    r2, s2 = step_host()
                    gather(r2)
                g1.record(stream)
                barrier()
                host_stages = eng.stage_host_ms()
                dev_stages = eng.stage_ms()
                wall = (time.perf_counter() - t0) / args.steps * 1e3
                ms2 = max(g0.elapsed_time(g1) / args.steps, wall)  # host-side work (staging, result decode) counts
                t2 = torch.tensor([ms2], device="cuda", dtype=torch.float64)
                if world_size > 1:
                    dist.all_reduce(t2, op=dist.ReduceOp.MAX)
                assert np.array_equal(res.transformations, r2.transformations)
                assert np.array_equal(sc, s2)
                return {"value": world_size * F / (float(t2.item()) * 1e-3), "unit": "pairs/s",
                        "h2d_bytes_per_step": int(n_raw * bytes_per_row), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": float(t2.item()),
                        "host_ms_last_step": {k: round(v, 2) for k, v in host_stages.items()},
                        "device_ms_last_step": {k: round(v, 2) for k, v in dev_stages.items()}}

            eng.set_profiling(True)
            e2e = timed_host(h32.numpy().reshape(-1, 3), 12)
            e2e["input"] = "pinned host float32 xyz records (as read from disk), widened on the device"
            e2e_f64 = timed_host(h64.numpy().reshape(-1, 3), 24)
            e2e_f64["input"] = "pinned host fp64 rows (PointCloud::Matrix)"
            eng.set_profiling(False)

        if rank != 0:
            if world_size > 1:
                dist.destroy_process_group()
            return

        # ---- roofline of the dominant stage (algorithmic bytes: SURVEY.md 8d / DESIGN.md)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        st = {k: v / args.steps for k, v in stage_acc.items()}
        N, M, T, Q = counts["raw_rows"], counts["voxel_rows"], counts["target_rows"], counts["nn_queries"]
        alg = {"voxel": 24 * N + 24 * M, "scan_context": 24 * M + 9600 * (F + 1), "index_build": 52 * T,
               "normals": 48 * T, "icp_loop": 72 * Q + 224 * F * max(counts["icp_iter_launches"], 1)}
        kern = {"voxel": "k_vox_insert (+ clear/list/sort/finalize/collect/patch)", "scan_context": "k_sc_compute",
                "index_build": "k_morton/k_sort_*/k_gather_leaves", "normals": "k_knn<1> (kNN + covariance + Jacobi)",
                "icp_loop": "k_icp_match/k_icp_fallback/k_icp_accum/k_icp_solve inside the WHILE graph"}
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        except Exception:
            pass
        dom = max(alg, key=lambda k: st.get(k, 0.0))
        dom_ms = st[dom]
        launches_dom = 4 * counts["icp_iter_launches"] + 1 if dom == "icp_loop" else 1
        achieved = alg[dom] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        roofline = {"bound": "hbm", "kernel": kern[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak,
                    # DRAM bytes of one launch of that kernel from the committed ncu capture (same command, 1000 frames)
                    "traffic": (traffic.get(dom, {}).get("dram_bytes_per_launch") if F == 1000 else None),
                    "traffic_source": "profiles/r01_traffic.json (ncu --set full)" if dom in traffic and F == 1000 else None,
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_step": int(alg[dom]), "launches_per_step": int(launches_dom),
                    "avg_launch_ms": dom_ms / max(launches_dom, 1),
                    "stages_ms": st, "stages_host_ms": {k: v / args.steps for k, v in host_acc.items()},
                    "stage_frac_of_peak": {k: (alg[k] / (st[k] * 1e-3) / 1e9 / peak if st.get(k, 0) > 0 else None)
                                           for k in alg}}

        # ---- cpu_baseline: the oracle, 1 thread (the reference is single-threaded), bounded sample
        orc = oracle_lib.Oracle()
        n_probe = 9
        h = d_raw[:int(off[n_probe]) * 3].cpu().numpy().reshape(-1, 3)
        scans = [h[int(off[i]):int(off[i + 1])] for i in range(n_probe)]
        t0 = time.perf_counter()
        cpu_res = cpu_pair(orc, scans[1], scans[0])
        t1 = time.perf_counter() - t0
        n_cpu = int(max(1, min(n_probe - 1, args.cpu_seconds / max(t1, 1e-3))))
        t0 = time.perf_counter()
        cpu_out = [cpu_pair(orc, scans[i + 1], scans[i]) for i in range(n_cpu)]
        cpu_dt = time.perf_counter() - t0
        # parity spot check on the same pairs
        max_dt = 0.0
        for i in range(n_cpu):
            dT = res[i].transformation @ np.linalg.inv(cpu_out[i]["transformation"])
            max_dt = max(max_dt, float(np.linalg.norm(dT[:3, 3])))
        cpu_baseline = {"value": n_cpu / cpu_dt, "unit": "pairs/s", "cores": 1, "kind": "port",
                        "sample": f"first {n_cpu} pairs of the same sequence, single thread, reference cost model "
                                  "(NN search twice per iteration); the reference's own sources build only over an "
                                  "Eigen stand-in (oracle/_ref), a checker whose eager loops are not Eigen's speed",
                        "parity_max_translation_diff_m": max_dt}
        try:  # for the record: the reference's own sources (oracle/_ref, Eigen stand-in) on the first two pairs
            import ref_lib
            if os.path.exists(ref_lib.SO):
                rf = ref_lib.Reference()
                t0 = time.perf_counter()
                for i in range(2):
                    a_ds = rf.voxel_downsample(scans[i + 1], VOXEL)
                    b_ds = rf.voxel_downsample(scans[i], VOXEL)
                    rf.sc_compute(a_ds)
                    rr = rf.icp_point_to_plane(a_ds, b_ds)
                cpu_baseline["reference_sources_over_eigen_standin"] = {
                    "value": 2 / (time.perf_counter() - t0), "unit": "pairs/s", "cores": 1,
                    "iterations_last_pair": rr["num_iterations"], "gpu_iterations_same_pair": int(res[1].num_iterations)}
        except Exception as ex:
            cpu_baseline["reference_sources_over_eigen_standin"] = {"error": repr(ex)}

        # ---- the other metrics BASELINE.json names, on their own configurations (short, device-resident, untimed
        # by the driver): C3 = k-NN (k=10) + normals on 128-beam scans at voxel 0.2; C4 = Scan Context search over a
        # 4000-keyframe database
        extras = {}
        try:
            s128 = dict(SENSOR, beams=128, azimuth_steps=2048)
            n3 = 48
            d3 = torch.empty(n3 * 128 * 2048 * 3, dtype=torch.float64, device="cuda")
            off3 = eng.synth_scans_dev(s128, world, poses[:n3], 5000, d3.data_ptr())
            cfg3 = eng.icp_config(max_iterations=0, normals_k=10)   # index + normals only; no ICP iterations
            p3s, p3t = np.arange(1, n3, dtype=np.int32), np.arange(0, n3 - 1, dtype=np.int32)
            for _ in range(3):
                eng.register_batch(None, off3, p3s, p3t, voxel=0.2, cfg=cfg3, device_ptr=d3.data_ptr())
            eng.set_profiling(True)
            eng.register_batch(None, off3, p3s, p3t, voxel=0.2, cfg=cfg3, device_ptr=d3.data_ptr())
            st3, c3 = eng.stage_ms(), eng.last_counts()
            eng.set_profiling(False)
            extras["c3_knn_normals"] = {
                "workload": "128 beams x 2048 az, voxel 0.2 m, k = 10, %d indexed clouds of %.0f points" %
                            (n3 - 1, c3["target_rows"] / (n3 - 1)),
                "knn_plus_normals_queries_per_s": c3["target_rows"] / (st3["normals"] * 1e-3),
                "ms": st3["normals"], "index_build_ms": st3["index_build"], "voxel_ms": st3["voxel"],
                "raw_points_per_s_voxel_grid": c3["raw_rows"] / (st3["voxel"] * 1e-3)}
            del d3
            rng = np.random.default_rng(3)
            det = slam_b200.LoopClosureDetector(eng, frame_gap=50, sc_distance_threshold=0.25)
            tiny = np.zeros((4, 3))
            base = np.where(rng.uniform(size=(64, 1200)) < 0.35, rng.uniform(-1.7, 9.0, (64, 1200)), 0.0)
            for i in range(4001):
                det.addFrame(tiny, i, desc=base[i % 64] + 0.01 * (i // 64))
            for _ in range(3):
                det.candidates_local()
            t0 = time.perf_counter()
            reps = 20
            for _ in range(reps):
                cd, ce = det.candidates_local()
            dt4 = (time.perf_counter() - t0) / reps
            extras["c4_scan_context_search"] = {
                "workload": "1 query vs 4000 descriptors (20x60, 60 column shifts), host call incl. result copy",
                "ms_per_query": dt4 * 1e3, "descriptor_pairs_per_s": 4000 / dt4,
                "db_bytes": 4000 * 9600, "gflops_fp64": 2 * 60 * 1200 * 4000 / dt4 / 1e9}
            det.close()
            # C2 as the reference runs it: one frame per call (SURVEY.md 8d "ms/frame odometry")
            sys.path.insert(0, os.path.join(ROOT, "scripts"))
            import stream_latency
            n_s = min(F, 150) + 1
            h_s = d_raw[:int(off[n_s]) * 3].cpu().numpy().reshape(-1, 3)
            scans_s = [np.ascontiguousarray(h_s[off[i]:off[i + 1]]) for i in range(n_s)]
            extras["c2_streaming"] = stream_latency.process_frames(eng, slam_b200, scans_s, VOXEL)
        except Exception as ex:  # the headline line must still be printed
            extras["error"] = repr(ex)

        iters = res.num_iterations
        out = {
            "metric": "ICP scan-pairs/s", "value": value, "unit": "pairs/s", "n_gpus": world_size, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(F),
            "ms_per_frame": ms_max / F, "clocks": clocks, "e2e": e2e, "e2e_f64": e2e_f64, "gpu_launches": int(launches), "per_rank": per_rank,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "extras": extras,
            "workload_stats": {"raw_points_per_scan": n_raw / (F + 1), "voxel_points_per_scan": M / (F + 1),
                               "icp_iterations_mean": float(iters.mean()), "icp_iterations_max": int(iters.max()),
                               "converged_frac": float(res.converged.mean())},
        }
        print(json.dumps(out))
        if world_size > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
