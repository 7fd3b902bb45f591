#!/usr/bin/env python3
"""bench.py -- ORB frames/s on synthetic KITTI-shaped frames (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config kitti|1080p|4k]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

A "step" is one pass of the whole hot path (pyramid -> FAST+NMS+Harris -> top-N -> orientation -> BRIEF) over one
batch of distinct synthetic frames per GPU.  --config selects BASELINE.json's workloads:
  kitti (default, the headline): 1000 frames of 1241x376, 8 levels, f = 1.2, N = 2000, thr 20, patch 31   (configs[2])
  1080p                        : 256 frames of 1920x1080, 8 levels, N = 5000                              (configs[3])
  4k                           : 64 frames of 3840x2160, 12 levels, N = 10000 (--frames sweeps the batch)  (configs[4])

  value : whole-job frames/s, frames and outputs resident in HBM, CUDA events on the launching stream, max over
          ranks.  The input batch is larger than L2 (the cache-hygiene rule used here).  ms_per_step is total / K; the
          per-step median and minimum are reported beside it.
  e2e   : same metric through the host-buffer C-ABI call (pinned host frames in, host records out), H2D and D2H
          inside the timed region.  e2e_link_bound is the copy-only figure: the same pinned buffers crossing the link
          in both directions with no kernels, i.e. what e2e cannot exceed on this host.
  roofline : the kernel with the largest share of the step; algorithmic bytes per frame = sum of pyramid pixels for
          k_pyramid (level 0 read + levels written) and k_fast (levels read), 44 B/record for k_describe; HBM peak from
          MEASURED_PEAKS.json.  config.pass_* report the whole-pass figure with B_min = sum(P) + 44 K.
          roofline_issue is the second roof (instruction issue) from the ncu capture recorded in profiles/latest_ncu.json.
  parity : the first frames of the bench batch through the CUDA path vs the CPU oracle (rank 0, N = 1, with the
          cpu_baseline leg): keypoint-set equality, descriptor identity, angle and Harris-response differences.
  cpu_baseline : the CPU oracle (port of the reference's orb_cpu path, multi-level composition) on this box's host
          cores, bounded sample, rank 0 only: all cores (frame-parallel) and one thread, median / min over passes.
Frames are independent, so ranks shard by frame and there is no data-path collective (weak scaling: fixed frames
per GPU).  --impl reference times the CPU oracle alone with all host threads.
"""
import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {   # BASELINE.json configs[2..4]
    "kitti": dict(W=1241, H=376, levels=8, nfeatures=2000, frames=1000),
    "1080p": dict(W=1920, H=1080, levels=8, nfeatures=5000, frames=256),
    "4k": dict(W=3840, H=2160, levels=12, nfeatures=10000, frames=64),
}
THR, PATCH, SCALE = 20, 31, 1.2
W = H = LEVELS = NFEAT = PITCH = SUM_P = P0 = 0


def level_sizes():
    out = [(W, H)]
    for l in range(1, LEVELS):
        s = np.float32(np.float64(np.float32(SCALE)) ** l)
        out.append((int(np.floor(np.float32(W) / s + np.float32(0.5))), int(np.floor(np.float32(H) / s + np.float32(0.5)))))
    return out


def workload_name(frames):
    return "kitti_synth_%dx%d_L%d_N%d_thr%d_patch%d_batch%d_per_gpu" % (W, H, LEVELS, NFEAT, THR, PATCH, frames)


def peaks():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(pk["hbm_gbs"]), "MEASURED_PEAKS.json (measured)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed regions (pynvml, 5 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop_evt, self.active = threading.Event(), False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._stop_evt.is_set():
            if self.active:
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
                except Exception:
                    pass
            time.sleep(0.005)

    def stop(self):
        self._stop_evt.set()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs NVML reports as local to its GPU (no-op when unavailable)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64 + 1)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


def all_host_cores():
    """Give the calling process every host CPU again (the GPU arm pins itself next to its GPU) and return their number."""
    try:
        os.sched_setaffinity(0, range(os.cpu_count() or 1))
    except Exception:
        pass
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def oracle_params(O):
    return O.params(nfeatures=NFEAT, scale_factor=SCALE, nlevels=LEVELS, fast_threshold=THR, orient_patch=PATCH,
                    select_policy=1, blur_levels=1, harris_k=0.04)


def cpu_passes(frames_np, threads, passes, warmup):
    """CPU oracle, frame-parallel over `threads` std::threads; returns the seconds of each timed pass."""
    from oracle import pyoracle as O
    p = oracle_params(O)
    for _ in range(warmup):
        O.detect_and_compute_batch(frames_np, p, NFEAT, threads)
    out = []
    for _ in range(passes):
        t0 = time.perf_counter()
        O.detect_and_compute_batch(frames_np, p, NFEAT, threads)
        out.append(time.perf_counter() - t0)
    return out


def cpu_sample_frames(V, cores):
    """Bounded sample of the workload for the CPU legs: ~a few seconds per pass on all cores."""
    per_frame_cost = (W * H) / float(1241 * 376)                      # relative to a KITTI frame (~60 ms on one core)
    n = int(max(cores, min(8 * cores, 256) / max(1.0, per_frame_cost)))
    return np.ascontiguousarray(V.synth_frames(n, W, H))


def cpu_baseline_leg(V):
    """All-core and one-thread CPU figures on a bounded sample of the workload (about 20 s of CPU work)."""
    cores = all_host_cores()
    frames = cpu_sample_frames(V, cores)
    n = frames.shape[0]
    dt1 = cpu_passes(frames, cores, 1, 1)[0]                          # calibration pass after one warm-up pass
    passes = int(min(40, max(3, round(10.0 / max(dt1, 1e-3)))))
    dts = cpu_passes(frames, cores, passes, 0)
    n1 = max(1, min(n, int(round(4.0 / max(dt1 * cores / n, 1e-3)))))  # ~4 s of single-thread work per pass
    dts1 = cpu_passes(frames[:n1], 1, 2, 0)
    return {"value": n / float(np.median(dts)), "unit": "frames/s", "cores": cores, "kind": "port",
            "value_best": n / min(dts), "value_mean": n * len(dts) / sum(dts),
            "one_thread": {"value": n1 / float(np.median(dts1)), "value_best": n1 / min(dts1), "frames": n1, "passes": len(dts1)},
            "sample": "%d passes over %d synthetic frames of the workload, frame-parallel over %d threads (value = median pass), %.1f s; "
                      "affinity reset to all host CPUs first (the GPU arm pins itself to its GPU's NUMA CPUs)" % (passes, n, cores, sum(dts))}


def parity_leg(V, frames_host, dev_index, n_check=8):
    """The first n_check frames of the bench batch: CUDA path (a small context with side arrays on) vs the CPU oracle.
    Follows the reference's own definition of the check (src/compare.cpp:82-107): keypoints matched by position,
    descriptors compared bit for bit."""
    from oracle import pyoracle as O
    n_check = min(n_check, frames_host.shape[0])
    fr = np.ascontiguousarray(frames_host[:n_check, :, :W])
    ctx = V.Context(V.make_params(nfeatures=NFEAT, scaleFactor=SCALE, nlevels=LEVELS, threshold=THR, patch_size=PATCH, device=dev_index,
                                  max_width=W, max_height=H, max_batch=1, max_keypoints=NFEAT, keep_side_arrays=1))
    p = oracle_params(O)
    tot = ident = set_equal = resp_equal = 0
    max_ulp = 0
    ham = []
    for f in range(n_check):
        k, a, d, npl = ctx.detect_and_compute(fr[f], NFEAT)
        xy, lid, rsp = ctx.get_side_arrays(0, len(k))
        r = O.detect_and_compute(fr[f], p, cap=NFEAT)
        got = {(int(l), int(x), int(y)): i for i, (l, x, y) in enumerate(zip(lid, xy["x"], xy["y"]))}
        ref = {(int(l), int(x), int(y)): i for i, (l, x, y) in enumerate(zip(r["level_id"], r["level_xy"]["x"], r["level_xy"]["y"]))}
        set_equal += int(set(got) == set(ref))
        common = [(got[key], ref[key]) for key in ref if key in got]
        tot += len(ref)
        gi = np.array([c[0] for c in common], np.int64)
        ri = np.array([c[1] for c in common], np.int64)
        if len(common):
            same = (d[gi] == r["desc"][ri]).all(axis=1)
            ident += int(same.sum())
            if (~same).any():
                ham.append(float(np.unpackbits(d[gi][~same] ^ r["desc"][ri][~same], axis=1).sum(axis=1).mean()))
            ua, ub = a[gi].view(np.int32).astype(np.int64), r["angles"][ri].view(np.int32).astype(np.int64)
            max_ulp = max(max_ulp, int(np.abs(ua - ub).max()))
            resp_equal += int(np.array_equal(rsp[gi].view(np.uint32), r["response"][ri].view(np.uint32)))
    ctx.close()
    return {"frames_checked": n_check, "keypoints_checked": tot, "keypoint_sets_equal": set_equal == n_check,
            "descriptor_identity": ident / max(1, tot), "mean_hamming_of_mismatches": (float(np.mean(ham)) if ham else 0.0),
            "max_angle_diff_ulp": max_ulp, "harris_bit_equal": resp_equal == n_check,
            "tolerance": "integer path bit-exact; angles and Harris responses 0 ulp (libm twins restate glibc); target >= 0.90 identity",
            "against": "CPU oracle (oracle/orb_oracle.cpp), the same frames"}


def ingest_leg(V, ctx, pool_frames, idx, cap, outs, n_kp_expected):
    """SURVEY 8(f)-3 (not part of the headline metric): the same batch read from 8-bit gray PNG files (page cache) through
    orb_detect_and_compute_files, host threads decoding into the pinned staging area while the GPU works.  Wall clock,
    because host threads are part of the path.  Beside it: single-thread decode rates of this library and of cv2."""
    import shutil
    import tempfile
    d = tempfile.mkdtemp(prefix="orb_ingest_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        files = []
        for i in range(len(pool_frames)):
            files.append(os.path.join(d, "%06d.png" % i))
            V.synth.write_png_gray8(files[-1], pool_frames[i], huffman_only=True)     # like the KITTI files: no LZ77 matches
        paths = [files[i] for i in idx]
        png_bytes = float(np.mean([os.path.getsize(f) for f in files]))
        out = tuple(t.numpy() for t in outs)
        threads = os.cpu_count() or 1
        ctx.detect_and_compute_files(paths, cap=cap, threads=threads, out=out)       # warm-up (pinned area, page cache)
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            ctx.detect_and_compute_files(paths, cap=cap, threads=threads, out=out)
        dt = (time.perf_counter() - t0) / reps
        assert int(out[3].sum()) == n_kp_expected, "file path and batch path disagree"
        res = {"host_decode_frames_per_s": len(paths) / dt, "frames": len(paths), "host_threads": threads,
               "png_bytes_per_frame": png_bytes, "png_encoding": "8-bit gray, Sub filter, Huffman-only deflate (as KITTI odometry files)"}
        # the same files with inflate + unfilter on the device (the host only reads, checks CRCs and uploads compressed bytes)
        ctx.detect_and_compute_files(paths, cap=cap, threads=threads, decode_on_device=True, out=out)
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.detect_and_compute_files(paths, cap=cap, threads=threads, decode_on_device=True, out=out)
        dt = (time.perf_counter() - t0) / reps
        assert int(out[3].sum()) == n_kp_expected, "device-decode path and batch path disagree"
        res["device_decode_frames_per_s"] = len(paths) / dt
        t0 = time.perf_counter()
        for f in files[:16]:
            V.imread_gray8(f)
        res["imread_1thread_frames_per_s"] = 16 / (time.perf_counter() - t0)
        try:
            import cv2
            t0 = time.perf_counter()
            for f in files[:16]:
                cv2.imread(f, cv2.IMREAD_GRAYSCALE)
            res["cv2_imread_1thread_frames_per_s"] = 16 / (time.perf_counter() - t0)
        except ImportError:
            pass
        return res
    finally:
        shutil.rmtree(d, ignore_errors=True)


def run_reference(args, rank):
    """--impl reference: the reference's CPU algorithm for this path on the host cores (oracle port; the
    reference's own orb_cpu.cpp is single-level and needs OpenCV)."""
    if rank != 0:
        return
    V = importlib.import_module("visual-odometry-gpu_b200")
    cores = all_host_cores()
    frames = cpu_sample_frames(V, cores)
    sample = frames.shape[0]
    dts = cpu_passes(frames, cores, args.steps, max(args.warmup, 1))
    dt = sum(dts) / len(dts)
    fps = sample / dt
    line = {"impl": "reference", "metric": "orb_frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "ms_per_step_median": float(np.median(dts)) * 1e3,
            "ms_per_step_min": min(dts) * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_name(args.frames), "frames_per_gpu_per_step": args.frames, "levels": LEVELS, "nfeatures": NFEAT,
                       "step_sample_frames": sample},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": "%d synthetic frames of the workload per step, frame-parallel over %d threads" % (sample, cores)},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=JSON_OUT, flush=True)


JSON_OUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="kitti", choices=sorted(CONFIGS), help="BASELINE.json workload (kitti = the headline)")
    ap.add_argument("--frames", type=int, default=None, help="frames per GPU per step (default: the config's batch)")
    ap.add_argument("--chunk", type=int, default=0, help="frames per kernel wave (0 = library default)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity leg")
    ap.add_argument("--no-ingest", action="store_true", help="skip the PNG-file ingest leg (SURVEY 8(f)-3)")
    ap.add_argument("--hot-only", action="store_true", help="only the ORB step: no cpu_baseline, matcher, tracker or ingest legs (ncu runs)")
    ap.add_argument("--shape", default=None, help="WxH of the synthetic frames (overrides the config)")
    ap.add_argument("--levels", type=int, default=None)
    ap.add_argument("--nfeatures", type=int, default=None)
    args = ap.parse_args()
    if args.hot_only:
        args.no_cpu = args.no_ingest = True
    global W, H, LEVELS, NFEAT, PITCH, SUM_P, P0
    cfg = CONFIGS[args.config]
    W, H, LEVELS, NFEAT = cfg["W"], cfg["H"], cfg["levels"], cfg["nfeatures"]
    if args.shape:
        W, H = (int(v) for v in args.shape.lower().split("x"))
    LEVELS = args.levels or LEVELS
    NFEAT = args.nfeatures or NFEAT
    args.frames = args.frames or cfg["frames"]
    PITCH = (W + 1 + 15) // 16 * 16                                # 16-byte rows, one spare byte after the last pixel
    SUM_P = sum(w * h for w, h in level_sizes())
    P0 = W * H
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the ORB path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    bind_to_gpu_numa_node(local_rank)            # pinned staging buffers get first-touched next to this GPU's PCIe root
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")       # stdout carries the JSON line only (NCCL prints a version banner)
        dist.init_process_group("nccl", device_id=dev)
    V = importlib.import_module("visual-odometry-gpu_b200")

    F, cap = args.frames, NFEAT
    lo, hi = V.shard_range(F * world, world, rank)          # global frame ids of this rank (weak scaling)
    # F distinct synthetic frames per rank (frame id = global frame index; the generator is not in the timed region)
    h_frames = torch.from_numpy(V.synth_frames(F, W, H, start=lo, pitch=PITCH)).pin_memory()      # (F, H, PITCH) pinned
    d_frames = h_frames.to(dev, non_blocking=True)
    d_k = torch.zeros(F, cap, 2, dtype=torch.int32, device=dev)
    d_a = torch.zeros(F, cap, dtype=torch.float32, device=dev)
    d_d = torch.zeros(F, cap, 32, dtype=torch.uint8, device=dev)
    d_n = torch.zeros(F, dtype=torch.int32, device=dev)
    h_k = torch.zeros(F, cap, 2, dtype=torch.int32).pin_memory()
    h_a = torch.zeros(F, cap, dtype=torch.float32).pin_memory()
    h_d = torch.zeros(F, cap, 32, dtype=torch.uint8).pin_memory()
    h_n = torch.zeros(F, dtype=torch.int32).pin_memory()

    ctx = V.Context(V.make_params(nfeatures=NFEAT, scaleFactor=SCALE, nlevels=LEVELS, threshold=THR, patch_size=PATCH,
                                  device=local_rank, max_width=W, max_height=H, max_batch=F, chunk_frames=args.chunk,
                                  max_keypoints=cap))
    stream = torch.cuda.Stream(device=dev)       # all ORB work and the timing events go to this stream
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    def step_device():
        ctx.detect_and_compute_batch_ptr(d_frames.data_ptr(), 1, F, W, H, PITCH, H * PITCH, cap, d_k.data_ptr(),
                                         d_a.data_ptr(), d_d.data_ptr(), d_n.data_ptr(), 1)

    def step_e2e():
        ctx.detect_and_compute_batch_ptr(h_frames.data_ptr(), 0, F, W, H, PITCH, H * PITCH, cap, h_k.data_ptr(),
                                         h_a.data_ptr(), h_d.data_ptr(), h_n.data_ptr(), 0)

    # copy-only step: the e2e step's bytes over the link, both directions at once, no kernels
    s_up, s_down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    link_start = torch.cuda.Event()

    def step_link():
        link_start.record(stream)
        s_up.wait_event(link_start)
        s_down.wait_event(link_start)
        with torch.cuda.stream(s_up):
            d_frames.copy_(h_frames, non_blocking=True)
        with torch.cuda.stream(s_down):
            h_k.copy_(d_k, non_blocking=True)
            h_a.copy_(d_a, non_blocking=True)
            h_d.copy_(d_d, non_blocking=True)
            h_n.copy_(d_n, non_blocking=True)
        stream.wait_stream(s_up)
        stream.wait_stream(s_down)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        """total ms of `steps` steps (max over ranks) and the per-step ms of this rank"""
        for _ in range(warmup):
            fn()
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        sampler.active = True
        ev[0].record(stream)
        for i in range(steps):
            fn()
            ev[i + 1].record(stream)
        torch.cuda.synchronize()
        sampler.active = False
        ms = ev[0].elapsed_time(ev[steps])
        per = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
        barrier()
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, per

    sampler = ClockSampler(local_rank)
    sampler.start()

    ms_dev, per_dev = timed(step_device, args.steps, args.warmup)
    ctx.synchronize()
    launches_per_step = ctx.launch_count()
    n_kp = int(d_n.sum().item())

    # per-kernel times of the same step (separate pass: the event brackets are not in the headline number)
    ctx.set_profiling(True)
    for _ in range(2):
        step_device()
    ctx.stage_ms()
    prof_steps = max(1, min(args.steps, 5))
    for _ in range(prof_steps):
        step_device()
    st_ms, st_n = ctx.stage_ms()
    ctx.set_profiling(False)

    # SURVEY 8(f) rank 2 (not part of the headline metric): exact Hamming 2-NN between consecutive frames of the batch,
    # descriptors still resident on the device
    d_m = torch.zeros(F - 1, cap, 4, dtype=torch.int32, device=dev) if F > 1 else None
    ms_match = None
    if d_m is not None and not args.hot_only:
        def step_match():
            ctx.match_knn2_batch_ptr(d_d.data_ptr(), d_n.data_ptr(), F, cap, d_m.data_ptr())
        ms_match = timed(step_match, 2, 1)[0] / 2

    # tensor-core roofline of the matcher (expansion kernel included in the time): 2 * 256 operations per descriptor pair actually
    # compared; peak = dense 8-bit rate = 2 x the measured bf16 cuBLAS rate of MEASURED_PEAKS.json (else the nominal 4500 TOP/s)
    match_roofline = None
    if ms_match:
        n_host = d_n.cpu().numpy().astype(np.int64)
        ops = 2.0 * 256.0 * float((n_host[:-1] * n_host[1:]).sum())
        try:
            tpeak, tsrc = 2.0 * float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]), "2 x MEASURED_PEAKS.json bf16_tflops"
        except Exception:
            tpeak, tsrc = 4500.0, "nominal dense INT8 / FP8"
        tach = ops / (ms_match * 1e-3) / 1e12
        match_roofline = {"bound": "tensor", "kernel": "k_match_expand + k_match_tc", "achieved": tach, "peak": tpeak, "unit": "TOP/s",
                          "frac": tach / tpeak, "peak_source": tsrc}

    e2e_steps = max(1, args.steps // 2)
    ms_e2e, per_e2e = timed(step_e2e, e2e_steps, 2)
    assert int(h_n.sum().item()) == n_kp, "host and device paths disagree"
    ms_link, _ = timed(step_link, e2e_steps, 1)
    sampler.stop()

    # SURVEY 8(f) rank 4 (not part of the headline metric): pyramidal LK of the reference's feature_tracking loop on the KITTI
    # fixture pair, from the ORB keypoints of the first frame; host call including both frame uploads and the result copy
    lk = None
    if rank == 0 and world == 1 and not args.hot_only:
        try:
            g = os.path.join(ROOT, "tests", "golden")
            f0, f1 = V.imread_gray8(os.path.join(g, "kitti_000000.png")), V.imread_gray8(os.path.join(g, "kitti_000001.png"))
            ctx_lk = V.Context(V.make_params(nfeatures=3000, max_width=f0.shape[1], max_height=f0.shape[0], max_batch=1, device=local_rank))
            kp0 = ctx_lk.detect_and_compute(f0)[0]
            pts = np.stack([kp0["x"], kp0["y"]], 1).astype(np.float32)
            ctx_lk.lk_track(f0, f1, pts)
            t0 = time.perf_counter()
            for _ in range(10):
                _, st_lk, _ = ctx_lk.lk_track(f0, f1, pts)
            lk = {"ms_per_pair": (time.perf_counter() - t0) / 10 * 1e3, "points": int(len(pts)), "tracked": int(st_lk.sum()),
                  "call": "win 21, 3 levels, 30 iterations, eps 0.01 (reference src/feature_tracking.cpp:174-180)"}
            try:
                import cv2
                t0 = time.perf_counter()
                for _ in range(3):
                    cv2.calcOpticalFlowPyrLK(f0, f1, pts.reshape(-1, 1, 2), None, winSize=(21, 21), maxLevel=3,
                                             criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01), flags=0, minEigThreshold=0.001)
                lk["cv2_ms_per_pair"] = (time.perf_counter() - t0) / 3 * 1e3
                lk["cv2_threads"] = cv2.getNumThreads()
            except ImportError:
                pass
            ctx_lk.close()
        except Exception as e:      # the fixture pair is optional for the headline number
            lk = {"error": str(e)}
        # the same tracker over the whole device-resident batch: the ORB keypoints of frame t tracked into frame t + 1
        # (frames, points and results stay on the device; every frame's pyramid is built once)
        try:
            if F > 1 and args.config == "kitti":
                d_pts = d_k[:F - 1].to(torch.float32).contiguous()
                d_nx = torch.zeros_like(d_pts)
                d_st = torch.zeros(F - 1, cap, dtype=torch.uint8, device=dev)
                d_er = torch.zeros(F - 1, cap, dtype=torch.float32, device=dev)

                def step_lk():
                    ctx.lk_track_batch_ptr(d_frames.data_ptr(), F, W, H, PITCH, H * PITCH, d_pts.data_ptr(), d_n.data_ptr(), cap,
                                           d_nx.data_ptr(), d_st.data_ptr(), d_er.data_ptr())
                ms_lk = timed(step_lk, 2, 1)[0] / 2
                n_trk = int(d_n[:F - 1].sum().item())
                lk = dict(lk or {}, batch={"ms_per_step": ms_lk, "pairs": F - 1, "points": n_trk, "tracked": int(d_st.sum().item()),
                                            "pairs_per_s": (F - 1) / (ms_lk * 1e-3), "points_per_s": n_trk / (ms_lk * 1e-3)})
        except Exception as e:
            lk = dict(lk or {}, batch={"error": str(e)})

    ingest = None
    if rank == 0 and world == 1 and not args.no_ingest and args.config == "kitti":
        pool = min(F, 64)               # 64 distinct files, listed F times (the decode work per step is F files either way)
        idx = np.arange(F) % pool
        pool_frames = h_frames[:pool, :, :W].numpy()
        n_expected = int(h_n[:pool].numpy()[idx].sum())
        ingest = ingest_leg(V, ctx, pool_frames, idx, cap, (h_k, h_a, h_d, h_n), n_expected)

    total_frames = F * world
    value = total_frames * args.steps / (ms_dev * 1e-3)
    e2e_value = total_frames * e2e_steps / (ms_e2e * 1e-3)
    link_value = total_frames * e2e_steps / (ms_link * 1e-3)
    peak, peak_src = peaks()
    # dominant kernel = largest share of the step; its algorithmic bytes per frame ("Kernels"):
    #   k_pyramid : level 0 read once + levels >= 1 written once          = sum(P)
    #   k_fast    : every level read once (box sums / candidates are ours) = sum(P)
    #   k_harris  : 8 B key read + written per candidate (count unknown here; use 4 K)
    #   k_select  : 8 B per candidate read                                 ~ 8 * 4 K
    #   k_describe: 44 B per output record written                         = 44 K
    names = ["k_pyramid", "k_fast", "k_harris", "k_select", "k_describe"]
    kpf = n_kp / F
    alg_bytes = [SUM_P, SUM_P, 64.0 * kpf, 32.0 * kpf, 44.0 * kpf]
    dom = int(np.argmax(st_ms))
    k1_ms = st_ms[dom] / max(1, st_n[dom])                    # average duration of one launch of that kernel
    frames_per_launch = F * prof_steps / max(1, st_n[dom])
    achieved = alg_bytes[dom] * frames_per_launch / (k1_ms * 1e-3) / 1e9
    kp_per_frame = n_kp / F
    b_min = SUM_P + 44.0 * kp_per_frame
    pass_gbs = b_min * F * args.steps / (ms_dev * 1e-3) / 1e9       # per GPU
    stage_share = [m / max(1e-9, sum(st_ms)) for m in st_ms]
    clocks = sampler.summary()

    # DRAM traffic and the instruction-issue roof come from the ncu capture recorded in profiles/latest_ncu.json (ncu cannot
    # run inside a timed bench); the capture names its commit and its frames per launch
    traffic = traffic_src = roof_issue = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "latest_ncu.json")))
        fpl = float(prof.get("_frames_per_launch", 128))
        if prof.get("_config", "kitti") == args.config:
            per_launch = prof.get(names[dom], {}).get("dram_bytes_per_launch")
            if per_launch is not None:
                traffic = per_launch / fpl * frames_per_launch
                traffic_src = "ncu --set full capture %s (commit %s, %d frames per launch), rescaled to this run's frames per launch" % (
                    prof.get("_tag"), prof.get("_commit"), int(fpl))
            ks = [k for k in names if k in prof and "warp_inst" in prof[k]]
            if ks:
                winst = sum(prof[k]["warp_inst"] for k in ks) / fpl          # warp instructions per frame, whole pass
                tsum = sum(prof[k]["time_us"] for k in ks)
                wavg = lambda key: sum(prof[k].get(key, 0.0) * prof[k]["time_us"] for k in ks) / tsum
                sm_hz = (clocks["sm_mhz"] or 1965.0) * 1e6
                sms = torch.cuda.get_device_properties(dev).multi_processor_count
                roof_issue = {"warp_inst_per_frame": winst, "thread_inst_per_pyramid_pixel": winst * 32.0 / SUM_P,
                              "issue_active_pct": wavg("issue_active_pct"), "alu_pct": wavg("alu_pct"), "fma_pct": wavg("fma_pct"),
                              "frac_of_issue_peak": winst * (value / world) / (sms * 4 * sm_hz),
                              "issue_peak_frames_per_s": sms * 4 * sm_hz / winst,
                              "source": "profiles/latest_ncu.json: %s, commit %s (time-weighted over the kernels; peak = SMs x 4 schedulers x SM clock)"
                                        % (prof.get("_tag"), prof.get("_commit"))}
    except Exception:
        pass

    line = {
        "metric": "orb_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "ms_per_step_median": float(np.median(per_dev)),
        "ms_per_step_min": float(min(per_dev)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(F), "preset": args.config, "frames_per_gpu_per_step": F, "levels": LEVELS, "nfeatures": NFEAT,
                   "keypoints_per_frame": kp_per_frame, "chunk_frames": args.chunk, "distinct_frames_per_gpu": F,
                   "cache_hygiene": "input batch %.0f MB > 126 MB L2; scratch arena reused per chunk" % (F * H * PITCH / 1e6),
                   "stage_names": names,
                   "stage_ms_per_step": [m / prof_steps for m in st_ms], "stage_share": stage_share,
                   "match_knn2_ms_per_step": ms_match, "match_pairs_per_step": F - 1, "match_roofline": match_roofline, "ingest_png": ingest, "lk_track": lk,
                   "pass_b_min_bytes_per_frame": b_min, "pass_hbm_gbs_per_gpu": pass_gbs, "pass_hbm_frac": pass_gbs / peak,
                   "peak_source": peak_src},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(F * H * PITCH) * world,
                "d2h_bytes_per_step": int(F * cap * 44 + F * 4) * world,
                "ms_per_step_median": float(np.median(per_e2e)), "ms_per_step_min": float(min(per_e2e))},
        "e2e_link_bound": {"value": link_value, "unit": "frames/s", "frac": e2e_value / link_value,
                           "what": "the e2e step's pinned buffers over the link, H2D and D2H at once, no kernels (max over ranks)"},
        "gpu_launches": int(launches_per_step * args.steps) * world,
        "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": alg_bytes[dom] * frames_per_launch, "avg_launch_ms": k1_ms},
        "roofline_issue": roof_issue,
        "clocks": clocks,
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        frames_host = h_frames.numpy()
        ctx.close()
        ctx = None
        line["parity"] = parity_leg(V, frames_host, local_rank)
        line["cpu_baseline"] = cpu_baseline_leg(V)
    if rank == 0:
        print(json.dumps(line), file=JSON_OUT, flush=True)
    if ctx is not None:
        ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    # stdout carries the JSON line and nothing else: libraries that write to file descriptor 1 themselves (NCCL prints a
    # version banner there) are pointed at stderr, the line goes to a private duplicate of the original stdout
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    main()
