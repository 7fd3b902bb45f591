"""CPU-side checks of the product: the C-ABI library builds, loads and exports every symbol include/orb_b200.h
declares; it fails loudly without a GPU (no CPU fallback); frame sharding logic incl. a world_size-2 gloo run."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(V):
    hdr = open(os.path.join(ROOT, "include", "orb_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(orb_[a-z_0-9]+)\s*\(", hdr)) | {"bit_pattern_31_"}
    assert declared == set(V.EXPORTS), declared ^ set(V.EXPORTS)
    lib = ctypes.CDLL(V.lib_path())
    for name in declared:
        assert hasattr(lib, name), name
    assert V.load_library().orb_abi_version() == 1


def test_exported_pattern_table(V, O):
    lib = ctypes.CDLL(V.lib_path())
    pat = np.ctypeslib.as_array((ctypes.c_int * 1024).in_dll(lib, "bit_pattern_31_"))
    assert np.array_equal(pat, O.pattern().astype(np.int32))


def test_params_struct_layout(V):
    p = V.default_params()
    assert ctypes.sizeof(V.Params) == 20 * 4
    assert (p.nfeatures, p.nlevels, p.fast_threshold, p.fast_n, p.nms_window, p.orient_patch) == (500, 8, 20, 9, 3, 31)
    assert abs(p.scale_factor - 1.2) < 1e-6 and abs(p.harris_k - 0.04) < 1e-7 and p.select_policy == 1


def test_create_rejects_bad_params_and_has_no_cpu_fallback(V):
    import torch
    for kw in (dict(nlevels=0), dict(nlevels=17), dict(n=0), dict(nms_window=5), dict(patch_size=0),
               dict(nfeatures=0), dict(max_width=0), dict(nfeatures=100000)):
        with pytest.raises(V.OrbError) as e:
            V.Context(V.make_params(**kw))
        assert e.value.code == -1
    if not torch.cuda.is_available():
        with pytest.raises(V.OrbError) as e:
            V.ORB()
        assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "visual-odometry-gpu_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in src and "orb_oracle" not in src and "oracle/" not in src, f
    for f in os.listdir(os.path.join(ROOT, "include")):
        assert not re.search(r"#\s*include[^\n]*oracle", open(os.path.join(ROOT, "include", f)).read()), f


def test_shard_range(V):
    for n in (0, 1, 7, 8, 1000, 1001):
        for ws in (1, 2, 3, 4, 8):
            spans = [V.shard_range(n, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) == -(-n // ws)
    with pytest.raises(ValueError):
        V.shard_range(10, 2, 2)


def test_synth_frames_are_deterministic(V):
    a = V.synth_frames(3, 1241, 376)
    b = V.synth_frames(2, 1241, 376, start=1)
    assert a.shape == (3, 376, 1241) and np.array_equal(a[1:], b)
    assert not np.array_equal(a[0], a[1])
    c = V.synth_frames(1, 640, 480, pitch=704)
    assert c.shape == (1, 480, 704) and not c[0, :, 640:].any()


GLOO_WORKER = r"""
import os, sys, importlib
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
V = importlib.import_module("visual-odometry-gpu_b200")
from oracle import pyoracle as O
dist.init_process_group("gloo")
rank, ws = dist.get_rank(), dist.get_world_size()
n = 5
lo, hi = V.shard_range(n, ws, rank)
frames = V.synth_frames(hi - lo, 320, 96, start=lo)
p = O.params(nfeatures=200, nlevels=3, fast_threshold=20)
counts = torch.zeros(n, dtype=torch.int32)
if hi > lo:
    counts[lo:hi] = torch.from_numpy(O.detect_and_compute_batch(frames, p, 200, 1))
dist.all_reduce(counts)                      # host-side gather of per-frame slots (no data-path collective)
t = torch.tensor([float(rank + 1)])
dist.all_reduce(t, op=dist.ReduceOp.MAX)     # max-over-ranks timing reduction used by bench.py
if rank == 0:
    ref = O.detect_and_compute_batch(V.synth_frames(n, 320, 96), p, 200, 1)
    assert np.array_equal(counts.numpy(), ref), (counts, ref)
    assert t.item() == ws
    print("GLOO_OK", counts.tolist())
dist.destroy_process_group()
"""


def test_frame_sharding_world_size_2_gloo(tmp_path):
    """N>1 host logic on CPU: two ranks shard a frame batch, results land in per-frame slots."""
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script), ROOT],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "GLOO_OK" in r.stdout, r.stdout[-3000:]


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU port on the host cores) must put exactly one JSON line on stdout, with the
    keys the driver reads; anything a library writes to file descriptor 1 is routed to stderr."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "orb_frames_per_s" and d["unit"] == "frames/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("kitti_synth_1241x376_L8_N2000")


@pytest.mark.parametrize("n,wave", [(1000, 64), (1000, 128), (256, 14), (64, 4), (64, 2), (7, 64), (256, 64), (255, 64), (1, 1), (5, 1), (0, 8)])
def test_wave_schedule_host_logic(V, n, wave):
    """Batch pipeline, host side (csrc/orb_api.cu wave_schedule): every frame in exactly one wave, no wave above the wave
    size; host-staged batches ramp up (so the kernels start early) and, when results go back to the host, ramp down (so
    little is left to compute after the last frame has arrived)."""
    for up in (False, True):
        for down in (False, True):
            b = V.orb.wave_schedule(n, wave, up, down)
            assert b[0] == 0 or n == 0
            assert b[-1] == n
            sizes = np.diff(b)
            assert (sizes > 0).all() and sizes.sum() == n and (sizes <= wave).all()
            if not up and not down and n:
                assert (sizes[:-1] == wave).all()
            if up and n > wave:
                assert sizes[0] == max(1, wave // 8)
                k = 1
                while k < len(sizes) and sizes[k - 1] < wave and sizes[k] <= wave and sizes[k] == 2 * sizes[k - 1]:
                    k += 1
                assert sizes[k - 1] >= min(wave, sizes[0] * 2 ** (k - 1))
            if down and n >= 4 * wave and wave >= 2:
                assert sizes[-1] <= max(1, wave // 2)
                assert sizes[-1] <= sizes[-2] or sizes[-2] <= wave // 2
