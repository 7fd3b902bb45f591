"""exact Hamming 2-NN + ratio test on the device against the CPU oracle (bit-exact: integer work)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(V):
    c = V.Context(V.make_params(nfeatures=2000, nlevels=8, max_width=1241, max_height=376, max_batch=4, max_keypoints=2000))
    yield c
    c.close()


def _as4(m):
    return np.stack([m["idx1"], m["dist1"], m["idx2"], m["dist2"]], axis=-1)


def test_knn2_random_with_ties(V, O, ctx):
    rng = np.random.default_rng(0)
    t = rng.integers(0, 256, (777, 32), dtype=np.uint8)
    q = rng.integers(0, 256, (301, 32), dtype=np.uint8)
    q[:50] = t[rng.integers(0, 777, 50)]            # exact duplicates -> distance 0
    t[100] = t[5]; t[200] = t[5]; q[60] = t[5]      # three identical train descriptors -> ties go to the lower index
    q[61] = t[5] ^ np.eye(1, 32, 3, dtype=np.uint8)[0]
    got = ctx.match_knn2(q, t)
    ref, keep = O.match_knn2(q, t)
    assert np.array_equal(_as4(got), ref)
    assert got["idx1"][60] == 5 and got["idx2"][60] == 100 and got["dist2"][60] == 0
    assert np.array_equal(ctx.ratio_test(got), keep)


@pytest.mark.parametrize("nq,nt", [(1, 0), (5, 1), (3, 2), (130, 64), (129, 65), (0, 10),
                                   # the kernel's shapes: 256 queries per work item (two blocks of 128), train tiles of 128
                                   (128, 128), (127, 129), (256, 127), (257, 256), (255, 257), (513, 383), (700, 1)])
def test_knn2_edge_sizes(V, O, ctx, nq, nt):
    rng = np.random.default_rng(nq * 100 + nt)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    got = ctx.match_knn2(q, t)
    ref, keep = O.match_knn2(q, t)
    assert np.array_equal(_as4(got), ref.reshape(-1, 4))
    assert np.array_equal(ctx.ratio_test(got), keep)
    if nt < 2 and nq:
        assert (got["idx2"] == -1).all() and (got["dist2"] == 2**31 - 1).all() and not ctx.ratio_test(got).any()


def test_match_consecutive_frames_of_a_batch(V, O, ctx, kitti0, kitti1):
    """The VO use: descriptors of frame t against frame t+1 (reference src/feature_tracking.cpp:201-219)."""
    frames = np.stack([kitti0, kitti1, kitti0, V.synth_frames(1, 1241, 376)[0]])
    k, a, d, n = ctx.detect_and_compute_batch(frames, 2000)
    m = ctx.match_knn2_batch(d, n)
    assert m.shape == (3, 2000)
    for p in range(3):
        ref, keep = O.match_knn2(d[p, :n[p]], d[p + 1, :n[p + 1]])
        assert np.array_equal(_as4(m[p, :n[p]]), ref)
        assert np.array_equal(ctx.ratio_test(m[p, :n[p]]), keep)
    # consecutive KITTI frames share most of the scene: a healthy share of the ratio-test survivors
    good = ctx.ratio_test(m[0, :n[0]])
    assert 100 < good.sum() < n[0]


def test_match_device_resident(V, O, ctx):
    """Descriptors never leave the device between detect and match."""
    import torch
    F, W, H, cap, pitch = 3, 1241, 376, 2000, 1248
    frames = V.synth_frames(F, W, H, pitch=pitch)
    d_frames = torch.from_numpy(frames).cuda()
    d_k = torch.zeros(F, cap, 2, dtype=torch.int32, device="cuda")
    d_a = torch.zeros(F, cap, dtype=torch.float32, device="cuda")
    d_d = torch.zeros(F, cap, 32, dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(F, dtype=torch.int32, device="cuda")
    d_m = torch.full((F - 1, cap, 4), -7, dtype=torch.int32, device="cuda")
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.detect_and_compute_batch_ptr(d_frames.data_ptr(), 1, F, W, H, pitch, H * pitch, cap, d_k.data_ptr(), d_a.data_ptr(),
                                     d_d.data_ptr(), d_n.data_ptr(), 1)
    ctx.match_knn2_batch_ptr(d_d.data_ptr(), d_n.data_ptr(), F, cap, d_m.data_ptr())
    ctx.synchronize()
    ctx.use_own_stream()
    n = d_n.cpu().numpy()
    d = d_d.cpu().numpy()
    m = d_m.cpu().numpy()
    for p in range(F - 1):
        ref, _ = O.match_knn2(d[p, :n[p]], d[p + 1, :n[p + 1]])
        assert np.array_equal(m[p, :n[p]], ref)
        assert (m[p, n[p]:] == -7).all()


def test_ragged_batch_with_empty_and_tiny_frames(V, O, ctx):
    """Pairs of a batch whose frames hold 0, 1, 2 or a full set of descriptors: every work item of the persistent kernel sees its
    own (query count, train count); an absent neighbour is (-1, INT32_MAX)."""
    rng = np.random.default_rng(5)
    cap = 300
    n = np.array([300, 0, 129, 1, 2, 257, 300, 128], dtype=np.int32)
    d = rng.integers(0, 256, (len(n), cap, 32), dtype=np.uint8)
    d[6, :40] = d[5, 100:140]                       # exact matches across a pair
    d[6, 200] = d[6, 7]                             # and a tie inside a train set
    m = ctx.match_knn2_batch(d, n)
    assert m.shape == (len(n) - 1, cap)
    for p in range(len(n) - 1):
        ref, _ = O.match_knn2(d[p, :n[p]], d[p + 1, :n[p + 1]])
        assert np.array_equal(_as4(m[p, :n[p]]), ref.reshape(-1, 4)), p


def test_many_pairs_walk_the_persistent_grid(V, O, ctx):
    """More work items than SMs: each CTA walks over several items and its pipeline state carries over from one to the next."""
    rng = np.random.default_rng(6)
    F, cap = 80, 520                                 # 79 pairs x 3 query blocks = 237 items
    n = rng.integers(cap - 140, cap + 1, F).astype(np.int32)
    d = rng.integers(0, 256, (F, cap, 32), dtype=np.uint8)
    m = ctx.match_knn2_batch(d, n)
    for p in (0, 1, 37, 78):
        ref, _ = O.match_knn2(d[p, :n[p]], d[p + 1, :n[p + 1]])
        assert np.array_equal(_as4(m[p, :n[p]]), ref.reshape(-1, 4)), p
