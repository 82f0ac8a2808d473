"""Property tests (hypothesis) of the CPU oracle -- SURVEY.md 8c test 7.  They hold for any input, so they also guard the
oracle's edge cases (duplicates, points on voxel faces, tiny clouds) that the fixed fixtures do not visit."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import oracle as O

pts = st.integers(min_value=0, max_value=2 ** 31 - 1).map(
    lambda seed: np.random.default_rng(seed).uniform(-3, 3, (int(np.random.default_rng(seed + 1).integers(1, 400)), 3)))
SET = dict(max_examples=25, deadline=None)


@settings(**SET)
@given(pts, st.sampled_from([0.05, 0.25, 1.0]))
def test_voxel_down_sample_is_a_partition(xyz, voxel):
    xyz = np.round(xyz / (voxel / 4)) * (voxel / 4)            # many points exactly on voxel faces, many duplicates
    out, _, keys = O.voxel_down_sample(xyz, voxel, return_keys=True)
    vmin = xyz.min(axis=0) - 0.5 * voxel
    k = np.floor((xyz - vmin) / voxel).astype(np.int64)
    uk, inv, cnt = np.unique(k, axis=0, return_inverse=True, return_counts=True)
    assert len(out) == len(uk)                                  # one output point per occupied voxel
    order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
    assert np.array_equal(keys[order], uk)
    sums = np.zeros((len(uk), 3)); np.add.at(sums, inv.reshape(-1), xyz)
    assert np.abs(out[order] - sums / cnt[:, None]).max() < 1e-12   # the mean of its members
    assert np.abs((out[order] * cnt[:, None]).sum(axis=0) - xyz.sum(axis=0)).max() < 1e-9     # mass is conserved


@settings(**SET)
@given(pts, st.floats(0.5, 4.0), st.booleans())
def test_crop_partitions_the_cloud(xyz, rmax, invert):
    c = O.cropper("MaxRadius", 0.0, rmax, center=(0.2, -0.1, 0.0), invert=invert)
    a, _ = O.crop(c, xyz)
    b, _ = O.crop(O.cropper("MaxRadius", 0.0, rmax, center=(0.2, -0.1, 0.0), invert=not invert), xyz)
    assert len(a) + len(b) == len(xyz)
    inside = np.sqrt(((xyz - [0.2, -0.1, 0.0]) ** 2).sum(axis=1)) <= rmax
    assert np.array_equal(a, xyz[inside != invert])             # order preserved


@settings(**SET)
@given(pts, st.floats(0.05, 0.95), st.integers(0, 2 ** 31 - 1))
def test_seeded_down_sample_size_order_and_permutation_independence(xyz, ratio, seed):
    xyz = np.unique(xyz, axis=0)
    out, _ = O.random_down_sample(xyz, ratio, seed)
    assert len(out) == int(len(xyz) * ratio)                    # [O3D]: size_t(n * ratio)
    idx = [np.flatnonzero((xyz == p).all(axis=1))[0] for p in out]
    assert idx == sorted(idx)                                   # original order kept
    perm = np.random.default_rng(seed).permutation(len(xyz))
    out2, _ = O.random_down_sample(xyz[perm], ratio, seed)      # the subset depends on the points, not on their order
    assert {tuple(p) for p in out} == {tuple(p) for p in out2}


@settings(**SET)
@given(pts, st.integers(0, 2 ** 31 - 1))
def test_icp_result_is_bounded_and_deterministic(xyz, seed):
    rng = np.random.default_rng(seed)
    nrm = rng.normal(size=xyz.shape); nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    src = xyz[: max(1, len(xyz) // 2)] + rng.normal(0, 0.02, (max(1, len(xyz) // 2), 3))
    r1 = O.registration_icp_p2plane(src, xyz, nrm, 0.3, np.eye(4), max_iter=10)
    r2 = O.registration_icp_p2plane(src, xyz, nrm, 0.3, np.eye(4), max_iter=10)
    assert 0.0 <= r1.fitness <= 1.0 and 0.0 <= r1.inlier_rmse <= 0.3 and 1 <= r1.iters <= 10
    assert np.array_equal(r1.T, r2.T) and r1.n_corr == r2.n_corr           # bit-reproducible (fixed reduction order)
    if np.isfinite(r1.T).all():
        R = r1.T[:3, :3]
        assert np.abs(R @ R.T - np.eye(3)).max() < 1e-6 and abs(np.linalg.det(R) - 1.0) < 1e-6


@settings(**SET)
@given(pts, st.floats(0.1, 0.5))
def test_overlap_flags_are_symmetric_in_the_voxel(xyz, voxel):
    half = len(xyz) // 2
    src, tgt = xyz[:half], xyz[half:]
    if len(src) == 0 or len(tgt) == 0:
        return
    fs, ft = O.overlap_flags(src, tgt, np.eye(4), voxel, 1)
    ks = {tuple(k) for k in np.floor(src[fs] * (1.0 / voxel)).astype(np.int64)}
    kt = {tuple(k) for k in np.floor(tgt[ft] * (1.0 / voxel)).astype(np.int64)}
    assert ks == kt                                             # the selected points of both clouds occupy the same voxels
