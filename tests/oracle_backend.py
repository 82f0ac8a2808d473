"""CPU backend of open3d_slam_b200.slam's control flow, built on the oracle (TEST INFRASTRUCTURE, never on the product path).

Same method names as slam.DeviceBackend; every operation is the oracle's restatement of the reference:
    step            = ScanToMapIcp::processForScanMatchingAndMerging + scanToMapRegistration + the gates of
                      Mapper::addRangeMeasurement + Submap::insertScan(carving = true) + insertScanDenseMap(carving = true)
    insert_scan     = Submap::insertScan(carving = false) (buffered overlap scans)
    revisit_fitness = SubmapCollection::isSwitchingSubmapsConsistant over a VoxelMap restated with a Python set of keys
"""
from __future__ import annotations

import numpy as np

from oracle import oracle as O


class OracleSubmap:
    def __init__(self, dense_voxel):
        self.xyz = np.zeros((0, 3)); self.nrm = np.zeros((0, 3))
        self.dense = O.DenseMap(dense_voxel, 1 << 22) if dense_voxel else None
        self.nScansInsertedMap = 0
        self.nScansInsertedDenseMap = 0
        self.cropperPose = np.eye(4)      # mapBuilderCropper_ pose: set after every insertion (Submap.cpp:71)
        self.voxel_keys = None            # VoxelMap keys (layer "map")
        self.carved_total = 0
        self.carve_runs = 0
        self.dense_carved_total = 0
        self.dense_carve_runs = 0


class OracleCloud:
    def __init__(self, xyz, nrm=None):
        self.xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        self.nrm = None if nrm is None else np.ascontiguousarray(nrm, dtype=np.float64).reshape(-1, 3)

    def __len__(self):
        return len(self.xyz)


class OracleBackend:
    def __init__(self, params, carving=True, dense=True):
        """params: open3d_slam_b200.engine.MapperParameters (plain dataclass, no device access)."""
        self.p = params
        sp, mb = params.scanProcessing, params.mapBuilder
        self.wide = self._cropper(mb.cropper)
        self.narrow = self._cropper(sp.cropper)
        self.carving, self.dense = carving, dense
        self.pose = np.eye(4)             # Mapper::mapToRangeSensorPrev_
        self._merge = None

    @staticmethod
    def _cropper(cp, center=(0.0, 0.0, 0.0)):
        return O.cropper(cp.cropperName, cp.croppingMinRadius, cp.croppingMaxRadius, cp.croppingMinZ, cp.croppingMaxZ, center=center)

    def _process(self, raw):
        p = self.p
        return O.process_scan(np.asarray(raw, dtype=np.float32).astype(np.float64), self.wide, self.narrow, p.scanProcessing.voxelSize, p.icp.knn,
                              p.icp.maxDistanceKnn, p.scanProcessing.downSamplingRatio, p.seed)

    # -- submaps
    def new_submap(self):
        return OracleSubmap(self.p.denseMapVoxelSize if self.dense else None)

    def _insert(self, sm, mx, mn, T):
        if len(mx) == 0:
            return
        sm.xyz, sm.nrm = O.submap_insert_scan(sm.xyz, sm.nrm, mx, mn, T, self.p.mapBuilder.mapVoxelSize, self.wide)
        sm.cropperPose = np.array(T)
        sm.nScansInsertedMap += 1

    def _carve(self, sm, raw64, T):
        """Submap::carve (Submap.cpp:109-123)"""
        cp = self.p.mapBuilder.carving
        if len(sm.xyz) == 0 or not (sm.nScansInsertedMap % cp.carveSpaceEveryNscans == 1):
            return
        scan_map, _ = O.transform(T, raw64)
        crop = self._cropper(self.p.mapBuilder.cropper, center=sm.cropperPose[:3, 3])
        rm = O.carve(sm.xyz, sm.nrm, scan_map, T[:3, 3], crop, cp.voxelSize, cp.maxRaytracingLength, cp.truncationDistance, cp.minDotProductWithNormal)
        sm.xyz, sm.nrm = sm.xyz[~rm], sm.nrm[~rm]
        sm.carve_runs += 1; sm.carved_total += int(rm.sum())

    def _dense(self, sm, raw64, T):
        """Submap::insertScanDenseMap(raw, T, carving = true)   Submap.cpp:77-92"""
        dc = self._cropper(self.p.denseMapCropper)
        cx, _ = O.crop(dc, raw64)
        tx, _ = O.transform(T, cx)                 # o3d_slam::transform, duplication quirk included
        sm.dense.insert(tx)
        cp = self.p.denseMapCarving
        if self.carving and sm.nScansInsertedDenseMap % cp.carveSpaceEveryNscans == 1:   # Submap.cpp:127 (the map is not empty here)
            n = sm.dense.carve(raw64, T[:3, 3], self.p.denseMapVoxelSize, cp.neighborhoodRadiusDenseMap, cp.truncationDistance, cp.maxRaytracingLength)
            sm.dense_carve_runs += 1; sm.dense_carved_total += n
        sm.nScansInsertedDenseMap += 1

    def first_scan(self, sm, raw):
        (mx, mn), _ = self._process(raw)
        self._insert(sm, mx, mn, np.eye(4))
        self.pose = np.eye(4)
        return OracleCloud(mx, mn)

    def step(self, sm, raw, odometryMotion):
        p = self.p
        raw64 = np.asarray(raw, dtype=np.float32).astype(np.float64)
        (mx, mn), (ax, an) = self._process(raw)
        self._merge = OracleCloud(mx, mn)
        patch = self._cropper(p.scanProcessing.cropper, center=self.pose[:3, 3])     # ScanToMapRegistration.cpp:58
        px, pn = O.crop(patch, sm.xyz, sm.nrm)
        res = O.registration_icp_p2plane(ax, px, pn, p.icp.maxCorrespondenceDistance, self.pose @ np.asarray(odometryMotion), max_iter=p.icp.maxNumIter)
        res.transformation_ = res.T; res.fitness_ = res.fitness; res.inlier_rmse_ = res.inlier_rmse
        accepted = p.isIgnoreMinRefinementFitness or not (res.fitness < p.minRefinementFitness)
        if accepted:
            self.pose = np.array(res.T)
            if self.carving:
                self._carve(sm, raw64, self.pose)
            self._insert(sm, mx, mn, self.pose)
            if self.dense:
                self._dense(sm, raw64, self.pose)
        return res, bool(accepted)

    def last_merge_cloud(self):
        return self._merge

    def insert_scan(self, sm, cloud, T):
        self._insert(sm, cloud.xyz, cloud.nrm, np.asarray(T))

    def set_pose(self, sm, T):
        pass

    def map_cloud(self, sm):
        return sm.xyz, sm.nrm

    def map_center(self, sm):
        return sm.xyz.mean(axis=0) if len(sm.xyz) else np.zeros(3)

    def build_voxel_map(self, sm):
        v = 2.5 * self.p.mapBuilder.mapVoxelSize
        inv = 1.0 / v
        k = np.floor(sm.xyz * inv).astype(np.int64)          # getVoxelIdx: floor(p * inverseVoxelSize)
        sm.voxel_keys = set(map(tuple, k))
        sm.voxel_inv = inv

    def revisit_fitness(self, sm, scan, T):
        if len(scan) == 0:
            return 0.0
        T = np.asarray(T)
        q = scan.xyz
        # Eigen isometry * vector: (R p) + t with the row sums associated left to right
        x = ((T[0, 0] * q[:, 0] + T[0, 1] * q[:, 1]) + T[0, 2] * q[:, 2]) + T[0, 3]
        y = ((T[1, 0] * q[:, 0] + T[1, 1] * q[:, 1]) + T[1, 2] * q[:, 2]) + T[1, 3]
        z = ((T[2, 0] * q[:, 0] + T[2, 1] * q[:, 1]) + T[2, 2] * q[:, 2]) + T[2, 3]
        k = np.floor(np.c_[x, y, z] * sm.voxel_inv).astype(np.int64)
        hits = sum(1 for t in map(tuple, k) if t in sm.voxel_keys)
        return hits / len(scan)

    # -- loop-closure refinement
    def submap_as_cloud(self, sm):
        return OracleCloud(sm.xyz, sm.nrm)

    def overlap(self, source, target, T0, voxel, min_pts):
        fs, ft = O.overlap_flags(source.xyz, target.xyz, T0, voxel, min_pts)
        return OracleCloud(source.xyz[fs], source.nrm[fs]), OracleCloud(target.xyz[ft], target.nrm[ft])

    def register_batch(self, sources, targets, inits, max_corr, max_iter):
        out = []
        for s, t, T0 in zip(sources, targets, inits):
            r = O.registration_icp_p2plane(s.xyz, t.xyz, t.nrm, max_corr, T0, max_iter=max_iter)
            r.transformation_ = r.T; r.fitness_ = r.fitness; r.inlier_rmse_ = r.inlier_rmse
            out.append(r)
        return out

    def information_matrix(self, source, target, max_corr, T):
        return O.information_matrix(source.xyz, target.xyz, max_corr, T)

    def cloud_size(self, c):
        return len(c)

    def dense_map(self, sm):
        x, _n, k = sm.dense.to_cloud()
        return x, k

    def counters(self, sm):
        return {"inserted_map": sm.nScansInsertedMap, "inserted_dense": sm.nScansInsertedDenseMap, "carve_runs": sm.carve_runs,
                "carved_points_total": sm.carved_total, "dense_carve_runs": sm.dense_carve_runs, "carved_voxels_total": sm.dense_carved_total}

    def close(self):
        pass
