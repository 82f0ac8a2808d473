"""Host control flow of the full mapper (config 5 of BASELINE.json) around the device hot path.

What stays on the host in the reference and is restated here as plain control flow (paths relative to
/root/reference/open3d_slam/open3d_slam/):
    Mapper::addRangeMeasurement                                   src/Mapper.cpp:101-181
    SubmapCollection::insertScan / updateActiveSubmap / createNewSubmap / insertBufferedScans / findClosestSubmap /
    isSwitchingSubmapsConsistant                                  src/SubmapCollection.cpp:83-131,133-158,172-207,352-364
    Submap::computeSubmapCenter / computeFeatures (voxel map part) src/Submap.cpp:228-259
    the refinement half of PlaceRecognition::buildLoopClosureConstraints (overlap -> ICP -> information matrix; the
    FPFH / RANSAC proposal in front of it is out of scope)   src/PlaceRecognition.cpp:96-149
Every arithmetic step is a call into a *backend*: `DeviceBackend` (below) drives libb2s.so; the parity tests run the very
same control flow over a CPU backend built on the oracle (tests/oracle_backend.py) -- the product never imports it.

The per-scan step of the device backend is ONE C call with host buffers (float32 scan in, RegistrationResult out) that
replays the captured CUDA graph of the chain S1 -> S2 -> gates -> [carving] -> F1 -> [dense map]; the host decisions
(submap hand-over, revisits) are taken from the returned result, like the reference's mapping thread does.
"""
from __future__ import annotations

import collections
from dataclasses import dataclass, field

import numpy as np

from . import engine as E


@dataclass
class SubmapParameters:
    """include/open3d_slam/Parameters.hpp:100-106"""
    radius: float = 20.0
    minNumRangeData: int = 5
    adjacencyBasedRevisitingMinFitness: float = 0.4
    numScansOverlap: int = 3


@dataclass
class LoopClosureParameters:
    """the PlaceRecognitionParameters the refinement half reads (Parameters.hpp:132-135) + magic.hpp:14"""
    maxIcpCorrespondenceDistance: float = 0.3
    minRefinementFitness: float = 0.7
    maxNumIter: int = 100                       # magic::icpRunUntilConvergenceNumberOfIterations
    voxelExpansionFactorOverlapComputation: float = 20.0
    minNumPointsPerVoxel: int = 1


VOXEL_EXPANSION_ADJACENCY_REVISITING = 2.5     # magic::voxelExpansionFactorAdjacencyBasedRevisiting
VOXEL_MAP_LAYER = "map"                        # Submap::voxelMapLayer


@dataclass
class SubmapRecord:
    """What SubmapCollection knows about one Submap besides its clouds."""
    handle: object                      # backend submap object
    id: int
    parent: int
    origin: np.ndarray                  # mapToSubmap_ translation at creation
    center: np.ndarray | None = None    # computeSubmapCenter() once finished
    has_voxel_map: bool = False

    def mapToSubmapCenter(self) -> np.ndarray:   # Submap::getMapToSubmapCenter
        return self.center if self.center is not None else self.origin


class SubmapCollection:
    """src/SubmapCollection.cpp, the parts Mapper::addRangeMeasurement reaches (no place recognition, no optimisation)."""

    def __init__(self, backend, params: SubmapParameters):
        self.backend = backend
        self.params = params
        self.submaps: list[SubmapRecord] = []
        self.activeSubmapIdx = 0
        self.numScansMergedInActiveSubmap = 0
        assert params.numScansOverlap <= 14   # DeviceBackend recycles merge_ clouds through a ring of 16
        self.overlapScansBuffer = collections.deque(maxlen=params.numScansOverlap)   # CircularBuffer, :216
        self.finishedSubmapsIdxs: list[int] = []
        self.adjacency: set[tuple[int, int]] = set()
        self.events: list[tuple] = []   # (scan index, what, ...) -- compared between backends by the parity test

    # -- helpers
    def getActiveSubmap(self) -> SubmapRecord:
        return self.submaps[self.activeSubmapIdx]

    def isAdjacent(self, a: int, b: int) -> bool:
        return (min(a, b), max(a, b)) in self.adjacency

    def createNewSubmap(self, mapToSubmap: np.ndarray) -> None:   # :133-145
        rec = SubmapRecord(self.backend.new_submap(), len(self.submaps), self.activeSubmapIdx, np.array(mapToSubmap[:3, 3], dtype=np.float64))
        self.submaps.append(rec)
        self.activeSubmapIdx = len(self.submaps) - 1
        self.numScansMergedInActiveSubmap = 0

    def findClosestSubmap(self, mapToRangeSensor: np.ndarray) -> int:   # :147-158 (std::min_element: first minimum)
        p0 = mapToRangeSensor[:3, 3]
        d = [float(np.linalg.norm(p0 - s.mapToSubmapCenter())) for s in self.submaps]
        return int(np.argmin(d))

    def updateActiveSubmap(self, mapToRangeSensor: np.ndarray, scan) -> None:   # :94-131
        if self.numScansMergedInActiveSubmap < self.params.minNumRangeData:
            return
        closest = self.findClosestSubmap(mapToRangeSensor)
        closestSubmap, activeSubmap = self.submaps[closest], self.submaps[self.activeSubmapIdx]
        p = mapToRangeSensor[:3, 3]
        if np.linalg.norm(p - closestSubmap.mapToSubmapCenter()) < self.params.radius:
            if closest == self.activeSubmapIdx:
                return
            if self.isAdjacent(closestSubmap.id, activeSubmap.id) and self.isSwitchingSubmapsConsistant(scan, closest, mapToRangeSensor):
                self.activeSubmapIdx = closest
            elif np.linalg.norm(p - activeSubmap.mapToSubmapCenter()) > self.params.radius:
                self.createNewSubmap(mapToRangeSensor)
        else:
            self.createNewSubmap(mapToRangeSensor)

    def isSwitchingSubmapsConsistant(self, scan, candidate: int, mapToRangeSensor: np.ndarray) -> bool:   # :352-364
        rec = self.submaps[candidate]
        if not rec.has_voxel_map:          # a default-constructed VoxelMap is empty: no point hits a voxel
            fitness = 0.0
        else:
            fitness = self.backend.revisit_fitness(rec.handle, scan, mapToRangeSensor)
        self.events.append(("revisit_check", candidate, fitness))
        return fitness > self.params.adjacencyBasedRevisitingMinFitness

    def finishSubmap(self, idx: int) -> None:
        """computeSubmapCenter (:184) and the voxel-map half of Submap::computeFeatures (Submap.cpp:233-237), which the
        reference runs asynchronously once a submap is finished (SlamWrapper::computeFeaturesIfReady)."""
        rec = self.submaps[idx]
        rec.center = self.backend.map_center(rec.handle)
        self.backend.build_voxel_map(rec.handle)
        rec.has_voxel_map = True

    # -- the call the mapper makes after an accepted registration (:172-207).  The scan itself has ALREADY been fused into the
    # submap that was active during the registration (with carving), because the device chain does that without a host
    # round trip; in the reference that is either branch of :185-204 (prevActive.insertScan / active.insertScan, both on
    # the submap that was active during the registration).
    def afterInsertion(self, scan_index: int, preProcessedScan, mapToRangeSensor: np.ndarray) -> None:
        if len(self.submaps) == 0:
            raise RuntimeError("SubmapCollection: no submap")
        self.overlapScansBuffer.append((preProcessedScan, np.array(mapToRangeSensor)))   # addScanToBuffer :83-85
        prev = self.activeSubmapIdx
        self.updateActiveSubmap(mapToRangeSensor, preProcessedScan)
        if prev != self.activeSubmapIdx:
            self.finishSubmap(prev)
            self.finishedSubmapsIdxs.append(prev)
            self.numScansMergedInActiveSubmap = 0
            a, b = self.submaps[prev].id, self.submaps[self.activeSubmapIdx].id
            self.adjacency.add((min(a, b), max(a, b)))
            self.events.append(("active_submap_changed", scan_index, prev, self.activeSubmapIdx))
            while self.overlapScansBuffer:   # insertBufferedScans :87-92 (no carving)
                cloud, T = self.overlapScansBuffer.popleft()
                self.backend.insert_scan(self.submaps[self.activeSubmapIdx].handle, cloud, T)
            self.backend.set_pose(self.submaps[self.activeSubmapIdx].handle, mapToRangeSensor)
        self.numScansMergedInActiveSubmap += 1


class SegmentMapper:
    """Mapper::addRangeMeasurement over a SubmapCollection, one scan per call (src/Mapper.cpp:101-181)."""

    def __init__(self, backend, submapParams: SubmapParameters | None = None):
        self.backend = backend
        self.submaps = SubmapCollection(backend, submapParams or SubmapParameters())
        self.mapToRangeSensor = np.eye(4)
        self.results: list = []
        self.poses: list[np.ndarray] = []
        self._k = 0

    def addRangeMeasurement(self, rawScanF32: np.ndarray, odometryMotion: np.ndarray):
        sc = self.submaps
        k = self._k
        self._k += 1
        if not sc.submaps:   # Mapper.cpp:105-114 / SubmapCollection.cpp:176-181
            sc.createNewSubmap(self.mapToRangeSensor)
            merge = self.backend.first_scan(sc.getActiveSubmap().handle, rawScanF32)
            sc.numScansMergedInActiveSubmap += 1
            self.poses.append(self.mapToRangeSensor.copy())
            self.results.append(None)
            del merge
            return None
        active = sc.getActiveSubmap()
        res, inserted = self.backend.step(active.handle, rawScanF32, odometryMotion)
        self.results.append(res)
        if inserted:
            self.mapToRangeSensor = np.array(res.transformation_, dtype=np.float64)
            sc.afterInsertion(k, self.backend.last_merge_cloud(), self.mapToRangeSensor)
        self.poses.append(self.mapToRangeSensor.copy())
        return res


def refineLoopClosures(backend, source_handle, target_handles, initial_guesses, mapVoxelSize: float, p: LoopClosureParameters | None = None):
    """The refinement half of PlaceRecognition::buildLoopClosureConstraints for one finished (source) submap against its
    candidate (target) submaps, src/PlaceRecognition.cpp:96-149: overlap selection with voxel = 20 x map voxel, ICP of the
    overlapping parts from the proposal, fitness gate, information matrix.  The n registrations run as one batch.
    Returns a list of dicts {overlap sizes, result, accepted, information}."""
    p = p or LoopClosureParameters()
    voxel = p.voxelExpansionFactorOverlapComputation * mapVoxelSize
    pairs = [backend.overlap(source_handle, t, T0, voxel, p.minNumPointsPerVoxel) for t, T0 in zip(target_handles, initial_guesses)]
    results = backend.register_batch([so for so, _to in pairs], [to for _so, to in pairs], initial_guesses, p.maxIcpCorrespondenceDistance, p.maxNumIter)
    out = []
    for (so, to), r in zip(pairs, results):
        acc = not (r.fitness_ < p.minRefinementFitness)
        info = backend.information_matrix(so, to, p.maxIcpCorrespondenceDistance, r.transformation_) if acc else None
        out.append({"n_source_overlap": backend.cloud_size(so), "n_target_overlap": backend.cloud_size(to), "result": r, "accepted": acc,
                    "information": info})
    return out


# ----------------------------------------------------------------------------------------------------------------------
# device backend
# ----------------------------------------------------------------------------------------------------------------------
class DeviceBackend:
    """All arithmetic on libb2s.so (one handle / one CUDA stream = one robot)."""

    def __init__(self, params: E.MapperParameters | None = None, device: int = 0, cuda_stream: int | None = None, submap_capacity: int = 900_000,
                 carving: bool = True, dense: bool = True, graph: bool = True, raw_capacity: int = 65536):
        self.params = params or E.MapperParameters()
        self.eng = E.Engine(self.params, device=device, cuda_stream=cuda_stream)
        self.mapper = E.Mapper(self.eng, 1024)     # its own first submap is a placeholder: submaps are created by new_submap()
        self.mapper.submap.free()
        self.mapper.submap = None
        self.submap_capacity = submap_capacity
        self.carving, self.dense, self.graph, self.raw_capacity = carving, dense, graph, raw_capacity
        self._voxel_maps = {}
        self._stagings = {}
        self._pin = None

    # -- submaps
    def new_submap(self):
        p = self.params
        sm = E.Submap(self.eng, self.submap_capacity)
        sm.setMapperOptions(minMovement=p.minMovementBetweenMappingSteps, carving=p.mapBuilder.carving if self.carving else None,
                            dense=self.dense, denseCarving=p.denseMapCarving if (self.dense and self.carving) else None,
                            denseCropper=p.denseMapCropper)
        return sm

    def _activate(self, sm):
        self.mapper.submap = sm
        if self.graph:
            if id(sm) not in self._stagings:
                self._stagings[id(sm)] = self.mapper.enableGraph(self.raw_capacity)
            self.mapper._staging = self._stagings[id(sm)]

    def _pinned(self, raw: np.ndarray):
        import torch
        raw = np.ascontiguousarray(raw, dtype=np.float32)
        if self._pin is None or self._pin.shape[0] < raw.shape[0]:
            self._pin = torch.empty((max(raw.shape[0], self.raw_capacity), 3), dtype=torch.float32).pin_memory()
        self._pin[:raw.shape[0]].copy_(torch.from_numpy(raw))
        return self._pin.data_ptr(), raw.shape[0]

    def first_scan(self, sm, raw: np.ndarray):
        """Mapper.cpp:109-112: pre-process and insert at Identity (carving is a no-op on the empty map)."""
        icp = self.mapper.scan2MapReg_
        raw_c = self.eng.cloud(np.ascontiguousarray(raw, dtype=np.float32))
        ps = icp.processForScanMatchingAndMerging(raw_c)
        sm.insertScan(raw_c, ps.merge_, np.eye(4))
        sm.setPose(np.eye(4))
        raw_c.free()
        return ps.merge_

    def step(self, sm, raw: np.ndarray, odometryMotion: np.ndarray):
        self._activate(sm)
        ptr, n = self._pinned(raw)
        res = self.mapper.addRangeMeasurementHost(ptr, n, odometryMotion)
        p = self.params
        accepted = p.isIgnoreMinRefinementFitness or not (res.fitness_ < p.minRefinementFitness)
        return res, bool(accepted)     # minMovementBetweenMappingSteps = 0 in every preset: accepted scans are inserted

    def last_merge_cloud(self):
        # ring of pre-allocated clouds (no cudaMalloc per scan); longer than SubmapCollection's overlap buffer, so a buffered cloud is
        # never overwritten while it can still be replayed into a new submap
        if not hasattr(self, "_merge_ring"):
            self._merge_ring = [E.Cloud(self.eng) for _ in range(16)]
            self._merge_pos = 0
        c = self._merge_ring[self._merge_pos % len(self._merge_ring)]
        self._merge_pos += 1
        return self.mapper.lastProcessedScan(merge_into=c).merge_

    def insert_scan(self, sm, cloud, T):
        sm.insertScan(None, cloud, T, isPerformCarving=False)

    def set_pose(self, sm, T):
        sm.setPose(T)

    def map_cloud(self, sm):
        return sm.getMapPointCloud()

    def map_center(self, sm) -> np.ndarray:
        xyz, _ = sm.getMapPointCloud()
        return xyz.mean(axis=0) if len(xyz) else np.zeros(3)   # [O3D] GetCenter

    def build_voxel_map(self, sm) -> None:
        v = VOXEL_EXPANSION_ADJACENCY_REVISITING * self.params.mapBuilder.mapVoxelSize
        vm = self._voxel_maps.get(id(sm))
        if vm is None:
            vm = E.VoxelMap(self.eng, v, 1 << 18)
            self._voxel_maps[id(sm)] = vm
        vm.clear()
        c = sm.toCloud()
        vm.insertCloud(VOXEL_MAP_LAYER, c)
        c.free()

    def revisit_fitness(self, sm, scan, mapToRangeSensor) -> float:
        vm = self._voxel_maps[id(sm)]
        n = len(scan)
        if n == 0:
            return 0.0
        _flags, hits = vm.hasVoxelContainingPoint(scan, mapToRangeSensor)
        return hits / n

    # -- loop-closure refinement
    def submap_as_cloud(self, sm):
        return sm.toCloud()

    def overlap(self, source, target, T0, voxel, min_pts):
        return E.computeOverlappingClouds(self.eng, source, target, T0, voxel, min_pts)

    def register_batch(self, sources, targets, inits, max_corr, max_iter):
        pc = E.CloudRegistrationParameters(icp=E.IcpParameters(maxNumIter=max_iter, maxCorrespondenceDistance=max_corr, knn=self.params.icp.knn,
                                                              maxDistanceKnn=self.params.icp.maxDistanceKnn))
        reg = E.RegistrationIcpPointToPlane(self.eng, pc)
        out = reg.registerCloudsBatch(sources, targets, inits)
        self.eng.set_parameters(self.params)   # the scan-to-map chain keeps its own ICP parameters
        return out

    def information_matrix(self, source, target, max_corr, T):
        return E.getInformationMatrixFromPointClouds(self.eng, source, target, max_corr, T)

    def cloud_size(self, c) -> int:
        return len(c)

    def dense_map(self, sm):
        return sm.getDenseMap()

    def counters(self, sm) -> dict:
        return sm.mapperCounters()

    def close(self):
        for vm in self._voxel_maps.values():
            vm.free()
        self._voxel_maps.clear()
