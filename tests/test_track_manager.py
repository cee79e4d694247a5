"""Track management: the product's xb_tm_* (libxb200.so, host code: no GPU needed) and the oracle
restatement against the reference's own track_manager.cpp compiled in place (oracle/_ref/libxref_tm.so).

Scenario: points drift through a 640 x 480 image over 60 frames, appear, disappear and are occasionally dropped by the
"tracker"; the camera rotates slowly.  Every frame all six lists (MSCKF, short MSCKF, new SLAM standard / MSCKF-SLAM,
normalised SLAM, opportunistic) and the lost-SLAM indexes must be IDENTICAL across the three implementations: integer
structure exactly, coordinates to 1e-15 (same arithmetic up to compiler contraction).  Between frames the "filter" rejects
some new persistent tracks and one persistent track, as VioUpdater / StateManager do through removeNewPersistentTracksAtIndexes
/ removePersistentTracksAtIndex."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from oracle.track_manager import TrackManagerOracle
from x_multi_agent_b200.track_manager import TrackManager

REF = Path(__file__).resolve().parents[1] / "oracle" / "_ref" / "libxref_tm.so"
need_ref = pytest.mark.skipif(not REF.exists(), reason="oracle/_ref/libxref_tm.so not built (oracle/ref_build/build_ref.sh)")

CAM = dict(fx=0.46, fy=0.61, cx=0.5, cy=0.5, s=0.95, w=640, h=480, bx=0.02, by=0.02, nth=3, ntw=4)


class RefTm:
    def __init__(self):
        self.l = C.CDLL(str(REF))
        self.l.xref_tm_create.restype = C.c_void_p
        self.l.xref_tm_create.argtypes = [C.c_double] * 5 + [C.c_uint, C.c_uint, C.c_double, C.c_double, C.c_uint, C.c_uint]
        self.l.xref_tm_manage.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int, C.c_uint, C.POINTER(C.c_double), C.c_int,
                                          C.c_int, C.c_int, C.c_int]
        self.l.xref_tm_list_size.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        self.l.xref_tm_get_list.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        self.l.xref_tm_lost.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int]
        self.l.xref_tm_remove_persistent.argtypes = [C.c_void_p, C.c_uint]
        self.l.xref_tm_remove_new_persistent.argtypes = [C.c_void_p, C.POINTER(C.c_uint), C.c_int]
        c = CAM
        self.h = self.l.xref_tm_create(c["fx"], c["fy"], c["cx"], c["cy"], c["s"], c["w"], c["h"], c["bx"], c["by"], c["nth"], c["ntw"])

    def manage_tracks(self, mv, rots, seq, n_poses_max, n_slam_max, min_len):
        mv = np.ascontiguousarray(mv, dtype=np.float64).reshape(-1, 10)
        rots = np.ascontiguousarray(rots, dtype=np.float64)
        self.l.xref_tm_manage(self.h, mv.ctypes.data_as(C.POINTER(C.c_double)), len(mv), seq,
                              rots.ctypes.data_as(C.POINTER(C.c_double)), len(rots), n_poses_max, n_slam_max, min_len)

    def get_list(self, which, size_out=0):
        nt, no = C.c_int(0), C.c_int(0)
        self.l.xref_tm_list_size(self.h, which, size_out, C.byref(nt), C.byref(no))
        off = np.zeros(nt.value + 1, dtype=np.int32)
        xy = np.zeros((no.value, 2))
        self.l.xref_tm_get_list(self.h, which, size_out, off.ctypes.data_as(C.POINTER(C.c_int)), xy.ctypes.data_as(C.POINTER(C.c_double)))
        return off, xy

    def lost(self):
        out = np.zeros(512, dtype=np.int32)
        n = self.l.xref_tm_lost(self.h, out.ctypes.data_as(C.POINTER(C.c_int)), 512)
        return out[:n]

    def remove_persistent(self, idx):
        self.l.xref_tm_remove_persistent(self.h, idx)

    def remove_new_persistent(self, idxs):
        a = np.ascontiguousarray(idxs, dtype=np.uint32)
        self.l.xref_tm_remove_new_persistent(self.h, a.ctypes.data_as(C.POINTER(C.c_uint)), len(a))


def scenario(seed, frames, n_points):
    """Per frame: (match vector [n, 10], camera attitudes [n_rots, 4])."""
    rng = np.random.default_rng(seed)
    pts = rng.uniform([20, 20], [620, 460], size=(n_points, 2))
    vel = rng.normal(0, 2.5, size=(n_points, 2)) + np.array([3.0, 1.0])
    alive = np.zeros(n_points, dtype=bool)
    alive[: n_points // 2] = True
    out = []
    rots = []
    ang = 0.0
    for k in range(frames):
        ang += 0.01 + 0.005 * np.sin(0.3 * k)
        q = np.array([np.sin(ang / 2) * 0.6, np.sin(ang / 2) * 0.8, 0.0, np.cos(ang / 2)])
        rots.append(q)
        new = pts + vel + rng.normal(0, 0.05, size=pts.shape)
        inside = (new[:, 0] > 5) & (new[:, 0] < 635) & (new[:, 1] > 5) & (new[:, 1] < 475)
        tracked = alive & inside & (rng.random(n_points) > 0.04)
        idx = np.flatnonzero(tracked)
        rng.shuffle(idx)     # the front end orders matches by FAST score, not by track
        mv = np.zeros((len(idx), 10))
        mv[:, 1] = 0.1 * k
        mv[:, 2:4] = pts[idx]
        mv[:, 4] = 0.1 * (k + 1)
        mv[:, 5:7] = new[idx]
        out.append((mv, np.array(rots)))
        # points that left or were dropped respawn somewhere else as new detections one frame later
        gone = ~tracked
        new[gone] = rng.uniform([20, 20], [620, 460], size=(gone.sum(), 2))
        vel[gone] = rng.normal(0, 2.5, size=(gone.sum(), 2)) + np.array([3.0, 1.0])
        alive = np.ones(n_points, dtype=bool) & (rng.random(n_points) > 0.02)
        pts = new
    return out


def _same(tag, a, b, tol=1e-15):
    assert np.array_equal(a[0], b[0]), f"{tag}: track structure differs {a[0]} vs {b[0]}"
    if len(a[1]):
        assert np.abs(a[1] - b[1]).max() <= tol, f"{tag}: coordinates differ by {np.abs(a[1] - b[1]).max():.2e}"


def _run(seed, n_poses_max, n_slam_max, min_len, frames=60, n_points=120, with_ref=True):
    c = CAM
    prod = TrackManager(c["fx"], c["fy"], c["cx"], c["cy"], c["s"], c["w"], c["h"], c["bx"], c["by"], c["nth"], c["ntw"])
    ora = TrackManagerOracle(c["fx"], c["fy"], c["cx"], c["cy"], c["s"], c["w"], c["h"], c["bx"], c["by"], c["nth"], c["ntw"])
    ref = RefTm() if with_ref else None
    rng = np.random.default_rng(seed + 1000)
    stats = np.zeros(7, dtype=int)
    for k, (mv, rots) in enumerate(scenario(seed, frames, n_points)):
        rots = rots[-(n_poses_max + 1):]    # window attitudes + the current one (vio_updater.cpp:150-153)
        if k == 0:
            mv = mv[:0]                     # first image: no matches are imported (vio.cpp:286-288)
        prod.manage_tracks(mv, rots, n_poses_max, n_slam_max, min_len)
        ora.manage_tracks(mv, rots, n_poses_max, n_slam_max, min_len)
        if ref:
            ref.manage_tracks(mv, rots, k + 1, n_poses_max, n_slam_max, min_len)
        for which in range(6):
            so = n_poses_max if which == 4 else 0
            lp = prod.get_list(which, so)
            _same(f"frame {k} list {which} product vs oracle", lp, ora.get_list(which, so))
            if ref:
                _same(f"frame {k} list {which} product vs reference", lp, ref.get_list(which, so))
            stats[which] += len(lp[0]) - 1
        lost = prod.lost_slam_idxs()
        assert list(lost) == list(ora.lost)
        if ref:
            assert list(lost) == list(ref.lost())
        stats[6] += len(lost)
        # the filter rejects some of the new persistent tracks (failed initialisation) and, now and then, a persistent one
        n_slam, n_new, _ = prod.counts()
        if n_new > 1 and rng.random() < 0.3:
            bad = sorted(rng.choice(n_new, size=min(2, n_new), replace=False).tolist())
            prod.remove_new_persistent_tracks(bad)
            ora.remove_new_persistent(bad)
            if ref:
                ref.remove_new_persistent(bad)
        if n_slam > 3 and rng.random() < 0.1:
            i = int(rng.integers(n_slam))
            prod.remove_persistent_track(i)
            ora.remove_persistent(i)
            if ref:
                ref.remove_persistent(i)
    prod.close()
    return np.concatenate([stats, ora.n_evicted])


@need_ref
@pytest.mark.parametrize("seed,n_poses_max,n_slam_max,min_len,expect_msckf",
                         [(0, 10, 15, 4, True), (1, 6, 8, 3, True), (2, 12, 40, 5, False), (3, 5, 6, 2, True)])
def test_track_manager_matches_the_compiled_reference(seed, n_poses_max, n_slam_max, min_len, expect_msckf):
    stats = _run(seed, n_poses_max, n_slam_max, min_len)
    # the scenario must exercise every list and the eviction / lost-track paths (with 40 SLAM slots every long track is
    # absorbed as a SLAM feature: that case exercises the slot filling, not the MSCKF list)
    assert stats[1] > 0 and stats[4] > 0 and stats[5] > 0 and stats[6] > 0 and (stats[2] + stats[3]) > 0, stats
    assert (stats[0] > 0) == expect_msckf, stats
    if seed == 1:
        assert stats[8] > 0, "the tile-balancing eviction of a persistent track was not exercised"


def test_track_manager_matches_the_oracle_without_the_reference_binary():
    """Same comparison product vs oracle only (runs wherever libxb200.so and libxsort.so are, e.g. on the GPU box)."""
    if not (REF.parent / "libxsort.so").exists():
        pytest.skip("oracle/_ref/libxsort.so not built")
    stats = _run(7, 8, 12, 3, with_ref=False)
    assert stats[0] > 0 and stats[4] > 0


def test_measurement_seam():
    """TrackManager.measurement() yields the Measurement the filter's set_measurement takes."""
    c = CAM
    tm = TrackManager(c["fx"], c["fy"], c["cx"], c["cy"], c["s"], c["w"], c["h"], c["bx"], c["by"], c["nth"], c["ntw"])
    for k, (mv, rots) in enumerate(scenario(5, 12, 60)):
        tm.manage_tracks(mv if k else mv[:0], rots[-7:], 6, 8, 3)
    m = tm.measurement(1.2, 6)
    assert len(m.slam_trks) == tm.counts()[0] and all(t.shape[1] == 2 and len(t) <= 6 for t in m.slam_trks)
    tm.close()


def test_cxx_track_manager_through_the_reference_api(tmp_path):
    """x::TrackManager::manageTracks (include/x/vio/track_manager.h, the reference's signature) == the C ABI on the same
    inputs; compiled with g++ against include/ and run on the host (no GPU: the track manager is host code)."""
    import os
    import subprocess
    root = Path(__file__).resolve().parents[1]
    exe = tmp_path / "test_tm"
    r = subprocess.run(["g++", "-std=c++17", "-O1", f"-I{root / 'include'}", f"-I{root / 'oracle' / 'ref_build' / 'shim'}",
                        os.fspath(root / "tests" / "cxx" / "test_track_manager.cpp"), "-o", os.fspath(exe),
                        f"-L{root / 'x_multi_agent_b200'}", "-lxb200", f"-Wl,-rpath,{root / 'x_multi_agent_b200'}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([os.fspath(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_delaunay_facet_lookup_matches_scipy():
    """TrackManager::featureTriangleAtPoint (track_manager.cpp:443-560) picks the Delaunay facet of the SLAM features'
    image positions that contains the LRF image point.  cv::Subdiv2D is not available here; the Delaunay triangulation of
    points in general position is unique, so qhull (scipy.spatial.Delaunay) on the same single-precision points is an
    independent oracle: same three vertices, or no facet on both sides when the point lies outside the hull."""
    import ctypes as C

    from scipy.spatial import Delaunay

    from x_multi_agent_b200 import lib as L
    lib = L.load()
    rng = np.random.default_rng(0)
    found = 0
    for trial in range(200):
        n = int(rng.integers(3, 200))
        xy = np.ascontiguousarray(rng.uniform([20, 20], [620, 460], (n, 2)).astype(np.float32).astype(np.float64))
        q = (320.5, 240.5) if trial % 2 else tuple(rng.uniform([60, 60], [580, 420]))
        ids = (C.c_int * 3)()
        r = lib.xb_tm_delaunay_facet(xy.ctypes.data_as(C.POINTER(C.c_double)), n, 640, 480, q[0], q[1], ids)
        tri = Delaunay(xy)
        s = tri.find_simplex(np.array([q], dtype=np.float32).astype(np.float64))[0]
        want = set(int(i) for i in tri.simplices[s]) if s >= 0 else None
        got = set(int(i) for i in ids) if r == 3 else None
        assert got == want, (trial, n, got, want)
        found += got is not None
    assert found > 150
