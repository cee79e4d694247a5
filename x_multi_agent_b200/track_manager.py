"""Host-side mirror of x::TrackManager (reference: include/x/vio/track_manager.h, src/x/vio/track_manager.cpp) over the
xb_tm_* entry points of libxb200.so: match vector in, the five track lists of VioUpdater::preProcess out
(src/x/vio/vio_updater.cpp:165-179)."""
import ctypes as C

import numpy as np

from . import lib as L
from .filter import Measurement

MSCKF, MSCKF_SHORT, NEW_SLAM_STD, NEW_SLAM_MSCKF, SLAM, OPP = range(6)


class TrackManager:
    def __init__(self, fx, fy, cx, cy, s, img_width, img_height, min_baseline_x_n, min_baseline_y_n, n_tiles_h=1,
                 n_tiles_w=1, multi_uav=False):
        self.lib = L.load()
        cfg = L.XbTmConfig(fx, fy, cx, cy, s, img_width, img_height, min_baseline_x_n, min_baseline_y_n, n_tiles_h,
                           n_tiles_w, int(multi_uav))
        self.h = self.lib.xb_tm_create(C.byref(cfg))
        if not self.h:
            raise ValueError("xb_tm_create: invalid configuration")

    def close(self):
        if self.h:
            self.lib.xb_tm_destroy(self.h)
            self.h = None

    def clear(self):
        self.lib.xb_tm_clear(self.h)

    def manage_tracks(self, match_vector, cam_rots, n_poses_max, n_slam_features_max, min_track_length):
        """VIO::importMatches + TrackManager::manageTracks.  match_vector: n x 10, cam_rots: n_rots x 4 (ax, ay, az, aw)."""
        mv = np.ascontiguousarray(match_vector, dtype=np.float64).reshape(-1, 10)
        cr = np.ascontiguousarray(cam_rots, dtype=np.float64).reshape(-1, 4)
        L.check(self.lib.xb_tm_manage_tracks(self.h, mv.ctypes.data_as(C.POINTER(C.c_double)), len(mv),
                                             cr.ctypes.data_as(C.POINTER(C.c_double)), len(cr), n_poses_max,
                                             n_slam_features_max, min_track_length))

    def get_list(self, which, size_out=0):
        """(offsets[n+1], xy[n_obs, 2]) of one of the lists (SLAM: normalizeSlamTracks(size_out))."""
        nt, no = C.c_int(0), C.c_int(0)
        L.check(self.lib.xb_tm_list_size(self.h, which, size_out, C.byref(nt), C.byref(no)))
        off = np.zeros(nt.value + 1, dtype=np.int32)
        xy = np.zeros((max(no.value, 0), 2), dtype=np.float64)
        L.check(self.lib.xb_tm_get_list(self.h, which, size_out, off.ctypes.data_as(C.POINTER(C.c_int)),
                                        xy.ctypes.data_as(C.POINTER(C.c_double)), None))
        return off, xy

    def lost_slam_idxs(self):
        n = self.lib.xb_tm_lost_slam_idxs(self.h, None, 0)
        out = np.zeros(max(n, 1), dtype=np.int32)
        self.lib.xb_tm_lost_slam_idxs(self.h, out.ctypes.data_as(C.POINTER(C.c_int)), n)
        return out[:n]

    def remove_persistent_track(self, idx):
        L.check(self.lib.xb_tm_remove_persistent_track(self.h, int(idx)))

    def remove_new_persistent_tracks(self, idxs):
        a = np.ascontiguousarray(idxs, dtype=np.uint32)
        L.check(self.lib.xb_tm_remove_new_persistent_tracks(self.h, a.ctypes.data_as(C.POINTER(C.c_uint)), len(a)))

    def counts(self):
        a, b, c = C.c_int(0), C.c_int(0), C.c_int(0)
        L.check(self.lib.xb_tm_counts(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def feature_triangle_at_point(self, x_dist, y_dist):
        """TrackManager::featureTriangleAtPoint (track_manager.cpp:443-560): ids of the three SLAM features whose Delaunay
        facet contains the image point, or []."""
        ids = (C.c_int * 3)()
        n = L.check(self.lib.xb_tm_feature_triangle_at_point(self.h, float(x_dist), float(y_dist), ids))
        return [int(i) for i in ids] if n == 3 else []

    def normalize_point(self, x_dist, y_dist):
        """Camera::undistort + Camera::normalize (camera.cpp:69-101) of one distorted pixel position."""
        out = (C.c_double * 2)()
        L.check(self.lib.xb_tm_normalize_point(self.h, float(x_dist), float(y_dist), out))
        return float(out[0]), float(out[1])

    def measurement(self, timestamp, n_poses_max):
        """The VioMeasurement-side seam: what VioUpdater::preProcess hands to the update (vio_updater.cpp:165-179)."""
        def tl(which, size_out=0):
            off, xy = self.get_list(which, size_out)
            return [xy[off[i]:off[i + 1]].copy() for i in range(len(off) - 1)]
        return Measurement(timestamp, tl(SLAM, n_poses_max), tl(MSCKF), tl(MSCKF_SHORT), tl(NEW_SLAM_STD), tl(NEW_SLAM_MSCKF),
                           [int(i) for i in self.lost_slam_idxs()])
