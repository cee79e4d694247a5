"""TEST INFRASTRUCTURE (oracle): plain-Python restatement of the reference's track management.

reference: src/x/vio/track_manager.cpp:115-436 (TrackManager::manageTracks), :576-636 (checkBaseline), :36-113 (getters),
           src/x/vio/vio.cpp:372-434 (VIO::importMatches), src/x/vision/camera.cpp:27-160, src/x/vision/tiled_image.cpp:139-158,
           src/x/vision/feature.cpp:47-67.

Pinned against the reference's own track_manager.cpp compiled in place (oracle/_ref/libxref_tm.so, tests/test_track_manager.py).
`std::sort` is not stable and its order among equal keys decides which tracks get the free SLAM slots, so the sort is
delegated to libstdc++ itself (oracle/_ref/libxsort.so, a 10-line helper built by oracle/ref_build/build_ref.sh).
Only tests/ may import this module.
"""
import ctypes as C
import math
import sys
from pathlib import Path

import numpy as np

_EPS = sys.float_info.epsilon
_MIN = sys.float_info.min
_MAX = sys.float_info.max
_sortlib = None


def _std_sort_desc(lengths):
    """Permutation std::sort(begin, end, [](a, b) { return a.size() > b.size(); }) applies (track_manager.cpp:274-276)."""
    global _sortlib
    if _sortlib is None:
        _sortlib = C.CDLL(str(Path(__file__).resolve().parent / "_ref" / "libxsort.so"))
        _sortlib.xsort_desc_by_len.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]
    n = len(lengths)
    a = (C.c_int * n)(*lengths)
    p = (C.c_int * n)()
    _sortlib.xsort_desc_by_len(a, n, p)
    return list(p)


def nearly_equal(a, b):   # feature.cpp:51-67
    if a == b:
        return True
    aa, ab, diff = abs(a), abs(b), abs(a - b)
    if a == 0 or b == 0 or aa + ab < _MIN:
        return diff < _EPS * _MIN
    return diff / min(aa + ab, _MAX) < _EPS


class Feat:
    __slots__ = ("t", "x", "y", "xd", "yd")

    def __init__(self, t=0.0, x=0.0, y=0.0, xd=0.0, yd=0.0):
        self.t, self.x, self.y, self.xd, self.yd = t, x, y, xd, yd

    def same(self, o):    # Feature::operator== (feature.cpp:47-49)
        return nearly_equal(self.x, o.x) and nearly_equal(self.y, o.y)


def _qmul(a, b):          # (w, x, y, z), Eigen's quaternion product
    return (a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
            a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
            a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3],
            a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1])


def _qconj(a):
    return (a[0], -a[1], -a[2], -a[3])


def _qnorm(a):
    n = math.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3])
    return (a[0] / n, a[1] / n, a[2] / n, a[3] / n)


class TrackManagerOracle:
    def __init__(self, fx, fy, cx, cy, s, w, h, bx, by, n_tiles_h, n_tiles_w):
        # Camera::Camera (camera.cpp:27-48)
        self.fx, self.fy, self.cx, self.cy, self.s = w * fx, h * fy, w * cx, h * cy, s
        self.inv_fx, self.inv_fy = 1.0 / self.fx, 1.0 / self.fy
        self.cx_n, self.cy_n = self.cx * self.inv_fx, self.cy * self.inv_fy
        self.s_term = 1.0 / (2.0 * math.tan(s / 2.0)) if s != 0.0 else float("inf")
        self.w, self.h, self.bx, self.by = w, h, bx, by
        self.nth, self.ntw = n_tiles_h, n_tiles_w
        self.tile_h, self.tile_w = h / n_tiles_h, w / n_tiles_w   # tiled_image.cpp:47-50
        self.slam, self.new_slam, self.opp = [], [], []
        self.msckf_n, self.short_n, self.new_std_n, self.new_msckf_n, self.lost = [], [], [], [], []
        self.n_evicted = [0, 0]   # test instrumentation: tile-balancing evictions of (new, persistent) SLAM tracks

    # ---- camera / image helpers
    def undistort(self, f):   # camera.cpp:69-87, :162-168
        dx, dy = f.xd * self.inv_fx - self.cx_n, f.yd * self.inv_fy - self.cy_n
        r = math.sqrt(dx * dx + dy * dy)
        k = 1.0
        if r > 0.01:
            k = (r if self.s == 0.0 else math.tan(r * self.s) * self.s_term) / r
        f.x, f.y = k * dx * self.fx + self.cx, k * dy * self.fy + self.cy

    def normalize(self, trk, max_size):   # camera.cpp:103-137
        n = len(trk)
        n_out = min(max_size, n) if max_size else n
        return [Feat(f.t, f.x * self.inv_fx - self.cx_n, f.y * self.inv_fy - self.cy_n,
                     f.xd * self.inv_fx - self.cx_n, f.yd * self.inv_fy - self.cy_n) for f in trk[n - n_out:]]

    def tile_of(self, f):                  # tiled_image.cpp:139-158
        c, col = f.xd - self.tile_w - 0.5, 0
        while c > 0:
            col += 1
            c -= self.tile_w
        r, row = self.h - f.yd - 0.5, self.nth - 1
        while r > self.tile_h:
            row -= 1
            r -= self.tile_h
        return row * self.ntw + col

    def check_baseline(self, trk, rots):   # track_manager.cpp:576-636; rots: list of (ax, ay, az, aw)
        n_q, n_obs = len(rots), len(trk)
        assert n_obs > 1 and n_q >= n_obs
        i_last = n_q - 1
        i_first = i_last - n_obs + 1
        min_x = max_x = trk[-1].x
        min_y = max_y = trk[-1].y
        qn = _qnorm((rots[i_last][3], rots[i_last][0], rots[i_last][1], rots[i_last][2]))
        for i in range(i_first, i_last + 1):
            qi = _qnorm((rots[i][3], rots[i][0], rots[i][1], rots[i][2]))
            rel = _qmul(_qconj(qi), qn)
            f = trk[i - i_first]
            rn = _qmul(_qmul(_qconj(rel), (0.0, f.x, f.y, 1.0)), rel)
            x, y = rn[1] / rn[3], rn[2] / rn[3]
            if x < min_x:
                min_x = x
            elif x > max_x:
                max_x = x
            if y < min_y:
                min_y = y
            elif y > max_y:
                max_y = y
        return (max_x - min_x) > self.bx or (max_y - min_y) > self.by

    # ---- VIO::importMatches (vio.cpp:372-434) + TrackManager::manageTracks (track_manager.cpp:115-436)
    def manage_tracks(self, match_vector, rots, n_poses_max, n_slam_max, min_track_length):
        mv = np.asarray(match_vector, dtype=float).reshape(-1, 10)
        matches = []
        for v in mv:
            p, c = Feat(v[1], xd=v[2], yd=v[3]), Feat(v[4], xd=v[5], yd=v[6])
            self.undistort(p)
            self.undistort(c)
            matches.append((p, c))
        rots = [tuple(r) for r in np.asarray(rots, dtype=float).reshape(-1, 4)]
        self.slam += self.new_slam                         # :120-124
        self.new_slam = []
        n_bins = self.nth * self.ntw
        bin_idx = [[] for _ in range(n_bins)]              # indexes as they are now
        bin_per = [[] for _ in range(n_bins)]              # indexes of the persistent tracks at entry
        fullest = 0
        t, n_lost = 0, 0
        self.lost = []
        while t < len(self.slam):                          # :139-187
            hit = next((m for m, (p, _) in enumerate(matches) if self.slam[t][-1].same(p)), None)
            if hit is None:
                self.lost.append(t + n_lost)
                del self.slam[t]
                n_lost += 1
                continue
            b = self.tile_of(matches[hit][1])
            bin_idx[b].append(t)
            bin_per[b].append(t + n_lost)
            if len(bin_idx[b]) > len(bin_idx[fullest]):
                fullest = b
            self.slam[t].append(matches[hit][1])
            del matches[hit]
            t += 1
        self.msckf_n, self.short_n = [], []                # :193-232
        prev_opp, self.opp = self.opp, []
        for p, c in matches:
            hit = next((k for k, trk in enumerate(prev_opp) if trk[-1].same(p)), None)
            if hit is not None:
                prev_opp[hit].append(c)
                self.opp.append(prev_opp.pop(hit))
            else:
                self.opp.append([p, c])
        rots_short = rots[:-1]                             # :263-273 (single-agent flavour)
        for dead in prev_opp:
            if len(dead) < 2:
                continue
            nt = self.normalize(dead, len(rots_short))
            if len(rots_short) >= len(nt) and self.check_baseline(nt, rots_short):
                self.short_n.append(nt)
        perm = _std_sort_desc([len(trk) for trk in self.opp])   # :274-276
        self.opp = [self.opp[i] for i in perm]
        t = 0
        while t < len(self.opp):                           # :277-398
            b = self.tile_of(self.opp[t][-1])
            if len(self.opp[t]) > (min_track_length - 1) % (1 << 64):
                if len(self.slam) + len(self.new_slam) < n_slam_max:
                    self.new_slam.append(self.opp.pop(t))
                    bin_idx[b].append(len(self.slam) + len(self.new_slam) - 1)
                    if len(bin_idx[b]) > len(bin_idx[fullest]):
                        fullest = b
                elif len(bin_idx[fullest]) > len(bin_idx[b]) + 1:
                    victim = bin_idx[fullest][-1]
                    self.n_evicted[0 if victim >= len(self.slam) else 1] += 1
                    if victim >= len(self.slam):
                        del self.new_slam[victim - len(self.slam)]
                    else:
                        self.lost.append(bin_per[fullest].pop())
                        del self.slam[victim]
                    for lst in bin_idx:
                        for j in range(len(lst)):
                            if lst[j] > victim:
                                lst[j] -= 1
                    bin_idx[fullest].pop()
                    self.new_slam.append(self.opp.pop(t))
                    bin_idx[b].append(len(self.slam) + len(self.new_slam) - 1)
                    for i in range(n_bins):
                        if len(bin_idx[i]) > len(bin_idx[fullest]):
                            fullest = i
                elif len(self.opp[t]) > (n_poses_max - 1) % (1 << 64):
                    nt = self.normalize(self.opp[t], len(rots))
                    if self.check_baseline(nt, rots):
                        self.msckf_n.append(nt)
                    del self.opp[t]
                else:
                    t += 1
            else:
                t += 1
        self.new_std_n, self.new_msckf_n = [], []          # :403-432
        with_b, without = [], []
        for trk in self.new_slam:
            nt = self.normalize(trk, len(rots))
            if self.check_baseline(nt, rots):
                self.new_msckf_n.append(nt)
                with_b.append(trk)
            else:
                self.new_std_n.append(nt)
                without.append(trk)
        self.new_slam = with_b + without

    # ---- getters (:36-61, :99-101), CSR form
    def get_list(self, which, size_out=0):
        if which == 4:
            lst = [self.normalize(t, size_out) for t in self.slam]
        elif which == 5:
            lst = [self.normalize(t, len(self.opp)) for t in self.opp]
        else:
            lst = [self.msckf_n, self.short_n, self.new_std_n, self.new_msckf_n][which]
        off = np.cumsum([0] + [len(t) for t in lst]).astype(np.int32)
        xy = np.array([[f.x, f.y] for t in lst for f in t], dtype=float).reshape(-1, 2)
        return off, xy

    def remove_persistent(self, idx):          # :83-85
        del self.slam[idx]

    def remove_new_persistent(self, idxs):     # :87-97
        for i in reversed(list(idxs)):
            del self.new_slam[i]
