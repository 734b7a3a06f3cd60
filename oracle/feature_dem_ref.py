"""Restatement of FLVIS's FeatureDEM (region-balanced corner selection) -- TEST INFRASTRUCTURE (oracle).

Follows /root/reference/src/processing/feature_dem.cpp:
  ctor :12-54, calHarrisR :59-88, fillIntoRegion :92-121, redetect :124-213, detect :215-266.
The GFTT call inside is oracle.gftt_ref (or cv2 when use_cv2=True, the reference's real library).
std::sort's tie order is taken from libstdc++ itself through oracle/_build/liboracle_helpers.so.
Reference quirks kept on purpose: patch[5] = (x+1,y+1); IX/IY integer division; Y2 = IY*IX, XY = IX*IX;
the `||` spacing test; redetect truncates the candidate to an int cv::Point.
"""
import ctypes
import math
import os
import subprocess
import numpy as np

from . import gftt_ref

f32 = np.float32
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def helpers_lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle_helpers.so")
        subprocess.check_call(["make", "-s", "-C", _HERE])     # no-op when up to date; rebuilds a stale helper library
        _LIB = ctypes.CDLL(path)
    return _LIB


def std_sort_desc(scores):
    s = np.ascontiguousarray(scores, np.float32)
    perm = np.zeros(len(s), np.int32)
    if len(s):
        helpers_lib().oracle_std_sort_desc(s.ctypes.data_as(ctypes.c_void_p), len(s),
                                           perm.ctypes.data_as(ctypes.c_void_p))
    return perm


class FeatureDEM:
    def __init__(self, width, height, f_para):
        self.width, self.height = width, height
        self.regionWidth = int(math.floor(width / 4.0))
        self.regionHeight = int(math.floor(height / 4.0))
        self.boundary_dis = int(math.floor(f_para[2] / 2.0))
        self.max_region_feature_num = int(f_para[0])
        self.min_region_feature_num = int(f_para[1])
        self.gftt_num = int(f_para[3])
        self.gftt_ql = float(f_para[4])
        self.gftt_dis = int(f_para[5])

    @staticmethod
    def harris_r(img, pt):
        xx, yy = int(pt[0]), int(pt[1])
        p = lambda x, y: int(img[y, x])
        patch = [p(xx - 1, yy - 1), p(xx, yy - 1), p(xx + 1, yy - 1), p(xx - 1, yy), p(xx, yy), p(xx + 1, yy + 1),
                 p(xx - 1, yy + 1), p(xx, yy + 1), p(xx + 1, yy + 1)]
        trunc = lambda a: int(a / 3) if a >= 0 else -int((-a) / 3)     # C++ int division truncates toward 0
        IX = f32(trunc(patch[0] + patch[3] + patch[6] - (patch[2] + patch[5] + patch[8])))
        IY = f32(trunc(patch[0] + patch[1] + patch[2] - (patch[6] + patch[7] + patch[8])))
        X2 = f32(IX * IX); Y2 = f32(IY * IX); XY = f32(IX * IX)
        t = f32(X2 + Y2)
        return f32(f32(f32(X2 * Y2) - f32(XY * XY)) - f32(f32(f32(0.05) * t) * t))

    def _region_of(self, pt):
        x, y = f32(pt[0]), f32(pt[1])
        if x >= 3 and x < (self.width - 3) and y >= 3 and y < (self.height - 3):
            return int(f32(f32(4) * f32(math.floor(f32(y / f32(self.regionHeight))))) + f32(x / f32(self.regionWidth)))
        return -1

    def _fill(self, img, pts, with_score):
        region = [[] for _ in range(16)]
        for pt in pts:
            r = self._region_of(pt)
            if r < 0:
                continue
            region[r].append(((f32(pt[0]), f32(pt[1])), self.harris_r(img, pt) if with_score else f32(99999.0)))
        return region

    def _gftt(self, img, n, use_cv2):
        if use_cv2:
            import cv2
            c = cv2.goodFeaturesToTrack(img, n, self.gftt_ql, self.gftt_dis)
            return np.zeros((0, 2), f32) if c is None else c.reshape(-1, 2)
        return gftt_ref.good_features_to_track(img, n, self.gftt_ql, self.gftt_dis)

    def detect(self, img, use_cv2=False, features=None):
        if features is None:
            features = self._gftt(img, self.gftt_num * 2, use_cv2)
        regions = self._fill(img, features, True)
        out = []
        bd = f32(self.boundary_dis)
        for i in range(16):
            tmp = regions[i]
            perm = std_sort_desc([s for _, s in tmp])
            kept = []
            for j in perm:
                pt = tmp[j][0]
                ok = True
                for k in kept:
                    if abs(f32(pt[0] - k[0])) <= bd or abs(f32(pt[1] - k[1])) <= bd:
                        ok = False
                if ok:
                    kept.append(pt)
                    if len(kept) >= self.max_region_feature_num:
                        break
            out.extend(kept)
        return np.array(out, f32).reshape(-1, 2)

    def redetect(self, img, existed_pts, use_cv2=False, features=None):
        existed = [(f32(p[0]), f32(p[1])) for p in np.asarray(existed_pts, np.float64).reshape(-1, 2)]
        region_key = [[k for k, _ in r] for r in self._fill(img, existed, False)]
        if features is None:
            features = self._gftt(img, self.gftt_num, use_cv2)
        prepare = self._fill(img, features, True)
        new = []
        bd = f32(self.boundary_dis)
        for i in range(16):
            perm = std_sort_desc([s for _, s in prepare[i]])
            for j in perm:
                p = prepare[i][j][0]
                pt = (f32(int(np.rint(p[0]))), f32(int(np.rint(p[1]))))      # cv::Point pt = Point2f (saturate_cast<int>)
                ok = True
                for k in region_key[i]:
                    if abs(f32(pt[0] - k[0])) <= bd or abs(f32(pt[1] - k[1])) <= bd:
                        ok = False
                if ok:
                    region_key[i].append(pt)
                    new.append(pt)
                    if len(region_key[i]) >= self.max_region_feature_num:
                        break
        return np.array(new, f32).reshape(-1, 2)
