"""The reference's on-disk inputs of the hot path, read into the layouts GridMapBuilder / DeviceFeatureDB take.

  * `clip_p32.hdf5`   one dataset per `<scan>_<viewpoint>`: CLIP ViT patch tokens of the 12 horizon views, [12, >= 50, 768]
                      (token 0 = CLS), written gzip'd as "float" by preprocess/get_map_feature.py:171-187 and read back as
                      float16 [12, 50, 768] by SemanticFeaturesDB (map_nav_src/r2r/env.py:98-113);
  * `depth.hdf5`      one dataset per key: the 36 discretised views' depth, [36, 128, 128(, 1)], uint16 in 0.25 mm units
                      (DepthFeaturesDB, r2r/env.py:80-95; sub-sampled at the 7 x 7 patch centres, r2r/env.py:279-285);
  * `viewpoint_info.json`   `<scan>_<viewpoint>` -> {"x", "y", "z"} (preprocess/get_viewpoint_info.py:69; r2r/env.py:168, 286).

h5py is the reference's own dependency and is imported only when a file is opened (this image does not ship it: the tests
exercise the conversion with an in-memory store of the same shape).  Nothing here touches the GPU except `preload()`, which
moves every viewpoint's tokens into a DeviceFeatureDB once -- after that a navigation step moves no feature bytes at all.
"""
import json

import numpy as np

VIEW_TOKENS = 50          # CLS + 7 x 7 patches per view (r2r/env.py:100)
DEPTH_HW = 128            # r2r/env.py:82 (WIDTH * HEIGHT = 128 * 128)


def _open_hdf5(path):
    try:
        import h5py
    except ImportError as e:      # pragma: no cover - depends on the environment
        raise ImportError("reading %s needs h5py (the reference's own dependency, requirements.txt); "
                          "pass open_fn= to use another store" % path) from e
    return h5py.File(path, "r")


class FeatureFiles:
    """Reader of the reference's feature files.  `open_fn(path)` must return a mapping whose values support `[...]`
    (h5py.File by default); the mapping is opened lazily and kept open."""

    def __init__(self, clip_file, depth_file, viewpoint_info_file=None, feat_dim=768, open_fn=None):
        self.clip_file, self.depth_file, self.feat_dim = clip_file, depth_file, int(feat_dim)
        self._open = open_fn or _open_hdf5
        self._clip = self._depth = None
        self.viewpoint_info = None
        if viewpoint_info_file is not None:
            with open(viewpoint_info_file) as f:
                self.viewpoint_info = json.load(f)

    @staticmethod
    def key(scan, viewpoint):
        return "%s_%s" % (scan, viewpoint)

    def clip_tokens(self, scan, viewpoint):
        """float16 [12, 50, D]: SemanticFeaturesDB.get_image_feature (r2r/env.py:104-113).  A file that stores all 36 views keeps
        the horizon ring 12..23, the only views the grid uses (r2r/env.py:296-300)."""
        if self._clip is None:
            self._clip = self._open(self.clip_file)
        ft = np.asarray(self._clip[self.key(scan, viewpoint)][...])[:, :VIEW_TOKENS].astype(np.float16)
        if ft.shape[0] == 36:
            ft = ft[12:24]
        if ft.shape != (12, VIEW_TOKENS, self.feat_dim):
            raise ValueError("CLIP tokens of %s have shape %s, expected (12, %d, %d)" % (self.key(scan, viewpoint), ft.shape, VIEW_TOKENS, self.feat_dim))
        return np.ascontiguousarray(ft)

    def depth_map(self, scan, viewpoint):
        """uint16 [36, 128, 128] in 0.25 mm units: DepthFeaturesDB.get_image_feature (r2r/env.py:86-95)."""
        if self._depth is None:
            self._depth = self._open(self.depth_file)
        d = np.asarray(self._depth[self.key(scan, viewpoint)][...])
        if d.ndim == 4:                                   # [36, 128, 128, 1]
            d = d[..., 0]
        if d.ndim == 2:                                   # [36, 128 * 128] flattened
            d = d[:, :DEPTH_HW * DEPTH_HW].reshape(d.shape[0], DEPTH_HW, DEPTH_HW)
        d = d.astype(np.uint16)
        if d.shape != (36, DEPTH_HW, DEPTH_HW):
            raise ValueError("depth of %s has shape %s, expected (36, 128, 128)" % (self.key(scan, viewpoint), d.shape))
        return d

    def position(self, scan, viewpoint):
        """(x, y) of the viewpoint (r2r/env.py:286, 290-291)."""
        if self.viewpoint_info is None:
            raise ValueError("no viewpoint_info.json was given")
        p = self.viewpoint_info[self.key(scan, viewpoint)]
        return float(p["x"]), float(p["y"])

    def step_inputs(self, scan_vps, with_clip=True):
        """Inputs of GridMapBuilder.step for one viewpoint per episode: depth_sub uint16 [B, 12, 49] (the 7 x 7 patch-centre
        pixels of the horizon views), clip float16 [B, 12, 50, D] (None with with_clip=False, the DeviceFeatureDB mode) and
        pos_xy float64 [B, 2]."""
        from .env import GridMapBuilder
        depth = np.stack([GridMapBuilder.subsample_depth(self.depth_map(s, v)) for s, v in scan_vps], 0).astype(np.uint16)
        clip = np.stack([self.clip_tokens(s, v) for s, v in scan_vps], 0) if with_clip else None
        pos = np.array([self.position(s, v) for s, v in scan_vps], dtype=np.float64) if self.viewpoint_info is not None else None
        return depth, clip, pos

    def preload(self, db, scan_vps):
        """Upload the tokens of every listed viewpoint into a DeviceFeatureDB (once; keys already present are skipped).
        Returns the number of viewpoints uploaded."""
        n = 0
        for s, v in scan_vps:
            k = self.key(s, v)
            if k not in db:
                db.put(k, self.clip_tokens(s, v))
                n += 1
        return n
