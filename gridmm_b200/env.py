"""Device-resident grid-map state: the B200 counterpart of the grid half of the reference's `EnvBatch`
(map_nav_src/r2r/env.py:125-400 -- `newEpisodes` :177-193, `getGlobalMap` :267-374, `getStates` :377-400).

The reference keeps, per episode, Python lists of numpy arrays, re-concatenates the whole [N,768] feature map every
step, rebuilds `grid_map` with a 196-iteration compare loop on the CPU and re-uploads everything to the GPU
(map_nav_src/r2r/agent.py:168).  Here the state lives in HBM for the whole episode:

    slab      fp16 [t_cap, B, 12*50, D]   the CLIP patch tokens of every visited viewpoint, written ONCE (one H2D copy of
                                          the step's B viewpoints); the CLS token stays in place and is skipped by indexing
    wx, wy    f32  [B, cap]               world coordinates of every accumulated point (cap = t_cap * 588)
    valid     u8   [B, cap]               depth != 0
    bounds    f32  [B, 4]                 running max_x, min_x, max_y, min_y

and one kernel launch per step (gridmm_grid_update) appends the new viewpoint, re-assigns all points to cells
(bit-exact with the reference arithmetic) and sorts them by cell for the pooling kernel.
"""
import math

import numpy as np
import torch

from . import ops

PTS = 588            # 12 horizon views x 49 patch centres (env.py:279-289)
VIEW_TOKENS = 50     # CLS + 49 patches (env.py:100)
_OFF7 = [-6 / 7, -4 / 7, -2 / 7, 0., 2 / 7, 4 / 7, 6 / 7]
PATCH_CENTRES = np.array([9 + 18 * i for i in range(7)])          # env.py:279
PATCH_CENTRES_CE = np.array([19 + 36 * i for i in range(7)])      # Policy_ViewSelection_GridMap.py (256x256 depth)
_VIEW_ANGLES = np.arange(12, dtype=np.float64) * math.pi / 6        # (ix - 12) * pi / 6 of the 12 horizon views, env.py:290


class Geometry:
    """Camera / rotation conventions."""

    def __init__(self, name, depth_scale, tan_half_fov, flip_y, angle_offset, negate_map_x, view_minus_heading, depth_is_f32,
                 pos_mode=0, max_dist=30.0):
        self.name = name
        self.depth_scale = depth_scale
        self.tan_half_fov = tan_half_fov
        self.flip_y = flip_y
        self.angle_offset = angle_offset
        self.negate_map_x = negate_map_x
        self.view_minus_heading = view_minus_heading
        self.depth_is_f32 = depth_is_f32
        self.pos_mode = pos_mode          # 1: the CE code's cell-feature convention (see include/gridmm_b200.h)
        self.max_dist = max_dist          # MAX_DIST: 30 (r2r/env.py:47), 25 / 40 (Policy_ViewSelection_GridMap.py:39, 269-285)
        # env.py:118: np.array(o, np.float32) * math.tan(...)  ->  f32(o_k) * f32(tan) in fp32
        self.off7 = (np.array(_OFF7, np.float32) * np.float32(tan_half_fov)).astype(np.float32)


GEOMETRIES = {
    # discrete envs: map_nav_src/{r2r,reverie,rxr}/env.py, pretrain_src/data/dataset.py
    "r2r": Geometry("r2r", 4000.0, math.tan(math.pi / 6), False, 0.0, False, False, False),
    # VLN_CE/vlnce_baselines/models/Policy_ViewSelection_GridMap.py:632-641, 689-825 (R2R-CE, hfov 90)
    "r2r_ce": Geometry("r2r_ce", 1.0, math.tan(math.pi / 4), True, math.pi, True, True, True, pos_mode=1, max_dist=25.0),
    # RxR-CE, hfov 79
    "rxr_ce": Geometry("rxr_ce", 1.0, math.tan(79 * math.pi / 360), True, math.pi, True, True, True, pos_mode=1, max_dist=40.0),
}


def target_patch_id(pos_xy, next_xy, heading, half_len, grid_w=14):
    """Pretraining label: 1 + cell of the next ground-truth viewpoint in this step's window, 0 when the path ends here
    (pretrain_src/data/dataset.py:361-368, 427-439; returned by the dataset next to the grid, not read by the model).
    Host scalar arithmetic like the reference's: the offset to the next viewpoint and its rotation by -heading are python
    floats, the fp32 window half-length then pulls the sum, the scale by grid_w (14 here, not grid_w - 1 as for the points) and
    the floor division into fp32.  `half_len` is this step's GridBatch.half_len entry."""
    if next_xy is None:
        return 0
    dx, dy = float(next_xy[0]) - float(pos_xy[0]), float(next_xy[1]) - float(pos_xy[1])
    c, s = math.cos(-heading), math.sin(-heading)
    h = np.float32(half_len)
    span = np.float32(2) * h
    idx = []
    for r in (dx * c + dy * s, dy * c - dx * s):
        q = np.floor_divide((np.float32(r) + h) * np.float32(grid_w), span)
        idx.append(min(max(int(q), 0), grid_w - 1))
    return 1 + idx[0] * grid_w + idx[1]


class GridBatch:
    """Handle to the device-resident grid map of a batch of episodes after a step (what the model's 'navigation' mode
    consumes instead of the reference's grid_fts / grid_map lists)."""

    def __init__(self, builder):
        b = builder
        self.batch = b.batch
        self.feat_dim = b.feat_dim
        self.grid_w = b.grid_w
        self.n_cells = b.grid_w * b.grid_w
        self.slab, self.slots, self.t_cap = b.slab, b.slots, b.t_cap
        self.slot_rows, self.view_rows, self.tok_off = 12 * VIEW_TOKENS, VIEW_TOKENS, 1
        self.cap = b.cap
        self.perm, self.cell_start, self.cell_rank, self.n_nonempty = b.perm, b.cell_start, b.cell_rank, b.n_nonempty
        self.cell, self.pos_fts, self.half_len, self.n_pts = b.cell, b.pos_fts, b.half_len, b.n_pts
        self.n_steps = b.n_steps.copy()
        self._builder = b
        self.pending = False      # True: gridmm_grid_update of this step has not been launched yet (step(lazy=True))

    def launch_update(self):
        """Launch the deferred gridmm_grid_update of a `step(lazy=True)` on the current stream (the model does this as the first
        kernel of its forward, so that the launch is part of the captured CUDA graph of the step); no-op otherwise."""
        if self.pending:
            self.pending = False
            self._builder._launch_update(None)

    def update_signature(self):
        """What a CUDA graph that contains this step's grid update depends on (static addresses / sizes)."""
        b = self._builder
        return (id(b), b.cap, b._d_pack.data_ptr(), b.wx.data_ptr(), b.cell.data_ptr(), b.bounds.data_ptr())

    # ---- views in the reference's formats (tests / drop-in consumers; these DO copy) ----
    def grid_map_numpy(self):
        """list of float64[N] with values in {-1, 0..n_cells-1} -- env.py:300-306,366-369."""
        self.launch_update()
        n = self.n_pts.cpu().numpy()
        cell = self.cell.cpu().numpy()
        return [cell[i, :n[i]].astype(np.float64) for i in range(self.batch)]

    def grid_fts_torch(self):
        """list of fp16[N, D] device tensors in point order (env.py:299-304)."""
        self.launch_update()
        out = []
        slab = self.slab.view(-1, 12, VIEW_TOKENS, self.feat_dim)
        slots = self.slots.view(self.batch, self.t_cap).cpu().numpy()
        for i in range(self.batch):
            t = int(self.n_steps[i])
            rows = slab[torch.as_tensor(slots[i, :t].astype(np.int64), device=slab.device)]   # [t,12,50,D]
            out.append(rows[:, :, 1:, :].reshape(-1, self.feat_dim))
        return out


class DeviceFeatureDB:
    """Device-resident counterpart of the reference's `SemanticFeaturesDB` (map_nav_src/r2r/env.py:98-113: an HDF5 file cached in a
    host dict, from which every step's [12,50,768] tokens are re-uploaded inside the growing map, r2r/agent.py:168).  The CLIP patch
    tokens of the 12 horizon views of every viewpoint are kept in HBM, uploaded ONCE per viewpoint -- the whole Matterport feature
    DB is ~10 GB, a B200 has 180 -- and a navigation step only names slots: GridMapBuilder(feature_db=db).step(depth, None, pos,
    heading, keys=[...]) moves no features at all (the pooling kernel gathers rows straight out of this buffer through its TMA
    gather4 tensor map)."""

    def __init__(self, capacity, feat_dim=768, device="cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("gridmm_b200.DeviceFeatureDB needs a CUDA device (there is no CPU path)")
        self.capacity, self.feat_dim, self.device = int(capacity), int(feat_dim), torch.device(device)
        self.buffer = torch.empty(self.capacity, 12 * VIEW_TOKENS, self.feat_dim, dtype=torch.float16, device=self.device)
        self.index = {}

    def __contains__(self, key):
        return key in self.index

    def __len__(self):
        return len(self.index)

    def put(self, key, fts):
        """fts: [12,50,D] fp16 tokens of the 12 horizon views (CLS first), or the reference's [36,50,D] (views 12..23 are taken,
        r2r/env.py:296); numpy / host tensor / device tensor.  Returns the slot; a key already present keeps its slot."""
        slot = self.index.get(key)
        if slot is not None:
            return slot
        if len(self.index) >= self.capacity:
            raise RuntimeError("DeviceFeatureDB is full (%d viewpoints)" % self.capacity)
        t = torch.as_tensor(fts)
        if t.shape[0] == 36:
            t = t[12:24]
        slot = len(self.index)
        self.buffer[slot].copy_(t.reshape(12 * VIEW_TOKENS, self.feat_dim).to(torch.float16), non_blocking=True)
        self.index[key] = slot
        return slot

    def slots(self, keys):
        try:
            return np.fromiter((self.index[k] for k in keys), dtype=np.int32, count=len(keys))
        except KeyError as e:
            raise KeyError("viewpoint %r is not in the device feature DB: put() it first" % (e.args[0],))


class GridMapBuilder:
    """`EnvBatch`-side replacement: persistent device buffers + one launch per step for the whole batch."""

    def __init__(self, batch_size, feat_dim=768, grid_w=14, geometry="r2r", max_steps=16, device="cuda", feature_db=None):
        if not torch.cuda.is_available():
            raise RuntimeError("gridmm_b200.GridMapBuilder needs a CUDA device (there is no CPU path)")
        self.feature_db = feature_db           # DeviceFeatureDB: the feature slab IS the DB, steps only name slots
        self.batch = int(batch_size)
        self.feat_dim = int(feat_dim)
        self.grid_w = int(grid_w)
        self.geom = GEOMETRIES[geometry] if isinstance(geometry, str) else geometry
        self.device = torch.device(device)
        self.t_cap = 0
        self._alloc(int(max_steps))
        B, nc = self.batch, self.grid_w * self.grid_w
        dev = self.device
        self.bounds = torch.empty(B, 4, dtype=torch.float32, device=dev)
        self.n_pts = torch.zeros(B, dtype=torch.int32, device=dev)
        self.half_len = torch.zeros(B, dtype=torch.float32, device=dev)
        self.cell_start = torch.zeros(B, nc + 1, dtype=torch.int32, device=dev)
        self.cell_rank = torch.zeros(B, nc, dtype=torch.int32, device=dev)
        self.n_nonempty = torch.zeros(B, dtype=torch.int32, device=dev)
        self.pos_fts = torch.zeros(B, nc, 5, dtype=torch.float32, device=dev)
        # per-step staging: ONE packed buffer [pose B x 4 f32 | view_cs B x 24 f32 | depth B x 588 (u16 or f32)] per step, written
        # by the host into a ring of pinned buffers (so the host never waits for the previous step's copy) and moved with one
        # H2D copy into a static device pack that gridmm_grid_update reads
        self._depth_elt = 4 if self.geom.depth_is_f32 else 2
        self._off_view = B * 4 * 4
        self._off_depth = self._off_view + B * 24 * 4
        self._off_slot = self._off_depth + B * PTS * self._depth_elt
        self._pack_bytes = self._off_slot + B * 4
        self._h_packs = [torch.empty(self._pack_bytes, dtype=torch.uint8).pin_memory() for _ in range(3)]
        self._h_evts = [None, None, None]
        self._h_next = 0
        self._d_pack = torch.empty(self._pack_bytes, dtype=torch.uint8, device=dev)
        self.d_pose = self._d_pack[:self._off_view].view(torch.float32).view(B, 4)
        self.d_view = self._d_pack[self._off_view:self._off_depth].view(torch.float32).view(B, 24)
        self.d_depth = self._d_pack[self._off_depth:self._off_slot].view(torch.float32 if self.geom.depth_is_f32 else torch.int16).view(B, PTS)
        self.d_new_slot = self._d_pack[self._off_slot:].view(torch.int32)
        self.h_clip = torch.empty(B, 12, VIEW_TOKENS, self.feat_dim, dtype=torch.float16).pin_memory()
        self._copy_stream = None
        self._staged_evt = None
        self._staged_step = -1
        self._clip_evt = None         # last asynchronous copy out of h_clip
        self._last_grid = None
        self.new_episodes()

    # ------------------------------------------------------------------ buffers
    def _alloc(self, t_cap):
        B, D, dev = self.batch, self.feat_dim, self.device
        old = self.t_cap
        cap = t_cap * PTS
        if cap > 65535:
            raise ValueError("at most 111 viewpoints per episode (cell sort uses 16-bit cursors)")
        db = getattr(self, "feature_db", None)
        slab = db.buffer if db is not None else torch.empty(t_cap, B, 12 * VIEW_TOKENS, D, dtype=torch.float16, device=dev)
        wx = torch.zeros(B, cap, dtype=torch.float32, device=dev)
        wy = torch.zeros(B, cap, dtype=torch.float32, device=dev)
        valid = torch.zeros(B, cap, dtype=torch.uint8, device=dev)
        if old:
            if db is None:
                slab[:old].copy_(self.slab)
            wx[:, :old * PTS].copy_(self.wx); wy[:, :old * PTS].copy_(self.wy); valid[:, :old * PTS].copy_(self.valid)
        self.slab, self.wx, self.wy, self.valid = slab, wx, wy, valid
        self.cell = torch.full((B, cap), -1, dtype=torch.int16, device=dev)
        self.perm = torch.zeros(B, cap, dtype=torch.int32, device=dev)
        # slot of (episode b, step t) = t * B + b  -> the B viewpoints of one step are one contiguous H2D copy.  (With per-episode
        # `active` masks an episode's t-th viewpoint is the one of the t'-th call, t' >= t: step() then rewrites its row of the table.)
        slots = (torch.arange(t_cap, dtype=torch.int32)[None, :] * B + torch.arange(B, dtype=torch.int32)[:, None]).contiguous()
        if old and getattr(self, "_slots_host", None) is not None:
            slots[:, :old] = self._slots_host
        self._slots_host = slots
        new_slots = slots.to(dev)
        if old and db is not None:
            new_slots[:, :old].copy_(self.slots)          # feature-DB mode: the table is written on the device (gridmm_grid_update)
        self.slots = new_slots
        self.t_cap, self.cap = t_cap, cap

    def _grow(self):
        """Double the capacity, up to the 111 viewpoints (65 268 points) the 16-bit sort cursors allow."""
        if self.t_cap >= 111:
            raise ValueError("at most 111 viewpoints per episode (cell sort uses 16-bit cursors)")
        self._alloc(min(self.t_cap * 2, 111))

    def new_episodes(self):
        """env.py:183-193."""
        self.bounds.copy_(torch.tensor([-10000.0, 10000.0, -10000.0, 10000.0]).repeat(self.batch, 1))
        self.n_pts.zero_()
        self.n_steps = np.zeros(self.batch, dtype=np.int64)      # viewpoints accumulated per episode
        self.n_calls = 0                                         # step() calls since the reset = next free slab row block
        if getattr(self, "_lockstep", True) is False:
            self._alloc_slots_reset()
        self._lockstep = True

    def _alloc_slots_reset(self):
        B, t_cap = self.batch, self.t_cap
        self._slots_host = (torch.arange(t_cap, dtype=torch.int32)[None, :] * B + torch.arange(B, dtype=torch.int32)[:, None]).contiguous()
        self.slots.copy_(self._slots_host)

    # ------------------------------------------------------------------ one navigation step
    @staticmethod
    def subsample_depth(depth_full, ce=False):
        """[..,36|12,H,W] depth map -> [..,12,49] patch-centre samples (env.py:279-285)."""
        c = PATCH_CENTRES_CE if ce else PATCH_CENTRES
        d = depth_full
        if d.shape[-3] == 36:
            d = d[..., 12:24, :, :]
        return d[..., c[:, None], c[None, :]].reshape(d.shape[:-2] + (49,))

    def host_pose(self, pos_xy, heading, out=None):
        """pose / view trigonometry, evaluated in double and rounded to fp32 like the reference
        (python float x np.float32 array, env.py:119-120, 290, 337, 347-348).  Returns [B, 28] fp32: px, py, cos, sin of the map
        angle, then (cos, sin) of the 12 view angles; `out` = (pose [B,4], view [B,24]) fp32 arrays to fill instead."""
        B, g = self.batch, self.geom
        pos_xy = np.asarray(pos_xy, dtype=np.float64).reshape(B, 2)
        heading = np.asarray(heading, dtype=np.float64).reshape(B)
        if out is None:
            full = np.empty((B, 28), dtype=np.float32)
            pose, view = full[:, :4], full[:, 4:]
        else:
            full, (pose, view) = None, out
        pose[:, 0:2] = pos_xy                      # float64 -> float32 rounding on assignment
        ang = -heading + g.angle_offset
        pose[:, 2] = np.cos(ang)
        pose[:, 3] = np.sin(ang)
        v = _VIEW_ANGLES
        va = v[None, :] - heading[:, None] if g.view_minus_heading else np.broadcast_to(v[None, :], (B, 12))
        view[:, 0::2] = np.cos(va)
        view[:, 1::2] = np.sin(va)
        return full

    def stage_features(self, clip, after=None):
        """Start the host->device copy of the NEXT viewpoint's CLIP tokens on this builder's copy stream and return at once.
        `step(..., clip=None)` then only waits for that copy on the device.  With two environment batches per GPU (two
        builders, one model) the 29.5 MB copy of one batch hides behind the other batch's kernels (bench.py's `e2e`).
        clip: pinned host tensor [B,12,50,D] fp16 (a pageable array is first copied into this builder's pinned buffer);
        after: optional CUDA event the copy must wait for (only needed when a slab slot is rewritten while an earlier step
        may still read it -- never the case in a real episode, where every step writes a new slot)."""
        grew = False
        if self.n_calls + 1 > self.t_cap:
            self._flush_pending()
            self._grow()
            grew = True
        t = self.n_calls
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            grew = True
        if not (isinstance(clip, torch.Tensor) and clip.is_pinned()):
            self._clip_free()
            self.h_clip.copy_(torch.as_tensor(clip).reshape(self.h_clip.shape))
            clip = self.h_clip
        cs = self._copy_stream
        if grew:      # the slab was (re)allocated on the compute stream: order that before the first copy into it
            cs.wait_stream(torch.cuda.current_stream(self.device))
        if after is not None:
            cs.wait_event(after)
        with torch.cuda.stream(cs):
            self.slab[t].copy_(clip.reshape(self.batch, 12 * VIEW_TOKENS, self.feat_dim), non_blocking=True)
            self._staged_evt = torch.cuda.Event()
            self._staged_evt.record(cs)
        if clip is self.h_clip:
            self._clip_evt = self._staged_evt
        self._staged_step = t

    def _clip_free(self):
        """h_clip (the pinned bounce buffer for pageable feature arrays) is reused every step: the asynchronous copy that read it
        last must have run before the host rewrites it."""
        if self._clip_evt is not None:
            self._clip_evt.synchronize()
            self._clip_evt = None

    def _flush_pending(self):
        """A step(lazy=True) whose grid update nobody launched yet must run before its inputs are overwritten."""
        g = self._last_grid
        if g is not None and g.pending:
            g.launch_update()

    def _launch_update(self, d_active):
        g = self.geom
        db = self.feature_db is not None
        ops.grid_update(self.batch, self.d_depth, g.depth_is_f32, g.depth_scale, self.d_pose, self.d_view, d_active, g.off7, g.flip_y,
                        g.negate_map_x, g.pos_mode, g.max_dist, self.grid_w, self.cap, self.wx, self.wy, self.valid, self.bounds, self.n_pts,
                        self.cell, self.half_len, self.perm, self.cell_start, self.cell_rank, self.n_nonempty, self.pos_fts,
                        new_slot=self.d_new_slot if db else None, slots=self.slots if db else None, t_cap=self.t_cap)

    def step(self, depth_sub, clip, pos_xy, heading, active=None, lazy=False, keys=None):
        """Append one viewpoint per episode and rebuild the grid assignment (getStates' grid half, env.py:392-398).

        depth_sub : [B,12,49] uint16 (0.25 mm) or float32 metres (CE); numpy (host) or a device tensor
        clip      : [B,12,50,D] fp16 CLIP tokens incl. CLS; numpy / host tensor (copied H2D) or a device tensor; None when the
                    copy was started earlier with stage_features()
        pos_xy    : [B,2] viewpoint x,y (python floats / float64);  heading : [B] radians
        keys      : feature-DB mode (GridMapBuilder(feature_db=...)): the B viewpoint keys (or int slots) of this step; `clip` is
                    ignored -- no feature bytes move, the step only names where they already are in HBM
        active    : optional [B] bools; episodes with 0 receive no viewpoint in this call
        lazy      : stage the inputs now but leave the launch of gridmm_grid_update to the consumer of the returned GridBatch
                    (forward('navigation') launches it as its first kernel, inside its CUDA graph when graphs are enabled)
        Returns a GridBatch.
        """
        B, g = self.batch, self.geom
        self._flush_pending()
        if self.n_calls + 1 > self.t_cap:
            self._grow()
        t = self.n_calls                  # slab row block of this call (all B viewpoints, active or not, land in slab[t])
        d_active = None
        db = self.feature_db
        if db is not None:
            if keys is None:
                raise ValueError("GridMapBuilder(feature_db=...) steps by viewpoint key: pass keys=[...]")
            new_slots = np.asarray(keys, dtype=np.int32) if isinstance(keys, np.ndarray) or isinstance(keys[0], (int, np.integer)) \
                else db.slots(keys)
            if active is not None and not bool(np.all(active)):
                d_active = torch.from_numpy(np.asarray(active).astype(np.uint8).reshape(B)).to(self.device)
        elif active is not None and not bool(np.all(active)):
            # Episodes with active[b] == 0 receive no viewpoint (the kernel leaves their points / bounds alone and only re-assigns
            # cells to the window of the pose passed for them); the others append theirs as viewpoint n_steps[b], which lives in
            # slab row block t = this call's index.
            # (The reference itself never skips: it re-adds the last viewpoint of ended episodes, r2r/env.py:392-398.)
            act = np.asarray(active).astype(bool).reshape(B)
            self._lockstep = False
            idx = torch.from_numpy(np.nonzero(act)[0])
            self._slots_host[idx, torch.from_numpy(self.n_steps[act])] = (t * B + idx).to(torch.int32)
            self.slots.copy_(self._slots_host, non_blocking=False)
            d_active = torch.from_numpy(act.astype(np.uint8)).to(self.device)
        elif not self._lockstep:
            idx = torch.arange(B)
            self._slots_host[idx, torch.from_numpy(self.n_steps)] = (t * B + idx).to(torch.int32)
            self.slots.copy_(self._slots_host, non_blocking=False)
        # features: one contiguous copy into slab[t] (feature-DB mode: nothing to copy)
        if db is not None:
            pass
        elif clip is None:
            if self._staged_step != t or self._staged_evt is None:
                raise RuntimeError("step(clip=None) needs stage_features() for this step first")
            torch.cuda.current_stream(self.device).wait_event(self._staged_evt)
            self._staged_step = -1
        elif isinstance(clip, torch.Tensor) and (clip.is_cuda or clip.is_pinned()):
            self.slab[t].copy_(clip.reshape(B, 12 * VIEW_TOKENS, self.feat_dim), non_blocking=True)
        else:
            self._clip_free()
            self.h_clip.copy_(torch.as_tensor(clip).reshape(self.h_clip.shape))
            self.slab[t].copy_(self.h_clip.view(B, 12 * VIEW_TOKENS, self.feat_dim), non_blocking=True)
            self._clip_evt = torch.cuda.Event()
            self._clip_evt.record()
        # pose, view trigonometry and depth: one packed pinned buffer (ring of three), one H2D copy
        k = self._h_next
        self._h_next = (k + 1) % len(self._h_packs)
        if self._h_evts[k] is not None:
            self._h_evts[k].synchronize()             # three steps old: long done unless the host runs that far ahead
        hp = self._h_packs[k]
        hp_np = hp.numpy()
        pose = hp_np[:self._off_view].view(np.float32).reshape(B, 4)
        view = hp_np[self._off_view:self._off_depth].view(np.float32).reshape(B, 24)
        self.host_pose(pos_xy, heading, out=(pose, view))
        if db is not None:
            hp_np[self._off_slot:].view(np.int32)[:] = new_slots
        depth_on_device = isinstance(depth_sub, torch.Tensor) and depth_sub.is_cuda
        if depth_on_device:
            n_h2d = self._off_depth
            if db is not None:
                self.d_new_slot.copy_(torch.from_numpy(new_slots), non_blocking=False)
        else:
            n_h2d = self._pack_bytes
            dst = hp_np[self._off_depth:self._off_slot].view(np.float32 if g.depth_is_f32 else np.uint16).reshape(B, PTS)
            src = depth_sub.numpy() if isinstance(depth_sub, torch.Tensor) else np.asarray(depth_sub)
            dst[...] = src.reshape(B, PTS).view(np.uint16) if (not g.depth_is_f32 and src.dtype == np.int16) else src.reshape(B, PTS)
        self._d_pack[:n_h2d].copy_(hp[:n_h2d], non_blocking=True)
        evt = torch.cuda.Event()
        evt.record()
        self._h_evts[k] = evt
        if depth_on_device:
            dd = depth_sub.reshape(B, PTS)
            self.d_depth.copy_(dd.view(torch.int16) if (not g.depth_is_f32 and dd.dtype == torch.uint16) else dd, non_blocking=True)
        self.n_steps += 1 if d_active is None else np.asarray(active).astype(np.int64).reshape(B)
        self.n_calls += 1
        grid = GridBatch(self)
        if lazy and d_active is None:
            grid.pending = True
        else:
            self._launch_update(d_active)
        self._last_grid = grid
        return grid

    def run_trajectory(self, depth_sub, clip, pos_xy, heading):
        """Whole ground-truth paths at once, as the pretraining dataset builds them (`get_traj_pano_fts`,
        pretrain_src/data/dataset.py:482-507: reset, then one getGlobalMap per viewpoint): depth_sub [B,T,12,49],
        clip [B,T,12,50,D], pos_xy [B,T,2], heading [B,T].  Returns the GridBatch after the last viewpoint; all T*588 points
        per path are assigned to the last viewpoint's window, which is what the pretraining model consumes."""
        self.new_episodes()
        grid = None
        for t in range(int(np.asarray(pos_xy).shape[1])):
            grid = self.step(depth_sub[:, t], clip[:, t], np.asarray(pos_xy)[:, t], np.asarray(heading)[:, t])
        return grid
