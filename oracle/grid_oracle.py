"""CPU restatement (numpy, explicit fp32) of GridMM's grid build -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module, and only as the checker or the timed CPU baseline.
The product path (gridmm_b200/) never imports it.

Parity status: PINNED IN THE AUTHORING CONTAINER against the reference's own
`EnvBatch.getGlobalMap` executed in-process (oracle/_refshim.py, numpy 2.3.5),
0 mismatching cells; the resulting vectors are committed under tests/golden/
(oracle/make_golden.py is the generating script).  The reference itself ships no
tests or golden vectors, so there is nothing else to pin against.

Follows, line by line:
  get_rel_position            map_nav_src/r2r/env.py:115-121
  EnvBatch.getGlobalMap       map_nav_src/r2r/env.py:267-374
  EnvBatch.get_gridmap_pos_fts map_nav_src/r2r/env.py:242-265
  calculate_vp_rel_pos_fts    map_nav_src/r2r/env.py:60-77
  get_angle_fts               map_nav_src/r2r/env.py:52-58
  target_patch_id (pretraining dataset only; its grid arithmetic is otherwise the R2R one)
                              pretrain_src/data/dataset.py:361-368, 427-439
Arithmetic contract: every op is an individually rounded IEEE fp32 op (numpy
elementwise semantics, no FMA); trig of view angles / heading is evaluated in
double on the host and rounded to fp32 (python float x fp32 array => fp32 under
NEP 50); int conversion truncates toward zero.

The grid width (14 in the reference, env.py:43-44) and the feature dim (768,
env.py:299) are parameters here so BASELINE config 1 (8x8 grid, 512-d) can run;
at 14/768 the functions are checked against the reference.
"""
import math

import numpy as np

f32 = np.float32

MAX_DIST = 30.0  # env.py:47


class R2RGeometry:
    """Camera/rotation conventions of the discrete-env trees (r2r/reverie/rxr/soon, pretrain)."""
    depth_scale = 4000.0                      # env.py:116 (uint16 0.25mm units -> metres)
    tan_half_fov = math.tan(math.pi / 6)      # env.py:118
    flip_y = False                            # global_y = rel_y + pos.y (env.py:292)
    angle_offset = 0.0                        # angle = -heading (env.py:337)
    negate_map_x = False
    ce_pos_fts = False
    max_dist = 30.0                           # env.py:47

    @staticmethod
    def view_angle(v, heading):
        return v * math.pi / 6                # env.py:290  (ix-12)*pi/6


class CEGeometry:
    """Continuous-env variant, VLN_CE/vlnce_baselines/models/Policy_ViewSelection_GridMap.py:632-641, 689-825."""
    depth_scale = 1.0                         # depth already metres (:633)
    tan_half_fov = math.tan(math.pi / 4)      # R2R-CE hfov 90 (:635)
    flip_y = True                             # global_y = -rel_y + pos.y (:735)
    angle_offset = math.pi                    # angle = -heading + pi (:781)
    negate_map_x = True                       # map_x = -(...) (:790)
    ce_pos_fts = True                         # calculate_vp_rel_pos_fts reads (x, z, y): VLN_CE/.../models/utils.py:125-144
    max_dist = 25.0                           # MAX_DIST (:39; 40 for RxR-CE :282-285)

    @staticmethod
    def view_angle(v, heading):
        return v * math.pi / 6 - heading      # :731


class RxRCEGeometry(CEGeometry):
    """RxR-CE: the same conventions with the 79-degree camera (Policy_ViewSelection_GridMap.py:637-638) and MAX_DIST 40 (:282-285)."""
    tan_half_fov = math.tan(math.pi * 79. / 360.)
    max_dist = 40.0


_OFF7 = [-6 / 7, -4 / 7, -2 / 7, 0., 2 / 7, 4 / 7, 6 / 7]


def rel_position(depth_row, angle, geom=R2RGeometry):
    """env.py:115-121.  depth_row: [1,49] (uint16 or f32); angle: python float."""
    if geom.depth_scale != 1.0:
        depth_y = depth_row.astype(f32) / f32(geom.depth_scale)
    else:
        depth_y = depth_row.astype(f32)
    off = np.array(_OFF7 * 7, f32) * f32(geom.tan_half_fov)
    depth_x = depth_y * off
    ca, sa = f32(math.cos(angle)), f32(math.sin(angle))
    rel_x = depth_x * ca + depth_y * sa
    rel_y = depth_y * ca - depth_x * sa
    return rel_x, rel_y


class GridState:
    """Per-episode accumulated state (env.py:142-151; reset by newEpisodes :183-193)."""

    def __init__(self):
        self.wx = []        # list of f32[588]
        self.wy = []
        self.mask = []      # list of bool[588]
        self.fts = []       # list of f16[588,D]
        self.max_x, self.min_x = f32(-10000), f32(10000)
        self.max_y, self.min_y = f32(-10000), f32(10000)


def grid_step(state, depth_sub, clip, pos_xy, heading, grid_w=14, geom=R2RGeometry, legacy_half=False):
    """One getGlobalMap call (env.py:267-374).

    legacy_half: evaluate the window half-length as the reference's PINNED numpy (1.20.3, value-based casting) would:
    `position - np.float32` is float64 there, so half_len is computed in double and only rounded to fp32 when it meets the
    fp32 point arrays.  Not the oracle of record (that is the reference run under this container's numpy 2, all fp32); kept to
    COUNT how many cell ids the two conventions can disagree on (hard parts).

    depth_sub: [12,49] uint16 (R2R) or f32 metres (CE); clip: f16[12,50,D] (CLS first) or None;
    pos_xy: python floats; heading: python float.
    Returns (grid_fts f16[N,D] or None, cell i32[N] in {-1,0..grid_w^2-1}, half_len f32).
    """
    px, py = f32(pos_xy[0]), f32(pos_xy[1])
    xs, ys = [], []
    for v in range(12):                                             # env.py:289-294
        rel_x, rel_y = rel_position(depth_sub[v:v + 1], geom.view_angle(v, heading), geom)
        xs.append(rel_x + px)
        ys.append((-rel_y if geom.flip_y else rel_y) + py)
    wx = np.concatenate(xs, 0).reshape(-1)
    wy = np.concatenate(ys, 0).reshape(-1)
    state.wx.append(wx)
    state.wy.append(wy)
    state.mask.append((depth_sub != 0).reshape(-1))                 # env.py:283-285
    if clip is not None:
        state.fts.append(clip[:, 1:].reshape(-1, clip.shape[-1]))   # env.py:299-304
    # running bounds over ALL new points, masked ones included (env.py:312-319)
    if wx.max() > state.max_x: state.max_x = wx.max()
    if wx.min() < state.min_x: state.min_x = wx.min()
    if wy.max() > state.max_y: state.max_y = wy.max()
    if wy.min() < state.min_y: state.min_y = wy.min()
    # window (env.py:322-331), all fp32
    a, b = px - state.min_x, state.max_x - px
    x_half = a if a > b else b
    a, b = py - state.min_y, state.max_y - py
    y_half = a if a > b else b
    half = x_half if x_half > y_half else y_half
    half = f32(f32(half * f32(2)) / f32(3))                         # half_len * 2/3
    if legacy_half:
        pxd, pyd = float(pos_xy[0]), float(pos_xy[1])
        a, b = pxd - float(state.min_x), float(state.max_x) - pxd
        xh = a if a > b else b
        a, b = pyd - float(state.min_y), float(state.max_y) - pyd
        yh = a if a > b else b
        half = f32((xh if xh > yh else yh) * 2 / 3)
    # index assignment for every accumulated point (env.py:337-369)
    ang = -heading + geom.angle_offset
    c, s = f32(math.cos(ang)), f32(math.sin(ang))
    gx = np.concatenate(state.wx, 0)
    gy = np.concatenate(state.wy, 0)
    tx = gx - px
    ty = gy - py
    mx = tx * c + ty * s
    my = ty * c - tx * s
    if geom.negate_map_x:
        mx = -mx
    two_half = f32(f32(2) * half)
    ix = ((mx + half) / two_half * f32(grid_w - 1)).astype(np.int32)
    iy = ((my + half) / two_half * f32(grid_w - 1)).astype(np.int32)
    ix = np.clip(ix, 0, grid_w - 1)
    iy = np.clip(iy, 0, grid_w - 1)
    cell = ix * grid_w + iy
    cell = np.where(np.concatenate(state.mask, 0), cell, -1).astype(np.int32)
    fts = np.concatenate(state.fts, 0) if clip is not None else None
    return fts, cell, half


def gridmap_pos_fts(half_len, grid_w=14, geom=R2RGeometry):
    """env.py:242-265 + :60-77 + :52-58 -> f32[grid_w^2, 5] = [sin h, cos h, sin e, cos e, dist/30].

    The reference evaluates this in numpy scalar arithmetic whose width follows
    `half_len`'s type (fp32 here); the result only feeds a Linear, so parity for this
    function is a 1e-6 tolerance, not bit-exactness.
    """
    half_len = f32(half_len)
    cell_len = half_len * f32(2) / f32(grid_w)
    hs, es, ds = [], [], []
    for i in range(grid_w):
        for j in range(grid_w):
            x = f32(i) * cell_len - half_len + cell_len / f32(2)
            y = f32(j) * cell_len - half_len + cell_len / f32(2)
            if geom.ce_pos_fts:
                # Policy_ViewSelection_GridMap.py:661-684 passes (x, y, 0) to a helper that reads (x, z, y): dz = y, dy = 0
                xy = max(np.sqrt(x * x), 1e-8)
                xyz = max(np.sqrt(x * x + y * y), 1e-8)
                hs.append(np.arcsin(x / xy))
                es.append(np.arcsin(y / xyz))
                ds.append(xyz / geom.max_dist)
                continue
            xy = max(np.sqrt(x * x + y * y), 1e-8)
            heading = np.arcsin(x / xy)
            if y < 0:
                heading = np.pi - heading
            hs.append(heading)
            es.append(np.arcsin(f32(0) / xy))
            ds.append(xy / geom.max_dist)
    hs = np.array(hs).astype(f32)
    es = np.array(es).astype(f32)
    ds = np.array(ds).astype(f32)
    return np.stack([np.sin(hs), np.cos(hs), np.sin(es), np.cos(es), ds], 1).astype(f32)


def target_patch_id(pos_xy, next_xy, heading, half, is_last, grid_w=14):
    """Cell (1-based; 0 = "stay") that holds the NEXT ground-truth viewpoint of a pretraining path -- an extra label the
    pretraining dataset derives from the same window (pretrain_src/data/dataset.py:361-368, 427-439); the model does not read it.

    pos_xy / next_xy: python floats (viewpoint_info); heading: python float; half: the fp32 half_len of this step's window.
    The offset and its rotation are python-float (double) arithmetic; adding the fp32 `half` makes the rest fp32 (NEP 50), and
    `//` is numpy's floor_divide on fp32 scalars.  Note the reference scales by 14 (GLOBAL_WIDTH) here, not by 13 as for points."""
    if is_last:
        return 0
    tx = float(next_xy[0]) - float(pos_xy[0])
    ty = float(next_xy[1]) - float(pos_xy[1])
    ang = -heading
    rx = tx * math.cos(ang) + ty * math.sin(ang)
    ry = ty * math.cos(ang) - tx * math.sin(ang)
    half = f32(half)
    two_half = f32(f32(2) * half)
    ix = int(np.floor_divide(f32(f32(f32(rx) + half) * f32(grid_w)), two_half))
    iy = int(np.floor_divide(f32(f32(f32(ry) + half) * f32(grid_w)), two_half))
    ix = min(max(ix, 0), grid_w - 1)
    iy = min(max(iy, 0), grid_w - 1)
    return 1 + ix * grid_w + iy
