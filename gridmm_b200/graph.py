"""Device-resident node embeddings of the episodes' topological maps (second half).

Mirrors what the reference agent does around GraphMap.update_node_embed / get_node_embed
(map_nav_src/models/graph_utils.py:114-125; map_nav_src/r2r/agent.py:306-320 update, :126-129 read-out) for a whole batch
of episodes: the sums and counts live in HBM, the host keeps the viewpoint -> slot dictionaries (the graph itself, its
shortest paths and position features stay with the reference's GraphMap: they are host-side control logic, out of scope).
"""
import numpy as np
import torch

from . import ops


def build_update_index(slot_maps, visited, cur_vpids, cand_vpids, ended, n_views, max_nodes):
    """Host half of one update step, pure Python (CPU-testable).  slot_maps / visited: per-episode dict vp -> slot / set of
    visited vps, both updated in place.  Returns (cur_slot i32 [B], cand_slot i32 [B, n_views]); cand_slot[b, j] is the node
    that pano token j feeds, -1 for tokens that are not unvisited candidates (agent.py:316-320) and for ended episodes."""
    B = len(cur_vpids)
    cur = np.full((B,), -1, dtype=np.int32)
    cand = np.full((B, n_views), -1, dtype=np.int32)
    for b in range(B):
        if ended is not None and ended[b]:
            continue
        sm, vis = slot_maps[b], visited[b]
        vp = cur_vpids[b]
        vis.add(vp)                                        # gmap.update_graph(ob) marks the current viewpoint visited first
        if vp not in sm:
            sm[vp] = len(sm)
        cur[b] = sm[vp]
        for j, cvp in enumerate(cand_vpids[b]):
            if cvp is None or cvp in vis:
                continue
            if cvp not in sm:
                sm[cvp] = len(sm)
            cand[b, j] = sm[cvp]
        if len(sm) > max_nodes:
            raise ValueError("episode %d has %d map nodes, capacity is %d" % (b, len(sm), max_nodes))
    return cur, cand


def build_gather_index(slot_maps, gmap_vpids):
    """slots i32 [B, G] for agent.py:126-129: None (the stop node) and padding -> -1."""
    G = max(len(v) for v in gmap_vpids)
    out = np.full((len(gmap_vpids), G), -1, dtype=np.int32)
    for b, vps in enumerate(gmap_vpids):
        for g, vp in enumerate(vps):
            if vp is not None:
                out[b, g] = slot_maps[b][vp]
    return out


class DeviceGraphMaps:
    """Node embedding sums / counts of B episodes in HBM.  One kernel launch per update, one per read-out."""

    def __init__(self, batch, dim=768, max_nodes=128, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("gridmm_b200 has no CPU path: DeviceGraphMaps needs a CUDA device")
        self.device = torch.device(device if device is not None else "cuda")
        self.batch, self.dim, self.max_nodes = batch, dim, max_nodes
        self.node_sum = torch.zeros(batch, max_nodes, dim, dtype=torch.float32, device=self.device)
        self.node_cnt = torch.zeros(batch, max_nodes, dtype=torch.float32, device=self.device)
        self.reset()

    def reset(self):
        """Start a new batch of episodes."""
        self.node_sum.zero_(); self.node_cnt.zero_()
        self.slot_maps = [dict() for _ in range(self.batch)]
        self.visited = [set() for _ in range(self.batch)]

    def update(self, pano_embeds, pano_masks, cur_vpids, cand_vpids, ended=None):
        """pano_embeds f32 [B, V, D] / pano_masks [B, V] as returned by forward('panorama'); cand_vpids[b][j] is the viewpoint
        that pano token j looks at (pano_inputs['cand_vpids'])."""
        B, V, D = pano_embeds.shape
        cur, cand = build_update_index(self.slot_maps, self.visited, cur_vpids, cand_vpids, ended, V, self.max_nodes)
        idx = torch.from_numpy(np.concatenate([cur, cand.reshape(-1)])).to(self.device, non_blocking=True)
        ops.gmap_update(pano_embeds.contiguous(), pano_masks.contiguous(), idx[:B], idx[B:].view(B, V), self.node_sum, self.node_cnt)

    def node_embeds(self, gmap_vpids, out=None):
        """gmap_img_embeds f32 [B, G, D] in the order of gmap_vpids (a zero row for None = the stop node, and for padding)."""
        slots = torch.from_numpy(build_gather_index(self.slot_maps, gmap_vpids)).to(self.device, non_blocking=True)
        B, G = slots.shape
        if out is None:
            out = torch.empty(B, G, self.dim, dtype=torch.float32, device=self.device)
        ops.gmap_gather(self.node_sum, self.node_cnt, slots, out)
        return out


def teacher_actions(obs, gmap_vpids, ended, shortest_distances, visited_masks=None, ignoreid=-100):
    """Imitation-learning targets of one step (map_nav_src/r2r/agent.py:207-237), as one masked arg-min per episode:
    an ended episode -> `ignoreid`; standing on the goal (last viewpoint of `gt_path`) -> 0 ([stop]); otherwise the map node j >= 1
    that is not visited and minimises  d(node_j, goal) + d(current, node_j)  over the scan's shortest-path table (the first such
    node on ties); no eligible node -> `ignoreid`.  obs[i] needs 'scan', 'viewpoint', 'gt_path'; gmap_vpids[i][0] is the stop slot.
    Returns int64 [B] (host array: the caller moves it where its loss lives)."""
    out = np.full(len(obs), ignoreid, dtype=np.int64)
    for i, ob in enumerate(obs):
        if ended[i]:
            continue
        cur, goal = ob["viewpoint"], ob["gt_path"][-1]
        if cur == goal:
            out[i] = 0
            continue
        table = shortest_distances[ob["scan"]]
        from_cur = table[cur]
        nodes = gmap_vpids[i]
        cost = np.full(len(nodes), np.inf)
        eligible = np.ones(len(nodes), dtype=bool)
        eligible[0] = False
        if visited_masks is not None:
            eligible &= ~np.asarray(visited_masks[i], dtype=bool)[:len(nodes)]
        idx = np.nonzero(eligible)[0]
        if idx.size:
            cost[idx] = [table[nodes[j]][goal] + from_cur[nodes[j]] for j in idx]
            j = int(np.argmin(cost))                      # first minimum, like the reference's strict `<`
            if np.isfinite(cost[j]):
                out[i] = j
    return out
