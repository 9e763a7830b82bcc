"""TEST INFRASTRUCTURE ONLY (never imported by gridmm_b200/): CPU restatement of the node-embedding bookkeeping of the
reference's topological map.

  * GraphMap.update_node_embed / get_node_embed            map_nav_src/models/graph_utils.py:114-125
  * the agent's per-step update of visited / unvisited nodes    map_nav_src/r2r/agent.py:306-320
  * the read-out into gmap_img_embeds ([stop] + nodes)          map_nav_src/r2r/agent.py:126-129, pad_tensors_wgrad

Pinned against the reference's own GraphMap class by tests/test_oracle_golden.py::test_graph_oracle_matches_reference_graphmap
(the class is imported from /root/reference when it is present).
"""
import torch


class NodeEmbeds:
    """Per-viewpoint running sums of one episode's map (graph_utils.py:114-125): `rewrite` replaces a node's sum by the new
    embedding with count 1 (the agent does that for the viewpoint it stands on), otherwise the embedding is added to the node's
    sum and its count grows by one (a viewpoint seen as a candidate from several places); the read-out is the mean."""

    def __init__(self):
        self.sums, self.counts = {}, {}

    @property
    def node_embeds(self):                       # same view as the reference's dict: vp -> [sum, count]
        return {vp: [self.sums[vp], self.counts[vp]] for vp in self.sums}

    def update_node_embed(self, vp, embed, rewrite=False):
        if rewrite or vp not in self.sums:
            self.sums[vp], self.counts[vp] = embed, 1
        else:
            self.sums[vp] = embed + self.sums[vp]
            self.counts[vp] += 1

    def get_node_embed(self, vp):
        return self.sums[vp] / self.counts[vp]


def step_update(maps, visited, pano_embeds, pano_masks, cur_vpids, cand_vpids, ended):
    """agent.py:306-320 for one step of a batch (visited[b]: set, updated like gmap.update_graph(ob) does before)."""
    m = pano_masks.to(pano_embeds.dtype)
    avg = torch.sum(pano_embeds * m.unsqueeze(2), 1) / torch.sum(m, 1, keepdim=True)
    for i, gm in enumerate(maps):
        if ended[i]:
            continue
        visited[i].add(cur_vpids[i])
        gm.update_node_embed(cur_vpids[i], avg[i], rewrite=True)
        for j, vp in enumerate(cand_vpids[i]):
            if vp not in visited[i]:
                gm.update_node_embed(vp, pano_embeds[i, j])


def read_out(maps, gmap_vpids):
    """agent.py:126-129 + pad_tensors_wgrad: zero row for the stop node, zero padding to the longest map."""
    rows = []
    for gm, vps in zip(maps, gmap_vpids):
        e = [gm.get_node_embed(vp) for vp in vps[1:]]
        rows.append(torch.stack([torch.zeros_like(e[0])] + e, 0))
    G = max(r.shape[0] for r in rows)
    out = torch.zeros(len(rows), G, rows[0].shape[1], dtype=rows[0].dtype)
    for i, r in enumerate(rows):
        out[i, :r.shape[0]] = r
    return out
