"""Device-side node embeddings of the topological maps (gridmm_gmap_update / gridmm_gmap_gather) against the oracle's restatement
of GraphMap.update_node_embed / get_node_embed and the agent loops around them."""
import numpy as np
import pytest
import torch


def _random_walk(B, steps, V, n_vps, seed):
    """Per step: current viewpoint and the candidates its first tokens look at (revisits allowed), plus ended flags."""
    rng = np.random.default_rng(seed)
    cur = [int(rng.integers(n_vps)) for _ in range(B)]
    seq = []
    ended = [False] * B
    for t in range(steps):
        cands = []
        for b in range(B):
            k = int(rng.integers(1, min(6, V)))
            cands.append(["vp%d" % int(x) for x in rng.choice(n_vps, size=k, replace=False)])
        seq.append((["vp%d" % c for c in cur], cands, list(ended)))
        for b in range(B):
            if rng.random() < 0.15:
                ended[b] = True
            nxt = cands[b][int(rng.integers(len(cands[b])))]
            cur[b] = int(nxt[2:])
    return seq


def test_update_index_is_pure_host_logic():
    from gridmm_b200.graph import build_gather_index, build_update_index
    sm, vis = [dict(), dict()], [set(), set()]
    cur, cand = build_update_index(sm, vis, ["a", "x"], [["b", "c"], ["y"]], [False, True], 4, 8)
    assert cur.tolist() == [0, -1] and cand.tolist() == [[1, 2, -1, -1], [-1, -1, -1, -1]]
    cur, cand = build_update_index(sm, vis, ["b", "x"], [["a", "d"], ["y"]], [False, False], 4, 8)
    assert cur.tolist() == [1, 0] and cand.tolist() == [[-1, 3, -1, -1], [1, -1, -1, -1]]      # "a" is visited: skipped
    assert build_gather_index(sm, [[None, "a", "b", "d"], [None, "y"]]).tolist() == [[-1, 0, 1, 3], [-1, 1, -1, -1]]
    with pytest.raises(ValueError):
        build_update_index([dict()], [set()], ["a"], [["b", "c", "d"]], None, 4, 2)


@pytest.mark.gpu
@pytest.mark.parametrize("B,steps,V,D", [(4, 6, 37, 768), (1, 3, 12, 512), (32, 8, 37, 768)])
def test_device_graph_maps_match_the_reference_bookkeeping(B, steps, V, D):
    from gridmm_b200.graph import DeviceGraphMaps
    from oracle import graph_oracle as go
    seq = _random_walk(B, steps, V, 14, seed=B * 10 + steps)
    g = torch.Generator().manual_seed(B)
    maps = [go.NodeEmbeds() for _ in range(B)]
    visited = [set() for _ in range(B)]
    dev = torch.device("cuda", 0)
    dg = DeviceGraphMaps(B, dim=D, max_nodes=32, device=dev)
    for cur, cands, ended in seq:
        pano = torch.randn(B, V, D, generator=g)
        lens = torch.randint(8, V + 1, (B,), generator=g)
        masks = torch.arange(V)[None, :] < lens[:, None]
        go.step_update(maps, visited, pano, masks, cur, cands, ended)
        dg.update(pano.to(dev), masks.to(dev), cur, cands, ended)
        # read-out in the agent's order: [stop] + visited + unvisited
        vpids = []
        for b in range(B):
            known = list(maps[b].node_embeds.keys())
            vpids.append([None] + [v for v in known if v in visited[b]] + [v for v in known if v not in visited[b]])
        ref = go.read_out(maps, vpids)
        got = dg.node_embeds(vpids).cpu()
        assert got.shape == ref.shape
        assert (got - ref).abs().max().item() < 2e-6


def test_teacher_actions_equal_the_reference_loop():
    """graph.teacher_actions against a literal restatement of the selection loop of map_nav_src/r2r/agent.py:207-237 on random
    maps (ties, visited masks, ended episodes, episodes standing on their goal, maps without any eligible node)."""
    from gridmm_b200.graph import teacher_actions
    rng = np.random.default_rng(11)
    vps = ["v%d" % i for i in range(9)]
    d = rng.integers(1, 5, (9, 9)).astype(float)          # small integers: ties are common
    d = np.minimum(d, d.T); np.fill_diagonal(d, 0.0)
    table = {"s": {a: {b: d[i, j] for j, b in enumerate(vps)} for i, a in enumerate(vps)}}
    for trial in range(200):
        B = 4
        obs, vpids, vis, ended = [], [], [], []
        for b in range(B):
            n = int(rng.integers(1, 7))
            nodes = [None] + list(rng.choice(vps, size=n, replace=False))
            cur = str(rng.choice(vps)); goal = cur if rng.random() < 0.2 else str(rng.choice(vps))
            obs.append({"scan": "s", "viewpoint": cur, "gt_path": ["x", goal]})
            vpids.append(nodes)
            m = rng.random(len(nodes)) < (1.0 if rng.random() < 0.1 else 0.4)
            vis.append(m.tolist())
            ended.append(bool(rng.random() < 0.2))
        masks = None if trial % 3 == 0 else vis
        want = np.zeros(B, dtype=np.int64)
        for i, ob in enumerate(obs):                      # the reference's loop, restated
            if ended[i]:
                want[i] = -100
            elif ob["viewpoint"] == ob["gt_path"][-1]:
                want[i] = 0
            else:
                best, best_d = -100, float("inf")
                for j, vp in enumerate(vpids[i]):
                    if j > 0 and (masks is None or not masks[i][j]):
                        dist = table["s"][vp][ob["gt_path"][-1]] + table["s"][ob["viewpoint"]][vp]
                        if dist < best_d:
                            best_d, best = dist, j
                want[i] = best
        got = teacher_actions(obs, vpids, ended, table, visited_masks=masks)
        assert np.array_equal(got, want), (trial, got, want)
