"""CPU: the trainable navigation model (gridmm_b200/train_nav.py) with torch fp32 linears (the tcgen05 LinearFn needs a GPU: its
path is covered by tests/test_gpu_train.py::test_trainable_nav_on_the_gpu) against the reference's own outputs (golden files
produced by oracle/make_golden.py from the unmodified reference) and against autograd through the oracle."""
import os

import numpy as np
import pytest
import torch

from gridmm_b200 import synth
from tests import helpers as H


def _nav_setup(name):
    from gridmm_b200.train_nav import TrainableNavCMT
    ep_kw, nav_kw, model_kw = H.NAV_CASES[name]
    cfg = H.make_config(**model_kw)
    w = H.make_weights(cfg, ep_kw["seed"])
    model = TrainableNavCMT(cfg, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    res = model.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()}, strict=True)
    assert not res.missing_keys and not res.unexpected_keys      # same state_dict keys as the reference
    ep = synth.make_episodes(dim=768, **ep_kw)
    cells, fts, _, pos = H.oracle_grid(ep)
    return model, cfg, w, H.nav_batch(ep_kw, nav_kw, cells, fts, pos)


@pytest.mark.parametrize("name", sorted(H.NAV_CASES))
def test_trainable_nav_matches_reference_golden(name):
    model, cfg, w, nav = _nav_setup(name)
    gold = np.load(os.path.join(H.GOLD, "nav_%s.npz" % name))
    model.train()                                             # dropout probabilities are 0: train() == eval() numerically
    out = model("navigation", nav)
    for k in gold.files:
        H.finite_close(out[k].detach(), gold[k], atol=5e-5)
    if "obj_logits" not in gold.files:
        assert out["obj_logits"] is None


@pytest.mark.parametrize("name", sorted(H.AUX_CASES))
def test_trainable_language_panorama_match_reference_golden(name):
    from gridmm_b200.train_nav import TrainableNavCMT
    seed, model_kw, in_kw = H.AUX_CASES[name]
    gold = np.load(os.path.join(H.GOLD, "aux_%s.npz" % name))
    cfg = H.make_config(**model_kw)
    model = TrainableNavCMT(cfg, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in H.make_weights(cfg, seed).items()})
    if name.startswith("lang"):
        out = model("language", synth.to_torch(synth.make_lang_inputs(seed=seed, **in_kw)))
        H.finite_close(out.detach(), gold["txt_embeds"], atol=5e-5)
    else:
        emb, masks = model("panorama", synth.to_torch(synth.make_pano_inputs(seed=seed, **in_kw)))
        assert np.array_equal(masks.numpy(), gold["pano_masks"])
        H.finite_close(emb.detach(), gold["pano_embeds"], atol=5e-5)


def test_trainable_nav_gradients_match_autograd_through_the_oracle():
    """The imitation-learning loss of the reference's train loop (cross entropy on the fused logits, r2r/agent.py:383-386) through
    this model and through the oracle's restatement of forward('navigation'): same loss, same gradients."""
    from oracle import model_oracle as mo
    model, cfg, w, nav = _nav_setup("reverie_small")
    B = nav["gmap_masks"].shape[0]
    finite = torch.isfinite(model("navigation", nav)["fused_logits"].detach())
    target = torch.tensor([int(torch.nonzero(finite[b])[-1]) for b in range(B)])
    model.zero_grad()
    out = model("navigation", nav)
    loss = torch.nn.functional.cross_entropy(out["fused_logits"], target, reduction="sum") + out["obj_logits"][torch.isfinite(out["obj_logits"])].sum() * 0.1
    loss.backward()
    sd = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in w.items()}
    ref = mo.navigation(sd, nav, n_x_layers=cfg.num_x_layers)
    ref_loss = torch.nn.functional.cross_entropy(ref["fused_logits"], target, reduction="sum") + ref["obj_logits"][torch.isfinite(ref["obj_logits"])].sum() * 0.1
    ref_loss.backward()
    assert abs(float(loss) - float(ref_loss)) < 1e-4
    tot = err = 0.0
    n_grad = 0
    for n, p in model.named_parameters():
        g_ref = sd[n].grad
        if g_ref is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
            continue
        n_grad += 1
        tot += float(g_ref.double().pow(2).sum())
        err += float((p.grad - g_ref).double().pow(2).sum())
    assert n_grad > 150 and (err / tot) ** 0.5 < 1e-4


def test_trainable_continuous_env_variant_matches_reference_golden():
    """The 14-tuple calling convention of the continuous-env policy (BASELINE config 4) against the reference's CE model."""
    from oracle import grid_oracle as go
    from gridmm_b200.train_nav import TrainableNavCMT
    ep_kw, nav_kw = H.CE_NAV_CASE
    gold = np.load(os.path.join(H.GOLD, "nav_ce_small.npz"))
    cfg = H.make_config(graph_sprels=False)
    model = TrainableNavCMT(cfg, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in H.make_weights(cfg, ep_kw["seed"]).items()})
    cells, fts, _, pos = H.oracle_grid(H.ce_episodes(ep_kw), geom=go.CEGeometry)
    tup = H.ce_nav_tuple(ep_kw, nav_kw, cells, fts, pos)
    out = model("navigation", tup)
    H.finite_close(out.detach(), gold["fused_logits"], atol=5e-5)
    out[torch.isfinite(out)].sum().backward()
    assert sum(1 for p in model.parameters() if p.grad is not None and float(p.grad.abs().max()) > 0) > 100


def test_wrapper_applies_feature_dropout_only_in_training():
    from types import SimpleNamespace
    from gridmm_b200.train_nav import VLNBertTrainable
    seed, model_kw, in_kw = H.AUX_CASES["pano_r2r"]
    args = SimpleNamespace(feat_dropout=0.4, num_l_layers=1, num_pano_layers=2, num_x_layers=4, image_feat_size=768, angle_feat_size=4,
                           obj_feat_size=0, graph_sprels=True, fusion="dynamic")
    vb = VLNBertTrainable(args)
    assert vb.vln_bert.p_hid == 0.1 and vb.drop_env.p == 0.4
    batch = synth.to_torch(synth.make_pano_inputs(seed=seed, **in_kw))
    vb.eval()
    with torch.no_grad():
        a, _ = vb("panorama", dict(batch))
        b, _ = vb("panorama", dict(batch))
        assert torch.equal(a, b)
        vb.train()
        c, _ = vb("panorama", dict(batch))
    assert not torch.equal(a, c)
    with pytest.raises(NotImplementedError):
        vb("waypoint", {})
    # weights travel between the trainable and the inference model: identical state_dict keys
    from gridmm_b200.model import param_spec
    assert list(vb.vln_bert.state_dict().keys()) == list(param_spec(vb.vln_bert.config).keys())
