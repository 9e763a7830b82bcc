"""forward_pretrain(batch, "sap", heads=True) -- the SAP action logits of the pretraining wrapper
(pretrain_src/model/pretrain_cmt.py:214-270) through the navigation heads -- against the reference's own
GlocalTextPathCMTPreTraining.forward_sap on the collated batch of tests/golden/pretrain_heads_small.npz.  The host side of this
path is covered on CPU (tests/test_cpu_host.py::test_pretrain_sap_heads_glue_masks_and_candidates) and every kernel on it is the
navigation step's.  The reference pools in fp16 there (the trunk outputs measured 2e-3), hence the looser tolerance."""
import json
import os

import numpy as np
import pytest
import torch

from gridmm_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_pretrain_sap_logits_match_reference_golden():
    from gridmm_b200.model import GlocalTextPathNavCMT, remap_pretrained_keys
    case = H.PRETRAIN_MODEL_CASE
    gold = np.load(os.path.join(H.GOLD, "pretrain_heads_small.npz"))
    shapes = json.load(open(os.path.join(H.GOLD, "pretrain_heads_small_spec.json")))
    w = synth.make_weights(shapes, seed=case["seed"])
    model = GlocalTextPathNavCMT(H.make_config(use_lang2visn_attn=True, graph_sprels=False, **case["model"]))
    res = model.load_state_dict(remap_pretrained_keys({k: torch.from_numpy(v) for k, v in w.items()}), strict=False)
    assert not res.missing_keys
    model = model.cuda().eval()
    batch = H.pretrain_batch(case)
    pb = synth.make_pretrain_batch(case["batch"], seed=case["seed"], txt_len=case["txt_len"], max_steps=case["max_steps"])
    batch["gmap_visited_masks"] = torch.from_numpy(synth.make_pretrain_labels(pb, seed=case["seed"])["gmap_visited_masks"])
    out = model.forward_pretrain(batch, task="sap", heads=True)
    torch.cuda.synchronize()
    errs = {k: H.finite_close(out[k], gold[k], atol=5e-3) for k in ("global_logits", "local_logits", "fused_logits")}
    print("pretraining SAP logits errors", errs)
