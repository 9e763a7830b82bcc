"""CPU: the trainable pretraining model (gridmm_b200/train_model.py) with torch fp32 linears (the tcgen05 LinearFn needs a GPU;
its GPU tests are tests/test_gpu_train.py) against the reference's own outputs and against autograd through the oracle."""
import os

import numpy as np
import torch

from tests import helpers as H
from tests.test_gpu_train import _oracle_grads, _setup


def test_pretrain_model_matches_reference_outputs_and_oracle_gradients():
    case = H.PRETRAIN_MODEL_CASE
    model, w, batch = _setup(case)
    gold = np.load(os.path.join(H.GOLD, "pretrain_heads_small.npz"))
    model.train()                                             # dropout probabilities are 0 in _setup
    gl, ll, fused = model(batch, "sap", compute_loss=False)
    for k, t in (("global_logits", gl), ("local_logits", ll), ("fused_logits", fused)):
        H.finite_close(t.detach(), gold[k], atol=2e-4)
    losses = model(batch, "sap")
    assert (losses.detach() - torch.from_numpy(gold["sap_losses"])).abs().max().item() < 5e-4
    scores = model(batch, "mlm", compute_loss=False)
    assert (scores.detach() - torch.from_numpy(gold["mlm_scores"])).abs().max().item() < 2e-4
    ref_loss, ref = _oracle_grads(w, batch, "sap", case)
    losses.mean().backward()
    tot = sum(float(g.double().pow(2).sum()) for g in ref.values())
    err = sum(float((p.grad - ref[n]).double().pow(2).sum()) for n, p in model.named_parameters() if n in ref)
    assert abs(float(losses.mean()) - ref_loss) < 1e-4
    assert (err / tot) ** 0.5 < 2e-3                          # the reference pools the grid cells in fp16, this model in fp32
    assert sum(1 for n, p in model.named_parameters() if p.grad is not None) > 200


def test_dropout_sites_only_act_in_training_mode():
    from gridmm_b200.model import NavConfig
    from gridmm_b200.train_model import PretrainModel
    case = H.PRETRAIN_MODEL_CASE
    model, w, batch = _setup(case)
    model.p_hid = model.p_att = 0.1
    model.eval()
    with torch.no_grad():
        a, b = model(batch, "sap"), model(batch, "sap")
        assert torch.equal(a, b)
        model.train()
        torch.manual_seed(0)
        c = model(batch, "sap")
        d = model(batch, "sap")
    assert not torch.equal(c, d) and not torch.equal(a, c)
    assert PretrainModel(NavConfig(pretrain_trunk=True, use_lang2visn_attn=True, **case["model"])).p_hid == 0.1      # reference default
