"""CPU: pins the oracle (oracle/rba_oracle.py) against the committed golden fixtures, which are outputs of the
unmodified reference run in the build container (oracle/make_golden.py)."""
import pytest
import torch

import rba_oracle as O
from conftest import load_golden
from golden_cases import CASES, case_images, case_model_config, state_checksum
from rba_b200 import weights

ORACLE_TOL = 5e-5   # oracle vs reference: fp32 round-off only (op-for-op restatement)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(name):
    fix = load_golden(f"model_{name}.pt")
    case = fix["case"]
    mc = case_model_config(case)
    sd = weights.init_state_dict(mc, seed=case["seed"], perturb=case["perturb"])
    assert abs(state_checksum(sd) - fix["state_checksum"]) <= 1e-6 * fix["state_checksum"], "weight RNG drifted"
    torch.set_num_threads(max(1, min(8, torch.get_num_threads())))
    out = O.forward(sd, mc, case_images(case), want_taps=True)
    # the case is well conditioned: no boolean attention-mask decision sits within 1e-3 of its threshold
    assert fix["am_margin"] >= 1e-3 and abs(out["taps"]["am_margin"] - fix["am_margin"]) < 1e-4
    assert (out["pred_logits"] - fix["pred_logits"]).abs().max() < ORACLE_TOL
    assert (out["pred_masks"] - fix["pred_masks"]).abs().max() < ORACLE_TOL
    sem = torch.stack(out["sem_seg"])
    rba = torch.stack(out["rba"])
    assert (sem[:, :, ::4, ::4] - fix["sem_seg_s4"]).abs().max() < ORACLE_TOL
    assert (rba - fix["rba"]).abs().max() < ORACLE_TOL
    if case.get("ood_prediction"):      # DenseHybrid head: model(..., return_ood_pred=True) + get_densehybrid_score
        assert (out["ood_pred"] - fix["ood_pred"]).abs().max() < ORACLE_TOL
        assert (torch.stack(out["densehybrid"]) - fix["densehybrid"]).abs().max() < ORACLE_TOL
    else:
        assert "ood_pred" not in out


def test_oracle_msda_matches_reference_golden():
    fix = load_golden("msda.pt")
    for f in (fix, fix["big"]):
        for fn in (O.msda_bilinear_gather, O.msda_core_grid_sample):
            out = fn(f["value"], f["shapes"], f["loc"], f["aw"])
            # tolerances of the reference's own float check (ops/test.py:59)
            assert torch.allclose(out, f["out"], rtol=1e-2, atol=1e-3)
            assert (out - f["out"]).abs().max() < 1e-5
    outd = O.msda_bilinear_gather(fix["value"].double(), fix["shapes"], fix["loc"].double(), fix["aw"].double())
    assert torch.allclose(outd, fix["out_double"])          # ops/test.py:43 (double, default allclose)


def test_oracle_msda_backward_matches_reference_gradients():
    """The oracle's explicit restatement of the reference CUDA backward (col2im_bilinear) against float64 autograd
    gradients of the reference's own ms_deform_attn_core_pytorch (oracle/make_golden_msda_grad.py)."""
    from test_kernels_gpu import _msda_grad_inputs
    fix = load_golden("msda_grad.pt")
    for name, f in fix.items():
        shapes, value, loc, aw, go = _msda_grad_inputs(f["case"])
        gv, gl, ga = O.msda_bilinear_backward(value.double(), shapes, loc.double(), aw.double(), go.double())
        for got, want, nm in ((gv, f["grad_value"], "grad_value"), (gl, f["grad_loc"], "grad_loc"), (ga, f["grad_aw"], "grad_aw")):
            assert (got.float() - want).abs().max() <= 2e-6 * max(1.0, want.abs().max().item()), (name, nm)


def test_oracle_outlier_loss_matches_reference_criterion():
    """The oracle's restatement of SetCriterion.outlier_loss against the reference's loss value and float64 gradients."""
    from test_kernels_gpu import _outlier_inputs
    fix = load_golden("outlier_loss.pt")
    for name, f in fix.items():
        c = f["case"]
        masks, logits, labels = _outlier_inputs(c)
        m, l = masks.double().requires_grad_(True), logits.double().requires_grad_(True)
        loss = O.outlier_loss(m, l, labels, c["target"], c["norm"], c["t_in"], c["t_out"])
        loss.backward()
        assert abs(float(loss) - f["loss"]) < 1e-9 * max(1.0, abs(f["loss"])), name
        assert (m.grad.float() - f["d_masks"]).abs().max() < 1e-9 and (l.grad.float() - f["d_logits"]).abs().max() < 1e-9, name


def test_oracle_score_matches_reference_golden():
    fix = load_golden("score.pt")
    for nm, f in fix.items():
        for b in range(f["masks"].shape[0]):
            h, w = f["masks"].shape[-2:]
            sem, rba = O.score_from_head_outputs(f["logits"][b], f["masks"][b], (4 * h, 4 * w), (f["H"], f["W"]))
            assert (sem - f["sem_seg"][b]).abs().max() < 1e-5
            assert (rba - f["rba"][b]).abs().max() < 1e-5


def test_oracle_shift_mask_and_index_are_analytic():
    """The kernels evaluate relative_position_index and the SW-MSA region mask analytically; pin the closed forms."""
    ws = 12
    idx = O.relative_position_index(ws)
    r = torch.arange(ws * ws)
    ri, rj = r // ws, r % ws
    closed = (ri[:, None] - ri[None, :] + ws - 1) * (2 * ws - 1) + (rj[:, None] - rj[None, :] + ws - 1)
    assert torch.equal(idx, closed)
    H, W, shift = 30, 41, 6
    m = O.shift_attn_mask(H, W, ws, shift)
    nWh, nWw = -(-H // ws), -(-W // ws)
    Hp, Wp = nWh * ws, nWw * ws
    for win in range(nWh * nWw):
        wh, ww = win // nWw, win % nWw
        hs, wsx = wh * ws + ri, ww * ws + rj
        rid = torch.where(hs < Hp - ws, 0, torch.where(hs < Hp - shift, 1, 2)) * 3 + \
            torch.where(wsx < Wp - ws, 0, torch.where(wsx < Wp - shift, 1, 2))
        closed_m = torch.where(rid[:, None] != rid[None, :], -100.0, 0.0)
        assert torch.equal(m[win], closed_m)
