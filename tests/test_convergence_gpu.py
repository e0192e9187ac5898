"""GPU: the trained Y0 against the reference's known answers (with Y == A the coupling vanishes, so the
exact Y0 is the closed-form price; mainMerton.py:68-73 and mainVG.py:65-70 plot the learned Y0 against it).  The runs are those of
scripts/convergence.py (full-length curves: profiles/r2_convergence_*.csv):

  SolverGlobalFBSDE at the mainMerton.py / mainVG.py defaults (10 paths x 5000 compensator samples, 120 x 100 steps): the learned
  Y0 must sit inside the 3-sigma Monte-Carlo confidence interval of a price estimate from the paths of ONE outer epoch (100 batches),
  and within 4 % of the closed form;
  SolverGlobalSumLocalReg on the d = 10 basket at ~2^16 paths per step: within 10 % of 0.1109224 after 6000 steps (the regression
  scheme's own error - shared time-dependent network over 100 steps - not Monte-Carlo noise: the same curve shape comes out of the
  reference-equivalent CPU restatement, profiles/r2_convergence_oracle_merton_reg.csv)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case,closed,rel", [("merton_global", 0.2714569, 0.04), ("vg_global", 0.1331402, 0.04)])
def test_global_solver_learns_the_closed_form_price_within_the_mc_interval(ctx, case, closed, rel):
    from convergence import run_case
    r = run_case(case, False, ctx, write=False)
    print({k: r[k] for k in ("case", "closed_form", "Y0_last", "Y0_tail_mean", "abs_err_tail_mean", "se_batch", "se_epoch", "mc_price",
                             "mc_price_se", "train_seconds")})
    assert abs(r["closed_form"] - closed) < 2e-5
    assert abs(r["mc_price"] - closed) < 4 * r["mc_price_se"]                 # the MC estimator itself brackets the closed form
    assert r["abs_err_tail_mean"] <= 3 * r["se_epoch"], r
    assert r["abs_err_tail_mean"] <= rel * closed, r


def test_basket_d10_reg_solver_approaches_the_closed_form(ctx):
    from convergence import run_case
    r = run_case("basket_d10", False, ctx, write=False)
    print({k: r[k] for k in ("case", "closed_form", "Y0_last", "Y0_tail_mean", "abs_err_tail_mean", "se_batch", "mc_price", "mc_price_se",
                             "train_seconds")})
    assert abs(r["closed_form"] - 0.1109224) < 2e-6 and abs(r["mc_price"] - r["closed_form"]) < 4 * r["mc_price_se"]
    assert r["abs_err_tail_mean"] <= 0.10 * r["closed_form"], r
