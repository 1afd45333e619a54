"""Pin the oracle: the CPU restatement (oracle/hsddp_oracle.cpp) against the reference's OWN solver sources, compiled
unmodified where they lie under /root/reference against the Eigen stand-in oracle/eigen_shim (oracle/Makefile target
`reference`; outputs in oracle/_ref/).  Both share the same documented evaluation order for the small dense products,
so the comparison is bit-exact: it checks control flow, quirks and every scalar of the restatement, not rounding."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_specs, to_orc_problem  # noqa: E402
from oracle import reference_py  # noqa: E402

pytestmark = pytest.mark.skipif(not reference_py.available("strict"), reason="reference library not built and reference tree absent")

CASES = [(1, 0, {}), (2, 1, {}), (2, 3, {}), (3, 1, {}), (4, 0, {}), (4, 5, {}), (2, 2, dict(DDP_thresh=-1.0, AL_thresh=-1.0)),
         (2, 1, dict(max_AL_iter=3.0, max_DDP_iter=4.0))]


@pytest.mark.parametrize("cfg,idx,opt", CASES)
def test_solve_bit_exact_against_reference_sources(cfg, idx, opt):
    from oracle.oracle_py import OracleSolver
    spec = make_specs(idx + 1, cfg)[idx]
    ref = reference_py.ReferenceSolver(spec, "strict", **opt)
    orc = OracleSolver(to_orc_problem(spec, **opt), variant="strict")
    ref.initialization()
    orc.initialization()
    modes, N, _ = orc.schedule()
    assert (modes, N) == ref.schedule()
    for p in range(len(N)):   # warm start
        assert np.array_equal(orc.field(p, "X_NOM"), ref.field(p, "X_NOM"))
        assert np.array_equal(orc.field(p, "U_NOM"), ref.field(p, "U_NOM"))
    ref.solve()
    orc.solve()
    so, sr = orc.scalars(), ref.scalars()
    assert so["cost"] == sr["cost"] and so["exp_cost_change"] == sr["exp_cost_change"] and so["viol"] == sr["viol"]
    for p in range(len(N)):
        for nm in ("X_NOM", "U_NOM", "Y_NOM", "K", "DU", "G", "H", "QX", "QU", "QXX", "QUX", "QUU"):
            assert np.array_equal(orc.field(p, nm), ref.field(p, nm)), (p, nm)


def test_stepwise_api_bit_exact():
    from oracle.oracle_py import OracleSolver
    spec = make_specs(2, 2)[1]
    ref = reference_py.ReferenceSolver(spec, "strict")
    orc = OracleSolver(to_orc_problem(spec), variant="strict")
    for s in (ref, orc):
        s.initialization()
    ref.forward_sweep(0.0); orc.forward_sweep(0.0)
    assert ref.scalars()["cost"] == orc.scalars()["cost"]
    ref.update_nominal_trajectory(); orc.update_nominal_trajectory()
    assert ref.backward_sweep(0.0) == orc.backward_sweep(0.0) == False  # noqa: E712  (singular Quu)
    assert ref.backward_sweep(1e-3) == orc.backward_sweep(1e-3) == True  # noqa: E712
    assert ref.scalars()["exp_cost_change"] == orc.scalars()["exp_cost_change"]
    assert ref.forward_iteration() == orc.forward_iteration()
    assert ref.scalars()["cost"] == orc.scalars()["cost"]


def test_fast_build_agrees_to_rounding():
    from oracle.oracle_py import OracleSolver
    spec = make_specs(1, 1)[0]
    ref = reference_py.ReferenceSolver(spec, "fast")
    orc = OracleSolver(to_orc_problem(spec), variant="strict")
    ref.initialization(); orc.initialization()
    ref.solve(); orc.solve()
    assert abs(ref.scalars()["cost"] - orc.scalars()["cost"]) <= 1e-7 * abs(orc.scalars()["cost"])


@pytest.mark.parametrize("cfg,idx,opt", [(1, 0, {}), (4, 0, {}), (2, 2, dict(DDP_thresh=-1.0, AL_thresh=-1.0)), (2, 3, dict(n_wbphase=3, n_fbphase=2))])
def test_update_problem_receding_horizon_bit_exact(cfg, idx, opt):
    """MHPCLocomotion::update_problem (MHPCLocomotion.cpp:107-158; never called by the reference's main but part of the
    compiled class): three solve -> shift -> new x0 -> solve ticks, restatement vs the reference sources."""
    from oracle.oracle_py import OracleSolver
    opt = dict(opt)
    spec = dict(make_specs(idx + 1, cfg)[idx])
    for k in ("n_wbphase", "n_fbphase"):
        if k in opt:
            spec[k] = opt.pop(k)
    ref = reference_py.ReferenceSolver(spec, "strict", **opt)
    orc = OracleSolver(to_orc_problem(spec, **opt), variant="strict")
    for s in (ref, orc):
        s.initialization()
        s.solve()
    for tick in range(3):
        xnew = orc.field(0, "X_NOM")[12]
        for s in (ref, orc):
            s.update_problem()
            s.set_x0(xnew)
            s.solve()
        modes, N, _ = orc.schedule()
        assert (modes, N) == ref.schedule()
        a, b = ref.scalars(), orc.scalars()
        assert a["cost"] == b["cost"] and a["viol"] == b["viol"] and a["exp_cost_change"] == b["exp_cost_change"]
        for p in range(len(N)):
            for nm in ("X_NOM", "U_NOM", "Y_NOM", "K", "DU", "G"):
                assert np.array_equal(orc.field(p, nm), ref.field(p, nm), equal_nan=True), (tick, p, nm)
