"""The timed CPU baseline of bench.py (oracle/dyn/oracle_step_omp.c: whole flat-task step in C, OpenMP over the
environments, built -O3 -march=native on the machine that runs it) against the numpy composition of the same oracle pieces
(oracle/cpu_baseline.py::CpuAnymalStep, the portable -O2 -ffp-contract=off build).  Both are test infrastructure."""
import numpy as np

from isaacgymenv_b200 import _abi
from oracle.cpu_baseline import CpuAnymalStep, CpuAnymalStepNative
from tests import kernel_checks as kc


def _pair(n, threads):
    art = kc.load_robot("anymal")
    sp = kc.flat_params()
    props = _abi.default_dof_props(art, _abi.DOF_MODE_POS, 85.0, 2.0)
    c = kc.anymal_cfg(art)
    model = _abi.pack_model(art)
    ref = CpuAnymalStep(model, sp, props, kc.cfg_dict(c, art.num_dofs), n, threads=1, dtype=np.float32)
    nat = CpuAnymalStepNative(model, sp, props, c, n, threads=threads)
    return ref, nat


def test_native_step_matches_numpy_composition():
    n = 48
    ref, nat = _pair(n, threads=3)
    rng = np.random.default_rng(5)
    for step in range(12):          # the robots land around step 6: free fall, first touch and contact are all covered
        a = (2 * rng.random((n, 12), dtype=np.float32) - 1)
        draws = rng.random((n, 27), dtype=np.float32)
        ref.rng = type("R", (), {"random": staticmethod(lambda shape, dtype=None, d=draws: d)})()
        o_r, oc_r, r_r, t_r = ref.step(a)
        o_n, oc_n, r_n, t_n = nat.step(a, draws)
        # same algorithm, different compiler flags (FMA contraction): agreement to rounding, masks equal
        np.testing.assert_allclose(o_n, o_r, rtol=2e-3, atol=2e-3, err_msg=f"obs step {step}")
        np.testing.assert_allclose(r_n, r_r, rtol=2e-3, atol=1e-5)
        assert (nat.reset == ref.state["reset"]).mean() > 0.95
        assert (t_n == t_r).all()
        # keep the two in lock-step so rounding does not compound through contact events
        nat.root[:] = ref.state["root"]; nat.dof[:, :, 0] = ref.state["dof_pos"]; nat.dof[:, :, 1] = ref.state["dof_vel"]
        nat.reset[:] = ref.state["reset"]; nat.commands[:] = ref.state["commands"]; nat.progress[:] = ref.state["progress"]


def test_native_step_thread_count_invariant():
    n = 32
    _, a1 = _pair(n, threads=1)
    _, a4 = _pair(n, threads=4)
    rng = np.random.default_rng(1)
    for _ in range(8):
        a = (2 * rng.random((n, 12), dtype=np.float32) - 1)
        d = rng.random((n, 27), dtype=np.float32)
        o1 = a1.step(a, d)[0].copy()
        o4 = a4.step(a, d)[0].copy()
        assert np.array_equal(o1, o4)
