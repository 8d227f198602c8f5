"""Property tests on the CPU (hypothesis): the two restatements of the reference's algorithm -- the numpy
oracle, pinned to outputs of the reference's own code, and the C oracle, which checks the CUDA path at sizes numpy cannot
reach -- must agree on RANDOM sampling rows, reference paths, Frenet states, obstacle sets and walls, not only on the golden
cases; and the CPU implementation of the C ABI (the timed baseline) must answer what the C oracle answers."""
import numpy as np
from hypothesis import given, settings, strategies as st, HealthCheck

from helpers import BAND, compare_with_oracle
from oracle import c_oracle, frenet_oracle as fo
from test_c_oracle import _as_dev
from test_gpu_properties import make_case

SETTINGS = dict(max_examples=40, deadline=None, derandomize=True,
                suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
case_args = dict(seed=st.integers(0, 2 ** 31 - 1), n_rows=st.integers(1, 400), path_kind=st.integers(0, 2), low_vel=st.booleans(),
                 draw=st.booleans(), debug=st.booleans(), n_obs=st.integers(0, 9), walls=st.integers(0, 3))
T_VALUES = np.round(np.arange(5, 31) * 0.1, 2)
_CPU_LIB = {}


@settings(**SETTINGS)
@given(**case_args)
def test_c_oracle_equals_numpy_oracle_on_random_inputs(seed, n_rows, path_kind, low_vel, draw, debug, n_obs, walls):
    S, ref, prm, preds, static = make_case(seed, n_rows, path_kind, low_vel, draw, debug, n_obs, walls)
    ora = fo.plan(S, ref, prm, preds, static_obbs=static)
    c = c_oracle.plan(S, ref, prm, preds, static_obbs=static, T_values=T_VALUES)
    # 1e-9 on the golden cases (test_c_oracle.py); random rows come close to the Frenet singularity now and then, where an ulp
    # of libm-vs-numpy difference in theta is amplified (600 random cases: worst 4e-9, in kappa)
    compare_with_oracle(_as_dev(c), ora, prm, tol=1e-7)
    assert np.array_equal(c["margins"] < BAND, ora["margins"] < BAND)


@settings(**SETTINGS)
@given(**case_args)
def test_cpu_abi_equals_c_oracle_on_random_inputs(seed, n_rows, path_kind, low_vel, draw, debug, n_obs, walls):
    from frenetix_motion_planner_b200 import _capi
    from helpers import configure_handler
    from oracle.build import build_cpu_abi
    S, ref, prm, preds, static = make_case(seed, n_rows, path_kind, low_vel, draw, debug, n_obs, walls)
    c = c_oracle.plan(S, ref, prm, preds, static_obbs=static, T_values=T_VALUES, check_all_collisions=False)   # lazy walk, like the ABI
    if "lib" not in _CPU_LIB:
        _CPU_LIB["lib"] = _capi.load_library(path=build_cpu_abi())
    h = _capi.Handler(0, library=_CPU_LIB["lib"])
    configure_handler(h, ref, prm, preds, static, T_values=T_VALUES)
    res = h.plan(np.ascontiguousarray(S))
    flags, traj_len = h.get_flags()
    costs, total = h.get_costs()
    assert int(res.argmin) == c["argmin"] and int(res.n_feasible) == c["n_feasible"] and int(res.collision_counter) == c["collision_counter"]
    assert np.array_equal(flags, c["flags"]) and np.array_equal(traj_len, c["traj_len"])
    assert np.array_equal(total, c["total"], equal_nan=True) and np.array_equal(costs, c["costs"], equal_nan=True)
    assert np.array_equal(h.get_states_range(), c["states"], equal_nan=True)
