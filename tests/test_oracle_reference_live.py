"""Live pin of the numpy oracle (build container only): the reference's own code, imported unmodified from /root/reference
under tests/golden/ref_stubs.py, against the oracle on RANDOM cases (tests/golden/sweep_reference_vs_oracle.py; a
400-case run of that script is recorded in DESIGN.md section 4).  Skipped where the reference tree is absent (GPU box)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(not os.path.isdir("/root/reference/frenetix_motion_planner"), reason="reference tree not present")
def test_oracle_equals_the_reference_on_random_cases():
    # its own process: the sweep installs import stubs and redirects the golden directory
    out = subprocess.run([sys.executable, os.path.join(HERE, "golden", "sweep_reference_vs_oracle.py"), "8", "77000"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "8 of 8 random cases: oracle == reference" in out.stdout


@pytest.mark.skipif(not os.path.isdir("/root/reference/frenetix_motion_planner"), reason="reference tree not present")
def test_frenet_initial_state_equals_the_reference_on_random_poses():
    out = subprocess.run([sys.executable, os.path.join(HERE, "golden", "sweep_reference_vs_oracle.py"), "--initial-states", "6", "88000"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "6 of 6 random ego poses: Frenet initial state == reference" in out.stdout


@pytest.mark.skipif(not os.path.isdir("/root/reference/frenetix_motion_planner"), reason="reference tree not present")
def test_collision_probability_equals_the_reference_on_random_cases():
    out = subprocess.run([sys.executable, os.path.join(HERE, "golden", "sweep_reference_vs_oracle.py"), "--collision-probability", "200", "99000"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "all equal" in out.stdout


@pytest.mark.skipif(not os.path.isdir("/root/reference/frenetix_motion_planner"), reason="reference tree not present")
def test_level_sets_and_their_order_equal_the_reference_on_random_configurations():
    out = subprocess.run([sys.executable, os.path.join(HERE, "golden", "sweep_reference_vs_oracle.py"), "--sampling-order", "300", "91000"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "all equal" in out.stdout


@pytest.mark.skipif(not os.path.isdir("/root/reference/frenetix_motion_planner"), reason="reference tree not present")
def test_reference_path_preparation_equals_the_reference_on_random_centre_lines():
    out = subprocess.run([sys.executable, os.path.join(HERE, "golden", "sweep_reference_vs_oracle.py"), "--refpath", "40", "93000"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "all equal" in out.stdout


@pytest.mark.skipif(not os.path.isdir("/root/reference/frenetix_motion_planner"), reason="reference tree not present")
def test_trajectory_pair_equals_the_reference_on_random_trajectories():
    out = subprocess.run([sys.executable, os.path.join(HERE, "golden", "sweep_reference_vs_oracle.py"), "--trajectory-pair", "60", "95000"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "all equal" in out.stdout


@pytest.mark.skipif(not os.path.isdir("/root/reference/frenetix_motion_planner"), reason="reference tree not present")
def test_boundary_harm_model_is_the_reference_function_with_the_reference_coefficients():
    """planner.py:370-381: harm of leaving the road = get_protected_inj_prob_log_reg_ignore_angle at the velocity of the first
    overlapping step, coefficients from configurations/harm_parameters.json."""
    import importlib.util
    import json
    import numpy as np
    from frenetix_motion_planner_b200 import trajectories
    coeff = json.load(open("/root/reference/configurations/harm_parameters.json"))
    assert coeff["log_reg"]["ignore_angle"] == trajectories.DEFAULT_HARM_COEFF
    spec = importlib.util.spec_from_file_location("lr_sym", "/root/reference/risk_assessment/utils/logistic_regression_symmetrical.py")
    lr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(lr)
    c = trajectories.DEFAULT_HARM_COEFF
    for v in np.random.default_rng(5).uniform(0.0, 40.0, 200):
        want = lr.get_protected_inj_prob_log_reg_ignore_angle(velocity=float(v), coeff=coeff)
        assert float(1.0 / (1.0 + np.exp(-c["const"] - c["speed"] * float(v)))) == want


@pytest.mark.skipif(not os.path.isdir("/root/reference/frenetix_motion_planner"), reason="reference tree not present")
def test_velocity_interval_and_input_record_equal_the_reference_methods():
    """Planner.set_desired_velocity (planner.py:292-310) and record_state_and_input (:244-262), run unmodified on a bare
    object, against the same methods of ReactivePlannerB200."""
    code = r'''
import sys, types
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import numpy as np
import ref_stubs
ref_stubs.install()
import importlib
pl = importlib.import_module("frenetix_motion_planner.planner")
pl.InputState = lambda **k: types.SimpleNamespace(**k)
from frenetix_motion_planner_b200.reactive_planner_b200 import ReactivePlannerB200
from frenetix_motion_planner_b200 import synthetic as syn
rng = np.random.default_rng(11)
class Log:
    def info(self, *a, **k): pass
for _ in range(500):
    got, want = [], []
    veh = types.SimpleNamespace(**dict(syn.VEHICLE_2, a_max=float(rng.uniform(2, 12)), v_max=float(rng.uniform(10, 60))))
    horizon = float(rng.choice([2.0, 3.0, 5.0]))
    a = types.SimpleNamespace(vehicle_params=veh, horizon=horizon, msg_logger=Log(), sampling_handler=types.SimpleNamespace(set_v_sampling=lambda lo, hi: want.append((lo, hi))))
    b = types.SimpleNamespace(vehicle_params=veh, horizon=horizon, sampling_handler=types.SimpleNamespace(set_v_sampling=lambda lo, hi: got.append((lo, hi))))
    v_des, v, lim = float(rng.uniform(0, 20)), float(rng.uniform(0, 40)), float(rng.choice([36.0, rng.uniform(5, 30)]))
    pl.Planner.set_desired_velocity(a, v_des, v, v_limit=lim)
    ReactivePlannerB200.set_desired_velocity(b, v_des, v, v_limit=lim)
    assert got == want and a.desired_velocity == b.desired_velocity, (got, want)
a = types.SimpleNamespace(record_state_list=[], record_input_list=[], dT=0.1)
b = types.SimpleNamespace(record_state_list=[], record_input_list=[], dT=0.1)
for k in range(50):
    st = types.SimpleNamespace(time_step=k, acceleration=float(rng.normal()), steering_angle=float(rng.normal(0, 0.2)))
    pl.Planner.record_state_and_input(a, st)
    ReactivePlannerB200.record_state_and_input(b, st)
assert [vars(x) for x in a.record_input_list] == [vars(x) for x in b.record_input_list]
print("planner methods equal")
'''
    out = subprocess.run([sys.executable, "-c", code, os.path.join(HERE, "golden"), os.path.dirname(HERE)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "planner methods equal" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


@pytest.mark.skipif(not os.path.isdir("/root/reference/frenetix_motion_planner"), reason="reference tree not present")
def test_optional_cost_terms_equal_the_reference_on_random_samples():
    out = subprocess.run([sys.executable, os.path.join(HERE, "golden", "sweep_reference_vs_oracle.py"), "--inactive-costs", "300", "97000"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "all equal" in out.stdout
