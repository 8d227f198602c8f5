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
