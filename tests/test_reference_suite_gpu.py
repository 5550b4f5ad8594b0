"""GPU: the REFERENCE'S OWN test classes for the hot path, run against the drop-in.

SURVEY.md §4 / §8(c) name them as the interface-conformance suite: the reference's tests are behavioural (shapes,
signs, bounds, determinism, exceptions), so they can be pointed at any implementation of the same interfaces.  The
test files are read from the git-ignored snapshot baseline/_ref/reference_tests (baseline/install_reference.py copies
them from /root/reference/tests, unmodified) and `quadcopter_tracking` is resolved through the test-only alias
package tests/refshim/ — i.e. every `from quadcopter_tracking... import X` in those files gets the B200 class.

Files run, whole (every class in them; /root/reference/tests): test_env_dynamics.py (target motion, env, PID / LQR /
Riccati-LQR / feedforward / LQI controllers, hover-thrust and axis-sign integration, ENU frame helpers),
test_eval.py (metrics, summaries, Evaluator incl. its plots, load_controller, controller selection, action schema,
LQI evaluation), test_tuning.py (search space, tuner strategies incl. CMA-ES and its checkpoint, Riccati tuning) and
test_training_loop.py (policy network, losses, Trainer and its modes; its TestDiagnostics class is a listed deviation).
SURVEY.md §8(c) names the classes TestTargetMotion, TestQuadcopterEnv, TestRiccatiLQR* / TestRiccatiLQI,
TestRiccatiControllerSelection, TestLQIController and TestRiccatiTuning; they are part of these files.
The reference itself passes all but one of them (a reference bug — the test expects the message
"q_int must have 3 elements", riccati_lqr.py:583-587 says "exactly 3 elements").  Every test that does not pass here is
listed in KNOWN_DEVIATIONS with its reason; anything else failing fails this test.

Round-2 correction: until session w the snapshot was run from baseline/_ref/reference_tests, a package directory whose
parent also holds the reference's own `quadcopter_tracking`; pytest's default import mode puts that parent first on
sys.path, so those runs imported the REFERENCE, not the drop-in.  The suite is now run from a copy with a clean parent
and the test asserts where `quadcopter_tracking` came from.
"""
import os
import pathlib
import subprocess
import sys
import xml.etree.ElementTree as ET

import pytest

pytestmark = pytest.mark.gpu

ROOT = pathlib.Path(__file__).resolve().parent.parent
REF_TESTS = ROOT / "baseline" / "_ref" / "reference_tests"

FILES = ["test_env_dynamics.py", "test_eval.py", "test_tuning.py", "test_training_loop.py"]

# test name -> why it does not pass against the drop-in (each is a deliberate, documented difference or a reference bug)
KNOWN_DEVIATIONS = {
    "test_lqi_invalid_q_int_length": "reference bug (SURVEY.md §4): the reference fails its own test too — the test "
                                          "matches 'q_int must have 3 elements' but riccati_lqr.py:583-587 (and the drop-in, "
                                          "which reproduces that message) say 'exactly 3 elements'",
}
# TestDiagnostics (test_training_loop.py:974-1290) exercises utils/diagnostics.py, the training-diagnostics recorder
# SURVEY.md §2 marks out of scope; it is not built and its thirteen tests are listed, not hidden
KNOWN_DEVIATIONS.update({name: "out of scope (SURVEY.md §2): the training diagnostics module utils/diagnostics.py is not built"
                         for name in (
    "test_diagnostics_config_defaults", "test_diagnostics_config_from_dict", "test_diagnostics_disabled_by_default",
    "test_diagnostics_enabled_logging", "test_diagnostics_gradient_stats", "test_diagnostics_observation_stats",
    "test_diagnostics_observation_stats_handles_nan", "test_diagnostics_action_stats", "test_diagnostics_save_files",
    "test_training_with_diagnostics_enabled", "test_diagnostics_throttling", "test_diagnostics_max_entries_per_epoch")})


def test_reference_test_classes_pass_against_the_drop_in(gpu_ctx, tmp_path):
    if not REF_TESTS.is_dir():
        pytest.skip("baseline/_ref/reference_tests is absent: run `python baseline/install_reference.py` in the build "
                    "container (the reference's files are never committed, only shipped)")
    # The snapshot directory is a package (tests/__init__.py) whose parent, baseline/_ref, also holds the reference's own
    # `quadcopter_tracking`: pytest's default import mode would put that parent first on sys.path and the reference's
    # tests would then test the reference.  They are therefore run from a copy whose parent directory holds nothing
    # else, and a conftest.py next to the copy records where `quadcopter_tracking` was imported from.
    import shutil

    suite = tmp_path / "suite" / "reference_tests"
    shutil.copytree(REF_TESTS, suite, ignore=shutil.ignore_patterns("__pycache__"))
    origin = tmp_path / "quadcopter_tracking_origin.txt"
    (tmp_path / "suite" / "conftest.py").write_text(
        "def pytest_sessionfinish(session, exitstatus):\n"
        "    import sys\n"
        "    mods = [m for n, m in sorted(sys.modules.items()) if n == 'quadcopter_tracking' or n.startswith('quadcopter_tracking.')]\n"
        f"    open({str(origin)!r}, 'w').write('\\n'.join(str(getattr(m, '__file__', None)) for m in mods))\n")
    nodes = [str(suite / f) for f in FILES]
    env = dict(os.environ)
    paths = [str(ROOT / "tests" / "refshim"), str(ROOT)]
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        paths.append(str(ROOT / "baseline" / "_ref" / "_stubs"))
    env["PYTHONPATH"] = os.pathsep.join(paths + [env.get("PYTHONPATH", "")])
    env["TUNING_OUTPUT_DIR"] = ""
    env.pop("TUNING_OUTPUT_DIR")
    xml = tmp_path / "reference_suite.xml"
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--rootdir", str(tmp_path),
                        "-o", "python_files=test_*.py", "--tb=short", f"--junitxml={xml}", *nodes],
                       cwd=tmp_path, env=env, capture_output=True, text=True, timeout=3000)
    assert xml.exists(), r.stdout[-3000:] + r.stderr[-3000:]
    files = origin.read_text().split()
    shim = str(ROOT / "tests" / "refshim")
    assert files and all(f.startswith(shim) for f in files), \
        f"the reference's tests did not import the drop-in: quadcopter_tracking modules came from {sorted(set(files))[:4]}"
    passed, failed, skipped = [], {}, []
    for case in ET.parse(xml).getroot().iter("testcase"):
        name = case.get("name")
        bad = case.find("failure") if case.find("failure") is not None else case.find("error")
        if bad is not None:
            failed[name] = (case.get("classname", ""), (bad.get("message") or "")[:300])
        elif case.find("skipped") is not None:
            skipped.append(name)
        else:
            passed.append(name)
    total = len(passed) + len(failed) + len(skipped)
    unexpected = {k: v for k, v in failed.items() if k not in KNOWN_DEVIATIONS}
    print(f"[conformance] reference test classes against the drop-in: {len(passed)} passed / {total} collected "
          f"({len(failed)} failed, {len(skipped)} skipped); the reference itself: all but test_lqi_invalid_q_int_length")
    for name, (cls, msg) in failed.items():
        print(f"[conformance]   FAILED {cls.split('.')[-1]}::{name}: {KNOWN_DEVIATIONS.get(name, 'UNEXPECTED: ' + msg)}")
    for name in skipped:
        print(f"[conformance]   skipped {name}")
    assert total >= 400, f"only {total} reference tests were collected\n{r.stdout[-2000:]}"
    assert not unexpected, f"reference tests failing against the drop-in: {unexpected}\n{r.stdout[-6000:]}"
    assert len(passed) >= total - len(KNOWN_DEVIATIONS)


def test_host_target_motion_draws_the_reference_patterns(gpu_ctx, golden):
    """TargetMotion(params, seed).reset() draws, for every motion type, exactly the pattern parameters the reference
    drew for that seed (fixtures from the reference's env.reset, which seeds its TargetMotion with the same seed), a
    reset without a seed continues the stream, and get_state applies the acceleration clamp (target_motion.py:387-411)."""
    import numpy as np

    from lqr_quadcopter_test_b200 import TargetMotion, TargetParams

    g = golden["reset"]
    for motion in ("stationary", "linear", "circular", "sinusoidal", "figure8"):
        for i in (0, 5, len(g["seeds"]) - 1):
            tm = TargetMotion(TargetParams(motion_type=motion), seed=int(g["seeds"][i]))
            tm.reset()
            np.testing.assert_array_equal(tm._tparams, g[f"{motion}_tparams"][:, i], err_msg=f"{motion} seed {g['seeds'][i]}")
            st = tm.get_state(0.0)
            assert set(st) == {"position", "velocity", "acceleration"} and st["position"].shape == (3,)
            assert np.linalg.norm(st["acceleration"]) <= tm.params.max_acceleration + 1e-9
    for motion in ("linear", "sinusoidal"):
        tm = TargetMotion(TargetParams(motion_type=motion), seed=7)
        for k in range(3):
            tm.reset()                                       # no seed: keeps drawing from the same generator
            np.testing.assert_array_equal(tm._tparams, g[f"{motion}_noseed_tparams"][:, k])
    with pytest.raises(ValueError, match="Invalid motion type"):
        TargetMotion(TargetParams(motion_type="invalid_type")).reset()
