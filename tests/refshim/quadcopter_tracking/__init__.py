"""TEST-ONLY alias: makes `import quadcopter_tracking...` resolve to the B200 drop-in (lqr_quadcopter_test_b200), so the
reference's own test classes (baseline/_ref/reference_tests, see tests/test_reference_suite_gpu.py) exercise the
drop-in through the reference's module paths.  Not product code; only the module paths the hot-path test classes import
exist here."""
__version__ = "0.5.0"
