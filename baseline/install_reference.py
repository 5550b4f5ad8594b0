"""Put the UNMODIFIED reference where the GPU box can import it: baseline/_ref/ (git-ignored, shipped by gpurun).

    python baseline/install_reference.py            # run in the build container, where /root/reference exists

1. `pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>`.
   In this image that fails: the reference's build backend (hatchling, pyproject.toml) is not installed and there is
   no network.  The reference is a pure-Python package, so the fallback installs exactly what its wheel would
   contain: the package directory src/quadcopter_tracking, byte for byte.
2. The reference's own tests (tests/*.py) go to baseline/_ref/reference_tests/ — the conformance suite that
   tests/test_reference_suite_gpu.py runs against the drop-in classes.
3. matplotlib is absent from the image and quadcopter_tracking.eval imports it at module level (eval.py:26-30): a
   two-file stub (the one SURVEY.md Appendix C describes) is written to baseline/_ref/_stubs/matplotlib/ and is only
   put on sys.path when the real package is missing.

Nothing here is product code and nothing under baseline/_ref/ is tracked: bench.py (cpu_baseline.python_reference)
and the conformance test read it, both as checkers / baselines only.
"""
from __future__ import annotations

import pathlib
import shutil
import subprocess
import sys
import tempfile

HERE = pathlib.Path(__file__).resolve().parent
REF = pathlib.Path("/root/reference")
DEST = HERE / "_ref"


def install(verbose: bool = True) -> str | None:
    """Returns how the package was installed ("pip" / "copy"), or None if /root/reference is absent."""
    if not (REF / "src" / "quadcopter_tracking").is_dir():
        return None
    if DEST.exists():
        shutil.rmtree(DEST)
    DEST.mkdir(parents=True)
    how = "copy"
    with tempfile.TemporaryDirectory() as tmp:
        src = pathlib.Path(tmp) / "reference"
        shutil.copytree(REF, src, ignore=shutil.ignore_patterns(".git", "__pycache__"))   # /root/reference is read-only
        r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                            "--find-links", "/opt/wheelhouse", "--target", str(DEST), str(src)],
                           capture_output=True, text=True)
        if r.returncode == 0 and (DEST / "quadcopter_tracking").is_dir():
            how = "pip"
        else:
            if verbose:
                tail = (r.stderr or r.stdout).strip().splitlines()[-1:]
                print(f"pip install failed ({' '.join(tail)}); copying the pure-Python package instead", file=sys.stderr)
            shutil.copytree(src / "src" / "quadcopter_tracking", DEST / "quadcopter_tracking",
                            ignore=shutil.ignore_patterns("__pycache__"))
        shutil.copytree(src / "tests", DEST / "reference_tests", ignore=shutil.ignore_patterns("__pycache__"))
        for extra in ("experiments",):           # the tuning / evaluation presets some reference tests read
            if (src / extra).is_dir():
                shutil.copytree(src / extra, DEST / extra)
    stub = DEST / "_stubs" / "matplotlib"
    stub.mkdir(parents=True)
    (stub / "__init__.py").write_text("def use(*a, **k):\n    pass\n")
    (stub / "pyplot.py").write_text("class Figure:\n    pass\n\n\nclass Axes:\n    pass\n\n\ndef close(*a, **k):\n    pass\n")
    (DEST / "INSTALLED.txt").write_text(f"quadcopter_tracking installed by baseline/install_reference.py via: {how}\n")
    return how


def reference_paths() -> list[str]:
    """sys.path entries that make `import quadcopter_tracking` resolve to the reference (empty if not installed)."""
    if not (DEST / "quadcopter_tracking").is_dir():
        return []
    paths = [str(DEST)]
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        paths.append(str(DEST / "_stubs"))
    return paths


if __name__ == "__main__":
    print(install())
