#!/usr/bin/env python
"""Install the UNMODIFIED reference into ``baseline/_ref`` (git-ignored, travels to the GPU box with gpurun).

    python baseline/install_ref.py [--reference /root/reference]

Step 1 is the contract's recipe: ``pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target
baseline/_ref <reference>``.  The reference is a flat script collection without ``setup.py`` / ``pyproject.toml``
(SURVEY.md §0), so pip refuses it; the outcome is recorded in ``baseline/_ref/INSTALL_LOG.txt`` and DESIGN.md.
Step 2 (what actually installs it): the reference's files are placed under ``baseline/_ref`` unchanged --
``results/`` (sample audio, 2.4 MB) is skipped, everything else (7.5 MB incl. the two bundled SpeechSR checkpoints and
``example/reference_1.wav``) is kept so that ``oracle/refload.py`` can import the reference's own modules on the GPU
box.  Nothing under ``baseline/_ref`` is tracked by git, nothing in the product path imports it: it serves the
reference arm of ``bench.py`` and the reference-vs-B200 GPU tests."""
import argparse
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")
SKIP = {"results", ".git", "__pycache__"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("HSV_REFERENCE_ROOT", "/root/reference"))
    args = ap.parse_args()
    src = args.reference
    if not os.path.isfile(os.path.join(src, "hierspeechpp_speechsynthesizer.py")):
        print(f"reference not found under {src}: nothing installed")
        return 1
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    log = []
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links",
           "/opt/wheelhouse", "--target", DST, src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    tail = (r.stdout + r.stderr).strip().splitlines()[-3:]
    log.append("$ " + " ".join(cmd))
    log.append(f"rc={r.returncode}: " + " | ".join(tail))
    pip_ok = r.returncode == 0 and os.path.isfile(os.path.join(DST, "hierspeechpp_speechsynthesizer.py"))
    if not pip_ok:
        log.append("pip cannot install the reference (no setup.py / pyproject.toml): placing its files unchanged")
        for name in sorted(os.listdir(src)):
            if name in SKIP:
                continue
            s, d = os.path.join(src, name), os.path.join(DST, name)
            if os.path.isdir(s):
                shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
            else:
                shutil.copy2(s, d)
    n = sum(len(f) for _, _, f in os.walk(DST))
    log.append(f"{n} files under baseline/_ref")
    with open(os.path.join(DST, "INSTALL_LOG.txt"), "w") as f:
        f.write("\n".join(log) + "\n")
    print("\n".join(log))
    return 0


if __name__ == "__main__":
    sys.exit(main())
