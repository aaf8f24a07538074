"""One-off offline install of the unmodified reference into baseline/_ref (build container only):

    python -m baseline.install_ref [--force]

= `pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy>`
run on a copy of /root/reference under /tmp (the source tree is read-only and versioneer writes into it).
`--no-deps`: nibabel / seaborn / nilearn / matplotlib are not in the wheelhouse and not on the resampling path."""
import os
import shutil
import subprocess
import sys
import tempfile

from .refshim import REF_DIR, reference_available

SOURCE = os.environ.get("PLSPY_REFERENCE", "/root/reference")


def install(force=False):
    if reference_available() and not force:
        return REF_DIR
    if not os.path.isdir(SOURCE):
        raise FileNotFoundError(f"{SOURCE} not present (the GPU box only uses the prebuilt baseline/_ref)")
    tmp = tempfile.mkdtemp(prefix="plspy_ref_")
    try:
        src = os.path.join(tmp, "src")
        shutil.copytree(SOURCE, src, ignore=shutil.ignore_patterns(".git", "docs", ".tox"))
        if os.path.isdir(REF_DIR):
            shutil.rmtree(REF_DIR)
        subprocess.check_call([sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation",
                               "--no-deps", "--find-links", "/opt/wheelhouse", "--target", REF_DIR, src])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    if not reference_available():
        raise RuntimeError("pip install finished but baseline/_ref/plspy is missing")
    return REF_DIR


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
