"""The unmodified reference (McIntosh-Lab/plspy), installed once into `baseline/_ref` (git-ignored, ships to the GPU
box with the gpurun snapshot) by `python -m baseline.install_ref`.  Test / benchmark infrastructure only: it is the
CPU arm of `bench.py --impl reference`, the `cpu_baseline` leg, and the host that `plspy_b200.install()` is tested
against.  Nothing under `plspy_b200/` imports it."""
from .refshim import REF_DIR, import_reference, reference_available  # noqa: F401
