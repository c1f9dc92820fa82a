"""pytest plugin (``-p ref_native_plugin``): ``import strata_fdtd`` resolves to the UNMODIFIED reference package with its
C++ kernels compiled by oracle/build_ref (oracle/ref_loader.py; h5py, which this image lacks, is the in-memory stand-in).
Used to run the reference's own tests against the build the oracle and the CPU baseline rely on."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for q in (str(ROOT), str(ROOT / "tests")):
    if q not in sys.path:
        sys.path.insert(0, q)

import fake_h5py  # noqa: E402

sys.modules.setdefault("h5py", fake_h5py)

from oracle import ref_loader  # noqa: E402

ref_loader.load_reference_package()
