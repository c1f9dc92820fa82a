"""pytest plugin (``-p ref_alias_plugin``): makes ``import strata_fdtd`` resolve to this package's mirrors with
``backend="b200"`` as the default, and ``h5py`` to the in-memory stand-in, before the reference's OWN test files are
collected (tests/test_reference_own_tests.py runs them, unmodified, from oracle/_ref/tests/)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for q in (str(ROOT), str(ROOT / "tests")):
    if q not in sys.path:
        sys.path.insert(0, q)

import fake_h5py  # noqa: E402

sys.modules.setdefault("h5py", fake_h5py)

from strata_fdtd_b200.compat import install_as_strata_fdtd  # noqa: E402

install_as_strata_fdtd(force_alias=True)
