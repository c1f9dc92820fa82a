"""python -m strata_fdtd_b200 script.py [args...]  -- run a strata_fdtd script on the b200 backend."""
import sys

from .compat import run_script

if __name__ == "__main__":
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    run_script(sys.argv[1], sys.argv[2:])
