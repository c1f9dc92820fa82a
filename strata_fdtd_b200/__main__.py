"""python -m strata_fdtd_b200 script.py [args...]  -- run a strata_fdtd script on the b200 backend.

On N GPUs:  python -m torch.distributed.run --nproc-per-node N -m strata_fdtd_b200 script.py
Every rank executes the script; ``FDTDSolver(...)`` then returns the slab solver (one slab per GPU) behind the same
methods.  Output of the ranks other than 0 is discarded unless STRATA_B200_ALL_RANKS_PRINT=1.
"""
import os
import sys

from .compat import run_script

if __name__ == "__main__":
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    if int(os.environ.get("RANK", "0")) != 0 and os.environ.get("STRATA_B200_ALL_RANKS_PRINT", "0") != "1":
        sys.stdout = open(os.devnull, "w")
    run_script(sys.argv[1], sys.argv[2:])
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()
    except ImportError:
        pass
