"""Run one of the reference's own driver scripts, unmodified, on the B200 engine.

    python -m cybayes_b200.run_reference_driver /path/to/CyBayes/mat_mcmc_gamma.py -i ... -m F81 -n 100000 -t 1000 -d bin -o out
    python -m cybayes_b200.run_reference_driver /path/to/CyBayes/mat_mcmc.py      -i ... -m JC  -n 1000   -t 10   -d multi -o out

The script is executed with ``runpy`` from outside its own directory, with ``cybayes_b200/compat`` first
on ``sys.path``: its ``import utils, config`` / ``from mcmc_gamma import *`` / ``from ML_gamma import *``
(mat_mcmc_gamma.py:2-5) resolve to this package.  (Starting ``mat_mcmc.py`` with plain ``python`` would put
the reference directory first on the path, where a dead legacy ``mcmc.py`` shadows the alias)
Byte-compiled copies (oracle/_ref/*.code) work too.
"""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit(__doc__)
    script, args = argv[0], argv[1:]
    compat = os.path.join(os.path.dirname(os.path.abspath(__file__)), "compat")
    script_dir = os.path.dirname(os.path.abspath(script))
    sys.path[:] = [compat] + [p for p in sys.path if os.path.abspath(p or ".") != script_dir]
    for name in ("config", "utils", "mcmc_gamma", "ML_gamma", "mcmc", "ML"):
        sys.modules.pop(name, None)
    sys.argv = [script] + args
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
