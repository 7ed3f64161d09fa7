"""python -m opendpd_b200.run_reference /path/to/OpenDPD [main.py arguments...]

Runs the UNMODIFIED reference entry point (main.py:13-37) with the native backbones: puts `<repo>/shim` in front of the reference
root on sys.path, so `import models` / `from quant import get_quant_model` in steps/*.py resolve to the native drop-ins, then
executes the reference's main.py as __main__ from the current directory (the reference writes ./save ./log ./dpd_out)."""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or not os.path.isfile(os.path.join(argv[0], "main.py")):
        raise SystemExit("usage: python -m opendpd_b200.run_reference /path/to/OpenDPD [--step train_pa ...]")
    ref = os.path.abspath(argv[0])
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shim = os.path.join(repo, "shim")
    for p in (ref, repo, shim):                       # final order: shim, repo, reference root, ...
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    sys.argv = [os.path.join(ref, "main.py")] + argv[1:]
    # run_path would put the script directory first again; execute the file's code ourselves with the path we built
    code = compile(open(sys.argv[0]).read(), sys.argv[0], "exec")
    exec(code, {"__name__": "__main__", "__file__": sys.argv[0]})


if __name__ == "__main__":
    main()
