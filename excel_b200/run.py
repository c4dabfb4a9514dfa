"""``python -m excel_b200.run <reference script> [args...]`` -- run an unmodified ExCEL script on the sm_100a path.
Must be started from the reference root (the scripts use ./-relative paths and ``sys.path.append("./")``)."""
import os
import runpy
import sys


def main():
    if len(sys.argv) < 2:
        sys.exit("usage: python -m excel_b200.run <script.py> [args...]")
    script = sys.argv[1]
    sys.argv = sys.argv[1:]
    sys.path.insert(0, os.getcwd())
    from .install import install
    install()
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
