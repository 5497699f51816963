#!/usr/bin/env python
"""minisat-b200: the reference's CLI surface over the B200 engine (see minisat-rust_b200/cli.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _b200pkg  # noqa: E402

pkg = _b200pkg.load()
from minisat_rust_b200 import cli  # noqa: E402

if __name__ == "__main__":
    sys.exit(cli.main())
