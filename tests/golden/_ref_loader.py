"""Development-time loader for the upstream reference, used by the fixture generators in this directory.

The loader itself lives in oracle/ref_loader.py (oracle/_ref copy first, /root/reference/src second); the generators run in the
authoring container.  Nothing here is imported by the `-m gpu` tests."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.ref_loader import load_reference  # noqa: E402,F401
