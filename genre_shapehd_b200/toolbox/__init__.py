"""Drop-in for the reference's ``toolbox`` package (toolbox/__init__.py is empty there).

Makes sure the repository root is importable so the op packages can reach
``genre_shapehd_b200._lib`` (the ctypes binding of the C ABI) however ``toolbox`` got on sys.path.
"""
import os
import sys

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _REPO_ROOT not in sys.path:
    sys.path.append(_REPO_ROOT)
