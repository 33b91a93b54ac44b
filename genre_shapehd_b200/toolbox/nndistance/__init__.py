"""Drop-in for the reference's ``nndistance`` package (toolbox/nndistance/).  The reference imports it
as a TOP-LEVEL package (``from nndistance.functions.nnd import nndistance``, functions/nnd.py:5,
modules/nnd.py:2); genre_shapehd_b200.install() puts this directory's parent on sys.path for that."""
import os
import sys

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _REPO_ROOT not in sys.path:
    sys.path.append(_REPO_ROOT)
