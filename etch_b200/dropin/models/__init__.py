"""Top-level ``models`` package for the reference's callers (src/eval.py:6-15, src/inference_demo.py:5,10).

Put THIS directory's parent (``etch_b200/dropin``) ahead of the reference's ``src/`` on PYTHONPATH and
``from models.models_pointcloud import GT_network_equiv`` / ``from models.fit_SMPL import fit_smpl`` resolve to the
B200 implementation; the scripts themselves stay unchanged.  The modules here only re-export ``etch_b200.models.*``.
"""
import os
import sys

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _REPO not in sys.path:      # make the etch_b200 package importable when only the drop-in directory is on PYTHONPATH
    sys.path.insert(0, _REPO)
