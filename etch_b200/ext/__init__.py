"""Drop-in modules carrying the names of the reference's native extensions.

Add this directory to ``sys.path`` (``etch_b200.ext.install()``) and the unmodified reference Python
(``vgtk.pc``, ``vgtk.so3conv.functional``, ``src/models/pointops.py``) binds to the B200 kernels.
"""
import os
import sys


def install():
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
