"""Import shim: the product package lives in the directory ``kinematics.jl_b200/`` (named after the
reference repo), which Python cannot import by name because of the dot.  This package points its
``__path__`` there, so ``import kinematics_jl_b200 as K`` / ``from kinematics_jl_b200 import lib`` work."""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "kinematics.jl_b200"))

from .api import *  # noqa: E402,F401,F403
from .api import __all__  # noqa: E402,F401
