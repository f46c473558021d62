import os as _os

from .. import _extend_with_reference

__path__ = [_os.path.dirname(_os.path.abspath(__file__))]
_extend_with_reference(__path__, ("models",))
