"""Importable alias of the `voicebox-pytorch_b200/` package directory (a hyphen is not a valid Python identifier).
`import voicebox_pytorch_b200` executes `voicebox-pytorch_b200/__init__.py` with this module's name and search path."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'voicebox-pytorch_b200')
__path__ = [_real]
with open(_os.path.join(_real, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))
del _os, _f, _real
