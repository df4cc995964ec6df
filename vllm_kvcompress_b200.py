"""Import alias: the package directory is ``vllm-kvcompress_b200/`` (not a valid Python
identifier), so ``import vllm_kvcompress_b200`` resolves through this one-file shim, which
turns itself into a package whose ``__path__`` is that directory."""
import os as _os

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "vllm-kvcompress_b200")
__path__ = [_dir]
__file__ = _os.path.join(_dir, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _os, _f, _dir
