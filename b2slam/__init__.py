"""Import alias for the product package.

The package directory is named after the reference repo
(`a-2d-lidar-based-slam-system-for-wheeled-mobile-robots_b200/`), which is not a valid Python
identifier; `import b2slam` resolves every submodule from that directory.
"""
import os as _os

_impl = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "a-2d-lidar-based-slam-system-for-wheeled-mobile-robots_b200")
__path__.insert(0, _impl)

with open(_os.path.join(_impl, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_impl, "__init__.py"), "exec"))
del _fh
