"""Import shim: the package directory is named `gpu-raytracer_b200` (not a valid Python identifier);
`import gpu_raytracer_b200` resolves to it."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "gpu-raytracer_b200")]
__package__ = __name__
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
__file__ = _os.path.join(__path__[0], "__init__.py")
exec(compile(open(__file__).read(), __file__, "exec"))
