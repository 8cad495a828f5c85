"""cpu-ray-tracer_b200: B200-native ray core behind the reference's Scene / Renderer surface.

The directory name carries a hyphen (it mirrors the reference's repository name), so import it with
`importlib.import_module("cpu-ray-tracer_b200")` or through the alias module `cpu_ray_tracer_b200`
at the repository root.
"""
from . import abi  # noqa: F401
from .scene_file import FlatScene  # noqa: F401
