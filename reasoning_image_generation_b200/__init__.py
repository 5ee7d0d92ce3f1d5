"""B200-native render + near-duplicate stages for the rule-based puzzle generator
MrGu11/reasoning_image_generation (drop-in for RPMGenerator._render_elements_to_canvas).

Host side (this package): scene emission (scene.py), ctypes binding of the C ABI (_abi.py, render.py).
Device side: csrc/ -> librig.so (hand-written sm_100a kernels).  No CPU rendering path exists here.
"""
from .scene import KINDS, Scene, SceneBuilder  # noqa: F401
from .render import Renderer, RigError, install, render_elements_to_canvas  # noqa: F401
