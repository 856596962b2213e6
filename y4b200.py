"""Import alias: the package directory is named after the reference ("yolo-v4-tf.keras_b200"), which is not a
valid Python identifier; `import y4b200` loads it under this name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'yolo-v4-tf.keras_b200')
_spec = importlib.util.spec_from_file_location('y4b200', os.path.join(_dir, '__init__.py'),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules['y4b200'] = _mod
_spec.loader.exec_module(_mod)
