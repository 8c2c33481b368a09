"""Import shim: `import sg3_b200` loads the package in ./stylegan3-editing_b200/ (whose directory
name, fixed by the project layout, contains a hyphen and so is not a valid module name)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'stylegan3-editing_b200')
_spec = importlib.util.spec_from_file_location('sg3_b200', os.path.join(_dir, '__init__.py'),
                                               submodule_search_locations=[_dir])
_pkg = importlib.util.module_from_spec(_spec)
sys.modules['sg3_b200'] = _pkg
_spec.loader.exec_module(_pkg)
