"""Import shim: the package lives in the directory ``di-engine_b200/`` (named after the reference repository, which
is not a valid python identifier).  ``import di_engine_b200`` loads that directory as the package ``di_engine_b200``.
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'di-engine_b200')
_spec = importlib.util.spec_from_file_location(
    'di_engine_b200', os.path.join(_dir, '__init__.py'), submodule_search_locations=[_dir]
)
_pkg = importlib.util.module_from_spec(_spec)
sys.modules['di_engine_b200'] = _pkg
_spec.loader.exec_module(_pkg)
