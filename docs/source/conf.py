"""Sphinx configuration (counterpart of the reference's doc/source/conf.py).

``sphinx-build -b html docs/source docs/_build`` renders the autodoc pages below; the compiled
extension and pyspark are mocked so that the pages build on a machine with neither (the reference
mocks ``pyspark`` and ``tensorflow`` the same way, doc/source/conf.py:22).  ``docs/api.md`` is the
same content rendered without sphinx (``python tools/gen_api_docs.py``).
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))

project = "tensorflowonspark_b200"
author = "tensorflowonspark_b200 developers"
release = "0.2.0"

extensions = ["sphinx.ext.autodoc", "sphinx.ext.napoleon", "sphinx.ext.viewcode"]
autodoc_mock_imports = ["pyspark", "tensorflowonspark_b200._ext"]
autodoc_default_options = {"members": True, "undoc-members": True, "show-inheritance": True}
templates_path = ["_templates"]
exclude_patterns = []
html_theme = "alabaster"
master_doc = "index"
