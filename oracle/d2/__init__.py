"""Stand-in for the parts of Detectron2 the reference path imports (test infrastructure only; see oracle/__init__.py)."""
