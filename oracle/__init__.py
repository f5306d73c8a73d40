"""CPU oracle for the UniT RoI stage -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``unit_b200`` never does: the
product path fails loudly when its CUDA library is missing.

Contents
  oracle.d2        restatement of the Detectron2 glue the reference calls (Detectron2 is an
                   un-vendored dependency of the reference, INSTALL.md:5 "detectron2 >= 0.2.1";
                   restated from the public v0.3/v0.4 source; ROIAlign / NMS are the compiled
                   torchvision 0.26 CPU ops, i.e. the kernels a current Detectron2 dispatches to)
  oracle.unit_ref  restatement of the UniT side (modeling/roi_heads/*.py, modeling/matcher.py)
  oracle.shim      installs ``detectron2`` / ``fvcore`` stand-in modules built from oracle.d2 so the
                   reference's own files can be imported verbatim from /root/reference (only in the
                   build container; used to pin oracle.unit_ref and to generate tests/golden/)
  oracle/c         plain-C restatement of the integer/byte-exact kernels (IoU, matcher, NMS, ROIAlign)

Parity pinning: the reference ships no tests and no golden vectors.  The oracle is
pinned by (a) the reference's own files executed verbatim through oracle.shim in the build container,
with their outputs committed as fixtures under tests/golden/ (script: tests/golden/make_golden.py),
(b) the shipped ``data/embeddings/glove_mean`` known-answer values, and (c) the torchvision CPU ops.
Detectron2 itself could not be installed (no wheel/sdist, no network), so the D2 glue is a restatement:
that part of parity is "restated, checked against torchvision + the verbatim UniT files", not pinned by
an upstream Detectron2 run.
"""
