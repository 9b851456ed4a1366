"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the PARADIS semi-Lagrangian advection hot path
(reference: model/advection.py:129-169, model/padding.py:11-39).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package.  The product
(``paradis_model_b200``) never imports it and has no CPU fallback.
"""
