"""mixed_graph_admm_b200 — B200-native solver hot path of Mixed-Graph-ADMM.

``from mixed_graph_admm_b200.ADMM import ADMM_algorithm`` is the drop-in for the reference's
``from ADMM import ADMM_algorithm``; ``mixed_graph_admm_b200.utils`` mirrors its ``utils.py`` graph
construction.  The kernels live in ``csrc/`` and are reached only through the C ABI of
``include/mga.h`` (``_lib/libmga.so``, built by ``python -m mixed_graph_admm_b200.build``).
"""
__version__ = "0.1.0"
