"""B200-native matching hot path of shoeprint-image-retrieval.

This directory is the implementation behind the reference's import path
``src.shoeprint_image_retrieval`` (``src/shoeprint_image_retrieval/__init__.py`` appends it to
its ``__path__``), so ``run.py`` and every caller of the reference API work unchanged:

* ``similarity.py``   compare_maps / get_similarity / normxcorr  -> CUDA (``libsir.so``)
* ``network.py``      Model.get_feature_maps / get_multiple_feature_maps
* ``parse_results.py`` cmp / cmp_all
* ``config.py``, ``customtypes.py``  schema + loader
* ``engine.py``       device-side orchestration (packing, variants, scoring, ranking, sharding)
* ``_native.py``      ctypes binding of the C ABI declared in ``include/sir.h``
* ``csrc/``           hand-written sm_100a CUDA (tcgen05 / TMA / TMEM) + the C ABI
"""
