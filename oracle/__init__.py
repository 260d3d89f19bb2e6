"""CPU oracle for the shoeprint matching hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy / scipy.fft, small pure-Python loops) of the
reference algorithm for the path  feature maps -> probe x gallery NCC -> rank -> S-scores.
It exists to CHECK the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
(``src/shoeprint_image_retrieval`` and ``shoeprint-image-retrieval_b200``) never does.

Parity pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself, generated in the build container
by ``tests/golden/make_golden.py`` (imports ``/root/reference`` unmodified) and committed as
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every oracle function here
against those vectors.

Reference locations restated (all relative to the reference repo):
  * ``src/shoeprint_image_retrieval/similarity.py:26-72``   normxcorr         -> ``ncc.py``
  * ``src/shoeprint_image_retrieval/similarity.py:75-108``  get_similarity    -> ``ncc.py``
  * ``src/shoeprint_image_retrieval/similarity.py:230-284`` _apply_transformations -> ``variants.py``
  * ``src/shoeprint_image_retrieval/similarity.py:287-375`` _comparison_worker -> ``compare.py``
  * ``src/shoeprint_image_retrieval/similarity.py:378-386`` _get_rank          -> ``compare.py``
  * ``src/shoeprint_image_retrieval/parse_results.py:4-35`` cmp / cmp_all      -> ``compare.py``
Third-party arithmetic restated from its published algorithm (not vendored in the
reference): scipy ``signal.convolve(mode="same")`` (pinned 1.14.0), Pillow ``Image.rotate``
(nearest, 16.16 fixed point) and ``Image.resize`` (bicubic a=-0.5) (pinned 10.2.0).
"""
