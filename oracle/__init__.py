"""TEST INFRASTRUCTURE ONLY — CPU oracle for the STonKGs hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it,
and only as the checker.  The product (``stonkgs_b200``) never imports this
package and fails loudly when its CUDA library is missing.

Contents
--------
* ``weights.py``        deterministic synthetic checkpoint (seeded, reproducible on any box
                        with the same torch build).
* ``stonkgs_oracle.py`` fp32 CPU restatement of ``STonKGsForPreTraining.forward``
                        (reference ``src/stonkgs/models/stonkgs_model.py:62-73,123-141,149-258``
                        plus the HuggingFace BERT arithmetic it delegates to).
* ``ref_shim.py``       side-effect-free import of the *real* reference module from
                        ``/root/reference`` (dev container only) used to pin the restatement and
                        to generate ``tests/golden/*.npz``.
* ``make_golden.py``    the script that generated the committed golden vectors.

Parity status: the reference has no tests on this path (``tests/test_version.py`` only), so the
restatement is pinned against outputs of the reference module itself, executed in the dev
container through ``ref_shim`` (``tests/golden/*.npz``; generator committed).
"""
