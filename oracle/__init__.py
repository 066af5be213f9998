"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the GNNome GatedGCN + edge-score hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the timed CPU baseline.

Contents
--------
* ``dgl_shim/``        a ~150-line stand-in for the slice of ``dgl==0.8.1`` that the
                       reference's hot path touches, so that the reference's OWN unmodified
                       ``layers/*.py`` + ``models/full_graph.py`` execute on CPU torch.
* ``reference_runner`` imports those files from ``/root/reference`` (build container only).
* ``restatement``      an independent straight-line torch restatement of the same algorithm
                       (fp32 or fp64), each function citing the reference file:line.  This is
                       what travels to the GPU box (``/root/reference`` does not exist there).
* ``make_golden``      the script that produced ``tests/golden/*.pt`` from the reference.

Parity status: the reference has no tests / golden vectors ("parity unpinned" at the DGL
boundary).  The oracle is pinned the only way available: the restatement
is checked against the reference's own files run here over the shim, and the fixtures in
``tests/golden`` were produced by that run with the shipped ``weights/weights.pt``.
"""
