"""CPU oracle for the PointConvFormer hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Everything under ``oracle/`` is a CPU restatement (numpy / torch-CPU / plain C) of the reference
algorithm for the hot path named in BASELINE.json, each function citing the reference file:line it
follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or as the reported CPU baseline.
The product package (``ml-pointconvformer_b200``) never imports it and has no CPU fallback.

Pinning status (see DESIGN.md "Oracle"):
  * layers / model (oracle.layers, oracle.model): PINNED -- checked against golden vectors generated
    by importing the unmodified reference (tests/golden/make_golden.py -> tests/golden/*.npz).
  * inverse map (oracle.inverse): PINNED -- against the reference's own `create_inverse_python`
    (cpp_wrappers/cpp_pcf_kernel/test_kernels.py:177-213) executed by make_golden.py.
  * grid subsampling (oracle.grid_subsample): PINNED -- against the reference C++ compiled from its
    own sources into oracle/_ref/ (oracle/Makefile), goldens committed.
  * kNN (oracle.knn): PARITY UNPINNED at the third-party boundary -- the reference's arithmetic lives
    in pykeops / cuVS / sklearn (knn_post_dataloader_utils.py:3-6,10-41), none vendored, none pinned,
    pykeops/cuVS not installed here.  The restatement follows the published KeOps formula
    ((x_i - y_j)**2).sum(-1).argKmin(K) in fp32 with ties broken by lowest index and is
    cross-checked at set level against sklearn.neighbors.KDTree (the reference's third option,
    knn_post_dataloader_utils.py:67-72) on tie-free inputs.
"""
