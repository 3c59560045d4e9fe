"""CPU oracle for the CMR-Agent geometric hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in here is on the product path: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or
as the reported CPU baseline.  The product (``cmr_agent_b200``) never imports
``oracle`` and raises when its CUDA library is missing.

Contents
--------
shims.py            in-memory stand-ins for the two python modules the
                    reference imports that this image lacks (torch_scatter,
                    open3d) - behaviour restated from SURVEY.md Appendix A.6.
reference_loader.py loads /root/reference/{environment/environment.py,
                    models/pointnet_util.py} by file path.  Works only in the
                    build container (the GPU box has no /root/reference); used
                    by tests/golden/make_golden.py and by the `-m "not gpu"`
                    tests that pin the restatement against the real reference.
env_oracle.py       torch-CPU restatement of environment.py (the "port" that
                    bench.py times as cpu_baseline).
pointnet_oracle.py  torch-CPU / numpy restatement of pointnet_util.py.
cmr_oracle.c        plain-C restatement of the integer-critical arithmetic
                    (projection, FPS, stable kNN, ball query, scatter-mean)
                    used for full-size bit-exact checks in seconds.
cref.py             ctypes binding of the compiled cmr_oracle.c.
dataset_oracle.py   numpy restatement of the datasets' float64 FPS and nearest-node search.
cost_volume_oracle.py  torch-CPU restatement of IterModel's cost-volume warp (pinned by
                    tests/golden/cost_volume.npz, made from the reference's own statements).
tower_oracle.py     the agent's 3-D tower in eval mode (no CUDA path yet: pins the algebra
                    of a later round's kernel against the reference's modules).

Parity pinning: the reference ships no tests or golden vectors,
so the pins are (1) the restatement == the real reference on seeded inputs,
checked here in the build container by tests/test_oracle_vs_reference.py, and
(2) golden fixtures under tests/golden/ generated from the real reference by
tests/golden/make_golden.py, which travel to the GPU box.
"""
