"""CPU oracle for the DiffMVS / CasDiffMVS inference hot path.  TEST INFRASTRUCTURE ONLY.

This package is a functional PyTorch restatement (plain fp32 CPU ops, no custom kernels) of the
algorithm in `/root/reference/models/{module,update,diffusion}.py`.  It exists so that the CUDA
product path in `diffmvs_b200/` can be checked on machines where `/root/reference` is absent
(the GPU box).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it; nothing under `diffmvs_b200/` does, and the product path
raises if its CUDA library is missing rather than falling back to this code.

Parity pin: the reference has no tests or golden vectors of its own.  The
oracle is pinned instead against outputs of the reference itself, imported in the build
container by `oracle/make_golden.py` and committed under `tests/golden/`; `tests/test_oracle_golden.py`
replays them.
"""
