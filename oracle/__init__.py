"""CPU oracle for the probayes hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import anything from here.  The
shipped package ``probayes_b200`` never does: it fails loudly when its CUDA
library is missing instead of falling back to this code.

Contents
--------
``np_oracle``   numpy restatement of the reference's algorithm for the path
                (MH step loop, normal / mvn densities, pscales clamps, box
                priors, ufun proposals, DGEI product + marginal algebra,
                CondCov Gibbs).  Every function cites the reference file:line it
                follows (paths are relative to the reference checkout root).
``philox``      numpy Philox4x32-10 + the u01 / Box-Muller conventions the CUDA
                kernels use, so native-RNG device runs can be replayed on the CPU
                through the *injected-stream* restatement.
``c/``          the same algorithms in plain C + OpenMP (liboracle), used as the
                all-host-cores CPU baseline and as a second opinion on numpy.
``ref_shim``    import shim that loads the *live* reference from /root/reference
                (development container only; absent on the GPU box).
``gen_golden``  drives the live reference through its public API on seeded
                inputs with injected proposal/threshold streams and writes the
                small fixtures under ``tests/golden/``.

Parity status: the reference's own tests hold no vectors for this path
(SURVEY.md section 8c), so the oracle is pinned against outputs of the
reference itself, generated here by ``gen_golden.py`` and committed under
``tests/golden/``; ``tests/test_oracle_golden.py`` replays them.
"""
