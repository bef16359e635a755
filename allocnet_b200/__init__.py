"""B200-native batched MINCO trajectory optimizer (AllocNet planner back-end slot).

Only what the hot path needs: csrc/ (sm_100a kernels + C-ABI), api.py (host binding mirroring
the MINCO / lbfgs interface), params.py (mincob_params), synth.py (seeded corridor problems),
build.py (in-tree nvcc build)."""
