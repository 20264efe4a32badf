"""pherofield — B200-native pheromone field and swarm step.

    abi          ctypes mirror of include/pherofield.h (loads nothing)
    field        PheroField / Swarm: the dense double-buffered field and the hot-path calls
    compat       ReferenceCompat: the reference's five methods + tick drivers over the dense field
    trail        TrailSwarm: the reference's sparse timestamped trail for N robots
    heatmap      HeatMap: graph.py's deposit + Gaussian loop
    sharded      row-strip driver over torch.distributed (NCCL) / in-process strips
    logs         heatmap.txt writer/reader in the reference's format

Everything that computes goes through libpherofield.so (sm_100a); there is no CPU fallback.
"""
__version__ = "0.1.0"
