"""machline_b200: B200-native (sm_100a) replacement for MachLine's two hot paths.

* ``machline_b200.host``   -- host-side case setup / post-processing (libmachline_host.so)
* ``machline_b200.gpu``    -- AIC assembly and dense solve (libmachline_gpu.so, CUDA only)
* ``machline_b200.solver`` -- ``run_case``: the end-to-end call a MachLine user makes
* ``machline_b200.meshgen``-- deterministic synthetic meshes for benchmarks
"""
__version__ = "0.1.0"
