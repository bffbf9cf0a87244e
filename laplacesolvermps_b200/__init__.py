"""B200-native tensor-train kernels behind the API of MScherbela/LaplaceSolverMPS (`laplace_mps`).

Modules mirror the reference package: `tensormethods` (TensorTrain), `solver`, `bpx2D`, `utils`;
`batched` holds the device-resident batched layer, `_lib` the ctypes binding of csrc/libqtt_b200.so.
"""
__all__ = ["tensormethods", "solver", "bpx2D", "utils", "batched"]
