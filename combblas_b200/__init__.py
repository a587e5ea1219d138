"""combblas_b200 -- B200-native semiring SpGEMM hot path behind the CombBLAS SpDCCols / semiring interface.

The product is ``libcbgpu.so`` (hand-written sm_100a CUDA behind the C ABI in ``include/cbgpu.h``).
This package is the thin Python host mirror used by the tests and the benchmark: it names things the way the
reference does (``SpDCCols``, ``LocalHybridSpGEMM``, ``MultiwayMerge``, ``Mult_AnXBn_Synch``,
``Mult_AnXBn_SUMMA3D``, ``PlusTimesSRing`` ...) and forwards every call through the C ABI.

There is NO CPU fallback: without the CUDA library or without a GPU every compute entry point raises.
"""
from .lib import (  # noqa: F401
    CbgpuError,
    Context,
    DeviceMatrix,
    Stats,
    DistStats,
    Grid,
    Comm,
    SlabPipeline,
    PruneStats,
    MemEffStats,
    lib_path,
    load_library,
    F64, F32, I64, I32, BOOL,
    DTYPE_TO_NUMPY,
)
from .host import (  # noqa: F401
    SpDCCols,
    SpTuples,
    SEMIRINGS,
    PlusTimesSRing_f64, PlusTimesSRing_f32, PlusTimesSRing_i64, SelectMaxSRing_bool_i64, MinPlusSRing_f64,
    OrAndSRing_bool, PlusTimesSRing_bool_f64, PlusTimesSRing_i32, SelectMaxSRing_i64,
    BoolCopy2ndSRing_f64, BoolCopy1stSRing_f64, BoolCopy2ndSRing_i64, BoolCopy1stSRing_i64, BoolCopy2ndSRing_bool, BoolCopy1stSRing_bool,
    semiring_types, load_user_semiring,
    LocalHybridSpGEMM, LocalSpGEMMHash, LocalSpGEMM, MultiwayMerge, MultiwayMergeHash, EstimateFLOP,
    MCLPruneRecoverySelect, MemEfficientSpGEMM, CalculateNumberOfPhases,
    block_range, block_owner, partition_2d, partition_3d,
)

__all__ = [n for n in dir() if not n.startswith("_")]
