"""minispark_b200 -- a B200-native execution engine behind minispark's DataFrame / SQL surface.

The hot path (BlockFile ingest -> scan -> filter -> project -> GROUP BY -> shuffle -> join) runs as
hand-written sm_100a CUDA kernels in ``lib/libminispark_cuda.so`` (sources in ``csrc/``, C-ABI in
``include/minispark_cuda.h``).  The Python modules mirror the reference's public interface
(``DataFrame``, ``Col``, ``Functions``, ``BlockFile``, ``ExecutionEngine``) so queries read unchanged.
"""

from .constants import ColumnType
from .dataframe import DataFrame
from .execution import CudaExecutionEngine, ExecutionEngine, ExecutionError
from .io import BlockFile
from .jobs import JobResult, OutputFile
from .sql import AggCol, Col, Functions, Lit
from .utils import TRACER, trace

__all__ = [
    "AggCol", "BlockFile", "Col", "ColumnType", "CudaExecutionEngine", "DataFrame", "ExecutionEngine",
    "ExecutionError", "Functions", "JobResult", "Lit", "OutputFile", "TRACER", "trace",
]
