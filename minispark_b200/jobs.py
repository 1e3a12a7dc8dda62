"""Result descriptors returned across the engine boundary.

``ExecutionEngine.execute_full_task`` returns ``list[JobResult]`` whose ``OutputFile`` paths are
BlockFiles read back by ``collect_results`` (reference ``src/mini_spark/jobs.py:27-37``,
``execution.py:41-55``).  The reference's ``ScanJob``/``LoadShuffleFilesJob``/``JoinJob`` byte
encodings (``jobs.py:45-79``) feed its subprocess workers over stdin and have no GPU analogue:
row-blocks are sharded over ranks and handed to kernels in-process instead.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path


@dataclass(frozen=True)
class OutputFile:
    file_path: Path
    partition: int = 0


@dataclass
class JobResult:
    job_id: str
    executor_id: str
    output_files: list[OutputFile] = field(default_factory=list)
    # CudaExecutionEngine with several ranks and replicate_results=False: True = the files hold this rank's part of the
    # result only (the union over ranks is the result); always False otherwise, as in the reference (jobs.py:27-31)
    result_partitioned: bool = False
