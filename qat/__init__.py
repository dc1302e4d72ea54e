"""QAT workloads of BASELINE.json (configs 1, 4, 5) built from brevitas_b200.nn layers, and the data-parallel
training-step harness (one process per GPU, NCCL gradient all-reduce).  Workloads, not product code."""
