"""Data-parallel plumbing: symmetric (peer-mapped) memory, the fused
all-reduce + optimizer step, parameter-server push/pull, process-group setup."""
