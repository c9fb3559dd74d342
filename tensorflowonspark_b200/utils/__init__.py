"""Checkpointing, metrics, profiling hooks and fault injection."""
