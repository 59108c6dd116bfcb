"""dbeel_b200 -- B200-native LSM compaction engine behind dbeel's storage-engine boundary.

The product is the C-ABI shared library built from csrc/ (libdbeel_compact.so, sm_100a only, no CPU fallback);
this package holds its ctypes bindings (capi, storage_engine), the SSTable byte formats (sstable), the synthetic
workloads of BASELINE.json (workloads, cfg5) and the multi-GPU job hand-off (shard_jobs).  Nothing here imports
the CPU oracle: that is test infrastructure (oracle/)."""

__all__ = ["capi", "cfg5", "shard_jobs", "sstable", "storage_engine", "workloads"]
