"""N-split (column-sharded) execution of the mmq path across the GPUs of one node."""
