"""``torch_cfd.tensor_utils``: imported by the reference scripts (src/inference.py:11) and never used on the rollout
path; present so that the import line resolves."""
