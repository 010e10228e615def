"""Test infrastructure only: CPU restatements of the reference's hot path.

Nothing under arflow_b200/ imports this package; tests/, __graft_entry__.smoke() and bench.py's CPU baseline do.
"""
